// sched_selftest.cu -- host-only check of the scheduler's dependency logic (no GPU needed).
//
// The scheduler (sched.h) reorders, batches, fuses and drops recorded calls; what it must never
// do is change what the program computes.  Here a pseudo-random program over the whole recorded
// API (limb add/sub/mul, the tmp-then-accumulate pattern that triggers mul+add fusion, zero, copy,
// fill, gather, encode, ModUp, ModDown, Rescale, alloc, free) is run twice through a Scheduler
// whose backend interprets the batches on HOST arrays with toy arithmetic (8 coefficients per
// limb, 16-bit primes): once call by call (eager) and once deferred.  At every synchronisation
// point the contents of all live blocks must agree.  The toy backend processes the chains, gathers
// and jobs of a batch in REVERSE order, so two ops that the scheduler wrongly put into one wave
// show up as a mismatch.  Built into libace_b200_selftest.so only (build.py), never into the product library; used by
// tests/test_cpu_sched.py.  TEST INFRASTRUCTURE: nothing on the product path calls this.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "sched.h"

namespace ace {
namespace {

constexpr u32 kN = 8, kK = 2, kPart = 3, kMods = 16;
const u64 kPrime[kMods] = {65521, 65519, 65497, 65479, 65449, 65447, 65437, 65423,
                           65419, 65413, 65407, 65393, 65381, 65371, 65357, 65353};

struct HostSimBackend : SchedBackend {
  std::map<const u64*, size_t> blocks;
  uint64_t                     alloc_counter = 0;
  ~HostSimBackend() override {
    for (auto& kv : blocks) delete[] kv.first;
  }
  u32 N() const override { return kN; }
  u32 K() const override { return kK; }
  u32 digit_start(u32 part) const override { return kPart * part; }
  u32 digit_len(u32 num_q, u32 part) const override {
    u32 beta = (num_q + kPart - 1) / kPart, st = kPart * part;
    return part == beta - 1 ? num_q - st : kPart;
  }
  // set by the program before each of ITS allocations; the scheduler's own buffers (kept ModUp
  // results, only in the deferred run) must not shift the pattern of the program's blocks
  int64_t next_tag = -1;
  u64* alloc(size_t n) override {
    u64* p = new u64[n * kN];
    // "uninitialised" memory with reproducible contents: the program's blocks get the same
    // pattern in both runs
    const uint64_t tag = next_tag >= 0 ? (uint64_t)next_tag : 0xABCDEFull;
    next_tag = -1;
    for (size_t i = 0; i < n * kN; i++) p[i] = (tag * 977 + i * 31 + 5) % 1000;
    alloc_counter++;
    blocks[p] = n;
    return p;
  }
  void free(u64* p) override {
    blocks.erase(p);
    delete[] p;
  }
  size_t block_limbs(const u64* p) const override { return blocks.at(p); }
  void   count_limb_op(int) override {}

  static u64 q_of(u32 g) { return kPrime[g % kMods]; }

  void run_chains(const ChainPack& P, u32 n_chains) override {
    for (int ch = (int)n_chains - 1; ch >= 0; ch--) {  // chains are independent: any order
      for (u32 i = 0; i < kN; i++) {                   // one "thread" per coefficient
        for (u32 k = P.chain_start[ch]; k < P.chain_start[ch + 1]; k++) {
          const ChainItem& it = P.it[k];
          const u64 q = q_of(it.g);
          if (it.op == OP_ZERO) { it.r[i] = 0; continue; }
          if (it.op == OP_COPY) { it.r[i] = it.a[i]; continue; }
          if (it.op == OP_FILL) { it.r[i] = (u64)(uintptr_t)it.a; continue; }
          const u64 x = it.a[i] % q, y = it.b[i] % q;
          u64 z;
          if (it.op == OP_ADD) z = (x + y) % q;
          else if (it.op == OP_SUB) z = (x + q - y) % q;
          else {
            z = (x * y) % q;
            if (it.op == OP_MAC) {
              if (it.t) it.t[i] = z;
              if (it.c) z = (it.c[i] % q + z) % q;
            }
          }
          it.r[i] = z;
        }
      }
    }
  }
  void run_gathers(const ChainPack& P, u32 n) override {
    // every source is read before any destination is written (what one parallel launch does
    // when, as the scheduler guarantees, no destination is another item's source)
    std::vector<std::vector<u64>> res(n, std::vector<u64>(kN));
    for (u32 k = 0; k < n; k++) {
      const int64_t* order = reinterpret_cast<const int64_t*>(P.it[k].b);
      for (u32 i = 0; i < kN; i++) res[k][i] = P.it[k].a[order[i]];
    }
    for (int k = (int)n - 1; k >= 0; k--)
      for (u32 i = 0; i < kN; i++) P.it[k].r[i] = res[k][i];
  }
  void run_encode(const EncodeJob* j, size_t n) override {
    for (int k = (int)n - 1; k >= 0; k--)
      for (u32 l = 0; l < j[k].level; l++)
        for (u32 i = 0; i < kN; i++)
          j[k].out[l * kN + i] = ((u64)(uintptr_t)j[k].src * 31 + l * 7 + i + j[k].sf_degree) % q_of(l);
  }
  void run_modup(const ModupJob* j, size_t n) override {
    // all inputs are read first (the real batch transforms every digit into temporaries)
    std::vector<std::vector<u64>> in(n);
    for (size_t k = 0; k < n; k++) {
      u32 len = digit_len(j[k].num_q, j[k].part);
      in[k].assign(j[k].digit, j[k].digit + len * kN);
    }
    for (int k = (int)n - 1; k >= 0; k--) {
      const u32 nq = j[k].num_q, st = digit_start(j[k].part), len = digit_len(nq, j[k].part);
      for (u32 o = 0; o < nq + kK; o++) {
        for (u32 i = 0; i < kN; i++) {
          if (o >= st && o < st + len) {
            if (j[k].copy_own) j[k].out[o * kN + i] = in[k][(o - st) * kN + i];
            continue;
          }
          u64 s = o + 1;
          for (u32 d = 0; d < len; d++) s += in[k][d * kN + i] * (d + 2);
          j[k].out[o * kN + i] = s % q_of(o);
        }
      }
    }
  }
  void run_moddown(const ModdownJob* j, size_t n) override {
    std::vector<std::vector<u64>> in(n);
    for (size_t k = 0; k < n; k++) in[k].assign(j[k].in, j[k].in + (j[k].num_q + kK) * kN);
    for (int k = (int)n - 1; k >= 0; k--)
      for (u32 l = 0; l < j[k].num_q; l++)
        for (u32 i = 0; i < kN; i++) {
          u64 s = in[k][l * kN + i];
          for (u32 p = 0; p < kK; p++) s += 3 * in[k][(j[k].num_q + p) * kN + i];
          j[k].out[l * kN + i] = s % q_of(l);
        }
  }
  void run_rescale(const RescaleJob* j, size_t n) override {
    std::vector<std::vector<u64>> in(n);
    for (size_t k = 0; k < n; k++) in[k].assign(j[k].in, j[k].in + j[k].num_q * kN);
    for (int k = (int)n - 1; k >= 0; k--)
      for (u32 l = 0; l + 1 < j[k].num_q; l++)
        for (u32 i = 0; i < kN; i++)
          j[k].out[l * kN + i] = (2 * in[k][l * kN + i] + in[k][(j[k].num_q - 1) * kN + i]) % q_of(l);
  }
};

struct Rng {
  uint64_t s;
  uint64_t next() {
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  u32 below(u32 n) { return (u32)(next() % n); }
};

struct Block {
  u64*    p;
  size_t  n;
  int64_t id;  // allocation number: addresses are recycled differently in the two runs
};

// runs the program; returns one hash per synchronisation point
std::vector<uint64_t> run_program(uint64_t seed, u32 n_ops, u32 sync_permille, bool eager,
                                  size_t* stats) {
  HostSimBackend* be = new HostSimBackend;
  Scheduler S(be);
  S.eager = eager;
  Rng rng{seed};
  std::vector<Block>    blk;
  std::vector<uint64_t> hashes;
  int64_t orders[3][kN];
  for (int t = 0; t < 3; t++)
    for (u32 i = 0; i < kN; i++) orders[t][i] = (int64_t)((i * (2 * t + 3) + t) % kN);  // permutations
  int64_t n_prog_allocs = 0;
  auto new_block = [&](size_t n, bool zeroed) {
    be->next_tag = n_prog_allocs++;
    blk.push_back(Block{S.alloc(n, zeroed), n, n_prog_allocs - 1});
  };
  int64_t last_mu_in = -1;  // the source block of the previous Decomp_modup: repeated now and then
  u32 last_mu_nq = 0, last_mu_part = 0;
  for (int i = 0; i < 6; i++) new_block(2 + rng.below(9), rng.below(2));
  // read-only blocks (cached plaintexts): written once here, only ever the source of an alias
  std::vector<Block> cblk;
  for (int i = 0; i < 3; i++) {
    be->next_tag = 1000 + i;
    Block c{S.alloc(8, false), 8, -1};
    S.encode(EncodeJob{c.p, (const void*)(uintptr_t)(7 + i), 0, 1, 8, 0, 1, 0});
    cblk.push_back(c);
  }
  S.flush();
  auto pick = [&](size_t min_limbs) -> int {  // a live block with at least min_limbs, or -1
    for (int tries = 0; tries < 16; tries++) {
      int b = (int)rng.below((u32)blk.size());
      if (blk[b].n >= min_limbs) return b;
    }
    return -1;
  };
  auto limb = [&](int b, u32 l) { return blk[b].p + (size_t)l * kN; };
  auto sync = [&] {
    S.flush();
    uint64_t h = 1469598103934665603ull;
    for (const Block& b : blk)
      for (size_t i = 0; i < b.n * kN; i++) h = (h ^ b.p[i]) * 1099511628211ull;
    if (getenv("ACE_SCHED_HASH")) {
      size_t bi = 0;
      for (const Block& b : blk) {
        for (size_t l = 0; l < b.n; l++) {
          uint64_t hl = 7;
          for (size_t i = 0; i < kN; i++) hl = hl * 31 + b.p[l * kN + i];
          fprintf(stderr, "%c sync%zu blk%zu.%zu %016llx\n", eager ? 'E' : 'D', hashes.size(), bi, l,
                  (unsigned long long)hl);
        }
        bi++;
      }
    }
    hashes.push_back(h);
  };
  const bool trace = !eager && getenv("ACE_SCHED_TRACE") != nullptr;
  auto where = [&](const u64* p) {  // "block.limb" of a limb address
    static char buf[8][48];
    static int  k = 0;
    char* o = buf[k++ & 7];
    for (size_t b = 0; b < blk.size(); b++)
      if (p >= blk[b].p && p < blk[b].p + blk[b].n * kN) {
        snprintf(o, 48, "B%zx.%zu[%p]", (size_t)((uintptr_t)blk[b].p >> 4) & 0xfff, (size_t)(p - blk[b].p) / kN, (const void*)p);
        return o;
      }
    snprintf(o, 48, "?");
    return o;
  };
#define TR(...) do { if (trace) fprintf(stderr, __VA_ARGS__); } while (0)
  for (u32 op = 0; op < n_ops; op++) {
    const u32 kind = rng.below(100);
    if (kind >= 37 && kind < 40) {  // Scheduler::alias: reads of a block are served from another one
                                    // (the plaintext cache of rt_shim.cu): later writes to either
                                    // side, frees and block consumers must all see the right data
      // (the source of an alias must stay unwritten while the alias lives: the constant blocks)
      int bd = pick(1);
      const Block& cs = cblk[rng.below((u32)cblk.size())];
      u32 n = 1 + rng.below((u32)std::min(blk[bd].n, cs.n));
      TR("%u alias %s <- const x%u\n", op, where(limb(bd, 0)), n);
      if (eager) S.copy(limb(bd, 0), cs.p, n);  // reference semantics: the data is there
      else S.alias(limb(bd, 0), cs.p, n);
    } else if (kind < 40) {  // limb add / sub / mul; operands share the limb index (one modulus per index,
                      // as in real programs: residues stay canonical), blocks may alias
      int br = pick(1), ba = pick(1), bb = pick(1);
      u32 l = rng.below((u32)std::min(blk[br].n, std::min(blk[ba].n, blk[bb].n)));
      SchedOp o = (SchedOp)(OP_ADD + rng.below(3));
      TR("%u ew%d %s = %s , %s\n", op, (int)o, where(limb(br, l)), where(limb(ba, l)), where(limb(bb, l)));
      S.ew(o, limb(br, l), limb(ba, l), limb(bb, l), l);
    } else if (kind >= 50 && kind < 55) {
      // the emitted convolutions (GEN20:1494-1500): mul t0 = a0*p; mul t1 = a1*p; add acc0 += t0;
      // add acc1 += t1 -- the first addition joins a multiplication that is NOT the last op.  Blocks
      // are drawn at random, so tmp / acc / operands alias now and then: those must not be fused
      int bt0 = pick(1), bt1 = pick(1), ba0 = pick(1), ba1 = pick(1), bp = pick(1), bc0 = pick(1), bc1 = pick(1);
      u32 lim = (u32)blk[bt0].n;
      for (int bx : {bt1, ba0, ba1, bp, bc0, bc1}) lim = std::min(lim, (u32)blk[bx].n);
      const u32 l = rng.below(lim);
      TR("%u conv t0=%s t1=%s a0=%s a1=%s p=%s acc0=%s acc1=%s\n", op, where(limb(bt0, l)), where(limb(bt1, l)),
         where(limb(ba0, l)), where(limb(ba1, l)), where(limb(bp, l)), where(limb(bc0, l)), where(limb(bc1, l)));
      S.ew(OP_MUL, limb(bt0, l), limb(ba0, l), limb(bp, l), l);
      S.ew(OP_MUL, limb(bt1, l), limb(ba1, l), limb(bp, l), l);
      if (rng.below(2)) S.ew(OP_ADD, limb(bc0, l), limb(bc0, l), limb(bt0, l), l);
      else S.ew(OP_ADD, limb(bc0, l), limb(bt0, l), limb(bc0, l), l);
      S.ew(OP_ADD, limb(bc1, l), limb(bc1, l), limb(bt1, l), l);
    } else if (kind < 55) {  // Hw_modmul(tmp, a, b); Hw_modadd(acc, acc, tmp)  (emitted pattern)
      int bt = pick(1), ba = pick(1), bb = pick(1), bc = pick(1);
      u32 l = rng.below((u32)std::min(std::min(blk[bt].n, blk[ba].n), std::min(blk[bb].n, blk[bc].n)));
      // now and then with another modulus for the addition: must NOT be fused
      u32 g2 = rng.below(16) ? l : l + 1;
      if (g2 != l && (bc == bt || blk[bc].n <= g2 || blk[bt].n <= g2)) g2 = l;
      TR("%u mac t=%s a=%s b=%s acc=%s g2=%u\n", op, where(limb(bt, l)), where(limb(ba, l)), where(limb(bb, l)),
         where(limb(bc, l)), g2);
      S.ew(OP_MUL, limb(bt, l), limb(ba, l), limb(bb, l), l);
      if (g2 != l) {
        // (values of limb l are canonical for q_l only; reduce through a multiplication by one
        // first so that the addition modulo q_{l+1} sees canonical operands in both modes)
        S.fill(limb(bc, g2), 1);
        S.ew(OP_MUL, limb(bt, g2), limb(bt, l), limb(bc, g2), g2);
        S.ew(OP_ADD, limb(bc, g2), limb(bc, g2), limb(bt, g2), g2);
      } else if (rng.below(2)) {
        S.ew(OP_ADD, limb(bc, l), limb(bc, l), limb(bt, l), l);
      } else {
        S.ew(OP_ADD, limb(bc, l), limb(bt, l), limb(bc, l), l);
      }
    } else if (kind < 60) {
      int b = pick(1);
      u32 l0 = rng.below((u32)blk[b].n);
      u32 zn = 1 + rng.below((u32)(blk[b].n - l0));
      TR("%u zero %s x%u\n", op, where(limb(b, l0)), zn);
      S.zero(limb(b, l0), zn);
    } else if (kind < 66) {
      int br = pick(1), ba = pick(1);
      u32 n = 1 + rng.below((u32)std::min(blk[br].n, blk[ba].n));
      TR("%u copy %s <- %s x%u\n", op, where(limb(br, 0)), where(limb(ba, 0)), br != ba ? n : 0);
      if (br != ba) S.copy(limb(br, 0), limb(ba, 0), n);
    } else if (kind < 69) {
      int b = pick(1);
      u64* fp = limb(b, rng.below((u32)blk[b].n));
      TR("%u fill %s\n", op, where(fp));
      S.fill(fp, rng.below(3) ? 1 + rng.below(60000) : 0);
    } else if (kind < 76) {
      int br = pick(1), ba = pick(1);
      u32 l = rng.below((u32)std::min(blk[br].n, blk[ba].n));
      TR("%u gather %s <- %s\n", op, where(limb(br, l)), where(limb(ba, l)));
      if (br != ba) S.gather(limb(br, l), limb(ba, l), orders[rng.below(3)], l);
    } else if (kind < 80) {
      int b = pick(1);
      u32 level = 1 + rng.below((u32)blk[b].n);
      TR("%u encode %s x%u\n", op, where(blk[b].p), level);
      S.encode(EncodeJob{blk[b].p, (const void*)(uintptr_t)(1 + rng.below(50)), 0, 1, level, 0,
                         1 + rng.below(2), 0});
    } else if (kind < 86) {  // Decomp_modup
      u32 nq = 1 + rng.below(6);
      int bi = pick(nq), bo = pick(nq + kK);
      u32 part = rng.below((nq + kPart - 1) / kPart);
      if (last_mu_in >= 0 && rng.below(2)) {  // the same digit of the same polynomial again (9 rotations
                                         // of one convolution input): must see intervening writes
        for (int b = 0; b < (int)blk.size(); b++)
          if (blk[b].id == last_mu_in && blk[b].n >= last_mu_nq) { bi = b; nq = last_mu_nq; part = last_mu_part; }
        bo = pick(nq + kK);
      }
      if (bi >= 0 && bo >= 0 && bi != bo) {
        TR("%u modup %s <- %s nq=%u part=%u\n", op, where(blk[bo].p), where(blk[bi].p), nq, part);
        S.modup(blk[bo].p, blk[bi].p, nq, part);
        last_mu_in = blk[bi].id; last_mu_nq = nq; last_mu_part = part;
      }
    } else if (kind < 90) {  // Mod_down
      u32 nq = 1 + rng.below(6);
      int bi = pick(nq + kK), bo = pick(nq);
      if (bi >= 0 && bo >= 0 && bi != bo) { TR("%u moddown %s <- %s nq=%u\n", op, where(blk[bo].p), where(blk[bi].p), nq); S.moddown(blk[bo].p, blk[bi].p, nq); }
    } else if (kind < 94) {  // Rescale
      u32 nq = 2 + rng.below(5);
      int bi = pick(nq), bo = pick(nq - 1);
      if (bi >= 0 && bo >= 0 && bi != bo) { TR("%u rescale %s <- %s nq=%u\n", op, where(blk[bo].p), where(blk[bi].p), nq); S.rescale(blk[bo].p, blk[bi].p, nq); }
    } else if (kind < 97) {  // Free_poly_data + Alloc_poly
      if (blk.size() > 4) {
        int b = (int)rng.below((u32)blk.size());
        TR("%u free %s\n", op, where(blk[b].p));
        S.free(blk[b].p);
        blk.erase(blk.begin() + b);
      }
      new_block(1 + rng.below(10), rng.below(2));
    } else if (rng.below(30) < sync_permille) {  // 3 % of the ops come here
      TR("%u sync\n", op);
      sync();
    }
  }
  sync();
#undef TR
  if (stats) {
    stats[0] = S.n_ops; stats[1] = S.n_flush; stats[2] = S.n_waves; stats[3] = S.n_fused;
    stats[4] = S.n_dead; stats[5] = S.n_chain_launches;
    stats[6] = S.n_modup; stats[7] = S.n_modup_shared;
  }
  for (const Block& b : blk) S.free(b.p);
  for (const Block& b : cblk) S.free(b.p);
  S.flush();
  return hashes;
}

}  // namespace
}  // namespace ace

// sync_permille: how often the program synchronises (30 = every ~33 ops, 1 = every ~1000, 0 = only
// at the end: one long deferred window).
// 0 = deferred execution reproduced call-by-call execution at every synchronisation point;
// k > 0 = first mismatch at synchronisation point k-1; stats (8 values, may be null) describe the
// deferred run: ops, flushes, waves, fused mul+add, dropped stores, chain launches, Decomp_modup
// calls, Decomp_modup calls served from an earlier result
extern "C" __attribute__((visibility("default"))) int ace_sched_selftest(uint64_t seed,
                                                                         uint32_t n_ops,
                                                                         uint32_t sync_permille,
                                                                         size_t* stats) {
  try {
    std::vector<uint64_t> a = ace::run_program(seed, n_ops, sync_permille, true, nullptr);
    std::vector<uint64_t> b = ace::run_program(seed, n_ops, sync_permille, false, stats);
    if (a.size() != b.size()) return -2;
    for (size_t i = 0; i < a.size(); i++)
      if (a[i] != b[i]) return (int)i + 1;
    return 0;
  } catch (const std::exception&) {
    return -1;
  }
}
