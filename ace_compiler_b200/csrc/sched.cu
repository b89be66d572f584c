// sched.cu -- see sched.h
#include "sched.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "prof.h"

namespace ace {

// ---------------------------------------------------------------------------- kernels
// one thread per coefficient, blockIdx.y = chain; the items of a chain run in program order.
// Plain (non-restrict, non-ldg) accesses: a later item may read what an earlier one wrote.
// (Keeping the last stored value in a register -- the accumulator of consecutive multiply-adds --
// and dropping stores that a later item of the chain overwrites was measured: 8 % SLOWER, 281 vs
// 260 ms per two ResNet-20 images; the extra dependences cost more memory-level parallelism than
// the saved loads give back.  Not kept.)
// V = 2: two adjacent coefficients per thread, 16-byte accesses (N a multiple of 512 -- every
// emitted ResNet); V = 1: any N.
template <int V> struct ChainVec;
template <> struct ChainVec<1> {
  typedef u64 type;
  static __device__ __forceinline__ u64 splat(u64 x) { return x; }
};
template <> struct ChainVec<2> {
  typedef ulonglong2 type;
  static __device__ __forceinline__ ulonglong2 splat(u64 x) { return make_ulonglong2(x, x); }
};
template <class F> __device__ __forceinline__ u64 chain_map(u64 x, u64 y, F f) { return f(x, y); }
template <class F> __device__ __forceinline__ ulonglong2 chain_map(ulonglong2 x, ulonglong2 y, F f) {
  return make_ulonglong2(f(x.x, y.x), f(x.y, y.y));
}
template <int V>
__global__ void __launch_bounds__(256) ew_chain_kernel(DeviceTables T,
                                                             const __grid_constant__ ChainPack P) {
  typedef typename ChainVec<V>::type W;
  pdl_enter();
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i * V >= T.N) return;
  const u32 k0 = P.chain_start[blockIdx.y], k1 = P.chain_start[blockIdx.y + 1];
  for (u32 k = k0; k < k1; k++) {
    const ChainItem& it = P.it[k];
    const u32 op = it.op;
    W* const r = reinterpret_cast<W*>(it.r);
    if (op == OP_ZERO) { r[i] = ChainVec<V>::splat(0); continue; }
    if (op == OP_COPY) { r[i] = reinterpret_cast<const W*>(it.a)[i]; continue; }
    if (op == OP_FILL) { r[i] = ChainVec<V>::splat((u64)(uintptr_t)it.a); continue; }
    const Modulus m = T.mod[it.g];
    const W x = reinterpret_cast<const W*>(it.a)[i], y = reinterpret_cast<const W*>(it.b)[i];
    W z;
    if (op == OP_ADD) z = chain_map(x, y, [&](u64 p, u64 q) { return add_mod(p, q, m.q); });
    else if (op == OP_SUB) z = chain_map(x, y, [&](u64 p, u64 q) { return sub_mod(p, q, m.q); });
    else {
      z = chain_map(x, y, [&](u64 p, u64 q) { return mul_mod(p, q, m); });
      if (op == OP_MAC) {
        if (it.t) reinterpret_cast<W*>(it.t)[i] = z;
        if (it.c) z = chain_map(reinterpret_cast<const W*>(it.c)[i], z, [&](u64 p, u64 q) { return add_mod(p, q, m.q); });
      }
    }
    r[i] = z;
  }
}

// independent gathers (Hw_rotate): blockIdx.y = item, b = the int64 order table.  A CTA owns 256
// consecutive outputs.  For the evaluation-domain tables of Auto_order() their sources form one
// aligned block of 256 as well (see ksw_inner_rot_kernel, kernels_ext.cu): the block is read
// coalesced into shared memory and permuted there.  Anything else (coefficient-domain tables,
// negative entries: number_theory.c:215-224) takes the direct path; the CTA decides by itself.
__global__ void __launch_bounds__(256) gather_batch_kernel(DeviceTables T,
                                                           const __grid_constant__ ChainPack P) {
  pdl_enter();
  __shared__ u64     stage[256];
  __shared__ int64_t first;
  const ChainItem& it    = P.it[blockIdx.y];
  const int64_t*   order = reinterpret_cast<const int64_t*>(it.b);
  const u64        q     = T.mod[it.g].q;
  const u32        i     = blockIdx.x * 256 + threadIdx.x;
  const bool       in    = i < T.N;
  const int64_t    k     = in ? order[i] : 0;
  if (threadIdx.x == 0) first = k;
  __syncthreads();
  const bool blocked = T.N % 256 == 0 && __syncthreads_and(k >= 0 && (k >> 8) == (first >> 8));
  if (blocked) {
    stage[threadIdx.x] = it.a[(first & ~(int64_t)255) + threadIdx.x];
    __syncthreads();
    it.r[i] = stage[k & 255];
    return;
  }
  if (in) it.r[i] = k >= 0 ? it.a[k] : q - it.a[-k];
}

// Deferred frees hold memory: all schedulers of the process (one per host thread that runs
// images) share a budget of 48 GB of blocks waiting for their flush.
static std::atomic<int> g_live_schedulers{0};

// ---------------------------------------------------------------------------- limb table
static inline u32 hash_addr(const u64* p) {
  return (u32)((((uint64_t)(uintptr_t)p) >> 9) * 0x9E3779B97F4A7C15ull >> 32);
}

// ---------------------------------------------------------------------------- product backend
namespace {
struct ContextBackend : SchedBackend {
  Context* c;
  explicit ContextBackend(Context* ctx) : c(ctx) {}
  u32    N() const override { return c->N; }
  u32    K() const override { return (u32)c->K; }
  u32    digit_start(u32 part) const override { return c->digit_start(part); }
  u32    digit_len(u32 num_q, u32 part) const override { return c->digit_len(num_q, part); }
  u64*   alloc(size_t n) override { return c->alloc_limbs(n, false); }
  void   free(u64* p) override { c->free_limbs(p); }
  size_t block_limbs(const u64* p) const override { return c->block_limbs(p); }
  void   count_limb_op(int kind) override {
    c->tr(kind == 0 ? Context::TR_LIMB_MUL : kind == 1 ? Context::TR_LIMB_ADD : Context::TR_LIMB_ROT, 0);
  }
  void run_chains(const ChainPack& pack, u32 n_chains) override {
    prof::Scope ps("ew_chain", c->stream);
    // a few long chains (the 576-term sums of the last convolutions run on 3 limbs: three chains of
    // 1 161 items per wave, ACE_SCHED_LONGCHAINS=1) leave most of the GPU idle: one coefficient
    // per thread then, twice the threads
    if (c->N % 512 == 0 && n_chains > 8)
      launch_chain(ew_chain_kernel<2>, dim3(c->N / 512, n_chains), dim3(256), 0, c->stream, c->T, pack);
    else
      launch_chain(ew_chain_kernel<1>, dim3((c->N + 255) / 256, n_chains), dim3(256), 0, c->stream, c->T, pack);
    c->launches++;
  }
  void run_gathers(const ChainPack& pack, u32 n) override {
    prof::Scope ps("gather_batch", c->stream);
    launch_chain(gather_batch_kernel, dim3((c->N + 255) / 256, n), dim3(256), 0, c->stream, c->T, pack);
    c->launches++;
  }
  void run_encode(const EncodeJob* j, size_t n) override { c->encode_batch(j, n); }
  void run_modup(const ModupJob* j, size_t n) override { c->modup_batch(j, n); }
  void run_moddown(const ModdownJob* j, size_t n) override { c->moddown_batch(j, n); }
  void run_rescale(const RescaleJob* j, size_t n) override { c->rescale_batch(j, n); }
};
}  // namespace
SchedBackend* make_context_backend(Context* c) { return new ContextBackend(c); }

Scheduler::Scheduler(SchedBackend* backend) : c_(backend) {
  g_live_schedulers++;
  table_.resize(1u << 16);
  memset(table_.data(), 0, table_.size() * sizeof(Limb));
  mask_ = (u32)table_.size() - 1;
  ops_.reserve(1 << 16);
  if (getenv("ACE_B200_NO_MODUP_SHARE")) share_modup = false;
  if (getenv("ACE_B200_MODUP_NOHIT")) modup_nohit_ = true;
}
Scheduler::~Scheduler() {
  g_live_schedulers--;
  delete c_;
}

void Scheduler::grow() {
  std::vector<Limb> old;
  old.swap(table_);
  table_.resize(old.size() * 2);
  memset(table_.data(), 0, table_.size() * sizeof(Limb));
  mask_ = (u32)table_.size() - 1;
  for (const Limb& l : old) {
    if (l.gen != gen_) continue;
    u32 h = hash_addr((const u64*)(uintptr_t)l.addr) & mask_;
    while (table_[h].gen == gen_) h = (h + 1) & mask_;
    table_[h] = l;
  }
}

// the entry of a limb, created on first sight.  References stay valid until the next limb()
// call that inserts (callers look up all operands of an op first: grow() only runs at entry)
Scheduler::Limb& Scheduler::limb(const u64* addr) {
  u32 h = hash_addr(addr) & mask_;
  for (;;) {
    Limb& l = table_[h];
    if (l.gen != gen_) {
      memset(&l, 0, sizeof(l));
      l.addr = (u64)(uintptr_t)addr;
      l.gen  = gen_;
      l.w_op = -1;
      live_++;
      return l;
    }
    if (l.addr == (u64)(uintptr_t)addr) return l;
    h = (h + 1) & mask_;
  }
}

// earliest wave of a reader / writer of limb l.  Chain ops may share a wave with the chain ops
// they depend on (program order inside the chain does the rest); anything else is one later.
u32 Scheduler::dep_read(Limb& l, bool heavy) const {
  if (!l.has_w) return 0;
  return (heavy || l.w_heavy) ? l.w_wave + 1 : l.w_wave;
}
u32 Scheduler::dep_write(Limb& l, bool heavy) const {
  u32 w = dep_read(l, heavy);
  if (l.has_r_chain) w = std::max(w, heavy ? l.r_wave_chain + 1 : l.r_wave_chain);
  if (l.has_r_heavy) w = std::max(w, l.r_wave_heavy + 1);
  return w;
}
void Scheduler::note_read(Limb& l, u32 wave, bool heavy) {
  l.read_since = 1;
  if (heavy) {
    l.r_wave_heavy = l.has_r_heavy ? std::max(l.r_wave_heavy, wave) : wave;
    l.has_r_heavy  = 1;
  } else {
    l.r_wave_chain = l.has_r_chain ? std::max(l.r_wave_chain, wave) : wave;
    l.has_r_chain  = 1;
  }
}
void Scheduler::note_write(Limb& l, u32 wave, bool heavy, int32_t op, bool as_t) {
  l.w_op = heavy ? -1 : op;
  l.has_w = 1;
  l.w_wave = wave;
  l.w_heavy = heavy ? 1 : 0;
  l.w_is_t = as_t ? 1 : 0;
  l.read_since = 0;
  l.has_r_chain = l.has_r_heavy = 0;
  l.is_zero = 0;
  l.w_seq = ++wseq_;
  l.alias = nullptr;
}

const u64* Scheduler::resolve(const u64* a) {
  const u64* t = limb(a).alias;
  return t ? t : a;
}

// consumers that take a block of consecutive limbs cannot follow per-limb renames: copy the kept
// data into place first (recorded like any other copy)
void Scheduler::materialize(const u64* p, size_t n) {
  for (size_t i = 0; i < n; i++) {
    u64* q = const_cast<u64*>(p) + i * c_->N();
    if (live_ + 8 > table_.size() / 2) grow();
    const u64* src = limb(q).alias;
    if (!src) continue;
    limb(q).alias = nullptr;
    copy(q, src, 1);
  }
}

// the limb is about to be overwritten (or its block freed): if nothing read what the last
// recorded chain op stored there, that store is dead
void Scheduler::kill_if_unread(Limb& l) {
  if (l.w_op < 0 || l.read_since) return;
  Op& o = ops_[l.w_op];
  if (l.w_is_t) {
    if (o.kind == OP_MAC) o.t_live = 0;
  } else if (o.kind == OP_MAC && o.t_live) {
    // only the product is still wanted: r = a * b stored to t
    o.kind = OP_MUL;
    o.r = o.t;
    o.t_live = 0;
    Limb& lt = limb(o.t);  // exists already
    if (lt.w_op == l.w_op) lt.w_is_t = 0;  // unless a later op has written t since
  } else if (o.kind != OP_NOP) {
    o.kind = OP_NOP;
  }
  n_dead++;
  l.w_op = -1;
}

void Scheduler::maybe_flush() {
  if (in_flush_) return;
  const size_t limit = ((size_t)48 << 30) / (size_t)std::max(1, g_live_schedulers.load());
  if (eager || ops_.size() > (1u << 18) || pending_free_bytes_ > limit) flush();
}

// ---------------------------------------------------------------------------- recording
u64* Scheduler::alloc(size_t n_limbs, bool zeroed) {
  u64* p = c_->alloc(n_limbs);
  if (zeroed) zero(p, n_limbs);
  return p;
}

void Scheduler::free(u64* block) {
  if (!block) return;
  const size_t n = c_->block_limbs(block);
  for (size_t i = 0; i < n; i++) {
    const u64* a = block + i * c_->N();
    u32 h = hash_addr(a) & mask_;
    while (table_[h].gen == gen_) {
      if (table_[h].addr == (u64)(uintptr_t)a) { kill_if_unread(table_[h]); table_[h].alias = nullptr; break; }
      h = (h + 1) & mask_;
    }
  }
  frees_.push_back(block);
  pending_free_bytes_ += n * c_->N() * sizeof(u64);
  maybe_flush();
}

void Scheduler::alias(u64* dst, const u64* src, size_t n_limbs) {
  if (live_ + n_limbs + 8 > table_.size() / 2) grow();
  if (eager) { copy(dst, src, n_limbs); return; }
  for (size_t i = 0; i < n_limbs; i++) {
    u64* o = dst + i * c_->N();
    Limb& l = limb(o);
    kill_if_unread(l);
    if (!l.alias) aliased_.push_back(o);
    l.alias = src + i * c_->N();
    l.is_zero = 0;
  }
  n_ops += n_limbs;
  maybe_flush();
}

void Scheduler::zero(u64* r, size_t n_limbs) {
  if (live_ + n_limbs + 8 > table_.size() / 2) grow();
  for (size_t i = 0; i < n_limbs; i++) {
    u64* p = r + i * c_->N();
    Limb& l = limb(p);
    kill_if_unread(l);
    const u32 wave = dep_write(l, false);
    Op o{};
    o.kind = OP_ZERO; o.r = p; o.wave = wave;
    ops_.push_back(o);
    note_write(l, wave, false, (int32_t)ops_.size() - 1, false);
    l.is_zero = 1;
  }
  n_ops += n_limbs;
  maybe_flush();
}

void Scheduler::fill(u64* r, u64 value) {
  if (value == 0) { zero(r, 1); return; }
  if (live_ + 8 > table_.size() / 2) grow();
  Limb& l = limb(r);
  kill_if_unread(l);
  const u32 wave = dep_write(l, false);
  Op o{};
  o.kind = OP_FILL; o.r = r; o.wave = wave;
  o.a = reinterpret_cast<const u64*>((uintptr_t)value);
  ops_.push_back(o);
  note_write(l, wave, false, (int32_t)ops_.size() - 1, false);
  n_ops++;
  maybe_flush();
}

void Scheduler::copy(u64* r, const u64* a, size_t n_limbs) {
  for (size_t i = 0; i < n_limbs; i++) {
    u64* rp = r + i * c_->N();
    const u64* ap = a + i * c_->N();
    if (rp == ap) continue;
    if (live_ + 8 > table_.size() / 2) grow();
    ap = resolve(ap);
    Limb& la = limb(ap);
    if (la.is_zero) { zero(rp, 1); continue; }
    const u32 wa = dep_read(la, false);
    note_read(la, wa, false);  // provisional; raised below if the write forces a later wave
    Limb& lr = limb(rp);
    kill_if_unread(lr);
    const u32 wave = std::max(wa, dep_write(lr, false));
    Op o{};
    o.kind = OP_COPY; o.r = rp; o.a = ap; o.wave = wave;
    ops_.push_back(o);
    note_write(lr, wave, false, (int32_t)ops_.size() - 1, false);
    if (wave != wa) note_read(limb(ap), wave, false);
    n_ops++;
  }
  maybe_flush();
}

void Scheduler::ew(SchedOp op, u64* r, const u64* a, const u64* b, u32 g) {
  if (live_ + 8 > table_.size() / 2) grow();
  c_->count_limb_op(op == OP_MUL ? 0 : 1);
  n_ops++;
  a = resolve(a);
  b = resolve(b);
  const bool za = limb(a).is_zero, zb = limb(b).is_zero;
  // ---- operands known to be zero: the op degenerates (0 + y = y, x * 0 = 0, x - 0 = x)
  if (za || zb) {
    if (op == OP_MUL || (za && zb)) { zero(r, 1); return; }
    if (op == OP_ADD) { copy(r, za ? b : a, 1); return; }
    if (op == OP_SUB && zb) { copy(r, a, 1); return; }
  }
  // ---- Hw_modmul(tmp, x, y) followed by Hw_modadd(r, acc, tmp): one multiply-add.  Directly
  // followed (the key inner products of Rotate / Relinearize: mul, add, mul, add), or with ONE
  // unrelated multiplication in between (the emitted convolutions: mul c0, mul c1, add c0, add
  // c1 -- GEN20:1494-1500): the op in between must not touch tmp, acc or r, so the addition
  // commutes with it and can join the multiplication where that stands.
  static const int max_back = getenv("ACE_B200_NO_FUSE2") ? 1 : 2;
  for (int back = 1; op == OP_ADD && back <= max_back && ops_.size() >= (size_t)back; back++) {
    const int32_t mi = (int32_t)ops_.size() - back;
    Op& m = ops_[mi];
    if (m.kind == OP_MUL && m.g == (uint16_t)g && m.r != r && ((b == m.r) != (a == m.r))) {
      const u64* acc = (b == m.r) ? a : b;
      u64*       tmp = m.r;
      if (back == 2) {
        const Op& x = ops_.back();
        const u64* touched[3] = {tmp, acc, r};
        bool clash = x.kind != OP_MUL;
        for (const u64* p : touched) clash |= x.r == p;
        clash |= x.a == tmp || x.a == r || x.b == tmp || x.b == r;
        // the multiply-add may move to a LATER wave than x (it waits for acc): x must not
        // overwrite what it still has to read
        clash |= x.r == m.a || x.r == m.b;
        if (clash) break;
      }
      Limb& lt = limb(tmp);
      if (lt.w_op == mi && !lt.read_since) {
        Limb& lacc = limb(acc);
        const bool acc_zero = lacc.is_zero;
        u32 wave = m.wave;
        if (!acc_zero) {
          wave = std::max(wave, dep_read(lacc, false));
          note_read(lacc, wave, false);
        }
        Limb& lr = limb(r);
        kill_if_unread(lr);
        wave = std::max(wave, dep_write(lr, false));
        m.kind = OP_MAC; m.c = acc_zero ? nullptr : acc; m.t = tmp; m.t_live = 1; m.r = r;
        m.wave = wave;
        const int32_t idx = mi;
        note_write(lr, wave, false, idx, false);
        Limb& lt2 = limb(tmp);
        lt2.w_wave = wave; lt2.w_is_t = 1; lt2.w_op = idx;
        // reads happen in the final wave (the value the destination held before is what is read)
        if (m.a != r && m.a != tmp) note_read(limb(m.a), wave, false);
        if (m.b != r && m.b != tmp) note_read(limb(m.b), wave, false);
        if (!acc_zero && acc != r) note_read(limb(acc), wave, false);
        n_fused++;
        maybe_flush();
        return;
      }
    }
  }
  Limb& la = limb(a);
  u32 wave = dep_read(la, false);
  Limb& lb = limb(b);
  wave = std::max(wave, dep_read(lb, false));
  note_read(limb(a), wave, false);
  note_read(limb(b), wave, false);
  Limb& lr = limb(r);
  kill_if_unread(lr);
  const u32 w2 = std::max(wave, dep_write(lr, false));
  Op o{};
  o.kind = op; o.r = r; o.a = a; o.b = b; o.g = (uint16_t)g; o.wave = w2;
  ops_.push_back(o);
  note_write(limb(r), w2, false, (int32_t)ops_.size() - 1, false);
  if (w2 != wave) {
    note_read(limb(a), w2, false);
    note_read(limb(b), w2, false);
    if (a == r || b == r) limb(r).read_since = 0;
  }
  // an op that reads its own destination: the read is of the OLD value and precedes the write
  if (a == r || b == r) { Limb& l = limb(r); l.read_since = 0; l.has_r_chain = 0; }
  maybe_flush();
}

void Scheduler::gather(u64* r, const u64* a, const int64_t* order, u32 g) {
  if (live_ + 8 > table_.size() / 2) grow();
  c_->count_limb_op(2);
  n_ops++;
  if (r == a) throw std::runtime_error("Hw_rotate in place is not supported");
  a = resolve(a);
  Limb& la = limb(a);
  u32 wave = dep_read(la, true);
  Limb& lr = limb(r);
  kill_if_unread(lr);
  wave = std::max(wave, dep_write(lr, true));
  note_read(limb(a), wave, true);
  Op o{};
  o.kind = OP_GATHER; o.r = r; o.a = a; o.b = reinterpret_cast<const u64*>(order);
  o.g = (uint16_t)g; o.wave = wave;
  ops_.push_back(o);
  note_write(limb(r), wave, true, -1, false);
  maybe_flush();
}

// a heavy op reading limbs [rd, rd + n_rd) and writing limbs [wr, wr + n_wr)
static inline size_t limb_off(u32 N, size_t i) { return i * (size_t)N; }

void Scheduler::encode(const EncodeJob& j) {
  const size_t nw = j.level + j.p_cnt;
  if (live_ + nw + 8 > table_.size() / 2) grow();
  u32 wave = 0;
  for (size_t i = 0; i < nw; i++) {
    Limb& l = limb(j.out + limb_off(c_->N(), i));
    kill_if_unread(l);
    wave = std::max(wave, dep_write(l, true));
  }
  for (size_t i = 0; i < nw; i++) note_write(limb(j.out + limb_off(c_->N(), i)), wave, true, -1, false);
  Op o{};
  o.kind = OP_ENCODE; o.wave = wave; o.p0 = (u32)enc_jobs_.size();
  enc_jobs_.push_back(j);
  ops_.push_back(o);
  n_ops++;
  maybe_flush();
}

void Scheduler::modup(u64* out, const u64* in, u32 num_q, u32 part) {
  const u32 N = c_->N(), W = num_q + (u32)c_->K();
  const u32 st = c_->digit_start(part), len = c_->digit_len(num_q, part);
  if (live_ + 2 * W + len + 8 > table_.size() / 2) grow();
  n_modup++;
  materialize(in + limb_off(N, st), len);
  u64* dst = out;
  if (share_modup && !eager) {
    // a kept result of the same digit of the same, unmodified polynomial?
    const ModupKey key{in + limb_off(N, st), num_q, part};
    auto it = modup_cache_.find(key);
    bool hit = it != modup_cache_.end() && !modup_nohit_;
    if (hit)
      for (u32 i = 0; i < len && hit; i++) hit = limb(in + limb_off(N, st + i)).w_seq <= it->second.stamp;
    if (!hit) {
      dst = c_->alloc(W);
      pending_free_bytes_ += (size_t)W * N * sizeof(u64);
      if (it != modup_cache_.end()) {
        frees_.push_back(it->second.buf);  // stale: released once the recorded ops are issued
        it->second = ModupEntry{dst, wseq_};
      } else {
        modup_cache_[key] = ModupEntry{dst, wseq_};
      }
    } else {
      dst = it->second.buf;
      n_modup_shared++;
    }
    // the caller's limbs now stand for the kept ones; what they held is unobservable
    for (u32 i = 0; i < W; i++) {
      u64* o = out + limb_off(N, i);
      Limb& l = limb(o);
      kill_if_unread(l);
      if (!l.alias) aliased_.push_back(o);
      l.alias = dst + limb_off(N, i);
      l.is_zero = 0;
    }
    n_ops++;
    if (hit) { maybe_flush(); return; }
  }
  u32 wave = 0;
  for (u32 i = 0; i < len; i++) wave = std::max(wave, dep_read(limb(in + limb_off(N, st + i)), true));
  const bool overlap = dst + limb_off(N, W) > in + limb_off(N, st) && dst < in + limb_off(N, st + len);
  for (u32 i = 0; i < W; i++) {
    Limb& l = limb(dst + limb_off(N, i));
    if (!overlap) kill_if_unread(l);
    wave = std::max(wave, dep_write(l, true));
  }
  for (u32 i = 0; i < len; i++) note_read(limb(in + limb_off(N, st + i)), wave, true);
  for (u32 i = 0; i < W; i++) note_write(limb(dst + limb_off(N, i)), wave, true, -1, false);
  Op o{};
  o.kind = OP_MODUP; o.r = dst; o.a = in + limb_off(N, st); o.p0 = num_q; o.p1 = part; o.wave = wave;
  ops_.push_back(o);
  if (dst == out) n_ops++;
  maybe_flush();
}

void Scheduler::moddown(u64* out, const u64* in, u32 num_q) {
  const u32 N = c_->N(), W = num_q + (u32)c_->K();
  materialize(in, W);
  if (live_ + W + num_q + 8 > table_.size() / 2) grow();
  u32 wave = 0;
  for (u32 i = 0; i < W; i++) wave = std::max(wave, dep_read(limb(in + limb_off(N, i)), true));
  const bool overlap = out + limb_off(N, num_q) > in && out < in + limb_off(N, W);
  for (u32 i = 0; i < num_q; i++) {
    Limb& l = limb(out + limb_off(N, i));
    if (!overlap) kill_if_unread(l);  // (in place: what the last op stored there is the input)
    wave = std::max(wave, dep_write(l, true));
  }
  for (u32 i = 0; i < W; i++) note_read(limb(in + limb_off(N, i)), wave, true);
  for (u32 i = 0; i < num_q; i++) note_write(limb(out + limb_off(N, i)), wave, true, -1, false);
  Op o{};
  o.kind = OP_MODDOWN; o.r = out; o.a = in; o.p0 = num_q; o.wave = wave;
  ops_.push_back(o);
  n_ops++;
  maybe_flush();
}

void Scheduler::rescale(u64* out, const u64* in, u32 num_q) {
  const u32 N = c_->N();
  if (num_q < 2) throw std::runtime_error("Rescale: level not enough");
  materialize(in, num_q);
  if (live_ + 2 * num_q + 8 > table_.size() / 2) grow();
  u32 wave = 0;
  for (u32 i = 0; i < num_q; i++) wave = std::max(wave, dep_read(limb(in + limb_off(N, i)), true));
  for (u32 i = 0; i + 1 < num_q; i++) {
    Limb& l = limb(out + limb_off(N, i));
    if (out != in) kill_if_unread(l);
    wave = std::max(wave, dep_write(l, true));
  }
  for (u32 i = 0; i < num_q; i++) note_read(limb(in + limb_off(N, i)), wave, true);
  for (u32 i = 0; i + 1 < num_q; i++) note_write(limb(out + limb_off(N, i)), wave, true, -1, false);
  Op o{};
  o.kind = OP_RESCALE; o.r = out; o.a = in; o.p0 = num_q; o.wave = wave;
  ops_.push_back(o);
  n_ops++;
  maybe_flush();
}

// ---------------------------------------------------------------------------- issuing
// chain ops of one wave: ops that touch a limb written in this wave belong to one chain
void Scheduler::run_chains(std::vector<u32>& idx) {
  const size_t n = idx.size();
  if (n == 0) return;
  // written limbs -> representative item (first writer), open addressing
  size_t cap = 16;
  while (cap < 4 * n) cap <<= 1;
  static thread_local std::vector<std::pair<const u64*, u32>> wmap;
  wmap.assign(cap, std::make_pair((const u64*)nullptr, 0u));
  auto wfind = [&](const u64* p, bool insert, u32 k) -> int32_t {
    size_t h = hash_addr(p) & (cap - 1);
    for (;;) {
      if (wmap[h].first == nullptr) {
        if (!insert) return -1;
        wmap[h] = std::make_pair(p, k);
        return (int32_t)k;
      }
      if (wmap[h].first == p) return (int32_t)wmap[h].second;
      h = (h + 1) & (cap - 1);
    }
  };
  std::vector<u32> parent(n);
  for (u32 k = 0; k < n; k++) parent[k] = k;
  auto find = [&](u32 x) {
    while (parent[x] != x) x = parent[x] = parent[parent[x]];
    return x;
  };
  auto unite = [&](u32 a, u32 b) {
    a = find(a); b = find(b);
    if (a != b) parent[std::max(a, b)] = std::min(a, b);
  };
  for (u32 k = 0; k < n; k++) {
    const Op& o = ops_[idx[k]];
    unite(k, (u32)wfind(o.r, true, k));
    if (o.kind == OP_MAC && o.t_live) unite(k, (u32)wfind(o.t, true, k));
  }
  for (u32 k = 0; k < n; k++) {
    const Op& o = ops_[idx[k]];
    const u64* rd[3] = {o.kind == OP_FILL ? nullptr : o.a, o.b, o.c};
    for (const u64* p : rd) {
      if (!p) continue;
      int32_t w = wfind(p, false, 0);
      if (w >= 0) unite(k, (u32)w);
    }
  }
  // group by chain, program order inside a chain (idx is in program order already)
  std::vector<int32_t> chain_id(n, -1);
  std::vector<u32>     chain_of(n);
  u32 n_chains = 0;
  for (u32 k = 0; k < n; k++) {
    u32 root = find(k);
    if (chain_id[root] < 0) chain_id[root] = (int32_t)n_chains++;
    chain_of[k] = (u32)chain_id[root];
  }
  std::vector<u32> cnt(n_chains + 1, 0);
  for (u32 k = 0; k < n; k++) cnt[chain_of[k] + 1]++;
  for (u32 ch = 0; ch < n_chains; ch++) cnt[ch + 1] += cnt[ch];
  std::vector<u32> order(n), pos(cnt.begin(), cnt.end() - 1);
  for (u32 k = 0; k < n; k++) order[pos[chain_of[k]]++] = k;
  // pack whole chains into launches; a chain longer than a launch continues in the next one
  static thread_local ChainPack pack;
  u32 items = 0, chains = 0;
  auto launch = [&] {
    if (items == 0) return;
    pack.chain_start[chains] = (uint16_t)items;
    pack.n_chains = chains;
    c_->run_chains(pack, chains);
    n_chain_launches++;
    items = chains = 0;
  };
  static const bool dump_long = getenv("ACE_SCHED_LONGCHAINS") != nullptr;  // debugging aid
  for (u32 ch = 0; ch < n_chains; ch++) {
    u32 len = cnt[ch + 1] - cnt[ch], at = cnt[ch];
    if (dump_long && len > (u32)kChainCap) {
      u32 kinds[16] = {0}, tlive = 0;
      std::vector<const u64*> rs, ts, as;
      for (u32 q = 0; q < len; q++) {
        const Op& o = ops_[idx[order[at + q]]];
        kinds[o.kind & 15]++;
        rs.push_back(o.r);
        if (o.kind == OP_MAC && o.t_live) { tlive++; ts.push_back(o.t); }
      }
      std::sort(rs.begin(), rs.end()); rs.erase(std::unique(rs.begin(), rs.end()), rs.end());
      std::sort(ts.begin(), ts.end()); ts.erase(std::unique(ts.begin(), ts.end()), ts.end());
      fprintf(stderr, "[sched] chain of %u items (of %zu in the wave, %u chains): distinct r %zu, live t %u (distinct %zu); kinds",
              len, n, n_chains, rs.size(), tlive, ts.size());
      for (int k = 0; k < 16; k++) if (kinds[k]) fprintf(stderr, " %d:%u", k, kinds[k]);
      fprintf(stderr, "\n");
      for (u32 q = 0; q < 6 && q < len; q++) {
        const Op& o = ops_[idx[order[at + q]]];
        fprintf(stderr, "    kind %d g %u r=%p a=%p b=%p c=%p t=%p tl=%d\n", (int)o.kind, o.g, (void*)o.r, (void*)o.a, (void*)o.b, (void*)o.c, (void*)o.t, (int)o.t_live);
      }
    }
    while (len > 0) {
      if (items == (u32)kChainCap || (items > 0 && items + len > (u32)kChainCap && len <= (u32)kChainCap))
        launch();
      const u32 take = std::min<u32>(len, (u32)kChainCap - items);
      pack.chain_start[chains++] = (uint16_t)items;
      for (u32 q = 0; q < take; q++) {
        const Op& o = ops_[idx[order[at + q]]];
        ChainItem& it = pack.it[items++];
        it.r = o.r; it.a = o.a; it.b = o.b; it.c = o.c;
        it.t = (o.kind == OP_MAC && o.t_live) ? o.t : nullptr;
        it.g = o.g; it.op = o.kind;
      }
      at += take; len -= take;
      if (len > 0) launch();  // the rest of this chain must run after this launch
    }
  }
  launch();
}

void Scheduler::flush() {
  // renamed limbs that are still alive get their data in place: after the window the caller may
  // read them by other means (the table of renames does not survive the window)
  if (!aliased_.empty()) {
    std::vector<u64*> al;
    al.swap(aliased_);
    in_flush_ = true;  // the copies recorded below must not re-enter flush()
    for (u64* q : al) materialize(q, 1);
    in_flush_ = false;
  }
  for (auto& kv : modup_cache_) frees_.push_back(kv.second.buf);
  modup_cache_.clear();
  if (ops_.empty()) {
    for (u64* p : frees_) c_->free(p);
    frees_.clear();
    pending_free_bytes_ = 0;
    return;
  }
  n_flush++;
  u32 max_wave = 0;
  for (const Op& o : ops_) max_wave = std::max(max_wave, o.wave);
  // counting sort of op indices by wave (stable: program order inside a wave)
  std::vector<u32> start(max_wave + 2, 0);
  for (const Op& o : ops_) if (o.kind != OP_NOP) start[o.wave + 1]++;
  for (u32 w = 0; w <= max_wave; w++) start[w + 1] += start[w];
  std::vector<u32> sorted(start[max_wave + 1]), pos(start.begin(), start.end() - 1);
  for (u32 i = 0; i < ops_.size(); i++)
    if (ops_[i].kind != OP_NOP) sorted[pos[ops_[i].wave]++] = i;
  static const bool dump = getenv("ACE_SCHED_DUMP") != nullptr;  // debugging aid of the self-test
  if (dump) {
    for (u32 w = 0; w <= max_wave; w++)
      for (u32 s2 = start[w]; s2 < start[w + 1]; s2++) {
        const Op& o = ops_[sorted[s2]];
        fprintf(stderr, "  wave %u op#%u kind %d r=%p a=%p b=%p c=%p t=%p tl=%d\n", w, sorted[s2], (int)o.kind,
                (void*)o.r, (void*)o.a, (void*)o.b, (void*)o.c, (void*)o.t, (int)o.t_live);
      }
  }
  std::vector<u32>        chain_idx, gather_idx;
  std::vector<EncodeJob>  enc;
  std::vector<ModupJob>   mu;
  std::vector<ModdownJob> md;
  std::vector<RescaleJob> rs;
  static thread_local ChainPack gpack;
  for (u32 w = 0; w <= max_wave; w++) {
    if (start[w] == start[w + 1]) continue;
    n_waves++;
    chain_idx.clear(); gather_idx.clear(); enc.clear(); mu.clear(); md.clear(); rs.clear();
    for (u32 s = start[w]; s < start[w + 1]; s++) {
      const Op& o = ops_[sorted[s]];
      switch (o.kind) {
        case OP_GATHER:  gather_idx.push_back(sorted[s]); break;
        case OP_ENCODE:  enc.push_back(enc_jobs_[o.p0]); break;
        case OP_MODUP:   mu.push_back(ModupJob{o.r, o.a, o.p0, o.p1, true}); break;
        case OP_MODDOWN: md.push_back(ModdownJob{o.r, o.a, o.p0}); break;
        case OP_RESCALE: rs.push_back(RescaleJob{o.r, o.a, o.p0}); break;
        default:         chain_idx.push_back(sorted[s]); break;
      }
    }
    run_chains(chain_idx);
    for (size_t at = 0; at < gather_idx.size(); at += kChainCap) {
      const u32 cnt = (u32)std::min<size_t>(kChainCap, gather_idx.size() - at);
      for (u32 k = 0; k < cnt; k++) {
        const Op& o = ops_[gather_idx[at + k]];
        ChainItem& it = gpack.it[k];
        it.r = o.r; it.a = o.a; it.b = o.b; it.c = nullptr; it.t = nullptr; it.g = o.g; it.op = OP_GATHER;
      }
      gpack.n_chains = cnt;
      c_->run_gathers(gpack, cnt);
    }
    if (!enc.empty()) c_->run_encode(enc.data(), enc.size());
    if (!mu.empty()) c_->run_modup(mu.data(), mu.size());
    if (!md.empty()) c_->run_moddown(md.data(), md.size());
    if (!rs.empty()) c_->run_rescale(rs.data(), rs.size());
  }
  ops_.clear();
  enc_jobs_.clear();
  for (u64* p : frees_) c_->free(p);
  frees_.clear();
  pending_free_bytes_ = 0;
  gen_++;
  live_ = 0;
  wseq_ = 0;
  if (gen_ == 0) {  // generation counter wrapped: really clear the table
    memset(table_.data(), 0, table_.size() * sizeof(Limb));
    gen_ = 1;
  }
}

}  // namespace ace
