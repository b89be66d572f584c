// client.cu -- key generation, encryption/decryption and CKKS encode/decode for the B200
// runtime.  Arithmetic runs on the GPU with the kernels of kernels.cu.  Sampling: by default
// ChaCha20 keyed from the operating system's entropy (keys are valid, not the reference's);
// keygen_reference() consumes the reference's own BLAKE2Xb / rand() streams (refrng.h) in the
// reference's order and reproduces its keys and encryptions bit for bit from the same seeds.
//
// Reference routines followed (paths under fhe-cmplr/rtlib/ant/):
//   keys       src/util/ckks_key_generator.c:69-336, src/util/polynomial.c:1349-1412
//   encrypt    src/util/ckks_encryptor.c:20-95        decrypt  src/util/ckks_decryptor.c:19-65
//   encode     src/util/ckks_encoder.c:199-299 (+464-528 for constants), src/util/ntt.c:713-753
//   decode     src/util/ckks_encoder.c:649-703, src/util/polynomial.c:467-497, ntt.c:672-711
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <sys/random.h>

#include <thread>

#include "context.h"
#include "refrng.h"
#include "prof.h"
#include "host_math.h"

namespace ace {

// ------------------------------------------------------------------------------ sampling
// Default sampler: ChaCha20 (RFC 8439 block function) as a counter-based generator.  The 256-bit
// key comes from the operating system (getrandom) unless the caller pins a seed (tests); every
// sample is addressed by (purpose, key id, digit, limb, coefficient) in the nonce / counter words,
// so streams of different keys, digits and encryptions never overlap (domain separation instead
// of seed arithmetic).  One 64-byte block per coefficient: 8 64-bit candidates for rejection
// sampling below the largest multiple of q (bias-free; all 8 rejected: probability < 2^-100).
// The reference's own generators (BLAKE2Xb + rand(), refrng.h) are used by keygen_reference().
enum : u32 { kPurposeUniform = 1, kPurposeError = 2, kPurposeEncU = 3, kPurposeEncE1 = 4, kPurposeEncE2 = 5 };

struct ChaChaKey { u32 k[8]; };

__host__ __device__ __forceinline__ u32 rotl32(u32 x, int n) { return (x << n) | (x >> (32 - n)); }
__host__ __device__ inline void chacha20_block(const u32 (&key)[8], u32 counter, u32 n0, u32 n1, u32 n2,
                                               u32 (&out)[16]) {
  u32 st[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3],
                key[4], key[5], key[6], key[7], counter, n0, n1, n2};
  u32 x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = st[i];
#define ACE_QR(a, b, c, d)                                   \
  x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12); \
  x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);
#pragma unroll
  for (int r = 0; r < 10; r++) {
    ACE_QR(0, 4, 8, 12) ACE_QR(1, 5, 9, 13) ACE_QR(2, 6, 10, 14) ACE_QR(3, 7, 11, 15)
    ACE_QR(0, 5, 10, 15) ACE_QR(1, 6, 11, 12) ACE_QR(2, 7, 8, 13) ACE_QR(3, 4, 9, 14)
  }
#undef ACE_QR
#pragma unroll
  for (int i = 0; i < 16; i++) out[i] = x[i] + st[i];
}

// uniform residues of limb l (modulus g[l]): nonce = (purpose | digit << 8 | limb << 16, id_lo, id_hi)
__global__ void uniform_kernel(DeviceTables T, LimbBatch b, ChaChaKey key, u32 digit, u64 id) {
  const u32     limb = blockIdx.y;
  const Modulus m    = T.mod[b.g[limb]];
  u64*          out  = b.base + (size_t)b.slot[limb] * T.N;
  const u64     lim  = 0 - m.c64;  // 2^64 - (2^64 mod q): the largest multiple of q below 2^64
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < T.N; i += gridDim.x * blockDim.x) {
    u32 w[16];
    chacha20_block(key.k, i, kPurposeUniform | (digit << 8) | (limb << 16), (u32)id, (u32)(id >> 32), w);
    u64 x = ((u64)w[1] << 32) | w[0];
#pragma unroll
    for (int c = 1; c < 8; c++)
      if (x >= lim) x = ((u64)w[2 * c + 1] << 32) | w[2 * c];
    u64 r = x - __umul64hi(x, m.mu_hi) * m.q;
    out[i] = r >= m.q ? r - m.q : r;
  }
}

// "triangle" samples (random_sample.c:78-97): -1, +1 with probability 1/4 each, else 0;
// the same small coefficient is written to every limb of the batch (Transform_values_to_rns)
__global__ void triangle_kernel(DeviceTables T, LimbBatch b, ChaChaKey key, u32 purpose, u32 digit, u64 id) {
  const u32 limb = blockIdx.y;
  const u64 q    = T.mod[b.g[limb]].q;
  u64*      out  = b.base + (size_t)b.slot[limb] * T.N;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < T.N; i += gridDim.x * blockDim.x) {
    u32 w[16];
    chacha20_block(key.k, i, purpose | (digit << 8), (u32)id, (u32)(id >> 32), w);
    const u32 r = w[0] & 3;
    out[i] = r == 0 ? q - 1 : (r == 1 ? 1 : 0);
  }
}

// small signed coefficients (host-sampled) -> residues
__global__ void small_to_rns_kernel(DeviceTables T, LimbBatch b, const int64_t* __restrict__ v) {
  const u32 limb = blockIdx.y;
  const u64 q    = T.mod[b.g[limb]].q;
  u64*      out  = b.base + (size_t)b.slot[limb] * T.N;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < T.N; i += gridDim.x * blockDim.x) {
    int64_t x = v[i];
    out[i]    = x < 0 ? q - (u64)(-x) % q : (u64)x % q;
    if (out[i] == q) out[i] = 0;
  }
}

// out = e + fac[l] * nk - a * old   on the limbs of the batch (fac = 0 outside the digit)
__global__ void swk_b_kernel(DeviceTables T, LimbBatch b, const u64* __restrict__ a,
                             const u64* __restrict__ e, const u64* __restrict__ nk,
                             const u64* __restrict__ old, const u64* __restrict__ fac) {
  const u32     limb = blockIdx.y;
  const Modulus m    = T.mod[b.g[limb]];
  const size_t  off  = (size_t)b.slot[limb] * T.N;
  const u64     f    = fac[limb];
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < T.N; i += gridDim.x * blockDim.x) {
    u64 v = e[off + i];
    if (f) v = add_mod(v, mul_mod(nk[off + i], f, m), m.q);
    b.base[off + i] = sub_mod(v, mul_mod(a[off + i], old[off + i], m), m.q);
  }
}

static LimbBatch all_limbs(u64* base, u32 g0, u32 n) {
  LimbBatch b;
  b.base = base; b.src = nullptr; b.n = n;
  for (u32 i = 0; i < n; i++) { b.slot[i] = (uint16_t)i; b.g[i] = (uint16_t)(g0 + i); }
  return b;
}

static dim3 grid_for(u32 N, u32 n) { return dim3((N + 255) / 256, n); }

// ---------------------------------------------------------------------------- generator state
// secure mode: key = 256 bits from the OS (or expanded from a pinned test seed);
// reference mode: the reference's own streams (refrng.h), consumed in the reference's order
struct Context::RngState {
  ChaChaKey          key;
  bool               reference = false;
  refrng::BulkPrng   prng;           // reference mode: BLAKE2Xb stream (uniform, ternary)
  u32                tri_base = 0;   // reference mode: Sample_triangle call k draws from
  u32                tri_calls = 0;  //   glibc random() after srandom(tri_base + k), or -- stream
  bool               tri_stream = false;  // mode -- from ONE random() stream seeded once, call k
  refrng::GlibcRandom tri_gen;            // starting tri_pos[k] draws after the seeding (the
  u64                tri_at = 0;          // reference's harness in pin mode 0: srand() swallowed;
  std::vector<u64>   tri_pos;             // positions recorded by tests/golden/make_tri_positions.py)
  std::vector<int64_t> host;         // staging for host-sampled values
};

void refrng::BulkPrng::refill() {
  words.resize(kBatch * 1024);
  const unsigned nt = std::max(1u, std::min(threads, std::thread::hardware_concurrency()));
  std::vector<std::thread> pool;
  const uint64_t c0 = counter;
  for (unsigned t = 0; t < nt; t++)
    pool.emplace_back([this, t, nt, c0] {
      for (size_t k = t; k < kBatch; k += nt) {
        const uint64_t c = c0 + k;
        uint8_t in[8];
        for (int i = 0; i < 8; i++) in[i] = (uint8_t)(c >> (8 * i));
        blake2xb(reinterpret_cast<uint8_t*>(words.data() + k * 1024), 4096, in, 8,
                 reinterpret_cast<const uint8_t*>(seed), 64);
      }
    });
  for (auto& th : pool) th.join();
  counter += kBatch;
  pos = 0;
}

Context::RngState* Context::rng() {
  if (!rng_) {
    rng_ = std::make_shared<RngState>();
    if (getrandom(rng_->key.k, sizeof(rng_->key.k), 0) != (ssize_t)sizeof(rng_->key.k))
      throw std::runtime_error("getrandom failed: no entropy for key generation");
  }
  return rng_.get();
}

// TEST ONLY: a reproducible key stream.  Anyone who knows the seed can regenerate every key.
void Context::rng_seed_for_tests(u64 seed) {
  if (!rng_) rng_ = std::make_shared<RngState>();
  const u32 k0[8] = {(u32)seed, (u32)(seed >> 32), 0x6b657973u, 0x65656421u, 0, 0, 0, 0};
  u32 w[16];
  chacha20_block(k0, 0, 0, 0, 0, w);
  for (int i = 0; i < 8; i++) rng_->key.k[i] = w[i];
  rng_->reference = false;
}

// the reference's generators from pinned seeds (what the test harness pins on the reference side)
void Context::rng_pin_reference(const u32* seed16, u64 counter, u32 tri_base) {
  if (!rng_) rng_ = std::make_shared<RngState>();
  rng_->reference = true;
  rng_->prng.pin(seed16, counter);
  rng_->tri_base = tri_base;
  rng_->tri_calls = 0;
  rng_->tri_stream = false;
}
void Context::rng_pin_reference_stream(const u32* seed16, u64 counter, u32 srandom_seed, const u64* tri_pos,
                                       size_t n_pos) {
  rng_pin_reference(seed16, counter, 0);
  rng_->tri_stream = true;
  rng_->tri_gen.srandom(srandom_seed);
  rng_->tri_at = 0;
  rng_->tri_pos.assign(tri_pos, tri_pos + n_pos);
}

// ---- the two samplers behind every key / encryption ------------------------------------------
// n_limbs uniform limbs (moduli g0 ..) at dst: Sample_uniform_poly (polynomial.c:1349-1372)
void Context::sample_uniform(u64* dst, u32 g0, u32 n_limbs, u32 digit, u64 id) {
  RngState* R = rng();
  LimbBatch b = all_limbs(dst, g0, n_limbs);
  if (!R->reference) {
    uniform_kernel<<<grid_for(N, n_limbs), 256, 0, stream>>>(T, b, R->key, digit, id);
    return;
  }
  R->host.resize((size_t)n_limbs * N);
  for (u32 l = 0; l < n_limbs; l++) R->prng.sample_uniform(R->host.data() + (size_t)l * N, N, mod[g0 + l]);
  h2d_sync(dst, R->host.data(), (size_t)n_limbs * N * sizeof(u64));
}

// one triangle polynomial, its residues written to n_limbs limbs at dst (coefficient form):
// Sample_triangle + Transform_values_to_qbase / _qpbase (ckks_key_generator.c:93-96, 173-174)
void Context::sample_triangle(u64* dst, u32 g0, u32 n_limbs, u32 purpose, u32 digit, u64 id) {
  RngState* R = rng();
  LimbBatch b = all_limbs(dst, g0, n_limbs);
  if (!R->reference) {
    triangle_kernel<<<grid_for(N, n_limbs), 256, 0, stream>>>(T, b, R->key, purpose, digit, id);
    return;
  }
  R->host.resize(N);
  if (R->tri_stream) {
    const u32 k = R->tri_calls++;
    if (k >= R->tri_pos.size() || R->tri_pos[k] < R->tri_at)
      throw std::runtime_error("reference stream: no (or an earlier) position recorded for this Sample_triangle call");
    for (; R->tri_at < R->tri_pos[k]; R->tri_at++) R->tri_gen.random();
    R->tri_gen.sample_triangle(R->host.data(), N);
    R->tri_at += N;
  } else {
    refrng::GlibcRandom g;
    g.srandom(R->tri_base + R->tri_calls++);
    g.sample_triangle(R->host.data(), N);
  }
  small_to_rns(dst, g0, n_limbs, R->host.data());
}

void Context::small_to_rns(u64* dst, u32 g0, u32 n_limbs, const int64_t* host_small) {
  int64_t* dv = nullptr;
  ACE_CUDA(cudaMallocAsync(&dv, N * sizeof(int64_t), stream));
  h2d_sync(dv, host_small, N * sizeof(int64_t));
  LimbBatch b = all_limbs(dst, g0, n_limbs);
  small_to_rns_kernel<<<grid_for(N, n_limbs), 256, 0, stream>>>(T, b, dv);
  ACE_CUDA(cudaFreeAsync(dv, stream));
}

// Generate_secret_key (ckks_key_generator.c:69-82): ternary s (Sample_ternary_poly), NTT
void Context::gen_secret_key() {
  RngState* R = rng();
  std::vector<int64_t> s(N, 0);
  size_t hw = params.hamming_weight;
  if (R->reference) {
    R->prng.sample_ternary(s.data(), N, (int64_t)hw);
  } else {
    // host-side draw from the ChaCha20 stream (purpose 0): same distribution as Sample_ternary
    u32 w[16];
    u32 blk = 0, at = 16;
    auto next = [&]() -> u32 {
      if (at == 16) { chacha20_block(R->key.k, blk++, 0, 0x736b, 0, w); at = 0; }
      return w[at++];
    };
    auto below = [&](u32 n) -> u32 {  // unbiased
      const u32 lim = 0xFFFFFFFFu - 0xFFFFFFFFu % n;
      u32 x;
      do x = next(); while (x >= lim);
      return x % n;
    };
    if (hw == 0) {  // uniform ternary (random_sample.c:127-131)
      for (auto& x : s) x = (int64_t)below(3) - 1;
    } else {
      if (hw > N) hw = N;
      size_t placed = 0;
      while (placed < hw) {
        const size_t idx = below(N);
        if (s[idx] == 0) { s[idx] = (next() & 1) ? 1 : -1; placed++; }
      }
    }
  }
  if (!sk_ntt) ACE_CUDA(cudaMalloc(&sk_ntt, G * (size_t)N * sizeof(u64)));
  small_to_rns(sk_ntt, 0, (u32)G, s.data());
  ntt(sk_ntt, 0, (u32)G);
  sync();
}

void Context::import_secret_key(const u64* host_ntt_qp) {
  if (!sk_ntt) ACE_CUDA(cudaMalloc(&sk_ntt, G * (size_t)N * sizeof(u64)));
  h2d_sync(sk_ntt, host_ntt_qp, G * (size_t)N * sizeof(u64));
}

void Context::import_public_key(const u64* h0, const u64* h1) {
  if (!pk0) ACE_CUDA(cudaMalloc(&pk0, L * (size_t)N * sizeof(u64)));
  if (!pk1) ACE_CUDA(cudaMalloc(&pk1, L * (size_t)N * sizeof(u64)));
  h2d_sync(pk0, h0, L * (size_t)N * sizeof(u64));
  h2d_sync(pk1, h1, L * (size_t)N * sizeof(u64));
}

// pk = (-a s + e, a) over Q  (ckks_key_generator.c:84-125; e is drawn before a)
void Context::gen_public_key() {
  if (!sk_ntt) throw std::runtime_error("secret key missing");
  if (!pk0) ACE_CUDA(cudaMalloc(&pk0, L * (size_t)N * sizeof(u64)));
  if (!pk1) ACE_CUDA(cudaMalloc(&pk1, L * (size_t)N * sizeof(u64)));
  u64* e = alloc_limbs(L, false);
  sample_triangle(e, 0, (u32)L, kPurposeError, 0, /*id*/ 1);
  sample_uniform(pk1, 0, (u32)L, 0, /*id*/ 1);
  ntt(e, 0, (u32)L);
  launch_ew(T, EW_MUL, pk0, pk1, sk_ntt, 0, (u32)L, stream);
  launch_ew(T, EW_SUB, pk0, e, pk0, 0, (u32)L, stream);
  free_limbs(e);
  sync();
}

// Generate_switching_key (ckks_key_generator.c:127-200): for every digit j
//   a_j uniform over Q u P,  b_j = e_j + [P]_q * new_key (digit limbs only) - a_j * old_key
// id: what the key is for (2 = relinearisation, 2^32 + automorphism index = rotation)
void Context::gen_switch_key(SwitchKey& key, const u64* new_key, const u64* old_key, u64 id) {
  const size_t per = G * (size_t)N;
  if (!key.k0) ACE_CUDA(cudaMalloc(&key.k0, dnum * per * sizeof(u64)));
  if (!key.k1) ACE_CUDA(cudaMalloc(&key.k1, dnum * per * sizeof(u64)));
  u64* e = alloc_limbs(G, false);
  std::vector<u64> fac(G);
  u64* dfac = nullptr;
  ACE_CUDA(cudaMalloc(&dfac, G * sizeof(u64)));
  for (size_t j = 0; j < dnum; j++) {
    u64* a = key.k1 + j * per;
    u64* b = key.k0 + j * per;
    LimbBatch bb = all_limbs(b, 0, (u32)G);
    sample_uniform(a, 0, (u32)G, (u32)j, id);
    sample_triangle(e, 0, (u32)G, kPurposeError, (u32)j, id);
    ntt(e, 0, (u32)G);
    for (size_t g = 0; g < G; g++) {
      fac[g] = 0;
      if (g < L && g >= j * part_size && g < (j + 1) * part_size) {
        u64 pm = 1;  // P mod q_g  (Get_pmodq)
        for (size_t k = 0; k < K; k++) pm = hm::mulmod(pm, mod[L + k] % mod[g], mod[g]);
        fac[g] = pm;
      }
    }
    ACE_CUDA(cudaMemcpyAsync(dfac, fac.data(), G * sizeof(u64), cudaMemcpyHostToDevice, stream));
    swk_b_kernel<<<grid_for(N, (u32)G), 256, 0, stream>>>(T, bb, a, e, new_key, old_key, dfac);
    sync();
  }
  free_limbs(e);
  sync();
  cudaFree(dfac);
}

void Context::gen_relin_key() {  // new = s^2, old = s (ckks_key_generator.c:203-215)
  u64* s2 = alloc_limbs(G, false);
  launch_ew(T, EW_MUL, s2, sk_ntt, sk_ntt, 0, (u32)G, stream);
  gen_switch_key(relin_key, s2, sk_ntt, /*id*/ 2);
  free_limbs(s2);
}

static u64 inv_mod_pow2(u64 a, u64 M) {  // odd a, M a power of two
  u64 x = 1;
  for (int i = 0; i < 7; i++) x = x * (2 - a * x);
  return x & (M - 1);
}

// "fast" rotation key for automorphism index k (ckks_key_generator.c:237-264; the conjugation
// key of Generate_conj_key, :217-235, is k = 2N - 1):
// old key = sigma_{k^-1}(s), new key = s, so that the automorphism is applied after the switch
void Context::gen_auto_key(u32 k) {
  const u64 M = 2 * (u64)N;
  u32 kinv = (u32)inv_mod_pow2(k, M);
  const int64_t* order = auto_order(kinv);
  u64* rot = alloc_limbs(G, false);
  launch_gather(T, rot, sk_ntt, order, 0, (u32)G, stream);
  gen_switch_key(rot_key(k), sk_ntt, rot, ((u64)1 << 32) | k);
  free_limbs(rot);
}

// Alloc_ckks_key_generator (ckks_key_generator.c:13-37): secret, public, relinearisation key,
// then the rotation keys in the order given (Generate_rot_maps skips repeated indices)
void Context::keygen(u64 seed, const int32_t* rots, size_t n_rots) {
  ACE_CUDA(cudaSetDevice(device));
  if (seed != 0) rng_seed_for_tests(seed);  // 0: the operating system's entropy
  gen_secret_key();
  gen_public_key();
  gen_relin_key();
  keygen_rotations(rots, n_rots);
}

// Generate_rot_maps (ckks_key_generator.c:288-336).  The reference remembers ROTATION VALUES, not
// automorphism indices: a second rotation value with the same automorphism (k and k + N/2, e.g.
// -2 and slots-2 of the bootstrap's list) generates the key again and replaces the first one.
// In reference mode this is reproduced (the streams must advance identically); otherwise a key
// that exists is kept.
void Context::keygen_rotations(const int32_t* rots, size_t n_rots) {
  const bool ref_mode = rng()->reference;
  for (size_t i = 0; i < n_rots; i++) {
    const u32 k = auto_index(rots[i]);
    if (ref_mode) {
      if (ref_rot_seen_.insert(rots[i]).second) gen_auto_key(k);
    } else if (!has_rot_key(k)) {
      gen_auto_key(k);
    }
  }
}

void Context::keygen_reference(const u32* seed16, u64 counter, u32 tri_base, const int32_t* rots,
                               size_t n_rots) {
  ACE_CUDA(cudaSetDevice(device));
  rng_pin_reference(seed16, counter, tri_base);
  ref_rot_seen_.clear();
  gen_secret_key();
  gen_public_key();
  gen_relin_key();
  keygen_rotations(rots, n_rots);
}

void Context::keygen_reference_stream(const u32* seed16, u64 counter, u32 srandom_seed, const u64* tri_pos,
                                      size_t n_pos, const int32_t* rots, size_t n_rots) {
  ACE_CUDA(cudaSetDevice(device));
  rng_pin_reference_stream(seed16, counter, srandom_seed, tri_pos, n_pos);
  ref_rot_seen_.clear();
  gen_secret_key();
  gen_public_key();
  gen_relin_key();
  keygen_rotations(rots, n_rots);
}

// Encrypt_msg (ckks_encryptor.c:20-95): c0 = pk0 u + e1 + m, c1 = pk1 u + e2; u, e1, e2 are
// three triangle draws in this order.  id: unique per encryption (the caller counts).
void Context::encrypt(u64* c0, u64* c1, const u64* pt, u32 level, u64 id) {
  if (!pk0) throw std::runtime_error("public key missing");
  u64* t = alloc_limbs(3 * (size_t)level, false);
  u64 *u = t, *e1 = t + (size_t)level * N, *e2 = t + 2 * (size_t)level * N;
  sample_triangle(u, 0, level, kPurposeEncU, 0, id);
  sample_triangle(e1, 0, level, kPurposeEncE1, 0, id);
  sample_triangle(e2, 0, level, kPurposeEncE2, 0, id);
  ntt(t, 0, level);
  ntt(e1, 0, level);
  ntt(e2, 0, level);
  launch_ew(T, EW_MUL, c0, pk0, u, 0, level, stream);
  launch_ew(T, EW_ADD, c0, c0, e1, 0, level, stream);
  launch_ew(T, EW_ADD, c0, c0, pt, 0, level, stream);
  launch_ew(T, EW_MUL, c1, pk1, u, 0, level, stream);
  launch_ew(T, EW_ADD, c1, c1, e2, 0, level, stream);
  free_limbs(t);
}

// Decrypt (ckks_decryptor.c:19-65): m = c0 + c1 s
void Context::decrypt(u64* pt, const u64* c0, const u64* c1, u32 level) {
  if (!sk_ntt) throw std::runtime_error("secret key missing");
  launch_ew(T, EW_MUL, pt, c1, sk_ntt, 0, level, stream);
  launch_ew(T, EW_ADD, pt, pt, c0, 0, level, stream);
}

// ------------------------------------------------------------------------------ encode
// FP64 arithmetic below must round exactly like the reference's x86-64 build (no FMA
// contraction): every operation is an explicit _rn intrinsic.
struct cplx { double re, im; };

// Embedding_inv (ntt.c:713-753): log2(slots) Gentleman-Sande stages (a, b) -> (a + b, (a - b) w),
// w = rou[(idx_mod - rot_group[i] % idx_mod) * gap] folded into tw[num2 - 1 + i].  Each
// butterfly performs exactly the reference's operations (unfused _rn arithmetic), so fusing
// stages into two launches does not change a single bit:
//   emb_inv_strided<SA>: the top SA = log2(slots) - 12 stages in registers (a thread owns the
//                        2^SA elements of one column at stride 4096), input conversion fused in;
//   emb_inv_tile:        the remaining <= 12 stages on a contiguous tile in shared memory.
enum SrcKind { SRC_F32 = 0, SRC_F64 = 1, SRC_C64 = 2 };

__device__ __forceinline__ cplx emb_load(const void* src, int kind, u32 idx, u32 len) {
  if (idx >= len) return cplx{0.0, 0.0};
  if (kind == SRC_F32) return cplx{(double)reinterpret_cast<const float*>(src)[idx], 0.0};
  if (kind == SRC_F64) return cplx{reinterpret_cast<const double*>(src)[idx], 0.0};
  return reinterpret_cast<const cplx*>(src)[idx];
}

__device__ __forceinline__ void emb_butterfly(cplx& a, cplx& b, const cplx w) {
  cplx s, d;
  s.re = __dadd_rn(a.re, b.re);
  s.im = __dadd_rn(a.im, b.im);
  d.re = __dsub_rn(a.re, b.re);
  d.im = __dsub_rn(a.im, b.im);
  a    = s;
  b.re = __dsub_rn(__dmul_rn(d.re, w.re), __dmul_rn(d.im, w.im));
  b.im = __dadd_rn(__dmul_rn(d.re, w.im), __dmul_rn(d.im, w.re));
}

constexpr u32 kEmbTileLog = 12, kEmbTile = 1u << kEmbTileLog;

// Messages of one batch: job j reads len[j] values of type kind[j] from src[j]; its N/2-slot
// work vector is v + j * slots.
constexpr int kMaxEnc = 64;
struct EncSrcBatch {
  u32         n;
  const void* src[kMaxEnc];
  u32         len[kMaxEnc];
  uint8_t     kind[kMaxEnc];
};
// Output limbs of one batch: limb y belongs to job[y], uses modulus g[y], is multiplied by
// pw[y] (Delta^(sf_degree-1) mod q, 0 = no factor) and written to out[y].
constexpr int kMaxEncLimbs = 160;
struct EncLimbBatch {
  u32      n;
  u64*     out[kMaxEncLimbs];
  u64      pw[kMaxEncLimbs];
  uint16_t g[kMaxEncLimbs];
  uint16_t job[kMaxEncLimbs];
};

template <int SA>
__global__ void __launch_bounds__(128) emb_inv_strided(cplx* __restrict__ vall,
                                                       const cplx* __restrict__ tw,
                                                       const __grid_constant__ EncSrcBatch B,
                                                       u32 logslots) {
  constexpr int R = 1 << SA;
  const u32 stride = kEmbTile, col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= stride) return;
  const u32   job  = blockIdx.y;
  cplx*       v    = vall + ((size_t)job << logslots);
  const void* src  = B.src[job];
  const int   kind = B.kind[job];
  const u32   len  = B.len[job];
  cplx x[R];
#pragma unroll
  for (int r = 0; r < R; r++) x[r] = emb_load(src, kind, r * stride + col, len);
#pragma unroll
  for (int s = 0; s < SA; s++) {
    const u32 num2 = 1u << (logslots - s - 1);  // half the butterfly span of this stage
    const int tr   = R >> (s + 1);
#pragma unroll
    for (int p = 0; p < R / 2; p++) {
      const int lo = (p / tr) * 2 * tr + (p % tr);
      const u32 i  = ((u32)lo * stride + col) & (num2 - 1);
      emb_butterfly(x[lo], x[lo + tr], tw[num2 - 1 + i]);
    }
  }
#pragma unroll
  for (int r = 0; r < R; r++) v[r * stride + col] = x[r];
}

// tile = min(slots, 4096) elements; from_src: no strided phase ran, convert the input here
__global__ void __launch_bounds__(512) emb_inv_tile(cplx* __restrict__ vall,
                                                    const cplx* __restrict__ tw,
                                                    const __grid_constant__ EncSrcBatch B,
                                                    u32 logtile, u32 logslots, int from_src) {
  extern __shared__ double emb_sm[];
  cplx* t = reinterpret_cast<cplx*>(emb_sm);
  const u32   job  = blockIdx.y;
  cplx*       v    = vall + ((size_t)job << logslots);
  const void* src  = B.src[job];
  const int   kind = B.kind[job];
  const u32   len  = B.len[job];
  const u32 tile = 1u << logtile, base = blockIdx.x * tile;
  for (u32 e = threadIdx.x; e < tile; e += blockDim.x)
    t[e] = from_src ? emb_load(src, kind, base + e, len) : v[base + e];
  __syncthreads();
  for (u32 logm = logtile; logm > 0; logm--) {
    const u32 num2 = 1u << (logm - 1);
    for (u32 p = threadIdx.x; p < tile / 2; p += blockDim.x) {
      const u32 i  = p & (num2 - 1);
      const u32 lo = ((p >> (logm - 1)) << logm) | i;
      emb_butterfly(t[lo], t[lo + num2], tw[num2 - 1 + i]);
    }
    __syncthreads();
  }
  for (u32 e = threadIdx.x; e < tile; e += blockDim.x) v[base + e] = t[e];
}

// bit-reverse, divide by slots, scale by Delta, round, spread with `gap`, reduce into every limb
// (ckks_encoder.c:247-263 + polynomial.c:362-392)
__global__ void __launch_bounds__(256) emb_round_rns_kernel(DeviceTables T,
                                                            const __grid_constant__ EncLimbBatch B,
                                                            const cplx* __restrict__ vall,
                                                            u32 slots, u32 logslots, double delta) {
  const u32     limb = blockIdx.y;
  const Modulus m    = T.mod[B.g[limb]];
  u64*          out  = B.out[limb];
  const cplx*   v    = vall + ((size_t)B.job[limb] << logslots);
  const u32     gap  = T.N / (2 * slots);
  const u64     pw   = B.pw[limb];
  for (u32 n = blockIdx.x * blockDim.x + threadIdx.x; n < T.N; n += gridDim.x * blockDim.x) {
    u64 res = 0;
    if (n % gap == 0) {
      const u32  k  = n / gap;           // < 2*slots
      const u32  i  = k < slots ? k : k - slots;
      const cplx c  = v[logslots ? (__brev(i) >> (32 - logslots)) : 0];
      double     x  = k < slots ? c.re : c.im;
      x             = __ddiv_rn(x, (double)slots);
      x             = __dadd_rn(__dmul_rn(x, delta), 0.5);
      long long  r  = llround(x);
      // r mod q (Max_64bit_value detour of the reference is the identity for |r| < 2^62)
      long long  q  = (long long)m.q;
      long long  t  = r % q;
      if (t < 0) t += q;
      res = (u64)t;
      if (pw) res = mul_mod(res, pw, m);
    }
    out[n] = res;
  }
}

__global__ void fill_limb_kernel(DeviceTables T, LimbBatch b, const ScalarPack val) {
  const u32 limb = blockIdx.y;
  u64*      out  = b.base + (size_t)b.slot[limb] * T.N;
  const u64 v    = val.v[limb];
  for (u32 n = blockIdx.x * blockDim.x + threadIdx.x; n < T.N; n += gridDim.x * blockDim.x)
    out[n] = v;
}

void Context::init_encoder() {
  if (enc_tw_ && enc_buf_) return;
  if (enc_tw_) {  // worker context: the tables are shared, the scratch buffers are not
    const size_t half = N / 2;
    ACE_CUDA(cudaMalloc(&enc_buf_, half * sizeof(cplx)));
    ACE_CUDA(cudaMalloc(&enc_pow_, G * sizeof(u64)));
    ACE_CUDA(cudaMallocHost(&enc_host_, half * sizeof(cplx)));
    return;
  }
  // Precompute_fft (ntt.c:585-610) with fft_length = 2N, folded per stage:
  //   tw[logm][i] = rou[(idx_mod - rot_group[i] % idx_mod) * gap]
  const size_t M = 2 * (size_t)N, half = N / 2;
  fft_rou_.resize(M);
  for (size_t i = 0; i < M; i++) {
    double angle = 2 * M_PI * i / M;
    fft_rou_[i]  = std::complex<double>(cos(angle), sin(angle));
  }
  rot_group_.resize(half);
  rot_group_[0] = 1;
  for (size_t i = 1; i < half; i++) rot_group_[i] = (5 * rot_group_[i - 1]) % M;
  // stage tables for every slot count share the same layout: entry (num2 - 1 + i)
  std::vector<cplx> tw(half > 0 ? half : 1);
  for (u32 logm = 1; (1u << logm) <= half; logm++) {
    size_t idx_mod = (size_t)1 << (logm + 2), gap = M / idx_mod, num2 = (size_t)1 << (logm - 1);
    for (size_t i = 0; i < num2; i++) {
      auto w = fft_rou_[(idx_mod - (rot_group_[i] % idx_mod)) * gap];
      tw[num2 - 1 + i] = cplx{w.real(), w.imag()};
    }
  }
  cplx* d = nullptr;
  ACE_CUDA(cudaMalloc(&d, tw.size() * sizeof(cplx)));
  h2d_sync(d, tw.data(), tw.size() * sizeof(cplx));
  enc_tw_ = d;
  ACE_CUDA(cudaMalloc(&enc_buf_, half * sizeof(cplx)));
  ACE_CUDA(cudaMalloc(&enc_pow_, G * sizeof(u64)));
  ACE_CUDA(cudaMallocHost(&enc_host_, half * sizeof(cplx)));
}

// Encode_impl (ckks_encoder.c:199-299).  vals: `len` real messages (host), zero padded to slots.
// out: level (+ p_cnt) limbs, NTT form.
void Context::encode(u64* out, const double* vals, size_t len, u32 level, u32 slots,
                     u32 sf_degree, u32 p_cnt) {
  encode_any(out, vals, nullptr, len, level, slots, sf_degree, p_cnt);
}
// complex messages (the C2S / S2C diagonals of the bootstrap, Encode_ext_at_level)
void Context::encode_cplx(u64* out, const std::complex<double>* vals, size_t len, u32 level,
                          u32 slots, u32 sf_degree, u32 p_cnt) {
  encode_any(out, nullptr, vals, len, level, slots, sf_degree, p_cnt);
}
void Context::encode_any(u64* out, const double* vals, const std::complex<double>* cvals,
                         size_t len, u32 level, u32 slots, u32 sf_degree, u32 p_cnt) {
  // host message -> device staging (stream-ordered; a pageable source is staged by the driver
  // before the call returns, so the caller's buffer is free again immediately)
  const size_t bytes = len * (cvals ? sizeof(cplx) : sizeof(double));
  void* d = nullptr;
  ACE_CUDA(cudaMallocAsync(&d, bytes ? bytes : 16, stream));
  if (bytes)
    ACE_CUDA(cudaMemcpyAsync(d, cvals ? (const void*)cvals : (const void*)vals, bytes,
                             cudaMemcpyHostToDevice, stream));
  encode_dev(out, d, cvals ? SRC_C64 : SRC_F64, len, level, slots, sf_degree, p_cnt);
  ACE_CUDA(cudaFreeAsync(d, stream));
}

// Encode_impl on a message that already lives in HBM (kind: 0 float32, 1 float64, 2 complex)
void Context::encode_dev(u64* out, const void* dev_src, int kind, size_t len, u32 level,
                         u32 slots, u32 sf_degree, u32 p_cnt) {
  EncodeJob j{out, dev_src, kind, (u32)len, level, slots, sf_degree, p_cnt};
  encode_batch(&j, 1);
}

// Encode_impl for a group of messages: the special IFFT of all of them in two launches, one
// rounding / RNS launch and one NTT launch over all their output limbs.
void Context::encode_batch(const EncodeJob* jobs_in, size_t n_jobs) {
  if (n_jobs == 0) return;
  init_encoder();
  std::vector<EncodeJob> jobs(jobs_in, jobs_in + n_jobs);
  for (EncodeJob& j : jobs) {
    if (j.slots == 0) j.slots = N / 2;
    if (j.level == 0) j.level = (u32)L;
    if (j.len > j.slots || j.slots > N / 2 || (j.slots & (j.slots - 1)))
      throw std::runtime_error("encode: bad slot count");
    if (j.level > L || j.sf_degree < 1 || j.p_cnt > K) throw std::runtime_error("encode: bad level");
    tr(TR_ENCODE, j.level + j.p_cnt);
  }
  // jobs of one launch share the slot count; emitted programs use N/2 throughout
  std::stable_sort(jobs.begin(), jobs.end(),
                   [](const EncodeJob& a, const EncodeJob& b) { return a.slots < b.slots; });
  const cplx*  tw    = (const cplx*)enc_tw_;
  const double delta = (double)((u64)1 << params.scaling_mod_size);
  prof::Scope  prof_scope_("encode(total)", stream);
  size_t at = 0;
  while (at < jobs.size()) {
    const u32 slots = jobs[at].slots;
    size_t    cnt   = 1;
    while (at + cnt < jobs.size() && jobs[at + cnt].slots == slots && cnt < (size_t)kMaxEnc) cnt++;
    u32 logslots = 0;
    while ((1u << logslots) < slots) logslots++;
    // work vectors: slots complex doubles per job (at most one limb's worth of bytes each)
    cplx* dv = cnt == 1 ? (cplx*)enc_buf_
                        : (cplx*)alloc_limbs((cnt * slots * sizeof(cplx) + (size_t)N * 8 - 1) / ((size_t)N * 8), false);
    EncSrcBatch sb;
    sb.n = (u32)cnt;
    for (size_t k = 0; k < cnt; k++) {
      sb.src[k] = jobs[at + k].src; sb.len[k] = jobs[at + k].len; sb.kind[k] = (uint8_t)jobs[at + k].kind;
    }
    const u32 logtile = logslots < kEmbTileLog ? logslots : kEmbTileLog;
    const int sa = (int)(logslots - logtile);
    {
      prof::Scope ps("encode_fft", stream);
      if (sa > 0) {
        dim3 grid(kEmbTile / 128, (u32)cnt);
        switch (sa) {
          case 1: emb_inv_strided<1><<<grid, 128, 0, stream>>>(dv, tw, sb, logslots); break;
          case 2: emb_inv_strided<2><<<grid, 128, 0, stream>>>(dv, tw, sb, logslots); break;
          case 3: emb_inv_strided<3><<<grid, 128, 0, stream>>>(dv, tw, sb, logslots); break;
          case 4: emb_inv_strided<4><<<grid, 128, 0, stream>>>(dv, tw, sb, logslots); break;
          default: throw std::runtime_error("encode: slot count too large");
        }
        launches++;
      }
      static bool attr = false;
      if (!attr) {
        cudaFuncSetAttribute(emb_inv_tile, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(kEmbTile * sizeof(cplx)));
        attr = true;
      }
      const u32 tile = 1u << logtile;
      emb_inv_tile<<<dim3(slots / tile, (u32)cnt), tile / 2 < 512 ? (tile / 2 ? tile / 2 : 1) : 512,
                     tile * sizeof(cplx), stream>>>(dv, tw, sb, logtile, logslots, sa == 0);
      launches++;
    }
    // rounding + RNS and the forward NTT over the output limbs of all jobs of the group
    EncLimbBatch lb;
    LimbPtrBatch nb;
    lb.n = 0; nb.n = 0;
    auto flush_round = [&] {
      if (lb.n == 0) return;
      prof::Scope ps("encode_round_rns", stream);
      emb_round_rns_kernel<<<grid_for(N, lb.n), 256, 0, stream>>>(T, lb, dv, slots, logslots, delta);
      launches++;
      lb.n = 0;
    };
    auto flush_ntt = [&] {
      if (nb.n == 0) return;
      launch_ntt(T, nb, stream);
      launches += (logN > 12) ? 2 : 1;
      nb.n = 0;
    };
    for (size_t k = 0; k < cnt; k++) {
      const EncodeJob& j = jobs[at + k];
      const u32 nl = j.level + j.p_cnt;
      for (u32 l = 0; l < nl; l++) {
        if (lb.n == (u32)kMaxEncLimbs) { flush_round(); flush_ntt(); }
        const u32 g = l < j.level ? l : (u32)L + (l - j.level);
        u64 pw = 0;
        if (j.sf_degree > 1 && l < j.level) {
          u64 q = mod[l];
          pw = (u64)delta % q;
          for (u32 d = 2; d < j.sf_degree; d++) pw = hm::mulmod(pw, (u64)delta % q, q);
        }
        u64* o = j.out + (size_t)l * N;
        lb.out[lb.n] = o; lb.pw[lb.n] = pw; lb.g[lb.n] = (uint16_t)g; lb.job[lb.n] = (uint16_t)k;
        lb.n++;
        nb.dst[nb.n] = o; nb.src[nb.n] = o; nb.g[nb.n] = (uint16_t)g;
        nb.n++;
      }
    }
    flush_round();
    flush_ntt();
    if (cnt > 1) free_limbs((u64*)dv);
    at += cnt;
  }
}

// Encode_val_at_level (ckks_encoder.c:464-528): constant plaintext, every coefficient of limb
// l equals the same residue (NTT of a constant polynomial... the reference marks it NTT as is)
std::vector<u64> Context::value_residues(double value, u32 level, u32 sf_degree) const {
  if (level == 0) level = (u32)L;
  const double sf = (double)((u64)1 << params.scaling_mod_size);
  const int MAX_BITS_IN_WORD = 61;
  int32_t log_sf     = (int32_t)ceil(log2(fabs(value * sf)));
  int32_t log_valid  = (log_sf <= MAX_BITS_IN_WORD) ? log_sf : MAX_BITS_IN_WORD;
  int32_t log_approx = log_sf - log_valid;
  double  approx_factor = pow(2, log_approx);
  double  scaled = value / approx_factor * sf + 0.5;
  int64_t val    = (int64_t)scaled;
  int64_t sf_deg_scalar = (int64_t)(sf + 0.5);
  std::vector<u64> res(level);
  for (u32 i = 0; i < level; i++) {
    int64_t q = (int64_t)mod[i];
    int64_t r = val % q;
    if (r < 0) r += q;
    for (u32 j = 1; j < sf_degree; j++) r = (int64_t)hm::mulmod((u64)r, (u64)(sf_deg_scalar % q), (u64)q);
    res[i] = (u64)r;
  }
  if (log_approx > 0) {  // Scale_back_up_by_approxfactor (ckks_encoder.c:410-459)
    int32_t log_step = (log_approx <= 60) ? log_approx : 60;  // MAX_LOG_STEP
    std::vector<u64> approx(level);
    int32_t left = log_approx;
    for (u32 i = 0; i < level; i++) approx[i] = ((u64)1 << log_step) % mod[i];
    left -= log_step;
    while (left > 0) {
      log_step = (left <= 60) ? left : 60;
      for (u32 i = 0; i < level; i++)
        approx[i] = hm::mulmod(approx[i], ((u64)1 << log_step) % mod[i], mod[i]);
      left -= log_step;
    }
    for (u32 i = 0; i < level; i++) res[i] = hm::mulmod(res[i], approx[i], mod[i]);
  }
  return res;
}

void Context::encode_value(u64* out, double value, u32 level, u32 sf_degree) {
  init_encoder();
  if (level == 0) level = (u32)L;
  std::vector<u64> res = value_residues(value, level, sf_degree);
  ScalarPack sp;
  for (u32 i = 0; i < level; i++) sp.v[i] = res[i];
  LimbBatch b = all_limbs(out, 0, level);
  fill_limb_kernel<<<grid_for(N, level), 256, 0, stream>>>(T, b, sp);
  launches++;
}

// ------------------------------------------------------------------------------ decode
// Decode (ckks_encoder.c:649-703): INTT, exact CRT reconstruction (centred), truncating
// conversion to double (mpz_get_d), divide by the scale, forward special FFT (Embedding).
void Context::decode(double* out_re, double* out_im, const u64* pt, u32 level, u32 slots,
                     double scale) {
  init_encoder();
  if (slots == 0) slots = N / 2;
  std::vector<u64> h((size_t)level * N);
  u64* coef = alloc_limbs(level, false);
  intt_from(coef, pt, 0, level);
  download(h.data(), coef, level);
  free_limbs(coef);
  const u32 half_n = N / 2, gap = half_n / slots;
  // mixed-radix (Garner) digits give the exact integer x in [0, Q_level)
  std::vector<u64> inv((size_t)level * level, 0);  // inv[j*level+i] = q_i^-1 mod q_j (i<j)
  for (u32 j = 1; j < level; j++)
    for (u32 i = 0; i < j; i++) inv[(size_t)j * level + i] = hm::invmod_prime(mod[i] % mod[j], mod[j]);
  // big modulus and its half as little-endian limbs
  auto mul_small = [](std::vector<u64>& a, u64 m) {
    u64 carry = 0;
    for (auto& w : a) { u128 t = (u128)w * m + carry; w = (u64)t; carry = (u64)(t >> 64); }
    if (carry) a.push_back(carry);
  };
  std::vector<u64> bigM(1, 1);
  for (u32 i = 0; i < level; i++) mul_small(bigM, mod[i]);
  const size_t W = bigM.size() + 1;
  bigM.resize(W, 0);
  std::vector<u64> halfM(W, 0);
  for (size_t k = 0; k < W; k++) halfM[k] = (bigM[k] >> 1) | (k + 1 < W ? bigM[k + 1] << 63 : 0);
  auto cmp = [&](const std::vector<u64>& a, const std::vector<u64>& b) {
    for (size_t k = W; k-- > 0;) {
      if (a[k] != b[k]) return a[k] < b[k] ? -1 : 1;
    }
    return 0;
  };
  auto to_double_trunc = [&](const std::vector<u64>& a) {  // mpz_get_d: truncate toward zero
    size_t top = W;
    while (top > 0 && a[top - 1] == 0) top--;
    if (top == 0) return 0.0;
    int lz = __builtin_clzll(a[top - 1]);
    size_t bits = top * 64 - lz;
    if (bits <= 53) return (double)a[0];
    size_t shift = bits - 53;  // keep the top 53 bits, drop the rest
    u64 mant = 0;
    size_t w = shift / 64, s = shift % 64;
    mant = a[w] >> s;
    if (s && w + 1 < W) mant |= a[w + 1] << (64 - s);
    mant &= (((u64)1 << 53) - 1);
    return ldexp((double)mant, (int)shift);
  };
  std::vector<std::complex<double>> msg(slots);
  std::vector<u64> v(level), x(W), tmp(W);
  for (u32 s = 0; s < 2 * slots; s++) {
    const u32 n = (s < slots) ? s * gap : (s - slots) * gap + half_n;
    for (u32 j = 0; j < level; j++) {
      u64 qj = mod[j];
      u64 t  = h[(size_t)j * N + n] % qj;
      for (u32 i = 0; i < j; i++) {
        u64 vi = v[i] % qj;
        t = hm::mulmod(t >= vi ? t - vi : t + qj - vi, inv[(size_t)j * level + i], qj);
      }
      v[j] = t;
    }
    // x = v0 + q0 (v1 + q1 (v2 + ...))
    std::fill(x.begin(), x.end(), 0);
    for (u32 j = level; j-- > 0;) {
      u64 carry = 0;
      for (size_t k = 0; k < W; k++) {
        u128 t = (u128)x[k] * mod[j] + carry;
        x[k] = (u64)t; carry = (u64)(t >> 64);
      }
      u64 c = v[j];
      for (size_t k = 0; k < W && c; k++) { u64 o = x[k]; x[k] += c; c = x[k] < o ? 1 : 0; }
    }
    double d;
    if (cmp(x, halfM) > 0) {  // x - M, negative
      u64 borrow = 0;
      for (size_t k = 0; k < W; k++) {
        u128 t = (u128)bigM[k] - x[k] - borrow;
        tmp[k] = (u64)t; borrow = (t >> 64) ? 1 : 0;
      }
      d = -to_double_trunc(tmp);
    } else {
      d = to_double_trunc(x);
    }
    d /= scale;
    if (s < slots) msg[s] = std::complex<double>(d, 0);
    else msg[s - slots] = std::complex<double>(msg[s - slots].real(), d);
  }
  // Embedding (ntt.c:672-711)
  const size_t M = 2 * (size_t)N;
  u32 logslots = 0;
  while ((1u << logslots) < slots) logslots++;
  std::vector<std::complex<double>> r(slots);
  for (u32 i = 0; i < slots; i++) r[hm::bit_reverse(i, logslots)] = msg[i];
  for (u32 logm = 1; logm <= logslots; logm++) {
    size_t idx_mod = (size_t)1 << (logm + 2), g = M / idx_mod, num = (size_t)1 << (logm - 1);
    for (size_t j = 0; j < slots; j += ((size_t)1 << logm)) {
      for (size_t i = 0; i < num; i++) {
        auto w  = fft_rou_[(rot_group_[i] % idx_mod) * g];
        auto b  = r[j + i + num];
        // complex product exactly as gcc expands it: (ac - bd) + (ad + bc) i
        std::complex<double> of(w.real() * b.real() - w.imag() * b.imag(),
                                w.real() * b.imag() + w.imag() * b.real());
        auto a  = r[j + i];
        r[j + i]       = std::complex<double>(a.real() + of.real(), a.imag() + of.imag());
        r[j + i + num] = std::complex<double>(a.real() - of.real(), a.imag() - of.imag());
      }
    }
  }
  for (u32 i = 0; i < slots; i++) {
    out_re[i] = r[i].real();
    if (out_im) out_im[i] = r[i].imag();
  }
}

}  // namespace ace
