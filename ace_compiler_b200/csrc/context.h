// context.h -- B200-resident CKKS evaluation context: parameter set, RNS tables in HBM,
// evaluation keys, a stream-ordered limb allocator and the polynomial / ciphertext ops
// that the C-ABI (include/ace_b200.h) exposes.
//
// Plays the role of the reference's global CKKS_CONTEXT (fhe-cmplr/rtlib/ant/src/rtlib/
// context.c:27-86) + CRT_CONTEXT (include/util/crt.h:873-878) for the evaluation path.
#pragma once
#include <cuda_runtime.h>

#include <complex>
#include <map>
#include <set>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "kernels.cuh"

namespace ace {

struct Params {
  u32    degree;
  size_t mul_depth;       // L = mul_depth + 1 Q primes
  size_t first_mod_size;  // bits of q0
  size_t scaling_mod_size;
  size_t num_q_parts;     // dnum
  size_t hamming_weight;
};

struct SwitchKey {
  u64* k0 = nullptr;  // [dnum][L+K][N], b part (Pk0_at)
  u64* k1 = nullptr;  // [dnum][L+K][N], a part (Pk1_at)
};

#define ACE_CUDA(x)                                                                       \
  do {                                                                                    \
    cudaError_t e_ = (x);                                                                 \
    if (e_ != cudaSuccess)                                                                \
      throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) +     \
                               " at " __FILE__ ":" + std::to_string(__LINE__));           \
  } while (0)

// Jobs of the batched polynomial-level primitives (batch.cu): what one reference call
// (Decomp_modup / Mod_down / Rescale / Encode_plain_from_float) asks for.  The scheduler
// (sched.h) collects independent calls and hands them over together, so that each step of the
// primitive is one launch over the limbs of all of them.
struct ModupJob   { u64* out; const u64* digit; u32 num_q, part; bool copy_own = true; };  // digit: the part's own limbs
struct ModdownJob { u64* out; const u64* in; u32 num_q; };            // in: [num_q | K]
struct RescaleJob { u64* out; const u64* in; u32 num_q; };
struct EncodeJob  { u64* out; const void* src; int kind; u32 len, level, slots, sf_degree, p_cnt; };

bool fused_tails();  // batch.cu: Rescale / ModDown steps folded into the forward transform (opt-in)

class Context {
 public:
  Context(const Params& p, int device);
  ~Context();
  // A worker context for another host thread (the reference runs Main_graph from several OpenMP
  // threads, fhe-cmplr/rtlib/ant/dataset/resnet_cifar.main.inc:81): shares every immutable device
  // table and all keys with this context, has its own stream, limb allocator, scratch buffers
  // and counters.  Create after key generation / bootstrap set-up; destroy before this context.
  Context* make_worker();

  // ---- parameter set (host copies)
  Params           params;
  u32              N, logN;
  size_t           L, K, G, dnum, part_size;
  std::vector<u64> mod;  // [G] Q primes then P primes
  std::vector<u64> psi;  // [G] 2N-th roots used for the NTT tables
  int              device;
  cudaStream_t     stream;
  DeviceTables     T;

  // ---- memory: limb arrays in HBM, stream-ordered
  // Blocks are recycled through per-size free lists: every consumer runs on `stream`, so a
  // block handed out again is only touched after the work queued on it before the free.
  u64* alloc_limbs(size_t n_limbs, bool zero);
  void free_limbs(u64* p);
  void trim_cache();  // return every cached block to the driver
  // device memory for tables and keys (plain cudaMalloc).  The stream-ordered pool behind the limb
  // allocator keeps what it once had; when a plain allocation fails, the cached limb blocks and
  // the pool's idle memory are given back and the allocation is tried again.
  void dev_malloc(void** p, size_t bytes);
  template <typename Tp> void dev_malloc(Tp** p, size_t bytes) { dev_malloc(reinterpret_cast<void**>(p), bytes); }
  static constexpr size_t kCacheCapBytes = (size_t)1 << 40;  // idle blocks per context: no static cap
  // memory pressure: an allocation failed somewhere in the process; every context gives its idle
  // blocks back at its next allocator call (free lists are touched by their own thread only)
  void relieve_pressure();
  unsigned pressure_seen_ = 0;
  size_t block_limbs(const u64* p) const {
    auto it = al_.block_limbs.find(p);
    if (it == al_.block_limbs.end()) throw std::runtime_error("unknown limb block");
    return it->second;
  }
  size_t cached_bytes = 0, live_bytes = 0, peak_bytes = 0;
  void upload(u64* dst, const u64* src, size_t n_limbs);
  void download(u64* dst, const u64* src, size_t n_limbs);
  void sync();
  // Host -> device copy that is complete, and ordered before later work on `stream`, when it
  // returns.  (A plain cudaMemcpy from pageable memory may return while the DMA is still in
  // flight on the legacy stream, which the non-blocking context stream does not wait for.)
  void h2d_sync(void* dst, const void* src, size_t bytes) {
    ACE_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
    ACE_CUDA(cudaStreamSynchronize(stream));
  }

  size_t num_decomp(size_t num_q) const {
    size_t n = (num_q + part_size - 1) / part_size;
    return n > dnum ? dnum : n;
  }
  u32 gidx(u32 o, u32 num_q) const { return o < num_q ? o : (u32)L + (o - num_q); }

  // ---- transforms on `n` consecutive limbs starting at modulus g0
  // standalone = false: part of a larger primitive that the op trace already counts
  void ntt(u64* data, u32 g0, u32 n, bool standalone = true);
  void intt(u64* data, u32 g0, u32 n, bool standalone = true);
  void intt_from(u64* dst, const u64* src, u32 g0, u32 n);

  // ---- reference polynomial-level API (a5, a6, a8)
  void decomp_modup(u64* out, const u64* in, u32 num_q, u32 part);
  void modup_from(u64* out, const u64* digit, u32 num_q, u32 part);
  u32  digit_start(u32 part) const { return (u32)(part_size * part); }
  u32  digit_len(u32 num_q, u32 part) const {
    u32 beta = (u32)num_decomp(num_q), st = digit_start(part);
    return part == beta - 1 ? num_q - st : (u32)part_size;
  }
  void mod_down(u64* out, const u64* in, u32 num_q);
  void rescale(u64* out, const u64* in, u32 num_q);
  // the same primitives over many independent polynomials at once (batch.cu)
  void modup_batch(const ModupJob* jobs, size_t n);
  void moddown_batch(const ModdownJob* jobs, size_t n);
  void rescale_batch(const RescaleJob* jobs, size_t n);
  void encode_batch(const EncodeJob* jobs, size_t n);

  // ---- keys and automorphisms
  u32            auto_index(int32_t rot_idx) const;
  const int64_t* auto_order(u32 auto_idx);  // device table, built on first use
  const int64_t* auto_order_inv(u32 auto_idx);  // table of the inverse automorphism (scatter form)
  SwitchKey&     rot_key(u32 auto_idx) {
    std::lock_guard<std::mutex> lk(sh_->mu);
    return sh_->rot_keys[auto_idx];
  }
  bool           has_rot_key(u32 auto_idx) const {
    std::lock_guard<std::mutex> lk(sh_->mu);
    return sh_->rot_keys.count(auto_idx) != 0;
  }
  std::vector<u32> rot_key_indices() const {
    std::lock_guard<std::mutex> lk(sh_->mu);
    std::vector<u32> v;
    for (auto& kv : sh_->rot_keys)
      if (kv.second.k0) v.push_back(kv.first);
    return v;
  }
  SwitchKey      relin_key;
  void           import_key_limbs(SwitchKey& key, u32 part, int which, const u64* host);

  // ---- fused ciphertext-level pipeline
  // hybrid key switch of d (num_q limbs, NTT form); out0 += nothing, optional add0 is added
  // to out0 after ModDown (the c0 of a rotation).
  void key_switch(u64* out0, u64* out1, const u64* d, u32 num_q, const SwitchKey& key,
                  const u64* add0, const u64* add1 = nullptr);
  void modup_all(u64* ext, const u64* d, u32 num_q);
  void ksw_acc(u64* acc0, u64* acc1, const u64* ext, const u64* d, u32 num_q,
               const SwitchKey& key);
  void mod_down_pair(u64* out0, u64* out1, const u64* a0, const u64* a1, u32 num_q,
                     const u64* add0, const u64* add1 = nullptr);
  void ct_rotate(u64* r0, u64* r1, const u64* c0, const u64* c1, u32 num_q, int32_t rot_idx);
  void ct_rotate_hoisted(u64* const* r0, u64* const* r1, const u64* c0, const u64* c1, u32 num_q,
                         const int32_t* rots, size_t n);
  void ct_mul_plain_acc(u64* acc0, u64* acc1, const u64* c0, const u64* c1, const u64* pt, u32 num_q,
                        bool first);
  void ct_mul_relin(u64* r0, u64* r1, const u64* a0, const u64* a1, const u64* b0,
                    const u64* b1, u32 num_q);

  // ---- client side (client.cu): keys, encryption, CKKS encode / decode
  u64* sk_ntt = nullptr;  // [G][N] secret key, NTT form over Q then P
  u64* pk0    = nullptr;  // [L][N]
  u64* pk1    = nullptr;  // [L][N]
  // seed == 0: keys from the operating system's entropy (getrandom); anything else pins a
  // reproducible stream and is for tests only.  Sampling: ChaCha20, domain-separated per key,
  // digit and encryption (client.cu).
  void keygen(u64 seed, const int32_t* rots, size_t n_rots);
  // the reference's own generators (BLAKE2Xb PRNG + glibc rand(), refrng.h) consumed in the
  // reference's order: the keys are bit-identical to the reference's from the same seeds
  void keygen_reference(const u32* seed16, u64 counter, u32 tri_base, const int32_t* rots, size_t n_rots);
  void keygen_reference_stream(const u32* seed16, u64 counter, u32 srandom_seed, const u64* tri_pos, size_t n_pos,
                               const int32_t* rots, size_t n_rots);
  void keygen_rotations(const int32_t* rots, size_t n_rots);
  std::set<int32_t> ref_rot_seen_;  // reference mode: rotation values that have a key (Generate_rot_maps)
  void gen_secret_key();
  void gen_public_key();
  void gen_relin_key();
  void gen_auto_key(u32 auto_idx);
  void gen_switch_key(SwitchKey& key, const u64* new_key, const u64* old_key, u64 id);
  struct RngState;
  std::shared_ptr<RngState> rng_;  // shared with worker contexts
  RngState* rng();
  void rng_seed_for_tests(u64 seed);
  void rng_pin_reference(const u32* seed16, u64 counter, u32 tri_base);
  void rng_pin_reference_stream(const u32* seed16, u64 counter, u32 srandom_seed, const u64* tri_pos, size_t n_pos);
  void sample_uniform(u64* dst, u32 g0, u32 n_limbs, u32 digit, u64 id);
  void sample_triangle(u64* dst, u32 g0, u32 n_limbs, u32 purpose, u32 digit, u64 id);
  void small_to_rns(u64* dst, u32 g0, u32 n_limbs, const int64_t* host_small);
  void import_secret_key(const u64* host_ntt_qp);
  void import_public_key(const u64* host_pk0, const u64* host_pk1);
  void encrypt(u64* c0, u64* c1, const u64* pt, u32 level, u64 id);  // id: unique per encryption
  // key / ciphertext files (keyfile.cu)
  void save_keys(const char* path, bool with_secret);
  void load_keys(const char* path);
  void save_ct(const char* path, const u64* c0, const u64* c1, u32 level, u32 slots, u32 sf_degree,
               double scale);
  void load_ct(const char* path, u64* c0, u64* c1, u32 max_level, u32* level, u32* slots,
               u32* sf_degree, double* scale);
  void decrypt(u64* pt, const u64* c0, const u64* c1, u32 level);
  void encode(u64* out, const double* vals, size_t len, u32 level, u32 slots, u32 sf_degree,
              u32 p_cnt);
  void encode_value(u64* out, double value, u32 level, u32 sf_degree);
  void decode(double* out_re, double* out_im, const u64* pt, u32 level, u32 slots,
              double scale);

  size_t launches = 0;  // kernels launched so far (bench.py reports the delta)

  // Op trace: how many reference-granularity primitives the work so far corresponds to, by
  // class and (where the cost depends on it) by number of Q limbs.  Control flow of emitted
  // programs is data independent, so the trace of one image is a constant of the model; bench.py
  // multiplies it with the reference's unit costs measured on the host for the CPU baseline.
  enum TraceClass { TR_MODUP_DIGIT, TR_MODDOWN_POLY, TR_RESCALE_POLY, TR_ENCODE, TR_LIMB_MUL,
                    TR_LIMB_ADD, TR_LIMB_ROT, TR_LIMB_NTT, TR_CLASSES };
  static constexpr int kTraceLevels = 72;
  uint64_t trace[TR_CLASSES][kTraceLevels] = {};
  void tr(TraceClass c, u32 level, uint64_t n = 1) { trace[c][level < kTraceLevels ? level : 0] += n; }

 private:
  typedef uint16_t u16;
  struct ModUpTab {
    u32              n_in, n_out, start;
    u64 *            hatinv, *hatinv_sh, *hatmod;  // device
    std::vector<u16> g_in, g_out, out_slot;
  };
  const ModUpTab&  modup_tab(u32 num_q, u32 part);
  // conversion-matrix entries as base_conv_kernel reads them (split at bit 30 when every modulus
  // is below 2^60, see kernels.cu)
  u64 pack_hat(u64 h) const {
    return T.small_moduli ? ((h >> 30) << 32) | (h & 0x3FFFFFFFull) : h;
  }
  void             fill_conv_desc(ConvDesc& d, const ModUpTab& t, const u64* x, u64* out);

  // Tables and keys that are built lazily at run time (ModUp tables per (level, digit),
  // automorphism tables, rotation keys of a Bootstrap at a new slot count) live in ONE store that
  // the primary context and all its workers share, behind a mutex: whichever thread needs an
  // entry first builds it, the others find it.  Entries are never removed and node addresses are
  // stable, so references handed out stay valid after the lock is released.  The store frees
  // its device memory when the last context that holds it goes away.
  struct Shared {
    std::mutex                               mu;
    std::map<std::pair<u32, u32>, ModUpTab>  modup_tabs;
    std::unordered_map<u32, int64_t*>        auto_orders;
    std::unordered_map<u32, SwitchKey>       rot_keys;
    std::vector<void*>                       owned;  // device tables behind the entries above
    int                                      device = 0;
    ~Shared();
  };
  std::shared_ptr<Shared>                  sh_;
  std::vector<void*>                       owned_;  // device tables of the primary (freed by it)
  bool   worker_ = false;          // shares tables and keys with a primary context
  // limb allocator: strictly per context (its thread, its stream).  Copying a context for a
  // worker must not even READ the primary's maps -- the primary may be inside alloc_limbs().
  struct AllocState {
    std::unordered_map<size_t, std::vector<u64*>> free_lists;  // by size in limbs
    std::unordered_map<const u64*, size_t>        block_limbs;
    AllocState() {}
    AllocState(const AllocState&) {}             // a copy starts empty
    AllocState& operator=(const AllocState&) { return *this; }
  };
  AllocState al_;
  // ModDown tables
  u64 *phat_inv_, *phat_inv_sh_, *phat_mod_q_;  // [K], [K], [L][K]
  u64 *pinv_mod_q_, *pinv_mod_q_sh_;            // [L]
 public:
  u64 *pmodq_, *pmodq_sh_;                      // [L] P mod q_j
  std::vector<u64> value_residues(double value, u32 level, u32 sf_degree) const;
  void encode_cplx(u64* out, const std::complex<double>* vals, size_t len, u32 level,
                   u32 slots, u32 sf_degree, u32 p_cnt);
  void encode_any(u64* out, const double* vals, const std::complex<double>* cvals, size_t len,
                  u32 level, u32 slots, u32 sf_degree, u32 p_cnt);
  void encode_dev(u64* out, const void* dev_src, int kind, size_t len, u32 level, u32 slots,
                  u32 sf_degree, u32 p_cnt);
 private:
  // Rescale tables, row l (dropping q_l), column i < l
  u64 *qlinv_, *qlinv_sh_, *negqlinv_, *negqlinv_sh_;  // [L][L]

  // encoder state
  void  init_encoder();
  void* enc_tw_   = nullptr;  // device: per-stage special-FFT twiddles
  void* enc_buf_  = nullptr;  // device: N/2 complex doubles
  u64*  enc_pow_  = nullptr;  // device: per-limb scalars
  void* enc_host_ = nullptr;  // pinned staging buffer
  std::vector<std::complex<double>> fft_rou_;
  std::vector<u64>                  rot_group_;

  template <typename Tp>
  Tp* to_device(const std::vector<Tp>& v);
  template <typename Tp>
  Tp* to_device_shared(const std::vector<Tp>& v);
};

}  // namespace ace
