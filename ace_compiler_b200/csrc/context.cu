// context.cu -- host-side setup (primes, roots, CRT constants -> HBM tables) and the
// polynomial / ciphertext level operations built from the kernels in kernels.cu.
//
// Reference routines restated here (paths under fhe-cmplr/rtlib/ant/):
//   parameter set         src/util/ckks_parameters.c:60-97, src/util/crt.c:574-585
//   Q / P primes          src/util/crt.c:46-125, 353-396
//   NTT tables            src/util/ntt.c:80-127
//   base-conversion and rescale constants  src/util/crt.c:206-533
//   Decompose_modup / Reduce_rns_base / Rescale_poly  src/util/polynomial.c:1241-1335, 928-967, 1097-1161
//   emitted Rotate()/Relinearize()  dataset/resnet20_cifar10_pre.onnx.inc:6972-7146
#include "context.h"
#include "prof.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <cstring>

#include "host_math.h"
#include "rou_table.h"

namespace ace {

static const size_t kAuxBits = 60;  // AUXBITS, include/util/fhe_types.h:27-29

template <typename Tp>
Tp* Context::to_device(const std::vector<Tp>& v) {
  Tp* d = nullptr;
  dev_malloc(&d, v.size() * sizeof(Tp) + 16);
  h2d_sync(d, v.data(), v.size() * sizeof(Tp));
  owned_.push_back(d);
  return d;
}

template <typename Tp>
Tp* Context::to_device_shared(const std::vector<Tp>& v) {  // caller holds sh_->mu
  Tp* d = nullptr;
  dev_malloc(&d, v.size() * sizeof(Tp) + 16);
  h2d_sync(d, v.data(), v.size() * sizeof(Tp));
  sh_->owned.push_back(d);
  return d;
}

Context::Context(const Params& p, int dev) : params(p), device(dev) {
  sh_ = std::make_shared<Shared>();
  sh_->device = dev;
  if (p.degree < 16 || (p.degree & (p.degree - 1)) || p.degree > (1u << 17))
    throw std::runtime_error("degree must be a power of two in [16, 2^17]");
  if (p.num_q_parts == 0) throw std::runtime_error("num_q_parts must be > 0");
  N    = p.degree;
  logN = 0;
  while ((1u << logN) < N) logN++;
  L         = p.mul_depth + 1;
  dnum      = p.num_q_parts;
  part_size = (L + dnum - 1) / dnum;  // ceil(L / dnum), crt.c:356
  if (L <= part_size * (dnum - 1)) throw std::runtime_error("invalid number of q parts");

  // ---- primes
  std::vector<u64> q = hm::q_chain(L, p.first_mod_size, p.scaling_mod_size, N);
  size_t max_bits = 0;
  for (size_t j = 0; j < dnum; j++) {
    size_t lo = j * part_size, hi = std::min(L, lo + part_size);
    max_bits  = std::max(max_bits, hm::product_bit_length(q.data() + lo, hi - lo));
  }
  K = (max_bits + kAuxBits - 1) / kAuxBits;  // crt.c:396
  G = L + K;
  if (G > 64 || part_size > 48 || K > 48) throw std::runtime_error("parameter set too large");
  mod = q;
  u64 prev = hm::first_prime(N, kAuxBits);  // crt.c:56-75
  for (size_t i = 0; i < K; i++) {
    u64 cand;
    bool dup;
    do {
      cand = hm::previous_prime(prev, 2 * (u64)N);
      dup  = false;
      for (size_t j = 0; j < L; j++) dup |= (cand == q[j]);
      prev = cand;
    } while (dup);
    mod.push_back(cand);
  }

  ACE_CUDA(cudaSetDevice(device));
  ACE_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  {  // keep freed limbs cached in the pool instead of returning them to the driver
    cudaMemPool_t pool;
    ACE_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thr = UINT64_MAX;
    ACE_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  }

  // ---- per-modulus constants + NTT tables
  std::vector<Modulus> mods(G);
  std::vector<u64>     tw(G * (size_t)N), tw_sh(G * (size_t)N), itw(G * (size_t)N),
      itw_sh(G * (size_t)N), n_inv(G), n_inv_sh(G);
  psi.resize(G);
  for (size_t g = 0; g < G; g++) {
    const u64 m = mod[g];
    Modulus&  M = mods[g];
    M.q         = m;
    // floor(2^128 / m) = floor((2^128 - 1) / m) because m is odd and > 1
    u128 mu = ~(u128)0 / m;
    M.mu_hi = (u64)(mu >> 64);
    M.mu_lo = (u64)mu;
    u32 nbits = 64 - (u32)__builtin_clzll(m);
    M.shift   = nbits - 2;
    M.mu64    = (u64)((((u128)1) << (62 + nbits)) / m);
    M.pad     = 0;
    M.c64     = (u64)((((u128)1) << 64) % m);
    M.c64_sh  = hm::shoup(M.c64, m);
    // psi = g^((q-1)/2N) with g the smallest generator (number_theory.c:132-157)
    // ... unless the reference's fixed-root table has an entry (fhe_std_parms.c:200-271)
    psi[g] = fixed_root(2 * (u64)N, m);
    if (psi[g] == 0) psi[g] = hm::powmod(hm::smallest_generator(m), (m - 1) / (2 * (u64)N), m);
    u64 psi_inv = hm::invmod_prime(psi[g], m), pw = 1, ipw = 1;
    u64* t  = tw.data() + g * (size_t)N;
    u64* ts = tw_sh.data() + g * (size_t)N;
    u64* it = itw.data() + g * (size_t)N;
    u64* is = itw_sh.data() + g * (size_t)N;
    for (u32 i = 0; i < N; i++) {
      u32 r = hm::bit_reverse(i, logN);
      t[r]  = pw;
      it[r] = ipw;
      pw    = hm::mulmod(pw, psi[g], m);
      ipw   = hm::mulmod(ipw, psi_inv, m);
    }
    for (u32 i = 0; i < N; i++) {
      ts[i] = hm::shoup(t[i], m);
      is[i] = hm::shoup(it[i], m);
    }
    n_inv[g]    = hm::invmod_prime(N % m, m);
    n_inv_sh[g] = hm::shoup(n_inv[g], m);
  }
  T.N = N; T.logN = logN; T.G = (u32)G;
  T.small_moduli = 1;
  for (size_t g = 0; g < G; g++)
    if (mod[g] >> 60) T.small_moduli = 0;
  T.mod    = to_device(mods);
  {  // the four twiddle tables in one block (one L2 access-policy window can cover them)
    const size_t per = G * (size_t)N;
    u64* blk = nullptr;
    dev_malloc(&blk, 4 * per * sizeof(u64));
    owned_.push_back(blk);
    h2d_sync(blk, tw.data(), per * sizeof(u64));
    h2d_sync(blk + per, tw_sh.data(), per * sizeof(u64));
    h2d_sync(blk + 2 * per, itw.data(), per * sizeof(u64));
    h2d_sync(blk + 3 * per, itw_sh.data(), per * sizeof(u64));
    T.tw = blk; T.tw_sh = blk + per; T.itw = blk + 2 * per; T.itw_sh = blk + 3 * per;
    // (Pinning the tables in the persisting part of the L2 with an access-policy window was
    // measured and rejected: key switch 553 -> 583 us, ResNet-20 1.12 -> 1.24 s/image -- the
    // ciphertext intermediates that flow from one kernel to the next need the L2 more.)
  }
  T.n_inv    = to_device(n_inv);
  T.n_inv_sh = to_device(n_inv_sh);
  T.ftw2 = T.itw2 = T.ips2 = nullptr;
  T.ftwd = T.itwd = T.ipsd = nullptr;
  T.fp64_max_q = 0;
  {  // N = 2^16 (ntt16.cu): interleaved {w, w'} tables; moduli must be below 2^61
    bool ok = logN == 16 && getenv("ACE_B200_OLD_NTT") == nullptr;
    for (size_t g = 0; g < G; g++) ok = ok && (mod[g] >> 61) == 0;
    if (ok) {
      const size_t per = G * (size_t)N;  // entries of 2 words
      std::vector<u64> f2(2 * per), i2(2 * per), p2(2 * per);
      for (size_t g = 0; g < G; g++) {
        const u64 m = mod[g];
        u64* f = f2.data() + 2 * g * (size_t)N;
        u64* iv = i2.data() + 2 * g * (size_t)N;
        u64* ps = p2.data() + 2 * g * (size_t)N;
        for (u32 i = 0; i < N; i++) {
          f[2 * i]     = tw[g * (size_t)N + i];
          f[2 * i + 1] = tw_sh[g * (size_t)N + i];
        }
        // decimation-in-time inverse: [mm + j] = omega_(2mm)^-j, omega = psi^2
        const u64 psi_inv = hm::invmod_prime(psi[g], m);
        const u64 om_inv  = hm::mulmod(psi_inv, psi_inv, m);
        iv[0] = iv[1] = 0;
        for (u32 mm = 1; mm < N; mm <<= 1) {
          const u64 base = hm::powmod(om_inv, N / (2 * mm), m);
          u64 pw = 1;
          for (u32 j = 0; j < mm; j++) {
            iv[2 * (mm + j)]     = pw;
            iv[2 * (mm + j) + 1] = hm::shoup(pw, m);
            pw = hm::mulmod(pw, base, m);
          }
        }
        u64 pw = n_inv[g];  // psi^-n N^-1
        for (u32 i = 0; i < N; i++) {
          ps[2 * i]     = pw;
          ps[2 * i + 1] = hm::shoup(pw, m);
          pw = hm::mulmod(pw, psi_inv, m);
        }
      }
      u64* blk = nullptr;
      dev_malloc(&blk, 6 * per * sizeof(u64));
      owned_.push_back(blk);
      h2d_sync(blk, f2.data(), 2 * per * sizeof(u64));
      h2d_sync(blk + 2 * per, i2.data(), 2 * per * sizeof(u64));
      h2d_sync(blk + 4 * per, p2.data(), 2 * per * sizeof(u64));
      T.ftw2 = reinterpret_cast<const ulonglong2*>(blk);
      T.itw2 = reinterpret_cast<const ulonglong2*>(blk + 2 * per);
      T.ips2 = reinterpret_cast<const ulonglong2*>(blk + 4 * per);
      // FP64 butterfly (ntt16.cu): the same values as doubles for the moduli below 1.5e15
      // (1.5 q < 2^51); other rows stay zero and are never read
      if (getenv("ACE_B200_NO_FP64_NTT") == nullptr) {
        const u64 max_q = 1500000000000000ull;
        bool any = false;
        std::vector<double> fd(per, 0.0), id(per, 0.0), pd(per, 0.0);
        for (size_t g = 0; g < G; g++) {
          if (mod[g] >= max_q) continue;
          any = true;
          for (size_t i = 0; i < N; i++) {
            fd[g * (size_t)N + i] = (double)f2[2 * (g * (size_t)N + i)];
            id[g * (size_t)N + i] = (double)i2[2 * (g * (size_t)N + i)];
            pd[g * (size_t)N + i] = (double)p2[2 * (g * (size_t)N + i)];
          }
        }
        if (any) {
          double* dblk = nullptr;
          dev_malloc(&dblk, 3 * per * sizeof(double));
          owned_.push_back(reinterpret_cast<u64*>(dblk));
          h2d_sync(dblk, fd.data(), per * sizeof(double));
          h2d_sync(dblk + per, id.data(), per * sizeof(double));
          h2d_sync(dblk + 2 * per, pd.data(), per * sizeof(double));
          T.ftwd = dblk; T.itwd = dblk + per; T.ipsd = dblk + 2 * per;
          T.fp64_max_q = max_q;
        }
      }
    }
  }

  // ---- ModDown constants (crt.c Precompute_primes(p) + Precompute_new_base(p, q))
  std::vector<u64> phi(K), phi_sh(K), phm(L * K), pinv(L), pinv_sh(L), pm(L), pm_sh(L);
  for (size_t i = 0; i < K; i++) {
    u64 pi = mod[L + i], hat = 1;
    for (size_t k = 0; k < K; k++)
      if (k != i) hat = hm::mulmod(hat, mod[L + k] % pi, pi);
    phi[i]    = hm::invmod_prime(hat, pi);
    phi_sh[i] = hm::shoup(phi[i], pi);
  }
  for (size_t j = 0; j < L; j++) {
    u64 qj = mod[j], prod = 1;
    for (size_t i = 0; i < K; i++) {
      u64 hat = 1;
      for (size_t k = 0; k < K; k++)
        if (k != i) hat = hm::mulmod(hat, mod[L + k] % qj, qj);
      phm[j * K + i] = hat;  // packed below, once small_moduli is known
      prod           = hm::mulmod(prod, mod[L + i] % qj, qj);
    }
    pinv[j]    = hm::invmod_prime(prod, qj);
    pinv_sh[j] = hm::shoup(pinv[j], qj);
    pm[j]      = prod;  // P mod q_j (Get_pmodq, crt.h:697-700)
    pm_sh[j]   = hm::shoup(prod, qj);
  }
  phat_inv_      = to_device(phi);
  phat_inv_sh_   = to_device(phi_sh);
  for (u64& h : phm) h = pack_hat(h);
  phat_mod_q_    = to_device(phm);
  pinv_mod_q_    = to_device(pinv);
  pinv_mod_q_sh_ = to_device(pinv_sh);
  // P mod q_j lifts a Q-basis polynomial into the extended basis (Switch_key_ext and the
  // add_first term of Fast_rotate_ext, ckks_evaluator.c:462-489, 566-573)
  pmodq_    = to_device(pm);
  pmodq_sh_ = to_device(pm_sh);

  // ---- Rescale constants (crt.c:270-330).  _ql_ql_inv_mod_ql_div_ql_mod_qi equals
  // -q_l^-1 mod q_i: (Q/q_l)*[(Q/q_l)^-1]_{q_l} = 1 + k*q_l and is 0 mod q_i, so the stored
  // floor quotient k satisfies k*q_l = -1 (mod q_i).
  std::vector<u64> qi(L * L, 0), qi_sh(L * L, 0), nq(L * L, 0), nq_sh(L * L, 0);
  for (size_t l = 1; l < L; l++) {
    for (size_t i = 0; i < l; i++) {
      u64 m = mod[i], inv = hm::invmod_prime(mod[l] % m, m);
      qi[l * L + i]    = inv;
      qi_sh[l * L + i] = hm::shoup(inv, m);
      nq[l * L + i]    = m - inv;
      nq_sh[l * L + i] = hm::shoup(m - inv, m);
    }
  }
  qlinv_       = to_device(qi);
  qlinv_sh_    = to_device(qi_sh);
  negqlinv_    = to_device(nq);
  negqlinv_sh_ = to_device(nq_sh);
}

Context::Shared::~Shared() {
  cudaSetDevice(device);
  for (auto& kv : rot_keys) { cudaFree(kv.second.k0); cudaFree(kv.second.k1); }
  for (auto& kv : auto_orders) cudaFree(kv.second);
  for (void* p : owned) cudaFree(p);
}

// contexts that run images next to the primary one (kernels.cuh: pdl_chain_enabled)
std::atomic<int> g_worker_contexts{0};

Context::~Context() {
  if (worker_) g_worker_contexts--;
  cudaSetDevice(device);
  cudaStreamSynchronize(stream);
  if (!worker_) {
    cudaFree(relin_key.k0);
    cudaFree(relin_key.k1);
    cudaFree(sk_ntt); cudaFree(pk0); cudaFree(pk1);
    cudaFree(enc_tw_);
  }
  cudaFree(enc_buf_); cudaFree(enc_pow_);
  if (enc_host_) cudaFreeHost(enc_host_);
  if (!worker_)
    for (void* p : owned_) cudaFree(p);
  cudaStreamSynchronize(stream);
  for (auto& kv : al_.block_limbs) cudaFreeAsync(const_cast<u64*>(kv.first), stream);
  cudaStreamSynchronize(stream);
  cudaStreamDestroy(stream);
}

Context* Context::make_worker() {
  ACE_CUDA(cudaSetDevice(device));
  g_worker_contexts++;
  ACE_CUDA(cudaStreamSynchronize(stream));  // everything the worker shares is in place
  // memberwise copy of what is immutable after set-up (tables, keys by pointer) plus the shared
  // store; the allocator state copies as empty (AllocState) and the counters are reset.  The
  // primary may be running: nothing it mutates at run time is read here (its allocator maps are
  // not copied, lazily built entries live behind sh_->mu, counters are overwritten below).
  Context* w = new Context(*this);
  w->worker_ = true;
  w->owned_.clear();
  w->stream = nullptr;
  ACE_CUDA(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking));
  w->cached_bytes = w->live_bytes = w->peak_bytes = 0;
  w->launches = 0;
  memset(w->trace, 0, sizeof(w->trace));
  w->enc_buf_ = nullptr; w->enc_pow_ = nullptr; w->enc_host_ = nullptr;  // scratch: per context
  return w;
}

// ---------------------------------------------------------------------------- memory
static std::atomic<unsigned> g_pressure_epoch{0};
void Context::relieve_pressure() {
  const unsigned e = g_pressure_epoch.load(std::memory_order_relaxed);
  if (e == pressure_seen_) return;
  pressure_seen_ = e;
  trim_cache();
}

// Blocks come in size classes (four per octave: 8, 10, 12, 14, 16, 20, ...), not exact sizes: a
// polynomial of 33 limbs reuses the block a polynomial of 34 limbs left behind.  With exact sizes
// every level of the modulus chain kept its own idle blocks (~30 GB per context at the ResNet set).
static size_t size_class(size_t n) {
  if (n <= 8) return n;
  size_t step = 2;
  while (8 * step <= n) step *= 2;  // n in (4*step, 8*step]
  return (n + step - 1) / step * step;
}

u64* Context::alloc_limbs(size_t n_limbs, bool zero) {
  relieve_pressure();
  n_limbs = size_class(std::max<size_t>(n_limbs, 1));
  const size_t bytes = n_limbs * N * sizeof(u64);
  u64* p = nullptr;
  auto it = al_.free_lists.find(n_limbs);
  if (it != al_.free_lists.end() && !it->second.empty()) {
    p = it->second.back();
    it->second.pop_back();
    cached_bytes -= bytes;
  } else {
    // new blocks come from the driver's stream-ordered pool (no device synchronisation)
    cudaError_t e = cudaMallocAsync(&p, bytes, stream);
    for (int attempt = 0; e != cudaSuccess && attempt < 40; attempt++) {
      // out of memory: give the idle blocks back -- ours now, the other contexts' at their next
      // allocator call -- and try again
      cudaGetLastError();
      ACE_CUDA(cudaStreamSynchronize(stream));
      trim_cache();
      g_pressure_epoch++;
      pressure_seen_ = g_pressure_epoch.load();
      if (attempt) std::this_thread::sleep_for(std::chrono::milliseconds(50));
      e = cudaMallocAsync(&p, bytes, stream);
    }
    ACE_CUDA(e);
    al_.block_limbs[p] = n_limbs;
  }
  live_bytes += bytes;
  if (live_bytes > peak_bytes) peak_bytes = live_bytes;
  if (zero) {
    prof::Scope ps("memset(alloc)", stream);
    ACE_CUDA(cudaMemsetAsync(p, 0, bytes, stream));
  }
  return p;
}
void Context::free_limbs(u64* p) {
  if (!p) return;
  relieve_pressure();
  auto it = al_.block_limbs.find(p);
  if (it == al_.block_limbs.end()) throw std::runtime_error("free_limbs: unknown block");
  const size_t bytes = it->second * N * sizeof(u64);
  live_bytes -= bytes;
  if (cached_bytes + bytes > kCacheCapBytes) {  // enough idle blocks already: back to the pool
    al_.block_limbs.erase(it);
    ACE_CUDA(cudaFreeAsync(p, stream));
    return;
  }
  al_.free_lists[it->second].push_back(p);
  cached_bytes += bytes;
}
void Context::dev_malloc(void** p, size_t bytes) {
  cudaError_t e = cudaMalloc(p, bytes);
  for (int attempt = 0; e != cudaSuccess && attempt < 40; attempt++) {
    cudaGetLastError();
    ACE_CUDA(cudaDeviceSynchronize());
    trim_cache();
    g_pressure_epoch++;
    pressure_seen_ = g_pressure_epoch.load();
    ACE_CUDA(cudaStreamSynchronize(stream));
    cudaMemPool_t pool;
    ACE_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    ACE_CUDA(cudaMemPoolTrimTo(pool, 0));
    if (attempt) std::this_thread::sleep_for(std::chrono::milliseconds(50));
    e = cudaMalloc(p, bytes);
  }
  ACE_CUDA(e);
}
void Context::trim_cache() {
  for (auto& kv : al_.free_lists) {
    for (u64* p : kv.second) {
      al_.block_limbs.erase(p);
      cudaFreeAsync(p, stream);
    }
    kv.second.clear();
  }
  cached_bytes = 0;
}
void Context::upload(u64* dst, const u64* src, size_t n_limbs) {
  ACE_CUDA(cudaMemcpyAsync(dst, src, n_limbs * N * sizeof(u64), cudaMemcpyHostToDevice, stream));
}
void Context::download(u64* dst, const u64* src, size_t n_limbs) {
  ACE_CUDA(cudaMemcpyAsync(dst, src, n_limbs * N * sizeof(u64), cudaMemcpyDeviceToHost, stream));
  ACE_CUDA(cudaStreamSynchronize(stream));
}
void Context::sync() { ACE_CUDA(cudaStreamSynchronize(stream)); }

static void copy_limbs(u64* dst, const u64* src, size_t n_limbs, u32 N, cudaStream_t s) {
  prof::Scope ps("memcpy_d2d", s);
  ACE_CUDA(cudaMemcpyAsync(dst, src, n_limbs * N * sizeof(u64), cudaMemcpyDeviceToDevice, s));
}

// ---------------------------------------------------------------------------- transforms
void Context::ntt(u64* data, u32 g0, u32 n, bool standalone) {
  if (standalone) tr(TR_LIMB_NTT, 0, n);
  for (u32 done = 0; done < n; done += kMaxBatch) {
    LimbBatch b;
    b.base = data; b.src = nullptr;
    b.n    = std::min<u32>(kMaxBatch, n - done);
    for (u32 i = 0; i < b.n; i++) { b.slot[i] = (u16)(done + i); b.g[i] = (u16)(g0 + done + i); }
    launch_ntt(T, b, stream);
    launches += (logN > 12) ? 2 : 1;
  }
}
void Context::intt(u64* data, u32 g0, u32 n, bool standalone) {
  if (standalone) tr(TR_LIMB_NTT, 0, n);
  for (u32 done = 0; done < n; done += kMaxBatch) {
    LimbBatch b;
    b.base = data; b.src = nullptr;
    b.n    = std::min<u32>(kMaxBatch, n - done);
    for (u32 i = 0; i < b.n; i++) { b.slot[i] = (u16)(done + i); b.g[i] = (u16)(g0 + done + i); }
    launch_intt(T, b, stream);
    launches += (logN > 12) ? 2 : 1;
  }
}

// out-of-place inverse transform: dst[i] = INTT(src[i]) for n consecutive limbs
void Context::intt_from(u64* dst, const u64* src, u32 g0, u32 n) {
  for (u32 done = 0; done < n; done += kMaxBatch) {
    LimbBatch b;
    b.base = dst; b.src = src;
    b.n    = std::min<u32>(kMaxBatch, n - done);
    for (u32 i = 0; i < b.n; i++) {
      b.slot[i] = b.src_slot[i] = (u16)(done + i);
      b.g[i]    = (u16)(g0 + done + i);
    }
    launch_intt(T, b, stream);
    launches += (logN > 12) ? 2 : 1;
  }
}

// ---------------------------------------------------------------------------- ModUp
// tables _l_hat_inv_modq[part][n_in-1] and _l_hat_modp[num_q-1][part] (crt.c:399-533)
const Context::ModUpTab& Context::modup_tab(u32 num_q, u32 part) {
  auto key = std::make_pair(num_q, part);
  std::lock_guard<std::mutex> lk(sh_->mu);
  auto it  = sh_->modup_tabs.find(key);
  if (it != sh_->modup_tabs.end()) return it->second;
  ModUpTab t;
  u32 beta = (u32)num_decomp(num_q);
  t.start  = (u32)(part_size * part);
  t.n_in   = (part == beta - 1) ? num_q - t.start : (u32)part_size;
  std::vector<u64> hi(t.n_in), his(t.n_in);
  for (u32 i = 0; i < t.n_in; i++) {
    u64 qi = mod[t.start + i], hat = 1;
    for (u32 k = 0; k < t.n_in; k++)
      if (k != i) hat = hm::mulmod(hat, mod[t.start + k] % qi, qi);
    hi[i]  = hm::invmod_prime(hat, qi);
    his[i] = hm::shoup(hi[i], qi);
    t.g_in.push_back((u16)(t.start + i));
  }
  for (u32 o = 0; o < num_q + K; o++) {
    if (o >= t.start && o < t.start + t.n_in) continue;
    t.g_out.push_back((u16)gidx(o, num_q));
    t.out_slot.push_back((u16)o);
  }
  t.n_out = (u32)t.g_out.size();
  std::vector<u64> hm_(std::max<size_t>(1, (size_t)t.n_out * t.n_in));
  for (u32 o = 0; o < t.n_out; o++) {
    u64 tm = mod[t.g_out[o]];
    for (u32 i = 0; i < t.n_in; i++) {
      u64 hat = 1;
      for (u32 k = 0; k < t.n_in; k++)
        if (k != i) hat = hm::mulmod(hat, mod[t.start + k] % tm, tm);
      hm_[(size_t)o * t.n_in + i] = pack_hat(hat);
    }
  }
  t.hatinv    = to_device_shared(hi);
  t.hatinv_sh = to_device_shared(his);
  t.hatmod    = to_device_shared(hm_);
  return sh_->modup_tabs.emplace(key, std::move(t)).first->second;
}

void Context::fill_conv_desc(ConvDesc& d, const ModUpTab& t, const u64* x, u64* out) {
  d.x = x; d.out = out;
  d.hatinv = t.hatinv; d.hatinv_sh = t.hatinv_sh; d.hatmod = t.hatmod;
  d.n_in = t.n_in; d.n_out = t.n_out;
  for (u32 i = 0; i < t.n_in; i++) d.g_in[i] = t.g_in[i];
  for (u32 o = 0; o < t.n_out; o++) { d.g_out[o] = t.g_out[o]; d.out_slot[o] = t.out_slot[o]; }
}

// Decomp_modup (poly_eval.c:28-34): out = num_q + K limbs
void Context::decomp_modup(u64* out, const u64* in, u32 num_q, u32 part) {
  const ModUpTab& t = modup_tab(num_q, part);
  modup_from(out, in + (size_t)t.start * N, num_q, part);
}

// Mod_up (poly_eval.c:19-26): `digit` holds only the digit's own limbs (output of Decomp)
void Context::modup_from(u64* out, const u64* digit, u32 num_q, u32 part) {
  const ModUpTab& t = modup_tab(num_q, part);
  tr(TR_MODUP_DIGIT, num_q);
  u64* coef = alloc_limbs(t.n_in, false);
  copy_limbs(out + (size_t)t.start * N, digit, t.n_in, N, stream);
  intt_from(coef, digit, t.start, t.n_in);
  ConvDesc d;
  fill_conv_desc(d, t, coef, out);
  launch_base_conv(T, &d, 1, stream);
  LimbBatch b;
  b.base = out; b.src = nullptr; b.n = t.n_out;
  for (u32 o = 0; o < t.n_out; o++) { b.slot[o] = t.out_slot[o]; b.g[o] = t.g_out[o]; }
  launch_ntt(T, b, stream);
  launches += 1 + ((logN > 12) ? 2 : 1);
  free_limbs(coef);
}

// Mod_down (poly_eval.c:36-41).  Unlike the reference the P part of `in` is left intact.
void Context::mod_down(u64* out, const u64* in, u32 num_q) {
  if (ntt16_usable(T) && fused_tails()) {  // fused tail (batch.cu; falls back by itself when out overlaps in)
    ModdownJob j{out, in, num_q};
    moddown_batch(&j, 1);
    return;
  }
  tr(TR_MODDOWN_POLY, num_q);
  u64* pc   = alloc_limbs(K, false);
  u64* conv = alloc_limbs(num_q, false);
  intt_from(pc, in + (size_t)num_q * N, (u32)L, (u32)K);
  ConvDesc d;
  d.x = pc; d.out = conv;
  d.hatinv = phat_inv_; d.hatinv_sh = phat_inv_sh_; d.hatmod = phat_mod_q_;
  d.n_in = (u32)K; d.n_out = num_q;
  for (u32 i = 0; i < K; i++) d.g_in[i] = (u16)(L + i);
  for (u32 o = 0; o < num_q; o++) { d.g_out[o] = (u16)o; d.out_slot[o] = (u16)o; }
  launch_base_conv(T, &d, 1, stream);
  ntt(conv, 0, num_q, false);
  launch_moddown_tail(T, out, in, conv, nullptr, pinv_mod_q_, pinv_mod_q_sh_, num_q, stream);
  launches += 2;
  free_limbs(pc);
  free_limbs(conv);
}

// Rescale (poly_eval.c:43-49): out gets num_q - 1 limbs
void Context::rescale(u64* out, const u64* in, u32 num_q) {
  if (num_q < 2) throw std::runtime_error("Rescale: level not enough");
  if (ntt16_usable(T) && fused_tails()) {  // fused prologue / epilogue (batch.cu; falls back when out == in)
    RescaleJob j{out, in, num_q};
    rescale_batch(&j, 1);
    return;
  }
  tr(TR_RESCALE_POLY, num_q);
  const u32 l = num_q - 1;
  u64* last = alloc_limbs(1, false);
  u64* tmp  = alloc_limbs(l, false);
  intt_from(last, in + (size_t)l * N, l, 1);
  launch_rescale_pre(T, tmp, last, l, negqlinv_ + (size_t)l * L, negqlinv_sh_ + (size_t)l * L,
                     stream);
  ntt(tmp, 0, l, false);
  launch_rescale_post(T, out, in, tmp, qlinv_ + (size_t)l * L, qlinv_sh_ + (size_t)l * L, l,
                      stream);
  launches += 2;
  free_limbs(last);
  free_limbs(tmp);
}

// ---------------------------------------------------------------------------- keys
u32 Context::auto_index(int32_t rot) const {  // number_theory.c:187-199 with modulus 2N
  const u64 M = 2 * (u64)N;
  if (rot == 0) return 1;
  if (rot == (int32_t)(M - 1)) return (u32)rot;
  u64 gen = 5;
  if (rot < 0) gen = hm::powmod(5, N / 2 - 1, M);  // 5 has order N/2 in Z_2N^*
  return (u32)hm::powmod(gen, (u64)(rot < 0 ? -(int64_t)rot : rot), M);
}

const int64_t* Context::auto_order(u32 k) {  // number_theory.c:201-214 (is_ntt = TRUE)
  std::lock_guard<std::mutex> lk(sh_->mu);
  auto it = sh_->auto_orders.find(k);
  if (it != sh_->auto_orders.end()) return it->second;
  std::vector<int64_t> ord(N);
  const u32 logm = logN + 1;
  for (u64 j = 0; j < N; j++) {
    u64 jt  = (j << 1) + 1;
    u64 idx = ((jt * k) - (((jt * k) >> logm) << logm)) >> 1;
    ord[hm::bit_reverse((u32)j, logN)] = (int64_t)hm::bit_reverse((u32)idx, logN);
  }
  int64_t* d = nullptr;
  dev_malloc(&d, N * sizeof(int64_t));
  h2d_sync(d, ord.data(), N * sizeof(int64_t));
  sh_->auto_orders[k] = d;
  return d;
}

// the inverse permutation of auto_order(k) is the table of k^-1 mod 2N
const int64_t* Context::auto_order_inv(u32 k) {
  const u64 M = 2 * (u64)N;
  u64 inv = k;  // Newton iteration for the inverse of an odd number modulo a power of two
  for (int i = 0; i < 6; i++) inv = (inv * (2 + M - (k * inv) % M)) % M;
  if ((inv * k) % M != 1) throw std::runtime_error("auto_order_inv: no inverse");
  return auto_order((u32)inv);
}

void Context::import_key_limbs(SwitchKey& key, u32 part, int which, const u64* host) {
  if (part >= dnum) throw std::runtime_error("key part out of range");
  size_t per = G * (size_t)N;
  u64**  dst = which ? &key.k1 : &key.k0;
  if (*dst == nullptr) dev_malloc(dst, dnum * per * sizeof(u64));
  h2d_sync(*dst + part * per, host, per * sizeof(u64));
}

// ---------------------------------------------------------------------------- key switch
// Switch_key_precompute (polynomial.c:1224-1239, 1337-1343): ext[j] = ModUp of digit j, j < beta.
// Only the complement limbs of ext[j] are written; the digit's own limbs are the limbs of `d`
// themselves and ksw_acc reads them from there.
void Context::modup_all(u64* ext, const u64* d, u32 num_q) {
  const u32 beta = (u32)num_decomp(num_q), W = num_q + (u32)K;
  tr(TR_MODUP_DIGIT, num_q, beta);
  u64* coef = alloc_limbs(num_q, false);
  intt_from(coef, d, 0, num_q);
  ConvDesc  descs[6];
  LimbBatch nb;
  nb.base = ext; nb.src = nullptr; nb.n = 0;
  for (u32 j = 0; j < beta; j++) {
    const ModUpTab& t = modup_tab(num_q, j);
    u64* ext_j = ext + (size_t)j * W * N;
    fill_conv_desc(descs[j % 6], t, coef + (size_t)t.start * N, ext_j);
    if (j % 6 == 5 || j == beta - 1) {
      launch_base_conv(T, descs, j % 6 + 1, stream);
      launches++;
    }
    for (u32 o = 0; o < t.n_out; o++) {
      if (nb.n == kMaxBatch) {
        launch_ntt(T, nb, stream);
        launches += (logN > 12) ? 2 : 1;
        nb.n = 0;
      }
      nb.slot[nb.n] = (u16)(j * W + t.out_slot[o]);
      nb.g[nb.n]    = t.g_out[o];
      nb.n++;
    }
  }
  launch_ntt(T, nb, stream);
  launches += (logN > 12) ? 2 : 1;
  free_limbs(coef);
}

// Fast_switch_key_ext (ckks_evaluator.c:418-460): acc0/acc1 = sum_j ext_j (.) key_j over the
// num_q + K limbs of the extended basis (no ModDown)
void Context::ksw_acc(u64* acc0, u64* acc1, const u64* ext, const u64* d, u32 num_q,
                      const SwitchKey& key) {
  if (!key.k0 || !key.k1) throw std::runtime_error("switch key not loaded");
  {  // the emitted loop: per digit and limb 2 Hw_modmul + 2 Hw_modadd (GEN20:7005-7032)
    const uint64_t n = 2ull * num_decomp(num_q) * (num_q + K);
    tr(TR_LIMB_MUL, 0, n);
    tr(TR_LIMB_ADD, 0, n);
  }
  launch_ksw_inner(T, acc0, acc1, ext, d, (u32)part_size, key.k0, key.k1,
                   (u32)num_decomp(num_q), num_q, (u32)L, (u32)K, stream);
  launches++;
}

// Reduce_rns_base of two extended polynomials at once (polynomial.c:928-967); a0/a1 are laid
// out [num_q | K]; add0 (optional) is added to out0.  a1 == nullptr: one polynomial only.
void Context::mod_down_pair(u64* out0, u64* out1, const u64* a0, const u64* a1, u32 num_q,
                            const u64* add0, const u64* add1) {
  const u32 np = a1 ? 2 : 1;
  tr(TR_MODDOWN_POLY, num_q, np);
  if (add0) tr(TR_LIMB_ADD, 0, num_q);
  if (add1) tr(TR_LIMB_ADD, 0, num_q);
  u64* pc   = alloc_limbs(np * K, false);
  u64* conv = alloc_limbs(np * (size_t)num_q, false);
  ConvDesc md[2];
  for (u32 h = 0; h < np; h++) {
    intt_from(pc + (size_t)h * K * N, (h ? a1 : a0) + (size_t)num_q * N, (u32)L, (u32)K);
    ConvDesc& c = md[h];
    c.x = pc + (size_t)h * K * N; c.out = conv + (size_t)h * num_q * N;
    c.hatinv = phat_inv_; c.hatinv_sh = phat_inv_sh_; c.hatmod = phat_mod_q_;
    c.n_in = (u32)K; c.n_out = num_q;
    for (u32 i = 0; i < K; i++) c.g_in[i] = (u16)(L + i);
    for (u32 o = 0; o < num_q; o++) { c.g_out[o] = (u16)o; c.out_slot[o] = (u16)o; }
  }
  launch_base_conv(T, md, np, stream);
  // out must not overlap the Q limbs of the inputs for the fused form (it reads them last)
  auto overlaps = [&](const u64* o, const u64* a) { return o && a && o + (size_t)num_q * N > a && o < a + (size_t)num_q * N; };
  if (ntt16_usable(T) && fused_tails() && !overlaps(out0, a0) && !overlaps(out0, a1) && !overlaps(out1, a0) &&
      !overlaps(out1, a1) && np * num_q <= (u32)kMaxFused) {
    // NTT of the converted limbs with the tail (old - conv) * P^-1 (+ add) in its last store
    NttFusedBatch fb;
    fb.n = 0; fb.pre = 0; fb.post = 2;
    fb.pre_w = fb.pre_w_sh = nullptr; fb.pre_stride = 0;
    fb.post_w = pinv_mod_q_; fb.post_w_sh = pinv_mod_q_sh_; fb.post_stride = 0;
    for (u32 h = 0; h < np; h++)
      for (u32 o = 0; o < num_q; o++) {
        const u32 k = fb.n++;
        fb.dst[k] = (h ? out1 : out0) + (size_t)o * N;
        fb.src[k] = conv + ((size_t)h * num_q + o) * N;
        fb.aux[k] = (h ? a1 : a0) + (size_t)o * N;
        const u64* ad = h ? add1 : add0;
        fb.add[k] = ad ? ad + (size_t)o * N : nullptr;
        fb.g[k] = (u16)o; fb.g_from[k] = 0;
      }
    launch_ntt16_fused(T, fb, stream);
    launches += 1 + 2;
    free_limbs(pc); free_limbs(conv);
    return;
  }
  LimbBatch cb;
  cb.base = conv; cb.src = nullptr; cb.n = np * num_q;
  for (u32 i = 0; i < np * num_q; i++) { cb.slot[i] = (u16)i; cb.g[i] = (u16)(i % num_q); }
  launch_ntt(T, cb, stream);
  if (a1)
    launch_moddown_tail2(T, out0, out1, a0, a1, conv, conv + (size_t)num_q * N, add0, add1, pinv_mod_q_,
                         pinv_mod_q_sh_, num_q, stream);
  else
    launch_moddown_tail(T, out0, a0, conv, add0, pinv_mod_q_, pinv_mod_q_sh_, num_q, stream);
  launches += 1 + 1 + ((logN > 12) ? 2 : 1);
  free_limbs(pc); free_limbs(conv);
}

// Same dataflow as the emitted Rotate()/Relinearize() bodies, but batched: one INTT launch
// for all digits, one base-conversion launch (one descriptor per digit), one NTT launch for
// all complement limbs, one inner-product launch that reads the key once, and one ModDown
// for both output polynomials.
void Context::key_switch(u64* out0, u64* out1, const u64* d, u32 num_q, const SwitchKey& key,
                         const u64* add0, const u64* add1) {
  if (!key.k0 || !key.k1) throw std::runtime_error("switch key not loaded");
  const u32 beta = (u32)num_decomp(num_q), W = num_q + (u32)K;
  u64* ext = alloc_limbs((size_t)beta * W, false);
  u64* acc = alloc_limbs(2 * (size_t)W, false);
  modup_all(ext, d, num_q);
  ksw_acc(acc, acc + (size_t)W * N, ext, d, num_q, key);
  mod_down_pair(out0, out1, acc, acc + (size_t)W * N, num_q, add0, add1);
  free_limbs(ext); free_limbs(acc);
}

// emitted Rotate(): key switch c1, add c0, apply the automorphism to both polynomials
void Context::ct_rotate(u64* r0, u64* r1, const u64* c0, const u64* c1, u32 num_q,
                        int32_t rot_idx) {
  u32 k = auto_index(rot_idx);
  if (!has_rot_key(k)) throw std::runtime_error("rotation key not loaded");
  const int64_t* order = auto_order(k);
  u64* s = alloc_limbs(2 * (size_t)num_q, false);
  u64 *s0 = s, *s1 = s + (size_t)num_q * N;
  key_switch(s0, s1, c1, num_q, rot_key(k), c0);
  launch_gather(T, r0, s0, order, 0, num_q, stream);
  launch_gather(T, r1, s1, order, 0, num_q, stream);
  tr(TR_LIMB_ROT, 0, 2 * num_q);
  launches += 2;
  free_limbs(s);
}

// n rotations of ONE ciphertext: Decomp_modup of c1 once, then per rotation the key inner
// product, ModDown (+ c0) and the automorphism -- what the reference's bootstrap does by hand
// (Switch_key_precompute reused by every Fast_rotate, ckks_bootstrap_context.c:1284-1299) and what
// the 9 input rotations of an emitted convolution could share (ut_ksw_opt.cxx:663-770).  Every
// output is bit-identical to ct_rotate() of the same index: the ModUp does not depend on the key.
void Context::ct_rotate_hoisted(u64* const* r0, u64* const* r1, const u64* c0, const u64* c1, u32 num_q,
                                const int32_t* rots, size_t n) {
  for (size_t i = 0; i < n; i++)
    if (!has_rot_key(auto_index(rots[i]))) throw std::runtime_error("rotation key not loaded");
  const u32 beta = (u32)num_decomp(num_q), W = num_q + (u32)K;
  u64* ext = alloc_limbs((size_t)beta * W, false);
  modup_all(ext, c1, num_q);
  // the ModDowns of a group of rotations run as ONE batch (one inverse transform over all their
  // P limbs, one base conversion, one forward transform, one tail): 2.5 -> 2.2 ms for nine
  // rotations at l = 34 against one ModDown pair per rotation (nine separate ct_rotate: 4.2 ms)
  constexpr size_t kGroup = 12;
  const size_t g_max = n < kGroup ? n : kGroup;
  u64* acc = alloc_limbs(g_max * 2 * (size_t)W, false);
  u64* s   = alloc_limbs(g_max * 2 * (size_t)num_q, false);
  std::vector<ModdownJob> jobs;
  for (size_t at = 0; at < n; at += kGroup) {
    const size_t m = n - at < kGroup ? n - at : kGroup;
    jobs.clear();
    for (size_t i = 0; i < m; i++) {
      u64* a0 = acc + i * 2 * (size_t)W * N;
      u64* s0 = s + i * 2 * (size_t)num_q * N;
      ksw_acc(a0, a0 + (size_t)W * N, ext, c1, num_q, rot_key(auto_index(rots[at + i])));
      jobs.push_back(ModdownJob{s0, a0, num_q});
      jobs.push_back(ModdownJob{s0 + (size_t)num_q * N, a0 + (size_t)W * N, num_q});
    }
    moddown_batch(jobs.data(), jobs.size());
    for (size_t i = 0; i < m; i++) {
      u64* s0 = s + i * 2 * (size_t)num_q * N;
      const int64_t* order = auto_order(auto_index(rots[at + i]));
      launch_ew(T, EW_ADD, s0, s0, c0, 0, num_q, stream);
      launch_gather(T, r0[at + i], s0, order, 0, num_q, stream);
      launch_gather(T, r1[at + i], s0 + (size_t)num_q * N, order, 0, num_q, stream);
      tr(TR_LIMB_ADD, 0, num_q);
      tr(TR_LIMB_ROT, 0, 2 * num_q);
      launches += 3;
    }
  }
  free_limbs(ext); free_limbs(acc); free_limbs(s);
}

// acc (+)= ct (.) pt, both polynomials, all limbs in one launch (Mul_plain + Add_ciph of the
// emitted convolution loops: 2 Hw_modmul + 2 Hw_modadd per limb)
void Context::ct_mul_plain_acc(u64* acc0, u64* acc1, const u64* c0, const u64* c1, const u64* pt, u32 num_q,
                               bool first) {
  launch_ct_mul_plain_acc(T, acc0, acc1, first ? nullptr : acc0, first ? nullptr : acc1, c0, c1, pt, num_q, stream);
  tr(TR_LIMB_MUL, 0, 2 * num_q);
  if (!first) tr(TR_LIMB_ADD, 0, 2 * num_q);
  launches++;
}

// tensor product + emitted Relinearize(): r0 = a0 b0 + ks0(a1 b1), r1 = a0 b1 + a1 b0 + ks1
void Context::ct_mul_relin(u64* r0, u64* r1, const u64* a0, const u64* a1, const u64* b0,
                           const u64* b1, u32 num_q) {
  u64* t  = alloc_limbs(3 * (size_t)num_q, false);
  u64 *d0 = t, *d1 = t + (size_t)num_q * N, *d2 = t + 2 * (size_t)num_q * N;
  launch_tensor(T, d0, d1, d2, a0, a1, b0, b1, num_q, stream);
  // the two final additions ride on the ModDown tails of the key switch
  key_switch(r0, r1, d2, num_q, relin_key, d0, d1);
  tr(TR_LIMB_MUL, 0, 4 * num_q);
  tr(TR_LIMB_ADD, 0, 1 * num_q);  // d1 = a0 b1 + a1 b0; the other two are counted by mod_down_pair
  launches += 1;
  free_limbs(t);
}

}  // namespace ace
