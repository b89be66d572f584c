// kernels.cu -- hand-written sm_100a kernels for the CKKS evaluation hot path.
//
// Reference CPU routines each kernel replaces (paths under fhe-cmplr/rtlib/ant/):
//   ntt_*            Forward_transform / Inverse_transform      src/util/ntt.c:190-353
//   ew_kernel        Hw_modadd / Hw_modmul                      src/poly/poly_arith.c:14-39
//   gather_kernel    Hw_rotate                                  src/poly/poly_arith.c:41-56
//   base_conv_kernel Fast_base_conv / Decompose_modup MAC loop  src/util/polynomial.c:755-807,1297-1320
//   ksw_inner_kernel emitted key inner-product loops            dataset/resnet20_cifar10_pre.onnx.inc:7005-7032
//   moddown_tail     Reduce_rns_base tail                       src/util/polynomial.c:953-965
//   rescale_*        Rescale_poly (NTT branch)                  src/util/polynomial.c:1123-1161
//
// All of these are integer kernels: NTT and base conversion are bound by the INT32 multiply
// pipe (IMAD), the rest by HBM bandwidth.  No tensor-core path is used.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "kernels.cuh"
#include "prof.h"

namespace ace {

// One image alone is bound by launch latencies: chaining the launches brings it from 0.934 to
// 0.885 s (ResNet-20).  With several images in flight per GPU (worker contexts, context.cu) the
// other streams fill those gaps already and CTAs parked at griddepcontrol.wait only take SM
// slots from them (1.383 against 1.409 images/s with three streams, profiles/r2_pdl_chain.txt):
// the attribute is set while the primary context is the only one running images.
extern std::atomic<int> g_worker_contexts;
bool pdl_chain_enabled() {
  static const int mode = getenv("ACE_B200_NO_PDL") ? 0 : (getenv("ACE_B200_PDL_ALWAYS") ? 2 : 1);
  return mode == 2 || (mode == 1 && g_worker_contexts.load(std::memory_order_relaxed) == 0);
}

// ------------------------------------------------------------------------------------
// NTT.  N = 2^logN.  A transform is split into a "tile" phase working on contiguous tiles
// of T = min(N, 4096) elements in shared memory (the min(logN,12) stages whose butterfly
// stride fits in a tile) and, for N > 4096, a "strided" phase doing the remaining
// logN-12 stages entirely in registers: a thread owns one column (R = N/4096 elements at
// stride 4096), adjacent threads own adjacent columns, so every access is coalesced.
// Forward (Cooley-Tukey, natural -> bit-reversed): strided phase first, then tiles.
// Inverse (Gentleman-Sande, bit-reversed -> natural): tiles first, then strided phase,
// which also folds in N^-1.
// ------------------------------------------------------------------------------------
constexpr int kTileLog   = 12;
constexpr int kTile      = 1 << kTileLog;
constexpr int kNttThreads = 512;

// limb i of a batch: where it is read from and where the result goes (see kernels.cuh)
__device__ __forceinline__ u64* batch_dst(const LimbBatch& b, u32 i, u32 N) {
  return b.base + (size_t)b.slot[i] * N;
}
__device__ __forceinline__ const u64* batch_src(const LimbBatch& b, u32 i, u32 N) {
  return b.src ? b.src + (size_t)b.src_slot[i] * N : b.base + (size_t)b.slot[i] * N;
}
__device__ __forceinline__ bool batch_has_src(const LimbBatch& b) { return b.src != nullptr; }
__device__ __forceinline__ u64* batch_dst(const LimbPtrBatch& b, u32 i, u32) { return b.dst[i]; }
__device__ __forceinline__ const u64* batch_src(const LimbPtrBatch& b, u32 i, u32) {
  return b.src[i];
}
__device__ __forceinline__ bool batch_has_src(const LimbPtrBatch&) { return true; }

// Harvey-style lazy butterflies: values stay in [0, 4q) (forward) / [0, 2q) (inverse) between
// stages, one conditional subtraction per butterfly; the final pass normalises to [0, q), so
// the stored result is the same canonical residue the reference produces (ntt.c:206-263).
__device__ __forceinline__ u64 csub(u64 a, u64 m) { return a >= m ? a - m : a; }

// forward: (u, v) -> (u + w v, u - w v); in/out in [0, 4q).
// LAZY (moduli of at most 55 bits): the conditional subtraction is skipped altogether -- u grows
// by at most 6q per stage (see mulhi_short), 16 stages keep every value below 98 q < 2^62 for the
// <= 55-bit moduli this is used for, the Shoup product takes any 64-bit input -- and the final
// normalisation is one Barrett step (normalize_any).
// floor(x * w / 2^64) short by at most 2: the product of the two low halves and the low halves of
// the cross products are left out (3 wide multiplies instead of 4)
__device__ __forceinline__ u64 mulhi_short(u64 x, u64 w) {
  const u32 x0 = (u32)x, x1 = (u32)(x >> 32), w0 = (u32)w, w1 = (u32)(w >> 32);
  const u64 m1 = (u64)x1 * w0, m2 = (u64)x0 * w1;
  return (u64)x1 * w1 + (m1 >> 32) + (m2 >> 32);
}
template <bool LAZY = false>
__device__ __forceinline__ void ct_butterfly(u64& u, u64& v, u64 w, u64 wsh, u64 q) {
  const u64 q2 = 2 * q;
  u64 a = LAZY ? u : csub(u, q2);
  // LAZY: the quotient estimate may be short by 3 in all, t < 5q; the bounds have room for it
  u64 t = LAZY ? v * w - mulhi_short(v, wsh) * q : mul_shoup_lazy(v, w, wsh, q);
  u     = a + t;
  v     = a - t + (LAZY ? 3 * q2 : q2);  // t < 5q in the lazy form
}
// inverse: (u, v) -> (u + v, (u - v) w); in/out in [0, 2q)
__device__ __forceinline__ void gs_butterfly(u64& u, u64& v, u64 w, u64 wsh, u64 q) {
  const u64 q2 = 2 * q;
  u64 d = u - v + q2;
  u     = csub(u + v, q2);
  v     = mul_shoup_lazy(d, w, wsh, q);
}
__device__ __forceinline__ u64 normalize4(u64 a, u64 q) { return csub(csub(a, 2 * q), q); }
// any a < 2^64: a mod q with mu = floor(2^64 / q) (quotient estimate short by at most 1)
__device__ __forceinline__ u64 normalize_any(u64 a, u64 q, u64 mu) {
  return csub(a - __umul64hi(a, mu) * q, q);
}
__device__ __forceinline__ bool lazy_modulus(const Modulus& m) { return m.shift <= 53; }  // <= 55 bits

// ---- strided phase, forward: stages 0 .. SA-1, R = 2^SA rows at stride N/R -------------
template <int SA, class B>
__global__ void __launch_bounds__(128) ntt_fwd_strided(DeviceTables T, const __grid_constant__ B b) {
  pdl_enter();
  constexpr int R = 1 << SA;
  const u32 limb  = blockIdx.y;
  const u32 g     = b.g[limb];
  const u64 q     = T.mod[g].q;
  u64*      data  = batch_dst(b, limb, T.N);
  const u64* in   = batch_src(b, limb, T.N);
  const u64* tw   = T.tw + (size_t)g * T.N;
  const u64* twsh = T.tw_sh + (size_t)g * T.N;
  const u32 stride = T.N >> SA;  // = 4096
  const u32 col    = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= stride) return;
  u64 x[R];
#pragma unroll
  for (int r = 0; r < R; r++) x[r] = in[(size_t)r * stride + col];
  const bool lazy = lazy_modulus(T.mod[g]);  // uniform per block
#pragma unroll
  for (int s = 0; s < SA; s++) {
    const int m  = 1 << s;
    const int tr = R >> (s + 1);
#pragma unroll
    for (int p = 0; p < R / 2; p++) {
      const int i  = p / tr;
      const int lo = i * 2 * tr + (p % tr);
      if (lazy) ct_butterfly<true>(x[lo], x[lo + tr], tw[m + i], twsh[m + i], q);
      else ct_butterfly<false>(x[lo], x[lo + tr], tw[m + i], twsh[m + i], q);
    }
  }
#pragma unroll
  for (int r = 0; r < R; r++) data[(size_t)r * stride + col] = x[r];
}

// ---- strided phase, inverse: stages with t = N/R .. N/2, then * N^-1 --------------------
template <int SA, class B>
__global__ void __launch_bounds__(128) ntt_inv_strided(DeviceTables T, const __grid_constant__ B b) {
  pdl_enter();
  constexpr int R = 1 << SA;
  const u32 limb  = blockIdx.y;
  const u32 g     = b.g[limb];
  const u64 q     = T.mod[g].q;
  u64*      data  = batch_dst(b, limb, T.N);
  const u64* tw   = T.itw + (size_t)g * T.N;
  const u64* twsh = T.itw_sh + (size_t)g * T.N;
  const u64 ninv = T.n_inv[g], ninv_sh = T.n_inv_sh[g];
  const u32 stride = T.N >> SA;
  const u32 col    = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= stride) return;
  u64 x[R];
#pragma unroll
  for (int r = 0; r < R; r++) x[r] = data[(size_t)r * stride + col];
#pragma unroll
  for (int s = SA - 1; s >= 0; s--) {  // m = 2^s groups, row stride tr
    const int m  = 1 << s;
    const int tr = R >> (s + 1);
#pragma unroll
    for (int p = 0; p < R / 2; p++) {
      const int i  = p / tr;
      const int lo = i * 2 * tr + (p % tr);
      gs_butterfly(x[lo], x[lo + tr], tw[m + i], twsh[m + i], q);
    }
  }
#pragma unroll
  for (int r = 0; r < R; r++)
    data[(size_t)r * stride + col] = mul_shoup(x[r], ninv, ninv_sh, q);
}

// ---- tile phase, forward: stages s0 .. logN-1 on a contiguous tile in shared memory ----
template <class B>
__global__ void __launch_bounds__(kNttThreads) ntt_fwd_tile(DeviceTables T, const __grid_constant__ B b) {
  pdl_enter();
  extern __shared__ u64 sm[];
  const u32 limb   = blockIdx.y;
  const u32 g      = b.g[limb];
  const u64 q      = T.mod[g].q;
  const u32 tile   = T.N < (u32)kTile ? T.N : (u32)kTile;
  const u32 tlog   = T.logN < (u32)kTileLog ? T.logN : (u32)kTileLog;
  const u32 tbase  = blockIdx.x * tile;
  u64*      data   = batch_dst(b, limb, T.N) + tbase;
  const u64* tw    = T.tw + (size_t)g * T.N;
  const u64* twsh  = T.tw_sh + (size_t)g * T.N;
  const u64* in    = batch_src(b, limb, T.N) + tbase;
  for (u32 i = threadIdx.x; i < tile; i += blockDim.x) sm[i] = in[i];
  __syncthreads();
  const u32 s0 = T.logN - tlog;
  for (u32 s = s0; s < T.logN; s++) {
    const u32 lt = T.logN - 1 - s;  // log2 of butterfly stride t
    const u32 t  = 1u << lt;
    const u32 m  = 1u << s;
    for (u32 p = threadIdx.x; p < tile / 2; p += blockDim.x) {
      const u32 lo = ((p >> lt) << (lt + 1)) | (p & (t - 1));
      const u32 i  = (tbase + lo) >> (lt + 1);
      u64 u = sm[lo], v = sm[lo + t];
      ct_butterfly(u, v, tw[m + i], twsh[m + i], q);
      sm[lo]     = u;
      sm[lo + t] = v;
    }
    __syncthreads();
  }
  for (u32 i = threadIdx.x; i < tile; i += blockDim.x) data[i] = normalize4(sm[i], q);
}

// ---- tile phase, inverse: stages with t = 1 .. tile/2; folds N^-1 when it is the only phase
template <class B>
__global__ void __launch_bounds__(kNttThreads) ntt_inv_tile(DeviceTables T, const __grid_constant__ B b) {
  pdl_enter();
  extern __shared__ u64 sm[];
  const u32 limb   = blockIdx.y;
  const u32 g      = b.g[limb];
  const u64 q      = T.mod[g].q;
  const u32 tile   = T.N < (u32)kTile ? T.N : (u32)kTile;
  const u32 tlog   = T.logN < (u32)kTileLog ? T.logN : (u32)kTileLog;
  const u32 tbase  = blockIdx.x * tile;
  u64*      data   = batch_dst(b, limb, T.N) + tbase;
  const u64* tw    = T.itw + (size_t)g * T.N;
  const u64* twsh  = T.itw_sh + (size_t)g * T.N;
  const u64* in    = batch_src(b, limb, T.N) + tbase;
  for (u32 i = threadIdx.x; i < tile; i += blockDim.x) sm[i] = in[i];
  __syncthreads();
  for (u32 lt = 0; lt < tlog; lt++) {
    const u32 t = 1u << lt;
    const u32 m = T.N >> (lt + 1);
    for (u32 p = threadIdx.x; p < tile / 2; p += blockDim.x) {
      const u32 lo = ((p >> lt) << (lt + 1)) | (p & (t - 1));
      const u32 i  = (tbase + lo) >> (lt + 1);
      u64 u = sm[lo], v = sm[lo + t];
      gs_butterfly(u, v, tw[m + i], twsh[m + i], q);
      sm[lo]     = u;
      sm[lo + t] = v;
    }
    __syncthreads();
  }
  if (T.logN <= (u32)kTileLog) {
    const u64 ninv = T.n_inv[g], ninv_sh = T.n_inv_sh[g];
    for (u32 i = threadIdx.x; i < tile; i += blockDim.x)
      data[i] = mul_shoup(sm[i], ninv, ninv_sh, q);
  } else {
    for (u32 i = threadIdx.x; i < tile; i += blockDim.x) data[i] = sm[i];
  }
}

// ---- tile phase for N >= 4096, register radix-8 version ---------------------------------
// 512 threads x 8 elements = one 4096-element tile.  The 12 stages run as 4 passes of three
// radix-2 stages held in registers; between passes the tile is transposed through one padded
// shared-memory buffer: a thread writes, at the end of a pass, exactly the positions it read at
// the start of that pass (no other thread touches them in that pass), so the one barrier between
// the write of pass p and the read of pass p+1 is all that is needed -- half the shared memory of
// two alternating buffers, which the L1 gets for the twiddle tables.  Pass 0 of the
// forward transform reads global memory directly in its register pattern (coalesced) and the
// last pass leaves 8 consecutive coefficients per thread, written with 16-byte stores; the
// inverse transform mirrors this.
__host__ __device__ constexpr u32 kPad(u32 i) { return i + (i >> 3); }  // 1 pad per 8 elements
constexpr int kTileSm = kTile + (kTile >> 3);

// position of element k (0..7) of thread tid in a pass whose innermost stride is 2^ltq
__device__ __forceinline__ u32 r8_base(u32 tid, u32 ltq) {
  return ((tid >> ltq) << (ltq + 3)) | (tid & ((1u << ltq) - 1));
}

// three forward stages (strides 4tq, 2tq, tq) on x[0..7]; ia = (tile_base + base) >> (ltq+3)
template <bool LAZY>
__device__ __forceinline__ void ct_radix8(u64 (&x)[8], const u64* __restrict__ tw,
                                          const u64* __restrict__ twsh, u32 ma, u32 ia, u64 q) {
  {
    const u64 w = __ldg(tw + ma + ia), ws = __ldg(twsh + ma + ia);
#pragma unroll
    for (int k = 0; k < 4; k++) ct_butterfly<LAZY>(x[k], x[k + 4], w, ws, q);
  }
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const u32 idx = 2 * ma + 2 * ia + h;
    const u64 w = __ldg(tw + idx), ws = __ldg(twsh + idx);
    ct_butterfly<LAZY>(x[4 * h + 0], x[4 * h + 2], w, ws, q);
    ct_butterfly<LAZY>(x[4 * h + 1], x[4 * h + 3], w, ws, q);
  }
#pragma unroll
  for (int h = 0; h < 4; h++) {
    const u32 idx = 4 * ma + 4 * ia + h;
    const u64 w = __ldg(tw + idx), ws = __ldg(twsh + idx);
    ct_butterfly<LAZY>(x[2 * h], x[2 * h + 1], w, ws, q);
  }
}

// three inverse stages (strides tq, 2tq, 4tq); mc = N / (8 tq) is the group count of the
// widest stage, ic = (tile_base + base) >> (ltq+3)
__device__ __forceinline__ void gs_radix8(u64 (&x)[8], const u64* __restrict__ tw,
                                          const u64* __restrict__ twsh, u32 mc, u32 ic, u64 q) {
#pragma unroll
  for (int h = 0; h < 4; h++) {
    const u32 idx = 4 * mc + 4 * ic + h;
    const u64 w = __ldg(tw + idx), ws = __ldg(twsh + idx);
    gs_butterfly(x[2 * h], x[2 * h + 1], w, ws, q);
  }
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const u32 idx = 2 * mc + 2 * ic + h;
    const u64 w = __ldg(tw + idx), ws = __ldg(twsh + idx);
    gs_butterfly(x[4 * h + 0], x[4 * h + 2], w, ws, q);
    gs_butterfly(x[4 * h + 1], x[4 * h + 3], w, ws, q);
  }
  {
    const u64 w = __ldg(tw + mc + ic), ws = __ldg(twsh + mc + ic);
#pragma unroll
    for (int k = 0; k < 4; k++) gs_butterfly(x[k], x[k + 4], w, ws, q);
  }
}

template <bool LAZY>
__device__ __forceinline__ void fwd_tile8_body(const DeviceTables& T, u64* sm, const u64* in,
                                               u64* data, const u64* tw, const u64* twsh, u32 tbase,
                                               const Modulus& mod) {
  const u64 q   = mod.q;
  const u32 tid = threadIdx.x;
  const u32 m0  = T.N >> kTileLog;  // groups of the first tile stage (stride 2048)
  u64 x[8];
#pragma unroll
  for (int k = 0; k < 8; k++) x[k] = in[tid + 512 * k];
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const u32 ltq  = 9 - 3 * p;  // log2 of the innermost stride of this pass
    const u32 base = r8_base(tid, ltq);
    if (p > 0) {
#pragma unroll
      for (int k = 0; k < 8; k++) x[k] = sm[kPad(base + (k << ltq))];
    }
    ct_radix8<LAZY>(x, tw, twsh, m0 << (3 * p), (tbase + base) >> (ltq + 3), q);
    if (p < 3) {
      // (the positions written here are the ones this thread read at the start of the pass)
#pragma unroll
      for (int k = 0; k < 8; k++) sm[kPad(base + (k << ltq))] = x[k];
      __syncthreads();
    }
  }
  ulonglong2* out = reinterpret_cast<ulonglong2*>(data + 8 * tid);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (LAZY)
      out[k] = make_ulonglong2(normalize_any(x[2 * k], q, mod.mu_hi),
                               normalize_any(x[2 * k + 1], q, mod.mu_hi));
    else
      out[k] = make_ulonglong2(normalize4(x[2 * k], q), normalize4(x[2 * k + 1], q));
  }
}

template <class B>
__global__ void __launch_bounds__(512, 2) ntt_fwd_tile8(DeviceTables T, const __grid_constant__ B b) {
  pdl_enter();
  extern __shared__ u64 sm[];
  const u32 limb  = blockIdx.y;
  const u32 g     = b.g[limb];
  const Modulus mod = T.mod[g];
  const u32 tbase = blockIdx.x * kTile;
  u64*      data  = batch_dst(b, limb, T.N) + tbase;
  const u64* tw   = T.tw + (size_t)g * T.N;
  const u64* twsh = T.tw_sh + (size_t)g * T.N;
  // the strided phase (if any) already moved the limb to its destination
  const u64* in = T.logN == (u32)kTileLog ? batch_src(b, limb, T.N) + tbase : data;
  if (lazy_modulus(mod)) fwd_tile8_body<true>(T, sm, in, data, tw, twsh, tbase, mod);
  else fwd_tile8_body<false>(T, sm, in, data, tw, twsh, tbase, mod);
}

template <class B>
__global__ void __launch_bounds__(512, 2) ntt_inv_tile8(DeviceTables T, const __grid_constant__ B b) {
  pdl_enter();
  extern __shared__ u64 sm[];
  u64* buf[2] = {sm, sm};
  const u32 limb  = blockIdx.y;
  const u32 g     = b.g[limb];
  const u64 q     = T.mod[g].q;
  const u32 tbase = blockIdx.x * kTile;
  u64*      data  = batch_dst(b, limb, T.N) + tbase;
  const u64* tw   = T.itw + (size_t)g * T.N;
  const u64* twsh = T.itw_sh + (size_t)g * T.N;
  const u32 tid   = threadIdx.x;
  u64 x[8];
  {
    const u64* src = batch_src(b, limb, T.N) + tbase;
    const ulonglong2* in = reinterpret_cast<const ulonglong2*>(src + 8 * tid);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      ulonglong2 v = in[k];
      x[2 * k] = v.x; x[2 * k + 1] = v.y;
    }
  }
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const u32 ltq  = 3 * p;
    const u32 base = r8_base(tid, ltq);
    if (p > 0) {
#pragma unroll
      for (int k = 0; k < 8; k++) x[k] = buf[(p - 1) & 1][kPad(base + (k << ltq))];
    }
    gs_radix8(x, tw, twsh, T.N >> (ltq + 3), (tbase + base) >> (ltq + 3), q);
    if (p < 3) {
#pragma unroll
      for (int k = 0; k < 8; k++) buf[p & 1][kPad(base + (k << ltq))] = x[k];
      __syncthreads();
    }
  }
  if (T.logN == (u32)kTileLog) {  // single-phase transform: fold N^-1 here
    const u64 ninv = T.n_inv[g], ninv_sh = T.n_inv_sh[g];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = mul_shoup(x[k], ninv, ninv_sh, q);
  }
#pragma unroll
  for (int k = 0; k < 8; k++) data[tid + 512 * k] = x[k];
}

template <int SA, class B>
static void launch_strided(bool fwd, const DeviceTables& T, const B& b, cudaStream_t s) {
  dim3 grid((T.N >> SA) / 128, b.n);
  if (fwd) {
    launch_chain(ntt_fwd_strided<SA, B>, grid, 128, 0, s, T, b);
  } else {
    launch_chain(ntt_inv_strided<SA, B>, grid, 128, 0, s, T, b);
  }
}

template <class B>
static void launch_strided_any(bool fwd, const DeviceTables& T, const B& b, cudaStream_t s) {
  switch (T.logN - kTileLog) {
    case 1: launch_strided<1, B>(fwd, T, b, s); break;
    case 2: launch_strided<2, B>(fwd, T, b, s); break;
    case 3: launch_strided<3, B>(fwd, T, b, s); break;
    case 4: launch_strided<4, B>(fwd, T, b, s); break;
    case 5: launch_strided<5, B>(fwd, T, b, s); break;
    default: break;
  }
}

// ACE_B200_NTTHIST=1: histogram of batch sizes (limbs per launch), printed at exit
static struct NttHist {
  bool on = getenv("ACE_B200_NTTHIST") != nullptr;
  size_t fwd[kMaxBatch + 1] = {0}, inv[kMaxBatch + 1] = {0};
  ~NttHist() {
    if (!on) return;
    for (int d = 0; d < 2; d++) {
      const size_t* h = d ? inv : fwd;
      size_t launches = 0, limbs = 0;
      for (int i = 0; i <= kMaxBatch; i++) { launches += h[i]; limbs += h[i] * i; }
      printf("[ace_b200 ntthist] %s: %zu launches, %zu limbs; limbs by batch size:", d ? "intt" : "ntt", launches, limbs);
      const int edges[] = {1, 2, 3, 5, 9, 17, 25, 37, 49, 73, 97, 145, 193};
      for (int e = 0; e + 1 < 13; e++) {
        size_t l = 0, n = 0;
        for (int i = edges[e]; i < edges[e + 1] && i <= kMaxBatch; i++) { l += h[i] * i; n += h[i]; }
        printf(" [%d-%d]: %zu limbs/%zu", edges[e], edges[e + 1] - 1, l, n);
      }
      printf("\n");
    }
  }
} g_ntt_hist;

template <class B>
static void launch_ntt_impl(const DeviceTables& T, const B& b, cudaStream_t s) {
  if (b.n == 0) return;
  if (g_ntt_hist.on) g_ntt_hist.fwd[b.n <= (u32)kMaxBatch ? b.n : kMaxBatch]++;
  prof::Scope prof_scope_("ntt", s);
  if (ntt16_usable(T)) { launch_ntt16(T, b, s); return; }
  const u32 tile = T.N < (u32)kTile ? T.N : (u32)kTile;
  if (T.logN > (u32)kTileLog) launch_strided_any<B>(true, T, b, s);
  dim3 grid(T.N / tile, b.n);
  if (T.logN >= (u32)kTileLog) {
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(ntt_fwd_tile8<B>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           kTileSm * (int)sizeof(u64));
      attr = true;
    }
    launch_chain(ntt_fwd_tile8<B>, grid, 512, kTileSm * sizeof(u64), s, T, b);
    return;
  }
  u32  threads = tile / 2 < (u32)kNttThreads ? tile / 2 : (u32)kNttThreads;
  launch_chain(ntt_fwd_tile<B>, grid, threads, tile * sizeof(u64), s, T, b);
}

template <class B>
static void launch_intt_impl(const DeviceTables& T, const B& b, cudaStream_t s) {
  if (b.n == 0) return;
  if (g_ntt_hist.on) g_ntt_hist.inv[b.n <= (u32)kMaxBatch ? b.n : kMaxBatch]++;
  prof::Scope prof_scope_("intt", s);
  if (ntt16_usable(T)) { launch_intt16(T, b, s); return; }
  const u32 tile = T.N < (u32)kTile ? T.N : (u32)kTile;
  dim3 grid(T.N / tile, b.n);
  if (T.logN >= (u32)kTileLog) {
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(ntt_inv_tile8<B>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           kTileSm * (int)sizeof(u64));
      attr = true;
    }
    launch_chain(ntt_inv_tile8<B>, grid, 512, kTileSm * sizeof(u64), s, T, b);
  } else {
    u32 threads = tile / 2 < (u32)kNttThreads ? tile / 2 : (u32)kNttThreads;
    launch_chain(ntt_inv_tile<B>, grid, threads, tile * sizeof(u64), s, T, b);
  }
  if (T.logN > (u32)kTileLog) launch_strided_any<B>(false, T, b, s);
}

void launch_ntt(const DeviceTables& T, const LimbBatch& b, cudaStream_t s) { launch_ntt_impl(T, b, s); }
void launch_intt(const DeviceTables& T, const LimbBatch& b, cudaStream_t s) { launch_intt_impl(T, b, s); }
void launch_ntt(const DeviceTables& T, const LimbPtrBatch& b, cudaStream_t s) { launch_ntt_impl(T, b, s); }
void launch_intt(const DeviceTables& T, const LimbPtrBatch& b, cudaStream_t s) { launch_intt_impl(T, b, s); }

// ------------------------------------------------------------------------------------
// Element-wise kernels (HBM-bound): one thread per coefficient, grid.y = limb.
// ------------------------------------------------------------------------------------
template <int OP>
__global__ void __launch_bounds__(256) ew_kernel(DeviceTables T, u64* __restrict__ r,
                                                 const u64* __restrict__ a,
                                                 const u64* __restrict__ b, u32 g0) {
  pdl_enter();
  const Modulus m   = T.mod[g0 + blockIdx.y];
  const size_t  off = (size_t)blockIdx.y * T.N;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < T.N; i += gridDim.x * blockDim.x) {
    u64 x = a[off + i], y = b[off + i];
    u64 z;
    if (OP == EW_ADD) z = add_mod(x, y, m.q);
    if (OP == EW_SUB) z = sub_mod(x, y, m.q);
    if (OP == EW_MUL) z = mul_mod(x, y, m);
    r[off + i] = z;
  }
}

static inline dim3 ew_grid(const DeviceTables& T, u32 n_limbs) {
  u32 bx = (T.N + 255) / 256;
  return dim3(bx, n_limbs);
}

void launch_ew(const DeviceTables& T, EwOp op, u64* r, const u64* a, const u64* b, u32 g0,
               u32 n_limbs, cudaStream_t s) {
  prof::Scope prof_scope_("ew", s);
  if (n_limbs == 0) return;
  dim3 grid = ew_grid(T, n_limbs);
  switch (op) {
    case EW_ADD: launch_chain(ew_kernel<EW_ADD>, grid, 256, 0, s, T, r, a, b, g0); break;
    case EW_SUB: launch_chain(ew_kernel<EW_SUB>, grid, 256, 0, s, T, r, a, b, g0); break;
    case EW_MUL: launch_chain(ew_kernel<EW_MUL>, grid, 256, 0, s, T, r, a, b, g0); break;
  }
}

__global__ void __launch_bounds__(256) gather_kernel(DeviceTables T, u64* __restrict__ r,
                                                     const u64* __restrict__ a,
                                                     const int64_t* __restrict__ order,
                                                     u32 g0) {
  pdl_enter();
  const u64    q   = T.mod[g0 + blockIdx.y].q;
  const size_t off = (size_t)blockIdx.y * T.N;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < T.N; i += gridDim.x * blockDim.x) {
    int64_t k = order[i];
    // negative entries only occur for coefficient-form tables (number_theory.c:215-224)
    r[off + i] = k >= 0 ? a[off + k] : q - a[off - k];
  }
}

void launch_gather(const DeviceTables& T, u64* r, const u64* a, const int64_t* order, u32 g0,
                   u32 n_limbs, cudaStream_t s) {
  prof::Scope prof_scope_("gather", s);
  if (n_limbs == 0) return;
  launch_chain(gather_kernel, ew_grid(T, n_limbs), 256, 0, s, T, r, a, order, g0);
}

__global__ void __launch_bounds__(256) mul_scalar_kernel(DeviceTables T, u64* __restrict__ r,
                                                         const u64* __restrict__ a,
                                                         const u64* __restrict__ sc,
                                                         const u64* __restrict__ sc_sh,
                                                         u32 g0) {
  pdl_enter();
  const u64    q   = T.mod[g0 + blockIdx.y].q;
  const u64    w = sc[blockIdx.y], wsh = sc_sh[blockIdx.y];
  const size_t off = (size_t)blockIdx.y * T.N;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < T.N; i += gridDim.x * blockDim.x)
    r[off + i] = mul_shoup(a[off + i], w, wsh, q);
}

void launch_mul_scalar(const DeviceTables& T, u64* r, const u64* a, const u64* sc,
                       const u64* sc_sh, u32 g0, u32 n_limbs, cudaStream_t s) {
  prof::Scope prof_scope_("mul_scalar", s);
  if (n_limbs == 0) return;
  launch_chain(mul_scalar_kernel, ew_grid(T, n_limbs), 256, 0, s, T, r, a, sc, sc_sh, g0);
}

// ------------------------------------------------------------------------------------
// Fast base conversion.  One thread per coefficient; the n_in scaled inputs y_i stay in
// registers and are re-used for every output limb, so HBM traffic is exactly
// (n_in + n_out) limbs; the work is n_in*n_out 64x64->128 MACs per coefficient (INT-bound).
// ------------------------------------------------------------------------------------
struct ConvDescPack {
  ConvDesc d[kMaxConvPack];
};

// Sum of products without carries.  When every modulus is below 2^60 (DeviceTables::small_moduli)
// the conversion matrices are stored split at bit 30 -- word = (h >> 30) << 32 | (h & (2^30-1)),
// Context::pack_hat -- and the scaled inputs y are split the same way in registers.  The four
// partial products a_i * b_j are < 2^60, so up to 16 of them add up in a 64-bit register: one MAC is
// four IMAD.WIDE.U32 with a 64-bit addend and nothing else (the generic mac128 takes ~20
// instructions).  fold() rebuilds the exact 128-bit sum, so the reduced result is unchanged.
struct DotAcc30 {
  u64 s00 = 0, s01 = 0, s10 = 0, s11 = 0;
  __device__ __forceinline__ void mac(u32 a0, u32 a1, u32 b0, u32 b1) {
    s00 += (u64)a0 * b0;
    s01 += (u64)a0 * b1;
    s10 += (u64)a1 * b0;
    s11 += (u64)a1 * b1;
  }
  // value = s00 + (s01 + s10) * 2^30 + s11 * 2^60
  __device__ __forceinline__ void fold(u64& lo, u64& hi) const {
    const u64 mid   = s01 + s10;                       // < 2^65: keep the carry
    const u64 mid_c = mid < s01 ? 1ull : 0ull;
    u64 l = s00, h = 0;
    const u64 m_lo = mid << 30, m_hi = (mid >> 34) | (mid_c << 30);
    l += m_lo; h += m_hi + (l < m_lo ? 1ull : 0ull);
    const u64 t_lo = s11 << 60, t_hi = s11 >> 4;
    l += t_lo; h += t_hi + (l < t_lo ? 1ull : 0ull);
    lo = l; hi = h;
  }
};

template <int NIN>
__device__ __forceinline__ void base_conv_fast(const DeviceTables& T, const ConvDesc& D,
                                               const u64* sh_hat, u32 n, u32 o_begin, u32 o_end) {
  u32 y0[NIN], y1[NIN];
#pragma unroll
  for (int i = 0; i < NIN; i++) {
    const u64 q = T.mod[D.g_in[i]].q;
    const u64 y = mul_shoup(D.x[(size_t)i * T.N + n], D.hatinv[i], D.hatinv_sh[i], q);
    y0[i] = (u32)y & 0x3FFFFFFFu;
    y1[i] = (u32)(y >> 30);
  }
  for (u32 o = o_begin; o < o_end; o++) {
    const Modulus m = T.mod[D.g_out[o]];
    const uint2*  h = reinterpret_cast<const uint2*>(sh_hat + (o - o_begin) * NIN);
    DotAcc30 acc;
#pragma unroll
    for (int i = 0; i < NIN; i++) {
      const uint2 w = h[i];
      acc.mac(y0[i], y1[i], w.x, w.y);
    }
    u64 lo, hi;
    acc.fold(lo, hi);
    D.out[(size_t)D.out_slot[o] * T.N + n] = reduce128(lo, hi, m);
  }
}

// One thread per coefficient; blockIdx.y = descriptor.  The n_in inputs are scaled ONCE (n_in
// Shoup products) and stay in registers while the thread walks over every output limb; the whole
// conversion matrix of the descriptor sits in shared memory.  (Round 1 split the outputs over
// blockIdx.z in groups of 8 and rescaled the inputs in every group: at l = 34 a digit's 12 inputs
// were scaled 5 times.)  n_in <= 16 with small moduli takes the exact-size carry-free path above;
// anything else the generic 128-bit accumulation.
template <int MAXIN>
__global__ void __launch_bounds__(128) base_conv_kernel(DeviceTables T,
                                                        const __grid_constant__ ConvDescPack P) {
  pdl_enter();
  extern __shared__ u64 sh_hat[];  // [n_out][n_in]
  const ConvDesc& D = P.d[blockIdx.y];
  const u32 n_in = D.n_in, n_out = D.n_out;
  for (u32 i = threadIdx.x; i < n_in * n_out; i += blockDim.x) sh_hat[i] = D.hatmod[i];
  __syncthreads();
  const u32 n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= T.N) return;
  if (MAXIN <= 16 && T.small_moduli) {
    switch (n_in) {
#define ACE_BC_CASE(k) case k: if (k <= MAXIN) base_conv_fast<(k <= MAXIN ? k : 1)>(T, D, sh_hat, n, 0, n_out); return;
      ACE_BC_CASE(1) ACE_BC_CASE(2) ACE_BC_CASE(3) ACE_BC_CASE(4) ACE_BC_CASE(5) ACE_BC_CASE(6)
      ACE_BC_CASE(7) ACE_BC_CASE(8) ACE_BC_CASE(9) ACE_BC_CASE(10) ACE_BC_CASE(11) ACE_BC_CASE(12)
      ACE_BC_CASE(13) ACE_BC_CASE(14) ACE_BC_CASE(15) ACE_BC_CASE(16)
#undef ACE_BC_CASE
      default: return;
    }
  }
  u64 y[MAXIN];
#pragma unroll
  for (int i = 0; i < MAXIN; i++) {
    if (i < (int)n_in) {
      const u64 q = T.mod[D.g_in[i]].q;
      y[i] = mul_shoup(D.x[(size_t)i * T.N + n], D.hatinv[i], D.hatinv_sh[i], q);
    }
  }
  for (u32 o = 0; o < n_out; o++) {
    const Modulus m = T.mod[D.g_out[o]];
    const u64*    h = sh_hat + o * n_in;
    u64 lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < MAXIN; i++) {
      if (i < (int)n_in) {
        u64 w = h[i];
        if (T.small_moduli) w = ((w >> 32) << 30) | (w & 0x3FFFFFFFull);  // stored split
        mac128(lo, hi, y[i], w);
      }
    }
    D.out[(size_t)D.out_slot[o] * T.N + n] = reduce128(lo, hi, m);
  }
}

void launch_base_conv(const DeviceTables& T, const ConvDesc* descs, u32 n_desc,
                      cudaStream_t s) {
  prof::Scope prof_scope_("base_conv", s);
  if (n_desc == 0) return;
  ConvDescPack P;
  u32 max_in = 0, max_out = 0;
  for (u32 i = 0; i < n_desc; i++) {
    P.d[i] = descs[i];
    if (descs[i].n_in > max_in) max_in = descs[i].n_in;
    if (descs[i].n_out > max_out) max_out = descs[i].n_out;
  }
  // ADVICE (round 1): more than 48 inputs were silently ignored, and the generic 128-bit
  // accumulation overflows beyond 16 terms of 2^62-sized factors: refuse instead
  if (max_in > 48 || max_out > 64 || (!T.small_moduli && max_in > 16))
    throw std::runtime_error("base conversion: unsupported digit size (n_in " + std::to_string(max_in) + ")");
  dim3   grid((T.N + 127) / 128, n_desc);
  size_t shm = (size_t)max_in * max_out * sizeof(u64);
  if (max_in <= 4) {
    launch_chain(base_conv_kernel<4>, grid, 128, shm, s, T, P);
  } else if (max_in <= 12) {
    launch_chain(base_conv_kernel<12>, grid, 128, shm, s, T, P);
  } else if (max_in <= 16) {
    launch_chain(base_conv_kernel<16>, grid, 128, shm, s, T, P);
  } else {
    launch_chain(base_conv_kernel<48>, grid, 128, shm, s, T, P);
  }
}

// ------------------------------------------------------------------------------------
// Key-switch inner product.  Streams the evaluation key exactly once (2*beta*(num_q+K)
// limbs) plus beta*(num_q+K) ext limbs; accumulates in 128 bits and reduces once.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ksw_inner_kernel(DeviceTables T, u64* __restrict__ acc0,
                                                        u64* __restrict__ acc1,
                                                        const u64* __restrict__ ext,
                                                        const u64* __restrict__ own,
                                                        u32 part_size,
                                                        const u64* __restrict__ key0,
                                                        const u64* __restrict__ key1, u32 beta,
                                                        u32 num_q, u32 L, u32 K) {
  pdl_enter();
  const u32     o = blockIdx.y;
  const u32     g = o < num_q ? o : L + (o - num_q);
  const u32     W = num_q + K;
  const Modulus m = T.mod[g];
  const u32     n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= T.N) return;
  u64 lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0;
  for (u32 j = 0; j < beta; j++) {
    const bool mine = own != nullptr && o < num_q && o / part_size == j;
    const u64 e = mine ? own[(size_t)o * T.N + n] : ext[((size_t)j * W + o) * T.N + n];
    const u64 k0 = key0[((size_t)j * (L + K) + g) * T.N + n];
    const u64 k1 = key1[((size_t)j * (L + K) + g) * T.N + n];
    mac128(lo0, hi0, e, k0);
    mac128(lo1, hi1, e, k1);
  }
  acc0[(size_t)o * T.N + n] = reduce128(lo0, hi0, m);
  acc1[(size_t)o * T.N + n] = reduce128(lo1, hi1, m);
}

void launch_ksw_inner(const DeviceTables& T, u64* acc0, u64* acc1, const u64* ext,
                      const u64* own, u32 part_size, const u64* key0, const u64* key1,
                      u32 beta, u32 num_q, u32 L, u32 K, cudaStream_t s) {
  prof::Scope prof_scope_("ksw_inner", s);
  dim3 grid((T.N + 255) / 256, num_q + K);
  launch_chain(ksw_inner_kernel, grid, 256, 0, s, T, acc0, acc1, ext, own, part_size, key0, key1, beta,
                                        num_q, L, K);
}

__global__ void __launch_bounds__(256) moddown_tail_kernel(
    DeviceTables T, u64* __restrict__ out, const u64* __restrict__ old,
    const u64* __restrict__ conv, const u64* __restrict__ add, const u64* __restrict__ pinv,
    const u64* __restrict__ pinv_sh) {
  pdl_enter();
  const u32    l   = blockIdx.y;
  const u64    q   = T.mod[l].q;
  const u64    w = pinv[l], wsh = pinv_sh[l];
  const size_t off = (size_t)l * T.N;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < T.N; i += gridDim.x * blockDim.x) {
    u64 v = mul_shoup(sub_mod(old[off + i], conv[off + i], q), w, wsh, q);
    if (add != nullptr) v = add_mod(v, add[off + i], q);
    out[off + i] = v;
  }
}

void launch_moddown_tail(const DeviceTables& T, u64* out, const u64* old, const u64* conv,
                         const u64* add, const u64* pinv, const u64* pinv_sh, u32 n_limbs,
                         cudaStream_t s) {
  prof::Scope prof_scope_("moddown_tail", s);
  if (n_limbs == 0) return;
  launch_chain(moddown_tail_kernel, ew_grid(T, n_limbs), 256, 0, s, T, out, old, conv, add, pinv,
                                                          pinv_sh);
}

// the tails of both polynomials of a key switch in one launch (blockIdx.z = polynomial)
__global__ void __launch_bounds__(256) moddown_tail2_kernel(DeviceTables T, u64* out0, u64* out1, const u64* old0,
                                                            const u64* old1, const u64* conv0, const u64* conv1,
                                                            const u64* add0, const u64* add1,
                                                            const u64* __restrict__ pinv,
                                                            const u64* __restrict__ pinv_sh) {
  pdl_enter();
  u64* __restrict__       out  = blockIdx.z ? out1 : out0;
  const u64* __restrict__ old  = blockIdx.z ? old1 : old0;
  const u64* __restrict__ conv = blockIdx.z ? conv1 : conv0;
  const u64* __restrict__ add  = blockIdx.z ? add1 : add0;
  const u32    l   = blockIdx.y;
  const u64    q   = T.mod[l].q;
  const u64    w = pinv[l], wsh = pinv_sh[l];
  const size_t off = (size_t)l * T.N;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < T.N; i += gridDim.x * blockDim.x) {
    u64 v = mul_shoup(sub_mod(old[off + i], conv[off + i], q), w, wsh, q);
    if (add != nullptr) v = add_mod(v, add[off + i], q);
    out[off + i] = v;
  }
}
void launch_moddown_tail2(const DeviceTables& T, u64* out0, u64* out1, const u64* old0, const u64* old1,
                          const u64* conv0, const u64* conv1, const u64* add0, const u64* add1, const u64* pinv,
                          const u64* pinv_sh, u32 n_limbs, cudaStream_t s) {
  prof::Scope prof_scope_("moddown_tail", s);
  if (n_limbs == 0) return;
  dim3 grid = ew_grid(T, n_limbs);
  grid.z = 2;
  launch_chain(moddown_tail2_kernel, grid, 256, 0, s, T, out0, out1, old0, old1, conv0, conv1, add0, add1, pinv,
               pinv_sh);
}

__global__ void __launch_bounds__(256) rescale_pre_kernel(DeviceTables T, u64* __restrict__ tmp,
                                                          const u64* __restrict__ last, u32 l,
                                                          const u64* __restrict__ negqlinv,
                                                          const u64* __restrict__ negqlinv_sh) {
  pdl_enter();
  const u32    i   = blockIdx.y;
  const u64    qi  = T.mod[i].q, ql = T.mod[l].q;
  const u64    w = negqlinv[i], wsh = negqlinv_sh[i];
  const size_t off = (size_t)i * T.N;
  for (u32 n = blockIdx.x * blockDim.x + threadIdx.x; n < T.N; n += gridDim.x * blockDim.x)
    tmp[off + n] = mul_shoup(switch_modulus(last[n], ql, qi), w, wsh, qi);
}

void launch_rescale_pre(const DeviceTables& T, u64* tmp, const u64* last, u32 l,
                        const u64* negqlinv, const u64* negqlinv_sh, cudaStream_t s) {
  prof::Scope prof_scope_("rescale_pre", s);
  if (l == 0) return;
  launch_chain(rescale_pre_kernel, ew_grid(T, l), 256, 0, s, T, tmp, last, l, negqlinv, negqlinv_sh);
}

__global__ void __launch_bounds__(256) rescale_post_kernel(DeviceTables T, u64* __restrict__ out,
                                                           const u64* __restrict__ c,
                                                           const u64* __restrict__ tmp,
                                                           const u64* __restrict__ qlinv,
                                                           const u64* __restrict__ qlinv_sh) {
  pdl_enter();
  const u32    i   = blockIdx.y;
  const u64    qi  = T.mod[i].q;
  const u64    w = qlinv[i], wsh = qlinv_sh[i];
  const size_t off = (size_t)i * T.N;
  for (u32 n = blockIdx.x * blockDim.x + threadIdx.x; n < T.N; n += gridDim.x * blockDim.x)
    out[off + n] = add_mod(mul_shoup(c[off + n], w, wsh, qi), tmp[off + n], qi);
}

void launch_rescale_post(const DeviceTables& T, u64* out, const u64* c, const u64* tmp,
                         const u64* qlinv, const u64* qlinv_sh, u32 n_limbs, cudaStream_t s) {
  prof::Scope prof_scope_("rescale_post", s);
  if (n_limbs == 0) return;
  launch_chain(rescale_post_kernel, ew_grid(T, n_limbs), 256, 0, s, T, out, c, tmp, qlinv, qlinv_sh);
}

// ------------------------------------------------------------------------------------
// Batched tails for polynomials that live in different allocations (scheduler path): one
// descriptor per output limb, blockIdx.y = descriptor.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) moddown_tail_batch_kernel(
    DeviceTables T, const __grid_constant__ Ptr3Batch P, const u64* __restrict__ pinv,
    const u64* __restrict__ pinv_sh) {
  pdl_enter();
  const u32  y = blockIdx.y, l = P.g[y];
  const u64  q = T.mod[l].q, w = pinv[l], wsh = pinv_sh[l];
  u64*       out = P.r[y];
  const u64 *old = P.a[y], *conv = P.b[y];
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < T.N; i += gridDim.x * blockDim.x)
    out[i] = mul_shoup(sub_mod(old[i], conv[i], q), w, wsh, q);
}
void launch_moddown_tail_batch(const DeviceTables& T, const Ptr3Batch& P, const u64* pinv,
                               const u64* pinv_sh, cudaStream_t s) {
  if (P.n == 0) return;
  prof::Scope prof_scope_("moddown_tail", s);
  launch_chain(moddown_tail_batch_kernel, ew_grid(T, P.n), 256, 0, s, T, P, pinv, pinv_sh);
}

__global__ void __launch_bounds__(256) rescale_pre_batch_kernel(
    DeviceTables T, const __grid_constant__ Ptr3Batch P, const u64* __restrict__ negqlinv,
    const u64* __restrict__ negqlinv_sh, u32 L) {
  pdl_enter();
  const u32  y = blockIdx.y, i = P.g[y], l = P.aux[y];
  const u64  qi = T.mod[i].q, ql = T.mod[l].q;
  const u64  w = negqlinv[(size_t)l * L + i], wsh = negqlinv_sh[(size_t)l * L + i];
  u64*       tmp  = P.r[y];
  const u64* last = P.a[y];
  for (u32 n = blockIdx.x * blockDim.x + threadIdx.x; n < T.N; n += gridDim.x * blockDim.x)
    tmp[n] = mul_shoup(switch_modulus(last[n], ql, qi), w, wsh, qi);
}
void launch_rescale_pre_batch(const DeviceTables& T, const Ptr3Batch& P, const u64* negqlinv,
                              const u64* negqlinv_sh, u32 L, cudaStream_t s) {
  if (P.n == 0) return;
  prof::Scope prof_scope_("rescale_pre", s);
  launch_chain(rescale_pre_batch_kernel, ew_grid(T, P.n), 256, 0, s, T, P, negqlinv, negqlinv_sh, L);
}

__global__ void __launch_bounds__(256) rescale_post_batch_kernel(
    DeviceTables T, const __grid_constant__ Ptr3Batch P, const u64* __restrict__ qlinv,
    const u64* __restrict__ qlinv_sh, u32 L) {
  pdl_enter();
  const u32  y = blockIdx.y, i = P.g[y], l = P.aux[y];
  const u64  qi = T.mod[i].q;
  const u64  w = qlinv[(size_t)l * L + i], wsh = qlinv_sh[(size_t)l * L + i];
  u64*       out = P.r[y];
  const u64 *c = P.a[y], *tmp = P.b[y];
  for (u32 n = blockIdx.x * blockDim.x + threadIdx.x; n < T.N; n += gridDim.x * blockDim.x)
    out[n] = add_mod(mul_shoup(c[n], w, wsh, qi), tmp[n], qi);
}
void launch_rescale_post_batch(const DeviceTables& T, const Ptr3Batch& P, const u64* qlinv,
                               const u64* qlinv_sh, u32 L, cudaStream_t s) {
  if (P.n == 0) return;
  prof::Scope prof_scope_("rescale_post", s);
  launch_chain(rescale_post_batch_kernel, ew_grid(T, P.n), 256, 0, s, T, P, qlinv, qlinv_sh, L);
}

// Tensor product of two ciphertexts in one pass (Mul_ciphertext3, ckks_evaluator.c:133-165):
//   d0 = a0 b0,  d1 = a0 b1 + a1 b0,  d2 = a1 b1     -- 4 limbs read, 3 written per limb index
__global__ void __launch_bounds__(256) tensor_kernel(DeviceTables T, u64* __restrict__ d0,
                                                     u64* __restrict__ d1, u64* __restrict__ d2,
                                                     const u64* __restrict__ a0,
                                                     const u64* __restrict__ a1,
                                                     const u64* __restrict__ b0,
                                                     const u64* __restrict__ b1) {
  pdl_enter();
  const Modulus m   = T.mod[blockIdx.y];
  const size_t  off = (size_t)blockIdx.y * T.N;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < T.N; i += gridDim.x * blockDim.x) {
    const u64 x0 = a0[off + i], x1 = a1[off + i], y0 = b0[off + i], y1 = b1[off + i];
    d0[off + i] = mul_mod(x0, y0, m);
    d1[off + i] = add_mod(mul_mod(x0, y1, m), mul_mod(x1, y0, m), m.q);
    d2[off + i] = mul_mod(x1, y1, m);
  }
}
void launch_tensor(const DeviceTables& T, u64* d0, u64* d1, u64* d2, const u64* a0,
                   const u64* a1, const u64* b0, const u64* b1, u32 n_limbs, cudaStream_t s) {
  if (n_limbs == 0) return;
  prof::Scope prof_scope_("tensor", s);
  launch_chain(tensor_kernel, ew_grid(T, n_limbs), 256, 0, s, T, d0, d1, d2, a0, a1, b0, b1);
}

}  // namespace ace
