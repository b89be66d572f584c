// ntt16.cu -- negacyclic NTT / inverse NTT for N = 2^16 (the degree of every ACE-emitted ResNet):
// the dominant kernel family of the runtime.
//
// Reference semantics: Forward_transform / Inverse_transform, fhe-cmplr/rtlib/ant/src/util/
// ntt.c:190-353 (natural order in -> bit-reversed out, and back); the stored residues are the
// canonical ones, so limbs are bit-identical to the reference's.
//
// Layout of the work.  The limb is a 256 x 256 matrix (element n = 256 r + c).  A transform is
// two kernels of 8 butterfly stages each, every kernel two passes of radix-16 held in registers
// (16 coefficients per thread, 4 stages, 32 butterflies) with ONE exchange through shared memory
// in between:
//   forward  K1 "cols": stages 0-7 run along r.  A CTA owns 16 adjacent columns (a 128-byte
//            segment of every row), 256 threads = 16 columns x 16 row groups.  The twiddles of
//            these stages do not depend on the column: 255 per limb in all.
//            K2 "rows": stages 8-15 run inside a row.  A CTA owns 16 rows (4096 contiguous
//            coefficients); the exchange between the two passes stays inside a half warp.
//   inverse  K1 "rows" then K2 "cols", mirrored.
// Butterflies are multiply-first in BOTH directions.  The forward transform is the usual
// psi-merged Cooley-Tukey.  The inverse is NOT the Gentleman-Sande network of the reference: it is
// a decimation-in-time cyclic inverse (bit-reversed in, natural out, twiddles omega^-j, the first
// 12 of every 64 products are by omega^0 = 1 and skipped) followed by one multiplication with
// psi^-n N^-1 per coefficient, which doubles as the final reduction -- the same function, hence
// the same canonical residues, but the butterfly is (u + w v, u - w v) with no correction on the
// sum, so both directions share one lazy butterfly.
//
// Three arithmetic back ends, chosen per limb by the size of its modulus:
//   FP64   q < 2^50.4 (the scaling primes q_1.. of the ResNet sets).  Residues are exact integers
//          in doubles; w v mod q = (h - c q) + l with h + l = w v exactly (one DMUL + one DFMA),
//          c = rint(h / q) (DFMA against 1.5 * 2^52), h - c q exact in one more DFMA.  8 FP64
//          instructions per butterfly plus 3 per coefficient every second stage to fold values
//          back to |x| <= q/2 -- measured 1.55 T butterflies/s against 0.98 T/s for the integer
//          form: B200 issues DFMA at (almost) the IMAD rate, and a 51-bit product needs 3 FP64
//          multiplies where the integer pipe needs 9 32-bit ones, 5 of them wide (half rate).
//   INT    q < 2^57: 64-bit lazy Shoup butterfly, no conditional subtraction anywhere.
//          t = w v - est q with a 3-multiply quotient estimate that may be short by 2 (t < 4q);
//          values grow by at most 4q per stage, 16 stages stay below 65 q < 2^64; one Barrett step
//          at the end.  5 IMAD.WIDE.U32 + 4 IMAD per butterfly (w v and -est q accumulate into one
//          64-bit value: t = v w + est (2^64 - q)).
//   INT/CS q < 2^61 (the 60-bit special primes): exact Shoup quotient (t < 2q, +2q per stage) and
//          ONE conditional subtraction per coefficient per radix-16 pass instead of one per
//          butterfly: a pass starts below 8q and ends below 16q <= 2^64 (q < 2^60; for
//          2^60 <= q < 2^61 the bounds are 4q / 8q with a second subtraction mid-pass).
//          Arithmetic ceiling 0.95 T butterflies/s against 0.79 T/s for round 2's first form
//          (conditional subtraction in every butterfly).
// Twiddles: {w, w'} interleaved (one 16-byte load per butterfly) for the integer forms, one double
// for the FP64 form.
#include <cooperative_groups.h>

#include <cstdlib>

#include "kernels.cuh"
#include "prof.h"

namespace ace {

namespace {

__device__ __forceinline__ u64* b_dst(const LimbBatch& b, u32 i, u32 N) {
  return b.base + (size_t)b.slot[i] * N;
}
__device__ __forceinline__ const u64* b_src(const LimbBatch& b, u32 i, u32 N) {
  return b.src ? b.src + (size_t)b.src_slot[i] * N : b.base + (size_t)b.slot[i] * N;
}
__device__ __forceinline__ u64* b_dst(const LimbPtrBatch& b, u32 i, u32) { return b.dst[i]; }
__device__ __forceinline__ const u64* b_src(const LimbPtrBatch& b, u32 i, u32) { return b.src[i]; }

__device__ __forceinline__ u64* b_dst(const NttFusedBatch& b, u32 i, u32) { return b.dst[i]; }
__device__ __forceinline__ const u64* b_src(const NttFusedBatch& b, u32 i, u32) { return b.src[i]; }

__device__ __forceinline__ u64 csub64(u64 a, u64 m) { return a >= m ? a - m : a; }

// What is folded into the first load / the last store of a forward transform (NttFusedBatch).
// NoHook: nothing, the calls vanish at compile time.
struct NoHook {
  static constexpr bool kActive = false;
  static constexpr u32  post_mode = 0;
  const u64*            aux = nullptr;
  __device__ __forceinline__ u64 pre(u64 v) const { return v; }
  __device__ __forceinline__ u64 post(u64 y, u64, u32) const { return y; }
};
struct FusedHook {
  static constexpr bool kActive = true;
  u32        pre_mode, post_mode;
  u64        q, q_from, pw, pwsh, ew, ewsh;
  const u64* aux;
  const u64* add;
  __device__ __forceinline__ FusedHook(const DeviceTables& T, const NttFusedBatch& b, u32 limb) {
    const u32 g = b.g[limb], gf = b.g_from[limb];
    pre_mode = b.pre; post_mode = b.post;
    q = T.mod[g].q;
    q_from = pre_mode ? T.mod[gf].q : 0;
    pw = pwsh = ew = ewsh = 0;
    if (pre_mode) { pw = b.pre_w[(size_t)gf * b.pre_stride + g]; pwsh = b.pre_w_sh[(size_t)gf * b.pre_stride + g]; }
    if (post_mode) { ew = b.post_w[(size_t)gf * b.post_stride + g]; ewsh = b.post_w_sh[(size_t)gf * b.post_stride + g]; }
    aux = b.aux[limb];
    add = b.add[limb];
  }
  __device__ __forceinline__ u64 pre(u64 v) const {
    return pre_mode ? mul_shoup(switch_modulus(v, q_from, q), pw, pwsh, q) : v;
  }
  // a = the epilogue operand aux[idx], fetched by the caller (staged through shared memory)
  __device__ __forceinline__ u64 post(u64 y, u64 a, u32 idx) const {
    if (post_mode == 1) return add_mod(mul_shoup(a, ew, ewsh, q), y, q);
    if (post_mode == 2) {
      u64 v = mul_shoup(sub_mod(a, y, q), ew, ewsh, q);
      if (add != nullptr) v = add_mod(v, add[idx], q);
      return v;
    }
    return y;
  }
};
// The epilogue operand of a tile (16 rows x 256 coefficients of `aux`) is copied into shared
// memory with cp.async at the START of the second kernel -- coalesced 16-byte chunks, no
// registers, in flight during both passes -- in groups of 16 coefficients 144 bytes apart, so
// that each thread later reads its 16 consecutive coefficients without bank conflicts.  (Loading
// it at the end with per-thread strided loads made the fused transform SLOWER than transform +
// separate tail kernel: 124.8 vs 111.6 us for a Mod_down at l = 34.)
constexpr u32 kAuxGroup = 18, kAuxRow = 16 * kAuxGroup;  // words
constexpr size_t kAuxSmem = 16 * kAuxRow * sizeof(u64);  // 36 864 bytes
__device__ __forceinline__ void stage_aux(u64* saux, const u64* aux_tile) {
  const u32 t = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const u32 id = t + 256 * i, row = id >> 7, p = (id & 127) * 2;
    const u32 dst = (u32)__cvta_generic_to_shared(saux + row * kAuxRow + kAuxGroup * (p >> 4) + (p & 15));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(aux_tile + row * 256 + p) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void stage_aux_wait() {
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
}
template <class B> struct HookOf { typedef NoHook type; static __device__ __forceinline__ NoHook make(const DeviceTables&, const B&, u32) { return NoHook(); } };
template <> struct HookOf<NttFusedBatch> {
  typedef FusedHook type;
  static __device__ __forceinline__ FusedHook make(const DeviceTables& T, const NttFusedBatch& b, u32 limb) { return FusedHook(T, b, limb); }
};

// ------------------------------------------------------------------------------------------
// integer arithmetic
// ------------------------------------------------------------------------------------------
// t = v w - est q  (mod 2^64),  est = floor(v w' / 2^64) - {0,1,2}:   0 <= t < 4q for any v < 2^64
template <class MK>
__device__ __forceinline__ u64 mul_lazy4(u64 v, const ulonglong2 tw, const MK& M) {
  const u32 v0 = (u32)v, v1 = (u32)(v >> 32);
  const u32 w0 = (u32)tw.x, w1 = (u32)(tw.x >> 32);
  const u32 s0 = (u32)tw.y, s1 = (u32)(tw.y >> 32);
  const u64 est = (u64)v1 * s1 + __umulhi(v1, s0) + __umulhi(v0, s1);
  const u32 e0 = (u32)est, e1 = (u32)(est >> 32);
  u64 t = (u64)v0 * w0;
  t     = (u64)e0 * M.nq0 + t;
  const u32 hi = (u32)(t >> 32) + v0 * w1 + v1 * w0 + e0 * M.nq1 + e1 * M.nq0;
  return ((u64)hi << 32) | (u32)t;
}

// CSUB = false: q < 2^57, short quotient estimate, nothing is ever subtracted (65 q < 2^64).
// CSUB = true : q < 2^61, exact quotient (t < 2q, +2q per stage) and ONE conditional subtraction
//               per coefficient per pass -- when it enters the pass (from_mid) -- instead of one per
//               butterfly: q < 2^60: in < 8q, four stages, < 16q <= 2^64;  2^60 <= q < 2^61: in
//               < 4q and a second subtraction after the second stage (< 8q <= 2^64).
template <bool CSUB>
struct ArithInt {
  typedef u64        E;   // a coefficient in registers
  typedef ulonglong2 TW;  // a twiddle
  struct Mod {
    u64 q;
    u32 nq0, nq1;  // 2^64 - q
    u64 q4;        // 4q
    u64 mu;        // floor(2^64 / q)
    u64 top;       // CSUB: bound enforced at the start of a pass (8q, or 4q for q >= 2^60)
    u64 q2;        // 2q
    bool mid;      // CSUB: q >= 2^60, subtract again after the second stage
  };
  static __device__ __forceinline__ Mod make(const Modulus& m) {
    const u64 nq = 0 - m.q;
    const bool big = (m.q >> 60) != 0;
    return Mod{m.q, (u32)nq, (u32)(nq >> 32), 4 * m.q, m.mu_hi, big ? 4 * m.q : 8 * m.q, 2 * m.q, big};
  }
  static __device__ __forceinline__ const TW* fwd_table(const DeviceTables& T, u32 g) { return T.ftw2 + (size_t)g * 65536; }
  static __device__ __forceinline__ const TW* inv_table(const DeviceTables& T, u32 g) { return T.itw2 + (size_t)g * 65536; }
  static __device__ __forceinline__ const TW* scale_table(const DeviceTables& T, u32 g) { return T.ips2 + (size_t)g * 65536; }
  static __device__ __forceinline__ TW ld(const TW* p) { return __ldg(p); }
  static __device__ __forceinline__ E from_canonical(u64 x, const Mod&) { return x; }
  // a value that went through shared or global memory: the start of a pass
  static __device__ __forceinline__ E from_mid(u64 x, const Mod& M) { return CSUB ? csub64(x, M.top) : x; }
  static __device__ __forceinline__ u64 to_mid(E x) { return x; }
  // (u, v) -> (u + w v, u - w v)
  static __device__ __forceinline__ void bfly(E& u, E& v, const TW tw, const Mod& M) {
    if (CSUB) {
      const u64 t = v * tw.x - __umul64hi(v, tw.y) * M.q;  // exact Shoup: t < 2q for any v
      const u64 a = u;
      u = a + t;
      v = a - t + M.q2;
    } else {
      const u64 t = mul_lazy4(v, tw, M);  // t < 4q
      const u64 a = u;
      u = a + t;
      v = a - t + M.q4;
    }
  }
  static __device__ __forceinline__ void bfly_first(E& u, E& v, const Mod& M) {  // canonical in, w = 1
    const u64 a = u;
    u = a + v;
    v = a + M.q - v;
  }
  static __device__ __forceinline__ void bfly_one(E& u, E& v, const Mod& M) {  // v < 2q, w = 1
    const u64 a = u, t = v;
    u = a + t;
    v = a - t + (CSUB ? M.q2 : M.q4);
  }
  static __device__ __forceinline__ void fold_after(E (&x)[16], const Mod& M, int stage) {
    if (CSUB && stage == 1 && M.mid) {
#pragma unroll
      for (int k = 0; k < 16; k++) x[k] = csub64(x[k], M.q4);
    }
  }
  static __device__ __forceinline__ u64 to_canonical(E x, const Mod& M) {
    return csub64(x - __umul64hi(x, M.mu) * M.q, M.q);  // the estimate is short by at most 1
  }
  static __device__ __forceinline__ u64 scale_to_canonical(E x, const TW tw, const Mod& M) {
    const u64 y = mul_lazy4(x, tw, M);  // < 4q for any x
    return csub64(csub64(y, 2 * M.q), M.q);
  }
};

// ------------------------------------------------------------------------------------------
// FP64 arithmetic (q < fp64_max_q): values are exact integers, |x| < 2^53 throughout
// ------------------------------------------------------------------------------------------
struct ModD { double q, qinv; };
constexpr double kMagic = 6755399441055744.0;   // 1.5 * 2^52: x + kMagic - kMagic = rint(x), |x| < 2^51
constexpr double kTwo52 = 4503599627370496.0;

// w v - rint(w v / q) q:  |result| <= (0.5 + |v| 2^-52) q   (|v| < 2^51, 0 <= w < q)
__device__ __forceinline__ double mulmod_dp(double v, double w, const ModD& M) {
  const double h = __dmul_rn(v, w);
  const double l = __fma_rn(v, w, -h);                                   // h + l = v w exactly
  const double c = __dadd_rn(__fma_rn(h, M.qinv, kMagic), -kMagic);      // rint(h / q)
  const double r = __fma_rn(-c, M.q, h);                                 // exact: |h - c q| < 2^53
  return __dadd_rn(r, l);
}

struct ArithDP {
  typedef double E;
  typedef double TW;
  typedef ModD   Mod;
  static __device__ __forceinline__ Mod make(const Modulus& m) {
    const double q = (double)m.q;
    return ModD{q, __ddiv_rn(1.0, q)};
  }
  static __device__ __forceinline__ const TW* fwd_table(const DeviceTables& T, u32 g) { return T.ftwd + (size_t)g * 65536; }
  static __device__ __forceinline__ const TW* inv_table(const DeviceTables& T, u32 g) { return T.itwd + (size_t)g * 65536; }
  static __device__ __forceinline__ const TW* scale_table(const DeviceTables& T, u32 g) { return T.ipsd + (size_t)g * 65536; }
  static __device__ __forceinline__ TW ld(const TW* p) { return __ldg(p); }
  static __device__ __forceinline__ E from_canonical(u64 x, const Mod&) {  // x < 2^52
    return __dadd_rn(__longlong_as_double((long long)(x | 0x4330000000000000ull)), -kTwo52);
  }
  static __device__ __forceinline__ E from_mid(u64 x, const Mod&) { return __longlong_as_double((long long)x); }
  static __device__ __forceinline__ u64 to_mid(E x) { return (u64)__double_as_longlong(x); }
  // |u| <= 1.5 q and |v| <= 1.5 q < 2^51 in; |out| <= |u| + q
  static __device__ __forceinline__ void bfly(E& u, E& v, const TW w, const Mod& M) {
    const double t = mulmod_dp(v, w, M);
    const double a = u;
    u = __dadd_rn(a, t);
    v = __dadd_rn(a, -t);
  }
  static __device__ __forceinline__ void bfly_first(E& u, E& v, const Mod&) {
    const double a = u;
    u = __dadd_rn(a, v);
    v = __dadd_rn(a, -v);
  }
  static __device__ __forceinline__ void bfly_one(E& u, E& v, const Mod& M) { bfly_first(u, v, M); }
  static __device__ __forceinline__ double fold(double x, const Mod& M) {  // |x| < 2^53 -> |r| <= q/2
    const double c = __dadd_rn(__fma_rn(x, M.qinv, kMagic), -kMagic);
    return __fma_rn(-c, M.q, x);
  }
  static __device__ __forceinline__ void fold_after(E (&x)[16], const Mod& M, int stage) {
    if (stage == 0 || stage == 2) {
#pragma unroll
      for (int k = 0; k < 16; k++) x[k] = fold(x[k], M);
    }
  }
  static __device__ __forceinline__ u64 nonneg_to_u64(double y, const Mod& M) {  // |y| < q
    const double z = y < 0.0 ? __dadd_rn(y, M.q) : y;
    return (u64)__double_as_longlong(__dadd_rn(z, kTwo52)) & 0x000FFFFFFFFFFFFFull;
  }
  static __device__ __forceinline__ u64 to_canonical(E x, const Mod& M) { return nonneg_to_u64(fold(x, M), M); }
  static __device__ __forceinline__ u64 scale_to_canonical(E x, const TW w, const Mod& M) {
    return nonneg_to_u64(mulmod_dp(x, w, M), M);  // |x| <= 1.5 q: |product| < 0.9 q
  }
};

// Bounds of the FP64 form (units of q, q < 2^50.415 so that 1.5 q < 2^51): a radix-16 pass is
//   stage, FOLD, stage, stage, FOLD, stage      with |x| <= 1.5 at its input:
//   1.5 -> 2.5 -> 0.5 -> 1.5 -> 2.5 -> 0.5 -> 1.5; every multiplied value is <= 1.5 q < 2^51 and
// every sum is <= 2.5 q < 2^53 (exact).  The integer forms ignore the folds.

// four stages on 16 registers: stage i pairs (k, k + (8 >> i)); TW(i, h) = twiddle of group h
// (0 <= h < 2^i) of stage i
#define ACE_R16_FWD(x, TWF)                                           \
  _Pragma("unroll") for (int i_ = 0; i_ < 4; i_++) {                  \
    const int tr_ = 8 >> i_;                                          \
    _Pragma("unroll") for (int h_ = 0; h_ < (1 << i_); h_++) {        \
      const typename A::TW tw_ = TWF(i_, h_);                         \
      _Pragma("unroll") for (int e_ = 0; e_ < tr_; e_++)              \
        A::bfly(x[2 * tr_ * h_ + e_], x[2 * tr_ * h_ + e_ + tr_], tw_, M); \
    }                                                                 \
    A::fold_after(x, M, i_);                                          \
  }

// decimation in time: stage i pairs (k, k + (1 << i)); TW(i, e) = twiddle of offset e (0 <= e < 2^i)
#define ACE_R16_DIT(x, TWF)                                           \
  _Pragma("unroll") for (int i_ = 0; i_ < 4; i_++) {                  \
    const int m_ = 1 << i_;                                           \
    _Pragma("unroll") for (int e_ = 0; e_ < m_; e_++) {               \
      const typename A::TW tw_ = TWF(i_, e_);                         \
      _Pragma("unroll") for (int h_ = 0; h_ < 8 / m_; h_++)           \
        A::bfly(x[2 * m_ * h_ + e_], x[2 * m_ * h_ + e_ + m_], tw_, M); \
    }                                                                 \
    A::fold_after(x, M, i_);                                          \
  }
// the very first pass of the inverse transform: canonical input, stage m = 1 and half of stage
// m = 2 multiply by omega^0 = 1
#define ACE_R16_DIT_HEAD(x, TWF)                                      \
  _Pragma("unroll") for (int h_ = 0; h_ < 8; h_++) A::bfly_first(x[2 * h_], x[2 * h_ + 1], M); \
  A::fold_after(x, M, 0);                                             \
  _Pragma("unroll") for (int h_ = 0; h_ < 4; h_++) A::bfly_one(x[4 * h_], x[4 * h_ + 2], M); \
  {                                                                   \
    const typename A::TW tw_ = TWF(1, 1);                             \
    _Pragma("unroll") for (int h_ = 0; h_ < 4; h_++) A::bfly(x[4 * h_ + 1], x[4 * h_ + 3], tw_, M); \
  }                                                                   \
  _Pragma("unroll") for (int i_ = 2; i_ < 4; i_++) {                  \
    const int m_ = 1 << i_;                                           \
    _Pragma("unroll") for (int e_ = 0; e_ < m_; e_++) {               \
      const typename A::TW tw_ = TWF(i_, e_);                         \
      _Pragma("unroll") for (int h_ = 0; h_ < 8 / m_; h_++)           \
        A::bfly(x[2 * m_ * h_ + e_], x[2 * m_ * h_ + e_ + m_], tw_, M); \
    }                                                                 \
    if (i_ == 2) A::fold_after(x, M, 2);                              \
  }

constexpr int kThreads = 256;
constexpr int kRowPad  = 17 * 16;  // a 256-coefficient row in shared memory, 1 pad word per 16

// Programmatic dependent launch: the second kernel of a transform is launched while the first is
// still running; its CTAs fetch their twiddles and then wait here for the first kernel's data.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }

// PRE, bit 0 / bit 1: the 15 twiddles of the first / second pass are loaded up front, in parallel
// with the coefficients, instead of stage by stage.  The kernels wait on memory, not on the
// pipes (top stall: long scoreboard), so loads are issued as early as the registers allow: both
// passes for small batches (one CTA per SM, no register limit), the second pass for the FP64
// form of large batches (a twiddle is one double there), nothing for the integer form of large
// batches (60 more registers would cost a resident CTA).
#define ACE_TW_DECL typename A::TW twa_[(PRE & 1) ? 15 : 1], twb_[(PRE & 2) ? 15 : 1]
#define ACE_TW_PRELOAD(PTRA, PTRB, NA)                                               \
  _Pragma("unroll") for (int i_ = 0; i_ < 4; i_++)                                   \
    _Pragma("unroll") for (int h_ = 0; h_ < (1 << i_); h_++) {                       \
      if ((PRE & 1) && (1 << i_) - 1 + h_ >= 15 - (NA)) twa_[(1 << i_) - 1 + h_] = A::ld(PTRA(i_, h_)); \
      if (PRE & 2) twb_[(1 << i_) - 1 + h_] = A::ld(PTRB(i_, h_));                   \
    }
#define ACE_TW_GET_A(PTR, i, h) ((PRE & 1) ? twa_[(1 << (i)) - 1 + (h)] : A::ld(PTR(i, h)))
#define ACE_TW_GET_B(PTR, i, h) ((PRE & 2) ? twb_[(1 << (i)) - 1 + (h)] : A::ld(PTR(i, h)))

// ---------------- forward K1: stages 0-7 along r, 16 columns per CTA -------------------------
// XCH = 1 (cluster form, one launch per transform): the 16 CTAs of a cluster own one limb; the
// first phase hands its results to the CTA that needs them in the second phase through
// distributed shared memory (`xin` of the peer) instead of global memory.
template <class A, int PRE, class H, int XCH = 0>
__device__ __forceinline__ void fwd_cols_body(const DeviceTables& T, u32 g, u64* sm, const u64* in, u64* out, const H& hook,
                                              u64* xin = nullptr) {
  typedef typename A::E E;
  const typename A::Mod M = A::make(T.mod[g]);
  const typename A::TW* __restrict__ tw = A::fwd_table(T, g);
  const u32 t = threadIdx.x, c = t & 15, j = t >> 4;
  const u32 col = blockIdx.x * 16 + c;
  if (!XCH) pdl_launch_dependents();
#define PA(i, h) (tw + (1 << (i)) + (h))
#define PB(i, h) (tw + (16 << (i)) + (j << (i)) + (h))
  ACE_TW_DECL;
  ACE_TW_PRELOAD(PA, PB, 15)
  E x[16];
  if (!XCH) pdl_wait();  // the kernel before this transform in the stream (kernels.cuh: pdl_enter)
#pragma unroll
  for (int k = 0; k < 16; k++) x[k] = A::from_canonical(hook.pre(in[(j + 16 * k) * 256 + col]), M);
#define TW_A(i, h) ACE_TW_GET_A(PA, i, h)
  ACE_R16_FWD(x, TW_A)
#undef TW_A
#pragma unroll
  for (int k = 0; k < 16; k++) sm[(j + 16 * k) * 16 + c] = A::to_mid(x[k]);
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 16; k++) x[k] = A::from_mid(sm[(16 * j + k) * 16 + c], M);
#define TW_B(i, h) ACE_TW_GET_B(PB, i, h)
  ACE_R16_FWD(x, TW_B)
#undef TW_B
#undef PA
#undef PB
  if (XCH) {
    // rows 16j .. 16j+15 belong to CTA j of the cluster: row k of its tile, columns 16 s + c
    cooperative_groups::cluster_group cl = cooperative_groups::this_cluster();
    u64* peer = cl.map_shared_rank(xin, j);
    cl.barrier_wait();  // (arrived at kernel start) every CTA of the cluster is running
#pragma unroll
    for (int k = 0; k < 16; k++) peer[k * kRowPad + 17 * blockIdx.x + c] = A::to_mid(x[k]);
    return;
  }
#pragma unroll
  for (int k = 0; k < 16; k++) out[(16 * j + k) * 256 + col] = A::to_mid(x[k]);
}

// ---------------- forward K2: stages 8-15 inside a row, 16 rows per CTA ----------------------
template <class A, int PRE, class H, int XCH = 0>
__device__ __forceinline__ void fwd_rows_body(const DeviceTables& T, u32 g, u64* sm, u64* data, const H& hook,
                                              const u64* xin = nullptr) {
  typedef typename A::E E;
  const typename A::Mod M = A::make(T.mod[g]);
  const typename A::TW* __restrict__ tw = A::fwd_table(T, g);
  const u32 t = threadIdx.x, j = t & 15, rho = t >> 4;
  const u32 r = blockIdx.x * 16 + rho;
  u64* row = data + r * 256;
  u64* srow = sm + rho * kRowPad;
#define PA(i, h) (tw + (256 << (i)) + (r << (i)) + (h))
#define PB(i, h) (tw + (4096 << (i)) + ((16 * r + j) << (i)) + (h))
  u64* saux = sm + 16 * kRowPad;  // (fused batches only: the launch provides the room)
  ACE_TW_DECL;
  ACE_TW_PRELOAD(PA, PB, 15)
  E x[16];
  if (XCH) {
    if (H::kActive && hook.post_mode) stage_aux(saux, hook.aux + blockIdx.x * 4096);
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = A::from_mid(xin[rho * kRowPad + 17 * k + j], M);
  } else {
    pdl_wait();  // the coefficients come from the first kernel
    pdl_launch_dependents();  // whatever follows the transform may be launched now
    if (H::kActive && hook.post_mode) stage_aux(saux, hook.aux + blockIdx.x * 4096);
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = A::from_mid(row[j + 16 * k], M);
  }
#define TW_A(i, h) ACE_TW_GET_A(PA, i, h)
  ACE_R16_FWD(x, TW_A)
#undef TW_A
#pragma unroll
  for (int k = 0; k < 16; k++) srow[17 * k + j] = A::to_mid(x[k]);
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 16; k++) x[k] = A::from_mid(srow[17 * j + k], M);
#define TW_B(i, h) ACE_TW_GET_B(PB, i, h)
  ACE_R16_FWD(x, TW_B)
#undef TW_B
#undef PA
#undef PB
  ulonglong2* o = reinterpret_cast<ulonglong2*>(row + 16 * j);
  if (H::kActive && hook.post_mode) {
    stage_aux_wait();
    const ulonglong2* sa = reinterpret_cast<const ulonglong2*>(saux + rho * kAuxRow + kAuxGroup * j);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const u32 at = r * 256 + 16 * j + 2 * k;
      const ulonglong2 a = sa[k];
      o[k] = make_ulonglong2(hook.post(A::to_canonical(x[2 * k], M), a.x, at),
                             hook.post(A::to_canonical(x[2 * k + 1], M), a.y, at + 1));
    }
    return;
  }
#pragma unroll
  for (int k = 0; k < 8; k++)
    o[k] = make_ulonglong2(A::to_canonical(x[2 * k], M), A::to_canonical(x[2 * k + 1], M));
}

// ---------------- inverse K1: DIT stages m = 1 .. 128 inside a row ---------------------------
template <class A, int PRE, class H, int XCH = 0>
__device__ __forceinline__ void inv_rows_body(const DeviceTables& T, u32 g, u64* sm, const u64* in, u64* out, const H&,
                                              u64* xin = nullptr) {
  typedef typename A::E E;
  const typename A::Mod M = A::make(T.mod[g]);
  const typename A::TW* __restrict__ tw = A::inv_table(T, g);
  const u32 t = threadIdx.x, j = t & 15, rho = t >> 4;
  const u32 r = blockIdx.x * 16 + rho;
  u64* srow = sm + rho * kRowPad;
  if (!XCH) pdl_launch_dependents();
#define PA(i, e) (tw + (1 << (i)) + (e))
#define PB(i, e) (tw + (16 << (i)) + j + 16 * (e))
  ACE_TW_DECL;
  ACE_TW_PRELOAD(PA, PB, 13)
  E x[16];
  if (!XCH) pdl_wait();  // the kernel before this transform in the stream
  const ulonglong2* i2 = reinterpret_cast<const ulonglong2*>(in + r * 256 + 16 * j);
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const ulonglong2 v = i2[k];
    x[2 * k] = A::from_canonical(v.x, M); x[2 * k + 1] = A::from_canonical(v.y, M);
  }
#define TW_A(i, e) ACE_TW_GET_A(PA, i, e)
  ACE_R16_DIT_HEAD(x, TW_A)
#undef TW_A
#pragma unroll
  for (int k = 0; k < 16; k++) srow[17 * j + k] = A::to_mid(x[k]);
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 16; k++) x[k] = A::from_mid(srow[17 * k + j], M);
#define TW_B(i, e) ACE_TW_GET_B(PB, i, e)
  ACE_R16_DIT(x, TW_B)
#undef TW_B
#undef PA
#undef PB
  if (XCH) {
    // columns 16k .. 16k+15 belong to CTA k of the cluster: row r of its 256 x 16 slab
    cooperative_groups::cluster_group cl = cooperative_groups::this_cluster();
    cl.barrier_wait();
#pragma unroll
    for (int k = 0; k < 16; k++) cl.map_shared_rank(xin, k)[r * 16 + j] = A::to_mid(x[k]);
    return;
  }
  u64* orow = out + r * 256;
#pragma unroll
  for (int k = 0; k < 16; k++) orow[j + 16 * k] = A::to_mid(x[k]);
}

// ---------------- inverse K2: DIT stages m = 256 .. 32768 along r, then * psi^-n N^-1 ---------
template <class A, int PRE, class H, int XCH = 0>
__device__ __forceinline__ void inv_cols_body(const DeviceTables& T, u32 g, u64* sm, u64* data, const H&,
                                              const u64* xin = nullptr) {
  typedef typename A::E E;
  const typename A::Mod M = A::make(T.mod[g]);
  const typename A::TW* __restrict__ tw = A::inv_table(T, g);
  const typename A::TW* __restrict__ ps = A::scale_table(T, g);
  const u32 t = threadIdx.x, c = t & 15, j = t >> 4;
  const u32 col = blockIdx.x * 16 + c;
#define PA(i, e) (tw + (256 << (i)) + 256 * (e) + col)
#define PB(i, e) (tw + (4096 << (i)) + (j + 16 * (e)) * 256 + col)
  ACE_TW_DECL;
  ACE_TW_PRELOAD(PA, PB, 15)
  E x[16];
  if (XCH) {
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = A::from_mid(xin[(16 * j + k) * 16 + c], M);
  } else {
    pdl_wait();  // the coefficients come from the first kernel
    pdl_launch_dependents();
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = A::from_mid(data[(16 * j + k) * 256 + col], M);
  }
#define TW_A(i, e) ACE_TW_GET_A(PA, i, e)
  ACE_R16_DIT(x, TW_A)
#undef TW_A
#pragma unroll
  for (int k = 0; k < 16; k++) sm[(16 * j + k) * 16 + c] = A::to_mid(x[k]);
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 16; k++) x[k] = A::from_mid(sm[(j + 16 * k) * 16 + c], M);
#define TW_B(i, e) ACE_TW_GET_B(PB, i, e)
  ACE_R16_DIT(x, TW_B)
#undef TW_B
#undef PA
#undef PB
#pragma unroll
  for (int k = 0; k < 16; k++) {
    const u32 n = (j + 16 * k) * 256 + col;
    data[n] = A::scale_to_canonical(x[k], A::ld(ps + n), M);
  }
}

// which arithmetic a limb takes: 0 FP64, 1 integer lazy, 2 integer with conditional subtraction
__device__ __forceinline__ int arith_of(const DeviceTables& T, const Modulus& m) {
  if (T.ftwd != nullptr && m.q < T.fp64_max_q) return 0;
  return m.shift <= 55 ? 1 : 2;
}

enum { FWD_COLS = 0, FWD_ROWS = 1, INV_ROWS = 2, INV_COLS = 3 };

template <int KIND, class B, bool PRE>
__global__ void __launch_bounds__(kThreads, PRE ? 1 : 3) ntt16_kernel(DeviceTables T, const __grid_constant__ B b) {
  extern __shared__ __align__(16) u64 sm[];  // kSmem<KIND, B>() bytes
  // last limbs first: the special primes (slowest arithmetic) sit at the end of a batch and
  // should not be the tail of the launch
  const u32 limb = gridDim.y - 1 - blockIdx.y, g = b.g[limb];
  const int ar = arith_of(T, T.mod[g]);
  const u64* src = b_src(b, limb, 65536);
  u64* dst = b_dst(b, limb, 65536);
  constexpr int PD = PRE ? 3 : 2, PI = PRE ? 3 : 0;  // what is preloaded: FP64 / integer form
  typedef typename HookOf<B>::type H;
  const H hook = HookOf<B>::make(T, b, limb);
#define ACE_DISPATCH(BODY, ...)                                                  \
  if (ar == 0) BODY<ArithDP, PD, H>(__VA_ARGS__, hook);                          \
  else if (ar == 1) BODY<ArithInt<false>, PI, H>(__VA_ARGS__, hook);             \
  else BODY<ArithInt<true>, PI, H>(__VA_ARGS__, hook);
  if (KIND == FWD_COLS) { ACE_DISPATCH(fwd_cols_body, T, g, sm, src, dst) }
  if (KIND == FWD_ROWS) { ACE_DISPATCH(fwd_rows_body, T, g, sm, dst) }
  if (KIND == INV_ROWS) { ACE_DISPATCH(inv_rows_body, T, g, sm, src, dst) }
  if (KIND == INV_COLS) { ACE_DISPATCH(inv_cols_body, T, g, sm, dst) }
#undef ACE_DISPATCH
}

// ---------------- one launch per transform: clusters of 16 CTAs, hand-over through DSMEM ------
// grid (16, limbs), cluster (16, 1, 1): cluster = limb, CTA rank = slab (first phase) = row group
// (second phase).  Saves the write + read of the intermediate limb through L2 and the second
// launch.  16 CTAs per cluster is above the portable limit of 8 (opt-in,
// cudaFuncAttributeNonPortableClusterSizeAllowed).  MEASURED AND NOT USED BY DEFAULT: bit-exact,
// but 51.4 us against 38.5 us per 45 limbs and 14.5 against 12.4 us for one limb (16-CTA
// clusters constrain placement to one GPC at a time and the remote stores are slower than the
// L2 round trip they replace; ResNet-20 1.108 s against 0.951 s per image).  Kept behind
// ACE_B200_NTT_CLUSTER=1 as the record of the experiment (profiles/r2_ntt_bench_v4.txt).
template <int DIR, class B, bool PRE>
__global__ void __launch_bounds__(kThreads, PRE ? 1 : 3) ntt16_cluster_kernel(DeviceTables T, const __grid_constant__ B b) {
  extern __shared__ __align__(16) u64 dyn_sm[];
  u64* sm  = dyn_sm;                             // exchange between the two passes (+ epilogue operand)
  u64* xin = dyn_sm + 16 * kRowPad + 16 * kAuxRow;  // what the peers hand over for the second phase
  cooperative_groups::cluster_group cl = cooperative_groups::this_cluster();
  cl.barrier_arrive();  // matched by the barrier_wait() before the first remote store
  const u32 limb = gridDim.y - 1 - blockIdx.y, g = b.g[limb];
  const int ar = arith_of(T, T.mod[g]);
  const u64* src = b_src(b, limb, 65536);
  u64* dst = b_dst(b, limb, 65536);
  constexpr int PD = PRE ? 3 : 2, PI = PRE ? 3 : 0;
  typedef typename HookOf<B>::type H;
  const H hook = HookOf<B>::make(T, b, limb);
#define ACE_DISPATCH(BODY, ...)                                                     \
  if (ar == 0) BODY<ArithDP, PD, H, 1>(__VA_ARGS__);                                \
  else if (ar == 1) BODY<ArithInt<false>, PI, H, 1>(__VA_ARGS__);                   \
  else BODY<ArithInt<true>, PI, H, 1>(__VA_ARGS__);
  if (DIR == 0) { ACE_DISPATCH(fwd_cols_body, T, g, sm, src, dst, hook, xin) }
  else { ACE_DISPATCH(inv_rows_body, T, g, sm, src, dst, hook, xin) }
  cl.sync();  // release / acquire: every slab has arrived
  if (DIR == 0) { ACE_DISPATCH(fwd_rows_body, T, g, sm, dst, hook, xin) }
  else { ACE_DISPATCH(inv_cols_body, T, g, sm, dst, hook, xin) }
#undef ACE_DISPATCH
}

// ---------------- arithmetic-only ceiling: the radix-16 pass on registers, no memory ----------
// form 0: FP64 butterfly, 1: integer lazy, 2: integer with conditional subtraction
template <class A>
__global__ void __launch_bounds__(kThreads) bfly_peak_kernel(DeviceTables T, u32 g, u64* out, int iters) {
  typedef typename A::E E;
  const typename A::Mod M = A::make(T.mod[g]);
  const typename A::TW* tw = A::fwd_table(T, g);
  E x[16];
  typename A::TW twr[15];
#pragma unroll
  for (int k = 0; k < 16; k++) x[k] = A::from_canonical((threadIdx.x * 977 + k * 131) & 0xFFFFF, M);
#pragma unroll
  for (int k = 0; k < 15; k++) twr[k] = A::ld(tw + 1 + k);
  for (int it = 0; it < iters; it++) {
#define TW_R(i, h) twr[(1 << (i)) - 1 + (h)]
    ACE_R16_FWD(x, TW_R)
#undef TW_R
  }
  u64 s = 0;
#pragma unroll
  for (int k = 0; k < 16; k++) s ^= A::to_mid(x[k]);
  out[blockIdx.x * kThreads + threadIdx.x] = s;
}

}  // namespace

// G butterflies/s of the radix-16 register pass alone: the arithmetic ceiling the NTT kernels are
// compared with (form 0 FP64 -- needs the FP64 tables --, 1 integer lazy, 2 integer + csub)
double ntt16_bfly_peak(const DeviceTables& T, int form, int ctas_per_sm, cudaStream_t st) {
  if (form == 0 && T.ftwd == nullptr) return -1.0;
  cudaDeviceProp prop;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaGetDeviceProperties(&prop, dev);
  const int blocks = prop.multiProcessorCount * ctas_per_sm, iters = 200;
  u64* out = nullptr;
  cudaMalloc(&out, (size_t)blocks * kThreads * sizeof(u64));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  const u32 g = form == 2 ? T.G - 1 : 1;  // a special prime for the csub form, q_1 otherwise
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0, st);
    if (form == 0) bfly_peak_kernel<ArithDP><<<blocks, kThreads, 0, st>>>(T, g, out, iters);
    else if (form == 1) bfly_peak_kernel<ArithInt<false>><<<blocks, kThreads, 0, st>>>(T, g, out, iters);
    else bfly_peak_kernel<ArithInt<true>><<<blocks, kThreads, 0, st>>>(T, g, out, iters);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return (double)blocks * kThreads * iters * 32.0 / (best * 1e-3) * 1e-9;
}

bool ntt16_usable(const DeviceTables& T) { return T.logN == 16 && T.ftw2 != nullptr; }

// first kernel: plain launch; second kernel: programmatic dependent launch (it may start, and
// fetch its twiddles, before the first has finished; pdl_wait() orders the data)
template <int KIND, class B>
constexpr size_t kernel_smem() {
  return (KIND == FWD_COLS || KIND == INV_COLS)
             ? 4096 * sizeof(u64)
             : 16 * kRowPad * sizeof(u64) + (HookOf<B>::type::kActive && KIND == FWD_ROWS ? kAuxSmem : 0);
}
template <int K1, int K2, class B, bool PRE>
static void launch_pair(const DeviceTables& T, const B& b, cudaStream_t s) {
  dim3 grid(16, b.n);
  static bool attr = false;
  if (!attr) {  // the fused second kernel needs more than the 48 KB that are available by default
    cudaFuncSetAttribute(ntt16_kernel<K2, B, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)kernel_smem<K2, B>());
    attr = true;
  }
  launch_chain(ntt16_kernel<K1, B, PRE>, grid, dim3(kThreads), kernel_smem<K1, B>(), s, T, b);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kernel_smem<K2, B>();
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, ntt16_kernel<K2, B, PRE>, T, b);
}
constexpr size_t kClusterSmem = 2 * 16 * kRowPad * sizeof(u64) + kAuxSmem;  // exchange + hand-over + epilogue operand
// 0: not probed, 1: clusters of 16 can be placed, -1: they cannot (or ACE_B200_NTT_NO_CLUSTER)
template <int DIR, class B, bool PRE>
static bool launch_cluster(const DeviceTables& T, const B& b, cudaStream_t s) {
  static int state = 0;
  auto kern = ntt16_cluster_kernel<DIR, B, PRE>;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(16, b.n);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kClusterSmem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (state == 0) {
    state = -1;
    if (getenv("ACE_B200_NTT_CLUSTER") != nullptr &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kClusterSmem) == cudaSuccess &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n > 0) state = 1;
    }
    cudaGetLastError();
  }
  if (state < 0) return false;
  return cudaLaunchKernelEx(&cfg, kern, T, b) == cudaSuccess;
}
// batches of at most kSmallBatch limbs (<= 2 CTAs per SM) take the latency-oriented variant
constexpr u32 kSmallBatch = 18;
static bool no_pre() {
  static int v = -1;
  if (v < 0) v = getenv("ACE_B200_NTT_NO_PRE") ? 1 : 0;
  return v != 0;
}
// Forward transforms never take it: measured on B200 (profiles/r2_ntt_nopre.txt) the preloading
// variant costs 16.2 against 12.4 us at 8 limbs, 20.5 against 18.5 us at 18 and 24.4 against 21.2 us
// for the 11 special primes; the inverse gains 0.4 us at 8-9 limbs from it.
template <class B>
static void ntt16_fwd_impl(const DeviceTables& T, const B& b, cudaStream_t s) {
  static const bool fwd_pre = getenv("ACE_B200_NTT_FWD_PRE") != nullptr;
  if (b.n <= kSmallBatch && fwd_pre) {
    if (!launch_cluster<0, B, true>(T, b, s)) launch_pair<FWD_COLS, FWD_ROWS, B, true>(T, b, s);
  } else if (!launch_cluster<0, B, false>(T, b, s)) {
    launch_pair<FWD_COLS, FWD_ROWS, B, false>(T, b, s);
  }
}
template <class B>
static void ntt16_inv_impl(const DeviceTables& T, const B& b, cudaStream_t s) {
  if (b.n <= kSmallBatch && !no_pre()) {
    if (!launch_cluster<1, B, true>(T, b, s)) launch_pair<INV_ROWS, INV_COLS, B, true>(T, b, s);
  } else if (!launch_cluster<1, B, false>(T, b, s)) {
    launch_pair<INV_ROWS, INV_COLS, B, false>(T, b, s);
  }
}
void launch_ntt16_fused(const DeviceTables& T, const NttFusedBatch& b, cudaStream_t s) {
  if (b.n == 0) return;
  ntt16_fwd_impl(T, b, s);
}
void launch_ntt16(const DeviceTables& T, const LimbBatch& b, cudaStream_t s) { ntt16_fwd_impl(T, b, s); }
void launch_ntt16(const DeviceTables& T, const LimbPtrBatch& b, cudaStream_t s) { ntt16_fwd_impl(T, b, s); }
void launch_intt16(const DeviceTables& T, const LimbBatch& b, cudaStream_t s) { ntt16_inv_impl(T, b, s); }
void launch_intt16(const DeviceTables& T, const LimbPtrBatch& b, cudaStream_t s) { ntt16_inv_impl(T, b, s); }

}  // namespace ace
