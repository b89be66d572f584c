// kernels.cuh -- launch wrappers of the hand-written sm_100a kernels (kernels.cu).
// All kernels work on limb-major RNS polynomials: limb l of a polynomial is the contiguous
// array data[l*N .. (l+1)*N) of canonical residues (int64 in the reference,
// fhe-cmplr/rtlib/ant/include/util/polynomial.h:35-44; u64 here, same bits).
#pragma once
#include <cuda_runtime.h>

#include "modarith.cuh"

namespace ace {

// ---- programmatic dependent launch along the whole stream ------------------------------------
// Every hot kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization and starts
// with pdl_enter(): griddepcontrol.wait (the kernel before it in the stream has completed and its
// writes are visible -- and, because that kernel waited as well, so has everything before it),
// then griddepcontrol.launch_dependents (the NEXT kernel may be launched as soon as every CTA of
// this one is running: its launch latency and CTA dispatch overlap this kernel's execution
// instead of following it).  Nothing is read from or written to global memory before the wait.
// An emitted ResNet-20 is ~46 000 dependent launches per image; ACE_B200_NO_PDL=1 turns the
// attribute off (plain stream order) for comparison.
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;");
}
bool pdl_chain_enabled();
template <class... P, class... A>
inline void launch_chain(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_chain_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, static_cast<A&&>(args)...);
}

constexpr int kMaxBatch = 192;  // limbs per batched launch (3 digits x 45 limbs fits)

// A batch of limbs: limb i is read from src + src_slot[i]*N, transformed modulo modulus g[i]
// and written to base + slot[i]*N.  src == nullptr means in place (src = base, same slots).
struct LimbBatch {
  u64*       base;
  const u64* src;
  u32        n;
  uint16_t   slot[kMaxBatch];
  uint16_t   src_slot[kMaxBatch];
  uint16_t   g[kMaxBatch];
};

// Device-resident per-context tables.
struct DeviceTables {
  u32            N, logN;
  u32            G;        // L + K
  u32            small_moduli;  // every modulus < 2^60: deferred-carry dot products are exact
  const Modulus* mod;      // [G]
  const u64*     tw;       // [G][N] psi powers, bit-reversed order   (ntt.c:95-101)
  const u64*     tw_sh;    // [G][N] Shoup companions                 (ntt.c:119-126)
  const u64*     itw;      // [G][N] inverse psi powers
  const u64*     itw_sh;
  const u64*     n_inv;    // [G]
  const u64*     n_inv_sh;
  // N = 2^16 transforms (ntt16.cu): twiddles interleaved with their Shoup companions {w, w'}
  const ulonglong2* ftw2;  // [G][N] forward, same order as tw
  const ulonglong2* itw2;  // [G][N] inverse, decimation in time: [m + j] = omega_(2m)^-j, j < m
  const ulonglong2* ips2;  // [G][N] psi^-n N^-1
  // the same three tables as exact integers in doubles, for the FP64 butterfly (moduli < fp64_max_q)
  const double* ftwd;
  const double* itwd;
  const double* ipsd;
  u64           fp64_max_q;
};

// Forward NTT with the neighbouring limb-wise steps of Rescale / Mod_down folded in (ntt16.cu):
//   prologue (pre = 1, Rescale_poly, polynomial.c:1123-1140): the input of limb i is
//       Switch_modulus(src, q_from, q_g) * pre_w[g_from * pre_stride + g]   (src = INTT of the dropped limb)
//   epilogue post = 1 (Rescale_poly, :1141-1161):  dst = aux * post_w[...] + NTT(x)
//            post = 2 (Reduce_rns_base tail, :953-965): dst = (aux - NTT(x)) * post_w[g] (+ add)
// so a Rescale is INTT(1 limb) + ONE fused transform instead of INTT + pre + NTT + post, and the
// tail of a ModDown disappears into its NTT: two passes over every limb less, two launches less.
constexpr int kMaxFused = 128;
struct NttFusedBatch {
  u32        n;
  u64*       dst[kMaxFused];
  const u64* src[kMaxFused];
  const u64* aux[kMaxFused];
  const u64* add[kMaxFused];   // post = 2 only; nullptr: none
  uint16_t   g[kMaxFused];
  uint16_t   g_from[kMaxFused];
  uint8_t    pre, post;
  const u64 *pre_w, *pre_w_sh, *post_w, *post_w_sh;
  u32        pre_stride, post_stride;  // table index = g_from * stride + g
};
void launch_ntt16_fused(const DeviceTables& T, const NttFusedBatch& b, cudaStream_t s);

// N = 2^16 only (ntt16.cu); launch_ntt / launch_intt route here when ntt16_usable()
bool ntt16_usable(const DeviceTables& T);
double ntt16_bfly_peak(const DeviceTables& T, int form, int ctas_per_sm, cudaStream_t st);
struct LimbBatch;
struct LimbPtrBatch;
void launch_ntt16(const DeviceTables& T, const LimbBatch& b, cudaStream_t s);
void launch_ntt16(const DeviceTables& T, const LimbPtrBatch& b, cudaStream_t s);
void launch_intt16(const DeviceTables& T, const LimbBatch& b, cudaStream_t s);
void launch_intt16(const DeviceTables& T, const LimbPtrBatch& b, cudaStream_t s);

void launch_ntt(const DeviceTables& T, const LimbBatch& b, cudaStream_t s);
void launch_intt(const DeviceTables& T, const LimbBatch& b, cudaStream_t s);

// The same transforms over limbs that live in different allocations (the scheduler batches
// the polynomials of many independent reference calls into one launch, sched.h): limb i is
// read from src[i], transformed modulo modulus g[i] and written to dst[i] (src[i] == dst[i]
// for an in-place transform).
constexpr int kMaxPtrBatch = 192;
struct LimbPtrBatch {
  u32        n;
  u64*       dst[kMaxPtrBatch];
  const u64* src[kMaxPtrBatch];
  uint16_t   g[kMaxPtrBatch];
};
void launch_ntt(const DeviceTables& T, const LimbPtrBatch& b, cudaStream_t s);
void launch_intt(const DeviceTables& T, const LimbPtrBatch& b, cudaStream_t s);

enum EwOp { EW_ADD = 0, EW_SUB = 1, EW_MUL = 2 };
// r[i] = a[i] op b[i] over n_limbs consecutive limbs, limb i uses modulus g0 + i
void launch_ew(const DeviceTables& T, EwOp op, u64* r, const u64* a, const u64* b, u32 g0,
               u32 n_limbs, cudaStream_t s);
// r[l][i] = a[l][order[i]]  (Hw_rotate with a non-negative order table, poly_arith.c:41-56)
void launch_gather(const DeviceTables& T, u64* r, const u64* a, const int64_t* order,
                   u32 g0, u32 n_limbs, cudaStream_t s);
// r = a * scalar[l] mod q_l, scalar given per limb with Shoup companion
void launch_mul_scalar(const DeviceTables& T, u64* r, const u64* a, const u64* sc,
                       const u64* sc_sh, u32 g0, u32 n_limbs, cudaStream_t s);

// Approximate fast base conversion (no correction term):
//   y_i   = x_i * hatinv_i mod b_i                     i < n_in
//   out_o = (sum_i y_i * hatmod[o][i]) mod t_o         o < n_out
// x: n_in limbs (coefficient form) at x + in_slot*N ..., moduli g_in[]; out limbs at
// out + out_slot[o]*N with moduli g_out[o].   (polynomial.c:755-807, 1297-1320)
struct ConvDesc {
  const u64* x;        // first input limb
  u64*       out;      // base of output polynomial
  const u64* hatinv;   // [n_in]
  const u64* hatinv_sh;
  const u64* hatmod;   // [n_out][n_in]
  u32        n_in, n_out;
  uint16_t   g_in[48];
  uint16_t   g_out[64];
  uint16_t   out_slot[64];
};
constexpr int kMaxConvPack = 8;  // descriptors per launch (n_desc <= kMaxConvPack)
void launch_base_conv(const DeviceTables& T, const ConvDesc* descs, u32 n_desc,
                      cudaStream_t s);

// key-switch inner product over digits (emitted loops, resnet20 .inc:7005-7032):
//   acc0[o] = sum_j ext_j[o] * key0_j[g(o)],  acc1 likewise, o < W = num_q + K
// ext: [beta][W][N]; key0/key1: [dnum][L+K][N]; g(o) = o < num_q ? o : L + o - num_q
// own != nullptr: the digit's own limbs (o in [j*part_size, (j+1)*part_size)) are read from
// own[o] (the key-switched polynomial itself) instead of ext_j[o].
void launch_ksw_inner(const DeviceTables& T, u64* acc0, u64* acc1, const u64* ext,
                      const u64* own, u32 part_size, const u64* key0, const u64* key1,
                      u32 beta, u32 num_q, u32 L, u32 K, cudaStream_t s);

// ModDown tail: out[l] = (old[l] - conv[l]) * pinv[l] (+ add[l] if add != nullptr)
void launch_moddown_tail(const DeviceTables& T, u64* out, const u64* old, const u64* conv,
                         const u64* add, const u64* pinv, const u64* pinv_sh, u32 n_limbs,
                         cudaStream_t s);

// d0 = a0 b0, d1 = a0 b1 + a1 b0, d2 = a1 b1 over n_limbs Q limbs (tensor product of two
// ciphertexts, ckks_evaluator.c:133-165) in one pass
void launch_moddown_tail2(const DeviceTables& T, u64* out0, u64* out1, const u64* old0, const u64* old1,
                          const u64* conv0, const u64* conv1, const u64* add0, const u64* add1, const u64* pinv,
                          const u64* pinv_sh, u32 n_limbs, cudaStream_t s);
void launch_tensor(const DeviceTables& T, u64* d0, u64* d1, u64* d2, const u64* a0,
                   const u64* a1, const u64* b0, const u64* b1, u32 n_limbs, cudaStream_t s);

// Rescale (polynomial.c:1097-1161):
//  pre : tmp[i] = switch_modulus(last, q_l, q_i) * negqlinv[i]          (coefficient form)
//  post: out[i] = c[i] * qlinv[i] + NTT(tmp)[i]
void launch_rescale_pre(const DeviceTables& T, u64* tmp, const u64* last, u32 l,
                        const u64* negqlinv, const u64* negqlinv_sh, cudaStream_t s);
void launch_rescale_post(const DeviceTables& T, u64* out, const u64* c, const u64* tmp,
                         const u64* qlinv, const u64* qlinv_sh, u32 n_limbs, cudaStream_t s);

// Batched forms of the three tails above for polynomials in different allocations: one
// descriptor per output limb y.
//   moddown tail : r = (a - b) * pinv[g]                      (a = old limb, b = converted limb)
//   rescale pre  : r = switch_modulus(a, q_aux, q_g) * negqlinv[aux][g]   (a = INTT of limb aux)
//   rescale post : r = a * qlinv[aux][g] + b
constexpr int kMaxP3 = 128;
struct Ptr3Batch {
  u32        n;
  u64*       r[kMaxP3];
  const u64* a[kMaxP3];
  const u64* b[kMaxP3];
  uint16_t   g[kMaxP3];
  uint16_t   aux[kMaxP3];
};
void launch_moddown_tail_batch(const DeviceTables& T, const Ptr3Batch& P, const u64* pinv,
                               const u64* pinv_sh, cudaStream_t s);
void launch_rescale_pre_batch(const DeviceTables& T, const Ptr3Batch& P, const u64* negqlinv,
                              const u64* negqlinv_sh, u32 L, cudaStream_t s);
void launch_rescale_post_batch(const DeviceTables& T, const Ptr3Batch& P, const u64* qlinv,
                               const u64* qlinv_sh, u32 L, cudaStream_t s);

}  // namespace ace

// ---- kernels_ext.cu: extended-basis (Q u P) variants and the bootstrap-only kernels ------
namespace ace {

// A polynomial over the extended basis is laid out [nq Q limbs | np P limbs] (Q limbs of the
// reference's POLYNOMIAL first, Get_p_coeffs after them, polynomial.h:214-217).  Limb y uses
// modulus index  y < nq ? y : pbase + (y - nq)   (pbase = L, the first P prime).
struct Basis {
  u32 nq, np, pbase;
  __host__ __device__ u32 width() const { return nq + np; }
  __host__ __device__ u32 g(u32 y) const { return y < nq ? y : pbase + (y - nq); }
};

// r = a op b over all limbs of the basis (Add_poly / Sub_poly / Multiply_ntt with p primes,
// polynomial.c:16-146)
void launch_ew_basis(const DeviceTables& T, EwOp op, u64* r, const u64* a, const u64* b,
                     Basis bs, cudaStream_t s);
// r[y][i] = a[y][order[i]] over all limbs (Automorphism_transform, polynomial.c:299-340)
void launch_ew_basis2(const DeviceTables& T, EwOp op, u64* r0, u64* r1, const u64* a0, const u64* a1,
                      const u64* b0, const u64* b1, Basis bs, cudaStream_t s);  // both polynomials, one launch
void launch_gather_basis(const DeviceTables& T, u64* r, const u64* a, const int64_t* order,
                         Basis bs, cudaStream_t s);
// Per-limb scalars passed by value (no staging copy, no sync): v[y] < q_y, sh[y] its Shoup
// companion floor(v * 2^64 / q).
struct ScalarPack {
  u64 v[64];
  u64 sh[64];
};
// r[y][i] = a[y][i] + v[y]   (Add_const: a constant plaintext in NTT form is the same residue
// in every coefficient, ckks_encoder.c:505-516; Add_plaintext, ckks_evaluator.c:103-118)
void launch_add_scalar(const DeviceTables& T, u64* r, const u64* a, const ScalarPack& sc,
                       u32 g0, u32 n_limbs, cudaStream_t s);
// r[y][i] = a[y][i] * v[y]   (Mul_const / Mul_integer, ckks_evaluator.c:208-232)
void launch_mul_scalar_pack(const DeviceTables& T, u64* r, const u64* a, const ScalarPack& sc,
                            Basis bs, cudaStream_t s);
// ModRaise (Transform_values_from_level0, ckks_bootstrap_context.c:1527-1550):
// out[0] = in, out[y] = Switch_modulus(in, q_0, q_y) for 0 < y < n_limbs; coefficient form
void launch_mul_scalar_pack2(const DeviceTables& T, u64* r0, u64* r1, const u64* a0, const u64* a1,
                             const ScalarPack& sc, Basis bs, cudaStream_t s);
void launch_mod_raise(const DeviceTables& T, u64* out, const u64* in, u32 n_limbs,
                      cudaStream_t s);

// out[y] = the monomial +-X^index in coefficient form over n_limbs Q limbs: coefficient `index`
// is 1 (or q_y - 1 when negative), all others 0   (Mul_by_monomial, ckks_evaluator.c:234-264)
void launch_monomial(const DeviceTables& T, u64* out, u32 index, bool negative, u32 n_limbs,
                     cudaStream_t s);

// Plaintext inner product of the baby-step/giant-step linear transform
// (Rotate_iteration, ckks_bootstrap_context.c:1299-1322):
//   out0 = sum_j a0[j] (.) pt[j],   out1 = sum_j a1[j] (.) pt[j]      over the basis `bs`
// a0/a1[j]: extended polynomials laid out [nq | np]; pt[j]: plaintext with its Q limbs at
// limb y and its P limbs starting at limb pt_pstart (a plaintext encoded at a higher level
// and derived down, Derive_plain).  128-bit accumulation, one reduction per output.
constexpr int kMaxDot = 32;
struct DotArgs {
  const u64* a0[kMaxDot];
  const u64* a1[kMaxDot];
  const u64* pt[kMaxDot];
  u32        n;
  u32        pt_pstart;
};
void launch_pt_dot(const DeviceTables& T, u64* out0, u64* out1, const DotArgs& args, Basis bs,
                   cudaStream_t s);

// acc0[y] += c0[y] * sc[y] on the Q limbs only, then nothing else: the "add_first" term of
// Fast_rotate_ext (ckks_evaluator.c:566-573); r may alias acc.
// out = sum_t w_t (.) ct_t over both polynomials (Eval_linear_wsum, ckks_chebyshev.c:282-323: Mul_const
// + Add_ciph per term): every term is read once, the sum is written once.  acc != 0: out is itself
// the first term (weight 1).  Canonical residues: the order of the modular additions does not matter.
constexpr int kMaxWsum = 12, kMaxWsumLimbs = 64;
struct WsumArgs {
  u32        n, acc;
  const u64* c0[kMaxWsum];
  const u64* c1[kMaxWsum];
  u64        w[kMaxWsum][kMaxWsumLimbs];
  u64        wsh[kMaxWsum][kMaxWsumLimbs];
};
void launch_ct_wsum(const DeviceTables& T, u64* out0, u64* out1, const WsumArgs& args, u32 n_limbs, cudaStream_t s);
void launch_ct_mul_plain_acc(const DeviceTables& T, u64* acc0, u64* acc1, const u64* in0, const u64* in1,
                             const u64* c0, const u64* c1, const u64* pt, u32 n_limbs, cudaStream_t s);
void launch_mul_scalar_add(const DeviceTables& T, u64* r, const u64* acc, const u64* c,
                           const u64* sc, const u64* sc_sh, u32 n_limbs, cudaStream_t s);

// ksw inner product fused with the epilogue of a rotation in the extended basis; see
// kernels_ext.cu.  c0 / scatter may be null; accN: add into outN instead of overwriting it.
void launch_ksw_inner_rot(const DeviceTables& T, u64* out0, u64* out1, const u64* ext,
                          const u64* own, u32 part_size, const u64* key0, const u64* key1,
                          u32 beta, u32 num_q, u32 L, u32 K, const u64* c0, const u64* pmodq,
                          const u64* pmodq_sh, const int64_t* scatter, bool acc0, bool acc1,
                          cudaStream_t s);
// r[y][i] += a[y][order[i]]
void launch_gather_add_basis(const DeviceTables& T, u64* r, const u64* a, const int64_t* order,
                             Basis bs, cudaStream_t s);

// all baby-step plaintext inner products of one BSGS level (see kernels_ext.cu):
// rotation j = (rot + j*rot_stride, rot + j*rot_stride + c1_offset); every pt[i*g + j] must be a
// valid plaintext (an all-zero one where the term is absent)
constexpr int kMaxDotBaby = 8, kMaxDotGiant = 16;
struct DotAllArgs {
  const u64* rot;
  size_t     rot_stride, c1_offset;
  const u64* pt[kMaxDotBaby * kMaxDotGiant];
  u64*       out0[kMaxDotBaby];
  u64*       out1[kMaxDotBaby];
  u32        b, g, pt_pstart;
};
void launch_pt_dot_all(const DeviceTables& T, const DotAllArgs& args, Basis bs, cudaStream_t s);

}  // namespace ace
