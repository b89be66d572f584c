// kernels_ext.cu -- element-wise kernels over the extended basis Q u P and the kernels only the
// CKKS bootstrap needs (ModRaise, scalar add, plaintext inner product).  All HBM-bound: one
// thread per coefficient, grid.y = limb, every output limb written exactly once.
#include "kernels.cuh"
#include "prof.h"

namespace ace {

static inline dim3 grid_for(const DeviceTables& T, u32 n_limbs) {
  return dim3((T.N + 255) / 256, n_limbs);
}

template <int OP>
__global__ void __launch_bounds__(256) ew_basis_kernel(DeviceTables T, u64* __restrict__ r,
                                                       const u64* __restrict__ a,
                                                       const u64* __restrict__ b, Basis bs) {
  pdl_enter();
  const Modulus m   = T.mod[bs.g(blockIdx.y)];
  const size_t  off = (size_t)blockIdx.y * T.N;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < T.N; i += gridDim.x * blockDim.x) {
    u64 x = a[off + i], y = b[off + i], z;
    if (OP == EW_ADD) z = add_mod(x, y, m.q);
    if (OP == EW_SUB) z = sub_mod(x, y, m.q);
    if (OP == EW_MUL) z = mul_mod(x, y, m);
    r[off + i] = z;
  }
}

void launch_ew_basis(const DeviceTables& T, EwOp op, u64* r, const u64* a, const u64* b,
                     Basis bs, cudaStream_t s) {
  prof::Scope prof_scope_("ew_basis", s);
  if (bs.width() == 0) return;
  dim3 grid = grid_for(T, bs.width());
  switch (op) {
    case EW_ADD: launch_chain(ew_basis_kernel<EW_ADD>, grid, 256, 0, s, T, r, a, b, bs); break;
    case EW_SUB: launch_chain(ew_basis_kernel<EW_SUB>, grid, 256, 0, s, T, r, a, b, bs); break;
    case EW_MUL: launch_chain(ew_basis_kernel<EW_MUL>, grid, 256, 0, s, T, r, a, b, bs); break;
  }
}

// both polynomials of a ciphertext in one launch (blockIdx.z = polynomial): r_z = a_z op b_z
template <int OP>
__global__ void __launch_bounds__(256) ew_basis2_kernel(DeviceTables T, u64* r0, u64* r1, const u64* a0,
                                                        const u64* a1, const u64* b0, const u64* b1, Basis bs) {
  pdl_enter();
  u64* __restrict__       r = blockIdx.z ? r1 : r0;
  const u64* __restrict__ a = blockIdx.z ? a1 : a0;
  const u64* __restrict__ b = blockIdx.z ? b1 : b0;
  const Modulus m   = T.mod[bs.g(blockIdx.y)];
  const size_t  off = (size_t)blockIdx.y * T.N;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < T.N; i += gridDim.x * blockDim.x) {
    u64 x = a[off + i], y = b[off + i], z;
    if (OP == EW_ADD) z = add_mod(x, y, m.q);
    if (OP == EW_SUB) z = sub_mod(x, y, m.q);
    if (OP == EW_MUL) z = mul_mod(x, y, m);
    r[off + i] = z;
  }
}
void launch_ew_basis2(const DeviceTables& T, EwOp op, u64* r0, u64* r1, const u64* a0, const u64* a1,
                      const u64* b0, const u64* b1, Basis bs, cudaStream_t s) {
  prof::Scope prof_scope_("ew_basis", s);
  if (bs.width() == 0) return;
  dim3 grid = grid_for(T, bs.width());
  grid.z = 2;
  switch (op) {
    case EW_ADD: launch_chain(ew_basis2_kernel<EW_ADD>, grid, 256, 0, s, T, r0, r1, a0, a1, b0, b1, bs); break;
    case EW_SUB: launch_chain(ew_basis2_kernel<EW_SUB>, grid, 256, 0, s, T, r0, r1, a0, a1, b0, b1, bs); break;
    case EW_MUL: launch_chain(ew_basis2_kernel<EW_MUL>, grid, 256, 0, s, T, r0, r1, a0, a1, b0, b1, bs); break;
  }
}

__global__ void __launch_bounds__(256) gather_basis_kernel(DeviceTables T, u64* __restrict__ r,
                                                           const u64* __restrict__ a,
                                                           const int64_t* __restrict__ order,
                                                           Basis bs) {
  pdl_enter();
  const u64    q   = T.mod[bs.g(blockIdx.y)].q;
  const size_t off = (size_t)blockIdx.y * T.N;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < T.N; i += gridDim.x * blockDim.x) {
    int64_t k = order[i];
    r[off + i] = k >= 0 ? a[off + k] : q - a[off - k];
  }
}

void launch_gather_basis(const DeviceTables& T, u64* r, const u64* a, const int64_t* order,
                         Basis bs, cudaStream_t s) {
  prof::Scope prof_scope_("gather_basis", s);
  if (bs.width() == 0) return;
  launch_chain(gather_basis_kernel, grid_for(T, bs.width()), 256, 0, s, T, r, a, order, bs);
}

__global__ void __launch_bounds__(256) add_scalar_kernel(DeviceTables T, u64* __restrict__ r,
                                                         const u64* __restrict__ a,
                                                         ScalarPack sc, u32 g0) {
  pdl_enter();
  const u64    q   = T.mod[g0 + blockIdx.y].q;
  const u64    v   = sc.v[blockIdx.y];
  const size_t off = (size_t)blockIdx.y * T.N;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < T.N; i += gridDim.x * blockDim.x)
    r[off + i] = add_mod(a[off + i], v, q);
}

void launch_add_scalar(const DeviceTables& T, u64* r, const u64* a, const ScalarPack& sc,
                       u32 g0, u32 n_limbs, cudaStream_t s) {
  prof::Scope prof_scope_("add_scalar", s);
  if (n_limbs == 0) return;
  launch_chain(add_scalar_kernel, grid_for(T, n_limbs), 256, 0, s, T, r, a, sc, g0);
}

__global__ void __launch_bounds__(256) mul_scalar_pack_kernel(DeviceTables T,
                                                              u64* __restrict__ r,
                                                              const u64* __restrict__ a,
                                                              ScalarPack sc, Basis bs) {
  pdl_enter();
  const u64    q   = T.mod[bs.g(blockIdx.y)].q;
  const u64    w = sc.v[blockIdx.y], wsh = sc.sh[blockIdx.y];
  const size_t off = (size_t)blockIdx.y * T.N;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < T.N; i += gridDim.x * blockDim.x)
    r[off + i] = mul_shoup(a[off + i], w, wsh, q);
}

void launch_mul_scalar_pack(const DeviceTables& T, u64* r, const u64* a, const ScalarPack& sc,
                            Basis bs, cudaStream_t s) {
  prof::Scope prof_scope_("mul_scalar_pack", s);
  if (bs.width() == 0) return;
  launch_chain(mul_scalar_pack_kernel, grid_for(T, bs.width()), 256, 0, s, T, r, a, sc, bs);
}

__global__ void __launch_bounds__(256) mul_scalar_pack2_kernel(DeviceTables T, u64* r0, u64* r1, const u64* a0,
                                                               const u64* a1, const __grid_constant__ ScalarPack sc,
                                                               Basis bs) {
  pdl_enter();
  u64* __restrict__       r = blockIdx.z ? r1 : r0;
  const u64* __restrict__ a = blockIdx.z ? a1 : a0;
  const u64    q   = T.mod[bs.g(blockIdx.y)].q;
  const u64    w = sc.v[blockIdx.y], wsh = sc.sh[blockIdx.y];
  const size_t off = (size_t)blockIdx.y * T.N;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < T.N; i += gridDim.x * blockDim.x)
    r[off + i] = mul_shoup(a[off + i], w, wsh, q);
}
void launch_mul_scalar_pack2(const DeviceTables& T, u64* r0, u64* r1, const u64* a0, const u64* a1,
                             const ScalarPack& sc, Basis bs, cudaStream_t s) {
  prof::Scope prof_scope_("mul_scalar_pack", s);
  if (bs.width() == 0) return;
  dim3 grid = grid_for(T, bs.width());
  grid.z = 2;
  launch_chain(mul_scalar_pack2_kernel, grid, 256, 0, s, T, r0, r1, a0, a1, sc, bs);
}

__global__ void __launch_bounds__(256) mod_raise_kernel(DeviceTables T, u64* __restrict__ out,
                                                        const u64* __restrict__ in) {
  pdl_enter();
  const u32    y   = blockIdx.y;
  const u64    q0 = T.mod[0].q, qy = T.mod[y].q;
  const size_t off = (size_t)y * T.N;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < T.N; i += gridDim.x * blockDim.x) {
    u64 v = in[i];
    out[off + i] = y == 0 ? v : switch_modulus(v, q0, qy);
  }
}

void launch_mod_raise(const DeviceTables& T, u64* out, const u64* in, u32 n_limbs,
                      cudaStream_t s) {
  prof::Scope prof_scope_("mod_raise", s);
  if (n_limbs == 0) return;
  launch_chain(mod_raise_kernel, grid_for(T, n_limbs), 256, 0, s, T, out, in);
}

__global__ void __launch_bounds__(256) monomial_kernel(DeviceTables T, u64* __restrict__ out,
                                                       u32 index, u32 negative) {
  pdl_enter();
  const u64    q   = T.mod[blockIdx.y].q;
  const size_t off = (size_t)blockIdx.y * T.N;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < T.N; i += gridDim.x * blockDim.x)
    out[off + i] = i == index ? (negative ? q - 1 : 1) : 0;
}

void launch_monomial(const DeviceTables& T, u64* out, u32 index, bool negative, u32 n_limbs,
                     cudaStream_t s) {
  prof::Scope prof_scope_("monomial", s);
  if (n_limbs == 0) return;
  launch_chain(monomial_kernel, grid_for(T, n_limbs), 256, 0, s, T, out, index, negative ? 1u : 0u);
}

__global__ void __launch_bounds__(256) pt_dot_kernel(DeviceTables T, u64* __restrict__ out0,
                                                     u64* __restrict__ out1, DotArgs A,
                                                     Basis bs) {
  pdl_enter();
  const u32     y    = blockIdx.y;
  const Modulus m    = T.mod[bs.g(y)];
  const size_t  off  = (size_t)y * T.N;
  const size_t  poff = (size_t)(y < bs.nq ? y : A.pt_pstart + (y - bs.nq)) * T.N;
  const u32     i    = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T.N) return;
  u64 lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0;
  for (u32 j = 0; j < A.n; j++) {
    const u64 p = A.pt[j][poff + i];
    mac128(lo0, hi0, A.a0[j][off + i], p);
    mac128(lo1, hi1, A.a1[j][off + i], p);
  }
  out0[off + i] = reduce128(lo0, hi0, m);
  out1[off + i] = reduce128(lo1, hi1, m);
}

void launch_pt_dot(const DeviceTables& T, u64* out0, u64* out1, const DotArgs& args, Basis bs,
                   cudaStream_t s) {
  prof::Scope prof_scope_("pt_dot", s);
  if (bs.width() == 0 || args.n == 0) return;
  launch_chain(pt_dot_kernel, grid_for(T, bs.width()), 256, 0, s, T, out0, out1, args, bs);
}

__global__ void __launch_bounds__(256) mul_scalar_add_kernel(
    DeviceTables T, u64* __restrict__ r, const u64* __restrict__ acc, const u64* __restrict__ c,
    const u64* __restrict__ sc, const u64* __restrict__ sc_sh) {
  pdl_enter();
  const u64    q   = T.mod[blockIdx.y].q;
  const u64    w = sc[blockIdx.y], wsh = sc_sh[blockIdx.y];
  const size_t off = (size_t)blockIdx.y * T.N;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < T.N; i += gridDim.x * blockDim.x)
    r[off + i] = add_mod(acc[off + i], mul_shoup(c[off + i], w, wsh, q), q);
}

void launch_mul_scalar_add(const DeviceTables& T, u64* r, const u64* acc, const u64* c,
                           const u64* sc, const u64* sc_sh, u32 n_limbs, cudaStream_t s) {
  prof::Scope prof_scope_("mul_scalar_add", s);
  if (n_limbs == 0) return;
  launch_chain(mul_scalar_add_kernel, grid_for(T, n_limbs), 256, 0, s, T, r, acc, c, sc, sc_sh);
}

__global__ void __launch_bounds__(256) ct_wsum_kernel(DeviceTables T, u64* out0, u64* out1,
                                                      const __grid_constant__ WsumArgs A) {
  pdl_enter();
  const u32    y   = blockIdx.y;
  const u64    q   = T.mod[y].q;
  const size_t off = (size_t)y * T.N;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < T.N; i += gridDim.x * blockDim.x) {
    u64 s0 = A.acc ? out0[off + i] : 0, s1 = A.acc ? out1[off + i] : 0;
#pragma unroll 4
    for (u32 t = 0; t < A.n; t++) {
      const u64 w = A.w[t][y], wsh = A.wsh[t][y];
      s0 = add_mod(s0, mul_shoup(A.c0[t][off + i], w, wsh, q), q);
      s1 = add_mod(s1, mul_shoup(A.c1[t][off + i], w, wsh, q), q);
    }
    out0[off + i] = s0;
    out1[off + i] = s1;
  }
}
void launch_ct_wsum(const DeviceTables& T, u64* out0, u64* out1, const WsumArgs& args, u32 n_limbs, cudaStream_t s) {
  prof::Scope prof_scope_("ct_wsum", s);
  if (n_limbs == 0 || args.n == 0) return;
  launch_chain(ct_wsum_kernel, grid_for(T, n_limbs), 256, 0, s, T, out0, out1, args);
}

// acc += ct (.) pt over all limbs of both polynomials: the plaintext limb is read once for c0 and
// c1 (emitted code: 2 Hw_modmul + 2 Hw_modadd per limb and term of a convolution).  acc may be
// the first operand of a fresh sum (acc_in == nullptr: acc = ct (.) pt).
__global__ void __launch_bounds__(256) ct_mul_plain_acc_kernel(DeviceTables T, u64* acc0, u64* acc1,
                                                               const u64* acc_in0, const u64* acc_in1,
                                                               const u64* __restrict__ c0,
                                                               const u64* __restrict__ c1,
                                                               const u64* __restrict__ pt) {
  pdl_enter();
  const Modulus m   = T.mod[blockIdx.y];
  const size_t  off = (size_t)blockIdx.y * T.N;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < T.N; i += gridDim.x * blockDim.x) {
    const u64 p = pt[off + i];
    u64 z0 = mul_mod(c0[off + i], p, m), z1 = mul_mod(c1[off + i], p, m);
    if (acc_in0) { z0 = add_mod(acc_in0[off + i], z0, m.q); z1 = add_mod(acc_in1[off + i], z1, m.q); }
    acc0[off + i] = z0;
    acc1[off + i] = z1;
  }
}
void launch_ct_mul_plain_acc(const DeviceTables& T, u64* acc0, u64* acc1, const u64* in0, const u64* in1,
                             const u64* c0, const u64* c1, const u64* pt, u32 n_limbs, cudaStream_t s) {
  prof::Scope prof_scope_("ct_mul_plain_acc", s);
  if (n_limbs == 0) return;
  launch_chain(ct_mul_plain_acc_kernel, grid_for(T, n_limbs), 256, 0, s, T, acc0, acc1, in0, in1, c0, c1, pt);
}

// Key inner product with the epilogue of a "fast" rotation in the extended basis
// (Fast_rotate_ext, ckks_evaluator.c:537-577 = Fast_switch_key_ext + P*c0 + automorphism):
//   v0[o] = sum_j ext_j[o] key0_j[g(o)]  (+ c0[o] * (P mod q_o) on the Q limbs)
//   v1[o] = sum_j ext_j[o] key1_j[g(o)]
//   out[o][scatter[n]] (+)= v[o][n]      scatter = the inverse automorphism table
// One pass instead of inner product, scalar multiply-add, two gathers and two additions; every
// value is the same canonical residue the separate kernels produce.
// STAGED: the automorphism permutation of the (bit-reversed) evaluation domain maps every aligned
// block of 2^m consecutive positions onto another such block (the low bits of the exponent
// 2 brv(i) + 1 are the reversed HIGH bits of i, and multiplication by an odd k keeps low bits among
// themselves).  A CTA owns 256 consecutive source positions: it permutes its results inside shared
// memory and writes (and, when accumulating, reads) the destination block with coalesced accesses
// instead of 8-byte scattered ones (round 1: 0.54 of the HBM peak, top stall long_scoreboard).
template <bool STAGED>
__global__ void __launch_bounds__(STAGED ? 128 : 256) ksw_inner_rot_kernel(
    DeviceTables T, u64* __restrict__ out0, u64* __restrict__ out1, const u64* __restrict__ ext,
    const u64* __restrict__ own, u32 part_size, const u64* __restrict__ key0,
    const u64* __restrict__ key1, u32 beta, u32 num_q, u32 L, u32 K, const u64* __restrict__ c0,
    const u64* __restrict__ pmodq, const u64* __restrict__ pmodq_sh,
    const int64_t* __restrict__ scatter, int acc0_flag, int acc1_flag) {
  pdl_enter();
  const u32     o = blockIdx.y;
  const u32     g = o < num_q ? o : L + (o - num_q);
  const u32     W = num_q + K;
  const Modulus m = T.mod[g];
  if (STAGED) {
    // 128 threads x 2 consecutive coefficients: every stream (beta digits, two key polynomials)
    // is read with 16-byte loads; the block of 256 results is permuted in shared memory
    __shared__ u64 st0[256], st1[256];
    __shared__ u32 blk;
    const u32 n = blockIdx.x * 256 + 2 * threadIdx.x;
    u64 lo0[2] = {0, 0}, hi0[2] = {0, 0}, lo1[2] = {0, 0}, hi1[2] = {0, 0};
#pragma unroll 3
    for (u32 j = 0; j < beta; j++) {
      const bool mine = own != nullptr && o < num_q && o / part_size == j;
      const u64* ep = mine ? own + (size_t)o * T.N + n : ext + ((size_t)j * W + o) * T.N + n;
      const ulonglong2 e  = *reinterpret_cast<const ulonglong2*>(ep);
      const ulonglong2 k0 = *reinterpret_cast<const ulonglong2*>(key0 + ((size_t)j * (L + K) + g) * T.N + n);
      const ulonglong2 k1 = *reinterpret_cast<const ulonglong2*>(key1 + ((size_t)j * (L + K) + g) * T.N + n);
      mac128(lo0[0], hi0[0], e.x, k0.x); mac128(lo0[1], hi0[1], e.y, k0.y);
      mac128(lo1[0], hi1[0], e.x, k1.x); mac128(lo1[1], hi1[1], e.y, k1.y);
    }
    u64 v0[2], v1[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      v0[h] = reduce128(lo0[h], hi0[h], m);
      v1[h] = reduce128(lo1[h], hi1[h], m);
    }
    if (c0 != nullptr && o < num_q) {
      const ulonglong2 c = *reinterpret_cast<const ulonglong2*>(c0 + (size_t)o * T.N + n);
      v0[0] = add_mod(v0[0], mul_shoup(c.x, pmodq[o], pmodq_sh[o], m.q), m.q);
      v0[1] = add_mod(v0[1], mul_shoup(c.y, pmodq[o], pmodq_sh[o], m.q), m.q);
    }
    u32 tg[2] = {n, n + 1};
    if (scatter) {
      const longlong2 sc = *reinterpret_cast<const longlong2*>(scatter + n);
      tg[0] = (u32)sc.x; tg[1] = (u32)sc.y;
    }
    st0[tg[0] & 255] = v0[0]; st0[tg[1] & 255] = v0[1];
    st1[tg[0] & 255] = v1[0]; st1[tg[1] & 255] = v1[1];
    if (threadIdx.x == 0) blk = tg[0] & ~255u;
    __syncthreads();
    const size_t pos = (size_t)o * T.N + blk + 2 * threadIdx.x;
    ulonglong2 r0 = make_ulonglong2(st0[2 * threadIdx.x], st0[2 * threadIdx.x + 1]);
    ulonglong2 r1 = make_ulonglong2(st1[2 * threadIdx.x], st1[2 * threadIdx.x + 1]);
    if (acc0_flag) {
      const ulonglong2 p = *reinterpret_cast<const ulonglong2*>(out0 + pos);
      r0.x = add_mod(p.x, r0.x, m.q); r0.y = add_mod(p.y, r0.y, m.q);
    }
    if (acc1_flag) {
      const ulonglong2 p = *reinterpret_cast<const ulonglong2*>(out1 + pos);
      r1.x = add_mod(p.x, r1.x, m.q); r1.y = add_mod(p.y, r1.y, m.q);
    }
    *reinterpret_cast<ulonglong2*>(out0 + pos) = r0;
    *reinterpret_cast<ulonglong2*>(out1 + pos) = r1;
    return;
  }
  const u32 n = blockIdx.x * blockDim.x + threadIdx.x;
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
  u64 v0 = reduce128(lo0, hi0, m), v1 = reduce128(lo1, hi1, m);
  if (c0 != nullptr && o < num_q)
    v0 = add_mod(v0, mul_shoup(c0[(size_t)o * T.N + n], pmodq[o], pmodq_sh[o], m.q), m.q);
  const size_t pos = (size_t)o * T.N + (scatter ? (u32)scatter[n] : n);
  if (acc0_flag) v0 = add_mod(out0[pos], v0, m.q);
  if (acc1_flag) v1 = add_mod(out1[pos], v1, m.q);
  out0[pos] = v0;
  out1[pos] = v1;
}

void launch_ksw_inner_rot(const DeviceTables& T, u64* out0, u64* out1, const u64* ext,
                          const u64* own, u32 part_size, const u64* key0, const u64* key1,
                          u32 beta, u32 num_q, u32 L, u32 K, const u64* c0, const u64* pmodq,
                          const u64* pmodq_sh, const int64_t* scatter, bool acc0, bool acc1,
                          cudaStream_t s) {
  prof::Scope prof_scope_("ksw_inner_rot", s);
  dim3 grid((T.N + 255) / 256, num_q + K);
  if (T.N % 256 == 0)
    launch_chain(ksw_inner_rot_kernel<true>, grid, 128, 0, s, T, out0, out1, ext, own, part_size, key0, key1, beta,
                                                    num_q, L, K, c0, pmodq, pmodq_sh, scatter,
                                                    acc0 ? 1 : 0, acc1 ? 1 : 0);
  else
    launch_chain(ksw_inner_rot_kernel<false>, grid, 256, 0, s, T, out0, out1, ext, own, part_size, key0, key1, beta,
                                                     num_q, L, K, c0, pmodq, pmodq_sh, scatter,
                                                     acc0 ? 1 : 0, acc1 ? 1 : 0);
}

// r[y][i] += a[y][order[i]] over all limbs of the basis (automorphism + accumulate)
__global__ void __launch_bounds__(256) gather_add_basis_kernel(DeviceTables T, u64* __restrict__ r,
                                                               const u64* __restrict__ a,
                                                               const int64_t* __restrict__ order,
                                                               Basis bs) {
  pdl_enter();
  const u64    q   = T.mod[bs.g(blockIdx.y)].q;
  const size_t off = (size_t)blockIdx.y * T.N;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < T.N; i += gridDim.x * blockDim.x) {
    const int64_t k = order[i];
    const u64 v = k >= 0 ? a[off + k] : q - a[off - k];
    r[off + i] = add_mod(r[off + i], v, q);
  }
}
void launch_gather_add_basis(const DeviceTables& T, u64* r, const u64* a, const int64_t* order,
                             Basis bs, cudaStream_t s) {
  prof::Scope prof_scope_("gather_add_basis", s);
  launch_chain(gather_add_basis_kernel, grid_for(T, bs.width()), 256, 0, s, T, r, a, order, bs);
}

// All baby-step inner sums of one BSGS level in one pass over the giant-step rotations:
//   out_i = sum_j rot_j (.) pt[i*g + j]        i < b, j < g  (absent terms: a zero plaintext)
// Each rotation limb is read once instead of b times; the sums are the same exact 128-bit
// accumulations as pt_dot_kernel's, so the results are identical.
template <int B>
__global__ void __launch_bounds__(256) pt_dot_all_kernel(DeviceTables T,
                                                         const __grid_constant__ DotAllArgs A,
                                                         Basis bs) {
  pdl_enter();
  const u32     y    = blockIdx.y;
  const Modulus m    = T.mod[bs.g(y)];
  const size_t  off  = (size_t)y * T.N;
  const size_t  poff = (size_t)(y < bs.nq ? y : A.pt_pstart + (y - bs.nq)) * T.N;
  const u32     n    = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= T.N) return;
  u64 lo0[B], hi0[B], lo1[B], hi1[B];
#pragma unroll
  for (int i = 0; i < B; i++) lo0[i] = hi0[i] = lo1[i] = hi1[i] = 0;
  // every table entry is a valid pointer (absent terms point at a zero plaintext), so the B + 2
  // loads of a step are independent and go out together -- and the loads of step j + 1 go out
  // before the multiply-accumulates of step j (the kernel waits on memory, not on the pipes):
  // 412 -> 396 us per launch at the ResNet set.  (Two coefficients per thread with 16-byte loads
  // instead: 451 us -- 126 registers leave two CTAs per SM.)
  u64 x0 = A.rot[off + n], x1 = A.rot[A.c1_offset + off + n], p[B];
#pragma unroll
  for (int i = 0; i < B; i++) p[i] = A.pt[(i < (int)A.b ? i : 0) * A.g][poff + n];
  for (u32 j = 0; j < A.g; j++) {
    const u32 jn = j + 1 < A.g ? j + 1 : j;  // (the last step reloads its own operands)
    const u64 nx0 = A.rot[(size_t)jn * A.rot_stride + off + n];
    const u64 nx1 = A.rot[(size_t)jn * A.rot_stride + A.c1_offset + off + n];
    u64 np[B];
#pragma unroll
    for (int i = 0; i < B; i++) np[i] = A.pt[(i < (int)A.b ? i : 0) * A.g + jn][poff + n];
#pragma unroll
    for (int i = 0; i < B; i++) {
      if (i < (int)A.b) {
        mac128(lo0[i], hi0[i], x0, p[i]);
        mac128(lo1[i], hi1[i], x1, p[i]);
      }
    }
    x0 = nx0; x1 = nx1;
#pragma unroll
    for (int i = 0; i < B; i++) p[i] = np[i];
  }
#pragma unroll
  for (int i = 0; i < B; i++) {
    if (i < (int)A.b) {
      A.out0[i][off + n] = reduce128(lo0[i], hi0[i], m);
      A.out1[i][off + n] = reduce128(lo1[i], hi1[i], m);
    }
  }
}

void launch_pt_dot_all(const DeviceTables& T, const DotAllArgs& args, Basis bs, cudaStream_t s) {
  prof::Scope prof_scope_("pt_dot_all", s);
  if (bs.width() == 0 || args.b == 0 || args.g == 0) return;
  if (args.b <= 4) launch_chain(pt_dot_all_kernel<4>, grid_for(T, bs.width()), 256, 0, s, T, args, bs);
  else launch_chain(pt_dot_all_kernel<kMaxDotBaby>, grid_for(T, bs.width()), 256, 0, s, T, args, bs);
}

}  // namespace ace
