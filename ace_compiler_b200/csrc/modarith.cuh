// modarith.cuh -- 64-bit RNS modular arithmetic for sm_100a (device) and the host setup code.
//
// Every routine returns the canonical residue in [0, q), so results are bit-identical to
// the reference's Barrett/Shoup routines (fhe-cmplr/rtlib/ant/include/util/fhe_utils.h:
// Add/Sub_int64_with_mod :192-215, Mod_barrett_128 :241-280, Mul_int64_mod_barret :290-300,
// Fast_mul_const_with_mod :311-318) whatever reduction strategy is used internally.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define ACE_HD __host__ __device__ __forceinline__
#define ACE_D __device__ __forceinline__
#else
#define ACE_HD inline
#define ACE_D inline
#endif

namespace ace {

typedef uint64_t           u64;
typedef uint32_t           u32;
typedef unsigned __int128  u128;

// Per-modulus constants (device copy lives in constant-like global memory, indexed by the
// global modulus index g: Q primes first, then P primes).
struct Modulus {
  u64 q;       // prime, < 2^62
  u64 mu_hi;   // floor(2^128 / q), high word   (Precompute_const_128, fhe_utils.h:385-401)
  u64 mu_lo;   //                  low word
  u64 mu64;    // floor(2^(62+nbits) / q): single-word Barrett ratio for products a*b, a,b<q
  u32 shift;   // nbits - 2, nbits = bit length of q
  u32 pad;
  u64 c64;     // 2^64 mod q
  u64 c64_sh;  // its Shoup companion floor(c64 * 2^64 / q)
};

ACE_HD u64 add_mod(u64 a, u64 b, u64 q) {
  u64 s = a + b;
  return s >= q ? s - q : s;
}
ACE_HD u64 sub_mod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }

#ifdef __CUDA_ARCH__
ACE_D u64 mul_hi64(u64 a, u64 b) { return __umul64hi(a, b); }
#else
inline u64 mul_hi64(u64 a, u64 b) { return (u64)(((u128)a * b) >> 64); }
#endif

// x * w mod q with w' = floor(w * 2^64 / q); valid for any x < 2^64, w < q.
ACE_HD u64 mul_shoup(u64 x, u64 w, u64 w_shoup, u64 q) {
  u64 hi = mul_hi64(x, w_shoup);
  u64 r  = x * w - hi * q;
  return r >= q ? r - q : r;
}
// same but leaves the result in [0, 2q) (lazy form for butterflies)
ACE_HD u64 mul_shoup_lazy(u64 x, u64 w, u64 w_shoup, u64 q) {
  u64 hi = mul_hi64(x, w_shoup);
  return x * w - hi * q;
}

// a * b mod q for a, b < q < 2^62: one-word Barrett on the product shifted by nbits-2.
// quotient estimate is within 2 of the true one, so two conditional subtractions finish it.
ACE_HD u64 mul_mod(u64 a, u64 b, const Modulus& m) {
  u64 lo = a * b;
  u64 hi = mul_hi64(a, b);
  u64 zs = (lo >> m.shift) | (hi << (64 - m.shift));
  u64 qh = mul_hi64(zs, m.mu64);
  u64 r  = lo - qh * m.q;
  r      = r >= 2 * m.q ? r - 2 * m.q : r;
  return r >= m.q ? r - m.q : r;
}

// 128-bit value (hi:lo) mod q:  hi * 2^64 + lo = hi * c64 + lo (mod q).  The first term is one
// Shoup product (exact for any hi < 2^64), the second a one-word Barrett step with
// mu_hi = floor(2^64 / q) (quotient estimate short by at most 1).  Canonical result, i.e. the same
// value as the reference's Mod_barrett_128 (fhe_utils.h:241-280), with 2 high + 3 low products
// instead of 5 + 4.
ACE_HD u64 reduce128(u64 lo, u64 hi, const Modulus& m) {
  const u64 r1 = mul_shoup(hi, m.c64, m.c64_sh, m.q);
  const u64 qh = mul_hi64(lo, m.mu_hi);
  u64 r0 = lo - qh * m.q;
  r0     = r0 >= m.q ? r0 - m.q : r0;
  return add_mod(r0, r1, m.q);
}

// 128-bit multiply-accumulate: (hi:lo) += a * b
ACE_HD void mac128(u64& lo, u64& hi, u64 a, u64 b) {
  u64 pl = a * b;
  u64 ph = mul_hi64(a, b);
  lo += pl;
  hi += ph + (lo < pl ? 1 : 0);
}

// Switch_modulus (fhe_utils.h:349-375): centred lift of val mod old_q into [0, new_q)
ACE_HD u64 switch_modulus(u64 val, u64 old_q, u64 new_q) {
  u64 half = old_q >> 1;
  if (new_q > old_q) {
    return val > half ? val + (new_q - old_q) : val;
  }
  u64 diff = new_q - (old_q % new_q);
  u64 r    = val > half ? val + diff : val;
  return r >= new_q ? r % new_q : r;
}

}  // namespace ace
