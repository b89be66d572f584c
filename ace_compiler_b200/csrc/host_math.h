// host_math.h -- host-side number theory for context setup (primes, roots, CRT constants).
// Restates what the reference computes with GMP + its own helpers, using only 128-bit
// integer arithmetic:
//   prime search          fhe-cmplr/rtlib/ant/src/util/crt.c:16-125
//   generator / psi       fhe-cmplr/rtlib/ant/src/util/number_theory.c:92-157
//   automorphism tables   fhe-cmplr/rtlib/ant/src/util/number_theory.c:187-225
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "modarith.cuh"

namespace ace {
namespace hm {

inline u64 mulmod(u64 a, u64 b, u64 m) { return (u64)((u128)a * b % m); }
inline u64 powmod(u64 a, u64 e, u64 m) {
  u64 r = 1;
  a %= m;
  for (; e; e >>= 1) {
    if (e & 1) r = mulmod(r, a, m);
    a = mulmod(a, a, m);
  }
  return r;
}
inline u64 invmod_prime(u64 a, u64 p) { return powmod(a % p, p - 2, p); }
inline u64 shoup(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }  // Precompute_const

// Deterministic Miller-Rabin (exact for 64-bit inputs).  The reference's Is_prime
// (number_theory.c:159-185) is the randomised version of the same test.
inline bool is_prime(u64 n) {
  static const u64 bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
  if (n < 2) return false;
  for (u64 b : bases) {
    if (n % b == 0) return n == b;
  }
  u64 d = n - 1;
  int s = 0;
  while (!(d & 1)) { d >>= 1; s++; }
  for (u64 b : bases) {
    u64 x = powmod(b, d, n);
    if (x == 1 || x == n - 1) continue;
    bool composite = true;
    for (int r = 1; r < s && composite; r++) {
      x = mulmod(x, x, n);
      if (x == n - 1) composite = false;
    }
    if (composite) return false;
  }
  return true;
}

inline u64 first_prime(u64 N, size_t bits) {  // crt.c:16-24
  u64 p = ((u64)1 << bits) + 2 * N + 1;
  while (!is_prime(p)) p += 2 * N;
  return p;
}
inline u64 previous_prime(u64 from, u64 step) {  // crt.c:26-32
  u64 p = from - step;
  while (!is_prime(p)) p -= step;
  return p;
}
inline u64 next_prime(u64 from, u64 step) {  // crt.c:34-42 (first candidate tested: from+2*step)
  u64 p = from + step;
  do { p += step; } while (!is_prime(p));
  return p;
}

// Q chain (crt.c:91-125): last prime just above 2^sf_bits, then alternately below/above,
// q0 below the first prime above 2^first_bits (or continuing downwards if sizes agree).
inline std::vector<u64> q_chain(size_t count, size_t first_bits, size_t sf_bits, u64 N) {
  std::vector<u64> q(count);
  const u64 step = 2 * N;
  u64 lowest = first_prime(N, sf_bits), highest = lowest;
  q[count - 1] = lowest;
  for (size_t k = 0; count >= 2 && k + 2 < count; k++) {
    size_t i = count - 2 - k;
    if (k % 2 == 0) {
      lowest = previous_prime(lowest, step);
      q[i]   = lowest;
    } else {
      highest = next_prime(highest, step);
      q[i]    = highest;
    }
  }
  q[0] = first_bits == sf_bits ? previous_prime(lowest, step)
                               : previous_prime(first_prime(N, first_bits), step);
  return q;
}

// bit length of a product of word-sized factors (mpz_sizeinbase(.,2) in crt.c:383)
inline size_t product_bit_length(const u64* f, size_t n) {
  std::vector<u64> limbs(1, 1);
  for (size_t i = 0; i < n; i++) {
    u64 carry = 0;
    for (auto& w : limbs) {
      u128 t = (u128)w * f[i] + carry;
      w      = (u64)t;
      carry  = (u64)(t >> 64);
    }
    if (carry) limbs.push_back(carry);
  }
  return (limbs.size() - 1) * 64 + (64 - (size_t)__builtin_clzll(limbs.back()));
}

// smallest generator of Z_q^* (number_theory.c:92-130)
inline u64 smallest_generator(u64 q) {
  const u64 phi = q - 1;
  std::vector<u64> pf;
  u64 rest = phi;
  for (u64 d = 2; d <= (u64)std::sqrt((double)rest); d++) {
    if (rest % d == 0) {
      pf.push_back(d);
      while (rest % d == 0) rest /= d;
    }
  }
  if (rest > 1) pf.push_back(rest);
  for (u64 r = 2; r <= phi; r++) {
    bool ok = true;
    for (u64 f : pf) {
      if (powmod(r, phi / f, q) == 1) { ok = false; break; }
    }
    if (ok) return r;
  }
  return 0;
}

inline u32 bit_reverse(u32 x, u32 bits) {
  u32 r = 0;
  for (u32 i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
  return r;
}

}  // namespace hm
}  // namespace ace
