/* oracle/ckks_oracle.c -- TEST INFRASTRUCTURE ONLY (see ckks_oracle.h).
 *
 * Plain-C restatement of the reference algorithms on the CKKS evaluation hot path.  Every
 * value the reference stores is the canonical residue of an exactly specified integer, so
 * this port computes with `unsigned __int128 %` everywhere instead of the reference's
 * Shoup/Barrett shortcuts; the residues are identical (checked limb-for-limb against the
 * compiled reference in tests/test_oracle_vs_ref.py).
 * Paths are relative to /root/reference/fhe-cmplr/rtlib/ant/.
 */
#include "ckks_oracle.h"
#include "rou_table.h"

#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t          u64;

#define AUXBITS 60 /* include/util/fhe_types.h:27-29 */

struct orc_ctx {
  uint32_t N, logN;
  size_t   L, K, dnum, part_size, sf_bits, first_bits;
  u64*     mod;     /* L+K moduli: Q then P */
  u64*     psi;     /* 2N-th root per modulus */
  u64**    rou;     /* rou[g][bitrev(i)] = psi^i          (src/util/ntt.c:95-101) */
  u64**    rou_inv; /* same for psi^-1                    (src/util/ntt.c:104-117) */
  u64*     n_inv;
  double complex* fft_rou; /* exp(2*pi*i*k/(2N)), k<2N    (src/util/ntt.c:585-593) */
  u64*            rot_group;
};

/* ---- scalar modular arithmetic ------------------------------------------- */
static u64 mulmod(u64 a, u64 b, u64 m) { return (u64)((u128)a * b % m); }
static u64 addmod(u64 a, u64 b, u64 m) { u64 s = a + b; return s >= m ? s - m : s; }
static u64 submod(u64 a, u64 b, u64 m) { return a >= b ? a - b : a + m - b; }
static u64 powmod(u64 a, u64 e, u64 m) {
  u64 r = 1;
  a %= m;
  while (e) {
    if (e & 1) r = mulmod(r, a, m);
    a = mulmod(a, a, m);
    e >>= 1;
  }
  return r;
}
static u64 invmod(u64 a, u64 m) { return powmod(a, m - 2, m); } /* number_theory.c:53-56 */

/* deterministic Miller-Rabin for 64-bit (stands in for the randomised Is_prime,
 * number_theory.c:159-185; both decide primality) */
static int is_prime(u64 n) {
  static const u64 bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
  if (n < 2) return 0;
  for (size_t i = 0; i < 12; i++) {
    if (n % bases[i] == 0) return n == bases[i];
  }
  u64 d = n - 1;
  int s = 0;
  while ((d & 1) == 0) { d >>= 1; s++; }
  for (size_t i = 0; i < 12; i++) {
    u64 x = powmod(bases[i], d, n);
    if (x == 1 || x == n - 1) continue;
    int comp = 1;
    for (int r = 1; r < s; r++) {
      x = mulmod(x, x, n);
      if (x == n - 1) { comp = 0; break; }
    }
    if (comp) return 0;
  }
  return 1;
}

/* src/util/crt.c:16-42 */
static u64 gen_first_prime(u64 N, size_t bits) {
  u64 order = 2 * N, p = ((u64)1 << bits) + order + 1;
  while (!is_prime(p)) p += order;
  return p;
}
static u64 gen_prev_prime(u64 m, u64 order) {
  u64 p = m - order;
  while (!is_prime(p)) p -= order;
  return p;
}
static u64 gen_next_prime(u64 m, u64 order) {
  u64 p = m + order; /* NB: the candidate m+order itself is never tested (crt.c:35-41) */
  do { p += order; } while (!is_prime(p));
  return p;
}

/* src/util/crt.c:91-125 */
static void gen_q_primes(u64* q, size_t n, size_t first_bits, size_t sf_bits, u64 N) {
  u64 order = 2 * N, cand = gen_first_prime(N, sf_bits);
  q[n - 1]   = cand;
  u64 q_next = cand, q_prev = cand;
  if (n > 1) {
    uint32_t cnt = 0;
    for (size_t i = n - 2; i >= 1; i--) {
      if ((cnt & 1) == 0) {
        q_prev = gen_prev_prime(q_prev, order);
        cand   = q_prev;
      } else {
        q_next = gen_next_prime(q_next, order);
        cand   = q_next;
      }
      q[i] = cand;
      cnt++;
    }
  }
  if (first_bits == sf_bits) {
    q[0] = gen_prev_prime(q_prev, order);
  } else {
    q[0] = gen_prev_prime(gen_first_prime(N, first_bits), order);
  }
}

/* bit length of a product of 64-bit primes (BI_SIZE_INBASE, crt.c:383) */
static size_t product_bits(const u64* f, size_t n) {
  u64    limbs[64] = {1};
  size_t len       = 1;
  for (size_t i = 0; i < n; i++) {
    u64 carry = 0;
    for (size_t k = 0; k < len; k++) {
      u128 t   = (u128)limbs[k] * f[i] + carry;
      limbs[k] = (u64)t;
      carry    = (u64)(t >> 64);
    }
    if (carry) limbs[len++] = carry;
  }
  return (len - 1) * 64 + (64 - (size_t)__builtin_clzll(limbs[len - 1]));
}

/* src/util/number_theory.c:92-157 */
static u64 find_generator(u64 q) {
  u64    phi = q - 1, number = phi, factor[64];
  size_t nf = 0;
  for (u64 i = 2; i <= (u64)sqrt((double)number); i++) {
    if (number % i == 0) {
      factor[nf++] = i;
      while (number % i == 0) number /= i;
    }
  }
  if (number > 1) factor[nf++] = number;
  for (u64 r = 2; r <= phi; r++) {
    int bad = 0;
    for (size_t i = 0; i < nf; i++) {
      if (powmod(r, phi / factor[i], q) == 1) { bad = 1; break; }
    }
    if (!bad) return r;
  }
  return 0;
}

static uint32_t bitrev(uint32_t x, uint32_t bits) {
  uint32_t r = 0;
  for (uint32_t i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
  return r;
}

orc_ctx* orc_create(uint32_t degree, size_t mul_depth, size_t first_mod_size,
                    size_t scaling_mod_size, size_t num_q_parts) {
  orc_ctx* c = (orc_ctx*)calloc(1, sizeof(orc_ctx));
  c->N       = degree;
  c->logN    = (uint32_t)round(log2(degree));
  c->L       = mul_depth + 1;
  c->dnum    = num_q_parts;
  c->sf_bits = scaling_mod_size;
  c->first_bits = first_mod_size;
  u64* q     = (u64*)calloc(c->L, sizeof(u64));
  gen_q_primes(q, c->L, first_mod_size, scaling_mod_size, degree);
  /* Precompute_qpart, crt.c:353-396 */
  c->part_size    = (size_t)ceil((double)c->L / (double)num_q_parts);
  size_t max_bits = 0;
  for (size_t j = 0; j < num_q_parts; j++) {
    size_t lo = j * c->part_size, hi = lo + c->part_size;
    if (hi > c->L) hi = c->L;
    size_t bits = product_bits(q + lo, hi - lo);
    if (bits > max_bits) max_bits = bits;
  }
  c->K   = (size_t)ceil((double)max_bits / AUXBITS);
  c->mod = (u64*)calloc(c->L + c->K, sizeof(u64));
  memcpy(c->mod, q, c->L * sizeof(u64));
  free(q);
  /* Generate_p_primes, crt.c:46-77 */
  u64 prev = gen_first_prime(degree, AUXBITS);
  for (size_t i = 0; i < c->K; i++) {
    u64 cand;
    int dup;
    do {
      cand = gen_prev_prime(prev, 2 * (u64)degree);
      dup  = 0;
      for (size_t j = 0; j < c->L; j++) dup |= (cand == c->mod[j]);
      prev = cand;
    } while (dup);
    c->mod[c->L + i] = cand;
  }
  /* Precompute_ntt, ntt.c:80-127 */
  size_t G   = c->L + c->K;
  c->psi     = (u64*)calloc(G, sizeof(u64));
  c->n_inv   = (u64*)calloc(G, sizeof(u64));
  c->rou     = (u64**)calloc(G, sizeof(u64*));
  c->rou_inv = (u64**)calloc(G, sizeof(u64*));
  for (size_t g = 0; g < G; g++) {
    u64 m     = c->mod[g];
    /* Root_of_unity (number_theory.c:140-157): fixed-root table first, then g^((q-1)/2N) */
    c->psi[g] = orc_fixed_root(2 * (u64)degree, m);
    if (c->psi[g] == 0) {
      u64 gen   = find_generator(m);
      c->psi[g] = powmod(gen, (m - 1) / (2 * (u64)degree), m);
    }
    c->n_inv[g]   = invmod(degree % m, m);
    c->rou[g]     = (u64*)calloc(degree, sizeof(u64));
    c->rou_inv[g] = (u64*)calloc(degree, sizeof(u64));
    u64 pinv = invmod(c->psi[g], m), pw = 1, pwi = 1;
    for (uint32_t i = 0; i < degree; i++) {
      uint32_t r       = bitrev(i, c->logN);
      c->rou[g][r]     = pw;
      c->rou_inv[g][r] = pwi;
      pw               = mulmod(pw, c->psi[g], m);
      pwi              = mulmod(pwi, pinv, m);
    }
  }
  /* Precompute_fft with fft_length = 2N (ntt.c:585-610) */
  size_t M     = 2 * (size_t)degree;
  c->fft_rou   = (double complex*)calloc(M, sizeof(double complex));
  c->rot_group = (u64*)calloc(degree / 2, sizeof(u64));
  for (size_t i = 0; i < M; i++) {
    double angle  = 2 * M_PI * i / M;
    c->fft_rou[i] = cos(angle) + sin(angle) * I;
  }
  c->rot_group[0] = 1;
  for (size_t i = 1; i < degree / 2; i++) c->rot_group[i] = (5 * c->rot_group[i - 1]) % M;
  return c;
}

void orc_destroy(orc_ctx* c) {
  if (!c) return;
  for (size_t g = 0; g < c->L + c->K; g++) {
    free(c->rou[g]);
    free(c->rou_inv[g]);
  }
  free(c->rou); free(c->rou_inv); free(c->psi); free(c->n_inv); free(c->mod);
  free(c->fft_rou); free(c->rot_group);
  free(c);
}
uint32_t orc_degree(const orc_ctx* c) { return c->N; }
size_t   orc_num_q(const orc_ctx* c) { return c->L; }
size_t   orc_num_p(const orc_ctx* c) { return c->K; }
size_t   orc_part_size(const orc_ctx* c) { return c->part_size; }
void     orc_get_primes(const orc_ctx* c, int64_t* q, int64_t* p) {
  for (size_t i = 0; i < c->L; i++) q[i] = (int64_t)c->mod[i];
  for (size_t i = 0; i < c->K; i++) p[i] = (int64_t)c->mod[c->L + i];
}
int64_t orc_psi(const orc_ctx* c, int is_p, size_t idx) {
  return (int64_t)c->psi[is_p ? c->L + idx : idx];
}

/* Forward_transform, src/util/ntt.c:190-264: Cooley-Tukey, natural in -> bit-reversed out */
void orc_ntt(const orc_ctx* c, size_t g, int64_t* data) {
  u64  q = c->mod[g];
  u64* a = (u64*)data;
  uint32_t n = c->N, t = n >> 1;
  for (uint32_t m = 1; m < n; m <<= 1, t >>= 1) {
    for (uint32_t i = 0; i < m; i++) {
      u64 w = c->rou[g][m + i];
      for (uint32_t j = 2 * i * t; j < 2 * i * t + t; j++) {
        u64 v    = mulmod(a[j + t], w, q);
        u64 u    = a[j];
        a[j]     = addmod(u, v, q);
        a[j + t] = submod(u, v, q);
      }
    }
  }
}

/* Inverse_transform, src/util/ntt.c:268-353: Gentleman-Sande, bit-reversed in -> natural out,
 * N^-1 folded in */
void orc_intt(const orc_ctx* c, size_t g, int64_t* data) {
  u64  q = c->mod[g];
  u64* a = (u64*)data;
  uint32_t n = c->N, t = 1;
  for (uint32_t m = n >> 1; m >= 1; m >>= 1, t <<= 1) {
    for (uint32_t i = 0; i < m; i++) {
      u64 w = c->rou_inv[g][m + i];
      for (uint32_t j = 2 * i * t; j < 2 * i * t + t; j++) {
        u64 u = a[j], v = a[j + t];
        a[j]     = addmod(u, v, q);
        a[j + t] = mulmod(submod(u, v, q), w, q);
      }
    }
  }
  for (uint32_t j = 0; j < n; j++) a[j] = mulmod(a[j], c->n_inv[g], q);
}

/* src/poly/poly_arith.c:14-56 */
void orc_hw_modadd(const orc_ctx* c, int64_t* r, const int64_t* a, const int64_t* b, size_t g) {
  for (uint32_t i = 0; i < c->N; i++) r[i] = (int64_t)addmod((u64)a[i], (u64)b[i], c->mod[g]);
}
void orc_hw_modmul(const orc_ctx* c, int64_t* r, const int64_t* a, const int64_t* b, size_t g) {
  for (uint32_t i = 0; i < c->N; i++) r[i] = (int64_t)mulmod((u64)a[i], (u64)b[i], c->mod[g]);
}
void orc_hw_rotate(const orc_ctx* c, int64_t* r, const int64_t* a, const int64_t* order,
                   size_t g) {
  for (uint32_t i = 0; i < c->N; i++) {
    int64_t k = order[i];
    r[i]      = k >= 0 ? a[k] : (int64_t)c->mod[g] - a[-k];
  }
}

/* Find_automorphism_index, src/util/number_theory.c:187-199 (modulus 2N) */
uint32_t orc_auto_index(const orc_ctx* c, int32_t rot_idx) {
  u64 M = 2 * (u64)c->N;
  if (rot_idx == 0) return 1;
  if (rot_idx == (int32_t)(M - 1)) return (uint32_t)rot_idx;
  u64 gen = 5;
  if (rot_idx < 0) { /* inverse of 5 mod 2N */
    u64 x = 1;
    for (u64 e = M / 2 - 1, b = 5; e; e >>= 1, b = b * b % M) {
      if (e & 1) x = x * b % M;
    }
    gen = x; /* 5^(phi(2N)-1) = 5^(N-1) */
  }
  u64 r = 1, b = gen;
  for (u64 e = (u64)(rot_idx < 0 ? -rot_idx : rot_idx); e; e >>= 1, b = b * b % M) {
    if (e & 1) r = r * b % M;
  }
  return (uint32_t)r;
}

/* Precompute_automorphism_order with is_ntt = TRUE, number_theory.c:201-214 */
void orc_auto_order(const orc_ctx* c, uint32_t k, int64_t* order) {
  size_t n = c->N, logm = c->logN + 1;
  for (size_t j = 0; j < n; j++) {
    size_t jt  = (j << 1) + 1;
    size_t idx = ((jt * k) - (((jt * k) >> logm) << logm)) >> 1;
    order[bitrev((uint32_t)j, c->logN)] = (int64_t)bitrev((uint32_t)idx, c->logN);
  }
}

size_t orc_num_decomp(const orc_ctx* c, size_t num_q) { /* polynomial.h:158-168 */
  size_t n = (num_q + c->part_size - 1) / c->part_size;
  return n > c->dnum ? c->dnum : n;
}

/* Decompose_modup, src/util/polynomial.c:1241-1335 with tables of crt.c:399-533.
 * in: num_q limbs NTT form. out: num_q + K limbs (Q limbs then P limbs), NTT form. */
void orc_decomp_modup(const orc_ctx* c, int64_t* out, const int64_t* in, size_t num_q,
                      size_t part) {
  size_t N = c->N, ps = c->part_size, beta = orc_num_decomp(c, num_q);
  size_t start = ps * part;
  size_t np2   = (part == beta - 1) ? num_q - ps * part : ps;
  /* part2: copy + INTT */
  u64* x = (u64*)malloc(sizeof(u64) * N * np2);
  for (size_t i = 0; i < np2; i++) {
    memcpy(out + (start + i) * N, in + (start + i) * N, sizeof(u64) * N);
    memcpy(x + i * N, in + (start + i) * N, sizeof(u64) * N);
    orc_intt(c, start + i, (int64_t*)(x + i * N));
  }
  /* y_i = x_i * (Qpart/q_i)^-1 mod q_i  (_l_hat_inv_modq[part][np2-1][i]) */
  for (size_t i = 0; i < np2; i++) {
    u64 qi = c->mod[start + i], hat = 1;
    for (size_t k = 0; k < np2; k++) {
      if (k != i) hat = mulmod(hat, c->mod[start + k] % qi, qi);
    }
    u64 hinv = invmod(hat, qi);
    for (size_t n = 0; n < N; n++) x[i * N + n] = mulmod(x[i * N + n], hinv, qi);
  }
  /* complement limbs: Q limbs [0,num_q) outside the part, then the K P limbs */
  for (size_t o = 0; o < num_q + c->K; o++) {
    if (o >= start && o < start + np2) continue;
    size_t g = o < num_q ? o : c->L + (o - num_q);
    u64    t = c->mod[g], hatmod[64];
    for (size_t i = 0; i < np2; i++) { /* _l_hat_modp[num_q-1][part][i][.] */
      u64 h = 1;
      for (size_t k = 0; k < np2; k++) {
        if (k != i) h = mulmod(h, c->mod[start + k] % t, t);
      }
      hatmod[i] = h;
    }
    u64* dst = (u64*)out + o * N;
    for (size_t n = 0; n < N; n++) {
      u128 sum = 0;
      for (size_t i = 0; i < np2; i++) sum += (u128)x[i * N + n] * hatmod[i];
      dst[n] = (u64)(sum % t);
    }
    orc_ntt(c, g, (int64_t*)dst);
  }
  free(x);
}

/* Reduce_rns_base + Fast_base_conv, src/util/polynomial.c:928-967, 755-807.
 * in: num_q + K limbs NTT form; out: num_q limbs. `in` is NOT modified here. */
void orc_mod_down(const orc_ctx* c, int64_t* out, const int64_t* in, size_t num_q) {
  size_t N = c->N, K = c->K, L = c->L;
  u64*   y = (u64*)malloc(sizeof(u64) * N * K);
  for (size_t i = 0; i < K; i++) {
    u64 pi = c->mod[L + i], hat = 1;
    memcpy(y + i * N, in + (num_q + i) * N, sizeof(u64) * N);
    orc_intt(c, L + i, (int64_t*)(y + i * N));
    for (size_t k = 0; k < K; k++) {
      if (k != i) hat = mulmod(hat, c->mod[L + k] % pi, pi);
    }
    u64 hinv = invmod(hat, pi);
    for (size_t n = 0; n < N; n++) y[i * N + n] = mulmod(y[i * N + n], hinv, pi);
  }
  for (size_t j = 0; j < num_q; j++) {
    u64 qj = c->mod[j], hatmod[64], pprod = 1;
    for (size_t i = 0; i < K; i++) {
      u64 h = 1;
      for (size_t k = 0; k < K; k++) {
        if (k != i) h = mulmod(h, c->mod[L + k] % qj, qj);
      }
      hatmod[i] = h;
      pprod     = mulmod(pprod, c->mod[L + i] % qj, qj);
    }
    u64  pinv = invmod(pprod, qj);
    u64* dst  = (u64*)out + j * N;
    for (size_t n = 0; n < N; n++) {
      u128 sum = 0;
      for (size_t i = 0; i < K; i++) sum += (u128)y[i * N + n] * hatmod[i];
      dst[n] = (u64)(sum % qj);
    }
    orc_ntt(c, j, (int64_t*)dst);
    for (size_t n = 0; n < N; n++) {
      dst[n] = mulmod(submod((u64)in[j * N + n], dst[n], qj), pinv, qj);
    }
  }
  free(y);
}

/* Switch_modulus, include/util/fhe_utils.h:349-375 */
static u64 switch_modulus(u64 val, u64 old_mod, u64 new_mod) {
  u64 res = val, half = old_mod >> 1;
  if (new_mod > old_mod) {
    if (res > half) res += new_mod - old_mod;
  } else {
    u64 diff = new_mod - (old_mod % new_mod);
    if (res > half) res += diff;
    if (res >= new_mod) res %= new_mod;
  }
  return res;
}

/* Rescale_poly (NTT branch), src/util/polynomial.c:1097-1161.  The table entry
 * _ql_ql_inv_mod_ql_div_ql_mod_qi (crt.c:300-317) equals -q_l^-1 mod q_i (the floor of
 * (Q/q_l * [(Q/q_l)^-1]_{q_l}) / q_l is k with k*q_l = -1 mod q_i).
 * in: num_q limbs; out: num_q - 1 limbs. */
void orc_rescale(const orc_ctx* c, int64_t* out, const int64_t* in, size_t num_q) {
  size_t N = c->N, l = num_q - 1;
  u64    ql = c->mod[l];
  u64*   last = (u64*)malloc(sizeof(u64) * N);
  u64*   tmp  = (u64*)malloc(sizeof(u64) * N);
  memcpy(last, in + l * N, sizeof(u64) * N);
  orc_intt(c, l, (int64_t*)last);
  for (size_t i = 0; i < l; i++) {
    u64 qi = c->mod[i], qlinv = invmod(ql % qi, qi), neg = qi - qlinv;
    for (size_t n = 0; n < N; n++) tmp[n] = mulmod(switch_modulus(last[n], ql, qi), neg, qi);
    orc_ntt(c, i, (int64_t*)tmp);
    for (size_t n = 0; n < N; n++) {
      out[i * N + n] = (int64_t)addmod(mulmod((u64)in[i * N + n], qlinv, qi), tmp[n], qi);
    }
  }
  free(last);
  free(tmp);
}

/* body of the emitted Rotate()/Relinearize() up to Mod_down
 * (dataset/resnet20_cifar10_pre.onnx.inc:6996-7036) */
void orc_key_switch(const orc_ctx* c, int64_t* out0, int64_t* out1, const int64_t* d,
                    size_t num_q, const int64_t* key0, const int64_t* key1) {
  size_t N = c->N, K = c->K, L = c->L, W = num_q + K, beta = orc_num_decomp(c, num_q);
  int64_t* ext  = (int64_t*)calloc(W * N, sizeof(int64_t));
  int64_t* acc0 = (int64_t*)calloc(W * N, sizeof(int64_t));
  int64_t* acc1 = (int64_t*)calloc(W * N, sizeof(int64_t));
  int64_t* tmp  = (int64_t*)calloc(N, sizeof(int64_t));
  for (size_t p = 0; p < beta; p++) {
    orc_decomp_modup(c, ext, d, num_q, p);
    const int64_t* k0 = key0 + p * (L + K) * N;
    const int64_t* k1 = key1 + p * (L + K) * N;
    for (size_t o = 0; o < W; o++) {
      size_t g = o < num_q ? o : L + (o - num_q); /* key P limbs sit at Poly_level(key)=L */
      orc_hw_modmul(c, tmp, k0 + g * N, ext + o * N, g);
      orc_hw_modadd(c, acc0 + o * N, acc0 + o * N, tmp, g);
      orc_hw_modmul(c, tmp, k1 + g * N, ext + o * N, g);
      orc_hw_modadd(c, acc1 + o * N, acc1 + o * N, tmp, g);
    }
  }
  orc_mod_down(c, out0, acc0, num_q);
  orc_mod_down(c, out1, acc1, num_q);
  free(ext); free(acc0); free(acc1); free(tmp);
}

void orc_ct_rotate(const orc_ctx* c, int64_t* r0, int64_t* r1, const int64_t* c0,
                   const int64_t* c1, size_t num_q, uint32_t auto_idx, const int64_t* key0,
                   const int64_t* key1) {
  size_t   N  = c->N;
  int64_t* s0 = (int64_t*)malloc(sizeof(int64_t) * N * num_q);
  int64_t* s1 = (int64_t*)malloc(sizeof(int64_t) * N * num_q);
  int64_t* order = (int64_t*)malloc(sizeof(int64_t) * N);
  orc_key_switch(c, s0, s1, c1, num_q, key0, key1);
  orc_auto_order(c, auto_idx, order);
  for (size_t l = 0; l < num_q; l++) {
    orc_hw_modadd(c, s0 + l * N, s0 + l * N, c0 + l * N, l);
    orc_hw_rotate(c, r0 + l * N, s0 + l * N, order, l);
    orc_hw_rotate(c, r1 + l * N, s1 + l * N, order, l);
  }
  free(s0); free(s1); free(order);
}

/* tensor product (emitted per-limb loop, e.g. resnet20 .inc:5440-5450) + Relinearize() */
void orc_ct_mul_relin(const orc_ctx* c, int64_t* r0, int64_t* r1, const int64_t* a0,
                      const int64_t* a1, const int64_t* b0, const int64_t* b1, size_t num_q,
                      const int64_t* key0, const int64_t* key1) {
  size_t   N  = c->N;
  int64_t* d2 = (int64_t*)malloc(sizeof(int64_t) * N * num_q);
  int64_t* s0 = (int64_t*)malloc(sizeof(int64_t) * N * num_q);
  int64_t* s1 = (int64_t*)malloc(sizeof(int64_t) * N * num_q);
  int64_t* t  = (int64_t*)malloc(sizeof(int64_t) * N);
  for (size_t l = 0; l < num_q; l++) {
    orc_hw_modmul(c, d2 + l * N, a1 + l * N, b1 + l * N, l);
  }
  orc_key_switch(c, s0, s1, d2, num_q, key0, key1);
  for (size_t l = 0; l < num_q; l++) {
    orc_hw_modmul(c, r0 + l * N, a0 + l * N, b0 + l * N, l);
    orc_hw_modadd(c, r0 + l * N, r0 + l * N, s0 + l * N, l);
    orc_hw_modmul(c, r1 + l * N, a0 + l * N, b1 + l * N, l);
    orc_hw_modmul(c, t, a1 + l * N, b0 + l * N, l);
    orc_hw_modadd(c, r1 + l * N, r1 + l * N, t, l);
    orc_hw_modadd(c, r1 + l * N, r1 + l * N, s1 + l * N, l);
  }
  free(d2); free(s0); free(s1); free(t);
}

/* Encode_impl (64-bit path), src/util/ckks_encoder.c:199-299, with Embedding_inv
 * (src/util/ntt.c:713-753) and Transform_values_to_rns (src/util/polynomial.c:362-392). */
void orc_encode(const orc_ctx* c, int64_t* out, const double* vals, size_t len, uint32_t level,
                uint32_t slots, uint32_t sf_degree) {
  size_t N = c->N, M = 2 * N;
  if (slots == 0) slots = (uint32_t)(N / 2);
  double complex* res = (double complex*)calloc(slots, sizeof(double complex));
  double complex* tmp = (double complex*)calloc(slots, sizeof(double complex));
  for (size_t i = 0; i < len; i++) res[i] = vals[i];
  uint32_t logn = (uint32_t)log2((double)slots);
  for (uint32_t logm = logn; logm > 0; logm--) {
    size_t idx_mod = (size_t)1 << (logm + 2), gap = M / idx_mod;
    size_t num1 = (size_t)1 << logm, num2 = num1 >> 1;
    for (size_t j = 0; j < slots; j += num1) {
      for (size_t i = 0; i < num2; i++) {
        size_t         ridx = (idx_mod - (c->rot_group[i] % idx_mod)) * gap;
        double complex p = res[j + i] + res[j + i + num2];
        double complex m = res[j + i] - res[j + i + num2];
        m *= c->fft_rou[ridx];
        res[j + i]        = p;
        res[j + i + num2] = m;
      }
    }
  }
  for (size_t i = 0; i < slots; i++) tmp[bitrev((uint32_t)i, logn)] = res[i];
  for (size_t i = 0; i < slots; i++) tmp[i] /= slots;
  /* Delta = 1UL << scaling_mod_size (src/util/ckks_parameters.c:66) */
  double   delta   = (double)((u64)1 << c->sf_bits);
  int64_t* message = (int64_t*)calloc(N, sizeof(int64_t));
  size_t   gap     = N / (2 * (size_t)slots);
  int64_t  big     = (int64_t)(((u64)1 << 63) - ((u64)1 << 9)) - 1; /* Max_64bit_value */
  for (size_t i = 0; i < slots; i++) {
    int64_t re = llround(creal(tmp[i]) * delta + 0.5);
    int64_t im = llround(cimag(tmp[i]) * delta + 0.5);
    message[i * gap]           = re < 0 ? big + re : re;
    message[(i + slots) * gap] = im < 0 ? big + im : im;
  }
  int64_t half = big >> 1;
  for (size_t l = 0; l < level; l++) {
    int64_t q = (int64_t)c->mod[l], diff = big - q;
    u64     powp = (u64)delta % (u64)q;
    for (uint32_t d = 2; d < sf_degree; d++) powp = mulmod(powp, (u64)delta % (u64)q, (u64)q);
    for (size_t n = 0; n < N; n++) {
      int64_t v = message[n] > half ? message[n] - diff : message[n];
      v %= q;
      if (v < 0) v += q;
      if (sf_degree > 1) v = (int64_t)mulmod((u64)v, powp, (u64)q);
      out[l * N + n] = v;
    }
    orc_ntt(c, l, out + l * N);
  }
  free(res); free(tmp); free(message);
}
