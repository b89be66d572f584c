/* oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Flat, ctypes-friendly entry points around the UNMODIFIED reference rtlib
 * (compiled from /root/reference by oracle/Makefile into oracle/_ref/libace_ref.so).
 * Every function here only marshals raw int64 limb buffers into the reference's
 * POLYNOMIAL / CIPHERTEXT structs and calls the reference's own API:
 *   Prepare_context            fhe-cmplr/rtlib/ant/src/rtlib/context.c:29
 *   Hw_modadd/modmul/rotate    fhe-cmplr/rtlib/ant/src/poly/poly_arith.c:14-56
 *   Decomp_modup/Mod_down/Rescale  fhe-cmplr/rtlib/ant/src/poly/poly_eval.c:28-49
 *   Ftt_fwd / Ftt_inv          fhe-cmplr/rtlib/ant/src/util/ntt.c:163-187
 *   Rotate_ciph/Mul_ciph/...   fhe-cmplr/rtlib/ant/src/ckks/cipher_eval.c:292-364
 *
 * Determinism: keygen/encrypt draw from (i) the BLAKE2 PRNG global `Prng`
 * (prng.c:13,56-70) which we pre-seed, and (ii) libc rand() re-seeded with
 * srand(time) before every triangle sample (random_sample.c:20-24) which we
 * neutralise by defining a no-op srand() in this shared object (linked with
 * -Bsymbolic) and seeding once through srandom().
 */
#define _GNU_SOURCE
#include <complex.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common/rt_api.h"
#include "rt_ant/rt_ant.h"
#include "util/ckks_bootstrap_context.h"
#include "util/ckks_decryptor.h"
#include "util/ckks_encoder.h"
#include "util/ckks_encryptor.h"
#include "util/ckks_evaluator.h"
#include "util/ckks_key_generator.h"
#include "util/ckks_parameters.h"
#include "util/crt.h"
#include "util/ntt.h"
#include "util/prng.h"
#include "util/random_sample.h"

/* ---- pinned randomness ------------------------------------------------- */
/* mode 0 (default, what every golden vector was made with): srand() is swallowed, rand() keeps
 * running from srandom(12345) -- its position then depends on every rand() call of the library,
 * Is_prime's 200 trials per FMT_ASSERT included (number_theory.c:160-185).
 * mode 1 (exact key generation tests): the k-th Sample_triangle of the run draws from
 * srandom(Tri_base + k), i.e. a function of its own call number only.  Sample_triangle reaches
 * srand through Srand_time() = gettimeofday() + srand() (random_sample.c:20-24), Is_prime
 * through time() + srand(): the two wrappers below tell the callers apart.  Both return the
 * real time (clock_gettime); mode 0 behaves exactly as before. */
#include <time.h>
#include <sys/time.h>
static int      Pin_mode = 0, Srand_after_tod = 0;
static uint32_t Tri_base = 0, Tri_count = 0;
__attribute__((visibility("hidden"))) int gettimeofday(struct timeval* tv, void* tz) {
  struct timespec ts;
  (void)tz;
  clock_gettime(CLOCK_REALTIME, &ts);
  if (tv) { tv->tv_sec = ts.tv_sec; tv->tv_usec = ts.tv_nsec / 1000; }
  Srand_after_tod = 1;
  return 0;
}
__attribute__((visibility("hidden"))) time_t time(time_t* t) {
  struct timespec ts;
  clock_gettime(CLOCK_REALTIME, &ts);
  if (t) *t = ts.tv_sec;
  Srand_after_tod = 0;
  return ts.tv_sec;
}
/* rand() is random() in glibc; counting the calls gives the position of the (mode 0) stream.
 * Tri_pos[k] = how many values had been drawn since srandom(12345) when the k-th
 * Sample_triangle of the run started: with these positions the B200 runtime's exact key
 * generation (csrc/refrng.h) consumes the same stream without the reference having to run
 * (tests/golden/make_tri_positions.py stores them next to a model's golden values). */
#define TRI_LOG_MAX 8192
static uint64_t Rand_count = 0, Tri_pos[TRI_LOG_MAX];
static uint32_t Tri_logged = 0;
__attribute__((visibility("hidden"))) int rand(void) {
  Rand_count++;
  return (int)random();
}
void srand(unsigned seed) { /* swallow srand(time) */
  (void)seed;
  if (Srand_after_tod && Tri_logged < TRI_LOG_MAX) Tri_pos[Tri_logged++] = Rand_count;
  if (Pin_mode == 1 && Srand_after_tod) srandom(Tri_base + Tri_count++);
  Srand_after_tod = 0;
}
void ref_pin_mode(int mode, uint32_t tri_base) { Pin_mode = mode; Tri_base = tri_base; Tri_count = 0; }
uint32_t ref_triangle_calls(void) { return Tri_count; }
uint32_t ref_triangle_positions(uint64_t* out, uint32_t cap) {
  for (uint32_t i = 0; i < Tri_logged && i < cap; i++) out[i] = Tri_pos[i];
  return Tri_logged;
}

extern BLAKE2_PRNG* Prng;
BLAKE2_PRNG*        Alloc_blake2_prng();

static void Pin_random(void) {
  Tri_count = 0;
  Tri_logged = 0;
  Rand_count = 0;
  srandom(12345u);
  if (Prng == NULL) Prng = Alloc_blake2_prng();
  for (uint32_t i = 0; i < SEED_CNT; i++) {
    Set_ui32_value(Prng->_seed, i, 0x9e3779b9u * (i + 1));
  }
  Prng->_counter    = 0;
  Prng->_buffer_idx = 0;
}

/* ---- callbacks the runtime expects from the emitted translation unit ----
 * Stand-alone (primitive tests) they answer from Harness_params; when an emitted model unit
 * (oracle/_ref/<model>_ref.so, built from the reference's .onnx.inc with these entry points
 * renamed) registers itself through ref_set_callbacks they forward to it. */
typedef struct {
  CKKS_PARAMS* (*get_params)(void);
  RT_DATA_INFO* (*get_data_info)(void);
  bool (*main_graph)(void);
  DATA_SCHEME* (*enc_scheme)(int);
  DATA_SCHEME* (*dec_scheme)(int);
} REF_CALLBACKS;
static REF_CALLBACKS Cb;
static CKKS_PARAMS*  Harness_params = NULL;
static RT_DATA_INFO  Data_info_override;
static char          Data_file_override[4096];

void ref_set_callbacks(void* get_params, void* get_data_info, void* main_graph, void* enc_scheme,
                       void* dec_scheme) {
  Cb.get_params    = (CKKS_PARAMS * (*)(void)) get_params;
  Cb.get_data_info = (RT_DATA_INFO * (*)(void)) get_data_info;
  Cb.main_graph    = (bool (*)(void))main_graph;
  Cb.enc_scheme    = (DATA_SCHEME * (*)(int)) enc_scheme;
  Cb.dec_scheme    = (DATA_SCHEME * (*)(int)) dec_scheme;
}
CKKS_PARAMS* Get_context_params() { return Cb.get_params ? Cb.get_params() : Harness_params; }
RT_DATA_INFO* Get_rt_data_info() {
  if (!Cb.get_data_info) return NULL;
  RT_DATA_INFO* info = Cb.get_data_info();
  if (info == NULL || Data_file_override[0] == 0) return info;
  Data_info_override            = *info; /* emitted paths are absolute (/app/release/...) */
  Data_info_override._file_name = Data_file_override;
  return &Data_info_override;
}
int          Get_input_count() { return Cb.main_graph ? 1 : 0; }
int          Get_output_count() { return Cb.main_graph ? 1 : 0; }
DATA_SCHEME* Get_encode_scheme(int idx) { return Cb.enc_scheme ? Cb.enc_scheme(idx) : NULL; }
DATA_SCHEME* Get_decode_scheme(int idx) { return Cb.dec_scheme ? Cb.dec_scheme(idx) : NULL; }
bool         Main_graph() { return Cb.main_graph ? Cb.main_graph() : true; }

/* whole-model runs: Prepare_context with the emitted unit's own parameters, rotation
 * indices and (overridden) weight file; randomness pinned first */
int ref_init_emitted(const char* data_file) {
  if (Context != NULL) return -1;
  if (!Cb.get_params || !Cb.main_graph) return -2;
  Data_file_override[0] = 0;
  if (data_file) strncpy(Data_file_override, data_file, sizeof(Data_file_override) - 1);
  Pin_random();
  Prepare_context();
  return 0;
}

/* ---- context ------------------------------------------------------------ */
int ref_init(uint32_t degree, size_t mul_depth, size_t first_mod_size,
             size_t scaling_mod_size, size_t num_q_parts, size_t hamming_weight,
             const int32_t* rot_idxs, size_t num_rot_idx, int with_bootstrap) {
  if (Context != NULL) return -1;
  Pin_random();
  size_t sz      = sizeof(CKKS_PARAMS) + sizeof(int32_t) * (num_rot_idx + 1);
  Harness_params = (CKKS_PARAMS*)calloc(1, sz);
  Harness_params->_provider         = LIB_ANT;
  Harness_params->_poly_degree      = degree;
  Harness_params->_sec_level        = 0;
  Harness_params->_mul_depth        = mul_depth;
  Harness_params->_first_mod_size   = first_mod_size;
  Harness_params->_scaling_mod_size = scaling_mod_size;
  Harness_params->_num_q_parts      = num_q_parts;
  Harness_params->_hamming_weight   = hamming_weight;
  Harness_params->_num_rot_idx      = num_rot_idx;
  for (size_t i = 0; i < num_rot_idx; i++) {
    Harness_params->_rot_idxs[i] = rot_idxs[i];
  }
  if (with_bootstrap) {
    Prepare_context(); /* includes Bootstrap_precom(N/2) when the depth allows it */
    return 0;
  }
  /* Same object graph as Prepare_context (context.c:29-86) minus Bootstrap_precom and the
   * weight file, for primitive-level runs where bootstrap keys would cost minutes. */
  Io_init();
  CKKS_PARAMETER* params = Alloc_ckks_parameter();
  Set_num_q_parts(params, num_q_parts);
  Init_ckks_parameters_with_prime_size(params, degree, Get_sec_level(0), mul_depth + 1,
                                       first_mod_size, scaling_mod_size, hamming_weight);
  CKKS_CONTEXT*       ctxt   = (CKKS_CONTEXT*)malloc(sizeof(CKKS_CONTEXT));
  CKKS_KEY_GENERATOR* keygen =
      Alloc_ckks_key_generator(params, Harness_params->_rot_idxs, num_rot_idx);
  CKKS_ENCODER*   encoder = Alloc_ckks_encoder(params);
  CKKS_ENCRYPTOR* encryptor =
      Alloc_ckks_encryptor(params, keygen->_public_key, keygen->_secret_key);
  CKKS_DECRYPTOR* decryptor = Alloc_ckks_decryptor(params, keygen->_secret_key);
  ctxt->_params        = (PTR_TY)params;
  ctxt->_key_generator = (PTR_TY)keygen;
  ctxt->_encoder       = (PTR_TY)encoder;
  ctxt->_encryptor     = (PTR_TY)encryptor;
  ctxt->_decryptor     = (PTR_TY)decryptor;
  ctxt->_evaluator = (PTR_TY)Alloc_ckks_evaluator(params, encoder, decryptor, keygen);
  Context          = ctxt;
  return 0;
}

void ref_fini(void) {
  if (Context) Finalize_context();
  free(Harness_params);
  Harness_params = NULL;
}

static CRT_CONTEXT* Crt(void) { return Get_crt_context(); }
uint32_t ref_degree(void) { return Degree(); }
size_t   ref_num_q(void) { return Get_primes_cnt(Get_q(Crt())); }
size_t   ref_num_p(void) { return Get_primes_cnt(Get_p(Crt())); }
size_t   ref_num_parts(void) { return Get_num_parts(Get_qpart(Crt())); }
size_t   ref_part_size(void) { return Get_per_part_size(Get_qpart(Crt())); }
double   ref_default_scale(void) { return Get_default_sc(); }

void ref_get_primes(int64_t* q, int64_t* p) {
  MODULUS* m = Q_modulus();
  for (size_t i = 0; i < ref_num_q(); i++) q[i] = Get_mod_val(m + i);
  m = P_modulus();
  for (size_t i = 0; i < ref_num_p(); i++) p[i] = Get_mod_val(m + i);
}

/* is_p: 0 -> i-th Q prime, 1 -> i-th P prime */
static CRT_PRIME* Prime_at(int is_p, size_t idx) {
  return Get_prime_at(is_p ? Get_p(Crt()) : Get_q(Crt()), idx);
}

/* root of unity actually used by the NTT tables: rou[bitrev(1)] = rou[N/2] */
int64_t ref_psi(int is_p, size_t idx) {
  NTT_CONTEXT* ntt = Get_ntt(Prime_at(is_p, idx));
  return Get_i64_value_at(ntt->_rou, ntt->_degree >> 1);
}

void ref_ntt(int is_p, size_t idx, int64_t* data) {
  VALUE_LIST vl;
  Init_i64_value_list_no_copy(&vl, Degree(), data);
  Ftt_fwd(&vl, Get_ntt(Prime_at(is_p, idx)), &vl);
}

void ref_intt(int is_p, size_t idx, int64_t* data) {
  VALUE_LIST vl;
  Init_i64_value_list_no_copy(&vl, Degree(), data);
  Ftt_inv(&vl, Get_ntt(Prime_at(is_p, idx)), &vl);
}

/* ---- per-limb "hardware" ops (boundary a4) -------------------------------- */
static MODULUS* Mod_at(int is_p, size_t idx) {
  return (is_p ? P_modulus() : Q_modulus()) + idx;
}
void ref_hw_modadd(int64_t* r, int64_t* a, int64_t* b, int is_p, size_t idx) {
  Hw_modadd(r, a, b, Mod_at(is_p, idx), Degree());
}
void ref_hw_modmul(int64_t* r, int64_t* a, int64_t* b, int is_p, size_t idx) {
  Hw_modmul(r, a, b, Mod_at(is_p, idx), Degree());
}
void ref_hw_rotate(int64_t* r, int64_t* a, int64_t* order, int is_p, size_t idx) {
  Hw_rotate(r, a, order, Mod_at(is_p, idx), Degree());
}

int ref_auto_order(int32_t rot_idx, int64_t* out) {
  memcpy(out, Auto_order(rot_idx), sizeof(int64_t) * Degree());
  return (int)Auto_idx(rot_idx);
}

/* ---- polynomial-level ops (a5, a6, a8) ------------------------------------- */
static void Wrap_poly(POLYNOMIAL* p, int64_t* data, size_t num_q, size_t num_p,
                      bool is_ntt) {
  memset(p, 0, sizeof(*p));
  Init_poly_data(p, Degree(), num_q, num_p, data);
  Set_is_ntt(p, is_ntt);
}

/* in: num_q limbs (NTT form). out: (num_q + K) limbs. */
void ref_decomp_modup(int64_t* out, int64_t* in, size_t num_q, uint32_t part) {
  POLYNOMIAL res, poly;
  Wrap_poly(&poly, in, num_q, 0, true);
  Wrap_poly(&res, out, num_q, ref_num_p(), true);
  Decomp_modup(&res, &poly, part);
}

/* unfused pair Decomp + Mod_up (poly_eval.c:11-26) */
void ref_decomp_then_modup(int64_t* out, int64_t* in, size_t num_q, uint32_t part) {
  POLYNOMIAL res, poly, dec;
  Wrap_poly(&poly, in, num_q, 0, true);
  Wrap_poly(&res, out, num_q, ref_num_p(), true);
  memset(&dec, 0, sizeof(dec));
  Alloc_poly_data(&dec, Degree(), ref_part_size(), 0);
  Decomp(&dec, &poly, part);
  Mod_up(&res, &dec, part);
  Free_poly_data(&dec);
}

/* in: (num_q + K) limbs NTT form. out: num_q limbs. NOTE: clobbers the P part of `in`
 * exactly like the reference (in-place INTT, polynomial.c:941-945). */
void ref_mod_down(int64_t* out, int64_t* in, size_t num_q) {
  POLYNOMIAL res, poly;
  Wrap_poly(&poly, in, num_q, ref_num_p(), true);
  Wrap_poly(&res, out, num_q, 0, true);
  Mod_down(&res, &poly);
}

/* in: num_q limbs NTT form. out: buffer of num_q limbs, first num_q-1 are valid. */
void ref_rescale(int64_t* out, int64_t* in, size_t num_q) {
  POLYNOMIAL res, poly;
  Wrap_poly(&poly, in, num_q, 0, true);
  Wrap_poly(&res, out, num_q, 0, true);
  Rescale(&res, &poly);
}

/* ---- keys ------------------------------------------------------------------- */
/* copies one key polynomial: (num_q_total + K) limbs, Q limbs first then P limbs */
int ref_swk_export(int is_rot, int32_t rot_idx, uint32_t part, int which,
                   int64_t* out) {
  SW_KEY swk = Swk(is_rot, rot_idx);
  if (part >= Get_swk_size(swk)) return -1;
  POLY p = which ? Pk1_at(swk, part) : Pk0_at(swk, part);
  memcpy(out, Get_poly_coeffs(p), Get_poly_mem_size(p));
  return (int)Get_num_alloc_primes(p);
}

/* same, looked up by automorphism index (the conjugation key 2N-1 has no rotation index);
 * returns -2 when the key does not exist */
int ref_swk_export_auto(uint32_t auto_idx, uint32_t part, int which, int64_t* out) {
  CKKS_KEY_GENERATOR* kg  = (CKKS_KEY_GENERATOR*)Get_key_gen(Context);
  SWITCH_KEY*         swk = Get_auto_key(kg, auto_idx);
  if (swk == NULL) return -2;
  if (part >= Get_swk_size(swk)) return -1;
  POLY p = which ? Pk1_at(swk, part) : Pk0_at(swk, part);
  memcpy(out, Get_poly_coeffs(p), Get_poly_mem_size(p));
  return (int)Get_num_alloc_primes(p);
}

/* secret key in NTT form over Q then P */
void ref_sk_export(int64_t* out) {
  CKKS_KEY_GENERATOR* kg = (CKKS_KEY_GENERATOR*)Get_key_gen(Context);
  POLYNOMIAL*         sk = Get_ntt_sk(kg->_secret_key);
  memcpy(out, Get_poly_coeffs(sk), Get_poly_mem_size(sk));
}

void ref_pk_export(int64_t* pk0, int64_t* pk1) {
  CKKS_KEY_GENERATOR* kg = (CKKS_KEY_GENERATOR*)Get_key_gen(Context);
  memcpy(pk0, Get_poly_coeffs(Get_pk0(kg->_public_key)),
         Get_poly_mem_size(Get_pk0(kg->_public_key)));
  memcpy(pk1, Get_poly_coeffs(Get_pk1(kg->_public_key)),
         Get_poly_mem_size(Get_pk1(kg->_public_key)));
}

/* ---- ciphertext-level ops ----------------------------------------------------- */
typedef struct {
  int64_t* c0;
  int64_t* c1;
  uint32_t level;
  uint32_t slots;
  uint32_t sf_degree;
  double   scale;
} REF_CT;

static void Wrap_ct(CIPHERTEXT* ct, const REF_CT* r) {
  memset(ct, 0, sizeof(*ct));
  Wrap_poly(Get_c0(ct), r->c0, r->level, 0, true);
  Wrap_poly(Get_c1(ct), r->c1, r->level, 0, true);
  ct->_slots          = r->slots;
  ct->_scaling_factor = r->scale;
  ct->_sf_degree      = r->sf_degree;
}

static void Unwrap_ct(REF_CT* r, CIPHERTEXT* ct) {
  size_t n = sizeof(int64_t) * Degree() * Get_ciph_level(ct);
  memcpy(r->c0, Get_poly_coeffs(Get_c0(ct)), n);
  memcpy(r->c1, Get_poly_coeffs(Get_c1(ct)), n);
  r->level     = Get_ciph_level(ct);
  r->slots     = Get_ciph_slots(ct);
  r->sf_degree = Get_ciph_sf_degree(ct);
  r->scale     = Get_ciph_sfactor(ct);
}

/* encode `len` real values into `slots` slots at `level`, sf_degree; out = level limbs */
void ref_encode(int64_t* out, const double* vals, size_t len, uint32_t level,
                uint32_t slots, uint32_t sf_degree, double* scale_out) {
  VALUE_LIST* vec = Alloc_value_list(DCMPLX_TYPE, len);
  for (size_t i = 0; i < len; i++) DCMPLX_VALUE_AT(vec, i) = vals[i];
  PLAINTEXT* plain = Alloc_plaintext();
  Encode_at_level_with_sf(plain, (CKKS_ENCODER*)Context->_encoder, vec, level,
                          slots, sf_degree);
  memcpy(out, Get_poly_coeffs(Get_plain_poly(plain)),
         sizeof(int64_t) * Degree() * level);
  if (scale_out) *scale_out = Get_plain_scaling_factor(plain);
  Free_plaintext(plain);
  Free_value_list(vec);
}

/* the run-time weight path: Encode_plain_from_float (plain_eval.c:25-42) */
void ref_encode_float(int64_t* out, float* vals, size_t len, uint32_t sc_degree,
                      uint32_t level, double* scale_out, uint32_t* slots_out) {
  PLAINTEXT* plain = Alloc_plaintext();
  Encode_plain_from_float(plain, vals, len, sc_degree, level);
  memcpy(out, Get_poly_coeffs(Get_plain_poly(plain)),
         sizeof(int64_t) * Degree() * Get_poly_level(Get_plain_poly(plain)));
  if (scale_out) *scale_out = Get_plain_scaling_factor(plain);
  if (slots_out) *slots_out = Get_plain_slots(plain);
  Free_plaintext(plain);
}

void ref_encode_double(int64_t* out, double* vals, size_t len, uint32_t sc_degree,
                       uint32_t level, double* scale_out, uint32_t* slots_out) {
  PLAINTEXT* plain = Alloc_plaintext();
  Encode_plain_from_double(plain, vals, len, sc_degree, level);
  memcpy(out, Get_poly_coeffs(Get_plain_poly(plain)),
         sizeof(int64_t) * Degree() * Get_poly_level(Get_plain_poly(plain)));
  if (scale_out) *scale_out = Get_plain_scaling_factor(plain);
  if (slots_out) *slots_out = Get_plain_slots(plain);
  Free_plaintext(plain);
}

/* encode + pk-encrypt; r->c0/c1 must hold `level` limbs */
void ref_encrypt(REF_CT* r, const double* vals, size_t len, uint32_t level,
                 uint32_t slots) {
  VALUE_LIST* vec = Alloc_value_list(DCMPLX_TYPE, len);
  for (size_t i = 0; i < len; i++) DCMPLX_VALUE_AT(vec, i) = vals[i];
  PLAINTEXT* plain = Alloc_plaintext();
  Encode_at_level_with_sf(plain, (CKKS_ENCODER*)Context->_encoder, vec, level,
                          slots, 1);
  CIPHERTEXT* ct = Alloc_ciphertext();
  Encrypt_msg(ct, (CKKS_ENCRYPTOR*)Context->_encryptor, plain);
  Unwrap_ct(r, ct);
  Free_ciphertext(ct);
  Free_plaintext(plain);
  Free_value_list(vec);
}

/* decrypt + decode; out must hold r->slots doubles (real parts) */
void ref_decrypt(double* out, const REF_CT* r) {
  CIPHERTEXT ct;
  Wrap_ct(&ct, r);
  double* msg = Get_msg(&ct);
  memcpy(out, msg, sizeof(double) * r->slots);
  free(msg);
}

void ref_ct_add(REF_CT* res, const REF_CT* a, const REF_CT* b) {
  CIPHERTEXT  x, y;
  CIPHERTEXT* z = Alloc_ciphertext();
  Wrap_ct(&x, a);
  Wrap_ct(&y, b);
  Add_ciph(z, &x, &y);
  Unwrap_ct(res, z);
  Free_ciphertext(z);
}

/* tensor product + relinearisation (no rescale) */
void ref_ct_mul(REF_CT* res, const REF_CT* a, const REF_CT* b) {
  CIPHERTEXT  x, y;
  CIPHERTEXT* z = Alloc_ciphertext();
  Wrap_ct(&x, a);
  Wrap_ct(&y, b);
  Mul_ciph(z, &x, &y);
  Unwrap_ct(res, z);
  Free_ciphertext(z);
}

void ref_ct_rescale(REF_CT* res, const REF_CT* a) {
  CIPHERTEXT  x;
  CIPHERTEXT* z = Alloc_ciphertext();
  Wrap_ct(&x, a);
  Rescale_ciph(z, &x);
  Unwrap_ct(res, z);
  Free_ciphertext(z);
}

void ref_ct_rotate(REF_CT* res, const REF_CT* a, int32_t rot) {
  CIPHERTEXT  x;
  CIPHERTEXT* z = Alloc_ciphertext();
  Wrap_ct(&x, a);
  Rotate_ciph(z, &x, rot);
  Unwrap_ct(res, z);
  Free_ciphertext(z);
}

/* plaintext (already encoded, `level` limbs, NTT form) times ciphertext */
void ref_ct_mul_plain(REF_CT* res, const REF_CT* a, int64_t* pt, double pt_scale,
                      uint32_t pt_sf_degree) {
  CIPHERTEXT  x;
  CIPHERTEXT* z = Alloc_ciphertext();
  PLAINTEXT   p;
  memset(&p, 0, sizeof(p));
  Wrap_ct(&x, a);
  Wrap_poly(Get_plain_poly(&p), pt, a->level, 0, true);
  p._slots          = a->slots;
  p._scaling_factor = pt_scale;
  p._sf_degree      = pt_sf_degree;
  Mul_plain(z, &x, &p);
  Unwrap_ct(res, z);
  Free_ciphertext(z);
}

/* ---- whole-model runs (the emitted Main_graph through the reference's own driver API,
 *      ant/src/rtlib/rtlib.c:41-87, common/src/rt_lib.c:16-20) -------------------------- */
void* Io_get_input(const char* name, size_t idx);
void* Io_get_output(const char* name, size_t idx);

void ref_prepare_input(const double* vals, size_t n, size_t c, size_t h, size_t w,
                       const char* name) {
  TENSOR* t = Alloc_tensor(n, c, h, w, vals);
  Prepare_input(t, name);
  Free_tensor(t);
}
static int Copy_io(REF_CT* out, CIPHERTEXT* ct) {
  if (ct == NULL) return -1;
  Unwrap_ct(out, ct);
  return (int)out->level;
}
/* copies (does not consume) the pending input / produced output ciphertext; buffers: L limbs */
int  ref_peek_input(const char* name, REF_CT* out) { return Copy_io(out, (CIPHERTEXT*)Io_get_input(name, 0)); }
int  ref_peek_output(const char* name, REF_CT* out) { return Copy_io(out, (CIPHERTEXT*)Io_get_output(name, 0)); }
void ref_run_main_graph(void) { Run_main_graph(); }
void ref_handle_output(const char* name, double* out, size_t n) {
  double* r = Handle_output(name);
  memcpy(out, r, n * sizeof(double));
  free(r);
}

/* bootstrap (a11); res buffers must hold level_after_bts(+) limbs: use num_q limbs */
void ref_ct_bootstrap(REF_CT* res, const REF_CT* a, uint32_t level_after_bts) {
  CIPHERTEXT  x;
  CIPHERTEXT* z = Alloc_ciphertext();
  Wrap_ct(&x, a);
  Bootstrap(z, &x, level_after_bts);
  Unwrap_ct(res, z);
  Free_ciphertext(z);
}

/* one diagonal plaintext of the bootstrap tables (u0hatt_pre_fft for encoding, u0_pre_fft for
 * decoding): copies num_q + num_p limbs, returns num_q (0: entry unused, -1: no tables) */
int ref_bts_plain(uint32_t slots, int encoding, uint32_t step, uint32_t idx, int64_t* out) {
  CKKS_BTS_CTX*    bts    = Get_bts_ctx((CKKS_EVALUATOR*)Get_eval(Context));
  CKKS_BTS_PRECOM* precom = Get_bts_precom(bts, slots);
  if (!precom) return -1;
  VL_VL_PLAIN* tab = encoding ? Get_u0hatt_pre_fft(precom) : Get_u0_pre_fft(precom);
  VALUE_LIST*  row = Get_vl_value_at(tab, step);
  if (idx >= LIST_LEN(row)) return 0;
  PLAINTEXT* pt = (PLAINTEXT*)Get_ptr_value_at(row, idx);
  if (pt == NULL) return 0;
  POLYNOMIAL* p = Get_plain_poly(pt);
  memcpy(out, Get_poly_coeffs(p), Get_poly_mem_size(p));
  return (int)Get_num_q(p);
}

/* Coeffs_to_slots / Slots_to_coeffs alone (ckks_bootstrap_context.c:1494-1504) */
void ref_bts_linear(REF_CT* res, const REF_CT* a, int encoding) {
  CKKS_BTS_CTX*    bts    = Get_bts_ctx((CKKS_EVALUATOR*)Get_eval(Context));
  CKKS_BTS_PRECOM* precom = Get_bts_precom(bts, a->slots);
  CIPHERTEXT       x;
  CIPHERTEXT*      z = Alloc_ciphertext();
  Wrap_ct(&x, a);
  if (encoding) Coeffs_to_slots(z, &x, Get_u0hatt_pre_fft(precom), bts);
  else Slots_to_coeffs(z, &x, Get_u0_pre_fft(precom), bts);
  Unwrap_ct(res, z);
  Free_ciphertext(z);
}

/* collapsed FFT diagonals of the bootstrap linear transforms (Coeff_collapse,
 * ckks_bootstrap_context.c:612-776) with the root table of Bootstrap_setup (:1107-1121);
 * out: [level][row][slots] complex doubles (re, im), rows as the reference sizes them;
 * returns the number of complex values written */
VALUE_LIST* Coeff_collapse(VALUE_LIST* ksipows, VALUE_LIST* rot_group, uint32_t level_budget,
                           bool flag, bool encoding);
size_t ref_coeff_collapse(uint32_t slots, uint32_t budget, int flag, int encoding, double* out) {
  uint32_t    slots_4   = 4 * slots;
  VALUE_LIST* rot_group = Alloc_value_list(UI32_TYPE, slots);
  uint32_t    five_pows = 1;
  FOR_ALL_ELEM(rot_group, idx) {
    UI32_VALUE_AT(rot_group, idx) = five_pows;
    five_pows *= 5;
    five_pows %= slots_4;
  }
  VALUE_LIST* ksi_pows = Alloc_value_list(DCMPLX_TYPE, slots_4 + 1);
  for (size_t idx = 0; idx < slots_4; idx++) {
    double angle                   = 2.0 * M_PI * idx / slots_4;
    DCMPLX_VALUE_AT(ksi_pows, idx) = cos(angle) + sin(angle) * I;
  }
  DCMPLX_VALUE_AT(ksi_pows, slots_4) = DCMPLX_VALUE_AT(ksi_pows, 0);
  VALUE_LIST* coeffs = Coeff_collapse(ksi_pows, rot_group, budget, flag, encoding);
  size_t      n      = 0;
  FOR_ALL_ELEM(coeffs, i) {
    VALUE_LIST* lvl = Get_vl_value_at(coeffs, i);
    FOR_ALL_ELEM(lvl, j) {
      VALUE_LIST* row = Get_vl_value_at(lvl, j);
      FOR_ALL_ELEM(row, k) {
        DCMPLX v     = Get_dcmplx_value_at(row, k);
        out[2 * n]   = creal(v);
        out[2 * n + 1] = cimag(v);
        n++;
      }
    }
  }
  Free_value_list(coeffs);
  Free_value_list(rot_group);
  Free_value_list(ksi_pows);
  return n;
}

/* ---- the raw random sources (tests of the runtime's restatement, csrc/refrng.h) ---------- */
void ref_prng_words(const uint32_t* seed16, uint64_t counter, uint32_t* out, size_t n) {
  BLAKE2_PRNG* saved = Prng;
  Prng = Alloc_blake2_prng();
  for (uint32_t i = 0; i < SEED_CNT; i++) Set_ui32_value(Prng->_seed, i, seed16[i]);
  Prng->_counter = counter;
  Prng->_buffer_idx = 0;
  for (size_t i = 0; i < n; i++) out[i] = Get_prng_value(Prng);
  Prng = saved;
}
void ref_sample_uniform(const uint32_t* seed16, uint64_t counter, int64_t* out, size_t n, uint64_t bound) {
  BLAKE2_PRNG* saved = Prng;
  Prng = Alloc_blake2_prng();
  for (uint32_t i = 0; i < SEED_CNT; i++) Set_ui32_value(Prng->_seed, i, seed16[i]);
  Prng->_counter = counter;
  Prng->_buffer_idx = 0;
  VALUE_LIST v;
  Init_i64_value_list_no_copy(&v, n, out);
  Sample_uniform(&v, bound);
  Prng = saved;
}
void ref_sample_ternary(const uint32_t* seed16, uint64_t counter, int64_t* out, size_t n, int64_t hw) {
  BLAKE2_PRNG* saved = Prng;
  Prng = Alloc_blake2_prng();
  for (uint32_t i = 0; i < SEED_CNT; i++) Set_ui32_value(Prng->_seed, i, seed16[i]);
  Prng->_counter = counter;
  Prng->_buffer_idx = 0;
  VALUE_LIST v;
  Init_i64_value_list_no_copy(&v, n, out);
  Sample_ternary(&v, hw);
  Prng = saved;
}
void ref_sample_triangle(uint32_t seed, int64_t* out, size_t n) { /* what mode 1 gives call k: seed = base + k */
  int saved = Pin_mode;
  Pin_mode  = 0;
  srandom(seed);
  VALUE_LIST v;
  Init_i64_value_list_no_copy(&v, n, out);
  Sample_triangle(&v);
  Pin_mode = saved;
}

/* ---- CPU baseline: the chain HMult+relin -> rescale -> rotate on independent ciphertexts,
 *      one per thread, sharing the (read-only) context like the reference's OpenMP image
 *      loop does (ant/dataset/resnet_cifar.main.inc:81).  Returns wall seconds. ---------- */
#include <pthread.h>
#include <time.h>

typedef struct {
  CIPHERTEXT* ct;
  int         iters;
  int32_t     rot;
} CHAIN_ARG;

static void* Chain_worker(void* p) {
  CHAIN_ARG* a = (CHAIN_ARG*)p;
  for (int i = 0; i < a->iters; i++) {
    CIPHERTEXT* m  = Alloc_ciphertext();
    CIPHERTEXT* rs = Alloc_ciphertext();
    CIPHERTEXT* r  = Alloc_ciphertext();
    Mul_ciph(m, a->ct, a->ct);
    Rescale_ciph(rs, m);
    Rotate_ciph(r, rs, a->rot);
    Free_ciphertext(m);
    Free_ciphertext(rs);
    Free_ciphertext(r);
  }
  return NULL;
}

double ref_bench_chain(int nthreads, int iters, uint32_t level, int32_t rot) {
  if (nthreads < 1) nthreads = 1;
  uint32_t    slots = Degree() / 2;
  CIPHERTEXT** cts  = (CIPHERTEXT**)malloc(sizeof(CIPHERTEXT*) * nthreads);
  VALUE_LIST* vec   = Alloc_value_list(DCMPLX_TYPE, slots);
  for (uint32_t i = 0; i < slots; i++) {
    DCMPLX_VALUE_AT(vec, i) = ((i * 2654435761u) % 1000) / 1000.0 - 0.5;
  }
  for (int t = 0; t < nthreads; t++) {
    PLAINTEXT* plain = Alloc_plaintext();
    Encode_at_level_with_sf(plain, (CKKS_ENCODER*)Context->_encoder, vec, level, slots, 1);
    cts[t] = Alloc_ciphertext();
    Encrypt_msg(cts[t], (CKKS_ENCRYPTOR*)Context->_encryptor, plain);
    Free_plaintext(plain);
  }
  pthread_t* th   = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  CHAIN_ARG* args = (CHAIN_ARG*)malloc(sizeof(CHAIN_ARG) * nthreads);
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < nthreads; t++) {
    args[t].ct = cts[t]; args[t].iters = iters; args[t].rot = rot;
    pthread_create(&th[t], NULL, Chain_worker, &args[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  for (int t = 0; t < nthreads; t++) Free_ciphertext(cts[t]);
  free(cts); free(th); free(args);
  Free_value_list(vec);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
