/* ace_b200.h -- C ABI of the B200-native CKKS evaluation runtime (libace_b200.so).
 *
 * Drop-in boundary for the polynomial-level "rt_ant" API that ACE-generated C calls
 * (reference: fhe-cmplr/rtlib/include/rt_ant/rt_ant.h:17-21 and the headers it pulls in).
 * Plain pointers and sizes only.  Every `int64_t*` limb pointer below is a DEVICE pointer
 * obtained from ace_alloc_limbs(); a "limb" is N int64 canonical residues, polynomials are
 * limb-major exactly like POLYNOMIAL._data (ant/include/util/polynomial.h:35-44): the
 * num_q Q-limbs first, then the K P-limbs.
 *
 * Modulus index `g`: 0..L-1 are the Q primes, L..L+K-1 the P primes -- the same order in
 * which the reference lays out its contiguous MODULUS arrays (Q_modulus()+i, P_modulus()+i,
 * ant/src/rtlib/context.c:156-160).
 *
 * All calls are asynchronous on the context's CUDA stream except ace_download/ace_sync.
 * Return value: 0 on success, negative on error (message via ace_last_error()); the
 * reference aborts on the same conditions (FMT_ASSERT, rtlib/include/common/error.h:23-29).
 */
#ifndef ACE_B200_H
#define ACE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ace_ctx ace_ctx;

/* ---- context: replaces Prepare_context's parameter/CRT/NTT-table part
 *      (ant/src/rtlib/context.c:29-47, ant/src/util/ckks_parameters.c:60-97).
 *      Arguments are the CKKS_PARAMS fields (rtlib/include/common/common.h:71-82). */
int  ace_ctx_create(ace_ctx** out, uint32_t poly_degree, size_t mul_depth,
                    size_t first_mod_size, size_t scaling_mod_size, size_t num_q_parts,
                    size_t hamming_weight, int device);
void ace_ctx_destroy(ace_ctx* ctx);          /* Finalize_context, context.c:88-138 */
const char* ace_last_error(void);

uint32_t ace_degree(const ace_ctx* ctx);     /* Degree(),      context.c:140-142 */
size_t   ace_num_q(const ace_ctx* ctx);      /* Get_primes_cnt(Get_q(crt)) */
size_t   ace_num_p(const ace_ctx* ctx);      /* Get_p_cnt(),   context.c:154 */
size_t   ace_num_q_parts(const ace_ctx* ctx);/* Get_q_parts(), context.c:148-150 */
size_t   ace_part_size(const ace_ctx* ctx);  /* Get_per_part_size(qpart) */
int      ace_get_primes(const ace_ctx* ctx, int64_t* q, int64_t* p); /* Q_modulus()/P_modulus() values */
int64_t  ace_psi(const ace_ctx* ctx, uint32_t g);   /* 2N-th root of unity behind NTT_CONTEXT._rou */
size_t   ace_num_decomp(const ace_ctx* ctx, size_t num_q); /* Num_decomp(), poly_eval.h:108-113 */
uint64_t ace_launch_count(const ace_ctx* ctx);      /* kernels launched so far */

/* ---- memory: Alloc_poly/Free_poly_data (ant/include/poly/poly_eval.h:29-37,
 *      ant/include/util/polynomial.h:54-77); zero != 0 reproduces the memset. */
int64_t* ace_alloc_limbs(ace_ctx* ctx, size_t n_limbs, int zero);
int      ace_free_limbs(ace_ctx* ctx, int64_t* dev);
int      ace_upload(ace_ctx* ctx, int64_t* dev_dst, const int64_t* host_src, size_t n_limbs);
int      ace_download(ace_ctx* ctx, int64_t* host_dst, const int64_t* dev_src, size_t n_limbs);
int      ace_copy_limbs(ace_ctx* ctx, int64_t* dev_dst, const int64_t* dev_src, size_t n_limbs); /* Set_coeffs, poly_eval.h:74-79 */
int      ace_zero_limbs(ace_ctx* ctx, int64_t* dev, size_t n_limbs);
int      ace_sync(ace_ctx* ctx);

/* ---- per-limb "hardware" ops (ant/src/poly/poly_arith.c:14-56).  n_limbs consecutive
 *      limbs are processed in one launch, limb i with modulus g0+i (n_limbs = 1 is the
 *      reference call). */
int ace_hw_modadd(ace_ctx* ctx, int64_t* res, const int64_t* a, const int64_t* b, uint32_t g0, uint32_t n_limbs);
int ace_hw_modsub(ace_ctx* ctx, int64_t* res, const int64_t* a, const int64_t* b, uint32_t g0, uint32_t n_limbs);
int ace_hw_modmul(ace_ctx* ctx, int64_t* res, const int64_t* a, const int64_t* b, uint32_t g0, uint32_t n_limbs);
int ace_hw_rotate(ace_ctx* ctx, int64_t* res, const int64_t* a, const int64_t* order_dev, uint32_t g0, uint32_t n_limbs);

/* ---- negacyclic NTT / INTT in place (Ftt_fwd / Ftt_inv, ant/src/util/ntt.c:163-187) */
int ace_ntt(ace_ctx* ctx, int64_t* data, uint32_t g0, uint32_t n_limbs);
int ace_intt(ace_ctx* ctx, int64_t* data, uint32_t g0, uint32_t n_limbs);

/* ---- polynomial-level ops (ant/src/poly/poly_eval.c:28-49).
 *      decomp_modup: in = num_q limbs, out = num_q + K limbs.
 *      mod_down:     in = num_q + K limbs, out = num_q limbs.
 *      rescale:      in = num_q limbs, out = num_q - 1 limbs.            NTT form throughout. */
int ace_decomp_modup(ace_ctx* ctx, int64_t* out, const int64_t* in, uint32_t num_q, uint32_t q_part_idx);
int ace_mod_down(ace_ctx* ctx, int64_t* out, const int64_t* in, uint32_t num_q);
int ace_rescale(ace_ctx* ctx, int64_t* out, const int64_t* in, uint32_t num_q);

/* ---- keys and automorphism tables (ant/include/rtlib/key_gen.h:28-75).
 *      A switch key is num_q_parts public keys; each polynomial has L+K limbs.
 *      ace_swk_import copies one host polynomial (Pk0_at: which=0, Pk1_at: which=1). */
uint32_t ace_auto_index(const ace_ctx* ctx, int32_t rot_idx);          /* Auto_idx   */
const int64_t* ace_auto_order(ace_ctx* ctx, int32_t rot_idx);          /* Auto_order (device table) */
int ace_swk_import(ace_ctx* ctx, int is_rot, int32_t rot_idx, uint32_t part, int which, const int64_t* host_poly);
const int64_t* ace_swk_poly(ace_ctx* ctx, int is_rot, int32_t rot_idx, uint32_t part, int which); /* device pointer, Pk0_at/Pk1_at */

/* ---- fused ciphertext-level entry points: same results as the emitted Rotate() /
 *      Relinearize() bodies (dataset/resnet20_cifar10_pre.onnx.inc:6972-7146) with all
 *      limbs and digits batched.  All pointers: num_q limbs. */
int ace_key_switch(ace_ctx* ctx, int64_t* out0, int64_t* out1, const int64_t* d, uint32_t num_q, int is_rot, int32_t rot_idx);
int ace_ct_rotate(ace_ctx* ctx, int64_t* r0, int64_t* r1, const int64_t* c0, const int64_t* c1, uint32_t num_q, int32_t rot_idx);
int ace_ct_mul_relin(ace_ctx* ctx, int64_t* r0, int64_t* r1, const int64_t* a0, const int64_t* a1, const int64_t* b0, const int64_t* b1, uint32_t num_q);
/*      ace_ct_rotate_hoisted: n rotations of one ciphertext with ONE Decomp_modup (Switch_key_precompute
 *      reused by every rotation, ckks_bootstrap_context.c:1284-1299; the fusion measured in
 *      unittest/ut_ksw_opt.cxx:663-770); r0[i] / r1[i] are bit-identical to ace_ct_rotate(rot_idxs[i]).
 *      ace_ct_mul_plain_acc: acc (+)= ct (.) pt over num_q limbs of both polynomials (first != 0:
 *      acc = ct (.) pt) -- Mul_plain + Add_ciph of the emitted convolution loops in one launch.
 *      These two are what a POLY emitter retargeted to ciphertext granularity would call
 *      (lib_provider.h:18-21). */
int ace_ct_rotate_hoisted(ace_ctx* ctx, int64_t* const* r0, int64_t* const* r1, const int64_t* c0, const int64_t* c1,
                          uint32_t num_q, const int32_t* rot_idxs, size_t n);
int ace_ct_mul_plain_acc(ace_ctx* ctx, int64_t* acc0, int64_t* acc1, const int64_t* c0, const int64_t* c1,
                         const int64_t* pt, uint32_t num_q, int first);
int ace_ct_rescale(ace_ctx* ctx, int64_t* r0, int64_t* r1, const int64_t* c0, const int64_t* c1, uint32_t num_q);

/* ---- client side: keys, encryption, CKKS encode/decode.
 *      keygen:  seed == 0: entropy from the operating system (getrandom -> ChaCha20 streams,
 *               domain-separated per key / digit / encryption); seed != 0 pins a reproducible
 *               stream, TESTS ONLY.  ace_encrypt's last argument is a per-encryption id.
 *               Alloc_ckks_key_generator (ant/src/util/ckks_key_generator.c:13-37): secret,
 *               public, relinearisation key and one rotation key per index.  Sampling uses the
 *               runtime's own generator (valid keys, not bit-identical to the reference's).
 *      import:  take the reference's keys instead (parity runs).
 *      encode:  Encode_at_level_with_sf (ant/src/util/ckks_encoder.c:199-299), FP64 on the GPU,
 *               bit-exact; `vals` is a HOST array of len reals, out = level (+p_cnt) limbs.
 *      encode_value: Encode_val_at_level (ckks_encoder.c:464-528).
 *      encrypt / decrypt: ant/src/util/ckks_encryptor.c:20-95, ckks_decryptor.c:19-65.
 *      decode:  Decode (ckks_encoder.c:649-703); pt = level limbs NTT form (device),
 *               out_re/out_im HOST arrays of `slots` doubles (out_im may be NULL). */
int ace_keygen(ace_ctx* ctx, uint64_t seed, const int32_t* rot_idxs, size_t num_rot_idx);
int ace_sk_import(ace_ctx* ctx, const int64_t* host_sk_ntt_qp);          /* L+K limbs */
int ace_pk_import(ace_ctx* ctx, const int64_t* host_pk0, const int64_t* host_pk1); /* L limbs each */
int ace_encode(ace_ctx* ctx, int64_t* out, const double* vals, size_t len, uint32_t level, uint32_t slots, uint32_t sf_degree, uint32_t p_cnt);
int ace_encode_value(ace_ctx* ctx, int64_t* out, double value, uint32_t level, uint32_t sf_degree);
int ace_encrypt(ace_ctx* ctx, int64_t* c0, int64_t* c1, const int64_t* pt, uint32_t level, uint64_t seed);
int ace_decrypt(ace_ctx* ctx, int64_t* pt, const int64_t* c0, const int64_t* c1, uint32_t level);
int ace_decode(ace_ctx* ctx, double* out_re, double* out_im, const int64_t* pt, uint32_t level, uint32_t slots, double scale);

/* ---- bootstrap (a11): Bootstrap -> Eval_bootstrap (ant/src/ckks/cipher_eval.c:366-404,
 *      ant/src/util/ckks_bootstrap_context.c:1050-1192, 1237-1860, ckks_chebyshev.c).
 *      depth:        Get_bootstrap_depth with level budget {3,3} (ckks_bootstrap_context.h:275-283)
 *      setup:        Bootstrap_setup for `slots` (0 = N/2): C2S/S2C plaintext tables into HBM
 *      rot_indices:  Find_rot_indices: the rotation keys Bootstrap_keygen generates; returns the
 *                    count (may exceed cap).  The conjugation key is rotation index 2N-1 in
 *                    ace_swk_import / ace_keygen_rotations.
 *      bootstrap:    r0/r1 must hold L limbs; level/scale/sf_degree of the result are returned.
 *      Environment switches follow the reference: RTLIB_BTS_EVEN_POLY, RT_BTS_CLEAR_IMAG. */
int ace_bootstrap_depth(const ace_ctx* ctx);
int ace_bootstrap_setup(ace_ctx* ctx, uint32_t slots);
int ace_bootstrap_rot_indices(ace_ctx* ctx, uint32_t slots, int32_t* out, size_t cap);
/*      linear:       Coeffs_to_slots (encoding != 0) / Slots_to_coeffs alone (:1494-1504)
 *      plain:        device pointer to diagonal plaintext [step][idx] of the C2S (encoding) or S2C
 *                    table: *level Q limbs followed by K P limbs; NULL where the table has no entry */
int ace_bootstrap_linear(ace_ctx* ctx, int64_t* r0, int64_t* r1, uint32_t* out_level, double* out_scale,
                         uint32_t* out_sf_degree, const int64_t* c0, const int64_t* c1, uint32_t level,
                         uint32_t slots, double scale, uint32_t sf_degree, int encoding);
const int64_t* ace_bootstrap_plain(ace_ctx* ctx, uint32_t slots, int encoding, uint32_t step, uint32_t idx, uint32_t* level);
/*      fft_diagonals: host-only (no device): the collapsed FFT diagonals Coeff_collapse
 *                    (ckks_bootstrap_context.c:612-776) produces for `slots` and a level budget,
 *                    flattened [level][row][slot] as (re, im) doubles; returns the count of
 *                    complex values.  out needs budget * (2^(ceil(log2(slots)/budget)+1)) * slots
 *                    * 2 doubles at most. */
size_t ace_bootstrap_fft_diagonals(uint32_t slots, uint32_t level_budget, int flag, int encoding, double* out);
int ace_keygen_rotations(ace_ctx* ctx, uint64_t seed, const int32_t* rot_idxs, size_t num_rot_idx);
int ace_bootstrap(ace_ctx* ctx, int64_t* r0, int64_t* r1, uint32_t* out_level, double* out_scale,
                  uint32_t* out_sf_degree, const int64_t* c0, const int64_t* c1, uint32_t level,
                  uint32_t slots, double scale, uint32_t sf_degree, uint32_t level_after_bts);

/* ---- timing helpers for the benchmark: CUDA events on the context's stream */
int ace_timer_start(ace_ctx* ctx);
int ace_timer_stop_ms(ace_ctx* ctx, float* ms);

/* ---- exact key generation: the reference's generators (BLAKE2Xb PRNG seed words + counter,
 *      prng.h:30-60; the k-th Sample_triangle draws from glibc rand() after srandom(tri_base + k))
 *      consumed in the reference's order (Alloc_ckks_key_generator, ckks_key_generator.c:13-37):
 *      secret, public, relinearisation key, rotation keys of rot_idxs.  Keys and later
 *      ace_encrypt results are bit-identical to the reference's from the same seeds
 *      (tests/test_gpu_keygen.py).  ace_keygen_autos continues the current stream for more
 *      automorphism indices (Bootstrap_keygen: rotation keys, then 2N-1 = conjugation).
 *      export: the counterparts of ace_sk_import / ace_pk_import / ace_swk_import. */
int ace_keygen_reference(ace_ctx* ctx, const uint32_t* seed16, uint64_t counter, uint32_t tri_base,
                         const int32_t* rot_idxs, size_t num_rot_idx);
/*      _stream: the same, for a reference whose rand() is seeded ONCE (srandom(srandom_seed)) and
 *      never re-seeded: the k-th Sample_triangle starts tri_pos[k] draws into that stream (other
 *      callers of rand() -- Is_prime, number_theory.c:160-185 -- advance it in between).  This is
 *      how the whole-model golden runs pinned the reference (oracle/ref_harness.c, mode 0). */
int ace_keygen_reference_stream(ace_ctx* ctx, const uint32_t* seed16, uint64_t counter, uint32_t srandom_seed,
                                const uint64_t* tri_pos, size_t n_pos, const int32_t* rot_idxs,
                                size_t num_rot_idx);
int ace_keygen_autos(ace_ctx* ctx, const uint32_t* auto_idx, size_t n);
int ace_sk_export(ace_ctx* ctx, int64_t* host_sk_ntt_qp);
int ace_pk_export(ace_ctx* ctx, int64_t* host_pk0, int64_t* host_pk1);
int ace_swk_export(ace_ctx* ctx, int is_rot, uint32_t auto_idx, uint32_t part, int which, int64_t* host_poly);

/* ---- key and ciphertext files (the reference has no serialisation).  Little-endian,
 *      magic "ACEB200K" / "ACEB200C", version, the parameter set (N, L, K, dnum, every modulus: a
 *      file only loads into a context with the same primes), then the limbs as stored in HBM.
 *      with_secret = 0 writes the evaluation side only (public, relinearisation, rotation keys). */
int ace_keys_save(ace_ctx* ctx, const char* path, int with_secret);
int ace_keys_load(ace_ctx* ctx, const char* path);
int ace_ct_save(ace_ctx* ctx, const char* path, const int64_t* c0, const int64_t* c1, uint32_t level,
                uint32_t slots, uint32_t sf_degree, double scale);
int ace_ct_load(ace_ctx* ctx, const char* path, int64_t* c0, int64_t* c1, uint32_t max_level, uint32_t* level,
                uint32_t* slots, uint32_t* sf_degree, double* scale);

/* ---- the reference's random sources restated on the host (csrc/refrng.h; no GPU needed): the
 *      BLAKE2Xb word stream of prng.h:42-60 for a given seed (16 words) and counter; Sample_uniform,
 *      Sample_ternary (random_sample.c:38-76, 99-152) on that stream; Sample_triangle (:78-97) on
 *      glibc's rand() after srandom(seed).  Used by the exact key generation (ace_keygen_reference). */
void ace_refrng_words(const uint32_t* seed16, uint64_t counter, uint32_t* out, size_t n);
void ace_refrng_uniform(const uint32_t* seed16, uint64_t counter, int64_t* out, size_t n, uint64_t bound);
void ace_refrng_ternary(const uint32_t* seed16, uint64_t counter, int64_t* out, size_t n, int64_t hamming_weight);
void ace_refrng_triangle(uint32_t seed, int64_t* out, size_t n);

/* ---- measurement: instruction-rate peaks of the device (csrc/peaks.cu), the roofline denominators
 *      of the integer-bound kernels (NTT, base conversion).  gops[0..6] = G thread-instructions/s of
 *      IMAD.WIDE.U32, IMAD (32-bit), IADD3, DFMA, IMAD.WIDE issued 1:1 with IADD3, IMAD.HI.U32,
 *      IMAD.WIDE issued 1:1 with DFMA.  No reference
 *      counterpart (the reference reports wall time only, rtlib/include/common/rt_stat.h). */
int ace_measure_pipe_peaks(int device, double* gops, int n);
/* G butterflies/s of the NTT's radix-16 register pass with no memory traffic (csrc/ntt16.cu): the
 * arithmetic ceiling of the transform kernels on this device.  form 0: FP64 butterfly (moduli below
 * 2^50.4), 1: 64-bit integer lazy butterfly, 2: integer with conditional subtraction (58-61 bit
 * moduli); < 0 if N != 2^16 or the form is not in use. */
double ace_ntt_bfly_peak(ace_ctx* ctx, int form, int ctas_per_sm);

#ifdef __cplusplus
}
#endif
#endif /* ACE_B200_H */
