/* pt_get_case.c -- pre-encoded weights through Pt_get, written in the style of an
 * ACE-emitted unit compiled with -P2C:ct_encode (fhe-cmplr/include/fhe/ckks/ir2c_ctx.h:75-92:
 * `dest = *(PLAIN)Pt_get(index, len, scale, level)`), because the reference checks in no program
 * that uses the plaintext path.
 *
 * The DE_PLAINTEXT data file is made by tests/test_gpu_ptmgr.py with the REFERENCE's own
 * Encode_plain_buffer (plain_eval.c:98-124) and laid out like RT_DATA_WRITER does
 * (rt_data_writer.h:28-106).  Checks:
 *   1. every plaintext handed out by Pt_get equals, limb for limb, the run-time encode of the
 *      same message (Encode_plain_from_float) -- file parsing and both encoders;
 *   2. output = sum_i input * w_i, recorded through the polynomial-level API while the ring of
 *      PT_ENTRY_COUNT slots is recycled under the deferred multiplications;
 *   3. the decrypted output against the expected values.
 * env: PT_CASE_MSGS = file with n_entries x len float32 messages. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common/rtlib.h"
#include "rt_ant/rt_ant.h"

#define N_ENT 5
#define MSG_LEN 2048

CKKS_PARAMS* Get_context_params() {
  static CKKS_PARAMS param = {LIB_ANT, 4096, 0, 7, 51, 50, 3, 192, 0};
  return &param;
}
DATA_SCHEME* Get_encode_scheme(int idx) {
  static DATA_SCHEME scheme_0 = {"input", {0, 0, 0, 0}, 1, {NORMAL, 0, 0, 0, 0}};
  return &scheme_0;
}
DATA_SCHEME* Get_decode_scheme(int idx) {
  static DATA_SCHEME scheme = {"output", {0, 0, 0, 0}, 1, {NORMAL, 0, 0, 0, 0}};
  return &scheme;
}
RT_DATA_INFO* Get_rt_data_info() {
  static RT_DATA_INFO info = {"pt_get_case.pt", "XXXXXXXX-XXXX-XXXX-XXXX-XXXXXXXXXXXX", DE_PLAINTEXT};
  return &info; /* the test points ACE_B200_DATA_FILE at the real file */
}
int Get_output_count() { return 1; }
int Get_input_count() { return 1; }

static float Msgs[N_ENT][MSG_LEN];
static int   Identical = 0;

static int same_limbs(PLAINTEXT* a, PLAINTEXT* b) {
  size_t n = a->_poly._num_primes * (size_t)a->_poly._ring_degree;
  if (a->_poly._num_primes != b->_poly._num_primes || a->_slots != b->_slots ||
      a->_scaling_factor != b->_scaling_factor || a->_sf_degree != b->_sf_degree)
    return 0;
  int64_t* ha = malloc(n * sizeof(int64_t));
  int64_t* hb = malloc(n * sizeof(int64_t));
  Ace_download_poly(ha, &a->_poly);
  Ace_download_poly(hb, &b->_poly);
  int same = memcmp(ha, hb, n * sizeof(int64_t)) == 0;
  free(ha);
  free(hb);
  return same;
}

bool Main_graph() {
  CIPHERTEXT input, output;
  uint32_t   degree = Degree();
  input = Get_input_data("input", 0);
  memset(&output, 0, sizeof(output));
  PLAINTEXT w0 = *(PLAIN)Pt_get(0 /* cst_0 */, MSG_LEN, 1, 0);
  Init_ciph_up_scale_plain(&output, &input, &w0);
  POLY tmp = Alloc_poly(degree, 1, 0);
  for (uint32_t i = 0; i < N_ENT; i++) {
    /* the slot of entry i - PT_ENTRY_COUNT is recycled here, its multiplications still deferred */
    PLAINTEXT w = *(PLAIN)Pt_get(i /* cst_i */, MSG_LEN, 1, 0);
    MODULUS*  modulus = Q_modulus();
    for (uint32_t l = 0; l < Level(&output); l++) {
      for (int c = 0; c < 2; c++) {
        POLY     src = c ? &input._c1_poly : &input._c0_poly;
        POLY     dst = c ? &output._c1_poly : &output._c0_poly;
        Hw_modmul(Coeffs(tmp, 0, degree), Coeffs(src, l, degree), Coeffs(&w._poly, l, degree), modulus, degree);
        Hw_modadd(Coeffs(dst, l, degree), Coeffs(dst, l, degree), Coeffs(tmp, 0, degree), modulus, degree);
      }
      modulus++;
    }
  }
  Free_poly(tmp);
  /* now that everything above is recorded: the plaintexts themselves, one at a time (each
   * comparison downloads, i.e. flushes) */
  for (uint32_t i = 0; i < N_ENT; i++) {
    PLAINTEXT w = *(PLAIN)Pt_get(i, MSG_LEN, 1, 0);
    PLAINTEXT e;
    memset(&e, 0, sizeof(e));
    Encode_plain_from_float(&e, Msgs[i], MSG_LEN, 1, 0);
    if (same_limbs(&w, &e)) Identical++;
    else printf("plaintext %u from the file differs from its run-time encode\n", i);
    Free_plain_poly(&e);
  }
  Pt_free(N_ENT - 1);
  Init_ciph_down_scale(&output, &output);
  Rescale(&(output._c0_poly), &(output._c0_poly));
  Rescale(&(output._c1_poly), &(output._c1_poly));
  Set_output_data("output", 0, &output);
  return true;
}

int main(int argc, char* argv[]) {
  const char* path = getenv("PT_CASE_MSGS");
  FILE*       f    = path ? fopen(path, "rb") : NULL;
  if (!f || fread(Msgs, sizeof(float), N_ENT * MSG_LEN, f) != N_ENT * MSG_LEN) {
    printf("PT_CASE_MSGS missing or short\n");
    return 2;
  }
  fclose(f);
  Prepare_context();
  double x[MSG_LEN];
  for (int k = 0; k < MSG_LEN; k++) x[k] = sin(0.37 * k) * 0.9;
  TENSOR* in = Alloc_tensor(1, 1, 1, MSG_LEN, x);
  Prepare_input(in, "input");
  Free_tensor(in);
  Run_main_graph();
  double* result = Handle_output("output");
  Finalize_context();
  double worst = 0;
  for (int k = 0; k < MSG_LEN; k++) {
    double want = 0;
    for (int i = 0; i < N_ENT; i++) want += x[k] * (double)Msgs[i][k];
    double err = fabs(result[k] - want);
    if (err > worst) worst = err;
  }
  free(result);
  printf("plaintexts identical to their run-time encode: %d of %d; max |error| of sum_i x*w_i: %.3e\n",
         Identical, N_ENT, worst);
  if (Identical == N_ENT && worst < 1e-6) {
    printf("SUCESS!\n");
    return 0;
  }
  printf("FAILED!\n");
  return 1;
}
