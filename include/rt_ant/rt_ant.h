/* rt_ant/rt_ant.h -- source-level drop-in for ACE's "ant" runtime provider, backed by the
 * B200 runtime (libace_b200.so).
 *
 * C code emitted by fhe_cmplr with -P2C:lib=ant includes "rt_ant/rt_ant.h"
 * (reference: fhe-cmplr/rtlib/include/rt_ant/rt_ant.h:17-21) and uses the types and
 * functions below by name, by value and by field.  This header re-declares that surface
 * with identical names, field names and argument meaning; every POLYNOMIAL._data pointer
 * (and hence everything Coeffs() returns) is a DEVICE pointer into HBM -- emitted code only
 * ever hands those pointers back to Hw_xxx / Set_coeffs, it never dereferences them.
 *
 * Reference declarations mirrored here (paths under fhe-cmplr/rtlib/):
 *   include/common/common.h:21-88   MAP_KIND, LIB_PROV, DATA_ENTRY_TYPE, MAP_DESC, SHAPE,
 *                                   DATA_SCHEME, CKKS_PARAMS, RT_DATA_INFO
 *   include/common/tensor.h:19-60   TENSOR
 *   include/common/rt_api.h:24-68   Prepare_context ... Run_main_graph + emitted callbacks
 *   include/common/pt_mgr.h:41-46   Pt_from_msg
 *   include/common/rt_stat.h        Tm_start / Tm_taken
 *   ant/include/util/fhe_utils.h:27-32      MODULUS
 *   ant/include/util/polynomial.h:35-44     POLYNOMIAL
 *   ant/include/util/ciphertext.h:32-38,346-353  CIPHERTEXT, CIPHERTEXT3
 *   ant/include/util/plaintext.h:29-34      PLAINTEXT
 *   ant/include/ckks/cipher_eval.h, plain_eval.h, ant/include/poly/poly_eval.h,
 *   poly_arith.h, ant/include/rtlib/context.h, key_gen.h     the function surface
 */
#ifndef ACE_B200_RT_ANT_H
#define ACE_B200_RT_ANT_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif


/* ---- common.h ------------------------------------------------------------------- */
typedef enum { NORMAL, CONV, CHANNEL, DIAGONAL } MAP_KIND;
typedef enum { LIB_ANT, LIB_SEAL, LIB_OPENFHE } LIB_PROV;
typedef enum { DE_MSG_F32, DE_MSG_F64, DE_PLAINTEXT } DATA_ENTRY_TYPE;

typedef struct {
  MAP_KIND _kind;
  int      _count;
  int      _start;
  int      _end;
  int      _stride;
} MAP_DESC;

typedef struct {
  size_t _n;
  size_t _c;
  size_t _h;
  size_t _w;
} SHAPE;

typedef struct {
  const char* _name;
  SHAPE       _shape;
  int         _count;
  MAP_DESC    _desc[];
} DATA_SCHEME;

typedef struct {
  LIB_PROV _provider;
  uint32_t _poly_degree;
  size_t   _sec_level;
  size_t   _mul_depth;
  size_t   _first_mod_size;
  size_t   _scaling_mod_size;
  size_t   _num_q_parts;
  size_t   _hamming_weight;
  size_t   _num_rot_idx;
  int32_t  _rot_idxs[];
} CKKS_PARAMS;

typedef struct {
  const char*     _file_name;
  const char*     _file_uuid;
  DATA_ENTRY_TYPE _entry_type;
} RT_DATA_INFO;

/* ---- tensor.h -------------------------------------------------------------------- */
typedef struct {
  SHAPE  _shape;
  double _vals[];
} TENSOR;
#define TENSOR_N(T) (T)->_shape._n
#define TENSOR_C(T) (T)->_shape._c
#define TENSOR_H(T) (T)->_shape._h
#define TENSOR_W(T) (T)->_shape._w
#define TENSOR_SIZE(T) (TENSOR_N(T) * TENSOR_C(T) * TENSOR_H(T) * TENSOR_W(T))
#define TENSOR_ELEM(T, n, c, h, w) \
  (T)->_vals[(w) + TENSOR_W(T) * ((h) + TENSOR_H(T) * ((c) + TENSOR_C(T) * (n)))]
static inline float* Slice(float* vec, size_t row_idx, size_t col) {
  return vec + row_idx * col;
}
TENSOR* Alloc_tensor(size_t n, size_t c, size_t h, size_t w, const double* val);
void    Free_tensor(TENSOR* tensor);
void    Print_tensor(FILE* fp, TENSOR* tensor);

/* ---- core types -------------------------------------------------------------------- */
typedef struct {
  int64_t           _val;
  int64_t           _br_k;
  int64_t           _br_m;
  unsigned __int128 _prec128;
} MODULUS;

typedef struct {
  uint32_t _ring_degree;
  size_t   _num_alloc_primes;
  size_t   _num_primes;
  size_t   _num_primes_p;
  bool     _is_ntt;
  int64_t* _data; /* DEVICE pointer */
} POLYNOMIAL;
typedef POLYNOMIAL* POLY;

typedef struct {
  POLYNOMIAL _c0_poly;
  POLYNOMIAL _c1_poly;
  uint32_t   _slots;
  double     _scaling_factor;
  uint32_t   _sf_degree;
} CIPHERTEXT;
typedef CIPHERTEXT* CIPHER;

typedef struct {
  POLYNOMIAL _c0_poly;
  POLYNOMIAL _c1_poly;
  POLYNOMIAL _c2_poly;
  uint32_t   _slots;
  double     _scaling_factor;
  uint32_t   _sf_degree;
} CIPHERTEXT3;
typedef CIPHERTEXT3* CIPHER3;

typedef struct {
  POLYNOMIAL _poly;
  uint32_t   _slots;
  double     _scaling_factor;
  uint32_t   _sf_degree;
} PLAINTEXT;
typedef PLAINTEXT* PLAIN;

typedef struct SWITCH_KEY SWITCH_KEY; /* opaque: lives in the runtime */
typedef SWITCH_KEY*       SW_KEY;

/* ---- rt_api.h: driver side ------------------------------------------------------------ */
void    Prepare_context(void);
void    Finalize_context(void);
void    Prepare_input(TENSOR* input, const char* name);
double* Handle_output(const char* name);
void    Run_main_graph(void);
/* provided by the emitted translation unit */
CKKS_PARAMS*  Get_context_params(void);
RT_DATA_INFO* Get_rt_data_info(void);
int           Get_input_count(void);
int           Get_output_count(void);
DATA_SCHEME*  Get_encode_scheme(int idx);
DATA_SCHEME*  Get_decode_scheme(int idx);
bool          Main_graph(void);

CIPHERTEXT Get_input_data(const char* name, size_t idx);
void       Set_output_data(const char* name, size_t idx, CIPHER data);

void Tm_start(const char* msg);
void Tm_taken(const char* msg);

/* ---- context.h ------------------------------------------------------------------------ */
uint32_t Degree(void);
double   Get_default_sc(void);
size_t   Get_q_parts(void);
size_t   Get_p_cnt(void);
MODULUS* Q_modulus(void);
MODULUS* P_modulus(void);

/* ---- poly_eval.h / poly_arith.h --------------------------------------------------------- */
POLY Alloc_poly(uint32_t degree, size_t q_primes, bool extend_p);
void Free_poly(POLY poly);
void Free_poly_data(POLY poly);
void Copy_poly(POLY res, POLY poly);
void Set_coeffs(POLY dst, uint32_t level, uint32_t degree, int64_t* src);
size_t Num_decomp(POLY poly);
static inline int64_t* Coeffs(POLY poly, size_t level, uint32_t degree) {
  return poly->_data + level * degree;
}
static inline size_t Poly_level(POLY poly) { return poly->_num_primes; }
static inline size_t Num_alloc(POLY poly) { return poly->_num_alloc_primes; }
static inline size_t Num_p(POLY poly) { return poly->_num_primes_p; }

POLY Decomp(POLY res, POLY poly, uint32_t q_part_idx);
POLY Mod_up(POLY new_poly, POLY old_poly, uint32_t q_part_idx);
POLY Decomp_modup(POLY res, POLY poly, uint32_t q_part_idx);
POLY Mod_down(POLY res, POLY poly);
POLY Rescale(POLY res, POLY poly);

int64_t* Hw_modadd(int64_t* res, int64_t* val1, int64_t* val2, MODULUS* modulus,
                   uint32_t degree);
int64_t* Hw_modmul(int64_t* res, int64_t* val1, int64_t* val2, MODULUS* modulus,
                   uint32_t degree);
int64_t* Hw_rotate(int64_t* res, int64_t* val, int64_t* rot_precomp, MODULUS* modulus,
                   uint32_t degree);

/* ---- key_gen.h ---------------------------------------------------------------------------- */
uint32_t Auto_idx(int32_t rot_idx);
int64_t* Auto_order(int32_t rot_idx);
SW_KEY   Swk(bool is_rot, int32_t rot_idx);
POLY     Pk0_at(SW_KEY swk, uint32_t idx);
POLY     Pk1_at(SW_KEY swk, uint32_t idx);

/* ---- cipher_eval.h ---------------------------------------------------------------------- */
void Init_ciph_same_scale(CIPHER res, CIPHER ciph1, CIPHER ciph2);
void Init_ciph_same_scale_plain(CIPHER res, CIPHER ciph, PLAIN plain);
void Init_ciph_same_scale_ciph3(CIPHER res, CIPHER3 ciph);
void Init_ciph3_same_scale_ciph3(CIPHER3 res, CIPHER3 ciph1, CIPHER3 ciph2);
void Init_ciph_up_scale(CIPHER res, CIPHER ciph1, CIPHER ciph2);
void Init_ciph_up_scale_plain(CIPHER res, CIPHER ciph, PLAIN plain);
void Init_ciph_down_scale(CIPHER res, CIPHER ciph);
void Init_ciph3_up_scale(CIPHER3 res, CIPHER ciph1, CIPHER ciph2);
void Copy_ciph(CIPHER res, CIPHER ciph);
void Zero_ciph(CIPHER ciph);
void Free_ciph_poly(CIPHER ciph, uint32_t cnt);
size_t   Level(CIPHER ciph);
uint32_t Sc_degree(CIPHER ciph);
uint32_t Get_slots(CIPHER ciph);
void     Set_slots(CIPHER ciph, uint32_t slots);
double*  Get_msg(CIPHER ciph);
void     Print_cipher_msg(FILE* fp, const char* name, CIPHER ciph, uint32_t len);

CIPHER  Add_ciph(CIPHER res, CIPHER ciph1, CIPHER ciph2);
CIPHER  Add_plain(CIPHER res, CIPHER ciph, PLAIN plain);
CIPHER  Sub_ciph(CIPHER res, CIPHER ciph1, CIPHER ciph2);
CIPHER  Mul_ciph(CIPHER res, CIPHER ciph1, CIPHER ciph2);
CIPHER3 Mul_ciph3(CIPHER3 res, CIPHER ciph1, CIPHER ciph2);
CIPHER  Mul_plain(CIPHER res, CIPHER ciph, PLAIN plain);
CIPHER  Relin(CIPHER res, CIPHER3 ciph);
CIPHER  Rescale_ciph(CIPHER res, CIPHER ciph);
CIPHER  Rotate_ciph(CIPHER res, CIPHER ciph, int32_t rotation);
CIPHER  Bootstrap(CIPHER res, CIPHER ciph, uint32_t level_after_bts);
CIPHER  Encrypt(CIPHER res, PLAIN plain);

/* ---- plain_eval.h / pt_mgr.h ---------------------------------------------------------------- */
void Encode_plain_from_float(PLAIN plain, float* input, size_t len, uint32_t sc_degree,
                             uint32_t level);
void Encode_plain_from_double(PLAIN plain, double* input, size_t len, uint32_t sc_degree,
                              uint32_t level);
void Free_plain_poly(PLAIN plain);
bool Pt_mgr_init(const char* fname);
void Pt_mgr_fini(void);
void Pt_from_msg(void* pt, uint32_t index, size_t len, uint32_t scale, uint32_t level);
void Pt_from_msg_validate(void* pt, float* buf, uint32_t index, size_t len, uint32_t scale,
                          uint32_t level); /* include/common/pt_mgr.h:44-46 */
/* pre-encoded weights (DE_PLAINTEXT data files; include/common/pt_mgr.h:28-38): Pt_get returns a
 * PLAINTEXT whose limbs are in HBM, valid until its slot (index % PT_ENTRY_COUNT) is used again */
void  Pt_prefetch(uint32_t index);
void* Pt_get(uint32_t index, size_t len, uint32_t scale, uint32_t level);
void* Pt_get_validate(float* buf, uint32_t index, size_t len, uint32_t scale, uint32_t level);
void  Pt_free(uint32_t index);

/* ---- B200 extensions (not in the reference): binding the runtime to a device and moving
 *      limbs across the host/device boundary for tests and client code ------------------- */
void  Ace_set_device(int device);
void* Ace_context(void); /* the underlying ace_ctx* (include/ace_b200.h) */
void  Ace_download_poly(int64_t* host_dst, POLY poly);  /* all q (+p) limbs */
void  Ace_upload_poly(POLY poly, const int64_t* host_src);
void  Ace_import_switch_key(bool is_rot, int32_t rot_idx, uint32_t part, int which,
                            const int64_t* host_poly);
void  Ace_set_input(const char* name, size_t idx, const int64_t* c0, const int64_t* c1,
                    uint32_t level, uint32_t slots, double scale, uint32_t sf_degree);
CIPHER Ace_get_output(const char* name, size_t idx);
void     Ace_timer_start(void);    /* CUDA event on the runtime's stream */
float    Ace_timer_stop_ms(void);  /* ms since Ace_timer_start, device time */
uint64_t Ace_launch_count(void);   /* kernels launched so far */
int      Ace_bootstrap_rot_indices(uint32_t slots, int32_t* out, size_t cap);
int      Ace_trace(uint64_t* out, size_t cap); /* op trace [8 classes][72 levels], see rt_shim.cu */

#ifdef __cplusplus
}
#endif
#endif /* ACE_B200_RT_ANT_H */
