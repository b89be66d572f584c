/* Declarations-only stand-in for <gmp.h> (TEST INFRASTRUCTURE, not product code).
 *
 * The image ships libgmp.so.10 but no development header.  The reference rtlib
 * reaches GMP only through the macros of
 *   fhe-cmplr/rtlib/ant/include/util/fhe_bignumber.h:19-89
 * so it is enough to declare the handful of exported __gmpz_ / __gmpf_ entry
 * points behind those macros.  Struct layouts follow the stable GMP 6 ABI
 * (libgmp.so.10).  Used only by oracle/Makefile to compile the reference
 * sources where they lie into oracle/_ref/.
 */
#ifndef ORACLE_GMP_SHIM_H
#define ORACLE_GMP_SHIM_H
#include <stddef.h>
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef unsigned long mp_limb_t;
typedef long          mp_exp_t;
typedef unsigned long mp_bitcnt_t;
typedef struct { int _mp_alloc; int _mp_size; mp_limb_t* _mp_d; } __mpz_struct;
typedef __mpz_struct mpz_t[1];
typedef __mpz_struct*       mpz_ptr;
typedef const __mpz_struct* mpz_srcptr;
typedef struct { int _mp_prec; int _mp_size; mp_exp_t _mp_exp; mp_limb_t* _mp_d; } __mpf_struct;
typedef __mpf_struct mpf_t[1];
typedef __mpf_struct*       mpf_ptr;
typedef const __mpf_struct* mpf_srcptr;
typedef struct { mpz_t _mp_seed; int _mp_alg; union { void* _mp_lc; } _mp_algdata; } __gmp_randstate_struct;
typedef __gmp_randstate_struct gmp_randstate_t[1];

#define mpz_init __gmpz_init
#define mpz_inits __gmpz_inits
#define mpz_clear __gmpz_clear
#define mpz_clears __gmpz_clears
#define mpz_init_set __gmpz_init_set
#define mpz_init_set_ui __gmpz_init_set_ui
#define mpz_init_set_si __gmpz_init_set_si
#define mpz_init_set_d __gmpz_init_set_d
#define mpz_set __gmpz_set
#define mpz_set_si __gmpz_set_si
#define mpz_set_ui __gmpz_set_ui
#define mpz_set_d __gmpz_set_d
#define mpz_set_str __gmpz_set_str
#define mpz_get_ui __gmpz_get_ui
#define mpz_get_si __gmpz_get_si
#define mpz_get_d __gmpz_get_d
#define mpz_get_d_2exp __gmpz_get_d_2exp
#define mpz_fits_slong_p __gmpz_fits_slong_p
#define mpz_add __gmpz_add
#define mpz_add_ui __gmpz_add_ui
#define mpz_sub __gmpz_sub
#define mpz_sub_ui __gmpz_sub_ui
#define mpz_mul __gmpz_mul
#define mpz_mul_ui __gmpz_mul_ui
#define mpz_mul_si __gmpz_mul_si
#define mpz_mul_2exp __gmpz_mul_2exp
#define mpz_addmul __gmpz_addmul
#define mpz_addmul_ui __gmpz_addmul_ui
#define mpz_pow_ui __gmpz_pow_ui
#define mpz_powm __gmpz_powm
#define mpz_sqrt __gmpz_sqrt
#define mpz_invert __gmpz_invert
#define mpz_fdiv_q __gmpz_fdiv_q
#define mpz_fdiv_r __gmpz_fdiv_r
#define mpz_fdiv_q_ui __gmpz_fdiv_q_ui
#define mpz_fdiv_r_ui __gmpz_fdiv_r_ui
#define mpz_fdiv_q_2exp __gmpz_fdiv_q_2exp
#define mpz_cmp __gmpz_cmp
#define mpz_cmp_si __gmpz_cmp_si
#define mpz_cmp_ui __gmpz_cmp_ui
#define mpz_sizeinbase __gmpz_sizeinbase
#define mpz_urandomm __gmpz_urandomm
#define mpf_init __gmpf_init
#define mpf_inits __gmpf_inits
#define mpf_clear __gmpf_clear
#define mpf_clears __gmpf_clears
#define mpf_init_set __gmpf_init_set
#define mpf_init_set_si __gmpf_init_set_si
#define mpf_init_set_d __gmpf_init_set_d
#define mpf_set __gmpf_set
#define mpf_set_z __gmpf_set_z
#define mpf_set_si __gmpf_set_si
#define mpf_set_d __gmpf_set_d
#define mpf_mul __gmpf_mul
#define mpf_mul_ui __gmpf_mul_ui
#define mpf_div __gmpf_div
#define mpf_get_d __gmpf_get_d
#define mpf_get_str __gmpf_get_str
#define gmp_printf __gmp_printf
#define gmp_fprintf __gmp_fprintf

void mpz_init(mpz_ptr);
void mpz_inits(mpz_ptr, ...);
void mpz_clear(mpz_ptr);
void mpz_clears(mpz_ptr, ...);
void mpz_init_set(mpz_ptr, mpz_srcptr);
void mpz_init_set_ui(mpz_ptr, unsigned long);
void mpz_init_set_si(mpz_ptr, long);
void mpz_init_set_d(mpz_ptr, double);
void mpz_set(mpz_ptr, mpz_srcptr);
void mpz_set_si(mpz_ptr, long);
void mpz_set_ui(mpz_ptr, unsigned long);
void mpz_set_d(mpz_ptr, double);
int  mpz_set_str(mpz_ptr, const char*, int);
unsigned long mpz_get_ui(mpz_srcptr);
long   mpz_get_si(mpz_srcptr);
double mpz_get_d(mpz_srcptr);
double mpz_get_d_2exp(long*, mpz_srcptr);
int    mpz_fits_slong_p(mpz_srcptr);
void mpz_add(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_add_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_sub(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_sub_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_mul(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_mul_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_mul_si(mpz_ptr, mpz_srcptr, long);
void mpz_mul_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void mpz_addmul(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_addmul_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_pow_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_powm(mpz_ptr, mpz_srcptr, mpz_srcptr, mpz_srcptr);
void mpz_sqrt(mpz_ptr, mpz_srcptr);
int  mpz_invert(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_fdiv_q(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_fdiv_r(mpz_ptr, mpz_srcptr, mpz_srcptr);
unsigned long mpz_fdiv_q_ui(mpz_ptr, mpz_srcptr, unsigned long);
unsigned long mpz_fdiv_r_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_fdiv_q_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
int  mpz_cmp(mpz_srcptr, mpz_srcptr);
int  mpz_cmp_si(mpz_srcptr, long);
int  mpz_cmp_ui(mpz_srcptr, unsigned long);
size_t mpz_sizeinbase(mpz_srcptr, int);
void mpz_urandomm(mpz_ptr, gmp_randstate_t, mpz_srcptr);
void mpf_init(mpf_ptr);
void mpf_inits(mpf_ptr, ...);
void mpf_clear(mpf_ptr);
void mpf_clears(mpf_ptr, ...);
void mpf_init_set(mpf_ptr, mpf_srcptr);
void mpf_init_set_si(mpf_ptr, long);
void mpf_init_set_d(mpf_ptr, double);
void mpf_set(mpf_ptr, mpf_srcptr);
void mpf_set_z(mpf_ptr, mpz_srcptr);
void mpf_set_si(mpf_ptr, long);
void mpf_set_d(mpf_ptr, double);
void mpf_mul(mpf_ptr, mpf_srcptr, mpf_srcptr);
void mpf_mul_ui(mpf_ptr, mpf_srcptr, unsigned long);
void mpf_div(mpf_ptr, mpf_srcptr, mpf_srcptr);
double mpf_get_d(mpf_srcptr);
char*  mpf_get_str(char*, mp_exp_t*, int, size_t, mpf_srcptr);
int gmp_printf(const char*, ...);
int gmp_fprintf(FILE*, const char*, ...);
#ifdef __cplusplus
}
#endif
#endif
