/* oracle/ckks_oracle.h -- TEST INFRASTRUCTURE ONLY.
 * CPU restatement (plain C, scalar, __int128 arithmetic) of the CKKS evaluation hot path of
 * ACE's rtlib.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * this; the product (ace_compiler_b200/) never does.
 * Pinned against the reference itself (oracle/_ref/libace_ref.so, built from /root/reference
 * by oracle/Makefile) in tests/test_oracle_vs_ref.py and against the committed golden
 * fixtures in tests/golden/.
 */
#ifndef CKKS_ORACLE_H
#define CKKS_ORACLE_H
#include <stddef.h>
#include <stdint.h>

typedef struct orc_ctx orc_ctx;

orc_ctx* orc_create(uint32_t degree, size_t mul_depth, size_t first_mod_size,
                    size_t scaling_mod_size, size_t num_q_parts);
void     orc_destroy(orc_ctx* c);
uint32_t orc_degree(const orc_ctx* c);
size_t   orc_num_q(const orc_ctx* c);
size_t   orc_num_p(const orc_ctx* c);
size_t   orc_part_size(const orc_ctx* c);
void     orc_get_primes(const orc_ctx* c, int64_t* q, int64_t* p);
int64_t  orc_psi(const orc_ctx* c, int is_p, size_t idx);

/* mod index g: 0..L-1 = Q primes, L..L+K-1 = P primes */
void orc_ntt(const orc_ctx* c, size_t g, int64_t* data);
void orc_intt(const orc_ctx* c, size_t g, int64_t* data);
void orc_hw_modadd(const orc_ctx* c, int64_t* r, const int64_t* a, const int64_t* b, size_t g);
void orc_hw_modmul(const orc_ctx* c, int64_t* r, const int64_t* a, const int64_t* b, size_t g);
void orc_hw_rotate(const orc_ctx* c, int64_t* r, const int64_t* a, const int64_t* order, size_t g);
uint32_t orc_auto_index(const orc_ctx* c, int32_t rot_idx);
void     orc_auto_order(const orc_ctx* c, uint32_t auto_idx, int64_t* order);

size_t orc_num_decomp(const orc_ctx* c, size_t num_q);
void   orc_decomp_modup(const orc_ctx* c, int64_t* out, const int64_t* in, size_t num_q,
                        size_t part);
void   orc_mod_down(const orc_ctx* c, int64_t* out, const int64_t* in, size_t num_q);
void   orc_rescale(const orc_ctx* c, int64_t* out, const int64_t* in, size_t num_q);

/* hybrid key switch of one polynomial d (num_q limbs, NTT form) with an imported key:
 * key0/key1 = dnum polys of (L+K) limbs each, laid out [part][limb][N].
 * Follows the emitted Rotate()/Relinearize() bodies. out0/out1: num_q limbs. */
void orc_key_switch(const orc_ctx* c, int64_t* out0, int64_t* out1, const int64_t* d,
                    size_t num_q, const int64_t* key0, const int64_t* key1);
/* rotation exactly as the emitted Rotate(): key switch c1, add c0, automorphism */
void orc_ct_rotate(const orc_ctx* c, int64_t* r0, int64_t* r1, const int64_t* c0,
                   const int64_t* c1, size_t num_q, uint32_t auto_idx, const int64_t* key0,
                   const int64_t* key1);
/* tensor product + relinearisation as emitted (Init_ciph3_up_scale + Relinearize()) */
void orc_ct_mul_relin(const orc_ctx* c, int64_t* r0, int64_t* r1, const int64_t* a0,
                      const int64_t* a1, const int64_t* b0, const int64_t* b1, size_t num_q,
                      const int64_t* key0, const int64_t* key1);

/* CKKS encode of real values (Encode_at_level_with_sf, 64-bit path); out = level limbs */
void orc_encode(const orc_ctx* c, int64_t* out, const double* vals, size_t len, uint32_t level,
                uint32_t slots, uint32_t sf_degree);
#endif
