// capi.cu -- extern "C" boundary (include/ace_b200.h) over ace::Context.
// No torch types, no exceptions across the boundary: errors become negative return codes
// plus a thread-local message.
#include <cstring>
#include <functional>
#include <string>

#include "../../include/ace_b200.h"
#include "evaluator.h"
#include "refrng.h"

using namespace ace;

struct ace_ctx {
  Context*    c;
  Evaluator*  ev;
  cudaEvent_t ev0, ev1;
};

static thread_local std::string g_err;

#define ACE_TRY(body)                     \
  try {                                   \
    if (!ctx) { g_err = "null context"; return -1; } \
    cudaSetDevice(ctx->c->device);        \
    body;                                 \
    return 0;                             \
  } catch (const std::exception& e) {     \
    g_err = e.what();                     \
    return -1;                            \
  }

static inline u64*       U(int64_t* p) { return reinterpret_cast<u64*>(p); }
static inline const u64* U(const int64_t* p) { return reinterpret_cast<const u64*>(p); }

extern "C" {

const char* ace_last_error(void) { return g_err.c_str(); }

int ace_ctx_create(ace_ctx** out, uint32_t poly_degree, size_t mul_depth, size_t first_mod_size,
                   size_t scaling_mod_size, size_t num_q_parts, size_t hamming_weight,
                   int device) {
  try {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
      g_err = "no CUDA device: libace_b200 has no CPU fallback";
      return -2;
    }
    Params p{poly_degree, mul_depth, first_mod_size, scaling_mod_size, num_q_parts,
             hamming_weight};
    ace_ctx* h = new ace_ctx;
    h->c       = new Context(p, device);
    h->ev      = new Evaluator(h->c);
    cudaEventCreate(&h->ev0);
    cudaEventCreate(&h->ev1);
    *out = h;
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

void ace_ctx_destroy(ace_ctx* ctx) {
  if (!ctx) return;
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  delete ctx->ev;
  delete ctx->c;
  delete ctx;
}

uint32_t ace_degree(const ace_ctx* ctx) { return ctx->c->N; }
size_t   ace_num_q(const ace_ctx* ctx) { return ctx->c->L; }
size_t   ace_num_p(const ace_ctx* ctx) { return ctx->c->K; }
size_t   ace_num_q_parts(const ace_ctx* ctx) { return ctx->c->dnum; }
size_t   ace_part_size(const ace_ctx* ctx) { return ctx->c->part_size; }
int      ace_get_primes(const ace_ctx* ctx, int64_t* q, int64_t* p) {
  for (size_t i = 0; i < ctx->c->L; i++) q[i] = (int64_t)ctx->c->mod[i];
  for (size_t i = 0; i < ctx->c->K; i++) p[i] = (int64_t)ctx->c->mod[ctx->c->L + i];
  return 0;
}
int64_t  ace_psi(const ace_ctx* ctx, uint32_t g) { return (int64_t)ctx->c->psi[g]; }
size_t   ace_num_decomp(const ace_ctx* ctx, size_t num_q) { return ctx->c->num_decomp(num_q); }
uint64_t ace_launch_count(const ace_ctx* ctx) { return ctx->c->launches; }

int64_t* ace_alloc_limbs(ace_ctx* ctx, size_t n_limbs, int zero) {
  try {
    cudaSetDevice(ctx->c->device);
    return reinterpret_cast<int64_t*>(ctx->c->alloc_limbs(n_limbs, zero != 0));
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
int ace_free_limbs(ace_ctx* ctx, int64_t* dev) { ACE_TRY(ctx->c->free_limbs(U(dev))) }
int ace_upload(ace_ctx* ctx, int64_t* d, const int64_t* s, size_t n) {
  ACE_TRY(ctx->c->upload(U(d), U(s), n))
}
int ace_download(ace_ctx* ctx, int64_t* d, const int64_t* s, size_t n) {
  ACE_TRY(ctx->c->download(U(d), U(s), n))
}
int ace_copy_limbs(ace_ctx* ctx, int64_t* d, const int64_t* s, size_t n) {
  ACE_TRY(ACE_CUDA(cudaMemcpyAsync(d, s, n * ctx->c->N * sizeof(u64), cudaMemcpyDeviceToDevice,
                                   ctx->c->stream)))
}
int ace_zero_limbs(ace_ctx* ctx, int64_t* d, size_t n) {
  ACE_TRY(ACE_CUDA(cudaMemsetAsync(d, 0, n * ctx->c->N * sizeof(u64), ctx->c->stream)))
}
int ace_sync(ace_ctx* ctx) { ACE_TRY(ctx->c->sync()) }

static void check_g(Context* c, uint32_t g0, uint32_t n) {
  if ((size_t)g0 + n > c->G) throw std::runtime_error("modulus index out of range");
}

int ace_hw_modadd(ace_ctx* ctx, int64_t* r, const int64_t* a, const int64_t* b, uint32_t g0,
                  uint32_t n) {
  ACE_TRY(check_g(ctx->c, g0, n); launch_ew(ctx->c->T, EW_ADD, U(r), U(a), U(b), g0, n, ctx->c->stream); ctx->c->launches++)
}
int ace_hw_modsub(ace_ctx* ctx, int64_t* r, const int64_t* a, const int64_t* b, uint32_t g0,
                  uint32_t n) {
  ACE_TRY(check_g(ctx->c, g0, n); launch_ew(ctx->c->T, EW_SUB, U(r), U(a), U(b), g0, n, ctx->c->stream); ctx->c->launches++)
}
int ace_hw_modmul(ace_ctx* ctx, int64_t* r, const int64_t* a, const int64_t* b, uint32_t g0,
                  uint32_t n) {
  ACE_TRY(check_g(ctx->c, g0, n); launch_ew(ctx->c->T, EW_MUL, U(r), U(a), U(b), g0, n, ctx->c->stream); ctx->c->launches++)
}
int ace_hw_rotate(ace_ctx* ctx, int64_t* r, const int64_t* a, const int64_t* order, uint32_t g0,
                  uint32_t n) {
  ACE_TRY(check_g(ctx->c, g0, n); launch_gather(ctx->c->T, U(r), U(a), order, g0, n, ctx->c->stream); ctx->c->launches++)
}
int ace_ntt(ace_ctx* ctx, int64_t* d, uint32_t g0, uint32_t n) {
  ACE_TRY(check_g(ctx->c, g0, n); ctx->c->ntt(U(d), g0, n))
}
int ace_intt(ace_ctx* ctx, int64_t* d, uint32_t g0, uint32_t n) {
  ACE_TRY(check_g(ctx->c, g0, n); ctx->c->intt(U(d), g0, n))
}

static void check_level(Context* c, uint32_t num_q) {
  if (num_q == 0 || num_q > c->L) throw std::runtime_error("num_q out of range");
}
int ace_decomp_modup(ace_ctx* ctx, int64_t* out, const int64_t* in, uint32_t num_q,
                     uint32_t part) {
  ACE_TRY(check_level(ctx->c, num_q);
          if (part >= ctx->c->num_decomp(num_q)) throw std::runtime_error("q_part_idx out of range");
          ctx->c->decomp_modup(U(out), U(in), num_q, part))
}
int ace_mod_down(ace_ctx* ctx, int64_t* out, const int64_t* in, uint32_t num_q) {
  ACE_TRY(check_level(ctx->c, num_q); ctx->c->mod_down(U(out), U(in), num_q))
}
int ace_rescale(ace_ctx* ctx, int64_t* out, const int64_t* in, uint32_t num_q) {
  ACE_TRY(check_level(ctx->c, num_q); ctx->c->rescale(U(out), U(in), num_q))
}

uint32_t ace_auto_index(const ace_ctx* ctx, int32_t rot) { return ctx->c->auto_index(rot); }
const int64_t* ace_auto_order(ace_ctx* ctx, int32_t rot) {
  try {
    cudaSetDevice(ctx->c->device);
    return ctx->c->auto_order(ctx->c->auto_index(rot));
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
static SwitchKey& key_of(Context* c, int is_rot, int32_t rot) {
  return is_rot ? c->rot_key(c->auto_index(rot)) : c->relin_key;
}
int ace_swk_import(ace_ctx* ctx, int is_rot, int32_t rot, uint32_t part, int which,
                   const int64_t* host) {
  ACE_TRY(ctx->c->import_key_limbs(key_of(ctx->c, is_rot, rot), part, which, U(host)))
}
const int64_t* ace_swk_poly(ace_ctx* ctx, int is_rot, int32_t rot, uint32_t part, int which) {
  Context* c = ctx->c;
  if (is_rot && !c->has_rot_key(c->auto_index(rot))) return nullptr;
  SwitchKey& k = key_of(c, is_rot, rot);
  u64* base = which ? k.k1 : k.k0;
  if (!base || part >= c->dnum) return nullptr;
  return reinterpret_cast<const int64_t*>(base + (size_t)part * c->G * c->N);
}

int ace_key_switch(ace_ctx* ctx, int64_t* o0, int64_t* o1, const int64_t* d, uint32_t num_q,
                   int is_rot, int32_t rot) {
  ACE_TRY(check_level(ctx->c, num_q);
          if (is_rot && !ctx->c->has_rot_key(ctx->c->auto_index(rot))) throw std::runtime_error("rotation key not loaded");
          ctx->c->key_switch(U(o0), U(o1), U(d), num_q, key_of(ctx->c, is_rot, rot), nullptr))
}
int ace_ct_rotate(ace_ctx* ctx, int64_t* r0, int64_t* r1, const int64_t* c0, const int64_t* c1,
                  uint32_t num_q, int32_t rot) {
  ACE_TRY(check_level(ctx->c, num_q); ctx->c->ct_rotate(U(r0), U(r1), U(c0), U(c1), num_q, rot))
}
int ace_ct_rotate_hoisted(ace_ctx* ctx, int64_t* const* r0, int64_t* const* r1, const int64_t* c0,
                          const int64_t* c1, uint32_t num_q, const int32_t* rots, size_t n) {
  ACE_TRY(check_level(ctx->c, num_q);
          ctx->c->ct_rotate_hoisted(reinterpret_cast<u64* const*>(r0), reinterpret_cast<u64* const*>(r1), U(c0),
                                    U(c1), num_q, rots, n))
}
int ace_ct_mul_plain_acc(ace_ctx* ctx, int64_t* acc0, int64_t* acc1, const int64_t* c0, const int64_t* c1,
                         const int64_t* pt, uint32_t num_q, int first) {
  ACE_TRY(check_level(ctx->c, num_q);
          ctx->c->ct_mul_plain_acc(U(acc0), U(acc1), U(c0), U(c1), U(pt), num_q, first != 0))
}
int ace_ct_mul_relin(ace_ctx* ctx, int64_t* r0, int64_t* r1, const int64_t* a0,
                     const int64_t* a1, const int64_t* b0, const int64_t* b1, uint32_t num_q) {
  ACE_TRY(check_level(ctx->c, num_q);
          ctx->c->ct_mul_relin(U(r0), U(r1), U(a0), U(a1), U(b0), U(b1), num_q))
}
int ace_ct_rescale(ace_ctx* ctx, int64_t* r0, int64_t* r1, const int64_t* c0, const int64_t* c1,
                   uint32_t num_q) {
  ACE_TRY(check_level(ctx->c, num_q); ctx->c->rescale(U(r0), U(c0), num_q);
          ctx->c->rescale(U(r1), U(c1), num_q))
}

int ace_keygen(ace_ctx* ctx, uint64_t seed, const int32_t* rots, size_t n) {
  ACE_TRY(ctx->c->keygen(seed, rots, n))
}
int ace_sk_import(ace_ctx* ctx, const int64_t* sk) { ACE_TRY(ctx->c->import_secret_key(U(sk))) }
int ace_pk_import(ace_ctx* ctx, const int64_t* p0, const int64_t* p1) {
  ACE_TRY(ctx->c->import_public_key(U(p0), U(p1)))
}
int ace_encode(ace_ctx* ctx, int64_t* out, const double* vals, size_t len, uint32_t level,
               uint32_t slots, uint32_t sf_degree, uint32_t p_cnt) {
  ACE_TRY(ctx->c->encode(U(out), vals, len, level, slots, sf_degree, p_cnt))
}
int ace_encode_value(ace_ctx* ctx, int64_t* out, double value, uint32_t level,
                     uint32_t sf_degree) {
  ACE_TRY(ctx->c->encode_value(U(out), value, level, sf_degree))
}
int ace_encrypt(ace_ctx* ctx, int64_t* c0, int64_t* c1, const int64_t* pt, uint32_t level,
                uint64_t seed) {
  ACE_TRY(check_level(ctx->c, level); ctx->c->encrypt(U(c0), U(c1), U(pt), level, seed))
}
int ace_decrypt(ace_ctx* ctx, int64_t* pt, const int64_t* c0, const int64_t* c1,
                uint32_t level) {
  ACE_TRY(check_level(ctx->c, level); ctx->c->decrypt(U(pt), U(c0), U(c1), level))
}
int ace_decode(ace_ctx* ctx, double* re, double* im, const int64_t* pt, uint32_t level,
               uint32_t slots, double scale) {
  ACE_TRY(check_level(ctx->c, level); ctx->c->decode(re, im, U(pt), level, slots, scale))
}

// ---- bootstrap
int ace_bootstrap_depth(const ace_ctx* ctx) {
  return (int)Evaluator::bootstrap_depth(ctx->c->params.hamming_weight);
}
int ace_bootstrap_setup(ace_ctx* ctx, uint32_t slots) { ACE_TRY(ctx->ev->bootstrap_setup(slots)) }
int ace_bootstrap_rot_indices(ace_ctx* ctx, uint32_t slots, int32_t* out, size_t cap) {
  try {
    if (!ctx) { g_err = "null context"; return -1; }
    std::vector<int32_t> v = ctx->ev->bootstrap_rot_indices(slots);
    for (size_t i = 0; i < v.size() && i < cap; i++) out[i] = v[i];
    return (int)v.size();
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
static void run_ct(ace_ctx* ctx, int64_t* r0, int64_t* r1, uint32_t* out_level, double* out_scale,
                   uint32_t* out_sf_degree, const int64_t* c0, const int64_t* c1, uint32_t level,
                   uint32_t slots, double scale, uint32_t sf_degree,
                   const std::function<void(Ct&, Ct&)>& op) {
  Ct in;
  in.c0 = const_cast<u64*>(U(c0)); in.c1 = const_cast<u64*>(U(c1));
  in.nq = level; in.np = 0; in.cap = level; in.sf = scale; in.sfd = sf_degree;
  in.slots = slots ? slots : ctx->c->N / 2;
  Ct out;
  op(out, in);
  size_t bytes = (size_t)out.nq * ctx->c->N * sizeof(u64);
  ACE_CUDA(cudaMemcpyAsync(r0, out.c0, bytes, cudaMemcpyDeviceToDevice, ctx->c->stream));
  ACE_CUDA(cudaMemcpyAsync(r1, out.c1, bytes, cudaMemcpyDeviceToDevice, ctx->c->stream));
  *out_level = out.nq; *out_scale = out.sf; *out_sf_degree = out.sfd;
  ctx->ev->release(out);
}
int ace_bootstrap_linear(ace_ctx* ctx, int64_t* r0, int64_t* r1, uint32_t* out_level,
                         double* out_scale, uint32_t* out_sf_degree, const int64_t* c0,
                         const int64_t* c1, uint32_t level, uint32_t slots, double scale,
                         uint32_t sf_degree, int encoding) {
  ACE_TRY(check_level(ctx->c, level);
          run_ct(ctx, r0, r1, out_level, out_scale, out_sf_degree, c0, c1, level, slots, scale,
                 sf_degree, [&](Ct& o, Ct& i) { ctx->ev->linear_transform(o, i, encoding != 0); }))
}
const int64_t* ace_bootstrap_plain(ace_ctx* ctx, uint32_t slots, int encoding, uint32_t step,
                                   uint32_t idx, uint32_t* level) {
  try {
    cudaSetDevice(ctx->c->device);
    return reinterpret_cast<const int64_t*>(ctx->ev->diagonal_plain(slots, encoding != 0, step, idx, level));
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
size_t ace_bootstrap_fft_diagonals(uint32_t slots, uint32_t level_budget, int flag, int encoding,
                                   double* out) {
  return Evaluator::fft_diagonals(slots, level_budget, flag != 0, encoding != 0, out);
}
int ace_keygen_rotations(ace_ctx* ctx, uint64_t seed, const int32_t* rots, size_t n) {
  ACE_TRY((void)seed;  // the stream is the context's (ace_keygen / ace_keygen_reference)
          ctx->c->keygen_rotations(rots, n))
}
int ace_keygen_reference(ace_ctx* ctx, const uint32_t* seed16, uint64_t counter, uint32_t tri_base,
                         const int32_t* rots, size_t n) {
  ACE_TRY(ctx->c->keygen_reference(seed16, counter, tri_base, rots, n))
}
int ace_keygen_reference_stream(ace_ctx* ctx, const uint32_t* seed16, uint64_t counter, uint32_t srandom_seed,
                                const uint64_t* tri_pos, size_t n_pos, const int32_t* rots, size_t n) {
  ACE_TRY(ctx->c->keygen_reference_stream(seed16, counter, srandom_seed, tri_pos, n_pos, rots, n))
}
int ace_keygen_autos(ace_ctx* ctx, const uint32_t* auto_idx, size_t n) {
  ACE_TRY(for (size_t i = 0; i < n; i++)
            if (!ctx->c->has_rot_key(auto_idx[i])) ctx->c->gen_auto_key(auto_idx[i]);)
}
int ace_sk_export(ace_ctx* ctx, int64_t* host_sk_ntt_qp) {
  ACE_TRY(if (!ctx->c->sk_ntt) throw std::runtime_error("secret key missing");
          ctx->c->download(U(host_sk_ntt_qp), ctx->c->sk_ntt, ctx->c->G))
}
int ace_pk_export(ace_ctx* ctx, int64_t* host_pk0, int64_t* host_pk1) {
  ACE_TRY(if (!ctx->c->pk0) throw std::runtime_error("public key missing");
          ctx->c->download(U(host_pk0), ctx->c->pk0, ctx->c->L);
          ctx->c->download(U(host_pk1), ctx->c->pk1, ctx->c->L))
}
int ace_swk_export(ace_ctx* ctx, int is_rot, uint32_t auto_idx, uint32_t part, int which, int64_t* host_poly) {
  ACE_TRY(if (is_rot && !ctx->c->has_rot_key(auto_idx)) throw std::runtime_error("no such rotation key");
          SwitchKey& k = is_rot ? ctx->c->rot_key(auto_idx) : ctx->c->relin_key;
          if (!k.k0 || part >= ctx->c->dnum) throw std::runtime_error("switch key missing");
          const size_t per = ctx->c->G * (size_t)ctx->c->N;
          ctx->c->download(U(host_poly), (which ? k.k1 : k.k0) + part * per, ctx->c->G))
}
int ace_keys_save(ace_ctx* ctx, const char* path, int with_secret) { ACE_TRY(ctx->c->save_keys(path, with_secret != 0)) }
int ace_keys_load(ace_ctx* ctx, const char* path) { ACE_TRY(ctx->c->load_keys(path)) }
int ace_ct_save(ace_ctx* ctx, const char* path, const int64_t* c0, const int64_t* c1, uint32_t level,
                uint32_t slots, uint32_t sf_degree, double scale) {
  ACE_TRY(check_level(ctx->c, level); ctx->c->save_ct(path, U(c0), U(c1), level, slots, sf_degree, scale))
}
int ace_ct_load(ace_ctx* ctx, const char* path, int64_t* c0, int64_t* c1, uint32_t max_level, uint32_t* level,
                uint32_t* slots, uint32_t* sf_degree, double* scale) {
  ACE_TRY(ctx->c->load_ct(path, U(c0), U(c1), max_level, level, slots, sf_degree, scale))
}
int ace_bootstrap(ace_ctx* ctx, int64_t* r0, int64_t* r1, uint32_t* out_level, double* out_scale,
                  uint32_t* out_sf_degree, const int64_t* c0, const int64_t* c1, uint32_t level,
                  uint32_t slots, double scale, uint32_t sf_degree, uint32_t level_after_bts) {
  ACE_TRY(check_level(ctx->c, level);
          run_ct(ctx, r0, r1, out_level, out_scale, out_sf_degree, c0, c1, level, slots, scale,
                 sf_degree, [&](Ct& o, Ct& i) { ctx->ev->bootstrap(o, i, level_after_bts); }))
}

void ace_refrng_words(const uint32_t* seed16, uint64_t counter, uint32_t* out, size_t n) {
  refrng::Blake2Prng g;
  g.pin(seed16, counter);
  for (size_t i = 0; i < n; i++) out[i] = g.next();
}
void ace_refrng_uniform(const uint32_t* seed16, uint64_t counter, int64_t* out, size_t n, uint64_t bound) {
  refrng::Blake2Prng g;
  g.pin(seed16, counter);
  g.sample_uniform(out, n, bound);
}
void ace_refrng_ternary(const uint32_t* seed16, uint64_t counter, int64_t* out, size_t n, int64_t hw) {
  refrng::Blake2Prng g;
  g.pin(seed16, counter);
  g.sample_ternary(out, n, hw);
}
void ace_refrng_triangle(uint32_t seed, int64_t* out, size_t n) {
  refrng::GlibcRandom g;
  g.srandom(seed);
  g.sample_triangle(out, n);
}
double ace_ntt_bfly_peak(ace_ctx* ctx, int form, int ctas_per_sm) {
  if (!ctx || !ntt16_usable(ctx->c->T)) return -1.0;
  cudaSetDevice(ctx->c->device);
  return ntt16_bfly_peak(ctx->c->T, form, ctas_per_sm, ctx->c->stream);
}
int ace_timer_start(ace_ctx* ctx) { ACE_TRY(ACE_CUDA(cudaEventRecord(ctx->ev0, ctx->c->stream))) }
int ace_timer_stop_ms(ace_ctx* ctx, float* ms) {
  ACE_TRY(ACE_CUDA(cudaEventRecord(ctx->ev1, ctx->c->stream));
          ACE_CUDA(cudaEventSynchronize(ctx->ev1));
          ACE_CUDA(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1)))
}

}  // extern "C"
