// rt_shim.cu -- the reference's rt_ant function surface (include/rt_ant/rt_ant.h) implemented
// on the B200 runtime.  One global context per process, like the reference's
// `CKKS_CONTEXT* Context` (fhe-cmplr/rtlib/ant/src/rtlib/context.c:27).
//
// Reference behaviour mirrored (paths under fhe-cmplr/rtlib/):
//   ant/src/rtlib/context.c:29-160, ant/src/rtlib/rtlib.c:41-87, common/src/io_lib.c,
//   ant/src/ckks/cipher_eval.c:23-127,292-364, ant/include/util/ciphertext.h:179-325,
//   ant/include/util/polynomial.h:54-77,335-347,425-438, ant/src/poly/poly_eval.c:11-49,
//   ant/src/poly/poly_arith.c:14-56, ant/include/rtlib/key_gen.h:28-75,
//   ant/src/ckks/plain_eval.c:19-80, common/src/pt_mgr.c, common/src/rt_stat.c:21-28
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <map>
#include <atomic>
#include <mutex>
#include <unordered_map>
#include <string>
#include <vector>

#include "../../include/rt_ant/rt_ant.h"
#include "evaluator.h"
#include "sched.h"
#include "prof.h"

using namespace ace;

#define API extern "C" __attribute__((visibility("default")))

// callbacks of the emitted translation unit: weak, so the library also loads stand-alone
extern "C" {
__attribute__((weak)) CKKS_PARAMS*  Get_context_params(void);
__attribute__((weak)) RT_DATA_INFO* Get_rt_data_info(void);
__attribute__((weak)) DATA_SCHEME*  Get_encode_scheme(int idx);
__attribute__((weak)) DATA_SCHEME*  Get_decode_scheme(int idx);
__attribute__((weak)) bool          Main_graph(void);
}

struct SWITCH_KEY {
  SwitchKey*  key;
  POLYNOMIAL* pk0;  // [dnum] views into the device key
  POLYNOMIAL* pk1;
};

namespace {

// Execution state is per host thread, like the reference's threadprivate Input_data / Output_data
// / Tm_stat (rtlib/common/src/io_lib.c:24-25, rt_stat.c:19): the thread that calls Prepare_context
// owns the primary context; any other thread that enters the API gets a worker context (own
// stream, limb allocator, scheduler; tables and keys shared, Context::make_worker) on its first
// call.  The reference's driver runs Run_main_graph from an OpenMP loop over images
// (ant/dataset/resnet_cifar.main.inc:81); here that puts several images in flight on one GPU.
Context*   g_primary    = nullptr;
Evaluator* g_primary_ev = nullptr;
thread_local Context*   g_ctx    = nullptr;
thread_local Evaluator* g_ev     = nullptr;
thread_local Scheduler* g_queue  = nullptr;  // deferred execution of the polynomial-level API
MODULUS*  g_mod    = nullptr;  // [G] Q then P, contiguous like the reference's arrays
int       g_device = 0;
std::atomic<uint64_t> g_enc_id{1};  // one id per encryption, process-wide (ChaCha20 nonce, client.cu)
thread_local std::map<std::string, CIPHERTEXT*> g_inputs, g_outputs;
std::map<u32, SWITCH_KEY*> g_swk;  // by automorphism index; 0 = relin key (guarded by g_mu)
std::mutex g_mu;
struct Worker { Context* c; Evaluator* ev; Scheduler* q; };
std::vector<Worker> g_workers;  // every thread's state but the primary's (guarded by g_mu)
bool g_eager = false;
thread_local clock_t g_tm_stamp;

// ACE_B200_STATS=1: per-entry-point call counts and host wall time (with a stream sync around
// the timed calls), printed by Finalize_context -- where Main_graph's time goes
struct Stat { const char* name; uint64_t calls; double secs; };
Stat g_stats[] = {{"Hw_modadd", 0, 0}, {"Hw_modmul", 0, 0}, {"Hw_rotate", 0, 0},
                  {"Decomp_modup", 0, 0}, {"Mod_down", 0, 0}, {"Rescale", 0, 0},
                  {"Encode(Pt_from_msg)", 0, 0}, {"Bootstrap", 0, 0}, {"Alloc/Init", 0, 0},
                  {"Free", 0, 0}, {"Copy/Set_coeffs", 0, 0}};
enum { ST_ADD, ST_MUL, ST_ROT, ST_MODUP, ST_MODDOWN, ST_RESCALE, ST_ENCODE, ST_BTS, ST_ALLOC,
       ST_FREE, ST_COPY, ST_COUNT };
// ACE_B200_QUIET=1: no [RT_STAT] lines and, more importantly, no stream synchronisation at
// every layer boundary (the emitted code calls Tm_taken ~65 times per image)
bool quiet() {
  static int q = -1;
  if (q < 0) q = (getenv("ACE_B200_QUIET") && getenv("ACE_B200_QUIET")[0] == '1') ? 1 : 0;
  return q == 1;
}
bool g_stats_on = false, g_stats_sync = false;  // =2: sync around every scope (true GPU time)
void stats_sync();
void stats_flush();
inline double wall() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
// only Bootstrap executes inside its call; everything else is deferred (sched.h)
const char* const kApiScope[] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                 "api.Bootstrap", nullptr, nullptr, nullptr};
cudaStream_t prof_stream();
struct StatScope {
  int id; double t0;
  int pslot = -1;
  explicit StatScope(int i) : id(i), t0(0) {
    if (prof::on && kApiScope[i]) {
      stats_flush();
      pslot = prof::begin(kApiScope[i], prof_stream());
    }
    if (!g_stats_on) return;
    if (g_stats_sync) stats_sync();
    t0 = wall();
  }
  ~StatScope() {
    if (pslot >= 0) prof::end(pslot, prof_stream());
    if (!g_stats_on) return;
    if (g_stats_sync) stats_sync();
    g_stats[id].calls++;
    g_stats[id].secs += wall() - t0;
  }
};

[[noreturn]] void die(const char* msg) {
  fprintf(stderr, "[ace_b200] fatal: %s\n", msg);
  abort();  // FMT_ASSERT semantics (rtlib/include/common/error.h:23-29)
}

cudaStream_t prof_stream() { return g_ctx ? g_ctx->stream : nullptr; }
void stats_flush() {
  if (g_queue) g_queue->flush();
}
void stats_sync() {
  if (!g_ctx) return;
  if (g_queue) g_queue->flush();
  cudaStreamSynchronize(g_ctx->stream);
}

// ctx_nf(): metadata only.  ctx(): the caller is about to touch device data or the stream, so
// every recorded limb op is issued first (keeps program order on the stream).
void attach_worker() {
  std::lock_guard<std::mutex> lk(g_mu);
  cudaSetDevice(g_device);
  Context* w = nullptr;
  try {
    w = g_primary->make_worker();
  } catch (const std::exception& e) {
    die(e.what());
  }
  g_ctx   = w;
  g_ev    = new Evaluator(w, g_primary_ev);
  g_queue = new Scheduler(w);
  g_queue->eager = g_eager;
  g_workers.push_back(Worker{g_ctx, g_ev, g_queue});
}
Context* ctx_nf() {
  if (!g_ctx) {
    if (!g_primary) die("context not prepared (call Prepare_context first)");
    attach_worker();
  }
  return g_ctx;
}
Context* ctx() {
  Context* c = ctx_nf();
  if (g_queue && !g_queue->empty()) g_queue->flush();
  return c;
}

inline u64*       U(int64_t* p) { return reinterpret_cast<u64*>(p); }
inline const u64* U(const int64_t* p) { return reinterpret_cast<const u64*>(p); }

template <typename F>
void guard(F&& f) {
  try {
    f();
  } catch (const std::exception& e) {
    die(e.what());
  }
}

std::string io_key(const char* name, size_t idx) { return std::string(name) + "#" + std::to_string(idx); }

// Alloc_poly_data (polynomial.h:54-64): zero-filled limbs in HBM
void alloc_poly_data(POLYNOMIAL* p, uint32_t degree, size_t nq, size_t np) {
  StatScope ss(ST_ALLOC);
  p->_ring_degree      = degree;
  p->_num_primes       = nq;
  p->_num_primes_p     = np;
  p->_num_alloc_primes = nq + np;
  p->_is_ntt           = false;
  // fresh memory: nothing recorded can refer to it, no flush needed
  guard([&] { ctx_nf(); p->_data = reinterpret_cast<int64_t*>(g_queue->alloc(nq + np, true)); });
}

void free_poly_data(POLYNOMIAL* p) {
  StatScope ss(ST_FREE);
  if (p->_data) {
    guard([&] { ctx_nf(); g_queue->free(U(p->_data)); });
    p->_data = nullptr;
  }
  p->_num_alloc_primes = 0;
}

size_t poly_len(const POLYNOMIAL* p) { return (p->_num_primes + p->_num_primes_p) * (size_t)p->_ring_degree; }

// Init_poly (polynomial.h:335-347): size like `poly`, contents zeroed
void init_poly(POLYNOMIAL* res, const POLYNOMIAL* poly) {
  if (res->_data == nullptr) {
    alloc_poly_data(res, poly->_ring_degree, poly->_num_primes, poly->_num_primes_p);
  } else if (res->_num_alloc_primes * (size_t)res->_ring_degree < poly_len(poly)) {
    free_poly_data(res);
    alloc_poly_data(res, poly->_ring_degree, poly->_num_primes, poly->_num_primes_p);
  } else {
    guard([&] { ctx_nf(); g_queue->zero(U(res->_data), res->_num_alloc_primes); });
    res->_ring_degree  = poly->_ring_degree;
    res->_num_primes   = poly->_num_primes;
    res->_num_primes_p = poly->_num_primes_p;
    res->_is_ntt       = false;
  }
}

int64_t* p_coeffs(const POLYNOMIAL* p) {  // Get_p_coeffs (polynomial.h:214-217)
  return p->_data + (p->_num_alloc_primes - p->_num_primes_p) * (size_t)p->_ring_degree;
}

void copy_polynomial(POLYNOMIAL* dst, const POLYNOMIAL* src) {  // polynomial.h:425-438
  StatScope ss(ST_COPY);
  guard([&] {
    ctx_nf();
    g_queue->copy(U(dst->_data), U(src->_data), dst->_num_primes);
    if (src->_num_primes_p) g_queue->copy(U(p_coeffs(dst)), U(p_coeffs(src)), dst->_num_primes_p);
  });
  dst->_is_ntt = src->_is_ntt;
}

// Init_ciphertext_from_ciph (ciphertext.h:211-221)
void init_ct_from_ct(CIPHERTEXT* res, CIPHERTEXT* ciph, double sf, uint32_t sf_degree) {
  res->_scaling_factor = sf;
  res->_sf_degree      = sf_degree;
  res->_slots          = ciph->_slots;
  if (res == ciph) return;
  init_poly(&res->_c0_poly, &ciph->_c0_poly);
  init_poly(&res->_c1_poly, &ciph->_c1_poly);
}

void set_level(CIPHERTEXT* c, size_t level) {
  c->_c0_poly._num_primes = level;
  c->_c1_poly._num_primes = level;
}

// Adjust_level with resize = false (ciphertext.h:283-319): the operand with fewer limbs
CIPHERTEXT* lower_level(CIPHERTEXT* a, CIPHERTEXT* b) {
  if (a->_c0_poly._data == nullptr) return b;
  if (b->_c0_poly._data == nullptr) die("poly coeffs of input ciph is invalid");
  return a->_c0_poly._num_primes > b->_c0_poly._num_primes ? b : a;
}

void init_cipher(CIPHERTEXT* res, CIPHERTEXT* ciph, double sc, uint32_t deg) {  // cipher_eval.c:23-30
  init_ct_from_ct(res, ciph, sc, deg);
  res->_c0_poly._is_ntt = res->_c1_poly._is_ntt = true;
  set_level(res, ciph->_c0_poly._num_primes);
}

// Bootstrap_keygen (ckks_bootstrap_context.c:1194-1226): rotation keys of the linear
// transforms plus the conjugation key (automorphism index 2N-1)
void bootstrap_keygen(u32 slots) {
  Context* c = g_ctx;
  std::vector<int32_t> rots = g_ev->bootstrap_rot_indices(slots);
  for (size_t i = 0; i < rots.size(); i++) {
    u32 k = c->auto_index(rots[i]);
    if (!c->has_rot_key(k)) c->gen_auto_key(k);
  }
  u32 conj = 2 * c->N - 1;
  if (!c->has_rot_key(conj)) c->gen_auto_key(conj);
}

u32 mod_index(MODULUS* m) {
  ptrdiff_t g = m - g_mod;
  if (g < 0 || (size_t)g >= ctx_nf()->G) die("MODULUS pointer does not belong to the context");
  return (u32)g;
}

SWITCH_KEY* wrap_key(u32 slot, SwitchKey* k) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_swk.find(slot);
  if (it != g_swk.end()) return it->second;
  Context* c = ctx_nf();
  SWITCH_KEY* s = new SWITCH_KEY;
  s->key = k;
  s->pk0 = new POLYNOMIAL[c->dnum];
  s->pk1 = new POLYNOMIAL[c->dnum];
  for (size_t j = 0; j < c->dnum; j++) {
    for (int w = 0; w < 2; w++) {
      POLYNOMIAL* p = w ? &s->pk1[j] : &s->pk0[j];
      p->_ring_degree = c->N; p->_num_primes = c->L; p->_num_primes_p = c->K;
      p->_num_alloc_primes = c->G; p->_is_ntt = true;
      p->_data = reinterpret_cast<int64_t*>((w ? k->k1 : k->k0) + j * c->G * (size_t)c->N);
    }
  }
  g_swk[slot] = s;
  return s;
}

// ---- weight file (DE_MSG_F32/F64): rt_data_def.h:90-109, pt_mgr.c:35-110 -----------------
struct DataFileHdr {
  char            magic[8];  // "!ANTFHE\0"
  uint32_t        rt_ver;
  uint16_t        flag;
  uint8_t         ent_type;   // DATA_ENTRY_TYPE
  uint8_t         ent_align;  // log2 of the entry alignment
  uint64_t        ent_count;
  uint64_t        lut_ofst;
  struct timespec ctime;
  char            model[48];
  char            uuid[40];
};
struct LutEntry {
  char     name[16];
  uint32_t index;
  uint32_t size;
  uint64_t ent_ofst;
};
std::vector<char> g_wfile;
// per host thread: plaintexts already encoded on this thread's stream, by (entry, level, degree)
thread_local std::unordered_map<u64, u64*> g_pt_cache;
thread_local size_t                        g_pt_cache_bytes = 0;
static size_t pt_cache_limit() {
  static const size_t lim = [] {
    const char* e = getenv("ACE_B200_PT_CACHE_GB");
    return (size_t)(e ? atof(e) : 24.0) << 30;
  }();
  return lim;
}
// the cache only takes memory nobody needs: it stops growing for good once less than 40 GB of the
// device would stay free (keys and the working sets of the images in flight come first)
thread_local bool g_pt_cache_full = false;
static bool pt_cache_has_room(size_t bytes) {
  if (g_pt_cache_full) return false;
  static thread_local unsigned calls = 0;
  if ((calls++ & 63) == 0) {  // cudaMemGetInfo is not free: look every 64th insertion
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess || fr < bytes + ((size_t)40 << 30)) g_pt_cache_full = true;
  }
  return !g_pt_cache_full;
}
// cache memory comes in 1 GiB slabs (6 044 cudaMalloc calls cost 17 s on the first image)
thread_local std::vector<char*> g_pt_slabs;
thread_local size_t             g_pt_slab_left = 0;
static u64* pt_cache_take(size_t bytes) {
  constexpr size_t kSlab = (size_t)1 << 30;
  if (bytes > kSlab) return nullptr;
  if (g_pt_slab_left < bytes) {
    char* p = nullptr;
    if (cudaMalloc(&p, kSlab) != cudaSuccess) {
      cudaGetLastError();
      g_pt_cache_full = true;
      return nullptr;
    }
    g_pt_slabs.push_back(p);
    g_pt_slab_left = kSlab;
  }
  char* at = g_pt_slabs.back() + (((size_t)1 << 30) - g_pt_slab_left);
  g_pt_slab_left -= bytes;
  return reinterpret_cast<u64*>(at);
}
char*             g_wfile_dev = nullptr;  // the same bytes in HBM: Pt_from_msg encodes from there
const LutEntry*   g_lut   = nullptr;
uint64_t          g_nent  = 0;
uint32_t          g_etype = 0;

}  // namespace

// =========================================================================== context
API void Ace_set_device(int device) { g_device = device; }
API void* Ace_context(void) { return g_ctx; }

// ---- B200 extensions used by bench.py / tests: device-side timing and counters
static thread_local cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
API void Ace_timer_start(void) {
  Context* c = ctx();
  if (!g_ev0) { cudaEventCreate(&g_ev0); cudaEventCreate(&g_ev1); }
  cudaEventRecord(g_ev0, c->stream);
}
API float Ace_timer_stop_ms(void) {
  Context* c = ctx();
  float ms = 0;
  cudaEventRecord(g_ev1, c->stream);
  cudaEventSynchronize(g_ev1);
  cudaEventElapsedTime(&ms, g_ev0, g_ev1);
  return ms;
}
API uint64_t Ace_launch_count(void) { return ctx()->launches; }
// rotation indices whose keys Bootstrap_keygen would generate for `slots` (0 = N/2); the
// conjugation key is rotation index 2N-1.  Returns the count.
API int Ace_bootstrap_rot_indices(uint32_t slots, int32_t* out, size_t cap) {
  ctx();
  if (!g_ev->bootstrap_supported()) return 0;
  std::vector<int32_t> v = g_ev->bootstrap_rot_indices(slots);
  for (size_t i = 0; i < v.size() && i < cap; i++) out[i] = v[i];
  return (int)v.size();
}
// TEST SUPPORT (whole-model parity against the golden runs, tests/model_case.py): after a
// Prepare_context with ACE_B200_NO_KEYGEN=1, generate the key set the REFERENCE's Prepare_context
// generates from the same pinned random streams (context.c:56-82: Alloc_ckks_key_generator, then
// Bootstrap_precom(N/2) -> Bootstrap_keygen: rotation keys by VALUE, then the conjugation key).
// Later encryptions (Prepare_input) continue the same streams.
API int Ace_keygen_reference_stream(const uint32_t* seed16, uint64_t counter, uint32_t srandom_seed,
                                    const uint64_t* tri_pos, size_t n_pos) {
  Context* c = ctx();
  if (c != g_ctx) die("Ace_keygen_reference_stream: call it from the thread that prepared the context");
  CKKS_PARAMS* p = Get_context_params();
  guard([&] {
    c->keygen_reference_stream(seed16, counter, srandom_seed, tri_pos, n_pos, p->_rot_idxs, p->_num_rot_idx);
    if (g_ev->bootstrap_supported()) {
      std::vector<int32_t> rots = g_ev->bootstrap_rot_indices(c->N / 2);
      c->keygen_rotations(rots.data(), rots.size());
      c->gen_auto_key(2 * c->N - 1);
    }
    c->sync();
  });
  return 0;
}
// op trace so far: out[class * 72 + level], classes in the order of Context::TraceClass
// (modup digit, moddown poly, rescale poly, encode, limb mul, limb add, limb rotate, limb ntt)
API int Ace_trace(uint64_t* out, size_t cap) {
  Context* c = ctx();
  size_t n = (size_t)Context::TR_CLASSES * Context::kTraceLevels;
  if (cap < n) return -1;
  memcpy(out, c->trace, n * sizeof(uint64_t));
  return (int)n;
}

API void Prepare_context(void) {
  if (g_primary) return;
  g_stats_on   = getenv("ACE_B200_STATS") && getenv("ACE_B200_STATS")[0] >= '1';
  g_stats_sync = g_stats_on && getenv("ACE_B200_STATS")[0] == '2';
  prof::enable(getenv("ACE_B200_PROF") && getenv("ACE_B200_PROF")[0] == '1');
  if (!Get_context_params) die("Get_context_params() not linked (emitted unit missing)");
  CKKS_PARAMS* p = Get_context_params();
  size_t parts = p->_num_q_parts;
  if (parts == 0) parts = p->_mul_depth > 3 ? 3 : (p->_mul_depth == 0 ? 1 : 2);  // fhe_std_parms.c:327-334
  guard([&] {
    Params prm{p->_poly_degree, p->_mul_depth, p->_first_mod_size, p->_scaling_mod_size, parts,
               p->_hamming_weight};
    g_ctx = g_primary = new Context(prm, g_device);
    g_mod = new MODULUS[g_ctx->G];
    for (size_t g = 0; g < g_ctx->G; g++) {
      g_mod[g]._val = (int64_t)g_ctx->mod[g];
      g_mod[g]._br_k = g_mod[g]._br_m = 0;
      g_mod[g]._prec128 = ~(unsigned __int128)0 / g_ctx->mod[g];
    }
    if (!quiet()) printf("ckks_param: _provider = %d, _poly_degree = %d, _sec_level = %ld, mul_depth = %ld, "
           "_first_mod_size = %ld, _scaling_mod_size = %ld, _num_q_parts = %ld, _num_p = %ld, "
           "_num_rot_idx = %ld,_hamming_wieght = %ld\n",
           p->_provider, p->_poly_degree, p->_sec_level, p->_mul_depth, p->_first_mod_size,
           p->_scaling_mod_size, parts, g_ctx->K, p->_num_rot_idx, p->_hamming_weight);
    g_ev = g_primary_ev = new Evaluator(g_ctx);
    g_queue = new Scheduler(g_ctx);
    // ACE_B200_EAGER=1: issue every call at once (no batching across calls)
    g_eager = getenv("ACE_B200_EAGER") && getenv("ACE_B200_EAGER")[0] == '1';
    g_queue->eager = g_eager;
    const char* no_keys = getenv("ACE_B200_NO_KEYGEN");  // parity runs import the oracle's keys
    const bool  own_keys = !(no_keys && no_keys[0] == '1');
    // keys come from the operating system's entropy (getrandom -> ChaCha20, client.cu), as the
    // reference seeds its PRNG from /dev/urandom and the clock (prng.c:28-84).  ACE_B200_SEED
    // pins a reproducible stream: TESTS ONLY -- whoever knows the seed can regenerate the keys.
    const char* seed_env = getenv("ACE_B200_SEED");
    const u64   seed = seed_env ? strtoull(seed_env, nullptr, 10) : 0;
    if (own_keys) g_ctx->keygen(seed, p->_rot_idxs, p->_num_rot_idx);
    // Bootstrap_precom(N/2) (context.c:80-82, 162-185): plaintext tables + bootstrap keys
    if (g_ev->bootstrap_supported()) {
      g_ev->bootstrap_setup(g_ctx->N / 2);
      if (own_keys) bootstrap_keygen(g_ctx->N / 2);
    }
  });
  if (Get_rt_data_info) {
    RT_DATA_INFO* info = Get_rt_data_info();
    if (info != nullptr) Pt_mgr_init(info->_file_name);
  }
}

API void Finalize_context(void) {  // called by the thread that prepared the context
  if (!g_primary) return;
  if (g_stats_on) {
    printf("[ace_b200 stats] kernels launched: %zu; scheduler: %zu ops in %zu flushes / %zu waves, "
           "%zu chain launches, %zu mul+add fused, %zu dead stores dropped, %zu of %zu Decomp_modup "
           "served from an earlier result; peak limb memory %.1f GB; %zu worker thread(s)\n",
           g_ctx->launches, g_queue->n_ops, g_queue->n_flush, g_queue->n_waves,
           g_queue->n_chain_launches, g_queue->n_fused, g_queue->n_dead, g_queue->n_modup_shared,
           g_queue->n_modup,
           g_ctx->peak_bytes / 1073741824.0, g_workers.size());
    for (int i = 0; i < ST_COUNT; i++)
      printf("[ace_b200 stats] %-22s calls %9llu  host time %8.3f s\n", g_stats[i].name,
             (unsigned long long)g_stats[i].calls, g_stats[i].secs);
  }
  {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& kv : g_swk) { delete[] kv.second->pk0; delete[] kv.second->pk1; delete kv.second; }
    g_swk.clear();
    for (Worker& w : g_workers) {  // their threads must have finished their images
      w.q->flush();
      delete w.q;
      delete w.ev;
      delete w.c;
    }
    g_workers.clear();
  }
  Pt_mgr_fini();
  if (g_queue) g_queue->flush();
  delete g_queue;
  g_queue = nullptr;
  delete g_primary_ev;
  g_ev = g_primary_ev = nullptr;
  delete g_primary;
  g_ctx = g_primary = nullptr;
  delete[] g_mod;
  g_mod = nullptr;
}

API uint32_t Degree(void) { return ctx_nf()->N; }
API double   Get_default_sc(void) { return (double)((u64)1 << ctx_nf()->params.scaling_mod_size); }
API size_t   Get_q_parts(void) { return ctx_nf()->dnum; }
API size_t   Get_p_cnt(void) { return ctx_nf()->K; }
API MODULUS* Q_modulus(void) { ctx_nf(); return g_mod; }
API MODULUS* P_modulus(void) { return g_mod + ctx_nf()->L; }

// =========================================================================== tensors / IO
API TENSOR* Alloc_tensor(size_t n, size_t c, size_t h, size_t w, const double* val) {
  size_t  cnt = n * c * h * w;
  TENSOR* t   = (TENSOR*)malloc(sizeof(TENSOR) + cnt * sizeof(double));
  t->_shape = SHAPE{n, c, h, w};
  if (val) memcpy(t->_vals, val, cnt * sizeof(double));
  else memset(t->_vals, 0, cnt * sizeof(double));
  return t;
}
API void Free_tensor(TENSOR* t) { free(t); }
API void Print_tensor(FILE* fp, TENSOR* t) {
  fprintf(fp, "\n[%zu %zu %zu %zu]: ", t->_shape._n, t->_shape._c, t->_shape._h, t->_shape._w);
  size_t cnt = TENSOR_SIZE(t);
  for (size_t i = 0; i < cnt && i < 16; i++) fprintf(fp, "%f ", t->_vals[i]);
  fprintf(fp, "\n");
}

// Prepare_input (rtlib.c:41-54): encode at full level with default slots, then pk-encrypt
API void Prepare_input(TENSOR* input, const char* name) {
  Context* c = ctx();
  size_t   len = TENSOR_SIZE(input);
  CIPHERTEXT* ct = (CIPHERTEXT*)calloc(1, sizeof(CIPHERTEXT));
  guard([&] {
    u64* pt = c->alloc_limbs(c->L, false);
    c->encode(pt, input->_vals, len, (u32)c->L, 0, 1, 0);
    alloc_poly_data(&ct->_c0_poly, c->N, c->L, 0);
    alloc_poly_data(&ct->_c1_poly, c->N, c->L, 0);
    ctx();  // the recorded zero fills go first
    c->encrypt(U(ct->_c0_poly._data), U(ct->_c1_poly._data), pt, (u32)c->L, g_enc_id.fetch_add(1));
    c->free_limbs(pt);
  });
  ct->_c0_poly._is_ntt = ct->_c1_poly._is_ntt = true;
  ct->_slots          = c->N / 2;
  ct->_scaling_factor = Get_default_sc();
  ct->_sf_degree      = 1;
  g_inputs[io_key(name, 0)] = ct;
}

API void Ace_set_input(const char* name, size_t idx, const int64_t* c0, const int64_t* c1,
                       uint32_t level, uint32_t slots, double scale, uint32_t sf_degree) {
  Context* c = ctx();
  CIPHERTEXT* ct = (CIPHERTEXT*)calloc(1, sizeof(CIPHERTEXT));
  alloc_poly_data(&ct->_c0_poly, c->N, level, 0);
  alloc_poly_data(&ct->_c1_poly, c->N, level, 0);
  ctx();  // the recorded zero fills go first
  guard([&] {
    c->upload(U(ct->_c0_poly._data), U(c0), level);
    c->upload(U(ct->_c1_poly._data), U(c1), level);
    c->sync();
  });
  ct->_c0_poly._is_ntt = ct->_c1_poly._is_ntt = true;
  ct->_slots = slots; ct->_scaling_factor = scale; ct->_sf_degree = sf_degree;
  g_inputs[io_key(name, idx)] = ct;
}

API CIPHER Ace_peek_input(const char* name, size_t idx) {  // test support: look, do not take
  auto it = g_inputs.find(io_key(name, idx));
  return it == g_inputs.end() ? nullptr : it->second;
}

API CIPHERTEXT Get_input_data(const char* name, size_t idx) {  // rtlib.c:74-80
  auto it = g_inputs.find(io_key(name, idx));
  if (it == g_inputs.end()) die("not find data");
  CIPHERTEXT ret = *it->second;
  free(it->second);
  g_inputs.erase(it);
  return ret;
}

API void Set_output_data(const char* name, size_t idx, CIPHER data) {  // rtlib.c:82-87
  CIPHERTEXT* out = (CIPHERTEXT*)calloc(1, sizeof(CIPHERTEXT));
  Copy_ciph(out, data);
  Free_ciph_poly(data, 1);
  auto key = io_key(name, idx);
  auto it  = g_outputs.find(key);
  if (it != g_outputs.end()) { Free_ciph_poly(it->second, 1); free(it->second); }
  g_outputs[key] = out;
}

API CIPHER Ace_get_output(const char* name, size_t idx) {
  auto it = g_outputs.find(io_key(name, idx));
  return it == g_outputs.end() ? nullptr : it->second;
}

API double* Get_msg(CIPHER ciph) {  // cipher_eval.c:129-148 (non-extended ciphertexts)
  Context* c = ctx();
  u32 level = (u32)ciph->_c0_poly._num_primes, slots = ciph->_slots;
  double* out = (double*)malloc(sizeof(double) * slots);
  guard([&] {
    u64* pt = c->alloc_limbs(level, false);
    c->decrypt(pt, U(ciph->_c0_poly._data), U(ciph->_c1_poly._data), level);
    c->decode(out, nullptr, pt, level, slots, ciph->_scaling_factor);
    c->free_limbs(pt);
  });
  return out;
}

API double* Handle_output(const char* name) {  // rtlib.c:56-72
  CIPHER ct = Ace_get_output(name, 0);
  if (!ct) die("not find data");
  double* r = Get_msg(ct);
  Free_ciph_poly(ct, 1);
  free(ct);
  g_outputs.erase(io_key(name, 0));
  return r;
}

API void Print_cipher_msg(FILE* fp, const char* name, CIPHER ciph, uint32_t len) {
  double* d = Get_msg(ciph);
  fprintf(fp, "\n[%s] ciph_info: %d %d %ld %ld\n[%s] msg: [ ", name, ciph->_sf_degree,
          ciph->_slots, ciph->_c0_poly._num_primes, ciph->_c0_poly._num_primes_p, name);
  for (uint32_t i = 0; i < len && i < ciph->_slots; i++) fprintf(fp, "%.17f ", d[i]);
  fprintf(fp, "]\n");
  free(d);
}

API void Run_main_graph(void) {  // common/src/rt_lib.c:16-20
  if (!Main_graph) die("Main_graph() not linked");
  if (prof::on) { ctx(); prof::reset(); }
  if (!Main_graph()) die("Main_graph failed");
  guard([&] { ctx()->sync(); });
  if (prof::on) prof::report("emitted");
}

API void Tm_start(const char* msg) {
  (void)msg;
  if (quiet()) return;
  guard([&] { ctx()->sync(); });
  g_tm_stamp = clock();
}
API void Tm_taken(const char* msg) {  // rt_stat.c:23-28; a stream sync makes it device-accurate
  if (quiet()) return;
  guard([&] { ctx()->sync(); });
  clock_t cur = clock();
  fprintf(stdout, "[RT_STAT] %s takes %.3f seconds.\n", msg,
          ((double)(cur - g_tm_stamp)) / (double)CLOCKS_PER_SEC);
  g_tm_stamp = cur;
}

// =========================================================================== polynomials
API POLY Alloc_poly(uint32_t degree, size_t q_primes, bool extend_p) {  // poly_eval.h:29-37
  if (q_primes == 0) die("Alloc_poly: q primes should not be NULL");
  POLY p = (POLY)calloc(1, sizeof(POLYNOMIAL));
  alloc_poly_data(p, degree, q_primes, extend_p ? ctx_nf()->K : 0);
  p->_is_ntt = true;
  return p;
}
API void Free_poly_data(POLY poly) { free_poly_data(poly); }
API void Free_poly(POLY poly) {
  free_poly_data(poly);
  free(poly);
}
API void Copy_poly(POLY res, POLY poly) { copy_polynomial(res, poly); }
API void Set_coeffs(POLY dst, uint32_t level, uint32_t degree, int64_t* src) {  // poly_eval.h:74-79
  StatScope ss(ST_COPY);
  guard([&] { ctx_nf(); g_queue->copy(U(dst->_data + (size_t)level * degree), U(src), 1); });
}
API size_t Num_decomp(POLY poly) { return ctx_nf()->num_decomp(poly->_num_primes); }

API void Ace_download_poly(int64_t* host_dst, POLY poly) {
  guard([&] { ctx()->download(U(host_dst), U(poly->_data), poly->_num_primes + poly->_num_primes_p); });
}
API void Ace_upload_poly(POLY poly, const int64_t* host_src) {
  guard([&] {
    ctx()->upload(U(poly->_data), U(host_src), poly->_num_primes + poly->_num_primes_p);
    ctx()->sync();
  });
}

// Hw_modadd / Hw_modmul / Hw_rotate (poly_arith.c:14-56) are recorded and issued in batches
// by the scheduler (sched.h); ACE_B200_EAGER=1 issues every call at once instead.
API int64_t* Hw_modadd(int64_t* res, int64_t* a, int64_t* b, MODULUS* m, uint32_t degree) {
  StatScope ss(ST_ADD);
  ctx_nf();
  guard([&] { g_queue->ew(OP_ADD, U(res), U(a), U(b), mod_index(m)); });
  return res + degree;
}
API int64_t* Hw_modmul(int64_t* res, int64_t* a, int64_t* b, MODULUS* m, uint32_t degree) {
  StatScope ss(ST_MUL);
  ctx_nf();
  guard([&] { g_queue->ew(OP_MUL, U(res), U(a), U(b), mod_index(m)); });
  return res + degree;
}
API int64_t* Hw_rotate(int64_t* res, int64_t* a, int64_t* order, MODULUS* m, uint32_t degree) {
  StatScope ss(ST_ROT);
  ctx_nf();
  guard([&] { g_queue->gather(U(res), U(a), order, mod_index(m)); });
  return res + degree;
}

API POLY Decomp(POLY res, POLY poly, uint32_t part) {  // Decompose_poly, polynomial.c:848-875
  Context* c = ctx();
  u32 nq = (u32)poly->_num_primes, len = c->digit_len(nq, part), st = c->digit_start(part);
  if (res->_num_alloc_primes < len) {
    free_poly_data(res);
    alloc_poly_data(res, poly->_ring_degree, len, 0);
  } else {
    res->_num_primes = len;
    res->_num_primes_p = 0;
  }
  guard([&] { g_queue->copy(U(res->_data), U(poly->_data + (size_t)st * c->N), len); });
  res->_is_ntt = poly->_is_ntt;
  return res;
}
// The kernels take extended polynomials as [num_q Q limbs | K P limbs] back to back.  The
// reference keeps the P part at _num_alloc_primes - _num_primes_p (Get_p_coeffs,
// polynomial.h:214-217): the two agree unless a polynomial allocated at a higher level is reused
// at a lower one.  No emitted program does that with an extended polynomial; refuse it loudly
// instead of reading the wrong limbs.
static void check_ext_layout(const char* who, POLY p) {
  if (p->_num_primes_p != 0 && p->_num_alloc_primes != p->_num_primes + p->_num_primes_p)
    die((std::string(who) + ": extended polynomial reused at a lower level (P limbs are not "
         "adjacent to the Q limbs); not supported").c_str());
}
API POLY Mod_up(POLY new_poly, POLY old_poly, uint32_t part) {  // poly_eval.c:19-26
  check_ext_layout("Mod_up", new_poly);
  guard([&] { ctx()->modup_from(U(new_poly->_data), U(old_poly->_data), (u32)new_poly->_num_primes, part); });
  new_poly->_is_ntt = old_poly->_is_ntt;
  return new_poly;
}
API POLY Decomp_modup(POLY res, POLY poly, uint32_t part) {  // poly_eval.c:28-34
  StatScope ss(ST_MODUP);
  check_ext_layout("Decomp_modup", res);
  guard([&] { ctx_nf(); g_queue->modup(U(res->_data), U(poly->_data), (u32)poly->_num_primes, part); });
  res->_is_ntt = true;
  return res;
}
API POLY Mod_down(POLY res, POLY poly) {  // poly_eval.c:36-41
  StatScope ss(ST_MODDOWN);
  check_ext_layout("Mod_down", poly);
  guard([&] { ctx_nf(); g_queue->moddown(U(res->_data), U(poly->_data), (u32)res->_num_primes); });
  res->_is_ntt = poly->_is_ntt;
  return res;
}
API POLY Rescale(POLY res, POLY poly) {  // poly_eval.c:43-49
  StatScope ss(ST_RESCALE);
  guard([&] { ctx_nf(); g_queue->rescale(U(res->_data), U(poly->_data), (u32)poly->_num_primes); });
  res->_is_ntt     = true;
  res->_num_primes = res->_num_primes - 1;  // Mod_down_q_primes
  return res;
}

// =========================================================================== keys
API uint32_t Auto_idx(int32_t rot_idx) { return ctx_nf()->auto_index(rot_idx); }
API int64_t* Auto_order(int32_t rot_idx) {
  int64_t* r = nullptr;
  guard([&] { r = const_cast<int64_t*>(ctx_nf()->auto_order(ctx_nf()->auto_index(rot_idx))); });
  return r;
}
API SW_KEY Swk(bool is_rot, int32_t rot_idx) {
  Context* c = ctx_nf();
  if (!is_rot) {
    if (!c->relin_key.k0) die("relinearisation key missing");
    return wrap_key(0, &c->relin_key);
  }
  u32 k = c->auto_index(rot_idx);
  if (!c->has_rot_key(k)) die("cannot find auto key");
  return wrap_key(k, &c->rot_key(k));
}
API POLY Pk0_at(SW_KEY swk, uint32_t idx) { return &swk->pk0[idx]; }
API POLY Pk1_at(SW_KEY swk, uint32_t idx) { return &swk->pk1[idx]; }

API void Ace_import_switch_key(bool is_rot, int32_t rot_idx, uint32_t part, int which,
                               const int64_t* host_poly) {
  guard([&] {
    Context* c = ctx();
    SwitchKey& k = is_rot ? c->rot_key(c->auto_index(rot_idx)) : c->relin_key;
    c->import_key_limbs(k, part, which, U(host_poly));
  });
}

// =========================================================================== ciphertexts
static void dbg_range(const char* where, CIPHER ct);
API void Init_ciph_same_scale(CIPHER res, CIPHER c1, CIPHER c2) {  // cipher_eval.c:32-43
  CIPHER ciph = c2 != nullptr ? lower_level(c1, c2) : c1;
  res->_scaling_factor = ciph->_scaling_factor;
  res->_sf_degree      = ciph->_sf_degree;
  res->_slots          = ciph->_slots;
  size_t level = ciph->_c0_poly._num_primes, np = ciph->_c0_poly._num_primes_p;
  if (res->_c0_poly._data == nullptr) alloc_poly_data(&res->_c0_poly, ciph->_c0_poly._ring_degree, level, np);
  if (res->_c1_poly._data == nullptr) alloc_poly_data(&res->_c1_poly, ciph->_c0_poly._ring_degree, level, np);
  res->_c0_poly._is_ntt = res->_c1_poly._is_ntt = true;
}
API void Init_ciph_same_scale_plain(CIPHER res, CIPHER ciph, PLAIN plain) {
  (void)plain;
  init_cipher(res, ciph, ciph->_scaling_factor, ciph->_sf_degree);
}
API void Init_ciph_up_scale(CIPHER res, CIPHER c1, CIPHER c2) {
  CIPHER ciph = lower_level(c1, c2);
  init_cipher(res, ciph, c1->_scaling_factor * c2->_scaling_factor, c1->_sf_degree + c2->_sf_degree);
}
API void Init_ciph_up_scale_plain(CIPHER res, CIPHER ciph, PLAIN plain) {
  dbg_range("mul_plain", ciph);
  init_cipher(res, ciph, ciph->_scaling_factor * plain->_scaling_factor,
              ciph->_sf_degree + plain->_sf_degree);
}
// ACE_B200_DEBUG_RANGE=<n> (own keys): decrypt the operand of the first n rescales / plaintext
// multiplications and print its message range -- a trace of activation magnitudes
static void dbg_range(const char* where, CIPHER ct) {
  static int left = getenv("ACE_B200_DEBUG_RANGE") ? atoi(getenv("ACE_B200_DEBUG_RANGE")) : 0;
  if (left <= 0 || ct->_c0_poly._data == nullptr || ct->_c0_poly._num_primes_p) return;
  left--;
  double* m = Get_msg(ct);
  double mx = 0;
  for (uint32_t i = 0; i < ct->_slots; i++) mx = std::max(mx, fabs(m[i]));
  printf("[ace_b200 range] %-12s level %2zu sf_degree %u scale 2^%.2f  max|msg| %.5g\n", where,
         (size_t)ct->_c0_poly._num_primes, ct->_sf_degree, log2(ct->_scaling_factor), mx);
  free(m);
}

API void Init_ciph_down_scale(CIPHER res, CIPHER ciph) {
  dbg_range("rescale", ciph);
  init_cipher(res, ciph, ciph->_scaling_factor / Get_default_sc(), ciph->_sf_degree - 1);
}
API void Init_ciph_same_scale_ciph3(CIPHER res, CIPHER3 ciph) {  // cipher_eval.c:65-75
  res->_scaling_factor = ciph->_scaling_factor;
  res->_sf_degree      = ciph->_sf_degree;
  res->_slots          = ciph->_slots;
  init_poly(&res->_c0_poly, &ciph->_c0_poly);
  init_poly(&res->_c1_poly, &ciph->_c1_poly);
  res->_c0_poly._is_ntt = res->_c1_poly._is_ntt = true;
  set_level(res, ciph->_c0_poly._num_primes);
}
static void init_ct3(CIPHER3 res, POLYNOMIAL* like, double sf, uint32_t deg, uint32_t slots) {
  res->_scaling_factor = sf;
  res->_sf_degree      = deg;
  res->_slots          = slots;
  init_poly(&res->_c0_poly, like);
  init_poly(&res->_c1_poly, like);
  init_poly(&res->_c2_poly, like);
  res->_c0_poly._is_ntt = res->_c1_poly._is_ntt = res->_c2_poly._is_ntt = true;
  res->_c0_poly._num_primes = res->_c1_poly._num_primes = res->_c2_poly._num_primes = like->_num_primes;
}
API void Init_ciph3_same_scale_ciph3(CIPHER3 res, CIPHER3 c1, CIPHER3 c2) {
  CIPHER3 ciph = c1;
  if (c2 != nullptr && c1->_c0_poly._data != nullptr &&
      c1->_c0_poly._num_primes > c2->_c0_poly._num_primes) ciph = c2;
  if (c2 != nullptr && c1->_c0_poly._data == nullptr) ciph = c2;
  if (res == ciph) return;
  init_ct3(res, &ciph->_c0_poly, ciph->_scaling_factor, ciph->_sf_degree, ciph->_slots);
}
API void Init_ciph3_up_scale(CIPHER3 res, CIPHER c1, CIPHER c2) {  // cipher_eval.c:93-107
  CIPHER ciph = lower_level(c1, c2);
  init_ct3(res, &ciph->_c0_poly, c1->_scaling_factor * c2->_scaling_factor,
           c1->_sf_degree + c2->_sf_degree, ciph->_slots);
}
API void Copy_ciph(CIPHER res, CIPHER ciph) {  // ciphertext.h:236-241
  init_ct_from_ct(res, ciph, ciph->_scaling_factor, ciph->_sf_degree);
  if (res == ciph) return;
  copy_polynomial(&res->_c0_poly, &ciph->_c0_poly);
  copy_polynomial(&res->_c1_poly, &ciph->_c1_poly);
}
API void Zero_ciph(CIPHER ciph) {
  free_poly_data(&ciph->_c0_poly);
  free_poly_data(&ciph->_c1_poly);
  memset(ciph, 0, sizeof(*ciph));
}
API void Free_ciph_poly(CIPHER ciph, uint32_t cnt) {
  for (uint32_t i = 0; i < cnt; i++) {
    free_poly_data(&ciph[i]._c0_poly);
    free_poly_data(&ciph[i]._c1_poly);
  }
}
API size_t   Level(CIPHER ciph) { return ciph->_c0_poly._num_primes; }
API uint32_t Sc_degree(CIPHER ciph) { return ciph->_sf_degree; }
API uint32_t Get_slots(CIPHER ciph) { return ciph->_slots; }
API void     Set_slots(CIPHER ciph, uint32_t slots) { ciph->_slots = slots; }

// ---- CKKS-level API (cipher_eval.c:292-364 -> ckks_evaluator.c), non-extended ciphertexts
static void ct_binary(EwOp op, CIPHER res, CIPHER a, CIPHER b) {
  CIPHER low = lower_level(a, b);
  u32 level = (u32)low->_c0_poly._num_primes;
  if (res != a && res != b) init_ct_from_ct(res, low, low->_scaling_factor, low->_sf_degree);
  Context* c = ctx();
  launch_ew(c->T, op, U(res->_c0_poly._data), U(a->_c0_poly._data), U(b->_c0_poly._data), 0, level, c->stream);
  launch_ew(c->T, op, U(res->_c1_poly._data), U(a->_c1_poly._data), U(b->_c1_poly._data), 0, level, c->stream);
  c->launches += 2;
  set_level(res, level);
  res->_c0_poly._is_ntt = res->_c1_poly._is_ntt = true;
}
API CIPHER Add_ciph(CIPHER res, CIPHER a, CIPHER b) { ct_binary(EW_ADD, res, a, b); return res; }
API CIPHER Sub_ciph(CIPHER res, CIPHER a, CIPHER b) { ct_binary(EW_SUB, res, a, b); return res; }

API CIPHER Add_plain(CIPHER res, CIPHER ciph, PLAIN plain) {  // ckks_evaluator.c:103-118
  init_ct_from_ct(res, ciph, ciph->_scaling_factor, ciph->_sf_degree);
  Context* c = ctx();
  u32 level = (u32)ciph->_c0_poly._num_primes;
  launch_ew(c->T, EW_ADD, U(res->_c0_poly._data), U(ciph->_c0_poly._data), U(plain->_poly._data), 0, level, c->stream);
  if (res != ciph) copy_polynomial(&res->_c1_poly, &ciph->_c1_poly);
  res->_c0_poly._is_ntt = true;
  c->launches++;
  return res;
}

API CIPHER Mul_plain(CIPHER res, CIPHER ciph, PLAIN plain) {  // ckks_evaluator.c:181-206
  double sf = ciph->_scaling_factor * plain->_scaling_factor;
  uint32_t deg = ciph->_sf_degree + plain->_sf_degree;
  init_ct_from_ct(res, ciph, sf, deg);
  Context* c = ctx();
  u32 level = (u32)ciph->_c0_poly._num_primes;
  launch_ew(c->T, EW_MUL, U(res->_c0_poly._data), U(ciph->_c0_poly._data), U(plain->_poly._data), 0, level, c->stream);
  launch_ew(c->T, EW_MUL, U(res->_c1_poly._data), U(ciph->_c1_poly._data), U(plain->_poly._data), 0, level, c->stream);
  res->_c0_poly._is_ntt = res->_c1_poly._is_ntt = true;
  c->launches += 2;
  return res;
}

API CIPHER3 Mul_ciph3(CIPHER3 res, CIPHER a, CIPHER b) {  // ckks_evaluator.c:133-165
  CIPHER low = lower_level(a, b);
  u32 level = (u32)low->_c0_poly._num_primes;
  POLYNOMIAL like = low->_c0_poly;
  init_ct3(res, &like, a->_scaling_factor * b->_scaling_factor, a->_sf_degree + b->_sf_degree, low->_slots);
  Context* c = ctx();
  guard([&] {
    u64* t = c->alloc_limbs(level, false);
    launch_ew(c->T, EW_MUL, U(res->_c0_poly._data), U(a->_c0_poly._data), U(b->_c0_poly._data), 0, level, c->stream);
    launch_ew(c->T, EW_MUL, U(res->_c1_poly._data), U(a->_c0_poly._data), U(b->_c1_poly._data), 0, level, c->stream);
    launch_ew(c->T, EW_MUL, t, U(a->_c1_poly._data), U(b->_c0_poly._data), 0, level, c->stream);
    launch_ew(c->T, EW_ADD, U(res->_c1_poly._data), U(res->_c1_poly._data), t, 0, level, c->stream);
    launch_ew(c->T, EW_MUL, U(res->_c2_poly._data), U(a->_c1_poly._data), U(b->_c1_poly._data), 0, level, c->stream);
    c->free_limbs(t);
    c->launches += 5;
  });
  return res;
}

API CIPHER Relin(CIPHER res, CIPHER3 ciph) {  // ckks_evaluator.c:266-282
  Context* c = ctx();
  u32 level = (u32)ciph->_c2_poly._num_primes;
  res->_scaling_factor = ciph->_scaling_factor;
  res->_sf_degree      = ciph->_sf_degree;
  res->_slots          = ciph->_slots;
  init_poly(&res->_c0_poly, &ciph->_c2_poly);
  init_poly(&res->_c1_poly, &ciph->_c2_poly);
  ctx();  // issue what was recorded for these blocks (zero fills) before launching directly
  guard([&] {
    u64* t = c->alloc_limbs(2 * (size_t)level, false);
    c->key_switch(t, t + (size_t)level * c->N, U(ciph->_c2_poly._data), level, c->relin_key, nullptr);
    launch_ew(c->T, EW_ADD, U(res->_c0_poly._data), t, U(ciph->_c0_poly._data), 0, level, c->stream);
    launch_ew(c->T, EW_ADD, U(res->_c1_poly._data), t + (size_t)level * c->N, U(ciph->_c1_poly._data), 0, level, c->stream);
    c->free_limbs(t);
    c->launches += 2;
  });
  res->_c0_poly._is_ntt = res->_c1_poly._is_ntt = true;
  return res;
}

API CIPHER Mul_ciph(CIPHER res, CIPHER a, CIPHER b) {  // ckks_evaluator.c:167-179
  CIPHER low = lower_level(a, b);
  u32 level = (u32)low->_c0_poly._num_primes;
  double sf = a->_scaling_factor * b->_scaling_factor;
  uint32_t deg = a->_sf_degree + b->_sf_degree, slots = low->_slots;
  Context* c = ctx();
  CIPHERTEXT tmp;
  memset(&tmp, 0, sizeof(tmp));
  alloc_poly_data(&tmp._c0_poly, c->N, level, 0);
  alloc_poly_data(&tmp._c1_poly, c->N, level, 0);
  ctx();  // issue what was recorded for these blocks (zero fills) before launching directly
  guard([&] {
    c->ct_mul_relin(U(tmp._c0_poly._data), U(tmp._c1_poly._data), U(a->_c0_poly._data),
                    U(a->_c1_poly._data), U(b->_c0_poly._data), U(b->_c1_poly._data), level);
  });
  free_poly_data(&res->_c0_poly);
  free_poly_data(&res->_c1_poly);
  res->_c0_poly = tmp._c0_poly;
  res->_c1_poly = tmp._c1_poly;
  res->_c0_poly._is_ntt = res->_c1_poly._is_ntt = true;
  res->_scaling_factor = sf; res->_sf_degree = deg; res->_slots = slots;
  return res;
}

API CIPHER Rescale_ciph(CIPHER res, CIPHER ciph) {  // ckks_evaluator.c:324-343
  Context* c = ctx();
  u32 level = (u32)ciph->_c0_poly._num_primes;
  if (level < 2) die("rescale: multiply level is not big enough");
  double sf = ciph->_scaling_factor / Get_default_sc();
  uint32_t deg = ciph->_sf_degree - 1, slots = ciph->_slots;
  CIPHERTEXT tmp;
  memset(&tmp, 0, sizeof(tmp));
  alloc_poly_data(&tmp._c0_poly, c->N, level, 0);
  alloc_poly_data(&tmp._c1_poly, c->N, level, 0);
  ctx();  // issue what was recorded for these blocks (zero fills) before launching directly
  guard([&] {
    c->rescale(U(tmp._c0_poly._data), U(ciph->_c0_poly._data), level);
    c->rescale(U(tmp._c1_poly._data), U(ciph->_c1_poly._data), level);
  });
  if (res != ciph) { free_poly_data(&res->_c0_poly); free_poly_data(&res->_c1_poly); }
  else { free_poly_data(&ciph->_c0_poly); free_poly_data(&ciph->_c1_poly); }
  res->_c0_poly = tmp._c0_poly;
  res->_c1_poly = tmp._c1_poly;
  set_level(res, level - 1);
  res->_c0_poly._is_ntt = res->_c1_poly._is_ntt = true;
  res->_scaling_factor = sf; res->_sf_degree = deg; res->_slots = slots;
  return res;
}

API CIPHER Rotate_ciph(CIPHER res, CIPHER ciph, int32_t rotation) {  // cipher_eval.c:353-364
  Context* c = ctx();
  u32 level = (u32)ciph->_c0_poly._num_primes;
  CIPHERTEXT tmp;
  memset(&tmp, 0, sizeof(tmp));
  alloc_poly_data(&tmp._c0_poly, c->N, level, 0);
  alloc_poly_data(&tmp._c1_poly, c->N, level, 0);
  ctx();  // issue what was recorded for these blocks (zero fills) before launching directly
  guard([&] {
    c->ct_rotate(U(tmp._c0_poly._data), U(tmp._c1_poly._data), U(ciph->_c0_poly._data),
                 U(ciph->_c1_poly._data), level, rotation);
  });
  double sf = ciph->_scaling_factor;
  uint32_t deg = ciph->_sf_degree, slots = ciph->_slots;
  free_poly_data(&res->_c0_poly);
  free_poly_data(&res->_c1_poly);
  res->_c0_poly = tmp._c0_poly;
  res->_c1_poly = tmp._c1_poly;
  res->_c0_poly._is_ntt = res->_c1_poly._is_ntt = true;
  res->_scaling_factor = sf; res->_sf_degree = deg; res->_slots = slots;
  return res;
}

API CIPHER Encrypt(CIPHER res, PLAIN plain) {  // cipher_eval.c:406-409
  Context* c = ctx();
  u32 level = (u32)plain->_poly._num_primes;
  res->_scaling_factor = plain->_scaling_factor;
  res->_sf_degree      = plain->_sf_degree;
  res->_slots          = plain->_slots;
  init_poly(&res->_c0_poly, &plain->_poly);
  init_poly(&res->_c1_poly, &plain->_poly);
  ctx();  // issue what was recorded for these blocks (zero fills) before launching directly
  guard([&] { c->encrypt(U(res->_c0_poly._data), U(res->_c1_poly._data), U(plain->_poly._data), level, g_enc_id.fetch_add(1)); });
  res->_c0_poly._is_ntt = res->_c1_poly._is_ntt = true;
  return res;
}

// Bootstrap (cipher_eval.c:366-404) -> Evaluator::bootstrap
API CIPHER Bootstrap(CIPHER res, CIPHER ciph, uint32_t level_after_bts) {
  Context* c = ctx();
  if (prof::on) prof::report("emitted");  // device time since the previous bootstrap
  // ACE_B200_DEBUG_BTS=1 (own keys only): decrypt before and after, print the message range and
  // how far the refreshed message is from the input -- tells a model whose activations leave the
  // bootstrap's (-1, 1) input range from a runtime problem
  static const bool dbg = getenv("ACE_B200_DEBUG_BTS") && getenv("ACE_B200_DEBUG_BTS")[0] == '1';
  double* before = dbg ? Get_msg(ciph) : nullptr;
  const uint32_t dbg_slots = ciph->_slots, dbg_level = (uint32_t)ciph->_c0_poly._num_primes;
  StatScope ss(ST_BTS);
  if (ciph->_c0_poly._num_primes_p != 0) die("Bootstrap: extended ciphertext");
  guard([&] {
    Ct in;
    in.c0 = U(ciph->_c0_poly._data); in.c1 = U(ciph->_c1_poly._data);
    in.nq = (u32)ciph->_c0_poly._num_primes; in.np = 0; in.cap = in.nq;
    in.sf = ciph->_scaling_factor; in.sfd = ciph->_sf_degree; in.slots = ciph->_slots;
    const bool own_keys = !(getenv("ACE_B200_NO_KEYGEN") && getenv("ACE_B200_NO_KEYGEN")[0] == '1');
    if (own_keys && in.slots != c->N / 2) bootstrap_keygen(in.slots);  // Bootstrap_precom(num_slots) on first use
    Ct out;
    g_ev->bootstrap(out, in, level_after_bts);
    if (res != ciph) { free_poly_data(&res->_c0_poly); free_poly_data(&res->_c1_poly); }
    else { free_poly_data(&ciph->_c0_poly); free_poly_data(&ciph->_c1_poly); }
    for (int w = 0; w < 2; w++) {
      POLYNOMIAL* p = w ? &res->_c1_poly : &res->_c0_poly;
      p->_ring_degree = c->N; p->_num_primes = out.nq; p->_num_primes_p = 0;
      p->_num_alloc_primes = out.cap; p->_is_ntt = true;
      p->_data = reinterpret_cast<int64_t*>(w ? out.c1 : out.c0);
    }
    res->_scaling_factor = out.sf; res->_sf_degree = out.sfd; res->_slots = out.slots;
  });
  if (prof::on) prof::report("bootstrap");
  if (dbg) {
    static int call = 0;
    double* after = Get_msg(res);
    double mx_in = 0, mx_out = 0, mx_diff = 0;
    for (uint32_t i = 0; i < dbg_slots; i++) {
      mx_in = std::max(mx_in, fabs(before[i]));
      mx_out = std::max(mx_out, fabs(after[i]));
      mx_diff = std::max(mx_diff, fabs(after[i] - before[i]));
    }
    printf("[ace_b200 bts %3d] level %u -> %zu, max|in| %.4g, max|out| %.4g, max|out-in| %.3g\n", call++,
           dbg_level, (size_t)res->_c0_poly._num_primes, mx_in, mx_out, mx_diff);
    free(before);
    free(after);
  }
  return res;
}

// =========================================================================== plaintexts
static void init_plain(PLAIN plain, uint32_t slots, size_t level, double sf, uint32_t deg) {
  Context* c = ctx_nf();
  plain->_scaling_factor = sf;
  plain->_sf_degree      = deg;
  plain->_slots          = slots;
  // always a fresh block: the encode that fills it is deferred (sched.h) and must not wait for
  // the ops still reading the previous contents
  free_poly_data(&plain->_poly);
  plain->_poly._ring_degree = c->N; plain->_poly._num_primes = level; plain->_poly._num_primes_p = 0;
  plain->_poly._num_alloc_primes = level;
  guard([&] { plain->_poly._data = reinterpret_cast<int64_t*>(g_queue->alloc(level, false)); });
  plain->_poly._is_ntt = true;
}

static void encode_plain(PLAIN plain, const double* vals, size_t len, uint32_t sc_degree,
                         uint32_t level) {
  StatScope ss(ST_ENCODE);
  Context* c = ctx_nf();
  if (level == 0) level = (uint32_t)c->L;
  double sf = Get_default_sc();
  if (len == 1) {  // plain_eval.c:28-34: constant fast path; one recorded fill per limb
    init_plain(plain, c->N / 2, level, pow(sf, sc_degree), sc_degree);
    guard([&] {
      std::vector<u64> res = c->value_residues(vals[0], level, sc_degree);
      for (uint32_t l = 0; l < level; l++)
        g_queue->fill(U(plain->_poly._data) + (size_t)l * c->N, res[l]);
    });
    return;
  }
  init_plain(plain, c->N / 2, level, pow(sf, sc_degree), sc_degree);
  ctx();  // the message is in host memory: staged and encoded at once
  guard([&] { c->encode(U(plain->_poly._data), vals, len, level, 0, sc_degree, 0); });
}

API void Encode_plain_from_float(PLAIN plain, float* input, size_t len, uint32_t sc_degree,
                                 uint32_t level) {
  std::vector<double> v(len);
  for (size_t i = 0; i < len; i++) v[i] = (double)input[i];
  encode_plain(plain, v.data(), len, sc_degree, level);
}
API void Encode_plain_from_double(PLAIN plain, double* input, size_t len, uint32_t sc_degree,
                                  uint32_t level) {
  encode_plain(plain, input, len, sc_degree, level);
}
API void Free_plain_poly(PLAIN plain) { free_poly_data(&plain->_poly); }

// ---- pre-encoded weights (DE_PLAINTEXT files) -----------------------------------
// Reference: Pt_mgr_init / Pt_get / Pt_free / Pt_prefetch (pt_mgr.c:35-176), Rt_data_prefetch
// (rt_data_file.c:61-80), PLAINTEXT_BUFFER and Cast_buffer_to_plain (rt_encode_api.h:22-27,
// plain_eval.c:132-160).  An entry is {magic "ANTPLAIN", version, size} + the PLAINTEXT struct with
// _poly._data == NULL + the limbs.  The reference keeps PT_ENTRY_COUNT (default 8) host buffers,
// slot = index % count, and returns a PLAINTEXT* into the slot; emitted code copies the struct
// (`dest = *(PLAIN)Pt_get(...)`, ir2c_ctx.h:84-91) and uses it until the slot comes round again.
// Here the limbs of a slot live in HBM: Pt_get reads the entry into pinned memory, validates it
// like Cast_buffer_to_plain, and copies the limbs to a FRESH block on the calling thread's
// stream; the block a slot held before is released through the scheduler's deferred free, so
// operations recorded against the previous occupant still read its limbs (the reference's
// synchronous semantics: what was returned stays valid for every call made before the slot
// is reused).  File and look-up table are shared, the ring is per host thread (as the I/O tables).
struct PlainBufHdr {  // PLAINTEXT_BUFFER without the flexible member
  char     magic[8];
  uint32_t version;
  uint32_t size;
};
constexpr uint32_t kRtVersionFull = 1;  // rt_version.h:15-22 (0.0.0 build 1)
int                   g_pt_fd = -1;
uint64_t              g_pt_fsize = 0;
std::vector<LutEntry> g_pt_lut;
uint32_t              g_pt_slots = 8, g_pt_prefetch = 2;
size_t                g_pt_max_entry = 0;
struct PtSlot {
  uint32_t    idx = (uint32_t)-1;
  PLAINTEXT   pt;
  char*       stage = nullptr;  // pinned
  cudaEvent_t copied = nullptr;
};
thread_local std::vector<PtSlot> g_pt_ring;

API void Pt_prefetch(uint32_t index);
static bool pt_file_open(int fd, uint64_t fsize, const DataFileHdr& h) {
  if (h.rt_ver != kRtVersionFull) die("rt data file version mismatch");
  if (h.lut_ofst > fsize || h.ent_count > (fsize - h.lut_ofst) / sizeof(LutEntry))
    die("weight data file: look-up table outside the file");
  g_pt_lut.resize(h.ent_count);
  const size_t lut_bytes = h.ent_count * sizeof(LutEntry);
  if (pread(fd, g_pt_lut.data(), lut_bytes, h.lut_ofst) != (ssize_t)lut_bytes) die("failed to read rt data file lookup table");
  g_pt_max_entry = 0;
  for (const LutEntry& e : g_pt_lut) {
    if (e.ent_ofst > fsize || e.size > fsize - e.ent_ofst) die("weight data file: entry outside the file");
    if (e.size < sizeof(PlainBufHdr) + sizeof(PLAINTEXT)) die("weight data file: plaintext entry too small");
    g_pt_max_entry = std::max(g_pt_max_entry, (size_t)e.size);
  }
  const char* ce = getenv("PT_ENTRY_COUNT");     // rt_env.h:25
  const char* pe = getenv("PT_PREFETCH_COUNT");  // rt_env.h:27
  g_pt_slots    = ce && atoi(ce) > 0 ? (uint32_t)atoi(ce) : 8;
  g_pt_prefetch = pe && atoi(pe) >= 0 ? (uint32_t)atoi(pe) : 2;
  g_pt_fd = fd;
  g_pt_fsize = fsize;
  g_etype = h.ent_type;
  g_nent  = h.ent_count;
  for (uint32_t i = 0; i < g_pt_prefetch; i++) Pt_prefetch(i);
  return true;
}
static void pt_ring_drop() {  // the calling thread's ring
  for (PtSlot& sl : g_pt_ring) {
    if (sl.idx != (uint32_t)-1 && sl.pt._poly._data) free_poly_data(&sl.pt._poly);
    if (sl.copied) { cudaEventSynchronize(sl.copied); cudaEventDestroy(sl.copied); }
    if (sl.stage) cudaFreeHost(sl.stage);
  }
  g_pt_ring.clear();
}
static void pt_file_close() {
  if (g_pt_fd < 0) return;
  pt_ring_drop();
  close(g_pt_fd);
  g_pt_fd = -1;
  g_pt_lut.clear();
}
// Pt_prefetch (pt_mgr.c:116-126): the reference starts an asynchronous read into the slot; here
// the page cache is asked to read ahead, the slot itself is only written by Pt_get
API void Pt_prefetch(uint32_t index) {
  if (g_pt_fd < 0 || index >= g_pt_lut.size()) return;  // rt_data_file.c:64: beyond the table is not an error
  posix_fadvise(g_pt_fd, (off_t)g_pt_lut[index].ent_ofst, (off_t)g_pt_lut[index].size, POSIX_FADV_WILLNEED);
}
API void* Pt_get(uint32_t index, size_t len, uint32_t scale, uint32_t level) {
  (void)len; (void)scale; (void)level;  // as in the reference: what was encoded at compile time counts
  if (g_pt_fd < 0) die("bad entry type");  // rt_data_file.c:63
  if (index >= g_pt_lut.size()) die("index out of entry range");
  StatScope ss(ST_ENCODE);
  Context* c = ctx_nf();
  if (g_pt_ring.empty()) g_pt_ring.resize(g_pt_slots);
  PtSlot& sl = g_pt_ring[index % g_pt_slots];
  const LutEntry& e = g_pt_lut[index];
  guard([&] {
    if (!sl.stage) {
      ACE_CUDA(cudaMallocHost(&sl.stage, g_pt_max_entry));
      ACE_CUDA(cudaEventCreateWithFlags(&sl.copied, cudaEventDisableTiming));
    } else {
      ACE_CUDA(cudaEventSynchronize(sl.copied));  // the previous occupant has left the staging buffer
    }
    size_t got = 0;
    while (got < e.size) {
      ssize_t r = pread(g_pt_fd, sl.stage + got, e.size - got, (off_t)(e.ent_ofst + got));
      if (r <= 0) die("failed to read rt data entry");
      got += (size_t)r;
    }
  });
  // Cast_buffer_to_plain (plain_eval.c:132-160), plus what a device copy has to know
  const PlainBufHdr* bh = reinterpret_cast<const PlainBufHdr*>(sl.stage);
  if (memcmp(bh->magic, "ANTPLAIN", 8) != 0) die("Plaintext buffer magic mismatch");
  if (bh->version != kRtVersionFull) die("Plaintext buffer version mismatch");
  if ((uint64_t)bh->size + sizeof(PlainBufHdr) > e.size) die("Plaintext buffer too small");
  PLAINTEXT pt;
  memcpy(&pt, sl.stage + sizeof(PlainBufHdr), sizeof(PLAINTEXT));
  if (pt._poly._data != nullptr) die("Plaintext poly data is not NULL");
  const uint64_t data_sz = sizeof(int64_t) * (uint64_t)pt._poly._num_alloc_primes * pt._poly._ring_degree;
  if (bh->size != data_sz + sizeof(PLAINTEXT)) die("Plaintext size mismatch");
  if (pt._poly._ring_degree != c->N || pt._poly._num_primes_p != 0 || pt._poly._num_primes == 0 ||
      pt._poly._num_primes > c->L || pt._poly._num_alloc_primes != pt._poly._num_primes)
    die("Plaintext does not fit the context (degree / number of primes)");
  // (the residues themselves are not range-checked: a value >= q gives a wrong product, never an
  // out-of-bounds access -- every address is derived from the checked sizes above)
  const size_t limbs = pt._poly._num_alloc_primes;
  const int64_t* src = reinterpret_cast<const int64_t*>(sl.stage + sizeof(PlainBufHdr) + sizeof(PLAINTEXT));
  if (sl.idx != (uint32_t)-1 && sl.pt._poly._data) free_poly_data(&sl.pt._poly);  // deferred: recorded readers go first
  guard([&] {
    ctx_nf();
    u64* dev = g_queue->alloc(limbs, false);
    ACE_CUDA(cudaMemcpyAsync(dev, src, data_sz, cudaMemcpyHostToDevice, c->stream));
    ACE_CUDA(cudaEventRecord(sl.copied, c->stream));
    pt._poly._data = reinterpret_cast<int64_t*>(dev);
  });
  sl.pt = pt;
  sl.idx = index;
  if (g_pt_prefetch > 0) Pt_prefetch(index + g_pt_prefetch);
  return &sl.pt;
}
API void* Pt_get_validate(float* buf, uint32_t index, size_t len, uint32_t scale, uint32_t level) {
  (void)buf; (void)index; (void)len; (void)scale; (void)level;
  die("TODO: not implemented");  // pt_mgr.c:161-164, word for word
  return nullptr;
}
// Pt_free (pt_mgr.c:166-176): the slot is given back; its limbs go once their readers are issued
API void Pt_free(uint32_t index) {
  if (g_pt_fd < 0 || g_pt_ring.empty()) return;
  PtSlot& sl = g_pt_ring[index % g_pt_slots];
  if (sl.idx != index) die("BLOCK_INFO state is not ready");
  if (sl.pt._poly._data) free_poly_data(&sl.pt._poly);
  sl.idx = (uint32_t)-1;
  if (g_pt_prefetch > 0) Pt_prefetch(index + g_pt_prefetch);
}

API bool Pt_mgr_init(const char* fname) {  // pt_mgr.c:35-110
  const char* override_path = getenv("ACE_B200_DATA_FILE");
  if (override_path && override_path[0]) fname = override_path;
  int fd = open(fname, O_RDONLY);
  if (fd < 0) {
    fprintf(stderr, "[ace_b200] cannot open weight data file %s\n", fname);
    die("weight data file missing");
  }
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size < (off_t)sizeof(DataFileHdr)) die("weight data file: cannot stat / too short");
  {
    DataFileHdr h0;
    if (pread(fd, &h0, sizeof(h0), 0) != (ssize_t)sizeof(h0)) die("failed to read rt data file header");
    if (memcmp(h0.magic, "!ANTFHE", 7) == 0 && h0.ent_type == DE_PLAINTEXT) return pt_file_open(fd, (uint64_t)st.st_size, h0);
  }
  g_wfile.resize(st.st_size);
  size_t got = 0;
  while (got < (size_t)st.st_size) {
    ssize_t r = pread(fd, g_wfile.data() + got, st.st_size - got, got);
    if (r <= 0) die("short read on weight data file");
    got += r;
  }
  close(fd);
  const DataFileHdr* h = reinterpret_cast<const DataFileHdr*>(g_wfile.data());
  if (memcmp(h->magic, "!ANTFHE", 7) != 0) die("bad weight data file magic");
  g_etype = h->ent_type;
  g_nent  = h->ent_count;
  // the header and the look-up table are untrusted input: every offset is checked against the
  // file size before anything on the host or the device is addressed through it
  if ((uint64_t)h->lut_ofst > g_wfile.size() ||
      (uint64_t)g_nent * sizeof(LutEntry) > g_wfile.size() - (uint64_t)h->lut_ofst)
    die("weight data file: look-up table outside the file");
  g_lut   = reinterpret_cast<const LutEntry*>(g_wfile.data() + h->lut_ofst);
  if (g_etype != DE_MSG_F32 && g_etype != DE_MSG_F64) die("only message data files are supported");
  for (size_t i = 0; i < g_nent; i++)
    if ((uint64_t)g_lut[i].ent_ofst > g_wfile.size() || (uint64_t)g_lut[i].size > g_wfile.size() - (uint64_t)g_lut[i].ent_ofst)
      die("weight data file: entry outside the file");
  guard([&] {
    Context* c = ctx();
    c->dev_malloc(&g_wfile_dev, g_wfile.size());
    c->h2d_sync(g_wfile_dev, g_wfile.data(), g_wfile.size());
  });
  return true;
}
API void Pt_mgr_fini(void) {
  pt_file_close();
  for (char* p : g_pt_slabs) cudaFree(p);  // the calling thread's plaintext cache
  g_pt_slabs.clear();
  g_pt_slab_left = 0;
  g_pt_cache.clear();
  g_pt_cache_bytes = 0;
  if (g_wfile_dev) cudaFree(g_wfile_dev);
  g_wfile_dev = nullptr;
  g_wfile.clear();
  g_wfile.shrink_to_fit();
  g_lut = nullptr;
  g_nent = 0;
}
// Pt_from_msg_validate (pt_mgr.c:193-205; emitted with the compiler's run-time validation option,
// ir2c_ctx.h:96-100): the constant the program carries must equal the file's entry, then as
// Pt_from_msg.  float32 message files only, as in the reference.
API void Pt_from_msg(void* pt, uint32_t index, size_t len, uint32_t scale, uint32_t level);
API void Pt_from_msg_validate(void* pt, float* buf, uint32_t index, size_t len, uint32_t scale, uint32_t level) {
  if (g_pt_fd >= 0) die("bad entry type");
  if (!g_lut || index >= g_nent) die("index out of entry range");
  const LutEntry& e = g_lut[index];
  if (g_etype != DE_MSG_F32 || e.size < len * sizeof(float)) die("entry size too small");
  const float* data = reinterpret_cast<const float*>(g_wfile.data() + e.ent_ofst);
  for (size_t i = 0; i < len; i++)
    if (!(fabs(buf[i] - data[i]) < 0.000001)) {
      fprintf(stderr, "Pt_from_msg_validate failed. index=%u, i=%zu: %f != %f.\n", index, i, buf[i], data[i]);
      die("Pt_from_msg_validate failed");
    }
  Pt_from_msg(pt, index, len, scale, level);
}

// Pt_from_msg (pt_mgr.c:182-191): look the message up and encode it at run time.  The message
// is read from the HBM copy of the weight file: no host->device copy, no synchronisation.
API void Pt_from_msg(void* pt, uint32_t index, size_t len, uint32_t scale, uint32_t level) {
  if (g_pt_fd >= 0) die("bad entry type");  // pt_mgr.c:184: a plaintext file has no messages
  if (!g_lut || index >= g_nent) die("Pt_from_msg: index out of range");
  const LutEntry& e = g_lut[index];
  const size_t esz = g_etype == DE_MSG_F32 ? sizeof(float) : sizeof(double);
  if (e.size < len * esz) die("Pt_from_msg: entry size too small");
  const char* data = g_wfile.data() + e.ent_ofst;
  if (len == 1) {  // constant fast path of Encode_plain_from_float (plain_eval.c:28-34)
    if (g_etype == DE_MSG_F32) Encode_plain_from_float((PLAIN)pt, (float*)data, len, scale, level);
    else Encode_plain_from_double((PLAIN)pt, (double*)data, len, scale, level);
    return;
  }
  StatScope ss(ST_ENCODE);
  PLAIN plain = (PLAIN)pt;
  Context* c = ctx_nf();
  if (level == 0) level = (uint32_t)c->L;
  init_plain(plain, c->N / 2, level, pow(Get_default_sc(), scale), scale);
  guard([&] {
    // Plaintext cache (the "encode cache" alternative to pre-encoded weight files):
    // the weights are constants of the model, so the plaintext of (entry, level, degree) is the
    // same for every image.  Each host thread keeps what it has encoded in HBM (its own stream
    // orders producer and consumers) and hands it out by renaming the caller's limbs -- no
    // copy, no encode; 11.5 GB per thread for ResNet-20.  ACE_B200_PT_CACHE_GB (default 24, 0 =
    // off) bounds it; beyond the bound plaintexts are encoded as before.
    const u64 key = (u64)index | ((u64)level << 32) | ((u64)scale << 48);
    const size_t bytes = (size_t)level * c->N * sizeof(u64);
    auto it = g_pt_cache.find(key);
    if (it != g_pt_cache.end()) {
      g_queue->alias(U(plain->_poly._data), it->second, level);
      c->tr(Context::TR_ENCODE, level);  // the reference would encode: keep the op trace complete
      return;
    }
    u64* dst = U(plain->_poly._data);
    u64* keep = nullptr;
    if (!g_eager && g_pt_cache_bytes + bytes <= pt_cache_limit() && pt_cache_has_room(bytes)) {
      keep = pt_cache_take(bytes);  // nullptr only means "do not cache", it must not disturb the run
      if (keep) {
        g_pt_cache[key] = keep;
        g_pt_cache_bytes += bytes;
        dst = keep;
      }
    }
    g_queue->encode(EncodeJob{dst, g_wfile_dev + e.ent_ofst,
                              g_etype == DE_MSG_F32 ? 0 : 1, (u32)len, level, 0, scale, 0});
    if (keep) g_queue->alias(U(plain->_poly._data), keep, level);
  });
}
