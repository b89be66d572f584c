// batch.cu -- the polynomial-level primitives over many independent polynomials at once.
//
// The reference runs Decomp_modup / Mod_down / Rescale one polynomial per call
// (fhe-cmplr/rtlib/ant/src/poly/poly_eval.c:28-49) and emitted programs call them that way,
// thousands of times per image at 2..12 limbs each -- far below one wave of the 148 SMs.  The
// scheduler (sched.h) collects the calls that do not depend on each other; here every step of a
// primitive (INTT, base conversion, NTT, tail) is ONE launch over the limbs of all collected
// jobs.  Per polynomial the arithmetic is exactly that of Context::modup_from / mod_down /
// rescale (context.cu), so results are bit-identical.
#include <algorithm>
#include <cstdlib>

#include "context.h"
#include "host_math.h"
#include "prof.h"

namespace ace {

// Folding the Rescale / ModDown limb-wise steps into the transform removes two passes and two
// launches per primitive -- and was measured SLOWER in its first form (Mod_down at l = 34: 124.8
// against 111.6 us, ResNet-20 0.978 against 0.955 s per image): the operand of the epilogue is
// loaded at the very end of the second kernel, its latency is not hidden, and the prologue's
// Switch_modulus sits on the transform's critical path.  Off unless ACE_B200_FUSED_TAILS=1.
bool fused_tails() {
  static const bool on = getenv("ACE_B200_FUSED_TAILS") != nullptr;
  return on;
}

namespace {
typedef uint16_t u16;

// collects limbs into LimbPtrBatch-sized launches
struct PtrBatcher {
  const DeviceTables& T;
  cudaStream_t        s;
  bool                inverse;
  size_t*             launches;
  u32                 per_launch;
  LimbPtrBatch        b;
  PtrBatcher(const DeviceTables& t, cudaStream_t st, bool inv, size_t* l)
      : T(t), s(st), inverse(inv), launches(l), per_launch(t.logN > 12 ? 2 : 1) { b.n = 0; }
  void add(u64* dst, const u64* src, u32 g) {
    if (b.n == kMaxPtrBatch) flush();
    b.dst[b.n] = dst; b.src[b.n] = src; b.g[b.n] = (u16)g;
    b.n++;
  }
  void flush() {
    if (b.n == 0) return;
    if (inverse) launch_intt(T, b, s); else launch_ntt(T, b, s);
    *launches += per_launch;
    b.n = 0;
  }
};

// forward transforms with the Rescale / ModDown limb-wise steps folded in (ntt16.cu)
struct FusedBatcher {
  const DeviceTables& T;
  cudaStream_t        s;
  size_t*             launches;
  NttFusedBatch       b;
  FusedBatcher(const DeviceTables& t, cudaStream_t st, size_t* l, int pre, int post, const u64* pre_w,
               const u64* pre_w_sh, u32 pre_stride, const u64* post_w, const u64* post_w_sh, u32 post_stride)
      : T(t), s(st), launches(l) {
    b.n = 0; b.pre = (uint8_t)pre; b.post = (uint8_t)post;
    b.pre_w = pre_w; b.pre_w_sh = pre_w_sh; b.post_w = post_w; b.post_w_sh = post_w_sh;
    b.pre_stride = pre_stride; b.post_stride = post_stride;
  }
  void add(u64* dst, const u64* src, const u64* aux, const u64* addend, u32 g, u32 g_from) {
    if (b.n == (u32)kMaxFused) flush();
    b.dst[b.n] = dst; b.src[b.n] = src; b.aux[b.n] = aux; b.add[b.n] = addend;
    b.g[b.n] = (u16)g; b.g_from[b.n] = (u16)g_from;
    b.n++;
  }
  void flush() {
    if (b.n == 0) return;
    launch_ntt16_fused(T, b, s);
    *launches += 2;
    b.n = 0;
  }
};

struct P3Batcher {
  Ptr3Batch P;
  P3Batcher() { P.n = 0; }
  bool full() const { return P.n == kMaxP3; }
  void add(u64* r, const u64* a, const u64* b, u32 g, u32 aux) {
    P.r[P.n] = r; P.a[P.n] = a; P.b[P.n] = b; P.g[P.n] = (u16)g; P.aux[P.n] = (u16)aux;
    P.n++;
  }
};
}  // namespace

// out-of-place copy of whole limbs, one descriptor per limb (kernel in kernels.cu's family)
__global__ void __launch_bounds__(256) copy_batch_kernel(u32 N, const __grid_constant__ Ptr3Batch P) {
  pdl_enter();
  const ulonglong2* a = reinterpret_cast<const ulonglong2*>(P.a[blockIdx.y]);
  ulonglong2*       r = reinterpret_cast<ulonglong2*>(P.r[blockIdx.y]);
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < N / 2; i += gridDim.x * blockDim.x)
    r[i] = a[i];
}

// ---------------------------------------------------------------------------- ModUp
void Context::modup_batch(const ModupJob* jobs, size_t n) {
  if (n == 0) return;
  if (n == 1 && jobs[0].copy_own) { modup_from(jobs[0].out, jobs[0].digit, jobs[0].num_q, jobs[0].part); return; }
  size_t total_in = 0;
  std::vector<const ModUpTab*> tabs(n);
  for (size_t j = 0; j < n; j++) {
    tabs[j] = &modup_tab(jobs[j].num_q, jobs[j].part);
    total_in += tabs[j]->n_in;
    tr(TR_MODUP_DIGIT, jobs[j].num_q);
  }
  u64* coef = alloc_limbs(total_in, false);
  {  // the digit's own limbs are copied through; everything else comes from the conversion
    P3Batcher cp;
    auto flush = [&] {
      if (cp.P.n == 0) return;
      prof::Scope ps("copy_batch", stream);
      launch_chain(copy_batch_kernel, dim3(N / 2 / 256 ? N / 2 / 256 : 1, cp.P.n), 256, 0, stream, N, cp.P);
      launches++;
      cp.P.n = 0;
    };
    for (size_t j = 0; j < n; j++) {
      const ModUpTab& t = *tabs[j];
      if (!jobs[j].copy_own) continue;  // the consumer reads the digit's own limbs from the source
      for (u32 i = 0; i < t.n_in; i++) {
        u64* dst = jobs[j].out + (size_t)(t.start + i) * N;
        const u64* src = jobs[j].digit + (size_t)i * N;
        if (dst == src) continue;
        if (cp.full()) flush();
        cp.add(dst, src, nullptr, 0, 0);
      }
    }
    flush();
  }
  {
    PtrBatcher ib(T, stream, true, &launches);
    size_t off = 0;
    for (size_t j = 0; j < n; j++) {
      const ModUpTab& t = *tabs[j];
      for (u32 i = 0; i < t.n_in; i++, off++)
        ib.add(coef + off * N, jobs[j].digit + (size_t)i * N, t.start + i);
    }
    ib.flush();
  }
  {
    ConvDesc descs[kMaxConvPack];
    u32      nd = 0;
    size_t   off = 0;
    for (size_t j = 0; j < n; j++) {
      const ModUpTab& t = *tabs[j];
      if (t.n_out) {
        fill_conv_desc(descs[nd++], t, coef + off * N, jobs[j].out);
        if (nd == (u32)kMaxConvPack) { launch_base_conv(T, descs, nd, stream); launches++; nd = 0; }
      }
      off += t.n_in;
    }
    if (nd) { launch_base_conv(T, descs, nd, stream); launches++; }
  }
  {
    PtrBatcher nb(T, stream, false, &launches);
    for (size_t j = 0; j < n; j++) {
      const ModUpTab& t = *tabs[j];
      for (u32 o = 0; o < t.n_out; o++) {
        u64* p = jobs[j].out + (size_t)t.out_slot[o] * N;
        nb.add(p, p, t.g_out[o]);
      }
    }
    nb.flush();
  }
  free_limbs(coef);
}

// ---------------------------------------------------------------------------- ModDown
void Context::moddown_batch(const ModdownJob* jobs, size_t n) {
  if (n == 0) return;
  if (n == 1 && !(ntt16_usable(T) && fused_tails())) { mod_down(jobs[0].out, jobs[0].in, jobs[0].num_q); return; }
  size_t total_q = 0;
  for (size_t j = 0; j < n; j++) { total_q += jobs[j].num_q; tr(TR_MODDOWN_POLY, jobs[j].num_q); }
  u64* pc   = alloc_limbs(n * K, false);
  u64* conv = alloc_limbs(total_q, false);
  {
    PtrBatcher ib(T, stream, true, &launches);
    for (size_t j = 0; j < n; j++)
      for (u32 i = 0; i < K; i++)
        ib.add(pc + (j * K + i) * N, jobs[j].in + (size_t)(jobs[j].num_q + i) * N, (u32)L + i);
    ib.flush();
  }
  {
    ConvDesc descs[kMaxConvPack];
    u32      nd = 0;
    size_t   off = 0;
    for (size_t j = 0; j < n; j++) {
      ConvDesc& d = descs[nd++];
      d.x = pc + j * K * N; d.out = conv + off * N;
      d.hatinv = phat_inv_; d.hatinv_sh = phat_inv_sh_; d.hatmod = phat_mod_q_;
      d.n_in = (u32)K; d.n_out = jobs[j].num_q;
      for (u32 i = 0; i < K; i++) d.g_in[i] = (u16)(L + i);
      for (u32 o = 0; o < jobs[j].num_q; o++) { d.g_out[o] = (u16)o; d.out_slot[o] = (u16)o; }
      off += jobs[j].num_q;
      if (nd == (u32)kMaxConvPack) { launch_base_conv(T, descs, nd, stream); launches++; nd = 0; }
    }
    if (nd) { launch_base_conv(T, descs, nd, stream); launches++; }
  }
  bool md_aliased = false;  // in-place Mod_down: the fused transform would overwrite its own operand
  for (size_t j = 0; j < n; j++)
    md_aliased |= jobs[j].out + (size_t)jobs[j].num_q * N > jobs[j].in &&
                  jobs[j].out < jobs[j].in + (size_t)(jobs[j].num_q + K) * N;
  if (ntt16_usable(T) && fused_tails() && !md_aliased) {
    // NTT of the converted limbs with the tail (old - conv) * P^-1 folded into its last store
    FusedBatcher fb(T, stream, &launches, 0, 2, nullptr, nullptr, 0, pinv_mod_q_, pinv_mod_q_sh_, 0);
    size_t off = 0;
    for (size_t j = 0; j < n; j++)
      for (u32 o = 0; o < jobs[j].num_q; o++, off++)
        fb.add(jobs[j].out + (size_t)o * N, conv + off * N, jobs[j].in + (size_t)o * N, nullptr, o, 0);
    fb.flush();
    free_limbs(pc);
    free_limbs(conv);
    return;
  }
  {
    PtrBatcher nb(T, stream, false, &launches);
    size_t off = 0;
    for (size_t j = 0; j < n; j++)
      for (u32 o = 0; o < jobs[j].num_q; o++, off++) nb.add(conv + off * N, conv + off * N, o);
    nb.flush();
  }
  {
    P3Batcher tb;
    size_t    off = 0;
    auto flush = [&] {
      launch_moddown_tail_batch(T, tb.P, pinv_mod_q_, pinv_mod_q_sh_, stream);
      launches++;
      tb.P.n = 0;
    };
    for (size_t j = 0; j < n; j++)
      for (u32 o = 0; o < jobs[j].num_q; o++, off++) {
        if (tb.full()) flush();
        tb.add(jobs[j].out + (size_t)o * N, jobs[j].in + (size_t)o * N, conv + off * N, o, 0);
      }
    if (tb.P.n) flush();
  }
  free_limbs(pc);
  free_limbs(conv);
}

// ---------------------------------------------------------------------------- Rescale
void Context::rescale_batch(const RescaleJob* jobs, size_t n) {
  if (n == 0) return;
  if (n == 1 && !(ntt16_usable(T) && fused_tails())) { rescale(jobs[0].out, jobs[0].in, jobs[0].num_q); return; }
  size_t total = 0;
  for (size_t j = 0; j < n; j++) {
    if (jobs[j].num_q < 2) throw std::runtime_error("Rescale: level not enough");
    total += jobs[j].num_q - 1;
    tr(TR_RESCALE_POLY, jobs[j].num_q);
  }
  u64* last = alloc_limbs(n, false);
  {
    PtrBatcher ib(T, stream, true, &launches);
    for (size_t j = 0; j < n; j++) {
      const u32 l = jobs[j].num_q - 1;
      ib.add(last + j * N, jobs[j].in + (size_t)l * N, l);
    }
    ib.flush();
  }
  bool aliased = false;  // in-place Rescale: the fused transform would overwrite its own operand
  for (size_t j = 0; j < n; j++) aliased |= jobs[j].out == jobs[j].in;
  if (ntt16_usable(T) && fused_tails() && !aliased) {
    // one fused transform per remaining limb: lift of the dropped limb (prologue), NTT,
    // c * q_l^-1 + t (epilogue)
    FusedBatcher fb(T, stream, &launches, 1, 1, negqlinv_, negqlinv_sh_, (u32)L, qlinv_, qlinv_sh_, (u32)L);
    for (size_t j = 0; j < n; j++) {
      const u32 l = jobs[j].num_q - 1;
      for (u32 i = 0; i < l; i++)
        fb.add(jobs[j].out + (size_t)i * N, last + j * N, jobs[j].in + (size_t)i * N, nullptr, i, l);
    }
    fb.flush();
    free_limbs(last);
    return;
  }
  u64* tmp  = alloc_limbs(total, false);
  {
    P3Batcher pb;
    size_t    off = 0;
    auto flush = [&] {
      launch_rescale_pre_batch(T, pb.P, negqlinv_, negqlinv_sh_, (u32)L, stream);
      launches++;
      pb.P.n = 0;
    };
    for (size_t j = 0; j < n; j++) {
      const u32 l = jobs[j].num_q - 1;
      for (u32 i = 0; i < l; i++, off++) {
        if (pb.full()) flush();
        pb.add(tmp + off * N, last + j * N, nullptr, i, l);
      }
    }
    if (pb.P.n) flush();
  }
  {
    PtrBatcher nb(T, stream, false, &launches);
    size_t off = 0;
    for (size_t j = 0; j < n; j++)
      for (u32 i = 0; i + 1 < jobs[j].num_q; i++, off++) nb.add(tmp + off * N, tmp + off * N, i);
    nb.flush();
  }
  {
    P3Batcher pb;
    size_t    off = 0;
    auto flush = [&] {
      launch_rescale_post_batch(T, pb.P, qlinv_, qlinv_sh_, (u32)L, stream);
      launches++;
      pb.P.n = 0;
    };
    for (size_t j = 0; j < n; j++) {
      const u32 l = jobs[j].num_q - 1;
      for (u32 i = 0; i < l; i++, off++) {
        if (pb.full()) flush();
        pb.add(jobs[j].out + (size_t)i * N, jobs[j].in + (size_t)i * N, tmp + off * N, i, l);
      }
    }
    if (pb.P.n) flush();
  }
  free_limbs(last);
  free_limbs(tmp);
}

}  // namespace ace
