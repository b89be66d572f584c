// sched.h -- deferred, dependency-aware execution of the reference's polynomial-level API.
//
// ACE-emitted C drives the runtime one limb or one polynomial at a time: per ResNet-20 image
// ~2.5 M Hw_modadd/Hw_modmul/Hw_rotate calls (fhe-cmplr/rtlib/ant/src/poly/poly_arith.c:14-56),
// ~6 000 run-time encodes (Pt_from_msg, rtlib/common/src/pt_mgr.c:182-191) and ~15 000
// Decomp_modup / Mod_down / Rescale calls (ant/src/poly/poly_eval.c:28-49), each far too small
// to fill 148 SMs, plus ~60 000 malloc+memset / free pairs.  Executed call by call the GPU is
// launch-bound and mostly idle.  The scheduler records the calls instead and issues them in
// *waves*:
//   * every recorded op knows the limbs (512 KiB units, identified by device address) it reads
//     and writes; its wave is one more than the latest wave among the ops it depends on
//     (read-after-write, write-after-write and write-after-read), except that coefficient-local
//     ops (add, sub, mul, mul-add, zero fill, copy) may share a wave with the coefficient-local
//     ops they depend on: such ops form a *chain* that one thread per coefficient executes in
//     program order (ew_chain_kernel);
//   * ops of one wave and one kind run as one batched launch sequence (Context::*_batch,
//     batch.cu): nine independent Rotate() calls of an emitted convolution become one INTT, one
//     base conversion, one NTT ... over nine polynomials; the nine encodes of the inner loop
//     become one FFT launch over nine messages;
//   * stores that nothing can observe are dropped: the zero fill of Alloc_poly / Init_ciph_*
//     when the first use overwrites the limb, the product of `Hw_modmul(tmp, a, b)` that the
//     next `Hw_modadd(acc, acc, tmp)` consumes (fused to one multiply-add) once tmp is
//     overwritten or freed, and additions of a limb known to be zero;
//   * Free_poly_data is deferred until the ops that use the block have been issued;
//   * Decomp_modup results are kept in scheduler-owned buffers and the caller's output limbs are
//     *renamed* to them (every later read of such a limb is redirected, nothing is copied unless a
//     multi-limb consumer or the end of the window needs the data in place).  A repeated
//     Decomp_modup(x, part) of a polynomial that has not been written since -- the nine Rotate()
//     calls of one convolution input, resnet20_cifar10_pre.onnx.inc:6972-7063, each run their own
//     ModUp of the same c1 -- is served from the first result: SURVEY 8(f4), the hoisting the
//     reference's own bootstrap does by hand (ckks_bootstrap_context.c:1284-1299).  Because the
//     emitted `ext` buffer is never physically written, the digits of one key switch no longer
//     wait for each other (write-after-read on `ext`) and run as one batch.
// Every op still computes the same canonical residues from the same operands in an order
// consistent with program order, so results are bit-identical to call-by-call execution
// (ACE_B200_EAGER=1 flushes after every call; tests compare the two modes).
#pragma once
#include <cstdint>
#include <map>
#include <vector>

#include "context.h"

namespace ace {

enum SchedOp : uint8_t {
  OP_NOP = 0, OP_ADD, OP_SUB, OP_MUL, OP_MAC, OP_ZERO, OP_COPY, OP_FILL,  // coefficient-local ("chain")
  OP_GATHER, OP_ENCODE, OP_MODUP, OP_MODDOWN, OP_RESCALE
};

// One chain item as the kernel sees it (40 bytes).
//   ADD/SUB/MUL: r = a op b      MAC: z = a*b; if (t) t = z; r = c + z
//   ZERO: r = 0                  COPY: r = a          FILL: r = (u64)a  (every coefficient)
struct ChainItem {
  u64*       r;
  const u64* a;
  const u64* b;
  const u64* c;
  u64*       t;
  u32        g;
  u32        op;
};
constexpr int kChainCap = 600;  // items per launch: 600 * 48 B + chain table < 32 KiB of params
struct ChainPack {
  ChainItem it[kChainCap];
  uint16_t  chain_start[kChainCap + 1];
  u32       n_chains;
};

// What the scheduler needs from the machine underneath: limb memory and the batched primitives.
// ContextBackend (sched.cu) is the product -- everything runs on the B200 through Context;
// HostSimBackend (sched_selftest.cu) interprets the same batches on host arrays with toy
// arithmetic so that the dependency logic can be tested without a GPU (tests/test_cpu_sched.py).
struct SchedBackend {
  virtual ~SchedBackend() {}
  virtual u32    N() const = 0;
  virtual u32    K() const = 0;
  virtual u32    digit_start(u32 part) const = 0;
  virtual u32    digit_len(u32 num_q, u32 part) const = 0;
  virtual u64*   alloc(size_t n_limbs) = 0;
  virtual void   free(u64* block) = 0;
  virtual size_t block_limbs(const u64* block) const = 0;
  virtual void   count_limb_op(int kind) = 0;  // 0 mul, 1 add, 2 rotate (op trace)
  virtual void   run_chains(const ChainPack& pack, u32 n_chains) = 0;
  virtual void   run_gathers(const ChainPack& pack, u32 n) = 0;
  virtual void   run_encode(const EncodeJob* jobs, size_t n) = 0;
  virtual void   run_modup(const ModupJob* jobs, size_t n) = 0;
  virtual void   run_moddown(const ModdownJob* jobs, size_t n) = 0;
  virtual void   run_rescale(const RescaleJob* jobs, size_t n) = 0;
};
SchedBackend* make_context_backend(Context* c);

class Scheduler {
 public:
  explicit Scheduler(SchedBackend* backend);  // takes ownership
  explicit Scheduler(Context* c) : Scheduler(make_context_backend(c)) {}
  ~Scheduler();

  // ---- recording (what rt_shim.cu calls)
  void ew(SchedOp op, u64* r, const u64* a, const u64* b, u32 g);  // OP_ADD / OP_SUB / OP_MUL
  void zero(u64* r, size_t n_limbs);
  void copy(u64* r, const u64* a, size_t n_limbs);
  void fill(u64* r, u64 value);  // one limb, every coefficient = value (constant plaintext)
  void gather(u64* r, const u64* a, const int64_t* order, u32 g);
  void encode(const EncodeJob& j);
  void modup(u64* out, const u64* in, u32 num_q, u32 part);   // Decomp_modup
  void moddown(u64* out, const u64* in, u32 num_q);           // Mod_down
  void rescale(u64* out, const u64* in, u32 num_q);           // Rescale
  u64* alloc(size_t n_limbs, bool zeroed);                    // zero fill is a recorded op
  void free(u64* block);                                      // deferred
  // reads of the n limbs at dst are served from src from now on (until dst is written or freed):
  // how a cached, read-only result is handed out without a copy (plaintext cache of rt_shim.cu)
  void alias(u64* dst, const u64* src, size_t n_limbs);
  // the caller is about to use the stream itself: issue everything recorded so far
  void flush();
  bool empty() const { return ops_.empty() && frees_.empty(); }

  bool   eager = false;  // flush after every call
  bool   share_modup = true;  // keep and re-use Decomp_modup results (off: write them in place)
  size_t n_flush = 0, n_ops = 0, n_dead = 0, n_fused = 0, n_waves = 0, n_chain_launches = 0;
  size_t n_modup = 0, n_modup_shared = 0;  // Decomp_modup calls recorded / served from an earlier one

 private:
  struct Op {
    SchedOp    kind;
    uint8_t    t_live;  // OP_MAC: the product is also stored to t
    uint16_t   g;
    u32        wave;
    u64*       r;
    const u64* a;
    const u64* b;
    const u64* c;
    u64*       t;
    u32        p0, p1;  // heavy ops: num_q, part / index into enc_jobs_
  };
  // what the recorded ops have done so far to one limb
  struct Limb {
    u64      addr;
    u32      gen;
    int32_t  w_op;        // last writer in ops_ (-1: none recorded)
    u32      w_wave;
    u32      r_wave_chain, r_wave_heavy;
    uint8_t  has_w, w_heavy, has_r_chain, has_r_heavy, read_since, is_zero, w_is_t;
    u32        w_seq;   // stamp of the last recorded write (0: not written in this window)
    const u64* alias;   // reads of this limb are served from there (a kept Decomp_modup result)
  };
  struct ModupKey {
    const u64* digit;
    u32        num_q, part;
    bool operator<(const ModupKey& o) const {
      if (digit != o.digit) return digit < o.digit;
      if (num_q != o.num_q) return num_q < o.num_q;
      return part < o.part;
    }
  };
  struct ModupEntry {
    u64* buf;    // [num_q + K] limbs, owned by the scheduler until the end of the window
    u32  stamp;  // wseq_ when it was recorded: valid while no source limb has a later w_seq
  };
  SchedBackend*       c_;
  std::vector<Op>     ops_;
  std::vector<EncodeJob> enc_jobs_;
  std::vector<u64*>   frees_;
  std::vector<Limb>   table_;
  u32                 gen_ = 1, mask_ = 0, live_ = 0, wseq_ = 0;
  size_t              pending_free_bytes_ = 0;
  bool                in_flush_ = false, modup_nohit_ = false;
  std::map<ModupKey, ModupEntry> modup_cache_;
  std::vector<u64*>   aliased_;  // limbs that were given an alias in this window

  Limb& limb(const u64* addr);
  const u64* resolve(const u64* a);           // where a read of limb a is served from
  void  materialize(const u64* p, size_t n);  // bring renamed limbs back in place (block consumers)
  u32   dep_read(Limb& l, bool heavy) const;
  u32   dep_write(Limb& l, bool heavy) const;
  void  note_read(Limb& l, u32 wave, bool heavy);
  void  note_write(Limb& l, u32 wave, bool heavy, int32_t op, bool as_t);
  void  kill_if_unread(Limb& l);
  void  grow();
  void  maybe_flush();
  void  run_chains(std::vector<u32>& idx);
};

}  // namespace ace
