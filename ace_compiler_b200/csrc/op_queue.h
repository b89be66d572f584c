// op_queue.h -- deferred execution of the per-limb "hardware" calls of emitted code.
//
// ACE-emitted C works at limb granularity: a ciphertext multiply-accumulate is a loop of
// 4*l calls  Hw_modmul, Hw_modmul, Hw_modadd, Hw_modadd  (fhe-cmplr/rtlib/ant/dataset/
// resnet20_cifar10_pre.onnx.inc:1493-1503, reference implementation ant/src/poly/
// poly_arith.c:14-56), ~350 000 calls per ResNet-20 image.  One kernel launch per call is
// launch-bound (a limb op moves 1.5 MiB, ~0.3 us of HBM time, against ~5 us per launch), so the
// runtime records the calls and executes them in batches:
//   * element-wise ops (add / mul) are coefficient-local, so a thread that owns coefficient i
//     can run a whole list of dependent ops for i in program order without any cross-thread
//     synchronisation.  Ops that share a limb (same pointer) form a chain; independent chains
//     run in different thread blocks (grid.y) of the same launch.
//   * gathers (Hw_rotate) read other coefficients, so they end the element-wise batch and are
//     batched among themselves (independent by construction in emitted code, verified here).
// Results are bit-identical to one-launch-per-call: the same residues are computed by the
// same arithmetic, only the launch boundaries move.
#pragma once
#include <unordered_map>
#include <vector>

#include "kernels.cuh"

namespace ace {

constexpr int kQueueCap = 448;  // items per launch: 448 * 32 B + chain table < 32 KiB of params

struct EwItem {
  u64*       r;
  const u64* a;
  const u64* b;   // second operand, or the int64 order table of a gather
  u32        g;   // modulus index
  u32        op;  // EwOp, or 3 = gather
};

struct EwPack {
  EwItem   it[kQueueCap];
  uint16_t chain_start[kQueueCap + 1];  // items of chain c: [chain_start[c], chain_start[c+1])
  u32      n_chains;
};

class OpQueue {
 public:
  OpQueue(const DeviceTables* T, cudaStream_t s, size_t* launch_counter, uint64_t* n_mul,
          uint64_t* n_add, uint64_t* n_rot)
      : T_(T), stream_(s), launches_(launch_counter), n_mul_(n_mul), n_add_(n_add),
        n_rot_(n_rot) {}
  void push_ew(EwOp op, u64* r, const u64* a, const u64* b, u32 g);
  void push_gather(u64* r, const u64* a, const int64_t* order, u32 g);
  void flush();
  bool empty() const { return items_.empty(); }
  size_t batches = 0, ops = 0;

 private:
  const DeviceTables* T_;
  cudaStream_t        stream_;
  size_t*             launches_;
  uint64_t *          n_mul_, *n_add_, *n_rot_;  // op-trace counters of the context
  std::vector<EwItem> items_;
  bool                gather_mode_ = false;
  std::unordered_map<const void*, int> owner_;  // limb pointer -> an item that touches it
  std::vector<int>    parent_;
  int  find(int x);
  void touch(const void* p, int k);
};

}  // namespace ace
