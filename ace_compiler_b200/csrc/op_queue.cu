// op_queue.cu -- see op_queue.h
#include "op_queue.h"
#include "prof.h"

#include <algorithm>

namespace ace {

// one thread per coefficient, blockIdx.y = chain; items of a chain run in program order
__global__ void __launch_bounds__(256) ew_chain_kernel(DeviceTables T,
                                                       const __grid_constant__ EwPack P) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T.N) return;
  const u32 k0 = P.chain_start[blockIdx.y], k1 = P.chain_start[blockIdx.y + 1];
  for (u32 k = k0; k < k1; k++) {
    const EwItem& it = P.it[k];
    const Modulus m  = T.mod[it.g];
    // plain (non-restrict, non-ldg) accesses: a later item may read what an earlier one wrote
    const u64 x = it.a[i], y = it.b[i];
    u64 z;
    if (it.op == EW_ADD) z = add_mod(x, y, m.q);
    else if (it.op == EW_SUB) z = sub_mod(x, y, m.q);
    else z = mul_mod(x, y, m);
    it.r[i] = z;
  }
}

// independent gathers: blockIdx.y = item
__global__ void __launch_bounds__(256) gather_batch_kernel(DeviceTables T,
                                                           const __grid_constant__ EwPack P) {
  const EwItem&  it    = P.it[blockIdx.y];
  const int64_t* order = reinterpret_cast<const int64_t*>(it.b);
  const u64      q     = T.mod[it.g].q;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < T.N; i += gridDim.x * blockDim.x) {
    const int64_t k = order[i];
    it.r[i] = k >= 0 ? it.a[k] : q - it.a[-k];
  }
}

int OpQueue::find(int x) {
  while (parent_[x] != x) x = parent_[x] = parent_[parent_[x]];
  return x;
}

void OpQueue::touch(const void* p, int k) {
  auto it = owner_.find(p);
  if (it == owner_.end()) {
    owner_.emplace(p, k);
  } else {
    int a = find(it->second), b = find(k);
    if (a != b) parent_[std::max(a, b)] = std::min(a, b);
  }
}

void OpQueue::push_ew(EwOp op, u64* r, const u64* a, const u64* b, u32 g) {
  if (gather_mode_ || items_.size() == (size_t)kQueueCap) flush();
  gather_mode_ = false;
  (op == EW_MUL ? *n_mul_ : *n_add_)++;
  int k = (int)items_.size();
  items_.push_back(EwItem{r, a, b, g, (u32)op});
  parent_.push_back(k);
  touch(r, k);
  touch(a, k);
  touch(b, k);
}

void OpQueue::push_gather(u64* r, const u64* a, const int64_t* order, u32 g) {
  if (!gather_mode_ || items_.size() == (size_t)kQueueCap) flush();
  gather_mode_ = true;
  // gathers of one batch must be independent: no source or destination may be a destination
  // of an earlier gather in the batch, and no destination may be an earlier source
  for (const EwItem& e : items_)
    if (e.r == a || e.r == r || e.a == r) { flush(); gather_mode_ = true; break; }
  (*n_rot_)++;
  items_.push_back(EwItem{r, a, reinterpret_cast<const u64*>(order), g, 3u});
}

void OpQueue::flush() {
  if (items_.empty()) return;
  static thread_local EwPack pack;
  const u32 n = (u32)items_.size();
  dim3 grid((T_->N + 255) / 256, 1);
  prof::Scope prof_scope_(gather_mode_ ? "gather_batch" : "ew_chain", stream_);
  if (gather_mode_) {
    for (u32 k = 0; k < n; k++) pack.it[k] = items_[k];
    pack.n_chains = n;
    grid.y = n;
    gather_batch_kernel<<<grid, 256, 0, stream_>>>(*T_, pack);
  } else {
    // order items by chain (stable: program order inside a chain)
    std::vector<int> root(n), chain_of(n, -1);
    u32 n_chains = 0;
    std::vector<int> chain_id(n, -1);
    for (u32 k = 0; k < n; k++) {
      root[k] = find((int)k);
      if (chain_id[root[k]] < 0) chain_id[root[k]] = (int)n_chains++;
      chain_of[k] = chain_id[root[k]];
    }
    std::vector<u32> cnt(n_chains + 1, 0);
    for (u32 k = 0; k < n; k++) cnt[chain_of[k] + 1]++;
    for (u32 c = 0; c < n_chains; c++) cnt[c + 1] += cnt[c];
    for (u32 c = 0; c <= n_chains; c++) pack.chain_start[c] = (uint16_t)cnt[c];
    std::vector<u32> pos(cnt.begin(), cnt.end() - 1);
    for (u32 k = 0; k < n; k++) pack.it[pos[chain_of[k]]++] = items_[k];
    pack.n_chains = n_chains;
    grid.y = n_chains;
    ew_chain_kernel<<<grid, 256, 0, stream_>>>(*T_, pack);
  }
  (*launches_)++;
  batches++;
  ops += n;
  items_.clear();
  parent_.clear();
  owner_.clear();
}

}  // namespace ace
