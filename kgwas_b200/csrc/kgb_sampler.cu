// Full-neighbour frontier expansion on the GPU: the bookkeeping behind
// NeighborLoader(data, num_neighbors=[-1]*L, input_nodes=('SNP', ids), batch_size) of kgwas/kgwas.py:99-113
// (SURVEY.md section 8 f-2, Appendix A.7).  Pure integer work, bit-exact against oracle/bookkeeping.py:
//
//   kgb_frontier_count   in-degrees of the frontier nodes of one relation -> exclusive offsets + total (one sync)
//   kgb_frontier_expand  every in-edge of every frontier node, frontier order then edge order: edge ids + sources
//   kgb_frontier_add     the sources not seen before, in FIRST-OCCURRENCE order, get the next batch-local ids
//                        (atomicMin of the slot position per node = first occurrence; flag; exclusive scan; commit)
//
// The host (kgwas_b200/loader.py) loops hops x relations in the reference's order -- the local-id table is updated
// after every relation, exactly like the CPU sampler -- and relabels the kept edges at the end.  The in-adjacency of a
// relation is the destination-major CSR of kgb_csr_build (stable edge order): rowptr, col = sources, eperm = edge ids.
#include "kgb_common.cuh"
#include <cub/cub.cuh>
#include <limits.h>

namespace kgb {

constexpr int kST = 256;

__global__ void k_frontier_deg(const int32_t* __restrict__ ptr, const int32_t* __restrict__ frontier, int32_t n_f,
                               int32_t* __restrict__ deg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_f) {
    const int v = __ldg(frontier + i);
    deg[i] = __ldg(ptr + v + 1) - __ldg(ptr + v);
  }
}

__global__ void k_frontier_total(const int32_t* __restrict__ deg, int32_t* __restrict__ offsets, int32_t n_f) {
  if (blockIdx.x == 0 && threadIdx.x == 0) offsets[n_f] = n_f ? offsets[n_f - 1] + deg[n_f - 1] : 0;
}

// one thread per output slot: binary search for the frontier node that owns the slot
__global__ void k_frontier_expand(const int32_t* __restrict__ ptr, const int32_t* __restrict__ col,
                                  const int32_t* __restrict__ eperm, const int32_t* __restrict__ frontier,
                                  const int32_t* __restrict__ offsets, int32_t n_f, int32_t total,
                                  int32_t* __restrict__ eids, int32_t* __restrict__ srcs) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < total; s += stride) {
    int lo = 0, hi = n_f - 1;                       // last i with offsets[i] <= s
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (__ldg(offsets + mid) <= s) lo = mid; else hi = mid - 1;
    }
    const int v = __ldg(frontier + lo);
    const int slot = __ldg(ptr + v) + (int)(s - __ldg(offsets + lo));
    eids[s] = __ldg(eperm + slot);
    srcs[s] = __ldg(col + slot);
  }
}

__global__ void k_add_mark(const int32_t* __restrict__ cand, int32_t n, const int32_t* __restrict__ local,
                           int32_t* __restrict__ firstpos) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int v = __ldg(cand + i);
    if (__ldg(local + v) < 0) atomicMin(firstpos + v, (int)i);
  }
}
__global__ void k_add_flag(const int32_t* __restrict__ cand, int32_t n, const int32_t* __restrict__ local,
                           const int32_t* __restrict__ firstpos, int32_t* __restrict__ flag) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int v = __ldg(cand + i);
    flag[i] = (__ldg(local + v) < 0 && __ldg(firstpos + v) == (int)i) ? 1 : 0;
  }
}
__global__ void k_add_commit(const int32_t* __restrict__ cand, int32_t n, const int32_t* __restrict__ flag,
                             const int32_t* __restrict__ pos, int32_t count_base, int32_t* __restrict__ local,
                             int32_t* __restrict__ firstpos, int32_t* __restrict__ new_nodes, int32_t* __restrict__ n_new) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (flag[i]) {
      const int v = __ldg(cand + i);
      const int p = pos[i];
      new_nodes[p] = v;
      local[v] = count_base + p;
      firstpos[v] = INT_MAX;                       // leave the scratch table clean for the next call
    }
    if (i == n - 1) *n_new = pos[i] + flag[i];
  }
}

static size_t scan_bytes(int32_t n) {
  size_t bytes = 0;
  cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, bytes, (int32_t*)nullptr, (int32_t*)nullptr, n, (cudaStream_t)0);
  if (e != cudaSuccess) { cudaGetLastError(); bytes = (size_t)(1 << 20); }
  return bytes;
}
inline unsigned grid_for(int64_t n) {
  int64_t c = (n + kST - 1) / kST;
  const int64_t cap = (int64_t)kNumSMs * 16;
  if (c > cap) c = cap;
  return (unsigned)(c < 1 ? 1 : c);
}

}  // namespace kgb

using namespace kgb;

extern "C" size_t kgb_frontier_workspace_bytes(int64_t n) {
  const int32_t m = (int32_t)(n > 0 ? n : 1);
  return 2 * align_up((size_t)m * 4, 256) + align_up(scan_bytes(m), 256) + 512;
}

extern "C" int kgb_frontier_count(const int32_t* ptr, const int32_t* frontier, int32_t n_f, int32_t* offsets,
                                  int32_t* h_total, void* workspace, size_t workspace_bytes, kgb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  KGB_REQUIRE(h_total && n_f >= 0, "frontier_count: bad argument");
  *h_total = 0;
  if (n_f == 0) return KGB_OK;
  KGB_REQUIRE(ptr && frontier && offsets, "frontier_count: null pointer");
  if (!workspace || workspace_bytes < kgb_frontier_workspace_bytes(n_f)) {
    set_error("frontier_count: workspace too small");
    return KGB_ERR_WORKSPACE;
  }
  Carver ws(workspace);
  int32_t* deg = ws.take<int32_t>(n_f);
  ws.take<int32_t>(n_f);
  size_t temp_bytes = scan_bytes(n_f);
  void* temp = ws.take<char>(temp_bytes);
  k_frontier_deg<<<grid_for(n_f), kST, 0, stream>>>(ptr, frontier, n_f, deg);
  KGB_LAUNCH_OK();
  KGB_CUDA_OK(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, deg, offsets, n_f, stream));
  k_frontier_total<<<1, 32, 0, stream>>>(deg, offsets, n_f);
  KGB_LAUNCH_OK();
  KGB_CUDA_OK(cudaMemcpyAsync(h_total, offsets + n_f, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  KGB_CUDA_OK(cudaStreamSynchronize(stream));
  return KGB_OK;
}

extern "C" int kgb_frontier_expand(const int32_t* ptr, const int32_t* col, const int32_t* eperm, const int32_t* frontier,
                                   const int32_t* offsets, int32_t n_f, int32_t total, int32_t* eids, int32_t* srcs,
                                   kgb_stream_t stream_) {
  if (total == 0 || n_f == 0) return KGB_OK;
  KGB_REQUIRE(ptr && col && eperm && frontier && offsets && eids && srcs && total > 0, "frontier_expand: bad argument");
  k_frontier_expand<<<grid_for(total), kST, 0, (cudaStream_t)stream_>>>(ptr, col, eperm, frontier, offsets, n_f, total, eids,
                                                                       srcs);
  KGB_LAUNCH_OK();
  return KGB_OK;
}

extern "C" int kgb_frontier_add(const int32_t* cand, int32_t n, int32_t* local, int32_t* firstpos, int32_t count_base,
                                int32_t* new_nodes, int32_t* h_n_new, void* workspace, size_t workspace_bytes,
                                kgb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  KGB_REQUIRE(h_n_new && n >= 0, "frontier_add: bad argument");
  *h_n_new = 0;
  if (n == 0) return KGB_OK;
  KGB_REQUIRE(cand && local && firstpos && new_nodes, "frontier_add: null pointer");
  if (!workspace || workspace_bytes < kgb_frontier_workspace_bytes(n)) {
    set_error("frontier_add: workspace too small");
    return KGB_ERR_WORKSPACE;
  }
  Carver ws(workspace);
  int32_t* flag = ws.take<int32_t>(n);
  int32_t* pos = ws.take<int32_t>(n);
  size_t temp_bytes = scan_bytes(n);
  void* temp = ws.take<char>(temp_bytes);
  int32_t* d_n_new = reinterpret_cast<int32_t*>(ws.take<char>(256));
  const unsigned g = grid_for(n);
  k_add_mark<<<g, kST, 0, stream>>>(cand, n, local, firstpos);
  KGB_LAUNCH_OK();
  k_add_flag<<<g, kST, 0, stream>>>(cand, n, local, firstpos, flag);
  KGB_LAUNCH_OK();
  KGB_CUDA_OK(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, flag, pos, n, stream));
  k_add_commit<<<g, kST, 0, stream>>>(cand, n, flag, pos, count_base, local, firstpos, new_nodes, d_n_new);
  KGB_LAUNCH_OK();
  KGB_CUDA_OK(cudaMemcpyAsync(h_n_new, d_n_new, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  KGB_CUDA_OK(cudaStreamSynchronize(stream));
  return KGB_OK;
}
