// Graph bookkeeping: COO -> CSR (+ transposed CSR) with stable edge order, heavy-row
// segmentation.  Integer work only; everything here must be bit-exact against
// oracle/bookkeeping.py (numpy stable argsort).
#include <cub/cub.cuh>

#include "kgb_common.cuh"

namespace kgb {

__global__ void k_prepare_keys(const int64_t* __restrict__ key64, int64_t n, int32_t* __restrict__ key32,
                               int32_t* __restrict__ val) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) {
    key32[i] = (int32_t)key64[i];
    val[i] = (int32_t)i;
  }
}

// rowptr[r] = first slot whose (sorted) key is >= r        (r in [0, n_rows])
__global__ void k_rowptr_from_sorted(const int32_t* __restrict__ sorted_keys, int64_t n, int32_t n_rows,
                                     int32_t* __restrict__ rowptr) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r > n_rows) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (sorted_keys[mid] < (int32_t)r) lo = mid + 1; else hi = mid;
  }
  rowptr[r] = (int32_t)lo;
}

// col[i] = other64[perm[i]]; key_next[i] = col[i]; val_next[i] = i
__global__ void k_gather_cols(const int64_t* __restrict__ other64, const int32_t* __restrict__ perm, int64_t n,
                              int32_t* __restrict__ col, int32_t* __restrict__ key_next,
                              int32_t* __restrict__ val_next) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) {
    int32_t c = (int32_t)other64[perm[i]];
    col[i] = c;
    if (key_next) { key_next[i] = c; val_next[i] = (int32_t)i; }
  }
}

// key32[i] = (int32) key64[idx[i]]
__global__ void k_gather_keys64(const int64_t* __restrict__ key64, const int32_t* __restrict__ idx, int64_t n,
                                int32_t* __restrict__ key32) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) key32[i] = (int32_t)key64[idx[i]];
}

// t_col[j] = dst_sorted[t_eperm[j]]
__global__ void k_gather_i32(const int32_t* __restrict__ table, const int32_t* __restrict__ idx, int64_t n,
                             int32_t* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = table[idx[i]];
}

static int bits_for(int64_t n) {  // number of key bits needed for values in [0, n)
  int b = 1;
  while ((int64_t(1) << b) < n) ++b;
  return b;
}

static size_t sort_temp_bytes(int64_t n) {
  size_t bytes = 0;
  cub::DoubleBuffer<int32_t> k(nullptr, nullptr), v(nullptr, nullptr);
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, (int)n, 0, 32, (cudaStream_t)0);
  if (e != cudaSuccess) {  // no device to ask (CPU-only host): conservative bound
    cudaGetLastError();
    bytes = (size_t)(1 << 20) + (size_t)n / 4;
  }
  return bytes;
}

}  // namespace kgb

using namespace kgb;

extern "C" size_t kgb_csr_build_workspace_bytes(int64_t n_edges, int64_t n_src, int64_t n_dst) {
  (void)n_src; (void)n_dst;
  size_t e = (size_t)(n_edges > 0 ? n_edges : 1);
  return 4 * align_up(e * sizeof(int32_t), 256) + align_up(sort_temp_bytes(n_edges), 256) + 256;
}

extern "C" int kgb_csr_build(const int64_t* src, const int64_t* dst, int64_t E, int64_t n_src, int64_t n_dst,
                             int32_t sort_cols, const int64_t* presort_key, int32_t* rowptr, int32_t* col, int32_t* eperm, int32_t* t_rowptr, int32_t* t_col,
                             int32_t* t_eperm, void* workspace, size_t workspace_bytes, kgb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  KGB_REQUIRE(E >= 0 && n_src >= 0 && n_dst >= 0, "csr_build: negative size");
  KGB_REQUIRE(E < (int64_t(1) << 31) && n_src < (int64_t(1) << 31) && n_dst < (int64_t(1) << 31),
              "csr_build: sizes must fit int32 (E=%lld)", (long long)E);
  KGB_REQUIRE(rowptr && (E == 0 || (src && dst && col && eperm)), "csr_build: null output");
  const bool want_t = (t_rowptr != nullptr);
  KGB_REQUIRE(!want_t || E == 0 || (t_col && t_eperm), "csr_build: t_col/t_eperm required with t_rowptr");
  if (workspace_bytes < kgb_csr_build_workspace_bytes(E, n_src, n_dst)) {
    set_error("csr_build: workspace %zu < %zu", workspace_bytes, kgb_csr_build_workspace_bytes(E, n_src, n_dst));
    return KGB_ERR_WORKSPACE;
  }
  const int T = 256;
  if (E == 0) {
    KGB_CUDA_OK(cudaMemsetAsync(rowptr, 0, (n_dst + 1) * sizeof(int32_t), stream));
    if (want_t) KGB_CUDA_OK(cudaMemsetAsync(t_rowptr, 0, (n_src + 1) * sizeof(int32_t), stream));
    return KGB_OK;
  }
  Carver ws(workspace);
  int32_t* k0 = ws.take<int32_t>(E);
  int32_t* k1 = ws.take<int32_t>(E);
  int32_t* v0 = ws.take<int32_t>(E);
  int32_t* v1 = ws.take<int32_t>(E);
  size_t temp_bytes = sort_temp_bytes(E);
  void* temp = ws.take<char>(temp_bytes);
  const unsigned gE = (unsigned)((E + T - 1) / T);

  // ---- by destination -------------------------------------------------------------
  if (sort_cols) {
    // pre-order the edge ids by source; the stable sort by destination below keeps that order inside a row
    k_prepare_keys<<<gE, T, 0, stream>>>(presort_key ? presort_key : src, E, k0, v0);
    KGB_LAUNCH_OK();
    cub::DoubleBuffer<int32_t> kb(k0, k1), vb(v0, v1);
    KGB_CUDA_OK(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, kb, vb, (int)E, 0, bits_for(n_src), stream));
    if (vb.Current() != v0) KGB_CUDA_OK(cudaMemcpyAsync(v0, vb.Current(), E * 4, cudaMemcpyDeviceToDevice, stream));
    k_gather_keys64<<<gE, T, 0, stream>>>(dst, v0, E, k0);   // k0[i] = dst[v0[i]]
    KGB_LAUNCH_OK();
  } else {
    k_prepare_keys<<<gE, T, 0, stream>>>(dst, E, k0, v0);
    KGB_LAUNCH_OK();
  }
  {
    cub::DoubleBuffer<int32_t> kb(k0, k1), vb(v0, v1);
    KGB_CUDA_OK(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, kb, vb, (int)E, 0, bits_for(n_dst), stream));
    // radix sort is stable: equal dst keep ascending original edge id
    if (kb.Current() != k1) KGB_CUDA_OK(cudaMemcpyAsync(k1, kb.Current(), E * 4, cudaMemcpyDeviceToDevice, stream));
    KGB_CUDA_OK(cudaMemcpyAsync(eperm, vb.Current(), E * 4, cudaMemcpyDeviceToDevice, stream));
  }
  // k1 = dst sorted ; eperm = original edge ids in dst-major order
  k_rowptr_from_sorted<<<(unsigned)((n_dst + 1 + T - 1) / T), T, 0, stream>>>(k1, E, (int32_t)n_dst, rowptr);
  KGB_LAUNCH_OK();
  k_gather_cols<<<gE, T, 0, stream>>>(src, eperm, E, col, want_t ? k0 : nullptr, want_t ? v0 : nullptr);
  KGB_LAUNCH_OK();
  if (!want_t) return KGB_OK;

  // ---- by source (transposed), stable w.r.t. CSR slot order ----------------------------
  // k0 = src of CSR slot i, v0 = i.  k1 (dst sorted) must survive -> sort k0 into a copy in v1's
  // partner buffers: use (k0 -> temp key buffer in t_col) to keep k1 intact.
  {
    cub::DoubleBuffer<int32_t> kb(k0, t_col), vb(v0, v1);
    KGB_CUDA_OK(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, kb, vb, (int)E, 0, bits_for(n_src), stream));
    if (kb.Current() != k0) KGB_CUDA_OK(cudaMemcpyAsync(k0, kb.Current(), E * 4, cudaMemcpyDeviceToDevice, stream));
    KGB_CUDA_OK(cudaMemcpyAsync(t_eperm, vb.Current(), E * 4, cudaMemcpyDeviceToDevice, stream));
  }
  k_rowptr_from_sorted<<<(unsigned)((n_src + 1 + T - 1) / T), T, 0, stream>>>(k0, E, (int32_t)n_src, t_rowptr);
  KGB_LAUNCH_OK();
  k_gather_i32<<<gE, T, 0, stream>>>(k1, t_eperm, E, t_col);
  KGB_LAUNCH_OK();
  return KGB_OK;
}

// ---- heavy-row segmentation ----------------------------------------------------------------
namespace kgb {

__global__ void k_heavy_flags(const int32_t* __restrict__ rowptr, int32_t n_rows, int32_t seg_len,
                              int32_t* __restrict__ flag, int32_t* __restrict__ nseg) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  int deg = rowptr[r + 1] - rowptr[r];
  bool heavy = deg > seg_len;
  flag[r] = heavy ? 1 : 0;
  nseg[r] = heavy ? (deg + seg_len - 1) / seg_len : 0;
}

__global__ void k_heavy_totals(const int32_t* flag, const int32_t* nseg, const int32_t* flag_scan,
                               const int32_t* seg_scan, int32_t n_rows, int32_t* totals) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    totals[0] = n_rows ? flag_scan[n_rows - 1] + flag[n_rows - 1] : 0;
    totals[1] = n_rows ? seg_scan[n_rows - 1] + nseg[n_rows - 1] : 0;
  }
}

__global__ void k_heavy_fill(const int32_t* __restrict__ flag, const int32_t* __restrict__ nseg,
                             const int32_t* __restrict__ flag_scan, const int32_t* __restrict__ seg_scan,
                             int32_t n_rows, int32_t n_hrows, int32_t n_hsegs, int32_t* __restrict__ hrow_id,
                             int32_t* __restrict__ hrow_segptr, int32_t* __restrict__ hseg_hrow) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r == 0) hrow_segptr[n_hrows] = n_hsegs;
  if (r >= n_rows || !flag[r]) return;
  int slot = flag_scan[r], s0 = seg_scan[r], ns = nseg[r];
  hrow_id[slot] = r;
  hrow_segptr[slot] = s0;
  for (int s = 0; s < ns; ++s) hseg_hrow[s0 + s] = slot;
}

static size_t scan_temp_bytes(int32_t n) {
  size_t bytes = 0;
  cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, bytes, (int32_t*)nullptr, (int32_t*)nullptr, n, (cudaStream_t)0);
  if (e != cudaSuccess) { cudaGetLastError(); bytes = (size_t)(1 << 20); }
  return bytes;
}
}  // namespace kgb

extern "C" size_t kgb_csr_heavy_workspace_bytes(int32_t n_rows) {
  size_t n = (size_t)(n_rows > 0 ? n_rows : 1);
  return 4 * align_up(n * 4, 256) + align_up(scan_temp_bytes(n_rows), 256) + 512;
}

extern "C" int kgb_csr_heavy_count(const int32_t* rowptr, int32_t n_rows, int32_t seg_len, int32_t* h_counts,
                                   void* workspace, size_t workspace_bytes, kgb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  KGB_REQUIRE(rowptr && h_counts && seg_len > 0 && n_rows >= 0, "csr_heavy_count: bad argument");
  if (workspace_bytes < kgb_csr_heavy_workspace_bytes(n_rows)) {
    set_error("csr_heavy_count: workspace too small");
    return KGB_ERR_WORKSPACE;
  }
  h_counts[0] = h_counts[1] = 0;
  if (n_rows == 0) return KGB_OK;
  Carver ws(workspace);
  int32_t* flag = ws.take<int32_t>(n_rows);
  int32_t* nseg = ws.take<int32_t>(n_rows);
  int32_t* flag_scan = ws.take<int32_t>(n_rows);
  int32_t* seg_scan = ws.take<int32_t>(n_rows);
  int32_t* totals = ws.take<int32_t>(2);
  size_t temp_bytes = scan_temp_bytes(n_rows);
  void* temp = ws.take<char>(temp_bytes);
  const int T = 256;
  k_heavy_flags<<<(n_rows + T - 1) / T, T, 0, stream>>>(rowptr, n_rows, seg_len, flag, nseg);
  KGB_LAUNCH_OK();
  KGB_CUDA_OK(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, flag, flag_scan, n_rows, stream));
  KGB_CUDA_OK(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, nseg, seg_scan, n_rows, stream));
  k_heavy_totals<<<1, 32, 0, stream>>>(flag, nseg, flag_scan, seg_scan, n_rows, totals);
  KGB_LAUNCH_OK();
  KGB_CUDA_OK(cudaMemcpyAsync(h_counts, totals, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  KGB_CUDA_OK(cudaStreamSynchronize(stream));
  return KGB_OK;
}

extern "C" int kgb_csr_heavy_fill(const int32_t* rowptr, int32_t n_rows, int32_t seg_len, int32_t n_hrows,
                                  int32_t n_hsegs, int32_t* hrow_id, int32_t* hrow_segptr, int32_t* hseg_hrow,
                                  void* workspace, size_t workspace_bytes, kgb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  (void)rowptr; (void)seg_len;
  KGB_REQUIRE(n_rows >= 0 && n_hrows >= 0 && n_hsegs >= 0, "csr_heavy_fill: bad argument");
  if (n_hrows == 0) return KGB_OK;
  KGB_REQUIRE(hrow_id && hrow_segptr && hseg_hrow, "csr_heavy_fill: null output");
  if (workspace_bytes < kgb_csr_heavy_workspace_bytes(n_rows)) {
    set_error("csr_heavy_fill: workspace too small");
    return KGB_ERR_WORKSPACE;
  }
  Carver ws(workspace);  // same carving as kgb_csr_heavy_count: the scans are still there
  int32_t* flag = ws.take<int32_t>(n_rows);
  int32_t* nseg = ws.take<int32_t>(n_rows);
  int32_t* flag_scan = ws.take<int32_t>(n_rows);
  int32_t* seg_scan = ws.take<int32_t>(n_rows);
  const int T = 256;
  k_heavy_fill<<<(n_rows + T - 1) / T, T, 0, stream>>>(flag, nseg, flag_scan, seg_scan, n_rows, n_hrows, n_hsegs,
                                                      hrow_id, hrow_segptr, hseg_hrow);
  KGB_LAUNCH_OK();
  return KGB_OK;
}
