// GATConv edge-level kernels (kgwas/conv.py:150-151, 200-228): attention coefficients per softmax
// group, their backward, and the sampled dense-dense product that feeds it.  Scalar-per-edge work
// rides the same "one warp per group, heavy groups cut into segments, last finisher folds the
// partials in segment order" scheme as the gather-reduce kernel, so results are deterministic.
//
// A "group" is (destination node t, relation slot k): row g = t*R + k of the group CSR.
//   a_src index of edge j:  col[j]*R + k   when the CSR columns are source NODES   (af jobs)
//                           col[j]         when they are already table rows s*R+k  (xf jobs)
#include "kgb_common.cuh"

namespace kgb {

constexpr int kGatThreads = 256;

struct AttArgs {
  const float* a_src;
  const float* a_dst;
  int R;
  int src_is_node;
  float slope;
  float inv_t;
  int mode;
};

__device__ __forceinline__ float edge_u(const AttArgs& p, const int32_t* __restrict__ col, int j, int g, int k) {
  const int c = __ldg(col + j);
  const int ai = p.src_is_node ? c * p.R + k : c;
  return __ldg(p.a_src + ai) + __ldg(p.a_dst + g);
}
__device__ __forceinline__ float lrelu(float u, float slope) { return u > 0.f ? u : slope * u; }

// (max, sum exp) of z/T over slots [s, e)
__device__ __forceinline__ void seg_stats(const AttArgs& p, const int32_t* __restrict__ col, int s, int e, int g, int k,
                                          int lane, float& m, float& l) {
  float mx = -INFINITY;
  for (int j = s + lane; j < e; j += 32) mx = fmaxf(mx, lrelu(edge_u(p, col, j, g, k), p.slope) * p.inv_t);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = s + lane; j < e; j += 32) sum += __expf(lrelu(edge_u(p, col, j, g, k), p.slope) * p.inv_t - mx);
  m = mx;
  l = warp_sum(sum);
}

__device__ __forceinline__ void seg_write_alpha(const AttArgs& p, const int32_t* __restrict__ col, int s, int e, int g,
                                                int k, int lane, float m, float l, float* __restrict__ alpha) {
  const float inv = 1.f / (l + 1e-16f);
  for (int j = s + lane; j < e; j += 32) {
    const float z = lrelu(edge_u(p, col, j, g, k), p.slope);
    float a;
    if (p.mode == KGB_ATT_SOFTMAX) a = __expf(z * p.inv_t - m) * inv;
    else if (p.mode == KGB_ATT_SIGMOID) a = 1.f / (1.f + __expf(-z * p.inv_t));
    else a = z;
    alpha[j] = a;
  }
}

// Groups with at most kLightMax edges: ONE THREAD per group.  A transform-first job has one group per (destination,
// relation) -- 4.7 M groups of 1.7 edges on the KGWAS graph -- and a warp per group spends its time on bookkeeping
// (measured: 756 us for 8 M edges); consecutive threads own consecutive slot ranges, so their loads still coalesce.
constexpr int kLightMax = 16;

__global__ void __launch_bounds__(kGatThreads)
k_gat_alpha_light(kgb_csr_t g, AttArgs p, float* __restrict__ alpha) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < g.n_rows; row += stride) {
    const int s = __ldg(g.rowptr + row), e = __ldg(g.rowptr + row + 1);
    if (e == s || e - s > kLightMax) continue;
    const int k = (int)(row % p.R);
    const float ad = __ldg(p.a_dst + row);
    float m = -INFINITY;
    for (int j = s; j < e; ++j) {                       // pass 1: z (kept in alpha[j]) and its maximum
      const int c = __ldg(g.col + j);
      const float u = __ldg(p.a_src + (p.src_is_node ? c * p.R + k : c)) + ad;
      const float z = lrelu(u, p.slope);
      float a = z;
      if (p.mode == KGB_ATT_SOFTMAX) { a = z * p.inv_t; m = fmaxf(m, a); }
      else if (p.mode == KGB_ATT_SIGMOID) a = 1.f / (1.f + __expf(-z * p.inv_t));
      alpha[j] = a;
    }
    if (p.mode != KGB_ATT_SOFTMAX) continue;
    float l = 0.f;
    for (int j = s; j < e; ++j) {
      const float ex = __expf(alpha[j] - m);
      alpha[j] = ex;
      l += ex;
    }
    const float inv = 1.f / (l + 1e-16f);
    for (int j = s; j < e; ++j) alpha[j] *= inv;
  }
}

__global__ void __launch_bounds__(kGatThreads)
k_gat_dsoftmax_light(kgb_csr_t g, AttArgs p, const float* __restrict__ alpha, const float* __restrict__ dalpha,
                     float* __restrict__ du, float* __restrict__ da_dst) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < g.n_rows; row += stride) {
    const int s = __ldg(g.rowptr + row), e = __ldg(g.rowptr + row + 1);
    if (e - s > kLightMax) continue;
    const int k = (int)(row % p.R);
    float S = 0.f;
    if (p.mode == KGB_ATT_SOFTMAX)
      for (int j = s; j < e; ++j) S = fmaf(__ldg(alpha + j), __ldg(dalpha + j), S);
    const float ad = e > s ? __ldg(p.a_dst + row) : 0.f;
    float tot = 0.f;
    for (int j = s; j < e; ++j) {
      const float a = __ldg(alpha + j), da = __ldg(dalpha + j);
      float dz;
      if (p.mode == KGB_ATT_SOFTMAX) dz = a * (da - S) * p.inv_t;
      else if (p.mode == KGB_ATT_SIGMOID) dz = a * (1.f - a) * da * p.inv_t;
      else dz = da;
      const int c = __ldg(g.col + j);
      const float u = __ldg(p.a_src + (p.src_is_node ? c * p.R + k : c)) + ad;
      const float d = u > 0.f ? dz : p.slope * dz;
      du[j] = d;
      tot += d;
    }
    da_dst[row] = tot;                                   // empty groups: 0
  }
}

struct HeavyScratch {
  int32_t* ticket;  // [n_hrows]
  float* seg_a;     // [n_hsegs]
  float* seg_b;     // [n_hsegs]
  float* row_a;     // [n_hrows]
  float* row_b;     // [n_hrows]
};

// phase A: light groups fully; heavy segments -> partial stats -> per-heavy-row stats
__global__ void __launch_bounds__(kGatThreads)
k_gat_alpha_a(kgb_csr_t g, AttArgs p, float* __restrict__ alpha, HeavyScratch hs) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_items = (int64_t)g.n_hsegs + (g.mid_row_id ? g.n_mid_rows : g.n_rows);
  for (int64_t item = warp0; item < n_items; item += n_warps) {
    if (item < g.n_hsegs) {
      const int seg = g.hseg_order ? __ldg(g.hseg_order + item) : (int)item;
      const int hr = __ldg(g.hseg_hrow + seg);
      const int row = __ldg(g.hrow_id + hr);
      const int seg0 = __ldg(g.hrow_segptr + hr), seg1 = __ldg(g.hrow_segptr + hr + 1);
      const int rs = __ldg(g.rowptr + row), re = __ldg(g.rowptr + row + 1);
      const int s = rs + (seg - seg0) * g.seg_len, e = min(re, s + g.seg_len);
      const int k = row % p.R;
      if (p.mode != KGB_ATT_SOFTMAX) {
        seg_write_alpha(p, g.col, s, e, row, k, lane, 0.f, 0.f, alpha);
        continue;
      }
      float m, l;
      seg_stats(p, g.col, s, e, row, k, lane, m, l);
      if (lane == 0) { hs.seg_a[seg] = m; hs.seg_b[seg] = l; }
      __threadfence();
      int t = 0;
      if (lane == 0) t = atomicAdd(hs.ticket + hr, 1);
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t == seg1 - seg0 - 1) {
        __threadfence();
        float M = -INFINITY;                                  // lanes stride over the segments: fixed mapping
        for (int q = seg0 + lane; q < seg1; q += 32) M = fmaxf(M, __ldcg(hs.seg_a + q));
        M = warp_max(M);
        float L = 0.f;
        for (int q = seg0 + lane; q < seg1; q += 32) L += __ldcg(hs.seg_b + q) * __expf(__ldcg(hs.seg_a + q) - M);
        L = warp_sum(L);
        if (lane == 0) {
          hs.row_a[hr] = M;
          hs.row_b[hr] = L;
          hs.ticket[hr] = 0;
        }
      }
    } else {
      const int row = g.mid_row_id ? __ldg(g.mid_row_id + (item - g.n_hsegs)) : (int)(item - g.n_hsegs);
      const int s = __ldg(g.rowptr + row), e = __ldg(g.rowptr + row + 1);
      if (e - s <= kLightMax || (e - s > g.seg_len && g.n_hsegs > 0)) continue;   // light: k_gat_alpha_light
      const int k = row % p.R;
      float m = 0.f, l = 0.f;
      if (p.mode == KGB_ATT_SOFTMAX) seg_stats(p, g.col, s, e, row, k, lane, m, l);
      seg_write_alpha(p, g.col, s, e, row, k, lane, m, l, alpha);
    }
  }
}

// phase B (softmax only): heavy segments write alpha with their row's final stats
__global__ void __launch_bounds__(kGatThreads)
k_gat_alpha_b(kgb_csr_t g, AttArgs p, float* __restrict__ alpha, HeavyScratch hs) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t seg = warp0; seg < g.n_hsegs; seg += n_warps) {
    const int hr = __ldg(g.hseg_hrow + seg);
    const int row = __ldg(g.hrow_id + hr);
    const int seg0 = __ldg(g.hrow_segptr + hr);
    const int rs = __ldg(g.rowptr + row), re = __ldg(g.rowptr + row + 1);
    const int s = rs + ((int)seg - seg0) * g.seg_len, e = min(re, s + g.seg_len);
    seg_write_alpha(p, g.col, s, e, row, row % p.R, lane, hs.row_a[hr], hs.row_b[hr], alpha);
  }
}

// ---- backward of the attention coefficients ------------------------------------------------
// S_g = sum_j alpha_j * dalpha_j over [s, e)
__device__ __forceinline__ float seg_S(const float* __restrict__ alpha, const float* __restrict__ dalpha, int s, int e,
                                       int lane) {
  float acc = 0.f;
  for (int j = s + lane; j < e; j += 32) acc = fmaf(__ldg(alpha + j), __ldg(dalpha + j), acc);
  return warp_sum(acc);
}
// du_j for slots [s, e); returns sum_j du_j
__device__ __forceinline__ float seg_du(const AttArgs& p, const int32_t* __restrict__ col, const float* __restrict__ alpha,
                                        const float* __restrict__ dalpha, int s, int e, int g, int k, int lane, float S,
                                        float* __restrict__ du) {
  float acc = 0.f;
  for (int j = s + lane; j < e; j += 32) {
    const float a = __ldg(alpha + j), da = __ldg(dalpha + j);
    float dz;
    if (p.mode == KGB_ATT_SOFTMAX) dz = a * (da - S) * p.inv_t;
    else if (p.mode == KGB_ATT_SIGMOID) dz = a * (1.f - a) * da * p.inv_t;
    else dz = da;
    const float u = edge_u(p, col, j, g, k);
    const float d = u > 0.f ? dz : p.slope * dz;
    du[j] = d;
    acc += d;
  }
  return warp_sum(acc);
}

__global__ void __launch_bounds__(kGatThreads)
k_gat_dsoftmax_a(kgb_csr_t g, AttArgs p, const float* __restrict__ alpha, const float* __restrict__ dalpha,
                 float* __restrict__ du, float* __restrict__ da_dst, HeavyScratch hs) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_items = (int64_t)g.n_hsegs + (g.mid_row_id ? g.n_mid_rows : g.n_rows);
  for (int64_t item = warp0; item < n_items; item += n_warps) {
    if (item < g.n_hsegs) {
      if (p.mode != KGB_ATT_SOFTMAX) continue;  // no group statistic needed: phase B does everything
      const int seg = g.hseg_order ? __ldg(g.hseg_order + item) : (int)item;
      const int hr = __ldg(g.hseg_hrow + seg);
      const int row = __ldg(g.hrow_id + hr);
      const int seg0 = __ldg(g.hrow_segptr + hr), seg1 = __ldg(g.hrow_segptr + hr + 1);
      const int rs = __ldg(g.rowptr + row), re = __ldg(g.rowptr + row + 1);
      const int s = rs + (seg - seg0) * g.seg_len, e = min(re, s + g.seg_len);
      const float S = seg_S(alpha, dalpha, s, e, lane);
      if (lane == 0) hs.seg_a[seg] = S;
      __threadfence();
      int t = 0;
      if (lane == 0) t = atomicAdd(hs.ticket + hr, 1);
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t == seg1 - seg0 - 1) {
        __threadfence();
        float tot = 0.f;
        for (int q = seg0 + lane; q < seg1; q += 32) tot += __ldcg(hs.seg_a + q);
        tot = warp_sum(tot);
        if (lane == 0) {
          hs.row_a[hr] = tot;
          hs.ticket[hr] = 0;
        }
      }
    } else {
      const int row = g.mid_row_id ? __ldg(g.mid_row_id + (item - g.n_hsegs)) : (int)(item - g.n_hsegs);
      const int s = __ldg(g.rowptr + row), e = __ldg(g.rowptr + row + 1);
      if (e - s <= kLightMax || (e - s > g.seg_len && g.n_hsegs > 0)) continue;   // light: k_gat_dsoftmax_light
      float tot = 0.f;
      if (e > s) {
        const float S = p.mode == KGB_ATT_SOFTMAX ? seg_S(alpha, dalpha, s, e, lane) : 0.f;
        tot = seg_du(p, g.col, alpha, dalpha, s, e, row, row % p.R, lane, S, du);
      }
      if (lane == 0) da_dst[row] = tot;
    }
  }
}

__global__ void __launch_bounds__(kGatThreads)
k_gat_dsoftmax_b(kgb_csr_t g, AttArgs p, const float* __restrict__ alpha, const float* __restrict__ dalpha,
                 float* __restrict__ du, float* __restrict__ da_dst, HeavyScratch hs) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t seg = warp0; seg < g.n_hsegs; seg += n_warps) {
    const int hr = __ldg(g.hseg_hrow + seg);
    const int row = __ldg(g.hrow_id + hr);
    const int seg0 = __ldg(g.hrow_segptr + hr), seg1 = __ldg(g.hrow_segptr + hr + 1);
    const int rs = __ldg(g.rowptr + row), re = __ldg(g.rowptr + row + 1);
    const int s = rs + ((int)seg - seg0) * g.seg_len, e = min(re, s + g.seg_len);
    const float S = p.mode == KGB_ATT_SOFTMAX ? hs.row_a[hr] : 0.f;
    const float part = seg_du(p, g.col, alpha, dalpha, s, e, row, row % p.R, lane, S, du);
    if (lane == 0) hs.seg_b[seg] = part;
    __threadfence();
    int t = 0;
    if (lane == 0) t = atomicAdd(hs.ticket + hr, 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t == seg1 - seg0 - 1) {
      __threadfence();
      float tot = 0.f;
      for (int q = seg0 + lane; q < seg1; q += 32) tot += __ldcg(hs.seg_b + q);
      tot = warp_sum(tot);
      if (lane == 0) {
        da_dst[row] = tot;
        hs.ticket[hr] = 0;
      }
    }
  }
}

// ---- sampled dense-dense product: out[j] = <xrow[row(j), :], x[col[j], :]> -------------------
template <int H>
__global__ void __launch_bounds__(kGatThreads)
k_sddmm(kgb_csr_t g, const float* __restrict__ xrow, int64_t ldr, const float* __restrict__ x, int64_t ldx,
        float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_items = (int64_t)g.n_hsegs + g.n_rows;
  for (int64_t item = warp0; item < n_items; item += n_warps) {
    int row, s, e;
    if (item < g.n_hsegs) {
      const int seg = g.hseg_order ? __ldg(g.hseg_order + item) : (int)item;
      const int hr = __ldg(g.hseg_hrow + seg);
      row = __ldg(g.hrow_id + hr);
      const int seg0 = __ldg(g.hrow_segptr + hr);
      const int rs = __ldg(g.rowptr + row), re = __ldg(g.rowptr + row + 1);
      s = rs + (seg - seg0) * g.seg_len;
      e = min(re, s + g.seg_len);
    } else {
      row = (int)(item - g.n_hsegs);
      s = __ldg(g.rowptr + row);
      e = __ldg(g.rowptr + row + 1);
      if (e == s || (e - s > g.seg_len && g.n_hsegs > 0)) continue;
    }
    RowVec<H> r;
    r.load(xrow + (int64_t)row * ldr, lane);
    for (int base = s; base < e; base += 32) {
      const int n = min(32, e - base);
      const int c = lane < n ? __ldg(g.col + base + lane) : 0;
      float mine = 0.f;
      int j = 0;
      for (; j + 4 <= n; j += 4) {
        RowVec<H> t[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) t[u].load(x + (int64_t)__shfl_sync(0xffffffffu, c, j + u) * ldx, lane);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float d = warp_sum(r.dot(t[u]));
          if (lane == j + u) mine = d;
        }
      }
      for (; j < n; ++j) {
        RowVec<H> t;
        t.load(x + (int64_t)__shfl_sync(0xffffffffu, c, j) * ldx, lane);
        const float d = warp_sum(r.dot(t));
        if (lane == j) mine = d;
      }
      if (lane < n) out[base + lane] = mine;
    }
  }
}

inline unsigned item_grid(int64_t n_items) {
  int64_t ctas = (n_items + kGatThreads / 32 - 1) / (kGatThreads / 32);
  const int64_t cap = (int64_t)kNumSMs * 32;
  if (ctas > cap) ctas = cap;
  return (unsigned)(ctas < 1 ? 1 : ctas);
}

inline unsigned thread_grid(int64_t n_threads) {
  int64_t ctas = (n_threads + kGatThreads - 1) / kGatThreads;
  const int64_t cap = (int64_t)kNumSMs * 16;
  if (ctas > cap) ctas = cap;
  return (unsigned)(ctas < 1 ? 1 : ctas);
}

int check_csr(const kgb_csr_t* g, const char* who);

static int carve_heavy(const kgb_csr_t* g, void* scratch, size_t bytes, HeavyScratch* hs, const char* who) {
  *hs = HeavyScratch{nullptr, nullptr, nullptr, nullptr, nullptr};
  if (g->n_hsegs == 0) return KGB_OK;
  if (!scratch || bytes < kgb_gat_scratch_bytes(g->n_hrows, g->n_hsegs)) {
    set_error("%s: scratch %zu < %zu", who, bytes, kgb_gat_scratch_bytes(g->n_hrows, g->n_hsegs));
    return KGB_ERR_WORKSPACE;
  }
  Carver ws(scratch);
  hs->ticket = ws.take<int32_t>(g->n_hrows);
  hs->seg_a = ws.take<float>(g->n_hsegs);
  hs->seg_b = ws.take<float>(g->n_hsegs);
  hs->row_a = ws.take<float>(g->n_hrows);
  hs->row_b = ws.take<float>(g->n_hrows);
  return KGB_OK;
}

}  // namespace kgb

using namespace kgb;

extern "C" size_t kgb_gat_scratch_bytes(int32_t n_hrows, int32_t n_hsegs) {
  return 3 * align_up((size_t)n_hrows * 4, 256) + 2 * align_up((size_t)n_hsegs * 4, 256) + 256;
}

extern "C" int kgb_gat_alpha(const kgb_csr_t* groups, const float* a_src, const float* a_dst, int32_t n_slots,
                             int32_t src_is_node, float* alpha, float negative_slope, float temperature, int32_t mode,
                             void* scratch, size_t scratch_bytes, kgb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = check_csr(groups, "gat_alpha")) return rc;
  if (groups->n_rows == 0) return KGB_OK;
  KGB_REQUIRE(a_src && a_dst && alpha && n_slots >= 1 && temperature != 0.f, "gat_alpha: bad argument");
  KGB_REQUIRE(mode >= KGB_ATT_SOFTMAX && mode <= KGB_ATT_RAW, "gat_alpha: unknown mode %d", mode);
  HeavyScratch hs;
  if (int rc = carve_heavy(groups, scratch, scratch_bytes, &hs, "gat_alpha")) return rc;
  const AttArgs p{a_src, a_dst, n_slots, src_is_node, negative_slope, 1.f / temperature, mode};
  k_gat_alpha_light<<<thread_grid(groups->n_rows), kGatThreads, 0, stream>>>(*groups, p, alpha);
  KGB_LAUNCH_OK();
  if (groups->n_hsegs > 0 || groups->n_mid_rows != 0) {      // rows with more than kLightMax edges, heavy segments
    const int64_t items = (int64_t)groups->n_hsegs + (groups->mid_row_id ? groups->n_mid_rows : groups->n_rows);
    k_gat_alpha_a<<<item_grid(items), kGatThreads, 0, stream>>>(*groups, p, alpha, hs);
    KGB_LAUNCH_OK();
  }
  if (groups->n_hsegs > 0 && mode == KGB_ATT_SOFTMAX) {
    k_gat_alpha_b<<<item_grid(groups->n_hsegs), kGatThreads, 0, stream>>>(*groups, p, alpha, hs);
    KGB_LAUNCH_OK();
  }
  return KGB_OK;
}

extern "C" int kgb_gat_dsoftmax(const kgb_csr_t* groups, const float* a_src, const float* a_dst, int32_t n_slots,
                                int32_t src_is_node, const float* alpha, const float* dalpha, float* du, float* da_dst,
                                float negative_slope, float temperature, int32_t mode, void* scratch,
                                size_t scratch_bytes, kgb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = check_csr(groups, "gat_dsoftmax")) return rc;
  if (groups->n_rows == 0) return KGB_OK;
  KGB_REQUIRE(a_src && a_dst && alpha && dalpha && du && da_dst && n_slots >= 1 && temperature != 0.f,
              "gat_dsoftmax: bad argument");
  KGB_REQUIRE(mode >= KGB_ATT_SOFTMAX && mode <= KGB_ATT_RAW, "gat_dsoftmax: unknown mode %d", mode);
  HeavyScratch hs;
  if (int rc = carve_heavy(groups, scratch, scratch_bytes, &hs, "gat_dsoftmax")) return rc;
  const AttArgs p{a_src, a_dst, n_slots, src_is_node, negative_slope, 1.f / temperature, mode};
  k_gat_dsoftmax_light<<<thread_grid(groups->n_rows), kGatThreads, 0, stream>>>(*groups, p, alpha, dalpha, du, da_dst);
  KGB_LAUNCH_OK();
  if (groups->n_hsegs > 0 || groups->n_mid_rows != 0) {
    const int64_t items = (int64_t)groups->n_hsegs + (groups->mid_row_id ? groups->n_mid_rows : groups->n_rows);
    k_gat_dsoftmax_a<<<item_grid(items), kGatThreads, 0, stream>>>(*groups, p, alpha, dalpha, du, da_dst, hs);
    KGB_LAUNCH_OK();
  }
  if (groups->n_hsegs > 0) {
    k_gat_dsoftmax_b<<<item_grid(groups->n_hsegs), kGatThreads, 0, stream>>>(*groups, p, alpha, dalpha, du, da_dst, hs);
    KGB_LAUNCH_OK();
  }
  return KGB_OK;
}

extern "C" int kgb_sddmm(const kgb_csr_t* csr, const float* xrow, int64_t ldr, const float* x, int64_t ldx, int32_t h,
                         float* out, kgb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = check_csr(csr, "sddmm")) return rc;
  if (csr->n_rows == 0) return KGB_OK;
  KGB_REQUIRE(xrow && x && out, "sddmm: null pointer");
  KGB_REQUIRE(aligned16(xrow) && aligned16(x) && ldr % 4 == 0 && ldx % 4 == 0, "sddmm: alignment");
  KGB_DISPATCH_H(h, (k_sddmm<H><<<item_grid((int64_t)csr->n_hsegs + csr->n_rows), kGatThreads, 0, stream>>>(
                        *csr, xrow, ldr, x, ldx, out)));
  KGB_LAUNCH_OK();
  return KGB_OK;
}
