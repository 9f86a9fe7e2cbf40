// Segmented gather-reduce: y[i,:] = beta*y[i,:] + sum_j ew[j] * x[col[j],:]  over a CSR.
//
// Two kernels behind kgb_spmm:
//   * lean::k_spmm_lean (kgb_spmm_lean.cuh) -- the SAGE mean / sum aggregation and its backward (no second scalar, no
//     weight permutation, h % 128 == 0): instruction-lean, software-pipelined; what the benchmark runs;
//   * k_spmm below -- the generic kernel: any supported width, weights through a slot permutation (transposed CSR of
//     the GAT backward), a second per-edge scalar summed per row into bins (attention-logit gradients).
//
// HBM-bound byte work: one warp owns one destination row (or one fixed-length
// segment of a heavy row); every gathered source row is one fully coalesced 128-bit-per-lane
// request (512 B at h=128), index/weight slices are read 32 at a time and broadcast by shuffle,
// four gathers are kept in flight per warp.  No [E,h] message tensor, no atomics on data.
// Heavy rows (hub genes: in-degree 1e4..1e5) are cut into seg_len-edge segments; the last warp
// to finish a row (ticket counter) folds the partials in segment order => deterministic.
#include "kgb_common.cuh"
#include "kgb_spmm_lean.cuh"
#include "kgb_spmm_hub.cuh"
#include <stdlib.h>

namespace kgb {

constexpr int kSpmmThreads = 256;  // 8 warps / CTA

struct EdgeW {            // per-edge scalars riding along the gather
  const float* ew;        // weight of the gathered row (NULL -> 1)
  const int32_t* wperm;   // weights are indexed ew[wperm[slot]] when non-NULL (transposed CSR)
  const float* ew2;       // second scalar, only row-summed into `bins` bins by col % bins (NULL -> unused)
  int bins;
};
struct Sum2 {
  float v[KGB_MAX_BINS];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int b = 0; b < KGB_MAX_BINS; ++b) v[b] = 0.f;
  }
  __device__ __forceinline__ void add(int bin, float x) {
#pragma unroll
    for (int b = 0; b < KGB_MAX_BINS; ++b) v[b] += (b == bin) ? x : 0.f;
  }
};

template <int H, int kUnroll>
__device__ __forceinline__ void gather_accumulate(RowVec<H>& acc, Sum2& sum2, const int32_t* __restrict__ col,
                                                  const EdgeW& e, const float* __restrict__ x,
                                                  int64_t ldx, int start, int end, int lane) {
  for (int base = start; base < end; base += kWarp) {
    const int n = min(kWarp, end - base);
    int c = 0;
    float w = 0.f;
    if (lane < n) {
      c = __ldg(col + base + lane);
      const int wi = e.wperm ? __ldg(e.wperm + base + lane) : base + lane;
      w = e.ew ? __ldg(e.ew + wi) : 1.f;
      if (e.ew2) sum2.add(e.bins > 1 ? c % e.bins : 0, __ldg(e.ew2 + wi));
    }
    int j = 0;
    for (; j + kUnroll <= n; j += kUnroll) {
      RowVec<H> t[kUnroll];
      float wj[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int cj = __shfl_sync(0xffffffffu, c, j + u);
        wj[u] = __shfl_sync(0xffffffffu, w, j + u);
        t[u].load(x + (int64_t)cj * ldx, lane);
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) acc.fma(wj[u], t[u]);
    }
    for (; j < n; ++j) {
      const int cj = __shfl_sync(0xffffffffu, c, j);
      const float wj = __shfl_sync(0xffffffffu, w, j);
      RowVec<H> t;
      t.load(x + (int64_t)cj * ldx, lane);
      acc.fma(wj, t);
    }
  }
}

template <int H>
__device__ __forceinline__ void write_row(const RowVec<H>& acc, float* __restrict__ yrow, float beta,
                                          const float* __restrict__ bias, int relu, int lane) {
  RowVec<H> out = acc;
  if (bias) {
    RowVec<H> bv;
    bv.load(bias, lane);
    out.add(bv);
  }
  if (beta != 0.f) {
    RowVec<H> old;
    old.load_plain(yrow, lane);
#pragma unroll
    for (int i = 0; i < RowVec<H>::N; ++i) out.v[i] = fmaf(beta, old.v[i], out.v[i]);
  }
  if (relu) {
#pragma unroll
    for (int i = 0; i < RowVec<H>::N; ++i) out.v[i] = fmaxf(out.v[i], 0.f);
  }
  out.store(yrow, lane);
}

struct HeavyBufs {
  int32_t* ticket;    // [n_hrows]    groups finished per heavy row
  int32_t* ticket1;   // [n_hgroups]  segments finished per group
  float* partial;     // [n_hsegs, H]
  float* gpartial;    // [n_hgroups, H]
  float* partial2;    // [n_hsegs, bins]
  float* gpartial2;   // [n_hgroups, bins]
};

// sum of n consecutive H-wide rows written earlier in this kernel (coherent loads), 8 in flight, index order
template <int H>
__device__ __forceinline__ RowVec<H> fold_rows(const float* base, int n, int lane) {
  RowVec<H> sum;
  sum.zero();
  int i = 0;
  for (; i + 8 <= n; i += 8) {
    RowVec<H> p[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) p[u].load_plain(base + (int64_t)(i + u) * H, lane);
#pragma unroll
    for (int u = 0; u < 8; ++u) sum.add(p[u]);
  }
  for (; i < n; ++i) {
    RowVec<H> p;
    p.load_plain(base + (int64_t)i * H, lane);
    sum.add(p);
  }
  return sum;
}

// Work items: [0, n_hsegs) heavy segments (long tasks first), then one item per row.
// UNROLL = gathers in flight per warp: 8 for long rows / segments (bandwidth), 4 with one more resident CTA per SM
// for CSRs dominated by short rows (latency: 784 k SNP rows with ~10 in-edges each).
template <int H, int UNROLL>
__global__ void __launch_bounds__(kSpmmThreads, H <= 128 ? 3 : 2)
k_spmm(kgb_csr_t g, EdgeW e, const float* __restrict__ x, int64_t ldx, float* __restrict__ y, int64_t ldy, float beta,
       const float* __restrict__ bias, int relu, float* __restrict__ rowsum2, HeavyBufs hb) {
  float* __restrict__ partial = hb.partial;
  float* __restrict__ partial2 = hb.partial2;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_items = (int64_t)g.n_hsegs + g.n_rows;
  for (int64_t item = warp0; item < n_items; item += n_warps) {
    if (item < g.n_hsegs) {
      const int seg = g.hseg_order ? __ldg(g.hseg_order + item) : (int)item;
      const int hr = __ldg(g.hseg_hrow + seg);
      const int row = __ldg(g.hrow_id + hr);
      const int seg0 = __ldg(g.hrow_segptr + hr), seg1 = __ldg(g.hrow_segptr + hr + 1);
      const int rs = __ldg(g.rowptr + row), re = __ldg(g.rowptr + row + 1);
      const int start = rs + (seg - seg0) * g.seg_len;
      const int end = min(re, start + g.seg_len);
      RowVec<H> acc;
      acc.zero();
      Sum2 s2;
      s2.zero();
      gather_accumulate<H, UNROLL>(acc, s2, g.col, e, x, ldx, start, end, lane);
      acc.store(partial + (int64_t)seg * H, lane);
      if (rowsum2) {
#pragma unroll
        for (int b = 0; b < KGB_MAX_BINS; ++b) {
          if (b < e.bins) {
            const float t2 = warp_sum(s2.v[b]);
            if (lane == 0) partial2[(int64_t)seg * e.bins + b] = t2;
          }
        }
      }
      __threadfence();  // publish this partial before taking a ticket
      // Two-level fold, each level done by the last finisher and always in index order (deterministic):
      // KGB_FOLD consecutive segments -> one group partial; the row's group partials -> the output row.
      const int P = seg1 - seg0;
      const int gi = (seg - seg0) / KGB_FOLD;
      const int gsz = min(KGB_FOLD, P - gi * KGB_FOLD);
      const int grp0 = __ldg(g.hrow_grpptr + hr), ngrp = __ldg(g.hrow_grpptr + hr + 1) - grp0;
      const int gid = grp0 + gi;
      int t = 0;
      if (lane == 0) t = atomicAdd(hb.ticket1 + gid, 1);
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t != gsz - 1) continue;
      __threadfence();
      {
        RowVec<H> sum = fold_rows<H>(partial + (int64_t)(seg0 + gi * KGB_FOLD) * H, gsz, lane);
        sum.store(hb.gpartial + (int64_t)gid * H, lane);
        if (rowsum2 && lane < e.bins) {
          float t2 = 0.f;
          const float* p2 = partial2 + (int64_t)(seg0 + gi * KGB_FOLD) * e.bins + lane;
          for (int s = 0; s < gsz; ++s) t2 += __ldcg(p2 + (int64_t)s * e.bins);
          hb.gpartial2[(int64_t)gid * e.bins + lane] = t2;
        }
        if (lane == 0) hb.ticket1[gid] = 0;  // leave the counters clean for the next launch
      }
      __threadfence();
      if (lane == 0) t = atomicAdd(hb.ticket + hr, 1);
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t != ngrp - 1) continue;
      __threadfence();
      {
        RowVec<H> sum = fold_rows<H>(hb.gpartial + (int64_t)grp0 * H, ngrp, lane);
        write_row<H>(sum, y + (int64_t)row * ldy, beta, bias, relu, lane);
        if (rowsum2 && lane < e.bins) {
          float t2 = 0.f;
          for (int s = 0; s < ngrp; ++s) t2 += __ldcg(hb.gpartial2 + (int64_t)(grp0 + s) * e.bins + lane);
          rowsum2[(int64_t)row * e.bins + lane] = t2;
        }
        if (lane == 0) hb.ticket[hr] = 0;
      }
    } else {
      const int row = (int)(item - g.n_hsegs);
      const int start = __ldg(g.rowptr + row), end = __ldg(g.rowptr + row + 1);
      if (end - start > g.seg_len && g.n_hsegs > 0) continue;  // heavy: handled above
      RowVec<H> acc;
      acc.zero();
      Sum2 s2;
      s2.zero();
      gather_accumulate<H, UNROLL>(acc, s2, g.col, e, x, ldx, start, end, lane);
      write_row<H>(acc, y + (int64_t)row * ldy, beta, bias, relu, lane);
      if (rowsum2) {
#pragma unroll
        for (int b = 0; b < KGB_MAX_BINS; ++b) {
          if (b < e.bins) {
            const float t2 = warp_sum(s2.v[b]);
            if (lane == 0) rowsum2[(int64_t)row * e.bins + b] = t2;
          }
        }
      }
    }
  }
}

inline unsigned spmm_grid(int64_t n_items, int h, bool heavy_dominated) {
  // Persistent grid-stride launch: exactly the number of CTAs that are resident at once (3 per SM at h <= 128,
  // 2 above), so that (a) at any moment the resident warps work on one contiguous range of items -- which is what
  // the L2-window ordering of the heavy segments needs -- and (b) 784 k one-row items do not pay for 98 k CTA launches.
  const int64_t warps_per_cta = kSpmmThreads / 32;
  int64_t ctas = (n_items + warps_per_cta - 1) / warps_per_cta;
  const int64_t resident = (int64_t)kNumSMs * (h <= 128 ? 3 : 2);
  // (measured on B200: CSRs whose edges sit mostly in heavy segments run ~8 % faster with one item per warp and
  //  CTAs dispatched in item order -- same contiguity, better tail balance -- while row-dominated CSRs lose 40 %)
  if (ctas > resident && !heavy_dominated) ctas = resident;
  if (ctas > 0x7fffffff) ctas = 0x7fffffff;
  if (ctas < 1) ctas = 1;
  return (unsigned)ctas;
}

int check_csr(const kgb_csr_t* g, const char* who) {
  KGB_REQUIRE(g && g->rowptr && g->n_rows >= 0, "%s: bad csr", who);
  KGB_REQUIRE(g->seg_len > 0, "%s: seg_len must be > 0", who);
  KGB_REQUIRE(g->n_hsegs == 0 || (g->hrow_id && g->hrow_segptr && g->hseg_hrow), "%s: heavy arrays missing", who);
  return KGB_OK;
}

}  // namespace kgb

using namespace kgb;

extern "C" size_t kgb_spmm_scratch_bytes(int32_t n_hrows, int32_t n_hsegs, int32_t n_hgroups, int32_t h) {
  return align_up((size_t)n_hrows * 4, 256) + align_up((size_t)n_hgroups * 4, 256) +
         align_up((size_t)n_hsegs * KGB_MAX_BINS * 4, 256) + align_up((size_t)n_hgroups * KGB_MAX_BINS * 4, 256) +
         align_up((size_t)n_hsegs * h * 4, 256) + align_up((size_t)n_hgroups * h * 4, 256) + 256;
}

extern "C" size_t kgb_spmm_scratch_bytes_csr(const kgb_csr_t* csr, int32_t h) {
  if (!csr) return 0;
  size_t n = kgb_spmm_scratch_bytes(csr->n_hrows, csr->n_hsegs, csr->n_hgroups, h);
  if (csr->hub_n > 0) n += align_up((size_t)csr->hub_n_cta * csr->hub_nv * h * 4, 256);
  return n;
}

namespace kgb {
// the hub rows of `csr` from shared-memory tiles (kgb_spmm_hub.cuh): tile kernel + fold of the per-CTA partials
template <int NV>
int launch_hub(const kgb_csr_t* csr, const float* x, float* y, int64_t ldy, const lean::Epi& ep, float* partial,
               cudaStream_t stream) {
  constexpr int H = NV * 128;
  const size_t smem = hub::smem_bytes(csr->hub_nv, csr->hub_tile_rows, csr->hub_chunk_cap, H);
  static size_t configured[16] = {0};
  int dev = 0;
  KGB_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 16 && configured[dev] < smem) {
    KGB_CUDA_OK(cudaFuncSetAttribute(hub::k_hub_tile<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[dev] = smem;
  } else if (dev >= 16) {
    KGB_CUDA_OK(cudaFuncSetAttribute(hub::k_hub_tile<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  hub::Params p{x, csr->hub_chunks, csr->hub_tile_off, partial, (int)csr->hub_n_cols, csr->hub_tile_rows,
                csr->hub_n_tiles, csr->hub_nv, csr->hub_chunk_cap};
  hub::k_hub_tile<NV><<<csr->hub_n_cta, hub::kThreads, smem, stream>>>(p);
  KGB_LAUNCH_OK();
  hub::k_hub_fold<NV><<<csr->hub_n, hub::kFoldWarps * 32, 0, stream>>>(partial, csr->hub_n_cta, csr->hub_nv, csr->hub_row,
                                                                      csr->hub_vptr, y, ldy, ep);
  KGB_LAUNCH_OK();
  return KGB_OK;
}
}  // namespace kgb

extern "C" int kgb_spmm(const kgb_csr_t* csr, const float* ew, const int32_t* wperm, const float* ew2, float* rowsum2,
                        int32_t rowsum2_bins, const float* x, int64_t ldx, float* y, int64_t ldy, int32_t h, float beta,
                        const float* bias, int32_t relu, const float* dot_w, float* dot_out, void* scratch,
                        size_t scratch_bytes, kgb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = check_csr(csr, "spmm")) return rc;
  if (csr->n_rows == 0) return KGB_OK;
  KGB_REQUIRE(x && y, "spmm: null x/y");
  KGB_REQUIRE(ldx >= h && ldy >= h && ldx % 4 == 0 && ldy % 4 == 0, "spmm: strides must be >= h and multiples of 4");
  KGB_REQUIRE(aligned16(x) && aligned16(y), "spmm: x/y must be 16-byte aligned");
  KGB_REQUIRE((ew2 == nullptr) == (rowsum2 == nullptr), "spmm: ew2 and rowsum2 go together");
  KGB_REQUIRE(!ew2 || (rowsum2_bins >= 1 && rowsum2_bins <= KGB_MAX_BINS), "spmm: rowsum2_bins must be in [1, %d]", KGB_MAX_BINS);
  KGB_REQUIRE(!bias || aligned16(bias), "spmm: bias must be 16-byte aligned");
  KGB_REQUIRE((dot_w == nullptr) == (dot_out == nullptr), "spmm: dot_w and dot_out go together");
  KGB_REQUIRE(!dot_w || aligned16(dot_w), "spmm: dot_w must be 16-byte aligned");
  HeavyBufs hb{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  float* hub_partial = nullptr;
  if (csr->n_hsegs > 0) {
    KGB_REQUIRE(csr->hrow_grpptr && csr->n_hgroups > 0, "spmm: hrow_grpptr / n_hgroups missing");
    const size_t need = kgb_spmm_scratch_bytes_csr(csr, h);
    if (scratch_bytes < need || !scratch) {
      set_error("spmm: scratch %zu < %zu", scratch_bytes, need);
      return KGB_ERR_WORKSPACE;
    }
    Carver ws(scratch);
    hb.ticket = ws.take<int32_t>(csr->n_hrows);  // counters first: caller zeroes them once
    hb.ticket1 = ws.take<int32_t>(csr->n_hgroups);
    hb.partial2 = ws.take<float>((size_t)csr->n_hsegs * KGB_MAX_BINS);
    hb.gpartial2 = ws.take<float>((size_t)csr->n_hgroups * KGB_MAX_BINS);
    hb.partial = ws.take<float>((size_t)csr->n_hsegs * h);
    hb.gpartial = ws.take<float>((size_t)csr->n_hgroups * h);
    if (csr->hub_n > 0) hub_partial = ws.take<float>((size_t)csr->hub_n_cta * csr->hub_nv * h);
  }
  const EdgeW e{ew, wperm, ew2, ew2 ? rowsum2_bins : 1};
  const bool heavy_dominated = csr->n_edges_hint > 0 && 2 * (int64_t)csr->n_hsegs * csr->seg_len > csr->n_edges_hint;
  const unsigned grid = spmm_grid((int64_t)csr->n_hsegs + csr->n_rows, h, heavy_dominated);
  // KGB_SPMM_VARIANT=0 forces the generic kernel (A/B measurements: profiles/r01_spmm_variants.md)
  const char* venv = getenv("KGB_SPMM_VARIANT");   // read per call: scratch/bench_spmm.py flips it at run time
  const bool want_lean = !(venv && atoi(venv) == 0);
  const bool lean_ok = !ew2 && !wperm && h % 128 == 0 && (csr->n_hsegs == 0 || (csr->hitem && csr->n_hitems == csr->n_hsegs));
  if (dot_w && !lean_ok) {
    set_error("spmm: the dot-product epilogue needs h %% 128 == 0, no ew2 / wperm, and csr.hitem when rows are segmented");
    return KGB_ERR_UNSUPPORTED;
  }
  if (lean_ok && (want_lean || dot_w)) {
    const lean::Epi lep{beta, bias, relu, dot_w, dot_out};
    const lean::Heavy lhb{hb.ticket, hb.ticket1, hb.partial, hb.gpartial};
    // Hub plan: the heaviest rows from shared-memory tiles (TMA), everything else through the pull kernel below with
    // the hub rows' segments left out of its item list.  KGB_SPMM_HUB=0 switches it off (A/B measurements).
    static const char* henv = getenv("KGB_SPMM_HUB");
    kgb_csr_t tail;
    const bool use_hub = csr->hub_n > 0 && ew && ew == csr->hub_ew && ldx == h && (h == 128 || h == 256) && hub_partial &&
                         !(henv && atoi(henv) == 0);
    if (use_hub) {
      KGB_REQUIRE(csr->hub_tile_rows > 0 && csr->hub_tile_rows <= 256 && csr->hub_chunk_cap % 16 == 0 && csr->hub_chunks &&
                      csr->hub_tile_off && csr->hub_row && csr->hub_vptr && csr->hub_n_cta > 0,
                  "spmm: malformed hub plan");
      KGB_REQUIRE(hub::smem_bytes(csr->hub_nv, csr->hub_tile_rows, csr->hub_chunk_cap, h) <= 232448,
                  "spmm: hub plan needs more shared memory than one SM has");
      KGB_REQUIRE(csr->n_hitems_tail == 0 || csr->hitem_tail, "spmm: hitem_tail missing");
      const int rc = h == 128 ? launch_hub<1>(csr, x, y, ldy, lep, hub_partial, stream)
                              : launch_hub<2>(csr, x, y, ldy, lep, hub_partial, stream);
      if (rc) return rc;
      tail = *csr;
      tail.hitem = csr->hitem_tail;
      tail.n_hitems = csr->n_hitems_tail;
      csr = &tail;
    }
    const int64_t resident = (int64_t)kNumSMs * (h <= 128 ? 3 : h <= 384 ? 2 : 1);
    unsigned lgrid = grid;
    // Row-dominated CSRs run `waves` x the resident CTA count instead of one persistent wave: CTA slots then turn
    // over every ~50 us and the high-priority side streams of the layer scheduler get their small kernels in between
    // (measured: a captured step is 7.35 ms with one wave, 6.83 ms with 8; the kernel alone loses ~3 %).
    static const char* wenv = getenv("KGB_SPMM_WAVES");
    const int waves = wenv ? atoi(wenv) : 8;
    if (use_hub) lgrid = spmm_grid((int64_t)csr->n_hitems + csr->n_rows, h, heavy_dominated);
    if (!heavy_dominated) {
      const int64_t all = ((int64_t)csr->n_hitems + csr->n_rows + 7) / 8;
      const int64_t cap = resident * (waves > 0 ? waves : 1);
      lgrid = (unsigned)(all < cap ? all : cap);
    }
#define KGB_LEAN(NV, MB)                                                                                             \
  do {                                                                                                               \
    if (ew) lean::k_spmm_lean<NV, true, MB, 1, true, 8><<<lgrid, lean::kThreads, 0, stream>>>(*csr, ew, x, ldx, y, ldy, lep, lhb);  \
    else lean::k_spmm_lean<NV, false, MB, 1, true, 8><<<lgrid, lean::kThreads, 0, stream>>>(*csr, ew, x, ldx, y, ldy, lep, lhb);   \
  } while (0)
    switch (h) {
      case 128: KGB_LEAN(1, 3); break;
      case 256: KGB_LEAN(2, 2); break;
      case 384: KGB_LEAN(3, 2); break;
      case 512: KGB_LEAN(4, 1); break;
      default:
        set_error("feature width %d unsupported (need 32,64,128,256,384,512)", (int)h);
        return KGB_ERR_UNSUPPORTED;
    }
#undef KGB_LEAN
    KGB_LAUNCH_OK();
    return KGB_OK;
  }
  KGB_DISPATCH_H(h, (k_spmm<H, 8><<<grid, kSpmmThreads, 0, stream>>>(*csr, e, x, ldx, y, ldy, beta, bias, relu, rowsum2, hb)));
  KGB_LAUNCH_OK();
  return KGB_OK;
}
