// Segmented gather-reduce: y[i,:] = beta*y[i,:] + sum_j ew[j] * x[col[j],:]  over a CSR.
//
// HBM-bound byte work: one warp owns one destination row (or one fixed-length
// segment of a heavy row); every gathered source row is one fully coalesced 128-bit-per-lane
// request (512 B at h=128), index/weight slices are read 32 at a time and broadcast by shuffle,
// four gathers are kept in flight per warp.  No [E,h] message tensor, no atomics on data.
// Heavy rows (hub genes: in-degree 1e4..1e5) are cut into seg_len-edge segments; the last warp
// to finish a row (ticket counter) folds the partials in segment order => deterministic.
#include "kgb_common.cuh"
#include "kgb_spmm_lean.cuh"
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

// One 32-slot slice of a CSR row held in registers: lane l owns slot base+l (column, weight, optional second scalar).
struct Chunk {
  int c;
  float w, w2;
};
template <bool kSum2>
__device__ __forceinline__ Chunk load_chunk(const int32_t* __restrict__ col, const EdgeW& e, int base, int end, int lane) {
  Chunk k{0, 0.f, 0.f};
  if (base + lane < end) {
    k.c = __ldg(col + base + lane);
    const int wi = e.wperm ? __ldg(e.wperm + base + lane) : base + lane;
    k.w = e.ew ? __ldg(e.ew + wi) : 1.f;
    if (kSum2) k.w2 = __ldg(e.ew2 + wi);
  }
  return k;
}

// acc += sum over slots [start, end) of w * x[col, :].  `first` = the slice at `start` (already loaded, possibly one
// item earlier: the row loop below prefetches it).  While a slice is being gathered the next one is already in flight.
// Gathers go out kUnroll at a time; a ragged tail is a PREDICATED full batch (all of its loads in flight together,
// w = 0 for the padding), never a serial one-by-one loop -- rows here are short (SNP rows: ~10 in-edges), the batch
// latency is paid once or twice per row, not once per edge.  Summation order = slot order (deterministic).
template <int H, int kUnroll, bool kSum2>
__device__ __forceinline__ void gather_accumulate(RowVec<H>& acc, Sum2& sum2, const int32_t* __restrict__ col,
                                                  const EdgeW& e, const float* __restrict__ x,
                                                  int64_t ldx, int start, int end, Chunk cur, int lane) {
  for (int base = start; base < end; base += kWarp) {
    const int n = min(kWarp, end - base);
    const Chunk nxt = load_chunk<kSum2>(col, e, base + kWarp, end, lane);
    if (kSum2) sum2.add(e.bins > 1 ? cur.c % e.bins : 0, cur.w2);
    for (int j = 0; j < n; j += kUnroll) {
      RowVec<H> t[kUnroll];
      float wj[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int cj = __shfl_sync(0xffffffffu, cur.c, (j + u) & 31);
        const float ww = __shfl_sync(0xffffffffu, cur.w, (j + u) & 31);
        const bool valid = j + u < n;                  // warp-uniform
        wj[u] = valid ? ww : 0.f;
        if (valid) t[u].load(x + (int64_t)cj * ldx, lane); else t[u].zero();
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) acc.fma(wj[u], t[u]);
    }
    cur = nxt;
  }
}

// What happens to a finished row: y = act(beta*y + bias + acc), and optionally the row's dot product with a fixed
// vector (the single-output head of kgwas/model.py:50,83 fused into the last writer of the SNP rows).
struct Epilogue {
  float beta;
  const float* bias;
  int relu;
  const float* dot_w;  // [H] or NULL
  float* dot_out;      // [n_rows]
};

template <int H>
__device__ __forceinline__ void write_row_with(const RowVec<H>& acc, const RowVec<H>& old, float* __restrict__ yrow,
                                               int64_t row, const Epilogue& ep, int lane) {
  RowVec<H> out = acc;
  if (ep.bias) {
    RowVec<H> bv;
    bv.load(ep.bias, lane);
    out.add(bv);
  }
  if (ep.beta != 0.f) {
#pragma unroll
    for (int i = 0; i < RowVec<H>::N; ++i) out.v[i] = fmaf(ep.beta, old.v[i], out.v[i]);
  }
  if (ep.relu) {
#pragma unroll
    for (int i = 0; i < RowVec<H>::N; ++i) out.v[i] = fmaxf(out.v[i], 0.f);
  }
  out.store(yrow, lane);
  if (ep.dot_w) {
    RowVec<H> wv;
    wv.load(ep.dot_w, lane);
    const float d = warp_sum(out.dot(wv));
    if (lane == 0) ep.dot_out[row] = d;
  }
}

template <int H>
__device__ __forceinline__ void write_row(const RowVec<H>& acc, float* __restrict__ yrow, int64_t row,
                                          const Epilogue& ep, int lane) {
  RowVec<H> old;
  old.zero();
  if (ep.beta != 0.f) old.load_plain(yrow, lane);
  write_row_with<H>(acc, old, yrow, row, ep, lane);
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

// Work items: [0, n_hsegs) heavy segments (long tasks first), then one item per row; a warp walks the items
// warp0, warp0 + n_warps, ... -- first its heavy segments, then its rows in a software pipeline:
//   iteration i   : row pointers of row i+2 requested, first index/weight slice of row i+1 requested, the old
//                   output row of row i requested (beta != 0), THEN the gathers of row i are issued and summed.
// Every dependent load of a row was therefore issued one full iteration before it is needed; what a row pays is the
// latency of its gather batches only (a short row used to pay five to seven dependent round trips).
// UNROLL = gathers in flight per warp and batch.
template <int H, int UNROLL, bool kSum2, bool kPipe, int kMinBlocks>
__global__ void __launch_bounds__(kSpmmThreads, (H <= 128 ? kMinBlocks : 2))
k_spmm(kgb_csr_t g, EdgeW e, const float* __restrict__ x, int64_t ldx, float* __restrict__ y, int64_t ldy, Epilogue ep,
       float* __restrict__ rowsum2, HeavyBufs hb) {
  float* __restrict__ partial = hb.partial;
  float* __restrict__ partial2 = hb.partial2;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  int64_t item = warp0;
  // ---------------------------------------------------------------- heavy segments
  for (; item < g.n_hsegs; item += n_warps) {
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
    gather_accumulate<H, UNROLL, kSum2>(acc, s2, g.col, e, x, ldx, start, end, load_chunk<kSum2>(g.col, e, start, end, lane), lane);
    acc.store(partial + (int64_t)seg * H, lane);
    if (kSum2) {
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
      if (kSum2 && lane < e.bins) {
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
      write_row<H>(sum, y + (int64_t)row * ldy, row, ep, lane);
      if (kSum2 && lane < e.bins) {
        float t2 = 0.f;
        for (int s = 0; s < ngrp; ++s) t2 += __ldcg(hb.gpartial2 + (int64_t)(grp0 + s) * e.bins + lane);
        rowsum2[(int64_t)row * e.bins + lane] = t2;
      }
      if (lane == 0) hb.ticket[hr] = 0;
    }
  }
  // ---------------------------------------------------------------- rows, software-pipelined
  // row pointers of a row: start < 0 marks "nothing to do" (past the end, or a heavy row handled above)
  auto row_meta = [&](int64_t r, int& s, int& en) {
    s = -1;
    en = -1;
    if (r < g.n_rows) {
      s = __ldg(g.rowptr + r);
      en = __ldg(g.rowptr + r + 1);
    }
  };
  auto skip = [&](int s, int en) { return s < 0 || (g.n_hsegs > 0 && en - s > g.seg_len); };
  int64_t row = item - g.n_hsegs;
  if (!kPipe) {  // plain loop (A/B reference for the pipelined one)
    for (; row < g.n_rows; row += n_warps) {
      const int start = __ldg(g.rowptr + row), end = __ldg(g.rowptr + row + 1);
      if (skip(start, end)) continue;
      RowVec<H> acc;
      acc.zero();
      Sum2 s2;
      s2.zero();
      gather_accumulate<H, UNROLL, kSum2>(acc, s2, g.col, e, x, ldx, start, end,
                                          load_chunk<kSum2>(g.col, e, start, end, lane), lane);
      write_row<H>(acc, y + row * ldy, row, ep, lane);
      if (kSum2) {
#pragma unroll
        for (int b = 0; b < KGB_MAX_BINS; ++b) {
          if (b < e.bins) {
            const float t2 = warp_sum(s2.v[b]);
            if (lane == 0) rowsum2[row * e.bins + b] = t2;
          }
        }
      }
    }
    return;
  }
  int sA, eA, sB, eB;
  row_meta(row, sA, eA);
  row_meta(row + n_warps, sB, eB);
  Chunk cA = skip(sA, eA) ? Chunk{0, 0.f, 0.f} : load_chunk<kSum2>(g.col, e, sA, eA, lane);
  for (; row < g.n_rows; row += n_warps) {
    int sC, eC;
    row_meta(row + 2 * n_warps, sC, eC);
    const Chunk cB = skip(sB, eB) ? Chunk{0, 0.f, 0.f} : load_chunk<kSum2>(g.col, e, sB, eB, lane);
    if (!skip(sA, eA)) {
      float* __restrict__ yrow = y + row * ldy;
      RowVec<H> old;
      old.zero();
      if (ep.beta != 0.f) old.load_plain(yrow, lane);
      RowVec<H> acc;
      acc.zero();
      Sum2 s2;
      s2.zero();
      gather_accumulate<H, UNROLL, kSum2>(acc, s2, g.col, e, x, ldx, sA, eA, cA, lane);
      write_row_with<H>(acc, old, yrow, row, ep, lane);
      if (kSum2) {
#pragma unroll
        for (int b = 0; b < KGB_MAX_BINS; ++b) {
          if (b < e.bins) {
            const float t2 = warp_sum(s2.v[b]);
            if (lane == 0) rowsum2[row * e.bins + b] = t2;
          }
        }
      }
    }
    sA = sB; eA = eB; cA = cB;
    sB = sC; eB = eC;
  }
}

// ---- variant 0: the pre-pipelining kernel, kept for A/B measurements (KGB_SPMM_VARIANT=0) ----
namespace v0 {
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

// Work items: [0, n_hsegs) heavy segments (long tasks first), then one item per row.
// UNROLL = gathers in flight per warp: 8 for long rows / segments (bandwidth), 4 with one more resident CTA per SM
// for CSRs dominated by short rows (latency: 784 k SNP rows with ~10 in-edges each).
template <int H, int UNROLL>
__global__ void __launch_bounds__(kSpmmThreads, (H <= 128 ? 3 : 2) + (UNROLL <= 4 && H <= 128 ? 1 : 0))
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

}  // namespace v0

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
  const Epilogue ep{beta, bias, relu, dot_w, dot_out};
  HeavyBufs hb{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  if (csr->n_hsegs > 0) {
    KGB_REQUIRE(csr->hrow_grpptr && csr->n_hgroups > 0, "spmm: hrow_grpptr / n_hgroups missing");
    const size_t need = kgb_spmm_scratch_bytes(csr->n_hrows, csr->n_hsegs, csr->n_hgroups, h);
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
  }
  const EdgeW e{ew, wperm, ew2, ew2 ? rowsum2_bins : 1};
  const bool heavy_dominated = csr->n_edges_hint > 0 && 2 * (int64_t)csr->n_hsegs * csr->seg_len > csr->n_edges_hint;
  const unsigned grid = spmm_grid((int64_t)csr->n_hsegs + csr->n_rows, h, heavy_dominated);
  // A/B switch for measurements (default 5): 0 = pre-pipelining generic kernel, 1 = generic with pipelined rows,
  // 2..4 = occupancy / pipelining variants of it, 5 = lean kernel (grid policy as before), 6 = lean, always persistent,
  // 7 = lean with 4 CTAs / SM
  static const char* venv = getenv("KGB_SPMM_VARIANT_FIXED");
  const char* vdyn = venv ? venv : getenv("KGB_SPMM_VARIANT");
  const int variant = vdyn ? atoi(vdyn) : 5;
  if (variant >= 5 && !ew2 && !wperm && h % 128 == 0 && (csr->n_hsegs == 0 || csr->hitem)) {
    const lean::Epi lep{beta, bias, relu, dot_w, dot_out};
    const lean::Heavy lhb{hb.ticket, hb.ticket1, hb.partial, hb.gpartial};
    const int64_t resident = (int64_t)kNumSMs * (h <= 128 ? (variant == 7 ? 4 : 3) : 2);
    unsigned lgrid = grid;
    if (variant >= 6 && lgrid > resident) lgrid = (unsigned)resident;
#define KGB_LEAN(NV, W, MB) lean::k_spmm_lean<NV, W, MB><<<lgrid, lean::kThreads, 0, stream>>>(*csr, ew, x, ldx, y, ldy, lep, lhb)
#define KGB_LEAN_W(NV, MB) do { if (ew) KGB_LEAN(NV, true, MB); else KGB_LEAN(NV, false, MB); } while (0)
    switch (h) {
      case 128: if (variant == 7) KGB_LEAN_W(1, 4); else KGB_LEAN_W(1, 3); break;
      case 256: KGB_LEAN_W(2, 2); break;
      case 384: KGB_LEAN_W(3, 2); break;
      default: KGB_LEAN_W(4, 1); break;
    }
#undef KGB_LEAN_W
#undef KGB_LEAN
    KGB_LAUNCH_OK();
    return KGB_OK;
  }
#define KGB_SPMM_LAUNCH(S2, PIPE, MB)                                                                           \
  KGB_DISPATCH_H(h, (k_spmm<H, 8, S2, PIPE, MB><<<grid, kSpmmThreads, 0, stream>>>(*csr, e, x, ldx, y, ldy, ep, rowsum2, hb)))
  if (variant == 0 && !dot_w) {
    KGB_DISPATCH_H(h, (v0::k_spmm<H, 8><<<grid, kSpmmThreads, 0, stream>>>(*csr, e, x, ldx, y, ldy, beta, bias, relu, rowsum2, hb)));
  } else if (ew2) {
    KGB_SPMM_LAUNCH(true, true, 3);
  } else if (variant == 2) {
    KGB_SPMM_LAUNCH(false, true, 2);
  } else if (variant == 3) {
    KGB_SPMM_LAUNCH(false, false, 3);
  } else if (variant == 4) {
    KGB_SPMM_LAUNCH(false, false, 2);
  } else {
    KGB_SPMM_LAUNCH(false, true, 3);
  }
#undef KGB_SPMM_LAUNCH
  KGB_LAUNCH_OK();
  return KGB_OK;
}
