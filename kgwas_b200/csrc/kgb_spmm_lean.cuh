// Instruction-lean segmented gather-reduce (the SAGE mean / sum aggregation and its backward).
//
// ncu on the generic kernel (profiles/r01_spmm_variants.md, profiles/r01_ncu_spmm_step.md): 32 warp instructions per edge, issue slots 45 % busy and
// every row paying five to seven dependent memory round trips -- the kernel is bound by instruction issue AND exposed
// latency long before L2 (38 %) or HBM.  This version
//   * spends ~7 instructions per edge: one SHFL for the column, one for the weight, one IMAD.WIDE for the row address,
//     one LDG.128 per 128 features, packed FFMA2 (fma.rn.f32x2, new on sm_100) for the accumulate;
//   * finishes a ragged tail as ONE exactly-sized batch (switch over 1..7), never a serial loop;
//   * software-pipelines the rows of a warp: row pointers are fetched 32 rows at a time (one per lane, two blocks
//     ahead), the first index/weight slice of row i+1 and the old output row of row i are requested before the gathers
//     of row i are issued -- a row pays for its gather batches only;
//   * reads everything a heavy segment needs from one 16-byte work-item record (start, length, segment, heavy row)
//     prepared at plan time, and prefetches the record and first slice of the warp's next segment.
// Summation order is the CSR slot order, exactly as in the generic kernel: results are bit-identical to it.
#pragma once
#include "kgb_common.cuh"

namespace kgb {
namespace lean {

constexpr unsigned kFull = 0xffffffffu;

template <int NV>
struct AccT {
  float2 v[2 * NV];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < 2 * NV; ++i) v[i] = make_float2(0.f, 0.f);
  }
  __device__ __forceinline__ void fma(float w, const float4 (&t)[NV]) {
    const float2 ww = make_float2(w, w);
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      v[2 * c] = __ffma2_rn(ww, make_float2(t[c].x, t[c].y), v[2 * c]);
      v[2 * c + 1] = __ffma2_rn(ww, make_float2(t[c].z, t[c].w), v[2 * c + 1]);
    }
  }
  __device__ __forceinline__ void add(const AccT& o) {
#pragma unroll
    for (int i = 0; i < 2 * NV; ++i) {
      v[i].x += o.v[i].x;
      v[i].y += o.v[i].y;
    }
  }
};

struct Slice {  // lane l: column and weight of CSR slot base + l
  int c;
  float w;
};

template <bool kHasW>
__device__ __forceinline__ Slice load_slice(const int32_t* __restrict__ col, const float* __restrict__ ew, int base, int n,
                                            int lane) {
  Slice s{0, 1.f};
  if (lane < n) {
    s.c = __ldg(col + base + lane);
    if (kHasW) s.w = __ldg(ew + base + lane);
  }
  return s;
}

// B gathers in flight, then B accumulates (slots j .. j+B-1 of the slice)
template <int NV, int B, bool kHasW>
__device__ __forceinline__ void batch(AccT<NV>& acc, const char* __restrict__ xl, uint32_t ldx_bytes, const Slice& s, int j) {
  float4 t[B][NV];
  float w[B];
#pragma unroll
  for (int u = 0; u < B; ++u) {
    const uint32_t c = (uint32_t)__shfl_sync(kFull, s.c, j + u);
    w[u] = kHasW ? __shfl_sync(kFull, s.w, j + u) : 1.f;
    const float4* p = reinterpret_cast<const float4*>(xl + (uint64_t)c * ldx_bytes);
#pragma unroll
    for (int q = 0; q < NV; ++q) t[u][q] = __ldg(p + 32 * q);
  }
#pragma unroll
  for (int u = 0; u < B; ++u) acc.fma(w[u], t[u]);
}

// acc += sum over n slots starting at `base`; `cur` is the slice at `base` (n > 32: further slices are fetched here,
// each one while the previous is being gathered)
template <int NV, bool kHasW, int kB>
__device__ __forceinline__ void gather(AccT<NV>& acc, const char* __restrict__ xl, uint32_t ldx_bytes,
                                       const int32_t* __restrict__ col, const float* __restrict__ ew, int base, int n,
                                       Slice cur, int lane) {
  while (true) {
    const int m = min(n, 32);
    Slice nxt{0, 1.f};
    if (n > 32) nxt = load_slice<kHasW>(col, ew, base + 32, n - 32, lane);
    int j = 0;
    for (; j + kB <= m; j += kB) batch<NV, kB, kHasW>(acc, xl, ldx_bytes, cur, j);
    switch (m - j) {
      case 1: batch<NV, 1, kHasW>(acc, xl, ldx_bytes, cur, j); break;
      case 2: batch<NV, 2, kHasW>(acc, xl, ldx_bytes, cur, j); break;
      case 3: batch<NV, 3, kHasW>(acc, xl, ldx_bytes, cur, j); break;
      case 4: batch<NV, 4, kHasW>(acc, xl, ldx_bytes, cur, j); break;
      case 5: batch<NV, 5, kHasW>(acc, xl, ldx_bytes, cur, j); break;
      case 6: batch<NV, 6, kHasW>(acc, xl, ldx_bytes, cur, j); break;
      case 7: batch<NV, 7, kHasW>(acc, xl, ldx_bytes, cur, j); break;
      default: break;
    }
    n -= 32;
    if (n <= 0) break;
    base += 32;
    cur = nxt;
  }
}

struct Epi {
  float beta;
  const float* bias;
  int relu;
  const float* dot_w;
  float* dot_out;
};

template <int NV>
__device__ __forceinline__ void load_old(float4 (&old)[NV], const float* yrow, int lane) {
#pragma unroll
  for (int q = 0; q < NV; ++q) old[q] = __ldcg(reinterpret_cast<const float4*>(yrow) + lane + 32 * q);
}

template <int NV>
__device__ __forceinline__ void finish_row(const AccT<NV>& acc, const float4 (&old)[NV], float* __restrict__ yrow,
                                           int64_t row, const Epi& ep, int lane) {
  float dot = 0.f;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    float4 o = make_float4(acc.v[2 * q].x, acc.v[2 * q].y, acc.v[2 * q + 1].x, acc.v[2 * q + 1].y);
    if (ep.bias) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias) + lane + 32 * q);
      o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
    }
    if (ep.beta != 0.f) {
      o.x = fmaf(ep.beta, old[q].x, o.x); o.y = fmaf(ep.beta, old[q].y, o.y);
      o.z = fmaf(ep.beta, old[q].z, o.z); o.w = fmaf(ep.beta, old[q].w, o.w);
    }
    if (ep.relu) {
      o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
    }
    *(reinterpret_cast<float4*>(yrow) + lane + 32 * q) = o;
    if (ep.dot_w) {
      const float4 d = __ldg(reinterpret_cast<const float4*>(ep.dot_w) + lane + 32 * q);
      dot = fmaf(o.x, d.x, dot); dot = fmaf(o.y, d.y, dot); dot = fmaf(o.z, d.z, dot); dot = fmaf(o.w, d.w, dot);
    }
  }
  if (ep.dot_w) {
    dot = warp_sum(dot);
    if (lane == 0) ep.dot_out[row] = dot;
  }
}

struct Heavy {
  int32_t* ticket;    // [n_hrows]
  int32_t* ticket1;   // [n_hgroups]
  float* partial;     // [n_hsegs, H]
  float* gpartial;    // [n_hgroups, H]
};

// sum of n consecutive rows of NV*128 floats written earlier in this kernel (L2-coherent loads), index order
template <int NV>
__device__ __forceinline__ AccT<NV> fold(const float* base, int n, int lane) {
  AccT<NV> sum;
  sum.zero();
  const float4* p = reinterpret_cast<const float4*>(base) + lane;
  int i = 0;
  for (; i + 8 <= n; i += 8) {
    float4 t[8][NV];
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int q = 0; q < NV; ++q) t[u][q] = __ldcg(p + (int64_t)(i + u) * (32 * NV) + 32 * q);
#pragma unroll
    for (int u = 0; u < 8; ++u) sum.fma(1.f, t[u]);
  }
  for (; i < n; ++i) {
    float4 t[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) t[q] = __ldcg(p + (int64_t)i * (32 * NV) + 32 * q);
    sum.fma(1.f, t);
  }
  return sum;
}

template <int NV>
__device__ __forceinline__ void store_acc(const AccT<NV>& a, float* row, int lane) {
#pragma unroll
  for (int q = 0; q < NV; ++q)
    *(reinterpret_cast<float4*>(row) + lane + 32 * q) = make_float4(a.v[2 * q].x, a.v[2 * q].y, a.v[2 * q + 1].x, a.v[2 * q + 1].y);
}

constexpr int kThreads = 256;

// kDist: how many rows ahead the index/weight slices are requested (1 or 2); kOldEarly: request the old output row
// (beta != 0) before the gathers instead of after them; kB (<= 8): gathers in flight per batch.  Measured on B200 at
// h=128 (profiles/r01_spmm_variants.md): (3 CTAs/SM, kDist 1, early, 8) is the fastest -- a second row of prefetch,
// a late old-row load, 2 or 4 CTAs per SM and 6- or 12-wide batches all lose 10-40 %; only that one is instantiated.
template <int NV, bool kHasW, int kMinBlocks, int kDist, bool kOldEarly, int kB>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
k_spmm_lean(kgb_csr_t g, const float* __restrict__ ew, const float* __restrict__ x, int64_t ldx, float* __restrict__ y,
            int64_t ldy, Epi ep, Heavy hb) {
  constexpr int H = NV * 128;
  const int lane = threadIdx.x & 31;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  // this lane's 16 bytes of row 0; kept opaque so that a gather address is ONE IMAD.WIDE (col * row bytes + xl)
  uint64_t xl_u = reinterpret_cast<uint64_t>(x) + (uint64_t)lane * 16;
  asm volatile("" : "+l"(xl_u));
  const char* __restrict__ xl = reinterpret_cast<const char*>(xl_u);
  const uint32_t ldx_bytes = (uint32_t)(ldx * 4);
  const int32_t* __restrict__ col = g.col;

  // ------------------------------------------------------------------ heavy segments (work items [0, n_hsegs))
  // work items [0, n_items): the heavy segments this launch is responsible for (all of them, or -- when the hub rows
  // run through hub::k_hub_tile -- the ones of the other heavy rows)
  const int n_items = g.n_hitems;
  int item = warp0;
  if (item < n_items) {
    const int4* __restrict__ items = reinterpret_cast<const int4*>(g.hitem);
    int4 it = __ldg(items + item);                               // (start, length, segment, heavy row)
    Slice cur = load_slice<kHasW>(col, ew, it.x, it.y, lane);
    while (true) {
      const int nitem = item + n_warps;
      const bool more = nitem < n_items;
      int4 nit = make_int4(0, 0, 0, 0);
      if (more) nit = __ldg(items + nitem);
      const int seg = it.z, hr = it.w;
      AccT<NV> acc;
      acc.zero();
      gather<NV, kHasW, kB>(acc, xl, ldx_bytes, col, ew, it.x, it.y, cur, lane);
      if (more) cur = load_slice<kHasW>(col, ew, nit.x, nit.y, lane);   // next segment's first slice: in flight during the fold
      store_acc<NV>(acc, hb.partial + (int64_t)seg * H, lane);
      __threadfence();  // publish this partial before taking a ticket
      // Two-level fold, each level by the last finisher and always in index order (deterministic):
      // KGB_FOLD consecutive segments -> one group partial; the row's group partials -> the output row.
      const int seg0 = __ldg(g.hrow_segptr + hr), seg1 = __ldg(g.hrow_segptr + hr + 1);
      const int gi = (seg - seg0) / KGB_FOLD;
      const int gsz = min((int)KGB_FOLD, seg1 - seg0 - gi * KGB_FOLD);
      const int grp0 = __ldg(g.hrow_grpptr + hr), ngrp = __ldg(g.hrow_grpptr + hr + 1) - grp0;
      const int gid = grp0 + gi;
      int t = 0;
      if (lane == 0) t = atomicAdd(hb.ticket1 + gid, 1);
      t = __shfl_sync(kFull, t, 0);
      if (t == gsz - 1) {
        __threadfence();
        AccT<NV> sum = fold<NV>(hb.partial + (int64_t)(seg0 + gi * KGB_FOLD) * H, gsz, lane);
        store_acc<NV>(sum, hb.gpartial + (int64_t)gid * H, lane);
        if (lane == 0) hb.ticket1[gid] = 0;  // leave the counters clean for the next launch
        __threadfence();
        if (lane == 0) t = atomicAdd(hb.ticket + hr, 1);
        t = __shfl_sync(kFull, t, 0);
        if (t == ngrp - 1) {
          __threadfence();
          sum = fold<NV>(hb.gpartial + (int64_t)grp0 * H, ngrp, lane);
          const int row = __ldg(g.hrow_id + hr);
          float* yrow = y + (int64_t)row * ldy;
          float4 old[NV];
#pragma unroll
          for (int q = 0; q < NV; ++q) old[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ep.beta != 0.f) load_old<NV>(old, yrow, lane);
          finish_row<NV>(sum, old, yrow, row, ep, lane);
          if (lane == 0) hb.ticket[hr] = 0;
        }
      }
      if (!more) break;
      item = nitem;
      it = nit;
    }
    item += n_warps;
  }

  // ------------------------------------------------------------------ rows: iteration i of this warp = row first + i*n_warps
  // (item indices continue after the segments: the warp's first row is the first element of its stride sequence
  //  that is >= n_items)
  const int first = item - n_items;
  if (first >= g.n_rows) return;
  const int cnt = (g.n_rows - 1 - first) / n_warps + 1;
  const int32_t* __restrict__ rowptr = g.rowptr;
  const int32_t* __restrict__ col_l = col + lane;         // this lane's element of a slice starting at slot 0
  const float* __restrict__ ew_l = kHasW ? ew + lane : nullptr;
  const bool has_heavy = g.n_hsegs > 0;
  const int seg_len = g.seg_len;
  // lane l holds (start, edge count) of iteration blk*32 + l; count < 0: nothing to do (heavy row, done above)
  auto block_meta = [&](int i0, int& ms, int& mn) {
    ms = 0;
    mn = -1;
    const int i = i0 + lane;
    if (i < cnt) {
      const int32_t* rp = rowptr + (first + (int64_t)i * n_warps);
      ms = __ldg(rp);
      mn = __ldg(rp + 1) - ms;
      if (has_heavy && mn > seg_len) mn = -1;
    }
  };
  auto slice_at = [&](int s, int n) {
    Slice k{0, 1.f};
    if (lane < n) {
      k.c = __ldg(col_l + s);
      if (kHasW) k.w = __ldg(ew_l + s);
    }
    return k;
  };
  // Pipeline: row pointers two 32-row blocks ahead; index/weight slices TWO rows ahead (they stream from HBM and take
  // longer than the gathers, which mostly hit L1/L2); the old output row at the top of its own iteration.
  int ms, mn, ms_nx, mn_nx;
  block_meta(0, ms, mn);
  block_meta(32, ms_nx, mn_nx);
  int s0 = __shfl_sync(kFull, ms, 0), n0 = __shfl_sync(kFull, mn, 0);
  int s1 = 0, n1 = -1;
  Slice c0 = slice_at(s0, n0), c1{0, 1.f};
  if (kDist == 2) {
    s1 = __shfl_sync(kFull, ms, 1);
    n1 = __shfl_sync(kFull, mn, 1);
    c1 = slice_at(s1, n1);
  }
  float* __restrict__ yrow = y + (int64_t)first * ldy;
  const int64_t ystep = (int64_t)n_warps * ldy;
  int row = first;
  for (int i = 0; i < cnt; ++i, yrow += ystep, row += n_warps) {
    // ---- requests for later iterations first: row pointers of row i+kDist (from the lane-held blocks), its slice
    int sN, nN;
    const int k = (i + kDist) & 31;
    if (k < kDist) {                   // that row lives in the next 32-row block
      sN = __shfl_sync(kFull, ms_nx, k);
      nN = __shfl_sync(kFull, mn_nx, k);
    } else {
      sN = __shfl_sync(kFull, ms, k);
      nN = __shfl_sync(kFull, mn, k);
    }
    if (k == kDist - 1) {              // every row of the old block that was still needed is out: rotate, request another
      ms = ms_nx;
      mn = mn_nx;
      block_meta(i + 1 + 32, ms_nx, mn_nx);
    }
    const Slice cN = slice_at(sN, nN);                            // nN < 0 (or past the end): no loads
    if (n0 >= 0) {
      float4 old[NV];
#pragma unroll
      for (int q = 0; q < NV; ++q) old[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kOldEarly && ep.beta != 0.f) load_old<NV>(old, yrow, lane);
      AccT<NV> acc;
      acc.zero();
      if (n0 > 0) gather<NV, kHasW, kB>(acc, xl, ldx_bytes, col, ew, s0, n0, c0, lane);
      if (!kOldEarly && ep.beta != 0.f) load_old<NV>(old, yrow, lane);
      finish_row<NV>(acc, old, yrow, row, ep, lane);
    }
    if (kDist == 2) {
      s0 = s1; n0 = n1; c0 = c1;
      s1 = sN; n1 = nN; c1 = cN;
    } else {
      s0 = sN; n0 = nN; c0 = cN;
    }
  }
}

}  // namespace lean
}  // namespace kgb
