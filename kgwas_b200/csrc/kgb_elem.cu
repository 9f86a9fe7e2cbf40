// Small fused elementwise / reduction helpers around the aggregation kernels.
#include "kgb_common.cuh"

namespace kgb {

// g = dy * (y > 0)
__global__ void k_relu_bwd(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ g, int64_t n) {
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = ldg_stream_f4(dy + 4 * i), b = ldg_stream_f4(y + 4 * i);
    st_f4(g + 4 * i, make_float4(b.x > 0.f ? a.x : 0.f, b.y > 0.f ? a.y : 0.f, b.z > 0.f ? a.z : 0.f, b.w > 0.f ? a.w : 0.f));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    g[i] = y[i] > 0.f ? dy[i] : 0.f;
  }
}

// Weighted column sums, stage 1:  part[b][r][:] = sum_{m in chunk b} w[m, r] * x[m, :]   (w == NULL -> 1, R = 1)
constexpr int kColsumThreads = 256;
template <int H, int R>
__global__ void __launch_bounds__(kColsumThreads)
k_wcolsum_stage1(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w, int64_t ldw, int64_t M,
                 int64_t rows_per_cta, float* __restrict__ part) {
  __shared__ float red[kColsumThreads / 32][H];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = min(M, r0 + rows_per_cta);
  RowVec<H> acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r].zero();
  for (int64_t m = r0 + warp; m < r1; m += kColsumThreads / 32) {
    RowVec<H> t;
    t.load(x + m * ldx, lane);
    if (w) {
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r].fma(__ldg(w + m * ldw + r), t);
    } else {
      acc[0].add(t);
    }
  }
  // fold the 8 warps in warp order (fixed => deterministic)
#pragma unroll
  for (int r = 0; r < R; ++r) {
    __syncthreads();
    acc[r].store(&red[warp][0], lane);
    __syncthreads();
    for (int c = threadIdx.x; c < H; c += kColsumThreads) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < kColsumThreads / 32; ++q) s += red[q][c];
      part[((int64_t)blockIdx.x * R + r) * H + c] = s;
    }
  }
}
// Fused backward of (ReLU -> single-output head) + bias column sums; see kgb_relu_bwd_fused in kgwas_b200.h.
// Two rows per warp and iteration in flight; part[b][0][:] = column sums of g, part[b][1][:] = sum dp[i] * y[i,:].
template <int H>
__global__ void __launch_bounds__(kColsumThreads)
k_relu_bwd_fused(const float* __restrict__ dy, int64_t lddy, const float* __restrict__ y, int64_t ldy,
                 const float* __restrict__ dp, const float* __restrict__ wv, float scale, float* __restrict__ g,
                 int64_t ldg, int64_t M, int64_t rows_per_cta, float* __restrict__ part) {
  __shared__ float red[kColsumThreads / 32][H];
  constexpr int NW = kColsumThreads / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = min(M, r0 + rows_per_cta);
  RowVec<H> acc_g, acc_w, wvec;
  acc_g.zero();
  acc_w.zero();
  wvec.zero();
  if (dp) wvec.load(wv, lane);
  auto load_row = [&](int64_t m, RowVec<H>& d, RowVec<H>& yv, float& p) {
    d.zero();
    yv.zero();
    p = 0.f;
    if (m < r1) {
      if (dy) d.load_stream(dy + m * lddy, lane);
      if (y) yv.load_stream(y + m * ldy, lane);
      if (dp) p = __ldg(dp + m);
    }
  };
  auto finish_row = [&](int64_t m, RowVec<H>& d, const RowVec<H>& yv, float p) {
    if (m >= r1) return;
    if (dp) d.fma(p, wvec);
#pragma unroll
    for (int i = 0; i < RowVec<H>::N; ++i) {
      const bool on = y ? yv.v[i] > 0.f : true;
      d.v[i] = on ? d.v[i] * scale : 0.f;
    }
    d.store(g + m * ldg, lane);
    acc_g.add(d);
    if (dp && y) acc_w.fma(p, yv);
  };
  for (int64_t m = r0 + warp; m < r1; m += 2 * NW) {
    RowVec<H> d0, y0, d1, y1;
    float p0, p1;
    load_row(m, d0, y0, p0);
    load_row(m + NW, d1, y1, p1);
    finish_row(m, d0, y0, p0);
    finish_row(m + NW, d1, y1, p1);
  }
  if (!part) return;
  // fold the 8 warps in warp order (fixed => deterministic)
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    __syncthreads();
    (r == 0 ? acc_g : acc_w).store(&red[warp][0], lane);
    __syncthreads();
    for (int c = threadIdx.x; c < H; c += kColsumThreads) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < NW; ++q) s += red[q][c];
      part[((int64_t)blockIdx.x * 2 + r) * H + c] = s;
    }
  }
}

// stage 2: out[i] = beta*out[i] + sum_b part[b][i].  One warp per output element: lanes stride over the chunks
// (fixed lane <-> chunk mapping, fixed shuffle tree => deterministic), 8 outputs per CTA.
__global__ void k_wcolsum_stage2(const float* __restrict__ part, int n_ctas, int RH, float* __restrict__ out, float beta) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= RH) return;
  float s = 0.f;
  for (int b = lane; b < n_ctas; b += 32) s += part[(int64_t)b * RH + i];
  s = warp_sum(s);
  if (lane == 0) out[i] = (beta != 0.f ? beta * out[i] : 0.f) + s;
}

// a[i, s] = <x[i, :h], v[s, :]> for s < n_slots <= 8 (slot_stride == 0: one input row, several vectors -- the folded
// attention logits of GATConv, kgwas/conv.py:150-151).  HBM-bound (one 4h-byte row per 4*n_slots bytes of output), so:
// the vectors live in registers, two rows are in flight per warp, and the n_slots lane-partials are reduced by a
// transposing butterfly (lane^16 keeps half of the values, lane^8 a quarter, lane^4 one; then two plain steps):
// 9 shuffles per row instead of 5 per slot.  Lane l with (l & 3) == 0 ends up with slot ((l>>4)&1)*4 + ((l>>3)&1)*2 + ((l>>2)&1).
template <int H>
__global__ void __launch_bounds__(256) k_rowdot_multi(const float* __restrict__ x, int64_t ldx, int64_t n_rows, int n_slots,
                                                      const float* __restrict__ v, float* __restrict__ a, int64_t lda) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  RowVec<H> vs[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    vs[s].zero();
    if (s < n_slots) vs[s].load(v + (int64_t)s * H, lane);
  }
  const int my_slot = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  auto reduce_store = [&](const RowVec<H>& xr, int64_t row) {
    float p[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) p[s] = xr.dot(vs[s]);
    float q[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {                       // lane^16: lower half keeps p[0..3], upper half p[4..7]
      const float send = (lane & 16) ? p[i] : p[i + 4];
      const float keep = (lane & 16) ? p[i + 4] : p[i];
      q[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    float r[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = (lane & 8) ? q[i] : q[i + 2];
      const float keep = (lane & 8) ? q[i + 2] : q[i];
      r[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    float t = ((lane & 4) ? r[1] : r[0]) + __shfl_xor_sync(0xffffffffu, (lane & 4) ? r[0] : r[1], 4);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    if ((lane & 3) == 0 && my_slot < n_slots) a[row * lda + my_slot] = t;
  };
  for (int64_t row = warp0; row < n_rows; row += 2 * n_warps) {
    RowVec<H> x0, x1;
    x0.load_stream(x + row * ldx, lane);
    const bool two = row + n_warps < n_rows;
    if (two) x1.load_stream(x + (row + n_warps) * ldx, lane);
    reduce_store(x0, row);
    if (two) reduce_store(x1, row + n_warps);
  }
}

// a[i, s] = <x[i, s*slot_stride : +h], v[s, :]>
template <int H>
__global__ void k_rowdot(const float* __restrict__ x, int64_t ldx, int64_t n_rows, int n_slots, int64_t slot_stride,
                         const float* __restrict__ v, float* __restrict__ a, int64_t lda) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp0; row < n_rows; row += n_warps) {
    RowVec<H> xr;
    if (slot_stride == 0) xr.load(x + row * ldx, lane);
    for (int s = 0; s < n_slots; ++s) {
      if (slot_stride != 0) xr.load(x + row * ldx + s * slot_stride, lane);
      RowVec<H> vs;
      vs.load(v + (int64_t)s * H, lane);
      const float d = warp_sum(xr.dot(vs));
      if (lane == 0) a[row * lda + s] = d;
    }
  }
}

// y[i, :] = beta*y[i, :] + sum_s a[i, s] * v[s, :]      (rank-n_slots update; n_slots small)
template <int H>
__global__ void k_rank_update(const float* __restrict__ a, int64_t lda, int n_slots, const float* __restrict__ v,
                              float* __restrict__ y, int64_t ldy, int64_t n_rows, float beta) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp0; row < n_rows; row += n_warps) {
    RowVec<H> acc;
    acc.zero();
    for (int s = 0; s < n_slots; ++s) {
      RowVec<H> vs;
      vs.load(v + (int64_t)s * H, lane);
      acc.fma(__ldg(a + row * lda + s), vs);
    }
    if (beta != 0.f) {
      RowVec<H> old;
      old.load_plain(y + row * ldy, lane);
#pragma unroll
      for (int i = 0; i < RowVec<H>::N; ++i) acc.v[i] = fmaf(beta, old.v[i], acc.v[i]);
    }
    acc.store(y + row * ldy, lane);
  }
}

__global__ void k_permute_f32(const float* __restrict__ w, const int32_t* __restrict__ perm, float* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __ldg(w + perm[i]);
}

inline unsigned warp_grid(int64_t n_warp_items, int threads) {
  int64_t ctas = (n_warp_items + threads / 32 - 1) / (threads / 32);
  const int64_t cap = (int64_t)kNumSMs * 32;
  if (ctas > cap) ctas = cap;
  return (unsigned)(ctas < 1 ? 1 : ctas);
}

static int64_t colsum_ctas(int64_t m) {
  int64_t c = (m + 63) / 64;
  if (c > 4 * kNumSMs) c = 4 * kNumSMs;
  return c < 1 ? 1 : c;
}

}  // namespace kgb

using namespace kgb;

extern "C" int kgb_relu_bwd(const float* dy, const float* y, float* g, int64_t n, kgb_stream_t stream_) {
  if (n == 0) return KGB_OK;
  KGB_REQUIRE(dy && y && g && n > 0, "relu_bwd: bad argument");
  KGB_REQUIRE(aligned16(dy) && aligned16(y) && aligned16(g), "relu_bwd: pointers must be 16-byte aligned");
  int64_t ctas = (n / 4 + 255) / 256;
  if (ctas > kNumSMs * 16) ctas = kNumSMs * 16;
  if (ctas < 1) ctas = 1;
  k_relu_bwd<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream_>>>(dy, y, g, n);
  KGB_LAUNCH_OK();
  return KGB_OK;
}

extern "C" size_t kgb_relu_bwd_fused_workspace_bytes(int64_t m, int32_t h) {
  return (size_t)colsum_ctas(m) * 2 * h * sizeof(float) + 256;
}

extern "C" int kgb_relu_bwd_fused(const float* dy, int64_t lddy, const float* y, int64_t ldy, const float* dp,
                                  const float* wv, float scale, float* g, int64_t ldg, int64_t m, int32_t h,
                                  float* sums, void* workspace, size_t workspace_bytes, kgb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (m == 0) {
    if (sums) KGB_CUDA_OK(cudaMemsetAsync(sums, 0, (size_t)2 * h * sizeof(float), stream));
    return KGB_OK;
  }
  KGB_REQUIRE(g && m > 0 && (dy || dp), "relu_bwd_fused: bad argument");
  KGB_REQUIRE((dp == nullptr) == (wv == nullptr), "relu_bwd_fused: dp and wv go together");
  KGB_REQUIRE(aligned16(g) && ldg % 4 == 0 && (!dy || (aligned16(dy) && lddy % 4 == 0)) &&
                  (!y || (aligned16(y) && ldy % 4 == 0)) && (!wv || aligned16(wv)),
              "relu_bwd_fused: pointers must be 16-byte aligned, strides multiples of 4");
  if (sums && (workspace_bytes < kgb_relu_bwd_fused_workspace_bytes(m, h) || !workspace)) {
    set_error("relu_bwd_fused: workspace too small");
    return KGB_ERR_WORKSPACE;
  }
  const int64_t ctas = colsum_ctas(m);
  const int64_t rows_per_cta = (m + ctas - 1) / ctas;
  float* part = sums ? static_cast<float*>(workspace) : nullptr;
  KGB_DISPATCH_H(h, (k_relu_bwd_fused<H><<<(unsigned)ctas, kColsumThreads, 0, stream>>>(dy, lddy, y, ldy, dp, wv, scale, g,
                                                                                     ldg, m, rows_per_cta, part)));
  KGB_LAUNCH_OK();
  if (sums) {
    const int RH = 2 * h;
    k_wcolsum_stage2<<<(RH + 7) / 8, 256, 0, stream>>>(part, (int)ctas, RH, sums, 0.f);
    KGB_LAUNCH_OK();
  }
  return KGB_OK;
}

extern "C" size_t kgb_wcolsum_workspace_bytes(int64_t m, int32_t n_slots, int32_t h) {
  return (size_t)colsum_ctas(m) * (size_t)(n_slots > 0 ? n_slots : 1) * h * sizeof(float) + 256;
}

extern "C" int kgb_wcolsum(const float* x, int64_t ldx, const float* w, int64_t ldw, int64_t m, int32_t n_slots,
                           int32_t h, float* out, float beta, void* workspace, size_t workspace_bytes,
                           kgb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  KGB_REQUIRE(x && out && m >= 0 && n_slots >= 1, "wcolsum: bad argument");
  KGB_REQUIRE(w != nullptr || n_slots == 1, "wcolsum: n_slots > 1 needs weights");
  KGB_REQUIRE(n_slots <= 8, "wcolsum: at most 8 slots per call");
  KGB_REQUIRE(aligned16(x) && ldx % 4 == 0, "wcolsum: x alignment");
  if (workspace_bytes < kgb_wcolsum_workspace_bytes(m, n_slots, h) || !workspace) {
    set_error("wcolsum: workspace too small");
    return KGB_ERR_WORKSPACE;
  }
  const int64_t ctas = colsum_ctas(m);
  const int64_t rows_per_cta = (m + ctas - 1) / ctas;
  float* part = static_cast<float*>(workspace);
#define KGB_WCS(RR) k_wcolsum_stage1<H, RR><<<(unsigned)ctas, kColsumThreads, 0, stream>>>(x, ldx, w, ldw, m, rows_per_cta, part)
  KGB_DISPATCH_H(h, {
    switch (n_slots) {
      case 1: KGB_WCS(1); break; case 2: KGB_WCS(2); break; case 3: KGB_WCS(3); break; case 4: KGB_WCS(4); break;
      case 5: KGB_WCS(5); break; case 6: KGB_WCS(6); break; case 7: KGB_WCS(7); break; default: KGB_WCS(8); break;
    }
  });
#undef KGB_WCS
  KGB_LAUNCH_OK();
  const int RH = n_slots * h;
  k_wcolsum_stage2<<<(RH + 7) / 8, 256, 0, stream>>>(part, (int)ctas, RH, out, beta);
  KGB_LAUNCH_OK();
  return KGB_OK;
}

extern "C" int kgb_rowdot(const float* x, int64_t ldx, int64_t n_rows, int32_t n_slots, int32_t h, int64_t slot_stride,
                          const float* v, float* a, int64_t lda, kgb_stream_t stream_) {
  if (n_rows == 0) return KGB_OK;
  KGB_REQUIRE(x && v && a && n_slots >= 1 && lda >= n_slots, "rowdot: bad argument");
  KGB_REQUIRE(aligned16(x) && aligned16(v) && ldx % 4 == 0 && slot_stride % 4 == 0, "rowdot: alignment");
  if (slot_stride == 0 && n_slots <= 8 && h <= 256) {
    const unsigned grid = warp_grid((n_rows + 1) / 2, 256);
    switch (h) {
      case 32: k_rowdot_multi<32><<<grid, 256, 0, (cudaStream_t)stream_>>>(x, ldx, n_rows, n_slots, v, a, lda); break;
      case 64: k_rowdot_multi<64><<<grid, 256, 0, (cudaStream_t)stream_>>>(x, ldx, n_rows, n_slots, v, a, lda); break;
      case 128: k_rowdot_multi<128><<<grid, 256, 0, (cudaStream_t)stream_>>>(x, ldx, n_rows, n_slots, v, a, lda); break;
      case 256: k_rowdot_multi<256><<<grid, 256, 0, (cudaStream_t)stream_>>>(x, ldx, n_rows, n_slots, v, a, lda); break;
      default: set_error("rowdot: feature width %d unsupported", (int)h); return KGB_ERR_UNSUPPORTED;
    }
    KGB_LAUNCH_OK();
    return KGB_OK;
  }
  KGB_DISPATCH_H(h, (k_rowdot<H><<<warp_grid(n_rows, 256), 256, 0, (cudaStream_t)stream_>>>(x, ldx, n_rows, n_slots,
                                                                                         slot_stride, v, a, lda)));
  KGB_LAUNCH_OK();
  return KGB_OK;
}

extern "C" int kgb_rank_update(const float* a, int64_t lda, int32_t n_slots, const float* v, float* y, int64_t ldy,
                               int64_t n_rows, int32_t h, float beta, kgb_stream_t stream_) {
  if (n_rows == 0) return KGB_OK;
  KGB_REQUIRE(a && v && y && n_slots >= 1, "rank_update: bad argument");
  KGB_REQUIRE(aligned16(v) && aligned16(y) && ldy % 4 == 0, "rank_update: alignment");
  KGB_DISPATCH_H(h, (k_rank_update<H><<<warp_grid(n_rows, 256), 256, 0, (cudaStream_t)stream_>>>(a, lda, n_slots, v, y,
                                                                                              ldy, n_rows, beta)));
  KGB_LAUNCH_OK();
  return KGB_OK;
}

extern "C" int kgb_permute_f32(const float* w, const int32_t* perm, float* out, int64_t n, kgb_stream_t stream_) {
  if (n == 0) return KGB_OK;
  KGB_REQUIRE(w && perm && out, "permute_f32: null pointer");
  k_permute_f32<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(w, perm, out, n);
  KGB_LAUNCH_OK();
  return KGB_OK;
}
