// fp32 FFMA GEMM for the dense contractions the tensor-core path does not take (small or odd
// shapes, and the split-K weight-gradient reduction).  128x128x16 CTA tile, 8x8 per thread.
//   NT: C = A[M,K] . B[N,K]^T      NN: C = A[M,K] . B[K,N]      TN: C = A[K,M]^T . B[K,N]
#include "kgb_common.cuh"

namespace kgb {

constexpr int BM = 128, BN = 128, BK = 16, GT = 256, PAD = 4;

// operand whose contiguous axis is k:  p[row*ld + k]   -> smem[k][row]
__device__ __forceinline__ void load_kcontig(const float* __restrict__ p, int64_t ld, int64_t row0, int64_t n_rows,
                                             int64_t k0, int64_t K, float4 (&r)[2], int tid) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int row = (tid >> 2) + 64 * i;
    const int kq = (tid & 3) * 4;
    const int64_t gr = row0 + row, gk = k0 + kq;
    if (gr < n_rows && gk < K) r[i] = __ldg(reinterpret_cast<const float4*>(p + gr * ld + gk));  // K % 4 == 0
    else r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ void store_kcontig(float (*s)[BM + PAD], const float4 (&r)[2], int tid) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int row = (tid >> 2) + 64 * i;
    const int kq = (tid & 3) * 4;
    s[kq + 0][row] = r[i].x; s[kq + 1][row] = r[i].y; s[kq + 2][row] = r[i].z; s[kq + 3][row] = r[i].w;
  }
}
// operand whose contiguous axis is the tile row (m or n):  p[k*ld + col]  -> smem[k][col]
__device__ __forceinline__ void load_rcontig(const float* __restrict__ p, int64_t ld, int64_t col0, int64_t n_cols,
                                             int64_t k0, int64_t kend, float4 (&r)[2], int tid) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int k = (tid >> 5) + 8 * i;
    const int cq = (tid & 31) * 4;
    const int64_t gk = k0 + k, gc = col0 + cq;
    if (gk < kend && gc < n_cols) r[i] = __ldg(reinterpret_cast<const float4*>(p + gk * ld + gc));  // n_cols % 4 == 0
    else r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ void store_rcontig(float (*s)[BM + PAD], const float4 (&r)[2], int tid) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int k = (tid >> 5) + 8 * i;
    const int cq = (tid & 31) * 4;
    *reinterpret_cast<float4*>(&s[k][cq]) = r[i];
  }
}

// When SPLIT: gridDim.z slices of K write raw partials to `part` [z][M][N]; otherwise full epilogue.
template <int LAYOUT, bool SPLIT>
__global__ void __launch_bounds__(GT, 2)
k_gemm_ffma(const float* __restrict__ a, int64_t lda, const float* __restrict__ b, int64_t ldb, float* __restrict__ c,
            int64_t ldc, int64_t M, int64_t N, int64_t K, float alpha, float beta, const float* __restrict__ bias,
            int relu, float* __restrict__ part, int64_t k_per_split) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  const int64_t kbeg = SPLIT ? (int64_t)blockIdx.z * k_per_split : 0;
  const int64_t kend = SPLIT ? min(K, kbeg + k_per_split) : K;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto fetch = [&](int64_t k0) {
    if constexpr (LAYOUT == KGB_TN) load_rcontig(a, lda, m0, M, k0, kend, ra, tid);
    else load_kcontig(a, lda, m0, M, k0, kend, ra, tid);
    if constexpr (LAYOUT == KGB_NT) load_kcontig(b, ldb, n0, N, k0, kend, rb, tid);
    else load_rcontig(b, ldb, n0, N, k0, kend, rb, tid);
  };
  auto commit = [&]() {
    if constexpr (LAYOUT == KGB_TN) store_rcontig(As, ra, tid); else store_kcontig(As, ra, tid);
    if constexpr (LAYOUT == KGB_NT) store_kcontig(Bs, rb, tid); else store_rcontig(Bs, rb, tid);
  };

  if (kbeg < kend) fetch(kbeg);
  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
    commit();
    __syncthreads();
    if (k0 + BK < kend) fetch(k0 + BK);  // register prefetch of the next k-tile overlaps the FMAs
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int64_t gn = n0 + jh * 64 + tx * 4;
      if (gn >= N) continue;  // N % 4 == 0
      float4 v = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
      if constexpr (SPLIT) {
        *reinterpret_cast<float4*>(part + ((int64_t)blockIdx.z * M + gm) * N + gn) = v;
      } else {
        v.x *= alpha; v.y *= alpha; v.z *= alpha; v.w *= alpha;
        float* cp = c + gm * ldc + gn;
        if (beta != 0.f) {
          const float4 o = *reinterpret_cast<const float4*>(cp);
          v.x = fmaf(beta, o.x, v.x); v.y = fmaf(beta, o.y, v.y); v.z = fmaf(beta, o.z, v.z); v.w = fmaf(beta, o.w, v.w);
        }
        if (bias) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + gn));
          v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
        }
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        *reinterpret_cast<float4*>(cp) = v;
      }
    }
  }
}

// C = act(alpha * sum_z part[z] + beta*C + bias).  256 threads = 64 float4 outputs x 4 slice lanes: lane q sums the
// slices z = q, q+4, ... (4 loads in flight), the 4 lane sums are folded in lane order => deterministic.
__global__ void __launch_bounds__(256)
k_splitk_reduce(const float* __restrict__ part, int splits, int64_t M, int64_t N, float* __restrict__ c,
                int64_t ldc, float alpha, float beta, const float* __restrict__ bias, int relu) {
  __shared__ float4 red[4][64];
  const int o = threadIdx.x & 63, q = threadIdx.x >> 6;
  const int64_t i = (int64_t)blockIdx.x * 64 + o;  // float4 index
  const int64_t nq = N / 4;
  const bool live = i < M * nq;
  const int64_t m = live ? i / nq : 0, n = live ? (i % nq) * 4 : 0;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) {
    const float* p = part + m * N + n;
    const int64_t zs = M * N;
    int z = q;
    for (; z + 12 < splits; z += 16) {
      const float4 p0 = *reinterpret_cast<const float4*>(p + (int64_t)z * zs);
      const float4 p1 = *reinterpret_cast<const float4*>(p + (int64_t)(z + 4) * zs);
      const float4 p2 = *reinterpret_cast<const float4*>(p + (int64_t)(z + 8) * zs);
      const float4 p3 = *reinterpret_cast<const float4*>(p + (int64_t)(z + 12) * zs);
      s.x += p0.x; s.y += p0.y; s.z += p0.z; s.w += p0.w;
      s.x += p1.x; s.y += p1.y; s.z += p1.z; s.w += p1.w;
      s.x += p2.x; s.y += p2.y; s.z += p2.z; s.w += p2.w;
      s.x += p3.x; s.y += p3.y; s.z += p3.z; s.w += p3.w;
    }
    for (; z < splits; z += 4) {
      const float4 p0 = *reinterpret_cast<const float4*>(p + (int64_t)z * zs);
      s.x += p0.x; s.y += p0.y; s.z += p0.z; s.w += p0.w;
    }
  }
  red[q][o] = s;
  __syncthreads();
  if (q != 0 || !live) return;
  s = red[0][o];
#pragma unroll
  for (int k = 1; k < 4; ++k) { s.x += red[k][o].x; s.y += red[k][o].y; s.z += red[k][o].z; s.w += red[k][o].w; }
  s.x *= alpha; s.y *= alpha; s.z *= alpha; s.w *= alpha;
  float* cp = c + m * ldc + n;
  if (beta != 0.f) {
    const float4 o4 = *reinterpret_cast<const float4*>(cp);
    s.x = fmaf(beta, o4.x, s.x); s.y = fmaf(beta, o4.y, s.y); s.z = fmaf(beta, o4.z, s.z); s.w = fmaf(beta, o4.w, s.w);
  }
  if (bias) { s.x += bias[n]; s.y += bias[n + 1]; s.z += bias[n + 2]; s.w += bias[n + 3]; }
  if (relu) { s.x = fmaxf(s.x, 0.f); s.y = fmaxf(s.y, 0.f); s.z = fmaxf(s.z, 0.f); s.w = fmaxf(s.w, 0.f); }
  *reinterpret_cast<float4*>(cp) = s;
}

int ffma_splits(int64_t m, int64_t n, int64_t k) {
  const int64_t tiles = ((m + BM - 1) / BM) * ((n + BN - 1) / BN);
  int64_t s = (2 * kNumSMs + tiles - 1) / tiles;
  const int64_t max_s = (k + 8 * BK - 1) / (8 * BK);  // at least 128 k per slice
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  return (int)s;
}

size_t gemm_ffma_workspace_bytes(int layout, int64_t m, int64_t n, int64_t k) {
  if (layout != KGB_TN) return 0;
  const int s = ffma_splits(m, n, k);
  return s > 1 ? (size_t)s * m * n * sizeof(float) + 256 : 0;
}

int gemm_ffma(int layout, const float* a, int64_t lda, const float* b, int64_t ldb, float* c, int64_t ldc, int64_t M,
              int64_t N, int64_t K, float alpha, float beta, const float* bias, int relu, void* ws, size_t ws_bytes,
              cudaStream_t stream) {
  dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM), 1);
  if (layout == KGB_NT) {
    k_gemm_ffma<KGB_NT, false><<<grid, GT, 0, stream>>>(a, lda, b, ldb, c, ldc, M, N, K, alpha, beta, bias, relu, nullptr, 0);
  } else if (layout == KGB_NN) {
    k_gemm_ffma<KGB_NN, false><<<grid, GT, 0, stream>>>(a, lda, b, ldb, c, ldc, M, N, K, alpha, beta, bias, relu, nullptr, 0);
  } else {
    const int s = ffma_splits(M, N, K);
    if (s <= 1) {
      k_gemm_ffma<KGB_TN, false><<<grid, GT, 0, stream>>>(a, lda, b, ldb, c, ldc, M, N, K, alpha, beta, bias, relu, nullptr, 0);
    } else {
      if (!ws || ws_bytes < gemm_ffma_workspace_bytes(layout, M, N, K)) {
        set_error("gemm: workspace %zu < %zu", ws_bytes, gemm_ffma_workspace_bytes(layout, M, N, K));
        return KGB_ERR_WORKSPACE;
      }
      int64_t kps = (K + s - 1) / s;
      kps = (kps + BK - 1) / BK * BK;
      grid.z = (unsigned)((K + kps - 1) / kps);
      float* part = static_cast<float*>(ws);
      k_gemm_ffma<KGB_TN, true><<<grid, GT, 0, stream>>>(a, lda, b, ldb, c, ldc, M, N, K, alpha, beta, bias, relu, part, kps);
      KGB_LAUNCH_OK();
      const int64_t n4 = M * N / 4;
      k_splitk_reduce<<<(unsigned)((n4 + 63) / 64), 256, 0, stream>>>(part, (int)grid.z, M, N, c, ldc, alpha, beta, bias, relu);
    }
  }
  KGB_LAUNCH_OK();
  return KGB_OK;
}

}  // namespace kgb
