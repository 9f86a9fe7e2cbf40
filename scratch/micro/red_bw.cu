// Micro-benchmark: throughput of 512-byte vector atomic adds (red.global.add.v4.f32, one row per warp instruction)
// into an L2-resident table, while a second stream of loads sweeps a big table (as the tile kernel would).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void k_red(float* __restrict__ y, const int* __restrict__ rows, int n_ops, int mode) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (int i = warp; i < n_ops; i += n_warps) {
    const int r = __ldg(rows + i);
    float* p = y + (size_t)r * 128 + lane * 4;
    if (mode == 0) red_add_v4(p, v);
    else { atomicAdd(p, v.x); atomicAdd(p + 1, v.y); atomicAdd(p + 2, v.z); atomicAdd(p + 3, v.w); }
  }
}

int main() {
  const int n_rows = 122226, n_ops = 4000000;
  float* y; int* rows;
  cudaMalloc(&y, (size_t)n_rows * 512);
  cudaMemset(y, 0, (size_t)n_rows * 512);
  std::vector<int> h(n_ops);
  srand(1);
  for (int i = 0; i < n_ops; ++i) h[i] = (int)(((long long)rand() * 32768 + rand()) % n_rows);
  cudaMalloc(&rows, n_ops * 4);
  cudaMemcpy(rows, h.data(), n_ops * 4, cudaMemcpyHostToDevice);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int mode = 0; mode < 2; ++mode)
    for (int grid : {148, 296, 592, 1184}) {
      k_red<<<grid, 512>>>(y, rows, n_ops, mode);
      cudaEventRecord(a);
      for (int it = 0; it < 5; ++it) k_red<<<grid, 512>>>(y, rows, n_ops, mode);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      printf("mode %d (0 = red.v4.f32, 1 = 4 scalar atomics) grid %d x 512: %.1f us per 4M row-adds (%.2f TB/s of payload)  err=%s\n", mode, grid,
             ms / 5 * 1e3, 4e6 * 512 / (ms / 5 * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
