// Shared helpers for libkgwas_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "kgwas_b200.h"

namespace kgb {

constexpr int kWarp = 32;
constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// thread-local error text behind kgb_last_error()
char* err_buf();
void set_error(const char* fmt, ...);
void count_launch();  // every kernel this library launches bumps kgb_launch_count()

#define KGB_CUDA_OK(expr)                                                              \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      kgb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return KGB_ERR_CUDA;                                                             \
    }                                                                                  \
  } while (0)

#define KGB_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      kgb::set_error(__VA_ARGS__);      \
      return KGB_ERR_INVALID;           \
    }                                   \
  } while (0)

#define KGB_LAUNCH_OK()                                                                  \
  do {                                                                                   \
    kgb::count_launch();                                                                 \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      kgb::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return KGB_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// carve a workspace into 256B-aligned pieces
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t n) {
    T* r = reinterpret_cast<T*>(base + off);
    off += align_up(n * sizeof(T), 256);
    return r;
  }
};

// ---- device helpers --------------------------------------------------------------------

// Read-only, L1-bypassing 128-bit load for streamed (touched-once) data.
__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
// Read-only 128-bit load that is allowed to stay in L1/L2 (gathered rows are re-used).
__device__ __forceinline__ float4 ldg_f4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void st_f4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// A feature row of H floats spread over one warp.  H % 128 == 0: lane owns float4 chunks at
// [c*128 + lane*4] (each warp-wide access is one fully coalesced 512 B request).
// H in {32, 64}: lane owns H/32 consecutive floats.
template <int H>
struct RowVec {
  static constexpr int N = H / 32;
  float v[N];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = 0.f;
  }
  __device__ __forceinline__ void load(const float* row, int lane) {
    if constexpr (H % 128 == 0) {
#pragma unroll
      for (int c = 0; c < H / 128; ++c) {
        float4 t = ldg_f4(row + c * 128 + lane * 4);
        v[4 * c] = t.x; v[4 * c + 1] = t.y; v[4 * c + 2] = t.z; v[4 * c + 3] = t.w;
      }
    } else if constexpr (N == 2) {
      float2 t = __ldg(reinterpret_cast<const float2*>(row + lane * 2));
      v[0] = t.x; v[1] = t.y;
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] = __ldg(row + lane * N + i);
    }
  }
  __device__ __forceinline__ void load_stream(const float* row, int lane) {  // touched once: do not keep in L1
    if constexpr (H % 128 == 0) {
#pragma unroll
      for (int c = 0; c < H / 128; ++c) {
        float4 t = ldg_stream_f4(row + c * 128 + lane * 4);
        v[4 * c] = t.x; v[4 * c + 1] = t.y; v[4 * c + 2] = t.z; v[4 * c + 3] = t.w;
      }
    } else {
      load(row, lane);
    }
  }
  __device__ __forceinline__ void load_plain(const float* row, int lane) {  // coherent load (data written in-kernel)
    if constexpr (H % 128 == 0) {
#pragma unroll
      for (int c = 0; c < H / 128; ++c) {
        float4 t = __ldcg(reinterpret_cast<const float4*>(row + c * 128 + lane * 4));
        v[4 * c] = t.x; v[4 * c + 1] = t.y; v[4 * c + 2] = t.z; v[4 * c + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] = __ldcg(row + lane * N + i);
    }
  }
  __device__ __forceinline__ void store(float* row, int lane) const {
    if constexpr (H % 128 == 0) {
#pragma unroll
      for (int c = 0; c < H / 128; ++c)
        st_f4(row + c * 128 + lane * 4, make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]));
    } else if constexpr (N == 2) {
      *reinterpret_cast<float2*>(row + lane * 2) = make_float2(v[0], v[1]);
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) row[lane * N + i] = v[i];
    }
  }
  __device__ __forceinline__ void fma(float w, const RowVec& x) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = fmaf(w, x.v[i], v[i]);
  }
  __device__ __forceinline__ void add(const RowVec& x) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] += x.v[i];
  }
  __device__ __forceinline__ float dot(const RowVec& x) const {  // per-lane partial
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) s = fmaf(v[i], x.v[i], s);
    return s;
  }
};

// dispatch a runtime feature width to a compile-time one
#define KGB_DISPATCH_H(h, ...)                                             \
  switch (h) {                                                             \
    case 32:  { constexpr int H = 32;  __VA_ARGS__; break; }               \
    case 64:  { constexpr int H = 64;  __VA_ARGS__; break; }               \
    case 128: { constexpr int H = 128; __VA_ARGS__; break; }               \
    case 256: { constexpr int H = 256; __VA_ARGS__; break; }               \
    case 384: { constexpr int H = 384; __VA_ARGS__; break; }               \
    case 512: { constexpr int H = 512; __VA_ARGS__; break; }               \
    default:                                                               \
      kgb::set_error("feature width %d unsupported (need 32,64,128,256,384,512)", (int)(h)); \
      return KGB_ERR_UNSUPPORTED;                                          \
  }

}  // namespace kgb
