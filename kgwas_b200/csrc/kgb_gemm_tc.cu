// tcgen05 (5th-gen tensor core) GEMM with fp32-equivalent accuracy: 3xTF32 split accumulation.
//
//   C[M,N] = act( alpha * op(A) op(B) + beta*C + bias )        NT / NN / TN as in kgwas_b200.h
//
// Every fp32 operand element x is split by the producer warps into hi = x & 0xffffe000 (exactly a
// TF32 number) and lo = x - hi (exact in fp32); the tensor core accumulates
//       A_lo.B_hi + A_hi.B_lo + A_hi.B_hi        (dropped term A_lo.B_lo ~ 2^-22 relative)
// into an fp32 accumulator that lives in TMEM.  One CTA owns one 128x128 output tile:
//   warps 0-3  producers: global -> registers -> split -> K-major "interleaved" (no-swizzle) smem
//              core-matrix layout, 3-stage mbarrier ring; afterwards the epilogue
//              (tcgen05.ld TMEM -> registers -> alpha/beta/bias/relu -> global)
//   warp  4    allocates TMEM, one elected lane issues tcgen05.mma.kind::tf32 (M=128, N=128, K=8)
//              and tcgen05.commit's to free smem stages / publish the accumulator.
// Because the producers rewrite the operands anyway, all three layouts end up in the same
// K-major smem layout: a K-contiguous global operand is read as float4 along k, a row-contiguous
// one (B of NN, A and B of TN) as coalesced scalars along the row and transposed on the fly.
// TN (weight gradients: the reduction runs over graph nodes) is split along K over gridDim.z and
// the partial tiles are folded in slice order by k_splitk_reduce (deterministic).
#include "kgb_common.cuh"
#include <stdlib.h>

namespace kgb {

namespace tc {

constexpr int TM = 128, TN_ = 128, BK = 16;          // CTA tile; k-block
constexpr int STAGES = 3;
constexpr int LBO = TM * 16 + 32;                    // byte stride between 16-byte k-chunks (+32: bank spread)
constexpr int SBO = 128;                             // byte stride between 8-row core matrices
constexpr int TILE_BYTES = (BK / 4) * LBO;           // one operand tile (hi or lo)
constexpr int STAGE_BYTES = 4 * TILE_BYTES;          // A_hi, A_lo, B_hi, B_lo
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256;
constexpr int THREADS = 224;                         // warps 0-2: A producers, 3-5: B producers, 6: MMA; 0-3 epilogue
constexpr int EPI_LD = TN_ + 4;                       // padded row stride (floats) of the epilogue transpose
constexpr int TMEM_COLS = 256;                       // [0,128): sum of hi.hi  [128,256): sum of the two cross terms

static_assert(4 * 32 * (TN_ + 4) * 4 <= STAGES * STAGE_BYTES, "epilogue transpose must fit in the stage ring");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(LBO >> 4) << 16) | ((uint64_t)(SBO >> 4) << 32) |
         ((uint64_t)1 << 46);  // version 1, SWIZZLE_NONE, K-major
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t (&r)[32], uint32_t taddr) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

__device__ __forceinline__ void split_store(char* hi_tile, char* lo_tile, int off, float4 v) {
  float4 h, l;
  h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
  h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
  h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
  h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
  l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
  *reinterpret_cast<float4*>(hi_tile + off) = h;
  *reinterpret_cast<float4*>(lo_tile + off) = l;
}

// One producer thread's share of a 128 x 16 operand block.
//   KC = true : the global operand is contiguous along k (p[row*ld + k]): 4 float4 per thread
//   KC = false: it is contiguous along the tile row (p[k*ld + col]): 16 coalesced scalars per thread
//               (thread t owns tile row t and transposes while storing)
template <bool KC>
struct Frag;
template <>
struct Frag<true> {
  float4 v[4];
  __device__ __forceinline__ void load(const float* __restrict__ p, int64_t ld, int64_t row0, int64_t n_rows, int64_t k0,
                                       int64_t kend, int t) {
    const int j = t & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t r = row0 + (t >> 2) + 32 * i, k = k0 + 4 * j;
      v[i] = (r < n_rows && k < kend) ? __ldg(reinterpret_cast<const float4*>(p + r * ld + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __device__ __forceinline__ void store(char* hi, char* lo, int t) const {
    const int j = t & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) split_store(hi, lo, j * LBO + ((t >> 2) + 32 * i) * 16, v[i]);
  }
};
template <>
struct Frag<false> {
  float v[16];
  __device__ __forceinline__ void load(const float* __restrict__ p, int64_t ld, int64_t col0, int64_t n_cols, int64_t k0,
                                       int64_t kend, int t) {
    const int64_t c = col0 + t;
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = (c < n_cols && k0 + k < kend) ? __ldg(p + (k0 + k) * ld + c) : 0.f;
  }
  __device__ __forceinline__ void store(char* hi, char* lo, int t) const {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      split_store(hi, lo, j * LBO + t * 16, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
  }
};

// One WARP's share = a whole 128 x 16 operand block (16 float4 per lane).  Producer warps each own complete
// k-blocks: a warp issues its loads, splits, stores, fences and signals without having any OTHER global load in
// flight -- `fence.proxy.async` lowers to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, and that MEMBAR drains every
// outstanding load of the thread, which silently serialises a register-prefetch pipeline (measured: 5.6 us of
// apparent load latency).  Latency is hidden ACROSS producer warps instead.
template <bool KC>
struct WFrag;
template <>
struct WFrag<true> {
  float4 v[16];
  __device__ __forceinline__ void load(const float* __restrict__ p, int64_t ld, int64_t row0, int64_t n_rows, int64_t k0,
                                       int64_t kend, int lane) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int q = i * 32 + lane;
      const int64_t r = row0 + (q >> 2), k = k0 + 4 * (q & 3);
      v[i] = (r < n_rows && k < kend) ? __ldg(reinterpret_cast<const float4*>(p + r * ld + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __device__ __forceinline__ void store(char* hi, char* lo, int lane) const {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int q = i * 32 + lane;
      split_store(hi, lo, (q & 3) * LBO + (q >> 2) * 16, v[i]);
    }
  }
};
template <>
struct WFrag<false> {
  float v[4][16];   // [column group][k]
  __device__ __forceinline__ void load(const float* __restrict__ p, int64_t ld, int64_t col0, int64_t n_cols, int64_t k0,
                                       int64_t kend, int lane) {
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int64_t c = col0 + lane + 32 * cc;
#pragma unroll
      for (int k = 0; k < 16; ++k) v[cc][k] = (c < n_cols && k0 + k < kend) ? __ldg(p + (k0 + k) * ld + c) : 0.f;
    }
  }
  __device__ __forceinline__ void store(char* hi, char* lo, int lane) const {
#pragma unroll
    for (int cc = 0; cc < 4; ++cc)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        split_store(hi, lo, j * LBO + (lane + 32 * cc) * 16,
                    make_float4(v[cc][4 * j], v[cc][4 * j + 1], v[cc][4 * j + 2], v[cc][4 * j + 3]));
  }
};

template <int LAYOUT>
__global__ void __launch_bounds__(THREADS, 2)
k_gemm_tc(const float* __restrict__ a, int64_t lda, const float* __restrict__ b, int64_t ldb, float* __restrict__ c,
          int64_t ldc, int64_t M, int64_t N, int64_t K, float alpha, float beta, const float* __restrict__ bias, int relu,
          float* __restrict__ part, int64_t k_per_split) {
  extern __shared__ __align__(128) char smem[];
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 1];
  __shared__ uint32_t tmem_base_slot;

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int64_t m0 = (int64_t)blockIdx.y * TM, n0 = (int64_t)blockIdx.x * TN_;
  const bool split = part != nullptr;
  const int64_t kbeg = split ? (int64_t)blockIdx.z * k_per_split : 0;
  const int64_t kend = split ? min(K, kbeg + k_per_split) : K;
  const int n_kb = (int)((kend - kbeg + BK - 1) / BK);

  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  const uint32_t accum_bar = bar0 + 8u * (2 * STAGES);

  if (t == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 2);   // the A-warp and the B-warp of this stage
      mbar_init(empty_bar(s), 1);  // tcgen05.commit
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 6) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;

  if (warp < 6) {
    // ================= producers =================
    // warp w < 3 fills the A half of stage w for k-blocks w, w+3, ...; warp 3+w the B half.  One warp = one whole
    // operand block, so nothing else of that warp is in flight when it fences (see WFrag).
    const int ws = warp % STAGES;
    const bool is_a = warp < STAGES;
    for (int kb = ws; kb < n_kb; kb += STAGES) {
      const int64_t k0 = kbeg + (int64_t)kb * BK;
      char* st = smem + (size_t)ws * STAGE_BYTES;
      if (is_a) {                 // loads are issued before the slot wait: they overlap it
        WFrag<LAYOUT != KGB_TN> f;
        f.load(a, lda, m0, M, k0, kend, lane);
        mbar_wait(empty_bar(ws), ((kb / STAGES) & 1) ^ 1);
        f.store(st, st + TILE_BYTES, lane);
      } else {
        WFrag<LAYOUT == KGB_NT> f;
        f.load(b, ldb, n0, N, k0, kend, lane);
        mbar_wait(empty_bar(ws), ((kb / STAGES) & 1) ^ 1);
        f.store(st + 2 * TILE_BYTES, st + 3 * TILE_BYTES, lane);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(ws));
    }
  }
  if (warp < 4) {
    // ================= epilogue =================
    // TMEM lane == tile row: each thread first owns one row.  Rows go through a padded smem transpose so that
    // every global access (C read for beta, C / partial write) is one coalesced 512-byte row segment per warp.
    mbar_wait(accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* ep = reinterpret_cast<float*>(smem) + (size_t)warp * 32 * EPI_LD;   // all MMAs are done: stages are free
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < TN_; c0 += 32) {
      uint32_t r[32], r2[32];
      tmem_ld32(r, taddr + (uint32_t)c0);
      tmem_ld32(r2, taddr + (uint32_t)(TN_ + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 v;
        v.x = __uint_as_float(r[4 * q]) + __uint_as_float(r2[4 * q]);
        v.y = __uint_as_float(r[4 * q + 1]) + __uint_as_float(r2[4 * q + 1]);
        v.z = __uint_as_float(r[4 * q + 2]) + __uint_as_float(r2[4 * q + 2]);
        v.w = __uint_as_float(r[4 * q + 3]) + __uint_as_float(r2[4 * q + 3]);
        *reinterpret_cast<float4*>(ep + lane * EPI_LD + c0 + 4 * q) = v;
      }
    }
    __syncwarp();
    const int64_t gn = n0 + 4 * lane;
    if (gn < N) {                                           // N % 4 == 0
      float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bias && !split) bb = __ldg(reinterpret_cast<const float4*>(bias + gn));
#pragma unroll 4
      for (int rr = 0; rr < 32; ++rr) {
        const int64_t gm = m0 + warp * 32 + rr;
        if (gm >= M) break;
        float4 v = *reinterpret_cast<const float4*>(ep + rr * EPI_LD + 4 * lane);
        if (split) {
          *reinterpret_cast<float4*>(part + ((int64_t)blockIdx.z * M + gm) * N + gn) = v;
        } else {
          v.x *= alpha; v.y *= alpha; v.z *= alpha; v.w *= alpha;
          float* cp = c + gm * ldc + gn;
          if (beta != 0.f) {
            const float4 o = *reinterpret_cast<const float4*>(cp);
            v.x = fmaf(beta, o.x, v.x); v.y = fmaf(beta, o.y, v.y); v.z = fmaf(beta, o.z, v.z); v.w = fmaf(beta, o.w, v.w);
          }
          v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
          if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          *reinterpret_cast<float4*>(cp) = v;
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else if (warp == 6) {
    // ================= MMA issuer (warp 6) =================
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN_ >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    for (int kb = 0; kb < n_kb; ++kb) {
      const int s = kb % STAGES;
      mbar_wait(full_bar(s), (kb / STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t st = smem_u32(smem + (size_t)s * STAGE_BYTES);
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
          const uint32_t koff = ks * 2 * LBO;
          const uint64_t a_hi = make_desc(st + koff), a_lo = make_desc(st + TILE_BYTES + koff);
          const uint64_t b_hi = make_desc(st + 2 * TILE_BYTES + koff), b_lo = make_desc(st + 3 * TILE_BYTES + koff);
          // The tensor core's fp32 accumulator truncates on every accumulate; keeping the (2^-11 smaller)
          // cross terms in their own accumulator makes their truncation error negligible and leaves the
          // main accumulator with K/8 instead of 3K/8 truncations.
          mma_tf32(tmem_base + TN_, a_lo, b_hi, idesc, (kb | ks) != 0);
          mma_tf32(tmem_base + TN_, a_hi, b_lo, idesc, 1);
          mma_tf32(tmem_base, a_hi, b_hi, idesc, (kb | ks) != 0);
        }
        mma_commit(empty_bar(s));                    // smem slot free once these MMAs have read it
        if (kb == n_kb - 1) mma_commit(accum_bar);   // accumulator complete
      }
      __syncwarp();
    }
    if (n_kb == 0 && lane == 0) mbar_arrive(accum_bar);  // empty K range: nothing to wait for (tile is garbage-free: see host)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 6) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}


// ------------------------------------------------------------------------------------------------
// Row-streaming variant for the node-row GEMMs  C[M, N<=128] = act(alpha * A[M, K<=128] op(B) + beta*C + bias)
// with M = number of graph nodes (784 k SNP rows): memory-bound, so the kernel is organised around keeping
// HBM busy.  One persistent CTA per SM:
//   * B (a [h,h] weight matrix) is split once and stays resident in smem for every tile of the CTA;
//   * warps 0-3 stream A k-blocks (one whole 128x16 block per warp and stage, see WFrag) across tile borders;
//   * warp 4 issues the MMAs into one of TWO TMEM accumulator sets;
//   * warps 5-8 drain the other set (tcgen05.ld -> smem transpose -> coalesced rows) while the next tile's MMAs
//     run, so loads, tensor work and stores of neighbouring tiles overlap inside one CTA.
namespace rs {
constexpr int STAGES_A = 4;
constexpr int KMAX = 128;
constexpr int B_TILE = (KMAX / 4) * LBO;                  // one resident B tile (hi or lo), full K
constexpr int A_STAGE = 2 * TILE_BYTES;                   // A_hi, A_lo of one k-block
constexpr int EPI_COLS = 32;                              // epilogue transposes 32 columns at a time
constexpr int EPI_LD2 = EPI_COLS + 4;
constexpr int EPI_WARP = 32 * EPI_LD2 * 4;
constexpr int SMEM = 2 * B_TILE + STAGES_A * A_STAGE + 4 * EPI_WARP + 256;
// PROD_WARPS (template parameter PW) x 8 KB of A in flight per SM; threads = producers, MMA issuer, 4 epilogue warps
constexpr int threads_rs(int pw) { return (pw + 1 + 4) * 32; }
static_assert(SMEM <= 227 * 1024, "row-streaming GEMM smem budget");

template <int LAYOUT, int PROD_WARPS>   // KGB_NT: B[N,K] k-contiguous; KGB_NN: B[K,N] row-contiguous; PROD_WARPS must be 2 * STAGES_A
__global__ void __launch_bounds__(threads_rs(PROD_WARPS), 1)
k_gemm_tc_rows(const float* __restrict__ a, int64_t lda, const float* __restrict__ b, int64_t ldb, float* __restrict__ c,
               int64_t ldc, int64_t M, int64_t N, int64_t K, float alpha, float beta, const float* __restrict__ bias,
               int relu) {
  extern __shared__ __align__(128) char smem[];
  __shared__ __align__(8) uint64_t bars[2 * STAGES_A + 5];
  __shared__ uint32_t tmem_base_slot;
  char* b_hi = smem;
  char* b_lo = smem + B_TILE;
  char* a_ring = smem + 2 * B_TILE;
  float* epi = reinterpret_cast<float*>(smem + 2 * B_TILE + STAGES_A * A_STAGE);

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  // blockIdx.y selects a 128-column slice of the output (N > 128: [nodes, h] x [h, R*h] products of the gene-sized
  // jobs): the slice's part of B stays resident, the CTAs of one slice share the row tiles between them
  {
    const int64_t n_off = (int64_t)blockIdx.y * TN_;
    b += LAYOUT == KGB_NT ? n_off * ldb : n_off;
    c += n_off;
    if (bias) bias += n_off;
    N = min((int64_t)TN_, N - n_off);
  }
  const int n_kb = (int)((K + BK - 1) / BK);
  const int64_t n_tiles = (M + TM - 1) / TM;
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (STAGES_A + s); };
  auto acc_full = [&](int s) { return bar0 + 8u * (2 * STAGES_A + s); };
  auto acc_empty = [&](int s) { return bar0 + 8u * (2 * STAGES_A + 2 + s); };
  const uint32_t b_ready = bar0 + 8u * (2 * STAGES_A + 4);

  if (t == 0) {
    for (int s = 0; s < STAGES_A; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), 4); }
    mbar_init(b_ready, PROD_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == PROD_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;

  if (warp < PROD_WARPS) {
    // ---------------- producers ----------------
    // resident B: all k-blocks, split once (warp w takes k-blocks w, w+4, ...)
    for (int kb = warp; kb < n_kb; kb += PROD_WARPS) {
      WFrag<LAYOUT == KGB_NT> fb;
      fb.load(b, ldb, 0, N, (int64_t)kb * BK, K, lane);
      fb.store(b_hi + kb * TILE_BYTES, b_lo + kb * TILE_BYTES, lane);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(b_ready);
    // A stream: a flat sequence of (tile, k-block) items; warp w takes items w, w+8, ... and item `it` goes
    // through stage it % 4.  A warp issues its loads first and only then waits for its slot, so up to 8 blocks
    // (64 KB) are in flight per SM while at most 4 are staged.
    const int64_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const int64_t n_items = my_tiles * n_kb;
    for (int64_t it = warp; it < n_items; it += PROD_WARPS) {
      const int64_t tile = blockIdx.x + (it / n_kb) * gridDim.x;
      const int s = (int)(it % STAGES_A);
      char* st = a_ring + (size_t)s * A_STAGE;
      const int64_t pos = it / STAGES_A;          // this item is the pos-th occupant of stage s
      WFrag<true> f;
      f.load(a, lda, tile * TM, M, (it % n_kb) * (int64_t)BK, K, lane);
      // Two warps alternate on a stage, so a warp's consecutive items are TWO barrier phases apart and a bare parity
      // wait could alias.  Waiting first for the previous occupant's publication pins this warp to at most one
      // phase ahead of a_empty[s] (that occupant could only publish after ITS slot wait had passed).
      if (pos > 0) mbar_wait(a_full(s), (uint32_t)((pos - 1) & 1));
      mbar_wait(a_empty(s), (uint32_t)((pos & 1) ^ 1));
      f.store(st, st + TILE_BYTES, lane);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full(s));
    }
  } else if (warp == PROD_WARPS) {
    // ---------------- MMA issuer ----------------
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN_ >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    mbar_wait(b_ready, 0);
    const uint32_t bh = smem_u32(b_hi), bl = smem_u32(b_lo);
    int64_t it = 0;
    int iter = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++iter) {
      const int set = iter & 1;
      mbar_wait(acc_empty(set), (uint32_t)(((iter >> 1) & 1) ^ 1));      // the epilogue has drained this set
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d_main = tmem_base + set * 256, d_cross = d_main + TN_;
      for (int kb = 0; kb < n_kb; ++kb, ++it) {
        const int s = (int)(it % STAGES_A);
        mbar_wait(a_full(s), (uint32_t)((it / STAGES_A) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint32_t st = smem_u32(a_ring + (size_t)s * A_STAGE);
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            const uint32_t ka = ks * 2 * LBO, kbo = kb * TILE_BYTES + ks * 2 * LBO;
            const uint64_t a_hi_d = make_desc(st + ka), a_lo_d = make_desc(st + TILE_BYTES + ka);
            const uint64_t b_hi_d = make_desc(bh + kbo), b_lo_d = make_desc(bl + kbo);
            mma_tf32(d_cross, a_lo_d, b_hi_d, idesc, (kb | ks) != 0);
            mma_tf32(d_cross, a_hi_d, b_lo_d, idesc, 1);
            mma_tf32(d_main, a_hi_d, b_hi_d, idesc, (kb | ks) != 0);
          }
          mma_commit(a_empty(s));
          if (kb == n_kb - 1) mma_commit(acc_full(set));
        }
        __syncwarp();
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else {
    // ---------------- epilogue (4 warps after the MMA warp; TMEM lane quarter = warp % 4) ----------------
    const int q = warp & 3;
    float* ep = epi + (size_t)(warp - PROD_WARPS - 1) * 32 * EPI_LD2;
    int iter = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++iter) {
      const int set = iter & 1;
      mbar_wait(acc_full(set), (uint32_t)((iter >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + set * 256 + ((uint32_t)(q * 32) << 16);
      const int64_t m_base = tile * TM + q * 32;
#pragma unroll 1
      for (int half = 0; half < TN_ / EPI_COLS; ++half) {
#pragma unroll 1
        for (int c0 = 0; c0 < EPI_COLS; c0 += 32) {
          uint32_t r[32], r2[32];
          tmem_ld32(r, taddr + (uint32_t)(half * EPI_COLS + c0));
          tmem_ld32(r2, taddr + (uint32_t)(TN_ + half * EPI_COLS + c0));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            float4 v;
            v.x = __uint_as_float(r[4 * k4]) + __uint_as_float(r2[4 * k4]);
            v.y = __uint_as_float(r[4 * k4 + 1]) + __uint_as_float(r2[4 * k4 + 1]);
            v.z = __uint_as_float(r[4 * k4 + 2]) + __uint_as_float(r2[4 * k4 + 2]);
            v.w = __uint_as_float(r[4 * k4 + 3]) + __uint_as_float(r2[4 * k4 + 3]);
            *reinterpret_cast<float4*>(ep + lane * EPI_LD2 + c0 + 4 * k4) = v;
          }
        }
        if (half == TN_ / EPI_COLS - 1) {   // every TMEM read of this tile is complete: hand the set back to the MMA warp
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty(set));
        }
        __syncwarp();
        // 8 lanes x float4 = one 128-byte row segment; a warp writes four rows per instruction
        const int64_t gn = half * EPI_COLS + 4 * (lane & 7);
        if (gn < N) {
          float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias) bb = __ldg(reinterpret_cast<const float4*>(bias + gn));
#pragma unroll 4
          for (int rr = lane >> 3; rr < 32; rr += 4) {
            const int64_t gm = m_base + rr;
            if (gm >= M) break;
            float4 v = *reinterpret_cast<const float4*>(ep + rr * EPI_LD2 + 4 * (lane & 7));
            v.x *= alpha; v.y *= alpha; v.z *= alpha; v.w *= alpha;
            float* cp = c + gm * ldc + gn;
            if (beta != 0.f) {
              const float4 o = *reinterpret_cast<const float4*>(cp);
              v.x = fmaf(beta, o.x, v.x); v.y = fmaf(beta, o.y, v.y); v.z = fmaf(beta, o.z, v.z); v.w = fmaf(beta, o.w, v.w);
            }
            v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
            if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            *reinterpret_cast<float4*>(cp) = v;
          }
        }
        __syncwarp();
      }
    }
  }
  __syncthreads();
  if (warp == PROD_WARPS) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}
}  // namespace rs

}  // namespace tc

__global__ void k_splitk_reduce(const float* __restrict__ part, int splits, int64_t M, int64_t N, float* __restrict__ c,
                                int64_t ldc, float alpha, float beta, const float* __restrict__ bias, int relu);

static int tc_splits(int64_t m, int64_t n, int64_t k) {
  const int64_t tiles = ((m + tc::TM - 1) / tc::TM) * ((n + tc::TN_ - 1) / tc::TN_);
  int64_t s = (2 * kNumSMs + tiles - 1) / tiles;       // two resident CTAs per SM
  const int64_t max_s = (k + 16 * tc::BK - 1) / (16 * tc::BK);  // at least 256 k per slice
  if (s > max_s) s = max_s;
  return (int)(s < 1 ? 1 : s);
}

bool gemm_tc_supported(int layout, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc) {
  (void)lda; (void)ldb; (void)ldc;
  if (k < 1) return false;
  // tiny problems: the FFMA kernel has less fixed cost than TMEM allocation + pipeline fill
  const double flops = 2.0 * (double)m * (double)n * (double)k;
  static const char* tenv = getenv("KGB_GEMM_TC_MIN_FLOPS");   // A/B knob for measurements
  static const double min_flops = tenv ? atof(tenv) : 3.0e7;
  if (flops < min_flops) return false;
  if (layout == KGB_TN) return true;
  return true;
}

size_t gemm_tc_workspace_bytes(int layout, int64_t m, int64_t n, int64_t k) {
  if (layout != KGB_TN) return 0;
  const int s = tc_splits(m, n, k);
  return s > 1 ? (size_t)s * m * n * sizeof(float) + 256 : 0;
}

int gemm_tc(int layout, const float* a, int64_t lda, const float* b, int64_t ldb, float* c, int64_t ldc, int64_t M,
            int64_t N, int64_t K, float alpha, float beta, const float* bias, int relu, void* ws, size_t ws_bytes,
            cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    KGB_CUDA_OK(cudaFuncSetAttribute(tc::k_gemm_tc<KGB_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    KGB_CUDA_OK(cudaFuncSetAttribute(tc::k_gemm_tc<KGB_NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    KGB_CUDA_OK(cudaFuncSetAttribute(tc::k_gemm_tc<KGB_TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    attr_set = true;
  }
  // node-row GEMMs with a resident [<=128, <=128] weight operand: persistent row-streaming kernel
  // (also the gene-sized [nodes, h] x [h, R*h] products, one 128-column slice of the output per blockIdx.y: a tile
  //  costs ~4.6 us here against ~20 us per CTA in the one-tile-per-CTA kernel below)
  static const char* rs_env = getenv("KGB_GEMM_ROWS_MIN_M");
  static const int64_t rs_min_m = rs_env ? atoll(rs_env) : 64 * 1024;   // measured: 4096 is neutral for the step (6.26 vs 6.23 ms)
  if (layout != KGB_TN && (N <= tc::TN_ || N % tc::TN_ == 0) && N <= 8 * tc::TN_ && K <= tc::rs::KMAX && K % tc::BK == 0 &&
      (M >= 64 * 1024 || (M >= rs_min_m && N >= tc::TN_))) {
    // (8 producer warps: two per stage of the 4-stage ring, which is what the phase bookkeeping of the producers assumes;
    //  12 warps -- three per stage -- deadlock, measured in round 2)
    static bool rs_attr = false;
    if (!rs_attr) {
      KGB_CUDA_OK(cudaFuncSetAttribute(tc::rs::k_gemm_tc_rows<KGB_NT, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::rs::SMEM));
      KGB_CUDA_OK(cudaFuncSetAttribute(tc::rs::k_gemm_tc_rows<KGB_NN, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::rs::SMEM));
      rs_attr = true;
    }
    const int64_t tiles = (M + tc::TM - 1) / tc::TM;
    const int64_t slices = (N + tc::TN_ - 1) / tc::TN_;
    int64_t gx = kNumSMs / slices;                       // one wave of CTAs: one CTA per SM (218 KB of shared memory)
    if (gx < 1) gx = 1;
    if (gx > tiles) gx = tiles;
    const dim3 g((unsigned)gx, (unsigned)slices, 1);
#define KGB_ROWS(LAY, PW) tc::rs::k_gemm_tc_rows<LAY, PW><<<g, tc::rs::threads_rs(PW), tc::rs::SMEM, stream>>>(a, lda, b, ldb, c, ldc, M, N, K, alpha, beta, bias, relu)
    if (layout == KGB_NT) KGB_ROWS(KGB_NT, 8);
    else KGB_ROWS(KGB_NN, 8);
#undef KGB_ROWS
    KGB_LAUNCH_OK();
    return KGB_OK;
  }
  dim3 grid((unsigned)((N + tc::TN_ - 1) / tc::TN_), (unsigned)((M + tc::TM - 1) / tc::TM), 1);
  if (layout == KGB_NT) {
    tc::k_gemm_tc<KGB_NT><<<grid, tc::THREADS, tc::SMEM_BYTES, stream>>>(a, lda, b, ldb, c, ldc, M, N, K, alpha, beta, bias,
                                                                         relu, nullptr, 0);
  } else if (layout == KGB_NN) {
    tc::k_gemm_tc<KGB_NN><<<grid, tc::THREADS, tc::SMEM_BYTES, stream>>>(a, lda, b, ldb, c, ldc, M, N, K, alpha, beta, bias,
                                                                         relu, nullptr, 0);
  } else {
    const int s = tc_splits(M, N, K);
    if (s <= 1) {
      tc::k_gemm_tc<KGB_TN><<<grid, tc::THREADS, tc::SMEM_BYTES, stream>>>(a, lda, b, ldb, c, ldc, M, N, K, alpha, beta,
                                                                           bias, relu, nullptr, 0);
    } else {
      if (!ws || ws_bytes < gemm_tc_workspace_bytes(layout, M, N, K)) {
        set_error("gemm_tc: workspace %zu < %zu", ws_bytes, gemm_tc_workspace_bytes(layout, M, N, K));
        return KGB_ERR_WORKSPACE;
      }
      int64_t kps = (K + s - 1) / s;
      kps = (kps + tc::BK - 1) / tc::BK * tc::BK;
      grid.z = (unsigned)((K + kps - 1) / kps);
      float* part = static_cast<float*>(ws);
      tc::k_gemm_tc<KGB_TN><<<grid, tc::THREADS, tc::SMEM_BYTES, stream>>>(a, lda, b, ldb, c, ldc, M, N, K, alpha, beta,
                                                                           bias, relu, part, kps);
      KGB_LAUNCH_OK();
      const int64_t n4 = M * N / 4;
      k_splitk_reduce<<<(unsigned)((n4 + 63) / 64), 256, 0, stream>>>(part, (int)grid.z, M, N, c, ldc, alpha, beta, bias,
                                                                        relu);
    }
  }
  KGB_LAUNCH_OK();
  return KGB_OK;
}

}  // namespace kgb
