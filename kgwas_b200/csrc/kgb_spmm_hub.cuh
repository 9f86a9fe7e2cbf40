// Hub rows of a gather-reduce from shared-memory tiles (TMA bulk copies), instead of one L2 / DRAM gather per edge.
//
// Why (profiles/r01_ncu_spmm_step.md): in the (gene, relation)-row launches of the KGWAS layer -- aggregate-first
// forward SNP -> Gene, transform-first backward Gene -> SNP -- the 128 heaviest rows hold half of the 8 M edges and each
// of them sweeps a large part of the 401 MB SNP table: the pull kernel moves 4.06 GB through L2 and re-reads the table
// 3.6x from DRAM.  Here the table is streamed ONCE: every CTA owns a contiguous range of gathered-table rows, stages it
// tile by tile in shared memory with cp.async.bulk (1-D TMA) behind an mbarrier ring, and reduces every hub edge that
// falls into the tile from shared memory.
//
//   plan time (kgwas_b200/_lib.py: Csr.build_hub): the hub rows are split into `nv` virtual hub slots (a hub heavier
//   than half a warp's share is cut into parts), every slot is owned by exactly one of the 16 consumer warps (LPT
//   balance), and the hub edges are re-sorted by (tile, warp, slot) into per-tile chunks
//       [ 24 x int32 header: first record of warp 0..15, end | records: {flush << 31 | slot << 8 | row in tile, weight} ]
//   so that ONE bulk copy brings a tile's whole edge list next to its feature rows.
//   kernel: producer warp = one thread issuing two bulk copies per tile (feature rows, chunk) into a 2-stage ring;
//   16 consumer warps walk their own record range: broadcast LDS of the record, one LDS.128 per lane for the row,
//   packed FFMA2, and -- when the record carries the flush bit (last edge of its slot in this tile) -- a plain
//   read-modify-write of the slot's accumulator row in shared memory (a slot has one owner: no atomics, fixed order,
//   bit-reproducible).  At the end the CTA writes its `nv` accumulator rows to partial[cta]; k_hub_fold sums the
//   partials of every hub row over CTAs and parts in index order and applies the usual epilogue.
// The remaining (non-hub) rows run through lean::k_spmm_lean with the hub rows' segments left out of its item list.
#pragma once
#include "kgb_common.cuh"
#include "kgb_spmm_lean.cuh"

namespace kgb {
namespace hub {

constexpr int kWarps = 16;                    // consumer warps
constexpr int kThreads = (kWarps + 1) * 32;   // + the producer warp
constexpr int kStages = 2;
constexpr int kHdrInts = 24;                  // chunk header: 17 offsets, padded to 96 bytes

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
// 1-D TMA: global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

struct Params {
  const float* x;           // gathered table: n_cols contiguous rows of H floats
  const char* chunks;       // per-tile chunks (header + records), 16-byte aligned
  const int64_t* tile_off;  // [n_tiles + 1] byte offset of every chunk
  float* partial;           // [gridDim.x][nv][H]
  int n_cols, tile_rows, n_tiles, nv, chunk_cap;
};

template <int NV>
__global__ void __launch_bounds__(kThreads, 1) k_hub_tile(Params p) {
  constexpr int H = NV * 128;
  extern __shared__ __align__(128) unsigned char smem[];
  float* acc = reinterpret_cast<float*>(smem);
  const uint32_t acc_bytes = (uint32_t)p.nv * H * 4;
  const uint32_t tile_bytes = (uint32_t)p.tile_rows * H * 4;
  const uint32_t stage_bytes = tile_bytes + (uint32_t)p.chunk_cap;
  unsigned char* stage0 = smem + acc_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage0 + kStages * stage_bytes);   // full[kStages], empty[kStages]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = (int)((int64_t)p.n_tiles * blockIdx.x / gridDim.x);
  const int t1 = (int)((int64_t)p.n_tiles * (blockIdx.x + 1) / gridDim.x);

  for (int i = threadIdx.x; i < p.nv * (H / 4); i += kThreads)
    reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(bars + s), 1);
      mbar_init(smem_u32(bars + kStages + s), kWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == kWarps) {
    // ------------------------------------------------------------------ producer: one thread, two bulk copies per tile
    if (lane == 0) {
      for (int t = t0, it = 0; t < t1; ++t, ++it) {
        const int s = it % kStages;
        if (it >= kStages) mbar_wait(smem_u32(bars + kStages + s), (uint32_t)(((it / kStages) & 1) ^ 1));
        const int64_t o0 = __ldg(p.tile_off + t), o1 = __ldg(p.tile_off + t + 1);
        const uint32_t cbytes = (uint32_t)(o1 - o0);
        const int rows = min(p.tile_rows, p.n_cols - t * p.tile_rows);
        const uint32_t xbytes = (uint32_t)rows * H * 4;
        const uint32_t full = smem_u32(bars + s);
        unsigned char* st = stage0 + (size_t)s * stage_bytes;
        mbar_arrive_expect_tx(full, xbytes + cbytes);
        bulk_g2s(smem_u32(st), p.x + (int64_t)t * p.tile_rows * H, xbytes, full);
        bulk_g2s(smem_u32(st + tile_bytes), p.chunks + o0, cbytes, full);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ consumers
    for (int t = t0, it = 0; t < t1; ++t, ++it) {
      const int s = it % kStages;
      mbar_wait(smem_u32(bars + s), (uint32_t)((it / kStages) & 1));
      const unsigned char* st = stage0 + (size_t)s * stage_bytes;
      const float4* xl = reinterpret_cast<const float4*>(st) + lane;      // this lane's 16 bytes of tile row 0
      const int* hdr = reinterpret_cast<const int*>(st + tile_bytes);
      const int2* rec = reinterpret_cast<const int2*>(hdr + kHdrInts);
      int e = hdr[warp];
      const int e1 = hdr[warp + 1];
      lean::AccT<NV> a;
      a.zero();
      for (; e < e1; e += 4) {
        int2 r[4];
        float4 tr[4][NV];
#pragma unroll
        for (int u = 0; u < 4; ++u) r[u] = (e + u < e1) ? rec[e + u] : make_int2(0, 0);   // broadcast LDS.64
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4* row = xl + (r[u].x & 0xff) * (H / 4);
#pragma unroll
          for (int q = 0; q < NV; ++q) tr[u][q] = row[32 * q];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (e + u < e1) {
            a.fma(__int_as_float(r[u].y), tr[u]);
            if (r[u].x < 0) {                                  // last edge of its slot in this tile: fold into the slot
              float4* ap = reinterpret_cast<float4*>(acc + (size_t)((r[u].x >> 8) & 0x7fffff) * H) + lane;
#pragma unroll
              for (int q = 0; q < NV; ++q) {
                float4 o = ap[32 * q];
                o.x += a.v[2 * q].x; o.y += a.v[2 * q].y; o.z += a.v[2 * q + 1].x; o.w += a.v[2 * q + 1].y;
                ap[32 * q] = o;
              }
              a.zero();
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(bars + kStages + s));
    }
  }
  __syncthreads();
  float4* dst = reinterpret_cast<float4*>(p.partial + (size_t)blockIdx.x * p.nv * H);
  for (int i = threadIdx.x; i < p.nv * (H / 4); i += kThreads) dst[i] = reinterpret_cast<const float4*>(acc)[i];
}

// y[hub row] = epilogue( sum over CTAs (outer) and parts (inner) of partial[cta][slot] ), fixed order
constexpr int kFoldWarps = 8;
template <int NV>
__global__ void __launch_bounds__(kFoldWarps * 32) k_hub_fold(const float* __restrict__ partial, int n_cta, int nv,
                                                              const int32_t* __restrict__ hub_row,
                                                              const int32_t* __restrict__ hub_vptr, float* __restrict__ y,
                                                              int64_t ldy, lean::Epi ep) {
  constexpr int H = NV * 128;
  __shared__ float red[kFoldWarps][H];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hb = blockIdx.x;
  const int v0 = __ldg(hub_vptr + hb), v1 = __ldg(hub_vptr + hb + 1);
  lean::AccT<NV> sum;
  sum.zero();
  const int per = (n_cta + kFoldWarps - 1) / kFoldWarps;
  const int c0 = warp * per, c1 = min(n_cta, c0 + per);
  for (int c = c0; c < c1; ++c) {
    for (int v = v0; v < v1; ++v) {
      const float4* p = reinterpret_cast<const float4*>(partial + ((size_t)c * nv + v) * H) + lane;
      float4 t[NV];
#pragma unroll
      for (int q = 0; q < NV; ++q) t[q] = __ldcg(p + 32 * q);
      sum.fma(1.f, t);
    }
  }
  lean::store_acc<NV>(sum, red[warp], lane);
  __syncthreads();
  if (warp == 0) {
    lean::AccT<NV> tot;
    tot.zero();
    for (int w = 0; w < kFoldWarps; ++w) {
      float4 t[NV];
#pragma unroll
      for (int q = 0; q < NV; ++q) t[q] = *(reinterpret_cast<const float4*>(red[w]) + lane + 32 * q);
      tot.fma(1.f, t);
    }
    const int row = __ldg(hub_row + hb);
    float* yrow = y + (int64_t)row * ldy;
    float4 old[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) old[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ep.beta != 0.f) lean::load_old<NV>(old, yrow, lane);
    lean::finish_row<NV>(tot, old, yrow, row, ep, lane);
  }
}

inline size_t smem_bytes(int nv, int tile_rows, int chunk_cap, int h) {
  return (size_t)nv * h * 4 + (size_t)kStages * ((size_t)tile_rows * h * 4 + chunk_cap) + 2 * kStages * 8 + 16;
}

}  // namespace hub
}  // namespace kgb
