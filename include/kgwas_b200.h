/*
 * kgwas_b200.h -- C ABI of libkgwas_b200.so, the sm_100a message-passing engine behind
 * kgwas_b200's drop-in HeteroGNN / HeteroConv / SAGEConv / GATConv.
 *
 * The reference (snap-stanford/KGWAS) has no FFI of its own: its hot path is Python calling
 * torch_geometric.  Each entry point below cites the reference interface whose arithmetic it
 * replaces (paths relative to /root/reference).  Conventions:
 *   - every pointer is a DEVICE pointer owned by the caller (torch), unless named h_*;
 *   - every call takes the CUDA stream to launch on and never synchronises the device,
 *     never allocates: scratch is a caller-supplied workspace sized by *_workspace_bytes();
 *   - return value 0 = ok, negative = KGB_ERR_*; kgb_last_error() gives the thread-local text;
 *   - node features are row-major fp32 with an explicit row stride (in floats); feature widths
 *     must be a multiple of 32 and <= 512; indices are int32 on the device side
 *     (int64 COO at the graph boundary, as PyG's edge_index is).
 */
#ifndef KGWAS_B200_H_
#define KGWAS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define KGB_API __attribute__((visibility("default")))
#else
#define KGB_API
#endif

typedef void* kgb_stream_t; /* cudaStream_t */

enum {
  KGB_OK = 0,
  KGB_ERR_INVALID = -1,   /* bad argument (shape, alignment, null pointer) */
  KGB_ERR_WORKSPACE = -2, /* workspace too small */
  KGB_ERR_CUDA = -3,      /* a CUDA runtime call failed; see kgb_last_error() */
  KGB_ERR_UNSUPPORTED = -4
};

/* ---- library identity ------------------------------------------------------------- */
KGB_API int kgb_version(void);            /* MAJOR*10000 + MINOR*100 + PATCH */
KGB_API int kgb_sm_arch(void);            /* 100 : compiled for sm_100a only */
KGB_API const char* kgb_last_error(void); /* thread-local, never NULL */
KGB_API long long kgb_launch_count(void); /* kernels launched by this library so far (process-wide) */

/* ---- graph bookkeeping ------------------------------------------------------------ */
/* COO (PyG edge_index [2,E] int64, unsorted, duplicates allowed; kgwas/model.py:53
 * `edge_index_dict`) -> CSR by destination + CSR by source ("transposed"), both with STABLE
 * edge order (ties keep original edge order) so results are bit-reproducible:
 *   rowptr  [n_dst+1]  col  [E] = src of the e-th edge in dst-major order
 *   eperm   [E]        original edge id of CSR slot i      (carries GAT alpha back to COO order)
 *   t_rowptr[n_src+1]  t_col[E] = dst of the j-th edge in src-major order
 *   t_eperm [E]        CSR slot of transposed slot j       (weights_csc = weights_csr[t_eperm])
 * sort_cols != 0: slots of one row are ordered by (source index, original edge id) instead of by
 * original edge id alone -- heavy rows then sweep the gathered table front to back, which is what
 * the L2-window scheduling of kgb_spmm relies on.  presort_key (nullable, values in [0, n_src)) replaces
 * the source index as that in-row ordering key.  The transposed CSR is column-sorted either way.
 * Replaces the index handling inside PyG MessagePassing.propagate / scatter (reached from
 * kgwas/conv.py:182 and kgwas/model.py:74).  Any of the t_* outputs may be NULL (all three). */
KGB_API size_t kgb_csr_build_workspace_bytes(int64_t n_edges, int64_t n_src, int64_t n_dst);
KGB_API int kgb_csr_build(const int64_t* src, const int64_t* dst, int64_t n_edges, int64_t n_src,
                  int64_t n_dst, int32_t sort_cols, const int64_t* presort_key, int32_t* rowptr,
                  int32_t* col, int32_t* eperm,
                  int32_t* t_rowptr, int32_t* t_col, int32_t* t_eperm, void* workspace,
                  size_t workspace_bytes, kgb_stream_t stream);

/* Rows longer than seg_len are split into fixed-length segments that are reduced by separate
 * warps and combined in segment order (deterministic).  Two-phase: count, then fill.
 *   hrow_id    [n_hrows]    row index of each heavy row (ascending)
 *   hrow_segptr[n_hrows+1]  prefix sum of segments per heavy row
 *   hseg_hrow  [n_hsegs]    heavy-row slot of each segment
 * (hrow_grpptr of kgb_csr_t is derived from hrow_segptr by the caller)
 * h_counts (HOST, 2 ints) receives {n_hrows, n_hsegs}; this one call synchronises the stream. */
KGB_API int kgb_csr_heavy_count(const int32_t* rowptr, int32_t n_rows, int32_t seg_len, int32_t* h_counts,
                        void* workspace, size_t workspace_bytes, kgb_stream_t stream);
KGB_API size_t kgb_csr_heavy_workspace_bytes(int32_t n_rows);
KGB_API int kgb_csr_heavy_fill(const int32_t* rowptr, int32_t n_rows, int32_t seg_len, int32_t n_hrows,
                       int32_t n_hsegs, int32_t* hrow_id, int32_t* hrow_segptr, int32_t* hseg_hrow,
                       void* workspace, size_t workspace_bytes, kgb_stream_t stream);

/* ---- segmented gather-reduce (the message-passing core) --------------------------- */
/* y[i,:] = act( beta*y[i,:] + bias + sum_{j in [rowptr[i], rowptr[i+1])} w_j * x[col[j], :] )
 *   w_j = ew ? ew[wperm ? wperm[j] : j] : 1
 *   rowsum2[i*bins + col[j] % bins] += ew2[wperm ? wperm[j] : j]      (bins >= 1; optional)
 * One pass: gather -> weighted segmented reduce -> write; no [E,h] message tensor, no atomics
 * on the data path.  Replaces index_select + scatter_add(/mean) of SAGEConv (PyG; instantiated
 * kgwas/model.py:38; mean = precomputed 1/deg edge weights) and `alpha.unsqueeze(-1) * x_j` +
 * 'add' aggregation of GATConv (kgwas/conv.py:227-228, :54); run on the transposed CSR (with
 * wperm = t_eperm so the weights stay in CSR order) it is their backward (index_add / gather).
 * dot_w [h] / dot_out [n_rows] (both or neither): dot_out[i] = <y[i,:], dot_w> of the row just
 * written -- the single-output head `self.lin` applied to the ReLU-ed SNP rows (kgwas/model.py:50,
 * 83-86) without a second pass over the [N_SNP, h] tensor.
 * Dispatch: ew2 == NULL, wperm == NULL and h % 128 == 0 (every launch of the SAGE path) runs the
 * instruction-lean, software-pipelined kernel; anything else the generic one.  Results do not
 * depend on the kernel (same summation order).  Environment knobs for A/B measurements only:
 * KGB_SPMM_VARIANT=0 forces the generic kernel, KGB_SPMM_WAVES=n sets the CTA waves of
 * row-dominated launches (default 8).
 * heavy_* come from kgb_csr_heavy_*; pass n_hrows = n_hsegs = 0 when no row exceeds seg_len.
 * scratch: kgb_spmm_scratch_bytes(), zero-initialised ONCE by the caller (the kernel leaves its
 * ticket counters zero again on exit); may be NULL when n_hsegs == 0. */
typedef struct {
  const int32_t* rowptr;
  const int32_t* col;
  int32_t n_rows;
  int32_t seg_len;
  int32_t n_hrows;
  int32_t n_hsegs;
  const int32_t* hrow_id;
  const int32_t* hrow_segptr;
  const int32_t* hseg_hrow;
  const int32_t* hseg_order; /* nullable: work item i processes segment hseg_order[i] (L2-window scheduling) */
  const int32_t* hrow_grpptr; /* [n_hrows+1] prefix sum of ceil(segments / KGB_FOLD) per heavy row */
  int32_t n_hgroups;          /* hrow_grpptr[n_hrows] */
  int64_t n_edges_hint;       /* rowptr[n_rows] if known on the host, else 0: grid policy (row- vs segment-dominated) */
  const int32_t* hitem;       /* nullable [n_hitems][4]: work item i = (first slot, slot count, segment id, heavy-row
                                 slot) in hseg_order order -- everything a warp needs to start a segment in ONE
                                 16-byte load (the lean kernel needs it when n_hsegs > 0) */
  int32_t n_hitems;           /* entries of hitem (= n_hsegs when hitem lists every segment) */
  /* ---- hub plan (optional; hub_n == 0: none).  The hub_n heaviest rows are reduced from shared-memory tiles of the
   * gathered table that are staged once per CTA with 1-D TMA (cp.async.bulk), instead of one L2 / DRAM gather per
   * edge (kgb_spmm_hub.cuh).  Used by kgb_spmm when ew == hub_ew (the weights are baked into the chunks), the lean
   * kernel applies (h % 128 == 0, no ew2 / wperm) and ldx == h; then hitem_tail / n_hitems_tail (the work items of
   * the NON-hub heavy rows) replace hitem / n_hitems for the pull kernel.  Built by the caller at plan time:
   *   slots: hub row i owns virtual slots [hub_vptr[i], hub_vptr[i+1]) (a hub is cut into parts for load balance);
   *   chunk of tile t (gathered rows [t*T, (t+1)*T)) at hub_chunks + hub_tile_off[t], 16-byte aligned:
   *     int32 hdr[24]: hdr[w] = first record of consumer warp w (w = 0..15), hdr[16] = end; then records
   *     {int32 (flush << 31) | (slot << 8) | (row - t*T), float weight}, sorted by (warp, slot); flush = last record
   *     of its slot in this tile; every slot belongs to exactly one warp. */
  int32_t hub_n;              /* hub rows */
  int32_t hub_nv;             /* virtual slots */
  int32_t hub_tile_rows;      /* T <= 256 */
  int32_t hub_n_tiles;        /* ceil(n gathered rows / T) */
  int32_t hub_n_cta;          /* CTAs of the tile kernel = partial sets in the scratch */
  int32_t hub_chunk_cap;      /* >= largest chunk + 32 bytes, multiple of 16 */
  int64_t hub_n_cols;         /* rows of the gathered table */
  const int64_t* hub_tile_off; /* [hub_n_tiles + 1] */
  const char* hub_chunks;
  const int32_t* hub_row;     /* [hub_n] row index of each hub */
  const int32_t* hub_vptr;    /* [hub_n + 1] */
  const float* hub_ew;        /* the edge-weight array the chunk weights were taken from (identity = contract) */
  const int32_t* hitem_tail;
  int32_t n_hitems_tail;
  int32_t n_mid_rows;         /* rows with 16 < degree <= seg_len if the caller knows it, else -1 (kgb_gat_*: 0 lets the
                                 warp-per-group pass be skipped when the one-thread-per-group pass covers every row) */
  const int32_t* mid_row_id;  /* nullable [n_mid_rows]: those rows, ascending -- the warp-per-group pass then visits only
                                 them instead of scanning all n_rows row pointers */
} kgb_csr_t;
enum { KGB_FOLD = 64 };       /* partial sums are folded 64 at a time (two levels) by the last finisher */

KGB_API size_t kgb_spmm_scratch_bytes(int32_t n_hrows, int32_t n_hsegs, int32_t n_hgroups, int32_t h);
/* the same plus the hub plan's partial sets ([hub_n_cta][hub_nv][h] floats) when csr->hub_n > 0 */
KGB_API size_t kgb_spmm_scratch_bytes_csr(const kgb_csr_t* csr, int32_t h);
enum { KGB_MAX_BINS = 8 };
KGB_API int kgb_spmm(const kgb_csr_t* csr, const float* ew, const int32_t* wperm, const float* ew2,
                     float* rowsum2, int32_t rowsum2_bins, const float* x, int64_t ldx, float* y,
                     int64_t ldy, int32_t h, float beta, const float* bias, int32_t relu,
                     const float* dot_w, float* dot_out, void* scratch, size_t scratch_bytes,
                     kgb_stream_t stream);

/* ---- dense contractions ([nodes x in] . [in x out]) ------------------------------- */
/* C[M,N] = act( alpha * op(A).op(B) + beta*C + bias[N] )      (row-major, strides in floats)
 *   layout KGB_NT : A[M,K] (lda), B[N,K] (ldb)   C = A.B^T     forward  x.W^T   (PyG Linear,
 *                                                              kgwas/model.py:38,50; conv.py:81-89)
 *   layout KGB_NN : A[M,K] (lda), B[K,N] (ldb)   C = A.B       dX = G.W
 *   layout KGB_TN : A[K,M] (lda), B[K,N] (ldb)   C = A^T.B     dW = G^T.X  (K = node rows)
 * fp32 in / fp32 out; accumulation is fp32-equivalent (3xTF32 split on the tensor cores for the
 * tcgen05 path, FFMA for small or odd shapes).  Needs N % 4 == 0, and K % 4 == 0 (NT/NN) or
 * M % 4 == 0 (TN).  workspace: kgb_gemm_workspace_bytes() (split-K partials). */
enum { KGB_NT = 0, KGB_NN = 1, KGB_TN = 2 };
KGB_API size_t kgb_gemm_workspace_bytes(int32_t layout, int64_t m, int64_t n, int64_t k);
KGB_API int kgb_gemm(int32_t layout, const float* a, int64_t lda, const float* b, int64_t ldb,
                     float* c, int64_t ldc, int64_t m, int64_t n, int64_t k, float alpha,
                     float beta, const float* bias, int32_t relu, void* workspace,
                     size_t workspace_bytes, kgb_stream_t stream);

/* ---- full-neighbour mini-batch assembly ---------------------- */
/* GPU replacement for the host-side sampling of NeighborLoader(data, num_neighbors=[-1]*L, input_nodes=('SNP', ids),
 * batch_size) (kgwas/kgwas.py:99-113; third-party C++ in the reference): L-hop frontier expansion with bit-exact node /
 * edge bookkeeping (oracle/bookkeeping.py: full_neighbor_subgraph_ref).  The in-adjacency of a relation is the
 * destination-major CSR kgb_csr_build returns (rowptr, col = sources, eperm = original edge ids, stable edge order).
 *   kgb_frontier_count : offsets[i] = number of in-edges of frontier[0..i) ([n_f + 1]); *h_total (HOST) = their sum.
 *   kgb_frontier_expand: for every frontier node in order, all its in-edges in edge order: eids[s], srcs[s], s < total.
 *   kgb_frontier_add   : candidates cand[0..n) (sources of the expanded edges, or the seeds); those whose local id is
 *                        still -1 get ids count_base, count_base+1, ... in FIRST-OCCURRENCE order: local[v] is set,
 *                        new_nodes[0..n_new) lists them in that order, *h_n_new (HOST) = n_new.  firstpos is a scratch
 *                        table [n_nodes of that type], all INT32_MAX on entry and again on exit.
 * kgb_frontier_count / kgb_frontier_add synchronise the stream once (a size has to reach the host). */
KGB_API size_t kgb_frontier_workspace_bytes(int64_t n);
KGB_API int kgb_frontier_count(const int32_t* ptr, const int32_t* frontier, int32_t n_f, int32_t* offsets,
                               int32_t* h_total, void* workspace, size_t workspace_bytes, kgb_stream_t stream);
KGB_API int kgb_frontier_expand(const int32_t* ptr, const int32_t* col, const int32_t* eperm, const int32_t* frontier,
                                const int32_t* offsets, int32_t n_f, int32_t total, int32_t* eids, int32_t* srcs,
                                kgb_stream_t stream);
KGB_API int kgb_frontier_add(const int32_t* cand, int32_t n, int32_t* local, int32_t* firstpos, int32_t count_base,
                             int32_t* new_nodes, int32_t* h_n_new, void* workspace, size_t workspace_bytes,
                             kgb_stream_t stream);

/* ---- small fused elementwise / reduction helpers ---------------------------------- */
/* g[i] = dy[i] * (y[i] > 0)          backward of `x.relu()` (kgwas/model.py:75)          */
KGB_API int kgb_relu_bwd(const float* dy, const float* y, float* g, int64_t n, kgb_stream_t stream);
/* Backward through the fused ReLU of a layer output, the single-output head, and the bias
 * column sums, in ONE pass over the [m, h] rows (autograd of kgwas/model.py:75,83-86 plus the
 * lin_l.bias gradient of every relation into this node type):
 *   g[i,:]    = scale * (y[i,:] > 0) * ( dy[i,:] + dp[i] * wv[:] )     dy, dp nullable (not both);
 *                                                                       y nullable (no mask)
 *   sums[0,:] = sum_i g[i,:]                                            (bias gradient)
 *   sums[1,:] = sum_i dp[i] * y[i,:]                                    (head weight gradient; needs dp, y)
 * sums is [2, h] (nullable).  Two-stage deterministic reduction; workspace from
 * kgb_relu_bwd_fused_workspace_bytes(). */
KGB_API size_t kgb_relu_bwd_fused_workspace_bytes(int64_t m, int32_t h);
KGB_API int kgb_relu_bwd_fused(const float* dy, int64_t lddy, const float* y, int64_t ldy,
                               const float* dp, const float* wv, float scale, float* g, int64_t ldg,
                               int64_t m, int32_t h, float* sums, void* workspace,
                               size_t workspace_bytes, kgb_stream_t stream);
/* out[s, :] = beta*out[s, :] + sum_m w[m, s] * x[m, :]   (w NULL -> plain column sum, n_slots 1)
 * db_l of SAGEConv.lin_l / GATConv.bias, and d(att-folded vectors) of GATConv. */
KGB_API size_t kgb_wcolsum_workspace_bytes(int64_t m, int32_t n_slots, int32_t h);
KGB_API int kgb_wcolsum(const float* x, int64_t ldx, const float* w, int64_t ldw, int64_t m,
                        int32_t n_slots, int32_t h, float* out, float beta, void* workspace,
                        size_t workspace_bytes, kgb_stream_t stream);
/* a[i, s] = <x[i, s*slot_stride : s*slot_stride + h], v[s, :]>   node-level attention logits
 * alpha_src / alpha_dst (kgwas/conv.py:150-151) for n_slots relations at once
 * (slot_stride 0: every slot reads the same row). */
KGB_API int kgb_rowdot(const float* x, int64_t ldx, int64_t n_rows, int32_t n_slots, int32_t h,
                       int64_t slot_stride, const float* v, float* a, int64_t lda,
                       kgb_stream_t stream);
/* y[i, :] = beta*y[i, :] + sum_s a[i, s] * v[s, :]   (backward of kgb_rowdot w.r.t. x) */
KGB_API int kgb_rank_update(const float* a, int64_t lda, int32_t n_slots, const float* v, float* y,
                            int64_t ldy, int64_t n_rows, int32_t h, float beta, kgb_stream_t stream);
/* out[j] = w[perm[j]]  (edge weights CSR order -> COO order / transposed order) */
KGB_API int kgb_permute_f32(const float* w, const int32_t* perm, float* out, int64_t n,
                            kgb_stream_t stream);

/* ---- GAT attention (kgwas/conv.py:150-151, 200-228) ------------------------------- */
/* `groups` is a CSR whose row g = t*n_slots + k is one softmax group (destination node t,
 * relation slot k) and whose slot order equals the slot order of the job's aggregation CSR.
 *   u_j = a_src[src_is_node ? col[j]*n_slots + k : col[j]] + a_dst[g]      (conv.py:205)
 *   z_j = leaky_relu(u_j, negative_slope)                                  (conv.py:217)
 *   SOFTMAX: alpha_j = exp(z_j/T - max_g) / (sum_g exp(.) + 1e-16)         (conv.py:223)
 *   SIGMOID: alpha_j = sigmoid(z_j/T)                                      (conv.py:220)
 *   RAW    : alpha_j = z_j   (`return_raw_attention_weights`)              (conv.py:222)
 * alpha [E] is written in slot order: it is the edge weight of the following kgb_spmm
 * (conv.py:227-228), the saved tensor of the backward pass and, permuted by eperm, the
 * `return_attention_weights` output (conv.py:192-194). */
enum { KGB_ATT_SOFTMAX = 0, KGB_ATT_SIGMOID = 1, KGB_ATT_RAW = 2 };
KGB_API size_t kgb_gat_scratch_bytes(int32_t n_hrows, int32_t n_hsegs);
KGB_API int kgb_gat_alpha(const kgb_csr_t* groups, const float* a_src, const float* a_dst,
                          int32_t n_slots, int32_t src_is_node, float* alpha, float negative_slope,
                          float temperature, int32_t mode, void* scratch, size_t scratch_bytes,
                          kgb_stream_t stream);
/* out[j] = <xrow[i, :], x[col[j], :]> for every slot j of row i: d(alpha_j) of the weighted
 * aggregation (dalpha_e = <G[dst(e)], H_s[src(e)]>). */
KGB_API int kgb_sddmm(const kgb_csr_t* csr, const float* xrow, int64_t ldr, const float* x, int64_t ldx,
                      int32_t h, float* out, kgb_stream_t stream);
/* Backward of kgb_gat_alpha:  S_g = sum_j alpha_j dalpha_j
 *   SOFTMAX dz_j = alpha_j (dalpha_j - S_g)/T | SIGMOID dz_j = alpha_j(1-alpha_j) dalpha_j/T | RAW dz_j = dalpha_j
 *   du_j = dz_j * (u_j > 0 ? 1 : negative_slope);   da_dst[g] = sum_j du_j;   du [E] in slot order
 * (d a_src is the row sum of du over the transposed CSR: kgb_spmm's rowsum2). */
KGB_API int kgb_gat_dsoftmax(const kgb_csr_t* groups, const float* a_src, const float* a_dst,
                             int32_t n_slots, int32_t src_is_node, const float* alpha,
                             const float* dalpha, float* du, float* da_dst, float negative_slope,
                             float temperature, int32_t mode, void* scratch, size_t scratch_bytes,
                             kgb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* KGWAS_B200_H_ */
