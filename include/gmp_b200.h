/* gmp_b200.h -- C ABI of the B200-native geometric message-passing core (libgmp_b200.so).
 *
 * This is the drop-in boundary for the hot path of NW-JEFF/Geometric-Message-Passing
 * (per-edge geometric message + scatter-sum aggregation inside the EGNN / SchNet / TFN / MACE
 * layers).  The reference has no FFI of its own: the native arithmetic on this path is reached
 * through third-party wheels.  Each entry point below names the reference call site (file:line
 * under /root/reference) whose third-party call it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless marked "host";
 *   - no allocation inside the library: callers pass outputs and workspaces;
 *   - all work is enqueued on `stream` (a cudaStream_t); nothing synchronises unless documented;
 *   - return value: GMP_OK (0) or a negative gmp_status; gmp_last_error() gives the message;
 *   - "CSR" = (rowptr int32[n+1], col int32[E], perm int32[E]): edges stably sorted by their
 *     aggregation index; col[k] is the gather-side node of sorted edge k, perm[k] its position in
 *     the caller's original edge_index.  Aggregation is a deterministic, atomics-free segmented
 *     reduction over this order.
 *   - fp32 row-major tensors; nn.Linear weights keep PyTorch's [out_features, in_features] layout.
 */
#ifndef GMP_B200_H
#define GMP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* gmp_stream_t; /* == cudaStream_t */

enum gmp_status {
    GMP_OK = 0,
    GMP_ERR_INVALID_ARGUMENT = -1,
    GMP_ERR_CUDA = -2,
    GMP_ERR_UNSUPPORTED = -3
};

/* precision modes of the fused edge kernels */
enum gmp_precision {
    GMP_FP32_STRICT = 0, /* fp32 FFMA on CUDA cores: matches the reference layers to 1e-5 relative */
    GMP_BF16_TC = 1      /* bf16 operands on tcgen05 tensor cores, fp32 accumulate in TMEM: 1e-2 relative */
};

int gmp_version(void);
const char* gmp_last_error(void); /* host string, valid until the next failing call on this thread */

/* ============================================================================================ */
/* Graph construction (integer path, bit-exact)                                                  */
/* ============================================================================================ */

/* torch_cluster.radius_graph in its CUDA canonical order (never called by the reference itself --
 * models/schnet.py:66-72 bypasses PyG's RadiusInteractionGraph -- but required by BASELINE.json).
 * One query per node, candidates scanned in ascending index order inside the node's own example,
 * fp32 squared distance accumulated x,y,z with FMA contraction, strict `<  r*r`, at most
 * max_num_neighbors (+1 when loop == 0, for the self match) hits kept, self loops then dropped.
 * Two phases so that the caller can size the edge list:
 *   count: deg[i]  = number of edges whose destination is i
 *   fill : writes edge_src / edge_dst (int64, dst-major, src ascending) using rowptr = exclusive
 *          scan of deg (rowptr[n] = E).
 * graph_ptr: int64[num_graphs+1], node ranges of the examples (batch must be sorted). */
int gmp_radius_graph_count(const float* pos, const int64_t* graph_ptr, int64_t num_graphs, int64_t n,
                           float r, int32_t max_num_neighbors, int32_t loop, int32_t* deg,
                           gmp_stream_t stream);
int gmp_radius_graph_fill(const float* pos, const int64_t* graph_ptr, int64_t num_graphs, int64_t n,
                          float r, int32_t max_num_neighbors, int32_t loop, const int64_t* rowptr,
                          int64_t* edge_src, int64_t* edge_dst, gmp_stream_t stream);

/* Same result for ONE large example (BASELINE config 5) through a uniform cell list of edge >= r.
 * cell_of int32[n] and cell_start int32[ncells+1] / cell_nodes int32[n] are produced by
 * gmp_cells_build; origin/dims are host arrays (dims[0]*dims[1]*dims[2] = ncells).
 * max_num_neighbors semantics are identical (lowest indices first), so the output equals the
 * brute-force entry points bit for bit. */
int gmp_cells_build(const float* pos, int64_t n, float cell, const float* origin_host, const int32_t* dims_host,
                    int32_t* cell_of, int32_t* cell_start, int32_t* cell_nodes, int32_t* cursor_ws,
                    gmp_stream_t stream);
int gmp_radius_cells_count(const float* pos, int64_t n, float r, float cell, const float* origin_host,
                           const int32_t* dims_host, const int32_t* cell_start, const int32_t* cell_nodes,
                           int32_t max_num_neighbors, int32_t loop, int32_t* deg, gmp_stream_t stream);
int gmp_radius_cells_fill(const float* pos, int64_t n, float r, float cell, const float* origin_host,
                          const int32_t* dims_host, const int32_t* cell_start, const int32_t* cell_nodes,
                          int32_t max_num_neighbors, int32_t loop, const int64_t* rowptr,
                          int64_t* edge_src, int64_t* edge_dst, gmp_stream_t stream);

/* exclusive prefix sum, int32 -> int64, out has n+1 entries (out[n] = total). ws: int64[ceil(n/1024)+1]. */
int gmp_exclusive_scan_i32(const int32_t* in, int64_t n, int64_t* out, int64_t* ws, gmp_stream_t stream);

/* Stable counting sort of edges by an int64 index (replaces the implicit ordering torch_scatter's
 * atomics ignore: models/layers/egnn_layer.py:77,79; models/layers/tfn_layer.py:87).
 *   count : counts[i] = #edges with index == i                  (counts must hold n int32)
 *   fill  : perm = stable argsort(index)  given rowptr = exclusive scan of counts (int32[n+1]);
 *           cursor_ws int32[n], tmp_ws int32[E] are scratch.
 *   gather_i64_to_i32: out[k] = (int32) src[perm[k]]   (builds `col` from the other edge_index row)
 *   is_sorted: *flag (device int32) = 1 iff index is non-decreasing. */
int gmp_csr_count(const int64_t* index, int64_t num_edges, int64_t n, int32_t* counts, gmp_stream_t stream);
int gmp_csr_fill(const int64_t* index, int64_t num_edges, int64_t n, const int32_t* rowptr, int32_t* cursor_ws,
                 int32_t* tmp_ws, int32_t* perm, gmp_stream_t stream);
int gmp_gather_i64_to_i32(const int64_t* src, const int32_t* perm, int64_t num_edges, int32_t* out,
                          gmp_stream_t stream);
int gmp_index_is_sorted(const int64_t* index, int64_t num_edges, int32_t* flag, gmp_stream_t stream);

/* Coalescing of an edge list (torch_geometric.utils.to_undirected / coalesce, called for every reference dataset at
 * experiments/utils/create_graphs.py:79,158,249,330): the pairs are first sorted lexicographically by (row, col) with two
 * stable counting-sort passes (gmp_csr_count / gmp_csr_fill: by col, then by row); `perm` is the composed permutation
 * (NULL = already in order).
 *   mark_unique_pairs: keep[k] = 1 iff sorted pair k differs from sorted pair k-1  (int32[E])
 *   compact_pairs    : out_row/out_col[pos[k]] = pair k for every kept k, pos = exclusive scan of keep (int64[E+1]). */
int gmp_mark_unique_pairs(const int64_t* row, const int64_t* col, const int32_t* perm, int64_t num_edges, int32_t* keep,
                          gmp_stream_t stream);
int gmp_compact_pairs(const int64_t* row, const int64_t* col, const int32_t* perm, const int32_t* keep, const int64_t* pos,
                      int64_t num_edges, int64_t* out_row, int64_t* out_col, gmp_stream_t stream);

/* ============================================================================================ */
/* Segmented reductions (torch_scatter.scatter / scatter_sum)                      */
/* ============================================================================================ */

/* out[r,:] = reduce_{k in [rowptr[r], rowptr[r+1])} src[perm ? perm[k] : k, :]     reduce = sum | mean.
 * Rows without edges are written as zeros (dim_size = n semantics).  F % 4 == 0.
 * Replaces torch_scatter.scatter at models/layers/egnn_layer.py:77,79,147 and tfn_layer.py:87. */
int gmp_segment_reduce_f32(const int32_t* rowptr, const int32_t* perm, const float* src, float* out,
                           int64_t n, int32_t F, int32_t mean, gmp_stream_t stream);

/* Same reduction for bf16 rows (fp32 accumulation and output), F = 128: the per-edge d(pre1) rows of the fused EGNN backward. */
int gmp_segment_sum_bf16_f32(const int32_t* rowptr, const int32_t* perm, const void* src_bf16, float* out, int64_t n,
                             int32_t F, gmp_stream_t stream);

/* K0: out[r,:] = sum_k x[col[k],:] * (w ? w[perm ? perm[k] : k, :] : 1).  The CFConv message with a
 * materialised filter (PyG CFConv.message, called at models/schnet.py:72).  F % 4 == 0. */
int gmp_gather_mul_segsum_f32(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const float* x,
                              const float* w, float* out, int64_t n, int32_t F, gmp_stream_t stream);

/* K0 with the per-edge factor stored as bf16 rows (gathered rows x fp32 or, with x_is_bf16, bf16; fp32 accumulation and output;
 * F = 128): dL/dx1 of the CFConv from the filter values the forward pass kept (gmp_schnet_cfconv_fwd_tc2_keep).
 * col / w_bf16 may be NULL for an edgeless graph (out = 0). */
int gmp_gather_mul_segsum_wbf16(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const void* x,
                                int32_t x_is_bf16, const void* w_bf16, float* out, int64_t n, int32_t F, gmp_stream_t stream);

/* out[k,:] = x[idx[k],:]  (backward of the reductions above; PyG propagate's index_select). */
int gmp_gather_rows_f32(const int32_t* idx, const float* x, float* out, int64_t num_rows, int32_t F,
                        gmp_stream_t stream);

/* deterministic sum of `nparts` partial buffers of `len` floats each: out[i] = sum_p part[p*len+i]. */
int gmp_reduce_partials_f32(const float* part, int32_t nparts, int64_t len, float* out, gmp_stream_t stream);
/* `count` independent reductions of that kind in one launch (host arrays: parts[k] is [nparts[k]][lens[k]], outs[k] is [lens[k]]). */
int gmp_reduce_partials_batch_f32(const float* const* parts, const int32_t* nparts, const int64_t* lens, float* const* outs,
                                  int32_t count, gmp_stream_t stream);

/* Edge lengths and their backward (models/schnet.py:66-67; models/tfn.py:171-172):
 *   fwd: dist[e] = || pos[src[e]] - pos[dst[e]] ||_2            (edge order of the caller)
 *   bwd: dpos[i] = sum_{e: src=i} g[e] u_e - sum_{e: dst=i} g[e] u_e,  u_e = (pos_src-pos_dst)/dist
 *        through the two CSRs (rows = src with perm_s, rows = dst with perm_d); no atomics. */
int gmp_edge_length_fwd(const float* pos, const int64_t* src, const int64_t* dst, int64_t num_edges, float* dist,
                        gmp_stream_t stream);
int gmp_edge_length_bwd(const float* pos, const int64_t* src, const int64_t* dst, const float* g_dist,
                        const int32_t* rowptr_s, const int32_t* perm_s, const int32_t* rowptr_d,
                        const int32_t* perm_d, int64_t n, float* dpos, gmp_stream_t stream);

/* TFN / MACE edge prologue (models/tfn.py:171-175 == models/mace.py:170-174), one pass over the edges:
 *   vec = pos[src] - pos[dst];  edge_sh [E,(L+1)^2] = e3nn SphericalHarmonics(normalize=True, 'component'), L <= 2;
 *   edge_feat [E,num_bessel] = BesselBasis(|vec|) * PolynomialCutoff(|vec|; r_max, p)  (models/mace_modules/radial.py). */
int gmp_edge_geometry_fwd(const float* pos, const int64_t* src, const int64_t* dst, int64_t num_edges,
                          int32_t max_ell, float r_max, int32_t num_bessel, float poly_p, float* edge_sh,
                          float* edge_feat, gmp_stream_t stream);

/* ============================================================================================ */
/* SchNet continuous-filter convolution (PyG InteractionBlock/CFConv called at models/schnet.py:72)*/
/* ============================================================================================ */
typedef struct gmp_schnet_filter {
    const float* w1; /* [F, G]   mlp.0.weight */
    const float* b1; /* [F]      mlp.0.bias   */
    const float* w2; /* [F, F]   mlp.2.weight */
    const float* b2; /* [F]      mlp.2.bias   */
    int32_t num_gaussians; /* G <= 64 */
    int32_t num_filters;   /* F in {64, 128} */
    float cutoff;          /* C(d) = 0.5 (cos(d pi / cutoff) + 1), no mask (PyG CFConv) */
    /* GaussianSmearing: rbf_k = exp(gauss_coeff * (d - gauss_offset[k])^2), gauss_offset = device float[G]
     * (the module's `offset` buffer, so the fp32 linspace values are the reference's own).
     * Used only when edge_attr == NULL (the fused path recomputes the expansion from d). */
    const float* gauss_offset;
    float gauss_coeff;
} gmp_schnet_filter;

/* agg[r,:] = sum_{k in row r} x1[col[k],:] * ( (ssp(rbf_k W1^T + b1) W2^T + b2) * C(d_k) )
 *   edge_weight: float[E] distances in the caller's edge order (read through perm)
 *   edge_attr  : float[E,G] in the caller's edge order, or NULL to recompute the Gaussian expansion
 * Per-edge filters never touch HBM.  Calling it with the transposed CSR and the output gradient in
 * place of x1 yields the gradient w.r.t. x1 (the message is symmetric in the two factors). */
int gmp_schnet_cfconv_fwd(const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t n,
                          int64_t num_edges, const float* edge_weight, const float* edge_attr, const float* x1,
                          const gmp_schnet_filter* filt /* host */, float* agg, int32_t precision,
                          gmp_stream_t stream);

/* number of partial-gradient slots (CTAs) gmp_schnet_cfconv_bwd writes for E edges */
int32_t gmp_schnet_bwd_num_parts(int64_t num_edges);
/* floats per slot: F*64 + F + F*F + F  (dW1 padded to 64 columns | db1 | dW2 | db2) */
int64_t gmp_schnet_bwd_part_len(int32_t num_gaussians, int32_t num_filters);

/* Filter-side backward over the dst-sorted CSR given g_agg = dL/dagg [n,F]:
 *   wgrad_parts[p] : per-CTA partial sums of (dW1 [F,Gp], db1 [F], dW2 [F,F], db2 [F]); reduce with
 *                    gmp_reduce_partials_f32 (deterministic).
 *   d_edge_weight  : float[E] (caller's edge order) or NULL;  d_edge_attr: float[E,G] or NULL
 *                    (when edge_attr == NULL the chain through the Gaussian expansion is folded into
 *                    d_edge_weight instead). */
int gmp_schnet_cfconv_bwd(const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t n,
                          int64_t num_edges, const float* edge_weight, const float* edge_attr, const float* x1,
                          const gmp_schnet_filter* filt /* host */, const float* g_agg, float* wgrad_parts,
                          float* d_edge_weight, float* d_edge_attr, int32_t precision, gmp_stream_t stream);

/* Hardware self test of the tcgen05 path: out[128,128] = bf16(A[128,K]) x bf16(B[128,K])^T with fp32 accumulation
 * in TMEM, through the same shared-memory descriptors and 128-byte swizzle the fused kernels use.  K in {64, 128}. */
int gmp_umma_selftest(const float* A, const float* B, float* out, int32_t K, gmp_stream_t stream);
/* Same for MN-major (transposed) operands: out[128,N] = sum_k bf16(A[k,m]) * bf16(B[k,n]), A [128,128], B [128,N], N in {64,128}
 * (the layout the weight-gradient GEMMs consume). */
int gmp_umma_selftest_mn(const float* A, const float* B, float* out, int32_t N, gmp_stream_t stream);

/* Pipelined GMP_BF16_TC forward of the same op (csrc/schnet_tc2.cu): warp-specialised (meta / MMA / two epilogue
 * groups), three tcgen05 products per tile -- the third one is the segmented row sum itself (msg^T x one-hot row
 * membership).  Restrictions: 128 filters, the Gaussian basis recomputed from edge_weight (edge_attr = NULL path),
 * x1 passed as bf16 rows.  agg [n,128] is overwritten (zeroed first); head [gmp_schnet_tc2_num_chunks(E),128] is
 * scratch for rows that straddle a CTA boundary; rowid int32[E] = CSR row of every sorted edge. */
int32_t gmp_schnet_tc2_num_chunks(int64_t num_edges);
int gmp_schnet_cfconv_fwd_tc2(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const int32_t* rowid,
                              int64_t n, int64_t num_edges, const float* edge_weight, const void* x1_bf16,
                              const gmp_schnet_filter* filter /* host */, float* agg, float* head, gmp_stream_t stream);

/* Same, and additionally stores the filter value of every edge, W(e) * C(e) as 128 bf16, into filter_out_bf16 [E,128] (may be
 * NULL = plain forward) at row filter_row[id] of the edge with caller's id (perm ? perm[k] : k); filter_row = NULL: at row id.
 * Training uses it so that the backward pass does not have to run the filter MLP again for dL/dx1 (PyG keeps the same [E,F]
 * tensor alive for autograd, CFConv.forward called at models/schnet.py:72); with filter_row = the position of each edge in
 * the source-sorted CSR, gmp_gather_mul_segsum_wbf16 then streams the factors in order (perm = NULL there). */
int gmp_schnet_cfconv_fwd_tc2_keep(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const int32_t* rowid,
                                   int64_t n, int64_t num_edges, const float* edge_weight, const void* x1_bf16,
                                   const gmp_schnet_filter* filter /* host */, float* agg, float* head,
                                   void* filter_out_bf16, const int32_t* filter_row, gmp_stream_t stream);

/* Pipelined GMP_BF16_TC variant of the filter-side backward (weight gradients of the filter MLP only; same partial layout
 * as gmp_schnet_cfconv_bwd: [dW1 128x64 | db1 | dW2 128x128 | db2] per CTA, `nparts` CTAs, each owning a contiguous range
 * of the sorted edges).  Restrictions: 128 filters, <= 63 lazily expanded Gaussians, x1 passed as bf16 rows. */
int gmp_schnet_cfconv_bwd_tc2(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const int32_t* rowid,
                              int64_t n, int64_t num_edges, const float* edge_weight, const void* x1_bf16,
                              const gmp_schnet_filter* filter /* host */, const float* g_agg, float* wgrad_parts,
                              int32_t nparts, gmp_stream_t stream);

/* Parameter gradients of a node-side nn.Linear (PyG CFConv.lin1 / lin2, InteractionBlock.lin; called at
 * models/schnet.py:72) in the GMP_BF16_TC mode: dW [out,in] = g^T x and db [out] = column sums of g, g [n,out], x [n,in],
 * out = 128, in in {64,128}.  Rows are split over the SMs; parts [gmp_linear_wgrad_num_parts(n)][out*in + out] holds
 * one partial (dW | db) per CTA: sum them with gmp_reduce_partials_f32. */
int32_t gmp_linear_wgrad_num_parts(int64_t n);
int gmp_linear_wgrad_tc(const float* g, const float* x, int64_t n, int32_t out_dim, int32_t in_dim, float* parts,
                        gmp_stream_t stream);

/* ============================================================================================ */
/* EGNN edge path (models/layers/egnn_layer.py:62-80: message + aggregate, fused)                 */
/* ============================================================================================ */
typedef struct gmp_egnn_edge_params {
    const float* wd;    /* [d]    mlp_msg.0.weight[:, 2d]   (the distance column of the first Linear) */
    const float* ln1_g; /* [d]    mlp_msg.1.weight */
    const float* ln1_b; /* [d]    mlp_msg.1.bias   */
    const float* w1;    /* [d,d]  mlp_msg.3.weight */
    const float* b1;    /* [d]    mlp_msg.3.bias   */
    const float* ln2_g; /* [d]    mlp_msg.4.weight */
    const float* ln2_b; /* [d]    mlp_msg.4.bias   */
    const float* w2;    /* [d,d]  mlp_pos.0.weight */
    const float* b2;    /* [d]    mlp_pos.0.bias   */
    const float* ln3_g; /* [d]    mlp_pos.1.weight */
    const float* ln3_b; /* [d]    mlp_pos.1.bias   */
    const float* w3;    /* [d]    mlp_pos.3.weight (shape [1,d]) */
    const float* b3;    /* [1]    mlp_pos.3.bias   */
    int32_t d;          /* emb_dim in {64, 128} */
    int32_t act;        /* 0 = relu, 1 = swish (SiLU) */
    float ln_eps;       /* LayerNorm eps (1e-5) */
    int32_t aggr_mean;  /* message aggregation: 0 = sum/add, 1 = mean (coordinates always use mean) */
} gmp_egnn_edge_params;

/* P = h W0[:, :d]^T + b0 and Q = h W0[:, d:2d]^T are the two node-side halves of mlp_msg's first Linear
 * (cat[h_i, h_j, dist] is never built).  CSR rows = edge_index[1] (i), col = edge_index[0] (j).
 *   msg_aggr[i,:] = reduce_e m_e ,  pos_aggr[i,:] = mean_e (pos_i - pos_j) * mlp_pos(m_e)            */
int gmp_egnn_edge_fwd(const int32_t* rowptr, const int32_t* col, int64_t n, int64_t num_edges, const float* P,
                      const float* Q, const float* pos, const gmp_egnn_edge_params* prm /* host */,
                      float* msg_aggr, float* pos_aggr, int32_t precision, gmp_stream_t stream);

int32_t gmp_egnn_bwd_num_parts(int64_t num_edges);
/* floats per slot: dW1 d*d | dW2 d*d | db1 | db2 | dln1_g | dln1_b | dln2_g | dln2_b | dln3_g | dln3_b | dw3 | dwd
 * (d each) | db3 (1) | 3 pad  = 2 d^2 + 10 d + 4 */
int64_t gmp_egnn_bwd_part_len(int32_t d);

/* Backward of gmp_egnn_edge_fwd given g_msg = dL/dmsg_aggr [n,d] and g_pos = dL/dpos_aggr [n,3]; the
 * forward is recomputed per tile.  Run it twice:
 *   src_pass = 0: (rowptr, col) = dst-sorted CSR -> d_node = dL/dP, d_pos = the pos_i part, wgrad_parts
 *   src_pass = 1: (rowptr, col) = src-sorted CSR (rows j, col i) -> d_node = dL/dQ, d_pos = the pos_j part
 * dst_rowptr is always the dst-sorted rowptr (in-degrees for the mean).  dL/dpos = sum of both d_pos. */
int gmp_egnn_edge_bwd(const int32_t* rowptr, const int32_t* col, const int32_t* dst_rowptr, int64_t n,
                      int64_t num_edges, const float* P, const float* Q, const float* pos,
                      const gmp_egnn_edge_params* prm /* host */, const float* g_msg, const float* g_pos,
                      int32_t src_pass, float* d_node, float* d_pos, float* wgrad_parts, int32_t precision,
                      gmp_stream_t stream);

/* ---- GMP_BF16_TC variant (csrc/egnn_tc.cu, emb_dim = 128): both edge GEMMs on tcgen05 (bf16 operands, fp32 accumulation
 * in tensor memory), 128-edge tiles; 1e-2 relative.  rowid int32[E] = CSR row of every sorted edge.  The col-side
 * operand is read as bf16 rows (256 B per edge instead of 512): Q_bf16 = bf16(Q) prepared by the caller. */
int gmp_egnn_tc_edge_fwd(const int32_t* rowptr, const int32_t* col, const int32_t* rowid, int64_t n, int64_t num_edges,
                         const float* P, const void* Q_bf16, const float* pos,
                         const gmp_egnn_edge_params* prm /* host */, float* msg_aggr, float* pos_aggr,
                         gmp_stream_t stream);
/* Second design of the same forward (csrc/egnn_tc2.cu; models/layers/egnn_layer.py:62-80): a thread owns a whole edge row
 * (LayerNorm statistics are thread-local, no barrier inside a tile), three independent tile streams per SM, aggregation
 * as a one-hot MMA with rows carried across tiles.  The sorted edge list is cut into gmp_egnn_tc2_num_chunks(E) contiguous
 * chunks; rows that straddle a chunk boundary are completed by a fix-up kernel from `head`
 * (float [num_chunks][132], scratch).  msg_aggr / pos_aggr are zeroed here (rows without edges stay zero). */
int32_t gmp_egnn_tc2_num_chunks(int64_t num_edges);
/* Destination-partitioned graph, gather fused with the halo transfer: with local numbering
 * [left halo | owned | right halo], Q_bf16 then holds the OWNED rows only and the rows of halo sources are read, inside the
 * kernel's gather, from the neighbouring ranks' Q arrays through peer-mapped pointers (NVLink loads; symmetric memory):
 * row j < n_left -> left + j, j < n_left + n_own -> Q_bf16 + (j - n_left), else right + (j - n_left - n_own), 128 bf16 each.
 * left / right already point at the first row this rank reads.  NULL = Q_bf16 holds every local row (single GPU). */
typedef struct gmp_peer_rows {
    const void* left;
    const void* right;
    int64_t n_left, n_own;
} gmp_peer_rows;
int gmp_egnn_tc2_edge_fwd(const int32_t* rowptr, const int32_t* col, const int32_t* rowid, int64_t n, int64_t num_edges,
                          const float* P, const void* Q_bf16, const float* pos,
                          const gmp_egnn_edge_params* prm /* host */, float* msg_aggr, float* pos_aggr, float* head,
                          const gmp_peer_rows* peer /* host, may be NULL */, gmp_stream_t stream);
/* Backward, two recompute passes as gmp_egnn_edge_bwd.  row_operand (fp32) / col_operand_bf16 are P / bf16(Q) in the
 * dst pass (src_pass = 0) and Q / bf16(P) in the src pass.  Both passes write per-CTA partials into the SAME
 * wgrad_parts [gmp_egnn_tc_bwd_num_parts(E)][gmp_egnn_bwd_part_len(128)]: the dst pass the two weight matrices
 * (dpre^T a1, dpre^T m accumulated in tensor memory), the src pass the ten vectors (column sums through the tensor
 * core) and db3; sum them with gmp_reduce_partials_f32. */
int32_t gmp_egnn_tc_bwd_num_parts(int64_t num_edges);
/* Single-pass variant: the dst pass computes every parameter gradient (full partial rows) and writes, per edge and in
 * the caller's edge order (perm = that of the dst-sorted CSR), d(pre1) as bf16 rows [E,128] and d(delta) as float4 [E];
 * dL/dQ = gmp_segment_sum_bf16_f32 over the src-sorted CSR, the pos_j part = -gmp_segment_reduce_f32(ddelta) likewise.
 * Costs 272 B of scratch per edge; the two-pass entry point below needs none. */
int gmp_egnn_tc_edge_bwd_fused(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const int32_t* rowid,
                               int64_t n, int64_t num_edges, const float* P, const void* Q_bf16, const float* pos,
                               const gmp_egnn_edge_params* prm /* host */, const float* g_msg, const float* g_pos,
                               float* dP, float* dpos_i, float* wgrad_parts, void* dpre1_bf16, float* ddelta,
                               const gmp_peer_rows* peer /* host, may be NULL: as in gmp_egnn_tc2_edge_fwd */, gmp_stream_t stream);
int gmp_egnn_tc_edge_bwd(const int32_t* rowptr, const int32_t* col, const int32_t* rowid, const int32_t* dst_rowptr,
                         int64_t n, int64_t num_edges, const float* row_operand, const void* col_operand_bf16,
                         const float* pos, const gmp_egnn_edge_params* prm /* host */, const float* g_msg,
                         const float* g_pos, int32_t src_pass, float* d_node, float* d_pos, float* wgrad_parts,
                         gmp_stream_t stream);

/* ============================================================================================ */
/* TFN / MACE tensor-product convolution (models/layers/tfn_layer.py:82-87)                       */
/* ============================================================================================ */

/* One contraction kernel serves the forward and the feature gradient (csrc/tpconv.cu):
 *   res[n, block] (+)= sum_{e in CSR row n} sum_a T_e[a,b] * ( sum_iA V[col_e][v_off + a*DA + iA] * Z_e[iA,kB] )
 * with T_e = fc(edge_feat_e) generated slice by slice in shared memory (the [E, weight_numel] tensor of the
 * reference never exists) and Z_e[iA,kB] = sum_j sh_e[j] * cg[iA][j][kB].
 *   forward   : CSR rows = edge_index[0] (aggregation), col = edge_index[1]; V = node_attr, res = TP output
 *   d/d node  : CSR rows = edge_index[1], col = edge_index[0];               V = dL/dout,  res = dL/dnode_attr
 * `passes` / `blocks` are device arrays of the 12-int32 / 8-int32 records documented in csrc/tpconv.cu
 * (struct TpPass / TpBlock), built once per layer by the host (gmp_b200/tfn.py) from the e3nn instruction
 * list; `cg` is the float table they index.  fc = Linear(R,H) -> ReLU -> Linear(H,numel): w1 [H,R], b1 [H],
 * w2 [numel,H], b2 [numel].  edge_sh [E,S] and edge_feat [E,R] are in the caller's edge order (perm). */
int gmp_tp_contract(const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t n, int64_t num_edges,
                    const float* V, int32_t v_len, float* res, int32_t r_len, const float* edge_sh, int32_t S,
                    const float* edge_feat, int32_t R, const float* w1, const float* b1, const float* w2,
                    const float* b2, int32_t H, const void* passes, const void* blocks, int32_t nblocks,
                    int32_t nunits, const float* cg, int32_t precision, gmp_stream_t stream);
int64_t gmp_tp_contract_smem_bytes(int32_t H, int32_t R);

/* Weight gradient of fc given g = dL/d(TP output) [n, g_len]: a CTA owns 64 rows of w2 (`units`: device array of
 * 12-int32 TpWUnit records) and streams all edges (CSR rows = edge_index[0], col = edge_index[1]).
 *   dW2 [numel,H], db2 [numel] are written directly; dW1/db1 come as per-unit partials
 *   w1_parts [nunits, H*16 + H] (dW1 padded to 16 columns | db1) to be summed by gmp_reduce_partials_f32. */
int gmp_tp_wgrad(const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t n, int64_t num_edges,
                 const float* x, int32_t x_len, const float* g, int32_t g_len, const float* edge_sh, int32_t S,
                 const float* edge_feat, int32_t R, const float* w1, const float* b1, const float* w2, int32_t H,
                 const void* units, int32_t nunits, const float* cg, float* dW2, float* db2, float* w1_parts,
                 int32_t precision, gmp_stream_t stream);
int64_t gmp_tp_wgrad_part_len(int32_t H);

/* ---- GMP_BF16_TC variant of the same layer (csrc/tpconv_tc.cu): fc's second Linear runs on tcgen05 (bf16 operands,
 * fp32 accumulation in tensor memory; 1e-2 relative), the generated weights are consumed out of tensor memory.
 * The operands are staged once per call as UMMA shared-memory images:
 *   hid_img : relu(w1 edge_feat + b1) of every 128-edge tile of the CSR order, bf16  (gmp_tp_tc_hid_bytes bytes)
 *   w2_img  : the rows of w2 of every 256-column "N-tile", bf16, in consumption order (gmp_tp_tc_w2_bytes bytes)
 * `ntile_table` [ntiles_n] (8 int32, struct TcNTile) and `ygroups` [nyg] (16 int32, struct TcYGroup) are device
 * tables built by the host from the e3nn instruction list (gmp_b200/tfn.py).  `res` [n, r_len] is overwritten;
 * `head` [gmp_tp_tc_num_chunks(E), r_len] is scratch.  fc's second bias is NOT applied here: it contributes
 * sum_a b2[a,b] YS[n][a,k] with YS from gmp_tp_ysum (fp32), a node-level GEMM the caller adds. */
int32_t gmp_tp_tc_num_chunks(int64_t num_edges);
int64_t gmp_tp_tc_hid_bytes(int64_t num_edges, int32_t H);
int64_t gmp_tp_tc_w2_bytes(int32_t ntiles_n, int32_t H);
int gmp_tp_tc_pack_hid(const int32_t* perm, int64_t num_edges, const float* edge_feat, int32_t R, const float* w1,
                       const float* b1, int32_t H, void* hid_img, gmp_stream_t stream);
int gmp_tp_tc_pack_w2(const float* w2, int32_t H, const void* ntile_table, int32_t ntiles_n, void* w2_img,
                      gmp_stream_t stream);
int gmp_tp_tc_contract(const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t n, int64_t num_edges,
                       const float* V, int32_t v_len, float* res, int32_t r_len, float* head, const float* edge_sh,
                       int32_t S, const void* hid_img, const void* w2_img, const void* ygroups, int32_t nyg,
                       int32_t ntiles_n, int32_t H, const float* cg, gmp_stream_t stream);
/* Parameter gradients of fc on the tensor cores (CSR rows = edge_index[0], col = edge_index[1]; x gathered at col,
 * g = dL/d(TP output) read at the row node; tables of the forward orientation):
 *   gmp_tp_tc_dhid : dpre[e, :] = relu'(pre_e) * sum_c dT_e[c] w2[c, :]   (caller's edge order; dW1 = dpre^T edge_feat,
 *                    db1 = column sums follow as plain GEMM / reduction)
 *   gmp_tp_tc_dw2  : dW2[c, :] = sum_e dT_e[c] hid_e[:]; `wtile_table` [ntiles_n] (16 int32, struct TcWTile), one CTA each
 * dT_e[(a,b)] = sum_k g[row_e][b,k] Y_e[a,k] is generated per tile as a bf16 UMMA operand and never stored. */
int gmp_tp_tc_dhid(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const int32_t* rowid, int64_t n,
                   int64_t num_edges, const float* x, int32_t x_len, const float* g, int32_t g_len, const float* edge_sh,
                   int32_t S, const float* edge_feat, int32_t R, const float* w1, const float* b1, const void* w2_img,
                   const void* ygroups, int32_t nyg, int32_t ntiles_n, int32_t H, const float* cg, float* dpre,
                   gmp_stream_t stream);
/* rowid int32[E]: CSR row of every sorted edge.  dW2 [ngroups][numel][H]: the edge tiles are split into `ngroups`
 * contiguous groups (grid = ntiles_n * ngroups CTAs, chosen by the caller to fill whole waves of SMs); each group
 * writes its own partial, the caller sums them in group order. */
int gmp_tp_tc_dw2(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const int32_t* rowid, int64_t n,
                  int64_t num_edges, const float* x, int32_t x_len, const float* g, int32_t g_len, const float* edge_sh,
                  int32_t S, const void* hid_img, const void* wtile_table, int32_t ntiles_n, int32_t H, const float* cg,
                  int64_t numel, int32_t ngroups, float* dW2, gmp_stream_t stream);
/* YS[n][y_off_p + a*DB_p + k] = sum_{e in CSR row n} sum_i V[col_e][v_off_p + a*DA_p + i] * Z^p_e[i][k], fp32:
 * the node-level aggregate in which both the bias term of the layer and db2 are linear.
 * `ypaths` [npaths] (8 int32, struct TcYPath), `zentries` [nz] (4 int32, struct TcZEntry); npairs = sum_p MA_p. */
int gmp_tp_ysum(const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t n, int64_t num_edges,
                const float* V, int32_t v_len, const float* edge_sh, int32_t S, const void* ypaths, int32_t npaths,
                int32_t npairs, const void* zentries, int32_t nz, const float* cg, float* YS, int32_t y_len,
                gmp_stream_t stream);

/* ============================================================================================ */
/* MACE symmetric contraction (models/mace_modules/symmetric_contraction.py:169-185)              */
/* ============================================================================================ */

/* out[b, C*off(K) + c*d(K) + k(K)] = sum_m coef[c][K][m] * x[b,c,i1(m)] * x[b,c,i2(m)] * x[b,c,i3(m)]
 *   x    [N, C, D]   node features after reshape_irreps (models/mace_modules/irreps_tools.py:69-79), D <= 15
 *   coef [C, K, M]   U matrices folded with the per-channel weights over the symmetric monomial basis (host, autograd)
 *   mono [M, 3]      component indices, value D = the constant 1 (pads monomials of degree < 3); M <= 256
 *   out_map [K, 3]   (component offset of the output irrep block, its dim d, local k), K <= 16
 * Replaces the three opt_einsum.contract calls per output irrep of Contraction.forward. */
/* 1 when the fully unrolled kernels for the model shape take the call (D = 9 components of l <= 2 features, K = 9,
 * correlation 3 = 219 monomials, irrep-major output of length 9 C), 0 when the generic kernels do. */
int32_t gmp_symcontract_fast_path(int32_t C, int32_t D, int32_t K, int32_t M, int32_t out_len);
int gmp_symcontract_fwd(const float* x, const float* coef, const int32_t* mono, const int32_t* out_map,
                        int64_t num_nodes, int32_t C, int32_t D, int32_t K, int32_t M, float* out, int32_t out_len,
                        gmp_stream_t stream);
/* dx [N,C,D] and per-block partials dcoef_parts [nparts, C, K, M] (sum with gmp_reduce_partials_f32). */
int32_t gmp_symcontract_bwd_num_parts(int64_t num_nodes);
int gmp_symcontract_bwd(const float* x, const float* coef, const int32_t* mono, const int32_t* out_map,
                        int64_t num_nodes, int32_t C, int32_t D, int32_t K, int32_t M, const float* g_out,
                        int32_t out_len, float* dx, float* dcoef_parts, gmp_stream_t stream);

/* ============================================================================================ */
/* ACEsuit-style MACE interaction: 'uvu' tensor product + scatter_sum           */
/* ============================================================================================ */

/* message[i] = sum_{e: receiver_e = i} TP_uvu(node_feats[sender_e], edge_attrs_e, w_e): replaces
 *   mji = self.conv_tp(node_feats[sender], edge_attrs, tp_weights); scatter_sum(mji, receiver, dim=0, dim_size=N)
 * (models/mace_modules/blocks.py:257-263, 319-325, 384-390, 446-455, 516-525) for node features C x (0e + 1o + 2e),
 * edge attributes 0e + 1o + 2e and the 11 'uvu' instructions of irreps_tools.py:14-44 (mid blocks sorted by irrep).
 *   rowptr/col/perm  CSR over the receivers (rows = edge_index[1], col = sender; perm = position in the caller's edge
 *                    order or NULL when already sorted)
 *   x          [N, 9 C]    e3nn layout          edge_attrs [E, 9]       w [E, 11 C] (instruction order, [path][channel])
 *   out        [N, 35 C]   3 x (C x 0e) + 4 x (C x 1o) + 4 x (C x 2e), every element written once (empty rows = 0) */
int gmp_uvu_conv_fwd(const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t num_nodes, int64_t num_edges,
                     const float* x, const float* edge_attrs, const float* w, int32_t C, float* out, gmp_stream_t stream);
/* dL/dx [N, 9 C] from g = dL/dmessage [N, 35 C]; the CSR is the transposed one (rows = senders, col = receivers). */
int gmp_uvu_conv_dx(const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t num_nodes, int64_t num_edges,
                    const float* g, const float* edge_attrs, const float* w, int32_t C, float* dx, gmp_stream_t stream);
/* dL/dw [E, 11 C] in the caller's edge order; sender / receiver = the two rows of edge_index (int64). */
int gmp_uvu_conv_dw(const int64_t* sender, const int64_t* receiver, int64_t num_edges, const float* x, const float* g,
                    const float* edge_attrs, int32_t C, float* dw, gmp_stream_t stream);

/* ============================================================================================ */
/* Node-side dense chains on tcgen05 (GMP_BF16_TC): nn.Linear layers with fused epilogues         */
/* ============================================================================================ */

/* Replaces, per 128-row tile and without intermediate HBM round trips, chains of up to three 128-wide nn.Linear layers
 * and their elementwise neighbours:
 *   SchNet  CFConv.lin2 -> ShiftedSoftplus -> InteractionBlock.lin -> residual -> next CFConv.lin1   (PyG blocks built at
 *           models/schnet.py:41-54, residual at :72) and the transposed chain of the backward pass;
 *   EGNN    mlp_upd = Linear(2d, d), LayerNorm, act, Linear(d, d), LayerNorm, act   (models/layers/egnn_layer.py:41-48, 82-86);
 *   single layers (EGNN P / Q projections, o3.Linear blocks).
 * Stage s computes  v = A_s W_s^T + bias;  [out_pre = v];  v = LayerNorm(v) * ln_g + ln_b;  v = act(v);
 *                   v *= f(mul_aux);  v += add_res;  [out_f32 = v];  [out_bf16 = bf16(v)];  A_{s+1} = bf16(v)
 * with bf16 operands and fp32 accumulation.  A_0 = a0 (fp32 [n,128]) or cat[a0, a1] along the columns (K = 256). */
enum { GMP_NODE_ACT_NONE = 0, GMP_NODE_ACT_SSP = 1, GMP_NODE_ACT_RELU = 2, GMP_NODE_ACT_SILU = 3 };
enum { GMP_NODE_MUL_PLAIN = 0, GMP_NODE_MUL_DSSP = 1 };   /* DSSP: multiply by 1 - exp(-(aux + ln 2)) = ssp'(pre) given aux = ssp(pre) */
typedef struct gmp_node_stage {
    const void* w_img;      /* operand image from gmp_node_pack_w (stage 0 with two sources: both images, contiguous) */
    const float* bias;      /* [128] or NULL */
    const float* ln_g;      /* LayerNorm weight / bias [128], both or neither */
    const float* ln_b;
    float ln_eps;
    int32_t act;            /* GMP_NODE_ACT_* */
    const float* mul_aux;   /* [n,128] or NULL */
    int32_t mul_mode;       /* GMP_NODE_MUL_* */
    const float* add_res;   /* [n,128] or NULL */
    float* out_f32;         /* [n,128] or NULL */
    void* out_bf16;         /* [n,128] bf16 or NULL */
    float* out_pre;         /* [n,128] pre-LayerNorm values (for the backward pass) or NULL */
} gmp_node_stage;
/* bytes of the operand image of W [out_dim, in_dim] (fp32, row-major, nn.Linear layout); -1 if the shape is not served.
 * transpose = 0: the image of W itself (y = x W^T, out_dim = 128, in_dim = 128 or 256);
 * transpose = 1: the image of W^T (dx = g W, in_dim = 128, out_dim = 128 or 256). */
int64_t gmp_node_w_image_bytes(int32_t out_dim, int32_t in_dim, int32_t transpose);
int gmp_node_pack_w(const float* w, int32_t out_dim, int32_t in_dim, int32_t transpose, void* img, gmp_stream_t stream);
/* `count` weights of shape [128, 128] in one launch (host arrays of pointers / flags; images of 32 KB each). */
int gmp_node_pack_w_batch(const float* const* w, const int32_t* transpose, void* const* img, int32_t count, gmp_stream_t stream);
int gmp_node_chain_tc(const float* a0, const float* a1, int64_t num_rows, int32_t nstage, const gmp_node_stage* stages,
                      gmp_stream_t stream);

/* Backward of a = act(LayerNorm(pre) * gamma + beta) over rows of 128 (EGNN mlp_upd, models/layers/egnn_layer.py:41-48):
 * d_pre [n,128]; optional act_out [n,128] = the recomputed activations; parts [num_parts, 256] = per-CTA partial
 * [d gamma | d beta] (sum with gmp_reduce_partials_f32).  act: GMP_NODE_ACT_NONE / _RELU / _SILU. */
int32_t gmp_ln_act_bwd_num_parts(int64_t num_rows);
int gmp_ln_act_bwd(const float* g_out, const float* pre, const float* gamma, const float* beta, float eps, int32_t act,
                   int64_t num_rows, float* d_pre, float* act_out, float* parts, gmp_stream_t stream);

/* ============================================================================================ */
/* e3nn Gate / scalar Activation of the TFN layer (models/layers/tfn_layer.py:45-63, 89-90)       */
/* ============================================================================================ */

/* x [n, ns + ng + nv] = [scalars | gates | gated] -> out [n, ns + nv] = [silu(s) c_silu | gated_j sigmoid(gate_expand[j]) c_sigmoid]
 * (c_* = e3nn's normalize2mom constants).  expand [nv]: gate index of every gated element; gate_start / gate_dim [ng]: the
 * contiguous range of gated elements each gate multiplies.  ng = nv = 0 is the all-scalar Activation. */
int gmp_gate_fwd(const float* x, const int32_t* expand, int64_t num_rows, int32_t num_scalars, int32_t num_gates,
                 int32_t num_gated, float c_silu, float c_sigmoid, float* out, gmp_stream_t stream);
int gmp_gate_bwd(const float* x, const float* g_out, const int32_t* expand, const int32_t* gate_start, const int32_t* gate_dim,
                 int64_t num_rows, int32_t num_scalars, int32_t num_gates, int32_t num_gated, float c_silu, float c_sigmoid,
                 float* dx, gmp_stream_t stream);

/* ============================================================================================ */
/* Destination-partitioned graph: halo rows over peer memory                 */
/* ============================================================================================ */

/* Copies (add_f32 = 0) or adds as fp32 (add_f32 = 1) up to 16 segments src[k] -> dst[k] of bytes[k] bytes (multiples of 4;
 * 16-byte vector path when every pointer and size is 16-byte aligned) in one launch.  src pointers may be peer-GPU
 * addresses mapped into this process (symmetric memory): the loads then travel over NVLink.  The caller orders the launch
 * after the peers' writes with a cross-GPU barrier.  Replaces the ncclSend / ncclRecv halo exchange + concatenation of the
 * partitioned EGNN (there is no reference counterpart: the reference is single-process). */
int gmp_halo_pull(const void* const* src, void* const* dst, const int64_t* bytes, int32_t num_segments, int32_t add_f32,
                  gmp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GMP_B200_H */
