/*
 * glab.h -- C ABI of libglab_b200.so: the B200 (sm_100a) edge-wise message-passing hot path
 * behind the GNN-as-linear-algebra layers of sandialabs/gnn-applied-linear-algebra.
 *
 * The reference has no FFI of its own: its seam is the Python call
 *     torch_scatter.scatter(src, edgeij_pair[0], dim=0, dim_size=n, reduce=...)
 * inside each layer's edge->vertex aggregation, wrapped by torch_geometric.nn.MetaLayer
 * (gather x[row], x[col] -> edge update -> aggregation -> vertex update -> global update).
 * Every entry point below replaces one such (gather + edge update + scatter + vertex update)
 * chain with ONE fused kernel launch; the citation on each function names the reference
 * lines it replaces (paths relative to <reference>/pytorch/).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - every function returns int: 0 = ok, >0 = cudaError_t, <0 = GLAB_E_* argument error.
 *     Nothing throws, nothing aborts, nothing synchronises the device unless documented.
 *   - calls operate on the CURRENT CUDA device, which must be the device the plan / pointers
 *     live on (one process per GPU is the intended model; several devices per process work as
 *     long as the caller switches devices).  The library keeps no data-path state besides
 *     opaque plans; calls are safe from one host thread per device.
 *   - all data pointers are BORROWED DEVICE pointers (owned by the caller, e.g. torch tensors);
 *     the library allocates only inside opaque plans (glab_plan_*, glab_halo_*).
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream).
 *   - dense vectors are row-major [n, k] with leading dimension k (k = number of RHS columns,
 *     k in {1,2,4,8}); per-edge arrays are in CSR slot order unless the name says `edge order`.
 *   - scalars that the reference keeps in the global attribute tensor `g` (omega, alpha, beta,
 *     1/norm) are read through DEVICE pointers so no step ever needs a host sync.
 *   - ALIGNMENT AND PADDING of caller arrays.  The pipeline kernels move CSR-ordered per-edge arrays
 *     (vals, S, aux) and dense vectors with 16-byte bulk copies whose source range is widened to
 *     16-byte granules: they may READ up to 15 bytes before the first and after the last element a
 *     call needs (never write).  Every such array must therefore live in an allocation that starts
 *     16-byte aligned and whose size is a multiple of 16 bytes -- true for cudaMalloc, torch and
 *     every pooled allocator (256-byte or larger granules); a C caller that sub-allocates must
 *     round its blocks the same way.  Dense vectors and diag must START 16-byte aligned; vectors
 *     that are not are still accepted by the non-halo entry points (slower kernel), the *_halo
 *     entry points return GLAB_E_ARG.
 *   - floating point: element-wise chains use IEEE mul/add/div in the reference's operation
 *     order with FMA contraction disabled, and row sums accumulate sequentially in edge order
 *     (the order of scatter_add_), so results match the reference's CPU path bit for bit
 *     wherever the reference itself is order-deterministic.
 */
#ifndef GLAB_H_
#define GLAB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLAB_VERSION 200
#define GLAB_MAX_PEERS 8

#define GLAB_E_ARG      (-1)  /* null pointer / negative size / unsupported k                    */
#define GLAB_E_RANGE    (-2)  /* nnz >= 2^31, row/col index outside [0, n)                       */
#define GLAB_E_NOMEM    (-3)  /* plan allocation failed                                          */
#define GLAB_E_UNSORTED (-4)  /* (internal) input not row-sorted and sorting was disabled        */
#define GLAB_E_PEER     (-5)  /* peer (IPC) mapping failed                                       */

typedef struct glab_plan glab_plan;   /* opaque: int32 CSR structure of one operator on one GPU */
typedef struct glab_halo glab_halo;   /* opaque: peer-memory halo exchange of one row block     */

int glab_version(void);
const char* glab_error_string(int code);

/* ------------------------------------------------------------------------------------------
 * Plan: int64 COO (edgeij_pair[0] = aggregation target i, edgeij_pair[1] = source j) ->
 * int32 CSR + stable permutation.  Replaces the implicit grouping done by
 * scatter(index=edgeij_pair[0]) at every call site (MatVecGNN.py:60, GNNResidual.py:60,
 * JacobiGNN.py:66, ChebyGNN.py:87, PowerMethodGNN.py:80, SOCClassicGNN.py:69,
 * DirectInterpGNN.py:89,92) and the x[row]/x[col] gathers of MetaLayer.
 * Arbitrary edge order and duplicate (i,j) pairs are accepted; within a row the CSR keeps the
 * caller's edge order (stable), so duplicates are summed in the order scatter_add_ would.
 * Synchronises `stream` once (it must read back nnz statistics).  ncols = length of the
 * vectors that get gathered (n for a square operator; > n_rows for a row block with halo).
 * ------------------------------------------------------------------------------------------ */
int glab_plan_create(int64_t n_rows, int64_t n_cols, int64_t nnz,
                     const int64_t* row, const int64_t* col, void* stream, glab_plan** out);
/* Same, from an existing int32 CSR (device pointers are copied into the plan). */
int glab_plan_create_csr(int64_t n_rows, int64_t n_cols, int64_t nnz,
                         const int32_t* rowptr, const int32_t* colidx, void* stream,
                         glab_plan** out);
int glab_plan_destroy(glab_plan* plan);
/* Introspection.  perm == NULL means "identity" (input was already row-sorted). */
int glab_plan_info(const glab_plan* plan, int64_t* n_rows, int64_t* n_cols, int64_t* nnz,
                   int32_t* max_row_nnz, int32_t* identity_perm);
int glab_plan_csr(const glab_plan* plan, const int32_t** rowptr, const int32_t** colidx,
                  const int32_t** perm);
/* 16-bit row-relative column indices.  Besides int32 colidx a plan holds coldelta[slot] =
 * col - row as int16 for every 256-row tile whose rows all satisfy |col - row| <= 32767, and the
 * pipeline kernels stream those 2 bytes instead of the 4-byte index in such tiles: all tiles of a
 * banded operator (the 2-D stencils up to a 32767-wide grid line), all but the wrap-around tiles
 * of a periodic one, all but the tiles that read the halo tail of a row block.  Built at plan
 * creation; environment variable GLAB_IDX16 = 0 disables it, 1 keeps it only when every tile
 * qualifies, 2 (default) also when at least half of the tiles do, 3 additionally lets the fused
 * multi-GPU halo kernels use it (measured neutral, hence not the default).  Results never depend on it.
 * index_width: 2 when every tile streams 16-bit indices, else 4.  index16_tiles: how many of the
 * plan's 256-row tiles do. */
int glab_plan_index_width(const glab_plan* plan, int32_t* bytes);
int glab_plan_index16_tiles(const glab_plan* plan, int64_t* tiles16, int64_t* tiles_total);
/* L2 residency for operators that fit the 126 MB L2 (e.g. one rank's block of a row-partitioned
 * operator): adopt copies the CSR-ordered values into plan-owned storage directly behind colidx
 * (*vals_out points at the copy; use it instead of the caller's array), l2_persist then marks
 * [colidx .. values] as persisting in L2 for every kernel subsequently launched on (or captured
 * from) `stream` (cudaStreamAttributeAccessPolicyWindow; hit ratio scaled to the device's
 * persisting-L2 capacity), so repeated sweeps stream the operator from L2 instead of HBM.
 * enable = 0 removes the window.  Both are optional and change no result.  EXPERIMENTAL: on B200 the
 * measured effect for an 84 MB operator was negative, so no layer enables it. */
int glab_plan_adopt_vals_f32(glab_plan* plan, const float* vals, const float** vals_out, void* stream);
int glab_plan_adopt_vals_f64(glab_plan* plan, const double* vals, const double** vals_out, void* stream);
int glab_plan_l2_persist(const glab_plan* plan, int enable, void* stream);

/* vals[slot] = edge_attr[perm[slot]*ld + column]  (CSR-ordered contiguous copy of A_ij). */
int glab_gather_vals_f32(const glab_plan* plan, const float* edge_attr, int64_t ld, int64_t column,
                         float* vals, void* stream);
int glab_gather_vals_f64(const glab_plan* plan, const double* edge_attr, int64_t ld, int64_t column,
                         double* vals, void* stream);
/* out_edge_order[perm[slot]*ld + column] = in_slot_order[slot]  (inverse of the above). */
int glab_scatter_edges_f32(const glab_plan* plan, const float* in_slots, float* out_edges,
                           int64_t ld, int64_t column, void* stream);
int glab_scatter_edges_f64(const glab_plan* plan, const double* in_slots, double* out_edges,
                           int64_t ld, int64_t column, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused SpMV-bearing layer steps.  Common arguments: plan, vals[nnz] (CSR order), k columns,
 * [row_begin,row_end) = rows to process (pass 0, n_rows for all; used for interior/boundary
 * overlap with the halo exchange).  Gathered vectors have n_cols rows, the others n_rows.
 * ------------------------------------------------------------------------------------------ */

/* y = A x.  Replaces MatVecGNN.py:66-84 (c_ij = A_ij x_j), :43-62 (sum_j), :95-114. */
int glab_spmm_f32(const glab_plan*, const float* vals, const float* x, int k, float* y,
                  int64_t row_begin, int64_t row_end, void* stream);
int glab_spmm_f64(const glab_plan*, const double* vals, const double* x, int k, double* y,
                  int64_t row_begin, int64_t row_end, void* stream);

/* r = b - A x.  Replaces GNNResidual.py:64-86, :43-62, :88-118. */
int glab_residual_f32(const glab_plan*, const float* vals, const float* x, const float* b, int k,
                      float* r, int64_t row_begin, int64_t row_end, void* stream);
int glab_residual_f64(const glab_plan*, const double* vals, const double* x, const double* b, int k,
                      double* r, int64_t row_begin, int64_t row_end, void* stream);

/* out = b + A x  (may alias b).  The coarse-grid correction x + P xc of VCycle.py:226. */
int glab_spmm_add_f32(const glab_plan*, const float* vals, const float* x, const float* b, int k,
                      float* out, int64_t row_begin, int64_t row_end, void* stream);
int glab_spmm_add_f64(const glab_plan*, const double* vals, const double* x, const double* b, int k,
                      double* out, int64_t row_begin, int64_t row_end, void* stream);

/* x_out = x_in + (omega*(b - A x_in))/diag  (one weighted-Jacobi sweep; x_out != x_in).
 * Replaces JacobiGNN.py:71-89, :52-69, :91-123 (op order of :119).  diag is [n_rows] (shared by
 * all k columns), omega_dev points to ONE device scalar (g[0], :112). */
int glab_jacobi_f32(const glab_plan*, const float* vals, const float* diag, const float* b,
                    const float* x_in, float* x_out, const float* omega_dev, int k,
                    int64_t row_begin, int64_t row_end, void* stream);
int glab_jacobi_f64(const glab_plan*, const double* vals, const double* diag, const double* b,
                    const double* x_in, double* x_out, const double* omega_dev, int k,
                    int64_t row_begin, int64_t row_end, void* stream);

/* n_sweeps weighted-Jacobi sweeps in ONE launch (the Python loop `for i in range(n_iters)` of
 * JacobiGNN.py:143-144 around the block above).  x ping-pongs xa -> xb -> xa ...: xa holds the start
 * vector, the result is in xb when n_sweeps is odd, in xa when it is even; both buffers are
 * overwritten.  A persistent, cooperatively launched kernel keeps its TMA ring running across the
 * sweep boundaries; tile t of sweep s starts once the tiles within the operator's band around t
 * have finished sweep s - 1 (per-tile completion counters inside the plan), so there is no grid
 * barrier and no pipeline drain between sweeps.  Same arithmetic, bit for bit, as n_sweeps calls of
 * glab_jacobi_*; operators that do not fit the pipeline (or devices without cooperative launch, or
 * GLAB_MS=0) are run as exactly those calls.  One multi-sweep launch at a time per plan. */
int glab_jacobi_sweeps_f32(const glab_plan*, const float* vals, const float* diag, const float* b,
                           float* xa, float* xb, const float* omega_dev, int k, int n_sweeps,
                           void* stream);
int glab_jacobi_sweeps_f64(const glab_plan*, const double* vals, const double* diag, const double* b,
                           double* xa, double* xb, const double* omega_dev, int k, int n_sweeps,
                           void* stream);

/* Chebyshev iteration 1:  r = b - A x_in;  p = r;  x_out = x_in + alpha*p.
 * Replaces ChebyGNN.py:49-70, :73-89, :91-121, :141-163.  alpha_dev -> 1/d (:137). */
int glab_cheby_first_f32(const glab_plan*, const float* vals, const float* b, const float* x_in,
                         float* x_out, float* r, float* p, const float* alpha_dev, int k,
                         int64_t row_begin, int64_t row_end, void* stream);
int glab_cheby_first_f64(const glab_plan*, const double* vals, const double* b, const double* x_in,
                         double* x_out, double* r, double* p, const double* alpha_dev, int k,
                         int64_t row_begin, int64_t row_end, void* stream);

/* Chebyshev iteration > 1:  r -= alpha_old*(A p_in);  p_out = r + beta*p_in;  x += alpha*p_out
 * (r and x updated in place, p ping-pongs).  Replaces ChebyGNN.py:166-183, :186-216, :219-243;
 * alpha_old is the previous iteration's alpha because MetaLayer runs the vertex update before
 * the global update (:214 vs :262-263). */
int glab_cheby_next_f32(const glab_plan*, const float* vals, const float* p_in, float* p_out,
                        float* r, float* x, const float* alpha_old_dev, const float* alpha_dev,
                        const float* beta_dev, int k, int64_t row_begin, int64_t row_end,
                        void* stream);
int glab_cheby_next_f64(const glab_plan*, const double* vals, const double* p_in, double* p_out,
                        double* r, double* x, const double* alpha_old_dev, const double* alpha_dev,
                        const double* beta_dev, int k, int64_t row_begin, int64_t row_end,
                        void* stream);

/* Power-method step with deferred normalisation:
 *     y = (A b_in) / n,   n = sqrt(sumsq_in[0]) if sumsq_in != NULL, else 1;
 *     sumsq_out[0] = sum_i y_i^2   (fp64 accumulation, deterministic two-level reduction).
 * b_in is the previous step's UN-normalised y and sumsq_in its squared norm, so
 * (A b_in)/n == A (b_in/n): the reference's b <- b/n (PowerMethodGNN.py:187-207) is folded into
 * the next step and never costs a pass over the vector.  Replaces PowerMethodGNN.py:86-106,
 * :64-83, :129-158 (b <- A b), :109-126 (y = b^2), :160-185 (n = sqrt(sum y)).
 * `workspace` = device scratch of glab_reduce_workspace_bytes() bytes, zero-initialised once
 * by the caller and then reusable by any number of stream-ordered calls. */
int64_t glab_reduce_workspace_bytes(void);
int glab_power_step_f32(const glab_plan*, const float* vals, const float* b_in, float* y,
                        const double* sumsq_in, double* sumsq_out, void* workspace,
                        int64_t row_begin, int64_t row_end, void* stream);
int glab_power_step_f64(const glab_plan*, const double* vals, const double* b_in, double* y,
                        const double* sumsq_in, double* sumsq_out, void* workspace,
                        int64_t row_begin, int64_t row_end, void* stream);

/* Rayleigh step:  bn = b_in / n (n as above);  Ab = (A b_in) / n;  y_out = bn*bn;
 *     sums_out[0] = sum bn_i*Ab_i   (n_A, PowerMethodGNN.py:235,:264)
 *     sums_out[1] = sum bn_i^2      (:124,:292)
 * b_out receives bn (the normalised iterate the reference returns in vertex_attr[:,0]).
 * Replaces PowerMethodGNN.py:209-237, :239-266, :268-294. */
int glab_rayleigh_f32(const glab_plan*, const float* vals, const float* b_in, float* b_out,
                      float* y_out, const double* sumsq_in, double* sums_out, void* workspace,
                      int64_t row_begin, int64_t row_end, void* stream);
int glab_rayleigh_f64(const glab_plan*, const double* vals, const double* b_in, double* b_out,
                      double* y_out, const double* sumsq_in, double* sums_out, void* workspace,
                      int64_t row_begin, int64_t row_end, void* stream);

/* Matrix-weighted norm pieces: sums_out[0] = sum_i x_i*(W x)_i.  MatrixWeightedNorm.py:49-161. */
int glab_xtax_f32(const glab_plan*, const float* vals, const float* x, double* sums_out,
                  void* workspace, int64_t row_begin, int64_t row_end, void* stream);
int glab_xtax_f64(const glab_plan*, const double* vals, const double* x, double* sums_out,
                  void* workspace, int64_t row_begin, int64_t row_end, void* stream);

/* c_slot = A_slot * x[col(slot)] for every edge (the `c_ij` / `z_ij` column the reference
 * layers return in edge_attr, e.g. MatVecGNN.py:80-84).  Written straight to the caller's
 * [nnz, ld] edge-order array at column `column` (k consecutive columns for k-column x). */
int glab_edge_messages_f32(const glab_plan*, const float* vals, const float* x, int k,
                           float* out_edges, int64_t ld, int64_t column, void* stream);
int glab_edge_messages_f64(const glab_plan*, const double* vals, const double* x, int k,
                           double* out_edges, int64_t ld, int64_t column, void* stream);

/* The reference's seam itself, for callers that compose their own graph-network blocks:
 *   segment_sum: out[i, :] = sum over the edges e with edgeij_pair[0][e] == i of src[e, :]
 *                == torch_scatter.scatter(src, edgeij_pair[0], dim=0, dim_size=n, reduce="sum")
 *                (MatVecGNN.py:60 and the other call sites), sequential in edge order;
 *   segment_max: reduce="max" with 0 for rows without edges (SOCClassicGNN.py:69).
 * src is [nnz, k] (k in {1,2,4,8}) contiguous in CSR slot order (use glab_gather_vals_* first
 * when the plan's permutation is not the identity). */
int glab_segment_sum_f32(const glab_plan*, const float* src_slots, int k, float* out, void* stream);
int glab_segment_sum_f64(const glab_plan*, const double* src_slots, int k, double* out, void* stream);
int glab_segment_max_f32(const glab_plan*, const float* src_slots, float* out, void* stream);
int glab_segment_max_f64(const glab_plan*, const double* src_slots, double* out, void* stream);
/* reduce="min" / reduce="mean" (sum / max(count, 1)), the other two members of the 4-way
 * aggregation of TrainableJacobiDiag/TrainableJacobiGNN.py:65-68; empty rows give 0. */
int glab_segment_min_f32(const glab_plan*, const float* src_slots, float* out, void* stream);
int glab_segment_min_f64(const glab_plan*, const double* src_slots, double* out, void* stream);
int glab_segment_mean_f32(const glab_plan*, const float* src_slots, int k, float* out, void* stream);
int glab_segment_mean_f64(const glab_plan*, const double* src_slots, int k, double* out, void* stream);

/* The 4-way aggregation in ONE pass: out[i, :] = [min | mean | sum | max] (F columns each, out is
 * [n_rows, 4 F]) over the slots of row i of src [nnz, F] (F <= 64, CSR slot order) -- what
 * TrainableJacobiDiag/TrainableJacobiGNN.py:53-70 and DiffCoeffs/LearnDiffusionCoeffs.py:291-342
 * compute with four torch_scatter.scatter calls.  The plan's rows are whatever the index groups by:
 * vertices (edge -> vertex aggregation) or the graphs of a batch (the reference's `batch` vector:
 * vertex / edge -> graph aggregation; long segments).  Rows of at most 64 slots are reduced
 * sequentially in edge order (sum / mean bit-identical to scatter_add_); longer rows by one warp each
 * (min / max exact, sum / mean in a fixed but different order).  Empty rows give 0. */
int glab_segment_agg4_f32(const glab_plan*, const float* src_slots, int F, float* out, void* stream);
int glab_segment_agg4_f64(const glab_plan*, const double* src_slots, int F, double* out, void* stream);

/* out_edges[e, 0] = A_ij and out_edges[e, 1..k] = A_ij * x_j in ONE pass: the reference layers'
 * returned edge_attr = torch.cat([A_ij, c_ij], 1) (MatVecGNN.py:84, JacobiGNN.py:88,
 * ChebyGNN.py:70,183, PowerMethodGNN.py:106).  out_edges is [nnz, ld], ld >= 1 + k. */
int glab_edge_attr_f32(const glab_plan*, const float* vals, const float* x, int k, float* out_edges,
                       int64_t ld, void* stream);
int glab_edge_attr_f64(const glab_plan*, const double* vals, const double* x, int k, double* out_edges,
                       int64_t ld, void* stream);

/* Layout glue of the drop-in API: the reference's layers take and return INTERLEAVED attribute
 * tensors (vertex_attr [n, F] = cat([A_ii, b, x], 1) ..., JacobiGNN.py:121, ChebyGNN.py:163,243,
 * PowerMethodGNN.py:126), the kernels work on dense column blocks.
 *   pack:   dst[i, off_j .. off_j + w_j) = src_j[i, 0..w_j)   for j < n_parts (dst is [n, ld])
 *   unpack: dst_j[i, 0..w_j) = src[i, off_j .. off_j + w_j)   (src is [n, ld])
 * parts / widths / offsets are HOST arrays (n_parts <= 8), copied at launch. */
#define GLAB_MAX_PARTS 8
int glab_pack_f32(int64_t n, int64_t ld, int n_parts, const float* const* parts, const int32_t* widths,
                  const int32_t* offsets, float* dst, void* stream);
int glab_pack_f64(int64_t n, int64_t ld, int n_parts, const double* const* parts, const int32_t* widths,
                  const int32_t* offsets, double* dst, void* stream);
int glab_unpack_f32(int64_t n, int64_t ld, int n_parts, float* const* parts, const int32_t* widths,
                    const int32_t* offsets, const float* src, void* stream);
int glab_unpack_f64(int64_t n, int64_t ld, int n_parts, double* const* parts, const int32_t* widths,
                    const int32_t* offsets, const double* src, void* stream);

/* ------------------------------------------------------------------------------------------
 * AMG setup kernels (row-local, one pass).  Per-edge outputs are written in the caller's
 * ORIGINAL edge order (out[perm[slot]]).
 * ------------------------------------------------------------------------------------------ */

/* Classical strength of connection on off-diagonal edges:
 *     v_i = max_k(-A_ik) (0 for empty rows);  S = relu(((-1*A_ij)/v_i) - theta).
 * Replaces SOCClassicGNN.py:50-72, :74-102, :104-129.  rowmax_out may be NULL. */
int glab_soc_classic_f32(const glab_plan*, const float* vals, float theta, float* S_edges,
                         float* rowmax_out, void* stream);
int glab_soc_classic_f64(const glab_plan*, const double* vals, double theta, double* S_edges,
                         double* rowmax_out, void* stream);

/* Smoothed-aggregation strength measure S_ij = (A_ij*A_ij)/(A_ii*A_jj); no threshold.
 * Replaces SOCSAGNN.py:49-71.  diag has n_cols entries. */
int glab_soc_sa_f32(const glab_plan*, const float* vals, const float* diag, float* S_edges,
                    void* stream);
int glab_soc_sa_f64(const glab_plan*, const double* vals, const double* diag, double* S_edges,
                    void* stream);

/* Direct interpolation weights:
 *     num_i = sum_k A_ik;  den_i = sum_k (A_ik*S_ik)*C_k;  alpha_i = (1/A_ii)*(num_i/den_i);
 *     w_ij = (1 - C_i)*((-A_ij)*alpha_i)        for EVERY edge (NaN = 0*inf kept, see header).
 * Replaces DirectInterpGNN.py:50-69, :71-97, :99-131, :133-152.  S[nnz] in CSR slot order,
 * Cflag has n_cols entries (1 = coarse), diag n_rows entries. */
int glab_direct_interp_f32(const glab_plan*, const float* vals, const float* S, const float* diag,
                           const float* Cflag, float* w_edges, void* stream);
int glab_direct_interp_f64(const glab_plan*, const double* vals, const double* S,
                           const double* diag, const double* Cflag, double* w_edges, void* stream);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU halo exchange over NVLink peer memory (one process per GPU).  Each rank owns a
 * contiguous row block; gathered vectors are [n_local + n_halo, k] with the halo tail filled
 * from the owners.  The exchange is a PUSH: a pack kernel stores this rank's boundary values
 * straight into the peers' halo tails through mapped peer pointers, followed by a release flag;
 * consumers wait on the flag in-stream (no NCCL call on the data path).
 * ------------------------------------------------------------------------------------------ */
int glab_ipc_handle_bytes(void);
/* Allocate a device buffer that peers can map; writes an opaque handle (glab_ipc_handle_bytes()). */
int glab_ipc_alloc(int64_t bytes, void** dev_ptr, void* handle_out);
int glab_ipc_open(const void* handle, void** peer_ptr);
int glab_ipc_close(void* peer_ptr);
int glab_ipc_free(void* dev_ptr);
/* One push = one kernel launch for ALL neighbours (gridDim.y = peer):
 *   for every peer q < n_peers:  dst[q][dst_offset[q] + i, :] = src[send_idx[q][i], :], i < count[q]
 * (k columns), followed by a system-scope release increment (+1) of *flag[q], a uint32 arrival
 * counter in peer q's memory, and by ++*pushed_local (this rank's own count of pushes of this
 * vector, device resident).  Counters only ever count up and live on the device, so a CUDA
 * graph that contains pushes and waits can be replayed any number of times.
 * `descs` is a HOST array, copied at launch.  n_peers == 0 still bumps *pushed_local. */
typedef struct glab_push_desc {
  const int32_t* send_idx;   /* device: local row ids to send to this peer        */
  int64_t first_row;         /* >= 0: send_idx[i] == first_row + i for all i (contiguous block:
                                copied with 16-byte vectors, no index loads); -1: use send_idx */
  int64_t count;
  void* dst;                 /* peer-mapped base of the destination vector        */
  int64_t dst_offset;        /* first destination row (the peer's halo tail slot) */
  uint32_t* flag;            /* peer-mapped arrival counter, may be NULL          */
} glab_push_desc;
int glab_halo_push_f32(const float* src, int k, int n_peers, const glab_push_desc* descs,
                       uint32_t* pushed_local, void* stream);
int glab_halo_push_f64(const double* src, int k, int n_peers, const glab_push_desc* descs,
                       uint32_t* pushed_local, void* stream);
/* One wait = one single-CTA kernel: spin, with system-scope acquire loads, until every arrival
 * counter *flags[i] (i < n_flags) has reached *pushed_local -- i.e. until each neighbour has
 * pushed this vector as often as this rank has (all ranks run the same sequence of pushes).
 * flags is a HOST array of device pointers. */
int glab_halo_wait(int n_flags, uint32_t* const* flags, const uint32_t* pushed_local, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused step + halo exchange: ONE kernel per sweep on a row-partitioned operator.
 * The step processes the rows of [interior_begin, interior_end) first (they gather only local
 * values), then the remaining boundary rows, whose producer acquires the arrival counters
 * wait_flags[i] (until each has reached *wait_target -- the neighbours' pushes of the GATHERED
 * vector); when the whole grid has finished, its last CTA stores this rank's boundary rows of
 * `push_src` (the vector the step PRODUCED and the next step will gather: x_out for Jacobi,
 * p_out for Chebyshev, y for the power step) into the neighbours' halo tails through the
 * peer-mapped pointers of `push`, release-increments their arrival counters and ++*pushed_counter.
 * interior_begin / interior_end must be multiples of 256.  All arrays are HOST arrays copied at
 * launch; done_counter is a zero-initialised device word owned by the caller.
 *
 * Residency: the CTAs of a fused step wait on each other (boundary tiles on the communication CTA's
 * counters), so the whole grid (SMs x occupancy) must be resident at once.  On a GPU the process owns
 * that holds by construction.  On a shared GPU (MPS, other streams, green contexts) set
 * GLAB_HALO_COOP=1: the steps are then launched with cudaLaunchAttributeCooperative -- the driver
 * starts the grid only when all of it fits -- at the price of the programmatic dependent launch
 * between consecutive steps.  Either way every in-kernel wait is bounded (`timeout_ms`): a step that
 * gives up sets GLAB_STATUS_TIMEOUT_* bits in *status instead of hanging, and its results are
 * undefined.  The multi-sweep entry points (glab_jacobi_sweeps_*) are always launched cooperatively.
 * ------------------------------------------------------------------------------------------ */
/* 1 if the fused halo steps (glab_*_halo_*) can take this operator with k right-hand-side columns of
 * `elem_bytes`-wide values: a 256-row tile of max_row_nnz slots plus the vertex streams must fit two
 * shared-memory stages.  0: use the separate kernels (glab_halo_wait, the row-range launches,
 * glab_halo_push_*) -- the fused entry points return GLAB_E_ARG for such operators (coarse-level
 * operators with very wide rows). */
int glab_halo_fits(const glab_plan* plan, int k, int elem_bytes);

/* Sum of per-rank partial sums over peer memory, fused into the reducing step kernels (power method):
 * every rank owns a mailbox  double mail[2][GLAB_MAX_PEERS][2]  and arrival counters
 * uint32 flag[GLAB_MAX_PEERS] (16 bytes apart), both peer-mapped.  The grid's last CTA stores this rank's
 * two partial sums into slot [parity][rank] of EVERY rank's mailbox (its own included) and
 * release-increments flag[rank] there; parity alternates per publishing launch (*parity_counter, a
 * device word the kernel advances).  A launch that CONSUMES sums (sumsq_in != NULL in
 * glab_power_step_halo_* / glab_rayleigh_halo_*) waits until all `world` counters have reached this
 * rank's own count, then adds the `world` partials of the previous parity in rank order -- the same
 * order on every rank, so all ranks use bit-identical norms.  Replaces one NCCL all-reduce per
 * iteration. */
typedef struct glab_peer_reduce {
  int32_t world, rank;
  double* mail_local;                    /* this rank's mailbox                                   */
  uint32_t* flag_local;                  /* this rank's arrival counters                          */
  double* mail_peer[GLAB_MAX_PEERS];     /* every rank's mailbox (peer-mapped; [rank] = local)    */
  uint32_t* flag_peer[GLAB_MAX_PEERS];   /* every rank's counters (peer-mapped)                   */
  uint32_t* parity_counter;              /* device word: number of publishes so far (local)       */
} glab_peer_reduce;

typedef struct glab_halo_step {
  int64_t interior_begin, interior_end;
  int32_t n_wait;
  uint32_t* const* wait_flags;
  const uint32_t* wait_target;
  int32_t n_push;
  const glab_push_desc* push;
  uint32_t* pushed_counter;
  const void* push_src;
  uint32_t* done_counter;
  uint32_t* status;        /* device word (may be NULL): OR-ed with GLAB_STATUS_* when an in-kernel
                              wait gave up after timeout_ms; the results of that launch are undefined */
  int64_t timeout_ms;      /* bound of every in-kernel wait on a flag written by another CTA / GPU;
                              0 = the library default (GLAB_SPIN_TIMEOUT_MS, else 20000), < 0 = forever */
  const glab_peer_reduce* reduce;   /* power_step / rayleigh only, may be NULL: sum the partial sums over
                              the ranks inside the kernels (see glab_peer_reduce) instead of leaving
                              sumsq_out rank-local                                                  */
} glab_halo_step;
#define GLAB_STATUS_TIMEOUT_PEER   1u  /* a neighbour's arrival counter did not reach its target   */
#define GLAB_STATUS_TIMEOUT_TILES  2u  /* this GPU's own boundary tiles did not finish             */
#define GLAB_STATUS_TIMEOUT_SWEEP  4u  /* multi-sweep kernel: a tile of the previous sweep did not */

int glab_spmm_halo_f32(const glab_plan*, const float* vals, const float* x, int k, float* y,
                       const glab_halo_step* halo, void* stream);
int glab_spmm_halo_f64(const glab_plan*, const double* vals, const double* x, int k, double* y,
                       const glab_halo_step* halo, void* stream);
int glab_residual_halo_f32(const glab_plan*, const float* vals, const float* x, const float* b, int k,
                           float* r, const glab_halo_step* halo, void* stream);
int glab_residual_halo_f64(const glab_plan*, const double* vals, const double* x, const double* b, int k,
                           double* r, const glab_halo_step* halo, void* stream);
int glab_jacobi_halo_f32(const glab_plan*, const float* vals, const float* diag, const float* b,
                         const float* x_in, float* x_out, const float* omega_dev, int k,
                         const glab_halo_step* halo, void* stream);
int glab_jacobi_halo_f64(const glab_plan*, const double* vals, const double* diag, const double* b,
                         const double* x_in, double* x_out, const double* omega_dev, int k,
                         const glab_halo_step* halo, void* stream);
/* Multi-sweep Jacobi on a row block: step_ab describes the sweeps that gather xa and produce xb (wait
 * on xa's arrival counters, push xb), step_ba the others.  Both buffers must be peer-mapped gathered
 * vectors [n_local + n_halo, k]; done_counter must point at 32 bytes of zero-initialised device
 * memory (two counters, 16 bytes apart). */
int glab_jacobi_sweeps_halo_f32(const glab_plan*, const float* vals, const float* diag, const float* b,
                                float* xa, float* xb, const float* omega_dev, int k, int n_sweeps,
                                const glab_halo_step* step_ab, const glab_halo_step* step_ba, void* stream);
int glab_jacobi_sweeps_halo_f64(const glab_plan*, const double* vals, const double* diag, const double* b,
                                double* xa, double* xb, const double* omega_dev, int k, int n_sweeps,
                                const glab_halo_step* step_ab, const glab_halo_step* step_ba, void* stream);
int glab_cheby_first_halo_f32(const glab_plan*, const float* vals, const float* b, const float* x_in,
                              float* x_out, float* r, float* p, const float* alpha_dev, int k,
                              const glab_halo_step* halo, void* stream);
int glab_cheby_first_halo_f64(const glab_plan*, const double* vals, const double* b, const double* x_in,
                              double* x_out, double* r, double* p, const double* alpha_dev, int k,
                              const glab_halo_step* halo, void* stream);
int glab_cheby_next_halo_f32(const glab_plan*, const float* vals, const float* p_in, float* p_out,
                             float* r, float* x, const float* alpha_old_dev, const float* alpha_dev,
                             const float* beta_dev, int k, const glab_halo_step* halo, void* stream);
int glab_cheby_next_halo_f64(const glab_plan*, const double* vals, const double* p_in, double* p_out,
                             double* r, double* x, const double* alpha_old_dev, const double* alpha_dev,
                             const double* beta_dev, int k, const glab_halo_step* halo, void* stream);
int glab_power_step_halo_f32(const glab_plan*, const float* vals, const float* b_in, float* y,
                             const double* sumsq_in, double* sumsq_out, void* workspace,
                             const glab_halo_step* halo, void* stream);
int glab_power_step_halo_f64(const glab_plan*, const double* vals, const double* b_in, double* y,
                             const double* sumsq_in, double* sumsq_out, void* workspace,
                             const glab_halo_step* halo, void* stream);
int glab_rayleigh_halo_f32(const glab_plan*, const float* vals, const float* b_in, float* b_out,
                           float* y_out, const double* sumsq_in, double* sums_out, void* workspace,
                           const glab_halo_step* halo, void* stream);
int glab_rayleigh_halo_f64(const glab_plan*, const double* vals, const double* b_in, double* b_out,
                           double* y_out, const double* sumsq_in, double* sums_out, void* workspace,
                           const glab_halo_step* halo, void* stream);

/* ------------------------------------------------------------------------------------------
 * AMG setup on the device: the glue either side of SOCClassicGNN /
 * DirectInterpGNN in the reference's two-grid cycle.  Setup work, once per operator; the calls
 * marked "host-synchronous" synchronise `stream` because they return a size to the host.
 * ------------------------------------------------------------------------------------------ */

/* Prolongator P = [I + W](:, C) as a sparse, (row, col)-sorted int64 COO.  Replaces
 * VCycle.py:126-137 (torch.eye(n) + W, .to_dense(), column slice, .to_sparse()), which builds a
 * dense n x n matrix.  `A_off` is the plan of the off-diagonal edges (the edge set DirectInterpGNN
 * ran on, UtilsGNN.py:69-72) with ascending columns inside each row (any coalesced operator),
 * w_slots[nnz] the DirectInterpGNN output in CSR slot order, cflag[n] the C/F splitting
 * (> 0 = coarse).  An entry (i, j) of W is kept when j is coarse and w_ij is not an exact zero
 * (NaN is kept), exactly what .to_sparse() keeps.  mode 0 = the Python reference as shipped
 * (a coarse row keeps its W entries, which are NaN when it has no strong coarse neighbour,
 * DirectInterpGNN.py:150); mode 1 = the MATLAB twin (coarse rows are identity rows,
 * matlab/test_direct_interpolation.m:130-132).
 *   workspace_bytes: size of the caller-owned scratch for count (scan state, < 1 MB); negative =
 *     error code.
 *   count (host-synchronous): p_rowptr[n+1] = row offsets of P, coarse_id[n+1] = coarse column
 *     number of every vertex (exclusive count of coarse points before it); returns nnz(P), n_coarse.
 *   fill: writes the nnz(P) entries; row i holds its kept W entries in ascending column order with
 *     the identity entry (i, coarse_id[i]) = 1 merged in.  nnz(P) must be < 2^31. */
int64_t glab_interp_workspace_bytes(int64_t n);
int glab_interp_count_f32(const glab_plan* A_off, const float* w_slots, const float* cflag, int mode,
                          void* workspace, int64_t workspace_bytes, int32_t* p_rowptr,
                          int32_t* coarse_id, int64_t* nnz_p, int64_t* n_coarse, void* stream);
int glab_interp_count_f64(const glab_plan* A_off, const double* w_slots, const double* cflag, int mode,
                          void* workspace, int64_t workspace_bytes, int32_t* p_rowptr,
                          int32_t* coarse_id, int64_t* nnz_p, int64_t* n_coarse, void* stream);
int glab_interp_fill_f32(const glab_plan* A_off, const float* w_slots, const float* cflag, int mode,
                         const int32_t* p_rowptr, const int32_t* coarse_id, int64_t* out_row,
                         int64_t* out_col, float* out_val, void* stream);
int glab_interp_fill_f64(const glab_plan* A_off, const double* w_slots, const double* cflag, int mode,
                         const int32_t* p_rowptr, const int32_t* coarse_id, int64_t* out_row,
                         int64_t* out_col, double* out_val, void* stream);

/* Sparse product Z = X * Y of two plans; called twice for the Galerkin operator A_c = P^T (A P).
 * Replaces VCycle.py:209 (`P.t() @ (A @ P)` on torch.sparse tensors).  The output is a
 * (row, col)-sorted int64 COO without duplicates -- the reference's edge layout
 * (UtilsGNN.py:74-78), so it can be handed to glab_plan_create and to every layer as the next
 * operator.  Products that meet in one entry are added sequentially in expansion order (X slot
 * order, then Y slot order): deterministic and independent of the launch geometry and of the
 * path taken.  Two paths: "row-local" (ONE pass during the symbolic call: a CTA expands the products
 * of 128 consecutive rows into shared memory and every thread insertion-sorts its row's segment in
 * place -- or, for rows too wide for that, one thread per row keeps the row's distinct columns and
 * sums in a shared-memory strip of 16 / 32 / 64 entries; the finished rows are parked in the
 * workspace and the numeric call compacts them; taken when no row has more than 64 distinct columns
 * or 128 rows' products fit 64 KB, i.e. for every stencil / Galerkin operator) and "ESC" (expand - stable radix sort - compress through the
 * workspace; the fallback for dense-ish rows).
 *   products (host-synchronous): number of scalar multiplications sum_{(i,j) in X} nnz(Y_j*) and
 *     the largest such count of a single row; scratch16 = 16 bytes of device memory.
 *     n_products must be < 2^31 - 64 (GLAB_E_RANGE otherwise).
 *   workspace_bytes: size of the caller-owned device workspace (elem_size 4 or 8); negative =
 *     error code.  12 bytes per X row + (4 + elem_size) bytes per product when
 *     max_row_products <= 64 (row-local path guaranteed), else about (16 + 2 * elem_size) bytes
 *     per product.
 *   symbolic (host-synchronous): returns nnz(Z) and leaves the parked rows / sorted runs in the
 *     workspace.  It reads the VALUES: the row-local path finishes the arithmetic here.
 *   numeric (host-synchronous): writes the nnz(Z) entries; MUST be given the same X, Y, values,
 *     workspace and counts as the symbolic call (a new set of values needs a new symbolic call). */
int glab_spgemm_products(const glab_plan* X, const glab_plan* Y, void* scratch16, int64_t* n_products,
                         int64_t* max_row_products, void* stream);
int64_t glab_spgemm_workspace_bytes(int64_t n_rows_x, int64_t n_products, int64_t max_row_products,
                                    int elem_size);
int glab_spgemm_symbolic_f32(const glab_plan* X, const float* x_vals, const glab_plan* Y,
                             const float* y_vals, void* workspace, int64_t workspace_bytes,
                             int64_t n_products, int64_t max_row_products, int64_t* nnz_out,
                             void* stream);
int glab_spgemm_symbolic_f64(const glab_plan* X, const double* x_vals, const glab_plan* Y,
                             const double* y_vals, void* workspace, int64_t workspace_bytes,
                             int64_t n_products, int64_t max_row_products, int64_t* nnz_out,
                             void* stream);
int glab_spgemm_numeric_f32(const glab_plan* X, const float* x_vals, const glab_plan* Y,
                            const float* y_vals, void* workspace, int64_t workspace_bytes,
                            int64_t n_products, int64_t max_row_products, int64_t nnz_out,
                            int64_t* out_row, int64_t* out_col, float* out_val, void* stream);
int glab_spgemm_numeric_f64(const glab_plan* X, const double* x_vals, const glab_plan* Y,
                            const double* y_vals, void* workspace, int64_t workspace_bytes,
                            int64_t n_products, int64_t max_row_products, int64_t nnz_out,
                            int64_t* out_row, int64_t* out_col, double* out_val, void* stream);

/* Coarse/fine splitting on the strength graph (PMIS: parallel modified independent set).  Stands
 * in for the pyamg CLJP call of the reference (VCycle.py:114, DirectInterpGNN.py:194; un-pinned
 * third-party package, so there is no reference output to match -- parity is against
 * oracle/cf_split.py, bit for bit).  `A_off` = plan of the off-diagonal edges, S_slots[nnz] = the
 * SOCClassicGNN output in slot order (strong <=> S > 0, VCycle.py:90).  Vertex keys are integers
 * ((number of rows that strongly depend on i) + 1) << 32 | mix32(i + seed), all distinct; each
 * round the undecided vertices whose key beats every undecided strong neighbour (either direction)
 * become coarse, then undecided rows that strongly depend on a coarse point become fine.  Every
 * fine point ends with at least one strong coarse neighbour; vertices without strong connections
 * become coarse.  cflag_out[n] = 1 (coarse) / 0 (fine) in the operator's dtype, the layout of
 * vertex_attr[:,1] of DirectInterpGNN.  Host-synchronous (one read-back per round). */
int64_t glab_cf_split_workspace_bytes(int64_t n);
int glab_cf_split_pmis_f32(const glab_plan* A_off, const float* S_slots, uint32_t seed, void* workspace,
                           int64_t workspace_bytes, float* cflag_out, int32_t* rounds_out, void* stream);
int glab_cf_split_pmis_f64(const glab_plan* A_off, const double* S_slots, uint32_t seed, void* workspace,
                           int64_t workspace_bytes, double* cflag_out, int32_t* rounds_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GLAB_H_ */
