/*
 * fdb200.h -- C ABI of libfdb200.so: the B200 (sm_100a) implementation of FlashDeconv's
 * data-parallel hot path.
 *
 * The reference (cafferychen777/flashdeconv v0.1.6) is pure Python and has NO FFI layer;
 * each entry point below names the reference function (file:line under the upstream repo)
 * whose arithmetic it replaces.  A maintainer binds them with ctypes exactly as
 * flashdeconv_b200/_native.py does.
 *
 * Conventions
 *   - every function returns 0 on success, a negative fdb_status otherwise; the message is
 *     available from fdb_last_error() (thread-local);
 *   - all pointers are DEVICE pointers unless the parameter name starts with `host_`;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on it,
 *     functions documented "syncs" block the host on that stream once;
 *   - nothing is allocated behind the caller's back: scratch memory is passed in as
 *     (workspace, workspace_bytes) sized by the matching *_workspace_bytes() query;
 *   - no torch / C++ types cross the boundary.
 *
 * Device layouts
 *   spots live in "tile order" (a spatially coherent permutation produced by
 *   fdb_graph_build); `order[p]` = original index of the spot at position p, `rank[i]` =
 *   position of original spot i.  Spot-by-type matrices (H, beta) are row-major
 *   n_rows x Kp float32 with Kp = fdb_padded_types(K) (K rounded up to a multiple of 8, so fp32
 *   rows are 32-byte and their fp16 images 16-byte aligned); padding columns are zero.
 */
#ifndef FDB200_H
#define FDB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    FDB_OK = 0,
    FDB_ERR_ARG = -1,        /* invalid argument                       */
    FDB_ERR_CUDA = -2,       /* a CUDA runtime call or launch failed   */
    FDB_ERR_WORKSPACE = -3,  /* workspace too small                    */
    FDB_ERR_UNSUPPORTED = -4 /* valid request outside what is built    */
} fdb_status;

#define FDB_MAX_TYPES 64     /* K handled by the register-resident BCD kernels and the fused sketch */
#define FDB_MAX_TYPES_WIDE 1024   /* K handled by the warp-per-spot entry points (fdb_*_wide, fdb_contract, fdb_finish) */

int fdb_abi_version(void);
const char *fdb_last_error(void);
int fdb_padded_types(int n_types);                       /* Kp */
long long fdb_launch_count(void);                        /* kernels launched by this library so far */

/* ---------------------------------------------------------------------------------------
 * (a1 + a3) log-CPM fused with the leverage-weighted CountSketch, one pass over the FULL
 * spot-by-gene CSR.  Replaces FlashDeconv._preprocess_data "log_cpm" (core/deconv.py:177-197)
 * + Y[:, gene_idx] (core/deconv.py:321) + project_to_sketch (core/sketching.py:190-199).
 *   gene_bucket[g] in [0,d) for selected genes, -1 otherwise;  gene_weight[g] = Omega entry.
 *   lib_i = sum of counts over SELECTED genes (0 -> 1), y~ = log1p(1e4 * count / lib_i).
 *   y_sketch: n_spots x d float32, row i at y_sketch + i*d (input spot order).
 * ------------------------------------------------------------------------------------- */
int fdb_sketch_logcpm_csr(const void *indptr, int indptr_is_int64, const int32_t *indices,
                          const float *counts, int64_t n_spots, int32_t n_genes,
                          const int32_t *gene_bucket, const float *gene_weight, int32_t d,
                          float *y_sketch, void *stream);

/* (a3 alone) Y_s = Y~ Omega for an ALREADY transformed CSR matrix -- the arithmetic of
 * project_to_sketch (core/sketching.py:160-206) without the log-CPM step. */
int fdb_sketch_project_csr(const void *indptr, int indptr_is_int64, const int32_t *indices,
                           const float *values, int64_t n_spots, int32_t n_genes,
                           const int32_t *gene_bucket, const float *gene_weight, int32_t d,
                           float *y_sketch, void *stream);

/* (a4) H = Y_s X_s^T and per-spot ||y_s||^2.  Replaces precompute_XtY (core/solver.py:204-223)
 * and YtY (core/solver.py:348).  x_sketch: K x d row-major.  h: n_spots x Kp.  ysq: n_spots f32. */
int fdb_contract(const float *y_sketch, const float *x_sketch, int64_t n_spots, int32_t d,
                 int32_t n_types, float *h, float *ysq, void *stream);

/* (a1 + a3 + a4) production form: never materialises Y_s.  Streams the CSR once and writes
 *   h[out(i)][:] = X_s . y_s,row(i)      (float32, Kp per row)        ysq[out(i)] = ||y_s,row(i)||^2
 * for i in [0, n_spots), where row(i) = row_ids ? row_ids[i] : i is the CSR row processed and
 * out(i) = row_map ? row_map[row(i)] : i.  row_map places results in tile order (single GPU);
 * row_ids selects the rows of one spatial tile (multi-GPU: row_ids = order[lo:hi], out = i).
 * x_sketch_t is the TRANSPOSED sketched reference, d x Kp row-major (padding columns zero).
 * n_selected = number of genes with gene_bucket >= 0 (sizes the shared-memory tables of the fastest
 * kernel), or -1 if unknown (a slower kernel that needs no such count is used). */
int fdb_sketch_contract_csr(const void *indptr, int indptr_is_int64, const int32_t *indices,
                            const float *counts, int64_t n_spots, int32_t n_genes,
                            const int32_t *gene_bucket, const float *gene_weight, int32_t d,
                            const float *x_sketch_t, int32_t n_types, const int32_t *row_map,
                            const int32_t *row_ids, int32_t n_selected, float *h, float *ysq, void *stream);

/* The same pass for FlashDeconv._preprocess_data's LINEAR branches, "raw" (core/deconv.py:227-229) and "pearson"
 * (:199-225): the preprocessed value of an entry is count * f_g with a per-gene factor f_g (1, or 1 / sigma_g with
 * sigma_g^2 = mu_g + mu_g^2 / 100, mu_g = column mean + 1e-6), which the caller folds into gene_weight; no library
 * size, no log1p.  Same arguments and outputs as fdb_sketch_contract_csr. */
int fdb_sketch_linear_contract_csr(const void *indptr, int indptr_is_int64, const int32_t *indices,
                                   const float *counts, int64_t n_spots, int32_t n_genes,
                                   const int32_t *gene_bucket, const float *gene_weight, int32_t d,
                                   const float *x_sketch_t, int32_t n_types, const int32_t *row_map,
                                   const int32_t *row_ids, int32_t n_selected, float *h, float *ysq, void *stream);

/* Multi-GPU form of the two calls above (linear != 0: the linear branches): this rank sketches the `n_spots` input rows
 * of the CSR slice it holds; row_map[i] is the global TILE position of slice row i, and the row lands in the H / ysq
 * buffers of the rank that owns that position (host_bounds[q] <= p < host_bounds[q + 1]) -- written over NVLink
 * through the peer mappings host_h[q] / host_ysq[q] (as seen from this process; own rows x Kp floats, own rows floats).
 * The sketch and its all-to-all in one kernel.  Needs n_selected >= 0. */
int fdb_sketch_contract_scatter_csr(const void *indptr, int indptr_is_int64, const int32_t *indices,
                                    const float *counts, int64_t n_spots, int32_t n_genes,
                                    const int32_t *gene_bucket, const float *gene_weight, int32_t d,
                                    const float *x_sketch_t, int32_t n_types, const int32_t *row_map,
                                    int32_t n_selected, int32_t linear, int32_t world, const int32_t *host_bounds,
                                    void *const *host_h, void *const *host_ysq, void *stream);

/* Per-gene sums of the raw counts, float64, ACCUMULATED into sums[n_genes] (zero it first): the column means of
 * preprocess "pearson" (Y.mean(axis=0), core/deconv.py:207). */
int fdb_gene_sums_csr(const int32_t *indices, const float *counts, int64_t nnz, int32_t n_genes, double *sums,
                      void *stream);

/* ---------------------------------------------------------------------------------------
 * (a5) spatial graph.  Replaces build_knn_graph (utils/graph.py:25-83), build_radius_graph
 * (:86-133) and build_grid_graph (:136-172): float64 squared distances, k nearest OTHER
 * spots (ties -> smaller original index), union-symmetrised, binary, ascending columns.
 *
 * fdb_graph_build (syncs: bounding box, cell table upload, nnz; "grid" additionally reads its radius back):
 *   coords        n_spots x 2 float64, input order
 *   mode          0 = kNN with `k`; 1 = radius graph with `radius` (d <= radius);
 *                 2 = "grid": radius = 1.5 * median nearest-neighbour distance
 *   order, rank   out, int32[n_spots]  (tile order <-> input order)
 *   indptr        out, int32[n_spots + 1], tile order
 *   indices       out, int32[indices_capacity], neighbours as tile-order positions, ascending
 *   host_nnz      out (host), number of stored entries
 *   host_radius   out (host, may be NULL), radius actually used for modes 1/2
 * Returns FDB_ERR_WORKSPACE if indices_capacity is too small (host_nnz then holds the need).
 * ------------------------------------------------------------------------------------- */
int64_t fdb_graph_workspace_bytes(int64_t n_spots, int32_t k);
/* The same for coords with `dims` = 1, 2 or 3 columns (row-major n_spots x dims) and any k <= 1024.  2-D coordinates
 * with k <= 32 run the grid-hash kernels; everything else an exhaustive search in float64 (exact, O(n^2): a fallback).
 * "grid" takes its median on the device (radix select); only the resulting radius crosses PCIe. */
int fdb_graph_build_nd(const double *coords, int64_t n_spots, int32_t dims, int32_t mode, int32_t k, double radius,
                       int32_t *order, int32_t *rank, int32_t *indptr, int32_t *indices,
                       int64_t indices_capacity, int64_t *host_nnz, double *host_radius,
                       void *workspace, int64_t workspace_bytes, void *stream);
int fdb_graph_build(const double *coords, int64_t n_spots, int32_t mode, int32_t k, double radius,
                    int32_t *order, int32_t *rank, int32_t *indptr, int32_t *indices,
                    int64_t indices_capacity, int64_t *host_nnz, double *host_radius,
                    void *workspace, int64_t workspace_bytes, void *stream);

/* Adjacency relabelled to INPUT order with ascending columns (what FlashDeconv.adjacency_
 * exposes, core/deconv.py:364).  out_indptr int32[n+1], out_indices int32[nnz]. */
int fdb_graph_to_input_order(const int32_t *indptr, const int32_t *indices, const int32_t *order,
                             const int32_t *rank, int64_t n_spots, int32_t *out_indptr,
                             int32_t *out_indices, void *workspace, int64_t workspace_bytes,
                             void *stream);

/* ---------------------------------------------------------------------------------------
 * (a8 + a9) Jacobi block-coordinate-descent.  Replaces _bcd_iteration_fused
 * (core/solver.py:104-184), update_spot_with_Xty (:29-101) and the loop of bcd_solve
 * (:385-413).
 *
 * State block (device, 64 bytes, zero it before the first sweep):
 *   [0] uint32 max|delta| bits   [1] uint32 max|old| bits   [2] uint32 blocks arrived
 *   [3] int32  sweeps completed  [4] int32 converged flag   [5] float  last rel_change
 *
 * fdb_bcd_sweep: one sweep over rows [0, n_rows) of beta_in -> beta_out (both n_total x Kp;
 * rows >= n_rows are halo rows that are only read).  If `finalize` != 0 the last block to
 * finish computes rel_change = max|delta| / (max|old| + 1e-10), bumps the sweep counter and
 * raises the converged flag when rel_change < tol; sweeps launched after the flag is up
 * return immediately (so a fixed number of launches can be enqueued with no host sync).
 * host_gram: K x K float32 (host memory; passed to the kernel by value).
 * ------------------------------------------------------------------------------------- */
int fdb_bcd_sweep(const float *h, const float *host_gram, const float *beta_in, float *beta_out,
                  const int32_t *indptr, const int32_t *indices, int64_t n_rows, int32_t n_types,
                  float lambda, float rho_scaled, float tol, int32_t finalize, void *state,
                  const void *plan, void *stream);

/* Gather plan: per-patch halo row lists + 16-bit neighbour codes, built ONCE per graph (the adjacency does not
 * change between sweeps) into a caller-provided device buffer of fdb_bcd_plan_bytes() bytes (0 = this row width
 * needs none).  Passing it to the sweep / solve entry points selects the fastest sweep kernel (shared-memory
 * fp16 gather tile); `plan` may be NULL everywhere (slower fp32 kernel, identical semantics). */
int64_t fdb_bcd_plan_bytes(int64_t n_rows, int64_t nnz, int32_t n_types);
int fdb_bcd_plan_build(const int32_t *indptr, const int32_t *indices, int64_t n_rows, int64_t nnz,
                       int32_t n_types, void *plan, int64_t plan_bytes, void *stream);

/* Single-thread kernel doing the finalize step on an externally reduced state block (the
 * multi-GPU path all-reduces words [0],[1] with MAX between sweep and finalize). */
int fdb_bcd_finalize(void *state, float tol, void *stream);

/* beta[:, :K] <- 1/K (core/solver.py:372), padding columns <- 0, state block <- 0 (state may be NULL). */
int fdb_bcd_init(float *beta, int64_t n_rows, int32_t n_types, void *state, void *stream);

/* Enqueues beta <- 1/K, then up to max_iter sweeps ping-ponging beta_a/beta_b (single GPU).
 * After a stream sync, state[3] holds n_iterations and the result is in beta_a when that
 * count is even, beta_b when odd. */
int fdb_bcd_solve(const float *h, const float *host_gram, float *beta_a, float *beta_b,
                  const int32_t *indptr, const int32_t *indices, int64_t n_rows, int32_t n_types,
                  float lambda, float rho_scaled, int32_t max_iter, float tol, void *state,
                  const void *plan, void *stream);

/* (a10) objective pieces in float64: out[0]=sum(beta*H) out[1]=sum_i b_i^T G b_i
 * out[2]=Tr(b^T L b) out[3]=sum|beta| out[4]=sum ysq.  Replaces compute_objective
 * (core/solver.py:226-284) and compute_laplacian (core/spatial.py:35-73; L is never built).
 * `out` (device, 5 doubles) is accumulated into: zero it first. */
int fdb_objective_terms(const float *beta, const float *h, const float *ysq, const float *host_gram,
                        const int32_t *indptr, const int32_t *indices, int64_t n_rows,
                        int32_t n_types, double *out, void *stream);

/* (a11) un-permute + widen: beta_out[order[p]][k] = beta[p][k] (float64, n x K) and the
 * row-normalised proportions (all-zero row -> 1/K).  Replaces normalize_proportions
 * (core/solver.py:431-452).  Either output may be NULL. */
int fdb_finish(const float *beta, const int32_t *order, int64_t n_rows, int32_t n_types,
               double *beta_out, double *prop_out, void *stream);

/* (f3) dominant cell type per spot: argmax_k beta[p][k] (first maximum on ties, numpy.argmax), written at
 * dominant[order[p]] (order may be NULL).  Replaces FlashDeconv.get_dominant_cell_type (core/deconv.py:467-478). */
int fdb_dominant_type(const float *beta, const int32_t *order, int64_t n_rows, int32_t n_types, int32_t *dominant,
                      void *stream);

/* ---------------------------------------------------------------------------------------
 * Any number of cell types (FDB_MAX_TYPES < K <= FDB_MAX_TYPES_WIDE).  The reference has no limit on K
 * (core/solver.py:287-428); these entry points restate the same sweep (maintained-residual coordinate
 * descent, core/solver.py:29-101, Jacobi neighbour sums, :104-184), stop test and objective with one warp
 * per spot.  gram_dev: DEVICE pointer, Kp x Kp float32 row-major, zero padded, diagonal included (no plan,
 * no host Gram).  State block, buffer ping-pong and n_iterations protocol as fdb_bcd_solve / fdb_bcd_sweep.
 * The sketch for such K is fdb_sketch_logcpm_csr / fdb_sketch_project_csr + fdb_contract (any K), the outputs
 * fdb_finish / fdb_dominant_type (any K).  They also accept K <= FDB_MAX_TYPES (tests compare the two paths).
 * ------------------------------------------------------------------------------------- */
int fdb_bcd_solve_wide(const float *h, const float *gram_dev, float *beta_a, float *beta_b,
                       const int32_t *indptr, const int32_t *indices, int64_t n_rows, int32_t n_types,
                       float lambda, float rho_scaled, int32_t max_iter, float tol, void *state, void *stream);
int fdb_bcd_sweep_wide(const float *h, const float *gram_dev, const float *beta_in, float *beta_out,
                       const int32_t *indptr, const int32_t *indices, int64_t n_rows, int32_t n_types,
                       float lambda, float rho_scaled, float tol, int32_t finalize, void *state, void *stream);
int fdb_objective_terms_wide(const float *beta, const float *h, const float *ysq, const float *gram_dev,
                             const int32_t *indptr, const int32_t *indices, int64_t n_rows,
                             int32_t n_types, double *out, void *stream);

/* (f3) per-group sums of a cells x genes CSR matrix, float64, ACCUMULATED into sums[n_groups x n_genes] (zero it
 * first); labels[i] < 0 skips cell i.  Divided by the group sizes this is the cell-type signature matrix of
 * load_reference / prepare_data (io/loader.py:119-136). */
int fdb_group_sums_csr(const void *indptr, int indptr_is_int64, const int32_t *indices, const float *values,
                       const int32_t *labels, int64_t n_rows, int32_t n_genes, int32_t n_groups, double *sums,
                       void *stream);

/* ---------------------------------------------------------------------------------------
 * (f1) gene statistics for HVG selection: per-gene sum and sum of squares of
 * log1p(1e4 * count / max(lib_all_genes, 1)).  Replaces the O(nnz) part of select_hvg
 * (utils/genes.py:52-83).  sums / sumsq: float64[n_genes], accumulated into (zero first).
 * ------------------------------------------------------------------------------------- */
int fdb_gene_moments_csr(const void *indptr, int indptr_is_int64, const int32_t *indices,
                         const float *counts, int64_t n_spots, int32_t n_genes, double *sums,
                         double *sumsq, void *stream);

/* ---------------------------------------------------------------------------------------
 * Multi-GPU (one process per GPU; no reference counterpart -- the reference is single process).
 * The communicator is NCCL, bound at run time (dlopen); the 128-byte unique id from rank 0's
 * fdb_comm_unique_id is distributed by the host layer (torch.distributed, MPI, a file ...).
 *
 * fdb_bcd_solve_tiled: the loop of fdb_bcd_solve for one spatial tile.  beta_a / beta_b hold
 * n_total = n_own + n_halo rows; halo slice [n_own + recv_first[r], +recv_count[r]) is owned by
 * recv_peer[r]; send_rows[s] (device, int32 local row ids) are the rows send_peer[s] needs, staged
 * through send_buf[s] (device, send_count[s] x Kp floats).  Per sweep: sweep -> pack -> grouped
 * ncclSend/ncclRecv -> ncclAllReduce(MAX) of the two max-norm words -> stop test, all on `stream`.
 * The host_* arrays are HOST arrays (of ints / of device pointers).
 * ------------------------------------------------------------------------------------- */
int fdb_comm_unique_id(char *host_id_128);
int fdb_comm_init(int32_t rank, int32_t world, const char *host_id_128, void **host_comm_out);
int fdb_comm_destroy(void *comm);
int fdb_bcd_solve_tiled(const float *h, const float *host_gram, float *beta_a, float *beta_b,
                        const int32_t *indptr, const int32_t *indices, int64_t n_own, int64_t n_total,
                        int32_t n_types, float lambda, float rho_scaled, int32_t max_iter, float tol,
                        void *state, int32_t n_recv, const int32_t *host_recv_peer,
                        const int64_t *host_recv_first, const int64_t *host_recv_count, int32_t n_send,
                        const int32_t *host_send_peer, const int32_t *const *host_send_rows,
                        const int64_t *host_send_count, float *const *host_send_buf, void *comm,
                        const void *plan, void *stream);

/* Partition of the solve over `world` ranks, derived on the device from the replicated global graph (tile order).
 * Rank r owns positions [host_bounds[r], host_bounds[r + 1]).  Phase 1 (fdb_tile_plan_counts, identical on every rank,
 * ONE stream synchronisation): host_counts[r * world + q] = number of rows of r with a neighbour owned by q, and
 * host_e = {indptr[lo], indptr[hi]} of `rank`.  From the counts the host derives every rank's halo size, the
 * capacity of the symmetric beta buffers and host_base[q] = n_own(q) + sum_{r < rank} counts[r][q], the first row
 * of peer q's buffers that receives this rank's boundary rows.  Phase 2 (fdb_tile_plan_build, no synchronisation)
 * writes this rank's local adjacency (own rows 0 .. n_own-1, halo rows after them in ascending global position),
 * the halo row list, the per-row push entries, and the patch order / boundary-patch count the fused sweep kernel
 * consumes (see fdb_bcd_solve_peer).  Same workspace for both phases (fdb_tile_plan_workspace_bytes). */
int64_t fdb_tile_plan_workspace_bytes(int64_t n, int64_t n_own_max, int32_t world);
int fdb_tile_plan_counts(const int32_t *indptr, const int32_t *indices, int64_t n, const int32_t *host_bounds,
                         int32_t world, int32_t rank, int64_t n_own_max, void *workspace, int64_t workspace_bytes,
                         int64_t *host_counts, int64_t *host_e, void *stream);
int fdb_tile_plan_build(const int32_t *indptr, const int32_t *indices, int64_t n, const int32_t *host_bounds,
                        int32_t world, int32_t rank, int64_t n_own_max, const int64_t *host_base, void *workspace,
                        int64_t workspace_bytes, int32_t *local_ptr, int32_t *local_idx, int64_t *halo_global,
                        int32_t *push_ptr, void *push_ent, int32_t *patch_order, int32_t *n_boundary, void *stream);

/* Peer-memory form of fdb_bcd_solve_tiled: no NCCL on the data path.  Every rank's beta buffers live in a
 * symmetric allocation mapped by all peers; host_peer_base[p] is rank p's base pointer as seen from THIS
 * process.  Layout (floats): beta_a [cap_rows x Kp], beta_b [cap_rows x Kp], comm [fdb_peer_comm_floats()],
 * identical on all ranks (cap_rows >= every rank's n_total).
 * Boundary rows are described per own row (device arrays): entries push_ptr[i] .. push_ptr[i + 1] of push_ent, each
 * an int32 pair (peer, row of the peer's buffers that receives own row i).  patch_order (int32, one entry per
 * 128-row patch of the own rows, patches that hold boundary rows FIRST) and n_boundary (device int32: how many
 * patches that is) let the sweep kernel write the boundary rows into the neighbours' halo slots from its own store
 * phase, ahead of the interior patches; its last block then publishes the two max-norm words + a sequence number to
 * every peer, waits for all peers' and applies the stop test: ONE launch per sweep.  patch_order / n_boundary may be
 * NULL (natural order; push + hand-shake then run as separate launches, as they also do with the fp32 fallback sweep).
 * seq_base must grow by more than max_iter between successive solves on the same allocation. */
int64_t fdb_peer_comm_floats(void);
int fdb_bcd_solve_peer(const float *h, const float *host_gram, void *const *host_peer_base, int32_t rank,
                       int32_t world, int64_t cap_rows, const int32_t *indptr, const int32_t *indices,
                       int64_t n_own, int64_t n_total, int32_t n_types, float lambda, float rho_scaled,
                       int32_t max_iter, float tol, void *state, const int32_t *push_ptr, const void *push_ent,
                       const int32_t *patch_order, const int32_t *n_boundary, uint32_t seq_base,
                       const void *plan, void *stream);

/* Multi-GPU helpers: gather / scatter whole beta rows by index list (halo exchange staging). */
int fdb_rows_gather(const float *src, const int32_t *rows, int64_t n_list, int32_t row_floats,
                    float *dst, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FDB200_H */
