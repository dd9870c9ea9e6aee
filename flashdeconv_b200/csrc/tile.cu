// Multi-GPU partition of the solve: local adjacency, halo rows and boundary-row push lists of ONE rank, derived on
// the device from the replicated global graph (tile order).  Host-side counterpart and documentation of the layout:
// flashdeconv_b200/tiling.py (plan_tile is the reference implementation these kernels are tested against).
//
// The ranks own contiguous position ranges [bounds[r], bounds[r + 1]).  Because the graph is undirected, the rows a
// peer q needs from rank r are exactly r's rows with a neighbour in q's range; a peer's halo rows are ordered by
// global position, i.e. grouped by owner rank and ascending inside a group.  So the halo slot of own row i in peer q is
//     n_own(q) + sum_{r' < r} count[r'][q] + (rank of i among r's rows that q needs),
// where count[r][q] = number of rows of r with a neighbour in q.  Every rank computes the whole R x R matrix from its
// own copy of the graph: no collective, ONE small device -> host read (the matrix and two row pointers).
#include <vector>
#include "bcd_common.cuh"

namespace fdb {

struct TileBounds {
    int32_t b[kMaxRanks + 1];
};
__device__ __forceinline__ int owner_of(const TileBounds &tb, int R, int p)
{
    int r = 0;
#pragma unroll 4
    for (int q = 1; q < R; ++q) r += p >= tb.b[q];
    return r;
}

// every row of the global graph: bitmask of the OTHER ranks that own a neighbour; counts[r][q] over the boundary rows
__global__ void __launch_bounds__(256)
tile_mask_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices, int n, TileBounds tb, int R,
                 uint16_t *__restrict__ row_mask, int32_t *__restrict__ counts)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = owner_of(tb, R, i);
    const int lo = tb.b[r], hi = tb.b[r + 1];
    unsigned mask = 0u;
    const int e1 = indptr[i + 1];
    for (int e = indptr[i]; e < e1; ++e) {
        const int j = indices[e];
        if (j < lo || j >= hi) mask |= 1u << owner_of(tb, R, j);
    }
    row_mask[i] = (uint16_t)mask;
    while (mask) {
        const int q = __ffs(mask) - 1;
        mask &= mask - 1;
        atomicAdd(counts + r * R + q, 1);
    }
}

// own rows: local row pointers, per-peer flags (peer-major, for the rank scan), push counts, halo flags
__global__ void __launch_bounds__(256)
tile_flag_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices, int lo, int hi, int R,
                 const uint16_t *__restrict__ row_mask, int32_t *__restrict__ local_ptr, int32_t *__restrict__ qflag,
                 int32_t *__restrict__ push_cnt, int32_t *__restrict__ hflag)
{
    const int n_own = hi - lo;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_own) return;
    const int e0 = indptr[lo];
    local_ptr[i] = indptr[lo + i] - e0;
    if (i == n_own) return;
    const unsigned mask = row_mask[lo + i];
    push_cnt[i] = __popc(mask);
    for (int q = 0; q < R; ++q) qflag[(int64_t)q * n_own + i] = (mask >> q) & 1u;
    if (mask) {
        const int e1 = indptr[lo + i + 1];
        for (int e = indptr[lo + i]; e < e1; ++e) {
            const int j = indices[e];
            if (j < lo || j >= hi) hflag[j] = 1;
        }
    }
}

struct TileBase {
    int32_t base[kMaxRanks];        // first halo slot of this rank's rows in peer q (its own rows included)
};

// own rows: local neighbour ids and the (peer, destination row) entries of the boundary rows
__global__ void __launch_bounds__(256)
tile_emit_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices, int lo, int hi, int R,
                 const uint16_t *__restrict__ row_mask, const int32_t *__restrict__ hscan,
                 const int32_t *__restrict__ qscan, const int32_t *__restrict__ push_ptr, TileBase tbase,
                 int32_t *__restrict__ local_idx, int2 *__restrict__ push_ent)
{
    const int n_own = hi - lo;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_own) return;
    const int e0 = indptr[lo], e1 = indptr[lo + i + 1];
    for (int e = indptr[lo + i]; e < e1; ++e) {
        const int j = indices[e];
        local_idx[e - e0] = (j < lo || j >= hi) ? n_own + hscan[j] : j - lo;
    }
    unsigned mask = row_mask[lo + i];
    int k = push_ptr[i];
    while (mask) {
        const int q = __ffs(mask) - 1;
        mask &= mask - 1;
        const int64_t at = (int64_t)q * n_own;
        push_ent[k++] = make_int2(q, tbase.base[q] + qscan[at + i] - qscan[at]);
    }
}

// all positions: halo rows in ascending global position
__global__ void __launch_bounds__(256)
tile_halo_kernel(const int32_t *__restrict__ hflag_scan_in, const int32_t *__restrict__ hflag, int n,
                 int64_t *__restrict__ halo_global)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n && hflag[p]) halo_global[hflag_scan_in[p]] = p;
}

// patches of `patch` own rows: does the patch hold a boundary row?
__global__ void __launch_bounds__(256)
tile_patch_flag_kernel(const int32_t *__restrict__ push_ptr, int n_own, int patch, int n_patches, int32_t *__restrict__ pflag)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_patches) return;
    const int a = min(p * patch, n_own), b = min(a + patch, n_own);
    pflag[p] = push_ptr[b] > push_ptr[a];
}
// boundary patches first, both groups in ascending order
__global__ void __launch_bounds__(256)
tile_patch_order_kernel(const int32_t *__restrict__ pflag, const int32_t *__restrict__ pscan, int n_patches,
                        int32_t *__restrict__ order, int32_t *__restrict__ n_boundary)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_patches) return;
    const int nb = pscan[n_patches];
    if (p == 0) *n_boundary = nb;
    order[pflag[p] ? pscan[p] : nb + (p - pscan[p])] = p;
}

}  // namespace fdb

using namespace fdb;
#define FDB_API extern "C" __attribute__((visibility("default")))

static int64_t al(int64_t x) { return round_up(x, 256); }

struct TileWs {
    uint16_t *row_mask;
    int32_t *counts, *hflag, *hscan, *qflag, *push_cnt, *pflag, *pscan, *block_sums;
    int64_t bytes;
};
static TileWs carve_tile(void *workspace, int64_t n, int64_t n_own_max, int world)
{
    const int64_t n_patches = ceil_div(n_own_max > 0 ? n_own_max : 1, 128);
    char *ws = (char *)workspace;
    TileWs w;
    w.row_mask = (uint16_t *)ws;    ws += al(2 * n);
    w.counts = (int32_t *)ws;       ws += al(4 * ((int64_t)world * world + 2));
    w.hflag = (int32_t *)ws;        ws += al(4 * (n + 1));
    w.hscan = (int32_t *)ws;        ws += al(4 * (n + 1));
    w.qflag = (int32_t *)ws;        ws += al(4 * ((int64_t)world * n_own_max + 1));
    w.push_cnt = (int32_t *)ws;     ws += al(4 * (n_own_max + 1));
    w.pflag = (int32_t *)ws;        ws += al(4 * (n_patches + 1));
    w.pscan = (int32_t *)ws;        ws += al(4 * (n_patches + 1));
    w.block_sums = (int32_t *)ws;   ws += al(4 * scan_blocks(std::max<int64_t>(n + 1, (int64_t)world * n_own_max + 1)));
    w.bytes = ws - (char *)workspace;
    return w;
}

FDB_API int64_t fdb_tile_plan_workspace_bytes(int64_t n, int64_t n_own_max, int32_t world)
{
    return carve_tile(nullptr, n, n_own_max, world).bytes + 256;
}

/* Phase 1 (all ranks identically): boundary-row masks of every row and the R x R count matrix.  Syncs once:
 * host_counts[r * world + q] and the two row pointers host_e[0] = indptr[lo], host_e[1] = indptr[hi] of `rank`. */
FDB_API int fdb_tile_plan_counts(const int32_t *indptr, const int32_t *indices, int64_t n, const int32_t *host_bounds,
                                 int32_t world, int32_t rank, int64_t n_own_max, void *workspace, int64_t workspace_bytes,
                                 int64_t *host_counts, int64_t *host_e, void *stream)
{
    FDB_REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "world must be in [1, %d]", kMaxRanks);
    FDB_REQUIRE(n >= 0 && n < ((int64_t)1 << 31) - 256 && indptr && host_bounds && workspace && host_counts && host_e, "bad arguments");
    if (workspace_bytes < fdb_tile_plan_workspace_bytes(n, n_own_max, world)) {
        set_error("tile plan workspace too small");
        return FDB_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const TileWs w = carve_tile(workspace, n, n_own_max, world);
    TileBounds tb;
    for (int r = 0; r <= kMaxRanks; ++r) tb.b[r] = host_bounds[std::min(r, (int)world)];
    FDB_CUDA(cudaMemsetAsync(w.counts, 0, 4 * ((size_t)world * world + 2), st));
    if (n > 0) {
        tile_mask_kernel<<<(int)ceil_div(n, 256), 256, 0, st>>>(indptr, indices, (int)n, tb, world, w.row_mask, w.counts);
        FDB_LAUNCH_CHECK("tile_mask_kernel");
    }
    std::vector<int32_t> hc((size_t)world * world);
    int32_t e[2] = {0, 0};
    FDB_CUDA(cudaMemcpyAsync(hc.data(), w.counts, hc.size() * 4, cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaMemcpyAsync(&e[0], indptr + host_bounds[rank], 4, cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaMemcpyAsync(&e[1], indptr + host_bounds[rank + 1], 4, cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    for (size_t i = 0; i < hc.size(); ++i) host_counts[i] = hc[i];
    host_e[0] = e[0];
    host_e[1] = e[1];
    return FDB_OK;
}

/* Phase 2: this rank's plan, enqueued on `stream` (no synchronisation).  Uses the workspace of phase 1 (row masks).
 *   host_base[q]   first row of peer q's buffers that receives this rank's boundary rows
 *   local_ptr      int32[n_own + 1]      local_idx  int32[nnz_local]  (own rows 0..n_own-1, halo rows after)
 *   halo_global    int64[n_halo]         push_ptr   int32[n_own + 1]   push_ent int32[2 * n_push]
 *   patch_order    int32[ceil(n_own / 128)]   n_boundary  int32[1] */
FDB_API int fdb_tile_plan_build(const int32_t *indptr, const int32_t *indices, int64_t n, const int32_t *host_bounds,
                                int32_t world, int32_t rank, int64_t n_own_max, const int64_t *host_base, void *workspace,
                                int64_t workspace_bytes, int32_t *local_ptr, int32_t *local_idx, int64_t *halo_global,
                                int32_t *push_ptr, void *push_ent, int32_t *patch_order, int32_t *n_boundary, void *stream)
{
    FDB_REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "world must be in [1, %d]", kMaxRanks);
    FDB_REQUIRE(indptr && host_bounds && host_base && workspace && local_ptr && local_idx && halo_global && push_ptr && push_ent &&
                patch_order && n_boundary, "null pointer");
    if (workspace_bytes < fdb_tile_plan_workspace_bytes(n, n_own_max, world)) {
        set_error("tile plan workspace too small");
        return FDB_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int lo = host_bounds[rank], hi = host_bounds[rank + 1], n_own = hi - lo;
    FDB_REQUIRE(n_own >= 0 && n_own <= n_own_max, "n_own_max too small");
    const int n_patches = (int)ceil_div(n_own > 0 ? n_own : 1, 128);
    const TileWs w = carve_tile(workspace, n, n_own_max, world);
    TileBase tbase;
    for (int q = 0; q < kMaxRanks; ++q) tbase.base[q] = q < world ? (int32_t)host_base[q] : 0;
    FDB_CUDA(cudaMemsetAsync(w.hflag, 0, 4 * (size_t)(n + 1), st));
    tile_flag_kernel<<<(int)ceil_div(n_own + 1, 256), 256, 0, st>>>(indptr, indices, lo, hi, world, w.row_mask, local_ptr,
                                                                    w.qflag, w.push_cnt, w.hflag);
    FDB_LAUNCH_CHECK("tile_flag_kernel");
    int rc = exclusive_scan(w.qflag, (int64_t)world * n_own, w.qflag, w.block_sums, st);        // rank per (peer, row)
    if (rc) return rc;
    rc = exclusive_scan(w.push_cnt, n_own, push_ptr, w.block_sums, st);
    if (rc) return rc;
    rc = exclusive_scan(w.hflag, n, w.hscan, w.block_sums, st);
    if (rc) return rc;
    if (n_own > 0) {
        tile_emit_kernel<<<(int)ceil_div(n_own, 256), 256, 0, st>>>(indptr, indices, lo, hi, world, w.row_mask, w.hscan,
                                                                    w.qflag, push_ptr, tbase, local_idx, (int2 *)push_ent);
        FDB_LAUNCH_CHECK("tile_emit_kernel");
    }
    if (n > 0) {
        tile_halo_kernel<<<(int)ceil_div(n, 256), 256, 0, st>>>(w.hscan, w.hflag, (int)n, halo_global);
        FDB_LAUNCH_CHECK("tile_halo_kernel");
    }
    tile_patch_flag_kernel<<<(int)ceil_div(n_patches, 256), 256, 0, st>>>(push_ptr, n_own, 128, n_patches, w.pflag);
    FDB_LAUNCH_CHECK("tile_patch_flag_kernel");
    rc = exclusive_scan(w.pflag, n_patches, w.pscan, w.block_sums, st);
    if (rc) return rc;
    tile_patch_order_kernel<<<(int)ceil_div(n_patches, 256), 256, 0, st>>>(w.pflag, w.pscan, n_patches, patch_order, n_boundary);
    FDB_LAUNCH_CHECK("tile_patch_order_kernel");
    return FDB_OK;
}
