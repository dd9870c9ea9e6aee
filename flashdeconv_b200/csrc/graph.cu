// Kernel family 3: grid-hashed 2-D spatial graph (kNN / radius / "grid") + tile ordering.
//
// Reference semantics (upstream file:line):
//   utils/graph.py:47-83    k_actual = min(k, N-1); cKDTree.query(k+1); drop index == row;
//                           A = A + A^T; data = 1  (union symmetrisation, binary)
//   utils/graph.py:108-133  all pairs with distance <= radius (query_pairs is inclusive)
//   utils/graph.py:157-172  "grid": radius = 1.5 * median nearest-neighbour distance
// Distances are float64 squared sums evaluated without FMA contraction so that the order of
// candidates is the order cKDTree sees; exact ties fall to the smaller original index.
//
// Pipeline (all on `stream`):
//   bbox -> [host: cell size, 8x8-cell tiles ranked along a Morton curve] -> cell histogram
//   -> scan -> scatter -> in-cell sort (determinism) -> ring search (thread per spot, top-k in
//   registers) -> reverse-edge test -> degree scan -> fill -> per-row sort.
// Spots are renumbered in cell order ("tile order"): position p <-> original index order[p].
// Everything downstream (H, beta, adjacency) lives in tile order, so a BCD thread block reads
// neighbours that are a few KB away in memory, and contiguous position ranges are compact
// spatial tiles for the multi-GPU partition.
#include <algorithm>
#include <cmath>
#include <vector>
#include "fdb_common.cuh"

namespace fdb {

struct GridSpec {
    double x0, y0;        // lower corner
    double inv_cell;      // 1 / cell side
    double cell;          // cell side
    int gx, gy;           // cells per axis
    int tw, th;           // tile shape in cells (powers of two <= 8)
    int tiles_x, tiles_y;
    int n_cells;          // tiles_x * tiles_y * tw * th
};

__device__ __forceinline__ int cell_id(const GridSpec &g, const int32_t *__restrict__ tile_rank,
                                       int cx, int cy)
{
    const int tx = cx / g.tw, ty = cy / g.th;
    const int lx = cx - tx * g.tw, ly = cy - ty * g.th;
    return (__ldg(tile_rank + ty * g.tiles_x + tx) * g.th + ly) * g.tw + lx;
}

__device__ __forceinline__ void cell_of_point(const GridSpec &g, double x, double y, int &cx, int &cy)
{
    cx = (int)floor((x - g.x0) * g.inv_cell);
    cy = (int)floor((y - g.y0) * g.inv_cell);
    cx = min(max(cx, 0), g.gx - 1);
    cy = min(max(cy, 0), g.gy - 1);
}

// ---------------------------------------------------------------- bounding box
__global__ void __launch_bounds__(256)
bbox_kernel(const double *__restrict__ coords, int64_t n, double *__restrict__ partial)
{
    double lo_x = INFINITY, lo_y = INFINITY, hi_x = -INFINITY, hi_y = -INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double2 c = *reinterpret_cast<const double2 *>(coords + 2 * i);
        lo_x = fmin(lo_x, c.x); hi_x = fmax(hi_x, c.x);
        lo_y = fmin(lo_y, c.y); hi_y = fmax(hi_y, c.y);
    }
    __shared__ double sm[4][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo_x = fmin(lo_x, __shfl_xor_sync(kFull, lo_x, o));
        lo_y = fmin(lo_y, __shfl_xor_sync(kFull, lo_y, o));
        hi_x = fmax(hi_x, __shfl_xor_sync(kFull, hi_x, o));
        hi_y = fmax(hi_y, __shfl_xor_sync(kFull, hi_y, o));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sm[0][warp] = lo_x; sm[1][warp] = lo_y; sm[2][warp] = hi_x; sm[3][warp] = hi_y; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            sm[0][0] = fmin(sm[0][0], sm[0][w]); sm[1][0] = fmin(sm[1][0], sm[1][w]);
            sm[2][0] = fmax(sm[2][0], sm[2][w]); sm[3][0] = fmax(sm[3][0], sm[3][w]);
        }
        for (int c = 0; c < 4; ++c) partial[4 * blockIdx.x + c] = sm[c][0];
    }
}

// ---------------------------------------------------------------- exclusive scan (int32)
constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads)
scan_reduce_kernel(const int32_t *__restrict__ in, int64_t n, int32_t *__restrict__ block_sums)
{
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    int s = 0;
    for (int i = threadIdx.x; i < kScanTile; i += kScanThreads)
        if (base + i < n) s += in[base + i];
    __shared__ int sm[kScanThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) t += sm[w];
        block_sums[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024)
scan_block_sums_kernel(int32_t *__restrict__ block_sums, int n_blocks, int32_t *__restrict__ total_out)
{
    __shared__ int sm[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n_blocks; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n_blocks ? block_sums[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) sm[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = sm[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(kFull, w, o);
                if (lane >= o) w += t;
            }
            sm[lane] = w;
        }
        __syncthreads();
        const int offset = carry + (warp ? sm[warp - 1] : 0);
        if (i < n_blocks) block_sums[i] = offset + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = offset + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(const int32_t *__restrict__ in, int64_t n, const int32_t *__restrict__ block_sums,
                  int32_t *__restrict__ out)
{
    // thread owns kScanItems consecutive items
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int v[kScanItems];
    int s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = base + i < n ? in[base + i] : 0;
        s += v[i];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += t;
    }
    __shared__ int sm[kScanThreads / 32];
    if (lane == 31) sm[warp] = inc;
    __syncthreads();
    if (warp == 0 && lane < kScanThreads / 32) {
        int w = sm[lane];
#pragma unroll
        for (int o = 1; o < kScanThreads / 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffu, w, o);
            if (lane >= o) w += t;
        }
        sm[lane] = w;
    }
    __syncthreads();
    int run = block_sums[blockIdx.x] + (warp ? sm[warp - 1] : 0) + inc - s;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
}

// out has n + 1 entries (out[n] = total).  `in` and `out` may alias.  (Also used by tile.cu.)
int exclusive_scan(const int32_t *in, int64_t n, int32_t *out, int32_t *block_sums, cudaStream_t st)
{
    if (n == 0) {
        FDB_CUDA(cudaMemsetAsync(out, 0, 4, st));
        return FDB_OK;
    }
    const int n_blocks = (int)ceil_div(n, kScanTile);
    scan_reduce_kernel<<<n_blocks, kScanThreads, 0, st>>>(in, n, block_sums);
    scan_block_sums_kernel<<<1, 1024, 0, st>>>(block_sums, n_blocks, out + n);
    scan_apply_kernel<<<n_blocks, kScanThreads, 0, st>>>(in, n, block_sums, out);
    count_launches(2);
    FDB_LAUNCH_CHECK("exclusive_scan");
    return FDB_OK;
}
int64_t scan_blocks(int64_t n) { return ceil_div(n > 0 ? n : 1, kScanTile); }

// ---------------------------------------------------------------- cell sort
__global__ void __launch_bounds__(256)
cell_count_kernel(const double *__restrict__ coords, int64_t n, GridSpec g,
                  const int32_t *__restrict__ tile_rank, int32_t *__restrict__ cell_of,
                  int32_t *__restrict__ hist)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 c = *reinterpret_cast<const double2 *>(coords + 2 * i);
    int cx, cy;
    cell_of_point(g, c.x, c.y, cx, cy);
    const int id = cell_id(g, tile_rank, cx, cy);
    cell_of[i] = id;
    atomicAdd(hist + id, 1);
}

__global__ void __launch_bounds__(256)
cell_scatter_kernel(const int32_t *__restrict__ cell_of, int64_t n,
                    const int32_t *__restrict__ cell_start, int32_t *__restrict__ cursor,
                    int32_t *__restrict__ order)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int id = cell_of[i];
    order[cell_start[id] + atomicAdd(cursor + id, 1)] = (int32_t)i;
}

// thread per cell: order the cell's members by original index (run-to-run determinism),
// then publish rank[] and the tile-ordered coordinate copy
__global__ void __launch_bounds__(256)
cell_finish_kernel(const double *__restrict__ coords, int n_cells,
                   const int32_t *__restrict__ cell_start, int32_t *__restrict__ order,
                   int32_t *__restrict__ rank, double2 *__restrict__ sorted_xy)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const int s = cell_start[c], e = cell_start[c + 1];
    for (int a = s + 1; a < e; ++a) {
        const int v = order[a];
        int b = a - 1;
        while (b >= s && order[b] > v) { order[b + 1] = order[b]; --b; }
        order[b + 1] = v;
    }
    for (int a = s; a < e; ++a) {
        const int o = order[a];
        rank[o] = a;
        sorted_xy[a] = *reinterpret_cast<const double2 *>(coords + 2 * (int64_t)o);
    }
}

// ---------------------------------------------------------------- ring search (kNN)
__device__ __forceinline__ double dist2(double2 a, double2 b)
{
    const double dx = __dsub_rn(a.x, b.x), dy = __dsub_rn(a.y, b.y);
    return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));     // no FMA contraction
}

template <int KMAX>
__global__ void __launch_bounds__(128)
knn_kernel(const double2 *__restrict__ xy, const int32_t *__restrict__ order, int64_t n, GridSpec g,
           const int32_t *__restrict__ tile_rank, const int32_t *__restrict__ cell_start, int k,
           int32_t *__restrict__ knn, double *__restrict__ kth_dist)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const double2 me = xy[p];
    int cx, cy;
    cell_of_point(g, me.x, me.y, cx, cy);
    double bd[KMAX];
    int bi[KMAX];
#pragma unroll
    for (int i = 0; i < KMAX; ++i) { bd[i] = INFINITY; bi[i] = -1; }
    const double slack = 1e-7 * g.cell;
    const int max_ring = max(g.gx, g.gy);
    auto visit = [&](int xx, int yy) {
        if (xx < 0 || xx >= g.gx) return;
        const int id = cell_id(g, tile_rank, xx, yy);
        const int s = __ldg(cell_start + id), e = __ldg(cell_start + id + 1);
        for (int q = s; q < e; ++q) {
            if (q == p) continue;
            const double d2 = dist2(me, xy[q]);
            const double worst = bd[KMAX - 1];
            bool better = d2 < worst;
            if (!better && d2 == worst && bi[KMAX - 1] >= 0)
                better = __ldg(order + q) < __ldg(order + bi[KMAX - 1]);
            if (!better) continue;
            // insert, keeping (distance, original index) ascending
            double cd = d2;
            int ci = q;
#pragma unroll
            for (int i = 0; i < KMAX; ++i) {
                bool before = cd < bd[i];
                if (!before && cd == bd[i] && bi[i] >= 0 && ci >= 0)
                    before = __ldg(order + ci) < __ldg(order + bi[i]);
                if (before) {
                    const double td = bd[i]; const int ti = bi[i];
                    bd[i] = cd; bi[i] = ci; cd = td; ci = ti;
                }
            }
        }
    };
    for (int r = 0; r <= max_ring; ++r) {
        for (int dy = -r; dy <= r; ++dy) {
            const int yy = cy + dy;
            if (yy < 0 || yy >= g.gy) continue;
            if (dy == -r || dy == r) {
                for (int dx = -r; dx <= r; ++dx) visit(cx + dx, yy);
            } else {
                visit(cx - r, yy);
                visit(cx + r, yy);
            }
        }
        // everything not yet visited lies outside the (2r+1)-cell block around the home cell
        double bound = INFINITY;
        if (cx - r > 0) bound = fmin(bound, (me.x - g.x0) - (double)(cx - r) * g.cell);
        if (cx + r + 1 < g.gx) bound = fmin(bound, (double)(cx + r + 1) * g.cell - (me.x - g.x0));
        if (cy - r > 0) bound = fmin(bound, (me.y - g.y0) - (double)(cy - r) * g.cell);
        if (cy + r + 1 < g.gy) bound = fmin(bound, (double)(cy + r + 1) * g.cell - (me.y - g.y0));
        if (bound == INFINITY) break;                       // whole grid visited
        bound -= slack;
        // KMAX slots are tracked but only the first k matter
        double kth = INFINITY;
#pragma unroll
        for (int i = 0; i < KMAX; ++i) if (i == k - 1) kth = bd[i];
        if (bound > 0.0 && kth < bound * bound) break;
    }
#pragma unroll
    for (int i = 0; i < KMAX; ++i)
        if (i < k) knn[p * k + i] = bi[i];
    if (kth_dist) {
        double kth = INFINITY;
#pragma unroll
        for (int i = 0; i < KMAX; ++i) if (i == k - 1) kth = bd[i];
        kth_dist[p] = sqrt(kth);
    }
}

// ---------------------------------------------------------------- general path: any dimension D <= 3, any k
// utils/graph.py:15-22 accepts coords with D >= 1 columns and any k.  The grid-hash kernels above are 2-D with the k
// best candidates in registers (k <= 32); everything else -- 1-D / 3-D coordinates, k > 32 -- goes through exhaustive
// search in tile order (N^2 / 2 distance evaluations in float64: exact, ~5 ms at 100k spots, a fallback by design).
// Tile order comes from the first two coordinates (any permutation is valid; it only has to be spatially coherent).
struct double3p { double x, y, z; };

__global__ void __launch_bounds__(256)
project_xy_kernel(const double *__restrict__ coords, int64_t n, int dims, double *__restrict__ xy)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    xy[2 * i] = coords[i * dims];
    xy[2 * i + 1] = dims > 1 ? coords[i * dims + 1] : 0.0;
}
__global__ void __launch_bounds__(256)
gather_nd_kernel(const double *__restrict__ coords, const int32_t *__restrict__ order, int64_t n, int dims,
                 double3p *__restrict__ sorted)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int64_t o = order[p];
    double3p c;
    c.x = coords[o * dims];
    c.y = dims > 1 ? coords[o * dims + 1] : 0.0;
    c.z = dims > 2 ? coords[o * dims + 2] : 0.0;
    sorted[p] = c;
}
__device__ __forceinline__ double dist2_nd(const double3p &a, const double3p &b)
{
    const double dx = __dsub_rn(a.x, b.x), dy = __dsub_rn(a.y, b.y), dz = __dsub_rn(a.z, b.z);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));     // no FMA contraction
}

constexpr int kBruteTile = 256;
template <int KB>
__global__ void __launch_bounds__(128)
knn_brute_kernel(const double3p *__restrict__ pts, const int32_t *__restrict__ order, int64_t n, int k,
                 int32_t *__restrict__ knn, double *__restrict__ kth_dist)
{
    __shared__ double3p tile[kBruteTile];
    __shared__ int32_t tile_o[kBruteTile];
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = p < n;
    const double3p me = pts[live ? p : 0];
    double bd[KB];
    int bi[KB], bo[KB];                                   // candidate position, its original index (tie order)
    for (int i = 0; i < k; ++i) { bd[i] = INFINITY; bi[i] = -1; bo[i] = 0x7fffffff; }
    for (int64_t base = 0; base < n; base += kBruteTile) {
        __syncthreads();
        for (int t = threadIdx.x; t < kBruteTile; t += blockDim.x)
            if (base + t < n) { tile[t] = pts[base + t]; tile_o[t] = order[base + t]; }
        __syncthreads();
        if (!live) continue;
        const int m = (int)min((int64_t)kBruteTile, n - base);
        for (int t = 0; t < m; ++t) {
            const int64_t q = base + t;
            if (q == p) continue;
            const double d2 = dist2_nd(me, tile[t]);
            const int oq = tile_o[t];
            if (d2 > bd[k - 1] || (d2 == bd[k - 1] && oq > bo[k - 1])) continue;
            int i = k - 1;                                // insertion keeps (distance, original index) ascending
            while (i > 0 && (bd[i - 1] > d2 || (bd[i - 1] == d2 && bo[i - 1] > oq))) {
                bd[i] = bd[i - 1]; bi[i] = bi[i - 1]; bo[i] = bo[i - 1];
                --i;
            }
            bd[i] = d2; bi[i] = (int)q; bo[i] = oq;
        }
    }
    if (!live) return;
    for (int i = 0; i < k; ++i) knn[p * k + i] = bi[i];
    if (kth_dist) kth_dist[p] = sqrt(bd[k - 1]);
}

template <bool FILL>
__global__ void __launch_bounds__(128)
radius_brute_kernel(const double3p *__restrict__ pts, int64_t n, double r2, int32_t *__restrict__ deg,
                    const int32_t *__restrict__ indptr, int32_t *__restrict__ indices)
{
    __shared__ double3p tile[kBruteTile];
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = p < n;
    const double3p me = pts[live ? p : 0];
    int cnt = 0;
    const int base_out = (FILL && live) ? indptr[p] : 0;
    for (int64_t base = 0; base < n; base += kBruteTile) {
        __syncthreads();
        for (int t = threadIdx.x; t < kBruteTile; t += blockDim.x)
            if (base + t < n) tile[t] = pts[base + t];
        __syncthreads();
        if (!live) continue;
        const int m = (int)min((int64_t)kBruteTile, n - base);
        for (int t = 0; t < m; ++t) {
            if (base + t == p) continue;
            if (dist2_nd(me, tile[t]) <= r2) {
                if (FILL) indices[base_out + cnt] = (int32_t)(base + t);
                ++cnt;
            }
        }
    }
    if (!FILL && live) deg[p] = cnt;
}

// ---------------------------------------------------------------- order statistics on the device (radix select)
// Non-negative doubles order like their bit patterns.  Eight passes of 8 bits from the top: a histogram of the
// current byte over the keys that match the prefix found so far, then one thread walks the 256 bins.  No host sync.
struct SelectState {
    unsigned long long prefix;      // key bits decided so far (high bytes)
    long long rank;                 // rank of the wanted element among the keys matching the prefix
    unsigned hist[256];
};
__global__ void select_init_kernel(SelectState *st, long long rank)
{
    st->prefix = 0ull;
    st->rank = rank;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) st->hist[i] = 0u;
}
__global__ void __launch_bounds__(256)
select_hist_kernel(const double *__restrict__ v, int64_t n, int pass, SelectState *st)
{
    __shared__ unsigned sh[256];
    sh[threadIdx.x] = 0u;
    __syncthreads();
    const int shift = 56 - 8 * pass;
    const unsigned long long prefix = st->prefix;
    const unsigned long long mask = pass == 0 ? 0ull : ~0ull << (shift + 8);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = (unsigned long long)__double_as_longlong(v[i]);
        if ((key & mask) == prefix) atomicAdd(&sh[(key >> shift) & 255ull], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], sh[threadIdx.x]);
}
__global__ void select_pick_kernel(SelectState *st, int pass, double *out)
{
    if (threadIdx.x == 0) {
        long long r = st->rank;
        int b = 0;
        for (; b < 255; ++b) {
            if (r < (long long)st->hist[b]) break;
            r -= st->hist[b];
        }
        st->rank = r;
        st->prefix |= (unsigned long long)b << (56 - 8 * pass);
        if (pass == 7) *out = __longlong_as_double((long long)st->prefix);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) st->hist[i] = 0u;
}
// out[0] <- the element of rank `rank` (0-based, ascending) of v[0..n), all values >= 0
static int device_select(const double *v, int64_t n, long long rank, SelectState *st, double *out, cudaStream_t stream)
{
    select_init_kernel<<<1, 256, 0, stream>>>(st, rank);
    const int blocks = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)kNumSM * 8);
    for (int pass = 0; pass < 8; ++pass) {
        select_hist_kernel<<<blocks, 256, 0, stream>>>(v, n, pass, st);
        select_pick_kernel<<<1, 256, 0, stream>>>(st, pass, out);
    }
    count_launches(16);
    FDB_LAUNCH_CHECK("device_select");
    return FDB_OK;
}

// ---------------------------------------------------------------- symmetrise the directed kNN lists
__global__ void __launch_bounds__(256)
reverse_count_kernel(const int32_t *__restrict__ knn, int64_t n, int k, int32_t *__restrict__ extra)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * k) return;
    const int p = (int)(t / k);
    const int q = knn[t];
    if (q < 0) return;
    bool mutual = false;
    for (int i = 0; i < k; ++i) mutual |= (knn[(int64_t)q * k + i] == p);
    if (!mutual) atomicAdd(extra + q, 1);
}

__global__ void __launch_bounds__(256)
degree_kernel(const int32_t *__restrict__ knn, const int32_t *__restrict__ extra, int64_t n, int k,
              int32_t *__restrict__ deg)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int own = 0;
    for (int i = 0; i < k; ++i) own += knn[p * k + i] >= 0;
    deg[p] = own + extra[p];
}

__global__ void __launch_bounds__(256)
knn_fill_kernel(const int32_t *__restrict__ knn, int64_t n, int k, const int32_t *__restrict__ indptr,
                int32_t *__restrict__ cursor, int32_t *__restrict__ indices)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int own = 0;
    for (int i = 0; i < k; ++i) {
        const int q = knn[p * k + i];
        if (q >= 0) indices[indptr[p] + own++] = q;
    }
    for (int i = 0; i < k; ++i) {
        const int q = knn[p * k + i];
        if (q < 0) continue;
        bool mutual = false;
        int q_own = 0;
        for (int j = 0; j < k; ++j) {
            const int v = knn[(int64_t)q * k + j];
            mutual |= (v == (int)p);
            q_own += v >= 0;
        }
        if (!mutual) indices[indptr[q] + q_own + atomicAdd(cursor + q, 1)] = (int32_t)p;
    }
}

__global__ void __launch_bounds__(256)
row_sort_kernel(const int32_t *__restrict__ indptr, int64_t n, int32_t *__restrict__ indices)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int s = indptr[p], e = indptr[p + 1];
    for (int a = s + 1; a < e; ++a) {
        const int v = indices[a];
        int b = a - 1;
        while (b >= s && indices[b] > v) { indices[b + 1] = indices[b]; --b; }
        indices[b + 1] = v;
    }
}

// ---------------------------------------------------------------- radius graph (cells are >= radius wide)
template <bool FILL>
__global__ void __launch_bounds__(128)
radius_kernel(const double2 *__restrict__ xy, int64_t n, GridSpec g, const int32_t *__restrict__ tile_rank,
              const int32_t *__restrict__ cell_start, double r2, int32_t *__restrict__ deg,
              const int32_t *__restrict__ indptr, int32_t *__restrict__ indices)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const double2 me = xy[p];
    int cx, cy;
    cell_of_point(g, me.x, me.y, cx, cy);
    int cnt = 0;
    const int base = FILL ? indptr[p] : 0;
    for (int dy = -1; dy <= 1; ++dy) {
        const int yy = cy + dy;
        if (yy < 0 || yy >= g.gy) continue;
        for (int dx = -1; dx <= 1; ++dx) {
            const int xx = cx + dx;
            if (xx < 0 || xx >= g.gx) continue;
            const int id = cell_id(g, tile_rank, xx, yy);
            const int s = __ldg(cell_start + id), e = __ldg(cell_start + id + 1);
            for (int q = s; q < e; ++q) {
                if (q == p) continue;
                if (dist2(me, xy[q]) <= r2) {
                    if (FILL) indices[base + cnt] = q;
                    ++cnt;
                }
            }
        }
    }
    if (!FILL) deg[p] = cnt;
}

// ---------------------------------------------------------------- relabel to input order
__global__ void __launch_bounds__(256)
input_degree_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ rank, int64_t n,
                    int32_t *__restrict__ deg)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = rank[i];
    deg[i] = indptr[p + 1] - indptr[p];
}

__global__ void __launch_bounds__(256)
input_fill_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                  const int32_t *__restrict__ order, const int32_t *__restrict__ rank, int64_t n,
                  const int32_t *__restrict__ out_indptr, int32_t *__restrict__ out_indices)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = rank[i];
    const int s = indptr[p], e = indptr[p + 1];
    const int o = out_indptr[i];
    for (int a = 0; a < e - s; ++a) {
        const int v = order[indices[s + a]];
        int b = a - 1;
        while (b >= 0 && out_indices[o + b] > v) { out_indices[o + b + 1] = out_indices[o + b]; --b; }
        out_indices[o + b + 1] = v;
    }
}

// ---------------------------------------------------------------- host side
struct Bump {
    char *base;
    int64_t off = 0;
    explicit Bump(void *b) : base((char *)b) {}
    template <typename T> T *take(int64_t count)
    {
        off = round_up(off, 256);
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += count * (int64_t)sizeof(T);
        return p;
    }
};

struct GraphScratch {
    double *bbox_partial;
    int32_t *cell_of, *hist, *cursor, *tile_rank, *block_sums, *knn, *extra, *deg;
    double2 *xy;
    double *kth;
    double *xy_proj;                 // general path: first two coordinates of every spot (input order)
    double3p *pts;                   // general path: coordinates in tile order, padded to three
    SelectState *select;
    double *select_out;              // two order statistics
    int64_t bytes;
};

static int64_t cell_capacity(int64_t n) { return 4 * std::max<int64_t>(n, 64) + 1024; }

static GraphScratch carve(void *ws, int64_t n, int k)
{
    Bump b(ws);
    GraphScratch s;
    const int64_t cells = cell_capacity(n);
    s.bbox_partial = b.take<double>(4 * 1024);
    s.cell_of = b.take<int32_t>(n);
    s.hist = b.take<int32_t>(cells + 1);
    s.cursor = b.take<int32_t>(std::max(cells, n) + 1);
    s.tile_rank = b.take<int32_t>(cells);
    s.block_sums = b.take<int32_t>(scan_blocks(std::max(cells, n) + 1) + 1);
    s.knn = b.take<int32_t>(n * std::max(k, 1));
    s.extra = b.take<int32_t>(n + 1);
    s.deg = b.take<int32_t>(n + 1);
    s.xy = b.take<double2>(n);
    s.kth = b.take<double>(n);
    s.xy_proj = b.take<double>(2 * n);
    s.pts = b.take<double3p>(n);
    s.select = b.take<SelectState>(1);
    s.select_out = b.take<double>(2);
    s.bytes = round_up(b.off, 256);
    return s;
}

static uint32_t morton2(uint32_t x, uint32_t y)
{
    auto spread = [](uint32_t v) {
        uint64_t w = v;
        w = (w | (w << 16)) & 0x0000FFFF0000FFFFull;
        w = (w | (w << 8)) & 0x00FF00FF00FF00FFull;
        w = (w | (w << 4)) & 0x0F0F0F0F0F0F0F0Full;
        w = (w | (w << 2)) & 0x3333333333333333ull;
        w = (w | (w << 1)) & 0x5555555555555555ull;
        return w;
    };
    return (uint32_t)(spread(x) | (spread(y) << 1));
}

static int pow2_floor_cap8(int v)
{
    int p = 1;
    while (p * 2 <= v && p < 8) p *= 2;
    return p;
}

// choose the hash grid for a bounding box; min_cell > 0 forces cells at least that wide
static GridSpec make_grid(double lo_x, double lo_y, double hi_x, double hi_y, int64_t n, double min_cell,
                          std::vector<int32_t> &tile_rank)
{
    GridSpec g;
    const double ex = std::max(hi_x - lo_x, 0.0), ey = std::max(hi_y - lo_y, 0.0);
    const double big = std::max(ex, ey);
    double cell = std::sqrt(ex * ey * 4.0 / (double)std::max<int64_t>(n, 1));   // ~4 spots per cell
    cell = std::max(cell, big * 4.0 / (double)std::max<int64_t>(n, 1));
    if (!(cell > 0.0) || !std::isfinite(cell)) cell = 1.0;
    cell = std::max(cell, min_cell);
    const int64_t cap = std::max<int64_t>(n, 64);
    for (;;) {
        const double fx = std::floor(ex / cell) + 1.0, fy = std::floor(ey / cell) + 1.0;
        if (fx * fy <= (double)cap && fx < 1e9 && fy < 1e9) break;
        cell *= 1.25;
    }
    g.x0 = lo_x; g.y0 = lo_y; g.cell = cell; g.inv_cell = 1.0 / cell;
    g.gx = (int)std::floor(ex / cell) + 1;
    g.gy = (int)std::floor(ey / cell) + 1;
    g.tw = pow2_floor_cap8(g.gx);
    g.th = pow2_floor_cap8(g.gy);
    g.tiles_x = (g.gx + g.tw - 1) / g.tw;
    g.tiles_y = (g.gy + g.th - 1) / g.th;
    const int n_tiles = g.tiles_x * g.tiles_y;
    g.n_cells = n_tiles * g.tw * g.th;
    std::vector<std::pair<uint32_t, int32_t>> keyed(n_tiles);
    for (int ty = 0; ty < g.tiles_y; ++ty)
        for (int tx = 0; tx < g.tiles_x; ++tx)
            keyed[ty * g.tiles_x + tx] = {morton2((uint32_t)tx, (uint32_t)ty), ty * g.tiles_x + tx};
    if (g.tiles_x <= 65535 && g.tiles_y <= 65535) std::sort(keyed.begin(), keyed.end());
    tile_rank.assign(n_tiles, 0);
    for (int r = 0; r < n_tiles; ++r) tile_rank[keyed[r].second] = r;
    return g;
}

static int grid1d(int64_t n, int threads) { return (int)std::max<int64_t>(1, ceil_div(n, threads)); }

// bin spots into cells and publish order / rank / sorted coordinates / cell_start (in s.hist)
static int bin_spots(const double *coords, int64_t n, const GridSpec &g, const std::vector<int32_t> &tile_rank,
                     GraphScratch &s, int32_t *order, int32_t *rank, cudaStream_t st)
{
    FDB_CUDA(cudaMemcpyAsync(s.tile_rank, tile_rank.data(), tile_rank.size() * 4, cudaMemcpyHostToDevice, st));
    FDB_CUDA(cudaStreamSynchronize(st));          // tile_rank is a host temporary
    FDB_CUDA(cudaMemsetAsync(s.hist, 0, ((int64_t)g.n_cells + 1) * 4, st));
    FDB_CUDA(cudaMemsetAsync(s.cursor, 0, ((int64_t)g.n_cells + 1) * 4, st));
    cell_count_kernel<<<grid1d(n, 256), 256, 0, st>>>(coords, n, g, s.tile_rank, s.cell_of, s.hist);
    int rc = exclusive_scan(s.hist, g.n_cells, s.hist, s.block_sums, st);
    if (rc) return rc;
    cell_scatter_kernel<<<grid1d(n, 256), 256, 0, st>>>(s.cell_of, n, s.hist, s.cursor, order);
    cell_finish_kernel<<<grid1d(g.n_cells, 256), 256, 0, st>>>(coords, g.n_cells, s.hist, order, rank, s.xy);
    count_launches(2);
    FDB_LAUNCH_CHECK("bin_spots");
    return FDB_OK;
}

template <int KMAX>
static void launch_knn(const GraphScratch &s, const int32_t *order, int64_t n, const GridSpec &g, int k,
                       double *kth, cudaStream_t st)
{
    knn_kernel<KMAX><<<grid1d(n, 128), 128, 0, st>>>(s.xy, order, n, g, s.tile_rank, s.hist, k, s.knn, kth);
}

static int run_knn(const GraphScratch &s, const int32_t *order, int64_t n, const GridSpec &g, int k,
                   double *kth, cudaStream_t st, bool general)
{
    if (general || k > 32) {                    // exhaustive search: any dimension, any k <= 1024
        if (k <= 32) knn_brute_kernel<32><<<grid1d(n, 128), 128, 0, st>>>(s.pts, order, n, k, s.knn, kth);
        else if (k <= 128) knn_brute_kernel<128><<<grid1d(n, 128), 128, 0, st>>>(s.pts, order, n, k, s.knn, kth);
        else if (k <= 1024) knn_brute_kernel<1024><<<grid1d(n, 128), 128, 0, st>>>(s.pts, order, n, k, s.knn, kth);
        else {
            set_error("k_neighbors > 1024 is not supported (got %d)", k);
            return FDB_ERR_UNSUPPORTED;
        }
        FDB_LAUNCH_CHECK("knn_brute_kernel");
        return FDB_OK;
    }
    if (k <= 8) launch_knn<8>(s, order, n, g, k, kth, st);
    else if (k <= 16) launch_knn<16>(s, order, n, g, k, kth, st);
    else launch_knn<32>(s, order, n, g, k, kth, st);
    FDB_LAUNCH_CHECK("knn_kernel");
    return FDB_OK;
}

}  // namespace fdb

using namespace fdb;

extern "C" __attribute__((visibility("default"))) int64_t fdb_graph_workspace_bytes(int64_t n_spots, int32_t k)
{
    if (n_spots < 0 || k < 0) return -1;
    return carve(nullptr, n_spots, k).bytes;
}

extern "C" __attribute__((visibility("default"))) int fdb_graph_build_nd(const double *coords_nd, int64_t n, int32_t dims, int32_t mode, int32_t k,
                               double radius, int32_t *order, int32_t *rank, int32_t *indptr, int32_t *indices,
                               int64_t indices_capacity, int64_t *host_nnz, double *host_radius,
                               void *workspace, int64_t workspace_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    FDB_REQUIRE(n >= 0 && n < (int64_t)1 << 31, "n_spots out of range");
    FDB_REQUIRE(dims >= 1 && dims <= 3, "coords must have 1, 2 or 3 columns, got %d", dims);
    FDB_REQUIRE(mode >= 0 && mode <= 2, "unknown graph mode %d", mode);
    FDB_REQUIRE(host_nnz != nullptr, "host_nnz is required");
    FDB_REQUIRE(mode != 0 || k >= 0, "k_neighbors must be non-negative, got %d", k);
    FDB_REQUIRE(mode != 1 || radius > 0, "radius must be positive, got %g", radius);
    *host_nnz = 0;
    if (host_radius) *host_radius = radius;
    if (n == 0) return FDB_OK;
    const int k_eff = mode == 0 ? (int)std::min<int64_t>(k, n - 1) : 1;
    GraphScratch s = carve(workspace, n, std::max(k_eff, 1));
    if (workspace_bytes < s.bytes || !workspace) {
        set_error("graph workspace too small: need %lld bytes, got %lld", (long long)s.bytes, (long long)workspace_bytes);
        return FDB_ERR_WORKSPACE;
    }
    FDB_REQUIRE(coords_nd && order && rank && indptr, "null pointer");
    // general path (exhaustive search in tile order): anything but 2-D coordinates with k <= 32
    const bool general = dims != 2 || (mode == 0 && k_eff > 32);
    const double *coords = coords_nd;
    if (dims != 2) {                             // tile order from the first two coordinates
        project_xy_kernel<<<grid1d(n, 256), 256, 0, st>>>(coords_nd, n, dims, s.xy_proj);
        FDB_LAUNCH_CHECK("project_xy_kernel");
        coords = s.xy_proj;
    }

    // bounding box (one host sync)
    const int bb_blocks = (int)std::min<int64_t>(1024, ceil_div(n, 256));
    bbox_kernel<<<bb_blocks, 256, 0, st>>>(coords, n, s.bbox_partial);
    FDB_LAUNCH_CHECK("bbox_kernel");
    std::vector<double> part(4 * bb_blocks);
    FDB_CUDA(cudaMemcpyAsync(part.data(), s.bbox_partial, part.size() * 8, cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    double lo_x = INFINITY, lo_y = INFINITY, hi_x = -INFINITY, hi_y = -INFINITY;
    for (int b = 0; b < bb_blocks; ++b) {
        lo_x = std::min(lo_x, part[4 * b]); lo_y = std::min(lo_y, part[4 * b + 1]);
        hi_x = std::max(hi_x, part[4 * b + 2]); hi_y = std::max(hi_y, part[4 * b + 3]);
    }
    FDB_REQUIRE(std::isfinite(lo_x) && std::isfinite(lo_y) && std::isfinite(hi_x) && std::isfinite(hi_y),
                "coords contain non-finite values");

    std::vector<int32_t> tile_rank;
    GridSpec g = make_grid(lo_x, lo_y, hi_x, hi_y, n, (mode == 1 && !general) ? radius : 0.0, tile_rank);
    int rc = bin_spots(coords, n, g, tile_rank, s, order, rank, st);
    if (rc) return rc;
    if (general) {
        gather_nd_kernel<<<grid1d(n, 256), 256, 0, st>>>(coords_nd, order, n, dims, s.pts);
        FDB_LAUNCH_CHECK("gather_nd_kernel");
    }

    if (mode == 2) {
        // radius = 1.5 * median nearest-neighbour distance (utils/graph.py:163-170)
        if (n <= 1) {
            FDB_CUDA(cudaMemsetAsync(indptr, 0, (n + 1) * 4, st));
            return FDB_OK;
        }
        rc = run_knn(s, order, n, g, 1, s.kth, st, general);
        if (rc) return rc;
        // numpy median on the device: the middle order statistic, or the mean of the two middle ones for even n
        const int64_t mid = n / 2;
        rc = device_select(s.kth, n, mid, s.select, s.select_out, st);
        if (rc) return rc;
        if (n % 2 == 0) {
            rc = device_select(s.kth, n, mid - 1, s.select, s.select_out + 1, st);
            if (rc) return rc;
        }
        double med2[2] = {0.0, 0.0};
        FDB_CUDA(cudaMemcpyAsync(med2, s.select_out, 16, cudaMemcpyDeviceToHost, st));
        FDB_CUDA(cudaStreamSynchronize(st));
        const double med = n % 2 == 0 ? (med2[1] + med2[0]) / 2.0 : med2[0];
        radius = med * 1.5;
        if (host_radius) *host_radius = radius;
        if (!general && radius > g.cell) {       // re-bin with cells at least `radius` wide
            g = make_grid(lo_x, lo_y, hi_x, hi_y, n, radius, tile_rank);
            rc = bin_spots(coords, n, g, tile_rank, s, order, rank, st);
            if (rc) return rc;
        }
    }

    const double r2 = radius * radius;
    if (mode == 0) {
        if (k_eff <= 0) {
            FDB_CUDA(cudaMemsetAsync(indptr, 0, (n + 1) * 4, st));
            return FDB_OK;
        }
        rc = run_knn(s, order, n, g, k_eff, nullptr, st, general);
        if (rc) return rc;
        FDB_CUDA(cudaMemsetAsync(s.extra, 0, (n + 1) * 4, st));
        reverse_count_kernel<<<grid1d(n * k_eff, 256), 256, 0, st>>>(s.knn, n, k_eff, s.extra);
        degree_kernel<<<grid1d(n, 256), 256, 0, st>>>(s.knn, s.extra, n, k_eff, s.deg);
        count_launches(1);
        FDB_LAUNCH_CHECK("degree_kernel");
    } else if (general) {
        radius_brute_kernel<false><<<grid1d(n, 128), 128, 0, st>>>(s.pts, n, r2, s.deg, nullptr, nullptr);
        FDB_LAUNCH_CHECK("radius_brute_kernel<count>");
    } else {
        radius_kernel<false><<<grid1d(n, 128), 128, 0, st>>>(s.xy, n, g, s.tile_rank, s.hist, r2, s.deg, nullptr, nullptr);
        FDB_LAUNCH_CHECK("radius_kernel<count>");
    }
    rc = exclusive_scan(s.deg, n, indptr, s.block_sums, st);
    if (rc) return rc;
    int32_t nnz32 = 0;
    FDB_CUDA(cudaMemcpyAsync(&nnz32, indptr + n, 4, cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    *host_nnz = nnz32;
    if (nnz32 > indices_capacity || (nnz32 > 0 && !indices)) {
        set_error("indices_capacity %lld too small for %d stored entries", (long long)indices_capacity, nnz32);
        return FDB_ERR_WORKSPACE;
    }
    if (mode == 0) {
        FDB_CUDA(cudaMemsetAsync(s.cursor, 0, (n + 1) * 4, st));
        knn_fill_kernel<<<grid1d(n, 256), 256, 0, st>>>(s.knn, n, k_eff, indptr, s.cursor, indices);
    } else if (general) {
        radius_brute_kernel<true><<<grid1d(n, 128), 128, 0, st>>>(s.pts, n, r2, nullptr, indptr, indices);
    } else {
        radius_kernel<true><<<grid1d(n, 128), 128, 0, st>>>(s.xy, n, g, s.tile_rank, s.hist, r2, nullptr, indptr, indices);
    }
    row_sort_kernel<<<grid1d(n, 256), 256, 0, st>>>(indptr, n, indices);
    count_launches(1);
    FDB_LAUNCH_CHECK("graph fill");
    return FDB_OK;
}

extern "C" __attribute__((visibility("default"))) int fdb_graph_build(const double *coords, int64_t n, int32_t mode, int32_t k, double radius,
                               int32_t *order, int32_t *rank, int32_t *indptr, int32_t *indices,
                               int64_t indices_capacity, int64_t *host_nnz, double *host_radius,
                               void *workspace, int64_t workspace_bytes, void *stream)
{
    return fdb_graph_build_nd(coords, n, 2, mode, k, radius, order, rank, indptr, indices, indices_capacity, host_nnz,
                              host_radius, workspace, workspace_bytes, stream);
}

extern "C" __attribute__((visibility("default"))) int fdb_graph_to_input_order(const int32_t *indptr, const int32_t *indices, const int32_t *order,
                                        const int32_t *rank, int64_t n, int32_t *out_indptr,
                                        int32_t *out_indices, void *workspace, int64_t workspace_bytes,
                                        void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    FDB_REQUIRE(n >= 0, "negative n_spots");
    if (n == 0) {
        if (out_indptr) FDB_CUDA(cudaMemsetAsync(out_indptr, 0, 4, st));
        return FDB_OK;
    }
    Bump b(workspace);
    int32_t *deg = b.take<int32_t>(n + 1);
    int32_t *block_sums = b.take<int32_t>(scan_blocks(n + 1) + 1);
    if (!workspace || workspace_bytes < round_up(b.off, 256)) {
        set_error("workspace too small: need %lld bytes", (long long)round_up(b.off, 256));
        return FDB_ERR_WORKSPACE;
    }
    input_degree_kernel<<<grid1d(n, 256), 256, 0, st>>>(indptr, rank, n, deg);
    int rc = exclusive_scan(deg, n, out_indptr, block_sums, st);
    if (rc) return rc;
    input_fill_kernel<<<grid1d(n, 256), 256, 0, st>>>(indptr, indices, order, rank, n, out_indptr, out_indices);
    count_launches(1);
    FDB_LAUNCH_CHECK("input_fill_kernel");
    return FDB_OK;
}
