// Kernel family 4: graph-regularised non-negative block-coordinate descent (Jacobi over
// spots, Gauss-Seidel over cell types inside a spot), objective and proportion output.
//
// Reference semantics (upstream file:line):
//   core/solver.py:149-184  per spot: neighbour sum over beta_in (Jacobi), cyclic CD, max-norm stats
//   core/solver.py:72-99    r = G b; for k: part = Xty_k - r_k + G_kk b_k (+ lam * nsum_k if deg > 0);
//                           b_k <- max(0, soft(part, rho) / (G_kk + lam*deg)) (0 if denom <= 1e-10);
//                           r += (b_k_new - b_k) * G[:, k] when the step is non-zero
//   core/solver.py:395-413  rel = max|delta| / (max|old| + 1e-10); stop when rel < tol
//   core/solver.py:269-284  objective;  core/solver.py:445-452 normalisation
//
// What lives where (HBM-bound: (12*Kp + 4*deg + 4) bytes per spot per sweep):
//   bcd_p.cuh / bcd_p_inst.cu   the PRODUCTION sweep kernel bcd_sweep_p_kernel (persistent, software-pipelined over
//                               128-spot patches, fp16 gather tile, pair-step descent), one translation unit per Kp
//   this file                   bcd_sweep_kernel: fp32-gather sweep, one patch per CTA -- the fallback when the spatial
//                               coupling is strong (fp16 neighbour values inadmissible) or no gather plan is given;
//                               bcd_plan_kernel: per-patch halo lists and neighbour codes, built once per graph;
//                               dispatcher, objective terms (float64 accumulation), proportions, C entry points.
//   Common to all sweep kernels: thread per spot for the strictly sequential K-step coordinate descent, evaluated in
//   the direct form part_k = c_k - sum_{j!=k} G_kj b_j (K^2 FMAs per spot, the dense minimum; the reference's
//   maintained-residual form costs 1.5-2 K^2 under SIMT) on packed fma.rn.f32x2 with the negated Gram matrix as a
//   by-value kernel parameter (constant bank); coalesced row traffic staged through shared memory; convergence
//   statistics by redux.sync max per warp, one atomicMax per CTA, last CTA finalises the stop test.
#include "bcd_common.cuh"

extern "C" __attribute__((visibility("default"))) int fdb_bcd_init(float *beta, int64_t n_rows, int32_t n_types, void *state, void *stream);

namespace fdb {

// any-K form (wide.cu)
int finish_wide(const float *beta, const int32_t *order, int64_t n_rows, int n_types, double *beta_out, double *prop_out,
                cudaStream_t st);

// Sweep kernel, tile-cached fp32-gather form (fallback: strong coupling / no plan).  One CTA = NW warps = TILE = 32*NW consecutive spots
// (tile order => a compact patch of the tissue).
//   step 1  the CTA streams its TILE beta_old rows and H rows into shared memory with fully
//           independent coalesced 128-bit loads (one global round trip, 2*Kp/4 loads in flight per thread);
//   step 2  neighbour sums, one spot per lane: a neighbour inside the CTA's patch (75-85 % of them) is
//           read from the shared tile, the rest from global/L2, through the same generic 128-bit loads;
//           c = H + lam * sum is left in shared memory;
//   step 3  one spot per lane: cyclic coordinate descent in the direct form on packed FFMA2;
//   step 4  the warp streams its 32 new rows back out.
template <int KP, int NW, int UNR>
__global__ void __launch_bounds__(NW * 32, (KP <= 32 ? (UNR == 1 ? 768 / (NW * 32) : 512 / (NW * 32)) : 1))
bcd_sweep_kernel(const float *__restrict__ h, const __grid_constant__ GramArg<KP> G,
                 const float *__restrict__ beta_in, float *__restrict__ beta_out,
                 const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                 int n_rows, int n_types, float lam, float rho, float tol, int finalize, SolveState *state)
{
    // the flag load overlaps the tile loads below; nothing is written to global memory before it is tested
    const int already_converged = *reinterpret_cast<volatile int *>(&state->converged);

    using L = TileLayout<KP>;
    constexpr int Q = L::Q, S = L::S, TILE = NW * 32;
    constexpr int SLOTS = 32 / Q;                       // rows walked per warp instruction
    constexpr int ITERS = (32 + SLOTS - 1) / SLOTS;
    extern __shared__ __align__(16) float sweep_smem[];
    float *c_tile = sweep_smem;
    float *b_tile = sweep_smem + TILE * S;
    int *idx_tile = reinterpret_cast<int *>(sweep_smem + 2 * TILE * S);
    __shared__ unsigned red[2][NW];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile_base = blockIdx.x * TILE;
    const int wrow = warp * 32;                          // first tile row of this warp
    int *iw = idx_tile + warp * kIdxCap;

    // ---------------- step 0/1: row pointers, then the tile's rows and the warp's index slice
    const int my_row = tile_base + wrow + lane;
    int my_s = 0, my_e = 0;
    if (my_row < n_rows) { my_s = __ldg(indptr + my_row); my_e = __ldg(indptr + my_row + 1); }
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const int idx = threadIdx.x + i * TILE;
        const int row = idx / Q, q = idx - row * Q;
        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f), hh = bb;
        if (tile_base + row < n_rows) {
            bb = ld4(beta_in + (size_t)(tile_base + row) * KP + 4 * q);
            hh = __ldcs(reinterpret_cast<const float4 *>(h + (size_t)(tile_base + row) * KP + 4 * q));
        }
        st4(b_tile + L::at(row, q), bb);
        st4(c_tile + L::at(row, q), hh);
    }
    const int my_deg = my_e - my_s;
    const int ibase = __shfl_sync(kFull, my_s, 0);
    const int icnt = __reduce_max_sync(kFull, my_e - ibase);
    const bool staged = icnt <= kIdxCap;
    if (staged)
        for (int t = lane; t < icnt; t += 32) iw[t] = __ldg(indices + ibase + t);
    __syncthreads();

    if (already_converged) return;                        // uniform across the grid

    // ---------------- step 2: c += lam * sum_j beta_old[j], one spot per lane.
    // A neighbour inside the CTA's patch is read from the staged tile, any other from global/L2; the
    // two cases share one instruction stream through generic 128-bit loads.  Lanes whose list is
    // shorter than the warp's longest re-read their own staged row with weight 0 (no divergence).
    {
        const int trow = wrow + lane;
        float ns[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) ns[k] = 0.f;
        const int maxdeg = __reduce_max_sync(kFull, my_deg);
        const int rs = my_s - ibase;
#pragma unroll UNR
        for (int u = 0; u < maxdeg; ++u) {
            const bool has = u < my_deg;
            int rel = trow;
            if (has) rel = (staged ? iw[rs + u] : __ldg(indices + my_s + u)) - tile_base;
            const bool in = (unsigned)rel < (unsigned)TILE;
            const float *base = in ? (b_tile + rel * S) : (beta_in + (size_t)(rel + tile_base) * KP);
            const int sw = (L::SWZ && in) ? (rel & 7) : 0;
            const float m = has ? 1.f : 0.f;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const float4 v = *reinterpret_cast<const float4 *>(base + 4 * (q ^ sw));
                ns[4 * q] = fmaf(v.x, m, ns[4 * q]);
                ns[4 * q + 1] = fmaf(v.y, m, ns[4 * q + 1]);
                ns[4 * q + 2] = fmaf(v.z, m, ns[4 * q + 2]);
                ns[4 * q + 3] = fmaf(v.w, m, ns[4 * q + 3]);
            }
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            float4 cc = ld4(c_tile + L::at(trow, q));
            cc.x = fmaf(lam, ns[4 * q], cc.x); cc.y = fmaf(lam, ns[4 * q + 1], cc.y);
            cc.z = fmaf(lam, ns[4 * q + 2], cc.z); cc.w = fmaf(lam, ns[4 * q + 3], cc.w);
            st4(c_tile + L::at(trow, q), cc);
        }
    }
    __syncwarp();

    // ---------------- step 3: one spot per lane, cyclic coordinate descent in the direct form
    //   part_k = c_k - sum_{j != k} G_kj b_j   (b_j already updated for j < k)
    // G.g holds the NEGATED Gram with a zero diagonal, so step k is Kp/2 packed FFMA2 (fma.rn.f32x2, two
    // fp32 FMAs per issue slot on sm_100) whose G operand pair comes from the constant bank via LDCU.128.
    float dmax = 0.f, amax = 0.f;
    {
        const int trow = wrow + lane;
        const float lam_deg = lam * (float)my_deg;
        float2 b2[KP / 2];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const float4 b4 = ld4(b_tile + L::at(trow, q));
            b2[2 * q] = make_float2(b4.x, b4.y);
            b2[2 * q + 1] = make_float2(b4.z, b4.w);
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const float4 c4 = ld4(c_tile + L::at(trow, q));
            float4 n4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = 4 * q + j;
                if (k >= KP - 3 && k >= n_types) continue;          // padding columns stay zero (warp-uniform)
                float2 a0 = make_float2(elem(c4, j), 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
                for (int jj = 0; jj < KP / 2; ++jj) {
                    const float2 g = make_float2(G.g[k * KP + 2 * jj], G.g[k * KP + 2 * jj + 1]);
                    if (jj & 1) ffma2(a1, g, b2[jj]); else ffma2(a0, g, b2[jj]);
                }
                const float part = (a0.x + a0.y) + (a1.x + a1.y);
                const float old = (k & 1) ? b2[k / 2].y : b2[k / 2].x;
                const float den = G.diag[k] + lam_deg;
                // max(0, soft(part, rho) / den) == max(0, (part - rho) / den) for den > 0
                const float nv = den > 1e-10f ? fmaxf(0.f, __fdividef(part - rho, den)) : 0.f;
                dmax = fmaxf(dmax, fabsf(nv - old));
                amax = fmaxf(amax, fabsf(old));
                if (k & 1) b2[k / 2].y = nv; else b2[k / 2].x = nv;
                set_elem(n4, j, nv);
            }
            st4(c_tile + L::at(trow, q), n4);                        // beta_new replaces c; beta_old stays readable
        }
        if (my_row >= n_rows) { dmax = 0.f; amax = 0.f; }
    }
    __syncwarp();

    // ---------------- step 4: stream the warp's new rows out
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const int idx = lane + 32 * i;
        const int lr = idx / Q, q = idx - lr * Q;
        const int p = tile_base + wrow + lr;
        if (p < n_rows) st4(beta_out + (size_t)p * KP + 4 * q, ld4(c_tile + L::at(wrow + lr, q)));
    }

    // ---------------- convergence statistics
    const unsigned wd = __reduce_max_sync(kFull, __float_as_uint(dmax));
    const unsigned wa = __reduce_max_sync(kFull, __float_as_uint(amax));
    if (lane == 0) { red[0][warp] = wd; red[1][warp] = wa; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned bd = 0u, ba = 0u;
#pragma unroll
        for (int w = 0; w < NW; ++w) { bd = max(bd, red[0][w]); ba = max(ba, red[1][w]); }
        if (bd > *reinterpret_cast<volatile unsigned *>(&state->max_diff_bits)) atomicMax(&state->max_diff_bits, bd);
        if (ba > *reinterpret_cast<volatile unsigned *>(&state->max_abs_bits)) atomicMax(&state->max_abs_bits, ba);
        if (finalize) {
            __threadfence();
            if (atomicAdd(&state->arrived, 1u) == gridDim.x - 1) {
                __threadfence();
                finalize_state(state, tol);
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// Gather plan for the production sweep kernel (bcd_p.cuh): the CTA's TILE beta_old rows are staged twice -- fp32 (own
// rows, needed exactly by the descent) and as an fp16 GATHER tile; neighbours outside the patch get a slot in a halo
// extension of the gather tile and are fetched ONCE per CTA with coalesced loads, so the gather itself is a pure
// shared-memory operation.  Why fp16 is admissible for the neighbour sum only: it enters the update as lam*sum with
// lam*deg ~ 0.5 % of G_kk (core/spatial.py:181-190), so a 5e-4 relative rounding of a neighbour value moves beta by
// ~1e-7 relative, three orders below the 1e-4 parity bar; the spot's own row, H and the Gram products stay fp32, and
// the tile is range-scaled per sweep (bcd_p.cuh) so magnitude never matters.
// ------------------------------------------------------------------------------------
template <int TILE>
__global__ void __launch_bounds__(TILE)
bcd_plan_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices, int n_rows,
                int32_t *__restrict__ halo_cnt, int32_t *__restrict__ halo_rows, uint16_t *__restrict__ codes,
                uint8_t *__restrict__ codes8)
{
    constexpr int HCAP = TILE;
    static_assert(TILE == 128, "byte codes assume 128-spot patches");
    __shared__ int keys[kPlanHash];
    __shared__ int slots[kPlanHash];
    __shared__ int count;
    for (int i = threadIdx.x; i < kPlanHash; i += TILE) keys[i] = -1;
    if (threadIdx.x == 0) count = 0;
    __syncthreads();
    const int tile_base = blockIdx.x * TILE;
    const int row = tile_base + threadIdx.x;
    int s = 0, e = 0;
    if (row < n_rows) { s = indptr[row]; e = indptr[row + 1]; }
    auto probe = [&](int g, bool insert) -> int {            // returns the table position of g, or -1
        unsigned hpos = ((unsigned)g * 2654435761u) >> 22;    // 10 bits
        for (int t = 0; t < kPlanHash; ++t) {
            const int cur = insert ? atomicCAS(&keys[hpos], -1, g) : keys[hpos];
            if (cur == g || (insert && cur == -1)) return (int)hpos;
            if (!insert && cur == -1) return -1;
            hpos = (hpos + 1) & (kPlanHash - 1);
        }
        return -1;
    };
    for (int j = s; j < e; ++j) {
        const int g = indices[j];
        if ((unsigned)(g - tile_base) >= (unsigned)TILE) probe(g, true);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kPlanHash; i += TILE) {
        if (keys[i] >= 0) {
            const int slot = atomicAdd(&count, 1);
            slots[i] = slot;
            if (slot < kHaloSlots) halo_rows[(int64_t)blockIdx.x * HCAP + slot] = keys[i];
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int maxdeg = __reduce_max_sync(kFull, e - s);
    uint8_t *c8 = codes8 + ((int64_t)blockIdx.x * (TILE / 32) + (threadIdx.x >> 5)) * (kCodeRounds * 32) + lane;
    const bool transposed = maxdeg <= kCodeRounds;
    for (int j = s; j < e; ++j) {
        const int g = indices[j];
        const int rel = g - tile_base;
        unsigned short code;
        if ((unsigned)rel < (unsigned)TILE) {
            code = (unsigned short)rel;
        } else {
            const int pos = probe(g, false);
            code = (pos >= 0 && slots[pos] < kHaloSlots) ? (unsigned short)(TILE + slots[pos]) : kCodeSlow;
        }
        codes[j] = code;
        if (transposed) c8[(j - s) * 32] = code == kCodeSlow ? (uint8_t)kCodeSlow8 : (uint8_t)code;
    }
    if (transposed)
        for (int u = e - s; u < maxdeg; ++u) c8[u * 32] = (uint8_t)kCodeZero8;
    if (threadIdx.x == 0) halo_cnt[blockIdx.x] = min(count, kHaloSlots);
    // unused slots carry -1 so that the sweep kernel can fetch the list without knowing its length first
    if ((int)threadIdx.x >= min(count, kHaloSlots)) halo_rows[(int64_t)blockIdx.x * HCAP + threadIdx.x] = -1;
}

__global__ void bcd_finalize_kernel(SolveState *state, float tol)
{
    if (state->converged) return;
    finalize_state(state, tol);
}

__global__ void __launch_bounds__(256)
bcd_init_kernel(float *__restrict__ beta, int64_t n_rows, int kp, int n_types, SolveState *state)
{
    // one 16-byte store per thread (kp is a multiple of 8, so a group of four never straddles two rows)
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && state) {
        SolveState z = {};
        z.last_max_abs = n_types > 0 ? 1.0f / (float)n_types : 0.f;      // bound on |beta| for the first sweep (fp16 tile scale)
        z.ov_new[0] = __float_as_uint(z.last_max_abs);                   // the same bound, overlapped multi-GPU mode
        *state = z;
    }
    const int q = kp >> 2;                                               // groups per row
    if (i >= n_rows * q) return;
    const int k = 4 * (int)(i % q);
    const float v = 1.0f / (float)n_types;
    st4(beta + 4 * i, make_float4(k < n_types ? v : 0.f, k + 1 < n_types ? v : 0.f, k + 2 < n_types ? v : 0.f,
                                  k + 3 < n_types ? v : 0.f));
}

// production sweep kernel: defined in bcd_p.cuh, instantiated per row width in bcd_p_inst.cu
template <int KP, bool COMM>
int launch_sweep_p(const float *h, const GramArg<KP> &G, int n_types, const float *beta_in, float *beta_out,
                   const int32_t *indptr, const int32_t *indices, int64_t n_rows, float lam, float rho, float tol,
                   int finalize, SolveState *state, const void *plan, cudaStream_t st, const SweepComm *comm);

// fp16 neighbour values are admissible while the spatial term is a small part of the diagonal
// (auto lambda: lam*deg = 0.5 % of G_kk); for strongly coupled problems the sweep stays in fp32
static bool weak_coupling(const float *host_gram, int n_types, float lam)
{
    float mean_diag = 0.f;
    for (int k = 0; k < n_types; ++k) mean_diag += host_gram[k * n_types + k];
    mean_diag /= (float)n_types;
    return lam * 8.f <= 0.02f * mean_diag;
}
static int sweep_variant()
{
    static const int variant = getenv("FDB_SWEEP_VARIANT") ? atoi(getenv("FDB_SWEEP_VARIANT")) : 0;
    return variant;
}

template <int KP>
static int launch_sweep(const float *h, const float *host_gram, int n_types, const float *beta_in,
                        float *beta_out, const int32_t *indptr, const int32_t *indices, int64_t n_rows,
                        float lam, float rho, float tol, int finalize, SolveState *state, const void *plan,
                        cudaStream_t st, const SweepComm *comm)
{
    GramArg<KP> G;
    for (int i = 0; i < KP * KP; ++i) G.g[i] = 0.f;
    for (int k = 0; k < KP; ++k) G.diag[k] = k < n_types ? host_gram[k * n_types + k] : 0.f;
    for (int k = 0; k < n_types; ++k)
        for (int a = 0; a < n_types; ++a) G.g[k * KP + a] = (a == k) ? 0.f : -host_gram[k * n_types + a];
    // FDB_SWEEP_VARIANT=5 forces the fp32-gather kernel (tests compare the two)
    const int variant = sweep_variant();
    auto go = [&](auto kern, int nw) -> int {
        const size_t smem = (size_t)nw * 32 * 2 * TileLayout<KP>::S * 4 + (size_t)nw * kIdxCap * 4;
        FDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = (int)ceil_div(n_rows, nw * 32);
        kern<<<grid, nw * 32, smem, st>>>(h, G, beta_in, beta_out, indptr, indices, (int)n_rows, n_types, lam, rho,
                                          tol, finalize, state);
        FDB_LAUNCH_CHECK("bcd_sweep_kernel");
        return FDB_OK;
    };
    if constexpr (KP % 8 == 0) {
        if (plan != nullptr && variant == 0 && weak_coupling(host_gram, n_types, lam))      // production kernel
            return comm != nullptr
                       ? launch_sweep_p<KP, true>(h, G, n_types, beta_in, beta_out, indptr, indices, n_rows, lam, rho, tol,
                                                  finalize, state, plan, st, comm)
                       : launch_sweep_p<KP, false>(h, G, n_types, beta_in, beta_out, indptr, indices, n_rows, lam, rho, tol,
                                                   finalize, state, plan, st, comm);
    }
    if (comm != nullptr) {
        set_error("the fused multi-GPU sweep needs the gather plan and weak coupling");
        return FDB_ERR_UNSUPPORTED;
    }
    if constexpr (KP <= 32) {
        if (variant == 1) return go(bcd_sweep_kernel<KP, 4, 1>, 4);
        return go(bcd_sweep_kernel<KP, 8, 1>, 8);        // variant 5 (or Kp % 8 != 0): fp32 gather, no halo staging
    } else {
        return go(bcd_sweep_kernel<KP, 4, 1>, 4);
    }
}

static int dispatch_sweep(const float *h, const float *host_gram, int n_types, const float *beta_in,
                          float *beta_out, const int32_t *indptr, const int32_t *indices, int64_t n_rows,
                          float lam, float rho, float tol, int finalize, SolveState *state, const void *plan,
                          cudaStream_t st, const SweepComm *comm = nullptr)
{
#define FDB_SWEEP_CASE(KP_)                                                                            \
    case KP_:                                                                                          \
        return launch_sweep<KP_>(h, host_gram, n_types, beta_in, beta_out, indptr, indices, n_rows,    \
                                 lam, rho, tol, finalize, state, plan, st, comm);
    switch (fdb_padded_types(n_types)) {
#ifdef FDB_DEV_ONLY_KP                                     // development builds: one row width only
        FDB_SWEEP_CASE(FDB_DEV_ONLY_KP)
#else
        FDB_SWEEP_CASE(8) FDB_SWEEP_CASE(16) FDB_SWEEP_CASE(24) FDB_SWEEP_CASE(32)
        FDB_SWEEP_CASE(40) FDB_SWEEP_CASE(48) FDB_SWEEP_CASE(56) FDB_SWEEP_CASE(64)
#endif
    default:
        set_error("n_types must be in [1, %d] for the register-resident sweep kernels, got %d (use fdb_bcd_solve_wide)", FDB_MAX_TYPES,
                  n_types);
        return FDB_ERR_UNSUPPORTED;
    }
#undef FDB_SWEEP_CASE
}

// ------------------------------------------------------------------------------------
// objective terms (float64 accumulation), warp per spot
// ------------------------------------------------------------------------------------
struct GramFull {
    float g[FDB_MAX_TYPES * FDB_MAX_TYPES];               // kp x kp row-major in the first kp*kp entries
};

template <int NK>
__global__ void __launch_bounds__(256)
objective_kernel(const float *__restrict__ beta, const float *__restrict__ h, const float *__restrict__ ysq,
                 const __grid_constant__ GramFull gram, const int32_t *__restrict__ indptr,
                 const int32_t *__restrict__ indices, int64_t n_rows, int kp, double *__restrict__ out)
{
    extern __shared__ float gs[];                          // kp x kp
    for (int i = threadIdx.x; i < kp * kp; i += blockDim.x) gs[i] = gram.g[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double cross = 0.0, quad = 0.0, lap = 0.0, l1 = 0.0, yy = 0.0;
    const bool on0 = lane < kp, on1 = NK == 2 && lane + 32 < kp;
    // the chain row pointers -> neighbour ids -> neighbour rows is three dependent global round trips per spot; the
    // pointers and the first 8 neighbour ids of the NEXT spot of this warp are fetched one iteration ahead
    int s = 0, e = 0, nb0 = -1;                            // lane u < 8 holds neighbour u of the current spot
    if (warp_global < n_rows) {
        s = __ldg(indptr + warp_global); e = __ldg(indptr + warp_global + 1);
        if (lane < 8 && s + lane < e) nb0 = __ldg(indices + s + lane);
    }
    for (int64_t p = warp_global; p < n_rows; p += n_warps) {
        const int64_t pn = p + n_warps;
        int s2 = 0, e2 = 0;
        if (pn < n_rows) { s2 = __ldg(indptr + pn); e2 = __ldg(indptr + pn + 1); }
        const float b0 = on0 ? beta[p * kp + lane] : 0.f;
        const float b1 = on1 ? beta[p * kp + 32 + lane] : 0.f;
        const float h0 = on0 ? h[p * kp + lane] : 0.f;
        const float h1 = on1 ? h[p * kp + 32 + lane] : 0.f;
        float n0 = 0.f, n1 = 0.f;
        for (int j0 = s; j0 < e; j0 += 8) {                 // 8 independent row reads in flight
            int nb[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int pre = __shfl_sync(kFull, nb0, u);
                nb[u] = j0 == s ? pre : (j0 + u < e ? __ldg(indices + j0 + u) : -1);
            }
            float v0[8], v1[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                v0[u] = (on0 && nb[u] >= 0) ? beta[(int64_t)nb[u] * kp + lane] : 0.f;
                v1[u] = (on1 && nb[u] >= 0) ? beta[(int64_t)nb[u] * kp + 32 + lane] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) { n0 += v0[u]; n1 += v1[u]; }
        }
        int nb2 = -1;
        if (lane < 8 && s2 + lane < e2) nb2 = __ldg(indices + s2 + lane);     // lands while the Gram product runs
        float gb0 = 0.f, gb1 = 0.f;
        unsigned m = __ballot_sync(kFull, b0 != 0.f);
        while (m) {
            const int c = __ffs(m) - 1;
            m &= m - 1;
            const float bc = __shfl_sync(kFull, b0, c);
            if (on0) gb0 = fmaf(gs[c * kp + lane], bc, gb0);
            if (on1) gb1 = fmaf(gs[c * kp + 32 + lane], bc, gb1);
        }
        if (NK == 2) {
            m = __ballot_sync(kFull, b1 != 0.f);
            while (m) {
                const int c = __ffs(m) - 1;
                m &= m - 1;
                const float bc = __shfl_sync(kFull, b1, c);
                if (on0) gb0 = fmaf(gs[(c + 32) * kp + lane], bc, gb0);
                if (on1) gb1 = fmaf(gs[(c + 32) * kp + 32 + lane], bc, gb1);
            }
        }
        const float deg = (float)(e - s);
        cross += (double)(b0 * h0) + (double)(b1 * h1);
        quad += (double)(b0 * gb0) + (double)(b1 * gb1);
        lap += (double)b0 * (double)(deg * b0 - n0) + (double)b1 * (double)(deg * b1 - n1);
        l1 += (double)fabsf(b0) + (double)fabsf(b1);
        if (lane == 0) yy += (double)ysq[p];
        s = s2; e = e2; nb0 = nb2;
    }
    __shared__ double red[5][8];
    cross = warp_sum(cross); quad = warp_sum(quad); lap = warp_sum(lap); l1 = warp_sum(l1); yy = warp_sum(yy);
    const int warp = threadIdx.x >> 5;
    if (lane == 0) { red[0][warp] = cross; red[1][warp] = quad; red[2][warp] = lap; red[3][warp] = l1; red[4][warp] = yy; }
    __syncthreads();
    if (threadIdx.x < 5) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
        atomicAdd(out + threadIdx.x, t);
    }
}

// ------------------------------------------------------------------------------------
// un-permute + widen + normalise, warp per spot
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
finish_kernel(const float *__restrict__ beta, const int32_t *__restrict__ order, int64_t n_rows, int kp,
              int n_types, double *__restrict__ beta_out, double *__restrict__ prop_out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp_global; p < n_rows; p += n_warps) {
        const double b0 = lane < n_types ? (double)beta[p * kp + lane] : 0.0;
        const double b1 = lane + 32 < n_types ? (double)beta[p * kp + 32 + lane] : 0.0;
        const double tot = warp_sum(b0 + b1);
        const int64_t o = order ? (int64_t)order[p] : p;
        const double den = tot > 1e-10 ? tot : 1e-10;
        const double uni = 1.0 / (double)n_types;
        if (lane < n_types) {
            if (beta_out) beta_out[o * n_types + lane] = b0;
            if (prop_out) prop_out[o * n_types + lane] = tot == 0.0 ? uni : b0 / den;
        }
        if (lane + 32 < n_types) {
            if (beta_out) beta_out[o * n_types + 32 + lane] = b1;
            if (prop_out) prop_out[o * n_types + 32 + lane] = tot == 0.0 ? uni : b1 / den;
        }
    }
}

// dominant cell type per spot (FlashDeconv.get_dominant_cell_type, core/deconv.py:467-478): argmax over the K abundances,
// first maximum on ties like numpy.argmax; equal to the argmax of the proportions (a positive row scaling), 0 for all-zero rows
__global__ void __launch_bounds__(256)
dominant_kernel(const float *__restrict__ beta, const int32_t *__restrict__ order, int64_t n_rows, int kp, int n_types,
                int32_t *__restrict__ out)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_rows) return;
    const float *row = beta + p * kp;
    float best = row[0];
    int arg = 0;
    for (int k = 1; k < n_types; ++k) {
        const float v = row[k];
        if (v > best) { best = v; arg = k; }
    }
    out[order ? (int64_t)order[p] : p] = arg;
}

__global__ void __launch_bounds__(256)
rows_gather_kernel(const float *__restrict__ src, const int32_t *__restrict__ rows, int64_t n_list,
                   int chunks, float *__restrict__ dst)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_list * chunks) return;
    const int64_t i = t / chunks;
    const int q = (int)(t - i * chunks);
    st4(dst + (i * chunks + q) * 4, ld4(src + ((int64_t)rows[i] * chunks + q) * 4));
}

}  // namespace fdb

using namespace fdb;

static int check_solver_args(const void *h, const void *gram, const void *a, const void *b, const void *ptr,
                             int64_t n_rows, int n_types, const void *state)
{
    FDB_REQUIRE(n_rows >= 0 && n_rows < ((int64_t)1 << 31) - 256, "n_rows out of range");
    FDB_REQUIRE(n_types >= 1 && n_types <= FDB_MAX_TYPES, "n_types must be in [1, %d], got %d", FDB_MAX_TYPES, n_types);
    FDB_REQUIRE(n_rows == 0 || (h && gram && a && b && ptr && state), "null pointer");
    return FDB_OK;
}

static int plan_tile_rows(int n_types)
{
    const int kp = fdb_padded_types(n_types);
    if (kp % 8 != 0) return 0;                         // no half gather tile for this row width: no plan needed
    return 128;                                        // must match the dispatcher in launch_sweep
}

extern "C" __attribute__((visibility("default"))) int64_t fdb_bcd_plan_bytes(int64_t n_rows, int64_t nnz, int32_t n_types)
{
    const int tile = plan_tile_rows(n_types);
    if (tile == 0 || n_rows <= 0) return 0;
    const int64_t n_ctas = ceil_div(n_rows, tile);
    return round_up(plan_off_codes(n_ctas, tile) + 2 * (nnz > 0 ? nnz : 1) + 32, 256);   // +32: 16-byte code chunks
}

extern "C" __attribute__((visibility("default"))) int fdb_bcd_plan_build(const int32_t *indptr, const int32_t *indices, int64_t n_rows,
                                          int64_t nnz, int32_t n_types, void *plan, int64_t plan_bytes, void *stream)
{
    const int tile = plan_tile_rows(n_types);
    if (tile == 0 || n_rows <= 0) return FDB_OK;
    FDB_REQUIRE(indptr && indices && plan, "null pointer");
    if (plan_bytes < fdb_bcd_plan_bytes(n_rows, nnz, n_types)) {
        set_error("plan buffer too small: need %lld bytes", (long long)fdb_bcd_plan_bytes(n_rows, nnz, n_types));
        return FDB_ERR_WORKSPACE;
    }
    const int64_t n_ctas = ceil_div(n_rows, tile);
    char *pbase = (char *)plan;
    int32_t *cnt = (int32_t *)(pbase + plan_off_cnt());
    int32_t *rows = (int32_t *)(pbase + plan_off_rows(n_ctas));
    uint16_t *codes = (uint16_t *)(pbase + plan_off_codes(n_ctas, tile));
    uint8_t *codes8 = (uint8_t *)(pbase + plan_off_codes8(n_ctas, tile));
    bcd_plan_kernel<128><<<(int)n_ctas, 128, 0, (cudaStream_t)stream>>>(indptr, indices, (int)n_rows, cnt, rows, codes,
                                                                        codes8);
    FDB_LAUNCH_CHECK("bcd_plan_kernel");
    return FDB_OK;
}

extern "C" __attribute__((visibility("default"))) int fdb_bcd_sweep(const float *h, const float *host_gram, const float *beta_in, float *beta_out,
                             const int32_t *indptr, const int32_t *indices, int64_t n_rows, int32_t n_types,
                             float lambda, float rho_scaled, float tol, int32_t finalize, void *state,
                             const void *plan, void *stream)
{
    int rc = check_solver_args(h, host_gram, beta_in, beta_out, indptr, n_rows, n_types, state);
    if (rc || n_rows == 0) return rc;
    return dispatch_sweep(h, host_gram, n_types, beta_in, beta_out, indptr, indices, n_rows, lambda, rho_scaled,
                          tol, finalize, (SolveState *)state, plan, (cudaStream_t)stream);
}

namespace fdb {
// used by peer.cu: one sweep with the boundary push and the inter-rank hand-shake fused in (production kernel only)
bool sweep_can_fuse_comm(const float *host_gram, int n_types, float lam, const void *plan)
{
    return plan != nullptr && sweep_variant() == 0 && fdb_padded_types(n_types) % 8 == 0 && weak_coupling(host_gram, n_types, lam);
}
int sweep_with_comm(const float *h, const float *host_gram, const float *beta_in, float *beta_out, const int32_t *indptr,
                    const int32_t *indices, int64_t n_rows, int32_t n_types, float lam, float rho, float tol, void *state,
                    const void *plan, void *stream, const SweepComm &comm)
{
    return dispatch_sweep(h, host_gram, n_types, beta_in, beta_out, indptr, indices, n_rows, lam, rho, tol, 1,
                          (SolveState *)state, plan, (cudaStream_t)stream, &comm);
}
}  // namespace fdb

extern "C" __attribute__((visibility("default"))) int fdb_bcd_finalize(void *state, float tol, void *stream)
{
    FDB_REQUIRE(state != nullptr, "null state");
    bcd_finalize_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((SolveState *)state, tol);
    FDB_LAUNCH_CHECK("bcd_finalize_kernel");
    return FDB_OK;
}

extern "C" __attribute__((visibility("default"))) int fdb_bcd_init(float *beta, int64_t n_rows, int32_t n_types, void *state, void *stream)
{
    FDB_REQUIRE(n_rows >= 0 && n_types >= 1, "bad shape");
    const int kp = fdb_padded_types(n_types);
    const int64_t total = n_rows * (kp / 4);
    bcd_init_kernel<<<(int)std::max<int64_t>(1, ceil_div(total, 256)), 256, 0, (cudaStream_t)stream>>>(
        beta, n_rows, kp, n_types, (SolveState *)state);
    FDB_LAUNCH_CHECK("bcd_init_kernel");
    return FDB_OK;
}

extern "C" __attribute__((visibility("default"))) int fdb_bcd_solve(const float *h, const float *host_gram, float *beta_a, float *beta_b,
                             const int32_t *indptr, const int32_t *indices, int64_t n_rows, int32_t n_types,
                             float lambda, float rho_scaled, int32_t max_iter, float tol, void *state,
                             const void *plan, void *stream)
{
    int rc = check_solver_args(h, host_gram, beta_a, beta_b, indptr, n_rows, n_types, state);
    if (rc) return rc;
    FDB_REQUIRE(max_iter >= 0, "max_iter must be non-negative, got %d", max_iter);
    rc = fdb_bcd_init(beta_a, n_rows, n_types, state, stream);
    if (rc || n_rows == 0) return rc;
    float *cur = beta_a, *nxt = beta_b;
    for (int it = 0; it < max_iter; ++it) {
        rc = dispatch_sweep(h, host_gram, n_types, cur, nxt, indptr, indices, n_rows, lambda, rho_scaled, tol, 1,
                            (SolveState *)state, plan, (cudaStream_t)stream);
        if (rc) return rc;
        float *t = cur; cur = nxt; nxt = t;
    }
    return FDB_OK;
}

extern "C" __attribute__((visibility("default"))) int fdb_objective_terms(const float *beta, const float *h, const float *ysq, const float *host_gram,
                                   const int32_t *indptr, const int32_t *indices, int64_t n_rows,
                                   int32_t n_types, double *out, void *stream)
{
    FDB_REQUIRE(n_rows >= 0, "negative n_rows");
    FDB_REQUIRE(n_types >= 1 && n_types <= FDB_MAX_TYPES, "n_types must be in [1, %d], got %d", FDB_MAX_TYPES, n_types);
    if (n_rows == 0) return FDB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int kp = fdb_padded_types(n_types);
    GramFull gram;
    for (int i = 0; i < kp * kp; ++i) gram.g[i] = 0.f;
    for (int a = 0; a < n_types; ++a)
        for (int c = 0; c < n_types; ++c) gram.g[a * kp + c] = host_gram[a * n_types + c];
    const int grid = (int)std::min<int64_t>(ceil_div(n_rows, 8), (int64_t)kNumSM * 8);
    const size_t smem = (size_t)kp * kp * 4;
    if (kp <= 32)
        objective_kernel<1><<<grid, 256, smem, st>>>(beta, h, ysq, gram, indptr, indices, n_rows, kp, out);
    else
        objective_kernel<2><<<grid, 256, smem, st>>>(beta, h, ysq, gram, indptr, indices, n_rows, kp, out);
    FDB_LAUNCH_CHECK("objective_kernel");
    return FDB_OK;
}

extern "C" __attribute__((visibility("default"))) int fdb_finish(const float *beta, const int32_t *order, int64_t n_rows, int32_t n_types,
                          double *beta_out, double *prop_out, void *stream)
{
    FDB_REQUIRE(n_rows >= 0, "negative n_rows");
    FDB_REQUIRE(n_types >= 1 && n_types <= FDB_MAX_TYPES_WIDE, "n_types must be in [1, %d], got %d", FDB_MAX_TYPES_WIDE, n_types);
    if (n_rows == 0) return FDB_OK;
    if (n_types > FDB_MAX_TYPES) return finish_wide(beta, order, n_rows, n_types, beta_out, prop_out, (cudaStream_t)stream);
    const int grid = (int)std::min<int64_t>(ceil_div(n_rows, 8), (int64_t)kNumSM * 16);
    finish_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(beta, order, n_rows, fdb_padded_types(n_types), n_types,
                                                         beta_out, prop_out);
    FDB_LAUNCH_CHECK("finish_kernel");
    return FDB_OK;
}

extern "C" __attribute__((visibility("default"))) int fdb_dominant_type(const float *beta, const int32_t *order, int64_t n_rows, int32_t n_types,
                                 int32_t *dominant, void *stream)
{
    FDB_REQUIRE(n_rows >= 0 && n_types >= 1 && n_types <= FDB_MAX_TYPES_WIDE, "bad shape");
    if (n_rows == 0) return FDB_OK;
    FDB_REQUIRE(beta && dominant, "null pointer");
    dominant_kernel<<<(int)ceil_div(n_rows, 256), 256, 0, (cudaStream_t)stream>>>(beta, order, n_rows, fdb_padded_types(n_types),
                                                                               n_types, dominant);
    FDB_LAUNCH_CHECK("dominant_kernel");
    return FDB_OK;
}

extern "C" __attribute__((visibility("default"))) int fdb_rows_gather(const float *src, const int32_t *rows, int64_t n_list, int32_t row_floats,
                               float *dst, void *stream)
{
    FDB_REQUIRE(n_list >= 0 && row_floats > 0 && row_floats % 4 == 0, "row_floats must be a positive multiple of 4");
    if (n_list == 0) return FDB_OK;
    const int chunks = row_floats / 4;
    rows_gather_kernel<<<(int)ceil_div(n_list * chunks, 256), 256, 0, (cudaStream_t)stream>>>(src, rows, n_list,
                                                                                             chunks, dst);
    FDB_LAUNCH_CHECK("rows_gather_kernel");
    return FDB_OK;
}
