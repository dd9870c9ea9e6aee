// libfdb200 -- any number of cell types (K > FDB_MAX_TYPES).  The reference has no limit on K (core/solver.py:287-428 works on
// K x K / K x N arrays of any size); the register-resident kernels of bcd.cu / bcd_p.cuh stop at K = 64 because a spot's whole
// row lives in one thread's registers.  Here a WARP owns a spot and the row lives in shared memory, lane l holding entries
// l, l + 32, ...; the Gram matrix stays in global memory (K^2 floats, L1/L2-resident) and is read row-wise (G is symmetric,
// so column k is row k).  The sweep is the reference's maintained-residual form restated for a warp:
//   core/solver.py:29-101   per-spot cyclic coordinate descent with resid = XtX beta kept up to date (update only when the
//                           step is non-zero: 85-90 % of the entries stay exactly 0)
//   core/solver.py:104-184  Jacobi sweep (neighbour sums from the previous iterate) + fused max-norm statistics
//   core/solver.py:269-284  objective terms;  core/solver.py:431-452  normalisation
// Stop test, state block and buffer ping-pong are those of fdb_bcd_solve.  This path is about coverage, not speed: one warp
// per spot costs ~K + K * nnz(beta) / 32 dependent steps per spot.
#include <algorithm>
#include "bcd_state.cuh"

namespace fdb {

constexpr int kWideWarps = 8;

// rows of the warp-private arrays: b (current beta row), r (Gram * b), c (H + lam * neighbour sum)
__global__ void __launch_bounds__(kWideWarps * 32)
bcd_sweep_wide_kernel(const float *__restrict__ h, const float *__restrict__ gram, const float *__restrict__ beta_in,
                      float *__restrict__ beta_out, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                      int64_t n_rows, int n_types, int kp, float lam, float rho, float tol, int finalize, SolveState *state)
{
    if (*reinterpret_cast<volatile int *>(&state->converged)) return;
    extern __shared__ __align__(16) float wide_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *diag = wide_smem;                                   // kp
    float *b = wide_smem + kp + warp * 3 * kp, *r = b + kp, *c = r + kp;
    __shared__ unsigned red[2][kWideWarps];
    for (int k = threadIdx.x; k < kp; k += blockDim.x) diag[k] = k < n_types ? gram[(size_t)k * kp + k] : 0.f;
    __syncthreads();
    const int64_t warp_global = (int64_t)blockIdx.x * kWideWarps + warp;
    const int64_t n_warps = (int64_t)gridDim.x * kWideWarps;
    float dmax = 0.f, amax = 0.f;
    for (int64_t p = warp_global; p < n_rows; p += n_warps) {
        const int s = __ldg(indptr + p), e = __ldg(indptr + p + 1);
        const float lam_deg = lam * (float)(e - s);
        for (int k = lane; k < kp; k += 32) {
            const float v = beta_in[p * kp + k];
            b[k] = v;
            r[k] = 0.f;
            c[k] = 0.f;
            amax = fmaxf(amax, fabsf(v));
        }
        for (int j = s; j < e; ++j) {                          // neighbour sums, previous iterate (core/solver.py:152-157)
            const float *nb = beta_in + (int64_t)__ldg(indices + j) * kp;
            for (int k = lane; k < kp; k += 32) c[k] += nb[k];
        }
        for (int k = lane; k < kp; k += 32) c[k] = fmaf(lam, c[k], h[p * kp + k]);
        __syncwarp();
        for (int j = 0; j < n_types; ++j) {                    // r = G b over the non-zero entries of b
            const float bj = b[j];
            if (bj != 0.f) {
                const float *g = gram + (size_t)j * kp;
                for (int k = lane; k < kp; k += 32) r[k] = fmaf(bj, g[k], r[k]);
            }
        }
        for (int k = 0; k < n_types; ++k) {                    // cyclic coordinate descent (core/solver.py:68-99)
            __syncwarp();
            const float bk = b[k], gkk = diag[k];
            const float part = c[k] - r[k] + gkk * bk;
            const float den = gkk + lam_deg;
            const float nv = den > 1e-10f ? fmaxf(0.f, (part - rho) / den) : 0.f;     // max(0, soft(part, rho) / den)
            const float step = nv - bk;
            dmax = fmaxf(dmax, fabsf(step));
            if (step != 0.f) {                                 // warp-uniform: every lane computed the same numbers
                __syncwarp();                                  // every lane has read b[k], r[k] before they change
                const float *g = gram + (size_t)k * kp;
                for (int j = lane; j < kp; j += 32) r[j] = fmaf(step, g[j], r[j]);
                if (lane == (k & 31)) b[k] = nv;
            }
        }
        __syncwarp();
        for (int k = lane; k < kp; k += 32) beta_out[p * kp + k] = b[k];
        __syncwarp();
    }
    const unsigned wd = __reduce_max_sync(kFull, __float_as_uint(dmax));
    const unsigned wa = __reduce_max_sync(kFull, __float_as_uint(amax));
    if (lane == 0) { red[0][warp] = wd; red[1][warp] = wa; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned bd = 0u, ba = 0u;
        for (int w = 0; w < kWideWarps; ++w) { bd = max(bd, red[0][w]); ba = max(ba, red[1][w]); }
        atomicMax(&state->max_diff_bits, bd);
        atomicMax(&state->max_abs_bits, ba);
        if (finalize) {
            __threadfence();
            if (atomicAdd(&state->arrived, 1u) == gridDim.x - 1) {
                __threadfence();
                finalize_state(state, tol);
            }
        }
    }
}

// objective terms, warp per spot, float64 accumulation (same five numbers as objective_kernel in bcd.cu)
__global__ void __launch_bounds__(kWideWarps * 32)
objective_wide_kernel(const float *__restrict__ beta, const float *__restrict__ h, const float *__restrict__ ysq,
                      const float *__restrict__ gram, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                      int64_t n_rows, int n_types, int kp, double *__restrict__ out)
{
    extern __shared__ __align__(16) float wide_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *b = wide_smem + warp * kp;
    const int64_t warp_global = (int64_t)blockIdx.x * kWideWarps + warp;
    const int64_t n_warps = (int64_t)gridDim.x * kWideWarps;
    double cross = 0.0, quad = 0.0, lap = 0.0, l1 = 0.0, yy = 0.0;
    for (int64_t p = warp_global; p < n_rows; p += n_warps) {
        const int s = __ldg(indptr + p), e = __ldg(indptr + p + 1);
        const float deg = (float)(e - s);
        for (int k = lane; k < kp; k += 32) b[k] = beta[p * kp + k];
        __syncwarp();
        for (int k = lane; k < n_types; k += 32) {
            const float bk = b[k];
            float ns = 0.f, gb = 0.f;
            for (int j = s; j < e; ++j) ns += beta[(int64_t)__ldg(indices + j) * kp + k];
            for (int j = 0; j < n_types; ++j) {
                const float bj = b[j];
                if (bj != 0.f) gb = fmaf(gram[(size_t)j * kp + k], bj, gb);
            }
            cross += (double)(bk * h[p * kp + k]);
            quad += (double)(bk * gb);
            lap += (double)bk * (double)(deg * bk - ns);
            l1 += (double)fabsf(bk);
        }
        if (lane == 0) yy += (double)ysq[p];
        __syncwarp();
    }
    __shared__ double red[5][kWideWarps];
    cross = warp_sum(cross); quad = warp_sum(quad); lap = warp_sum(lap); l1 = warp_sum(l1); yy = warp_sum(yy);
    if (lane == 0) { red[0][warp] = cross; red[1][warp] = quad; red[2][warp] = lap; red[3][warp] = l1; red[4][warp] = yy; }
    __syncthreads();
    if (threadIdx.x < 5) {
        double t = 0.0;
        for (int w = 0; w < kWideWarps; ++w) t += red[threadIdx.x][w];
        atomicAdd(out + threadIdx.x, t);
    }
}

// un-permute + widen + normalise for any K, warp per spot (core/solver.py:431-452)
__global__ void __launch_bounds__(256)
finish_wide_kernel(const float *__restrict__ beta, const int32_t *__restrict__ order, int64_t n_rows, int kp, int n_types,
                   double *__restrict__ beta_out, double *__restrict__ prop_out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp_global; p < n_rows; p += n_warps) {
        double part = 0.0;
        for (int k = lane; k < n_types; k += 32) part += (double)beta[p * kp + k];
        const double tot = warp_sum(part);
        const int64_t o = order ? (int64_t)order[p] : p;
        const double den = tot > 1e-10 ? tot : 1e-10;
        const double uni = 1.0 / (double)n_types;
        for (int k = lane; k < n_types; k += 32) {
            const double v = (double)beta[p * kp + k];
            if (beta_out) beta_out[o * n_types + k] = v;
            if (prop_out) prop_out[o * n_types + k] = tot == 0.0 ? uni : v / den;
        }
    }
}

int finish_wide(const float *beta, const int32_t *order, int64_t n_rows, int n_types, double *beta_out, double *prop_out,
                cudaStream_t st)
{
    const int grid = (int)std::min<int64_t>(ceil_div(n_rows, 8), (int64_t)kNumSM * 16);
    finish_wide_kernel<<<grid, 256, 0, st>>>(beta, order, n_rows, fdb_padded_types(n_types), n_types, beta_out, prop_out);
    FDB_LAUNCH_CHECK("finish_wide_kernel");
    return FDB_OK;
}

static int launch_sweep_wide(const float *h, const float *gram, const float *beta_in, float *beta_out, const int32_t *indptr,
                             const int32_t *indices, int64_t n_rows, int n_types, float lam, float rho, float tol, int finalize,
                             SolveState *state, cudaStream_t st)
{
    const int kp = fdb_padded_types(n_types);
    const size_t smem = (size_t)(1 + 3 * kWideWarps) * kp * 4;
    FDB_CUDA(cudaFuncSetAttribute(bcd_sweep_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_sm = std::max<int>(1, std::min<int>(8, (int)((size_t)200 * 1024 / (smem + 1024))));
    const int grid = (int)std::min<int64_t>(ceil_div(n_rows, kWideWarps), (int64_t)kNumSM * per_sm);
    bcd_sweep_wide_kernel<<<grid, kWideWarps * 32, smem, st>>>(h, gram, beta_in, beta_out, indptr, indices, n_rows, n_types, kp, lam,
                                                              rho, tol, finalize, state);
    FDB_LAUNCH_CHECK("bcd_sweep_wide_kernel");
    return FDB_OK;
}

}  // namespace fdb

using namespace fdb;

static int check_wide_args(const void *h, const void *gram, const void *a, const void *b, const void *ptr, int64_t n_rows, int n_types,
                           const void *state)
{
    FDB_REQUIRE(n_rows >= 0 && n_rows < ((int64_t)1 << 31) - 256, "n_rows out of range");
    FDB_REQUIRE(n_types >= 1 && n_types <= FDB_MAX_TYPES_WIDE, "n_types must be in [1, %d], got %d", FDB_MAX_TYPES_WIDE, n_types);
    FDB_REQUIRE(n_rows == 0 || (h && gram && a && b && ptr && state), "null pointer");
    return FDB_OK;
}

extern "C" __attribute__((visibility("default"))) int fdb_bcd_sweep_wide(const float *h, const float *gram_dev, const float *beta_in,
                                                                         float *beta_out, const int32_t *indptr, const int32_t *indices,
                                                                         int64_t n_rows, int32_t n_types, float lambda, float rho_scaled,
                                                                         float tol, int32_t finalize, void *state, void *stream)
{
    int rc = check_wide_args(h, gram_dev, beta_in, beta_out, indptr, n_rows, n_types, state);
    if (rc || n_rows == 0) return rc;
    return launch_sweep_wide(h, gram_dev, beta_in, beta_out, indptr, indices, n_rows, n_types, lambda, rho_scaled, tol, finalize,
                             (SolveState *)state, (cudaStream_t)stream);
}

extern "C" __attribute__((visibility("default"))) int fdb_bcd_solve_wide(const float *h, const float *gram_dev, float *beta_a, float *beta_b,
                                                                         const int32_t *indptr, const int32_t *indices, int64_t n_rows,
                                                                         int32_t n_types, float lambda, float rho_scaled, int32_t max_iter,
                                                                         float tol, void *state, void *stream)
{
    int rc = check_wide_args(h, gram_dev, beta_a, beta_b, indptr, n_rows, n_types, state);
    if (rc) return rc;
    FDB_REQUIRE(max_iter >= 0, "max_iter must be non-negative, got %d", max_iter);
    rc = fdb_bcd_init(beta_a, n_rows, n_types, state, stream);
    if (rc || n_rows == 0) return rc;
    float *cur = beta_a, *nxt = beta_b;
    for (int it = 0; it < max_iter; ++it) {
        rc = launch_sweep_wide(h, gram_dev, cur, nxt, indptr, indices, n_rows, n_types, lambda, rho_scaled, tol, 1, (SolveState *)state,
                               (cudaStream_t)stream);
        if (rc) return rc;
        float *t = cur; cur = nxt; nxt = t;
    }
    return FDB_OK;
}

extern "C" __attribute__((visibility("default"))) int fdb_objective_terms_wide(const float *beta, const float *h, const float *ysq,
                                                                               const float *gram_dev, const int32_t *indptr,
                                                                               const int32_t *indices, int64_t n_rows, int32_t n_types,
                                                                               double *out, void *stream)
{
    FDB_REQUIRE(n_rows >= 0, "negative n_rows");
    FDB_REQUIRE(n_types >= 1 && n_types <= FDB_MAX_TYPES_WIDE, "n_types must be in [1, %d], got %d", FDB_MAX_TYPES_WIDE, n_types);
    if (n_rows == 0) return FDB_OK;
    FDB_REQUIRE(beta && h && ysq && gram_dev && indptr && out, "null pointer");
    const int kp = fdb_padded_types(n_types);
    const size_t smem = (size_t)kWideWarps * kp * 4;
    FDB_CUDA(cudaFuncSetAttribute(objective_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<int64_t>(ceil_div(n_rows, kWideWarps), (int64_t)kNumSM * 4);
    objective_wide_kernel<<<grid, kWideWarps * 32, smem, (cudaStream_t)stream>>>(beta, h, ysq, gram_dev, indptr, indices, n_rows, n_types,
                                                                                kp, out);
    FDB_LAUNCH_CHECK("objective_wide_kernel");
    return FDB_OK;
}
