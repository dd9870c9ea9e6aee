/*
 * oracle/bcd_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C (float64, OpenMP) restatement of the reference's Jacobi block-
 * coordinate-descent sweep and of the per-row log-CPM + CountSketch
 * projection.  It exists only as the CPU checker for the CUDA path
 * (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl
 * reference legs).  Nothing under flashdeconv_b200/ may link or call it.
 *
 * Reference lines restated (paths into the upstream repo):
 *   flashdeconv/core/solver.py:18-26    soft threshold
 *   flashdeconv/core/solver.py:29-101   per-spot cyclic coordinate descent
 *   flashdeconv/core/solver.py:104-184  Jacobi sweep + fused max-norm stats
 *   flashdeconv/core/solver.py:431-452  proportion normalisation
 *   flashdeconv/core/deconv.py:181-188  sparse log-CPM (lib over selected genes, 0 -> 1)
 *   flashdeconv/core/sketching.py:195   Y_tilde @ Omega with 1 nnz / row of Omega
 *
 * Layout differs from the reference on purpose: H is spot-major (N x K), so
 * the oracle takes `h_spot_major`; numerically the sweep is identical.
 *
 * Pinned against the imported reference by oracle/pin_against_reference.py
 * (agreement to ~1e-12 relative; the reference's BLAS matvec orders its sums
 * differently, so bit-equality is not claimed for float64 sums).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline double shrink(double v, double t)
{
    if (v > t)  return v - t;
    if (v < -t) return v + t;
    return 0.0;
}

int fdo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* explicit thread count (bench.py: torchrun exports OMP_NUM_THREADS=1 to its workers) */
void fdo_set_num_threads(int n)
{
    if (n > 0) omp_set_num_threads(n);
}


/* One Jacobi sweep over all spots.  beta_prev is read-only; beta_next gets the
 * update.  diff_out / abs_out receive per-spot max|new-old| and max|old|.   */
void fdo_bcd_sweep(const double *h_spot_major,   /* N x K  */
                   const double *gram,           /* K x K  */
                   const double *beta_prev,      /* N x K  */
                   double *beta_next,            /* N x K  */
                   const int64_t *nbr_ptr,       /* N + 1  */
                   const int64_t *nbr_idx,
                   int64_t n_spots, int n_types,
                   double lam, double rho_scaled,
                   double *diff_out, double *abs_out)
{
    const int K = n_types;
#pragma omp parallel
    {
        double *resid = (double *)malloc(sizeof(double) * (size_t)K);
        double *nsum  = (double *)malloc(sizeof(double) * (size_t)K);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n_spots; ++i) {
            const double *b_old = beta_prev + i * K;
            double *b = beta_next + i * K;
            const double *hi = h_spot_major + i * K;
            const int64_t lo = nbr_ptr[i], hi_ptr = nbr_ptr[i + 1];
            const int64_t deg = hi_ptr - lo;

            memcpy(b, b_old, sizeof(double) * (size_t)K);
            memset(nsum, 0, sizeof(double) * (size_t)K);
            for (int64_t e = lo; e < hi_ptr; ++e) {
                const double *bj = beta_prev + nbr_idx[e] * K;
                for (int k = 0; k < K; ++k) nsum[k] += bj[k];
            }
            /* maintained product  resid = gram * b  */
            for (int a = 0; a < K; ++a) {
                double acc = 0.0;
                const double *g = gram + (size_t)a * K;
                for (int c = 0; c < K; ++c) acc += g[c] * b[c];
                resid[a] = acc;
            }
            for (int k = 0; k < K; ++k) {
                const double gkk = gram[(size_t)k * K + k];
                const double prev = b[k];
                double part = hi[k] - resid[k] + gkk * prev;
                if (deg > 0) part += lam * nsum[k];
                const double den = gkk + lam * (double)deg;
                double nv = 0.0;
                if (den > 1e-10) {
                    nv = shrink(part, rho_scaled) / den;
                    if (nv < 0.0) nv = 0.0;
                }
                b[k] = nv;
                const double step = nv - prev;
                if (step != 0.0)
                    for (int a = 0; a < K; ++a) resid[a] += step * gram[(size_t)a * K + k];
            }
            double dmax = 0.0, amax = 0.0;
            for (int k = 0; k < K; ++k) {
                const double d = fabs(b[k] - b_old[k]);
                const double a = fabs(b_old[k]);
                if (d > dmax) dmax = d;
                if (a > amax) amax = a;
            }
            diff_out[i] = dmax;
            abs_out[i] = amax;
        }
        free(resid);
        free(nsum);
    }
}

/* Row-normalise; all-zero rows become uniform 1/K. */
void fdo_normalize(const double *beta, double *prop, int64_t n_spots, int n_types)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_spots; ++i) {
        double s = 0.0;
        for (int k = 0; k < n_types; ++k) s += beta[i * n_types + k];
        if (s == 0.0) {
            for (int k = 0; k < n_types; ++k) prop[i * n_types + k] = 1.0 / n_types;
        } else {
            const double den = s > 1e-10 ? s : 1e-10;
            for (int k = 0; k < n_types; ++k) prop[i * n_types + k] = beta[i * n_types + k] / den;
        }
    }
}

/* log-CPM + CountSketch of a CSR count matrix over the FULL gene axis.
 * gene_bucket[g] < 0 marks an unselected gene; lib size sums selected genes
 * only (deconv.py:321 subsets before :183 sums).  y_sketch is N x d, zeroed
 * by the caller.                                                            */
void fdo_sketch_csr(const int64_t *indptr, const int32_t *indices, const double *counts,
                    int64_t n_spots, const int32_t *gene_bucket, const double *gene_weight,
                    int d, double *y_sketch)
{
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n_spots; ++i) {
        double lib = 0.0;
        for (int64_t e = indptr[i]; e < indptr[i + 1]; ++e)
            if (gene_bucket[indices[e]] >= 0) lib += counts[e];
        if (lib == 0.0) lib = 1.0;
        const double scale = 1e4 / lib;
        double *row = y_sketch + i * (int64_t)d;
        for (int64_t e = indptr[i]; e < indptr[i + 1]; ++e) {
            const int32_t g = indices[e];
            const int32_t b = gene_bucket[g];
            if (b >= 0) row[b] += log1p(scale * counts[e]) * gene_weight[g];
        }
    }
}
