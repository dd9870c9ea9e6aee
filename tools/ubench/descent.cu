// Microbenchmark of the wide-row (Kp = 56, K = 50) pair-step descent of bcd_sweep_p_kernel in isolation: 3 CTAs x 4 warps per SM,
// every warp streaming the 12.5 KB Gram operand from the constant bank once per 32 spots.  Variants probe whether the
// constant-cache misses (L1 constant = 2 KB, miss = 102 cycles: tools/ubench/ubench.cu) can be hidden:
//   0 baseline   1 CTA barrier per pair   2 barrier + cooperative touch of the next pair's lines   3 every warp touches its own next lines
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I flashdeconv_b200/csrc -o tools/ubench/descent tools/ubench/descent.cu
#include <cstdio>
#include <vector>
#include "bcd_common.cuh"
using namespace fdb;

// symmetric block storage: 2x2 blocks B(m,n), m < n, as (-G[2m][2n], -G[2m+1][2n], -G[2m][2n+1], -G[2m+1][2n+1]);
// pair step m uses B(m,n), n > m, column-wise (FFMA2 with a broadcast beta) and B(n,m), n < m, row-wise (FFMA2 with the
// beta PAIR of pair n and a horizontal sum at the end): half the constant footprint of the pair-row layout
template <int KP>
struct alignas(16) GramSymArg {
    float blk[(KP / 2) * (KP / 2 - 1) / 2 * 4];
    float cross[KP];
    float diag[KP];
};
template <int NPAIR>
__host__ __device__ constexpr int bidx(int m, int n) { return m * (2 * NPAIR - m - 1) / 2 + (n - m - 1); }

template <int KP, int PADC, int NCONST, int MODE>
__global__ void __launch_bounds__(128, KP <= 48 ? 4 : 3)
descent_sym(const __grid_constant__ GramSymArg<KP> G, float *state, int n_types, float lam, float rho, int rounds, long long *cycles)
{
    extern __shared__ __align__(16) float smem[];
    constexpr int Q = KP / 4, S = KP + 4, NPAIR = (KP - PADC) / 2, NB = NPAIR * (NPAIR - 1) / 2;
    const int own = threadIdx.x;
    float *crow = smem + own * S;
    float *sG = smem + 128 * S;                            // blocks NCONST.. of the Gram operand
    for (int i = threadIdx.x; i < (NB > NCONST ? (NB - NCONST) * 4 : 0); i += 128) sG[i] = G.blk[NCONST * 4 + i];
    float b[KP];
    for (int q = 0; q < Q; ++q) {
        const float4 v = ld4(state + ((size_t)blockIdx.x * 128 + own) * KP + 4 * q);
        b[4 * q] = v.x; b[4 * q + 1] = v.y; b[4 * q + 2] = v.z; b[4 * q + 3] = v.w;
        st4(crow + 4 * q, make_float4(0.3f + 0.01f * q, 0.2f, 0.1f + 0.001f * own, 0.25f));
    }
    __syncthreads();
    float dmax = 0.f;
    const long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < rounds; ++r) {
        const float lam_deg = lam * 6.f, neg_rho = -rho;
        float dm = 0.f;
        // two opaque, uniform, loop-variant zeros: the compiler can neither hoist the Gram loads out of the loop nor merge
        // the column-wise and the row-wise read of a block (which would keep 5 KB of values live across the whole body)
        const int zv = MODE == 0 ? (r >> 28) & 4 : 0, zh = MODE == 2 ? (b[0] < -1e30f ? 4 : 0) + ((r >> 27) & 8) : (r >> 27) & 8;
        static_for<0, Q>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            if constexpr (4 * q >= KP - PADC) return;
            const float4 c4 = ld4(crow + 4 * q);
            static_for<0, 2>([&](auto mc) {
                constexpr int m = 2 * q + decltype(mc)::value;
                constexpr int k0 = 2 * m;
                if constexpr (k0 >= KP - PADC) return;
                if (k0 >= KP - 8 && k0 >= n_types) return;
                const float den0 = G.diag[k0] + lam_deg, den1 = G.diag[k0 + 1] + lam_deg;
                const float ri0 = den0 > 1e-10f ? rcp_fast(den0) : 0.f;
                const float ri1 = den1 > 1e-10f ? rcp_fast(den1) : 0.f;
                u64 a0 = pack2(elem(c4, k0 & 3), elem(c4, (k0 & 3) + 1));
                u64 a1 = pack2(neg_rho, neg_rho);
                u64 h0 = pack2(0.f, 0.f), h1 = pack2(0.f, 0.f);
                auto block = [&](auto ic, int z, u64 &g01, u64 &g23) {
                    constexpr int idx = decltype(ic)::value;
                    if constexpr (idx < NCONST) {
                        const float4 g = *reinterpret_cast<const float4 *>(G.blk + 4 * idx + z);
                        g01 = pack2(g.x, g.y);
                        g23 = pack2(g.z, g.w);
                    } else {
                        const float4 g = ld4(sG + 4 * (idx - NCONST));
                        g01 = pack2(g.x, g.y);
                        g23 = pack2(g.z, g.w);
                    }
                };
                // not yet updated pairs n > m: column-wise
                static_for<m + 1, NPAIR>([&](auto nc) {
                    constexpr int n = decltype(nc)::value;
                    u64 g01, g23;
                    block(std::integral_constant<int, bidx<NPAIR>(m, n)>{}, zv, g01, g23);
                    a0 = fma2(g01, pack2(b[2 * n], b[2 * n]), a0);
                    a1 = fma2(g23, pack2(b[2 * n + 1], b[2 * n + 1]), a1);
                });
                // already updated pairs n < m, oldest first: row-wise
                static_for<0, m>([&](auto nc) {
                    constexpr int n = decltype(nc)::value;
                    u64 g01, g23;
                    block(std::integral_constant<int, bidx<NPAIR>(n, m)>{}, zh, g01, g23);
                    const u64 bp = pack2(b[2 * n], b[2 * n + 1]);
                    h0 = fma2(g01, bp, h0);
                    h1 = fma2(g23, bp, h1);
                });
                float p0, p1, x0, x1, y0, y1;
                unpack2(add2q(a0, a1), p0, p1);
                unpack2(h0, x0, x1);
                unpack2(h1, y0, y1);
                p0 += x0 + x1;
                p1 += y0 + y1;
                p0 = fmaf(G.cross[k0], b[k0 + 1], p0);
                const float nv0 = fmaxf(0.f, p0 * ri0);
                dm = fmaxf(dm, fabsf(nv0 - b[k0]));
                b[k0] = nv0;
                p1 = fmaf(G.cross[k0 + 1], nv0, p1);
                const float nv1 = fmaxf(0.f, p1 * ri1);
                dm = fmaxf(dm, fabsf(nv1 - b[k0 + 1]));
                b[k0 + 1] = nv1;
            });
            st4(crow + 4 * q, make_float4(b[4 * q] + 0.3f, b[4 * q + 1] + 0.2f, b[4 * q + 2] + 0.1f, b[4 * q + 3] + 0.25f));
        });
        dmax = fmaxf(dmax, dm);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    for (int q = 0; q < Q; ++q)
        st4(state + ((size_t)blockIdx.x * 128 + own) * KP + 4 * q, make_float4(b[4 * q], b[4 * q + 1] + dmax, b[4 * q + 2], b[4 * q + 3]));
}

template <int KP, int PADC, int VAR>
__global__ void __launch_bounds__(128, KP <= 48 ? 4 : 3)
descent_bench(const __grid_constant__ GramPairArg<KP> G, float *state, int n_types, float lam, float rho, int rounds, long long *cycles)
{
    extern __shared__ __align__(16) float smem[];
    constexpr int Q = KP / 4, S = KP + 4;
    const int own = threadIdx.x, warp = threadIdx.x >> 5;
    float *crow = smem + own * S;
    float b[KP];
    for (int q = 0; q < Q; ++q) {
        const float4 v = ld4(state + ((size_t)blockIdx.x * 128 + own) * KP + 4 * q);
        b[4 * q] = v.x; b[4 * q + 1] = v.y; b[4 * q + 2] = v.z; b[4 * q + 3] = v.w;
        st4(crow + 4 * q, make_float4(0.3f + 0.01f * q, 0.2f, 0.1f + 0.001f * own, 0.25f));
    }
    __syncthreads();
    float dmax = 0.f;
    const long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < rounds; ++r) {
        const float lam_deg = lam * 6.f, neg_rho = -rho;
        float dm = 0.f;
        static_for<0, Q>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            if constexpr (4 * q >= KP - PADC) return;
            const float4 c4 = ld4(crow + 4 * q);
            static_for<0, 2>([&](auto mc) {
                constexpr int m = 2 * q + decltype(mc)::value;
                constexpr int k0 = 2 * m;
                if constexpr (k0 >= KP - PADC) return;
                if (k0 >= KP - 8 && k0 >= n_types) return;
                if (VAR == 1 || VAR == 2) __syncthreads();
                float t0v = 0.f, t1v = 0.f;
                if (VAR == 2) {                         // warp w touches lines w and w + 4 of the next pair's row
                    constexpr int nm = (m + 1) % (KP / 2);
                    t0v = G.g2[nm * KP * 2 + min(16 * warp, 2 * KP - 1)];
                    t1v = G.g2[nm * KP * 2 + min(16 * (warp + 4), 2 * KP - 1)];
                }
                float tv[8];
                if (VAR == 3) {
                    constexpr int nm = (m + 1) % (KP / 2);
#pragma unroll
                    for (int l = 0; l < 8; ++l) tv[l] = G.g2[nm * KP * 2 + min(16 * l + (warp & 1), 2 * KP - 1)];
                }
                const float den0 = G.diag[k0] + lam_deg, den1 = G.diag[k0 + 1] + lam_deg;
                const float ri0 = den0 > 1e-10f ? rcp_fast(den0) : 0.f;
                const float ri1 = den1 > 1e-10f ? rcp_fast(den1) : 0.f;
                u64 a0 = pack2(elem(c4, k0 & 3), elem(c4, (k0 & 3) + 1));
                u64 a1 = pack2(neg_rho, neg_rho);
                static_for<2, KP>([&](auto ic) {
                    constexpr int i = decltype(ic)::value;
                    constexpr int jj = (k0 + i) % KP;
                    if constexpr (jj < KP - PADC) {
                        const u64 g = pack2(G.g2[(m * KP + jj) * 2], G.g2[(m * KP + jj) * 2 + 1]);
                        if constexpr (i & 1) a1 = fma2(g, pack2(b[jj], b[jj]), a1);
                        else a0 = fma2(g, pack2(b[jj], b[jj]), a0);
                    }
                });
                float p0, p1;
                unpack2(add2q(a0, a1), p0, p1);
                p0 = fmaf(G.cross[k0], b[k0 + 1], p0);
                const float nv0 = fmaxf(0.f, p0 * ri0);
                dm = fmaxf(dm, fabsf(nv0 - b[k0]));
                b[k0] = nv0;
                p1 = fmaf(G.cross[k0 + 1], nv0, p1);
                const float nv1 = fmaxf(0.f, p1 * ri1);
                float dd = fabsf(nv1 - b[k0 + 1]);
                if (VAR == 2) dd = fmaf(t0v, 0.f, fmaf(t1v, 0.f, dd));
                if (VAR == 3) {
#pragma unroll
                    for (int l = 0; l < 8; ++l) dd = fmaf(tv[l], 0.f, dd);
                }
                dm = fmaxf(dm, dd);
                b[k0 + 1] = nv1;
            });
            st4(crow + 4 * q, make_float4(b[4 * q] + 0.3f, b[4 * q + 1] + 0.2f, b[4 * q + 2] + 0.1f, b[4 * q + 3] + 0.25f));
        });
        dmax = fmaxf(dmax, dm);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    for (int q = 0; q < Q; ++q)
        st4(state + ((size_t)blockIdx.x * 128 + own) * KP + 4 * q, make_float4(b[4 * q], b[4 * q + 1] + dmax, b[4 * q + 2], b[4 * q + 3]));
}


template <typename Kern, typename Arg>
double launch(Kern kern, const Arg &G, float *d_state, long long *d_cyc, int K, int cps, size_t smem, std::vector<float> &out, int KP, const char *name)
{
    const int grid = 148 * cps, rounds = 200;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaMemset(d_state, 0, (size_t)148 * 6 * 128 * KP * 4);
    kern<<<grid, 128, smem>>>(G, d_state, K, 0.01f, 0.001f, 3, d_cyc);        // 3 rounds from zero: the numerics sample
    out.resize((size_t)128 * KP);
    cudaMemcpy(out.data(), d_state, out.size() * 4, cudaMemcpyDeviceToHost);
    cudaEventRecord(e0);
    kern<<<grid, 128, smem>>>(G, d_state, K, 0.01f, 0.001f, rounds, d_cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), d_cyc, grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (auto c : h) avg += (double)c;
    avg /= grid;
    printf("%-28s Kp %d K %d, %d CTAs/SM: %6.0f cycles per descent of 32 spots per warp (kernel %.3f ms, %s)\n", name, KP, K, cps, avg / rounds, ms,
           cudaGetErrorString(cudaGetLastError()));
    return avg / rounds;
}

template <int KP, int PADC, int NCONST>
void compare(int K, int cps)
{
    static GramPairArg<KP> G;
    static GramSymArg<KP> Gs;
    constexpr int NPAIR = (KP - PADC) / 2;
    std::vector<float> full(KP * KP, 0.f);
    for (int k = 0; k < K; ++k)
        for (int j = 0; j < K; ++j) full[k * KP + j] = k == j ? 1.0f + 0.01f * k : 0.01f + 1e-4f * (((j + k) * 7 + j * k) % 13);
    for (int i = 0; i < KP * KP; ++i) G.g2[i] = 0.f;
    for (auto &x : Gs.blk) x = 0.f;
    for (int k = 0; k < KP; ++k) {
        G.diag[k] = Gs.diag[k] = full[k * KP + k];
        G.cross[k] = Gs.cross[k] = -full[k * KP + (k ^ 1)];
        for (int j = 0; j < KP; ++j)
            if ((j >> 1) != (k >> 1)) G.g2[((k >> 1) * KP + j) * 2 + (k & 1)] = -full[k * KP + j];
    }
    for (int m = 0; m < NPAIR; ++m)
        for (int n = m + 1; n < NPAIR; ++n) {
            float *blk = Gs.blk + 4 * bidx<NPAIR>(m, n);
            blk[0] = -full[(2 * m) * KP + 2 * n];
            blk[1] = -full[(2 * m + 1) * KP + 2 * n];
            blk[2] = -full[(2 * m) * KP + 2 * n + 1];
            blk[3] = -full[(2 * m + 1) * KP + 2 * n + 1];
        }
    float *d_state; long long *d_cyc;
    cudaMalloc(&d_state, (size_t)148 * 6 * 128 * KP * 4);
    cudaMalloc(&d_cyc, 148 * 6 * 8);
    std::vector<float> o0, o1;
    const size_t smem = (KP <= 48 ? 50 : 62) * 1024;
    launch(descent_bench<KP, PADC, 0>, G, d_state, d_cyc, K, cps, smem, o0, KP, "pair rows (constant)");
    launch(descent_sym<KP, PADC, NCONST, 0>, Gs, d_state, d_cyc, K, cps, smem, o1, KP, "sym, both uses indexed");
    std::vector<float> o2;
    launch(descent_sym<KP, PADC, NCONST, 1>, Gs, d_state, d_cyc, K, cps, smem, o2, KP, "sym, row-wise use indexed");
    double md2 = 0;
    for (size_t i = 0; i < o0.size(); ++i) md2 = fmax(md2, fabs((double)o0[i] - o2[i]));
    launch(descent_sym<KP, PADC, NCONST, 2>, Gs, d_state, d_cyc, K, cps, smem, o2, KP, "sym, row-wise use vector LDC");
    for (size_t i = 0; i < o0.size(); ++i) md2 = fmax(md2, fabs((double)o0[i] - o2[i]));
    printf("    mixed forms: max |difference| %.3g  (constant blocks %d of %d)\n", md2, NCONST < NPAIR * (NPAIR - 1) / 2 ? NCONST : NPAIR * (NPAIR - 1) / 2, NPAIR * (NPAIR - 1) / 2);
    double md = 0, mx = 0;
    for (size_t i = 0; i < o0.size(); ++i) { md = fmax(md, fabs((double)o0[i] - o1[i])); mx = fmax(mx, fabs((double)o0[i])); }
    printf("    max |difference| after 3 rounds %.3g (max |beta| %.3g)\n", md, mx);
    cudaFree(d_state); cudaFree(d_cyc);
}

int main()
{
    setvbuf(stdout, NULL, _IONBF, 0);
    compare<40, 0, 1 << 20>(40, 4);
    compare<56, 6, 224>(50, 3);
    compare<56, 6, 200>(50, 3);
    compare<48, 0, 200>(48, 4);
    return 0;
}
