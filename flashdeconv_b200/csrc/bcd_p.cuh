// Production BCD sweep kernel (persistent, software-pipelined, pair-step descent) and its launcher.  One translation
// unit per (row width, single-/multi-GPU form) instantiates launch_sweep_p<KP, COMM> (bcd_p_inst.cu, compiled once per
// FDB_P_KP x FDB_P_COMM), so the sixteen fully unrolled forms build in parallel; bcd.cu only declares the launcher.
#pragma once
#include "bcd_common.cuh"

namespace fdb {

// ------------------------------------------------------------------------------------
// Sweep kernel, persistent software-pipelined form (production; every K, rows padded to Kp = 8 ceil(K / 8)).
//
// bcd_sweep_h_kernel runs one patch per CTA, and because every CTA of a wave takes the same time the waves stay
// in lock-step: all resident CTAs wait on their start-of-patch DRAM round trips together (37 % of the warp time
// in the round-1 profile), then all compute together while HBM idles.  Here a CTA is persistent and walks
// patches blockIdx.x, blockIdx.x + gridDim.x, ...; everything patch p+1 needs is requested while patch p is
// being computed, with no extra shared memory for rows (residency stays at 6 CTAs/SM for Kp <= 32):
//   * row pointers and halo row ids of p+1: plain loads into registers, consumed one iteration later;
//   * neighbour codes of p+1: one 16-byte cp.async per lane into the alternate code buffer (after the gather of
//     p, when the row pointers have landed);
//   * beta_old rows, H rows and halo rows of p+1: prefetch.global.L2, so that the loads at the start of p+1 are
//     L2 hits (the prefetched-ahead footprint of the whole grid is ~32 MB of the 126 MB L2);
//   * H rows of p: cp.async into c_tile once the lane has its own beta_old row in registers, consumed after
//     the gather; c_tile then receives beta_new and is streamed out.
// Per patch: two CTA barriers (gather tile free / gather tile complete); statistics stay in registers until
// the CTA has finished all its patches.
// ------------------------------------------------------------------------------------
// Descent in PAIR STEPS: steps k = 2m and 2m+1 share one accumulator pair (part_2m, part_2m+1) that is fed, for every
// column j outside the pair, by ONE FFMA2 whose beta_j operand is a broadcast scalar register and whose Gram operand
// is the uniform pair (-G[2m][j], -G[2m+1][j]) (FFMA2 R, R.F32, UR.F32x2): Kp-2 FFMA2 per two steps with no
// horizontal reduction and no 64-bit re-packing of beta (it lives in Kp scalar registers); the two cross terms
// G[2m][2m+1] beta_old and G[2m+1][2m] beta_new are scalar FFMAs in the short serial tail.  Columns are walked from
// the least to the most recently updated one, so only the last links of the two chains wait for the previous pair.
// PADC = trailing all-padding columns (Kp - K >= PADC) left out at compile time.
template <int KP, int NW, int MINB, int PADC, bool COMM>
__global__ void __launch_bounds__(NW * 32, MINB)
bcd_sweep_p_kernel(const float *__restrict__ h, const __grid_constant__ GramPairArg<KP> G,
                   const float *__restrict__ beta_in, float *__restrict__ beta_out,
                   const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices, PlanView plan,
                   int n_rows, int n_types, float lam, float rho, float tol, int finalize, SolveState *state,
                   int n_patches, const __grid_constant__ SweepComm comm)
{
    static_assert(KP % 8 == 0, "half gather rows need Kp % 8 == 0");
    using L = TileLayout<KP>;
    constexpr int Q = L::Q, S = L::S, TILE = NW * 32, HCAP = TILE;
    constexpr int GQ = Q / 2;                            // 16-byte chunks per fp16 gather row
    constexpr int GROW = KP / 2;                         // 32-bit words per gather row
    extern __shared__ __align__(16) float sweep_smem[];
    float *c_tile = sweep_smem;                                                   // TILE x S fp32: beta_old, H, beta_new
    uint32_t *g_tile = reinterpret_cast<uint32_t *>(c_tile + TILE * S);           // (TILE + HCAP) x GROW words
    uint8_t *idx_tile = reinterpret_cast<uint8_t *>(g_tile + (TILE + HCAP) * GROW);     // NW x kCodeRounds x 32 byte codes
    int *scal = reinterpret_cast<int *>(idx_tile + NW * kCodeRounds * 32);        // 3 x TILE: row start, end, halo id
    __shared__ unsigned red[COMM ? 3 : 2][NW];
    __shared__ int s_flag;

    if (*reinterpret_cast<volatile int *>(&state->converged)) return;      // uniform across the grid (and across ranks)
    // COMM = false (single GPU) compiles the multi-GPU extension out: the loop below is the round-1 kernel unchanged
    const bool comm_on = COMM && comm.world > 1;

    // ---------------- multi-GPU: block 0 carries the hand-shake of the PREVIOUS sweep and nothing else
    if (comm_on && blockIdx.x == 0) {
        const int pp = (comm.sweep + 1) & 1;                                // parity of the sweep being closed
        const int pn3 = (comm.sweep + 2) % 3;                              // its slot of max|beta_new| ((sweep - 1) mod 3)
        __shared__ int s_timeout;
        if (threadIdx.x == 0) s_timeout = 0;
        __syncthreads();
        if ((int)threadIdx.x < comm.world) {
            const int peer = threadIdx.x;
            unsigned *pc = reinterpret_cast<unsigned *>(comm.peer_base[peer] + comm.comm_off);   // the peer's comm block
            unsigned *slot = pc + kMaxRanks + (pp * kMaxRanks + comm.rank) * kCommStatWords;
            // plain stores and ONE release: st.release.sys orders them -- and every write of the previous launches on
            // this stream (the sweep's pushes) -- before the flag
            slot[0] = *reinterpret_cast<volatile unsigned *>(&state->ov_diff[pp]);
            slot[1] = *reinterpret_cast<volatile unsigned *>(&state->ov_abs[pp]);
            slot[2] = *reinterpret_cast<volatile unsigned *>(&state->ov_new[pn3]);
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pc + comm.rank), "r"(comm.seq) : "memory");
            const unsigned *mine = reinterpret_cast<const unsigned *>(comm.peer_base[comm.rank] + comm.comm_off);
            const long long t0 = clock64();
            for (;;) {
                unsigned v;
                asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine + peer) : "memory");
                if ((int)(v - comm.seq) >= 0 || (comm.debug & 1)) break;
                if (clock64() - t0 > 120000000000LL) { s_timeout = 1; break; }                   // ~60 s: a peer is gone
            }
            asm volatile("fence.acq_rel.sys;" ::: "memory");      // the peers' rows and norm words are visible from here on
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (s_timeout) {
                state->converged = 2;                                                            // surfaced as an error by the host
            } else {
                const unsigned *mine = reinterpret_cast<const unsigned *>(comm.peer_base[comm.rank] + comm.comm_off);
                unsigned md = 0u, ma = 0u, mn = 0u;
                for (int p = 0; p < comm.world; ++p) {
                    const volatile unsigned *sl = mine + kMaxRanks + (pp * kMaxRanks + p) * kCommStatWords;
                    md = max(md, sl[0]);
                    ma = max(ma, sl[1]);
                    mn = max(mn, sl[2]);
                }
                if (comm.finalize_prev) {
                    state->max_diff_bits = md;
                    state->max_abs_bits = ma;
                    finalize_state(state, tol);
                }
                state->last_max_abs = __uint_as_float(mn);        // max over ALL ranks of |beta| after the closed sweep
                state->last_max_diff = 0.f;
                state->ov_diff[pp] = 0u;
                state->ov_abs[pp] = 0u;
                state->ov_new[(comm.sweep + 1) % 3] = 0u;         // the slot the NEXT sweep accumulates into
            }
            __threadfence();
            *reinterpret_cast<volatile unsigned *>(&state->hs_done) = comm.seq;
        }
        return;
    }
    const int n_boundary = comm_on ? __ldg(comm.n_boundary) : 0;
    const int n_workers = comm_on ? (int)gridDim.x - 1 : (int)gridDim.x;
    const int worker = comm_on ? (int)blockIdx.x - 1 : (int)blockIdx.x;
    // multi-GPU: the order array lists the boundary patches first; it is walked BACKWARDS (interior patches first)
    auto patch_at = [&](int pi) { return comm_on ? __ldg(comm.patch_order + (n_patches - 1 - pi)) : pi; };
    const int first_boundary = n_patches - n_boundary;    // positions >= this hold boundary rows (multi-GPU only)

    // Range of the fp16 gather tile.  With 2^x <= bound < 2^(x+1) on the magnitude of every beta_in the tile stores
    // beta * 2^(8-x): magnitudes below 512, sums of up to 64 neighbours below 32768 < 65504, and values down to
    // 1e-7 of the largest one keep all 11 bits.  Powers of two: results are bit-identical to an unscaled tile whenever
    // that one neither overflows nor underflows.  lam_s folds the factor back in.
    // Single GPU: bound = last_max_abs + last_max_diff (max|beta| of the sweep before plus its largest step;
    // fdb_bcd_init seeds 1/K).  Multi-GPU: interior patches gather this rank's rows only, so the rank's own
    // max|beta_new| of the previous sweep bounds them; boundary patches use the all-rank maximum the hand-shake left
    // in last_max_abs (they wait for it anyway).
    float inv_s, lam_s;
    auto set_scale = [&](float bound) {
        const int ef = min(max((int)((__float_as_uint(bound) >> 23) & 255u), 9), 245);
        inv_s = __uint_as_float((unsigned)(262 - ef) << 23);
        lam_s = lam * __uint_as_float((unsigned)(ef - 8) << 23);
    };
    if (comm_on) set_scale(__uint_as_float(*reinterpret_cast<volatile unsigned *>(&state->ov_new[(comm.sweep + 2) % 3])));
    else set_scale(*reinterpret_cast<volatile float *>(&state->last_max_abs) +
                   *reinterpret_cast<volatile float *>(&state->last_max_diff));

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wrow = warp * 32;
    const int own = wrow + lane;
    uint8_t *iw = idx_tile + warp * (kCodeRounds * 32) + lane;

    // asynchronous copy of the warp's 32 rows of `src` (patch `pp`) into c_tile
    auto rows_async = [&](const float *__restrict__ src, int pp) {
#pragma unroll
        for (int i = 0; i < Q; ++i) {
            const int idx = lane + 32 * i;
            const int lr = idx / Q, q = idx - lr * Q;
            const int p = pp * TILE + wrow + lr;
            const uint32_t d = (uint32_t)__cvta_generic_to_shared(c_tile + L::at(wrow + lr, q));
            const int nbytes = p < n_rows ? 16 : 0;                               // rows past the end: zero fill
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d),
                         "l"(src + (size_t)min(p, n_rows - 1) * KP + 4 * q), "r"(nbytes));
        }
    };
    // this thread's row pointers and halo row id of patch `pp` -> its three private words of `scal`
    // (asynchronously: nothing is held in registers while the current patch is computed)
    auto scalars_async = [&](int pp) {
        const int r = pp * TILE + own;
        const int nb = r < n_rows ? 4 : 0;
        const int32_t *src = indptr + min(r, n_rows - 1);
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(scal + threadIdx.x);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(nb));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d + 4 * TILE), "l"(src + 1), "r"(nb));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 8 * TILE),
                     "l"(plan.halo_rows + (size_t)pp * HCAP + threadIdx.x));    // slot = thread, -1 = unused
    };

    // ---------------- prologue: the first patch's scalars
    // (single GPU: `patch` is the only induction variable, exactly the round-1 loop -- at the 80-register cap one more
    // live value makes ptxas rematerialise address arithmetic all over the loop body: +4 % instructions, +8 % time)
    int pi = COMM ? worker : 0;                           // multi-GPU: position in the processing order
    int patch = COMM ? (pi < n_patches ? patch_at(pi) : n_patches) : (int)blockIdx.x;
    if (patch < n_patches) scalars_async(patch);
    asm volatile("cp.async.commit_group;");
    asm volatile("cp.async.wait_all;");

    float dmax = 0.f, amax = 0.f, nmax = 0.f;
    bool waited = false;
#pragma unroll 1
    for (; COMM ? pi < n_patches : patch < n_patches;) {
        if (COMM && comm_on && pi >= first_boundary && !waited) {
            // first patch with boundary rows: the hand-shake of the previous sweep has to be complete (peers' rows of
            // that sweep are in the halo slots; its stop test decides whether this sweep exists at all)
            if (threadIdx.x == 0) {
                const long long t0 = clock64();
                while (*reinterpret_cast<volatile unsigned *>(&state->hs_done) != comm.seq &&
                       !*reinterpret_cast<volatile int *>(&state->converged)) {
                    if (clock64() - t0 > 130000000000LL) { state->converged = 2; break; }       // the hand-shake never came
                    __nanosleep(200);
                }
                __threadfence();
                s_flag = *reinterpret_cast<volatile int *>(&state->converged);
            }
            __syncthreads();
            if (s_flag) break;                             // the previous sweep converged (or a peer is gone): stop here
            set_scale(*reinterpret_cast<volatile float *>(&state->last_max_abs));
            waited = true;
        }
        const int tile_base = patch * TILE;
        const int next = COMM ? (pi + n_workers < n_patches ? patch_at(pi + n_workers) : n_patches) : patch + (int)gridDim.x;
        const int my_row = tile_base + own;
        const int my_s = scal[threadIdx.x], my_e = scal[TILE + threadIdx.x];
        const int halo_id = scal[2 * TILE + threadIdx.x];
        const int my_deg = my_e - my_s;
        // the warp's transposed byte codes: 32 bytes per gather round
        const int maxdeg = __reduce_max_sync(kFull, my_deg);
        const bool staged = maxdeg <= kCodeRounds;
        if (staged && lane < 2 * maxdeg) {
            const uint32_t d = (uint32_t)__cvta_generic_to_shared(iw + 15 * lane);          // iw + lane + 15 lane = 16-byte chunk
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d),
                         "l"(plan.codes8 + ((size_t)patch * NW + warp) * (kCodeRounds * 32) + 16 * lane));
        }
        // the warp's beta_old rows -> c_tile (its rows of c_tile were streamed out at the end of the previous patch)
        rows_async(beta_in, patch);
        asm volatile("cp.async.commit_group;");
        // the patch's halo rows (ids travel by shuffle); like everything above these are L2 hits: the lines were
        // prefetched while the previous patch was computed
        float4 hrow[Q];
#pragma unroll
        for (int i = 0; i < Q; ++i) {
            const int idx = lane + 32 * i;
            const int lr = idx / Q, q = idx - lr * Q;
            const int g = __shfl_sync(kFull, halo_id, lr);
            hrow[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g >= 0) hrow[i] = ld4(beta_in + (size_t)g * KP + 4 * q);
        }
        asm volatile("cp.async.wait_all;");
        __syncthreads();                 // (1) every warp is past the gather of the previous patch: g_tile is free
        auto to_gather = [&](int grow, int q, const float4 bb) {
            const __half2 lo = __floats2half2_rn(bb.x * inv_s, bb.y * inv_s), hi = __floats2half2_rn(bb.z * inv_s, bb.w * inv_s);
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t *>(&lo);
            pk.y = *reinterpret_cast<const uint32_t *>(&hi);
            *reinterpret_cast<uint2 *>(g_tile + grow * GROW + 4 * ((q >> 1) ^ gsw<GQ>(grow)) + 2 * (q & 1)) = pk;
        };
#pragma unroll
        for (int i = 0; i < Q; ++i) {
            const int idx = lane + 32 * i;
            const int lr = idx / Q, q = idx - lr * Q;
            to_gather(TILE + wrow + lr, q, hrow[i]);
            to_gather(wrow + lr, q, ld4(c_tile + L::at(wrow + lr, q)));
        }
        // own beta_old row -> registers (fp32 scalars)
        float b[KP];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const float4 b4 = ld4(c_tile + L::at(own, q));
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(b4.x), fabsf(b4.y)), fmaxf(fabsf(b4.z), fabsf(b4.w))));
            b[4 * q] = b4.x; b[4 * q + 1] = b4.y; b[4 * q + 2] = b4.z; b[4 * q + 3] = b4.w;
        }
        __syncthreads();                 // (2) gather tile complete; the warp's fp32 rows are consumed

        // ---------------- requests: H rows of this patch -> c_tile; the next patch's scalars -> scal, rows -> L2
        rows_async(h, patch);
        if (next < n_patches) {
            scalars_async(next);
            const size_t base = (size_t)next * TILE * KP;
            const int lines = min(TILE, n_rows - next * TILE) * KP / 32;       // 128-byte lines of the patch's rows
            for (int l = threadIdx.x; l < lines; l += TILE) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(beta_in + base + (size_t)l * 32));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(h + base + (size_t)l * 32));
            }
        }
        asm volatile("cp.async.commit_group;");

        // ---------------- neighbour sums from the fp16 gather tile, one spot per lane.  Round u adds, for every
        // lane, the gather-tile row named by its u-th byte code (padding codes name the all-zero row 254).
        __half2 acc[KP / 2];
        {
#pragma unroll
            for (int i = 0; i < KP / 2; ++i) acc[i] = __floats2half2_rn(0.f, 0.f);
            auto add_row = [&](int grow) {
                const uint4 *row = reinterpret_cast<const uint4 *>(g_tile + grow * GROW);
                const int sw = gsw<GQ>(grow);
#pragma unroll
                for (int q = 0; q < GQ; ++q) {
                    const uint4 w = row[q ^ sw];
                    acc[4 * q] = __hadd2(*reinterpret_cast<const __half2 *>(&w.x), acc[4 * q]);
                    acc[4 * q + 1] = __hadd2(*reinterpret_cast<const __half2 *>(&w.y), acc[4 * q + 1]);
                    acc[4 * q + 2] = __hadd2(*reinterpret_cast<const __half2 *>(&w.z), acc[4 * q + 2]);
                    acc[4 * q + 3] = __hadd2(*reinterpret_cast<const __half2 *>(&w.w), acc[4 * q + 3]);
                }
            };
            auto add_slow = [&](int u) {                 // foreign row without a halo slot: fp32 row from global
                const float *src = beta_in + (size_t)__ldg(indices + my_s + u) * KP;
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const float4 v = ld4(src + 4 * q);
                    acc[2 * q] = __hadd2(acc[2 * q], __floats2half2_rn(v.x * inv_s, v.y * inv_s));
                    acc[2 * q + 1] = __hadd2(acc[2 * q + 1], __floats2half2_rn(v.z * inv_s, v.w * inv_s));
                }
            };
            if (staged) {
                int code = maxdeg > 0 ? iw[0] : kCodeZero8;
#pragma unroll 1
                for (int u = 0; u < maxdeg; ++u) {
                    const int cur = code;
                    code = iw[min(u + 1, kCodeRounds - 1) * 32];            // next round's code (stale past the end)
                    if (__any_sync(kFull, cur == kCodeSlow8)) {              // rare
                        if (cur == kCodeSlow8) add_slow(u);
                    }
                    add_row(cur == kCodeSlow8 ? kCodeZero8 : cur);
                }
            } else {                                     // a row with more than kCodeRounds neighbours: CSR-order codes
#pragma unroll 1
                for (int u = 0; u < maxdeg; ++u) {
                    unsigned code = kCodeZero8;
                    if (u < my_deg) code = plan.codes[my_s + u];
                    if (code == kCodeSlow) { add_slow(u); code = kCodeZero8; }
                    add_row((int)code);
                    if ((u & 63) == 63 && u + 1 < maxdeg) {
                        // 64 neighbours summed: move the half-precision partial sums into the fp32 H row (which has to have
                        // landed first) so that very high degrees cannot overflow the accumulator
                        asm volatile("cp.async.wait_all;");
                        __syncwarp();
#pragma unroll
                        for (int q = 0; q < Q; ++q) {
                            float4 c4 = ld4(c_tile + L::at(own, q));
                            const float2 s01 = __half22float2(acc[2 * q]), s23 = __half22float2(acc[2 * q + 1]);
                            c4.x = fmaf(lam_s, s01.x, c4.x); c4.y = fmaf(lam_s, s01.y, c4.y);
                            c4.z = fmaf(lam_s, s23.x, c4.z); c4.w = fmaf(lam_s, s23.y, c4.w);
                            st4(c_tile + L::at(own, q), c4);
                            acc[2 * q] = __floats2half2_rn(0.f, 0.f);
                            acc[2 * q + 1] = __floats2half2_rn(0.f, 0.f);
                        }
                    }
                }
            }
        }
        asm volatile("cp.async.wait_all;");              // H rows of this patch, scalars of the next
        __syncwarp();
        // the next patch's halo rows and code slice -> L2
        if (next < n_patches) {
            const int nx_halo = scal[2 * TILE + threadIdx.x];
            if (nx_halo >= 0) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(beta_in + (size_t)nx_halo * KP));
                if (KP > 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(beta_in + (size_t)nx_halo * KP + 32));
            }
            if (lane < 4)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(plan.codes8 + ((size_t)next * NW + warp) * (kCodeRounds * 32) + 128 * lane));
        }

        // ---------------- cyclic coordinate descent in pair steps (see the kernel header):
        //   part_k = c_k - rho - sum_{j != k} G_kj b_j  (b_j already updated for j < k),  b_k <- max(0, part_k / den_k)
        // Padding columns need no special case: their H, beta and Gram entries are 0, so they stay exactly 0.
        {
            const float lam_deg = lam * (float)my_deg;
            const float neg_rho = -rho;
            float dm = 0.f;
            static_for<0, Q>([&](auto qc) {
                constexpr int q = decltype(qc)::value;
                if constexpr (4 * q >= KP - PADC) return;
                const float4 c4 = ld4(c_tile + L::at(own, q));
                const float2 s01 = __half22float2(acc[2 * q]), s23 = __half22float2(acc[2 * q + 1]);
                const float4 ns4 = make_float4(s01.x, s01.y, s23.x, s23.y);
                static_for<0, 2>([&](auto mc) {
                    constexpr int m = 2 * q + decltype(mc)::value;         // pair index
                    constexpr int k0 = 2 * m;
                    if constexpr (k0 >= KP - PADC) return;
                    if (k0 >= KP - 8 && k0 >= n_types) return;              // a pair of padding columns (warp-uniform)
                    const float den0 = G.diag[k0] + lam_deg, den1 = G.diag[k0 + 1] + lam_deg;
                    const float ri0 = den0 > 1e-10f ? rcp_fast(den0) : 0.f;      // core/solver.py:87-90
                    const float ri1 = den1 > 1e-10f ? rcp_fast(den1) : 0.f;
                    u64 a0 = pack2(fmaf(lam_s, elem(ns4, k0 & 3), elem(c4, k0 & 3)),
                                   fmaf(lam_s, elem(ns4, (k0 & 3) + 1), elem(c4, (k0 & 3) + 1)));
                    u64 a1 = pack2(neg_rho, neg_rho);
                    static_for<2, KP>([&](auto ic) {
                        constexpr int i = decltype(ic)::value;
                        constexpr int jj = (k0 + i) % KP;                   // k0+2, ..., Kp-1, 0, ..., k0-1
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
                    if (COMM) nmax = fmaxf(nmax, nv0);
                    dm = fmaxf(dm, fabsf(nv0 - b[k0]));
                    b[k0] = nv0;
                    p1 = fmaf(G.cross[k0 + 1], nv0, p1);
                    const float nv1 = fmaxf(0.f, p1 * ri1);
                    if (COMM) nmax = fmaxf(nmax, nv1);
                    dm = fmaxf(dm, fabsf(nv1 - b[k0 + 1]));
                    b[k0 + 1] = nv1;
                });
                st4(c_tile + L::at(own, q), make_float4(b[4 * q], b[4 * q + 1], b[4 * q + 2], b[4 * q + 3]));
            });
            if (my_row < n_rows) dmax = fmaxf(dmax, dm);
        }
        __syncwarp();

        // ---------------- stream the warp's new rows out
#pragma unroll
        for (int i = 0; i < Q; ++i) {
            const int idx = lane + 32 * i;
            const int lr = idx / Q, q = idx - lr * Q;
            const int p = tile_base + wrow + lr;
            if (p < n_rows) st4(beta_out + (size_t)p * KP + 4 * q, ld4(c_tile + L::at(wrow + lr, q)));
        }
        // ---------------- boundary rows -> the neighbouring tiles' halo slots (peer memory), from registers
        if (COMM && comm_on && pi >= first_boundary && my_row < n_rows && !(comm.debug & 2)) {
            const int pe = __ldg(comm.push_ptr + my_row + 1);
            for (int u = __ldg(comm.push_ptr + my_row); u < pe; ++u) {
                const int2 ent = __ldg(comm.push_ent + u);
                float *dst = comm.peer_base[ent.x] + comm.out_off + (long long)ent.y * KP;
#pragma unroll
                for (int q = 0; q < Q; ++q)
                    st4(dst + 4 * q, make_float4(b[4 * q], b[4 * q + 1], b[4 * q + 2], b[4 * q + 3]));
            }
            // this thread's peer writes are ordered before the CTA's arrival (barrier + thread 0's fence below) and, through
            // the arrival counter and the last CTA's release, before the flag the peers acquire
            if (pe > __ldg(comm.push_ptr + my_row)) __threadfence_system();
        }
        if (COMM) { patch = next; pi += n_workers; }
        else patch += gridDim.x;
    }

    const unsigned wd = __reduce_max_sync(kFull, __float_as_uint(dmax));
    const unsigned wa = __reduce_max_sync(kFull, __float_as_uint(amax));
    const unsigned wn = COMM ? __reduce_max_sync(kFull, __float_as_uint(nmax)) : 0u;
    if (lane == 0) {
        red[0][warp] = wd;
        red[1][warp] = wa;
        if (COMM) red[COMM ? 2 : 0][warp] = wn;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned bd = 0u, ba = 0u, bn = 0u;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            bd = max(bd, red[0][w]);
            ba = max(ba, red[1][w]);
            if (COMM) bn = max(bn, red[COMM ? 2 : 0][w]);
        }
        if (comm_on) {
            // overlapped mode: this sweep's norms go to their parity slot; the NEXT launch's hand-shake block closes the sweep
            const int p = comm.sweep & 1;
            if (bd > *reinterpret_cast<volatile unsigned *>(&state->ov_diff[p])) atomicMax(&state->ov_diff[p], bd);
            if (ba > *reinterpret_cast<volatile unsigned *>(&state->ov_abs[p])) atomicMax(&state->ov_abs[p], ba);
            if (bn > *reinterpret_cast<volatile unsigned *>(&state->ov_new[comm.sweep % 3])) atomicMax(&state->ov_new[comm.sweep % 3], bn);
            return;
        }
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


// host side: Gram operand in pair-row layout, plan view, residency-sized persistent grid
template <int KP, bool COMM>
int launch_sweep_p(const float *h, const GramArg<KP> &G, int n_types, const float *beta_in, float *beta_out,
                   const int32_t *indptr, const int32_t *indices, int64_t n_rows, float lam, float rho, float tol,
                   int finalize, SolveState *state, const void *plan, cudaStream_t st, const SweepComm *comm)
{
    // COMM = (comm != nullptr): the caller picks the instantiation
    // 128-spot patches (4 warps); residency by row width (shared memory): 6 CTAs/SM at 36 KB (Kp <= 32), 4 at 46-54 KB
    // (Kp = 40, 48), 3 at 62-68 KB (Kp = 56, 64)
    constexpr int NWH = 4;
    constexpr int MINB = KP <= 32 ? 6 : (KP <= 48 ? 4 : 3);
    constexpr int tile = NWH * 32;
    const int64_t n_ctas = ceil_div(n_rows, tile);
    const PlanView pv = plan_view(plan, n_ctas, tile);
    GramPairArg<KP> P;
    for (int i = 0; i < KP * KP; ++i) P.g2[i] = 0.f;
    for (int k = 0; k < KP; ++k) {
        P.diag[k] = G.diag[k];
        P.cross[k] = G.g[k * KP + (k ^ 1)];
        for (int j = 0; j < KP; ++j)
            if ((j >> 1) != (k >> 1)) P.g2[((k >> 1) * KP + j) * 2 + (k & 1)] = G.g[k * KP + j];
    }
    const size_t smem = (size_t)tile * TileLayout<KP>::S * 4 + (size_t)2 * tile * (KP / 2) * 4 +
                        (size_t)NWH * kCodeRounds * 32 + (size_t)3 * tile * 4;
    SweepComm cm;
    if (comm) cm = *comm;
    else {
        cm = SweepComm();
        cm.patch_order = nullptr; cm.n_boundary = nullptr; cm.push_ptr = nullptr; cm.push_ent = nullptr;
        for (int p = 0; p < kMaxRanks; ++p) cm.peer_base[p] = nullptr;
        cm.out_off = cm.comm_off = 0; cm.rank = 0; cm.world = 1; cm.seq = 0u; cm.sweep = 1; cm.hs_only = 0; cm.finalize_prev = 0; cm.debug = 0;
    }
    auto run_p = [&](auto kern) -> int {
        int resident = 0;
        FDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, tile, smem));
        // FDB_SWEEP_MAX_CTAS caps the persistent grid (tests: several patches per CTA on small problems)
        static const int cap = getenv("FDB_SWEEP_MAX_CTAS") ? std::max(atoi(getenv("FDB_SWEEP_MAX_CTAS")), 1) : 1 << 30;
        // persistent grid: as many CTAs as are resident, but balanced -- with R = ceil(patches / slots) rounds every CTA
        // walks R (or R - 1) patches instead of leaving a ragged last round to a few CTAs (matters when a rank's tile
        // is barely more than one round: 977 patches on 888 slots at 8 GPUs)
        const int64_t slots = (int64_t)kNumSM * std::max(resident, 1);
        const int64_t rounds = std::max<int64_t>(ceil_div(n_ctas, slots), 1);
        // (only for short walks: with many rounds a full grid keeps every SM slot busy, which is worth more than balance)
        int grid = (int)std::min<int64_t>(std::min<int64_t>(n_ctas, cap), rounds <= 3 ? ceil_div(n_ctas, rounds) : slots);
        if (comm != nullptr) {                              // block 0 = hand-shake; the workers leave it a slot
            const int64_t wslots = std::max<int64_t>(slots - 1, 1);
            const int64_t wrounds = std::max<int64_t>(ceil_div(n_ctas, wslots), 1);
            grid = 1 + (cm.hs_only ? 0 : (int)std::min<int64_t>(std::min<int64_t>(n_ctas, cap), ceil_div(n_ctas, wrounds)));
        }
        kern<<<grid, tile, smem, st>>>(h, P, beta_in, beta_out, indptr, indices, pv, (int)n_rows, n_types, lam, rho, tol,
                                       finalize, state, cm.hs_only ? 0 : (int)n_ctas, cm);
        FDB_LAUNCH_CHECK("bcd_sweep_p_kernel");
        return FDB_OK;
    };
    // trailing padding columns (Kp - K, rounded down to even) are left out at compile time
    const int pad = KP - n_types;
    if (pad >= 6) return run_p(bcd_sweep_p_kernel<KP, NWH, MINB, 6, COMM>);
    if (pad >= 4) return run_p(bcd_sweep_p_kernel<KP, NWH, MINB, 4, COMM>);
    if (pad >= 2) return run_p(bcd_sweep_p_kernel<KP, NWH, MINB, 2, COMM>);
    return run_p(bcd_sweep_p_kernel<KP, NWH, MINB, 0, COMM>);
}

}  // namespace fdb
