// Shared pieces of the BCD sweep kernels (bcd.cu, bcd_p.cuh): Gram operands, packed-f32x2 helpers, tile layout,
// gather-plan layout.
#pragma once
#include <algorithm>
#include <type_traits>
#include <cuda_fp16.h>
#include <stdlib.h>
#include "fdb_common.cuh"
#include "bcd_state.cuh"

namespace fdb {

template <int KP>
struct alignas(16) GramArg {
    float g[KP * KP];               // g[k*KP+j] = -G[k][j] for j != k, 0 on the diagonal; zero padded
    float diag[KP];                 // G[k][k]
};

// Gram operand of the pair-step descent (bcd_sweep_p_kernel): rows 2m and 2m+1 interleaved column by column
template <int KP>
struct alignas(16) GramPairArg {
    float g2[KP * KP];              // g2[(m*KP + j)*2 + r] = -G[2m+r][j] for j outside {2m, 2m+1}, else 0; zero padded
    float cross[KP];                // cross[k] = -G[k][k^1]
    float diag[KP];                 // G[k][k]
};

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ void add4(float4 &a, const float4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
// packed fp32x2 FMA (sm_100+): d = a * b + d on both halves, one issue slot
__device__ __forceinline__ void ffma2(float2 &d, const float2 a, const float2 b)
{
    unsigned long long ua = *reinterpret_cast<const unsigned long long *>(&a);
    unsigned long long ub = *reinterpret_cast<const unsigned long long *>(&b);
    unsigned long long ud = *reinterpret_cast<unsigned long long *>(&d);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(ud) : "l"(ua), "l"(ub));
    d = *reinterpret_cast<float2 *>(&ud);
}
// 64-bit register-pair forms: keeping beta pairs and accumulators as b64 values lets ptxas hold them in
// aligned even/odd register pairs, so FFMA2 needs no MOVs to assemble its operands
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi)
{
    return (u64)__float_as_uint(lo) | ((u64)__float_as_uint(hi) << 32);      // folds to mov.b64 {lo, hi}
}
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi)
{
    lo = __uint_as_float((unsigned)v);
    hi = __uint_as_float((unsigned)(v >> 32));
}
__device__ __forceinline__ void ffma2q(u64 &d, const u64 a, const u64 b)
{
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ u64 add2q(const u64 a, const u64 b)
{
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float rcp_fast(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float2 fmul2(const float2 a, const float2 b)
{
    unsigned long long ua = *reinterpret_cast<const unsigned long long *>(&a);
    unsigned long long ub = *reinterpret_cast<const unsigned long long *>(&b);
    unsigned long long ud;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(ud) : "l"(ua), "l"(ub));
    return *reinterpret_cast<float2 *>(&ud);
}
// compile-time loop: f(integral_constant<int, I>) for I in [I0, N)
template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F &&f)
{
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}
__device__ __forceinline__ float elem(const float4 &v, int j) { return j == 0 ? v.x : j == 1 ? v.y : j == 2 ? v.z : v.w; }
__device__ __forceinline__ void set_elem(float4 &v, int j, float x)
{
    if (j == 0) v.x = x; else if (j == 1) v.y = x; else if (j == 2) v.z = x; else v.w = x;
}

constexpr int kIdxCap = 256;        // neighbour indices staged in shared memory per warp (32 spots)

template <int KP>
struct TileLayout {
    static constexpr int Q = KP / 4;                                      // float4 chunks per row
    static constexpr bool SWZ = (Q % 8 == 0);                             // XOR swizzle instead of padding
    static constexpr int S = SWZ ? KP : ((Q % 2 == 1) ? KP : KP + 4);     // floats per staged row
    __device__ static __forceinline__ int at(int row, int q)              // float offset of chunk q of a row
    {
        return SWZ ? row * S + 4 * (q ^ (row & 7)) : row * S + 4 * q;
    }
};

// 16-byte chunk swizzle of the fp16 gather rows: spreads rows that share a 128-byte bank window
template <int GQ>
__device__ __forceinline__ int gsw(int row)
{
    return GQ == 4 ? ((row >> 1) & 3) : (GQ == 2 ? ((row >> 2) & 1) : 0);
}

// ---- gather plan (built once per solve: the graph does not change between sweeps) ----------------------
// For every CTA patch of `TILE` spots: the out-of-patch neighbour rows it needs (deduplicated, at most HCAP
// = TILE of them) and, per neighbour reference in CSR order, a 16-bit code = row of the CTA's gather tile
// (patch row, or TILE + halo slot) or 0xFFFF when the patch needs more than HCAP foreign rows.
// The same codes are also kept per warp (32 consecutive rows), transposed and padded to the warp's largest
// degree, as bytes: codes8[warp][u][lane], u < max degree <= kCodeRounds, 512 bytes per warp -- round u of the
// gather reads one byte per lane, conflict-free, with no row-pointer arithmetic.  Byte values are gather-tile rows
// directly: 0..127 patch row, 128 + slot (slot < kHaloSlots = 126) halo row, 254 = the all-zero row (padding; it
// is halo slot 126, which is never assigned), 255 = slow (foreign row without a slot).
constexpr int kPlanHash = 1024;
constexpr unsigned short kCodeSlow = 0xFFFF;
constexpr int kHaloSlots = 126;                    // usable halo slots per 128-spot patch
constexpr int kCodeRounds = 16;                    // largest per-warp max degree with a transposed code block
constexpr int kCodeZero8 = 254, kCodeSlow8 = 255;

struct PlanView {
    const int32_t *halo_cnt;      // [n_ctas]
    const int32_t *halo_rows;     // [n_ctas * HCAP]
    const uint16_t *codes;        // [nnz], CSR order
    const uint8_t *codes8;        // [n_ctas * 4 warps * kCodeRounds * 32], transposed per warp
};

__host__ __device__ inline int64_t plan_off_cnt() { return 64; }
__host__ __device__ inline int64_t plan_off_rows(int64_t n_ctas) { return 64 + round_up(n_ctas * 4, 16); }
__host__ __device__ inline int64_t plan_off_codes8(int64_t n_ctas, int tile) { return round_up(plan_off_rows(n_ctas) + n_ctas * tile * 4, 512); }
__host__ __device__ inline int64_t plan_off_codes(int64_t n_ctas, int tile)
{
    return plan_off_codes8(n_ctas, tile) + n_ctas * (tile / 32) * (kCodeRounds * 32);
}

__device__ __forceinline__ u64 packm(float lo, float hi)                // explicit mov.b64: the pair is assembled once
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpackm(u64 v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 fma2(const u64 a, const u64 b, const u64 c)
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 mul2q(const u64 a, const u64 b)
{
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 sub2q(const u64 a, const u64 b)
{
    u64 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}


// Multi-GPU extension of the production sweep kernel (world <= 1: unused).
//   * The rank's boundary rows -- own rows that a neighbouring tile reads as halo -- are written straight into the
//     peers' beta_out buffers (NVLink peer memory) from the store phase of the patch that produced them.
//   * The inter-rank hand-shake (max norms for the stop test, core/solver.py:395-397, and "my rows of the last sweep
//     are in your halo slots") costs two system-scope round trips plus rank skew -- 10 + 16 us measured on 8 GPUs
//     against 23 us of compute for a 125k-spot tile.  It is therefore OVERLAPPED: the kernel of sweep t carries the
//     hand-shake of sweep t - 1 in one dedicated CTA (block 0), the other CTAs walk the INTERIOR patches (no peer
//     data needed) first, and only the patches that hold boundary rows -- last in the order -- wait for that CTA's
//     `hs_done` word (and for its verdict: a sweep found converged makes them stop, the speculative interior work is
//     simply never read).  A peer cannot overwrite halo rows this rank still reads: it pushes sweep t + 1 only after
//     its own hand-shake of sweep t, which needs this rank's flag, which this rank publishes at the start of its
//     kernel t + 1 -- after its kernel t has finished (stream order).  One launch per sweep, no NCCL, no host sync.
constexpr int kMaxRanks = 16;
constexpr int kCommStatWords = 4;                       // per (parity, rank): max|delta|, max|old|, max|new|, spare
constexpr int kCommWords = kMaxRanks + 2 * kMaxRanks * kCommStatWords;      // 144 words; the host reserves 256
struct SweepComm {
    const int32_t *patch_order;     // [n_patches] patches holding boundary rows first (walked backwards: interior first)
    const int32_t *n_boundary;      // device scalar: how many patches hold boundary rows
    const int32_t *push_ptr;        // [n_rows + 1] push entries of every own row
    const int2 *push_ent;           // (peer, destination row inside the peer's beta buffers)
    float *peer_base[kMaxRanks];    // symmetric buffers: [beta_a cap_rows*Kp][beta_b cap_rows*Kp][comm words]...
    long long out_off;              // float offset of beta_out inside a symmetric buffer
    long long comm_off;             // float offset of the comm block
    int rank, world;
    unsigned seq;                   // flag value of THIS launch's hand-shake ("sweep `sweep - 1` finished everywhere")
    int sweep;                      // 1-based index of the sweep this launch computes (max_iter + 1 for the closing launch)
    int hs_only;                    // closing launch: hand-shake of the last sweep, no patches
    int finalize_prev;              // the hand-shake closes a sweep (stop test, sweep counter)
    int debug;                      // timing experiments only: bit 0 = do not wait for the peers, bit 1 = do not push rows
};

inline PlanView plan_view(const void *plan, int64_t n_ctas, int tile)
{
    const char *pbase = (const char *)plan;
    PlanView pv;
    pv.halo_cnt = (const int32_t *)(pbase + plan_off_cnt());
    pv.halo_rows = (const int32_t *)(pbase + plan_off_rows(n_ctas));
    pv.codes = (const uint16_t *)(pbase + plan_off_codes(n_ctas, tile));
    pv.codes8 = (const uint8_t *)(pbase + plan_off_codes8(n_ctas, tile));
    return pv;
}

}  // namespace fdb
