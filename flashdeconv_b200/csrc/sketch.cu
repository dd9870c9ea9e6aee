// Kernel family 1/2: log-CPM + leverage-weighted CountSketch of the spot-by-gene CSR,
// and the thin contraction against the sketched reference.
//
// Reference semantics (upstream file:line):
//   core/deconv.py:321        Y_subset = Y[:, gene_idx]   -> here: gene_bucket[g] < 0 masks a gene out
//   core/deconv.py:183-188    lib = row sum over selected genes, 0 -> 1;  y~ = log1p(1e4/lib * y)
//   core/sketching.py:195     Y_s = Y~ @ Omega, Omega has ONE entry per gene (bucket, weight)
//   core/solver.py:223,348    H = X_s Y_s^T,  YtY = sum Y_s^2
//
// Mapping: one warp per spot (CSR row).  The row is streamed once with coalesced loads that
// bypass L1 (read-once data) while the per-gene tables stay L1-resident; a register cache
// keeps (bucket, count, weight) of the first 32*kCache entries so the second pass (which
// needs the library size) does not touch memory again.  The 512-bucket accumulator is a
// per-warp shared-memory array updated with shared atomics (fp32 RED).
//
// HBM-bound: algorithmic bytes = 8*nnz + 4*(N+1) read, + 4*d*N written (materialising form)
// or 4*(Kp+1)*N written (fused form).
#include <stdlib.h>
#include <algorithm>
#include "fdb_common.cuh"

namespace fdb {

// value transform of the fused kernels: scale > 0 -> log1p(v * scale) (log-CPM, core/deconv.py:177-197); scale <= 0 ->
// v itself (preprocess "raw" / "pearson": a per-gene factor folded into the gene weights, core/deconv.py:199-229)
__device__ __forceinline__ float xform_value(float v, float scale) { return scale > 0.f ? log1pf(v * scale) : v; }


constexpr int kCache = 16;   // register-cached chunks of 32 entries per row

struct RowCache {
    int b[kCache];
    float v[kCache];
    float w[kCache];
};

// Pass 1 of a row: library size over selected genes + fill the register cache.
// Entries past 32*kCache are summed here and re-read by the caller's tail loop.
__device__ __forceinline__ float row_pass1(const int32_t *__restrict__ indices,
                                           const float *__restrict__ counts,
                                           const int32_t *__restrict__ gene_bucket,
                                           const float *__restrict__ gene_weight, int64_t s,
                                           int64_t e, int lane, RowCache &rc)
{
    float lib = 0.f;
#pragma unroll
    for (int c = 0; c < kCache; ++c) {
        const int64_t j = s + lane + 32 * c;
        int b = -1;
        float v = 0.f, w = 0.f;
        if (j < e) {
            const int g = ld_stream(indices + j);
            v = ld_stream(counts + j);
            b = __ldg(gene_bucket + g);
            if (b >= 0) {
                w = __ldg(gene_weight + g);
                lib += v;
            }
        }
        rc.b[c] = b;
        rc.v[c] = v;
        rc.w[c] = w;
    }
    for (int64_t j = s + lane + 32 * kCache; j < e; j += 32) {
        const int g = ld_stream(indices + j);
        if (__ldg(gene_bucket + g) >= 0) lib += ld_stream(counts + j);
    }
    lib = warp_sum(lib);
    return lib == 0.f ? 1.f : lib;
}

// ------------------------------------------------------------------------------------
// materialising form: writes Y_s (N x d)
// ------------------------------------------------------------------------------------
template <typename IndPtr, bool LOGCPM>
__global__ void __launch_bounds__(512)
sketch_rows_kernel(const IndPtr *__restrict__ indptr, const int32_t *__restrict__ indices,
                   const float *__restrict__ counts, int64_t n_spots,
                   const int32_t *__restrict__ gene_bucket, const float *__restrict__ gene_weight,
                   int d, float *__restrict__ y_sketch)
{
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warps_per_cta = blockDim.x >> 5;
    float *acc = smem + (size_t)warp * d;
    for (int c = lane; c < d; c += 32) acc[c] = 0.f;
    __syncwarp();

    RowCache rc;
    for (int64_t row = (int64_t)blockIdx.x * warps_per_cta + warp; row < n_spots;
         row += (int64_t)gridDim.x * warps_per_cta) {
        const int64_t s = load_ptr(indptr, row), e = load_ptr(indptr, row + 1);
        const float scale = 1e4f / row_pass1(indices, counts, gene_bucket, gene_weight, s, e, lane, rc);
        auto xform = [&](float v) { return LOGCPM ? log1pf(v * scale) : v; };
#pragma unroll
        for (int c = 0; c < kCache; ++c)
            if (rc.b[c] >= 0) atomicAdd(acc + rc.b[c], xform(rc.v[c]) * rc.w[c]);
        for (int64_t j = s + lane + 32 * kCache; j < e; j += 32) {
            const int g = ld_stream(indices + j);
            const int b = __ldg(gene_bucket + g);
            if (b >= 0) atomicAdd(acc + b, xform(ld_stream(counts + j)) * __ldg(gene_weight + g));
        }
        __syncwarp();
        float *out = y_sketch + row * (int64_t)d;
        for (int c = lane * 4; c < d; c += 128) {            // d % 4 == 0 checked on the host
            const float4 val = *reinterpret_cast<float4 *>(acc + c);
            *reinterpret_cast<float4 *>(acc + c) = make_float4(0.f, 0.f, 0.f, 0.f);
            __stcs(reinterpret_cast<float4 *>(out + c), val);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------
// fused form: H[i,:] = sum_e c_e * X_s^T[bucket_e, :]  (no Y_s), ysq_i = sum_b acc_b^2
// NK = ceil(Kp / 32): lane owns types lane (+32).
// ------------------------------------------------------------------------------------
template <typename IndPtr, int NK>
__global__ void __launch_bounds__(512)
sketch_contract_kernel(const IndPtr *__restrict__ indptr, const int32_t *__restrict__ indices,
                       const float *__restrict__ counts, int64_t n_spots,
                       const int32_t *__restrict__ gene_bucket,
                       const float *__restrict__ gene_weight, int d,
                       const float *__restrict__ x_sketch_t, int kp,
                       const int32_t *__restrict__ row_map, const int32_t *__restrict__ row_ids,
                       float *__restrict__ h, float *__restrict__ ysq, int linear)
{
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warps_per_cta = blockDim.x >> 5;
    const int xrow = NK * 32;                       // smem row stride of X_s^T (padded, zero filled)
    float *xs = smem;                               // d x xrow
    float *acc = smem + (size_t)d * xrow + (size_t)warp * d;
    for (int i = threadIdx.x; i < d * xrow; i += blockDim.x) {
        const int r = i / xrow, c = i - r * xrow;
        xs[i] = c < kp ? __ldg(x_sketch_t + (size_t)r * kp + c) : 0.f;
    }
    for (int c = lane; c < d; c += 32) acc[c] = 0.f;
    __syncthreads();

    RowCache rc;
    for (int64_t it = (int64_t)blockIdx.x * warps_per_cta + warp; it < n_spots;
         it += (int64_t)gridDim.x * warps_per_cta) {
        const int64_t row = row_ids ? (int64_t)__ldg(row_ids + it) : it;       // input row processed by this warp
        const int64_t s = load_ptr(indptr, row), e = load_ptr(indptr, row + 1);
        const float lib1 = row_pass1(indices, counts, gene_bucket, gene_weight, s, e, lane, rc);
        const float scale = linear ? -1.f : 1e4f / lib1;
        float h0 = 0.f, h1 = 0.f;
        auto consume = [&](int b, float c) {
            // every lane calls this with its own (b, c); selected entries are broadcast one by one
            if (b >= 0) atomicAdd(acc + b, c);
            unsigned m = __ballot_sync(kFull, b >= 0);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const float cb = __shfl_sync(kFull, c, src);
                const int bb = __shfl_sync(kFull, b, src);
                h0 = fmaf(cb, xs[bb * xrow + lane], h0);
                if (NK == 2) h1 = fmaf(cb, xs[bb * xrow + 32 + lane], h1);
            }
        };
#pragma unroll
        for (int c = 0; c < kCache; ++c) {
            if (s + 32 * c >= e) break;             // warp-uniform
            consume(rc.b[c], rc.b[c] >= 0 ? xform_value(rc.v[c], scale) * rc.w[c] : 0.f);
        }
        for (int64_t j0 = s + 32 * kCache; j0 < e; j0 += 32) {
            const int64_t j = j0 + lane;
            int b = -1;
            float c = 0.f;
            if (j < e) {
                const int g = ld_stream(indices + j);
                b = __ldg(gene_bucket + g);
                if (b >= 0) c = xform_value(ld_stream(counts + j), scale) * __ldg(gene_weight + g);
            }
            consume(b, c);
        }
        __syncwarp();
        float sq = 0.f;
        for (int c = lane * 4; c < d; c += 128) {
            const float4 a = *reinterpret_cast<float4 *>(acc + c);
            *reinterpret_cast<float4 *>(acc + c) = make_float4(0.f, 0.f, 0.f, 0.f);
            sq = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, sq))));
        }
        sq = warp_sum(sq);
        const int64_t orow = row_map ? (int64_t)__ldg(row_map + row) : it;
        float *out = h + orow * kp;
        if (lane < kp) out[lane] = h0;
        if (NK == 2 && 32 + lane < kp) out[32 + lane] = h1;
        if (lane == 0) ysq[orow] = sq;
        __syncwarp();
    }
}

constexpr int kPrefetch = 16;       // register-prefetched chunks of 32 entries (first 512 entries of a row)

// ------------------------------------------------------------------------------------
// fused form, v5 (production): the v3 structure (one row per warp, u16 gene -> slot table, cross-row register
// prefetch) with the three changes the round-2 profiles asked for:
//   * conflict-free AXPY: X_s^T rows are exactly 128 bytes (2 x 128 for Kp > 32) and lane j = lane & 7 reads the 16-byte
//     chunk (j XOR t) of its entry's row at step t, so the eight lanes of a quarter-warp always hit eight different bank
//     groups whatever rows they read (v3: padded rows, 57 % of all shared wavefronts were conflict replays);
//   * select-free reduction: lane j holds chunk (j XOR t) in register block t, so hv[t] += shfl_xor(hv[t + 4], 4);
//     hv[t] += shfl_xor(hv[t + 2], 2); hv[0] += shfl_xor(hv[1], 1) leaves chunk j in block 0 of every lane of a
//     quarter-warp, and two more xor steps fold the four quarters (36 shuffles, no FSEL; v3: 31 shuffles + 62 FSEL +
//     31 FADD);
//   * ||y_s||^2 without CAS loops: in log-CPM mode |c| <= log1p(1e4) |w|, so with W = max_b sum_{g in b} |w_g| every
//     bucket sum is below 9.22 W and the entries are added in fixed point (scale 2^k, native integer ATOMS.ADD,
//     order-independent -> deterministic); the sums are read back with atomicExch(acc[b], 0) by the entries themselves
//     (first reader of a bucket gets S_b, later ones 0), so only touched buckets are visited and nothing is re-zeroed.
//     The linear branches (raw / pearson: no bound on the values) keep float atomics.
// ------------------------------------------------------------------------------------
typedef unsigned long long sk_u64;
__device__ __forceinline__ sk_u64 sk_pack(float lo, float hi)
{
    sk_u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void sk_unpack(sk_u64 v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ sk_u64 sk_fma2(sk_u64 a, sk_u64 b, sk_u64 c)
{
    sk_u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ sk_u64 sk_add2(sk_u64 a, sk_u64 b)
{
    sk_u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ sk_u64 sk_shfl_xor(sk_u64 v, int m)
{
    float lo, hi;
    sk_unpack(v, lo, hi);
    lo = __shfl_xor_sync(kFull, lo, m);
    hi = __shfl_xor_sync(kFull, hi, m);
    return sk_pack(lo, hi);
}
__device__ __forceinline__ void sk_lds128(unsigned addr, sk_u64 &a, sk_u64 &b)
{
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
// log1p(v * scale) with MUFU.LG2 for 1 + x >= 1.5 (error ~2 ulp of the result, far inside the 1e-5 sketch tolerance);
// scale <= 0: the linear branches
__device__ __forceinline__ float sk_xform(float v, float scale)
{
    if (!(scale > 0.f)) return v;
    const float x = v * scale;
    if (x >= 0.5f) {
        float l2;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(1.f + x));
        return l2 * 0.693147180559945f;
    }
    return log1pf(x);
}

constexpr int kV5List = 256;        // compacted selected entries per flush (8-byte records)

// Multi-GPU output scatter of the fused kernel (world <= 1: unused).  Each rank sketches a slice of INPUT rows (the rows
// it uploaded) while the solve wants H in TILE order, sharded by spatial tile: the output row index delivered by
// row_map is a global tile position p; the row is written straight into the H / ||y_s||^2 buffers of the rank that owns
// p (peer memory over NVLink) -- the sketch and its all-to-all in one kernel.
constexpr int kSketchMaxRanks = 16;
struct SketchScatter {
    int world;
    int32_t bounds[kSketchMaxRanks + 1];      // rank q owns positions [bounds[q], bounds[q + 1])
    float *h[kSketchMaxRanks];                // q's H buffer (own rows x Kp), as mapped in THIS process
    float *ysq[kSketchMaxRanks];
};

// TAB = true: u16 gene -> slot table (2 bytes per gene); TAB = false: one membership bit per gene + a rank prefix per
// 32-gene word (slot = prefix + popc of the bits below), for wide gene axes / wide rows where the table does not fit
// next to X_s^T -- the list then carries the gene and the slot is computed for the selected entries only.
template <typename IndPtr, int NK, bool FIXED, bool TAB>
__global__ void __launch_bounds__(NK == 1 ? 512 : 384, 1)               // wide rows: 64 accumulator registers per lane -> 168 registers, no spills
sketch_contract_v5_kernel(const IndPtr *__restrict__ indptr, const int32_t *__restrict__ indices,
                          const float *__restrict__ counts, int64_t n_spots, int n_genes, int n_selected,
                          const int32_t *__restrict__ gene_bucket, const float *__restrict__ gene_weight,
                          int d, const float *__restrict__ x_sketch_t, int kp,
                          const int32_t *__restrict__ row_map, const int32_t *__restrict__ row_ids,
                          float *__restrict__ h, float *__restrict__ ysq, int linear,
                          const __grid_constant__ SketchScatter sc)
{
    constexpr int XS = NK * 32;                                                         // floats per staged X_s^T row
    constexpr int PF = NK == 1 ? 16 : 8;      // register-prefetched chunks of 32 entries (wide rows hold 64 accumulators)
    extern __shared__ unsigned char smem_raw[];
    __shared__ int scan_warp[32];
    __shared__ float s_wmax;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warps_per_cta = blockDim.x >> 5;
    const unsigned raw_addr = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned base_addr = (raw_addr + 127u) & ~127u;                               // XOR chunk addressing: 128-byte rows
    unsigned char *sm = smem_raw + (base_addr - raw_addr);
    const int per_warp_bytes = d * 4 + kV5List * 8;
    float *xs = reinterpret_cast<float *>(sm);                                          // d x XS
    int2 *slot_bw = reinterpret_cast<int2 *>(sm + (size_t)d * XS * 4);                  // n_selected (bucket, weight bits)
    unsigned char *warp_area = reinterpret_cast<unsigned char *>(slot_bw + ((n_selected + 1) & ~1));
    int *acc = reinterpret_cast<int *>(warp_area + (size_t)warp * per_warp_bytes);      // d words
    float2 *list = reinterpret_cast<float2 *>(reinterpret_cast<unsigned char *>(acc) + d * 4);   // (count, slot)
    unsigned char *tab_area = warp_area + (size_t)warps_per_cta * per_warp_bytes;
    unsigned short *gslot = reinterpret_cast<unsigned short *>(tab_area);                           // TAB: n_genes + 1
    const int nwords = (n_genes + 31) >> 5;
    unsigned *bitmap = reinterpret_cast<unsigned *>(tab_area);                                       // !TAB: nwords + 1
    unsigned *prefix = bitmap + nwords + 1;                                                          // !TAB: nwords + 1
    const unsigned list_addr = (unsigned)__cvta_generic_to_shared(list);

    for (int i = threadIdx.x; i < d * XS; i += blockDim.x) {
        const int r = i / XS, c = i - r * XS;
        xs[i] = c < kp ? __ldg(x_sketch_t + (size_t)r * kp + c) : 0.f;
    }
    for (int c = lane; c < d; c += 32) acc[c] = 0;
    if (threadIdx.x == 0) s_wmax = 0.f;
    if (TAB) {   // gene -> slot (rank among selected genes): block-wide exclusive scan over contiguous gene chunks
        const int per = (n_genes + (int)blockDim.x - 1) / (int)blockDim.x;
        const int g0 = min((int)threadIdx.x * per, n_genes), g1 = min(g0 + per, n_genes);
        int mine = 0;
        for (int g = g0; g < g1; ++g) mine += __ldg(gene_bucket + g) >= 0;
        int inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) scan_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = lane < warps_per_cta ? scan_warp[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(kFull, w, o);
                if (lane >= o) w += t;
            }
            scan_warp[lane] = w;
        }
        __syncthreads();
        int slot = (warp ? scan_warp[warp - 1] : 0) + inc - mine;
        for (int g = g0; g < g1; ++g) {
            const int b = __ldg(gene_bucket + g);
            unsigned short code = 0xFFFF;
            if (b >= 0 && slot < n_selected && slot < 0xFFFF) {
                slot_bw[slot] = make_int2(b, __float_as_int(__ldg(gene_weight + g)));
                code = (unsigned short)slot;
                ++slot;
            }
            gslot[g] = code;
        }
        if (threadIdx.x == 0) gslot[n_genes] = 0xFFFF;                                  // the pad gene of masked lanes
    } else {     // membership bitmap + rank prefix per word (word nwords stays empty: the pad gene of masked lanes)
        unsigned lt;
        asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt));
        for (int w = warp; w <= nwords; w += warps_per_cta) {
            const int g = 32 * w + lane;
            const bool sel = g < n_genes && __ldg(gene_bucket + g) >= 0;
            const unsigned bits = __ballot_sync(kFull, sel);
            if (lane == 0) { bitmap[w] = bits; prefix[w] = __popc(bits); }
        }
        __syncthreads();
        const int n = nwords + 1;
        const int per = (n + (int)blockDim.x - 1) / (int)blockDim.x;
        const int w0 = min((int)threadIdx.x * per, n), w1 = min(w0 + per, n);
        int tot = 0;
        for (int w = w0; w < w1; ++w) tot += (int)prefix[w];
        int inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) scan_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = lane < warps_per_cta ? scan_warp[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(kFull, w, o);
                if (lane >= o) w += t;
            }
            scan_warp[lane] = w;
        }
        __syncthreads();
        int run = (warp ? scan_warp[warp - 1] : 0) + inc - tot;
        for (int w = w0; w < w1; ++w) {
            const int c = (int)prefix[w];
            prefix[w] = (unsigned)run;
            run += c;
        }
        __syncthreads();
        for (int w = warp; w < nwords; w += warps_per_cta) {
            const int g = 32 * w + lane;
            const unsigned bits = bitmap[w];
            const int slot = (int)prefix[w] + __popc(bits & lt);
            const bool keep = ((bits >> lane) & 1u) && slot < n_selected;   // genes ranked past n_selected are dropped
            if (keep) slot_bw[slot] = make_int2(__ldg(gene_bucket + g), __float_as_int(__ldg(gene_weight + g)));
            const unsigned kept = __ballot_sync(kFull, keep);
            if (lane == 0) bitmap[w] = kept;
        }
    }
    __syncthreads();
    if (FIXED) {   // W = max over buckets of sum |w|: warp 0's accumulator as float scratch, then re-zeroed
        float *scratch = reinterpret_cast<float *>(warp_area);
        for (int sidx = threadIdx.x; sidx < n_selected; sidx += blockDim.x)
            atomicAdd(scratch + slot_bw[sidx].x, fabsf(__int_as_float(slot_bw[sidx].y)));
        __syncthreads();
        if (warp == 0) {
            float m = 0.f;
            for (int c = lane; c < d; c += 32) { m = fmaxf(m, scratch[c]); scratch[c] = 0.f; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
            if (lane == 0) s_wmax = m;
        }
        __syncthreads();
    }
    // fixed point: |bucket sum| <= 9.2104 W < 2^(ex+1)  =>  scale 2^(29-ex) keeps every sum below 2^30
    float q_scale = 0.f, q_inv = 0.f;
    if (FIXED) {
        const int bexp = (int)((__float_as_uint(9.2104f * s_wmax) >> 23) & 255u);
        if (bexp >= 30 && bexp <= 254) {
            q_scale = __uint_as_float((unsigned)(283 - bexp) << 23);
            q_inv = __uint_as_float((unsigned)(bexp - 29) << 23);
        }
    }

    unsigned lt_mask;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
    const int64_t stride = (int64_t)gridDim.x * warps_per_cta;
    const int j8 = lane & 7;
    const unsigned xs_lane = base_addr + 16u * j8;
    const int pad_gene = TAB ? n_genes : 32 * nwords;

    sk_u64 hv[NK * 16];                                                                 // NK x 8 chunk blocks (float pairs)
    // lane-parallel AXPY over list[0, n) with the row's scale
    auto flush = [&](int n, float scale) {
        // software-pipelined: the record and (bucket, weight) of the next iteration are fetched before this
        // iteration's eight row chunks, so the dependent LDS chain overlaps the FFMA2 block
        float2 rec = make_float2(0.f, 0.f);
        int2 bw = make_int2(0, 0);
        auto slot_of = [&](float code) -> int {
            const int x = __float_as_int(code);
            if (TAB) return x;
            return (int)prefix[x >> 5] + __popc(bitmap[x >> 5] & ((1u << (x & 31)) - 1u));
        };
        if (lane < n) {
            rec = list[lane];
            bw = slot_bw[slot_of(rec.y)];
        }
#pragma unroll 1
        for (int t0 = 0; t0 < n; t0 += 32) {
            const bool on = t0 + lane < n;
            const float v = rec.x;
            const int2 cur = bw;
            const int tn = t0 + 32 + lane;
            if (tn < n) {
                rec = list[tn];
                bw = slot_bw[slot_of(rec.y)];
            }
            if (on) {
                const float c = sk_xform(v, scale) * __int_as_float(cur.y);
                reinterpret_cast<int *>(list)[2 * (t0 + lane) + 1] = cur.x;             // the read-back pass wants the bucket
                if (FIXED) atomicAdd(acc + cur.x, __float2int_rn(c * q_scale));
                else atomicAdd(reinterpret_cast<float *>(acc) + cur.x, c);
                const unsigned a0 = xs_lane + (unsigned)cur.x * (XS * 4);
                const sk_u64 cc = sk_pack(c, c);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    sk_u64 x0, x1;
                    sk_lds128(a0 ^ (16u * q), x0, x1);
                    hv[2 * q] = sk_fma2(cc, x0, hv[2 * q]);
                    hv[2 * q + 1] = sk_fma2(cc, x1, hv[2 * q + 1]);
                    if (NK == 2) {
                        sk_lds128((a0 ^ (16u * q)) + 128u, x0, x1);
                        hv[16 + 2 * q] = sk_fma2(cc, x0, hv[16 + 2 * q]);
                        hv[16 + 2 * q + 1] = sk_fma2(cc, x1, hv[16 + 2 * q + 1]);
                    }
                }
            }
        }
    };
    // one chunk of 32 entries: look the slot up, add to the library size, append selected entries to the list
    auto take = [&](int g, float v, int &cnt, float &lib) {
        unsigned sl;
        bool sel;
        if (TAB) {
            sl = gslot[g];
            sel = sl != 0xFFFFu;
        } else {
            const unsigned word = bitmap[g >> 5];
            sl = (unsigned)g;
            sel = (__funnelshift_r(word, word, g) & 1u) != 0u;
        }
        const unsigned m = __ballot_sync(kFull, sel);
        const int pos = cnt + __popc(m & lt_mask);
        if (sel) {
            lib += v;
            if (pos < kV5List)
                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(list_addr + 8u * pos), "f"(v), "r"(sl) : "memory");
        }
        cnt += __popc(m);
    };
    auto row_of = [&](int64_t it) { return row_ids ? (int64_t)__ldg(row_ids + it) : it; };

    // 32-bit offsets inside a row (rows longer than 2^31 entries are rejected on the host)
    int64_t it = (int64_t)blockIdx.x * warps_per_cta + warp;
    int pg[PF];
    float pv[PF];
    int64_t s = 0, row = 0;
    int len = 0;
    if (it < n_spots) {
        row = row_of(it);
        s = load_ptr(indptr, row);
        len = (int)(load_ptr(indptr, row + 1) - s);
    }
    {
        const int32_t *ip = indices + s;
        const float *vp = counts + s;
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int j = 32 * u + lane;
            pg[u] = j < len ? ld_stream(ip + j) : pad_gene;
            pv[u] = j < len ? ld_stream(vp + j) : 0.f;
        }
    }
    while (it < n_spots) {
        const int64_t it_next = it + stride;
        // output row: loaded now so that the final store does not wait on a dependent global load
        const int64_t orow = row_map ? (int64_t)__ldg(row_map + row) : it;
        int64_t s2 = 0, row2 = 0;
        int len2 = 0;
        if (it_next < n_spots) {
            row2 = row_of(it_next);
            s2 = load_ptr(indptr, row2);
            len2 = (int)(load_ptr(indptr, row2 + 1) - s2);
        }
#pragma unroll
        for (int i = 0; i < NK * 16; ++i) asm volatile("mov.b64 %0, 0;" : "=l"(hv[i]));
        int cnt = 0;
        float lib = 0.f;
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            if (32 * u >= len) break;                                     // warp-uniform
            take(pg[u], pv[u], cnt, lib);
        }
        {
            const int32_t *ip = indices + s;
            const float *vp = counts + s;
            for (int j0 = 32 * PF; j0 < len; j0 += 128) {          // long rows: the rest, 4 chunks at a time
                int g[4];
                float v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + 32 * u + lane;
                    g[u] = j < len ? ld_stream(ip + j) : pad_gene;
                    v[u] = j < len ? ld_stream(vp + j) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (j0 + 32 * u >= len) break;
                    take(g[u], v[u], cnt, lib);
                }
            }
        }
        lib = warp_sum(lib);
        if (lib == 0.f) lib = 1.f;
        const float scale = linear ? -1.f : 1e4f / lib;
        const bool overflow = cnt > kV5List;
        // next row's first 512 entries start streaming now and land while this row's AXPY runs
        {
            const int32_t *ip = indices + s2;
            const float *vp = counts + s2;
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int j = 32 * u + lane;
                pg[u] = j < len2 ? ld_stream(ip + j) : pad_gene;
                pv[u] = j < len2 ? ld_stream(vp + j) : 0.f;
            }
        }
        __syncwarp();
        float sq = 0.f;
        if (!overflow) {
            flush(cnt, scale);
            __syncwarp();
            // bucket sums back: the first entry of a bucket to get there takes S_b and leaves 0
#pragma unroll 1
            for (int t = lane; t < cnt; t += 32) {
                const int b = __float_as_int(list[t].y);
                float sb;
                if (FIXED) sb = (float)atomicExch(acc + b, 0) * q_inv;
                else sb = atomicExch(reinterpret_cast<float *>(acc) + b, 0.f);
                sq = fmaf(sb, sb, sq);
            }
        } else {                                   // rare: more selected entries than the list holds -> re-stream
            int c2 = 0;
            float dummy = 0.f;
            const int32_t *ip = indices + s;
            const float *vp = counts + s;
            for (int j0 = 0; j0 < len; j0 += 32) {
                const int j = j0 + lane;
                const int g = j < len ? ld_stream(ip + j) : pad_gene;
                const float v = j < len ? ld_stream(vp + j) : 0.f;
                if (c2 + 32 > kV5List) {
                    __syncwarp();
                    flush(c2, scale);
                    __syncwarp();
                    c2 = 0;
                }
                take(g, v, c2, dummy);
            }
            __syncwarp();
            flush(c2, scale);
            __syncwarp();
            for (int c = lane; c < d; c += 32) {
                const float a = FIXED ? (float)acc[c] * q_inv : __int_as_float(acc[c]);
                acc[c] = 0;
                sq = fmaf(a, a, sq);
            }
        }
        __syncwarp();
        // reduction: eight lanes of a quarter (chunk j XOR t in block t), then the four quarters
#pragma unroll
        for (int hf = 0; hf < NK; ++hf) {
            sk_u64 *v = hv + 16 * hf;
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = sk_add2(v[i], sk_shfl_xor(v[i + 8], 4));
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = sk_add2(v[i], sk_shfl_xor(v[i + 4], 2));
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                v[i] = sk_add2(v[i], sk_shfl_xor(v[i + 2], 1));
                v[i] = sk_add2(v[i], sk_shfl_xor(v[i], 8));
                v[i] = sk_add2(v[i], sk_shfl_xor(v[i], 16));
            }
        }
        sq = warp_sum(sq);
        float *out = h + orow * kp;
        float *out_sq = ysq + orow;
        if (sc.world > 1) {                                               // orow is a global tile position: find its owner
            int q = 0;
            for (int r = 1; r < sc.world; ++r) q += orow >= sc.bounds[r];
            const int64_t local = orow - sc.bounds[q];
            out = sc.h[q] + local * kp;
            out_sq = sc.ysq[q] + local;
        }
        if (lane < 8) {                                                   // lane j: chunk j (and 8 + j) of H[orow]
            if (4 * lane < kp) {
                float4 o;
                sk_unpack(hv[0], o.x, o.y);
                sk_unpack(hv[1], o.z, o.w);
                *reinterpret_cast<float4 *>(out + 4 * lane) = o;
            }
            if (NK == 2 && 32 + 4 * lane < kp) {
                float4 o;
                sk_unpack(hv[16], o.x, o.y);
                sk_unpack(hv[17], o.z, o.w);
                *reinterpret_cast<float4 *>(out + 32 + 4 * lane) = o;
            }
        }
        if (lane == 0) *out_sq = sq;
        __syncwarp();
        it = it_next; s = s2; len = len2; row = row2;
    }
}

// ------------------------------------------------------------------------------------
// unfused contraction (API / parity form): H = Y_s X_s^T, ysq = rowwise ||y_s||^2
// one warp per spot, X_s staged in shared memory
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
contract_kernel(const float *__restrict__ y_sketch, const float *__restrict__ x_sketch,
                int64_t n_spots, int d, int k0, int nk, int n_types, int kp, float *__restrict__ h,
                float *__restrict__ ysq)
{
    // one launch handles the type columns [k0, k0 + nk) (nk x d floats of X_s fit in shared memory); the last chunk also
    // zeroes the padding columns [n_types, kp)
    extern __shared__ __align__(16) float xs[];     // nk x d
    for (int i = threadIdx.x; i < nk * d; i += blockDim.x) xs[i] = __ldg(x_sketch + (size_t)k0 * d + i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warps_per_cta = blockDim.x >> 5;
    const int k_end = k0 + nk == n_types ? kp : k0 + nk;
    for (int64_t row = (int64_t)blockIdx.x * warps_per_cta + warp; row < n_spots;
         row += (int64_t)gridDim.x * warps_per_cta) {
        const float *y = y_sketch + row * (int64_t)d;
        if (k0 == 0) {
            float sq = 0.f;
            for (int c = lane; c < d; c += 32) {
                const float v = y[c];
                sq = fmaf(v, v, sq);
            }
            sq = warp_sum(sq);
            if (lane == 0) ysq[row] = sq;
        }
        for (int k = k0; k < k_end; ++k) {
            float a = 0.f;
            if (k < n_types)
                for (int c = lane; c < d; c += 32) a = fmaf(y[c], xs[(k - k0) * d + c], a);
            a = warp_sum(a);
            if (lane == 0) h[row * kp + k] = a;
        }
    }
}

// ------------------------------------------------------------------------------------
// (f1) per-gene moments of log1p(CP10k) over ALL genes: utils/genes.py:52-83
// warp per row; fp64 atomics into the G-long accumulators (L2-resident, 2*8*G bytes)
// ------------------------------------------------------------------------------------
template <typename IndPtr>
__global__ void __launch_bounds__(256)
gene_moments_kernel(const IndPtr *__restrict__ indptr, const int32_t *__restrict__ indices,
                    const float *__restrict__ counts, int64_t n_spots, double *__restrict__ sums,
                    double *__restrict__ sumsq)
{
    // float64 throughout: the ranking built from these moments must reproduce the reference's gene
    // selection exactly, so only the summation order may differ from numpy (utils/genes.py:57-75)
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp_global; row < n_spots; row += n_warps) {
        const int64_t s = load_ptr(indptr, row), e = load_ptr(indptr, row + 1);
        double lib = 0.0;
        for (int64_t j = s + lane; j < e; j += 32) lib += (double)__ldg(counts + j);
        lib = warp_sum(lib);
        const double scale = 10000.0 / fmax(lib, 1.0);
        for (int64_t j = s + lane; j < e; j += 32) {
            const double z = log1p(scale * (double)__ldg(counts + j));
            const int g = ld_stream(indices + j);
            atomicAdd(sums + g, z);
            atomicAdd(sumsq + g, z * z);
        }
    }
}

static int pick_grid(int64_t n_rows, int warps_per_cta, int ctas_per_sm)
{
    int64_t want = ceil_div(n_rows, warps_per_cta);
    int64_t cap = (int64_t)kNumSM * ctas_per_sm;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace fdb

using namespace fdb;

template <typename IndPtr, bool LOGCPM>
static int launch_rows(const void *indptr, const int32_t *indices, const float *counts, int64_t n_spots,
                       const int32_t *gene_bucket, const float *gene_weight, int d, float *y_sketch,
                       cudaStream_t st)
{
    int warps = 16;
    while (warps > 1 && (size_t)warps * d * 4 > 96 * 1024) warps >>= 1;
    const size_t smem = (size_t)warps * d * 4;
    auto kern = sketch_rows_kernel<IndPtr, LOGCPM>;
    FDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<pick_grid(n_spots, warps, 2), warps * 32, smem, st>>>((const IndPtr *)indptr, indices, counts, n_spots,
                                                                 gene_bucket, gene_weight, d, y_sketch);
    FDB_LAUNCH_CHECK("sketch_rows_kernel");
    return FDB_OK;
}

static int sketch_rows(const void *indptr, int is64, const int32_t *indices, const float *counts, int64_t n_spots,
                       int32_t n_genes, const int32_t *gene_bucket, const float *gene_weight, int32_t d,
                       float *y_sketch, void *stream, bool logcpm)
{
    FDB_REQUIRE(n_spots >= 0 && n_genes >= 0, "negative shape");
    FDB_REQUIRE(d > 0 && d % 4 == 0 && d <= 8192, "sketch_dim must be a positive multiple of 4 (<= 8192), got %d", d);
    if (n_spots == 0) return FDB_OK;
    FDB_REQUIRE(indptr && gene_bucket && gene_weight && y_sketch, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (is64)
        return logcpm ? launch_rows<int64_t, true>(indptr, indices, counts, n_spots, gene_bucket, gene_weight, d, y_sketch, st)
                      : launch_rows<int64_t, false>(indptr, indices, counts, n_spots, gene_bucket, gene_weight, d, y_sketch, st);
    return logcpm ? launch_rows<int32_t, true>(indptr, indices, counts, n_spots, gene_bucket, gene_weight, d, y_sketch, st)
                  : launch_rows<int32_t, false>(indptr, indices, counts, n_spots, gene_bucket, gene_weight, d, y_sketch, st);
}

extern "C" __attribute__((visibility("default"))) int fdb_sketch_logcpm_csr(
    const void *indptr, int indptr_is_int64, const int32_t *indices, const float *counts, int64_t n_spots,
    int32_t n_genes, const int32_t *gene_bucket, const float *gene_weight, int32_t d, float *y_sketch, void *stream)
{
    return sketch_rows(indptr, indptr_is_int64, indices, counts, n_spots, n_genes, gene_bucket, gene_weight, d,
                       y_sketch, stream, true);
}

extern "C" __attribute__((visibility("default"))) int fdb_sketch_project_csr(
    const void *indptr, int indptr_is_int64, const int32_t *indices, const float *values, int64_t n_spots,
    int32_t n_genes, const int32_t *gene_bucket, const float *gene_weight, int32_t d, float *y_sketch, void *stream)
{
    return sketch_rows(indptr, indptr_is_int64, indices, values, n_spots, n_genes, gene_bucket, gene_weight, d,
                       y_sketch, stream, false);
}

extern "C" __attribute__((visibility("default"))) int fdb_contract(const float *y_sketch, const float *x_sketch, int64_t n_spots, int32_t d,
                            int32_t n_types, float *h, float *ysq, void *stream)
{
    FDB_REQUIRE(n_spots >= 0 && d > 0 && n_types > 0 && n_types <= FDB_MAX_TYPES_WIDE, "bad shape");
    if (n_spots == 0) return FDB_OK;
    FDB_REQUIRE((size_t)d * 4 <= 200 * 1024, "sketch_dim %d: one row of X_s does not fit in shared memory", d);
    const int kp = fdb_padded_types(n_types);
    // type columns in chunks whose X_s rows fit in shared memory (one chunk for K <= 100 at d = 512)
    const int chunk = (int)std::min<size_t>((size_t)n_types, std::max<size_t>(1, (size_t)200 * 1024 / ((size_t)d * 4)));
    for (int k0 = 0; k0 < n_types; k0 += chunk) {
        const int nk = std::min(chunk, n_types - k0);
        const size_t smem = (size_t)nk * d * 4;
        FDB_CUDA(cudaFuncSetAttribute(contract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = pick_grid(n_spots, 8, smem > 100 * 1024 ? 1 : 2);
        contract_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(y_sketch, x_sketch, n_spots, d, k0, nk, n_types, kp, h, ysq);
        FDB_LAUNCH_CHECK("contract_kernel");
    }
    return FDB_OK;
}

template <typename IndPtr, int NK>
static int launch_fused(const void *indptr, const int32_t *indices, const float *counts,
                        int64_t n_spots, int n_genes, int n_selected, const int32_t *gene_bucket,
                        const float *gene_weight, int d, const float *x_sketch_t, int kp, const int32_t *row_map,
                        const int32_t *row_ids, float *h, float *ysq, int linear, cudaStream_t st,
                        const SketchScatter *scatter = nullptr)
{
    SketchScatter sc;
    if (scatter) sc = *scatter;
    else sc.world = 1;
    // production: v5 (v3 structure + conflict-free XOR-phased AXPY, select-free reduction, integer atomics); the u16
    // gene -> slot table when it leaves room for at least 12 warps, else the membership bitmap + rank prefix
    if (n_selected >= 0 && getenv("FDB_SKETCH_V1") == nullptr) {
        const int force_tab = getenv("FDB_SKETCH_TAB") ? atoi(getenv("FDB_SKETCH_TAB")) : -1;
        const size_t common = (size_t)d * NK * 128 + (size_t)((n_selected + 1) & ~1) * 8 + 16 + 128;
        const size_t per_warp = (size_t)d * 4 + kV5List * 8;
        const size_t limit = 227 * 1024 - 256;
        auto warps_for = [&](size_t table_bytes) {
            for (int w : {16, 14, 12, 10, 8, 6, 4})
                if (w * 32 <= (NK == 1 ? 512 : 384) && common + table_bytes + w * per_warp <= limit) return w;
            return 0;
        };
        const size_t tab_bytes = (size_t)(n_genes + 1) * 2, bit_bytes = (size_t)(((n_genes + 31) >> 5) + 1) * 8;
        const int w_tab = n_selected < 0xFFFF ? warps_for(tab_bytes) : 0, w_bit = warps_for(bit_bytes);
        const bool use_tab = force_tab >= 0 ? (force_tab != 0 && w_tab > 0) : (w_tab >= 12 || w_tab >= w_bit);
        const int warps = use_tab ? w_tab : w_bit;
        if (warps) {
            const size_t smem = common + (use_tab ? tab_bytes : bit_bytes) + warps * per_warp;
            const int grid = pick_grid(n_spots, warps, 1);
            auto go = [&](auto kern) -> int {
                FDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                kern<<<grid, warps * 32, smem, st>>>((const IndPtr *)indptr, indices, counts, n_spots, n_genes, n_selected,
                                                     gene_bucket, gene_weight, d, x_sketch_t, kp, row_map, row_ids, h, ysq,
                                                     linear, sc);
                FDB_LAUNCH_CHECK("sketch_contract_v5_kernel");
                return FDB_OK;
            };
            if (use_tab)
                return linear ? go(sketch_contract_v5_kernel<IndPtr, NK, false, true>)
                              : go(sketch_contract_v5_kernel<IndPtr, NK, true, true>);
            return linear ? go(sketch_contract_v5_kernel<IndPtr, NK, false, false>)
                          : go(sketch_contract_v5_kernel<IndPtr, NK, true, false>);
        }
    }
    if (scatter) {
        set_error("the scattering sketch needs the shared-memory kernel (n_selected known, tables that fit)");
        return FDB_ERR_UNSUPPORTED;
    }
    // fallback: v1 (tables in global memory) for very wide gene axes / sketches
    const size_t xs_bytes = (size_t)d * NK * 32 * 4;
    int warps = 16;
    while (warps > 1 && xs_bytes + (size_t)warps * d * 4 > 110 * 1024) warps >>= 1;
    size_t smem = xs_bytes + (size_t)warps * d * 4;
    int per_sm = 2;
    if (smem > 110 * 1024) {                       // large X_s^T: one CTA per SM, as many warps as fit
        per_sm = 1;
        warps = 16;
        while (warps > 1 && xs_bytes + (size_t)warps * d * 4 > 220 * 1024) warps >>= 1;
        smem = xs_bytes + (size_t)warps * d * 4;
    }
    if (smem > 227 * 1024) {
        set_error("sketch_dim %d x %d types does not fit in shared memory", d, kp);
        return FDB_ERR_UNSUPPORTED;
    }
    auto kern = sketch_contract_kernel<IndPtr, NK>;
    FDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = pick_grid(n_spots, warps, per_sm);
    kern<<<grid, warps * 32, smem, st>>>((const IndPtr *)indptr, indices, counts, n_spots, gene_bucket,
                                         gene_weight, d, x_sketch_t, kp, row_map, row_ids, h, ysq, linear);
    FDB_LAUNCH_CHECK("sketch_contract_kernel");
    return FDB_OK;
}

static int sketch_contract_impl(const void *indptr, int indptr_is_int64, const int32_t *indices, const float *counts,
                                int64_t n_spots, int32_t n_genes, const int32_t *gene_bucket, const float *gene_weight,
                                int32_t d, const float *x_sketch_t, int32_t n_types, const int32_t *row_map,
                                const int32_t *row_ids, int32_t n_selected, float *h, float *ysq, int linear, void *stream,
                                const SketchScatter *scatter = nullptr)
{
    FDB_REQUIRE(n_spots >= 0 && n_genes >= 0, "negative shape");
    FDB_REQUIRE(d > 0 && d % 4 == 0, "sketch_dim must be a positive multiple of 4, got %d", d);
    FDB_REQUIRE(n_types > 0 && n_types <= FDB_MAX_TYPES, "n_types must be in [1, %d], got %d", FDB_MAX_TYPES, n_types);
    if (n_spots == 0) return FDB_OK;
    FDB_REQUIRE(indptr && gene_bucket && gene_weight && x_sketch_t && ((h && ysq) || scatter), "null pointer");
    const int kp = fdb_padded_types(n_types);
    cudaStream_t st = (cudaStream_t)stream;
    if (kp <= 32)
        return indptr_is_int64
                   ? launch_fused<int64_t, 1>(indptr, indices, counts, n_spots, n_genes, n_selected, gene_bucket, gene_weight, d, x_sketch_t, kp, row_map, row_ids, h, ysq, linear, st, scatter)
                   : launch_fused<int32_t, 1>(indptr, indices, counts, n_spots, n_genes, n_selected, gene_bucket, gene_weight, d, x_sketch_t, kp, row_map, row_ids, h, ysq, linear, st, scatter);
    return indptr_is_int64
               ? launch_fused<int64_t, 2>(indptr, indices, counts, n_spots, n_genes, n_selected, gene_bucket, gene_weight, d, x_sketch_t, kp, row_map, row_ids, h, ysq, linear, st, scatter)
               : launch_fused<int32_t, 2>(indptr, indices, counts, n_spots, n_genes, n_selected, gene_bucket, gene_weight, d, x_sketch_t, kp, row_map, row_ids, h, ysq, linear, st, scatter);
}

extern "C" __attribute__((visibility("default"))) int fdb_sketch_contract_csr(const void *indptr, int indptr_is_int64, const int32_t *indices,
                                       const float *counts, int64_t n_spots, int32_t n_genes,
                                       const int32_t *gene_bucket, const float *gene_weight, int32_t d,
                                       const float *x_sketch_t, int32_t n_types, const int32_t *row_map,
                                       const int32_t *row_ids, int32_t n_selected, float *h, float *ysq, void *stream)
{
    return sketch_contract_impl(indptr, indptr_is_int64, indices, counts, n_spots, n_genes, gene_bucket, gene_weight, d,
                                x_sketch_t, n_types, row_map, row_ids, n_selected, h, ysq, 0, stream);
}

extern "C" __attribute__((visibility("default"))) int fdb_sketch_linear_contract_csr(const void *indptr, int indptr_is_int64, const int32_t *indices,
                                       const float *counts, int64_t n_spots, int32_t n_genes,
                                       const int32_t *gene_bucket, const float *gene_weight, int32_t d,
                                       const float *x_sketch_t, int32_t n_types, const int32_t *row_map,
                                       const int32_t *row_ids, int32_t n_selected, float *h, float *ysq, void *stream)
{
    return sketch_contract_impl(indptr, indptr_is_int64, indices, counts, n_spots, n_genes, gene_bucket, gene_weight, d,
                                x_sketch_t, n_types, row_map, row_ids, n_selected, h, ysq, 1, stream);
}

extern "C" __attribute__((visibility("default"))) int fdb_sketch_contract_scatter_csr(
    const void *indptr, int indptr_is_int64, const int32_t *indices, const float *counts, int64_t n_spots, int32_t n_genes,
    const int32_t *gene_bucket, const float *gene_weight, int32_t d, const float *x_sketch_t, int32_t n_types,
    const int32_t *row_map, int32_t n_selected, int32_t linear, int32_t world, const int32_t *host_bounds,
    void *const *host_h, void *const *host_ysq, void *stream)
{
    FDB_REQUIRE(world >= 1 && world <= kSketchMaxRanks && host_bounds && host_h && host_ysq && row_map, "bad scatter arguments");
    SketchScatter sc;
    sc.world = world;
    for (int q = 0; q <= kSketchMaxRanks; ++q) sc.bounds[q] = host_bounds[std::min(q, (int)world)];
    for (int q = 0; q < kSketchMaxRanks; ++q) {
        sc.h[q] = q < world ? (float *)host_h[q] : nullptr;
        sc.ysq[q] = q < world ? (float *)host_ysq[q] : nullptr;
    }
    if (world == 1)      // plain form: row_map indexes the single rank's buffers
        return sketch_contract_impl(indptr, indptr_is_int64, indices, counts, n_spots, n_genes, gene_bucket, gene_weight, d,
                                    x_sketch_t, n_types, row_map, nullptr, n_selected, sc.h[0], sc.ysq[0], linear, stream);
    return sketch_contract_impl(indptr, indptr_is_int64, indices, counts, n_spots, n_genes, gene_bucket, gene_weight, d,
                                x_sketch_t, n_types, row_map, nullptr, n_selected, nullptr, nullptr, linear, stream, &sc);
}

// per-gene sums of the raw counts (float64), for the per-gene scale of preprocess "pearson" (core/deconv.py:206-212)
__global__ void __launch_bounds__(256)
gene_sums_kernel(const int32_t *__restrict__ indices, const float *__restrict__ counts, int64_t nnz, double *__restrict__ sums)
{
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nnz; j += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(sums + ld_stream(indices + j), (double)ld_stream(counts + j));
}

extern "C" __attribute__((visibility("default"))) int fdb_gene_sums_csr(const int32_t *indices, const float *counts, int64_t nnz, int32_t n_genes,
                                 double *sums, void *stream)
{
    FDB_REQUIRE(nnz >= 0 && n_genes >= 0, "negative shape");
    if (nnz == 0) return FDB_OK;
    FDB_REQUIRE(indices && counts && sums, "null pointer");
    const int grid = (int)std::min<int64_t>(ceil_div(nnz, 256), (int64_t)kNumSM * 16);
    gene_sums_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(indices, counts, nnz, sums);
    FDB_LAUNCH_CHECK("gene_sums_kernel");
    return FDB_OK;
}

// (f3) per-type mean expression of a cells x genes reference (io/loader.py:119-136): float64 sums per (label, gene),
// warp per cell; the caller divides by the group sizes
template <typename IndPtr>
__global__ void __launch_bounds__(256)
group_sums_kernel(const IndPtr *__restrict__ indptr, const int32_t *__restrict__ indices, const float *__restrict__ values,
                  const int32_t *__restrict__ labels, int64_t n_rows, int n_genes, double *__restrict__ sums)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp_global; row < n_rows; row += n_warps) {
        const int64_t s = load_ptr(indptr, row), e = load_ptr(indptr, row + 1);
        const int label = labels[row];
        if (label < 0) continue;
        double *out = sums + (int64_t)label * n_genes;
        for (int64_t j = s + lane; j < e; j += 32) atomicAdd(out + ld_stream(indices + j), (double)ld_stream(values + j));
    }
}

extern "C" __attribute__((visibility("default"))) int fdb_group_sums_csr(const void *indptr, int indptr_is_int64, const int32_t *indices,
                                  const float *values, const int32_t *labels, int64_t n_rows, int32_t n_genes,
                                  int32_t n_groups, double *sums, void *stream)
{
    FDB_REQUIRE(n_rows >= 0 && n_genes >= 0 && n_groups >= 0, "negative shape");
    if (n_rows == 0) return FDB_OK;
    FDB_REQUIRE(indptr && labels && sums, "null pointer");
    const int grid = pick_grid(n_rows, 8, 8);
    if (indptr_is_int64)
        group_sums_kernel<int64_t><<<grid, 256, 0, (cudaStream_t)stream>>>((const int64_t *)indptr, indices, values, labels, n_rows, n_genes, sums);
    else
        group_sums_kernel<int32_t><<<grid, 256, 0, (cudaStream_t)stream>>>((const int32_t *)indptr, indices, values, labels, n_rows, n_genes, sums);
    FDB_LAUNCH_CHECK("group_sums_kernel");
    return FDB_OK;
}

extern "C" __attribute__((visibility("default"))) int fdb_gene_moments_csr(const void *indptr, int indptr_is_int64, const int32_t *indices,
                                    const float *counts, int64_t n_spots, int32_t n_genes, double *sums,
                                    double *sumsq, void *stream)
{
    FDB_REQUIRE(n_spots >= 0 && n_genes >= 0, "negative shape");
    if (n_spots == 0) return FDB_OK;
    const int grid = pick_grid(n_spots, 8, 8);
    if (indptr_is_int64)
        gene_moments_kernel<int64_t><<<grid, 256, 0, (cudaStream_t)stream>>>((const int64_t *)indptr, indices, counts, n_spots, sums, sumsq);
    else
        gene_moments_kernel<int32_t><<<grid, 256, 0, (cudaStream_t)stream>>>((const int32_t *)indptr, indices, counts, n_spots, sums, sumsq);
    FDB_LAUNCH_CHECK("gene_moments_kernel");
    return FDB_OK;
}
