// Production fused sketch kernel: log-CPM (or a linear per-gene scaling) + CountSketch + contraction against the
// sketched reference in ONE pass over the spot-by-gene CSR, never materialising Y_s.
//
// Reference semantics (upstream file:line):
//   core/deconv.py:321        Y[:, gene_idx]            -> a bitmap over the gene axis masks unselected genes
//   core/deconv.py:183-188    lib over selected genes (0 -> 1), y~ = log1p(1e4 * y / lib)
//   core/sketching.py:195     y_s = y~ Omega (one (bucket, weight) per selected gene)
//   core/solver.py:223, 348   H[i, :] = X_s y_s = sum_e c_e X_s^T[bucket_e, :],   ||y_s||^2
//
// Why it looks the way it does (ncu, round 1 and the first cuts of this kernel: the shared-memory/LSU pipe is the
// limiter -- 403 wavefronts per row, 57 % of them bank-conflict replays of random-row LDS.128 reads of X_s^T; float
// shared atomics are CAS spin loops; a fully unrolled stream of 7000 SASS instructions starves on the I-cache):
//   * a warp owns a batch of FOUR consecutive rows.  Each row is streamed as aligned 128-entry groups (512 bytes of
//     column indices + 512 bytes of counts) through a per-warp shared-memory ring that one lane fills with 1-D TMA
//     bulk copies (cp.async.bulk + one mbarrier per stage): no per-lane address arithmetic, no registers held by
//     loads in flight, no scoreboard coupling between prefetch depth and first use (a register ring of four 16-byte
//     loads per lane measured 26 % long-scoreboard stalls: SASS has six scoreboards), and a rolled loop of a few
//     hundred instructions.  The producer runs NST groups ahead, across rows and into the next batch.
//   * membership is ONE bit per gene (bitmap in shared memory: 32 sorted column indices touch ~45 consecutive words,
//     so the lookup is nearly conflict-free); selected entries are compacted (count, gene) into a per-warp list with
//     one ballot per 32 entries, predicated (no divergent regions).
//   * the AXPY runs with EIGHT lanes per row (the four rows of the batch side by side).  Lane j of a row reads the
//     16-byte chunk (j XOR t) of its entry's X_s^T row at step t: the eight lanes of a quarter-warp always hit
//     eight different 16-byte bank groups, whatever rows they read -> conflict-free by construction
//     (X_s^T rows are padded to 128 bytes, or 2 x 128 for Kp > 32).  The rows' list regions start 64-byte aligned
//     with alternating 64-byte phase, so the four groups' record reads/writes fall into two wavefronts.
//   * because lane j holds chunk (j XOR t) in register block t, the cross-lane reduction is three xor-shuffle
//     steps on static registers, no selects: hv[t] += shfl_xor(hv[t + 4], 4); hv[t] += shfl_xor(hv[t + 2], 2);
//     hv[0] += shfl_xor(hv[1], 1) -- 28 shuffles for FOUR rows, and lane j ends with the float4 chunk j of H[i].
//   * gene -> (bucket, weight) goes through a rank, only for the ~20 % of entries that are selected:
//     slot = prefix[word] + popc(bits below).
//   * ||y_s||^2 = sum_b S_b^2 with S_b the bucket sums: per row the transformed entries are added into a per-warp
//     512-word accumulator in block fixed point (scale 2^k chosen from the row's n * max|c|, so the sums cannot
//     overflow and carry > 29 bits: native integer ATOMS.ADD, order-independent -> deterministic) and read back with
//     atomicExch(acc[b], 0): the first reader of a bucket gets S_b and leaves 0 for the next row, later readers get 0,
//     so only touched buckets are visited and nothing is re-zeroed.
//   * rows whose selected entries do not fit the list (dense inputs) take a single-row path: re-stream (library size
//     first, then flush whenever the list fills), all 32 lanes on the one row.
// Bulk copies start at the row start rounded down to a multiple of four entries and end at the row end rounded up
// to one (neighbouring rows' entries are masked) -- at the very end of the arrays that is at most 12 bytes inside the
// last valid 16-byte block, which is why `indices` and `counts` must be 16-byte aligned for this kernel.
#pragma once
#include "fdb_common.cuh"

namespace fdb {

constexpr int kV4Rows = 4;                      // rows per warp batch (8 lanes each in the AXPY phase)
constexpr int kV4Group = 128;                   // entries per ring stage (512 bytes per array)

typedef unsigned long long v4u64;
__device__ __forceinline__ v4u64 v4_pack(float lo, float hi)
{
    v4u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void v4_unpack(v4u64 v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ v4u64 v4_fma2(v4u64 a, v4u64 b, v4u64 c)
{
    v4u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ v4u64 v4_add2(v4u64 a, v4u64 b)
{
    v4u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ v4u64 v4_shfl_xor(v4u64 v, int m)
{
    float lo, hi;
    v4_unpack(v, lo, hi);
    lo = __shfl_xor_sync(kFull, lo, m);
    hi = __shfl_xor_sync(kFull, hi, m);
    return v4_pack(lo, hi);
}
__device__ __forceinline__ void v4_lds128(unsigned addr, v4u64 &a, v4u64 &b)
{
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
__device__ __forceinline__ int4 v4_lds_i4(unsigned addr)
{
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 v4_lds_f4(unsigned addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned v4_lds_u32(unsigned addr)
{
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void v4_sts_rec(unsigned addr, float v, int g)
{
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "f"(v), "r"(g) : "memory");
}
// ---- mbarrier / 1-D TMA bulk copy (global -> this CTA's shared memory)
__device__ __forceinline__ void v4_mbar_init(unsigned bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void v4_mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void v4_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void v4_mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "V4_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra V4_DONE;\n"
        "bra V4_WAIT;\n"
        "V4_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// log-CPM value: log1p(v * scale) (scale > 0), or v itself for the linear branches (scale <= 0).  1 + x >= 1.5 goes
// through MUFU.LG2 (2 ulp-class error, far inside the 1e-5 sketch tolerance); small arguments keep the exact routine.
__device__ __forceinline__ float v4_xform(float v, float scale)
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

struct V4Layout {                    // byte offsets into the (128-byte aligned) dynamic shared memory block
    int xs, bitmap, prefix, slot_bw, per_warp, list, ring, bars, per_warp_bytes, total;
    int nwords, cap, nst;
};

// host + device: identical carve-up.  nh = 1 (Kp <= 32) or 2.  Per warp: accumulator (d words), list (cap x 8 bytes),
// ring (nst stages x (512 B indices + 512 B counts)), nst mbarriers.
__host__ __device__ inline V4Layout v4_layout(int d, int nh, int n_genes, int n_selected, int warps, int cap, int nst)
{
    V4Layout L;
    L.nwords = (n_genes + 31) >> 5;
    L.cap = cap;
    L.nst = nst;
    L.xs = 0;
    L.bitmap = L.xs + d * 128 * nh;
    L.prefix = L.bitmap + (int)round_up((L.nwords + 1) * 4, 16);
    L.slot_bw = L.prefix + (int)round_up((L.nwords + 1) * 4, 16);
    L.per_warp = (int)round_up(L.slot_bw + (int64_t)(n_selected > 0 ? n_selected : 1) * 8, 128);
    L.list = (int)round_up((int64_t)d * 4, 128);
    L.ring = L.list + (int)round_up((int64_t)cap * 8, 128);
    L.bars = L.ring + nst * 8 * kV4Group;
    L.per_warp_bytes = (int)round_up(L.bars + nst * 8, 128);
    L.total = L.per_warp + warps * L.per_warp_bytes;
    return L;
}

template <typename IndPtr, int NH>
__global__ void __launch_bounds__(512, 1)
sketch_contract_v4_kernel(const IndPtr *__restrict__ indptr, const int32_t *__restrict__ indices,
                          const float *__restrict__ counts, int64_t n_spots, int n_genes, int n_selected,
                          const int32_t *__restrict__ gene_bucket, const float *__restrict__ gene_weight,
                          int d, const float *__restrict__ x_sketch_t, int kp,
                          const int32_t *__restrict__ row_map, const int32_t *__restrict__ row_ids,
                          float *__restrict__ h, float *__restrict__ ysq, int linear, int cap, int nst)
{
    extern __shared__ unsigned char v4_smem_raw[];
    __shared__ int scan_warp[32];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warps_per_cta = blockDim.x >> 5;
    const V4Layout L = v4_layout(d, NH, n_genes, n_selected, warps_per_cta, cap, nst);
    // 128-byte aligned base (the XOR chunk addressing needs X_s^T rows on 128-byte boundaries)
    const unsigned raw = (unsigned)__cvta_generic_to_shared(v4_smem_raw);
    const unsigned base = (raw + 127u) & ~127u;
    unsigned char *sm = v4_smem_raw + (base - raw);
    float *xs = reinterpret_cast<float *>(sm + L.xs);
    unsigned *bitmap = reinterpret_cast<unsigned *>(sm + L.bitmap);
    unsigned *prefix = reinterpret_cast<unsigned *>(sm + L.prefix);
    int2 *slot_bw = reinterpret_cast<int2 *>(sm + L.slot_bw);
    unsigned char *mine = sm + L.per_warp + (size_t)warp * L.per_warp_bytes;
    int *acc = reinterpret_cast<int *>(mine);
    float2 *list = reinterpret_cast<float2 *>(mine + L.list);
    const unsigned xs_addr = base + L.xs;
    const unsigned bitmap_addr = base + L.bitmap;
    const unsigned mine_addr = base + L.per_warp + (unsigned)warp * L.per_warp_bytes;
    const unsigned list_addr = mine_addr + L.list;
    const unsigned ring_addr = mine_addr + L.ring;
    const unsigned bars_addr = mine_addr + L.bars;
    constexpr int XS = 32 * NH;                                        // floats per staged X_s^T row

    unsigned lt_mask;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));

    // ---------------- per-CTA tables: membership bitmap, rank prefix, slot -> (bucket, weight), X_s^T
    const int nwords = L.nwords;
    {
        for (int w = warp; w <= nwords; w += warps_per_cta) {
            const int g = 32 * w + lane;
            const bool s = g < n_genes && __ldg(gene_bucket + g) >= 0;
            const unsigned bits = __ballot_sync(kFull, s);
            if (lane == 0) { bitmap[w] = bits; prefix[w] = __popc(bits); }
        }
        for (int i = threadIdx.x; i < d * XS; i += blockDim.x) {
            const int r = i / XS, c = i - r * XS;
            xs[i] = c < kp ? __ldg(x_sketch_t + (size_t)r * kp + c) : 0.f;
        }
        for (int c = threadIdx.x; c < (L.per_warp_bytes * warps_per_cta) / 4; c += blockDim.x)
            reinterpret_cast<int *>(sm + L.per_warp)[c] = 0;           // accumulators, lists and rings start at zero
        __syncthreads();
        // exclusive scan of the per-word counts (contiguous slice per thread)
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
            const bool s = (bits >> lane) & 1u;
            const int slot = (int)prefix[w] + __popc(bits & lt_mask);
            // genes ranked past n_selected (caller passed too small a count) are dropped from the bitmap
            const bool keep = s && slot < n_selected;
            if (keep) slot_bw[slot] = make_int2(__ldg(gene_bucket + g), __float_as_int(__ldg(gene_weight + g)));
            const unsigned kept = __ballot_sync(kFull, keep);
            if (lane == 0) bitmap[w] = kept;
        }
        if (lane == 0) {
            for (int s = 0; s < nst; ++s) v4_mbar_init(bars_addr + 8 * s, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // zero-filled rings before the first bulk write
    }
    __syncthreads();

    const int j8 = lane & 7, grp = lane >> 3;
    const int64_t n_batches = (n_spots + kV4Rows - 1) / kV4Rows;
    const int64_t bstride = (int64_t)gridDim.x * warps_per_cta;

    // ---- batch metadata: lane r < 4 holds row r of the batch -------------------------------------------------
    struct Raw {
        int64_t s, orow;
        int len;
    };
    // the loads are issued one batch before their values are needed
    auto load_raw = [&](int64_t bt) {
        Raw w;
        w.s = 0; w.len = 0; w.orow = -1;
        const int64_t it = bt * kV4Rows + lane;
        if (lane < kV4Rows && bt < n_batches && it < n_spots) {
            const int64_t row = row_ids ? (int64_t)__ldg(row_ids + it) : it;
            w.s = load_ptr(indptr, row);
            w.len = (int)(load_ptr(indptr, row + 1) - w.s);
            w.orow = row_map ? (int64_t)__ldg(row_map + row) : it;
        }
        return w;
    };

    // ---- producer: the stream of 128-entry groups, rows of `cur` (0..3) then of `nxt` (4..7) -------------------
    Raw cur = load_raw((int64_t)blockIdx.x * warps_per_cta + warp);
    Raw nxt = load_raw((int64_t)blockIdx.x * warps_per_cta + warp + bstride);
    int p_row = 0;                       // next row to open
    int p_left = 0;                      // groups left in the open row
    int64_t p_off = 0, p_end = 0;        // next group's first entry; the row's end rounded up to 4 entries
    int n_issued = 0, n_consumed = 0;    // stream counters (groups)
    int p_stage = 0;                     // ring stage of the next group to issue
    // warp-uniform broadcast of lane r's value through REDUX: the result lives in a uniform register, so loops and
    // bulk-copy operands derived from it need no divergence checks (BRA.DIV) and no operand waterfall (R2UR loops)
    auto bcast = [&](unsigned v, int r) -> unsigned { return __reduce_or_sync(kFull, lane == r ? v : 0u); };
    auto bcast64 = [&](int64_t v, int r) -> int64_t {
        const unsigned lo = bcast((unsigned)v, r), hi = bcast((unsigned)((unsigned long long)v >> 32), r);
        return (int64_t)(((unsigned long long)hi << 32) | lo);
    };
    auto pump = [&]() {
#pragma unroll 1
        while (n_issued - n_consumed < nst) {
            while (p_left == 0 && p_row < 2 * kV4Rows) {               // open the next non-empty row
                const int r = p_row & 3;
                const bool in_cur = p_row < kV4Rows;
                const int len = (int)bcast((unsigned)(in_cur ? cur.len : nxt.len), r);
                ++p_row;
                if (len > 0) {
                    const int64_t s = bcast64(in_cur ? cur.s : nxt.s, r);
                    p_off = s & ~(int64_t)3;
                    p_end = (s + len + 3) & ~(int64_t)3;
                    p_left = (int)((p_end - p_off + kV4Group - 1) / kV4Group);
                }
            }
            if (p_left == 0) break;
            {
                const unsigned bytes = (unsigned)min((int64_t)kV4Group, p_end - p_off) * 4u;
                const unsigned bar = bars_addr + 8u * p_stage;
                const unsigned dst = ring_addr + (unsigned)p_stage * (8u * kV4Group);
                asm volatile(
                    "{\n"
                    ".reg .pred P1;\n"
                    "elect.sync _|P1, 0xffffffff;\n"
                    "@P1 mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
                    "@P1 cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%2], [%3], %4, [%0];\n"
                    "@P1 cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%5], [%6], %4, [%0];\n"
                    "}\n" ::"r"(bar), "r"(2u * bytes), "r"(dst), "l"(indices + p_off), "r"(bytes),
                    "r"(dst + 4u * kV4Group), "l"(counts + p_off) : "memory");
            }
            p_off += kV4Group;
            --p_left;
            ++n_issued;
            p_stage = p_stage + 1 == nst ? 0 : p_stage + 1;
        }
    };

    // one entry class (element c of every lane's 16-byte block): membership bit, library size, append (count, gene)
    unsigned wr_addr = list_addr;        // shared address of the next free list record
    const unsigned wr_last = list_addr + 8u * (unsigned)(cap - 32);   // a take needs room for 32 records
    auto take = [&](int g, float v, bool valid, float &lib) {
        const unsigned word = v4_lds_u32(bitmap_addr + 4u * (unsigned)(g >> 5));
        const bool sel = valid && (__funnelshift_r(word, word, g) & 1u) != 0u;
        const unsigned m = __ballot_sync(kFull, sel);
        if (sel) {
            v4_sts_rec(wr_addr + 8u * __popc(m & lt_mask), v, g);
            lib += v;
        }
        wr_addr += 8u * __popc(m);
    };

    v4u64 hv[NH * 16];                                                 // NH x 8 chunk blocks of 4 floats (as pairs)
    // entry e of the list: transform, weight, AXPY with the lane's chunk phase.  Fast mode writes (c, bucket) back for
    // the ||y_s||^2 passes and returns |c|; slow mode (single-row path) adds c into the float accumulator directly.
    auto axpy_entry = [&](int e, float scale, bool slow_mode) -> float {
        const float2 rec = list[e];
        const int gene = __float_as_int(rec.y);
        const int wi = gene >> 5;
        const int slot = (int)prefix[wi] + __popc(bitmap[wi] & ((1u << (gene & 31)) - 1u));
        const int2 bw = slot_bw[slot];
        const float c = v4_xform(rec.x, scale) * __int_as_float(bw.y);
        if (slow_mode) atomicAdd(reinterpret_cast<float *>(acc) + bw.x, c);
        else list[e] = make_float2(c, __int_as_float(bw.x));
        const unsigned a0 = xs_addr + (unsigned)bw.x * (XS * 4) + 16u * j8;
        const v4u64 cc = v4_pack(c, c);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            v4u64 x0, x1;
            v4_lds128(a0 ^ (16u * t), x0, x1);
            hv[2 * t] = v4_fma2(cc, x0, hv[2 * t]);
            hv[2 * t + 1] = v4_fma2(cc, x1, hv[2 * t + 1]);
            if (NH == 2) {
                v4_lds128((a0 ^ (16u * t)) + 128u, x0, x1);
                hv[16 + 2 * t] = v4_fma2(cc, x0, hv[16 + 2 * t]);
                hv[16 + 2 * t + 1] = v4_fma2(cc, x1, hv[16 + 2 * t + 1]);
            }
        }
        return fabsf(c);
    };
    // xor-shuffle reduction over the 8 lanes of a row: lane j ends with chunk j (and 8 + j) in hv[0..1] (hv[16..17])
    auto reduce8 = [&]() {
#pragma unroll
        for (int hf = 0; hf < NH; ++hf) {
            v4u64 *v = hv + 16 * hf;
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = v4_add2(v[i], v4_shfl_xor(v[i + 8], 4));
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = v4_add2(v[i], v4_shfl_xor(v[i + 4], 2));
#pragma unroll
            for (int i = 0; i < 2; ++i) v[i] = v4_add2(v[i], v4_shfl_xor(v[i + 2], 1));
        }
    };
    auto store_h = [&](int64_t orow) {                                 // lane j: chunk j (and 8 + j) of H[orow]
        float *out = h + orow * kp;
        if (4 * j8 < kp) {
            float4 o;
            v4_unpack(hv[0], o.x, o.y);
            v4_unpack(hv[1], o.z, o.w);
            *reinterpret_cast<float4 *>(out + 4 * j8) = o;
        }
        if (NH == 2 && 32 + 4 * j8 < kp) {
            float4 o;
            v4_unpack(hv[16], o.x, o.y);
            v4_unpack(hv[17], o.z, o.w);
            *reinterpret_cast<float4 *>(out + 32 + 4 * j8) = o;
        }
    };
    // four per-lane partial sums (one per row of the batch) -> every lane of group g gets the warp total of row g
    auto sum_rows = [&](float a0, float a1, float a2, float a3) -> float {
        const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
        float a = up16 ? a2 : a0, b = up16 ? a3 : a1;
        a += __shfl_xor_sync(kFull, up16 ? a0 : a2, 16);
        b += __shfl_xor_sync(kFull, up16 ? a1 : a3, 16);
        float c = up8 ? b : a;
        c += __shfl_xor_sync(kFull, up8 ? a : b, 8);
        c += __shfl_xor_sync(kFull, c, 4);
        c += __shfl_xor_sync(kFull, c, 2);
        c += __shfl_xor_sync(kFull, c, 1);
        return c;
    };

    // ---------------- main loop over batches of four rows ---------------------------------------------------
    int64_t bt = (int64_t)blockIdx.x * warps_per_cta + warp;
    int c_stage = 0;                     // ring stage of the next group to consume
    unsigned c_parity = 0u;
    pump();

    while (bt < n_batches) {
        const Raw ahead = load_raw(bt + 2 * bstride);                  // becomes `nxt` at the end of this batch
        unsigned slow = 0u;              // bit r: row r did not fit the list -> single-row path
        int stv = 0, endv = 0;           // lane r < 4: list range [stv, endv) of row r
        float libv = 0.f;                // lane r < 4: library size of row r
        wr_addr = list_addr;
#pragma unroll 1
        for (int r = 0; r < kV4Rows; ++r) {
            const int len = (int)bcast((unsigned)cur.len, r);
            const int srel = (int)(bcast((unsigned)cur.s, r) & 3u);
            const int ng = len > 0 ? (srel + len + kV4Group - 1) / kV4Group : 0;
            const unsigned row_st = wr_addr;
            float lib = 0.f;
            bool over = false;
#pragma unroll 1
            for (int k = 0; k < ng; ++k) {
                v4_mbar_wait(bars_addr + 8u * c_stage, c_parity);
                const unsigned sa = ring_addr + (unsigned)c_stage * (8u * kV4Group) + 16u * lane;
                const int4 g4 = v4_lds_i4(sa);
                const float4 v4 = v4_lds_f4(sa + 4u * kV4Group);
                const int jr = kV4Group * k + 4 * lane - srel;         // row-relative index of the lane's element 0
                if (wr_addr + 8u * (kV4Group - 32) <= wr_last) {       // room for a whole group: no per-take checks
                    const int lim = over ? 0 : len;
                    take(g4.x, v4.x, (unsigned)jr < (unsigned)lim, lib);
                    take(g4.y, v4.y, (unsigned)(jr + 1) < (unsigned)lim, lib);
                    take(g4.z, v4.z, (unsigned)(jr + 2) < (unsigned)lim, lib);
                    take(g4.w, v4.w, (unsigned)(jr + 3) < (unsigned)lim, lib);
                } else {                                               // near the end of the list: check every take;
                    over = over || wr_addr > wr_last;                  // a row that does not fit goes to the slow path
                    take(g4.x, v4.x, !over && (unsigned)jr < (unsigned)len, lib);
                    over = over || wr_addr > wr_last;
                    take(g4.y, v4.y, !over && (unsigned)(jr + 1) < (unsigned)len, lib);
                    over = over || wr_addr > wr_last;
                    take(g4.z, v4.z, !over && (unsigned)(jr + 2) < (unsigned)len, lib);
                    over = over || wr_addr > wr_last;
                    take(g4.w, v4.w, !over && (unsigned)(jr + 3) < (unsigned)len, lib);
                }
                ++n_consumed;
                if (++c_stage == nst) { c_stage = 0; c_parity ^= 1u; }
                pump();
            }
            lib = warp_sum(lib);
            if (over) { slow |= 1u << r; wr_addr = row_st; }
            if (lane == r) {
                stv = (int)(row_st - list_addr) >> 3;
                endv = (int)(wr_addr - list_addr) >> 3;
                libv = lib;
            }
            // next row's region: 64-byte aligned with 64-byte phase (r + 1) & 1 -> two wavefronts per record access
            wr_addr = (wr_addr + 63u) & ~63u;
            if ((((wr_addr - list_addr) >> 6) & 1u) != ((unsigned)(r + 1) & 1u)) wr_addr += 64u;
        }
        __syncwarp();

        const float my_lib0 = __shfl_sync(kFull, libv, grp);
        const float my_lib = my_lib0 == 0.f ? 1.f : my_lib0;
        const float my_scale = linear ? -1.f : 1e4f / my_lib;
        const int my_lo = __shfl_sync(kFull, stv, grp);
        const int my_hi = __shfl_sync(kFull, endv, grp);
        const int64_t my_orow = __shfl_sync(kFull, cur.orow, grp);
        const bool my_slow = (slow >> grp) & 1u;

        // ---- AXPY, eight lanes per row
#pragma unroll
        for (int i = 0; i < NH * 16; ++i) hv[i] = 0ull;
        float amax = 0.f;
        {
            const int n_it = ((int)__reduce_max_sync(kFull, (unsigned)(endv - stv)) + 7) >> 3;
#pragma unroll 1
            for (int it = 0; it < n_it; ++it) {
                const int e = my_lo + j8 + 8 * it;
                if (e < my_hi) amax = fmaxf(amax, axpy_entry(e, my_scale, false));
            }
        }
        reduce8();
        if (my_orow >= 0 && !my_slow) store_h(my_orow);
        // block fixed point of the row: n * max|c| < 2^(ex+1)  =>  every bucket sum * 2^(29-ex) < 2^30
        amax = fmaxf(amax, __shfl_xor_sync(kFull, amax, 4));
        amax = fmaxf(amax, __shfl_xor_sync(kFull, amax, 2));
        amax = fmaxf(amax, __shfl_xor_sync(kFull, amax, 1));
        const int bexp = (int)((__float_as_uint(amax * (float)(my_hi - my_lo)) >> 23) & 255u);
        const float q_scale = bexp >= 30 ? __uint_as_float((unsigned)(283 - bexp) << 23) : 0.f;
        const float q_inv = bexp >= 30 ? __uint_as_float((unsigned)(bexp - 29) << 23) : 0.f;
        __syncwarp();

        // ---- ||y_s||^2 of the four rows, one row at a time over the per-warp accumulator
        {
            float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll
            for (int r = 0; r < kV4Rows; ++r) {
                const int lo = __shfl_sync(kFull, stv, r), hi = __shfl_sync(kFull, endv, r);
                const float qs = __shfl_sync(kFull, q_scale, 8 * r), qi = __shfl_sync(kFull, q_inv, 8 * r);
#pragma unroll 1
                for (int e = lo + lane; e < hi; e += 32) {
                    const float2 rec = list[e];
                    atomicAdd(acc + __float_as_int(rec.y), __float2int_rn(rec.x * qs));
                }
                __syncwarp();
                float part = 0.f;
#pragma unroll 1
                for (int e = lo + lane; e < hi; e += 32) {
                    const float sb = (float)atomicExch(acc + __float_as_int(list[e].y), 0) * qi;
                    part = fmaf(sb, sb, part);
                }
                __syncwarp();
                if (r == 0) p0 = part; else if (r == 1) p1 = part; else if (r == 2) p2 = part; else p3 = part;
            }
            const float sq = sum_rows(p0, p1, p2, p3);
            if (j8 == 0 && my_orow >= 0 && !my_slow) ysq[my_orow] = sq;
        }

        // ---- rows that did not fit: single-row path (library size first, then list flushed whenever it fills)
        if (slow) {
#pragma unroll 1
            for (int r = 0; r < kV4Rows; ++r) {
                if (!((slow >> r) & 1u)) continue;
                const int64_t s = __shfl_sync(kFull, cur.s, r);
                const int len = __shfl_sync(kFull, cur.len, r);
                const int64_t orow = __shfl_sync(kFull, cur.orow, r);
                const int32_t *ip = indices + s;
                const float *vp = counts + s;
                float lib = 0.f;
                for (int j = lane; j < len; j += 32) {
                    const int g = ld_stream(ip + j);
                    if ((bitmap[g >> 5] >> (g & 31)) & 1u) lib += ld_stream(vp + j);
                }
                lib = warp_sum(lib);
                const float scale = linear ? -1.f : 1e4f / (lib == 0.f ? 1.f : lib);
#pragma unroll
                for (int i = 0; i < NH * 16; ++i) hv[i] = 0ull;
                float *facc = reinterpret_cast<float *>(acc);
                int j0 = 0;
#pragma unroll 1
                while (j0 < len) {
                    float dummy = 0.f;
                    wr_addr = list_addr;
                    for (; j0 < len && (int)(wr_addr - list_addr) + 8 * 32 <= 8 * cap; j0 += 32) {
                        const int j = j0 + lane;
                        const bool ok = j < len;
                        take(ok ? ld_stream(ip + j) : 0, ok ? ld_stream(vp + j) : 0.f, ok, dummy);
                    }
                    const int c2 = (int)(wr_addr - list_addr) >> 3;
                    __syncwarp();
                    for (int e = lane; e < c2; e += 32) axpy_entry(e, scale, true);
                    __syncwarp();
                }
                reduce8();
#pragma unroll
                for (int hf = 0; hf < NH; ++hf) {                      // fold the four lane groups
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        v4u64 &x = hv[16 * hf + i];
                        x = v4_add2(x, v4_shfl_xor(x, 8));
                        x = v4_add2(x, v4_shfl_xor(x, 16));
                    }
                }
                float sq = 0.f;
                for (int c = lane * 4; c < d; c += 128) {
                    const float4 a = *reinterpret_cast<float4 *>(facc + c);
                    *reinterpret_cast<float4 *>(facc + c) = make_float4(0.f, 0.f, 0.f, 0.f);
                    sq = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, sq))));
                }
                sq = warp_sum(sq);
                if (orow >= 0) {
                    if (grp == 0) store_h(orow);
                    if (lane == 0) ysq[orow] = sq;
                }
                __syncwarp();
            }
        }

        bt += bstride;
        cur = nxt;
        nxt = ahead;
        p_row = max(p_row, kV4Rows) - kV4Rows;
        pump();                                                        // rows of the new `nxt` can be opened now
    }
}

}  // namespace fdb
