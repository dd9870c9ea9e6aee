// Multi-GPU solve loop over NVLink peer memory (no NCCL on the data path).
//
// Every rank's beta buffers live in a symmetric allocation that all peers have mapped (the host layer
// obtains the mapping, e.g. torch symmetric memory / CUDA IPC, and passes the peers' base pointers).
// Per sweep, on ONE stream and without host synchronisation, ONE launch (bcd_p.cuh, SweepComm):
//   1. the sweep kernel walks the patches that hold boundary rows first and writes those rows STRAIGHT into the
//      neighbours' halo slots over NVLink from its store phase (128-byte rows, 16-byte stores), overlapping the
//      interior patches;
//   2. its last CTA publishes this rank's two max-norm words and a sweep sequence number into every peer's comm
//      block (fence.sys + release store), spins until every peer's sequence number has arrived here (acquire
//      loads), then reduces the max norms and runs the stop test -- the MAX all-reduce of core/solver.py:395-397
//      and the halo hand-shake.
// (With the fp32-gather fallback kernel the same three steps run as three launches: sweep, peer_push_kernel,
// peer_sync_kernel.)
// Ordering argument (two beta buffers X, Y alternate): a rank starts sweep t+1 only after every peer's
// flag t, which a peer writes after its sweep t and push t completed; so rows pushed for sweep t+1 can never
// overwrite halo rows a peer is still reading in sweep t, and the rows read in sweep t+2 were pushed
// before flag t+1.  The statistics slots are double-buffered by sweep parity for the same reason.
//
// Symmetric buffer layout (floats): [beta_a: cap_rows*Kp][beta_b: cap_rows*Kp][comm: kCommWords u32]
//   comm: flags[kMaxRanks], stats[2][kMaxRanks][kCommStatWords]
#include "bcd_common.cuh"

extern "C" int fdb_bcd_sweep(const float *h, const float *host_gram, const float *beta_in, float *beta_out,
                             const int32_t *indptr, const int32_t *indices, int64_t n_rows, int32_t n_types,
                             float lambda, float rho_scaled, float tol, int32_t finalize, void *state, const void *plan,
                             void *stream);
extern "C" int fdb_bcd_init(float *beta, int64_t n_rows, int32_t n_types, void *state, void *stream);

namespace fdb {

bool sweep_can_fuse_comm(const float *host_gram, int n_types, float lam, const void *plan);
int sweep_with_comm(const float *h, const float *host_gram, const float *beta_in, float *beta_out, const int32_t *indptr,
                    const int32_t *indices, int64_t n_rows, int32_t n_types, float lam, float rho, float tol, void *state,
                    const void *plan, void *stream, const SweepComm &comm);


struct PeerBases {
    float *base[kMaxRanks];
};

// unfused fallback (fp32-gather sweep kernel): one thread per (own row, 16-byte chunk); rows without entries cost a load
__global__ void __launch_bounds__(256)
peer_push_kernel(const float *__restrict__ beta_local, PeerBases pb, int64_t buf_off, const int32_t *__restrict__ push_ptr,
                 const int2 *__restrict__ push_ent, int64_t n_own, int chunks, const SolveState *st)
{
    if (st->converged) return;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_own * chunks) return;
    const int64_t row = t / chunks;
    const int q = (int)(t - row * chunks);
    const int pe = push_ptr[row + 1];
    for (int u = push_ptr[row]; u < pe; ++u) {
        const int2 ent = push_ent[u];
        const float4 v = *reinterpret_cast<const float4 *>(beta_local + (row * chunks + q) * 4);
        *reinterpret_cast<float4 *>(pb.base[ent.x] + buf_off + ((int64_t)ent.y * chunks + q) * 4) = v;
    }
}

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// start-of-solve rendezvous (do_finalize = 0) and the hand-shake of the unfused fallback loop
__global__ void __launch_bounds__(kMaxRanks)
peer_sync_kernel(SolveState *st, PeerBases pb, int64_t comm_off, int rank, int world, unsigned seq, float tol,
                 int do_finalize)
{
    if (st->converged) return;                       // every rank converges at the same sweep (same reduced norms)
    const int peer = threadIdx.x;
    const int parity = (int)(seq & 1u);
    __shared__ int timed_out;
    if (threadIdx.x == 0) timed_out = 0;
    __syncthreads();
    if (peer < world) {
        unsigned *pc = reinterpret_cast<unsigned *>(pb.base[peer] + comm_off);       // the peer's comm block
        unsigned *slot = pc + kMaxRanks + (parity * kMaxRanks + rank) * kCommStatWords;
        __threadfence_system();                      // order after this rank's sweep + push (earlier kernels)
        slot[0] = st->max_diff_bits;
        slot[1] = st->max_abs_bits;
        __threadfence_system();
        st_release_sys(pc + rank, seq);
        const unsigned *mine = reinterpret_cast<const unsigned *>(pb.base[rank] + comm_off);
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(mine + peer) - seq) < 0) {
            if (clock64() - t0 > 120000000000LL) { timed_out = 1; break; }           // ~60 s: a peer is gone
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (timed_out) { st->converged = 2; return; }                                // surfaced as an error by the host
        if (!do_finalize) return;                                                    // start-of-solve rendezvous only
        const unsigned *mine = reinterpret_cast<const unsigned *>(pb.base[rank] + comm_off);
        unsigned md = 0u, ma = 0u;
        for (int p = 0; p < world; ++p) {
            md = max(md, ld_acquire_sys(mine + kMaxRanks + (parity * kMaxRanks + p) * kCommStatWords));
            ma = max(ma, ld_acquire_sys(mine + kMaxRanks + (parity * kMaxRanks + p) * kCommStatWords + 1));
        }
        st->max_diff_bits = md;
        st->max_abs_bits = ma;
        finalize_state(st, tol);
    }
}

}  // namespace fdb

using namespace fdb;

#define FDB_API extern "C" __attribute__((visibility("default")))

FDB_API int64_t fdb_peer_comm_floats(void) { return 256; }

FDB_API int fdb_bcd_solve_peer(const float *h, const float *host_gram, void *const *host_peer_base, int32_t rank,
                               int32_t world, int64_t cap_rows, const int32_t *indptr, const int32_t *indices,
                               int64_t n_own, int64_t n_total, int32_t n_types, float lambda, float rho_scaled,
                               int32_t max_iter, float tol, void *state, const int32_t *push_ptr, const void *push_ent,
                               const int32_t *patch_order, const int32_t *n_boundary, uint32_t seq_base,
                               const void *plan, void *stream)
{
    FDB_REQUIRE(host_peer_base && state, "null peer table / state");
    FDB_REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "world must be in [1, %d]", kMaxRanks);
    FDB_REQUIRE(n_own >= 0 && n_total >= n_own && n_total <= cap_rows && max_iter >= 0, "bad sizes");
    FDB_REQUIRE(n_types >= 1 && n_types <= FDB_MAX_TYPES, "n_types must be in [1, %d], got %d", FDB_MAX_TYPES, n_types);
    FDB_REQUIRE(world == 1 || (push_ptr && push_ent), "null push lists");
    cudaStream_t st = (cudaStream_t)stream;
    const int kp = fdb_padded_types(n_types);
    PeerBases pb;
    for (int p = 0; p < kMaxRanks; ++p) pb.base[p] = p < world ? (float *)host_peer_base[p] : nullptr;
    const int64_t off_a = 0, off_b = cap_rows * kp, off_comm = 2 * cap_rows * kp;
    float *mine = pb.base[rank];
    // beta_a (own + halo rows) starts at 1/K.  beta_b is NOT initialised: its own rows are written by sweep 1 and
    // its halo rows by the peers' first push, which may arrive before this rank gets here.
    int rc = fdb_bcd_init(mine + off_a, n_total, n_types, state, stream);
    if (rc) return rc;
    // start-of-solve rendezvous: nobody pushes rows before every rank has initialised its buffers and finished
    // reading the previous solve's result (those kernels precede this one in stream order)
    peer_sync_kernel<<<1, kMaxRanks, 0, st>>>((SolveState *)state, pb, off_comm, rank, world, seq_base, tol, 0);
    FDB_LAUNCH_CHECK("peer_sync_kernel");
    // one launch per sweep when the production sweep kernel runs (push + hand-shake fused in); otherwise three
    const bool fused = n_own > 0 && world > 1 && sweep_can_fuse_comm(host_gram, n_types, lambda, plan);
    SweepComm cm;
    cm.patch_order = patch_order;
    cm.n_boundary = n_boundary;
    cm.push_ptr = push_ptr;
    cm.push_ent = (const int2 *)push_ent;
    for (int p = 0; p < kMaxRanks; ++p) cm.peer_base[p] = pb.base[p];
    cm.comm_off = off_comm;
    cm.rank = rank;
    cm.world = world;
    cm.debug = 0;
    int64_t cur = off_a, nxt = off_b;
    const int chunks = kp / 4;
    if (fused && n_boundary != nullptr) {
        // overlapped form: launch t computes sweep t and carries the hand-shake that closes sweep t - 1 (block 0); a closing
        // launch does the hand-shake of the last sweep.  Flag value of launch t: seq_base + t.
        cm.debug = getenv("FDB_PEER_DEBUG") ? atoi(getenv("FDB_PEER_DEBUG")) : 0;
        for (int it = 1; it <= max_iter + 1; ++it) {
            cm.out_off = nxt;
            cm.seq = seq_base + (unsigned)it;
            cm.sweep = it;
            cm.hs_only = it == max_iter + 1;
            cm.finalize_prev = it > 1;
            if (cm.hs_only && max_iter == 0) break;
            rc = sweep_with_comm(h, host_gram, mine + cur, mine + nxt, indptr, indices, n_own, n_types, lambda, rho_scaled,
                                 tol, state, plan, stream, cm);
            if (rc) return rc;
            const int64_t t = cur; cur = nxt; nxt = t;
        }
        return FDB_OK;
    }
    for (int it = 0; it < max_iter; ++it) {
        const unsigned seq = seq_base + (unsigned)it + 1u;
        if (n_own > 0) {
            rc = fdb_bcd_sweep(h, host_gram, mine + cur, mine + nxt, indptr, indices, n_own, n_types, lambda,
                               rho_scaled, tol, 0, state, plan, stream);
            if (rc) return rc;
            if (world > 1) {
                peer_push_kernel<<<(int)ceil_div(n_own * chunks, 256), 256, 0, st>>>(
                    mine + nxt, pb, nxt, push_ptr, (const int2 *)push_ent, n_own, chunks, (const SolveState *)state);
                FDB_LAUNCH_CHECK("peer_push_kernel");
            }
        }
        peer_sync_kernel<<<1, kMaxRanks, 0, st>>>((SolveState *)state, pb, off_comm, rank, world, seq, tol, 1);
        FDB_LAUNCH_CHECK("peer_sync_kernel");
        const int64_t t = cur; cur = nxt; nxt = t;
    }
    return FDB_OK;
}
