// Multi-GPU solve loop with the collectives issued from native code.
//
// One process per GPU.  The tile partition and halo maps come from the host layer (tiling.py); this file
// runs the per-sweep sequence
//     sweep own rows -> pack boundary rows -> grouped ncclSend/ncclRecv into the halo slices
//     -> ncclAllReduce(MAX) of the two max-norm words -> finalize (stop test)
// on ONE stream with no host synchronisation, so a 100-sweep solve is ~700 asynchronous enqueues instead of
// a Python round trip per collective.  NCCL is bound at run time (dlopen "libnccl.so.2": the copy torch
// already loaded when running under torch.distributed, else the system one), so libfdb200.so has no
// link-time NCCL dependency and still loads on a CPU-only box.
//
// Reference semantics: Jacobi sweep (core/solver.py:149-184) is order independent across spots, and the
// stop test uses GLOBAL max norms (core/solver.py:395-397) -> MAX all-reduce, not SUM.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>
#include "fdb_common.cuh"

extern "C" int fdb_bcd_sweep(const float *h, const float *host_gram, const float *beta_in, float *beta_out,
                             const int32_t *indptr, const int32_t *indices, int64_t n_rows, int32_t n_types,
                             float lambda, float rho_scaled, float tol, int32_t finalize, void *state, const void *plan,
                             void *stream);
extern "C" int fdb_bcd_finalize(void *state, float tol, void *stream);
extern "C" int fdb_bcd_init(float *beta, int64_t n_rows, int32_t n_types, void *state, void *stream);
extern "C" int fdb_rows_gather(const float *src, const int32_t *rows, int64_t n_list, int32_t row_floats, float *dst,
                               void *stream);

namespace fdb {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int load_nccl()
{
    if (g_nccl.handle) return FDB_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *hnd = nullptr;
    for (const char *nm : names)
        if ((hnd = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!hnd) {
        set_error("cannot load NCCL (libnccl.so.2): %s", dlerror());
        return FDB_ERR_UNSUPPORTED;
    }
#define FDB_NCCL_SYM(field, sym)                                                   \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(hnd, sym));      \
    if (!g_nccl.field) { set_error("NCCL symbol %s missing", sym); return FDB_ERR_UNSUPPORTED; }
    FDB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    FDB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    FDB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    FDB_NCCL_SYM(GroupStart, "ncclGroupStart")
    FDB_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    FDB_NCCL_SYM(Send, "ncclSend")
    FDB_NCCL_SYM(Recv, "ncclRecv")
    FDB_NCCL_SYM(AllReduce, "ncclAllReduce")
    FDB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef FDB_NCCL_SYM
    g_nccl.handle = hnd;
    return FDB_OK;
}

static int nccl_fail(ncclResult_t r, const char *what)
{
    set_error("NCCL error %d (%s) at %s", (int)r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?", what);
    return FDB_ERR_CUDA;
}

#define FDB_NCCL(call)                                                \
    do {                                                              \
        ncclResult_t r__ = (call);                                    \
        if (r__ != ncclSuccess) return fdb::nccl_fail(r__, #call);    \
    } while (0)

}  // namespace fdb

using namespace fdb;

#define FDB_API extern "C" __attribute__((visibility("default")))

FDB_API int fdb_comm_unique_id(char *host_id_128)
{
    FDB_REQUIRE(host_id_128 != nullptr, "null id buffer");
    int rc = load_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    FDB_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(host_id_128, id.internal, NCCL_UNIQUE_ID_BYTES);
    return FDB_OK;
}

FDB_API int fdb_comm_init(int32_t rank, int32_t world, const char *host_id_128, void **host_comm_out)
{
    FDB_REQUIRE(host_id_128 && host_comm_out && world > 0 && rank >= 0 && rank < world, "bad communicator arguments");
    int rc = load_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    memcpy(id.internal, host_id_128, NCCL_UNIQUE_ID_BYTES);
    ncclComm_t comm = nullptr;
    FDB_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
    *host_comm_out = comm;
    return FDB_OK;
}

FDB_API int fdb_comm_destroy(void *comm)
{
    if (!comm || !g_nccl.handle) return FDB_OK;
    FDB_NCCL(g_nccl.CommDestroy((ncclComm_t)comm));
    return FDB_OK;
}

// Halo exchange of `beta` (n_total x Kp): receive peers' rows into the halo slices, send packed boundary rows.
static int exchange(float *beta, int64_t n_own, int kp, int n_recv, const int32_t *recv_peer,
                    const int64_t *recv_first, const int64_t *recv_count, int n_send, const int32_t *send_peer,
                    const int32_t *const *send_rows, const int64_t *send_count, float *const *send_buf,
                    ncclComm_t comm, cudaStream_t st)
{
    for (int s = 0; s < n_send; ++s) {
        int rc = fdb_rows_gather(beta, send_rows[s], send_count[s], kp, send_buf[s], st);
        if (rc) return rc;
    }
    if (n_recv + n_send == 0) return FDB_OK;
    FDB_NCCL(g_nccl.GroupStart());
    for (int r = 0; r < n_recv; ++r)
        FDB_NCCL(g_nccl.Recv(beta + (n_own + recv_first[r]) * kp, (size_t)recv_count[r] * kp, ncclFloat32, recv_peer[r],
                             comm, st));
    for (int s = 0; s < n_send; ++s)
        FDB_NCCL(g_nccl.Send(send_buf[s], (size_t)send_count[s] * kp, ncclFloat32, send_peer[s], comm, st));
    FDB_NCCL(g_nccl.GroupEnd());
    return FDB_OK;
}

FDB_API int fdb_bcd_solve_tiled(const float *h, const float *host_gram, float *beta_a, float *beta_b,
                                const int32_t *indptr, const int32_t *indices, int64_t n_own, int64_t n_total,
                                int32_t n_types, float lambda, float rho_scaled, int32_t max_iter, float tol,
                                void *state, int32_t n_recv, const int32_t *host_recv_peer,
                                const int64_t *host_recv_first, const int64_t *host_recv_count, int32_t n_send,
                                const int32_t *host_send_peer, const int32_t *const *host_send_rows,
                                const int64_t *host_send_count, float *const *host_send_buf, void *comm,
                                const void *plan, void *stream)
{
    FDB_REQUIRE(comm != nullptr && state != nullptr, "null communicator / state");
    FDB_REQUIRE(n_own >= 0 && n_total >= n_own && max_iter >= 0, "bad sizes");
    FDB_REQUIRE(n_types >= 1 && n_types <= FDB_MAX_TYPES, "n_types must be in [1, %d], got %d", FDB_MAX_TYPES, n_types);
    int rc = load_nccl();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int kp = fdb_padded_types(n_types);
    rc = fdb_bcd_init(beta_a, n_total, n_types, state, stream);          // halo rows start at 1/K like everyone else
    if (rc) return rc;
    rc = fdb_bcd_init(beta_b, n_total, n_types, nullptr, stream);
    if (rc) return rc;
    float *cur = beta_a, *nxt = beta_b;
    for (int it = 0; it < max_iter; ++it) {
        if (n_own > 0) {
            rc = fdb_bcd_sweep(h, host_gram, cur, nxt, indptr, indices, n_own, n_types, lambda, rho_scaled, tol, 0,
                               state, plan, stream);
            if (rc) return rc;
        }
        rc = exchange(nxt, n_own, kp, n_recv, host_recv_peer, host_recv_first, host_recv_count, n_send, host_send_peer,
                      host_send_rows, host_send_count, host_send_buf, (ncclComm_t)comm, st);
        if (rc) return rc;
        // words [0], [1] of the state block: bit patterns of non-negative floats order like unsigned integers
        FDB_NCCL(g_nccl.AllReduce(state, state, 2, ncclUint32, ncclMax, (ncclComm_t)comm, st));
        rc = fdb_bcd_finalize(state, tol, stream);
        if (rc) return rc;
        float *t = cur; cur = nxt; nxt = t;
    }
    return FDB_OK;
}
