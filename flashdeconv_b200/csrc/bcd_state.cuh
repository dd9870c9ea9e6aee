// Solver state block shared by the sweep kernels (bcd.cu) and the peer-memory exchange (peer.cu).
#pragma once
#include "fdb_common.cuh"

namespace fdb {

struct SolveState {                 // mirrors the 64-byte block documented in fdb200.h
    unsigned max_diff_bits;
    unsigned max_abs_bits;
    unsigned arrived;
    int sweeps;
    int converged;
    float rel_change;
    float last_max_diff;
    float last_max_abs;
    int pad[8];
};
static_assert(sizeof(SolveState) == 64, "state block is 64 bytes");

__device__ __forceinline__ void finalize_state(SolveState *st, float tol)
{
    const float md = __uint_as_float(st->max_diff_bits), ma = __uint_as_float(st->max_abs_bits);
    const float rel = md / (ma + 1e-10f);
    st->rel_change = rel;
    st->last_max_diff = md;
    st->last_max_abs = ma;
    st->sweeps += 1;
    if (rel < tol) st->converged = 1;
    st->max_diff_bits = 0u;
    st->max_abs_bits = 0u;
    st->arrived = 0u;
}

}  // namespace fdb
