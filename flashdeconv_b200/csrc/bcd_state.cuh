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
    // overlapped multi-GPU mode (bcd_p.cuh, SweepComm): sweep t accumulates its max norms into slot t & 1 while the
    // hand-shake of sweep t - 1 consumes the other one; max|beta_new| of this rank rotates over three slots because
    // sweep t + 1 still reads the value of sweep t (range of its fp16 tile) while the hand-shake of t - 1 recycles one
    unsigned ov_diff[2];
    unsigned ov_abs[2];
    unsigned ov_new[3];
    unsigned hs_done;               // sequence number of the last completed hand-shake
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
