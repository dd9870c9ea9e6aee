// One row width of the production sweep kernel per translation unit: nvcc -DFDB_P_KP=<Kp> (see build.py).
#include "bcd_p.cuh"
#ifndef FDB_P_KP
#error "compile with -DFDB_P_KP=<8|16|24|32|40|48|56|64>"
#endif
namespace fdb {
template int launch_sweep_p<FDB_P_KP>(const float *, const GramArg<FDB_P_KP> &, int, const float *, float *, const int32_t *,
                                      const int32_t *, int64_t, float, float, float, int, SolveState *, const void *,
                                      cudaStream_t, const SweepComm *);
}
