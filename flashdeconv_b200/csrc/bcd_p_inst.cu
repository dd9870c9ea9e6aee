// One row width and one form (single-GPU / multi-GPU extension) of the production sweep kernel per translation unit:
// nvcc -DFDB_P_KP=<Kp> -DFDB_P_COMM=<0|1> (see build.py).
#include "bcd_p.cuh"
#if !defined(FDB_P_KP) || !defined(FDB_P_COMM)
#error "compile with -DFDB_P_KP=<8|16|24|32|40|48|56|64> -DFDB_P_COMM=<0|1>"
#endif
namespace fdb {
template int launch_sweep_p<FDB_P_KP, (FDB_P_COMM != 0)>(const float *, const GramArg<FDB_P_KP> &, int, const float *, float *,
                                                         const int32_t *, const int32_t *, int64_t, float, float, float, int,
                                                         SolveState *, const void *, cudaStream_t, const SweepComm *);
}
