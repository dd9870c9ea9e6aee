// Error plumbing + tiny ABI helpers.
#include <stdarg.h>
#include "fdb_common.cuh"

namespace fdb {
static thread_local char g_err[512] = "";
long long g_launches = 0;

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what)
{
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return FDB_ERR_CUDA;
}
}  // namespace fdb

#define FDB_API extern "C" __attribute__((visibility("default")))
FDB_API int fdb_abi_version(void) { return 1; }
FDB_API const char *fdb_last_error(void) { return fdb::g_err; }
FDB_API long long fdb_launch_count(void) { return fdb::g_launches; }
// rows are padded to a multiple of 8 floats: 32-byte aligned fp32 rows and 16-byte aligned half-precision gather
// rows, so that every K runs the halo-staged sweep kernel (padding columns are exactly 0 everywhere)
FDB_API int fdb_padded_types(int n_types)
{
    if (n_types <= 0) return 0;
    return (int)fdb::round_up(n_types, 8);
}
