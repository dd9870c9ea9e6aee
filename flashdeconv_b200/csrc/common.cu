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
// rows are 16-byte aligned (multiple of 4 floats); above 32 types they are padded to a multiple of 8 so that the
// half-precision gather rows of the sweep kernel are 16-byte aligned too
FDB_API int fdb_padded_types(int n_types)
{
    if (n_types <= 0) return 0;
    return (int)fdb::round_up(n_types, n_types > 32 ? 8 : 4);
}
