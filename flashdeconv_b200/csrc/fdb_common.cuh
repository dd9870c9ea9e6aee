// Shared helpers for libfdb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/fdb200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libfdb200 is written for sm_100a (B200) only"
#endif

namespace fdb {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kNumSM = 148;                       // B200

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define FDB_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) return fdb::cuda_fail(e__, #call);            \
    } while (0)

extern long long g_launches;      // kernels launched through this library (bench.py's gpu_launches)

#define FDB_LAUNCH_CHECK(name)                                                \
    do {                                                                      \
        ++fdb::g_launches;                                                    \
        cudaError_t e__ = cudaGetLastError();                                 \
        if (e__ != cudaSuccess) return fdb::cuda_fail(e__, name);             \
    } while (0)

#define FDB_REQUIRE(cond, ...)                                                \
    do {                                                                      \
        if (!(cond)) {                                                        \
            fdb::set_error(__VA_ARGS__);                                      \
            return FDB_ERR_ARG;                                               \
        }                                                                     \
    } while (0)

inline void count_launches(int n) { g_launches += n; }

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// streaming (read-once) loads: keep them out of L1 so the tables stay resident
__device__ __forceinline__ int ld_stream(const int *p)
{
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_stream(const float *p)
{
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// device-wide exclusive scan of int32 (graph.cu): out has n + 1 entries (out[n] = total), `in` and `out` may alias;
// block_sums needs scan_blocks(n) words
int exclusive_scan(const int32_t *in, int64_t n, int32_t *out, int32_t *block_sums, cudaStream_t st);
int64_t scan_blocks(int64_t n);

template <typename IndPtr>
__device__ __forceinline__ int64_t load_ptr(const IndPtr *p, int64_t i) { return (int64_t)p[i]; }

}  // namespace fdb
