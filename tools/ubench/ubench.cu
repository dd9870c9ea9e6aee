// Microbenchmarks behind two DESIGN.md decisions about the wide-row sweep (Kp >= 40):
//   A  latency of a dependent constant-bank load chain as a function of footprint (L1 / L1.5 constant cache sizes)
//   B  throughput of warp-broadcast LDS.32 / LDS.128 against conflict-free LDS.128 (is a broadcast 128-bit load 1 or 4 cycles?)
//   C  streaming LDCU.128-style reads of a Gram-sized constant region by 12 warps at different phases
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/ubench tools/ubench/ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

__constant__ int c_chain[8192];           // 32 KB
__constant__ float4 c_stream[1984];       // 31 KB

__global__ void chase(int start, int iters, long long *out, int *sink)
{
    int idx = start;
    // warm
    for (int i = 0; i < iters; ++i) idx = c_chain[idx];
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) idx = c_chain[idx];
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; sink[0] = idx; }
}

template <int MODE>
__global__ void lds_tp(int iters, long long *out, float *sink)
{
    __shared__ float4 buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_float4(i, 1.f, 2.f, 3.f);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    float4 a0 = make_float4(0, 0, 0, 0), a1 = a0, a2 = a0, a3 = a0;
    float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int base = (threadIdx.x >> 5) * 8;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {          // broadcast LDS.128
            const float4 *p = buf + ((base + i) & 511);
            float4 v0 = p[0], v1 = p[64], v2 = p[128], v3 = p[192];
            a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
            a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
            a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
            a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
        } else if (MODE == 1) {   // conflict-free distinct LDS.128
            const float4 *p = buf + ((base + i) & 255) + lane;
            float4 v0 = p[0], v1 = p[64], v2 = p[128], v3 = p[192];
            a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
            a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
            a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
            a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
        } else if (MODE == 2) {   // broadcast LDS.32
            const float *p = reinterpret_cast<const float *>(buf) + ((base + i) & 511);
            s0 += p[0]; s1 += p[256]; s2 += p[512]; s3 += p[768];
        } else {                  // broadcast LDS.64
            const float2 *p = reinterpret_cast<const float2 *>(buf) + ((base + i) & 511);
            float2 v0 = p[0], v1 = p[128], v2 = p[256], v3 = p[384];
            s0 += v0.x + v0.y; s1 += v1.x + v1.y; s2 += v2.x + v2.y; s3 += v3.x + v3.y;
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = a0.x + a0.y + a0.z + a0.w + a1.x + a1.y + a1.z + a1.w + a2.x + a2.y + a2.z + a2.w +
                                                  a3.x + a3.y + a3.z + a3.w + s0 + s1 + s2 + s3;
}

// every warp walks the first SPAN float4 of c_stream with compile-time addresses (LDCU.128 + uniform-register operands, as
// in the sweep kernel's descent), warps staggered by `stagger` cycles so that they are at different places of the region
template <int SPAN>
__global__ void cstream(int rounds, int stagger, long long *out, float *sink)
{
    const int warp = threadIdx.x >> 5;
    float x = threadIdx.x * 1e-3f;
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    const long long w0 = clock64();
    while (clock64() - w0 < (long long)warp * stagger) { }
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < rounds; ++r) {
#pragma unroll
        for (int u = 0; u < SPAN; ++u) {
            const float4 g = c_stream[u];
            a0 = fmaf(g.x, x, a0); a1 = fmaf(g.y, x, a1); a2 = fmaf(g.z, x, a2); a3 = fmaf(g.w, x, a3);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}

template <int SPAN>
void run_cstream(long long *d_out, float *d_fs)
{
    long long h;
    for (int warps : {1, 12, 21}) {
        for (int stagger : {3000}) {
            const int rounds = 64;
            cstream<SPAN><<<1, warps * 32>>>(rounds, stagger, d_out, d_fs);
            cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
            printf("  region %5.1f KB warps %2d stagger %4d: %.2f cycles per LDCU.128 (+4 FFMA) for warp 0\n", SPAN * 16 / 1024.0, warps, stagger,
                   (double)h / ((double)rounds * SPAN));
        }
    }
}

int main()
{
    long long *d_out, h_out[4];
    int *d_sink;
    float *d_fs;
    cudaMalloc(&d_out, 64);
    cudaMalloc(&d_sink, 64);
    cudaMalloc(&d_fs, 4 * 1024 * 1024);
    static int chain[8192];
    printf("A: dependent constant loads, stride 64 B\n");
    for (int kb : {1, 2, 3, 4, 6, 8, 12, 16, 24, 32}) {
        const int n = kb * 1024 / 4, step = 16;
        for (int i = 0; i < 8192; ++i) chain[i] = 0;
        for (int i = 0; i < n; i += step) chain[i] = (i + step) % n;
        cudaMemcpyToSymbol(c_chain, chain, sizeof(chain));
        const int iters = 4096;
        chase<<<1, 32>>>(0, iters, d_out, d_sink);
        cudaMemcpy(h_out, d_out, 8, cudaMemcpyDeviceToHost);
        printf("  footprint %2d KB: %.1f cycles/load\n", kb, (double)h_out[0] / iters);
    }
    printf("B: shared-memory load throughput, one CTA of 1024 threads on one SM (4 loads per thread per iteration)\n");
    const char *names[4] = {"LDS.128 broadcast", "LDS.128 distinct ", "LDS.32  broadcast", "LDS.64  broadcast"};
    for (int mode = 0; mode < 4; ++mode) {
        const int iters = 4096;
        if (mode == 0) lds_tp<0><<<1, 1024>>>(iters, d_out, d_fs);
        if (mode == 1) lds_tp<1><<<1, 1024>>>(iters, d_out, d_fs);
        if (mode == 2) lds_tp<2><<<1, 1024>>>(iters, d_out, d_fs);
        if (mode == 3) lds_tp<3><<<1, 1024>>>(iters, d_out, d_fs);
        cudaMemcpy(h_out, d_out, 8, cudaMemcpyDeviceToHost);
        printf("  %s: %.2f cycles per warp-instruction (SM-wide)\n", names[mode], (double)h_out[0] / ((double)iters * 4 * 32));
    }
    printf("C: 12 warps (3 CTAs x 4 warps on one SM would need 3 blocks; here 1 CTA of W warps) streaming a constant region\n");
    static float4 st[1984];
    for (int i = 0; i < 1984; ++i) st[i] = make_float4(1e-3f * i, 1.f, 0.5f, 0.25f);
    cudaMemcpyToSymbol(c_stream, st, sizeof(st));
    run_cstream<64>(d_out, d_fs);
    run_cstream<128>(d_out, d_fs);
    run_cstream<192>(d_out, d_fs);
    run_cstream<224>(d_out, d_fs);
    run_cstream<256>(d_out, d_fs);
    run_cstream<288>(d_out, d_fs);
    run_cstream<320>(d_out, d_fs);
    run_cstream<352>(d_out, d_fs);
    run_cstream<400>(d_out, d_fs);
    run_cstream<800>(d_out, d_fs);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
