// Microbenchmark: fp64 FMA pipe vs fp64 tensor (DMMA, mma.sync m8n8k4) on sm_100a,
// alone and mixed.  Answers "are they separate pipes on B200?" for the design
// of the high-order path.  nvcc -arch=sm_100a -O3 -o microbench_fp64
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double *out, int iters, double a, double b)
{
    double r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) r[i] = fma(r[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void k_dmma(double *out, int iters, double a, double b)
{
    double c[8];
#pragma unroll
    for (int i = 0; i < 8; i++) c[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) dmma(c[2 * i], c[2 * i + 1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ratio: per iteration 4 DMMA (4*256 FMA/warp) and NF*8 DFMA (NF*8*32 FMA/warp)
template <int NF>
__global__ void k_mixed(double *out, int iters, double a, double b)
{
    double c[8], r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i] = threadIdx.x * 1e-3 + i; r[i] = c[i] + 1; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) dmma(c[2 * i], c[2 * i + 1], a, b);
#pragma unroll
        for (int j = 0; j < NF; j++)
#pragma unroll
            for (int i = 0; i < 8; i++) r[i] = fma(r[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i] + r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float timeit(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int threads = 256, blocks = sms * 8, iters = 20000;
    double *out; cudaMalloc(&out, sizeof(double) * threads * blocks);
    double nthreads = (double)threads * blocks;
    for (int warps = 0; warps < 1; warps++) {
        float ms = timeit([&] { k_dfma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("DFMA only : %.2f TFLOP/s\n", nthreads * iters * 8 * 2 / ms / 1e9);
        ms = timeit([&] { k_dmma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("DMMA only : %.2f TFLOP/s\n", (nthreads / 32) * iters * 4 * 256 * 2 / ms / 1e9);
        ms = timeit([&] { k_mixed<1><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("mixed 4 DMMA : 8 DFMA  -> DMMA %.2f + DFMA %.2f TFLOP/s\n",
               (nthreads / 32) * iters * 4 * 256 * 2 / ms / 1e9, nthreads * iters * 8 * 2 / ms / 1e9);
        ms = timeit([&] { k_mixed<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("mixed 4 DMMA : 32 DFMA -> DMMA %.2f + DFMA %.2f TFLOP/s\n",
               (nthreads / 32) * iters * 4 * 256 * 2 / ms / 1e9, nthreads * iters * 32 * 2 / ms / 1e9);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
