// DFMA dependent-issue latency and throughput vs ILP / warps per SMSP on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double *out, long long *cyc, int iters, double a, double b)
{
    double r[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) r[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) r[i] = fma(r[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
void run(int warps_per_sm)
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 8 * 2048 * 200); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    // one CTA per SM with `warps_per_sm` warps (4 SMSPs): warps/SMSP = warps_per_sm/4
    k<ILP><<<148, 32 * warps_per_sm>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    k<ILP><<<148, 32 * warps_per_sm>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / ((double)iters * ILP);
    printf("ILP %2d warps/SM %2d (%.2f/SMSP): %.2f cycles per DFMA per warp -> pipe util %.0f%%\n", ILP,
           warps_per_sm, warps_per_sm / 4.0, per, 100.0 * (warps_per_sm / 4.0) * 2.0 / per);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    for (int w : {4, 8, 12, 16}) {
        run<1>(w); run<2>(w); run<3>(w); run<4>(w); run<6>(w); run<8>(w); run<16>(w);
    }
    return 0;
}
