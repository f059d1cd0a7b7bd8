// Does a second pipe issue in the shadow of a DFMA on sm_100a?
// A warp-wide DFMA occupies the 16-lane fp64 pipe of an SMSP for 2 cycles.  Each test interleaves
// 8 independent DFMA chains with K instructions of another pipe per 8 DFMAs and reports cycles per
// 8-DFMA group per SMSP (16 = the fp64 pipe is the only limiter; 16 + K = the companion
// instructions cost an issue cycle each ON TOP of the two cycles of every DFMA).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench_issue microbench_issue.cu
#include <cstdio>
#include <cuda_runtime.h>

enum Mix { NONE, IMAD, LOP, LDS64, MOVS, FFMA, DREG };

template <int MIXK, int K>
__global__ void __launch_bounds__(512) kern(double *out, long long *cyc, int iters, double a, double b, int ia)
{
    __shared__ double sm[512 * 4];
    double r[8];
    int q[16];
    double l[4] = {0, 0, 0, 0};
    float f[8];
    double breg[8], creg[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r[i] = threadIdx.x * 1e-3 + i * a;
        f[i] = threadIdx.x + i;
        breg[i] = a + i * 1e-9;
        creg[i] = b + i * 1e-12;
    }
#pragma unroll
    for (int i = 0; i < 16; i++) q[i] = threadIdx.x + i;
    for (int i = 0; i < 4; i++) sm[threadIdx.x * 4 + i] = i;
    __syncthreads();
    const double *mine = sm + threadIdx.x;   // conflict-free 64-bit loads
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MIXK == DREG) r[i] = fma(r[i], breg[i], creg[(i + 3) & 7]);
            else if (MIXK == MOVS) r[i] = fma(r[i], breg[i], b);
            else r[i] = fma(r[i], a, b);
            // K companions spread over the 8 DFMAs
#pragma unroll
            for (int j = 0; j < (K * (i + 1)) / 8 - (K * i) / 8; j++) {
                const int s = (K * i) / 8 + j;
                if (MIXK == IMAD) q[s & 15] = q[s & 15] * ia + 3;
                if (MIXK == LOP) q[s & 15] = (q[s & 15] ^ ia) & it;
                if (MIXK == LDS64) {
                    double v;
                    asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v)
                                 : "r"((unsigned)__cvta_generic_to_shared(mine + 512 * (s & 3))) : "memory");
                    l[s & 3] = v;
                }
                if (MIXK == FFMA) f[s & 7] = fmaf(f[s & 7], 1.0001f, 0.5f);
            }
        }
        if (MIXK == MOVS) {
            // register rotation of the DFMA chains in a rolled loop (as the action kernel's U / Vp):
            // K/2 doubles move by one slot = K 32-bit moves
            const double r0 = r[0];
#pragma unroll
            for (int i = 0; i + 1 < K / 2; i++) r[i] = r[i + 1];
            r[K / 2 - 1] = r0;
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += r[i] + f[i];
#pragma unroll
    for (int i = 0; i < 16; i++) s += q[i];
#pragma unroll
    for (int i = 0; i < 4; i++) s += l[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MIXK, int K>
void run(const char *name, int warps_per_sm)
{
    double *out;
    long long *cyc, h;
    cudaMalloc(&out, 8 * 512 * 148);
    cudaMalloc(&cyc, 8);
    const int iters = 4096;
    for (int rep = 0; rep < 2; rep++) kern<MIXK, K><<<148, 32 * warps_per_sm>>>(out, cyc, iters, 1.0000001, 1e-9, 3);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double wps = warps_per_sm / 4.0;
    // cycles per 8-DFMA group per SMSP (all of its warps together)
    const double per = (double)h / iters / wps;
    printf("%-6s K=%2d warps/SMSP %.0f: %6.2f cycles per (8 DFMA + K) per SMSP   [fp64 floor 16, serial %d]\n", name, K,
           wps, per, 16 + K);
    cudaFree(out);
    cudaFree(cyc);
}

template <int MIXK>
void sweep(const char *name)
{
    for (int w : {4, 12}) {
        run<MIXK, 2>(name, w);
        run<MIXK, 4>(name, w);
        run<MIXK, 8>(name, w);
        run<MIXK, 16>(name, w);
    }
}

int main()
{
    for (int w : {4, 8, 12}) run<NONE, 0>("none", w);
    for (int w : {4, 8, 12}) run<DREG, 0>("dreg", w);
    sweep<IMAD>("imad");
    sweep<LOP>("lop3");
    sweep<FFMA>("ffma");
    sweep<MOVS>("mov");
    sweep<LDS64>("lds64");
    return 0;
}
