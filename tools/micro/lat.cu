// Latency microbenchmarks (one warp): dependent DFMA / DADD chains, SHFL + DFMA chains, LDS, REDUX.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu && ./lat
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, double x, double y)
{
    __shared__ double sm[64];
    sm[threadIdx.x] = x; sm[threadIdx.x + 32] = y;
    __syncthreads();
    double a = x + threadIdx.x;
    long long t0, t1;
    const int N = 512;
    // 0: DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) a = fma(a, y, x);
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = (t1 - t0);
    // 1: DADD chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) a = a + y;
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = (t1 - t0);
    // 2: SHFL.UP(double) + 2 DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { double w = __shfl_up_sync(0xffffffffu, a, 1); double n0 = fma(y, w, x); a = fma(y, n0, x); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = (t1 - t0);
    // 3: shfl only chain (32-bit)
    int ia = (int)a;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) ia = __shfl_up_sync(0xffffffffu, ia, 1) + 1;
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = (t1 - t0);
    // 4: LDS dependent chain
    int idx = threadIdx.x & 31;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) idx = (int)sm[idx] & 31;
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = (t1 - t0);
    // 5: redux.sync add chain
    unsigned r = ia;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) r = __reduce_add_sync(0xffffffffu, r) + i;
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = (t1 - t0);
    // 6: 8 independent DFMA chains (throughput per warp)
    double b0 = a, b1 = a + 1, b2 = a + 2, b3 = a + 3, b4 = a + 4, b5 = a + 5, b6 = a + 6, b7 = a + 7;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; i++) { b0 = fma(b0, y, x); b1 = fma(b1, y, x); b2 = fma(b2, y, x); b3 = fma(b3, y, x); b4 = fma(b4, y, x); b5 = fma(b5, y, x); b6 = fma(b6, y, x); b7 = fma(b7, y, x); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = (t1 - t0);
    // 7: FFMA chain (fp32) for comparison
    float f = (float)a;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) f = fmaf(f, (float)y, (float)x);
    t1 = clock64(); if (threadIdx.x == 0) cyc[7] = (t1 - t0);
    // 8: F2I.S64 conversion chain
    long long li = 0; double c = a;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { li = __double2ll_rn(c); c = (double)(li & 1023) + y; }
    t1 = clock64(); if (threadIdx.x == 0) cyc[8] = (t1 - t0);
    out[threadIdx.x + blockIdx.x * blockDim.x] = a + ia + idx + r + b0 + b1 + b2 + b3 + b4 + b5 + b6 + b7 + f + c;
}
int main()
{
    double *out; long long *cyc, h[16];
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 128);
    const char *names[] = {"DFMA chain", "DADD chain", "SHFL.UP f64 + 2 DFMA", "SHFL.UP s32 + IADD", "LDS chain (+cvt)", "REDUX.SUM + IADD", "8x indep DFMA (per 8)", "FFMA chain", "F2I.S64+I2F+DADD chain"};
    for (int threads : {32, 128, 512}) {
        for (int rep = 0; rep < 2; rep++) { k<<<1, threads>>>(out, cyc, 1.0000001, 0.999999); cudaDeviceSynchronize(); }
        cudaMemcpy(h, cyc, 128, cudaMemcpyDeviceToHost);
        printf("threads/CTA %d (one CTA):\n", threads);
        for (int i = 0; i < 9; i++) printf("  %-26s %7.2f cycles/iter\n", names[i], h[i] / 512.0);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
