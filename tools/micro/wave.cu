// Cycles per step of the three-plane transport wavefront loop in isolation (one warp).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int wavefront_row_stride(int ny) { return ((ny + 1 - 2 + 7) / 8) * 8 + 2; }
__device__ __forceinline__ double2 lds128(const void *p)
{
    double2 v;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
}
template <typename R, int NY, int VAR>
__device__ __noinline__ void wf(uint32_t offA, uint32_t offW, uint32_t offS, int lanes, int lane)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int RS = wavefront_row_stride(NY);
    const bool on = lane < lanes;
    const int l = on ? lane : 0;
    double2 *AA = reinterpret_cast<double2 *>(smem_raw + offA) + l * RS;
    const double2 *WW = reinterpret_cast<const double2 *>(smem_raw + offW) + l * RS, *SS = reinterpret_cast<const double2 *>(smem_raw + offS) + l * RS;
    R p0 = AA[1].x, p1 = AA[1].y, w0 = WW[1].x, w1 = WW[1].y;
    const double2 *An = AA + (2 - l), *Wn = WW + (2 - l), *Sn = SS + (2 - l);
    double2 *Out = AA + (1 - l);
    double2 an = An[0], wn = Wn[0], sn = Sn[0];
    R last_new = R(0);
    const int c0 = on ? -lane : -(1 << 20);
    const int steps = ((NY + lanes - 1 + 5) / 6) * 6;
#pragma unroll 6
    for (int t = 0; t < steps; t++) {
        double2 an2, wn2, sn2;
        if (VAR == 3) { an2 = lds128(An + t + 1); wn2 = lds128(Wn + t + 1); sn2 = lds128(Sn + t + 1); }
        else { an2 = An[t + 1]; wn2 = Wn[t + 1]; sn2 = Sn[t + 1]; }
        R wv;
        if (VAR == 0) wv = __shfl_up_sync(0xffffffffu, last_new, 1);
        else wv = last_new * 0.999;                      // no shuffle: isolates the FMA / load chain
        const R n0 = fma(w0, wv, p0);
        const R n1 = fma(w1, n0, p1);
        last_new = n1;
        if (VAR != 2) { if ((unsigned)(c0 + t) < (unsigned)NY) { double2 o; o.x = n0; o.y = n1; Out[t] = o; } }
        p0 = fma(sn.x, n0, an.x); p1 = fma(sn.y, n1, an.y);
        w0 = wn.x; w1 = wn.y;
        an = an2; wn = wn2; sn = sn2;
    }
    if (last_new == 12345.678) Out[0].x = last_new;
}
template <int VAR> __global__ void k(long long *cyc, int lanes)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s = reinterpret_cast<double *>(smem_raw);
    for (int i = threadIdx.x; i < 3 * 5600; i += blockDim.x) s[i] = 0.3 + 1e-3 * (i % 97);
    __syncthreads();
    long long t0 = clock64();
    if (threadIdx.x < 32) wf<double, 100, VAR>(0, 44800, 89600, lanes, threadIdx.x);
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[VAR] = t1 - t0;
}
int main()
{
    long long *d, h[4];
    cudaMalloc(&d, 64);
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140000);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140000);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140000);
    cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140000);
    for (int threads : {32, 512}) {
        for (int rep = 0; rep < 2; rep++) { k<0><<<1, threads, 140000>>>(d, 26); k<1><<<1, threads, 140000>>>(d, 26); k<2><<<1, threads, 140000>>>(d, 26); k<3><<<1, threads, 140000>>>(d, 26); cudaDeviceSynchronize(); }
        cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        printf("threads %3d: full loop %.1f cycles/step, without shuffle %.1f, without store %.1f, asm loads %.1f  (126 steps)  %s\n", threads, h[0] / 126.0, h[1] / 126.0, h[2] / 126.0, h[3] / 126.0,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
