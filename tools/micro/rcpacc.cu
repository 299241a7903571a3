// Accuracy of the branch-free reciprocal / division variants against IEEE division (fp64).
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ double rel(double x, double ref) { return fabs((x - ref) / ref); }
__global__ void k(double *out, int n)
{
    double m0 = 0, m2 = 0, m3 = 0, mq2 = 0, mq3 = 0, mq0 = 0, mh = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        // b spans many binades and mantissas; a arbitrary
        unsigned long long h = (unsigned long long)i * 0x9E3779B97F4A7C15ull + 12345;
        double mant = 1.0 + (double)(h >> 11) * (1.0 / 9007199254740992.0);
        int ex = (int)((h >> 3) % 120) - 60;
        double b = ldexp(mant, ex) * ((h & 1) ? 1 : -1);
        double a = 1.0 + (double)((h * 7919) >> 11) * (1.0 / 9007199254740992.0) * 3.0;
        double y0; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(b));
        double r = 1.0 / b;
        m0 = fmax(m0, rel(y0, r));
        double e = fma(-b, y0, 1.0), y = fma(y0, e, y0); e = fma(-b, y, 1.0); y = fma(y, e, y);
        m2 = fmax(m2, rel(y, r)); mq2 = fmax(mq2, rel(a * y, a / b));
        e = fma(-b, y0, 1.0); double y3 = fma(y0, fma(e, e, e), y0);
        m3 = fmax(m3, rel(y3, r)); mq3 = fmax(mq3, rel(a * y3, a / b));
        // the shipped forms (csrc/common.cuh): fdiv = (a y0)(1 + e + e^2), fdiv_half = (a (y0/2))(1 + e + e^2)
        double q0 = a * y0, t = fma(e, e, e);
        mq0 = fmax(mq0, rel(fma(q0, t, q0), a / b));
        double hy = __hiloint2double(__double2hiint(y0) - 0x00100000, 0), hq0 = a * hy;
        mh = fmax(mh, rel(fma(hq0, t, hq0), 0.5 * (a / b)));
    }
    out[(blockIdx.x * blockDim.x + threadIdx.x) * 7 + 0] = m0;
    out[(blockIdx.x * blockDim.x + threadIdx.x) * 7 + 1] = m2;
    out[(blockIdx.x * blockDim.x + threadIdx.x) * 7 + 2] = m3;
    out[(blockIdx.x * blockDim.x + threadIdx.x) * 7 + 3] = mq2;
    out[(blockIdx.x * blockDim.x + threadIdx.x) * 7 + 4] = mq3;
    out[(blockIdx.x * blockDim.x + threadIdx.x) * 7 + 5] = mq0;
    out[(blockIdx.x * blockDim.x + threadIdx.x) * 7 + 6] = mh;
}
int main()
{
    const int G = 256, T = 256; double *d, *h = new double[G * T * 7];
    cudaMalloc(&d, G * T * 7 * 8);
    k<<<G, T>>>(d, 1 << 28); cudaMemcpy(h, d, G * T * 7 * 8, cudaMemcpyDeviceToHost);
    double m[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < G * T; i++) for (int j = 0; j < 7; j++) m[j] = fmax(m[j], h[i * 7 + j]);
    printf("rcp.approx seed max rel err      %.3e (2^%.1f)\n", m[0], log2(m[0]));
    printf("2 Newton steps: 1/b              %.3e  a/b %.3e (ulp = 1.1e-16)\n", m[1], m[3]);
    printf("1 cubic step:   1/b              %.3e  a/b %.3e\n", m[2], m[4]);
    printf("shipped fdiv (a y0)(1+e+e^2):    a/b %.3e   fdiv_half: 0.5 a/b %.3e\n", m[5], m[6]);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
