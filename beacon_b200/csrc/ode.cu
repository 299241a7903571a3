// lorenz-v0 and vortex-v0 — low-dimensional ODE environments, five-stage fourth-order
// low-storage Runge-Kutta (Carpenter-Kennedy).
//
// Reference: /root/reference/beacon/lorenz/lorenz.py — solve() :120-153, lsrk4 :267-297,
// get_obs :156-164, get_rwd :167-172;  /root/reference/beacon/vortex/vortex.py — solve()
// :149-183, get_obs :186-194, get_rwd :197-208.
//
// One THREAD per environment (the state is 3-4 numbers); any number of fused actions per
// launch.  As in the reference the register `x` doubles as the low-storage residual and `xk`
// as the solution inside a step.
#include "common.cuh"

namespace beacon {

__device__ __constant__ double LSRK_A[5] = {0.000000000000000, -0.417890474499852, -1.192151694642677,
                                            -1.697784692471528, -1.514183444257156};
__device__ __constant__ double LSRK_B[5] = {0.149659021999229, 0.379210312999627, 0.822955029386982,
                                            0.699450455949122, 0.153057247968152};

template <typename R> struct LorArgs {
    int B, ndt_act, n_act, mode, n_fused;
    R dt, sigma, rho, beta, x0[3], forcing[3];
    R *x, *fx;
    int32_t *stp;
    const int32_t *actions;
    const uint8_t *mask;
    R *obs, *rwd;
    uint8_t *done, *trunc;
    int32_t *status;
};

template <typename R> __global__ void __launch_bounds__(128) lorenz_kernel(const LorArgs<R> a)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    if (a.mode == 1) {                                         // reset_fields, lorenz.py:69-95
        if (a.mask && !a.mask[b]) return;
        for (int i = 0; i < 3; i++) { a.x[3 * b + i] = a.x0[i]; a.fx[3 * b + i] = R(0); a.obs[6 * b + i] = a.x0[i]; a.obs[6 * b + 3 + i] = R(0); }
        a.stp[b] = 0;
        return;
    }
    R x[3], xk[3], f[3];
    for (int i = 0; i < 3; i++) { x[i] = a.x[3 * b + i]; f[i] = a.fx[3 * b + i]; }
    int stp = a.stp[b];
    int st = 0;
    for (int act = 0; act < a.n_fused; act++) {
        const size_t orow = (size_t)act * a.B + b;
        int u = a.actions[orow];
        u = u < 0 ? 0 : (u > 2 ? 2 : u);
        const R force = a.forcing[u];
        for (int it = 0; it < a.ndt_act; it++) {
            for (int i = 0; i < 3; i++) xk[i] = x[i];
#pragma unroll
            for (int j = 0; j < 5; j++) {
                f[0] = a.sigma * (xk[1] - xk[0]);                                       // :137-139
                f[1] = xk[0] * (a.rho - xk[2]) - xk[1];
                f[2] = xk[0] * xk[1] - a.beta * xk[2];
                f[1] += force;                                                          // :142
                const R A = (R)LSRK_A[j], Bc = (R)LSRK_B[j];
#pragma unroll
                for (int i = 0; i < 3; i++) { x[i] = A * x[i] + a.dt * f[i]; xk[i] += Bc * x[i]; }   // :293-297
            }
            for (int i = 0; i < 3; i++) x[i] = xk[i];
        }
        for (int i = 0; i < 3; i++) { a.obs[orow * 6 + i] = x[i]; a.obs[orow * 6 + 3 + i] = f[i]; }
        a.rwd[orow] = x[0] < R(0) ? R(1) : R(0);
        bool horizon = stp == a.n_act - 1;
        a.done[orow] = horizon; a.trunc[orow] = horizon;
        stp += 1;
        if (!finite_(x[0]) || !finite_(x[1]) || !finite_(x[2])) st |= BEACON_STATUS_NONFINITE;
    }
    for (int i = 0; i < 3; i++) { a.x[3 * b + i] = x[i]; a.fx[3 * b + i] = f[i]; }
    a.stp[b] = stp;
    if (a.status) a.status[b] = st;
}

template <typename R> class LorenzEnv : public Env {
    DeviceBuffer x, fx, stp;
    LorArgs<R> base{};

public:
    LorenzEnv(const beacon_common &c, const beacon_lorenz_params &p)
    {
        common = c;
        const int B = c.batch;
        BEACON_REQUIRE(p.ndt_act > 0, "lorenz: bad sizes");
        host_zero_copy = false;
        info.kind = BEACON_LORENZ; info.batch = B; info.dtype = real_traits<R>::dtype; info.device = c.device;
        info.n_obs = 6; info.act_dim = 1; info.act_is_int = 1; info.rwd_dim = 1; info.n_act = p.n_act; info.noise_dim = 0;
        x.alloc((size_t)B * 3 * sizeof(R)); fx.alloc((size_t)B * 3 * sizeof(R)); stp.alloc((size_t)B * 4);
        add_field("x", x.ptr, 3); add_field("fx", fx.ptr, 3); add_field("stp", stp.ptr, 1, true);
        LorArgs<R> &a = base;
        a.B = B; a.ndt_act = p.ndt_act; a.n_act = p.n_act; a.dt = (R)p.dt; a.sigma = (R)p.sigma; a.rho = (R)p.rho; a.beta = (R)p.beta;
        for (int i = 0; i < 3; i++) { a.x0[i] = (R)p.x0[i]; a.forcing[i] = (R)p.forcing[i]; }
        a.x = x.as<R>(); a.fx = fx.as<R>(); a.stp = stp.as<int32_t>();
    }
    void run(const LorArgs<R> &a, cudaStream_t s)
    {
        lorenz_kernel<R><<<(a.B + 127) / 128, 128, 0, s>>>(a);
        BEACON_CUDA_CHECK(cudaGetLastError());
        launches++;
    }
    void reset(const ResetArgs &r) override
    {
        BEACON_REQUIRE(r.obs != nullptr, "reset: obs must not be NULL");
        LorArgs<R> a = base; a.mode = 1; a.mask = r.mask; a.obs = (R *)r.obs;
        run(a, r.stream);
    }
    void step(const StepArgs &s) override
    {
        LorArgs<R> a = base; a.mode = 0; a.n_fused = s.n_fused; a.actions = (const int32_t *)s.actions;
        a.obs = (R *)s.obs; a.rwd = (R *)s.rwd; a.done = s.done; a.trunc = s.trunc; a.status = s.status;
        run(a, s.stream);
    }
};

Env *make_lorenz(const beacon_common &c, const beacon_lorenz_params &p)
{
    if (c.dtype == BEACON_F64) return new LorenzEnv<double>(c, p);
    if (c.dtype == BEACON_F32) return new LorenzEnv<float>(c, p);
    throw Error(BEACON_ERR_INVALID, "unknown dtype");
}

// ---------------------------------------------------------------------------------------
// vortex
// ---------------------------------------------------------------------------------------
template <typename R> struct VorArgs {
    int B, ndt_act, n_act, mode, n_fused;
    R dt, lmbda_re, lmbda_cx, mu_re, mu_cx, alpha_re, alpha_cx, ire, omega_s, omega_f, domega, gamma, beta_m, weight,
        mod_min, mod_max, phase_min, phase_max, x0[4];
    R *x, *fx, *t, *y;
    int32_t *stp;
    const R *actions;
    const uint8_t *mask;
    R *obs, *rwd;
    uint8_t *done, *trunc;
    int32_t *status;
};

template <typename R> __global__ void __launch_bounds__(128) vortex_kernel(const VorArgs<R> a)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    if (a.mode == 1) {                                         // reset_fields, vortex.py:88-122
        if (a.mask && !a.mask[b]) return;
        for (int i = 0; i < 4; i++) { a.x[4 * b + i] = a.x0[i]; a.fx[4 * b + i] = R(0); a.obs[8 * b + i] = a.x0[i]; a.obs[8 * b + 4 + i] = R(0); }
        a.t[b] = R(0);
        a.y[b] = R(2) * (a.x0[2] * R(1) - a.x0[3] * R(0));     // cos(0)=1, sin(0)=0  (:98-99)
        a.stp[b] = 0;
        return;
    }
    R x[4], xk[4], f[4];
    for (int i = 0; i < 4; i++) { x[i] = a.x[4 * b + i]; f[i] = a.fx[4 * b + i]; }
    R t = a.t[b], y = a.y[b];
    int stp = a.stp[b], st = 0;
    for (int act = 0; act < a.n_fused; act++) {
        const size_t orow = (size_t)act * a.B + b;
        const R u0 = a.actions[orow * 2], u1 = a.actions[orow * 2 + 1];
        const R kmod = a.mod_min + R(0.5) * (u0 + R(1)) * (a.mod_max - a.mod_min);           // :156-157
        const R kphase = a.phase_min + R(0.5) * (u1 + R(1)) * (a.phase_max - a.phase_min);
        const R ck = (R)cos((double)kphase), sk = (R)sin((double)kphase);
        for (int it = 0; it < a.ndt_act; it++) {
            for (int i = 0; i < 4; i++) xk[i] = x[i];
#pragma unroll
            for (int j = 0; j < 5; j++) {
                const R n2 = xk[0] * xk[0] + xk[1] * xk[1];                                 // :169-172
                f[0] = a.ire * (a.lmbda_re * xk[0] - a.lmbda_cx * xk[1]) - (a.mu_re * xk[0] - a.mu_cx * xk[1]) * n2 +
                       (a.alpha_re * xk[2] - a.alpha_cx * xk[3]) + xk[0] * kmod * ck - xk[1] * kmod * sk;
                f[1] = a.ire * (a.lmbda_re * xk[1] + a.lmbda_cx * xk[0]) - (a.mu_re * xk[1] + a.mu_cx * xk[0]) * n2 +
                       (a.alpha_re * xk[3] + a.alpha_cx * xk[2]) + xk[0] * kmod * sk + xk[1] * kmod * ck;
                f[2] = -a.omega_f * a.gamma * xk[2] - a.domega * xk[3] + a.beta_m * xk[0];
                f[3] = -a.omega_f * a.gamma * xk[3] + a.domega * xk[2] + a.beta_m * xk[1];
                const R A = (R)LSRK_A[j], Bc = (R)LSRK_B[j];
#pragma unroll
                for (int i = 0; i < 4; i++) { x[i] = A * x[i] + a.dt * f[i]; xk[i] += Bc * x[i]; }
            }
            for (int i = 0; i < 4; i++) x[i] = xk[i];
            t += a.dt;                                                                       // :180
        }
        for (int i = 0; i < 4; i++) { a.obs[orow * 8 + i] = x[i]; a.obs[orow * 8 + 4 + i] = f[i]; }
        // get_rwd(), :197-208
        const R yp = y;
        const R cw = (R)cos((double)(a.omega_f * t)), sw = (R)sin((double)(a.omega_f * t));
        y = R(2) * (x[2] * cw - x[3] * sw);
        R cost = R(2) * kmod * ck * (x[0] * cw - x[1] * sw) - R(2) * kmod * sk * (x[1] * cw + x[0] * sw);
        cost = R(0.5) * cost * cost;
        const R dy = (y - yp) / a.dt;
        a.rwd[orow] = R(2) * a.omega_s * a.gamma * (dy * dy) - a.weight * cost;
        bool horizon = stp == a.n_act - 1;
        a.done[orow] = horizon; a.trunc[orow] = horizon;
        stp += 1;
        if (!finite_(x[0]) || !finite_(x[1]) || !finite_(x[2]) || !finite_(x[3])) st |= BEACON_STATUS_NONFINITE;
    }
    for (int i = 0; i < 4; i++) { a.x[4 * b + i] = x[i]; a.fx[4 * b + i] = f[i]; }
    a.t[b] = t; a.y[b] = y; a.stp[b] = stp;
    if (a.status) a.status[b] = st;
}

template <typename R> class VortexEnv : public Env {
    DeviceBuffer x, fx, t, y, stp;
    VorArgs<R> base{};

public:
    VortexEnv(const beacon_common &c, const beacon_vortex_params &p)
    {
        common = c;
        const int B = c.batch;
        BEACON_REQUIRE(p.ndt_act > 0, "vortex: bad sizes");
        host_zero_copy = false;
        info.kind = BEACON_VORTEX; info.batch = B; info.dtype = real_traits<R>::dtype; info.device = c.device;
        info.n_obs = 8; info.act_dim = 2; info.act_is_int = 0; info.rwd_dim = 1; info.n_act = p.n_act; info.noise_dim = 0;
        x.alloc((size_t)B * 4 * sizeof(R)); fx.alloc((size_t)B * 4 * sizeof(R)); t.alloc((size_t)B * sizeof(R));
        y.alloc((size_t)B * sizeof(R)); stp.alloc((size_t)B * 4);
        add_field("x", x.ptr, 4); add_field("fx", fx.ptr, 4); add_field("t", t.ptr, 1); add_field("y", y.ptr, 1);
        add_field("stp", stp.ptr, 1, true);
        VorArgs<R> &a = base;
        a.B = B; a.ndt_act = p.ndt_act; a.n_act = p.n_act; a.dt = (R)p.dt;
        a.lmbda_re = (R)p.lmbda_re; a.lmbda_cx = (R)p.lmbda_cx; a.mu_re = (R)p.mu_re; a.mu_cx = (R)p.mu_cx;
        a.alpha_re = (R)p.alpha_re; a.alpha_cx = (R)p.alpha_cx; a.ire = (R)p.ire; a.omega_s = (R)p.omega_s; a.omega_f = (R)p.omega_f;
        a.domega = (R)p.domega; a.gamma = (R)p.gamma; a.beta_m = (R)p.beta_m; a.weight = (R)p.weight;
        a.mod_min = (R)p.mod_min; a.mod_max = (R)p.mod_max; a.phase_min = (R)p.phase_min; a.phase_max = (R)p.phase_max;
        for (int i = 0; i < 4; i++) a.x0[i] = (R)p.x0[i];
        a.x = x.as<R>(); a.fx = fx.as<R>(); a.t = t.as<R>(); a.y = y.as<R>(); a.stp = stp.as<int32_t>();
    }
    void run(const VorArgs<R> &a, cudaStream_t s)
    {
        vortex_kernel<R><<<(a.B + 127) / 128, 128, 0, s>>>(a);
        BEACON_CUDA_CHECK(cudaGetLastError());
        launches++;
    }
    void reset(const ResetArgs &r) override
    {
        BEACON_REQUIRE(r.obs != nullptr, "reset: obs must not be NULL");
        VorArgs<R> a = base; a.mode = 1; a.mask = r.mask; a.obs = (R *)r.obs;
        run(a, r.stream);
    }
    void step(const StepArgs &s) override
    {
        VorArgs<R> a = base; a.mode = 0; a.n_fused = s.n_fused; a.actions = (const R *)s.actions;
        a.obs = (R *)s.obs; a.rwd = (R *)s.rwd; a.done = s.done; a.trunc = s.trunc; a.status = s.status;
        run(a, s.stream);
    }
};

Env *make_vortex(const beacon_common &c, const beacon_vortex_params &p)
{
    if (c.dtype == BEACON_F64) return new VortexEnv<double>(c, p);
    if (c.dtype == BEACON_F32) return new VortexEnv<float>(c, p);
    throw Error(BEACON_ERR_INVALID, "unknown dtype");
}

}  // namespace beacon
