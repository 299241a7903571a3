// burgers-v0 — 1D inviscid Burgers with inlet noise and a point-forcing actuator.
//
// Reference: /root/reference/beacon/burgers/burgers.py — solve() :119-151, derx :231-243,
// rhs :253-255, dert :247-249, get_obs :154-159, get_rwd :162-166, step :98-116.
//
// One CTA per environment; each thread keeps C consecutive points of u, up, upp in registers
// for all ndt_act sub-steps of all fused actions (same scheme as shkadov.cu): per sub-step only
// the chunk edges (first u, last two u) cross shared memory, one __syncthreads per sub-step.
#include <cstdlib>

#include "common.cuh"

namespace beacon {

template <typename R> struct BurArgs {
    int nx, ndt_act, n_act, ctrl_pos, n_obs, off, B, mode, n_fused;
    R inv_dx, two_dt, amp, u_target, dx;
    double sigma;
    uint64_t seed;
    int64_t env_base;
    R *u, *up, *upp, *a_cur;
    int32_t *stp;
    unsigned long long *draws;
    const R *actions, *noise;
    const uint8_t *mask;
    R *obs, *rwd;
    uint8_t *done, *trunc;
    int32_t *status;
};

template <typename R, int C, int T>
__global__ void __launch_bounds__(T) burgers_kernel(const BurArgs<R> a)
{
    __shared__ R ex[2][3][T];
    __shared__ R s_u[C * T];
    __shared__ R s_red[T / 32];
    const int tid = threadIdx.x, b = blockIdx.x, nx = a.nx;
    const bool resetting = a.mode == 1;
    if (resetting && a.mask && !a.mask[b]) return;
    const int a0 = tid * C - a.off;
    const int tl = tid > 0 ? tid - 1 : 0, tr = tid < T - 1 ? tid + 1 : T - 1;
    const size_t row = (size_t)b * nx;

    R u[C], up[C], upp[C];
#pragma unroll
    for (int m = 0; m < C; m++) {
        int i = a0 + m;
        bool real = i >= 0 && i < nx;
        if (resetting || !real) { u[m] = up[m] = upp[m] = a.u_target; }      // reset_fields, :77-95
        else { u[m] = a.u[row + i]; up[m] = a.up[row + i]; upp[m] = a.upp[row + i]; }
    }
    int stp = resetting ? 0 : a.stp[b];
    unsigned long long draws = a.draws[b];
    R act_val = resetting ? R(0) : a.a_cur[b];
    const int n_actions = resetting ? 0 : a.n_fused;
    int status = 0;                                     // OR of the beacon_status bits of all fused actions

    for (int act = 0; act < n_actions; act++) {
        const size_t orow = (size_t)act * a.B + b;
        act_val = a.actions[orow];
        // one noise draw per action, burgers.py:127
        R nz = a.noise ? a.noise[orow] : (R)philox_uniform_pm(a.seed, (uint64_t)(a.env_base + b), draws, a.sigma);
        draws += 1ull;
        const R forcing = act_val * a.amp;

        // one sub-step; run unrolled by two (the BDF2 history rotation upp <- up <- u then costs no copies)
        auto substep = [&](const int it) {
            R(*X)[T] = ex[it & 1];
#pragma unroll
            for (int m = 0; m < C; m++) { upp[m] = up[m]; up[m] = u[m]; }                 // :135-136
#pragma unroll
            for (int m = 0; m < C; m++) {                                                // :139-140
                int i = a0 + m;
                if (i == 0) u[m] = a.u_target + nz;
                if (m > 0 && i == nx - 1) u[m] = u[m - 1];
            }
            X[0][tid] = u[0]; X[1][tid] = u[C - 2]; X[2][tid] = u[C - 1];
            __syncthreads();
            R e[C + 3];   // u at a0-2 .. a0+C
            e[0] = X[1][tl]; e[1] = X[2][tl];
#pragma unroll
            for (int m = 0; m < C; m++) e[m + 2] = u[m];
            e[C + 2] = X[0][tr];
            R d[C + 2];
#pragma unroll
            for (int k = 0; k < C + 2; k++) d[k] = e[k + 1] - e[k];
            R F[C + 1];   // faces a0-1 .. a0+C-1 ; derx :231-243
#pragma unroll
            for (int m = 0; m < C + 1; m++) {
                int f = a0 - 1 + m;
                // van Leer, burgers.py:231-243: r = a / b, phi = (r + |r|) / (1 + r) with a = d[m], b = d[m+1] + 1e-8.
                // In ONE division: a, b of equal sign -> phi / 2 = a / (a + b); opposite sign or a = 0 -> 0; a = -b
                // (r = -1 exactly) -> NaN like the reference.  Halves the dependent division chain of the sub-step.
                const R av = d[m], bv = d[m + 1] + R(1.0e-8), sum = av + bv;
                R hphi = fdiv(av, sum);
                if (!same_sign(av, bv)) hphi = is_zero(sum) ? qnan<R>() : R(0);    // sign / zero tests on the integer pipe
                if (f <= 0) hphi = R(0);
                F[m] = e[m + 1] + hphi * d[m + 1];
            }
#pragma unroll
            for (int m = 0; m < C; m++) {
                int i = a0 + m;
                R du = (F[m + 1] - F[m]) * a.inv_dx;
                R rhs = u[m] * du;                                                       // rhs(), :253-255
                if (i == a.ctrl_pos) rhs += forcing;                                     // :149
                if (i >= 1 && i <= nx - 2) u[m] = div3(R(4) * up[m] - upp[m] - a.two_dt * rhs);        // dert(), :247-249
            }
        };
        {
            int it = 0;
            for (; it + 1 < a.ndt_act; it += 2) { substep(it); substep(it + 1); }
            for (; it < a.ndt_act; it++) substep(it);
        }
        // ---- obs / reward ------------------------------------------------------------------
        __syncthreads();
        R part = R(0);
        bool nonfinite = false;
#pragma unroll
        for (int m = 0; m < C; m++) {
            int i = a0 + m;
            if (i >= 0 && i < nx) {
                s_u[i] = u[m];
                if (i >= a.ctrl_pos) part += rabs(u[m] - a.u_target);                   // :164
                nonfinite |= !finite_(u[m]);
            }
        }
        R tot = block_sum(part, s_red);
        int bad = __syncthreads_or(nonfinite ? 1 : 0);
        for (int k = tid; k < a.n_obs; k += T) a.obs[orow * a.n_obs + k] = s_u[a.ctrl_pos - a.n_obs + k];   // :154-159
        if (tid == 0) {
            a.rwd[orow] = -tot * a.dx;
            bool horizon = stp == a.n_act - 1;
            a.done[orow] = horizon; a.trunc[orow] = horizon;
            if (bad) status |= BEACON_STATUS_NONFINITE;
        }
        stp += 1;
        __syncthreads();
    }

    if (resetting) {
        for (int k = tid; k < a.n_obs; k += T) a.obs[(size_t)b * a.n_obs + k] = a.u_target;
    }
#pragma unroll
    for (int m = 0; m < C; m++) {
        int i = a0 + m;
        if (i >= 0 && i < nx) { a.u[row + i] = u[m]; a.up[row + i] = up[m]; a.upp[row + i] = upp[m]; }
    }
    if (tid == 0) {
        a.stp[b] = stp; a.draws[b] = draws; a.a_cur[b] = act_val;
        if (!resetting && a.status) a.status[b] = status;
    }
}

template <typename R> class BurgersEnv : public Env {
    beacon_burgers_params p;
    DeviceBuffer u, up, upp, a_cur, stp, draws;
    BurArgs<R> base{};
    // two variants: 4 points per thread (throughput, large batches) and 2 points per thread with
    // twice the warps (latency of a single env: configs[0] steps ONE env, the per-sub-step
    // dependent chain is what counts)
    int C = 4, T = 128;
    void (*kernel)(const BurArgs<R>) = nullptr;

public:
    BurgersEnv(const beacon_common &c, const beacon_burgers_params &pp) : p(pp)
    {
        common = c;
        const int B = c.batch, nx = p.nx;
        BEACON_REQUIRE(nx >= 8 && p.ndt_act > 0, "burgers: bad sizes");
        if (nx > 511) throw Error(BEACON_ERR_UNSUPPORTED, "burgers: nx > 511 not supported");
        if (B <= 296 && !getenv("BEACON_BURGERS_C4")) { C = 2; T = 256; kernel = burgers_kernel<R, 2, 256>; }
        else { C = 4; T = 128; kernel = burgers_kernel<R, 4, 128>; }
        BEACON_REQUIRE(p.ctrl_pos - p.n_obs_pts >= 0 && p.ctrl_pos >= 1 && p.ctrl_pos <= nx - 2, "burgers: control point outside the domain");
        info.kind = BEACON_BURGERS; info.batch = B; info.dtype = real_traits<R>::dtype; info.device = c.device;
        info.n_obs = p.n_obs_pts; info.act_dim = 1; info.act_is_int = 0; info.rwd_dim = 1; info.n_act = p.n_act; info.noise_dim = 1;
        size_t nb = (size_t)B * nx * sizeof(R);
        u.alloc(nb); up.alloc(nb); upp.alloc(nb); a_cur.alloc(B * sizeof(R)); stp.alloc(B * 4); draws.alloc(B * 8);
        add_field("u", u.ptr, nx); add_field("up", up.ptr, nx); add_field("upp", upp.ptr, nx);
        add_field("a", a_cur.ptr, 1); add_field("stp", stp.ptr, 1, true);
        add_field("draws", draws.ptr, 2, true);       // uint64 Philox draw counter as two int32 words (checkpoint / resume)
        BurArgs<R> &a = base;
        a.nx = nx; a.ndt_act = p.ndt_act; a.n_act = p.n_act; a.ctrl_pos = p.ctrl_pos; a.n_obs = p.n_obs_pts;
        a.off = ((nx - 1) % C == 0) ? 1 : 0; a.B = B;
        a.inv_dx = (R)(1.0 / p.dx); a.two_dt = (R)(2.0 * p.dt); a.amp = (R)p.amp; a.u_target = (R)p.u_target; a.dx = (R)p.dx;
        a.sigma = p.sigma; a.seed = c.seed; a.env_base = c.env_index_base;
        a.u = u.as<R>(); a.up = up.as<R>(); a.upp = upp.as<R>(); a.a_cur = a_cur.as<R>();
        a.stp = stp.as<int32_t>(); a.draws = draws.as<unsigned long long>();
    }
    void reset(const ResetArgs &r) override
    {
        BEACON_REQUIRE(r.obs != nullptr, "reset: obs must not be NULL");
        BurArgs<R> a = base;
        a.mode = 1; a.mask = r.mask; a.obs = (R *)r.obs;
        kernel<<<a.B, T, 0, r.stream>>>(a);
        BEACON_CUDA_CHECK(cudaGetLastError());
        launches++;
    }
    void step(const StepArgs &s) override
    {
        BurArgs<R> a = base;
        a.mode = 0; a.n_fused = s.n_fused; a.actions = (const R *)s.actions; a.noise = (const R *)s.noise;
        a.obs = (R *)s.obs; a.rwd = (R *)s.rwd; a.done = s.done; a.trunc = s.trunc; a.status = s.status;
        kernel<<<a.B, T, 0, s.stream>>>(a);
        BEACON_CUDA_CHECK(cudaGetLastError());
        launches++;
    }
};

Env *make_burgers(const beacon_common &c, const beacon_burgers_params &p)
{
    if (c.dtype == BEACON_F64) return new BurgersEnv<double>(c, p);
    if (c.dtype == BEACON_F32) return new BurgersEnv<float>(c, p);
    throw Error(BEACON_ERR_INVALID, "unknown dtype");
}

}  // namespace beacon
