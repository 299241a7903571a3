// sloshing-v0 — 1D shallow-water (Saint-Venant) tank, Rusanov fluxes + Adams-Bashforth 2.
//
// Reference: /root/reference/beacon/sloshing/sloshing.py — solve() :168-224, rusanov :323-325,
// adams :329-331, get_obs :227-232, get_rwd :235-244, step :141-165, reset :89-125.
//
// One CTA per environment, C consecutive cells (ghosts included: nx+2 entries) per thread in
// registers across all sub-steps; per sub-step the chunk edges of (h, q, q^2/h+g h^2/2,
// |v|+sqrt(g h)) cross a double-buffered shared-memory exchange, one __syncthreads per sub-step.
#include "common.cuh"

namespace beacon {

template <typename R> struct SloArgs {
    int nx, n2, ndt_act, n_act, n_interp, obs_smpl, n_obs, off, B, mode, n_fused;
    R inv_dx, hdt, g, half_g, amp, alpha_pen, dx, blow_lo, blow_hi;
    R *h, *q, *rhsh, *rhsq, *u_cur, *u_prev;
    int32_t *stp;
    const R *h_init, *q_init;
    const R *actions;
    const uint8_t *mask;
    R *obs, *rwd;
    uint8_t *done, *trunc;
    int32_t *status;
};

template <typename R, int C, int T>
__global__ void __launch_bounds__(T) sloshing_kernel(const SloArgs<R> a)
{
    __shared__ R ex[2][8][T];
    __shared__ R s_q[C * T];
    __shared__ R s_red[T / 32];
    __shared__ R s_alpha[256];                              // min(i / n_interp, 1): one IEEE division per entry, once per launch
    const int tid = threadIdx.x, b = blockIdx.x, nx = a.nx, n2 = a.n2;
    const bool resetting = a.mode == 1;
    if (resetting && a.mask && !a.mask[b]) return;
    for (int i = tid; i < a.ndt_act && i < 256; i += T) s_alpha[i] = (R)fmin((double)i / (double)a.n_interp, 1.0);
    __syncthreads();
    const int a0 = tid * C - a.off;
    const int tl = tid > 0 ? tid - 1 : 0, tr = tid < T - 1 ? tid + 1 : T - 1;
    const size_t row = (size_t)b * n2;

    R h[C], q[C], rh[C], rq[C];
#pragma unroll
    for (int m = 0; m < C; m++) {
        int i = a0 + m;
        bool real = i >= 0 && i < n2;
        if (resetting) {                                   // reset, sloshing.py:89-125
            h[m] = real ? a.h_init[i] : R(1); q[m] = real ? a.q_init[i] : R(0); rh[m] = R(0); rq[m] = R(0);
        } else {
            h[m] = real ? a.h[row + i] : R(1); q[m] = real ? a.q[row + i] : R(0);
            rh[m] = real ? a.rhsh[row + i] : R(0); rq[m] = real ? a.rhsq[row + i] : R(0);
        }
    }
    int stp = resetting ? 0 : a.stp[b];
    R ucur = resetting ? R(0) : a.u_cur[b], uprev = resetting ? R(0) : a.u_prev[b];
    const int n_actions = resetting ? 0 : a.n_fused;
    int status = 0;                                     // OR of the beacon_status bits of all fused actions

    for (int act = 0; act < n_actions; act++) {
        const size_t orow = (size_t)act * a.B + b;
        uprev = ucur;                                      // :173-174
        ucur = a.actions[orow];
        // one sub-step; the loop below runs it unrolled by two so that the rotation of the state
        // registers (old values are still needed while the new ones are formed) costs no copies
        auto substep = [&](const int it) {
            R(*X)[T] = ex[it & 1];
            // wall boundary conditions, :183-186 (ghost and its neighbour share a chunk)
#pragma unroll
            for (int m = 0; m < C; m++) {
                int i = a0 + m;
                if (m < C - 1 && i == 0) { h[m] = h[m + 1]; q[m] = R(0); }
                if (m > 0 && i == nx + 1) { h[m] = h[m - 1]; q[m] = R(0); }
            }
            R w[C], s[C];   // w = q^2/h + g h^2/2 (:194), s = |v| + sqrt(g h) (:197-199)
#pragma unroll
            for (int m = 0; m < C; m++) {
                R v = fdiv(q[m], h[m]);
                w[m] = fdiv(q[m] * q[m], h[m]) + a.half_g * (h[m] * h[m]);
                s[m] = rabs(v) + fsqrt(a.g * h[m]);
            }
            X[0][tid] = h[0]; X[1][tid] = q[0]; X[2][tid] = w[0]; X[3][tid] = s[0];
            X[4][tid] = h[C - 1]; X[5][tid] = q[C - 1]; X[6][tid] = w[C - 1]; X[7][tid] = s[C - 1];
            __syncthreads();
            R eh[C + 2], eq[C + 2], ew[C + 2], es[C + 2];   // cells a0-1 .. a0+C
            eh[0] = X[4][tl]; eq[0] = X[5][tl]; ew[0] = X[6][tl]; es[0] = X[7][tl];
#pragma unroll
            for (int m = 0; m < C; m++) { eh[m + 1] = h[m]; eq[m + 1] = q[m]; ew[m + 1] = w[m]; es[m + 1] = s[m]; }
            eh[C + 1] = X[0][tr]; eq[C + 1] = X[1][tr]; ew[C + 1] = X[2][tr]; es[C + 1] = X[3][tr];
            R fh[C + 1], fq[C + 1];   // Rusanov fluxes at faces (a0-1|a0) .. (a0+C-1|a0+C); rusanov(), :323-325
#pragma unroll
            for (int m = 0; m < C + 1; m++) {
                R c = np_max(es[m], es[m + 1]);
                fh[m] = R(0.5) * (eq[m] + eq[m + 1]) - (R(0.5) * c) * (eh[m + 1] - eh[m]);
                fq[m] = R(0.5) * (ew[m] + ew[m + 1]) - (R(0.5) * c) * (eq[m + 1] - eq[m]);
            }
            R alpha = it < 256 ? s_alpha[it] : (R)fmin((double)it / (double)a.n_interp, 1.0);   // :218-220
            R f = ((R(1) - alpha) * uprev + alpha * ucur) * a.amp;
#pragma unroll
            for (int m = 0; m < C; m++) {
                int i = a0 + m;
                if (i >= 1 && i <= nx) {
                    R nrh = (fh[m + 1] - fh[m]) * a.inv_dx;                                // :214-215
                    R nrq = (fq[m + 1] - fq[m]) * a.inv_dx + f;
                    h[m] += a.hdt * (R(-3) * nrh + rh[m]);                                 // adams(), :329-331
                    q[m] += a.hdt * (R(-3) * nrq + rq[m]);
                    rh[m] = nrh; rq[m] = nrq;
                }
            }
        };
        {
            int it = 0;
            for (; it + 1 < a.ndt_act; it += 2) { substep(it); substep(it + 1); }
            for (; it < a.ndt_act; it++) substep(it);
        }
        // ---- guards / obs / reward, step() :141-165 --------------------------------------------
        __syncthreads();
        R part = R(0);
        int fl = 0;
#pragma unroll
        for (int m = 0; m < C; m++) {
            int i = a0 + m;
            if (i >= 0 && i < n2) {
                s_q[i] = q[m];
                if (i >= 1 && i <= nx) { R d = h[m] - R(1); part += d * d; }
                if (h[m] < a.blow_lo || h[m] > a.blow_hi) fl |= 1;
                if (!finite_(h[m]) || !finite_(q[m])) fl |= 4;
            }
        }
        R tot = block_sum(part, s_red);
        fl = __syncthreads_or(fl);
        for (int k = tid; k < a.n_obs; k += T) a.obs[orow * a.n_obs + k] = s_q[1 + k * a.obs_smpl];   // :227-232
        if (tid == 0) {
            a.rwd[orow] = R(0) - rsqrt_(tot) * a.dx - a.alpha_pen * rabs(a.amp * ucur);                // :235-244
            bool horizon = stp == a.n_act - 1, blow = fl & 1;
            a.done[orow] = horizon || blow; a.trunc[orow] = horizon && !blow;
            status |= (blow ? BEACON_STATUS_BLOWUP : 0) | ((fl & 4) ? BEACON_STATUS_NONFINITE : 0);
        }
        stp += 1;
        __syncthreads();
    }

    if (resetting) {
#pragma unroll
        for (int m = 0; m < C; m++) { int i = a0 + m; if (i >= 0 && i < n2) s_q[i] = q[m]; }
        __syncthreads();
        for (int k = tid; k < a.n_obs; k += T) a.obs[(size_t)b * a.n_obs + k] = s_q[1 + k * a.obs_smpl];
    }
#pragma unroll
    for (int m = 0; m < C; m++) {
        int i = a0 + m;
        if (i >= 0 && i < n2) { a.h[row + i] = h[m]; a.q[row + i] = q[m]; a.rhsh[row + i] = rh[m]; a.rhsq[row + i] = rq[m]; }
    }
    if (tid == 0) {
        a.stp[b] = stp; a.u_cur[b] = ucur; a.u_prev[b] = uprev;
        if (!resetting && a.status) a.status[b] = status;
    }
}

template <typename R> class SloshingEnv : public Env {
    beacon_sloshing_params p;
    DeviceBuffer h, q, rhsh, rhsq, u_cur, u_prev, stp, h_init, q_init;
    SloArgs<R> base{};
    static constexpr int C = 4, T = 64;

public:
    SloshingEnv(const beacon_common &c, const beacon_sloshing_params &pp, const double *h0, const double *q0) : p(pp)
    {
        common = c;
        const int B = c.batch, nx = p.nx, n2 = nx + 2;
        BEACON_REQUIRE(nx >= 4 && p.ndt_act > 0 && p.n_interp > 0 && p.obs_smpl > 0, "sloshing: bad sizes");
        if (n2 > C * T - 1) throw Error(BEACON_ERR_UNSUPPORTED, "sloshing: nx > 253 not supported");
        BEACON_REQUIRE(1 + (p.n_obs - 1) * p.obs_smpl <= nx, "sloshing: observation probes outside the domain");
        info.kind = BEACON_SLOSHING; info.batch = B; info.dtype = real_traits<R>::dtype; info.device = c.device;
        info.n_obs = p.n_obs; info.act_dim = 1; info.act_is_int = 0; info.rwd_dim = 1; info.n_act = p.n_act; info.noise_dim = 0;
        size_t nb = (size_t)B * n2 * sizeof(R);
        h.alloc(nb); q.alloc(nb); rhsh.alloc(nb); rhsq.alloc(nb); u_cur.alloc(B * sizeof(R)); u_prev.alloc(B * sizeof(R)); stp.alloc(B * 4);
        upload_as<R>(h_init, h0, n2); upload_as<R>(q_init, q0, n2);
        add_field("h", h.ptr, n2); add_field("q", q.ptr, n2); add_field("rhsh", rhsh.ptr, n2); add_field("rhsq", rhsq.ptr, n2);
        add_field("u", u_cur.ptr, 1); add_field("up", u_prev.ptr, 1); add_field("stp", stp.ptr, 1, true);
        SloArgs<R> &a = base;
        a.nx = nx; a.n2 = n2; a.ndt_act = p.ndt_act; a.n_act = p.n_act; a.n_interp = p.n_interp; a.obs_smpl = p.obs_smpl; a.n_obs = p.n_obs;
        a.off = ((nx + 1) % C == 0) ? 1 : 0;          // ghost nx+1 must share a chunk with cell nx; ghost 0 with cell 1
        a.B = B;
        a.inv_dx = (R)(1.0 / p.dx); a.hdt = (R)(0.5 * p.dt); a.g = (R)p.g; a.half_g = (R)(0.5 * p.g); a.amp = (R)p.amp;
        a.alpha_pen = (R)p.alpha; a.dx = (R)p.dx; a.blow_lo = (R)p.blow_lo; a.blow_hi = (R)p.blow_hi;
        a.h = h.as<R>(); a.q = q.as<R>(); a.rhsh = rhsh.as<R>(); a.rhsq = rhsq.as<R>(); a.u_cur = u_cur.as<R>(); a.u_prev = u_prev.as<R>();
        a.stp = stp.as<int32_t>(); a.h_init = h_init.as<R>(); a.q_init = q_init.as<R>();
    }
    void reset(const ResetArgs &r) override
    {
        BEACON_REQUIRE(r.obs != nullptr, "reset: obs must not be NULL");
        SloArgs<R> a = base;
        a.mode = 1; a.mask = r.mask; a.obs = (R *)r.obs;
        sloshing_kernel<R, C, T><<<a.B, T, 0, r.stream>>>(a);
        BEACON_CUDA_CHECK(cudaGetLastError());
        launches++;
    }
    void step(const StepArgs &s) override
    {
        SloArgs<R> a = base;
        a.mode = 0; a.n_fused = s.n_fused; a.actions = (const R *)s.actions;
        a.obs = (R *)s.obs; a.rwd = (R *)s.rwd; a.done = s.done; a.trunc = s.trunc; a.status = s.status;
        sloshing_kernel<R, C, T><<<a.B, T, 0, s.stream>>>(a);
        BEACON_CUDA_CHECK(cudaGetLastError());
        launches++;
    }
};

Env *make_sloshing(const beacon_common &c, const beacon_sloshing_params &p, const double *h0, const double *q0)
{
    BEACON_REQUIRE(h0 && q0, "sloshing: init fields must not be NULL");
    if (c.dtype == BEACON_F64) return new SloshingEnv<double>(c, p, h0, q0);
    if (c.dtype == BEACON_F32) return new SloshingEnv<float>(c, p, h0, q0);
    throw Error(BEACON_ERR_INVALID, "unknown dtype");
}

}  // namespace beacon
