/*
 * beacon_oracle.c — CPU restatement of the solver hot path of jviquerat/beacon.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path and the
 * `cpu_baseline` / `--impl reference` arm of bench.py.  It must never be linked, imported
 * or called by the product package (beacon_b200/); only tests/, __graft_entry__.smoke()
 * and bench.py's CPU-baseline legs may use it.
 *
 * Parity is PINNED: every function below is checked bit-for-bit (fields) against the
 * unmodified Python reference imported in the build container (oracle/refload.py) and
 * against the committed golden vectors tests/golden/<env>.npz generated from it by
 * oracle/gen_golden.py (tests/test_oracle_golden.py, tests/test_oracle_vs_reference.py).
 *
 * Arithmetic follows the reference operation-for-operation (same association, true
 * divisions, no FMA contraction: build with -ffp-contract=off, no -ffast-math), so that
 * numba/numpy and this file round identically.  All citations are file:line under
 * /root/reference/beacon/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

#define ORC_API __attribute__((visibility("default")))

/* numpy's maximum/minimum propagate NaN from either argument. */
static inline double np_max(double a, double b) { return isnan(a) ? a : (a > b ? a : b); }
static inline double np_min(double a, double b) { return isnan(a) ? a : (a < b ? a : b); }

/* numpy pairwise summation (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum_DOUBLE):
 * used by np.sum / np.mean on contiguous float64; reproduced so rewards match bitwise. */
static double np_pairwise_sum(const double *a, long n)
{
    if (n < 8) {
        double res = 0.0;
        /* numpy starts from -0.0 to preserve the sign of -0.0 sums; value-identical otherwise */
        res = -0.0;
        for (long i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        long i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8) {
            r[0] += a[i + 0]; r[1] += a[i + 1]; r[2] += a[i + 2]; r[3] += a[i + 3];
            r[4] += a[i + 4]; r[5] += a[i + 5]; r[6] += a[i + 6]; r[7] += a[i + 7];
        }
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
    }
}

ORC_API double orc_np_sum(const double *a, long n) { return np_pairwise_sum(a, n); }

/* --- tiny pthread parallel-for (the image's gcc has no usable libgomp spec) ------------ */
static int g_threads = 0;

ORC_API int orc_num_threads(void)
{
    if (g_threads > 0) return g_threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

ORC_API void orc_set_num_threads(int n) { g_threads = n; }

typedef void (*orc_body_fn)(int b, void *ctx, void *tls);
typedef struct {
    orc_body_fn fn; void *ctx; int B; int *next; pthread_mutex_t *mu; size_t tls_bytes;
} orc_pf_job;

static void *orc_pf_worker(void *arg)
{
    orc_pf_job *job = (orc_pf_job *)arg;
    void *tls = job->tls_bytes ? calloc(1, job->tls_bytes) : NULL;
    for (;;) {
        pthread_mutex_lock(job->mu);
        int b = (*job->next)++;
        pthread_mutex_unlock(job->mu);
        if (b >= job->B) break;
        job->fn(b, job->ctx, tls);
    }
    free(tls);
    return NULL;
}

/* Runs fn(b, ctx, tls) for b in [0,B) on orc_num_threads() threads; tls is a zeroed
 * per-thread scratch block of tls_bytes. */
static void orc_parallel_for(int B, orc_body_fn fn, void *ctx, size_t tls_bytes)
{
    int nt = orc_num_threads();
    if (nt > B) nt = B;
    if (nt < 1) nt = 1;
    int next = 0;
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    orc_pf_job job = {fn, ctx, B, &next, &mu, tls_bytes};
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nt);
    for (int t = 1; t < nt; t++) pthread_create(&th[t], NULL, orc_pf_worker, &job);
    orc_pf_worker(&job);
    for (int t = 1; t < nt; t++) pthread_join(th[t], NULL);
    free(th);
}

/* ------------------------------------------------------------------------------------
 * shkadov  (shkadov/shkadov.py)
 * ---------------------------------------------------------------------------------- */

typedef struct {
    int nx, ndt_act, n_interp, n_jets, jet_pos, jet_space, jet_hw, l_obs, l_rwd, obs_stride, n_obs;
    double dx, dt, delta, eps, jet_amp;
} orc_shkadov_cfg;

/* d1tvd, shkadov.py:494-504 — minmod-limited upwind first derivative.  phi is scratch[nx]. */
static void shk_d1tvd(const double *u, double *du, int nx, double dx, double *phi)
{
    phi[0] = 0.0;
    if (nx > 1) phi[nx - 1] = 0.0;
    for (int i = 1; i < nx - 1; i++) {
        double r = (u[i] - u[i - 1]) / (u[i + 1] - u[i] + 1.0e-8);
        phi[i] = np_max(0.0, np_min(r, 1.0));
    }
    for (int i = 1; i < nx - 1; i++) {
        double d = u[i] + (0.5 * phi[i]) * (u[i + 1] - u[i]);
        d -= u[i - 1] + (0.5 * phi[i - 1]) * (u[i] - u[i - 1]);
        d /= dx;
        du[i] = d;
    }
}

/* d3o2u, shkadov.py:485-491 — third derivative, 2nd-order upwind-biased, two closures. */
static void shk_d3o2u(const double *u, double *du, int nx, double dx)
{
    double den = 2.0 * dx * dx * dx;
    for (int i = 1; i < nx - 3; i++)
        du[i] = (-u[i + 3] + 6.0 * u[i + 2] - 12.0 * u[i + 1] + 10.0 * u[i] - 3.0 * u[i - 1]) / den;
    double d3 = dx * dx * dx;
    du[nx - 3] = (u[nx - 1] - 3.0 * u[nx - 2] + 3.0 * u[nx - 3] - u[nx - 4]) / d3;
    du[nx - 2] = (-u[nx - 4] + 3.0 * u[nx - 3] - 3.0 * u[nx - 2] + u[nx - 1]) / d3;
}

/* solve(), shkadov.py:188-236.  `up`/`u` are previous/current actions [n_jets];
 * noise[ndt_act] replaces the np.random.uniform draw of :204 (one per sub-step).
 * scratch: 5*nx doubles (q2h, dq2h, dddh, phi, spare). */
ORC_API void orc_shkadov_solve(const orc_shkadov_cfg *c, double *h, double *q, double *rhsh,
                               double *rhsq, double *rhshp, double *rhsqp, const double *up,
                               const double *u, const double *noise, double *scratch)
{
    const int nx = c->nx;
    double *q2h = scratch, *dq2h = scratch + nx, *dddh = scratch + 2 * nx, *phi = scratch + 3 * nx;
    /* dq2h/dddh edge entries are never written by the reference (stay 0). */
    dq2h[0] = dq2h[nx - 1] = 0.0;
    dddh[0] = dddh[nx - 1] = 0.0;
    for (int it = 0; it < c->ndt_act; it++) {
        memcpy(rhshp, rhsh, sizeof(double) * nx);               /* :200 */
        memcpy(rhsqp, rhsq, sizeof(double) * nx);               /* :201 */
        h[0] = 1.0 + (noise ? noise[it] : 0.0);                 /* :204 */
        q[0] = 1.0;                                             /* :205 */
        h[nx - 1] = h[nx - 2];                                  /* :206 */
        q[nx - 1] = q[nx - 2];                                  /* :207 */
        shk_d1tvd(q, rhsh, nx, c->dx, phi);                     /* :210 */
        for (int i = 0; i < nx; i++) q2h[i] = q[i] * q[i] / (h[i] + c->eps); /* :213 */
        shk_d1tvd(q2h, dq2h, nx, c->dx, phi);                   /* :214 */
        shk_d3o2u(h, dddh, nx, c->dx);                          /* :217 */
        {                                                       /* rhsq(), :507-512 */
            double p = 1.0 / (5.0 * c->delta);
            for (int i = 1; i < nx - 1; i++)
                rhsq[i] = 1.2 * dq2h[i] - p * (h[i] * (dddh[i] + 1.0) - q[i] / (h[i] * h[i] + c->eps));
        }
        {                                                       /* jets, :223-232 */
            double alpha = fmin((double)it / (double)c->n_interp, 1.0);
            for (int j = 0; j < c->n_jets; j++) {
                double uj = (1.0 - alpha) * up[j] + alpha * u[j];
                int s = c->jet_pos + j * c->jet_space - c->jet_hw;
                int e = s + 2 * c->jet_hw;
                double den = 0.25 * (double)((long)(e - s) * (long)(e - s));
                for (int k = s; k <= e; k++) {
                    double v = (double)((long)(k - s) * (long)(e - k)) / den;
                    double dq = c->jet_amp * uj * v;
                    rhsq[k] += dq;
                }
            }
        }
        {                                                       /* adams(), :515-518 */
            double hdt = 0.5 * c->dt;
            for (int i = 1; i < nx - 1; i++) h[i] += hdt * (-3.0 * rhsh[i] + rhshp[i]);
            for (int i = 1; i < nx - 1; i++) q[i] += hdt * (-3.0 * rhsq[i] + rhsqp[i]);
        }
    }
}

/* get_obs(), shkadov.py:239-250: q[s:e:stride] per jet, jet-major. */
ORC_API void orc_shkadov_obs(const orc_shkadov_cfg *c, const double *q, double *obs)
{
    for (int j = 0; j < c->n_jets; j++) {
        int s = c->jet_pos + j * c->jet_space - c->l_obs;
        for (int k = 0; k < c->n_obs; k++) obs[j * c->n_obs + k] = q[s + k * c->obs_stride];
    }
}

/* get_rwd(), shkadov.py:253-264 (total) and :469-481 (per jet, separable): per_jet[j] is the
 * separable reward of jet j; return value is the joint reward. */
ORC_API double orc_shkadov_rwd(const orc_shkadov_cfg *c, const double *h, double *per_jet)
{
    double rwd = 0.0;
    double *sq = (double *)malloc(sizeof(double) * (c->l_rwd > 0 ? c->l_rwd : 1));
    for (int j = 0; j < c->n_jets; j++) {
        int s = c->jet_pos + j * c->jet_space;
        for (int k = 0; k < c->l_rwd; k++) {
            double d = h[s + k] - 1.0;
            sq[k] = d * d;
        }
        double term = np_pairwise_sum(sq, c->l_rwd) * c->dx;
        rwd -= term;
        if (per_jet) {
            double r = 0.0;
            r -= term;
            r /= (double)(c->n_jets * c->l_rwd);
            per_jet[j] = r;
        }
    }
    free(sq);
    rwd /= (double)(c->n_jets * c->l_rwd);
    return rwd;
}

/* blow-up guard, shkadov.py:176: any(h < -5 h_max | h > 5 h_max) with h_max = 5. */
ORC_API int orc_shkadov_blowup(const orc_shkadov_cfg *c, const double *h)
{
    for (int i = 0; i < c->nx; i++)
        if (h[i] < -25.0 || h[i] > 25.0) return 1;
    return 0;
}

/* Batched driver for the CPU baseline: one env per worker thread, state arrays [B, nx],
 * actions [B, n_jets]; u_prev/u_cur [B, n_jets] are updated (up<-u, u<-a, :193-194). */
typedef struct {
    const orc_shkadov_cfg *c; double *h, *q, *rhsh, *rhsq, *u_prev, *u_cur;
    const double *actions, *noise; double *obs, *rwd; uint8_t *blowup;
} shk_batch_ctx;

static void shk_batch_body(int b, void *vctx, void *tls)
{
    shk_batch_ctx *k = (shk_batch_ctx *)vctx;
    const orc_shkadov_cfg *c = k->c;
    const int nx = c->nx;
    double *scratch = (double *)tls;
    double *up = k->u_prev + (size_t)b * c->n_jets, *uc = k->u_cur + (size_t)b * c->n_jets;
    for (int j = 0; j < c->n_jets; j++) { up[j] = uc[j]; uc[j] = k->actions[(size_t)b * c->n_jets + j]; }
    orc_shkadov_solve(c, k->h + (size_t)b * nx, k->q + (size_t)b * nx, k->rhsh + (size_t)b * nx,
                      k->rhsq + (size_t)b * nx, scratch + 5 * nx, scratch + 6 * nx, up, uc,
                      k->noise ? k->noise + (size_t)b * c->ndt_act : NULL, scratch);
    orc_shkadov_obs(c, k->q + (size_t)b * nx, k->obs + (size_t)b * c->n_obs * c->n_jets);
    k->rwd[b] = orc_shkadov_rwd(c, k->h + (size_t)b * nx, NULL);
    k->blowup[b] = (uint8_t)orc_shkadov_blowup(c, k->h + (size_t)b * nx);
    if (k->blowup[b]) k->rwd[b] = -1.0;
}

ORC_API void orc_shkadov_step_batch(const orc_shkadov_cfg *c, int B, double *h, double *q,
                                    double *rhsh, double *rhsq, double *u_prev, double *u_cur,
                                    const double *actions, const double *noise, double *obs,
                                    double *rwd, uint8_t *blowup)
{
    shk_batch_ctx k = {c, h, q, rhsh, rhsq, u_prev, u_cur, actions, noise, obs, rwd, blowup};
    orc_parallel_for(B, shk_batch_body, &k, sizeof(double) * c->nx * 7);
}

/* ------------------------------------------------------------------------------------
 * burgers  (burgers/burgers.py)
 * ---------------------------------------------------------------------------------- */

/* derx(), burgers.py:231-243 — van-Leer limited flux form.  phi is scratch[nx]. */
static void bur_derx(const double *u, double *du, int nx, double dx, double *phi)
{
    phi[0] = 0.0;
    phi[nx - 1] = 0.0;
    for (int i = 1; i < nx - 1; i++) {
        double r = (u[i] - u[i - 1]) / (u[i + 1] - u[i] + 1.0e-8);
        phi[i] = (r + fabs(r)) / (1.0 + r);
    }
    for (int i = 1; i < nx - 1; i++) {
        double fp = u[i] + (0.5 * phi[i]) * (u[i + 1] - u[i]);
        double fm = u[i - 1] + (0.5 * phi[i - 1]) * (u[i] - u[i - 1]);
        du[i] = (fp - fm) / dx;
    }
}

/* solve(), burgers.py:119-151.  noise = the single np.random.uniform draw of :127.
 * scratch: 3*nx doubles. */
ORC_API void orc_burgers_solve(int nx, double dx, double dt, int ndt_act, int ctrl_pos, double amp,
                               double u_target, double *u, double *up, double *upp, double a,
                               double noise, double *scratch)
{
    double *du = scratch, *rhs = scratch + nx, *phi = scratch + 2 * nx;
    du[0] = du[nx - 1] = 0.0;
    rhs[0] = rhs[nx - 1] = 0.0;
    for (int it = 0; it < ndt_act; it++) {
        memcpy(upp, up, sizeof(double) * nx);                   /* :135 */
        memcpy(up, u, sizeof(double) * nx);                     /* :136 */
        u[0] = u_target + noise;                                /* :139 */
        u[nx - 1] = u[nx - 2];                                  /* :140 */
        bur_derx(u, du, nx, dx, phi);                           /* :143 */
        for (int i = 1; i < nx - 1; i++) rhs[i] = u[i] * du[i]; /* rhs(), :253-255 */
        rhs[ctrl_pos] += a * amp;                               /* :149 */
        for (int i = 1; i < nx - 1; i++)                        /* dert(), :247-249 */
            u[i] = (4.0 * up[i] - upp[i] - 2.0 * dt * rhs[i]) / 3.0;
    }
}

/* get_rwd(), burgers.py:162-166 */
ORC_API double orc_burgers_rwd(int nx, double dx, int ctrl_pos, double u_target, const double *u)
{
    int n = nx - ctrl_pos;
    double *t = (double *)malloc(sizeof(double) * (n > 0 ? n : 1));
    for (int i = 0; i < n; i++) t[i] = fabs(u[ctrl_pos + i] - u_target);
    double r = -np_pairwise_sum(t, n) * dx;
    free(t);
    return r;
}

typedef struct {
    int nx, ndt_act, ctrl_pos, n_obs; double dx, dt, amp, u_target;
    double *u, *up, *upp; const double *actions, *noise; double *obs, *rwd;
} bur_batch_ctx;

static void bur_batch_body(int b, void *vctx, void *tls)
{
    bur_batch_ctx *k = (bur_batch_ctx *)vctx;
    const int nx = k->nx;
    double *ub = k->u + (size_t)b * nx;
    orc_burgers_solve(nx, k->dx, k->dt, k->ndt_act, k->ctrl_pos, k->amp, k->u_target, ub,
                      k->up + (size_t)b * nx, k->upp + (size_t)b * nx, k->actions[b],
                      k->noise ? k->noise[b] : 0.0, (double *)tls);
    for (int i = 0; i < k->n_obs; i++) k->obs[(size_t)b * k->n_obs + i] = ub[k->ctrl_pos - k->n_obs + i];
    k->rwd[b] = orc_burgers_rwd(nx, k->dx, k->ctrl_pos, k->u_target, ub);
}

ORC_API void orc_burgers_step_batch(int B, int nx, double dx, double dt, int ndt_act, int ctrl_pos,
                                    double amp, double u_target, int n_obs, double *u, double *up,
                                    double *upp, const double *actions, const double *noise,
                                    double *obs, double *rwd)
{
    bur_batch_ctx k = {nx, ndt_act, ctrl_pos, n_obs, dx, dt, amp, u_target, u, up, upp, actions, noise, obs, rwd};
    orc_parallel_for(B, bur_batch_body, &k, sizeof(double) * nx * 3);
}

/* ------------------------------------------------------------------------------------
 * sloshing  (sloshing/sloshing.py)
 * ---------------------------------------------------------------------------------- */

/* solve(), sloshing.py:168-224.  Arrays have nx+2 entries (ghosts at 0 and nx+1).
 * scratch: 7*(nx+2) doubles. */
ORC_API void orc_sloshing_solve(int nx, double dx, double dt, int ndt_act, int n_interp, double g,
                                double amp, double *h, double *q, double *rhsh, double *rhsq,
                                double *rhshp, double *rhsqp, double u_prev, double u_cur,
                                double *scratch)
{
    const int n2 = nx + 2;
    double *v = scratch, *qgh = scratch + n2, *cc = scratch + 2 * n2, *fhg = scratch + 3 * n2,
           *fhd = scratch + 4 * n2, *fqg = scratch + 5 * n2, *fqd = scratch + 6 * n2;
    for (int it = 0; it < ndt_act; it++) {
        h[0] = h[1];                                            /* :183 */
        q[0] = 0.0;
        h[nx + 1] = h[nx];
        q[nx + 1] = 0.0;
        for (int i = 1; i <= nx; i++) { rhshp[i] = rhsh[i]; rhsqp[i] = rhsq[i]; } /* :189-190 */
        for (int i = 0; i < n2; i++) {                          /* :193-194 */
            v[i] = q[i] / h[i];
            qgh[i] = q[i] * q[i] / h[i] + 0.5 * g * (h[i] * h[i]);
        }
        for (int i = 0; i <= nx; i++)                           /* :197-199 */
            cc[i] = np_max(fabs(v[i]) + sqrt(g * h[i]), fabs(v[i + 1]) + sqrt(g * h[i + 1]));
        for (int i = 1; i <= nx; i++) {                         /* rusanov(), :323-325; calls :202-211 */
            fhg[i] = 0.5 * (q[i - 1] + q[i]) - 0.5 * cc[i - 1] * (h[i] - h[i - 1]);
            fhd[i] = 0.5 * (q[i] + q[i + 1]) - 0.5 * cc[i] * (h[i + 1] - h[i]);
            fqg[i] = 0.5 * (qgh[i - 1] + qgh[i]) - 0.5 * cc[i - 1] * (q[i] - q[i - 1]);
            fqd[i] = 0.5 * (qgh[i] + qgh[i + 1]) - 0.5 * cc[i] * (q[i + 1] - q[i]);
        }
        for (int i = 1; i <= nx; i++) {                         /* :214-215 */
            rhsh[i] = (fhd[i] - fhg[i]) / dx;
            rhsq[i] = (fqd[i] - fqg[i]) / dx;
        }
        {                                                       /* :218-220 */
            double alpha = fmin((double)it / (double)n_interp, 1.0);
            double uu = (1.0 - alpha) * u_prev + alpha * u_cur;
            double f = uu * amp;
            for (int i = 1; i <= nx; i++) rhsq[i] += f;
        }
        {                                                       /* adams(), :329-331 */
            double hdt = 0.5 * dt;
            for (int i = 1; i <= nx; i++) h[i] += hdt * (-3.0 * rhsh[i] + rhshp[i]);
            for (int i = 1; i <= nx; i++) q[i] += hdt * (-3.0 * rhsq[i] + rhsqp[i]);
        }
    }
}

typedef struct {
    int nx, ndt_act, n_interp; double dx, dt, g, amp, alpha_pen;
    double *h, *q, *rhsh, *rhsq, *u_prev, *u_cur; const double *actions; double *obs, *rwd;
} slo_batch_ctx;

static void slo_batch_body(int b, void *vctx, void *tls)
{
    slo_batch_ctx *k = (slo_batch_ctx *)vctx;
    const int nx = k->nx, n2 = nx + 2;
    const int n_obs = nx / 2 + (nx % 2 != 0);
    double *scratch = (double *)tls;
    double *hb = k->h + (size_t)b * n2, *qb = k->q + (size_t)b * n2;
    k->u_prev[b] = k->u_cur[b];
    k->u_cur[b] = k->actions[b];
    orc_sloshing_solve(nx, k->dx, k->dt, k->ndt_act, k->n_interp, k->g, k->amp, hb, qb,
                       k->rhsh + (size_t)b * n2, k->rhsq + (size_t)b * n2, scratch + 7 * n2,
                       scratch + 8 * n2, k->u_prev[b], k->u_cur[b], scratch);
    for (int i = 0; i < n_obs; i++) k->obs[(size_t)b * n_obs + i] = qb[1 + 2 * i];
    double s = 0.0;
    for (int i = 1; i <= nx; i++) s += (hb[i] - 1.0) * (hb[i] - 1.0);
    k->rwd[b] = -sqrt(s) * k->dx - k->alpha_pen * fabs(k->amp * k->u_cur[b]);
}

ORC_API void orc_sloshing_step_batch(int B, int nx, double dx, double dt, int ndt_act, int n_interp,
                                     double g, double amp, double alpha_pen, double *h, double *q,
                                     double *rhsh, double *rhsq, double *u_prev, double *u_cur,
                                     const double *actions, double *obs, double *rwd)
{
    slo_batch_ctx k = {nx, ndt_act, n_interp, dx, dt, g, amp, alpha_pen, h, q, rhsh, rhsq, u_prev, u_cur, actions, obs, rwd};
    orc_parallel_for(B, slo_batch_body, &k, sizeof(double) * (nx + 2) * 9);
}

/* ------------------------------------------------------------------------------------
 * lorenz / vortex  (lorenz/lorenz.py, vortex/vortex.py) — 5-stage LSRK4
 * ---------------------------------------------------------------------------------- */

/* lsrk4 coefficients, lorenz.py:272-277 (vortex.py identical). */
static const double LSRK_A[5] = {0.000000000000000, -0.417890474499852, -1.192151694642677,
                                 -1.697784692471528, -1.514183444257156};
static const double LSRK_B[5] = {0.149659021999229, 0.379210312999627, 0.822955029386982,
                                 0.699450455949122, 0.153057247968152};

/* solve(), lorenz.py:120-153 with lsrk4.update :293-297.  x is used as the low-storage
 * register, xk as the solution, exactly as in the reference; fx keeps the last-stage rhs. */
ORC_API void orc_lorenz_solve(double sigma, double rho, double beta, double dt, int ndt_act,
                              double forcing, double *x, double *xk, double *fx)
{
    for (int it = 0; it < ndt_act; it++) {
        for (int i = 0; i < 3; i++) xk[i] = x[i];
        for (int j = 0; j < 5; j++) {
            fx[0] = sigma * (xk[1] - xk[0]);
            fx[1] = xk[0] * (rho - xk[2]) - xk[1];
            fx[2] = xk[0] * xk[1] - beta * xk[2];
            fx[1] += forcing;
            for (int i = 0; i < 3; i++) {
                x[i] = LSRK_A[j] * x[i] + dt * fx[i];
                xk[i] += LSRK_B[j] * x[i];
            }
        }
        for (int i = 0; i < 3; i++) x[i] = xk[i];
    }
}

typedef struct {
    double sigma, rho, beta, dt; int ndt_act; double *x, *fx; const int32_t *actions; double *obs, *rwd;
    int chunk, B;
} lor_batch_ctx;

static void lor_batch_body(int blk, void *vctx, void *tls)
{
    (void)tls;
    lor_batch_ctx *k = (lor_batch_ctx *)vctx;
    static const double F[3] = {-1.0, 0.0, 1.0};
    int b1 = (blk + 1) * k->chunk;
    if (b1 > k->B) b1 = k->B;
    for (int b = blk * k->chunk; b < b1; b++) {
        double xk[3];
        orc_lorenz_solve(k->sigma, k->rho, k->beta, k->dt, k->ndt_act, F[k->actions[b]], k->x + 3 * b, xk, k->fx + 3 * b);
        for (int i = 0; i < 3; i++) { k->obs[6 * b + i] = k->x[3 * b + i]; k->obs[6 * b + 3 + i] = k->fx[3 * b + i]; }
        k->rwd[b] = k->x[3 * b] < 0.0 ? 1.0 : 0.0;
    }
}

ORC_API void orc_lorenz_step_batch(int B, double sigma, double rho, double beta, double dt,
                                   int ndt_act, double *x, double *fx, const int32_t *actions,
                                   double *obs, double *rwd)
{
    lor_batch_ctx k = {sigma, rho, beta, dt, ndt_act, x, fx, actions, obs, rwd, 1024, B};
    orc_parallel_for((B + k.chunk - 1) / k.chunk, lor_batch_body, &k, 0);
}

typedef struct {
    double lmbda_re, lmbda_cx, mu_re, mu_cx, alpha_re, alpha_cx, ire, omega_f, gamma, domega, beta_m;
} orc_vortex_cfg;

/* solve(), vortex.py:149-183.  kmod/kphase from the action (:156-157) are computed by the caller. */
ORC_API void orc_vortex_solve(const orc_vortex_cfg *c, double dt, int ndt_act, double kmod,
                              double kphase, double *x, double *xk, double *fx)
{
    const double ck = cos(kphase), sk = sin(kphase);
    for (int it = 0; it < ndt_act; it++) {
        for (int i = 0; i < 4; i++) xk[i] = x[i];
        for (int j = 0; j < 5; j++) {
            double n2 = xk[0] * xk[0] + xk[1] * xk[1];
            fx[0] = c->ire * (c->lmbda_re * xk[0] - c->lmbda_cx * xk[1]) -
                    (c->mu_re * xk[0] - c->mu_cx * xk[1]) * n2 +
                    (c->alpha_re * xk[2] - c->alpha_cx * xk[3]) + xk[0] * kmod * ck - xk[1] * kmod * sk;
            fx[1] = c->ire * (c->lmbda_re * xk[1] + c->lmbda_cx * xk[0]) -
                    (c->mu_re * xk[1] + c->mu_cx * xk[0]) * n2 +
                    (c->alpha_re * xk[3] + c->alpha_cx * xk[2]) + xk[0] * kmod * sk + xk[1] * kmod * ck;
            fx[2] = -c->omega_f * c->gamma * xk[2] - c->domega * xk[3] + c->beta_m * xk[0];
            fx[3] = -c->omega_f * c->gamma * xk[3] + c->domega * xk[2] + c->beta_m * xk[1];
            for (int i = 0; i < 4; i++) {
                x[i] = LSRK_A[j] * x[i] + dt * fx[i];
                xk[i] += LSRK_B[j] * x[i];
            }
        }
        for (int i = 0; i < 4; i++) x[i] = xk[i];
    }
}

/* ------------------------------------------------------------------------------------
 * rayleigh / mixing  (rayleigh/rayleigh.py, mixing/mixing.py) — 2D MAC projection
 * Arrays are [(nx+2),(ny+2)] C-order: index i*(ny+2)+j.
 * ---------------------------------------------------------------------------------- */

#define IX(i, j) ((size_t)(i) * (size_t)(ny + 2) + (size_t)(j))

typedef struct {
    int nx, ny, ndt_act, kind;          /* kind 0 = rayleigh, 1 = mixing */
    double dx, dy, dt;
    double pr, ra;                      /* rayleigh */
    double re, pe;                      /* mixing */
    double tol;                         /* 1e-8 rayleigh.py:414, 1e-4 mixing.py:423 */
    int itmax;                          /* 300000 */
} orc_mac_cfg;

/* predictor(): rayleigh.py:371-407 (buoyancy +T, diff*sqrt(pr/ra)); mixing.py:382-416 (diff/re). */
static void mac_predictor(const orc_mac_cfg *c, const double *u, const double *v, double *us,
                          double *vs, const double *p, const double *T)
{
    const int nx = c->nx, ny = c->ny;
    const double dx = c->dx, dy = c->dy, dt = c->dt;
    const double dx2 = dx * dx, dy2 = dy * dy;
    const double sq = (c->kind == 0) ? sqrt(c->pr / c->ra) : 0.0;
    for (int i = 2; i <= nx; i++)
        for (int j = 1; j <= ny; j++) {
            double uE = 0.5 * (u[IX(i + 1, j)] + u[IX(i, j)]);
            double uW = 0.5 * (u[IX(i, j)] + u[IX(i - 1, j)]);
            double uN = 0.5 * (u[IX(i, j + 1)] + u[IX(i, j)]);
            double uS = 0.5 * (u[IX(i, j)] + u[IX(i, j - 1)]);
            double vN = 0.5 * (v[IX(i, j + 1)] + v[IX(i - 1, j + 1)]);
            double vS = 0.5 * (v[IX(i, j)] + v[IX(i - 1, j)]);
            double conv = (uE * uE - uW * uW) / dx + (uN * vN - uS * vS) / dy;
            double diff = ((u[IX(i + 1, j)] - 2.0 * u[IX(i, j)] + u[IX(i - 1, j)]) / dx2 +
                           (u[IX(i, j + 1)] - 2.0 * u[IX(i, j)] + u[IX(i, j - 1)]) / dy2);
            if (c->kind == 0) diff *= sq; else diff = diff / c->re;
            double pres = (p[IX(i, j)] - p[IX(i - 1, j)]) / dx;
            us[IX(i, j)] = u[IX(i, j)] + dt * (diff - conv - pres);
        }
    for (int i = 1; i <= nx; i++)
        for (int j = 2; j <= ny; j++) {
            double vE = 0.5 * (v[IX(i + 1, j)] + v[IX(i, j)]);
            double vW = 0.5 * (v[IX(i, j)] + v[IX(i - 1, j)]);
            double uE = 0.5 * (u[IX(i + 1, j)] + u[IX(i + 1, j - 1)]);
            double uW = 0.5 * (u[IX(i, j)] + u[IX(i, j - 1)]);
            double vN = 0.5 * (v[IX(i, j + 1)] + v[IX(i, j)]);
            double vS = 0.5 * (v[IX(i, j)] + v[IX(i, j - 1)]);
            double conv = (uE * vE - uW * vW) / dx + (vN * vN - vS * vS) / dy;
            double diff = ((v[IX(i + 1, j)] - 2.0 * v[IX(i, j)] + v[IX(i - 1, j)]) / dx2 +
                           (v[IX(i, j + 1)] - 2.0 * v[IX(i, j)] + v[IX(i, j - 1)]) / dy2);
            if (c->kind == 0) diff *= sq; else diff = diff / c->re;
            double pres = (p[IX(i, j)] - p[IX(i, j - 1)]) / dy;
            if (c->kind == 0)
                vs[IX(i, j)] = v[IX(i, j)] + dt * (diff - conv - pres + T[IX(i, j)]);
            else
                vs[IX(i, j)] = v[IX(i, j)] + dt * (diff - conv - pres);
        }
}

/* poisson(): rayleigh.py:412-456, mixing.py:421-465.  Jacobi; b is sweep-invariant so it is
 * evaluated once per solve (identical values to the per-sweep evaluation of :428-429).
 * Residual = sum over the whole ghost-inclusive array of (phi-phin)^2 (np.dot :449; BLAS
 * summation order is not reproducible, it only feeds the `err > tol` test).
 * work: 2*(nx+2)*(ny+2) doubles (phin, b).  Returns itp; *ovf set on overflow. */
static int mac_poisson(const orc_mac_cfg *c, const double *us, const double *vs, double *phi,
                       double *work, int *ovf)
{
    const int nx = c->nx, ny = c->ny;
    const double dx = c->dx, dy = c->dy, dt = c->dt;
    const size_t n = (size_t)(nx + 2) * (size_t)(ny + 2);
    double *phin = work, *bb = work + n;
    double err = 1.0e10;
    int itp = 0;
    *ovf = 0;
    memset(phi, 0, sizeof(double) * n);
    memset(phin, 0, sizeof(double) * n);
    for (int i = 1; i <= nx; i++)
        for (int j = 1; j <= ny; j++) {
            double b = ((us[IX(i + 1, j)] - us[IX(i, j)]) / dx + (vs[IX(i, j + 1)] - vs[IX(i, j)]) / dy) / dt;
            bb[IX(i, j)] = b * dx * dx * dy * dy;
        }
    const double den = dx * dx + dy * dy;
    while (err > c->tol) {
        memcpy(phin, phi, sizeof(double) * n);
        for (int i = 1; i <= nx; i++)
            for (int j = 1; j <= ny; j++)
                phi[IX(i, j)] = 0.5 * ((phin[IX(i + 1, j)] + phin[IX(i - 1, j)]) * dy * dy +
                                       (phin[IX(i, j + 1)] + phin[IX(i, j - 1)]) * dx * dx - bb[IX(i, j)]) / den;
        for (int j = 1; j <= ny; j++) { phi[IX(0, j)] = phi[IX(1, j)]; phi[IX(nx + 1, j)] = phi[IX(nx, j)]; }
        for (int i = 1; i <= nx; i++) {
            phi[IX(i, ny + 1)] = (c->kind == 0) ? phi[IX(i, ny)] : 0.0;   /* rayleigh.py:442 / mixing.py:451 */
            phi[IX(i, 0)] = phi[IX(i, 1)];
        }
        err = 0.0;
        for (size_t k = 0; k < n; k++) { double d = phi[k] - phin[k]; err += d * d; }
        itp += 1;
        if (itp > c->itmax) { *ovf = 1; break; }
    }
    return itp;
}

/* corrector(): rayleigh.py:461-464, mixing.py:470-473 */
static void mac_corrector(const orc_mac_cfg *c, double *u, double *v, const double *us,
                          const double *vs, const double *phi)
{
    const int nx = c->nx, ny = c->ny;
    for (int i = 2; i <= nx; i++)
        for (int j = 1; j <= ny; j++)
            u[IX(i, j)] = us[IX(i, j)] - c->dt * (phi[IX(i, j)] - phi[IX(i - 1, j)]) / c->dx;
    for (int i = 1; i <= nx; i++)
        for (int j = 2; j <= ny; j++)
            v[IX(i, j)] = vs[IX(i, j)] - c->dt * (phi[IX(i, j)] - phi[IX(i, j - 1)]) / c->dy;
}

/* transport(): rayleigh.py:469-487 (diff/sqrt(pr*ra)), mixing.py:478-495 (diff/pe);
 * in place, lexicographic i-outer j-inner (new W,S neighbours, old E,N). */
static void mac_transport(const orc_mac_cfg *c, const double *u, const double *v, double *T)
{
    const int nx = c->nx, ny = c->ny;
    const double dx = c->dx, dy = c->dy, dt = c->dt;
    const double dx2 = dx * dx, dy2 = dy * dy;
    const double sq = (c->kind == 0) ? sqrt(c->pr * c->ra) : c->pe;
    for (int i = 1; i <= nx; i++)
        for (int j = 1; j <= ny; j++) {
            double uE = u[IX(i + 1, j)], uW = u[IX(i, j)], vN = v[IX(i, j + 1)], vS = v[IX(i, j)];
            double TE = 0.5 * (T[IX(i + 1, j)] + T[IX(i, j)]);
            double TW = 0.5 * (T[IX(i - 1, j)] + T[IX(i, j)]);
            double TN = 0.5 * (T[IX(i, j + 1)] + T[IX(i, j)]);
            double TS = 0.5 * (T[IX(i, j - 1)] + T[IX(i, j)]);
            double conv = (uE * TE - uW * TW) / dx + (vN * TN - vS * TS) / dy;
            double diff = ((T[IX(i + 1, j)] - 2.0 * T[IX(i, j)] + T[IX(i - 1, j)]) / dx2 +
                           (T[IX(i, j + 1)] - 2.0 * T[IX(i, j)] + T[IX(i, j - 1)]) / dy2);
            diff = diff / sq;
            T[IX(i, j)] += dt * (diff - conv);
        }
}

/* Boundary conditions: rayleigh.py:180-202 (seg_val[n_sgts] = Th + a_j already conditioned;
 * nx_sgts cells per segment), mixing.py:153-171 (wall = {u_t,u_b,v_l,v_r}). */
static void mac_bcs(const orc_mac_cfg *c, double *u, double *v, double *T, double Tc,
                    const double *seg_val, int n_sgts, int nx_sgts, const double *wall)
{
    const int nx = c->nx, ny = c->ny;
    if (c->kind == 0) {
        for (int j = 1; j <= ny; j++) u[IX(1, j)] = 0.0;
        for (int j = 2; j <= ny; j++) v[IX(0, j)] = -v[IX(1, j)];
        for (int j = 1; j <= ny; j++) T[IX(0, j)] = T[IX(1, j)];
        for (int j = 1; j <= ny; j++) u[IX(nx + 1, j)] = 0.0;
        for (int j = 2; j <= ny; j++) v[IX(nx + 1, j)] = -v[IX(nx, j)];
        for (int j = 1; j <= ny; j++) T[IX(nx + 1, j)] = T[IX(nx, j)];
        for (int i = 1; i <= nx + 1; i++) u[IX(i, ny + 1)] = -u[IX(i, ny)];
        for (int i = 1; i <= nx; i++) v[IX(i, ny + 1)] = 0.0;
        for (int i = 1; i <= nx; i++) T[IX(i, ny + 1)] = 2.0 * Tc - T[IX(i, ny)];
        for (int i = 1; i <= nx + 1; i++) u[IX(i, 0)] = -u[IX(i, 1)];
        for (int i = 1; i <= nx; i++) v[IX(i, 1)] = 0.0;
        for (int s = 0; s < n_sgts; s++)
            for (int i = 1 + s * nx_sgts; i < 1 + (s + 1) * nx_sgts; i++)
                T[IX(i, 0)] = 2.0 * seg_val[s] - T[IX(i, 1)];
    } else {
        const double u_t = wall[0], u_b = wall[1], v_l = wall[2], v_r = wall[3];
        for (int j = 1; j <= ny; j++) u[IX(1, j)] = 0.0;
        for (int j = 2; j <= ny; j++) v[IX(0, j)] = 2.0 * v_l - v[IX(1, j)];
        for (int j = 1; j <= ny; j++) T[IX(0, j)] = T[IX(1, j)];
        for (int j = 1; j <= ny; j++) u[IX(nx + 1, j)] = 0.0;
        for (int j = 2; j <= ny; j++) v[IX(nx + 1, j)] = 2.0 * v_r - v[IX(nx, j)];
        for (int j = 1; j <= ny; j++) T[IX(nx + 1, j)] = T[IX(nx, j)];
        for (int i = 1; i <= nx + 1; i++) u[IX(i, ny + 1)] = 2.0 * u_t - u[IX(i, ny)];
        for (int i = 1; i <= nx; i++) v[IX(i, ny + 1)] = 0.0;
        for (int i = 1; i <= nx; i++) T[IX(i, ny + 1)] = T[IX(i, ny)];
        for (int i = 1; i <= nx + 1; i++) u[IX(i, 0)] = 2.0 * u_b - u[IX(i, 1)];
        for (int i = 1; i <= nx; i++) v[IX(i, 1)] = 0.0;
        for (int i = 1; i <= nx; i++) T[IX(i, 0)] = T[IX(i, 1)];
    }
}

/* solve() sub-step loop: rayleigh.py:174-240, mixing.py:147-209.
 * seg_val: rayleigh Th + a_j (a already zero-meaned / scaled, :164-171), wall: mixing lid speeds.
 * iters_out[ndt_act] (nullable) receives itp of every Poisson solve.
 * work: 4*(nx+2)*(ny+2) doubles.  Returns 1 on Poisson overflow (reference calls exit(1)). */
ORC_API int orc_mac_solve(const orc_mac_cfg *c, double *u, double *v, double *p, double *T,
                          double *us, double *vs, double *phi, double Tc, const double *seg_val,
                          int n_sgts, int nx_sgts, const double *wall, int32_t *iters_out,
                          double *work)
{
    const size_t n = (size_t)(c->nx + 2) * (size_t)(c->ny + 2);
    for (int it = 0; it < c->ndt_act; it++) {
        mac_bcs(c, u, v, T, Tc, seg_val, n_sgts, nx_sgts, wall);
        mac_predictor(c, u, v, us, vs, p, T);
        int ovf;
        int itp = mac_poisson(c, us, vs, phi, work, &ovf);
        if (iters_out) iters_out[it] = itp;
        for (size_t k = 0; k < n; k++) p[k] += phi[k];          /* rayleigh.py:219 */
        if (ovf) return 1;
        mac_corrector(c, u, v, us, vs, phi);
        mac_transport(c, u, v, T);
    }
    return 0;
}

/* Batched CPU-baseline driver: state arrays [B, (nx+2)(ny+2)]; per-env seg_val [B,n_sgts]
 * (rayleigh) or wall [B,4] (mixing); iters_total[B] = sum of itp over the action. */
typedef struct {
    const orc_mac_cfg *c; double *u, *v, *p, *T; double Tc; const double *seg_val; int n_sgts, nx_sgts;
    const double *wall; int64_t *iters_total;
} mac_batch_ctx;

static void mac_batch_body(int b, void *vctx, void *tls)
{
    mac_batch_ctx *k = (mac_batch_ctx *)vctx;
    const orc_mac_cfg *c = k->c;
    const size_t n = (size_t)(c->nx + 2) * (size_t)(c->ny + 2);
    double *work = (double *)tls;
    double *us = work + 4 * n, *vs = work + 5 * n, *phi = work + 6 * n;
    int32_t *it = (int32_t *)(work + 7 * n);
    memset(us, 0, sizeof(double) * 3 * n);
    orc_mac_solve(c, k->u + b * n, k->v + b * n, k->p + b * n, k->T + b * n, us, vs, phi, k->Tc,
                  k->seg_val ? k->seg_val + (size_t)b * k->n_sgts : NULL, k->n_sgts, k->nx_sgts,
                  k->wall ? k->wall + (size_t)b * 4 : NULL, it, work);
    int64_t s = 0;
    for (int i = 0; i < c->ndt_act; i++) s += it[i];
    if (k->iters_total) k->iters_total[b] = s;
}

ORC_API void orc_mac_solve_batch(const orc_mac_cfg *c, int B, double *u, double *v, double *p,
                                 double *T, double Tc, const double *seg_val, int n_sgts,
                                 int nx_sgts, const double *wall, int64_t *iters_total)
{
    const size_t n = (size_t)(c->nx + 2) * (size_t)(c->ny + 2);
    mac_batch_ctx k = {c, u, v, p, T, Tc, seg_val, n_sgts, nx_sgts, wall, iters_total};
    orc_parallel_for(B, mac_batch_body, &k, sizeof(double) * n * 7 + sizeof(int32_t) * (c->ndt_act + 2));
}
