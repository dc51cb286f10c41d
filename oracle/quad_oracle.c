/* TEST INFRASTRUCTURE ONLY — plain-C (float64) restatement of the reference hot path.
 *
 * Used (a) as the CPU baseline / `--impl reference` arm of bench.py on the GPU box, where the Python
 * reference cannot travel, and (b) by tests/ as a second checker.  The product never links or calls it.
 * Parity status: PINNED through oracle/quad_oracle.py (tests/test_oracle_c.py checks this file against the
 * NumPy restatement, which is itself pinned to the reference's golden vectors and shipped logs).
 *
 * It is a literal scalar port of what the reference executes per env step — including the work the
 * reference repeats in every RHS call (f2F :247-272, the 10-point beam-drag loop :328-334, J^-1 :384) —
 * so that its timing is the reference's ALGORITHM in compiled code, not an optimised re-formulation.
 * Citations: environment/quadrotor_env.py (reference root) and scipy/integrate/_ivp/{rk,common}.py.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* constants :30-80 */
#define BB_POS 5.0
#define BB_VEL 10.0
#define BB_ANG (M_PI / 2)
#define QM 1.03
#define QG 9.82
#define RHO 1.2041
#define C_D 1.1
#define K_F 1.435e-5
#define K_M 2.4086e-7
#define I_R 5e-5
#define T2WR 2.0
static const double JD[3] = {16.83e-3, 16.83e-3, 28.34e-3};
#define ARM 0.26
#define BEAM 0.05
static const double TR[3] = {0.005, 0.01, 0.1};
static const double TR_P[3] = {3, 2, 1};

typedef struct {
    int direct;
    double a[4];     /* direct: clipped normalised thrusts; indirect: [F,Mx,My,Mz] after the mixer */
    double w[4];     /* rotor speeds (indirect mode: from f2w) */
} ctrl_t;

/* f2F :247-272 */
static void f2F(const double a[4], double w[4], double* F, double Mo[3]) {
    double f[4];
    for (int k = 0; k < 4; ++k) { f[k] = (a[k] + 1) * T2WR * QM * QG / 8; w[k] = sqrt(f[k] / K_F); }
    *F = f[0] + f[1] + f[2] + f[3];
    Mo[0] = (f[2] - f[0]) * ARM; Mo[1] = (f[1] - f[3]) * ARM; Mo[2] = (-f[0] + f[1] - f[2] + f[3]) * K_M / K_F;
}

/* f2w :197-245 (closed-form inverse of the 4x4 mixer) */
static void f2w(int clipped, const double fm[4], double eff[4], double w[4], double fmn[4]) {
    double uf = fm[0] / (4 * K_F), ux = fm[1] / (2 * ARM * K_F), uy = fm[2] / (2 * ARM * K_F), uz = fm[3] / (4 * K_M);
    double u[4] = {uf - ux - uz, uf + uy + uz, uf + ux - uz, uf - uy + uz};
    const double umax = T2WR * QM * QG / 4 / K_F;
    for (int k = 0; k < 4; ++k) {
        if (clipped) { if (u[k] < 0) u[k] = 0; if (u[k] > umax) u[k] = umax; w[k] = sqrt(u[k]); }
        else w[k] = sqrt(fabs(u[k])) * (u[k] < 0 ? -1.0 : 1.0);
        eff[k] = (u[k] * K_F / (T2WR * QM * QG / 4) * 2) - 1;
    }
    fmn[0] = K_F * (u[0] + u[1] + u[2] + u[3]);
    fmn[1] = ARM * K_F * (u[2] - u[0]);
    fmn[2] = ARM * K_F * (u[1] - u[3]);
    fmn[3] = K_M * (-u[0] + u[1] - u[2] + u[3]);
}

/* drone_eq :274-406; vq_out (nullable) receives V_q :392 */
static void drone_eq(const ctrl_t* c, const double x[13], double dx[13], double* vq_out) {
    double w[4], F, Mo[3];
    if (c->direct) f2F(c->a, w, &F, Mo);
    else { F = c->a[0]; Mo[0] = c->a[1]; Mo[1] = c->a[2]; Mo[2] = c->a[3]; memcpy(w, c->w, sizeof(w)); }
    double q0 = x[6], q1 = x[7], q2 = x[8], q3 = x[9];
    double nq = sqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
    double a = q0 / nq, b = q1 / nq, cc = q2 / nq, d = q3 / nq;
    double R[3][3] = {{a * a + b * b - cc * cc - d * d, 2 * b * cc - 2 * a * d, 2 * b * d + 2 * a * cc},
                      {2 * b * cc + 2 * a * d, a * a - b * b + cc * cc - d * d, 2 * cc * d - 2 * a * b},
                      {2 * b * d - 2 * a * cc, 2 * cc * d + 2 * a * b, a * a - b * b - cc * cc + d * d}};
    double v[3] = {x[1], x[3], x[5]}, vb[3], fd[3];
    const double area[3] = {BEAM * 2 * ARM, BEAM * 2 * ARM, BEAM * 2 * ARM * 2};
    for (int i = 0; i < 3; ++i) {
        vb[i] = R[0][i] * v[0] + R[1][i] * v[1] + R[2][i] * v[2];
        fd[i] = -0.5 * RHO * C_D * area[i] * (fabs(vb[i]) * vb[i]);
    }
    double W[3] = {x[10], x[11], x[12]}, md[3] = {0, 0, 0};
    for (int k = 0; k < 10; ++k) {                                   /* :328-334 */
        double xx = ARM * k / 9.0;
        md[0] += -RHO * C_D * BEAM * ARM / 10 * (fabs(xx * W[0]) * (xx * W[0])) * xx;
        md[1] += -RHO * C_D * BEAM * ARM / 10 * (fabs(xx * W[1]) * (xx * W[1])) * xx;
        md[2] += -2 * RHO * C_D * BEAM * ARM / 10 * (fabs(xx * W[2]) * (xx * W[2])) * xx;
    }
    double omega_r = (-w[0] + w[1] - w[2] + w[3]) * I_R;              /* :345 */
    double mg[3] = {-W[0] * omega_r, W[1] * omega_r, 0};
    double fb[3] = {fd[0], fd[1], fd[2] + F}, acc[3];
    for (int i = 0; i < 3; ++i) acc[i] = (R[i][0] * fb[0] + R[i][1] * fb[1] + R[i][2] * fb[2]) / QM;
    acc[2] -= QG;
    double JW[3] = {JD[0] * W[0], JD[1] * W[1], JD[2] * W[2]};
    double cr[3] = {W[1] * JW[2] - W[2] * JW[1], W[2] * JW[0] - W[0] * JW[2], W[0] * JW[1] - W[1] * JW[0]};
    double vq[4] = {0.5 * (-W[0] * b - W[1] * cc - W[2] * d), 0.5 * (W[0] * a + W[2] * cc - W[1] * d),
                    0.5 * (W[1] * a - W[2] * b + W[0] * d), 0.5 * (W[2] * a + W[1] * b - W[0] * cc)};
    dx[0] = v[0]; dx[1] = acc[0]; dx[2] = v[1]; dx[3] = acc[1]; dx[4] = v[2]; dx[5] = acc[2];
    for (int i = 0; i < 4; ++i) dx[6 + i] = vq[i];
    for (int i = 0; i < 3; ++i) dx[10 + i] = (Mo[i] + mg[i] + md[i] - cr[i]) * (1.0 / JD[i]);
    if (vq_out) memcpy(vq_out, vq, sizeof(vq));
}

static double rms13(const double* v) {
    double s = 0;
    for (int j = 0; j < 13; ++j) s += v[j] * v[j];
    return sqrt(s) / sqrt(13.0);
}

/* solve_ivp(drone_eq,(0,tb),y) with defaults: rk.py:85-105,111-183,14-70; common.py:68-134 */
static int rk45(const ctrl_t* c, double y[13], double tb) {
    static const double A[6][5] = {{0}, {1.0 / 5}, {3.0 / 40, 9.0 / 40}, {44.0 / 45, -56.0 / 15, 32.0 / 9},
                                   {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729},
                                   {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
    static const double B[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
    static const double E[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};
    const double rtol = 1e-3, atol = 1e-6;
    double K[7][13], f[13], tmp[13], sc[13], yn[13];
    double t = 0;
    int nfev = 1;
    drone_eq(c, y, f, 0);
    for (int j = 0; j < 13; ++j) { sc[j] = atol + fabs(y[j]) * rtol; tmp[j] = y[j] / sc[j]; }
    double d0 = rms13(tmp);
    for (int j = 0; j < 13; ++j) tmp[j] = f[j] / sc[j];
    double d1 = rms13(tmp);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    if (h0 > tb) h0 = tb;
    for (int j = 0; j < 13; ++j) tmp[j] = y[j] + h0 * f[j];
    drone_eq(c, tmp, K[1], 0); ++nfev;
    for (int j = 0; j < 13; ++j) tmp[j] = (K[1][j] - f[j]) / sc[j];
    double d2 = rms13(tmp) / h0, h1;
    if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
    else h1 = pow(0.01 / fmax(d1, d2), 1.0 / 5);
    double h_abs = fmin(fmin(100 * h0, h1), tb);
    for (int guard = 0; guard < 100000; ++guard) {
        double min_step = 10 * fabs(nextafter(t, INFINITY) - t);
        if (h_abs < min_step) h_abs = min_step;
        int rejected = 0, accepted = 0, failed = 0;
        double t_new = t;
        while (!accepted) {
            if (h_abs < min_step) { failed = 1; break; }
            t_new = t + h_abs;
            if (t_new - tb > 0) t_new = tb;
            double h = t_new - t;
            h_abs = fabs(h);
            memcpy(K[0], f, sizeof(f));
            for (int s = 1; s < 6; ++s) {
                for (int j = 0; j < 13; ++j) {
                    double dy = 0;
                    for (int m = 0; m < s; ++m) dy += K[m][j] * A[s][m];
                    tmp[j] = y[j] + dy * h;
                }
                drone_eq(c, tmp, K[s], 0);
            }
            for (int j = 0; j < 13; ++j) {
                double dy = 0;
                for (int m = 0; m < 6; ++m) dy += K[m][j] * B[m];
                yn[j] = y[j] + h * dy;
            }
            drone_eq(c, yn, K[6], 0);
            nfev += 6;
            for (int j = 0; j < 13; ++j) {
                double e = 0;
                for (int m = 0; m < 7; ++m) e += K[m][j] * E[m];
                tmp[j] = e * h / (atol + fmax(fabs(y[j]), fabs(yn[j])) * rtol);
            }
            double err = rms13(tmp);
            if (err < 1) {
                double factor = err == 0 ? 10 : fmin(10, 0.9 * pow(err, -0.2));
                if (rejected) factor = fmin(1, factor);
                h_abs *= factor; accepted = 1;
            } else if (err != err) { accepted = 1; failed = 1; t_new = tb; }
            else { h_abs *= fmax(0.2, 0.9 * pow(err, -0.2)); rejected = 1; }
        }
        if (failed && !accepted) break;
        t = t_new;
        memcpy(y, yn, sizeof(yn)); memcpy(f, K[6], sizeof(f));
        if (failed || t - tb >= 0) break;
    }
    return nfev;
}

static void rk4(const ctrl_t* c, double y[13], double tb, int substeps) {
    double h = tb / substeps, k1[13], k2[13], k3[13], k4[13], yt[13];
    for (int s = 0; s < substeps; ++s) {
        drone_eq(c, y, k1, 0);
        for (int j = 0; j < 13; ++j) yt[j] = y[j] + 0.5 * h * k1[j];
        drone_eq(c, yt, k2, 0);
        for (int j = 0; j < 13; ++j) yt[j] = y[j] + 0.5 * h * k2[j];
        drone_eq(c, yt, k3, 0);
        for (int j = 0; j < 13; ++j) yt[j] = y[j] + h * k3[j];
        drone_eq(c, yt, k4, 0);
        for (int j = 0; j < 13; ++j) y[j] += (h / 6.0) * (k1[j] + 2 * k2[j] + 2 * k3[j] + k4[j]);
    }
}

typedef struct {
    int64_t n;
    int integrator, substeps, direct, clipped, training, n_limit;
    double dt;
    double* state;        /* [n][13] */
    double* prev_ang;     /* [n][3]  */
    double* prev_shaping; /* [n]     */
    uint8_t* flags;       /* [n] bit0 done, bit1 has_prev_shaping, bit2 solved */
    int32_t* step_i;      /* [n]     */
    double* abs_sum;      /* [n]     */
} qo_envs;

/* quad.step :458-498 (+ done_condition :500-509, reward_function :511-573, control_effort :575-577) for env j */
static void step_one(const qo_envs* e, int64_t j, const double* act_in, double* obs, double* reward, uint8_t* done_out,
                     int32_t* nfev_out) {
    double* y = e->state + 13 * j;
    ctrl_t c;
    double act[4], eff[4];
    e->step_i[j] += 1;
    c.direct = e->direct;
    if (e->direct) {
        for (int k = 0; k < 4; ++k) { double v = act_in[k]; v = v < -1 ? -1 : v; v = v > 1 ? 1 : v; act[k] = v; eff[k] = v; c.a[k] = v; }
    } else {
        for (int k = 0; k < 4; ++k) act[k] = act_in[k];
        f2w(e->clipped, act_in, eff, c.w, c.a);
    }
    int nfev = 0;
    if (e->integrator == 1) nfev = rk45(&c, y, e->dt); else rk4(&c, y, e->dt, e->substeps);
    double dx[13], vq[4];
    drone_eq(&c, y, dx, vq);
    for (int k = 0; k < 10; ++k) obs[k] = y[k];
    for (int k = 0; k < 4; ++k) obs[10 + k] = vq[k];
    double nq = sqrt(y[6] * y[6] + y[7] * y[7] + y[8] * y[8] + y[9] * y[9]);
    double q0 = y[6] / nq, q1 = y[7] / nq, q2 = y[8] / nq, q3 = y[9] / nq;
    double ang[3] = {atan2(2 * (q0 * q1 + q2 * q3), 1 - 2 * (q1 * q1 + q2 * q2)), asin(2 * (q0 * q2 - q3 * q1)),
                     atan2(2 * (q0 * q3 + q1 * q2), 1 - 2 * (q2 * q2 + q3 * q3))};
    for (int k = 0; k < 3; ++k) e->prev_ang[3 * j + k] = ang[k];
    int done = e->flags[j] & 1;
    const double bb[9] = {BB_VEL, BB_VEL, BB_VEL, BB_ANG, BB_ANG, 3.0 / 4 * M_PI, BB_VEL * 2, BB_VEL * 2, BB_VEL * 2};
    const double cx[9] = {y[1], y[3], y[5], ang[0], ang[1], ang[2], y[10], y[11], y[12]};
    for (int k = 0; k < 9; ++k) if (fabs(cx[k]) >= bb[k]) done = 1;
    double v2 = y[1] * y[1] + y[3] * y[3] + y[5] * y[5], e2 = ang[0] * ang[0] + ang[1] * ang[1], psi = ang[2];
    double shaping = -5.0 / 20.0 * (15 * sqrt(v2) / BB_VEL + 4 * fabs(psi / 4) + 1 * sqrt(e2) / BB_ANG);
    double nr = sqrt(v2 + psi * psi), ne = sqrt(e2);
    for (int k = 0; k < 3; ++k) {
        if (nr < sqrt(4 * TR[k] * TR[k])) {
            shaping += TR_P[k];
            if (ne < sqrt(2 * (TR[k] * 4) * (TR[k] * 4))) shaping += TR_P[k];
            break;
        }
    }
    double r = (e->flags[j] & 2) ? shaping - e->prev_shaping[j] : 0.0;
    e->prev_shaping[j] = shaping;
    double pen = 0;
    for (int k = 0; k < 4; ++k) {
        double zc = e->direct ? (2 / T2WR - 1) : (k == 0 ? QM * QG : 0.0);
        pen += (act[k] - zc) * (act[k] - zc);
    }
    r += -pen * 0.003;
    double cur = v2 + e2 + psi * psi + y[10] * y[10] + y[11] * y[11] + y[12] * y[12];
    int solved = (e->flags[j] >> 2) & 1;
    if (cur < 9 * (TR[0] * TR[0])) { r += 20; solved = 1; if (e->training) done = 1; }
    else if (e->step_i[j] >= e->n_limit) { solved = 0; done = 1; }
    else if (done) { r += -20; solved = 0; }
    e->flags[j] = (uint8_t)((done ? 1 : 0) | 2 | (solved ? 4 : 0));
    e->abs_sum[j] += sqrt(eff[0] * eff[0] + eff[1] * eff[1] + eff[2] * eff[2] + eff[3] * eff[3]);
    *reward = r; *done_out = (uint8_t)done;
    if (nfev_out) *nfev_out = nfev;
}

/* One lock-step env step for n envs.  action [n][4], obs [n][14], reward [n], done [n], nfev [n] or NULL. */
void qo_step_batch(const qo_envs* e, const double* action, double* obs, double* reward, uint8_t* done, int32_t* nfev,
                   int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
#endif
    for (int64_t j = 0; j < e->n; ++j)
        step_one(e, j, action + 4 * j, obs + 14 * j, reward + j, done + j, nfev ? nfev + j : 0);
}

int qo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
