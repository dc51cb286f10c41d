// controller_rollout.cuh — the reference's classical comparison controllers as per-env device functions, fused into a
// K-step rollout of an indirect-control handle (action = [F, Mx, My, Mz]), one thread per env, state in registers:
//   * LQR  environment/controller/lqr_quad.py:129-157      (gains K_t 3x6, K_att 4x6 from the two AREs, :25-111)
//   * PID  environment/controller/pid_vel_control.py:29-127 (cascaded velocity -> attitude PIDs, class pid :113-127)
// so that the controller comparison of the reference's README (LQR / PID / PPO on the same initial states) runs on
// millions of envs without a host round trip per step.  Included by quadsim.cu (single translation unit).
#pragma once

template <typename R> struct CtrlDev {
    int32_t kind;                 // QS_CTRL_LQR | QS_CTRL_PID
    R k_t[3][6], k_att[4][6];
    R p[6], i[6], d[6];           // the six scalar PIDs: x, y, z velocity loops, phi, theta, psi attitude loops
    R xd[3], psd;                 // velocity / yaw set-points (pid_vel_control.py:150-153)
    R pid_ts;                     // class pid's own time step (default argument 0.01, :114)
    R mass, g, dt;
    R j[3];
};

struct ControlIO {
    int32_t horizon;
    void* ctrl_state;             // [QS_CTRL_STATE_DIM][N] in/out or NULL
    void* obs_out;                // [K][14][N]
    void* action_out;             // [K][4][N]   the [F, Mx, My, Mz] command applied at step t
    void* reward_out;             // [K][N]
    uint8_t* done_out;            // [K][N]
    void* aux_out;                // [K][10][N]  ang(3), ang_vel(3), step_effort(4) — the columns of the reference's logs after vel
    const void* target_traj;      // [K][3] or NULL: velocity set-point in force at step t (mission.velocity), PID law only
};

template <typename R> struct CtrlMem {
    R ang_vel[3];                 // quad.ang_vel (:492) as of the last step — the LQR's rate feedback
    R x_old[6], ix[6];            // class pid memory
    R ang_d_ant[3];               // pid_control.ang_d_ant
    R pending[4];                 // PID: action computed after the previous step, applied at the next (pid_vel_control.py:144-153)
};

template <typename R> __device__ __forceinline__ void ctrl_fresh(const CtrlDev<R>& cd, CtrlMem<R>& m) {
#pragma unroll
    for (int k = 0; k < 6; ++k) { m.x_old[k] = R(0); m.ix[k] = R(0); }
#pragma unroll
    for (int k = 0; k < 3; ++k) m.ang_d_ant[k] = R(0);
    m.pending[0] = cd.g * cd.mass;                                   // action = np.array([9.82*1.03, 0, 0, 0])  :144
    m.pending[1] = m.pending[2] = m.pending[3] = R(0);
}

// lqr_quad.py:129-157 (deuler_t, :148, is computed by the script and never used)
template <typename R>
__device__ __forceinline__ void lqr_law(const CtrlDev<R>& cd, const R y[13], const R ang[3], const R ang_vel[3], R a[4]) {
    const R st[6] = {R(0), y[1], R(0), y[3], R(0), y[5]};            // :129
    R F[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        R s = R(0);
#pragma unroll
        for (int c = 0; c < 6; ++c) s += cd.k_t[r][c] * st[c];       // :130
        F[r] = s;
    }
    const R theta_t = M_<R>::atan2(F[0], F[2] + cd.g);               // :135
    R st_, ct_, sp_, cp_;
    M_<R>::sincos(theta_t, &st_, &ct_);
    const R phi_t = M_<R>::atan2(-F[1] * ct_, F[2] + cd.g);          // :137
    M_<R>::sincos(phi_t, &sp_, &cp_);
    const R U_1 = cd.mass * (F[2] + cd.g) / (ct_ * cp_);             // :141
    const R e0 = ang[0] - phi_t, e1 = ang[1] - theta_t, e2 = ang[2]; // :144
    const R sa[6] = {e0, ang_vel[0], e1, ang_vel[1], e2, ang_vel[2]};   // :152
#pragma unroll
    for (int r = 1; r < 4; ++r) {
        R s = R(0);
#pragma unroll
        for (int c = 0; c < 6; ++c) s += cd.k_att[r][c] * sa[c];     // :153
        a[r] = s;
    }
    a[0] = U_1;                                                      // :154
}

// class pid :113-127 — the dx argument of pid.pid is overwritten by the finite difference of x
template <typename R>
__device__ __forceinline__ R pid_scalar(const CtrlDev<R>& cd, CtrlMem<R>& m, int k, R x, R x_d, R dx_d) {
    const R dx = (x - m.x_old[k]) / cd.pid_ts;
    m.x_old[k] = x;
    m.ix[k] = m.ix[k] + (x_d - x) * cd.pid_ts;
    return cd.p[k] * (x_d - x) + cd.d[k] * (dx_d - dx) - cd.i[k] * m.ix[k];
}

// pid_control.control :97-110 = lower_control :48-63 + upper_control :66-95 (3x3 inverse in closed form)
template <typename R>
__device__ __forceinline__ void pid_law(const CtrlDev<R>& cd, CtrlMem<R>& m, const R y[13], const R ang[3], R a[4],
                                        const R xd[3]) {              // xd: the script's constant set-point (:150) or the mission's
    const R u_1 = pid_scalar(cd, m, 0, y[1], xd[0], R(0));
    const R u_2 = pid_scalar(cd, m, 1, y[3], xd[1], R(0));
    const R u_3 = pid_scalar(cd, m, 2, y[5], xd[2], R(0));
    const R theta_d = M_<R>::atan2(u_1, u_3 + cd.g);
    R std_, ctd_, spd_, cpd_;
    M_<R>::sincos(theta_d, &std_, &ctd_);
    const R phi_d = M_<R>::atan2(-u_2 * ctd_, u_3 + cd.g);
    M_<R>::sincos(phi_d, &spd_, &cpd_);
    a[0] = cd.mass * (u_3 + cd.g) / (ctd_ * cpd_);
    const R ang_d[3] = {phi_d, theta_d, cd.psd};                     // :103
    R v_ang_d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { v_ang_d[k] = (ang_d[k] - m.ang_d_ant[k]) / cd.dt; m.ang_d_ant[k] = ang_d[k]; }   // :104,:106
    const R u_5 = pid_scalar(cd, m, 3, ang[0], ang_d[0], v_ang_d[0]);
    const R u_6 = pid_scalar(cd, m, 4, ang[1], ang_d[1], v_ang_d[1]);
    const R u_7 = pid_scalar(cd, m, 5, ang[2], ang_d[2], v_ang_d[2]);
    R sp, cp, st, ct;
    M_<R>::sincos(ang[0], &sp, &cp);
    M_<R>::sincos(ang[1], &st, &ct);
    const R tt = st / ct;
    const R b1 = R(1) / cd.j[0], b2 = tt * sp / cd.j[1], b3 = tt * cp / cd.j[2];       // :81-87
    const R b4 = cp / cd.j[1], b5 = -sp / cd.j[2], b6 = sp / ct / cd.j[1], b7 = cp / ct / cd.j[2];
    // [U2,U3,U4] = inv([[b1,b2,b3],[0,b4,b5],[0,b6,b7]]) @ [u5,u6,u7]   :89-93
    const R det = b4 * b7 - b5 * b6;
    const R U_3 = (b7 * u_6 - b5 * u_7) / det;
    const R U_4 = (b4 * u_7 - b6 * u_6) / det;
    const R U_2 = (u_5 - b2 * U_3 - b3 * U_4) / b1;
    a[1] = U_2; a[2] = U_3; a[3] = U_4;
}

template <typename R, int INTEG, bool ROBUST = false>
__global__ void __launch_bounds__(kBlock)
control_rollout_kernel(const __grid_constant__ DevParams<R> p, const __grid_constant__ SimView<R> v,
                       const __grid_constant__ CtrlDev<R> cd, const __grid_constant__ ControlIO io) {
    LocalStats ls;
    ls.clear();
    bool any_end = false;
    R* cs = (R*)io.ctrl_state;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < v.N; n += stride) {
        Env<R> e;
        load_env(v, n, e);
        CtrlMem<R> m;
        if (cs) {
#pragma unroll
            for (int k = 0; k < 3; ++k) m.ang_vel[k] = cs[(int64_t)k * v.N + n];
#pragma unroll
            for (int k = 0; k < 6; ++k) { m.x_old[k] = cs[(int64_t)(3 + k) * v.N + n]; m.ix[k] = cs[(int64_t)(9 + k) * v.N + n]; }
#pragma unroll
            for (int k = 0; k < 3; ++k) m.ang_d_ant[k] = cs[(int64_t)(15 + k) * v.N + n];
#pragma unroll
            for (int k = 0; k < 4; ++k) m.pending[k] = cs[(int64_t)(18 + k) * v.N + n];
        } else {
            ctrl_fresh(cd, m);
            m.ang_vel[0] = m.ang_vel[1] = m.ang_vel[2] = R(0);
        }
        if (v.ang_vel) {                                             // QS_FLAG_AUX handles carry quad.ang_vel themselves
#pragma unroll
            for (int k = 0; k < 3; ++k) m.ang_vel[k] = v.ang_vel[k * v.ld + n];
        }
        Ctrl<R> c_last;
        StepOut<R> o;
        R reward = R(0);
        bool done = false, solved = false, warm_last = false;
        int32_t gust_count = ROBUST ? v.gust_count[n] : 0;           // robust_control: kept in a register over the horizon
        const RobustCtx rc{v.seed, v.env_id_offset + (uint32_t)n, &gust_count};
        for (int t = 0; t < io.horizon; ++t) {
            R a[4];
            if (cd.kind == QS_CTRL_LQR) {
                lqr_law(cd, e.y, e.prev_ang, m.ang_vel, a);          // env.ang == prev_ang after a step (:493)
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) a[k] = m.pending[k];
            }
            bool warm = false;
            if (p.flags & F_ASYNC_RESET) warm = async_warmup_prologue(p, e, a);
            const bool was_done = (e.flags & EF_DONE) != 0;
            step_core<R, INTEG, false, ROBUST>(p, e, a, o, &c_last, &rc);
            if (warm) o.reward = R(0); else e.ep_return += o.reward;
            reward = o.reward; done = o.done; solved = o.solved; warm_last = warm;
#pragma unroll
            for (int k = 0; k < 3; ++k) m.ang_vel[k] = o.ang_vel[k];
            if (done && !was_done) { count_episode(ls, p, e, o); any_end = true; }
            if (io.aux_out) {                                        // logged BEFORE an asynchronous re-sample, like the scripts' memory_step
                R* xt = (R*)io.aux_out + (int64_t)t * 10 * v.N;
#pragma unroll
                for (int k = 0; k < 3; ++k) { xt[(int64_t)k * v.N + n] = o.ang[k]; xt[(int64_t)(3 + k) * v.N + n] = o.ang_vel[k]; }
#pragma unroll
                for (int k = 0; k < 4; ++k) xt[(int64_t)(6 + k) * v.N + n] = o.effort[k];
            }
            if (io.obs_out) {
                R* ot = (R*)io.obs_out + (int64_t)t * 14 * v.N;
#pragma unroll
                for (int k = 0; k < 10; ++k) ot[k * v.N + n] = e.y[k];
#pragma unroll
                for (int k = 0; k < 4; ++k) ot[(10 + k) * v.N + n] = o.vq[k];
            }
            if ((p.flags & F_ASYNC_RESET) && done) {
                async_resample(p, v.seed, v.env_id_offset + (uint32_t)n, e, o.vq);
                ctrl_fresh(cd, m);                                   // controller = pid_control(drone) after every reset (:143)
            } else if (warm) {
                ctrl_fresh(cd, m);                                   // ... i.e. after reset's T hover steps: nothing is remembered across them
            } else if (cd.kind == QS_CTRL_PID) {
                // controller.control(...) after the step (:153); with a mission the set-point is the one in force at this step
                R xd[3] = {cd.xd[0], cd.xd[1], cd.xd[2]};
                if (io.target_traj) {
                    const R* tp = (const R*)io.target_traj + 3 * t;
                    xd[0] = tp[0]; xd[1] = tp[1]; xd[2] = tp[2];
                }
                pid_law(cd, m, e.y, e.prev_ang, m.pending, xd);
            }
            if (io.action_out) {
                R* at = (R*)io.action_out + (int64_t)t * 4 * v.N;
#pragma unroll
                for (int k = 0; k < 4; ++k) at[k * v.N + n] = a[k];
            }
            if (io.reward_out) ((R*)io.reward_out)[(int64_t)t * v.N + n] = reward;
            if (io.done_out) io.done_out[(int64_t)t * v.N + n] = (uint8_t)((done ? 1 : 0) | (warm ? 2 : 0));
        }
        store_env(v, n, e, o.vq);
        if (ROBUST) v.gust_count[n] = gust_count;
        v.reward[n] = reward;
        v.done[n] = (uint8_t)((done ? 1 : 0) | (warm_last ? 2 : 0));
        v.solved[n] = solved;
        if (v.ang_vel) store_aux(p, v, n, e, o, c_last);
        if (cs) {
#pragma unroll
            for (int k = 0; k < 3; ++k) cs[(int64_t)k * v.N + n] = m.ang_vel[k];
#pragma unroll
            for (int k = 0; k < 6; ++k) { cs[(int64_t)(3 + k) * v.N + n] = m.x_old[k]; cs[(int64_t)(9 + k) * v.N + n] = m.ix[k]; }
#pragma unroll
            for (int k = 0; k < 3; ++k) cs[(int64_t)(15 + k) * v.N + n] = m.ang_d_ant[k];
#pragma unroll
            for (int k = 0; k < 4; ++k) cs[(int64_t)(18 + k) * v.N + n] = m.pending[k];
        }
    }
    flush_stats(ls, any_end, v.stats);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&v.stats[7], (double)v.N * io.horizon);
}

template <typename R> static CtrlDev<R> make_ctrl_dev(const qs_sim* s, const qs_controller* c) {
    CtrlDev<R> d;
    d.kind = c->kind;
    for (int r = 0; r < 3; ++r) for (int k = 0; k < 6; ++k) d.k_t[r][k] = (R)c->k_t[r][k];
    for (int r = 0; r < 4; ++r) for (int k = 0; k < 6; ++k) d.k_att[r][k] = (R)c->k_att[r][k];
    const double* grp[6] = {c->pid_xy, c->pid_xy, c->pid_z, c->pid_att, c->pid_att, c->pid_psi};
    for (int k = 0; k < 6; ++k) { d.p[k] = (R)grp[k][0]; d.i[k] = (R)grp[k][1]; d.d[k] = (R)grp[k][2]; }
    for (int k = 0; k < 3; ++k) { d.xd[k] = (R)c->target_vel[k]; d.j[k] = (R)s->cfg.params.j[k]; }
    d.psd = (R)c->target_psi;
    d.pid_ts = (R)(c->pid_ts > 0 ? c->pid_ts : 0.01);
    d.mass = (R)s->cfg.params.mass; d.g = (R)s->cfg.params.gravity; d.dt = (R)s->cfg.t_step;
    return d;
}

template <typename R, int INTEG, bool DIRECT>
static void launch_control_rollout(qs_sim* s, const qs_controller* c, const ControlIO& io, cudaStream_t st) {
    if (s->cfg.flags & QS_FLAG_ROBUST)       // robust_control: the same laws against perturbed dynamics and wind gusts
        control_rollout_kernel<R, INTEG, true><<<grid_for(s, s->N), kBlock, 0, st>>>(params_of<R>(s), make_view<R>(s), make_ctrl_dev<R>(s, c), io);
    else
        control_rollout_kernel<R, INTEG><<<grid_for(s, s->N), kBlock, 0, st>>>(params_of<R>(s), make_view<R>(s), make_ctrl_dev<R>(s, c), io);
}

extern "C" int qs_default_controller(qs_controller* c, int kind, int clipped) {
    if (!c || (kind != QS_CTRL_LQR && kind != QS_CTRL_PID)) return fail(QS_EINVAL, "qs_default_controller: bad argument");
    memset(c, 0, sizeof(*c));
    c->kind = kind;
    c->pid_ts = 0.01;
    // pid_vel_control.py:17-27
    const double cl[4][3] = {{1, -0.0, 0}, {0.4, -0.0, 0}, {20, 0, 20}, {5, 0, 5}};
    const double nc[4][3] = {{2, -0.0, 0}, {1, -0.0, 0}, {180, 0, 50}, {40, 0, 20}};
    const double (*g)[3] = clipped ? cl : nc;
    for (int k = 0; k < 3; ++k) { c->pid_xy[k] = g[0][k]; c->pid_z[k] = g[1][k]; c->pid_att[k] = g[2][k]; c->pid_psi[k] = g[3][k]; }
    // LQR gains are the solutions of two AREs (lqr_quad.py:82-111): computed by the host framework (SciPy) and passed in
    return QS_OK;
}

extern "C" int qs_control_rollout(qs_handle h, const qs_controller* c, const qs_control_rollout_args* a, void* stream) {
    if (!h || !c || !a) return fail(QS_EINVAL, "qs_control_rollout: NULL argument");
    if (a->horizon < 1) return fail(QS_EINVAL, "qs_control_rollout: horizon must be >= 1");
    if (c->kind != QS_CTRL_LQR && c->kind != QS_CTRL_PID) return fail(QS_EINVAL, "qs_control_rollout: bad controller kind");
    const uint32_t f = h->cfg.flags;
    if (f & QS_FLAG_DIRECT_CONTROL)
        return fail(QS_ESTATE, "qs_control_rollout: the LQR / PID laws command [F,Mx,My,Mz]: needs a direct_control=0 handle");
    if (f & (QS_FLAG_SENSOR_NOISE | QS_FLAG_AUTO_RESET))
        return fail(QS_ESTATE, "qs_control_rollout: not available with SENSOR_NOISE / strict AUTO_RESET (use ASYNC_RESET)");
    if ((f & QS_FLAG_AUX) && (f & QS_FLAG_ASYNC_RESET))
        return fail(QS_ESTATE, "qs_control_rollout: AUX rows are not maintained across asynchronous resets");
    if (a->target_traj && c->kind != QS_CTRL_PID)
        return fail(QS_EINVAL, "qs_control_rollout: target_traj (velocity set-points) applies to the PID law only");
    QS_USE_DEVICE(h);
    ControlIO io{a->horizon, a->ctrl_state, a->obs_out, a->action_out, a->reward_out, a->done_out, a->aux_out, a->target_traj};
    cudaStream_t st = (cudaStream_t)stream;
    QS_DISPATCH(h, launch_control_rollout, h, c, io, st);
    QS_CUDA(cudaGetLastError());
    return QS_OK;
}
