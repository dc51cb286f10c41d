"""TEST INFRASTRUCTURE ONLY — CPU (NumPy, float64) restatement of the reference hot path.

This file is the *checker* for the CUDA path.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The product package never does.

Parity status: PINNED.  The restatement is checked (tests/test_oracle_*.py) against
  * the unmodified reference imported in the build container (``oracle/ref_import.py``),
  * fixtures generated from the reference by ``oracle/gen_golden.py`` (tests/golden/*.npz),
  * the reference's own shipped trajectory logs (``environment/controller/
    classical_controller_results/{lqr,pid}_log_same_start.npy``; a slice is committed
    under tests/golden/).

Every function is batched over a leading env axis N and cites the reference lines it
follows (paths relative to the reference root; ``scipy/`` = the SciPy the reference
calls for its integrator, algorithm unchanged since the pinned scipy==1.6.0).
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# constants — environment/quadrotor_env.py:30-80
# --------------------------------------------------------------------------------------
BB_POS = 5
BB_VEL = 10
BB_ANG = np.pi / 2
M, G = 1.03, 9.82
RHO = 1.2041
C_D = 1.1
K_F = 1.435e-5
K_M = 2.4086e-7
I_R = 5e-5
T2WR = 2
J_DIAG = np.array([16.83e-3, 16.83e-3, 28.34e-3])
D = 0.26
BEAM_THICKNESS = 0.05
A_X = BEAM_THICKNESS * 2 * D
A_Y = BEAM_THICKNESS * 2 * D
A_Z = BEAM_THICKNESS * 2 * D * 2
AREA = np.array([A_X, A_Y, A_Z])
SOLVED_REWARD = 20
BROKEN_REWARD = -20
SHAPING_WEIGHT = 5
SHAPING_INTERNAL_WEIGHTS = [15, 4, 1]
P_C = 0.003
TR = [0.005, 0.01, 0.1]
TR_P = [3, 2, 1]

# quad.__init__ :139-143
BB_COND = np.array([BB_VEL, BB_VEL, BB_VEL, BB_ANG, BB_ANG, 3 / 4 * np.pi,
                    BB_VEL * 2, BB_VEL * 2, BB_VEL * 2], dtype=np.float64)
# quad.__init__ :178-180
D_XX = np.linspace(0, D, 10)

# Dormand–Prince 5(4) tableau — scipy/integrate/_ivp/rk.py (class RK45: C, A, B, E)
RK45_C = np.array([0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1])
RK45_A = np.array([
    [0, 0, 0, 0, 0],
    [1 / 5, 0, 0, 0, 0],
    [3 / 40, 9 / 40, 0, 0, 0],
    [44 / 45, -56 / 15, 32 / 9, 0, 0],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729, 0],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
])
RK45_B = np.array([35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84])
RK45_E = np.array([-71 / 57600, 0, 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40])
RK_SAFETY, RK_MIN_FACTOR, RK_MAX_FACTOR = 0.9, 0.2, 10.0
RK45_RTOL, RK45_ATOL = 1e-3, 1e-6  # solve_ivp defaults, scipy/integrate/_ivp/ivp.py (solve_ivp signature)


# --------------------------------------------------------------------------------------
# environment/quaternion_euler_utility.py
# --------------------------------------------------------------------------------------
def euler_quat(ang):
    """(N,3) 3-2-1 Euler -> (N,4) unit quaternion. quaternion_euler_utility.py:17-36."""
    ang = np.asarray(ang, dtype=np.float64)
    phi, theta, psi = ang[..., 0], ang[..., 1], ang[..., 2]
    cp, sp = np.cos(phi / 2), np.sin(phi / 2)
    ct, st = np.cos(theta / 2), np.sin(theta / 2)
    cps, sps = np.cos(psi / 2), np.sin(psi / 2)
    q = np.stack([cp * ct * cps + sp * st * sps,
                  sp * ct * cps - cp * st * sps,
                  cp * st * cps + sp * ct * sps,
                  cp * ct * sps - sp * st * cps], axis=-1)
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def quat_euler(q):
    """(N,4) quaternion -> (N,3) Euler. quaternion_euler_utility.py:39-48 (no asin clamp: NaN propagates)."""
    q = np.asarray(q, dtype=np.float64)
    q0, q1, q2, q3 = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    with np.errstate(invalid="ignore"):
        phi = np.arctan2(2 * (q0 * q1 + q2 * q3), 1 - 2 * (q1 ** 2 + q2 ** 2))
        theta = np.arcsin(2 * (q0 * q2 - q3 * q1))
        psi = np.arctan2(2 * (q0 * q3 + q1 * q2), 1 - 2 * (q2 ** 2 + q3 ** 2))
    return np.stack([phi, theta, psi], axis=-1)


def deriv_quat(w, q):
    """(N,3),(N,4) -> (N,4) quaternion derivative 1/2*Omega(w)*q. quaternion_euler_utility.py:58-69."""
    wx, wy, wz = w[..., 0], w[..., 1], w[..., 2]
    q0, q1, q2, q3 = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    z = np.zeros_like(wx)
    # rows of omega (:63-66) dotted with q
    d0 = z * q0 + (-wx) * q1 + (-wy) * q2 + (-wz) * q3
    d1 = wx * q0 + z * q1 + wz * q2 + (-wy) * q3
    d2 = wy * q0 + (-wz) * q1 + z * q2 + wx * q3
    d3 = wz * q0 + wy * q1 + (-wx) * q2 + z * q3
    return 1 / 2 * np.stack([d0, d1, d2, d3], axis=-1)


def quat_rot_mat(q):
    """(N,4) -> (N,3,3). quaternion_euler_utility.py:71-80."""
    a, b, c, d = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3), dtype=np.float64)
    R[..., 0, 0] = a ** 2 + b ** 2 - c ** 2 - d ** 2
    R[..., 0, 1] = 2 * b * c - 2 * a * d
    R[..., 0, 2] = 2 * b * d + 2 * a * c
    R[..., 1, 0] = 2 * b * c + 2 * a * d
    R[..., 1, 1] = a ** 2 - b ** 2 + c ** 2 - d ** 2
    R[..., 1, 2] = 2 * c * d - 2 * a * b
    R[..., 2, 0] = 2 * b * d - 2 * a * c
    R[..., 2, 1] = 2 * c * d + 2 * a * b
    R[..., 2, 2] = a ** 2 - b ** 2 - c ** 2 + d ** 2
    return R


# --------------------------------------------------------------------------------------
# rotor / mixer maps
# --------------------------------------------------------------------------------------
def f2F(f_action, kf=None):
    """Direct mode: normalised rotor commands (N,4) -> w(N,4), F(N), M(N,3). quadrotor_env.py:247-272.
    kf (N,4): robust_control's episode_kf, applied after w (:265-266)."""
    f = (f_action + 1) * T2WR * M * G / 8
    with np.errstate(invalid="ignore"):
        w = np.sqrt(f / K_F)
    if kf is not None:
        f = f - kf * f
    F_new = f[:, 0] + f[:, 1] + f[:, 2] + f[:, 3]
    M_new = np.stack([(f[:, 2] - f[:, 0]) * D,
                      (f[:, 1] - f[:, 3]) * D,
                      (-f[:, 0] + f[:, 1] - f[:, 2] + f[:, 3]) * K_M / K_F], axis=-1)
    return w, F_new, M_new


_MIXER = np.array([[K_F, K_F, K_F, K_F],
                   [-D * K_F, 0, D * K_F, 0],
                   [0, D * K_F, 0, -D * K_F],
                   [-K_M, +K_M, -K_M, +K_M]])


def f2w(f, m, clipped=True, kf=None):
    """Indirect mode mixer: F(N), M(N,3) -> step_effort(N,4), w(N,4), F_new(N), M_new(N,3).
    quadrotor_env.py:197-245 (np.linalg.solve of the 4x4 mixer, clip or signed sqrt, FM_new = x.u).
    kf (N,4): robust_control's episode_kf, applied to u after w and before FM_new / step_effort (:235-236)."""
    y = np.concatenate([np.asarray(f, dtype=np.float64)[:, None], np.asarray(m, dtype=np.float64)], axis=1)
    u = np.linalg.solve(_MIXER, y.T).T
    if clipped:
        u = np.clip(u, 0, T2WR * M * G / 4 / K_F)
        w = np.sqrt(u)
    else:
        w = np.sqrt(np.abs(u)) * np.where(u < 0, -1.0, 1.0)
    if kf is not None:
        u = u - u * kf
    FM_new = u @ _MIXER.T
    step_effort = (u * K_F / (T2WR * M * G / 4) * 2) - 1
    return step_effort, w, FM_new[:, 0], FM_new[:, 1:4]


# --------------------------------------------------------------------------------------
# drone_eq — environment/quadrotor_env.py:274-406
# --------------------------------------------------------------------------------------
def drone_eq(x, F, Mact, w_rotor, want_aux=False, rb=None):
    """RHS of the 13-state ODE, batched.

    rb (robust_control, :84-109; None = off): dict with wind (N,3) added to the inertial velocity the drag model sees
    (:318-320), ir (N,4) = episode_ir (:341-343), m (N) = episode_m (:360-361), J (N,3) = diag(episode_J) (:381-382).

    x (N,13) = [x,vx,y,vy,z,vz,q0..q3,wx,wy,wz]; F (N) body thrust; Mact (N,3) body moments;
    w_rotor (N,4) rotor speeds (for the gyroscopic term, :345).  Returns dx (N,13) and, if
    ``want_aux``, a dict with the side-effect attributes the reference leaves behind
    (V_q :392, accel :368, mat_rot :315, accelerometer_read :371).
    """
    vel = x[:, 1:6:2]
    q = x[:, 6:10]
    W = x[:, 10:13]
    with np.errstate(invalid="ignore", divide="ignore"):
        q = q / np.linalg.norm(q, axis=1, keepdims=True)                       # :311-312
    R = quat_rot_mat(q)                                                       # :315
    v_in = vel if rb is None else vel + rb["wind"]                            # :318-320
    v_body = np.einsum("nji,nj->ni", R, v_in)                                 # :322  R^T v
    f_drag = -0.5 * RHO * C_D * AREA[None, :] * (np.abs(v_body) * v_body)     # :323
    m_drag = np.zeros_like(W)
    for xx in D_XX:                                                           # :328-334
        m_drag[:, 0] += -RHO * C_D * BEAM_THICKNESS * D / 10 * (np.abs(xx * W[:, 0]) * (xx * W[:, 0])) * xx
        m_drag[:, 1] += -RHO * C_D * BEAM_THICKNESS * D / 10 * (np.abs(xx * W[:, 1]) * (xx * W[:, 1])) * xx
        m_drag[:, 2] += -2 * RHO * C_D * BEAM_THICKNESS * D / 10 * (np.abs(xx * W[:, 2]) * (xx * W[:, 2])) * xx
    if rb is None:
        omega_r = (-w_rotor[:, 0] + w_rotor[:, 1] - w_rotor[:, 2] + w_rotor[:, 3]) * I_R   # :345
    else:
        ir = I_R * (1.0 + rb["ir"])                                           # :342
        omega_r = -w_rotor[:, 0] * ir[:, 0] + w_rotor[:, 1] * ir[:, 1] - w_rotor[:, 2] * ir[:, 2] + w_rotor[:, 3] * ir[:, 3]
    m_gyro = np.stack([-W[:, 0] * omega_r, W[:, 1] * omega_r, np.zeros_like(omega_r)], axis=1)  # :347-349
    f_body = f_drag.copy()
    f_body[:, 2] += F                                                         # :352-353
    f_inertial = np.einsum("nij,nj->ni", R, f_body)                           # :357
    quad_m = M if rb is None else (M * (1.0 + rb["m"]))[:, None]              # :360-363
    accel = f_inertial / quad_m                                               # :365-367
    accel[:, 2] -= G
    JW = W * J_DIAG[None, :]
    m_in = Mact + m_gyro + m_drag - np.cross(W, JW)                           # :378
    if rb is None:
        accel_ang = m_in * (1.0 / J_DIAG)[None, :]                            # :384-388 (J diagonal)
    else:
        accel_ang = m_in / (J_DIAG[None, :] + J_DIAG[None, :] * rb["J"])      # :382  inv(J + J*episode_J)
    V_q = deriv_quat(W, q)                                                    # :392
    out = np.empty_like(x)
    out[:, 0:6:2] = vel
    out[:, 1:6:2] = accel
    out[:, 6:10] = V_q
    out[:, 10:13] = accel_ang
    if want_aux:
        acc_read = np.einsum("nji,nj->ni", R, accel + np.array([0, 0, -G])[None, :])   # :371
        return out, dict(V_q=V_q, accel=accel, mat_rot=R, accelerometer_read=acc_read)
    return out


# --------------------------------------------------------------------------------------
# integrators
# --------------------------------------------------------------------------------------
def _rms(x):
    """scipy/integrate/_ivp/common.py:63-65  norm(x) = ||x||_2 / sqrt(n)."""
    return np.linalg.norm(x, axis=1) / x.shape[1] ** 0.5


def rk45_solve(fun, y0, t_bound, rtol=RK45_RTOL, atol=RK45_ATOL, max_attempts=10000):
    """Per-env replica of ``solve_ivp(fun,(0,t_bound),y0)`` with all defaults (RK45), batched with masks.

    Follows scipy/integrate/_ivp/rk.py:85-105 (RungeKutta.__init__), common.py:68-134
    (select_initial_step), rk.py:111-183 (_step_impl), rk.py:14-70 (rk_step), base.py:179-210 (step).
    ``fun(y, idx)`` evaluates the RHS for env subset ``idx``.
    Returns y_final (N,13), nfev (N), n_accepted (N).
    """
    N, n = y0.shape
    all_idx = np.arange(N)
    t = np.zeros(N)
    y = y0.astype(np.float64).copy()
    f = fun(y, all_idx)
    nfev = np.ones(N, dtype=np.int64)
    nacc = np.zeros(N, dtype=np.int64)
    # --- select_initial_step
    with np.errstate(all="ignore"):
        scale = atol + np.abs(y) * rtol
        d0 = _rms(y / scale)
        d1 = _rms(f / scale)
        h0 = np.where((d0 < 1e-5) | (d1 < 1e-5), 1e-6, 0.01 * d0 / d1)
        h0 = np.minimum(h0, t_bound)
        y1 = y + h0[:, None] * f
        f1 = fun(y1, all_idx)
        nfev += 1
        d2 = _rms((f1 - f) / scale) / h0
        h1 = np.where((d1 <= 1e-15) & (d2 <= 1e-15), np.maximum(1e-6, h0 * 1e-3),
                      (0.01 / np.maximum(d1, d2)) ** (1 / 5))
        h_abs = np.minimum(np.minimum(100 * h0, h1), t_bound)
    running = np.ones(N, dtype=bool)
    new_step = np.ones(N, dtype=bool)          # entering _step_impl afresh
    rejected = np.zeros(N, dtype=bool)
    min_step = np.zeros(N)
    K = np.zeros((N, 7, n))
    for _ in range(max_attempts):
        idx = np.nonzero(running)[0]
        if idx.size == 0:
            break
        with np.errstate(all="ignore"):
            # _step_impl prologue (only when a new solver.step() begins)
            ns = idx[new_step[idx]]
            if ns.size:
                min_step[ns] = 10 * np.abs(np.nextafter(t[ns], np.inf) - t[ns])
                h_abs[ns] = np.where(h_abs[ns] < min_step[ns], min_step[ns], h_abs[ns])
                rejected[ns] = False
                new_step[ns] = False
            # too-small step -> solver fails, solve_ivp stops with the last accepted y
            fail = idx[h_abs[idx] < min_step[idx]]
            if fail.size:
                running[fail] = False
                idx = np.nonzero(running)[0]
                if idx.size == 0:
                    break
            ti, yi, fi = t[idx], y[idx], f[idx]
            t_new = ti + h_abs[idx]
            t_new = np.where(t_new - t_bound > 0, t_bound, t_new)
            h = t_new - ti
            ha = np.abs(h)
            # rk_step
            Ki = K[idx]
            Ki[:, 0] = fi
            for s in range(1, 6):
                dy = np.einsum("nsk,s->nk", Ki[:, :s], RK45_A[s, :s]) * h[:, None]
                Ki[:, s] = fun(yi + dy, idx)
            y_new = yi + h[:, None] * np.einsum("nsk,s->nk", Ki[:, :6], RK45_B)
            f_new = fun(y_new, idx)
            Ki[:, 6] = f_new
            nfev[idx] += 6
            scale = atol + np.maximum(np.abs(yi), np.abs(y_new)) * rtol
            err = _rms(np.einsum("nsk,s->nk", Ki, RK45_E) * h[:, None] / scale)
            acc = err < 1
            # NaN error norm: `error_norm < 1` is False -> rejection path with NaN factor; the
            # reference then loops until h_abs<min_step is False forever (NaN) -> we stop these envs.
            nanerr = np.isnan(err)
            factor_acc = np.where(err == 0, RK_MAX_FACTOR,
                                  np.minimum(RK_MAX_FACTOR, RK_SAFETY * err ** -0.2))
            factor_acc = np.where(rejected[idx], np.minimum(1.0, factor_acc), factor_acc)
            factor_rej = np.maximum(RK_MIN_FACTOR, RK_SAFETY * err ** -0.2)
            h_abs[idx] = np.where(acc, ha * factor_acc, ha * factor_rej)
            K[idx] = Ki
            a = idx[acc]
            if a.size:
                t[a] = t_new[acc]
                y[a] = y_new[acc]
                f[a] = f_new[acc]
                nacc[a] += 1
                new_step[a] = True
                running[a] = ~(t[a] - t_bound >= 0)
            r = idx[~acc]
            rejected[r] = True
            if nanerr.any():
                # poisoned env: take the NaN state (what the caller would eventually see is undefined)
                bad = idx[nanerr]
                y[bad] = y_new[nanerr]
                running[bad] = False
    return y, nfev, nacc


def rk4_solve(fun, y0, t_bound, substeps=1):
    """Classical fixed-step RK4 with ``substeps`` equal sub-intervals (the FP32 production integrator)."""
    N = y0.shape[0]
    idx = np.arange(N)
    y = y0.astype(np.float64).copy()
    h = t_bound / substeps
    for _ in range(substeps):
        k1 = fun(y, idx)
        k2 = fun(y + 0.5 * h * k1, idx)
        k3 = fun(y + 0.5 * h * k2, idx)
        k4 = fun(y + h * k3, idx)
        y = y + (h / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
    return y


# --------------------------------------------------------------------------------------
# Philox4x32-10 (counter-based RNG used by the CUDA reset / noise sub-passes)
# --------------------------------------------------------------------------------------
_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = np.uint32(0x9E3779B9)
_PHILOX_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(counter, key):
    """counter (N,4) uint32, key (N,2) uint32 -> (N,4) uint32.  Salmon et al. 2011, 10 rounds
    (integer work: the CUDA implementation must match bit-exactly)."""
    c = np.array(counter, dtype=np.uint32).reshape(-1, 4).copy()
    k = np.array(key, dtype=np.uint32).reshape(-1, 2).copy()
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _PHILOX_M0 * c[:, 0].astype(np.uint64)
            p1 = _PHILOX_M1 * c[:, 2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & mask).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & mask).astype(np.uint32)
            c = np.stack([hi1 ^ c[:, 1] ^ k[:, 0], lo1, hi0 ^ c[:, 3] ^ k[:, 1], lo0], axis=1)
            k = np.stack([k[:, 0] + _PHILOX_W0, k[:, 1] + _PHILOX_W1], axis=1)
    return c


def u32_to_unit(u):
    """uint32 -> float64 in (0,1):  (u + 0.5) * 2^-32  (same map as the CUDA side)."""
    return (u.astype(np.float64) + 0.5) * (1.0 / 4294967296.0)


def box_muller(u1, u2):
    r = np.sqrt(-2.0 * np.log(u1))
    return r * np.cos(2 * np.pi * u2), r * np.sin(2 * np.pi * u2)


# stream ids for the Philox counter word 3 (shared with csrc/philox.cuh)
STREAM_RESET = 0
STREAM_SENSOR = 1
STREAM_ACTION = 2
STREAM_POLICY = 3
STREAM_ROBUST = 4
STREAM_GUST = 5


def philox_block(seed, env_id, episode, block, stream):
    """counter = (env_id, episode, block, stream), key = (seed_lo, seed_hi)."""
    env_id = np.asarray(env_id, dtype=np.uint32)
    n = env_id.shape[0]
    ctr = np.stack([env_id,
                    np.broadcast_to(np.asarray(episode, dtype=np.uint32), (n,)),
                    np.broadcast_to(np.asarray(block, dtype=np.uint32), (n,)),
                    np.full(n, stream, dtype=np.uint32)], axis=1)
    key = np.stack([np.full(n, seed & 0xFFFFFFFF, dtype=np.uint32),
                    np.full(n, (seed >> 32) & 0xFFFFFFFF, dtype=np.uint32)], axis=1)
    return philox4x32_10(ctr, key)


def sample_reset_state(seed, env_id, episode):
    """Device-side restatement of the random branch of quad.reset (quadrotor_env.py:439-445) with
    Philox in place of NumPy's MT19937 (which a counter-based generator cannot reproduce):
      ang ~ U(-0.5,0.5)^3; pos ~ clip(N(0,2),+-2.5); vel ~ clip(N(0,2),+-5); w ~ clip(N(0,2),-15,+7.5).
    Draw order (fixed, shared with csrc): block0 -> u(ang0..2), u_spare; blocks 1..3 -> 4 uniforms each ->
    2 Box-Muller pairs each: normals n0..n11; pos=n0..2, vel=n3..5, w=n6..8.
    Returns state (N,13) float64 and ang (N,3).
    """
    b0 = u32_to_unit(philox_block(seed, env_id, episode, 0, STREAM_RESET))
    ang = b0[:, 0:3] - 0.5
    normals = []
    for blk in (1, 2, 3):
        u = u32_to_unit(philox_block(seed, env_id, episode, blk, STREAM_RESET))
        a, b = box_muller(u[:, 0], u[:, 1])
        c, d = box_muller(u[:, 2], u[:, 3])
        normals += [a, b, c, d]
    nrm = np.stack(normals, axis=1)
    st = np.zeros((len(env_id), 13))
    st[:, 0:5:2] = np.clip(nrm[:, 0:3] * 2, -BB_POS / 2, BB_POS / 2)
    st[:, 1:6:2] = np.clip(nrm[:, 3:6] * 2, -BB_VEL / 2, BB_VEL / 2)
    st[:, 6:10] = euler_quat(ang)
    st[:, 10:13] = np.clip(nrm[:, 6:9] * 2, -BB_VEL * 1.5, BB_POS * 1.5)
    return st, ang


# --------------------------------------------------------------------------------------
# robust_control — environment/quadrotor_env.py:84-109 with Philox in place of NumPy's global stream
# (same draw order as csrc/quad_device.cuh: robust_prepare / robust_gust)
# --------------------------------------------------------------------------------------
ROBUST_DEFAULTS = dict(d_kf=0.1, d_m=0.3, d_ir=0.1, d_j=(0.1, 0.1, 0.1), gust_std=(5.0, 5.0, 2.0), gust_period=500)   # :85-93


def robust_episode(seed, env_id, episode, par=ROBUST_DEFAULTS):
    """robust_control.reset (:98-102): episode_kf = U(0,1)^4 D_KF; episode_m = N(0, D_M); episode_ir = U(0,1)^4 D_IR;
    diag(episode_J) = N(0, D_J)^3.  Blocks 0, 1, 2 of stream ROBUST of (env, episode)."""
    a = u32_to_unit(philox_block(seed, env_id, episode, 0, STREAM_ROBUST))
    b = u32_to_unit(philox_block(seed, env_id, episode, 1, STREAM_ROBUST))
    c = u32_to_unit(philox_block(seed, env_id, episode, 2, STREAM_ROBUST))
    n0, n1 = box_muller(c[:, 0], c[:, 1])
    n2, n3 = box_muller(c[:, 2], c[:, 3])
    return dict(kf=a * par["d_kf"], ir=b * par["d_ir"], m=n0 * par["d_m"],
                J=np.stack([n1, n2, n3], axis=1) * np.asarray(par["d_j"])[None, :])


def robust_gust(seed, env_id, count, par=ROBUST_DEFAULTS):
    """Gust number `count` of each env (count <= 0: no wind yet): N(0, gust_std), block 0 of stream GUST of (env, count)."""
    count = np.asarray(count, dtype=np.int64)
    u = u32_to_unit(philox_block(seed, env_id, np.maximum(count, 0), 0, STREAM_GUST))
    n0, n1 = box_muller(u[:, 0], u[:, 1])
    n2, _ = box_muller(u[:, 2], u[:, 3])
    g = np.stack([n0, n1, n2], axis=1) * np.asarray(par["gust_std"])[None, :]
    return np.where((count > 0)[:, None], g, 0.0)


def wind_ramp(last_gust, gust, i, period):
    """robust_control.wind (:104-109) between two gusts: np.linspace(last, gust, P)[(i % P) - 1]  (index -1 = last element)."""
    index = (np.asarray(i, dtype=np.int64) % period) - 1
    t = np.where(index < 0, 1.0, index / (period - 1.0))
    return last_gust + (gust - last_gust) * t[:, None]


# --------------------------------------------------------------------------------------
# the environment: N independent copies of `quad`, advanced in lock-step
# --------------------------------------------------------------------------------------
class BatchQuadOracle:
    """N independent reference ``quad`` environments (quadrotor_env.py:111-577), vectorised.

    integrator: "rk45" = replica of the reference's solve_ivp call (:483); "rk4" = fixed-step RK4.
    """

    def __init__(self, n_envs, t_step, n, training=True, direct_control=1, T=1, clipped=True,
                 integrator="rk45", substeps=1, robust=None):
        """robust: None, or dict(seed=..., env_id=(N,) global ids[, par=ROBUST_DEFAULTS-like]) = quad.robust_control True."""
        self.N = n_envs
        self.robust = robust
        if robust is not None:
            self.episode = np.zeros(n_envs, dtype=np.int64)
            self.gust_count = np.zeros(n_envs, dtype=np.int64)
        self.t_step = t_step
        self.T = T
        self.n = n + T                                                        # :157
        self.training = bool(training)
        self.direct = bool(direct_control)
        self.clipped = clipped
        self.integrator = integrator
        self.substeps = substeps
        self.zero_control = (np.ones(4) * (2 / T2WR - 1)) if self.direct else np.array([M * G, 0, 0, 0])  # :164-167
        self.ang_vel = np.zeros((n_envs, 3))
        self.prev_ang = np.zeros((n_envs, 3))                                 # :171-172 (never cleared by reset)
        self.ang = np.zeros((n_envs, 3))
        self.abs_sum = np.zeros(n_envs)
        self.done = np.ones(n_envs, dtype=bool)                               # :154
        self.solved = np.zeros(n_envs, dtype=np.int64)
        self.i = np.zeros(n_envs, dtype=np.int64)
        self.prev_shaping = np.zeros(n_envs)
        self.has_prev_shaping = np.zeros(n_envs, dtype=bool)
        self.previous_state = np.zeros((n_envs, 13))
        self.state = np.zeros((n_envs, 13))
        self.nfev = np.zeros(n_envs, dtype=np.int64)

    # -- reset :408-454
    def reset(self, det_state, mask=None):
        """Deterministic-state reset for envs in ``mask`` (all if None), then T hover steps.
        Returns obs_hist (T,N,14), act_hist (T,N,4) (rows of non-reset envs are whatever step produced)."""
        m = np.ones(self.N, dtype=bool) if mask is None else np.asarray(mask, dtype=bool)
        self.solved[m] = 0
        self.done[m] = False
        self.i[m] = 0
        self.has_prev_shaping[m] = False
        self.abs_sum[m] = 0
        self.previous_state[m] = np.asarray(det_state, dtype=np.float64)[m]
        self.ang[m] = quat_euler(self.previous_state[m, 6:10])                # :437-438
        obs_h, act_h = [], []
        for _ in range(self.T):
            a = np.broadcast_to(self.zero_control, (self.N, 4)).copy()
            obs, _, _ = self.step(a, mask=m)
            obs_h.append(obs)
            act_h.append(a)
        return np.array(obs_h), np.array(act_h)

    # -- step :458-498
    def step(self, action, mask=None):
        m = np.ones(self.N, dtype=bool) if mask is None else np.asarray(mask, dtype=bool)
        idx_all = np.nonzero(m)[0]
        action = np.asarray(action, dtype=np.float64)
        self.i[m] += 1
        rb = None
        if self.robust is not None:
            par = self.robust.get("par", ROBUST_DEFAULTS)
            seed, ids = self.robust["seed"], np.asarray(self.robust["env_id"])
            rb = robust_episode(seed, ids, self.episode, par)
            new_gust = m & ((self.i % par["gust_period"]) - 1 == 0)               # :105-108
            self.gust_count[new_gust] += 1
            rb["wind"] = wind_ramp(robust_gust(seed, ids, self.gust_count - 1, par), robust_gust(seed, ids, self.gust_count, par),
                                   self.i, par["gust_period"])
            self.rb = rb
        kf = None if rb is None else rb["kf"]
        if self.direct:
            act = np.clip(action, -1, 1)                                      # :470
            self.action = act
            self.clipped_action = act
            step_effort = act
            w, F, Mact = f2F(act, kf)
        else:
            self.action = action
            step_effort, w, F, Mact = f2w(action[:, 0], action[:, 1:4], self.clipped, kf)   # :476
            self.clipped_action = np.concatenate([F[:, None], Mact], axis=1)
        self.w = w

        def sub(g):
            return None if rb is None else {k: v[g] for k, v in rb.items()}

        def fun(y, idx):
            g = idx_all[idx]
            return drone_eq(y, F[g], Mact[g], w[g], rb=sub(g))

        y0 = self.previous_state[m]
        if self.integrator == "rk45":
            y, nfev, _ = rk45_solve(fun, y0, self.t_step)
            self.nfev[m] = nfev
        else:
            y = rk4_solve(fun, y0, self.t_step, self.substeps)
        _, aux = drone_eq(y, F[m], Mact[m], w[m], want_aux=True, rb=sub(m))  # FSAL stage f(t+h,y_new): last drone_eq call
        state = self.state.copy()
        state[m] = y
        self.state = state
        V_q = np.zeros((self.N, 4))
        V_q[m] = aux["V_q"]
        self.V_q = V_q
        self.accel = np.zeros((self.N, 3)); self.accel[m] = aux["accel"]
        self.accelerometer_read = np.zeros((self.N, 3)); self.accelerometer_read[m] = aux["accelerometer_read"]
        self.mat_rot = np.zeros((self.N, 3, 3)); self.mat_rot[m] = aux["mat_rot"]
        self.f_in = F
        quat_state = np.concatenate([self.state[:, 0:10], V_q], axis=1)       # :486
        with np.errstate(invalid="ignore", divide="ignore"):
            q = self.state[:, 6:10] / np.linalg.norm(self.state[:, 6:10], axis=1, keepdims=True)   # :488-489
        ang = quat_euler(q)
        self.ang = np.where(m[:, None], ang, self.ang)
        self.ang_vel = np.where(m[:, None], (self.ang - self.prev_ang) / self.t_step, self.ang_vel)  # :492
        self.prev_ang = np.where(m[:, None], self.ang, self.prev_ang)
        self.previous_state = np.where(m[:, None], self.state, self.previous_state)
        self.step_effort = step_effort
        self._done_condition(m)
        reward = self._reward_function(m)
        self.abs_sum = np.where(m, self.abs_sum + np.sqrt(np.sum(np.square(step_effort), axis=1)), self.abs_sum)  # :575-577
        self.reward = reward
        return quat_state, reward, self.done.copy()

    # -- done_condition :500-509
    def _done_condition(self, m):
        cond_x = np.concatenate([self.state[:, 1:6:2], self.ang, self.state[:, 10:13]], axis=1)
        with np.errstate(invalid="ignore"):
            hit = np.any(np.abs(cond_x) >= BB_COND[None, :], axis=1)
        self.done = np.where(m, self.done | hit, self.done)

    # -- reward_function :511-573
    def _reward_function(self, m):
        velocity = self.state[:, 1:6:2]
        euler = self.ang
        psi = self.ang[:, 2]
        body_ang_vel = self.state[:, 10:13]
        sw = SHAPING_INTERNAL_WEIGHTS
        shaping = -SHAPING_WEIGHT / np.sum(sw) * (sw[0] * np.linalg.norm(velocity / BB_VEL, axis=1)
                                                  + sw[1] * np.abs(psi / 4)
                                                  + sw[2] * np.linalg.norm(euler[:, 0:2] / BB_ANG, axis=1))
        r_state = np.concatenate([velocity, psi[:, None]], axis=1)
        nr = np.linalg.norm(r_state, axis=1)
        ne = np.linalg.norm(euler[:, 0:2], axis=1)
        taken = np.zeros(self.N, dtype=bool)
        with np.errstate(invalid="ignore"):
            for TR_i, TR_Pi in zip(TR, TR_P):                                  # :535-542
                c1 = (nr < np.linalg.norm(np.ones(4) * TR_i)) & ~taken
                c2 = c1 & (ne < np.linalg.norm(np.ones(2) * TR_i * 4))
                shaping = shaping + np.where(c1, TR_Pi, 0) + np.where(c2, TR_Pi, 0)
                taken |= c1
        reward = np.where(self.has_prev_shaping, shaping - self.prev_shaping, 0.0)     # :545-547
        self.prev_shaping = np.where(m, shaping, self.prev_shaping)
        self.has_prev_shaping = self.has_prev_shaping | m
        reward = reward + (-np.sum(np.square(self.action - self.zero_control[None, :]), axis=1) * P_C)   # :553-554
        target_state = 9 * (TR[0] ** 2)                                        # :557
        current_state = np.sum(np.square(np.concatenate([velocity, euler, body_ang_vel], axis=1)), axis=1)
        self.current_state = current_state
        with np.errstate(invalid="ignore"):
            is_solved = current_state < target_state
        timeout = ~is_solved & (self.i >= self.n)
        broken = ~is_solved & ~timeout & self.done
        reward = reward + np.where(is_solved, SOLVED_REWARD, 0.0) + np.where(broken, BROKEN_REWARD, 0.0)
        # :563-573  solved=1 | (timeout or broken) solved=0 | otherwise unchanged
        solved_new = np.where(is_solved, 1, np.where(timeout | broken, 0, self.solved))
        self.solved = np.where(m, solved_new, self.solved)
        new_done = self.done | timeout | (is_solved & self.training)
        self.done = np.where(m, new_done, self.done)
        return np.where(m, reward, 0.0)

    # -- lock-step auto-reset (restatement of the CUDA kernel's predicated reset sub-pass) ---------------
    def step_autoreset(self, action, seed, env_id_offset=0):
        """step(); envs that are done are re-sampled with Philox (episode counter + 1) and warmed up for T hover
        steps, exactly like the in-kernel sub-pass.  Returns (obs after reset, terminal reward, terminal done)."""
        if not hasattr(self, "ep_return"):
            if not hasattr(self, "episode"):
                self.episode = np.zeros(self.N, dtype=np.int64)
            self.ep_return = np.zeros(self.N)
            self.stats = dict(sum_return=0.0, sum_length=0.0, n_episodes=0, n_solved=0, n_broken=0, n_timeout=0,
                              sum_effort=0.0)
        was_done = self.done.copy()
        obs, rew, done = self.step(action)
        self.ep_return += rew
        ended = done & ~was_done
        if ended.any():
            s = self.stats
            s["sum_return"] += float(self.ep_return[ended].sum())
            s["sum_length"] += float((self.i[ended] - self.T).sum())
            s["n_episodes"] += int(ended.sum())
            s["n_solved"] += int(self.solved[ended].sum())
            timeout = ended & (self.solved == 0) & (self.i >= self.n)
            s["n_timeout"] += int(timeout.sum())
            s["n_broken"] += int((ended & (self.solved == 0) & ~timeout).sum())
            s["sum_effort"] += float(self.abs_sum[ended].sum())
        if done.any():
            self.episode[done] += 1
            ids = np.arange(self.N) + env_id_offset
            st, _ = sample_reset_state(seed, ids[done], self.episode[done])
            full = np.zeros((self.N, 13))
            full[done] = st
            oh, _ = self.reset(full, mask=done)
            obs = np.where(done[:, None], oh[-1], obs)
            self.ep_return[done] = 0.0
        return obs, rew, done

    # -- asynchronous warm-up (restatement of QS_FLAG_ASYNC_RESET) ---------------------------------------
    def step_async(self, action, seed, env_id_offset=0):
        """A warm-up step applies zero_control instead of the caller's action and returns reward 0.  An env that
        returns done begins its next episode at the END of that step: Philox-sampled state (episode counter + 1),
        bookkeeping cleared, T warm-up steps owed; the observation returned with done is the new episode's initial
        observation [s0[0:10], V_q(s0)].  Returns (obs, reward, done, warm)."""
        if not hasattr(self, "ep_return"):
            if not hasattr(self, "episode"):
                self.episode = np.zeros(self.N, dtype=np.int64)
            self.ep_return = np.zeros(self.N)
            self.stats = dict(sum_return=0.0, sum_length=0.0, n_episodes=0, n_solved=0, n_broken=0, n_timeout=0,
                              sum_effort=0.0)
        if not hasattr(self, "warm"):
            self.warm = np.zeros(self.N, dtype=np.int64)
        w = self.warm > 0
        a = np.where(w[:, None], self.zero_control[None, :], np.asarray(action, dtype=np.float64))
        self.warm[w] -= 1
        was_done = self.done.copy()
        obs, rew, done = self.step(a)
        rew = np.where(w, 0.0, rew)
        self.ep_return += rew
        ended = done & ~was_done
        if ended.any():
            s = self.stats
            s["sum_return"] += float(self.ep_return[ended].sum())
            s["sum_length"] += float((self.i[ended] - self.T).sum())
            s["n_episodes"] += int(ended.sum())
            s["n_solved"] += int(self.solved[ended].sum())
            timeout = ended & (self.solved == 0) & (self.i >= self.n)
            s["n_timeout"] += int(timeout.sum())
            s["n_broken"] += int((ended & (self.solved == 0) & ~timeout).sum())
            s["sum_effort"] += float(self.abs_sum[ended].sum())
        if done.any():
            self.episode[done] += 1
            ids = np.arange(self.N) + env_id_offset
            st, _ = sample_reset_state(seed, ids[done], self.episode[done])
            self.previous_state[done] = st
            self.state[done] = st
            self.solved[done] = 0
            self.done[done] = False
            self.i[done] = 0
            self.has_prev_shaping[done] = False
            self.abs_sum[done] = 0
            self.ep_return[done] = 0
            self.warm[done] = self.T
            obs = obs.copy()
            obs[done] = np.concatenate([st[:, 0:10], deriv_quat(st[:, 10:13], st[:, 6:10])], axis=1)
        return obs, rew, done, w


# --------------------------------------------------------------------------------------
# sensor model — environment/quadrotor_env.py:579-724 in the canonical call order
# accel_int -> gyro_int -> gyro -> gps -> triad (visual_landing/rl_worker.py:164-175)
# --------------------------------------------------------------------------------------
MAGNET_VEC = np.array([-4047, 12911, -9899]) * 0.01          # :651


def sensor_normals(seed, env_id, episode, step):
    """27 normals per env step; every 32-bit Philox word gives two 16-bit uniforms (one Box-Muller pair).  Mapping shared with
    csrc/sensor_device.cuh: blocks 0..2 of the step give the physical normals P[0..23]; z[0..14] = P[0..14], z[21..26] =
    P[15..20]; the six GPS normals z[15..20] come from block 3 (the device draws them only when the GPS blend reads them)."""
    env_id = np.asarray(env_id)
    P = np.zeros((env_id.shape[0], 32))
    step = np.asarray(step, dtype=np.uint32)
    for b in range(4):
        w = philox_block(seed, env_id, episode, step * np.uint32(4) + np.uint32(b), STREAM_SENSOR)
        for k in range(4):
            u1 = ((w[:, k] & np.uint32(0xFFFF)).astype(np.float64) + 0.5) / 65536.0
            u2 = ((w[:, k] >> np.uint32(16)).astype(np.float64) + 0.5) / 65536.0
            P[:, 8 * b + 2 * k], P[:, 8 * b + 2 * k + 1] = box_muller(u1, u2)
    z = np.zeros_like(P)
    z[:, 0:15] = P[:, 0:15]
    z[:, 21:27] = P[:, 15:21]
    z[:, 15:21] = P[:, 24:30]
    return z


def _unit(a):
    return a / np.linalg.norm(a, axis=-1, keepdims=True)


def _triad_inertial():
    gv = _unit(np.array([0, 0, -G], dtype=np.float64))
    mv = _unit(MAGNET_VEC)
    t2 = _unit(np.cross(gv, mv))
    t3 = _unit(np.cross(gv, t2))
    return np.vstack((gv, t2, t3)).T                           # ti, :691


def _triad(grav_body, mag_body):
    """sensor.triad :664-693 -> R (N,3,3)."""
    g = _unit(grav_body)
    m = _unit(mag_body)
    t2 = _unit(np.cross(g, m))
    t3 = _unit(np.cross(g, t2))
    tb = np.stack((g, t2, t3), axis=2)                         # columns
    return tb @ _triad_inertial().T


def rot_to_quat(R):
    """q returned by sensor.triad (:695-696): Rotation.from_matrix(R.T).as_quat() re-ordered scalar-first.  Restatement of
    SciPy's conversion for an orthogonal matrix (scipy/spatial/transform, `from_matrix`: Markley's method — the largest of
    the three diagonal entries and the trace selects the row the quaternion is built from, then normalise).  R (N,3,3)."""
    m = np.swapaxes(np.asarray(R, dtype=np.float64), 1, 2)
    tr = m[:, 0, 0] + m[:, 1, 1] + m[:, 2, 2]
    choice = np.argmax(np.stack([m[:, 0, 0], m[:, 1, 1], m[:, 2, 2], tr], axis=1), axis=1)
    q = np.zeros((m.shape[0], 4))                                       # scalar-last like SciPy, re-ordered below
    for i in range(3):
        j, k = (i + 1) % 3, (i + 2) % 3
        c = choice == i
        q[c, i] = 1 - tr[c] + 2 * m[c, i, i]
        q[c, j] = m[c, j, i] + m[c, i, j]
        q[c, k] = m[c, k, i] + m[c, i, k]
        q[c, 3] = m[c, k, j] - m[c, j, k]
    c = choice == 3
    q[c, 0] = m[c, 2, 1] - m[c, 1, 2]
    q[c, 1] = m[c, 0, 2] - m[c, 2, 0]
    q[c, 2] = m[c, 1, 0] - m[c, 0, 1]
    q[c, 3] = 1 + tr[c]
    q = q / np.linalg.norm(q, axis=1, keepdims=True)
    return np.concatenate([q[:, 3:4], q[:, 0:3]], axis=1)


class SensorOracle:
    """The reference's `sensor` class (:579-724) for N envs, one method per reference method; every random draw is supplied
    by the caller as STANDARD normals z (the reference's np.random.normal(loc, scale) = loc + scale * z), so that the class
    can be fed the very stream the reference consumed (tests/golden/sensor_vectors.npz) or the Philox stream of the kernels
    (sensor_normals).  Methods update the state of the envs in `mask` only."""

    def __init__(self, n_envs, t_step, accel_std=0.1, accel_bias_drift=0.0005, gyro_std=0.035, gyro_bias_drift=0.00015,
                 magnet_std=15, gps_std_p=1.71, gps_std_v=0.5, gps_blend=0.0):
        self.N, self.dt = n_envs, t_step
        self.gps_p, self.gps_v, self.gps_blend = gps_std_p, gps_std_v, gps_blend      # sensor.__init__ :591; math_trajectory.py:104-105
        self.a_std, self.a_drift, self.g_std, self.g_drift, self.m_std = accel_std, accel_bias_drift, gyro_std, gyro_bias_drift, magnet_std
        self.a_b = np.zeros(n_envs); self.g_b = np.zeros(n_envs)
        self.a_b_d = np.zeros(n_envs); self.g_b_d = np.zeros(n_envs)
        self.vel = np.zeros((n_envs, 3)); self.pos = np.zeros((n_envs, 3)); self.quat = np.zeros((n_envs, 4))
        self.acc0 = np.zeros((n_envs, 3))
        self.R = np.tile(np.eye(3), (n_envs, 1, 1))                                   # :597

    def _m(self, mask):
        return np.ones(self.N, bool) if mask is None else np.asarray(mask, bool)

    @property
    def Rc2(self):
        return self.R[:, :, 2]

    def reset_u(self, u, y, mask=None, reset_R=False):
        """sensor.reset :630-640 + bias_reset :600-608 with the U(0,1) draws u (N,>=2) (third draw: m_b_d, dead).  reset_R: the
        kernels also restore self.R = I (the reference only does so in __init__; documented deviation)."""
        m = self._m(mask)
        self.a_b[m] = 0; self.g_b[m] = 0
        self.a_b_d[m] = (u[m, 0] - 0.5) * 2 * self.a_drift                              # :602
        self.g_b_d[m] = (u[m, 1] - 0.5) * 2 * self.g_drift                              # :604
        self.vel[m] = y[m][:, 1:6:2]; self.pos[m] = y[m][:, 0:5:2]; self.quat[m] = y[m][:, 6:10]
        self.acc0[m] = 0
        if reset_R:
            self.R[m] = np.eye(3)

    def reset(self, seed, env_id, episode, y, mask=None):
        """sensor.reset as the kernels do it: bias draws from the env's Philox stream, self.R = I."""
        m = self._m(mask)
        u = np.zeros((self.N, 4))
        u[m] = u32_to_unit(philox_block(seed, np.asarray(env_id)[m], np.asarray(episode)[m], 0xFFFFFFF0, STREAM_SENSOR))
        self.reset_u(u, y, m, reset_R=True)

    def accel(self, z, acc_read, mask=None):                                            # :611-620
        m = self._m(mask)
        self.a_b = np.where(m, self.a_b + self.a_b_d * self.dt, self.a_b)
        return acc_read + (self.a_b[:, None] + self.a_std * z)

    def gyro(self, z, y, mask=None):                                                    # :622-628
        m = self._m(mask)
        self.g_b = np.where(m, self.g_b + self.g_b_d * self.dt, self.g_b)
        return (self.g_b[:, None] + self.g_std * z) + y[:, 10:13]

    def gps(self, z, y):                                                                # :642-647
        return (self.gps_p * z[:, 0:3]) + y[:, 0:5:2], (self.gps_v * z[:, 3:6]) + y[:, 1:6:2]

    def triad(self, z, acc_read, rot, f_m, mask=None):                                  # :649-697
        m = self._m(mask)
        ind = -(self.R @ np.array([0.0, 0.0, -G]))                                      # :658  f_in/M - R@[0,0,-G]
        ind[:, 2] += f_m
        gb = self.accel(z[:, 0:3], acc_read, m) - ind
        mb = np.einsum("nji,nj->ni", rot, MAGNET_VEC[None, :] + self.m_std * z[:, 3:6])  # :662
        Rm = _triad(gb, mb)
        self.R = np.where(m[:, None, None], Rm, self.R)
        return rot_to_quat(Rm), Rm

    def accel_int(self, z, acc_read, rot, f_m, mask=None):                              # :700-715
        m = self._m(mask)
        acc1 = self.accel(z[:, 0:3], acc_read, m)
        _, Rm = self.triad(z[:, 3:9], acc_read, rot, f_m, m)
        a_in = np.einsum("nji,nj->ni", Rm, acc1) + np.array([0, 0, G])
        vel = self.vel + a_in * self.dt
        pos = self.pos + vel * self.dt
        self.acc0 = np.where(m[:, None], a_in, self.acc0)
        self.vel = np.where(m[:, None], vel, self.vel); self.pos = np.where(m[:, None], pos, self.pos)
        return a_in, vel, pos

    def gyro_int(self, z, y, mask=None):                                                # :717-724
        m = self._m(mask)
        w = self.gyro(z, y, m)
        qg = self.quat + deriv_quat(w, self.quat) * self.dt
        self.quat = np.where(m[:, None], qg / np.linalg.norm(qg, axis=1, keepdims=True), self.quat)
        return qg

    def step(self, z, y, acc_read, rot, f_m, mask=None):
        """One env step in the canonical call order of every caller (rl_worker.py:164-175, math_trajectory.py:61-83 incl. its
        GPS blend); z (N,27+); returns the 14-float sensed observation (N,14)."""
        m = self._m(mask)
        _, vel, pos = self.accel_int(z[:, 0:9], acc_read, rot, f_m, m)
        qg = self.gyro_int(z[:, 9:12], y, m)
        w2 = self.gyro(z[:, 12:15], y, m)
        qv = deriv_quat(w2, qg)
        pos_gps, vel_gps = self.gps(z[:, 15:21], y)
        if self.gps_blend > 0:                                                          # math_trajectory.py:71-77
            pos = ((100 - self.gps_blend) * pos + self.gps_blend * pos_gps) / 100
            vel = ((100 - self.gps_blend) * vel + self.gps_blend * vel_gps) / 100
            self.pos = np.where(m[:, None], pos, self.pos); self.vel = np.where(m[:, None], vel, self.vel)
        self.triad(z[:, 21:27], acc_read, rot, f_m, m)
        self.last = dict(w=w2, pos_gps=pos_gps, vel_gps=vel_gps)
        return np.concatenate([np.stack([pos[:, 0], vel[:, 0], pos[:, 1], vel[:, 1], pos[:, 2], vel[:, 2]], axis=1), qg, qv], axis=1)


# --------------------------------------------------------------------------------------
# classical controllers the reference compares against PPO — batched restatements
# --------------------------------------------------------------------------------------
def lqr_gains(clipped=True):
    """Gains exactly as environment/controller/lqr_quad.py:25-111 computes them (two continuous-time AREs)."""
    from scipy.linalg import solve_continuous_are as solve_lqr
    if clipped:                                                               # :25-43
        Q_att = np.diag([5, 1, 5, 1, 0.05, 0.01]) * 50.0
        R_att = np.diag(np.ones(4)) * 40.0
        Q_t = np.diag([1e-08, 1, 1e-08, 1, 1e-08, 0.8]) * 10.0
        R_t = np.diag(np.ones(3)) * 10.0
    else:                                                                     # :44-62
        Q_att = np.diag([5, 0.3, 5, 0.3, 2, 0.3]) * 160.0
        R_att = np.diag(np.ones(4)) * 40.0
        Q_t = np.diag([1e-08, 1, 1e-08, 1, 1e-08, 0.5]) * 60.0
        R_t = np.diag(np.ones(3)) * 5.0
    A = np.zeros((6, 6)); A[0, 1] = A[2, 3] = A[4, 5] = 1.0                   # :67-72, :88-93 (same chain of integrators)
    B_att = np.zeros((6, 4)); B_att[1, 1] = 1 / J_DIAG[0]; B_att[3, 2] = 1 / J_DIAG[1]; B_att[5, 3] = 1 / J_DIAG[2]   # :74-79
    B_t = np.zeros((6, 3)); B_t[1, 0] = B_t[3, 1] = B_t[5, 2] = 1 / M         # :98-103
    K_att = -np.dot(np.linalg.inv(R_att), np.dot(B_att.T, solve_lqr(A, B_att, Q_att, R_att)))     # :82-86
    K_t = -np.dot(np.linalg.inv(R_t), np.dot(B_t.T, solve_lqr(A, B_t, Q_t, R_t)))                 # :107-111
    return K_t, K_att


def lqr_law(K_t, K_att, state, ang, ang_vel):
    """environment/controller/lqr_quad.py:129-157 for N envs: state (N,13), ang (N,3), ang_vel (N,3) -> action (N,4) =
    [F, Mx, My, Mz] for an indirect-control quad.  (deuler_t, :148, is computed by the script but never used.)"""
    z = np.zeros(len(state))
    state_t = np.stack([z, state[:, 1], z, state[:, 3], z, state[:, 5]], axis=1)                 # :129
    F = state_t @ K_t.T                                                                           # :130
    theta_t = np.arctan2(F[:, 0], F[:, 2] + G)                                                    # :135
    phi_t = np.arctan2(-F[:, 1] * np.cos(theta_t), F[:, 2] + G)                                   # :137
    U_1 = M * (F[:, 2] + G) / (np.cos(theta_t) * np.cos(phi_t))                                   # :141
    euler = ang - np.stack([phi_t, theta_t, z], axis=1)                                           # :139,:144
    state_att = np.stack([euler[:, 0], ang_vel[:, 0], euler[:, 1], ang_vel[:, 1], euler[:, 2], ang_vel[:, 2]], axis=1)  # :152
    action = state_att @ K_att.T                                                                  # :153
    action[:, 0] = U_1                                                                            # :154
    return action


PID_GAINS_CLIPPED = dict(xy=(1, -0.0, 0), z=(0.4, -0.0, 0), att=(20, 0, 20), psi=(5, 0, 5))       # pid_vel_control.py:18-22
PID_GAINS_NOT_CLIPPED = dict(xy=(2, -0.0, 0), z=(1, -0.0, 0), att=(180, 0, 50), psi=(40, 0, 20))  # :24-27


class PidControllerOracle:
    """environment/controller/pid_vel_control.py:29-127 (class pid_control + class pid) for N envs.
    Controller memory per env: x_old(6), ix(6) of the six scalar PIDs [x, y, z, phi, theta, psi] and ang_d_ant(3)."""

    def __init__(self, n_envs, t_step=0.01, gains=PID_GAINS_CLIPPED):
        self.N, self.ts = n_envs, t_step
        g = gains
        self.p = np.array([g["xy"][0], g["xy"][0], g["z"][0], g["att"][0], g["att"][0], g["psi"][0]], dtype=np.float64)
        self.i = np.array([g["xy"][1], g["xy"][1], g["z"][1], g["att"][1], g["att"][1], g["psi"][1]], dtype=np.float64)
        self.d = np.array([g["xy"][2], g["xy"][2], g["z"][2], g["att"][2], g["att"][2], g["psi"][2]], dtype=np.float64)
        self.reset()

    def reset(self, mask=None):
        if mask is None:
            self.x_old = np.zeros((self.N, 6)); self.ix = np.zeros((self.N, 6)); self.ang_d_ant = np.zeros((self.N, 3))
        else:
            self.x_old[mask] = 0; self.ix[mask] = 0; self.ang_d_ant[mask] = 0

    def _pid(self, k, x, x_d, dx_d):                                           # class pid :114-127 (the dx argument is overwritten)
        dx = (x - self.x_old[:, k]) / 0.01                                     # pid.ts defaults to 0.01 (:115)
        self.x_old[:, k] = x
        self.ix[:, k] = self.ix[:, k] + (x_d - x) * 0.01
        return self.p[k] * (x_d - x) + self.d[k] * (dx_d - dx) - self.i[k] * self.ix[:, k]

    def control(self, state, ang, xd, psd):
        """pid_control.control :97-110: state (N,13), ang (N,3), velocity set-point xd (3,), yaw set-point psd."""
        z = np.zeros(self.N)
        u_1 = self._pid(0, state[:, 1], xd[0] + z, z)                          # lower_control :48-63
        u_2 = self._pid(1, state[:, 3], xd[1] + z, z)
        u_3 = self._pid(2, state[:, 5], xd[2] + z, z)
        theta_d = np.arctan2(u_1, u_3 + G)
        phi_d = np.arctan2(-u_2 * np.cos(theta_d), u_3 + G)
        U_1 = M * (u_3 + G) / (np.cos(theta_d) * np.cos(phi_d))
        ang_d = np.stack([phi_d, theta_d, psd + z], axis=1)                    # :103
        v_ang_d = (ang_d - self.ang_d_ant) / self.ts                           # :104
        u_5 = self._pid(3, ang[:, 0], ang_d[:, 0], v_ang_d[:, 0])              # upper_control :66-95
        u_6 = self._pid(4, ang[:, 1], ang_d[:, 1], v_ang_d[:, 1])
        u_7 = self._pid(5, ang[:, 2], ang_d[:, 2], v_ang_d[:, 2])
        sp, cp = np.sin(ang[:, 0]), np.cos(ang[:, 0])
        ct, tt = np.cos(ang[:, 1]), np.tan(ang[:, 1])
        Mm = np.zeros((self.N, 3, 3))
        Mm[:, 0, 0] = 1 / J_DIAG[0]; Mm[:, 0, 1] = tt * sp / J_DIAG[1]; Mm[:, 0, 2] = tt * cp / J_DIAG[2]
        Mm[:, 1, 1] = cp / J_DIAG[1]; Mm[:, 1, 2] = -sp / J_DIAG[2]
        Mm[:, 2, 1] = sp / ct / J_DIAG[1]; Mm[:, 2, 2] = cp / ct / J_DIAG[2]
        U = np.einsum("nij,nj->ni", np.linalg.inv(Mm), np.stack([u_5, u_6, u_7], axis=1))    # :93
        self.ang_d_ant = ang_d                                                 # :106
        return np.concatenate([U_1[:, None], U], axis=1)                       # :107


# --------------------------------------------------------------------------------------
# PPO.get_advantages — environment/controller/ppo.py:125-141
# --------------------------------------------------------------------------------------
def gae_advantages(values, masks, rewards, gamma=0.99, lmbda=0.99):
    """Literal restatement for ONE concatenated sequence: values has len(rewards)+1 entries (ppo.py:384 appends 0),
    masks = not terminal.  Returns (returns, normalised advantages)."""
    values = np.asarray(values, dtype=np.float64)
    returns = np.zeros(len(rewards))
    gae = 0.0
    for i in reversed(range(len(rewards))):
        delta = rewards[i] + gamma * values[i + 1] * masks[i] - values[i]          # :135 (the `i == len(rewards)` branch is dead)
        gae = delta + gamma * lmbda * masks[i] * gae                                # :136
        returns[i] = gae + values[i]                                                # :137
    adv = returns - values[:-1]                                                     # :139
    return returns, (adv - np.mean(adv)) / (np.std(adv) + 1e-10)                    # :141


def gae_batched(reward, value, done, gamma=0.99, lmbda=0.99):
    """Time-major batched form the CUDA kernel implements: reward (K,N), value (K+1,N), done u8 (K,N) with bit0 = done and
    bit1 = asynchronous warm-up step (not a transition).  Returns (returns (K,N), raw advantages (K,N), valid (K,N) bool,
    normalised advantages (K,N) with zeros at invalid entries)."""
    K, N = reward.shape
    mask = ((done & 1) == 0).astype(np.float64)
    valid = (done & 2) == 0
    ret = np.zeros((K, N)); adv = np.zeros((K, N))
    gae = np.zeros(N)
    for t in reversed(range(K)):
        delta = reward[t] + gamma * value[t + 1] * mask[t] - value[t]
        gae = delta + gamma * lmbda * mask[t] * gae
        ret[t] = gae + value[t]; adv[t] = gae
    a = adv[valid]
    norm = np.where(valid, (adv - a.mean()) / (a.std() + 1e-10), 0.0)
    return ret, adv, valid, norm
