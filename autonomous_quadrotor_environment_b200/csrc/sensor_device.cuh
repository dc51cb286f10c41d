// sensor_device.cuh — in-kernel restatement of the reference's `sensor` class (environment/quadrotor_env.py:579-724)
// for one environment: IMU / gyro / GPS / magnetometer noise with drifting biases, TRIAD attitude and the two
// dead-reckoning integrators, evaluated in the canonical per-step call order of every user of the class
// (visual_landing/rl_worker.py:164-175, environment/position.py:225-230):
//        accel_int() -> gyro_int() -> gyro() -> gps() -> triad()            = 27 normal draws per env step
// and producing the 14-float sensor-based observation  [p_ins/v_ins interleaved (6), q_gyro (4), 1/2 Omega(gyro) q_gyro (4)].
//
// The reference draws from NumPy's global MT19937 stream, which a counter-based generator cannot reproduce:
// this is a statistical model with the same distributions and the same arithmetic on the draws (the test-side
// checker restates the identical Philox -> normal mapping, so both sides agree to rounding).
//
// Per-env sensor state (QS_SENSOR_STATE_DIM = 20 rows):
//   0 a_b_accel  1 g_b  2 a_b_d  3 g_b_d          scalar biases and their per-episode drift rates  (:600-608,:613,:624)
//   4..6 velocity_t0   7..9 position_t0   10..13 quaternion_t0                                      (:634-638)
//   14..16 third column of the last TRIAD rotation R (the only part of self.R that is read, :658)
//   17..19 acceleration_t0 (write-only, :711)
// Deliberate deviations (documented in DESIGN.md): the aliasing bug of sensor.reset/gyro_int that perturbs the TRUE
// quaternion once per episode (:636-638,:721-722) is not replicated; self.R is re-initialised to identity at
// every sensor reset (the reference only does so in __init__); the magnetometer bias is never used by the
// reference either (m_b, m_b_d are dead stores).
#pragma once
#include "quad_device.cuh"

namespace qs {

constexpr int kSensorStateDim = 20;

// 27 normals per env step; every 32-bit Philox word yields two 16-bit uniforms (u = (h + 0.5) / 65536), i.e. one Box-Muller
// pair.  16-bit resolution truncates the noise at 4.8 sigma with 1.5e-5 granularity — irrelevant for a sensor-noise model
// and it halves the integer work of the generator, the dominant cost of this sub-pass.
// Mapping (the test-side checker restates it): blocks 0..2 of the step give the physical normals P[0..23];
// z[0..14] = P[0..14] (accel_int, first triad, gyro_int, gyro), z[21..26] = P[15..20] (second triad); the six GPS normals
// z[15..20] come from block 3 and are drawn only when the GPS blend consumes them (gps = p.s_gps_blend > 0; the reference's
// gps() draws them on every step and throws them away, :642-647) — with a counter-based generator unused draws cost nothing.
template <typename R>
__device__ __forceinline__ void sensor_block_normals(uint64_t seed, uint32_t env_id, uint32_t episode, uint32_t step, int b, R out[8]) {
    const uint4 u = philox_block(seed, env_id, episode, step * 4u + (uint32_t)b, RNG_SENSOR);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const R u1 = (R(w[k] & 0xFFFFu) + R(0.5)) * R(1.0 / 65536.0);
        const R u2 = (R(w[k] >> 16) + R(0.5)) * R(1.0 / 65536.0);
        box_muller(u1, u2, &out[2 * k], &out[2 * k + 1]);
    }
}
template <typename R>
__device__ __forceinline__ void sensor_normals(uint64_t seed, uint32_t env_id, uint32_t episode, uint32_t step, bool gps, R z[32]) {
    R P[24];
#pragma unroll
    for (int b = 0; b < 3; ++b) sensor_block_normals(seed, env_id, episode, step, b, &P[8 * b]);
#pragma unroll
    for (int k = 0; k < 15; ++k) z[k] = P[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) { z[21 + k] = P[15 + k]; z[15 + k] = R(0); }
    if (gps) {
        R g[8];
        sensor_block_normals(seed, env_id, episode, step, 3, g);
#pragma unroll
        for (int k = 0; k < 6; ++k) z[15 + k] = g[k];
    }
}

template <typename R> __device__ __forceinline__ void cross3(const R a[3], const R b[3], R c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
template <typename R> __device__ __forceinline__ void normalize3(R a[3]) {
    const R inv = M_<R>::rsqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    a[0] *= inv; a[1] *= inv; a[2] *= inv;
}

// sensor.triad :649-697 — returns R = tb @ ti^T (row-major); ti (inertial triad, constant) comes from DevParams.
template <typename R>
__device__ __forceinline__ void triad(const DevParams<R>& p, const R grav_body_in[3], const R mag_body_in[3], R Rm[9]) {
    R g[3] = {grav_body_in[0], grav_body_in[1], grav_body_in[2]};
    R m[3] = {mag_body_in[0], mag_body_in[1], mag_body_in[2]};
    normalize3(g);                                        // :668
    normalize3(m);                                        // :670
    R t2[3], t3[3];
    cross3(g, m, t2); normalize3(t2);                     // :675-676  (t1b = g, already unit)
    cross3(g, t2, t3); normalize3(t3);                    // :678-679
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            Rm[3 * r + c] = g[r] * p.s_ti[c] + t2[r] * p.s_ti[3 + c] + t3[r] * p.s_ti[6 + c];   // :693  tb @ ti.T
}

// quad.mat_rot and quad.accelerometer_read (:315, :371) at state y, as the trailing drone_eq call of a step leaves them.
// With accel = R f_b / M - G z^ (:364-367):  accelerometer_read = R^T (accel - G z^) = f_b / M - 2 G (r6, r7, r8),
// which saves the two matrix-vector products of the literal form (used by the FP32 production kernels; the FP64 parity
// path keeps the literal form).
template <typename R>
__device__ __forceinline__ void accel_read(const DevParams<R>& p, const Ctrl<R>& c, const R y[13], R rot[9], R acc[3]) {
    R qn[4];
    quat_normalize(&y[6], qn);
    quat_rot_mat(qn, rot);
    const R vbx = rot[0] * y[1] + rot[3] * y[3] + rot[6] * y[5];
    const R vby = rot[1] * y[1] + rot[4] * y[3] + rot[7] * y[5];
    const R vbz = rot[2] * y[1] + rot[5] * y[3] + rot[8] * y[5];
    const R g2 = R(-2) * p.g;
    acc[0] = g2 * rot[6] - p.kd_m[0] * (M_<R>::abs(vbx) * vbx);
    acc[1] = g2 * rot[7] - p.kd_m[1] * (M_<R>::abs(vby) * vby);
    acc[2] = g2 * rot[8] + (c.f_m - p.kd_m[2] * (M_<R>::abs(vbz) * vbz));
}

// sensor.reset :630-640 + bias_reset :600-608 (called once per episode, after quad.reset's warm-up steps)
template <typename R>
__device__ __forceinline__ void sensor_reset(const DevParams<R>& p, uint64_t seed, uint32_t env_id, uint32_t episode,
                                             const R y[13], R s[kSensorStateDim]) {
    const uint4 u = philox_block(seed, env_id, episode, 0xFFFFFFF0u, RNG_SENSOR);
    s[0] = R(0); s[1] = R(0);
    s[2] = (u32_to_unit<R>(u.x) - R(0.5)) * R(2) * p.s_accel_drift;       // :602
    s[3] = (u32_to_unit<R>(u.y) - R(0.5)) * R(2) * p.s_gyro_drift;        // :604
    s[4] = y[1]; s[5] = y[3]; s[6] = y[5];                                // velocity_t0 :637
    s[7] = y[0]; s[8] = y[2]; s[9] = y[4];                                // position_t0 :636
    s[10] = y[6]; s[11] = y[7]; s[12] = y[8]; s[13] = y[9];               // quaternion_t0 :638
    s[14] = R(0); s[15] = R(0); s[16] = R(1);                             // R = I
    s[17] = R(0); s[18] = R(0); s[19] = R(0);                             // acceleration_t0 :633
}

// One env step of the sensor model.  y = TRUE state after the step, acc_read = quad.accelerometer_read (:371),
// rot = quad.mat_rot (:315), f_m = F/M (induced acceleration of the rotors, :658).  Updates s, writes obs14.
template <typename R>
__device__ __forceinline__ void sensor_step(const DevParams<R>& p, const R z[32], const R y[13], const R acc_read[3],
                                            const R rot[9], R f_m, R s[kSensorStateDim], R obs[14]) {
    const R dt = p.dt;
    // ---- accel_int :700-715
    s[0] += s[2] * dt;                                                                     // accel() :613
    const R acc1[3] = {acc_read[0] + s[0] + p.s_accel_std * z[0], acc_read[1] + s[0] + p.s_accel_std * z[1],
                       acc_read[2] + s[0] + p.s_accel_std * z[2]};
    R Rm[9];
    {   // triad()
        s[0] += s[2] * dt;
        const R ind[3] = {p.g * s[14], p.g * s[15], f_m + p.g * s[16]};                    // :658  f_in/M - R@[0,0,-G]
        const R gb[3] = {acc_read[0] + s[0] + p.s_accel_std * z[3] - ind[0], acc_read[1] + s[0] + p.s_accel_std * z[4] - ind[1],
                         acc_read[2] + s[0] + p.s_accel_std * z[5] - ind[2]};
        const R mi[3] = {p.s_mag[0] + p.s_mag_std * z[6], p.s_mag[1] + p.s_mag_std * z[7], p.s_mag[2] + p.s_mag_std * z[8]};
        const R mb[3] = {rot[0] * mi[0] + rot[3] * mi[1] + rot[6] * mi[2], rot[1] * mi[0] + rot[4] * mi[1] + rot[7] * mi[2],
                         rot[2] * mi[0] + rot[5] * mi[1] + rot[8] * mi[2]};               // :662  mat_rot.T @ ...
        triad(p, gb, mb, Rm);
    }
    const R a_in[3] = {Rm[0] * acc1[0] + Rm[3] * acc1[1] + Rm[6] * acc1[2], Rm[1] * acc1[0] + Rm[4] * acc1[1] + Rm[7] * acc1[2],
                       Rm[2] * acc1[0] + Rm[5] * acc1[1] + Rm[8] * acc1[2] + p.g};        // :705  R.T @ accel_body + [0,0,G]
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        s[4 + k] += a_in[k] * dt;                                                          // velocity :707
        s[7 + k] += s[4 + k] * dt;                                                         // position :708
        s[17 + k] = a_in[k];
    }
    // ---- gyro_int :717-724
    s[1] += s[3] * dt;                                                                     // gyro() :624
    const R w1[3] = {y[10] + s[1] + p.s_gyro_std * z[9], y[11] + s[1] + p.s_gyro_std * z[10], y[12] + s[1] + p.s_gyro_std * z[11]};
    R dq[4], qg[4];
    deriv_quat(w1, &s[10], dq);
#pragma unroll
    for (int k = 0; k < 4; ++k) qg[k] = s[10 + k] + dq[k] * dt;                            // :721-722 (returned un-normalised)
    quat_normalize(qg, &s[10]);                                                            // :723
    // ---- gyro :622-628
    s[1] += s[3] * dt;
    const R w2[3] = {y[10] + s[1] + p.s_gyro_std * z[12], y[11] + s[1] + p.s_gyro_std * z[13], y[12] + s[1] + p.s_gyro_std * z[14]};
    R qv[4];
    deriv_quat(w2, qg, qv);                                                                // rl_worker.py:168
    // ---- gps :642-647 consumes z[15..20].  Its readings enter only through the optional complementary blend of the landing
    //      stack (visual_landing/math_trajectory.py:71-77, GPS_P per cent; off by default like the script's GPS = False),
    //      which also writes the blended estimate back into the dead-reckoning integrators.
    if (p.s_gps_blend > R(0)) {
        const R wg = p.s_gps_blend, wa = R(100) - p.s_gps_blend;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const R pos_gps = y[2 * k] + p.s_gps_p * z[15 + k];
            const R vel_gps = y[2 * k + 1] + p.s_gps_v * z[18 + k];
            s[7 + k] = (wa * s[7 + k] + wg * pos_gps) / R(100);
            s[4 + k] = (wa * s[4 + k] + wg * vel_gps) / R(100);
        }
    }
    // ---- triad :649-697 (updates self.R for the next step)
    {
        s[0] += s[2] * dt;
        const R ind[3] = {p.g * Rm[2], p.g * Rm[5], f_m + p.g * Rm[8]};
        const R gb[3] = {acc_read[0] + s[0] + p.s_accel_std * z[21] - ind[0], acc_read[1] + s[0] + p.s_accel_std * z[22] - ind[1],
                         acc_read[2] + s[0] + p.s_accel_std * z[23] - ind[2]};
        const R mi[3] = {p.s_mag[0] + p.s_mag_std * z[24], p.s_mag[1] + p.s_mag_std * z[25], p.s_mag[2] + p.s_mag_std * z[26]};
        const R mb[3] = {rot[0] * mi[0] + rot[3] * mi[1] + rot[6] * mi[2], rot[1] * mi[0] + rot[4] * mi[1] + rot[7] * mi[2],
                         rot[2] * mi[0] + rot[5] * mi[1] + rot[8] * mi[2]};
        R R2[9];
        triad(p, gb, mb, R2);
        s[14] = R2[2]; s[15] = R2[5]; s[16] = R2[8];
    }
    obs[0] = s[7]; obs[1] = s[4]; obs[2] = s[8]; obs[3] = s[5]; obs[4] = s[9]; obs[5] = s[6];   // rl_worker.py:171-173
#pragma unroll
    for (int k = 0; k < 4; ++k) { obs[6 + k] = qg[k]; obs[10 + k] = qv[k]; }
}

}  // namespace qs
