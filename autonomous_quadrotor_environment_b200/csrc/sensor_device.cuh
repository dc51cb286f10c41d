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

// ---- the methods of the class, one device function each.  s = the 20-row sensor state; z = the standard normals the
// reference's np.random.normal calls of that method consume, in call order.  The same functions serve the fused per-step
// model (sensor_step, Philox normals) and the per-method entry point qs_sensor_call (caller-provided normals), which is what
// pins the arithmetic to the reference class itself (tests/golden/sensor_vectors.npz).

// sensor.accel :611-620 — read_accel_body + N(a_b_accel, a_std)
template <typename R>
__device__ __forceinline__ void sensor_accel(const DevParams<R>& p, R s[kSensorStateDim], const R acc_read[3], const R z[3], R out[3]) {
    s[0] += s[2] * p.dt;                                                                   // :613
#pragma unroll
    for (int k = 0; k < 3; ++k) out[k] = acc_read[k] + (s[0] + p.s_accel_std * z[k]);      // :615-619
}

// sensor.gyro :622-628 — read_gyro + N(g_b, g_std)
template <typename R>
__device__ __forceinline__ void sensor_gyro(const DevParams<R>& p, R s[kSensorStateDim], const R y[13], const R z[3], R out[3]) {
    s[1] += s[3] * p.dt;                                                                   // :624
#pragma unroll
    for (int k = 0; k < 3; ++k) out[k] = (s[1] + p.s_gyro_std * z[k]) + y[10 + k];         // :626-628
}

// sensor.gps :642-647 — z[0..2] position errors, z[3..5] velocity errors
template <typename R>
__device__ __forceinline__ void sensor_gps(const DevParams<R>& p, const R y[13], const R z[6], R pos[3], R vel[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        pos[k] = p.s_gps_p * z[k] + y[2 * k];
        vel[k] = p.s_gps_v * z[3 + k] + y[2 * k + 1];
    }
}

// sensor.triad :649-697 — z[0..2] feed its accel() call, z[3..5] the magnetometer noise.  Reads the third column of the
// previous self.R (s[14..16], :658), writes the new one, returns R = tb @ ti^T (row-major).
template <typename R>
__device__ __forceinline__ void sensor_triad(const DevParams<R>& p, R s[kSensorStateDim], const R acc_read[3], const R rot[9], R f_m,
                                             const R z[6], R Rm[9]) {
    const R ind[3] = {p.g * s[14], p.g * s[15], f_m + p.g * s[16]};                        // :658  f_in/M - R@[0,0,-G]
    R gb[3];
    sensor_accel(p, s, acc_read, z, gb);
#pragma unroll
    for (int k = 0; k < 3; ++k) gb[k] -= ind[k];                                           // :659
    const R mi[3] = {p.s_mag[0] + p.s_mag_std * z[3], p.s_mag[1] + p.s_mag_std * z[4], p.s_mag[2] + p.s_mag_std * z[5]};
    const R mb[3] = {rot[0] * mi[0] + rot[3] * mi[1] + rot[6] * mi[2], rot[1] * mi[0] + rot[4] * mi[1] + rot[7] * mi[2],
                     rot[2] * mi[0] + rot[5] * mi[1] + rot[8] * mi[2]};                   // :662  mat_rot.T @ ...
    triad(p, gb, mb, Rm);
    s[14] = Rm[2]; s[15] = Rm[5]; s[16] = Rm[8];                                           // self.R = tb @ ti.T  :693
}

// q of sensor.triad's return value (:695-696): Rotation.from_matrix(R.T).as_quat() re-ordered scalar-first.  SciPy's
// conversion of an orthogonal matrix (spatial/transform/_rotation: Markley's method — pick the largest of the diagonal
// entries and the trace, build the quaternion from that row, normalise); m = R^T row-major.
template <typename R>
__device__ __forceinline__ void rot_to_quat_scipy(const R Rm[9], R q[4]) {
    const R m00 = Rm[0], m01 = Rm[3], m02 = Rm[6], m10 = Rm[1], m11 = Rm[4], m12 = Rm[7], m20 = Rm[2], m21 = Rm[5], m22 = Rm[8];
    const R tr = m00 + m11 + m22;
    const R dec[4] = {m00, m11, m22, tr};
    int choice = 0;
#pragma unroll
    for (int k = 1; k < 4; ++k) choice = (dec[k] > dec[choice]) ? k : choice;              // argmax: first maximum
    R x, y, zq, w;
    if (choice == 0)      { x = R(1) - tr + R(2) * m00; y = m10 + m01; zq = m20 + m02; w = m21 - m12; }
    else if (choice == 1) { x = m10 + m01; y = R(1) - tr + R(2) * m11; zq = m21 + m12; w = m02 - m20; }
    else if (choice == 2) { x = m20 + m02; y = m21 + m12; zq = R(1) - tr + R(2) * m22; w = m10 - m01; }
    else                  { x = m21 - m12; y = m02 - m20; zq = m10 - m01; w = R(1) + tr; }
    const R inv = R(1) / M_<R>::sqrt(x * x + y * y + zq * zq + w * w);
    q[0] = w * inv; q[1] = x * inv; q[2] = y * inv; q[3] = zq * inv;
}

// sensor.accel_int :700-715 — z[0..2] its own accel() call, z[3..8] the triad() call.  out: acceleration, velocity, position
template <typename R>
__device__ __forceinline__ void sensor_accel_int(const DevParams<R>& p, R s[kSensorStateDim], const R acc_read[3], const R rot[9],
                                                 R f_m, const R z[9], R a_in[3], R* R_out = nullptr) {
    R acc1[3], Rm[9];
    sensor_accel(p, s, acc_read, z, acc1);                                                 // :702
    sensor_triad(p, s, acc_read, rot, f_m, z + 3, Rm);                                     // :703
    a_in[0] = Rm[0] * acc1[0] + Rm[3] * acc1[1] + Rm[6] * acc1[2];                         // :705  R.T @ accel_body + [0,0,G]
    a_in[1] = Rm[1] * acc1[0] + Rm[4] * acc1[1] + Rm[7] * acc1[2];
    a_in[2] = Rm[2] * acc1[0] + Rm[5] * acc1[1] + Rm[8] * acc1[2] + p.g;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        s[4 + k] += a_in[k] * p.dt;                                                        // velocity :707
        s[7 + k] += s[4 + k] * p.dt;                                                       // position :708
        s[17 + k] = a_in[k];                                                               // :710
    }
    if (R_out) {
#pragma unroll
        for (int k = 0; k < 9; ++k) R_out[k] = Rm[k];                                      // self.R after the call (set by triad)
    }
}

// sensor.gyro_int :717-724 — returns q BEFORE normalisation (the reference returns the array it updated in place)
template <typename R>
__device__ __forceinline__ void sensor_gyro_int(const DevParams<R>& p, R s[kSensorStateDim], const R y[13], const R z[3], R qg[4]) {
    R w1[3], dq[4];
    sensor_gyro(p, s, y, z, w1);                                                           // :718
    deriv_quat(w1, &s[10], dq);                                                            // :720
#pragma unroll
    for (int k = 0; k < 4; ++k) qg[k] = s[10 + k] + dq[k] * p.dt;                          // :721-722
    quat_normalize(qg, &s[10]);                                                            // :723
}

// One env step of the sensor model in the canonical call order.  y = TRUE state after the step, acc_read =
// quad.accelerometer_read (:371), rot = quad.mat_rot (:315), f_m = F/M (induced acceleration of the rotors, :658).
// Updates s, writes obs14 (rl_worker.py:171-173 / math_trajectory.py:61-83).
template <typename R>
__device__ __forceinline__ void sensor_step(const DevParams<R>& p, const R z[32], const R y[13], const R acc_read[3],
                                            const R rot[9], R f_m, R s[kSensorStateDim], R obs[14]) {
    R a_in[3], qg[4], w2[3], qv[4];
    sensor_accel_int(p, s, acc_read, rot, f_m, z, a_in);                                   // z[0..8]
    sensor_gyro_int(p, s, y, z + 9, qg);                                                   // z[9..11]
    sensor_gyro(p, s, y, z + 12, w2);                                                      // z[12..14]
    deriv_quat(w2, qg, qv);                                                                // rl_worker.py:168
    // gps :642-647 consumes z[15..20].  Its readings enter only through the optional complementary blend of the landing
    // stack (visual_landing/math_trajectory.py:71-77, GPS_P per cent; off by default like the script's GPS = False),
    // which also writes the blended estimate back into the dead-reckoning integrators.
    if (p.s_gps_blend > R(0)) {
        R pos_gps[3], vel_gps[3];
        sensor_gps(p, y, z + 15, pos_gps, vel_gps);
        const R wg = p.s_gps_blend, wa = R(100) - p.s_gps_blend;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            s[7 + k] = (wa * s[7 + k] + wg * pos_gps[k]) / R(100);
            s[4 + k] = (wa * s[4 + k] + wg * vel_gps[k]) / R(100);
        }
    }
    R R2[9];
    sensor_triad(p, s, acc_read, rot, f_m, z + 21, R2);                                    // z[21..26]; updates self.R for the next step
    obs[0] = s[7]; obs[1] = s[4]; obs[2] = s[8]; obs[3] = s[5]; obs[4] = s[9]; obs[5] = s[6];   // rl_worker.py:171-173
#pragma unroll
    for (int k = 0; k < 4; ++k) { obs[6 + k] = qg[k]; obs[10 + k] = qv[k]; }
}

}  // namespace qs
