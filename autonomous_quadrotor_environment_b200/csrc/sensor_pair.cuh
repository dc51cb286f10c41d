// sensor_pair.cuh — the reference's `sensor` class (environment/quadrotor_env.py:579-724) for an env PAIR held as packed
// register pairs {A, B}: the same per-step sequence as sensor_step (sensor_device.cuh) —
//        accel_int() -> gyro_int() -> gyro() -> gps() -> triad()          = 27 normal draws per env step
// with every FP32 operation on the packed pipe (FFMA2 / FMUL2 / FADD2: one issue slot for both envs).  Only the Philox
// rounds, the int->float conversions and the MUFU calls (log2, sqrt, rsqrt, sin, cos) remain per env.
//
// Normals are produced one Philox block (8 normals per env) at a time, right before their first use (mapping: sensor_normals,
// sensor_device.cuh — blocks 0..2 for the model, block 3 only when the GPS blend reads it), so that at most two blocks are live: the sensor phase then fits the register budget of the two-envs-per-lane step kernel (step_pair.cuh).
// The mapping counter -> normal is the one of sensor_normals (16-bit uniforms, Box-Muller at 2*pi*(u2-1/2)); results agree
// with the scalar routine to FP32 rounding (the test compares the two kernels field by field).
#pragma once
#include "packed_device.cuh"
#include "sensor_device.cuh"

namespace qs {

__device__ __forceinline__ P2 psub(P2 a, P2 b) { return pfma(b, bc(-1.f), a); }
__device__ __forceinline__ P2 pneg(P2 a) { return pmul(a, bc(-1.f)); }
__device__ __forceinline__ P2 psel(bool ca, bool cb, P2 t, P2 f) { return pk(ca ? t.v.x : f.v.x, cb ? t.v.y : f.v.y); }

struct SensorRng2 { uint64_t seed; const uint32_t* rk; uint32_t id[2], ep[2], step[2]; };   // rk = SimView::rk (round keys of seed)

// 8 normals per env from Philox block b of the env's sensor stream (sensor_normals, sensor_device.cuh)
__device__ __forceinline__ void sensor_normals_block2(const SensorRng2& r, int b, P2 z[8]) {
#ifdef QS_PHILOX_KEYS_IN_REGS
    const uint4 ua = philox_block(r.seed, r.id[0], r.ep[0], r.step[0] * 4u + (uint32_t)b, RNG_SENSOR);
    const uint4 ub = philox_block(r.seed, r.id[1], r.ep[1], r.step[1] * 4u + (uint32_t)b, RNG_SENSOR);
#else
    const uint4 ua = philox4x32_10_rk(make_uint4(r.id[0], r.ep[0], r.step[0] * 4u + (uint32_t)b, RNG_SENSOR), r.rk);
    const uint4 ub = philox4x32_10_rk(make_uint4(r.id[1], r.ep[1], r.step[1] * 4u + (uint32_t)b, RNG_SENSOR), r.rk);
#endif
    const uint32_t wa[4] = {ua.x, ua.y, ua.z, ua.w}, wb[4] = {ub.x, ub.y, ub.z, ub.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const P2 h1 = pk((float)(wa[k] & 0xFFFFu), (float)(wb[k] & 0xFFFFu));
        const P2 h2 = pk((float)(wa[k] >> 16), (float)(wb[k] >> 16));
        const P2 u1 = pfma(h1, bc(1.0f / 65536.0f), bc(0.5f / 65536.0f));                  // (h + 1/2) / 65536, exact
        const P2 ang = pfma(h2, bc(6.28318548f / 65536.0f), bc(6.28318548f * (0.5f / 65536.0f - 0.5f)));   // 2 pi (u2 - 1/2)
        const P2 l = pmul(pk(fast_log2f(u1.v.x), fast_log2f(u1.v.y)), bc(-2.0f * 0.693147182f));  // -2 ln u1
        const P2 nr = pk(-fast_sqrtf(l.v.x), -fast_sqrtf(l.v.y));
        const P2 cs = pk(__cosf(ang.v.x), __cosf(ang.v.y)), sn = pk(__sinf(ang.v.x), __sinf(ang.v.y));
        z[2 * k] = pmul(nr, cs);
        z[2 * k + 1] = pmul(nr, sn);
    }
}

__device__ __forceinline__ void pnormalize3(P2 a[3]) {
    const P2 inv = prsqrt(pfma(a[0], a[0], pfma(a[1], a[1], pmul(a[2], a[2]))));
    a[0] = pmul(a[0], inv); a[1] = pmul(a[1], inv); a[2] = pmul(a[2], inv);
}
__device__ __forceinline__ void pcross3(const P2 a[3], const P2 b[3], P2 c[3]) {
    const P2 n0 = pneg(a[0]), n1 = pneg(a[1]), n2 = pneg(a[2]);
    c[0] = pfma(a[1], b[2], pmul(n2, b[1]));
    c[1] = pfma(a[2], b[0], pmul(n0, b[2]));
    c[2] = pfma(a[0], b[1], pmul(n1, b[0]));
}

// sensor.triad :649-697 — R = tb @ ti^T (row-major).  FULL = false: only the third column (Rm[2], Rm[5], Rm[8]), the
// part of self.R the next step reads (:658).
template <bool FULL>
__device__ __forceinline__ void triad2(const DevParams<float>& p, P2 g[3], P2 m[3], P2 Rm[9]) {
    // FP32 production form: two of the reference's four normalisations are mathematically redundant and are skipped here
    // (each is a serial FMA-FMA-FMA-MUFU.RSQ-FMUL chain in a latency-bound phase): t2 = (g x m)/|g x m| does not depend on |m|
    // (:670), and t3 = t1 x t2 of two orthogonal unit vectors is a unit vector already (:679).  The results move by FP32
    // rounding only; the FP64 parity path (sensor_device.cuh: triad) keeps all four.
    pnormalize3(g);                                       // :668
    P2 t2[3], t3[3];
    pcross3(g, m, t2); pnormalize3(t2);                   // :675-676
    pcross3(g, t2, t3);                                   // :678
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = FULL ? 0 : 2; c < 3; ++c)
            Rm[3 * r + c] = pfma(g[r], bc(p.s_ti[c]), pfma(t2[r], bc(p.s_ti[3 + c]), pmul(t3[r], bc(p.s_ti[6 + c]))));   // :693
}

// deriv_quat utility:58-69 for a pair
__device__ __forceinline__ void deriv_quat2(const P2 w[3], const P2 q[4], P2 dq[4]) {
    const P2 hx = pmul(w[0], bc(0.5f)), hy = pmul(w[1], bc(0.5f)), hz = pmul(w[2], bc(0.5f));
    const P2 nx = pmul(w[0], bc(-0.5f)), ny = pmul(w[1], bc(-0.5f)), nz = pmul(w[2], bc(-0.5f));
    dq[0] = pfma(nx, q[1], pfma(ny, q[2], pmul(nz, q[3])));
    dq[1] = pfma(hx, q[0], pfma(hz, q[2], pmul(ny, q[3])));
    dq[2] = pfma(hy, q[0], pfma(nz, q[1], pmul(hx, q[3])));
    dq[3] = pfma(hz, q[0], pfma(hy, q[1], pmul(nx, q[2])));
}

// quad.mat_rot and quad.accelerometer_read (:315, :371) at state y for a pair.  With accel = R f_b / M - G z^ (:364-367),
//   accelerometer_read = R^T (accel - G z^) = f_b / M - 2 G R^T z^ = f_b / M - 2 G (r6, r7, r8).
__device__ __forceinline__ void accel_read2(const DevParams<float>& p, P2 f_m, const P2 y[13], P2 rot[9], P2 acc[3]) {
    const P2 inv = prsqrt(pfma(y[6], y[6], pfma(y[7], y[7], pfma(y[8], y[8], pmul(y[9], y[9])))));
    const P2 a = pmul(y[6], inv), b = pmul(y[7], inv), cq = pmul(y[8], inv), d = pmul(y[9], inv);
    const P2 one = bc(1.f), m2 = bc(-2.f);
    const P2 a2 = padd(a, a), b2 = padd(b, b), c2 = padd(cq, cq);
    const P2 nb2 = pmul(b, m2), nc2 = pmul(cq, m2), nd2 = pmul(d, m2);
    const P2 tb = pfma(nb2, b, one);
    rot[0] = pfma(nd2, d, pfma(nc2, cq, one)); rot[4] = pfma(nd2, d, tb); rot[8] = pfma(nc2, cq, tb);
    const P2 bc2 = pmul(b2, cq), bd2 = pmul(b2, d), cd2 = pmul(c2, d);
    rot[1] = pfma(nd2, a, bc2); rot[3] = pfma(a2, d, bc2);
    rot[2] = pfma(a2, cq, bd2); rot[6] = pfma(nc2, a, bd2);
    rot[5] = pfma(nb2, a, cd2); rot[7] = pfma(a2, b, cd2);
    const P2 vx = y[1], vy = y[3], vz = y[5];
    const P2 vbx = pfma(rot[0], vx, pfma(rot[3], vy, pmul(rot[6], vz)));
    const P2 vby = pfma(rot[1], vx, pfma(rot[4], vy, pmul(rot[7], vz)));
    const P2 vbz = pfma(rot[2], vx, pfma(rot[5], vy, pmul(rot[8], vz)));
    const P2 g2 = bc(-2.f * p.g);
    acc[0] = pfma(g2, rot[6], pmul(bc(-p.kd_m[0]), pabsmul(vbx)));
    acc[1] = pfma(g2, rot[7], pmul(bc(-p.kd_m[1]), pabsmul(vby)));
    acc[2] = pfma(g2, rot[8], pfma(bc(-p.kd_m[2]), pabsmul(vbz), f_m));
}

// One env step of the sensor model for a pair.  y = TRUE state after the step, acc_read / rot from accel_read2, f_m = F/M.
// Updates s, writes obs14 (rl_worker.py:171-173).
// zpre: NULL, or the 24 normals of blocks 0..2 drawn ahead of time (integrate_rk4_2_fused: zpre[8 b + k] = normal k of block b).
// ZS > 0: zpre is NOT null and normal j of the pair sits at zpre[j * ZS] (the producer warps of step_kernel_pair leave the normals
// of a chunk as rows of 64 floats in shared memory: ZS = 32); the Philox / Box-Muller code of blocks 0..2 is then not instantiated.
template <int ZS = 0>
__device__ __forceinline__ void sensor_step2(const DevParams<float>& p, const SensorRng2& rng, const P2 y[13], const P2 acc_read[3],
                                             const P2 rot[9], P2 f_m, P2 s[kSensorStateDim], P2 obs[14], const P2* zpre = nullptr) {
    const P2 dt = bc(p.dt), sa = bc(p.s_accel_std), sg = bc(p.s_gyro_std), sm = bc(p.s_mag_std), ng = bc(-p.g);
    constexpr int kZs = ZS > 0 ? ZS : 1;
    P2 z0[8], z1[8];
    if (ZS > 0 || zpre) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { z0[k] = zpre[k * kZs]; z1[k] = zpre[(8 + k) * kZs]; }
    } else
    sensor_normals_block2(rng, 0, z0);                                                     // z[0..7]
    // ---- accel_int :700-715
    s[0] = pfma(s[2], dt, s[0]);                                                           // accel() :613
    P2 acc1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) acc1[k] = pfma(sa, z0[k], padd(acc_read[k], s[0]));
    if (ZS == 0 && !zpre) sensor_normals_block2(rng, 1, z1);                               // z[8..15]
    P2 Rm[9];
    {   // triad()
        s[0] = pfma(s[2], dt, s[0]);
        P2 gb[3], mi[3], mb[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) gb[k] = pfma(ng, s[14 + k], pfma(sa, z0[3 + k], padd(acc_read[k], s[0])));   // :658
        gb[2] = psub(gb[2], f_m);
        mi[0] = pfma(sm, z0[6], bc(p.s_mag[0])); mi[1] = pfma(sm, z0[7], bc(p.s_mag[1])); mi[2] = pfma(sm, z1[0], bc(p.s_mag[2]));
#pragma unroll
        for (int c = 0; c < 3; ++c) mb[c] = pfma(rot[c], mi[0], pfma(rot[3 + c], mi[1], pmul(rot[6 + c], mi[2])));   // :662
        triad2<true>(p, gb, mb, Rm);
    }
    P2 a_in[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) a_in[c] = pfma(Rm[c], acc1[0], pfma(Rm[3 + c], acc1[1], pmul(Rm[6 + c], acc1[2])));   // :705
    a_in[2] = padd(a_in[2], bc(p.g));
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        s[4 + k] = pfma(a_in[k], dt, s[4 + k]);                                            // velocity :707
        s[7 + k] = pfma(s[4 + k], dt, s[7 + k]);                                           // position :708
        s[17 + k] = a_in[k];
    }
    // ---- gyro_int :717-724
    s[1] = pfma(s[3], dt, s[1]);                                                           // gyro() :624
    P2 w1[3], dq[4], qg[4];
#pragma unroll
    for (int k = 0; k < 3; ++k) w1[k] = pfma(sg, z1[1 + k], padd(y[10 + k], s[1]));        // z[9..11]
    deriv_quat2(w1, &s[10], dq);
#pragma unroll
    for (int k = 0; k < 4; ++k) qg[k] = pfma(dq[k], dt, s[10 + k]);                        // :721-722 (returned un-normalised)
    {
        const P2 inv = prsqrt(pfma(qg[0], qg[0], pfma(qg[1], qg[1], pfma(qg[2], qg[2], pmul(qg[3], qg[3])))));
#pragma unroll
        for (int k = 0; k < 4; ++k) s[10 + k] = pmul(qg[k], inv);                          // :723
    }
    // ---- gyro :622-628
    s[1] = pfma(s[3], dt, s[1]);
    P2 w2[3], qv[4];
#pragma unroll
    for (int k = 0; k < 3; ++k) w2[k] = pfma(sg, z1[4 + k], padd(y[10 + k], s[1]));        // z[12..14]
    deriv_quat2(w2, qg, qv);                                                               // rl_worker.py:168
    // ---- gps :642-647: z[15..20] = block 3, drawn only when the optional complementary blend reads them (math_trajectory.py:71-77)
    const P2 z21 = z1[7];                                                                  // P[15] = z[21]
    if (p.s_gps_blend > 0.f) {
        const P2 wg = bc(p.s_gps_blend * 0.01f), wa = bc((100.f - p.s_gps_blend) * 0.01f);
        sensor_normals_block2(rng, 3, z1);
        const P2 zp[3] = {z1[0], z1[1], z1[2]}, zv[3] = {z1[3], z1[4], z1[5]};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const P2 pos_gps = pfma(bc(p.s_gps_p), zp[k], y[2 * k]);
            const P2 vel_gps = pfma(bc(p.s_gps_v), zv[k], y[2 * k + 1]);
            s[7 + k] = pfma(wa, s[7 + k], pmul(wg, pos_gps));
            s[4 + k] = pfma(wa, s[4 + k], pmul(wg, vel_gps));
        }
    }
    // ---- triad :649-697 (updates self.R for the next step)
    if (ZS > 0 || zpre) {
#pragma unroll
        for (int k = 0; k < 8; ++k) z0[k] = zpre[(16 + k) * kZs];
    } else
    sensor_normals_block2(rng, 2, z0);                                                     // P[16..23]: z[22..26] = P[16..20]
    {
        s[0] = pfma(s[2], dt, s[0]);
        P2 gb[3], mi[3], mb[3], R2[9];
        gb[0] = pfma(ng, Rm[2], pfma(sa, z21, padd(acc_read[0], s[0])));                   // z[21..23]
        gb[1] = pfma(ng, Rm[5], pfma(sa, z0[0], padd(acc_read[1], s[0])));
        gb[2] = psub(pfma(ng, Rm[8], pfma(sa, z0[1], padd(acc_read[2], s[0]))), f_m);
#pragma unroll
        for (int k = 0; k < 3; ++k) mi[k] = pfma(sm, z0[2 + k], bc(p.s_mag[k]));           // z[24..26]
#pragma unroll
        for (int c = 0; c < 3; ++c) mb[c] = pfma(rot[c], mi[0], pfma(rot[3 + c], mi[1], pmul(rot[6 + c], mi[2])));
        triad2<false>(p, gb, mb, R2);
        s[14] = R2[2]; s[15] = R2[5]; s[16] = R2[8];
    }
    obs[0] = s[7]; obs[1] = s[4]; obs[2] = s[8]; obs[3] = s[5]; obs[4] = s[9]; obs[5] = s[6];
#pragma unroll
    for (int k = 0; k < 4; ++k) { obs[6 + k] = qg[k]; obs[10 + k] = qv[k]; }
}

// ---------------------------------------------------------------------------------------------------------------------
// STREAMING form of sensor_step2: the pair's 20 state rows stay in shared memory and are read right before / written right
// after each use, and every output is handed to `out` the moment it exists, so that neither the state (40 registers) nor the
// sensed observation (28) is ever held in registers as a whole: the sensor phase then fits the register budget of 12 warps
// per SM (168) next to the dynamics.  Same operations in the same order per quantity as sensor_step2 (the GPS blend reloads the
// integrators it has just stored), hence the same results.
// ---------------------------------------------------------------------------------------------------------------------
struct SensorMem {
    float2* st;              // shared memory; row k of this lane's pair = st[k * 32]
    bool any_warm;           // warp-uniform: some env of the warp is in a warm-up step of quad.reset (state kept, sensor bypassed)
    bool w0, w1;             // ... this pair's halves
    __device__ __forceinline__ P2 ld(int k) const { P2 r; r.v = st[k * 32]; return r; }
    __device__ __forceinline__ void stv(int k, P2 x) const {
        if (any_warm) x = psel(w0, w1, ld(k), x);
        st[k * 32] = x.v;
    }
};

// out(k, value): row k of the 14-float sensed observation of the pair (computed value; the caller substitutes the true
// observation for envs in a warm-up step)
template <typename Out>
__device__ __forceinline__ void sensor_step2_stream(const DevParams<float>& p, const SensorRng2& rng, const P2 y[13], P2 f_m,
                                                    const SensorMem& sm, Out&& out, const P2* zpre = nullptr) {
    const P2 dt = bc(p.dt), sa = bc(p.s_accel_std), sg = bc(p.s_gyro_std), smg = bc(p.s_mag_std), ng = bc(-p.g);
    P2 rot[9], acc_read[3];
    accel_read2(p, f_m, y, rot, acc_read);
    P2 z0[8], z1[8];
    if (zpre) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { z0[k] = zpre[k]; z1[k] = zpre[8 + k]; }
    } else
    sensor_normals_block2(rng, 0, z0);                                                     // z[0..7]
    // ---- accel_int :700-715
    const P2 drift_a = sm.ld(2);
    P2 ab = pfma(drift_a, dt, sm.ld(0));                                                   // accel() :613
    P2 acc1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) acc1[k] = pfma(sa, z0[k], padd(acc_read[k], ab));
    if (!zpre) sensor_normals_block2(rng, 1, z1);                                          // z[8..15]
    P2 rc[3];                                                                              // third column of the first TRIAD rotation
    {
        P2 Rm[9];
        {   // triad()
            ab = pfma(drift_a, dt, ab);
            P2 gb[3], mi[3], mb[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) gb[k] = pfma(ng, sm.ld(14 + k), pfma(sa, z0[3 + k], padd(acc_read[k], ab)));   // :658
            gb[2] = psub(gb[2], f_m);
            mi[0] = pfma(smg, z0[6], bc(p.s_mag[0])); mi[1] = pfma(smg, z0[7], bc(p.s_mag[1])); mi[2] = pfma(smg, z1[0], bc(p.s_mag[2]));
#pragma unroll
            for (int c = 0; c < 3; ++c) mb[c] = pfma(rot[c], mi[0], pfma(rot[3 + c], mi[1], pmul(rot[6 + c], mi[2])));   // :662
            triad2<true>(p, gb, mb, Rm);
        }
        P2 a_in[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) a_in[c] = pfma(Rm[c], acc1[0], pfma(Rm[3 + c], acc1[1], pmul(Rm[6 + c], acc1[2])));   // :705
        a_in[2] = padd(a_in[2], bc(p.g));
        rc[0] = Rm[2]; rc[1] = Rm[5]; rc[2] = Rm[8];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const P2 vel = pfma(a_in[k], dt, sm.ld(4 + k));                                // velocity :707
            const P2 pos = pfma(vel, dt, sm.ld(7 + k));                                    // position :708
            sm.stv(4 + k, vel); sm.stv(7 + k, pos); sm.stv(17 + k, a_in[k]);
        }
    }
    // ---- gyro_int :717-724
    const P2 drift_g = sm.ld(3);
    P2 gbias = pfma(drift_g, dt, sm.ld(1));                                                // gyro() :624
    P2 qg[4];
    {
        P2 w1[3], dq[4], q0[4];
#pragma unroll
        for (int k = 0; k < 3; ++k) w1[k] = pfma(sg, z1[1 + k], padd(y[10 + k], gbias));   // z[9..11]
#pragma unroll
        for (int k = 0; k < 4; ++k) q0[k] = sm.ld(10 + k);
        deriv_quat2(w1, q0, dq);
#pragma unroll
        for (int k = 0; k < 4; ++k) qg[k] = pfma(dq[k], dt, q0[k]);                        // :721-722 (returned un-normalised)
        const P2 inv = prsqrt(pfma(qg[0], qg[0], pfma(qg[1], qg[1], pfma(qg[2], qg[2], pmul(qg[3], qg[3])))));
#pragma unroll
        for (int k = 0; k < 4; ++k) { sm.stv(10 + k, pmul(qg[k], inv)); out(6 + k, qg[k]); }   // :723
    }
    // ---- gyro :622-628
    gbias = pfma(drift_g, dt, gbias);
    sm.stv(1, gbias);
    {
        P2 w2[3], qv[4];
#pragma unroll
        for (int k = 0; k < 3; ++k) w2[k] = pfma(sg, z1[4 + k], padd(y[10 + k], gbias));   // z[12..14]
        deriv_quat2(w2, qg, qv);                                                           // rl_worker.py:168
#pragma unroll
        for (int k = 0; k < 4; ++k) out(10 + k, qv[k]);
    }
    // ---- gps :642-647: z[15..20] = block 3, drawn only when the optional complementary blend reads them (math_trajectory.py:71-77)
    const P2 z21 = z1[7];                                                                  // P[15] = z[21]
    if (p.s_gps_blend > 0.f) {
        const P2 wg = bc(p.s_gps_blend * 0.01f), wa = bc((100.f - p.s_gps_blend) * 0.01f);
        sensor_normals_block2(rng, 3, z1);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const P2 pos_gps = pfma(bc(p.s_gps_p), z1[k], y[2 * k]);
            const P2 vel_gps = pfma(bc(p.s_gps_v), z1[3 + k], y[2 * k + 1]);
            sm.stv(7 + k, pfma(wa, sm.ld(7 + k), pmul(wg, pos_gps)));
            sm.stv(4 + k, pfma(wa, sm.ld(4 + k), pmul(wg, vel_gps)));
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { out(2 * k, sm.ld(7 + k)); out(2 * k + 1, sm.ld(4 + k)); }
    // ---- triad :649-697 (updates self.R for the next step)
    if (zpre) {
#pragma unroll
        for (int k = 0; k < 8; ++k) z0[k] = zpre[16 + k];
    } else
    sensor_normals_block2(rng, 2, z0);                                                     // P[16..23]: z[22..26] = P[16..20]
    {
        ab = pfma(drift_a, dt, ab);
        sm.stv(0, ab);
        P2 gb[3], mi[3], mb[3], R2[9];
        gb[0] = pfma(ng, rc[0], pfma(sa, z21, padd(acc_read[0], ab)));                     // z[21..23]
        gb[1] = pfma(ng, rc[1], pfma(sa, z0[0], padd(acc_read[1], ab)));
        gb[2] = psub(pfma(ng, rc[2], pfma(sa, z0[1], padd(acc_read[2], ab))), f_m);
#pragma unroll
        for (int k = 0; k < 3; ++k) mi[k] = pfma(smg, z0[2 + k], bc(p.s_mag[k]));          // z[24..26]
#pragma unroll
        for (int c = 0; c < 3; ++c) mb[c] = pfma(rot[c], mi[0], pfma(rot[3 + c], mi[1], pmul(rot[6 + c], mi[2])));
        triad2<false>(p, gb, mb, R2);
        sm.stv(14, R2[2]); sm.stv(15, R2[5]); sm.stv(16, R2[8]);
    }
}

}  // namespace qs
