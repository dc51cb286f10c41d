// quad_device.cuh — device-side building blocks of the batched quadrotor hot path (sm_100a).
//
// One CUDA thread advances one environment; the 13-float state, the rotor command and all RK stage
// vectors live in registers.  Everything is templated on the arithmetic type R (float = production
// mode, double = parity mode) and cites the reference lines it re-implements (paths relative to the
// reference root; scipy/ = SciPy's integrate/_ivp package the reference calls at quadrotor_env.py:483).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

namespace qs {

// --------------------------------------------------------------------------------------------
// TMA (1-D bulk async copy) + mbarrier primitives used to stage state / action tiles in shared memory
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA engine, SASS UBLKCP); bytes % 16 == 0, both addresses 16-byte aligned;
// completion is signalled on `bar` as transaction bytes.
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// --------------------------------------------------------------------------------------------
// math dispatch
// --------------------------------------------------------------------------------------------
template <typename R> struct M_;
__device__ __forceinline__ float fma_(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_(double a, double b, double c) { return ::fma(a, b, c); }
// FP32 production-mode math: branch-free, MUFU-based, a few ulp — far inside the FP32 parity bound
// (1e-4 relative + 1e-5 absolute) and ~3x fewer instructions than the IEEE-exact libdevice routines.
__device__ __forceinline__ float fast_sqrtf(float x) {            // max rel. error 2^-23 (PTX ISA)
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_rsqrtf(float x) {           // MUFU.RSQ without the denormal pre-scaling of rsqrtf()
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_log2f(float x) {            // MUFU.LG2 without the denormal pre-scaling of __log2f()
    float r;                                                      // (FSETP + FMUL 2^24 + FADD -24 around every call)
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_rcpf(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// atan on [0,1]: odd minimax polynomial (max error ~1.5 ulp), then octant reconstruction.
__device__ __forceinline__ float fast_atan2f(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float a = mn * fast_rcpf(mx);
    a = (mx == 0.f) ? 0.f : a;                                    // atan2(0,0) = 0 like numpy
    const float s = a * a;
    float r = 0.0027856871f;
    r = fmaf(r, s, -0.0158660002f);
    r = fmaf(r, s, 0.042472221f);
    r = fmaf(r, s, -0.0749753043f);
    r = fmaf(r, s, 0.106448799f);
    r = fmaf(r, s, -0.142070308f);
    r = fmaf(r, s, 0.199934542f);
    r = fmaf(r, s, -0.333331466f);
    r = r * s;
    r = fmaf(r, a, a);
    r = (ay > ax) ? (1.57079637f - r) : r;
    r = (x < 0.f) ? (3.14159274f - r) : r;
    return copysignf(r, y);
}
// asin on [-1,1]: |x|<=0.5 -> x + x^3 P(x^2); else pi/2 - 2 asin(sqrt((1-|x|)/2))  (Cephes asinf coefficients)
__device__ __forceinline__ float fast_asinf(float x) {
    const float ax = fabsf(x);
    const bool big = ax > 0.5f;
    const float z = big ? fmaf(-0.5f, ax, 0.5f) : ax * ax;
    const float s = big ? fast_sqrtf(z) : ax;
    float pz = 4.2163199048e-2f;
    pz = fmaf(pz, z, 2.4181311049e-2f);
    pz = fmaf(pz, z, 4.5470025998e-2f);
    pz = fmaf(pz, z, 7.4953002686e-2f);
    pz = fmaf(pz, z, 1.6666752422e-1f);
    float r = fmaf(s * z, pz, s);
    r = big ? fmaf(-2.f, r, 1.57079637f) : r;
    r = (ax > 1.f) ? __int_as_float(0x7fc00000) : r;              // NaN outside the domain, like asin
    return copysignf(r, x);
}

template <> struct M_<float> {
    static __device__ __forceinline__ float rsqrt(float x) { return fast_rsqrtf(x); }
    static __device__ __forceinline__ float sqrt(float x) { return fast_sqrtf(x); }
    static __device__ __forceinline__ float abs(float x) { return fabsf(x); }
    static __device__ __forceinline__ float atan2(float y, float x) { return fast_atan2f(y, x); }
    static __device__ __forceinline__ float asin(float x) { return fast_asinf(x); }
    static __device__ __forceinline__ float pow(float x, float y) { return powf(x, y); }
    static __device__ __forceinline__ float log(float x) { return __logf(x); }
    static __device__ __forceinline__ float fmin(float a, float b) { return fminf(a, b); }
    static __device__ __forceinline__ float fmax(float a, float b) { return fmaxf(a, b); }
    static __device__ __forceinline__ void sincos(float x, float* s, float* c) { __sincosf(x, s, c); }
    static __device__ __forceinline__ void sincospi(float x, float* s, float* c) { __sincosf(3.14159274f * x, s, c); }
    static __device__ __forceinline__ float next_up(float x) { return nextafterf(x, CUDART_INF_F); }
    static __device__ __forceinline__ float inf() { return CUDART_INF_F; }
};
template <> struct M_<double> {
    static __device__ __forceinline__ double rsqrt(double x) { return 1.0 / ::sqrt(x); }
    static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
    static __device__ __forceinline__ double abs(double x) { return ::fabs(x); }
    static __device__ __forceinline__ double atan2(double y, double x) { return ::atan2(y, x); }
    static __device__ __forceinline__ double asin(double x) { return ::asin(x); }
    static __device__ __forceinline__ double pow(double x, double y) { return ::pow(x, y); }
    static __device__ __forceinline__ double log(double x) { return ::log(x); }
    static __device__ __forceinline__ double fmin(double a, double b) { return ::fmin(a, b); }
    static __device__ __forceinline__ double fmax(double a, double b) { return ::fmax(a, b); }
    static __device__ __forceinline__ void sincos(double x, double* s, double* c) { ::sincos(x, s, c); }
    static __device__ __forceinline__ void sincospi(double x, double* s, double* c) { ::sincospi(x, s, c); }
    static __device__ __forceinline__ double next_up(double x) { return ::nextafter(x, CUDART_INF); }
    static __device__ __forceinline__ double inf() { return CUDART_INF; }
};

// --------------------------------------------------------------------------------------------
// per-handle constants, derived on the host in double from qs_params (quadrotor_env.py:30-80)
// --------------------------------------------------------------------------------------------
enum : uint32_t {
    F_DIRECT = 0x01u, F_CLIPPED = 0x02u, F_TRAINING = 0x04u, F_AUTO_RESET = 0x08u,
    F_SENSOR = 0x10u, F_AUX = 0x20u, F_ASYNC_RESET = 0x40u, F_ROBUST = 0x80u
};
enum : uint32_t { EF_DONE = 1u, EF_HAS_SHAPING = 2u, EF_SOLVED = 4u, EF_WARM_SHIFT = 3u, EF_LOW = 7u };

template <typename R> struct DevParams {
    // rotor map (f2F :247-272, f2w :197-245)
    R c8;            // T2WR*M*G/8
    R inv_kf;        // 1/K_F
    R arm;           // D
    R km_over_kf;    // K_M/K_F
    R i_r;           // I_R
    R k_f, k_m, dkf; // K_F, K_M, D*K_F
    R mix_f, mix_m, mix_z;   // 1/(4K_F), 1/(2 D K_F), 1/(4 K_M)
    R u_max;         // T2WR*M*G/4/K_F
    R effort_scale;  // K_F/(T2WR*M*G/4)*2
    // rigid body (drone_eq :274-406)
    R inv_m, g;
    R kd_m[3];       // 0.5*RHO*C_D*A_i / M
    R kdm_j[3];      // beam drag-moment coefficient / J_i   (:328-334 summed in closed form)
    R inv_j[3];
    R cross_j[3];    // (Jz-Jy)/Jx, (Jx-Jz)/Jy, (Jy-Jx)/Jz
    // stepping
    R dt, h_sub, inv_dt;   // env step, RK4 sub-interval, 1/dt
    R bb[9];         // bb_cond :139-143
    // reward (:511-573)
    R sh_v, sh_psi, sh_ang;          // shaping weights folded with SHAPING_WEIGHT/sum and the normalisers
    R tr_r[3], tr_e[3], tr_p[3];     // norm(ones(4)*TR_i), norm(ones(2)*TR_i*4), TR_P
    R p_c, target_state, solved_reward, broken_reward;
    R zero_control[4];
    // reset distribution (:439-445)
    R pos_clip, vel_clip, w_clip_lo, w_clip_hi;
    // sensor (:587-608)
    R s_accel_std, s_accel_drift, s_gyro_std, s_gyro_drift, s_mag_std, s_mag_drift, s_gps_p, s_gps_v;
    R s_gps_blend;   // GPS_P in per cent (visual_landing/math_trajectory.py:71-77); 0 = off
    R s_mag[3];      // magnetic field vector, mG (:651)
    R s_ti[9];       // inertial TRIAD basis t1i,t2i,t3i (:682-691), constant
    // robust_control (:84-109): perturbation scales and wind gusts
    R rb_kf, rb_m, rb_ir, rb_j[3], rb_gust_std[3];
    int32_t rb_gust_period;
    int32_t n_limit; // n + T  :157
    int32_t T;
    int32_t substeps;
    uint32_t flags;
};

// rotor command after the action map; constant across the RK stages of one env step
template <typename R> struct Ctrl {
    R f_m;           // F / M
    R tau_j[3];      // M_i / J_i
    R gyro_j[2];     // omega_r / Jx, omega_r / Jy
    // robust_control only (template flag ROBUST; never read otherwise, so they cost nothing in the production kernels)
    R wind[3];       // wind(i), added to the inertial velocity seen by the drag model  :318-320
    R sm;            // 1 / (1 + episode_m)                                             :360-361
    R sj[3];         // 1 / (1 + episode_J_ii)                                          :381-382
};

// robust_control (:84-109) for one env and one step: the per-episode perturbations and the wind of this step
template <typename R> struct Robust {
    R kf[4];         // episode_kf = U(0,1) * D_KF: rotor thrust loss                   :99,:236,:266
    R ir[4];         // 1 + episode_ir, episode_ir = U(0,1) * D_IR                      :101,:342
    R sm, sj[3], wind[3];
};

// --------------------------------------------------------------------------------------------
// quaternion / Euler utilities — environment/quaternion_euler_utility.py
// --------------------------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ void quat_normalize(const R q[4], R qn[4]) {
    R inv = M_<R>::rsqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    qn[0] = q[0] * inv; qn[1] = q[1] * inv; qn[2] = q[2] * inv; qn[3] = q[3] * inv;
}

// quat_rot_mat :71-80 (row-major r[3*i+j])
template <typename R>
__device__ __forceinline__ void quat_rot_mat(const R q[4], R r[9]) {
    R a = q[0], b = q[1], c = q[2], d = q[3];
    R aa = a * a, bb = b * b, cc = c * c, dd = d * d;
    R bc = b * c, ad = a * d, bd = b * d, ac = a * c, cd = c * d, ab = a * b;
    r[0] = aa + bb - cc - dd; r[1] = R(2) * (bc - ad);  r[2] = R(2) * (bd + ac);
    r[3] = R(2) * (bc + ad);  r[4] = aa - bb + cc - dd; r[5] = R(2) * (cd - ab);
    r[6] = R(2) * (bd - ac);  r[7] = R(2) * (cd + ab);  r[8] = aa - bb - cc + dd;
}

// deriv_quat :58-69
template <typename R>
__device__ __forceinline__ void deriv_quat(const R w[3], const R q[4], R dq[4]) {
    dq[0] = R(0.5) * (-w[0] * q[1] - w[1] * q[2] - w[2] * q[3]);
    dq[1] = R(0.5) * (w[0] * q[0] + w[2] * q[2] - w[1] * q[3]);
    dq[2] = R(0.5) * (w[1] * q[0] - w[2] * q[1] + w[0] * q[3]);
    dq[3] = R(0.5) * (w[2] * q[0] + w[1] * q[1] - w[0] * q[2]);
}

// quat_euler :39-48.  FP64 (parity mode): no asin clamp, NaN propagates exactly like the reference.
// FP32 (production mode): the argument is clamped to [-1,1] — rounding of a unit quaternion can push it one
// ulp past 1 near gimbal lock, and a NaN pitch would poison reward/done for the rest of the episode
// (SURVEY.md §5.3); with the clamp theta = +-pi/2 trips the bounding box (:500-509) instead.
template <typename R> __device__ __forceinline__ R asin_arg(R s) { return s; }
template <> __device__ __forceinline__ float asin_arg<float>(float s) { return fminf(fmaxf(s, -1.f), 1.f); }

template <typename R>
__device__ __forceinline__ void quat_euler(const R q[4], R ang[3]) {
    ang[0] = M_<R>::atan2(R(2) * (q[0] * q[1] + q[2] * q[3]), R(1) - R(2) * (q[1] * q[1] + q[2] * q[2]));
    ang[1] = M_<R>::asin(asin_arg<R>(R(2) * (q[0] * q[2] - q[3] * q[1])));
    ang[2] = M_<R>::atan2(R(2) * (q[0] * q[3] + q[1] * q[2]), R(1) - R(2) * (q[2] * q[2] + q[3] * q[3]));
}

// euler_quat :17-36
template <typename R>
__device__ __forceinline__ void euler_quat(const R ang[3], R q[4]) {
    R sp, cp, st, ct, sps, cps;
    M_<R>::sincos(ang[0] * R(0.5), &sp, &cp);
    M_<R>::sincos(ang[1] * R(0.5), &st, &ct);
    M_<R>::sincos(ang[2] * R(0.5), &sps, &cps);
    R t[4];
    t[0] = cp * ct * cps + sp * st * sps;
    t[1] = sp * ct * cps - cp * st * sps;
    t[2] = cp * st * cps + sp * ct * sps;
    t[3] = cp * ct * sps - sp * st * cps;
    quat_normalize(t, q);
}

// --------------------------------------------------------------------------------------------
// rotor maps
// --------------------------------------------------------------------------------------------
// f2F :247-272 (direct mode).  a[] already clipped to [-1,1] (:470).
template <typename R>
__device__ __forceinline__ void rotor_direct(const DevParams<R>& p, const R a[4], R w[4], R fm[4], const R* kf = nullptr) {
    R f0 = (a[0] + R(1)) * p.c8, f1 = (a[1] + R(1)) * p.c8, f2 = (a[2] + R(1)) * p.c8, f3 = (a[3] + R(1)) * p.c8;
    w[0] = M_<R>::sqrt(f0 * p.inv_kf); w[1] = M_<R>::sqrt(f1 * p.inv_kf);
    w[2] = M_<R>::sqrt(f2 * p.inv_kf); w[3] = M_<R>::sqrt(f3 * p.inv_kf);
    if (kf) { f0 = f0 - kf[0] * f0; f1 = f1 - kf[1] * f1; f2 = f2 - kf[2] * f2; f3 = f3 - kf[3] * f3; }   // robust :265-266
    fm[0] = f0 + f1 + f2 + f3;
    fm[1] = (f2 - f0) * p.arm;
    fm[2] = (f1 - f3) * p.arm;
    fm[3] = (-f0 + f1 - f2 + f3) * p.km_over_kf;
}

// f2w :197-245 (indirect mode): closed-form inverse of the 4x4 mixer the reference solves with LU.
template <typename R>
__device__ __forceinline__ void rotor_indirect(const DevParams<R>& p, bool clipped, const R fm_in[4],
                                               R effort[4], R w[4], R fm[4], const R* kf = nullptr) {
    R uf = fm_in[0] * p.mix_f, ux = fm_in[1] * p.mix_m, uy = fm_in[2] * p.mix_m, uz = fm_in[3] * p.mix_z;
    R u[4] = {uf - ux - uz, uf + uy + uz, uf + ux - uz, uf - uy + uz};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (clipped) {
            // np.clip(u, 0, max) == minimum(maximum(u, 0), max); NaN propagates
            R v = u[k];
            v = (v < R(0)) ? R(0) : v;
            v = (v > p.u_max) ? p.u_max : v;
            u[k] = v;
            w[k] = M_<R>::sqrt(v);
        } else {
            R s = (u[k] < R(0)) ? R(-1) : R(1);
            w[k] = M_<R>::sqrt(M_<R>::abs(u[k])) * s;
        }
        if (kf) u[k] = u[k] - u[k] * kf[k];                      // robust :235-236 (after w, before FM_new and step_effort)
        effort[k] = u[k] * p.effort_scale - R(1);
    }
    fm[0] = p.k_f * (u[0] + u[1] + u[2] + u[3]);
    fm[1] = p.dkf * (u[2] - u[0]);
    fm[2] = p.dkf * (u[1] - u[3]);
    fm[3] = p.k_m * (-u[0] + u[1] - u[2] + u[3]);
}

template <typename R>
__device__ __forceinline__ Ctrl<R> make_ctrl(const DevParams<R>& p, const R fm[4], const R w[4], const Robust<R>* rb = nullptr) {
    Ctrl<R> c;
    R omega_r = (-w[0] + w[1] - w[2] + w[3]) * p.i_r;          // :345
    if (rb) {                                                    // :341-343  ir_k = I_R (1 + episode_ir_k)
        omega_r = (-w[0] * rb->ir[0] + w[1] * rb->ir[1] - w[2] * rb->ir[2] + w[3] * rb->ir[3]) * p.i_r;
        c.sm = rb->sm;
#pragma unroll
        for (int k = 0; k < 3; ++k) { c.wind[k] = rb->wind[k]; c.sj[k] = rb->sj[k]; }
    }
    c.f_m = fm[0] * p.inv_m;
    c.tau_j[0] = fm[1] * p.inv_j[0]; c.tau_j[1] = fm[2] * p.inv_j[1]; c.tau_j[2] = fm[3] * p.inv_j[2];
    c.gyro_j[0] = omega_r * p.inv_j[0]; c.gyro_j[1] = omega_r * p.inv_j[1];
    return c;
}

// --------------------------------------------------------------------------------------------
// drone_eq :274-406 — RHS of the 13-state ODE.  y = [x,vx,y,vy,z,vz,q0..q3,wx,wy,wz]
// --------------------------------------------------------------------------------------------
// ROBUST (robust_control, :318-320,:360-361,:381-382): wind added to the velocity the drag model sees, body force divided by
// the perturbed mass, angular acceleration through the perturbed inertia (the w x Jw term keeps the nominal J, :378).
template <typename R, bool ROBUST = false>
__device__ __forceinline__ void drone_rhs(const DevParams<R>& p, const Ctrl<R>& c, const R y[13], R dy[13]) {
    R qn[4];
    quat_normalize(&y[6], qn);                                   // :311-312
    R r[9];
    quat_rot_mat(qn, r);                                         // :315
    R vx = y[1], vy = y[3], vz = y[5];
    R ux = vx, uy = vy, uz = vz;
    if (ROBUST) { ux += c.wind[0]; uy += c.wind[1]; uz += c.wind[2]; }
    R vbx = r[0] * ux + r[3] * uy + r[6] * uz;                   // :322  R^T v
    R vby = r[1] * ux + r[4] * uy + r[7] * uz;
    R vbz = r[2] * ux + r[5] * uy + r[8] * uz;
    R fx = -p.kd_m[0] * (M_<R>::abs(vbx) * vbx);                 // :323 (already / M)
    R fy = -p.kd_m[1] * (M_<R>::abs(vby) * vby);
    R fz = c.f_m - p.kd_m[2] * (M_<R>::abs(vbz) * vbz);          // :352-353
    if (ROBUST) { fx *= c.sm; fy *= c.sm; fz *= c.sm; }
    dy[0] = vx; dy[2] = vy; dy[4] = vz;
    dy[1] = r[0] * fx + r[1] * fy + r[2] * fz;                   // :357-367
    dy[3] = r[3] * fx + r[4] * fy + r[5] * fz;
    dy[5] = r[6] * fx + r[7] * fy + r[8] * fz - p.g;
    R wx = y[10], wy = y[11], wz = y[12];
    // m_in = m_action + m_gyro + m_drag - w x Jw  (:378), times J^-1 (:384-388, J diagonal)
    dy[10] = c.tau_j[0] - c.gyro_j[0] * wx - p.kdm_j[0] * (M_<R>::abs(wx) * wx) - p.cross_j[0] * (wy * wz);
    dy[11] = c.tau_j[1] + c.gyro_j[1] * wy - p.kdm_j[1] * (M_<R>::abs(wy) * wy) - p.cross_j[1] * (wx * wz);
    dy[12] = c.tau_j[2] - p.kdm_j[2] * (M_<R>::abs(wz) * wz) - p.cross_j[2] * (wx * wy);
    if (ROBUST) { dy[10] *= c.sj[0]; dy[11] *= c.sj[1]; dy[12] *= c.sj[2]; }
    deriv_quat(&y[10], qn, &dy[6]);                              // :392
}

// --------------------------------------------------------------------------------------------
// integrators
// --------------------------------------------------------------------------------------------
// Fixed-step classical RK4, S sub-intervals of length p.h_sub.  The four stages are a ROLLED loop (one copy of
// the RHS in the instruction stream instead of four): the kernel is instruction-issue bound and the unrolled
// form overflowed the instruction cache (ncu: stall_no_instructions) — the cost is 13 dead FMAs per substep.
template <typename R, bool ROBUST = false>
__device__ __forceinline__ void integrate_rk4(const DevParams<R>& p, const Ctrl<R>& c, R y[13]) {
    const R h = p.h_sub, hh = p.h_sub * R(0.5), h6 = p.h_sub * R(1.0 / 6.0);
    for (int s = 0; s < p.substeps; ++s) {
        R k[13], acc[13], yt[13];
#pragma unroll
        for (int j = 0; j < 13; ++j) { acc[j] = R(0); yt[j] = y[j]; }
#pragma unroll 1
        for (int st = 0; st < 4; ++st) {
            drone_rhs<R, ROBUST>(p, c, yt, k);
            const R wgt = (st == 0 || st == 3) ? R(1) : R(2);
            const R cc = (st < 2) ? hh : h;
#pragma unroll
            for (int j = 0; j < 13; ++j) { acc[j] = fma_(wgt, k[j], acc[j]); yt[j] = fma_(cc, k[j], y[j]); }
        }
#pragma unroll
        for (int j = 0; j < 13; ++j) y[j] = fma_(h6, acc[j], y[j]);
    }
}

// Dormand–Prince 5(4) tableau — scipy/integrate/_ivp/rk.py (class RK45: A, B, E)
#define QS_DP_A21 (1.0 / 5)
#define QS_DP_A31 (3.0 / 40)
#define QS_DP_A32 (9.0 / 40)
#define QS_DP_A41 (44.0 / 45)
#define QS_DP_A42 (-56.0 / 15)
#define QS_DP_A43 (32.0 / 9)
#define QS_DP_A51 (19372.0 / 6561)
#define QS_DP_A52 (-25360.0 / 2187)
#define QS_DP_A53 (64448.0 / 6561)
#define QS_DP_A54 (-212.0 / 729)
#define QS_DP_A61 (9017.0 / 3168)
#define QS_DP_A62 (-355.0 / 33)
#define QS_DP_A63 (46732.0 / 5247)
#define QS_DP_A64 (49.0 / 176)
#define QS_DP_A65 (-5103.0 / 18656)
#define QS_DP_B1 (35.0 / 384)
#define QS_DP_B3 (500.0 / 1113)
#define QS_DP_B4 (125.0 / 192)
#define QS_DP_B5 (-2187.0 / 6784)
#define QS_DP_B6 (11.0 / 84)
#define QS_DP_E1 (-71.0 / 57600)
#define QS_DP_E3 (71.0 / 16695)
#define QS_DP_E4 (-71.0 / 1920)
#define QS_DP_E5 (17253.0 / 339200)
#define QS_DP_E6 (-22.0 / 525)
#define QS_DP_E7 (1.0 / 40)

template <typename R>
__device__ __forceinline__ R rms13(const R v[13]) {           // scipy/.../common.py:63-65
    R s = R(0);
#pragma unroll
    for (int j = 0; j < 13; ++j) s += v[j] * v[j];
    return M_<R>::sqrt(s) / M_<R>::sqrt(R(13));
}

// Replica of solve_ivp(drone_eq, (0, t_bound), y0) with all defaults, as the reference calls it at
// quadrotor_env.py:483: RungeKutta.__init__ (rk.py:85-105), select_initial_step (common.py:68-134),
// _step_impl (rk.py:111-183), rk_step (rk.py:14-70), OdeSolver.step (base.py:179-210).
// Keeps only y(t_bound) like the caller (`self.y[:, -1]`, :485).  Returns the number of RHS calls.
template <typename R, bool ROBUST = false>
__device__ __noinline__ int integrate_rk45(const DevParams<R>& p, const Ctrl<R>& c, R y[13]) {
    const R rtol = R(1e-3), atol = R(1e-6);
    const R t_bound = p.dt;
    R t = R(0);
    R f[13], tmp[13], scale[13];
    drone_rhs<R, ROBUST>(p, c, y, f);
    int nfev = 1;
    // ---- select_initial_step
    R h_abs;
    {
#pragma unroll
        for (int j = 0; j < 13; ++j) scale[j] = atol + M_<R>::abs(y[j]) * rtol;
#pragma unroll
        for (int j = 0; j < 13; ++j) tmp[j] = y[j] / scale[j];
        R d0 = rms13(tmp);
#pragma unroll
        for (int j = 0; j < 13; ++j) tmp[j] = f[j] / scale[j];
        R d1 = rms13(tmp);
        R h0 = (d0 < R(1e-5) || d1 < R(1e-5)) ? R(1e-6) : R(0.01) * d0 / d1;
        h0 = M_<R>::fmin(h0, t_bound);
        R y1[13], f1[13];
#pragma unroll
        for (int j = 0; j < 13; ++j) y1[j] = y[j] + h0 * f[j];
        drone_rhs<R, ROBUST>(p, c, y1, f1);
        ++nfev;
#pragma unroll
        for (int j = 0; j < 13; ++j) tmp[j] = (f1[j] - f[j]) / scale[j];
        R d2 = rms13(tmp) / h0;
        R h1;
        if (d1 <= R(1e-15) && d2 <= R(1e-15)) h1 = M_<R>::fmax(R(1e-6), h0 * R(1e-3));
        else h1 = M_<R>::pow(R(0.01) / M_<R>::fmax(d1, d2), R(1.0 / 5));
        h_abs = M_<R>::fmin(M_<R>::fmin(R(100) * h0, h1), t_bound);
    }
    // ---- solver.step() until t >= t_bound
    // rk_step (rk.py:14-70) as ONE rolled stage loop: stage s = 1..5 evaluates K[s] at y + h * (A[s][0..s-1] . K[0..s-1]), "stage 6" is
    // y_new = y + h * (B . K[0..5]) with K[6] = f(y_new) (Dormand-Prince: the last row of A is B), so the RHS has a single call site and
    // the code of the loop body exists once (the unrolled form, six inlined copies of drone_eq, stalled on instruction fetch and kept
    // 3.2 KB of stage vectors and spills per thread in local memory).  Like SciPy's np.dot over K[:s], the zero coefficients take part.
    static constexpr double kA[6][6] = {
        {QS_DP_A21, 0, 0, 0, 0, 0},
        {QS_DP_A31, QS_DP_A32, 0, 0, 0, 0},
        {QS_DP_A41, QS_DP_A42, QS_DP_A43, 0, 0, 0},
        {QS_DP_A51, QS_DP_A52, QS_DP_A53, QS_DP_A54, 0, 0},
        {QS_DP_A61, QS_DP_A62, QS_DP_A63, QS_DP_A64, QS_DP_A65, 0},
        {QS_DP_B1, 0, QS_DP_B3, QS_DP_B4, QS_DP_B5, QS_DP_B6}};
    static constexpr double kE[7] = {QS_DP_E1, 0, QS_DP_E3, QS_DP_E4, QS_DP_E5, QS_DP_E6, QS_DP_E7};
    R K[7][13];                                  // K[0] = f(y) (first same as last), K[1..5] stages 2..6, K[6] = f(y_new)
    R yn[13];
#pragma unroll
    for (int j = 0; j < 13; ++j) {
        K[0][j] = f[j];
#pragma unroll
        for (int q = 1; q < 7; ++q) K[q][j] = R(0);            // rows a stage does not use yet are loaded (and not added)
    }
    for (int guard = 0; guard < 100000; ++guard) {
        R min_step = R(10) * M_<R>::abs(M_<R>::next_up(t) - t);
        if (h_abs < min_step) h_abs = min_step;
        bool rejected = false, accepted = false, failed = false;
        while (!accepted) {
            if (h_abs < min_step) { failed = true; break; }
            R t_new = t + h_abs;
            if (t_new - t_bound > R(0)) t_new = t_bound;
            R h = t_new - t;
            h_abs = M_<R>::abs(h);
#pragma unroll 1
            for (int s = 1; s <= 6; ++s) {
#pragma unroll
                for (int j = 0; j < 13; ++j) tmp[j] = K[0][j] * R(kA[s - 1][0]);
                // q unrolled with a uniform predicate: the stage vectors live in local memory and their loads must be in flight
                // together, not one row per trip of a rolled loop
#pragma unroll
                for (int q = 1; q < 6; ++q) {
                    const R a = R(kA[s - 1][q]);
                    const bool use = q < s;
#pragma unroll
                    for (int j = 0; j < 13; ++j) tmp[j] = use ? tmp[j] + K[q][j] * a : tmp[j];
                }
#pragma unroll
                for (int j = 0; j < 13; ++j) yn[j] = y[j] + tmp[j] * h;
                drone_rhs<R, ROBUST>(p, c, yn, K[s]);
            }
            nfev += 6;
#pragma unroll
            for (int j = 0; j < 13; ++j) tmp[j] = K[0][j] * R(kE[0]);
#pragma unroll
            for (int q = 1; q < 7; ++q) {
                const R eq = R(kE[q]);
#pragma unroll
                for (int j = 0; j < 13; ++j) tmp[j] += K[q][j] * eq;
            }
#pragma unroll
            for (int j = 0; j < 13; ++j) {
                R sc = atol + M_<R>::fmax(M_<R>::abs(y[j]), M_<R>::abs(yn[j])) * rtol;
                tmp[j] = tmp[j] * h / sc;
            }
            R err = rms13(tmp);
            if (err < R(1)) {
                R factor = (err == R(0)) ? R(10) : M_<R>::fmin(R(10), R(0.9) * M_<R>::pow(err, R(-0.2)));
                if (rejected) factor = M_<R>::fmin(R(1), factor);
                h_abs *= factor;
                accepted = true;
                t = t_new;
            } else if (err != err) {
                // NaN error norm: the reference would spin forever; leave the poisoned state and stop.
                accepted = true; failed = true; t = t_bound;
            } else {
                h_abs *= M_<R>::fmax(R(0.2), R(0.9) * M_<R>::pow(err, R(-0.2)));
                rejected = true;
            }
        }
        if (failed && !accepted) break;          // TOO_SMALL_STEP: solve_ivp stops, last accepted y is kept
#pragma unroll
        for (int j = 0; j < 13; ++j) { y[j] = yn[j]; K[0][j] = K[6][j]; }
        if (failed || t - t_bound >= R(0)) break;
    }
    return nfev;
}

// --------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) — counter-based RNG for resets / noise / action sampling
// --------------------------------------------------------------------------------------------
enum : uint32_t { RNG_RESET = 0, RNG_SENSOR = 1, RNG_ACTION = 2, RNG_POLICY = 3, RNG_ROBUST = 4, RNG_GUST = 5 };

#ifndef QS_PHILOX_UNROLL
#define QS_PHILOX_UNROLL 10
#endif
constexpr int kPhiloxUnroll = QS_PHILOX_UNROLL;
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll kPhiloxUnroll
    for (int r = 0; r < 10; ++r) {
        // (__umulhi + * = IMAD.HI + IMAD: measured faster than one mul.wide.u32 = IMAD.WIDE per word on sm_100)
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}

// Same function with the ten round keys (k + r * (0x9E3779B9, 0xBB67AE85)) precomputed on the host (SimView::rk): the key is
// the handle's seed, i.e. uniform over the launch, so the round keys can be constant-bank operands of the XORs instead of
// twenty live registers.
__device__ __forceinline__ uint4 philox4x32_10_rk(uint4 c, const uint32_t* __restrict__ rk) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ rk[2 * r], lo1, hi0 ^ c.w ^ rk[2 * r + 1], lo0);
    }
    return c;
}
__host__ __device__ inline void philox_round_keys(uint64_t seed, uint32_t rk[20]) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) { rk[2 * r] = k0; rk[2 * r + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
}

__device__ __forceinline__ uint4 philox_block(uint64_t seed, uint32_t env_id, uint32_t episode, uint32_t block,
                                              uint32_t stream) {
    return philox4x32_10(make_uint4(env_id, episode, block, stream),
                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

template <typename R> __device__ __forceinline__ R u32_to_unit(uint32_t u) {     // (u + 0.5) * 2^-32 in (0,1]
    return (R(u) + R(0.5)) * R(1.0 / 4294967296.0);
}

template <typename R> __device__ __forceinline__ void box_muller(R u1, R u2, R* n0, R* n1) {
    R r = M_<R>::sqrt(R(-2) * M_<R>::log(u1));
    R s, c;
    M_<R>::sincospi(R(2) * u2, &s, &c);
    *n0 = r * c; *n1 = r * s;
}
// FP32: MUFU sin/cos are most accurate on [-pi,pi] -> evaluate at 2*pi*(u2-1/2) and flip both signs.
template <> __device__ __forceinline__ void box_muller<float>(float u1, float u2, float* n0, float* n1) {
    float r = fast_sqrtf((-2.f * 0.693147182f) * fast_log2f(u1));     // u1 >= 2^-17: never denormal
    float s, c;
    __sincosf(6.28318548f * (u2 - 0.5f), &s, &c);
    *n0 = -r * c; *n1 = -r * s;
}

template <typename R> __device__ __forceinline__ R clampr(R v, R lo, R hi) {
    return M_<R>::fmin(M_<R>::fmax(v, lo), hi);
}

// Random branch of quad.reset (:439-445): ang ~ U(-.5,.5)^3, pos ~ clip(N(0,2),+-2.5),
// vel ~ clip(N(0,2),+-5), w ~ clip(N(0,2),-15,+7.5) (asymmetric, as in the reference).
template <typename R>
__device__ __forceinline__ void sample_reset_state(const DevParams<R>& p, uint64_t seed, uint32_t env_id,
                                                   uint32_t episode, R y[13], R ang[3]) {
    uint4 b0 = philox_block(seed, env_id, episode, 0, RNG_RESET);
    ang[0] = u32_to_unit<R>(b0.x) - R(0.5);
    ang[1] = u32_to_unit<R>(b0.y) - R(0.5);
    ang[2] = u32_to_unit<R>(b0.z) - R(0.5);
    R n[12];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        uint4 u = philox_block(seed, env_id, episode, b + 1, RNG_RESET);
        box_muller(u32_to_unit<R>(u.x), u32_to_unit<R>(u.y), &n[4 * b + 0], &n[4 * b + 1]);
        box_muller(u32_to_unit<R>(u.z), u32_to_unit<R>(u.w), &n[4 * b + 2], &n[4 * b + 3]);
    }
    y[0] = clampr(n[0] * R(2), -p.pos_clip, p.pos_clip);
    y[2] = clampr(n[1] * R(2), -p.pos_clip, p.pos_clip);
    y[4] = clampr(n[2] * R(2), -p.pos_clip, p.pos_clip);
    y[1] = clampr(n[3] * R(2), -p.vel_clip, p.vel_clip);
    y[3] = clampr(n[4] * R(2), -p.vel_clip, p.vel_clip);
    y[5] = clampr(n[5] * R(2), -p.vel_clip, p.vel_clip);
    euler_quat(ang, &y[6]);
    y[10] = clampr(n[6] * R(2), p.w_clip_lo, p.w_clip_hi);
    y[11] = clampr(n[7] * R(2), p.w_clip_lo, p.w_clip_hi);
    y[12] = clampr(n[8] * R(2), p.w_clip_lo, p.w_clip_hi);
}

// --------------------------------------------------------------------------------------------
// robust_control (:84-109): domain randomisation.  The reference draws from NumPy's global stream; here every draw is a
// pure function of (seed, global env id, episode / gust counter), so no perturbation has to be stored and any sharding
// reproduces it.  Per episode (robust_control.reset :98-102, called by quad.reset :426): episode_kf = U(0,1)^4 D_KF,
// episode_m = N(0, D_M), episode_ir = U(0,1)^4 D_IR, diag(episode_J) = N(0, D_J)^3  [Philox stream RNG_ROBUST of
// (env, episode): block 0 -> kf, block 1 -> ir, block 2 -> normals m, Jx, Jy, Jz].
// Wind (:104-109): a new gust N(0, gust_std) whenever i % gust_period == 1 (i = quad.i after the increment of this step, so
// every episode starts one), the wind ramping linearly from the previous gust to the new one over gust_period steps:
// np.linspace(last, gust, P)[(i % P) - 1].  The gust sequence of an env continues across episodes (the reference never
// resets it), numbered by the env's gust counter [stream RNG_GUST of (env, gust counter); gust 0 = 0].
// Deliberate deviation: the reference re-draws the gust at EVERY drone_eq call of the step in which i % P == 1 (8-14 draws
// in that step, each evaluation seeing the previous draw) — dead code there, not replicated: one draw per period.
// --------------------------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ void robust_gust(const DevParams<R>& p, uint64_t seed, uint32_t env_id, int32_t count, R g[3]) {
    if (count <= 0) { g[0] = g[1] = g[2] = R(0); return; }
    const uint4 u = philox_block(seed, env_id, (uint32_t)count, 0, RNG_GUST);
    R n[4];
    box_muller(u32_to_unit<R>(u.x), u32_to_unit<R>(u.y), &n[0], &n[1]);
    box_muller(u32_to_unit<R>(u.z), u32_to_unit<R>(u.w), &n[2], &n[3]);
#pragma unroll
    for (int k = 0; k < 3; ++k) g[k] = n[k] * p.rb_gust_std[k];
}

// i = quad.i of this step (after the increment); gust_count is the env's persistent gust counter (updated here)
template <typename R>
__device__ __forceinline__ void robust_prepare(const DevParams<R>& p, uint64_t seed, uint32_t env_id, uint32_t episode, int32_t i,
                                               int32_t& gust_count, Robust<R>& rb) {
    const uint4 a = philox_block(seed, env_id, episode, 0, RNG_ROBUST);
    const uint4 b = philox_block(seed, env_id, episode, 1, RNG_ROBUST);
    const uint4 c = philox_block(seed, env_id, episode, 2, RNG_ROBUST);
    rb.kf[0] = u32_to_unit<R>(a.x) * p.rb_kf; rb.kf[1] = u32_to_unit<R>(a.y) * p.rb_kf;
    rb.kf[2] = u32_to_unit<R>(a.z) * p.rb_kf; rb.kf[3] = u32_to_unit<R>(a.w) * p.rb_kf;
    rb.ir[0] = R(1) + u32_to_unit<R>(b.x) * p.rb_ir; rb.ir[1] = R(1) + u32_to_unit<R>(b.y) * p.rb_ir;
    rb.ir[2] = R(1) + u32_to_unit<R>(b.z) * p.rb_ir; rb.ir[3] = R(1) + u32_to_unit<R>(b.w) * p.rb_ir;
    R n[4];
    box_muller(u32_to_unit<R>(c.x), u32_to_unit<R>(c.y), &n[0], &n[1]);
    box_muller(u32_to_unit<R>(c.z), u32_to_unit<R>(c.w), &n[2], &n[3]);
    rb.sm = R(1) / (R(1) + n[0] * p.rb_m);
#pragma unroll
    for (int k = 0; k < 3; ++k) rb.sj[k] = R(1) / (R(1) + n[1 + k] * p.rb_j[k]);
    const int32_t P = p.rb_gust_period;
    const int32_t index = (i % P) - 1;                             // :105
    if (index == 0) gust_count += 1;                               // :106-108
    R g0[3], g1[3];
    robust_gust(p, seed, env_id, gust_count - 1, g0);
    robust_gust(p, seed, env_id, gust_count, g1);
    const R t = (index < 0) ? R(1) : R(index) / R(P - 1);          // np.linspace(last, gust, P)[index]; index -1 = the last element
#pragma unroll
    for (int k = 0; k < 3; ++k) rb.wind[k] = g0[k] + (g1[k] - g0[k]) * t;
}

// --------------------------------------------------------------------------------------------
// one environment, in registers
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ float div_dt(float x, const DevParams<float>& p) { return x * p.inv_dt; }
__device__ __forceinline__ double div_dt(double x, const DevParams<double>& p) { return x / p.dt; }

template <typename R> struct Env {
    R y[13];
    R prev_ang[3];       // quad.prev_ang — NOT cleared by reset (reference quirk, :171-172/:492-493)
    R prev_shaping;
    R abs_sum;
    R ep_return;
    int32_t i;
    uint32_t flags;      // EF_*
    uint32_t episode;
};

template <typename R> struct StepOut {
    R vq[4];             // V_q of the trailing drone_eq call (FSAL stage) :392,:486
    R ang[3];
    R ang_vel[3];
    R reward;
    R effort[4];
    R w[4];
    R clipped[4];        // quad.clipped_action :472,:477
    R fm[4];             // body thrust and moments applied
    bool done;           // value `quad.step` returns
    bool solved;
    bool broken, timeout;
};

// quad.step :458-498 for one env, in three phases so that kernels which integrate two envs per thread with the packed
// FP32 instructions (step_pair.cuh) share the scalar action map and the done/reward evaluation with every other kernel.
// `a_in` = action as given by the caller.
//
// phase 1 — :467-477: step counter, action clip, rotor map (f2F / f2w) -> rotor command held constant over the RK stages
template <typename R, bool DIRECT>
__device__ __forceinline__ Ctrl<R> step_pre(const DevParams<R>& p, Env<R>& e, const R a_in[4], StepOut<R>& o, R act[4],
                                            const Robust<R>* rb = nullptr) {
    e.i += 1;                                                    // :467
    R fm[4];
    if (DIRECT) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {                            // np.clip(action,-1,1) :470
            R v = a_in[k];
            v = (v < R(-1)) ? R(-1) : v;
            v = (v > R(1)) ? R(1) : v;
            act[k] = v; o.effort[k] = v;
        }
        rotor_direct(p, act, o.w, fm, rb ? rb->kf : nullptr);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) act[k] = a_in[k];            // reward uses the RAW action :476,:553
        rotor_indirect(p, (p.flags & F_CLIPPED) != 0, a_in, o.effort, o.w, fm, rb ? rb->kf : nullptr);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { o.fm[k] = fm[k]; o.clipped[k] = DIRECT ? act[k] : fm[k]; }
    return make_ctrl(p, fm, o.w, rb);
}

// phase 3 — :486-498 after the integration: observation tail, Euler angles, done_condition, reward_function, control_effort
template <typename R>
__device__ __forceinline__ void step_post(const DevParams<R>& p, Env<R>& e, const R act[4], StepOut<R>& o) {
    // observation tail: V_q = 1/2 Omega(w_new) normalize(q_new)  (:392 evaluated at the FSAL stage)
    R qn[4];
    quat_normalize(&e.y[6], qn);                                 // :488-489
    deriv_quat(&e.y[10], qn, o.vq);
    quat_euler(qn, o.ang);                                       // :491
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        o.ang_vel[k] = div_dt(o.ang[k] - e.prev_ang[k], p);      // :492
        e.prev_ang[k] = o.ang[k];                                // :493
    }
    // done_condition :500-509 (>=, sticky; NaN compares false).  Everything from here on is written without
    // short-circuit operators or if/else bodies so that it compiles to straight-line predicated code: kernels that
    // instantiate this phase for two envs per thread get both dependency chains into one basic block.
    bool done = (e.flags & EF_DONE) != 0;
    {
        const R cx[9] = {e.y[1], e.y[3], e.y[5], o.ang[0], o.ang[1], o.ang[2], e.y[10], e.y[11], e.y[12]};
#pragma unroll
        for (int k = 0; k < 9; ++k) done = done | (M_<R>::abs(cx[k]) >= p.bb[k]);
    }
    // reward_function :511-573
    R v2 = e.y[1] * e.y[1] + e.y[3] * e.y[3] + e.y[5] * e.y[5];
    R e2 = o.ang[0] * o.ang[0] + o.ang[1] * o.ang[1];
    R psi = o.ang[2];
    R nv = M_<R>::sqrt(v2), ne = M_<R>::sqrt(e2);
    R shaping = -(p.sh_v * nv + p.sh_psi * M_<R>::abs(psi) + p.sh_ang * ne);          // :529-531
    R nr = M_<R>::sqrt(v2 + psi * psi);
    {
        bool taken = false;
#pragma unroll
        for (int k = 0; k < 3; ++k) {                            // cascade :535-542
            const bool c1 = (!taken) & (nr < p.tr_r[k]);
            const bool c2 = c1 & (ne < p.tr_e[k]);
            shaping += c1 ? p.tr_p[k] : R(0);
            shaping += c2 ? p.tr_p[k] : R(0);
            taken = taken | c1;
        }
    }
    R reward = (e.flags & EF_HAS_SHAPING) ? (shaping - e.prev_shaping) : R(0);       // :545-547
    e.prev_shaping = shaping;
    R pen = R(0);
#pragma unroll
    for (int k = 0; k < 4; ++k) { R d = act[k] - p.zero_control[k]; pen += d * d; }
    reward += -pen * p.p_c;                                      // :553-554
    R cur = v2 + (e2 + psi * psi) + (e.y[10] * e.y[10] + e.y[11] * e.y[11] + e.y[12] * e.y[12]);   // :558
    // :562-573  precedence solved > time limit > broken
    const bool is_solved = cur < p.target_state;
    const bool timeout = (!is_solved) & (e.i >= p.n_limit);
    const bool broken = (!is_solved) & (!timeout) & done;
    reward = is_solved ? reward + p.solved_reward : (broken ? reward + p.broken_reward : reward);
    const bool solved = is_solved | (((e.flags & EF_SOLVED) != 0) & (!timeout) & (!broken));
    done = done | (is_solved & ((p.flags & F_TRAINING) != 0)) | timeout;
    e.flags = (e.flags & ~EF_LOW) | (done ? EF_DONE : 0u) | EF_HAS_SHAPING | (solved ? EF_SOLVED : 0u);
    e.abs_sum += M_<R>::sqrt(o.effort[0] * o.effort[0] + o.effort[1] * o.effort[1] + o.effort[2] * o.effort[2] +
                             o.effort[3] * o.effort[3]);         // :575-577
    o.reward = reward; o.done = done; o.solved = solved; o.broken = broken; o.timeout = timeout;
}

// ---- FP32 twin of phase 3 with the rounding of every operation written out.
// The pair kernels evaluate this phase on the packed FP32 pipe (step_post2, step_pair.cuh): fma.rn.f32x2 / mul.rn.f32x2 / add.rn.f32x2
// round each half exactly like the scalar fma.rn / mul.rn / add.rn.  The FP32 scalar phase below is the same sequence of operations
// as explicit intrinsics (the compiler neither contracts nor re-associates them), so from the same state the one-env-per-thread
// kernels and the pair kernels return bit-identical rewards, angles and flags; the per-env tail (thresholds, cascade, flag logic) is
// one function for both.  (The RK4 stages before it are not twinned: scalar drone_eq is contracted by the compiler.)
struct PostSums { float v2, e2, ne, nr2, cur, shaping0, pen_c, ef2; };

__device__ __forceinline__ void post_tail(const DevParams<float>& p, Env<float>& e, StepOut<float>& o, const float ang[3],
                                          const float vq[4], const PostSums& q) {
#pragma unroll
    for (int k = 0; k < 4; ++k) o.vq[k] = vq[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        o.ang[k] = ang[k];
        o.ang_vel[k] = div_dt(ang[k] - e.prev_ang[k], p);        // :492
        e.prev_ang[k] = ang[k];                                  // :493
    }
    bool done = (e.flags & EF_DONE) != 0;                        // done_condition :500-509 (>=, sticky; NaN compares false)
    const float cx9[9] = {e.y[1], e.y[3], e.y[5], ang[0], ang[1], ang[2], e.y[10], e.y[11], e.y[12]};
#pragma unroll
    for (int k = 0; k < 9; ++k) done = done | (fabsf(cx9[k]) >= p.bb[k]);
    const float nrh = fast_sqrtf(q.nr2), neh = q.ne;
    float shaping = q.shaping0;
    bool taken = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {                                // cascade :535-542
        const bool c1 = (!taken) & (nrh < p.tr_r[k]);
        const bool c2 = c1 & (neh < p.tr_e[k]);
        shaping += c1 ? p.tr_p[k] : 0.f;
        shaping += c2 ? p.tr_p[k] : 0.f;
        taken = taken | c1;
    }
    float reward = (e.flags & EF_HAS_SHAPING) ? (shaping - e.prev_shaping) : 0.f;     // :545-547
    e.prev_shaping = shaping;
    reward += q.pen_c;                                           // :553-554
    const bool is_solved = q.cur < p.target_state;               // :562-573  precedence solved > time limit > broken
    const bool timeout = (!is_solved) & (e.i >= p.n_limit);
    const bool broken = (!is_solved) & (!timeout) & done;
    reward = is_solved ? reward + p.solved_reward : (broken ? reward + p.broken_reward : reward);
    const bool solved = is_solved | (((e.flags & EF_SOLVED) != 0) & (!timeout) & (!broken));
    done = done | (is_solved & ((p.flags & F_TRAINING) != 0)) | timeout;
    e.flags = (e.flags & ~EF_LOW) | (done ? EF_DONE : 0u) | EF_HAS_SHAPING | (solved ? EF_SOLVED : 0u);
    e.abs_sum += fast_sqrtf(q.ef2);                              // :575-577
    o.reward = reward; o.done = done; o.solved = solved; o.broken = broken; o.timeout = timeout;
}

__device__ __forceinline__ void step_post(const DevParams<float>& p, Env<float>& e, const float act[4], StepOut<float>& o) {
    const float* y = e.y;
    auto fma_ = [](float a, float b, float c) { return __fmaf_rn(a, b, c); };
    auto mul_ = [](float a, float b) { return __fmul_rn(a, b); };
    auto add_ = [](float a, float b) { return __fadd_rn(a, b); };
    const float inv = fast_rsqrtf(fma_(y[6], y[6], fma_(y[7], y[7], fma_(y[8], y[8], mul_(y[9], y[9])))));          // :488-489
    const float q[4] = {mul_(y[6], inv), mul_(y[7], inv), mul_(y[8], inv), mul_(y[9], inv)};
    float vq[4];
    {   // deriv_quat :58-69 in the operation order of deriv_quat2 (sensor_pair.cuh)                                :392
        const float hx = mul_(y[10], 0.5f), hy = mul_(y[11], 0.5f), hz = mul_(y[12], 0.5f);
        const float nx = mul_(y[10], -0.5f), ny = mul_(y[11], -0.5f), nz = mul_(y[12], -0.5f);
        vq[0] = fma_(nx, q[1], fma_(ny, q[2], mul_(nz, q[3])));
        vq[1] = fma_(hx, q[0], fma_(hz, q[2], mul_(ny, q[3])));
        vq[2] = fma_(hy, q[0], fma_(nz, q[1], mul_(hx, q[3])));
        vq[3] = fma_(hz, q[0], fma_(hy, q[1], mul_(nx, q[2])));
    }
    // quat_euler utility:39-48
    const float sx = mul_(2.f, fma_(q[0], q[1], mul_(q[2], q[3])));
    const float cx = fma_(-2.f, fma_(q[1], q[1], mul_(q[2], q[2])), 1.f);
    const float sy = mul_(2.f, fma_(q[0], q[2], mul_(mul_(q[3], -1.f), q[1])));
    const float sz = mul_(2.f, fma_(q[0], q[3], mul_(q[1], q[2])));
    const float cz = fma_(-2.f, fma_(q[2], q[2], mul_(q[3], q[3])), 1.f);
    const float ang[3] = {fast_atan2f(sx, cx), fast_asinf(asin_arg<float>(sy)), fast_atan2f(sz, cz)};
    // reward_function :511-573, the sums
    PostSums r;
    r.v2 = fma_(y[1], y[1], fma_(y[3], y[3], mul_(y[5], y[5])));
    r.e2 = fma_(ang[0], ang[0], mul_(ang[1], ang[1]));
    const float psi2 = mul_(ang[2], ang[2]);
    const float w2 = fma_(y[10], y[10], fma_(y[11], y[11], mul_(y[12], y[12])));
    r.cur = add_(r.v2, add_(add_(r.e2, psi2), w2));                                                               // :558
    r.nr2 = add_(r.v2, psi2);
    float pen = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float d = add_(act[k], -p.zero_control[k]); pen = fma_(d, d, pen); }
    r.ef2 = fma_(o.effort[0], o.effort[0], fma_(o.effort[1], o.effort[1], fma_(o.effort[2], o.effort[2], mul_(o.effort[3], o.effort[3]))));
    const float nv = fast_sqrtf(r.v2);
    r.ne = fast_sqrtf(r.e2);
    r.shaping0 = mul_(-1.f, fma_(p.sh_v, nv, fma_(p.sh_psi, fabsf(ang[2]), mul_(p.sh_ang, r.ne))));                // :529-531
    r.pen_c = mul_(pen, -p.p_c);
    post_tail(p, e, o, ang, vq, r);
}

// robust_control context of an env (ROBUST kernels only): where its Philox streams and its gust counter live
struct RobustCtx { uint64_t seed; uint32_t env_id; int32_t* gust_count; };

template <typename R, int INTEG, bool DIRECT, bool ROBUST = false>
__device__ __forceinline__ void step_core(const DevParams<R>& p, Env<R>& e, const R a_in[4], StepOut<R>& o,
                                          Ctrl<R>* ctrl_out = nullptr, const RobustCtx* rc = nullptr) {
    R act[4];
    Robust<R> rb;
    if (ROBUST) {
        int32_t gc = *rc->gust_count;
        robust_prepare(p, rc->seed, rc->env_id, e.episode, e.i + 1, gc, rb);
        *rc->gust_count = gc;
    }
    const Ctrl<R> c = step_pre<R, DIRECT>(p, e, a_in, o, act, ROBUST ? &rb : nullptr);
    if (ctrl_out) *ctrl_out = c;
    if (INTEG == 1) integrate_rk45<R, ROBUST>(p, c, e.y);        // :483
    else integrate_rk4<R, ROBUST>(p, c, e.y);
    step_post(p, e, act, o);
}

// QS_FLAG_ASYNC_RESET, start of a step: an env that still owes warm-up steps gets the neutral action (:448)
// for this step.  Returns true if this step is a warm-up step.
template <typename R>
__device__ __forceinline__ bool async_warmup_prologue(const DevParams<R>& p, Env<R>& e, R a[4]) {
    const bool warm = (e.flags >> EF_WARM_SHIFT) != 0u;
    e.flags -= warm ? (1u << EF_WARM_SHIFT) : 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] = warm ? p.zero_control[k] : a[k];
    return warm;
}

// QS_FLAG_ASYNC_RESET, end of the step that returned done: begin the next episode — Philox-sampled initial
// state (random branch of quad.reset :439-445), bookkeeping cleared (:428-433), T warm-up steps owed.
// vq receives V_q of the initial state so that the observation rows are the new episode's first observation.
template <typename R>
__device__ __forceinline__ void async_resample(const DevParams<R>& p, uint64_t seed, uint32_t env_id, Env<R>& e, R vq[4]) {
    e.episode += 1;
    R ang[3];
    sample_reset_state(p, seed, env_id, e.episode, e.y, ang);
    e.flags = (uint32_t)p.T << EF_WARM_SHIFT;          // solved=0, done=False, prev_shaping=None
    e.i = 0;
    e.abs_sum = R(0);
    e.ep_return = R(0);
    deriv_quat(&e.y[10], &e.y[6], vq);                 // euler_quat returns a unit quaternion
}

// head of quad.reset :428-438 for one env (state already chosen); the T warm-up steps follow in the caller
template <typename R>
__device__ __forceinline__ void reset_head(Env<R>& e) {
    e.flags = 0u;            // solved=0, done=False, prev_shaping=None
    e.i = 0;
    e.abs_sum = R(0);
    e.ep_return = R(0);
}

}  // namespace qs
