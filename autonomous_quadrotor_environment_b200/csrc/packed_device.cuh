// packed_device.cuh — drone_eq and the RK4 integration for TWO environments per thread on the sm_100 packed-FP32
// pipe (PTX fma/mul/add.f32x2, SASS FFMA2/FMUL2/FADD2: one issue slot, two IEEE-rn FP32 results).
//
// Why: the scalar step kernel is instruction-issue bound (profiles/r01_prof_step_warp_*.txt: ~1260 warp-instructions per
// 32 env-steps, FMA pipe 39 % busy).  A lane that owns the env pair (A, B) keeps every quantity as one aligned 64-bit
// register pair {A, B}; the arithmetic of drone_eq then needs half the issue slots per env, and every scalar
// per-env phase around it (action map, Euler angles, done/reward) sees two independent dependency chains.
//
// The formulation is written out operation by operation (no compiler contraction involved): same mathematics as
// drone_rhs<float> (quad_device.cuh), signs folded into the constants so that only fma/mul/add are needed.
#pragma once
#include "quad_device.cuh"

namespace qs {

struct P2 { float2 v; };     // .x = env A, .y = env B

__device__ __forceinline__ P2 pk(float a, float b) { return P2{make_float2(a, b)}; }
__device__ __forceinline__ P2 bc(float c) { return P2{make_float2(c, c)}; }      // FFMA2 takes a scalar register / immediate broadcast
__device__ __forceinline__ P2 pfma(P2 a, P2 b, P2 c) { return P2{__ffma2_rn(a.v, b.v, c.v)}; }
__device__ __forceinline__ P2 pmul(P2 a, P2 b) { return P2{__fmul2_rn(a.v, b.v)}; }
__device__ __forceinline__ P2 padd(P2 a, P2 b) { return P2{__fadd2_rn(a.v, b.v)}; }
__device__ __forceinline__ P2 prsqrt(P2 a) { return pk(fast_rsqrtf(a.v.x), fast_rsqrtf(a.v.y)); }          // MUFU.RSQ is scalar
__device__ __forceinline__ P2 pabsmul(P2 a) { return pk(fabsf(a.v.x) * a.v.x, fabsf(a.v.y) * a.v.y); }   // |x| is a free operand modifier of the scalar FMUL

// rotor command of an env pair; signs folded: g0 = -omega_r/Jx, g1 = +omega_r/Jy  (drone_eq :345,:378)
struct Ctrl2 { P2 f_m, tau[3], g0, g1; };

__device__ __forceinline__ Ctrl2 pack_ctrl(const Ctrl<float>& a, const Ctrl<float>& b) {
    Ctrl2 c;
    c.f_m = pk(a.f_m, b.f_m);
#pragma unroll
    for (int k = 0; k < 3; ++k) c.tau[k] = pk(a.tau_j[k], b.tau_j[k]);
    c.g0 = pk(-a.gyro_j[0], -b.gyro_j[0]);
    c.g1 = pk(a.gyro_j[1], b.gyro_j[1]);
    return c;
}

// drone_eq :274-406 for an env pair.  y = [x,vx,y,vy,z,vz,q0..q3,wx,wy,wz]; dy[0], dy[2], dy[4] are NOT written (they are
// y[1], y[3], y[5]: the caller reads them from the stage state).  rot (nullable at compile time) receives quat_rot_mat.
template <bool WANT_ROT>
__device__ __forceinline__ void drone_rhs2(const DevParams<float>& p, const Ctrl2& c, const P2 y[13], P2 dy[13], P2* rot) {
    // q^ = q/|q|  :311-312
    const P2 n2 = pfma(y[6], y[6], pfma(y[7], y[7], pfma(y[8], y[8], pmul(y[9], y[9]))));
    const P2 inv = prsqrt(n2);
    const P2 a = pmul(y[6], inv), b = pmul(y[7], inv), cq = pmul(y[8], inv), d = pmul(y[9], inv);
    // quat_rot_mat utility:71-80 with |q^| = 1:  r_ii = 1 - 2(.. + ..), r_ij = 2(.. +- ..)
    const P2 one = bc(1.f), m2 = bc(-2.f);
    const P2 a2 = padd(a, a), b2 = padd(b, b), c2 = padd(cq, cq);
    const P2 nb2 = pmul(b, m2), nc2 = pmul(cq, m2), nd2 = pmul(d, m2);
    const P2 tb = pfma(nb2, b, one);
    const P2 r0 = pfma(nd2, d, pfma(nc2, cq, one)), r4 = pfma(nd2, d, tb), r8 = pfma(nc2, cq, tb);
    const P2 bc2 = pmul(b2, cq), bd2 = pmul(b2, d), cd2 = pmul(c2, d);
    const P2 r1 = pfma(nd2, a, bc2), r3 = pfma(a2, d, bc2);
    const P2 r2 = pfma(a2, cq, bd2), r6 = pfma(nc2, a, bd2);
    const P2 r5 = pfma(nb2, a, cd2), r7 = pfma(a2, b, cd2);
    if (WANT_ROT) { rot[0] = r0; rot[1] = r1; rot[2] = r2; rot[3] = r3; rot[4] = r4; rot[5] = r5; rot[6] = r6; rot[7] = r7; rot[8] = r8; }
    // body velocity R^T v :322, quadratic drag :323 (already / M), thrust :352-353
    const P2 vx = y[1], vy = y[3], vz = y[5];
    const P2 vbx = pfma(r0, vx, pfma(r3, vy, pmul(r6, vz)));
    const P2 vby = pfma(r1, vx, pfma(r4, vy, pmul(r7, vz)));
    const P2 vbz = pfma(r2, vx, pfma(r5, vy, pmul(r8, vz)));
    const P2 fx = pmul(bc(-p.kd_m[0]), pabsmul(vbx));
    const P2 fy = pmul(bc(-p.kd_m[1]), pabsmul(vby));
    const P2 fz = pfma(bc(-p.kd_m[2]), pabsmul(vbz), c.f_m);
    dy[1] = pfma(r0, fx, pfma(r1, fy, pmul(r2, fz)));            // :357-367
    dy[3] = pfma(r3, fx, pfma(r4, fy, pmul(r5, fz)));
    dy[5] = pfma(r6, fx, pfma(r7, fy, pfma(r8, fz, bc(-p.g))));
    // J^-1 (m_action + m_gyro + m_drag - w x Jw)  :378-388
    const P2 wx = y[10], wy = y[11], wz = y[12];
    dy[10] = pfma(bc(-p.cross_j[0]), pmul(wy, wz), pfma(bc(-p.kdm_j[0]), pabsmul(wx), pfma(c.g0, wx, c.tau[0])));
    dy[11] = pfma(bc(-p.cross_j[1]), pmul(wx, wz), pfma(bc(-p.kdm_j[1]), pabsmul(wy), pfma(c.g1, wy, c.tau[1])));
    dy[12] = pfma(bc(-p.cross_j[2]), pmul(wx, wy), pfma(bc(-p.kdm_j[2]), pabsmul(wz), c.tau[2]));
    // deriv_quat utility:58-69 on q^
    const P2 hf = bc(0.5f), nhf = bc(-0.5f);
    const P2 hx = pmul(wx, hf), hy = pmul(wy, hf), hz = pmul(wz, hf);
    const P2 nhx = pmul(wx, nhf), nhy = pmul(wy, nhf), nhz = pmul(wz, nhf);
    dy[6] = pfma(nhx, b, pfma(nhy, cq, pmul(nhz, d)));
    dy[7] = pfma(hx, a, pfma(hz, cq, pmul(nhy, d)));
    dy[8] = pfma(hy, a, pfma(nhz, b, pmul(hx, d)));
    dy[9] = pfma(hz, a, pfma(hy, b, pmul(nhx, cq)));
}

// Classical RK4, p.substeps sub-intervals; rolled stage loop as in integrate_rk4 (one copy of the RHS in the
// instruction stream).  Position rows advance with the stage velocities (dy[0,2,4] = yt[1,3,5]).
__device__ __forceinline__ void integrate_rk4_2(const DevParams<float>& p, const Ctrl2& c, P2 y[13]) {
    const float h = p.h_sub, hh = p.h_sub * 0.5f, h6 = p.h_sub * (1.0f / 6.0f);
    for (int s = 0; s < p.substeps; ++s) {
        P2 k[13], acc[13], yt[13];
#pragma unroll
        for (int j = 0; j < 13; ++j) { acc[j] = bc(0.f); yt[j] = y[j]; }
#pragma unroll 1
        for (int st = 0; st < 4; ++st) {
            drone_rhs2<false>(p, c, yt, k, nullptr);
            k[0] = yt[1]; k[2] = yt[3]; k[4] = yt[5];
            const P2 wgt = bc((st == 0 || st == 3) ? 1.f : 2.f);
            const P2 cc = bc((st < 2) ? hh : h);
#pragma unroll
            for (int j = 0; j < 13; ++j) { acc[j] = pfma(wgt, k[j], acc[j]); yt[j] = pfma(cc, k[j], y[j]); }
        }
#pragma unroll
        for (int j = 0; j < 13; ++j) y[j] = pfma(bc(h6), acc[j], y[j]);
    }
}

// The same integration with the four stages UNROLLED and a caller-provided piece of independent work placed after each stage
// (`between(st)`, st a compile-time constant after unrolling): everything lands in one basic block, so that the instruction
// scheduler can interleave that work — the integer Philox rounds and the MUFU chains of the sensor model's normal draws — with
// the FP32 stage arithmetic inside ONE warp (the kernels run 8 warps per SM: too few to hide those latencies across warps).
// First sub-interval only; further sub-intervals (S > 1) use the rolled loop.
template <typename F>
__device__ __forceinline__ void integrate_rk4_2_fused(const DevParams<float>& p, const Ctrl2& c, P2 y[13], F&& between) {
    const float h = p.h_sub, hh = p.h_sub * 0.5f, h6 = p.h_sub * (1.0f / 6.0f);
    {
        P2 k[13], acc[13], yt[13];
#pragma unroll
        for (int j = 0; j < 13; ++j) { acc[j] = bc(0.f); yt[j] = y[j]; }
#pragma unroll
        for (int st = 0; st < 4; ++st) {
            drone_rhs2<false>(p, c, yt, k, nullptr);
            k[0] = yt[1]; k[2] = yt[3]; k[4] = yt[5];
            const P2 wgt = bc((st == 0 || st == 3) ? 1.f : 2.f);
            const P2 cc = bc((st < 2) ? hh : h);
#pragma unroll
            for (int j = 0; j < 13; ++j) { acc[j] = pfma(wgt, k[j], acc[j]); yt[j] = pfma(cc, k[j], y[j]); }
            between(st);
        }
#pragma unroll
        for (int j = 0; j < 13; ++j) y[j] = pfma(bc(h6), acc[j], y[j]);
    }
    for (int s = 1; s < p.substeps; ++s) {
        P2 k[13], acc[13], yt[13];
#pragma unroll
        for (int j = 0; j < 13; ++j) { acc[j] = bc(0.f); yt[j] = y[j]; }
#pragma unroll 1
        for (int st = 0; st < 4; ++st) {
            drone_rhs2<false>(p, c, yt, k, nullptr);
            k[0] = yt[1]; k[2] = yt[3]; k[4] = yt[5];
            const P2 wgt = bc((st == 0 || st == 3) ? 1.f : 2.f);
            const P2 cc = bc((st < 2) ? hh : h);
#pragma unroll
            for (int j = 0; j < 13; ++j) { acc[j] = pfma(wgt, k[j], acc[j]); yt[j] = pfma(cc, k[j], y[j]); }
        }
#pragma unroll
        for (int j = 0; j < 13; ++j) y[j] = pfma(bc(h6), acc[j], y[j]);
    }
}

}  // namespace qs
