// rk4_packed_probe.cu — developer experiment: the RK4 integration loop of drone_eq (environment/quadrotor_env.py:274-406)
// written once over a value type V and instantiated for V = float (one env per thread) and V = float2 lanes issued as the
// sm_100 packed-FP32 instructions FFMA2/FMUL2/FADD2 (two envs per thread).  Register-resident, no memory traffic in the
// loop: isolates the question "do the packed instructions relieve the issue-slot bound of the step kernel?".
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/rk4_packed_probe tools/rk4_packed_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

struct P2 { float2 v; };

__device__ __forceinline__ float vfma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ float vmul(float a, float b) { return a * b; }
__device__ __forceinline__ float vadd(float a, float b) { return a + b; }
__device__ __forceinline__ float vsub(float a, float b) { return a - b; }
__device__ __forceinline__ float vrsqrt(float a) { return rsqrtf(a); }
__device__ __forceinline__ float vabsmul(float a) { return fabsf(a) * a; }
__device__ __forceinline__ void vset(float& d, float c) { d = c; }

__device__ __forceinline__ P2 vfma(P2 a, P2 b, P2 c) { return P2{__ffma2_rn(a.v, b.v, c.v)}; }
__device__ __forceinline__ P2 vmul(P2 a, P2 b) { return P2{__fmul2_rn(a.v, b.v)}; }
__device__ __forceinline__ P2 vadd(P2 a, P2 b) { return P2{__fadd2_rn(a.v, b.v)}; }
__device__ __forceinline__ P2 vsub(P2 a, P2 b) { return P2{__ffma2_rn(b.v, make_float2(-1.f, -1.f), a.v)}; }
__device__ __forceinline__ P2 vrsqrt(P2 a) { return P2{make_float2(rsqrtf(a.v.x), rsqrtf(a.v.y))}; }
__device__ __forceinline__ P2 vabsmul(P2 a) { return P2{make_float2(fabsf(a.v.x) * a.v.x, fabsf(a.v.y) * a.v.y)}; }
__device__ __forceinline__ void vset(P2& d, float c) { d.v = make_float2(c, c); }

template <class V> struct K {      // constants, broadcast to V
    V kd[3], g, kdm[3], cj[3], half, one, two;
};

// ctrl c[6] = F/M, M_i/J_i (3), omega_r/Jx, omega_r/Jy
template <class V>
__device__ __forceinline__ void rhs(const K<V>& k, const V* c, const V* y, V* dy) {
    V n2 = vfma(y[6], y[6], vfma(y[7], y[7], vfma(y[8], y[8], vmul(y[9], y[9]))));
    V inv = vrsqrt(n2);
    V a = vmul(y[6], inv), b = vmul(y[7], inv), cq = vmul(y[8], inv), d = vmul(y[9], inv);
    V a2 = vadd(a, a), b2 = vadd(b, b), c2 = vadd(cq, cq);
    V ad = vmul(a2, d), bc = vmul(b2, cq), ac = vmul(a2, cq), bd = vmul(b2, d), ab = vmul(a2, b), cd = vmul(c2, d);
    V bb = vmul(b2, b), cc = vmul(c2, cq), dd = vmul(vadd(d, d), d);
    V r0 = vsub(vsub(k.one, cc), dd), r4 = vsub(vsub(k.one, bb), dd), r8 = vsub(vsub(k.one, bb), cc);
    V r1 = vsub(bc, ad), r3 = vadd(bc, ad), r2 = vadd(bd, ac), r6 = vsub(bd, ac), r5 = vsub(cd, ab), r7 = vadd(cd, ab);
    V vx = y[1], vy = y[3], vz = y[5];
    V vbx = vfma(r0, vx, vfma(r3, vy, vmul(r6, vz)));
    V vby = vfma(r1, vx, vfma(r4, vy, vmul(r7, vz)));
    V vbz = vfma(r2, vx, vfma(r5, vy, vmul(r8, vz)));
    V fx = vmul(k.kd[0], vabsmul(vbx)), fy = vmul(k.kd[1], vabsmul(vby));        // kd = -0.5 rho C_D A / M
    V fz = vfma(k.kd[2], vabsmul(vbz), c[0]);
    dy[0] = vx; dy[2] = vy; dy[4] = vz;
    dy[1] = vfma(r0, fx, vfma(r1, fy, vmul(r2, fz)));
    dy[3] = vfma(r3, fx, vfma(r4, fy, vmul(r5, fz)));
    dy[5] = vfma(r6, fx, vfma(r7, fy, vfma(r8, fz, k.g)));                        // g = -G
    V wx = y[10], wy = y[11], wz = y[12];
    dy[10] = vfma(k.cj[0], vmul(wy, wz), vfma(k.kdm[0], vabsmul(wx), vfma(c[4], wx, c[1])));   // c[4] = -omega_r/Jx
    dy[11] = vfma(k.cj[1], vmul(wx, wz), vfma(k.kdm[1], vabsmul(wy), vfma(c[5], wy, c[2])));
    dy[12] = vfma(k.cj[2], vmul(wx, wy), vfma(k.kdm[2], vabsmul(wz), c[3]));
    V hx = vmul(wx, k.half), hy = vmul(wy, k.half), hz = vmul(wz, k.half);
    dy[6] = vsub(k.one, k.one);                      // placeholder overwritten below (keeps V default-constructible)
    V t0 = vfma(hx, b, vfma(hy, cq, vmul(hz, d)));
    dy[6] = vsub(dy[6], t0);
    dy[7] = vfma(hx, a, vsub(vmul(hz, cq), vmul(hy, d)));
    dy[8] = vfma(hy, a, vsub(vmul(hx, d), vmul(hz, b)));
    dy[9] = vfma(hz, a, vsub(vmul(hy, b), vmul(hx, cq)));
}

template <class V>
__device__ __forceinline__ void rk4(const K<V>& k, const V* c, V* y, V h, V hh, V h6) {
    V kk[13], acc[13], yt[13];
#pragma unroll
    for (int j = 0; j < 13; ++j) { acc[j] = vsub(y[j], y[j]); yt[j] = y[j]; }
#pragma unroll 1
    for (int st = 0; st < 4; ++st) {
        rhs(k, c, yt, kk);
        V wgt, cc;
        vset(wgt, (st == 0 || st == 3) ? 1.f : 2.f);
        cc = (st < 2) ? hh : h;
#pragma unroll
        for (int j = 0; j < 13; ++j) { acc[j] = vfma(wgt, kk[j], acc[j]); yt[j] = vfma(cc, kk[j], y[j]); }
    }
#pragma unroll
    for (int j = 0; j < 13; ++j) y[j] = vfma(h6, acc[j], y[j]);
}

template <class V> __device__ __forceinline__ K<V> make_k() {
    K<V> k;
    const float M = 1.03f, rho = 1.2041f, cd = 1.1f;
    const float A[3] = {0.026f, 0.026f, 0.052f}, J[3] = {16.83e-3f, 16.83e-3f, 28.34e-3f};
    const float kdm = 8.406517802222225e-05f;
    for (int i = 0; i < 3; ++i) vset(k.kd[i], -0.5f * rho * cd * A[i] / M);
    vset(k.kdm[0], -kdm / J[0]); vset(k.kdm[1], -kdm / J[1]); vset(k.kdm[2], -2 * kdm / J[2]);
    vset(k.cj[0], -(J[2] - J[1]) / J[0]); vset(k.cj[1], -(J[0] - J[2]) / J[1]); vset(k.cj[2], -(J[1] - J[0]) / J[2]);
    vset(k.g, -9.82f); vset(k.half, 0.5f); vset(k.one, 1.f); vset(k.two, 2.f);
    return k;
}

// state [13][n], ctrl [6][n]
__global__ void __launch_bounds__(256, 2) k_scalar(int64_t n, int steps, const float* __restrict__ st, const float* __restrict__ ct, float* __restrict__ out) {
    const K<float> k = make_k<float>();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float y[13], c[6];
#pragma unroll
        for (int j = 0; j < 13; ++j) y[j] = st[j * n + i];
#pragma unroll
        for (int j = 0; j < 6; ++j) c[j] = ct[j * n + i];
        for (int s = 0; s < steps; ++s) rk4<float>(k, c, y, 0.01f, 0.005f, 0.01f / 6);
#pragma unroll
        for (int j = 0; j < 13; ++j) out[j * n + i] = y[j];
    }
}

template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_packed(int64_t n, int steps, const float* __restrict__ st, const float* __restrict__ ct, float* __restrict__ out) {
    const K<P2> k = make_k<P2>();
    P2 h, hh, h6;
    vset(h, 0.01f); vset(hh, 0.005f); vset(h6, 0.01f / 6);
    const int64_t n2 = n / 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
        P2 y[13], c[6];
#pragma unroll
        for (int j = 0; j < 13; ++j) y[j].v = *reinterpret_cast<const float2*>(st + j * n + 2 * i);
#pragma unroll
        for (int j = 0; j < 6; ++j) c[j].v = *reinterpret_cast<const float2*>(ct + j * n + 2 * i);
        for (int s = 0; s < steps; ++s) rk4<P2>(k, c, y, h, hh, h6);
#pragma unroll
        for (int j = 0; j < 13; ++j) *reinterpret_cast<float2*>(out + j * n + 2 * i) = y[j].v;
    }
}

int main(int argc, char** argv) {
    const int64_t n = 1 << 20;
    const int steps = argc > 1 ? atoi(argv[1]) : 32;
    std::vector<float> hs(13 * n), hc(6 * n);
    srand(1);
    auto U = [] { return rand() / (float)RAND_MAX * 2 - 1; };
    for (int64_t i = 0; i < n; ++i) {
        for (int j = 0; j < 13; ++j) hs[j * n + i] = U();
        hs[6 * n + i] = 1.f + 0.1f * U();
        hc[0 * n + i] = 9.82f + U();
        for (int j = 1; j < 6; ++j) hc[j * n + i] = U();
    }
    float *st, *ct, *o1, *o2;
    cudaMalloc(&st, 13 * n * 4); cudaMalloc(&ct, 6 * n * 4); cudaMalloc(&o1, 13 * n * 4); cudaMalloc(&o2, 13 * n * 4);
    cudaMemcpy(st, hs.data(), 13 * n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(ct, hc.data(), 6 * n * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](auto launch, const char* name) {
        launch();
        cudaEventRecord(e0);
        for (int r = 0; r < 5; ++r) launch();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
        printf("%-28s %8.3f ms  %7.2f us per RK4 step of 1M envs  %.3e env-substeps/s\n", name, ms, ms * 1e3 / steps, n * (double)steps / (ms * 1e-3));
    };
    time([&] { k_scalar<<<148 * 2, 256>>>(n, steps, st, ct, o1); }, "scalar  (2 CTA/SM)");
    time([&] { k_scalar<<<148 * 8, 256>>>(n, steps, st, ct, o1); }, "scalar  (grid 8/SM)");
    time([&] { k_packed<1><<<148 * 1, 256>>>(n, steps, st, ct, o2); }, "packed  (1 CTA/SM, <=255 reg)");
    time([&] { k_packed<2><<<148 * 2, 256>>>(n, steps, st, ct, o2); }, "packed  (2 CTA/SM, <=128 reg)");
    std::vector<float> a(13 * n), b(13 * n);
    cudaMemcpy(a.data(), o1, 13 * n * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(b.data(), o2, 13 * n * 4, cudaMemcpyDeviceToHost);
    double worst = 0; int64_t bad = 0;
    for (int64_t i = 0; i < 13 * n; ++i) {
        if (!std::isfinite(a[i]) || !std::isfinite(b[i])) { ++bad; continue; }
        double e = fabs((double)a[i] - b[i]) / (1e-5 + 1e-4 * fabs((double)a[i]));
        if (e > worst) worst = e;
    }
    printf("scalar vs packed: worst err/bound %.3g, non-finite %lld; %s\n", worst, (long long)bad, cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
