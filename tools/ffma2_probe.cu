// ffma2_probe.cu — developer micro-benchmark: is the packed FP32 pipe (fma.rn.f32x2 / SASS FFMA2, sm_100+) worth
// restructuring the step kernel around two environments per thread?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2_probe tools/ffma2_probe.cu && /tmp/ffma2_probe
// Cases (8 independent chains per thread, 1184 CTAs x 256 threads):
//   scalar      : FFMA only                       -> FLOP/s of the scalar pipe (the roof bench.py quotes)
//   packed      : FFMA2 only                      -> FLOP/s of the packed form
//   scalar+alu  : 1 FFMA : 1 integer LOP3/IADD    -> issue-slot contention (what the step kernel looks like)
//   packed+alu  : 1 FFMA2 : 2 integer ops         -> same FLOPs and ALU work as scalar+alu in fewer issue slots
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    uint64_t ra = *reinterpret_cast<uint64_t*>(&a), rb = *reinterpret_cast<uint64_t*>(&b), rc = *reinterpret_cast<uint64_t*>(&c), rd;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float seed) {
    float a = seed, b = 0.999f;
    float s[16];
    float2 v[8];
    uint32_t u[8];
#pragma unroll
    for (int k = 0; k < 16; ++k) s[k] = threadIdx.x * 1e-3f + k;
#pragma unroll
    for (int k = 0; k < 8; ++k) { v[k] = make_float2(s[k], s[k + 8]); u[k] = threadIdx.x + k; }
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < 16; ++k) s[k] = fmaf(s[k], b, a);
        } else if (MODE == 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = ffma2(v[k], b2, a2);
        } else if (MODE == 2) {
#pragma unroll
            for (int k = 0; k < 16; ++k) { s[k] = fmaf(s[k], b, a); u[k & 7] = (u[k & 7] ^ (uint32_t)i) + 0x9E3779B9u * (k + 1); }
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                v[k] = ffma2(v[k], b2, a2);
                u[k] = (u[k] ^ (uint32_t)i) + 0x9E3779B9u * (k + 1);
                u[(k + 1) & 7] = (u[(k + 1) & 7] ^ (uint32_t)i) + 0x9E3779B9u * (k + 9);
            }
        }
    }
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) r += s[k];
#pragma unroll
    for (int k = 0; k < 8; ++k) r += v[k].x + v[k].y + __uint_as_float(u[k] & 0x3fffffffu);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE> void run(const char* name, float* out) {
    const int blocks = 148 * 8, threads = 256, iters = 1 << 13;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<blocks, threads>>>(out, iters, 0.5f);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) probe<MODE><<<blocks, threads>>>(out, iters, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double fmas = 16.0 * iters * blocks * threads;
    printf("%-12s %8.3f ms  %7.2f TFLOP/s (FP32 FMA = 2 FLOP)\n", name, ms, 2 * fmas / (ms * 1e-3) / 1e12);
}

int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    run<0>("scalar", out); run<1>("packed", out); run<2>("scalar+alu", out); run<3>("packed+alu", out);
    cudaError_t err = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(err));
    return err != cudaSuccess;
}
