// MUFU throughput probe (B200): tanh.approx.f32 vs tanh.approx.f16 (scalar) vs tanh.approx.f16x2 / bf16x2, ex2.approx.f32 as a reference.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_probe tools/mufu_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(unsigned* out, int iters) {
    unsigned x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = 0x3c003800u + threadIdx.x + j * 977u;      // f16x2 / f32 bit patterns, values stay bounded under tanh
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+r"(x[j]));
            if (MODE == 1) asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %0; tanh.approx.f16 lo, lo; mov.b32 %0, {lo, hi};}" : "+r"(x[j]));
            if (MODE == 2) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(x[j]));
            if (MODE == 3) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(x[j]));
            if (MODE == 4) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(x[j]));
            if (MODE == 5) asm volatile("cvt.rn.bf16x2.f32 %0, %0, %1;" : "+r"(x[j]) : "r"(x[(j + 1) & 7]));        // F2FP.BF16.F32.PACK_AB
            if (MODE == 6) asm volatile("{.reg .b32 t; add.u32 t, %0, 0x8000; prmt.b32 %0, t, %1, 0x7632;}" : "+r"(x[j]) : "r"(x[(j + 1) & 7]));   // integer pack
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s ^= x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int per_instr) {
    unsigned* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 4096;
    k<MODE><<<148 * 8, 256>>>(d, 16); cudaDeviceSynchronize();
    cudaEventRecord(a); k<MODE><<<148 * 8, 256>>>(d, iters); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double ops = 148.0 * 8 * 256 * iters * 8 * per_instr;
    printf("%-22s %8.3f ms  %7.1f G results/s  (%.2f results per clock per SM at 1.93 GHz)\n", name, ms, ops / ms / 1e6, ops / (ms * 1e-3) / 148 / 1.93e9);
    cudaFree(d);
}
int main() {
    run<0>("tanh.approx.f32", 1); run<1>("tanh.approx.f16", 1); run<2>("tanh.approx.f16x2", 2); run<3>("tanh.approx.bf16x2", 2); run<4>("ex2.approx.ftz.f32", 1);
    run<5>("cvt.rn.bf16x2.f32", 1); run<6>("add + prmt pack", 1);
    return 0;
}
