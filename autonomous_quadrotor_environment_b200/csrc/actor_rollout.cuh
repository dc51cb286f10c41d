// actor_rollout.cuh — the reference's actor MLP (environment/controller/model.py:27-34: Linear(75,H)-Tanh-Linear(H,H)-
// Tanh-Linear(H,4)-Tanh) and its observation-history input (environment/controller/dl_auxiliary.py:15-32) fused into
// the K-step rollout: per group of 128 threads 128 envs = one UMMA M=128 tile; the two hidden layers are tcgen05.mma (BF16
// operands staged in shared memory in the canonical K-major layout, FP32 accumulators in TMEM, biases folded into the
// GEMMs), the first epilogue reads TMEM with tcgen05.ld, applies tanh and writes the next layer's A operand back to shared
// memory, the second one applies tanh and the 128 -> 4 output layer on the FP32 pipe (weights in constant memory); the
// dynamics run on the same 128 threads (thread = env = TMEM lane) with the env state in registers.
// Own translation unit (actor_rollout.cu).
#pragma once
#include "umma.cuh"

// ---------------------------------------------------------------------------------------------------------------
// self-test: D[128][N] = A[128][K] * B[N][K]^T through the same operand layout / descriptors / TMEM path
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_umma_selftest(int N, int K, const float* A, const float* B, float* D) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sA = smem;
    unsigned char* sB = smem + 128 * K * 2;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int k = 0; k < K; k += 8) {
        uint4 v;
        v.x = pack_bf16x2(A[tid * K + k + 0], A[tid * K + k + 1]);
        v.y = pack_bf16x2(A[tid * K + k + 2], A[tid * K + k + 3]);
        v.z = pack_bf16x2(A[tid * K + k + 4], A[tid * K + k + 5]);
        v.w = pack_bf16x2(A[tid * K + k + 6], A[tid * K + k + 7]);
        *reinterpret_cast<uint4*>(sA + umma_canon_offset(tid, k, K)) = v;
        if (tid < N) {
            v.x = pack_bf16x2(B[tid * K + k + 0], B[tid * K + k + 1]);
            v.y = pack_bf16x2(B[tid * K + k + 2], B[tid * K + k + 3]);
            v.z = pack_bf16x2(B[tid * K + k + 4], B[tid * K + k + 5]);
            v.w = pack_bf16x2(B[tid * K + k + 6], B[tid * K + k + 7]);
            *reinterpret_cast<uint4*>(sB + umma_canon_offset(tid, k, K)) = v;
        }
    }
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_base_slot, 128);
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    if (tid == 0) {
        umma_gemm_k(tmem_base, smem_u32(sA), K, 0, smem_u32(sB), K, 0, K, N, false);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < N; c += 16) {
        float v[16];
        tmem_ld_32x32b_x16(lane_addr + (uint32_t)c, v);
        for (int i = 0; i < 16; ++i) D[tid * N + c + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 128);
}

// the same product with the A operand staged in TENSOR memory (thread = row writes its packed BF16 row with tcgen05.st)
__global__ void __launch_bounds__(128) k_umma_selftest_ts(int N, int K, const float* A, const float* B, float* D) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sB = smem;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int k = 0; k < K; k += 8) {
        if (tid < N) {
            uint4 v;
            v.x = pack_bf16x2(B[tid * K + k + 0], B[tid * K + k + 1]);
            v.y = pack_bf16x2(B[tid * K + k + 2], B[tid * K + k + 3]);
            v.z = pack_bf16x2(B[tid * K + k + 4], B[tid * K + k + 5]);
            v.w = pack_bf16x2(B[tid * K + k + 6], B[tid * K + k + 7]);
            *reinterpret_cast<uint4*>(sB + umma_canon_offset(tid, k, K)) = v;
        }
    }
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_base_slot, 256);
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t a_col = 128;                                      // A operand: columns [128, 128 + K/2)
    for (int k = 0; k < K; k += 32) {
        uint32_t r[16];
        for (int i = 0; i < 16; ++i) {
            const int kk = k + 2 * i;
            r[i] = kk < K ? pack_bf16x2(A[tid * K + kk], A[tid * K + kk + 1]) : 0u;
        }
        tmem_st_32x32b_x16(lane_addr + a_col + (uint32_t)(k >> 1), r);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        const uint32_t idesc = umma_idesc_bf16_f32(128, N);
        for (int k = 0; k < K; k += 16) {
            const uint64_t db = umma_smem_desc(smem_u32(sB) + (uint32_t)(k >> 3) * 128u, 128u, (uint32_t)(K >> 3) * 128u);
            umma_bf16_ts(tmem_base, tmem_base + a_col + (uint32_t)(k >> 1), db, idesc, k > 0);
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c = 0; c < N; c += 16) {
        float v[16];
        tmem_ld_32x32b_x16(lane_addr + (uint32_t)c, v);
        for (int i = 0; i < 16; ++i) D[tid * N + c + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

extern "C" int qs_umma_selftest_ts(int N, int K, const float* A, const float* B, float* D, void* stream) {
    if (N < 16 || N > 128 || (N % 16) || K < 32 || K > 128 || (K % 32) || !A || !B || !D)
        return fail(QS_EINVAL, "qs_umma_selftest_ts: need 16 <= N <= 128 (multiple of 16), 32 <= K <= 128 (multiple of 32)");
    const size_t smem = (size_t)N * K * 2;
    QS_CUDA(cudaFuncSetAttribute(k_umma_selftest_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_umma_selftest_ts<<<1, 128, smem, (cudaStream_t)stream>>>(N, K, A, B, D);
    QS_CUDA(cudaGetLastError());
    return QS_OK;
}

extern "C" int qs_umma_selftest(int N, int K, const float* A, const float* B, float* D, void* stream) {
    if (N < 16 || N > 128 || (N % 16) || K < 16 || K > 128 || (K % 16) || !A || !B || !D)
        return fail(QS_EINVAL, "qs_umma_selftest: need 16 <= N,K <= 128, multiples of 16");
    const size_t smem = (size_t)(128 + N) * K * 2;
    QS_CUDA(cudaFuncSetAttribute(k_umma_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_umma_selftest<<<1, 128, smem, (cudaStream_t)stream>>>(N, K, A, B, D);
    QS_CUDA(cudaGetLastError());
    return QS_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// fused policy rollout (BASELINE.json configs[4]: 1M envs x 128-step horizon with the actor MLP on tensor cores)
// ---------------------------------------------------------------------------------------------------------------
constexpr int kPM = 128;        // envs per CTA = UMMA M
constexpr int kPH = 128;        // hidden width of the shipped actors (solved/nn_*_128_*.pth)
constexpr int kPSlots = 5;      // history length T of dl_in_gen (ppo.py:306, T = 5)
constexpr int kPSlotK = 16;     // 15 floats per history entry [action(4), v(3), q(4), dq(4)] padded to one UMMA K block
constexpr int kPKin = kPSlots * kPSlotK;   // 80

struct ActorView {
    const float *w1, *b1, *w2, *b2, *w3, *b3;   // PyTorch Linear layout [out][in]: (128,75) (128,128) (4,128)
    float action_std;                            // sigma of the fixed-std Normal policy (model.py:62); <= 0: deterministic
    const float *cw1, *cb1, *cw2, *cb2;          // critic (model.py:36-43) hidden layers, same shapes; NULL = no critic
};
struct PolicyIO {
    int32_t horizon;
    float* obs_out;       // [K][14][N] or NULL
    float* action_out;    // [K][4][N]  or NULL  (the sampled, unclipped action PPO stores, model.py:64-68)
    float* logprob_out;   // [K][4][N]  or NULL  (per-dimension log-prob, model.py:66)
    float* reward_out;    // [K][N]     or NULL
    uint8_t* done_out;    // [K][N]     or NULL  (bit0 done, bit1 warm-up step)
    float* hist;          // [75][N] in/out: dl_in_gen.deep_learning_input per env (oldest entry first), or NULL
    float* value_out;     // [K+1][N] or NULL (critic handles): V of the network input of step t; row K = V of the input after the last step
    float* sensed_out;    // [K][14][N] or NULL (QS_FLAG_SENSOR_NOISE handles): the sensed observation the history takes
};

#ifndef QS_POLICY_GROUPS
#define QS_POLICY_GROUPS 3      // 128-env tiles in flight per CTA (one per group of 4 warps), sharing one copy of the weights
#endif
#ifndef QS_POLICY_TS            // 1 (default): hidden activations stay in TENSOR memory (TS form of tcgen05.mma), 4 tiles in flight
#define QS_POLICY_TS 1
#endif

// Two data paths for the hidden activations H1 = tanh(X W1^T + b1), selected at compile time:
//   TS = false  H1 is packed to BF16 into a 32 KB shared-memory tile per group (A operand of layer 2 from shared memory);
//               with the 20 KB history tile that is 52 KB per group: THREE tiles in flight per SM next to the weights.
//   TS = true   H1 never leaves tensor memory: the epilogue writes the packed BF16 row back with tcgen05.st IN PLACE over the
//               first 64 of the 128 accumulator columns it has just read, and layer 2 takes its A operand from there.  Layer 2
//               runs as two N = 64 halves into the other 64 columns (the output layer is fused into its epilogue, so nothing
//               of H2 is stored): 128 TMEM columns and 20 KB of shared memory per group -> FOUR tiles in flight per SM
//               (512 TMEM columns, 140 KB), which is what the kernel needs: every tile is a serial chain MMA -> tanh -> MMA ->
//               tanh -> dynamics and the SM is only busy while other tiles fill its gaps.
template <bool TS, bool CRITIC = false> struct PolicyCfg {
    static constexpr int kG = TS ? 4 : QS_POLICY_GROUPS;                 // tiles (groups of 128 threads) per CTA
    static constexpr int kXH = kPM * kPKin * 2 + (TS ? 0 : kPM * kPH * 2);   // per group: history/A tile (+ hidden tile)
    static constexpr int kX = 0;
    static constexpr int kHd = kPM * kPKin * 2;
    static constexpr int kW1 = kG * kXH;
    static constexpr int kW2 = kW1 + kPH * kPKin * 2;
    static constexpr int kOnes = kW2 + kPH * kPH * 2;                    // A operand of the bias block of layer 2: [128][16], columns 0,1 = 1
    static constexpr int kW2x = kOnes + kPM * 16 * 2;                    // B operand of that block: [128][16], columns 0,1 = b2 (hi, lo)
    static constexpr int kW1c = kW2x + kPH * 16 * 2;                     // CRITIC: the critic's W1 / W2 / b2 operand tiles (one copy per CTA)
    static constexpr int kW2c = kW1c + kPH * kPKin * 2;
    static constexpr int kW2xc = kW2c + kPH * kPH * 2;
    static constexpr int kBytes = CRITIC ? kW2xc + kPH * 16 * 2 : kW1c;
};

// Output layer (Linear(128,4) + Tanh, model.py:33-34) on the FP32 pipe: 128 -> 4 is 512 FMAs per env, and as a third UMMA it
// cost a full issue -> commit -> mbarrier round trip (~700 cycles, 5 % of the step: ncu long_scoreboard on the wait) plus a
// BF16 pack + 16 STS.128 per thread to stage its A operand.  The weights sit in constant memory (FFMA takes them as
// c[bank][imm] operands, no load instruction) and the FMAs issue in the shadow of the MUFU-bound tanh sequence of the second
// hidden epilogue; the hidden activations enter in FP32, not rounded to BF16.
// The symbol is written by a stream-ordered device-to-device copy in qs_policy_rollout: rollouts with DIFFERENT actors must
// not run concurrently on different streams of one device.
__constant__ float c_actor_w3[kPH * 4];      // [j][k] = w3[k][j]
__constant__ float c_actor_b3[4];
__constant__ float c_critic_w3[kPH];         // critic output layer Linear(128,1) (model.py:42)
__constant__ float c_critic_b3[1];

__global__ void k_pack_w3(const float* __restrict__ w3, const float* __restrict__ b3, float* __restrict__ out) {
    const int j = threadIdx.x;               // out: [128][4] then b3[4]
    if (j < kPH) { for (int k = 0; k < 4; ++k) out[j * 4 + k] = w3[k * kPH + j]; }
    if (j < 4) out[kPH * 4 + j] = b3[j];
}

// The biases of the two hidden layers ride in the GEMMs: every history entry is 15 floats padded to one K=16 block, the pad
// column of the A operand holds 1 and the pad columns of W1's first two K blocks hold b1 split in two BF16 pieces
// (hi = bf16(b), lo = bf16(b - hi): 16 significant bits); layer 2 gets one extra K=16 block [1 1 0 ..] x [b2_hi b2_lo 0 ..].
// That removes 128 FADD + 32 LDS.128 per thread from each hidden epilogue (the kernel is issue- and MUFU-bound, the tensor
// pipe is 26 % busy).
__device__ __forceinline__ float bf16_hi(float x) { return __bfloat162float(__float2bfloat16(x)); }

// barrier among the 128 threads of one group (ids 1..kPG; 0 is __syncthreads)
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(kPM) : "memory"); }

// tanh epilogue of one hidden layer: TMEM accumulators (thread = row; bias already inside) -> BF16 A operand of the next layer
__device__ __forceinline__ void actor_hidden_epilogue(uint32_t lane_addr, unsigned char* sH, int row) {
#pragma unroll 1
    for (int c = 0; c < kPH; c += 32) {
        float acc[32];
        tmem_ld_32x32b_x32(lane_addr + (uint32_t)c, acc);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_bf16x2(tanh_fast(acc[8 * j + 0]), tanh_fast(acc[8 * j + 1]));
            o.y = pack_bf16x2(tanh_fast(acc[8 * j + 2]), tanh_fast(acc[8 * j + 3]));
            o.z = pack_bf16x2(tanh_fast(acc[8 * j + 4]), tanh_fast(acc[8 * j + 5]));
            o.w = pack_bf16x2(tanh_fast(acc[8 * j + 6]), tanh_fast(acc[8 * j + 7]));
            *reinterpret_cast<uint4*>(sH + umma_canon_offset(row, c + 8 * j, kPH)) = o;
        }
    }
}

// tanh epilogue of the second hidden layer fused with the output layer: m[k] += sum_j tanh(acc_j) w3[k][j] over the hidden
// units [J0, J0 + NJ) whose accumulators sit in the NJ columns starting at taddr; the caller starts from b3 and applies tanh
template <int J0, int NJ>
__device__ __forceinline__ void actor_output_accumulate(uint32_t taddr, float m[4]) {
    float2 m01 = make_float2(m[0], m[1]), m23 = make_float2(m[2], m[3]);
#pragma unroll
    for (int c = 0; c < NJ; c += 32) {
        float acc[32];
        tmem_ld_32x32b_x32(taddr + (uint32_t)c, acc);
#pragma unroll
        for (int i = 0; i < 32; ++i) {                                // two packed FMAs per hidden unit: h x (w0,w1), h x (w2,w3);
            const float h = tanh_fast(acc[i]);                       // the weight pairs arrive in uniform registers (LDCU.128)
            const float2 hh = make_float2(h, h);
            const float* w = &c_actor_w3[(J0 + c + i) * 4];
            m01 = __ffma2_rn(hh, make_float2(w[0], w[1]), m01);
            m23 = __ffma2_rn(hh, make_float2(w[2], w[3]), m23);
        }
    }
    m[0] = m01.x; m[1] = m01.y; m[2] = m23.x; m[3] = m23.y;
}

// critic: tanh epilogue of the second hidden layer fused with Linear(128,1): v += sum_j tanh(acc_j) w3c[j]
template <int J0, int NJ>
__device__ __forceinline__ float critic_output_accumulate(uint32_t taddr, float vacc) {
    float v0 = vacc, v1 = 0.f;
#pragma unroll
    for (int c = 0; c < NJ; c += 32) {
        float acc[32];
        tmem_ld_32x32b_x32(taddr + (uint32_t)c, acc);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            v0 = fmaf(tanh_fast(acc[i]), c_critic_w3[J0 + c + i], v0);
            v1 = fmaf(tanh_fast(acc[i + 1]), c_critic_w3[J0 + c + i + 1], v1);
        }
    }
    return v0 + v1;
}

// tanh epilogue of the first hidden layer, TS path: accumulator columns [0,128) -> packed BF16 row in columns [0,64), in place
// (chunk c reads columns [32c, 32c+32) and writes [16c, 16c+16): always columns this thread has already consumed)
__device__ __forceinline__ void actor_hidden_epilogue_tmem(uint32_t lane_addr) {
#pragma unroll 1
    for (int c = 0; c < kPH; c += 32) {
        float acc[32];
        tmem_ld_32x32b_x32(lane_addr + (uint32_t)c, acc);
        uint32_t o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = pack_bf16x2(tanh_fast(acc[2 * i]), tanh_fast(acc[2 * i + 1]));
        tmem_st_32x32b_x16(lane_addr + (uint32_t)(c >> 1), o);
    }
    tmem_st_wait();
}

template <bool TS, bool CRITIC, bool SENSOR>
__global__ void __launch_bounds__(kPM * PolicyCfg<TS, CRITIC>::kG, 1)
policy_rollout_kernel(const __grid_constant__ DevParams<float> p, const __grid_constant__ SimView<float> v,
                      const __grid_constant__ ActorView act, const __grid_constant__ PolicyIO io) {
    static_assert(TS || !CRITIC, "the critic head exists on the TS path only");
    using PolicySmem = PolicyCfg<TS, CRITIC>;
    constexpr int kPG = PolicySmem::kG;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid_all = threadIdx.x, grp = tid_all / kPM, tid = tid_all % kPM, warp = tid >> 5;   // group-local thread / warp
    unsigned char* sX = smem + grp * PolicySmem::kXH + PolicySmem::kX;
    unsigned char* sH = smem + grp * PolicySmem::kXH + PolicySmem::kHd;
    unsigned char* sW1 = smem + PolicySmem::kW1;
    unsigned char* sW2 = smem + PolicySmem::kW2;
    unsigned char* sOnes = smem + PolicySmem::kOnes;
    unsigned char* sW2x = smem + PolicySmem::kW2x;
    unsigned char* sW1c = smem + PolicySmem::kW1c;
    unsigned char* sW2c = smem + PolicySmem::kW2c;
    unsigned char* sW2xc = smem + PolicySmem::kW2xc;
    __shared__ uint64_t bars[kPG];
    __shared__ uint32_t tmem_slot;
    uint64_t& bar = bars[grp];

    // ---- one-time: weights fp32 (global) -> bf16 canonical K-major operand tiles (shared, one copy per CTA)
    for (int idx = tid_all; idx < kPH * kPKin; idx += kPM * kPG) {   // W1: input j = 15*slot + e  ->  K index 16*slot + e
        const int n = idx / kPKin, kk = idx % kPKin, a = kk / kPSlotK, e = kk % kPSlotK;
        float w = 0.f;
        if (e < 15) w = act.w1[n * 75 + a * 15 + e];
        else if (a == 0) w = bf16_hi(act.b1[n]);                     // pad column (A holds 1 there): bias, hi piece
        else if (a == 1) w = act.b1[n] - bf16_hi(act.b1[n]);         //                                bias, lo piece
        *reinterpret_cast<__nv_bfloat16*>(sW1 + umma_canon_offset(n, kk, kPKin)) = __float2bfloat16(w);
    }
    for (int idx = tid_all; idx < kPH * kPH; idx += kPM * kPG) {
        const int n = idx / kPH, kk = idx % kPH;
        *reinterpret_cast<__nv_bfloat16*>(sW2 + umma_canon_offset(n, kk, kPH)) = __float2bfloat16(act.w2[n * kPH + kk]);
    }
    for (int idx = tid_all; idx < kPH * 16; idx += kPM * kPG) {
        const int n = idx / 16, kk = idx % 16;
        const float b = act.b2[n];
        *reinterpret_cast<__nv_bfloat16*>(sOnes + umma_canon_offset(n, kk, 16)) = __float2bfloat16(kk < 2 ? 1.f : 0.f);
        *reinterpret_cast<__nv_bfloat16*>(sW2x + umma_canon_offset(n, kk, 16)) = __float2bfloat16(kk == 0 ? bf16_hi(b) : (kk == 1 ? b - bf16_hi(b) : 0.f));
    }
    if constexpr (CRITIC) {
        for (int idx = tid_all; idx < kPH * kPKin; idx += kPM * kPG) {
            const int n = idx / kPKin, kk = idx % kPKin, a = kk / kPSlotK, e = kk % kPSlotK;
            float w = 0.f;
            if (e < 15) w = act.cw1[n * 75 + a * 15 + e];
            else if (a == 0) w = bf16_hi(act.cb1[n]);
            else if (a == 1) w = act.cb1[n] - bf16_hi(act.cb1[n]);
            *reinterpret_cast<__nv_bfloat16*>(sW1c + umma_canon_offset(n, kk, kPKin)) = __float2bfloat16(w);
        }
        for (int idx = tid_all; idx < kPH * kPH; idx += kPM * kPG) {
            const int n = idx / kPH, kk = idx % kPH;
            *reinterpret_cast<__nv_bfloat16*>(sW2c + umma_canon_offset(n, kk, kPH)) = __float2bfloat16(act.cw2[n * kPH + kk]);
        }
        for (int idx = tid_all; idx < kPH * 16; idx += kPM * kPG) {
            const int n = idx / 16, kk = idx % 16;
            const float b = act.cb2[n];
            *reinterpret_cast<__nv_bfloat16*>(sW2xc + umma_canon_offset(n, kk, 16)) = __float2bfloat16(kk == 0 ? bf16_hi(b) : (kk == 1 ? b - bf16_hi(b) : 0.f));
        }
    }
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    constexpr uint32_t kTmemCols = kPG * kPH <= 128 ? 128 : (kPG * kPH <= 256 ? 256 : 512);      // power of two >= 128 accumulator columns per group
    if (tid_all < 32) tmem_alloc(&tmem_slot, kTmemCols);
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot + (uint32_t)(grp * kPH);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint32_t phase = 0;
    LocalStats ls;
    ls.clear();
    bool any_end = false;
    const float sigma = act.action_std;
    const float log_norm = (sigma > 0.f) ? (-__logf(sigma) - 0.918938533f) : 0.f;

    const int64_t n_tiles = (v.N + kPM - 1) / kPM;
    for (int64_t tile = (int64_t)blockIdx.x * kPG + grp; tile < n_tiles; tile += (int64_t)gridDim.x * kPG) {
        const int64_t n = tile * kPM + tid;
        const bool active = n < v.N;
        Env<float> e;
        if (active) {
            load_env(v, n, e);
        } else {
#pragma unroll
            for (int k = 0; k < 13; ++k) e.y[k] = 0.f;
            e.y[6] = 1.f;
            e.prev_ang[0] = e.prev_ang[1] = e.prev_ang[2] = 0.f;
            e.prev_shaping = e.abs_sum = e.ep_return = 0.f;
            e.i = 0; e.flags = 0; e.episode = 0;
        }
        // history -> A operand: physical slot s holds age s at tile start (head = 0 is the OLDEST entry)
#pragma unroll
        for (int s = 0; s < kPSlots; ++s) {
            float h[16];
#pragma unroll
            for (int q = 0; q < 15; ++q) h[q] = (active && io.hist) ? io.hist[(int64_t)(s * 15 + q) * v.N + n] : 0.f;
            h[15] = 1.f;                                             // pad column = 1: carries b1 through the GEMM
            uint4 lo, hi;
            lo.x = pack_bf16x2(h[0], h[1]); lo.y = pack_bf16x2(h[2], h[3]); lo.z = pack_bf16x2(h[4], h[5]); lo.w = pack_bf16x2(h[6], h[7]);
            hi.x = pack_bf16x2(h[8], h[9]); hi.y = pack_bf16x2(h[10], h[11]); hi.z = pack_bf16x2(h[12], h[13]); hi.w = pack_bf16x2(h[14], h[15]);
            *reinterpret_cast<uint4*>(sX + umma_canon_offset(tid, s * kPSlotK, kPKin)) = lo;
            *reinterpret_cast<uint4*>(sX + umma_canon_offset(tid, s * kPSlotK + 8, kPKin)) = hi;
        }
        int head = 0;
        StepOut<float> o;
        float reward = 0.f;
        bool done = false, solved = false, warm_last = false;
        // Critic head (model.py:36-43, Linear-Tanh-Linear-Tanh-Linear(128,1)) on the SAME history tile as the actor: the group's
        // 128 TMEM columns are re-used once the actor's output is out (the two networks of a step form one serial chain per tile;
        // the other tiles of the CTA fill its gaps), the hidden layers are the same two UMMA rounds with the critic's operand
        // tiles, the 128 -> 1 output layer rides in the second tanh epilogue.  Called by every thread of the group.
        auto critic_value = [&]() -> float {
            tc_fence_before();
            fence_proxy_async_smem();
            group_sync(grp);                                             // the actor's accumulators are consumed; sX is visible
            if (tid == 0) {
                tc_fence_after();
#pragma unroll
                for (int a = 0; a < kPSlots; ++a) {
                    int s_ = head + a; s_ = s_ >= kPSlots ? s_ - kPSlots : s_;
                    umma_gemm_k(tmem_base, smem_u32(sX), kPKin, s_ * kPSlotK, smem_u32(sW1c), kPKin, a * kPSlotK, kPSlotK, kPH, a > 0);
                }
                umma_commit(&bar);
            }
            mbar_wait(&bar, phase); phase ^= 1;
            tc_fence_after();
            actor_hidden_epilogue_tmem(lane_addr);
            tc_fence_before();
            group_sync(grp);
            float val = c_critic_b3[0];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if (tid == 0) {
                    tc_fence_after();
                    constexpr uint32_t idesc = umma_idesc_bf16_f32(128, 64);
                    const uint32_t b_rows = smem_u32(sW2c) + (uint32_t)half * umma_canon_offset(64, 0, kPH);
#pragma unroll
                    for (int k = 0; k < kPH; k += 16)
                        umma_bf16_ts(tmem_base + 64u, tmem_base + (uint32_t)(k >> 1),
                                     umma_smem_desc(b_rows + (uint32_t)(k >> 3) * 128u, 128u, (uint32_t)(kPH >> 3) * 128u), idesc, k > 0);
                    umma_bf16(tmem_base + 64u, umma_smem_desc(smem_u32(sOnes), 128u, 256u),
                              umma_smem_desc(smem_u32(sW2xc) + (uint32_t)half * umma_canon_offset(64, 0, 16), 128u, 256u), idesc, true);
                    umma_commit(&bar);
                }
                mbar_wait(&bar, phase); phase ^= 1;
                tc_fence_after();
                val = half == 0 ? critic_output_accumulate<0, 64>(lane_addr + 64u, val) : critic_output_accumulate<64, 64>(lane_addr + 64u, val);
                tc_fence_before();
                if (half == 0) group_sync(grp);
            }
            return val;
        };
        for (int t = 0; t < io.horizon; ++t) {
            // ---------------- layer 1: [128 x 80] x W1^T, one K=16 MMA per history slot, oldest first
            fence_proxy_async_smem();
            group_sync(grp);
            if (tid == 0) {
                tc_fence_after();
#pragma unroll
                for (int a = 0; a < kPSlots; ++a) {
                    int s = head + a; s = s >= kPSlots ? s - kPSlots : s;
                    umma_gemm_k(tmem_base, smem_u32(sX), kPKin, s * kPSlotK, smem_u32(sW1), kPKin, a * kPSlotK, kPSlotK, kPH, a > 0);
                }
                umma_commit(&bar);
            }
            mbar_wait(&bar, phase); phase ^= 1;
            tc_fence_after();
            float mean[4] = {c_actor_b3[0], c_actor_b3[1], c_actor_b3[2], c_actor_b3[3]};
            if constexpr (!TS) {
                actor_hidden_epilogue(lane_addr, sH, tid);
                tc_fence_before();
                fence_proxy_async_smem();
                group_sync(grp);
                // ---------------- layer 2: [128 x 128] x W2^T
                if (tid == 0) {
                    tc_fence_after();
                    umma_gemm_k(tmem_base, smem_u32(sH), kPH, 0, smem_u32(sW2), kPH, 0, kPH, kPH, false);
                    umma_gemm_k(tmem_base, smem_u32(sOnes), 16, 0, smem_u32(sW2x), 16, 0, 16, kPH, true);      // + b2
                    umma_commit(&bar);
                }
                mbar_wait(&bar, phase); phase ^= 1;
                tc_fence_after();
                // ---------------- second tanh + layer 3 (128 -> 4) on the FP32 pipe, straight from the accumulators
                actor_output_accumulate<0, kPH>(lane_addr, mean);
                tc_fence_before();
            } else {
                actor_hidden_epilogue_tmem(lane_addr);                   // H1 -> TMEM columns [0,64) of the group
                tc_fence_before();
                group_sync(grp);
                // ---------------- layer 2 in two N = 64 halves: D[64,128) = H1 (TMEM) x W2[64 half ..][:]^T + b2
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    if (tid == 0) {
                        tc_fence_after();
                        constexpr uint32_t idesc = umma_idesc_bf16_f32(128, 64);
                        const uint32_t b_rows = smem_u32(sW2) + (uint32_t)half * umma_canon_offset(64, 0, kPH);
#pragma unroll
                        for (int k = 0; k < kPH; k += 16)
                            umma_bf16_ts(tmem_base + 64u, tmem_base + (uint32_t)(k >> 1),
                                         umma_smem_desc(b_rows + (uint32_t)(k >> 3) * 128u, 128u, (uint32_t)(kPH >> 3) * 128u), idesc, k > 0);
                        umma_bf16(tmem_base + 64u, umma_smem_desc(smem_u32(sOnes), 128u, 256u),
                                  umma_smem_desc(smem_u32(sW2x) + (uint32_t)half * umma_canon_offset(64, 0, 16), 128u, 256u), idesc, true);   // + b2
                        umma_commit(&bar);
                    }
                    mbar_wait(&bar, phase); phase ^= 1;
                    tc_fence_after();
                    if (half == 0) actor_output_accumulate<0, 64>(lane_addr + 64u, mean);
                    else actor_output_accumulate<64, 64>(lane_addr + 64u, mean);
                    tc_fence_before();
                    if (half == 0) group_sync(grp);                      // every thread has read the half before it is overwritten
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) mean[k] = tanh_fast(mean[k]);
            if constexpr (CRITIC) {                                      // V of the input the actor has just seen (memory.values, model.py:68)
                const float val = critic_value();
                if (active && io.value_out) io.value_out[(int64_t)t * v.N + n] = val;
            }
            // ---------------- a ~ Normal(mean, sigma)  (model.py:60-66), per-dimension log-prob
            float a[4], logp[4];
            if (sigma > 0.f) {
                const uint4 u = philox_block(v.seed, v.env_id_offset + (uint32_t)n, e.episode, (uint32_t)e.i, RNG_POLICY);
                float z[4];
                box_muller(u32_to_unit<float>(u.x), u32_to_unit<float>(u.y), &z[0], &z[1]);
                box_muller(u32_to_unit<float>(u.z), u32_to_unit<float>(u.w), &z[2], &z[3]);
#pragma unroll
                for (int k = 0; k < 4; ++k) { a[k] = fmaf(sigma, z[k], mean[k]); logp[k] = fmaf(-0.5f * z[k], z[k], log_norm); }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) { a[k] = mean[k]; logp[k] = 0.f; }
            }
            // ---------------- quad.step (same code path as rollout_kernel)
            bool warm = false;
            if (p.flags & F_ASYNC_RESET) warm = async_warmup_prologue(p, e, a);
            const bool was_done = (e.flags & EF_DONE) != 0;
            Ctrl<float> c;
            step_core<float, 0, true>(p, e, a, o, SENSOR ? &c : nullptr);
            if (warm) o.reward = 0.f; else e.ep_return += o.reward;
            reward = o.reward; done = o.done; solved = o.solved; warm_last = warm;
            if (active && done && !was_done) { count_episode(ls, p, e, o); any_end = true; }
            // QS_FLAG_SENSOR_NOISE (quadrotor_env.py:579-724): the sensor model after the step, exactly as in rollout_kernel — the sensor
            // rows stay in HBM, one out-of-line copy of the model per kernel; warm-up steps bypass it, the last one re-initialises it
            if (SENSOR && active)
                sensor_update(p, v, n, e, c, o.vq[0], o.vq[1], o.vq[2], o.vq[3], warm ? ((e.flags >> EF_WARM_SHIFT) ? 2 : 1) : 0);
            if ((p.flags & F_ASYNC_RESET) && done) {
                async_resample(p, v.seed, v.env_id_offset + (uint32_t)n, e, o.vq);
                if (SENSOR && active) {
#pragma unroll
                    for (int k = 0; k < 10; ++k) v.sensed_obs[k * v.ld + n] = e.y[k];
#pragma unroll
                    for (int k = 0; k < 4; ++k) v.sensed_obs[(10 + k) * v.ld + n] = o.vq[k];
                }
            }
            // what the policy observes: the sensed observation on sensor handles (the loop of visual_landing/rl_worker.py:164-175 feeds
            // sensor_sp's states_sens to the network), the true one otherwise
            float so[14];
            if (SENSOR) {
#pragma unroll
                for (int k = 0; k < 14; ++k) so[k] = active ? v.sensed_obs[k * v.ld + n] : 0.f;
            } else {
#pragma unroll
                for (int k = 0; k < 10; ++k) so[k] = e.y[k];
#pragma unroll
                for (int k = 0; k < 4; ++k) so[10 + k] = o.vq[k];
            }
            if (active) {
                if (SENSOR && io.sensed_out) {
                    float* st_ = io.sensed_out + (int64_t)t * 14 * v.N;
#pragma unroll
                    for (int k = 0; k < 14; ++k) st_[k * v.N + n] = so[k];
                }
                if (io.obs_out) {
                    float* ot = io.obs_out + (int64_t)t * 14 * v.N;
#pragma unroll
                    for (int k = 0; k < 10; ++k) ot[k * v.N + n] = e.y[k];
#pragma unroll
                    for (int k = 0; k < 4; ++k) ot[(10 + k) * v.N + n] = o.vq[k];
                }
                if (io.action_out) {
                    float* at = io.action_out + (int64_t)t * 4 * v.N;
#pragma unroll
                    for (int k = 0; k < 4; ++k) at[k * v.N + n] = a[k];
                }
                if (io.logprob_out) {
                    float* lt = io.logprob_out + (int64_t)t * 4 * v.N;
#pragma unroll
                    for (int k = 0; k < 4; ++k) lt[k * v.N + n] = logp[k];
                }
                if (io.reward_out) io.reward_out[(int64_t)t * v.N + n] = reward;
                if (io.done_out) io.done_out[(int64_t)t * v.N + n] = (uint8_t)((done ? 1 : 0) | (warm ? 2 : 0));
            }
            // ---------------- dl_in_gen.dl_input (dl_auxiliary.py:25-32): the new entry replaces the oldest slot
            {
                uint4 lo, hi;
                lo.x = pack_bf16x2(a[0], a[1]); lo.y = pack_bf16x2(a[2], a[3]);
                lo.z = pack_bf16x2(so[1], so[3]); lo.w = pack_bf16x2(so[5], so[6]);
                hi.x = pack_bf16x2(so[7], so[8]); hi.y = pack_bf16x2(so[9], so[10]);
                hi.z = pack_bf16x2(so[11], so[12]); hi.w = pack_bf16x2(so[13], 1.f);
                *reinterpret_cast<uint4*>(sX + umma_canon_offset(tid, head * kPSlotK, kPKin)) = lo;
                *reinterpret_cast<uint4*>(sX + umma_canon_offset(tid, head * kPSlotK + 8, kPKin)) = hi;
                head = head + 1 == kPSlots ? 0 : head + 1;
            }
        }
        if constexpr (CRITIC) {                                          // bootstrap value of the input after the last step (GAE, ppo.py:384)
            const float val = critic_value();
            if (active && io.value_out) io.value_out[(int64_t)io.horizon * v.N + n] = val;
        }
        if (active) {
            store_env(v, n, e, o.vq);
            v.reward[n] = reward;
            v.done[n] = (uint8_t)((done ? 1 : 0) | (warm_last ? 2 : 0));
            v.solved[n] = solved;
            if (io.hist) {                                           // back to oldest-first order
#pragma unroll
                for (int a = 0; a < kPSlots; ++a) {
                    int s = head + a; s = s >= kPSlots ? s - kPSlots : s;
                    const uint4 lo = *reinterpret_cast<const uint4*>(sX + umma_canon_offset(tid, s * kPSlotK, kPKin));
                    const uint4 hi = *reinterpret_cast<const uint4*>(sX + umma_canon_offset(tid, s * kPSlotK + 8, kPKin));
                    const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
                    for (int q = 0; q < 15; ++q) {
                        const uint32_t bits = (q & 1) ? (w[q >> 1] & 0xFFFF0000u) : (w[q >> 1] << 16);
                        io.hist[(int64_t)(a * 15 + q) * v.N + n] = __uint_as_float(bits);
                    }
                }
            }
        }
        group_sync(grp);                                             // sX is re-initialised by the next tile
    }
    flush_stats(ls, any_end, v.stats);
    if (blockIdx.x == 0 && tid_all == 0) atomicAdd(&v.stats[7], (double)v.N * io.horizon);
    tc_fence_before();
    __syncthreads();
    if (tid_all < 32) tmem_dealloc(tmem_slot, kTmemCols);
}

extern "C" int qs_policy_rollout(qs_handle h, const qs_actor* actor, const qs_policy_rollout_args* args, void* stream) {
    if (!h || !actor || !args) return fail(QS_EINVAL, "qs_policy_rollout: NULL argument");
    if (args->horizon < 1) return fail(QS_EINVAL, "qs_policy_rollout: horizon must be >= 1");
    if (actor->hidden != kPH || actor->in_dim != 75)
        return fail(QS_EINVAL, "qs_policy_rollout: only the 75-128-128-4 actor (history T=5) is supported");
    if (!actor->w1 || !actor->b1 || !actor->w2 || !actor->b2 || !actor->w3 || !actor->b3)
        return fail(QS_EINVAL, "qs_policy_rollout: NULL weight pointer");
    const uint32_t f = h->cfg.flags;
    if (h->cfg.precision != QS_F32 || h->cfg.integrator != QS_RK4 || !(f & QS_FLAG_DIRECT_CONTROL))
        return fail(QS_ESTATE, "qs_policy_rollout: needs an FP32 / RK4 / direct-control handle");
    if (f & (QS_FLAG_AUX | QS_FLAG_AUTO_RESET | QS_FLAG_ROBUST))
        return fail(QS_ESTATE, "qs_policy_rollout: not available with AUX / ROBUST / strict AUTO_RESET (use ASYNC_RESET)");
    const bool sensor = (f & QS_FLAG_SENSOR_NOISE) != 0;
    if (args->sensed_obs_out && !sensor) return fail(QS_ESTATE, "qs_policy_rollout: sensed_obs_out needs a QS_FLAG_SENSOR_NOISE handle");
    QS_USE_DEVICE(h);
    const bool critic = actor->cw1 != nullptr;
    if (critic && (!actor->cb1 || !actor->cw2 || !actor->cb2 || !actor->cw3 || !actor->cb3))
        return fail(QS_EINVAL, "qs_policy_rollout: the critic needs all six of cw1, cb1, cw2, cb2, cw3, cb3");
    if (args->value_out && !critic) return fail(QS_EINVAL, "qs_policy_rollout: value_out needs the critic weights (qs_actor.cw1 ...)");
    ActorView av{actor->w1, actor->b1, actor->w2, actor->b2, actor->w3, actor->b3, actor->action_std,
                 actor->cw1, actor->cb1, actor->cw2, actor->cb2};
    PolicyIO io{args->horizon, (float*)args->obs_out, (float*)args->action_out, (float*)args->logprob_out,
                (float*)args->reward_out, args->done_out, (float*)args->hist, (float*)args->value_out, (float*)args->sensed_obs_out};
    constexpr bool kTS = QS_POLICY_TS != 0;
    if (critic && !kTS) return fail(QS_ESTATE, "qs_policy_rollout: the critic head needs the TS build (QS_POLICY_TS=1)");
    {   // output layers -> constant memory, stream-ordered (the action stage of the handle is free during a policy rollout)
        float* tmp = (float*)h->action_stage;
        k_pack_w3<<<1, kPH, 0, (cudaStream_t)stream>>>(actor->w3, actor->b3, tmp);
        QS_CUDA(cudaMemcpyToSymbolAsync(c_actor_w3, tmp, sizeof(float) * kPH * 4, 0, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        QS_CUDA(cudaMemcpyToSymbolAsync(c_actor_b3, tmp + kPH * 4, sizeof(float) * 4, 0, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        if (critic) {
            QS_CUDA(cudaMemcpyToSymbolAsync(c_critic_w3, actor->cw3, sizeof(float) * kPH, 0, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
            QS_CUDA(cudaMemcpyToSymbolAsync(c_critic_b3, actor->cb3, sizeof(float), 0, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        }
    }
    const int64_t tiles = (h->N + kPM - 1) / kPM;
    int64_t grid = (int64_t)h->sm_count;                             // one CTA per SM, kG tiles in flight each
    constexpr int kG = PolicyCfg<kTS>::kG;
    const int64_t need = (tiles + kG - 1) / kG;
    if (grid > need) grid = need;
#define QS_LAUNCH_POLICY(CR, SE)                                                                                                  \
    do {                                                                                                                          \
        using Cfg = PolicyCfg<kTS, CR>;                                                                                           \
        QS_SET_SMEM_ONCE(h, (policy_rollout_kernel<kTS, CR, SE>), Cfg::kBytes);                                                    \
        policy_rollout_kernel<kTS, CR, SE><<<(int)grid, kPM * kG, Cfg::kBytes, (cudaStream_t)stream>>>(h->pf, make_view<float>(h), av, io); \
    } while (0)
    if constexpr (kTS) {
        if (critic) {
            if (sensor) QS_LAUNCH_POLICY(true, true); else QS_LAUNCH_POLICY(true, false);
            QS_CUDA(cudaGetLastError());
            return QS_OK;
        }
    }
    if (sensor) QS_LAUNCH_POLICY(false, true); else QS_LAUNCH_POLICY(false, false);
#undef QS_LAUNCH_POLICY
    QS_CUDA(cudaGetLastError());
    return QS_OK;
}
