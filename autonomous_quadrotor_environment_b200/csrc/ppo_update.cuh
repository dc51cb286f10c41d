// ppo_update.cuh — the network update of the reference's PPO trainer (environment/controller/ppo.py:143-209, model.py:19-88) as
// hand-written sm_100a kernels on the time-major rollout buffers ([K][C][N], never leaving HBM):
//
//   ppo_grad_kernel<NET>   one full-batch gradient of the clipped-surrogate loss for ONE network (NET = 0 actor 75-128-128-4 with a
//                          tanh output, NET = 1 critic 75-128-128-1): forward AND backward of the MLP on tcgen05 tensor cores (BF16
//                          operands, FP32 TMEM accumulators), FP32 gradients accumulated over the launch
//   adam_kernel            torch.optim.Adam's step (ppo.py:105, default eps 1e-8, no weight decay) on FP32 master weights
//
// One CTA of 128 threads per SM walks env tiles of 128 envs through all K recorded steps; thread = sample = TMEM lane.
// Per (tile, step), with X the [128 x 80] history tile of the rollout kernel (15 floats per entry padded to 16, pad = 1: bias column):
//   Z1 = X W1^T           SS-form MMA, 5 x K16 (one per history slot of the ring)                          -> TMEM [64,192)
//   H1 = tanh(Z1)         packed BF16 -> TMEM [0,64) (A operand of layer 2, TS form) and, TRANSPOSED, -> smem H1^T [j][s]
//   Z2 = H1 W2^T + b2     TS-form MMA, 8 x K16 (+ one K16 block [1 1 0..] x [b2_hi b2_lo 0..])                  -> TMEM [64,192)
//   H2 = tanh(Z2); out = W3 H2 + b3 (FP32 pipe, weights in constant memory); loss gradient dZ3 per sample (registers)
//   dW3 += H2^T dZ3       H2^T [j][s] written transposed to smem, SS-form MMA M=128 (j) N=16 K=128 (s)          -> TMEM [480,496)
//   dZ2 = (W3^T dZ3) (1 - H2^2)   packed BF16 -> TMEM [192,256) (A operand, TS form) and transposed -> smem dZ^T [j][s]
//   dW2 += dZ2^T H1       SS-form MMA M=128 (j2) N=128 (j1) K=128 (s); db2 += dZ2^T 1 (N=16 block against a ones tile)  -> TMEM [256,384), [464,480)
//   dH1 = dZ2 W2          TS-form MMA against W2^T staged K(=j2)-major                                          -> TMEM [64,192)
//   dZ1 = dH1 (1 - H1^2)  (H1 re-read from TMEM [0,64)) transposed -> smem dZ^T
//   dW1 += dZ1^T X        5 x (M=128, N=16, K=128) against X^T [i][s] kept as a second, transposed ring; the pad columns of
//                         slots 0 / 1 carry b1 (hi / lo pieces in the forward): db1 = column 15 of dW1                  -> TMEM [384,464)
// The weight-gradient accumulators stay in TENSOR MEMORY for the whole launch (496 of the 512 columns) and are added to the FP32
// gradient buffers in HBM once per CTA.  Every operand with K = sample index needs the transpose of what a thread (= sample) holds:
// those tiles are written with 2-byte scattered shared-memory stores in the canonical K-major layout (umma.cuh).
#pragma once
#include "umma.cuh"

namespace ppo {

constexpr int kM = 128;          // samples per tile = UMMA M / K
constexpr int kH = 128;          // hidden width
constexpr int kSlots = 5, kSlotK = 16, kKin = kSlots * kSlotK;   // 80

struct Batch {
    int64_t N;                   // envs
    int32_t K;                   // recorded steps
    int32_t flags;               // QS_PPO_RECORD_LOGP: write logp_old from this forward pass (ratio = 1) instead of reading it
    const float* hist0;          // [75][N]   dl_in_gen buffer at rollout start, oldest entry first
    const float* entries;        // [K][15][N] history entry pushed after step t: [action(4), v(3), q(4), dq(4)] (BF16-rounded)
    const float* actions;        // [K][4][N]
    float* logp_old;             // [K][4][N]  (actor; written when flags & QS_PPO_RECORD_LOGP)
    const float* adv;            // [K][N]     normalised advantages (actor)
    const float* ret;            // [K][N]     returns (critic)
    const float* weight;         // [K][N]     1 = transition, 0 = warm-up step of an asynchronous reset
};
struct Net {                     // FP32, PyTorch Linear layout [out][in]
    const float *w1, *b1, *w2, *b2, *w3, *b3;
};
struct Grad {                    // FP32, same shapes, ACCUMULATED (atomicAdd); loss: double accumulator of the summed per-sample loss
    float *w1, *b1, *w2, *b2, *w3, *b3;
    double* loss;
};

// shared-memory map (bytes)
constexpr int oX = 0;                                 // [128 s][80]   K-major in the input index (A of layer 1)
constexpr int oXT = oX + kM * kKin * 2;               // [80 i][128 s] K-major in the sample index (B of dW1)
constexpr int oH1T = oXT + kKin * kM * 2;             // [128 j][128 s]  H1^T, then H2^T (B of dW2 / A of dW3)
constexpr int oDZT = oH1T + kH * kM * 2;              // [128 j][128 s]  dZ2^T, then dZ1^T (A of dW2 / dW1)
constexpr int oW1 = oDZT + kH * kM * 2;               // [128 j][80]
constexpr int oW2 = oW1 + kH * kKin * 2;              // [128 j2][128 j1]  (B of layer 2)
constexpr int oW2T = oW2 + kH * kH * 2;               // [128 j1][128 j2]  (B of dH1 = dZ2 W2)
constexpr int oOnes = oW2T + kH * kH * 2;             // [128 s][16]  columns 0,1 = 1 (A of the b2 block, K-major in the 16)
constexpr int oW2x = oOnes + kM * 16 * 2;             // [128 j][16]  columns 0,1 = b2 (hi, lo)
constexpr int oOnesT = oW2x + kH * 16 * 2;            // [16 n][128 s] row 0 = 1 (B of db2), K-major in s
constexpr int oDZ3T = oOnesT + 16 * kM * 2;           // [16 k][128 s] dZ3^T (B of dW3)
constexpr int kSmemBytes = oDZ3T + 16 * kM * 2;

// TMEM column map
constexpr uint32_t cH1 = 0, cZ = 64, cDZ2 = 192, cDW2 = 256, cDW1 = 384, cDB2 = 464, cDW3 = 480, kCols = 512;

__constant__ float c_w3[2][kH * 4];     // [net][j*4 + k] = w3[k][j]  (critic: k = 0 only)
__constant__ float c_b3[2][4];

using namespace qs;

__device__ __forceinline__ void st_bf16(unsigned char* base, uint32_t off, float x) {
    *reinterpret_cast<__nv_bfloat16*>(base + off) = __float2bfloat16(x);
}

template <int NET>
__global__ void __launch_bounds__(kM, 1)
ppo_grad_kernel(const __grid_constant__ Batch b, const __grid_constant__ Net w, const __grid_constant__ Grad g, float sigma,
                float eps_clip, float inv_count) {
    constexpr int OUT = NET == 0 ? 4 : 1;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __shared__ float s_loss;
    const int tid = threadIdx.x, warp = tid >> 5;

    // ---- one-time: FP32 master weights -> BF16 operand tiles
    for (int idx = tid; idx < kH * kKin; idx += kM) {            // W1: input 15*slot + e -> K index 16*slot + e; pad columns carry b1
        const int n = idx / kKin, kk = idx % kKin, a = kk / kSlotK, e = kk % kSlotK;
        float x = 0.f;
        if (e < 15) x = w.w1[n * 75 + a * 15 + e];
        else if (a == 0) x = __bfloat162float(__float2bfloat16(w.b1[n]));
        else if (a == 1) x = w.b1[n] - __bfloat162float(__float2bfloat16(w.b1[n]));
        st_bf16(smem + oW1, umma_canon_offset(n, kk, kKin), x);
    }
    for (int idx = tid; idx < kH * kH; idx += kM) {
        const int n = idx / kH, kk = idx % kH;
        const float x = w.w2[n * kH + kk];
        st_bf16(smem + oW2, umma_canon_offset(n, kk, kH), x);    // [j2][j1]
        st_bf16(smem + oW2T, umma_canon_offset(kk, n, kH), x);   // [j1][j2]
    }
    for (int idx = tid; idx < kH * 16; idx += kM) {
        const int n = idx / 16, kk = idx % 16;
        const float bb = w.b2[n], hi = __bfloat162float(__float2bfloat16(bb));
        st_bf16(smem + oOnes, umma_canon_offset(n, kk, 16), kk < 2 ? 1.f : 0.f);
        st_bf16(smem + oW2x, umma_canon_offset(n, kk, 16), kk == 0 ? hi : (kk == 1 ? bb - hi : 0.f));
    }
    for (int idx = tid; idx < 16 * kM; idx += kM) {
        const int n = idx / kM, kk = idx % kM;
        st_bf16(smem + oOnesT, umma_canon_offset(n, kk, kM), n == 0 ? 1.f : 0.f);
        st_bf16(smem + oDZ3T, umma_canon_offset(n, kk, kM), 0.f);
    }
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); s_loss = 0.f; }
    if (tid < 32) tmem_alloc(&tmem_slot, kCols);
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_slot;
    const uint32_t lane = tb + ((uint32_t)(warp * 32) << 16);
    uint32_t phase = 0;
    bool have_acc = false;                                       // the weight-gradient accumulators hold something
    float loss_local = 0.f;
    float db3_local[4] = {0.f, 0.f, 0.f, 0.f};                   // db3[k] = sum_s dZ3[s][k]
    const float inv_var = 1.f / (sigma * sigma);
    const float log_norm = -__logf(sigma) - 0.918938533f;

    auto commit_wait = [&]() {                                   // all threads: wait for the MMAs thread 0 has issued
        if (tid == 0) umma_commit(&bar);
        mbar_wait(&bar, phase); phase ^= 1;
        tc_fence_after();
    };
    auto sync_before_issue = [&]() {                             // operands written by all threads -> visible to the tensor core
        tc_fence_before();
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) tc_fence_after();
    };

    const int64_t n_tiles = (b.N + kM - 1) / kM;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t n = tile * kM + tid;
        const bool active = n < b.N;
        // history ring: physical slot s_ holds age s_ at t = 0 (hist0, oldest first)
        __syncthreads();                                         // previous tile's MMAs on sX / sXT are complete (waited), all threads done
#pragma unroll
        for (int s_ = 0; s_ < kSlots; ++s_) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float x = q < 15 ? (active ? b.hist0[(int64_t)(s_ * 15 + q) * b.N + n] : 0.f) : 1.f;
                st_bf16(smem + oX, umma_canon_offset(tid, s_ * kSlotK + q, kKin), x);
                st_bf16(smem + oXT, umma_canon_offset(s_ * kSlotK + q, tid, kM), x);
            }
        }
        int head = 0;
        for (int t = 0; t < b.K; ++t) {
            const int64_t tn = (int64_t)t * b.N + n;
            const float wgt = active ? b.weight[tn] * inv_count : 0.f;
            // ================= forward
            sync_before_issue();
            if (tid == 0) {
#pragma unroll
                for (int a = 0; a < kSlots; ++a) {
                    int s_ = head + a; s_ = s_ >= kSlots ? s_ - kSlots : s_;
                    umma_gemm_k(tb + cZ, smem_u32(smem + oX), kKin, s_ * kSlotK, smem_u32(smem + oW1), kKin, a * kSlotK, kSlotK, kH, a > 0);
                }
            }
            commit_wait();
            // H1 = tanh(Z1): packed BF16 -> TMEM [0,64) and transposed -> sH1T
#pragma unroll 1
            for (int c = 0; c < kH; c += 32) {
                float acc[32];
                tmem_ld_32x32b_x32(lane + cZ + (uint32_t)c, acc);
                uint32_t o[16];
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[i] = tanh_fast(acc[i]);
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = pack_bf16x2(acc[2 * i], acc[2 * i + 1]);
                tmem_st_32x32b_x16(lane + cH1 + (uint32_t)(c >> 1), o);
#pragma unroll
                for (int i = 0; i < 32; ++i) st_bf16(smem + oH1T, umma_canon_offset(c + i, tid, kM), acc[i]);
            }
            tmem_st_wait();
            sync_before_issue();
            if (tid == 0) {
                constexpr uint32_t idesc = umma_idesc_bf16_f32(128, kH);
#pragma unroll
                for (int k = 0; k < kH; k += 16)
                    umma_bf16_ts(tb + cZ, tb + cH1 + (uint32_t)(k >> 1),
                                 umma_smem_desc(smem_u32(smem + oW2) + (uint32_t)(k >> 3) * 128u, 128u, (uint32_t)(kH >> 3) * 128u), idesc, k > 0);
                umma_bf16(tb + cZ, umma_smem_desc(smem_u32(smem + oOnes), 128u, 256u), umma_smem_desc(smem_u32(smem + oW2x), 128u, 256u), idesc, true);
            }
            commit_wait();
            // output layer on the FP32 pipe
            float out[4] = {c_b3[NET][0], c_b3[NET][1], c_b3[NET][2], c_b3[NET][3]};
#pragma unroll 1
            for (int c = 0; c < kH; c += 32) {
                float acc[32];
                tmem_ld_32x32b_x32(lane + cZ + (uint32_t)c, acc);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float h2 = tanh_fast(acc[i]);
#pragma unroll
                    for (int k = 0; k < OUT; ++k) out[k] = fmaf(h2, c_w3[NET][(c + i) * 4 + k], out[k]);
                }
            }
            // ================= loss and its gradient w.r.t. the pre-activation of the output layer (per sample, x weight / count)
            float dz3[4] = {0.f, 0.f, 0.f, 0.f};
            if (NET == 0) {
                float mean[4], a[4], lp_new = 0.f, lp_old = 0.f;
                const bool record = (b.flags & QS_PPO_RECORD_LOGP) != 0;                      // memory.logprobs of policy_old == policy (ppo.py:206)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    mean[k] = tanh_fast(out[k]);                                              // model.py:33-34
                    a[k] = active ? b.actions[((int64_t)t * 4 + k) * b.N + n] : 0.f;
                    const float d = a[k] - mean[k];
                    const float lp = -0.5f * d * d * inv_var + log_norm;                      // Normal.log_prob, model.py:82
                    lp_new += lp;
                    if (record) { if (active) b.logp_old[((int64_t)t * 4 + k) * b.N + n] = lp; lp_old += lp; }
                    else lp_old += active ? b.logp_old[((int64_t)t * 4 + k) * b.N + n] : 0.f;
                }
                const float adv = active ? b.adv[tn] : 0.f;
                const float ratio = __expf(lp_new - lp_old);                                  // ppo.py:187
                const float surr1 = ratio * adv;                                              // :192
                const float rc = fminf(fmaxf(ratio, 1.f - eps_clip), 1.f + eps_clip);
                const float surr2 = rc * adv;                                                 // :193
                // torch.min(surr1, surr2): the gradient flows through the selected branch; clamp passes it inside [1-eps, 1+eps] only
                const bool through = (surr1 <= surr2) || (ratio >= 1.f - eps_clip && ratio <= 1.f + eps_clip);
                const float dl_dlp = through ? -adv * ratio : 0.f;                            // d(-min)/d(sum of log-probs)
                loss_local += wgt * (-fminf(surr1, surr2));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    dz3[k] = wgt * dl_dlp * (a[k] - mean[k]) * inv_var * (1.f - mean[k] * mean[k]);
            } else {
                const float ret = active ? b.ret[tn] : 0.f;
                const float d = out[0] - ret;                                                 // 0.5 * MSE, ppo.py:194
                loss_local += wgt * 0.5f * d * d;
                dz3[0] = wgt * d;
            }
#pragma unroll
            for (int k = 0; k < OUT; ++k) db3_local[k] += dz3[k];
            // ================= backward
            // H2^T -> smem (over H1^T: the layer-2 MMAs that read TMEM H1 are complete; H1^T itself is needed again only by dW2, which
            // is issued AFTER H1^T has been rewritten below — so H2^T goes to the dZ^T buffer, not here) and dZ3^T
#pragma unroll 1
            for (int c = 0; c < kH; c += 32) {
                float acc[32];
                tmem_ld_32x32b_x32(lane + cZ + (uint32_t)c, acc);
#pragma unroll
                for (int i = 0; i < 32; ++i) st_bf16(smem + oDZT, umma_canon_offset(c + i, tid, kM), tanh_fast(acc[i]));
            }
#pragma unroll
            for (int k = 0; k < OUT; ++k) st_bf16(smem + oDZ3T, umma_canon_offset(k, tid, kM), dz3[k]);
            sync_before_issue();
            if (tid == 0) {                                       // dW3[j][k] += sum_s H2^T[j][s] dZ3^T[k][s]
                umma_gemm_k(tb + cDW3, smem_u32(smem + oDZT), kM, 0, smem_u32(smem + oDZ3T), kM, 0, kM, 16, have_acc);
            }
            commit_wait();
            // dZ2 = (W3^T dZ3) (1 - H2^2): packed -> TMEM [192,256), transposed -> sDZT (the dW3 MMA that read it has completed)
#pragma unroll 1
            for (int c = 0; c < kH; c += 32) {
                float acc[32];
                tmem_ld_32x32b_x32(lane + cZ + (uint32_t)c, acc);
                uint32_t o[16];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float h2 = tanh_fast(acc[i]);
                    float dh = 0.f;
#pragma unroll
                    for (int k = 0; k < OUT; ++k) dh = fmaf(dz3[k], c_w3[NET][(c + i) * 4 + k], dh);
                    acc[i] = dh * (1.f - h2 * h2);
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = pack_bf16x2(acc[2 * i], acc[2 * i + 1]);
                tmem_st_32x32b_x16(lane + cDZ2 + (uint32_t)(c >> 1), o);
#pragma unroll
                for (int i = 0; i < 32; ++i) st_bf16(smem + oDZT, umma_canon_offset(c + i, tid, kM), acc[i]);
            }
            tmem_st_wait();
            sync_before_issue();
            if (tid == 0) {
                // dW2[j2][j1] += sum_s dZ2^T[j2][s] H1^T[j1][s];  db2[j2] += sum_s dZ2^T[j2][s] 1
                umma_gemm_k(tb + cDW2, smem_u32(smem + oDZT), kM, 0, smem_u32(smem + oH1T), kM, 0, kM, kH, have_acc);
                umma_gemm_k(tb + cDB2, smem_u32(smem + oDZT), kM, 0, smem_u32(smem + oOnesT), kM, 0, kM, 16, have_acc);
                // dH1[s][j1] = sum_j2 dZ2[s][j2] W2[j2][j1]: A = dZ2 from TMEM, B = W2^T [j1][j2]
                constexpr uint32_t idesc = umma_idesc_bf16_f32(128, kH);
#pragma unroll
                for (int k = 0; k < kH; k += 16)
                    umma_bf16_ts(tb + cZ, tb + cDZ2 + (uint32_t)(k >> 1),
                                 umma_smem_desc(smem_u32(smem + oW2T) + (uint32_t)(k >> 3) * 128u, 128u, (uint32_t)(kH >> 3) * 128u), idesc, k > 0);
            }
            commit_wait();
            // dZ1 = dH1 (1 - H1^2), transposed -> sDZT (the dW2 / db2 MMAs that read it have completed)
#pragma unroll 1
            for (int c = 0; c < kH; c += 32) {
                float acc[32], h1p[16];
                tmem_ld_32x32b_x32(lane + cZ + (uint32_t)c, acc);
                tmem_ld_32x32b_x16(lane + cH1 + (uint32_t)(c >> 1), h1p);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint32_t pk = __float_as_uint(h1p[i]);
                    const float ha = __uint_as_float(pk << 16), hb = __uint_as_float(pk & 0xFFFF0000u);
                    st_bf16(smem + oDZT, umma_canon_offset(c + 2 * i, tid, kM), acc[2 * i] * (1.f - ha * ha));
                    st_bf16(smem + oDZT, umma_canon_offset(c + 2 * i + 1, tid, kM), acc[2 * i + 1] * (1.f - hb * hb));
                }
            }
            sync_before_issue();
            if (tid == 0) {                                       // dW1[j][16 a + e] += sum_s dZ1^T[j][s] X^T[16 slot + e][s]
#pragma unroll
                for (int a = 0; a < kSlots; ++a) {
                    int s_ = head + a; s_ = s_ >= kSlots ? s_ - kSlots : s_;
                    const uint32_t brow = smem_u32(smem + oXT) + umma_canon_offset(s_ * kSlotK, 0, kM);
                    umma_gemm_k(tb + cDW1 + (uint32_t)(a * kSlotK), smem_u32(smem + oDZT), kM, 0, brow, kM, 0, kM, 16, have_acc);
                }
            }
            commit_wait();
            have_acc = true;
            // ================= next step's input: the entry recorded after step t replaces the oldest slot (dl_auxiliary.py:25-32)
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float x = q < 15 ? (active ? b.entries[((int64_t)t * 15 + q) * b.N + n] : 0.f) : 1.f;
                st_bf16(smem + oX, umma_canon_offset(tid, head * kSlotK + q, kKin), x);
                st_bf16(smem + oXT, umma_canon_offset(head * kSlotK + q, tid, kM), x);
            }
            head = head + 1 == kSlots ? 0 : head + 1;
        }
    }
    // ---- accumulators -> HBM gradients (thread = row j of every accumulator)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (have_acc) {
        const int j = tid;
#pragma unroll 1
        for (int c = 0; c < kH; c += 32) {
            float acc[32];
            tmem_ld_32x32b_x32(lane + cDW2 + (uint32_t)c, acc);
#pragma unroll
            for (int i = 0; i < 32; ++i) atomicAdd(&g.w2[j * kH + c + i], acc[i]);
        }
        float d1[16];
#pragma unroll 1
        for (int a = 0; a < kSlots; ++a) {
            tmem_ld_32x32b_x16(lane + cDW1 + (uint32_t)(a * kSlotK), d1);
#pragma unroll
            for (int e = 0; e < 15; ++e) atomicAdd(&g.w1[j * 75 + a * 15 + e], d1[e]);
            if (a == 0) atomicAdd(&g.b1[j], d1[15]);             // the bias rides in the pad column (every slot's pad column sees the same sum)
        }
        tmem_ld_32x32b_x16(lane + cDB2, d1);
        atomicAdd(&g.b2[j], d1[0]);
        tmem_ld_32x32b_x16(lane + cDW3, d1);
#pragma unroll
        for (int k = 0; k < OUT; ++k) atomicAdd(&g.w3[k * kH + j], d1[k]);
    }
    // b3 gradient and the loss: per-thread sums -> warp shuffle -> HBM
#pragma unroll
    for (int k = 0; k < OUT; ++k) {
        float x = db3_local[k];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
        if ((tid & 31) == 0 && x != 0.f) atomicAdd(&g.b3[k], x);
    }
    atomicAdd(&s_loss, loss_local);
    __syncthreads();
    if (tid == 0 && g.loss) atomicAdd(g.loss, (double)s_loss);
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tb, kCols);
}

// [j][4] packing of an output layer for constant memory (+ the bias), OUT rows of w3
__global__ void k_pack_out(const float* __restrict__ w3, const float* __restrict__ b3, int out_dim, float* __restrict__ dst) {
    const int j = threadIdx.x;
    if (j < kH) { for (int k = 0; k < 4; ++k) dst[j * 4 + k] = k < out_dim ? w3[k * kH + j] : 0.f; }
    if (j < 4) dst[kH * 4 + j] = j < out_dim ? b3[j] : 0.f;
}

// torch.optim.Adam.step (ppo.py:105; amsgrad off, weight_decay 0, eps 1e-8):  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
// p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps).  One flat FP32 parameter vector.
__global__ void adam_kernel(int64_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            float lr, float beta1, float beta2, float eps, float bc1, float bc2_sqrt) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gi = g[i];
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi; v[i] = vi;
        p[i] -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
    }
}

}  // namespace ppo

// ------------------------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------------------------
extern "C" int qs_ppo_grad(const qs_ppo_batch* bt, const qs_ppo_net* net, const qs_ppo_net* grad, int which, float sigma, float eps_clip,
                           double count, double* loss_sum, void* scratch, void* stream) {
    if (!bt || !net || !grad || !scratch) return fail(QS_EINVAL, "qs_ppo_grad: NULL argument");
    if (which != QS_PPO_ACTOR && which != QS_PPO_CRITIC) return fail(QS_EINVAL, "qs_ppo_grad: which must be QS_PPO_ACTOR or QS_PPO_CRITIC");
    if (bt->n_envs < 1 || bt->horizon < 1 || !bt->hist0 || !bt->entries || !bt->weight) return fail(QS_EINVAL, "qs_ppo_grad: bad batch");
    if (which == QS_PPO_ACTOR && (!bt->actions || !bt->logp_old || !bt->adv || !(sigma > 0.f)))
        return fail(QS_EINVAL, "qs_ppo_grad: the actor needs actions, logp_old, adv and sigma > 0");
    if (which == QS_PPO_CRITIC && !bt->ret) return fail(QS_EINVAL, "qs_ppo_grad: the critic needs returns");
    if (!(count > 0)) return fail(QS_EINVAL, "qs_ppo_grad: count must be > 0");
    const float* w[6] = {net->w1, net->b1, net->w2, net->b2, net->w3, net->b3};
    float* gr[6] = {(float*)grad->w1, (float*)grad->b1, (float*)grad->w2, (float*)grad->b2, (float*)grad->w3, (float*)grad->b3};
    for (int k = 0; k < 6; ++k) if (!w[k] || !gr[k]) return fail(QS_EINVAL, "qs_ppo_grad: NULL weight / gradient pointer");
    cudaStream_t st = (cudaStream_t)stream;
    static std::atomic<int> attr_done{0};                        // (per-device contexts share the attribute of the current device only:
    int dev = 0;                                                 //  set it on every device the entry point is used on)
    QS_CUDA(cudaGetDevice(&dev));
    if (!(attr_done.load() & (1 << (dev & 31)))) {
        QS_CUDA(cudaFuncSetAttribute(ppo::ppo_grad_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, ppo::kSmemBytes));
        QS_CUDA(cudaFuncSetAttribute(ppo::ppo_grad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ppo::kSmemBytes));
        attr_done.fetch_or(1 << (dev & 31));
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int out_dim = which == QS_PPO_ACTOR ? 4 : 1;
    float* tmp = (float*)scratch;                                // >= 516 floats
    ppo::k_pack_out<<<1, ppo::kH, 0, st>>>(net->w3, net->b3, out_dim, tmp);
    QS_CUDA(cudaMemcpyToSymbolAsync(ppo::c_w3, tmp, sizeof(float) * ppo::kH * 4, (size_t)which * sizeof(float) * ppo::kH * 4, cudaMemcpyDeviceToDevice, st));
    QS_CUDA(cudaMemcpyToSymbolAsync(ppo::c_b3, tmp + ppo::kH * 4, sizeof(float) * 4, (size_t)which * sizeof(float) * 4, cudaMemcpyDeviceToDevice, st));
    ppo::Batch b{bt->n_envs, bt->horizon, bt->flags, bt->hist0, bt->entries, bt->actions, (float*)bt->logp_old, bt->adv, bt->ret, bt->weight};
    ppo::Net nw{net->w1, net->b1, net->w2, net->b2, net->w3, net->b3};
    ppo::Grad g{gr[0], gr[1], gr[2], gr[3], gr[4], gr[5], loss_sum};
    const int64_t tiles = (bt->n_envs + ppo::kM - 1) / ppo::kM;
    const int grid = (int)(tiles < sms ? tiles : sms);
    const float inv_count = (float)(1.0 / count);
    if (which == QS_PPO_ACTOR) ppo::ppo_grad_kernel<0><<<grid, ppo::kM, ppo::kSmemBytes, st>>>(b, nw, g, sigma, eps_clip, inv_count);
    else ppo::ppo_grad_kernel<1><<<grid, ppo::kM, ppo::kSmemBytes, st>>>(b, nw, g, sigma, eps_clip, inv_count);
    QS_CUDA(cudaGetLastError());
    return QS_OK;
}

extern "C" int qs_adam_step(int64_t n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int32_t step, float lr,
                            float beta1, float beta2, float eps, void* stream) {
    if (n < 1 || !param || !grad || !exp_avg || !exp_avg_sq || step < 1) return fail(QS_EINVAL, "qs_adam_step: bad argument");
    const float bc1 = 1.f - powf(beta1, (float)step), bc2 = sqrtf(1.f - powf(beta2, (float)step));
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    ppo::adam_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(n, param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, bc1, bc2);
    QS_CUDA(cudaGetLastError());
    return QS_OK;
}
