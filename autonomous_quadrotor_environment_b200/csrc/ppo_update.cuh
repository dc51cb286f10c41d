// ppo_update.cuh — the network update of the reference's PPO trainer (environment/controller/ppo.py:143-209, model.py:19-88) as
// hand-written sm_100a kernels on the time-major rollout buffers ([K][C][N], never leaving HBM):
//
//   ppo_grad_kernel<NET>   one full-batch gradient of the clipped-surrogate loss for ONE network (NET = 0 actor 75-128-128-4 with a
//                          tanh output, NET = 1 critic 75-128-128-1): forward AND backward of the MLP on tcgen05 tensor cores (BF16
//                          operands, FP32 TMEM accumulators), FP32 gradients accumulated over the launch
//   adam_kernel            torch.optim.Adam's step (ppo.py:105, default eps 1e-8, no weight decay) on FP32 master weights
//
// One CTA of 512 threads per SM walks work units = (128-env tile, chunk of recorded steps); sample = TMEM lane = tid & 127, and the four
// threads of a sample (tid >> 7) own 32 of the 128 hidden columns each.  Every activation / gradient tile is written ONCE, row = sample,
// with 16-byte shared-memory stores in the canonical [row][col] layout of umma.cuh, and serves the tensor core twice: as a K-major
// operand (K = hidden index) of the forward / input-gradient products and as an MN-major operand (K = sample index) of the
// weight-gradient products — no transposed copies exist.  Per (tile, step), X = [128 x 80] the five history entries oldest first
// (15 floats per entry + pad = 1: bias column), double-buffered: a sample's row is shifted by one entry into the other buffer while the
// products of this step still read the current one:
//   Z1 = X W1^T             5 x K16, N = 128                                                                           -> TMEM Z
//   H1 = tanh(Z1)           -> smem H1 [s][144]: 128 activations + the constant columns [1 1 0..] of the bias fold
//   Z2 = [H1 1 1] [W2 b2_hi b2_lo]^T      9 x K16, N = 128                                                              -> TMEM Z
//   H2 = tanh(Z2) -> smem H2 [s][j]; out = W3 H2 + b3 on the FP32 pipe (partial sums of the four column owners exchanged through
//   shared memory); the loss and dZ3 per sample in registers; dZ2 = (W3^T dZ3)(1 - H2^2) -> smem dZ [s][j]
//   group 1:  dW3 += H2^T dZ3  (A = H2 MN-major, B = dZ3 [s][16] MN-major, N = 16)
//             dH1  = dZ2 W2    (A = dZ K-major, B = W2 [j2][j1] MN-major, N = 128)                                       -> TMEM Z
//   group 2:  [dW2 | db2] += dZ2^T [H1 1 1]   (A = dZ MN-major, B = the H1 tile MN-major, N = 144) — runs while the threads compute
//   dZ1 = dH1 (1 - H1^2)    -> smem, into the H2 tile (dW3 has completed)
//   group 3:  next step's Z1 (waited for first), then dW1 += dZ1^T X (A = dZ1 MN-major, B = X MN-major, N = 80; the pad columns of X
//             make column 15 of every 16-wide block db1) — runs under the next step's first epilogue.
// The weight-gradient accumulators stay in TENSOR MEMORY for the whole launch (240 columns next to the 128 of Z and 64 of parked
// tanh' factors) and are added to the FP32 gradient buffers in HBM once per CTA.
#pragma once
#include "umma.cuh"

namespace ppo {

constexpr int kM = 128;          // samples per tile = UMMA M / K
constexpr int kH = 128;          // hidden width
constexpr int kSlots = 5, kSlotK = 16, kKin = kSlots * kSlotK;   // 80
constexpr int kH1ext = kH + 16;                                  // H1 / W2 tiles carry the 16-wide bias block of layer 2 as columns 128..143
constexpr int kSplit = 4, kThreads = kM * kSplit, kCW = kH / kSplit;   // 512 threads, 32 hidden columns per thread

struct Batch {
    int64_t N;                   // envs
    int32_t K;                   // recorded steps
    int32_t flags;               // QS_PPO_RECORD_LOGP: write logp_old from this forward pass (ratio = 1) instead of reading it
    const float* hist0;          // [75][N]   dl_in_gen buffer at rollout start, oldest entry first
    const float* entries;        // [K][15][N] history entry pushed after step t: [action(4), v(3), q(4), dq(4)]; nullptr: built from
    const float* obs;            // [K][14][N] the recorded observations (rows 1,3,5 and 6..13) and `actions`
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

// shared-memory map (bytes); every tile is [row][col] in umma_canon_offset(row, col, cols)
constexpr int oX = 0;                                 // 2 x [128 s][80]  history, oldest entry first (A of layer 1, B of dW1), ping-pong
constexpr int oH1 = oX + 2 * kM * kKin * 2;           // [128 s][144]   H1 | 1 1 0.. (A of layer 2, B of dW2 | db2)
constexpr int oH2 = oH1 + kM * kH1ext * 2;            // [128 s][128 j] H2 (A of dW3), then dZ1 (A of dW1)
constexpr int oDZ = oH2 + kM * kH * 2;                // [128 s][128 j] dZ2 (A of dW2 | db2 and of dH1)
constexpr int oW1 = oDZ + kM * kH * 2;                // [128 j][80]    (B of layer 1; column = 16 age + e)
constexpr int oW2 = oW1 + kH * kKin * 2;              // [128 j2][144]  W2 | b2_hi b2_lo 0.. (B of layer 2 K-major, B of dH1 MN-major)
constexpr int oDZ3 = oW2 + kH * kH1ext * 2;           // [128 s][16]    dZ3 in columns 0..OUT-1 (B of dW3)
constexpr int oPart = oDZ3 + kM * 16 * 2;             // float [4 part][4 k][128 s]  partial sums of the output layer
constexpr int oIn = oPart + kSplit * 4 * kM * 4;      // float [10][128 s]  per-sample inputs of the loss: act(4), logp_old(4), adv | ret, weight
constexpr int oW3 = oIn + 10 * kM * 4;                 // float4 [128 j] = w3[0..3][j] (critic: .x only), then float [4] b3: the FP32 output layer
constexpr int kSmemBytes = oW3 + kH * 16 + 16;

// TMEM column map
constexpr uint32_t cZ = 0, cDW2 = 128, cDB2 = cDW2 + 128, cDW1 = 272, cDW3 = 352, cT1 = 368, kCols = 512;   // [dW2 | db2] = 144 columns; cT1: 64 columns, 1 - H1^2 packed BF16

using namespace qs;

__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src) {        // global -> shared without a register stop-over
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void st_bf16(unsigned char* base, uint32_t off, float x) {
    *reinterpret_cast<__nv_bfloat16*>(base + off) = __float2bfloat16(x);
}
// 32 consecutive columns of a thread's own row: four 16-byte stores / loads (a quarter-warp covers 128 contiguous bytes: no conflicts)
__device__ __forceinline__ void st_row32(unsigned char* tile, int ext, int row, int col0, const float v[32]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint4 q;
        q.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]); q.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
        q.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]); q.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
        *reinterpret_cast<uint4*>(tile + umma_canon_offset(row, col0 + 8 * i, ext)) = q;
    }
}
// element q (0..14) of the history entry pushed after step t (dl_auxiliary.py:27-30: action(4), obs[1,3,5], obs[6:14])
__device__ __forceinline__ float entry_value(const Batch& b, int t, int q, int64_t n) {
    if (b.entries) return b.entries[((int64_t)t * 15 + q) * b.N + n];
    if (q < 4) return b.actions[((int64_t)t * 4 + q) * b.N + n];
    const int row = q < 7 ? 2 * (q - 4) + 1 : q - 1;
    return b.obs[((int64_t)t * 14 + row) * b.N + n];
}

template <int NET>
__global__ void __launch_bounds__(kThreads, 1)
ppo_grad_kernel(const __grid_constant__ Batch b, const __grid_constant__ Net w, const __grid_constant__ Grad g, float sigma,
                float eps_clip, float inv_count, int chunk_len, int n_chunks) {
    constexpr int OUT = NET == 0 ? 4 : 1;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __shared__ float s_loss;
    const int tid = threadIdx.x, s = tid & (kM - 1), part = tid >> 7, warp = tid >> 5;
    float* s_part = reinterpret_cast<float*>(smem + oPart);
    float* s_in = reinterpret_cast<float*>(smem + oIn);
    for (int idx = tid; idx < 10 * kM; idx += kThreads) s_in[idx] = 0.f;
    float* s_w3 = reinterpret_cast<float*>(smem + oW3);          // [j][4], every lane of a warp reads the same address (broadcast)
    float* s_b3 = s_w3 + kH * 4;
    for (int idx = tid; idx < kH * 4; idx += kThreads) s_w3[idx] = (idx & 3) < OUT ? w.w3[(idx & 3) * kH + (idx >> 2)] : 0.f;
    if (tid < 4) s_b3[tid] = tid < OUT ? w.b3[tid] : 0.f;

    // ---- one-time: FP32 master weights -> BF16 operand tiles
    for (int idx = tid; idx < kH * kKin; idx += kThreads) {      // W1: input 15*age + e -> column 16*age + e; pad columns carry b1
        const int n = idx / kKin, kk = idx % kKin, a = kk / kSlotK, e = kk % kSlotK;
        float x = 0.f;
        if (e < 15) x = w.w1[n * 75 + a * 15 + e];
        else if (a == 0) x = __bfloat162float(__float2bfloat16(w.b1[n]));
        else if (a == 1) x = w.b1[n] - __bfloat162float(__float2bfloat16(w.b1[n]));
        st_bf16(smem + oW1, umma_canon_offset(n, kk, kKin), x);
    }
    for (int idx = tid; idx < kH * kH; idx += kThreads) {
        const int n = idx / kH, kk = idx % kH;
        st_bf16(smem + oW2, umma_canon_offset(n, kk, kH1ext), w.w2[n * kH + kk]);    // [j2][j1]
    }
    for (int idx = tid; idx < kH * 16; idx += kThreads) {
        const int n = idx / 16, kk = idx % 16;
        const float bb = w.b2[n], hi = __bfloat162float(__float2bfloat16(bb));
        st_bf16(smem + oH1, umma_canon_offset(n, kH + kk, kH1ext), kk < 2 ? 1.f : 0.f);
        st_bf16(smem + oW2, umma_canon_offset(n, kH + kk, kH1ext), kk == 0 ? hi : (kk == 1 ? bb - hi : 0.f));
        st_bf16(smem + oDZ3, umma_canon_offset(n, kk, 16), 0.f);
    }
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); s_loss = 0.f; }
    if (tid < 32) tmem_alloc(&tmem_slot, kCols);
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_slot;
    const uint32_t lane = tb + ((uint32_t)((warp & 3) * 32) << 16);     // a warp reads the 32 TMEM lanes of its quarter
    const uint32_t uX = smem_u32(smem + oX), uH1 = smem_u32(smem + oH1), uH2 = smem_u32(smem + oH2), uDZ = smem_u32(smem + oDZ),
                   uW1 = smem_u32(smem + oW1), uW2 = smem_u32(smem + oW2), uDZ3 = smem_u32(smem + oDZ3);
    constexpr uint32_t kXbytes = kM * kKin * 2;
    uint32_t phase = 0;
    bool have_acc = false;                                       // the weight-gradient accumulators hold something
    float loss_local = 0.f;
    float db3_local[4] = {0.f, 0.f, 0.f, 0.f};                   // db3[k] = sum_s dZ3[s][k]  (part 0 only)
    const float inv_var = 1.f / (sigma * sigma);
    const float log_norm = -__logf(sigma) - 0.918938533f;
    const int col0 = part * kCW;

    auto commit = [&]() { if (tid == 0) umma_commit(&bar); };
    auto wait = [&]() { mbar_wait(&bar, phase); phase ^= 1; tc_fence_after(); };
    auto sync_before_issue = [&]() {                             // operands written by all threads -> visible to the tensor core
        tc_fence_before();
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) tc_fence_after();
    };
    auto issue_layer1 = [&](int cur) { umma_gemm_k(tb + cZ, uX + (uint32_t)cur * kXbytes, kKin, 0, uW1, kKin, 0, kKin, kH, false); };
    // the four history values this thread stages per entry: elements 4*part .. 4*part+3 of the 16-wide block (element 15 = 1)
    auto entry_store = [&](int buf, int age, const float x[4]) {
        uint2 q;
        q.x = pack_bf16x2(x[0], x[1]); q.y = pack_bf16x2(x[2], x[3]);
        *reinterpret_cast<uint2*>(smem + oX + buf * kXbytes + umma_canon_offset(s, age * kSlotK + 4 * part, kKin)) = q;
    };

    const int64_t n_tiles = (b.N + kM - 1) / kM;
    const int64_t n_units = n_tiles * n_chunks;
    for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
        const int64_t tile = u / n_chunks;
        const int t0 = (int)(u % n_chunks) * chunk_len;
        const int t1 = t0 + chunk_len < b.K ? t0 + chunk_len : b.K;
        if (t0 >= t1) continue;
        const int64_t n = tile * kM + s;
        const bool active = n < b.N;
        // history at step t0 = seq[t0 .. t0+4], seq = [the five entries of hist0 (oldest first), entries[0], entries[1], ...]
        // (every MMA of the previous unit has completed: the unit ends with a wait)
#pragma unroll
        for (int a = 0; a < kSlots; ++a) {
            const int i = t0 + a;
            float x[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int q = 4 * part + r;
                x[r] = q < 15 ? (active ? (i < kSlots ? b.hist0[(int64_t)(i * 15 + q) * b.N + n] : entry_value(b, i - kSlots, q, n)) : 0.f) : 1.f;
            }
            entry_store(0, a, x);
        }
        int cur = 0;
        sync_before_issue();
        if (tid == 0) issue_layer1(cur);
        commit();
        for (int t = t0; t < t1; ++t) {
            const int64_t tn = (int64_t)t * b.N + n;
            // ---- this step's per-sample inputs (in flight while the MMAs run)
            float e_new[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int q = 4 * part + r;
                e_new[r] = q < 15 ? (active ? entry_value(b, t, q, n) : 0.f) : 1.f;
            }
            // the loss's inputs go straight to shared memory (cp.async), one item per column owner; rows of padding samples stay 0
            const bool record = NET == 0 && (b.flags & QS_PPO_RECORD_LOGP) != 0;   // memory.logprobs of policy_old == policy (ppo.py:206)
            if (active) {
                if (NET == 0) {
                    cp_async4(&s_in[part * kM + s], &b.actions[((int64_t)t * 4 + part) * b.N + n]);
                    if (!record) cp_async4(&s_in[(4 + part) * kM + s], &b.logp_old[((int64_t)t * 4 + part) * b.N + n]);
                    if (part == 0) cp_async4(&s_in[8 * kM + s], &b.adv[tn]);
                } else if (part == 0) cp_async4(&s_in[8 * kM + s], &b.ret[tn]);
                if (part == 1) cp_async4(&s_in[9 * kM + s], &b.weight[tn]);
            }
            // ================= forward
            wait();                                               // Z1 (and the previous step's dW1)
            float acc[32];
            tmem_ld_32x32b_x32(lane + cZ + (uint32_t)col0, acc);
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = tanh_fast(acc[i]);
            st_row32(smem + oH1, kH1ext, s, col0, acc);
            {   // tanh' = 1 - H1^2 from the FP32 activation, parked in tensor memory until dZ1 (BF16 of the FACTOR: a saturated unit's
                // 1 - h^2 computed from a BF16-rounded h would be off by tens of percent)
                uint32_t tp[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) tp[i] = pack_bf16x2(1.f - acc[2 * i] * acc[2 * i], 1.f - acc[2 * i + 1] * acc[2 * i + 1]);
                tmem_st_32x32b_x16(lane + cT1 + (uint32_t)(col0 >> 1), tp);
                tmem_st_wait();
            }
            sync_before_issue();
            if (tid == 0) umma_gemm_k(tb + cZ, uH1, kH1ext, 0, uW2, kH1ext, 0, kH1ext, kH, false);
            commit();
            wait();
            // H2 and the output layer on the FP32 pipe: this thread's 32 hidden units
            tmem_ld_32x32b_x32(lane + cZ + (uint32_t)col0, acc);
            float po[OUT];
#pragma unroll
            for (int k = 0; k < OUT; ++k) po[k] = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                acc[i] = tanh_fast(acc[i]);
                if (OUT == 4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(&s_w3[(col0 + i) * 4]);
                    po[0] = fmaf(acc[i], w4.x, po[0]); po[1 % OUT] = fmaf(acc[i], w4.y, po[1 % OUT]);
                    po[2 % OUT] = fmaf(acc[i], w4.z, po[2 % OUT]); po[3 % OUT] = fmaf(acc[i], w4.w, po[3 % OUT]);
                } else po[0] = fmaf(acc[i], s_w3[(col0 + i) * 4], po[0]);
            }
            st_row32(smem + oH2, kH, s, col0, acc);
#pragma unroll
            for (int k = 0; k < OUT; ++k) s_part[(part * 4 + k) * kM + s] = po[k];
            cp_async_wait_all();
            __syncthreads();
            const float wgt = s_in[9 * kM + s] * inv_count;
            float out[OUT];
#pragma unroll
            for (int k = 0; k < OUT; ++k)
                out[k] = s_b3[k] + ((s_part[(0 * 4 + k) * kM + s] + s_part[(1 * 4 + k) * kM + s]) + (s_part[(2 * 4 + k) * kM + s] + s_part[(3 * 4 + k) * kM + s]));
            // ================= loss and its gradient w.r.t. the pre-activation of the output layer (per sample, x weight / count)
            float dz3[4] = {0.f, 0.f, 0.f, 0.f};
            if (NET == 0) {
                float mean[4], lp[4], act[4], lp_new = 0.f, lp_old = 0.f;
                const float adv = s_in[8 * kM + s];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    mean[k] = tanh_fast(out[k]);                                              // model.py:33-34
                    act[k] = s_in[k * kM + s];
                    const float d = act[k] - mean[k];
                    lp[k] = -0.5f * d * d * inv_var + log_norm;                               // Normal.log_prob, model.py:82
                    lp_new += lp[k];
                    lp_old += record ? lp[k] : s_in[(4 + k) * kM + s];
                }
                if (record && active && part == 0) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) b.logp_old[((int64_t)t * 4 + k) * b.N + n] = lp[k];
                }
                const float ratio = __expf(lp_new - lp_old);                                  // ppo.py:187
                const float surr1 = ratio * adv;                                              // :192
                const float rc = fminf(fmaxf(ratio, 1.f - eps_clip), 1.f + eps_clip);
                const float surr2 = rc * adv;                                                 // :193
                // torch.min(surr1, surr2): the gradient flows through the selected branch; clamp passes it inside [1-eps, 1+eps] only
                const bool through = (surr1 <= surr2) || (ratio >= 1.f - eps_clip && ratio <= 1.f + eps_clip);
                const float dl_dlp = through ? -adv * ratio : 0.f;                            // d(-min)/d(sum of log-probs)
                if (part == 0) loss_local += wgt * (-fminf(surr1, surr2));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    dz3[k] = wgt * dl_dlp * (act[k] - mean[k]) * inv_var * (1.f - mean[k] * mean[k]);
            } else {
                const float d = out[0] - s_in[8 * kM + s];                                    // 0.5 * MSE, ppo.py:194
                if (part == 0) loss_local += wgt * 0.5f * d * d;
                dz3[0] = wgt * d;
            }
            if (part == 0) {
#pragma unroll
                for (int k = 0; k < OUT; ++k) db3_local[k] += dz3[k];
                uint2 q;
                q.x = pack_bf16x2(dz3[0], dz3[1]); q.y = pack_bf16x2(dz3[2], dz3[3]);
                *reinterpret_cast<uint2*>(smem + oDZ3 + umma_canon_offset(s, 0, 16)) = q;
            }
            // ================= backward
            // dZ2 = (W3^T dZ3)(1 - H2^2)
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float dh;
                if (OUT == 4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(&s_w3[(col0 + i) * 4]);
                    dh = fmaf(dz3[3], w4.w, fmaf(dz3[2], w4.z, fmaf(dz3[1], w4.y, dz3[0] * w4.x)));
                } else dh = dz3[0] * s_w3[(col0 + i) * 4];
                acc[i] = dh * (1.f - acc[i] * acc[i]);
            }
            st_row32(smem + oDZ, kH, s, col0, acc);
            sync_before_issue();
            if (tid == 0) {
                umma_gemm_mn(tb + cDW3, uH2, kH, 0, uDZ3, 16, 0, 16, have_acc);       // dW3[j][k]   += sum_s H2[s][j] dZ3[s][k]
                umma_gemm_k_mn(tb + cZ, uDZ, kH, uW2, kH1ext, 0, kH, kH, false);        // dH1[s][j1]   = sum_j2 dZ2[s][j2] W2[j2][j1]
            }
            commit();
            // [dW2 | db2][j2][j1 | 128] += sum_s dZ2[s][j2] [H1 | 1][s][..]: no commit of its own — it runs under the dZ1 epilogue below and
            // is covered by the next commit
            if (tid == 0) umma_gemm_mn(tb + cDW2, uDZ, kH, 0, uH1, kH1ext, 0, kH1ext, have_acc);
            wait();
            // dZ1 = dH1 (1 - H1^2) into the H2 tile (dW3, its reader, has completed)
            {
                float tp[16];
                tmem_ld_32x32b_x32(lane + cZ + (uint32_t)col0, acc);
                tmem_ld_32x32b_x16(lane + cT1 + (uint32_t)(col0 >> 1), tp);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint32_t pk = __float_as_uint(tp[i]);
                    acc[2 * i] *= __uint_as_float(pk << 16);
                    acc[2 * i + 1] *= __uint_as_float(pk & 0xFFFF0000u);
                }
                st_row32(smem + oH2, kH, s, col0, acc);
            }
            // next step's input (dl_auxiliary.py:25-32): this sample's row moves up by one entry into the other buffer (this thread: the
            // 16 columns of entry part+1 -> entry part), the entry recorded after step t becomes the newest
            {
                const unsigned char* src = smem + oX + cur * kXbytes;
                unsigned char* dst = smem + oX + (cur ^ 1) * kXbytes;
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    *reinterpret_cast<uint4*>(dst + umma_canon_offset(s, part * kSlotK + 8 * i, kKin)) =
                        *reinterpret_cast<const uint4*>(src + umma_canon_offset(s, (part + 1) * kSlotK + 8 * i, kKin));
                entry_store(cur ^ 1, kSlots - 1, e_new);
            }
            sync_before_issue();
            if (tid == 0 && t + 1 < t1) issue_layer1(cur ^ 1);   // first in the queue after dW2: the next epilogue waits for it alone
            commit();
            // dW1[j][16 a + e] += sum_s dZ1[s][j] X[s][16 a + e]; covered by the next commit (layer 2 of the next step / the unit's end)
            if (tid == 0) umma_gemm_mn(tb + cDW1, uH2, kH, 0, uX + (uint32_t)cur * kXbytes, kKin, 0, kKin, have_acc);
            have_acc = true;
            cur ^= 1;
        }
        wait();                                                   // the last step's commit ...
        commit();
        wait();                                                   // ... and the unit's last dW1
    }
    // ---- accumulators -> HBM gradients (thread = row j of every accumulator; the four parts split the columns)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (have_acc) {
        const int j = s;
        float acc[32];
        tmem_ld_32x32b_x32(lane + cDW2 + (uint32_t)col0, acc);
#pragma unroll
        for (int i = 0; i < 32; ++i) atomicAdd(&g.w2[j * kH + col0 + i], acc[i]);
        float d1[16];
        for (int a = part; a < kSlots; a += kSplit) {
            tmem_ld_32x32b_x16(lane + cDW1 + (uint32_t)(a * kSlotK), d1);
#pragma unroll
            for (int e = 0; e < 15; ++e) atomicAdd(&g.w1[j * 75 + a * 15 + e], d1[e]);
            if (a == 0) atomicAdd(&g.b1[j], d1[15]);             // the bias rides in the pad column (every slot's pad column sees the same sum)
        }
        if (part == 1) {
            tmem_ld_32x32b_x16(lane + cDB2, d1);
            atomicAdd(&g.b2[j], d1[0]);
        }
        if (part == 2) {
            tmem_ld_32x32b_x16(lane + cDW3, d1);
#pragma unroll
            for (int k = 0; k < OUT; ++k) atomicAdd(&g.w3[k * kH + j], d1[k]);
        }
    }
    // b3 gradient and the loss: per-thread sums (part 0) -> warp shuffle -> HBM
    if (part == 0) {
#pragma unroll
        for (int k = 0; k < OUT; ++k) {
            float x = db3_local[k];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
            if ((tid & 31) == 0 && x != 0.f) atomicAdd(&g.b3[k], x);
        }
        float x = loss_local;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
        if ((tid & 31) == 0) atomicAdd(&s_loss, x);
    }
    __syncthreads();
    if (tid == 0 && g.loss) atomicAdd(g.loss, (double)s_loss);
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tb, kCols);
}

// ---------------------------------------------------------------------------------------------------------------
// self-test of the MN-major operand path: mode 0: D[128][N] = At^T Bt, At [128 k][128 m], Bt [128 k][N] (both MN-major);
// mode 1: D[128][N] = A Bt, A [128 m][128 k] K-major, Bt [128 k][N] MN-major.  Row-major FP32 in / out.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_umma_selftest_mn(int mode, int N, const float* A, const float* B, float* D) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sA = smem;
    unsigned char* sB = smem + 128 * 128 * 2;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int c = 0; c < 128; ++c) st_bf16(sA, umma_canon_offset(tid, c, 128), A[tid * 128 + c]);
    for (int c = 0; c < N; ++c) st_bf16(sB, umma_canon_offset(tid, c, N), B[tid * N + c]);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_base_slot, 128);
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    if (tid == 0) {
        if (mode == 0) umma_gemm_mn(tmem_base, smem_u32(sA), 128, 0, smem_u32(sB), N, 0, N, false);
        else umma_gemm_k_mn(tmem_base, smem_u32(sA), 128, smem_u32(sB), N, 0, 128, N, false);
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
                           double count, double* loss_sum, void* stream) {
    if (!bt || !net || !grad) return fail(QS_EINVAL, "qs_ppo_grad: NULL argument");
    if (which != QS_PPO_ACTOR && which != QS_PPO_CRITIC) return fail(QS_EINVAL, "qs_ppo_grad: which must be QS_PPO_ACTOR or QS_PPO_CRITIC");
    if (bt->n_envs < 1 || bt->horizon < 1 || !bt->hist0 || !bt->weight) return fail(QS_EINVAL, "qs_ppo_grad: bad batch");
    if (!bt->entries && !(bt->obs && bt->actions)) return fail(QS_EINVAL, "qs_ppo_grad: needs entries, or obs and actions to build them from");
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
    ppo::Batch b{bt->n_envs, bt->horizon, bt->flags, bt->hist0, bt->entries, bt->obs, bt->actions, (float*)bt->logp_old, bt->adv, bt->ret, bt->weight};
    ppo::Net nw{net->w1, net->b1, net->w2, net->b2, net->w3, net->b3};
    ppo::Grad g{gr[0], gr[1], gr[2], gr[3], gr[4], gr[5], loss_sum};
    // work units = (128-env tile, chunk of steps): the history at any step is a window of the entries buffer, so a tile's steps split
    // freely; enough units for ~6 per SM (tail balance), chunks of at least 16 steps (a chunk start costs ~1.5 steps)
    const int64_t tiles = (bt->n_envs + ppo::kM - 1) / ppo::kM;
    int64_t n_chunks = (6 * (int64_t)sms + tiles - 1) / tiles;
    const int64_t max_chunks = bt->horizon / 16 > 1 ? bt->horizon / 16 : 1;
    if (n_chunks > max_chunks) n_chunks = max_chunks;
    if (n_chunks < 1) n_chunks = 1;
    const int chunk_len = (int)((bt->horizon + n_chunks - 1) / n_chunks);
    n_chunks = (bt->horizon + chunk_len - 1) / chunk_len;
    const int64_t units = tiles * n_chunks;
    const int grid = (int)(units < sms ? units : sms);
    const float inv_count = (float)(1.0 / count);
    if (which == QS_PPO_ACTOR)
        ppo::ppo_grad_kernel<0><<<grid, ppo::kThreads, ppo::kSmemBytes, st>>>(b, nw, g, sigma, eps_clip, inv_count, chunk_len, (int)n_chunks);
    else
        ppo::ppo_grad_kernel<1><<<grid, ppo::kThreads, ppo::kSmemBytes, st>>>(b, nw, g, sigma, eps_clip, inv_count, chunk_len, (int)n_chunks);
    QS_CUDA(cudaGetLastError());
    return QS_OK;
}

extern "C" int qs_umma_selftest_mn(int mode, int N, const float* A, const float* B, float* D, void* stream) {
    if ((mode != 0 && mode != 1) || N < 16 || N > 128 || (N % 16) || !A || !B || !D)
        return fail(QS_EINVAL, "qs_umma_selftest_mn: mode 0/1, 16 <= N <= 128 (multiple of 16)");
    const size_t smem = (size_t)128 * 128 * 2 + (size_t)128 * N * 2;
    QS_CUDA(cudaFuncSetAttribute(ppo::k_umma_selftest_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ppo::k_umma_selftest_mn<<<1, 128, smem, (cudaStream_t)stream>>>(mode, N, A, B, D);
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
