// ppo_kernels.cuh — the per-iteration reductions of the reference's PPO trainer that sit between the fused rollout and
// the network update, on [K][N] time-major rollout buffers that never leave HBM:
//   * gae_kernel        PPO.get_advantages  environment/controller/ppo.py:125-141  (backward GAE scan, gamma = lambda = 0.99
//                       in the reference) + masked first/second moments of the advantages
//   * adv_norm_kernel   the normalisation of :141  (adv - mean) / (std + 1e-10), population std like np.std
// One thread per env walks its K transitions backwards in register tiles of 8 (the loads of a tile are independent of
// the recursion, so they are all in flight before the first dependent FMA); a warp reads 128 contiguous bytes per row.
#pragma once

constexpr int kGaeTile = 8;

__global__ void __launch_bounds__(256)
gae_kernel(int64_t N, int K, float gamma, float lambda, const float* __restrict__ reward, const float* __restrict__ value,
           const uint8_t* __restrict__ done, float* __restrict__ returns_out, float* __restrict__ adv_out, double* moments) {
    float s_cnt = 0.f, s_sum = 0.f, s_sq = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += stride) {
        float gae = 0.f;                                              // :127
        float v_next = value[(int64_t)K * N + n];                     // values[K]: bootstrap (the reference appends 0, ppo.py:384)
        for (int t1 = K; t1 > 0; t1 -= kGaeTile) {
            const int t0 = t1 - kGaeTile > 0 ? t1 - kGaeTile : 0;
            float r[kGaeTile], v[kGaeTile];
            uint8_t d[kGaeTile];
#pragma unroll
            for (int j = 0; j < kGaeTile; ++j) {
                const int t = t1 - 1 - j;
                if (t >= t0) {
                    r[j] = reward[(int64_t)t * N + n]; v[j] = value[(int64_t)t * N + n]; d[j] = done[(int64_t)t * N + n];
                }
            }
#pragma unroll
            for (int j = 0; j < kGaeTile; ++j) {
                const int t = t1 - 1 - j;
                if (t >= t0) {
                    const float mask = (d[j] & 1) ? 0.f : 1.f;        // np.logical_not(is_terminals)  :162
                    const float delta = r[j] + gamma * v_next * mask - v[j];       // :135
                    gae = delta + gamma * lambda * mask * gae;        // :136
                    const bool valid = !(d[j] & 2);                   // warm-up steps of an asynchronous reset are not transitions
                    returns_out[(int64_t)t * N + n] = gae + v[j];     // :137
                    adv_out[(int64_t)t * N + n] = gae;                // :139  returns - values[:-1]
                    if (valid) { s_cnt += 1.f; s_sum += gae; s_sq += gae * gae; }
                    v_next = v[j];
                }
            }
        }
    }
    __shared__ float sh[3];
    if (threadIdx.x < 3) sh[threadIdx.x] = 0.f;
    __syncthreads();
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s_cnt += __shfl_xor_sync(full, s_cnt, o); s_sum += __shfl_xor_sync(full, s_sum, o); s_sq += __shfl_xor_sync(full, s_sq, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&sh[0], s_cnt); atomicAdd(&sh[1], s_sum); atomicAdd(&sh[2], s_sq); }
    __syncthreads();
    if (threadIdx.x < 3 && sh[0] != 0.f) atomicAdd(&moments[threadIdx.x], (double)sh[threadIdx.x]);
}

__global__ void __launch_bounds__(256)
adv_norm_kernel(int64_t total, const uint8_t* __restrict__ done, const double* __restrict__ moments, float* __restrict__ adv,
                float* __restrict__ weight) {
    const double cnt = moments[0] > 0 ? moments[0] : 1.0;
    const double mean = moments[1] / cnt;
    double var = moments[2] / cnt - mean * mean;
    var = var > 0 ? var : 0;
    const float m = (float)mean, inv = (float)(1.0 / (sqrt(var) + 1e-10));       // :141
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const bool valid = !(done[i] & 2);
        adv[i] = valid ? (adv[i] - m) * inv : 0.f;
        if (weight) weight[i] = valid ? 1.f : 0.f;
    }
}

extern "C" int qs_gae(int64_t n_envs, int32_t horizon, float gamma, float lambda, const float* reward, const float* value,
                      const uint8_t* done, float* returns_out, float* adv_out, double* moments, void* stream) {
    if (n_envs < 1 || horizon < 1 || !reward || !value || !done || !returns_out || !adv_out || !moments)
        return fail(QS_EINVAL, "qs_gae: bad argument");
    int64_t blocks = (n_envs + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    gae_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(n_envs, horizon, gamma, lambda, reward, value, done, returns_out,
                                                             adv_out, moments);
    QS_CUDA(cudaGetLastError());
    return QS_OK;
}

extern "C" int qs_adv_normalize(int64_t total, const uint8_t* done, const double* moments, float* adv, float* weight, void* stream) {
    if (total < 1 || !done || !moments || !adv) return fail(QS_EINVAL, "qs_adv_normalize: bad argument");
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    adv_norm_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(total, done, moments, adv, weight);
    QS_CUDA(cudaGetLastError());
    return QS_OK;
}
