// Generalised advantage estimation over a rollout buffer laid out like the simulator's outputs ([T, n], env
// index fastest): the consumer that sits directly behind futbol_rollout / futbol_step in the reference's flow
// (stable-baselines PPO2 runner, colab_notebook.ipynb:852; gamma 0.99, lambda 0.95 in the saved models' JSON).
//   delta_t = r_t + gamma * V_{t+1} * (1 - done_t) - V_t
//   A_t     = delta_t + gamma * lambda * (1 - done_t) * A_{t+1},   A_T = 0;   R_t = A_t + V_t
// One thread per environment walks its column backwards; every access of a warp is a run of 32 consecutive
// elements.  HBM-bound: 4 (r) + 1 (done) + 4 (V) read, 8 written per element = 17 bytes.
#include <cuda_runtime.h>
#include <stdint.h>
#include <cuda_bf16.h>
#include "../../include/futbol_b200.h"
#include "sampler.cuh"

namespace futbol {

__global__ void __launch_bounds__(256) gae_kernel(const float *__restrict__ reward, const uint8_t *__restrict__ done,
                                                  const float *__restrict__ value, float gamma, float lam,
                                                  float *__restrict__ adv, float *__restrict__ ret, int T, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a = 0.0f;
    float v_next = value[(size_t)T * n + i];
#pragma unroll 4
    for (int t = T - 1; t >= 0; --t) {
        const size_t k = (size_t)t * n + i;
        const float nd = done[k] ? 0.0f : 1.0f;
        const float v = value[k];
        const float delta = __fsub_rn(__fadd_rn(reward[k], __fmul_rn(__fmul_rn(gamma, v_next), nd)), v);
        a = __fadd_rn(delta, __fmul_rn(__fmul_rn(__fmul_rn(gamma, lam), nd), a));
        adv[k] = a;
        ret[k] = __fadd_rn(a, v);
        v_next = v;
    }
}

// Minibatch gather: row idx[j] of every rollout-buffer column into row j of the minibatch, ONE pass over the index
// (stable-baselines' runner slices its flattened [T * n] buffers per minibatch, colab_notebook.ipynb:852; on the device
// that is six indexing launches that each re-read the permutation).  One warp per sample: lanes copy the observation
// row (obs_dim floats, a contiguous run), the last lanes the action byte and up to four float scalars (old log-prob,
// advantage, return, value).  HBM-bound: (4 obs_dim + 17) bytes read and written per sample + 8 of index.
__global__ void __launch_bounds__(256) gather_minibatch_kernel(const long long *__restrict__ idx, long long m, long long rows,
                                                               const float *__restrict__ obs, int obs_dim, float *__restrict__ obs_out,
                                                               const uint8_t *__restrict__ act, uint8_t *__restrict__ act_out,
                                                               const float *__restrict__ c0, float *__restrict__ c0_out,
                                                               const float *__restrict__ c1, float *__restrict__ c1_out,
                                                               const float *__restrict__ c2, float *__restrict__ c2_out,
                                                               const float *__restrict__ c3, float *__restrict__ c3_out,
                                                               unsigned long long *__restrict__ bad)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long j = warp; j < m; j += n_warps) {
        const long long r = idx[j];
        if (r < 0 || r >= rows) {                                 // never dereferenced: counted, the row is zero-filled
            if (lane == 0 && bad != nullptr) atomicAdd(bad, 1ull);
            if (obs != nullptr) for (int k = lane; k < obs_dim; k += 32) obs_out[j * obs_dim + k] = 0.0f;
            continue;
        }
        if (obs != nullptr) for (int k = lane; k < obs_dim; k += 32) obs_out[j * obs_dim + k] = __ldg(obs + r * obs_dim + k);
        if (lane == 31 && act != nullptr) act_out[j] = act[r];
        if (lane == 30 && c0 != nullptr) c0_out[j] = c0[r];
        if (lane == 29 && c1 != nullptr) c1_out[j] = c1[r];
        if (lane == 28 && c2 != nullptr) c2_out[j] = c2[r];
        if (lane == 27 && c3 != nullptr) c3_out[j] = c3[r];
    }
}

cudaError_t launch_gather_minibatch(const long long *idx, long long m, long long rows, const float *obs, int obs_dim, float *obs_out,
                                    const uint8_t *act, uint8_t *act_out, const float *c0, float *c0_out, const float *c1,
                                    float *c1_out, const float *c2, float *c2_out, const float *c3, float *c3_out,
                                    unsigned long long *bad, cudaStream_t st)
{
    long long blocks = (m + 7) / 8;                               // 8 warps per block, one sample per warp per trip
    if (blocks > 148 * 64) blocks = 148 * 64;
    gather_minibatch_kernel<<<(int)blocks, 256, 0, st>>>(idx, m, rows, obs, obs_dim, obs_out, act, act_out, c0, c0_out, c1, c1_out,
                                                         c2, c2_out, c3, c3_out, bad);
    return cudaGetLastError();
}

// Categorical sampling between the policy's forward pass and futbol_step (the row rule lives in sampler.cuh, which also
// compiles for the host: tests/host_shim/sampler_host.cpp): one thread per row, ONE launch instead of softmax + multinomial +
// gather + cast.  The row (n_actions <= 32 values) is read three times; the second and third pass hit L1.
template <typename T> __device__ __forceinline__ float logit_at(const T *p, int k);
template <> __device__ __forceinline__ float logit_at<float>(const float *p, int k) { return __ldg(p + k); }
template <> __device__ __forceinline__ float logit_at<__nv_bfloat16>(const __nv_bfloat16 *p, int k) { return __bfloat162float(p[k]); }

template <typename T>
__global__ void __launch_bounds__(256) sample_actions_kernel(const T *__restrict__ logits, long long n, int n_actions, PhiloxKey key,
                                                             const unsigned long long *__restrict__ t_base, unsigned long long t_off,
                                                             uint8_t *__restrict__ actions, float *__restrict__ logp)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T *row = logits + i * n_actions;
    const unsigned long long t = (t_base != nullptr ? *t_base : 0ull) + t_off;
    int pick;
    float lp;
    sample_row([row](int k) { return logit_at(row, k); }, n_actions, key, t, (unsigned long long)i, pick, lp);
    actions[i] = (uint8_t)pick;
    if (logp != nullptr) logp[i] = lp;
}

cudaError_t launch_sample_actions(const void *logits, int bf16, long long n, int n_actions, unsigned long long seed,
                                  const unsigned long long *t_base, unsigned long long t_off, uint8_t *actions, float *logp, cudaStream_t st)
{
    const PhiloxKey key = philox_expand_key(seed);
    const int blocks = (int)((n + 255) / 256);
    if (bf16) sample_actions_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16 *)logits, n, n_actions, key, t_base, t_off, actions, logp);
    else sample_actions_kernel<float><<<blocks, 256, 0, st>>>((const float *)logits, n, n_actions, key, t_base, t_off, actions, logp);
    return cudaGetLastError();
}

cudaError_t launch_gae(const float *reward, const uint8_t *done, const float *value, float gamma, float lam, float *adv,
                       float *ret, int T, int n, cudaStream_t st)
{
    gae_kernel<<<(n + 255) / 256, 256, 0, st>>>(reward, done, value, gamma, lam, adv, ret, T, n);
    return cudaGetLastError();
}

}  // namespace futbol
