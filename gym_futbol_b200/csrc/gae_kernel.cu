// Generalised advantage estimation over a rollout buffer laid out like the simulator's outputs ([T, n], env
// index fastest): the consumer that sits directly behind futbol_rollout / futbol_step in the reference's flow
// (stable-baselines PPO2 runner, colab_notebook.ipynb:852; gamma 0.99, lambda 0.95 in the saved models' JSON).
//   delta_t = r_t + gamma * V_{t+1} * (1 - done_t) - V_t
//   A_t     = delta_t + gamma * lambda * (1 - done_t) * A_{t+1},   A_T = 0;   R_t = A_t + V_t
// One thread per environment walks its column backwards; every access of a warp is a run of 32 consecutive
// elements.  HBM-bound: 4 (r) + 1 (done) + 4 (V) read, 8 written per element = 17 bytes.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/futbol_b200.h"

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

cudaError_t launch_gae(const float *reward, const uint8_t *done, const float *value, float gamma, float lam, float *adv,
                       float *ret, int T, int n, cudaStream_t st)
{
    gae_kernel<<<(n + 255) / 256, 256, 0, st>>>(reward, done, value, gamma, lam, adv, ret, T, n);
    return cudaGetLastError();
}

}  // namespace futbol
