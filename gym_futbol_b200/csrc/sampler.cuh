// One categorical draw from a row of unnormalised log-probabilities (futbol_sample_actions, include/futbol_b200.h): what
// stable-baselines' CategoricalProbabilityDistribution.sample() / .neglogp() do for the reference's Discrete(16) action
// space (colab_notebook.ipynb:852 runner; envs/futbol_env.py:143).
//   m = max l, s = sum exp(l - m), u = a 24-bit Philox uniform of (seed, t, row) on stream 4,
//   action = the first k whose running sum of exp(l - m) exceeds u s (the last action with a non-zero term if rounding never
//   lets it), logp = l[action] - m - log s.
// Device code that also compiles for the host (tests/host_shim/sampler_host.cpp shims the qualifiers and __umulhi).
#pragma once
#include <math.h>
#include <stdint.h>
#include "philox.cuh"

namespace futbol {

constexpr uint32_t kStreamSampler = 4;

template <typename LogitAt>
__device__ __forceinline__ void sample_row(LogitAt at, int n_actions, const PhiloxKey &key, unsigned long long t,
                                           unsigned long long row, int &pick_out, float &logp_out)
{
    float m = at(0);
    for (int k = 1; k < n_actions; ++k) m = fmaxf(m, at(k));
    float s = 0.0f;
    for (int k = 0; k < n_actions; ++k) s += expf(at(k) - m);
    const Philox4 r = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), (uint32_t)row, kStreamSampler ^ ((uint32_t)(row >> 32) << 8), key);
    const float target = (float)(r.x >> 8) * (1.0f / 16777216.0f) * s;
    int pick = 0;
    float run = 0.0f, lp = at(0);
    bool found = false;
    for (int k = 0; k < n_actions; ++k) {
        const float l = at(k), e = expf(l - m);
        run += e;
        if (!found && e > 0.0f) { pick = k; lp = l; }          // the last action with a non-zero term so far
        if (!found && run > target) found = true;
    }
    pick_out = pick;
    logp_out = lp - m - logf(s);
}

}  // namespace futbol
