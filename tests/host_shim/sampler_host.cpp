// TEST-ONLY: compiles the row rule of futbol_sample_actions (gym_futbol_b200/csrc/sampler.cuh) with g++ so that it can be
// checked without a GPU (tests/test_sampler_host.py).  Nothing in the package loads this.
#include <cmath>
#include <cstdint>
#define __device__
#define __host__
#define __forceinline__ inline
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }
#include "../../gym_futbol_b200/csrc/sampler.cuh"

extern "C" void host_sample_actions(const float *logits, long long n, int n_actions, uint64_t seed, uint64_t t, uint8_t *actions,
                                    float *logp)
{
    const futbol::PhiloxKey key = futbol::philox_expand_key(seed);
    for (long long i = 0; i < n; ++i) {
        const float *row = logits + i * n_actions;
        int pick;
        float lp;
        futbol::sample_row([row](int k) { return row[k]; }, n_actions, key, t, (unsigned long long)i, pick, lp);
        actions[i] = (uint8_t)pick;
        logp[i] = lp;
    }
}
