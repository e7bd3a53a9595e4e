// Philox4x32-10, per-step counter layout.  Specification (shared with the CPU oracle, which
// implements it independently): key = (seed_lo, seed_hi); counter = (t_lo, (t_hi & 0xffff) |
// (block << 16), global_env_id, stream); draw j of a step = word (j & 3) of block (j >> 2).
// A step's randomness depends only on (seed, env id, t): no RNG state lives in HBM.
#pragma once
#include <stdint.h>

namespace futbol {

constexpr uint32_t kStreamDynamics = 0;   // v0 environment draws
constexpr uint32_t kStreamActions = 1;    // synthetic AI actions (bench / in-kernel random policy)
constexpr uint32_t kStreamV1Opp = 2;      // v1 opponent actions
constexpr uint32_t kStreamV1Dynamics = 3; // v1 environment draws
constexpr uint32_t kNormalBlock0 = 0x8000u; // first Philox block of a step's normal() slots

struct Philox4 { uint32_t x, y, z, w; };

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c2 = hi0 ^ c3 ^ k1;
        c1 = lo1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

__device__ __forceinline__ Philox4 philox_step_block(uint64_t seed, uint32_t env_id, uint32_t stream,
                                                     uint64_t t, uint32_t block)
{
    return philox4x32_10((uint32_t)t, ((uint32_t)(t >> 32) & 0xFFFFu) | (block << 16), env_id, stream,
                         (uint32_t)seed, (uint32_t)(seed >> 32));
}

// The sequential per-step draw stream.  The first kPre*4 words are produced up front by every
// lane (uniform control flow); later words (only the rare "shoot" path reaches them) on demand.
template <int kPreBlocks>
struct StepRng {
    uint32_t w[kPreBlocks * 4];
    uint64_t seed, t;
    uint32_t env_id, stream, j, normal_calls;

    __device__ __forceinline__ void begin(uint64_t seed_, uint32_t env_id_, uint32_t stream_, uint64_t t_)
    {
        seed = seed_; env_id = env_id_; stream = stream_; t = t_; j = 0; normal_calls = 0;
#pragma unroll
        for (int b = 0; b < kPreBlocks; ++b) {
            const Philox4 p = philox_step_block(seed, env_id, stream, t, b);
            w[4 * b + 0] = p.x; w[4 * b + 1] = p.y; w[4 * b + 2] = p.z; w[4 * b + 3] = p.w;
        }
    }

    __device__ __forceinline__ uint32_t word_at(uint32_t idx) const
    {
        if (idx < (uint32_t)(kPreBlocks * 4)) {
            // register select tree (a dynamically indexed array would live in local memory)
            uint32_t r = w[0];
#pragma unroll
            for (int i = 1; i < kPreBlocks * 4; ++i) r = (idx == (uint32_t)i) ? w[i] : r;
            return r;
        }
        const Philox4 p = philox_step_block(seed, env_id, stream, t, idx >> 2);
        const uint32_t lo = (idx & 1u) ? p.y : p.x, hi = (idx & 1u) ? p.w : p.z;
        return (idx & 2u) ? hi : lo;
    }

    __device__ __forceinline__ uint32_t next_u32() { return word_at(j++); }
    // random.random(): (w >> 8) * 2^-24
    __device__ __forceinline__ double random() { return (double)(next_u32() >> 8) * (1.0 / 16777216.0); }
    // random.randint(a, b): a + ((w * (b - a + 1)) >> 32)
    __device__ __forceinline__ int randint(int a, int b) { return a + (int)__umulhi(next_u32(), (uint32_t)(b - a + 1)); }
};

__device__ __forceinline__ int philox_action(uint64_t seed, uint32_t env_id, uint64_t t, uint32_t n_actions)
{
    return (int)__umulhi(philox_step_block(seed, env_id, kStreamActions, t, 0).x, n_actions);
}

}  // namespace futbol
