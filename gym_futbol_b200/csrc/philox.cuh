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

// Threads per block of every kernel that steps environments = stride of the shared-memory draw buffer.
#ifndef FUTBOL_ENV_THREADS
#define FUTBOL_ENV_THREADS 128
#endif
constexpr int kEnvThreads = FUTBOL_ENV_THREADS;
constexpr int kPreDraws = 8;              // draws generated up front per step (2 Philox blocks)

struct Philox4 { uint32_t x, y, z, w; };

// The ten round keys of a seed (k0 + r * 0x9E3779B9, k1 + r * 0xBB67AE85).  The seed is the same for every
// environment, so the host expands it once and the keys reach the kernel as launch constants: a round is
// two 32x32->64 multiplies and two three-input XORs.
struct PhiloxKey { uint32_t k0[10], k1[10]; };

__host__ __device__ inline PhiloxKey philox_expand_key(uint64_t seed)
{
    PhiloxKey K;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) { K.k0[r] = k0; K.k1[r] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    return K;
}

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKey &K)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ K.k0[r];
        c2 = hi0 ^ c3 ^ K.k1[r];
        c1 = lo1;
        c3 = lo0;
    }
    return Philox4{c0, c1, c2, c3};
}

__device__ __forceinline__ Philox4 philox_step_block(const PhiloxKey &K, uint32_t env_id, uint32_t stream,
                                                     uint64_t t, uint32_t block)
{
    return philox4x32_10((uint32_t)t, ((uint32_t)(t >> 32) & 0xFFFFu) | (block << 16), env_id, stream, K);
}

// draws past the pre-generated ones: out of line, a step needs at most 10 and almost always <= 8
static __device__ __noinline__ uint32_t philox_step_word(uint64_t seed, uint32_t env_id, uint32_t stream, uint64_t t,
                                                         uint32_t idx)
{
    const PhiloxKey K = philox_expand_key(seed);
    const Philox4 p = philox_step_block(K, env_id, stream, t, idx >> 2);
    const uint32_t lo = (idx & 1u) ? p.y : p.x, hi = (idx & 1u) ? p.w : p.z;
    return (idx & 2u) ? hi : lo;
}

// The sequential per-step draw stream.  The reference consumes a state-dependent number of draws in
// a state-dependent order, so the position `j` of the next draw is data; a register array indexed by
// data would cost a select tree per draw (or live in local memory).  The first kPreDraws words of the
// step are therefore parked in shared memory, one column per thread (word k of thread `tid` at
// buf[k * kEnvThreads + tid]: every lane hits its own bank whatever its j), and a draw is one LDS.
struct StepRng {
    const uint32_t *col;
    const PhiloxKey *key;
    uint64_t t;
    uint32_t env_id, stream, j;

    __device__ __forceinline__ uint64_t seed() const { return (uint64_t)key->k0[0] | ((uint64_t)key->k1[0] << 32); }

    __device__ __forceinline__ void begin(uint32_t *col_, const PhiloxKey &key_, uint32_t env_id_, uint32_t stream_, uint64_t t_)
    {
        col = col_; key = &key_; env_id = env_id_; stream = stream_; t = t_; j = 0;
#pragma unroll
        for (int b = 0; b < kPreDraws / 4; ++b) {
            const Philox4 p = philox_step_block(*key, env_id, stream, t, b);
            col_[(4 * b + 0) * kEnvThreads] = p.x; col_[(4 * b + 1) * kEnvThreads] = p.y;
            col_[(4 * b + 2) * kEnvThreads] = p.z; col_[(4 * b + 3) * kEnvThreads] = p.w;
        }
    }

    __device__ __forceinline__ uint32_t word_at(uint32_t idx) const
    {
        if (idx < (uint32_t)kPreDraws) return col[idx * kEnvThreads];
        return philox_step_word(seed(), env_id, stream, t, idx);
    }

    // CHECKED = false: the caller guarantees that the cursor is below kPreDraws at this site (v0: every
    // site before ai_2's turn, see the draw budget in v0_step.cuh), so the draw is a bare LDS.
    template <bool CHECKED>
    __device__ __forceinline__ uint32_t take()
    {
        const uint32_t w = CHECKED ? word_at(j) : col[j * kEnvThreads];
        j += 1;
        return w;
    }
    // the word at the cursor; consumed only if `c` (the value is ignored by the caller otherwise)
    template <bool CHECKED>
    __device__ __forceinline__ uint32_t take_if(bool c)
    {
        uint32_t w = 0;
        if (!CHECKED || j < (uint32_t)kPreDraws) w = col[j * kEnvThreads];
        else if (c) w = philox_step_word(seed(), env_id, stream, t, j);
        j += c ? 1u : 0u;
        return w;
    }
};

typedef StepRng V0Rng;

__device__ __forceinline__ int philox_action(const PhiloxKey &K, uint32_t env_id, uint64_t t, uint32_t n_actions)
{
    return (int)__umulhi(philox_step_block(K, env_id, kStreamActions, t, 0).x, n_actions);
}

}  // namespace futbol
