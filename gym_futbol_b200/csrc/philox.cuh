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

// Per-environment working storage lives in shared memory as one column per lane of the owning warp:
// element k of lane l at base[k * kLanes + l], so a warp-wide access is conflict-free whatever k is.
#ifndef FUTBOL_LANES
#define FUTBOL_LANES 32
#endif
constexpr int kLanes = FUTBOL_LANES;
constexpr int kSeqBlocks = 3;             // sequential draws generated per step: 12 words (a step uses <= 10)
constexpr int kDrawWords = 4 * kSeqBlocks;

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

// The sequential draws of one step.  The reference consumes a state-dependent number of draws in a
// state-dependent order, so the position of the next draw is data.  All kDrawWords words a step can need
// are generated up front and parked in this lane's shared-memory column; a draw is then one LDS at a
// data-dependent row.  The blocks are unrolled side by side: ten rounds are a serial chain of multiply -> xor,
// and three independent chains interleave where one would leave the warp waiting.  (A loop over the rounds with the three
// blocks side by side is 70 instructions shorter and 1.3-3.4 % slower at 2^20 envs: profiles/r2_v0_history.md, r2g.)
__device__ __forceinline__ void philox_fill_block(uint32_t *col, const PhiloxKey &K, uint32_t env_id, uint32_t stream, uint64_t t, int b)
{
    const Philox4 p = philox_step_block(K, env_id, stream, t, (uint32_t)b);
    uint32_t *dst = col + 4 * b * kLanes;
    dst[0] = p.x; dst[kLanes] = p.y; dst[2 * kLanes] = p.z; dst[3 * kLanes] = p.w;
}

// blocks [0, n_blocks) of the step's sequential draws
__device__ __forceinline__ void philox_fill_step(uint32_t *col, const PhiloxKey &K, uint32_t env_id, uint32_t stream, uint64_t t,
                                                 int n_blocks = kSeqBlocks)
{
#pragma unroll
    for (int b = 0; b < kSeqBlocks; ++b)
        if (b < n_blocks) philox_fill_block(col, K, env_id, stream, t, b);
}

__device__ __forceinline__ int philox_action(const PhiloxKey &K, uint32_t env_id, uint64_t t, uint32_t n_actions)
{
    return (int)__umulhi(philox_step_block(K, env_id, kStreamActions, t, 0).x, n_actions);
}

}  // namespace futbol
