// v0 kernels: reset, per-step, fused K-step rollout, AoS<->SoA state access.
//
// HBM layout of the state buffer (structure of arrays, `np` = n_envs rounded up to 256 so every
// section starts 256-byte aligned and every warp's accesses are fully coalesced):
//   double  f[25][np]     rows ai_1, ai_2, opp_1, opp_2, ball x (x, y, tx, ty, speed)
//   uint64  t_total[np]
//   int32   ep_step[np], ai_score[np], opp_score[np]
//   uint8   owner[np], last_owner[np], flags[np]
//   uint32  sched[64 + np / 32]   work queue of the time-sliced rollout (below): [0] next unit, [64 + w] chunks done by
//                                 the envs of block-group w; zeroed by the launcher before every such launch
// = 223 bytes per env.  Inside a kernel the 25 doubles of an environment sit in shared memory (one column
// per lane, v0_step.cuh) and the scalars in registers; HBM is touched at the two ends of a launch only.
#include <cuda_runtime.h>
#include <mutex>
#include <stdint.h>
#include "../../include/futbol_b200.h"
#include "v0_step.cuh"
#include "v0_kernels.h"

namespace futbol {

struct StateView {
    double *f; uint64_t *t_total; int32_t *ep_step, *ai_score, *opp_score; uint8_t *owner, *last_owner, *flags;
    uint32_t *sched;
    size_t np;
};
constexpr int kSchedHead = 64;                                        // words before the per-group counters (256 B)
__host__ __device__ inline size_t v0_sched_words(size_t np) { return kSchedHead + np / 32; }

__host__ __device__ inline size_t v0_padded(int n) { return ((size_t)n + 255) & ~(size_t)255; }

size_t v0_state_bytes(int n_envs) { return v0_padded(n_envs) * (25 * 8 + 8 + 3 * 4 + 3) + v0_sched_words(v0_padded(n_envs)) * 4; }

__host__ __device__ inline StateView make_view(void *base, int n)
{
    StateView v;
    v.np = v0_padded(n);
    char *p = (char *)base;
    v.f = (double *)p;            p += v.np * 25 * 8;
    v.t_total = (uint64_t *)p;    p += v.np * 8;
    v.ep_step = (int32_t *)p;     p += v.np * 4;
    v.ai_score = (int32_t *)p;    p += v.np * 4;
    v.opp_score = (int32_t *)p;   p += v.np * 4;
    v.owner = (uint8_t *)p;       p += v.np;
    v.last_owner = (uint8_t *)p;  p += v.np;
    v.flags = (uint8_t *)p;       p += v.np;
    v.sched = (uint32_t *)p;
    return v;
}

#ifndef FUTBOL_ENV_THREADS
#define FUTBOL_ENV_THREADS 128
#endif
constexpr int kEnvThreads = FUTBOL_ENV_THREADS;                   // threads per block of every env kernel
constexpr int kEnvSmemBytes = (kEnvThreads / 32) * kWarpSmemBytes;   // dynamic shared memory per block

// The loads go to L2 (ld.global.cg): in the time-sliced rollout another SM may have written this state a moment ago,
// and L1 is not coherent.
__device__ __forceinline__ void load_state(const StateView &v, int i, Lane L, V0Regs &s)
{
    const double *f = v.f + i;
#pragma unroll
    for (int k = 0; k < 25; ++k) L.f(k * kLanes) = __ldcg(f + (size_t)k * v.np);
    s.t_total = __ldcg(v.t_total + i);
    s.ep_step = __ldcg(v.ep_step + i); s.ai_score = __ldcg(v.ai_score + i); s.opp_score = __ldcg(v.opp_score + i);
    s.owner = __ldcg(v.owner + i); s.last_owner = __ldcg(v.last_owner + i);
}

__device__ __forceinline__ void store_state(const StateView &v, int i, Lane L, const V0Regs &s, int flags)
{
    double *f = v.f + i;
#pragma unroll
    for (int k = 0; k < 25; ++k) f[(size_t)k * v.np] = L.f(k * kLanes);
    v.t_total[i] = s.t_total;
    v.ep_step[i] = s.ep_step; v.ai_score[i] = s.ai_score; v.opp_score[i] = s.opp_score;
    v.owner[i] = (uint8_t)s.owner; v.last_owner[i] = (uint8_t)s.last_owner; v.flags[i] = (uint8_t)flags;
}

// ---- observation output ------------------------------------------------------------------------
// A thread's observation is 30 consecutive values, so lane-strided stores would touch 30 cache lines
// per instruction.  fp32 rows are therefore staged through shared memory per warp and written out as
// consecutive 128-bit stores (a warp's 32 rows are one contiguous 3840-byte span of [n, 30]).  The
// staging area reuses the space of the step's draw words (dead by now), hence the leading __syncwarp.
// Bulk asynchronous copy shared -> global (cp.async.bulk, SASS UBLKCP): one lane hands the warp's finished 3840-byte tile to
// the copy engine instead of 32 lanes moving it with 8 LDS.128 + 8 STG.128 each.  The staging area may be rewritten (by
// the next step's draw words) only after the engine has READ it: bulk_wait_read() before that.
// Measured (2^20 envs, K = 64, profiles/r2_v0_history.md): +1.4 % with random opponents (1.221 -> 1.238e10 env-steps/s),
// -6.7 % with the hard-coded opponents (1.056 -> 0.985e10; the larger kernel, instruction-fetch bound), so only the
// random-opponent kernels use it.
#ifndef FUTBOL_NO_BULK_STORE
#define FUTBOL_BULK_STORE 1
#else
#define FUTBOL_BULK_STORE 0
#endif
__device__ __forceinline__ void bulk_store_s2g(void *gdst, const void *ssrc, uint32_t bytes)
{
    const uint32_t src = (uint32_t)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(src), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// BULK: the tile leaves through the copy engine when it is a whole, 16-byte aligned warp tile (the rollout kernels, which
// call bulk_wait_read() before the staging area is reused); otherwise, and in the per-step kernel, through the lanes.
template <bool BULK = false>
__device__ __forceinline__ void warp_store_obs_f32(Lane L, const V0Regs &s, float *stage, float *gdst_warp_row0,
                                                   int lane, int rows_in_warp, bool vec_ok)
{
    __syncwarp();
    float *mine = stage + lane * kObsDim;
    // one row (5 values) per trip: loads, conversions and stores of a row overlap, and the loop is a quarter of the
    // unrolled code (instruction-cache footprint, r1_history.md r1i)
#pragma unroll 1
    for (int r = 0; r < 5; ++r) {
        double d[5];
#pragma unroll
        for (int c = 0; c < 5; ++c) d[c] = L.f((5 * r + c) * kLanes);
#pragma unroll
        for (int c = 0; c < 5; ++c) mine[5 * r + c] = (float)d[c];
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) mine[25 + k] = (float)obs_owner_elem(s, k);
    if (BULK && FUTBOL_BULK_STORE && vec_ok && rows_in_warp == 32) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // this lane's staged values -> visible to the async proxy
        __syncwarp();
        if (lane == 0) bulk_store_s2g(gdst_warp_row0, stage, 32 * kObsDim * 4);
        return;
    }
    __syncwarp();
    const int total = rows_in_warp * kObsDim;
    if (vec_ok) {
        const int nvec = total >> 2;
        const float4 *src = reinterpret_cast<const float4 *>(stage);
        float4 *dst = reinterpret_cast<float4 *>(gdst_warp_row0);
        for (int q = lane; q < nvec; q += 32) __stcs(dst + q, src[q]);
        for (int q = (nvec << 2) + lane; q < total; q += 32) __stcs(gdst_warp_row0 + q, stage[q]);
    } else {
        for (int q = lane; q < total; q += 32) __stcs(gdst_warp_row0 + q, stage[q]);
    }
    __syncwarp();
}

// The same tile in THREE passes of ten values per environment through the 1536-byte area of the draw words (dense
// rollout kernels: 7936 B of shared memory per warp = seven blocks = 28 warps per SM, which holds a 131,072-env rank
// batch -- 4096 warps on 148 x 28 = 4144 slots -- in ONE wave).  A pass stages [32 envs][10 floats] = 160 float2; float2 u of
// the area belongs to env u / 5 and goes to float2 15 (u / 5) + 5 pass + u % 5 of the warp's tile: 40-byte runs, 8-byte stores.
__device__ __forceinline__ void warp_store_obs_f32_3pass(Lane L, const V0Regs &s, float *stage, float *gdst_warp_row0,
                                                         int lane, int rows_in_warp, bool vec_ok)
{
    float2 *mine = reinterpret_cast<float2 *>(stage) + lane * 5;
#pragma unroll 1
    for (int p = 0; p < 3; ++p) {
        __syncwarp();                                  // pass 0: the draws are dead; later: the previous pass has been read
        if (p < 2) {
#pragma unroll
            for (int c = 0; c < 5; ++c) mine[c] = make_float2((float)L.f((10 * p + 2 * c) * kLanes), (float)L.f((10 * p + 2 * c + 1) * kLanes));
        } else {
#pragma unroll
            for (int c = 0; c < 2; ++c) mine[c] = make_float2((float)L.f((20 + 2 * c) * kLanes), (float)L.f((21 + 2 * c) * kLanes));
            mine[2] = make_float2((float)L.f(24 * kLanes), (float)obs_owner_elem(s, 0));
            mine[3] = make_float2((float)obs_owner_elem(s, 1), (float)obs_owner_elem(s, 2));
            mine[4] = make_float2((float)obs_owner_elem(s, 3), (float)obs_owner_elem(s, 4));
        }
        __syncwarp();
        const float2 *src = reinterpret_cast<const float2 *>(stage);
        if (vec_ok) {
            float2 *dst = reinterpret_cast<float2 *>(gdst_warp_row0);
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const int u = lane + 32 * i, e = (u * 205) >> 10;            // u / 5 for u < 160
                if (e < rows_in_warp) __stcs(dst + 15 * e + 5 * p + (u - 5 * e), src[u]);
            }
        } else {
            for (int q = lane; q < 320; q += 32) {
                const int e = q / 10;
                if (e < rows_in_warp) __stcs(gdst_warp_row0 + 30 * e + 10 * p + (q - 10 * e), stage[q]);
            }
        }
    }
    __syncwarp();
}

template <typename T>
__device__ __forceinline__ void thread_store_obs(T *dst_row, Lane L, const V0Regs &s)
{
#pragma unroll
    for (int k = 0; k < 25; ++k) dst_row[k] = (T)L.f(k * kLanes);
#pragma unroll
    for (int k = 0; k < 5; ++k) dst_row[25 + k] = (T)obs_owner_elem(s, k);
}

// ---- reset ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kEnvThreads) v0_reset_kernel(V0Params P, StateView v, const uint8_t *mask, T *obs, int init)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_envs) return;
    if (mask != nullptr && mask[i] == 0) return;
    const Lane L = make_lane(threadIdx.x >> 5, threadIdx.x & 31);
    V0Regs s;
    s.t_total = init ? 0 : v.t_total[i];
    reset_env(L, s);
    store_state(v, i, L, s, 0);
    if (obs != nullptr) thread_store_obs(obs + (size_t)i * kObsDim, L, s);
}

#ifndef FUTBOL_MIN_BLOCKS
#define FUTBOL_MIN_BLOCKS 5     // resident blocks per SM the register allocation is sized for (shared memory allows 5)
#endif

// ---- per-step API ----------------------------------------------------------------------------------
template <typename T, bool RANDOM_OPP>
__global__ void __launch_bounds__(kEnvThreads, FUTBOL_MIN_BLOCKS) v0_step_kernel(V0Params P, StateView v, const uint8_t *actions,
                                                              const uint8_t *opp_actions, T *obs, T *reward, uint8_t *done,
                                                              T *final_obs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warp_env0 = i - lane;
    if (warp_env0 >= P.n_envs) return;            // whole warp out of range
    const bool live = i < P.n_envs;
    const int rows_in_warp = min(32, P.n_envs - warp_env0);
    const Lane L = make_lane(warp, lane);
    V0Regs s;
    StepResult r;
    r.reward = 0.0; r.done = 0; r.flags = 0;
    if (live) {
        load_state(v, i, L, s);
        r = v0_step<RANDOM_OPP>(L, s, P, P.env_id_offset + (uint32_t)i, actions[i] & 15,
                                opp_actions != nullptr ? (int)(opp_actions[i] & 15) : -1);
        if (r.done && P.auto_reset) {
            if (final_obs != nullptr) thread_store_obs(final_obs + (size_t)i * kObsDim, L, s);
            reset_env(L, s);
        }
        store_state(v, i, L, s, r.flags);
        if (reward != nullptr) reward[i] = (T)r.reward;
        if (done != nullptr) done[i] = (uint8_t)r.done;
    } else {
        s.t_total = 0; reset_env(L, s);           // padding lanes of the last warp: a private dummy env for the staged store
    }
    if (obs != nullptr) {
        if (sizeof(T) == 4) {                     // fp32: staged through shared memory, 128-bit coalesced stores
            float *stage = reinterpret_cast<float *>(futbol_smem + warp * kWarpSmemBytes + kWarpStateBytes);
            float *dst = reinterpret_cast<float *>(obs) + (size_t)warp_env0 * kObsDim;
            const bool vec_ok = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
            warp_store_obs_f32(L, s, stage, dst, lane, rows_in_warp, vec_ok);
        } else if (live) {
            thread_store_obs(obs + (size_t)i * kObsDim, L, s);
        }
    }
}

// ---- fused K-step rollout ----------------------------------------------------------------------------

// Steps [k0, k1) of the rollout for the 128 envs of block-group `group`: state HBM -> shared memory, the steps,
// state back.  One call per block in the plain rollout; one call per work unit in the time-sliced one.
template <bool RANDOM_OPP, bool DENSE = false, bool SLICED = false>
__device__ __forceinline__ void rollout_span(const V0Params &P, const StateView &v, int group, int k0, int k1,
                                             const uint8_t *__restrict__ actions, const uint8_t *__restrict__ opp_actions,
                                             float *__restrict__ obs, float *__restrict__ reward, uint8_t *__restrict__ done,
                                             FutbolStats *stats)
{
    const int i = group * kEnvThreads + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warp_env0 = i - lane;
    if (warp_env0 >= P.n_envs) return;            // whole warp out of range
    const bool live = i < P.n_envs;
    const int rows_in_warp = min(32, P.n_envs - warp_env0);
    const size_t n = (size_t)P.n_envs;
    // 128-bit stores need every step's row block 16-byte aligned: n*30*4 % 16 == 0  <=>  n even
    const bool vec_ok = ((n & 1) == 0) && ((reinterpret_cast<uintptr_t>(obs) & 15) == 0);   // (8-byte stores: always true then)
    const uint32_t env_id = P.env_id_offset + (uint32_t)i;
    constexpr int kWB = DENSE ? kWarpSmemBytesDense : kWarpSmemBytes;
    const Lane L = make_lane(warp, lane, kWB);
    float *stage = reinterpret_cast<float *>(futbol_smem + warp * kWB + kWarpStateBytes);
    constexpr bool kBulk = FUTBOL_BULK_STORE && RANDOM_OPP && !DENSE;

    V0Regs s;
    if (live) load_state(v, i, L, s);
    else { s.t_total = 0; reset_env(L, s); }      // padding lanes of the last warp step a private dummy env

    double reward_sum = 0.0;
    uint32_t episodes = 0, goals_ai = 0, goals_opp = 0, fixes = 0;
    int last_flags = 0;

    int a_next = (actions != nullptr && live && k0 < k1) ? (int)__ldg(actions + (size_t)k0 * n + (size_t)i) : 0;
#pragma unroll 1
    for (int k = k0; k < k1; ++k) {
        const size_t slot = (size_t)k * n + (size_t)i;
        int a;
        if (actions != nullptr) {                 // this step's byte was requested one step ago: no wait on HBM here
            a = a_next & 15;
            if (live && k + 1 < k1) a_next = __ldg(actions + slot + n);
        } else a = philox_action(P.key, env_id, s.t_total, 16);
        const int ai_before = s.ai_score;
        const int oa = (opp_actions != nullptr && live) ? (int)(__ldg(opp_actions + slot) & 15) : -1;
        // the draw words are parked where the previous step's observation tile was staged: with the bulk store, wait
        // until the copy engine has read that tile
        const bool wait_tile = kBulk && obs != nullptr && k > k0;
        auto before_draws = [&]() {
            if (kBulk) {
                if (wait_tile) { if (lane == 0) bulk_wait_read(); __syncwarp(); }
            }
        };
#ifndef FUTBOL_SLICED_LAYOUT
#define FUTBOL_SLICED_LAYOUT 8
#endif
#ifndef FUTBOL_PLAIN_LAYOUT
#define FUTBOL_PLAIN_LAYOUT 0
#endif
        const StepResult r = v0_step<RANDOM_OPP, decltype(before_draws), SLICED ? FUTBOL_SLICED_LAYOUT : FUTBOL_PLAIN_LAYOUT>(L, s, P, env_id, a, oa, before_draws);
        last_flags = r.flags;
        reward_sum += r.reward;
        goals_ai += (r.flags & kFlagGoal) && s.ai_score != ai_before;
        goals_opp += (r.flags & kFlagGoal) && s.ai_score == ai_before;
        fixes += (r.flags & kFlagFix) != 0;
        episodes += r.done;
        if (r.done && P.auto_reset) reset_env(L, s);
        if (obs != nullptr) {
            float *tile = obs + ((size_t)k * n + (size_t)warp_env0) * kObsDim;
            if (DENSE) warp_store_obs_f32_3pass(L, s, stage, tile, lane, rows_in_warp, (reinterpret_cast<uintptr_t>(tile) & 7) == 0);
            else warp_store_obs_f32<kBulk>(L, s, stage, tile, lane, rows_in_warp, vec_ok);
        }
        if (live) {
            if (reward != nullptr) __stcs(reward + slot, (float)r.reward);
            if (done != nullptr) done[slot] = (uint8_t)r.done;
        }
    }
    if (live) store_state(v, i, L, s, last_flags);
    if (kBulk && obs != nullptr && lane == 0) bulk_wait_all();                  // every tile of this span has landed

    if (stats != nullptr) {
        if (!live) { reward_sum = 0.0; episodes = goals_ai = goals_opp = fixes = 0; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            reward_sum += __shfl_xor_sync(0xffffffffu, reward_sum, o);
            episodes += __shfl_xor_sync(0xffffffffu, episodes, o);
            goals_ai += __shfl_xor_sync(0xffffffffu, goals_ai, o);
            goals_opp += __shfl_xor_sync(0xffffffffu, goals_opp, o);
            fixes += __shfl_xor_sync(0xffffffffu, fixes, o);
        }
        if (lane == 0) {
            atomicAdd(&stats->reward_sum, reward_sum);
            atomicAdd((unsigned long long *)&stats->env_steps, (unsigned long long)rows_in_warp * (unsigned long long)(k1 - k0));
            atomicAdd((unsigned long long *)&stats->episodes, (unsigned long long)episodes);
            atomicAdd((unsigned long long *)&stats->goals_ai, (unsigned long long)goals_ai);
            atomicAdd((unsigned long long *)&stats->goals_opp, (unsigned long long)goals_opp);
            atomicAdd((unsigned long long *)&stats->out_of_field, (unsigned long long)fixes);
        }
    }
}


template <bool RANDOM_OPP>
__global__ void __launch_bounds__(kEnvThreads, FUTBOL_MIN_BLOCKS)
v0_rollout_kernel(V0Params P, StateView v, int K, const uint8_t *__restrict__ actions, const uint8_t *__restrict__ opp_actions,
                  float *__restrict__ obs, float *__restrict__ reward, uint8_t *__restrict__ done, FutbolStats *stats)
{
    rollout_span<RANDOM_OPP>(P, v, blockIdx.x, 0, K, actions, opp_actions, obs, reward, done, stats);
}

// The dense variant: 7936 B of shared memory per warp and 72 registers -> seven blocks = 28 warps per SM (4144 warp slots on a
// B200): the rank-sized batches of the 2^20 job (131,072 / 262,144 / 524,288 envs = 4096 / 8192 / 16384 warps) are then one, two
// and four nearly full waves, where 20 warps per SM leave 1.38 / 2.77 / 5.5.
constexpr int kEnvSmemBytesDense = (kEnvThreads / 32) * kWarpSmemBytesDense;
template <bool RANDOM_OPP>
__global__ void __launch_bounds__(kEnvThreads, 7)
v0_rollout_dense_kernel(V0Params P, StateView v, int K, const uint8_t *__restrict__ actions, const uint8_t *__restrict__ opp_actions,
                        float *__restrict__ obs, float *__restrict__ reward, uint8_t *__restrict__ done, FutbolStats *stats)
{
    rollout_span<RANDOM_OPP, true>(P, v, blockIdx.x, 0, K, actions, opp_actions, obs, reward, done, stats);
}

// Time-sliced rollout for batches of only a few waves of blocks (131,072 envs per GPU = 1024 blocks on 740 block slots:
// the last 0.38 wave would run on a mostly idle GPU).  The K steps are cut into `chunks` slices of `chunk_steps`; a work
// unit is (slice c, block-group g), numbered c * groups + g, and a grid that just fills the GPU takes units from a
// counter.  Unit (c, g) needs (c - 1, g): that unit has a SMALLER number, so it was taken earlier by a block that is
// running and never waits on a later unit -- no deadlock whatever the residency.  With groups >= grid it has normally
// finished long ago; otherwise thread 0 polls the group's counter.  Hand-over of the state between blocks goes through
// HBM/L2: writer __threadfence + barrier + counter store, reader counter load + __threadfence + barrier + ld.cg loads.
template <bool RANDOM_OPP>
__global__ void __launch_bounds__(kEnvThreads, FUTBOL_MIN_BLOCKS)
v0_rollout_sliced_kernel(V0Params P, StateView v, int K, int chunk_steps, int chunks, int groups,
                         const uint8_t *__restrict__ actions, const uint8_t *__restrict__ opp_actions,
                         float *__restrict__ obs, float *__restrict__ reward, uint8_t *__restrict__ done, FutbolStats *stats)
{
    __shared__ uint32_t unit_s;
    volatile uint32_t *progress = v.sched + kSchedHead;
    const uint32_t units = (uint32_t)chunks * (uint32_t)groups;
    for (;;) {
        if (threadIdx.x == 0) unit_s = atomicAdd(v.sched, 1u);
        __syncthreads();
        const uint32_t u = unit_s;
        if (u >= units) break;
        const int c = (int)(u / (uint32_t)groups), g = (int)(u % (uint32_t)groups);
        if (c > 0) {
            if (threadIdx.x == 0) {
                while (progress[g] < (uint32_t)c) __nanosleep(200);
                __threadfence();
            }
        }
        __syncthreads();                       // also keeps unit_s from being overwritten while others still read it
        rollout_span<RANDOM_OPP, false, true>(P, v, g, c * chunk_steps, min(K, (c + 1) * chunk_steps), actions, opp_actions, obs,
                                              reward, done, stats);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) progress[g] = (uint32_t)(c + 1);
    }
}

// ---- AoS <-> SoA -------------------------------------------------------------------------------------
__global__ void v0_get_state_kernel(int n, StateView v, FutbolV0EnvState *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    FutbolV0EnvState e;
    for (int k = 0; k < 25; ++k) e.rows[k / 5][k % 5] = v.f[(size_t)k * v.np + i];
    e.t_total = v.t_total[i]; e.ep_step = v.ep_step[i]; e.ai_score = v.ai_score[i]; e.opp_score = v.opp_score[i];
    e.owner = v.owner[i]; e.last_owner = v.last_owner[i]; e.flags = v.flags[i]; e.pad_ = 0;
    out[i] = e;
}

__global__ void v0_set_state_kernel(int n, StateView v, const FutbolV0EnvState *in)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const FutbolV0EnvState e = in[i];
    for (int k = 0; k < 25; ++k) v.f[(size_t)k * v.np + i] = e.rows[k / 5][k % 5];
    v.t_total[i] = e.t_total; v.ep_step[i] = e.ep_step; v.ai_score[i] = e.ai_score; v.opp_score[i] = e.opp_score;
    v.owner[i] = e.owner; v.last_owner[i] = e.last_owner; v.flags[i] = e.flags;
}

// ---- host launchers ------------------------------------------------------------------------------------
static inline int blocks_for(int n, int t) { return (n + t - 1) / t; }

cudaError_t v0_launch_reset(const V0Params &P, void *state, const uint8_t *mask, void *obs, int obs_f64, int init,
                            cudaStream_t st)
{
    const StateView v = make_view(state, P.n_envs);
    const int blocks = blocks_for(P.n_envs, kEnvThreads);
    if (obs_f64) v0_reset_kernel<double><<<blocks, kEnvThreads, kEnvSmemBytes, st>>>(P, v, mask, (double *)obs, init);
    else v0_reset_kernel<float><<<blocks, kEnvThreads, kEnvSmemBytes, st>>>(P, v, mask, (float *)obs, init);
    return cudaGetLastError();
}

template <typename T, bool RANDOM_OPP>
static void launch_step(const V0Params &P, const StateView &v, const uint8_t *actions, const uint8_t *opp_actions, void *obs,
                        void *reward, uint8_t *done, void *final_obs, cudaStream_t st)
{
    v0_step_kernel<T, RANDOM_OPP><<<blocks_for(P.n_envs, kEnvThreads), kEnvThreads, kEnvSmemBytes, st>>>(
        P, v, actions, opp_actions, (T *)obs, (T *)reward, done, (T *)final_obs);
}

cudaError_t v0_launch_step(const V0Params &P, void *state, const uint8_t *actions, const uint8_t *opp_actions, void *obs,
                           void *reward, uint8_t *done, void *final_obs, int out_f64, cudaStream_t st)
{
    const StateView v = make_view(state, P.n_envs);
    if (out_f64) {
        if (P.random_opp) launch_step<double, true>(P, v, actions, opp_actions, obs, reward, done, final_obs, st);
        else launch_step<double, false>(P, v, actions, opp_actions, obs, reward, done, final_obs, st);
    } else {
        if (P.random_opp) launch_step<float, true>(P, v, actions, opp_actions, obs, reward, done, final_obs, st);
        else launch_step<float, false>(P, v, actions, opp_actions, obs, reward, done, final_obs, st);
    }
    return cudaGetLastError();
}

// resident blocks of the rollout kernels on the CURRENT device (SMs x blocks per SM), queried once per device and variant
// (a process may drive several GPUs, or create environments from several threads)
template <bool RANDOM_OPP>
static int rollout_block_slots()
{
    static std::mutex mu;
    static int slots[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    int &s = slots[dev & 63];
    if (s == 0) {
        int sms = 0, per_sm = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, v0_rollout_sliced_kernel<RANDOM_OPP>, kEnvThreads, kEnvSmemBytes);
        s = sms * per_sm > 0 ? sms * per_sm : 1;
    }
    return s;
}

// How a K-step rollout of this batch is launched (profiles/r2_slices.md has the measurements behind the rule).
//   slices: 0 = automatic: the time-sliced work-queue launch (four slices at K = 64) when the batch is between one and two
//   waves of blocks -- 131,072 envs, one of eight ranks of the 2^20 job: +8 % over the plain launch -- and the plain
//   launch otherwise: from about 2.8 waves on the plain launch wins (262,144 envs: 1.677 vs 1.738 ms), because blocks of a
//   thinning last wave run faster than blocks of a full one, while every slice pays the hand-over.  1 = never slice,
//   n > 1 = n equal slices (futbol_set_rollout_slices: tests, tuning).
//   variant: 0 / 1 = the standard kernel, 2 = the dense kernel (28 warps per SM).  Never chosen automatically: its
//   three-pass observation staging and 72-register budget cost more than its whole waves gain (2^20 envs: 6.91 vs 6.37 ms;
//   131,072: 0.942 vs 0.889 ms sliced); kept selectable so that the measurement can be repeated.
V0RolloutChoice v0_plan_rollout(const V0Params &P, int K, int slices, int variant)
{
    V0RolloutChoice c;
    c.kernel = variant == 2 ? 1 : 0;
    c.slices = 1;
    if (c.kernel == 1) return c;
    const int groups = blocks_for(P.n_envs, kEnvThreads);
    const int slots = P.random_opp ? rollout_block_slots<true>() : rollout_block_slots<false>();
    int n = 1;
    if (slices > 0) n = slices < K ? slices : K;
    else if (groups > slots) {      // measured per wave count, profiles/r2_slices.md ("Slices per wave count")
        const long long g = groups, sl = slots;
        if (g < 2 * sl) n = (int)((11 * sl + 2 * g - 1) / (2 * g));     // one to two waves: five to six waves of units
        else if (g <= 6 * sl) n = 3;
        else if (g <= 10 * sl) n = 2;
        if (n > K / 4) n = K / 4 > 0 ? K / 4 : 1;
    }
    if (n > 1) {
        const int chunk_steps = (K + n - 1) / n;
        c.slices = (K + chunk_steps - 1) / chunk_steps;
    }
    return c;
}

template <bool RANDOM_OPP>
static cudaError_t launch_rollout(const V0Params &P, const StateView &v, int K, const uint8_t *actions, const uint8_t *opp_actions,
                                  float *obs, float *reward, uint8_t *done, FutbolStats *stats, int slices, int variant, cudaStream_t st)
{
    const int groups = blocks_for(P.n_envs, kEnvThreads);
    const V0RolloutChoice c = v0_plan_rollout(P, K, slices, variant);
    if (c.kernel == 1) {
        v0_rollout_dense_kernel<RANDOM_OPP><<<groups, kEnvThreads, kEnvSmemBytesDense, st>>>(P, v, K, actions, opp_actions, obs, reward, done, stats);
        return cudaGetLastError();
    }
    if (c.slices == 1) {
        v0_rollout_kernel<RANDOM_OPP><<<groups, kEnvThreads, kEnvSmemBytes, st>>>(P, v, K, actions, opp_actions, obs, reward, done, stats);
        return cudaGetLastError();
    }
    const int slots = rollout_block_slots<RANDOM_OPP>();
    const int chunk_steps = (K + c.slices - 1) / c.slices;
    cudaError_t e = cudaMemsetAsync(v.sched, 0, v0_sched_words(v.np) * 4, st);
    if (e != cudaSuccess) return e;
    const long long units = (long long)c.slices * groups;
    const int grid = (int)(units < slots ? units : slots);
    v0_rollout_sliced_kernel<RANDOM_OPP><<<grid, kEnvThreads, kEnvSmemBytes, st>>>(P, v, K, chunk_steps, c.slices, groups, actions,
                                                                                 opp_actions, obs, reward, done, stats);
    return cudaGetLastError();
}

cudaError_t v0_launch_rollout(const V0Params &P, void *state, int K, const uint8_t *actions, const uint8_t *opp_actions,
                              float *obs, float *reward, uint8_t *done, FutbolStats *stats, int slices, int variant, cudaStream_t st)
{
    const StateView v = make_view(state, P.n_envs);
    return P.random_opp ? launch_rollout<true>(P, v, K, actions, opp_actions, obs, reward, done, stats, slices, variant, st)
                        : launch_rollout<false>(P, v, K, actions, opp_actions, obs, reward, done, stats, slices, variant, st);
}

cudaError_t v0_launch_get_state(int n, const void *state, void *aos, cudaStream_t st)
{
    v0_get_state_kernel<<<blocks_for(n, 128), 128, 0, st>>>(n, make_view(const_cast<void *>(state), n), (FutbolV0EnvState *)aos);
    return cudaGetLastError();
}

cudaError_t v0_launch_set_state(int n, void *state, const void *aos, cudaStream_t st)
{
    v0_set_state_kernel<<<blocks_for(n, 128), 128, 0, st>>>(n, make_view(state, n), (const FutbolV0EnvState *)aos);
    return cudaGetLastError();
}

}  // namespace futbol
