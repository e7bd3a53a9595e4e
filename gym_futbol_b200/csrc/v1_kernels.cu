// v1 kernels (N-vs-N rigid-body variant): reset, per-step, fused K-step rollout, state export.
//
// HBM layout of the state buffer (structure of arrays over environments, `np` = n_envs rounded up to 256),
// B = 2N + 1 bodies, P = B(B-1)/2 + 12 B shape pairs:
//   double  body[6 B][np]   per body x, y, vx, vy, v_bias_x, v_bias_y (team A, team B, ball)
//   uint64  t_total[np];  uint32 stamp[np];  int32 ep_step[np];  uint8 owner_side[np], flags[np]
//   CacheRec cache[P][np]                     the arbiter cache, 16 bytes per shape pair: accumulated normal impulse and
//                                             the stamp of the space step in which the pair last touched
//   uint32  sched[64 + np / 32]               work queue of the time-sliced rollout (zeroed by the launcher)
// Inside a kernel the 6 B doubles sit in shared memory (one column per lane), scalars in registers; the
// arbiter cache stays in HBM and is touched only by pairs in contact.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <mutex>
#include <stdint.h>
#include "../../include/futbol_b200.h"
#include "v1_step.cuh"
#include "v1_kernels.h"

namespace futbol {
namespace v1 {

struct StateView {
    double *body; uint64_t *t_total; uint32_t *stamp; int32_t *ep_step; uint8_t *owner_side, *flags; CacheRec *cache;
    uint32_t *sched;      // work queue of the time-sliced rollout: unit counter, then one progress word per warp of envs
    size_t np;
};
constexpr int kSchedHead = 64;                                        // words before the per-warp counters (256 B)
__host__ __device__ inline size_t sched_words(size_t np) { return kSchedHead + np / 32; }

__host__ __device__ inline size_t padded(int n) { return ((size_t)n + 255) & ~(size_t)255; }

size_t state_bytes(int n_envs, int n_players)
{
    const size_t B = 2 * n_players + 1, P = n_pairs((int)B);
    return padded(n_envs) * (6 * B * 8 + 8 + 4 + 4 + 1 + 1 + P * sizeof(CacheRec)) + sched_words(padded(n_envs)) * 4;
}

__host__ __device__ inline StateView make_view(void *base, int n, int n_players)
{
    const size_t B = 2 * n_players + 1, P = n_pairs((int)B);
    StateView v;
    v.np = padded(n);
    char *p = (char *)base;
    v.body = (double *)p;        p += v.np * 6 * B * 8;
    v.cache = (CacheRec *)p;     p += v.np * P * sizeof(CacheRec);
    v.t_total = (uint64_t *)p;   p += v.np * 8;
    v.stamp = (uint32_t *)p;     p += v.np * 4;
    v.ep_step = (int32_t *)p;    p += v.np * 4;
    v.owner_side = (uint8_t *)p; p += v.np;
    v.flags = (uint8_t *)p;      p += v.np;
    v.sched = (uint32_t *)p;
    return v;
}

// CG: loads from L2 (ld.global.cg) -- the time-sliced rollout reads state that a block on another SM wrote in this launch
template <bool CG = false>
__device__ __forceinline__ void load_state(const StateView &v, int i, Lane L, V1Regs &s, int B)
{
    if (CG) {
        for (int k = 0; k < 6 * B; ++k) L.f(k * kCol) = __ldcg(v.body + (size_t)k * v.np + i);
        s.t_total = __ldcg(v.t_total + i); s.stamp = __ldcg(v.stamp + i); s.ep_step = __ldcg(v.ep_step + i);
        s.owner_side = __ldcg(v.owner_side + i);
        return;
    }
    for (int k = 0; k < 6 * B; ++k) L.f(k * kCol) = v.body[(size_t)k * v.np + i];
    s.t_total = v.t_total[i]; s.stamp = v.stamp[i]; s.ep_step = v.ep_step[i]; s.owner_side = v.owner_side[i];
}

__device__ __forceinline__ void store_state(const StateView &v, int i, Lane L, const V1Regs &s, int B, int flags)
{
    for (int k = 0; k < 6 * B; ++k) v.body[(size_t)k * v.np + i] = L.f(k * kCol);
    v.t_total[i] = s.t_total; v.stamp[i] = s.stamp; v.ep_step[i] = s.ep_step; v.owner_side[i] = (uint8_t)s.owner_side;
    v.flags[i] = (uint8_t)flags;
}

template <typename T>
__device__ __forceinline__ void thread_store_obs(T *dst_row, Lane L, int N)
{
    const int D = obs_dim(N);
    for (int k = 0; k < D; ++k) dst_row[k] = (T)obs_elem(L, N, k);
}

// fp32 observations of a warp's 32 environments = one contiguous span of [n, D].  Every lane can read every lane's
// column (same warp, shared memory), so no staging copy is needed.  Lane l serves field l >> 3 (x, y, vx, vy) of
// environment 8 g + (l & 7) in pass g: the field's average, range and the refined reciprocal of the range
// (ieee_fast.cuh: same quotient as obs_elem's division) are then per-lane constants of the call, the body index is the
// loop counter, and an instruction writes the four fields = 16 contiguous bytes of one body for eight environments
// (the other bodies' stores complete the 32-byte sectors in L2).  The previous version wrote 32 consecutive floats
// per instruction but spent 53 instructions per store on index arithmetic and a full division (r1_history.md, r1k).
__device__ __forceinline__ void warp_store_obs_f32(Lane L, int N, float *gdst_warp_row0, int lane, int rows_in_warp)
{
    const int D = obs_dim(N), B = 2 * N + 1;
    __syncwarp();                                     // every lane's step has finished writing its column
    const int fld = lane >> 3, sub = lane & 7;
    const double avg = fld == 0 ? 52.5 : (fld == 1 ? 34.0 : 0.0);
    const double rng_ball = fld == 0 ? 52.5 : (fld == 1 ? 34.0 : 25.0), rng_player = fld == 0 ? 55.5 : (fld == 1 ? 34.0 : 10.0);
    const double rcp_ball = frcp_refined(rng_ball), rcp_player = frcp_refined(rng_player);
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {
        const int env = 8 * g + sub;
        if (env >= rows_in_warp) continue;
        Lane src;                                     // field `fld` of body 0 of that environment's column
        src.st = L.st - (uint32_t)lane + (uint32_t)(env + fld * kCol);
        float *dst = gdst_warp_row0 + (size_t)env * D + fld;
        {   // the ball leads the vector, :154-180
            const double num = dsub(src.f(2 * N * kBodyStride), avg);
            const bool z = num == 0.0;                               // 0 / range = that same zero (obs_elem)
            const double q = fdiv_with(pick(z, 1.0, num), rng_ball, rcp_ball);
            __stcs(dst, (float)(z ? num : q));
        }
#pragma unroll 1
        for (int b = 0; b < B - 1; ++b) {
            const double num = dsub(src.f(b * kBodyStride), avg);
            const bool z = num == 0.0;
            const double q = fdiv_with(pick(z, 1.0, num), rng_player, rcp_player);
            __stcs(dst + 4 + 4 * b, (float)(z ? num : q));
        }
    }
    __syncwarp();                                     // before the next step overwrites the columns
}

template <typename T>
__global__ void v1_reset_kernel(const __grid_constant__ V1Params P, StateView v, const uint8_t *mask, T *obs, int init)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_envs) return;
    if (mask != nullptr && mask[i] == 0) return;
    const int N = P.n_players, B = 2 * N + 1;
    const Lane L = make_lane(threadIdx.x >> 5, threadIdx.x & 31, N);
    const uint32_t env_id = P.env_id_offset + (uint32_t)i;
    V1Regs s;
    if (init) init_env(L, s, P, env_id);
    else { load_state(v, i, L, s, B); reset_env(L, s, P, env_id); }
    store_state(v, i, L, s, B, 0);
    if (obs != nullptr) thread_store_obs(obs + (size_t)i * obs_dim(N), L, N);
}

template <typename T, int REGC>
__global__ void v1_step_kernel(const __grid_constant__ V1Params P, StateView v, const uint8_t *actions, const uint8_t *opp_actions, T *obs, T *reward,
                               uint8_t *done, T *final_obs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int warp_env0 = i - lane;
    if (warp_env0 >= P.n_envs) return;
    const bool live = i < P.n_envs;
    const int rows_in_warp = min(32, P.n_envs - warp_env0);
    const int N = P.n_players, B = 2 * N + 1, D = obs_dim(N);
    const Lane L = make_lane(threadIdx.x >> 5, lane, N);
    const uint32_t env_id = P.env_id_offset + (uint32_t)i;
    V1Regs s;
    if (live) {
        const PairCache C{v.cache + i, v.np};
        Contact con[kMaxContacts];
        load_state(v, i, L, s, B);
        const StepResult r = v1_step<REGC>(L, s, P, env_id, actions + (size_t)i * 2 * N, C, con,
                                           opp_actions != nullptr ? opp_actions + (size_t)i * 2 * N : nullptr);
        if (r.done && P.auto_reset) {
            if (final_obs != nullptr) thread_store_obs(final_obs + (size_t)i * D, L, N);
            reset_env(L, s, P, env_id);
        }
        store_state(v, i, L, s, B, r.flags);
        if (reward != nullptr) reward[i] = (T)r.reward;
        if (done != nullptr) done[i] = (uint8_t)r.done;
    }
    if (obs != nullptr) {
        if (sizeof(T) == 4) warp_store_obs_f32(L, N, reinterpret_cast<float *>(obs) + (size_t)warp_env0 * D, lane, rows_in_warp);
        else if (live) thread_store_obs(obs + (size_t)i * D, L, N);
    }
}

// MINB: resident 64-thread blocks per SM the register allocation is capped for.  Measured (tools/exp_variants_v1.sh, end of
// round 2): the small teams, whose shared-memory state is small, gain from more warps -- 1v1 +3.9 % at 12 blocks (85
// registers), 2v2 +1.8 % at 10 (102) -- from 3v3 on the uncapped allocation wins.
// Steps [k0, k1) of the rollout for the warp of environments `wg` (envs 32 wg .. 32 wg + 31), run by the calling warp in its
// own shared-memory columns.  SLICED: the state was written by another block of this launch (loads from L2).
template <int REGC, bool SLICED>
__device__ __forceinline__ void v1_rollout_span(const V1Params &P, const StateView &v, int wg, int k0, int k1, const uint8_t *__restrict__ actions,
                                                const uint8_t *__restrict__ opp_actions, float *__restrict__ obs,
                                                float *__restrict__ reward, uint8_t *__restrict__ done, FutbolStats *stats)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warp_env0 = wg * 32, i = warp_env0 + lane;          // the caller guarantees warp_env0 < n_envs
    const bool live = i < P.n_envs;
    const int rows_in_warp = min(32, P.n_envs - warp_env0);
    const int N = P.n_players, B = 2 * N + 1, D = obs_dim(N);
    const size_t n = (size_t)P.n_envs;
    const uint32_t env_id = P.env_id_offset + (uint32_t)i;
    const Lane L = make_lane(warp, lane, N);
    // padding lanes of the last warp step env 0's cache column?  No: they get a private dummy state and never touch HBM
    const int ci = live ? i : 0;
    const PairCache C{v.cache + ci, v.np};
    Contact con[kMaxContacts];

    V1Regs s;
    if (live) load_state<SLICED>(v, i, L, s, B);
    else init_env(L, s, P, env_id);

    double reward_sum = 0.0;
    uint32_t episodes = 0, goals_l = 0, goals_r = 0, outs = 0, contacts = 0, overflow = 0;
    int last_flags = 0;
#pragma unroll 1
    for (int k = k0; k < k1; ++k) {
        const size_t slot = (size_t)k * n + (size_t)i;
        StepResult r;
        if (live) {
            r = v1_step<REGC>(L, s, P, env_id, actions != nullptr ? actions + slot * 2 * N : nullptr, C, con,
                              opp_actions != nullptr ? opp_actions + slot * 2 * N : nullptr);
            if (r.done && P.auto_reset) reset_env(L, s, P, env_id);
        } else {
            r.reward = 0.0; r.done = 0; r.flags = 0; r.contacts = 0; r.overflow = 0;
        }
        last_flags = r.flags;
        reward_sum += r.reward;
        goals_l += (r.flags & kFlagGoalLeft) != 0;
        goals_r += (r.flags & kFlagGoal) && !(r.flags & kFlagGoalLeft);
        outs += (r.flags & kFlagOut) != 0;
        episodes += r.done;
        contacts += r.contacts; overflow += r.overflow;
        if (obs != nullptr) warp_store_obs_f32(L, N, obs + ((size_t)k * n + (size_t)warp_env0) * D, lane, rows_in_warp);
        if (live) {
            if (reward != nullptr) __stcs(reward + slot, (float)r.reward);
            if (done != nullptr) done[slot] = (uint8_t)r.done;
        }
    }
    if (live) store_state(v, i, L, s, B, last_flags);

    if (stats != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            reward_sum += __shfl_xor_sync(0xffffffffu, reward_sum, o);
            episodes += __shfl_xor_sync(0xffffffffu, episodes, o);
            goals_l += __shfl_xor_sync(0xffffffffu, goals_l, o);
            goals_r += __shfl_xor_sync(0xffffffffu, goals_r, o);
            outs += __shfl_xor_sync(0xffffffffu, outs, o);
            contacts += __shfl_xor_sync(0xffffffffu, contacts, o);
            overflow += __shfl_xor_sync(0xffffffffu, overflow, o);
        }
        if (lane == 0) {
            atomicAdd(&stats->reward_sum, reward_sum);
            atomicAdd((unsigned long long *)&stats->env_steps, (unsigned long long)rows_in_warp * (unsigned long long)(k1 - k0));
            atomicAdd((unsigned long long *)&stats->episodes, (unsigned long long)episodes);
            atomicAdd((unsigned long long *)&stats->goals_ai, (unsigned long long)goals_l);
            atomicAdd((unsigned long long *)&stats->goals_opp, (unsigned long long)goals_r);
            atomicAdd((unsigned long long *)&stats->out_of_field, (unsigned long long)outs);
            atomicAdd((unsigned long long *)&stats->reserved[0], (unsigned long long)contacts);
            atomicAdd((unsigned long long *)&stats->reserved[1], (unsigned long long)overflow);
        }
    }
}

template <int REGC, int MINB>
__global__ void __launch_bounds__(64, MINB) v1_rollout_kernel(const __grid_constant__ V1Params P, StateView v, int K, const uint8_t *__restrict__ actions,
                                  const uint8_t *__restrict__ opp_actions, float *__restrict__ obs,
                                  float *__restrict__ reward, uint8_t *__restrict__ done, FutbolStats *stats)
{
    const int wg = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (wg * 32 >= P.n_envs) return;
    v1_rollout_span<REGC, false>(P, v, wg, 0, K, actions, opp_actions, obs, reward, done, stats);
}

// The same rollout as a queue of time slices (cf. v0_rollout_sliced_kernel, v0_kernels.cu): a few waves of K-step-long warps
// leave the SMs idle through the partial last wave (5v5, 2^18 envs: 8192 warps on 1924 slots = 4.26 waves).  The K steps are cut
// into `chunks` slices; a unit is (slice c, warp of envs g), numbered c * groups + g, and every WARP of a grid that just fills
// the GPU takes units from a counter (lane 0 + shuffle: no block barrier, no shared memory).  Unit (c, g) needs (c - 1, g),
// which has a smaller number: it was taken earlier by a warp that is running and never waits on a later unit -- no deadlock
// whatever the residency.  Hand-over through L2: every lane __threadfence()s after its stores, lane 0 publishes the progress
// word; the reader polls it, fences, and loads the state and the arbiter-cache records with ld.global.cg.
template <int REGC, int MINB>
__global__ void __launch_bounds__(64, MINB) v1_rollout_sliced_kernel(const __grid_constant__ V1Params P, StateView v, int K, int chunk_steps,
                                  int chunks, int groups, const uint8_t *__restrict__ actions,
                                  const uint8_t *__restrict__ opp_actions, float *__restrict__ obs,
                                  float *__restrict__ reward, uint8_t *__restrict__ done, FutbolStats *stats)
{
    volatile uint32_t *progress = v.sched + kSchedHead;
    const uint32_t units = (uint32_t)chunks * (uint32_t)groups;
    const int lane = threadIdx.x & 31;
    for (;;) {
        uint32_t mine = 0;
        if (lane == 0) mine = atomicAdd(v.sched, 1u);
        // broadcast as a warp reduction: REDUX writes a UNIFORM register, so the compiler knows that the unit -- and every
        // branch and address derived from it -- is the same in all lanes (a shuffle's result is not: convergence barriers and
        // vector address arithmetic all over the step, -10 % at 5v5)
        const uint32_t u = __reduce_add_sync(0xffffffffu, mine);
        if (u >= units) break;
        const int c = (int)(u / (uint32_t)groups), g = (int)(u % (uint32_t)groups);
        if (c > 0) {
            if (lane == 0) {
                while (progress[g] < (uint32_t)c) __nanosleep(200);
                __threadfence();
            }
            __syncwarp();
        }
        v1_rollout_span<REGC, true>(P, v, g, c * chunk_steps, min(K, (c + 1) * chunk_steps), actions, opp_actions, obs, reward, done, stats);
        __threadfence();
        __syncwarp();
        if (lane == 0) progress[g] = (uint32_t)(c + 1);
    }
}

// AoS record = FutbolV1EnvState header + double jn[P] + uint32 last[P], padded to 8 bytes (include/futbol_b200.h)
__host__ __device__ inline size_t record_bytes(int n_players)
{
    const size_t P = n_pairs(2 * n_players + 1);
    return (sizeof(FutbolV1EnvState) + P * 12 + 7) & ~(size_t)7;
}
size_t env_state_bytes(int n_players) { return record_bytes(n_players); }

__global__ void v1_get_state_kernel(int n, int n_players, StateView v, unsigned char *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int B = 2 * n_players + 1, P = n_pairs(B);
    FutbolV1EnvState *e = reinterpret_cast<FutbolV1EnvState *>(out + (size_t)i * record_bytes(n_players));   // written field by field
    for (int b = 0; b < 21; ++b)
        for (int f = 0; f < 6; ++f) e->body[b][f] = b < B ? v.body[(size_t)(6 * b + f) * v.np + i] : 0.0;
    e->t_total = v.t_total[i]; e->stamp = v.stamp[i]; e->ep_step = v.ep_step[i]; e->owner_side = v.owner_side[i];
    e->flags = v.flags[i];
    for (int k = 0; k < 6; ++k) e->pad_[k] = 0;
    double *jn = reinterpret_cast<double *>(e + 1);
    uint32_t *last = reinterpret_cast<uint32_t *>(jn + P);
    for (int q = 0; q < P; ++q) { const CacheRec r = v.cache[(size_t)q * v.np + i]; jn[q] = r.jn; last[q] = r.last; }
    if (P & 1) last[P] = 0;
}

__global__ void v1_set_state_kernel(int n, int n_players, StateView v, const unsigned char *in)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int B = 2 * n_players + 1, P = n_pairs(B);
    const FutbolV1EnvState *e = reinterpret_cast<const FutbolV1EnvState *>(in + (size_t)i * record_bytes(n_players));
    for (int b = 0; b < B; ++b)
        for (int f = 0; f < 6; ++f) v.body[(size_t)(6 * b + f) * v.np + i] = e->body[b][f];
    v.t_total[i] = e->t_total; v.stamp[i] = e->stamp; v.ep_step[i] = e->ep_step; v.owner_side[i] = e->owner_side;
    v.flags[i] = e->flags;
    const double *jn = reinterpret_cast<const double *>(e + 1);
    const uint32_t *last = reinterpret_cast<const uint32_t *>(jn + P);
    for (int q = 0; q < P; ++q) { CacheRec r; r.jn = jn[q]; r.last = last[q]; r.pad_ = 0; v.cache[(size_t)q * v.np + i] = r; }
}

// ---- host launchers ------------------------------------------------------------------------------------
static inline int blocks_for(int n, int t) { return (n + t - 1) / t; }
#ifdef FUTBOL_V1_THREADS
static inline int threads_for(int) { return FUTBOL_V1_THREADS; }   // tuning builds (tools/build_variant.py)
#else
// The block size that lets the most warps be resident (228 KB of shared memory per SM, 1 KB reserved per block; a warp holds
// 1536 (2N + 1) B of state): two warps per block up to 4v4 (16 warps at 4v4), one from 5v5 on (13 instead of 12 warps at 5v5,
// 11 instead of 10 at 6v6, 7 at 10v10).
static inline int threads_for(int n_players) { return n_players <= 4 ? 64 : 32; }
#endif
#ifdef FUTBOL_V1_REGC
static inline int regc_for(int) { return FUTBOL_V1_REGC; }     // tuning builds (tools/build_variant.py)
#else
static inline int regc_for(int n_players) { return n_players >= 7 ? 3 : (n_players >= 5 ? 2 : (n_players >= 3 ? 1 : 0)); }
#endif   // contacts kept in registers by the solver (v1_step.cuh space_step)
static inline int smem_for(int n_players)
{
    static const int pad = getenv("FUTBOL_V1_SMEM_PAD") ? atoi(getenv("FUTBOL_V1_SMEM_PAD")) : 0;   // occupancy experiments only
    return block_smem_bytes(n_players, threads_for(n_players) / 32) + pad;
}

cudaError_t launch_reset(const V1Params &P, void *state, const uint8_t *mask, void *obs, int obs_f64, int init, cudaStream_t st)
{
    const StateView v = make_view(state, P.n_envs, P.n_players);
    const int t = threads_for(P.n_players), sm = smem_for(P.n_players);
    if (init) {   // first construction: an empty arbiter cache
        const size_t P_ = n_pairs(2 * P.n_players + 1);
        cudaError_t e = cudaMemsetAsync(v.cache, 0, v.np * P_ * sizeof(CacheRec), st);
        if (e != cudaSuccess) return e;
    }
    if (obs_f64) v1_reset_kernel<double><<<blocks_for(P.n_envs, t), t, sm, st>>>(P, v, mask, (double *)obs, init);
    else v1_reset_kernel<float><<<blocks_for(P.n_envs, t), t, sm, st>>>(P, v, mask, (float *)obs, init);
    return cudaGetLastError();
}

cudaError_t launch_step(const V1Params &P, void *state, const uint8_t *actions, const uint8_t *opp_actions, void *obs, void *reward,
                        uint8_t *done, void *final_obs, int out_f64, cudaStream_t st)
{
    const StateView v = make_view(state, P.n_envs, P.n_players);
    const int t = threads_for(P.n_players), sm = smem_for(P.n_players);
    const int g = blocks_for(P.n_envs, t), rc = regc_for(P.n_players);
#define FUTBOL_V1_STEP(T, RC) v1_step_kernel<T, RC><<<g, t, sm, st>>>(P, v, actions, opp_actions, (T *)obs, (T *)reward, done, (T *)final_obs)
    if (out_f64) { if (rc == 3) FUTBOL_V1_STEP(double, 3); else if (rc == 2) FUTBOL_V1_STEP(double, 2); else if (rc == 1) FUTBOL_V1_STEP(double, 1); else FUTBOL_V1_STEP(double, 0); }
    else { if (rc == 3) FUTBOL_V1_STEP(float, 3); else if (rc == 2) FUTBOL_V1_STEP(float, 2); else if (rc == 1) FUTBOL_V1_STEP(float, 1); else FUTBOL_V1_STEP(float, 0); }
#undef FUTBOL_V1_STEP
    return cudaGetLastError();
}

// resident WARPS of the time-sliced rollout kernel on the current device for this team size (queried once per device and size)
template <int REGC, int MINB>
static int rollout_warp_slots(int n_players)
{
    static std::mutex mu;
    static int slots[64][kMaxN + 1] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    int &s = slots[dev & 63][n_players];
    if (s == 0) {
        int sms = 0, per_sm = 0;
        const int t = threads_for(n_players);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, v1_rollout_sliced_kernel<REGC, MINB>, t, smem_for(n_players));
        s = sms * per_sm > 0 ? sms * per_sm * (t / 32) : 1;
    }
    return s;
}

static int warp_slots_for(int n_players)
{
    const int rc = regc_for(n_players);
    // register caps of the queue kernel: the plain kernel's residency (REGC 2: 168 registers = 12 warps per SM; REGC 1: 128 = 16)
    if (rc == 3) return rollout_warp_slots<3, 1>(n_players);
    if (rc == 2) return rollout_warp_slots<2, 6>(n_players);
    if (rc == 1) return rollout_warp_slots<1, 8>(n_players);
    return n_players == 1 ? rollout_warp_slots<0, 12>(n_players) : rollout_warp_slots<0, 10>(n_players);
}

// Time slices of a K-step rollout of this batch.  slices: 0 = automatic, 1 = never slice, n > 1 = n equal slices
// (futbol_set_rollout_slices).  Automatic, from the sweep in profiles/r2_v1_history.md ("time slices"), w = warps of envs /
// resident warps: up to one wave the plain launch (slicing costs 5-12 % there); between one and two waves about eight waves
// of units (5v5, 65,536 envs: +39 %; 2v2, 131,072: +26 %); four slices up to four waves (5v5 at 2^17 envs, 2.3 waves: +19 %;
// 2v2 at 2^18: +8.6 %); two slices beyond (5v5 at 2^18 envs, 4.6 waves: +5.2 % -- four slices would give +5.8 % for twice the
// extra state traffic; 2v2 at 2^20 envs, 11 waves: +3 %).
int plan_rollout_slices(const V1Params &P, int K, int slices)
{
    const int groups = blocks_for(P.n_envs, 32);
    int n = 1;
    if (slices > 0) n = slices < K ? slices : K;
    else {
        const long long g = groups, sl = warp_slots_for(P.n_players);
        if (g > sl) {
            if (g < 2 * sl) n = (int)((8 * sl + g - 1) / g);
            else if (g < 4 * sl) n = 4;
            else n = 2;
        }
        if (n > K / 4) n = K / 4 > 0 ? K / 4 : 1;
    }
    if (n <= 1) return 1;
    const int chunk_steps = (K + n - 1) / n;
    return (K + chunk_steps - 1) / chunk_steps;
}

cudaError_t launch_rollout(const V1Params &P, void *state, int K, const uint8_t *actions, const uint8_t *opp_actions, float *obs,
                           float *reward, uint8_t *done, FutbolStats *stats, int slices, cudaStream_t st)
{
    const StateView v = make_view(state, P.n_envs, P.n_players);
    const int t = threads_for(P.n_players), sm = smem_for(P.n_players);
    const int g = blocks_for(P.n_envs, t), rc = regc_for(P.n_players);
    const int chunks = plan_rollout_slices(P, K, slices);
    static const bool force_queue = getenv("FUTBOL_V1_FORCE_QUEUE") != nullptr;      // experiments only: one slice through the queue kernel
    if (chunks > 1 || force_queue) {
        const int groups = blocks_for(P.n_envs, 32), chunk_steps = (K + chunks - 1) / chunks;
        cudaError_t e = cudaMemsetAsync(v.sched, 0, sched_words(v.np) * 4, st);
        if (e != cudaSuccess) return e;
        const long long unit_blocks = ((long long)chunks * groups + t / 32 - 1) / (t / 32);
        const long long slot_blocks = warp_slots_for(P.n_players) / (t / 32);
        const int grid = (int)(unit_blocks < slot_blocks ? unit_blocks : slot_blocks);
#define FUTBOL_V1_SLICED(RC, MB) v1_rollout_sliced_kernel<RC, MB><<<grid, t, sm, st>>>(P, v, K, chunk_steps, chunks, groups, actions, opp_actions, obs, reward, done, stats)
        if (rc == 3) FUTBOL_V1_SLICED(3, 1);
        else if (rc == 2) FUTBOL_V1_SLICED(2, 6);
        else if (rc == 1) FUTBOL_V1_SLICED(1, 8);
        else if (P.n_players == 1) FUTBOL_V1_SLICED(0, 12);
        else FUTBOL_V1_SLICED(0, 10);
#undef FUTBOL_V1_SLICED
        return cudaGetLastError();
    }
    if (rc == 3) v1_rollout_kernel<3, 1><<<g, t, sm, st>>>(P, v, K, actions, opp_actions, obs, reward, done, stats);
    else if (rc == 2) v1_rollout_kernel<2, 1><<<g, t, sm, st>>>(P, v, K, actions, opp_actions, obs, reward, done, stats);
    else if (rc == 1) v1_rollout_kernel<1, 1><<<g, t, sm, st>>>(P, v, K, actions, opp_actions, obs, reward, done, stats);
    else if (P.n_players == 1) v1_rollout_kernel<0, 12><<<g, t, sm, st>>>(P, v, K, actions, opp_actions, obs, reward, done, stats);
    else v1_rollout_kernel<0, 10><<<g, t, sm, st>>>(P, v, K, actions, opp_actions, obs, reward, done, stats);
    return cudaGetLastError();
}

cudaError_t launch_get_state(int n, int n_players, const void *state, void *aos, cudaStream_t st)
{
    v1_get_state_kernel<<<blocks_for(n, 128), 128, 0, st>>>(n, n_players, make_view(const_cast<void *>(state), n, n_players), (unsigned char *)aos);
    return cudaGetLastError();
}

cudaError_t launch_set_state(int n, int n_players, void *state, const void *aos, cudaStream_t st)
{
    v1_set_state_kernel<<<blocks_for(n, 128), 128, 0, st>>>(n, n_players, make_view(state, n, n_players), (const unsigned char *)aos);
    return cudaGetLastError();
}

}  // namespace v1
}  // namespace futbol
