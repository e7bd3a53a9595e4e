// TEST-ONLY development aid (see v0_step_host.cpp): the DEVICE header gym_futbol_b200/csrc/v1_step.cuh compiled
// with g++ (CUDA intrinsics shimmed, one lane) so that the v1 kernel's step logic can be compared with the
// oracle on a machine without a GPU.  Nothing in the package loads this.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#define __device__
#define __host__
#define __forceinline__ inline
#define __align__(n) __attribute__((aligned(n)))
#define __noinline__
#define FUTBOL_LANES 1
#define FUTBOL_HOST_SHIM 1
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
#include "../../gym_futbol_b200/csrc/v1_step.cuh"

using namespace futbol;
using namespace futbol::v1;

extern "C" {
// n envs with global ids env_id0.., constructed + reset, then `steps` steps each with auto-reset (VecEnv semantics).
// left_actions: uint8 [steps][n][2N] or null (synthetic, Philox stream 1).  Outputs may be null.
void host_v1_rollout(uint64_t seed, uint32_t env_id0, int n_players, int ep_limit, double damping_dt, double bias_coef,
                     const double *form_x, const double *form_y, int n, int steps, const uint8_t *left_actions,
                     double *obs, double *reward, uint8_t *done, uint8_t *flags, int32_t *contacts)
{
    V1Params P;
    std::memset(&P, 0, sizeof(P));
    P.seed = seed; P.key = philox_expand_key(seed); P.env_id_offset = env_id0; P.n_envs = n; P.n_players = n_players;
    P.ep_limit = ep_limit; P.auto_reset = 1; P.damping_dt = damping_dt; P.bias_coef = bias_coef;
    for (int which = 0; which < 2; ++which) {      // largest s with sqrt(s) <= max (capi.cu computes the same bound)
        const double mx = which ? kBallMaxV : kPlayerMaxV;
        double sq = mx * mx;
        while (std::sqrt(sq) > mx) sq = std::nextafter(sq, -INFINITY);
        while (std::sqrt(std::nextafter(sq, INFINITY)) <= mx) sq = std::nextafter(sq, INFINITY);
        (which ? P.clamp_sq_ball : P.clamp_sq_player) = sq;
    }
    for (int i = 0; i < 2 * n_players; ++i) { P.form_x[i] = form_x[i]; P.form_y[i] = form_y[i]; }
    const int N = n_players, B = 2 * N + 1, D = 4 + 8 * N, NP = n_pairs(B);
    const Lane L = make_lane(0, 0, N);
    std::vector<CacheRec> cache(NP);
    Contact con[kMaxContacts];
    for (int i = 0; i < n; ++i) {
        std::memset(cache.data(), 0, NP * sizeof(CacheRec));
        PairCache C{cache.data(), 1};
        V1Regs s;
        const uint32_t env_id = env_id0 + (uint32_t)i;
        init_env(L, s, P, env_id);
        for (int k = 0; k < steps; ++k) {
            const size_t slot = (size_t)k * n + i;
            const StepResult r = (N >= 7) ? v1_step<3>(L, s, P, env_id, left_actions ? left_actions + slot * 2 * N : nullptr, C, con)
                                 : (N >= 4) ? v1_step<2>(L, s, P, env_id, left_actions ? left_actions + slot * 2 * N : nullptr, C, con)
                                 : (N >= 2) ? v1_step<1>(L, s, P, env_id, left_actions ? left_actions + slot * 2 * N : nullptr, C, con)
                                          : v1_step<0>(L, s, P, env_id, left_actions ? left_actions + slot * 2 * N : nullptr, C, con);
            if (r.done) reset_env(L, s, P, env_id);
            if (obs) for (int e = 0; e < D; ++e) obs[slot * D + e] = obs_elem(L, N, e);
            if (reward) reward[slot] = r.reward;
            if (done) done[slot] = (uint8_t)r.done;
            if (flags) flags[slot] = (uint8_t)r.flags;
            if (contacts) contacts[slot] = r.contacts;
        }
    }
}
}
