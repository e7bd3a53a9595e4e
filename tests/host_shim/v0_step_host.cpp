// TEST-ONLY development aid: compiles the DEVICE header gym_futbol_b200/csrc/v0_step.cuh with g++ by
// shimming the handful of CUDA intrinsics it uses, so that the kernel's step logic can be compared with
// the oracle on a machine without a GPU (tests/test_v0_step_host.py).  Nothing in the package loads this.
// Built with -ffp-contract=off: every shimmed operation is one IEEE double operation, as on the device.
// One "lane": the shared-memory columns of the device layout collapse to a plain array (FUTBOL_LANES = 1).
#include <cmath>
#include <cstdint>
#include <cstring>
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define FUTBOL_LANES 1
#define FUTBOL_HOST_SHIM 1
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
static inline long long __double_as_longlong(double x) { long long v; std::memcpy(&v, &x, 8); return v; }
static inline double __longlong_as_double(long long v) { double x; std::memcpy(&x, &v, 8); return x; }
using std::rint;
#include "../../gym_futbol_b200/csrc/v0_step.cuh"
#include "../../include/futbol_b200.h"

using namespace futbol;

extern "C" {
// Steps `n` AoS env records `steps` times with actions[steps][n]; auto-reset as the rollout kernel does.
// Outputs (any may be null): obs[steps][n][30], reward[steps][n], done[steps][n], flags[steps][n].
void host_v0_rollout(uint64_t seed, uint32_t env_id0, int random_opp, int one_goal_end, int only_reward_goal,
                     int auto_reset, int ep_limit, int shoot_speed, double player_speed, double reach_sq_max,
                     FutbolV0EnvState *envs, int n, int steps, const uint8_t *actions,
                     double *obs, double *reward, uint8_t *done, uint8_t *flags)
{
    V0Params P;
    P.seed = seed; P.key = philox_expand_key(seed); P.env_id_offset = env_id0; P.n_envs = n; P.random_opp = random_opp;
    P.one_goal_end = one_goal_end; P.only_reward_goal = only_reward_goal; P.auto_reset = auto_reset; P.ep_limit = ep_limit;
    P.shoot_speed = shoot_speed; P.player_speed = player_speed; P.reach_sq_max = reach_sq_max;
    const Lane L = make_lane(0, 0);
    for (int i = 0; i < n; ++i) {
        V0Regs s;
        FutbolV0EnvState &e = envs[i];
        for (int k = 0; k < 25; ++k) L.f(k * kLanes) = e.rows[k / 5][k % 5];
        s.t_total = e.t_total; s.ep_step = e.ep_step; s.ai_score = e.ai_score; s.opp_score = e.opp_score;
        s.owner = e.owner; s.last_owner = e.last_owner;
        int last_flags = e.flags;
        for (int k = 0; k < steps; ++k) {
            const size_t slot = (size_t)k * n + i;
            const StepResult r = random_opp ? v0_step<true>(L, s, P, env_id0 + (uint32_t)i, actions[slot] & 15)
                                            : v0_step<false>(L, s, P, env_id0 + (uint32_t)i, actions[slot] & 15);
            last_flags = r.flags;
            if (r.done && auto_reset) reset_env(L, s);
            if (obs) {
                for (int j = 0; j < 25; ++j) obs[slot * 30 + j] = L.f(j * kLanes);
                for (int j = 0; j < 5; ++j) obs[slot * 30 + 25 + j] = obs_owner_elem(s, j);
            }
            if (reward) reward[slot] = r.reward;
            if (done) done[slot] = (uint8_t)r.done;
            if (flags) flags[slot] = (uint8_t)r.flags;
        }
        for (int k = 0; k < 25; ++k) e.rows[k / 5][k % 5] = L.f(k * kLanes);
        e.t_total = s.t_total; e.ep_step = s.ep_step; e.ai_score = s.ai_score; e.opp_score = s.opp_score;
        e.owner = (uint8_t)s.owner; e.last_owner = (uint8_t)s.last_owner; e.flags = (uint8_t)last_flags;
    }
}
}
