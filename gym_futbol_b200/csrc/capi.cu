// The C ABI declared in include/futbol_b200.h.  Thin: validates arguments, fills the kernel
// parameter block, launches on the caller's stream.  No torch types, no host synchronisation.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <new>
#include "../../include/futbol_b200.h"
#include "v0_kernels.h"
#include "v1_kernels.h"

using namespace futbol;

namespace futbol {
cudaError_t launch_selftest_arith(const double *a, const double *b, unsigned long long *mismatch, size_t n, cudaStream_t st);
cudaError_t launch_gather_minibatch(const long long *idx, long long m, long long rows, const float *obs, int obs_dim, float *obs_out,
                                    const uint8_t *act, uint8_t *act_out, const float *c0, float *c0_out, const float *c1,
                                    float *c1_out, const float *c2, float *c2_out, const float *c3, float *c3_out,
                                    unsigned long long *bad, cudaStream_t st);
cudaError_t launch_sample_actions(const void *logits, int bf16, long long n, int n_actions, unsigned long long seed,
                                  const unsigned long long *t_base, unsigned long long t_off, uint8_t *actions, float *logp, cudaStream_t st);
cudaError_t launch_gae(const float *reward, const uint8_t *done, const float *value, float gamma, float lam, float *adv,
                       float *ret, int T, int n, cudaStream_t st);
}

struct FutbolHandle {
    FutbolConfig cfg;
    V0Params v0;
    v1::V1Params v1;
    bool is_v1;
    uint64_t launches;
    bool initialised;   // first futbol_reset zeroes t_total
    int rollout_slices; // 0 = chosen per launch from the batch size (v0_kernels.cu)
    int rollout_variant; // 0 = automatic, 1 = standard kernel, 2 = dense kernel (futbol_set_rollout_variant)
};

static thread_local char g_err[256] = "";

static int fail(int code, const char *fmt, const char *detail = "")
{
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}

static int cuda_fail(cudaError_t e) { return fail(FUTBOL_ERR_CUDA, "CUDA error: %s", cudaGetErrorString(e)); }

// Number of `time += 0.1` additions after which `time >= game_time` first holds
// (futbol_env.py:712-716: the test precedes the increment, so done is first returned by step
// ep_limit + 1; 400 -> step 401 for game_time = 40).
static int episode_limit(double game_time)
{
    double t = 0.0;   // the reference starts from int 0; 0 + 0.1 is the same double
    int k = 0;
    while (!(t >= game_time) && k < (1 << 30)) { t += 0.1; ++k; }
    return k;
}

// v1: `current_time += 0.1; done = current_time > total_time` (envs_v1/futbol_env.py:478-481): the step count
// at which done first holds (300 for total_time = 30: the float sum of 300 x 0.1 is 30.000000000000156).
static int episode_limit_v1(double total_time)
{
    double t = 0.0;
    int k = 0;
    do { t += 0.1; ++k; } while (!(t > total_time) && k < (1 << 30));
    return k;
}

// kick-off formation of one side, envs_v1/team.py:52-112 (Python float arithmetic, left to right)
static void formation_v1(int n, bool right, double *xs, double *ys)
{
    const double w = 105.0, h = 68.0;
    if (n <= 3) {
        for (int i = 0; i < n; ++i) { xs[i] = right ? w * 0.75 : w * 0.25; ys[i] = (h / (n + 1)) * (i + 1); }
    } else if (n <= 6) {
        for (int i = 0; i < n; ++i) xs[i] = i < 3 ? (right ? w * 5 / 6 : w * 1 / 6) : (right ? w * 4 / 6 : w * 2 / 6);
        for (int i = 0; i < 3; ++i) ys[i] = (h / (3 + 1)) * (i + 1);
        for (int i = 0; i < n - 3; ++i) ys[3 + i] = (h / (n - 3 + 1)) * (i + 1);
    } else {
        for (int i = 0; i < n; ++i)
            xs[i] = i < 4 ? (right ? w * 7 / 8 : w * 1 / 8) : (i < 7 ? (right ? w * 6 / 8 : w * 2 / 8) : (right ? w * 5 / 8 : w * 3 / 8));
        for (int i = 0; i < 4; ++i) ys[i] = (h / (4 + 1)) * (i + 1);
        for (int i = 0; i < 3; ++i) ys[4 + i] = (h / (3 + 1)) * (i + 1);
        for (int i = 0; i < n - 7; ++i) ys[7 + i] = (h / (n - 7 + 1)) * (i + 1);
    }
}

// Largest double s with sqrt(s) < r (IEEE sqrt is correctly rounded, hence monotone), or -1 if none:
// lets the kernel test `sqrt(s) < r` as `s <= bound` without taking the root.
static double sqrt_less_than_bound(double r)
{
    if (!(r > 0.0)) return -1.0;
    double s = r * r;
    while (sqrt(s) >= r) s = nextafter(s, -INFINITY);
    while (sqrt(nextafter(s, INFINITY)) < r) s = nextafter(s, INFINITY);
    return s;
}

extern "C" {

int futbol_abi_version(void) { return FUTBOL_ABI_VERSION; }
const char *futbol_last_error(void) { return g_err; }

int futbol_create(const FutbolConfig *cfg, FutbolHandle **out)
{
    if (cfg == nullptr || out == nullptr) return fail(FUTBOL_ERR_ARG, "null argument%s");
    if (cfg->abi_version != FUTBOL_ABI_VERSION) return fail(FUTBOL_ERR_ARG, "abi_version mismatch%s");
    if (cfg->n_envs <= 0) return fail(FUTBOL_ERR_ARG, "n_envs must be positive%s");
    const bool is_v1 = cfg->variant == FUTBOL_VARIANT_V1;
    if (cfg->variant != FUTBOL_VARIANT_V0 && !is_v1) return fail(FUTBOL_ERR_UNSUPPORTED, "unknown variant%s");
    if (!is_v1) {
        if (cfg->n_players != 2) return fail(FUTBOL_ERR_ARG, "v0 is 2v2: n_players must be 2%s");
        // upper bounds: the kernels' guard-free division / square root (csrc/ieee_fast.cuh) are proven for pitch-scale
        // operands only, shoot_speed - 16 + r must not overflow, and the episode limit is found by repeated addition
        if (!(cfg->game_time >= 0.0 && cfg->game_time <= 1e6) || !(cfg->player_speed >= 0.0 && cfg->player_speed <= 1e4) ||
            cfg->shoot_speed < 16 || cfg->shoot_speed > 10000)
            return fail(FUTBOL_ERR_ARG, "bad game_time (0..1e6) / player_speed (0..1e4) / shoot_speed (16..10000)%s");
    } else {
        if (cfg->n_players < 1 || cfg->n_players > v1::kMaxN) return fail(FUTBOL_ERR_ARG, "v1: n_players must be 1..10%s");
        if (!(cfg->game_time >= 0.0 && cfg->game_time <= 1e6)) return fail(FUTBOL_ERR_ARG, "bad total_time (0..1e6)%s");
    }
    int dev_count = 0;
    cudaError_t e = cudaGetDeviceCount(&dev_count);
    if (e != cudaSuccess || dev_count == 0)
        return fail(FUTBOL_ERR_CUDA, "no CUDA device: %s (there is no CPU fallback)", cudaGetErrorString(e));
    FutbolHandle *h = new (std::nothrow) FutbolHandle();
    if (h == nullptr) return fail(FUTBOL_ERR_ARG, "out of host memory%s");
    h->cfg = *cfg;
    h->launches = 0;
    h->initialised = false;
    h->rollout_slices = 0;
    h->rollout_variant = 0;
    h->is_v1 = is_v1;
    if (is_v1) {
        v1::V1Params &Q = h->v1;
        memset(&Q, 0, sizeof(Q));
        Q.seed = cfg->seed;
        Q.key = philox_expand_key(cfg->seed);
        Q.env_id_offset = cfg->env_id_offset;
        Q.n_envs = cfg->n_envs;
        Q.n_players = cfg->n_players;
        Q.ep_limit = episode_limit_v1(cfg->game_time);
        Q.auto_reset = cfg->auto_reset != 0;
        // Literals, not pow() calls: the compiler may fold pow() of constants with correct rounding while Chipmunk
        // and CPython call the C library at run time (glibc pow is not correctly rounded).  These are the run-time
        // values (tests/test_oracle_v1_golden.py checks them against math.pow and the oracle's configuration).
        Q.damping_dt = 0x1.fd6168eb56e59p-1;     // pow(0.95, 0.1): space.damping ** TIME_STEP (:99)
        Q.bias_coef = 0x1.dfcdf3e02c8a4p-2;      // 1 - pow(pow(1.0f - 0.1f, 60.0f), 0.1): cpSpace.c's default collision_bias
        Q.clamp_sq_player = sqrt_less_than_bound(nextafter(v1::kPlayerMaxV, INFINITY));   // sqrt(s) > 10  <=>  s > bound
        Q.clamp_sq_ball = sqrt_less_than_bound(nextafter(v1::kBallMaxV, INFINITY));
        formation_v1(cfg->n_players, false, Q.form_x, Q.form_y);
        formation_v1(cfg->n_players, true, Q.form_x + cfg->n_players, Q.form_y + cfg->n_players);
    }
    V0Params &P = h->v0;
    P.seed = cfg->seed;
    P.key = philox_expand_key(cfg->seed);
    P.env_id_offset = cfg->env_id_offset;
    P.n_envs = cfg->n_envs;
    P.random_opp = cfg->random_opp != 0;
    P.one_goal_end = cfg->one_goal_end != 0;
    P.only_reward_goal = cfg->only_reward_goal != 0;
    P.auto_reset = cfg->auto_reset != 0;
    P.ep_limit = episode_limit(cfg->game_time);
    P.shoot_speed = cfg->shoot_speed;
    P.player_speed = cfg->player_speed;
    P.reach_sq_max = sqrt_less_than_bound(0.1 * cfg->player_speed);   // futbol_env.py:972 `< STEP_SIZE * player_speed`
    *out = h;
    return FUTBOL_OK;
}

int futbol_destroy(FutbolHandle *h)
{
    delete h;
    return FUTBOL_OK;
}

size_t futbol_state_bytes(const FutbolHandle *h)
{
    if (!h) return 0;
    return h->is_v1 ? v1::state_bytes(h->cfg.n_envs, h->cfg.n_players) : v0_state_bytes(h->cfg.n_envs);
}
size_t futbol_env_state_bytes(const FutbolHandle *h) { return h ? (h->is_v1 ? v1::env_state_bytes(h->cfg.n_players) : sizeof(FutbolV0EnvState)) : 0; }
int futbol_obs_dim(const FutbolHandle *h) { return h ? (h->is_v1 ? v1::obs_dim(h->cfg.n_players) : 30) : 0; }
int futbol_act_dim(const FutbolHandle *h) { return h ? (h->is_v1 ? 2 * h->cfg.n_players : 1) : 0; }
int futbol_draw_limit_steps(const FutbolHandle *h) { return h ? (h->is_v1 ? h->v1.ep_limit : h->v0.ep_limit + 1) : 0; }
uint64_t futbol_launch_count(const FutbolHandle *h) { return h ? h->launches : 0; }

int futbol_rollout_slices(FutbolHandle *h, int K)
{
    if (h == nullptr || K <= 0) return fail(FUTBOL_ERR_ARG, "null handle or K <= 0%s");
    return h->is_v1 ? v1::plan_rollout_slices(h->v1, K, h->rollout_slices) : v0_plan_rollout(h->v0, K, h->rollout_slices, h->rollout_variant).slices;
}

int futbol_rollout_kernel(FutbolHandle *h, int K)
{
    if (h == nullptr || K <= 0) return fail(FUTBOL_ERR_ARG, "null handle or K <= 0%s");
    if (h->is_v1) return v1::plan_rollout_slices(h->v1, K, h->rollout_slices) > 1 ? 1 : 0;
    const V0RolloutChoice c = v0_plan_rollout(h->v0, K, h->rollout_slices, h->rollout_variant);
    return c.kernel == 1 ? 2 : (c.slices > 1 ? 1 : 0);
}

int futbol_set_rollout_variant(FutbolHandle *h, int variant)
{
    if (h == nullptr || variant < 0 || variant > 2) return fail(FUTBOL_ERR_ARG, "null handle or variant not in 0..2%s");
    h->rollout_variant = variant;
    return FUTBOL_OK;
}

int futbol_set_rollout_slices(FutbolHandle *h, int slices)
{
    if (h == nullptr || slices < 0) return fail(FUTBOL_ERR_ARG, "null handle or negative slice count%s");
    h->rollout_slices = slices;
    return FUTBOL_OK;
}

int futbol_reset(FutbolHandle *h, void *state, const uint8_t *mask, void *obs, int obs_dtype, void *stream)
{
    if (h == nullptr || state == nullptr) return fail(FUTBOL_ERR_ARG, "null handle/state%s");
    if (obs_dtype != 0 && obs_dtype != 1) return fail(FUTBOL_ERR_ARG, "obs_dtype must be 0 (f32) or 1 (f64)%s");
    const int init = (!h->initialised && mask == nullptr) ? 1 : 0;
    if (!h->initialised && mask != nullptr) return fail(FUTBOL_ERR_ARG, "first reset must cover all envs (mask = NULL)%s");
    cudaError_t e = h->is_v1 ? v1::launch_reset(h->v1, state, mask, obs, obs_dtype, init, (cudaStream_t)stream)
                             : v0_launch_reset(h->v0, state, mask, obs, obs_dtype, init, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e);
    h->initialised = true;
    h->launches += 1;
    return FUTBOL_OK;
}

int futbol_step_vs(FutbolHandle *h, void *state, const uint8_t *actions, const uint8_t *opp_actions, void *obs, void *reward,
                   uint8_t *done, void *final_obs, int out_dtype, void *stream)
{
    if (h != nullptr && opp_actions != nullptr && !h->is_v1 && !h->cfg.random_opp)
        return fail(FUTBOL_ERR_ARG, "v0: opponent actions can only replace the RANDOM opponents (create with random_opp = 1)%s");
    if (h == nullptr || state == nullptr || actions == nullptr) return fail(FUTBOL_ERR_ARG, "null handle/state/actions%s");
    if (out_dtype != 0 && out_dtype != 1) return fail(FUTBOL_ERR_ARG, "out_dtype must be 0 (f32) or 1 (f64)%s");
    if (!h->initialised) return fail(FUTBOL_ERR_ARG, "futbol_reset must be called before futbol_step%s");
    if (h->is_v1 && ((uintptr_t)actions & 1u)) return fail(FUTBOL_ERR_ARG, "v1: the action buffer must be 2-byte aligned%s");
    cudaError_t e = h->is_v1 ? v1::launch_step(h->v1, state, actions, opp_actions, obs, reward, done, final_obs, out_dtype, (cudaStream_t)stream)
                             : v0_launch_step(h->v0, state, actions, opp_actions, obs, reward, done, final_obs, out_dtype, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e);
    h->launches += 1;
    return FUTBOL_OK;
}

int futbol_step(FutbolHandle *h, void *state, const uint8_t *actions, void *obs, void *reward, uint8_t *done,
                void *final_obs, int out_dtype, void *stream)
{
    return futbol_step_vs(h, state, actions, nullptr, obs, reward, done, final_obs, out_dtype, stream);
}

int futbol_rollout_vs(FutbolHandle *h, void *state, int K, const uint8_t *actions, const uint8_t *opp_actions, float *obs,
                      float *reward, uint8_t *done, FutbolStats *stats, void *stream)
{
    if (h != nullptr && opp_actions != nullptr && !h->is_v1 && !h->cfg.random_opp)
        return fail(FUTBOL_ERR_ARG, "v0: opponent actions can only replace the RANDOM opponents (create with random_opp = 1)%s");
    if (h == nullptr || state == nullptr) return fail(FUTBOL_ERR_ARG, "null handle/state%s");
    if (K <= 0) return fail(FUTBOL_ERR_ARG, "K must be positive%s");
    if (!h->initialised) return fail(FUTBOL_ERR_ARG, "futbol_reset must be called before futbol_rollout%s");
    if (h->is_v1 && ((uintptr_t)actions & 1u)) return fail(FUTBOL_ERR_ARG, "v1: the action buffer must be 2-byte aligned%s");
    cudaError_t e = h->is_v1 ? v1::launch_rollout(h->v1, state, K, actions, opp_actions, obs, reward, done, stats, h->rollout_slices, (cudaStream_t)stream)
                             : v0_launch_rollout(h->v0, state, K, actions, opp_actions, obs, reward, done, stats, h->rollout_slices, h->rollout_variant, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e);
    h->launches += 1;
    return FUTBOL_OK;
}

int futbol_rollout(FutbolHandle *h, void *state, int K, const uint8_t *actions, float *obs, float *reward,
                   uint8_t *done, FutbolStats *stats, void *stream)
{
    return futbol_rollout_vs(h, state, K, actions, nullptr, obs, reward, done, stats, stream);
}

int futbol_gae(const float *reward, const uint8_t *done, const float *value, float gamma, float lam, float *adv, float *ret,
               int T, int n, void *stream)
{
    if (reward == nullptr || done == nullptr || value == nullptr || adv == nullptr || ret == nullptr)
        return fail(FUTBOL_ERR_ARG, "null argument%s");
    if (T <= 0 || n <= 0) return fail(FUTBOL_ERR_ARG, "T and n must be positive%s");
    cudaError_t e = launch_gae(reward, done, value, gamma, lam, adv, ret, T, n, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e);
    return FUTBOL_OK;
}

int futbol_gather_minibatch(const int64_t *idx, int64_t m, int64_t rows, const float *obs, int obs_dim, float *obs_out,
                            const uint8_t *act, uint8_t *act_out, const float *c0, float *c0_out, const float *c1, float *c1_out,
                            const float *c2, float *c2_out, const float *c3, float *c3_out, uint64_t *bad, void *stream)
{
    if (idx == nullptr || m <= 0 || rows <= 0) return fail(FUTBOL_ERR_ARG, "null index or empty minibatch%s");
    if ((obs != nullptr) != (obs_out != nullptr) || (act != nullptr) != (act_out != nullptr) || (c0 != nullptr) != (c0_out != nullptr) ||
        (c1 != nullptr) != (c1_out != nullptr) || (c2 != nullptr) != (c2_out != nullptr) || (c3 != nullptr) != (c3_out != nullptr))
        return fail(FUTBOL_ERR_ARG, "every source column needs its destination (and vice versa)%s");
    if (obs != nullptr && obs_dim <= 0) return fail(FUTBOL_ERR_ARG, "obs_dim must be positive%s");
    cudaError_t e = launch_gather_minibatch((const long long *)idx, m, rows, obs, obs_dim, obs_out, act, act_out, c0, c0_out, c1, c1_out,
                                            c2, c2_out, c3, c3_out, (unsigned long long *)bad, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e);
    return FUTBOL_OK;
}

int futbol_sample_actions(const void *logits, int logits_dtype, int64_t n, int n_actions, uint64_t seed, const uint64_t *t_base,
                          uint64_t t_off, uint8_t *actions, float *logp, void *stream)
{
    if (logits == nullptr || actions == nullptr) return fail(FUTBOL_ERR_ARG, "null logits or actions%s");
    if (logits_dtype != 0 && logits_dtype != 1) return fail(FUTBOL_ERR_ARG, "logits_dtype must be 0 (f32) or 1 (bf16)%s");
    if (n <= 0 || n_actions < 1 || n_actions > 32) return fail(FUTBOL_ERR_ARG, "n must be positive and n_actions in 1..32%s");
    cudaError_t e = launch_sample_actions(logits, logits_dtype, n, n_actions, seed, (const unsigned long long *)t_base, t_off, actions, logp,
                                          (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e);
    return FUTBOL_OK;
}

int futbol_selftest_arith(const double *a, const double *b, uint64_t *mismatch, size_t n, void *stream)
{
    if (a == nullptr || b == nullptr || mismatch == nullptr || n == 0) return fail(FUTBOL_ERR_ARG, "null argument%s");
    cudaError_t e = launch_selftest_arith(a, b, (unsigned long long *)mismatch, n, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e);
    return FUTBOL_OK;
}

int futbol_get_state(FutbolHandle *h, const void *state, void *aos_out, void *stream)
{
    if (h == nullptr || state == nullptr || aos_out == nullptr) return fail(FUTBOL_ERR_ARG, "null argument%s");
    cudaError_t e = h->is_v1 ? v1::launch_get_state(h->cfg.n_envs, h->cfg.n_players, state, aos_out, (cudaStream_t)stream)
                             : v0_launch_get_state(h->cfg.n_envs, state, aos_out, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e);
    h->launches += 1;
    return FUTBOL_OK;
}

int futbol_set_state(FutbolHandle *h, void *state, const void *aos_in, void *stream)
{
    if (h == nullptr || state == nullptr || aos_in == nullptr) return fail(FUTBOL_ERR_ARG, "null argument%s");
    cudaError_t e = h->is_v1 ? v1::launch_set_state(h->cfg.n_envs, h->cfg.n_players, state, aos_in, (cudaStream_t)stream)
                             : v0_launch_set_state(h->cfg.n_envs, state, aos_in, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e);
    h->initialised = true;
    h->launches += 1;
    return FUTBOL_OK;
}

}  // extern "C"
