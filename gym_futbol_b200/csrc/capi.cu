// The C ABI declared in include/futbol_b200.h.  Thin: validates arguments, fills the kernel
// parameter block, launches on the caller's stream.  No torch types, no host synchronisation.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <new>
#include "../../include/futbol_b200.h"
#include "v0_kernels.h"

using namespace futbol;

struct FutbolHandle {
    FutbolConfig cfg;
    V0Params v0;
    uint64_t launches;
    bool initialised;   // first futbol_reset zeroes t_total
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
    if (cfg->variant != FUTBOL_VARIANT_V0) return fail(FUTBOL_ERR_UNSUPPORTED, "variant not built: %s", "v1");
    if (cfg->n_players != 2) return fail(FUTBOL_ERR_ARG, "v0 is 2v2: n_players must be 2%s");
    if (!(cfg->game_time >= 0.0) || !(cfg->player_speed >= 0.0) || cfg->shoot_speed < 16)
        return fail(FUTBOL_ERR_ARG, "bad game_time / player_speed / shoot_speed%s");
    int dev_count = 0;
    cudaError_t e = cudaGetDeviceCount(&dev_count);
    if (e != cudaSuccess || dev_count == 0)
        return fail(FUTBOL_ERR_CUDA, "no CUDA device: %s (there is no CPU fallback)", cudaGetErrorString(e));
    FutbolHandle *h = new (std::nothrow) FutbolHandle();
    if (h == nullptr) return fail(FUTBOL_ERR_ARG, "out of host memory%s");
    h->cfg = *cfg;
    h->launches = 0;
    h->initialised = false;
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

size_t futbol_state_bytes(const FutbolHandle *h) { return h ? v0_state_bytes(h->cfg.n_envs) : 0; }
size_t futbol_env_state_bytes(const FutbolHandle *h) { return h ? sizeof(FutbolV0EnvState) : 0; }
int futbol_obs_dim(const FutbolHandle *h) { return h ? 30 : 0; }
int futbol_act_dim(const FutbolHandle *h) { return h ? 1 : 0; }
int futbol_draw_limit_steps(const FutbolHandle *h) { return h ? h->v0.ep_limit + 1 : 0; }
uint64_t futbol_launch_count(const FutbolHandle *h) { return h ? h->launches : 0; }

int futbol_reset(FutbolHandle *h, void *state, const uint8_t *mask, void *obs, int obs_dtype, void *stream)
{
    if (h == nullptr || state == nullptr) return fail(FUTBOL_ERR_ARG, "null handle/state%s");
    if (obs_dtype != 0 && obs_dtype != 1) return fail(FUTBOL_ERR_ARG, "obs_dtype must be 0 (f32) or 1 (f64)%s");
    const int init = (!h->initialised && mask == nullptr) ? 1 : 0;
    if (!h->initialised && mask != nullptr) return fail(FUTBOL_ERR_ARG, "first reset must cover all envs (mask = NULL)%s");
    cudaError_t e = v0_launch_reset(h->v0, state, mask, obs, obs_dtype, init, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e);
    h->initialised = true;
    h->launches += 1;
    return FUTBOL_OK;
}

int futbol_step(FutbolHandle *h, void *state, const uint8_t *actions, void *obs, void *reward, uint8_t *done,
                void *final_obs, int out_dtype, void *stream)
{
    if (h == nullptr || state == nullptr || actions == nullptr) return fail(FUTBOL_ERR_ARG, "null handle/state/actions%s");
    if (out_dtype != 0 && out_dtype != 1) return fail(FUTBOL_ERR_ARG, "out_dtype must be 0 (f32) or 1 (f64)%s");
    if (!h->initialised) return fail(FUTBOL_ERR_ARG, "futbol_reset must be called before futbol_step%s");
    cudaError_t e = v0_launch_step(h->v0, state, actions, obs, reward, done, final_obs, out_dtype, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e);
    h->launches += 1;
    return FUTBOL_OK;
}

int futbol_rollout(FutbolHandle *h, void *state, int K, const uint8_t *actions, float *obs, float *reward,
                   uint8_t *done, FutbolStats *stats, void *stream)
{
    if (h == nullptr || state == nullptr) return fail(FUTBOL_ERR_ARG, "null handle/state%s");
    if (K <= 0) return fail(FUTBOL_ERR_ARG, "K must be positive%s");
    if (!h->initialised) return fail(FUTBOL_ERR_ARG, "futbol_reset must be called before futbol_rollout%s");
    cudaError_t e = v0_launch_rollout(h->v0, state, K, actions, obs, reward, done, stats, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e);
    h->launches += 1;
    return FUTBOL_OK;
}

int futbol_get_state(FutbolHandle *h, const void *state, void *aos_out, void *stream)
{
    if (h == nullptr || state == nullptr || aos_out == nullptr) return fail(FUTBOL_ERR_ARG, "null argument%s");
    cudaError_t e = v0_launch_get_state(h->cfg.n_envs, state, aos_out, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e);
    h->launches += 1;
    return FUTBOL_OK;
}

int futbol_set_state(FutbolHandle *h, void *state, const void *aos_in, void *stream)
{
    if (h == nullptr || state == nullptr || aos_in == nullptr) return fail(FUTBOL_ERR_ARG, "null argument%s");
    cudaError_t e = v0_launch_set_state(h->cfg.n_envs, state, aos_in, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e);
    h->initialised = true;
    h->launches += 1;
    return FUTBOL_OK;
}

}  // extern "C"
