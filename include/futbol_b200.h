/*
 * futbol_b200 -- C ABI of the B200-native batched gym-futbol environment step.
 *
 * This is the drop-in boundary for ONE hot path of yc2454/gym-futbol: the environment
 * step (reference: gym_futbol/envs/futbol_env.py FutbolEnv.reset :205-245 / .step :628-717
 * and gym_futbol/envs_v1/futbol_env.py Futbol.reset :145-150 / .step :427-483).  The
 * reference has no native interface (it is pure Python; v1 calls Chipmunk2D through
 * pymunk/cffi), so every entry point below names the Python method it replaces.  The
 * reference-side binding a maintainer would add is the ctypes stub in INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; every data pointer is a DEVICE pointer owned by the caller
 *     (e.g. a torch CUDA tensor's data_ptr()); `stream` is a cudaStream_t passed as void*.
 *   - all work is enqueued on `stream`; no call synchronises the host.
 *   - return value: 0 = ok, <0 = error (see futbol_last_error()); nothing throws.
 *   - one handle per (process, device); calls on a handle are serialised by the caller.
 *   - the library allocates nothing persistent on the device except a 64-byte statistics
 *     scratch inside the handle.
 */
#ifndef FUTBOL_B200_H
#define FUTBOL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FUTBOL_ABI_VERSION 2

enum { FUTBOL_VARIANT_V0 = 0, /* FutbolEnv: 2v2 kinematic, possession state machine */
       FUTBOL_VARIANT_V1 = 1  /* Futbol: NvN, circle/segment rigid-body physics     */ };

enum { FUTBOL_OK = 0, FUTBOL_ERR_ARG = -1, FUTBOL_ERR_CUDA = -2, FUTBOL_ERR_UNSUPPORTED = -3 };

/* Constructor arguments of the reference classes (v0: futbol_env.py:134-138,
 * v1: envs_v1/futbol_env.py:63-65) plus what batching adds. */
typedef struct FutbolConfig {
    int32_t  abi_version;      /* FUTBOL_ABI_VERSION */
    int32_t  variant;          /* FUTBOL_VARIANT_* */
    int32_t  n_envs;           /* envs held by this handle (this device's shard) */
    uint32_t env_id_offset;    /* global id of local env 0: RNG is keyed by global id, so
                                  trajectories do not depend on how envs are sharded */
    uint64_t seed;
    int32_t  n_players;        /* players per team: v0 must be 2; v1 1..10 (number_of_player) */
    int32_t  random_opp;       /* v0 random_opp (:138) */
    int32_t  one_goal_end;     /* v0 one_goal_end (:137) */
    int32_t  only_reward_goal; /* v0 only_reward_goal (:137) */
    int32_t  auto_reset;       /* 1: VecEnv semantics -- an env that returns done is reset in the
                                  same call and its obs slot holds the reset observation */
    int32_t  shoot_speed;      /* v0 shoot_speed (:136), default 20 */
    double   game_time;        /* v0 game_time (:135) = 40; v1 total_time (:63) = 30 */
    double   player_speed;     /* v0 player_speed (:135) = 12 */
} FutbolConfig;

typedef struct FutbolHandle FutbolHandle;

/* One env's complete v0 state in array-of-structs form, for get/set_state (checkpoint,
 * parity tests).  In HBM the state is structure-of-arrays; see DESIGN.md. */
typedef struct FutbolV0EnvState {
    double   rows[5][5];   /* ai_1, ai_2, opp_1, opp_2, ball: x, y, tx, ty, speed (obs rows 0-4) */
    uint64_t t_total;      /* steps since creation = Philox step index (not cleared by reset) */
    int32_t  ep_step;      /* steps since the last reset (reference `time` = ep_step additions of 0.1) */
    int32_t  ai_score, opp_score;
    uint8_t  owner, last_owner;  /* BallOwner 0..4 (ballowner.py:3-7) */
    uint8_t  flags;        /* of the last step: 1 goal, 2 out-of-field fix, 4 done */
    uint8_t  pad_;
} FutbolV0EnvState;

/* One env's v1 state, for futbol_get_state / futbol_set_state: this header, followed by the arbiter cache of the
 * env's P = B(B-1)/2 + 12 B shape pairs (B = 2N + 1 bodies):  double jn[P]; uint32_t last[P];  then padding to a
 * multiple of 8 bytes.  futbol_env_state_bytes() is the size of one whole record.
 * Bodies: team A players 0..N-1, team B players N..2N-1, ball 2N; unused rows are zero.
 * Pair ids: circle/circle (i < j) -> j(j-1)/2 + i; circle/segment -> B(B-1)/2 + 12 body + segment (segments in the
 * order of _setup_walls, envs_v1/futbol_env.py:184-224).  jn = the normal impulse the pair accumulated the last
 * time it touched; last = the stamp of that space step (0 = never). */
typedef struct FutbolV1EnvState {
    double   body[21][6];  /* x, y, vx, vy, v_bias_x, v_bias_y */
    uint64_t t_total;      /* steps since creation = Philox step index */
    uint32_t stamp;        /* space steps taken (0.1 s steps and the 1e-4 s kick-off steps), counted from 8 */
    int32_t  ep_step;
    uint8_t  owner_side;   /* ball_owner_side: 0 left, 1 right (envs_v1/futbol_env.py:147) */
    uint8_t  flags;        /* of the last step: 1 goal, 2 out of bounds, 4 done, 8 the goal was scored by the left team */
    uint8_t  pad_[6];
} FutbolV1EnvState;

/* Rollout statistics (sums over all envs and steps of one futbol_rollout call).
 * v1: goals_ai = goals of the left team, goals_opp = of the right team, out_of_field = out-of-bounds fixes,
 * reserved[0] = contacts solved, reserved[1] = contacts dropped (more than 32 in one space step; expected 0). */
typedef struct FutbolStats {
    double   reward_sum;
    uint64_t env_steps, episodes, goals_ai, goals_opp, out_of_field;
    uint64_t reserved[2];
} FutbolStats;

/* ---- lifetime ---------------------------------------------------------------------- */
/* replaces FutbolEnv.__init__ (futbol_env.py:134-201) / Futbol.__init__ (envs_v1:63-127) */
int futbol_create(const FutbolConfig *cfg, FutbolHandle **out);
int futbol_destroy(FutbolHandle *h);
const char *futbol_last_error(void);
int futbol_abi_version(void);

/* ---- sizes ------------------------------------------------------------------------- */
size_t futbol_state_bytes(const FutbolHandle *h); /* bytes of the opaque SoA state buffer */
int futbol_obs_dim(const FutbolHandle *h);        /* v0: 30 (=6x5); v1: 4 + 8*n_players */
int futbol_act_dim(const FutbolHandle *h);        /* v0: 1 (Discrete(16)); v1: 2*n_players */
int futbol_draw_limit_steps(const FutbolHandle *h); /* v0: episode length in steps (401 at game_time 40) */

/* ---- reset: FutbolEnv.reset (futbol_env.py:205-245) / Futbol.reset (envs_v1:145-150) --
 * mask: NULL = all envs, else uint8[n_envs], non-zero = reset that env.
 * obs: NULL or [n_envs, obs_dim] in `obs_dtype` (0 = float32, 1 = float64). */
int futbol_reset(FutbolHandle *h, void *state, const uint8_t *mask, void *obs, int obs_dtype, void *stream);

/* ---- step: FutbolEnv.step (futbol_env.py:628-717) / Futbol.step (envs_v1:427-483) -----
 * actions: v0 uint8[n_envs] in 0..15 (ai_1 = a/4, ai_2 = a%4, :653);
 *          v1 uint8[n_envs, 2*n_players] (arrow, key) per left-team player.
 * obs: [n_envs, obs_dim]; reward: [n_envs]; both in `out_dtype` (0 = float32, 1 = float64).
 * done: uint8[n_envs].  final_obs: NULL or [n_envs, obs_dim]: with auto_reset, the terminal
 * observation of envs that finished in this call (other rows untouched). */
int futbol_step(FutbolHandle *h, void *state, const uint8_t *actions, void *obs, void *reward,
                uint8_t *done, void *final_obs, int out_dtype, void *stream);

/* ---- fused K-step rollout: the caller's `for t: env.step(a_t)` loop in one launch ------
 * State stays in registers across the K steps.  actions: uint8 [K, n_envs(, act_dim)] or NULL
 * (NULL = uniform random actions generated in-kernel from Philox stream 1 = synthetic load).
 * obs: float32 [K, n_envs, obs_dim]; reward: float32 [K, n_envs]; done: uint8 [K, n_envs];
 * any of the three may be NULL (not written).  stats: NULL or a device FutbolStats that the
 * call ACCUMULATES into.  auto_reset semantics as futbol_step. */
int futbol_rollout(FutbolHandle *h, void *state, int K, const uint8_t *actions, float *obs,
                   float *reward, uint8_t *done, FutbolStats *stats, void *stream);

/* ---- the same with the OPPONENTS' actions supplied by the caller (self-play / learned opponents) ----------
 * Replaces the reference's opponent sources -- v0: `randint(0, 15)` (futbol_env.py:639-645; the handle must have
 * been created with random_opp = 1, the draw is not taken); v1: `action_space.sample()` (envs_v1/futbol_env.py:429).
 * opp_actions: v0 uint8[n] / [K, n] in 0..15 (opp_1 = a / 4, opp_2 = a % 4); v1 uint8[n, 2N] / [K, n, 2N] for the
 * right team.  NULL = the reference's own opponents (then identical to futbol_step / futbol_rollout). */
int futbol_step_vs(FutbolHandle *h, void *state, const uint8_t *actions, const uint8_t *opp_actions, void *obs,
                   void *reward, uint8_t *done, void *final_obs, int out_dtype, void *stream);
int futbol_rollout_vs(FutbolHandle *h, void *state, int K, const uint8_t *actions, const uint8_t *opp_actions,
                      float *obs, float *reward, uint8_t *done, FutbolStats *stats, void *stream);

/* ---- state access (device AoS records: FutbolV0EnvState / FutbolV1EnvState) ---------- */
size_t futbol_env_state_bytes(const FutbolHandle *h); /* sizeof one AoS record */
int futbol_get_state(FutbolHandle *h, const void *state, void *aos_out, void *stream);
/* Values must be pitch-scale: finite, |value| <= 1e6 and either zero or >= 1e-60 (the kernel's correctly rounded
 * division / square root run without a range guard, csrc/ieee_fast.cuh); the Python binding validates this on the
 * host before the call. */
int futbol_set_state(FutbolHandle *h, void *state, const void *aos_in, void *stream);

/* ---- rollout-buffer glue: generalised advantage estimation -----------------------------
 * The consumer directly behind the path in the reference's flow (stable-baselines PPO2 runner,
 * colab_notebook.ipynb:852).  reward float32 [T, n], done uint8 [T, n] (done_t: the episode ended AT
 * step t, so nothing is bootstrapped across it), value float32 [T + 1, n] (V of obs_t; row T = V of the
 * observation after the last step); adv, ret float32 [T, n].  No handle: a pure function of its inputs. */
int futbol_gae(const float *reward, const uint8_t *done, const float *value, float gamma, float lam,
               float *adv, float *ret, int T, int n, void *stream);

/* ---- rollout-buffer glue: minibatch gather ----------------------------------------------------
 * Row idx[j] (int64, 0 <= idx[j] < rows) of every column of a flattened rollout buffer into row j of the minibatch, in
 * ONE pass over the index: obs float32 [rows, obs_dim], act uint8 [rows], c0..c3 float32 [rows] (e.g. old log-prob,
 * advantage, return, value).  Any column may be NULL together with its destination.  An out-of-range index is never
 * dereferenced: its observation row is zero-filled and, when `bad` is not NULL, the device counter *bad is incremented.
 * Replaces the per-minibatch slicing of stable-baselines' runner (colab_notebook.ipynb:852).  No handle. */
int futbol_gather_minibatch(const int64_t *idx, int64_t m, int64_t rows, const float *obs, int obs_dim, float *obs_out,
                            const uint8_t *act, uint8_t *act_out, const float *c0, float *c0_out, const float *c1, float *c1_out,
                            const float *c2, float *c2_out, const float *c3, float *c3_out, uint64_t *bad, void *stream);

/* ---- rollout-buffer glue: action sampling -------------------------------------------------------
 * One categorical draw per row of unnormalised log-probabilities, in one launch: what stable-baselines'
 * CategoricalProbabilityDistribution.sample() and .neglogp() compute between the policy's forward pass and env.step for
 * the reference's Discrete(16) action space (envs/futbol_env.py:143; runner: colab_notebook.ipynb:852).
 * logits [n, n_actions] row-major, float32 (logits_dtype 0) or bfloat16 (1), 1 <= n_actions <= 32, finite.
 * u = 24-bit uniform from Philox4x32-10 (key seed, counter (t, row), stream 4) with t = (*t_base if t_base else 0) + t_off:
 * t_base is a DEVICE counter, so a captured CUDA graph draws fresh numbers on every replay once the caller advances it.
 * action[i] = the first k with sum_{j<=k} exp(l_j - max l) > u * sum_j exp(l_j - max l); logp[i] (may be NULL) =
 * log-softmax(l)[action].  actions uint8 [n], logp float32 [n].  No handle. */
int futbol_sample_actions(const void *logits, int logits_dtype, int64_t n, int n_actions, uint64_t seed, const uint64_t *t_base,
                          uint64_t t_off, uint8_t *actions, float *logp, void *stream);

/* ---- self test ----------------------------------------------------------------------------
 * Compares the kernel's guard-free fp64 division / square-root sequences (csrc/ieee_fast.cuh) with the
 * compiler's __ddiv_rn / __dsqrt_rn on n operand pairs: a[i] / b[i], (a[i], |b[i]|) / b[i] through the
 * shared-reciprocal form, sqrt(|b[i]|); pairs with b[i] == 0 are skipped.  mismatch: device uint64[3]
 * (division, two-numerator division, square root), accumulated into. */
int futbol_selftest_arith(const double *a, const double *b, uint64_t *mismatch, size_t n, void *stream);

/* number of kernels this handle has launched (bench.py's gpu_launches) */
uint64_t futbol_launch_count(const FutbolHandle *h);

/* ---- launch tuning --------------------------------------------------------------------------
 * futbol_rollout on a batch of only a few waves of thread blocks (e.g. 131,072 envs: one of eight ranks of the 2^20 job)
 * cuts the K steps into time slices and lets a grid that just fills the GPU take (slice, env-block) units from a queue,
 * so that no SM idles through a partial last wave.  Results do not depend on the slicing.  slices: 0 = chosen per
 * launch from the batch size (default: sliced from one to ten waves of thread blocks), 1 = never slice, n > 1 = n equal
 * slices.  v0 and v1 (v1: units are warps of 32 envs).  No reference counterpart. */
int futbol_set_rollout_slices(FutbolHandle *h, int slices);
/* the number of time slices futbol_rollout will use for K steps on the current device (1 = the plain kernel) */
int futbol_rollout_slices(FutbolHandle *h, int K);
/* which kernel futbol_rollout will launch for K steps on the current device: 0 = standard (v0: 20 warps per SM), 1 = standard,
 * time-sliced, 2 = dense (v0 only: 28 warps per SM) */
int futbol_rollout_kernel(FutbolHandle *h, int K);
/* 0 / 1 = the standard kernel (default), 2 = the dense kernel: 7936 B of shared memory per warp and 72 registers, so that a
 * batch of a few waves fills whole waves; measured slower, kept selectable.  Results do not depend on it. */
int futbol_set_rollout_variant(FutbolHandle *h, int variant);

#ifdef __cplusplus
}
#endif
#endif /* FUTBOL_B200_H */
