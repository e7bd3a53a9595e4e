/*
 * CPU restatement of the v1 hot path: gym_futbol/envs_v1/futbol_env.py Futbol.step/reset with
 * team.py, player.py, ball.py, plus the subset of Chipmunk2D 7.0.x (pymunk 5.6.0, the version the
 * reference's authors ran: colab_notebook.ipynb:118,130) that those files drive.
 *
 * TEST INFRASTRUCTURE (see oracle/README.md).
 * PINNED, game logic: tests/golden/v1_golden.npz holds traces of the reference's OWN, UNMODIFIED Python
 * (envs_v1/futbol_env.py, team.py, player.py, ball.py) executed by oracle/ref_harness_v1.py with the Philox
 * streams below injected; tests/test_oracle_v1_golden.py holds this file to them bit for bit (arith = 1).
 * PARITY UNPINNED, physics: pymunk/Chipmunk is a third-party dependency that is neither vendored in
 * /root/reference nor installable here, so those traces ran over oracle/pymunk_standin.py, a second, independent
 * restatement of the same specification, not over the real library.  The physics below
 * restates Chipmunk's published algorithm (cpSpaceStep: position integration, circle/circle and circle/segment
 * narrow phase, arbiter pre-step, damped velocity integration, warm start, 10 sequential-impulse iterations)
 * from knowledge of its source.  External pins that are also checked (tests/test_oracle_v1.py): episode length
 * 300 (gym_futbol/envs_v1/2v2/logs/evaluations.npz), observation and action shapes (saved-model JSON), kick-off
 * formations (closed form of team.py:52-112), single-body closed forms (impulse -> delta v, damping 0.95^dt,
 * speed clamps), two-body restitution 0.04.
 *
 * What is specified here because Chipmunk leaves it implementation-defined:
 *   - body order: team A players 0..N-1, team B players N..2N-1, ball 2N (the order they are added to
 *     the space, futbol_env.py:102-125);
 *   - contact (arbiter) order = ascending pair id: circle/circle pair (i < j) -> j(j-1)/2 + i, with a = i,
 *     b = j; then circle/segment -> CC + 12*body + segment, segments in the order of _setup_walls
 *     (:184-224: six boundary segments, then six goal-box segments);
 *   - the rotational terms of k_scalar vanish (contact offsets are parallel to the normal), friction is
 *     0 (mu_a * mu_b, circles have friction 0), so bodies never spin and tangential impulses are 0;
 *   - cached impulses (cpArbiterUpdate / cpArbiterApplyCachedImpulse): a pair that touches again within the
 *     3-step persistence window inherits its accumulated normal impulse jnAcc (the contact hashes of circle
 *     pairs are all 0), but the warm-start impulse itself is applied only if the pair also touched in the
 *     PREVIOUS space step: a cached arbiter that separated for one or two steps comes back in state
 *     FIRST_COLLISION, for which cpArbiterApplyCachedImpulse returns early;
 *   - dt_coef of the warm start is 1: the only step with another dt is the 1e-4 step after a kick-off
 *     (:142); nothing touches in it (all bodies are teleported to the formation), so no pair is warm-started
 *     in it or in the step after it;
 *   - cpSpace defaults as cpSpace.c writes them, in C float literals: collision_slop = 0.1f,
 *     collision_bias = pow(1.0f - 0.1f, 60.0f) (both exact in double: 0x1.99999ap-4, 0x1.ccccccp-1);
 *   - a segment's closest point is taken by clamping the centre's coordinate (all segments are axis-aligned),
 *     and at most V1_MAX_CONTACTS = 32 contacts are solved per space step (pairs beyond that, in pair order,
 *     are ignored and counted in `overflow`; never reached in any test or bench run);
 *   - contact-point distance dist = (p2 - p1) . n with p1 = c_a + n r_a, p2 = c_b - n r_b (closest - n r_s
 *     for a segment), i.e. Chipmunk's (r2 - r1 + body_delta) . n without the round trip through r1, r2.
 * Arithmetic: one IEEE double operation per written operation, no contraction (-ffp-contract=off).  Two modes
 * (as futbol_v0_oracle.c): arith = 0 "kernel": x**2 is x*x everywhere -- what the CUDA kernel computes, so the
 * two agree bit for bit; arith = 1 "libm": the Python-level squares of the reference -- get_vec (:57),
 * _ball_to_team_distance_arr (:490) and pymunk's Vec2d.length in the speed clamps (player.py:47, ball.py:51) --
 * are libm pow(x, 2.0) as CPython's float ** computes them: bit-identical to the Python run on the same libm.
 * The squares inside Chipmunk (C: cpvlengthsq, cpvdot) are products in both modes.
 *
 * Randomness (specification: oracle/philox.py; the reference is unseeded):
 *   stream 2, step t: right-team actions, player p: arrow = word(2p)*5 >> 32, key = word(2p+1)*5 >> 32 (:429)
 *   stream 1, step t: synthetic left-team actions, same layout
 *   stream 3, step t: sequential draws j = 0, 1, ...: random.choices of a pass (team.py:141-178, one
 *             or two per pass), random.choice of the player who gets an out-of-bounds ball (:275,:278),
 *             random.choice of the side after a goal (:475); random() = (w>>8)*2^-24, choice among n =
 *             w*n >> 32, choices among c equally weighted = ((w>>8)*c) >> 24
 *   stream 3, block 0x4000 word 0 at the env's current t: the side drawn by reset() (:147)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <pthread.h>

void futbol_oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#define V1_MAX_N 10
#define V1_MAX_BODIES (2 * V1_MAX_N + 1)
#define V1_NSEG 12
#define V1_MAX_PAIRS (V1_MAX_BODIES * (V1_MAX_BODIES - 1) / 2 + V1_MAX_BODIES * V1_NSEG)
#define V1_MAX_CONTACTS 32

/* futbol_env.py:19-37 */
#define WIDTH 105.0
#define HEIGHT 68.0
#define GOAL_SIZE 20.0
#define TIME_STEP 0.1
#define BALL_MAX_VELOCITY 25.0
#define PLAYER_MAX_VELOCITY 10.0
#define BALL_WEIGHT 10.0
#define PLAYER_WEIGHT 20.0
#define PLAYER_FORCE_LIMIT 40.0
#define BALL_FORCE_LIMIT 120.0
#define R_PLAYER 1.5 /* player.py:7 */
#define R_BALL 1.0   /* ball.py:7 */
#define R_SEG 1.0    /* :187 */
#define ELASTICITY 0.2

enum { V1_FLAG_GOAL = 1, V1_FLAG_OUT = 2, V1_FLAG_DONE = 4, V1_FLAG_GOAL_LEFT = 8 };

typedef struct {
    uint64_t seed;
    int32_t n_players;     /* number_of_player, :65 */
    int32_t ep_limit;      /* first k with k additions of 0.1 > total_time (300 for 30) */
    double damping_dt;     /* pow(0.95, 0.1): space.damping ** dt, :99 */
    double bias_coef;      /* 1 - pow(pow(1.0f - 0.1f, 60), 0.1): Chipmunk collision_bias default */
    double slop;           /* 0.1f: Chipmunk collision_slop default */
    int32_t arith, pad_;   /* 1 = libm pow for the Python-level squares, 0 = kernel arithmetic (x*x) */
    double form_x[2 * V1_MAX_N], form_y[2 * V1_MAX_N];   /* kick-off formation, team.py:52-112 */
} OracleV1Config;

typedef struct {
    double p[V1_MAX_BODIES][2], v[V1_MAX_BODIES][2], vb[V1_MAX_BODIES][2];   /* position, velocity, bias velocity */
    double jn[V1_MAX_PAIRS];     /* normal impulse accumulated by the pair the last time it touched */
    uint8_t age[V1_MAX_PAIRS];   /* space steps since the pair last touched (255 = never) */
    uint64_t t_total;
    int32_t ep_step;
    int32_t owner_side;          /* 0 = left, 1 = right (ball_owner_side, :147) */
    uint32_t env_id;
    uint32_t step_draws;
    int32_t goals_left, goals_right;   /* bookkeeping for statistics (the reference keeps no score) */
    int32_t flags;
    int32_t contacts;            /* contacts of the last 0.1 space step (statistics) */
    int32_t overflow;            /* contacts dropped because more than V1_MAX_CONTACTS touched in one step */
    int32_t pad_;
} OracleV1Env;

size_t futbol_v1_oracle_env_bytes(void) { return sizeof(OracleV1Env); }
size_t futbol_v1_oracle_cfg_bytes(void) { return sizeof(OracleV1Config); }
int futbol_v1_oracle_present(void) { return 1; }

/* ---- segments, _setup_walls :182-224 ------------------------------------------------- */
static const double SEG[V1_NSEG][4] = {
    {0, 0, 0, HEIGHT / 2 - GOAL_SIZE / 2},
    {0, HEIGHT / 2 + GOAL_SIZE / 2, 0, HEIGHT},
    {0, HEIGHT, WIDTH, HEIGHT},
    {WIDTH, 0, WIDTH, HEIGHT / 2 - GOAL_SIZE / 2},
    {WIDTH, HEIGHT / 2 + GOAL_SIZE / 2, WIDTH, HEIGHT},
    {0, 0, WIDTH, 0},
    {-2, HEIGHT / 2 - GOAL_SIZE / 2, -2, HEIGHT / 2 + GOAL_SIZE / 2},
    {-2, HEIGHT / 2 - GOAL_SIZE / 2, 0, HEIGHT / 2 - GOAL_SIZE / 2},
    {-2, HEIGHT / 2 + GOAL_SIZE / 2, 0, HEIGHT / 2 + GOAL_SIZE / 2},
    {WIDTH + 2, HEIGHT / 2 - GOAL_SIZE / 2, WIDTH + 2, HEIGHT / 2 + GOAL_SIZE / 2},
    {WIDTH, HEIGHT / 2 - GOAL_SIZE / 2, WIDTH + 2, HEIGHT / 2 - GOAL_SIZE / 2},
    {WIDTH, HEIGHT / 2 + GOAL_SIZE / 2, WIDTH + 2, HEIGHT / 2 + GOAL_SIZE / 2},
};

/* ---- configuration -------------------------------------------------------------------- */
/* kick-off formation of one side, team.py:52-112 (Python float arithmetic, left to right) */
static void formation(int n, int right, double *xs, double *ys)
{
    double w = WIDTH, h = HEIGHT;
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

/* gcc folds pow(x, 2.0) into x*x; the volatile pointer (and -fno-builtin-pow) keeps the libm call */
static double (*volatile libm_pow_v1)(double, double) = pow;
static double sq(const OracleV1Config *c, double x) { return c->arith ? libm_pow_v1(x, 2.0) : x * x; }

int futbol_v1_oracle_config(OracleV1Config *cfg, uint64_t seed, int n_players, double total_time, int arith)
{
    if (n_players < 1 || n_players > V1_MAX_N) return -1;
    memset(cfg, 0, sizeof(*cfg));
    cfg->seed = seed;
    cfg->n_players = n_players;
    double t = 0.0;               /* :478-481: current_time += 0.1; done = current_time > total_time */
    int k = 0;
    do { t += TIME_STEP; ++k; } while (!(t > total_time) && k < (1 << 30));
    cfg->ep_limit = k;
    cfg->damping_dt = libm_pow_v1(0.95, TIME_STEP);
    cfg->slop = (double)0.1f;
    cfg->bias_coef = 1.0 - libm_pow_v1(libm_pow_v1((double)(1.0f - 0.1f), 60.0), TIME_STEP);
    cfg->arith = arith ? 1 : 0;
    formation(n_players, 0, cfg->form_x, cfg->form_y);
    formation(n_players, 1, cfg->form_x + n_players, cfg->form_y + n_players);
    return 0;
}

/* ---- RNG -------------------------------------------------------------------------------- */
static uint32_t word_at(uint64_t seed, uint32_t env_id, uint32_t stream, uint64_t t, uint32_t block, uint32_t w)
{
    uint32_t ctr[4] = { (uint32_t)t, ((uint32_t)(t >> 32) & 0xFFFFu) | (block << 16), env_id, stream };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t out[4];
    futbol_oracle_philox(ctr, key, out);
    return out[w];
}
static uint32_t draw(const OracleV1Config *c, OracleV1Env *e)
{
    uint32_t j = e->step_draws++;
    return word_at(c->seed, e->env_id, 3, e->t_total, j >> 2, j & 3);
}
/* MultiDiscrete([5,5]*N).sample() for one team from `stream` at this env's step t */
void futbol_v1_oracle_team_actions(uint64_t seed, uint32_t env_id, uint32_t stream, uint64_t t, int n, uint8_t *out)
{
    for (int j = 0; j < 2 * n; ++j)
        out[j] = (uint8_t)(((uint64_t)word_at(seed, env_id, stream, t, (uint32_t)j >> 2, (uint32_t)j & 3) * 5u) >> 32);
}

/* ---- kick-off / reset --------------------------------------------------------------------- */
static void age_pairs(OracleV1Env *e, int npairs)
{
    for (int q = 0; q < npairs; ++q) if (e->age[q] != 255) e->age[q] += 1;
}
static int n_pairs(int B) { return B * (B - 1) / 2 + B * V1_NSEG; }

/* _position_to_initial, :129-143: teleport to the formation, zero velocities, space.step(1e-4).
 * With all velocities zero the 1e-4 step only consumes the bias velocities (p += v_bias * 1e-4, v_bias = 0),
 * finds no contact (formation spacing >= 13.6) and ages the cached arbiters by one step. */
static void position_to_initial(const OracleV1Config *c, OracleV1Env *e)
{
    int N = c->n_players, B = 2 * N + 1;
    for (int i = 0; i < 2 * N; ++i) { e->p[i][0] = c->form_x[i]; e->p[i][1] = c->form_y[i]; e->v[i][0] = 0.0; e->v[i][1] = 0.0; }
    e->p[2 * N][0] = WIDTH * 0.5; e->p[2 * N][1] = HEIGHT * 0.5; e->v[2 * N][0] = 0.0; e->v[2 * N][1] = 0.0;
    for (int i = 0; i < B; ++i) {
        e->p[i][0] = e->p[i][0] + (0.0 + e->vb[i][0]) * 0.0001;
        e->p[i][1] = e->p[i][1] + (0.0 + e->vb[i][1]) * 0.0001;
        e->vb[i][0] = 0.0; e->vb[i][1] = 0.0;
    }
    age_pairs(e, n_pairs(B));
}

void futbol_v1_oracle_reset(const OracleV1Config *c, OracleV1Env *e)
{   /* Futbol.reset, :145-150 */
    e->ep_step = 0;
    e->owner_side = (int)(((uint64_t)word_at(c->seed, e->env_id, 3, e->t_total, 0x4000u, 0) * 2u) >> 32);
    position_to_initial(c, e);
}

void futbol_v1_oracle_init(const OracleV1Config *c, OracleV1Env *e, uint32_t env_id)
{   /* Futbol.__init__, :63-127: builds the space and calls reset() (:127) */
    memset(e, 0, sizeof(*e));
    e->env_id = env_id;
    memset(e->age, 255, sizeof(e->age));
    futbol_v1_oracle_reset(c, e);
}

/* ---- observation, :154-180 ------------------------------------------------------------------ */
void futbol_v1_oracle_obs(const OracleV1Config *c, const OracleV1Env *e, double *obs)
{
    int N = c->n_players, b = 2 * N;
    obs[0] = (e->p[b][0] - 52.5) / 52.5; obs[1] = (e->p[b][1] - 34.0) / 34.0;
    obs[2] = (e->v[b][0] - 0.0) / 25.0;  obs[3] = (e->v[b][1] - 0.0) / 25.0;
    for (int i = 0; i < 2 * N; ++i) {
        obs[4 + 4 * i + 0] = (e->p[i][0] - 52.5) / 55.5; obs[4 + 4 * i + 1] = (e->p[i][1] - 34.0) / 34.0;
        obs[4 + 4 * i + 2] = (e->v[i][0] - 0.0) / 10.0;  obs[4 + 4 * i + 3] = (e->v[i][1] - 0.0) / 10.0;
    }
}

/* ---- action processing, :309-422 ---------------------------------------------------------------- */
static int touching(const OracleV1Env *e, int p, int ball)
{   /* Ball.has_contact_with, ball.py:39-40 = Chipmunk CircleToCircle: |delta|^2 < (r1 + r2)^2 */
    double dx = e->p[p][0] - e->p[ball][0], dy = e->p[p][1] - e->p[ball][1];
    return dx * dx + dy * dy < (R_BALL + R_PLAYER) * (R_BALL + R_PLAYER);
}

/* Team.get_pass_target_teammate, team.py:136-180; returns the body index of the target */
static int pass_target(const OracleV1Config *c, OracleV1Env *e, int p, int arrow)
{
    int N = c->n_players, base = p < N ? 0 : N, k = p - base;
    if (N == 1) return p;                                        /* :137-138 */
    uint32_t w = draw(c, e);                                     /* :141-142: any other teammate */
    int r = (int)(((uint64_t)(w >> 8) * (uint64_t)(N - 1)) >> 24);
    int target = r < k ? r : r + 1;
    if (arrow != 0) {                                            /* :148-178 */
        int elig[V1_MAX_N], cnt = 0;
        for (int i = 0; i < N; ++i) {
            double mx = e->p[base + i][0] - e->p[p][0], my = e->p[base + i][1] - e->p[p][1];
            int ok = arrow == 1 ? my > 0 : (arrow == 2 ? mx > 0 : (arrow == 3 ? my < 0 : mx < 0));
            if (ok) elig[cnt++] = i;
        }
        if (cnt > 0) {
            uint32_t w2 = draw(c, e);
            target = elig[(int)(((uint64_t)(w2 >> 8) * (uint64_t)cnt) >> 24)];
        }
    }
    return base + target;
}

static void process_action(const OracleV1Config *c, OracleV1Env *e, int p, int arrow, int key)
{
    const int N = c->n_players, ball = 2 * N, side = p < N ? 0 : 1;
    const double m_inv_p = 1.0 / PLAYER_WEIGHT, m_inv_b = 1.0 / BALL_WEIGHT;
    double fx = 0, fy = 0;                                       /* :312-327 */
    if (arrow == 1) fy = 1; else if (arrow == 2) fx = 1; else if (arrow == 3) fy = -1; else if (arrow == 4) fx = -1;
    if (key == 0 || key == 1) {                                  /* noop :331-335, dash :338-341 */
        double f = key == 0 ? PLAYER_WEIGHT : PLAYER_FORCE_LIMIT;
        e->v[p][0] = e->v[p][0] + (f * fx) * m_inv_p;            /* apply_impulse_at_local_point: v += j * m_inv */
        e->v[p][1] = e->v[p][1] + (f * fy) * m_inv_p;
        if (touching(e, p, ball)) { e->v[ball][0] = e->v[p][0]; e->v[ball][1] = e->v[p][1]; }   /* :300-304 */
    } else if (key == 2 || key == 4) {                           /* shoot :344-366, pass :394-416 */
        if (touching(e, p, ball)) {
            double gx, gy, force, div;
            if (key == 2) { gx = side == 0 ? WIDTH : 0.0; gy = HEIGHT / 2; force = BALL_FORCE_LIMIT; div = 2.0; }
            else { int t = pass_target(c, e, p, arrow); gx = e->p[t][0]; gy = e->p[t][1]; force = BALL_FORCE_LIMIT - 20; div = 10.0; }
            double vx = gx - e->p[ball][0], vy = gy - e->p[ball][1];
            double mag = sqrt(sq(c, vx) + sq(c, vy));                /* get_vec, :55-58 */
            double bfx = force * vx / mag, bfy = force * vy / mag;
            e->v[ball][0] = e->v[ball][0] / div; e->v[ball][1] = e->v[ball][1] / div;
            e->owner_side = side;
            e->v[ball][0] = e->v[ball][0] + bfx * m_inv_b;
            e->v[ball][1] = e->v[ball][1] + bfy * m_inv_b;
        }
    } else if (key == 3) {                                       /* press :371-391 */
        if (!touching(e, p, ball) && arrow == 0) {
            double vx = e->p[ball][0] - e->p[p][0], vy = e->p[ball][1] - e->p[p][1];
            double mag = sqrt(sq(c, vx) + sq(c, vy));
            double pfx = PLAYER_FORCE_LIMIT * vx / mag, pfy = PLAYER_FORCE_LIMIT * vy / mag;
            e->v[p][0] = e->v[p][0] + pfx * m_inv_p;
            e->v[p][1] = e->v[p][1] + pfy * m_inv_p;
        }
    }
}

/* ---- Chipmunk subset ----------------------------------------------------------------------------- */
typedef struct { int a, b, q, warm; double nx, ny, n_mass, bias, bounce, jn, jbias; } Contact;

static double radius_of(int i, int ball) { return i == ball ? R_BALL : R_PLAYER; }
static double minv_of(int i, int ball) { return i == ball ? 1.0 / BALL_WEIGHT : 1.0 / PLAYER_WEIGHT; }

/* closest point of segment s to (cx, cy).  Chipmunk's CircleToSegment computes a + (b - a) * clamp01(((b - a) .
 * (c - a)) / |b - a|^2); every segment of this scene is axis-aligned (:184-224), for which that point is the
 * centre's coordinate clamped to the segment's extent -- the form used here (and by the kernel). */
static void seg_closest(int s, double cx, double cy, double *qx, double *qy)
{
    double ax = SEG[s][0], ay = SEG[s][1], bx = SEG[s][2], by = SEG[s][3];
    if (ax == bx) { *qx = ax; *qy = cy < ay ? ay : (cy > by ? by : cy); }     /* vertical, ay < by */
    else          { *qy = ay; *qx = cx < ax ? ax : (cx > bx ? bx : cx); }     /* horizontal, ax < bx */
}

static int ball_touches_segment(const OracleV1Env *e, int ball, int s)
{
    double qx, qy;
    seg_closest(s, e->p[ball][0], e->p[ball][1], &qx, &qy);
    double dx = qx - e->p[ball][0], dy = qy - e->p[ball][1];
    return dx * dx + dy * dy < (R_BALL + R_SEG) * (R_BALL + R_SEG);
}

/* cpSpaceStep(dt = 0.1) */
static void space_step(const OracleV1Config *c, OracleV1Env *e)
{
    const int N = c->n_players, B = 2 * N + 1, ball = 2 * N, CC = B * (B - 1) / 2;
    const double dt = TIME_STEP, slop = c->slop;
    Contact con[V1_MAX_CONTACTS];
    int nc = 0;
    /* 1. integrate positions (cpBodyUpdatePosition): p += (v + v_bias) dt; v_bias = 0 */
    for (int i = 0; i < B; ++i) {
        e->p[i][0] = e->p[i][0] + (e->v[i][0] + e->vb[i][0]) * dt;
        e->p[i][1] = e->p[i][1] + (e->v[i][1] + e->vb[i][1]) * dt;
        e->vb[i][0] = 0.0; e->vb[i][1] = 0.0;
    }
    /* 2. narrow phase in pair-id order + 5. arbiter pre-step (uses the velocities before step 6) */
    for (int q = 0; q < n_pairs(B); ++q) {
        int a, b;                                                 /* b < 0: static segment -1 - b */
        if (q < CC) { int j = 1; while (j * (j + 1) / 2 <= q) ++j; a = q - j * (j - 1) / 2; b = j; }
        else { a = (q - CC) / V1_NSEG; b = -1 - (q - CC) % V1_NSEG; }
        double ra = radius_of(a, ball), rb, tx, ty;              /* (tx, ty): centre of b or closest point */
        if (b >= 0) { rb = radius_of(b, ball); tx = e->p[b][0]; ty = e->p[b][1]; }
        else { rb = R_SEG; seg_closest(-1 - b, e->p[a][0], e->p[a][1], &tx, &ty); }
        double dx = tx - e->p[a][0], dy = ty - e->p[a][1];
        double distsq = dx * dx + dy * dy, mind = ra + rb;
        int touch = distsq < mind * mind;
        if (touch && nc == V1_MAX_CONTACTS) { e->overflow += 1; touch = 0; }
        if (!touch) { if (e->age[q] != 255) e->age[q] += 1; continue; }
        Contact *k = &con[nc++];
        double dist = sqrt(distsq);
        if (dist != 0.0) { double inv = 1.0 / dist; k->nx = dx * inv; k->ny = dy * inv; }
        else if (b >= 0) { k->nx = 1.0; k->ny = 0.0; }
        else {   /* segment normal: perp(normalize(b - a)) */
            double sx = SEG[-1 - b][2] - SEG[-1 - b][0], sy = SEG[-1 - b][3] - SEG[-1 - b][1], sl = sqrt(sx * sx + sy * sy);
            k->nx = -(sy / sl); k->ny = sx / sl;
        }
        k->a = a; k->b = b; k->q = q;
        double p1x = e->p[a][0] + k->nx * ra, p1y = e->p[a][1] + k->ny * ra;
        double p2x = tx + k->nx * (-rb), p2y = ty + k->ny * (-rb);
        double pen = (p2x - p1x) * k->nx + (p2y - p1y) * k->ny;
        double ma = minv_of(a, ball), mb = b >= 0 ? minv_of(b, ball) : 0.0;
        k->n_mass = 1.0 / (ma + mb);
        double m = pen + slop; m = m < 0.0 ? m : 0.0;            /* cpfmin(0, dist + slop) */
        k->bias = -c->bias_coef * m / dt;
        k->jbias = 0.0;
        double vbx = b >= 0 ? e->v[b][0] : 0.0, vby = b >= 0 ? e->v[b][1] : 0.0;
        double el = b >= 0 ? ELASTICITY * ELASTICITY : ELASTICITY * 0.0;
        k->bounce = ((vbx - e->v[a][0]) * k->nx + (vby - e->v[a][1]) * k->ny) * el;
        k->jn = e->age[q] <= 2 ? e->jn[q] : 0.0;                 /* cached arbiter: collision_persistence = 3 */
        k->warm = e->age[q] == 0;                                /* state NORMAL: the pair touched in the previous step too */
        e->age[q] = 0;
    }
    e->contacts = nc;
    /* 6. integrate velocities through velocity_func (player.py:45-50, ball.py:49-54) */
    for (int i = 0; i < B; ++i) {
        double vx = e->v[i][0] * c->damping_dt + 0.0, vy = e->v[i][1] * c->damping_dt + 0.0;
        double l = sqrt(sq(c, vx) + sq(c, vy)), mx = i == ball ? BALL_MAX_VELOCITY : PLAYER_MAX_VELOCITY;
        if (l > mx) { double sc = mx / l; vx = vx * sc; vy = vy * sc; }
        e->v[i][0] = vx; e->v[i][1] = vy;
    }
    /* 7. warm start (cpArbiterApplyCachedImpulse, dt_coef = 1; returns early for first-contact arbiters) */
    for (int i = 0; i < nc; ++i) {
        Contact *k = &con[i];
        if (!k->warm) continue;
        double jx = k->nx * k->jn, jy = k->ny * k->jn, ma = minv_of(k->a, ball);
        e->v[k->a][0] = e->v[k->a][0] - jx * ma; e->v[k->a][1] = e->v[k->a][1] - jy * ma;
        if (k->b >= 0) { double mb = minv_of(k->b, ball); e->v[k->b][0] = e->v[k->b][0] + jx * mb; e->v[k->b][1] = e->v[k->b][1] + jy * mb; }
    }
    /* 8. ten iterations of cpArbiterApplyImpulse over the contacts in order */
    for (int it = 0; it < 10; ++it) {
        for (int i = 0; i < nc; ++i) {
            Contact *k = &con[i];
            int a = k->a, b = k->b;
            double ma = minv_of(a, ball), mb = b >= 0 ? minv_of(b, ball) : 0.0;
            double vb2x = b >= 0 ? e->vb[b][0] : 0.0, vb2y = b >= 0 ? e->vb[b][1] : 0.0;
            double v2x = b >= 0 ? e->v[b][0] : 0.0, v2y = b >= 0 ? e->v[b][1] : 0.0;
            double vbn = (vb2x - e->vb[a][0]) * k->nx + (vb2y - e->vb[a][1]) * k->ny;
            double vrn = (v2x - e->v[a][0]) * k->nx + (v2y - e->v[a][1]) * k->ny;
            double jbn = (k->bias - vbn) * k->n_mass, jbn_old = k->jbias;
            double t1 = jbn_old + jbn;
            k->jbias = t1 > 0.0 ? t1 : 0.0;
            double jn = -(k->bounce + vrn) * k->n_mass, jn_old = k->jn;
            double t2 = jn_old + jn;
            k->jn = t2 > 0.0 ? t2 : 0.0;
            double db = k->jbias - jbn_old, dj = k->jn - jn_old;
            double bx = k->nx * db, by = k->ny * db, jx = k->nx * dj, jy = k->ny * dj;
            e->vb[a][0] = e->vb[a][0] - bx * ma; e->vb[a][1] = e->vb[a][1] - by * ma;
            e->v[a][0] = e->v[a][0] - jx * ma;   e->v[a][1] = e->v[a][1] - jy * ma;
            if (b >= 0) {
                e->vb[b][0] = e->vb[b][0] + bx * mb; e->vb[b][1] = e->vb[b][1] + by * mb;
                e->v[b][0] = e->v[b][0] + jx * mb;   e->v[b][1] = e->v[b][1] + jy * mb;
            }
        }
    }
    for (int i = 0; i < nc; ++i) e->jn[con[i].q] = con[i].jn;
}

/* ---- Futbol.step, :427-483 ------------------------------------------------------------------------ */
/* right_actions: NULL = action_space.sample() (:429, Philox stream 2); else 2N bytes supplied by the caller for the
 * right team (the self-play hook). */
int futbol_v1_oracle_step_vs(const OracleV1Config *c, OracleV1Env *e, const uint8_t *left_actions, const uint8_t *right_actions,
                             double *reward_out)
{
    const int N = c->n_players, ball = 2 * N;
    uint8_t right[2 * V1_MAX_N];
    if (right_actions) memcpy(right, right_actions, (size_t)(2 * N));
    else futbol_v1_oracle_team_actions(c->seed, e->env_id, 2, e->t_total, N, right);   /* :429 */
    e->step_draws = 0;
    e->flags = 0;
    double init_d[V1_MAX_N];                                      /* :433 */
    for (int i = 0; i < N; ++i) { double dx = e->p[i][0] - e->p[ball][0], dy = e->p[i][1] - e->p[ball][1]; init_d[i] = sqrt(sq(c, dx) + sq(c, dy)); }
    double bix = e->p[ball][0], biy = e->p[ball][1];              /* :435 */
    double reward = 0.0;

    for (int p = 0; p < 2 * N; ++p) {                             /* :447-453 */
        const uint8_t *act = p < N ? left_actions + 2 * p : right + 2 * (p - N);
        process_action(c, e, p, act[0] % 5, act[1] % 5);
        if (touching(e, p, ball)) e->owner_side = p < N ? 0 : 1;
    }

    int out = 0;                                                  /* check_and_fix_out_bounds, :256-287 */
    for (int s = 0; s < 6 && !out; ++s) {
        if (!ball_touches_segment(e, ball, s)) continue;
        out = 1;
        double bx = e->p[ball][0], by = e->p[ball][1], dbx = 0, dby = 0, dpx = 0, dpy = 0;
        if (s == 0 || s == 1) { dbx = 3.5; dpx = 1; } else if (s == 3 || s == 4) { dbx = -3.5; dpx = -1; }
        else if (s == 2) { dby = -3.5; dpy = -1; } else { dby = 3.5; dpy = 1; }
        e->p[ball][0] = bx + dbx; e->p[ball][1] = by + dby; e->v[ball][0] = 0.0; e->v[ball][1] = 0.0;
        int pick = (int)(((uint64_t)draw(c, e) * (uint64_t)N) >> 32);
        int g = e->owner_side == 1 ? pick : N + pick;             /* the other side gets the ball */
        e->owner_side = e->owner_side == 1 ? 0 : 1;
        e->p[g][0] = bx + dpx; e->p[g][1] = by + dpy; e->v[g][0] = 0.0; e->v[g][1] = 0.0;
    }
    if (out) e->flags |= V1_FLAG_OUT;

    space_step(c, e);                                             /* :459 */

    if (!out) {                                                   /* :463-467 */
        double best = 0.0;
        int first = 1;
        for (int i = (N == 5 ? 3 : 0); i < N; ++i) {              /* :501-504 */
            double dx = e->p[i][0] - e->p[ball][0], dy = e->p[i][1] - e->p[ball][1];
            double diff = init_d[i] - sqrt(sq(c, dx) + sq(c, dy));
            if (first || diff > best) { best = diff; first = 0; }
        }
        reward = reward + best * 10;
        double ax = e->p[ball][0] - WIDTH, ay = e->p[ball][1] - HEIGHT / 2, ix = bix - WIDTH, iy = biy - HEIGHT / 2;
        reward = reward + (sqrt(sq(c, ix) + sq(c, iy)) - sqrt(sq(c, ax) + sq(c, ay))) * 10;
    }

    int goal = 0;                                                 /* ball_contact_goal, :291-296 */
    for (int s = 6; s < 12; ++s) goal = goal || ball_touches_segment(e, ball, s);
    if (goal) {                                                   /* :469-475 */
        int left_scored = e->p[ball][0] > WIDTH - 2;
        reward = reward + (left_scored ? 1000.0 : -1000.0);
        if (left_scored) e->goals_left += 1; else e->goals_right += 1;
        position_to_initial(c, e);
        e->owner_side = (int)(((uint64_t)draw(c, e) * 2u) >> 32);
        e->flags |= V1_FLAG_GOAL | (left_scored ? V1_FLAG_GOAL_LEFT : 0);
    }
    e->ep_step += 1;                                              /* :478-481 */
    e->t_total += 1;
    int done = e->ep_step >= c->ep_limit;
    if (done) e->flags |= V1_FLAG_DONE;
    *reward_out = reward;
    return done;
}

int futbol_v1_oracle_step(const OracleV1Config *c, OracleV1Env *e, const uint8_t *left_actions, double *reward_out)
{
    return futbol_v1_oracle_step_vs(c, e, left_actions, NULL, reward_out);
}

/* ---- batched rollout (threads over envs) -------------------------------------------------------------
 * actions: NULL = synthetic left actions from stream 1, else uint8 [steps][n][2N].
 * autoreset: 0 none; 2 VecEnv semantics (reset in the same step, obs slot holds the reset observation). */
typedef struct {
    const OracleV1Config *cfg; OracleV1Env *envs; int n, steps, lo, hi, autoreset; const uint8_t *actions, *right_actions;
    double *obs, *reward; uint8_t *done, *flags;
} Job;

static void *job_main(void *arg)
{
    Job *j = (Job *)arg;
    const int N = j->cfg->n_players, D = 4 + 8 * N;
    for (int i = j->lo; i < j->hi; ++i) {
        OracleV1Env *e = &j->envs[i];
        for (int k = 0; k < j->steps; ++k) {
            size_t slot = (size_t)k * j->n + i;
            uint8_t synth[2 * V1_MAX_N];
            const uint8_t *act;
            if (j->actions) act = j->actions + slot * 2 * N;
            else { futbol_v1_oracle_team_actions(j->cfg->seed, e->env_id, 1, e->t_total, N, synth); act = synth; }
            double r;
            int d = futbol_v1_oracle_step_vs(j->cfg, e, act, j->right_actions ? j->right_actions + slot * 2 * N : NULL, &r);
            int fl = e->flags;
            if (d && j->autoreset) futbol_v1_oracle_reset(j->cfg, e);
            if (j->obs) futbol_v1_oracle_obs(j->cfg, e, j->obs + slot * D);
            if (j->reward) j->reward[slot] = r;
            if (j->done) j->done[slot] = (uint8_t)d;
            if (j->flags) j->flags[slot] = (uint8_t)fl;
        }
    }
    return NULL;
}

void futbol_v1_oracle_rollout_vs(const OracleV1Config *cfg, OracleV1Env *envs, int n, int steps, const uint8_t *actions,
                                 const uint8_t *right_actions, int autoreset, int n_threads, double *obs, double *reward,
                                 uint8_t *done, uint8_t *flags)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n) n_threads = n;
    pthread_t th[256];
    Job jobs[256];
    if (n_threads > 256) n_threads = 256;
    for (int t = 0; t < n_threads; ++t) {
        jobs[t] = (Job){ cfg, envs, n, steps, (int)((long long)n * t / n_threads), (int)((long long)n * (t + 1) / n_threads),
                         autoreset, actions, right_actions, obs, reward, done, flags };
        if (n_threads == 1) job_main(&jobs[t]); else pthread_create(&th[t], NULL, job_main, &jobs[t]);
    }
    if (n_threads > 1) for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
}

void futbol_v1_oracle_rollout(const OracleV1Config *cfg, OracleV1Env *envs, int n, int steps, const uint8_t *actions,
                              int autoreset, int n_threads, double *obs, double *reward, uint8_t *done, uint8_t *flags)
{
    futbol_v1_oracle_rollout_vs(cfg, envs, n, steps, actions, NULL, autoreset, n_threads, obs, reward, done, flags);
}
