/*
 * CPU restatement of the reference v0 environment step -- TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for gym_futbol_b200's CUDA path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * build, load or call it.  The product (gym_futbol_b200/) never links or calls it.
 *
 * It restates, function by function, yc2454/gym-futbol:
 *   gym_futbol/envs/futbol_env.py   (FutbolEnv: reset :205-245, step :628-717, ...)
 *   gym_futbol/envs/easy_agent.py   (Easy_Agent.get_action_type :53-98)
 * in plain C, in the reference's own order of floating-point operations, including the
 * behaviours listed as quirks Q1..Q12 in SURVEY.md section 8a.  Each function cites the
 * lines it follows.  Parity status: PINNED -- tests/test_oracle_v0.py checks this file
 * against golden traces produced by running the unmodified reference
 * (tests/golden/make_golden_v0.py, oracle/ref_harness.py) and against the RNG-free
 * known-answer trace of SURVEY.md Appendix A.
 *
 * Randomness: the reference uses unseeded global generators; parity is defined on the
 * injected Philox4x32-10 stream specified in oracle/philox.py (same maps here).
 *
 * Arithmetic modes (``arith``).  The reference's floats depend on the host libm in two places:
 * numpy-scalar ``x**2`` (futbol_env.py:64,562; easy_agent.py:12) is libm ``pow(x, 2.0)``, which is
 * not correctly rounded (differs from x*x for ~0.09 % of inputs), and math.sin/cos/log
 * (:108-110, Box-Muller in the injected normal()).  No GPU can reproduce glibc's last-bit choices,
 * and the game has structural knife edges (e.g. an opponent chasing its own shot closes on the ball
 * by exactly 0.1 per step and is then tested with ``distance <= 1``), so last-bit differences do
 * flip integer outcomes now and then.  Therefore two modes:
 *   arith = 1 "libm":   pow(x,2.0) + libm sin/cos/log -- bit-identical to the Python reference on
 *                       the same glibc; this is the mode the golden vectors pin.
 *   arith = 0 "kernel": x*x + the fully specified fm_log / fm_sincos below (fdlibm-style polynomials,
 *                       only IEEE add/mul/div, no FMA) -- the arithmetic the CUDA kernel implements, so
 *                       kernel and oracle can be compared BIT-exactly at any scale.
 * tests/test_oracle_v0.py quantifies libm-vs-kernel mode: identical integers except where a compare was
 * decided by a last-bit difference, floats ~1e-13 otherwise.
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -fno-builtin-pow -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <pthread.h>

/* ---- constants: futbol_env.py:18-58 ---------------------------------------------- */
#define FIELD_LEN 105.0
#define FIELD_WID 68.0
#define GOAL_UPPER 39.0 /* FIELD_WID/2 + GOAL_SIZE/2  (:23) */
#define GOAL_LOWER 29.0 /* :24 */
#define SHOOT_SPEED 20.0
#define GOAL_REWARD 1000.0
#define PLAYER_ADV_REWARD_BASE 0.2
#define OUT_OF_FIELD_PENALTY (-0.6)
#define BAD_ACTION_PENALTY (-0.5)
#define BALL_CONTROL 0.3
#define STEP_SIZE 0.1
#define NORMAL_MISS 10.0
#define UNDER_DEFENCE_MISS 20.0
#define MAX_INTERCEPT_PROB 0.9
#define MAX_INTERCEPT_DIST 2.0
#define MIN_INTERCEPT_DIST 1.0
#define DECELERATION 0.0

/* action.py:3-6, ballowner.py:3-7 */
enum { A_RUN = 0, A_INTERCEPT = 1, A_SHOOT = 2, A_ASSIST = 3 };
enum { AI_1 = 0, AI_2 = 1, OPP_1 = 2, OPP_2 = 3, NOONE = 4, BALL = 4, OWNER_ROW = 5 };

typedef struct {
    uint64_t seed;
    int32_t random_opp;       /* futbol_env.py:138 */
    int32_t one_goal_end;     /* :137 */
    int32_t only_reward_goal; /* :137 */
    int32_t arith;            /* 1 = libm (pow, sin, cos, log), 0 = kernel arithmetic (x*x, fm_*) */
    int32_t rng_const;        /* 1 = constant RNG of SURVEY.md Appendix A (randint->a, random->0.5,
                                 uniform->(a+b)/2, normal->0); draws are still counted */
    int32_t pad_;
    double game_time;         /* :135 (GAME_TIME = 40) */
    double player_speed;      /* :135 (12) */
    double shoot_speed;       /* :136 (20) */
} OracleV0Config;

typedef struct {
    double obs[6][5];    /* rows ai_1, ai_2, opp_1, opp_2, ball, owner one-hot*10 */
    double kick[4][2];   /* frozen kickoff views of the four Easy_Agents (Q1) */
    double time;         /* float accumulator, :239,:716 */
    uint64_t draw_ctr;   /* draws consumed since construction (bookkeeping only) */
    uint64_t t_total;    /* total steps taken = Philox step index; not cleared by reset */
    uint32_t step_draws; /* draws consumed inside the current step (index j of the next draw) */
    uint32_t normal_calls; /* normal() calls inside the current step */
    uint32_t env_id;     /* global env id */
    int32_t owner, last_owner;
    int32_t ai_score, opp_score;
    int32_t flags;       /* set by the last step: 1 = goal, 2 = out-of-field fix */
} OracleV0Env;

/* ---- Philox4x32-10 (specification: oracle/philox.py) ------------------------------ */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void futbol_oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    philox4x32_10(ctr, key, out);
}

/* ---- "kernel arithmetic": fully specified elementary functions ----------------------------------
 * Domain needed: log on (0, 1] (Box-Muller), sin/cos on |x| <= 2*pi (Box-Muller angle and the kick
 * swing angle, bounded by 5.77 sigma <= 289 degrees).  fdlibm-style kernels (e_log.c, k_sin.c, k_cos.c
 * polynomials; two-term Cody-Waite reduction by pi/2), every operation an individually rounded IEEE
 * double op in exactly this order.  Max error vs libm: 1 ulp (log), 2.3e-16 abs (sin/cos). */
static double fm_log(double x)
{
    static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
        Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
        Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
        Lg7 = 1.479819860511658591e-01;
    uint64_t b;
    memcpy(&b, &x, 8);
    int k = (int)((b >> 52) & 0x7ff) - 1023;
    b = (b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double m;
    memcpy(&m, &b, 8);
    if (m > 1.4142135623730951) { m = m * 0.5; k += 1; }
    double f = m - 1.0;
    double s = f / (2.0 + f);
    double z = s * s, w = z * z;
    double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    double R = t2 + t1;
    double hfsq = (0.5 * f) * f;
    double dk = (double)k;
    return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
}
static double fm_ksin(double x)
{
    static const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
        S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08,
        S6 = 1.58969099521155010221e-10;
    double z = x * x, v = z * x;
    double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    return x + v * (S1 + z * r);
}
static double fm_kcos(double x)
{
    static const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
        C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09,
        C6 = -1.13596475577881948265e-11;
    double z = x * x;
    double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
    return (1.0 - 0.5 * z) + z * r;
}
static void fm_sincos(double x, double *s, double *c)
{
    static const double invpio2 = 6.36619772367581382433e-01, pio2_1 = 1.57079632673412561417e+00,
        pio2_1t = 6.07710050650619224932e-11;
    double fn = rint(x * invpio2);
    int n = (int)fn;
    double r = (x - fn * pio2_1) - fn * pio2_1t;
    double sr = fm_ksin(r), cr = fm_kcos(r);
    switch (n & 3) {
    case 0: *s = sr; *c = cr; break;
    case 1: *s = cr; *c = -sr; break;
    case 2: *s = -sr; *c = -cr; break;
    default: *s = -cr; *c = sr; break;
    }
}
void futbol_oracle_fm(double x, double out[3]) { out[0] = x > 0 ? fm_log(x) : 0.0; fm_sincos(x, &out[1], &out[2]); }

typedef struct { const OracleV0Config *cfg; OracleV0Env *e; } Ctx;

/* word j (per-step draw index) of step t -- counter layout: oracle/philox.py */
static uint32_t draw_word(uint64_t seed, uint32_t env_id, uint32_t stream, uint64_t t, uint32_t j)
{
    uint32_t blk = j >> 2;
    uint32_t ctr[4] = { (uint32_t)t, ((uint32_t)(t >> 32) & 0xFFFFu) | (blk << 16), env_id, stream };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t out[4];
    philox4x32_10(ctr, key, out);
    return out[j & 3];
}

static uint32_t next_u32(Ctx *c)
{
    c->e->draw_ctr++;
    return draw_word(c->cfg->seed, c->e->env_id, 0, c->e->t_total, c->e->step_draws++);
}
static double rng_random(Ctx *c)
{
    uint32_t w = next_u32(c);
    return c->cfg->rng_const ? 0.5 : (double)(w >> 8) * (1.0 / 16777216.0);
}
static int rng_randint(Ctx *c, int a, int b)
{
    uint32_t w = next_u32(c);
    return c->cfg->rng_const ? a : a + (int)(((uint64_t)w * (uint64_t)(b - a + 1)) >> 32);
}
static double rng_uniform(Ctx *c, double a, double b)
{
    if (c->cfg->rng_const) { next_u32(c); return (a + b) / 2; }
    return a + (b - a) * rng_random(c);
}
/* normal(mu, sd, 10): slot k comes from block 0x8000 + 8*call + (k >> 1), words 2(k&1), 2(k&1)+1;
 * no sequential draws are consumed (specification: oracle/philox.py) */
static void rng_normal10(Ctx *c, double mu, double sd, double out[10])
{
    const uint32_t base = 0x8000u + 8u * c->e->normal_calls++;
    if (c->cfg->rng_const) { for (int j = 0; j < 10; ++j) out[j] = 0.0; return; }
    for (int k = 0; k < 10; ++k) {
        uint32_t w0 = draw_word(c->cfg->seed, c->e->env_id, 0, c->e->t_total, 4u * (base + (uint32_t)(k >> 1)) + 2u * (k & 1));
        uint32_t w1 = draw_word(c->cfg->seed, c->e->env_id, 0, c->e->t_total, 4u * (base + (uint32_t)(k >> 1)) + 2u * (k & 1) + 1u);
        double u1 = (double)((w0 >> 8) + 1u) * (1.0 / 16777216.0);
        double u2 = (double)(w1 >> 8) * (1.0 / 16777216.0);
        double z;
        if (c->cfg->arith) {
            z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
        } else {
            double sn, cs;
            fm_sincos(6.283185307179586 * u2, &sn, &cs);
            z = sqrt(-2.0 * fm_log(u1)) * cs;
        }
        out[k] = mu + sd * z;
    }
}

int futbol_oracle_action(uint64_t seed, uint32_t env_id, uint64_t t, int n_actions)
{
    uint32_t ctr[4] = { (uint32_t)t, (uint32_t)(t >> 32) & 0xFFFFu, env_id, 1u };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t out[4];
    philox4x32_10(ctr, key, out);
    return (int)(((uint64_t)out[0] * (uint64_t)n_actions) >> 32);
}

/* ---- helpers: futbol_env.py:62-129 -------------------------------------------------- */
/* gcc folds pow(x, 2.0) into x*x; the volatile pointer (and -fno-builtin-pow) keeps the libm call */
static double (*volatile libm_pow)(double, double) = pow;
static double sq(const Ctx *c, double x) { return c->cfg->arith ? libm_pow(x, 2.0) : x * x; }
double futbol_oracle_libm_sq(double x) { return libm_pow(x, 2.0); }

/* get_vec, :62-65 (duplicate easy_agent.py:10-13) */
static double get_vec(const Ctx *c, const double *t, const double *o, double v[2])
{
    v[0] = t[0] - o[0];
    v[1] = t[1] - o[1];
    return sqrt(sq(c, v[0]) + sq(c, v[1]));
}

/* lock_in, :68-74 */
static double lock_in(double val, double max) { return val < 0 ? 0.0 : (val > max ? max : val); }

/* bigger_than, :76-82 */
static int bigger_than(double x1, double x2, double v)
{
    if (x1 <= v && x2 <= v) return 2;
    if (x1 > v && x2 > v) return 0;
    return 1;
}

/* screw_vec, :101-116 */
static void screw_vec(Ctx *c, const double vec[2], double mag, double accuracy, double out[2])
{
    double nd[10];
    rng_normal10(c, 0.0, accuracy, nd);                 /* :103 */
    double cs = vec[0] * 1.0 / mag;                     /* :105 */
    double sn = vec[1] * 1.0 / mag;                     /* :106 */
    int seed = rng_randint(c, 0, 9);                    /* :107 */
    double swing = (nd[seed] / 180) * 3.141592653589793; /* :108 */
    double ssin, scos;                                  /* :109-110 */
    if (c->cfg->arith) { ssin = sin(swing); scos = cos(swing); }
    else fm_sincos(swing, &ssin, &scos);
    double tcos = (cs * scos) - (sn * ssin);            /* :113 */
    double tsin = (sn * scos) + (cs * ssin);            /* :114 */
    out[0] = tcos * mag;                                /* :115 */
    out[1] = tsin * mag;
}

/* intercept_chance, :122-129 */
static double intercept_chance(double d, double d1, double d2)
{
    if (d < d1) return MAX_INTERCEPT_PROB;
    if (d >= d1 && d <= d2) {
        double k = MAX_INTERCEPT_PROB / (d1 - d2);
        return k * (d - d2);
    }
    return 0.0;
}

/* ---- reset, :205-245 ------------------------------------------------------------------ */
static void kickoff_rows(double obs[6][5])
{
    memset(obs, 0, sizeof(double) * 30);
    obs[BALL][0] = FIELD_LEN / 2;      obs[BALL][1] = FIELD_WID / 2;       /* :211 */
    obs[AI_1][0] = FIELD_LEN / 2 - 9;  obs[AI_1][1] = FIELD_WID / 2 + 5;   /* :213 */
    obs[AI_2][0] = FIELD_LEN / 2 - 9;  obs[AI_2][1] = FIELD_WID / 2 - 5;   /* :215 */
    obs[OPP_1][0] = FIELD_LEN / 2 + 9; obs[OPP_1][1] = FIELD_WID / 2 + 5;  /* :217 */
    obs[OPP_2][0] = FIELD_LEN / 2 + 9; obs[OPP_2][1] = FIELD_WID / 2 - 5;  /* :219 */
    /* owner row stays all zeros (:223) until the first ball_owner_array_update */
}

void futbol_v0_oracle_reset(OracleV0Env *e)
{
    kickoff_rows(e->obs);
    for (int p = 0; p < 4; ++p) { e->kick[p][0] = e->obs[p][0]; e->kick[p][1] = e->obs[p][1]; }
    e->owner = NOONE;      /* :235 */
    e->last_owner = NOONE; /* :236 */
    e->time = 0;           /* :239 */
    e->ai_score = 0;       /* :242-243 */
    e->opp_score = 0;
    e->flags = 0;
}

void futbol_v0_oracle_init(OracleV0Env *e, uint32_t env_id)
{
    memset(e, 0, sizeof(*e));
    e->env_id = env_id;
    futbol_v0_oracle_reset(e);
}

/* defence_near, :280-289.  Q1: agent.agent_observation is the view bound at construction
 * (easy_agent.py:36), frozen at the kickoff spot once reset() has re-allocated the obs
 * array; with random_opp=False the opponents' views are refreshed every step by
 * get_action_type (easy_agent.py:55-57) and are live. */
static int defence_near(const Ctx *c, int agent)
{
    const OracleV0Env *e = c->e;
    const double *view = (agent >= OPP_1 && !c->cfg->random_opp) ? e->obs[agent] : e->kick[agent];
    double v[2];
    if (agent <= AI_2) {                                   /* team 'left' */
        double o1 = get_vec(c, e->obs[OPP_1], view, v);
        double o2 = get_vec(c, e->obs[OPP_2], view, v);
        return bigger_than(o1, o2, 2);
    }
    double a1 = get_vec(c, e->obs[AI_1], view, v);
    double a2 = get_vec(c, e->obs[AI_2], view, v);
    return bigger_than(a1, a2, 2);
}

/* _set_vector_observation, :300-530 */
static void set_vector_observation(Ctx *c, int agent, int has_ball, int action, int set_target,
                                   const double target[2])
{
    OracleV0Env *e = c->e;
    double *ao = e->obs[agent];
    double *bo = e->obs[BALL];
    const int right = agent >= OPP_1;
    const double goal_x = right ? 0.0 : FIELD_LEN;
    double target_y = (double)rng_randint(c, 32, 36);       /* :306 (goal_down+3 .. goal_up-3) */
    double tgt[2] = { goal_x, target_y };
    double v[2];

    if (has_ball) {
        if (action == A_INTERCEPT) {                        /* :318-321 */
            ao[2] = ao[3] = ao[4] = 0;
            bo[2] = bo[3] = bo[4] = 0;
        } else if (action == A_RUN) {                       /* :330-356 */
            ao[4] = c->cfg->player_speed;
            if (set_target) {
                ao[2] = target[0]; ao[3] = target[1];
            } else {
                get_vec(c, tgt, ao, v);                     /* :348/:350 */
                ao[2] = v[0]; ao[3] = v[1];
            }
            if (rng_random(c) < 0.05)                       /* :353 */
                e->owner = NOONE;                           /* Q3: last_owner, ball untouched */
            else
                memcpy(bo, ao, sizeof(double) * 5);         /* :356 */
        } else if (action == A_SHOOT) {                     /* :362-383 */
            double accuracy = NORMAL_MISS + defence_near(c, agent) * UNDER_DEFENCE_MISS; /* :364 */
            int sp = (int)c->cfg->shoot_speed;
            bo[4] = rng_randint(c, sp - 16, sp) * 1.0;      /* :367 */
            double mag = get_vec(c, tgt, bo, v);            /* :373/:377 */
            double tw[2];
            screw_vec(c, v, mag, accuracy, tw);             /* :374/:378 */
            bo[2] = tw[0]; bo[3] = tw[1];
            e->last_owner = e->owner;                       /* :381 */
            e->owner = NOONE;                               /* :382 */
            ao[2] = ao[3] = ao[4] = 0;                      /* :383 */
        } else {                                            /* assist, :385-423 */
            const double *mate = e->obs[agent ^ 1];         /* :391-398 */
            double mag = get_vec(c, mate, bo, v);           /* :412 */
            double sp = mag / STEP_SIZE;                    /* :413 */
            if (sp > SHOOT_SPEED) sp = SHOOT_SPEED;         /* :414-415 */
            bo[4] = rng_uniform(c, sp - 1, sp + 1);         /* :416 */
            bo[2] = v[0]; bo[3] = v[1];                     /* :418 */
            e->last_owner = e->owner;                       /* :421 */
            e->owner = NOONE;                               /* :422 */
            ao[2] = ao[3] = ao[4] = 0;                      /* :423 */
        }
    } else {
        double b2a[2], g2a[2];
        double b2a_mag = get_vec(c, bo, ao, b2a);           /* :432 */
        double gc[2] = { goal_x, FIELD_WID / 2 };
        get_vec(c, gc, ao, g2a);                            /* :433-436 */
        if (action == A_INTERCEPT) {                        /* :452-476 ; Q4: player keeps moving */
            int success = rng_random(c) < intercept_chance(b2a_mag, MIN_INTERCEPT_DIST, MAX_INTERCEPT_DIST);
            if (success || (e->owner == NOONE && b2a_mag < MAX_INTERCEPT_DIST + 2)) { /* :462-463 */
                bo[2] = ao[2]; bo[3] = ao[3]; bo[4] = ao[4]; /* :465 */
                bo[0] = ao[0]; bo[1] = ao[1];               /* :466 */
                e->last_owner = e->owner;                   /* :467 */
                e->owner = agent;                           /* :468 */
            }
        } else if (action == A_RUN) {                       /* :483-503 */
            ao[4] = c->cfg->player_speed;
            if (set_target) {
                ao[2] = target[0]; ao[3] = target[1];
            } else if (e->owner != agent) {
                ao[2] = b2a[0]; ao[3] = b2a[1];             /* :501 */
            } else {
                ao[2] = g2a[0]; ao[3] = g2a[1];             /* :503 (dead, Q12) */
            }
        } else {                                            /* shoot / assist without ball: stop, :509-525 */
            ao[2] = ao[3] = ao[4] = 0;
        }
    }
}

/* _agent_set_vector_observation with action_set=True, :536-555 */
static void agent_set_vector_observation(Ctx *c, int agent, int action)
{
    int has_ball = (c->e->owner == agent);                  /* :538-546 */
    static const double zero[2] = { 0, 0 };
    set_vector_observation(c, agent, has_ball, action, 0, zero);
}

/* _step_by_observation, :560-571 */
static void step_by_observation(const Ctx *c, double *o, int is_ball)
{
    double tx = o[2], ty = o[3];
    double mag = sqrt(sq(c, tx) + sq(c, ty));               /* :562 */
    if (mag != 0) {
        o[0] += o[4] * (tx * STEP_SIZE / mag);              /* :567 */
        o[1] += o[4] * (ty * STEP_SIZE / mag);              /* :568 */
    }
    if (is_ball && c->e->owner == NOONE) o[4] -= DECELERATION * STEP_SIZE; /* :570-571 */
}

/* out, :574-577 */
static int is_out(const double *o) { return (o[0] < 0 || o[0] > FIELD_LEN) || (o[1] < 0 || o[1] > FIELD_WID); }

/* score, :580-583 */
static int is_score(const OracleV0Env *e)
{
    const double *b = e->obs[BALL];
    int ai_in = b[0] <= 0 && (b[1] > GOAL_LOWER && b[1] < GOAL_UPPER);
    int opp_in = b[0] >= FIELD_LEN && (b[1] > GOAL_LOWER && b[1] < GOAL_UPPER);
    return ai_in || opp_in;
}

/* fix, :587-604 (Q8) */
static void fix(OracleV0Env *e, int player)
{
    int new_owner = (player == OPP_1 || player == OPP_2) ? AI_1 : OPP_1;
    double *b = e->obs[BALL];
    b[0] = lock_in(b[0], FIELD_LEN);
    b[1] = lock_in(b[1], FIELD_WID);
    e->owner = new_owner;
    b[2] = b[3] = b[4] = 0;
    memcpy(e->obs[new_owner], b, sizeof(double) * 5);       /* :601-604 */
}

/* out_of_field, :621-625 (Q7) */
static int out_of_field(const OracleV0Env *e)
{
    const double *b = e->obs[BALL];
    int x_out = b[0] < 0 || b[0] > FIELD_LEN;
    int y_out = b[1] < 0 || b[1] > FIELD_WID;
    int y_score = b[1] > GOAL_LOWER - 2 && b[1] < GOAL_UPPER + 2;
    return (x_out && !y_score) || y_out;
}

/* ball_owner_array_update, :720-736 */
static void ball_owner_array_update(OracleV0Env *e)
{
    for (int i = 0; i < 5; ++i) e->obs[OWNER_ROW][i] = (i == e->owner) ? 10 : 0;
}

/* _get_reward, :752-861.  ball/ai_1/ai_2/owner_arr are the PRE-step snapshots (:630-635). */
static double get_reward(const Ctx *c, const double *ball, const double *ai_1, const double *ai_2,
                         const double *owner_arr, int action1, int action2)
{
    const OracleV0Env *e = c->e;
    double v[2];
    double ball_to_ai_1 = get_vec(c, ball, ai_1, v);        /* :757 */
    double ball_to_ai_2 = get_vec(c, ball, ai_2, v);        /* :758 */
    double running_r, player_adv_r, bad1, bad2, out_r, get_ball, score, get_scored;

    running_r = (action1 == A_RUN || action2 == A_RUN) ? 10 * PLAYER_ADV_REWARD_BASE : 0;       /* :772-775 */
    if ((owner_arr[AI_1] == 10 && action2 == A_RUN) || (owner_arr[AI_2] == 10 && action1 == A_RUN))
        player_adv_r = 10 * PLAYER_ADV_REWARD_BASE;                                               /* :777-781 */
    else
        player_adv_r = 0;

    if (owner_arr[AI_1] == 0) {                                                                   /* :783-794 */
        if (action1 == A_ASSIST || action1 == A_SHOOT) bad1 = 2 * BAD_ACTION_PENALTY;
        else if (ball_to_ai_1 > 2 && action1 == A_INTERCEPT) bad1 = 1 * BAD_ACTION_PENALTY;
        else bad1 = 0;
    } else {
        bad1 = (action1 == A_INTERCEPT) ? 2 * BAD_ACTION_PENALTY : 0;
    }
    if (owner_arr[AI_2] == 0) {                                                                   /* :796-807 */
        if (action2 == A_ASSIST || action2 == A_SHOOT) bad2 = 2 * BAD_ACTION_PENALTY;
        else if (ball_to_ai_2 > 2 && action1 == A_INTERCEPT) bad2 = 1 * BAD_ACTION_PENALTY;      /* Q5: action1 */
        else bad2 = 0;
    } else {
        bad2 = (action2 == A_INTERCEPT) ? 2 * BAD_ACTION_PENALTY : 0;
    }
    double bad_action_p = bad1 + bad2;                                                            /* :809 */

    out_r = (is_out(e->obs[AI_1]) || is_out(e->obs[AI_2])) ? OUT_OF_FIELD_PENALTY : 0;            /* :823-826 */

    if ((e->owner == AI_1 || e->owner == AI_2) && (owner_arr[AI_1] == 0 && owner_arr[AI_2] == 0)) { /* :828-836 */
        if (ball[2] > ball[3] && ball[2] > 0 && ball[0] > ai_1[0] && ball[0] > ai_2[0] && owner_arr[BALL] == 10)
            get_ball = -50 * BALL_CONTROL;                                                        /* Q6 */
        else
            get_ball = 60 * BALL_CONTROL;
    } else if ((e->owner == AI_1 && owner_arr[AI_1] == 10) || (e->owner == AI_2 && owner_arr[AI_2] == 10)) {
        get_ball = 30 * BALL_CONTROL;                                                             /* :837-839 */
    } else {
        get_ball = 0;
    }

    score = (is_score(e) && e->obs[BALL][0] >= FIELD_LEN) ? GOAL_REWARD : 0;                      /* :843-848 */
    get_scored = (is_score(e) && e->obs[BALL][0] <= 0) ? -GOAL_REWARD : 0;                        /* :850-855 */

    if (c->cfg->only_reward_goal) return score + get_scored;                                      /* :857-858 */
    return get_ball + score + get_scored + out_r + bad_action_p + player_adv_r + running_r;      /* :861 */
}

/* Easy_Agent.get_action_type, easy_agent.py:53-98 (agents are the two 'right' opponents) */
static int easy_get_action_type(Ctx *c, int agent, int has_ball, int team_has_ball)
{
    const OracleV0Env *e = c->e;
    const double *ao = e->obs[agent], *mo = e->obs[agent ^ 1], *bo = e->obs[BALL];
    double v[2];
    double ball_mag = get_vec(c, bo, ao, v);                /* :67-68 */
    double mate_mag = get_vec(c, mo, ao, v);                /* :69-70 */
    const double shoot_x = 0 + 20;                          /* :30-32, shoot_range = 20 (futbol_env.py:196) */
    if (has_ball) {
        if (ao[0] <= shoot_x) return A_SHOOT;               /* :77-79 (team == 'right') */
        if ((mo[0] < ao[0] || mo[1] < ao[1] - 7 || mo[1] > ao[1] + 7)
            && rng_random(c) > 0.8 && mate_mag > 12)        /* :81-83, short-circuit order kept */
            return A_ASSIST;
        return A_RUN;
    }
    if (ball_mag <= 1 && !team_has_ball) return A_INTERCEPT; /* :90-92 */
    return A_RUN;
}

/* _opp_team_set_vector_observation, :864-982 */
static void opp_team_set_vector_observation(Ctx *c)
{
    OracleV0Env *e = c->e;
    const double length = FIELD_LEN, width = FIELD_WID;
    int opp_1_has = (e->owner == OPP_1), opp_2_has = (e->owner == OPP_2);   /* :866-877 */
    int team_has = opp_1_has || opp_2_has;

    int a1_type = easy_get_action_type(c, OPP_1, opp_1_has, team_has);     /* :879 */
    int a2_type = easy_get_action_type(c, OPP_2, opp_2_has, team_has);     /* :880 */
    const int a1 = a1_type, a2 = a2_type;   /* opp1_action / opp2_action: NOT updated by the override below */

    int t1_set = 0, t2_set = 0;
    double t1[2] = { 0, 0 }, t2[2] = { 0, 0 };
    double *o1 = e->obs[OPP_1], *o2 = e->obs[OPP_2], *b = e->obs[BALL];

    if (opp_1_has) {                                                        /* :893-909 */
        if (a1 == A_RUN) {
            if (o1[1] > width * 0.2) { t1_set = 1; t1[0] = -1; t1[1] = -1; }
            if (a2 == A_RUN && o2[0] > length * 0.1) {
                if (o2[1] < width * 0.8) { t2_set = 1; t2[0] = -1; t2[1] = 1; }
            }
        }
    }
    if (opp_2_has) {                                                        /* :911-928 */
        if (a2 == A_RUN) {
            if (o2[1] < width * 0.8) { t2_set = 1; t2[0] = -1; t2[1] = 1; }
            if (a1 == A_RUN && o1[0] > length * 0.1) {
                if (o1[1] > width * 0.2) { t1_set = 1; t1[0] = -1; t1[1] = -1; }
            }
        }
    }
    if (e->owner == AI_1 || e->owner == AI_2) {                             /* :931-947 */
        if (b[0] < length * 0.6) {
            double dp[2] = { length * 0.75, width * 0.5 };
            if (o1[0] > o2[0]) { a1_type = A_RUN; t1_set = 1; get_vec(c, dp, o1, t1); }
            else               { a2_type = A_RUN; t2_set = 1; get_vec(c, dp, o2, t2); }
        }
    }

    set_vector_observation(c, OPP_1, opp_1_has, a1_type, t1_set, t1);       /* :951-954 */
    set_vector_observation(c, OPP_2, opp_2_has, a2_type, t2_set, t2);       /* :956-959 */

    if (e->owner == NOONE && a1 == A_RUN && a2 == A_RUN) {                  /* :962-982 */
        double nb[5];
        memcpy(nb, b, sizeof(nb));
        {   /* _step_by_observation(ball_next_obs), is_ball=False */
            double tx = nb[2], ty = nb[3];
            double mag = sqrt(sq(c, tx) + sq(c, ty));
            if (mag != 0) {
                nb[0] += nb[4] * (tx * STEP_SIZE / mag);
                nb[1] += nb[4] * (ty * STEP_SIZE / mag);
            }
        }
        double v1[2], v2[2];
        double m1 = get_vec(c, nb, o1, v1);                                 /* :969 */
        double m2 = get_vec(c, nb, o2, v2);                                 /* :970 */
        if (m1 < STEP_SIZE * c->cfg->player_speed) {                        /* :972-976 */
            o1[2] = v1[0]; o1[3] = v1[1]; o1[4] = m1 / STEP_SIZE;
        } else if (m2 < STEP_SIZE * c->cfg->player_speed) {                 /* :978-982 */
            o2[2] = v2[0]; o2[3] = v2[1]; o2[4] = m2 / STEP_SIZE;
        }
    }
}

/* step, :628-717.  Returns done. */
/* opp_action: -1 = the reference's own opponents (random draw :641 or the hard-coded team :649); 0..15 = actions
 * supplied by the caller for opp_1 (a / 4) and opp_2 (a % 4) -- the self-play hook: they go through the same
 * _agent_set_vector_observation path as the random opponents (:642-645) and the randint(0, 15) draw is not taken. */
int futbol_v0_oracle_step_vs(const OracleV0Config *cfg, OracleV0Env *e, int ai_action, int opp_action, double *reward_out)
{
    Ctx ctx = { cfg, e };
    Ctx *c = &ctx;
    double o_b[5], o_ai_1[5], o_ai_2[5], o_owner[5];
    memcpy(o_b, e->obs[BALL], sizeof(o_b));                 /* :630-635 */
    memcpy(o_ai_1, e->obs[AI_1], sizeof(o_ai_1));
    memcpy(o_ai_2, e->obs[AI_2], sizeof(o_ai_2));
    memcpy(o_owner, e->obs[OWNER_ROW], sizeof(o_owner));
    e->flags = 0;
    e->step_draws = 0;
    e->normal_calls = 0;

    if (opp_action >= 0 || cfg->random_opp) {               /* :639-645 */
        int r = opp_action >= 0 ? (opp_action & 15) : rng_randint(c, 0, 15);
        agent_set_vector_observation(c, OPP_1, r / 4);
        agent_set_vector_observation(c, OPP_2, r % 4);
    } else {
        opp_team_set_vector_observation(c);                 /* :649 */
    }
    int action1 = ai_action / 4, action2 = ai_action % 4;   /* :653 */
    agent_set_vector_observation(c, AI_1, action1);         /* :655 */
    agent_set_vector_observation(c, AI_2, action2);         /* :656 */

    for (int p = 0; p < 4; ++p) step_by_observation(c, e->obs[p], 0); /* :661 */
    step_by_observation(c, e->obs[BALL], 1);                /* :663 */

    double reward = get_reward(c, o_b, o_ai_1, o_ai_2, o_owner, action1, action2); /* :666 */
    int done = 0;

    if (is_score(e)) {                                      /* :670-699 */
        if (e->obs[BALL][0] <= 0) e->opp_score += 1; else e->ai_score += 1;
        if (cfg->one_goal_end) done = 1;
        kickoff_rows(e->obs);
        e->owner = NOONE;
        e->last_owner = NOONE;
        e->flags |= 1;
    }
    if (out_of_field(e)) {                                  /* :701-707 */
        fix(e, e->last_owner);
        if (cfg->one_goal_end) done = 1;
        e->flags |= 2;
    }
    ball_owner_array_update(e);                             /* :709 */
    if (e->time >= cfg->game_time) done = 1;                /* :712-713 */
    e->time += STEP_SIZE;                                   /* :716 */
    e->t_total += 1;
    *reward_out = reward;
    return done;
}

/* ---- batch drivers (harness level, not part of the reference) ------------------------- */
int futbol_v0_oracle_step(const OracleV0Config *cfg, OracleV0Env *e, int ai_action, double *reward_out)
{
    return futbol_v0_oracle_step_vs(cfg, e, ai_action, -1, reward_out);
}

/*
 * Steps n envs (global ids env_id0 .. env_id0+n-1) ``steps`` times.
 *   actions: [steps][n] uint8 or NULL (=> Philox action stream, index = env total step)
 *   autoreset: 0 none; 1 record the terminal obs then reset (golden-harness semantics);
 *              2 VecEnv semantics: on done the obs slot holds the reset obs.
 * Any output pointer may be NULL.  obs: [steps][n][30] doubles.
 */
typedef struct {
    const OracleV0Config *cfg; OracleV0Env *envs; int n, steps, lo, hi, autoreset;
    const uint8_t *actions, *opp_actions;
    double *obs, *reward; uint8_t *done, *owner, *last_owner; int32_t *ai_score, *opp_score;
    uint64_t *draws; uint8_t *flags;
} RolloutJob;

static void *rollout_worker(void *arg)
{
    const RolloutJob *j = (const RolloutJob *)arg;
    const int n = j->n;
    for (int i = j->lo; i < j->hi; ++i) {
        OracleV0Env *e = &j->envs[i];
        for (int t = 0; t < j->steps; ++t) {
            size_t k = (size_t)t * n + i;
            int a = j->actions ? j->actions[k] : futbol_oracle_action(j->cfg->seed, e->env_id, e->t_total, 16);
            double r;
            int d = futbol_v0_oracle_step_vs(j->cfg, e, a, j->opp_actions ? j->opp_actions[k] : -1, &r);
            int fl = e->flags;
            if (j->obs && (j->autoreset != 2 || !d)) memcpy(j->obs + k * 30, e->obs, sizeof(double) * 30);
            if (j->reward) j->reward[k] = r;
            if (j->done) j->done[k] = (uint8_t)d;
            if (j->owner) j->owner[k] = (uint8_t)e->owner;
            if (j->last_owner) j->last_owner[k] = (uint8_t)e->last_owner;
            if (j->ai_score) j->ai_score[k] = e->ai_score;
            if (j->opp_score) j->opp_score[k] = e->opp_score;
            if (j->flags) j->flags[k] = (uint8_t)fl;
            if (d && j->autoreset) {
                futbol_v0_oracle_reset(e);
                if (j->autoreset == 2 && j->obs) memcpy(j->obs + k * 30, e->obs, sizeof(double) * 30);
            }
            if (j->draws) j->draws[k] = e->draw_ctr;
        }
    }
    return NULL;
}

void futbol_v0_oracle_rollout_vs(const OracleV0Config *cfg, OracleV0Env *envs, int n, int steps,
                              const uint8_t *actions, const uint8_t *opp_actions, int autoreset, int n_threads,
                              double *obs, double *reward, uint8_t *done, uint8_t *owner,
                              uint8_t *last_owner, int32_t *ai_score, int32_t *opp_score,
                              uint64_t *draws, uint8_t *flags)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    if (n_threads > n) n_threads = n > 0 ? n : 1;
    RolloutJob jobs[256];
    pthread_t th[256];
    for (int w = 0; w < n_threads; ++w) {
        RolloutJob j = { cfg, envs, n, steps, (int)((long)n * w / n_threads), (int)((long)n * (w + 1) / n_threads),
                         autoreset, actions, opp_actions, obs, reward, done, owner, last_owner, ai_score, opp_score, draws, flags };
        jobs[w] = j;
    }
    if (n_threads == 1) { rollout_worker(&jobs[0]); return; }
    for (int w = 0; w < n_threads; ++w) pthread_create(&th[w], NULL, rollout_worker, &jobs[w]);
    for (int w = 0; w < n_threads; ++w) pthread_join(th[w], NULL);
}

void futbol_v0_oracle_rollout(const OracleV0Config *cfg, OracleV0Env *envs, int n, int steps,
                              const uint8_t *actions, int autoreset, int n_threads,
                              double *obs, double *reward, uint8_t *done, uint8_t *owner,
                              uint8_t *last_owner, int32_t *ai_score, int32_t *opp_score,
                              uint64_t *draws, uint8_t *flags)
{
    futbol_v0_oracle_rollout_vs(cfg, envs, n, steps, actions, NULL, autoreset, n_threads, obs, reward, done, owner, last_owner,
                                ai_score, opp_score, draws, flags);
}

size_t futbol_v0_oracle_env_bytes(void) { return sizeof(OracleV0Env); }
size_t futbol_v0_oracle_cfg_bytes(void) { return sizeof(OracleV0Config); }
