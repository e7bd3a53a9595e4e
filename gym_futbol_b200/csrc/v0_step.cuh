// Device-side v0 environment step: one thread owns one environment, all 25 state doubles live in
// registers.  Replaces gym_futbol/envs/futbol_env.py FutbolEnv.step (:628-717) and everything it
// calls, including Easy_Agent.get_action_type (gym_futbol/envs/easy_agent.py:53-98).
//
// Arithmetic contract ("kernel arithmetic"): fp64, the reference's operation order, no FMA
// contraction (every product and sum goes through __dmul_rn/__dadd_rn/..., which the compiler never
// fuses), IEEE sqrt/div.  Numpy-scalar x**2 is computed as x*x, and log/sin/cos are the fully
// specified fm_log/fm_sincos below (fdlibm-style polynomials, individually rounded IEEE operations in a
// fixed order), so that a CPU restatement of the same specification agrees BIT for bit.
//
// Shape of the code ("branch-lean").  The reference picks one of eight branches per player
// (has-ball x action); the 32 environments of a warp pick 32 different ones, so a branchy transcription
// executes the union of all branches at ~13 active lanes and, inlined four times, overflows the 32 KB
// instruction cache (ncu: 76 % of stall samples "no instruction").  Here every player turn is ONE
// straight-line block: the case is classified into predicates, the single vector magnitude any case
// needs is computed once, results are committed with selects.  What stays behind a branch is rare:
// the kick (screw_vec: log, sin, cos) -- deferred to one shared site per step because at most one player
// can shoot per step -- and draws past the eight pre-generated Philox words.  Threshold tests on a
// distance (d <= 2, d > 12, ...) are done on the squared distance with the exactly equivalent bound
// (kSq* below), which removes the square root without changing a single decision.
#pragma once
#include <stdint.h>
#include "philox.cuh"

namespace futbol {

// action.py:3-6 / ballowner.py:3-7
enum : int { kRun = 0, kIntercept = 1, kShoot = 2, kAssist = 3 };
enum : int { kAI1 = 0, kAI2 = 1, kOpp1 = 2, kOpp2 = 3, kNoOne = 4 };
enum : int { kFlagGoal = 1, kFlagFix = 2, kFlagDone = 4 };

constexpr double kFieldLen = 105.0, kFieldWid = 68.0;   // futbol_env.py:18-19
constexpr double kGoalLower = 29.0, kGoalUpper = 39.0;  // :23-24
constexpr double kStepSize = 0.1;                       // :45

// For s = fl(fl(dx*dx) + fl(dy*dy)) and d = sqrt_rn(s) (correctly rounded, hence monotone):
//   d <= 1.0  <=>  s <= nextafter(1, +inf)
//   d <= 2.0  <=>  s <= nextafter(4, +inf)
//   d > 12.0  <=>  s >  144.0
// (sqrt(s) rounds to c or below exactly when it lies below the midpoint of c and its successor; squaring
// that midpoint gives the bound.  tests/test_sqrt_thresholds.py checks every neighbouring double.)
constexpr double kSqLe1 = 0x1.0000000000001p+0;
constexpr double kSqLe2 = 0x1.0000000000001p+2;
constexpr double kSqGt12 = 144.0;
// products of reference constants, rounded once as the reference's float64 multiply rounds them
// (__dmul_rn is not folded by the compiler; tests/test_sqrt_thresholds.py::test_constant_products)
constexpr double kWid02 = 13.600000000000001;   // width * 0.2, :895
constexpr double kWid08 = 54.400000000000006;   // width * 0.8, :913
constexpr double kLen01 = 10.5;                 // length * 0.1, :903
constexpr double kLen06 = 63.0;                 // length * 0.6, :931
constexpr double kDefendX = 78.75, kDefendY = 34.0;   // (length * 0.75, width * 0.5), :933
constexpr double kRewStolen = -15.0, kRewGained = 18.0, kRewKept = 9.0;   // -50*0.3, 60*0.3, 30*0.3, :832-839

struct Row { double x, y, tx, ty, sp; };

struct V0Params {
    uint64_t seed;
    PhiloxKey key;       // philox_expand_key(seed), set by the host
    uint32_t env_id_offset;
    int n_envs;
    int random_opp, one_goal_end, only_reward_goal, auto_reset;
    int ep_limit;        // first ep_step at which `time >= game_time` holds (400 for game_time 40)
    int shoot_speed;
    double player_speed;
    double reach_sq_max; // largest s with sqrt_rn(s) < fl(0.1 * player_speed) (:972-976); set by the host
};

struct V0State {
    Row p[4];           // ai_1, ai_2, opp_1, opp_2
    Row b;              // ball
    uint64_t t_total;
    int ep_step, ai_score, opp_score, owner, last_owner;
};

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
// c ? a : b that the optimiser cannot look through.  Used where a lane is handed a benign operand to keep
// it on the fast path of the IEEE sqrt/div sequence: with a plain ternary the compiler rewrites
// sqrt(c ? q : 1.0) into c ? sqrt(q) : 1.0 and the zero operand is back (seen in ncu as 1-6 lanes per warp
// inside __cuda_sm20_dsqrt_rn_f64_mediumpath / div_rn_f64_full on every step).
#ifndef FUTBOL_HOST_SHIM
__device__ __forceinline__ double pick(bool c, double a, double b)
{
    double r;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\tselp.f64 %0, %1, %2, p;\n\t}" : "=d"(r) : "d"(a), "d"(b), "r"((int)c));
    return r;
}
#else
inline double pick(bool c, double a, double b) { return c ? a : b; }
#endif
__device__ __forceinline__ double sqsum(double vx, double vy) { return dadd(dmul(vx, vx), dmul(vy, vy)); }
__device__ __forceinline__ double hyp(double vx, double vy) { return __dsqrt_rn(sqsum(vx, vy)); }   // get_vec, :62-65

// ---- specified elementary functions (domain: log on (0,1]; sin/cos on |x| <= 2*pi) --------------------
__device__ __forceinline__ double fm_log(double x)
{
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
        Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
        Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
        Lg7 = 1.479819860511658591e-01;
    long long b = __double_as_longlong(x);
    int k = (int)((b >> 52) & 0x7ff) - 1023;
    double m = __longlong_as_double((b & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
    if (m > 1.4142135623730951) { m = dmul(m, 0.5); k += 1; }
    const double f = dsub(m, 1.0);
    const double s = ddiv(f, dadd(2.0, f));
    const double z = dmul(s, s), w = dmul(z, z);
    const double t1 = dmul(w, dadd(Lg2, dmul(w, dadd(Lg4, dmul(w, Lg6)))));
    const double t2 = dmul(z, dadd(Lg1, dmul(w, dadd(Lg3, dmul(w, dadd(Lg5, dmul(w, Lg7)))))));
    const double R = dadd(t2, t1);
    const double hfsq = dmul(dmul(0.5, f), f);
    const double dk = (double)k;
    return dsub(dmul(dk, ln2_hi), dsub(dsub(hfsq, dadd(dmul(s, dadd(hfsq, R)), dmul(dk, ln2_lo))), f));
}

// out of line: only the kick uses it (twice), a few lanes per warp
static __device__ __noinline__ void fm_sincos(double x, double &sn, double &cs)
{
    const double invpio2 = 6.36619772367581382433e-01, pio2_1 = 1.57079632673412561417e+00,
        pio2_1t = 6.07710050650619224932e-11;
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
        S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
        C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const double fn = rint(dmul(x, invpio2));
    const int n = (int)fn;
    const double r = dsub(dsub(x, dmul(fn, pio2_1)), dmul(fn, pio2_1t));
    const double z = dmul(r, r), v = dmul(z, r);
    const double rs = dadd(S2, dmul(z, dadd(S3, dmul(z, dadd(S4, dmul(z, dadd(S5, dmul(z, S6))))))));
    const double sr = dadd(r, dmul(v, dadd(S1, dmul(z, rs))));
    const double rc = dmul(z, dadd(C1, dmul(z, dadd(C2, dmul(z, dadd(C3, dmul(z, dadd(C4, dmul(z, dadd(C5, dmul(z, C6)))))))))));
    const double cr = dadd(dsub(1.0, dmul(0.5, z)), dmul(z, rc));
    switch (n & 3) {
    case 0: sn = sr; cs = cr; break;
    case 1: sn = cr; cs = -sr; break;
    case 2: sn = -sr; cs = -cr; break;
    default: sn = -cr; cs = sr; break;
    }
}

__device__ __forceinline__ void zero_motion(Row &r) { r.tx = 0.0; r.ty = 0.0; r.sp = 0.0; }

__device__ __forceinline__ void kickoff(V0State &s)
{   // futbol_env.py:211-223 (also the goal re-kickoff :684-692)
    s.b = Row{kFieldLen / 2, kFieldWid / 2, 0, 0, 0};
    s.p[kAI1] = Row{kFieldLen / 2 - 9, kFieldWid / 2 + 5, 0, 0, 0};
    s.p[kAI2] = Row{kFieldLen / 2 - 9, kFieldWid / 2 - 5, 0, 0, 0};
    s.p[kOpp1] = Row{kFieldLen / 2 + 9, kFieldWid / 2 + 5, 0, 0, 0};
    s.p[kOpp2] = Row{kFieldLen / 2 + 9, kFieldWid / 2 - 5, 0, 0, 0};
    s.owner = kNoOne;
    s.last_owner = kNoOne;
}

__device__ __forceinline__ void reset_env(V0State &s)
{   // FutbolEnv.reset, :205-245.  t_total (the Philox step index) is deliberately kept.
    kickoff(s);
    s.ep_step = 0;
    s.ai_score = 0;
    s.opp_score = 0;
}

// A kick waiting to be resolved.  At most one player can shoot in a step: shooting needs the ball at
// the player's own turn, a shot leaves the ball with no one, and the only way to gain it back inside
// the step is an intercept -- which is that player's single action of the step (for the hard-coded
// opponents has_ball is latched before either acts, :866-877, and cannot hold for both).  So the
// expensive part of shoot (defence_near + screw_vec, :364-378) is evaluated once per step, after the
// last player turn: everything it reads (player and ball POSITIONS) only changes in the kinematics
// phase, and the vector it writes is first read there.  A later successful intercept overwrites the
// whole ball row (:465-466), which cancels the pending kick.
struct PendingShot { int shooter; int target_y; uint32_t pick_idx; };

// Draw budget of a step (sequential draws; the kick's normal() slots are addressed separately).  In turn
// order: [random_opp: 1 | hard-coded: <= 1, only the opponent holding the ball can draw in get_action_type],
// then per player 1 (target_y) + 1 if it draws (+1 more for the single possible shooter).  The cursor
// before a site is therefore at most: opp_1 turn 1, 2, 3; opp_2 turn 4, 5; ai_1 turn 6, 7 (+ pick 8);
// ai_2 turn 8, 9 (+ pick 10) -- a step consumes at most 10 draws, and only ai_2's sites (and a kick's pick)
// can reach past the kPreDraws = 8 words parked in shared memory, so only those carry the range check.

// _set_vector_observation, :300-530, for one player: straight-line, predicated.
template <int AGENT>
__device__ __forceinline__ void player_turn(V0State &s, V0Rng &rng, const V0Params &P, bool has_ball, int action,
                                            bool set_target, double tgx, double tgy, PendingShot &shot)
{
    constexpr bool right = AGENT >= kOpp1;
    constexpr double goal_x = right ? 0.0 : kFieldLen;
    Row &ao = s.p[AGENT];
    const Row &mate = s.p[AGENT ^ 1];

    constexpr bool kChecked = AGENT == kAI2;                             // see "Draw budget" above
    const int target_y = 32 + (int)__umulhi(rng.take<kChecked>(), 5u);   // randint(32, 36), :306 -- always drawn first
    const bool is_run = action == kRun, is_int = action == kIntercept;
    const bool hb_run = has_ball && is_run, hb_int = has_ball && is_int;
    const bool hb_shoot = has_ball && action == kShoot, hb_assist = has_ball && action == kAssist;
    const bool nb_int = !has_ball && is_int;
    // one more draw in: has-ball run (:353), shoot (:367), assist (:416); no-ball intercept (:459, even when far)
    const uint32_t w = rng.take_if<kChecked>(has_ball != is_int);
    const double u = (double)(w >> 8) * (1.0 / 16777216.0);

    const double bx = dsub(s.b.x, ao.x), by = dsub(s.b.y, ao.y);         // :432
    const double vx = hb_assist ? dsub(mate.x, s.b.x) : bx;              // :412
    const double vy = hb_assist ? dsub(mate.y, s.b.y) : by;
    // The one magnitude a turn needs (no-ball intercept: |ball - player|; has-ball assist: |mate - ball|).
    // sqrt(0) and 0/x leave the fast path of the IEEE sequences (a subroutine call for the lanes that
    // hold the ball, whose ball-player vector is exactly zero), so those lanes are fed a benign operand
    // and the exact result (0) is selected afterwards.
    const double q = sqsum(vx, vy);
    const bool q_zero = q == 0.0;
    const bool need_mag = nb_int || hb_assist;
    const double root = __dsqrt_rn(pick(need_mag && !q_zero, q, 1.0));
    const double mag = q_zero ? 0.0 : root;

    // has-ball assist, :413-416
    const double quot = ddiv(pick(hb_assist && !q_zero, mag, 1.0), kStepSize);
    double pass = q_zero ? 0.0 : quot;
    pass = pass > 20.0 ? 20.0 : pass;
    const double lo = dsub(pass, 1.0), hi = dadd(pass, 1.0);
    const double pass_speed = dadd(lo, dmul(dsub(hi, lo), u));           // random.uniform
    // no-ball intercept, :452-463; intercept_chance(d, 1, 2) :122-129 with k = 0.9 / (1 - 2) = -0.9 exactly
    const double chance = mag < 1.0 ? 0.9 : (mag <= 2.0 ? dmul(-0.9, dsub(mag, 2.0)) : 0.0);
    const bool take = nb_int && (u < chance || (s.owner == kNoOne && mag < 4.0));
    // has-ball run, :353-356 (Q3)
    const bool drop = hb_run && u < 0.05;
    const bool carry = hb_run && !drop;

    // the player's own row (Q4: a no-ball intercept keeps the previous vector and speed)
    if (is_run) {                                                        // :330-352, :483-503 (:503 unreachable, Q12)
        ao.sp = P.player_speed;
        ao.tx = set_target ? tgx : (has_ball ? dsub(goal_x, ao.x) : bx);
        ao.ty = set_target ? tgy : (has_ball ? dsub((double)target_y, ao.y) : by);
    } else if (!nb_int) {
        zero_motion(ao);                                                 // :318-321, :383, :423, :509-525
    }
    // ball and possession
    if (carry || take) s.b = ao;                                         // :356, :465-466
    if (hb_int) zero_motion(s.b);                                        // :321
    if (hb_assist) { s.b.sp = pass_speed; s.b.tx = vx; s.b.ty = vy; }    // :416-418
    if (hb_shoot) {                                                      // :367; vector resolved later (PendingShot)
        s.b.sp = (double)(P.shoot_speed - 16 + (int)__umulhi(w, 17u));
        shot.shooter = AGENT; shot.target_y = target_y; shot.pick_idx = rng.j;
        rng.j += 1;                                                      // randint(0, 9) of screw_vec, :107
    }
    if (take) shot.shooter = -1;
    if (hb_shoot || hb_assist || take) s.last_owner = s.owner;           // :381, :421, :467
    s.owner = take ? (int)AGENT : ((drop || hb_shoot || hb_assist) ? (int)kNoOne : s.owner);   // :354, :382, :422, :468
}

// defence_near (:280-289, with the stale-view behaviour Q1) + screw_vec (:101-116) for the pending kick.
// By specification (oracle/philox.py) a step's normal() call consumes no sequential draws: slot k lives in
// Philox block 0x8000 + (k >> 1), words 2(k&1), 2(k&1)+1, so only the block of the picked slot is evaluated.
__device__ __forceinline__ void resolve_shot(V0State &s, const V0Rng &rng, const V0Params &P, const PendingShot &shot)
{
    const int a = shot.shooter;
    const bool right = a >= kOpp1;
    // the shooter's own position is the frozen kickoff spot, except for hard-coded opponents (views refreshed)
    double px = right ? kFieldLen / 2 + 9 : kFieldLen / 2 - 9;
    double py = (a == kAI1 || a == kOpp1) ? kFieldWid / 2 + 5 : kFieldWid / 2 - 5;
    if (right && !P.random_opp) {
        px = a == kOpp1 ? s.p[kOpp1].x : s.p[kOpp2].x;
        py = a == kOpp1 ? s.p[kOpp1].y : s.p[kOpp2].y;
    }
    const double d1x = right ? s.p[kAI1].x : s.p[kOpp1].x, d1y = right ? s.p[kAI1].y : s.p[kOpp1].y;
    const double d2x = right ? s.p[kAI2].x : s.p[kOpp2].x, d2y = right ? s.p[kAI2].y : s.p[kOpp2].y;
    // bigger_than(d1, d2, 2), :76-82 = how many of the two defenders are within 2.0
    const int near = (sqsum(dsub(d1x, px), dsub(d1y, py)) <= kSqLe2 ? 1 : 0) +
                     (sqsum(dsub(d2x, px), dsub(d2y, py)) <= kSqLe2 ? 1 : 0);
    const double accuracy = dadd(10.0, dmul((double)near, 20.0));        // :364
    const double vx = dsub(right ? 0.0 : kFieldLen, s.b.x), vy = dsub((double)shot.target_y, s.b.y);   // :373-376
    const double mag = hyp(vx, vy);

    const uint32_t pick = __umulhi(rng.word_at(shot.pick_idx), 10u);     // randint(0, 9), :107
    const Philox4 nb = philox_step_block(P.key, rng.env_id, rng.stream, rng.t, kNormalBlock0 + (pick >> 1));
    const uint32_t w0 = (pick & 1u) ? nb.z : nb.x, w1 = (pick & 1u) ? nb.w : nb.y;
    const double u1 = (double)((w0 >> 8) + 1u) * (1.0 / 16777216.0);
    const double u2 = (double)(w1 >> 8) * (1.0 / 16777216.0);
    double bm_sin, bm_cos;
    fm_sincos(dmul(6.283185307179586, u2), bm_sin, bm_cos);
    const double z = dmul(__dsqrt_rn(dmul(-2.0, fm_log(u1))), bm_cos);
    const double nd = dadd(0.0, dmul(accuracy, z));                      // np.random.normal(0, accuracy), :103
    const double c = ddiv(dmul(vx, 1.0), mag), sn = ddiv(dmul(vy, 1.0), mag);  // :105-106
    const double swing = dmul(ddiv(nd, 180.0), 3.141592653589793);       // :108
    double ss, sc;
    fm_sincos(swing, ss, sc);                                            // :109-110
    const double tc = dsub(dmul(c, sc), dmul(sn, ss));                   // :113
    const double ts = dadd(dmul(sn, sc), dmul(c, ss));                   // :114
    s.b.tx = dmul(tc, mag);                                              // :115
    s.b.ty = dmul(ts, mag);
}

// Easy_Agent.get_action_type for a 'right' opponent, easy_agent.py:53-98
template <int AGENT>
__device__ __forceinline__ int easy_action(const V0State &s, V0Rng &rng, bool has_ball, bool team_has_ball)
{
    const Row &ao = s.p[AGENT];
    const Row &mo = s.p[AGENT ^ 1];
    const bool in_range = ao.x <= 20.0;                                  // :77-79, shoot_x = 0 + 20
    const bool open_mate = mo.x < ao.x || mo.y < dsub(ao.y, 7.0) || mo.y > dadd(ao.y, 7.0);   // :81-83
    // the draw happens only when the geometric clause holds (short-circuit `and`, :81-85)
    const uint32_t w = rng.take_if<false>(has_ball && !in_range && open_mate);
    const bool lucky = (double)(w >> 8) * (1.0 / 16777216.0) > 0.8;
    const bool far_mate = sqsum(dsub(mo.x, ao.x), dsub(mo.y, ao.y)) > kSqGt12;                // distance > 12
    const bool ball_close = sqsum(dsub(s.b.x, ao.x), dsub(s.b.y, ao.y)) <= kSqLe1;           // distance <= 1.0, :90
    const int with_ball = in_range ? (int)kShoot : ((open_mate && lucky && far_mate) ? (int)kAssist : (int)kRun);
    const int without = (!team_has_ball && ball_close) ? (int)kIntercept : (int)kRun;         // :90-96
    return has_ball ? with_ball : without;
}

// _step_by_observation, :560-571 (DECELERATION = 0: the ball's speed update is `sp -= 0.0`).
// Straight-line: a stopped row (|t| == 0, :563) and a zero component (0 / mag = that same signed zero) are
// fed benign operands so that every lane stays on the fast path of sqrt/div, and the five rows of a step
// interleave in the instruction stream instead of being fenced by branches.
__device__ __forceinline__ void advance(Row &o)
{
    const double s2 = sqsum(o.tx, o.ty);                                 // :562; sqrt(s2) == 0 <=> s2 == 0
    const bool moving = s2 != 0.0;
    const double mag = __dsqrt_rn(pick(moving, s2, 1.0));
    const double nx = dmul(o.tx, kStepSize), ny = dmul(o.ty, kStepSize);
    const bool zx = nx == 0.0, zy = ny == 0.0;
    const double qx = ddiv(pick(zx, mag, nx), mag), qy = ddiv(pick(zy, mag, ny), mag);
    const double x1 = dadd(o.x, dmul(o.sp, zx ? nx : qx));               // :567
    const double y1 = dadd(o.y, dmul(o.sp, zy ? ny : qy));               // :568
    o.x = moving ? x1 : o.x;
    o.y = moving ? y1 : o.y;
}

// _opp_team_set_vector_observation, :864-982
__device__ __forceinline__ void opp_team(V0State &s, V0Rng &rng, const V0Params &P, PendingShot &shot)
{
    const bool has1 = s.owner == kOpp1, has2 = s.owner == kOpp2;         // :866-877 (latched before either acts)
    const bool team_has = has1 || has2;
    const int a1 = easy_action<kOpp1>(s, rng, has1, team_has);           // :879
    const int a2 = easy_action<kOpp2>(s, rng, has2, team_has);           // :880
    const bool run1 = a1 == kRun, run2 = a2 == kRun;
    int a1_type = a1, a2_type = a2;                                      // overrides do not touch a1 / a2
    const Row &o1 = s.p[kOpp1], &o2 = s.p[kOpp2];
    const bool diag1 = o1.y > kWid02, diag2 = o2.y < kWid08;
    // carrier runs on the diagonal, its running mate mirrors it, :893-928
    bool set1 = diag1 && ((has1 && run1) || (has2 && run2 && run1 && o1.x > kLen01));
    bool set2 = diag2 && ((has2 && run2) || (has1 && run1 && run2 && o2.x > kLen01));
    double t1x = -1.0, t1y = -1.0, t2x = -1.0, t2y = 1.0;
    if ((s.owner == kAI1 || s.owner == kAI2) && s.b.x < kLen06) {   // :931-947: the deeper opp defends
        const double dpx = kDefendX, dpy = kDefendY;
        if (o1.x > o2.x) { a1_type = kRun; set1 = true; t1x = dsub(dpx, o1.x); t1y = dsub(dpy, o1.y); }
        else             { a2_type = kRun; set2 = true; t2x = dsub(dpx, o2.x); t2y = dsub(dpy, o2.y); }
    }
    player_turn<kOpp1>(s, rng, P, has1, a1_type, set1, t1x, t1y, shot);  // :951-954
    player_turn<kOpp2>(s, rng, P, has2, a2_type, set2, t2x, t2y, shot);  // :956-959
    {   // :962-982 anticipate the ball: whoever can reach its next position lands exactly on it
        Row nb = s.b;
        advance(nb);
        const double v1x = dsub(nb.x, s.p[kOpp1].x), v1y = dsub(nb.y, s.p[kOpp1].y);
        const double v2x = dsub(nb.x, s.p[kOpp2].x), v2y = dsub(nb.y, s.p[kOpp2].y);
        const double q1 = sqsum(v1x, v1y), q2 = sqsum(v2x, v2y);
        const bool on = s.owner == kNoOne && run1 && run2;
        const bool c1 = on && q1 <= P.reach_sq_max;                      // hyp(v1) < 0.1 * player_speed
        const bool c2 = on && !c1 && q2 <= P.reach_sq_max;
        const double qs = c1 ? q1 : q2;
        const bool live = (c1 || c2) && qs != 0.0;                       // others: benign operand, result unused
        const double ms = __dsqrt_rn(pick(live, qs, 1.0));
        const double sp = qs == 0.0 ? 0.0 : ddiv(pick(live, ms, 1.0), kStepSize);
        if (c1) { s.p[kOpp1].tx = v1x; s.p[kOpp1].ty = v1y; s.p[kOpp1].sp = sp; }
        if (c2) { s.p[kOpp2].tx = v2x; s.p[kOpp2].ty = v2y; s.p[kOpp2].sp = sp; }
    }
}

__device__ __forceinline__ bool player_out(const Row &o)
{   // out, :574-577
    return (o.x < 0.0 || o.x > kFieldLen) || (o.y < 0.0 || o.y > kFieldWid);
}

struct StepResult { double reward; int done; int flags; };

// FutbolEnv.step, :628-717.  `ai_action` in 0..15.  `rng_col`: this thread's column of the shared-memory
// draw buffer (StepRng).  RANDOM_OPP = the constructor's random_opp (:138), a compile-time variant so that each
// kernel carries only its own opponent code.
template <bool RANDOM_OPP>
__device__ __forceinline__ StepResult v0_step(V0State &s, const V0Params &P, uint32_t env_id, int ai_action,
                                              uint32_t *rng_col)
{
    V0Rng rng;
    rng.begin(rng_col, P.key, env_id, kStreamDynamics, s.t_total);
    PendingShot shot;
    shot.shooter = -1; shot.target_y = 0; shot.pick_idx = 0;

    // pre-step snapshot used by the reward (:630-635).  The owner one-hot row of the observation is all
    // zeros between reset() and the end of the first step, otherwise 10 * onehot(owner).
    const bool fresh = s.ep_step == 0;
    const bool pre_ai1 = !fresh && s.owner == kAI1, pre_ai2 = !fresh && s.owner == kAI2;
    const bool pre_none = !fresh && s.owner == kNoOne;
    // everything _get_reward reads from the snapshot is a comparison of pre-step values: evaluate them now
    const bool far1 = sqsum(dsub(s.b.x, s.p[kAI1].x), dsub(s.b.y, s.p[kAI1].y)) > kSqLe2;   // distance > 2, :757, :790
    const bool far2 = sqsum(dsub(s.b.x, s.p[kAI2].x), dsub(s.b.y, s.p[kAI2].y)) > kSqLe2;   // :758, :799
    const bool own_forward_pass = s.b.tx > s.b.ty && s.b.tx > 0.0 && s.b.x > s.p[kAI1].x && s.b.x > s.p[kAI2].x && pre_none;  // :831 (Q6)

    if (RANDOM_OPP) {                                                    // :639-645
        const int r = (int)__umulhi(rng.take<false>(), 16u);             // randint(0, 15)
        player_turn<kOpp1>(s, rng, P, s.owner == kOpp1, r >> 2, false, 0, 0, shot);
        player_turn<kOpp2>(s, rng, P, s.owner == kOpp2, r & 3, false, 0, 0, shot);
    } else {
        opp_team(s, rng, P, shot);                                       // :649
    }
    const int action1 = ai_action >> 2, action2 = ai_action & 3;         // :653
    player_turn<kAI1>(s, rng, P, s.owner == kAI1, action1, false, 0, 0, shot);   // :655
    player_turn<kAI2>(s, rng, P, s.owner == kAI2, action2, false, 0, 0, shot);   // :656
    if (shot.shooter >= 0) resolve_shot(s, rng, P, shot);

#pragma unroll
    for (int i = 0; i < 4; ++i) advance(s.p[i]);                         // :661
    advance(s.b);                                                        // :663

    // ---- _get_reward, :752-861 (evaluated before the goal re-kickoff) ----
    const bool in_mouth = s.b.y > kGoalLower && s.b.y < kGoalUpper;
    const bool goal_for = s.b.x >= kFieldLen && in_mouth, goal_against = s.b.x <= 0.0 && in_mouth;   // score(), :580-583
    double reward;
    {
        const double score = goal_for ? 1000.0 : 0.0, get_scored = goal_against ? -1000.0 : 0.0;
        if (P.only_reward_goal) {
            reward = dadd(score, get_scored);                            // :857-858
        } else {
            const double running_r = (action1 == kRun || action2 == kRun) ? 2.0 : 0.0;       // :772-775
            const double adv_r = ((pre_ai1 && action2 == kRun) || (pre_ai2 && action1 == kRun)) ? 2.0 : 0.0;  // :777-781
            double bad1, bad2;
            if (!pre_ai1) bad1 = (action1 == kAssist || action1 == kShoot) ? -1.0 : ((far1 && action1 == kIntercept) ? -0.5 : 0.0);
            else bad1 = action1 == kIntercept ? -1.0 : 0.0;              // :783-794
            if (!pre_ai2) bad2 = (action2 == kAssist || action2 == kShoot) ? -1.0 : ((far2 && action1 == kIntercept) ? -0.5 : 0.0);  // Q5
            else bad2 = action2 == kIntercept ? -1.0 : 0.0;              // :796-807
            const double out_r = (player_out(s.p[kAI1]) || player_out(s.p[kAI2])) ? -0.6 : 0.0;   // :823-826
            const bool ai_owns = s.owner == kAI1 || s.owner == kAI2;
            double get_ball;
            if (ai_owns && !pre_ai1 && !pre_ai2)                          // :828-836 (Q6)
                get_ball = own_forward_pass ? kRewStolen : kRewGained;
            else if ((s.owner == kAI1 && pre_ai1) || (s.owner == kAI2 && pre_ai2))
                get_ball = kRewKept;                                 // :837-839
            else
                get_ball = 0.0;
            reward = dadd(dadd(dadd(dadd(dadd(dadd(get_ball, score), get_scored), out_r), dadd(bad1, bad2)), adv_r), running_r);  // :861
        }
    }

    StepResult res;
    res.done = 0;
    res.flags = 0;
    if (goal_for || goal_against) {                                      // :670-699
        if (s.b.x <= 0.0) s.opp_score += 1; else s.ai_score += 1;
        if (P.one_goal_end) res.done = 1;
        kickoff(s);
        res.flags |= kFlagGoal;
    }
    {   // out_of_field + fix, :621-625, :587-604, :701-707 (Q7, Q8)
        const bool x_out = s.b.x < 0.0 || s.b.x > kFieldLen, y_out = s.b.y < 0.0 || s.b.y > kFieldWid;
        const bool y_score = s.b.y > kGoalLower - 2 && s.b.y < kGoalUpper + 2;
        if ((x_out && !y_score) || y_out) {
            const int new_owner = (s.last_owner == kOpp1 || s.last_owner == kOpp2) ? kAI1 : kOpp1;
            s.b.x = s.b.x < 0.0 ? 0.0 : (s.b.x > kFieldLen ? kFieldLen : s.b.x);   // lock_in, :68-74
            s.b.y = s.b.y < 0.0 ? 0.0 : (s.b.y > kFieldWid ? kFieldWid : s.b.y);
            zero_motion(s.b);
            s.owner = new_owner;
            if (new_owner == kAI1) s.p[kAI1] = s.b; else s.p[kOpp1] = s.b;
            if (P.one_goal_end) res.done = 1;
            res.flags |= kFlagFix;
        }
    }
    if (s.ep_step >= P.ep_limit) res.done = 1;                           // :712-713 (`time >= game_time`)
    s.ep_step += 1;                                                      // :716
    s.t_total += 1;
    if (res.done) res.flags |= kFlagDone;
    res.reward = reward;
    return res;
}

// observation element k (0..29) of the (6,5) array the reference returns (:717)
__device__ __forceinline__ double obs_elem_owner(const V0State &s, int idx)
{   // ball_owner_array_update, :720-736; all zeros right after reset (:223)
    return (s.ep_step != 0 && s.owner == idx) ? 10.0 : 0.0;
}

template <typename F>
__device__ __forceinline__ void for_each_obs(const V0State &s, F f)
{
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        f(r * 5 + 0, s.p[r].x); f(r * 5 + 1, s.p[r].y); f(r * 5 + 2, s.p[r].tx); f(r * 5 + 3, s.p[r].ty); f(r * 5 + 4, s.p[r].sp);
    }
    f(20, s.b.x); f(21, s.b.y); f(22, s.b.tx); f(23, s.b.ty); f(24, s.b.sp);
#pragma unroll
    for (int i = 0; i < 5; ++i) f(25 + i, obs_elem_owner(s, i));
}

}  // namespace futbol
