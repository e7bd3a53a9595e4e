// Device-side v0 environment step: one thread owns one environment, all 25 state doubles live in
// registers.  Replaces gym_futbol/envs/futbol_env.py FutbolEnv.step (:628-717) and everything it
// calls, including Easy_Agent.get_action_type (gym_futbol/envs/easy_agent.py:53-98).
//
// Arithmetic contract ("kernel arithmetic"): fp64, the reference's operation order, no FMA
// contraction (every product and sum goes through __dmul_rn/__dadd_rn/..., which the compiler never
// fuses), IEEE sqrt/div.  Numpy-scalar x**2 is computed as x*x, and log/sin/cos are the fully
// specified fm_log/fm_sincos below (fdlibm-style polynomials, individually rounded IEEE operations in a
// fixed order), so that a CPU restatement of the same specification agrees BIT for bit.
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
constexpr int kV0PreBlocks = 2;                         // 8 draws cover every non-shoot step (>99.9 %)

struct Row { double x, y, tx, ty, sp; };

struct V0Params {
    uint64_t seed;
    uint32_t env_id_offset;
    int n_envs;
    int random_opp, one_goal_end, only_reward_goal, auto_reset;
    int ep_limit;       // first ep_step at which `time >= game_time` holds (400 for game_time 40)
    int shoot_speed;
    double player_speed;
};

struct V0State {
    Row p[4];           // ai_1, ai_2, opp_1, opp_2
    Row b;              // ball
    uint64_t t_total;
    int ep_step, ai_score, opp_score, owner, last_owner;
};

typedef StepRng<kV0PreBlocks> V0Rng;

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double hyp(double vx, double vy) { return __dsqrt_rn(dadd(dmul(vx, vx), dmul(vy, vy))); }

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

__device__ __forceinline__ void fm_sincos(double x, double &sn, double &cs)
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

// bigger_than(x1, x2, 2), :76-82
__device__ __forceinline__ int within_count(double d1, double d2)
{
    return (d1 <= 2.0 && d2 <= 2.0) ? 2 : ((d1 > 2.0 && d2 > 2.0) ? 0 : 1);
}

// defence_near, :280-289, with the stale-view behaviour: the shooter's own position is
// the frozen kickoff spot, except for hard-coded opponents whose views are refreshed every step.
template <int AGENT>
__device__ __forceinline__ int defence_near(const V0State &s, const V0Params &P)
{
    constexpr bool right = AGENT >= kOpp1;
    double vx = (AGENT == kAI1 || AGENT == kAI2) ? kFieldLen / 2 - 9 : kFieldLen / 2 + 9;
    double vy = (AGENT == kAI1 || AGENT == kOpp1) ? kFieldWid / 2 + 5 : kFieldWid / 2 - 5;
    if (right && !P.random_opp) { vx = s.p[AGENT].x; vy = s.p[AGENT].y; }
    const Row &d1 = right ? s.p[kAI1] : s.p[kOpp1];
    const Row &d2 = right ? s.p[kAI2] : s.p[kOpp2];
    return within_count(hyp(dsub(d1.x, vx), dsub(d1.y, vy)), hyp(dsub(d2.x, vx), dsub(d2.y, vy)));
}

// screw_vec, :101-116.  The reference draws 10 normals (np.random.normal(0, accuracy, 10), :103) and
// then an index (randint(0, 9), :107).  By specification (oracle/philox.py) the c-th normal() call of a
// step consumes no sequential draws: slot k lives in Philox block 0x8000 + 8c + (k >> 1), words
// 2(k&1), 2(k&1)+1, so only the block of the indexed slot is evaluated here.
__device__ __forceinline__ void screw_vec(V0Rng &rng, double vx, double vy, double mag, double accuracy,
                                          double &ox, double &oy)
{
    const uint32_t call = rng.normal_calls++;
    const uint32_t pick = (uint32_t)rng.randint(0, 9);                   // :107
    const Philox4 nb = philox_step_block(rng.seed, rng.env_id, rng.stream, rng.t, kNormalBlock0 + 8u * call + (pick >> 1));
    const uint32_t w0 = (pick & 1u) ? nb.z : nb.x, w1 = (pick & 1u) ? nb.w : nb.y;
    const double u1 = (double)((w0 >> 8) + 1u) * (1.0 / 16777216.0);
    const double u2 = (double)(w1 >> 8) * (1.0 / 16777216.0);
    double bm_sin, bm_cos;
    fm_sincos(dmul(6.283185307179586, u2), bm_sin, bm_cos);
    const double z = dmul(__dsqrt_rn(dmul(-2.0, fm_log(u1))), bm_cos);
    const double nd = dadd(0.0, dmul(accuracy, z));                      // np.random.normal(0, accuracy)
    const double c = ddiv(dmul(vx, 1.0), mag), sn = ddiv(dmul(vy, 1.0), mag);  // :105-106
    const double swing = dmul(ddiv(nd, 180.0), 3.141592653589793);       // :108
    double ss, sc;
    fm_sincos(swing, ss, sc);                                            // :109-110
    const double tc = dsub(dmul(c, sc), dmul(sn, ss));                   // :113
    const double ts = dadd(dmul(sn, sc), dmul(c, ss));                   // :114
    ox = dmul(tc, mag);                                                  // :115
    oy = dmul(ts, mag);
}

// intercept_chance(d, 1, 2), :122-129
__device__ __forceinline__ double intercept_chance(double d)
{
    if (d < 1.0) return 0.9;
    if (d <= 2.0) return dmul(ddiv(0.9, dsub(1.0, 2.0)), dsub(d, 2.0));
    return 0.0;
}

// _set_vector_observation, :300-530, for one player.
template <int AGENT>
__device__ __forceinline__ void set_vector_observation(V0State &s, V0Rng &rng, const V0Params &P, bool has_ball,
                                                       int action, bool set_target, double tgx, double tgy)
{
    constexpr bool right = AGENT >= kOpp1;
    constexpr double goal_x = right ? 0.0 : kFieldLen;
    Row &ao = s.p[AGENT];
    const double target_y = (double)rng.randint(32, 36);                 // :306 -- always drawn first

    if (has_ball) {
        if (action == kIntercept) {                                      // :318-321
            zero_motion(ao);
            zero_motion(s.b);
        } else if (action == kRun) {                                     // :330-356
            ao.sp = P.player_speed;
            if (set_target) { ao.tx = tgx; ao.ty = tgy; }
            else { ao.tx = dsub(goal_x, ao.x); ao.ty = dsub(target_y, ao.y); }
            if (rng.random() < 0.05) s.owner = kNoOne;                   // :353-354 (Q3)
            else s.b = ao;                                               // :356
        } else if (action == kShoot) {                                   // :362-383
            const double accuracy = dadd(10.0, dmul((double)defence_near<AGENT>(s, P), 20.0));  // :364
            s.b.sp = (double)rng.randint(P.shoot_speed - 16, P.shoot_speed);                      // :367
            const double vx = dsub(goal_x, s.b.x), vy = dsub(target_y, s.b.y);
            screw_vec(rng, vx, vy, hyp(vx, vy), accuracy, s.b.tx, s.b.ty);                        // :373-378
            s.last_owner = s.owner;                                      // :381
            s.owner = kNoOne;                                            // :382
            zero_motion(ao);                                             // :383
        } else {                                                         // assist, :385-423
            const Row &mate = s.p[AGENT ^ 1];
            const double vx = dsub(mate.x, s.b.x), vy = dsub(mate.y, s.b.y);   // :412
            double sp = ddiv(hyp(vx, vy), kStepSize);                    // :413
            if (sp > 20.0) sp = 20.0;                                    // :414-415
            const double lo = dsub(sp, 1.0), hi = dadd(sp, 1.0);
            s.b.sp = dadd(lo, dmul(dsub(hi, lo), rng.random()));         // random.uniform, :416
            s.b.tx = vx; s.b.ty = vy;                                    // :418
            s.last_owner = s.owner;                                      // :421
            s.owner = kNoOne;                                            // :422
            zero_motion(ao);                                             // :423
        }
    } else {
        const double bx = dsub(s.b.x, ao.x), by = dsub(s.b.y, ao.y);     // :432
        if (action == kIntercept) {                                      // :452-476 (Q4: player keeps moving)
            const double mag = hyp(bx, by);
            const bool success = rng.random() < intercept_chance(mag);   // :459 -- drawn even when far
            if (success || (s.owner == kNoOne && mag < 4.0)) {           // :462-463
                s.b = ao;                                                // :465-466
                s.last_owner = s.owner;                                  // :467
                s.owner = AGENT;                                         // :468
            }
        } else if (action == kRun) {                                     // :483-503
            ao.sp = P.player_speed;
            if (set_target) { ao.tx = tgx; ao.ty = tgy; }
            else if (s.owner != AGENT) { ao.tx = bx; ao.ty = by; }       // :501
            else { ao.tx = dsub(goal_x, ao.x); ao.ty = dsub(kFieldWid / 2, ao.y); }  // :503 (dead, Q12)
        } else {                                                         // shoot / assist without the ball, :509-525
            zero_motion(ao);
        }
    }
}

// Easy_Agent.get_action_type for a 'right' opponent, easy_agent.py:53-98
template <int AGENT>
__device__ __forceinline__ int easy_action(const V0State &s, V0Rng &rng, bool has_ball, bool team_has_ball)
{
    const Row &ao = s.p[AGENT];
    const Row &mo = s.p[AGENT ^ 1];
    if (has_ball) {
        if (ao.x <= 20.0) return kShoot;                                 // :77-79, shoot_x = 0 + 20
        if (mo.x < ao.x || mo.y < dsub(ao.y, 7.0) || mo.y > dadd(ao.y, 7.0)) {   // :81-83 (short-circuit order)
            if (rng.random() > 0.8 && hyp(dsub(mo.x, ao.x), dsub(mo.y, ao.y)) > 12.0) return kAssist;
        }
        return kRun;
    }
    if (!team_has_ball && hyp(dsub(s.b.x, ao.x), dsub(s.b.y, ao.y)) <= 1.0) return kIntercept;  // :90-92
    return kRun;
}

// _step_by_observation, :560-571 (DECELERATION = 0: the ball's speed update is `sp -= 0.0`)
__device__ __forceinline__ void advance(Row &o)
{
    const double mag = hyp(o.tx, o.ty);                                  // :562
    if (mag != 0.0) {
        o.x = dadd(o.x, dmul(o.sp, ddiv(dmul(o.tx, kStepSize), mag)));   // :567
        o.y = dadd(o.y, dmul(o.sp, ddiv(dmul(o.ty, kStepSize), mag)));   // :568
    }
}

// _opp_team_set_vector_observation, :864-982
__device__ __forceinline__ void opp_team(V0State &s, V0Rng &rng, const V0Params &P)
{
    const bool has1 = s.owner == kOpp1, has2 = s.owner == kOpp2;         // :866-877
    const bool team_has = has1 || has2;
    const int a1 = easy_action<kOpp1>(s, rng, has1, team_has);           // :879
    const int a2 = easy_action<kOpp2>(s, rng, has2, team_has);           // :880
    int a1_type = a1, a2_type = a2;                                      // overrides do not touch a1 / a2
    bool set1 = false, set2 = false;
    double t1x = 0, t1y = 0, t2x = 0, t2y = 0;
    const Row &o1 = s.p[kOpp1], &o2 = s.p[kOpp2];
    const bool diag1 = o1.y > dmul(kFieldWid, 0.2), diag2 = o2.y < dmul(kFieldWid, 0.8);
    if (has1 && a1 == kRun) {                                            // :893-909
        if (diag1) { set1 = true; t1x = -1; t1y = -1; }
        if (a2 == kRun && o2.x > dmul(kFieldLen, 0.1) && diag2) { set2 = true; t2x = -1; t2y = 1; }
    }
    if (has2 && a2 == kRun) {                                            // :911-928
        if (diag2) { set2 = true; t2x = -1; t2y = 1; }
        if (a1 == kRun && o1.x > dmul(kFieldLen, 0.1) && diag1) { set1 = true; t1x = -1; t1y = -1; }
    }
    if ((s.owner == kAI1 || s.owner == kAI2) && s.b.x < dmul(kFieldLen, 0.6)) {   // :931-947
        const double dpx = dmul(kFieldLen, 0.75), dpy = dmul(kFieldWid, 0.5);
        if (o1.x > o2.x) { a1_type = kRun; set1 = true; t1x = dsub(dpx, o1.x); t1y = dsub(dpy, o1.y); }
        else             { a2_type = kRun; set2 = true; t2x = dsub(dpx, o2.x); t2y = dsub(dpy, o2.y); }
    }
    set_vector_observation<kOpp1>(s, rng, P, has1, a1_type, set1, t1x, t1y);   // :951-954
    set_vector_observation<kOpp2>(s, rng, P, has2, a2_type, set2, t2x, t2y);   // :956-959
    if (s.owner == kNoOne && a1 == kRun && a2 == kRun) {                 // :962-982 anticipate the ball
        Row nb = s.b;
        advance(nb);
        const double v1x = dsub(nb.x, s.p[kOpp1].x), v1y = dsub(nb.y, s.p[kOpp1].y);
        const double v2x = dsub(nb.x, s.p[kOpp2].x), v2y = dsub(nb.y, s.p[kOpp2].y);
        const double m1 = hyp(v1x, v1y), m2 = hyp(v2x, v2y);
        const double reach = dmul(kStepSize, P.player_speed);
        if (m1 < reach) { s.p[kOpp1].tx = v1x; s.p[kOpp1].ty = v1y; s.p[kOpp1].sp = ddiv(m1, kStepSize); }
        else if (m2 < reach) { s.p[kOpp2].tx = v2x; s.p[kOpp2].ty = v2y; s.p[kOpp2].sp = ddiv(m2, kStepSize); }
    }
}

__device__ __forceinline__ bool player_out(const Row &o)
{   // out, :574-577
    return (o.x < 0.0 || o.x > kFieldLen) || (o.y < 0.0 || o.y > kFieldWid);
}

struct StepResult { double reward; int done; int flags; };

// FutbolEnv.step, :628-717.  `ai_action` in 0..15.
__device__ __forceinline__ StepResult v0_step(V0State &s, const V0Params &P, uint32_t env_id, int ai_action)
{
    V0Rng rng;
    rng.begin(P.seed, env_id, kStreamDynamics, s.t_total);

    // pre-step snapshot used by the reward (:630-635).  The owner one-hot row of the observation is all
    // zeros between reset() and the end of the first step, otherwise 10 * onehot(owner).
    const bool fresh = s.ep_step == 0;
    const double ob_x = s.b.x, ob_tx = s.b.tx, ob_ty = s.b.ty, ob_y = s.b.y;
    const double o1_x = s.p[kAI1].x, o1_y = s.p[kAI1].y, o2_x = s.p[kAI2].x, o2_y = s.p[kAI2].y;
    const bool pre_ai1 = !fresh && s.owner == kAI1, pre_ai2 = !fresh && s.owner == kAI2;
    const bool pre_none = !fresh && s.owner == kNoOne;

    if (P.random_opp) {                                                  // :639-645
        const int r = rng.randint(0, 15);
        set_vector_observation<kOpp1>(s, rng, P, s.owner == kOpp1, r >> 2, false, 0, 0);
        set_vector_observation<kOpp2>(s, rng, P, s.owner == kOpp2, r & 3, false, 0, 0);
    } else {
        opp_team(s, rng, P);                                             // :649
    }
    const int action1 = ai_action >> 2, action2 = ai_action & 3;         // :653
    set_vector_observation<kAI1>(s, rng, P, s.owner == kAI1, action1, false, 0, 0);   // :655
    set_vector_observation<kAI2>(s, rng, P, s.owner == kAI2, action2, false, 0, 0);   // :656

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
            const double d1 = hyp(dsub(ob_x, o1_x), dsub(ob_y, o1_y));   // :757
            const double d2 = hyp(dsub(ob_x, o2_x), dsub(ob_y, o2_y));   // :758
            const double running_r = (action1 == kRun || action2 == kRun) ? 2.0 : 0.0;       // :772-775
            const double adv_r = ((pre_ai1 && action2 == kRun) || (pre_ai2 && action1 == kRun)) ? 2.0 : 0.0;  // :777-781
            double bad1, bad2;
            if (!pre_ai1) bad1 = (action1 == kAssist || action1 == kShoot) ? -1.0 : ((d1 > 2.0 && action1 == kIntercept) ? -0.5 : 0.0);
            else bad1 = action1 == kIntercept ? -1.0 : 0.0;              // :783-794
            if (!pre_ai2) bad2 = (action2 == kAssist || action2 == kShoot) ? -1.0 : ((d2 > 2.0 && action1 == kIntercept) ? -0.5 : 0.0);  // Q5
            else bad2 = action2 == kIntercept ? -1.0 : 0.0;              // :796-807
            const double out_r = (player_out(s.p[kAI1]) || player_out(s.p[kAI2])) ? -0.6 : 0.0;   // :823-826
            const bool ai_owns = s.owner == kAI1 || s.owner == kAI2;
            double get_ball;
            if (ai_owns && !pre_ai1 && !pre_ai2)                          // :828-836 (Q6)
                get_ball = (ob_tx > ob_ty && ob_tx > 0.0 && ob_x > o1_x && ob_x > o2_x && pre_none) ? dmul(-50.0, 0.3) : dmul(60.0, 0.3);
            else if ((s.owner == kAI1 && pre_ai1) || (s.owner == kAI2 && pre_ai2))
                get_ball = dmul(30.0, 0.3);                              // :837-839
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
