// Device-side v0 environment step: one thread owns one environment.  Replaces
// gym_futbol/envs/futbol_env.py FutbolEnv.step (:628-717) and everything it calls, including
// Easy_Agent.get_action_type (gym_futbol/envs/easy_agent.py:53-98).
//
// Arithmetic contract ("kernel arithmetic"): fp64, the reference's operation order, no FMA
// contraction (every product and sum goes through __dmul_rn/__dadd_rn/..., which the compiler never
// fuses), IEEE sqrt/div.  Numpy-scalar x**2 is computed as x*x, and log/sin/cos are the fully
// specified fm_log/fm_sincos below (fdlibm-style polynomials, individually rounded IEEE operations in a
// fixed order), so that a CPU restatement of the same specification agrees BIT for bit.
//
// Shape of the code (each choice is an ncu finding, profiles/).
//  * Branch-lean.  The reference picks one of eight branches per player (has-ball x action); the 32
//    environments of a warp pick 32 different ones, so a branchy transcription runs the union of all
//    branches at ~13 active lanes.  Here a player turn is ONE straight-line block: the case is classified
//    into predicates, the single vector magnitude any case needs is computed once, results are committed
//    with predicated stores.  What stays behind a branch is rare: the kick (screw_vec: log, sin, cos) --
//    deferred to one site per step because at most one player can shoot per step.
//  * Small.  The step is instruction-FETCH bound when unrolled (four inlined player turns + six inlined
//    kinematics updates = 58 KB of SASS: SM I-cache hit rate 70 %, GPC instruction-cache requests at 95 %
//    of peak).  The 25 state doubles of an environment therefore live in shared memory, one column per
//    lane (conflict-free), which makes rows addressable by a RUN-TIME index: one copy of the turn code
//    looped over the four players, one copy of the kinematics code looped over the five rows.
//  * Threshold tests on a distance (d <= 2, d > 12, ...) are done on the squared distance with the
//    exactly equivalent bound (kSq* below): no square root, not a single decision changed.
#pragma once
#include <stdint.h>
#include "philox.cuh"
#include "ieee_fast.cuh"

namespace futbol {

// action.py:3-6 / ballowner.py:3-7
enum : int { kRun = 0, kIntercept = 1, kShoot = 2, kAssist = 3 };
enum : int { kAI1 = 0, kAI2 = 1, kOpp1 = 2, kOpp2 = 3, kNoOne = 4 };
enum : int { kFlagGoal = 1, kFlagFix = 2, kFlagDone = 4 };

constexpr double kFieldLen = 105.0, kFieldWid = 68.0;   // futbol_env.py:18-19
constexpr double kGoalLower = 29.0, kGoalUpper = 39.0;  // :23-24
constexpr double kStepSize = 0.1;                       // :45
constexpr double kNaN = __builtin_nan("");

// For s = fl(fl(dx*dx) + fl(dy*dy)) and d = sqrt_rn(s) (correctly rounded, hence monotone):
//   d <= 1.0  <=>  s <= nextafter(1, +inf)
//   d <= 2.0  <=>  s <= nextafter(4, +inf)
//   d > 12.0  <=>  s >  144.0
// (sqrt(s) rounds to c or below exactly when it lies below the midpoint of c and its successor; squaring
// that midpoint gives the bound.  tests/test_sqrt_thresholds.py checks every neighbouring double.)
//   d <  1.0  <=>  s <  1.0          d <  4.0  <=>  s <  16.0      (the predecessor of c*c has a root below c)
constexpr double kSqLe1 = 0x1.0000000000001p+0;
constexpr double kSqLe2 = 0x1.0000000000001p+2;
constexpr double kSqGt12 = 144.0;
constexpr double kSqLt1 = 1.0, kSqLt4 = 16.0;
// random() = k * 2^-24 (k = w >> 8, exact in double):  random() < 0.9  <=>  k <= 15099494;  random() < 0.05  <=>  k <= 838860;
// random() > 0.8  <=>  k >= 13421773
constexpr uint32_t kULt0p9 = 15099494u, kULt0p05 = 838860u, kUGt0p8 = 13421773u;
// products of reference constants, rounded once as the reference's float64 multiply rounds them
// (__dmul_rn is not folded by the compiler; tests/test_sqrt_thresholds.py::test_constant_products)
constexpr double kWid02 = 13.600000000000001;   // width * 0.2, :895
constexpr double kWid08 = 54.400000000000006;   // width * 0.8, :913
constexpr double kLen01 = 10.5;                 // length * 0.1, :903
constexpr double kLen06 = 63.0;                 // length * 0.6, :931
constexpr double kDefendX = 78.75, kDefendY = 34.0;   // (length * 0.75, width * 0.5), :933
constexpr double kRewStolen = -15.0, kRewGained = 18.0, kRewKept = 9.0;   // -50*0.3, 60*0.3, 30*0.3, :832-839

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

// ---- per-environment working storage in shared memory ---------------------------------------------------
// One block of kWarpSmemBytes per warp; element k of lane l of a section at section[k * kLanes + l].
//   double  st[25][kLanes]   k = 5 * row + field, rows ai_1, ai_2, opp_1, opp_2, ball; fields x, y, tx, ty,
//                            speed (= observation rows 0-4)
//   then, overlapping in time:  uint32 draws[kDrawWords][kLanes]   (during the step: the Philox words)
//                               float  stage[kLanes * 30]          (observation staging, after the step)
constexpr int kStateWords = 25;
constexpr int kObsDim = 30;
constexpr int kWarpStateBytes = kStateWords * kLanes * 8;
constexpr int kStepScratchBytes = kDrawWords * kLanes * 4;
constexpr int kWarpScratchBytes = (kLanes * kObsDim * 4 > kStepScratchBytes) ? kLanes * kObsDim * 4 : kStepScratchBytes;
constexpr int kWarpSmemBytes = kWarpStateBytes + kWarpScratchBytes;
constexpr int kWarpSmemBytesDense = kWarpStateBytes + kStepScratchBytes;   // 7936 B: seven 128-thread blocks per SM
constexpr int kX = 0, kY = kLanes, kTX = 2 * kLanes, kTY = 3 * kLanes, kSP = 4 * kLanes;   // field offsets in a row
constexpr int kRowStride = 5 * kLanes;
constexpr int kBallRow = 4;

#ifndef FUTBOL_HOST_SHIM
extern __shared__ __align__(16) unsigned char futbol_smem[];   // dynamic: (threads / 32) * kWarpSmemBytes
#else
static unsigned char futbol_smem[kWarpSmemBytes] __attribute__((aligned(16)));
#endif

// This lane's columns, as offsets into futbol_smem (an offset, unlike a pointer, keeps its address space
// through the out-of-line helpers below: every access stays an LDS/STS).
struct Lane {
    uint32_t st;   // index (in doubles) of state element 0
    uint32_t dw;   // index (in uint32) of draw word 0
    __device__ __forceinline__ double &f(int k) const { return reinterpret_cast<double *>(futbol_smem)[st + k]; }
    __device__ __forceinline__ uint32_t &draw(uint32_t j) const { return reinterpret_cast<uint32_t *>(futbol_smem)[dw + j * kLanes]; }
};

// warp_bytes: shared memory per warp of the calling kernel (kWarpSmemBytes, or kWarpSmemBytesDense for the rollout
// kernels that stage the observation in three passes through the draw words' area)
__device__ __forceinline__ Lane make_lane(int warp_in_block, int lane, int warp_bytes = kWarpSmemBytes)
{
    Lane L;
    L.st = (uint32_t)(warp_in_block * (warp_bytes / 8) + lane);
    L.dw = (uint32_t)(warp_in_block * (warp_bytes / 4) + kWarpStateBytes / 4 + lane);
    return L;
}

// the scalar part of an environment's state (registers)
struct V0Regs {
    uint64_t t_total;
    int ep_step, ai_score, opp_score, owner, last_owner;
};

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
// c ? a : b that the optimiser cannot look through.  Used where a lane is handed a benign operand to keep
// it on the fast path of the IEEE sqrt/div sequence: with a plain ternary the compiler rewrites
// sqrt(c ? q : 1.0) into c ? sqrt(q) : 1.0 and the zero operand is back (seen in ncu as 1-6 lanes per warp
// inside __cuda_sm20_dsqrt_rn_f64_mediumpath / div_rn_f64_full on every step).  The hot-path operations now use
// the guard-free sequences of ieee_fast.cuh, for which the benign operand is a REQUIREMENT (fsqrt(0) is not 0).
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
    const double s = fdiv(f, dadd(2.0, f));            // f is +0 or 2^-24 <= |f| <= 0.42, the divisor in [1.7, 2.42]
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
    // quadrant n & 3:  0: (sr, cr)   1: (cr, -sr)   2: (-sr, -cr)   3: (-cr, sr)
    const bool odd = (n & 1) != 0;
    const double s0 = odd ? cr : sr, c0 = odd ? sr : cr;
    sn = (n & 2) ? -s0 : s0;
    cs = ((n + 1) & 2) ? -c0 : c0;
}

// kickoff formation, futbol_env.py:211-223 (also the goal re-kickoff :684-692).  Out of line: three call
// sites (goal, reset, padding lanes), 25 stores.
static __device__ __noinline__ void kickoff_rows(Lane L)
{
    const double px[5] = {kFieldLen / 2 - 9, kFieldLen / 2 - 9, kFieldLen / 2 + 9, kFieldLen / 2 + 9, kFieldLen / 2};
    const double py[5] = {kFieldWid / 2 + 5, kFieldWid / 2 - 5, kFieldWid / 2 + 5, kFieldWid / 2 - 5, kFieldWid / 2};
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        L.f(r * kRowStride + kX) = px[r]; L.f(r * kRowStride + kY) = py[r];
        L.f(r * kRowStride + kTX) = 0.0; L.f(r * kRowStride + kTY) = 0.0; L.f(r * kRowStride + kSP) = 0.0;
    }
}

__device__ __forceinline__ void kickoff(Lane L, V0Regs &s)
{
    kickoff_rows(L);
    s.owner = kNoOne;
    s.last_owner = kNoOne;
}

__device__ __forceinline__ void reset_env(Lane L, V0Regs &s)
{   // FutbolEnv.reset, :205-245.  t_total (the Philox step index) is deliberately kept.
    kickoff(L, s);
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
// The same holds for a pass (assist): it needs the ball too, so a step has at most ONE pending kick of either kind.  The
// pass's speed (|mate - ball| / 0.1 capped at 20, +- 1 uniform, :413-416: a square root and a division) is resolved at that
// same site from the positions, which do not change before the kinematics phase, and the draw taken at the turn.
struct PendingShot { int shooter; int target_y; uint32_t pick_idx; int passer; uint32_t pass_w; };

// Draw budget of a step (sequential draws; the kick's normal() slots are addressed separately).  In turn
// order: [random_opp: 1 | hard-coded: <= 1, only the opponent holding the ball can draw in get_action_type],
// then per player 1 (target_y) + 1 if it draws (+1 more for the single possible shooter): at most
// 1 + 4 + 4 + 1 = 10 draws, all inside the kDrawWords = 12 words generated per step, so a draw is a bare LDS.

// _set_vector_observation, :300-530, for player `a` (run-time index): straight-line, predicated.
__device__ __forceinline__ void player_turn(Lane L, uint32_t &j, V0Regs &s, const V0Params &P, int a, bool has_ball,
                                            int action, bool set_target, double tgx, double tgy, PendingShot &shot)
{
    const int ao = a * kRowStride, mo = (a ^ 1) * kRowStride, bo = kBallRow * kRowStride;
    const double ax = L.f(ao + kX), ay = L.f(ao + kY), atx = L.f(ao + kTX), aty = L.f(ao + kTY), asp = L.f(ao + kSP);
    const double mx = L.f(mo + kX), my = L.f(mo + kY);
    const double ball_x = L.f(bo + kX), ball_y = L.f(bo + kY);
    const double goal_x = a >= kOpp1 ? 0.0 : kFieldLen;                   // the goal this player attacks

    const int target_y = 32 + (int)__umulhi(L.draw(j), 5u);              // randint(32, 36), :306 -- always drawn first
    j += 1;
    const bool is_run = action == kRun, is_int = action == kIntercept;
    const bool hb_run = has_ball && is_run, hb_int = has_ball && is_int;
    const bool hb_shoot = has_ball && action == kShoot, hb_assist = has_ball && action == kAssist;
    const bool nb_int = !has_ball && is_int;
    // one more draw in: has-ball run (:353), shoot (:367), assist (:416); no-ball intercept (:459, even when far)
    const uint32_t w = L.draw(j);
    j += (has_ball != is_int) ? 1u : 0u;
    // random() = k * 2^-24 with k = w >> 8: the fixed thresholds are integer compares on k (kU* below, exact:
    // tests/test_sqrt_thresholds.py); the double is only formed where a computed chance is compared
    const uint32_t uk = w >> 8;

    const double bx = dsub(ball_x, ax), by = dsub(ball_y, ay);           // :432
    const double vx = hb_assist ? dsub(mx, ball_x) : bx;                 // :412
    const double vy = hb_assist ? dsub(my, ball_y) : by;
    // no-ball intercept, :452-463: intercept_chance(d, 1, 2) (:122-129, k = 0.9 / (1 - 2) = -0.9 exactly) is 0.9 below 1,
    // 0 above 2, and an unowned ball within 4 is taken whatever the chance: all decided on the SQUARED distance (exact
    // bounds above).  The root itself is needed only for an owned ball at 1 <= d <= 2 -- about one turn in a hundred --
    // and is taken behind a branch that most warps skip (it was a full sqrt per player per step: ~45 instructions).
    const double q = sqsum(bx, by);
    const bool lt1 = q < kSqLt1, le2 = q <= kSqLe2, free_ball = s.owner == kNoOne;
    const bool mid = nb_int && !lt1 && le2 && !free_ball;
    bool take_mid = false;
    if (mid) {                                                           // q in [1, 4]: inside the guard-free sqrt's domain
        const double chance_mid = dmul(-0.9, dsub(fsqrt(q), 2.0));
        take_mid = (double)uk * (1.0 / 16777216.0) < chance_mid;
    }
    const bool take = nb_int && ((free_ball && q < kSqLt4) || (lt1 ? uk <= kULt0p9 : take_mid));
    // has-ball run, :353-356 (Q3)
    const bool drop = hb_run && uk <= kULt0p05;
    const bool carry = hb_run && !drop;

    // the player's own row after the turn (Q4: a no-ball intercept keeps the previous vector and speed;
    // :330-352, :483-503 run (:503 unreachable, Q12); everything else stops :318-321, :383, :423, :509-525)
    const double run_tx = set_target ? tgx : (has_ball ? dsub(goal_x, ax) : bx);
    const double run_ty = set_target ? tgy : (has_ball ? dsub((double)target_y, ay) : by);
    const double rtx = nb_int ? atx : (is_run ? run_tx : 0.0);
    const double rty = nb_int ? aty : (is_run ? run_ty : 0.0);
    const double rsp = nb_int ? asp : (is_run ? P.player_speed : 0.0);
    if (!nb_int) { L.f(ao + kTX) = rtx; L.f(ao + kTY) = rty; L.f(ao + kSP) = rsp; }
    // ball and possession
    if (carry || take) { L.f(bo + kX) = ax; L.f(bo + kY) = ay; }         // ball row := player row, :356, :465-466
    if (carry || take || hb_int || hb_assist) {                          // :321 stop, :416-418 pass
        L.f(bo + kTX) = hb_int ? 0.0 : (hb_assist ? vx : rtx);
        L.f(bo + kTY) = hb_int ? 0.0 : (hb_assist ? vy : rty);
    }
    if (carry || take || hb_int || hb_shoot) {                           // :367 kick speed; the vector (and a pass's speed) are resolved later
        const double kick = (double)(P.shoot_speed - 16 + (int)__umulhi(w, 17u));
        L.f(bo + kSP) = hb_int ? 0.0 : (hb_shoot ? kick : rsp);
    }
    if (hb_shoot) { shot.shooter = a; shot.target_y = target_y; shot.pick_idx = j; }
    if (hb_assist) { shot.passer = a; shot.pass_w = w; }
    j += hb_shoot ? 1u : 0u;                                             // randint(0, 9) of screw_vec, :107
    if (take) { shot.shooter = -1; shot.passer = -1; }
    if (hb_shoot || hb_assist || take) s.last_owner = s.owner;           // :381, :421, :467
    s.owner = take ? a : ((drop || hb_shoot || hb_assist) ? (int)kNoOne : s.owner);   // :354, :382, :422, :468
}

// defence_near (:280-289, with the stale-view behaviour Q1) + screw_vec (:101-116) for the pending kick.
// By specification a step's normal() call consumes no sequential draws: slot k lives in
// Philox block 0x8000 + (k >> 1), words 2(k&1), 2(k&1)+1, so only the block of the picked slot is evaluated.
__device__ __forceinline__ void resolve_shot(Lane L, const V0Regs &s, const V0Params &P, bool random_opp, uint32_t env_id,
                                             const PendingShot &shot)
{
    const int a = shot.shooter;
    const bool right = a >= kOpp1;
    // the shooter's own position is the frozen kickoff spot, except for hard-coded opponents (views refreshed)
    double px = right ? kFieldLen / 2 + 9 : kFieldLen / 2 - 9;
    double py = (a == kAI1 || a == kOpp1) ? kFieldWid / 2 + 5 : kFieldWid / 2 - 5;
    if (right && !random_opp) { px = L.f(a * kRowStride + kX); py = L.f(a * kRowStride + kY); }
    const int d1 = (right ? kAI1 : kOpp1) * kRowStride, d2 = (right ? kAI2 : kOpp2) * kRowStride;
    // bigger_than(d1, d2, 2), :76-82 = how many of the two defenders are within 2.0
    const int near = (sqsum(dsub(L.f(d1 + kX), px), dsub(L.f(d1 + kY), py)) <= kSqLe2 ? 1 : 0) +
                     (sqsum(dsub(L.f(d2 + kX), px), dsub(L.f(d2 + kY), py)) <= kSqLe2 ? 1 : 0);
    const double accuracy = dadd(10.0, dmul((double)near, 20.0));        // :364
    const int bo = kBallRow * kRowStride;
    const double vx = dsub(right ? 0.0 : kFieldLen, L.f(bo + kX)), vy = dsub((double)shot.target_y, L.f(bo + kY));   // :373-376
    // The kicker holds the ball, so the ball is on the pitch and off the goal mouth it aims at: |v| > 0 (a ball at
    // (0 | 105, 32..36) has scored, :581-582).  The guard-free sequences need just that; were it ever 0 the
    // reference would divide 0 by 0, and so does this.
    const double mag2 = sqsum(vx, vy);
    const bool aimed = mag2 != 0.0;
    const double mag = aimed ? fsqrt(pick(aimed, mag2, 1.0)) : 0.0;      // get_vec, :62-65

    const uint32_t pick_slot = __umulhi(L.draw(shot.pick_idx), 10u);     // randint(0, 9), :107
    const Philox4 nb = philox_step_block(P.key, env_id, kStreamDynamics, s.t_total, kNormalBlock0 + (pick_slot >> 1));
    const uint32_t w0 = (pick_slot & 1u) ? nb.z : nb.x, w1 = (pick_slot & 1u) ? nb.w : nb.y;
    const double u1 = (double)((w0 >> 8) + 1u) * (1.0 / 16777216.0);
    const double u2 = (double)(w1 >> 8) * (1.0 / 16777216.0);
    double bm_sin, bm_cos;
    fm_sincos(dmul(6.283185307179586, u2), bm_sin, bm_cos);
    const double m2l = dmul(-2.0, fm_log(u1));                           // -0 for u1 == 1, else >= 1.19e-7
    const double z = dmul(m2l == 0.0 ? m2l : fsqrt(pick(m2l != 0.0, m2l, 1.0)), bm_cos);
    const double nd = dadd(0.0, dmul(accuracy, z));                      // np.random.normal(0, accuracy), :103; never -0
    double c, sn;
    fdiv2(dmul(vx, 1.0), dmul(vy, 1.0), pick(aimed, mag, 1.0), c, sn);   // :105-106 (numerators: never -0)
    if (!aimed) c = sn = kNaN;
    const double swing = dmul(fdiv(nd, 180.0), 3.141592653589793);       // :108 (|nd| is 0 or > 1e-16)
    double ss, sc;
    fm_sincos(swing, ss, sc);                                            // :109-110
    const double tc = dsub(dmul(c, sc), dmul(sn, ss));                   // :113
    const double ts = dadd(dmul(sn, sc), dmul(c, ss));                   // :114
    L.f(bo + kTX) = dmul(tc, mag);                                       // :115
    L.f(bo + kTY) = dmul(ts, mag);
}

// the pending pass: ball speed = uniform(s - 1, s + 1), s = min(|mate - ball| / 0.1, 20), :413-416
__device__ __forceinline__ void resolve_pass(Lane L, const PendingShot &shot)
{
    const int mo = (shot.passer ^ 1) * kRowStride, bo = kBallRow * kRowStride;
    const double q = sqsum(dsub(L.f(mo + kX), L.f(bo + kX)), dsub(L.f(mo + kY), L.f(bo + kY)));
    const bool q_zero = q == 0.0;                                        // mate on the ball: sqrt(0) / 0.1 = 0 as in the reference
    const double mag = fsqrt(pick(!q_zero, q, 1.0));
    const double quot = fdiv(pick(!q_zero, mag, 1.0), kStepSize);
    double pass = q_zero ? 0.0 : quot;
    pass = pass > 20.0 ? 20.0 : pass;
    const double lo = dsub(pass, 1.0), hi = dadd(pass, 1.0);
    const double u = (double)(shot.pass_w >> 8) * (1.0 / 16777216.0);
    L.f(bo + kSP) = dadd(lo, dmul(dsub(hi, lo), u));                     // random.uniform
}

// the same, out of line (LAYOUT bits 0 / 1): layouts measured in profiles/r2_v0_history.md (r2d, r2e, r2g); not in the shipped kernels
static __device__ __noinline__ void resolve_pass_outlined(Lane L, const PendingShot &shot) { resolve_pass(L, shot); }
static __device__ __noinline__ void resolve_shot_outlined(Lane L, const V0Regs &s, const V0Params &P, bool random_opp, uint32_t env_id,
                                                          const PendingShot &shot) { resolve_shot(L, s, P, random_opp, env_id, shot); }

// Easy_Agent.get_action_type for 'right' opponent `a`, easy_agent.py:53-98
__device__ __forceinline__ int easy_action(Lane L, uint32_t &j, int a, bool has_ball, bool team_has_ball)
{
    const int ao = a * kRowStride, mo = (a ^ 1) * kRowStride, bo = kBallRow * kRowStride;
    const double ax = L.f(ao + kX), ay = L.f(ao + kY), mx = L.f(mo + kX), my = L.f(mo + kY);
    const bool in_range = ax <= 20.0;                                    // :77-79, shoot_x = 0 + 20
    const bool open_mate = mx < ax || my < dsub(ay, 7.0) || my > dadd(ay, 7.0);   // :81-83
    // the draw happens only when the geometric clause holds (short-circuit `and`, :81-85)
    const uint32_t w = L.draw(j);
    j += (has_ball && !in_range && open_mate) ? 1u : 0u;
    const bool lucky = (w >> 8) >= kUGt0p8;                              // random() > 0.8
    const bool far_mate = sqsum(dsub(mx, ax), dsub(my, ay)) > kSqGt12;                         // distance > 12
    const bool ball_close = sqsum(dsub(L.f(bo + kX), ax), dsub(L.f(bo + kY), ay)) <= kSqLe1;  // distance <= 1.0, :90
    const int with_ball = in_range ? (int)kShoot : ((open_mate && lucky && far_mate) ? (int)kAssist : (int)kRun);
    const int without = (!team_has_ball && ball_close) ? (int)kIntercept : (int)kRun;          // :90-96
    return has_ball ? with_ball : without;
}

// _step_by_observation, :560-571 (DECELERATION = 0: the ball's speed update is `sp -= 0.0`) for one row: reads
// row `src`, returns the new (x, y).  Straight-line: a stopped row (|t| == 0, :563) and a zero component
// (0 / mag = that same signed zero) are fed benign operands (a requirement of the guard-free sqrt/div).
__device__ __forceinline__ void advance_xy(Lane L, int src, double &xo, double &yo)
{
    const double x = L.f(src + kX), y = L.f(src + kY), tx = L.f(src + kTX), ty = L.f(src + kTY), sp = L.f(src + kSP);
    const double s2 = sqsum(tx, ty);                                     // :562; sqrt(s2) == 0 <=> s2 == 0
    const bool moving = s2 != 0.0;
    const double mag = fsqrt(pick(moving, s2, 1.0));
    const double nx = dmul(tx, kStepSize), ny = dmul(ty, kStepSize);
    const bool zx = nx == 0.0, zy = ny == 0.0;
    double qx, qy;
    fdiv2(pick(zx, mag, nx), pick(zy, mag, ny), mag, qx, qy);            // one refined reciprocal for both components
    const double x1 = dadd(x, dmul(sp, zx ? nx : qx));                   // :567
    const double y1 = dadd(y, dmul(sp, zy ? ny : qy));                   // :568
    xo = moving ? x1 : x;
    yo = moving ? y1 : y;
}

// The kinematics phase, :661-663: all five rows in one straight-line block (independent rows: the scheduler
// overlaps their sqrt / reciprocal chains).  4.6 KB of code; a rolled loop over one out-of-line copy fits the
// instruction cache better (hit rate 94.5 % -> 99.4 %) but is 3 % slower for the lost overlap (r1_history.md, r1i).
// PAIRS: two trips over a pair of player rows + the ball row -- 3/5 of the code, two chains still overlap.  The time-sliced
// kernel runs 1.7 % faster with it, the plain kernel 1.2 % slower (profiles/r2_v0_history.md, r2g).
template <bool PAIRS = false>
__device__ __forceinline__ void advance_all(Lane L)
{
    if (PAIRS) {
#pragma unroll 1
        for (int p = 0; p < 2; ++p) {
            const int r0 = 2 * p * kRowStride, r1 = r0 + kRowStride;
            double ax, ay, bx, by;
            advance_xy(L, r0, ax, ay);
            advance_xy(L, r1, bx, by);
            L.f(r0 + kX) = ax; L.f(r0 + kY) = ay; L.f(r1 + kX) = bx; L.f(r1 + kY) = by;
        }
        double cx, cy;
        advance_xy(L, kBallRow * kRowStride, cx, cy);
        L.f(kBallRow * kRowStride + kX) = cx; L.f(kBallRow * kRowStride + kY) = cy;
        return;
    }
    double nx[5], ny[5];
#pragma unroll
    for (int r = 0; r < 5; ++r) advance_xy(L, r * kRowStride, nx[r], ny[r]);
#pragma unroll
    for (int r = 0; r < 5; ++r) { L.f(r * kRowStride + kX) = nx[r]; L.f(r * kRowStride + kY) = ny[r]; }
}

// single row, out of line: the opponents' anticipation of the ball (:962-965)
struct XY { double x, y; };
static __device__ __noinline__ XY advance_row(Lane L, int src)
{
    XY o;
    advance_xy(L, src, o.x, o.y);
    return o;
}

__device__ __forceinline__ bool out_of_pitch(double x, double y)
{   // out, :574-577
    return (x < 0.0 || x > kFieldLen) || (y < 0.0 || y > kFieldWid);
}

struct StepResult { double reward; int done; int flags; };

// FutbolEnv.step, :628-717.  `ai_action` in 0..15.  RANDOM_OPP = the constructor's random_opp (:138), a
// compile-time variant so that each kernel carries only its own opponent code.
// `opp_action`: -1 = the reference's own opponents; 0..15 (RANDOM_OPP variant only) = actions supplied by the caller
// for opp_1 (a / 4) and opp_2 (a % 4), the self-play hook: same path as the random opponents (:642-645), the
// randint(0, 15) draw is not taken.
struct NoHook { __device__ __forceinline__ void operator()() const {} };
// LAYOUT (code layout of the rare / repeated parts; results are identical): bit 0 = the pending pass resolved through a call,
// bit 1 = the pending shot resolved through a call, bit 2 = the five kinematics rows through one out-of-line copy,
// bit 3 = the kinematics as two trips over a pair of player rows + the ball row
template <bool RANDOM_OPP, typename Hook = NoHook, int LAYOUT = 0>
__device__ __forceinline__ StepResult v0_step(Lane L, V0Regs &s, const V0Params &P, uint32_t env_id, int ai_action, int opp_action = -1,
                                              Hook before_draw_store = Hook())
{
    // All three Philox blocks (twelve words) up front.  Generating the third one -- draws 9 and 10: never consumed with the
    // hard-coded opponents in 5,400 reference steps, in 1.2 % of the steps with random ones -- only on demand saves 39
    // instructions per warp-step and was measured SLOWER (2^20 envs: 1.052e10 against 1.078e10 env-steps/s): the test before
    // the last turn and the second copy of the block cost the instruction-fetch-bound loop more (profiles/r2_v0_history.md).
    before_draw_store();
    philox_fill_step(&L.draw(0), P.key, env_id, kStreamDynamics, s.t_total);
    uint32_t j = 0;
    PendingShot shot;
    shot.shooter = -1; shot.target_y = 0; shot.pick_idx = 0; shot.passer = -1; shot.pass_w = 0;
    const int bo = kBallRow * kRowStride;

    // pre-step snapshot used by the reward (:630-635).  The owner one-hot row of the observation is all
    // zeros between reset() and the end of the first step, otherwise 10 * onehot(owner).  Everything
    // _get_reward reads from the snapshot is a comparison of pre-step values: evaluate them now.
    const bool fresh = s.ep_step == 0;
    const bool pre_ai1 = !fresh && s.owner == kAI1, pre_ai2 = !fresh && s.owner == kAI2;
    const bool pre_none = !fresh && s.owner == kNoOne;
    bool far1, far2, own_forward_pass;
    {
        const double b_x = L.f(bo + kX), b_y = L.f(bo + kY), b_tx = L.f(bo + kTX), b_ty = L.f(bo + kTY);
        const double a1x = L.f(kAI1 * kRowStride + kX), a1y = L.f(kAI1 * kRowStride + kY);
        const double a2x = L.f(kAI2 * kRowStride + kX), a2y = L.f(kAI2 * kRowStride + kY);
        far1 = sqsum(dsub(b_x, a1x), dsub(b_y, a1y)) > kSqLe2;           // distance > 2, :757, :790
        far2 = sqsum(dsub(b_x, a2x), dsub(b_y, a2y)) > kSqLe2;           // :758, :799
        own_forward_pass = b_tx > b_ty && b_tx > 0.0 && b_x > a1x && b_x > a2x && pre_none;   // :831 (Q6)
    }

    // ---- what the two opponents will do ----
    int opp_a1, opp_a2;                      // the action each opponent's turn is run with
    bool has1 = false, has2 = false, set1 = false, set2 = false, run1 = false, run2 = false;
    double t1x = -1.0, t1y = -1.0, t2x = -1.0, t2y = 1.0;
    if (RANDOM_OPP) {                                                    // :639-645
        const bool given = opp_action >= 0;
        const int r = given ? (opp_action & 15) : (int)__umulhi(L.draw(j), 16u);   // randint(0, 15)
        j += given ? 0u : 1u;
        opp_a1 = r >> 2; opp_a2 = r & 3;
    } else {                                                             // _opp_team_set_vector_observation, :864-947
        has1 = s.owner == kOpp1; has2 = s.owner == kOpp2;                // :866-877 (latched before either acts)
        const bool team_has = has1 || has2;
        int both = 0;                                                    // :879-880, one copy of the rule
#pragma unroll 1
        for (int o = 0; o < 2; ++o) both |= easy_action(L, j, kOpp1 + o, s.owner == kOpp1 + o, team_has) << (2 * o);
        const int a1 = both & 3, a2 = both >> 2;
        run1 = a1 == kRun; run2 = a2 == kRun;
        opp_a1 = a1; opp_a2 = a2;                                        // overrides do not touch a1 / a2
        const double o1x = L.f(kOpp1 * kRowStride + kX), o1y = L.f(kOpp1 * kRowStride + kY);
        const double o2x = L.f(kOpp2 * kRowStride + kX), o2y = L.f(kOpp2 * kRowStride + kY);
        const bool diag1 = o1y > kWid02, diag2 = o2y < kWid08;
        // carrier runs on the diagonal, its running mate mirrors it, :893-928
        set1 = diag1 && ((has1 && run1) || (has2 && run2 && run1 && o1x > kLen01));
        set2 = diag2 && ((has2 && run2) || (has1 && run1 && run2 && o2x > kLen01));
        if ((s.owner == kAI1 || s.owner == kAI2) && L.f(bo + kX) < kLen06) {   // :931-947: the deeper opp defends
            if (o1x > o2x) { opp_a1 = kRun; set1 = true; t1x = dsub(kDefendX, o1x); t1y = dsub(kDefendY, o1y); }
            else           { opp_a2 = kRun; set2 = true; t2x = dsub(kDefendX, o2x); t2y = dsub(kDefendY, o2y); }
        }
    }
    const int action1 = ai_action >> 2, action2 = ai_action & 3;         // :653

    // ---- the four turns, in the reference's order opp_1, opp_2, ai_1, ai_2 (:639-656): one copy of the code ----
#pragma unroll 1
    for (int t = 0; t < 4; ++t) {
        const int a = (t + 2) & 3;
        const bool opp = t < 2;
        const bool latched = !RANDOM_OPP && opp;
        const bool has_ball = latched ? (t == 0 ? has1 : has2) : s.owner == a;
        const int action = t == 0 ? opp_a1 : (t == 1 ? opp_a2 : (t == 2 ? action1 : action2));
        const bool set_target = latched && (t == 0 ? set1 : set2);

        player_turn(L, j, s, P, a, has_ball, action, set_target, t == 0 ? t1x : t2x, t == 0 ? t1y : t2y, shot);
        if (!RANDOM_OPP && t == 1) {
            // :962-982 anticipate the ball: whoever can reach its next position lands exactly on it
            const XY nb = advance_row(L, bo);
            const double nbx = nb.x, nby = nb.y;
            const double v1x = dsub(nbx, L.f(kOpp1 * kRowStride + kX)), v1y = dsub(nby, L.f(kOpp1 * kRowStride + kY));
            const double v2x = dsub(nbx, L.f(kOpp2 * kRowStride + kX)), v2y = dsub(nby, L.f(kOpp2 * kRowStride + kY));
            const double q1 = sqsum(v1x, v1y), q2 = sqsum(v2x, v2y);
            const bool on = s.owner == kNoOne && run1 && run2;
            const bool c1 = on && q1 <= P.reach_sq_max;                  // hyp(v1) < 0.1 * player_speed
            const bool c2 = on && !c1 && q2 <= P.reach_sq_max;
            const double qs = c1 ? q1 : q2;
            const bool live = (c1 || c2) && qs != 0.0;                   // others: benign operand, result unused
            const double ms = fsqrt(pick(live, qs, 1.0));
            const double sp = qs == 0.0 ? 0.0 : fdiv(pick(live, ms, 1.0), kStepSize);
            if (c1 || c2) {
                const int ro = (c1 ? kOpp1 : kOpp2) * kRowStride;
                L.f(ro + kTX) = c1 ? v1x : v2x; L.f(ro + kTY) = c1 ? v1y : v2y; L.f(ro + kSP) = sp;
            }
        }
    }
    if ((shot.shooter & shot.passer) >= 0) {                             // at most one of the two is pending (both -1: skip)
        if (shot.shooter >= 0) {
            if (LAYOUT & 2) resolve_shot_outlined(L, s, P, RANDOM_OPP, env_id, shot);
            else resolve_shot(L, s, P, RANDOM_OPP, env_id, shot);
        } else if (LAYOUT & 1) resolve_pass_outlined(L, shot);
        else resolve_pass(L, shot);
    }

    // ---- kinematics, :661-663 ----
    if (LAYOUT & 4) {
#pragma unroll 1
        for (int r = 0; r < 5; ++r) { const XY n = advance_row(L, r * kRowStride); L.f(r * kRowStride + kX) = n.x; L.f(r * kRowStride + kY) = n.y; }
    } else advance_all<(LAYOUT & 8) != 0>(L);

    // ---- _get_reward, :752-861 (evaluated before the goal re-kickoff) ----
    const double ball_x = L.f(bo + kX), ball_y = L.f(bo + kY);
    const bool in_mouth = ball_y > kGoalLower && ball_y < kGoalUpper;
    const bool goal_for = ball_x >= kFieldLen && in_mouth, goal_against = ball_x <= 0.0 && in_mouth;   // score(), :580-583
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
            const bool ai_out = out_of_pitch(L.f(kAI1 * kRowStride + kX), L.f(kAI1 * kRowStride + kY)) ||
                                out_of_pitch(L.f(kAI2 * kRowStride + kX), L.f(kAI2 * kRowStride + kY));
            const double out_r = ai_out ? -0.6 : 0.0;                    // :823-826
            const bool ai_owns = s.owner == kAI1 || s.owner == kAI2;
            double get_ball;
            if (ai_owns && !pre_ai1 && !pre_ai2)                          // :828-836 (Q6)
                get_ball = own_forward_pass ? kRewStolen : kRewGained;
            else if ((s.owner == kAI1 && pre_ai1) || (s.owner == kAI2 && pre_ai2))
                get_ball = kRewKept;                                     // :837-839
            else
                get_ball = 0.0;
            reward = dadd(dadd(dadd(dadd(dadd(dadd(get_ball, score), get_scored), out_r), dadd(bad1, bad2)), adv_r), running_r);  // :861
        }
    }

    StepResult res;
    res.done = 0;
    res.flags = 0;
    double fx = ball_x, fy = ball_y;                                     // the ball as out_of_field sees it
    if (goal_for || goal_against) {                                      // :670-699
        if (ball_x <= 0.0) s.opp_score += 1; else s.ai_score += 1;
        if (P.one_goal_end) res.done = 1;
        kickoff(L, s);
        fx = kFieldLen / 2; fy = kFieldWid / 2;
        res.flags |= kFlagGoal;
    }
    {   // out_of_field + fix, :621-625, :587-604, :701-707 (Q7, Q8)
        const bool x_out = fx < 0.0 || fx > kFieldLen, y_out = fy < 0.0 || fy > kFieldWid;
        const bool y_score = fy > kGoalLower - 2 && fy < kGoalUpper + 2;
        if ((x_out && !y_score) || y_out) {
            const int new_owner = (s.last_owner == kOpp1 || s.last_owner == kOpp2) ? kAI1 : kOpp1;
            const double cx = fx < 0.0 ? 0.0 : (fx > kFieldLen ? kFieldLen : fx);   // lock_in, :68-74
            const double cy = fy < 0.0 ? 0.0 : (fy > kFieldWid ? kFieldWid : fy);
            const int po = new_owner * kRowStride;                       // the new owner is put on the ball, :601-604
            L.f(bo + kX) = cx; L.f(bo + kY) = cy; L.f(bo + kTX) = 0.0; L.f(bo + kTY) = 0.0; L.f(bo + kSP) = 0.0;
            L.f(po + kX) = cx; L.f(po + kY) = cy; L.f(po + kTX) = 0.0; L.f(po + kTY) = 0.0; L.f(po + kSP) = 0.0;
            s.owner = new_owner;
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

// observation element 25 + idx of the (6,5) array the reference returns (:717): ball_owner_array_update,
// :720-736; all zeros right after reset (:223)
__device__ __forceinline__ double obs_owner_elem(const V0Regs &s, int idx)
{
    return (s.ep_step != 0 && s.owner == idx) ? 10.0 : 0.0;
}

}  // namespace futbol
