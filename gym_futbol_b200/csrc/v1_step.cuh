// Device-side v1 environment step (N-vs-N, rigid-body contacts): one thread owns one environment.
// Replaces gym_futbol/envs_v1/futbol_env.py Futbol.step (:427-483) / reset (:145-150) with team.py,
// player.py, ball.py, and the subset of Chipmunk2D 7 (pymunk) those files drive: cpSpaceStep with
// circle/circle and circle/segment contacts, arbiter pre-step, damped velocity integration with the
// reference's speed clamps, warm start, ten sequential-impulse iterations.
//
// Parity: the game logic is pinned to traces of the reference's own Python (executed over a pymunk
// stand-in, tests/golden/v1_golden.npz); the physics at the pymunk boundary is UNPINNED: it restates Chipmunk's published
// algorithm, and what Chipmunk leaves implementation-defined (body order, contact order, ...) is specified in DESIGN.md
// and implemented identically by the CPU checker, with which this code agrees bit for bit.
// Arithmetic: fp64, one IEEE operation per written operation, no FMA contraction.
//
// Why one thread per environment and not one warp (players on lanes): the per-environment parallelism is
// tiny (5 to 21 bodies) and the expensive parts are sequential by construction (the 2N action turns share the
// ball; sequential impulses are Gauss-Seidel), so lanes-as-bodies leaves most of a warp idle in exactly the
// phases that cost the most, while environments are independent and plentiful.  Layout: the 6(2N+1) body
// doubles in shared memory, one column per lane (run-time indexable, conflict-free); the arbiter cache
// (accumulated impulse + step stamp per shape pair, dense, touched only by pairs in contact) in HBM; the
// step's contact list in local memory.
#pragma once
#include <stdint.h>
#include "philox.cuh"
#include "ieee_fast.cuh"

namespace futbol {
namespace v1 {

constexpr int kMaxN = 10, kMaxBodies = 2 * kMaxN + 1, kNSeg = 12, kMaxContacts = 32;
#ifndef FUTBOL_V1_SOLVER_ITERS
#define FUTBOL_V1_SOLVER_ITERS 10          // pymunk's Space.iterations default; other values only in timing builds (breaks parity)
#endif
constexpr int kSolverIterations = FUTBOL_V1_SOLVER_ITERS;
constexpr double kWidth = 105.0, kHeight = 68.0, kGoalSize = 20.0, kDt = 0.1;      // futbol_env.py:19-26
constexpr double kBallMaxV = 25.0, kPlayerMaxV = 10.0;                              // :31-32
constexpr double kBallWeight = 10.0, kPlayerWeight = 20.0;                          // :34-35
constexpr double kPlayerForce = 40.0, kBallForce = 120.0;                           // :37-38
constexpr double kRPlayer = 1.5, kRBall = 1.0, kRSeg = 1.0, kElasticity = 0.2;      // player.py:7, ball.py:7, :187
constexpr double kSlop = (double)0.1f;                                              // Chipmunk collision_slop: cpSpace.c writes the float literal 0.1f
enum : int { kFlagGoal = 1, kFlagOut = 2, kFlagDone = 4, kFlagGoalLeft = 8 };
constexpr uint32_t kResetBlock = 0x4000u;   // Philox block of the side drawn by reset()
constexpr uint32_t kStamp0 = 8;             // first space-step stamp (cache entries start at 0 = "never touched")

// the 12 segments of _setup_walls (:184-224): six boundary segments, then six goal-box segments; (ax, ay, bx, by)
#ifndef FUTBOL_HOST_SHIM
__constant__
#endif
static const double kSegments[kNSeg][4] = {
    {0, 0, 0, kHeight / 2 - kGoalSize / 2},
    {0, kHeight / 2 + kGoalSize / 2, 0, kHeight},
    {0, kHeight, kWidth, kHeight},
    {kWidth, 0, kWidth, kHeight / 2 - kGoalSize / 2},
    {kWidth, kHeight / 2 + kGoalSize / 2, kWidth, kHeight},
    {0, 0, kWidth, 0},
    {-2, kHeight / 2 - kGoalSize / 2, -2, kHeight / 2 + kGoalSize / 2},
    {-2, kHeight / 2 - kGoalSize / 2, 0, kHeight / 2 - kGoalSize / 2},
    {-2, kHeight / 2 + kGoalSize / 2, 0, kHeight / 2 + kGoalSize / 2},
    {kWidth + 2, kHeight / 2 - kGoalSize / 2, kWidth + 2, kHeight / 2 + kGoalSize / 2},
    {kWidth, kHeight / 2 - kGoalSize / 2, kWidth + 2, kHeight / 2 - kGoalSize / 2},
    {kWidth, kHeight / 2 + kGoalSize / 2, kWidth + 2, kHeight / 2 + kGoalSize / 2},
};
__device__ __forceinline__ void segment(int s, double &ax, double &ay, double &bx, double &by)
{
    ax = kSegments[s][0]; ay = kSegments[s][1]; bx = kSegments[s][2]; by = kSegments[s][3];
}

struct V1Params {
    uint64_t seed;
    PhiloxKey key;
    uint32_t env_id_offset;
    int n_envs;
    int n_players;        // number_of_player, :65
    int ep_limit;         // first k with k additions of 0.1 > total_time (300 for 30), :478-481
    int auto_reset;
    double damping_dt;    // pow(0.95, 0.1): space.damping ** dt, :99
    double bias_coef;     // 1 - pow(pow(1.0f - 0.1f, 60), 0.1): Chipmunk collision_bias default (float literals, cpSpace.c)
    double clamp_sq_player, clamp_sq_ball;   // largest |v|^2 whose correctly rounded root is <= 10 / 25: `sqrt(s) > max` as `s > bound`
    double form_x[2 * kMaxN], form_y[2 * kMaxN];   // kick-off formation, team.py:52-112
};

// ---- per-environment working storage --------------------------------------------------------------------
// shared memory, one column per lane: element (6 * body + field) of lane l at st[(6 * body + field) * kCol + l],
// kCol = 32: a lane walking its own column is conflict-free (consecutive lanes, consecutive words).  (Round 1 padded the
// columns to 33 for an observation writer that read ONE lane's column element by element; the present writer reads four
// fields of eight environments per instruction, for which 32 and 33 conflict alike, and without the padding -- and with
// the kick-off formation read from the launch parameters instead of a shared-memory copy -- a 10v10 warp's state is
// 32,256 B: SEVEN resident warps per SM instead of six.)
// fields x, y, vx, vy, v_bias_x, v_bias_y; bodies: team A 0..N-1, team B N..2N-1, ball 2N.
#ifndef FUTBOL_HOST_SHIM
extern __shared__ __align__(16) unsigned char futbol_smem[];
#else
static unsigned char futbol_smem[8 * 6 * kMaxBodies + 16] __attribute__((aligned(16)));
#endif
constexpr int kCol = kLanes;
constexpr int kPX = 0, kPY = kCol, kVX = 2 * kCol, kVY = 3 * kCol, kBX = 4 * kCol, kBY = 5 * kCol;
constexpr int kBodyStride = 6 * kCol;

struct Lane {
    uint32_t st;          // index (in doubles) of this lane's element 0
    __device__ __forceinline__ double &f(int k) const { return reinterpret_cast<double *>(futbol_smem)[st + k]; }
};

__host__ __device__ constexpr int warp_state_bytes(int n_players) { return ((6 * (2 * n_players + 1) * kCol * 8) + 15) & ~15; }
__host__ __device__ constexpr int obs_dim(int n_players) { return 4 + 8 * n_players; }
__host__ __device__ constexpr int warp_smem_bytes(int n_players) { return warp_state_bytes(n_players); }
__host__ __device__ constexpr int block_smem_bytes(int n_players, int warps) { return warps * warp_smem_bytes(n_players); }
__host__ __device__ constexpr int n_pairs(int bodies) { return bodies * (bodies - 1) / 2 + bodies * kNSeg; }

__device__ __forceinline__ Lane make_lane(int warp_in_block, int lane, int n_players)
{
    Lane L;
    L.st = (uint32_t)(warp_in_block * (warp_smem_bytes(n_players) / 8) + lane);
    return L;
}

// the arbiter cache of this environment in HBM: pair q at rec[q * stride], one 16-byte record per shape pair (the
// accumulated normal impulse and the stamp of the space step in which the pair last touched), so that a touching pair
// costs one 32-byte sector each way (as two arrays it cost two: ncu counted 1.17 x the algorithmic DRAM bytes at 5v5)
struct __align__(16) CacheRec { double jn; uint32_t last, pad_; };
struct PairCache { CacheRec *rec; size_t stride; };
__device__ __forceinline__ CacheRec cache_load(const PairCache &C, int q)
{
#ifndef FUTBOL_HOST_SHIM
    // One 128-bit load through L1.  (The time-sliced rollout hands an env's records from SM to SM: its reader fences after
    // polling the progress word -- MEMBAR + CCTL.IVALL in SASS, which drops every L1 line of the SM -- so a plain load cannot
    // see a stale line.  ld.global.cg here instead costs 30-50 % of the whole rollout: profiles/r2_v1_history.md.)
    const double2 raw = *reinterpret_cast<const double2 *>(C.rec + (size_t)q * C.stride);
    CacheRec r;
    r.jn = raw.x; r.last = (uint32_t)__double2loint(raw.y); r.pad_ = 0;
    return r;
#else
    return C.rec[(size_t)q * C.stride];
#endif
}
__device__ __forceinline__ void cache_store(const PairCache &C, int q, double jn, uint32_t stamp)
{
#ifndef FUTBOL_HOST_SHIM
    *reinterpret_cast<double2 *>(C.rec + (size_t)q * C.stride) = make_double2(jn, __hiloint2double(0, (int)stamp));
#else
    CacheRec r; r.jn = jn; r.last = stamp; r.pad_ = 0;
    C.rec[(size_t)q * C.stride] = r;
#endif
}

// scalars (registers)
struct V1Regs {
    uint64_t t_total;
    uint32_t stamp;       // space steps taken so far (0.1 steps and the 1e-4 kick-off steps), starts at kStamp0
    int ep_step, owner_side;
};

// q: pair id, plus kWarmBit when the pair also touched in the previous space step (arbiter state NORMAL: only those are
// warm-started, cpArbiterApplyCachedImpulse returns early for FIRST_COLLISION)
struct Contact { double nx, ny, n_mass, bias, bounce, jn, jbias; int a, b, q; };
constexpr int kWarmBit = 0x8000, kPairMask = 0x7fff;

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double dsqrt(double a) { return __dsqrt_rn(a); }
// c ? a : b that the optimiser cannot look through: keeps a zero operand off the slow path of the IEEE sqrt
// sequence (see v0_step.cuh `pick`)
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

// sqrt for the hot path: the guard-free sequence of ieee_fast.cuh with the zero operand selected around it
__device__ __forceinline__ double sqrt0(double x)
{
    const bool nz = x != 0.0;
    const double r = fsqrt(pick(nz, x, 1.0));
    return nz ? r : 0.0;
}

// One Philox word for the rare draws (pass target, out-of-bounds receiver, side after a goal / reset): out of line
// and keyed by the seed (the round keys are re-derived) so that the hot loop carries a single inlined Philox.
static __device__ __noinline__ uint32_t philox_word_cold(uint64_t seed, uint32_t env_id, uint32_t stream, uint64_t t, uint32_t block, uint32_t w)
{
    const PhiloxKey K = philox_expand_key(seed);
    const Philox4 p = philox_step_block(K, env_id, stream, t, block);
    const uint32_t lo = (w & 1u) ? p.y : p.x, hi = (w & 1u) ? p.w : p.z;
    return (w & 2u) ? hi : lo;
}
// sequential dynamics draw (stream 3)
__device__ __forceinline__ uint32_t draw(const V1Params &P, uint32_t env_id, uint64_t t, uint32_t &j)
{
    const uint32_t w = philox_word_cold(P.seed, env_id, kStreamV1Dynamics, t, j >> 2, j & 3u);
    j += 1;
    return w;
}

// _position_to_initial, :129-143: teleport to the formation, zero velocities, space.step(1e-4).  With all
// velocities zero the 1e-4 step only consumes the bias velocities (p += v_bias * 1e-4, v_bias = 0), finds no
// contact (formation spacing >= 13.6) and ages the cached arbiters by one step.  Out of line (three call sites);
// the formation is read from the launch parameters (constant bank; the kernels declare them __grid_constant__).
static __device__ __noinline__ void position_to_initial(Lane L, const V1Params &P)
{
    const int N = P.n_players, B = 2 * N + 1;
#pragma unroll 1
    for (int i = 0; i < B; ++i) {
        const int o = i * kBodyStride;
        const double x = i < 2 * N ? P.form_x[i] : dmul(kWidth, 0.5), y = i < 2 * N ? P.form_y[i] : dmul(kHeight, 0.5);
        L.f(o + kPX) = dadd(x, dmul(dadd(0.0, L.f(o + kBX)), 0.0001));
        L.f(o + kPY) = dadd(y, dmul(dadd(0.0, L.f(o + kBY)), 0.0001));
        L.f(o + kVX) = 0.0; L.f(o + kVY) = 0.0; L.f(o + kBX) = 0.0; L.f(o + kBY) = 0.0;
    }
}

__device__ __forceinline__ void reset_env(Lane L, V1Regs &s, const V1Params &P, uint32_t env_id)
{   // Futbol.reset, :145-150 (t_total, the Philox step index, is deliberately kept; so is the arbiter cache)
    s.ep_step = 0;
    s.owner_side = (int)__umulhi(philox_word_cold(P.seed, env_id, kStreamV1Dynamics, s.t_total, kResetBlock, 0u), 2u);
    position_to_initial(L, P);
    s.stamp += 1;
}

// first construction (Futbol.__init__ -> reset): zero bias velocities, fresh stamps
__device__ __forceinline__ void init_env(Lane L, V1Regs &s, const V1Params &P, uint32_t env_id)
{
    const int B = 2 * P.n_players + 1;
    for (int i = 0; i < B; ++i) { L.f(i * kBodyStride + kBX) = 0.0; L.f(i * kBodyStride + kBY) = 0.0; }
    s.t_total = 0;
    s.stamp = kStamp0;
    reset_env(L, s, P, env_id);
}

// observation element k of the normalised vector [ball, team A, team B], :154-180
__device__ __forceinline__ double obs_elem(Lane L, int N, int k)
{
    const int body = k < 4 ? 2 * N : (k - 4) >> 2, fld = k & 3;
    const double v = L.f(body * kBodyStride + fld * kCol);
    const double avg = fld == 0 ? 52.5 : (fld == 1 ? 34.0 : 0.0);
    const double rng = fld == 0 ? (k < 4 ? 52.5 : 55.5) : (fld == 1 ? 34.0 : (k < 4 ? 25.0 : 10.0));
    const double num = dsub(v, avg);
    const bool z = num == 0.0;                                           // 0 / rng = that same zero: keep it off the divider's slow path
    const double q = fdiv(pick(z, 1.0, num), rng);
    return z ? num : q;
}

__device__ __forceinline__ bool touching(Lane L, int p, int ball)
{   // Ball.has_contact_with, ball.py:39-40 = Chipmunk CircleToCircle: |delta|^2 < (r1 + r2)^2
    const double dx = dsub(L.f(p * kBodyStride + kPX), L.f(ball * kBodyStride + kPX));
    const double dy = dsub(L.f(p * kBodyStride + kPY), L.f(ball * kBodyStride + kPY));
    return dadd(dmul(dx, dx), dmul(dy, dy)) < (kRBall + kRPlayer) * (kRBall + kRPlayer);
}

// Team.get_pass_target_teammate, team.py:136-180; returns the body index of the target
__device__ __forceinline__ int pass_target(Lane L, const V1Params &P, uint32_t env_id, uint64_t t, uint32_t &j, int p, int arrow)
{
    const int N = P.n_players, base = p < N ? 0 : N, k = p - base;
    if (N == 1) return p;                                                // :137-138
    const uint32_t w = draw(P, env_id, t, j);                            // :141-142: any other teammate
    const int r = (int)(((uint64_t)(w >> 8) * (uint64_t)(N - 1)) >> 24);
    int target = r < k ? r : r + 1;
    if (arrow != 0) {                                                    // :148-178
        const double px = L.f(p * kBodyStride + kPX), py = L.f(p * kBodyStride + kPY);
        uint32_t elig = 0;
        for (int i = 0; i < N; ++i) {
            const double mx = dsub(L.f((base + i) * kBodyStride + kPX), px), my = dsub(L.f((base + i) * kBodyStride + kPY), py);
            const bool ok = arrow == 1 ? my > 0.0 : (arrow == 2 ? mx > 0.0 : (arrow == 3 ? my < 0.0 : mx < 0.0));
            elig |= ok ? (1u << i) : 0u;
        }
        const int cnt = __popc(elig);
        if (cnt > 0) {
            const uint32_t w2 = draw(P, env_id, t, j);
            int pick = (int)(((uint64_t)(w2 >> 8) * (uint64_t)cnt) >> 24);
            for (int i = 0; i < N; ++i) if (elig & (1u << i)) { if (pick == 0) { target = i; break; } pick -= 1; }
        }
    }
    return base + target;
}

// _process_action, :309-422.  The common keys (noop, dash, press) are one straight-line block with selects; only
// a kick (shoot / pass while touching the ball: a few percent of the turns) stays behind a branch.
__device__ __forceinline__ void process_action(Lane L, V1Regs &s, const V1Params &P, uint32_t env_id, uint32_t &j, int p, int arrow, int key)
{
    const int N = P.n_players, ball = 2 * N, side = p < N ? 0 : 1;
    const int po = p * kBodyStride, bo = ball * kBodyStride;
    const double m_inv_p = 1.0 / kPlayerWeight, m_inv_b = 1.0 / kBallWeight;
    const double fx = arrow == 2 ? 1.0 : (arrow == 4 ? -1.0 : 0.0), fy = arrow == 1 ? 1.0 : (arrow == 3 ? -1.0 : 0.0);   // :312-327
    const double px = L.f(po + kPX), py = L.f(po + kPY), pvx = L.f(po + kVX), pvy = L.f(po + kVY);
    const double dx = dsub(L.f(bo + kPX), px), dy = dsub(L.f(bo + kPY), py);          // player -> ball
    const double d2 = dadd(dmul(dx, dx), dmul(dy, dy));
    // Ball.has_contact_with, ball.py:39-40 = Chipmunk CircleToCircle: |delta|^2 < (r1 + r2)^2 (delta = p - ball there:
    // the same squares)
    const bool touch = d2 < (kRBall + kRPlayer) * (kRBall + kRPlayer);
    const bool is_move = key <= 1;                                       // noop :331-335, dash :338-341
    const bool is_press = key == 3 && !touch && arrow == 0;              // press :371-391 (not touching: d2 >= 6.25 > 0)
    const double f = key == 0 ? kPlayerWeight : kPlayerForce;
    const double mag = fsqrt(pick(is_press, d2, 1.0));
    double pfx, pfy;
    fdiv2(dmul(kPlayerForce, pick(is_press, dx, 1.0)), dmul(kPlayerForce, pick(is_press, dy, 1.0)), mag, pfx, pfy);
    // apply_impulse_at_local_point: v += j * m_inv
    const double ivx = is_move ? dmul(dmul(f, fx), m_inv_p) : dmul(pfx, m_inv_p);
    const double ivy = is_move ? dmul(dmul(f, fy), m_inv_p) : dmul(pfy, m_inv_p);
    const double nvx = dadd(pvx, ivx), nvy = dadd(pvy, ivy);
    if (is_move || is_press) { L.f(po + kVX) = nvx; L.f(po + kVY) = nvy; }
    if (is_move && touch) { L.f(bo + kVX) = nvx; L.f(bo + kVY) = nvy; }  // _ball_move_with_player, :300-304
    if ((key == 2 || key == 4) && touch) {                               // shoot :344-366, pass :394-416
        double gx, gy, force, div;
        if (key == 2) { gx = side == 0 ? kWidth : 0.0; gy = kHeight / 2; force = kBallForce; div = 2.0; }
        else {
            const int tg = pass_target(L, P, env_id, s.t_total, j, p, arrow);
            gx = L.f(tg * kBodyStride + kPX); gy = L.f(tg * kBodyStride + kPY); force = kBallForce - 20; div = 10.0;
        }
        const double vx = dsub(gx, L.f(bo + kPX)), vy = dsub(gy, L.f(bo + kPY));
        const double k2 = dadd(dmul(vx, vx), dmul(vy, vy));
        if (k2 != 0.0) {
            // guard-free sequences (ieee_fast.cuh): |v| is a pitch-scale distance; a zero numerator keeps its sign
            // through the select; the ball's velocity is 0 or far above the sequences' lower bound
            const double kmag = fsqrt(k2);
            const double nx_ = dmul(force, vx), ny_ = dmul(force, vy), ovx = L.f(bo + kVX), ovy = L.f(bo + kVY);
            const bool zx = nx_ == 0.0, zy = ny_ == 0.0, zvx = ovx == 0.0, zvy = ovy == 0.0;
            double qx, qy, hx, hy;
            fdiv2(pick(zx, 1.0, nx_), pick(zy, 1.0, ny_), kmag, qx, qy);
            fdiv2(pick(zvx, 1.0, ovx), pick(zvy, 1.0, ovy), div, hx, hy);
            L.f(bo + kVX) = dadd(zvx ? ovx : hx, dmul(zx ? nx_ : qx, m_inv_b));
            L.f(bo + kVY) = dadd(zvy ? ovy : hy, dmul(zy ? ny_ : qy, m_inv_b));
        } else {                                                         // kicked at a point exactly under the ball: 0 / 0 as in the reference
            const double kmag = dsqrt(k2);
            const double bfx = ddiv(dmul(force, vx), kmag), bfy = ddiv(dmul(force, vy), kmag);
            L.f(bo + kVX) = dadd(ddiv(L.f(bo + kVX), div), dmul(bfx, m_inv_b));
            L.f(bo + kVY) = dadd(ddiv(L.f(bo + kVY), div), dmul(bfy, m_inv_b));
        }
    }
    if (touch) s.owner_side = side;                                      // :364, :414, :450-451
}

// closest point of segment s to (cx, cy): every segment is axis-aligned, so it is the centre's coordinate
// clamped to the segment's extent (Chipmunk CircleToSegment: a + (b - a) clamp01(((b - a).(c - a)) / |b - a|^2))
__device__ __forceinline__ void seg_closest(int sg, double cx, double cy, double &qx, double &qy)
{
    double ax, ay, bx, by;
    segment(sg, ax, ay, bx, by);
    if (ax == bx) { qx = ax; qy = cy < ay ? ay : (cy > by ? by : cy); }
    else { qy = ay; qx = cx < ax ? ax : (cx > bx ? bx : cx); }
}

__device__ __forceinline__ bool ball_touches_segment(Lane L, int ball, int sg)
{
    const double cx = L.f(ball * kBodyStride + kPX), cy = L.f(ball * kBodyStride + kPY);
    double qx, qy;
    seg_closest(sg, cx, cy, qx, qy);
    const double dx = dsub(qx, cx), dy = dsub(qy, cy);
    return dadd(dmul(dx, dx), dmul(dy, dy)) < (kRBall + kRSeg) * (kRBall + kRSeg);
}

// one solver iteration for one contact (cpArbiterApplyImpulse)
__device__ __forceinline__ void solve_contact(Lane L, Contact &k, int ball)
{
    const double m_inv_p = 1.0 / kPlayerWeight, m_inv_b = 1.0 / kBallWeight;
    const int a = k.a, b = k.b, ao = a * kBodyStride, bo = (b >= 0 ? b : 0) * kBodyStride;
    const double ma = a == ball ? m_inv_b : m_inv_p, mb = b >= 0 ? (b == ball ? m_inv_b : m_inv_p) : 0.0;
    const double vb2x = b >= 0 ? L.f(bo + kBX) : 0.0, vb2y = b >= 0 ? L.f(bo + kBY) : 0.0;
    const double v2x = b >= 0 ? L.f(bo + kVX) : 0.0, v2y = b >= 0 ? L.f(bo + kVY) : 0.0;
    const double vbn = dadd(dmul(dsub(vb2x, L.f(ao + kBX)), k.nx), dmul(dsub(vb2y, L.f(ao + kBY)), k.ny));
    const double vrn = dadd(dmul(dsub(v2x, L.f(ao + kVX)), k.nx), dmul(dsub(v2y, L.f(ao + kVY)), k.ny));
    const double jbn = dmul(dsub(k.bias, vbn), k.n_mass), jbn_old = k.jbias;
    const double t1 = dadd(jbn_old, jbn);
    k.jbias = pick(t1 > 0.0, t1, 0.0);      // cpfmax(x, 0) as compare + select: the ternary compiles to the NaN-quieting DSETP.MAX sequence
    const double jn = dmul(-dadd(k.bounce, vrn), k.n_mass), jn_old = k.jn;
    const double t2 = dadd(jn_old, jn);
    k.jn = pick(t2 > 0.0, t2, 0.0);
    const double db = dsub(k.jbias, jbn_old), dj = dsub(k.jn, jn_old);
    const double bx = dmul(k.nx, db), by = dmul(k.ny, db), jx = dmul(k.nx, dj), jy = dmul(k.ny, dj);
    L.f(ao + kBX) = dsub(L.f(ao + kBX), dmul(bx, ma)); L.f(ao + kBY) = dsub(L.f(ao + kBY), dmul(by, ma));
    L.f(ao + kVX) = dsub(L.f(ao + kVX), dmul(jx, ma)); L.f(ao + kVY) = dsub(L.f(ao + kVY), dmul(jy, ma));
    if (b >= 0) {
        L.f(bo + kBX) = dadd(L.f(bo + kBX), dmul(bx, mb)); L.f(bo + kBY) = dadd(L.f(bo + kBY), dmul(by, mb));
        L.f(bo + kVX) = dadd(L.f(bo + kVX), dmul(jx, mb)); L.f(bo + kVY) = dadd(L.f(bo + kVY), dmul(jy, mb));
    }
}

// warm start of one contact (cpArbiterApplyCachedImpulse, dt_coef = 1)
__device__ __forceinline__ void warm_start_contact(Lane L, const Contact &k, int ball)
{
    if (!(k.q & kWarmBit)) return;
    const double m_inv_p = 1.0 / kPlayerWeight, m_inv_b = 1.0 / kBallWeight;
    const double jx = dmul(k.nx, k.jn), jy = dmul(k.ny, k.jn), ma = k.a == ball ? m_inv_b : m_inv_p;
    const int ao = k.a * kBodyStride;
    L.f(ao + kVX) = dsub(L.f(ao + kVX), dmul(jx, ma)); L.f(ao + kVY) = dsub(L.f(ao + kVY), dmul(jy, ma));
    if (k.b >= 0) {
        const double mb = k.b == ball ? m_inv_b : m_inv_p;
        const int bo = k.b * kBodyStride;
        L.f(bo + kVX) = dadd(L.f(bo + kVX), dmul(jx, mb)); L.f(bo + kVY) = dadd(L.f(bo + kVY), dmul(jy, mb));
    }
}

// cpSpaceStep(dt = 0.1).  Returns the number of contacts; `overflow` counts contacts beyond kMaxContacts.
// The first REGC contacts of a step live in registers (c0, c1, c2), the rest in the local-memory list `con` (index
// i - REGC): a step rarely has more than two, and the ten solver iterations would otherwise wait on local-memory
// loads (L1 is small next to 200 KB of shared memory: ncu showed 53 % of the long-scoreboard stalls there).
// REGC = how many contacts are register-resident, chosen by measurement (end of round 2, env-steps/s, tools/exp_variants_v1.sh):
// 0 for 1v1 and 2v2 (with a register cap, v1_kernels.cu), 1 for 3v3 and 4v4 (4v4: +4 % over 2), 2 for 5v5 and 6v6
// (5v5: 0 -8 %, 1 +0.3 %, 3 -15 %), 3 for N >= 7.
template <int REGC>
__device__ __forceinline__ int space_step(Lane L, V1Regs &s, const V1Params &P, const PairCache &C, Contact *con, int &overflow)
{
    const int N = P.n_players, B = 2 * N + 1, ball = 2 * N, CC = B * (B - 1) / 2;
    const double m_inv_p = 1.0 / kPlayerWeight, m_inv_b = 1.0 / kBallWeight;
    int nc = 0;
    Contact c0, c1, c2;
    c0.a = c0.b = c0.q = 0; c1.a = c1.b = c1.q = 0;
    c0.nx = c0.ny = c0.n_mass = c0.bias = c0.bounce = c0.jn = c0.jbias = 0.0; c1 = c0; c2 = c0;
    // 1. integrate positions (cpBodyUpdatePosition): p += (v + v_bias) dt; v_bias = 0
#pragma unroll 1
    for (int i = 0; i < B; ++i) {
        const int o = i * kBodyStride;
        L.f(o + kPX) = dadd(L.f(o + kPX), dmul(dadd(L.f(o + kVX), L.f(o + kBX)), kDt));
        L.f(o + kPY) = dadd(L.f(o + kPY), dmul(dadd(L.f(o + kVY), L.f(o + kBY)), kDt));
        L.f(o + kBX) = 0.0; L.f(o + kBY) = 0.0;
    }
    // 2. narrow phase in pair-id order.
    //    pass 0: circle/circle pairs (i < j), id j(j-1)/2 + i: outer loop over j, inner over i;
    //    pass 1: circle/segment pairs, id CC + 12 * body + segment.
    //    The scan is the SAME instruction stream for every lane: for body j a branch-free loop over all i < j builds the
    //    mask of touching partners (two shared-memory loads and six fp64 operations per pair, independent across
    //    pairs, so they pipeline), and only a non-empty mask -- a few percent of the bodies -- enters the branch that
    //    records the hits.  (The first version walked a per-lane candidate list through a resumable state machine: one
    //    pair per loop trip, 14 to 24 of 32 lanes active, about 90 warp-instructions per pair -- a third of all
    //    instructions of a 5v5 step and 44 % at 10v10, profiles/r2_v1_history.md.)
    //    Segments: only those that can be reached at all are tested -- r + r_segment <= 2.5, the left-hand segments
    //    {0, 1, 6, 7, 8} lie at x <= 0, the right-hand ones {3, 4, 9, 10, 11} at x >= 105, segment 5 on y = 0 and
    //    segment 2 on y = 68 -- so a body strictly inside the pitch tests none.
    //    A hit is RECORDED as an 11-bit code (in the q field of its future contact); the contacts are built afterwards,
    //    all lanes together, so that the 90-instruction block runs as many times as the busiest environment has
    //    contacts, not once per pair that touches in any of the warp's 32 environments.
    {
        auto record = [&](uint32_t code) {
            if (nc == kMaxContacts) { overflow += 1; return; }
            if (REGC > 0 && nc == 0) c0.q = (int)code; else if (REGC > 1 && nc == 1) c1.q = (int)code; else if (REGC > 2 && nc == 2) c2.q = (int)code;
            else con[nc - REGC].q = (int)code;
            nc += 1;
        };
        constexpr double kPP2 = (kRPlayer + kRPlayer) * (kRPlayer + kRPlayer), kPB2 = (kRPlayer + kRBall) * (kRPlayer + kRBall);
#pragma unroll 1
        for (int jb = 1; jb < B; ++jb) {
            const double jx = L.f(jb * kBodyStride + kPX), jy = L.f(jb * kBodyStride + kPY);
            const double mind2 = jb == ball ? kPB2 : kPP2;               // a = ii < jb: only b can be the ball
            uint32_t hj = 0u;
#pragma unroll 4
            for (int ii = 0; ii < jb; ++ii) {
                const double dx = dsub(jx, L.f(ii * kBodyStride + kPX)), dy = dsub(jy, L.f(ii * kBodyStride + kPY));
                const double distsq = dadd(dmul(dx, dx), dmul(dy, dy));
                hj |= distsq < mind2 ? (1u << ii) : 0u;
            }
            while (hj != 0u) {
                const int ii = __ffs((int)hj) - 1;
                hj &= hj - 1u;
                record(((uint32_t)jb << 1) | ((uint32_t)ii << 6));
            }
        }
#pragma unroll 1
        for (int jb = 0; jb < B; ++jb) {
            const double jx = L.f(jb * kBodyStride + kPX), jy = L.f(jb * kBodyStride + kPY);
            // segments 0 / 3 (x = 0 / 105, y in [0, 24]) need y < 26.5; 1 / 4 (y in [44, 68]) need y > 41.5; the goal boxes
            // {6, 7, 8} / {9, 10, 11} (y in [24, 44], x beyond the line) need 21.5 < y < 46.5
            const uint32_t ylo = jy < 26.5 ? 0x009u : 0u, yhi = jy > 41.5 ? 0x012u : 0u, ymid = (jy > 21.5 && jy < 46.5) ? 0xFC0u : 0u;
            uint32_t cand = ((jx < 2.5 ? 0x1C3u : 0u) | (jx > kWidth - 2.5 ? 0xE18u : 0u)) & (ylo | yhi | ymid);
            cand |= (jy < 2.5 ? 0x020u : 0u) | (jy > kHeight - 2.5 ? 0x004u : 0u);
            const double mind = (jb == ball ? kRBall : kRPlayer) + kRSeg;
            while (cand != 0u) {
                const int ii = __ffs((int)cand) - 1;
                cand &= cand - 1u;
                double tx, ty;
                seg_closest(ii, jx, jy, tx, ty);
                const double dx = dsub(tx, jx), dy = dsub(ty, jy);
                if (dadd(dmul(dx, dx), dmul(dy, dy)) < mind * mind) record(1u | ((uint32_t)jb << 1) | ((uint32_t)ii << 6));
            }
        }
    }
    // 5. arbiter pre-step for the recorded pairs (uses the velocities before step 6)
    {
        // one copy of the block for every contact: the code of contact h comes from, and the finished contact goes to,
        // c0 / c1 / c2 (h < REGC) or the local-memory list
#pragma unroll 1
        for (int h = 0; h < nc; ++h) {
            Contact k;
            const uint32_t code = (uint32_t)((REGC > 0 && h == 0) ? c0.q : ((REGC > 1 && h == 1) ? c1.q : ((REGC > 2 && h == 2) ? c2.q : con[h - REGC].q)));
            const int hp = (int)(code & 1u), hj = (int)((code >> 1) & 31u), ii = (int)(code >> 6);
            int a, b, q;                                                 // b < 0: static segment -1 - b
            if (hp == 0) { a = ii; b = hj; q = hj * (hj - 1) / 2 + ii; }
            else { a = hj; b = -1 - ii; q = CC + hj * kNSeg + ii; }
            // cached arbiter of the pair, requested first: the ~90 instructions below run under the HBM latency
            const CacheRec cached = cache_load(C, q);
            const int ao = a * kBodyStride;
            const double pax = L.f(ao + kPX), pay = L.f(ao + kPY);
            const double ra = a == ball ? kRBall : kRPlayer;
            double rb, tx, ty;                                           // (tx, ty): centre of b or closest point
            if (b >= 0) { rb = b == ball ? kRBall : kRPlayer; tx = L.f(b * kBodyStride + kPX); ty = L.f(b * kBodyStride + kPY); }
            else { rb = kRSeg; seg_closest(ii, pax, pay, tx, ty); }
            const double dx = dsub(tx, pax), dy = dsub(ty, pay);
            const double distsq = dadd(dmul(dx, dx), dmul(dy, dy));
            const double dist = sqrt0(distsq);
            if (dist != 0.0) { const double inv = fdiv(1.0, dist); k.nx = dmul(dx, inv); k.ny = dmul(dy, inv); }
            else if (b >= 0) { k.nx = 1.0; k.ny = 0.0; }
            else {   // segment normal: perp(normalize(b - a))
                double sax, say, sbx, sby;
                segment(ii, sax, say, sbx, sby);
                const double sx = dsub(sbx, sax), sy = dsub(sby, say), sl = dsqrt(dadd(dmul(sx, sx), dmul(sy, sy)));
                k.nx = -ddiv(sy, sl); k.ny = ddiv(sx, sl);
            }
            k.a = a; k.b = b;
            const double p1x = dadd(pax, dmul(k.nx, ra)), p1y = dadd(pay, dmul(k.ny, ra));
            const double p2x = dadd(tx, dmul(k.nx, -rb)), p2y = dadd(ty, dmul(k.ny, -rb));
            const double pen = dadd(dmul(dsub(p2x, p1x), k.nx), dmul(dsub(p2y, p1y), k.ny));
            // 1 / (m_inv_a + m_inv_b): four possible pairs of masses, each quotient folded by the compiler (IEEE)
            const double nm_pp = 1.0 / (m_inv_p + m_inv_p), nm_pb = 1.0 / (m_inv_p + m_inv_b), nm_ps = 1.0 / (m_inv_p + 0.0),
                nm_bs = 1.0 / (m_inv_b + 0.0);
            k.n_mass = b >= 0 ? ((a == ball || b == ball) ? nm_pb : nm_pp) : (a == ball ? nm_bs : nm_ps);
            double m = dadd(pen, kSlop);
            m = pick(m < 0.0, m, 0.0);                                   // cpfmin(0, dist + slop)
            const double bnum = dmul(-P.bias_coef, m);                   // -0.0 unless the pair overlaps by more than the slop
            const bool bz = bnum == 0.0;
            const double bq = fdiv(pick(bz, 1.0, bnum), kDt);
            k.bias = bz ? bnum : bq;                                     // 0 / dt = that same zero, kept off the divider's slow path
            k.jbias = 0.0;
            const double vbx = b >= 0 ? L.f(b * kBodyStride + kVX) : 0.0, vby = b >= 0 ? L.f(b * kBodyStride + kVY) : 0.0;
            const double el = b >= 0 ? kElasticity * kElasticity : kElasticity * 0.0;
            k.bounce = dmul(dadd(dmul(dsub(vbx, L.f(ao + kVX)), k.nx), dmul(dsub(vby, L.f(ao + kVY)), k.ny)), el);
            // cached arbiter (collision_persistence = 3): a pair that touched within 3 steps inherits its impulse;
            // it is warm-started with it only if it touched in the previous step too
            k.jn = (s.stamp - cached.last <= 3u) ? cached.jn : 0.0;
            k.q = q | (s.stamp - cached.last == 1u ? kWarmBit : 0);
            if (REGC > 0 && h == 0) c0 = k; else if (REGC > 1 && h == 1) c1 = k; else if (REGC > 2 && h == 2) c2 = k;
            else con[h - REGC] = k;
        }
    }
    // 6. integrate velocities through velocity_func (player.py:45-50, ball.py:49-54)
#pragma unroll 1
    for (int i = 0; i < B; ++i) {
        const int o = i * kBodyStride;
        double vx = dadd(dmul(L.f(o + kVX), P.damping_dt), 0.0), vy = dadd(dmul(L.f(o + kVY), P.damping_dt), 0.0);
        const double l2 = dadd(dmul(vx, vx), dmul(vy, vy));
        // `sqrt(l2) > max` decided on l2 (IEEE sqrt is monotone: the bound is exact); the root is taken only to clamp
        if (l2 > (i == ball ? P.clamp_sq_ball : P.clamp_sq_player)) {
            const double mx = i == ball ? kBallMaxV : kPlayerMaxV, sc = fdiv(mx, fsqrt(l2));
            vx = dmul(vx, sc); vy = dmul(vy, sc);
        }
        L.f(o + kVX) = vx; L.f(o + kVY) = vy;
    }
    // 7. warm start (cpArbiterApplyCachedImpulse, dt_coef = 1)
    if (REGC > 0 && nc > 0) warm_start_contact(L, c0, ball);
    if (REGC > 1 && nc > 1) warm_start_contact(L, c1, ball);
    if (REGC > 2 && nc > 2) warm_start_contact(L, c2, ball);
#pragma unroll 1
    for (int i = REGC; i < nc; ++i) warm_start_contact(L, con[i - REGC], ball);
    // 8. ten iterations of cpArbiterApplyImpulse over the contacts in order
    if (nc > 0) {
        // The contacts beyond the register-resident ones live in local memory, which misses L1 two times out of three (ncu, 5v5:
        // 188 local loads per warp-step at a 35 % hit rate).  For the big teams (REGC = 3: N >= 7, where a step often has more
        // than three contacts) each of them is requested one contact AHEAD -- the next contact's nine words travel while the
        // current contact's dependent chain runs -- and only its two accumulators are written back: 10v10 +4 %.  For the
        // smaller teams the 19 extra registers cost more than the overlap gains (3v3 -10 %, 4v4 -5 %, 5v5 +1 %): plain loop.
        if (REGC >= 3) {
            Contact nxt = c0;
            if (nc > REGC) nxt = con[0];
#pragma unroll 1
            for (int it = 0; it < kSolverIterations; ++it) {
                solve_contact(L, c0, ball);
                if (nc > 1) solve_contact(L, c1, ball);
                if (nc > 2) solve_contact(L, c2, ball);
#pragma unroll 1
                for (int i = REGC; i < nc; ++i) {
                    Contact cur = nxt;
                    const int ni = i + 1 < nc ? i + 1 - REGC : 0;        // the next contact: of this iteration, or the first of the next
                    nxt = con[ni];
                    solve_contact(L, cur, ball);
                    con[i - REGC].jn = cur.jn; con[i - REGC].jbias = cur.jbias;
                    if (ni == i - REGC) { nxt.jn = cur.jn; nxt.jbias = cur.jbias; }   // a single overflow contact: the prefetched copy is stale
                }
            }
        } else {
#pragma unroll 1
            for (int it = 0; it < kSolverIterations; ++it) {
                if (REGC > 0) solve_contact(L, c0, ball);
                if (REGC > 1 && nc > 1) solve_contact(L, c1, ball);
#pragma unroll 1
                for (int i = REGC; i < nc; ++i) solve_contact(L, con[i - REGC], ball);
            }
        }
        if (REGC > 0) cache_store(C, c0.q & kPairMask, c0.jn, s.stamp);
        if (REGC > 1 && nc > 1) cache_store(C, c1.q & kPairMask, c1.jn, s.stamp);
        if (REGC > 2 && nc > 2) cache_store(C, c2.q & kPairMask, c2.jn, s.stamp);
#pragma unroll 1
        for (int i = REGC; i < nc; ++i) cache_store(C, con[i - REGC].q & kPairMask, con[i - REGC].jn, s.stamp);
    }
    s.stamp += 1;
    return nc;
}

struct StepResult { double reward; int done; int flags; int contacts; int overflow; };

// Futbol.step, :427-483.  `left`: this env's 2N action bytes (arrow, key per left player) or nullptr =
// synthetic uniform actions from Philox stream 1.  `right`: the right team's 2N action bytes supplied by the caller (the
// self-play hook) or nullptr = action_space.sample() (:429, Philox stream 2).
template <int REGC>
__device__ __forceinline__ StepResult v1_step(Lane L, V1Regs &s, const V1Params &P, uint32_t env_id, const uint8_t *left,
                                              const PairCache &C, Contact *con, const uint8_t *right = nullptr)
{
    const int N = P.n_players, ball = 2 * N, bo = ball * kBodyStride;
    StepResult res;
    res.flags = 0; res.overflow = 0;
    uint32_t j = 0;                                                      // sequential dynamics draws of this step
    double init_d[kMaxN];                                                // :433
    const double bix = L.f(bo + kPX), biy = L.f(bo + kPY);               // :435
#pragma unroll 1
    for (int i = (N == 5 ? 3 : 0); i < N; ++i) {                         // only the players get_team_reward looks at (:501-504)
        const double dx = dsub(L.f(i * kBodyStride + kPX), bix), dy = dsub(L.f(i * kBodyStride + kPY), biy);
        init_d[i] = sqrt0(dadd(dmul(dx, dx), dmul(dy, dy)));
    }
    double reward = 0.0;

    // the left team's 2N action bytes, fetched up front (independent loads, one memory latency) and packed 6 bits per
    // player; inside the sequential turn loop below they would cost one exposed HBM latency per player
    uint64_t packed = 0;
    if (left != nullptr) {
        const uint16_t *pair = reinterpret_cast<const uint16_t *>(left);   // (arrow, key) of player p; 2N bytes per env: 2-byte aligned
#pragma unroll
        for (int p = 0; p < kMaxN; ++p)
            if (p < N) { const uint32_t w = pair[p]; packed |= (uint64_t)(((w & 0xffu) % 5u) | ((((w >> 8) & 0xffu) % 5u) << 3)) << (6 * p); }
    }
    Philox4 blk;                                                         // one Philox block = the actions of two players
    blk.x = blk.y = blk.z = blk.w = 0u;
#pragma unroll 1
    for (int p = 0; p < 2 * N; ++p) {                                    // :447-453, right team = action_space.sample() (:429)
        int arrow, key;
        if (p < N && left != nullptr) { arrow = (int)((packed >> (6 * p)) & 7u); key = (int)((packed >> (6 * p + 3)) & 7u); }
        else if (p >= N && right != nullptr) { arrow = right[2 * (p - N)] % 5; key = right[2 * (p - N) + 1] % 5; }
        else {
            const uint32_t stream = p < N ? kStreamActions : kStreamV1Opp;
            const int q = p < N ? p : p - N;
            if ((q & 1) == 0) blk = philox_step_block(P.key, env_id, stream, s.t_total, (uint32_t)q >> 1);
            const uint32_t w0 = (q & 1) ? blk.z : blk.x, w1 = (q & 1) ? blk.w : blk.y;
            arrow = (int)__umulhi(w0, 5u); key = (int)__umulhi(w1, 5u);
        }
        process_action(L, s, P, env_id, j, p, arrow, key);
    }

    bool out = false;                                                    // check_and_fix_out_bounds, :256-287
    // only boundary segments the ball can reach (r + r_segment = 2) are tested, in the reference's order (first hit wins)
    uint32_t ocand;
    {
        const double bx = L.f(bo + kPX), by = L.f(bo + kPY);
        ocand = ((bx < 2.0 ? 0x03u : 0u) | (bx > kWidth - 2.0 ? 0x18u : 0u)) & ((by < 26.0 ? 0x09u : 0u) | (by > 42.0 ? 0x12u : 0u));
        ocand |= (by > kHeight - 2.0 ? 0x04u : 0u) | (by < 2.0 ? 0x20u : 0u);
    }
#pragma unroll 1
    while (ocand != 0u && !out) {
        const int sg = __ffs((int)ocand) - 1;
        ocand &= ocand - 1u;
        if (!ball_touches_segment(L, ball, sg)) continue;
        out = true;
        const double bx = L.f(bo + kPX), by = L.f(bo + kPY);
        double dbx = 0, dby = 0, dpx = 0, dpy = 0;
        if (sg == 0 || sg == 1) { dbx = 3.5; dpx = 1; } else if (sg == 3 || sg == 4) { dbx = -3.5; dpx = -1; }
        else if (sg == 2) { dby = -3.5; dpy = -1; } else { dby = 3.5; dpy = 1; }
        L.f(bo + kPX) = dadd(bx, dbx); L.f(bo + kPY) = dadd(by, dby); L.f(bo + kVX) = 0.0; L.f(bo + kVY) = 0.0;
        const int pick = (int)__umulhi(draw(P, env_id, s.t_total, j), (uint32_t)N);
        const int g = s.owner_side == 1 ? pick : N + pick;               // the other side gets the ball
        s.owner_side = s.owner_side == 1 ? 0 : 1;
        const int go = g * kBodyStride;
        L.f(go + kPX) = dadd(bx, dpx); L.f(go + kPY) = dadd(by, dpy); L.f(go + kVX) = 0.0; L.f(go + kVY) = 0.0;
    }
    if (out) res.flags |= kFlagOut;

    res.contacts = space_step<REGC>(L, s, P, C, con, res.overflow);            // :459

    if (!out) {                                                          // :463-467
        double best = 0.0;
        bool first = true;
        const double bx = L.f(bo + kPX), by = L.f(bo + kPY);
        for (int i = (N == 5 ? 3 : 0); i < N; ++i) {                     // :501-504
            const double dx = dsub(L.f(i * kBodyStride + kPX), bx), dy = dsub(L.f(i * kBodyStride + kPY), by);
            const double diff = dsub(init_d[i], sqrt0(dadd(dmul(dx, dx), dmul(dy, dy))));
            if (first || diff > best) { best = diff; first = false; }
        }
        reward = dadd(reward, dmul(best, 10.0));
        const double ax = dsub(bx, kWidth), ay = dsub(by, kHeight / 2), ix = dsub(bix, kWidth), iy = dsub(biy, kHeight / 2);
        reward = dadd(reward, dmul(dsub(sqrt0(dadd(dmul(ix, ix), dmul(iy, iy))), sqrt0(dadd(dmul(ax, ax), dmul(ay, ay)))), 10.0));
    }

    bool goal = false;                                                   // ball_contact_goal, :291-296
    {   // the goal-box segments lie beyond the goal lines between y = 24 and y = 44: reachable only from x < 2 or x > 103
        const double bx = L.f(bo + kPX), by = L.f(bo + kPY);
        uint32_t gcand = ((bx < 2.0 ? 0x1C0u : 0u) | (bx > kWidth - 2.0 ? 0xE00u : 0u)) & ((by > 22.0 && by < 46.0) ? 0xFC0u : 0u);
#pragma unroll 1
        while (gcand != 0u && !goal) {
            const int sg = __ffs((int)gcand) - 1;
            gcand &= gcand - 1u;
            goal = ball_touches_segment(L, ball, sg);
        }
    }
    if (goal) {                                                          // :469-475
        const bool left_scored = L.f(bo + kPX) > kWidth - 2;
        reward = dadd(reward, left_scored ? 1000.0 : -1000.0);
        position_to_initial(L, P);
        s.stamp += 1;
        s.owner_side = (int)__umulhi(draw(P, env_id, s.t_total, j), 2u);
        res.flags |= kFlagGoal | (left_scored ? (int)kFlagGoalLeft : 0);
    }
    s.ep_step += 1;                                                      // :478-481
    s.t_total += 1;
    res.done = s.ep_step >= P.ep_limit ? 1 : 0;
    if (res.done) res.flags |= kFlagDone;
    res.reward = reward;
    return res;
}

}  // namespace v1
}  // namespace futbol
