"""A pure-Python stand-in for the part of ``pymunk`` 5.6 / Chipmunk2D 7.0 that gym_futbol.envs_v1 drives.

TEST INFRASTRUCTURE ONLY (oracle/README.md).  pymunk is a third-party dependency of the reference
(setup.py:5; 5.6.0 in the authors' run, colab_notebook.ipynb:118,130) that is neither vendored under
/root/reference nor installable here.  This module lets the reference's OWN, UNMODIFIED Python
(envs_v1/futbol_env.py, team.py, player.py, ball.py) execute: ``install()`` places it in ``sys.modules`` as
``pymunk`` (+ ``pymunk.vec2d``, ``pymunk.matplotlib_util``) exactly as oracle/ref_harness.py does for gym and
matplotlib.  What the reference's game logic computes -- actions, impulses, possession, out-of-bounds fix,
rewards, goals, re-kick-off, time limit, observation -- is then the reference's code; what ``space.step`` and
``shapes_collide`` compute is THIS file, i.e. the specification of DESIGN.md section 10, written from
knowledge of Chipmunk2D 7.0.x's sources (cpSpaceStep.c, cpArbiter.c, cpCollision.c, cpBody.c).  It is NOT the
real library: rows b3 / b6 of SURVEY.md section 8 stay "parity unpinned".

Structure follows Chipmunk (an arbiter object per shape pair with Chipmunk's state machine, a cached-arbiter
table filtered by time stamps, shape positions cached at ``space.step`` so that ``shapes_collide`` sees what the
C library would see) rather than the flat arrays of oracle/futbol_v1_oracle.c, so that the two are independent
restatements of the same specification:

  * body order = order of ``space.add``; arbiter (contact) order = circle/circle pairs (i < j) ascending
    j(j-1)/2 + i, then circle/segment pairs body-major in the segments' ``add`` order.  (Chipmunk's own order
    comes out of its BB-tree and is implementation-defined.)
  * narrow phase: CircleToCircle / CircleToSegment of cpCollision.c; the closest point of the (axis-aligned)
    segments is taken by clamping the centre's coordinate; end-cap contacts are kept (the reference sets no
    neighbour tangents).
  * cpArbiterPreStep: nMass = 1 / (m_inv_a + m_inv_b) (the rotational terms vanish: the contact offsets are
    parallel to the normal); bias = -biasCoef min(0, dist + slop) / dt with dist = (p2 - p1) . n;
    bounce = (v_b - v_a) . n  e_a e_b;  friction mu_a mu_b = 0 for every pair of this scene, so no tangential
    impulse and no spin.
  * cpSpace defaults as cpSpace.c writes them, with C float literals: collision_slop = 0.1f,
    collision_bias = pow(1.0f - 0.1f, 60.0f), collision_persistence = 3, iterations = 10.
  * cached impulses: a pair that touches again after 1 or 2 steps apart finds its cached arbiter, inherits
    jnAcc (contact hashes are all 0) but is in state FIRST_COLLISION, for which cpArbiterApplyCachedImpulse
    returns early: the warm-start impulse is applied only to pairs that also touched in the previous step.
  * velocity integration through each body's ``velocity_func`` (the reference's Python clamp, player.py:45-50,
    ball.py:49-54, which calls ``Body.update_velocity``: v = v damping^dt + (g + f m_inv) dt).

Arithmetic: Python floats = IEEE doubles, one operation per written operation; ``Vec2d.length`` is
``sqrt(x**2 + y**2)`` as in pymunk's vec2d.py (``**`` = libm pow: the oracle's ``arith = 1`` mode follows it).
"""
from __future__ import annotations

import math
import struct
import sys
import types


def _f32(x):
    return struct.unpack("f", struct.pack("f", x))[0]


class Vec2d:
    """pymunk.vec2d.Vec2d: the handful of operations the reference uses."""

    __slots__ = ("x", "y")

    def __init__(self, x=0.0, y=None):
        if y is None:
            x, y = x
        self.x, self.y = float(x), float(y)

    def __iter__(self):
        yield self.x
        yield self.y

    def __len__(self):
        return 2

    def __getitem__(self, i):
        return (self.x, self.y)[i]

    def __eq__(self, other):
        try:
            ox, oy = other
        except TypeError:
            return False
        return self.x == ox and self.y == oy

    def __ne__(self, other):
        return not self.__eq__(other)

    def __repr__(self):
        return "Vec2d(%r, %r)" % (self.x, self.y)

    def __add__(self, o):
        return Vec2d(self.x + o[0], self.y + o[1])

    def __sub__(self, o):
        return Vec2d(self.x - o[0], self.y - o[1])

    def __mul__(self, s):
        return Vec2d(self.x * s, self.y * s)

    __rmul__ = __mul__

    def __truediv__(self, s):
        return Vec2d(self.x / s, self.y / s)

    def __neg__(self):
        return Vec2d(-self.x, -self.y)

    def get_length_sqrd(self):
        return self.x**2 + self.y**2

    def get_length(self):
        return math.sqrt(self.x**2 + self.y**2)

    length = property(get_length)


def moment_for_circle(mass, inner_radius, outer_radius, offset=(0, 0)):
    """cpMomentForCircle: m (0.5 (r1^2 + r2^2) + |offset|^2)."""
    ox, oy = offset
    return mass * (0.5 * (inner_radius * inner_radius + outer_radius * outer_radius) + (ox * ox + oy * oy))


class Body:
    DYNAMIC, KINEMATIC, STATIC = 0, 1, 2

    def __init__(self, mass=0.0, moment=0.0, body_type=0):
        self.body_type = body_type
        if body_type == Body.DYNAMIC:
            self.mass, self.moment = float(mass), float(moment)
            self.m_inv = 1.0 / self.mass
        else:
            self.mass = self.moment = math.inf
            self.m_inv = 0.0
        self._p = [0.0, 0.0]
        self._v = [0.0, 0.0]
        self._v_bias = [0.0, 0.0]
        self._f = [0.0, 0.0]
        self.velocity_func = Body.update_velocity
        self.position_func = Body.update_position
        self.shapes = []
        self.space = None

    # cpBodySetPosition does NOT refresh the shapes' cached positions: collision queries see the new
    # position only after the next space.step (why the reference steps by 1e-4 after teleporting, :140-142)
    @property
    def position(self):
        return Vec2d(self._p[0], self._p[1])

    @position.setter
    def position(self, value):
        x, y = value
        self._p = [float(x), float(y)]

    @property
    def velocity(self):
        return Vec2d(self._v[0], self._v[1])

    @velocity.setter
    def velocity(self, value):
        x, y = value
        self._v = [float(x), float(y)]

    def apply_impulse_at_local_point(self, impulse, point=(0, 0)):
        """cpBodyApplyImpulseAtLocalPoint at the centre of gravity of a body that never rotates: v += j m_inv."""
        if tuple(point) != (0, 0):
            raise NotImplementedError("stand-in: impulses are applied at the centre only")
        jx, jy = impulse
        self._v[0] = self._v[0] + jx * self.m_inv
        self._v[1] = self._v[1] + jy * self.m_inv

    @staticmethod
    def update_velocity(body, gravity, damping, dt):
        """cpBodyUpdateVelocity: v = v damping + (g + f m_inv) dt; forces reset."""
        gx, gy = gravity
        body._v[0] = body._v[0] * damping + (gx + body._f[0] * body.m_inv) * dt
        body._v[1] = body._v[1] * damping + (gy + body._f[1] * body.m_inv) * dt
        body._f = [0.0, 0.0]

    @staticmethod
    def update_position(body, dt):
        """cpBodyUpdatePosition: p += (v + v_bias) dt; v_bias = 0."""
        body._p[0] = body._p[0] + (body._v[0] + body._v_bias[0]) * dt
        body._p[1] = body._p[1] + (body._v[1] + body._v_bias[1]) * dt
        body._v_bias = [0.0, 0.0]


class ContactPointSet:
    def __init__(self, normal, points):
        self.normal, self.points = normal, points


class _Contact:
    __slots__ = ("p1", "p2", "n_mass", "bias", "j_bias", "bounce", "jn_acc")

    def __init__(self, p1, p2):
        self.p1, self.p2 = p1, p2
        self.n_mass = self.bias = self.j_bias = self.bounce = self.jn_acc = 0.0


class Shape:
    def __init__(self, body):
        self.body = body
        self.elasticity = 0.0
        self.friction = 0.0
        self.collision_type = 0
        self.space = None
        if body is not None:
            body.shapes.append(self)

    def shapes_collide(self, other):
        """cpShapesCollide: the narrow phase on the CACHED shape positions."""
        info = _collide(self, other)
        if info is None:
            return ContactPointSet(Vec2d(0, 0), [])
        _, _, n, con = info
        return ContactPointSet(Vec2d(n[0], n[1]), [(con.p1, con.p2)])


class Circle(Shape):
    def __init__(self, body, radius, offset=(0, 0)):
        super().__init__(body)
        if tuple(offset) != (0, 0):
            raise NotImplementedError("stand-in: circles are centred on their body")
        self.radius = float(radius)
        self.tc = (0.0, 0.0)

    def cache(self):
        self.tc = (self.body._p[0], self.body._p[1])


class Segment(Shape):
    def __init__(self, body, a, b, radius):
        super().__init__(body)
        self.a, self.b, self.radius = (float(a[0]), float(a[1])), (float(b[0]), float(b[1])), float(radius)
        if not (self.a[0] == self.b[0] or self.a[1] == self.b[1]):
            raise NotImplementedError("stand-in: axis-aligned segments only (all the reference creates)")
        self.ta, self.tb = self.a, self.b

    def cache(self):
        if self.body.body_type != Body.STATIC:
            raise NotImplementedError("stand-in: segments live on the static body")
        self.ta, self.tb = self.a, self.b

    def closest(self, cx, cy):
        (ax, ay), (bx, by) = self.ta, self.tb
        if ax == bx:
            lo, hi = (ay, by) if ay < by else (by, ay)
            return ax, (lo if cy < lo else (hi if cy > hi else cy))
        lo, hi = (ax, bx) if ax < bx else (bx, ax)
        return (lo if cx < lo else (hi if cx > hi else cx)), ay

    def normal(self):
        """segment->tn = perp(normalize(b - a))"""
        sx, sy = self.tb[0] - self.ta[0], self.tb[1] - self.ta[1]
        sl = math.sqrt(sx * sx + sy * sy)
        return -(sy / sl), sx / sl


def _collide(a, b):
    """cpCollide for circle/circle and circle/segment: None, or (a, b, n, contact) with the circle first."""
    if isinstance(a, Segment):
        a, b = b, a
    if not isinstance(a, Circle):
        raise NotImplementedError("stand-in: segment/segment pairs do not occur")
    cx, cy = a.tc
    if isinstance(b, Circle):
        tx, ty = b.tc
    else:
        tx, ty = b.closest(cx, cy)
    dx, dy = tx - cx, ty - cy
    distsq = dx * dx + dy * dy
    mindist = a.radius + b.radius
    if not distsq < mindist * mindist:
        return None
    dist = math.sqrt(distsq)
    if dist != 0.0:
        inv = 1.0 / dist
        n = (dx * inv, dy * inv)
    else:
        n = (1.0, 0.0) if isinstance(b, Circle) else b.normal()
    p1 = (cx + n[0] * a.radius, cy + n[1] * a.radius)
    p2 = (tx + n[0] * (-b.radius), ty + n[1] * (-b.radius))
    return a, b, n, _Contact(p1, p2)


class _Arbiter:
    FIRST_COLLISION, NORMAL, IGNORE, CACHED = range(4)

    def __init__(self, a, b):
        self.a, self.b = a, b
        self.state = _Arbiter.FIRST_COLLISION
        self.stamp = 0
        self.contacts = []
        self.n = (0.0, 0.0)
        self.e = self.u = 0.0


class Space:
    def __init__(self, threaded=False):
        self.gravity = (0.0, 0.0)
        self.damping = 1.0
        self.iterations = 10
        self.collision_slop = _f32(0.1)                                   # cpSpace.c: 0.1f
        self.collision_bias = math.pow(_f32(_f32(1.0) - _f32(0.1)), 60.0)  # cpfpow(1.0f - 0.1f, 60.0f)
        self.collision_persistence = 3
        self.static_body = Body(body_type=Body.STATIC)
        self.bodies = []
        self.dynamic_shapes = []
        self.static_shapes = []
        self.stamp = 0
        self.curr_dt = 0.0
        self.arbiters = []
        self.cached_arbiters = {}
        self.step_log = []      # (dt, contacts solved) per step: instrumentation for the harness
        self.counters = {"warm_started": 0, "inherited_not_warm": 0, "new": 0}   # arbiters by kind, instrumentation

    def add(self, *objs):
        for o in objs:
            if isinstance(o, (list, tuple)):
                self.add(*o)
            elif isinstance(o, Body):
                o.space = self
                self.bodies.append(o)
            elif isinstance(o, Shape):
                o.space = self
                o.cache()                                                # cpSpaceAddShape -> cpShapeUpdate
                (self.static_shapes if o.body.body_type == Body.STATIC else self.dynamic_shapes).append(o)
            else:
                raise TypeError("stand-in: cannot add %r" % (o,))

    def debug_draw(self, options):
        raise NotImplementedError("stand-in: no drawing")

    # ---------------------------------------------------------------- cpSpaceStep
    def step(self, dt):
        if dt == 0.0:
            return
        self.stamp += 1
        prev_dt, self.curr_dt = self.curr_dt, dt
        for arb in self.arbiters:
            arb.state = _Arbiter.NORMAL
        self.arbiters = []
        for body in self.bodies:
            body.position_func(body, dt)
        for s in self.dynamic_shapes:
            s.cache()
        shapes = self.dynamic_shapes
        for j in range(1, len(shapes)):
            for i in range(j):
                self._collide_shapes(shapes[i], shapes[j])
        for s in shapes:
            for w in self.static_shapes:
                self._collide_shapes(s, w)
        # cpSpaceArbiterSetFilter
        for key in list(self.cached_arbiters):
            arb = self.cached_arbiters[key]
            ticks = self.stamp - arb.stamp
            if ticks >= 1 and arb.state != _Arbiter.CACHED:
                arb.state = _Arbiter.CACHED
            if ticks >= self.collision_persistence:
                arb.contacts = []
                del self.cached_arbiters[key]
        slop = self.collision_slop
        bias_coef = 1.0 - math.pow(self.collision_bias, dt)
        for arb in self.arbiters:
            self._pre_step(arb, dt, slop, bias_coef)
        damping = math.pow(self.damping, dt)
        for body in self.bodies:
            body.velocity_func(body, self.gravity, damping, dt)
        dt_coef = 0.0 if prev_dt == 0.0 else dt / prev_dt
        for arb in self.arbiters:
            self._apply_cached_impulse(arb, dt_coef)
        for _ in range(self.iterations):
            for arb in self.arbiters:
                self._apply_impulse(arb)
        self.step_log.append((dt, sum(len(a.contacts) for a in self.arbiters)))

    def _collide_shapes(self, sa, sb):
        info = _collide(sa, sb)
        if info is None:
            return
        a, b, n, con = info
        key = (id(a), id(b))
        arb = self.cached_arbiters.get(key)
        if arb is None:
            arb = self.cached_arbiters[key] = _Arbiter(a, b)
        # cpArbiterUpdate: contact hashes are 0 for both shape pairs, so an old contact always matches
        con.jn_acc = arb.contacts[0].jn_acc if arb.contacts else 0.0
        self.counters["new" if not arb.contacts else ("inherited_not_warm" if arb.state == _Arbiter.CACHED else "warm_started")] += 1
        arb.contacts = [con]
        arb.n = n
        arb.e = a.elasticity * b.elasticity
        arb.u = a.friction * b.friction
        if arb.u != 0.0:
            raise NotImplementedError("stand-in: frictionless pairs only (circles have friction 0)")
        if arb.state == _Arbiter.CACHED:
            arb.state = _Arbiter.FIRST_COLLISION
        self.arbiters.append(arb)
        arb.stamp = self.stamp

    @staticmethod
    def _pre_step(arb, dt, slop, bias_coef):
        a, b, (nx, ny) = arb.a.body, arb.b.body, arb.n
        for con in arb.contacts:
            con.n_mass = 1.0 / (a.m_inv + b.m_inv)
            dist = (con.p2[0] - con.p1[0]) * nx + (con.p2[1] - con.p1[1]) * ny
            m = dist + slop
            m = m if m < 0.0 else 0.0
            con.bias = -bias_coef * m / dt
            con.j_bias = 0.0
            con.bounce = ((b._v[0] - a._v[0]) * nx + (b._v[1] - a._v[1]) * ny) * arb.e

    @staticmethod
    def _apply_cached_impulse(arb, dt_coef):
        if arb.state == _Arbiter.FIRST_COLLISION:
            return
        a, b, (nx, ny) = arb.a.body, arb.b.body, arb.n
        for con in arb.contacts:
            jx, jy = nx * con.jn_acc * dt_coef, ny * con.jn_acc * dt_coef
            a._v[0] = a._v[0] - jx * a.m_inv
            a._v[1] = a._v[1] - jy * a.m_inv
            if b.body_type == Body.DYNAMIC:
                b._v[0] = b._v[0] + jx * b.m_inv
                b._v[1] = b._v[1] + jy * b.m_inv

    @staticmethod
    def _apply_impulse(arb):
        a, b, (nx, ny) = arb.a.body, arb.b.body, arb.n
        dyn = b.body_type == Body.DYNAMIC
        for con in arb.contacts:
            vbn = (b._v_bias[0] - a._v_bias[0]) * nx + (b._v_bias[1] - a._v_bias[1]) * ny
            vrn = (b._v[0] - a._v[0]) * nx + (b._v[1] - a._v[1]) * ny
            jbn = (con.bias - vbn) * con.n_mass
            jbn_old = con.j_bias
            t1 = jbn_old + jbn
            con.j_bias = t1 if t1 > 0.0 else 0.0
            jn = -(con.bounce + vrn) * con.n_mass
            jn_old = con.jn_acc
            t2 = jn_old + jn
            con.jn_acc = t2 if t2 > 0.0 else 0.0
            db, dj = con.j_bias - jbn_old, con.jn_acc - jn_old
            bx, by, jx, jy = nx * db, ny * db, nx * dj, ny * dj
            a._v_bias[0] = a._v_bias[0] - bx * a.m_inv
            a._v_bias[1] = a._v_bias[1] - by * a.m_inv
            a._v[0] = a._v[0] - jx * a.m_inv
            a._v[1] = a._v[1] - jy * a.m_inv
            if dyn:
                b._v_bias[0] = b._v_bias[0] + bx * b.m_inv
                b._v_bias[1] = b._v_bias[1] + by * b.m_inv
                b._v[0] = b._v[0] + jx * b.m_inv
                b._v[1] = b._v[1] + jy * b.m_inv


def install():
    """Place this module in sys.modules as pymunk (idempotent); returns the module object."""
    if "pymunk" in sys.modules and getattr(sys.modules["pymunk"], "__futbol_standin__", False):
        return sys.modules["pymunk"]
    if "pymunk" in sys.modules:
        raise RuntimeError("a real pymunk is already imported: use it instead of the stand-in")
    pm = types.ModuleType("pymunk")
    pm.__futbol_standin__ = True
    pm.__version__ = "5.6.0-standin"
    vec2d = types.ModuleType("pymunk.vec2d")
    vec2d.Vec2d = Vec2d
    mpl_util = types.ModuleType("pymunk.matplotlib_util")

    class DrawOptions:
        def __init__(self, ax):
            self.ax = ax

    mpl_util.DrawOptions = DrawOptions
    for name, obj in (("Space", Space), ("Body", Body), ("Circle", Circle), ("Segment", Segment), ("Shape", Shape),
                      ("Vec2d", Vec2d), ("moment_for_circle", moment_for_circle), ("ContactPointSet", ContactPointSet),
                      ("vec2d", vec2d), ("matplotlib_util", mpl_util)):
        setattr(pm, name, obj)
    sys.modules.update({"pymunk": pm, "pymunk.vec2d": vec2d, "pymunk.matplotlib_util": mpl_util})
    return pm
