"""CPU ORACLE (test infrastructure, NOT product code) -- reference-shaped scalar form.

One instance, one cycle at a time, a Python loop over control cycles, YARP ports
replaced by in-process fakes: the shape of the reference's own loop
(``scripts/bridge:600-634`` driving ``scripts/vf:312-466``, ``scripts/nullspace:162-184``,
``scripts/joint_p_controller:116-146`` and ``src/command_mixer.py:46-82``).  It exists to
(1) run BASELINE config 1 (single LWR, 1 goal, 3 obstacles, 1000 cycles), (2) be the
"reference CPU loop" leg of ``bench.py`` (kind "port"), and (3) cross-check
``oracle/batch.py``, which must give the same numbers.

The three un-vendored libraries are restated as small classes with the API surface the
reference uses: ``KdlFrame``/``kdl_diff``/``Twist`` (PyKDL),
``Lafik`` (arcospyu.robot_tools), ``vfl_library``/``VectorField``/``ScalarField`` (vfl).
PARITY UNPINNED for those three (see oracle/batch.py header and ORACLE_CHOICES);
``CommandMixer`` and the nullspace functions are pinned against the real reference code
by ``tests/golden/`` (``oracle/gen_golden.py``).
"""
from __future__ import annotations

import math
import time

import numpy as np
from numpy.linalg import norm, pinv, svd

from .batch import (FLAG_AT_GOAL, FLAG_CLAMPED, FLAG_NAN, FLAG_NS_LIMIT, J_ROTX, J_ROTY, J_ROTZ,
                    Params, _AXIS)


# ----------------------------------------------------------------------------- fake YARP

class Value:
    def __init__(self, v):
        self.v = v

    def asDouble(self):
        return float(self.v)

    def asInt(self):
        return int(self.v)

    def asString(self):
        return str(self.v)

    toString = asString

    def asList(self):
        return self.v


class Bottle:
    def __init__(self, items=()):
        self.items = list(items)

    def size(self):
        return len(self.items)

    def get(self, i):
        return Value(self.items[i])

    def clear(self):
        self.items = []

    def addDouble(self, v):
        self.items.append(float(v))

    def addInt(self, v):
        self.items.append(int(v))


class Port:
    """Latest-value port: ``read(False)`` returns the newest bottle once, then None."""

    def __init__(self):
        self._b = None

    def write_list(self, l):
        self._b = Bottle(l)

    def read(self, wait=False):
        b, self._b = self._b, None
        return b


# ----------------------------------------------------------------------------- PyKDL surface

class KdlFrame:
    def __init__(self, R=None, p=None):
        self.M = np.eye(3) if R is None else np.asarray(R, dtype=np.float64)
        self.p = np.zeros(3) if p is None else np.asarray(p, dtype=np.float64)

    def __mul__(self, o):
        return KdlFrame(self.M @ o.M, self.M @ o.p + self.p)


def listToKdlFrame(l):
    a = np.asarray(l, dtype=np.float64).reshape(4, 4)
    return KdlFrame(a[:3, :3].copy(), a[:3, 3].copy())


def kdlFrameToList(f):
    a = np.eye(4)
    a[:3, :3] = f.M
    a[:3, 3] = f.p
    return a.reshape(16).tolist()


def _axis_angle(R):
    """Same quaternion route as oracle.batch.rot_axis_angle (scalar)."""
    t = R[0, 0] + R[1, 1] + R[2, 2]
    if t > 0.0:
        s = math.sqrt(1.0 + t) * 2.0
        w, x, y, z = 0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s
    else:
        i = int(np.argmax([R[0, 0], R[1, 1], R[2, 2]]))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(max(1.0 + 2.0 * R[i, i] - t, 0.0)) * 2.0
        xyz = [0.0, 0.0, 0.0]
        xyz[i] = 0.25 * s
        xyz[j] = (R[j, i] + R[i, j]) / s
        xyz[k] = (R[k, i] + R[i, k]) / s
        w = (R[k, j] - R[j, k]) / s
        x, y, z = xyz
    if w < 0.0:
        w, x, y, z = -w, -x, -y, -z
    n = math.sqrt(x * x + y * y + z * z)
    ang = 2.0 * math.atan2(n, w)
    axis = np.array([x, y, z]) / n if n > 0.0 else np.zeros(3)
    return axis, ang


class Twist:
    def __init__(self, vel, rot):
        self.vel = np.asarray(vel, dtype=np.float64)
        self.rot = np.asarray(rot, dtype=np.float64)

    def RefPoint(self, v_base_AB):
        """KDL: ``Twist(vel + rot x v_base_AB, rot)``."""
        return Twist(self.vel + np.cross(self.rot, v_base_AB), self.rot)


def kdl_diff(A, B):
    """``PyKDL.diff(F_a, F_b)`` with dt = 1: ``vel = p_b - p_a``, ``rot = R_a * rotvec(R_a^T R_b)``."""
    axis, ang = _axis_angle(A.M.T @ B.M)
    return Twist(B.p - A.p, A.M @ (axis * ang))


# ----------------------------------------------------------------------------- Lafik surface

class Lafik:
    """``arcospyu.robot_tools.Lafik`` surface used by the reference."""

    def __init__(self, chain, prm: Params):
        self.chain = chain
        self.numJnts = chain.n_joints
        self.jnt_pos = [0.0] * self.numJnts
        self.joint_limits = [[float(a), float(b)] for a, b in zip(chain.q_lo, chain.q_hi)]
        self.tweights = np.diag(np.asarray(prm.w_task, dtype=np.float64))
        wj = np.ones(self.numJnts) if prm.w_joint is None else np.asarray(prm.w_joint, dtype=np.float64)[:self.numJnts]
        self.jweights = np.diag(wj)
        self.ik_lambda = prm.ik_lambda
        self._fk()

    # jntsList assignment triggers FK (scripts/vf:316)
    @property
    def jntsList(self):
        return list(self.jnt_pos)

    @jntsList.setter
    def jntsList(self, q):
        self.jnt_pos = [float(v) for v in q]
        self._fk()

    def _fk(self):
        ch = self.chain
        R = ch.base[:9].reshape(3, 3).copy()
        p = ch.base[9:12].copy()
        self._z, self._p = [], []
        for i in range(self.numJnts):
            jt = int(ch.joint_type[i])
            ax = _AXIS[jt]
            self._z.append(R[:, ax].copy())
            self._p.append(p.copy())
            qi = self.jnt_pos[i]
            if jt in (J_ROTX, J_ROTY, J_ROTZ):
                c, s = math.cos(qi), math.sin(qi)
                Rq = np.eye(3)
                a, b = (ax + 1) % 3, (ax + 2) % 3
                Rq[a, a], Rq[a, b], Rq[b, a], Rq[b, b] = c, -s, s, c
                R = R @ Rq
            else:
                p = p + R[:, ax] * qi
            p = p + R @ ch.tip[i, 9:12]
            R = R @ ch.tip[i, :9].reshape(3, 3)
        self.kdlframe = KdlFrame(R, p)
        self.frame = kdlFrameToList(self.kdlframe)

    def jac_list(self):
        self._fk()          # nullspace writes jnt_pos[i] directly (scripts/nullspace:165-166)
        N = self.numJnts
        J = np.zeros((6, N))
        pe = self.kdlframe.p
        for i in range(N):
            jt = int(self.chain.joint_type[i])
            if jt in (J_ROTX, J_ROTY, J_ROTZ):
                J[0:3, i] = np.cross(self._z[i], pe - self._p[i])
                J[3:6, i] = self._z[i]
            else:
                J[0:3, i] = self._z[i]
        return J.tolist()

    def get_limits(self):
        return self.joint_limits

    def set_tweights(self, W):
        self.tweights = np.asarray(W, dtype=np.float64)

    def set_jweights(self, W):
        self.jweights = np.asarray(W, dtype=np.float64)

    def getIKV(self, vel, rot):
        J = np.asarray(self.jac_list())
        Jw = self.tweights @ J @ self.jweights
        A = Jw @ Jw.T + self.ik_lambda ** 2 * np.eye(6)
        t = self.tweights @ np.concatenate([np.asarray(vel, dtype=np.float64), np.asarray(rot, dtype=np.float64)])
        return (self.jweights @ (Jw.T @ np.linalg.solve(A, t))).tolist()


# ----------------------------------------------------------------------------- vfl surface

class _NullField:
    def setParams(self, params):
        pass

    def getVector(self, frame):
        return np.zeros(6)

    def getScalar(self, frame):
        return (1.0, 1.0)


class _PointAttractor:
    """vfl type 1: params = 16 goal frame floats + slowdown distance (scripts/object_feeder:229-241)."""
    rot_slowdown = 0.09

    def setParams(self, params):
        g = np.asarray(params[:16], dtype=np.float64).reshape(4, 4)
        self.Rg, self.pg = g[:3, :3].copy(), g[:3, 3].copy()
        self.slow = float(params[16]) if len(params) > 16 else 0.03

    def _err(self, frame):
        f = np.asarray(frame, dtype=np.float64).reshape(4, 4)
        e = self.pg - f[:3, 3]
        dist = math.sqrt(float(e @ e))
        axis, ang = _axis_angle(self.Rg @ f[:3, :3].T)
        return e, dist, axis, ang

    def getVector(self, frame):
        e, dist, axis, _ = self._err(frame)
        return np.concatenate([e / dist if dist > 0.0 else np.zeros(3), axis])

    def getScalar(self, frame):
        _, dist, _, ang = self._err(frame)
        s0 = min(1.0, dist / self.slow) if self.slow > 0.0 else 1.0
        s1 = min(1.0, ang / self.rot_slowdown) if self.rot_slowdown > 0.0 else 1.0
        return (s0, s1)


class _DecayRepeller:
    """vfl type 2: params = xyz, radius, safe distance, decay order (scripts/object_feeder:326-333)."""

    def setParams(self, params):
        self.o = np.asarray(params[:3], dtype=np.float64)
        self.radius, self.safe, self.order = float(params[3]), float(params[4]), float(params[5])

    def getVector(self, frame):
        p = np.array([frame[3], frame[7], frame[11]], dtype=np.float64)
        dv = self.o - p
        d = math.sqrt(float(dv @ dv))
        out = np.zeros(6)
        if d > 0.0 and self.radius > 0.0:
            out[:3] = dv * ((self.radius / max(d, self.safe)) ** self.order / d)
        return out

    def getScalar(self, frame):
        return (1.0, 1.0)


class _HemisphereRepeller:
    """vfl type 4: params = xyz, normal, safe distance, decay order (scripts/object_feeder:335-354)."""

    def setParams(self, params):
        self.o = np.asarray(params[0:3], dtype=np.float64)
        self.n = np.asarray(params[3:6], dtype=np.float64)
        self.safe, self.order = float(params[6]), float(params[7])

    def getVector(self, frame):
        p = np.array([frame[3], frame[7], frame[11]], dtype=np.float64)
        out = np.zeros(6)
        nn = math.sqrt(float(self.n @ self.n))
        if nn > 0.0:
            nh = self.n / nn
            h = float((p - self.o) @ nh)
            out[:3] = -nh * (self.safe / max(h, self.safe)) ** self.order
        return out

    def getScalar(self, frame):
        return (1.0, 1.0)


class _FunnelAttractor:
    """vfl type 5: params = goal xyz, axis, cut angle, angle-decay order, cut distance, distance-decay order
    (scripts/object_feeder:262-280)."""

    def setParams(self, params):
        self.g = np.asarray(params[0:3], dtype=np.float64)
        self.a = np.asarray(params[3:6], dtype=np.float64)
        self.cut_a, self.ord_a, self.cut_d, self.ord_d = (float(x) for x in params[6:10])

    def getVector(self, frame):
        p = np.array([frame[3], frame[7], frame[11]], dtype=np.float64)
        out = np.zeros(6)
        an = math.sqrt(float(self.a @ self.a))
        if an == 0.0:
            return out
        ah = self.a / an
        r = p - self.g
        s = float(r @ ah)
        rp = r - s * ah
        rho = math.sqrt(float(rp @ rp))
        if rho == 0.0:
            return out
        theta = math.atan2(rho, s)
        R = math.sqrt(float(r @ r))
        wa = 1.0 if theta <= self.cut_a else (self.cut_a / theta) ** self.ord_a
        wd = 1.0 if R <= self.cut_d else (self.cut_d / R) ** self.ord_d
        out[:3] = -rp / rho * wa * wd
        return out

    def getScalar(self, frame):
        return (1.0, 1.0)


def vfl_library():
    """``vfl.vfl.vectorFieldLibrary()``: type id -> class (scripts/vf:146,238,283)."""
    return {0: _NullField, 1: _PointAttractor, 2: _DecayRepeller, 4: _HemisphereRepeller, 5: _FunnelAttractor}


class VectorField:
    """``+=``, ``* float``, ``.normCart()``, ``.getVector`` (scripts/vf:150,287-292)."""

    def __init__(self, fn):
        self.fn = fn

    def getVector(self, x):
        return self.fn(x)

    def __add__(self, o):
        a, b = self.fn, o.fn
        return VectorField(lambda x: a(x) + b(x))

    __iadd__ = __add__

    def __mul__(self, k):
        a = self.fn
        return VectorField(lambda x: a(x) * k)

    def normCart(self):
        a = self.fn

        def f(x):
            v = np.array(a(x), dtype=np.float64)
            n = math.sqrt(float(v[:3] @ v[:3]))
            if n > 0.0:
                v[:3] = v[:3] / n
            return v
        return VectorField(f)


class ScalarField:
    def __init__(self, fn):
        self.fn = fn

    def getScalar(self, x):
        return self.fn(x)

    def __mul__(self, o):
        a, b = self.fn, o.fn
        return ScalarField(lambda x: tuple(u * v for u, v in zip(a(x), b(x))))

    __imul__ = __mul__


# ----------------------------------------------------------------------------- command mixer (restated)

class CommandMixer:
    """Restatement of ``src/command_mixer.py:32-82`` (checked against the real class in tests)."""

    def __init__(self, ports, weight_port, n, guard_time, weights, clock=time.time):
        self.nChannels = n
        self.ports = ports
        self.weight_port = weight_port
        self.clock = clock
        if len(ports) != len(weights):
            print('wrong number of initial weights. Resetting to zeros.')
            self.weights = [0.0] * len(ports)
        else:
            self.weights = weights
        self.guard_time = guard_time
        self.last_command = [[0.0] * n] * len(self.ports)
        self.last_command_time = [self.clock()] * len(self.ports)
        self.nan_seen = []

    def read(self):
        if self.weight_port:
            b = self.weight_port.read(False)
            if b:
                for i in range(min(b.size(), len(self.ports))):
                    self.weights[i] = b.get(i).asDouble()
        for p in range(len(self.ports)):
            b = self.ports[p].read(False)
            if b and b.size() == self.nChannels:
                self.last_command_time[p] = self.clock()
                self.last_command[p] = [b.get(i).asDouble() for i in range(self.nChannels)]
            elif self.clock() - self.last_command_time[p] > self.guard_time:
                self.last_command[p] = [0.0] * self.nChannels
            elif b:
                print('wrong length for data bottle')
        self.nan_seen = [(i, j) for i, c in enumerate(self.last_command) for j, v in enumerate(c) if math.isnan(v)]
        result = [0.0] * self.nChannels
        for v, w in zip(self.last_command, self.weights):
            for i in range(len(v)):
                result[i] += v[i] * w
        return result


# ----------------------------------------------------------------------------- nullspace (restated)

class Nullspace:
    """Restatement of ``scripts/nullspace:67-131`` with its module globals as attributes."""

    def __init__(self, n_joints, ns_lambda=0.0):
        self.nJoints = n_joints
        self.sig = [1] * n_joints
        self.lastvec = np.zeros((n_joints, n_joints))
        self.ns_lambda = ns_lambda

    def restrict(self, P, J):
        pJ = np.asarray(P) @ np.asarray(J)
        if self.ns_lambda == 0.0:
            pJt = pinv(pJ)
        else:
            pJt = pJ.T @ np.linalg.inv(pJ @ pJ.T + self.ns_lambda ** 2 * np.eye(pJ.shape[0]))
        return np.eye(self.nJoints) - pJt @ pJ

    def nullspace(self, P, J):
        B = self.restrict(P, J)
        u, s, vh = svd(B.T)
        i = 0
        while i < self.nJoints and s[i] >= 1e-8:
            if norm(self.sig[i] * u[:, i] - self.lastvec[:, i]) > norm(self.sig[i] * u[:, i] + self.lastvec[:, i]):
                self.sig[i] = -self.sig[i]
            u[:, i] = u[:, i] * self.sig[i]
            self.lastvec[:, i] = u[:, i]
            i += 1
        return u[:, 0:i].T

    def move_in_nullspace(self, P, J, control):
        ns = self.nullspace(P, J)
        n = min(self.nJoints, len(control), ns.shape[0])
        qdot = np.zeros(self.nJoints)
        for i in range(n):
            qdot = qdot + ns[i, :] * control[i]
        return [float(qdot[i]) for i in range(self.nJoints)]

    @staticmethod
    def check_limits(q, qdot, limits):
        scale = 0.3
        margin = 0.0
        n = len(limits)
        for i in range(n):
            d = q[i] + scale * qdot[i]
            if d < limits[i][0] + margin or d > limits[i][1] - margin:
                return [0] * n, True
        return qdot, False


# ----------------------------------------------------------------------------- the loop

class ControlLoop:
    """Single-process restatement of the 4 hot-path processes with stubbed ports.

    Scene = one goal (field id 1, type 1, force +1) and M ``ObstacleP`` decay repellers
    (ids 5+n, type 2, force -10) in the parameter layouts ``object_feeder`` emits
    (``scripts/object_feeder:229-241,317-334``); composition follows ``scripts/vf:276-293``.
    Nullspace: ``ns_mode`` 1 applies the reference's ``restrict()`` projector to the
    limit-avoidance gradient (north_star's ``(I - J^+ J) qdot0``; equal to
    ``sum_i (u_i . qdot0) u_i`` over the reference's SVD basis when ns_lambda = 0,
    SURVEY.md App. C.2 step 6); ``ns_mode`` 2 is the reference's 4-float control interface.
    """

    def __init__(self, chain, prm: Params, q0, goal17, obstacles=(), jp_ref=None, extra_fields=()):
        self.chain, self.prm = chain, prm
        N = chain.n_joints
        self.N = N
        self.q = [float(v) for v in q0]
        self.lafik = Lafik(chain, prm)
        self.lafik_ns = Lafik(chain, prm)             # nullspace runs its own Lafik (scripts/nullspace:60)
        _PointAttractor.rot_slowdown = prm.rot_slowdown
        self.vfDB = vfl_library()
        self.vectorFields = {}
        self.vectorFields[1] = [prm.goal_force, 1, list(goal17)]
        for n, ob in enumerate(obstacles):
            ob = list(ob)
            if len(ob) == 4:
                ob = ob + [prm.obst_safe, prm.obst_order]
            self.vectorFields[5 + n] = [prm.obst_force, 2, ob]
        for fid, force, vtype, params in extra_fields:
            self.vectorFields[int(fid)] = [float(force), int(vtype), [float(x) for x in params]]
        self._compose()
        self.toolFrame = np.eye(4)
        t = np.asarray(prm.tool, dtype=np.float64)
        self.toolFrame[:3, :3] = t[:9].reshape(3, 3)
        self.toolFrame[:3, 3] = t[9:12]
        self.toolFrame = self.toolFrame.reshape(16).tolist()
        self.ns = Nullspace(N, prm.ns_lambda)
        self.control = list(prm.ns_control)
        self.ref = list(jp_ref) if jp_ref is not None else (
            list(prm.jp_ref)[:N] if prm.jp_ref is not None else [0.0] * N)
        self.ports = [Port() for _ in range(6)]
        self.mixer = CommandMixer(self.ports, None, N, 2.0, list(prm.mixer_w))
        self.last = {}

    def _compose(self):
        """scripts/vf:276-293."""
        vftemp = self.vfDB[0]()
        vftemp.setParams([])
        total = VectorField(vftemp.getVector)
        totalS = ScalarField(vftemp.getScalar)
        for num in self.vectorFields:
            force, tVF, params = self.vectorFields[num]
            f = self.vfDB[tVF]()
            f.setParams(params)
            total += VectorField(f.getVector) * force
            totalS *= ScalarField(f.getScalar)
        self.totalVF = total.normCart()
        self.totalSF = totalS

    def cycle(self):
        prm, N = self.prm, self.N
        q = list(self.q)                                        # bridge.read_pos -> /bridge/encoders
        flags = 0
        # ---- vf (scripts/vf:312-466)
        self.lafik.jntsList = q
        kdlframe = self.lafik.kdlframe
        newkdlframe = kdlframe * listToKdlFrame(self.toolFrame)
        diff = kdl_diff(newkdlframe, kdlframe)
        frame = kdlFrameToList(newkdlframe)
        velvector = self.totalVF.getVector(frame)
        scalars = self.totalSF.getScalar(frame)
        velPos = prm.speed_scale * scalars[0] * velvector[0:3]
        velRot = prm.speed_scale * scalars[1] * velvector[3:6]
        tw = Twist(velPos, velRot).RefPoint(diff.vel)
        qd_vf = self.lafik.getIKV(tw.vel, tw.rot)
        self.ports[0].write_list(qd_vf)
        # ---- nullspace (scripts/nullspace:162-184)
        if prm.ns_mode != 0:
            for i in range(N):
                self.lafik_ns.jnt_pos[i] = q[i]
            limits = self.lafik_ns.get_limits()
            J = np.asarray(self.lafik_ns.jac_list())
            if prm.ns_mode == 1:
                mid = 0.5 * (self.chain.q_lo + self.chain.q_hi)
                rng = self.chain.q_hi - self.chain.q_lo
                qd0 = -prm.ns_limit_gain * (np.asarray(q) - mid) / (rng * rng)
                B = self.ns.restrict(np.eye(6), J)             # scripts/nullspace:75-79
                qd_ns = [float(v) for v in (B @ qd0)]
            else:
                qd_ns = self.ns.move_in_nullspace(np.eye(6), J, self.control)
            qd_ns, bad = Nullspace.check_limits(q, qd_ns, limits)
            if bad:
                flags |= FLAG_NS_LIMIT
            qd_ns = [v * prm.ns_gain for v in qd_ns]
            self.ports[1].write_list(qd_ns)
        else:
            qd_ns = [0.0] * N
        # ---- joint_p_controller (scripts/joint_p_controller:79-89,116-146)
        limits = [[float(a), float(b)] for a, b in zip(self.chain.q_lo, self.chain.q_hi)]
        ref_out = list(self.ref)
        for i in range(len(limits)):
            if self.ref[i] < limits[i][0]:
                ref_out[i] = limits[i][0]
            elif self.ref[i] > limits[i][1]:
                ref_out[i] = limits[i][1]
        self.ref = ref_out
        error = np.asarray(self.ref) - np.asarray(q)
        qd_jp = (error * prm.jp_kp).tolist()
        if all(x < prm.jp_delta for x in error):
            flags |= FLAG_AT_GOAL
        self.ports[2].write_list(qd_jp)
        # ---- bridge: mixer + clamp (scripts/bridge:604,625 and :188-203)
        direct = all(w == 0 for w in self.mixer.weights) if prm.direct_control < 0 else bool(prm.direct_control)
        mix = self.mixer.read()
        if self.mixer.nan_seen:
            flags |= FLAG_NAN
        leading_vel = max(map(abs, mix))
        if leading_vel > prm.max_vel:
            ratio = prm.max_vel / leading_vel
            flags |= FLAG_CLAMPED
        else:
            ratio = 1.0
        qdot_lim = [v * ratio for v in mix]
        if prm.bridge_kind == 1:
            # Powercube_Bridge.set_vel, scripts/bridge:295-303 (statement for statement; `ratio` is the variable above)
            shoulder_vel = qdot_lim[0]
            if shoulder_vel > prm.shoulder_vel[0]:
                ratio = abs(prm.shoulder_vel[0] / shoulder_vel)
            elif shoulder_vel < prm.shoulder_vel[1]:
                ratio = abs(prm.shoulder_vel[1] / shoulder_vel)
            qdot_lim = [i * ratio for i in qdot_lim]
        cmd = [qdot_lim[i] if (direct or prm.bridge_kind != 0) else (-q[i] + q[i] + qdot_lim[i]) for i in range(N)]
        # ---- plant (external joint_sim): explicit Euler
        if prm.integrate:
            self.q = [q[i] + prm.dt * qdot_lim[i] for i in range(N)]
        self.last = dict(qdot_vf=qd_vf, qdot_ns=qd_ns, qdot_jp=qd_jp, qdot_mix=mix, qdot=qdot_lim,
                         cmd=cmd, pose=frame, flags=flags)
        return qdot_lim

    def run(self, k_cycles):
        for _ in range(k_cycles):
            self.cycle()
        return self.q
