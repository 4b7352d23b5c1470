"""CPU ORACLE (test infrastructure, NOT product code) -- exact operation count of one control cycle.

SURVEY.md section 8(d) asks for the FLOP estimate ``F(N, M) ~ 187 N + 20 M + 560`` to be replaced by an exact count
of the oracle's arithmetic, frozen as a constant, with ``roofline.achieved`` of the compute-bound (K-fused) shape
computed from that constant.  This module *executes* the cycle of ``oracle.batch.step`` (same steps, same order,
same formulas; the test checks the values agree to 1e-12) on a counting scalar type and reports what it executed.

Counting convention (stated once, used everywhere):
* add, subtract, multiply, divide, square root: 1 FLOP each (an FMA is therefore 2);
* sin, cos, atan2, pow: counted separately as ``transcendentals`` (not FLOPs);
* comparisons, min / max / abs / negation / selects: free;
* arithmetic whose operands are all robot constants is free (it is folded on the host), and so are the structural
  identities ``x * 0``, ``x * (+-1)``, ``x + 0`` with a *constant* 0 / 1 -- so a DH-structured chain (twists of
  +-90 degrees, sparse offsets), identity weights and an identity tool cost what their non-trivial entries cost;
* symmetric matrices are formed once per distinct entry (21 of 36 for ``J J^T``); the 6x6 systems are solved by
  Cholesky (n^3/3-type count), and the nullspace projector reuses the factor when ``ns_lambda == ik_lambda`` and the
  weights are identity (same matrix).
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np

from . import batch as ob


class Counter:
    def __init__(self):
        self.flops = 0
        self.transcendentals = 0
        self.by_stage = {}
        self._stage = "misc"

    def stage(self, name):
        self._stage = name

    def add(self, n=1):
        self.flops += n
        self.by_stage[self._stage] = self.by_stage.get(self._stage, 0) + n

    def trans(self, n=1):
        self.transcendentals += n
        key = self._stage + ":transcendental"
        self.by_stage[key] = self.by_stage.get(key, 0) + n


_C: Optional[Counter] = None


def _v(x):
    return x.v if isinstance(x, F) else float(x)


class F:
    """A run-time scalar: every arithmetic operation on it is counted.  Plain Python floats are constants."""
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = float(v)

    def _bin(self, o, fn, ident=None, zero_absorbs=False):
        if not isinstance(o, F):
            o = float(o)
            if ident is not None and o == ident:
                return self
            if zero_absorbs and o == 0.0:
                return 0.0
            if zero_absorbs and o == -1.0:
                return F(-self.v)
        _C.add()
        return F(fn(self.v, _v(o)))

    def __add__(self, o): return self._bin(o, lambda a, b: a + b, ident=0.0)
    __radd__ = __add__
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b, ident=0.0)
    def __rsub__(self, o):
        if float(o) == 0.0:
            return F(-self.v)
        _C.add()
        return F(float(o) - self.v)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b, ident=1.0, zero_absorbs=True)
    __rmul__ = __mul__
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b, ident=1.0)
    def __rtruediv__(self, o):
        _C.add()
        return F(float(o) / self.v)
    def __neg__(self): return F(-self.v)
    def __lt__(self, o): return self.v < _v(o)
    def __gt__(self, o): return self.v > _v(o)
    def __le__(self, o): return self.v <= _v(o)
    def __ge__(self, o): return self.v >= _v(o)
    def __abs__(self): return F(abs(self.v))
    def __float__(self): return self.v


def fsqrt(x):
    _C.add()
    return F(math.sqrt(_v(x)))


def fsin(x):
    _C.trans()
    return F(math.sin(_v(x)))


def fcos(x):
    _C.trans()
    return F(math.cos(_v(x)))


def fatan2(y, x):
    _C.trans()
    return F(math.atan2(_v(y), _v(x)))


def fpow(x, y):
    _C.trans()
    return F(math.pow(_v(x), _v(y)))


def fmin(a, b):
    return a if _v(a) <= _v(b) else b


def fmax(a, b):
    return a if _v(a) >= _v(b) else b


def dot(a, b):
    s = 0.0
    for x, y in zip(a, b):
        s = s + x * y
    return s


def cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def norm3(a):
    return fsqrt(dot(a, a))


def matmul3(A, B):
    return [[dot(A[r], [B[0][c], B[1][c], B[2][c]]) for c in range(3)] for r in range(3)]


def chol_solve_factor(A):
    """Cholesky of a dense symmetric 6x6 given as rows (only the lower triangle is read)."""
    n = len(A)
    L = [[0.0] * n for _ in range(n)]
    for j in range(n):
        s = A[j][j]
        for k in range(j):
            s = s - L[j][k] * L[j][k]
        L[j][j] = fsqrt(s)
        for i in range(j + 1, n):
            s = A[i][j]
            for k in range(j):
                s = s - L[i][k] * L[j][k]
            L[i][j] = s / L[j][j]
    return L


def chol_solve(L, b):
    n = len(b)
    y = [0.0] * n
    for i in range(n):
        s = b[i]
        for k in range(i):
            s = s - L[i][k] * y[k]
        y[i] = s / L[i][i]
    x = [0.0] * n
    for i in reversed(range(n)):
        s = y[i]
        for k in range(i + 1, n):
            s = s - L[k][i] * x[k]
        x[i] = s / L[i][i]
    return x


def cycle(chain, prm: ob.Params, q, goal, obst, counter: Optional[Counter] = None):
    """One control cycle of ONE instance (steps 1-10 of ``oracle.batch.step``; projector nullspace on the built-in
    limit gradient, no auxiliary fields, no external mixer ports) on counting scalars.

    q [N], goal [13], obst [M, 4|6].  Returns (dict of plain-float results, Counter).
    """
    global _C
    _C = counter if counter is not None else Counter()
    C = _C
    N = chain.n_joints
    qf = [F(x) for x in q]
    g = [F(x) for x in goal]
    ob_rows = [[F(x) for x in row] for row in np.asarray(obst, dtype=np.float64)]

    # 1. FK + Jacobian (oracle.batch.fk_jac)
    C.stage("fk")
    R = [[float(chain.base[3 * r + c]) for c in range(3)] for r in range(3)]
    p = [float(chain.base[9 + k]) for k in range(3)]
    zs, ps = [], []
    for i in range(N):
        jt = int(chain.joint_type[i])
        ax = ob._AXIS[jt]
        zs.append([R[0][ax], R[1][ax], R[2][ax]])
        ps.append(list(p))
        tipR = [[float(chain.tip[i, 3 * r + c]) for c in range(3)] for r in range(3)]
        tipP = [float(chain.tip[i, 9 + k]) for k in range(3)]
        # snap the 6e-17 entries of cos(pi/2): structurally zero in a DH chain
        tipR = [[0.0 if abs(x) < 1e-15 else (1.0 if abs(x - 1) < 1e-15 else (-1.0 if abs(x + 1) < 1e-15 else x)) for x in row]
                for row in tipR]
        if jt in (ob.J_ROTX, ob.J_ROTY, ob.J_ROTZ):
            c_, s_ = fcos(qf[i]), fsin(qf[i])
            a, b = [(1, 2), (2, 0), (0, 1)][ax]        # the two columns the rotation mixes
            for r in range(3):
                ra, rb = R[r][a], R[r][b]
                R[r][a] = c_ * ra + s_ * rb
                R[r][b] = c_ * rb - s_ * ra
        else:
            p = [p[r] + R[r][ax] * qf[i] for r in range(3)]
        p = [p[r] + dot(R[r], tipP) for r in range(3)]
        R = matmul3(R, tipR)
    C.stage("jacobian")
    J = [[0.0] * N for _ in range(6)]
    for i in range(N):
        jt = int(chain.joint_type[i])
        if jt in (ob.J_ROTX, ob.J_ROTY, ob.J_ROTZ):
            lin = cross(zs[i], [p[k] - ps[i][k] for k in range(3)])
            for k in range(3):
                J[k][i] = lin[k]
                J[3 + k][i] = zs[i][k]
        else:
            for k in range(3):
                J[k][i] = zs[i][k]
    # tool compose (identity by default: constant 0 / 1 entries are free)
    C.stage("tool")
    tool = [float(x) for x in prm.tool]
    Rtool = [[tool[3 * r + c] for c in range(3)] for r in range(3)]
    identity_tool = tool == [1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0]
    Rt = R if identity_tool else matmul3(R, Rtool)
    pt = p if identity_tool else [p[r] + dot(R[r], tool[9:12]) for r in range(3)]
    dp = [0.0, 0.0, 0.0] if identity_tool else [p[k] - pt[k] for k in range(3)]

    # 2-3. field + saturation (oracle.batch.field_eval)
    C.stage("attractor")
    e = [g[9 + k] - pt[k] for k in range(3)]
    dist = norm3(e)
    unit = [x / dist for x in e]
    Rg = [[g[3 * r + c] for c in range(3)] for r in range(3)]
    E = [[dot(Rg[r], Rt[c]) for c in range(3)] for r in range(3)]         # Rg @ Rt^T
    t = E[0][0] + E[1][1] + E[2][2]
    if _v(t) > 0.0:
        s4 = fsqrt(1.0 + t) * 2.0
        qw = 0.25 * s4
        qx, qy, qz = (E[2][1] - E[1][2]) / s4, (E[0][2] - E[2][0]) / s4, (E[1][0] - E[0][1]) / s4
    else:
        d3 = [_v(E[0][0]), _v(E[1][1]), _v(E[2][2])]
        i = int(np.argmax(d3)); j = (i + 1) % 3; k = (i + 2) % 3
        s4 = fsqrt(fmax(1.0 + 2.0 * E[i][i] - t, 0.0)) * 2.0
        xyz = [0.0, 0.0, 0.0]
        xyz[i] = 0.25 * s4
        xyz[j] = (E[j][i] + E[i][j]) / s4
        xyz[k] = (E[k][i] + E[i][k]) / s4
        qw = (E[k][j] - E[j][k]) / s4
        qx, qy, qz = xyz
    if _v(qw) < 0.0:
        qw, qx, qy, qz = -qw, -qx, -qy, -qz
    nq = norm3([qx, qy, qz])
    angle = 2.0 * fatan2(nq, qw)
    axis = [qx / nq, qy / nq, qz / nq] if _v(nq) > 0 else [0.0, 0.0, 0.0]
    V = [prm.goal_force * u for u in unit]
    Vr = [prm.goal_force * a for a in axis]
    S0 = fmin(1.0, dist / g[12]) if _v(g[12]) > 0 else 1.0
    S1 = fmin(1.0, angle / prm.rot_slowdown) if prm.rot_slowdown > 0 else 1.0
    C.stage("repulsors")
    for row in ob_rows:
        dv = [row[k] - pt[k] for k in range(3)]
        d = norm3(dv)
        safe, order = (row[4], row[5]) if len(row) >= 6 else (prm.obst_safe, prm.obst_order)
        decay = fpow(row[3] / fmax(d, safe), order)
        w = decay / d
        V = [V[k] + prm.obst_force * (w * dv[k]) for k in range(3)]
    C.stage("normcart+saturation")
    n = norm3(V)
    V = [x / n for x in V]
    sv = prm.speed_scale * S0
    sr = prm.speed_scale * S1
    v = [sv * x for x in V]
    om = [sr * x for x in Vr]
    # 4. RefPoint
    C.stage("refpoint")
    wxd = cross(om, dp)
    tw = [v[k] + wxd[k] for k in range(3)] + om

    # 5. DLS velocity IK (oracle.batch.ikv_dls)
    C.stage("normal matrix")
    wt = [float(x) for x in prm.w_task]
    wj = [1.0] * N if prm.w_joint is None else [float(x) for x in prm.w_joint][:N]
    Jw = [[wt[r] * J[r][c] * wj[c] for c in range(N)] for r in range(6)]
    lam2 = prm.ik_lambda ** 2
    A = [[0.0] * 6 for _ in range(6)]
    for r in range(6):
        for c in range(r + 1):
            A[r][c] = dot(Jw[r], Jw[c]) + (lam2 if r == c else 0.0)
    C.stage("cholesky")
    L = chol_solve_factor(A)
    C.stage("ik solve")
    y = chol_solve(L, [wt[k] * tw[k] for k in range(6)])
    qd_vf = [wj[c] * dot([Jw[r][c] for r in range(6)], y) for c in range(N)]

    # 6. nullspace projector on the limit gradient (oracle.batch.ns_project, ns_limit_gradient, ns_check_limits)
    C.stage("nullspace")
    qd_ns = [0.0] * N
    ns_limit = False
    if prm.ns_mode == 1:
        mid = 0.5 * (chain.q_lo + chain.q_hi)
        rng = chain.q_hi - chain.q_lo
        x = [(qf[i] - float(mid[i])) * float(-prm.ns_limit_gain / (rng[i] * rng[i])) for i in range(N)]
        unit_w = all(w == 1.0 for w in wt) and all(w == 1.0 for w in wj)
        if unit_w and prm.ns_lambda == prm.ik_lambda:
            Lns = L
        else:
            An = [[0.0] * 6 for _ in range(6)]
            for r in range(6):
                for c in range(r + 1):
                    An[r][c] = dot(J[r], J[c]) + ((prm.ns_lambda ** 2) if r == c else 0.0)
            Lns = chol_solve_factor(An)
        Jx = [dot(J[r], x) for r in range(6)]
        yn = chol_solve(Lns, Jx)
        raw = [x[c] - dot([J[r][c] for r in range(6)], yn) for c in range(N)]
        look = [qf[i] + prm.ns_lookahead * raw[i] for i in range(N)]
        ns_limit = any(_v(look[i]) < chain.q_lo[i] or _v(look[i]) > chain.q_hi[i] for i in range(N))
        if ns_limit:
            raw = [0.0] * N
        qd_ns = [r * prm.ns_gain for r in raw]

    # 7. joint P controller
    C.stage("joint p")
    ref = [0.0] * N if prm.jp_ref is None else [float(x) for x in prm.jp_ref][:N]
    refc = [min(max(ref[i], float(chain.q_lo[i])), float(chain.q_hi[i])) for i in range(N)]
    w = [float(x) for x in prm.mixer_w]
    qd_jp = [0.0] * N
    if w[2] != 0.0:                       # nothing observes it otherwise
        err = [refc[i] - qf[i] for i in range(N)]
        qd_jp = [e_ * prm.jp_kp for e_ in err]
    # 8. mixer
    C.stage("mixer+clamp+euler")
    mix = [0.0] * N
    for cmd, wp in zip([qd_vf, qd_ns, qd_jp], w[:3]):
        mix = [mix[i] + cmd[i] * wp for i in range(N)]
    # 9. clamp
    lead = max(abs(_v(m)) for m in mix)
    ratio = (prm.max_vel / F(lead)) if lead > prm.max_vel else 1.0
    qd = [m * ratio for m in mix]
    # 10. Euler
    qn = [qf[i] + prm.dt * qd[i] for i in range(N)] if prm.integrate else qf
    out = dict(qdot=np.array([_v(x) for x in qd]), q=np.array([_v(x) for x in qn]),
               qdot_vf=np.array([_v(x) for x in qd_vf]), qdot_ns=np.array([_v(x) for x in qd_ns]))
    return out, C


def count(chain, n_obstacles: int, prm: Optional[ob.Params] = None, seed: int = 0, samples: int = 8):
    """Operation count of one instance-cycle at (chain, M): the count is data-independent up to the quaternion branch,
    the limit check and the clamp, so it is taken as the maximum over a few seeded samples (the branch-heavy side)."""
    prm = prm or ob.Params()
    rng = np.random.default_rng(seed)
    best = None
    for _ in range(samples):
        q = rng.uniform(0.9 * chain.q_lo, 0.9 * chain.q_hi)
        quat = rng.normal(size=4); quat /= np.linalg.norm(quat)
        w_, x, y, z = quat
        Rg = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w_), 2 * (x * z + y * w_)],
                       [2 * (x * y + z * w_), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w_)],
                       [2 * (x * z - y * w_), 2 * (y * z + x * w_), 1 - 2 * (x * x + y * y)]])
        goal = np.concatenate([Rg.reshape(9), rng.uniform(-0.5, 0.5, size=3) + chain.base[9:12], [0.05]])
        obst = np.concatenate([rng.uniform(-0.8, 0.8, size=(n_obstacles, 3)) + chain.base[9:12],
                               rng.uniform(0.03, 0.10, size=(n_obstacles, 1))], axis=1)
        _, c = cycle(chain, prm, q, goal, obst, Counter())
        if best is None or c.flops > best.flops:
            best = c
    return best


if __name__ == "__main__":
    import json
    import sys
    sys.path.insert(0, ".")
    from vfclik_b200 import workloads
    from vfclik_b200.config import PACKAGE_CONFIG_DIR, chain_from_config, config_filename, load_config
    lwr = chain_from_config(load_config(config_filename(PACKAGE_CONFIG_DIR + "/lwr/", "lwr", "right")))
    res = {}
    for name, ch, M in (("7x32", lwr, 32), ("7x256", lwr, 256), ("7x3", lwr, 3), ("17x64", workloads.dual_arm_torso_chain(), 64)):
        c = count(ch, M)
        res[name] = dict(flops=c.flops, transcendentals=c.transcendentals, by_stage=c.by_stage)
    print(json.dumps(res, indent=1))
