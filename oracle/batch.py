"""CPU ORACLE (test infrastructure, NOT product code) -- vectorised FP64 numpy form.

Restates one vfclik control cycle for a batch of independent
instances.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline legs may import this package; the product (``vfclik_b200``) never does.

PARITY STATUS: **partially pinned.**  The parts of the path whose source is in the
reference and runs in a py3 container -- ``CommandMixer.read`` (src/command_mixer.py:46-82),
the ``scripts/nullspace`` functions (:67-131), the bridge back-ends' ``set_vel``
(scripts/bridge:182-210,288-312,507-530), ``joint_p_controller.check_limits`` (:79-89) and
``vf.get_weight_matrix`` (scripts/vf:164-179) -- are pinned by golden vectors
generated from the real reference code (``oracle/gen_golden.py`` ->
``tests/golden/``).  Everything behind the un-vendored, un-pinned PyKDL / arcospyu.Lafik /
vfl boundary is **parity unpinned**: the reference ships no test, fixture or golden
vector for it, so this file *defines* it (``ORACLE_CHOICES`` below) following public
KDL conventions and BASELINE.json's north_star formulas.

Every function cites the reference lines it follows.  Layout: leading axis is the
instance axis (``q[I, N]``), i.e. AoS-by-instance; the GPU uses the transposed SoA.
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np

ORACLE_CHOICES = [
    ("chain constants", "nominal KUKA LWR 4+ (0.31/0.40/0.39/0.078 m, +-pi/2 twists); the reference's config file is external"),
    ("FK / Jacobian", "public KDL conventions: pose_i = Joint_i(q_i)*F_tip_i; geometric Jacobian in the base frame, reference point = flange origin, rows [linear; angular]"),
    ("attractor (vfl type 1) linear part", "unit vector from tool position to goal position (zero at the goal)"),
    ("attractor rotational part", "unit rotation axis of R_goal * R_tool^T (base frame), extracted through the unit quaternion with w >= 0"),
    ("attractor scalars", "S0 = min(1, dist/slowdown) (1 if slowdown <= 0); S1 = min(1, angle/rot_slowdown)"),
    ("decay repeller (vfl type 2)", "V = (o - p)/d * (radius / max(d, safe))^order, zero rotational part, scalars (1, 1); force < 0 repels"),
    ("hemisphere repeller (vfl type 4)", "n = normal/|normal|, h = (p - o).n; V = -n * (safe / max(h, safe))^order, zero rotational part, scalars (1, 1)"),
    ("funnel attractor (vfl type 5)", "a = axis/|axis|, r = p - goal, s = r.a, r_perp = r - s a, theta = atan2(|r_perp|, s); V = -r_perp/|r_perp| * wa * wd with wa = 1 if theta <= cut_angle else (cut_angle/theta)^order_a and wd = 1 if |r| <= cut_dist else (cut_dist/|r|)^order_d"),
    ("normCart", "divides the translational part of the summed field by its Euclidean norm (zero stays zero); rotational part untouched"),
    ("velocity IK", "north_star closed form qdot = Wj Jw^T (Jw Jw^T + lambda^2 I)^-1 Wt t, Jw = Wt J Wj; lambda explicit"),
    ("nullspace projector", "ns_lambda > 0: B = I - J^T (J J^T + ns_lambda^2 I)^-1 J (north_star's damped form); ns_lambda = 0: the reference's own B = I - pinv(J) J with numpy's SVD pinv (scripts/nullspace:75-79)"),
    ("limit-avoidance qdot0", "qdot0_i = -k (q_i - mid_i) / (hi_i - lo_i)^2"),
    ("nullspace basis for the 4-float control interface", "the reference takes the sigma >= 1e-8 left singular vectors of B^T (scripts/nullspace:95-107): ANY orthonormal basis of null(J), k = N - rank(J) vectors -- for k = 1 unique up to sign, for k > 1 whatever rotation of the k-fold singular value 1 LAPACK's gesdd happens to return (not a property of the reference's algorithm). The oracle fixes it as columns 6..N-1 of the Q of a Householder QR of J^T with LAPACK's dgeqr2/dlarfg sign convention (== numpy.linalg.qr(J.T, 'complete')); each vector then follows the reference's sign-continuity rule against lastvec, the first cycle keeping the raw sign as the reference does (sig starts at +1). Pinned against the executed reference by the projector sum_i u_i u_i^T (unique) and, for k = 1, by the vector itself up to the first-cycle sign. J is assumed to have full row rank (k = N - 6); at an exactly rank-deficient posture the reference finds one more vector"),
    ("velocity IK vs KDL", "Lafik.getIKV wraps a KDL velocity solver whose source is not in the reference. [recollection, unverifiable here] KDL's ChainIkSolverVel_wdls applies lambda only to singular values below eps and is a plain weighted pseudo-inverse elsewhere, which would differ from north_star's uniformly damped form by O(lambda^2/sigma^2) away from singularities. Params.ik_mode = 1 evaluates that truncated form (SVD; FP64 oracle only) so a future Lafik comparison is one flag away; the product implements ik_mode 0, north_star's closed form"),
    ("plant", "explicit Euler q += dt * qdot_lim (the reference's integrator is the external joint_sim)"),
    ("synchronous cycle", "vf, nullspace and joint_p_controller all see the same q in one cycle (the reference's processes run asynchronously on latest values)"),
]

J_NONE, J_ROTX, J_ROTY, J_ROTZ, J_TRANSX, J_TRANSY, J_TRANSZ = range(7)

FLAG_AT_GOAL = 1       # joint_p_controller at_goal (scripts/joint_p_controller:135-146)
FLAG_NS_LIMIT = 2      # nullspace check_limits zeroed the command (scripts/nullspace:120-131)
FLAG_NAN = 4           # a NaN was seen in a mixer input (src/command_mixer.py:71-75)
FLAG_CLAMPED = 8       # bridge velocity clamp was active (scripts/bridge:188-196)


@dataclasses.dataclass
class Params:
    """Mirror of ``vfk_params`` (include/vfk.h); defaults = ``vfk_default_params``."""
    ik_lambda: float = 0.1
    ns_lambda: float = 0.1
    dt: float = 0.01
    speed_scale: float = 0.2
    max_vel: float = 1.0
    jp_kp: float = 1.5
    jp_delta: float = 0.087            # scripts/joint_p_controller:57
    ns_gain: float = 0.5               # scripts/nullspace:62
    ns_lookahead: float = 0.3          # scripts/nullspace:121
    ns_limit_gain: float = 1.0
    rot_slowdown: float = 0.09         # scripts/vf:138 (min_rot)
    goal_force: float = 1.0            # scripts/object_feeder:236
    obst_force: float = -10.0          # scripts/object_feeder:322
    obst_safe: float = 0.001           # scripts/object_feeder:331
    obst_order: float = 20.0           # old/README.old:75
    mixer_w: tuple = (1.0, 1.0, 0.0, 0.0, 0.0, 0.0)   # scripts/bridge:596
    w_task: tuple = (1.0,) * 6
    w_joint: Optional[tuple] = None    # None -> ones(N)
    tool: tuple = (1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0)   # identity (scripts/vf:154)
    jp_ref: Optional[tuple] = None     # None -> zeros(N); config.initial_joint_pos
    ns_control: tuple = (0.0, 0.0, 0.0, 0.0)          # scripts/nullspace:137
    ns_mode: int = 1                   # 0 off, 1 projector, 2 control (reference interface)
    direct_control: int = -1           # -1 auto: all mixer weights zero (scripts/bridge:604)
    integrate: int = 1
    bridge_kind: int = 0               # 0 LWR_Bridge, 1 Powercube_Bridge, 2 ICUB_Bridge (scripts/bridge:102-111)
    shoulder_vel: tuple = (0.0, 0.0)   # Powercube: config.max_vel_shoulder_pos (> 0), max_vel_shoulder_neg (< 0)
    ik_mode: int = 0                   # 0: north_star's uniformly damped closed form; 1: KDL-wdls-style (see ORACLE_CHOICES; oracle only)
    ik_eps: float = 1e-5               # ik_mode 1: singular values below this are damped, the rest inverted plainly


# --------------------------------------------------------------------------- FK / J

def _joint_rot(jtype: int, q: np.ndarray) -> np.ndarray:
    """Rotation of a revolute KDL joint about its own axis, [I,3,3]."""
    c, s = np.cos(q), np.sin(q)
    R = np.zeros(q.shape + (3, 3))
    R[..., 0, 0] = R[..., 1, 1] = R[..., 2, 2] = 1.0
    if jtype == J_ROTX:
        R[..., 1, 1], R[..., 1, 2], R[..., 2, 1], R[..., 2, 2] = c, -s, s, c
    elif jtype == J_ROTY:
        R[..., 0, 0], R[..., 0, 2], R[..., 2, 0], R[..., 2, 2] = c, s, -s, c
    elif jtype == J_ROTZ:
        R[..., 0, 0], R[..., 0, 1], R[..., 1, 0], R[..., 1, 1] = c, -s, s, c
    return R


_AXIS = {J_ROTX: 0, J_ROTY: 1, J_ROTZ: 2, J_TRANSX: 0, J_TRANSY: 1, J_TRANSZ: 2}


def fk_jac(chain, q: np.ndarray):
    """Flange frame and geometric Jacobian (a1, a7; ``scripts/vf:316-318``, ``scripts/nullspace:175``).

    ``T_{i+1} = T_i * Joint_i(q_i) * F_tip_i``.  Column i of J for a revolute joint with
    axis ``z_i = R_i e_axis`` through ``p_i``: ``[z_i x (p_e - p_i); z_i]``; prismatic: ``[z_i; 0]``.
    Returns R[I,3,3], p[I,3], J[I,6,N].
    """
    I, N = q.shape
    assert N == chain.n_joints
    R = np.broadcast_to(chain.base[:9].reshape(3, 3), (I, 3, 3)).copy()
    p = np.broadcast_to(chain.base[9:12], (I, 3)).copy()
    zs, ps = [], []
    for i in range(N):
        jt = int(chain.joint_type[i])
        ax = _AXIS[jt]
        zs.append(R[:, :, ax].copy())
        ps.append(p.copy())
        Rt = chain.tip[i, :9].reshape(3, 3)
        pt = chain.tip[i, 9:12]
        if jt in (J_ROTX, J_ROTY, J_ROTZ):
            R = R @ _joint_rot(jt, q[:, i])
        else:
            p = p + R[:, :, ax] * q[:, i:i + 1]
        p = p + R @ pt
        R = R @ Rt
    J = np.zeros((I, 6, N))
    for i in range(N):
        jt = int(chain.joint_type[i])
        if jt in (J_ROTX, J_ROTY, J_ROTZ):
            J[:, 0:3, i] = np.cross(zs[i], p - ps[i])
            J[:, 3:6, i] = zs[i]
        else:
            J[:, 0:3, i] = zs[i]
    return R, p, J


# --------------------------------------------------------------------------- field

def rot_axis_angle(Rerr: np.ndarray):
    """Unit axis and angle in [0, pi] of a rotation matrix via the unit quaternion.

    Shepperd's branch selection keeps the extraction well conditioned near pi, where
    the antisymmetric-part formula loses the axis.  Returns axis[I,3] (zero when the
    angle is zero), angle[I].
    """
    m = Rerr
    t = m[:, 0, 0] + m[:, 1, 1] + m[:, 2, 2]
    I = m.shape[0]
    quat = np.zeros((I, 4))   # w, x, y, z
    d = np.stack([m[:, 0, 0], m[:, 1, 1], m[:, 2, 2]], axis=1)
    br = np.where(t > 0.0, 3, np.argmax(d, axis=1))
    for b in range(4):
        sel = np.nonzero(br == b)[0]
        if sel.size == 0:
            continue
        ms = m[sel]
        if b == 3:
            s = np.sqrt(1.0 + t[sel]) * 2.0          # 4w
            quat[sel, 0] = 0.25 * s
            quat[sel, 1] = (ms[:, 2, 1] - ms[:, 1, 2]) / s
            quat[sel, 2] = (ms[:, 0, 2] - ms[:, 2, 0]) / s
            quat[sel, 3] = (ms[:, 1, 0] - ms[:, 0, 1]) / s
        else:
            i, j, k = b, (b + 1) % 3, (b + 2) % 3
            s = np.sqrt(np.maximum(1.0 + 2.0 * ms[:, i, i] - t[sel], 0.0)) * 2.0   # 4 x_i
            quat[sel, 1 + i] = 0.25 * s
            quat[sel, 1 + j] = (ms[:, j, i] + ms[:, i, j]) / s
            quat[sel, 1 + k] = (ms[:, k, i] + ms[:, i, k]) / s
            quat[sel, 0] = (ms[:, k, j] - ms[:, j, k]) / s
    neg = quat[:, 0] < 0.0
    quat[neg] = -quat[neg]
    n = np.linalg.norm(quat[:, 1:4], axis=1)
    angle = 2.0 * np.arctan2(n, quat[:, 0])
    with np.errstate(invalid="ignore", divide="ignore"):
        axis = np.where(n[:, None] > 0.0, quat[:, 1:4] / n[:, None], 0.0)
    return axis, angle


def aux_field_eval(pt: np.ndarray, aux: np.ndarray) -> np.ndarray:
    """Sum of force * V over the auxiliary field records aux[I, A, 12] = {type, force, p0..p9} (types 4 and 5, see
    ORACLE_CHOICES; parameter layouts of ``scripts/object_feeder:262-280,335-354``).  Returns [I, 3]."""
    I, A, _ = aux.shape
    out = np.zeros((I, 3))
    for s in range(A):
        typ, force, p = aux[:, s, 0].astype(int), aux[:, s, 1], aux[:, s, 2:12]
        r = pt - p[:, 0:3]
        an = np.linalg.norm(p[:, 3:6], axis=1)
        with np.errstate(invalid="ignore", divide="ignore"):
            a = np.where(an[:, None] > 0, p[:, 3:6] / an[:, None], 0.0)
            sa = np.einsum("ij,ij->i", r, a)
            # type 4
            v4 = -a * np.power(p[:, 6] / np.maximum(sa, p[:, 6]), p[:, 7])[:, None]
            # type 5
            rp = r - sa[:, None] * a
            rho = np.linalg.norm(rp, axis=1)
            theta = np.arctan2(rho, sa)
            R = np.linalg.norm(r, axis=1)
            wa = np.where(theta <= p[:, 6], 1.0, np.power(p[:, 6] / theta, p[:, 7]))
            wd = np.where(R <= p[:, 8], 1.0, np.power(p[:, 8] / R, p[:, 9]))
            v5 = np.where(rho[:, None] > 0, -rp / rho[:, None] * (wa * wd)[:, None], 0.0)
        ok = an > 0
        out += np.where(((typ == 4) & ok)[:, None], force[:, None] * v4, 0.0)
        out += np.where(((typ == 5) & ok)[:, None], force[:, None] * v5, 0.0)
    return out


def field_eval(prm: Params, Rt: np.ndarray, pt: np.ndarray, goal: np.ndarray, obst: Optional[np.ndarray],
               aux: Optional[np.ndarray] = None):
    """Composed vector field and scalar field at the tool frame (a3, a4).

    ``scripts/vf:276-293``: ``totalVF = normCart( null + sum_i force_i * V_i )``,
    ``totalSF = prod_i S_i``; ``scripts/vf:344-347``: ``velPos = speedScale*S0*V[0:3]``,
    ``velRot = speedScale*S1*V[3:6]``.  Field parameter layouts follow
    ``scripts/object_feeder:229-241`` (goal: 16-float frame + slowdown) and ``:317-334``
    (ObstacleP: xyz, radius, safe 0.001, order).

    goal[I,13]: R_goal row-major (9), p_goal (3), slowdown distance (1).
    obst[I,M,4|6]: x, y, z, radius [, safe, order]; radius 0 = inactive padding.
    Returns v[I,3], w[I,3].
    """
    I = pt.shape[0]
    Rg = goal[:, 0:9].reshape(I, 3, 3)
    pg = goal[:, 9:12]
    slow = goal[:, 12]
    e = pg - pt
    dist = np.linalg.norm(e, axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        unit = np.where(dist[:, None] > 0.0, e / dist[:, None], 0.0)
    axis, angle = rot_axis_angle(Rg @ np.transpose(Rt, (0, 2, 1)))
    V = np.zeros((I, 6))
    V[:, 0:3] = prm.goal_force * unit
    V[:, 3:6] = prm.goal_force * axis
    with np.errstate(invalid="ignore", divide="ignore"):
        S0 = np.where(slow > 0.0, np.minimum(1.0, dist / slow), 1.0)
    S1 = np.minimum(1.0, angle / prm.rot_slowdown) if prm.rot_slowdown > 0 else np.ones(I)
    if obst is not None and obst.shape[1] > 0:
        o = obst[:, :, 0:3]
        rad = obst[:, :, 3]
        if obst.shape[2] >= 6:
            safe, order = obst[:, :, 4], obst[:, :, 5]
        else:
            safe, order = prm.obst_safe, prm.obst_order
        dv = o - pt[:, None, :]
        d = np.linalg.norm(dv, axis=2)
        with np.errstate(invalid="ignore", divide="ignore"):
            decay = np.power(rad / np.maximum(d, safe), order)
            decay = np.where(rad > 0.0, decay, 0.0)
            w = np.where(d > 0.0, decay / d, 0.0)
        # sequential accumulation in obstacle order, like the += chain of scripts/vf:280-290
        for k in range(obst.shape[1]):
            V[:, 0:3] += prm.obst_force * (w[:, k:k + 1] * dv[:, k, :])
    if aux is not None and aux.shape[1] > 0:
        V[:, 0:3] += aux_field_eval(pt, np.asarray(aux, dtype=np.float64))
    n = np.linalg.norm(V[:, 0:3], axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        V[:, 0:3] = np.where(n[:, None] > 0.0, V[:, 0:3] / n[:, None], V[:, 0:3])
    v = prm.speed_scale * S0[:, None] * V[:, 0:3]
    w_ = prm.speed_scale * S1[:, None] * V[:, 3:6]
    return v, w_


# --------------------------------------------------------------------------- IK / nullspace

def ikv_dls(prm: Params, J: np.ndarray, tw: np.ndarray, N: int) -> np.ndarray:
    """``lafik.getIKV`` (a6, ``scripts/vf:461``) as north_star's closed form.

    ``qdot = Wj Jw^T (Jw Jw^T + lambda^2 I)^-1 Wt t``, ``Jw = Wt J Wj`` with the
    diagonal weights of ``set_tweights/set_jweights`` (``scripts/vf:296-309``).
    """
    wt = np.asarray(prm.w_task, dtype=np.float64)
    wj = np.ones(N) if prm.w_joint is None else np.asarray(prm.w_joint, dtype=np.float64)[:N]
    Jw = wt[None, :, None] * J * wj[None, None, :]
    if getattr(prm, "ik_mode", 0) == 1:
        # [recollection of KDL's ChainIkSolverVel_wdls] qdot = Wj V diag(f(sigma)) U^T Wt t with f = 1/sigma above eps and
        # sigma/(sigma^2 + lambda^2) below it
        U, sg, Vt = np.linalg.svd(Jw, full_matrices=False)
        with np.errstate(divide="ignore"):
            f = np.where(sg < prm.ik_eps, sg / (sg * sg + prm.ik_lambda ** 2), 1.0 / sg)
        y = np.einsum("ikr,ik->ir", U, wt[None, :] * tw) * f
        return wj[None, :] * np.einsum("irn,ir->in", Vt, y)
    A = Jw @ np.transpose(Jw, (0, 2, 1)) + (prm.ik_lambda ** 2) * np.eye(6)[None]
    y = np.linalg.solve(A, (wt[None, :] * tw)[..., None])[..., 0]
    return wj[None, :] * np.einsum("ikn,ik->in", Jw, y)


def ns_project(prm: Params, J: np.ndarray, x: np.ndarray) -> np.ndarray:
    """``B x`` with ``B = I - pinv(PJ) PJ``, ``P = I6`` (a8, ``scripts/nullspace:75-79,136``).

    ``ns_lambda = 0``: the reference's own form, ``pinv`` being numpy's SVD pseudo-inverse.
    ``ns_lambda > 0``: north_star's damped form ``J^T (J J^T + ns_lambda^2 I)^-1``.
    """
    if prm.ns_lambda == 0:
        Jp = np.linalg.pinv(J)                                   # [I, N, 6]
        return x - np.einsum("ink,ik->in", Jp, np.einsum("ikn,in->ik", J, x))
    A = J @ np.transpose(J, (0, 2, 1)) + (prm.ns_lambda ** 2) * np.eye(6)[None]
    Jx = np.einsum("ikn,in->ik", J, x)
    y = np.linalg.solve(A, Jx[..., None])[..., 0]
    return x - np.einsum("ikn,ik->in", J, y)


def ns_limit_gradient(prm: Params, chain, q: np.ndarray) -> np.ndarray:
    mid = 0.5 * (chain.q_lo + chain.q_hi)
    rng = chain.q_hi - chain.q_lo
    return -prm.ns_limit_gain * (q - mid[None, :]) / (rng * rng)[None, :]


def ns_ctrl_vectors(n_joints: int) -> int:
    """Basis vectors the four-float control interface can address: ``min(nJoints, len(control), k)`` with ``k = N - 6``
    (``scripts/nullspace:113``)."""
    return max(0, min(4, n_joints - 6))


def ns_basis(J: np.ndarray, n_vec: Optional[int] = None) -> np.ndarray:
    """Orthonormal vectors of null(J): columns ``6 .. 6 + n_vec - 1`` of the Q of a Householder QR of ``J^T`` [I, N, 6].

    Restates LAPACK's published ``dgeqr2`` / ``dlarfg`` (column by column: ``beta = -sign(alpha) |x|``,
    ``tau = (beta - alpha) / beta``, ``v = x / (alpha - beta)`` with ``v_0 = 1``, ``H = I - tau v v^T``), vectorised over the
    instance axis; ``tests/test_oracle.py`` checks it against ``numpy.linalg.qr(J.T, 'complete')`` and, through the
    projector ``sum_i u_i u_i^T``, against the executed ``scripts/nullspace:91-107`` (see ORACLE_CHOICES).
    Returns [I, n_vec, N] (default: all ``N - 6``)."""
    I, _, N = J.shape
    k = max(0, N - 6)
    n_vec = k if n_vec is None else min(n_vec, k)
    a = np.transpose(J, (0, 2, 1)).copy()                        # [I, N, 6]
    tau = np.zeros((I, 6))
    for j in range(min(6, N)):
        alpha = a[:, j, j].copy()
        xn2 = np.einsum("ir,ir->i", a[:, j + 1:, j], a[:, j + 1:, j])
        live = xn2 > 0.0
        beta = -np.copysign(np.sqrt(alpha * alpha + xn2), alpha)
        with np.errstate(invalid="ignore", divide="ignore"):
            t = np.where(live, (beta - alpha) / beta, 0.0)
            sc = np.where(live, 1.0 / (alpha - beta), 1.0)
        tau[:, j] = t
        a[:, j + 1:, j] *= sc[:, None]
        a[:, j, j] = np.where(live, beta, alpha)
        v = a[:, j + 1:, j]
        for c in range(j + 1, 6):
            w = (a[:, j, c] + np.einsum("ir,ir->i", v, a[:, j + 1:, c])) * t
            a[:, j, c] -= w
            a[:, j + 1:, c] -= v * w[:, None]
    out = np.zeros((I, n_vec, N))
    for i in range(n_vec):
        w = np.zeros((I, N))
        w[:, 6 + i] = 1.0
        for j in range(min(6, N) - 1, -1, -1):
            v = a[:, j + 1:, j]
            d = (w[:, j] + np.einsum("ir,ir->i", v, w[:, j + 1:])) * tau[:, j]
            w[:, j] -= d
            w[:, j + 1:] -= v * d[:, None]
        out[:, i, :] = w
    return out


def ns_control_motion(J: np.ndarray, lastvec: np.ndarray, control: np.ndarray):
    """``move_in_nullspace`` (a9, a10; ``scripts/nullspace:91-117``) on a batch.

    ``lastvec`` [I, kk * N] is the reference's ``lastvec`` state (vector i at ``[i * N, (i + 1) * N)``, zeros at the start),
    ``control`` [I, 4].  Each basis vector is flipped when ``|s u - last| > |s u + last|``, i.e. when it points away from
    the previous cycle's vector; then ``qdot = sum_{i < kk} control_i u_i``.  Returns (qdot [I, N], new lastvec)."""
    I, _, N = J.shape
    kk = ns_ctrl_vectors(N)
    if kk == 0:
        return np.zeros((I, N)), lastvec
    u = ns_basis(J, kk)                                           # [I, kk, N]
    last = lastvec.reshape(I, kk, N)
    dotl = np.einsum("ikn,ikn->ik", u, last)
    u = np.where((dotl < 0.0)[:, :, None], -u, u)
    qd = np.einsum("ik,ikn->in", control[:, :kk], u)
    return qd, u.reshape(I, kk * N)


def ns_basis_1d(prm: Params, J: np.ndarray, lastvec: np.ndarray) -> np.ndarray:
    """The single nullspace vector of a 6 x 7 Jacobian with the reference's sign continuity (kept for the 7-joint tests)."""
    assert J.shape[2] == 7
    return ns_control_motion(J, lastvec, np.ones((J.shape[0], 4)))[1]


def ns_check_limits(prm: Params, chain, q: np.ndarray, qdot: np.ndarray):
    """All-or-nothing lookahead limit check (a11, ``scripts/nullspace:120-131``)."""
    d = q + prm.ns_lookahead * qdot
    bad = np.any((d < chain.q_lo[None, :]) | (d > chain.q_hi[None, :]), axis=1)
    return np.where(bad[:, None], 0.0, qdot), bad


def bridge_set_vel(prm: Params, mix: np.ndarray, q: np.ndarray, qc: np.ndarray, direct: bool):
    """``set_vel`` of the three bridge back-ends on a batch: returns (qdot_lim, cmd, leading-clamp-active).

    LWR (``scripts/bridge:182-210``): ratio = max_vel / max|qdot| when exceeded; cmd = qdot_lim (direct) or
    ``-q_cmded + q + qdot_lim``.  Powercube (``:288-312``): the same clamp, then the shoulder-speed clamp of joint 0 --
    one ``ratio`` variable serves both, so when the shoulder limit is not hit the leading ratio is applied a second
    time (kept: bug-compatible); cmd = qdot_lim.  iCub (``:507-530``): leading clamp, cmd = qdot_lim.
    Pinned by golden vectors produced by executing those reference methods (``oracle/gen_golden.py:gen_bridge``)."""
    lead = np.max(np.abs(mix), axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        ratio = np.where(lead > prm.max_vel, prm.max_vel / lead, 1.0)
    qd = mix * ratio[:, None]
    if prm.bridge_kind == 1:
        sh = qd[:, 0]
        with np.errstate(invalid="ignore", divide="ignore"):
            r2 = np.where(sh > prm.shoulder_vel[0], np.abs(prm.shoulder_vel[0] / sh),
                          np.where(sh < prm.shoulder_vel[1], np.abs(prm.shoulder_vel[1] / sh), ratio))
        qd = qd * r2[:, None]
    cmd = qd if (direct or prm.bridge_kind != 0) else (-qc + q + qd)
    return qd, cmd, lead > prm.max_vel


# --------------------------------------------------------------------------- full cycle

def step(chain, prm: Params, q, goal, obst=None, jp_ref=None, ns_in=None, lastvec=None,
         q_cmded=None, ext_cmd=(None, None, None), k_cycles: int = 1, aux=None, jp_limits=None):
    """K synchronous control cycles.

    ``jp_limits = (lo[I, N], hi[I, N])``: what ``config.updateJntLimits(q)`` returned per instance (posture-dependent on the
    iCub, ``scripts/joint_p_controller:79-89``); the clamped reference then replaces the reference, as the loop's ``ref =
    check_limits(ref, q)`` does (``:121``), and is returned as ``jp_ref``.  Default: the chain's static limits.

    Returns a dict with the last cycle's ``qdot_vf, qdot_ns, qdot_jp, qdot_mix, qdot`` (clamped),
    ``cmd``, ``pose`` ([I,12]: R row-major, p of the tool frame), ``flags`` and the final ``q``
    (and ``lastvec`` in control mode).
    """
    q = np.array(q, dtype=np.float64, copy=True)
    I, N = q.shape
    goal = np.asarray(goal, dtype=np.float64)
    tool = np.asarray(prm.tool, dtype=np.float64)
    Rtool, ptool = tool[:9].reshape(3, 3), tool[9:12]
    w = np.asarray(prm.mixer_w, dtype=np.float64)
    direct = bool(np.all(w == 0.0)) if prm.direct_control < 0 else bool(prm.direct_control)
    if lastvec is not None:
        lastvec = np.array(lastvec, dtype=np.float64, copy=True)
    out = {}
    for _ in range(k_cycles):
        flags = np.zeros(I, dtype=np.int32)
        # 1. FK, tool compose, delta p (scripts/vf:316-331)
        R, p, J = fk_jac(chain, q)
        Rt = R @ Rtool
        pt = p + R @ ptool
        dp = p - pt                                   # PyKDL.diff(newkdlframe, kdlframe).vel
        # 2-3. field + saturation (scripts/vf:344-347)
        v, om = field_eval(prm, Rt, pt, goal, obst, aux)
        out_twist = np.concatenate([v, om], axis=1)
        # 4. Twist.RefPoint(dp): v' = v + w x dp (scripts/vf:456-459)
        tw = np.concatenate([v + np.cross(om, dp), om], axis=1)
        # 5. velocity IK (scripts/vf:461)
        qd_vf = ikv_dls(prm, J, tw, N)
        # 6. nullspace (scripts/nullspace:159-184)
        if prm.ns_mode == 0:
            qd_ns = np.zeros((I, N))
        else:
            if prm.ns_mode == 1:
                x = ns_limit_gradient(prm, chain, q) if ns_in is None else np.asarray(ns_in, dtype=np.float64)
                raw = ns_project(prm, J, x)
            else:
                if lastvec is None:
                    lastvec = np.zeros((I, max(1, ns_ctrl_vectors(N)) * N))
                ctrl = (np.broadcast_to(np.asarray(prm.ns_control, dtype=np.float64), (I, 4))
                        if ns_in is None else np.asarray(ns_in, dtype=np.float64))
                raw, lastvec = ns_control_motion(J, lastvec, ctrl)     # min(N, len(control), k) terms (scripts/nullspace:113-116)
            raw, bad = ns_check_limits(prm, chain, q, raw)
            flags |= np.where(bad, FLAG_NS_LIMIT, 0).astype(np.int32)
            qd_ns = raw * prm.ns_gain                  # scripts/nullspace:183
        # 7. joint P controller (scripts/joint_p_controller:79-89,126-146)
        if jp_ref is None:
            ref = np.broadcast_to(np.zeros(N) if prm.jp_ref is None else np.asarray(prm.jp_ref, dtype=np.float64)[:N], (I, N))
        else:
            ref = np.asarray(jp_ref, dtype=np.float64)
        lo, hi = (chain.q_lo[None, :], chain.q_hi[None, :]) if jp_limits is None else (np.asarray(jp_limits[0], dtype=np.float64),
                                                                                        np.asarray(jp_limits[1], dtype=np.float64))
        refc = np.where(ref < lo, lo, np.where(ref > hi, hi, ref))
        if jp_limits is not None and jp_ref is not None:
            jp_ref = refc                                # scripts/joint_p_controller:121: the clamp persists
        err = refc - q
        qd_jp = err * prm.jp_kp
        flags |= np.where(np.all(err < prm.jp_delta, axis=1), FLAG_AT_GOAL, 0).astype(np.int32)
        # 8. mixer (src/command_mixer.py:71-82; port order scripts/bridge:593-596)
        cmds = [qd_vf, qd_ns, qd_jp] + [np.zeros((I, N)) if e is None else np.asarray(e, dtype=np.float64) for e in ext_cmd]
        mix = np.zeros((I, N))
        nan = np.zeros(I, dtype=bool)
        for c, wp in zip(cmds, w):
            nan |= np.any(np.isnan(c), axis=1)
            mix = mix + c * wp
        flags |= np.where(nan, FLAG_NAN, 0).astype(np.int32)
        # 9. velocity clamp + command forming (scripts/bridge:188-203, 288-305, 507-530)
        qc = q if q_cmded is None else np.asarray(q_cmded, dtype=np.float64)
        qd, cmd, clamped = bridge_set_vel(prm, mix, q, qc, direct)
        flags |= np.where(clamped, FLAG_CLAMPED, 0).astype(np.int32)
        out = dict(qdot_vf=qd_vf, qdot_ns=qd_ns, qdot_jp=qd_jp, qdot_mix=mix, qdot=qd, cmd=cmd,
                   pose=np.concatenate([Rt.reshape(I, 9), pt], axis=1), flags=flags, twist=out_twist)
        # 10. plant: explicit Euler (joint_sim is external to the reference)
        if prm.integrate:
            q = q + prm.dt * qd
    out["q"] = q
    if jp_limits is not None and jp_ref is not None:
        out["jp_ref"] = np.asarray(jp_ref)
    if lastvec is not None:
        out["lastvec"] = lastvec
    return out
