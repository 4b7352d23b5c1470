"""Host-side chain description in the style of PyKDL (config files only).

The reference builds its robot model from ``config.segments`` through
``arcospyu.robot_tools.Lafik`` -> ``PyKDL.Chain`` (reference call sites
``scripts/vf:153``, ``scripts/nullspace:60``).  PyKDL is not available, and on
this path nothing is *computed* on the host: the classes here only let a
``config-<robot>-<instance>.py`` file *describe* a serial chain with the same
vocabulary (``Segment(Joint(Joint.RotZ), Frame(Rotation.RotX(a), Vector(..)))``)
so that :func:`vfclik_b200.config.load_config` can flatten it into the plain
``vfk_chain_desc`` the CUDA library takes.  FK / Jacobian / IK live in
``csrc/vfk_kernels.cu``.

Conventions (public KDL conventions):
``Frame * Frame``: ``R = R1 R2``, ``p = R1 p2 + p1``;
``Segment.pose(q) = Joint.pose(q) * f_tip``;
``Frame.DH_Craig1989(a, alpha, d, theta)``: ``R = RotX(alpha) RotZ(theta)``,
``p = (a, -sin(alpha) d, cos(alpha) d)``.
"""
from __future__ import annotations

import math
from typing import Iterable, List, Sequence


def _snap(x: float) -> float:
    """cos/sin of multiples of pi/2 come out as 6e-17; store exact 0/+-1."""
    r = round(x)
    return float(r) if abs(x - r) < 1e-15 else float(x)


class Vector:
    __slots__ = ("v",)

    def __init__(self, x: float = 0.0, y: float = 0.0, z: float = 0.0):
        self.v = [float(x), float(y), float(z)]

    def __getitem__(self, i):
        return self.v[i]

    def __setitem__(self, i, val):
        self.v[i] = float(val)

    def __iter__(self):
        return iter(self.v)

    def __add__(self, o):
        return Vector(*(a + b for a, b in zip(self.v, o.v)))

    def __sub__(self, o):
        return Vector(*(a - b for a, b in zip(self.v, o.v)))

    def __repr__(self):
        return "Vector(%r, %r, %r)" % tuple(self.v)


class Rotation:
    """Row-major 3x3 rotation."""
    __slots__ = ("m",)

    def __init__(self, *vals):
        if not vals:
            vals = (1, 0, 0, 0, 1, 0, 0, 0, 1)
        if len(vals) != 9:
            raise ValueError("Rotation takes 9 row-major values")
        self.m = [float(v) for v in vals]

    @staticmethod
    def Identity():
        return Rotation()

    @staticmethod
    def RotX(a):
        c, s = _snap(math.cos(a)), _snap(math.sin(a))
        return Rotation(1, 0, 0, 0, c, -s, 0, s, c)

    @staticmethod
    def RotY(a):
        c, s = _snap(math.cos(a)), _snap(math.sin(a))
        return Rotation(c, 0, s, 0, 1, 0, -s, 0, c)

    @staticmethod
    def RotZ(a):
        c, s = _snap(math.cos(a)), _snap(math.sin(a))
        return Rotation(c, -s, 0, s, c, 0, 0, 0, 1)

    def __getitem__(self, ij):
        i, j = ij
        return self.m[3 * i + j]

    def __setitem__(self, ij, val):
        i, j = ij
        self.m[3 * i + j] = float(val)

    def __mul__(self, o):
        if isinstance(o, Rotation):
            return Rotation(*[
                sum(self.m[3 * i + k] * o.m[3 * k + j] for k in range(3))
                for i in range(3) for j in range(3)
            ])
        if isinstance(o, Vector):
            return Vector(*[
                sum(self.m[3 * i + k] * o.v[k] for k in range(3))
                for i in range(3)
            ])
        return NotImplemented

    def Inverse(self):
        m = self.m
        return Rotation(m[0], m[3], m[6], m[1], m[4], m[7], m[2], m[5], m[8])


class Frame:
    __slots__ = ("M", "p")

    def __init__(self, a=None, b=None):
        if isinstance(a, Rotation) and isinstance(b, Vector):
            self.M, self.p = a, b
        elif isinstance(a, Rotation) and b is None:
            self.M, self.p = a, Vector()
        elif isinstance(a, Vector) and b is None:
            self.M, self.p = Rotation(), a
        elif a is None and b is None:
            self.M, self.p = Rotation(), Vector()
        else:
            raise TypeError("Frame(Rotation, Vector) | Frame(Rotation) | Frame(Vector) | Frame()")

    @staticmethod
    def Identity():
        return Frame()

    @staticmethod
    def DH_Craig1989(a, alpha, d, theta):
        ct, st = _snap(math.cos(theta)), _snap(math.sin(theta))
        ca, sa = _snap(math.cos(alpha)), _snap(math.sin(alpha))
        return Frame(
            Rotation(ct, -st, 0, st * ca, ct * ca, -sa, st * sa, ct * sa, ca),
            Vector(a, -sa * d, ca * d))

    def __mul__(self, o):
        if isinstance(o, Frame):
            return Frame(self.M * o.M, self.M * o.p + self.p)
        if isinstance(o, Vector):
            return self.M * o + self.p
        return NotImplemented

    def to_list16(self) -> List[float]:
        """Row-major 4x4, translation at indices 3, 7, 11 (``src/handlers.py:313-315``)."""
        m, p = self.M.m, self.p.v
        return [m[0], m[1], m[2], p[0], m[3], m[4], m[5], p[1],
                m[6], m[7], m[8], p[2], 0.0, 0.0, 0.0, 1.0]

    def to_list12(self) -> List[float]:
        """R row-major (9) followed by p (3): the layout ``vfk_chain_desc`` uses."""
        return list(self.M.m) + list(self.p.v)

    @staticmethod
    def from_list16(l: Sequence[float]) -> "Frame":
        if len(l) != 16:
            raise ValueError("expected 16 values")
        return Frame(Rotation(l[0], l[1], l[2], l[4], l[5], l[6], l[8], l[9], l[10]),
                     Vector(l[3], l[7], l[11]))


class Joint:
    # same enumeration order the C ABI uses (include/vfk.h: VFK_JOINT_*)
    NoJoint, RotX, RotY, RotZ, TransX, TransY, TransZ = range(7)
    # PyKDL spells the fixed joint ``Joint.None``; ``None`` is a keyword in py3.
    Fixed = NoJoint

    __slots__ = ("type",)

    def __init__(self, jtype=NoJoint):
        if jtype not in range(7):
            raise ValueError("unknown joint type %r" % (jtype,))
        self.type = jtype


class Segment:
    __slots__ = ("joint", "f_tip")

    def __init__(self, joint: Joint, f_tip: Frame | None = None):
        self.joint = joint
        self.f_tip = f_tip if f_tip is not None else Frame()


class Chain:
    def __init__(self, segments: Iterable[Segment] = ()):
        self.segments: List[Segment] = list(segments)

    def addSegment(self, seg: Segment):
        self.segments.append(seg)

    def getNrOfJoints(self) -> int:
        return sum(1 for s in self.segments if s.joint.type != Joint.NoJoint)

    def getNrOfSegments(self) -> int:
        return len(self.segments)
