"""Robot config loading: ``config-<robot>-<instance>.py`` -> flat chain description.

The reference executes a Python config module through
``arcospyu.config_parser.ConfigFileParser(sys.argv)`` (``scripts/vf:63-64``,
``scripts/vfclik:80-85``) and reads the attributes listed in SURVEY.md App. B.5.
The config file itself is not part of the reference, so this package ships its
own (``config_data/lwr/config-lwr-right.py``) with the same attribute names.
"""
from __future__ import annotations

import dataclasses
import optparse
import os
import types
from typing import List, Sequence

import numpy as np

from . import kdl

VFK_MAX_JOINTS = 17   # include/vfk.h

PACKAGE_CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config_data")


@dataclasses.dataclass
class ChainDesc:
    """Flattened serial chain: what ``vfk_chain_desc`` (include/vfk.h) carries.

    ``pose_i(q_i) = Joint_i(q_i) * tip_i``; flange = ``base * prod_i pose_i``.
    Fixed segments of the config are folded into ``base`` / the preceding tip.
    Frames are 12 doubles: R row-major (9) then p (3).
    """
    n_joints: int
    base: np.ndarray          # [12]
    joint_type: np.ndarray    # [N] int32, kdl.Joint.* codes (never NoJoint)
    tip: np.ndarray           # [N, 12]
    q_lo: np.ndarray          # [N]
    q_hi: np.ndarray          # [N]

    def validate(self):
        n = self.n_joints
        if not (1 <= n <= VFK_MAX_JOINTS):
            raise ValueError("n_joints must be in 1..%d, got %d" % (VFK_MAX_JOINTS, n))
        if self.tip.shape != (n, 12) or self.joint_type.shape != (n,):
            raise ValueError("chain arrays do not match n_joints")
        if self.q_lo.shape != (n,) or self.q_hi.shape != (n,):
            raise ValueError("limit arrays do not match n_joints")
        if np.any(self.q_lo >= self.q_hi):
            raise ValueError("joint limits must satisfy lo < hi")
        return self


def chain_from_segments(segments: Sequence[kdl.Segment], limits: Sequence[Sequence[float]]) -> ChainDesc:
    base = kdl.Frame()
    tips: List[kdl.Frame] = []
    types_: List[int] = []
    for seg in segments:
        if seg.joint.type == kdl.Joint.NoJoint:
            if tips:
                tips[-1] = tips[-1] * seg.f_tip
            else:
                base = base * seg.f_tip
        else:
            types_.append(seg.joint.type)
            tips.append(seg.f_tip)
    n = len(tips)
    lim = np.asarray(limits, dtype=np.float64)
    if lim.shape != (n, 2):
        raise ValueError("limits must be [nJoints][2], got %r for %d joints" % (lim.shape, n))
    return ChainDesc(
        n_joints=n,
        base=np.asarray(base.to_list12(), dtype=np.float64),
        joint_type=np.asarray(types_, dtype=np.int32),
        tip=np.asarray([t.to_list12() for t in tips], dtype=np.float64).reshape(n, 12),
        q_lo=lim[:, 0].copy(),
        q_hi=lim[:, 1].copy(),
    ).validate()


def load_config(filename: str) -> types.SimpleNamespace:
    """Execute a config file and return its globals as a namespace.

    Mirrors what ``ConfigFileParser.get_all()`` hands to every module as
    ``config`` (``scripts/vf:63-64``).  Raises ``FileNotFoundError`` when the
    file is missing (the launcher turns that into exit code -1,
    ``scripts/vfclik:83-85``).
    """
    if not os.path.exists(filename):
        raise FileNotFoundError(filename)
    glb = {"__file__": os.path.abspath(filename), "__name__": "vfclik_config"}
    with open(filename, "r") as fh:
        code = compile(fh.read(), filename, "exec")
    exec(code, glb)
    glb.pop("__builtins__", None)
    cfg = types.SimpleNamespace(**glb)
    for attr in ("nJoints", "segments", "limits", "robotarm_portbasename"):
        if not hasattr(cfg, attr):
            raise AttributeError("config %s lacks required attribute %r" % (filename, attr))
    return cfg


def chain_from_config(config) -> ChainDesc:
    chain = chain_from_segments(config.segments, config.limits)
    if chain.n_joints != config.nJoints:
        raise ValueError("config.nJoints=%d but segments hold %d joints" % (config.nJoints, chain.n_joints))
    return chain


def config_filename(config_dir: str, robot: str, instance: str) -> str:
    """``scripts/vfclik:80-81``: plain string concatenation, no path join."""
    return config_dir + "config-" + robot + "-" + instance + ".py"


class ConfigFileParser:
    """Stand-in for ``arcospyu.config_parser.ConfigFileParser`` (module CLIs).

    ``ConfigFileParser(sys.argv).get_all() -> (options, args, config)`` with
    ``-c <config file>``, ``-n <namespace>`` (``scripts/vfclik:88-105``) and the
    bridge's ``-s`` (``scripts/bridge:59-67``).
    """

    def __init__(self, argv: Sequence[str], extra_options=()):
        parser = optparse.OptionParser("usage: %prog [options]")
        parser.add_option("-c", "--config_filename", dest="config_filename", default=None, type="string")
        parser.add_option("-n", "--namespace", dest="namespace", default="", type="string")
        parser.add_option("-s", "--simulation", action="store_true", dest="sim", default=False)
        for opt_args, opt_kwargs in extra_options:
            parser.add_option(*opt_args, **opt_kwargs)
        self.options, self.args = parser.parse_args(list(argv)[1:])
        if self.options.config_filename is None:
            parser.error("a config file is required (-c)")
        self.config = load_config(self.options.config_filename)

    def get_all(self):
        return self.options, self.args, self.config
