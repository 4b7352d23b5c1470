"""Shared batched runtime behind the reference-shaped host modules.

The reference runs ``vf``, ``nullspace``, ``joint_p_controller`` and ``bridge`` as separate
processes that each recompute kinematics from the same ``/bridge/encoders`` sample
(SURVEY.md section 3.2).  Here they are thin port-facing objects over ONE
``ControlRuntime``: a host-buffer session of the fused CUDA kernel (``vfk_session_*``)
whose scene (goal + obstacle slots) stays resident on the GPU.  The first module that asks
for the results of a given ``q`` triggers one kernel launch; the others read the same
outputs (the synchronous idealisation of DESIGN.md section 1).

Scene protocol = ``scripts/vf:210-293``: fields are ``{id: [force, type, params]}``;
type 1 (point attractor, 16 frame floats + slowdown) becomes the instance's goal, type 2
(decay repeller: x, y, z, radius, safe, order) takes an obstacle slot, types 4 / 5
(hemisphere repeller / funnel attractor) take an auxiliary-field
slot.  Any other type gets the reference's "Unknown vector field type, ignoring" warning.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, Optional

import numpy as np

from .config import chain_from_config
from .engine import Engine, Params, NS_CONTROL, NS_OFF, NS_PROJECTOR  # noqa: F401

SUPPORTED_FIELD_TYPES = (0, 1, 2, 4, 5)
MAX_AUX_SLOTS = 4
CONFIG_MAX_SPEED_SCALE = 0.41          # scripts/vf:134


def dprint(*args):
    print(*args)


class ControlRuntime:
    def __init__(self, config, n_instances: int = 1, max_obstacles: int = 16, precision: int = 64, device: int = 0,
                 params: Optional[Params] = None, simulate_plant: bool = False):
        self.config = config
        self.chain = chain_from_config(config)
        self.N = self.chain.n_joints
        self.I = int(n_instances)
        self.M = int(max_obstacles)
        prm = params if params is not None else Params.from_config(config)
        prm = dataclasses.replace(prm, integrate=1 if simulate_plant else 0, goal_force=0.0)
        self.engine = Engine(self.chain, precision=precision, device=device, params=prm)
        self.session = self.engine.session(self.I, self.M, obst_ext=True)
        self.session.enable("qdot_vf", "qdot_ns", "qdot_jp", "cmd", "pose", "twist")
        dt = self.engine.np_dtype
        self.goal = np.zeros((13, self.I), dtype=dt)
        self.goal[0] = self.goal[4] = self.goal[8] = 1.0
        self.obst = np.zeros((self.M, self.I, 4), dtype=dt)            # radius 0 = empty slot
        self.obst_ext = np.zeros((self.M, self.I, 2), dtype=dt)
        self.obst_ext[:, :, 0] = prm.obst_safe
        self.obst_ext[:, :, 1] = prm.obst_order
        self.aux = np.zeros((MAX_AUX_SLOTS, 12, self.I), dtype=dt)       # type 0 = empty slot
        self.vectorFields: Dict[int, list] = {}                         # scripts/vf:145
        self._slot_of: Dict[int, int] = {}
        self._scene_dirty = True
        self._q_cached = None
        self._out = None
        self.cycles = 0
        self._mon = None
        self._mon_out = None
        self._mon_cycle = -1

    # ------------------------------------------------------------------ parameters
    @property
    def params(self) -> Params:
        return self.engine.params

    def set_params(self, **kw):
        self.engine.set_params(**kw)
        self._q_cached = None

    def set_speed_scale(self, v: float) -> bool:
        """``scripts/vf:197-207``: accepted iff 0 <= v <= 0.41."""
        if 0.0 <= v <= CONFIG_MAX_SPEED_SCALE:
            self.set_params(speed_scale=float(v))
            return True
        dprint("Value of speedScale not between 0.0 and config_max_vel. Ignoring")
        return False

    def set_max_vel(self, v: float, cap: float) -> bool:
        """``scripts/bridge:613-623``: accepted iff 0 <= v <= config.max_vel."""
        if 0.0 <= v <= cap:
            self.set_params(max_vel=float(v))
            return True
        print("Value of max_vel not between 0.0 and config.max_vel. Ignoring")
        return False

    def set_tool(self, frame16):
        f = [float(x) for x in frame16]
        self.set_params(tool=(f[0], f[1], f[2], f[4], f[5], f[6], f[8], f[9], f[10], f[3], f[7], f[11]))

    # ------------------------------------------------------------------ scene (scripts/vf:210-293)
    def add_field(self, field_id: int, force: float, vtype: int, params) -> bool:
        if vtype not in SUPPORTED_FIELD_TYPES:
            dprint("Unknown vector field type, ignoring")
            return False
        self.vectorFields[int(field_id)] = [float(force), int(vtype), [float(p) for p in params]]
        self._rebuild()
        return True

    def remove_field(self, field_id: int) -> bool:
        if int(field_id) in self.vectorFields:
            del self.vectorFields[int(field_id)]
            self._rebuild()
            return True
        return False

    def _rebuild(self):
        """Field list -> kernel scene arrays for instance 0 (the port-driven single-robot case)."""
        prm = self.params
        goal_force = 0.0
        self.obst[:, 0, :] = 0.0
        self.aux[:, :, 0] = 0.0
        self._slot_of.clear()
        slot = 0
        aslot = 0
        have_goal = False
        for fid in sorted(self.vectorFields):
            force, vtype, p = self.vectorFields[fid]
            if vtype == 1:
                if have_goal:
                    dprint("More than one point attractor: the GPU scene holds one goal per instance, ignoring field", fid)
                    continue
                if len(p) < 16:
                    dprint("Wrong number of values for a point attractor, ignoring")
                    continue
                g = self.goal[:, 0]
                g[0:9] = [p[0], p[1], p[2], p[4], p[5], p[6], p[8], p[9], p[10]]
                g[9:12] = [p[3], p[7], p[11]]
                g[12] = p[16] if len(p) > 16 else 0.03
                goal_force = force
                have_goal = True
            elif vtype == 2:
                if len(p) < 6:
                    dprint("Wrong number of values for a decay repeller, ignoring")
                    continue
                if slot >= self.M:
                    dprint("No free obstacle slot (max_obstacles = %d), ignoring field %d" % (self.M, fid))
                    continue
                radius, safe, order = p[3], p[4], p[5]
                if order <= 0.0:
                    dprint("Decay order must be > 0, ignoring field", fid)
                    continue
                # the kernel has one repeller force: fold force / obst_force into the radius (r'^n = ratio * r^n)
                ratio = force / prm.obst_force if prm.obst_force != 0.0 else 0.0
                if ratio <= 0.0:
                    dprint("Repeller force %g has the wrong sign for obst_force %g, ignoring field %d" % (force, prm.obst_force, fid))
                    continue
                self.obst[slot, 0, 0:3] = p[0:3]
                self.obst[slot, 0, 3] = radius * ratio ** (1.0 / order)
                self.obst_ext[slot, 0, 0] = safe
                self.obst_ext[slot, 0, 1] = order
                self._slot_of[fid] = slot
                slot += 1
            elif vtype in (4, 5):
                want = 8 if vtype == 4 else 10
                if len(p) < want:
                    dprint("Wrong number of values for field type %d, ignoring" % vtype)
                    continue
                if aslot >= MAX_AUX_SLOTS:
                    dprint("No free auxiliary field slot (%d), ignoring field %d" % (MAX_AUX_SLOTS, fid))
                    continue
                self.aux[aslot, 0, 0] = vtype
                self.aux[aslot, 1, 0] = force
                self.aux[aslot, 2:2 + want, 0] = p[:want]
                aslot += 1
        if goal_force != prm.goal_force:
            self.engine.set_params(goal_force=goal_force)
        self._scene_dirty = True
        self._q_cached = None

    # batched scene setters (many instances; bypass the single-robot field protocol)
    def set_goals(self, goal: np.ndarray, goal_force: float = 1.0):
        self.goal[...] = goal
        self.engine.set_params(goal_force=goal_force)
        self._scene_dirty = True
        self._q_cached = None

    def set_obstacles(self, obst: np.ndarray, obst_ext: Optional[np.ndarray] = None):
        self.obst[...] = 0.0
        self.obst[:obst.shape[0]] = obst
        if obst_ext is not None:
            self.obst_ext[:obst_ext.shape[0]] = obst_ext
        self._scene_dirty = True
        self._q_cached = None

    def set_jp_ref(self, ref):
        ref = np.asarray(ref, dtype=self.engine.np_dtype)
        if ref.ndim == 1:
            ref = np.repeat(ref[:, None], self.I, axis=1)
        self.session.set_jp_ref(ref)
        self._q_cached = None

    def set_ns_control(self, control4):
        c = [float(x) for x in control4][:4] + [0.0] * max(0, 4 - len(control4))
        self.set_params(ns_control=tuple(c))

    # ------------------------------------------------------------------ the cycle
    def cycle(self, q, k_cycles: int = 1) -> dict:
        """Outputs of the control cycle at joint positions ``q`` ([N] or [N, I]); cached per distinct q."""
        dt = self.engine.np_dtype
        q = np.asarray(q, dtype=dt)
        if q.ndim == 1:
            q = q[:, None]
        if q.shape != (self.N, self.I):
            raise ValueError("q must be [%d, %d], got %r" % (self.N, self.I, q.shape))
        if self._q_cached is not None and np.array_equal(self._q_cached, q) and k_cycles == 1:
            return self._out
        if self._scene_dirty:
            self._upload_scene()
        qdot = np.empty((self.N, self.I), dtype=dt)
        q_out = np.empty((self.N, self.I), dtype=dt)
        flags = np.empty(self.I, dtype=np.int32)
        self.session.cycle(q_in=q, k_cycles=k_cycles, qdot_out=qdot, q_out=q_out, flags_out=flags)
        self._out = dict(qdot=qdot, q=q_out, flags=flags, qdot_vf=self.session.read("qdot_vf"),
                         qdot_ns=self.session.read("qdot_ns"), qdot_jp=self.session.read("qdot_jp"),
                         cmd=self.session.read("cmd"), pose=self.session.read("pose"), twist=self.session.read("twist"))
        self._q_cached = q.copy()
        self.cycles += k_cycles
        return self._out

    def _upload_scene(self):
        self.session.set_goal(self.goal)
        self.session.set_obstacles(self.obst, self.obst_ext)
        used = int(np.any(self.aux[:, 0, :] != 0, axis=1).sum())
        self.session.set_aux(self.aux[:used] if used else None)
        self._scene_dirty = False

    # ------------------------------------------------------------------ monitoring
    def monitor(self, advance: bool = False) -> dict:
        """Tracking diagnostics (``scripts/vf:349-428``) and goal distance / majority state
        (``scripts/monitor_distance:76-84,148-219``) of the last cycle, from ONE ``vfk_monitor`` launch per cycle: the vf
        module publishes ``track`` on ``/track_error``, the distance monitor reads all three.  ``seen`` counts the frames
        the history holds (the reference publishes tracking errors from the 6th frame on)."""
        import torch
        # advance=True (the vf module, once per iteration with joint data -- the reference appends a frame every iteration,
        # moving or not) always evaluates; otherwise the result of the current cycle is reused
        if not advance and getattr(self, "_mon_cycle", -1) == self.cycles and self._mon_out is not None:
            return self._mon_out
        e = self.engine
        if getattr(self, "_mon", None) is None:
            self._mon = dict(f=e.alloc(32, self.I), i=e.alloc(6, self.I, dtype=torch.int32), track=e.alloc(8, self.I),
                             dist=e.alloc(2, self.I), state=e.alloc(2, self.I, dtype=torch.int32))
            self._mon_seen = 0
        m = self._mon
        b = _session_device_view(self)
        e.monitor(int(b.pose), int(b.twist), int(b.goal), m["f"], m["i"], self.I, track_out=m["track"], dist_out=m["dist"],
                  tracking_state_out=m["state"])
        self._mon_seen += 1

        def unblock(t):
            comps = t.shape[1]
            if t.dtype == torch.int32:
                return t.cpu().numpy().transpose(1, 0, 2).reshape(comps, -1)[:, :self.I]
            d = torch.empty((comps, self.I), dtype=t.dtype, device=t.device)
            e.unpack(t, d, comps, 1, self.I)
            return d.cpu().numpy()
        self._mon_out = dict(track=unblock(m["track"]), dist=unblock(m["dist"]), state=unblock(m["state"]), seen=self._mon_seen)
        self._mon_cycle = self.cycles
        return self._mon_out

    def close(self):
        self.session.close()
        self.engine.close()


def _session_device_view(runtime: "ControlRuntime"):
    import ctypes as C
    from ._lib import BuffersC
    b = BuffersC()
    runtime.engine._check(runtime.engine._lib.vfk_session_buffers(runtime.session._s, C.byref(b)))
    return b


def field_query(runtime: "ControlRuntime", pose12: np.ndarray) -> np.ndarray:
    """Twist the composed field commands at arbitrary tool poses (``scripts/vf:469-503``).

    pose12: [12, I] (R row-major, p).  Returns [6, I] = [v; w] already scaled by speedScale and the
    slowdown scalars, evaluated by ``vfk_field_eval`` on the session's resident scene.
    """
    import torch
    e = runtime.engine
    if runtime._scene_dirty:
        runtime._upload_scene()
    b = _session_device_view(runtime)
    dev = "cuda:%d" % e.device
    dense = torch.from_numpy(np.ascontiguousarray(pose12, dtype=e.np_dtype)).to(dev)
    pose_b = e.alloc(12, runtime.I)
    tw_b = e.alloc(6, runtime.I)
    tw_d = torch.empty((6, runtime.I), dtype=e.torch_dtype, device=dev)
    e.pack(dense, pose_b, 12, 1, runtime.I)
    e.field_eval(pose_b, int(b.goal), int(b.obst) if b.obst else None, tw_b, runtime.I, runtime.M,
                 obst_ext=int(b.obst_ext) if b.obst_ext else None, aux=int(b.aux) if b.aux else None, n_aux=int(b.n_aux))
    e.unpack(tw_b, tw_d, 6, 1, runtime.I)
    return tw_d.cpu().numpy()
