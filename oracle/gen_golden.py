"""Generate ``tests/golden/*.npz`` by running the REAL reference code (test infrastructure).

Run in the build container, where ``/root/reference`` exists:

    python -m oracle.gen_golden

These pieces of the reference's hot path run under python 3:

* ``src/command_mixer.py::CommandMixer`` -- imported with ``yarp`` and
  ``arcospyu.config_parser`` stubbed (they are only used for ports / CLI parsing);
* ``scripts/nullspace`` functions ``matrixrank, restrict, sign, nullspace,
  move_in_nullspace, check_limits`` (``:67-131``) -- the module cannot be imported
  (it opens YARP ports at import time and numpy 2 removed ``numpy.mat``), so the
  ``FunctionDef`` nodes are extracted with ``ast`` and executed with
  ``mat = numpy.asmatrix`` and the module globals ``nJoints, sig, lastvec``.

* the three bridge back-ends' ``set_vel`` and ``ICUB_Bridge.set_torso_cjoints`` (``scripts/bridge``),
  ``joint_p_controller.check_limits`` and ``vf.get_weight_matrix`` -- those files hold python-2 ``print`` statements
  elsewhere, so the wanted ``def`` is cut out by indentation and executed on its own with stand-in ports.

Nothing is copied from the reference: its code is executed where it lies and only the
inputs and outputs are stored.  The vectors pin ``oracle/refshape.py``'s restatements
(``CommandMixer``, ``Nullspace``) and, through them, ``oracle/batch.py`` and the GPU.
"""
from __future__ import annotations

import ast
import io
import os
import sys
import types
from contextlib import redirect_stdout

import numpy as np

REF = os.environ.get("VFCLIK_REFERENCE", "/root/reference")
OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


# ------------------------------------------------------------------------- loading the real code

def load_real_command_mixer():
    yarp = types.ModuleType("yarp")
    arcospyu = types.ModuleType("arcospyu")
    cp = types.ModuleType("arcospyu.config_parser")
    cp.ConfigFileParser = object
    arcospyu.config_parser = cp
    saved = {k: sys.modules.get(k) for k in ("yarp", "arcospyu", "arcospyu.config_parser")}
    sys.modules.update({"yarp": yarp, "arcospyu": arcospyu, "arcospyu.config_parser": cp})
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_command_mixer", os.path.join(REF, "src", "command_mixer.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod.CommandMixer


def load_real_nullspace(n_joints: int) -> dict:
    src = open(os.path.join(REF, "scripts", "nullspace")).read()
    tree = ast.parse(src)
    wanted = {"matrixrank", "restrict", "sign", "nullspace", "move_in_nullspace", "check_limits"}
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    assert {n.name for n in body} == wanted
    mod = ast.Module(body=body, type_ignores=[])
    from numpy import eye, matrix, zeros, sum, where
    from numpy.linalg import norm, pinv, svd
    glb = dict(svd=svd, norm=norm, pinv=pinv, sum=sum, where=where, mat=np.asmatrix, eye=eye, zeros=zeros,
               matrix=matrix, nJoints=n_joints, sig=[1] * n_joints, lastvec=np.asmatrix(zeros((n_joints, n_joints))))
    exec(compile(mod, "<reference scripts/nullspace>", "exec"), glb)
    return glb


class FakeValue:
    def __init__(self, v):
        self.v = v

    def asDouble(self):
        return float(self.v)


class FakeBottle:
    def __init__(self, vals):
        self.vals = list(vals)

    def size(self):
        return len(self.vals)

    def get(self, i):
        return FakeValue(self.vals[i])


class FakePort:
    def __init__(self):
        self.q = []

    def push(self, vals):
        self.q.append(FakeBottle(vals))

    def read(self, wait=False):
        return self.q.pop(0) if self.q else None


class ScriptIn:
    """Input port stand-in for the executed loops: messages queued by the script, ``None`` when there is nothing."""

    def __init__(self):
        self.q = []

    def read(self, wait=False):
        return self.q.pop(0) if self.q else None


class ScriptOut:
    """Output port stand-in: every written bottle is kept as a plain (nested) list."""

    def __init__(self):
        self.sent, self.b = [], None

    def prepare(self):
        from vfclik_b200 import ports as yarp
        self.b = yarp.Bottle()
        return self.b

    def write(self, *a):
        self.sent.append(self.b.to_list())

    writeStrict = write


def compile_iteration(body: str, global_names, label: str):
    """One pass through a reference loop body as a function: ``while True: <body>; break`` keeps its ``continue`` legal."""
    head = "def _iteration():\n" + ("    global " + ", ".join(global_names) + "\n" if global_names else "")
    return compile(head + "    while True:\n" + "".join("        " + l + "\n" for l in body.splitlines()) + "        break\n",
                   label, "exec")


# ------------------------------------------------------------------------- generators

def gen_mixer(rng):
    """A scripted sequence of port events through the real CommandMixer.read()."""
    CommandMixer = load_real_command_mixer()
    import time as _time
    n, n_ports, steps = 7, 6, 12
    now = [1000.0]
    real_time = _time.time
    _time.time = lambda: now[0]
    try:
        ports = [FakePort() for _ in range(n_ports)]
        wport = FakePort()
        mixer = CommandMixer(ports, wport, n, 2.0, [1.0, 1.0, 0.0, 0.0, 0.0, 0.0])
        events = np.full((steps, n_ports, n), np.nan)      # NaN row = no bottle on that port in that step
        wevents = np.full((steps, n_ports), np.nan)
        short = np.zeros((steps, n_ports), dtype=bool)      # a wrong-length bottle arrived
        dts = np.zeros(steps)
        outs = np.zeros((steps, n))
        weights_after = np.zeros((steps, n_ports))
        for s in range(steps):
            dts[s] = [0.01, 0.5, 0.01, 1.0, 0.8, 0.01, 0.3, 2.5, 0.01, 0.01, 0.7, 0.01][s]
            now[0] += dts[s]
            for p in range(n_ports):
                r = rng.random()
                if s == 0 or r < 0.45:
                    v = rng.normal(size=n)
                    events[s, p] = v
                    ports[p].push(v)
                elif r < 0.55:
                    short[s, p] = True
                    ports[p].push(rng.normal(size=n - 2))
            if s in (3, 8):
                w = rng.uniform(0, 1, size=(4 if s == 3 else 6))
                wevents[s, :w.size] = w
                wport.push(w)
            with redirect_stdout(io.StringIO()):
                outs[s] = mixer.read()
            weights_after[s] = mixer.weights
        # constructor contract: wrong number of initial weights -> zeros
        with redirect_stdout(io.StringIO()):
            m2 = CommandMixer([FakePort(), FakePort()], None, 3, 1.0, [1.0])
        bad_init = np.asarray(m2.weights)
    finally:
        _time.time = real_time
    return dict(mixer_events=events, mixer_wevents=wevents, mixer_short=short, mixer_dts=dts, mixer_out=outs,
                mixer_weights_after=weights_after, mixer_bad_init_weights=bad_init)


def lwr_chain():
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT_DIR)))
    from vfclik_b200.config import PACKAGE_CONFIG_DIR, chain_from_config, config_filename, load_config
    cfg = load_config(config_filename(PACKAGE_CONFIG_DIR + "/lwr/", "lwr", "right"))
    return chain_from_config(cfg), cfg


def gen_nullspace(rng):
    """Real restrict / nullspace / move_in_nullspace / check_limits on LWR Jacobians along a short
    joint trajectory (sign continuity is stateful, so the call order is part of the fixture)."""
    from oracle import batch
    chain, cfg = lwr_chain()
    steps = 10
    q = np.zeros((steps, 7))
    q[0] = cfg.initial_joint_pos
    for s in range(1, steps):
        q[s] = q[s - 1] + rng.normal(scale=0.05, size=7)
    _, _, J = batch.fk_jac(chain, q)
    glb = load_real_nullspace(7)
    P = np.asmatrix(np.eye(6))
    control = [0.7, -0.2, 0.1, 0.0]
    B = np.zeros((steps, 7, 7))
    qd = np.zeros((steps, 7))
    basis = np.zeros((steps, 7))
    limited = np.zeros((steps, 7))
    hit = np.zeros(steps, dtype=bool)
    limits = [[float(a), float(b)] for a, b in zip(chain.q_lo, chain.q_hi)]
    for s in range(steps):
        Jm = np.asmatrix(J[s])
        B[s] = np.asarray(glb["restrict"](P, Jm))
        qd[s] = glb["move_in_nullspace"](P, Jm, control)
        basis[s] = np.asarray(glb["lastvec"])[:, 0]
        with redirect_stdout(io.StringIO()):
            out = glb["check_limits"](list(q[s]), list(qd[s] * (25.0 if s % 3 == 2 else 1.0)), limits)
        limited[s] = out
        hit[s] = all(v == 0 for v in out)
    rank = int(glb["matrixrank"](np.asmatrix(J[0])))
    # random (non-LWR) full-rank 6x7 Jacobians, fresh state per sample: projector only
    Jr = rng.normal(size=(16, 6, 7))
    Br = np.stack([np.asarray(load_real_nullspace(7)["restrict"](P, np.asmatrix(Jr[k]))) for k in range(16)])
    return dict(ns_q=q, ns_J=J, ns_B=B, ns_control=np.asarray(control), ns_qdot=qd, ns_basis=basis,
                ns_limited=limited, ns_hit=hit, ns_rank=np.asarray(rank), ns_Jrand=Jr, ns_Brand=Br)


def gen_nullspace_wide(rng=None):
    """The same reference functions with ``nJoints = 10`` (the reference's iCub shape, ``scripts/bridge:344-345``: a 4-D
    nullspace) and 8 (2-D) along short joint trajectories of a torso + arm chain: projector, the ``k = N - 6`` basis vectors
    ``nullspace()`` returns with their sign-continuity state, and ``move_in_nullspace`` on four control floats.  For
    ``k > 1`` LAPACK's choice of basis inside the degenerate singular subspace is arbitrary, so what the record pins is the
    projector ``sum_i u_i u_i^T``, the span, ``k`` and the norm of the motion (own seed: the older keys stay as they are)."""
    from oracle import batch
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT_DIR)))
    from vfclik_b200 import workloads
    rng = np.random.default_rng(20261019)
    out = {}
    for n in (10, 8):
        chain = workloads.torso_arm_chain(n)
        steps = 8
        q = np.zeros((steps, n))
        q[0] = rng.uniform(0.4 * chain.q_lo, 0.4 * chain.q_hi)
        for s in range(1, steps):
            q[s] = q[s - 1] + rng.normal(scale=0.03, size=n)
        _, _, J = batch.fk_jac(chain, q)
        glb = load_real_nullspace(n)
        P = np.asmatrix(np.eye(6))
        control = [0.7, -0.2, 0.1, 0.4]
        B = np.zeros((steps, n, n))
        basis = np.zeros((steps, n - 6, n))
        qd = np.zeros((steps, n))
        for s in range(steps):
            Jm = np.asmatrix(J[s])
            B[s] = np.asarray(glb["restrict"](P, Jm))
            qd[s] = glb["move_in_nullspace"](P, Jm, control)
            basis[s] = np.asarray(glb["lastvec"])[:, :n - 6].T
        tag = "ns%d_" % n
        out.update({tag + "q": q, tag + "J": J, tag + "B": B, tag + "basis": basis, tag + "qdot": qd,
                    tag + "control": np.asarray(control)})
    return out


def load_reference_function(relpath: str, name: str, cls: str = None):
    """Source text of one function (or method of ``cls``) of a reference script, dedented, ready for ``exec``.

    ``scripts/vf`` and ``scripts/bridge`` contain python-2 ``print`` statements elsewhere in the file, so they cannot be
    parsed as a whole; the wanted ``def`` itself is python-3 clean and is cut out by indentation."""
    import textwrap
    lines = open(os.path.join(REF, relpath)).read().splitlines()
    start = 0
    if cls is not None:
        start = next(i for i, l in enumerate(lines) if l.startswith("class %s(" % cls) or l.startswith("class %s:" % cls)) + 1
    indent = "    " if cls is not None else ""
    head = indent + "def %s(" % name
    i0 = next(i for i in range(start, len(lines)) if lines[i].startswith(head))
    i1 = i0 + 1
    while i1 < len(lines) and (lines[i1].strip() == "" or lines[i1].startswith(indent + " ")):
        i1 += 1
    return textwrap.dedent("\n".join(lines[i0:i1])) + "\n"


class CapturePort:
    """Output port stand-in: ``prepare()`` returns a bottle that records ``addDouble`` / ``addString``."""

    def __init__(self):
        self.items = []

    def prepare(self):
        return self

    def clear(self):
        self.items = []

    def addDouble(self, v):
        self.items.append(float(v))

    def addString(self, v):
        self.items.append(str(v))

    def toString(self):
        return " ".join(str(v) for v in self.items)

    def write(self, *a):
        pass


def gen_bridge(rng):
    """The three back-ends' ``set_vel`` (scripts/bridge:182-210, 288-312, 507-530), ``ICUB_Bridge.set_torso_cjoints``
    (:470-506), ``joint_p_controller.check_limits`` (:79-89) and ``vf.get_weight_matrix`` (:164-179), executed as they
    stand in the reference on random inputs."""
    N = 7
    cases = 64
    qdot = rng.normal(scale=0.7, size=(cases, N))
    qdot[::5] *= 0.05                                              # some below every limit
    q = rng.normal(size=(cases, N))
    qc = q + rng.normal(scale=0.02, size=(cases, N))
    cfg = types.SimpleNamespace(max_vel=0.6, max_vel_shoulder_pos=0.15, max_vel_shoulder_neg=-0.1)
    out = {"br_qdot": qdot, "br_q": q, "br_qcmded": qc,
           "br_cfg": np.array([cfg.max_vel, cfg.max_vel_shoulder_pos, cfg.max_vel_shoulder_neg])}
    sink = io.StringIO()
    # LWR: both command forms
    for direct in (0, 1):
        glb = {"direct_control": bool(direct), "max_vel": cfg.max_vel}
        exec(load_reference_function("scripts/bridge", "set_vel", cls="LWR_Bridge"), glb)
        rows = []
        for k in range(cases):
            me = types.SimpleNamespace(nJoints=N, last_q=list(q[k]), last_qcmded=list(qc[k]), qcmd_port=CapturePort())
            with redirect_stdout(sink):
                glb["set_vel"](me, list(qdot[k]))
            rows.append(me.qcmd_port.items)
        out["br_lwr_cmd_direct%d" % direct] = np.asarray(rows)
    # Powercube and iCub, simulation branch
    for cls, key in (("Powercube_Bridge", "br_powercube_cmd"), ("ICUB_Bridge", "br_icub_cmd")):
        glb = {"config": cfg}
        exec(load_reference_function("scripts/bridge", "set_vel", cls=cls), glb)
        rows = []
        for k in range(cases):
            me = types.SimpleNamespace(sim=True, qcmd_port=CapturePort())
            with redirect_stdout(sink):
                glb["set_vel"](me, list(qdot[k]))
            rows.append(me.qcmd_port.items)
        out[key] = np.asarray(rows)
    # iCub torso (de)activation -> joint weights message
    glb = {}
    exec(load_reference_function("scripts/bridge", "set_torso_cjoints", cls="ICUB_Bridge"), glb)
    msgs = []
    for cj in ([1, 0, 1], [0, 0, 0], [1, 1, 1]):
        me = types.SimpleNamespace(sim=True, icub_torso_num_joints=3, icub_torso_cjoints=[True] * 3,
                                   control_weights_port=CapturePort())
        with redirect_stdout(sink):
            glb["set_torso_cjoints"](me, cj)
        assert me.control_weights_port.items[0] == "j"
        msgs.append(me.control_weights_port.items[1:])
    out["br_icub_weights"] = np.asarray(msgs)
    # joint_p_controller.check_limits with posture-dependent limits
    lim_of = lambda cur: [[-0.5 - abs(cur[0]), 0.4 + abs(cur[1])]] * N
    glb = {"config": types.SimpleNamespace(updateJntLimits=lim_of)}
    exec(load_reference_function("scripts/joint_p_controller", "check_limits"), glb)
    ref = rng.normal(scale=1.0, size=(cases, N))
    with redirect_stdout(sink):
        out["jp_ref_clamped"] = np.asarray([glb["check_limits"](list(ref[k]), list(q[k])) for k in range(cases)])
    out["jp_ref"] = ref
    # vf.get_weight_matrix: good and wrong sizes
    glb = {"zeros": np.zeros, "dprint": lambda *a: None}
    exec(load_reference_function("scripts/vf", "get_weight_matrix"), glb)

    class WB(FakeBottle):
        def get(self, i):
            v = FakeValue(self.vals[i])
            v.asString = lambda: str(self.vals[i])
            return v
    wt = rng.uniform(0.1, 2.0, size=6)
    wj = rng.uniform(0.1, 2.0, size=N)
    out["vf_wt"], out["vf_wj"] = wt, wj
    out["vf_wt_matrix"] = glb["get_weight_matrix"](WB(["t"] + list(wt)), 6)
    out["vf_wj_matrix"] = glb["get_weight_matrix"](WB(["j"] + list(wj)), N)
    assert glb["get_weight_matrix"](WB(["j"] + list(wj[:5])), N) is None          # wrong size: ignored
    return out


def load_reference_loop_body(relpath: str, head: str, end_marker: str) -> str:
    """Body of a top-level ``while`` loop of a reference script (from the line after ``head`` to the line before
    ``end_marker``), dedented -- the module programs keep their logic inline in that loop."""
    import textwrap
    lines = open(os.path.join(REF, relpath)).read().splitlines()
    i0 = next(i for i, l in enumerate(lines) if l.startswith(head)) + 1
    i1 = next(i for i in range(i0, len(lines)) if lines[i].startswith(end_marker))
    body = [l for l in lines[i0:i1] if not l.lstrip().startswith("#")]      # column-0 comments would defeat the dedent
    return textwrap.dedent("\n".join(body)) + "\n"


class Py2IntKeyDict(dict):
    """The reference runs under python 2, whose dicts iterate small non-negative int keys in ascending order;
    python 3 keeps insertion order.  Used for the feeder's ``objects`` so the emission order is the reference's."""

    def __iter__(self):
        return iter(sorted(dict.keys(self)))


FEEDER_SCRIPT = [
    ["set", "ObstacleP", 0, [1, 0, 0, 0.0, 0, 1, 0, -0.4, 0, 0, 1, 0.4, 0, 0, 0, 1, 0.05, 20]],          # old/README.old:75 (before any goal)
    ["set", "goal", [0, 1, 0, 0, -1, 0, 0, 0.3, 0, 0, 1, 1.1, 0, 0, 0, 1, 0.1]],                         # old/README.old:69
    ["set", "ObstacleH", 1, [1, 0, 0, 0, 0, 1, 0, -0.4, 0, 0, 1, 0.3, 0, 0, 0, 1, 0, 0, 1, 0.001, 5]],   # old/README.old:78
    ["set", "goalAndNormal", [1, 0, 0, 0.4, 0, -1, 0, -0.4, 0, 0, -1, 0.4, 0, 0, 0, 1, 0, -1, 0, 0.1, 0.15, 0.15]],   # :72-73
    ["set", "goal", [1, 0, 0, 0.5, 0, 1, 0, 0.1, 0, 0, 1, 0.9, 0, 0, 0, 1]],                             # 16 values: default slowdown
    ["set", "goal", [1.0, 2.0, 3.0]],                                                                   # wrong length
    ["set", "ObstacleP", 2, [1, 0, 0, 0.2, 0, 1, 0, 0.1, 0, 0, 1, 0.7, 0, 0, 0, 1, 0.08, 7]],
    ["remove", 0],
    ["remove", 5],                                                                                       # does not exist
    ["fly", "goal", [1.0]],                                                                              # unknown action
    ["set", "goalAndNormal", [1, 0, 0, 0.4, 0, -1, 0, -0.4, 0, 0, -1, 0.4, 0, 0, 0, 1, 0, 0, 2, 0.1, 0.15]],   # 21 values
]


def gen_feeder(rng=None):
    """``scripts/object_feeder``'s loop body (:93-359) executed message by message on the scripted sequence above, with
    this repo's in-process bottles / ports standing in for YARP and ``vfl.vfl.length`` taken as the Euclidean norm.
    Returns the messages it wrote to ``/param`` and ``/objectOut`` as one JSON string per input message."""
    import json
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from vfclik_b200 import ports as yarp
    body = load_reference_loop_body("scripts/object_feeder", "while not stop:", "paramPort.close()")
    code = compile_iteration(body, ['init_pose'], "<reference scripts/object_feeder loop>")

    class EndOfScript(Exception):
        pass

    class FeederIn(ScriptIn):
        def read(self, wait=False):
            if not self.q:
                raise EndOfScript()           # the loop came back for another message (`continue`): iteration over
            return self.q.pop(0)

    param, objout, objf, obj = ScriptOut(), ScriptOut(), ScriptOut(), FeederIn()
    glb = dict(yarp=types.SimpleNamespace(Bottle=yarp.Bottle, Time_delay=lambda t: None), yarp_ctrl=types.SimpleNamespace(update=lambda: None),
               objectPort=obj, object_f_port=objf, paramPort=param, objectOutPort=objout, objects=Py2IntKeyDict(),
               init_pose=False, config=types.SimpleNamespace(), dprint=lambda *a: None, array=np.array,
               length=lambda v: float(np.linalg.norm(v)), recur=None, stop=False)
    exec(code, glb)
    rows = []
    for msg in FEEDER_SCRIPT:
        obj.q.append(yarp.Bottle.from_list(msg))
        param.sent, objout.sent = [], []
        with redirect_stdout(io.StringIO()):
            try:
                glb["_iteration"]()
            except EndOfScript:
                pass
        rows.append(json.dumps({"param": param.sent, "objectOut": objout.sent}))
    return {"feeder_script": np.array([json.dumps(m) for m in FEEDER_SCRIPT]), "feeder_out": np.array(rows)}


def gen_jp(rng):
    """``scripts/joint_p_controller``'s loop body (:96-146) with the reference's own ``check_limits`` (:79-89), one
    iteration per scripted step: an optional new reference on ``/ref``, joint positions on ``/in``; records ``/out`` and
    ``/at_goal``.  Two limit hooks: the LWR's static limits and posture-dependent ones (the clamped reference persists
    between cycles -- the loop overwrites ``ref`` -- which only shows with the latter).  ``map`` is given its python-2
    meaning (a list), which the loop relies on for its side effects."""
    import builtins
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from vfclik_b200 import ports as yarp
    N, steps, kp, delta = 7, 18, 1.5, 0.087
    deg = np.pi / 180.0
    static = [[-170 * deg, 170 * deg], [-120 * deg, 120 * deg]] * 3 + [[-170 * deg, 170 * deg]]
    hooks = {"static": lambda cur: static, "dynamic": lambda cur: [[-0.5 - abs(cur[0]), 0.4 + abs(cur[1])]] * N}
    q = rng.uniform(-1.0, 1.0, size=(steps, N))
    q[12:] = q[11] + rng.normal(scale=0.01, size=(steps - 12, N))          # nearly at rest near the last reference
    refs = {0: rng.uniform(-3.5, 3.5, size=N), 5: rng.uniform(-0.3, 0.3, size=N), 11: q[11] + 0.05}
    body = load_reference_loop_body("scripts/joint_p_controller", "while not stop:", "inPort.close()")
    code = compile_iteration(body, ['ref', 'waitRef', 'refbottle'], "<reference scripts/joint_p_controller loop>")

    out = {"jp_q": q, "jp_ref_steps": np.array(sorted(refs)), "jp_ref_msgs": np.array([refs[k] for k in sorted(refs)]),
           "jp_static_limits": np.array(static)}
    for name, hook in hooks.items():
        refp, inp, outp, goalp = ScriptIn(), ScriptIn(), ScriptOut(), ScriptOut()
        cfg = types.SimpleNamespace(nJoints=N, initial_joint_pos=[0.0, -1.2, 0.7, 1.4, 0.35, -1.4, 0.0], updateJntLimits=hook)
        glb = dict(yarp=types.SimpleNamespace(Bottle=yarp.Bottle, Value=yarp.Value, Time=types.SimpleNamespace(delay=lambda t: None)),
                   yarp_ctrl=types.SimpleNamespace(update=lambda: None), refPort=refp, inPort=inp, outPort=outp, atGoalPort=goalp,
                   config=cfg, kp=kp, delta=delta, array=np.array, stop=False,
                   map=lambda f, *a: list(builtins.map(f, *a)), ref=cfg.initial_joint_pos, waitRef=False, refbottle=yarp.Bottle())
        with redirect_stdout(io.StringIO()):
            exec(load_reference_function("scripts/joint_p_controller", "check_limits"), glb)
            exec(code, glb)
            for k in range(steps):
                if k in refs:
                    refp.q.append(yarp.Bottle.from_list([float(v) for v in refs[k]]))
                inp.q.append(yarp.Bottle.from_list([float(v) for v in q[k]]))
                glb["_iteration"]()
        assert len(outp.sent) == steps and len(goalp.sent) == steps
        out["jp_out_" + name] = np.asarray(outp.sent)
        out["jp_at_goal_" + name] = np.asarray([g[0] for g in goalp.sent])
    out["jp_kp_delta"] = np.array([kp, delta])
    return out


def gen_vf(rng):
    """``scripts/vf``'s loop body (:193-505) executed cycle by cycle.  PyKDL, arcospyu.Lafik and vfl are un-vendored, so the
    oracle's stand-ins (``oracle/refshape.py``) take their names: what this pins is everything the reference's own loop does
    around them -- field bookkeeping and composition order, weight and tool handling, the speed-scale rule, the pose /
    tracking-error / vector_out / goal_out ports and their cadence, the tool twist shift and the call into the IK.
    The plant is closed around it (q += rate * qdot) so the tracking-error block sees a moving arm."""
    import builtins
    import json
    from math import acos, sqrt
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from vfclik_b200 import ports as yarp
    from . import refshape
    from .batch import Params
    cfgfile = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vfclik_b200", "config_data", "lwr",
                           "config-lwr-right.py")
    from vfclik_b200.config import chain_from_config, load_config
    cfg = load_config(cfgfile)
    chain = chain_from_config(cfg)
    prm = Params(ik_lambda=cfg.ik_lambda, speed_scale=cfg.speedScale)
    body = load_reference_loop_body("scripts/vf", "while not stop:", "qdotOutPort.close()")
    state = ["speedScale", "start_attractor", "first_arm_data", "first_arm_frame", "oldtoolFrame", "totalVF", "totalSF", "vftemp",
             "vfparams", "counter_test", "reporting_port_counter", "frame_list", "cmd_buffer", "vectorFields", "first_cycle"]
    code = compile_iteration(body, [s for s in state if s != "speedScale"], "<reference scripts/vf loop>")

    ins = {k: ScriptIn() for k in ("maxvel_port", "paramPort", "weightPort", "qInPort", "toolPort", "pose_in_port")}
    outs = {k: ScriptOut() for k in ("qdotOutPort", "posePort", "pose_no_tool_Port", "tracking_error_port", "vector_port", "goal_port")}

    def sendListPort(port, l):
        b = port.prepare()
        for v in l:
            b.addDouble(v)
        port.write()

    def readListPort(port, blocking=False):
        b = port.read(blocking)
        return None if b is None else [b.get(i).asDouble() for i in range(b.size())]

    vfDB = refshape.vfl_library()
    refshape._PointAttractor.rot_slowdown = prm.rot_slowdown
    vft = vfDB[0]()
    vft.setParams([])
    glb = dict(yarp=types.SimpleNamespace(Value=yarp.Value, Value_makeString=yarp.Value_makeString, Bottle=yarp.Bottle,
                                          Time=types.SimpleNamespace(delay=lambda t: None)),
               yarp_ctrl=types.SimpleNamespace(update=lambda: None), config=cfg, numJntsArm=cfg.nJoints, config_max_vel=0.41,
               speedScale=cfg.speedScale, vfDB=vfDB, VectorField=refshape.VectorField, ScalarField=refshape.ScalarField,
               lafik=refshape.Lafik(chain, prm), PyKDL=types.SimpleNamespace(diff=refshape.kdl_diff, Twist=refshape.Twist,
                                                                           Vector=lambda x, y, z: np.array([x, y, z])),
               listToKdlFrame=refshape.listToKdlFrame, kdlFrameToList=refshape.kdlFrameToList, sendListPort=sendListPort,
               readListPort=readListPort, sqrt=sqrt, acos=acos, dot=np.dot, array=np.array, dprint=lambda *a: None,
               map=lambda f, *a: list(builtins.map(f, *a)), stop=False,
               vectorFields=Py2IntKeyDict(), vfparams=[], vftemp=vft, totalVF=refshape.VectorField(vft.getVector),
               totalSF=refshape.ScalarField(vft.getScalar), oldtoolFrame=[1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1],
               frame_list=[], cmd_buffer=[], cmd_buffer_size=4, frame_list_size=5, check_delay=4, first_cycle=True,
               start_attractor=False, first_arm_data=True, first_arm_frame=None, reporting_port_counter=0, counter_test=0)
    glb.update(ins)
    glb.update(outs)
    with redirect_stdout(io.StringIO()):
        exec(load_reference_function("scripts/vf", "get_weight_matrix"), dict(zeros=np.zeros, dprint=lambda *a: None), glb)
    glb["zeros"] = np.zeros
    exec(load_reference_function("scripts/vf", "get_weight_matrix"), glb)
    exec(code, glb)

    goal = [float(v) for v in cfg.initial_vf_pose[2]]
    script = {
        1: {"paramPort": ["add", 1, 1.0, 1, goal]},
        2: {"paramPort": ["add", 5, -10.0, 2, [0.55, 0.10, 1.05, 0.06, 0.001, 20.0]]},
        3: {"paramPort": ["add", 6, -10.0, 2, [0.30, 0.25, 1.20, 0.05, 0.001, 20.0]]},
        6: {"weightPort": ["t", 1.0, 1.0, 1.0, 0.5, 0.5, 0.5]},
        8: {"toolPort": [1.0, 0.0, 0.0, 0.02, 0.0, 1.0, 0.0, -0.01, 0.0, 0.0, 1.0, 0.12, 0.0, 0.0, 0.0, 1.0]},
        10: {"maxvel_port": [0.3]},
        12: {"maxvel_port": [0.9]},                                        # out of range: ignored
        14: {"paramPort": ["add", 9, -10.0, 7, [0.0, 0.0, 0.0]]},           # unknown field type: ignored
        16: {"weightPort": ["j", 1.0, 1.0, 0.7, 1.0, 1.0, 0.4]},            # wrong length: ignored
        18: {"paramPort": ["remove", 6]},
        20: {"weightPort": ["j", 1.0, 1.0, 0.7, 1.0, 1.0, 0.4, 1.0]},
        24: {"pose_in_port": [1.0, 0.0, 0.0, 0.5, 0.0, 1.0, 0.0, 0.2, 0.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0, 1.0]},
        30: {"paramPort": ["add", 2, 30.0, 5, [goal[3], goal[7], goal[11], 0.0, 0.0, -1.0, 0.3, 10.0, 0.15, 2.0]]},
    }
    steps = 48
    q = [float(v) for v in cfg.initial_joint_pos]
    rows, qs = [], []
    for k in range(steps):
        for port, msg in script.get(k, {}).items():
            ins[port].q.append(yarp.Bottle.from_list(msg))
        ins["qInPort"].q.append(yarp.Bottle.from_list(q))
        for o in outs.values():
            o.sent = []
        with redirect_stdout(io.StringIO()):
            glb["_iteration"]()
        rows.append(json.dumps({name: o.sent for name, o in outs.items()}))
        qs.append(list(q))
        if outs["qdotOutPort"].sent:
            qd = outs["qdotOutPort"].sent[-1]
            q = [q[i] + float(cfg.rate) * 25.0 * qd[i] for i in range(len(q))]      # exaggerated step: visible motion in 48 cycles
    return {"vf_script": np.array([json.dumps({str(k): v for k, v in script.items()})]), "vf_q": np.asarray(qs),
            "vf_out": np.array(rows)}


def gen_bridge_loop(rng):
    """``scripts/bridge``'s main loop body (:600-634) with the reference's own ``LWR_Bridge.read_pos`` / ``set_vel``
    (:163-210), ``ut_writebottle`` / ``ut_bottle2list`` (:540-549) and the real ``CommandMixer``: one iteration per scripted
    step.  Inputs: plant positions, /cmded feedback, the three controller command ports, mixer weights (incl. all-zero =
    direct control) and max_vel messages (one out of range); records ``/encoders``, the robot command and
    ``/current_weights``."""
    import json
    import time as _time
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from vfclik_b200 import ports as yarp
    N, steps = 7, 24
    CommandMixer = load_real_command_mixer()
    import textwrap
    lines = open(os.path.join(REF, "scripts", "bridge")).read().splitlines()
    i0 = next(i for i, l in enumerate(lines) if l.startswith("    while (not stop):")) + 1
    i1 = next(i for i in range(i0, len(lines)) if lines[i].startswith("    encoders_port.close()"))
    body = textwrap.dedent("\n".join(l for l in lines[i0:i1] if not l.lstrip().startswith("#")))
    code = compile_iteration(body, ['direct_control', 'max_vel'], "<reference scripts/bridge loop>")

    cmd_ports = [ScriptIn() for _ in range(6)]
    weight_port, maxvel_port, qin, qcmded = ScriptIn(), ScriptIn(), ScriptIn(), ScriptIn()
    encoders, qcmd, cur_w = ScriptOut(), ScriptOut(), ScriptOut()
    cfg = types.SimpleNamespace(nJoints=N, max_vel=0.5, torso_joints=[])
    glb = dict(config=cfg, config_max_vel=cfg.max_vel, max_vel=cfg.max_vel, direct_control=False, rate=0.0, time=_time, stop=False,
               yarp_ctrl=types.SimpleNamespace(update=lambda: None), encoders_port=encoders, current_weight_port=cur_w,
               maxvel_port=maxvel_port)
    for name in ("ut_bottle2list", "ut_writebottle"):
        exec(load_reference_function("scripts/bridge", name), glb)
    ns = {}
    for name in ("read_pos", "set_vel"):
        exec(load_reference_function("scripts/bridge", name, cls="LWR_Bridge"), glb, ns)
    bridge = types.SimpleNamespace(nJoints=N, last_q=N * [0.0], last_qcmded=[], torso_joints=[], qin_port=qin, qcmded_port=qcmded,
                                   qcmd_port=qcmd)
    bridge.read_pos = lambda: ns["read_pos"](bridge)
    bridge.set_vel = lambda qdot: ns["set_vel"](bridge, qdot)
    glb["bridge"] = bridge
    glb["mixer"] = CommandMixer(cmd_ports, weight_port, N, 2.0, [1.0, 1.0, 0.0, 0.0, 0.0, 0.0])
    exec(code, glb)

    q = rng.uniform(-1, 1, size=N)
    script, rows = [], []
    for k in range(steps):
        q = q + rng.normal(scale=0.01, size=N)
        ev = {"q": q.tolist(), "cmd": {}}
        if k > 0:
            ev["cmded"] = (q + rng.normal(scale=0.003, size=N)).tolist()
        for p, scale in ((0, 0.6), (1, 0.05), (2, 0.3)):
            if k % (p + 1) == 0:
                ev["cmd"][str(p)] = rng.normal(scale=scale, size=N).tolist()
        if k == 6:
            ev["weights"] = [0.0, 0.0, 1.0]
        if k == 12:
            ev["weights"] = [0.0, 0.0, 0.0, 0.0, 0.0, 0.0]
        if k == 17:
            ev["weights"] = [1.0, 0.5, 0.0, 0.0, 0.0, 0.0]
        if k == 9:
            ev["max_vel"] = 0.2
        if k == 15:
            ev["max_vel"] = 3.0                                  # out of range: ignored
        script.append(ev)
        qin.q.append(yarp.Bottle.from_list(ev["q"]))
        if "cmded" in ev:
            qcmded.q.append(yarp.Bottle.from_list(ev["cmded"]))
        for p, v in ev["cmd"].items():
            cmd_ports[int(p)].q.append(yarp.Bottle.from_list(v))
        if "weights" in ev:
            weight_port.q.append(yarp.Bottle.from_list(ev["weights"]))
        if "max_vel" in ev:
            maxvel_port.q.append(yarp.Bottle.from_list([ev["max_vel"]]))
        encoders.sent, qcmd.sent, cur_w.sent = [], [], []
        with redirect_stdout(io.StringIO()):
            glb["_iteration"]()
        rows.append(json.dumps({"encoders": encoders.sent, "qcmd": qcmd.sent, "current_weights": cur_w.sent}))
    return {"bl_script": np.array([json.dumps(e) for e in script]), "bl_out": np.array(rows),
            "bl_cfg": np.array([cfg.max_vel])}


def gen_bridge_readpos(rng=None):
    """``LWR_Bridge.read_pos`` (``scripts/bridge:163-180``) executed without any ``/cmded`` feedback, as in ``bridge -s`` where
    that port is never connected (``:136-139``): ``last_qcmded`` is set ONCE, to the first measured q (the same list object),
    and never follows q afterwards, so the non-direct command ``-q_cmded + q + qdot`` carries ``q - q_first``.  Pinned here
    so that the deviation this repo's bridge makes in simulation (it refreshes q_cmded every cycle) is a
    documented choice and its feedback-less real-robot path is checked against the reference (own seed)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from vfclik_b200 import ports as yarp
    rng = np.random.default_rng(20261020)
    N, steps = 7, 5
    glb = {}
    exec(load_reference_function("scripts/bridge", "ut_bottle2list"), glb)
    ns = {}
    exec(load_reference_function("scripts/bridge", "read_pos", cls="LWR_Bridge"), glb, ns)
    qin, qcmded = ScriptIn(), ScriptIn()
    bridge = types.SimpleNamespace(nJoints=N, last_q=N * [0.0], last_qcmded=[], torso_joints=[], qin_port=qin, qcmded_port=qcmded)
    q = rng.uniform(-1, 1, size=(steps, N))
    seen_q, seen_c = [], []
    for k in range(steps):
        qin.q.append(yarp.Bottle.from_list([float(v) for v in q[k]]))
        seen_q.append(list(ns["read_pos"](bridge)))
        seen_c.append(list(bridge.last_qcmded))
    return {"rp_q": q, "rp_last_q": np.asarray(seen_q), "rp_last_qcmded": np.asarray(seen_c)}


def gen_dmonitor(rng):
    """``scripts/monitor_distance``'s loop body (:107-221) with its own ``orientLength`` (:76-84); PyKDL ``Frame`` / ``diff``
    and ``vfl.length`` replaced by the oracle's stand-ins.  One iteration per scripted step: objects, the tool pose and the
    vf module's tracking errors in; ``/distOut`` and the ``/tracking_state`` change messages out (20-sample majority,
    including the reference sending the *xyz* state under the ``rot`` tag)."""
    import builtins
    import json
    from math import pi, sqrt
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from vfclik_b200 import ports as yarp
    from . import refshape
    body = load_reference_loop_body("scripts/monitor_distance", "while not stop:", "track_error_in_port.close()")
    state = ["track_error_xyz", "track_error_rot", "tracking_buffer", "last_tracking_xyz_state", "last_tracking_rot_state",
             "tracking_xyz_state", "tracking_rot_state", "objects"]
    code = compile_iteration(body, state, "<reference scripts/monitor_distance loop>")

    objs, terr, pose = ScriptIn(), ScriptIn(), ScriptIn()
    dist, tstate = ScriptOut(), ScriptOut()
    glb = dict(yarp=types.SimpleNamespace(Value=yarp.Value, Time=types.SimpleNamespace(delay=lambda t: None), Time_delay=lambda t: None),
               yarp_ctrl=types.SimpleNamespace(update=lambda: None), objectsInPort=objs, track_error_in_port=terr, currentPosIn=pose,
               distOutPort=dist, tracking_state_port=tstate, objects=Py2IntKeyDict(), array=np.array,
               length=lambda v: float(np.linalg.norm(v)), Frame=refshape.KdlFrame, diff=refshape.kdl_diff, sqrt=sqrt, pi=pi,
               map=lambda f, *a: list(builtins.map(f, *a)), zip=lambda *a: list(builtins.zip(*a)), stop=False,
               track_error_xyz=0.0, track_error_rot=0.0, distanceXYZ_th=0.02, track_error_xyz_th=0.1, distanceOrient_th=1.0,
               track_error_rot_th=0.1, tracking_buffer=[], tracking_buffer_size=20, last_tracking_xyz_state="on goal",
               last_tracking_rot_state="on goal", tracking_xyz_state="on goal", tracking_rot_state="on goal")
    exec(load_reference_function("scripts/monitor_distance", "orientLength"), glb)
    exec(code, glb)

    def frame16(R, p):
        T = np.eye(4); T[:3, :3] = R; T[:3, 3] = p
        return T.reshape(16).tolist()

    def rot(axis, ang):
        a = np.asarray(axis, dtype=float); a /= np.linalg.norm(a)
        K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)

    Rg, pg = rot([0.2, 1.0, 0.3], 0.7), np.array([0.5, 0.1, 0.9])
    steps = 70
    script, rows = [], []
    for k in range(steps):
        ev = {}
        if k == 0:
            ev["objects"] = [["add", 0, frame16(Rg, pg)], ["add", 3, frame16(np.eye(3), [0.2, -0.3, 0.6])]]
        if k == 40:
            ev["objects"] = [["remove", 3]]
        s_ = max(0.0, 1.0 - k / 50.0)                                   # approach: the distance shrinks to zero at step 50
        R = Rg @ rot([1.0, 0.2, -0.4], 0.9 * s_)
        p = pg + np.array([0.3, -0.2, 0.25]) * s_
        ev["pose"] = frame16(R, p)
        # tracking errors: fine, then a stretch where the arm does not follow (both), then fine again
        bad = 12 <= k < 36
        ev["track_error"] = [0.5 if bad else 0.02, 0.4 if bad else 0.03, 0.1, 0.1, 0.1, 0.1, 0.0, 1]
        script.append(ev)
        for m in ev.get("objects", []):
            objs.q.append(yarp.Bottle.from_list(m))
        terr.q.append(yarp.Bottle.from_list(ev["track_error"]))
        pose.q.append(yarp.Bottle.from_list(ev["pose"]))
        n_iter = max(1, len(ev.get("objects", [])))                     # the loop reads one /objectsIn message per iteration
        dist.sent, tstate.sent = [], []
        with redirect_stdout(io.StringIO()):
            for _ in range(n_iter):
                glb["_iteration"]()
        rows.append(json.dumps({"distOut": dist.sent, "tracking_state": tstate.sent}))
    return {"dm_script": np.array([json.dumps(e) for e in script]), "dm_out": np.array(rows)}


def gen_nullspace_loop(rng):
    """``scripts/nullspace``'s main loop body (:159-187) with the reference's own functions (:67-131) and the oracle's Lafik
    stand-in as ``rob``: joint positions and four-float control bottles in, ``/qdotout`` (gain applied) out."""
    import json
    import textwrap
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from vfclik_b200 import ports as yarp
    from . import refshape
    from .batch import Params
    chain, _ = lwr_chain()
    glb = load_real_nullspace(chain.n_joints)
    lines = open(os.path.join(REF, "scripts", "nullspace")).read().splitlines()
    i0 = next(i for i, l in enumerate(lines) if l.startswith("    while not stop:")) + 1
    i1 = next(i for i in range(i0, len(lines)) if lines[i].startswith("    qin_port.close()"))
    body = textwrap.dedent("\n".join(l for l in lines[i0:i1] if not l.lstrip().startswith("#")))
    code = compile_iteration(body, ['control'], "<reference scripts/nullspace loop>")

    qin, ctl, out = ScriptIn(), ScriptIn(), ScriptOut()
    glb.update(yarp=types.SimpleNamespace(Time_delay=lambda t: None), yarp_ctrl=types.SimpleNamespace(update=lambda: None),
               qin_port=qin, control_port=ctl, qdotout_port=out, rob=refshape.Lafik(chain, Params()), P=np.asmatrix(np.eye(6)),
               control=[0] * 4, gain=0.5, stop=False)
    exec(code, glb)
    steps = 16
    q = rng.uniform(0.5 * chain.q_lo, 0.5 * chain.q_hi)
    script, rows = [], []
    for k in range(steps):
        q = q + rng.normal(scale=0.02, size=chain.n_joints)
        ev = {"q": q.tolist()}
        if k == 2:
            ev["control"] = [0.4, 0.0, 0.0, 0.0]
        if k == 8:
            ev["control"] = [-0.8, 0.1, 0.0, 0.0]
        script.append(ev)
        qin.q.append(yarp.Bottle.from_list(ev["q"]))
        if "control" in ev:
            ctl.q.append(yarp.Bottle.from_list(ev["control"]))
        out.sent = []
        with redirect_stdout(io.StringIO()):
            glb["_iteration"]()
        rows.append(out.sent[0])
    return {"nl_script": np.array([json.dumps(e) for e in script]), "nl_qdotout": np.asarray(rows)}


HANDLER_SINKS = ["/ofeeder/object", "/robot/stiffness", "/vectorField/tool", "/bridge/weight", "/bridge/weights",
                 "/vectorField/weight", "/jpctrl/ref", "/joint_sim/qin", "/bridge/torso_cjoints:i"]


def handler_calls(mod, prefix, sinks, drain):
    """The scripted client calls (shared by the generator and the test): every sender of the four handler classes."""
    rec = []

    def note(label):
        rec.append([label, drain()])
    arm = mod.HandleArmNew(namespace="/0", module_name="/handle_arm", arm_namespace="/0", robot="/lwr", arm="/right", sim=True)
    frame = [0.0, 1.0, 0.0, 0.5, -1.0, 0.0, 0.0, 0.1, 0.0, 0.0, 1.0, 0.9, 0.0, 0.0, 0.0, 1.0]
    arm.set_sim_arm_q([0.1, -0.2, 0.3, 0.4, -0.5, 0.6, 0.7]); note("new.set_sim_arm_q")
    arm.go_cart(frame); note("new.go_cart")
    arm.go_joint([0.0, -1.2, 0.7, 1.4, 0.35, -1.4, 0.0]); note("new.go_joint")
    arm.set_controller_mixer(cart=True, joint=True, null=False); note("new.set_controller_mixer")
    arm.set_cartesian_control(); note("new.set_cartesian_control")
    arm.set_joint_control(); note("new.set_joint_control")
    arm.set_wik_joint_weights([1.0, 1.0, 0.5, 1.0, 1.0, 0.25, 1.0]); note("new.set_wik_joint_weights")
    arm.set_wik_cart_weights([1.0, 1.0, 1.0, 0.5, 0.5, 0.5]); note("new.set_wik_cart_weights")
    arm.set_tool([1.0, 0.0, 0.0, 0.02, 0.0, 1.0, 0.0, -0.01, 0.0, 0.0, 1.0, 0.12, 0.0, 0.0, 0.0, 1.0]); note("new.set_tool")
    old = mod.HandleArm(prefix, namespace="", handlername="/HandlerArm")
    old.set_stiffness([200.0] * 7); note("arm.set_stiffness")
    old.sendFrame(); note("arm.sendFrame")
    old.gotoPos([0.4, -0.1, 1.0]); note("arm.gotoPos")
    old.setOrient([0.0, -1.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0]); note("arm.setOrient")
    old.gotoPose([0.3, 0.2, 0.8], [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0]); note("arm.gotoPose")
    br = mod.HandleBridge(prefix, handlername="HandlerArmBridge", torso=True)
    br.joint_controller(); note("bridge.joint_controller")
    br.cartesian_controller(); note("bridge.cartesian_controller")
    br.torso_joints([1, 0, 1]); note("bridge.torso_joints")
    br.set_weights("task", [1.0, 1.0, 1.0, 0.2, 0.2, 0.2]); note("bridge.set_weights task")
    br.set_VFW("joint", [1.0] * 6 + [0.5]); note("bridge.set_VFW joint")
    jc = mod.HandleJController(prefix, handlername="HandlerArmJoint")
    res = jc.set_ref_js([0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7]); note("jctrl.set_ref_js")
    rec.append(["jctrl.set_ref_js result", [bool(res[0]), [float(v) for v in res[1]]]])
    return rec


def gen_handlers(rng=None):
    """``src/handlers.py`` (the client API) executed with this repo's in-process ports as ``yarp``: every sender of
    ``HandleArmNew``, ``HandleArm``, ``HandleBridge`` and ``HandleJController`` is called once and the messages arriving at the
    arm-side port names it connects to are recorded.  The file's python-2 ``print`` statements are rewritten in memory
    (``print x`` -> ``print(x)``); nothing else is touched and nothing is stored."""
    import json
    import re
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from vfclik_b200 import ports as yarp
    from . import refshape
    src = open(os.path.join(REF, "src", "handlers.py")).read()
    src = re.sub(r"^(\s*)print (.+)$", r"\1print(\2)", src, flags=re.M)
    for line in ("import yarp\n", "from arcospyu.yarp_tools.yarp_comm_helpers import yarp_connect_blocking, \\\n    new_port\n",
                 "from arcospyu.kdl_helpers.kdl_helpers import frame_to_list\n"):
        assert line in src, line
        src = src.replace(line, "")
    yarp.Network.reset()
    prefix = "/0/lwr/right"
    sinks = {}
    for name in HANDLER_SINKS:
        p = yarp.BufferedPortBottle(); p.open(prefix + name); p.setStrict(True)
        sinks[name] = p
    for name in ("/vectorField/pose", "/dmonitor/distOut", "/bridge/encoders"):       # sources the handlers subscribe to
        p = yarp.BufferedPortBottle(); p.open(prefix + name)
        sinks["src:" + name] = p

    def new_port(name, direction, remote, timeout=None):
        p = yarp.BufferedPortBottle(); p.open(name)
        if direction == "out":
            yarp.Network.connect(name, remote)
        else:
            yarp.Network.connect(remote, name)
        return p

    def drain():
        got = {}
        for name in HANDLER_SINKS:
            while True:
                b = sinks[name].read(False)
                if b is None:
                    break
                got.setdefault(name, []).append(b.to_list())
        return got
    mod = types.ModuleType("ref_handlers")
    mod.__dict__.update(yarp=yarp, yarp_connect_blocking=lambda a, b, timeout=None: yarp.Network.connect(a, b), new_port=new_port,
                        frame_to_list=refshape.kdlFrameToList)
    with redirect_stdout(io.StringIO()):
        exec(compile(src, "<reference src/handlers.py>", "exec"), mod.__dict__)
        rec = handler_calls(mod, prefix, sinks, drain)
    yarp.Network.reset()
    return {"handlers_record": np.array([json.dumps(rec)])}


def main():
    if not os.path.isdir(REF):
        raise SystemExit("reference not found at %s: golden vectors can only be regenerated in the build container" % REF)
    os.makedirs(OUT_DIR, exist_ok=True)
    rng = np.random.default_rng(20261018)
    data = {}
    data.update(gen_mixer(rng))
    data.update(gen_nullspace(rng))
    data.update(gen_bridge(rng))
    data.update(gen_feeder())
    data.update(gen_jp(rng))
    data.update(gen_vf(rng))
    data.update(gen_bridge_loop(rng))
    data.update(gen_dmonitor(rng))
    data.update(gen_nullspace_loop(rng))
    data.update(gen_handlers())
    data.update(gen_nullspace_wide())
    data.update(gen_bridge_readpos())
    path = os.path.join(OUT_DIR, "reference_vectors.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, {k: v.shape for k, v in data.items()})


if __name__ == "__main__":
    main()
