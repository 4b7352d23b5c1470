"""GPU tests of the reference-shaped host modules (ports in, ports out) against the oracle's reference-shaped loop."""
import io
import time
from contextlib import redirect_stdout

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def fresh_ports():
    from vfclik_b200 import ports as yarp
    yarp.Network.reset()
    yield yarp
    yarp.Network.reset()


def test_vfclik_app_config1_matches_reference_shaped_loop(lwr, built_lib, fresh_ports):
    """BASELINE config 1 through the port-driven modules: single LWR, the reference's first goal, three ObstacleP,
    simulation plant.  Every published qdot and the joint trajectory equal oracle/refshape.py's loop."""
    from oracle import batch, refshape
    from vfclik_b200 import workloads
    from vfclik_b200.launcher import Vfclik
    chain, cfg = lwr
    w = workloads.config1(chain, cfg)
    app = Vfclik(cfg, namespace="/0", sim=True, precision=64)
    prm = batch.Params(jp_ref=tuple(cfg.initial_joint_pos), speed_scale=cfg.speedScale, dt=cfg.rate, max_vel=cfg.max_vel,
                       jp_kp=cfg.jpctrl_kp, ik_lambda=cfg.ik_lambda, ns_lambda=cfg.ns_lambda)
    try:
        obj = _out_port(fresh_ports, "/0/test/obj", app.ofeeder.objectPort.getName())
        # the feeder's first iteration replaces whatever it read by config.initial_vf_pose (scripts/object_feeder:98-103),
        # so let it publish its first goal before the obstacles are sent
        assert app.ofeeder.update() and sorted(app.ofeeder.objects) == [0]
        for m in range(3):                               # "set ObstacleP n (16 + radius + order)" (old/README.old:75)
            o = w["obst"][m, 0]
            frame = [1, 0, 0, o[0], 0, 1, 0, o[1], 0, 0, 1, o[2], 0, 0, 0, 1]
            fresh_ports.write_bottle_lists(obj, ["set", "ObstacleP", m, [float(v) for v in frame] + [float(o[3]), 20.0]], strict=True)
        loop = refshape.ControlLoop(chain, prm, cfg.initial_joint_pos, cfg.initial_vf_pose[2],
                                    obstacles=[list(w["obst"][m, 0]) for m in range(3)])
        K = 60
        for k in range(K):
            cmd = app.step()
            loop.cycle()
            if k == 0:
                assert sorted(app.runtime.vectorFields) == [1, 5, 6, 7]            # ids of scripts/object_feeder:229-334
                continue
            assert np.allclose(app.vf.last_qdot, loop.last["qdot_vf"], rtol=1e-9, atol=1e-12), k
            assert np.allclose(app.nullspace.last_qdot, loop.last["qdot_ns"], rtol=1e-9, atol=1e-12), k
        # the modular loop lags one period behind the synchronous oracle (cold start), so compare one step shifted
        loop2 = refshape.ControlLoop(chain, prm, cfg.initial_joint_pos, cfg.initial_vf_pose[2],
                                     obstacles=[list(w["obst"][m, 0]) for m in range(3)])
        loop2.run(K - 1)
        assert np.allclose(app.joint_sim.q, loop2.q, rtol=1e-9, atol=1e-11)
        assert app.runtime.cycles == K and app.bridge.direct_control is False
        # controller switching (src/handlers.py:189-211): joint mode [0,0,1,0,0,0]
        fresh_ports.write_bottle_lists(_out_port(fresh_ports, "/0/test/w", app.bridge.weight_port.getName()), [0.0, 0.0, 1.0])
        app.step(); app.step()
        assert app.bridge.mixer.weights[:3] == [0.0, 0.0, 1.0]
        q = np.asarray(app.bridge.last_q)
        want = cfg.jpctrl_kp * (np.asarray(cfg.initial_joint_pos) - q)
        assert np.allclose(app.bridge.last_cmd, want, rtol=1e-9, atol=1e-12)
    finally:
        app.close()


def _out_port(yarp, name, dst):
    p = yarp.BufferedPortBottle()
    p.open(name)
    yarp.Network.connect(name, dst)
    return p


def test_vf_module_protocol_edge_cases(lwr, built_lib, fresh_ports):
    """Malformed messages are ignored with a warning, never raised (SURVEY.md section 8b error conventions)."""
    from vfclik_b200.runtime import ControlRuntime
    from vfclik_b200.vf import VectorFieldModule
    yarp = fresh_ports
    chain, cfg = lwr
    rt = ControlRuntime(cfg, precision=64)
    vf = VectorFieldModule(rt, "/0")
    base = "/0" + cfg.robotarm_portbasename + "/vectorField"
    param = _out_port(yarp, "/t/param", base + "/param")
    weight = _out_port(yarp, "/t/weight", base + "/weight")
    maxvel = _out_port(yarp, "/t/maxvel", base + "/max_vel")
    qin = _out_port(yarp, "/t/q", base + "/qIn")
    try:
        buf = io.StringIO()
        with redirect_stdout(buf):
            yarp.write_bottle_lists(param, ["add", 1, 1.0, 1, list(cfg.initial_vf_pose[2])], strict=True)
            yarp.write_bottle_lists(param, ["add", 9, -50.0, 7, [0.0] * 8], strict=True)     # a type vfl does not have
            yarp.write_bottle_lists(param, ["add", 3, 1.0], strict=True)                      # wrong arity
            yarp.write_bottle_lists(param, ["remove", 77], strict=True)                       # unknown id: silently nothing
            yarp.write_bottle_lists(weight, ["t", 1.0, 1.0], strict=True)                     # wrong size
            yarp.sendListPort(maxvel, [0.9])                                                  # above the 0.41 cap
            vf.update()
        out = buf.getvalue()
        assert "Unknown vector field type, ignoring" in out and "expected 5" in out and "Wrong size" in out
        assert "speedScale" in out and rt.params.speed_scale == cfg.speedScale
        assert sorted(rt.vectorFields) == [1] and rt.params.w_task == (1.0,) * 6
        yarp.sendListPort(maxvel, [0.3])
        yarp.write_bottle_lists(weight, ["j", 1, 1, 0.5, 1, 1, 1, 1], strict=True)
        yarp.sendListPort(qin, [0.1] * 6)                                                      # wrong length: ignored
        vf.update()
        assert rt.params.speed_scale == 0.3 and rt.params.w_joint[2] == 0.5 and vf.last_qdot is None
        yarp.sendListPort(qin, cfg.initial_joint_pos)
        vf.update()
        assert vf.last_qdot is not None and np.all(np.isfinite(vf.last_qdot)) and np.max(np.abs(vf.last_qdot)) > 1e-3
    finally:
        vf.close()
        rt.close()


def test_command_mixer_class_reproduces_reference_golden_sequence(golden, built_lib, monkeypatch):
    """The product CommandMixer (port logic on the host, sum in vfk_mix) on the scripted sequence recorded from the
    real src/command_mixer.py."""
    from vfclik_b200 import command_mixer, ports as yarp
    ev, wev, short = golden["mixer_events"], golden["mixer_wevents"], golden["mixer_short"]
    steps, n_ports, n = ev.shape
    now = [1000.0]
    monkeypatch.setattr(time, "time", lambda: now[0])
    ports = [yarp.BufferedPortBottle() for _ in range(n_ports)]
    wport = yarp.BufferedPortBottle()
    mixer = command_mixer.CommandMixer(ports, wport, n, 2.0, [1.0, 1.0, 0.0, 0.0, 0.0, 0.0])
    assert mixer.nChannels == n and mixer.guard_time == 2.0 and len(mixer.last_command) == n_ports
    for s in range(steps):
        now[0] += float(golden["mixer_dts"][s])
        for p in range(n_ports):
            if not np.isnan(ev[s, p, 0]):
                ports[p]._deliver(yarp.Bottle(list(ev[s, p])))
            elif short[s, p]:
                ports[p]._deliver(yarp.Bottle([0.0] * (n - 2)))
        wv = wev[s][~np.isnan(wev[s])]
        if wv.size:
            wport._deliver(yarp.Bottle(list(wv)))
        with redirect_stdout(io.StringIO()):
            out = mixer.read()
        assert np.allclose(out, golden["mixer_out"][s], rtol=1e-14, atol=1e-15), s      # fma vs mul+add: <= 1 ulp
        assert np.array_equal(np.asarray(mixer.weights), golden["mixer_weights_after"][s])
    with redirect_stdout(io.StringIO()) as buf:
        m2 = command_mixer.CommandMixer([yarp.BufferedPortBottle(), yarp.BufferedPortBottle()], None, 3, 1.0, [1.0])
    assert m2.weights == [0.0, 0.0]


def test_set_vel_kernel(lwr, built_lib):
    """vfk_set_vel == LWR_Bridge.set_vel (scripts/bridge:188-203) on a batch, both command forms."""
    from vfclik_b200.engine import DeviceBatch, Engine
    chain, _ = lwr
    e = Engine(chain, precision=64)
    try:
        n = 777
        rng = np.random.default_rng(21)
        qd = rng.normal(scale=0.8, size=(7, n)); q = rng.normal(size=(7, n)); qc = q + rng.normal(scale=0.01, size=(7, n))
        db = DeviceBatch(e, n, 0, outputs=("cmd", "qdot"))
        b_qd, b_q, b_qc = db.to_blocked(qd), db.to_blocked(q), db.to_blocked(qc)
        lead = np.max(np.abs(qd), axis=0)
        ratio = np.where(lead > 1.0, 1.0 / lead, 1.0)
        for direct in (False, True):
            e.set_vel(b_qd, b_q, db.t["cmd"], 1.0, direct, 7, n, q_cmded=b_qc, qdot_lim_out=db.t["qdot"])
            want = qd * ratio if direct else (-qc + q + qd * ratio)
            assert np.allclose(db.download("cmd"), want, rtol=1e-14, atol=1e-15)
            assert np.allclose(db.download("qdot"), qd * ratio, rtol=1e-14, atol=1e-15)
        assert np.any(ratio < 1) and np.any(ratio == 1)
    finally:
        e.close()


def test_monitor_kernel_matches_reference_shaped_monitor(lwr, built_lib):
    """vfk_monitor (tracking diagnostics + distance monitor) over 40 cycles of 200 instances vs the
    scalar restatement of scripts/vf:349-428 and scripts/monitor_distance:148-219."""
    from oracle import monitor as omon
    from oracle.refshape import KdlFrame
    from vfclik_b200 import workloads
    from vfclik_b200.engine import DeviceBatch, Engine, Params
    chain, cfg = lwr
    e = Engine(chain, precision=64, params=Params.from_config(cfg, speed_scale=0.4, dt=0.02))
    try:
        n, M, K = 200, 4, 40
        w = workloads.random_batch(chain, n, M, seed=61)
        w["goal"][12] = 0.3                                   # wide slow-down ball: some instances reach "on goal"
        db = DeviceBatch(e, n, M, outputs=("qdot", "pose", "twist"))
        db.upload("q", w["q"]); db.upload("goal", w["goal"]); db.upload("obst", w["obst"])
        diag = [omon.TrackingDiagnostics() for _ in range(n)]
        dmon = [omon.DistanceMonitor() for _ in range(n)]
        goals = [KdlFrame(w["goal"][0:9, i].reshape(3, 3), w["goal"][9:12, i]) for i in range(n)]
        seen_states = set()
        for k in range(K):
            db.step(1)
            db.monitor()
            pose, tw = db.download("pose"), db.download("twist")
            track, dist, st = db.download("track"), db.download("dist"), db.download("tracking_state")
            for i in range(0, n, 7):
                f = KdlFrame(pose[0:9, i].reshape(3, 3), pose[9:12, i])
                te = diag[i].update(f, tw[0:3, i], tw[3:6, i])
                dx, dd, mx, mr = dmon[i].update(f, goals[i], te)
                if te is None:
                    assert np.all(track[:, i] == 0.0)
                else:
                    assert np.allclose(track[:, i], te, rtol=1e-6, atol=1e-7), (k, i, track[:, i], te)
                assert np.isclose(dist[0, i], dx, rtol=1e-12, atol=1e-14) and np.isclose(dist[1, i], dd, rtol=1e-9, atol=1e-9)
                want = (-1 if mx is None else omon.STATES.index(mx), -1 if mr is None else omon.STATES.index(mr))
                assert (st[0, i], st[1, i]) == want, (k, i, st[:, i], want)
                seen_states.update(want)
        assert {-1, 1}.issubset(seen_states)
    finally:
        e.close()


def test_handlers_drive_the_arm_through_ports(lwr, built_lib, fresh_ports):
    """src/handlers.py surface: HandleArm.gotoFrame waits on /dmonitor/distOut, HandleBridge switches controllers,
    HandleJController.set_ref_js waits on /bridge/encoders -- all through ports, spinning the single-process app."""
    from vfclik_b200 import handlers
    from vfclik_b200.launcher import Vfclik
    chain, cfg = lwr
    app = Vfclik(cfg, namespace="/0", sim=True, precision=64)
    try:
        prename = "/0" + cfg.robotarm_portbasename
        arm = handlers.HandleArm(cfg.robotarm_portbasename, namespace="/0", spin=app.step)
        bridge = handlers.HandleBridge(prename, spin=app.step, torso=False)
        jctrl = handlers.HandleJController(prename, spin=app.step)
        for _ in range(3):
            app.step()
        pose0 = arm.getPose()
        assert len(pose0) == 16 and abs(pose0[15] - 1.0) < 1e-12
        # a reachable Cartesian goal 12 cm from the start pose, same orientation
        goal = list(pose0)
        goal[3] += 0.08; goal[7] -= 0.06; goal[11] += 0.06
        app.runtime.set_speed_scale(0.4)
        arm.current_slowdown_distance = 0.05
        ok, diff = arm.gotoFrame(goal, wait=60.0, goal_precision=[0.004, 0.02])
        assert ok and diff[0] < 0.004 and diff[1] < 0.02, diff
        assert app.dmonitor.last["dist"][0][0] == 0
        # joint mode: go back to the initial posture
        bridge.joint_controller()
        ref = list(cfg.initial_joint_pos)
        ok, diff = jctrl.set_ref_js(ref, wait=60.0, goal_precision=[0.01] * 7)
        assert ok and np.max(np.abs(diff)) <= 0.01
        assert app.bridge.mixer.weights[:4] == [0.0, 0.0, 1.0, 0.0] and app.jpctrl.at_goal == 1
        bridge.cartesian_controller()
        bridge.set_weights("task", [1, 1, 1, 0.5, 0.5, 0.5])
        app.step(); app.step()
        assert app.bridge.mixer.weights[:4] == [1.0, 1.0, 0.0, 0.0] and app.runtime.params.w_task[3] == 0.5
        assert len(bridge.read_joint_angles()) == 7
    finally:
        app.close()


def test_object_feeder_goal_and_normal_and_table(lwr, built_lib, fresh_ports):
    """f2: "set goalAndNormal" (attractor + funnel type 5 + near-goal repeller) and "set ObstacleH" (hemisphere type 4)
    through /ofeeder/object -> /vectorField/param -> the GPU's auxiliary-field slots, vs the reference-shaped loop."""
    from oracle import batch, refshape
    from vfclik_b200.launcher import Vfclik
    chain, cfg = lwr
    app = Vfclik(cfg, namespace="/0", sim=True, precision=64)
    prm = batch.Params(jp_ref=tuple(cfg.initial_joint_pos), speed_scale=cfg.speedScale, dt=cfg.rate, max_vel=cfg.max_vel,
                       jp_kp=cfg.jpctrl_kp, ik_lambda=cfg.ik_lambda, ns_lambda=cfg.ns_lambda)
    try:
        obj = _out_port(fresh_ports, "/0/test/obj", app.ofeeder.objectPort.getName())
        app.ofeeder.update()                                             # first goal from config.initial_vf_pose
        g = list(cfg.initial_vf_pose[2][:16])
        gan = g + [0.0, 0.3, -1.0, 0.5, 0.15, 0.04]                    # axis, cut angle, cut length, slowdown
        table = [1, 0, 0, 0.7, 0, 1, 0, 0.1, 0, 0, 1, 0.9, 0, 0, 0, 1] + [0.0, 0.0, 1.0, 0.02, 3.0]   # normal, safe, order
        fresh_ports.write_bottle_lists(obj, ["set", "goalAndNormal", [float(v) for v in gan]], strict=True)
        fresh_ports.write_bottle_lists(obj, ["set", "ObstacleH", 0, [float(v) for v in table]], strict=True)
        app.step()
        vfields = app.runtime.vectorFields
        assert sorted(vfields) == [1, 2, 3, 5] and [vfields[k][1] for k in (1, 2, 3, 5)] == [1, 5, 2, 4]
        assert [vfields[k][0] for k in (1, 2, 3, 5)] == [1.0, 30.0, -10.0, -50.0]          # scripts/object_feeder:236,268,288,341
        extra = [(k, vfields[k][0], vfields[k][1], vfields[k][2]) for k in (2, 3, 5)]
        loop = refshape.ControlLoop(chain, prm, cfg.initial_joint_pos, vfields[1][2], extra_fields=extra)
        loop.cycle()
        assert np.allclose(app.vf.last_qdot, loop.last["qdot_vf"], rtol=1e-9, atol=1e-12)
        for k in range(25):
            app.step(); loop.cycle()
            assert np.allclose(app.vf.last_qdot, loop.last["qdot_vf"], rtol=1e-8, atol=1e-11), k
        # removing the table through the feeder frees the auxiliary slot again
        fresh_ports.write_bottle_lists(obj, ["remove", 0], strict=True)
        app.step()
        assert 5 not in app.runtime.vectorFields
    finally:
        app.close()


def test_bridge_backends_powercube_icub_and_torso_sharing(lwr, built_lib, fresh_ports):
    """SURVEY.md 8 row f4 through the port-driven bridge module: Powercube shoulder clamp, iCub torso (de)activation
    weights, and the LWR bridge taking its torso joints from the other arm's plant (scripts/bridge:140-147,174-178)."""
    import copy
    from vfclik_b200.launcher import Vfclik
    _, cfg = lwr

    # --- Powercube: leading clamp + shoulder clamp, command = qdot_lim (scripts/bridge:288-312)
    c1 = copy.copy(cfg)
    c1.arm_type, c1.max_vel_shoulder_pos, c1.max_vel_shoulder_neg, c1.speedScale, c1.max_vel = "powercube", 0.004, -0.004, 0.41, 0.05
    app = Vfclik(c1, namespace="/1", sim=True, precision=64)
    try:
        hit = 0
        for _ in range(6):
            cmd = app.step()
        for _ in range(10):
            cmd = app.step()
            mix = [a + b for a, b in zip(app.vf.last_qdot, app.nullspace.last_qdot)]      # mixer weights [1,1,0,0,0,0]
            leading_vel = max(map(abs, mix))
            ratio = c1.max_vel / leading_vel if leading_vel > c1.max_vel else 1.0
            qdot_lim = [v * ratio for v in mix]
            shoulder_vel = qdot_lim[0]
            if shoulder_vel > c1.max_vel_shoulder_pos:
                ratio = abs(c1.max_vel_shoulder_pos / shoulder_vel); hit += 1
            elif shoulder_vel < c1.max_vel_shoulder_neg:
                ratio = abs(c1.max_vel_shoulder_neg / shoulder_vel); hit += 1
            qdot_lim = [i * ratio for i in qdot_lim]
            assert np.allclose(cmd, qdot_lim, rtol=1e-9, atol=1e-13)
        assert hit > 0
        assert app.bridge.qin_port.getName() == "/1" + cfg.robotarm_portbasename + "/bridge/qin"
    finally:
        app.close()
    fresh_ports.Network.reset()

    # --- iCub: torso_cjoints -> ['j', w0..w9] on /control_weights:o (scripts/bridge:470-506)
    c2 = copy.copy(cfg)
    c2.arm_type, c2.icub_torso_cjoints = "icub", [True, True, True]
    app = Vfclik(c2, namespace="/2", sim=True, precision=64)
    try:
        base = "/2" + cfg.robotarm_portbasename + "/bridge"
        probe = fresh_ports.BufferedPortBottle(); probe.open("/2/test/weights_probe")
        fresh_ports.Network.connect(base + "/control_weights:o", "/2/test/weights_probe")
        tc = _out_port(fresh_ports, "/2/test/tc", base + "/torso_cjoints:i")
        fresh_ports.write_bottle_lists(tc, [1, 0, 1], strict=True)
        with redirect_stdout(io.StringIO()) as log:
            app.step()
        b = probe.read(False)
        assert b is not None and b.get(0).asString() == "j"
        assert [b.get(i).asDouble() for i in range(1, b.size())] == [1.0, 0.0, 1.0] + [1.0] * 7
        assert app.bridge.icub_torso_cjoints == [True, False, True]
        assert "Received a torso_joints bottle" in log.getvalue()
        fresh_ports.write_bottle_lists(tc, [1, 0], strict=True)
        with redirect_stdout(io.StringIO()) as log:
            cmd = app.step()
        assert "wrong length" in log.getvalue() and probe.read(False) is None
        # the iCub bridge commands qdot_lim itself, never the LWR offset form
        mix = np.asarray(app.vf.last_qdot) + np.asarray(app.nullspace.last_qdot)
        lead = np.max(np.abs(mix))
        assert np.allclose(cmd, mix * (min(1.0, cfg.max_vel / lead)), rtol=1e-9, atol=1e-13)
    finally:
        app.close()
    fresh_ports.Network.reset()

    # --- LWR with the torso on the other arm's plant
    c3 = copy.copy(cfg)
    c3.torso_joints, c3.torso_instance, c3.torso_qin_portname = [0, 1], "left", "/torso_qin"
    other = fresh_ports.BufferedPortBottle(); other.open("/3/lwr/left/joint_sim/qout")
    app = Vfclik(c3, namespace="/3", sim=True, precision=64)
    try:
        fresh_ports.write_bottle_lists(other, [0.11, -0.22, 9, 9, 9, 9, 9], strict=True)
        app.step()
        assert app.bridge.last_q[:2] == [0.11, -0.22]                     # torso joints: the other arm's plant
        assert np.allclose(app.bridge.last_q[2:], cfg.initial_joint_pos[2:])   # the rest: this arm's own joint_sim
    finally:
        app.close()


def test_module_programs_run_standalone(built_lib):
    """`vf -c <config> -n <ns>`, `nullspace ...`, `joint_p_controller ...`, `bridge ... -s` (scripts/vfclik:88-105) each start
    on their own runtime, poll for --cycles iterations and exit 0."""
    import os
    import subprocess
    import sys
    from vfclik_b200.config import PACKAGE_CONFIG_DIR, config_filename
    cfgfile = config_filename(PACKAGE_CONFIG_DIR + "/lwr/", "lwr", "right")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for mod, extra in (("vf", []), ("nullspace", []), ("joint_p_controller", []), ("monitor_distance", []), ("bridge", ["-s"])):
        r = subprocess.run([sys.executable, "-m", "vfclik_b200." + mod, "-c", cfgfile, "-n", "/9", "--cycles", "3", "--no_sleep"] + extra,
                           cwd=root, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, (mod, r.stdout[-500:], r.stderr[-1500:])
        assert "iterations: 3" in r.stdout, (mod, r.stdout[-500:])


def test_joint_controller_position_dependent_limits(lwr, built_lib, fresh_ports, capsys):
    """Row a12: the joint reference is clamped into config.updateJntLimits(q), which may depend on the posture
    (scripts/joint_p_controller:79-89)."""
    import copy
    from vfclik_b200.joint_p_controller import JointPControllerModule
    from vfclik_b200.runtime import ControlRuntime
    _, cfg = lwr
    c = copy.copy(cfg)
    c.updateJntLimits = lambda q: [[-0.5 - abs(q[0]), 0.5 + abs(q[0])]] * 7        # widens with |q0|
    rt = ControlRuntime(c, n_instances=1, precision=64)
    jp = JointPControllerModule(rt, "/5")
    try:
        enc = _out_port(fresh_ports, "/5/test/enc", jp.inPort.getName())
        ref = _out_port(fresh_ports, "/5/test/ref", jp.refPort.getName())
        out = fresh_ports.BufferedPortBottle(); out.open("/5/test/out")
        fresh_ports.Network.connect(jp.outPort.getName(), "/5/test/out")
        fresh_ports.sendListPort(ref, [1.0, -1.0, 0.2, 0.0, 0.0, 0.0, 0.0])
        ref_now = np.array([1.0, -1.0, 0.2, 0, 0, 0, 0])
        for q0, lim in ((0.3, 0.8), (0.0, 0.5), (0.3, 0.8)):
            q = [q0, 0.1, 0.0, 0.0, 0.0, 0.0, 0.0]
            fresh_ports.sendListPort(enc, q)
            assert jp.update()
            b = out.read(False)
            got = [b.get(i).asDouble() for i in range(7)]
            ref_now = np.clip(ref_now, -lim, lim)        # the clamp persists (scripts/joint_p_controller:121): 0.8 -> 0.5 -> stays 0.5
            want = cfg.jpctrl_kp * (ref_now - np.asarray(q))
            assert np.allclose(got, want, rtol=1e-12, atol=1e-14), (q0, got, want)
        assert np.allclose(ref_now[:3], [0.5, -0.5, 0.2])
        text = capsys.readouterr().out
        assert "Limiting high 0" in text and "Limiting low 1" in text
    finally:
        jp.close(); rt.close()


def test_set_vel_kernel_against_reference_golden_vectors(lwr, golden, built_lib):
    """vfk_set_vel for the LWR (both command forms), Powercube and iCub back-ends against vectors produced by executing the
    reference's own set_vel methods (scripts/bridge:182-210, 288-312, 507-530; oracle/gen_golden.py:gen_bridge)."""
    import dataclasses
    from vfclik_b200.engine import BRIDGE_ICUB, BRIDGE_LWR, BRIDGE_POWERCUBE, DeviceBatch, Engine, Params
    chain, cfg = lwr
    g = golden
    qd, q, qc = g["br_qdot"].T.copy(), g["br_q"].T.copy(), g["br_qcmded"].T.copy()          # [7, cases]
    max_vel, sh_pos, sh_neg = [float(v) for v in g["br_cfg"]]
    n = qd.shape[1]
    for kind, direct, key in ((BRIDGE_LWR, False, "br_lwr_cmd_direct0"), (BRIDGE_LWR, True, "br_lwr_cmd_direct1"),
                              (BRIDGE_POWERCUBE, False, "br_powercube_cmd"), (BRIDGE_ICUB, False, "br_icub_cmd")):
        prm = dataclasses.replace(Params.from_config(cfg), bridge_kind=kind, shoulder_vel=(sh_pos, sh_neg), max_vel=max_vel)
        e = Engine(chain, precision=64, params=prm)
        try:
            db = DeviceBatch(e, n, 0, outputs=("cmd", "qdot"))
            e.set_vel(db.to_blocked(qd), db.to_blocked(q), db.t["cmd"], max_vel, direct, 7, n, q_cmded=db.to_blocked(qc),
                      qdot_lim_out=db.t["qdot"])
            assert np.allclose(db.download("cmd").T, g[key], rtol=1e-14, atol=1e-15), key
        finally:
            e.close()


def test_joint_controller_module_matches_the_reference_loop(lwr, golden, built_lib, fresh_ports, capsys):
    """/jpctrl/out and /jpctrl/at_goal against what the reference's own loop body (scripts/joint_p_controller:96-146, with
    its check_limits) produced for a scripted sequence of /ref and /in messages -- static LWR limits and posture-dependent
    limits, where the persisting clamp and the signed at-goal compare both show."""
    import copy
    from vfclik_b200.joint_p_controller import JointPControllerModule
    from vfclik_b200.runtime import ControlRuntime
    _, cfg = lwr
    g = golden
    steps = {int(k): g["jp_ref_msgs"][i] for i, k in enumerate(g["jp_ref_steps"])}
    kp, delta = [float(v) for v in g["jp_kp_delta"]]
    hooks = {"static": lambda cur: [list(r) for r in g["jp_static_limits"]],
             "dynamic": lambda cur: [[-0.5 - abs(cur[0]), 0.4 + abs(cur[1])]] * 7}
    for name, hook in hooks.items():
        fresh_ports.Network.reset()
        c = copy.copy(cfg)
        c.updateJntLimits, c.jpctrl_kp = hook, kp
        rt = ControlRuntime(c, n_instances=1, precision=64)
        jp = JointPControllerModule(rt, "/6")
        try:
            enc = _out_port(fresh_ports, "/6/test/enc", jp.inPort.getName())
            ref = _out_port(fresh_ports, "/6/test/ref", jp.refPort.getName())
            out = fresh_ports.BufferedPortBottle(); out.open("/6/test/out")
            goal = fresh_ports.BufferedPortBottle(); goal.open("/6/test/goal")
            fresh_ports.Network.connect(jp.outPort.getName(), "/6/test/out")
            fresh_ports.Network.connect(jp.atGoalPort.getName(), "/6/test/goal")
            for k in range(g["jp_q"].shape[0]):
                if k in steps:
                    fresh_ports.sendListPort(ref, [float(v) for v in steps[k]])
                fresh_ports.sendListPort(enc, [float(v) for v in g["jp_q"][k]])
                assert jp.update()
                b, a = out.read(False), goal.read(False)
                got = [b.get(i).asDouble() for i in range(7)]
                assert np.allclose(got, g["jp_out_" + name][k], rtol=1e-12, atol=1e-14), (name, k)
                assert a.get(0).asInt() == int(g["jp_at_goal_" + name][k]), (name, k)
        finally:
            jp.close(); rt.close()
    capsys.readouterr()
    assert g["jp_at_goal_static"].max() == 1 and g["jp_at_goal_static"].min() == 0


def test_vf_module_matches_the_reference_loop_cycle_by_cycle(lwr, golden, built_lib, fresh_ports, capsys):
    """Every port of the vf module against what the reference's own loop body (scripts/vf:193-505; PyKDL / Lafik / vfl replaced
    by the oracle's stand-ins, oracle/gen_golden.py:gen_vf) published for a scripted run: field add / remove / unknown type,
    task and joint weights (one with a wrong length), a tool frame, speed-scale changes (one out of range), a pose query, the
    message swallowed by `start_attractor`, tracking errors from the 6th frame, vector_out / goal_out every 21st cycle."""
    import json
    from vfclik_b200.runtime import ControlRuntime
    from vfclik_b200.vf import VectorFieldModule
    _, cfg = lwr
    g = golden
    script = {int(k): v for k, v in json.loads(str(g["vf_script"][0])).items()}
    rt = ControlRuntime(cfg, n_instances=1, precision=64)
    vf = VectorFieldModule(rt, "/8")
    names = {"qdotOutPort": vf.qdotOutPort, "posePort": vf.posePort, "pose_no_tool_Port": vf.pose_no_tool_Port,
             "tracking_error_port": vf.tracking_error_port, "vector_port": vf.vector_port, "goal_port": vf.goal_port}
    ins = {"maxvel_port": vf.maxvel_port, "paramPort": vf.paramPort, "weightPort": vf.weightPort, "qInPort": vf.qInPort,
           "toolPort": vf.toolPort, "pose_in_port": vf.pose_in_port}
    try:
        sinks = {}
        for name, port in names.items():
            p = fresh_ports.BufferedPortBottle(); p.open("/8/sink/" + name); p.setStrict(True)
            fresh_ports.Network.connect(port.getName(), "/8/sink/" + name)
            sinks[name] = p
        feeds = {name: _out_port(fresh_ports, "/8/feed/" + name, port.getName()) for name, port in ins.items()}
        for k in range(g["vf_q"].shape[0]):
            for port, msg in script.get(k, {}).items():
                fresh_ports.write_bottle_lists(feeds[port], msg, strict=True)
            fresh_ports.sendListPort(feeds["qInPort"], [float(v) for v in g["vf_q"][k]])
            vf.update()
            want = json.loads(str(g["vf_out"][k]))
            for name, p in sinks.items():
                got = []
                while True:
                    b = p.read(False)
                    if b is None:
                        break
                    got.append(b.to_list())
                assert len(got) == len(want[name]), (k, name, len(got), len(want[name]))
                for a, w in zip(got, want[name]):
                    if name == "tracking_error_port":
                        assert np.allclose(a[:7], w[:7], rtol=1e-6, atol=1e-9), (k, name, a, w)
                        assert int(a[7]) == int(w[7]), (k, name)
                    else:
                        assert np.allclose(a, w, rtol=1e-8, atol=1e-11), (k, name, a, w)
    finally:
        vf.close(); rt.close()
    capsys.readouterr()


def test_bridge_module_matches_the_reference_main_loop(lwr, golden, built_lib, fresh_ports, capsys, monkeypatch):
    """/bridge/encoders, the robot command and /bridge/current_weights against what the reference's own main loop body
    (scripts/bridge:600-634, with its read_pos / set_vel and the real CommandMixer; oracle/gen_golden.py:gen_bridge_loop)
    produced for a scripted run: /cmded feedback, three command ports at different rates, weight changes down to all-zero
    (direct control, one iteration late like the reference), max_vel changes (one refused)."""
    import copy
    import json
    from vfclik_b200.bridge import BridgeModule
    from vfclik_b200.runtime import ControlRuntime
    import types
    from vfclik_b200 import command_mixer
    _, cfg = lwr
    g = golden
    # the guard time is wall-clock (src/command_mixer.py:64-66): the first CUDA calls of a fresh process take longer than its
    # 2 s, which would zero the slower ports after the first iteration -- run the mixer on a scripted clock instead
    ticks = iter(range(10 ** 9))
    monkeypatch.setattr(command_mixer, "time", types.SimpleNamespace(time=lambda: 0.001 * next(ticks)))
    c = copy.copy(cfg)
    c.max_vel = float(g["bl_cfg"][0])
    rt = ControlRuntime(c, n_instances=1, precision=64)
    base = "/4" + c.robotarm_portbasename
    robot = {name: fresh_ports.BufferedPortBottle() for name in ("pos", "cmded", "cmd")}
    for name, p in robot.items():
        p.open(base + "/robot/" + name)
    robot["cmd"].setStrict(True)
    br = BridgeModule(rt, "/4", sim=False)
    try:
        sinks = {}
        for name, port in (("encoders", br.encoders_port), ("current_weights", br.current_weight_port)):
            p = fresh_ports.BufferedPortBottle(); p.open("/4/sink/" + name); p.setStrict(True)
            fresh_ports.Network.connect(port.getName(), "/4/sink/" + name)
            sinks[name] = p
        sinks["qcmd"] = robot["cmd"]
        feeds = [_out_port(fresh_ports, "/4/feed/cmd%d" % i, br.cmd_ports[i].getName()) for i in range(3)]
        wfeed = _out_port(fresh_ports, "/4/feed/w", br.weight_port.getName())
        mfeed = _out_port(fresh_ports, "/4/feed/mv", br.maxvel_port.getName())
        for k in range(g["bl_script"].shape[0]):
            ev = json.loads(str(g["bl_script"][k]))
            fresh_ports.sendListPort(robot["pos"], ev["q"])
            if "cmded" in ev:
                fresh_ports.sendListPort(robot["cmded"], ev["cmded"])
            for p, v in ev["cmd"].items():
                fresh_ports.sendListPort(feeds[int(p)], v)
            if "weights" in ev:
                fresh_ports.sendListPort(wfeed, ev["weights"])
            if "max_vel" in ev:
                fresh_ports.sendListPort(mfeed, [ev["max_vel"]])
            br.update()
            br.finish()
            want = json.loads(str(g["bl_out"][k]))
            for name, p in sinks.items():
                got = []
                while True:
                    b = p.read(False)
                    if b is None:
                        break
                    got.append(b.to_list())
                assert len(got) == len(want[name]) == 1, (k, name, len(got))
                assert np.allclose(got[0], want[name][0], rtol=1e-12, atol=1e-14), (k, name, got[0], want[name][0])
    finally:
        br.close(); rt.close()
        for p in robot.values():
            p.close()
    capsys.readouterr()


def test_vf_module_drains_queued_param_messages(lwr, built_lib, fresh_ports, capsys):
    """Deliberate deviation: one ``update()`` is one control cycle and applies EVERY ``/param`` message
    queued since the last one (the reference's loop takes one per pass but spins ~10 passes per control period), except on
    the pass after the first joint data, where the message read is swallowed by ``start_attractor`` as in the reference."""
    from vfclik_b200.runtime import ControlRuntime
    from vfclik_b200.vf import VectorFieldModule
    _, cfg = lwr
    rt = ControlRuntime(cfg, n_instances=1, precision=64)
    vf = VectorFieldModule(rt, "/6")
    try:
        pfeed = _out_port(fresh_ports, "/6/feed/param", vf.paramPort.getName())
        qfeed = _out_port(fresh_ports, "/6/feed/q", vf.qInPort.getName())
        obstacle = lambda i: ["add", 5 + i, -10.0, 2, [0.4 + 0.05 * i, 0.1, 0.6, 0.05, 0.001, 20.0]]
        q = [float(v) for v in cfg.initial_joint_pos]
        fresh_ports.sendListPort(qfeed, q)
        vf.update()                                                   # first joint data: arms start_attractor
        n0 = len(rt.vectorFields)
        for i in range(3):
            fresh_ports.write_bottle_lists(pfeed, obstacle(i), strict=True)
        fresh_ports.sendListPort(qfeed, q)
        vf.update()                                                   # swallows ONE message (the reference's quirk), applies the rest
        assert len(rt.vectorFields) == n0 + 2 and 5 not in rt.vectorFields
        for i in range(3, 7):
            fresh_ports.write_bottle_lists(pfeed, obstacle(i), strict=True)
        fresh_ports.write_bottle_lists(pfeed, ["remove", 6], strict=True)
        fresh_ports.sendListPort(qfeed, q)
        vf.update()                                                   # all five applied within this one cycle, in order
        assert sorted(k for k in rt.vectorFields if k >= 5) == [7, 8, 9, 10, 11]
    finally:
        vf.close(); rt.close()
    capsys.readouterr()


def test_bridge_read_pos_without_cmded_feedback(lwr, golden, built_lib, fresh_ports, capsys):
    """``LWR_Bridge.read_pos`` with no ``/cmded`` feedback (tests/golden rp_*: the reference's own method executed): the
    commanded position is taken from the FIRST measured q and then kept.  The real-robot path of this repo's bridge does the
    same; in simulation (``sim=True``) it deliberately follows q instead (with the reference's rule the
    non-direct command ``-q_cmded + q + qdot`` sent to the simulated plant as a velocity would carry the whole displacement
    since start-up)."""
    import copy
    from vfclik_b200.bridge import BridgeModule
    from vfclik_b200.runtime import ControlRuntime
    _, cfg = lwr
    g = golden
    for sim in (False, True):
        rt = ControlRuntime(copy.copy(cfg), n_instances=1, precision=64)
        ns = "/8" if sim else "/9"
        base = ns + cfg.robotarm_portbasename
        br = BridgeModule(rt, ns, sim=sim)
        try:
            feed = _out_port(fresh_ports, ns + "/feed/pos", br.qin_port.getName())
            for k in range(g["rp_q"].shape[0]):
                fresh_ports.sendListPort(feed, [float(v) for v in g["rp_q"][k]])
                q = br.read_pos()
                assert np.array_equal(q, g["rp_last_q"][k])
                want = g["rp_q"][k] if sim else g["rp_last_qcmded"][k]
                assert np.array_equal(br.last_qcmded, want), (sim, k)
        finally:
            br.close(); rt.close()
    capsys.readouterr()


def test_nullspace_module_matches_the_reference_main_loop(lwr, golden, built_lib, fresh_ports, capsys):
    """/nullspace/qdotout against the reference's own main loop body (scripts/nullspace:159-187) running its own
    restrict / nullspace / move_in_nullspace / check_limits on the oracle's Lafik stand-in: the four-float control interface,
    sticky control bottles, gain 0.5.  LAPACK's first-cycle sign of the nullspace vector is arbitrary (ORACLE_CHOICES), the
    reference then keeps it by continuity, so the records agree up to ONE global sign."""
    import json
    from vfclik_b200.nullspace import NullspaceModule
    from vfclik_b200.runtime import NS_CONTROL, ControlRuntime
    _, cfg = lwr
    g = golden
    rt = ControlRuntime(cfg, n_instances=1, precision=64)
    rt.set_params(ns_mode=NS_CONTROL, ns_lambda=0.0)
    ns = NullspaceModule(rt, "/7")
    try:
        qfeed = _out_port(fresh_ports, "/7/feed/q", ns.qin_port.getName())
        cfeed = _out_port(fresh_ports, "/7/feed/c", ns.control_port.getName())
        sink = fresh_ports.BufferedPortBottle(); sink.open("/7/sink"); sink.setStrict(True)
        fresh_ports.Network.connect(ns.qdotout_port.getName(), "/7/sink")
        sign = None
        for k in range(g["nl_script"].shape[0]):
            ev = json.loads(str(g["nl_script"][k]))
            if "control" in ev:
                fresh_ports.sendListPort(cfeed, ev["control"])
            fresh_ports.sendListPort(qfeed, ev["q"])
            assert ns.update()
            b = sink.read(False)
            got = np.asarray([b.get(i).asDouble() for i in range(7)])
            want = g["nl_qdotout"][k]
            if sign is None and np.max(np.abs(want)) > 0:
                sign = 1.0 if np.dot(got, want) > 0 else -1.0
            assert np.allclose(got * (sign or 1.0), want, rtol=1e-9, atol=1e-12), (k, got, want)
        assert sign is not None
    finally:
        ns.close(); rt.close()
    capsys.readouterr()
