"""``bridge`` module: the timed robot loop with the reference's port contract (``scripts/bridge``).

``BridgeModule.update()`` + ``finish()`` are one iteration of ``scripts/bridge:600-634``: read the plant's joint
positions, publish ``/encoders``, accept a new ``/max_vel`` in ``[0, config.max_vel]`` (``:613-623``), read the
command mixer over the six command ports in the reference's order (``:593-596``), clamp the leading joint
velocity to ``max_vel`` and form the command (``LWR_Bridge.set_vel``, ``:182-210``), publish the current
weights (``:627``).  The mixer sum goes through ``vfk_mix`` and the clamp / command through ``vfk_set_vel``; the
batched driver for millions of instances gets both from the fused cycle kernel instead.

Simulation (``-s``): the reference wires the command to an external ``joint_sim`` process that integrates it
(``scripts/bridge:136-139``); here ``JointSim`` is the explicit-Euler plant ``q += rate * cmd``.

Back-ends (``construct_arm_bridge``, ``scripts/bridge:102-111``), selected by
``config.arm_type``; only their simulation / YARP side exists here (the Player, iCub ``remote_controlboard`` and RSI
device drivers are hardware I/O):

* ``lwr``       offset command ``-q_cmded + q + qdot_lim``; **torso-joint sharing**: when ``config.torso_joints`` is not
  empty the positions of those joints come from the *other* arm's plant (``config.torso_qin_portname`` connected to
  ``/<arm_type>/<torso_instance>/joint_sim/qout``, ``scripts/bridge:140-147,174-178``);
* ``powercube`` ports ``/qin`` ``/qcmd`` under the bridge prefix, leading clamp + shoulder-speed clamp
  (``config.max_vel_shoulder_pos/neg``, ``scripts/bridge:288-305``, bug-compatible double ratio), command = qdot_lim;
* ``icub``      ports ``/qin`` ``/qcmd`` ``/torso_cjoints:i`` ``/control_weights:o``: a torso_cjoints bottle of three ints
  (de)activates torso joints by sending ``['j', w0..w9]`` joint weights to the vector field
  (``scripts/bridge:437-445,470-506``), command = qdot_lim.
"""
from __future__ import annotations

import numpy as np

from . import ports as yarp
from .command_mixer import CommandMixer
from .ports import sendListPort
from .runtime import ControlRuntime

from .engine import BRIDGE_ICUB, BRIDGE_KINDS, BRIDGE_LWR, BRIDGE_POWERCUBE

MODULE_NAME = "/bridge"
COMMAND_PORTS = ("/vectorfieldcmd", "/nullcmd", "/jointcmd", "/mechanismcmd", "/xtra1cmd", "/xtra2cmd")
DEFAULT_WEIGHTS = [1.0, 1.0, 0.0, 0.0, 0.0, 0.0]      # scripts/bridge:596
GUARD_TIME = 2.0


class JointSim:
    """Stand-in for the external ``joint_sim`` process: ``/qvin`` velocity command in, ``/qout`` positions out."""

    def __init__(self, config, namespace: str = ""):
        base = config.robotarm_portbasename
        self.yarp_ctrl = yarp.ArcosYarp(ports_name_prefix=namespace, module_name_prefix=base + "/joint_sim")
        self.qvin = self.yarp_ctrl.create_yarp_port("/qvin", strict=False)
        self.qout = self.yarp_ctrl.create_yarp_port("/qout", input_port=False)
        self.qin = self.yarp_ctrl.create_yarp_port("/qin", strict=False)       # teleport the simulated arm (HandleArmNew.set_sim_arm_q)
        self.q = [float(v) for v in config.initial_joint_pos]
        self.dt = float(config.rate)

    def update(self):
        t = self.qin.read(False)
        if t and t.size() == len(self.q):
            self.q = [t.get(i).asDouble() for i in range(len(self.q))]
        b = self.qvin.read(False)
        if b and b.size() == len(self.q):
            self.q = [self.q[i] + self.dt * b.get(i).asDouble() for i in range(len(self.q))]
        sendListPort(self.qout, self.q)

    def close(self):
        self.yarp_ctrl.close()


class BridgeModule:
    def __init__(self, runtime: ControlRuntime, namespace: str = "", sim: bool = True):
        self.rt = runtime
        cfg = runtime.config
        self.config = cfg
        base = cfg.robotarm_portbasename
        self.nJoints = cfg.nJoints
        self.sim = sim
        self.yarp_ctrl = yarp.ArcosYarp(ports_name_prefix=namespace, module_name_prefix=base + MODULE_NAME)
        y = self.yarp_ctrl
        self.encoders_port = y.create_yarp_port("/encoders", input_port=False)
        self.cmd_ports = [y.create_yarp_port(name, strict=False) for name in COMMAND_PORTS]
        self.weight_port = y.create_yarp_port("/weight")
        self.current_weight_port = y.create_yarp_port("/current_weights", input_port=False)
        self.maxvel_port = y.create_yarp_port("/max_vel")
        y.connect(self.encoders_port, base + "/vectorField", "/qIn")
        y.connect(self.encoders_port, base + "/nullspace", "/qin", necessary=False)
        y.connect(self.encoders_port, base + "/jpctrl", "/in")
        arm_type = str(getattr(cfg, "arm_type", "lwr")).lower()
        if arm_type not in BRIDGE_KINDS:
            print('warning: do not know class "%s"' % arm_type)            # scripts/bridge:110 (the reference then crashes)
            raise ValueError("unknown config.arm_type %r" % arm_type)
        self.kind = BRIDGE_KINDS[arm_type]
        if runtime.params.bridge_kind != self.kind:
            shoulder = (float(getattr(cfg, "max_vel_shoulder_pos", 0.0)), float(getattr(cfg, "max_vel_shoulder_neg", 0.0)))
            runtime.set_params(bridge_kind=self.kind, shoulder_vel=shoulder)
        self.torso_joints = list(getattr(cfg, "torso_joints", []) or [])
        self.torso_qin_port = None
        self.qcmded_port = None
        if self.kind == BRIDGE_LWR:
            # LWR_Bridge.open (scripts/bridge:127-155)
            self.qin_port = y.create_yarp_port(cfg.qin_portname, strict=False)
            self.qcmded_port = y.create_yarp_port(cfg.qcmded_portname, strict=False)
            self.qcmd_port = y.create_yarp_port(cfg.qcmd_portname, input_port=False)
            if sim:
                y.connect(self.qcmd_port, base + "/joint_sim", "/qvin")
                y.connect(self.qin_port, base + "/joint_sim", "/qout")
                if self.torso_joints:
                    print("Torso is another arm!")
                    remote = "/" + arm_type + "/" + cfg.torso_instance + "/joint_sim"
                    self.torso_qin_port = y.create_yarp_port(cfg.torso_qin_portname, strict=False)
                    y.connect(self.torso_qin_port, remote, "/qout", necessary=False)
                else:
                    print("This is the torso!")
            else:
                y.connect(self.qcmd_port, base + "/robot", "/cmd")
                y.connect(self.qcmded_port, base + "/robot", "/cmded")
                y.connect(self.qin_port, base + "/robot", "/pos")
        else:
            if not sim:
                raise NotImplementedError("the %s hardware driver (Player / remote_controlboard) is out of scope; use sim=True"
                                          % arm_type)
            # Powercube_Bridge.open / ICUB_Bridge.open, simulation branch (scripts/bridge:261-267,423-429)
            self.qin_port = y.create_yarp_port("/qin", strict=False)
            self.qcmd_port = y.create_yarp_port("/qcmd", input_port=False)
            y.connect(self.qcmd_port, base + "/joint_sim", "/qvin", necessary=False)
            y.connect(self.qin_port, base + "/joint_sim", "/qout", necessary=False)
        if self.kind == BRIDGE_ICUB:
            # scripts/bridge:343-355
            self.icub_torso_cjoints = list(getattr(cfg, "icub_torso_cjoints", [True, True, True]))
            self.icub_torso_num_joints = 3
            self.torso_cjoints_port = y.create_yarp_port("/torso_cjoints:i")
            self.control_weights_port = y.create_yarp_port("/control_weights:o", input_port=False)
            y.connect(self.control_weights_port, base + "/vectorField", "/weight", necessary=False)
        self.mixer = CommandMixer(self.cmd_ports, self.weight_port, self.nJoints, GUARD_TIME, list(DEFAULT_WEIGHTS),
                                  engine=runtime.engine)
        self.max_vel = float(cfg.max_vel)
        self.last_q = self.nJoints * [0.0]
        self.last_qcmded = []
        self.direct_control = False
        self.last_cmd = None

    # LWR_Bridge.read_pos (scripts/bridge:163-180); non-blocking here (single-threaded runtime)
    def read_pos(self):
        if self.kind == BRIDGE_ICUB:
            # Added at VVV09 (scripts/bridge:437-445): torso joint (de)activation
            bin_ = self.torso_cjoints_port.read(False)
            if bin_:
                print("Received a torso_joints bottle")
                self.set_torso_cjoints([bin_.get(i).asInt() for i in range(bin_.size())])
        bottle = self.qin_port.read(False)
        if bottle and bottle.size() == self.nJoints:
            self.last_q = [bottle.get(i).asDouble() for i in range(self.nJoints)]
        bottle = self.qcmded_port.read(False) if self.qcmded_port is not None else None
        if bottle is not None and bottle.size() == self.nJoints:
            self.last_qcmded = [bottle.get(i).asDouble() for i in range(self.nJoints)]
        elif self.last_qcmded == [] or self.sim:
            # no /cmded feedback (always the case in simulation): commanded == measured, so the command is qdot_lim
            self.last_qcmded = self.last_q
        if self.torso_qin_port is not None:
            # the torso belongs to the other arm's plant (scripts/bridge:174-178)
            bottle = self.torso_qin_port.read(False)
            if bottle:
                torso_q = [bottle.get(i).asDouble() for i in range(bottle.size())]
                for joint in self.torso_joints:
                    if joint < len(torso_q):
                        self.last_q[joint] = torso_q[joint]
        return self.last_q

    # ICUB_Bridge.set_torso_cjoints (scripts/bridge:470-506)
    def set_torso_cjoints(self, cjoints):
        num_weights = 10      # the reference hard-codes the iCub arm+torso size (scripts/bridge:473)
        weights = [1.0] * num_weights
        if len(cjoints) == self.icub_torso_num_joints:
            for i in range(self.icub_torso_num_joints):
                if cjoints[i] == 1:
                    self.icub_torso_cjoints[i] = True
                    weights[i] = 1.0
                else:
                    self.icub_torso_cjoints[i] = False
                    weights[i] = 0.0
            # tell the Vector Field about the new weights
            bout = self.control_weights_port.prepare()
            bout.clear()
            bout.addString("j")
            for w in weights:
                bout.addDouble(w)
            print("Sending weights: ")
            print(bout.toString())
            self.control_weights_port.write(True)
            return weights
        print("WARNING: set_torso_cjoints received a vector of controlled joints of a wrong length")
        return None

    # LWR_Bridge / Powercube_Bridge / ICUB_Bridge .set_vel (scripts/bridge:182-210,288-312,507-530): clamp(s) + command
    # forming run in the vfk_set_vel kernel, which follows params.bridge_kind
    def set_vel(self, qdot):
        if len(qdot) != self.nJoints or self.last_qcmded == []:
            print('got %d velocities, expected %d. Ignoring input' % (len(qdot), self.nJoints))
            return None
        import torch
        e = self.rt.engine
        n = self.nJoints
        dev = "cuda:%d" % e.device
        host = np.asarray([qdot, self.last_q, self.last_qcmded], dtype=e.np_dtype).reshape(3 * n, 1)
        dense = torch.from_numpy(host).to(dev)
        blocked = [e.alloc(n, 1) for _ in range(4)]
        for k in range(3):
            e.pack(dense[k * n:(k + 1) * n], blocked[k], n, 1, 1)
        e.set_vel(blocked[0], blocked[1], blocked[3], self.max_vel, self.direct_control, n, 1, q_cmded=blocked[2])
        out = torch.empty((n, 1), dtype=e.torch_dtype, device=dev)
        e.unpack(blocked[3], out, n, 1, 1)
        cmd = [float(v) for v in out[:, 0].cpu().numpy()]
        sendListPort(self.qcmd_port, cmd)
        self.last_cmd = cmd
        return cmd

    def update(self):
        """First half of the loop body (``scripts/bridge:601-623``): plant -> encoders, max_vel updates."""
        # if there is no controller, then avoid drifting (scripts/bridge:603-604)
        self.direct_control = all(w == 0 for w in self.mixer.weights)
        q = self.read_pos()
        sendListPort(self.encoders_port, q)
        botin = self.maxvel_port.read(False)
        if botin:
            new_max_vel = botin.get(0).asDouble()
            if self.rt.set_max_vel(new_max_vel, float(self.config.max_vel)):
                self.max_vel = new_max_vel
        return q

    def finish(self):
        """Second half (``scripts/bridge:625-627``): collect velocities from the mixer and write to the robot."""
        qdot = self.mixer.read()
        if tuple(self.mixer.weights) != tuple(self.rt.params.mixer_w):
            self.rt.set_params(mixer_w=tuple(self.mixer.weights))       # keep the fused kernel's mixer in step
        cmd = self.set_vel(qdot)
        sendListPort(self.current_weight_port, self.mixer.weights)
        return cmd

    def close(self):
        self.yarp_ctrl.close()


def main(argv=None):
    """``bridge -c <config> [-s] -n <namespace>`` (``scripts/vfclik:100-105``, ``scripts/bridge:58-66``)."""
    import sys
    from .module_cli import run_module

    def build(rt, opt, cfg):
        mods = [JointSim(cfg, opt.namespace)] if opt.sim else []
        return mods + [BridgeModule(rt, opt.namespace, sim=opt.sim)]
    sim_opt = (("-s", "--simulation"), dict(action="store_true", dest="sim", default=False, help="Simulation"))
    return run_module(sys.argv if argv is None else argv, build, extra_options=[sim_opt])


if __name__ == "__main__":
    import sys
    sys.exit(main())
