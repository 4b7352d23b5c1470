"""``vf`` -- the vector-field executor module with the reference's port contract (``scripts/vf``).

One ``VectorFieldModule.update()`` is one iteration of the reference's ``while not stop`` loop
(``scripts/vf:194-521``): speed-scale updates (``:197-207``), field add / remove messages
(``:210-293``), IK weight messages (``:296-309``), and -- when a joint vector arrived on ``/qIn`` --
the FK -> field -> IK chain (``:312-466``), whose arithmetic runs in the fused CUDA kernel through
``ControlRuntime.cycle``.  Ports: in ``/qIn /tool /param /weight /max_vel /pose_in``;
out ``/qdotOut /pose /pose_no_tool /vector_out /goal_out /track_error``.
"""
from __future__ import annotations

import numpy as np

from . import ports as yarp
from .ports import readListPort, sendListPort
from .runtime import ControlRuntime, dprint, field_query

MODULE_NAME = "/vectorField"


def get_weight_matrix(wbottle, n_vars):
    """``scripts/vf:164-179``: bottle ``['t'|'j', w...]`` -> diagonal weights (n_vars) or None if the size is wrong."""
    if wbottle.size() == n_vars + 1:
        return [wbottle.get(i + 1).asDouble() for i in range(n_vars)]
    dprint("WARNING: Wrong size of %s weights. Ignored" % wbottle.get(0).asString())
    return None


def pose12_to_list16(p):
    return [p[0], p[1], p[2], p[9], p[3], p[4], p[5], p[10], p[6], p[7], p[8], p[11], 0.0, 0.0, 0.0, 1.0]


def list16_to_pose12(l):
    return [l[0], l[1], l[2], l[4], l[5], l[6], l[8], l[9], l[10], l[3], l[7], l[11]]


class VectorFieldModule:
    def __init__(self, runtime: ControlRuntime, namespace: str = ""):
        self.rt = runtime
        cfg = runtime.config
        self.numJntsArm = cfg.nJoints
        base = cfg.robotarm_portbasename
        self.yarp_ctrl = yarp.ArcosYarp(ports_name_prefix=namespace, module_name_prefix=base + MODULE_NAME)
        y = self.yarp_ctrl
        self.qdotOutPort = y.create_yarp_port("/qdotOut", input_port=False)
        self.qInPort = y.create_yarp_port("/qIn", strict=False)
        self.toolPort = y.create_yarp_port("/tool", strict=False)
        self.paramPort = y.create_yarp_port("/param", strict=True)
        self.posePort = y.create_yarp_port("/pose", input_port=False)
        self.pose_no_tool_Port = y.create_yarp_port("/pose_no_tool", input_port=False)
        self.weightPort = y.create_yarp_port("/weight", strict=True)
        self.tracking_error_port = y.create_yarp_port("/track_error", input_port=False)
        self.maxvel_port = y.create_yarp_port("/max_vel", strict=False)
        self.pose_in_port = y.create_yarp_port("/pose_in", strict=False)
        self.vector_port = y.create_yarp_port("/vector_out", input_port=False)
        self.goal_port = y.create_yarp_port("/goal_out", input_port=False)
        y.connect(self.posePort, base + "/dmonitor", "/currentPosIn", necessary=False)
        y.connect(self.tracking_error_port, base + "/dmonitor", "/track_error_in", necessary=False)
        y.connect(self.qdotOutPort, base + "/bridge", "/vectorfieldcmd")
        self.oldtoolFrame = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]       # scripts/vf:154
        self.reporting_port_counter = 0
        self.last_qdot = None
        self.first_arm_data = True        # scripts/vf:183-184,335-338
        self.start_attractor = False

    # -- one loop iteration ------------------------------------------------------------------
    def update(self):
        rt = self.rt
        # speed scale (scripts/vf:197-207)
        botin = self.maxvel_port.read(False)
        if botin:
            rt.set_speed_scale(botin.get(0).asDouble())

        # field parameters (scripts/vf:210-293); strict port: drain everything that is queued
        parambottle = self.paramPort.read(False)
        if self.start_attractor:
            # scripts/vf:212-225: the iteration after the first joint data takes the "start_attractor" action instead of
            # the message it has just read from /param -- that message is consumed and lost (kept: bug-compatible; the
            # feeder re-sends every object with its next message)
            self.start_attractor = False
            parambottle = self.paramPort.read(False)
        while parambottle is not None:
            if parambottle.size() >= 1:
                action = parambottle.get(0).toString()
                if action == "add":
                    if parambottle.size() == 5:
                        force = parambottle.get(2).asDouble()
                        tVF = parambottle.get(3).asInt()
                        pl = parambottle.get(4).asList()
                        params = [pl.get(i).asDouble() for i in range(pl.size())]
                        rt.add_field(parambottle.get(1).asInt(), force, tVF, params)
                    else:
                        dprint("Wrong number of values, expected 5, ignoring")
                elif action == "remove":
                    if parambottle.size() == 2:
                        rt.remove_field(parambottle.get(1).asInt())
                    else:
                        dprint("Wrong number of values, expected 2")
            parambottle = self.paramPort.read(False)

        # IK weights (scripts/vf:296-309)
        weightbottle = self.weightPort.read(False)
        if weightbottle and weightbottle.size() >= 1:
            weight_type = weightbottle.get(0).asString()
            if weight_type == 't':
                w = get_weight_matrix(weightbottle, 6)
                if w is not None:
                    rt.set_params(w_task=tuple(w))
            elif weight_type == 'j':
                w = get_weight_matrix(weightbottle, self.numJntsArm)
                if w is not None:
                    rt.set_params(w_joint=tuple(w))

        # joint positions -> qdot (scripts/vf:312-466)
        qInbottle = self.qInPort.read(False)
        if qInbottle and qInbottle.size() == self.numJntsArm:
            q = [qInbottle.get(i).asDouble() for i in range(qInbottle.size())]
            toolFrame = readListPort(self.toolPort)
            if (not toolFrame) or (len(toolFrame) != 16):
                toolFrame = self.oldtoolFrame
            elif toolFrame != self.oldtoolFrame:
                self.oldtoolFrame = toolFrame
                rt.set_tool(toolFrame)
            out = rt.cycle(np.asarray(q))
            if self.first_arm_data:
                self.first_arm_data = False
                self.start_attractor = True
            pose16 = pose12_to_list16(out["pose"][:, 0])
            sendListPort(self.posePort, pose16)
            # flange frame = tool frame * tool^-1 (scripts/vf:329-342 publishes both)
            T = np.asarray(pose16).reshape(4, 4) @ np.linalg.inv(np.asarray(self.oldtoolFrame, dtype=np.float64).reshape(4, 4))
            sendListPort(self.pose_no_tool_Port, T.reshape(16).tolist())
            self.last_qdot = out["qdot_vf"][:, 0]
            mon = rt.monitor(advance=True)                                # scripts/vf:349-428, from the 6th frame on
            if mon["seen"] > 5:
                t = mon["track"][:, 0]
                b = self.tracking_error_port.prepare()
                b.clear()
                for v in t[:7]:
                    b.addDouble(float(v))
                b.addInt(int(t[7]))
                self.tracking_error_port.write()
            self.reporting_port_counter += 1
            if self.reporting_port_counter > 20:                      # scripts/vf:432-453
                self.reporting_port_counter = 0
                tw = field_query(rt, out["pose"])
                sendListPort(self.vector_port, tw[:, 0])
                if 1 in rt.vectorFields:
                    sendListPort(self.goal_port, rt.vectorFields[1][2])
            sendListPort(self.qdotOutPort, self.last_qdot)

        # field visualisation query (scripts/vf:469-503)
        pose_in_bottle = self.pose_in_port.read(False)
        if pose_in_bottle and pose_in_bottle.size() == 16:
            pose_in = [pose_in_bottle.get(i).asDouble() for i in range(16)]
            p12 = np.repeat(np.asarray(list16_to_pose12(pose_in))[:, None], rt.I, axis=1)
            sendListPort(self.vector_port, field_query(rt, p12)[:, 0])

    def close(self):
        self.yarp_ctrl.close()


def main(argv=None):
    """``vf -c <config> -n <namespace>`` (``scripts/vfclik:89``)."""
    import sys
    from .module_cli import run_module
    return run_module(sys.argv if argv is None else argv, lambda rt, opt, cfg: [VectorFieldModule(rt, opt.namespace)])


if __name__ == "__main__":
    import sys
    sys.exit(main())
