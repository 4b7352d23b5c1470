"""``joint_p_controller`` module with the reference's port contract (``scripts/joint_p_controller``).

Ports: ``/in`` (joint positions), ``/ref`` (joint reference; starts at ``config.initial_joint_pos``,
``:92``), ``/out`` -> ``/bridge/jointcmd``, ``/at_goal`` (1 int).  The clamp of the reference to the joint
limits (``:79-89``), ``qdot = kp (ref - q)`` (``:126-127``) and the signed at-goal test against
``delta = 0.087`` (``:135-146``) run in the fused CUDA kernel.
"""
from __future__ import annotations

import numpy as np

from . import ports as yarp
from ._lib import FLAG_AT_GOAL
from .ports import sendListPort
from .runtime import ControlRuntime

MODULE_NAME = "/jpctrl"
delta = 0.087               # scripts/joint_p_controller:57


class JointPControllerModule:
    def __init__(self, runtime: ControlRuntime, namespace: str = ""):
        self.rt = runtime
        cfg = runtime.config
        base = cfg.robotarm_portbasename
        self.nJoints = cfg.nJoints
        self.yarp_ctrl = yarp.ArcosYarp(ports_name_prefix=namespace, module_name_prefix=base + MODULE_NAME)
        self.inPort = self.yarp_ctrl.create_yarp_port("/in", strict=False)
        self.refPort = self.yarp_ctrl.create_yarp_port("/ref", strict=False)
        self.outPort = self.yarp_ctrl.create_yarp_port("/out", input_port=False)
        self.atGoalPort = self.yarp_ctrl.create_yarp_port("/at_goal", input_port=False)
        self.yarp_ctrl.connect(self.outPort, base + "/bridge", "/jointcmd")
        self.yarp_ctrl.connect(self.inPort, base + "/bridge", "/encoders")
        self.ref = list(cfg.initial_joint_pos)
        runtime.set_jp_ref(self.ref)
        self.at_goal = 0

    def update(self) -> bool:
        refbottle = self.refPort.read(False)
        if refbottle and refbottle.size() == self.nJoints:
            self.ref = [refbottle.get(i).asDouble() for i in range(self.nJoints)]
            self.rt.set_jp_ref(self.ref)
        inbottle = self.inPort.read(False)
        if not inbottle or inbottle.size() != self.nJoints:
            return False
        q = [inbottle.get(i).asDouble() for i in range(self.nJoints)]
        out = self.rt.cycle(np.asarray(q))
        sendListPort(self.outPort, out["qdot_jp"][:, 0])
        self.at_goal = 1 if (int(out["flags"][0]) & FLAG_AT_GOAL) else 0
        b = self.atGoalPort.prepare()
        b.clear()
        b.addInt(self.at_goal)
        self.atGoalPort.write(True)
        return True

    def close(self):
        self.yarp_ctrl.close()


def main(argv=None):
    """``joint_p_controller -c <config> -n <namespace>`` (``scripts/vfclik:92``)."""
    import sys
    from .module_cli import run_module
    return run_module(sys.argv if argv is None else argv, lambda rt, opt, cfg: [JointPControllerModule(rt, opt.namespace)])


if __name__ == "__main__":
    import sys
    sys.exit(main())
