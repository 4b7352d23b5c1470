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
            self._ref_sent = list(self.ref)
        inbottle = self.inPort.read(False)
        if not inbottle or inbottle.size() != self.nJoints:
            return False
        q = [inbottle.get(i).asDouble() for i in range(self.nJoints)]
        self._apply_position_dependent_limits(q)
        out = self.rt.cycle(np.asarray(q))
        sendListPort(self.outPort, out["qdot_jp"][:, 0])
        self.at_goal = 1 if (int(out["flags"][0]) & FLAG_AT_GOAL) else 0
        b = self.atGoalPort.prepare()
        b.clear()
        b.addInt(self.at_goal)
        self.atGoalPort.write(True)
        return True

    def _apply_position_dependent_limits(self, q):
        """``check_limits`` (``scripts/joint_p_controller:79-89``) clamps the reference into ``config.updateJntLimits(q)``,
        which may depend on the current posture (iCub shoulder coupling).  The kernel clamps into the chain's static
        limits; when the config's hook returns anything tighter for this q, the reference handed to the kernel is
        clamped here first (a user-supplied Python function cannot run on the device), with the reference's messages."""
        hook = getattr(self.rt.config, "updateJntLimits", None)
        if hook is None:
            return
        limits = hook(q)
        ref_out = list(self.ref)
        for i in range(min(len(limits), self.nJoints)):
            if self.ref[i] < limits[i][0]:
                print("Limiting low", i)
                ref_out[i] = limits[i][0]
            elif self.ref[i] > limits[i][1]:
                print("Limiting high", i)
                ref_out[i] = limits[i][1]
        if ref_out != getattr(self, "_ref_sent", self.ref):
            self.rt.set_jp_ref(ref_out)
        self._ref_sent = ref_out
        self.ref = ref_out            # the reference's loop overwrites `ref` with the clamped value (:121): the clamp persists

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
