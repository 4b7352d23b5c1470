"""``nullspace`` module with the reference's port contract (``scripts/nullspace``).

Ports: ``/qin`` (joint positions from ``/bridge/encoders``), ``/control`` ("four-float-bottles",
sticky, ``:169-173``), ``/qdotout`` -> ``/bridge/nullcmd``.  The arithmetic -- Jacobian, projector
``I - pinv(J) J`` (``:75-79``), basis with sign continuity (``:91-107``), motion along it
(``:110-117``), the all-or-nothing lookahead limit check (``:120-131``) and the 0.5 gain
(``:62,183``) -- runs inside the fused CUDA kernel; this class only moves bottles.

Two modes (``vfk_params.ns_mode``): ``NS_CONTROL`` is the reference's 4-float control interface
(any joint count: the four floats mix ``min(4, nJoints - 6)`` sign-continuous basis vectors of null(J), undamped like the
reference); ``NS_PROJECTOR`` is north_star's joint-limit avoidance ``(I - J^+ J) qdot0``.
"""
from __future__ import annotations

import numpy as np

from . import ports as yarp
from .ports import sendListPort
from .runtime import ControlRuntime

MODULE_NAME = "/nullspace"
gain = 0.5                  # scripts/nullspace:62


class NullspaceModule:
    def __init__(self, runtime: ControlRuntime, namespace: str = ""):
        self.rt = runtime
        cfg = runtime.config
        base = cfg.robotarm_portbasename
        self.nJoints = cfg.nJoints
        self.yarp_ctrl = yarp.ArcosYarp(ports_name_prefix=namespace, module_name_prefix=base + MODULE_NAME)
        self.qin_port = self.yarp_ctrl.create_yarp_port("/qin", strict=False)
        self.qdotout_port = self.yarp_ctrl.create_yarp_port("/qdotout", input_port=False)
        self.control_port = self.yarp_ctrl.create_yarp_port("/control", strict=False)     # expects four-float-bottles
        self.yarp_ctrl.connect(self.qdotout_port, base + "/bridge", "/nullcmd")
        self.yarp_ctrl.connect(self.qin_port, base + "/bridge", "/encoders")
        self.control = [0] * 4                                                             # scripts/nullspace:137
        self.last_qdot = None

    def update(self) -> bool:
        """One iteration of ``scripts/nullspace:159-187``; returns False when there was no joint data."""
        bin_ = self.qin_port.read(False)
        if not bin_:
            return False
        q = [bin_.get(i).asDouble() for i in range(bin_.size())]
        bcontrol = self.control_port.read(False)
        if bcontrol is not None:
            self.control = [bcontrol.get(i).asDouble() for i in range(bcontrol.size())]
            self.rt.set_ns_control(self.control)
        if len(q) != self.nJoints:
            return False
        out = self.rt.cycle(np.asarray(q))
        self.last_qdot = out["qdot_ns"][:, 0]
        sendListPort(self.qdotout_port, self.last_qdot)            # gain already applied in the kernel
        return True

    def close(self):
        self.yarp_ctrl.close()


def main(argv=None):
    """``nullspace -c <config> -n <namespace>`` (``scripts/vfclik:96-98``)."""
    import sys
    from .module_cli import run_module
    return run_module(sys.argv if argv is None else argv, lambda rt, opt, cfg: [NullspaceModule(rt, opt.namespace)])


if __name__ == "__main__":
    import sys
    sys.exit(main())
