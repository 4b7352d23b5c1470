"""``vfclik`` launcher with the reference's command line (``scripts/vfclik:27-114``).

The reference spawns one OS process per module through ``PManager``; here the same modules are objects in one
process over one ``ControlRuntime`` (one fused CUDA kernel launch per control period).  Flags kept:
``-r/--robot`` (lwr), ``-i/--instance`` (right), ``-n/--namespace`` (/0), ``-d/--config_dir``, ``-s/--simulation``,
``--no_nullspace``; the config file is ``<config_dir>config-<robot>-<instance>.py`` and a missing file exits -1.
Extra (batched operation): ``--cycles`` bounds the loop, ``--precision`` picks the kernel arithmetic.
"""
from __future__ import annotations

import optparse
import os
import sys
import time

from . import ports as yarp
from .config import PACKAGE_CONFIG_DIR, config_filename, load_config


def build_parser():
    parser = optparse.OptionParser("usage: %prog [options]")
    parser.add_option("-r", "--robot", dest="robot", default="lwr", type="string", help="robot name")
    parser.add_option("-i", "--instance", dest="instance", default="right", type="string", help="instance")
    parser.add_option("-n", "--namespace", dest="namespace", default="/0", type="string", help="ARCOS-Lab yarp basename")
    parser.add_option("-d", "--config_dir", dest="config_dir", default="../config_data/lwr/", type="string",
                      help="config data directory")
    parser.add_option("-s", "--simulation", action="store_true", dest="sim", default=False, help="Simulation")
    parser.add_option("--no_nullspace", action="store_true", dest="no_nullspace", default=False, help="No nullspace module")
    parser.add_option("--cycles", dest="cycles", default=0, type="int", help="stop after this many control cycles (0 = run)")
    parser.add_option("--precision", dest="precision", default=64, type="int", help="kernel arithmetic: 32 or 64")
    parser.add_option("--no_sleep", action="store_true", dest="no_sleep", default=False, help="do not pace the loop at config.rate")
    return parser


class Vfclik:
    """All modules of one arm wired together (``scripts/vfclik:88-105``)."""

    def __init__(self, config, namespace="/0", sim=True, no_nullspace=False, precision=64, device=0):
        from .bridge import BridgeModule, JointSim
        from .joint_p_controller import JointPControllerModule
        from .monitor_distance import MonitorDistanceModule
        from .nullspace import NullspaceModule
        from .object_feeder import ObjectFeederModule
        from .runtime import ControlRuntime, NS_OFF, NS_PROJECTOR
        from .vf import VectorFieldModule
        self.config = config
        self.runtime = ControlRuntime(config, n_instances=1, precision=precision, device=device)
        self.runtime.set_params(ns_mode=NS_OFF if no_nullspace else NS_PROJECTOR)
        self.vf = VectorFieldModule(self.runtime, namespace)
        self.ofeeder = ObjectFeederModule(config, namespace)
        self.dmonitor = MonitorDistanceModule(self.runtime, namespace)
        self.jpctrl = JointPControllerModule(self.runtime, namespace)
        self.nullspace = None if no_nullspace else NullspaceModule(self.runtime, namespace)
        self.joint_sim = JointSim(config, namespace) if sim else None
        self.bridge = BridgeModule(self.runtime, namespace, sim=sim)
        self.namespace = namespace
        self.cycles = 0
        # the first goal comes from object_feeder's own first message (scripts/object_feeder:102-108): config.initial_vf_pose

    def _param_port(self):
        return self.vf.paramPort

    def set_goal(self, params17):
        """``set goal (16 [+ slowdown])`` as object_feeder forwards it (``scripts/object_feeder:122-135,229-247``)."""
        p = [float(x) for x in params17]
        if len(p) == 16:
            p.append(0.03)
        self.runtime.add_field(1, 1.0, 1, p)
        self.runtime.remove_field(2)
        self.runtime.remove_field(3)

    def set_obstacle_p(self, n, frame16, radius, order):
        """``set ObstacleP n (16 + radius + order)`` (``scripts/object_feeder:156-170,317-334``)."""
        f = [float(x) for x in frame16]
        self.runtime.add_field(4 + (n + 1), -10.0, 2, [f[3], f[7], f[11], float(radius), 0.001, float(order)])

    def remove_obstacle(self, n):
        self.runtime.remove_field(4 + (n + 1))

    def step(self):
        """One control period: plant -> bridge -> controllers -> mixer -> command."""
        if self.joint_sim is not None:
            self.joint_sim.update()
        self.bridge.update()
        while self.ofeeder.update():
            pass
        self.vf.update()
        if self.nullspace is not None:
            self.nullspace.update()
        self.jpctrl.update()
        self.dmonitor.update()
        cmd = self.bridge.finish()
        self.cycles += 1
        return cmd

    def close(self):
        for m in (self.vf, self.ofeeder, self.dmonitor, self.jpctrl, self.nullspace, self.joint_sim, self.bridge):
            if m is not None:
                m.close()
        self.runtime.close()


def main(argv=None):
    parser = build_parser()
    (options, args) = parser.parse_args(sys.argv[1:] if argv is None else argv)
    filename = config_filename(options.config_dir, options.robot, options.instance)
    print("Config filename: ", filename)
    if not os.path.exists(filename):
        alt = config_filename(os.path.join(PACKAGE_CONFIG_DIR, options.robot) + "/", options.robot, options.instance)
        if options.config_dir == "../config_data/lwr/" and os.path.exists(alt):
            filename = alt              # the reference's default directory is external; fall back to the packaged config
            print("Config filename: ", filename)
        else:
            print("Config filename: ", filename, " not found, exiting")
            sys.exit(-1)
    config = load_config(filename)
    yarp.Network.init()
    app = Vfclik(config, namespace=options.namespace, sim=options.sim, no_nullspace=options.no_nullspace,
                 precision=options.precision)
    try:
        while options.cycles == 0 or app.cycles < options.cycles:
            t0 = time.time()
            app.step()
            if not options.no_sleep:
                left = float(config.rate) - (time.time() - t0)
                if left > 0:
                    time.sleep(left)
    except KeyboardInterrupt:
        pass
    finally:
        q = app.bridge.last_q
        app.close()
    print("cycles:", app.cycles, "q:", [round(v, 6) for v in q])
    return 0


if __name__ == "__main__":
    sys.exit(main())
