"""``object_feeder`` module with the reference's port contract (``scripts/object_feeder``).

Translates the user-level scene messages on ``/ofeeder/object`` --
``("set", "goal", (16 [+ slowdown]))``, ``("set", "goalAndNormal", (21 | 22))``,
``("set", "ObstacleP", n, (16 + radius + order))``, ``("set", "ObstacleH", n, (16 + normal(3) + safe + order))``,
``("remove", n)`` (``:111-212``) -- into the ``add`` / ``remove`` field messages ``vf`` consumes on ``/param``
(ids, forces, vfl types and parameter layouts of ``:214-354``), echoes every message on ``/objectf`` and forwards the
object frames to the distance monitor on ``/objectOut``.  Pure protocol translation: no arithmetic beyond the
near-goal repeller's position ``goal + axis/|axis| * 0.05`` (``:281-303``).
"""
from __future__ import annotations

import math

from . import ports as yarp
from .ports import write_bottle_lists
from .runtime import dprint

MODULE_NAME = "/ofeeder"
DEFAULT_SLOWDOWN = 0.03          # scripts/object_feeder:131,150
GOAL_TO_VOBS_D = 0.05            # scripts/object_feeder:283


class ObjectFeederModule:
    def __init__(self, config, namespace: str = ""):
        base = config.robotarm_portbasename
        self.config = config
        self.yarp_ctrl = yarp.ArcosYarp(ports_name_prefix=namespace, module_name_prefix=base + MODULE_NAME)
        y = self.yarp_ctrl
        self.objectPort = y.create_yarp_port("/object", strict=True)
        self.paramPort = y.create_yarp_port("/param", input_port=False)
        self.object_f_port = y.create_yarp_port("/objectf", input_port=False)
        self.objectOutPort = y.create_yarp_port("/objectOut", input_port=False)
        y.connect(self.paramPort, base + "/vectorField", "/param")
        y.connect(self.objectOutPort, base + "/dmonitor", "/objectsIn", necessary=False)
        self.objects = {}                                   # scripts/object_feeder:89
        self.init_pose = hasattr(config, "initial_vf_pose")

    # ---------------------------------------------------------------------------------------
    def update(self) -> bool:
        """One iteration of ``scripts/object_feeder:93-359``; returns False when there was nothing to do."""
        objectbottle = self.objectPort.read(False)
        if not objectbottle and not self.init_pose:
            return False
        if self.init_pose:
            dprint("Setting first goal")
            objectbottle = yarp.Bottle.from_list(self.config.initial_vf_pose)
            self.init_pose = False
        # echo (scripts/object_feeder:105-109)
        fb = self.object_f_port.prepare()
        fb.clear()
        for i in range(objectbottle.size()):
            fb.add(objectbottle.get(i))
        self.object_f_port.writeStrict()
        if objectbottle.size() < 2:
            return True
        action = objectbottle.get(0).toString()
        if action == "set":
            kind = objectbottle.get(1).toString()
            if objectbottle.size() == 3:
                pl = objectbottle.get(2).asList()
                params = [pl.get(i).asDouble() for i in range(pl.size())] if pl is not None else []
                if kind == "goal":
                    if len(params) == 16:
                        params.append(DEFAULT_SLOWDOWN)
                    if len(params) == 17:
                        self.objects[0] = params
                    else:
                        dprint("Wrong number of values, expected 16")
                if kind == "goalAndNormal":
                    if len(params) == 21:
                        params.append(DEFAULT_SLOWDOWN)
                    if len(params) == 22:
                        self.objects[0] = params
                    else:
                        dprint("Wrong number of values, expected 21")
            elif objectbottle.size() == 4:
                pl = objectbottle.get(3).asList()
                params = [pl.get(i).asDouble() for i in range(pl.size())] if pl is not None else []
                if kind == "ObstacleP":
                    if len(params) == 18:
                        self.objects[objectbottle.get(2).asInt() + 1] = ['ObstacleP'] + params
                    else:
                        dprint("Wrong number of values, expected 18")
                if kind == "ObstacleH":
                    if len(params) == 21:
                        self.objects[objectbottle.get(2).asInt() + 1] = ['ObstacleH'] + params
                    else:
                        dprint("Wrong number of values, expected 21")
            else:
                dprint("Wrong number of values, expected 3 or 4")
        elif action == "remove":
            objectNum = objectbottle.get(1).asInt()
            if objectNum + 1 in self.objects:
                del self.objects[objectNum + 1]
                write_bottle_lists(self.paramPort, ["remove", 5 + objectNum], strict=True)
                write_bottle_lists(self.objectOutPort, ["remove", objectNum + 1], strict=True)
            else:
                dprint("Object doesn't exist, doing nothing")
                return True
        else:
            dprint("Action not recognized")
        self._send_all()
        return True

    def _add(self, field_id, force, vtype, params):
        write_bottle_lists(self.paramPort, ["add", int(field_id), float(force), int(vtype), [float(p) for p in params]],
                           strict=True)

    def _send_all(self):
        """``scripts/object_feeder:214-359``: re-send every object as field messages once a goal exists."""
        if 0 not in self.objects:
            dprint("Not setting repellers, waiting for a goal. Please set a goal, to apply all the repellers")
            return
        for objectNum in sorted(self.objects):      # Python 2 dicts iterate small int keys in ascending order: goal first
            pl = self.objects[objectNum]
            if objectNum == 0:
                write_bottle_lists(self.objectOutPort, ["add", objectNum, [float(v) for v in pl[:16]]], strict=True)
                if len(pl) == 17:                                        # normal goal
                    self._add(1, 1, 1, pl[:17])
                    for i in range(2):
                        write_bottle_lists(self.paramPort, ["remove", i + 2], strict=True)
                elif len(pl) == 22:                                      # goal with approach vector
                    dprint("sending goal with approach")
                    self._add(1, 1, 1, list(pl[:16]) + [pl[21]])
                    goal = [pl[3], pl[7], pl[11]]
                    self._add(2, 30, 5, goal + list(pl[16:19]) + [pl[19], 10, pl[20], 2])        # funnel attractor
                    n = math.sqrt(sum(v * v for v in pl[16:19]))
                    if n == 0.0:
                        dprint("goalAndNormal with a zero approach axis, near-goal repeller skipped")
                        continue
                    vob = [goal[k] + pl[16 + k] / n * GOAL_TO_VOBS_D for k in range(3)]
                    self._add(3, -10, 2, vob + [pl[20] + GOAL_TO_VOBS_D, 0.001, 5])              # near-goal repeller
                else:
                    dprint("wrong number of parameters")
            else:
                write_bottle_lists(self.objectOutPort, ["add", objectNum, [float(v) for v in pl[1:17]]], strict=True)
                if pl[0] == "ObstacleP":
                    dprint("Sending point obstacle")
                    self._add(4 + objectNum, -10, 2, [pl[4], pl[8], pl[12], pl[17], 0.001, pl[18]])
                if pl[0] == "ObstacleH":
                    dprint("Sending table repeller")
                    self._add(4 + objectNum, -50, 4, [pl[4], pl[8], pl[12], pl[17], pl[18], pl[19], pl[20], pl[21]])

    def close(self):
        self.yarp_ctrl.close()


def main(argv=None):
    """``object_feeder -c <config> -n <namespace>`` (``scripts/vfclik:90``)."""
    import sys
    from .module_cli import run_module
    return run_module(sys.argv if argv is None else argv, lambda rt, opt, cfg: [ObjectFeederModule(cfg, opt.namespace)],
                      needs_runtime=False)


if __name__ == "__main__":
    import sys
    sys.exit(main())
