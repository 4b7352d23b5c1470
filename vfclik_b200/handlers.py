"""Client-side convenience API over the ports, with the reference's class surface (``src/handlers.py``).

``HandleArm`` (``:232-440``), ``HandleBridge`` (``:443-522``), ``HandleJController`` (``:525-576``) and ``HandleArmNew``
(``:32-230``) only write and read bottles; they contain no arithmetic.  The one difference: the reference's blocking
loops ``sleep(0.01)`` while other *processes* advance the robot; here everything lives in one process, so every handler
takes a ``spin`` callable that is invoked instead of sleeping (pass ``Vfclik.step`` to advance one control period).
"""
from __future__ import annotations

import time
from math import pi
from typing import Callable, Optional, Sequence

import numpy as np

from . import ports as yarp


def _default_spin():
    time.sleep(0.01)


def _connect(src, dst):
    yarp.Network.connect(src, dst)


def _write(port, data, strict=True):
    bottle = port.prepare()
    bottle.clear()
    for i in data:
        if isinstance(i, bool) or isinstance(i, int):
            bottle.addInt(int(i))
        elif isinstance(i, float):
            bottle.addDouble(i)
        elif isinstance(i, str):
            bottle.addString(i)
    port.writeStrict() if strict else port.write()


def _goal_distance(bottle):
    """``(dist_xyz, dist_rad)`` of object 0 in a ``/dmonitor/distOut`` bottle, or None."""
    for i in range(bottle.size()):
        line = bottle.get(i).asList()
        if line is not None and line.get(0).asInt() == 0:
            return line.get(1).asDouble(), line.get(2).asDouble() * pi / 180.0
    return None


class HandleArm(object):
    def __init__(self, arm_portbasename, namespace="", handlername="/HandlerArm", spin: Optional[Callable] = None):
        prename = namespace + arm_portbasename
        full_name = prename + handlername
        self.spin = spin or _default_spin
        self.outp = yarp.BufferedPortBottle(); self.outp.open(full_name + "/toObjectFeeder")
        self.stiffness_port = yarp.BufferedPortBottle(); self.stiffness_port.open(full_name + "/stiffness")
        self.goaldistp = yarp.BufferedPortBottle(); self.goaldistp.open(full_name + "/fromGoalDistance")
        self.posep = yarp.BufferedPortBottle(); self.posep.open(full_name + "/pose:i")
        self.toolp = yarp.BufferedPortBottle(); self.toolp.open(full_name + "/toToolin")
        _connect(full_name + "/toObjectFeeder", prename + "/ofeeder/object")
        _connect(full_name + "/stiffness", prename + "/robot/stiffness")
        _connect(prename + "/dmonitor/distOut", full_name + "/fromGoalDistance")
        _connect(prename + "/vectorField/pose", full_name + "/pose:i")
        _connect(full_name + "/toToolin", prename + "/vectorField/tool")
        self.current_frame = [1.0, 0.0, 0.0, 0.0, 0.0, 0.1, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0]   # src/handlers.py:261-264
        self.current_slowdown_distance = 0.1
        self.goal_threshold = 0.01

    def setTool(self, toolframe):
        """``toolframe``: 16 floats (row-major 4x4) or an object with ``to_list16()``."""
        vals = toolframe.to_list16() if hasattr(toolframe, "to_list16") else list(toolframe)
        _write(self.toolp, [float(v) for v in vals])

    def set_stiffness(self, stiffness):
        _write(self.stiffness_port, [float(v) for v in stiffness])

    def sendFrame(self):
        yarp.write_bottle_lists(self.outp, ["set", "goal", [float(v) for v in self.current_frame] + [self.current_slowdown_distance]],
                                strict=True)

    def gotoPos(self, pos):
        self.current_frame[3], self.current_frame[7], self.current_frame[11] = pos[0], pos[1], pos[2]
        self.sendFrame()

    def setOrient(self, orient):
        for i in range(3):
            for j in range(3):
                self.current_frame[j + 4 * i] = orient[j + i * 3]
        self.sendFrame()

    def getPose(self, blocking=True):
        pose_b = self.posep.read(False)
        while pose_b is None and blocking:
            self.spin()
            pose_b = self.posep.read(False)
        return None if pose_b is None else [pose_b.get(i).asDouble() for i in range(pose_b.size())]

    def gotoPose(self, pos, orient):
        self.current_frame[3], self.current_frame[7], self.current_frame[11] = pos[0], pos[1], pos[2]
        for i in range(3):
            for j in range(3):
                self.current_frame[j + 4 * i] = orient[j + i * 3]
        self.sendFrame()

    def gotoFrame(self, frame, wait=10.0, goal_precision=[]):
        """``src/handlers.py:346-387``: send the goal, then poll the goal distance until it is inside
        ``goal_precision = [trans, rot_rad]`` or ``wait`` seconds have passed.  Returns ``(result, difference)``."""
        for i in range(len(frame)):
            self.current_frame[i] = frame[i]
        self.sendFrame()
        init_time = cur_time = time.time()
        difference = np.array([0.0, 0.0])
        result = False
        while self.goaldistp.read(False) is not None:       # drop stale distance reports
            pass
        if len(goal_precision) == 2 and wait > 0.0:
            first_read = True
            while cur_time - init_time < wait:
                b = self.goaldistp.read(False)
                if b and not first_read:
                    d = _goal_distance(b)
                    if d is not None:
                        difference = np.array(d)
                        if d[0] < goal_precision[0] and d[1] < goal_precision[1]:
                            result = True
                            break
                if b:
                    first_read = False
                self.spin()
                cur_time = time.time()
        return (result, difference)

    def gotThere(self):
        b = self.goaldistp.read(False)
        dist = 1000.0
        if b:
            d = _goal_distance(b)
            if d is not None:
                dist = d[0]
            return dist < self.goal_threshold
        return False

    def gotoPosBlocking(self, pos, timeout=20):
        self.gotoPos(pos)
        startTime = time.time()
        for i in range(10):                                   # ignore the first reports
            self.spin()
            self.gotThere()
        while (time.time() - startTime) < timeout:
            self.spin()
            if self.gotThere():
                return True
        return False

    gotoPosBlockingGrasp = gotoPosBlocking


class HandleBridge(object):
    def __init__(self, arm_portbasename, handlername="HandlerArmBridge", torso=True, spin: Optional[Callable] = None):
        self.torso = torso
        self.spin = spin or _default_spin
        prename = arm_portbasename
        full_name = prename + "/" + handlername
        self.outp = yarp.BufferedPortBottle(); self.outp.open(full_name + "/toBridge_weights")
        if self.torso:
            self.torso_port = yarp.BufferedPortBottle(); self.torso_port.open(full_name + "/to_torso_cjoints")
            _connect(full_name + "/to_torso_cjoints", prename + "/bridge/torso_cjoints:i")
        self.VFW_port = yarp.BufferedPortBottle(); self.VFW_port.open(full_name + "/to_VF_weight:o")
        self.encoders_port = yarp.BufferedPortBottle(); self.encoders_port.open(full_name + "/encoders:i")
        # the reference connects to "/bridge/weights" (src/handlers.py:472) while the bridge opens "/bridge/weight"
        # (scripts/bridge:571); connect to both so the handler works against this bridge as intended
        _connect(full_name + "/toBridge_weights", prename + "/bridge/weights")
        _connect(full_name + "/toBridge_weights", prename + "/bridge/weight")
        _connect(full_name + "/to_VF_weight:o", prename + "/vectorField/weight")
        _connect(prename + "/bridge/encoders", full_name + "/encoders:i")

    def read_joint_angles(self):
        b = self.encoders_port.read(False)
        while b is None:
            self.spin()
            b = self.encoders_port.read(False)
        return [b.get(i).asDouble() for i in range(b.size())]

    def joint_controller(self):
        _write(self.outp, [0, 0, 1, 0])

    def cartesian_controller(self):
        _write(self.outp, [1, 1, 0, 0])

    def torso_joints(self, cjoints):
        if self.torso:
            _write(self.torso_port, [int(i) for i in cjoints])
        else:
            print("There's no torso")

    def set_VFW(self, type_of="joint", weights=[1] * 7):        # back compatibility
        print("deprecated, use set_weights instead")
        self.set_weights(type_of, weights)

    def set_weights(self, type_of="joint", weights=[1] * 7):
        _write(self.VFW_port, ['t' if type_of == "task" else 'j'] + [float(w) for w in weights])


class HandleJController(object):
    def __init__(self, arm_portbasename, handlername="HandlerArmJoint", spin: Optional[Callable] = None):
        self.spin = spin or _default_spin
        prename = arm_portbasename
        full_name = prename + "/" + handlername
        self.outp = yarp.BufferedPortBottle(); self.outp.open(full_name + "/to_js")
        self.inp = yarp.BufferedPortBottle(); self.inp.open(full_name + "/q")
        _connect(full_name + "/to_js", prename + "/jpctrl/ref")
        _connect(prename + "/bridge/encoders", full_name + "/q")

    def set_ref_js(self, js, wait=0.0, goal_precision=[]):
        """``src/handlers.py:544-576``: send the joint reference; optionally wait until every joint is within
        ``goal_precision`` of it (``wait`` seconds, -1 = forever).  Returns ``(result, difference)``."""
        js = np.asarray(js, dtype=np.float64)
        _write(self.outp, [float(v) for v in js])
        init_time = cur_time = time.time()
        difference = np.array([0.0] * len(js))
        result = False
        if len(goal_precision) == len(js) and wait != 0.0:
            gp = np.asarray(goal_precision, dtype=np.float64)
            while (cur_time - init_time < wait) or wait == -1:
                b = self.inp.read(False)
                if b:
                    q = np.array([b.get(i).asDouble() for i in range(b.size())])
                    difference = js - q
                    if (((js - gp) <= q) * ((js + gp) >= q)).all():
                        result = True
                        break
                self.spin()
                cur_time = time.time()
        return (result, difference)


class HandleArmNew:
    def __init__(self, namespace="/0", module_name="/handle_arm", arm_namespace="/0", robot="/lwr", arm="/right", sim=True,
                 spin: Optional[Callable] = None):
        self.sim = True
        self.spin = spin or _default_spin
        self.module_name, self.namespace, self.arm_namespace = module_name, namespace, arm_namespace
        arm_base = arm_namespace + robot + arm
        own = namespace + module_name + arm
        links = {            # attribute: (own port suffix, remote port, direction)
            "object_port": ("/object", arm_base + "/ofeeder/object", "out"),
            "stiffness_port": ("/stiffness", arm_base + "/robot/stiffness", "out"),
            "pose_port": ("/pose", arm_base + "/vectorField/pose", "in"),
            "distout_port": ("/distOut", arm_base + "/dmonitor/distOut", "in"),
            "tool_port": ("/tool", arm_base + "/vectorField/tool", "out"),
            "bridge_weight_port": ("/bridge/weight", arm_base + "/bridge/weight", "out"),
            "vf_weight_port": ("/vectorField/weight", arm_base + "/vectorField/weight", "out"),
            "bridge_encoders_port": ("/encoders", arm_base + "/bridge/encoders", "in"),
            "joint_ref_port": ("/joint_ref", arm_base + "/jpctrl/ref", "out"),
            "joint_sim_qin_port": ("/joint_sim/qin", arm_base + "/joint_sim/qin", "out"),
        }
        for attr, (suffix, remote, direction) in links.items():
            p = yarp.BufferedPortBottle()
            p.open(own + suffix)
            setattr(self, attr, p)
            _connect(own + suffix, remote) if direction == "out" else _connect(remote, own + suffix)
        self.current_slowdown_distance = 0.1
        self.cart_goal, self.joint_goal = None, None

    def _write_yarp_port(self, port, data, strict=True):
        _write(port, data, strict)

    def _read_blocking(self, port):
        b = port.read(False)
        while b is None:
            self.spin()
            b = port.read(False)
        return b

    def set_sim_arm_q(self, q):
        self._write_yarp_port(self.joint_sim_qin_port, [float(v) for v in q], strict=True)

    def go_cart(self, frame):
        self.cart_goal = frame
        yarp.write_bottle_lists(self.object_port, ["set", "goal", [float(v) for v in frame] + [self.current_slowdown_distance]],
                                strict=True)
        self.set_cartesian_control()

    def go_joint(self, angles):
        self.joint_goal = angles
        self._write_yarp_port(self.joint_ref_port, [float(v) for v in angles])
        self.set_joint_control()

    def get_cart_pose(self):
        b = self._read_blocking(self.pose_port)
        return [b.get(i).asDouble() for i in range(b.size())]

    def get_dist_cart_goal(self):
        while self.distout_port.read(False) is not None:
            pass
        while True:
            d = _goal_distance(self._read_blocking(self.distout_port))
            if d is not None:
                return [d[0], d[1]]

    def get_dist_joint_goal(self):
        b = self._read_blocking(self.bridge_encoders_port)
        cur = [b.get(i).asDouble() for i in range(b.size())]
        return [i - j for i, j in zip(self.joint_goal, cur)]

    def set_controller_mixer(self, cart=True, joint=False, null=False):
        self._write_yarp_port(self.bridge_weight_port, [1 if cart else 0, 1 if null else 0, 1 if joint else 0, 0])

    def set_cartesian_control(self):
        self.set_controller_mixer(cart=True, null=True)

    def set_joint_control(self):
        self.set_controller_mixer(cart=False, null=False, joint=True)

    def set_wik_joint_weights(self, joint_weights):
        _write(self.vf_weight_port, ["j"] + [float(w) for w in joint_weights])

    def set_wik_cart_weights(self, cart_weights):
        _write(self.vf_weight_port, ["t"] + [float(w) for w in cart_weights])

    def set_tool(self, tool_frame):
        self._write_yarp_port(self.tool_port, [float(v) for v in tool_frame])
