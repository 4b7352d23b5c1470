"""``monitor_distance`` module with the reference's port contract (``scripts/monitor_distance``).

Ports: ``/currentPosIn`` (tool pose from ``/vectorField/pose``), ``/objectsIn`` (``add`` / ``remove`` object frames from
``/ofeeder/objectOut``), ``/track_error_in``; out ``/distOut`` (list of ``(id, dist_xyz, dist_deg)``, ``:160-172``) and
``/tracking_state`` (``("xyz" | "rot", state)`` on change, ``:196-219``).  Distances, the tracking-error diagnostics of
``scripts/vf:349-428`` and the 20-sample majority vote run in the ``vfk_monitor`` CUDA kernel on the runtime's resident
pose / twist / goal; this class moves bottles and remembers the last published state.
"""
from __future__ import annotations

import numpy as np

from . import ports as yarp
from .runtime import ControlRuntime

MODULE_NAME = "/dmonitor"
STATES = ["on goal", "follow", "not follow"]


class MonitorDistanceModule:
    def __init__(self, runtime: ControlRuntime, namespace: str = ""):
        self.rt = runtime
        base = runtime.config.robotarm_portbasename
        self.yarp_ctrl = yarp.ArcosYarp(ports_name_prefix=namespace, module_name_prefix=base + MODULE_NAME)
        y = self.yarp_ctrl
        self.currentPosIn = y.create_yarp_port("/currentPosIn", strict=False)
        self.objectsInPort = y.create_yarp_port("/objectsIn")
        self.distOutPort = y.create_yarp_port("/distOut", input_port=False)
        self.track_error_in_port = y.create_yarp_port("/track_error_in", strict=False)
        self.tracking_state_port = y.create_yarp_port("/tracking_state", input_port=False)
        self.objects = {}
        self.last_tracking_xyz_state = "on goal"
        self.last_tracking_rot_state = "on goal"
        self.last = None
        self._mon = None
        self._seen_cycle = -1

    def _device_state(self):
        if self._mon is None:
            import torch
            e = self.rt.engine
            self._mon = dict(f=e.alloc(32, self.rt.I), i=e.alloc(6, self.rt.I, dtype=torch.int32), track=e.alloc(8, self.rt.I),
                             dist=e.alloc(2, self.rt.I), state=e.alloc(2, self.rt.I, dtype=torch.int32),
                             scratch_f=e.alloc(32, self.rt.I), scratch_i=e.alloc(6, self.rt.I, dtype=torch.int32),
                             obj=e.alloc(13, self.rt.I), odist=e.alloc(2, self.rt.I))
        return self._mon

    def _unblock(self, t):
        import torch
        comps = t.shape[1]
        if t.dtype == torch.int32:
            return t.cpu().numpy().transpose(1, 0, 2).reshape(comps, -1)[:, :self.rt.I]
        d = torch.empty((comps, self.rt.I), dtype=t.dtype, device=t.device)
        self.rt.engine.unpack(t, d, comps, 1, self.rt.I)
        return d.cpu().numpy()

    def update(self) -> bool:
        """One iteration of ``scripts/monitor_distance:107-221``."""
        import torch
        from .runtime import _session_device_view
        ob = self.objectsInPort.read(False)
        while ob is not None:
            if ob.size() >= 2:
                if ob.get(0).toString() == "add" and ob.size() == 3:
                    lst = ob.get(2).asList()
                    if lst is not None and lst.size() == 16:
                        self.objects[ob.get(1).asInt()] = [lst.get(i).asDouble() for i in range(16)]
                if ob.get(0).toString() == "remove":
                    self.objects.pop(ob.get(1).asInt(), None)
            ob = self.objectsInPort.read(False)
        pose_b = self.currentPosIn.read(False)
        if not pose_b or pose_b.size() != 16 or len(self.objects) == 0 or self.rt.cycles == self._seen_cycle:
            return False
        self._seen_cycle = self.rt.cycles
        e, m = self.rt.engine, self._device_state()
        b = _session_device_view(self.rt)
        shared = self.rt.monitor()                    # one vfk_monitor launch per cycle, shared with vf's /track_error
        dist, state, track = shared["dist"], shared["state"], shared["track"]
        out = self.distOutPort.prepare()
        out.clear()
        entries = []
        for oid in self.objects:
            if oid == 0:
                dxyz, ddeg = float(dist[0, 0]), float(dist[1, 0])
            else:                                   # any other object: same kernel, that object's frame as the target
                f = self.objects[oid]
                g = np.zeros((13, self.rt.I), dtype=e.np_dtype)
                g[0:9, 0] = [f[0], f[1], f[2], f[4], f[5], f[6], f[8], f[9], f[10]]
                g[9:12, 0] = [f[3], f[7], f[11]]
                dense = torch.from_numpy(g).to(m["obj"].device)
                e.pack(dense, m["obj"], 13, 1, self.rt.I)
                m["scratch_i"].zero_()
                e.monitor(int(b.pose), int(b.twist), m["obj"], m["scratch_f"], m["scratch_i"], self.rt.I, dist_out=m["odist"])
                od = self._unblock(m["odist"])
                dxyz, ddeg = float(od[0, 0]), float(od[1, 0])
            lst = out.addList()
            lst.addDouble(oid); lst.addDouble(dxyz); lst.addDouble(ddeg)
            entries.append((oid, dxyz, ddeg))
        self.distOutPort.write()
        sx, sr = int(state[0, 0]), int(state[1, 0])
        if sx >= 0 and STATES[sx] != self.last_tracking_xyz_state:
            self.last_tracking_xyz_state = STATES[sx]
            yarp.write_bottle_lists(self.tracking_state_port, ["xyz", STATES[sx]], strict=True)
        if sr >= 0 and STATES[sr] != self.last_tracking_rot_state:
            self.last_tracking_rot_state = STATES[sr]
            # the reference writes tracking_xyz_state under the "rot" tag (scripts/monitor_distance:215): kept, bug-compatible
            yarp.write_bottle_lists(self.tracking_state_port, ["rot", STATES[sx] if sx >= 0 else STATES[sr]], strict=True)
        self.last = dict(dist=entries, track_error=track[:, 0].tolist(), xyz_state=sx, rot_state=sr)
        return True

    def close(self):
        self.yarp_ctrl.close()


def main(argv=None):
    """``monitor_distance -c <config> -n <namespace>`` (``scripts/vfclik:91``)."""
    import sys
    from .module_cli import run_module
    return run_module(sys.argv if argv is None else argv, lambda rt, opt, cfg: [MonitorDistanceModule(rt, opt.namespace)])


if __name__ == "__main__":
    import sys
    sys.exit(main())
