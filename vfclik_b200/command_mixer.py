"""``CommandMixer`` with the reference's class interface (``src/command_mixer.py:32-82``).

Port handling (weight updates, latest-value reads, guard-time zeroing, the wrong-length and NaN
reports) stays on the host exactly as the reference orders it; the weighted sum itself
(``:78-82``) is the ``vfk_mix`` CUDA kernel.  Inside the fused control cycle the same sum is
computed in registers (``vfk_cycle_kernel`` step 8); this class exists so that code written
against the reference's mixer keeps working, including its stand-alone ``main`` use.

Port objects need only ``read(False) -> bottle | None``; bottles ``size()`` and
``get(i).asDouble()`` -- the contract of SURVEY.md section 8b.
"""
from __future__ import annotations

import math
import time
from typing import List, Optional, Sequence

import numpy as np

basename = '/mixer'

_UTIL_ENGINE = None


def _utility_engine(precision: int = 64):
    """A process-wide engine for the chain-independent kernels (mixer sum, layout conversion)."""
    global _UTIL_ENGINE
    if _UTIL_ENGINE is None:
        from . import kdl
        from .config import chain_from_segments
        from .engine import Engine
        segs = [kdl.Segment(kdl.Joint(kdl.Joint.RotZ), kdl.Frame(kdl.Vector(0, 0, 0.1))) for _ in range(6)]
        _UTIL_ENGINE = Engine(chain_from_segments(segs, [[-1.0, 1.0]] * 6), precision=precision)
    return _UTIL_ENGINE


class CommandMixer:
    def __init__(self, ports: Sequence, weight_port, n: int, guard_time: float, weights: List[float], engine=None):
        self.nChannels = n
        self.ports = ports
        self.weight_port = weight_port
        if len(ports) != len(weights):
            print('wrong number of initial weights. Resetting to zeros.')
            self.weights = [0.0] * len(ports)
        else:
            self.weights = weights
        self.guard_time = guard_time
        self.last_command = [[0.0] * n] * len(self.ports)
        self.last_command_time = [time.time()] * len(self.ports)
        self._engine = engine
        self._dev = None

    # -- device side -------------------------------------------------------------------------
    def _mix_on_device(self) -> List[float]:
        """One ``vfk_mix`` launch: the single instance's tile-blocked arrays (element (c, 0) at ``c * 32``) are laid out on
        the host, so the call is one H2D copy, one kernel, one D2H copy."""
        import torch
        e = self._engine or _utility_engine()
        P, n = len(self.ports), self.nChannels
        if P > 8:
            raise ValueError("vfk_mix takes at most 8 command ports, got %d" % P)
        if self._dev is None or self._dev[0] is not e:
            dev = "cuda:%d" % e.device
            self._dev = (e, np.zeros((P, n, 32), dtype=e.np_dtype), torch.zeros((P, n, 32), dtype=e.torch_dtype, device=dev),
                         torch.zeros((n, 32), dtype=e.torch_dtype, device=dev))
        _, host, blocked, out_b = self._dev
        host[:, :, 0] = np.asarray(self.last_command, dtype=e.np_dtype)
        blocked.copy_(torch.from_numpy(host))
        e.mix([blocked[p] for p in range(P)], self.weights, out_b, n, 1)
        return [float(v) for v in out_b[:, 0].cpu().numpy()]

    # -- reference interface -----------------------------------------------------------------
    def read(self) -> List[float]:
        # update weights (src/command_mixer.py:48-53)
        if self.weight_port:
            bottle = self.weight_port.read(False)
            if bottle:
                for i in range(min(bottle.size(), len(self.ports))):
                    self.weights[i] = bottle.get(i).asDouble()
        # update inputs (:56-69)
        now = time.time
        for p in range(len(self.ports)):
            bottle = self.ports[p].read(False)
            if bottle and bottle.size() == self.nChannels:
                self.last_command_time[p] = now()
                self.last_command[p] = [bottle.get(i).asDouble() for i in range(self.nChannels)]
            elif now() - self.last_command_time[p] > self.guard_time:
                self.last_command[p] = [0.0] * self.nChannels
            elif bottle:
                print('wrong length for data bottle')
        # NaN report (:71-75)
        for i, cmd in enumerate(self.last_command):
            for j, v in enumerate(cmd):
                if math.isnan(v):
                    print('nan: %d - %d' % (i, j))
        # weighted sum of all port data (:78-82) -> vfk_mix
        return self._mix_on_device()
