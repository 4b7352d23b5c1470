"""In-process transport with the YARP surface the reference uses.

The reference glues its modules with YARP ports (``yarp.BufferedPortBottle``,
``read(False)`` latest-value reads, ``prepare()/write()``; e.g. ``scripts/vf:70-84,
312-315,462-466``).  ``yarp`` is not installed here and the batched runtime lives in one
process, so this module provides the same names with an in-process registry:
``Network.connect(src, dst)`` wires an output port to input ports, ``write()`` delivers a
copy of the prepared bottle.

Using a real YARP (row f4 of SURVEY.md section 8): every port the host modules own is created by
``ArcosYarp.create_yarp_port`` and wired by ``ArcosYarp.connect``, and both go through the
*transport* selected here.  ``use_transport(yarp)`` -- with the imported ``yarp`` python
module, or anything with the same surface -- makes them create ``yarp.BufferedPortBottle``
objects and call ``yarp.Network.connect``; the modules themselves only use the bottle / port
calls this file implements (``read(False)``, ``prepare``, ``write``, ``writeStrict``, ``size``,
``get(i).asDouble / asInt / asString / asList``, ``clear``, ``addDouble / addInt / addString /
addList``), so nothing else changes.  ``use_transport(None)`` returns to the in-process
registry.  ``yarp`` is not installed in this image, so the adapter is exercised by a
duck-typed stand-in (``tests/test_host.py::test_injected_transport``); against a real
``yarp`` it is untested.

Semantics kept from YARP: non-strict input ports keep only the newest unread bottle
(``read(False)`` returns it once, then ``None``); strict ports queue; ``read(True)`` on an
empty port raises here instead of blocking forever (single-threaded process).
"""
from __future__ import annotations

import time as _time
from collections import deque
from typing import Dict, List, Optional


class Value:
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = v

    def asDouble(self) -> float:
        return float(self.v)

    asFloat64 = asDouble

    def asInt(self) -> int:
        return int(self.v)

    asInt32 = asInt

    def asString(self) -> str:
        return str(self.v)

    def toString(self) -> str:
        return self.v.toString() if isinstance(self.v, Bottle) else str(self.v)

    def asList(self) -> Optional["Bottle"]:
        return self.v if isinstance(self.v, Bottle) else None

    def isDouble(self) -> bool:
        return isinstance(self.v, float)

    def isInt(self) -> bool:
        return isinstance(self.v, int) and not isinstance(self.v, bool)

    def isString(self) -> bool:
        return isinstance(self.v, str)

    def isList(self) -> bool:
        return isinstance(self.v, Bottle)

    @staticmethod
    def makeString(s: str) -> "Value":
        return Value(str(s))


def Value_makeString(s: str) -> Value:      # old SWIG spelling used at scripts/vf:220
    return Value.makeString(s)


class Bottle:
    def __init__(self, items=None):
        self.items: List[object] = list(items) if items is not None else []

    def clear(self):
        self.items = []

    def size(self) -> int:
        return len(self.items)

    def get(self, i: int) -> Value:
        return Value(self.items[i])

    def addDouble(self, v):
        self.items.append(float(v))

    addFloat64 = addDouble

    def addInt(self, v):
        self.items.append(int(v))

    addInt32 = addInt

    def addString(self, s):
        self.items.append(str(s))

    def addList(self) -> "Bottle":
        b = Bottle()
        self.items.append(b)
        return b

    def add(self, value):
        self.items.append(value.v if isinstance(value, Value) else value)

    def copy(self) -> "Bottle":
        return Bottle([x.copy() if isinstance(x, Bottle) else x for x in self.items])

    def toString(self) -> str:
        return " ".join("(%s)" % x.toString() if isinstance(x, Bottle) else str(x) for x in self.items)

    def to_list(self) -> list:
        return [x.to_list() if isinstance(x, Bottle) else x for x in self.items]

    @staticmethod
    def from_list(l) -> "Bottle":
        b = Bottle()
        for x in l:
            b.items.append(Bottle.from_list(x) if isinstance(x, (list, tuple)) else x)
        return b


class _Registry:
    def __init__(self):
        self.ports: Dict[str, "BufferedPortBottle"] = {}
        self.links: Dict[str, List[str]] = {}

    def reset(self):
        self.ports.clear()
        self.links.clear()


_REG = _Registry()


_TRANSPORT = None          # None: the in-process classes of this module


def use_transport(module=None):
    """Select what ``ArcosYarp`` creates ports with and connects them through: ``module.BufferedPortBottle`` and
    ``module.Network`` (e.g. the real ``yarp`` python module); ``None`` restores the in-process registry."""
    global _TRANSPORT
    _TRANSPORT = module


def transport():
    """The module ports are created from (this one unless ``use_transport`` chose another)."""
    import sys
    return _TRANSPORT if _TRANSPORT is not None else sys.modules[__name__]


class ContactStyle:
    persistent = False


class Network:
    @staticmethod
    def init():
        pass

    @staticmethod
    def fini():
        pass

    @staticmethod
    def connect(src: str, dst: str, style=None) -> bool:
        lst = _REG.links.setdefault(src, [])
        if dst not in lst:
            lst.append(dst)
        return True

    @staticmethod
    def disconnect(src: str, dst: str) -> bool:
        if dst in _REG.links.get(src, []):
            _REG.links[src].remove(dst)
        return True

    @staticmethod
    def isConnected(src: str, dst: str) -> bool:
        return dst in _REG.links.get(src, []) and src in _REG.ports and dst in _REG.ports

    @staticmethod
    def exists(name: str) -> bool:
        return name in _REG.ports

    @staticmethod
    def reset():
        """Forget every port and connection (test isolation)."""
        _REG.reset()


class BufferedPortBottle:
    def __init__(self):
        self.name: Optional[str] = None
        self.strict = False
        self._inbox = deque()
        self._out = Bottle()

    def open(self, name: str) -> bool:
        self.name = name
        _REG.ports[name] = self
        return True

    def close(self):
        if self.name and _REG.ports.get(self.name) is self:
            del _REG.ports[self.name]
        self.name = None

    def getName(self) -> str:
        return self.name or ""

    def setStrict(self, strict: bool = True):
        self.strict = bool(strict)

    # -- receiving
    def _deliver(self, bottle: Bottle):
        if not self.strict:
            self._inbox.clear()
        self._inbox.append(bottle)

    def read(self, shouldWait: bool = True) -> Optional[Bottle]:
        if self._inbox:
            return self._inbox.popleft()
        if shouldWait:
            raise RuntimeError("blocking read on empty in-process port %s (single-threaded runtime)" % self.name)
        return None

    def getPendingReads(self) -> int:
        return len(self._inbox)

    # -- sending
    def prepare(self) -> Bottle:
        return self._out

    def write(self, forceStrict: bool = False):
        for dst in _REG.links.get(self.name, []):
            port = _REG.ports.get(dst)
            if port is not None:
                port._deliver(self._out.copy())

    def writeStrict(self):
        self.write(True)


class Time:
    @staticmethod
    def delay(seconds: float):
        _time.sleep(seconds)

    @staticmethod
    def now() -> float:
        return _time.time()


def Time_delay(seconds: float):            # old SWIG spelling (scripts/nullspace:187)
    Time.delay(seconds)


# ------------------------------------------------------------------ arcospyu.yarp_tools helpers (App. B.4)

def sendListPort(port: BufferedPortBottle, values):
    b = port.prepare()
    b.clear()
    for v in values:
        b.addDouble(v)
    port.write()


def readListPort(port: BufferedPortBottle, blocking: bool = False):
    b = port.read(blocking)
    if b is None:
        return None
    return [b.get(i).asDouble() for i in range(b.size())]


def write_bottle_lists(port: BufferedPortBottle, items, strict: bool = False):
    """Write a (possibly nested) python list as a bottle (``arcospyu`` ``recur`` + ``write``)."""
    b = port.prepare()
    b.clear()
    for x in items:
        if isinstance(x, (list, tuple)):
            sub = b.addList()
            for y in x:
                if isinstance(y, (int, float)) and not isinstance(y, bool):
                    sub.addDouble(float(y))
                else:
                    sub.addString(y)
        elif isinstance(x, bool):
            b.addInt(int(x))
        elif isinstance(x, int):
            b.addInt(x)
        elif isinstance(x, float):
            b.addDouble(x)
        else:
            b.addString(x)
    port.writeStrict() if strict else port.write()


class ArcosYarp:
    """``arcospyu.yarp_tools.yarp_comm_helpers.ArcosYarp`` surface (``scripts/vf:70-84,129-131,188-194``).

    Port names are ``<ports_name_prefix><module_name_prefix><name>``; ``connect(port, remote_module,
    remote_port)`` links ``port`` with ``<prefix><remote_module><remote_port>`` in the direction
    implied by the local port (output ports write to the remote, input ports read from it).
    """

    def __init__(self, ports_name_prefix: str = "", module_name_prefix: str = ""):
        self.prefix = ports_name_prefix
        self.module = module_name_prefix
        self._ports = []
        self._wanted = []
        self._is_input = {}

    def create_yarp_port(self, name: str, input_port: bool = True, strict: bool = True) -> BufferedPortBottle:
        p = transport().BufferedPortBottle()
        p.open(self.prefix + self.module + name)
        self._is_input[id(p)] = input_port                 # kept here, not on the port: a SWIG proxy need not take attributes
        if input_port:
            p.setStrict(strict)
        self._ports.append(p)
        return p

    def connect(self, port: BufferedPortBottle, remote_module: str, remote_port: str, necessary: bool = True):
        remote = self.prefix + remote_module + remote_port
        src, dst = (remote, port.getName()) if self._is_input.get(id(port), True) else (port.getName(), remote)
        transport().Network.connect(src, dst)
        self._wanted.append((src, dst, necessary))

    def is_ready(self) -> bool:
        return all(transport().Network.isConnected(s, d) for s, d, nec in self._wanted if nec)

    def update(self):
        pass

    def close(self):
        for p in self._ports:
            p.close()
