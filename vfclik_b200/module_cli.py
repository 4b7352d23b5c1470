"""Command-line contract of the reference's per-module programs.

The launcher starts every module as ``<module> -c <config file> -n <namespace>`` (``scripts/vfclik:88-105``); each
module parses that with ``arcospyu.config_parser.ConfigFileParser(sys.argv).get_all()`` -> ``(options, args, config)``
(``scripts/vf:63-64``, ``scripts/nullspace:31-32``, ``scripts/joint_p_controller:52-53``) and ``bridge`` adds
``-s/--simulation`` through ``config_parser.parser.add_option`` (``scripts/bridge:58-66``).  ``ConfigFileParser`` here
keeps that surface; ``run_module`` is the single-threaded polling loop with the SIGINT / SIGTERM stop flag every
reference module has (``scripts/vf:52-61``).

``python -m vfclik_b200.vf -c <config> -n /0`` therefore starts exactly one module on its own batched runtime.  The
in-process port registry (``ports.py``) does not cross processes, so stand-alone modules are only useful behind a real
YARP transport adapter; the launcher (``python -m vfclik_b200.launcher``) wires all of them in one process instead.
"""
from __future__ import annotations

import optparse
import signal
import sys
import time

from .config import load_config


class ConfigFileParser:
    def __init__(self, argv):
        self.argv = list(argv)
        self.parser = optparse.OptionParser("usage: %prog [options]")
        self.parser.add_option("-c", "--config_filename", dest="config_filename", default="", type="string",
                               help="config filename")
        self.parser.add_option("-n", "--namespace", dest="namespace", default="", type="string",
                               help="ARCOS-Lab yarp basename")
        # not in the reference: bounded runs for tests / batch jobs
        self.parser.add_option("--cycles", dest="cycles", default=0, type="int", help="stop after this many iterations (0 = run)")
        self.parser.add_option("--precision", dest="precision", default=64, type="int", help="kernel arithmetic: 32 or 64")
        self.parser.add_option("--no_sleep", action="store_true", dest="no_sleep", default=False, help="do not pace at config.rate")

    def get_all(self):
        options, args = self.parser.parse_args(self.argv[1:])
        if not options.config_filename:
            self.parser.error("a config file is required (-c)")
        config = load_config(options.config_filename)
        return options, args, config


class StopFlag:
    """``stop`` global + signal handler of every reference module (``scripts/vf:52-61``)."""

    def __init__(self):
        self.stop = False

    def install(self):
        signal.signal(signal.SIGINT, self)
        signal.signal(signal.SIGTERM, self)
        return self

    def __call__(self, sig, frame):
        print("Terminating, ", __file__)
        self.stop = True


def run_module(argv, build, extra_options=None, simulate_plant: bool = False, needs_runtime: bool = True):
    """Parse the module's command line, build it on a one-instance ControlRuntime and poll it at ``config.rate``.

    build(runtime, options, config) -> list of objects with ``update()`` (and optionally ``finish()`` / ``close()``),
    called in order every period.
    """
    from . import ports as yarp
    from .runtime import ControlRuntime
    parser = ConfigFileParser(argv)
    for args, kw in (extra_options or []):
        parser.parser.add_option(*args, **kw)
    options, _, config = parser.get_all()
    yarp.Network.init()
    runtime = (ControlRuntime(config, n_instances=1, precision=options.precision, simulate_plant=simulate_plant)
               if needs_runtime else None)
    modules = build(runtime, options, config)
    flag = StopFlag().install()
    n = 0
    try:
        while not flag.stop and (options.cycles == 0 or n < options.cycles):
            t0 = time.time()
            for m in modules:
                m.update()
            for m in modules:
                if hasattr(m, "finish"):
                    m.finish()
            n += 1
            if not options.no_sleep:
                left = float(config.rate) - (time.time() - t0)
                if left > 0:
                    time.sleep(left)
    finally:
        for m in modules:
            if hasattr(m, "close"):
                m.close()
        if runtime is not None:
            runtime.close()
    print("iterations:", n)
    return 0
