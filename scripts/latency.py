"""Single-robot / small-batch latency: microseconds per control cycle with K cycles fused (cooperative vs solo shape)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vfclik_b200 import workloads
from vfclik_b200.config import PACKAGE_CONFIG_DIR, chain_from_config, config_filename, load_config
from vfclik_b200.engine import DeviceBatch, Engine, Params
cfg = load_config(config_filename(PACKAGE_CONFIG_DIR + "/lwr/", "lwr", "right"))
chain = chain_from_config(cfg)
K = 1000
for precision in (64, 32):
    e = Engine(chain, precision=precision, params=Params.from_config(cfg))
    for n, M in ((1, 3), (1, 32), (1, 256), (64, 32), (1024, 32), (4096, 32)):
        w = workloads.random_batch(chain, n, M, seed=3, dtype=np.float32 if precision == 32 else np.float64)
        db = DeviceBatch(e, n, M, outputs=("qdot",))
        db.upload("q", w["q"]); db.upload("goal", w["goal"]); db.upload("obst", w["obst"])
        db.step(K); torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(3):
            db.step(K)
        ev1.record(); torch.cuda.synchronize()
        us = ev0.elapsed_time(ev1) * 1e3 / 3 / K
        print("fp%d coop=%s n=%d M=%d: %.2f us per control cycle (%.3e inst-cycles/s)" % (
            precision, "off" if os.environ.get("VFK_NO_COOP") else "auto", n, M, us, n / (us * 1e-6)))
    e.close()
