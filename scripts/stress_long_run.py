"""Stress: many fused cycles on a large batch; every state must stay finite and converge or keep moving sensibly.
usage: python scripts/stress_long_run.py [precision] [n] [M] [K]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vfclik_b200 import workloads
from vfclik_b200.config import PACKAGE_CONFIG_DIR, chain_from_config, config_filename, load_config
from vfclik_b200.engine import DeviceBatch, Engine, Params

prec = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18
M = int(sys.argv[3]) if len(sys.argv) > 3 else 32
K = int(sys.argv[4]) if len(sys.argv) > 4 else 2000
cfg = load_config(config_filename(PACKAGE_CONFIG_DIR + "/lwr/", "lwr", "right"))
chain = chain_from_config(cfg)
e = Engine(chain, precision=prec, params=Params.from_config(cfg))
dt = np.float32 if prec == 32 else np.float64
w = workloads.random_batch(chain, n, M, seed=123, dtype=dt)
db = DeviceBatch(e, n, M, outputs=("qdot", "pose", "flags"))
db.upload("q", w["q"]); db.upload("goal", w["goal"]); db.upload("obst", w["obst"])
for block in range(4):
    db.step(K // 4)
    q = db.download("q"); qd = db.download("qdot"); pose = db.download("pose"); fl = db.download("flags")[0]
    dist = np.linalg.norm(pose[9:12] - w["goal"][9:12], axis=0)
    print("fp%d after %5d cycles: finite q %s qdot %s | max|q| %.2f | goal distance median %.4f p90 %.4f | clamped %.3f ns-limit %.3f nan-flag %d" % (
        prec, (block + 1) * (K // 4), bool(np.isfinite(q).all()), bool(np.isfinite(qd).all()), float(np.abs(q).max()),
        float(np.median(dist)), float(np.quantile(dist, 0.9)), float(np.mean((fl & 8) != 0)), float(np.mean((fl & 2) != 0)),
        int(np.sum((fl & 4) != 0))))
    bad = ~np.isfinite(q).all(axis=0)
    if bad.any():
        i = int(np.nonzero(bad)[0][0])
        print("  first bad instance", i, "q0", w["q"][:, i], "goal", w["goal"][:, i])
e.close()
