"""FP32-mode accuracy report: per-instance relative qdot error vs the FP64 oracle on N random LWR instances.

usage: python scripts/fp32_error.py [n] [n_obst] [order] [ik_lambda]
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import oracle_params, rel_err, to_oracle
from oracle import batch
from vfclik_b200 import workloads
from vfclik_b200.config import PACKAGE_CONFIG_DIR, chain_from_config, config_filename, load_config
from vfclik_b200.engine import DeviceBatch, Engine, Params

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
M = int(sys.argv[2]) if len(sys.argv) > 2 else 32
order = float(sys.argv[3]) if len(sys.argv) > 3 else 20.0
lam = float(sys.argv[4]) if len(sys.argv) > 4 else 0.1
cfg = load_config(config_filename(PACKAGE_CONFIG_DIR + "/lwr/", "lwr", "right"))
chain = chain_from_config(cfg)
e = Engine(chain, precision=32, params=Params.from_config(cfg, obst_order=order, ik_lambda=lam, ns_lambda=lam))
w = workloads.random_batch(chain, n, M, seed=1, dtype=np.float32)
db = DeviceBatch(e, n, M, outputs=("qdot", "qdot_vf", "qdot_ns", "twist"))
db.upload("q", w["q"]); db.upload("goal", w["goal"])
if M:
    db.upload("obst", w["obst"])
db.step(1)
q, goal, obst = to_oracle(w, M)
ref = batch.step(chain, oracle_params(e.params), q, goal, obst)
tw = db.download("twist").T.astype(np.float64)
terr = np.max(np.abs(tw - ref["twist"]), axis=1) / np.maximum(np.max(np.abs(ref["twist"]), axis=1), 1e-3)
print("twist n=%d M=%d order=%g lambda=%g rel err: median %.2e p99.9 %.2e p99.99 %.2e max %.2e" % (
    n, M, order, lam, np.median(terr), np.quantile(terr, 0.999), np.quantile(terr, 0.9999), terr.max()))
for k in ("qdot_vf", "qdot_ns", "qdot"):
    err = rel_err(db.download(k).T.astype(np.float64), ref[k])
    print("%s lib=%s n=%d rel err: median %.2e p99 %.2e p99.9 %.2e p99.99 %.2e max %.2e  frac>1e-4 %.2e" % (
        k, os.path.basename(os.environ.get("VFK_LIB", "libvfk.so")), n, np.median(err), np.quantile(err, 0.99),
        np.quantile(err, 0.999), np.quantile(err, 0.9999), err.max(), np.mean(err > 1e-4)))
