"""FP32-mode accuracy report: per-instance relative qdot error vs the FP64 oracle on N random LWR instances."""
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
cfg = load_config(config_filename(PACKAGE_CONFIG_DIR + "/lwr/", "lwr", "right"))
chain = chain_from_config(cfg)
e = Engine(chain, precision=32, params=Params.from_config(cfg))
w = workloads.random_batch(chain, n, 32, seed=1, dtype=np.float32)
db = DeviceBatch(e, n, 32, outputs=("qdot", "qdot_vf", "qdot_ns"))
db.upload("q", w["q"]); db.upload("goal", w["goal"]); db.upload("obst", w["obst"])
db.step(1)
q, goal, obst = to_oracle(w, 32)
ref = batch.step(chain, oracle_params(e.params), q, goal, obst)
for k in ("qdot_vf", "qdot_ns", "qdot"):
    err = rel_err(db.download(k).T.astype(np.float64), ref[k])
    print("%s lib=%s n=%d rel err: median %.2e p99 %.2e p99.9 %.2e max %.2e  frac>1e-4 %.2e" % (
        k, os.path.basename(os.environ.get("VFK_LIB", "libvfk.so")), n, np.median(err), np.quantile(err, 0.99),
        np.quantile(err, 0.999), err.max(), np.mean(err > 1e-4)))
