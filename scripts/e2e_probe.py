"""e2e step time of the host-buffer session (1 M instances, FP32) under a given environment (VFK_SESSION_DIRECT etc.)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
from vfclik_b200 import workloads
from vfclik_b200.config import PACKAGE_CONFIG_DIR, chain_from_config, config_filename, load_config
from vfclik_b200.engine import Engine, Params
cfg = load_config(config_filename(PACKAGE_CONFIG_DIR + "/lwr/", "lwr", "right"))
chain = chain_from_config(cfg)
eng = Engine(chain, precision=32, params=Params.from_config(cfg))
n, M = 1 << 20, 32
w = workloads.random_batch(chain, n, M, seed=1, dtype=np.float32)
s = eng.session(n, M); s.set_goal(w["goal"]); s.set_obstacles(w["obst"])
q = torch.from_numpy(w["q"]).pin_memory().numpy(); qd = torch.empty((7, n), dtype=torch.float32).pin_memory().numpy()
ref = None
out = {}
for mode in sys.argv[1:]:
    if mode == "-": os.environ.pop("VFK_SESSION_DIRECT", None)
    else: os.environ["VFK_SESSION_DIRECT"] = mode
    for _ in range(3): s.cycle(q_in=q, qdot_out=qd)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): s.cycle(q_in=q, qdot_out=qd)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 20 * 1e3
    if ref is None: ref = qd.copy()
    out[mode] = {"ms": round(ms, 4), "same": bool(np.array_equal(ref, qd))}
print(json.dumps(out))
