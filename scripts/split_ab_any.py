"""Lane-split vs one-thread-per-instance kernel on a DH chain of given length / precision: us per launch (lean call)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from vfclik_b200 import workloads
from vfclik_b200.engine import Engine, Params
n_joints, precision, n, M = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
chain = workloads.dual_arm_torso_chain(n_joints)
e = Engine(chain, precision=precision, params=Params())
db = workloads.random_batch_device(e, n, M, seed=4)
res = {}
for mode in ("0", "1"):
    os.environ["VFK_SPLIT"] = mode
    for _ in range(5): db.step(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(100): db.step(1)
    b.record(); torch.cuda.synchronize()
    res[mode] = a.elapsed_time(b) * 10
print("N=%d fp%d n=%d M=%d  solo %.1f us  split %.1f us" % (n_joints, precision, n, M, res["0"], res["1"]))
