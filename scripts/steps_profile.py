"""Per-block launch time over a long run of the headline launch: does it drift with the state (q integrates towards the goals)
or with the box (power / thermal)?  Blocks of 100 launches; second pass restores q before every block."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as ge
ge.build()
from vfclik_b200 import workloads
from vfclik_b200.config import PACKAGE_CONFIG_DIR, chain_from_config, config_filename, load_config
from vfclik_b200.engine import DeviceBatch, Engine, Params
cfg = load_config(config_filename(PACKAGE_CONFIG_DIR + "/lwr/", "lwr", "right"))
chain = chain_from_config(cfg)
eng = Engine(chain, precision=32, params=Params.from_config(cfg))
n, M = 1 << 20, 32
w = workloads.random_batch(chain, n, M, seed=1, dtype=np.float32)
db = DeviceBatch(eng, n, M, outputs=("qdot",))
db.upload("q", w["q"]); db.upload("goal", w["goal"]); db.upload("obst", w["obst"])
q0 = db.t["q"].clone()
def block(k=100):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(k):
        db.step(1)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / k * 1e3
import pynvml
pynvml.nvmlInit()
hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
def clocks():
    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(hnd)
    return [pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_MEM),
            round(pynvml.nvmlDeviceGetPowerUsage(hnd) / 1000.0), int(r)]
# the same question for a plain copy of the same traffic (2 x 340 MB), from a cold start
import time
a = torch.empty(340 << 20, dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
def copy_block(k=100):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(k):
        b.copy_(a)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k * 1e3
time.sleep(3.0)
copy_block(5)
cp, ck = [], []
for _ in range(30):
    cp.append(round(copy_block(), 2)); ck.append(clocks())
del a, b
out = {"copy_680MB_us": cp, "copy_clocks_sm_mem_w_reasons": ck}
time.sleep(3.0)
for _ in range(5):
    db.step(1)
dr, dk = [], []
for _ in range(30):
    dr.append(round(block(), 2)); dk.append(clocks())
out["drift_us"] = dr
out["drift_clocks_sm_mem_w_reasons"] = dk
db.t["q"].copy_(q0)
reset = []
for _ in range(30):
    db.t["q"].copy_(q0)
    reset.append(round(block(), 2))
out["reset_each_block_us"] = reset
qd = db.download("qdot")
out["qdot_absmax_after"] = float(np.abs(qd).max())
out["frac_tiny"] = float((np.abs(qd) < 1e-30).mean())
print(json.dumps(out))
