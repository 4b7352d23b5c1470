"""Stress of the general (non-lean) kernel shapes at scale: per-obstacle safe/order rows, auxiliary fields, tool frame,
weights, all outputs, nullspace control mode -- every output finite, FP32 against FP64 GPU results on the same inputs.
usage: python scripts/stress_general.py [n]"""
import dataclasses, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vfclik_b200 import workloads
from vfclik_b200.config import PACKAGE_CONFIG_DIR, chain_from_config, config_filename, load_config
from vfclik_b200.engine import DeviceBatch, Engine, Params

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
cfg = load_config(config_filename(PACKAGE_CONFIG_DIR + "/lwr/", "lwr", "right"))
chain = chain_from_config(cfg)
M = 24
outs = ("qdot_vf", "qdot_ns", "qdot_jp", "qdot", "cmd", "pose", "twist", "flags")
base = Params.from_config(cfg)
cases = {
    "ext+aux+tool+weights": dataclasses.replace(base, tool=(0, -1, 0, 1, 0, 0, 0, 0, 1, 0.02, -0.01, 0.12), w_task=(1, 1, 1, 0.5, 0.5, 0.5),
                                                w_joint=(1, 1, 0.7, 1, 1, 0.4, 1), mixer_w=(1, 1, 0.3, 0, 0, 0)),
    "ns control mode": dataclasses.replace(base, ns_mode=2, ns_lambda=0.05, ns_control=(0.3, 0, 0, 0)),
}
res = {}
for prec in (64, 32):
    dt = np.float32 if prec == 32 else np.float64
    w = workloads.random_batch(chain, n, M, seed=321, dtype=dt, obst_ext=True)
    rng = np.random.default_rng(7)
    w["obst_ext"][:, :, 0] = rng.uniform(0.0005, 0.01, size=(M, n)).astype(dt)
    w["obst_ext"][:, :, 1] = rng.choice([2.0, 5.0, 7.5, 20.0], size=(M, n)).astype(dt)
    aux = np.zeros((2, 12, n), dtype=dt)
    aux[0, 0] = 4; aux[0, 1] = -50; aux[0, 2:5] = rng.uniform(-0.5, 0.5, size=(3, n)) + chain.base[9:12, None]; aux[0, 7] = 1.0; aux[0, 8] = 0.02; aux[0, 9] = 3
    aux[1, 0] = 5; aux[1, 1] = 30; aux[1, 2:5] = w["goal"][9:12]; aux[1, 6] = -1.0; aux[1, 8] = 0.3; aux[1, 9] = 10; aux[1, 10] = 0.15; aux[1, 11] = 2
    for name, prm in cases.items():
        e = Engine(chain, precision=prec, params=prm)
        db = DeviceBatch(e, n, M, obst_ext=True, outputs=outs)
        db.upload("q", w["q"]); db.upload("goal", w["goal"]); db.upload("obst", w["obst"]); db.upload("obst_ext", w["obst_ext"])
        if name.startswith("ext"):
            db.set_aux(aux)
        else:
            db.upload("ns_lastvec", np.zeros((7, n), dtype=dt))
        db.step(3)
        got = {k: db.download(k) for k in outs + ("q",)}
        fin = {k: bool(np.isfinite(v).all()) for k, v in got.items() if v.dtype.kind == "f"}
        res[(name, prec)] = got
        print("fp%d %-22s finite: %s" % (prec, name, "all" if all(fin.values()) else fin))
        e.close()
for name in cases:
    a, b = res[(name, 32)], res[(name, 64)]
    for k in ("qdot_vf", "qdot"):
        num = np.max(np.abs(a[k].astype(np.float64) - b[k]), axis=0)
        den = np.maximum(np.max(np.abs(b[k]), axis=0), 1e-3)
        err = num / den
        print("%-22s %-8s fp32 vs fp64 (3 cycles): median %.1e p99.9 %.1e max %.1e  frac>1e-3 %.1e" % (
            name, k, np.median(err), np.quantile(err, 0.999), err.max(), np.mean(err > 1e-3)))
