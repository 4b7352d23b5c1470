"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np


def to_oracle(batch_arrays, n_obst):
    """GPU layouts (q [N,I], goal [13,I], obst [M,I,4], obst_ext [M,I,2]) -> the oracle's instance-major arrays."""
    q = batch_arrays["q"].T.astype(np.float64)
    goal = batch_arrays["goal"].T.astype(np.float64)
    obst = None
    if n_obst > 0:
        obst = batch_arrays["obst"].transpose(1, 0, 2).astype(np.float64)
        if "obst_ext" in batch_arrays:
            obst = np.concatenate([obst, batch_arrays["obst_ext"].transpose(1, 0, 2).astype(np.float64)], axis=2)
    return q, goal, obst


def rel_err(a, b, floor=1e-3):
    """Per-instance relative error of joint-velocity vectors: max_i |a_i - b_i| / max(max_i |b_i|, floor).

    a, b: [I, N].  The floor keeps instances whose reference velocity is ~0 from dividing by noise.
    """
    num = np.max(np.abs(a - b), axis=1)
    den = np.maximum(np.max(np.abs(b), axis=1), floor)
    return num / den


def oracle_params(params):
    """vfclik_b200.engine.Params -> oracle.batch.Params (same field names)."""
    import dataclasses
    from oracle import batch
    d = dataclasses.asdict(params)
    for k in ("mixer_w", "w_task", "tool", "ns_control"):
        d[k] = tuple(d[k])
    for k in ("w_joint", "jp_ref"):
        d[k] = None if d[k] is None else tuple(d[k])
    return batch.Params(**d)


# ---------------------------------------------------------------------------------------------- the oracle over all host cores
_PAR = {}


def _oracle_chunk(span):
    """One chunk of instances through oracle.batch.step (runs in a forked worker: the arrays are inherited, not pickled)."""
    from oracle import batch
    i0, i1 = span
    q, goal, obst = to_oracle({k: (v[:, i0:i1] if k != "obst" else v[:, i0:i1, :]) for k, v in _PAR["w"].items()}, _PAR["m"])
    out = batch.step(_PAR["chain"], _PAR["prm"], q, goal, obst, k_cycles=_PAR["k"])
    return i0, {k: out[k] for k in _PAR["keys"]}


def oracle_parallel(chain, params, w, n_obst, k=1, keys=("qdot",), chunk=16384):
    """oracle.batch.step over a large batch, fanned over the host cores in chunks (fork: zero-copy inputs).  The vectorised
    oracle does ~1.7e5 instance-cycles/s per core, so the full 1,048,576-instance configuration takes seconds, not minutes."""
    import multiprocessing as mp
    import os
    n = w["q"].shape[1]
    _PAR.update(chain=chain, prm=oracle_params(params), w=w, m=n_obst, k=k, keys=tuple(keys))
    spans = [(i, min(n, i + chunk)) for i in range(0, n, chunk)]
    procs = max(1, min(len(spans), len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)))
    with mp.get_context("fork").Pool(procs) as pool:
        parts = pool.map(_oracle_chunk, spans)
    _PAR.clear()
    out = {}
    for key in keys:
        first = parts[0][1][key]
        full = np.empty((n,) + first.shape[1:], dtype=first.dtype)
        for i0, d in parts:
            full[i0:i0 + d[key].shape[0]] = d[key]
        out[key] = full
    return out
