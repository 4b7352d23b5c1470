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
