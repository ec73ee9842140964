"""Shared test helpers (oracle = checker only)."""
import numpy as np
import scipy.sparse as sp

import chebnet_oracle as O


def rel_err(y, ref, per_graph_off=None):
    """max over graphs of |y-ref|_inf / max(|ref|_inf per graph, tiny)  (SURVEY 7.2)."""
    y = np.asarray(y, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    if per_graph_off is None:
        return np.abs(y - ref).max() / max(np.abs(ref).max(), 1e-30)
    worst = 0.0
    for a, b in zip(per_graph_off[:-1], per_graph_off[1:]):
        if b > a:
            den = max(np.abs(ref[a:b]).max(), 1e-30)
            worst = max(worst, np.abs(y[a:b] - ref[a:b]).max() / den)
    return worst


def oracle_batch_forward(mats, X, weights, acts=None, slope=0.2):
    """Reference semantics: ONE GRAPH AT A TIME (gnn_offloading_agent.py:149), fp64."""
    outs, o = [], 0
    for A in mats:
        n = A.shape[0]
        outs.append(O.cheb_stack_forward(A, X[o:o + n], weights, acts, slope))
        o += n
    return np.concatenate(outs, axis=0)


def random_weights(specs, rng, scale=1.0, bias=0.05):
    ws = []
    for s in specs:
        lim = np.sqrt(6.0 / (s.K * s.f_in + s.K * s.f_out)) * scale
        ws.append((rng.uniform(-lim, lim, size=(s.K, s.f_in, s.f_out)), rng.normal(size=s.f_out) * bias))
    return ws
