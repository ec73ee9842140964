"""Shared test helpers (oracle = checker only)."""
import numpy as np
import scipy.sparse as sp

import chebnet_oracle as O


def rel_err(y, ref, per_graph_off=None, scale=None):
    """max over graphs of |y-ref|_inf / max(|ref|_inf per graph, scale_g, tiny)  (SURVEY 7.2).

    `scale` (optional, one value per graph) is the inf-norm of the last layer's PRE-activation:
    a relu output that is (almost) entirely dead has |y_ref| << |z_ref|, and an error of one fp32
    ulp of z would otherwise read as a huge "relative" error of y ("relu outputs that
    are exactly 0 in one and 1e-9 in the other must not fail a pure-relative check")."""
    y = np.asarray(y, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    if per_graph_off is None:
        den = max(np.abs(ref).max(), 1e-30 if scale is None else float(np.max(scale)))
        return np.abs(y - ref).max() / den
    worst = 0.0
    for g, (a, b) in enumerate(zip(per_graph_off[:-1], per_graph_off[1:])):
        if b > a:
            den = max(np.abs(ref[a:b]).max(), 1e-30, 0.0 if scale is None else float(scale[g]))
            worst = max(worst, np.abs(y[a:b] - ref[a:b]).max() / den)
    return worst


def oracle_batch_forward(mats, X, weights, acts=None, slope=0.2, return_scale=False):
    """Reference semantics: ONE GRAPH AT A TIME (gnn_offloading_agent.py:149), fp64.
    return_scale: also the per-graph inf-norm of the last layer's pre-activation."""
    outs, scales, o = [], [], 0
    for A in mats:
        n = A.shape[0]
        y, cache = O.cheb_stack_forward(A, X[o:o + n], weights, acts, slope, return_cache=True)
        outs.append(y)
        scales.append(np.abs(cache[-1][1]).max() if n else 0.0)
        o += n
    Y = np.concatenate(outs, axis=0)
    return (Y, np.asarray(scales)) if return_scale else Y


def random_weights(specs, rng, scale=1.0, bias=0.05):
    ws = []
    for s in specs:
        lim = np.sqrt(6.0 / (s.K * s.f_in + s.K * s.f_out)) * scale
        ws.append((rng.uniform(-lim, lim, size=(s.K, s.f_in, s.f_out)), rng.normal(size=s.f_out) * bias))
    return ws


def numpy_fp32_forward(mats, X, weights, acts, slope=0.2):
    """The same recurrence in plain numpy fp32 (scipy CSR @ + matmul): the error a straightforward
    fp32 implementation makes against the fp64 oracle - the yardstick for 'fp32-grade' parity."""
    outs, o = [], 0
    for A in mats:
        n = A.shape[0]
        A32 = sp.csr_matrix(A).astype(np.float32)
        h = X[o:o + n].astype(np.float32)
        for (W, b), act in zip(weights, acts):
            W = W.astype(np.float32); K = W.shape[0]
            Ts = [h]
            if K > 1: Ts.append(A32 @ h)
            for _ in range(2, K): Ts.append(np.float32(2.0) * (A32 @ Ts[-1]) - Ts[-2])
            z = sum(T @ W[k] for k, T in enumerate(Ts)) + b.astype(np.float32)
            h = z if act == 0 else (np.maximum(z, 0) if act == 1 else np.where(z > 0, z, np.float32(slope) * z))
        outs.append(h); o += n
    return np.concatenate(outs, axis=0)
