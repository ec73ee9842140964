"""CPU fp64 oracle for the ChebConv hot path of zhongyuanzhao/multihop-offload.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker (or
as the CPU thing being timed), never as the thing shipped.

PARITY PINNING: the arithmetic restated here lives in two un-vendored, un-pinned
third-party packages of the reference (``spektral`` and ``tensorflow``,
``src/requirements.txt:9,12``) that cannot be installed in this sandbox, and the
reference holds no tests or golden outputs for this path.  The oracle is
therefore pinned *statistically* (``tests/test_agent_host.py::test_statistical_pin_forward_env (30 files) and ::test_statistical_pin_150_files (slow)`` replays the
reference's AdHoc_test protocol through this oracle + the reference's own
environment and compares per-size mean tau with the shipped result CSV) and by
algebraic invariants (Chebyshev polynomials on a diagonal operator, finite
differences for the VJP).  At the Spektral/TF boundary itself parity is
"unpinned" in the strict sense; DESIGN.md says so too.

What each function restates (file:line into /root/reference):
  cheb_layer_forward   spektral.layers.ChebConv.call [upstream, Spektral>=1.0],
                       called from src/gnn_offloading_agent.py:95-110 via :149
  cheb_stack_forward   ACOAgent._build_model      src/gnn_offloading_agent.py:81-123
  cheb_stack_backward  g.gradient(..., weights)   src/gnn_offloading_agent.py:448
  queue_head_forward   ACOAgent.forward           src/gnn_offloading_agent.py:229-276
  queue_head_vjp       tape through :229-274 seeded at :448
  keras_adam_replay    ACOAgent.replay + Adam(clipnorm=1) + max_norm(1)
                       src/gnn_offloading_agent.py:104-121,156-169
  read_tf_bundle       ACOAgent.load              src/gnn_offloading_agent.py:125-129
"""
from __future__ import annotations

import os
import struct

import numpy as np
import scipy.sparse as sp

ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2
DEFAULT_SLOPE = 0.2  # tf.nn.leaky_relu / keras.activations.leaky_relu default


# --------------------------------------------------------------------------- #
# ChebConv layer / stack
# --------------------------------------------------------------------------- #
def _act(z, act, slope):
    if act == ACT_NONE:
        return z
    if act == ACT_RELU:
        return np.maximum(z, 0.0)
    if act == ACT_LEAKY:
        return np.where(z > 0, z, slope * z)
    raise ValueError(act)


def _act_grad(z, act, slope):
    if act == ACT_NONE:
        return np.ones_like(z)
    if act == ACT_RELU:
        return (z > 0).astype(z.dtype)
    if act == ACT_LEAKY:
        return np.where(z > 0, 1.0, slope)
    raise ValueError(act)


def cheb_basis(A, X, K):
    """[T_0 X, ..., T_{K-1} X] with T_0=I, T_1=A, T_k = 2 A T_{k-1} - T_{k-2}."""
    Ts = [X]
    if K > 1:
        Ts.append(A @ X)
    for _ in range(2, K):
        Ts.append(2.0 * (A @ Ts[-1]) - Ts[-2])
    return Ts


def cheb_layer_forward(A, X, W, b, act=ACT_LEAKY, slope=DEFAULT_SLOPE, return_pre=False):
    """One ChebConv layer.  A: scipy sparse (n,n); X: (n,F_in); W: (K,F_in,F_out); b: (F_out,)."""
    X = np.asarray(X, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    K = W.shape[0]
    Ts = cheb_basis(A, X, K)
    z = Ts[0] @ W[0]
    for k in range(1, K):
        z = z + Ts[k] @ W[k]
    if b is not None:
        z = z + np.asarray(b, dtype=np.float64)
    y = _act(z, act, slope)
    if return_pre:
        return y, z, Ts
    return y


def stack_activations(n_layers):
    """leaky_relu x (L-1) then relu  (src/gnn_offloading_agent.py:88-93)."""
    return [ACT_LEAKY] * (n_layers - 1) + [ACT_RELU]


def cheb_stack_forward(A, X, weights, acts=None, slope=DEFAULT_SLOPE, return_cache=False):
    """weights: list of (W[K,Fi,Fo], b[Fo]).  Returns Y (n, F_last)."""
    if acts is None:
        acts = stack_activations(len(weights))
    A = sp.csr_matrix(A).astype(np.float64)
    h = np.asarray(X, dtype=np.float64)
    cache = []
    for (W, b), act in zip(weights, acts):
        y, z, Ts = cheb_layer_forward(A, h, W, b, act, slope, return_pre=True)
        cache.append((Ts, z, act))
        h = y
    if return_cache:
        return h, cache
    return h


def cheb_stack_backward(A, weights, cache, dY, slope=DEFAULT_SLOPE):
    """VJP of the stack.  Returns ([(dW, db), ...], dX)."""
    A = sp.csr_matrix(A).astype(np.float64)
    At = A.T.tocsr()
    g = np.asarray(dY, dtype=np.float64)
    grads = [None] * len(weights)
    for li in range(len(weights) - 1, -1, -1):
        W, _ = weights[li]
        W = np.asarray(W, dtype=np.float64)
        Ts, z, act = cache[li]
        K = W.shape[0]
        G = g * _act_grad(z, act, slope)
        db = G.sum(axis=0)
        dW = np.stack([Ts[k].T @ G for k in range(K)], axis=0)
        U = [G @ W[k].T for k in range(K)]
        for k in range(K - 1, 1, -1):
            U[k - 1] = U[k - 1] + 2.0 * (At @ U[k])
            U[k - 2] = U[k - 2] - U[k]
        g = U[0] + (At @ U[1]) if K > 1 else U[0]
        grads[li] = (dW, db)
    return grads, g


# --------------------------------------------------------------------------- #
# flat parameter layout shared with the C-ABI (kernel, bias per layer, creation order)
# --------------------------------------------------------------------------- #
def flatten_params(weights):
    return np.concatenate([np.concatenate([np.asarray(W).ravel(), np.asarray(b).ravel()]) for W, b in weights])


def unflatten_params(flat, shapes):
    """shapes: list of (K, Fi, Fo)."""
    out, o = [], 0
    for K, Fi, Fo in shapes:
        W = np.asarray(flat[o:o + K * Fi * Fo]).reshape(K, Fi, Fo); o += K * Fi * Fo
        b = np.asarray(flat[o:o + Fo]); o += Fo
        out.append((W, b))
    assert o == len(flat)
    return out


def glorot_weights(dims, K, rng):
    """Keras glorot_uniform on a (K,Fi,Fo) kernel: fan_in=K*Fi? No: Keras computes fans for a
    rank-3 shape as receptive_field=K, fan_in=Fi*K, fan_out=Fo*K (keras initializers
    _compute_fans [upstream]); zeros bias (src/gnn_offloading_agent.py:101-102)."""
    ws = []
    for fi, fo in zip(dims[:-1], dims[1:]):
        lim = np.sqrt(6.0 / (K * fi + K * fo))
        ws.append((rng.uniform(-lim, lim, size=(K, fi, fo)), np.zeros(fo)))
    return ws


# --------------------------------------------------------------------------- #
# Queue-model head (post-GNN), src/gnn_offloading_agent.py:229-276
# --------------------------------------------------------------------------- #
def queue_head_forward(lam, maps_ol_el, maps_on_el, link_rates, cf_degs, proc_bws, adj_i, T,
                       return_cache=False):
    """lam (n_ext,1) -> link_delay (L,1), node_delay (n_comp,1).  :231-254."""
    lam = np.asarray(lam, dtype=np.float64).reshape(-1, 1)
    ll = lam[np.asarray(maps_ol_el)]
    nl = lam[np.asarray(maps_on_el)]
    proc = np.asarray(proc_bws, dtype=np.float64)
    node_mu = proc[proc > 0].reshape(-1, 1)
    rates = np.asarray(link_rates, dtype=np.float64).reshape(-1, 1)
    link_mu = rates / (np.asarray(cf_degs, dtype=np.float64).reshape(-1, 1) + 1.0)
    Ai = sp.csr_matrix(adj_i).astype(np.float64)
    mus = [link_mu]
    for _ in range(10):
        busy = np.clip(ll / link_mu, 0.0, 1.0)
        link_mu = rates / (1.0 + Ai @ busy)
        mus.append(link_mu)
    ld = 1.0 / (link_mu - ll)
    nd = 1.0 / (node_mu - nl)
    lc = (ll - link_mu) > 0
    nc = (nl - node_mu) > 0
    ld = np.where(lc, float(T) * ll / (101.0 * link_mu), ld)
    nd = np.where(nc, float(T) * nl / (100.0 * node_mu), nd)
    if return_cache:
        return ld, nd, dict(ll=ll, nl=nl, mus=mus, node_mu=node_mu, rates=rates, Ai=Ai, lc=lc, nc=nc, T=float(T))
    return ld, nd


def queue_head_vjp(cache, g_ld, g_nd, n_ext, maps_ol_el, maps_on_el):
    """VJP of queue_head_forward wrt lam, differentiating through all 10 fixed-point
    iterations like the TF tape does (:240-244).  clip_by_value passes gradient where
    0 <= x <= 1 (TF: gradient of minimum/maximum, ties go to the input)."""
    ll, nl, mus = cache["ll"], cache["nl"], cache["mus"]
    node_mu, rates, Ai, lc, nc, T = (cache[k] for k in ("node_mu", "rates", "Ai", "lc", "nc", "T"))
    mu = mus[-1]
    g_ld = np.asarray(g_ld, dtype=np.float64).reshape(-1, 1)
    g_nd = np.asarray(g_nd, dtype=np.float64).reshape(-1, 1)
    # link delay
    d = mu - ll
    g_ll = np.where(lc, g_ld * T / (101.0 * mu), g_ld / d ** 2)
    g_mu = np.where(lc, -g_ld * T * ll / (101.0 * mu ** 2), -g_ld / d ** 2)
    # unroll the fixed point backwards: mu_{t+1} = rates / (1 + Ai @ clip(ll/mu_t))
    for t in range(9, -1, -1):
        mu_t = mus[t]
        r = ll / mu_t
        busy = np.clip(r, 0.0, 1.0)
        denom = 1.0 + Ai @ busy
        g_den = -g_mu * rates / denom ** 2
        g_busy = Ai.T @ g_den
        pas = ((r >= 0.0) & (r <= 1.0)).astype(np.float64)
        g_r = g_busy * pas
        g_ll = g_ll + g_r / mu_t
        g_mu = -g_r * ll / mu_t ** 2  # gradient wrt mu_t (mu_0 is a constant)
    dn = node_mu - nl
    g_nl = np.where(nc, g_nd * T / (100.0 * node_mu), g_nd / dn ** 2)
    g_lam = np.zeros((n_ext, 1))
    np.add.at(g_lam, np.asarray(maps_ol_el), g_ll)
    np.add.at(g_lam, np.asarray(maps_on_el), g_nl)
    return g_lam


def delay_matrix(ld, nd, num_nodes, edges, link_matrix, comp_nodes, bug_compatible=True):
    """:257-274.  bug_compatible=True is the numpy twin (np.fill_diagonal cycling an
    (n_comp,1) array, :269); False is the TF tensor (relays = +inf, :270-274)."""
    D = np.full((num_nodes, num_nodes), np.nan)
    for (e0, e1) in edges:
        D[e0, e1] = D[e1, e0] = ld[link_matrix[e0, e1], 0]
    if bug_compatible:
        np.fill_diagonal(D, nd)
    else:
        diag = np.full(num_nodes, np.inf)
        diag[np.asarray(comp_nodes)] = nd[:, 0]
        np.fill_diagonal(D, diag)
    return D


# --------------------------------------------------------------------------- #
# Optimizer: Keras-2 Adam(clipnorm=1.0) + max_norm(1.0, axis=0) constraint  
# --------------------------------------------------------------------------- #
class KerasAdam:
    def __init__(self, shapes, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-7, clipnorm=1.0,
                 max_norm=1.0, decay_rate=1.0, decay_steps=100):
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.clipnorm, self.max_norm = clipnorm, max_norm
        self.decay_rate, self.decay_steps = decay_rate, decay_steps
        self.m = [np.zeros(s) for s in shapes]
        self.v = [np.zeros(s) for s in shapes]
        self.t = 0

    def apply(self, params, grads):
        """params/grads: list of ndarrays (kernel, bias, kernel, bias, ...); updates in place."""
        lr = self.lr if self.decay_rate == 1.0 else self.lr * self.decay_rate ** (self.t / self.decay_steps)
        self.t += 1
        t = self.t
        alpha = lr * np.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t)
        for i, (p, g) in enumerate(zip(params, grads)):
            g = np.asarray(g, dtype=np.float64)
            if self.clipnorm is not None:
                nrm = np.sqrt((g * g).sum())
                # tf.clip_by_norm: g * clipnorm / max(norm, clipnorm)
                g = g * self.clipnorm / max(nrm, self.clipnorm)
            self.m[i] = self.b1 * self.m[i] + (1 - self.b1) * g
            self.v[i] = self.b2 * self.v[i] + (1 - self.b2) * g * g
            p -= alpha * self.m[i] / (np.sqrt(self.v[i]) + self.eps)
            if self.max_norm is not None:
                # keras.constraints.MaxNorm(max_value, axis=0): w *= clip(norm,0,max)/(1e-7+norm)
                nr = np.sqrt((p * p).sum(axis=0, keepdims=True))
                p *= np.clip(nr, 0, self.max_norm) / (1e-7 + nr)
        return params


# --------------------------------------------------------------------------- #
# TF tensor-bundle reader, independent of the product reader
# --------------------------------------------------------------------------- #
def _varint(buf, pos):
    out = shift = 0
    while True:
        c = buf[pos]; pos += 1
        out |= (c & 0x7F) << shift
        if c < 0x80:
            return out, pos
        shift += 7


def _parse_proto(buf):
    """Minimal protobuf wire parser -> list of (field, wiretype, value)."""
    pos, out = 0, []
    while pos < len(buf):
        key, pos = _varint(buf, pos)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = bytes(buf[pos:pos + ln]); pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]; pos += 4
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]; pos += 8
        else:
            raise ValueError("wire type %d" % wt)
        out.append((f, wt, v))
    return out


def read_tf_bundle(prefix):
    """prefix = '<dir>/cp-0000.ckpt'.  Returns {key: ndarray(float64)} for DT_DOUBLE/DT_FLOAT entries."""
    idx = open(prefix + ".index", "rb").read()
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    assert idx[-8:] == bytes.fromhex("57fb808b247547db"), "bad sstable magic"
    footer = idx[-48:]
    _, p = _varint(footer, 0); _, p = _varint(footer, p)       # metaindex handle
    ioff, p = _varint(footer, p); isz, p = _varint(footer, p)  # index handle
    tensors = {}

    def block_entries(off, size):
        blk = idx[off:off + size]
        nrest = struct.unpack_from("<I", blk, len(blk) - 4)[0]
        end = len(blk) - 4 - 4 * nrest
        pos, key = 0, b""
        while pos < end:
            sh, pos = _varint(blk, pos); ns, pos = _varint(blk, pos); vl, pos = _varint(blk, pos)
            key = key[:sh] + blk[pos:pos + ns]; pos += ns
            val = blk[pos:pos + vl]; pos += vl
            yield key, val

    for _, handle in block_entries(ioff, isz):
        boff, p = _varint(handle, 0); bsz, _ = _varint(handle, p)
        for key, val in block_entries(boff, bsz):
            if key == b"":
                continue
            dtype, shape, offset, size = 0, [], 0, 0
            for f, wt, v in _parse_proto(val):
                if f == 1: dtype = v
                elif f == 2:
                    for f2, _, v2 in _parse_proto(v):
                        if f2 == 2:
                            dim = 0
                            for f3, _, v3 in _parse_proto(v2):
                                if f3 == 1: dim = v3
                            shape.append(dim)
                elif f == 4: offset = v
                elif f == 5: size = v
            if dtype == 2:
                tensors[key.decode()] = np.frombuffer(data, "<f8", size // 8, offset).reshape(shape).copy()
            elif dtype == 1:
                tensors[key.decode()] = np.frombuffer(data, "<f4", size // 4, offset).reshape(shape).astype(np.float64)
    return tensors


def load_reference_weights(ckpt_dir):
    """ACOAgent.load (src/gnn_offloading_agent.py:125-129): follow the 'checkpoint' text file."""
    name = None
    for line in open(os.path.join(ckpt_dir, "checkpoint")):
        if line.startswith("model_checkpoint_path:"):
            name = line.split('"')[1]
    t = read_tf_bundle(os.path.join(ckpt_dir, name))
    ws, li = [], 0
    while "layer_with_weights-%d/kernel/.ATTRIBUTES/VARIABLE_VALUE" % li in t:
        ws.append((t["layer_with_weights-%d/kernel/.ATTRIBUTES/VARIABLE_VALUE" % li],
                   t["layer_with_weights-%d/bias/.ATTRIBUTES/VARIABLE_VALUE" % li]))
        li += 1
    return ws


# --------------------------------------------------------------------------- #
# Synthetic inputs of the benchmark: BA graphs -> block-diagonal CSR batch
# --------------------------------------------------------------------------- #
def ba_adjacency(n, m=2, seed=0):
    """networkx.barabasi_albert_graph (same generator as src/offloading_v3.py:40)."""
    import networkx as nx
    g = nx.barabasi_albert_graph(int(n), m, seed=int(seed))
    return sp.csr_matrix(nx.adjacency_matrix(g)).astype(np.float64)


def cheb_laplacian(A):
    """Spektral ChebConv.preprocess: 2 L_sym / lambda_max - I (never called by the reference)."""
    A = sp.csr_matrix(A).astype(np.float64)
    d = np.asarray(A.sum(1)).ravel()
    dinv = np.where(d > 0, 1.0 / np.sqrt(np.maximum(d, 1e-300)), 0.0)
    L = sp.identity(A.shape[0]) - sp.diags(dinv) @ A @ sp.diags(dinv)
    lmax = np.linalg.eigvalsh(L.toarray()).max()
    return sp.csr_matrix(2.0 / lmax * L - sp.identity(A.shape[0]))


def make_batch(sizes, seed0=1000, operator="raw-adj"):
    """Returns list of per-graph CSR operators."""
    mats = []
    for i, n in enumerate(sizes):
        A = ba_adjacency(n, 2, seed0 + i)
        if operator == "cheb-lap":
            A = cheb_laplacian(A)
        mats.append(A)
    return mats


def concat_batch(mats):
    """Block-diagonal concatenation -> (graph_off[B+1], rowptr[N+1], colidx[nnz] global, vals[nnz])."""
    graph_off = np.zeros(len(mats) + 1, dtype=np.int32)
    rp, ci, va = [np.zeros(1, dtype=np.int64)], [], []
    noff = zoff = 0
    for i, A in enumerate(mats):
        A = sp.csr_matrix(A); A.sort_indices()
        n = A.shape[0]
        rp.append(A.indptr[1:].astype(np.int64) + zoff)
        ci.append(A.indices.astype(np.int64) + noff)
        va.append(A.data.astype(np.float64))
        noff += n; zoff += A.nnz
        graph_off[i + 1] = noff
    return (graph_off, np.concatenate(rp).astype(np.int32), np.concatenate(ci).astype(np.int32),
            np.concatenate(va))
