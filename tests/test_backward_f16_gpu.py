"""GPU parity of the tensor-core VJP (csrc/cheb_backward_f16.cu): one 32 -> 32 ChebConv layer, 2 <= K <= 10, binary
operator, graphs of <= 128 nodes, no input gradient - through the C-ABI (mho_cheb_backward), against the fp64 oracle.

Tolerance: 2e-5 relative (the VJP tests' GTOL: sums over all nodes of a graph), per graph and PER PARAMETER BLOCK
(dW_0 .. dW_K-1, db): |g - g_ref|_inf / |g_ref|_inf of the block - dW_K-1 is orders of magnitude larger than dW_0 on a
raw adjacency, a whole-vector norm would hide the small blocks."""
import numpy as np
import pytest
import scipy.sparse as sp

import chebnet_oracle as O
from helpers import random_weights

pytestmark = pytest.mark.gpu
GTOL = 2e-5


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


def _oracle(mats, X, ws, act, dY, slope=0.2):
    gs, o = [], 0
    for A in mats:
        n = A.shape[0]
        _, cache = O.cheb_stack_forward(A, X[o:o + n], ws, [act], slope, return_cache=True)
        g, _ = O.cheb_stack_backward(A, ws, cache, dY[o:o + n], slope)
        gs.append(O.flatten_params(g))
        o += n
    return np.stack(gs)


def _run(torch, K, act, ws, mats, X, dY):
    from multihop_offload_b200 import ChebNet, GraphBatch, LayerSpec
    net = ChebNet([LayerSpec(K, 32, 32, act, 0.2)], device="cuda:0")
    net.set_weights(ws)
    batch = GraphBatch.from_scipy(mats, device="cuda:0")
    Xd = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)).cuda()
    Y, saved = net.forward(batch, Xd, save=True)
    before = net.ctx.launch_count()
    gpg, gsum, _ = net.backward(batch, Xd, Y, saved, torch.from_numpy(np.ascontiguousarray(dY, dtype=np.float32)).cuda())
    torch.cuda.synchronize()
    assert net.ctx.launch_count() - before == 2   # VJP kernel + the one-launch deterministic sum
    return gpg.cpu().numpy(), gsum.cpu().numpy()


def _block_err(g, ref, K):
    """worst relative inf-norm error over the blocks dW_0 .. dW_K-1, db of one flat gradient"""
    worst, where = 0.0, None
    for k in range(K + 1):
        a, b = k * 1024, (k * 1024 + 1024 if k < K else K * 1024 + 32)
        den = max(np.abs(ref[a:b]).max(), 1e-300)
        e = np.abs(g[a:b].astype(np.float64) - ref[a:b]).max() / den
        if e > worst:
            worst, where = e, k
    return worst, where


def _check(gpg, gsum, ref, K, tag):
    for gi in range(ref.shape[0]):
        e, k = _block_err(gpg[gi], ref[gi], K)
        assert e < GTOL, (tag, "graph", gi, "block", k, e)
    den = np.abs(ref.sum(0)).max()
    assert np.abs(gsum - ref.sum(0)).max() / den < GTOL, tag
    # the sum kernel adds exactly the per-graph vectors it was given (fixed order, fp32)
    assert np.abs(gsum - gpg.astype(np.float64).sum(0)).max() / den < 1e-6, tag


@pytest.mark.parametrize("K", [2, 3, 4, 5, 6, 8, 10])
def test_all_orders(torch_cuda, K):
    """Random BA batches, more graphs than CTA slots for the small orders (several graphs per CTA pass), every activation."""
    rng = np.random.default_rng(300 + K)
    sizes = rng.integers(3, 129, size=(700 if K == 5 else 90))
    mats = O.make_batch(sizes, seed0=7000 + K)
    n = int(sizes.sum())
    for act in (O.ACT_LEAKY, O.ACT_RELU, O.ACT_NONE):
        from multihop_offload_b200 import LayerSpec
        ws = random_weights([LayerSpec(K, 32, 32, act, 0.2)], rng, bias=0.2)
        X = rng.normal(size=(n, 32))
        dY = rng.normal(size=(n, 32))
        gpg, gsum = _run(torch_cuda, K, act, ws, mats, X, dY)
        _check(gpg, gsum, _oracle(mats, X, ws, act, dY), K, (K, act))


def test_magnitudes_and_degenerate_graphs(torch_cuda):
    """One power-of-two scale per graph and step: graphs of wildly different magnitude next to each other, rows of very
    different magnitude inside a graph, single nodes, two-node paths, a star (one hub of degree n - 1), a full 128-node graph."""
    from multihop_offload_b200 import LayerSpec
    rng = np.random.default_rng(77)
    K = 5

    def path(n):
        return sp.diags([np.ones(n - 1), np.ones(n - 1)], [-1, 1], format="csr") if n > 1 else sp.csr_matrix((1, 1))

    def star(n):
        A = sp.lil_matrix((n, n)); A[0, 1:] = 1; A[1:, 0] = 1
        return A.tocsr()

    full = sp.csr_matrix(np.ones((128, 128)) - np.eye(128))
    mats = [path(1), path(2), path(3), star(128), full, path(128), star(17), path(1)] + O.make_batch([128, 100, 20, 64], seed0=9)
    sizes = [m.shape[0] for m in mats]
    n = int(sum(sizes))
    X = rng.normal(size=(n, 32))
    dY = rng.normal(size=(n, 32))
    off = np.concatenate([[0], np.cumsum(sizes)])
    for gi, s in enumerate([1e-6, 1e6, 1.0, 1e3, 1e-3, 1e5, 1e-5, 1.0, 1.0, 1e4, 1e-4, 1.0]):
        X[off[gi]:off[gi + 1]] *= s
        dY[off[gi]:off[gi + 1]] *= 1.0 / s if gi % 2 else s
    X[off[8]:off[8] + 10] *= 1e-4   # rows of very different magnitude inside one graph
    dY[off[9]:off[9] + 7] *= 1e4
    X[off[10]:off[11]] = 0.0        # an all-zero input
    ws = random_weights([LayerSpec(K, 32, 32, O.ACT_LEAKY, 0.2)], rng, bias=0.2)
    gpg, gsum = _run(torch_cuda, K, O.ACT_LEAKY, ws, mats, X, dY)
    ref = _oracle(mats, X, ws, O.ACT_LEAKY, dY)
    for gi in range(len(mats)):
        e, k = _block_err(gpg[gi], ref[gi], K)
        if np.abs(ref[gi]).max() == 0.0:
            assert np.abs(gpg[gi]).max() == 0.0, gi
            continue
        # blocks that are exactly zero in the reference (single nodes: T_k = 0 for odd k) must be exactly zero here
        for kk in range(K):
            if np.abs(ref[gi][kk * 1024:(kk + 1) * 1024]).max() == 0.0:
                assert np.abs(gpg[gi][kk * 1024:(kk + 1) * 1024]).max() == 0.0, (gi, kk)
        assert e < GTOL, ("graph", gi, "block", k, e)


def test_repeatable_and_independent_of_grid(torch_cuda):
    """Bit-identical gradients run to run (fixed summation orders, no atomics on data)."""
    from multihop_offload_b200 import LayerSpec
    rng = np.random.default_rng(5)
    sizes = rng.integers(20, 111, size=400)
    mats = O.make_batch(sizes, seed0=123)
    n = int(sizes.sum())
    ws = random_weights([LayerSpec(5, 32, 32, O.ACT_LEAKY, 0.2)], rng)
    X, dY = rng.normal(size=(n, 32)), rng.normal(size=(n, 32))
    a, sa = _run(torch_cuda, 5, O.ACT_LEAKY, ws, mats, X, dY)
    b, sb = _run(torch_cuda, 5, O.ACT_LEAKY, ws, mats, X, dY)
    assert np.array_equal(a, b) and np.array_equal(sa, sb)


def test_long_pipelines_many_graphs_per_cta(torch_cuda):
    """4096 graphs = fourteen graphs per CTA pass through the VJP kernel's role hand-offs; every 16th graph against the
    fp64 oracle, the deterministic sum against the sum of the per-graph rows, and bit-identical on a second run."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from multihop_offload_b200 import ChebNet, GraphBatch, LayerSpec
    torch = torch_cuda
    w = bench.make_workload(4096, 0, None, K=5)
    net = ChebNet([LayerSpec(5, 32, 32, 2, 0.2)], device="cuda:0")
    ws = [(w["W"], w["b"])]
    net.set_weights(ws)
    batch = GraphBatch(w["graph_off"], w["rowptr"], w["colidx"], None, device="cuda:0")
    Xd = torch.from_numpy(w["X"]).cuda()
    rng = np.random.default_rng(3)
    dY = rng.normal(size=w["X"].shape).astype(np.float32)
    Y, saved = net.forward(batch, Xd, save=True)
    g1, s1, _ = net.backward(batch, Xd, Y, saved, torch.from_numpy(dY).cuda())
    g2, s2, _ = net.backward(batch, Xd, Y, saved, torch.from_numpy(dY).cuda())
    torch.cuda.synchronize()
    assert torch.equal(g1, g2) and torch.equal(s1, s2)
    g1 = g1.cpu().numpy()
    assert np.abs(s1.cpu().numpy() - g1.astype(np.float64).sum(0)).max() / np.abs(g1.astype(np.float64).sum(0)).max() < 1e-6
    import scipy.sparse as sp
    off, rp, ci = w["graph_off"], w["rowptr"], w["colidx"]
    for gi in range(0, 4096, 16):
        a, b = int(off[gi]), int(off[gi + 1])
        n = b - a
        indptr = rp[a:b + 1] - rp[a]
        A = sp.csr_matrix((np.ones(int(indptr[-1])), ci[rp[a]:rp[b]] - a, indptr), shape=(n, n))
        ref = _oracle([A], w["X"][a:b].astype(np.float64), ws, O.ACT_LEAKY, dY[a:b].astype(np.float64))[0]
        e, k = _block_err(g1[gi], ref, 5)
        assert e < GTOL, (gi, k, e)
