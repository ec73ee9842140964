"""GPU parity of the fused K = 1 stack kernel (csrc/cheb_mlp_f16.cu: every layer K = 1, <= 32 features - the model the
reference ships, gnn_offloading_agent.py:81-123 with Spektral's default K) through the C-ABI, against the fp64 oracle.

Tolerance: 1e-5 relative per graph (|y - y_ref|_inf / max(|y_ref|_inf, |z_ref|_inf)); the activations kept for the VJP
(`saved`) are held to the same bound against the oracle's layer inputs."""
import numpy as np
import pytest

import chebnet_oracle as O
from helpers import oracle_batch_forward, random_weights, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


def _specs(dims, acts):
    from multihop_offload_b200 import LayerSpec
    return [LayerSpec(1, dims[i], dims[i + 1], acts[i], 0.2) for i in range(len(dims) - 1)]


@pytest.mark.parametrize("dims,acts", [
    ((4, 32, 32, 32, 32, 1), (O.ACT_LEAKY,) * 4 + (O.ACT_RELU,)),       # the shipped model
    ((4, 32, 1), (O.ACT_LEAKY, O.ACT_NONE)),
    ((8, 16, 32, 8), (O.ACT_RELU, O.ACT_LEAKY, O.ACT_NONE)),            # <= 16 inputs: one K slice; 8 outputs: direct store
    ((32, 32), (O.ACT_LEAKY,)),                                          # one layer, whole 128 B output lines
    ((12, 20, 28, 32, 3), (O.ACT_LEAKY, O.ACT_RELU, O.ACT_LEAKY, O.ACT_NONE)),
])
def test_k1_stacks(torch_cuda, dims, acts):
    from multihop_offload_b200 import ChebNet, GraphBatch
    torch = torch_cuda
    rng = np.random.default_rng(sum(dims))
    sizes = rng.integers(1, 129, size=300)            # 128-row tiles ignore graph boundaries: ragged on purpose
    mats = O.make_batch(np.maximum(sizes, 3), seed0=11)
    n = int(sum(m.shape[0] for m in mats))
    specs = _specs(dims, acts)
    ws = random_weights(specs, rng, bias=0.3)
    X = rng.normal(size=(n, dims[0]))
    X[: n // 7] *= 1e4                                  # rows of very different magnitude (the scale is per row)
    X[n // 7: n // 5] *= 1e-4
    X[n // 2] = 0.0
    net = ChebNet(specs, device="cuda:0")
    net.set_weights(ws)
    batch = GraphBatch.from_scipy(mats, device="cuda:0")
    Xd = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)).cuda()
    before = net.ctx.launch_count()
    Y = net.forward(batch, Xd)
    Y2, saved = net.forward(batch, Xd, save=True)
    torch.cuda.synchronize()
    assert net.ctx.launch_count() - before <= 3        # weight image once + one fused launch per forward
    ref, zs = oracle_batch_forward(mats, X, ws, list(acts), 0.2, return_scale=True)
    err = rel_err(Y.cpu().numpy(), ref, batch.graph_off, zs)
    assert err < TOL, (dims, err)
    assert torch.equal(Y, Y2), "keeping the activations must not change the output"
    if len(dims) > 2:
        # saved = the inputs of layers 1.. (row-major [n, f_in_l], concatenated)
        sv = saved.cpu().numpy().ravel()
        o = 0
        h = X
        for li, ((W, b), act) in enumerate(zip(ws[:-1], acts[:-1])):
            h = O._act(h @ W[0] + b, act, 0.2)
            blk = sv[o:o + n * dims[li + 1]].reshape(n, dims[li + 1])
            # per row, relative to the larger of the activation and the pre-activation scale of that row
            den = np.maximum(np.abs(h).max(axis=1), 1e-30)
            pre = np.abs(h).max(axis=1) if act != O.ACT_RELU else np.maximum(np.abs(h).max(axis=1), 1e-6 * np.abs(h).max())
            assert (np.abs(blk - h).max(axis=1) / np.maximum(den, pre)).max() < 2e-5, (dims, li)
            o += n * dims[li + 1]


# ------------------------------------------------------------------------------------------------------------------------
# VJP of K = 1 stacks on the tensor cores (csrc/cheb_mlp_backward_f16.cu)
# ------------------------------------------------------------------------------------------------------------------------
GTOL = 2e-5


def _layer_blocks(dims):
    """(start, end) of every layer's block (kernel + bias) in the flat parameter vector.  The bias of a 1-wide layer is ONE number,
    a sum over the graph's nodes that may cancel: judged together with its kernel (both are sums of the same G)"""
    blocks, o = [], 0
    for i in range(len(dims) - 1):
        n = dims[i] * dims[i + 1] + dims[i + 1]
        blocks.append((o, o + n)); o += n
    return blocks


@pytest.mark.parametrize("dims,acts", [
    ((4, 32, 32, 32, 32, 1), (O.ACT_LEAKY,) * 4 + (O.ACT_RELU,)),       # the shipped model
    ((8, 32, 32, 2), (O.ACT_RELU, O.ACT_LEAKY, O.ACT_NONE)),
    ((32, 32, 4), (O.ACT_LEAKY, O.ACT_LEAKY)),
    ((12, 32, 32, 32, 32, 32, 1), (O.ACT_LEAKY, O.ACT_NONE, O.ACT_RELU, O.ACT_LEAKY, O.ACT_LEAKY, O.ACT_NONE)),
])
def test_k1_stack_vjp(torch_cuda, dims, acts):
    """Per-graph gradients of K = 1 stacks (no input gradient: the tensor-core path) against the fp64 oracle, per graph and per
    parameter block; the deterministic sum; bit-identical on a second run."""
    from multihop_offload_b200 import ChebNet, GraphBatch
    torch = torch_cuda
    rng = np.random.default_rng(7 + sum(dims))
    sizes = np.concatenate([rng.integers(3, 129, size=420), [128, 127, 3, 16, 17]])
    mats = O.make_batch(sizes, seed0=77)
    n = int(sizes.sum())
    specs = _specs(dims, acts)
    ws = random_weights(specs, rng, scale=1.5, bias=0.3)
    X = rng.normal(size=(n, dims[0]))
    dY = rng.normal(size=(n, dims[-1]))
    off = np.concatenate([[0], np.cumsum(sizes)])
    for gi, sc in enumerate([1e-5, 1e5, 1e3, 1e-3]):        # graphs of very different magnitude
        X[off[gi]:off[gi + 1]] *= sc
        dY[off[gi]:off[gi + 1]] *= 1.0 / sc if gi % 2 else sc
    net = ChebNet(specs, device="cuda:0")
    net.set_weights(ws)
    batch = GraphBatch.from_scipy(mats, device="cuda:0")
    Xd = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)).cuda()
    dYd = torch.from_numpy(np.ascontiguousarray(dY, dtype=np.float32)).cuda()
    Y, saved = net.forward(batch, Xd, save=True)
    g1, s1, _ = net.backward(batch, Xd, Y, saved, dYd)
    before = net.ctx.launch_count()
    g2, s2, _ = net.backward(batch, Xd, Y, saved, dYd)
    torch.cuda.synchronize()
    assert net.ctx.launch_count() - before in (2, 3)  # VJP kernel + the sum (one launch when the parameter count is a multiple of 4)
    assert torch.equal(g1, g2) and torch.equal(s1, s2)
    g1 = g1.cpu().numpy()
    blocks = _layer_blocks(dims)
    Yh = Y.cpu().numpy().astype(np.float64)
    worst = 0.0
    for gi, A in enumerate(mats):
        a, b = int(off[gi]), int(off[gi + 1])
        _, cache = O.cheb_stack_forward(A, X[a:b], ws, list(acts), 0.2, return_cache=True)
        gr, _ = O.cheb_stack_backward(A, ws, cache, dY[a:b], 0.2)
        ref = O.flatten_params(gr)
        for (p0, p1) in blocks:
            den = np.abs(ref[p0:p1]).max()
            if den == 0.0:
                assert np.abs(g1[gi][p0:p1]).max() == 0.0, (dims, gi, p0)
                continue
            e = np.abs(g1[gi][p0:p1] - ref[p0:p1]).max() / den
            worst = max(worst, e)
            assert e < GTOL, (dims, "graph", gi, "block", (p0, p1), e)
    ref_sum = g1.astype(np.float64).sum(0)
    assert np.abs(s1.cpu().numpy() - ref_sum).max() / np.abs(ref_sum).max() < 1e-6
