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
