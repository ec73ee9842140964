"""GPU parity: libmho forward (through the C-ABI) vs the fp64 oracle / committed goldens.

Tolerance (north_star): 1e-5 relative, measured per graph as |y - y_ref|_inf / |y_ref|_inf.
"""
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp

import chebnet_oracle as O
from helpers import numpy_fp32_forward, oracle_batch_forward, random_weights, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _net(specs, weights):
    from multihop_offload_b200 import ChebNet
    net = ChebNet(specs, device="cuda:0")
    net.set_weights(weights)
    return net


def _run(torch, net, mats, X, tile_rows=128, binary=None):
    from multihop_offload_b200 import GraphBatch
    batch = GraphBatch.from_scipy(mats, tile_rows=tile_rows, binary=binary, device="cuda:0")
    Xd = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)).cuda()
    Y = net.forward(batch, Xd)
    torch.cuda.synchronize()
    return Y.cpu().numpy(), batch


def test_golden_layer_k5_f32(torch_cuda, golden_dir):
    from multihop_offload_b200 import GraphBatch, LayerSpec
    z = np.load(os.path.join(golden_dir, "layer_K5_F32.npz"))
    net = _net([LayerSpec(5, 32, 32, O.ACT_LEAKY, 0.2)], [(z["W"], z["b"])])
    for tile_rows in (128, 64, 256, 512):
        batch = GraphBatch(z["graph_off"], z["rowptr"], z["colidx"], None, tile_rows=tile_rows, device="cuda:0")
        Y = net.forward(batch, torch_cuda.from_numpy(z["X"].astype(np.float32)).cuda()).cpu().numpy()
        err = rel_err(Y, z["Y"], z["graph_off"])
        assert err < TOL, (tile_rows, err)


def test_golden_rollout_cases_shipped_checkpoint(torch_cuda, golden_dir):
    from multihop_offload_b200 import reference_stack
    for tag, key, K in (("BAT800", "lam", 1), ("K3", "lam_K3", 3)):
        w = np.load(os.path.join(golden_dir, "weights_%s.npz" % tag))
        ws = [(w["W%d" % i], w["b%d" % i]) for i in range(5)]
        net = _net(reference_stack(K=K), ws)
        for f in sorted(glob.glob(os.path.join(golden_dir, "case*.npz"))):
            z = np.load(f)
            n = z["X"].shape[0]
            A = sp.csr_matrix((z["vals"], z["colidx"], z["rowptr"]), shape=(n, n))
            Y, _ = _run(torch_cuda, net, [A], z["X"], tile_rows=128)
            err = rel_err(Y, z[key])
            assert err < TOL, (tag, os.path.basename(f), err)
            assert (Y >= 0).all()


def test_random_batches_vs_oracle(torch_cuda):
    from multihop_offload_b200 import LayerSpec, reference_stack
    rng = np.random.default_rng(42)
    configs = [
        ([LayerSpec(5, 32, 32)], "raw-adj"),
        ([LayerSpec(1, 32, 32)], "raw-adj"),
        ([LayerSpec(2, 32, 32, O.ACT_RELU)], "raw-adj"),
        ([LayerSpec(10, 32, 32)], "cheb-lap"),
        ([LayerSpec(3, 4, 32), LayerSpec(3, 32, 1, O.ACT_RELU)], "raw-adj"),
        ([LayerSpec(4, 7, 13), LayerSpec(2, 13, 20, O.ACT_NONE), LayerSpec(3, 20, 5, O.ACT_RELU)], "cheb-lap"),
        (reference_stack(K=5), "cheb-lap"),
    ]
    for specs, op in configs:
        sizes = rng.choice([20, 30, 40, 50, 60, 70, 80, 90, 100, 110], size=37)
        mats = O.make_batch(sizes, seed0=int(rng.integers(1 << 20)), operator=op)
        X = rng.normal(size=(int(sizes.sum()), specs[0].f_in))
        ws = random_weights(specs, rng, scale=1.0 if op == "cheb-lap" else 0.5)
        net = _net(specs, ws)
        Y, batch = _run(torch_cuda, net, mats, X)
        ref, zscale = oracle_batch_forward(mats, X, ws, [s.act for s in specs], 0.2, return_scale=True)
        err = rel_err(Y, ref, batch.graph_off, zscale)
        err32 = rel_err(numpy_fp32_forward(mats, X, ws, [s.act for s in specs], 0.2), ref, batch.graph_off, zscale)
        print([(s.K, s.f_in, s.f_out) for s in specs], op, "err", err, "numpy-fp32 err", err32)
        assert err < TOL, ([(s.K, s.f_in, s.f_out) for s in specs], op, err, err32)


def test_edge_cases(torch_cuda):
    from multihop_offload_b200 import GraphBatch, LayerSpec
    rng = np.random.default_rng(7)
    specs = [LayerSpec(4, 32, 32)]
    ws = random_weights(specs, rng, 0.5)
    net = _net(specs, ws)
    # empty batch
    b0 = GraphBatch(np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(0, np.int32), device="cuda:0")
    Y0 = net.forward(b0, torch_cuda.empty((0, 32), dtype=torch_cuda.float32, device="cuda"))
    assert Y0.shape == (0, 32)
    # ragged: single-node graph without edges, empty rows, a 2-node graph, the largest supported graphs
    mats = [sp.csr_matrix((1, 1)), sp.csr_matrix(np.array([[0., 1.], [1., 0.]])), O.ba_adjacency(3, 2, 1),
            O.ba_adjacency(300, 2, 2), O.ba_adjacency(512, 2, 3), sp.csr_matrix((5, 5)), O.ba_adjacency(17, 2, 4)]
    X = rng.normal(size=(sum(m.shape[0] for m in mats), 32))
    for tr in (128, 512):
        Y, batch = _run(torch_cuda, net, mats, X, tile_rows=tr)
        ref = oracle_batch_forward(mats, X, ws, [s.act for s in specs], 0.2)
        assert rel_err(Y, ref, batch.graph_off) < TOL
    # dense-ish hub graph whose CSR slice cannot be staged next to the tiles (falls back to L1/L2 reads)
    hub = sp.random(500, 500, 0.2, random_state=1, format="csr")
    hub.data[:] = rng.normal(size=hub.nnz) * 0.05
    Xh = rng.normal(size=(500, 32))
    Y, batch = _run(torch_cuda, net, [hub], Xh, tile_rows=512, binary=False)
    assert rel_err(Y, oracle_batch_forward([hub], Xh, ws, [s.act for s in specs], 0.2)) < TOL


def test_non_symmetric_weighted_operator(torch_cuda):
    from multihop_offload_b200 import LayerSpec
    rng = np.random.default_rng(9)
    specs = [LayerSpec(5, 32, 32), LayerSpec(3, 32, 8, O.ACT_RELU)]
    ws = random_weights(specs, rng, 0.7)
    mats = []
    for i in range(20):
        n = int(rng.integers(5, 120))
        A = sp.random(n, n, min(1.0, 6.0 / n), random_state=i, format="csr")
        A.data[:] = rng.uniform(-0.3, 0.3, size=A.nnz)
        mats.append(A)
    X = rng.normal(size=(sum(m.shape[0] for m in mats), 32))
    Y, batch = _run(torch_cuda, _net(specs, ws), mats, X, binary=False)
    ref = oracle_batch_forward(mats, X, ws, [s.act for s in specs], 0.2)
    assert rel_err(Y, ref, batch.graph_off) < TOL


def test_host_buffer_api(torch_cuda, golden_dir):
    from multihop_offload_b200 import LayerSpec
    z = np.load(os.path.join(golden_dir, "layer_K5_F32.npz"))
    net = _net([LayerSpec(5, 32, 32, O.ACT_LEAKY, 0.2)], [(z["W"], z["b"])])
    Y = net.forward_host(z["graph_off"], z["rowptr"], z["colidx"], None, z["X"])
    assert rel_err(Y, z["Y"], z["graph_off"]) < TOL


def test_full_size_linearity_property(torch_cuda):
    """BASELINE config 2 size (1024 graphs): with act=none the layer is linear in X, and
    permuting the graph order permutes the output blocks - size-independent properties."""
    from multihop_offload_b200 import GraphBatch, LayerSpec
    rng = np.random.default_rng(0)
    sizes = rng.choice(np.arange(20, 111, 10), size=1024)
    mats = O.make_batch(sizes, seed0=1000)
    specs = [LayerSpec(5, 32, 32, O.ACT_NONE)]
    ws = random_weights(specs, rng, 0.3, bias=0.0)
    net = _net(specs, ws)
    batch = GraphBatch.from_scipy(mats, device="cuda:0")
    n = batch.total_nodes
    X1 = torch_cuda.randn(n, 32, device="cuda"); X2 = torch_cuda.randn(n, 32, device="cuda")
    Y1 = net.forward(batch, X1); Y2 = net.forward(batch, X2)
    Y12 = net.forward(batch, (2.0 * X1 - 0.5 * X2).contiguous())
    scale = Y12.abs().max().item()
    assert (Y12 - (2.0 * Y1 - 0.5 * Y2)).abs().max().item() <= 2e-5 * scale
    # graph-order permutation
    perm = rng.permutation(len(mats))
    batch_p = GraphBatch.from_scipy([mats[i] for i in perm], device="cuda:0")
    off = batch.graph_off
    idx = np.concatenate([np.arange(off[i], off[i + 1]) for i in perm])
    Yp = net.forward(batch_p, X1[torch_cuda.from_numpy(idx).cuda()].contiguous())
    # (not bit-equal: hub rows that straddle stream segments are summed in a position-dependent order)
    Yref = Y1[torch_cuda.from_numpy(idx).cuda()]
    assert (Yp - Yref).abs().max().item() <= 2e-5 * Yref.abs().max().item()
    # and a sample of graphs against the oracle at full batch size
    Xh = X1.cpu().numpy().astype(np.float64)
    for gi in rng.choice(len(mats), size=16, replace=False):
        a, b = off[gi], off[gi + 1]
        ref = O.cheb_stack_forward(mats[gi], Xh[a:b], ws, [O.ACT_NONE])
        assert rel_err(Y1[a:b].cpu().numpy(), ref) < TOL


def test_host_api_multichunk_pipeline_equals_device_api(torch_cuda):
    """1024 graphs -> ~600 tiles -> the host call runs as pipelined chunks of ~300 tiles (upload / kernel / download
    streams); the result must equal the single-launch device path bit for bit (same tiles, same order of
    floating-point operations inside every tile), with and without operator values."""
    from multihop_offload_b200 import GraphBatch, LayerSpec
    rng = np.random.default_rng(3)
    sizes = rng.choice(np.arange(20, 111, 10), size=1024)
    mats = O.make_batch(sizes, seed0=5000)
    for weighted in (False, True):
        if weighted:
            mats = [sp.csr_matrix((rng.uniform(-0.3, 0.3, size=m.nnz), m.indices, m.indptr), shape=m.shape) for m in mats]
        specs = [LayerSpec(3, 32, 32)]
        ws = random_weights(specs, rng, 0.5)
        net = _net(specs, ws)
        batch = GraphBatch.from_scipy(mats, binary=not weighted, device="cuda:0")
        X = rng.normal(size=(batch.total_nodes, 32)).astype(np.float32)
        Yd = net.forward(batch, torch_cuda.from_numpy(X).cuda()).cpu().numpy()
        Yh = net.forward_host(batch.graph_off, batch.rowptr, batch.colidx, batch.vals, X)
        assert np.array_equal(Yd, Yh)
        off = batch.graph_off
        for gi in rng.choice(len(mats), size=8, replace=False):
            ref = O.cheb_stack_forward(mats[gi], X[off[gi]:off[gi + 1]].astype(np.float64), ws, [specs[0].act])
            assert rel_err(Yh[off[gi]:off[gi + 1]], ref) < TOL


def test_weights_change_is_picked_up(torch_cuda):
    """The packed weight images (TF32 hi/lo, bf16 x3, fp16 x2) are cached per context: set_weights must invalidate them
    (the optimizer: tests/test_f16_kernel_gpu.py::test_forward_after_optimizer_replay_uses_new_weights)."""
    from multihop_offload_b200 import GraphBatch, LayerSpec
    rng = np.random.default_rng(4)
    mats = O.make_batch([40, 50], seed0=1)
    specs = [LayerSpec(2, 32, 32)]
    net = _net(specs, random_weights(specs, rng, 0.5))
    batch = GraphBatch.from_scipy(mats, device="cuda:0")
    X = rng.normal(size=(90, 32))
    Xd = torch_cuda.from_numpy(X.astype(np.float32)).cuda()
    for _ in range(3):
        ws = random_weights(specs, rng, 0.5)
        net.set_weights(ws)
        Y = net.forward(batch, Xd).cpu().numpy()
        assert rel_err(Y, oracle_batch_forward(mats, X, ws, [s.act for s in specs], 0.2), batch.graph_off) < TOL


def test_dense_kernel_bit_rows_vs_csr_input(torch_cuda, golden_dir):
    """The tensor-core forward takes the operator either as precomputed bit rows (mho_batch_t.adj_bits) or derives
    the bits from the CSR slice in the kernel: same bits, same arithmetic -> identical results; both match the golden."""
    from multihop_offload_b200 import GraphBatch, LayerSpec
    z = np.load(os.path.join(golden_dir, "layer_K5_F32.npz"))
    net = _net([LayerSpec(5, 32, 32, O.ACT_LEAKY, 0.2)], [(z["W"], z["b"])])
    X = torch_cuda.from_numpy(z["X"].astype(np.float32)).cuda()
    batch = GraphBatch(z["graph_off"], z["rowptr"], z["colidx"], None, tile_rows=128, device="cuda:0")
    assert "adj_bits" in batch.dev
    Y_bits = net.forward(batch, X).cpu().numpy()
    del batch.dev["adj_bits"]
    batch._struct_cache = {}
    Y_csr = net.forward(batch, X).cpu().numpy()
    assert np.array_equal(Y_bits, Y_csr)
    assert rel_err(Y_bits, z["Y"], z["graph_off"]) < TOL


def test_host_api_async_pipeline(torch_cuda):
    """mho_cheb_forward_host_async / mho_host_wait: several calls through the two staging slots, each with its own
    page-locked buffers and its own inputs, give exactly the blocking call's results."""
    from multihop_offload_b200 import LayerSpec
    from multihop_offload_b200._lib import PinnedArray, pinned_like
    rng = np.random.default_rng(21)
    mats = O.make_batch(rng.choice(np.arange(20, 111, 10), size=700), seed0=7000)
    g_off, rowptr, colidx, vals = O.concat_batch(mats)
    specs = [LayerSpec(5, 32, 32, O.ACT_LEAKY, 0.2)]
    net = _net(specs, random_weights(specs, rng, 0.5))
    n = int(g_off[-1])
    g, r, c = (pinned_like(np.ascontiguousarray(a, dtype=np.int32)) for a in (g_off, rowptr, colidx))
    Xs = [pinned_like(rng.standard_normal((n, 32)).astype(np.float32)) for _ in range(5)]
    Ys = [PinnedArray((n, 32), np.float32) for _ in range(5)]
    refs = [net.forward_host(g.array, r.array, c.array, None, X.array).copy() for X in Xs]
    tickets = [net.forward_host_async(g.array, r.array, c.array, None, Xs[i].array, Ys[i].array) for i in range(5)]
    for t in tickets:
        net.host_wait(t)
    for i in range(5):
        assert np.array_equal(Ys[i].array, refs[i]), i


def test_dense_kernel_edge_cases(torch_cuda):
    """Tensor-core forward (binary operator, tiles <= 128 nodes): degenerate graphs, a full 128-node tile, a complete
    graph (every adjacency bit set), unaligned / narrow feature widths, fused stacks with saved activations."""
    from multihop_offload_b200 import GraphBatch, LayerSpec
    rng = np.random.default_rng(77)
    full = np.ones((128, 128)) - np.eye(128)
    mats = [sp.csr_matrix((1, 1)), sp.csr_matrix(np.array([[0., 1.], [1., 0.]])), sp.csr_matrix((5, 5)),
            O.ba_adjacency(128, 2, 11), sp.csr_matrix(full), O.ba_adjacency(3, 2, 1), O.ba_adjacency(127, 2, 12),
            sp.csr_matrix((1, 1)), O.ba_adjacency(64, 2, 13), O.ba_adjacency(64, 2, 14)]
    n = sum(m.shape[0] for m in mats)
    for specs, scale in (([LayerSpec(5, 32, 32, O.ACT_LEAKY, 0.2)], 0.05),
                         ([LayerSpec(2, 7, 13, O.ACT_RELU), LayerSpec(3, 13, 32, O.ACT_NONE), LayerSpec(1, 32, 1, O.ACT_RELU)], 0.1),
                         ([LayerSpec(4, 4, 32), LayerSpec(1, 32, 32), LayerSpec(5, 32, 8, O.ACT_RELU)], 0.05)):
        ws = random_weights(specs, rng, scale)
        net = _net(specs, ws)
        X = rng.normal(size=(n, specs[0].f_in))
        batch = GraphBatch.from_scipy(mats, tile_rows=128, device="cuda:0")
        assert batch.adj_bits is not None and batch.max_tile_rows <= 128
        Xd = torch_cuda.from_numpy(X.astype(np.float32)).cuda()
        Y, saved = net.forward(batch, Xd, save=True)
        ref, zscale = oracle_batch_forward(mats, X, ws, [s.act for s in specs], 0.2, return_scale=True)
        assert rel_err(Y.cpu().numpy(), ref, batch.graph_off, zscale) < TOL, [(s.K, s.f_in, s.f_out) for s in specs]
        if len(specs) > 1:   # saved hidden activations = the next layer's inputs (what the VJP consumes)
            acts = [s.act for s in specs]
            h1 = oracle_batch_forward(mats, X, ws[:1], acts[:1], 0.2)
            got = saved[: n * specs[1].f_in].view(n, specs[1].f_in).cpu().numpy()
            assert rel_err(got, h1, batch.graph_off) < TOL


def test_two_streams_two_contexts_equal_single_stream(torch_cuda, golden_dir):
    """Independent batches on two CUDA streams, one private library context each (tile counters, weight images are
    per context): the overlapped launches give exactly the results of launches issued one after the other."""
    from multihop_offload_b200 import ChebNet, GraphBatch, LayerSpec
    rng = np.random.default_rng(31)
    z = np.load(os.path.join(golden_dir, "layer_K5_F32.npz"))
    specs = [LayerSpec(5, 32, 32, O.ACT_LEAKY, 0.2)]
    nets = [ChebNet(specs, device="cuda:0", private_context=True) for _ in range(2)]
    for n_ in nets:
        n_.set_weights([(z["W"], z["b"])])
    assert nets[0].ctx is not nets[1].ctx
    mats = O.make_batch(rng.choice(np.arange(20, 111, 10), size=900), seed0=9000)
    batch = GraphBatch.from_scipy(mats, device="cuda:0")
    Xs = [torch_cuda.randn((batch.total_nodes, 32), device="cuda") for _ in range(6)]
    want = [nets[0].forward(batch, X).clone() for X in Xs]
    torch_cuda.cuda.synchronize()
    streams = [torch_cuda.cuda.Stream() for _ in range(2)]
    outs = [torch_cuda.empty_like(want[0]) for _ in Xs]
    for rep in range(3):
        for i, X in enumerate(Xs):
            with torch_cuda.cuda.stream(streams[i & 1]):
                nets[i & 1].forward(batch, X, out=outs[i])
    torch_cuda.cuda.synchronize()
    for i in range(len(Xs)):
        assert torch_cuda.equal(outs[i], want[i]), i


def test_dense_kernel_streams_weights_of_deep_k_stacks(torch_cuda):
    """Five-layer stacks with K = 3 / 4 / 5: the bf16 weight images of all layers do not fit next to the tiles, so the
    tensor-core kernel streams them layer by layer through two shared-memory slots (many tiles per CTA: the slots are
    recycled across tiles); hidden activations are saved for the VJP."""
    from multihop_offload_b200 import GraphBatch, reference_stack
    rng = np.random.default_rng(91)
    sizes = rng.choice(np.arange(20, 111, 10), size=700)
    mats = O.make_batch(sizes, seed0=12000)
    n = int(sizes.sum())
    batch = GraphBatch.from_scipy(mats, tile_rows=128, device="cuda:0")
    assert batch.n_tiles > 2 * 296 // 2   # several tiles per CTA
    for K in (3, 4, 5):
        specs = reference_stack(K=K)
        ws = random_weights(specs, rng, 0.15 if K < 5 else 0.06)
        net = _net(specs, ws)
        X = rng.normal(size=(n, 4))
        Y, saved = net.forward(batch, torch_cuda.from_numpy(X.astype(np.float32)).cuda(), save=True)
        acts = [s.act for s in specs]
        ref, zscale = oracle_batch_forward(mats, X, ws, acts, 0.2, return_scale=True)
        err = rel_err(Y.cpu().numpy(), ref, batch.graph_off, zscale)
        # five raw-adjacency layers of order K amplify round-off (values grow by (2 rho)^K per layer): the yardstick is
        # what plain fp32 arithmetic (numpy) loses on the same stack
        err32 = rel_err(numpy_fp32_forward(mats, X, ws, acts, 0.2), ref, batch.graph_off, zscale)
        print("K", K, "err", err, "numpy-fp32 err", err32)
        assert err < max(TOL, 1.5 * err32), (K, err, err32)   # measured table: profiles/r2_deep_stack_errors.json
        h3 = oracle_batch_forward(mats, X, ws[:3], acts[:3], 0.2)
        off = n * (32 + 32)
        got = saved[off: off + n * 32].view(n, 32).cpu().numpy()
        assert rel_err(got, h3, batch.graph_off) < TOL, K
