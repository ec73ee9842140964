"""GPU: the drop-in ACOAgent on libmho - forward (lambda -> delay matrices) and the :448 VJP against the
committed goldens (shipped checkpoint x shipped networks).  The reference simulator is not available on
the GPU box, so obj/env are rebuilt from the golden fixtures (fields of offloading_v3.py:284-333)."""
import glob
import os
import sys
import types

import networkx as nx
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def stub_case(z):
    n = z["X"].shape[0]
    A = sp.csr_matrix((z["vals"], z["colidx"], z["rowptr"]), shape=(n, n))
    obj = types.SimpleNamespace()
    obj.gi_ext = nx.from_scipy_sparse_array(A)
    obj.num_edges_ext = n
    obj.edge_self_loop, obj.edge_rate_ext = z["X"][:, 0], z["X"][:, 1]
    obj.jobs_arrivals, obj.edge_as_server = z["X"][:, 2], z["X"][:, 3]
    obj.maps_ol_el, obj.maps_on_el = z["maps_ol_el"], z["maps_on_el"]
    env = types.SimpleNamespace()
    env.num_nodes, env.T = int(z["num_nodes"]), int(z["T"])
    env.link_rates, env.cf_degs, env.proc_bws = z["link_rates"], z["cf_degs"], z["proc_bws"]
    L = len(z["link_rates"])
    env.num_links = L
    env.adj_i = sp.csr_matrix((z["adj_i_vals"], z["adj_i_colidx"], z["adj_i_rowptr"]), shape=(L, L))
    env.link_matrix = z["link_matrix"]
    g = nx.Graph()
    g.add_nodes_from(range(env.num_nodes))
    g.add_edges_from([tuple(e) for e in z["edges"]])
    env.graph_c = g
    return obj, env


@pytest.fixture()
def agent(golden_dir, monkeypatch):
    monkeypatch.setattr(sys, "argv", ["test"])
    from multihop_offload_b200.gnn_offloading_agent import ACOAgent, FLAGS
    FLAGS.device = "cuda:0"
    FLAGS.K = 1
    FLAGS.fix_diag = False
    a = ACOAgent(FLAGS, 100)
    a.load(os.path.join(golden_dir, "ckpt_BAT800"))
    return a


def test_forward_matches_golden_delay_matrices(agent, golden_dir):
    for f in sorted(glob.glob(os.path.join(golden_dir, "case*.npz"))):
        z = np.load(f)
        obj, env = stub_case(z)
        state, D_ts, D_np = agent.forward(obj, env)
        lam = agent.predict(state).cpu().numpy()
        assert np.abs(lam - z["lam"]).max() <= 1e-5 * np.abs(z["lam"]).max()
        # numpy twin (bug-compatible wrapped diagonal) and TF tensor twin (relays = +inf)
        gold = z["delay_mtx_bug"]
        assert np.array_equal(np.isnan(D_np), np.isnan(gold))
        m = ~np.isnan(gold)
        # the head is 1/(mu - lambda): a 1e-5 relative error of the fp32 GNN output lambda is amplified by
        # d(delay)/d(lambda) = delay^2 near congestion, so the bound is conditioning-aware
        lam_max = np.abs(z["lam"]).max()
        tol = 2e-5 * np.abs(gold[m]) + 2e-5 * lam_max * gold[m] ** 2
        assert (np.abs(D_np[m] - gold[m]) <= tol).all()
        Dt = D_ts.cpu().numpy()
        gt = np.nan_to_num(z["delay_mtx_ts"], nan=0.0, posinf=np.inf)
        fin = np.isfinite(gt)
        assert np.array_equal(np.isinf(Dt), np.isinf(gt))
        assert (np.abs(Dt[fin] - gt[fin]) <= 2e-5 * np.abs(gt[fin]) + 2e-5 * lam_max * gt[fin] ** 2).all()


def test_vjp_from_grad_dist_matches_golden(agent, golden_dir):
    for f in sorted(glob.glob(os.path.join(golden_dir, "case*.npz"))):
        z = np.load(f)
        obj, env = stub_case(z)
        agent.forward(obj, env, save=True)
        g = agent.vjp_from_grad_dist(z["grad_dist"]).cpu().numpy()
        want = z["grad_flat"]
        err = np.abs(g - want).max() / np.abs(want).max()
        # the head's Jacobian d(delay)/d(lambda) = delay^2 is itself evaluated at the fp32 lambda: its relative
        # perturbation is ~ 2 * 1e-5 * lambda * delay (the pure GNN VJP is checked at 2e-5 in test_backward_gpu)
        d_max = max(np.abs(z["link_delay"]).max(), np.abs(z["node_delay"]).max())
        tol = 5e-5 * (1.0 + 2.0 * np.abs(z["lam"]).max() * d_max)
        print(os.path.basename(f), "vjp err", err, "tol", tol)
        assert err < tol, (os.path.basename(f), err, tol)


def test_predict_batch_equals_single_calls(agent, golden_dir):
    states = []
    for f in sorted(glob.glob(os.path.join(golden_dir, "case*.npz"))):
        z = np.load(f)
        obj, env = stub_case(z)
        states.append(agent.makestate(nx.adjacency_matrix(obj.gi_ext), z["X"]))
    outs = agent.predict_batch(states)
    for s, y in zip(states, outs):
        y1 = agent.predict(s)
        assert (y - y1).abs().max().item() <= 1e-6 * max(y1.abs().max().item(), 1e-30)


def test_replay_and_checkpoint_roundtrip_on_device(agent, tmp_path):
    import torch
    rng = np.random.default_rng(0)
    for i in range(5):
        agent.memorize(torch.as_tensor(rng.normal(size=agent.net.n_params).astype(np.float32)).cuda(), float(i), 0.0)
    w0 = agent.optimizer.get_master()
    loss = agent.replay(4)
    assert np.isfinite(loss) and agent.optimizer.iterations == 4
    w1 = agent.optimizer.get_master()
    assert np.abs(w1 - w0).max() > 0
    np.testing.assert_array_equal(agent.net.get_flat().astype(np.float32), w1.astype(np.float32))
    agent.save(str(tmp_path / "m" / "cp-0001.ckpt"))
    from multihop_offload_b200 import tf_bundle
    back = tf_bundle.load_weights(tf_bundle.latest_checkpoint(str(tmp_path / "m")))
    flat = np.concatenate([np.concatenate([W.ravel(), b.ravel()]) for W, b in back])
    np.testing.assert_array_equal(flat, w1)   # fp64 master weights survive bit for bit


def test_fused_queue_head_kernels_match_golden(golden_dir):
    """mho_queue_head_forward / _backward (fp64) vs the oracle's head on the golden cases, batched 6 graphs at once."""
    import torch
    import chebnet_oracle as O
    from multihop_offload_b200 import _lib
    from multihop_offload_b200 import queue_head as qh
    ctx = _lib.Context.get(0)
    cases = [np.load(f) for f in sorted(glob.glob(os.path.join(golden_dir, "case*.npz")))]
    his, sizes, lams = [], [], []
    for z in cases:
        obj, env = stub_case(z)
        his.append(qh.HeadInputs(obj, env, None)); sizes.append(z["X"].shape[0]); lams.append(z["lam"][:, 0])
    hb = qh.HeadBatch(his, sizes, ctx, "cuda:0")
    lam = torch.as_tensor(np.concatenate(lams).astype(np.float32)).cuda()
    ld, nd = hb.forward(lam, save=True)
    ld, nd = ld.cpu().numpy(), nd.cpu().numpy()
    rng = np.random.default_rng(0)
    g_link = rng.normal(size=ld.shape); g_node = rng.normal(size=nd.shape)
    g_lam = hb.backward(torch.as_tensor(g_link).cuda(), torch.as_tensor(g_node).cuda()).cpu().numpy()[:, 0]
    for gi, z in enumerate(cases):
        L = len(z["link_rates"])
        Ai = sp.csr_matrix((z["adj_i_vals"], z["adj_i_colidx"], z["adj_i_rowptr"]), shape=(L, L))
        lam32 = z["lam"].astype(np.float32).astype(np.float64)   # the kernels see the fp32 GNN output
        rld, rnd, hc = O.queue_head_forward(lam32, z["maps_ol_el"], z["maps_on_el"], z["link_rates"], z["cf_degs"],
                                            z["proc_bws"], Ai, int(z["T"]), return_cache=True)
        a, b = hb.link_off[gi], hb.link_off[gi + 1]
        c, d = hb.comp_off[gi], hb.comp_off[gi + 1]
        np.testing.assert_allclose(ld[a:b], rld[:, 0], rtol=1e-12)
        np.testing.assert_allclose(nd[c:d], rnd[:, 0], rtol=1e-12)
        want = O.queue_head_vjp(hc, g_link[a:b], g_node[c:d], lam32.shape[0], z["maps_ol_el"], z["maps_on_el"])[:, 0]
        e0, e1 = hb.ext_off[gi], hb.ext_off[gi + 1]
        assert np.abs(g_lam[e0:e1] - want).max() <= 2e-7 * max(np.abs(want).max(), 1e-30)   # fp32 output


def test_shortest_path_matrices_on_device_equal_dijkstra(agent, golden_dir):
    """ACOAgent._shortest_paths (sp_gnn / sp_hop of gnn_offloading_agent.py:286-287) runs mho_apsp on a CUDA device:
    bit-identical to Dijkstra on the delay matrix the agent just produced; the plan is cached per topology."""
    import apsp_oracle as AO
    f = sorted(glob.glob(os.path.join(golden_dir, "case*.npz")))[0]
    z = np.load(f)
    obj, env = stub_case(z)
    state, D_ts, D_np = agent.forward(obj, env)
    sp_gnn, sp_hop = agent._shortest_paths(env, D_np, None)
    edges = list(env.graph_c.edges)
    w = [D_np[a, b] for (a, b) in edges]
    assert np.array_equal(sp_gnn, AO.apsp_lengths(env.num_nodes, edges, w))
    assert np.array_equal(sp_hop, AO.apsp_lengths(env.num_nodes, edges, None))
    plan = env.__dict__["_mho_apsp"]
    agent._shortest_paths(env, D_np, None)
    assert env.__dict__["_mho_apsp"] is plan


def test_forward_instances_equals_per_instance_launches(agent, golden_dir):
    """ten job instances of one network (same extended line graph, only the arrival column differs) through ONE
    GNN launch + ONE queue-head launch + ONE shortest-path launch give, instance by instance, exactly what ten forward()
    calls + ten APSP launches give (K = 1: per-node MLP; bit-identical)."""
    for f in sorted(glob.glob(os.path.join(golden_dir, "case*.npz")))[:3]:
        z = np.load(f)
        obj, env = stub_case(z)
        rng = np.random.default_rng(5)
        feats, single = [], []
        for i in range(10):
            o = types.SimpleNamespace(**vars(obj))
            o.jobs_arrivals = obj.jobs_arrivals * rng.uniform(0.0, 2.0, size=obj.jobs_arrivals.shape)
            feats.append(agent.instance_features(o))
            _, _, D_np = agent.forward(o, env)
            sp_gnn, sp_hop = agent._shortest_paths(env, D_np, None)
            single.append((D_np, sp_gnn, sp_hop))
        l0 = agent.net.ctx.launch_count()
        pre = agent.forward_instances(obj, env, feats, None)
        launches = agent.net.ctx.launch_count() - l0
        assert launches <= 4, launches      # GNN forward, queue head, weighted APSP (+ hop counts when not cached)
        for (D1, s1, h1), (D2, s2, h2) in zip(single, pre):
            assert np.array_equal(np.isnan(D1), np.isnan(D2))
            m = ~np.isnan(D1)
            assert np.array_equal(D1[m], D2[m])
            assert np.array_equal(s1, s2) and np.array_equal(h1, h2)
