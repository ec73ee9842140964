"""All-pairs shortest path lengths: oracle vs outputs of the reference's own util.all_pairs_shortest_paths
(tests/golden/apsp_cases.npz, oracle/make_golden_apsp.py), and the CUDA kernel vs both - bit-exact."""
import os

import networkx as nx
import numpy as np
import pytest

import apsp_oracle as AO


def _cases(golden_dir):
    z = np.load(os.path.join(golden_dir, "apsp_cases.npz"))
    for i in range(int(z["n_cases"])):
        yield int(z["n%d" % i]), [tuple(int(x) for x in e) for e in z["edges%d" % i]], z["w%d" % i], z["sp_delay%d" % i], z["sp_hop%d" % i]


def test_oracle_matches_reference_outputs(golden_dir):
    for n, edges, w, sp_delay, sp_hop in _cases(golden_dir):
        assert np.array_equal(AO.apsp_lengths(n, edges, w), sp_delay), n
        assert np.array_equal(AO.apsp_lengths(n, edges, None), sp_hop), n


def test_oracle_properties():
    rng = np.random.default_rng(3)
    g = nx.barabasi_albert_graph(40, 2, seed=9)
    edges = list(g.edges)
    w = rng.uniform(0.1, 5.0, size=len(edges))
    d = AO.apsp_lengths(40, edges, w)
    assert np.all(np.diag(d) == 0) and np.allclose(d, d.T, rtol=1e-13)
    assert np.all(np.isfinite(d))   # connected
    for k in range(40):   # triangle inequality up to rounding
        assert np.all(d <= d[:, [k]] + d[[k], :] + 1e-12)
    # two components: +inf across
    d2 = AO.apsp_lengths(4, [(0, 1), (2, 3)], None)
    assert d2[0, 1] == 1 and np.isinf(d2[0, 2]) and np.isinf(d2[3, 1])


@pytest.mark.gpu
def test_kernel_bit_exact_vs_reference_outputs(golden_dir):
    import torch
    from multihop_offload_b200.apsp import ApspPlan
    assert torch.cuda.is_available()
    graphs, mats, want_d, want_h = [], [], [], []
    for n, edges, w, sp_delay, sp_hop in _cases(golden_dir):
        g = nx.Graph()
        g.add_nodes_from(range(n))
        g.add_edges_from(edges)
        M = np.zeros((n, n))
        for (a, b), x in zip(edges, w):
            M[a, b] = M[b, a] = x
        graphs.append(g); mats.append(M); want_d.append(sp_delay); want_h.append(sp_hop)
    # degenerate members of the batch: one node, two components (+inf), a self loop
    g1 = nx.Graph(); g1.add_node(0)
    g2 = nx.Graph(); g2.add_nodes_from(range(4)); g2.add_edges_from([(0, 1), (2, 3), (1, 1)])
    graphs += [g1, g2]; mats += [np.zeros((1, 1)), np.ones((4, 4))]
    want_d += [AO.apsp_lengths(1, [], None), AO.apsp_lengths(4, [(0, 1), (2, 3)], [1.0, 1.0])]
    want_h += [AO.apsp_lengths(1, [], None), AO.apsp_lengths(4, [(0, 1), (2, 3)], None)]
    plan = ApspPlan(graphs, device="cuda:0")
    got_d = plan.lengths(plan.entry_weights(mats))
    got_h = plan.hops()
    for i in range(len(graphs)):
        assert np.array_equal(got_d[i], want_d[i]), i
        assert np.array_equal(got_h[i], want_h[i]), i
