"""Import shim for the reference's CPU environment (TEST INFRASTRUCTURE ONLY).

The reference's simulator (src/offloading_v3.py, src/util.py) is plain numpy/networkx and
imports here once ``matplotlib`` is stubbed; it supplies *real* GNN inputs
(``graph_expand()``, src/offloading_v3.py:262-339) for the golden vectors and the statistical
pin.  Nothing is copied: the modules are imported from where they lie under the reference
checkout (``MHO_REFERENCE_SRC`` or /root/reference/src).  Not available on the GPU box.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import scipy.io as sio
import scipy.sparse as sp

REF_ROOT = os.environ.get("MHO_REFERENCE_ROOT", "/root/reference")
REF_SRC = os.environ.get("MHO_REFERENCE_SRC", os.path.join(REF_ROOT, "src"))


def available():
    return os.path.isfile(os.path.join(REF_SRC, "offloading_v3.py"))


def import_env():
    """Returns (AdhocCloud, all_pairs_shortest_paths) from the reference checkout."""
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    from offloading_v3 import AdhocCloud  # noqa
    from util import all_pairs_shortest_paths  # noqa
    return AdhocCloud, all_pairs_shortest_paths


def build_env(filepath, T=1000):
    """AdHoc_test.py:84-109 restated: env from a shipped .mat (links_init uses np.random)."""
    AdhocCloud, _ = import_env()
    mat = sio.loadmat(filepath)
    net_cfg = mat["network"][0, 0]
    link_rates = mat["link_rate"].flatten()
    nodes_info = mat["nodes_info"]
    seed = int(net_cfg["seed"].flatten()[0])
    n = int(net_cfg["num_nodes"].flatten()[0])
    env = AdhocCloud(n, T, seed, cf_radius=0.0, gtype=filepath, trace=True)
    # networkx>=3 returns sparse arrays; offloading_v3.py:448,:503 need matrix semantics
    env.adj_c = sp.csr_matrix(env.adj_c)
    env.adj_i = sp.csr_matrix(env.adj_i)
    env.links_init(link_rates)
    for i in range(n):
        if nodes_info[i, 0] == 2:
            env.add_relay(i)
        elif nodes_info[i, 0] == 1:
            env.add_server(i, float(nodes_info[i, 1]))
        elif nodes_info[i, 0] == 0:
            env.proc_bws[i] = nodes_info[i, 1]
    return env, nodes_info


def sample_jobs(env, nodes_info, arrival_scale=0.15):
    """AdHoc_test.py:113-121 restated (uses the global np.random stream)."""
    env.clear_all_jobs()
    mobile, = np.nonzero(nodes_info[:, 0] == 0)
    mobile = np.random.permutation(mobile)
    num_jobs = np.random.randint(int(0.3 * mobile.size), mobile.size)
    rates = np.random.uniform(0.1, 0.5, (num_jobs,))
    for i in range(num_jobs):
        env.add_job(mobile[i], rate=arrival_scale * rates[i])
    return num_jobs


def gnn_inputs(obj):
    """src/gnn_offloading_agent.py:218-224: operator and the 4 node features."""
    import networkx as nx
    adj = sp.csr_matrix(nx.adjacency_matrix(obj.gi_ext)).astype(np.float64)
    X = np.zeros((obj.num_edges_ext, 4))
    X[:, 0] = obj.edge_self_loop
    X[:, 1] = obj.edge_rate_ext
    X[:, 2] = obj.jobs_arrivals
    X[:, 3] = obj.edge_as_server
    return adj, X
