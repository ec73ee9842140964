"""Writes tests/golden/apsp_cases.npz from the REFERENCE's own util.all_pairs_shortest_paths (src/util.py:101-110),
imported from the read-only checkout: BA graphs like env.graph_c (offloading_v3.py:40) with random positive delays.
Run in the build container (needs /root/reference); the fixture travels, the reference does not."""
import os
import sys

import networkx as nx
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_env  # noqa: E402

_, all_pairs_shortest_paths = ref_env.import_env()
rng = np.random.default_rng(2024)
out = {}
for i, n in enumerate([5, 20, 37, 64, 110, 170]):
    g = nx.barabasi_albert_graph(n, 2, seed=100 + i)
    edges = np.asarray(list(g.edges), dtype=np.int32)
    w = rng.uniform(0.05, 30.0, size=len(edges))
    if i == 2:
        w[:] = np.round(w)          # many ties
        w[w == 0] = 1.0
    for (a, b), x in zip(edges, w):
        g[a][b]["delay"] = float(x)
    out["n%d" % i] = np.int32(n)
    out["edges%d" % i] = edges
    out["w%d" % i] = w
    out["sp_delay%d" % i] = all_pairs_shortest_paths(g, weight="delay")
    out["sp_hop%d" % i] = all_pairs_shortest_paths(g, weight=None)
out["n_cases"] = np.int32(6)
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "apsp_cases.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, os.path.getsize(dst), "bytes")
