"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's all-pairs shortest path lengths
(src/util.py:101-110: networkx ``all_pairs_dijkstra_path_length`` written into a dense matrix).

Dijkstra with a binary heap from every source, distances accumulated as ``dist[u] + w(u, v)`` exactly like
networkx's ``_dijkstra_multisource`` does, so the numbers are bit-identical to the reference's; pinned against
outputs of the reference function itself (tests/golden/apsp_cases.npz, written by oracle/make_golden_apsp.py).
"""
from __future__ import annotations

import heapq

import numpy as np


def apsp_lengths(n, edges, weights=None):
    """n nodes 0..n-1, undirected `edges` [(a, b)], `weights` per edge or None (hop counts).  Returns [n, n] fp64
    (+inf where the reference would raise KeyError)."""
    adj = [[] for _ in range(n)]
    for i, (a, b) in enumerate(edges):
        w = 1.0 if weights is None else float(weights[i])
        adj[a].append((b, w))
        adj[b].append((a, w))
    out = np.full((n, n), np.inf)
    for s in range(n):
        dist = {}
        seen = {s: 0.0}
        heap = [(0.0, s)]
        while heap:
            d, v = heapq.heappop(heap)
            if v in dist:
                continue
            dist[v] = d
            for (u, w) in adj[v]:
                vu = dist[v] + w
                if u not in dist and (u not in seen or vu < seen[u]):
                    seen[u] = vu
                    heapq.heappush(heap, (vu, u))
        for v, d in dist.items():
            out[s, v] = d
    return out
