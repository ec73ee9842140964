"""Batched all-pairs shortest path lengths on the GPU (mho_apsp): the replacement of the reference's
``util.all_pairs_shortest_paths`` (src/util.py:101-110), which runs networkx Dijkstra from every node on the CPU -
the largest cost of a rollout step (call sites gnn_offloading_agent.py:286-287,304-305, AdHoc_test.py:135-136).

``ApspPlan`` turns undirected graphs (networkx, nodes 0..n-1 like ``env.graph_c``) into the concatenated CSR the kernel
reads, once per topology; ``lengths()`` evaluates one set of edge weights (or hop counts).  Results are bit-identical
to Dijkstra's algorithm (see csrc/apsp.cu).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class ApspPlan(object):
    def __init__(self, graphs, device="cuda:0", ctx=None):
        import torch
        if not isinstance(graphs, (list, tuple)):
            graphs = [graphs]
        self.device = torch.device(device)
        self.ctx = ctx if ctx is not None else _lib.Context(self.device.index or 0)
        node_off, rowptr, cols, esrc, edst = [0], [0], [], [], []
        for g in graphs:
            n = g.number_of_nodes()
            assert sorted(g.nodes) == list(range(n)), "ApspPlan: nodes must be 0..n-1 (as in env.graph_c)"
            nbrs = [[] for _ in range(n)]
            for (a, b) in g.edges:          # orientation as the reference reads the weight: graph[a][b] <- M[a, b]
                if a == b:
                    continue                # a self loop never shortens a path
                nbrs[a].append((b, a, b))
                nbrs[b].append((a, a, b))
            base = node_off[-1]
            for v in range(n):
                for (u, a, b) in sorted(nbrs[v]):
                    cols.append(base + u); esrc.append(a); edst.append(b)
                rowptr.append(len(cols))
            node_off.append(base + n)
        self.n_graphs = len(graphs)
        self.sizes = np.diff(np.asarray(node_off, dtype=np.int64))
        self.node_off = np.asarray(node_off, dtype=np.int32)
        self.out_off = np.concatenate([[0], np.cumsum(self.sizes ** 2)]).astype(np.int64)
        self.rowptr = np.asarray(rowptr, dtype=np.int32)
        self.colidx = np.asarray(cols, dtype=np.int32)
        self.e_src = np.asarray(esrc, dtype=np.int64)   # per directed entry: the (src, dst) of its undirected edge,
        self.e_dst = np.asarray(edst, dtype=np.int64)   # LOCAL node ids, in the orientation graph.edges yields
        self.e_graph = np.repeat(np.arange(self.n_graphs), np.diff(self.rowptr[self.node_off]))
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)  # noqa: E731
        self.dev = dict(node_off=up(self.node_off), rowptr=up(self.rowptr), colidx=up(self.colidx), out_off=up(self.out_off[:-1].copy()))
        self._hops = None

    def entry_weights(self, mats):
        """Per directed entry weights from per-graph matrices M (weight of edge (a, b) = M[a, b], gnn_offloading_agent.py:282-283)."""
        if not isinstance(mats, (list, tuple)):
            mats = [mats]
        w = np.empty(self.colidx.size, dtype=np.float64)
        for g, M in enumerate(mats):
            sel = self.e_graph == g
            w[sel] = np.asarray(M, dtype=np.float64)[self.e_src[sel], self.e_dst[sel]]
        return w

    def lengths(self, weights=None, as_numpy=True):
        """weights: None (hop counts) or fp64 per directed entry (entry_weights).  Returns a list of n x n arrays."""
        import torch
        out = torch.empty(int(self.out_off[-1]), dtype=torch.float64, device=self.device)
        wd = None
        if weights is not None:
            wd = torch.as_tensor(np.ascontiguousarray(weights, dtype=np.float64)).to(self.device) if not torch.is_tensor(weights) else weights
            assert wd.dtype == torch.float64 and wd.numel() == self.colidx.size
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        rc = self.ctx.lib.mho_apsp(self.ctx.handle, self.n_graphs, self.dev["node_off"].data_ptr(), self.dev["rowptr"].data_ptr(),
                                   self.dev["colidx"].data_ptr() if self.colidx.size else None,
                                   wd.data_ptr() if wd is not None else None, self.dev["out_off"].data_ptr(), out.data_ptr(), st)
        _lib.check(rc, "mho_apsp")
        if not as_numpy:
            return out
        flat = out.cpu().numpy()
        return [flat[self.out_off[g]:self.out_off[g + 1]].reshape(int(self.sizes[g]), int(self.sizes[g])) for g in range(self.n_graphs)]

    def hops(self):
        if self._hops is None:
            self._hops = self.lengths(None)
        return self._hops
