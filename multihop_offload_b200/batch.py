"""GraphBatch: many independent graph instances concatenated block-diagonally (mho_batch_t).

The reference converts one scipy adjacency per call (spektral.utils.sp_matrix_to_sp_tensor,
call site src/gnn_offloading_agent.py:148) - coordinates sorted row-major, i.e. CSR.  Here the
conversion happens once per batch on the host (numpy), the arrays then live in HBM.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _csr_parts(A):
    import scipy.sparse as sp
    A = sp.csr_matrix(A)
    if not A.has_sorted_indices:
        A = A.sorted_indices()
    return A.indptr, A.indices, A.data, A.shape[0]


def pack_order(sizes, tile_rows=128):
    """Order of independent graphs that makes consecutive tiles (runs of <= tile_rows nodes) nearly full:
    first-fit decreasing bin packing, bins concatenated.  The order of graph instances inside a batch carries
    no meaning, so whoever assembles a batch (a rollout driver, the benchmark) can lay it out this way; the
    forward kernel's cost per tile is almost independent of how full the tile is, so fewer, fuller tiles win."""
    sizes = np.asarray(sizes, dtype=np.int64)
    order = np.argsort(-sizes, kind="stable")
    bins, room = [], []
    for g in order:
        n = int(sizes[g])
        for b in range(len(bins)):
            if room[b] >= n:
                bins[b].append(int(g)); room[b] -= n
                break
        else:
            bins.append([int(g)]); room.append(tile_rows - n)
    return np.asarray([g for b in bins for g in b], dtype=np.int64)


class GraphBatch:
    """Host CSR arrays (+ device copies) of a block-diagonal batch and its tile plan."""

    def __init__(self, graph_off, rowptr, colidx, vals=None, symmetric=True, tile_rows=128, device=None,
                 transpose=None):
        self.graph_off = np.ascontiguousarray(graph_off, dtype=np.int32)
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        self.colidx = np.ascontiguousarray(colidx, dtype=np.int32)
        self.vals = None if vals is None else np.ascontiguousarray(vals, dtype=np.float32)
        self.n_graphs = int(self.graph_off.size - 1)
        self.total_nodes = int(self.graph_off[-1]) if self.graph_off.size else 0
        self.total_nnz = int(self.rowptr[-1]) if self.rowptr.size else 0
        self.symmetric = bool(symmetric)
        self.transpose = transpose  # (rowptr_t, colidx_t, vals_t) host arrays when not symmetric
        assert self.rowptr.size == self.total_nodes + 1 and self.colidx.size == self.total_nnz
        self.tile_rows = int(tile_rows)
        self._plan(self.tile_rows)
        self.dev = {}
        self.device = None
        self._struct_cache = {}
        if device is not None:
            self.to(device)

    # ---- constructors -------------------------------------------------------------------
    @classmethod
    def from_scipy(cls, mats, binary=None, **kw):
        """mats: list of scipy sparse operators, one per graph instance."""
        g_off = np.zeros(len(mats) + 1, dtype=np.int64)
        rps, cis, vas = [np.zeros(1, dtype=np.int64)], [], []
        noff = zoff = 0
        sym = True
        for i, A in enumerate(mats):
            indptr, indices, data, n = _csr_parts(A)
            rps.append(indptr[1:].astype(np.int64) + zoff)
            cis.append(indices.astype(np.int64) + noff)
            vas.append(np.asarray(data, dtype=np.float64))
            noff += n
            zoff += indices.size
            g_off[i + 1] = noff
            if sym:
                import scipy.sparse as sp
                A2 = sp.csr_matrix(A)
                sym = (abs(A2 - A2.T)).nnz == 0
        rowptr = np.concatenate(rps)
        colidx = np.concatenate(cis) if cis else np.zeros(0, dtype=np.int64)
        vals = np.concatenate(vas) if vas else np.zeros(0)
        if binary is None:
            binary = bool(np.all(vals == 1.0))
        transpose = None
        if not sym:
            import scipy.sparse as sp
            Ablk = sp.csr_matrix((vals, colidx, rowptr), shape=(noff, noff)).T.tocsr()
            Ablk.sort_indices()
            transpose = (Ablk.indptr.astype(np.int32), Ablk.indices.astype(np.int32), Ablk.data.astype(np.float32))
        return cls(g_off, rowptr, colidx, None if binary else vals, symmetric=sym, transpose=transpose, **kw)

    # ---- tile plan ----------------------------------------------------------------------
    def _plan(self, tile_rows):
        lib = _lib.load_library()
        tile_off = np.zeros(self.n_graphs + 1, dtype=np.int32)
        nt, mr, mz = C.c_int32(), C.c_int32(), C.c_int32()
        _lib.check(lib.mho_plan_tiles(self.graph_off.ctypes.data, self.rowptr.ctypes.data, self.n_graphs,
                                      int(tile_rows), tile_off.ctypes.data, C.byref(nt), C.byref(mr), C.byref(mz)),
                   "mho_plan_tiles")
        self.n_tiles, self.max_tile_rows, self.max_tile_nnz = nt.value, mr.value, mz.value
        self.tile_off = tile_off[: self.n_tiles + 1].copy()
        self.tile_info = np.zeros((max(self.n_tiles, 1), 4), dtype=np.int32)
        self.graph_info = np.zeros((max(self.n_graphs, 1), 4), dtype=np.int32)
        _lib.check(lib.mho_fill_tile_info(self.graph_off.ctypes.data, self.rowptr.ctypes.data, self.tile_off.ctypes.data,
                                          self.n_tiles, self.tile_info.ctypes.data), "mho_fill_tile_info")
        _lib.check(lib.mho_fill_tile_info(self.graph_off.ctypes.data, self.rowptr.ctypes.data, None,
                                          self.n_graphs, self.graph_info.ctypes.data), "mho_fill_tile_info")
        # binary operator, tiles of <= 128 nodes: bit rows for the tensor-core forward (16 B per node)
        self.adj_bits = None
        if self.vals is None and self.n_tiles and self.max_tile_rows <= 128 and self.total_nodes:
            bits = np.zeros((self.total_nodes, 4), dtype=np.uint32)
            _lib.check(lib.mho_fill_adj_bits(self.graph_off.ctypes.data, self.rowptr.ctypes.data,
                                             self.colidx.ctypes.data if self.total_nnz else None, self.tile_off.ctypes.data,
                                             self.n_tiles, bits.ctypes.data), "mho_fill_adj_bits")
            self.adj_bits = bits
        # row tiles: plain runs of 128 nodes that ignore graph boundaries - valid (and perfectly full) whenever no layer
        # touches the operator (every K = 1, the reference's shipped model), whatever the graph sizes
        nrt = (self.total_nodes + 127) // 128
        n0 = 128 * np.arange(max(nrt, 1), dtype=np.int64)
        n1 = np.minimum(n0 + 128, self.total_nodes)
        self.row_tile_info = np.zeros((max(nrt, 1), 4), dtype=np.int32)
        if nrt:
            self.row_tile_info[:, 0] = n0
            self.row_tile_info[:, 1] = n1 - n0
            self.row_tile_info[:, 2] = self.rowptr[n0]
            self.row_tile_info[:, 3] = self.rowptr[n1] - self.rowptr[n0]
        self.n_row_tiles = int(nrt)
        self.max_row_tile_nnz = int(self.row_tile_info[:, 3].max()) if nrt else 0
        # largest tile first: the kernel's CTAs pull tiles in this order from a global counter
        self.tile_graph0 = self.tile_off[: max(self.n_tiles, 1)].astype(np.int32).copy()   # first graph of each listed tile
        if self.n_tiles > 1:
            cost = 3 * self.tile_info[:, 1].astype(np.int64) + self.tile_info[:, 3]
            perm = np.argsort(-cost, kind="stable")
            self.tile_info = np.ascontiguousarray(self.tile_info[perm])
            self.tile_graph0 = np.ascontiguousarray(self.tile_graph0[perm])
        # one-graph-per-tile statistics (the backward runs one graph per CTA)
        if self.n_graphs:
            sizes = np.diff(self.graph_off)
            nnzs = self.rowptr[self.graph_off[1:]] - self.rowptr[self.graph_off[:-1]]
            self.max_graph_rows, self.max_graph_nnz = int(sizes.max()), int(nnzs.max())
        else:
            self.max_graph_rows = self.max_graph_nnz = 0

    # ---- device residency ---------------------------------------------------------------
    def to(self, device):
        import torch
        device = torch.device(device)
        self.device = device
        self._struct_cache = {}

        def up(a):
            return torch.from_numpy(np.ascontiguousarray(a)).to(device, non_blocking=False)

        self.dev = dict(graph_off=up(self.graph_off), rowptr=up(self.rowptr), colidx=up(self.colidx),
                        tile_off=up(self.tile_off), tile_info=up(self.tile_info), graph_info=up(self.graph_info),
                        tile_graph0=up(self.tile_graph0),
                        row_tile_info=up(self.row_tile_info))
        if self.vals is not None:
            self.dev["vals"] = up(self.vals)
        if getattr(self, "adj_bits", None) is not None:
            self.dev["adj_bits"] = torch.from_numpy(self.adj_bits.view(np.int32)).to(device)
        if self.transpose is not None:
            self.dev["rowptr_t"], self.dev["colidx_t"], self.dev["vals_t"] = (up(a) for a in self.transpose)
        return self

    def _graph_bits(self):
        """Bit rows relative to each GRAPH's first node (the tensor-core VJP runs one graph per CTA pass); built and
        uploaded on first use.  None for weighted operators or graphs of more than 128 nodes."""
        if "adj_bits_graph" in self.dev:
            return self.dev["adj_bits_graph"]
        import torch
        t = None
        if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
            raise _lib.MhoError("GraphBatch: the per-graph bit rows of the tensor-core VJP are built (host) and uploaded on first use - "
                                "run one backward (or batch.struct(per_graph_tiles=True)) on this batch before capturing a CUDA graph")
        if self.vals is None and self.n_graphs and self.total_nodes and self.max_graph_rows <= 128:
            if self.adj_bits is not None and self.n_tiles == self.n_graphs:
                t = self.dev["adj_bits"]   # every tile is one graph
            else:
                bits = np.zeros((self.total_nodes, 4), dtype=np.uint32)
                one_per_tile = np.arange(self.n_graphs + 1, dtype=np.int32)
                _lib.check(_lib.load_library().mho_fill_adj_bits(
                    self.graph_off.ctypes.data, self.rowptr.ctypes.data, self.colidx.ctypes.data if self.total_nnz else None,
                    one_per_tile.ctypes.data, self.n_graphs, bits.ctypes.data), "mho_fill_adj_bits")
                t = torch.from_numpy(bits.view(np.int32)).to(self.device)
        self.dev["adj_bits_graph"] = t
        return t

    def struct_ref(self, per_graph_tiles=False, row_tiles=False):
        """Cached ctypes byref of the mho_batch_t (device arrays never move after .to())."""
        key = (bool(per_graph_tiles), bool(row_tiles))
        c = self._struct_cache.get(key)
        if c is None:
            st = self.struct(per_graph_tiles=key[0], row_tiles=key[1])
            c = (st, C.byref(st))
            self._struct_cache[key] = c
        return c[1]

    def struct(self, per_graph_tiles=False, row_tiles=False):
        """mho_batch_t over the device arrays.  per_graph_tiles=True => tile_off NULL (backward); row_tiles=True =>
        128-node row tiles that ignore graph boundaries (only for stacks whose layers all have K = 1)."""
        assert self.dev, "GraphBatch.to(device) first"
        b = _lib.mho_batch_t()
        b.n_graphs, b.total_nodes, b.total_nnz = self.n_graphs, self.total_nodes, self.total_nnz
        b.graph_off = self.dev["graph_off"].data_ptr()
        b.rowptr = self.dev["rowptr"].data_ptr()
        b.colidx = self.dev["colidx"].data_ptr() if self.total_nnz else None
        b.vals = self.dev["vals"].data_ptr() if "vals" in self.dev else None
        if self.transpose is not None:
            b.rowptr_t = self.dev["rowptr_t"].data_ptr()
            b.colidx_t = self.dev["colidx_t"].data_ptr()
            b.vals_t = self.dev["vals_t"].data_ptr() if "vals" in self.dev else None
        if row_tiles:
            b.tile_off, b.n_tiles = self.dev["tile_off"].data_ptr(), self.n_row_tiles   # non-NULL marks "tiled"
            b.tile_info = self.dev["row_tile_info"].data_ptr()
            b.max_tile_rows, b.max_tile_nnz = min(128, self.total_nodes), self.max_row_tile_nnz
        elif per_graph_tiles:
            b.tile_off, b.n_tiles = None, self.n_graphs
            b.tile_info = self.dev["graph_info"].data_ptr()
            b.max_tile_rows, b.max_tile_nnz = self.max_graph_rows, self.max_graph_nnz
            gb = self._graph_bits()
            b.adj_bits = gb.data_ptr() if gb is not None else None
        else:
            b.tile_off, b.n_tiles = self.dev["tile_off"].data_ptr(), self.n_tiles
            b.tile_info = self.dev["tile_info"].data_ptr()
            b.max_tile_rows, b.max_tile_nnz = self.max_tile_rows, self.max_tile_nnz
            b.adj_bits = self.dev["adj_bits"].data_ptr() if "adj_bits" in self.dev else None
            b.tile_graph0 = self.dev["tile_graph0"].data_ptr() if "tile_graph0" in self.dev else None
        return b

    # ---- sharding across ranks (partition by graph, balanced by rows+nnz) ------
    def shard(self, rank, world):
        """Contiguous shard of graphs for `rank`, balanced on (nodes + nnz)."""
        if world == 1:
            return self
        sizes = np.diff(self.graph_off).astype(np.int64)
        nnzs = (self.rowptr[self.graph_off[1:]] - self.rowptr[self.graph_off[:-1]]).astype(np.int64)
        cost = np.concatenate([[0], np.cumsum(32 * sizes + nnzs)])
        bounds = [int(np.searchsorted(cost, cost[-1] * r / world, side="left")) for r in range(world + 1)]
        bounds[0], bounds[-1] = 0, self.n_graphs
        g0, g1 = bounds[rank], max(bounds[rank], bounds[rank + 1])
        n0, n1 = int(self.graph_off[g0]), int(self.graph_off[g1])
        z0, z1 = int(self.rowptr[n0]), int(self.rowptr[n1])
        transpose = None
        if not self.symmetric:
            # the VJP walks the transposed operator: rebuild it for the shard's block (graphs are block-diagonal, so the
            # transpose of the slice is the slice of the transpose)
            import scipy.sparse as sp
            vals = np.ones(z1 - z0, dtype=np.float32) if self.vals is None else self.vals[z0:z1]
            At = sp.csr_matrix((vals, self.colidx[z0:z1] - n0, self.rowptr[n0:n1 + 1] - z0), shape=(n1 - n0, n1 - n0)).T.tocsr()
            At.sort_indices()
            transpose = (At.indptr.astype(np.int32), At.indices.astype(np.int32), At.data.astype(np.float32))
        sub = GraphBatch(self.graph_off[g0:g1 + 1] - n0, self.rowptr[n0:n1 + 1] - z0, self.colidx[z0:z1] - n0,
                         None if self.vals is None else self.vals[z0:z1], symmetric=self.symmetric,
                         tile_rows=self.tile_rows, transpose=transpose)
        sub.node_range = (n0, n1)
        sub.graph_range = (g0, g1)
        return sub
