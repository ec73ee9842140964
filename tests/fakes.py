"""Test doubles (tests/ only): an oracle-backed stand-in for ChebNet / KerasAdamReplay so that the HOST
logic of the agent and drivers can be exercised on the CPU box against the reference's real simulator.
The product never imports this; on a GPU the real libmho path is tested in test_agent_gpu.py."""
import numpy as np
import scipy.sparse as sp
import torch

import chebnet_oracle as O


class OracleChebNet:
    def __init__(self, specs, device="cpu", params=None, seed=0):
        self.specs = list(specs)
        self.device = "cpu"
        self.n_params = int(sum(s.n_params for s in self.specs))
        rng = np.random.default_rng(seed)
        ws = O.glorot_weights([self.specs[0].f_in] + [s.f_out for s in self.specs], self.specs[0].K, rng)
        self.params = torch.zeros(self.n_params, dtype=torch.float64)
        self.set_weights(ws)

    def set_flat(self, flat):
        self.params.copy_(torch.as_tensor(np.asarray(flat, dtype=np.float64).ravel()))

    def get_flat(self):
        return self.params.numpy().copy()

    def set_weights(self, ws):
        self.set_flat(O.flatten_params(ws))

    def get_weights(self):
        return O.unflatten_params(self.get_flat(), [(s.K, s.f_in, s.f_out) for s in self.specs])

    def weights_changed(self):
        pass

    def _mats(self, batch):
        out = []
        for a, b in zip(batch.graph_off[:-1], batch.graph_off[1:]):
            z0, z1 = batch.rowptr[a], batch.rowptr[b]
            vals = np.ones(z1 - z0) if batch.vals is None else batch.vals[z0:z1]
            out.append(sp.csr_matrix((vals, batch.colidx[z0:z1] - a, batch.rowptr[a:b + 1] - z0), shape=(b - a, b - a)))
        return out

    def forward(self, batch, X, save=False, out=None, per_graph_tiles=False):
        ws, acts = self.get_weights(), [s.act for s in self.specs]
        Xn = X.numpy().astype(np.float64)
        ys, caches, o = [], [], 0
        for A in self._mats(batch):
            n = A.shape[0]
            y, c = O.cheb_stack_forward(A, Xn[o:o + n], ws, acts, self.specs[0].slope, return_cache=True)
            ys.append(y); caches.append(c); o += n
        Y = torch.as_tensor(np.concatenate(ys, 0))
        return (Y, caches) if save else Y

    def backward(self, batch, X, Y, saved, dY, need_dx=False, need_sum=True):
        ws = self.get_weights()
        g, o = [], 0
        dYn = dY.numpy().astype(np.float64)
        for A, cache in zip(self._mats(batch), saved):
            n = A.shape[0]
            grads, _ = O.cheb_stack_backward(A, ws, cache, dYn[o:o + n], self.specs[0].slope)
            g.append(O.flatten_params(grads)); o += n
        gpg = torch.as_tensor(np.stack(g))
        return gpg, (gpg.sum(0) if need_sum else None), None


class OracleAdam:
    def __init__(self, net, learning_rate=1e-4, clipnorm=1.0, max_norm=1.0, decay_rate=1.0, decay_steps=100, **kw):
        self.net = net
        shapes = []
        for s in net.specs:
            shapes += [(s.K, s.f_in, s.f_out), (s.f_out,)]
        self.shapes = shapes
        self.opt = O.KerasAdam(shapes, lr=learning_rate, clipnorm=clipnorm, max_norm=max_norm, decay_rate=decay_rate,
                               decay_steps=decay_steps)
        self.iterations = 0
        self.master = net.params

    def set_master(self, flat):
        self.net.set_flat(flat)

    def get_master(self):
        return self.net.get_flat()

    def apply(self, grads):
        if grads.dim() == 1:
            grads = grads.unsqueeze(0)
        for g in grads.numpy():
            flat = self.net.get_flat()
            ps, gs, o = [], [], 0
            for sh in self.shapes:
                n = int(np.prod(sh)); ps.append(flat[o:o + n].reshape(sh)); gs.append(g[o:o + n].reshape(sh)); o += n
            self.opt.apply(ps, gs)
            self.net.set_flat(np.concatenate([p.ravel() for p in ps]))
            self.iterations += 1


class FakeGraphBatch:
    """GraphBatch without libmho's planner (CPU tests of host logic only)."""

    def __init__(self, graph_off, rowptr, colidx, vals):
        self.graph_off, self.rowptr, self.colidx, self.vals = graph_off, rowptr, colidx, vals
        self.n_graphs = len(graph_off) - 1
        self.total_nodes = int(graph_off[-1])

    @classmethod
    def from_scipy(cls, mats, device=None, **kw):
        g, rp, ci, va = O.concat_batch(mats)
        return cls(g, rp, ci, va)


def install(monkeypatch):
    """Swap the libmho-backed classes inside the agent module for the oracle-backed doubles."""
    import multihop_offload_b200.gnn_offloading_agent as mod
    monkeypatch.setattr(mod, "ChebNet", OracleChebNet)
    monkeypatch.setattr(mod, "KerasAdamReplay", OracleAdam)
    monkeypatch.setattr(mod, "GraphBatch", FakeGraphBatch)
    return mod
