"""Queue-model head that follows the GNN, and the analytic critic - host/torch side (fp64).

Restates (does not copy) the TensorFlow code of the reference:
  * ``ACOAgent.forward``   src/gnn_offloading_agent.py:229-276  (lambda -> link/node delays -> N x N matrix)
  * critic in ``forward_backward`` :333-374 (routes -> loss, d loss / d routes under a nested tape)
torch.autograd plays the role of tf.GradientTape; everything is float64 like the reference.  This is
the part SURVEY 8(f) lists as "next" after the ChebConv path: tiny tensors (L <= 216 links), kept in
torch ops rather than a hand-written kernel for now.
"""
from __future__ import annotations

import numpy as np
import torch


class _MulNoNan(torch.autograd.Function):
    """tf.math.multiply_no_nan(x, y): x*y but exactly 0 where y == 0, with TensorFlow's gradient
    (grad_x = mul_no_nan(grad, y), grad_y = mul_no_nan(x, grad))."""

    @staticmethod
    def forward(ctx, x, y):
        ctx.save_for_backward(x, y)
        return torch.where(y == 0, torch.zeros((), dtype=x.dtype, device=x.device), x * y)

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        zero = torch.zeros((), dtype=g.dtype, device=g.device)
        gx = torch.where(y == 0, zero, g * y)
        gy = torch.where(g == 0, zero, x * g)
        # un-broadcast
        while gx.dim() > x.dim():
            gx = gx.sum(0)
        for i, (a, b) in enumerate(zip(gx.shape, x.shape)):
            if a != b:
                gx = gx.sum(i, keepdim=True)
        while gy.dim() > y.dim():
            gy = gy.sum(0)
        for i, (a, b) in enumerate(zip(gy.shape, y.shape)):
            if a != b:
                gy = gy.sum(i, keepdim=True)
        return gx, gy


def mul_no_nan(x, y):
    return _MulNoNan.apply(x, y)


def queue_delays(link_lambda, node_lambda, link_rates, cf_degs, node_mu, adj_i, T):
    """:233-254.  link_lambda [L,1], node_lambda [nc,1] -> (link_delay [L,1], node_delay [nc,1]).

    10 iterations of the conflict fixed point mu <- r / (1 + A_i clip(lambda/mu, 0, 1)), M/M/1 delay
    1/(mu - lambda), congested entries replaced by T*lambda/(101 mu) (links) / T*lambda/(100 mu) (nodes)."""
    link_mu = (link_rates / (cf_degs + 1.0)).reshape(-1, 1)
    rates = link_rates.reshape(-1, 1)
    for _ in range(10):
        busy = torch.clamp(link_lambda / link_mu, 0.0, 1.0)
        link_mu = rates * (1.0 / (1.0 + adj_i @ busy))
    link_delay = 1.0 / (link_mu - link_lambda)
    node_delay = 1.0 / (node_mu - node_lambda)
    link_delay = torch.where((link_lambda - link_mu) > 0, float(T) * (link_lambda / (101.0 * link_mu)), link_delay)
    node_delay = torch.where((node_lambda - node_mu) > 0, float(T) * (node_lambda / (100.0 * node_mu)), node_delay)
    return link_delay, node_delay


class HeadInputs:
    """Per-(network, instance) constants of the head, on `device` in float64 (env fields of :233-238)."""

    def __init__(self, obj, env, device):
        f64 = dict(dtype=torch.float64, device=device if device is not None else "cpu")
        self.maps_ol_el_host = np.asarray(obj.maps_ol_el, dtype=np.int64)
        self.maps_on_el_host = np.asarray(obj.maps_on_el, dtype=np.int64)
        proc = np.asarray(env.proc_bws, dtype=np.float64).reshape(-1)
        self.comp_nodes = np.nonzero(proc > 0)[0]
        self.node_mu_host = proc[self.comp_nodes]
        self.link_rates_host = np.asarray(env.link_rates, dtype=np.float64).reshape(-1)
        self.cf_degs_host = np.asarray(env.cf_degs, dtype=np.float64).reshape(-1)
        self.adj_i_host = env.adj_i
        if device is not None:  # torch copies for the autograd implementation (critic on the CPU, tests)
            self.maps_ol_el = torch.as_tensor(self.maps_ol_el_host, device=device)
            self.maps_on_el = torch.as_tensor(self.maps_on_el_host, device=device)
            self.node_mu = torch.as_tensor(self.node_mu_host.reshape(-1, 1), **f64)
            self.link_rates = torch.as_tensor(self.link_rates_host, **f64)
            self.cf_degs = torch.as_tensor(self.cf_degs_host, **f64)
            adj = env.adj_i
            adj = adj.toarray() if hasattr(adj, "toarray") else np.asarray(adj)
            self.adj_i = torch.as_tensor(adj.astype(np.float64), **f64)
        self.T = float(env.T)
        self.num_nodes = int(env.num_nodes)
        self.num_links = int(env.num_links)
        edges = np.asarray(list(env.graph_c.edges), dtype=np.int64).reshape(-1, 2)
        self.e0, self.e1 = edges[:, 0], edges[:, 1]
        self.edge_link = np.asarray(env.link_matrix)[self.e0, self.e1].astype(np.int64)


def delay_matrices(link_delay, node_delay, hi: HeadInputs, bug_compatible=True):
    """:257-274 -> (delay_mtx tensor [N,N] like the TF one: 0 off-graph, relays' diagonal +inf;
    delay_mtx_np like the numpy twin: NaN off-graph, diagonal = np.fill_diagonal of the (nc,1) array,
    which numpy CYCLES when nc < N - the reference's behaviour)."""
    N = hi.num_nodes
    ld = link_delay.detach().cpu().numpy()
    nd = node_delay.detach().cpu().numpy()
    D_np = np.full((N, N), np.nan)
    D_np[hi.e0, hi.e1] = ld[hi.edge_link, 0]
    D_np[hi.e1, hi.e0] = ld[hi.edge_link, 0]
    if bug_compatible:
        np.fill_diagonal(D_np, nd)
    else:
        diag = np.full(N, np.inf)
        diag[hi.comp_nodes] = nd[:, 0]
        np.fill_diagonal(D_np, diag)
    D_ts = torch.zeros((N, N), dtype=link_delay.dtype, device=link_delay.device)
    e0 = torch.as_tensor(hi.e0, device=D_ts.device)
    e1 = torch.as_tensor(hi.e1, device=D_ts.device)
    el = torch.as_tensor(hi.edge_link, device=D_ts.device)
    D_ts[e0, e1] = link_delay[el, 0]
    D_ts[e1, e0] = link_delay[el, 0]
    diag = torch.full((N,), float("inf"), dtype=D_ts.dtype, device=D_ts.device)
    diag[torch.as_tensor(hi.comp_nodes, device=D_ts.device)] = node_delay[:, 0]
    D_ts = D_ts.clone()
    D_ts.fill_diagonal_(0.0)
    D_ts = D_ts + torch.diag(diag)
    return D_ts, D_np


def seed_from_grad_dist(grad_dist, hi: HeadInputs, device):
    """output_gradients=grad_dist_np of :448 pulled back through the scatter of :260-274:
    link l receives gD[e0,e1] + gD[e1,e0]; computing node c receives gD[c,c]."""
    gD = np.asarray(grad_dist, dtype=np.float64)
    g_ld = np.zeros((hi.num_links, 1))
    np.add.at(g_ld[:, 0], hi.edge_link, gD[hi.e0, hi.e1] + gD[hi.e1, hi.e0])
    g_nd = gD[hi.comp_nodes, hi.comp_nodes].reshape(-1, 1)
    return (torch.as_tensor(g_ld, dtype=torch.float64, device=device),
            torch.as_tensor(g_nd, dtype=torch.float64, device=device))


def critic(routes_np, jobs_load, jobs_data, obj, hi_cpu: HeadInputs, num_edges_ext):
    """:333-374 - analytic M/M/1 critic under autograd.  Returns (loss, grad_routes, delay_job_edge,
    unit_delay_edge) as numpy (the reference .numpy()s them right away, :377-382)."""
    routes = torch.tensor(routes_np, dtype=torch.float64, requires_grad=True)
    load = torch.tensor(jobs_load, dtype=torch.float64)
    data = torch.tensor(jobs_data, dtype=torch.float64)
    link_load = routes @ load
    ll = link_load[hi_cpu.maps_ol_el]
    nl = link_load[hi_cpu.maps_on_el]
    ld, nd = queue_delays(ll, nl, hi_cpu.link_rates, hi_cpu.cf_degs, hi_cpu.node_mu, hi_cpu.adj_i, hi_cpu.T)
    unit = torch.zeros((num_edges_ext,), dtype=torch.float64)
    unit = unit.index_put((hi_cpu.maps_ol_el,), ld[:, 0])
    unit = unit.index_put((hi_cpu.maps_on_el,), nd[:, 0])
    unit = unit.reshape(num_edges_ext, 1)
    unit_job_edge = mul_no_nan(unit, routes)
    delay_job_edge = mul_no_nan(data, unit_job_edge)
    # tf.maximum sends the gradient to its first argument on ties (x >= y); torch.maximum would split it
    delay_job_edge = torch.where(delay_job_edge >= routes, delay_job_edge, routes)
    loss = delay_job_edge.sum()
    (grad_routes,) = torch.autograd.grad(loss, routes)
    return float(loss.item()), grad_routes.numpy(), delay_job_edge.detach().numpy(), unit.detach().numpy()


# ---------------------------------------------------------------------------------------------
# fused, batched head on the device (csrc/queue_head.cu through the C-ABI)
# ---------------------------------------------------------------------------------------------
class HeadBatch:
    """Device-resident head constants of B (network, instance) graphs, concatenated (mho_head_t)."""

    def __init__(self, his, ext_sizes, ctx, device):
        import ctypes as C
        from . import _lib
        self.ctx, self.device, self.his = ctx, torch.device(device), list(his)
        B = len(self.his)
        ext_off = np.concatenate([[0], np.cumsum(ext_sizes)]).astype(np.int32)
        link_off = np.concatenate([[0], np.cumsum([h.num_links for h in self.his])]).astype(np.int32)
        comp_off = np.concatenate([[0], np.cumsum([len(h.comp_nodes) for h in self.his])]).astype(np.int32)
        rp, ci, z = [np.zeros(1, dtype=np.int64)], [], 0
        for h in self.his:
            import scipy.sparse as sp
            A = sp.csr_matrix(h.adj_i_host); A.sort_indices()
            rp.append(A.indptr[1:].astype(np.int64) + z); ci.append(A.indices.astype(np.int32)); z += A.nnz
        cat = lambda xs, dt: np.ascontiguousarray(np.concatenate(xs) if xs else np.zeros(0), dtype=dt)
        host = dict(ext_off=ext_off, link_off=link_off, comp_off=comp_off,
                    maps_ol_el=cat([h.maps_ol_el_host for h in self.his], np.int32),
                    maps_on_el=cat([h.maps_on_el_host for h in self.his], np.int32),
                    link_rates=cat([h.link_rates_host for h in self.his], np.float64),
                    cf_degs=cat([h.cf_degs_host for h in self.his], np.float64),
                    node_mu=cat([h.node_mu_host for h in self.his], np.float64),
                    adj_rowptr=np.concatenate(rp).astype(np.int32), adj_colidx=cat(ci, np.int32))
        self.dev = {k: torch.from_numpy(v).to(self.device) for k, v in host.items()}
        self.total_links, self.total_comp, self.total_ext = int(link_off[-1]), int(comp_off[-1]), int(ext_off[-1])
        self.link_off, self.comp_off, self.ext_off = link_off, comp_off, ext_off
        st = _lib.mho_head_t()
        st.n_graphs, st.max_links = B, int(max([h.num_links for h in self.his] + [0]))
        st.total_links, st.total_comp, st.total_adj_nnz = self.total_links, self.total_comp, int(z)
        for k in ("ext_off", "link_off", "comp_off", "maps_ol_el", "maps_on_el", "link_rates", "cf_degs", "node_mu",
                  "adj_rowptr", "adj_colidx"):
            setattr(st, k, self.dev[k].data_ptr())
        st.T = float(self.his[0].T) if self.his else 0.0
        self.struct, self._C, self._lib = st, C, _lib

    def _stream(self):
        return self._C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def forward(self, lam, save=True):
        """lam: float32 device tensor [total_ext] or [total_ext, 1] -> (link_delay [L], node_delay [nc]) fp64."""
        lam = lam.reshape(-1).contiguous()
        assert lam.dtype == torch.float32 and lam.numel() == self.total_ext
        ld = torch.empty(self.total_links, dtype=torch.float64, device=self.device)
        nd = torch.empty(self.total_comp, dtype=torch.float64, device=self.device)
        saved_mu = torch.empty(11 * max(self.total_links, 1), dtype=torch.float64, device=self.device) if save else None
        rc = self.ctx.lib.mho_queue_head_forward(self.ctx.handle, self._C.byref(self.struct), lam.data_ptr(), ld.data_ptr(),
                                                 nd.data_ptr(), saved_mu.data_ptr() if save else None, self._stream())
        self._lib.check(rc, "mho_queue_head_forward")
        # per-call state of the VJP: the constants of this object are shared (cached per network), so a later forward
        # (e.g. a save=False evaluation of another instance) must not clobber what a taped forward left behind
        if save:
            self.last_tape = (lam, saved_mu)
        return ld, nd

    def backward(self, g_link, g_node, tape=None):
        """gradients wrt (link_delay, node_delay) fp64 -> g_lam float32 [total_ext, 1].  tape = (lam, saved_mu) of the taped
        forward (default: the last forward(save=True) of this object)."""
        lam, saved_mu = tape if tape is not None else self.last_tape
        g_link = g_link.reshape(-1).to(torch.float64).contiguous()
        g_node = g_node.reshape(-1).to(torch.float64).contiguous()
        g_lam = torch.empty(self.total_ext, dtype=torch.float32, device=self.device)
        rc = self.ctx.lib.mho_queue_head_backward(self.ctx.handle, self._C.byref(self.struct), lam.data_ptr(),
                                                  saved_mu.data_ptr(), g_link.data_ptr(), g_node.data_ptr(),
                                                  g_lam.data_ptr(), self._stream())
        self._lib.check(rc, "mho_queue_head_backward")
        return g_lam.reshape(-1, 1)
