"""ChebNet: the stack of ChebConv layers of the reference's actor, executed by libmho.

Mirrors ``ACOAgent._build_model`` (src/gnn_offloading_agent.py:81-123): ``num_layer`` ChebConv
layers, widths n_features -> 32 -> ... -> 32 -> 1, leaky_relu on all but the last (relu), bias,
Chebyshev order K (Spektral default 1 in the shipped checkpoints; any K<=16 here).

All parameters live in ONE flat fp32 device buffer in variable-creation order
(kernel_0, bias_0, kernel_1, ...): the layout of the checkpoint's data file, of the gradient
list the reference memorises (:142,:450) and of the buffer a data-parallel all-reduce ships.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import ACT_LEAKY, ACT_RELU


@dataclass(frozen=True)
class LayerSpec:
    K: int
    f_in: int
    f_out: int
    act: int = ACT_LEAKY
    slope: float = 0.2  # tf.nn.leaky_relu default; unpinned in the reference

    @property
    def n_params(self):
        return self.K * self.f_in * self.f_out + self.f_out


def reference_stack(K=1, num_layer=5, n_features=4, hidden=32, out=1, slope=0.2):
    """Layer specs of ACOAgent._build_model (gnn_offloading_agent.py:87-110)."""
    specs, fi = [], n_features
    for l in range(num_layer):
        last = l == num_layer - 1
        fo = out if last else hidden
        specs.append(LayerSpec(K, fi, fo, ACT_RELU if last else ACT_LEAKY, slope))
        fi = fo
    return specs


def glorot_uniform_(specs, rng):
    """Keras glorot_uniform for a (K, f_in, f_out) kernel (fan = K*f_in, K*f_out), zeros bias."""
    parts = []
    for s in specs:
        lim = np.sqrt(6.0 / (s.K * s.f_in + s.K * s.f_out))
        parts.append(rng.uniform(-lim, lim, size=s.K * s.f_in * s.f_out))
        parts.append(np.zeros(s.f_out))
    return np.concatenate(parts)


class ChebNet:
    def __init__(self, specs, device="cuda:0", params=None, seed=0, private_context=False):
        """private_context=True: this net gets its own mho_ctx_t (tile-scheduler counters, weight-image cache, staging
        slots) instead of the per-device shared one - required when several nets / streams launch concurrently: a
        context serialises nothing itself and must only be used from one stream at a time."""
        import torch
        self.specs = list(specs)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.MhoError("ChebNet needs a CUDA device (no CPU fallback by design)")
        self.ctx = _lib.Context(self.device.index or 0) if private_context else _lib.Context.get(self.device.index or 0)
        self.n_params = int(sum(s.n_params for s in self.specs))
        if params is None:
            params = glorot_uniform_(self.specs, np.random.default_rng(seed))
        self.params = torch.empty(self.n_params, dtype=torch.float32, device=self.device)
        self.set_flat(params)
        self._offsets = []
        o = 0
        for s in self.specs:
            self._offsets.append((o, o + s.K * s.f_in * s.f_out))
            o += s.n_params

    # ---- parameters ---------------------------------------------------------------------
    def set_flat(self, flat):
        import torch
        flat = np.ascontiguousarray(np.asarray(flat, dtype=np.float32).ravel())
        assert flat.size == self.n_params, (flat.size, self.n_params)
        self.params.copy_(torch.from_numpy(flat))
        self.weights_changed()

    def weights_changed(self):
        """Tell libmho the parameter buffer was modified in place (packed weight images are stale)."""
        self.ctx.lib.mho_invalidate_weights(self.ctx.handle)

    def get_flat(self):
        return self.params.detach().cpu().numpy().astype(np.float64)

    def set_weights(self, weights):
        """weights: list of (W[K,f_in,f_out], b[f_out]) - the Keras get_weights() pairing."""
        self.set_flat(np.concatenate([np.concatenate([np.asarray(W).ravel(), np.asarray(b).ravel()])
                                      for W, b in weights]))

    def get_weights(self):
        flat, out, o = self.get_flat(), [], 0
        for s in self.specs:
            nW = s.K * s.f_in * s.f_out
            out.append((flat[o:o + nW].reshape(s.K, s.f_in, s.f_out).copy(), flat[o + nW:o + nW + s.f_out].copy()))
            o += s.n_params
        return out

    def layer_structs(self, params=None):
        if params is None:
            if getattr(self, "_layers_cache", None) is None:
                self._layers_cache = self.layer_structs(self.params)  # params buffer is never reallocated
            return self._layers_cache
        p = params
        base = p.data_ptr()
        arr = (_lib.mho_layer_t * len(self.specs))()
        for i, (s, (ow, ob)) in enumerate(zip(self.specs, self._offsets)):
            arr[i].K, arr[i].f_in, arr[i].f_out, arr[i].act, arr[i].slope = s.K, s.f_in, s.f_out, s.act, s.slope
            arr[i].W = base + 4 * ow
            arr[i].b = base + 4 * ob
        return arr

    # ---- forward / backward -------------------------------------------------------------
    def _stream(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def saved_floats(self, batch):
        return int(sum(batch.total_nodes * s.f_in for s in self.specs[1:]))

    def forward(self, batch, X, save=False, out=None, per_graph_tiles=False):
        """X: float32 device tensor [total_nodes, f_in].  Returns Y (and the saved activations)."""
        import torch
        assert X.is_cuda and X.dtype == torch.float32 and X.is_contiguous()
        assert X.shape[0] == batch.total_nodes and X.shape[1] == self.specs[0].f_in, X.shape
        Y = out if out is not None else torch.empty((batch.total_nodes, self.specs[-1].f_out), dtype=torch.float32,
                                                    device=self.device)
        saved = torch.empty(max(self.saved_floats(batch), 1), dtype=torch.float32, device=self.device) if save else None
        # no layer touches the operator (all K = 1, the reference's shipped model): full 128-node row tiles, any graph size
        row_tiles = (not per_graph_tiles) and all(s.K == 1 for s in self.specs) \
            and all(s.f_in <= 32 and s.f_out <= 32 for s in self.specs)
        rc = self.ctx.lib.mho_cheb_forward(self.ctx.handle, batch.struct_ref(per_graph_tiles, row_tiles), self.layer_structs(),
                                           len(self.specs), X.data_ptr(), Y.data_ptr(),
                                           saved.data_ptr() if save else None, self._stream())
        if rc:
            _lib.check(rc, "mho_cheb_forward")
        return (Y, saved) if save else Y

    def backward(self, batch, X, Y, saved, dY, need_dx=False, need_sum=True):
        """VJP (gnn_offloading_agent.py:448).  Returns (grads_per_graph [B,P], grads_sum [P] | None, dX | None)."""
        import torch
        P = self.n_params
        gpg = torch.empty((batch.n_graphs, P), dtype=torch.float32, device=self.device)
        gsum = torch.empty(P, dtype=torch.float32, device=self.device) if need_sum else None
        dX = torch.empty_like(X) if need_dx else None
        b = batch.struct(per_graph_tiles=True)
        layers = self.layer_structs()
        assert dY.is_contiguous() and dY.dtype == torch.float32 and dY.shape == Y.shape
        rc = self.ctx.lib.mho_cheb_backward(self.ctx.handle, C.byref(b), layers, len(self.specs), X.data_ptr(),
                                            Y.data_ptr(), saved.data_ptr() if saved is not None else None,
                                            dY.data_ptr(), gpg.data_ptr(), gsum.data_ptr() if need_sum else None,
                                            dX.data_ptr() if need_dx else None, self._stream())
        _lib.check(rc, "mho_cheb_backward")
        return gpg, gsum, dX

    def forward_host(self, graph_off, rowptr, colidx, vals, X_host, Y_host=None):
        """Host-buffer call (numpy in, numpy out) through mho_cheb_forward_host; synchronises."""
        graph_off = np.ascontiguousarray(graph_off, dtype=np.int32)
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        colidx = np.ascontiguousarray(colidx, dtype=np.int32)
        X_host = np.ascontiguousarray(X_host, dtype=np.float32)
        if vals is not None:
            vals = np.ascontiguousarray(vals, dtype=np.float32)
        n = int(graph_off[-1])
        if Y_host is None:
            Y_host = np.empty((n, self.specs[-1].f_out), dtype=np.float32)
        rc = self.ctx.lib.mho_cheb_forward_host(self.ctx.handle, graph_off.size - 1, graph_off.ctypes.data,
                                                rowptr.ctypes.data, colidx.ctypes.data,
                                                vals.ctypes.data if vals is not None else None,
                                                self.layer_structs(), len(self.specs), X_host.ctypes.data,
                                                Y_host.ctypes.data, self._stream())
        _lib.check(rc, "mho_cheb_forward_host")
        return Y_host

    def forward_host_async(self, graph_off, rowptr, colidx, vals, X_host, Y_host):
        """Pipelined host-buffer call (mho_cheb_forward_host_async): returns a ticket right after enqueuing; Y_host is
        complete after host_wait(ticket).  At most two calls in flight, each with its own (page-locked) host buffers;
        the arrays must already be contiguous int32 / float32 (no hidden copies here: they would be freed too early)."""
        for a, dt in ((graph_off, np.int32), (rowptr, np.int32), (colidx, np.int32), (X_host, np.float32), (Y_host, np.float32)):
            assert a.dtype == dt and a.flags["C_CONTIGUOUS"], "forward_host_async needs contiguous int32/float32 arrays"
        assert vals is None or (vals.dtype == np.float32 and vals.flags["C_CONTIGUOUS"])
        ticket = C.c_int32(-1)
        rc = self.ctx.lib.mho_cheb_forward_host_async(self.ctx.handle, graph_off.size - 1, graph_off.ctypes.data,
                                                      rowptr.ctypes.data, colidx.ctypes.data,
                                                      vals.ctypes.data if vals is not None else None,
                                                      self.layer_structs(), len(self.specs), X_host.ctypes.data,
                                                      Y_host.ctypes.data, self._stream(), C.byref(ticket))
        _lib.check(rc, "mho_cheb_forward_host_async")
        return ticket.value

    def host_wait(self, ticket):
        _lib.check(self.ctx.lib.mho_host_wait(self.ctx.handle, int(ticket)), "mho_host_wait")
