"""KerasAdamReplay: the reference's optimizer semantics executed by libmho's replay kernel.

Mirrors ``Adam(learning_rate, clipnorm=1.0)`` + ``max_norm(1.0)`` constraints and the sequential
``apply_gradients`` loop of ``ACOAgent.replay`` (src/gnn_offloading_agent.py:104-121,156-169).
Master weights / moments are fp64 on the device (the reference trains in fp64); the ChebNet's fp32
parameter buffer is refreshed by the same launch.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class KerasAdamReplay:
    def __init__(self, net, learning_rate=1e-4, beta_1=0.9, beta_2=0.999, epsilon=1e-7, clipnorm=1.0, max_norm=1.0,
                 decay_rate=1.0, decay_steps=100):
        import torch
        self.net = net
        self.cfg = _lib.mho_adam_t(float(learning_rate), float(beta_1), float(beta_2), float(epsilon),
                                   float(clipnorm or 0.0), float(max_norm or 0.0), float(decay_rate), int(decay_steps))
        self.master = net.params.detach().to(torch.float64).clone()
        self.m = torch.zeros_like(self.master)
        self.v = torch.zeros_like(self.master)
        self.iterations = 0

    def set_master(self, flat64):
        """Load exact fp64 weights (e.g. from a checkpoint) into the master copy and the fp32 mirror."""
        import torch
        flat64 = np.ascontiguousarray(np.asarray(flat64, dtype=np.float64).ravel())
        assert flat64.size == self.net.n_params
        self.master.copy_(torch.from_numpy(flat64))
        self.net.params.copy_(self.master.to(torch.float32))
        self.net.weights_changed()

    def get_master(self):
        return self.master.detach().cpu().numpy().copy()

    def apply(self, grads):
        """grads: float32 device tensor [n_steps, n_params] applied one after the other (replay order)."""
        import torch
        if grads.dim() == 1:
            grads = grads.unsqueeze(0)
        assert grads.is_cuda and grads.dtype == torch.float32 and grads.is_contiguous()
        assert grads.shape[1] == self.net.n_params
        n = int(grads.shape[0])
        rc = self.net.ctx.lib.mho_adam_replay(self.net.ctx.handle, self.net.layer_structs(), len(self.net.specs),
                                              C.byref(self.cfg), self.master.data_ptr(), self.m.data_ptr(),
                                              self.v.data_ptr(), self.net.params.data_ptr(), grads.data_ptr(), n,
                                              self.iterations, self.net._stream())
        _lib.check(rc, "mho_adam_replay")
        self.iterations += n
