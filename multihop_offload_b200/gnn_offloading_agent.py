"""Drop-in replacement of the reference's ``gnn_offloading_agent`` module on the B200-native path.

Same surface as /root/reference/src/gnn_offloading_agent.py: module-level ``FLAGS`` with the flag
names of :44-60 and ``ACOAgent(input_flags, memory_size)`` with ``load, save, makestate, memorize,
predict, act, replay, forward, forward_env, forward_backward`` (:64-453), the attributes ``model,
optimizer, memory, epsilon, flags`` and the 7-tuple returned by ``forward_backward`` (:453).
The ChebConv stack (``self.model([x_in, a_in])``, :149), its VJP (:448) and the optimizer replay
(:156-169) run in libmho's sm_100a kernels; TensorFlow and Spektral are not needed.

New, optional flags (defaults reproduce the reference): ``--K`` Chebyshev order (Spektral default 1),
``--leaky_slope`` (0.2), ``--fix_diag`` (False = keep the wrapped np.fill_diagonal of :269),
``--device``, ``--dp_mode`` (``replay`` = every rank replays the same all-gathered gradients, the
reference's algorithm; ``allreduce`` = one averaged step per replay).
"""
from __future__ import division, print_function

import os
import random
import sys
from collections import deque
from copy import deepcopy

import numpy as np
import scipy.sparse as sp

# ---------------------------------------------------------------------------------------------
# flags: same names/defaults as the reference's tf.compat.v1.flags (= absl), parsed lazily with
# unknown flags tolerated, so drivers can read FLAGS.x at module level exactly as before.
# ---------------------------------------------------------------------------------------------
from absl import flags as _absl_flags


class _LazyFlags(object):
    def __init__(self, fv):
        object.__setattr__(self, "_fv", fv)

    def _parse(self):
        fv = object.__getattribute__(self, "_fv")
        if not fv.is_parsed():
            fv(sys.argv, known_only=True)
        return fv

    def __getattr__(self, name):
        return getattr(self._parse(), name)

    def __setattr__(self, name, value):
        setattr(self._parse(), name, value)


flags = _absl_flags
_FV = _absl_flags.FlagValues()


def _define():
    d = dict(flag_values=_FV)
    flags.DEFINE_string('datapath', '../data_100', 'input data path.', **d)
    flags.DEFINE_string('out', '../out', 'output data path.', **d)
    flags.DEFINE_integer('T', 1000, 'Number of time slots with traffic inputs.', **d)
    flags.DEFINE_boolean('prob', False, 'If probabilistic decision.', **d)
    flags.DEFINE_string('training_set', 'BAm2', 'Name of training dataset', **d)
    flags.DEFINE_float('learning_rate', 0.0001, 'Initial learning rate.', **d)
    flags.DEFINE_float('learning_decay', 1.0, 'Initial learning rate.', **d)
    flags.DEFINE_float('arrival_scale', 0.1, 'Scale of arrival rate.', **d)
    flags.DEFINE_integer('epochs', 201, 'Number of epochs to train.', **d)
    flags.DEFINE_integer('num_layer', 5, 'number of layers.', **d)
    flags.DEFINE_float('dropout', 0, 'Dropout rate (1 - keep probability).', **d)
    flags.DEFINE_float('weight_decay', 5e-4, 'Weight for L2 loss on embedding matrix.', **d)
    flags.DEFINE_float('epsilon', 1.0, 'initial exploration rate', **d)
    flags.DEFINE_float('epsilon_min', 0.001, 'minimal exploration rate', **d)
    flags.DEFINE_float('epsilon_decay', 0.985, 'exploration rate decay per replay', **d)
    flags.DEFINE_float('gamma', 1.0, 'gamma', **d)
    flags.DEFINE_integer('batch', 100, 'batch size.', **d)
    # --- additions of this build (defaults = reference behaviour)
    flags.DEFINE_integer('K', 1, 'Chebyshev order of every ChebConv layer (Spektral default 1).', **d)
    flags.DEFINE_float('leaky_slope', 0.2, 'negative slope of leaky_relu.', **d)
    flags.DEFINE_boolean('fix_diag', False, 'align node delays on the diagonal instead of np.fill_diagonal wrap.', **d)
    flags.DEFINE_string('device', 'cuda:0', 'CUDA device of the GNN.', **d)
    flags.DEFINE_string('dp_mode', 'replay', 'multi-GPU training: replay (all-gather, faithful) | allreduce.', **d)
    flags.DEFINE_string('ref_src', os.environ.get('MHO_REFERENCE_SRC', ''), 'path of the reference src/ (environment simulator).', **d)
    flags.DEFINE_string('modeldir', os.path.join('..', 'model'), 'directory holding model_ChebConv_* checkpoints.', **d)
    flags.DEFINE_integer('max_files', 0, 'process at most this many network files (0 = all).', **d)
    flags.DEFINE_integer('seed', -1, 'numpy/random seed of the drivers (-1 = unseeded like the reference).', **d)
    flags.DEFINE_boolean('batch_instances', False, 'AdHoc_test: evaluate the 10 job instances of a network file in ONE GNN / queue-head / '
                         'shortest-path launch each (same CSV rows as the per-instance loop under a fixed seed).', **d)


_define()
FLAGS = _LazyFlags(_FV)

from . import tf_bundle  # noqa: E402
from .batch import GraphBatch  # noqa: E402
from .chebnet import ChebNet, reference_stack  # noqa: E402
from .optim import KerasAdamReplay  # noqa: E402
from . import queue_head as qh  # noqa: E402


def _apsp():
    """util.all_pairs_shortest_paths of the reference checkout (src/util.py:101-110) - environment side,
    out of scope of this build, imported from where the user keeps the reference."""
    try:
        from util import all_pairs_shortest_paths
        return all_pairs_shortest_paths
    except ImportError:
        src = FLAGS.ref_src
        if src and src not in sys.path:
            sys.path.insert(0, src)
        from util import all_pairs_shortest_paths
        return all_pairs_shortest_paths


def _apsp_or_none():
    """On a CUDA device mho_apsp replaces util.all_pairs_shortest_paths; the reference checkout is only needed for
    the environment (offloading_v3), so a missing util module is not an error there."""
    try:
        return _apsp()
    except Exception:
        return False


class _Model(object):
    """Stand-in for the Keras model attribute: ``agent.model.trainable_weights`` / ``get_weights``."""

    def __init__(self, net):
        self.net = net

    @property
    def trainable_weights(self):
        out = []
        for W, b in self.net.get_weights():
            out += [W, b]
        return out

    def get_weights(self):
        return self.trainable_weights

    def set_weights(self, ws):
        self.net.set_weights([(ws[2 * i], ws[2 * i + 1]) for i in range(len(ws) // 2)])

    def summary(self):
        print("ChebNet on libmho: " + " -> ".join("%d" % s.f_in for s in self.net.specs) + " -> %d, K=%d, %d params"
              % (self.net.specs[-1].f_out, self.net.specs[0].K, self.net.n_params))

    def __call__(self, inputs):
        raise TypeError("call ACOAgent.predict(state); the model runs in libmho, not as a Keras callable")


# Agent
class ACOAgent:
    def __init__(self, input_flags, memory_size=5000):
        self.flags = input_flags
        self.learning_rate = self.flags.learning_rate
        self.n_node_features = 4
        self.output_size = 1
        self.max_degree = 1
        self.num_supports = 1 + self.max_degree
        self.l2_reg = self.flags.weight_decay  # inert in the reference too (model.losses never used)
        self.epsilon = self.flags.epsilon
        self.device = getattr(self.flags, "device", "cuda:0")
        self.K = int(getattr(self.flags, "K", 1))
        self.slope = float(getattr(self.flags, "leaky_slope", 0.2))
        self.bug_compatible = not bool(getattr(self.flags, "fix_diag", False))
        self.model = self._build_model()
        self.memory = deque(maxlen=memory_size)
        self.reward_mem = deque(maxlen=memory_size)
        self._tape = None
        self._adj_cache = {}
        self._head_cache = {}

    def _build_model(self):
        """:81-123 - num_layer ChebConv layers, 4 -> 32 -> ... -> 1, leaky_relu x (L-1) + relu, Adam(clipnorm=1)."""
        specs = reference_stack(K=self.K, num_layer=self.flags.num_layer, n_features=self.n_node_features,
                                hidden=32, out=self.output_size, slope=self.slope)
        self.net = ChebNet(specs, device=self.device, seed=random.randrange(1 << 30))
        decay = float(self.flags.learning_decay)
        self.optimizer = KerasAdamReplay(self.net, learning_rate=self.learning_rate, clipnorm=1.0, max_norm=1.0,
                                         decay_rate=decay, decay_steps=100)
        model = _Model(self.net)
        model.summary()
        return model

    # ---- checkpoints (TF tensor-bundle layout, src/gnn_offloading_agent.py:125-132) -------------
    def load(self, name):
        ckpt = tf_bundle.latest_checkpoint(name)
        if ckpt:
            ws = tf_bundle.load_weights(ckpt)
            shapes = [(s.K, s.f_in, s.f_out) for s in self.net.specs]
            got = [tuple(W.shape) for W, _ in ws]
            if got != shapes:
                raise ValueError("checkpoint %s holds kernels %s but the model is %s (use --K)" % (ckpt, got, shapes))
            flat = np.concatenate([np.concatenate([W.ravel(), b.ravel()]) for W, b in ws])
            self.optimizer.set_master(flat)  # exact fp64 master + fp32 mirror
            print('Actor loaded ' + ckpt)

    def save(self, checkpoint_path):
        flat, ws, o = self.optimizer.get_master(), [], 0
        for s in self.net.specs:
            nW = s.K * s.f_in * s.f_out
            ws.append((flat[o:o + nW].reshape(s.K, s.f_in, s.f_out), flat[o + nW:o + nW + s.f_out]))
            o += s.n_params
        tf_bundle.save_weights(checkpoint_path, ws)

    def makestate(self, adj, node_features):
        return {"node_features": node_features, "support": adj}

    def memorize(self, grad, loss, reward):
        self.memory.append((grad.clone() if hasattr(grad, "clone") else grad.copy(), loss, reward))

    # ---- GNN ---------------------------------------------------------------------------------
    def _batch_of(self, adj):
        """CSR conversion of the support (spektral.utils.sp_matrix_to_sp_tensor, :148).  The extended line
        graph is identical for the 10 instances of a file, so the device copy is cached on its structure."""
        A = sp.csr_matrix(adj)
        key = (A.shape[0], A.nnz, hash(A.indptr.tobytes()), hash(A.indices.tobytes()), hash(A.data.tobytes()))
        b = self._adj_cache.get(key)
        if b is None:
            if len(self._adj_cache) > 64:
                self._adj_cache.clear()
            b = GraphBatch.from_scipy([A], device=self.device)
            self._adj_cache[key] = b
        return b

    def predict(self, state, save=False):
        """:144-150 - returns the (n_ext, 1) output as a torch tensor on the device."""
        import torch
        batch = self._batch_of(state["support"])
        X = torch.as_tensor(np.ascontiguousarray(state["node_features"], dtype=np.float32), device=self.device)
        if save:
            Y, saved = self.net.forward(batch, X, save=True)
            self._tape = dict(batch=batch, X=X, Y=Y, saved=saved)
            return Y
        return self.net.forward(batch, X)

    def predict_batch(self, states):
        """Many (network, instance) states in one launch - the batched form the reference lacks."""
        import torch
        mats = [sp.csr_matrix(s["support"]) for s in states]
        batch = GraphBatch.from_scipy(mats, device=self.device)
        X = np.concatenate([np.asarray(s["node_features"], dtype=np.float32) for s in states], axis=0)
        Y = self.net.forward(batch, torch.as_tensor(X, device=self.device))
        return [Y[a:b] for a, b in zip(batch.graph_off[:-1], batch.graph_off[1:])]

    def act(self, state, save=False):
        return self.predict(state, save=save)

    def replay(self, batch_size):
        """:156-169 - sample batch_size memorised gradients and apply them one after the other."""
        import torch
        if len(self.memory) < batch_size:
            return float('NaN')
        self.reward_mem.clear()
        minibatch = random.sample(self.memory, batch_size)
        grads = torch.stack([g.reshape(-1) for g, _, _ in minibatch]).contiguous()
        self.optimizer.apply(grads)
        losses = [loss for _, loss, _ in minibatch]
        if self.epsilon > self.flags.epsilon_min:
            self.epsilon *= self.flags.epsilon_decay
        return np.nanmean(losses)

    # ---- forward: features -> GNN -> queue head -> delay matrix (:211-276) ----------------------
    def forward(self, obj, env, save=False):
        import networkx as nx
        import torch
        adj = nx.adjacency_matrix(obj.gi_ext)
        nn = obj.num_edges_ext
        node_features = np.zeros((nn, 4))
        node_features[:, 0] = obj.edge_self_loop
        node_features[:, 1] = obj.edge_rate_ext
        node_features[:, 2] = obj.jobs_arrivals
        node_features[:, 3] = obj.edge_as_server
        state = self.makestate(adj, node_features)
        lambda_array = self.act(state, save=save)

        fused = str(self.device).startswith("cuda") and hasattr(self.net, "ctx")
        if fused:
            # fused fp64 head kernels (csrc/queue_head.cu); the per-network constants are cached on the device
            hi = qh.HeadInputs(obj, env, None)
            # every constant the cached device copy holds is part of the key (id(env) is not: ids are recycled)
            import scipy.sparse as sp_
            adj_i = sp_.csr_matrix(hi.adj_i_host)
            key = (hi.link_rates_host.tobytes(), hi.node_mu_host.tobytes(), hi.maps_ol_el_host.tobytes(), hi.maps_on_el_host.tobytes(),
                   hi.cf_degs_host.tobytes(), adj_i.indptr.tobytes(), adj_i.indices.tobytes(), float(hi.T), int(nn))
            hb = self._head_cache.get(key)
            if hb is None:
                if len(self._head_cache) > 32:
                    self._head_cache.clear()
                hb = qh.HeadBatch([hi], [nn], self.net.ctx, self.device)
                self._head_cache[key] = hb
            ld, nd = hb.forward(lambda_array, save=save)
            head_tape = hb.last_tape if save else None   # (lam, saved_mu) of THIS call; the HeadBatch itself is shared
            link_delay, node_delay = ld.reshape(-1, 1), nd.reshape(-1, 1)
            lam64 = None
        else:
            hi = qh.HeadInputs(obj, env, self.device)
            hb = None
            head_tape = None
            lam64 = lambda_array.detach().to(torch.float64)
            if save:
                lam64.requires_grad_(True)
            link_lambda = lam64[hi.maps_ol_el]
            node_lambda = lam64[hi.maps_on_el]
            link_delay, node_delay = qh.queue_delays(link_lambda, node_lambda, hi.link_rates, hi.cf_degs, hi.node_mu,
                                                     hi.adj_i, hi.T)
        delay_mtx_ts, delay_mtx_np = qh.delay_matrices(link_delay, node_delay, hi, self.bug_compatible)
        if save:
            self._tape.update(lam64=lam64, link_delay=link_delay, node_delay=node_delay, hi=hi, hb=hb, head_tape=head_tape)
        return state, delay_mtx_ts, delay_mtx_np

    def _on_gpu(self):
        return str(self.device).startswith("cuda") and hasattr(self.net, "ctx")

    # ---- the instances of ONE network file in one launch each -------------------------
    @staticmethod
    def instance_features(obj):
        """The four feature columns of :218-226 for one job instance (only column 2 differs between instances)."""
        X = np.zeros((obj.num_edges_ext, 4))
        X[:, 0] = obj.edge_self_loop
        X[:, 1] = obj.edge_rate_ext
        X[:, 2] = obj.jobs_arrivals
        X[:, 3] = obj.edge_as_server
        return X

    def forward_instances(self, obj, env, feature_list, cpu_apsp=None):
        """forward() + the shortest-path matrices of forward_env() (:278-287) for B job instances of one network: ONE
        mho_cheb_forward launch over B copies of the extended line graph, ONE fused queue-head launch, ONE mho_apsp
        launch.  obj / env: any instance of the network (topology, rates and maps are instance-invariant,
        offloading_v3.py:262-339).  Returns a list of (delay_mtx_np, sp_gnn, sp_hop), one per instance - what
        env_step_from() consumes."""
        import networkx as nx
        import torch
        B = len(feature_list)
        nn = obj.num_edges_ext
        A = sp.csr_matrix(nx.adjacency_matrix(obj.gi_ext))
        if not self._on_gpu():
            # host stand-in (tests): the same quantities instance by instance through forward()'s own code path
            out = []
            hi = qh.HeadInputs(obj, env, self.device)
            for X in feature_list:
                lam64 = self.act(self.makestate(A, X)).detach().to(torch.float64)
                ld, nd = qh.queue_delays(lam64[hi.maps_ol_el], lam64[hi.maps_on_el], hi.link_rates, hi.cf_degs, hi.node_mu, hi.adj_i, hi.T)
                _, D_np = qh.delay_matrices(ld, nd, hi, self.bug_compatible)
                for (src, dst) in env.graph_c.edges:
                    env.graph_c[src][dst]["delay"] = D_np[src, dst]
                out.append((D_np, cpu_apsp(env.graph_c, weight="delay"), cpu_apsp(env.graph_c, weight=None)))
            return out
        key = ("inst", B, A.shape[0], A.nnz, hash(A.indptr.tobytes()), hash(A.indices.tobytes()), hash(A.data.tobytes()))
        batch = self._adj_cache.get(key)
        if batch is None:
            if len(self._adj_cache) > 64:
                self._adj_cache.clear()
            batch = GraphBatch.from_scipy([A] * B, device=self.device)
            self._adj_cache[key] = batch
        X = torch.as_tensor(np.ascontiguousarray(np.concatenate(feature_list, axis=0), dtype=np.float32), device=self.device)
        lam = self.net.forward(batch, X)                                   # [B * nn, 1]
        hi = qh.HeadInputs(obj, env, None)
        adj_i = sp.csr_matrix(hi.adj_i_host)
        hkey = ("inst", B, hi.link_rates_host.tobytes(), hi.node_mu_host.tobytes(), hi.maps_ol_el_host.tobytes(), hi.maps_on_el_host.tobytes(),
                hi.cf_degs_host.tobytes(), adj_i.indptr.tobytes(), adj_i.indices.tobytes(), float(hi.T), int(nn))
        hb = self._head_cache.get(hkey)
        if hb is None:
            if len(self._head_cache) > 32:
                self._head_cache.clear()
            hb = qh.HeadBatch([hi] * B, [nn] * B, self.net.ctx, self.device)
            self._head_cache[hkey] = hb
        ld, nd = hb.forward(lam, save=False)
        L, nc = hi.num_links, len(hi.comp_nodes)
        from .apsp import ApspPlan
        plan = env.__dict__.get("_mho_apsp_inst")
        if plan is None or plan.n_graphs != B or plan.n_edges != env.graph_c.number_of_edges():
            plan = ApspPlan([env.graph_c] * B, device=self.device, ctx=self.net.ctx)
            plan.n_edges = env.graph_c.number_of_edges()
            env.__dict__["_mho_apsp_inst"] = plan
        mats = []
        for i in range(B):
            _, D_np = qh.delay_matrices(ld[i * L:(i + 1) * L].reshape(-1, 1), nd[i * nc:(i + 1) * nc].reshape(-1, 1), hi, self.bug_compatible)
            mats.append(D_np)
        sps = plan.lengths(plan.entry_weights(mats))
        hop = self._shortest_paths(env, mats[0], cpu_apsp)[1]              # topology only: cached on the env
        return [(mats[i], sps[i], hop.copy()) for i in range(B)]

    @staticmethod
    def env_step_from(pre, env):
        """The environment half of forward_env() (:281-290) from a precomputed (delay_mtx_np, sp_gnn, sp_hop)."""
        delay_mtx_np, sp_gnn, sp_hop = pre
        for (src, dst) in env.graph_c.edges:
            env.graph_c[src][dst]["delay"] = delay_mtx_np[src, dst]
        sp_gnn = np.array(sp_gnn, copy=True)
        np.fill_diagonal(sp_gnn, np.diagonal(delay_mtx_np))
        env.offloading(sp_gnn, sp_hop)
        return env.run()

    def _shortest_paths(self, env, delay_mtx_np, cpu_apsp):
        """sp_gnn / sp_hop of :286-287.  On a CUDA device: mho_apsp (bit-identical to the reference's Dijkstra); the CSR
        of env.graph_c and the hop-count matrix depend on the topology only and are cached on the env object."""
        if not self._on_gpu():
            return cpu_apsp(env.graph_c, weight="delay"), cpu_apsp(env.graph_c, weight=None)
        from .apsp import ApspPlan
        plan = env.__dict__.get("_mho_apsp")
        if plan is None or plan.n_edges != env.graph_c.number_of_edges():
            plan = ApspPlan(env.graph_c, device=self.device, ctx=self.net.ctx)
            plan.n_edges = env.graph_c.number_of_edges()
            env.__dict__["_mho_apsp"] = plan
        sp_gnn = plan.lengths(plan.entry_weights(delay_mtx_np))[0]
        return sp_gnn, plan.hops()[0].copy()

    def forward_env(self, obj, env):
        """:278-291."""
        all_pairs_shortest_paths = _apsp_or_none() if self._on_gpu() else _apsp()
        state, delay_mtx_ts, delay_mtx_np = self.forward(obj, env)
        for (src, dst) in env.graph_c.edges:
            env.graph_c[src][dst]["delay"] = delay_mtx_np[src, dst]
        delay_servers = np.diagonal(delay_mtx_np)
        sp_gnn, sp_hop = self._shortest_paths(env, delay_mtx_np, all_pairs_shortest_paths)
        np.fill_diagonal(sp_gnn, delay_servers)
        decisions, delay_est = env.offloading(sp_gnn, sp_hop)
        delay_links_gnn, delay_nodes_gnn, delay_unit_gnn = env.run()
        return delay_links_gnn, delay_nodes_gnn, delay_unit_gnn

    @staticmethod
    def _link_index(obj):
        idx = {}
        for i, (a, b) in enumerate(obj.link_list_ext):
            idx.setdefault((a, b), i)
        return idx

    def forward_backward(self, obj, env, explore=0.0):
        """:293-453 - forward, environment step, analytic critic, route gradient, VJP to the weights."""
        import torch
        all_pairs_shortest_paths = _apsp_or_none() if self._on_gpu() else _apsp()
        state, delay_mtx_ts, delay_mtx_np = self.forward(obj, env, save=True)
        for (src, dst) in env.graph_c.edges:
            env.graph_c[src][dst]["delay"] = delay_mtx_np[src, dst]
        delay_servers = np.diagonal(delay_mtx_np)
        sp_gnn, sp_hop = self._shortest_paths(env, delay_mtx_np, all_pairs_shortest_paths)
        np.fill_diagonal(sp_gnn, delay_servers)
        decisions, delay_est = env.offloading(sp_gnn, sp_hop, explore)
        delay_links_gnn, delay_nodes_gnn, delay_unit_gnn = env.run()

        # routes (edges_ext x jobs), :310-331.  (a dict replaces the reference's list.index searches;
        # first occurrence wins, like list.index)
        lidx_of = self._link_index(obj)

        def find(n0, n1):
            if (n0, n1) in lidx_of:
                return lidx_of[(n0, n1)]
            if (n1, n0) in lidx_of:
                return lidx_of[(n1, n0)]
            raise ValueError("Link not exist, check route")

        routes_np = np.zeros((obj.num_edges_ext, env.num_jobs))
        jobs_load = np.zeros((env.num_jobs, 1))
        jobs_data = np.zeros((1, env.num_jobs))
        for i in range(env.num_jobs):
            src = env.jobs[i].source_node
            jobs_load[i, 0] += env.jobs[i].arrival_rate * env.jobs[i].ul_data
            jobs_data[0, i] += env.jobs[i].ul_data + env.jobs[i].dl_data
            n0 = src
            if n0 != env.flows[i].dst:
                for n1 in env.flows[i].route[1:]:
                    routes_np[find(n0, n1), i] = 1
                    n0 = n1
            routes_np[lidx_of[(n0, n0 + env.num_nodes)], i] = 1

        # critic with a nested tape, :333-374
        hi_cpu = qh.HeadInputs(obj, env, "cpu")
        loss_fn, grad_routes_np, delay_job_edge, unit_delay_edge = qh.critic(routes_np, jobs_load, jobs_data, obj, hi_cpu,
                                                                              obj.num_edges_ext)

        # gradient toward distances, "method 2" :384-416: bias[l_m, j] = sum of unit delays from the
        # destination's compute edge back to hop m, so d(sum -grad_routes * bias)/d unit_delay[l_i] is the
        # suffix sum of -grad_routes over the hops at or after i in the walk (closed form of the `gl` tape)
        grad_edge_np = np.zeros((obj.num_edges_ext,))
        for jidx in range(env.num_jobs):
            job, flow = env.jobs[jidx], env.flows[jidx]
            n1 = flow.dst + env.num_nodes
            walk = []
            for n0 in reversed(flow.route):
                walk.append(find(n0, n1))
                if n0 == job.source_node:
                    break
                n1 = n0
            suffix = 0.0
            for lidx in reversed(walk):
                suffix += -grad_routes_np[lidx, jidx]
                grad_edge_np[lidx] += suffix
        grad_dist_np = np.zeros_like(delay_mtx_np)
        for lidx in range(len(obj.link_list_ext)):
            n0, n1 = obj.link_list_ext[lidx]
            if n1 >= env.num_nodes:
                grad_dist_np[n0, n0] = grad_edge_np[lidx]
            else:
                grad_dist_np[n0, n1] = grad_edge_np[lidx]
                grad_dist_np[n1, n0] = grad_edge_np[lidx]

        # MSE term, :440-444
        delay_unit_gnn[np.isinf(delay_unit_gnn)] = np.nan
        loss_mse = np.nanmean((delay_mtx_np - delay_unit_gnn) ** 2)
        grad_dist_np += np.nan_to_num(0.001 * (delay_mtx_np - delay_unit_gnn), nan=0.0)

        gradients = self.vjp_from_grad_dist(grad_dist_np)
        self.memorize(gradients, loss_fn, loss_mse)
        self._tape = None

        flows_gnn = deepcopy(env.flows)
        return delay_mtx_np, delay_links_gnn, delay_nodes_gnn, delay_unit_gnn, flows_gnn, loss_fn, loss_mse

    def vjp_from_grad_dist(self, grad_dist_np):
        """``g.gradient(delay_mtx_ts, weights, output_gradients=grad_dist_np)`` (:448) for the forward taped by
        ``forward(obj, env, save=True)``: grad_dist -> scatter of :260-274 -> queue head (torch autograd, fp64)
        -> libmho's ChebConv VJP.  Returns the flat gradient (kernel_0, bias_0, ...) as a device tensor."""
        import torch
        tape = self._tape
        if tape is None or "hi" not in tape:
            raise RuntimeError("no taped forward: call forward(obj, env, save=True) first")
        g_ld, g_nd = qh.seed_from_grad_dist(grad_dist_np, tape["hi"], self.device)
        if tape["hb"] is not None:
            dY = tape["hb"].backward(g_ld, g_nd, tape=tape.get("head_tape")).contiguous()
        else:
            (g_lam,) = torch.autograd.grad([tape["link_delay"], tape["node_delay"]], tape["lam64"], [g_ld, g_nd])
            dY = g_lam.to(torch.float32).contiguous()
        gpg, _, _ = self.net.backward(tape["batch"], tape["X"], tape["Y"], tape["saved"], dY, need_sum=False)
        return gpg[0]

    # TensorBoard helpers of the reference (:455-468) are dead code there; kept as no-ops for drop-in use.
    def log_init(self):
        pass

    def log_scalar(self, name, variable, step, test=False):
        pass
