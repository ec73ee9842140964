"""Generate tests/golden/*.npz (TEST INFRASTRUCTURE; run in the build container only).

Inputs come from the reference itself: shipped networks (data/aco_data_ba_10/*.mat) run through
the reference's own ``AdhocCloud.graph_expand()`` (imported from /root/reference/src, not
copied), and the shipped checkpoints (model/model_ChebConv_BAT{800,950}_a5_c5_ACO_agent).
Outputs are this oracle's fp64 results (the reference's TF/Spektral arithmetic cannot be
installed here - see chebnet_oracle.py header).

    python oracle/make_golden.py          # rewrites tests/golden/
"""
from __future__ import annotations

import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import chebnet_oracle as O  # noqa: E402
import ref_env  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
FILES = ["aco_case_seed500_m2_n20_s4.mat", "aco_case_seed501_m2_n50_s", "aco_case_seed502_m2_n80_s",
         "aco_case_seed503_m2_n100_s", "aco_case_seed504_m2_n110_s", "aco_case_seed505_m2_n30_s"]


def csr_parts(A):
    A = sp.csr_matrix(A); A.sort_indices()
    return A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)


def main():
    os.makedirs(OUT, exist_ok=True)
    datadir = os.path.join(ref_env.REF_ROOT, "data", "aco_data_ba_10")
    names = sorted(os.listdir(datadir))
    for tag in ("BAT800", "BAT950"):
        ws = O.load_reference_weights(os.path.join(ref_env.REF_ROOT, "model", "model_ChebConv_%s_a5_c5_ACO_agent" % tag))
        np.savez_compressed(os.path.join(OUT, "weights_%s.npz" % tag),
                            **{"W%d" % i: W for i, (W, _) in enumerate(ws)},
                            **{"b%d" % i: b for i, (_, b) in enumerate(ws)})
    ws = O.load_reference_weights(os.path.join(ref_env.REF_ROOT, "model", "model_ChebConv_BAT800_a5_c5_ACO_agent"))
    rng = np.random.default_rng(7)
    # K=3 variant of the same architecture (the --K extension), fixed-seed weights
    ws_k3 = O.glorot_weights([4, 32, 32, 32, 32, 1], 3, np.random.default_rng(11))
    ws_k3 = [(W * 0.25, b + 0.05) for W, b in ws_k3]
    np.savez_compressed(os.path.join(OUT, "weights_K3.npz"),
                        **{"W%d" % i: W for i, (W, _) in enumerate(ws_k3)},
                        **{"b%d" % i: b for i, (_, b) in enumerate(ws_k3)})
    np.random.seed(20260921)
    for ci, pat in enumerate(FILES):
        fn = [n for n in names if n.startswith(pat)][0]
        env, nodes_info = ref_env.build_env(os.path.join(datadir, fn), T=1000)
        ref_env.sample_jobs(env, nodes_info, 0.15)
        obj = env.graph_expand()
        adj, X = ref_env.gnn_inputs(obj)
        rp, cidx, vals = csr_parts(adj)
        # --- shipped K=1 model
        lam, cache = O.cheb_stack_forward(adj, X, ws, return_cache=True)
        ld, nd, hc = O.queue_head_forward(lam, obj.maps_ol_el, obj.maps_on_el, env.link_rates, env.cf_degs,
                                          env.proc_bws, env.adj_i, env.T, return_cache=True)
        comp_nodes = np.nonzero(env.proc_bws > 0)[0]
        edges = np.array(list(env.graph_c.edges), dtype=np.int32)
        D_bug = O.delay_matrix(ld, nd, env.num_nodes, edges, env.link_matrix, comp_nodes, True)
        D_ts = O.delay_matrix(ld, nd, env.num_nodes, edges, env.link_matrix, comp_nodes, False)
        # --- VJP seeded with a synthetic grad_dist (N x N), as :448 does
        gD = rng.normal(size=(env.num_nodes, env.num_nodes)) * 0.01
        g_ld = np.zeros_like(ld); g_nd = np.zeros_like(nd)
        for (e0, e1) in edges:  # delay_mtx_ts scatters link_delay to [e0,e1] and [e1,e0]
            li = env.link_matrix[e0, e1]
            g_ld[li, 0] += gD[e0, e1] + gD[e1, e0]
        g_nd[:, 0] = gD[comp_nodes, comp_nodes]
        g_lam = O.queue_head_vjp(hc, g_ld, g_nd, lam.shape[0], obj.maps_ol_el, obj.maps_on_el)
        grads, _ = O.cheb_stack_backward(adj, ws, cache, g_lam)
        # --- K=3 variant on the same graph (forward + backward with a direct dY seed)
        lam3, cache3 = O.cheb_stack_forward(adj, X, ws_k3, return_cache=True)
        dY3 = rng.normal(size=lam3.shape)
        grads3, dX3 = O.cheb_stack_backward(adj, ws_k3, cache3, dY3)
        irp, icidx, ivals = csr_parts(env.adj_i)
        np.savez_compressed(
            os.path.join(OUT, "case%d.npz" % ci), filename=fn, num_nodes=env.num_nodes, T=env.T,
            rowptr=rp, colidx=cidx, vals=vals, X=X, lam=lam,
            maps_ol_el=obj.maps_ol_el, maps_on_el=obj.maps_on_el, link_rates=env.link_rates,
            cf_degs=env.cf_degs, proc_bws=env.proc_bws, adj_i_rowptr=irp, adj_i_colidx=icidx, adj_i_vals=ivals,
            link_delay=ld, node_delay=nd, edges=edges, link_matrix=env.link_matrix, comp_nodes=comp_nodes,
            delay_mtx_bug=D_bug, delay_mtx_ts=D_ts, grad_dist=gD, g_lam=g_lam,
            grad_flat=O.flatten_params(grads),
            lam_K3=lam3, dY_K3=dY3, grad_flat_K3=O.flatten_params(grads3), dX_K3=dX3)
        print(fn, "n_ext", lam.shape[0], "nnz", adj.nnz, "lam range", lam.min(), lam.max())

    # --- synthetic headline layer: K=5, 32->32, BA graphs of several sizes
    sizes = [20, 30, 50, 70, 100, 110, 64, 33]
    mats = O.make_batch(sizes, seed0=1000)
    g_off, rp, cidx, vals = O.concat_batch(mats)
    rng = np.random.default_rng(1)
    Xs = rng.normal(size=(g_off[-1], 32))
    W = O.glorot_weights([32, 32], 5, np.random.default_rng(2))[0][0]
    b = np.random.default_rng(3).normal(size=32) * 0.1
    Ablk = sp.block_diag(mats, format="csr")
    Y = O.cheb_layer_forward(Ablk, Xs, W, b, O.ACT_LEAKY, 0.2)
    np.savez_compressed(os.path.join(OUT, "layer_K5_F32.npz"), sizes=np.array(sizes), graph_off=g_off,
                        rowptr=rp, colidx=cidx, vals=vals, X=Xs, W=W, b=b, Y=Y)
    print("layer_K5_F32", Y.shape, np.abs(Y).max())


if __name__ == "__main__":
    main()
