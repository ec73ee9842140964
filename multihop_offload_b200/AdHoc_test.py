"""AdHoc_test - the reference's test driver (src/AdHoc_test.py) on the B200-native agent.

Same flags (--datapath --out --T --arrival_scale --training_set ...), same checkpoint directory
naming (model_ChebConv_{training_set}_a5_c5_ACO_agent, :60), same CSV name and columns (:41-48).
    python -m multihop_offload_b200.AdHoc_test --datapath=../data/aco_data_ba_100 --arrival_scale=0.15 \
        --training_set=BAT800 --ref_src=/path/to/reference/src
Under torchrun the network files are sharded over the ranks (one GPU each, no collective on the data
path) and rank 0 writes the merged CSV.
"""
from __future__ import absolute_import, division, print_function

import os
import time

import numpy as np
import pandas as pd

from . import parallel
from .drivers_common import import_reference_env, load_case, lookahead_instances, result_row, run_method, sample_jobs
from .gnn_offloading_agent import ACOAgent, FLAGS

COLUMNS = ["filename", "seed", "num_nodes", "m", "num_mobile", "num_servers", "num_relays", "num_jobs", "n_instance",
           "Algo", "runtime", "tau", "congest_jobs", "gnn_bl_ratio", "gap_2_bl"]


def main():
    rank, world = parallel.init_from_env()
    if world > 1:
        FLAGS.device = "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0"))
    if FLAGS.seed >= 0:
        np.random.seed(FLAGS.seed + rank)
    AdhocCloud, apsp = import_reference_env(FLAGS.ref_src)
    agent = ACOAgent(FLAGS, 1000)
    arrival_scale, T, datapath = FLAGS.arrival_scale, FLAGS.T, FLAGS.datapath
    val_mat_names = sorted(os.listdir(datapath))
    if FLAGS.max_files > 0:
        val_mat_names = val_mat_names[:FLAGS.max_files]
    output_csv = os.path.join(FLAGS.out, "Adhoc_test_data_{}_load_{:.2f}_T_{}.csv".format(
        datapath.rstrip("/").split("/")[-1], arrival_scale, T))
    os.makedirs(FLAGS.out, exist_ok=True)
    actor_model = os.path.join(FLAGS.modeldir, 'model_ChebConv_{}_a{}_c{}_ACO_agent'.format(FLAGS.training_set, 5, 5))
    try:
        agent.load(actor_model)
    except Exception as e:  # the reference swallows load failures too (:63-66)
        print("unable to load {} ({})".format(actor_model, e))

    rows, num_instances = [], 10
    for fname in parallel.shard(val_mat_names, rank, world):
        env, nodes_info, seed, num_nodes, m = load_case(AdhocCloud, os.path.join(datapath, fname), T)
        t_case = time.time()
        batched = bool(FLAGS.batch_instances)
        pre, t_pre = None, 0.0
        if batched:   # the GNN side of all instances of this file in one launch each (forward only: the reference's test
            t0 = time.time()   # driver also runs a backward whose gradients it never uses, AdHoc_test.py:152)
            pre = lookahead_instances(env, nodes_info, arrival_scale, num_instances, agent, apsp)
            t_pre = (time.time() - t0) / num_instances
        for ni in range(num_instances):
            num_jobs = sample_jobs(env, nodes_info, arrival_scale)
            delay_dict = {}
            for method in ["baseline", "local", "GNN"]:
                t0 = time.time()
                if batched and method == "GNN":
                    delay_emp, _ = run_method("GNN-pre", env, agent, apsp, pre=pre[ni])
                    t0 -= t_pre
                else:
                    delay_emp, _ = run_method(method, env, agent, apsp)
                runtime = time.time() - t0
                delay_dict[method] = delay_emp
                base = {"filename": fname, "seed": seed, "n_instance": ni, "num_nodes": num_nodes, "m": m}
                rows.append(result_row(base, "Algo", method, runtime, delay_emp, delay_dict, env, num_jobs))
        print("Runtime {:.3f}s".format(time.time() - t_case),
              " for network of {} nodes, {} servers, {} relays".format(num_nodes, len(env.servers), len(env.relays)))
        if world == 1:
            pd.DataFrame(rows, columns=COLUMNS).to_csv(output_csv, index=False)
    gathered = parallel.gather_objects(rows)
    if rank == 0:
        allrows = [r for part in gathered for r in part]
        pd.DataFrame(allrows, columns=COLUMNS).to_csv(output_csv, index=False)
        print("wrote", output_csv, len(allrows), "rows")
    return 0


if __name__ == "__main__":
    main()
