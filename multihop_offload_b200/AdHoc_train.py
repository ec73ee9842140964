"""AdHoc_train - the reference's training driver (src/AdHoc_train.py) on the B200-native agent.

Same flags, checkpoint naming (cp-{epoch:04d}.ckpt under model_ChebConv_{training_set}_a5_c5_ACO_agent,
:59,:204-206) and CSV (:40-46).  Under torchrun the shuffled files of an epoch are sharded over the ranks;
after every round of files the ranks exchange what they memorised:
  --dp_mode=replay     all-gather the new per-instance gradients, every rank then runs the reference's
                       replay (100 stored gradients applied one by one) with a shared RNG -> identical weights
  --dp_mode=allreduce  each rank sums batch/world sampled gradients, one NCCL all-reduce of the flat
                       buffer, ONE Adam step with the mean (classic data-parallel step; a different algorithm)
"""
from __future__ import absolute_import, division, print_function

import os
import random
import time

import numpy as np
import pandas as pd

from . import parallel
from .drivers_common import import_reference_env, load_case, result_row, run_method, sample_jobs
from .gnn_offloading_agent import ACOAgent, FLAGS

COLUMNS = ["fid", "filename", "seed", "num_nodes", "m", "num_mobile", "num_servers", "num_relays", "num_jobs",
           "n_instance", "method", "runtime", "gap_2_bl", "gnn_bl_ratio", "tau", "congest_jobs"]


def sync_and_replay(agent, batch_size, n_new, dp_mode, step_seed):
    """Exchange the gradients memorised since the last call, then update the weights."""
    import torch
    rank, world = parallel.world()
    if world == 1:
        return agent.replay(batch_size)
    new = list(agent.memory)[len(agent.memory) - n_new:] if n_new else []
    P = agent.net.n_params
    g = torch.stack([x[0].reshape(-1) for x in new]) if new else torch.zeros((0, P), device=agent.device)
    meta = torch.tensor([[float(x[1]), float(x[2])] for x in new], dtype=torch.float32,
                        device=agent.device).reshape(-1, 2)
    if dp_mode == "replay":
        for _ in range(n_new):
            agent.memory.pop()
        g_all, meta_all = parallel.allgather_rows(g.contiguous(), meta)
        for i in range(g_all.shape[0]):
            agent.memory.append((g_all[i].clone(), float(meta_all[i, 0]), float(meta_all[i, 1])))
        random.seed(step_seed)  # same sample on every rank -> identical sequential replay
        return agent.replay(batch_size)
    # allreduce mode
    share = max(1, batch_size // world)
    if len(agent.memory) < share:
        flag = torch.zeros(1, device=agent.device)
    else:
        flag = torch.ones(1, device=agent.device)
    parallel.allreduce_mean_(flag)
    if float(flag.item()) < 1.0:
        return float('NaN')
    mini = random.sample(agent.memory, share)
    gsum = torch.stack([x[0].reshape(-1) for x in mini]).mean(0, keepdim=True).contiguous()
    parallel.allreduce_mean_(gsum)
    agent.optimizer.apply(gsum)
    if agent.epsilon > agent.flags.epsilon_min:
        agent.epsilon *= agent.flags.epsilon_decay
    return float(np.nanmean([x[1] for x in mini]))


def main():
    rank, world = parallel.init_from_env()
    if world > 1:
        FLAGS.device = "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0"))
    if FLAGS.seed >= 0:
        np.random.seed(FLAGS.seed + rank)
        random.seed(FLAGS.seed)
    AdhocCloud, apsp = import_reference_env(FLAGS.ref_src)
    agent = ACOAgent(FLAGS, 5000)
    arrival_scale, T, batch_size, datapath = FLAGS.arrival_scale, FLAGS.T, FLAGS.batch, FLAGS.datapath
    val_mat_names = sorted(os.listdir(datapath))
    if FLAGS.max_files > 0:
        val_mat_names = val_mat_names[:FLAGS.max_files]
    os.makedirs(FLAGS.out, exist_ok=True)
    output_csv = os.path.join(FLAGS.out, "aco_training_data_{}_load_{:.2f}_T_{}.csv".format(
        datapath.rstrip("/").split("/")[-1], arrival_scale, T))
    actor_model = os.path.join(FLAGS.modeldir, 'model_ChebConv_{}_a{}_c{}_ACO_agent'.format(FLAGS.training_set, 5, 5))
    os.makedirs(FLAGS.modeldir, exist_ok=True)
    try:
        agent.load(actor_model)
    except Exception as e:
        print("unable to load {} ({})".format(actor_model, e))
    if world > 1:  # identical starting weights everywhere
        import torch
        parallel.broadcast_(agent.optimizer.master)
        agent.net.params.copy_(agent.optimizer.master.to(torch.float32))
        agent.net.weights_changed()

    gidx, losses, rows = 0, [], []
    num_instances, explore, explore_decay = 10, 0.1, 0.99
    for epoch in range(FLAGS.epochs):
        order = np.random.RandomState(1000 + epoch).permutation(len(val_mat_names)) if world > 1 \
            else np.random.permutation(len(val_mat_names))
        # ranks walk the shuffled list in lock-step rounds of `world` files
        for r0 in range(0, len(order), world):
            mine = order[r0 + rank] if r0 + rank < len(order) else None
            n_new = 0
            if mine is not None:
                fname = val_mat_names[mine]
                env, nodes_info, seed, num_nodes, m = load_case(AdhocCloud, os.path.join(datapath, fname), T)
                for ni in range(num_instances):
                    num_jobs = sample_jobs(env, nodes_info, arrival_scale)
                    delay_dict = {}
                    for method in ["baseline", "local", "GNN", "GNN-test"]:
                        t0 = time.time()
                        delay_emp, _ = run_method(method, env, agent, apsp, explore)
                        runtime = time.time() - t0
                        delay_dict[method] = delay_emp
                        if method == "GNN":
                            n_new += 1
                        base = {"fid": gidx, "filename": fname, "seed": seed, "n_instance": ni, "num_nodes": num_nodes, "m": m}
                        rows.append(result_row(base, "method", method, runtime, delay_emp, delay_dict, env, num_jobs))
            loss = sync_and_replay(agent, batch_size, n_new, FLAGS.dp_mode, 7919 * epoch + r0)
            losses.append(loss)
            if rank == 0:
                print("{} Loss: {:.2f}, explore: {:.4f}".format(gidx, np.nanmean(losses) if losses else float('nan'), explore))
            if not np.isnan(loss):
                if rank == 0:
                    agent.save(os.path.join(actor_model, 'cp-{epoch:04d}.ckpt'.format(epoch=epoch)))
                explore = np.clip(explore * explore_decay, 0., 1.)
                losses = []
            gidx += 1
            if world == 1:
                pd.DataFrame(rows, columns=COLUMNS).to_csv(output_csv, index=False)
    gathered = parallel.gather_objects(rows)
    if rank == 0:
        pd.DataFrame([r for part in gathered for r in part], columns=COLUMNS).to_csv(output_csv, index=False)
    return 0


if __name__ == "__main__":
    main()
