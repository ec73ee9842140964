"""CPU, world_size 2 over gloo: the N>1 host logic (sharding, gradient all-gather / all-reduce, result
gathering) that the multi-GPU drivers and bench rely on.  No CUDA involved."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys, json
    sys.path.insert(0, %r)
    import numpy as np, torch
    import torch.distributed as dist
    dist.init_process_group("gloo")
    from multihop_offload_b200 import parallel
    rank, world = parallel.world()
    assert world == 2
    # 1. file sharding: disjoint, exhaustive
    items = list(range(11))
    mine = parallel.shard(items)
    gathered = parallel.gather_objects(mine)
    # 2. all-gather of a different number of per-instance gradients per rank
    P = 37
    n_local = 3 if rank == 0 else 5
    g = torch.arange(n_local * P, dtype=torch.float32).reshape(n_local, P) + 1000 * rank
    meta = torch.full((n_local, 2), float(rank))
    g_all, m_all = parallel.allgather_rows(g, meta)
    # 3. all-reduce mean of a flat buffer
    flat = torch.full((P,), float(rank + 1))
    parallel.allreduce_mean_(flat)
    # 4. broadcast of master weights
    w = torch.full((P,), float(rank + 7), dtype=torch.float64)
    parallel.broadcast_(w, 0)
    out = dict(rank=rank, g_shape=list(g_all.shape), g_sum=float(g_all.sum()), m=m_all[:, 0].tolist(),
               flat=float(flat[0]), w=float(w[0]))
    if rank == 0:
        out["gathered"] = gathered
    open(os.path.join(%r, "rank%%d.json" %% rank), "w").write(json.dumps(out))
    dist.destroy_process_group()
''')


def test_two_rank_gloo_collectives(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % (ROOT, str(tmp_path)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert res.returncode == 0, res.stderr[-2000:]
    import json
    outs = [json.loads((tmp_path / ("rank%d.json" % r)).read_text()) for r in range(2)]
    outs.sort(key=lambda o: o["rank"])
    assert outs[0]["gathered"] == [[0, 2, 4, 6, 8, 10], [1, 3, 5, 7, 9]]
    P = 37
    g0 = np.arange(3 * P, dtype=np.float64).reshape(3, P)
    g1 = np.arange(5 * P, dtype=np.float64).reshape(5, P) + 1000
    for o in outs:
        assert o["g_shape"] == [8, P]
        assert o["g_sum"] == g0.sum() + g1.sum()
        assert o["m"] == [0.0] * 3 + [1.0] * 5
        assert o["flat"] == 1.5
        assert o["w"] == 7.0


def test_graphbatch_shard_partitions_the_batch(built_lib):
    import scipy.sparse as sp
    from multihop_offload_b200 import GraphBatch
    rng = np.random.default_rng(0)
    mats = []
    for i in range(23):
        n = int(rng.integers(3, 60))
        A = sp.random(n, n, 0.2, random_state=i, format="csr"); A = A + A.T; A.data[:] = 1.0
        mats.append(sp.csr_matrix(A))
    b = GraphBatch.from_scipy(mats)
    tot_nodes = tot_nnz = 0
    ranges = []
    for r in range(4):
        s = b.shard(r, 4)
        tot_nodes += s.total_nodes; tot_nnz += s.total_nnz
        ranges.append(s.graph_range)
        if s.n_graphs:
            assert s.colidx.min() >= 0 and s.colidx.max() < s.total_nodes
    assert tot_nodes == b.total_nodes and tot_nnz == b.total_nnz
    assert ranges[0][0] == 0 and ranges[-1][1] == b.n_graphs
    assert all(ranges[i][1] == ranges[i + 1][0] for i in range(3))
