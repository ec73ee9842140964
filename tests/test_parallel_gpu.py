"""NCCL gradient exchange of AdHoc_train (the allreduce site of gnn_offloading_agent.py:156-169) on real GPUs.
Runs only under a multi-process launch:  torchrun --nproc-per-node 2 -m pytest tests/test_parallel_gpu.py -m gpu -q"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
WORLD = int(os.environ.get("WORLD_SIZE", "1"))


@pytest.mark.skipif(WORLD < 2, reason="needs WORLD_SIZE >= 2 (torchrun), one rank per GPU")
def test_nccl_exchange_matches_single_process_replay():
    import torch
    import torch.distributed as dist
    from multihop_offload_b200 import ChebNet, parallel, reference_stack
    from multihop_offload_b200.optim import KerasAdamReplay
    rank, world = parallel.init_from_env()
    assert world == WORLD and dist.get_backend() == "nccl"
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    for K in (1, 5):
        specs = reference_stack(K=K)
        P = sum(s.n_params for s in specs)
        w0 = np.random.default_rng(1).normal(size=P) * 0.1               # identical start on every rank
        rows = [np.random.default_rng(100 + r).normal(size=(10, P)).astype(np.float32) * 1e-2 for r in range(world)]
        mine = torch.from_numpy(rows[rank]).to(dev)
        # ---- dp_mode = replay: all-gather the memorised rows, every rank replays the same sequence
        net = ChebNet(specs, device=dev, params=w0, private_context=True)
        opt = KerasAdamReplay(net, learning_rate=1e-3)
        g_all = parallel.allgather_rows(mine.contiguous())
        assert g_all.shape == (10 * world, P)
        opt.apply(g_all.contiguous())
        ref_net = ChebNet(specs, device=dev, params=w0, private_context=True)
        ref_opt = KerasAdamReplay(ref_net, learning_rate=1e-3)
        ref_opt.apply(torch.from_numpy(np.concatenate(rows)).to(dev))
        assert torch.equal(opt.master, ref_opt.master), "replay after the all-gather differs from the sequential replay"
        # ---- dp_mode = allreduce: mean gradient, one step; identical weights on every rank
        net2 = ChebNet(specs, device=dev, params=w0, private_context=True)
        opt2 = KerasAdamReplay(net2, learning_rate=1e-3)
        g = mine.mean(0, keepdim=True).contiguous()
        parallel.allreduce_mean_(g)
        want = torch.from_numpy(np.mean([r.mean(0) for r in rows], axis=0)).to(dev)
        assert torch.allclose(g[0], want, rtol=1e-5, atol=1e-8)
        opt2.apply(g)
        gathered = [torch.zeros_like(opt2.master) for _ in range(world)]
        dist.all_gather(gathered, opt2.master)
        for t in gathered[1:]:
            assert torch.equal(t, gathered[0]), "ranks diverged after the all-reduced step"
    # the timed series bench.py reports at N > 1
    out = parallel.bench_exchange(torch, dist, dev, world, lambda: (dist.barrier(), torch.cuda.synchronize()), iters=5)
    assert out["n_ranks"] == world and all(v > 0 for v in out["params_3361"].values())
    dist.barrier()
