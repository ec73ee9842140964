"""Data parallelism over independent graph instances: one process per GPU.

Forward / rollout: network files (or graphs of a batch) are sharded across ranks, no collective on the
data path.  Training: the only exchange is the flat gradient buffer (3 361 floats for the shipped
architecture) - an all-gather when the reference's replay semantics must be kept bit-for-bit on every
rank (``dp_mode=replay``), an all-reduce for a classic averaged step (``dp_mode=allreduce``).
NCCL over NVLink on GPUs; the same code runs on gloo/CPU tensors in the tests.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)


def init_from_env(device=None):
    """Initialise torch.distributed from torchrun's environment (no-op for a single process)."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws <= 1 or dist.is_initialized():
        return world()
    use_cuda = torch.cuda.is_available()
    if use_cuda:
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")
    return world()


def shard(items, rank=None, world_size=None):
    """Round-robin shard of a list of work items (files, graphs): rank r takes items[r::world]."""
    if rank is None:
        rank, world_size = world()
    return list(items[rank::world_size])


def allreduce_mean_(flat):
    """In-place mean over ranks of a flat gradient buffer (ncclAllReduce SUM, then scale)."""
    r, w = world()
    if w > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(w)
    return flat


def allgather_rows(local, extra=None):
    """All-gather a variable number of rows per rank.  local: [n_r, P]; extra: [n_r, E] or None.
    Returns the concatenation over ranks in rank order (identical on every rank)."""
    r, w = world()
    if w == 1:
        return (local, extra) if extra is not None else local
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n) for _ in range(w)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    mx = max(counts) if counts else 0

    def gather(t):
        pad = torch.zeros((mx, t.shape[1]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        bufs = [torch.zeros_like(pad) for _ in range(w)]
        dist.all_gather(bufs, pad)
        return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)

    out = gather(local)
    if extra is not None:
        return out, gather(extra)
    return out


def broadcast_(t, src=0):
    r, w = world()
    if w > 1:
        dist.broadcast(t, src=src)
    return t


def gather_objects(obj, dst=0):
    """Gather picklable per-rank results (CSV rows) on rank dst; returns the list there, None elsewhere."""
    r, w = world()
    if w == 1:
        return [obj]
    out = [None] * w if r == dst else None
    dist.gather_object(obj, out, dst=dst)
    return out


def bench_exchange(torch, dist_, dev, world_size, barrier, iters=20):
    """Device-timed cost of AdHoc_train's gradient exchange (the allreduce site of gnn_offloading_agent.py:156-169) on
    NCCL tensors, for the shipped architecture (3 361 parameters) and its K = 5 variant (16 289): per optimizer step
      allreduce: ncclAllReduce(SUM) of the flat mean gradient + one mho_adam_replay step;
      replay:    all-gather of the 10 gradient rows (+ loss / reward) every rank memorised for one file, then the
                 reference's sequential replay of 100 stored gradients in one mho_adam_replay launch.
    CUDA events on the current stream, max over ranks.  Returns a dict of microseconds."""
    from .chebnet import ChebNet, reference_stack
    from .optim import KerasAdamReplay
    out = {"n_ranks": world_size, "unit": "us per exchange (CUDA events, max over ranks)"}

    def timed(fn):
        for _ in range(3):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) * 1e3 / iters], dtype=torch.float64, device=dev)
        if world_size > 1:
            dist_.all_reduce(t, op=dist_.ReduceOp.MAX)
        return float(t.item())

    for K in (1, 5):
        net = ChebNet(reference_stack(K=K), device=dev, private_context=True)
        opt = KerasAdamReplay(net, learning_rate=1e-6)
        P = net.n_params
        g1 = torch.randn((1, P), device=dev) * 1e-3
        rows = torch.randn((10, P), device=dev) * 1e-3
        meta = torch.zeros((10, 2), device=dev)
        pool = torch.randn((100, P), device=dev) * 1e-3

        def ar_only():
            allreduce_mean_(g1)

        def ar_step():
            allreduce_mean_(g1)
            opt.apply(g1)

        def gather_only():
            allgather_rows(rows, meta)

        def replay_step():
            allgather_rows(rows, meta)
            opt.apply(pool)

        out["params_%d" % P] = {"allreduce": timed(ar_only), "allreduce_plus_adam_step": timed(ar_step),
                               "allgather_10_rows_per_rank": timed(gather_only), "allgather_plus_replay_100": timed(replay_step)}
    return out
