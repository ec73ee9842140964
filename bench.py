#!/usr/bin/env python
"""bench.py - graph-steps/sec of the batched ChebConv forward (BASELINE.json metric).

Workload (BASELINE.json configs[1]): one ChebConv layer, K=5, 32 -> 32 features, bias +
leaky_relu, over a batch of 1024 Barabasi-Albert (m=2) graphs of 20..110 nodes.  A "step" is
one pass of the hot path over one such batch.  Weak scaling: every rank gets its own 1024 graphs.

    python bench.py --gpus 1 --steps 200 --warmup 20           # this repo's CUDA path
    python bench.py --impl reference --steps 3 --warmup 1      # the CPU restatement of the reference path
    torchrun --nproc-per-node N bench.py --gpus N ...          # one rank per GPU

Timing: CUDA events on the launching stream around exactly K back-to-back steps, barrier +
synchronize on both sides, max over ranks.  L2 hygiene: the timed loop rotates over R distinct
copies of the whole input/output set with R * bytes > 2 * 126 MB, so no step finds its inputs in L2.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "graph-steps/sec (ChebConv K=5 fwd, BA 20-110 nodes)"
UNIT = "graph-steps/s"
L2_BYTES = 126 * 1024 * 1024
SIZES = np.arange(20, 111, 10)


# --------------------------------------------------------------------------------------------
# workload (pure numpy/networkx; shared by both arms)
# --------------------------------------------------------------------------------------------
def ba_csr(n, seed):
    """CSR of networkx.barabasi_albert_graph(n, 2, seed) (generator of src/offloading_v3.py:40)."""
    import networkx as nx
    g = nx.barabasi_albert_graph(int(n), 2, seed=int(seed))
    indptr = np.zeros(n + 1, dtype=np.int64)
    cols = []
    for i in range(n):
        nb = sorted(g.adj[i])
        cols.append(np.asarray(nb, dtype=np.int64))
        indptr[i + 1] = indptr[i] + len(nb)
    return indptr, (np.concatenate(cols) if cols else np.zeros(0, dtype=np.int64))


def bind_to_gpu_numa_node(torch, local):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off (sysfs), so that cudaHostAlloc'd staging buffers
    are local to the GPU's PCIe root (one rank per GPU: 8 ranks otherwise crowd the first socket).  Best effort."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def make_workload(n_graphs, rank=0, fixed_n=None, K=5, F=32, pack=True):
    rng = np.random.default_rng(0 + 7919 * rank)
    sizes = np.full(n_graphs, fixed_n) if fixed_n else rng.choice(SIZES, size=n_graphs)
    if pack and not fixed_n:
        # same multiset of graphs, laid out in tile-packing order (first-fit decreasing): the order of independent
        # graph instances in a batch is the batch builder's choice
        from multihop_offload_b200.batch import pack_order
        perm = pack_order(sizes, 128)
        seeds = (1000 + np.arange(n_graphs) + 100003 * rank)[perm]
        sizes = sizes[perm]
    else:
        seeds = 1000 + np.arange(n_graphs) + 100003 * rank
    goff = np.zeros(n_graphs + 1, dtype=np.int64)
    rps, cis = [np.zeros(1, dtype=np.int64)], []
    noff = zoff = 0
    for i, n in enumerate(sizes):
        ip, ci = ba_csr(int(n), int(seeds[i]))
        rps.append(ip[1:] + zoff)
        cis.append(ci + noff)
        noff += int(n); zoff += ci.size
        goff[i + 1] = noff
    rowptr = np.concatenate(rps).astype(np.int32)
    colidx = np.concatenate(cis).astype(np.int32)
    X = np.random.default_rng(1 + rank).normal(size=(noff, F)).astype(np.float32)
    lim = np.sqrt(6.0 / (K * F + K * F))
    W = np.random.default_rng(2).uniform(-lim, lim, size=(K, F, F))
    b = np.zeros(F)
    return dict(sizes=sizes, graph_off=goff.astype(np.int32), rowptr=rowptr, colidx=colidx, X=X, W=W, b=b, K=K, F=F)


def algorithmic_bytes(w):
    """4 n F_in + 4 n F_out + 4 (n+1) + 8 nnz per graph (fp32 values, int32 ids;
    binary operator => the 4 B/nnz value stream is not read: 4 nnz; T_k stay on chip)."""
    n = int(w["graph_off"][-1]); nnz = int(w["rowptr"][-1]); B = len(w["sizes"]); F = w["F"]
    return 4 * n * F + 4 * n * F + 4 * (n + B) + 4 * nnz


def algorithmic_flops(w):
    n = int(w["graph_off"][-1]); nnz = int(w["rowptr"][-1]); F = w["F"]; K = w["K"]
    return 2 * nnz * F * (K - 1) + 2 * n * K * F * F


# --------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's per-graph fp64 path
# --------------------------------------------------------------------------------------------
def cpu_reference_rate(w, seconds=10.0, threads=0, max_passes=1000):
    """graph-steps/s of the reference's CPU path restated in C (fp64, one graph at a time as the
    reference's eager call does, graphs spread over `threads` OpenMP threads; 0 = all)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle
    ws = [(w["W"], w["b"])]
    X64 = w["X"].astype(np.float64)
    c_oracle.stack_forward(w["graph_off"], w["rowptr"], w["colidx"], None, ws, [2], 0.2, X64, threads)  # warm
    t0 = time.perf_counter(); passes = 0
    while True:
        c_oracle.stack_forward(w["graph_off"], w["rowptr"], w["colidx"], None, ws, [2], 0.2, X64, threads)
        passes += 1
        dt = time.perf_counter() - t0
        if dt >= seconds or passes >= max_passes:
            break
    cores = threads if threads > 0 else c_oracle.max_threads()
    return len(w["sizes"]) * passes / dt, cores, passes, dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w = make_workload(args.graphs, 0, args.fixed_n, pack=not args.no_pack)
    os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")   # idle OpenMP threads sleep instead of spinning (CPU quotas)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle
    ws = [(w["W"], w["b"])]
    X64 = w["X"].astype(np.float64)
    # all the host threads it can USE: more OpenMP threads than the container's CPU quota / the memory system can feed
    # make the pass slower, so the thread count is tuned first (one pass each, halving from the maximum)
    def one_pass(nt):
        t_ = time.perf_counter()
        c_oracle.stack_forward(w["graph_off"], w["rowptr"], w["colidx"], None, ws, [2], 0.2, X64, nt)
        return time.perf_counter() - t_
    cand, nt = [], c_oracle.max_threads()
    while nt >= 1:
        cand.append(nt)
        nt //= 2
    one_pass(cand[0])
    timing = {nt: min(one_pass(nt), one_pass(nt)) for nt in cand}
    cores = min(timing, key=timing.get)
    for _ in range(max(args.warmup, 1)):
        one_pass(cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_pass(cores)
    dt = time.perf_counter() - t0
    val = args.graphs * args.steps / dt
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, w),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "threads_tried": {str(k): round(v * 1e3, 2) for k, v in timing.items()},
                         "sample": "%d full passes over the %d-graph batch (oracle/cheb_oracle.c, fp64, one graph at a "
                                   "time, OpenMP over graphs)" % (args.steps, args.graphs)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference hot path is TensorFlow+Spektral (not installable offline): timed arm is the oracle port",
    }
    print(json.dumps(out))
    return 0


def workload_config(args, w):
    return {"workload": "configs[1]: ChebConv K=%d forward, %d->%d, bias+leaky_relu, batch %d BA(m=2) graphs, n in %s, "
                        "raw-adjacency operator" % (w["K"], w["F"], w["F"], args.graphs,
                                                    "{%d}" % args.fixed_n if args.fixed_n else "{20..110 step 10}"),
            "graphs_per_gpu": args.graphs, "nodes_per_gpu": int(w["graph_off"][-1]), "nnz_per_gpu": int(w["rowptr"][-1]),
            "parallelism": "graph-instance sharding, no data-path collective",
            "l2": "rotating over distinct input/output sets > 2x L2",
            "numa_node": getattr(args, "numa_node", None), "streams": "%d CUDA streams, one library context each; consecutive steps are independent batches and may overlap at their boundaries" % int(getattr(args, "streams", 1)),
            "batch_order": "graphs laid out in tile-packing order (first-fit decreasing, multihop_offload_b200.pack_order)"}


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from multihop_offload_b200 import ChebNet, GraphBatch, LayerSpec

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback "
                         "(use --impl reference for the CPU restatement)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    args.numa_node = bind_to_gpu_numa_node(torch, local)   # page-locked staging buffers land next to this rank's GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    w = make_workload(args.graphs, rank, args.fixed_n, pack=not args.no_pack)
    net = ChebNet([LayerSpec(w["K"], w["F"], w["F"], 2, 0.2)], device=dev)
    net.set_weights([(w["W"], w["b"])])
    n_nodes = int(w["graph_off"][-1])
    alg_bytes = algorithmic_bytes(w)
    R = max(2, int(np.ceil(2.2 * L2_BYTES / alg_bytes)))
    batches = [GraphBatch(w["graph_off"], w["rowptr"], w["colidx"], None, tile_rows=args.tile_rows, device=dev)
               for _ in range(R)]
    X0 = torch.from_numpy(w["X"]).to(dev)
    Xs = [X0] + [torch.randn_like(X0) for _ in range(R - 1)]
    Ys = [torch.empty((n_nodes, w["F"]), dtype=torch.float32, device=dev) for _ in range(R)]

    # Steps are independent batches.  With --streams S (default 2) they are issued round-robin on S CUDA streams, each
    # with its own library context (tile counters, weight images): stream order is kept inside a stream, and the tail
    # of one step (the last tiles of its 2.007 rounds) overlaps the head of the next instead of idling 146 SMs
    # (measured on B200: 22.8 / 17.6 / 17.6 us per step with 1 / 2 / 3 streams; the results are checked against a lone launch).
    n_str = max(1, int(args.streams))
    nets = [net]
    for _ in range(n_str - 1):
        n2 = ChebNet([LayerSpec(w["K"], w["F"], w["F"], 2, 0.2)], device=dev, private_context=True)
        n2.set_weights([(w["W"], w["b"])])
        nets.append(n2)
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_str)] if n_str > 1 else [torch.cuda.current_stream(dev)]

    def step(i):
        j = i % R
        k = i % n_str
        if n_str == 1:
            nets[0].forward(batches[j], Xs[j], out=Ys[j])
        else:
            with torch.cuda.stream(streams[k]):
                nets[k].forward(batches[j], Xs[j], out=Ys[j])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3) + n_str):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = sum(n_.ctx.launch_count() for n_ in nets)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    main = torch.cuda.current_stream(dev)
    ev0.record(main)
    if n_str > 1:
        for s_ in streams:
            s_.wait_event(ev0)
    for i in range(args.steps):
        step(i)
    if n_str > 1:
        for s_ in streams:
            done = torch.cuda.Event()
            done.record(s_)
            main.wait_event(done)
    ev1.record(main)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = sum(n_.ctx.launch_count() for n_ in nets) - l0
    # the overlapped launches computed what a lone launch computes (same buffers, same tiles)
    chk = [(args.steps - 1 - d) % R for d in range(min(n_str, args.steps))]
    kept = [Ys[j].clone() for j in chk]
    for j, y in zip(chk, kept):
        net.forward(batches[j], Xs[j], out=Ys[j])
        torch.cuda.synchronize()
        assert torch.equal(Ys[j], y), "multi-stream step differs from a single launch"
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: host buffers in, host buffers out, through the C-ABI host call
    from multihop_offload_b200._lib import PinnedArray, pinned_like
    goff_h, rp_h, ci_h = (pinned_like(np.ascontiguousarray(w[k], dtype=np.int32)) for k in ("graph_off", "rowptr", "colidx"))
    # two calls in flight, each with its own page-locked X / Y: the upload of step i+1 overlaps the download of step i
    Xh = [pinned_like(np.ascontiguousarray(w["X"], dtype=np.float32)) for _ in range(2)]
    Yh = [PinnedArray((n_nodes, w["F"]), np.float32) for _ in range(2)]
    e2e_steps = max(4, min(args.steps, 50))

    def e2e_run(n):
        tickets = [None, None]
        for i in range(n):
            b = i & 1
            if tickets[b] is not None:
                net.host_wait(tickets[b])          # step i-2's result is on the host: its buffers may be reused
            tickets[b] = net.forward_host_async(goff_h.array, rp_h.array, ci_h.array, None, Xh[b].array, Yh[b].array)
        for t_ in tickets:
            if t_ is not None:
                net.host_wait(t_)

    e2e_run(4)
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h2d = int(Xh[0].array.nbytes + rp_h.array.nbytes + ci_h.array.nbytes)  # graph_off stays on the host (tile planning)
    d2h = int(Yh[0].array.nbytes)

    # ---- second series: the reference's shipped 5-layer stack 4-32-32-32-32-1, K=1, on the same graphs
    # (one fused launch per step; informational, not part of `value`)
    from multihop_offload_b200 import reference_stack
    rs = np.random.default_rng(5)
    specs5 = reference_stack(K=1)
    net5 = ChebNet(specs5, device=dev)
    net5.set_weights([((rs.standard_normal((sp_.K, sp_.f_in, sp_.f_out)) * 0.2).astype(np.float32),
                       np.zeros(sp_.f_out, np.float32)) for sp_ in specs5])
    X5 = torch.randn((n_nodes, 4), device=dev)
    Y5 = torch.empty((n_nodes, 1), dtype=torch.float32, device=dev)
    for _ in range(5):
        net5.forward(batches[0], X5, out=Y5)
    barrier()
    e5a, e5b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n5 = max(10, min(args.steps, 100))
    e5a.record()
    for i in range(n5):
        net5.forward(batches[i % R], X5, out=Y5)
    e5b.record()
    barrier()
    stack5_ms = e5a.elapsed_time(e5b) / n5
    # third series: forward (activations kept) + VJP to one flat gradient per graph + deterministic sum
    dYt = torch.randn((n_nodes, w["F"]), device=dev)
    def train_step():
        Yt, saved = net.forward(batches[0], Xs[0], save=True)
        net.backward(batches[0], Xs[0], Yt, saved, dYt)
    for _ in range(3):
        train_step()
    barrier()
    eta, etb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nt = max(5, min(args.steps, 30))
    eta.record()
    for _ in range(nt):
        train_step()
    etb.record()
    barrier()
    train_ms = eta.elapsed_time(etb) / nt

    t = torch.tensor([ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_s = float(t[0]), float(t[1])
    value = world * args.graphs * args.steps / (ms * 1e-3)
    e2e_val = world * args.graphs * e2e_steps / e2e_s

    if rank == 0:
        peaks, peak_src = None, "fallback"
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy burst)"
        except Exception:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
        per_launch_ms = ms / max(launches, 1)
        achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("dram_bytes_per_launch")
        except Exception:
            pass
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (fp32 storage and accumulation; both products run on tcgen05 with operands split into three bf16 parts, fp32-grade)", "data": "synthetic",
            "config": workload_config(args, w),
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "mho_cheb_forward_host_async + mho_host_wait, two steps in flight (page-locked host buffers from mho_host_alloc; every step uploads X + CSR and downloads Y; chunked upload/kernel/download pipeline)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "cheb_dense_kernel",
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "algorithmic_bytes_survey_formula": alg_bytes + 4 * int(w["rowptr"][-1]),
                         "tflops_algorithmic": algorithmic_flops(w) / (per_launch_ms * 1e-3) / 1e12},
        }
        out["series"] = {"reference_stack_4_32_32_32_32_1_K1": {
            "value": args.graphs / (stack5_ms * 1e-3), "unit": "graph forwards/s per GPU (5 fused layers, rank 0)", "ms_per_step": stack5_ms},
            "forward_backward_K5_32_32": {
            "value": args.graphs / (train_ms * 1e-3), "unit": "graph forward+VJP steps/s per GPU (per-graph gradients + their sum, rank 0)", "ms_per_step": train_ms}}
        if world == 1 and not args.no_cpu:
            v, cores, passes, dt = cpu_reference_rate(w, seconds=args.cpu_seconds, threads=1)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": "%d passes over the same %d-graph batch in %.1f s (oracle/cheb_oracle.c: fp64, "
                                             "one graph at a time like the reference's eager call, 1 thread)" % (passes, args.graphs, dt)}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--graphs", type=int, default=1024, help="graphs per GPU per step")
    ap.add_argument("--fixed-n", type=int, default=None, help="all graphs of this size (e.g. 100)")
    ap.add_argument("--tile-rows", type=int, default=128)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--streams", type=int, default=2, help="independent steps are issued round-robin on this many CUDA streams")
    ap.add_argument("--no-pack", action="store_true", help="keep the random graph order instead of tile-packing order")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 20:
            pass  # each step is a full pass (~0.1 s with 8 threads); the driver picks K
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
