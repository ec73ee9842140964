"""Kernel time vs Chebyshev order K (same batch): separates per-tile fixed cost from per-step cost."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from multihop_offload_b200 import ChebNet, GraphBatch, LayerSpec

w = bench.make_workload(int(sys.argv[1]) if len(sys.argv) > 1 else 1024)
dev = torch.device("cuda:0")
n = int(w["graph_off"][-1])
R = 16
batches = [GraphBatch(w["graph_off"], w["rowptr"], w["colidx"], None, device=dev) for _ in range(R)]
Xs = [torch.randn(n, 32, device=dev) for _ in range(R)]
Ys = [torch.empty(n, 32, device=dev) for _ in range(R)]
for K in (1, 2, 3, 5, 8):
    net = ChebNet([LayerSpec(K, 32, 32, 2, 0.2)], device=dev)
    for i in range(10): net.forward(batches[i % R], Xs[i % R], out=Ys[i % R])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(200): net.forward(batches[i % R], Xs[i % R], out=Ys[i % R])
    e1.record(); torch.cuda.synchronize()
    print("K=%d  %.1f us/step" % (K, e0.elapsed_time(e1) * 1000 / 200))
