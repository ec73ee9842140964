"""Timing of the VJP (mho_cheb_backward) next to the forward on the benchmark batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from multihop_offload_b200 import ChebNet, GraphBatch, LayerSpec, reference_stack
w = bench.make_workload(1024)
dev = torch.device("cuda:0")
b = GraphBatch(w["graph_off"], w["rowptr"], w["colidx"], None, device=dev)
n = int(w["graph_off"][-1])
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e6
for name, specs, fi in (("K5 32->32", [LayerSpec(5, 32, 32, 2, 0.2)], 32), ("shipped stack K1", reference_stack(K=1), 4), ("stack K3", reference_stack(K=3), 4)):
    net = ChebNet(specs, device=dev)
    X = torch.randn(n, fi, device=dev)
    Y, saved = net.forward(b, X, save=True, per_graph_tiles=True)
    dY = torch.randn_like(Y)
    tf = t(lambda: net.forward(b, X))
    tfs = t(lambda: net.forward(b, X, save=True))
    tb = t(lambda: net.backward(b, X, Y, saved, dY))
    print("%-18s forward %.1f us, forward(save) %.1f us, backward %.1f us  (1024 graphs)" % (name, tf, tfs, tb))
