"""Device time of forward(kept activations) / VJP (+ sum) of the shipped 5-layer K = 1 stack over the benchmark's 1024 graphs.
usage: mlp_bwd_time.py [once]   ("once": a single eager VJP, for the -DMHO_PROBE clock marks)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from multihop_offload_b200 import ChebNet, GraphBatch, reference_stack
w = bench.make_workload(1024); dev = torch.device("cuda:0")
specs = reference_stack(K=1); net = ChebNet(specs, device=dev)
b = GraphBatch(w["graph_off"], w["rowptr"], w["colidx"], None, device=dev)
n = int(w["graph_off"][-1]); X = torch.randn(n, 4, device=dev); dY = torch.randn(n, 1, device=dev)
Y, saved = net.forward(b, X, save=True)
if len(sys.argv) > 1:
    net.backward(b, X, Y, saved, dY); torch.cuda.synchronize(); sys.exit(0)
def t(fn, name, steps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(steps): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3 / steps)
    print(name, "%.1f us" % np.median(ts[2:]))
t(lambda: net.forward(b, X, save=True), "fwd(save)")
t(lambda: net.backward(b, X, Y, saved, dY), "bwd+sum")
t(lambda: net.backward(b, X, Y, saved, dY, need_sum=False), "bwd no sum")
