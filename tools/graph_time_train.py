"""Device time of forward(save) + VJP (per-graph gradients + sum) for the benchmark layer, CUDA-graph replays.
usage: graph_time_train.py [steps=10]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from multihop_offload_b200 import ChebNet, GraphBatch, LayerSpec
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
w = bench.make_workload(1024)
dev = torch.device("cuda:0")
net = ChebNet([LayerSpec(w["K"], 32, 32)], device=dev)
rs = np.random.default_rng(5)
net.set_weights([((rs.standard_normal((w["K"], 32, 32)) * 0.1).astype(np.float32), np.zeros(32, np.float32))])
b = GraphBatch(w["graph_off"], w["rowptr"], w["colidx"], None, device=dev)
n = int(w["graph_off"][-1])
X = torch.randn(n, 32, device=dev); dY = torch.randn(n, 32, device=dev)

def timeit(fn, name):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(steps): fn()
    torch.cuda.synchronize()
    ts = []
    for rep in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / steps)
    eager = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps): fn()
        e1.record(); torch.cuda.synchronize()
        eager.append(e0.elapsed_time(e1) * 1e3 / steps)
    print("%-28s graph %.1f us  eager %.1f us  -> %.2f M graph-steps/s" % (name, np.median(ts[2:]), np.median(eager), 1024 / np.median(ts[2:])))

Y, saved = net.forward(b, X, save=True)
timeit(lambda: net.forward(b, X, save=True), "forward(save)")
timeit(lambda: net.backward(b, X, Y, saved, dY), "backward")
timeit(lambda: net.backward(b, X, Y, saved, dY, need_sum=False), "backward (no sum)")
def both():
    Yt, sv = net.forward(b, X, save=True); net.backward(b, X, Yt, sv, dY)
timeit(both, "forward+backward")
