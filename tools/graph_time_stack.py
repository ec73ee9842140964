"""Device time of the shipped 5-layer K=1 stack (4-32-32-32-32-1) over the benchmark's 1024 graphs, CUDA-graph replays.
usage: graph_time_stack.py [streams=1] [steps=20]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from multihop_offload_b200 import ChebNet, GraphBatch, reference_stack
n_str = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
w = bench.make_workload(1024)
dev = torch.device("cuda:0")
specs = reference_stack(K=1)
rs = np.random.default_rng(5)
net = ChebNet(specs, device=dev)
net.set_weights([((rs.standard_normal((s.K, s.f_in, s.f_out)) * 0.2).astype(np.float32), np.zeros(s.f_out, np.float32)) for s in specs])
R = 8
batches = [GraphBatch(w["graph_off"], w["rowptr"], w["colidx"], None, device=dev) for _ in range(R)]
n = int(w["graph_off"][-1])
Xs = [torch.randn(n, 4, device=dev) for _ in range(R)]
Ys = [torch.empty(n, 1, device=dev) for _ in range(R)]
for i in range(5):
    net.forward(batches[i % R], Xs[i % R], out=Ys[i % R])
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
side = [torch.cuda.Stream() for _ in range(n_str)]
with torch.cuda.stream(s):
    with torch.cuda.graph(g, stream=s):
        if n_str == 1:
            for i in range(steps): net.forward(batches[i % R], Xs[i % R], out=Ys[i % R])
        else:
            for ss in side: ss.wait_stream(s)
            for i in range(steps):
                with torch.cuda.stream(side[i % n_str]): net.forward(batches[i % R], Xs[i % R], out=Ys[i % R])
            for ss in side: s.wait_stream(ss)
torch.cuda.synchronize()
ts = []
for rep in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3 / steps)
ts = np.array(ts[2:])
print("streams %d: us/step median %.2f min %.2f -> %.1f M graph forwards/s (rows %d, tiles %d)" % (n_str, np.median(ts), ts.min(), 1024 / np.median(ts), n, batches[0].n_row_tiles))
