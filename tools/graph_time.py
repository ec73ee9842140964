"""Device time of the benchmark step without host launch cost: `steps` launches captured in one CUDA graph, replayed.
usage: graph_time.py [graphs=1024] [steps=20] [fixed_n=0] [K=5]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from multihop_offload_b200 import ChebNet, GraphBatch, LayerSpec
graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
fixed_n = int(sys.argv[3]) if len(sys.argv) > 3 else 0
K = int(sys.argv[4]) if len(sys.argv) > 4 else 5
n_str = int(sys.argv[5]) if len(sys.argv) > 5 else 1
w = bench.make_workload(graphs, 0, fixed_n or None, K=K)
dev = torch.device("cuda:0")
net = ChebNet([LayerSpec(K, 32, 32, 2, 0.2)], device=dev)
net.set_weights([(w["W"], w["b"])])
alg = bench.algorithmic_bytes(w)
R = max(2, int(np.ceil(2.2 * bench.L2_BYTES / alg)))
batches = [GraphBatch(w["graph_off"], w["rowptr"], w["colidx"], None, device=dev) for _ in range(R)]
n = int(w["graph_off"][-1])
Xs = [torch.randn(n, 32, device=dev) for _ in range(R)]
Ys = [torch.empty(n, 32, device=dev) for _ in range(R)]
print("tiles", batches[0].n_tiles, "nodes", n, "R", R, "alg MB %.2f" % (alg / 1e6))
for i in range(5):
    net.forward(batches[i % R], Xs[i % R], out=Ys[i % R])
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
side = [torch.cuda.Stream() for _ in range(n_str)]
with torch.cuda.stream(s):
    with torch.cuda.graph(g, stream=s):
        if n_str == 1:
            for i in range(steps):
                net.forward(batches[i % R], Xs[i % R], out=Ys[i % R])
        else:
            for ss in side:
                ss.wait_stream(s)
            for i in range(steps):
                with torch.cuda.stream(side[i % n_str]):
                    net.forward(batches[i % R], Xs[i % R], out=Ys[i % R])
            for ss in side:
                s.wait_stream(ss)
torch.cuda.synchronize()
ts = []
for rep in range(12):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3 / steps)
ts = np.array(ts[2:])
peak = 6580.9
print("streams", n_str, "graph replay: us/step median %.2f min %.2f max %.2f  -> %.1f M graph-steps/s, %.1f %% of HBM roofline (alg bytes, 4 nnz)" % (
    np.median(ts), ts.min(), ts.max(), graphs / np.median(ts), 100 * alg / (np.median(ts) * 1e-6) / 1e9 / peak))
# eager loop for comparison
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for i in range(steps):
    net.forward(batches[i % R], Xs[i % R], out=Ys[i % R])
e1.record(); torch.cuda.synchronize()
print("eager loop: us/step %.2f" % (e0.elapsed_time(e1) * 1e3 / steps))
