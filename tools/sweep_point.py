"""One point of the cfg-5 sweep for profiling: usage sweep_point.py n K B [launches]."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from multihop_offload_b200 import ChebNet, GraphBatch, LayerSpec
n, K, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
w = bench.make_workload(B, 0, n, K=K, unique=64)
dev = torch.device("cuda:0")
b = GraphBatch(w["graph_off"], w["rowptr"], w["colidx"], None, device=dev)
net = ChebNet([LayerSpec(K, 32, 32, 2, 0.2)], device=dev)
net.set_weights([(w["W"], w["b"])])
X = torch.from_numpy(w["X"]).to(dev)
for _ in range(reps):
    net.forward(b, X)
torch.cuda.synchronize()
print("done", n, K, B, "tiles", b.n_tiles)
