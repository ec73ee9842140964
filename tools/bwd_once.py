"""A few eager forward(save)+VJP steps of the benchmark layer (for ncu launch lists)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from multihop_offload_b200 import ChebNet, GraphBatch, LayerSpec
w = bench.make_workload(1024)
dev = torch.device("cuda:0")
net = ChebNet([LayerSpec(w["K"], 32, 32)], device=dev)
rs = np.random.default_rng(5)
net.set_weights([((rs.standard_normal((w["K"], 32, 32)) * 0.1).astype(np.float32), np.zeros(32, np.float32))])
b = GraphBatch(w["graph_off"], w["rowptr"], w["colidx"], None, device=dev)
n = int(w["graph_off"][-1])
X = torch.randn(n, 32, device=dev); dY = torch.randn(n, 32, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    Y, saved = net.forward(b, X, save=True)
    net.backward(b, X, Y, saved, dY)
torch.cuda.synchronize()
