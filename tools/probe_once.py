import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from multihop_offload_b200 import ChebNet, GraphBatch, LayerSpec
w = bench.make_workload(int(sys.argv[1]) if len(sys.argv) > 1 else 1024)
dev = torch.device("cuda:0")
b = GraphBatch(w["graph_off"], w["rowptr"], w["colidx"], None, device=dev)
net = ChebNet([LayerSpec(5, 32, 32, 2, 0.2)], device=dev)
X = torch.randn(int(w["graph_off"][-1]), 32, device=dev)
for _ in range(3): net.forward(b, X)
torch.cuda.synchronize()
