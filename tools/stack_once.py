"""A few eager forwards of the shipped 5-layer K=1 stack over the benchmark's graphs (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from multihop_offload_b200 import ChebNet, GraphBatch, reference_stack
w = bench.make_workload(1024)
dev = torch.device("cuda:0")
specs = reference_stack(K=1)
rs = np.random.default_rng(5)
net = ChebNet(specs, device=dev)
net.set_weights([((rs.standard_normal((s.K, s.f_in, s.f_out)) * 0.2).astype(np.float32), np.zeros(s.f_out, np.float32)) for s in specs])
b = GraphBatch(w["graph_off"], w["rowptr"], w["colidx"], None, device=dev)
n = int(w["graph_off"][-1])
X = torch.randn(n, 4, device=dev); Y = torch.empty(n, 1, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    net.forward(b, X, out=Y)
torch.cuda.synchronize()
