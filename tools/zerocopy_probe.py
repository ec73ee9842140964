"""Does the forward kernel run well with X / Y (and the CSR slice) left in page-locked HOST memory (zero-copy over PCIe)?"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from multihop_offload_b200 import ChebNet, GraphBatch, LayerSpec
from multihop_offload_b200._lib import pinned_like
w = bench.make_workload(1024)
dev = torch.device("cuda:0")
net = ChebNet([LayerSpec(5, 32, 32, 2, 0.2)], device=dev)
b = GraphBatch(w["graph_off"], w["rowptr"], w["colidx"], None, device=dev)
n = int(w["graph_off"][-1])
Xh = pinned_like(np.random.default_rng(0).standard_normal((n, 32)).astype(np.float32))
Yh = pinned_like(np.zeros((n, 32), np.float32))
Xd = torch.from_numpy(Xh.array).to(dev)
Yd = net.forward(b, Xd); torch.cuda.synchronize()
lib, st = net.ctx.lib, net._stream()
def run(xp, yp):
    rc = lib.mho_cheb_forward(net.ctx.handle, b.struct_ref(False), net.layer_structs(), 1, C.c_void_p(xp), C.c_void_p(yp), None, st)
    assert rc == 0
    torch.cuda.synchronize()
for name, xp, yp in (("device X, device Y", Xd.data_ptr(), Yd.data_ptr()), ("host X, device Y", Xh.array.ctypes.data, Yd.data_ptr()),
                     ("device X, host Y", Xd.data_ptr(), Yh.array.ctypes.data), ("host X, host Y", Xh.array.ctypes.data, Yh.array.ctypes.data)):
    for _ in range(3): run(xp, yp)
    t = time.perf_counter()
    for _ in range(20): run(xp, yp)
    print("%-20s %.1f us/call" % (name, (time.perf_counter() - t) / 20 * 1e6))
print("max |Y_host - Y_dev| = %.3g" % float(np.abs(Yh.array - Yd.cpu().numpy()).max()))
