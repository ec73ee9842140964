"""Debug probe for the dense-adjacency forward path: runs golden cases one by one and prints errors."""
import glob, os, sys
import numpy as np, scipy.sparse as sp, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from helpers import rel_err
from multihop_offload_b200 import ChebNet, GraphBatch, reference_stack
g = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
order = sys.argv[1:] or ["case0", "case5", "case1", "case5", "case0"]
for tag, key, K in (("BAT800", "lam", 1), ("K3", "lam_K3", 3)):
    w = np.load(os.path.join(g, "weights_%s.npz" % tag))
    ws = [(w["W%d" % i], w["b%d" % i]) for i in range(5)]
    net = ChebNet(reference_stack(K=K), device="cuda:0"); net.set_weights(ws)
    for c in order:
        z = np.load(os.path.join(g, c + ".npz"))
        n = z["X"].shape[0]
        A = sp.csr_matrix((z["vals"], z["colidx"], z["rowptr"]), shape=(n, n))
        b = GraphBatch.from_scipy([A], tile_rows=128, device="cuda:0")
        Y = net.forward(b, torch.from_numpy(z["X"].astype(np.float32)).cuda()).cpu().numpy()
        ref = z[key]
        bad = np.where(np.abs(Y - ref).ravel() > 1e-4 * np.abs(ref).max())[0]
        print(tag, c, n, "err %.3g" % rel_err(Y, ref), "n_bad", len(bad), "first bad rows", bad[:8], "Y", Y.ravel()[bad[:4]], "ref", ref.ravel()[bad[:4]])
