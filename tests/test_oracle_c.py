"""CPU: the plain-C oracle agrees with the numpy oracle (both are checkers)."""
import os

import numpy as np

import c_oracle
import chebnet_oracle as O
from helpers import oracle_batch_forward


def test_c_oracle_matches_numpy_oracle(golden_dir):
    z = np.load(os.path.join(golden_dir, "layer_K5_F32.npz"))
    Y = c_oracle.stack_forward(z["graph_off"], z["rowptr"], z["colidx"], None, [(z["W"], z["b"])], [O.ACT_LEAKY], 0.2,
                               z["X"], n_threads=2)
    np.testing.assert_allclose(Y, z["Y"], rtol=1e-11, atol=1e-9)


def test_c_oracle_stack_and_values():
    rng = np.random.default_rng(0)
    sizes = [5, 3, 33, 64]
    mats = O.make_batch(sizes, seed0=3, operator="cheb-lap")
    import scipy.sparse as sp
    mats[1] = sp.csr_matrix((1, 1))  # single node, no edges
    g, rp, ci, va = O.concat_batch(mats)
    X = rng.normal(size=(g[-1], 4))
    ws = O.glorot_weights([4, 16, 1], 3, rng)
    ws = [(W, b + 0.1) for W, b in ws]
    acts = [O.ACT_LEAKY, O.ACT_RELU]
    Y = c_oracle.stack_forward(g, rp, ci, va, ws, acts, 0.2, X)
    np.testing.assert_allclose(Y, oracle_batch_forward(mats, X, ws, acts, 0.2), rtol=1e-10, atol=1e-12)
