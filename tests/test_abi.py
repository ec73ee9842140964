"""CPU: the C-ABI library builds, loads and exports every symbol include/mho.h declares."""
import ctypes as C
import os
import re

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(built_lib):
    from multihop_offload_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "mho.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mho_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no prototypes parsed"
    lib = C.CDLL(built_lib)
    for name in sorted(declared):
        assert hasattr(lib, name), "libmho.so does not export %s" % name
    bound = {n for n, _, _ in _lib.PROTOTYPES}
    assert declared == bound, (declared - bound, bound - declared)
    assert _lib.load_library().mho_version() == 100


def test_plan_tiles_packs_consecutive_graphs(built_lib):
    from multihop_offload_b200 import GraphBatch
    sizes = [20, 30, 110, 100, 28, 50, 50, 3, 128, 129, 1]
    goff = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    rowptr = np.arange(goff[-1] + 1, dtype=np.int32) * 2
    colidx = np.zeros(rowptr[-1], dtype=np.int32)
    b = GraphBatch(goff, rowptr, colidx, tile_rows=128)
    assert b.tile_off[0] == 0 and b.tile_off[-1] == len(sizes)
    rows = [goff[b.tile_off[i + 1]] - goff[b.tile_off[i]] for i in range(b.n_tiles)]
    assert rows == [50, 110, 128, 103, 128, 129, 1]
    assert b.max_tile_rows == 129 and b.max_tile_nnz == 258


def test_plan_tiles_rejects_oversized_graph(built_lib):
    from multihop_offload_b200 import GraphBatch, MhoError
    goff = np.array([0, 600], dtype=np.int32)
    rowptr = np.zeros(601, dtype=np.int32)
    with pytest.raises(MhoError):
        GraphBatch(goff, rowptr, np.zeros(0, dtype=np.int32))


def test_empty_batch_plans(built_lib):
    from multihop_offload_b200 import GraphBatch
    b = GraphBatch(np.zeros(1, dtype=np.int32), np.zeros(1, dtype=np.int32), np.zeros(0, dtype=np.int32))
    assert b.n_graphs == 0 and b.n_tiles == 0


def test_no_cpu_fallback_without_gpu(built_lib):
    """On a box without a GPU the product path must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from multihop_offload_b200 import ChebNet, MhoError, reference_stack
    with pytest.raises((MhoError, RuntimeError, AssertionError)):
        ChebNet(reference_stack(), device="cuda:0")
