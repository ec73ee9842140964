"""multihop_offload_b200 - B200-native (sm_100a) batched ChebConv hot path of
zhongyuanzhao/multihop-offload behind a thin C-ABI (include/mho.h, libmho.so).

Host side mirrors the reference's operator interface (``gnn_offloading_agent.ACOAgent``);
PyTorch is used only for device memory, streams and torch.distributed plumbing.
There is no CPU fallback: importing works anywhere, running needs libmho.so and a B200.
"""
from ._lib import MhoError, lib_path, load_library  # noqa: F401
from .batch import GraphBatch, pack_order  # noqa: F401
from .chebnet import ChebNet, LayerSpec, reference_stack  # noqa: F401

__all__ = ["MhoError", "lib_path", "load_library", "GraphBatch", "pack_order", "ChebNet", "LayerSpec", "reference_stack"]
