"""B200-native conv-GAT hot path: Python host side over libcgat_b200.so (no CPU fallback)."""
from . import _lib
from .functional import (AttnConfig, adam_step_, adjacency_norm, conv2d_nhwc, graph_attention, loss_and_grad,
                         IMPL_AUTO, IMPL_DIRECT, IMPL_TC)
from .layers import GATMultiHead2D, GATMultiHead3D, GraphAttentionLayer2D

__all__ = [
    "AttnConfig", "adam_step_", "adjacency_norm", "conv2d_nhwc", "graph_attention", "loss_and_grad",
    "GATMultiHead2D", "GATMultiHead3D", "GraphAttentionLayer2D", "IMPL_AUTO", "IMPL_DIRECT", "IMPL_TC",
]
