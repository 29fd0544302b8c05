"""Drop-ins for ``convolutional_gat/baseline_model.py`` (reference :13-270) on the B200 kernels."""
import torch as t
from torch import nn

from cgat.layers import GATMultiHead, GATMultiHead2D, GraphAttentionLayer, GraphAttentionLayer2D

__all__ = ["GraphAttentionLayer", "GATMultiHead", "GraphAttentionLayer2D", "GATMultiHead2D", "BaselineModel2D",
           "BaselineModel"]


class BaselineModel2D(nn.Module):
    """reference :200-233 -- two single-head 2-D GAT layers then tanh."""

    def __init__(self, *, image_width: int, image_height: int, n_vertices: int, time_steps: int = 4,
                 mapping_type="linear"):
        super().__init__()
        self.mapping_type = mapping_type
        self.hidden_layer = GATMultiHead2D(nfeat=time_steps, nhid=time_steps, n_vertices=n_vertices, alpha=0.2, nheads=1)
        self.output_layer = GATMultiHead2D(nfeat=time_steps, nhid=time_steps, n_vertices=n_vertices, alpha=0.2, nheads=1)

    def forward(self, x):
        B, H, W, T, V = x.shape
        x = x.reshape(B, H * W, T, V)  # :229
        x = self.output_layer(self.hidden_layer(x))  # :230-231
        return t.tanh(x.view(B, H, W, T, V))  # :232-233


class BaselineModel(nn.Module):
    """reference :236-270 -- two single-head 1-D GAT layers on ``F = H*W*T`` features per vertex, then tanh.

    The reference ``.view``s the ``[B, V, F]`` result back to ``[B, H, W, T, V]`` WITHOUT permuting (:269); that
    axis scramble is part of its behaviour and is kept.
    """

    def __init__(self, *, image_width: int, image_height: int, n_vertices: int, time_steps: int = 4,
                 mapping_type="linear"):
        super().__init__()
        self.mapping_type = mapping_type
        n_features = time_steps * image_height * image_width
        self.hidden_layer = GATMultiHead(nfeat=n_features, nhid=n_features, n_vertices=n_vertices, alpha=0.2, nheads=1)
        self.output_layer = GATMultiHead(nfeat=n_features, nhid=n_features, n_vertices=n_vertices, alpha=0.2, nheads=1)

    def forward(self, x):
        B, H, W, T, V = x.shape
        x = x.reshape(B, H * W * T, V).permute(0, 2, 1)  # :266
        x = self.output_layer(self.hidden_layer(x))  # :267-268
        return t.tanh(x.reshape(B, H, W, T, V))  # :269-270 (raw view of [B, V, F])
