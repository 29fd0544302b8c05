"""Drop-ins for ``convolutional_gat/baseline_model.py`` (reference :13-270) on the B200 kernels."""
import torch as t
from torch import nn

from cgat.layers import GATMultiHead2D, GraphAttentionLayer2D

__all__ = ["GraphAttentionLayer2D", "GATMultiHead2D", "BaselineModel2D"]


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
