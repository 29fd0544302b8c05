"""``UnetModel`` of ``convolutional_gat/unet_model.py`` (reference :8-29): one shared SmaAt-UNet(4 -> 4) per vertex.

The reference applies the SAME UNet to each vertex in a Python loop (:25-26); in train mode its BatchNorm layers
therefore see per-vertex batch statistics and make V sequential running-stat updates per step.  That order is kept
(folding V into the batch would change the numbers -- SURVEY.md 3.4).
"""
import torch as t
from torch import nn

from .GAT3D.smaat_unet.SmaAt_UNet import SmaAt_UNet


class UnetModel(nn.Module):
    def __init__(self, *, image_width: int, image_height: int, n_vertices: int, attention_type: str,
                 mapping_type: str = "conv"):
        super().__init__()
        self.mapping_type = mapping_type
        self.unet = SmaAt_UNet(n_channels=4, n_classes=4)

    def forward(self, x):
        x = x.permute(4, 0, 3, 1, 2)  # [V, B, T, H, W]        (:24)
        acc = [self.unet(x[i]) for i in range(x.shape[0])]  # (:25-26)
        return t.stack(acc).permute(1, 3, 4, 2, 0)  # [B, H, W, T, V]   (:27-28)
