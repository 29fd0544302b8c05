"""``UnetModel`` of ``convolutional_gat/unet_model.py`` (reference :8-29): one shared SmaAt-UNet(4 -> 4) per vertex.

The reference applies the SAME UNet to each vertex in a Python loop (:25-26); in train mode its BatchNorm layers
therefore see per-vertex batch statistics and make V sequential running-stat updates per step (folding V into the batch
naively would change the numbers -- SURVEY.md 3.4).  Here the V passes run as ONE pass over a ``[V*B, T, H, W]`` batch:
every op of the net is per-sample except BatchNorm, and the BatchNorm kernels take the vertices as ``sets`` -- each set
normalised with its own statistics, the running statistics updated once per set in vertex order
(``cgat_bn_stats_sets``).  Same outputs, gradients and buffers as the loop, 1/V of the launches, V times the work per
launch.  ``batched = False`` keeps the literal loop.
"""
import torch as t
from torch import nn

from cgat.norm_act import BatchNormAct2d

from .GAT3D.smaat_unet.SmaAt_UNet import SmaAt_UNet


class UnetModel(nn.Module):
    def __init__(self, *, image_width: int, image_height: int, n_vertices: int, attention_type: str,
                 mapping_type: str = "conv"):
        super().__init__()
        self.mapping_type = mapping_type
        self.unet = SmaAt_UNet(n_channels=4, n_classes=4)
        self.batched = True

    def _set_sets(self, sets: int):
        for m in self.unet.modules():
            if isinstance(m, BatchNormAct2d):
                m.sets = sets

    def forward(self, x):
        B, H, W, T, V = x.shape
        x = x.permute(4, 0, 3, 1, 2)  # [V, B, T, H, W]        (:24)
        if self.batched and x.is_cuda:
            self._set_sets(V)
            try:
                out = self.unet(x.reshape(V * B, T, H, W))  # vertex-major: vertex v = images [v*B, (v+1)*B)
            finally:
                self._set_sets(1)
            return out.reshape(V, B, *out.shape[1:]).permute(1, 3, 4, 2, 0)  # [B, H, W, T, V]   (:27-28)
        acc = [self.unet(x[i]) for i in range(V)]  # (:25-26)
        return t.stack(acc).permute(1, 3, 4, 2, 0)
