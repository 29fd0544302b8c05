"""``from .GAT3D.GATMultistream import Model`` (convolutional_gat/train.py:10, utils.py:9).

The upstream source is missing; the constructor signature is pinned by its one call site
(train.py:199-205): ``Model(image_width=, image_height=, n_vertices=, attention_type=, mapping_type=)``
and the ``.mapping_type`` attribute read at train.py:208.  Built here as the wrapper of
``convolutional_gat/model.py`` matching ``attention_type`` (PARITY UNPINNED: upstream's parameter count of
43,936 at 20x20/V=6/temporal/conv, compare_models/results/results.json:9, is not reproduced).
"""
import torch.nn as nn

from ..model import MultiStreamModel, SpatialModel, TemporalModel

_BY_TYPE = {"spatial": SpatialModel, "temporal": TemporalModel, "multi_stream": MultiStreamModel}


class Model(nn.Module):
    def __init__(self, *, image_width: int, image_height: int, n_vertices: int, attention_type: str,
                 mapping_type: str = "linear", time_steps: int = 4):
        super().__init__()
        if attention_type not in _BY_TYPE:
            raise ValueError(f"attention_type must be one of {sorted(_BY_TYPE)}, got {attention_type!r}")
        self.mapping_type = mapping_type
        self.attention_type = attention_type
        self.net = _BY_TYPE[attention_type](image_width=image_width, image_height=image_height,
                                            n_vertices=n_vertices, time_steps=time_steps, mapping_type=mapping_type)

    def forward(self, x):
        return self.net(x)
