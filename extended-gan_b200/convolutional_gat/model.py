"""Model wrappers of ``convolutional_gat/model.py`` (reference :8-248) over the B200 ``GATMultiHead3D``.

Same class names, keyword-only constructor arguments, ``hidden_layer`` / ``output_layer`` attribute
names and ``mapping_type`` attribute.  The reference's quirk of registering an ``output_layer`` whose
forward is commented out (:44-47, :85-88) is kept: its parameters exist in ``state_dict()`` and in
``parameters()`` but receive no gradient -- ``torch.optim.Adam`` skips parameters whose ``grad`` is ``None`` (weight
decay included), and so does ``cgat.train_step.TrainStep``: they keep their initial values.
"""
import torch.nn as nn

from .GAT3D.GATMultiHead3D import GATMultiHead3D


def _layer(time_steps, nheads, type_, mapping_type, image_height, image_width, n_vertices):
    return GATMultiHead3D(nfeat=time_steps, nhid=time_steps, alpha=0.2, nheads=nheads, type_=type_,
                          mapping_type=mapping_type, image_height=image_height, image_width=image_width,
                          n_vertices=n_vertices)


class _Wrapper(nn.Module):
    _type = None
    _hidden_heads = 3
    _output_heads = 1  # None: no output layer
    _use_output = False

    def __init__(self, *, image_width: int, image_height: int, n_vertices: int, time_steps: int = 4,
                 mapping_type="linear"):
        super().__init__()
        self.mapping_type = mapping_type
        self.hidden_layer = _layer(time_steps, self._hidden_heads, self._type, mapping_type, image_height,
                                   image_width, n_vertices)
        if self._output_heads is not None:
            self.output_layer = _layer(time_steps, self._output_heads, self._type, mapping_type, image_height,
                                       image_width, n_vertices)

    def forward(self, x):
        x = self.hidden_layer(x)
        if self._use_output:
            x = self.output_layer(x)
        return x


class SpatialModel(_Wrapper):  # reference :8-47
    _type = "spatial"


class TemporalModel(_Wrapper):  # reference :50-88
    _type = "temporal"


class TemporalModel4h(_Wrapper):  # reference :91-117
    _type = "temporal"
    _hidden_heads = 4
    _output_heads = None


class TemporalModel2l(_Wrapper):  # reference :120-158
    _type = "temporal"
    _output_heads = 3
    _use_output = True


class MultiStreamModel(_Wrapper):  # reference :210-248
    _type = "multi_stream"
    _hidden_heads = 1
    _use_output = True


class ConvGAT(nn.Module):  # reference :161-166 (an empty stub there too)
    def forward(self, x):
        pass
