"""The model registry of ``convolutional_gat/utils.py`` (reference :13-22); metrics/plots there are out of scope."""
from .GAT3D.GATMultistream import Model as GatModel
from .unet_model import UnetModel

model_classes = {
    "unet": UnetModel,
    "temporal": GatModel,
    "spatial": GatModel,
    "multi_stream": GatModel,
}


def get_number_parameters(model):
    return sum(p.numel() for p in model.parameters() if p.requires_grad)
