"""``from .GAT3D.GATMultiHead3D import GATMultiHead3D`` (convolutional_gat/model.py:3)."""
from cgat.layers import GATMultiHead3D

__all__ = ["GATMultiHead3D"]
