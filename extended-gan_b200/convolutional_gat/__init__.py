"""Drop-in mirror of the reference's ``convolutional_gat`` package for the hot path only.

Put ``extended-gan_b200/`` on ``sys.path`` and the reference's own import lines keep working:
``from convolutional_gat.model import TemporalModel``, ``from convolutional_gat.GAT3D.GATMultiHead3D import
GATMultiHead3D`` (the sub-module the reference is missing), ``from convolutional_gat.baseline_model import
BaselineModel2D``.  Data loaders, plotting, preprocessing and the CLI are out of scope.
"""
