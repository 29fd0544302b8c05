"""The ``GAT3D`` sub-module the reference imports but does not ship (convolutional_gat/model.py:3,
train.py:10, unet_model.py:4), provided here on the B200 kernels."""
