"""DCGAN next-frame nets of ``dcgan/model.py`` (reference :19-179) on the B200 conv kernels.

Same class names, ``params`` dict (``nc``, ``ndf``), sub-module names and ``state_dict`` keys
(``layers.{k}.layers.0.{weight,bias}``, ``layers.{k}.layers.1.*`` for G/TD; ``conv{1..5}.weight``, ``bn{2..4}.*`` for
FD), so reference checkpoints load unchanged.  Convolutions run in the CUDA kernels (tcgen05 implicit GEMM where
the shape is served, the direct kernel otherwise); BatchNorm2d (train-mode statistics, running-statistics update),
Dropout2d and the activation of a block are ONE fused op on the same channels_last tensors
(cgat.norm_act: csrc/norm_act_kernels.cu).  ``weights_init`` is kept as the no-op it is in the reference (it looks for
lower-case "conv"/"bn" in class names, :7-16).
"""
import torch as t
import torch.nn as nn
import torch.nn.functional as F

from cgat.conv_layers import Conv2d
from cgat.norm_act import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, ActDropout2d, BatchNormAct2d

_ACT_CODE = {}


def weights_init(w):
    classname = w.__class__.__name__
    if classname.find("conv") != -1:
        nn.init.normal_(w.weight.data, 0.0, 0.02)
    elif classname.find("bn") != -1:
        nn.init.normal_(w.weight.data, 1.0, 0.02)
        nn.init.constant_(w.bias.data, 0)


class ConvBlock(nn.Module):  # reference :19-52
    """conv -> BatchNorm2d -> Dropout2d -> activation.  ``layers.0`` is the conv and ``layers.1`` the BatchNorm, as in the
    reference's ``nn.Sequential`` (state_dict keys ``layers.0.*``, ``layers.1.*``); dropout and activation have no state
    and are fused into ``layers.1`` (or, without BatchNorm, into the stateless ``post``)."""

    def __init__(self, chin, chout, kernel_size, *, bias=True, stride=1, padding=0, dropout=0.01, act=F.relu,
                 batchnorm=True):
        super().__init__()
        code = _ACT_CODE.get(act)
        if code is None:
            raise NotImplementedError("ConvBlock activations: F.relu, leaky_relu(0.2), torch.sigmoid (dcgan/model.py)")
        layers = [Conv2d(chin, chout, kernel_size=kernel_size, stride=stride, padding=padding, bias=bias)]
        self.post = None
        if batchnorm and not (code == ACT_SIGMOID and dropout > 0):
            layers.append(BatchNormAct2d(chout, act=code, slope=0.2, dropout=dropout))
        else:
            if batchnorm:  # (sigmoid behind dropout: normalise first, then the two-pass dropout + sigmoid)
                layers.append(BatchNormAct2d(chout, act=ACT_NONE))
            self.post = ActDropout2d(act=code, slope=0.2, dropout=dropout)
        self.act = act
        self.layers = nn.Sequential(*layers)

    def forward(self, x):
        x = self.layers(x)
        return x if self.post is None else self.post(x)


class Generator(nn.Module):  # reference :55-76
    def __init__(self, params):
        super().__init__()
        self.params = params
        nc = params["nc"]
        self.layers = nn.Sequential(
            ConvBlock(nc, nc * 8, kernel_size=4, padding="same"),
            ConvBlock(nc * 8, nc * 4, 4, padding="same"),
            ConvBlock(nc * 4, nc * 2, 4, padding="same"),
            ConvBlock(nc * 2, nc, 4, padding="same"),
            ConvBlock(nc, nc, 4, padding="same", act=t.sigmoid, batchnorm=False),
        )

    def forward(self, x):
        return self.layers(x)


def _lrelu(x):
    return F.leaky_relu(x, 0.2, True)


_ACT_CODE.update({F.relu: ACT_RELU, _lrelu: ACT_LRELU, t.sigmoid: ACT_SIGMOID, None: ACT_NONE})


class TemporalDiscriminator(nn.Module):  # reference :79-142
    def __init__(self, params):
        super().__init__()
        nc, ndf = params["nc"], params["ndf"]
        self.layers = nn.Sequential(
            ConvBlock(2 * nc, ndf, kernel_size=4, stride=2, bias=False, batchnorm=False, padding=1, act=_lrelu),
            ConvBlock(ndf, 2 * ndf, kernel_size=4, stride=2, padding=1, bias=False, act=_lrelu),
            ConvBlock(2 * ndf, 4 * ndf, kernel_size=4, stride=2, padding=1, bias=False, act=_lrelu),
            ConvBlock(4 * ndf, 8 * ndf, kernel_size=4, stride=2, padding=1, bias=False, act=_lrelu),
            ConvBlock(8 * ndf, 1, kernel_size=4, stride=4, padding=0, bias=False, batchnorm=False, act=t.sigmoid),
        )

    def forward(self, x):
        return self.layers(x).squeeze()


class FrameDiscriminator(nn.Module):  # reference :145-179
    def __init__(self, params):
        super().__init__()
        nc, ndf = params["nc"], params["ndf"]
        self.conv1 = Conv2d(nc, ndf, 4, 2, 1, bias=False)
        self.conv2 = Conv2d(ndf, ndf * 2, 4, 2, 1, bias=False)
        self.bn2 = BatchNormAct2d(ndf * 2, act=ACT_LRELU, slope=0.2)  # LeakyReLU(0.2) of :172-174 fused in
        self.conv3 = Conv2d(ndf * 2, ndf * 4, 4, 2, 1, bias=False)
        self.bn3 = BatchNormAct2d(ndf * 4, act=ACT_LRELU, slope=0.2)
        self.conv4 = Conv2d(ndf * 4, ndf * 8, 4, 2, 1, bias=False)
        self.bn4 = BatchNormAct2d(ndf * 8, act=ACT_LRELU, slope=0.2)
        self.conv5 = Conv2d(ndf * 8, 1, 4, 1, 0, bias=False)
        self.act1 = ActDropout2d(act=ACT_LRELU, slope=0.2)
        self.act5 = ActDropout2d(act=ACT_SIGMOID)

    def forward(self, x):
        x = self.act1(self.conv1(x))
        x = self.bn2(self.conv2(x))
        x = self.bn3(self.conv3(x))
        x = self.bn4(self.conv4(x))
        return self.act5(self.conv5(x)).squeeze()
