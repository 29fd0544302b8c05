"""``from .GAT3D.smaat_unet.SmaAt_UNet import SmaAt_UNet`` (convolutional_gat/unet_model.py:4).

The upstream file is missing from the reference tree; this is the public SmaAt-UNet architecture
(depthwise-separable double convs with kernels_per_layer=2, CBAM with reduction_ratio=16, bilinear up-sampling)
whose parameter count, 4,032,548 at (n_channels=4, n_classes=4), is the number the reference recorded
(compare_models/results/results.json:18).  Every op runs in our CUDA kernels: the convolutions (depthwise 3x3, pointwise
1x1, CBAM 7x7, output 1x1) in the conv kernels, BatchNorm + ReLU / sigmoid fused (cgat.norm_act), max-pooling, bilinear
up-sampling + pad + concat in one pass, and CBAM's channel / spatial gates with their pooling and MLP (cgat.unet_ops;
SURVEY.md 8f rank 2).  PARITY UNPINNED beyond the parameter count; checked against oracle/spec.py SpecSmaAtUNet (same
module tree, same state_dict keys).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from cgat import unet_ops
from cgat.conv_layers import Conv2d
from cgat.norm_act import ACT_RELU, ACT_SIGMOID, BatchNormAct2d


class DepthwiseSeparableConv(nn.Module):
    def __init__(self, cin, cout, kernel_size, padding=0, kernels_per_layer=1):
        super().__init__()
        self.depthwise = Conv2d(cin, cin * kernels_per_layer, kernel_size, padding=padding, groups=cin)
        self.pointwise = Conv2d(cin * kernels_per_layer, cout, 1)

    def forward(self, x):
        return self.pointwise(self.depthwise(x))


class DoubleConvDS(nn.Module):
    def __init__(self, cin, cout, mid=None, kernels_per_layer=1):
        super().__init__()
        mid = mid or cout
        self.double_conv = nn.Sequential(
            # (BatchNorm2d + ReLU as one fused op at the BatchNorm's index; an Identity keeps the ReLU's index, so the
            # state_dict keys double_conv.{0,1,3,4}.* are those of the public SmaAt-UNet)
            DepthwiseSeparableConv(cin, mid, 3, padding=1, kernels_per_layer=kernels_per_layer),
            BatchNormAct2d(mid, act=ACT_RELU), nn.Identity(),
            DepthwiseSeparableConv(mid, cout, 3, padding=1, kernels_per_layer=kernels_per_layer),
            BatchNormAct2d(cout, act=ACT_RELU), nn.Identity(),
        )

    def forward(self, x):
        return self.double_conv(x)


class DownDS(nn.Module):
    def __init__(self, cin, cout, kernels_per_layer=1):
        super().__init__()
        self.maxpool_conv = nn.Sequential(unet_ops.MaxPool2d(), DoubleConvDS(cin, cout, kernels_per_layer=kernels_per_layer))

    def forward(self, x):
        return self.maxpool_conv(x)


class UpDS(nn.Module):
    def __init__(self, cin, cout, bilinear=True, kernels_per_layer=1):
        super().__init__()
        if not bilinear:
            raise NotImplementedError("the reference configuration is bilinear")
        self.conv = DoubleConvDS(cin, cout, cin // 2, kernels_per_layer=kernels_per_layer)

    def forward(self, x1, x2):
        # Upsample(scale_factor=2, bilinear, align_corners=True) -> F.pad to x2's size -> cat([x2, x1], dim=1): one kernel
        return self.conv(unet_ops.upsample_pad_concat(x1, x2))


class ChannelAttention(nn.Module):
    def __init__(self, c, reduction_ratio=16):
        super().__init__()
        self.MLP = nn.Sequential(nn.Flatten(), nn.Linear(c, c // reduction_ratio), nn.ReLU(),
                                 nn.Linear(c // reduction_ratio, c))

    def forward(self, x):
        # x * sigmoid(MLP(avg_pool(x)) + MLP(max_pool(x))): pooling, both MLP passes and the gate in three launches
        return unet_ops.channel_gate(x, self.MLP[1].weight, self.MLP[1].bias, self.MLP[3].weight, self.MLP[3].bias)


class SpatialAttention(nn.Module):
    def __init__(self, kernel_size=7):
        super().__init__()
        self.conv = Conv2d(2, 1, kernel_size, padding=(kernel_size - 1) // 2, bias=False)
        self.bn = BatchNormAct2d(1, act=ACT_SIGMOID)  # BatchNorm2d(1) + the sigmoid of the gate

    def forward(self, x):
        # x * sigmoid(bn(conv7x7(cat(mean_c(x), max_c(x)))))
        return unet_ops.pixel_gate(x, self.bn(self.conv(unet_ops.channel_pool(x))))


class CBAM(nn.Module):
    def __init__(self, c, reduction_ratio=16, kernel_size=7):
        super().__init__()
        self.channel_att = ChannelAttention(c, reduction_ratio)
        self.spatial_att = SpatialAttention(kernel_size)

    def forward(self, x):
        return self.spatial_att(self.channel_att(x))


class SmaAt_UNet(nn.Module):
    def __init__(self, n_channels, n_classes, kernels_per_layer=2, bilinear=True, reduction_ratio=16):
        super().__init__()
        self.n_channels, self.n_classes = n_channels, n_classes
        k, r = kernels_per_layer, reduction_ratio
        self.inc = DoubleConvDS(n_channels, 64, kernels_per_layer=k)
        self.cbam1 = CBAM(64, r)
        self.down1 = DownDS(64, 128, k)
        self.cbam2 = CBAM(128, r)
        self.down2 = DownDS(128, 256, k)
        self.cbam3 = CBAM(256, r)
        self.down3 = DownDS(256, 512, k)
        self.cbam4 = CBAM(512, r)
        self.down4 = DownDS(512, 512, k)
        self.cbam5 = CBAM(512, r)
        self.up1 = UpDS(1024, 256, bilinear, k)
        self.up2 = UpDS(512, 128, bilinear, k)
        self.up3 = UpDS(256, 64, bilinear, k)
        self.up4 = UpDS(128, 64, bilinear, k)
        self.outc = Conv2d(64, n_classes, 1)

    def forward(self, x):
        x1 = self.inc(x); a1 = self.cbam1(x1)
        x2 = self.down1(x1); a2 = self.cbam2(x2)
        x3 = self.down2(x2); a3 = self.cbam3(x3)
        x4 = self.down3(x3); a4 = self.cbam4(x4)
        x5 = self.down4(x4); a5 = self.cbam5(x5)
        x = self.up1(a5, a4)
        x = self.up2(x, a3)
        x = self.up3(x, a2)
        x = self.up4(x, a1)
        return self.outc(x)
