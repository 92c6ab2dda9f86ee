"""MRF residual blocks on the sm_100a conv kernels (reference hifigan/common.py:21-79).  x is (B, L, C)."""
from torch import nn

from msmctts._b200 import layers as Ly

LRELU_SLOPE = 0.1


def get_padding(kernel_size, dilation=1):
    return int((kernel_size * dilation - dilation) / 2)


class ResBlock1(nn.Module):
    """3 x [lrelu -> dilated conv -> lrelu -> conv -> + x]; the leaky-ReLUs ride on the conv's operand load and the
    residual add on its epilogue, so each pair is exactly two kernel launches."""

    def __init__(self, channels, kernel_size=3, dilation=(1, 3, 5)):
        super().__init__()
        self.convs1 = nn.ModuleList([
            Ly.WNConv1d(channels, channels, kernel_size, dilation=d, padding=get_padding(kernel_size, d))
            for d in dilation])
        self.convs2 = nn.ModuleList([
            Ly.WNConv1d(channels, channels, kernel_size, dilation=1, padding=get_padding(kernel_size, 1))
            for _ in dilation])

    def forward(self, x):
        for c1, c2 in zip(self.convs1, self.convs2):
            xt = c1(x, pre_slope=LRELU_SLOPE)
            x = c2(xt, pre_slope=LRELU_SLOPE, residual=x)
        return x

    def remove_weight_norm(self):
        for layer in list(self.convs1) + list(self.convs2):
            layer.remove_weight_norm()


class ResBlock2(nn.Module):
    def __init__(self, channels, kernel_size=3, dilation=(1, 3)):
        super().__init__()
        self.convs = nn.ModuleList([
            Ly.WNConv1d(channels, channels, kernel_size, dilation=d, padding=get_padding(kernel_size, d))
            for d in dilation])

    def forward(self, x):
        for c in self.convs:
            x = c(x, pre_slope=LRELU_SLOPE, residual=x)
        return x

    def remove_weight_norm(self):
        for layer in self.convs:
            layer.remove_weight_norm()
