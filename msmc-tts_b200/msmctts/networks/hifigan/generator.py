"""HifiGAN generator on the sm_100a conv kernels (reference hifigan/generator.py:9-64).
forward(mel (B, C, L)) -> (B, 1, L*prod(upsample_rates)) keeps the reference signature; forward_cl works on
(B, L, C) directly and is what MSMCVQGAN calls (no NCL<->NLC transposes on the hot path)."""
from torch import nn

from msmctts._b200 import functional as Fn
from msmctts._b200 import layers as Ly
from .common import LRELU_SLOPE, ResBlock1


class Generator(nn.Module):
    def __init__(self, resblock_kernel_sizes, resblock_dilation_sizes, upsample_rates, upsample_initial_channel,
                 upsample_kernel_sizes, num_mels=80):
        super().__init__()
        self.num_kernels = len(resblock_kernel_sizes)
        self.num_upsamples = len(upsample_rates)
        c0 = upsample_initial_channel
        self.conv_pre = Ly.WNConv1d(num_mels, c0, 7, 1, padding=3)
        self.ups = nn.ModuleList([
            Ly.WNConvTranspose1d(c0 // (2 ** i), c0 // (2 ** (i + 1)), k, u, padding=(k - u) // 2)
            for i, (u, k) in enumerate(zip(upsample_rates, upsample_kernel_sizes))])
        self.resblocks = nn.ModuleList()
        for i in range(len(self.ups)):
            ch = c0 // (2 ** (i + 1))
            for k, d in zip(resblock_kernel_sizes, resblock_dilation_sizes):
                self.resblocks.append(ResBlock1(ch, k, d))
        self.conv_post = Ly.WNConv1d(ch, 1, 7, 1, padding=3)

    def forward_cl(self, x):
        x = self.conv_pre(x)
        for i in range(self.num_upsamples):
            x = self.ups[i](x, pre_slope=LRELU_SLOPE)
            blocks = self.resblocks[i * self.num_kernels:(i + 1) * self.num_kernels]
            rs = Fn.run_branches([(lambda b=b, x=x: b(x)) for b in blocks])      # independent MRF branches
            xs = rs[0]
            for r in rs[1:]:
                xs = xs + r
            x = xs / self.num_kernels
        # F.leaky_relu default slope 0.01 (generator.py:52), conv_post, tanh -- one launch
        return self.conv_post(x, pre_slope=0.01, post="tanh")

    def forward(self, mel):
        return self.forward_cl(mel.transpose(1, 2)).transpose(1, 2)

    def remove_weight_norm(self):
        """reference generator.py:57-64: fold g * v / ||v|| into plain weights (inference).  The GEMM-layout weights
        and their tensor-core operand images are then baked once, on the first no-grad forward (layers._prepped)."""
        print("Removing weight norm...")
        for layer in self.ups:
            layer.remove_weight_norm()
        for block in self.resblocks:
            block.remove_weight_norm()
        self.conv_pre.remove_weight_norm()
        self.conv_post.remove_weight_norm()
