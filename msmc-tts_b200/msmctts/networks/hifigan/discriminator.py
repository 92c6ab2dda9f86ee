"""UnivNet discriminator (multi-resolution STFT + multi-period) on the sm_100a kernels.
Same classes / kwargs / state_dict keys / return structure as reference hifigan/discriminator.py:15-190.
Internally everything is channels-last; the returned scores and feature maps are permuted VIEWS in the
reference's (B, C, H, W) layout, so the trainer's losses see identical tensors."""
import torch
import torch.nn.functional as F
from torch import nn

from msmctts._b200 import functional as Fn
from msmctts._b200 import layers as Ly
from msmctts.utils.audio import TorchSTFT
from .common import get_padding

LRELU_SLOPE = 0.2


class _Layer(nn.Module):
    """names the conv like the reference's nn.Sequential slot ('1' in [pad, conv], '2' in [lrelu, pad, conv])"""

    def __init__(self, slot, conv):
        super().__init__()
        self.slot = slot
        self.add_module(slot, conv)

    def forward(self, x, **kw):
        return getattr(self, self.slot)(x, **kw)


class DiscriminatorR(nn.Module):
    def __init__(self, in_channels, hidden_channels=512):
        super().__init__()
        h = hidden_channels
        chans = [in_channels, h // 32, h // 16, h // 8, h // 4, h // 2, h, 1]
        strides = [1, 2, 1, 2, 1, 2, 1]
        self.discriminator = nn.ModuleList([
            _Layer("1" if i == 0 else "2",
                   Ly.WNConv2d(chans[i], chans[i + 1], (3, 3), stride=(strides[i],) * 2, padding=(1, 1),
                               reflect=True, swap_hw=True)) for i in range(7)])

    def forward_cl(self, x):
        """x (B, frames, F, C).  The reference's in-place LeakyReLU(0.2, True) mutates the tensors it has already
        stored in `hiddens` (discriminator.py:70-76), so its feature maps are post-activation: the activation is
        fused into the producing conv's epilogue here."""
        hiddens = []
        n = len(self.discriminator)
        for i, layer in enumerate(self.discriminator):
            x = layer(x, post=("lrelu", LRELU_SLOPE) if i < n - 1 else "none")
            hiddens.append(x)
        return x, hiddens[:-1]

    def forward(self, x):
        score, hid = self.forward_cl(x.permute(0, 3, 2, 1))
        return score.permute(0, 3, 2, 1), [h.permute(0, 3, 2, 1) for h in hid]


class MultiResolutionDiscriminator(nn.Module):
    def __init__(self, hop_lengths=[15, 30, 50, 120, 240, 480], hidden_channels=[128, 128, 256, 256, 512, 512],
                 domain="double", mel_scale=True, sample_rate=24000):
        super().__init__()
        self.stfts = nn.ModuleList([
            TorchSTFT(fft_size=x * 4, hop_size=x, win_size=x * 4, normalized=True, domain=domain,
                      mel_scale=mel_scale, sample_rate=sample_rate) for x in hop_lengths])
        self.domain = domain
        self.discriminators = nn.ModuleList([
            DiscriminatorR(2 if domain == "double" else 1, c) for _, c in zip(hop_lengths, hidden_channels)])

    def forward(self, x):
        outs = Fn.run_branches(self.branches(x))
        return [o[0] for o in outs], [o[1] for o in outs]

    def branches(self, x):
        """one closure per resolution -> (score, feature maps) in the reference's (B, C, H, W) layout"""
        wav = x.reshape(x.shape[0], -1)

        def one(stft, disc):
            score, feat = disc.forward_cl(stft.transform_cl(wav))
            return score.permute(0, 3, 2, 1), [f.permute(0, 3, 2, 1) for f in feat]
        return [(lambda s=s, d=d: one(s, d)) for s, d in zip(self.stfts, self.discriminators)]


class DiscriminatorP(nn.Module):
    def __init__(self, period, ch=32, max_ch=1024, kernel_size=5, stride=3, use_spectral_norm=False):
        super().__init__()
        if use_spectral_norm:
            raise NotImplementedError("spectral_norm is unused by the in-tree configs")
        self.period = period
        c1, c2, c3, c4 = ch, ch * 4, min(max_ch, ch * 16), min(max_ch, ch * 32)
        pad = (get_padding(kernel_size, 1), 0)
        self.convs = nn.ModuleList([
            Ly.WNConv2d(1, c1, (kernel_size, 1), (stride, 1), padding=pad),
            Ly.WNConv2d(c1, c2, (kernel_size, 1), (stride, 1), padding=pad),
            Ly.WNConv2d(c2, c3, (kernel_size, 1), (stride, 1), padding=pad),
            Ly.WNConv2d(c3, c4, (kernel_size, 1), (stride, 1), padding=pad),
            Ly.WNConv2d(c4, c4, (5, 1), 1, padding=(2, 0))])
        self.conv_post = Ly.WNConv2d(c4, 1, (3, 1), 1, padding=(1, 0))

    def forward(self, x):
        """x (B, 1, T) -> (flattened score, [fmaps (B, C, H, period)])"""
        b, c, t = x.shape
        if t % self.period != 0:
            x = F.pad(x, (0, self.period - (t % self.period)), "reflect")
            t = x.shape[-1]
        h = x.reshape(b, t // self.period, self.period, 1)          # channels-last (B, H, W=period, 1)
        fmap = []
        for i, conv in enumerate(self.convs):
            h = conv(h, pre_slope=LRELU_SLOPE if i > 0 else None)    # pre-activation maps, lrelu on operand load
            fmap.append(h.permute(0, 3, 1, 2))
        h = self.conv_post(h, pre_slope=LRELU_SLOPE)
        return torch.flatten(h, 1, -1), fmap


class MultiPeriodDiscriminator(nn.Module):
    def __init__(self, periods=[2, 3, 5, 7, 11], channels=32, max_channels=1024):
        super().__init__()
        self.discriminators = nn.ModuleList([DiscriminatorP(p, channels, max_channels) for p in periods])

    def forward(self, y):
        outs = Fn.run_branches(self.branches(y))
        return [o[0] for o in outs], [o[1] for o in outs]

    def branches(self, y):
        return [(lambda d=d: d(y)) for d in self.discriminators]


class Discriminator(nn.Module):
    def __init__(self, mrd_config, mpd_config):
        super().__init__()
        self.mrd = MultiResolutionDiscriminator(**mrd_config)
        self.mpd = MultiPeriodDiscriminator(**mpd_config)

    def forward(self, y):
        if y.dim() == 2:
            y = y.unsqueeze(1)
        # all 3 + 5 sub-discriminators are independent: one fork, eight branches, one join
        outs = Fn.run_branches(self.mrd.branches(y) + self.mpd.branches(y))
        return [o[0] for o in outs], [o[1] for o in outs]
