"""tcgen05 (TF32 tensor-core) convolution path vs a plain PyTorch fp32 CPU reference.
3xTF32 (default): same tolerance as the CUDA-core fp32 kernels (2e-5 of the tensor's max);
plain TF32: 3e-3 (10-bit mantissa operands, fp32 accumulate)."""
import zlib

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def close(a, b, tol, msg=""):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (msg, a.shape, b.shape)
    scale = float(b.abs().max()) + 1e-12
    err = float((a - b).abs().max())
    assert err <= tol * scale + 1e-7, "%s max|diff| %.3e, scale %.3e (tol %.1e)" % (msg, err, scale, tol)


CASES = [
    # name, B, H, W, Ci, Co, KH, KW, stride, dil, pad, reflect, pre_slope, post, residual
    ("1x1 Cs32 Cd32", 2, 1, 300, 32, 32, 1, 1, (1, 1), (1, 1), (0, 0), False, None, "none", False),
    ("k3 Cs64 Cd64 M=517", 1, 1, 517, 64, 64, 1, 3, (1, 1), (1, 1), (0, 1), False, None, "relu", False),
    ("mrf k11 d5 Cs64", 2, 1, 400, 64, 64, 1, 11, (1, 1), (1, 5), (0, 25), False, 0.1, "none", True),
    ("mrf k7 d3 Cs32 Cd32", 3, 1, 1000, 32, 32, 1, 7, (1, 1), (1, 3), (0, 9), False, 0.1, "none", True),
    ("conv_post Cd1", 2, 1, 900, 32, 1, 1, 7, (1, 1), (1, 1), (0, 3), False, 0.01, "tanh", False),
    ("Cd100 ragged N", 2, 1, 260, 96, 100, 1, 3, (1, 1), (1, 1), (0, 1), False, None, "none", False),
    ("Cd200 two N tiles", 2, 1, 300, 64, 200, 1, 3, (1, 1), (1, 1), (0, 1), False, None, "tanh", False),
    ("ffn-like Cs256 Cd512", 2, 1, 240, 256, 512, 1, 3, (1, 1), (1, 1), (0, 1), False, None, "relu", False),
    ("mpd (5,1) s(3,1) Cs64", 2, 90, 5, 64, 128, 5, 1, (3, 1), (1, 1), (2, 0), False, 0.2, "none", False),
    ("mpd (5,1) s1 Cs64 W7", 2, 40, 7, 64, 64, 5, 1, (1, 1), (1, 1), (2, 0), False, 0.2, "none", False),
    ("mpd (3,1) s1 Cs64 Cd1", 2, 30, 11, 64, 1, 3, 1, (1, 1), (1, 1), (1, 0), False, 0.2, "none", False),
    ("k3 d1 L=1000 Cs32 (3 tiles/batch + tail)", 3, 1, 1000, 32, 48, 1, 3, (1, 1), (1, 1), (0, 1), False, None, "none", False),
    ("mrd 3x3 reflect s2 Cs32", 2, 40, 30, 32, 64, 3, 3, (2, 2), (1, 1), (1, 1), True, None, "lrelu", False),
    ("mrd 3x3 reflect s1 Cs64", 2, 21, 18, 64, 32, 3, 3, (1, 1), (1, 1), (1, 1), True, None, "lrelu", False),
    # channel counts that are not multiples of 32 (the AM's d_model = 600, 1456-wide decoder input): the last
    # 32-channel chunk of the reduction is ragged and zero-filled in the operand tiles / weight images
    ("AM ffn1 Cs600 Cd1536 k3", 2, 1, 240, 600, 1536, 1, 3, (1, 1), (1, 1), (0, 1), False, None, "relu", False),
    ("AM ffn2 Cs1536 Cd600 k3", 2, 1, 240, 1536, 600, 1, 3, (1, 1), (1, 1), (0, 1), False, None, "none", False),
    ("AM downsampler Cs600 k9", 2, 1, 300, 600, 600, 1, 9, (1, 1), (1, 1), (0, 4), False, None, "none", False),
    ("AM linear 1456->600", 1, 1, 480, 1456, 600, 1, 1, (1, 1), (1, 1), (0, 0), False, None, "none", False),
    ("ragged Cs40 strided 2-D", 2, 30, 20, 40, 48, 3, 3, (2, 1), (1, 1), (1, 1), False, 0.2, "none", False),
    # fewer than 32 source channels (second layers of the MRD / MPD stacks): one ragged chunk
    ("mrd 3x3 reflect s(1,2) Cs16", 2, 51, 60, 16, 32, 3, 3, (1, 2), (1, 1), (1, 1), True, None, "lrelu", False),
    ("mpd (5,1) s(3,1) Cs16", 2, 200, 7, 16, 64, 5, 1, (3, 1), (1, 1), (2, 0), False, 0.2, "none", False),
    ("mrd 3x3 reflect s1 Cs8", 2, 60, 40, 8, 16, 3, 3, (1, 1), (1, 1), (1, 1), True, None, "lrelu", False),
    ("k3 Cs16 Cd16 1-D", 2, 1, 700, 16, 16, 1, 3, (1, 1), (1, 1), (0, 1), False, 0.1, "none", True),
]


def mean_close(a, b, tol, msg=""):
    """plain TF32 flips a few relu/lrelu masks near zero relative to the fp32 reference, which moves isolated
    gradient entries by O(1); compare the mean error instead of the max"""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    err = float((a - b).abs().mean()) / (float(b.abs().mean()) + 1e-12)
    assert err <= tol, "%s mean rel err %.3e (tol %.1e)" % (msg, err, tol)


def _run_case(case, tol, cmp=None):
    cmp = cmp or close
    from msmctts._b200 import functional as Fn
    (name, B, H, W, Ci, Co, KH, KW, stride, dil, pad, reflect, pre_slope, post, use_res) = case
    gen = torch.Generator().manual_seed(zlib.crc32(name.encode()))
    x = torch.randn(B, H, W, Ci, generator=gen)
    v = torch.randn(Co, Ci, KH, KW, generator=gen) * (1.0 / (Ci * KH * KW) ** 0.5)
    g = torch.rand(Co, 1, 1, 1, generator=gen) + 0.5
    bias = torch.randn(Co, generator=gen) * 0.1
    Ho = Fn.conv_out_size(H, KH, stride[0], dil[0], pad[0], False)
    Wo = Fn.conv_out_size(W, KW, stride[1], dil[1], pad[1], False)
    res = torch.randn(B, Ho, Wo, Co, generator=gen)
    wgt = torch.randn(B, Ho, Wo, Co, generator=gen)
    pslope = 0.2 if post == "lrelu" else 0.0
    xr, vr, gr, br, rr = (t.clone().requires_grad_(True) for t in (x, v, g, bias, res))
    w_ref = vr * (gr / vr.norm(2, dim=(1, 2, 3), keepdim=True))
    xin = xr.permute(0, 3, 1, 2)
    if pre_slope is not None:
        xin = F.leaky_relu(xin, pre_slope)
    p = pad
    if reflect:
        xin = F.pad(xin, (pad[1], pad[1], pad[0], pad[0]), mode="reflect")
        p = (0, 0)
    y_ref = F.conv2d(xin, w_ref, br, stride=stride, padding=p, dilation=dil)
    y_ref = {"relu": F.relu, "tanh": torch.tanh, "lrelu": lambda t: F.leaky_relu(t, pslope),
             "none": lambda t: t}[post](y_ref).permute(0, 2, 3, 1)
    if use_res:
        y_ref = y_ref + rr
    (y_ref * wgt).sum().backward()
    xc, vc, gc, bc, rc = (t.to(DEV).requires_grad_(True) for t in (x, v, g, bias, res))
    w = Fn.prep_conv_weight(vc, gc)
    from msmctts._b200 import lib as L
    L.profile_begin()
    y = Fn.conv_cl(xc, w, bc, rc if use_res else None, kernel=(KH, KW), stride=stride, dilation=dil, padding=pad,
                   reflect=reflect, pre_slope=pre_slope, post=(post, pslope))
    (y * wgt.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    names = [n for n, _, _, _ in L.profile_end()]
    names = [n.replace("msmc_conv_forward_umma_reuse", "msmc_conv_forward_umma") for n in names]
    assert "msmc_conv_forward_umma" in names, "the tensor-core kernel did not run: %s" % names
    assert "msmc_conv_wgrad_umma" in names, "the tensor-core weight gradient did not run: %s" % names
    close(y, y_ref, tol, "y")
    cmp(xc.grad, xr.grad, tol, "dx")
    cmp(vc.grad, vr.grad, max(tol, 1e-4), "dv")
    cmp(bc.grad, br.grad, max(tol, 1e-4), "dbias")
    return names


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_umma_3xtf32(case, monkeypatch):
    from msmctts._b200 import functional as Fn
    monkeypatch.setattr(Fn, "CONV_MATH", "3xtf32")
    monkeypatch.setattr(Fn, "UMMA_MIN_CS", 8)      # (the default policy keeps Cs < 32 on the CUDA-core kernels)
    # 2e-5 of the tensor max up to reductions of ~1.5k terms (the GAN step's longest is 3 x 1024).  The hi / lo split
    # TRUNCATES (hi = top 19 bits, the tensor core truncates lo to TF32 again), so every partial product carries a
    # relative error of ~2^-22 with the SIGN OF THE PRODUCT: it does not average out, the bound grows with the
    # reduction length (the AM's longest reductions are 9 x 600 and 3 x 1536 terms)
    k_red = max(case[4], case[5]) * case[6] * case[7]
    names = _run_case(case, 2e-5 * max(1.0, k_red / 1536.0))
    stride = case[8]
    if stride == (1, 1) and not case[11] and case[5] % 32 == 0:
        assert names.count("msmc_conv_forward_umma") == 2, "forward and stride-1 data gradient both on tensor cores"


@pytest.mark.parametrize("case", [CASES[2], CASES[7], CASES[9]], ids=[CASES[2][0], CASES[7][0], CASES[9][0]])
def test_conv_umma_plain_tf32(case, monkeypatch):
    from msmctts._b200 import functional as Fn
    monkeypatch.setattr(Fn, "CONV_MATH", "tf32")
    _run_case(case, 1e-2, cmp=mean_close)


def test_tap_reuse_kernel_is_selected_and_matches_plain_kernel(monkeypatch):
    """stride-1 1-D conv: the tap-reuse kernel (shifted descriptors with base offset) must give the same numbers as
    the per-tap staging kernel, which is already pinned against the CPU reference"""
    from msmctts._b200 import functional as Fn
    from msmctts._b200 import lib as L
    monkeypatch.setattr(Fn, "CONV_MATH", "3xtf32")
    gen = torch.Generator().manual_seed(11)
    for (Ln, Ci, Co, K, d) in ((700, 64, 64, 11, 5), (333, 32, 96, 7, 3), (260, 256, 32, 3, 1)):
        x = torch.randn(3, 1, Ln, Ci, generator=gen).to(DEV)
        w = (torch.randn(1, K, Ci, Co, generator=gen) / (Ci * K) ** 0.5).to(DEV)
        pad = (K * d - d) // 2
        outs = []
        for reuse in (True, False):
            monkeypatch.setattr(Fn, "USE_TAP_REUSE", reuse)
            L.profile_begin()
            y = Fn.conv_cl(x, w, None, None, kernel=(1, K), dilation=(1, d), padding=(0, pad), pre_slope=0.1)
            torch.cuda.synchronize()
            names = [n for n, _, _, _ in L.profile_end()]
            assert ("msmc_conv_forward_umma_reuse" in names) == reuse, names
            outs.append(y)
        close(outs[0], outs[1], 1e-5, "reuse vs plain L=%d" % Ln)


def test_linear_umma(monkeypatch):
    from msmctts._b200 import functional as Fn
    monkeypatch.setattr(Fn, "CONV_MATH", "3xtf32")
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(16, 60, 256, generator=gen)
    W = torch.randn(384, 256, generator=gen) / 16
    b = torch.randn(384, generator=gen)
    xr, Wr, br = (t.clone().requires_grad_(True) for t in (x, W, b))
    y_ref = F.linear(xr, Wr, br)
    wgt = torch.randn(y_ref.shape, generator=gen)
    (y_ref * wgt).sum().backward()
    xc, Wc, bc = (t.to(DEV).requires_grad_(True) for t in (x, W, b))
    y = Fn.linear_cl(xc, Wc, bc)
    (y * wgt.to(DEV)).sum().backward()
    close(y, y_ref, 2e-5, "y")
    close(xc.grad, xr.grad, 2e-5, "dx")
    close(Wc.grad, Wr.grad, 1e-4, "dW")
    close(bc.grad, br.grad, 1e-4, "db")


PERSIST_SHAPES = [
    # B, L, Ci, Co, K, dilation, residual : chosen so that the persistent kernel runs several rounds of work items
    # per CTA, partial last tiles, several channel tiles and several 32-channel chunks
    (4, 12000, 32, 32, 11, 5, True),      # 376 items at NACC=1 (2.5 rounds), reach 50
    (3, 6000, 64, 64, 7, 3, True),        # KC=2
    (5, 1200, 128, 128, 3, 1, False),     # KC=4, BN=128
    (16, 240, 256, 512, 3, 1, False),     # FFN-like: 4 channel tiles, L < 256
    (16, 60, 512, 256, 3, 1, False),      # short sequences (60 of 128 rows used)
    (2, 1000, 96, 200, 5, 2, False),      # ragged channels (Cd not a multiple of the tile)
]


@pytest.mark.parametrize("nacc", [0, 1, 2, 4], ids=["auto", "nacc1", "nacc2", "nacc4"])
@pytest.mark.parametrize("shape", PERSIST_SHAPES, ids=["B%d_L%d_C%d-%d_k%d_d%d" % s[:6] for s in PERSIST_SHAPES])
def test_persistent_reuse_kernel_vs_one_tile_kernel_and_cpu(shape, nacc, monkeypatch):
    """conv_reuse_persist_kernel (persistent grid, warp-specialised roles, NACC accumulators per weight pass,
    double-buffered TMEM) against (1) the round-1 one-tile-per-CTA tap-reuse kernel: identical MMAs in identical
    order per output element -> bit-identical results, and (2) torch fp32 on the CPU (2e-5 of the tensor max).
    Forward and the stride-1 data gradient (same kernel, reversed-tap weight image) at every forced NACC."""
    from msmctts._b200 import functional as Fn
    monkeypatch.setattr(Fn, "CONV_MATH", "3xtf32")
    B, Ln, Ci, Co, K, d, use_res = shape
    gen = torch.Generator().manual_seed(Ln + Ci + K)
    x = torch.randn(B, 1, Ln, Ci, generator=gen)
    w = torch.randn(1, K, Ci, Co, generator=gen) / (Ci * K) ** 0.5
    bias = torch.randn(Co, generator=gen) * 0.1
    res = torch.randn(B, 1, Ln, Co, generator=gen) if use_res else None
    wgt = torch.randn(B, 1, Ln, Co, generator=gen)
    pad = (K * d - d) // 2
    outs = []
    monkeypatch.setenv("MSMC_REUSE_KSPLIT", "1")     # (split-K clusters change the summation order)
    for persist in ("1", "0"):
        monkeypatch.setenv("MSMC_REUSE_PERSIST", persist)
        if nacc:
            monkeypatch.setenv("MSMC_PERSIST_NACC", str(nacc))
        xc = x.to(DEV).requires_grad_(True)
        y = Fn.conv_cl(xc, w.to(DEV), bias.to(DEV), res.to(DEV) if use_res else None, kernel=(1, K),
                       dilation=(1, d), padding=(0, pad), pre_slope=0.1)
        (y * wgt.to(DEV)).sum().backward()
        torch.cuda.synchronize()
        outs.append((y.detach(), xc.grad.detach()))
    assert torch.equal(outs[0][0], outs[1][0]), "forward: persistent vs one-tile kernel must be bit-identical"
    assert torch.equal(outs[0][1], outs[1][1]), "data gradient: persistent vs one-tile kernel must be bit-identical"
    xr = x.clone().requires_grad_(True)
    y_ref = F.conv1d(F.leaky_relu(xr[:, 0].transpose(1, 2), 0.1), w[0].permute(2, 1, 0).contiguous(), bias,
                     padding=pad, dilation=d).transpose(1, 2).unsqueeze(1)
    if use_res:
        y_ref = y_ref + res
    (y_ref * wgt).sum().backward()
    close(outs[0][0], y_ref, 2e-5, "y vs cpu")
    close(outs[0][1], xr.grad, 2e-5, "dx vs cpu")


@pytest.mark.parametrize("ks", [0, 2, 4], ids=["auto", "ks2", "ks4"])
@pytest.mark.parametrize("shape", [(16, 240, 1024, 256, 3, 1, False),    # FFN second conv: 64 tiles x 96 MMA steps
                                   (16, 60, 1024, 256, 3, 1, False),     # short sequences: 32 tiles
                                   (16, 240, 256, 256, 11, 5, True),     # MRF, 8 chunks x 11 taps
                                   (3, 200, 160, 72, 5, 1, False)],      # ragged channels, 5 chunks
                         ids=["ffn2", "ffn2_T60", "mrf256k11", "ragged"])
def test_split_k_cluster_reuse_kernel(shape, ks, monkeypatch):
    """split-K thread-block clusters of the tap-reuse kernel (CTA z reduces its slice of the 32-channel chunks, the
    accumulators of CTAs 1.. travel through distributed shared memory to CTA 0, which sums them in rank order and
    runs the epilogue): forward and stride-1 data gradient vs torch fp32 on the CPU (2e-5 of the tensor max) and vs
    the unsplit kernel (4e-5: two results that are each within 2e-5 of the reference; the summation order differs)."""
    from msmctts._b200 import functional as Fn
    monkeypatch.setattr(Fn, "CONV_MATH", "3xtf32")
    monkeypatch.setenv("MSMC_REUSE_PERSIST", "0")
    B, Ln, Ci, Co, K, d, use_res = shape
    gen = torch.Generator().manual_seed(Ln + Ci + K)
    x = torch.randn(B, 1, Ln, Ci, generator=gen)
    w = torch.randn(1, K, Ci, Co, generator=gen) / (Ci * K) ** 0.5
    bias = torch.randn(Co, generator=gen) * 0.1
    res = torch.randn(B, 1, Ln, Co, generator=gen) if use_res else None
    wgt = torch.randn(B, 1, Ln, Co, generator=gen)
    pad = (K * d - d) // 2
    outs = []
    for split in ([str(ks)] if ks else [None]) + ["1"]:
        if split is None:
            monkeypatch.delenv("MSMC_REUSE_KSPLIT", raising=False)
        else:
            monkeypatch.setenv("MSMC_REUSE_KSPLIT", split)
        xc = x.to(DEV).requires_grad_(True)
        y = Fn.conv_cl(xc, w.to(DEV), bias.to(DEV), res.to(DEV) if use_res else None, kernel=(1, K),
                       dilation=(1, d), padding=(0, pad), pre_slope=0.1)
        (y * wgt.to(DEV)).sum().backward()
        torch.cuda.synchronize()
        outs.append((y.detach(), xc.grad.detach()))
    xr = x.clone().requires_grad_(True)
    y_ref = F.conv1d(F.leaky_relu(xr[:, 0].transpose(1, 2), 0.1), w[0].permute(2, 1, 0).contiguous(), bias,
                     padding=pad, dilation=d).transpose(1, 2).unsqueeze(1)
    if use_res:
        y_ref = y_ref + res
    (y_ref * wgt).sum().backward()
    close(outs[0][0], y_ref, 2e-5, "y vs cpu")
    close(outs[0][1], xr.grad, 2e-5, "dx vs cpu")
    close(outs[0][0], outs[1][0], 4e-5, "y vs unsplit")      # (each is within 2e-5 of the fp32 reference)
    close(outs[0][1], outs[1][1], 4e-5, "dx vs unsplit")
