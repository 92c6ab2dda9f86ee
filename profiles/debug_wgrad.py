"""Debug helper: run the tensor-core weight gradient on structured inputs and print how the result relates to the
exact answer (used while bringing up the MN-major operand path)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "msmc-tts_b200"))
import torch  # noqa: E402
from msmctts._b200 import functional as Fn  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def run(tag, x, gy, K=1):
    B, _, L, Ci = x.shape
    Co = gy.shape[-1]
    gw = torch.full((1, K, Ci, Co), -7.0, device=dev)
    gb = torch.full((Co,), -7.0, device=dev)
    pad = (K - 1) // 2
    Fn._launch_wgrad(x, gy, gw, (K * Ci * Co, Ci * Co, Co, 1), gb, 1, K, 1, 1, 1, 1, 0, pad, False)
    torch.cuda.synchronize()
    xp = torch.nn.functional.pad(x[:, 0], (0, 0, pad, pad))
    ref = torch.stack([torch.einsum("blc,bld->cd", xp[:, k:k + L], gy[:, 0]) for k in range(K)])[None]
    err = (gw - ref).abs().max().item()
    print("%-28s max|ref| %.3e  max|out| %.3e  err %.3e  frac(out==0) %.3f  frac(out==-7) %.3f  bias_err %.3e" % (
        tag, ref.abs().max().item(), gw.abs().max().item(), err, (gw == 0).float().mean().item(),
        (gw == -7).float().mean().item(), (gb - gy.sum(dim=(0, 1, 2))).abs().max().item()))
    return gw, ref


for mode in ("3xtf32", "tf32"):
    Fn.CONV_MATH = mode
    print("mode", mode)
    B, L, Ci, Co = 2, 512, 32, 32
    run("ones", torch.ones(B, 1, L, Ci, device=dev), torch.ones(B, 1, L, Co, device=dev))
    x = torch.zeros(B, 1, L, Ci, device=dev); x[..., 3] = 1.0
    g = torch.zeros(B, 1, L, Co, device=dev); g[..., 5] = 1.0
    gw, ref = run("onehot c3 x n5", x, g)
    nz = (gw[0, 0] != 0).nonzero()
    print("   nonzero at", nz[:8].tolist(), "value", gw[0, 0][gw[0, 0] != 0][:4].tolist())
    x = torch.zeros(B, 1, L, Ci, device=dev); x[0, 0, 7, :] = torch.arange(Ci, device=dev).float() + 1
    g = torch.zeros(B, 1, L, Co, device=dev); g[0, 0, 7, :] = (torch.arange(Co, device=dev).float() + 1) * 100
    gw, ref = run("single position ramp", x, g)
    print("   out[0:3,0:3]", gw[0, 0, :3, :3].tolist(), " ref", ref[0, 0, :3, :3].tolist())
    run("random C32", torch.randn(B, 1, L, Ci, device=dev), torch.randn(B, 1, L, Co, device=dev))
    run("random C64->128 k3", torch.randn(B, 1, L, 64, device=dev), torch.randn(B, 1, L, 128, device=dev), K=3)
