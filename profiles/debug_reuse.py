"""Bring-up helper: tap-reuse kernel vs plain kernel, both descriptor base-offset conventions, plus wgrad timing."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "msmc-tts_b200"))
import torch  # noqa: E402
from msmctts._b200 import functional as Fn  # noqa: E402
dev = torch.device("cuda:0")
torch.manual_seed(0)
mode = os.environ.get("MSMC_REUSE_BASEOFF", "0")
for (Ln, Ci, Co, K, d) in ((256, 32, 32, 2, 1), (256, 32, 32, 2, 8), (700, 64, 64, 11, 5), (260, 256, 32, 3, 1)):
    x = torch.randn(2, 1, Ln, Ci, device=dev)
    w = torch.randn(1, K, Ci, Co, device=dev) / (Ci * K) ** 0.5
    pad = (K * d - d) // 2
    outs = []
    for reuse in (True, False):
        Fn.USE_TAP_REUSE = reuse
        outs.append(Fn.conv_cl(x, w, None, None, kernel=(1, K), dilation=(1, d), padding=(0, pad)))
    torch.cuda.synchronize()
    err = (outs[0] - outs[1]).abs().max().item()
    print("baseoff_mode=%s L=%d Ci=%d K=%d d=%d: max|reuse-plain| = %.3e (scale %.2f)" % (mode, Ln, Ci, K, d, err, outs[1].abs().max().item()))
if mode == "0":
    # wgrad timing: 3xTF32 (12 MMAs/stage) vs TF32 (4 MMAs/stage) on an M-heavy shape
    for cm in ("3xtf32", "tf32", "fp32"):
        Fn.CONV_MATH = cm
        B, L, Ci, Co, K = 16, 6000, 64, 64, 11
        x = torch.randn(B, 1, L, Ci, device=dev); gy = torch.randn(B, 1, L, Co, device=dev)
        gw = torch.empty(1, K, Ci, Co, device=dev)
        for _ in range(3):
            Fn._launch_wgrad(x, gy, gw, (K * Ci * Co, Ci * Co, Co, 1), None, 1, K, 1, 1, 1, 1, 0, 5, False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            Fn._launch_wgrad(x, gy, gw, (K * Ci * Co, Ci * Co, Co, 1), None, 1, K, 1, 1, 1, 1, 0, 5, False)
        e1.record(); torch.cuda.synchronize()
        print("wgrad C64 k11 L6000 %s: %.3f ms" % (cm, e0.elapsed_time(e1) / 10))
