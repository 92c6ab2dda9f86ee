"""Time stride-1 conv shapes of the train step with the N tile forced to 32 / 64 / 128 (MSMC_FORCE_BN), each as a
CUDA graph of 20 launches so that host launch overhead is excluded.  Output feeds umma_pick_bn's policy."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "msmc-tts_b200"))
import torch  # noqa: E402
from msmctts._b200 import functional as Fn  # noqa: E402

dev = torch.device("cuda:0")
# name: (B, H, W, Ci, Co, KH, KW, dil, pre_slope)
CASES = {
    "ffn1_T240": (16, 1, 240, 256, 1024, 1, 3, 1, None),
    "ffn2_T240": (16, 1, 240, 1024, 256, 1, 3, 1, None),
    "ffn1_T60": (16, 1, 60, 256, 1024, 1, 3, 1, None),
    "ffn2_T60": (16, 1, 60, 1024, 256, 1, 3, 1, None),
    "mrf256_k3": (16, 1, 240, 256, 256, 1, 3, 1, 0.1),
    "mrf256_k11": (16, 1, 240, 256, 256, 1, 11, 1, 0.1),
    "mrf128_k3": (16, 1, 1200, 128, 128, 1, 3, 1, 0.1),
    "mrf128_k11": (16, 1, 1200, 128, 128, 1, 11, 5, 0.1),
    "mrf64_k3": (16, 1, 6000, 64, 64, 1, 3, 1, 0.1),
    "mrf64_k11": (16, 1, 6000, 64, 64, 1, 11, 3, 0.1),
    "mrf32_k7": (16, 1, 12000, 32, 32, 1, 7, 1, 0.1),
    "lin_256_384_M3840": (3840, 1, 1, 256, 384, 1, 1, 1, None),
    "lin_384_256_M3840": (3840, 1, 1, 384, 256, 1, 1, 1, None),
    "lin_256_256_M960": (960, 1, 1, 256, 256, 1, 1, 1, None),
    "mpd512_14x11_B32": (32, 14, 11, 512, 512, 5, 1, 1, None),
    "mpd512_75x2_B16": (16, 75, 2, 512, 512, 5, 1, 1, None),
    "prior_256_512_k5": (16, 1, 240, 256, 512, 1, 5, 1, None),
}
print("%-20s %8s %8s %8s   (us per launch; TF/s at best)" % ("case", "BN32", "BN64", "BN128"))
for name, (B, H, W, Ci, Co, KH, KW, d, slope) in CASES.items():
    x = torch.randn(B, H, W, Ci, device=dev)
    w = torch.randn(KH, KW, Ci, Co, device=dev) * 0.05
    bias = torch.randn(Co, device=dev)
    ph, pw = (KH * d - d) // 2, (KW * d - d) // 2
    res = []
    for bn in (32, 64, 128):
        if bn > 32 and Co <= bn // 2:
            res.append(float("nan"))
            continue
        os.environ["MSMC_FORCE_BN"] = str(bn)

        def run():
            return Fn.conv_cl(x, w, bias, kernel=(KH, KW), dilation=(d if KH > 1 else 1, d if KW > 1 else 1),
                              padding=(ph, pw), pre_slope=slope)
        with torch.no_grad():
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(20):
                    run()
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) * 1e3 / 100)
    os.environ.pop("MSMC_FORCE_BN", None)
    fl = 2.0 * B * H * W * KH * KW * Ci * Co
    best = min(r for r in res if r == r)
    print("%-20s %8.1f %8.1f %8.1f   %6.1f TF/s" % (name, res[0], res[1], res[2], fl / best / 1e6))
