"""Run a few launches of one conv shape through the C-ABI (for ncu captures).  usage: run_conv_case.py CASE [fwd|bwd]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "msmc-tts_b200"))
import torch  # noqa: E402
from msmctts._b200 import functional as Fn  # noqa: E402

CASES = {
    # name: (B, L, Ci, Co, K, dilation)
    "ffn2": (16, 240, 1024, 256, 3, 1),
    "ffn1": (16, 240, 256, 1024, 3, 1),
    "mrf32": (16, 12000, 32, 32, 11, 1),
    "mrf64": (16, 6000, 64, 64, 11, 3),
    "mrf128": (16, 1200, 128, 128, 11, 5),
    "mrf256": (16, 240, 256, 256, 11, 1),
}
name = sys.argv[1]
mode = sys.argv[2] if len(sys.argv) > 2 else "fwd"
B, L, Ci, Co, K, d = CASES[name]
dev = torch.device("cuda:0")
x = torch.randn(B, 1, L, Ci, device=dev, requires_grad=(mode == "bwd"))
v = torch.randn(Co, Ci, K, device=dev, requires_grad=True)
bias = torch.randn(Co, device=dev, requires_grad=True)
pad = (K * d - d) // 2
for it in range(4):
    w = Fn.prep_conv_weight(v)
    y = Fn.conv_cl(x, w, bias, kernel=(1, K), dilation=(1, d), padding=(0, pad), pre_slope=0.1)
    if mode == "bwd":
        y.sum().backward()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    w = Fn.prep_conv_weight(v)
    Fn.conv_cl(x, w, bias, kernel=(1, K), dilation=(1, d), padding=(0, pad), pre_slope=0.1)
    e0.record()
    for _ in range(10):
        Fn.conv_cl(x, w, bias, kernel=(1, K), dilation=(1, d), padding=(0, pad), pre_slope=0.1)
    e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
fl = 2.0 * B * L * K * Ci * Co
print("%s fwd %.3f ms  %.1f TFLOP/s  (%.1f GB/s algorithmic)" % (name, ms, fl / ms / 1e9,
      4.0 * (B * L * (Ci + Co) + K * Ci * Co) / ms / 1e6))
