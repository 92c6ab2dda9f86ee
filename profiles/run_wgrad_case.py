"""A few launches of the tensor-core weight-gradient kernel on one shape (for ncu captures / timing)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "msmc-tts_b200"))
import torch  # noqa: E402
from msmctts._b200 import functional as Fn  # noqa: E402
dev = torch.device("cuda:0")
CASES = {"mrf64": (16, 6000, 64, 64, 11, 5), "mrf32": (16, 12000, 32, 32, 11, 5), "ffn2": (16, 240, 1024, 256, 3, 1)}
B, L, Ci, Co, K, pad = CASES[sys.argv[1] if len(sys.argv) > 1 else "mrf64"]
x = torch.randn(B, 1, L, Ci, device=dev)
gy = torch.randn(B, 1, L, Co, device=dev)
gw = torch.empty(1, K, Ci, Co, device=dev)
for _ in range(4):
    Fn._launch_wgrad(x, gy, gw, (K * Ci * Co, Ci * Co, Co, 1), None, 1, K, 1, 1, 1, 1, 0, pad, False)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    Fn._launch_wgrad(x, gy, gw, (K * Ci * Co, Ci * Co, Co, 1), None, 1, K, 1, 1, 1, 1, 0, pad, False)
e1.record()
torch.cuda.synchronize()
print("wgrad %s: %.3f ms, %.1f TFLOP/s" % (sys.argv[1] if len(sys.argv) > 1 else "mrf64", e0.elapsed_time(e1) / 10,
                                           2.0 * B * L * K * Ci * Co / (e0.elapsed_time(e1) / 10) / 1e9))
