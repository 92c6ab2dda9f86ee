"""A few launches of the VQ search at one size (for ncu captures).  usage: run_vq_case.py N K [umma]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "msmc-tts_b200"))
import torch  # noqa: E402
from msmctts._b200 import functional as Fn  # noqa: E402

n, K = int(sys.argv[1]), int(sys.argv[2])
Fn.VQ_UMMA = len(sys.argv) > 3 and sys.argv[3] == "umma"
dev = torch.device("cuda:0")
z = torch.randn(n, 256, device=dev)
e = torch.randn(4, 64, K, device=dev)
with torch.no_grad():
    for _ in range(5):
        Fn.vq_quantize(z, e, 4, 64)
torch.cuda.synchronize()
