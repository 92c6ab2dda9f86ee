"""VQ search kernel alone (CUDA graph of 20 launches): default kernels vs the two-phase tensor-core kernel
(MSMC_VQ_UMMA=1), K = 64 / 256, at-config row counts and an N sweep."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "msmc-tts_b200"))
import torch  # noqa: E402
from msmctts._b200 import functional as Fn  # noqa: E402

dev = torch.device("cuda:0")
heads, dim = 4, 64


def run(n, K, umma, reps=20):
    Fn.VQ_UMMA = umma
    embed = torch.randn(heads, dim, K, device=dev)
    z = torch.randn(n, heads * dim, device=dev)
    with torch.no_grad():
        for _ in range(3):
            Fn.vq_quantize(z, embed, heads, dim)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                Fn.vq_quantize(z, embed, heads, dim)
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    byt = n * heads * dim * 4 * 3 + n * dim * 4 + n * heads * 8 + heads * dim * K * 4
    return us, byt / us / 1e3


for K in (64, 256):
    for n in (960, 3840, 1 << 14, 1 << 16, 1 << 18, 1 << 20, 1 << 22):
        a = run(n, K, False, 20 if n < (1 << 20) else 3)
        b = run(n, K, True, 20 if n < (1 << 20) else 3)
        print("K=%3d n=%8d | simt %8.1f us %7.1f GB/s | umma %8.1f us %7.1f GB/s" % (K, n, a[0], a[1], b[0], b[1]),
              flush=True)
