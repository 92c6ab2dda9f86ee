"""Why is the single-tile tensor-core VQ launch 19 us in some runs and 31 us in others?  Same process, fresh
allocations per trial, K = 256, 3840 and 960 rows: per-trial time and the buffers' addresses."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "msmc-tts_b200"))
import torch  # noqa: E402
from msmctts._b200 import functional as Fn  # noqa: E402

dev = torch.device("cuda:0")
Fn.VQ_UMMA = True
heads, dim, K = 4, 64, 256
keep = []
for trial in range(8):
    for n in (3840, 960):
        embed = torch.randn(heads, dim, K, device=dev)
        z = torch.randn(n, heads * dim, device=dev)
        with torch.no_grad():
            for _ in range(3):
                Fn.vq_quantize(z, embed, heads, dim)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(20):
                    Fn.vq_quantize(z, embed, heads, dim)
        g.replay()
        ts = []
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / 20)
        print("trial %d n=%4d  us/launch %s  z@%x (mod 2MB %7d)  embed@%x" %
              (trial, n, " ".join("%5.1f" % t for t in ts), z.data_ptr(), z.data_ptr() % (2 << 20), embed.data_ptr()),
              flush=True)
        keep.append((z, embed, g) if trial % 2 == 0 else None)   # vary what stays allocated
        if trial % 3 == 2:
            _junk = torch.empty(int(1.3e6) * (trial + 1), device=dev)   # shift later allocations
            keep.append(_junk)
