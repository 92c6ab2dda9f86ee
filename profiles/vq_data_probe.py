"""Is the two-valued single-tile VQ launch time a property of the DATA or of the buffers?  One CUDA graph (fixed
buffers), the inputs re-randomised in place between timings."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "msmc-tts_b200"))
import torch  # noqa: E402
from msmctts._b200 import functional as Fn  # noqa: E402

dev = torch.device("cuda:0")
Fn.VQ_UMMA = True
heads, dim, K = 4, 64, 256
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
                q, d, i = Fn.vq_quantize(z, embed, heads, dim)
    out = []
    for trial in range(12):
        if trial % 3 == 1:
            z.normal_()
        if trial % 3 == 2:
            embed.normal_()
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        out.append("%s%.1f" % ("" if trial % 3 == 0 else ("z:" if trial % 3 == 1 else "e:"), e0.elapsed_time(e1) * 1e3 / 20))
    print("n=%4d same graph, z: = new rows, e: = new codebook ->" % n, " ".join(out), flush=True)
