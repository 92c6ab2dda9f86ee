"""Single-launch latency of the tensor-core VQ search at 3840 rows (K = 256) against the number of clusters
(MSMC_VQ_MAX_CLUSTERS): 30 clusters x 1 tile ... 8 clusters x 4 tiles."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "msmc-tts_b200"))
import torch  # noqa: E402
from msmctts._b200 import functional as Fn  # noqa: E402

dev = torch.device("cuda:0")
Fn.VQ_UMMA = True
heads, dim = 4, 64
for K in (256, 64):
    for n in (3840, 1920):
        for cap in (33, 20, 15, 10, 8, 5):
            os.environ["MSMC_VQ_MAX_CLUSTERS"] = str(cap)
            res = []
            for trial in range(3):
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
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                res.append(e0.elapsed_time(e1) * 1e3 / 20)
            print("K=%3d n=%4d clusters<=%2d  us/launch %s" % (K, n, cap, " ".join("%5.1f" % t for t in res)), flush=True)
