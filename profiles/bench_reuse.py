"""A/B timing of the stride-1 tap-reuse convolution kernels on the train step's own shapes (CUDA graph of 10 launches
per measurement, so host-side call cost is excluded).  usage: bench_reuse.py [case ...]
Configurations: one-tile-per-CTA kernel (round 1) and the persistent kernel at NACC = auto / 1 / 2 / 4."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "msmc-tts_b200"))
import torch  # noqa: E402
from msmctts._b200 import functional as Fn  # noqa: E402

CASES = {
    # name: (B, L, Ci, Co, K, dilation, residual)
    "mrf32k3": (16, 12000, 32, 32, 3, 1, True),
    "mrf32k11": (16, 12000, 32, 32, 11, 5, True),
    "mrf64k3": (16, 6000, 64, 64, 3, 1, True),
    "mrf64k11": (16, 6000, 64, 64, 11, 5, True),
    "mrf128k3": (16, 1200, 128, 128, 3, 1, True),
    "mrf128k11": (16, 1200, 128, 128, 11, 5, True),
    "mrf256k11": (16, 240, 256, 256, 11, 5, True),
    "ffn1": (16, 240, 256, 1024, 3, 1, False),
    "ffn2": (16, 240, 1024, 256, 3, 1, False),
    "ffn2_60": (16, 60, 1024, 256, 3, 1, False),
}
CONFIGS = [("one-tile", {"MSMC_REUSE_PERSIST": "0", "MSMC_REUSE_KSPLIT": "1"}),
           ("one-tile-ks2", {"MSMC_REUSE_PERSIST": "0", "MSMC_REUSE_KSPLIT": "2"}),
           ("one-tile-ks4", {"MSMC_REUSE_PERSIST": "0", "MSMC_REUSE_KSPLIT": "4"}),
           ("persist-auto", {"MSMC_REUSE_PERSIST": "1"}),
           ("persist-nacc1", {"MSMC_REUSE_PERSIST": "1", "MSMC_PERSIST_NACC": "1"}),
           ("persist-nacc2", {"MSMC_REUSE_PERSIST": "1", "MSMC_PERSIST_NACC": "2"}),
           ("persist-nacc4", {"MSMC_REUSE_PERSIST": "1", "MSMC_PERSIST_NACC": "4"})]
extra = os.environ.get("BENCH_REUSE_EXTRA")          # e.g. "MSMC_PERSIST_NBS=2"
if os.environ.get("BENCH_REUSE_CONFIGS"):            # e.g. "persist-nacc1,one-tile"
    CONFIGS = [c for c in CONFIGS if c[0] in os.environ["BENCH_REUSE_CONFIGS"].split(",")]
names = sys.argv[1:] or list(CASES)
dev = torch.device("cuda:0")
for name in names:
    B, L, Ci, Co, K, d, use_res = CASES[name]
    x = torch.randn(B, 1, L, Ci, device=dev)
    w = (torch.randn(1, K, Ci, Co, device=dev) / (Ci * K) ** 0.5)
    bias = torch.randn(Co, device=dev)
    res = torch.randn(B, 1, L, Co, device=dev) if use_res else None
    pad = (K * d - d) // 2
    fl = 2.0 * B * L * K * Ci * Co
    byt = 4.0 * (B * L * (Ci + Co * (2 if use_res else 1)) + K * Ci * Co)
    row = []
    for cname, env in CONFIGS:
        for k in ("MSMC_REUSE_PERSIST", "MSMC_PERSIST_NACC", "MSMC_PERSIST_NBS", "MSMC_REUSE_KSPLIT"):
            os.environ.pop(k, None)
        os.environ.update(env)
        if extra:
            k, v = extra.split("=")
            os.environ[k] = v

        def run():
            return Fn.conv_cl(x, w, bias, res, kernel=(1, K), dilation=(1, d), padding=(0, pad),
                              pre_slope=0.1 if use_res else None)
        with torch.no_grad():
            try:
                for _ in range(3):
                    run()
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for _ in range(10):
                        run()
                g.replay()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 100.0
                row.append("%s %6.1f us %6.1f TF/s" % (cname, us, fl / us / 1e6))
            except Exception as ex:
                row.append("%s FAILED %s" % (cname, str(ex)[:40]))
    print("%-10s (%.2f GF, %.1f MB) | %s" % (name, fl / 1e9, byt / 1e6, " | ".join(row)), flush=True)
