"""Which Python lines launch the step's non-msmc (aten) CUDA kernels: torch.profiler with stacks over one eager
train step, kernels grouped by the innermost repo frame of the op that launched them."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
cfg = bench.load_cfg()
trainer = bench.build_gpu_trainer(cfg, dev, False, 0, 1, use_graph=False)
batch = bench.synth_batch(bench.B_PER_GPU, 1000, device=dev)
win = [(100, 100 + bench.WIN_FRAMES)] * bench.B_PER_GPU
for i in range(2):
    trainer.train_step(batch, iteration=10 + i, frame_windows=win)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=True) as prof:
    trainer.train_step(batch, iteration=12, frame_windows=win)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CPU or not ev.name.startswith("aten::"):
        continue
    ktime = sum(k.duration for k in ev.kernels)
    if not ev.kernels:
        continue
    site = "?"
    for fr in ev.stack or []:
        if "/msmc-tts_b200/" in fr or "/bench.py" in fr:
            site = fr.split("/msmc-tts_b200/")[-1].strip()
            break
    if site == "?" and ev.stack:
        site = "autograd engine / " + ev.stack[0].strip()[-60:]
    key = (ev.name, site[:110])
    agg[key][0] += len(ev.kernels)
    agg[key][1] += ktime
tot = sum(v[1] for v in agg.values())
print("aten kernels: %d launches, %.2f ms" % (sum(v[0] for v in agg.values()), tot / 1e3))
for (name, site), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%5d %8.1f us  %-28s %s" % (n, t, name, site))
