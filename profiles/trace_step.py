"""Kernel timeline of the CUDA-graph-replayed train step (CUPTI through torch.profiler): every kernel with its
stream, start and duration -> gpurun_out/<tag>_timeline.csv.  The ncu launch list is serialised and cold-cache; this
is the real concurrent schedule, which is what decides the step time."""
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
cfg = bench.load_cfg()
trainer = bench.build_gpu_trainer(cfg, dev, False, 0, 1, use_graph=True)
batch = bench.synth_batch(bench.B_PER_GPU, 1000, device=dev)
win = [(100, 100 + bench.WIN_FRAMES)] * bench.B_PER_GPU
for i in range(8):
    trainer.train_step(batch, iteration=10 + i, frame_windows=win)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(2):
        trainer.train_step(batch, iteration=20 + i, frame_windows=win)
    torch.cuda.synchronize()
rows = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        rows.append((ev.time_range.start, ev.time_range.end - ev.time_range.start, getattr(ev, "device_index", 0),
                     ev.name[:160]))
rows.sort()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
# the stream id is only in the chrome trace: export it and pull (name, ts, dur, stream) from there
trace = os.path.join(ROOT, "gpurun_out", tag + "_trace.json")
prof.export_chrome_trace(trace)
import json  # noqa: E402
with open(trace) as f:
    tr = json.load(f)
out = []
for e in tr["traceEvents"]:
    if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and e.get("ph") == "X":
        out.append((e["ts"], e["dur"], e.get("args", {}).get("stream", -1), e.get("cat"), e["name"][:160]))
out.sort()
with open(os.path.join(ROOT, "gpurun_out", tag + "_timeline.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["ts_us", "dur_us", "stream", "cat", "name"])
    w.writerows(out)
os.remove(trace)
print("kernels", len(out))
