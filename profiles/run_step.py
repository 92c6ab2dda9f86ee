"""One eager MSMC-VQ-GAN train step (bench.py's workload) inside a cudaProfilerStart/Stop range, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python profiles/run_step.py
Two untimed steps run first (optimizer state, weight-norm caches, kernel attributes)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
cfg = bench.load_cfg()
trainer = bench.build_gpu_trainer(cfg, dev, False, 0, 1, use_graph=False)
batch = bench.synth_batch(bench.B_PER_GPU, 1000, device=dev)
win = [(100, 100 + bench.WIN_FRAMES)] * bench.B_PER_GPU
for i in range(2):
    trainer.train_step(batch, iteration=10 + i, frame_windows=win)
torch.cuda.synchronize()
torch.cuda.profiler.start()
trainer.train_step(batch, iteration=12, frame_windows=win)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
