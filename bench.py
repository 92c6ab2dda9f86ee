#!/usr/bin/env python
"""bench.py -- mel-frames/s of one full MSMC-VQ-GAN GAN train step (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload ("config.workload"): the reference's VQGANTrainer.train_step after warm-up (trainers/msmctts_trainer.py:115-209)
on the CSMSC yaml architecture with 256 codewords/head as BASELINE.json words it: autoencoder forward/backward
(2-stage 4-head VQ + EMA, FFT encoders/decoder, HifiGAN generator on a 40-frame window), MelLoss, discriminator
(UnivNet MRD + MPD) step, generator adversarial + feature-matching step, grad clip, AdamW on both -- B=16 per GPU,
T=240, 80-dim mel, 24 kHz / hop 300 (the reference has no 22.05 kHz config), synthetic data,
random-init weights, fp32.

One JSON line on rank 0.  `value` = frames/s with inputs resident in HBM; `e2e` = the same through the public
trainer API with per-step pinned-host -> device copies of the batch and a device -> host read of the loss;
`roofline` = the dominant kernel family from a CUDA-event pass over instrumented steps; `cpu_baseline` = the CPU
oracle port of the same step on the box's host cores (rank 0, N=1 only).
`--impl reference` times that CPU port alone (the reference tree does not exist on the GPU box and its
discriminator / MelLoss do not run unmodified on torch >= 2, SURVEY section 0 B4/B5: kind = "port").
"""
import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "msmc-tts_b200"))

import torch  # noqa: E402

B_PER_GPU, T_FRAMES, N_MELS, HOP, WIN_FRAMES, K_CODEWORDS = 16, 240, 80, 300, 40, 256
CPU_SAMPLE_B = 16          # the CPU arm runs the SAME batch as the GPU arm (B=16); it is bounded by its step count


def load_cfg():
    with open(os.path.join(ROOT, "tests", "golden", "csmsc_config.json")) as f:
        cfg = json.load(f)
    cfg["autoencoder"]["quantizer_config"]["embedding_sizes"] = K_CODEWORDS
    return cfg


def synth_batch(B, seed, device="cpu", pin=False):
    g = torch.Generator().manual_seed(seed)
    mel = (1.5 * torch.randn(B, T_FRAMES, N_MELS, generator=g)).clamp_(-4, 4)
    wav = (0.3 * torch.randn(B, T_FRAMES * HOP, 1, generator=g)).clamp_(-1, 1)
    length = torch.full((B,), T_FRAMES, dtype=torch.int64)
    batch = {"mel": mel, "mel_length": length, "wav": wav}
    if pin:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    if device != "cpu":
        batch = {k: v.to(device) for k, v in batch.items()}
    return batch


AM_B, AM_L, AM_DUR = 32, 24, 10          # BASELINE.json configs[4]: B=32/GPU, 24 phonemes x 10 frames = T 240


def load_am_cfg():
    with open(os.path.join(ROOT, "tests", "golden", "csmsc_am_config.json")) as f:
        return json.load(f)


def synth_am_batch(B, seed, device="cpu", pin=False):
    """text (B, 24, 3) ints in [1,100) x [1,10) x {0,1}, duration 10 frames per phoneme, mel as above"""
    g = torch.Generator().manual_seed(seed)
    text = torch.stack([torch.randint(1, 100, (B, AM_L), generator=g), torch.randint(1, 10, (B, AM_L), generator=g),
                        torch.randint(0, 2, (B, AM_L), generator=g)], dim=-1)
    T = AM_L * AM_DUR
    batch = {"text": text, "text_length": torch.full((B,), AM_L, dtype=torch.int64),
             "dur": torch.full((B, AM_L), AM_DUR, dtype=torch.int64),
             "mel": (1.5 * torch.randn(B, T, N_MELS, generator=g)).clamp_(-4, 4),
             "mel_length": torch.full((B,), T, dtype=torch.int64)}
    if pin:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    if device != "cpu":
        batch = {k: v.to(device) for k, v in batch.items()}
    return batch


def build_am_trainer(cfg, am, device, distributed, rank, world):
    """PredictorTrainer on the CSMSC AM architecture with a frozen, random-init CSMSC autoencoder injected
    (the reference loads it from a checkpoint, trainers/msmctts_trainer.py:288-295; there is none here)"""
    from msmctts.networks.vqgantts import MSMCVQGAN
    from msmctts.tasks.msmc_tts import MSMCTTS
    from msmctts.trainers.msmctts_trainer import PredictorTrainer
    from msmctts.utils.config import Config, ConfigItem
    ycfg = {"id": "bench_am", "task": {"_name": "MSMCTTS", "_mode": "train_predictor",
                                       "predictor": dict(am["predictor"], _name="MultiStagePredictor")},
            "trainer": dict(am["trainer"]), "optimizer": am["optimizer"],
            "dataset": {"_name": "SyntheticMelDataset", "samplerate": 24000, "feature": ["text", "dur", "mel"],
                        "frameshift": [None, None, HOP]},
            "dataloader": {"batch_size": AM_B * world, "num_workers": 0}}
    config = Config(ycfg)
    torch.manual_seed(config.seed)
    task = MSMCTTS(config, mode="train")
    kwargs = config.trainer.to_dict()
    kwargs.pop("_name")
    trainer = PredictorTrainer(config, task, num_gpus=world, rank=rank, **kwargs)
    trainer.build_optimizer()
    c = copy.deepcopy(cfg["autoencoder"])
    torch.manual_seed(4321)
    ae = MSMCVQGAN(c["in_dim"], c["n_model_size"], ConfigItem(c["encoder_config"]), ConfigItem(c["quantizer_config"]),
                   ConfigItem(c["frame_decoder_config"]), ConfigItem(c["decoder_config"]), c["pred_mel"])
    for p in ae.parameters():
        p.requires_grad_(False)
    trainer.build_autoencoder(ae.to(device))
    task.train()
    return trainer


def am_workload_config(n):
    return {"workload": "MSMC-VQ-GAN-AM PredictorTrainer.train_step: frozen CSMSC autoencoder analysis + "
                        "MultiStagePredictor (FFT x 6 + 2 x FFT x 6, d_model 600, 114.8 M params) forward / backward, "
                        "mse + triplet-sum embedding loss, duration loss, clip 10, Adam (BASELINE.json configs[4])",
            "batch_per_gpu": AM_B, "global_batch": AM_B * n, "phonemes": AM_L, "mel_frames": AM_L * AM_DUR,
            "n_mels": N_MELS, "codewords_per_head": K_CODEWORDS, "parallelism": "dp%d" % n,
            "l2_policy": "per-step working set (114.8 M fp32 params + grads + Adam moments = 1.8 GB) exceeds the "
                         "126 MB L2; no explicit flush"}


def time_cpu_am_steps(cfg, am, steps, warmup):
    from oracle.train_step import OraclePredictorTrainer
    from msmctts.networks.acoustic_models import MultiStagePredictor
    threads, ncpu = pick_cpu_threads(cfg)
    torch.set_num_threads(threads)
    torch.manual_seed(1234)
    sd_ae, _ = init_state_dicts(cfg, 4321)
    sd_p = MultiStagePredictor(**copy.deepcopy(am["predictor"])).state_dict()
    tr = OraclePredictorTrainer(sd_p, sd_ae, am["predictor"], cfg["autoencoder"], am["trainer"],
                                am["optimizer"]["_default"])
    b = synth_am_batch(AM_B, 99)
    args = (b["text"], b["text_length"], b["dur"], b["mel"], b["mel_length"])
    for _ in range(warmup):
        tr.step(*args)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.step(*args)
    dt = (time.perf_counter() - t0) / max(1, steps)
    return AM_B * AM_L * AM_DUR / dt, dt, threads, ncpu


def run_am(args, rank, world, local_rank):
    """bench line of BASELINE.json configs[4] (`--config am`)"""
    import torch.distributed as dist
    from msmctts._b200 import lib as L
    cfg, am = load_cfg(), load_am_cfg()
    if args.impl == "reference":
        if rank != 0:
            return
        steps, warmup = min(args.steps, 10), min(args.warmup, 2)
        value, dt, threads, ncpu = time_cpu_am_steps(cfg, am, steps, warmup)
        sample = "PredictorTrainer step at B=%d, %d timed steps after %d warm-up, %d torch threads of %d CPUs" % (
            AM_B, steps, warmup, threads, ncpu)
        print(json.dumps({"impl": "reference", "metric": "mel-frames/sec MSMC-VQ-GAN-AM predictor train step",
                          "value": value, "unit": "mel-frames/s", "n_gpus": args.gpus, "steps": steps,
                          "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": am_workload_config(args.gpus),
                          "cpu_baseline": {"value": value, "unit": "mel-frames/s", "cores": threads, "kind": "port",
                                           "sample": sample},
                          "e2e": {"value": value, "unit": "mel-frames/s", "h2d_bytes_per_step": 0,
                                  "d2h_bytes_per_step": 0}}))
        return
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    L.load()
    distributed = world > 1
    if distributed and not dist.is_initialized():
        dist.init_process_group("nccl")
    trainer = build_am_trainer(cfg, am, device, distributed, rank, world)
    dev_batch = synth_am_batch(AM_B, 1000 + rank, device=device)
    host_batches = [synth_am_batch(AM_B, 2000 + rank * 16 + i, pin=True) for i in range(4)]

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
            torch.cuda.synchronize()

    def step_resident(i):
        return trainer.train_step(dev_batch, iteration=i)

    def step_e2e(i):
        hb = host_batches[i % len(host_batches)]
        log = trainer.train_step({k: v.to(device, non_blocking=True) for k, v in hb.items()}, iteration=i)
        return float(log["loss"]["total_loss"])

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if distributed:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    for i in range(max(3, args.warmup)):
        step_resident(i)
    torch.cuda.synchronize()
    l0 = L.launch_count
    step_resident(99)
    launches = L.launch_count - l0
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_step = timed(step_resident, args.steps)
    clocks = sampler.stop() if sampler else None
    for i in range(2):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)
    frames = AM_B * AM_L * AM_DUR * world
    h2d = sum(v.numel() * v.element_size() for v in host_batches[0].values())
    # roofline: the predictor's dense contractions (409 GMAC forward; x3 for fwd + dgrad + wgrad)
    pk = peaks()
    flops = 3 * 2 * 409e9 * world
    ach = flops / (ms_step * 1e-3) / 1e12
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt, threads, ncpu = time_cpu_am_steps(cfg, am, 2, 1)
        cpu = {"value": v, "unit": "mel-frames/s", "cores": threads, "kind": "port",
               "sample": "oracle port of the same PredictorTrainer step at the same batch (B=%d), 2 timed steps after "
                         "1 warm-up (%.1f s/step); %d torch threads of the host's %d logical CPUs" % (
                             AM_B, dt, threads, ncpu)}
    if rank == 0:
        print(json.dumps({
            "metric": "mel-frames/sec MSMC-VQ-GAN-AM predictor train step", "value": frames / (ms_step * 1e-3),
            "unit": "mel-frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": am_workload_config(world), "clocks": clocks,
            "e2e": {"value": frames / (ms_e2e * 1e-3), "unit": "mel-frames/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": launches * args.steps, "gpu_launches_per_step": launches, "cuda_graph": False,
            "roofline": {"kernel": "whole step (conv / linear contractions of the predictor)", "bound": "tensor",
                         "achieved": ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"],
                         "traffic": None, "peak_source": pk["source"],
                         "note": "algorithmic flops = 3 x 2 x 409 GMAC (SURVEY 8a a19) over the device-timed step"},
            "cpu_baseline": cpu}), flush=True)
    if distributed:
        torch.cuda.synchronize()
        sys.stdout.flush()
        threading.Timer(30.0, lambda: os._exit(0)).start()
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)


def family_traffic(entry):
    """DRAM bytes per launch of one C-ABI entry point's kernels, from the newest committed ncu capture of this same
    step (profiles/r*_family_traffic.json, written by profiles/family_traffic.py); None when there is none"""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_family_traffic.json")))
    if not files:
        return None, None
    try:
        with open(files[-1]) as f:
            d = json.load(f)["families"].get(entry)
        return (d["dram_bytes_per_launch"] if d else None), os.path.relpath(files[-1], ROOT)
    except Exception:
        return None, None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                p = json.load(f)
            return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p.get("bf16_tflops_sustained",
                                                                           p["bf16_tflops"])),
                    "source": "measured (MEASURED_PEAKS.json)"}
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled during the timed region"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
def build_cpu_trainer(cfg, seed=1234):
    from oracle.train_step import OracleTrainer
    sd_ae, sd_d = init_state_dicts(cfg, seed)
    return OracleTrainer(sd_ae, sd_d, cfg, cfg["trainer"], cfg["optimizer"]["_default"], use_dropout=True)


def init_state_dicts(cfg, seed):
    """random-init weights of the CSMSC architecture (constructed on CPU; no kernels involved)"""
    from msmctts.networks.hifigan import UnivNetDiscriminator
    from msmctts.networks.vqgantts import MSMCVQGAN
    from msmctts.utils.config import ConfigItem
    torch.manual_seed(seed)
    c = copy.deepcopy(cfg["autoencoder"])
    ae = MSMCVQGAN(c["in_dim"], c["n_model_size"], ConfigItem(c["encoder_config"]),
                   ConfigItem(c["quantizer_config"]), ConfigItem(c["frame_decoder_config"]),
                   ConfigItem(c["decoder_config"]), c["pred_mel"])
    d = UnivNetDiscriminator(ConfigItem(cfg["discriminator"]["mrd_config"]),
                             ConfigItem(cfg["discriminator"]["mpd_config"]))
    return ae.state_dict(), d.state_dict()


def pick_cpu_threads(cfg):
    """The oracle port is thousands of small torch ops: on a many-core host all-cores oversubscribes badly
    (measured: 128 threads were ~60x slower than 8).  Calibrate on an autoencoder forward and keep the best."""
    from oracle import ref_modules as O
    ncpu = os.cpu_count() or 1
    sd_ae, _ = init_state_dicts(cfg, 1234)
    batch = synth_batch(2, 7)
    best, best_t = 1, float("inf")
    for n in sorted({min(ncpu, c) for c in (8, 16, 32, 64, ncpu)}):
        torch.set_num_threads(n)
        with torch.no_grad():
            O.msmcvqgan_forward(sd_ae, cfg["autoencoder"], batch["mel"], batch["mel_length"], False, [(100, 140)] * 2)
            t0 = time.perf_counter()
            O.msmcvqgan_forward(sd_ae, cfg["autoencoder"], batch["mel"], batch["mel_length"], False, [(100, 140)] * 2)
            dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = n, dt
    return best, ncpu


def time_cpu_steps(cfg, steps, warmup, B):
    threads, ncpu = pick_cpu_threads(cfg)
    torch.set_num_threads(threads)
    tr = build_cpu_trainer(cfg)
    batch = synth_batch(B, 99)
    win = [(100, 100 + WIN_FRAMES)] * B
    for _ in range(warmup):
        tr.step(batch["mel"], batch["mel_length"], batch["wav"], win)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.step(batch["mel"], batch["mel_length"], batch["wav"], win)
    dt = (time.perf_counter() - t0) / max(1, steps)
    return B * T_FRAMES / dt, dt, threads, ncpu


def run_reference(args, rank):
    cfg = load_cfg()
    if rank != 0:
        return
    # same config (B=16 per step) and the caller's step / warm-up counts; a step takes ~4 s on the host cores, so the
    # counts are only capped where the whole run would leave the "few minutes" budget
    steps, warmup = min(args.steps, 40), min(args.warmup, 5)
    value, dt, threads, ncpu = time_cpu_steps(cfg, steps, warmup, CPU_SAMPLE_B)
    sample = ("full GAN train step at the GPU arm's own batch (B=%d, T=%d), %d timed steps after %d warm-up; %d torch "
              "threads = fastest of a sweep up to the host's %d logical CPUs") % (
        CPU_SAMPLE_B, T_FRAMES, steps, warmup, threads, ncpu)
    print(json.dumps({
        "impl": "reference", "metric": "mel-frames/sec MSMC-VQ-GAN train step", "value": value,
        "unit": "mel-frames/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "mel-frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "mel-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def time_reference_gpu(cfg, device, steps=5, warmup=2):
    """The comparator SURVEY 8(d) asks for beside the CPU arm: the reference's own step (the oracle port -- plain
    torch functionals, i.e. cuDNN / cuBLAS / cuFFT with torch's default TF32 policy: cudnn.allow_tf32 = True,
    matmul fp32) run EAGERLY on the same B200, same batch, CUDA events.  It keeps the reference's host syncs (the
    per-sample EMA slicing, `.item()`-style reads), which is what a user of the reference gets on this GPU."""
    from oracle.train_step import OracleTrainer
    try:
        sd_ae, sd_d = init_state_dicts(cfg, 1234)
        sd_ae = {k: v.to(device) for k, v in sd_ae.items()}
        sd_d = {k: v.to(device) for k, v in sd_d.items()}
        tr = OracleTrainer(sd_ae, sd_d, cfg, cfg["trainer"], cfg["optimizer"]["_default"], use_dropout=True)
        batch = synth_batch(B_PER_GPU, 99, device=device)
        win = [(100, 100 + WIN_FRAMES)] * B_PER_GPU
        for _ in range(warmup):
            tr.step(batch["mel"], batch["mel_length"], batch["wav"], win)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            tr.step(batch["mel"], batch["mel_length"], batch["wav"], win)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out = {"value": B_PER_GPU * T_FRAMES / (ms * 1e-3), "unit": "mel-frames/s", "ms_per_step": ms,
               "steps": steps, "warmup": warmup, "kind": "port",
               "what": "oracle port of the reference step (torch functionals -> cuDNN/cuBLAS/cuFFT, eager, torch "
                       "default precision flags: cudnn.allow_tf32=%s, matmul.allow_tf32=%s) on the same GPU and batch"
                       % (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)}
        del tr, sd_ae, sd_d
        torch.cuda.empty_cache()
        return out
    except Exception as e:          # report, never take the bench line down
        return {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}


def workload_config(n):
    return {"workload": "MSMC-VQ-GAN full GAN train step (AE fwd/bwd + MelLoss + UnivNet MRD/MPD D-step + G-step + "
                        "AdamW), CSMSC yaml architecture, 2-stage 4-head VQ with %d codewords/head" % K_CODEWORDS,
            "batch_per_gpu": B_PER_GPU, "global_batch": B_PER_GPU * n, "mel_frames": T_FRAMES, "n_mels": N_MELS,
            "sample_rate": 24000, "hop": HOP, "vocoder_window_frames": WIN_FRAMES, "parallelism": "dp%d" % n,
            "l2_policy": "per-step working set (activations + 51M fp32 params, grads, Adam moments ~ 1.5 GB) "
                         "exceeds the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------ GPU arm
def build_gpu_trainer(cfg, device, distributed, rank, world, use_graph=False, reference_schedule=False):
    from msmctts.tasks.msmc_tts import MSMCTTS
    from msmctts.trainers.msmctts_trainer import VQGANTrainer
    from msmctts.utils.config import Config
    ycfg = {"id": "bench", "task": {"_name": "MSMCTTS", "_mode": "train_autoencoder",
                                    "autoencoder": dict(cfg["autoencoder"], _name="MSMCVQGAN"),
                                    "discriminator": dict(cfg["discriminator"], _name="UnivNetDiscriminator")},
            "trainer": dict(cfg["trainer"]), "optimizer": cfg["optimizer"],
            "dataset": {"_name": "SyntheticMelDataset", "samplerate": 24000, "feature": ["mel", "wav"],
                        "frameshift": [HOP, 1]},
            "dataloader": {"batch_size": B_PER_GPU * world, "num_workers": 0}}
    config = Config(ycfg)
    torch.manual_seed(config.seed)
    task = MSMCTTS(config, mode="train")
    kwargs = config.trainer.to_dict()
    kwargs.pop("_name")
    kwargs["warmup_steps"] = 0      # bench the post-warm-up (GAN) step
    kwargs["cuda_graph"] = use_graph
    kwargs["reference_schedule"] = reference_schedule
    trainer = VQGANTrainer(config, task, num_gpus=world, rank=rank, **kwargs)
    trainer.build_optimizer()
    task.train()
    return trainer


def vq_bandwidth(device, pk, K=K_CODEWORDS):
    """VQ search kernel alone: at-config (both stages of one forward: N = 3840 + 960 rows) and an N sweep."""
    from msmctts._b200 import functional as Fn
    out = {"codewords_per_head": K,
           "kernel": "two-phase tcgen05 search (vq_search_umma_kernel) from %d rows on, CUDA-core cluster kernel "
                     "below (MSMC_VQ_UMMA=%s); identical results" % (Fn.VQ_UMMA_MIN_ROWS, Fn.VQ_UMMA)}
    heads, dim = 4, 64
    embed = torch.randn(heads, dim, K, device=device)

    def run(n, reps):
        # `reps` launches captured in one CUDA graph: the kernel at the training shape is shorter than the host-side
        # cost of one eager call (4 output allocations + ctypes), which an eager loop would measure instead
        z = torch.randn(n, heads * dim, device=device)
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
        ms = e0.elapsed_time(e1) / reps
        del g
        # algorithmic bytes: read z, codebooks; write quant_raw, quant_st, diff, idx
        byt = n * heads * dim * 4 * 3 + n * dim * 4 + n * heads * 8 + heads * dim * K * 4
        return byt / (ms * 1e-3) / 1e9, ms

    # the single-tile launches take one of two values depending on where the graph's buffers land (DESIGN.md
    # section 5, profiles/r02_vq_small_probe.txt): five fresh graphs each, the MEDIAN is reported, all trials listed
    t1 = sorted(run(3840, 20)[1] for _ in range(5))
    t2 = sorted(run(960, 20)[1] for _ in range(5))
    ms1, ms2 = t1[2], t2[2]
    n_bytes = (3840 + 960) * (heads * dim * 4 * 3 + dim * 4 + heads * 8) + 2 * heads * dim * K * 4
    out["at_config"] = {"rows": [3840, 960], "us": [ms1 * 1e3, ms2 * 1e3],
                        "us_trials": [[round(t * 1e3, 1) for t in t1], [round(t * 1e3, 1) for t in t2]],
                        "gbs": n_bytes / ((ms1 + ms2) * 1e-3) / 1e9,
                        "frac_of_hbm_peak": n_bytes / ((ms1 + ms2) * 1e-3) / 1e9 / pk["hbm_gbs"],
                        "note": "fixed-cost bound, not HBM bound: %.1f MB of traffic is < 3 us at the HBM peak "
                                "(DESIGN.md section 3)" % (n_bytes / 1e6)}
    sweep = {}
    for p in (12, 14, 16, 18, 20, 22):
        g, ms = run(1 << p, 10 if p < 20 else 3)
        sweep["2^%d" % p] = round(g, 1)
    out["sweep_gbs"] = sweep
    out["asymptote_frac_of_hbm_peak"] = max(sweep.values()) / pk["hbm_gbs"]
    return out


def vq_cpu_port(threads):
    """the reference's MultiHeadQuantize search (oracle port of vqgantts/modules.py:24-33,137-151, eval mode) on the
    host cores at the training shape: the CPU figure that belongs next to `vq_argmin.at_config`"""
    from oracle import ref_modules as O
    torch.set_num_threads(threads)
    heads, dim, K = 4, 64, K_CODEWORDS
    g = torch.Generator().manual_seed(5)
    sd = {"q.quantizers.%d.embed" % h: torch.randn(dim, K, generator=g) for h in range(heads)}
    us, n_bytes = [], 0
    for t in (T_FRAMES, T_FRAMES // 4):
        x = torch.randn(B_PER_GPU, t, heads * dim, generator=g)
        with torch.no_grad():
            O.multihead_quantize(sd, "q.", x, heads)
            reps = 5
            t0 = time.perf_counter()
            for _ in range(reps):
                O.multihead_quantize(sd, "q.", x, heads)
            us.append((time.perf_counter() - t0) / reps * 1e6)
        n = B_PER_GPU * t
        n_bytes += n * (heads * dim * 4 * 3 + dim * 4 + heads * 8) + heads * dim * K * 4
    return {"rows": [B_PER_GPU * T_FRAMES, B_PER_GPU * T_FRAMES // 4], "us": us, "cores": threads, "kind": "port",
            "gbs": n_bytes / (sum(us) * 1e-6) / 1e9}


def run_b200(args, rank, world, local_rank):
    import torch.distributed as dist
    from msmctts._b200 import lib as L
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    L.load()
    distributed = world > 1
    if distributed and not dist.is_initialized():
        dist.init_process_group("nccl")
    cfg = load_cfg()
    trainer = build_gpu_trainer(cfg, device, distributed, rank, world, use_graph=not args.no_graph,
                                reference_schedule=args.reference_schedule)
    pk = peaks()
    fixed_win = [(100, 100 + WIN_FRAMES)] * B_PER_GPU
    dev_batch = synth_batch(B_PER_GPU, 1000 + rank, device=device)
    host_batches = [synth_batch(B_PER_GPU, 2000 + rank * 16 + i, pin=True) for i in range(4)]

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
            torch.cuda.synchronize()

    def step_resident(i):
        return trainer.train_step(dev_batch, iteration=10 + i, frame_windows=fixed_win)

    def step_e2e(i):
        hb = host_batches[i % len(host_batches)]
        batch = {k: v.to(device, non_blocking=True) for k, v in hb.items()}
        log = trainer.train_step(batch, iteration=10 + i, frame_windows=fixed_win)
        return float(log["loss"]["g_loss"])          # device -> host read of the step's loss

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if distributed:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    # the first three calls are eager (the graph is captured on the fourth): count the kernels of the THIRD one, a
    # steady-state step (the first step builds operand images inline, later ones batch them in the weight prefetch)
    step_resident(0)
    step_resident(1)
    l0 = L.launch_count
    step_resident(2)
    launches = L.launch_count - l0
    for i in range(max(3, args.warmup) + 1):
        step_resident(3 + i)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_step = timed(step_resident, args.steps)
    clocks = sampler.stop() if sampler else None
    for i in range(2):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)
    frames = B_PER_GPU * T_FRAMES * world
    value = frames / (ms_step * 1e-3)
    e2e_value = frames / (ms_e2e * 1e-3)
    h2d = sum(v.numel() * v.element_size() for v in host_batches[0].values())

    roof, families, vq = None, None, None
    # ---- roofline pass: CUDA events around every C-ABI call of two more EAGER steps (same stream).  Every rank runs
    # the steps (they contain the gradient all-reduce); rank 0 alone records and reports.
    trainer.use_cuda_graph = False
    # one stream for this pass: with the branch / weight-gradient / prefetch side streams on, the time between two
    # events on one stream would include other streams' kernels sharing the GPU
    from msmctts._b200 import functional as Fn
    Fn.BRANCH_STREAMS = Fn.WGRAD_STREAM = Fn.PREFETCH_WEIGHTS = False
    if rank == 0:
        L.profile_begin()
    for i in range(2):
        # An eager step is ~2000 ctypes calls at ~40 us of host time each: on an idle GPU an event pair around a call
        # would time the HOST (the start event fires long before the launch arrives).  A spin kernel keeps the device
        # busy while the host enqueues the step, so every event pair brackets back-to-back device work only.
        torch.cuda.synchronize()
        torch.cuda._sleep(int(0.16 * 1.9e9))
        step_resident(100 + i)
    torch.cuda.synchronize()
    if rank == 0:
        prof = L.profile_end()
        fam = {}
        for name, meta, e0, e1 in prof:
            f = fam.setdefault(name, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "calls": 0})
            f["ms"] += e0.elapsed_time(e1)
            f["calls"] += 1
            if meta:
                f["flops"] += meta["flops"]
                f["bytes"] += meta["bytes"]
        if os.environ.get("MSMC_BENCH_DUMP"):
            shapes = {}
            for name, meta, e0, e1 in prof:
                key = name + " | " + (meta["shape"] if meta else "-")
                d = shapes.setdefault(key, {"ms": 0.0, "calls": 0, "gflop": 0.0})
                d["ms"] += e0.elapsed_time(e1) / 2
                d["calls"] += 0.5
                d["gflop"] += (meta["flops"] / 2e9) if meta else 0.0
            rows = sorted(shapes.items(), key=lambda kv: -kv[1]["ms"])
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", os.environ["MSMC_BENCH_DUMP"]), "w") as f:
                for k, d in rows:
                    f.write("%9.3f ms  %5.1f calls  %9.2f GFLOP  %7.2f TF/s  %s\n" % (
                        d["ms"], d["calls"], d["gflop"], d["gflop"] / max(d["ms"], 1e-9), k))
        tot_ms = sum(f["ms"] for f in fam.values())
        families = {k: {"ms_per_step": round(v["ms"] / 2, 3), "calls_per_step": v["calls"] // 2,
                        "share": round(v["ms"] / tot_ms, 3)} for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
        top = max(fam.items(), key=lambda kv: kv[1]["ms"])
        tname, t = top
        traffic, traffic_src = family_traffic(tname)
        if t["flops"] > 0:
            ach = t["flops"] / (t["ms"] * 1e-3) / 1e12
            roof = {"kernel": tname, "bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops"],
                    "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"], "traffic": traffic,
                    "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                    "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": t["bytes"] / max(1, t["calls"]),
                    "peak_source": pk["source"] + "; the kernel computes fp32-accurate results as 3xTF32 (three "
                                   "tcgen05 kind::tf32 MMAs per K-step), its algorithmic fp32 flops are measured "
                                   "against the dense bf16 tensor peak",
                    "tf32_mode_ceilings": {
                        "unit": "TFLOP/s of algorithmic fp32 work",
                        "plain_tf32": pk["bf16_tflops"] / 2, "three_x_tf32": pk["bf16_tflops"] / 6,
                        "frac_of_three_x_tf32": ach / (pk["bf16_tflops"] / 6),
                        "note": "kind::tf32 retires half the bf16 MACs per cycle and 3xTF32 issues three MMAs per "
                                "K-step: the pipe can deliver at most peak/6 of fp32-accurate work"},
                    "launches_per_step": t["calls"] // 2, "avg_launch_us": t["ms"] * 1e3 / t["calls"],
                    "algorithmic_gflop_per_step": t["flops"] / 2 / 1e9,
                    "share_of_kernel_time": round(t["ms"] / tot_ms, 3)}
        else:
            ach = t["bytes"] / (t["ms"] * 1e-3) / 1e9 if t["bytes"] else 0.0
            roof = {"kernel": tname, "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / pk["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": pk["source"]}
        vq = vq_bandwidth(device, pk, K_CODEWORDS)
        vq["k64"] = vq_bandwidth(device, pk, 64)       # the reference yaml's own codebook size (HBM-bound there)

    cpu, ref_gpu = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref_gpu = time_reference_gpu(cfg, device)
        v, dt, threads, ncpu = time_cpu_steps(cfg, 5, 1, CPU_SAMPLE_B)
        cpu = {"value": v, "unit": "mel-frames/s", "cores": threads, "kind": "port",
               "sample": "oracle port of the same full GAN step at the same batch (B=%d), 5 timed steps after 1 "
                         "warm-up (%.1f s/step); %d torch threads = fastest of a sweep up to the host's %d logical "
                         "CPUs" % (CPU_SAMPLE_B, dt, threads, ncpu)}
        if vq is not None:
            vq["cpu_port_at_config"] = vq_cpu_port(threads)
    if rank == 0:
        line = json.dumps({
            "metric": "mel-frames/sec MSMC-VQ-GAN train step", "value": value, "unit": "mel-frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(world), schedule=(
                "reference (4 separate D passes, D gradients computed and discarded in the G step)"
                if args.reference_schedule else
                "same losses and parameter updates as the reference; D step scores cat(fake, real) in one pass, G "
                "step back-propagates only into the autoencoder (the reference zeroes D's G-step gradients unread); "
                "--reference-schedule replays the reference's launches")), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "mel-frames/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4},
            "gpu_launches": launches * args.steps, "gpu_launches_per_step": launches,
            "cuda_graph": not args.no_graph,
            "roofline": roof, "kernel_families": families, "vq_argmin": vq, "cpu_baseline": cpu,
            "reference_gpu": ref_gpu})
        print(line, flush=True)
    if distributed:
        # Tear down in the order that works: the captured graphs hold NCCL all-reduce nodes, so they go first
        # (VQGANTrainer.release_graphs), then the communicator.  In round 1 destroy_process_group() after
        # graph-captured collectives never returned; a watchdog thread still ends the process if that recurs (the
        # JSON line is already out and all device work is complete).
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        threading.Timer(30.0, lambda: os._exit(0)).start()
        trainer.release_graphs()
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of as a CUDA graph")
    ap.add_argument("--reference-schedule", action="store_true",
                    help="replay the reference's launch schedule: 4 separate discriminator passes and discriminator "
                         "gradients computed (then discarded) in the generator step")
    ap.add_argument("--config", default="gan", choices=["gan", "am"],
                    help="gan: full GAN train step (the headline, BASELINE.json configs[1-3]); am: the multi-stage "
                         "predictor train step of configs[4]")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.config == "am":
        if args.impl != "reference" and not torch.cuda.is_available():
            raise SystemExit("bench.py --impl b200 needs a CUDA device (the hot path has no CPU fallback)")
        run_am(args, rank, world, local_rank)
        return
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (the hot path has no CPU fallback)")
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
