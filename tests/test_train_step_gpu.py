"""Whole-step parity: VQGANTrainer.train_step on the CUDA path vs the CPU oracle port of the reference's step
(oracle/train_step.py) from the SAME initial state and batch, dropout off, two consecutive steps (the second one
sees the AdamW-updated weights and EMA-updated codebooks of the first).  GAN training amplifies differences, so the
comparison is on one/two steps from identical state, not on trajectories."""
import copy
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _no_dropout(model):
    for mod in model.modules():
        if hasattr(mod, "p_dropout"):
            mod.p_dropout = 0.0
        if hasattr(mod, "attn_dropout"):
            mod.attn_dropout = 0.0
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
        if mod.__class__.__name__ == "MultiStageQuantizer":
            mod.dropout = 0.0


@pytest.mark.parametrize("reference_schedule", [False, True], ids=["fused-schedule", "reference-schedule"])
def test_two_train_steps_vs_oracle(reference_schedule):
    """both launch schedules (see VQGANTrainer.__init__) must reproduce the reference's losses and updates"""
    import bench
    from oracle.train_step import OracleTrainer
    cfg = bench.load_cfg()
    cfg["autoencoder"]["quantizer_config"]["embedding_sizes"] = 64
    dev = torch.device("cuda:0")
    trainer = bench.build_gpu_trainer(cfg, dev, False, 0, 1, reference_schedule=reference_schedule)
    _no_dropout(trainer.model)
    sd_ae = {k: v.detach().cpu().clone() for k, v in trainer.model.autoencoder.state_dict().items()}
    sd_d = {k: v.detach().cpu().clone() for k, v in trainer.model.discriminator.state_dict().items()}
    oracle = OracleTrainer(sd_ae, sd_d, cfg, cfg["trainer"], cfg["optimizer"]["_default"], use_dropout=False)
    B = 3
    batch = bench.synth_batch(B, 5)
    batch["mel_length"] = torch.tensor([240, 200, 131])
    win = [(100, 140), (60, 100), (0, 40)]
    gbatch = {k: v.to(dev) for k, v in batch.items()}
    for step in range(2):
        log = trainer.train_step(gbatch, iteration=10 + step, frame_windows=win)["loss"]
        ref = oracle.step(batch["mel"], batch["mel_length"], batch["wav"], win)
        for k in ("vq_loss", "frame_loss", "stft_loss", "d_loss_real", "d_loss_fake", "d_loss", "fm_loss",
                  "adv_loss", "g_loss"):
            a, b = float(log[k]), ref[k]
            tol = 2e-3 if step == 0 else 2e-2     # step 2 inherits Adam's sign-like first update
            assert abs(a - b) <= tol * max(abs(b), 1e-3), "step %d %s: %.6f vs %.6f" % (step, k, a, b)
    # codebooks after two EMA updates
    sd_gpu = trainer.model.autoencoder.state_dict()
    for k, v in oracle.sd_ae.items():
        if k.split(".")[-1] == "cluster_size":
            assert torch.allclose(sd_gpu[k].cpu(), v.detach(), rtol=1e-2, atol=1e-3), k


def test_cuda_graph_replay_matches_eager_steps():
    """bench.py times the step as a CUDA-graph replay (3 eager warm-ups, capture, replay): the captured step --
    forked branch / weight-gradient / prefetch streams, graph-private pointer tables of the multi-tensor kernels,
    device-side step counter -- must produce the same losses and the same parameters as the eager step."""
    import bench
    cfg = bench.load_cfg()
    cfg["autoencoder"]["quantizer_config"]["embedding_sizes"] = 64
    dev = torch.device("cuda:0")
    B = 2
    batch = {k: v.to(dev) for k, v in bench.synth_batch(B, 11).items()}
    win = [(100, 140), (20, 60)]
    logs, params = [], []
    for use_graph in (False, True):
        trainer = bench.build_gpu_trainer(cfg, dev, False, 0, 1, use_graph=use_graph)
        _no_dropout(trainer.model)
        out = []
        for step in range(6):            # graph mode: steps 0-2 eager warm-up, step 3 capture + replay, 4-5 replays
            log = trainer.train_step(batch, iteration=10 + step, frame_windows=win)["loss"]
            out.append({k: float(v) for k, v in log.items() if torch.is_tensor(v)})
        torch.cuda.synchronize()
        logs.append(out)
        params.append([p.detach().clone() for p in trainer.model.parameters()])
    for step, (a, b) in enumerate(zip(*logs)):
        for k in ("vq_loss", "frame_loss", "stft_loss", "d_loss", "fm_loss", "adv_loss", "g_loss"):
            assert abs(a[k] - b[k]) <= 1e-4 * max(abs(a[k]), 1e-3), "step %d %s: eager %.7f graph %.7f" % (
                step, k, a[k], b[k])
    worst = max(float((x - y).abs().max()) for x, y in zip(*params))
    assert worst <= 1e-4, "parameters after 6 steps differ by %.3e" % worst
