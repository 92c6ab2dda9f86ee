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
        if mod.__class__.__name__ in ("MultiStageQuantizer", "DurationPredictor"):
            mod.dropout = 0.0


@pytest.mark.parametrize("K", [256, 64])
@pytest.mark.parametrize("reference_schedule", [False, True], ids=["fused-schedule", "reference-schedule"])
def test_two_train_steps_vs_oracle(reference_schedule, K):
    """both launch schedules (see VQGANTrainer.__init__) must reproduce the reference's losses and updates, at the
    bench's codebook size (K=256) and the reference yaml's (K=64): every logged loss of both steps, and after the
    second update every parameter of both networks and every codebook buffer."""
    import bench
    from oracle.train_step import OracleTrainer
    cfg = bench.load_cfg()
    cfg["autoencoder"]["quantizer_config"]["embedding_sizes"] = K
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
    def codebooks_close(max_bad_codewords, when):
        """EMA buffers vs the oracle's.  One flipped index (the two sides' z differ by fp32 summation order in step 1
        and additionally by Adam's sign-like first update in step 2) moves a count by 0.01 and two embed_avg columns
        by 0.01 * z: count the codewords that are off instead of failing on the first element."""
        sd_gpu = trainer.model.autoencoder.state_dict()
        for k, v in oracle.sd_ae.items():
            last = k.split(".")[-1]
            if last not in ("embed", "embed_avg", "cluster_size"):
                continue
            a, b = sd_gpu[k].detach().cpu(), v.detach()
            bad = ~torch.isclose(a, b, rtol=1e-3, atol=1e-4)
            n_bad = int(bad.reshape(-1, bad.shape[-1]).any(dim=0).sum())
            assert n_bad <= max_bad_codewords, "%s %s: %d codewords differ" % (when, k, n_bad)
            if last == "cluster_size":       # totals are flip-invariant
                assert abs(float(a.sum()) - float(b.sum())) <= 1e-3 * float(b.sum()), k

    for step in range(2):
        log = trainer.train_step(gbatch, iteration=10 + step, frame_windows=win)["loss"]
        ref = oracle.step(batch["mel"], batch["mel_length"], batch["wav"], win)
        for k in ("vq_loss", "frame_loss", "stft_loss", "d_loss_real", "d_loss_fake", "d_loss", "fm_loss",
                  "adv_loss", "g_loss"):
            a, b = float(log[k]), ref[k]
            tol = 2e-3 if step == 0 else 2e-2     # step 2 inherits Adam's sign-like first update
            assert abs(a - b) <= tol * max(abs(b), 1e-3), "step %d %s: %.6f vs %.6f" % (step, k, a, b)
        # 3600 indices per step: <= 2 flips (4 codewords) from summation order alone, a few more after the update
        codebooks_close(4 if step == 0 else max(8, K // 16), "after step %d" % (step + 1))
    # after two AdamW (+ clip) updates and two EMA updates.  lr = 2e-4: an Adam step moves a weight by <= ~lr, and
    # Adam's first updates are sign-like (g / |g|), so a gradient that is ~0 by cancellation may legitimately take
    # the opposite sign on the two sides -> absolute tolerance of 2 steps x lr on parameters; the codebooks (EMA of
    # data, no optimizer) are compared tightly.
    # (the codebooks -- EMA of data, no optimizer -- were compared after each step above)
    lr = cfg["optimizer"]["_default"]["learning_rate"]
    for name, module, ref_sd in (("autoencoder", trainer.model.autoencoder, oracle.sd_ae),
                                 ("discriminator", trainer.model.discriminator, oracle.sd_d)):
        sd_gpu = module.state_dict()
        n_checked, n_tight = 0, 0
        for k, v in ref_sd.items():
            if not v.is_floating_point():
                continue
            a, b = sd_gpu[k].detach().cpu(), v.detach()
            last = k.split(".")[-1]
            if last in ("embed", "embed_avg", "cluster_size"):
                continue                    # checked per step above
            else:
                err = (a - b).abs()
                assert float(err.max()) <= 2.2 * 2 * lr, "%s.%s differs by %.3e" % (name, k, float(err.max()))
                n_tight += int((err <= 0.1 * lr).sum())
                n_checked += err.numel()
        # ...and the overwhelming majority of weights agree to a tenth of one step
        assert n_tight >= 0.98 * n_checked, "%s: only %.2f %% of the weights within 0.1 lr" % (
            name, 100.0 * n_tight / max(1, n_checked))


def test_two_train_steps_vs_reference_trainer_golden():
    """Whole-step parity DIRECTLY against the unmodified reference trainer: tests/golden/train_step.pt holds two
    consecutive `VQGANTrainer.train_step` calls of the reference itself (oracle/make_golden.py gen_train_step, small
    config, 20 samples per frame).  Same initial state_dicts, batch and windows; losses of both steps and the
    parameters after the two AdamW + EMA updates.  Tolerances as above: 2e-3 relative on step-1 losses, 2e-2 on
    step 2, 1e-3 absolute on parameters (lr 2e-4 x 2 steps bounds any parameter change by 4e-4 per step)."""
    from msmctts.tasks.msmc_tts import MSMCTTS
    from msmctts.trainers.msmctts_trainer import VQGANTrainer
    from msmctts.utils.config import Config
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = torch.load(os.path.join(root, "tests", "golden", "train_step.pt"), map_location="cpu", weights_only=False)
    cfg, tcfg = g["cfg"], dict(g["trainer"])
    hop = tcfg.pop("frameshift")
    tcfg.pop("sample_rate")
    ycfg = {"id": "golden", "task": {"_name": "MSMCTTS", "_mode": "train_autoencoder",
                                     "autoencoder": dict(cfg["autoencoder"], _name="MSMCVQGAN"),
                                     "discriminator": dict(cfg["discriminator"], _name="UnivNetDiscriminator")},
            "trainer": dict(tcfg, _name="VQGANTrainer"), "optimizer": {"_default": g["optimizer"]},
            "dataset": {"_name": "SyntheticMelDataset", "samplerate": 24000, "feature": ["mel", "wav"],
                        "frameshift": [hop, 1]},
            "dataloader": {"batch_size": 2, "num_workers": 0}}
    config = Config(ycfg)
    task = MSMCTTS(config, mode="train")
    task.autoencoder.load_state_dict(g["sd_ae"], strict=True)
    task.discriminator.load_state_dict(g["sd_d"], strict=True)
    kwargs = config.trainer.to_dict()
    kwargs.pop("_name")
    kwargs["cuda_graph"] = False
    trainer = VQGANTrainer(config, task, num_gpus=1, rank=0, **kwargs)
    trainer.build_optimizer()
    task.train()
    _no_dropout(task)           # the fixture's harness tweak: ResStack's hard-wired Dropout(0.1) -> 0
    dev = torch.device("cuda:0")
    batch = {"mel": g["mel"].to(dev), "mel_length": g["length"].to(dev), "wav": g["wav"].to(dev)}
    keys = ("vq_loss", "frame_loss", "stft_loss", "d_loss_real", "d_loss_fake", "d_loss", "fm_loss", "adv_loss",
            "g_loss")
    for n, st in enumerate(g["steps"]):
        log = trainer.train_step(batch, iteration=1 + n, frame_windows=st["windows"])["loss"]
        tol = 2e-3 if n == 0 else 2e-2
        for k in keys:
            a, b = float(log[k]), st["losses"][k]
            assert abs(a - b) <= tol * max(abs(b), 1e-3), "step %d %s: %.6f vs reference %.6f" % (n, k, a, b)
    for name, module, ref_sd in (("autoencoder", task.autoencoder, g["sd_ae_after"]),
                                 ("discriminator", task.discriminator, g["sd_d_after"])):
        sd = module.state_dict()
        for k, v in ref_sd.items():
            if v.is_floating_point() and k.split(".")[-1] not in ("embed", "embed_avg"):
                err = float((sd[k].detach().cpu() - v).abs().max())
                assert err <= 1e-3, "%s.%s differs from the reference by %.3e after two steps" % (name, k, err)


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


def test_predictor_train_step_vs_oracle():
    """BASELINE.json configs[4] at its real architecture (CSMSC AM yaml: d_model 600, FFT x 6 encoder + 2 x FFT x 6
    decoders, 114.8 M parameters; frozen CSMSC autoencoder, K=256; 24 phonemes x 10 frames = T 240) on a B=4 slice:
    PredictorTrainer.train_step (reference trainers/msmctts_trainer.py:237-286) vs the CPU oracle port from the same
    state -- every logged loss of two consecutive steps (mse + triplet-sum embedding losses, duration loss, gradient
    norm) and every parameter after the two clipped Adam updates."""
    import bench
    from oracle.train_step import OraclePredictorTrainer
    cfg, am = bench.load_cfg(), bench.load_am_cfg()
    dev = torch.device("cuda:0")
    trainer = bench.build_am_trainer(cfg, am, dev, False, 0, 1)
    _no_dropout(trainer.model)
    sd_p = {k: v.detach().cpu().clone() for k, v in trainer.model.predictor.state_dict().items()}
    sd_ae = {k: v.detach().cpu().clone() for k, v in trainer.autoencoder.state_dict().items()}
    oracle = OraclePredictorTrainer(sd_p, sd_ae, am["predictor"], cfg["autoencoder"], am["trainer"],
                                    am["optimizer"]["_default"])
    B = 4
    batch = bench.synth_am_batch(B, 21)
    batch["text_length"] = torch.tensor([24, 24, 20, 17])
    dur = batch["dur"].clone()
    dur[0, :4] = torch.tensor([12, 8, 11, 9])                     # uneven durations, same total
    for b, n in enumerate(batch["text_length"].tolist()):          # padded phonemes: id 0, duration 0
        batch["text"][b, n:] = 0
        dur[b, n:] = 0
    batch["dur"] = dur
    batch["mel_length"] = dur.sum(1)
    keys = ("total_loss", "embed_loss_mse_0", "embed_loss_triple_sum_0", "embed_loss_mse_1", "embed_loss_triple_sum_1",
            "dur_loss", "grad_norm")
    for step in range(2):
        log = trainer.train_step({k: v.to(dev) for k, v in batch.items()}, iteration=step)["loss"]
        ref = oracle.step(batch["text"], batch["text_length"], batch["dur"], batch["mel"], batch["mel_length"])
        bad = []
        for k in keys:
            a, b = float(log[k]), ref[k]
            tol = 2e-3 if step == 0 else 2e-2
            if not abs(a - b) <= tol * max(abs(b), 1e-3):
                bad.append("step %d %s: %.6f vs %.6f" % (step, k, a, b))
        assert not bad, "; ".join(bad)
    lr = am["optimizer"]["_default"]["learning_rate"]
    sd_gpu = trainer.model.predictor.state_dict()
    n_checked = n_tight = 0
    for k, v in oracle.sd_p.items():
        if not v.is_floating_point():
            continue
        err = (sd_gpu[k].detach().cpu() - v.detach()).abs()
        assert float(err.max()) <= 2.2 * 2 * lr, "predictor.%s differs by %.3e" % (k, float(err.max()))
        n_tight += int((err <= 0.1 * lr).sum())
        n_checked += err.numel()
    assert n_tight >= 0.98 * n_checked, "only %.2f %% of the weights within 0.1 lr" % (100.0 * n_tight / n_checked)
