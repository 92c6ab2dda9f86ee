"""PENDING (written at the end of round 1 with no GPU time left; NOT collected by pytest -- the file name does not
match test_*.py).  First thing to validate on a B200 in round 2, then move to tests/test_train_step_gpu.py:

    python -m pytest tests/pending/gpu_reference_trainer_parity.py -q -p no:cacheprovider

Whole-step parity of the CUDA trainer DIRECTLY against the unmodified reference trainer: tests/golden/train_step.pt
holds two consecutive `VQGANTrainer.train_step` calls of the reference itself (oracle/make_golden.py gen_train_step,
small config, 20 samples per frame).  Same initial state_dicts, same batch, same windows; losses of both steps and the
parameters after the two AdamW + EMA updates are compared.  Tolerances follow tests/test_train_step_gpu.py: 2e-3
relative on step-1 losses, 2e-2 on step 2 (Adam's first update is sign-like), 1e-3 absolute on parameters
(lr 2e-4 x 2 steps bounds any parameter change by 4e-4 per step).
"""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "msmc-tts_b200"))

pytestmark = pytest.mark.gpu


def test_two_train_steps_vs_reference_trainer_golden():
    from msmctts.tasks.msmc_tts import MSMCTTS
    from msmctts.trainers.msmctts_trainer import VQGANTrainer
    from msmctts.utils.config import Config
    g = torch.load(os.path.join(ROOT, "tests", "golden", "train_step.pt"), map_location="cpu", weights_only=False)
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
    trainer = VQGANTrainer(config, task, num_gpus=1, rank=0, **kwargs)
    trainer.build_optimizer()
    task.train()
    for mod in task.modules():          # the fixture's harness tweak: ResStack's hard-wired Dropout(0.1) -> 0
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
        if hasattr(mod, "p_dropout"):
            mod.p_dropout = 0.0
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
