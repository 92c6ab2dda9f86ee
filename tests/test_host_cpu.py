"""CPU-only checks: the C-ABI library loads and exports every symbol include/msmc_b200.h declares, the drop-in
boundary (names, kwargs, state_dict keys, yaml registry) matches the reference, the product path never touches the
oracle and fails loudly without CUDA, and the data-parallel plumbing works at world_size 2 over gloo."""
import json
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "msmc-tts_b200")


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from msmctts._b200 import lib as L
    lib = L.load()
    header = open(os.path.join(ROOT, "include", "msmc_b200.h")).read()
    declared = set(re.findall(r"\b(msmc_[a-z0-9_]+)\s*\(", header))
    declared -= {"msmc_status", "msmc_xform", "msmc_conv_geom"}
    assert len(declared) >= 25
    for name in sorted(declared):
        assert hasattr(lib, name), "declared in the header but not exported: " + name
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    assert lib.msmc_version() >= 100


def test_conv_geom_struct_matches_header_layout():
    from msmctts._b200.lib import ConvGeom
    import ctypes
    # 17 int32 (+pad) , 9 int64, then (int32,float) x2
    assert ctypes.sizeof(ConvGeom) == 17 * 4 + 4 + 9 * 8 + 16


def test_state_dict_keys_match_reference(golden):
    import copy
    from msmctts.networks.hifigan import UnivNetDiscriminator
    from msmctts.networks.vqgantts import MSMCVQGAN
    from msmctts.utils.config import ConfigItem
    with open(os.path.join(ROOT, "tests", "golden", "csmsc_config.json")) as f:
        cfg = json.load(f)
    with open(os.path.join(ROOT, "tests", "golden", "csmsc_state_dict_keys.json")) as f:
        ref = json.load(f)
    c = copy.deepcopy(cfg["autoencoder"])
    ae = MSMCVQGAN(c["in_dim"], c["n_model_size"], ConfigItem(c["encoder_config"]), ConfigItem(c["quantizer_config"]),
                   ConfigItem(c["frame_decoder_config"]), ConfigItem(c["decoder_config"]), c["pred_mel"])
    d = UnivNetDiscriminator(ConfigItem(cfg["discriminator"]["mrd_config"]), ConfigItem(cfg["discriminator"]["mpd_config"]))
    for name, mod in (("autoencoder", ae), ("discriminator", d)):
        mine = {k: list(v.shape) for k, v in mod.state_dict().items()}
        assert list(mine.keys()) == list(ref[name].keys()), name + ": key order"
        assert mine == ref[name], name + ": shapes"
        assert sum(p.numel() for p in mod.parameters()) == ref[name + "_params"]
    # a reference checkpoint loads strictly
    ae.load_state_dict({k: torch.zeros(s) for k, s in ref["autoencoder"].items()}, strict=True)


def test_yaml_registry_builds_task_by_name(tmp_path):
    from msmctts.tasks import build_task
    from msmctts.utils.config import Config
    with open(os.path.join(ROOT, "tests", "golden", "csmsc_config.json")) as f:
        cfg = json.load(f)
    y = {"task": {"_name": "MSMCTTS", "_mode": "train_autoencoder",
                  "autoencoder": dict(cfg["autoencoder"], _name="MSMCVQGAN"),
                  "discriminator": dict(cfg["discriminator"], _name="UnivNetDiscriminator")},
         "dataset": {"samplerate": 24000, "feature": ["mel", "wav"], "frameshift": [300, 1]}}
    task = build_task(Config(y), "train")
    assert [n for n, _ in task.named_children()] == ["autoencoder", "discriminator"]
    assert task.autoencoder.quantizer.quantizer[0].n_head == 4
    assert Config(y).seed == 1234 and Config(y).distributed.dist_backend == "nccl"


def test_multihead_codebook_views_survive_to_and_load_state_dict():
    from msmctts.networks.vqgantts.modules import MultiHeadQuantize
    q = MultiHeadQuantize(256, 64, 4)
    e, a, c = q._stacked()
    assert q.quantizers[2].embed.data_ptr() == e[2].data_ptr()
    sd = {k: torch.randn_like(v) for k, v in q.state_dict().items()}
    q.load_state_dict(sd)
    e2, _, _ = q._stacked()
    assert torch.equal(e2[3], sd["quantizers.3.embed"])
    q.double().float()       # _apply replaces the buffers: the views must be re-established
    e3, _, _ = q._stacked()
    assert q.quantizers[1].embed.data_ptr() == e3[1].data_ptr() and torch.equal(e3[3], sd["quantizers.3.embed"])


def test_product_path_has_no_cpu_fallback_and_never_imports_oracle():
    from msmctts._b200 import functional as Fn
    from msmctts._b200.lib import MsmcError
    with pytest.raises(MsmcError):
        Fn.vq_quantize(torch.randn(4, 256), torch.randn(4, 64, 64), 4, 64)
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(dirpath, f)
                assert "/root/reference" not in src, os.path.join(dirpath, f)


def test_missing_library_raises(monkeypatch):
    from msmctts._b200 import lib as L
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", "/nonexistent/libmsmc_b200.so")
    with pytest.raises(L.MsmcError):
        L.load()


def test_window_gather_matches_python_slicing():
    from msmctts.trainers.msmctts_trainer import VQGANTrainer
    x = torch.randn(3, 50, 7)
    starts = torch.tensor([0, 10, 37])
    got = VQGANTrainer.gather_windows(x, starts, 13)
    want = torch.stack([x[i, s:s + 13] for i, s in enumerate(starts.tolist())])
    assert torch.equal(got, want)


_DDP = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from msmctts.distributed.distributed import init_distributed, apply_gradient_allreduce
rank, world = int(sys.argv[2]), 2
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=sys.argv[3], RANK=str(rank), WORLD_SIZE="2")
init_distributed(rank, world, "g", dist_backend="gloo")
torch.manual_seed(100 + rank)                      # ranks start with DIFFERENT weights
task = torch.nn.Module()
task.autoencoder = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4))
task.discriminator = torch.nn.Linear(4, 1)
task.autoencoder.register_buffer("embed", torch.randn(3))
apply_gradient_allreduce(task)                     # broadcast from rank 0
ref = [p.detach().clone() for p in task.parameters()]
gathered = [torch.zeros_like(ref[0]) for _ in range(world)]
dist.all_gather(gathered, ref[0])
assert torch.equal(gathered[0], gathered[1]), "initial broadcast failed"
x = torch.randn(5, 8)                              # different data per rank
red = task.grad_reducers["autoencoder"]
red.arm()
task.discriminator(task.autoencoder(x)).sum().backward()
red.finish()
g = task.autoencoder[0].weight.grad.clone()
gathered = [torch.zeros_like(g) for _ in range(world)]
dist.all_gather(gathered, g)
assert torch.allclose(gathered[0], gathered[1]), "autoencoder grads were not averaged"
# zero-copy hand-back: after finish() every .grad is a view into its bucket's reduced flat buffer
ps = [p for p in task.autoencoder.parameters()]
assert len({p.grad.untyped_storage().data_ptr() for p in ps}) == 1, "grads must alias one flat bucket"
assert all(p.grad.shape == p.shape and p.grad.is_contiguous() for p in ps)
gd = task.discriminator.weight.grad.clone()
gathered = [torch.zeros_like(gd) for _ in range(world)]
dist.all_gather(gathered, gd)
assert not torch.allclose(gathered[0], gathered[1]), "discriminator grads must stay local in the G step"
# averaged gradient equals the mean of the two local gradients
task.zero_grad()
task.discriminator(task.autoencoder(x)).sum().backward()
local = task.autoencoder[0].weight.grad.clone()
both = [torch.zeros_like(local) for _ in range(world)]
dist.all_gather(both, local)
assert torch.allclose(g, (both[0] + both[1]) / 2, atol=1e-6)
dist.barrier()
print("rank", rank, "ok")
'''


def test_gradient_allreduce_world_size_2_gloo(tmp_path):
    script = tmp_path / "ddp.py"
    script.write_text(_DDP)
    port = str(29500 + os.getpid() % 1000)
    procs = [subprocess.Popen([sys.executable, str(script), PKG, str(r), port], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_tile_policy_and_workspace_queries_run_without_a_gpu():
    """host-side planning entry points of the C-ABI are pure functions of the geometry (no device work): the N-tile
    policy (widest tile unless the grid would fall below 16 CTAs, never pad N by more than 2x) and
    the operand-image / workspace sizes the Python side allocates from"""
    from msmctts._b200 import lib as L
    lib = L.load()
    assert lib.msmc_umma_tile_n(1024, 3840) == 128      # FFN-1
    assert lib.msmc_umma_tile_n(256, 3840) == 128       # FFN-2: 60 CTAs of N = 128 beat 240 of N = 32
    assert lib.msmc_umma_tile_n(256, 960) == 128        # 8 x 2 = 16 CTAs: still the wide tile
    assert lib.msmc_umma_tile_n(256, 256) == 32         # 2 x 2 = 4 CTAs -> narrow tiles for parallelism
    assert lib.msmc_umma_tile_n(32, 192000) == 32       # never pad 32 channels to 64
    assert lib.msmc_umma_tile_n(64, 96000) == 64
    assert lib.msmc_umma_tile_n(1, 12000) == 32
    # image size: T taps x (K/32) chunks x n-tiles x BN rows x 32 floats x (hi, lo planes)
    assert lib.msmc_weight_image_elems(3, 256, 1024, 0, 1, 128) == 3 * 8 * 8 * 128 * 32 * 2
    assert lib.msmc_weight_image_elems(3, 100, 64, 0, 1, 64) == 3 * 4 * 1 * 64 * 32 * 2   # ragged K: 4 chunks, padded
    assert lib.msmc_weight_image_elems(3, 102, 64, 0, 1, 64) == -1        # K must be a multiple of 4 (16-byte rows)
    assert lib.msmc_adam_chunk_elems() > 0 and lib.msmc_l1_chunk_elems() > 0


def test_dense_permutation_of_feature_map_views():
    """the fused feature-matching loss runs on the underlying channels-last buffers: `_dense_perm` must find the
    permutation that makes a permuted view contiguous, and its inverse must restore shape and strides"""
    import torch
    from msmctts._b200.functional import _dense_perm
    base = torch.arange(2 * 5 * 7 * 3, dtype=torch.float32).reshape(2, 5, 7, 3)      # (B, H, W, C)
    for dims in [(0, 3, 1, 2), (0, 3, 2, 1), (0, 1, 2, 3), (3, 0, 2, 1)]:
        view = base.permute(*dims)
        perm = _dense_perm(view)
        assert perm is not None and view.permute(perm).is_contiguous()
        inv = [0] * len(perm)
        for i, d in enumerate(perm):
            inv[d] = i
        back = view.permute(perm).permute(inv)
        assert back.shape == view.shape and back.stride() == view.stride()
    assert _dense_perm(base[:, :, ::2]) is None        # strided slice: not dense under any permutation


def _cli_yaml(tmp_path, **over):
    import yaml
    with open(os.path.join(ROOT, "tests", "golden", "csmsc_config.json")) as f:
        cfg = json.load(f)
    y = {"id": "cli",
         "task": {"_name": "MSMCTTS", "_mode": "train_autoencoder",
                  "autoencoder": dict(cfg["autoencoder"], _name="MSMCVQGAN"),
                  "discriminator": dict(cfg["discriminator"], _name="UnivNetDiscriminator")},
         "trainer": dict(cfg["trainer"]), "optimizer": cfg["optimizer"],
         "lr_scheduler": {"_name": "ExponentialDecayLRScheduler", "warmup_steps": 200000, "decay_scale": 200000,
                          "decay_learning_rate": 0.5, "final_learning_rate": 1e-5},
         # the reference yaml's per-feature lists (examples/csmsc/configs/msmc_vq_gan.yaml)
         "dataset": {"_name": "SyntheticMelDataset", "samplerate": 24000, "feature": ["mel", "wav"],
                     "frameshift": [300, 1], "n_items": 16},
         "dataloader": {"batch_size": 2, "num_workers": 0},
         "training_steps": 3, "iters_per_checkpoint": 2, "save_checkpoint_dir": str(tmp_path / "ckpt")}
    y.update(over)
    path = tmp_path / "cli.yaml"
    with open(path, "w") as f:
        yaml.safe_dump(y, f)
    return str(path)


def test_trainer_loop_checkpoints_and_resumes(tmp_path):
    """host side of `train.py` (reference trainers/base_trainer.py:16-142): yaml -> task -> trainer -> dataloader ->
    lr schedule -> step -> log -> checkpoint -> resume, with the device step stubbed out (no GPU here)"""
    import torch
    from msmctts.tasks import build_task
    from msmctts.trainers import build_trainer
    from msmctts.utils.config import Config
    config = Config(_cli_yaml(tmp_path))
    trainer = build_trainer(config, build_task(config, "train"), num_gpus=0, rank=0)
    seen = []

    def fake_step(batch, iteration):
        seen.append((iteration, tuple(batch["mel"].shape), tuple(batch["wav"].shape)))
        return {"loss": {"g_loss": torch.tensor(1.0 + iteration), "d_loss": 0.5}}
    trainer.train_step = fake_step
    trainer.train()
    assert [s[0] for s in seen] == [0, 1, 2, 3]
    assert seen[0][1:] == ((2, 240, 80), (2, 72000, 1))           # hop = the mel entry of dataset.frameshift
    assert sorted(os.listdir(config.save_checkpoint_dir)) == ["model_2", "train.log"]
    again = build_trainer(config, build_task(config, "train"), num_gpus=0, rank=0)
    again.build_optimizer()
    assert again.attempt_load_checkpoint() == 3                     # resumes after the saved iteration
    for a, b in zip(trainer.model.state_dict().values(), again.model.state_dict().values()):
        assert torch.equal(a, b)
    # checkpoint -> task / sub-network (reference tasks/__init__.py: load_task, load_model; the predictor trainer
    # loads its frozen autoencoder this way)
    from msmctts.tasks import build_task as bt, load_model
    ckpt = os.path.join(config.save_checkpoint_dir, "model_2")
    ae = load_model("autoencoder", ckpt)
    for k, v in trainer.model.autoencoder.state_dict().items():
        assert torch.equal(ae.state_dict()[k], v), k
    assert [n for n, _ in bt(checkpoint=ckpt, mode="infer").named_children()] == ["autoencoder", "discriminator"]


def test_train_cli_fails_loudly_without_a_gpu(tmp_path):
    """`python train.py -c cfg.yaml` (the reference's CLI) builds everything from the yaml and, on a machine
    without CUDA, stops at the first kernel call with MsmcError -- it never falls back to a CPU path"""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without CUDA")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "msmc-tts_b200", "train.py"), "-c", _cli_yaml(tmp_path)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0
    assert "MsmcError" in r.stderr and "no CPU fallback" in r.stderr


def test_batched_lsgan_loss_equals_sum_of_mse_terms():
    """trainer.lsgan_loss == sum_i F.mse_loss(s_i, target) (reference trainers/msmctts_trainer.py:165-168,184-185),
    values and gradients, on score tensors of different sizes and on slices of a shared buffer (the discriminator
    step scores cat(fake, real) and slices the halves).  Tolerance 1e-6 relative: summation order only."""
    import torch
    import torch.nn.functional as F
    from msmctts.trainers.msmctts_trainer import lsgan_loss
    torch.manual_seed(0)
    sizes = [(4, 1, 25, 13), (4, 120), (4, 1, 7), (4, 333), (4, 1, 2000, 2)]
    both = [torch.randn((8,) + s[1:], requires_grad=True) for s in sizes]
    ref_in = [b.detach().clone().requires_grad_(True) for b in both]
    for target in (1.0, 0.0):
        mine = lsgan_loss([b[:4] for b in both], target) + 0.5 * lsgan_loss([b[4:] for b in both], 1.0 - target)
        ref = sum(F.mse_loss(r[:4], torch.full_like(r[:4], target)) for r in ref_in) + \
            0.5 * sum(F.mse_loss(r[4:], torch.full_like(r[4:], 1.0 - target)) for r in ref_in)
        assert abs(float(mine) - float(ref)) <= 1e-6 * abs(float(ref))
        for t in both + ref_in:
            t.grad = None
        mine.backward()
        ref.backward()
        for a, b in zip(both, ref_in):
            assert torch.allclose(a.grad, b.grad, rtol=1e-5, atol=1e-9)


def _write_wav(path, x, sr=24000):
    import wave
    import numpy as np
    with wave.open(str(path), "wb") as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(sr)
        f.writeframes((np.clip(x, -1, 1) * 32767).astype("<i2").tobytes())


def test_file_backed_datasets_follow_the_reference_yaml_interface(tmp_path):
    """SURVEY 8f rank 4 data path: MelDataset / TTSDataset built from the reference yaml's own `dataset` keys
    (id_list, feature, feature_path templates and books, dimension, frameshift, padding_value, segment_length,
    pre_load): windows of mel and wav stay aligned, batches are sorted by length and padded with padding_value."""
    import numpy as np
    import torch
    from msmctts.datasets import build_dataloader
    from msmctts.utils.config import ConfigItem
    rng = np.random.default_rng(0)
    (tmp_path / "mel").mkdir()
    (tmp_path / "wav").mkdir()
    ids, lens = ["000001", "000002", "000003", "000004"], [57, 80, 41, 66]
    hop = 30
    for uid, n in zip(ids, lens):
        np.save(tmp_path / "mel" / (uid + ".npy"), rng.standard_normal((n, 8)).astype(np.float32))
        # sample k of the waveform encodes its own index so alignment can be checked
        _write_wav(tmp_path / "wav" / (uid + ".wav"), (np.arange(n * hop) % 900) / 2000.0)   # 900 = 30 hops
    (tmp_path / "train.list").write_text("\n".join(ids) + "\n")
    phones = {uid: rng.integers(1, 50, size=5 + i) for i, uid in enumerate(ids)}
    (tmp_path / "phone.txt").write_text("".join(
        "%s|%s\n" % (u, " ".join("%d_%d_%d" % (p, p % 7 + 1, p % 2) for p in ph)) for u, ph in phones.items()))
    durs = {}
    for uid, n in zip(ids, lens):
        k = len(phones[uid])
        d = np.full(k, n // k)
        d[-1] += n - d.sum()
        durs[uid] = d
    (tmp_path / "dur.txt").write_text("".join("%s|%s\n" % (u, " ".join(str(int(v)) for v in d))
                                              for u, d in durs.items()))
    # ---- MelDataset (examples/csmsc/configs/msmc_vq_gan.yaml: dataset block), random 20-frame segments
    cfg = ConfigItem({"_name": "MelDataset", "id_list": str(tmp_path / "train.list"), "samplerate": 24000,
                      "feature": ["mel", "wav"],
                      "feature_path": [str(tmp_path / "mel" / "{}.npy"), str(tmp_path / "wav" / "{}.wav")],
                      "dimension": [8, 1], "frameshift": [hop, 1], "padding_value": [-4, 0], "pre_load": False,
                      "segment_length": 20 * hop})
    ds, _, loader = build_dataloader(cfg, ConfigItem({"batch_size": 4, "num_workers": 0}))
    assert len(ds) >= 3200                       # an epoch is at least MIN_DATASET_SIZE draws, like the reference
    batch = next(iter(loader))
    assert batch["mel"].shape == (4, 20, 8) and batch["wav"].shape == (4, 20 * hop, 1)
    assert batch["mel_length"].tolist() == [20] * 4 and batch["wav_length"].tolist() == [20 * hop] * 4
    w = batch["wav"][:, :, 0] * 2000.0           # recovered sample indices: consecutive, starting at a hop multiple
    assert torch.allclose((w[:, 1:] - w[:, :-1]) % 900, torch.ones(4, 20 * hop - 1), atol=0.1)
    assert torch.allclose(torch.round(w[:, 0]) % hop, torch.zeros(4), atol=0.1)
    # whole utterances: sorted by length, padded
    cfg["segment_length"] = -1
    ds, _, loader = build_dataloader(cfg, ConfigItem({"batch_size": 4, "num_workers": 0}))
    batch = next(iter(loader))
    got = batch["mel_length"].tolist()           # (a shuffled epoch of >= 3200 draws may repeat an utterance)
    assert got == sorted(got, reverse=True) and set(got) <= set(lens)
    assert batch["mel"].shape == (4, got[0], 8) and batch["wav"].shape == (4, got[0] * hop, 1)
    for row, n in enumerate(got):
        if n < got[0]:
            assert float(batch["mel"][row, n:].max()) == -4.0 and float(batch["mel"][row, n:].min()) == -4.0
            assert float(batch["wav"][row, n * hop:].abs().max()) == 0
    # ---- TTSDataset (msmc_vq_gan_am.yaml: text and durations come from books)
    cfg = ConfigItem({"_name": "TTSDataset", "id_list": str(tmp_path / "train.list"), "samplerate": 24000,
                      "feature": ["text", "dur", "mel"],
                      "feature_path": [str(tmp_path / "phone.txt"), str(tmp_path / "dur.txt"),
                                       str(tmp_path / "mel" / "{}.npy")],
                      "dimension": [3, 1, 8], "padding_value": [0, 0, -4], "frameshift": [None, None, hop],
                      "pre_load": True, "segment_length": -1})
    ds, _, loader = build_dataloader(cfg, ConfigItem({"batch_size": 4, "num_workers": 0}))
    batch = next(iter(loader))
    tl = batch["text_length"].tolist()
    assert tl == sorted(tl, reverse=True) and set(tl) <= {5, 6, 7, 8}
    assert batch["text"].shape == (4, tl[0], 3) and batch["dur"].shape == (4, tl[0])
    assert batch["dur"].sum(1).tolist() == batch["mel_length"].tolist()       # durations add up to the mel frames
    for row, n in enumerate(tl):
        if n < tl[0]:
            assert float(batch["text"][row, n:].abs().max()) == 0 and float(batch["dur"][row, n:].abs().max()) == 0


def test_device_prefetcher_yields_every_batch_in_order():
    import torch
    from msmctts.datasets import DevicePrefetcher
    batches = [{"x": torch.full((2, 3), float(i)), "n": torch.tensor(i)} for i in range(5)]
    got = list(DevicePrefetcher(batches))        # no CUDA here: pass-through, same order, nothing dropped
    assert [int(b["n"]) for b in got] == list(range(5)) and all(float(b["x"][0, 0]) == i for i, b in enumerate(got))
