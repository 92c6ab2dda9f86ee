"""Generate tests/golden/*.pt from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py

Each fixture holds: the config, the reference module's state_dict (small configs only), seeded inputs, and the
reference's outputs (and, where noted, gradients).  tests/test_oracle_golden.py pins oracle/ against them;
the GPU parity tests compare the CUDA path with the same fixtures.  Harness-only tweaks (no reference source
is modified): the shims of oracle/ref_import.py, and ResStack's hard-wired Dropout(0.1) set to p=0 on the
instantiated module so train-mode forwards are deterministic.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_import as R  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

SMALL_AE = dict(
    in_dim=20, n_model_size=64,
    encoder_config=dict(downsample_scales=[1, 4], max_seq_len=100, n_layers=1, n_head=2, d_k=64, d_v=64,
                        d_inner=96, fft_conv1d_kernel=3, fft_conv1d_padding=1, dropout=0.0, attn_dropout=0.0,
                        fused_layernorm=False),
    quantizer_config=dict(embedding_sizes=32, embedding_dims=64, n_heads=2,
                          prior_config=dict(kernel_size=5, dilation_rate=1, n_layers=1), norm=False, dropout=0.0),
    frame_decoder_config=dict(max_seq_len=100, n_layers=1, n_head=2, d_k=64, d_v=64, d_inner=96,
                              fft_conv1d_kernel=3, fft_conv1d_padding=1, dropout=0.0, attn_dropout=0.0,
                              fused_layernorm=False),
    pred_mel=True,
    decoder_config=dict(upsample_rates=[3, 2], upsample_kernel_sizes=[6, 4], upsample_initial_channel=32,
                        resblock_kernel_sizes=[3, 5], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5]]),
)
SMALL_D = dict(
    mrd_config=dict(hop_lengths=[15, 30], hidden_channels=[32, 64], domain="double", mel_scale=True,
                    sample_rate=24000),
    mpd_config=dict(periods=[2, 3], channels=4, max_channels=16),
)


def cfgitem(d):
    C = R.ref("msmctts.utils.config")
    return C.ConfigItem(d)


def clone_sd(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def save(name, obj):
    path = os.path.join(OUT, name)
    torch.save(obj, path)
    print("%-28s %8.1f KB" % (name, os.path.getsize(path) / 1024))


def gen_quantize():
    M = R.ref("msmctts.networks.vqgantts.modules")
    torch.manual_seed(11)
    for name, heads, K in (("mh4_k64", 4, 64), ("mh4_k256", 4, 256), ("single_k64", 1, 64)):
        q = M.MultiHeadQuantize(256, K, heads) if heads > 1 else M.Quantize(256, K)
        # make the codebook "data-like": a couple of EMA steps first
        q.train()
        lengths = torch.tensor([24, 17, 9])
        for _ in range(2):
            q(torch.randn(3, 24, 256), lengths)
        sd0 = clone_sd(q)
        x = torch.randn(3, 24, 256, requires_grad=True)
        quant, diff, ind = q(x, lengths)
        (quant.sum() * 0.5 + (diff * torch.linspace(0, 1, diff.shape[-1])).sum()).backward()
        save("quantize_%s.pt" % name, dict(heads=heads, K=K, sd_before=sd0, sd_after=clone_sd(q), x=x.detach(),
                                           lengths=lengths, quant=quant.detach(), diff=diff.detach(), ind=ind,
                                           grad_x=x.grad.clone()))
    # triplet loss
    q = M.MultiHeadQuantize(256, 64, 4)
    pred = torch.randn(2, 10, 256, requires_grad=True)
    tgt = torch.randint(0, 64, (2, 10, 4))
    out = {}
    for red in ("mean", "sum"):
        l = q.compute_triple_loss(pred, tgt, reduction=red)
        (g,) = torch.autograd.grad(l.sum(), pred)
        out[red] = l.detach()
        out["grad_" + red] = g
    save("triple_loss.pt", dict(sd=clone_sd(q), pred=pred.detach(), target=tgt, **out))


def gen_fft():
    T = R.ref("msmctts.networks.acoustic_models.transformer")
    torch.manual_seed(12)
    cfg = dict(max_seq_len=100, n_layers=2, n_head=2, d_k=64, d_v=64, d_model=64, d_inner=128,
               fft_conv1d_kernel=3, fft_conv1d_padding=1, dropout=0.0, attn_dropout=0.0, name="t")
    m = T.FFTBlocks(**cfg)
    m.train()
    for p in m.parameters():
        if p.dim() == 1:
            p.data.add_(0.1 * torch.randn_like(p))
    seq = torch.randn(3, 37, 64, requires_grad=True)
    lengths = torch.tensor([37, 20, 5])
    pos = torch.arange(1, 38).view(1, -1).repeat(3, 1)
    pos.masked_fill_(torch.arange(37).view(1, -1) >= lengths.view(-1, 1), 0)
    out, _ = m(seq, pos)
    w = torch.randn_like(out)
    (out * w).sum().backward()
    grads = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    save("fftblocks.pt", dict(cfg=cfg, sd=clone_sd(m), seq=seq.detach(), lengths=lengths, pos=pos, out=out.detach(),
                              w=w, grad_seq=seq.grad.clone(), grads=grads))


def gen_generator():
    G = R.ref("msmctts.networks.hifigan.generator")
    torch.manual_seed(13)
    dcfg = dict(SMALL_AE["decoder_config"], num_mels=16)
    m = G.Generator(**dcfg)
    for p in m.parameters():   # default init std 0.01 makes everything tiny; widen so errors are visible
        p.data.mul_(3.0).add_(0.02 * torch.randn_like(p))
    x = torch.randn(2, 16, 11, requires_grad=True)
    y = m(x)
    w = torch.randn_like(y)
    (y * w).sum().backward()
    grads = {k: p.grad.clone() for k, p in m.named_parameters()}
    save("generator.pt", dict(cfg=dcfg, sd=clone_sd(m), x=x.detach(), y=y.detach(), w=w, grad_x=x.grad.clone(),
                              grads=grads))


def gen_discriminator():
    D = R.ref("msmctts.networks.hifigan.discriminator")
    torch.manual_seed(14)
    m = D.Discriminator(cfgitem(SMALL_D["mrd_config"]), cfgitem(SMALL_D["mpd_config"]))
    y = (0.3 * torch.randn(2, 1201)).clamp(-1, 1).requires_grad_(True)   # odd length exercises the period padding
    scores, feats = m(y)
    loss = sum((s * torch.linspace(-1, 1, s.numel()).view_as(s)).sum() for s in scores) + \
        sum(f.abs().mean() for fl in feats for f in fl)
    loss.backward()
    grads = {k: p.grad.clone() for k, p in m.named_parameters()}
    save("discriminator.pt", dict(cfg=SMALL_D, sd=clone_sd(m), y=y.detach(), scores=[s.detach() for s in scores],
                                  feats=[[f.detach() for f in fl] for fl in feats], loss=loss.detach(),
                                  grad_y=y.grad.clone(), grads=grads))
    # spectral front end alone (utils/audio.py TorchSTFT.transform) at two of the CSMSC resolutions
    A = R.ref("msmctts.utils.audio")
    fe = {}
    x = (0.3 * torch.randn(2, 2400)).clamp(-1, 1)
    for hop in (15, 120):
        st = A.TorchSTFT(fft_size=hop * 4, hop_size=hop, win_size=hop * 4, normalized=True, domain="double",
                         mel_scale=True, sample_rate=24000)
        mag, _ = st.transform(x)
        fe["hop%d" % hop] = mag
        fe["fb%d" % hop] = A.create_fb_matrix(hop * 2 + 1, 0.0, 12000.0, hop * 2 + 1, 24000)
    save("stft_frontend.pt", dict(x=x, **fe))


def gen_melloss():
    S = R.ref("msmctts.trainers.criterions.stft_loss")
    torch.manual_seed(15)
    ml = S.MelLoss(2048, 300, 1200, 24000, 128)
    a = (0.3 * torch.randn(2, 3000)).clamp(-1, 1).requires_grad_(True)
    b = (0.3 * torch.randn(2, 3000)).clamp(-1, 1)
    l = ml(a, b)
    l.backward()
    save("melloss.pt", dict(pred=a.detach(), target=b, loss=l.detach(), grad_pred=a.grad.clone(),
                            note="mel filterbank = oracle restatement of librosa (parity unpinned)"))


def gen_autoencoder():
    V = R.ref("msmctts.networks.vqgantts.msmc_vqgan")
    torch.manual_seed(16)
    cfg = json.loads(json.dumps(SMALL_AE))
    m = V.MSMCVQGAN(cfg["in_dim"], cfg["n_model_size"], cfgitem(cfg["encoder_config"]),
                    cfgitem(cfg["quantizer_config"]), cfgitem(cfg["frame_decoder_config"]),
                    cfgitem(cfg["decoder_config"]), cfg["pred_mel"])
    for k, p in m.named_parameters():
        if k.startswith("decoder."):
            p.data.mul_(3.0).add_(0.02 * torch.randn_like(p))
    for pr in m.quantizer.predictor:   # harness tweak: ResStack's hard-wired Dropout(0.1) -> 0
        pr.enc.drop.p = 0.0
    m.train()
    mel = (1.5 * torch.randn(2, 24, 20)).clamp(-4, 4)
    length = torch.tensor([24, 18])
    with torch.no_grad():
        for _ in range(2):
            m(mel, length, warmup=True)   # settle the codebooks with two EMA steps
    sd0 = clone_sd(m)
    window = [(4, 12), (2, 10)]
    out = m(mel, length, warmup=False, window=window)
    loss = out["decoder_outputs"].pow(2).mean() * 10 + out["mel_outputs"].pow(2).mean() + \
        sum(d.mean() for d in out["encoder_diffs"]) + out["decoder_diffs"]["total_loss"]
    loss.backward()
    grads = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    save("autoencoder_train.pt", dict(
        cfg=SMALL_AE, sd_before=sd0,
        sd_after={k: v for k, v in clone_sd(m).items() if k.split('.')[-1] in ('embed', 'embed_avg', 'cluster_size')}, mel=mel, length=length, window=window,
        decoder_outputs=out["decoder_outputs"].detach(), mel_outputs=out["mel_outputs"].detach(),
        encoder_indices=[i for i in out["encoder_indices"]], encoder_diffs=[d.detach() for d in out["encoder_diffs"]],
        decoder_total=out["decoder_diffs"]["total_loss"].detach(), loss=loss.detach(), grads=grads))
    # eval-mode analysis-synthesis (config 1 call path: tasks/msmc_tts.py:129-133), single stage / single head
    cfg1 = json.loads(json.dumps(SMALL_AE))
    cfg1["encoder_config"]["downsample_scales"] = [1]
    cfg1["quantizer_config"].update(n_heads=1, embedding_sizes=64, embedding_dims=256)
    cfg1["in_dim"] = 80
    torch.manual_seed(17)
    m1 = V.MSMCVQGAN(cfg1["in_dim"], cfg1["n_model_size"], cfgitem(cfg1["encoder_config"]),
                     cfgitem(cfg1["quantizer_config"]), cfgitem(cfg1["frame_decoder_config"]),
                     cfgitem(cfg1["decoder_config"]), cfg1["pred_mel"])
    for k, p in m1.named_parameters():
        if k.startswith("decoder."):
            p.data.mul_(3.0).add_(0.02 * torch.randn_like(p))
    m1.eval()
    mel1 = (1.5 * torch.randn(2, 64, 80)).clamp(-4, 4)
    len1 = torch.tensor([64, 64])
    with torch.no_grad():
        o1 = m1(mel1, len1)
    save("autoencoder_config1.pt", dict(cfg=cfg1, sd=clone_sd(m1), mel=mel1, length=len1,
                                        decoder_outputs=o1["decoder_outputs"], mel_outputs=o1["mel_outputs"],
                                        encoder_indices=[i for i in o1["encoder_indices"]]))


def gen_predictor():
    P = R.ref("msmctts.networks.acoustic_models.multi_stage_predictor")
    torch.manual_seed(18)
    fft = dict(n_layers=1, n_head=2, d_k=64, d_v=64, d_model=64, d_inner=96, fft_conv1d_kernel=3,
               fft_conv1d_padding=1, dropout=0.0, attn_dropout=0.0, fused_layernorm=False)
    cfg = dict(n_symbols=[20, 5, 2], n_model_size=64, n_pred_size=32, n_pred_scale=[4, 1],
               encoder_config=dict(fft, max_seq_len=40, name="phoneme_side"),
               adaptor_config=dict(input_size=64, duration_predictor_filter_size=48,
                                   duration_predictor_kernel_size=3, dropout=0.0, fused_layernorm=False),
               decoder_config=dict(fft, max_seq_len=200, name="mel_side"))
    m = P.MultiStagePredictor(**json.loads(json.dumps(cfg)))
    m.train()
    B, Lt = 2, 6
    text = torch.stack([torch.randint(1, 20, (B, Lt)), torch.randint(1, 5, (B, Lt)), torch.randint(0, 2, (B, Lt))], -1)
    text_length = torch.tensor([6, 4])
    text[1, 4:] = 0
    dur = torch.randint(2, 7, (B, Lt)).float()
    dur[1, 4:] = 0
    T = int(dur.sum(1).max())
    fl1 = dur.sum(1).long()
    fl0 = torch.ceil(fl1 / 4).long()
    feat = [torch.randn(B, int(fl0.max()), 32), torch.randn(B, T, 32)]
    out = m(text, text_length, dur=dur, feat=feat, feat_length=[fl0, fl1])
    loss = sum((p * torch.linspace(-1, 1, p.numel()).view_as(p)).sum() for p in out["feat"]) + out["duration"].sum()
    loss.backward()
    grads = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    save("predictor.pt", dict(cfg=cfg, sd=clone_sd(m), text=text, text_length=text_length, dur=dur, feat=feat,
                              feat_length=[fl0, fl1], preds=[p.detach() for p in out["feat"]],
                              duration=out["duration"].detach(), loss=loss.detach(), grads=grads))


def gen_train_step():
    """Two consecutive `VQGANTrainer.train_step` calls of the UNMODIFIED reference trainer
    (trainers/msmctts_trainer.py:115-209: losses, D step, G step, clip, AdamW through the reference's own
    build_optimizer) on CPU, small config.  Pins oracle/train_step.py -- the port that the whole-step GPU parity test
    and bench.py's CPU arm run.  The window draw (random.randrange, :211-219) is made reproducible by seeding
    Python's RNG before each step; the fixture records the windows it produced."""
    import random
    V = R.ref("msmctts.networks.vqgantts.msmc_vqgan")
    D = R.ref("msmctts.networks.hifigan.discriminator")
    T = R.ref("msmctts.trainers.msmctts_trainer")
    O = R.ref("msmctts.trainers.optimizers")
    torch.manual_seed(21)
    cfg = json.loads(json.dumps(SMALL_AE))
    cfg["decoder_config"].update(upsample_rates=[10, 2], upsample_kernel_sizes=[20, 4])    # 20 samples per frame
    ae = V.MSMCVQGAN(cfg["in_dim"], cfg["n_model_size"], cfgitem(cfg["encoder_config"]),
                     cfgitem(cfg["quantizer_config"]), cfgitem(cfg["frame_decoder_config"]),
                     cfgitem(cfg["decoder_config"]), cfg["pred_mel"])
    for k, p in ae.named_parameters():
        if k.startswith("decoder."):
            p.data.mul_(3.0).add_(0.02 * torch.randn_like(p))
    for pr in ae.quantizer.predictor:   # harness tweak: ResStack's hard-wired Dropout(0.1) -> 0
        pr.enc.drop.p = 0.0
    disc = D.Discriminator(cfgitem(SMALL_D["mrd_config"]), cfgitem(SMALL_D["mpd_config"]))
    model = torch.nn.Module()
    model.autoencoder, model.discriminator = ae, disc
    model.train()
    tcfg = dict(warmup_steps=0, lambda_frame=450, grad_clip_thresh=1.0, sample_lengths=1200, lambda_vq=1,
                lambda_pr=0.1, lambda_fm=2, lambda_stft=45)
    ocfg = dict(_name="AdamW", learning_rate=0.0002, betas=[0.8, 0.99], eps=1e-8, weight_decay=0.0)
    config = cfgitem(dict(dataset=dict(samplerate=24000, feature=["mel", "wav"], frameshift=[20, 1]),
                          optimizer=dict(_default=ocfg)))
    trainer = T.VQGANTrainer(config, model, num_gpus=0, rank=0, **tcfg)
    trainer.optimizer = O.build_optimizer(model, config.optimizer)
    B, frames = 2, 80
    mel = (1.5 * torch.randn(B, frames, cfg["in_dim"])).clamp(-4, 4)
    wav = (0.3 * torch.randn(B, frames * 20, 1)).clamp(-1, 1)
    length = torch.tensor([80, 71])
    batch = dict(mel=mel, mel_length=length, wav=wav, wav_length=length * 20)
    sd_ae0, sd_d0 = clone_sd(ae), clone_sd(disc)
    steps = []
    for it in (1, 2):
        random.seed(100 + it)
        windows = [(s, s + 60) for s in (random.randrange(max(1, int(n) - 60)) for n in length)]
        random.seed(100 + it)
        log = trainer.train_step(batch, iteration=it)["loss"]
        steps.append(dict(windows=windows, losses={k: float(v) for k, v in log.items()}))
    # after two steps: every 4th tensor plus all codebook buffers (keeps the fixture small)
    keep = lambda sd: {k: v for i, (k, v) in enumerate(sd.items())
                       if i % 4 == 0 or k.split(".")[-1] in ("embed", "embed_avg", "cluster_size")}
    save("train_step.pt", dict(cfg=dict(autoencoder=cfg, discriminator=SMALL_D), trainer=dict(tcfg, frameshift=20,
                               sample_rate=24000), optimizer=ocfg, sd_ae=sd_ae0, sd_d=sd_d0, mel=mel, wav=wav,
                               length=length, steps=steps, sd_ae_after=keep(clone_sd(ae)),
                               sd_d_after=keep(clone_sd(disc))))


def gen_keys():
    """state_dict names + shapes of the full CSMSC models (examples/csmsc/configs/msmc_vq_gan.yaml)."""
    C = R.ref("msmctts.utils.config")
    cfg = C.Config(os.path.join(R.REF_ROOT, "examples/csmsc/configs/msmc_vq_gan.yaml"))
    V = R.ref("msmctts.networks.vqgantts.msmc_vqgan")
    D = R.ref("msmctts.networks.hifigan.discriminator")
    ae_cfg = {k: v for k, v in cfg.task.autoencoder.items() if not k.startswith("_")}
    d_cfg = {k: v for k, v in cfg.task.discriminator.items() if not k.startswith("_")}
    ae = V.MSMCVQGAN(**ae_cfg)
    dd = D.Discriminator(**d_cfg)
    out = dict(autoencoder={k: list(v.shape) for k, v in ae.state_dict().items()},
               discriminator={k: list(v.shape) for k, v in dd.state_dict().items()},
               autoencoder_params=sum(p.numel() for p in ae.parameters()),
               discriminator_params=sum(p.numel() for p in dd.parameters()))
    with open(os.path.join(OUT, "csmsc_state_dict_keys.json"), "w") as f:
        json.dump(out, f)
    ae_cfg_plain = {k: v for k, v in cfg.task.autoencoder.to_dict().items() if not k.startswith('_')}
    d_cfg_plain = {k: v for k, v in cfg.task.discriminator.to_dict().items() if not k.startswith('_')}
    # the yaml blocks themselves (examples/csmsc/configs/msmc_vq_gan.yaml:10-67, 89-98), for tests and bench.py
    with open(os.path.join(OUT, "csmsc_config.json"), "w") as f:
        json.dump(dict(autoencoder=ae_cfg_plain, discriminator=d_cfg_plain,
                       trainer=cfg.trainer.to_dict(), optimizer=cfg.optimizer.to_dict(),
                       dataset=dict(samplerate=cfg.dataset.samplerate, frameshift=cfg.dataset.frameshift,
                                    feature=cfg.dataset.feature, padding_value=cfg.dataset.padding_value)),
                  f, indent=1)
    print("csmsc keys: ae %d tensors %.2fM params, d %d tensors %.2fM params" % (
        len(out["autoencoder"]), out["autoencoder_params"] / 1e6, len(out["discriminator"]),
        out["discriminator_params"] / 1e6))


def gen_am_config():
    """BASELINE.json configs[4]: the yaml blocks of examples/csmsc/configs/msmc_vq_gan_am.yaml (predictor, trainer,
    optimizer) for tests and bench.py --config am, plus the predictor's state_dict names / shapes."""
    C = R.ref("msmctts.utils.config")
    cfg = C.Config(os.path.join(R.REF_ROOT, "examples/csmsc/configs/msmc_vq_gan_am.yaml"))
    P = R.ref("msmctts.networks.acoustic_models.multi_stage_predictor")
    p_cfg = {k: v for k, v in cfg.task.predictor.to_dict().items() if not k.startswith("_")}
    m = P.MultiStagePredictor(**{k: v for k, v in cfg.task.predictor.items() if not k.startswith("_")})
    with open(os.path.join(OUT, "csmsc_am_config.json"), "w") as f:
        json.dump(dict(predictor=p_cfg, trainer=cfg.trainer.to_dict(), optimizer=cfg.optimizer.to_dict(),
                       predictor_keys={k: list(v.shape) for k, v in m.state_dict().items()},
                       predictor_params=sum(p.numel() for p in m.parameters())), f, indent=1)
    print("csmsc am: %d tensors %.2fM params" % (len(m.state_dict()), sum(p.numel() for p in m.parameters()) / 1e6))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    R.install()
    gen_quantize()
    gen_fft()
    gen_generator()
    gen_discriminator()
    gen_melloss()
    gen_autoencoder()
    gen_predictor()
    gen_train_step()
    gen_keys()
    gen_am_config()
