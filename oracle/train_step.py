"""CPU restatement of one full VQGANTrainer.train_step (reference trainers/msmctts_trainer.py:115-209) on top of
oracle/ref_modules.py: forward, D step, G step, clip, AdamW.  TEST INFRASTRUCTURE / CPU BASELINE ONLY.

Used by bench.py (`cpu_baseline`, `--impl reference`: the reference tree is absent on the GPU box and its
discriminator / MelLoss do not run unmodified on torch>=2 -- SURVEY section 0, B4/B5 -- so the timed CPU arm is this
port, kind "port") and by tests/ for whole-step parity.
"""
import torch
import torch.nn.functional as F

from . import ref_modules as O


class OracleTrainer(object):
    def __init__(self, sd_ae, sd_d, cfg, tcfg, ocfg, use_dropout=True):
        """sd_* : state dicts keyed like the reference's; cfg = {'autoencoder':..., 'discriminator':...};
        tcfg = trainer block; ocfg = optimizer._default block"""
        self.cfg, self.tcfg, self.use_dropout = cfg, tcfg, use_dropout
        self.sd_ae = {k: v.detach().clone() for k, v in sd_ae.items()}
        self.sd_d = {k: v.detach().clone() for k, v in sd_d.items()}
        buffers = ("embed", "embed_avg", "cluster_size")
        self.p_ae = [k for k, v in self.sd_ae.items() if v.is_floating_point() and k.split(".")[-1] not in buffers
                     and not k.endswith("position.weight")]
        self.p_d = [k for k, v in self.sd_d.items() if v.is_floating_point()]
        for k in self.p_ae:
            self.sd_ae[k].requires_grad_(True)
        for k in self.p_d:
            self.sd_d[k].requires_grad_(True)
        kw = dict(lr=ocfg["learning_rate"], betas=tuple(ocfg["betas"]), eps=ocfg["eps"],
                  weight_decay=ocfg["weight_decay"])
        self.opt_ae = torch.optim.AdamW([self.sd_ae[k] for k in self.p_ae], **kw)
        self.opt_d = torch.optim.AdamW([self.sd_d[k] for k in self.p_d], **kw)

    def step(self, mel, mel_length, wav, windows):
        t = self.tcfg
        losses, g_loss, predict, target, _ = O.generator_losses(
            self.sd_ae, self.sd_d, self.cfg, mel, mel_length, wav, windows,
            dict(dict(frameshift=300, sample_rate=24000), **t), training=True, use_dropout=self.use_dropout)
        self.opt_d.zero_grad(set_to_none=True)
        losses["d_loss"].backward()
        self.opt_d.step()
        adv = O.generator_adv_losses(self.sd_d, self.cfg, predict, target, g_loss, t.get("lambda_fm", 2))
        losses.update(adv)
        self.opt_ae.zero_grad(set_to_none=True)
        for k in self.p_d:
            self.sd_d[k].grad = None
        adv["g_loss"].backward()
        torch.nn.utils.clip_grad_norm_([self.sd_ae[k] for k in self.p_ae], t.get("grad_clip_thresh", 1.0))
        self.opt_ae.step()
        return {k: float(v) for k, v in losses.items() if torch.is_tensor(v)}


class OraclePredictorTrainer(object):
    """CPU restatement of PredictorTrainer.train_step (reference trainers/msmctts_trainer.py:237-286): frozen
    autoencoder analysis (eval mode), teacher-forced MultiStagePredictor, embedding loss (mse + triple_sum) and
    duration loss, clip_grad_norm_, Adam.  TEST INFRASTRUCTURE / CPU BASELINE ONLY (dropout off)."""

    def __init__(self, sd_p, sd_ae, pcfg, ae_cfg, tcfg, ocfg):
        self.pcfg, self.ae_cfg, self.tcfg = pcfg, ae_cfg, tcfg
        self.sd_ae = {k: v.detach().clone() for k, v in sd_ae.items()}
        self.sd_p = {k: v.detach().clone() for k, v in sd_p.items()}
        self.params = [k for k, v in self.sd_p.items() if v.is_floating_point() and not k.endswith("position.weight")]
        for k in self.params:
            self.sd_p[k].requires_grad_(True)
        kw = dict(lr=ocfg["learning_rate"], betas=tuple(ocfg["betas"]), eps=ocfg["eps"],
                  weight_decay=ocfg["weight_decay"])
        opt = torch.optim.AdamW if ocfg.get("_name", "Adam") == "AdamW" else torch.optim.Adam
        self.opt = opt([self.sd_p[k] for k in self.params], **kw)

    def losses(self, text, text_length, dur, mel, mel_length):
        t = self.tcfg
        with torch.no_grad():
            qs = O.msmcvqgan_forward(self.sd_ae, self.ae_cfg, mel, mel_length, warmup=True, training=False)
        feat, feat_length = qs["quantizer_outputs"], qs["encoder_lengths"]
        preds, duration = O.multistage_predictor(self.sd_p, self.pcfg, text, text_length, dur, feat, feat_length,
                                                 training=True)
        states = [dict(predictor_outputs=preds[i], target_outputs=feat[i], target_indices=qs["encoder_indices"][i],
                       target_lengths=feat_length[i]) for i in range(len(preds))]
        n_heads = self.ae_cfg["quantizer_config"].get("n_heads", 4)
        emb = O.embedding_loss(self.sd_ae, "quantizer.", states, n_heads, tuple(t.get("training_methods", ["mse"])),
                               t.get("loss_weights", [1.0]))
        losses = {"total_loss": emb.pop("total_loss")}
        losses.update(emb)
        dl = F.mse_loss(duration, dur.float(), reduction="none")
        dl = dl.masked_fill(O.mask_from_lengths(text_length, dl.shape[1]), 0)
        dl = dl.sum() / text_length.sum()
        losses["dur_loss"] = dl
        losses["total_loss"] = losses["total_loss"] + t.get("lambda_dur", 1.0) * dl
        return losses, preds, duration

    def step(self, text, text_length, dur, mel, mel_length):
        losses, _, _ = self.losses(text, text_length, dur, mel, mel_length)
        self.opt.zero_grad(set_to_none=True)
        losses["total_loss"].backward()
        clip = self.tcfg.get("grad_clip_thresh", 1.0)
        if clip is not None:
            losses["grad_norm"] = torch.nn.utils.clip_grad_norm_([self.sd_p[k] for k in self.params], clip)
        self.opt.step()
        return {k: float(v) for k, v in losses.items() if torch.is_tensor(v)}
