"""GAN / predictor train steps driving the sm_100a network kernels (reference trainers/msmctts_trainer.py:12-294).

Same losses and update order as VQGANTrainer.train_step (:115-209): VQ + frame + MelLoss, discriminator step on
(fake.detach(), real), then generator step with adversarial + feature-matching terms, gradient clipping, AdamW.
Host-side differences only: no `.item()` host syncs inside the step (the reference has 8 explicit ones plus ~265
implicit) -- losses are returned as 0-dim device tensors and read when they are logged; the window
gather is tensor arithmetic instead of per-sample Python slicing, so the step is CUDA-graph capturable.
"""
import contextlib
import random

import torch
import torch.nn as nn
import torch.nn.functional as F

from msmctts.tasks import load_model
from msmctts.utils.utils import get_mask_from_lengths
from .base_trainer import BaseTrainer
from .criterions.stft_loss import MelLoss


class DurationLoss(nn.Module):
    def __init__(self, lambda_dur=1):
        super().__init__()
        self.lambda_dur = lambda_dur

    def forward(self, outputs, targets):
        dur_target = targets["dur"].float()
        dur_length = targets["text_length"]
        loss = F.mse_loss(outputs["duration"], dur_target, reduction="none")
        loss = loss.masked_fill(get_mask_from_lengths(dur_length, loss.shape[1]), 0)
        dur_loss = loss.sum() / dur_length.sum()
        return {"total_loss": self.lambda_dur * dur_loss, "dur_loss": dur_loss}


class QuantizerLoss(nn.Module):
    """masked commitment loss per stage + weighted prior-prediction loss (reference :39-71)"""

    def __init__(self, lambda_vq=1, lambda_pr=1):
        super().__init__()
        self.lambda_vq, self.lambda_pr = lambda_vq, lambda_pr

    def forward(self, outputs):
        loss = {"vq_loss": 0}
        latent_losses = outputs["encoder_diffs"]
        if not isinstance(latent_losses, (tuple, list)):
            latent_losses = [latent_losses]
        for i, term in enumerate(latent_losses):
            length = outputs["encoder_lengths"][i]
            mask = get_mask_from_lengths(length, term.shape[1])
            term = term.masked_fill(mask.unsqueeze(-1), 0)
            term = term.sum() / length.sum() / term.shape[2]
            loss["latent_loss_{}_{}".format(i, 0)] = term
            loss["vq_loss"] = loss["vq_loss"] + self.lambda_vq * term
        dd = outputs.get("decoder_diffs")
        if isinstance(dd, dict):
            dd = dict(dd)
            loss["vq_loss"] = loss["vq_loss"] + self.lambda_pr * dd.pop("total_loss")
            loss.update(dd)
        return loss


_LSGAN_WEIGHTS = {}


def lsgan_loss(scores, target):
    """sum_i F.mse_loss(s_i, full_like(s_i, target)) over a list of score tensors (reference :165-168, :184-185) in
    one pass: the scores are concatenated and each element carries the weight 1 / numel(s_i).  ~7 tiny launches per
    score (fill, sub, pow, mean and their backward) become ~10 per LOSS; they sit on the serial stretch between the
    discriminator's forward and backward."""
    scores = list(scores)
    key = (tuple(int(s.numel()) for s in scores), scores[0].device)
    w = _LSGAN_WEIGHTS.get(key)
    if w is None:
        if scores[0].is_cuda and torch.cuda.is_current_stream_capturing():
            # never cache memory that belongs to a CUDA graph's private pool: per-score form for this one call
            return sum(F.mse_loss(s, torch.full_like(s, float(target))) for s in scores)
        w = torch.cat([torch.full((n,), 1.0 / n, dtype=torch.float32, device=key[1]) for n in key[0]])
        _LSGAN_WEIGHTS[key] = w
    flat = torch.cat([s.reshape(-1) for s in scores])
    d = flat - float(target)
    return (d * d * w).sum()


@contextlib.contextmanager
def _frozen(module):
    """parameters of `module` do not require grad inside the block (custom autograd Functions decide at forward time
    which gradients they will produce, so restricting backward(inputs=...) alone would not skip the weight gradients)"""
    params = [p for p in module.parameters() if p.requires_grad]
    for p in params:
        p.requires_grad_(False)
    try:
        yield
    finally:
        for p in params:
            p.requires_grad_(True)


class VQGANTrainer(BaseTrainer):
    def __init__(self, config, model, num_gpus=1, rank=0, warmup_steps=0, lambda_frame=1.0,
                 eval_inteval_iters=1000, grad_clip_thresh=1.0, sample_lengths=24000, lambda_vq=1, lambda_pr=1,
                 lambda_fm=2, lambda_stft=45, stft_loss_func="mel_loss", stft_loss_config=None, cuda_graph=None,
                 reference_schedule=False, max_graphs=4, graph_bucket_frames=0):
        super().__init__(config, model, num_gpus, rank)
        # reference_schedule=False (default) keeps every loss value and every parameter update of the reference's
        # step but drops work whose result the reference discards:
        #   * the discriminator step scores cat(fake, real) in ONE pass (no batch statistics anywhere in D, so the
        #     per-sample math is unchanged; the two weight-gradient sums become one);
        #   * the generator step back-propagates only into the autoencoder: the reference also fills D's .grad
        #     there (msmctts_trainer.py:182-207) and zeroes it before it is ever read (:166), and the real branch of
        #     the feature-matching pass needs no graph at all.
        # reference_schedule=True replays the reference's exact launch schedule (4 separate D passes, D gradients
        # computed in the G step).
        self.reference_schedule = bool(reference_schedule)
        # cuda_graph: the sync-free step body is captured once per (shape, phase) into a CUDA graph and replayed --
        # ~2k kernel launches and all Python/autograd dispatch collapse into one graph launch.  Default (None) = on
        # whenever the model lives on a CUDA device, so `train.py -c <reference yaml>` (no such key) gets the graphed
        # step.  At most `max_graphs` shapes are captured (each holds its own activation pool); further shapes run the
        # eager step.  graph_bucket_frames > 0 pads every batch on the right (mel with the dataset's pad value, wav
        # with zeros; lengths unchanged, so every masked quantity is unaffected) to a multiple of that many frames so
        # that variable-length batches share a few graphs.
        self.use_cuda_graph = bool(next(model.parameters()).is_cuda) if cuda_graph is None else bool(cuda_graph)
        self.max_graphs, self.graph_bucket_frames = int(max_graphs), int(graph_bucket_frames)
        self._graphs = {}
        self.lambda_frame, self.warmup_steps = lambda_frame, warmup_steps
        self.frameshift = self.config.dataset.frameshift[self.config.dataset.feature.index("mel")]
        self.frame_lengths = -1 if sample_lengths == -1 else sample_lengths // self.frameshift
        self.grad_clip_thresh = grad_clip_thresh
        self.vq_criterion = QuantizerLoss(lambda_vq=lambda_vq, lambda_pr=lambda_pr)
        self.sample_lengths, self.lambda_fm, self.lambda_stft = sample_lengths, lambda_fm, lambda_stft
        if stft_loss_func != "mel_loss":
            raise NotImplementedError("only the default 'mel_loss' criterion is built on the B200 path")
        sr = config.dataset.samplerate
        kwargs = dict(sample_rate=sr, win_size=sr // 20, hop_size=sr // 80, num_mels=128)
        kwargs["fft_size"] = 2048 if kwargs["win_size"] > 1024 else 1024
        if stft_loss_config is not None:
            kwargs.update(stft_loss_config)
        self.stft_criterion = MelLoss(**kwargs)
        if next(model.parameters()).is_cuda:
            self.stft_criterion = self.stft_criterion.cuda()

    # ------------------------------------------------------------------ window selection (reference :211-219)
    def random_select(self, mel_length):
        lengths = mel_length.tolist() if torch.is_tensor(mel_length) else list(mel_length)
        frame_windows, sample_windows = [], []
        for n in lengths:
            start = random.randrange(max(1, int(n) - self.frame_lengths))
            end = start + self.frame_lengths
            frame_windows.append((start, end))
            sample_windows.append((start * self.frameshift, end * self.frameshift))
        return frame_windows, sample_windows

    @staticmethod
    def gather_windows(x, starts, length):
        """x (B, T, ...), starts (B,) device int64 -> (B, length, ...) without host-side slicing"""
        idx = starts.view(-1, 1) + torch.arange(length, device=x.device).view(1, -1)
        idx = idx.view(idx.shape + (1,) * (x.dim() - 2)).expand(-1, -1, *x.shape[2:])
        return torch.gather(x, 1, idx)

    # ------------------------------------------------------------------------------------------ the step
    def train_step(self, batch, iteration, frame_windows=None):
        mel, mel_length = batch["mel"], batch["mel_length"]
        wav = batch["wav"]
        step = self._graphed_step if (self.use_cuda_graph and mel.is_cuda) else self._step
        if iteration < self.warmup_steps:
            return step(mel, mel_length, None, None, warmup=True, gan=False)
        if frame_windows is None:
            frame_windows, _ = self.random_select(mel_length.cpu())
        starts = torch.as_tensor([w[0] for w in frame_windows], dtype=torch.int64)
        if mel.is_cuda:
            starts = starts.pin_memory().to(mel.device, non_blocking=True)
        return step(mel, mel_length, wav, starts, warmup=False, gan=iteration > self.warmup_steps)

    def _graphed_step(self, mel, mel_length, wav, starts, warmup, gan):
        """CUDA-graph replay of `_step` on static input buffers (one graph per input shape and phase)."""
        if self.graph_bucket_frames > 0:
            T, Tb = mel.shape[1], -(-mel.shape[1] // self.graph_bucket_frames) * self.graph_bucket_frames
            if Tb != T:
                pad_value = -4.0
                pv = self.config.dataset.get("padding_value", None)
                if isinstance(pv, (list, tuple)):       # one value per feature (reference yaml: [-4, 0])
                    pad_value = float(pv[list(self.config.dataset.feature).index("mel")])
                mel = F.pad(mel, (0, 0, 0, Tb - T), value=pad_value)
                if wav is not None:
                    wav = F.pad(wav, (0, 0, 0, (Tb - T) * self.frameshift) if wav.dim() == 3
                                else (0, (Tb - T) * self.frameshift))
        key = (tuple(mel.shape), None if wav is None else tuple(wav.shape), warmup, gan)
        st = self._graphs.get(key)
        if st is None and len(self._graphs) >= self.max_graphs:
            return self._step(mel, mel_length, wav, starts, warmup=warmup, gan=gan)
        if st is None:
            st = {"n": 0, "inp": [None if t is None else torch.empty_like(t) for t in (mel, mel_length, wav, starts)]}
            self._graphs[key] = st
        for dst, src in zip(st["inp"], (mel, mel_length, wav, starts)):
            if dst is not None:
                dst.copy_(src, non_blocking=True)
        if st["n"] < 3:
            # eager warm-up on a side stream (initialises optimizer state, weight-norm caches, autotuned attributes)
            st["n"] += 1
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                out = self._step(*st["inp"], warmup=warmup, gan=gan)
            torch.cuda.current_stream().wait_stream(side)
            return out
        if "graph" not in st:
            self.optimizer.zero_grad()
            graph = torch.cuda.CUDAGraph()
            from msmctts._b200.functional import MAIN_PRIORITY
            # thread_local: the DataLoader's pin-memory thread (cudaHostAlloc) and NCCL's watchdog keep making CUDA
            # calls from THEIR threads while this one captures; in the default "global" mode those invalidate the capture
            with torch.cuda.graph(graph, stream=torch.cuda.Stream(priority=MAIN_PRIORITY),
                                  capture_error_mode="thread_local"):
                st["out"] = self._step(*st["inp"], warmup=warmup, gan=gan)
            st["graph"] = graph
        st["graph"].replay()
        return st["out"]

    def release_graphs(self):
        """drop the captured graphs (and their private memory pools).  Must run before the NCCL communicator is
        destroyed when the graphs hold captured all-reduces: destroy_process_group() otherwise never returns."""
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        for st in self._graphs.values():
            st.pop("graph", None)
            st.pop("out", None)
        self._graphs = {}
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def _step(self, mel, mel_length, wav, starts, warmup, gan):
        if mel.is_cuda:
            from msmctts._b200.functional import prep_scope
            with prep_scope():
                return self._step_body(mel, mel_length, wav, starts, warmup, gan)
        return self._step_body(mel, mel_length, wav, starts, warmup, gan)

    def _step_body(self, mel, mel_length, wav, starts, warmup, gan):
        """sync-free body; everything inside is device work (graph-capturable)"""
        losses = {}
        ae, disc = self.model.autoencoder, getattr(self.model, "discriminator", None)
        if mel.is_cuda:
            from msmctts._b200.functional import DeviceRng
            DeviceRng.advance(mel.device)      # device-side seed bump: a graph replay draws fresh dropout masks
        prefetch = None
        if mel.is_cuda and not self.reference_schedule:
            # weight re-parametrisation + operand images of every layer run ahead of the forward on a side stream
            from msmctts._b200.layers import prefetch_weights
            prefetch = prefetch_weights(ae)
            if gan and not warmup and disc is not None:
                prefetch_weights(disc)
        if warmup:
            output = ae(mel, mel_length, warmup=True)
        else:
            target = self.gather_windows(wav, starts * self.frameshift, self.frame_lengths * self.frameshift)
            output = ae(mel, mel_length, warmup=False, window=(starts, self.frame_lengths))
        vq = self.vq_criterion(output)
        losses.update(vq)
        g_loss = vq["vq_loss"]
        if "mel_outputs" in output:
            ml = F.mse_loss(mel, output["mel_outputs"], reduction="none")
            ml = ml.masked_fill(get_mask_from_lengths(mel_length, mel.shape[1]).unsqueeze(-1), 0)
            ml = ml.sum() / mel_length.sum() / ml.shape[2]
            losses["frame_loss"] = ml
            g_loss = g_loss + self.lambda_frame * ml
        if gan:
            predict = output["decoder_outputs"].squeeze(-1)
            target = target.reshape(predict.shape)
            stft_loss = self.stft_criterion(predict, target)
            losses["stft_loss"] = stft_loss
            g_loss = g_loss + self.lambda_stft * stft_loss
            # ---- discriminator step (reference :162-179)
            if self.reference_schedule:
                fake_scores, _ = disc(predict.detach())
                real_scores, _ = disc(target)
            else:
                nb = predict.shape[0]
                both, _ = disc(torch.cat([predict.detach(), target], dim=0))
                fake_scores, real_scores = [s[:nb] for s in both], [s[nb:] for s in both]
            if self.reference_schedule:
                d_real = sum(F.mse_loss(s, torch.ones_like(s)) for s in real_scores)
                d_fake = sum(F.mse_loss(s, torch.zeros_like(s)) for s in fake_scores)
            else:
                d_real, d_fake = lsgan_loss(real_scores, 1.0), lsgan_loss(fake_scores, 0.0)
            d_loss = d_real + d_fake
            losses.update(d_loss_real=d_real, d_loss_fake=d_fake, d_loss=d_loss)
            self.optimizer.zero_grad(["discriminator"])
            self.backward(d_loss, "discriminator")
            self.optimizer.step(["discriminator"])
            # ---- generator step (reference :182-201): D has already been updated, both passes are recomputed
            if self.reference_schedule:
                fake_scores, fake_feats = disc(predict)
                _, real_feats = disc(target)
            else:
                with _frozen(disc):         # D is a fixed function here: data gradients only, no weight gradients
                    if prefetch is not None:
                        prefetch_weights(disc)      # D was just updated: new parameter versions
                    fake_scores, fake_feats = disc(predict)
                    with torch.no_grad():
                        _, real_feats = disc(target)
            adv = sum(F.mse_loss(s, torch.ones_like(s)) for s in fake_scores) if self.reference_schedule \
                else lsgan_loss(fake_scores, 1.0)
            if self.reference_schedule or not predict.is_cuda:
                fm = sum(F.l1_loss(a, b) for fa, fb in zip(fake_feats, real_feats) for a, b in zip(fa, fb))
            else:
                # the 55 L1 terms in one multi-tensor kernel (real features carry no graph in this schedule)
                from msmctts._b200.functional import l1_multi
                fm = l1_multi([a for fa in fake_feats for a in fa], [b for fb in real_feats for b in fb])
            scale = self.lambda_fm if self.lambda_fm != "auto" else (g_loss / fm).detach()
            adv_loss = adv + fm * scale
            g_loss = g_loss + adv_loss
            losses.update(fm_loss=fm, adv_loss=adv_loss, g_loss=g_loss)
        self.optimizer.zero_grad(["autoencoder"])
        only = None if (self.reference_schedule or not gan) else list(self.model.autoencoder.parameters())
        self.backward(g_loss, "autoencoder", inputs=only)
        # clip_grad_norm_ + AdamW (reference :203-207) in the optimizer's two fused launches
        self.optimizer.step(["autoencoder"], max_grad_norm=self.grad_clip_thresh)
        if prefetch is not None:
            torch.cuda.current_stream().wait_stream(prefetch)      # join the side stream (graph capture needs it)
        return {"loss": {k: (v.detach() if torch.is_tensor(v) else v) for k, v in losses.items()}}


class PredictorTrainer(BaseTrainer):
    """multi-stage predictor training against a frozen autoencoder (reference :222-295)"""

    def __init__(self, config, model, num_gpus=1, rank=0, grad_clip_thresh=1.0, eval_inteval_iters=1000,
                 training_methods=["mse"], loss_weights=[1.0], lambda_dur=1.0):
        super().__init__(config, model, num_gpus, rank)
        self.training_methods, self.loss_weights = training_methods, loss_weights
        self.grad_clip_thresh = grad_clip_thresh
        self.dur_loss = DurationLoss(lambda_dur)

    def train_step(self, batch, iteration):
        batch = dict(batch)
        if not hasattr(self, "autoencoder"):
            self.build_autoencoder()
        self.autoencoder.eval()
        with torch.no_grad():
            qs = self.autoencoder.analysis(batch.pop("mel"), batch.pop("mel_length").int())
            batch["feat"], batch["feat_length"] = qs["quantizer_outputs"], qs["quantizer_lengths"]
        output = self.model.predictor(**batch)
        emb = self.autoencoder.compute_embedding_loss(output["feat"], output["feat_length"], qs,
                                                      methods=self.training_methods, loss_weights=self.loss_weights)
        losses = {"total_loss": emb.pop("total_loss")}
        losses.update(emb)
        dur = self.dur_loss(output, batch)
        losses["total_loss"] = losses["total_loss"] + dur.pop("total_loss")
        losses.update(dur)
        self.optimizer.zero_grad(["predictor"])
        self.backward(losses["total_loss"], "predictor")
        if self.grad_clip_thresh is not None:
            losses["grad_norm"] = self.optimizer.step(["predictor"], max_grad_norm=self.grad_clip_thresh)
        else:
            self.optimizer.step(["predictor"])
        return {"loss": {k: (v.detach() if torch.is_tensor(v) else v) for k, v in losses.items()}}

    def build_autoencoder(self, autoencoder=None):
        if autoencoder is None:
            cfg = self.config.task.autoencoder
            autoencoder = load_model("autoencoder", cfg._checkpoint, cfg.get("_config"))
        self.autoencoder = autoencoder.cuda() if torch.cuda.is_available() else autoencoder
