"""MSMCTTS task: dispatch by `_mode` (reference tasks/msmc_tts.py:93-159)."""
import torch

from . import load_model
from .base_task import BaseTask


class MSMCTTS(BaseTask):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        ds = self.config.dataset
        self.samplerate = ds.samplerate
        self.fs = {ds.feature[i]: ds.frameshift[i] for i in range(len(ds.feature))}
        self.training_mode = self.config.task._mode
        self.load_modules = False

    def train_step(self, input_dict, mode=None):
        mode = mode or self.training_mode
        return self.analysis_synthesis(input_dict) if mode == "train_autoencoder" else self.predict(input_dict)

    def infer_step(self, input_dict, mode=None):
        mode = mode or self.training_mode
        if mode == "train_autoencoder":
            return self.analysis_synthesis(input_dict)
        if not self.load_modules:
            self.pre_infer()
        return self.predict(input_dict)

    def analysis_synthesis(self, input_dict):
        out = self.autoencoder(**input_dict)
        return {"wav": out["decoder_outputs"].squeeze(-1)}

    def predict(self, input_dict):
        input_dict = dict(input_dict)
        input_dict.pop("mel", None)
        input_dict.pop("mel_length", None)
        output_dict = self.predictor(**input_dict)
        feats, lengths = output_dict["feat"], output_dict["feat_length"]
        wavs = self.autoencoder.synthesis(feats, lengths)[..., 0]
        wav_lengths = (lengths[-1] * wavs.shape[1] / feats[-1].shape[1]).int()
        output_dict["wav"] = [x[:l] for x, l in zip(wavs, wav_lengths)]
        output_dict["embedding"] = feats[-1]
        return output_dict

    def pre_infer(self):
        self.load_modules = True
        ae_cfg = self.config.task.get("autoencoder")
        if ae_cfg is not None and "_checkpoint" in ae_cfg:
            model = load_model("autoencoder", ae_cfg._checkpoint, ae_cfg.get("_config"))
            self.autoencoder = (model.cuda() if torch.cuda.is_available() else model).eval()
        if hasattr(self, "predictor") and hasattr(self, "autoencoder"):
            self.predictor.autoencoder = self.autoencoder
            self.predictor.quantizers = self.autoencoder.quantizer.quantizer
