"""Names the yaml `_name` lookup resolves in this sub-package: the multi-stage predictor (reference
networks/acoustic_models/__init__.py:1) plus the FFT block stack the autoencoder imports from here."""
from . import multi_stage_predictor as _msp
from . import transformer as _transformer

FFTBlocks = _transformer.FFTBlocks
MultiStagePredictor = _msp.MultiStagePredictor

__all__ = ["FFTBlocks", "MultiStagePredictor"]
