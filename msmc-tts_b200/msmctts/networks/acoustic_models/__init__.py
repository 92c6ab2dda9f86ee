from .multi_stage_predictor import MultiStagePredictor
from .transformer import FFTBlocks
