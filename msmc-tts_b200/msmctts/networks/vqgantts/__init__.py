"""Name the yaml `_name` lookup resolves in this sub-package (reference networks/vqgantts/__init__.py:1; its second
export, the QS-TTS `MSMCVQGANEmb`, points at a file that is not in the reference tree -- SURVEY section 0, B2)."""
from . import msmc_vqgan as _msmc_vqgan

MSMCVQGAN = _msmc_vqgan.MSMCVQGAN

__all__ = ["MSMCVQGAN"]
