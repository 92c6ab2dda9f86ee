"""Names the yaml `_name` lookup resolves in this sub-package (the reference exports the same two,
networks/hifigan/__init__.py:1-2); the classes run on the sm_100a kernels."""
from . import discriminator as _discriminator
from . import generator as _generator

HifiGANGenerator = _generator.Generator
UnivNetDiscriminator = _discriminator.Discriminator

__all__ = ["HifiGANGenerator", "UnivNetDiscriminator"]
