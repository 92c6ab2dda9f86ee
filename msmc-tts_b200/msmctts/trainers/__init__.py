"""Trainer lookup: the yaml `trainer` block names a class of this package and carries its keyword arguments
(interface of reference trainers/__init__.py:6-12)."""
import os

from msmctts.utils.utils import module_search

_HERE = os.path.dirname(__file__)


def build_trainer(config, model, num_gpus=1, rank=0):
    options = dict(config.trainer.to_dict())
    cls = module_search(options.pop("_name"), _HERE, "msmctts.trainers")
    return cls(config, model, num_gpus=num_gpus, rank=rank, **options)
