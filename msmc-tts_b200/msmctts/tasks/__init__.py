"""Task construction and checkpoint loading (interface of reference tasks/__init__.py:9-43: `build_task`,
`load_task`, `load_model`; own implementation)."""
import os

import torch

from msmctts.utils.config import Config
from msmctts.utils.utils import load_checkpoint, module_search

_TASK_DIR = os.path.dirname(__file__)


def _task_class(config):
    """yaml `task._name` -> class, searched in this package like the network registry"""
    return module_search(config.task._name, _TASK_DIR, "msmctts.tasks")


def build_task(config=None, mode="train", checkpoint=None, *args, **kwargs):
    """A task from a Config / yaml path, or (when `checkpoint` is given) from a checkpoint file whose stored config
    is used unless `config` overrides it."""
    if checkpoint is not None:
        return load_task(checkpoint, config, mode)
    if config is None:
        raise AssertionError("build_task needs a config or a checkpoint")
    cfg = Config(config) if isinstance(config, str) else config
    return _task_class(cfg)(cfg, mode=mode, *args, **kwargs)


def load_task(checkpoint_path, config_path=None, mode="infer"):
    """Rebuild the task a checkpoint was written by and load its weights (on the CPU; the caller moves it)."""
    state = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
    cfg = Config(state["config"] if config_path is None else config_path)
    task = build_task(cfg, mode)
    load_checkpoint(state, task)
    return task


def load_model(name, checkpoint_path, config_path=None):
    """One sub-network (`autoencoder`, `predictor`, ...) of a checkpointed task"""
    return getattr(load_task(checkpoint_path, config_path), name)
