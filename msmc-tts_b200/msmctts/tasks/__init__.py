from os.path import dirname

import torch

from msmctts.utils.config import Config
from msmctts.utils.utils import load_checkpoint, module_search


def load_model(name, checkpoint_path, config_path=None):
    return getattr(load_task(checkpoint_path, config_path), name)


def load_task(checkpoint_path, config_path=None, mode="infer"):
    checkpoint = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
    config = Config(config_path if config_path is not None else checkpoint["config"])
    task = build_task(config, mode)
    load_checkpoint(checkpoint, task)
    return task


def build_task(config=None, mode="train", checkpoint=None, *args, **kwargs):
    assert config is not None or checkpoint is not None
    if checkpoint is not None:
        return load_task(checkpoint, config, mode)
    if isinstance(config, str):
        config = Config(config)
    TaskClass = module_search(config.task._name, dirname(__file__), "msmctts.tasks")
    return TaskClass(config, mode=mode, *args, **kwargs)
