"""Container of the named sub-networks a yaml `task` block describes (interface of reference
tasks/base_task.py:6-33: sub-networks become attributes, `forward` dispatches on the construction mode)."""
import torch

from msmctts.networks import find_modules


class BaseTask(torch.nn.Module):
    _STEP_OF_MODE = {"train": "train_step", "infer": "infer_step", "debug": "debug_step"}

    def __init__(self, config, mode="train"):
        super().__init__()
        self.config, self.mode = config, mode
        task = config.task
        if hasattr(task, "network"):
            blocks = task.network
        else:                      # every non-directive entry that names a class is a sub-network
            blocks = {k: v for k, v in task.items() if not k.startswith("_") and "_name" in v}
        for name, network in find_modules(blocks):
            self.add_module(name, network)

    def forward(self, features):
        return getattr(self, self._STEP_OF_MODE[self.mode])(features)

    # subclasses override the steps they support
    def train_step(self, features):
        return None

    def infer_step(self, features):
        return None

    def debug_step(self, features):
        return None
