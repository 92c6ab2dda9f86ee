from os.path import dirname

from msmctts.utils.utils import module_search


def build_trainer(config, model, num_gpus=1, rank=0):
    """yaml `trainer._name` -> class (reference trainers/__init__.py:6-12)"""
    kwargs = config.trainer.to_dict()
    name = kwargs.pop("_name")
    Trainer = module_search(name, dirname(__file__), "msmctts.trainers")
    return Trainer(config, model, num_gpus=num_gpus, rank=rank, **kwargs)
