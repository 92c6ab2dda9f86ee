"""yaml -> attribute dict with defaults (same behaviour as reference utils/config.py:6-110; own implementation)."""
import os
import re

import yaml

DEFAULT_DICT = {
    "id": "null",
    "save_checkpoint_dir": "",
    "pretrain_checkpoint_path": "",
    "restore_checkpoint_path": "",
    "resume_training": True,
    "training_steps": 1000000,
    "iters_per_checkpoint": 50000,
    "seed": 1234,
    "cudnn": {"enabled": True, "benchmark": False},
    "distributed": {"dist_backend": "nccl", "dist_url": "tcp://localhost:54321"},
}

_FLOAT = re.compile(r"""^(?:[-+]?(?:[0-9][0-9_]*)\.[0-9_]*(?:[eE][-+]?[0-9]+)?
                        |[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)
                        |\.[0-9_]+(?:[eE][-+][0-9]+)?
                        |[-+]?\.(?:inf|Inf|INF)|\.(?:nan|NaN|NAN))$""", re.X)


def load_yaml(path):
    class _Loader(yaml.SafeLoader):
        pass
    _Loader.add_implicit_resolver("tag:yaml.org,2002:float", _FLOAT, list("-+0123456789."))   # accept 2e-4
    with open(path) as f:
        return yaml.load(f, Loader=_Loader)


class ConfigItem(dict):
    __slots__ = ()

    def __init__(self, source=None):
        super().__init__()
        for key, value in (source or {}).items():
            self[key] = self._wrap(value)

    @staticmethod
    def _wrap(value):
        if isinstance(value, dict):
            return ConfigItem(value)
        if isinstance(value, (list, tuple)):
            return [ConfigItem(v) if isinstance(v, dict) else v for v in value]
        if isinstance(value, str) and value.lower() == "none":
            return None
        return value

    def __getattr__(self, item):
        try:
            return self[item]
        except KeyError:
            raise AttributeError(item)

    def __setattr__(self, name, value):
        self[name] = value

    def to_dict(self, recursive=True):
        return {k: (v.to_dict(True) if recursive and isinstance(v, ConfigItem) else v) for k, v in self.items()}

    def update(self, other):
        for k, v in other.items():
            if k in self and isinstance(v, dict) and isinstance(self[k], ConfigItem):
                self[k].update(v)
            else:
                self[k] = self._wrap(v) if not isinstance(v, ConfigItem) else v


class Config(ConfigItem):
    def __init__(self, yaml_object):
        super().__init__(DEFAULT_DICT)
        if isinstance(yaml_object, str):
            if not os.path.isfile(yaml_object):
                raise FileNotFoundError(yaml_object)
            yaml_object = load_yaml(yaml_object)
        self.update(ConfigItem(yaml_object) if not isinstance(yaml_object, ConfigItem) else yaml_object)
