"""msmctts.networks -- the drop-in boundary: yaml `_name` -> class lookup (reference networks/__init__.py:6-11)."""
import os

from msmctts.utils.utils import module_search


def find_modules(conf):
    module_names, confs = zip(*conf.items())
    names = [x["_name"] for x in confs]
    kwargs = [{k: v for k, v in c.items() if k[:1] != "_"} for c in confs]
    modules = module_search(names, os.path.dirname(__file__), "msmctts.networks")
    return [(x, modules[i](**kwargs[i])) for i, x in enumerate(module_names)]
