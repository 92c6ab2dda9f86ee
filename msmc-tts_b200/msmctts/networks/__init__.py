"""msmctts.networks -- the drop-in boundary: every entry of the yaml `task` block names a class (`_name`) exported by
one of the sub-packages here and carries its constructor kwargs (keys starting with `_` are directives, not kwargs).
Interface of reference networks/__init__.py:6-11."""
import os

from msmctts.utils.utils import module_search

_HERE = os.path.dirname(__file__)


def _constructor_kwargs(block):
    return {key: value for key, value in block.items() if not key.startswith("_")}


def find_modules(conf):
    """{attribute name: yaml block} -> [(attribute name, constructed network)] in yaml order"""
    attr_names = list(conf.keys())
    classes = module_search([conf[a]["_name"] for a in attr_names], _HERE, "msmctts.networks")
    return [(a, cls(**_constructor_kwargs(conf[a]))) for a, cls in zip(attr_names, classes)]
