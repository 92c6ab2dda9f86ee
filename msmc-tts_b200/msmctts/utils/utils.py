"""Host-side helpers the reference's entry points expect under msmctts.utils.utils (own implementation;
interface follows reference utils/utils.py:137-157, 207-316)."""
import glob
import importlib
import inspect
import os
import re

import torch


def get_mask_from_lengths(lengths, max_len=None):
    """True on padding.  Unlike the reference (utils.py:155) no `.item()` host sync is needed when max_len is given."""
    if max_len is None:
        max_len = int(torch.max(lengths).item())
    ids = torch.arange(0, max_len, device=lengths.device)
    return ~(ids < lengths.unsqueeze(1))


def lengths_i32(lengths, device):
    return lengths.to(device=device, dtype=torch.int32).contiguous()


def to_gpu(x):
    x = x.contiguous()
    if torch.cuda.is_available():
        x = x.cuda(non_blocking=True)
    return x


def to_model(x):
    if isinstance(x, (tuple, list)):
        return [to_model(v) for v in x]
    if isinstance(x, dict):
        return {k: to_model(v) for k, v in x.items()}
    return to_gpu(torch.as_tensor(x))


def load_checkpoint(checkpoint_object, model, optimizer=None, module=None):
    """Same contract as reference utils.py:207-250: path | dict | list of [regex, path]; returns the iteration."""
    if isinstance(checkpoint_object, (tuple, list)):
        it = 0
        for pattern, obj in checkpoint_object:
            it = max(it, load_checkpoint(obj, model, optimizer, pattern))
        return it
    if isinstance(checkpoint_object, str):
        if not os.path.isfile(checkpoint_object):
            raise FileNotFoundError(checkpoint_object)
        ckpt = torch.load(checkpoint_object, map_location="cpu", weights_only=False)
    elif isinstance(checkpoint_object, dict):
        ckpt = checkpoint_object
    else:
        raise TypeError("Unacceptable type: %s" % type(checkpoint_object))
    params = ckpt["model"]
    iteration = ckpt.get("iteration", 0)
    if module is not None:
        wanted = {k: params[k] for k in model.state_dict().keys() if re.match(module, k)}
        model.load_state_dict(wanted, strict=False)
    else:
        try:
            model.load_state_dict(params)
            if optimizer is not None:
                optimizer.load_state_dict(ckpt["optimizer"])
        except Exception:
            print("Loaded model is not the same as the current one")
            model.load_state_dict(params, strict=False)
    print("Checkpoint loading is completed.")
    return iteration


def save_checkpoint(checkpoint_dict, filepath, autoclean=False, save_interval=50000):
    torch.save(checkpoint_dict, filepath)


def module_search(names, directory, package=None):
    """Resolve class names by importing every module / sub-package of `directory` (reference utils.py:276-316)."""
    single = isinstance(names, str)
    anchors = [names] if single else list(names)
    files = glob.glob(os.path.join(directory, "*.py")) + glob.glob(os.path.join(directory, "*", "__init__.py"))
    mods = []
    for f in files:
        rel = f[len(directory):][:-3].replace(os.path.sep, ".").replace(".__init__", "")
        if rel and not rel.endswith("__init__") and rel != ".":
            mods.append(rel)
    found = [None] * len(anchors)
    for i, name in enumerate(anchors):
        cls_name = name.split(".")[-1]
        sub = name[: -len(cls_name) - 1]
        space = [package + "." + sub] if sub else mods
        for mf in space:
            if mf.split(".")[-1].startswith("_"):
                continue
            mod = importlib.import_module(mf, package=package)
            if not hasattr(mod, cls_name):
                continue
            cls = getattr(mod, cls_name)
            if found[i] is not None:
                if inspect.getfile(found[i]) != inspect.getfile(cls):
                    raise RuntimeError("Repeated Module for %s" % cls_name)
                continue
            found[i] = cls
    if None in found:
        raise RuntimeError("Found dismatched modules for {}".format(names))
    return found[0] if single else found
