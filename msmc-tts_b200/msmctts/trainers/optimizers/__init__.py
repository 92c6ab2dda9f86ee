"""Per-child-module optimizers (reference trainers/optimizers/__init__.py:9-78).  On CUDA the step is the fused
multi-tensor kernel (fused.py) so a whole train step can be graph-captured; CPU tensors (host-side
tests) get torch.optim."""
import re

import torch
from torch.optim import Adam, AdamW


def get_optimizer(parameters, config):
    parameters = list(parameters)
    name = config._name
    kw = dict(lr=config.learning_rate, betas=tuple(config.betas), eps=config.eps, weight_decay=config.weight_decay)
    if parameters and parameters[0].is_cuda:
        if name in ("Adam", "AdamW"):
            # one fused multi-tensor launch per optimizer, device-resident lr and step (CUDA-graph replayable)
            from .fused import FusedAdam
            return FusedAdam(parameters, decoupled=(name == "AdamW"), **kw)
    if name == "Adam":
        return Adam(parameters, **kw)
    if name == "AdamW":
        return AdamW(parameters, **kw)
    raise ValueError("optimizer %s is not provided by the B200 backend (Adam / AdamW are)" % name)


def build_optimizer(model, config):
    optimizers, configs = {}, {}
    for module_name, module in model.named_children():
        try:
            module_config = config[module_name]
        except KeyError:
            assert hasattr(config, "_default"), "Both {} and _default not found".format(module_name)
            module_config = config._default
        configs[module_name] = module_config
        # every parameter, frozen ones included, exactly like the reference (:38): param-group sizes then match the
        # reference's optimizer checkpoints whatever `config.freeze` says; parameters without a gradient are skipped
        # by the step (torch and FusedAdam alike)
        parameters = list(module.parameters())
        if hasattr(module_config, "parameters"):
            parameters = []
            for name, p in module.named_parameters():
                if re.match(module_config.parameters, name):
                    parameters.append(p)
                else:
                    p.requires_grad = False
        optimizers[module_name] = get_optimizer(parameters, module_config)
    return Optimizer(optimizers, configs)


class Optimizer(object):
    def __init__(self, optimizers_dict, config):
        self.optimizers = optimizers_dict
        self.config = config

    def _names(self, names):
        if names is None:
            return tuple(self.optimizers.keys())
        return names if isinstance(names, (list, tuple)) else [names]

    def load_state_dict(self, state):
        for key in self.optimizers:
            self.optimizers[key].load_state_dict(state[key])

    def state_dict(self):
        return {key: opt.state_dict() for key, opt in self.optimizers.items()}

    def zero_grad(self, names=None):
        for key in self._names(names):
            self.optimizers[key].zero_grad(set_to_none=True)

    def step(self, names=None, max_grad_norm=None):
        """max_grad_norm: clip the global gradient norm of each named optimizer's parameters first (reference
        trainers/msmctts_trainer.py:203-207); fused into the update launches on CUDA.  Returns the norm(s)."""
        norms = []
        for key in self._names(names):
            opt = self.optimizers[key]
            if max_grad_norm is None:
                opt.step()
            elif hasattr(opt, "last_grad_norm"):
                opt.step(max_grad_norm=max_grad_norm)
                norms.append(opt.last_grad_norm)
            else:
                params = [p for g in opt.param_groups for p in g["params"]]
                norms.append(torch.nn.utils.clip_grad_norm_(params, max_grad_norm))
                opt.step()
        return norms[0] if len(norms) == 1 else norms
