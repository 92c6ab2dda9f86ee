"""Learning-rate schedule (reference trainers/lr_schedulers/): exponential decay after a flat warm-up."""


class ExponentialDecayLRScheduler(object):
    def __init__(self, warmup_steps=50000, decay_scale=50000, decay_learning_rate=0.5, final_learning_rate=1e-5):
        self.warmup_steps, self.decay_scale = warmup_steps, decay_scale
        self.decay_learning_rate, self.final_learning_rate = decay_learning_rate, final_learning_rate

    def step(self, optimizer, iteration):
        for key, opt in optimizer.optimizers.items():
            base = optimizer.config[key].learning_rate
            lr = base
            if iteration >= self.warmup_steps:
                lr = base * self.decay_learning_rate ** ((iteration - self.warmup_steps) / self.decay_scale)
            lr = max(lr, self.final_learning_rate)
            for group in opt.param_groups:
                if hasattr(group["lr"], "fill_"):
                    group["lr"].fill_(lr)       # device tensor (capturable optimizers): no re-capture needed
                else:
                    group["lr"] = lr


def build_lr_scheduler(config):
    kwargs = {k: v for k, v in config.items() if not k.startswith("_")}
    name = config.get("_name", "ExponentialDecayLRScheduler")
    if name != "ExponentialDecayLRScheduler":
        raise ValueError(name)
    return ExponentialDecayLRScheduler(**kwargs)
