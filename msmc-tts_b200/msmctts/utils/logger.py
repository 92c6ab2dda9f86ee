"""Text logger with windowed averages (reference utils/logger.py; tensorboardX is optional here).  Loss values may be
0-dim device tensors: they are only read (one host sync) when a window is flushed."""
import os
import time

import torch


class Logger(object):
    def __init__(self, log_dir, prefix="", filename="train.log", window=100):
        self.window, self.prefix = window, prefix
        self.sums, self.count = {}, 0
        self.file = None
        if log_dir:
            os.makedirs(log_dir, exist_ok=True)
            self.file = open(os.path.join(log_dir, filename), "a")

    def info(self, text):
        line = "[%s] %s" % (time.strftime("%Y-%m-%d %H:%M:%S"), text)
        if self.file:
            self.file.write(line + "\n")
            self.file.flush()

    def log(self, iteration, log):
        for k, v in (log or {}).get("loss", {}).items():
            v = v.detach() if torch.is_tensor(v) else torch.as_tensor(float(v))
            self.sums[k] = self.sums[k] + v if k in self.sums else v.clone()
        self.count += 1
        if self.count >= self.window:
            self.info("iter %d: " % iteration + ", ".join(
                "%s=%.5f" % (k, float(v) / self.count) for k, v in sorted(self.sums.items())))
            self.sums, self.count = {}, 0
