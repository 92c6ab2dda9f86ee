"""Text logger with windowed averages (reference utils/logger.py; tensorboardX is optional here).  Loss values may be
0-dim device tensors: they are only read (one host sync) when a window is flushed."""
import os
import time

import torch


class Logger(object):
    def __init__(self, log_dir, prefix="", filename="train.log", window=100, echo=False):
        self.window, self.prefix, self.echo = window, prefix, echo
        self.sums, self.count = {}, 0
        self.t_window = time.perf_counter()
        self.file = None
        if log_dir:
            os.makedirs(log_dir, exist_ok=True)
            self.file = open(os.path.join(log_dir, filename), "a")

    def info(self, text):
        line = "[%s] %s" % (time.strftime("%Y-%m-%d %H:%M:%S"), text)
        if self.file:
            self.file.write(line + "\n")
            self.file.flush()
        if self.echo:
            print(self.prefix + line, flush=True)

    def log(self, iteration, log):
        for k, v in (log or {}).get("loss", {}).items():
            v = v.detach() if torch.is_tensor(v) else torch.as_tensor(float(v))
            self.sums[k] = self.sums[k] + v if k in self.sums else v.clone()
        self.count += 1
        if self.count >= self.window:
            # float(v) is the one host sync of the window: the wall time per step below is honest device time
            text = ", ".join("%s=%.5f" % (k, float(v) / self.count) for k, v in sorted(self.sums.items()))
            now = time.perf_counter()
            self.info("iter %d: %s | %.2f ms/step over the last %d steps" % (
                iteration, text, (now - self.t_window) * 1e3 / self.count, self.count))
            self.sums, self.count, self.t_window = {}, 0, now
