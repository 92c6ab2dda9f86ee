"""Trainer loop, checkpoint find/load/save (reference trainers/base_trainer.py:16-142; host-side, own code)."""
import glob
import json
import os
import re

import torch

from msmctts.datasets import DevicePrefetcher, build_dataloader
from msmctts.distributed.distributed import apply_gradient_allreduce
from msmctts.utils.logger import Logger
from msmctts.utils.utils import load_checkpoint, to_model
from .lr_schedulers import build_lr_scheduler
from .optimizers import build_optimizer


class BaseTrainer(object):
    def __init__(self, config, model, num_gpus=1, rank=0):
        self.config = config
        self.distributed = num_gpus > 1
        self.rank = rank
        freeze = getattr(config, "freeze", "") if hasattr(config, "freeze") else ""
        if freeze:
            for name, p in model.named_parameters():
                if re.match(freeze, name):
                    p.requires_grad = False
        if num_gpus > 0 and torch.cuda.is_available():
            model = model.cuda()
        if self.distributed:
            model = apply_gradient_allreduce(model)
        self.model = model
        self.optimizer = None

    def build_optimizer(self):
        self.optimizer = build_optimizer(self.model, self.config.optimizer)
        return self.optimizer

    def backward(self, loss, module_name, inputs=None):
        """backward + (when data-parallel) the bucketed, overlapped all-reduce of that sub-module's gradients;
        `inputs` restricts the pass to those leaves (gradients of anything else are neither computed nor stored)"""
        reducer = getattr(self.model, "grad_reducers", {}).get(module_name) if self.distributed else None
        if reducer is not None:
            reducer.arm()
        if inputs is None:
            loss.backward()
        else:
            loss.backward(inputs=[p for p in inputs if p.requires_grad])
        if reducer is not None:
            reducer.finish()

    def train(self):
        _, sampler, loader = build_dataloader(self.config.dataset, self.config.dataloader, self.distributed)
        self.build_optimizer()
        lr_scheduler = build_lr_scheduler(self.config.lr_scheduler)
        iteration = self.attempt_load_checkpoint()
        logger = Logger(self.config.save_checkpoint_dir, "GPU_%d_" % self.rank if self.distributed else "",
                        "GPU_%d.log" % self.rank if self.distributed else "train.log",
                        window=int(self.config.get("log_window", 100)), echo=bool(self.config.get("log_echo", False)))
        logger.info(json.dumps(self.config.to_dict(), indent=2, default=str))
        self.model.train()
        while True:
            epoch = iteration // max(1, len(loader))
            if sampler is not None:
                sampler.set_epoch(epoch)
            # batch i+1 is uploaded from pinned memory on a side stream while step i runs (datasets/prefetch.py)
            use_prefetch = torch.cuda.is_available() and next(self.model.parameters()).is_cuda
            for batch in (DevicePrefetcher(loader) if use_prefetch else loader):
                lr_scheduler.step(self.optimizer, iteration)
                batch = to_model(batch)
                self.optimizer.zero_grad()         # reference base_trainer.py:84-85 (set_to_none: no launches)
                log = self.train_step(batch, iteration)
                logger.log(iteration, log)
                if self.rank == 0 and iteration > 0 and iteration % self.config.iters_per_checkpoint == 0:
                    self.save_checkpoint("{}/model_{}".format(self.config.save_checkpoint_dir, iteration), iteration)
                if iteration >= self.config.training_steps:
                    return
                iteration += 1

    def train_step(self, batch, iteration):
        raise NotImplementedError

    def attempt_load_checkpoint(self):
        restore = self.config.restore_checkpoint_path
        latest = self.find_latest_checkpoint()
        if self.config.resume_training and latest != "":
            restore = latest
        if restore != "":
            return load_checkpoint(restore, self.model, self.optimizer) + 1
        if self.config.pretrain_checkpoint_path != "":
            load_checkpoint(self.config.pretrain_checkpoint_path, self.model)
        return 0

    def find_latest_checkpoint(self):
        directory = self.config.save_checkpoint_dir
        if not os.path.exists(directory):
            return ""
        its = [int(x.split("_")[-1]) for x in glob.glob(os.path.join(directory, "model_*"))]
        return os.path.join(directory, "model_%d" % max(its)) if its and max(its) > 0 else ""

    def save_checkpoint(self, filepath, iteration):
        os.makedirs(os.path.dirname(filepath), exist_ok=True)
        torch.save({"model": self.model.state_dict(), "optimizer": self.optimizer.state_dict(),
                    "iteration": iteration, "config": self.config.to_dict()}, filepath)
