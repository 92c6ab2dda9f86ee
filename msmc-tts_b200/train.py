"""Entry point with the reference's CLI (train.py:32-67): `python train.py -c config.yaml [-r RANK -g GROUP]`.
Also accepts `--num_gpus` (train_dist.py passes it; the reference's own train.py rejects it -- SURVEY section 0, B6)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch  # noqa: E402

from msmctts.distributed.distributed import init_distributed  # noqa: E402
from msmctts.tasks import build_task  # noqa: E402
from msmctts.trainers import build_trainer  # noqa: E402
from msmctts.utils.config import Config  # noqa: E402


def train(config, num_gpus, rank, group_name):
    torch.manual_seed(config["seed"])
    if torch.cuda.is_available():
        torch.cuda.manual_seed(config["seed"])
    if num_gpus > 1:
        init_distributed(rank, num_gpus, group_name, **config.distributed)
        config.dataloader.batch_size = config.dataloader.batch_size // num_gpus
        print(f"Batch size per GPU is changed to {config.dataloader.batch_size}.")
    task = build_task(config, "train")
    trainer = build_trainer(config, task, num_gpus=num_gpus, rank=rank)
    trainer.train()
    if num_gpus > 1:
        # captured graphs hold NCCL all-reduce nodes: they go before the communicator (destroy_process_group() after
        # graph-captured collectives otherwise never returns)
        import torch.distributed as dist
        if hasattr(trainer, "release_graphs"):
            trainer.release_graphs()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
    print("Training done!")


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("-c", "--config", type=str, required=True, help="YAML file for configuration")
    parser.add_argument("-r", "--rank", type=int, default=0, help="rank of process for distributed")
    parser.add_argument("-g", "--group_name", type=str, default="", help="name of group for distributed")
    parser.add_argument("--num_gpus", type=int, default=None, help="world size (sent by train_dist.py)")
    args = parser.parse_args()
    config = Config(args.config)
    if "save_checkpoint_dir" not in config or not config.save_checkpoint_dir:
        config.save_checkpoint_dir = os.path.join(os.path.dirname(args.config), "checkpoints")
    num_gpus = args.num_gpus if args.num_gpus is not None else torch.cuda.device_count()
    if num_gpus > 1 and args.group_name == "":
        print("WARNING: Multiple GPUs detected but no distributed group set")
        num_gpus = 1
    if num_gpus == 1 and args.rank != 0:
        raise Exception("Doing single GPU training on rank > 0")
    train(config, num_gpus, args.rank, args.group_name)


if __name__ == "__main__":
    main()
