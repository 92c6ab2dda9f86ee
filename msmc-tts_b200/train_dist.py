"""One train.py process per visible GPU (reference train_dist.py:13-35), rendezvous on 127.0.0.1.
Unlike the reference it fails fast: if any rank dies the others are terminated instead of hanging in NCCL."""
import argparse
import os
import subprocess
import sys
import time

import torch

TRAIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "train.py")


def multi_gpu(stdout_dir, config_file, num_gpus=None):
    num_gpus = num_gpus or torch.cuda.device_count()
    group = "group_{}".format(time.strftime("%Y_%m_%d-%H%M%S"))
    os.makedirs(stdout_dir, exist_ok=True)
    workers = []
    for i in range(num_gpus):
        cmd = [sys.executable, TRAIN, "-c", config_file, "--num_gpus={}".format(num_gpus), "--rank={}".format(i),
               "--group_name={}".format(group)]
        out = None if i == 0 else open(os.path.join(stdout_dir, "GPU_{}.log".format(i)), "w")
        workers.append(subprocess.Popen(cmd, stdout=out))
    rc = 0
    while workers:
        for p in list(workers):
            r = p.poll()
            if r is None:
                continue
            workers.remove(p)
            if r != 0:
                rc = r
                for q in workers:
                    q.terminate()
        time.sleep(0.5)
    return rc


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("-s", "--stdout_dir", type=str, default="/tmp/msmc-tts/logs")
    parser.add_argument("-c", "--config_file", type=str, required=True)
    parser.add_argument("-n", "--num_gpus", type=int, default=None)
    args = parser.parse_args()
    sys.exit(multi_gpu(args.stdout_dir, args.config_file, args.num_gpus))


if __name__ == "__main__":
    main()
