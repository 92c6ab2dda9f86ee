"""Inference entry point with the reference's CLI (infer.py:98-133): `python infer.py -m CHECKPOINT [-c CONFIG]
[-t TEST_LIST] [-o OUT_DIR] [-j JOBS]`.  Loads the task from a checkpoint, runs `task.infer_step` over the test set
and writes every feature named in the yaml's `save_features` block ([name, extension, sample_rate])."""
import argparse
import os
import re
import sys
import wave

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch.utils.data import DataLoader, SequentialSampler  # noqa: E402

from msmctts.datasets import build_dataset  # noqa: E402
from msmctts.tasks import build_task  # noqa: E402
from msmctts.utils.utils import to_model  # noqa: E402


def output_base(checkpoint_path):
    m = re.match(r".*_([0-9]+)$", checkpoint_path)
    return os.path.join(os.path.dirname(checkpoint_path), "eval-%d" % int(m.group(1)) if m else "eval")


def save_feature(path, feat, fmt, sample_rate):
    if fmt == ".npy":
        np.save(path, feat)
    elif fmt == ".txt":
        np.savetxt(path, feat, fmt="%.6f")
    elif fmt == ".dat":
        feat.astype(np.float32).tofile(path)
    elif fmt == ".wav":
        feat = np.asarray(feat, dtype=np.float64).reshape(-1)
        peak = float(np.abs(feat).max()) if feat.size else 0.0
        if peak > 1:
            feat = feat / peak
        with wave.open(path, "wb") as f:
            f.setnchannels(1)
            f.setsampwidth(2)
            f.setframerate(int(sample_rate))
            f.writeframes((feat * 32767.0).astype("<i2").tobytes())
    else:
        raise ValueError("unsupported output format %s" % fmt)


def run(task, testset, output_dir, jobs=1):
    loader = DataLoader(testset, batch_size=jobs, num_workers=0, shuffle=False, drop_last=False,
                        sampler=SequentialSampler(testset), collate_fn=getattr(testset, "collate_fn", None))
    if torch.cuda.is_available():
        task = task.cuda()
    task.eval()
    if not hasattr(task.config, "save_features"):
        raise ValueError("the config names no `save_features`")
    dirs = {}
    for name, _, _ in task.config.save_features:
        dirs[name] = os.path.join(output_dir, name)
        os.makedirs(dirs[name], exist_ok=True)
    with torch.no_grad():
        for features in loader:
            ids = [testset.id_list[int(i)] for i in features.pop("_id")]
            out = task(to_model(features))
            for i, uid in enumerate(ids):
                for name, fmt, sr in task.config.save_features:
                    feat = out[name][i]
                    if torch.is_tensor(feat):
                        feat = feat.detach().cpu().numpy()
                    stem = uid[0] if isinstance(uid, (tuple, list)) else uid
                    save_feature(os.path.join(dirs[name], "%s%s" % (stem, fmt)), feat, fmt, sr)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-m", "--model", required=True)
    ap.add_argument("-c", "--config", default=None)
    ap.add_argument("-t", "--test_config", default=None)
    ap.add_argument("-j", "--jobs", type=int, default=1)
    ap.add_argument("-o", "--output_dir", default=None)
    ap.add_argument("--debug", action="store_true")
    args = ap.parse_args()
    task = build_task(args.config, mode="debug" if args.debug else "infer", checkpoint=args.model)
    ds_cfg = task.config.testset if hasattr(task.config, "testset") else task.config.dataset
    ds_cfg["training"] = False
    if args.test_config is not None:
        ds_cfg["id_list"] = args.test_config
    dataset = build_dataset(ds_cfg)
    out_dir = args.output_dir or output_base(args.model)
    os.makedirs(out_dir, exist_ok=True)
    run(task, dataset, out_dir, jobs=args.jobs)


if __name__ == "__main__":
    main()
