"""Build the sm_100a C-ABI library of the MSMC-VQ-GAN hot path, in-tree.

    python msmc-tts_b200/build.py            # -> msmc-tts_b200/msmctts/_b200/libmsmc_b200.so

nvcc cross-compiles without a GPU.  Every translation unit under csrc/ is compiled with
`-gencode arch=compute_100a,code=sm_100a -lineinfo`; the objects are cached by source hash.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "msmctts", "_b200")
OBJ_DIR = os.environ.get("MSMC_OBJ_DIR", "/tmp/msmc_b200_obj")   # object cache lives outside the tree
LIB = os.path.join(OUT_DIR, "libmsmc_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
STAMP = os.path.join(OUT_DIR, "libmsmc_b200.stamp")


def _digest(path, extra):
    h = hashlib.sha256()
    h.update(" ".join(FLAGS).encode())
    for p in [path] + extra:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def build(verbose=True):
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(OUT_DIR, exist_ok=True)
    headers = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    headers.append(os.path.join(HERE, "..", "include", "msmc_b200.h"))
    sources = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    # the library travels to the GPU box without the object cache: a stamp of all source digests lets build()
    # recognise an up-to-date .so instead of recompiling everything there
    stamp = " ".join(_digest(os.path.join(CSRC, src), headers) for src in sources)
    if os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read() == stamp:
        return LIB
    objs, jobs = [], []
    for src in sources:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ_DIR, "%s.%s.o" % (src[:-3], _digest(path, headers)))
        if not os.path.exists(obj):
            jobs.append([NVCC] + FLAGS + ["-c", path, "-o", obj + ".tmp.o"])
        objs.append(obj)
    rebuilt = bool(jobs)
    if jobs:
        # translation units compile side by side (conv_umma.cu alone takes ~2.5 min); objects appear atomically
        from concurrent.futures import ThreadPoolExecutor

        def run(cmd):
            if verbose:
                print("[build]", " ".join(cmd), flush=True)
            subprocess.check_call(cmd)
            os.replace(cmd[-1], cmd[-1][:-len(".tmp.o")])
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as pool:
            list(pool.map(run, jobs))
    if rebuilt or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-cudart", "shared", "-Xlinker", "-rpath,/usr/local/cuda/lib64", "-o", LIB] + objs
        if verbose:
            print("[build]", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    with open(STAMP, "w") as f:
        f.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build())
