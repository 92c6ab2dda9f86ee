"""VQ oracle: ctypes wrapper of vq_oracle.c plus a numpy restatement.  TEST INFRASTRUCTURE ONLY.
Follows reference msmctts/networks/vqgantts/modules.py:24-67 (Quantize.forward) and 137-151 (MultiHeadQuantize)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libvq_oracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def search_c(z, embed):
    """z (n_rows, H*dim) f32, embed (H, dim, K) f32 -> quant_raw, quant_st, diff (n_rows, dim), idx (n_rows, H) i64"""
    z = np.ascontiguousarray(z, dtype=np.float32)
    embed = np.ascontiguousarray(embed, dtype=np.float32)
    H, dim, K = embed.shape
    n = z.shape[0]
    q_raw = np.empty((n, H * dim), np.float32)
    q_st = np.empty_like(q_raw)
    diff = np.empty((n, dim), np.float32)
    idx = np.empty((n, H), np.int64)
    _load().vq_oracle_search(_p(z), C.c_int64(z.shape[1]), _p(embed), _p(q_raw), _p(q_st), _p(diff), _p(idx),
                             n, H, dim, K)
    return q_raw, q_st, diff, idx


def ema_c(z, idx, lengths, batch, t, embed, embed_avg, cluster_size, decay=0.99, eps=1e-5):
    """in-place on copies; returns (cluster_size, embed_avg, embed)"""
    z = np.ascontiguousarray(z, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    lengths = np.ascontiguousarray(lengths, dtype=np.int32)
    embed = np.array(embed, dtype=np.float32, copy=True)
    embed_avg = np.array(embed_avg, dtype=np.float32, copy=True)
    cluster_size = np.array(cluster_size, dtype=np.float32, copy=True)
    H, dim, K = embed.shape
    _load().vq_oracle_ema(_p(z), C.c_int64(z.shape[1]), _p(idx), _p(lengths), batch, t, H, dim, K,
                          C.c_float(decay), C.c_float(eps), _p(cluster_size), _p(embed_avg), _p(embed))
    return cluster_size, embed_avg, embed


def search_np(z, embed):
    """numpy restatement of modules.py:25-33 in fp64 (formulation check / tie diagnostics)."""
    H, dim, K = embed.shape
    n = z.shape[0]
    idx = np.empty((n, H), np.int64)
    gap = np.empty((n, H), np.float64)
    for h in range(H):
        zh = z[:, h * dim:(h + 1) * dim].astype(np.float64)
        E = embed[h].astype(np.float64)
        dist = (zh ** 2).sum(1, keepdims=True) - 2 * zh @ E + (E ** 2).sum(0, keepdims=True)
        order = np.argsort(dist, axis=1, kind="stable")
        idx[:, h] = order[:, 0]
        srt = np.take_along_axis(dist, order[:, :2], axis=1)
        gap[:, h] = srt[:, 1] - srt[:, 0]
    return idx, gap
