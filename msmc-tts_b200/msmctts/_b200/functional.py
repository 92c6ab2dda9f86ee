"""Autograd glue over the C-ABI kernels (host code stays Python/PyTorch; every op below is a hand-written
sm_100a kernel in csrc/).  Activations are channels-last: (B, H, W, C) with H == 1 for sequences.
"""
import ctypes as C
import os
from collections import namedtuple

import torch

from . import lib as L

# Contraction math of the conv / linear kernels:
#   "3xtf32" (default) tcgen05 tensor cores, fp32 operands split hi/lo -> fp32-accurate results
#   "tf32"             tcgen05 tensor cores, plain TF32 (the reference's cuDNN default on Ampere+)
#   "fp32"             CUDA-core FMA kernels only
CONV_MATH = os.environ.get("MSMC_CONV_MATH", "3xtf32")
UMMA_MIN_ROWS = 256
# fewest source channels routed to the tensor-core conv kernels.  The kernels take >= 8 (csrc UM_MIN_CS: one ragged,
# zero-filled 32-channel chunk) and are parity-tested there, but measured on the train step the CUDA-core kernels win
# below 32: 37.4 ms/step at 32, 38.1 at 16, 40.1 at 8 (profiles/README.md) -- those layers are all position tiles
# and no reduction, so the per-tile fixed cost of the tensor-core kernels dominates
UMMA_MIN_CS = int(os.environ.get("MSMC_UMMA_MIN_CS", "32"))
USE_TAP_REUSE = os.environ.get("MSMC_TAP_REUSE", "1") != "0"
# VQ search kernel choice: "auto" = the two-phase tensor-core kernel (csrc/vq_umma.cu, bit-identical results) from
# VQ_UMMA_MIN_ROWS rows on -- below that a launch is one row tile per CTA and the CUDA-core cluster kernel's shorter
# prologue wins (profiles/r02_bench_vq_*.txt) --; True / "1" forces it, False / "0" disables it
VQ_UMMA = {"0": False, "1": True}.get(os.environ.get("MSMC_VQ_UMMA", "auto"), "auto")
VQ_UMMA_MIN_ROWS = int(os.environ.get("MSMC_VQ_UMMA_MIN_ROWS", "2048"))

ConvCfg = namedtuple("ConvCfg", "KH KW sh sw dh dw ph pw reflect transposed wstr Cd pre_slope post out_hw")
# wstr = element strides of the weight tensor for (kh, kw, cs, cd), cs = channels of the op's INPUT


# ------------------------------------------------------------------------------------------------ helpers
def _rows(x):
    """channels-last 4-D tensor with unit channel stride and packed rows; returns (tensor, row_pitch)."""
    assert x.dim() == 4
    B, H, W, Cn = x.shape
    ok = (Cn == 1 or x.stride(3) == 1)
    ld, expect = None, None
    for n, s in ((W, x.stride(2)), (H, x.stride(1)), (B, x.stride(0))):
        if n == 1:
            continue
        if ld is None:
            ld, expect = s, s * n
        else:
            ok = ok and s == expect
            expect = s * n
    if ld is None:
        ld = Cn
    if not ok or ld < Cn:
        x = x.contiguous()
        ld = Cn
    return x, ld


def conv_out_size(n, k, s, d, p, transposed):
    if transposed:
        return (n - 1) * s - 2 * p + d * (k - 1) + 1
    return (n + 2 * p - d * (k - 1) - 1) // s + 1


def _weight_image(w, T, Cs, Cd, role, split, bn):
    """swizzled tcgen05 operand image of a GEMM-layout weight [T][Cs][Cd]; cached on the tensor for its lifetime"""
    cache = getattr(w, "_msmc_img", None)
    if cache is None:
        cache = {}
        w._msmc_img = cache
    key = (role, split, bn)
    owner = getattr(w, "_msmc_owner", None)
    if owner is not None and not PREFETCHING[0]:
        keys = owner.__dict__.setdefault("_msmc_img_keys", set())
        keys.add((T, Cs, Cd, role, split, bn))
    img = cache.get(key)
    if img is None:
        n = L.load().msmc_weight_image_elems(T, Cs, Cd, role, split, bn)
        img = torch.empty(n, dtype=torch.float32, device=w.device)
        if DEFERRED[0] is not None:
            # prefetch_weights batches every image of the sub-network into one msmc_weight_image_multi launch
            DEFERRED[0]["img"].append((w, img, T, Cs, Cd, bn, role, split, n // (2 if split else 1)))
        else:
            L.call("msmc_weight_image", L.ptr(w), L.ptr(img), T, Cs, Cd, role, split, bn)
        cache[key] = img
    return img


# ---- deferred multi-tensor weight preparation (layers.prefetch_weights): job lists + pointer-table staging
DEFERRED = [None]
_prep_staging = {}


def _job_table(key, words, device):
    """int64 job table -> device through pinned staging (graph-capturable H2D copy).  Same discipline as the other
    multi-tensor pointer tables: a table used inside a CUDA-graph capture gets a PRIVATE staging pair (taken from
    spares allocated by an earlier eager call), because the captured copy re-reads the pinned buffer on every
    replay; an eager call waits for its previous copy out of the pinned buffer before rewriting it."""
    capturing = torch.cuda.is_current_stream_capturing()
    n = len(words)
    st = _prep_staging.get(key)
    if st is None or st["n"] != n:
        if capturing:
            raise L.MsmcError("weight prefetch: run one eager step with this model before capturing a CUDA graph")
        st = {"n": n, "pair": _staging(n, device), "spares": [_staging(n, device) for _ in range(3)],
              "event": None, "graph_pairs": []}
        _prep_staging[key] = st
    if capturing:
        if not st["spares"]:
            raise L.MsmcError("weight prefetch: no private staging buffer left for another CUDA-graph capture")
        host, table = st["spares"].pop()
        st["graph_pairs"].append((host, table))
    else:
        if st["event"] is not None:
            st["event"].synchronize()
        host, table = st["pair"]
    host.copy_(torch.tensor(words, dtype=torch.int64))
    table.copy_(host, non_blocking=True)
    if not capturing:
        st["event"] = torch.cuda.Event()
        st["event"].record()
    return table


def flush_deferred(jobs, key):
    """launch the batched re-parametrisation and the batched operand images recorded in `jobs`"""
    wn, img = jobs["wn"], jobs["img"]
    if wn:
        words, row0 = [], 0
        for (v, g, w, inv, O, I, J, so, si, sj) in wn:
            words += [v.data_ptr(), g.data_ptr() if g is not None else 0, w.data_ptr(),
                      inv.data_ptr() if inv is not None else 0, so, si, sj, O, I, J, row0]
            row0 += O
        table = _job_table((key, "wn"), words, wn[0][0].device)
        L.call("msmc_weight_norm_fwd_multi", L.ptr(table), len(wn), C.c_int64(row0))
    if img:
        words, blk0 = [], 0
        for (w, im, T, Cs, Cd, bn, role, split, elems) in img:
            words += [w.data_ptr(), im.data_ptr(), T, Cs, Cd, bn, role, split, blk0]
            blk0 += (elems + 1023) // 1024
        table = _job_table((key, "img"), words, img[0][0].device)
        L.call("msmc_weight_image_multi", L.ptr(table), len(img), C.c_int64(blk0))


def _umma_ok(src, ld_src, w, wstr_gemm, KH, KW, Cs, Cd_gemm, rows, saux, ld_saux):
    if CONV_MATH == "fp32" or Cs % 4 != 0 or Cs < UMMA_MIN_CS or rows < UMMA_MIN_ROWS:      # a ragged last 32-channel chunk is fine
        return False
    if not w.is_contiguous() or w.numel() != KH * KW * wstr_gemm[0] * wstr_gemm[1]:
        return False
    if ld_src % 4 != 0 or src.data_ptr() % 16 != 0:
        return False
    if saux is not None and (ld_saux % 4 != 0 or saux.data_ptr() % 16 != 0):
        return False
    return True


def _launch_conv(src, w, wstr, bias, residual, dst, KH, KW, sh, sw, dh, dw, ph, pw, reflect, transposed,
                 src_xf=(L.XF_NONE, 0.0, None), dst_xf=(L.XF_NONE, 0.0, None), w_role=0, w_dims=None):
    """w_role/w_dims: when the caller runs a stride-1 data gradient as a forward-form conv (taps reversed, channels
    swapped) it passes the ORIGINAL GEMM-layout weight with w_role=1 and w_dims=(Cs_op, Cd_op); only the
    tensor-core path can consume that, so the caller must have checked eligibility."""
    src, ld_src = _rows(src)
    B, Hs, Ws, Cs = src.shape
    _, Hd, Wd, Cd = dst.shape
    assert dst.is_contiguous()
    g = L.ConvGeom()
    g.B, g.Hs, g.Ws, g.Cs, g.Hd, g.Wd, g.Cd = B, Hs, Ws, Cs, Hd, Wd, Cd
    g.KH, g.KW, g.sh, g.sw, g.dh, g.dw, g.ph, g.pw = KH, KW, sh, sw, dh, dw, ph, pw
    g.pad_reflect, g.transposed = int(reflect), int(transposed)
    g.ld_src, g.ld_dst = ld_src, Cd
    res = None
    if residual is not None:
        res, g.ld_res = _rows(residual)
    saux = daux = None
    if src_xf[2] is not None:
        saux, g.ld_saux = _rows(src_xf[2])
    if dst_xf[2] is not None:
        daux, g.ld_daux = _rows(dst_xf[2])
    g.ws_kh, g.ws_kw, g.ws_cs, g.ws_cd = wstr
    g.src_xf, g.src_slope = src_xf[0], float(src_xf[1])
    g.dst_xf, g.dst_slope = dst_xf[0], float(dst_xf[1])
    L.require_cuda(src, w, dst)
    meta = None
    if L._profile is not None:
        pos = B * (Hs * Ws if transposed else Hd * Wd)
        meta = {"flops": 2.0 * pos * KH * KW * Cs * Cd,
                "bytes": 4.0 * (src.numel() + dst.numel() + KH * KW * Cs * Cd + (res.numel() if res is not None else 0)),
                "shape": "B%d %dx%d C%d->%d k%dx%d s%d%s" % (B, Hs, Ws, Cs, Cd, KH, KW, sw, "T" if transposed else "")}
    gemm_contig = tuple(wstr) == (KW * Cs * Cd, Cs * Cd, Cd, 1) and not (transposed and reflect)
    if w_role != 0 or (gemm_contig and _umma_ok(src, ld_src, w, (Cs, Cd), KH, KW, Cs, Cd, B * Hd * Wd, saux,
                                                  g.ld_saux)):
        split = 0 if CONV_MATH == "tf32" else 1
        cs_op, cd_op = w_dims if w_role != 0 else (Cs, Cd)
        bn = L.load().msmc_umma_tile_n(Cd, B * Hd * Wd)
        img = _weight_image(w, KH * KW, cs_op, cd_op, w_role, split, bn)
        name = "msmc_conv_forward_umma"
        if USE_TAP_REUSE and L.load().msmc_conv_reuse_eligible(C.byref(g)):
            name = "msmc_conv_forward_umma_reuse"     # stride-1: one staged operand tile serves every tap
        L.call(name, C.byref(g), L.ptr(src), L.ptr(saux), L.ptr(img), L.ptr(bias), L.ptr(res), L.ptr(daux),
               L.ptr(dst), split, bn, meta=meta)
        return dst
    L.call("msmc_conv_forward", C.byref(g), L.ptr(src), L.ptr(saux), L.ptr(w), L.ptr(bias), L.ptr(res),
           L.ptr(daux), L.ptr(dst), meta=meta)
    return dst


def _launch_wgrad(src, gout, dw, wstr, dbias, KH, KW, sh, sw, dh, dw_, ph, pw, reflect,
                  src_xf=(L.XF_NONE, 0.0, None), gout_xf=(L.XF_NONE, 0.0, None)):
    """dW[kh,kw,cs,cd] = sum_rows xf(src)[gathered] * xf(gout);  rows = positions of gout (forward form)."""
    src, ld_src = _rows(src)
    gout, ld_g = _rows(gout)
    B, Hs, Ws, Cs = src.shape
    _, Hd, Wd, Cd = gout.shape
    g = L.ConvGeom()
    g.B, g.Hs, g.Ws, g.Cs, g.Hd, g.Wd, g.Cd = B, Hs, Ws, Cs, Hd, Wd, Cd
    g.KH, g.KW, g.sh, g.sw, g.dh, g.dw, g.ph, g.pw = KH, KW, sh, sw, dh, dw_, ph, pw
    g.pad_reflect, g.transposed = int(reflect), 0
    g.ld_src, g.ld_dst = ld_src, ld_g
    saux = daux = None
    if src_xf[2] is not None:
        saux, g.ld_saux = _rows(src_xf[2])
    if gout_xf[2] is not None:
        daux, g.ld_daux = _rows(gout_xf[2])
    g.ws_kh, g.ws_kw, g.ws_cs, g.ws_cd = wstr
    g.src_xf, g.src_slope = src_xf[0], float(src_xf[1])
    g.dst_xf, g.dst_slope = gout_xf[0], float(gout_xf[1])
    lib = L.load()
    meta = None
    if L._profile is not None:
        meta = {"flops": 2.0 * B * Hd * Wd * KH * KW * Cs * Cd,
                "bytes": 4.0 * (src.numel() + gout.numel() + KH * KW * Cs * Cd),
                "shape": "wgrad B%d %dx%d C%d->%d k%dx%d" % (B, Hs, Ws, Cs, Cd, KH, KW)}
    use_umma = (CONV_MATH != "fp32" and Cs % 4 == 0 and Cs >= UMMA_MIN_CS and ld_src % 4 == 0 and src.data_ptr() % 16 == 0
                and B * Hd * Wd >= UMMA_MIN_ROWS
                and (saux is None or (g.ld_saux % 4 == 0 and saux.data_ptr() % 16 == 0)))
    if use_umma:
        nbytes = lib.msmc_conv_wgrad_umma_workspace(C.byref(g))
        ws = torch.empty(max(nbytes // 4, 1), dtype=torch.float32, device=src.device)
        L.call("msmc_conv_wgrad_umma", C.byref(g), L.ptr(src), L.ptr(saux), L.ptr(gout), L.ptr(daux), L.ptr(dw),
               L.ptr(dbias), L.ptr(ws), C.c_int64(nbytes), 0 if CONV_MATH == "tf32" else 1, meta=meta)
        return
    nbytes = lib.msmc_conv_wgrad_workspace(C.byref(g))
    ws = torch.empty(max(nbytes // 4, 1), dtype=torch.float32, device=src.device)
    L.call("msmc_conv_wgrad", C.byref(g), L.ptr(src), L.ptr(saux), L.ptr(gout), L.ptr(daux), L.ptr(dw),
           L.ptr(dbias), L.ptr(ws), C.c_int64(nbytes), meta=meta)


# post = (kind, slope): result transform fused in the conv epilogue, and the operand modifier its backward needs
_POST_FWD = {"none": L.XF_NONE, "relu": L.XF_RELU, "tanh": L.XF_TANH, "lrelu": L.XF_LRELU}
_POST_BWD = {"none": L.XF_NONE, "relu": L.XF_MUL_DRELU, "tanh": L.XF_MUL_DTANH, "lrelu": L.XF_MUL_DLRELU}


class _ConvFn(torch.autograd.Function):
    """y = post(conv(pre(x), w) + bias) + residual, forward form or conv-transpose form."""

    @staticmethod
    def forward(ctx, x, w, bias, residual, cfg):
        x, _ = _rows(x)
        B, Hs, Ws, Cs = x.shape
        if cfg.out_hw is not None:
            Hd, Wd = cfg.out_hw
        else:
            Hd = conv_out_size(Hs, cfg.KH, cfg.sh, cfg.dh, cfg.ph, cfg.transposed)
            Wd = conv_out_size(Ws, cfg.KW, cfg.sw, cfg.dw, cfg.pw, cfg.transposed)
        y = torch.empty((B, Hd, Wd, cfg.Cd), dtype=torch.float32, device=x.device)
        pre = (L.XF_LRELU, cfg.pre_slope, None) if cfg.pre_slope is not None else (L.XF_NONE, 0.0, None)
        kind, pslope = cfg.post
        assert not (kind != "none" and residual is not None)
        _launch_conv(x, w, cfg.wstr, bias, residual, y, cfg.KH, cfg.KW, cfg.sh, cfg.sw, cfg.dh, cfg.dw, cfg.ph,
                     cfg.pw, cfg.reflect, cfg.transposed, src_xf=pre, dst_xf=(_POST_FWD[kind], pslope, None))
        ctx.cfg = cfg
        ctx.has_bias = bias is not None
        ctx.has_res = residual is not None
        ctx.save_for_backward(x, w, y if kind != "none" else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        cfg = ctx.cfg
        x, w, y = ctx.saved_tensors
        gy = gy.contiguous()
        kind, pslope = cfg.post
        # (the sign of a leaky-relu output equals the sign of its input, so y serves as the aux tensor)
        gmod = (_POST_BWD[kind], pslope, y if kind != "none" else None)
        if kind != "none" and PREMASK_GRAD and gy.is_cuda and y.is_contiguous() and gy.data_ptr() % 16 == 0:
            # pre-activation gradient g * act'(y) formed ONCE by a vectorised pass: the data-gradient and the
            # weight-gradient kernels then read a plain operand (their aux-reading variants are 2-4x slower)
            gpre = torch.empty_like(gy)
            L.call("msmc_xform_apply", L.ptr(gy), L.ptr(y), L.ptr(gpre), C.c_int64(gy.numel()),
                   _POST_BWD[kind], C.c_float(pslope))
            gy, kind = gpre, "none"
            gmod = (L.XF_NONE, 0.0, None)
        post_eff = (kind, pslope)
        pre = (L.XF_LRELU, cfg.pre_slope, None) if cfg.pre_slope is not None else (L.XF_NONE, 0.0, None)
        s_kh, s_kw, s_cs, s_cd = cfg.wstr
        gx = gw = gb = None
        # The weight gradient is off the critical path (only the optimizer reads it) while the data gradient feeds
        # the next layer's backward: enqueue wgrad on a side stream first, the data gradient on the current stream,
        # and join before returning.  Outputs are allocated on the current stream (allocator ownership).
        side = None
        if ctx.needs_input_grad[1]:
            gw = torch.empty_like(w) if w.is_contiguous() else torch.zeros_like(w)
            want_b = ctx.has_bias and ctx.needs_input_grad[2]
            if want_b and not cfg.transposed:
                gb = torch.empty(cfg.Cd, dtype=torch.float32, device=x.device)
            cur = None
            if WGRAD_STREAM and ctx.needs_input_grad[0] and x.is_cuda:
                cur, side = _wgrad_side_stream()
                side.wait_stream(cur)
            with torch.cuda.stream(side) if side is not None else _NullCtx():
                if not cfg.transposed:
                    _launch_wgrad(x, gy, gw, cfg.wstr, gb, cfg.KH, cfg.KW, cfg.sh, cfg.sw, cfg.dh, cfg.dw, cfg.ph,
                                  cfg.pw, cfg.reflect, src_xf=pre, gout_xf=gmod)
                else:
                    # roles swap: the op's output plays the gathered source, the op's input plays "gout"
                    _launch_wgrad(gy, x, gw, (s_kh, s_kw, s_cd, s_cs), None, cfg.KH, cfg.KW, cfg.sh, cfg.sw, cfg.dh,
                                  cfg.dw, cfg.ph, cfg.pw, False, src_xf=gmod, gout_xf=pre)
            if want_b and cfg.transposed:
                gb = _post_grad(gy, y, post_eff).sum(dim=(0, 1, 2))
        elif ctx.has_bias and ctx.needs_input_grad[2]:
            gb = _post_grad(gy, y, post_eff).sum(dim=(0, 1, 2))
        if ctx.needs_input_grad[0]:
            dmod = (L.XF_MUL_DLRELU, cfg.pre_slope, x) if cfg.pre_slope is not None else (L.XF_NONE, 0.0, None)
            if cfg.reflect:
                # gradient w.r.t. the reflect-padded input, then fold the borders back
                B, Hs, Ws, Cs = x.shape
                gpad = torch.empty((B, Hs + 2 * cfg.ph, Ws + 2 * cfg.pw, Cs), dtype=torch.float32, device=x.device)
                gy_r, ld_gy = _rows(gy)
                aux_r, ld_aux = _rows(gmod[2]) if gmod[2] is not None else (None, 0)
                if (cfg.sh == 1 and cfg.sw == 1 and cfg.wstr == (cfg.KW * Cs * cfg.Cd, Cs * cfg.Cd, cfg.Cd, 1)
                        and _umma_ok(gy_r, ld_gy, w, (Cs, cfg.Cd), cfg.KH, cfg.KW, cfg.Cd, Cs,
                                     gpad.shape[0] * gpad.shape[1] * gpad.shape[2], aux_r, ld_aux)):
                    # stride 1: full correlation with reversed taps on the tensor cores
                    _launch_conv(gy, w, (cfg.KW * cfg.Cd * Cs, cfg.Cd * Cs, Cs, 1), None, None, gpad, cfg.KH, cfg.KW,
                                 1, 1, cfg.dh, cfg.dw, cfg.dh * (cfg.KH - 1), cfg.dw * (cfg.KW - 1), False, False,
                                 src_xf=gmod, w_role=1, w_dims=(Cs, cfg.Cd))
                elif (cfg.wstr == (cfg.KW * Cs * cfg.Cd, Cs * cfg.Cd, cfg.Cd, 1)
                      and _umma_ok(gy_r, ld_gy, w, (Cs, cfg.Cd), cfg.KH, cfg.KW, cfg.Cd, Cs,
                                   gpad.shape[0] * gpad.shape[1] * gpad.shape[2], aux_r, ld_aux)):
                    # strided: conv-transpose form on the tensor cores (phase-decomposed)
                    _launch_conv(gy, w, (s_kh, s_kw, s_cd, s_cs), None, None, gpad, cfg.KH, cfg.KW, cfg.sh, cfg.sw,
                                 cfg.dh, cfg.dw, 0, 0, False, True, src_xf=gmod, w_role=2, w_dims=(Cs, cfg.Cd))
                else:
                    _launch_conv(gy, w, (s_kh, s_kw, s_cd, s_cs), None, None, gpad, cfg.KH, cfg.KW, cfg.sh, cfg.sw,
                                 cfg.dh, cfg.dw, 0, 0, False, True, src_xf=gmod)
                gx = torch.empty_like(x)
                L.call("msmc_reflect_pad_fold", L.ptr(gpad), L.ptr(gx), B, Hs, Ws, Cs, cfg.ph, cfg.pw)
                if cfg.pre_slope is not None:
                    gx = torch.where(x > 0, gx, gx * cfg.pre_slope)
            else:
                gx = torch.empty_like(x)
                Cs_op = x.shape[-1]
                gy_r, ld_gy = _rows(gy)
                aux_r, ld_aux = _rows(gmod[2]) if gmod[2] is not None else (None, 0)
                if (not cfg.transposed and cfg.sh == 1 and cfg.sw == 1
                        and cfg.wstr == (cfg.KW * Cs_op * cfg.Cd, Cs_op * cfg.Cd, cfg.Cd, 1)
                        and _umma_ok(gy_r, ld_gy, w, (Cs_op, cfg.Cd), cfg.KH, cfg.KW, cfg.Cd, Cs_op,
                                     gx.shape[0] * gx.shape[1] * gx.shape[2], aux_r, ld_aux)):
                    # stride-1 data gradient == forward-form conv with reversed taps: tensor-core path
                    _launch_conv(gy, w, (cfg.KW * cfg.Cd * Cs_op, cfg.Cd * Cs_op, Cs_op, 1), None, None, gx, cfg.KH,
                                 cfg.KW, 1, 1, cfg.dh, cfg.dw, cfg.dh * (cfg.KH - 1) - cfg.ph,
                                 cfg.dw * (cfg.KW - 1) - cfg.pw, False, False, src_xf=gmod, dst_xf=dmod, w_role=1,
                                 w_dims=(Cs_op, cfg.Cd))
                elif (cfg.wstr == (cfg.KW * Cs_op * cfg.Cd, Cs_op * cfg.Cd, cfg.Cd, 1)
                      and _umma_ok(gy_r, ld_gy, w, (Cs_op, cfg.Cd), cfg.KH, cfg.KW, cfg.Cd, Cs_op,
                                   gx.shape[0] * gx.shape[1] * gx.shape[2], aux_r, ld_aux)):
                    # strided conv (-> conv-transpose form) or conv-transpose op (-> forward form): same weight
                    # with channels swapped and taps in place (image role 2)
                    _launch_conv(gy, w, (s_kh, s_kw, s_cd, s_cs), None, None, gx, cfg.KH, cfg.KW, cfg.sh, cfg.sw,
                                 cfg.dh, cfg.dw, cfg.ph, cfg.pw, False, not cfg.transposed, src_xf=gmod,
                                 dst_xf=dmod, w_role=2, w_dims=(Cs_op, cfg.Cd))
                else:
                    _launch_conv(gy, w, (s_kh, s_kw, s_cd, s_cs), None, None, gx, cfg.KH, cfg.KW, cfg.sh, cfg.sw,
                                 cfg.dh, cfg.dw, cfg.ph, cfg.pw, False, not cfg.transposed, src_xf=gmod,
                                 dst_xf=dmod)
        if side is not None:
            cur.wait_stream(side)
        gres = gy if (ctx.has_res and ctx.needs_input_grad[3]) else None
        return gx, gw, gb, gres, None


WGRAD_STREAM = os.environ.get("MSMC_WGRAD_STREAM", "1") != "0"
PREMASK_GRAD = os.environ.get("MSMC_PREMASK_GRAD", "1") != "0"
_wgrad_streams = {}


class _NullCtx(object):
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


def _wgrad_side_stream():
    """(current stream, its dedicated weight-gradient side stream)"""
    cur = torch.cuda.current_stream()
    key = (cur.device.index, cur.cuda_stream)
    s = _wgrad_streams.get(key)
    if s is None:
        s = _wgrad_streams[key] = torch.cuda.Stream(device=cur.device)
    return cur, s


def _post_grad(gy, y, post):
    kind, slope = post
    if kind == "relu":
        return gy * (y > 0)
    if kind == "tanh":
        return gy * (1 - y * y)
    if kind == "lrelu":
        return torch.where(y > 0, gy, gy * slope)
    return gy


def conv_cl(x, w, bias=None, residual=None, *, kernel=(1, 1), stride=(1, 1), dilation=(1, 1), padding=(0, 0),
            reflect=False, transposed=False, wstr=None, out_channels=None, pre_slope=None, post="none",
            out_hw=None):
    """Channels-last convolution-shaped contraction.  `w` is any tensor whose element for (kh, kw, cs, cd) sits
    at the strides `wstr` (default: the GEMM layout [KH][KW][Cs][Cd], contiguous)."""
    KH, KW = kernel
    Cs = x.shape[-1]
    if wstr is None:
        Cd = w.shape[-1]
        wstr = (KW * Cs * Cd, Cs * Cd, Cd, 1)
    else:
        Cd = out_channels
    if isinstance(post, str):
        post = (post, 0.0)
    cfg = ConvCfg(KH, KW, stride[0], stride[1], dilation[0], dilation[1], padding[0], padding[1], bool(reflect),
                  bool(transposed), tuple(int(s) for s in wstr), int(Cd), pre_slope, (post[0], float(post[1])),
                  out_hw)
    return _ConvFn.apply(x, w, bias, residual, cfg)


def linear_cl(x, weight, bias=None, residual=None, post="none", pre_slope=None, owner=None):
    """x (..., Ci) @ weight(Co, Ci)^T + bias, weight consumed in torch's native nn.Linear / 1x1-conv layout.
    `owner` (the layer module) lets the re-laid-out weight be shared / prefetched per parameter version."""
    Ci = x.shape[-1]
    Co = weight.shape[0]
    lead = x.shape[:-1]
    x4 = x.reshape(-1, 1, 1, Ci)
    r4 = residual.reshape(-1, 1, 1, Co) if residual is not None else None
    umma = CONV_MATH != "fp32" and Ci % 4 == 0 and Ci >= UMMA_MIN_CS and x4.shape[0] >= UMMA_MIN_ROWS and weight.requires_grad
    if owner is not None:
        owner.__dict__["_msmc_lin_umma"] = bool(umma)
    if umma:
        # tensor-core path wants the GEMM layout [1][Ci][Co]; the re-layout is one tiny launch
        if owner is not None:
            from .layers import _prepped
            w_g = _prepped(owner, weight.reshape(Co, Ci, 1), None)
        else:
            w_g = prep_conv_weight(weight.reshape(Co, Ci, 1))
        y = conv_cl(x4, w_g, bias, r4, post=post, pre_slope=pre_slope)
    else:
        y = conv_cl(x4, weight.reshape(Co, Ci), bias, r4, wstr=(0, 0, 1, Ci), out_channels=Co, post=post,
                    pre_slope=pre_slope)
    return y.reshape(*lead, Co)


# ------------------------------------------------------------------------------------------------- STFT
class _StftFn(torch.autograd.Function):
    """reflect-padded, windowed DFT of a waveform as a strided conv with a fixed basis; the backward is a dense
    GEMM (spectrum gradient x basis^T -> per-frame time gradients, tensor-core eligible because the spectrum
    channels are padded to a multiple of 32) followed by overlap-add + reflect fold, instead of a 1-output-channel
    conv-transpose over hop phases."""

    @staticmethod
    def forward(ctx, x, basis, basis_t, hop, pad, win):
        B, Ln = x.shape
        win_p, F2 = basis.shape                  # basis rows are zero-padded from win to a multiple of 32
        frames = (Ln + 2 * pad - win) // hop + 1
        x = x.contiguous()
        if CONV_MATH != "fp32" and B * frames >= UMMA_MIN_ROWS and win_p % 32 == 0:
            # unfold (a few MB) + one tensor-core GEMM over the frames
            fr = torch.empty((B * frames, 1, 1, win_p), dtype=torch.float32, device=x.device)
            L.require_cuda(x)
            L.call("msmc_frame_unfold", L.ptr(x), L.ptr(fr), B, Ln, frames, win, win_p, hop, pad)
            y = torch.empty((B * frames, 1, 1, F2), dtype=torch.float32, device=x.device)
            _launch_conv(fr, basis, (win_p * F2, win_p * F2, F2, 1), None, None, y, 1, 1, 1, 1, 1, 1, 0, 0, False,
                         False)
        else:
            y = torch.empty((B, 1, frames, F2), dtype=torch.float32, device=x.device)
            _launch_conv(x.reshape(B, 1, Ln, 1), basis, (0, F2, F2, 1), None, None, y, 1, win, 1, hop, 1, 1, 0, pad,
                         True, False)
        ctx.save_for_backward(basis_t)
        ctx.dims = (B, Ln, win, F2, frames, hop, pad)
        return y.reshape(B, frames, F2)

    @staticmethod
    def backward(ctx, gy):
        (basis_t,) = ctx.saved_tensors          # (F2, win): GEMM layout [1][Cs = F2][Cd = win]
        B, Ln, win, F2, frames, hop, pad = ctx.dims
        gy = gy.contiguous()
        gframes = torch.empty((B * frames, 1, 1, win), dtype=torch.float32, device=gy.device)
        _launch_conv(gy.reshape(B * frames, 1, 1, F2), basis_t, (F2 * win, F2 * win, win, 1), None, None, gframes,
                     1, 1, 1, 1, 1, 1, 0, 0, False, False)
        gx = torch.empty((B, Ln), dtype=torch.float32, device=gy.device)
        L.call("msmc_overlap_add_fold", L.ptr(gframes), L.ptr(gx), B, frames, win, hop, Ln, pad)
        return gx, None, None, None, None, None


def stft_frames(x, basis, basis_t, hop, pad, win):
    """x (B, L); basis (win_p, 2Fp) = hann * [cos | 0 | -sin | 0] with rows >= win zero; basis_t (2Fp, win) the
    transpose of its first win rows -> (B, frames, 2Fp)"""
    return _StftFn.apply(x, basis, basis_t, int(hop), int(pad), int(win))


# ------------------------------------------------------------------------------------------ weight prep
class _PrepWeightFn(torch.autograd.Function):
    """(v[, g]) -> GEMM-layout weight; weight_norm (dim=0) when g is given."""

    @staticmethod
    def forward(ctx, v, g, O, I, J, so, si, sj, out_shape, token):
        ctx.token = token
        v = v.contiguous()
        w = torch.empty(out_shape, dtype=torch.float32, device=v.device)
        inv = torch.empty(O, dtype=torch.float32, device=v.device) if g is not None else None
        L.require_cuda(v, g)
        meta = {"flops": 0.0, "bytes": 8.0 * O * I * J, "shape": "O%d I%d J%d %s" % (O, I, J, "wn" if g is not None else "relayout")} \
            if L._profile is not None else None
        if DEFERRED[0] is not None:
            DEFERRED[0]["wn"].append((v, g, w, inv, O, I, J, so, si, sj))     # one batched launch later
        else:
            L.call("msmc_weight_norm_fwd", L.ptr(v), L.ptr(g), L.ptr(w), L.ptr(inv), O, I, J,
                   C.c_int64(so), C.c_int64(si), C.c_int64(sj), meta=meta)
        ctx.dims = (O, I, J, so, si, sj)
        ctx.has_g = g is not None
        ctx.save_for_backward(v, g, inv)
        return w

    @staticmethod
    def backward(ctx, gw):
        v, g, inv = ctx.saved_tensors
        if ctx.token is not None:
            ctx.token["spent"] = True        # the layer-level cache must not hand this node out again
        O, I, J, so, si, sj = ctx.dims
        gw = gw.contiguous()
        dv = torch.empty_like(v)
        dg = torch.empty(O, dtype=torch.float32, device=v.device) if ctx.has_g else None
        meta = {"flops": 0.0, "bytes": 8.0 * O * I * J, "shape": "O%d I%d J%d %s" % (O, I, J, "wn" if ctx.has_g else "relayout")} \
            if L._profile is not None else None
        L.call("msmc_weight_norm_bwd", L.ptr(gw), C.c_int64(so), C.c_int64(si), C.c_int64(sj), L.ptr(v), L.ptr(g),
               L.ptr(inv), L.ptr(dv), L.ptr(dg), O, I, J, meta=meta)
        if ctx.has_g:
            dg = dg.reshape(g.shape)
        return dv, dg, None, None, None, None, None, None, None, None


# ------------------------------------------------------------------------------------------ stream forks
# Independent sub-networks (the 3 + 5 sub-discriminators, the parallel MRF blocks of one generator stage) are
# enqueued on forked CUDA streams and joined afterwards: most of their kernels fill only part of the 148 SMs, so
# running branches side by side hides tails, prologues and launch gaps.  autograd replays each branch's backward
# on the stream its forward ran on, and a CUDA-graph capture records the fork/join as parallel graph branches.
BRANCH_STREAMS = os.environ.get("MSMC_BRANCH_STREAMS", "1") != "0"
_side_streams = {}
_in_branch = [False]


def run_branches(fns):
    """[f() for f in fns] with every f on its own side stream (sequential when disabled, nested or on CPU)"""
    if not BRANCH_STREAMS or len(fns) < 2 or _in_branch[0] or not torch.cuda.is_available():
        return [f() for f in fns]
    main = torch.cuda.current_stream()
    pool = _side_streams.setdefault(main.device, [])
    while len(pool) < len(fns):
        pool.append(torch.cuda.Stream(device=main.device, priority=MAIN_PRIORITY))
    outs = []
    _in_branch[0] = True
    try:
        for f, s in zip(fns, pool):
            s.wait_stream(main)
            with torch.cuda.stream(s):
                outs.append(f())
    finally:
        _in_branch[0] = False
    for s in pool[:len(fns)]:
        main.wait_stream(s)
    return outs


# weight prefetch (layers.prefetch_weights): side stream per consumer stream, re-entrancy flag
PREFETCH_WEIGHTS = os.environ.get("MSMC_PREFETCH_WEIGHTS", "1") != "0"
PREFETCHING = [False]
# stream priorities (lower = more urgent): critical-path streams (graph capture stream, sub-network branches) vs the
# weight-gradient side streams (always 0, the least urgent) vs the prefetch stream
MAIN_PRIORITY = int(os.environ.get("MSMC_MAIN_PRIORITY", "0"))
PREFETCH_PRIORITY = int(os.environ.get("MSMC_PREFETCH_PRIORITY", str(MAIN_PRIORITY - 1)))
_prefetch_streams = {}


def prefetch_stream(cur):
    key = (cur.device.index, cur.cuda_stream)
    s = _prefetch_streams.get(key)
    if s is None:
        # high priority: the prefetched kernels are tiny and the forward waits on them, so when both are runnable
        # the block scheduler should take them before the next wave of a long convolution
        s = _prefetch_streams[key] = torch.cuda.Stream(device=cur.device, priority=PREFETCH_PRIORITY)
    return s


# id of the trainer step in flight (0 = none): layers share one re-parametrised weight per parameter version inside it
PREP_SCOPE = [0]
_prep_scope_ids = [0]


class prep_scope(object):
    """`with prep_scope():` -- layers called more than once on unchanged parameters reuse their prepared weight"""

    def __enter__(self):
        _prep_scope_ids[0] += 1
        self.prev = PREP_SCOPE[0]
        PREP_SCOPE[0] = _prep_scope_ids[0]

    def __exit__(self, *exc):
        PREP_SCOPE[0] = self.prev
        return False


def prep_conv_weight(v, g=None, transposed=False, token=None):
    """torch conv weight -> GEMM layout [KH][KW][Cs][Cd] (contiguous).
    Conv:          v (Co, Ci, KH, KW) or (Co, Ci, K);  weight_norm over (Ci, K...) per Co.
    ConvTranspose: v (Cin, Cout, K);                    weight_norm over (Cout, K) per Cin (torch dim=0)."""
    if v.dim() == 3:
        O, I, K = v.shape
        KH, KW = 1, K
    else:
        O, I, KH, KW = v.shape
    J = KH * KW
    if not transposed:
        # o = cd, i = cs:  offset = j*(I*O) + i*O + o
        shape, so, si, sj = (KH, KW, I, O), 1, O, I * O
    else:
        # o = cs (Cin), i = cd (Cout): offset = j*(O*I) + o*I + i
        shape, so, si, sj = (KH, KW, O, I), I, 1, O * I
    return _PrepWeightFn.apply(v, g, O, I, J, so, si, sj, shape, token)


# ------------------------------------------------------------------------------------------------ VQ
class _VQFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, embed, n_heads, dim):
        z2 = z.reshape(-1, n_heads * dim)
        if z2.stride(-1) != 1:
            z2 = z2.contiguous()
        n_rows = z2.shape[0]
        K = embed.shape[-1]
        L.require_cuda(z2, embed)
        q_raw = torch.empty((n_rows, n_heads * dim), dtype=torch.float32, device=z.device)
        q_st = torch.empty_like(q_raw)
        diff = torch.empty((n_rows, dim), dtype=torch.float32, device=z.device)
        idx = torch.empty((n_rows, n_heads), dtype=torch.int64, device=z.device)
        entry = "msmc_vq_search"
        use_umma = VQ_UMMA is True or (VQ_UMMA == "auto" and n_rows >= VQ_UMMA_MIN_ROWS)
        if use_umma and dim == 64 and K in (64, 128, 256) and n_heads in (1, 2, 4, 8) and z2.stride(0) % 4 == 0 \
                and z2.data_ptr() % 16 == 0:
            entry = "msmc_vq_search_umma"      # two-phase tensor-core search
        L.call(entry, L.ptr(z2), C.c_int64(z2.stride(0)), L.ptr(embed), L.ptr(q_raw), L.ptr(q_st),
               L.ptr(diff), L.ptr(idx), n_rows, n_heads, dim, K)
        ctx.save_for_backward(z2, q_raw)
        ctx.hd = (n_heads, dim)
        ctx.zshape = z.shape
        lead = z.shape[:-1]
        idx = idx.reshape(*lead, n_heads)
        ctx.mark_non_differentiable(idx)
        return q_st.reshape(*lead, n_heads * dim), diff.reshape(*lead, dim), idx

    @staticmethod
    def backward(ctx, g_q, g_diff, _g_idx):
        z2, q_raw = ctx.saved_tensors
        n_heads, dim = ctx.hd
        n_rows = q_raw.shape[0]
        zc = z2.contiguous()
        gq = g_q.reshape(n_rows, -1).contiguous() if g_q is not None else None
        gd = g_diff.reshape(n_rows, -1).contiguous() if g_diff is not None else None
        gz = torch.empty_like(q_raw)
        L.call("msmc_vq_backward", L.ptr(gq), L.ptr(gd), L.ptr(zc), L.ptr(q_raw), L.ptr(gz), n_rows, n_heads, dim)
        return gz.reshape(ctx.zshape), None, None, None


def vq_quantize(z, embed, n_heads, dim):
    """z (..., n_heads*dim), embed (n_heads, dim, K) -> (quant_st, diff (..., dim), idx (..., n_heads) int64)"""
    return _VQFn.apply(z, embed, n_heads, dim)


def vq_ema_update(z, idx, lengths, embed, embed_avg, cluster_size, decay, eps):
    """in-place EMA codebook update on the stacked buffers (n_heads, dim, K) / (n_heads, K)"""
    n_heads, dim, K = embed.shape
    B, t = z.shape[0], z.shape[1]
    z2 = z.detach().reshape(B * t, n_heads * dim)
    if z2.stride(-1) != 1:
        z2 = z2.contiguous()
    idx2 = idx.reshape(B * t, n_heads).contiguous()
    lengths = lengths.to(device=z.device, dtype=torch.int32).contiguous()
    L.require_cuda(z2, embed, embed_avg, cluster_size)
    assert embed.is_contiguous() and embed_avg.is_contiguous() and cluster_size.is_contiguous()
    L.call("msmc_vq_ema_update", L.ptr(z2), C.c_int64(z2.stride(0)), L.ptr(idx2), L.ptr(lengths), B, t, n_heads,
           dim, K, C.c_float(decay), C.c_float(eps), L.ptr(cluster_size), L.ptr(embed_avg), L.ptr(embed))


class _TripleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, embed, target, n_heads, dim, margin, reduce_mean):
        p2 = pred.reshape(-1, n_heads * dim).contiguous()
        n_rows = p2.shape[0]
        K = embed.shape[-1]
        t2 = target.reshape(n_rows, n_heads).contiguous()
        loss = torch.empty((n_rows, n_heads), dtype=torch.float32, device=pred.device)
        gp = torch.empty_like(p2)
        L.require_cuda(p2, embed, t2)
        L.call("msmc_vq_triple_loss", L.ptr(p2), C.c_int64(p2.stride(0)), L.ptr(embed), L.ptr(t2), L.ptr(loss),
               L.ptr(gp), n_rows, n_heads, dim, K, C.c_float(margin), int(reduce_mean))
        ctx.save_for_backward(gp)
        ctx.pshape = pred.shape
        return loss.sum(dim=1).reshape(pred.shape[:-1])

    @staticmethod
    def backward(ctx, gl):
        (gp,) = ctx.saved_tensors
        g = gp * gl.reshape(-1, 1)
        return g.reshape(ctx.pshape), None, None, None, None, None, None


def vq_triple_loss(pred, embed, target, n_heads, dim, margin=1e-6, reduction="mean"):
    return _TripleFn.apply(pred, embed, target, n_heads, dim, margin, reduction == "mean")


# ------------------------------------------------------------------------------ multi-tensor L1 (feature matching)
_l1_plans = {}
_l1_graph_plans = []


def _dense_perm(t):
    """permutation under which `t` is contiguous (feature maps are permuted views of channels-last buffers), or None"""
    order = sorted(range(t.dim()), key=lambda d: (-t.stride(d), d))
    return order if t.permute(order).is_contiguous() else None


def _staging(rows_x_n, device):
    return (torch.empty(rows_x_n, dtype=torch.int64).pin_memory(),
            torch.empty(rows_x_n, dtype=torch.int64, device=device))


def _l1_plan(sizes, device, with_grad):
    """chunk map (device, immutable) + pointer-table staging.  A plan used inside a CUDA-graph capture gets PRIVATE
    staging buffers (taken from spares allocated by an earlier eager call): the captured H2D copy re-reads the pinned
    buffer on every replay, so later eager calls must not overwrite it."""
    capturing = torch.cuda.is_current_stream_capturing()
    key = (tuple(sizes), device.index, with_grad)
    plan = _l1_plans.get(key)
    rows = 3 if with_grad else 2
    if plan is None:
        if capturing:
            raise L.MsmcError("l1_multi: run one eager step with these shapes before capturing a CUDA graph")
        chunk = L.load().msmc_l1_chunk_elems()
        ct, ci = [], []
        for t, n in enumerate(sizes):
            for c in range((n + chunk - 1) // chunk):
                ct.append(t)
                ci.append(c)
        host, table = _staging(rows * len(sizes), device)
        plan = {"sizes": torch.tensor(sizes, dtype=torch.int64, device=device),
                "ct": torch.tensor(ct, dtype=torch.int32, device=device),
                "ci": torch.tensor(ci, dtype=torch.int32, device=device),
                "host": host, "table": table, "n_chunks": len(ct), "event": None,
                "spares": [_staging(rows * len(sizes), device) for _ in range(2)]}
        _l1_plans[key] = plan
    if capturing:
        if not plan["spares"]:
            raise L.MsmcError("l1_multi: no private staging buffer left for another CUDA-graph capture")
        host, table = plan["spares"].pop()
        plan = dict(plan, host=host, table=table, spares=None)
        _l1_graph_plans.append(plan)
    return plan


def _l1_fill_table(plan, rows):
    """pointer table -> device through the plan's pinned staging buffer (graph-capturable H2D copy)"""
    capturing = torch.cuda.is_current_stream_capturing()
    if not capturing and plan["event"] is not None:
        plan["event"].synchronize()          # the previous eager copy out of the staging buffer has run
    host = plan["host"]
    k = 0
    for row in rows:
        for t in row:
            host[k] = t.data_ptr()
            k += 1
    plan["table"].copy_(host, non_blocking=True)
    if not capturing:
        plan["event"] = torch.cuda.Event()
        plan["event"].record()


class _L1MultiFn(torch.autograd.Function):
    """sum_t mean|a_t - b_t| over a list of pairs in 2 launches (forward) + 1 (backward); gradients only w.r.t. a_t"""

    @staticmethod
    def forward(ctx, n, *tensors):
        a_list, b_list = tensors[:n], tensors[n:]
        da, db, perms = [], [], []
        for a, b in zip(a_list, b_list):
            assert a.shape == b.shape
            L.require_cuda(a, b)
            perm = _dense_perm(a)
            if perm is None:
                perm = list(range(a.dim()))
                a = a.contiguous()
            a_d = a.permute(perm)
            b_d = b.permute(perm)
            if not b_d.is_contiguous():
                b_d = b_d.contiguous()
            da.append(a_d)
            db.append(b_d)
            perms.append(perm)
        dev = da[0].device
        plan = _l1_plan([t.numel() for t in da], dev, False)
        _l1_fill_table(plan, (da, db))
        partial = torch.empty(plan["n_chunks"], dtype=torch.float32, device=dev)
        out = torch.empty((), dtype=torch.float32, device=dev)
        L.call("msmc_l1_multi_fwd", L.ptr(plan["table"]), n, L.ptr(plan["sizes"]), L.ptr(plan["ct"]),
               L.ptr(plan["ci"]), plan["n_chunks"], L.ptr(partial), L.ptr(out))
        ctx.n, ctx.perms = n, perms
        ctx.save_for_backward(*da, *db)
        return out

    @staticmethod
    def backward(ctx, gout):
        n = ctx.n
        saved = ctx.saved_tensors
        da, db = saved[:n], saved[n:]
        dev = da[0].device
        ga = [torch.empty(t.shape, dtype=torch.float32, device=dev) for t in da]
        plan = _l1_plan([t.numel() for t in da], dev, True)
        _l1_fill_table(plan, (da, db, ga))
        g = gout.reshape(1).contiguous()
        L.call("msmc_l1_multi_bwd", L.ptr(plan["table"]), n, L.ptr(plan["sizes"]), L.ptr(plan["ct"]),
               L.ptr(plan["ci"]), plan["n_chunks"], L.ptr(g))
        grads = []
        for t, perm in zip(ga, ctx.perms):
            inv = [0] * len(perm)
            for i, d in enumerate(perm):
                inv[d] = i
            grads.append(t.permute(inv))     # same shape AND strides as the forward's feature-map view
        return (None,) + tuple(grads) + (None,) * n


def l1_multi(a_list, b_list):
    """sum over pairs of F.l1_loss(a, b) (mean reduction); b_list is treated as constant"""
    a_list, b_list = list(a_list), list(b_list)
    assert len(a_list) == len(b_list) and a_list
    return _L1MultiFn.apply(len(a_list), *a_list, *[b.detach() for b in b_list])


# ------------------------------------------------------------------------------------------- dropout rng
class DeviceRng:
    """Device-resident 64-bit seed; `advance()` is a captured device op so graph replays draw fresh masks."""
    _state = {}

    @classmethod
    def get(cls, device):
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        st = cls._state.get(key)
        if st is None:
            st = cls._state[key] = {"seed": torch.full((1,), int(torch.initial_seed()) & 0x7FFFFFFFFFFFFFFF,
                                                       dtype=torch.int64, device=device), "salt": 0}
        return st

    @classmethod
    def next_salt(cls, device):
        st = cls.get(device)
        st["salt"] += 1
        return st["seed"], st["salt"]

    @classmethod
    def advance(cls, device):
        st = cls.get(device)
        st["seed"].add_(0x632BE59BD9B4E019)
        st["salt"] = 0


# --------------------------------------------------------------------------------------------- attention
class _AttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, lengths, n_head, d, inv_temp, drop_p):
        qkv = qkv.contiguous()
        B, t, _ = qkv.shape
        out = torch.empty((B, t, n_head * d), dtype=torch.float32, device=qkv.device)
        lse = torch.empty((n_head * B, t), dtype=torch.float32, device=qkv.device)
        seed, salt = (None, 0)
        if drop_p > 0:
            seed, salt = DeviceRng.next_salt(qkv.device)
        L.require_cuda(qkv, lengths)
        L.call("msmc_attention_fwd", L.ptr(qkv), L.ptr(lengths), L.ptr(out), L.ptr(lse), B, t, n_head, d,
               C.c_float(inv_temp), C.c_float(drop_p), L.ptr(seed), C.c_uint64(salt))
        ctx.save_for_backward(qkv, lengths, out, lse, seed)
        ctx.args = (n_head, d, inv_temp, drop_p, salt)
        return out

    @staticmethod
    def backward(ctx, gout):
        qkv, lengths, out, lse, seed = ctx.saved_tensors
        n_head, d, inv_temp, drop_p, salt = ctx.args
        B, t, _ = qkv.shape
        gout = gout.contiguous()
        gqkv = torch.empty_like(qkv)
        L.call("msmc_attention_bwd", L.ptr(qkv), L.ptr(lengths), L.ptr(out), L.ptr(lse), L.ptr(gout), L.ptr(gqkv),
               B, t, n_head, d, C.c_float(inv_temp), C.c_float(drop_p), L.ptr(seed), C.c_uint64(salt))
        return gqkv, None, None, None, None, None


def attention(qkv, lengths, n_head, d, temperature, drop_p=0.0):
    return _AttnFn.apply(qkv, lengths, n_head, d, 1.0 / float(temperature), float(drop_p))


# --------------------------------------------------------------------------------------------- layernorm
class _AddLNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, r, gamma, beta, lengths, eps, drop_p):
        a = a.contiguous()
        r = r.contiguous() if r is not None else None
        B, t, Cn = a.shape
        y = torch.empty_like(a)
        xhat = torch.empty_like(a)
        rstd = torch.empty(B * t, dtype=torch.float32, device=a.device)
        seed, salt = (None, 0)
        if drop_p > 0:
            seed, salt = DeviceRng.next_salt(a.device)
        L.require_cuda(a, r, gamma, beta)
        L.call("msmc_add_layernorm_fwd", L.ptr(a), L.ptr(r), L.ptr(gamma), L.ptr(beta), L.ptr(lengths), L.ptr(y),
               L.ptr(xhat), L.ptr(rstd), B, t, Cn, C.c_float(eps), C.c_float(drop_p), L.ptr(seed),
               C.c_uint64(salt))
        ctx.save_for_backward(xhat, rstd, gamma, lengths, seed)
        ctx.args = (drop_p, salt, r is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        xhat, rstd, gamma, lengths, seed = ctx.saved_tensors
        drop_p, salt, has_r = ctx.args
        B, t, Cn = xhat.shape
        gy = gy.contiguous()
        ga = torch.empty_like(xhat)
        gr = torch.empty_like(xhat) if has_r else None
        dgamma = torch.empty_like(gamma)
        dbeta = torch.empty_like(gamma)
        nbytes = L.load().msmc_add_layernorm_bwd_workspace(Cn)
        ws = torch.empty(nbytes // 4, dtype=torch.float32, device=xhat.device)
        L.call("msmc_add_layernorm_bwd", L.ptr(gy), L.ptr(xhat), L.ptr(rstd), L.ptr(gamma), L.ptr(lengths),
               L.ptr(ga), L.ptr(gr), L.ptr(dgamma), L.ptr(dbeta), L.ptr(ws), B, t, Cn, C.c_float(drop_p),
               L.ptr(seed), C.c_uint64(salt))
        return ga, gr, dgamma, dbeta, None, None, None


def add_layernorm(a, r, gamma, beta, lengths=None, eps=1e-5, drop_p=0.0):
    """mask * LayerNorm(dropout(a) + r) over the last dim; a, r : (B, t, C); lengths int32 (B,) or None"""
    return _AddLNFn.apply(a, r, gamma, beta, lengths, float(eps), float(drop_p))


# --------------------------------------------------------------------------------------------- pointwise
class _SpecMagFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec, n_freq, floor_, floor_add):
        spec = spec.contiguous()
        Fp = spec.shape[-1] // 2
        rows = spec.numel() // (2 * Fp)
        mag = torch.empty(spec.shape[:-1] + (n_freq,), dtype=torch.float32, device=spec.device)
        L.require_cuda(spec)
        L.call("msmc_spec_magnitude_fwd", L.ptr(spec), L.ptr(mag), C.c_int64(rows), n_freq, Fp, C.c_float(floor_),
               int(floor_add))
        ctx.save_for_backward(spec, mag)
        ctx.args = (rows, n_freq, Fp, floor_, floor_add)
        return mag

    @staticmethod
    def backward(ctx, gmag):
        spec, mag = ctx.saved_tensors
        rows, n_freq, Fp, floor_, floor_add = ctx.args
        gmag = gmag.contiguous()
        gspec = torch.empty_like(spec)
        L.call("msmc_spec_magnitude_bwd", L.ptr(gmag), L.ptr(spec), L.ptr(mag), L.ptr(gspec), C.c_int64(rows),
               n_freq, Fp, C.c_float(floor_), int(floor_add))
        return gspec, None, None, None


def spec_magnitude(spec, floor_, floor_add=False, n_freq=None):
    """spec (..., 2*Fp) = [re | pad | im | pad] -> (..., n_freq); n_freq defaults to Fp (no padding)"""
    if n_freq is None:
        n_freq = spec.shape[-1] // 2
    return _SpecMagFn.apply(spec, int(n_freq), float(floor_), bool(floor_add))


class _MelDoubleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mel, ref_db, min_db):
        mel = mel.contiguous()
        out = torch.empty(mel.shape + (2,), dtype=torch.float32, device=mel.device)
        L.require_cuda(mel)
        L.call("msmc_mel_double_fwd", L.ptr(mel), L.ptr(out), C.c_int64(mel.numel()), C.c_float(ref_db),
               C.c_float(min_db))
        ctx.save_for_backward(mel)
        ctx.args = (ref_db, min_db)
        return out

    @staticmethod
    def backward(ctx, gout):
        (mel,) = ctx.saved_tensors
        gout = gout.contiguous()
        gmel = torch.empty_like(mel)
        L.call("msmc_mel_double_bwd", L.ptr(gout), L.ptr(mel), L.ptr(gmel), C.c_int64(mel.numel()),
               C.c_float(ctx.args[0]), C.c_float(ctx.args[1]))
        return gmel, None, None


def mel_double(mel, ref_db=20.0, min_db=-100.0):
    """mel (...,) -> (..., 2): [lin, clamp((20 log10(lin) - ref - min)/-min, 0, 1)]"""
    return _MelDoubleFn.apply(mel, float(ref_db), float(min_db))


class _LogClampFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, clip):
        x = x.contiguous()
        y = torch.empty_like(x)
        L.require_cuda(x)
        L.call("msmc_log_clamp_fwd", L.ptr(x), L.ptr(y), C.c_int64(x.numel()), C.c_float(clip))
        ctx.save_for_backward(x)
        ctx.clip = clip
        return y

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        gy = gy.contiguous()
        gx = torch.empty_like(x)
        L.call("msmc_log_clamp_bwd", L.ptr(gy), L.ptr(x), L.ptr(gx), C.c_int64(x.numel()), C.c_float(ctx.clip))
        return gx, None


def log_clamp(x, clip=1e-5):
    return _LogClampFn.apply(x, float(clip))


class _GatedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        C2 = x.shape[-1]
        Cn = C2 // 2
        rows = x.numel() // C2
        y = torch.empty(x.shape[:-1] + (Cn,), dtype=torch.float32, device=x.device)
        L.require_cuda(x)
        L.call("msmc_gated_act_fwd", L.ptr(x), L.ptr(y), C.c_int64(rows), Cn)
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        gy = gy.contiguous()
        gx = torch.empty_like(x)
        C2 = x.shape[-1]
        L.call("msmc_gated_act_bwd", L.ptr(gy), L.ptr(x), L.ptr(gx), C.c_int64(x.numel() // C2), C2 // 2)
        return gx


def gated_act(x):
    """x (..., 2C) -> tanh(x[..., :C]) * sigmoid(x[..., C:])   (channels-last form of modules.py:172-179)"""
    return _GatedFn.apply(x)
