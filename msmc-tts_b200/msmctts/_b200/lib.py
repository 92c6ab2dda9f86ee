"""ctypes binding of libmsmc_b200.so (the C-ABI declared in include/msmc_b200.h).

There is no CPU fallback: if the library is missing, or a call returns a non-zero status, this raises.
PyTorch is used only for device memory and streams; every argument crossing the boundary is a raw
device pointer, a size, or the current CUDA stream handle.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmsmc_b200.so")

XF_NONE, XF_LRELU, XF_RELU, XF_TANH, XF_MUL_DLRELU, XF_MUL_DRELU, XF_MUL_DTANH = range(7)


class ConvGeom(C.Structure):
    """mirror of msmc_conv_geom"""
    _fields_ = [
        ("B", C.c_int32), ("Hs", C.c_int32), ("Ws", C.c_int32), ("Cs", C.c_int32),
        ("Hd", C.c_int32), ("Wd", C.c_int32), ("Cd", C.c_int32),
        ("KH", C.c_int32), ("KW", C.c_int32),
        ("sh", C.c_int32), ("sw", C.c_int32), ("dh", C.c_int32), ("dw", C.c_int32),
        ("ph", C.c_int32), ("pw", C.c_int32),
        ("pad_reflect", C.c_int32), ("transposed", C.c_int32),
        ("ld_src", C.c_int64), ("ld_dst", C.c_int64), ("ld_res", C.c_int64),
        ("ld_saux", C.c_int64), ("ld_daux", C.c_int64),
        ("ws_kh", C.c_int64), ("ws_kw", C.c_int64), ("ws_cs", C.c_int64), ("ws_cd", C.c_int64),
        ("src_xf", C.c_int32), ("src_slope", C.c_float),
        ("dst_xf", C.c_int32), ("dst_slope", C.c_float),
    ]


_P = C.c_void_p
_I32, _I64, _F, _U64 = C.c_int32, C.c_int64, C.c_float, C.c_uint64
_G = C.POINTER(ConvGeom)

# name -> (restype, argtypes); the exported-symbol test walks this table against include/msmc_b200.h
SIGNATURES = {
    "msmc_conv_forward": (C.c_int, [_G, _P, _P, _P, _P, _P, _P, _P, _P]),
    "msmc_conv_wgrad_workspace": (C.c_int64, [_G]),
    "msmc_conv_wgrad": (C.c_int, [_G, _P, _P, _P, _P, _P, _P, _P, _I64, _P]),
    "msmc_umma_tile_n": (C.c_int, [_I32, _I64]),
    "msmc_weight_image_elems": (C.c_int64, [_I32, _I32, _I32, _I32, _I32, _I32]),
    "msmc_weight_image": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    "msmc_conv_forward_umma": (C.c_int, [_G, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _P]),
    "msmc_conv_reuse_eligible": (C.c_int, [_G]),
    "msmc_conv_forward_umma_reuse": (C.c_int, [_G, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _P]),
    "msmc_conv_wgrad_umma_workspace": (C.c_int64, [_G]),
    "msmc_conv_wgrad_umma": (C.c_int, [_G, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _P]),
    "msmc_weight_norm_fwd": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I64, _I64, _I64, _P]),
    "msmc_weight_norm_fwd_multi": (C.c_int, [_P, _I32, _I64, _P]),
    "msmc_weight_image_multi": (C.c_int, [_P, _I32, _I64, _P]),
    "msmc_weight_norm_bwd": (C.c_int, [_P, _I64, _I64, _I64, _P, _P, _P, _P, _P, _I32, _I32, _I32, _P]),
    "msmc_reflect_pad_fold": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    "msmc_vq_search": (C.c_int, [_P, _I64, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P]),
    "msmc_vq_search_umma": (C.c_int, [_P, _I64, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P]),
    "msmc_vq_ema_update": (C.c_int, [_P, _I64, _P, _P, _I32, _I32, _I32, _I32, _I32, _F, _F, _P, _P, _P, _P]),
    "msmc_vq_backward": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _I32, _P]),
    "msmc_vq_triple_loss": (C.c_int, [_P, _I64, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _F, _I32, _P]),
    "msmc_attention_fwd": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _F, _F, _P, _U64, _P]),
    "msmc_attention_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _F, _F, _P, _U64, _P]),
    "msmc_add_layernorm_fwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _F, _F, _P, _U64, _P]),
    "msmc_add_layernorm_bwd_workspace": (C.c_int64, [_I32]),
    "msmc_add_layernorm_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _F, _P, _U64, _P]),
    "msmc_spec_magnitude_fwd": (C.c_int, [_P, _P, _I64, _I32, _I32, _F, _I32, _P]),
    "msmc_spec_magnitude_bwd": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _I32, _F, _I32, _P]),
    "msmc_mel_double_fwd": (C.c_int, [_P, _P, _I64, _F, _F, _P]),
    "msmc_mel_double_bwd": (C.c_int, [_P, _P, _P, _I64, _F, _F, _P]),
    "msmc_xform_apply": (C.c_int, [_P, _P, _P, _I64, _I32, _F, _P]),
    "msmc_log_clamp_fwd": (C.c_int, [_P, _P, _I64, _F, _P]),
    "msmc_log_clamp_bwd": (C.c_int, [_P, _P, _P, _I64, _F, _P]),
    "msmc_adam_chunk_elems": (C.c_int, []),
    "msmc_adam_multi": (C.c_int, [_P, _I32, _P, _P, _P, _I32, _P, _P, _P, _F, _P, _P, _F, _F, _F, _F, _I32, _P]),
    "msmc_l1_chunk_elems": (C.c_int, []),
    "msmc_l1_multi_fwd": (C.c_int, [_P, _I32, _P, _P, _P, _I32, _P, _P, _P]),
    "msmc_l1_multi_bwd": (C.c_int, [_P, _I32, _P, _P, _P, _I32, _P, _P]),
    "msmc_frame_unfold": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    "msmc_overlap_add_fold": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    "msmc_gated_act_fwd": (C.c_int, [_P, _P, _I64, _I32, _P]),
    "msmc_gated_act_bwd": (C.c_int, [_P, _P, _P, _I64, _I32, _P]),
    "msmc_version": (C.c_int, []),
    "msmc_num_sms": (C.c_int, []),
}

_STATUS = {1: "MSMC_ERR_BAD_ARG", 2: "MSMC_ERR_LAUNCH", 3: "MSMC_ERR_UNSUPPORTED"}
_lib = None
# kernels each entry point launches (bench.py's `gpu_launches` is the sum over the timed region)
KERNELS_PER_CALL = {"msmc_l1_multi_fwd": 2, "msmc_adam_multi": 2, "msmc_conv_wgrad": 2, "msmc_conv_wgrad_umma": 2, "msmc_vq_ema_update": 2, "msmc_attention_bwd": 2,
                    "msmc_add_layernorm_bwd": 2}
launch_count = 0   # kernels launched through the C-ABI so far
_profile = None    # when a list: (name, meta, start_event, end_event) per call, for bench.py's roofline pass


class MsmcError(RuntimeError):
    pass


def load():
    """dlopen the in-tree library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MsmcError(
            "libmsmc_b200.so not found at %s -- build it with `python msmc-tts_b200/build.py` "
            "(there is no CPU / PyTorch fallback for the hot path)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def profile_begin():
    global _profile
    _profile = []


def profile_end():
    global _profile
    out, _profile = _profile, None
    return out


def call(name, *args, meta=None):
    """Invoke a status-returning entry point on the current stream."""
    global launch_count
    lib = load()
    if _profile is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args, stream_ptr())
        e1.record()
        _profile.append((name, meta, e0, e1))
    else:
        rc = getattr(lib, name)(*args, stream_ptr())
    launch_count += KERNELS_PER_CALL.get(name, 1)
    if rc != 0:
        raise MsmcError("%s failed: %s" % (name, _STATUS.get(rc, rc)))


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise MsmcError("msmctts hot-path ops run on CUDA (sm_100a) only; got a %s tensor -- "
                            "there is no CPU fallback" % t.device)
        if t is not None and t.dtype not in (torch.float32, torch.int64, torch.int32):
            raise MsmcError("unsupported dtype %s" % t.dtype)
