"""Parameter containers whose names/shapes match torch.nn (and old-style torch.nn.utils.weight_norm) modules so the
reference's checkpoints load unchanged, with forwards that run on the channels-last sm_100a kernels."""
import math
import warnings

import torch
from torch import nn

from . import functional as Fn


def _conv_init(weight, bias, fan_in):
    nn.init.kaiming_uniform_(weight, a=math.sqrt(5))
    if bias is not None and fan_in > 0:
        bound = 1 / math.sqrt(fan_in)
        nn.init.uniform_(bias, -bound, bound)


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def _prepped(mod, v, g, transposed=False, swap_taps=False):
    """GEMM-layout (weight-normalised) weight of `mod`, re-used while the parameters are unchanged.

    The discriminator runs twice per phase (fake, real) on the same weights: the re-parametrisation, its
    tensor-core operand images and its backward then run once per parameter version instead of once per call.
    The entry is dropped as soon as a backward pass has consumed it (its saved tensors are gone by then).
    An entry produced by `prefetch_weights` carries the event recorded after its kernels on the prefetch stream;
    every consumer stream waits on it."""
    scope = Fn.PREP_SCOPE[0]
    if scope == 0:       # sharing is only safe inside a trainer step (nobody edits .data behind autograd's back there)
        if not torch.is_grad_enabled():
            # inference (no autograd graph): the re-parametrised GEMM-layout weight and the tensor-core operand images
            # hanging off it are baked ONCE per parameter version (in-place updates bump `_version`), not per call
            key = (v._version, -1 if g is None else g._version, v.data_ptr(), transposed, swap_taps)
            hit = mod.__dict__.get("_msmc_infer_prep")
            if hit is not None and hit[0] == key:
                return hit[1]
            w = Fn.prep_conv_weight(v.transpose(2, 3) if swap_taps else v, g, transposed=transposed)
            mod.__dict__["_msmc_infer_prep"] = (key, w)
            return w
        return Fn.prep_conv_weight(v.transpose(2, 3) if swap_taps else v, g, transposed=transposed)
    grad = torch.is_grad_enabled() and (v.requires_grad or (g is not None and g.requires_grad))
    key = (scope, v._version, -1 if g is None else g._version, grad, v.data_ptr())
    hit = mod.__dict__.get("_msmc_prep")
    if hit is not None and hit[0] == key and not hit[2]["spent"]:
        ev = hit[2].get("event")
        if ev is not None and not Fn.PREFETCHING[0]:
            torch.cuda.current_stream().wait_event(ev)
        return hit[1]
    token = {"spent": False, "event": None}
    w = Fn.prep_conv_weight(v.transpose(2, 3) if swap_taps else v, g, transposed=transposed, token=token)
    w._msmc_owner = mod          # lets _weight_image remember which operand images this layer asks for
    mod.__dict__["_msmc_prep"] = (key, w, token)
    return w


# prefetched weights are re-parametrised on the prefetch stream, so their backward nodes run there too while the
# AccumulateGrad nodes stay on the stream that created them; autograd orders the two with events (correct, and the
# step is captured on a non-default stream), but says so once per backward call
warnings.filterwarnings("ignore", message="The AccumulateGrad node's stream does not match")


import os

PREFETCH_GROUP_BYTES = int(os.environ.get("MSMC_PF_GROUP_MB", "8")) << 20


def prefetch_weights(root):
    """Re-parametrise every conv / linear weight under `root` (and rebuild the tensor-core operand images each
    layer used on its previous call) on a dedicated side stream, in module order, ahead of the forward pass.

    Issued inline these are ~700 launches of 3-10 us per step in front of every conv.  Here the layers are walked
    WITHOUT launching (each layer's autograd node and output buffers are created, the work is recorded as a job),
    and every ~8 MB of operands the recorded jobs run as two multi-tensor launches (msmc_weight_norm_fwd_multi,
    msmc_weight_image_multi) on the side stream; the layers of a group share one event and a consuming stream waits
    only for its own group.  Must be called inside a `prep_scope()`; returns the side stream (the caller joins it
    before the step ends) or None when there is nothing to do."""
    if Fn.PREP_SCOPE[0] == 0 or not Fn.PREFETCH_WEIGHTS:
        return None
    mods = [m for m in root.modules() if hasattr(m, "_prep_spec")]
    if not mods or not next(root.parameters()).is_cuda:
        return None
    cur = torch.cuda.current_stream()
    side = Fn.prefetch_stream(cur)
    side.wait_stream(cur)
    Fn.PREFETCHING[0] = True
    # the n-th prefetch of `root` inside one step keeps its own pointer-table staging (D is prefetched twice per step:
    # before its own update and again, frozen, for the generator step)
    seen = root.__dict__.get("_msmc_pf_calls")
    nth = seen[1] + 1 if (seen is not None and seen[0] == Fn.PREP_SCOPE[0]) else 0
    root.__dict__["_msmc_pf_calls"] = (Fn.PREP_SCOPE[0], nth)
    state = {"jobs": {"wn": [], "img": []}, "tokens": [], "bytes": 0, "group": 0}

    def flush():
        # two launches for the whole group (round 1: one per layer and per image, ~700 per train step); the group's
        # consumers wait on ONE event.  Groups are cut by operand bytes so that the first layers of the forward pass
        # do not wait for the re-parametrisation of the whole network.
        Fn.DEFERRED[0] = None
        if state["tokens"]:
            Fn.flush_deferred(state["jobs"], (id(root), nth, state["group"]))
            ev = torch.cuda.Event()
            ev.record(side)
            for token in state["tokens"]:
                token["event"] = ev
        state.update(jobs={"wn": [], "img": []}, tokens=[], bytes=0, group=state["group"] + 1)
        Fn.DEFERRED[0] = state["jobs"]

    Fn.DEFERRED[0] = state["jobs"]
    try:
        with torch.cuda.stream(side):
            for m in mods:
                spec = m._prep_spec()
                if spec is None:
                    continue
                w = _prepped(m, *spec)             # allocates + records the job (autograd node included), no launch
                token = m.__dict__["_msmc_prep"][2]
                if token.get("event") is not None or token.get("inline"):
                    continue                       # already prepared in this scope
                keys = m.__dict__.get("_msmc_img_keys", ())
                for key in keys:
                    Fn._weight_image(w, *key)
                state["tokens"].append(token)
                state["bytes"] += w.numel() * 4 * (1 + 2 * len(keys))
                if state["bytes"] >= PREFETCH_GROUP_BYTES:
                    flush()
            flush()
    finally:
        Fn.DEFERRED[0] = None
        Fn.PREFETCHING[0] = False
    return side


class Linear(nn.Linear):
    """nn.Linear parameters; forward = msmc_conv_forward with the weight consumed in its native (Co, Ci) layout."""

    def _prep_spec(self):
        # only when the previous call took the tensor-core path (it needs the GEMM-layout copy of the weight)
        if not self.__dict__.get("_msmc_lin_umma") or not self.weight.requires_grad:
            return None
        return (self.weight.reshape(self.out_features, self.in_features, 1), None)

    def forward(self, x, post="none", residual=None):
        return Fn.linear_cl(x, self.weight, self.bias, residual=residual, post=post, owner=self)


class Conv1d(nn.Module):
    """nn.Conv1d parameters `weight` (Co, Ci, K), `bias`; x and y are (B, L, C)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, bias=True):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, kernel_size))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        _conv_init(self.weight, self.bias, in_channels * kernel_size)

    def gemm_weight(self):
        return self.weight, None

    def _prep_spec(self):
        v, g = self.gemm_weight()
        if self.kernel_size == 1 and g is None:
            if not self.__dict__.get("_msmc_lin_umma") or not v.requires_grad:
                return None
            return (v.reshape(v.shape[0], v.shape[1], 1), None)
        return (v, g)

    def forward(self, x, pre_slope=None, post="none", residual=None):
        B, L, _ = x.shape
        v, g = self.gemm_weight()
        if self.kernel_size == 1 and g is None:
            return Fn.linear_cl(x, v, self.bias, residual=residual, post=post, pre_slope=pre_slope, owner=self)
        w = _prepped(self, v, g)
        r4 = residual.unsqueeze(1) if residual is not None else None
        y = Fn.conv_cl(x.unsqueeze(1), w, self.bias, r4, kernel=(1, self.kernel_size), stride=(1, self.stride),
                       dilation=(1, self.dilation), padding=(0, self.padding), pre_slope=pre_slope, post=post)
        return y.squeeze(1)


def _fold_weight_norm(mod):
    """torch.nn.utils.remove_weight_norm semantics (reference hifigan/generator.py:57-64, common.py:43-51): replace
    the (weight_g, weight_v) pair by the plain parameter `weight` = g * v / ||v|| (norm over all dims but 0)."""
    if not hasattr(mod, "weight_v"):
        raise ValueError("weight_norm of '%s' not found" % mod.__class__.__name__)
    with torch.no_grad():
        v, g = mod.weight_v, mod.weight_g
        w = v * (g / v.norm(2, dim=tuple(range(1, v.dim())), keepdim=True))
    del mod.weight_g
    del mod.weight_v
    mod.weight = nn.Parameter(w)
    mod.__dict__.pop("_msmc_prep", None)
    mod.__dict__.pop("_msmc_infer_prep", None)


class WNConv1d(Conv1d):
    """weight_norm(nn.Conv1d): parameters `bias`, `weight_g` (Co,1,1), `weight_v` (Co,Ci,K) in that order."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        v = self.weight.data
        del self.weight
        self.weight_g = nn.Parameter(v.norm(2, dim=(1, 2), keepdim=True))
        self.weight_v = nn.Parameter(v)

    def gemm_weight(self):
        if hasattr(self, "weight_v"):
            return self.weight_v, self.weight_g
        return self.weight, None                    # after remove_weight_norm()

    def remove_weight_norm(self):
        _fold_weight_norm(self)


class WNConvTranspose1d(nn.Module):
    """weight_norm(nn.ConvTranspose1d): `weight_v` (Cin, Cout, K), `weight_g` (Cin,1,1) (torch's dim=0)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding=0):
        super().__init__()
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding
        v = torch.empty(in_channels, out_channels, kernel_size)
        self.bias = nn.Parameter(torch.empty(out_channels))
        _conv_init(v, self.bias, out_channels * kernel_size)
        self.weight_g = nn.Parameter(v.norm(2, dim=(1, 2), keepdim=True))
        self.weight_v = nn.Parameter(v)

    def _vg(self):
        return (self.weight_v, self.weight_g) if hasattr(self, "weight_v") else (self.weight, None)

    def remove_weight_norm(self):
        _fold_weight_norm(self)

    def _prep_spec(self):
        return self._vg() + (True,)

    def forward(self, x, pre_slope=None):
        w = _prepped(self, *self._vg(), transposed=True)
        y = Fn.conv_cl(x.unsqueeze(1), w, self.bias, kernel=(1, self.kernel_size), stride=(1, self.stride),
                       padding=(0, self.padding), transposed=True, pre_slope=pre_slope)
        return y.squeeze(1)


class WNConv2d(nn.Module):
    """weight_norm(nn.Conv2d): `bias`, `weight_g` (Co,1,1,1), `weight_v` (Co,Ci,KH,KW); x is (B, H, W, C).
    swap_hw=True runs on a tensor whose two spatial axes are exchanged w.r.t. the reference's (taps transposed)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, reflect=False, swap_hw=False):
        super().__init__()
        self.kernel_size, self.stride, self.padding = _pair(kernel_size), _pair(stride), _pair(padding)
        self.reflect, self.swap_hw = reflect, swap_hw
        self.in_channels, self.out_channels = in_channels, out_channels
        v = torch.empty(out_channels, in_channels, *self.kernel_size)
        self.bias = nn.Parameter(torch.empty(out_channels))
        _conv_init(v, self.bias, in_channels * self.kernel_size[0] * self.kernel_size[1])
        self.weight_g = nn.Parameter(v.norm(2, dim=(1, 2, 3), keepdim=True))
        self.weight_v = nn.Parameter(v)

    def _vg(self):
        return (self.weight_v, self.weight_g) if hasattr(self, "weight_v") else (self.weight, None)

    def remove_weight_norm(self):
        _fold_weight_norm(self)

    def _prep_spec(self):
        return self._vg() + (False, self.swap_hw)

    def forward(self, x, pre_slope=None, post="none"):
        KH, KW = self.kernel_size
        if not self.swap_hw:
            w = _prepped(self, *self._vg())   # [kh][kw][ci][co]
            return Fn.conv_cl(x, w, self.bias, kernel=(KH, KW), stride=self.stride, padding=self.padding,
                              reflect=self.reflect, pre_slope=pre_slope, post=post)
        # exchanged spatial axes: transpose the taps, then the standard contiguous GEMM layout [kw][kh][ci][co]
        # (the weight norm runs over (ci, kh, kw), so it is unaffected by the permutation)
        w = _prepped(self, *self._vg(), swap_taps=True)
        return Fn.conv_cl(x, w, self.bias, kernel=(KW, KH), stride=self.stride[::-1], padding=self.padding[::-1],
                          reflect=self.reflect, pre_slope=pre_slope, post=post)


class LayerNormParams(nn.LayerNorm):
    """nn.LayerNorm parameters; the math runs fused in add_layernorm (residual + dropout + LN + pad mask)."""

    def forward(self, a, r=None, lengths=None, drop_p=0.0):
        return Fn.add_layernorm(a, r, self.weight, self.bias, lengths, self.eps, drop_p)
