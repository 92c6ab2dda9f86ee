"""Adam / AdamW whose whole step -- global gradient-norm clip included -- is ONE call of msmc_adam_multi (two
launches; SURVEY 8f rank 1, reference trainers/msmctts_trainer.py:203-207 + optimizers/__init__.py:53-78).

State layout and hyper-parameter names follow torch.optim.Adam(W) (`step`, `exp_avg`, `exp_avg_sq` per parameter,
param_groups with lr/betas/eps/weight_decay), so reference optimizer checkpoints (reference
trainers/optimizers/__init__.py:47-57 stores torch's state_dict) load unchanged.  `lr` is a device scalar: the
scheduler can change it between CUDA-graph replays.  Update rule == torch's single-tensor path
(torch/optim/adam.py): decoupled decay for AdamW, L2 term added to the gradient for Adam.

Step counters are PER PARAMETER, exactly like torch: a parameter's `step` advances only on steps where it has a
gradient (the HiFiGAN decoder gets none during the reference's 50k warm-up steps, so its first real update must use
the bias correction of step 1).  Every `state[p]['step']` is a 0-dim view into one device vector per group; the
kernel advances the slots of the parameters it updates."""
import ctypes as C

import torch

from msmctts._b200 import lib as L


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False):
        params = list(params)
        if not params:
            raise ValueError("optimizer got an empty parameter list")
        dev = params[0].device
        if not torch.is_tensor(lr):
            lr = torch.tensor(float(lr), dtype=torch.float32, device=dev)
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, decoupled=decoupled,
                        amsgrad=False, maximize=False, capturable=True, foreach=None, fused=None,
                        differentiable=False)
        super().__init__(params, defaults)
        self._plans = {}
        self._graph_plans = []
        self.last_grad_norm = None       # device scalar written by the last clipped step

    # ---------------------------------------------------------------------------------------------- state
    def _ensure_state(self, group):
        """torch-compatible per-parameter state; `step` entries are views into the group's device vector"""
        steps = group.get("_steps")
        if steps is None:
            dev = group["params"][0].device
            steps = torch.zeros(len(group["params"]), dtype=torch.float32, device=dev)
            for i, p in enumerate(group["params"]):
                st = self.state.get(p)
                if st and "step" in st:          # loaded from a (torch or fused) checkpoint
                    steps[i] = float(st["step"])
            group["_steps"] = steps
            for i, p in enumerate(group["params"]):
                st = self.state.get(p)
                if st and "step" in st:
                    st["step"] = steps[i]
        for i, p in enumerate(group["params"]):
            if p.grad is None and p not in self.state:
                continue                         # like torch: state is created on the first step with a gradient
            st = self.state[p]
            if "exp_avg" not in st:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            if "step" not in st:
                st["step"] = steps[i]
        return steps

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        for group in self.param_groups:          # re-tie the per-parameter counters to one device vector
            group.pop("_steps", None)
            group.setdefault("decoupled", self.defaults["decoupled"])   # torch's groups do not carry it
            if not torch.is_tensor(group["lr"]):
                group["lr"] = torch.tensor(float(group["lr"]), dtype=torch.float32,
                                           device=group["params"][0].device)
        self._plans = {}

    def state_dict(self):
        sd = super().state_dict()
        for g in sd["param_groups"]:
            g.pop("_steps", None)
        for st in sd["state"].values():          # detach the views: a checkpoint holds plain 0-dim tensors
            if "step" in st and torch.is_tensor(st["step"]):
                st["step"] = st["step"].detach().clone()
        return sd

    # ----------------------------------------------------------------------------------------------- step
    def _plan(self, gi, active):
        """static per (group, set of parameters with gradients): sizes and chunk map on the device.  A plan used
        inside a CUDA-graph capture gets PRIVATE pointer-table staging (spares allocated by an earlier eager step):
        the captured H2D copy re-reads the pinned buffer on every replay, so eager steps must not overwrite it."""
        capturing = torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
        key = (gi, tuple(active))
        plan = self._plans.get(key)
        if plan is None:
            if capturing:
                raise L.MsmcError("FusedAdam: run one eager step before capturing a CUDA graph")
            group = self.param_groups[gi]
            ps = [group["params"][i] for i in active]
            dev = ps[0].device
            chunk = L.load().msmc_adam_chunk_elems()
            sizes = [p.numel() for p in ps]
            ct, ci = [], []
            for t, n in enumerate(sizes):
                for c in range((n + chunk - 1) // chunk):
                    ct.append(t)
                    ci.append(c)

            def staging():
                return (torch.empty(4 * len(ps), dtype=torch.int64).pin_memory(),
                        torch.empty(4 * len(ps), dtype=torch.int64, device=dev))
            host, table = staging()
            plan = {
                "sizes": torch.tensor(sizes, dtype=torch.int64, device=dev),
                "ct": torch.tensor(ct, dtype=torch.int32, device=dev),
                "ci": torch.tensor(ci, dtype=torch.int32, device=dev),
                "sidx": torch.tensor(list(active), dtype=torch.int32, device=dev),
                "partial": torch.empty(len(ct), dtype=torch.float32, device=dev),
                "norm": torch.zeros((), dtype=torch.float32, device=dev),
                "host": host, "table": table, "n_chunks": len(ct), "event": None, "ptrs": None,
                "spares": [staging() for _ in range(2)],
            }
            self._plans[key] = plan
        if capturing:
            if not plan["spares"]:
                raise L.MsmcError("FusedAdam: no private staging buffer left for another CUDA-graph capture")
            host, table = plan["spares"].pop()
            plan = dict(plan, host=host, table=table, spares=None, event=None, ptrs=None,
                        partial=torch.empty_like(plan["partial"]), norm=torch.zeros_like(plan["norm"]))
            self._graph_plans.append(plan)
        return plan

    @torch.no_grad()
    def step(self, closure=None, max_grad_norm=None):
        """max_grad_norm: clip the global l2 norm of this optimizer's gradients to it inside the same launches
        (torch.nn.utils.clip_grad_norm_ semantics; the norm lands in `self.last_grad_norm`, a device scalar)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if max_grad_norm is not None and len(self.param_groups) != 1:
            raise L.MsmcError("FusedAdam: the fused gradient clip spans one parameter group")
        for gi, group in enumerate(self.param_groups):
            active = [i for i, p in enumerate(group["params"]) if p.grad is not None]
            if not active:
                continue
            steps = self._ensure_state(group)
            plan = self._plan(gi, active)
            capturing = torch.cuda.is_current_stream_capturing()
            n = len(active)
            ptrs = []
            for i in active:
                p = group["params"][i]
                g = p.grad
                if not (p.is_contiguous() and g.is_contiguous() and p.dtype == torch.float32 and
                        g.dtype == torch.float32):
                    raise L.MsmcError("FusedAdam needs contiguous fp32 parameters and gradients")
                st = self.state[p]
                ptrs.append((p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()))
            if capturing or plan["ptrs"] != ptrs:
                # the pinned staging buffer may still be the source of the previous step's queued H2D copy (the
                # step is sync-free, the host can run ahead): wait for that copy before rewriting it
                if not capturing and plan["event"] is not None:
                    plan["event"].synchronize()
                host = plan["host"]
                for k, quad in enumerate(ptrs):
                    host[k], host[n + k], host[2 * n + k], host[3 * n + k] = quad
                plan["table"].copy_(host, non_blocking=True)      # graph-capturable H2D from pinned memory
                if not capturing:
                    plan["event"] = torch.cuda.Event()
                    plan["event"].record()
                    plan["ptrs"] = ptrs
            b1, b2 = group["betas"]
            clip = float(max_grad_norm) if max_grad_norm is not None else 0.0
            L.call("msmc_adam_multi", L.ptr(plan["table"]), n, L.ptr(plan["sizes"]), L.ptr(plan["ct"]),
                   L.ptr(plan["ci"]), plan["n_chunks"], L.ptr(plan["sidx"]), L.ptr(steps),
                   L.ptr(plan["partial"]) if clip > 0 else None, C.c_float(clip),
                   L.ptr(plan["norm"]) if clip > 0 else None, L.ptr(group["lr"]), C.c_float(b1), C.c_float(b2),
                   C.c_float(group["eps"]), C.c_float(group["weight_decay"]), 1 if group["decoupled"] else 0)
            if clip > 0:
                self.last_grad_norm = plan["norm"]
        return loss
