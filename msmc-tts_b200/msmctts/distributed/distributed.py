"""Data-parallel plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch (reference
distributed/distributed.py:21-33 init_distributed, :154-204 apply_gradient_allreduce).

Differences from the reference, all host-side:
  * initial state is broadcast as ONE flat buffer per dtype instead of one collective per tensor (:160-163);
  * gradients are all-reduced per trainable sub-module in ~25 MB buckets that are launched from
    post-accumulate-grad hooks as soon as a bucket's last gradient is produced, on a side stream, so the
    collective overlaps the rest of backward (the reference runs one blocking all_reduce after backward, :165-189);
  * only the sub-module being stepped is reduced (the reference re-sends the discriminator's gradients with the
    generator's, :171-176);
  * the 1/world_size scale is applied by NCCL's AVG reduction (pre-divide on gloo).
Codebook EMA buffers are NOT synchronised across ranks -- exactly like the reference (each rank's codebooks follow
its own batches; rank 0's are checkpointed).  See DESIGN.md "multi-GPU" for the alternative.
"""
import os

import torch
import torch.distributed as dist

BUCKET_BYTES = 25 << 20


def init_distributed(rank, num_gpus, group_name, dist_backend="nccl", dist_url="tcp://127.0.0.1:54321"):
    if dist_backend == "nccl":
        assert torch.cuda.is_available(), "Distributed mode requires CUDA."
        torch.cuda.set_device(rank % torch.cuda.device_count())
    if dist.is_initialized():
        return
    if "MASTER_ADDR" in os.environ and "RANK" in os.environ:
        dist.init_process_group(dist_backend)        # torchrun / bench.py launch
    else:
        dist.init_process_group(dist_backend, init_method=dist_url.replace("localhost", "127.0.0.1"),
                                world_size=num_gpus, rank=rank)


def broadcast_state(module, src=0):
    """rank-`src` parameters and buffers to everyone, one flat buffer per dtype"""
    by_dtype = {}
    for t in module.state_dict().values():
        if torch.is_tensor(t):
            by_dtype.setdefault(t.dtype, []).append(t)
    for tensors in by_dtype.values():
        flat = torch.cat([t.reshape(-1) for t in tensors])
        dist.broadcast(flat, src)
        off = 0
        for t in tensors:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n


class GradientReducer(object):
    """Bucketed, overlapped gradient all-reduce for ONE sub-module (e.g. task.autoencoder)."""

    def __init__(self, module, bucket_bytes=BUCKET_BYTES):
        self.world = dist.get_world_size()
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.use_avg = dist.get_backend() == "nccl"
        # buckets in reverse registration order ~ the order backward produces gradients
        self.buckets, cur, size = [], [], 0
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * p.element_size()
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self.bucket_of = {id(p): i for i, b in enumerate(self.buckets) for p in b}
        self.pending = [0] * len(self.buckets)
        self.handles = []
        self._handle_params = []
        self.enabled = False
        self.stream = torch.cuda.Stream() if torch.cuda.is_available() and self.params and self.params[0].is_cuda \
            else None
        for p in self.params:
            p.register_post_accumulate_grad_hook(self._hook)

    def arm(self):
        """call before the backward whose gradients should be reduced"""
        self.enabled = True
        self.pending = [len(b) for b in self.buckets]
        self.handles = []
        self._handle_params = []

    def _hook(self, p):
        if not self.enabled:
            return
        i = self.bucket_of[id(p)]
        self.pending[i] -= 1
        if self.pending[i] == 0:
            self._launch(i)

    def _launch(self, i):
        params = [p for p in self.buckets[i] if p.grad is not None]
        grads = [p.grad for p in params]
        if not grads:
            return
        self._handle_params.append(params)
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                flat = torch.cat([g.reshape(-1) for g in grads])
                self._reduce(flat)
            self.handles.append((flat, grads))
        else:
            flat = torch.cat([g.reshape(-1) for g in grads])
            self._reduce(flat)
            self.handles.append((flat, grads))

    def _reduce(self, flat):
        if self.use_avg:
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        else:
            flat.div_(self.world)
            dist.all_reduce(flat)

    def finish(self):
        """after backward: flush buckets whose params got no gradient, wait, and RE-POINT every `.grad` at its slice
        of the reduced flat buffer.  No copy back: round 1 scattered the reduced values with one `copy_` per gradient
        tensor (~640 tiny kernels serial on the main stream = the fixed +1.9 ms per step at N >= 2); the optimizer
        (fused Adam + clip) reads gradients through pointers, so a view of the bucket is as good as the original."""
        for i, n in enumerate(self.pending):
            if n > 0:
                self._launch(i)
                self.pending[i] = 0
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        for (flat, grads), params in zip(self.handles, self._handle_params):
            off = 0
            for p, g in zip(params, grads):
                n = g.numel()
                p.grad = flat[off:off + n].view_as(g)
                off += n
        self.handles = []
        self._handle_params = []
        self.enabled = False


def apply_gradient_allreduce(module):
    """Reference-compatible entry point: broadcast the initial state and attach one GradientReducer per child
    (task.autoencoder, task.discriminator, ...).  Trainers call `module.grad_reducers[name].arm()/finish()`."""
    broadcast_state(module, 0)
    module.grad_reducers = {name: GradientReducer(child) for name, child in module.named_children()
                            if any(p.requires_grad for p in child.parameters())}
    return module
