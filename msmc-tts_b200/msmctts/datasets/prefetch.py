"""Host -> device batch prefetch.  The reference copies each batch synchronously from pageable
memory right before the step (utils.py:137-151, pin_memory=False); at B200 step times (tens of ms) that copy and
the loader's collate sit on the critical path.  Here batch i+1 is copied from PINNED memory on a side stream while
step i runs; the consumer only waits on an event."""
import torch


class DevicePrefetcher(object):
    def __init__(self, loader, device=None):
        self.loader = loader
        self.device = device or (torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available()
                                 else None)
        self.stream = torch.cuda.Stream(device=self.device) if self.device is not None else None

    def __len__(self):
        return len(self.loader)

    def _upload(self, batch):
        if self.stream is None:
            return batch, None
        with torch.cuda.stream(self.stream):
            out = {}
            for k, v in batch.items():
                if torch.is_tensor(v):
                    if not v.is_pinned():
                        v = v.pin_memory()
                    v = v.to(self.device, non_blocking=True)
                out[k] = v
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return out, ev

    def __iter__(self):
        it = iter(self.loader)
        try:
            nxt = self._upload(next(it))
        except StopIteration:
            return
        while nxt is not None:
            cur, ev = nxt
            try:
                nxt = self._upload(next(it))            # overlaps with the step that consumes `cur`
            except StopIteration:
                nxt = None
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)
                for v in cur.values():                  # the tensors were allocated on the side stream
                    if torch.is_tensor(v) and v.is_cuda:
                        v.record_stream(torch.cuda.current_stream())
            yield cur
