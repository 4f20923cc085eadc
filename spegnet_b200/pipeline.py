"""Host-to-host streaming of batches through the drop-in module with the copies overlapped on their own streams.

The reference's engines move one batch at a time: ``images.to(device)`` -> ``model(images)`` -> ``.cpu()`` on the
default stream, synchronising between (engine/evaluator.py:505-511,522-554; engine/predictor.py:332-365), so the PCIe
time of every batch (201 MB in, 68 MB out at batch 64) is added to its compute time.  ``HostPipeline`` keeps the same
per-batch contract -- pinned host images in, host logits out, in order -- but runs three streams: the host->device copy
of batch i+1 and the device->host copy of batch i-1 proceed on the two copy engines while the kernels of batch i run.

    pipe = HostPipeline(model)
    for out in pipe.run(host_batches):        # out = {'prediction': [B,1,S,S] fp32 pinned, 'edge': [B,1,S/8,S/8]}
        ...                                    # valid until the generator is advanced again (ring of `depth` buffers)

Only memory movement and stream plumbing live here; the forward itself is the module's (no arithmetic in PyTorch).
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, List, Optional

import torch


class HostPipeline:
    def __init__(self, model, device: Optional[torch.device] = None, depth: int = 2):
        if depth < 2:
            raise ValueError("depth must be >= 2 (one batch in flight per stage)")
        self.model = model
        self.device = device or next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("HostPipeline needs the model on a CUDA (B200) device; there is no CPU fallback")
        self.depth = depth
        self.s_in = torch.cuda.Stream(device=self.device)
        self.s_out = torch.cuda.Stream(device=self.device)
        self._dev_in: List[Optional[torch.Tensor]] = [None] * depth
        self._host_out: List[Optional[Dict[str, torch.Tensor]]] = [None] * depth
        self._fwd_done: List[Optional[torch.cuda.Event]] = [None] * depth
        self._out_done: List[Optional[torch.cuda.Event]] = [None] * depth

    def _stage_in(self, slot: int, host: torch.Tensor) -> torch.cuda.Event:
        buf = self._dev_in[slot]
        fresh = buf is None or buf.shape != host.shape
        compute = torch.cuda.current_stream(self.device)  # taken BEFORE switching to the copy stream below
        if fresh:
            buf = torch.empty(host.shape, dtype=torch.float32, device=self.device)
            self._dev_in[slot] = buf
        with torch.cuda.stream(self.s_in):
            if fresh:
                # The block comes from the caching allocator of the COMPUTE stream and may have just been released by
                # kernels that are still queued there (an earlier forward's workspace or outputs): the copy stream must
                # not write it before everything enqueued on the compute stream so far has run.
                self.s_in.wait_stream(compute)
            elif self._fwd_done[slot] is not None:  # the forward that last read this buffer must be done
                self.s_in.wait_event(self._fwd_done[slot])
            buf.copy_(host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.s_in)
        return ev

    def _forward(self, slot: int, ready: torch.cuda.Event) -> None:
        compute = torch.cuda.current_stream(self.device)
        compute.wait_event(ready)
        out = self.model(self._dev_in[slot])
        pred, edge = out["predictions"][-1], out["edge"]
        done = torch.cuda.Event()
        done.record(compute)
        self._fwd_done[slot] = done
        host = self._host_out[slot]
        if host is None or host["prediction"].shape != pred.shape:
            host = {"prediction": torch.empty(pred.shape, dtype=pred.dtype).pin_memory(),
                    "edge": torch.empty(edge.shape, dtype=edge.dtype).pin_memory()}
            self._host_out[slot] = host
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(done)
            host["prediction"].copy_(pred, non_blocking=True)
            host["edge"].copy_(edge, non_blocking=True)
            pred.record_stream(self.s_out)  # fresh allocations of the compute stream, read on the copy stream
            edge.record_stream(self.s_out)
            ev = torch.cuda.Event()
            ev.record(self.s_out)
        self._out_done[slot] = ev

    @torch.no_grad()
    def run(self, host_batches: Iterable[torch.Tensor]) -> Iterator[Dict[str, torch.Tensor]]:
        """Yields one dict of pinned host tensors per input batch, in order.  Input batches should be pinned
        ([B,3,S,S] fp32) for the copies to be asynchronous."""
        it = iter(host_batches)
        pending: List[int] = []  # slots whose results have not been yielded yet, oldest first
        i = 0
        nxt = next(it, None)
        staged = self._stage_in(0, nxt) if nxt is not None else None
        while nxt is not None:
            slot = i % self.depth
            ready = staged
            nxt = next(it, None)
            if nxt is not None:  # prefetch the next batch while this one computes
                staged = self._stage_in((i + 1) % self.depth, nxt)
            self._forward(slot, ready)  # overwrites the host buffers of batch i - depth (yielded `depth - 1` steps ago)
            pending.append(slot)
            i += 1
            if len(pending) == self.depth:
                old = pending.pop(0)
                self._out_done[old].synchronize()
                yield self._host_out[old]
        for old in pending:
            self._out_done[old].synchronize()
            yield self._host_out[old]
