"""Host <-> device streaming around the block/net forward.

The reference uploads the whole operator tuple with blocking ``.to(device)`` calls before every forward
(``experiments/human_segmentation_original/human_segmentation_original.py:111-120``) and reads results back
synchronously.  ``StreamedForward`` keeps the same per-step traffic (nothing is cached across steps) but puts
the upload of step i+1, the kernels of step i and the download of step i-1 on three CUDA streams, so a sequence
of forwards runs at the speed of the slowest of the three instead of their sum.
"""
from __future__ import annotations

import torch


class _Ticket:
    __slots__ = ("out", "fin", "dev")

    def __init__(self, out, fin, dev):
        self.out, self.fin, self.dev = out, fin, dev

    def __getitem__(self, k):          # dict-style access kept for callers
        return getattr(self, k)


class StreamedForward:
    """``fn(**device_inputs) -> device tensor`` driven from pinned host buffers.

    ``submit(host_inputs)`` enqueues upload -> compute -> download and returns a ticket;
    ``result(ticket)`` blocks until that step's output has landed in its pinned host buffer.
    At most ``depth`` steps are in flight (their device copies are alive simultaneously).
    """

    def __init__(self, fn, device, depth=2):
        self.fn = fn
        self.device = torch.device(device)
        self.depth = depth
        self.s_in = torch.cuda.Stream(device=self.device)
        self.s_out = torch.cuda.Stream(device=self.device)
        self.inflight = []
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def submit(self, host_inputs, out_host=None):
        while len(self.inflight) >= self.depth:
            self._retire(self.inflight.pop(0))
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.s_in):
            dev = {}
            for k, t in host_inputs.items():
                d = t.to(self.device, non_blocking=True)
                d.record_stream(main)
                dev[k] = d
                self.h2d_bytes += t.numel() * t.element_size()
            up = torch.cuda.Event()
            up.record(self.s_in)
        main.wait_event(up)
        out = self.fn(**dev)
        done = torch.cuda.Event()
        done.record(main)
        if out_host is None:
            out_host = torch.empty(out.shape, dtype=out.dtype).pin_memory()
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(done)
            out.record_stream(self.s_out)
            out_host.copy_(out, non_blocking=True)
            fin = torch.cuda.Event()
            fin.record(self.s_out)
        self.d2h_bytes += out_host.numel() * out_host.element_size()
        ticket = _Ticket(out_host, fin, dev)
        self.inflight.append(ticket)
        return ticket

    @staticmethod
    def _retire(ticket):
        ticket.fin.synchronize()
        ticket.dev = None

    def result(self, ticket):
        self._retire(ticket)
        self.inflight = [t for t in self.inflight if t is not ticket]
        return ticket.out

    def drain(self):
        while self.inflight:
            self._retire(self.inflight.pop(0))
