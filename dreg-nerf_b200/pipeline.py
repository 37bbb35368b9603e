"""Pairs of a batch in flight on several CUDA streams.

The reference registers one pair per call (conerf/register/nerf_regtr.py:144-147 index batch element 0 only;
train_nerf_regtr.py:171-239 and eval_nerf_regtr.py:278-281 loop over pairs), and so does ``NeRFRegTr.forward``.
One pair alone cannot fill a B200: of the ~460 kernels of a 128^3 forward only the five level-1 FPN convolutions
occupy all 148 SMs, the ~100 deep-level GEMMs and the BatchNorm / point kernels are a few CTAs each.  The pairs of a
batch are independent, so ``PairPipeline`` keeps ``streams`` of them in flight: one worker thread per slot, each
with its own CUDA stream and its own engine (``nerf_regtr.set_engine_slot``: an engine's workspaces belong to the
work queued on one stream), all engines bound to the SAME parameter tensors.  Kernels of different pairs then
overlap on the device; per-pair results are bit-identical to the sequential loop (every pair still runs the same
kernels in the same order on its own stream).

Semantics that differ from a sequential loop, on purpose:
  * training-mode BatchNorm running statistics are updated by slot 0's pairs only (engine slots > 0 are created
    with drb_engine_set_update_running(0)) - the replicas of a DistributedDataParallel job behave the same way;
  * gradients of pairs that finish out of order are accumulated by autograd in completion order (fp32 sums in a
    different order, not bit-reproducible run to run).

ctypes releases the GIL for the duration of every C-ABI call and torch does for its kernels, so the worker threads
overlap on the host as well.
"""
import queue
import threading

import torch

from . import nerf_regtr


def _record_stream(obj, stream):
    """Tensors produced on a worker stream and consumed on the caller's: tell the caching allocator."""
    if torch.is_tensor(obj):
        if obj.is_cuda:
            obj.record_stream(stream)
    elif isinstance(obj, dict):
        for v in obj.values():
            _record_stream(v, stream)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            _record_stream(v, stream)


class PairPipeline:
    """``PairPipeline(device, streams=4).map(fn, items)`` -> ``[fn(item) for item in items]`` with up to ``streams``
    calls in flight, each on its own CUDA stream / engine slot.  ``streams=1`` (or a CPU device) degenerates to the
    plain loop on the caller's thread and stream."""

    def __init__(self, device=None, streams=4):
        self.device = torch.device(device) if device is not None else None
        self.n = max(1, int(streams))
        self.cuda = self.device is not None and self.device.type == "cuda"
        self._threads = []
        self._jobs = [queue.Queue() for _ in range(self.n)]
        self._done = queue.Queue()
        self._streams = [None] * self.n
        self._closed = False

    # ---------------------------------------------------------------------------------------------
    def _worker(self, slot):
        nerf_regtr.set_engine_slot(slot)
        if self.cuda:
            torch.cuda.set_device(self.device)
            self._streams[slot] = torch.cuda.Stream(self.device)
        self._done.put(("ready", slot, None))
        while True:
            job = self._jobs[slot].get()
            if job is None:
                return
            fn, items, nxt, lock, results, start_event, grad = job
            err = None
            end_event = None
            try:
                with torch.set_grad_enabled(grad):
                    if self.cuda:
                        stream = self._streams[slot]
                        stream.wait_event(start_event)          # everything the caller queued before map()
                        with torch.cuda.stream(stream):
                            self._drain(fn, items, nxt, lock, results)
                            end_event = torch.cuda.Event()
                            end_event.record(stream)
                    else:
                        self._drain(fn, items, nxt, lock, results)
            except BaseException as exc:  # noqa: BLE001 - re-raised in the caller
                err = exc
                with lock:
                    nxt[0] = len(items)                         # the other workers stop after their current item
            self._done.put(("done", slot, (err, end_event)))

    @staticmethod
    def _drain(fn, items, nxt, lock, results):
        while True:
            with lock:
                i = nxt[0]
                if i >= len(items):
                    return
                nxt[0] = i + 1
            results[i] = fn(items[i])

    def _start(self):
        if self._threads:
            return
        for slot in range(self.n):
            t = threading.Thread(target=self._worker, args=(slot,), name="drb-pipeline-%d" % slot, daemon=True)
            t.start()
            self._threads.append(t)
        for _ in range(self.n):
            self._done.get()

    # ---------------------------------------------------------------------------------------------
    def map(self, fn, items):
        items = list(items)
        if self._closed:
            raise RuntimeError("PairPipeline is closed")
        if self.n == 1 or len(items) <= 1:
            return [fn(it) for it in items]
        self._start()
        results = [None] * len(items)
        nxt, lock = [0], threading.Lock()
        start_event = None
        caller = None
        if self.cuda:
            caller = torch.cuda.current_stream(self.device)
            start_event = torch.cuda.Event()
            start_event.record(caller)
        used = min(self.n, len(items))
        grad = torch.is_grad_enabled()
        for slot in range(used):
            self._jobs[slot].put((fn, items, nxt, lock, results, start_event, grad))
        first_err = None
        for _ in range(used):
            _, _, (err, end_event) = self._done.get()
            if err is not None and first_err is None:
                first_err = err
            if end_event is not None:
                caller.wait_event(end_event)                    # the caller's stream continues after every pair
        if first_err is not None:
            raise first_err
        if self.cuda:
            _record_stream(results, caller)
        return results

    def close(self):
        if self._closed:
            return
        self._closed = True
        for q in self._jobs[:len(self._threads)]:
            q.put(None)
        for t in self._threads:
            t.join(timeout=5)
        self._threads = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
