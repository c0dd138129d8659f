"""Serving side of the reference's second entry point (reference demo.py:39-73), without the Flask UI.

The reference's demo handles one request per Flask thread (``app.run(threaded=True)``) and every request runs its own
batch-1 forward. Here concurrent requests are BATCHED: ``RequestBatcher`` collects the requests that arrive within a short
window, groups them by input size and runs each group as one forward (``Engine.inference_u8``: the codecs of
demo.py:52-53,64-66 run on the device); ``DemoProcessor.process_image`` is the reference's ``process_image`` around it
(floor the size to a multiple of 8, PIL resize in, forward, PIL resize back).

    proc = DemoProcessor(models.create_model(opt), max_batch=16, max_wait_ms=2.0)
    result_pil = proc.process_image(image_pil, mask_pil)       # callable from any number of threads
    proc.close()

Everything except the forward itself (``run_batch``) is plain host logic and is unit-tested on the CPU with a fake forward.
"""
import threading
import time
from collections import OrderedDict, deque

import numpy as np


class _Request:
    __slots__ = ("key", "payload", "event", "result", "error", "t_submit")

    def __init__(self, key, payload):
        self.key, self.payload = key, payload
        self.event = threading.Event()
        self.result = self.error = None
        self.t_submit = time.monotonic()


class RequestBatcher:
    """Thread-safe request batching.

    ``run_batch(key, payloads) -> list of results`` (same length and order) is called from ONE worker thread with all the
    pending requests that share ``key`` (at most ``max_batch``). A request is dispatched as soon as ``max_batch`` requests of
    its key are pending or ``max_wait_ms`` after it was submitted, whichever comes first; keys are served oldest request first.
    ``submit`` blocks the calling thread until its result is ready and re-raises the worker's exception for that batch.
    """

    def __init__(self, run_batch, max_batch=16, max_wait_ms=2.0):
        if max_batch < 1:
            raise ValueError("max_batch must be >= 1")
        self.run_batch, self.max_batch, self.max_wait = run_batch, int(max_batch), max_wait_ms / 1e3
        self._cv = threading.Condition()
        self._pending = OrderedDict()          # key -> deque of requests, keys in order of their oldest pending request
        self._closed = False
        self.batches = []                      # (key, size) of every dispatched batch (observability / tests)
        self._worker = threading.Thread(target=self._loop, name="sketchedit-batcher", daemon=True)
        self._worker.start()

    def submit(self, key, payload):
        req = _Request(key, payload)
        with self._cv:
            if self._closed:
                raise RuntimeError("RequestBatcher is closed")
            self._pending.setdefault(key, deque()).append(req)
            self._cv.notify_all()
        req.event.wait()
        if req.error is not None:
            raise req.error
        return req.result

    def close(self):
        with self._cv:
            self._closed = True
            self._cv.notify_all()
        self._worker.join()

    # -- worker
    def _take(self):
        """Under the lock: the next batch to run, or (None, seconds to sleep) / (None, None) when closed and drained."""
        now = time.monotonic()
        best_wait = None
        for key, q in self._pending.items():
            age = now - q[0].t_submit
            if len(q) >= self.max_batch or age >= self.max_wait or self._closed:
                reqs = [q.popleft() for _ in range(min(self.max_batch, len(q)))]
                if not q:
                    del self._pending[key]
                else:
                    self._pending.move_to_end(key)          # the rest of this key queues behind the other keys
                return reqs, None
            w = self.max_wait - age
            best_wait = w if best_wait is None else min(best_wait, w)
        if self._closed and not self._pending:
            return None, None
        return None, (best_wait if best_wait is not None else 3600.0)

    def _loop(self):
        while True:
            with self._cv:
                reqs, wait = self._take()
                while reqs is None:
                    if wait is None:
                        return
                    self._cv.wait(timeout=wait)
                    reqs, wait = self._take()
            key = reqs[0].key
            try:
                results = self.run_batch(key, [r.payload for r in reqs])
                if len(results) != len(reqs):
                    raise RuntimeError("run_batch returned %d results for %d requests" % (len(results), len(reqs)))
                for r, res in zip(reqs, results):
                    r.result = res
            except BaseException as e:      # noqa: BLE001 - delivered to every requester of this batch
                for r in reqs:
                    r.error = e
            self.batches.append((key, len(reqs)))
            for r in reqs:
                r.event.set()


def floor8(n):
    return n // 8 * 8


class DemoProcessor:
    """``process_image`` of the reference demo (demo.py:39-73) on the batched uint8 forward.

    Differences from the reference function, none of them numerical: it returns the PIL result instead of writing
    ``static/results/<name>``, and concurrent calls share forwards. ``precision``: 'bf16' | 'fp32' | 'fp32_direct'.
    """

    def __init__(self, model, precision=None, max_batch=16, max_wait_ms=2.0):
        import torch
        self._torch = torch
        self.model = model
        self.precision = precision or getattr(model, "precision", "bf16")
        self.engine = model.engine()
        self.batcher = RequestBatcher(self._run_batch, max_batch=max_batch, max_wait_ms=max_wait_ms)

    def close(self):
        self.batcher.close()

    def _run_batch(self, key, payloads):
        torch = self._torch
        img = torch.from_numpy(np.stack([p[0] for p in payloads])).cuda(non_blocking=True)     # [B,H,W,3] RGB uint8
        msk = torch.from_numpy(np.stack([p[1] for p in payloads])).cuda(non_blocking=True)     # [B,H,W] uint8 (> 0 = stroke)
        with torch.no_grad():
            bgr, _ = self.engine.inference_u8(img, msk, precision=self.precision)
        rgb = bgr.cpu().numpy()[..., ::-1]                                                     # demo.py keeps RGB (test.py swaps to BGR)
        return [np.ascontiguousarray(rgb[i]) for i in range(len(payloads))]

    def process_image(self, img, mask):
        """img: PIL image; mask: PIL 'L' image of the same size (non-zero = sketch stroke). Returns the edited PIL image at the
        input's size. Sizes are floored to a multiple of 8 for the network exactly like demo.py:43."""
        from PIL import Image
        img = img.convert("RGB")
        w_raw, h_raw = img.size
        h_t, w_t = floor8(h_raw), floor8(w_raw)
        if h_t < 16 or w_t < 16:
            raise ValueError("image smaller than 16x16 (two stride-2 convolutions, 4x4 mask pool, stride-2 patch grid)")
        img_t = np.ascontiguousarray(np.array(img.resize((w_t, h_t))), dtype=np.uint8)
        mask_t = np.array(mask.resize((w_t, h_t)))
        mask_t = np.ascontiguousarray((mask_t > 0).astype(np.uint8) * 255)
        out = self.batcher.submit((h_t, w_t), (img_t, mask_t))
        return Image.fromarray(out).resize((w_raw, h_raw))
