"""Serving path (reference demo.py:39-73): request batching (host logic, CPU) and the demo's process_image flow (GPU)."""
import threading
import time

import numpy as np
import pytest
import torch

from sketchedit_b200.serving import RequestBatcher, floor8


def test_floor8():
    assert [floor8(n) for n in (8, 15, 16, 250, 256, 641)] == [8, 8, 16, 248, 256, 640]


def test_batcher_groups_concurrent_requests_by_key_and_routes_results():
    calls = []

    def run(key, payloads):
        calls.append((key, list(payloads)))
        time.sleep(0.01)                                   # a "forward": lets the next requests pile up
        return [(key, p * 10) for p in payloads]

    b = RequestBatcher(run, max_batch=4, max_wait_ms=20.0)
    out = {}

    def worker(i):
        out[i] = b.submit("A" if i % 3 else "B", i)

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(24)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    b.close()
    assert out == {i: ("A" if i % 3 else "B", i * 10) for i in range(24)}          # every requester got ITS result
    assert all(len(p) <= 4 for _, p in calls)                                        # max_batch respected
    assert all(len({("A" if x % 3 else "B") for x in p}) == 1 for _, p in calls)      # a batch never mixes keys (= input sizes)
    assert sum(len(p) for _, p in calls) == 24 and len(calls) < 24                   # requests really were batched
    assert b.batches == [(k, len(p)) for k, p in calls]


def test_batcher_dispatches_a_lone_request_after_the_wait_window():
    b = RequestBatcher(lambda k, ps: [p + 1 for p in ps], max_batch=8, max_wait_ms=5.0)
    t0 = time.monotonic()
    assert b.submit((64, 64), 41) == 42
    assert time.monotonic() - t0 < 1.0
    b.close()
    with pytest.raises(RuntimeError):
        b.submit((64, 64), 1)


def test_batcher_full_batch_does_not_wait_and_errors_reach_every_requester():
    def run(key, payloads):
        if key == "bad":
            raise ValueError("forward failed")
        return payloads

    b = RequestBatcher(run, max_batch=2, max_wait_ms=10_000.0)       # only a full batch can be dispatched quickly
    res, errs = [], []

    def ok(i):
        res.append(b.submit("ok", i))

    def bad(i):
        try:
            b.submit("bad", i)
        except ValueError as e:
            errs.append(str(e))

    ts = [threading.Thread(target=ok, args=(i,)) for i in range(2)] + [threading.Thread(target=bad, args=(i,)) for i in range(2)]
    t0 = time.monotonic()
    [t.start() for t in ts]
    [t.join(timeout=5.0) for t in ts]
    assert time.monotonic() - t0 < 5.0 and sorted(res) == [0, 1] and errs == ["forward failed"] * 2
    b.close()


@pytest.mark.gpu
def test_process_image_matches_the_reference_demo_flow():
    """DemoProcessor.process_image == reference demo.py:39-73 evaluated with the CPU oracle: floor to a multiple of 8, PIL resize,
    (x/255-0.5)/0.5 and mask > 0, forward, clamp, (g+1)/2*255 truncated to uint8 (RGB kept), PIL resize back. Eight requests of two
    different sizes from eight threads share forwards. fp32 path: the uint8 result may differ by one level where a float lands
    within ~1e-5 of an integer."""
    from PIL import Image

    from oracle import sketchedit_oracle as O
    from sketchedit_b200 import synth
    from sketchedit_b200.serving import DemoProcessor
    from tests.test_gpu_configs import _model
    from tests.util_parity import weights
    model = _model("fp32_direct")      # the flow is under test, not the arithmetic mode: no threshold ties to reason about
    proc = DemoProcessor(model, max_batch=4, max_wait_ms=50.0)
    WM, WG = weights()
    rs = np.random.RandomState(11)
    cases = []
    for i in range(8):
        h, w = ((75, 100), (64, 90))[i % 2]
        img = Image.fromarray(rs.randint(0, 256, (h, w, 3), dtype=np.uint8))
        m = np.zeros((h, w), np.uint8)
        m[10 + i:40, 20:22 + i] = 255
        cases.append((img, Image.fromarray(m)))
    got = [None] * 8

    def worker(i):
        got[i] = proc.process_image(*cases[i])

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    proc.close()
    assert sum(n for _, n in proc.batcher.batches) == 8 and len(proc.batcher.batches) < 8       # batched, two sizes never mixed
    assert {k for k, _ in proc.batcher.batches} == {(72, 96), (64, 88)}
    for (img, mask), res in zip(cases, got):
        w_raw, h_raw = img.size
        h_t, w_t = h_raw // 8 * 8, w_raw // 8 * 8
        it = torch.from_numpy(np.array(img.resize((w_t, h_t))).transpose(2, 0, 1)).float()
        it = ((it / 255 - 0.5) / 0.5)[None]
        mt = (torch.from_numpy(np.array(mask.resize((w_t, h_t)))).float() > 0).float()[None, None]
        ref = O.inference(WM, WG, it, mt)["composed"]
        ref = ((torch.clamp(ref, -1, 1) + 1) / 2 * 255).numpy().astype(np.uint8)[0].transpose(1, 2, 0)
        want = np.array(Image.fromarray(ref).resize((w_raw, h_raw)))
        assert res.size == (w_raw, h_raw)
        d = np.abs(np.array(res).astype(int) - want.astype(int))
        assert d.max() <= 2 and (d != 0).mean() <= 5e-3, (d.max(), (d != 0).mean())
