"""Benchmark of the SketchEdit generator forward pass (BASELINE.json metric: images/sec, 256x256
CelebA-HQ-shaped inputs, synthetic seeded weights of the real architecture).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--dtype bf16|fp32] [--batch B]

One "step" = one forward of `model(data, mode='inference')` (netM + threshold + netG incl. contextual
attention + blend) over one batch. Default workload = BASELINE.json configs[2]: batch 128 per GPU, bf16
tensor-core path (== the per-GPU shard of configs[4], 1024 images over 8 GPUs; weak scaling). `--dtype fp32
--batch 32` runs configs[1] (fp32 parity path).

Prints ONE JSON line (rank 0). `value` is whole-job throughput with inputs resident in HBM; `e2e` is the same
metric through the reference-facing module API with pinned host tensors in and host tensors out.
`--impl reference` times the CPU oracle port of the reference path on the host cores (the reference itself is
Python + PyTorch and is not shipped to the GPU box).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

H = W = 256
METRIC = "images/sec 256x256 CelebA-HQ generator fwd"
UNIT = "images/s"


def workload_name(dtype, batch, n):
    return "CelebA-HQ 256x256 generator fwd (netM+netG+CAM), batch %d/GPU x %d GPU, %s, synthetic weights+inputs" % (batch, n, dtype)


def make_inputs(batch):
    from sketchedit_b200 import synth
    base_img, base_sk = synth.synth_inputs(8, H, W, seed=0)
    reps = (batch + 7) // 8
    return base_img.repeat(reps, 1, 1, 1)[:batch].contiguous(), base_sk.repeat(reps, 1, 1, 1)[:batch].contiguous()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 25 ms during the timed regions (device-resident + e2e)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    PERIOD_MS = 25

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", str(self.PERIOD_MS)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()          # exact PID we started
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


_BEST_THREADS = None


def best_cpu_threads(WM, WG):
    """torch's CPU convolutions stop scaling (and regress) well before 100+ threads at batch 4: pick the
    fastest intra-op thread count from a short calibration so the CPU arm is not handicapped."""
    global _BEST_THREADS
    if _BEST_THREADS is not None:
        return _BEST_THREADS
    from oracle import sketchedit_oracle as O
    ncpu = os.cpu_count() or 1
    cands = sorted({t for t in (8, 16, 32, 64, ncpu) if t <= ncpu})
    img, sk = make_inputs(2)
    best, best_t = cands[0], float("inf")
    for t in cands:
        torch.set_num_threads(t)
        O.inference(WM, WG, img, sk)
        t0 = time.perf_counter()
        O.inference(WM, WG, img, sk)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = t, dt
    _BEST_THREADS = best
    return best


def cpu_oracle_throughput(n_img, passes, warm):
    """images/s of the CPU oracle port of the reference path on the host cores, bounded sample."""
    from oracle import sketchedit_oracle as O
    from sketchedit_b200 import synth
    WM, WG = synth.synth_state_dict("M"), synth.synth_state_dict("G")
    torch.set_num_threads(best_cpu_threads(WM, WG))
    img, sk = make_inputs(n_img)
    for _ in range(warm):
        O.inference(WM, WG, img, sk)
    t0 = time.perf_counter()
    for _ in range(passes):
        O.inference(WM, WG, img, sk)
    dt = time.perf_counter() - t0
    return n_img * passes / dt, dt / passes


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_img = 4     # one step = a 4-image slice of the workload batch (bounded sample)
    ips, step_s = cpu_oracle_throughput(n_img, args.steps, args.warmup)
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic",
        "config": {"workload": workload_name(args.dtype, args.batch, args.gpus), "step": "%d-image slice per step on CPU" % n_img},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "%d steps x %d images 256x256, torch CPU fp32 oracle port, %d threads (fastest of a calibration over 8..%d)" % (args.steps, n_img, torch.get_num_threads(), cores)},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# one ncu capture of the dominant kernel at the default workload (profiles/r01_dram_traffic_b128.md): 101.3 MB read + 60.8 MB written
DOMINANT_KERNEL_DRAM_BYTES = 162.1e6


def run_b200(args):
    import torch.distributed as dist
    from argparse import Namespace

    import models
    from sketchedit_b200 import _lib, synth
    from sketchedit_b200.arch import cam_flops_per_image, conv_flops_per_image

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU for --impl b200 (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world)
    B = args.batch
    prec = args.dtype

    opt = Namespace(gpu_ids=[local], isTrain=False, isSkip=True, netG="deepfillc2", init_type="xavier", init_variance=0.02,
                    use_cam=True, pool_type="max", no_mask_cc=False, no_mask_coarse=False, joint_train_inp=True,
                    model="editline2", precision=prec)
    model = models.create_model(opt)
    model.netM.load_state_dict(synth.synth_state_dict("M"))
    model.netG.load_state_dict(synth.synth_state_dict("G"))
    model.eval()
    eng = model.engine()
    lib = _lib.load()

    img_h, sk_h = make_inputs(B)                     # each rank: its own contiguous shard (same synthetic content)
    img_h, sk_h = img_h.pin_memory(), sk_h.pin_memory()
    img_d, sk_d = img_h.cuda(non_blocking=True), sk_h.cuda(non_blocking=True)
    gathered = torch.empty(world * B, 4, H, W, device="cuda") if world > 1 else None

    from sketchedit_b200 import parallel

    def step_device():
        composed, mask, _ = eng.inference(img_d, sk_d, precision=prec)
        if world > 1:   # the path's single collective: all-gather of the packed output tiles over NVLink
            parallel.all_gather_outputs(composed, mask, out=gathered)
        return composed, mask

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    launches_per_step = eng.launches()

    # ---------------- timed region: device-resident inputs
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if prec == "bf16":
        lib.se_tc_timing_enable(1)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    tc_ms, tc_n, tc_fl = 0.0, 0, 0.0
    if prec == "bf16":
        import ctypes
        a, b, c = ctypes.c_double(), ctypes.c_int(), ctypes.c_double()
        _lib.check(lib.se_tc_time_ms(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        tc_ms, tc_n, tc_fl = a.value, b.value, c.value
        lib.se_tc_timing_enable(0)
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms / 1e3)

    # ---------------- e2e: reference-facing module API, pinned host tensors in, host tensors out
    comp_h = torch.empty(B, 3, H, W).pin_memory()
    mask_h = torch.empty(B, 1, H, W).pin_memory()

    def step_e2e():
        with torch.no_grad():
            composed, mask = model({"image": img_h, "mask": sk_h}, mode="inference")
        comp_h.copy_(composed, non_blocking=True)
        mask_h.copy_(mask, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def timed(fn):
        barrier()
        t0 = time.perf_counter()
        fn()
        barrier()
        t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # (a) one blocking call per batch, as the reference's test.py loop does
    for _ in range(2):
        step_e2e()
    e2e_serial = world * B * args.steps / timed(lambda: [step_e2e() for _ in range(args.steps)])

    # (b) the package's pipelined loop over the same batches (models.EditLine2Model.inference_stream): every step
    # still copies its own inputs host->device and its own outputs device->host, on side streams
    def run_stream(n):
        with torch.no_grad():
            for comp, msk in model.inference_stream(({"image": img_h, "mask": sk_h} for _ in range(n))):
                pass
        return comp, msk

    run_stream(6)   # > depth + 2 batches: the pinned output ring is allocated (cudaHostAlloc is slow) before the timed loop
    e2e_value = world * B * args.steps / timed(lambda: run_stream(args.steps))
    clocks = sampler.stop() if rank == 0 else None
    h2d = B * 4 * H * W * 4
    d2h = B * 4 * H * W * 4

    if rank == 0:
        peaks, src = measured_peaks()
        roof = None
        if prec == "bf16" and tc_ms > 0:
            ach = tc_fl / (tc_ms / 1e3) / 1e12
            peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
            roof = {"bound": "tensor", "kernel": "conv_c8_kernel + conv_tc_kernel (every tcgen05 implicit-GEMM launch: 71 gated convs on channel-blocked activations, 5 attention GEMM-conv launches)",
                    "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": DOMINANT_KERNEL_DRAM_BYTES if (B == 128 and prec == "bf16") else None,
                    "traffic_note": "dram__bytes_read+write per launch of conv_c8_kernel<1,1,4,PAIR> (96->192, 31 launches/step) at batch 128 from one ncu capture (profiles/r01_dram_traffic_b128.md); algorithmic bytes of that launch: 201.7e6",
                    "peak_source": "%s bf16_tflops_sustained (kernel timed inside a long step)" % src,
                    "launches_timed": tc_n, "flops_convention": "algorithmic 2*MAC of the reference ops (SURVEY.md 8d), per launch summed",
                    "kernel_share_of_step": tc_ms / ms}
        cpu_ips, cpu_step = cpu_oracle_throughput(4, 2, 1) if world == 1 else (None, None)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": prec, "data": "synthetic",
            "config": {"workload": workload_name(prec, B, world), "global_batch": world * B, "H": H, "W": W,
                       "l2": "inputs + per-step activations (%.1f GB workspace) far exceed the 126 MB L2; no explicit flush" % (eng.workspace_bytes() / 1e9),
                       "parallelism": "dp%d (batch shards, one NCCL all-gather of outputs)" % world if world > 1 else "single GPU",
                       "algorithmic_gflop_per_image": (conv_flops_per_image(H, W) + cam_flops_per_image(H, W)) / 1e9},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "models.create_model(opt).inference_stream(batches): pinned CPU tensors in -> pinned CPU outputs, "
                           "per-step H2D/D2H on side streams overlapping compute",
                    "serial_value": e2e_serial,
                    "serial_api": "models.create_model(opt)(data, mode='inference') + .copy_ to pinned CPU, one blocking call per batch"},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roof,
        }
        if cpu_ips is not None:
            line["cpu_baseline"] = {"value": cpu_ips, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": "2 passes x 4 images 256x256 (1 warm-up), torch CPU fp32 oracle port of the reference forward, %d threads (fastest of a calibration over 8..%d host cores)" % (torch.get_num_threads(), os.cpu_count())}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=("b200", "reference"))
    ap.add_argument("--dtype", default="bf16", choices=("bf16", "fp32"))
    ap.add_argument("--batch", type=int, default=None, help="images per GPU per step (default 128 bf16 / 32 fp32)")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = 128 if args.dtype == "bf16" else 32
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
