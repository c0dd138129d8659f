"""Which host-side factor moves the pipelined end-to-end number? (sampler period, pinned ring vs fresh allocation)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from argparse import Namespace
import bench, models
from sketchedit_b200 import synth

B = 128
opt = Namespace(gpu_ids=[0], isTrain=False, isSkip=True, netG="deepfillc2", init_type="xavier", init_variance=0.02, use_cam=True,
                pool_type="max", no_mask_cc=False, no_mask_coarse=False, joint_train_inp=True, model="editline2", precision="bf16")
model = models.create_model(opt)
model.netM.load_state_dict(synth.synth_state_dict("M"))
model.netG.load_state_dict(synth.synth_state_dict("G"))
model.eval()
img_h, sk_h = bench.make_inputs(B)
img_h, sk_h = img_h.pin_memory(), sk_h.pin_memory()
comp_h = torch.empty(B, 3, 256, 256).pin_memory()
mask_h = torch.empty(B, 1, 256, 256).pin_memory()


def serial(n):
    for _ in range(n):
        with torch.no_grad():
            c, m = model({"image": img_h, "mask": sk_h}, mode="inference")
        comp_h.copy_(c, non_blocking=True)
        mask_h.copy_(m, non_blocking=True)
        torch.cuda.current_stream().synchronize()


def stream(n, ring):
    with torch.no_grad():
        for c, m in model.inference_stream(({"image": img_h, "mask": sk_h} for _ in range(n)), pinned_ring=ring):
            pass


def timed(fn, n=10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn(n)
    torch.cuda.synchronize()
    return B * n / (time.perf_counter() - t0)


serial(2); stream(3, True); stream(3, False)
for period in (None, 200, 25):
    s = None
    if period:
        bench.ClockSampler.PERIOD_MS = period
        s = bench.ClockSampler(0)
        s.start()
        time.sleep(0.3)
    print("sampler %s ms: serial %.0f  stream(ring) %.0f  stream(fresh) %.0f  stream(ring) again %.0f img/s" % (
        period, timed(serial), timed(lambda n: stream(n, True)), timed(lambda n: stream(n, False)), timed(lambda n: stream(n, True))), flush=True)
    if s:
        s.stop()
