"""Per-tile timelines (SE_TC_DEBUG=2, se_conv_c8.cu C8_TRACE) of every tcgen05 conv launch of ONE forward at the bench shape:
    SE_TC_DEBUG=2 python tools/trace_forward.py [B] 2> gpurun_out/trace_full.log"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sketchedit_b200 import synth
from tests.util_parity import engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
img, sk = synth.synth_inputs(B, 256, 256, seed=3)
eng = engine()
eng.inference(img.cuda(), sk.cuda(), precision="bf16")
torch.cuda.synchronize()
sys.stderr.write("==== traced forward\n")
eng.inference(img.cuda(), sk.cuda(), precision="bf16")
torch.cuda.synchronize()
