import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sketchedit_b200 import synth
from tests.util_parity import engine
img, sk = synth.synth_inputs(2, 64, 96, seed=11)
m, s = engine().netM(img.cuda(), sk.cuda(), precision="bf16")
torch.cuda.synchronize()
print("mask nan", torch.isnan(m).sum().item(), "img nan", torch.isnan(s).sum().item())
