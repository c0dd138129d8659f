"""Role-timer probe of the tcgen05 conv kernel: SE_TC_DEBUG=1 python tools/tc_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.util_parity import engine, rand_act
from sketchedit_b200.arch import layer_map
B = int(os.environ.get("PB", "32"))
cases = [("M", "conv5", 64, 64), ("G", "conv11", 64, 64), ("M", "conv3", 128, 128), ("M", "conv1", 256, 256),
         ("M", "conv13_upsample_conv", 64, 64), ("M", "conv15_upsample_conv", 128, 128), ("M", "conv16", 256, 256),
         ("M", "conv2_downsample", 256, 256), ("M", "conv4_downsample", 128, 128), ("G", "xconv4_downsample", 128, 128), ("M", "conv10_atrous", 64, 64)]
sel = os.environ.get("SE_PROBE_CASES")
if sel:
    cases = [c for c in cases if c[1] in sel.split(",")]
for net, name, H, W in cases:
    spec = layer_map(net)[name]
    x = rand_act((B, spec.cin, H, W), seed=1).cuda()
    for _ in range(2):
        sys.stderr.write("== %s.%s\n" % (net, name))
        engine().gated_conv(net, name, x, precision="bf16")
        torch.cuda.synchronize()
