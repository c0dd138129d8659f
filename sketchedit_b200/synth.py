"""Seeded synthetic checkpoints and inputs (no real CelebA-HQ / Places weights exist offline).

The reference's ``checkpoints/`` is empty and its download script needs the network
(reference download/download_model.sh:1-8), so every parity test and benchmark in this
repo uses state_dicts produced here. They have exactly the key names / shapes / dtypes
of ``latest_net_M.pth`` / ``latest_net_G.pth`` (reference util/util.py:214-225), so a
real checkpoint drops in unchanged.

PyTorch-default init makes activations decay and the predicted mask collapse to ~0.49
(SURVEY.md section 7.3-7); weights here use a larger gain and the mask head gets a spatially
structured response so that 10-40 % of pixels binarise to 1. The gain (2.3) is the largest
for which the random network is not chaotic: a bf16-rounded evaluation of the oracle stays
within ~6e-3 of the fp32 one (same margin the reference's own autocast shows), while the
attention is moderately peaked (mean max weight ~0.65).
numpy's legacy RandomState is used because its stream is frozen across numpy versions.
"""
import numpy as np
import torch

from .arch import NET_LAYERS


def synth_state_dict(net, seed=1234, gain=2.3):
    """Return an OrderedDict-like {name.weight, name.bias} of fp32 torch tensors (OIHW)."""
    rs = np.random.RandomState(seed + (0 if net == "M" else 1))
    sd = {}
    for l in NET_LAYERS[net]:
        fan_in = l.cin * l.k * l.k
        bound = gain / np.sqrt(fan_in)
        w = rs.uniform(-bound, bound, size=(l.cout, l.cin, l.k, l.k)).astype(np.float32)
        b = rs.uniform(-0.1, 0.1, size=(l.cout,)).astype(np.float32)
        if l.act is not None:
            # open the gates a little so signal survives 17 layers
            b[l.cout // 2:] += 1.0
        if net == "M" and l.name == "conv_mask_17":
            w *= 4.0          # spread the mask logits so the soft mask leaves 0.5 +- 0.01
            b[:] = 0.95       # 10-40 % of pixels binarise to 1 on the synthetic inputs
        sd[l.name + ".weight"] = torch.from_numpy(w)
        sd[l.name + ".bias"] = torch.from_numpy(b)
    return sd


def synth_inputs(B, H, W, seed=0):
    """image in [-1, 1] (smooth + noise), sketch in {0, 1} made of a few polylines."""
    rs = np.random.RandomState(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, H, dtype=np.float32),
                         np.linspace(0, 1, W, dtype=np.float32), indexing="ij")
    img = np.empty((B, 3, H, W), np.float32)
    sk = np.zeros((B, 1, H, W), np.float32)
    for b in range(B):
        for c in range(3):
            f = rs.uniform(1.0, 6.0, size=4)
            p = rs.uniform(0, 2 * np.pi, size=2)
            img[b, c] = 0.6 * np.sin(f[0] * xx * 6.28 + p[0]) * np.cos(f[1] * yy * 6.28 + p[1]) \
                + 0.25 * np.sin((f[2] * xx + f[3] * yy) * 6.28)
        img[b] += rs.uniform(-0.15, 0.15, size=(3, H, W)).astype(np.float32)
        for _ in range(3):
            # random-walk stroke
            y, x = rs.randint(H // 8, H - H // 8), rs.randint(W // 8, W - W // 8)
            ang = rs.uniform(0, 2 * np.pi)
            for _ in range(max(H, W) // 2):
                sk[b, 0, int(y) % H, int(x) % W] = 1.0
                ang += rs.uniform(-0.4, 0.4)
                y += np.sin(ang)
                x += np.cos(ang)
    img = np.clip(img, -1, 1)
    return torch.from_numpy(img), torch.from_numpy(sk)
