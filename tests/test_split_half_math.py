"""CPU restatement of the arithmetic the fp32-on-tensor-cores mode runs (se_gemm_split.cu, se_conv_c8.cu split-half twins), checked
against the oracle: host logic only (torch-CPU emulation of fp16 hi/lo operands with fp32 accumulation), no GPU, no product code.

  * the attention as two explicit GEMMs over patch matrices + fold-sum (the formulation of se_gemm_split.cu) equals the oracle's
    contextual attention (reference models/networks/splitcam.py:37-108,132-174);
  * three fp16 products hi*hi + hi*lo + lo*hi with power-of-two operand scaling reproduce the fp32 result to ~1e-6, and the scaling
    matters: without it the lo halves fall into fp16's subnormal range.
"""
import torch
import torch.nn.functional as F

from oracle import sketchedit_oracle as O

S_Q, S_K, S_P = 64.0, 32768.0, 16384.0        # se_gemm_split.cu: kScaleQ (= kSplitActScale), kScaleK, kScaleP


def split(x, scale):
    s = (x * scale).clamp(-65000.0, 65000.0)
    hi = s.half()
    lo = (s - hi.float()).half()
    return hi.float(), lo.float()


def gemm3(a, b, sa, sb):
    """a [.., M, K] @ b [.., N, K]^T with split-half operands: products of two fp16 numbers are exact in fp32; fp32 accumulation."""
    ah, al = split(a, sa)
    bh, bl = split(b, sb)
    acc = ah @ bh.transpose(-1, -2) + ah @ bl.transpose(-1, -2) + al @ bh.transpose(-1, -2)
    return acc / (sa * sb)


def patches_uvc(x):
    """x [B, C, h, w] -> [B, L, (u, v, c)] raw 4x4 / stride-2 patches in the column order of se_gemm_split.cu (tap-major, channel-minor)."""
    B, C, h, w = x.shape
    p = F.unfold(x, kernel_size=4, stride=2)                       # [B, C*16, L], rows ordered (c, u, v)
    L = p.shape[-1]
    return p.view(B, C, 16, L).permute(0, 3, 2, 1).reshape(B, L, 16 * C)


def attention_as_gemms(feat, mask_s, scaled=True):
    B, C, h, w = feat.shape
    hs, ws = (h - 4) // 2 + 1, (w - 4) // 2 + 1
    sq, sk, sp = (S_Q, S_K, S_P) if scaled else (1.0, 1.0, 1.0)
    rnorm = 1.0 / torch.sqrt((feat ** 2).sum((2, 3)) + 1e-8)        # [B, C]
    Q = patches_uvc(feat)                                           # [B, L, 16C]
    K = Q * rnorm.repeat(1, 16)[:, None, :]                         # channel-minor columns: rnorm tiles 16 times
    valid = 1.0 - mask_s
    m = (F.unfold(valid, kernel_size=4, stride=2).mean(1) > 0.1).float()      # [B, L] per key
    S = 10.0 * m[:, None, :] * gemm3(Q, K, sq, sk)                  # [B, N, L]
    P = torch.softmax(S, dim=2)
    Ot = gemm3(P, Q.transpose(1, 2), sp, sq)                        # [B, N, 16C]: P [N, L] @ (Q^T [16C, L])^T
    out = torch.zeros(B, h, w, C)
    Ov = Ot.view(B, hs, ws, 4, 4, C)
    for u in range(4):
        for v in range(4):
            out[:, u:u + 2 * hs:2, v:v + 2 * ws:2, :] += Ov[:, :, :, u, v, :]
    return out.permute(0, 3, 1, 2)


def _case(h, w, B=2, C=32, seed=0):
    g = torch.Generator().manual_seed(seed)
    feat = F.relu(torch.randn(B, C, h, w, generator=g) * 0.5)
    mask = torch.zeros(B, 1, 4 * h, 4 * w)
    mask[:, :, h:3 * h, w:2 * w + 8] = 1.0
    return feat, F.avg_pool2d(mask, 4, 4)


def test_attention_as_two_gemms_and_fold_equals_the_oracle():
    for h, w in ((16, 16), (12, 20)):
        feat, mask_s = _case(h, w, seed=h * w)
        ref, _ = O.contextual_attention(feat, mask_s)
        got = attention_as_gemms(feat, mask_s)
        tol = 2e-5 * max(1.0, float(ref.abs().max()))
        assert float((got - ref).abs().max()) <= tol, (float((got - ref).abs().max()), tol)


def test_operand_scaling_keeps_the_lo_halves_out_of_the_fp16_subnormals():
    g = torch.Generator().manual_seed(7)
    a = torch.randn(64, 864, generator=g) * 0.05                   # activations of a 96-channel 3x3 layer (typical magnitudes)
    wgt = torch.randn(192, 864, generator=g) * 0.01                # weights of that size
    ref = (a.double() @ wgt.double().t()).float()
    wscale = 2.0 ** (13 - int(torch.floor(torch.log2(wgt.abs().max()))))          # pack_class: largest weight in [8192, 16384)
    err_scaled = float((gemm3(a, wgt, 64.0, wscale) - ref).abs().max())
    err_plain = float((gemm3(a, wgt, 1.0, 1.0) - ref).abs().max())
    err_fp32 = float(((a @ wgt.t()) - ref).abs().max())
    assert err_scaled <= 4 * max(err_fp32, 1e-7), (err_scaled, err_fp32)            # as good as an fp32 GEMM
    assert err_plain >= 2 * err_scaled, (err_plain, err_scaled)                     # the unscaled split loses much of its lo half
