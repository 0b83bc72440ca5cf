"""Geometry sweep of vmb_conv3x3_relu against torch (same bf16-rounded operands, fp32 math): every kernel variant
(single-CTA / pair / big box / halo / MT) gets hit by some combination.  usage: python tools/conv_sweep.py"""
import itertools
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, __import__("glob").glob(__file__.rsplit("/", 2)[0] + "/audio-*_b200")[0])
from b200 import _lib, engine  # noqa: E402

dev = torch.device("cuda:0")
L = _lib.lib()
bad = 0
g = torch.Generator().manual_seed(0)
for H, W, Cin, Cout, n, pool in itertools.product((8, 12, 24, 48), (8, 16, 32, 48), (64, 128), (128, 256), (1, 7, 60, 400), (0, 1)):
    if n * H * W * Cout * Cin > 2.5e11:
        continue
    x = torch.randn(n, Cin, H, W, generator=g).to(dev).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * (2.0 / (9 * Cin)) ** 0.5).to(dev).bfloat16()
    b = (torch.randn(Cout, generator=g) * 0.1).to(dev)
    xn = x.permute(0, 2, 3, 1).contiguous()
    wk = w.permute(0, 2, 3, 1).contiguous().reshape(Cout, 9 * Cin)
    out = torch.full((n, H // 2, W // 2, Cout) if pool else (n, H, W, Cout), float("nan"), device=dev, dtype=torch.bfloat16)
    rc = L.vmb_conv3x3_relu(xn.data_ptr(), wk.data_ptr(), b.data_ptr(), out.data_ptr(), n, H, W, Cin, Cout, pool, engine.stream_ptr())
    if rc:
        print(f"H={H} W={W} Cin={Cin} Cout={Cout} n={n} pool={pool}: refused: {L.vmb_last_error().decode()}")
        continue
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(x.float(), w.float(), b, padding=1))
    if pool:
        ref = F.max_pool2d(ref, 2, 2)
    got = out.permute(0, 3, 1, 2).float()
    err = ((got - ref).abs().max() / ref.abs().max()).item()
    ok = err < 0.006 and not torch.isnan(got).any()
    bad += not ok
    if not ok:
        print(f"H={H} W={W} Cin={Cin} Cout={Cout} n={n} pool={pool}: rel-max-err {err:.3e}  <-- FAIL")
print("conv sweep done, failures:", bad)
sys.exit(1 if bad else 0)
