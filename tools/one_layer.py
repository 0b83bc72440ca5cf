"""Run one conv layer shape through vmb_conv3x3_relu a few times (target for a single-kernel ncu capture).
usage: python tools/one_layer.py LIB.so H W Cin Cout pool [n_images]"""
import ctypes as C
import sys

import torch

L = C.CDLL(sys.argv[1])
H, W, Cin, Cout, pool = map(int, sys.argv[2:7])
n = int(sys.argv[7]) if len(sys.argv) > 7 else 2560
vp, ll, ci = C.c_void_p, C.c_longlong, C.c_int
L.vmb_conv3x3_relu.argtypes = [vp, vp, vp, vp, ll, ci, ci, ci, ci, ci, vp]
dev = torch.device("cuda:0")
x = torch.randn(n, H, W, Cin, device=dev).bfloat16()
w = (torch.randn(Cout, 9 * Cin, device=dev) * 0.02).bfloat16()
b = torch.randn(Cout, device=dev)
o = torch.empty(n, H // (2 if pool else 1), W // (2 if pool else 1), Cout, device=dev, dtype=torch.bfloat16)
for _ in range(4):
    assert L.vmb_conv3x3_relu(x.data_ptr(), w.data_ptr(), b.data_ptr(), o.data_ptr(), n, H, W, Cin, Cout, pool,
                              torch.cuda.current_stream().cuda_stream) == 0
torch.cuda.synchronize()
print("ok", float(o.float().abs().mean()))
