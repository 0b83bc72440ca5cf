"""Compare the layer-level entry points of two builds on identical inputs (bit-for-bit expected when only the
schedule differs).  usage: python tools/cmp_libs.py ref.so new.so"""
import ctypes as C
import sys

import torch

dev = torch.device("cuda:0")
vp, ll, ci = C.c_void_p, C.c_longlong, C.c_int
libs = []
for path in sys.argv[1:3]:
    L = C.CDLL(path)
    L.vmb_conv3x3_relu.argtypes = [vp, vp, vp, vp, ll, ci, ci, ci, ci, ci, vp]
    L.vmb_linear.argtypes = [vp, vp, vp, vp, ci, ci, ll, ci, ci, vp]
    L.vmb_last_error.restype = C.c_char_p
    libs.append(L)
st = lambda: torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
bad = 0
for n in (1, 3, 10, 77, 640):
    for (H, W, Cin, Cout, pool) in [(24, 16, 128, 256, 0), (24, 16, 256, 256, 1), (12, 8, 256, 512, 0), (12, 8, 512, 512, 1),
                                     (48, 32, 64, 128, 1)]:
        x = torch.randn(n, H, W, Cin, device=dev).bfloat16()
        w = (torch.randn(Cout, 9 * Cin, device=dev) * 0.03).bfloat16()
        b = torch.randn(Cout, device=dev)
        outs = []
        for L in libs:
            o = torch.full((n, H // (2 if pool else 1), W // (2 if pool else 1), Cout), -7.0, device=dev, dtype=torch.bfloat16)
            rc = L.vmb_conv3x3_relu(x.data_ptr(), w.data_ptr(), b.data_ptr(), o.data_ptr(), n, H, W, Cin, Cout, pool, st())
            assert rc == 0, L.vmb_last_error()
            torch.cuda.synchronize()
            outs.append(o)
        same = torch.equal(outs[0], outs[1])
        d = (outs[0].float() - outs[1].float()).abs().max().item()
        print(f"conv n={n} {H}x{W} {Cin}->{Cout} p{pool}: equal={same} maxdiff={d:.3g}", flush=True)
        bad += not same
for (M, N, K) in [(1, 4096, 4096), (130, 4096, 12288), (256, 4096, 4096), (300, 256, 128), (2560, 4096, 12288)]:
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    b = torch.randn(N, device=dev)
    outs = []
    for L in libs:
        o = torch.full((M, N), -7.0, device=dev, dtype=torch.bfloat16)
        rc = L.vmb_linear(a.data_ptr(), w.data_ptr(), b.data_ptr(), o.data_ptr(), 0, 1, M, N, K, st())
        assert rc == 0, L.vmb_last_error()
        torch.cuda.synchronize()
        outs.append(o)
    same = torch.equal(outs[0], outs[1])
    d = (outs[0].float() - outs[1].float()).abs().max().item()
    print(f"linear {M}x{N}x{K}: equal={same} maxdiff={d:.3g}", flush=True)
    bad += not same
print("MISMATCHES", bad)
sys.exit(1 if bad else 0)
