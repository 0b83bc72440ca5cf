"""A/B timing of the layer-level C-ABI entry points across several builds of the library on the SAME box (box-to-box
clock / power differences are larger than most kernel changes).  usage: python tools/ab_layers.py lib1.so lib2.so ..."""
import ctypes as C
import sys

import torch

dev = torch.device("cuda:0")
vp, ll, ci = C.c_void_p, C.c_longlong, C.c_int
libs = []
for path in sys.argv[1:]:
    L = C.CDLL(path)
    L.vmb_conv3x3_relu.argtypes = [vp, vp, vp, vp, ll, ci, ci, ci, ci, ci, vp]
    L.vmb_linear.argtypes = [vp, vp, vp, vp, ci, ci, ll, ci, ci, vp]
    L.vmb_conv1_relu_pool.argtypes = [vp, vp, vp, vp, ll, vp]
    L.vmb_logmel.argtypes = [vp, ll, ll, ll, ll, vp, vp]
    libs.append((path.split("/")[-1], L))
n = 2560
st = lambda: torch.cuda.current_stream().cuda_stream
cases = []
for (H, W, Cin, Cout, pool) in [(48, 32, 64, 128, 1), (24, 16, 128, 256, 0), (24, 16, 256, 256, 1), (12, 8, 256, 512, 0), (12, 8, 512, 512, 1)]:
    x = torch.randn(n, H, W, Cin, device=dev).bfloat16()
    w = (torch.randn(Cout, 9 * Cin, device=dev) * 0.02).bfloat16()
    b = torch.randn(Cout, device=dev)
    o = torch.empty(n, H // (2 if pool else 1), W // (2 if pool else 1), Cout, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * n * H * W * Cout * 9 * Cin
    cases.append((f"conv {H}x{W} {Cin}->{Cout} p{pool}", fl,
                  lambda L, x=x, w=w, b=b, o=o, H=H, W=W, Cin=Cin, Cout=Cout, pool=pool:
                  L.vmb_conv3x3_relu(x.data_ptr(), w.data_ptr(), b.data_ptr(), o.data_ptr(), n, H, W, Cin, Cout, pool, st())))
for (M, N, K) in [(2560, 4096, 12288), (2560, 4096, 4096)]:
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    b = torch.randn(N, device=dev)
    o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    cases.append((f"linear {M}x{N}x{K}", 2.0 * M * N * K,
                  lambda L, a=a, w=w, b=b, o=o, M=M, N=N, K=K: L.vmb_linear(a.data_ptr(), w.data_ptr(), b.data_ptr(), o.data_ptr(), 0, 1, M, N, K, st())))
ex = torch.randn(n, 96, 64, device=dev)
w1, b1 = torch.randn(64, 9, device=dev), torch.randn(64, device=dev)
o1 = torch.empty(n, 48, 32, 64, device=dev, dtype=torch.bfloat16)
cases.append(("conv1", 2.0 * n * 96 * 64 * 64 * 9, lambda L: L.vmb_conv1_relu_pool(ex.data_ptr(), w1.data_ptr(), b1.data_ptr(), o1.data_ptr(), n, st())))
wave = (torch.rand(256, 160000, device=dev) * 2 - 1)
lm = torch.empty(256, 960, 64, device=dev)
cases.append(("logmel 256 clips", 256 * (2.0 * 400 * 514 + 2.0 * 257 * 64) * 998,
              lambda L: L.vmb_logmel(wave.data_ptr(), 256, 160000, 160000, 960, lm.data_ptr(), st())))
for name, fl, fn in cases:
    res = {nm: [] for nm, _ in libs}
    for rep in range(4):                       # interleave the builds
        for nm, L in libs:
            for _ in range(3):
                assert fn(L) == 0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn(L)
            e1.record()
            torch.cuda.synchronize()
            res[nm].append(e0.elapsed_time(e1) / 20)
    print(f"{name:28s} " + "  ".join(f"{nm}: {min(v):.4f} ms ({fl / min(v) / 1e9:7.1f} TF)" for nm, v in res.items()), flush=True)
