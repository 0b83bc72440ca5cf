import ctypes as C, sys, torch
L = C.CDLL(sys.argv[1]); vp, ll, ci = C.c_void_p, C.c_longlong, C.c_int
L.vmb_conv3x3_relu.argtypes = [vp, vp, vp, vp, ll, ci, ci, ci, ci, ci, vp]
dev = torch.device("cuda:0"); n = 2560
for (H, W, Cin, Cout, pool) in [(48, 32, 64, 128, 1), (48, 32, 64, 256, 1), (48, 32, 64, 256, 0), (48, 32, 64, 128, 0)]:
    x = torch.randn(n, H, W, Cin, device=dev).bfloat16(); w = (torch.randn(Cout, 9 * Cin, device=dev) * 0.02).bfloat16(); b = torch.randn(Cout, device=dev)
    o = torch.empty(n, H // (2 if pool else 1), W // (2 if pool else 1), Cout, device=dev, dtype=torch.bfloat16)
    fn = lambda: L.vmb_conv3x3_relu(x.data_ptr(), w.data_ptr(), b.data_ptr(), o.data_ptr(), n, H, W, Cin, Cout, pool, torch.cuda.current_stream().cuda_stream)
    for _ in range(3): assert fn() == 0
    best = 1e9
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) / 20)
    print(f"conv {H}x{W} {Cin}->{Cout} p{pool}: {best:.4f} ms {2.0*n*H*W*Cout*9*Cin/best/1e9:.1f} TF", flush=True)
