"""GPU bring-up probe: exercises the layer-level C-ABI entry points against torch references and prints
detailed diagnostics (not a pytest file; used while bringing the kernels up on a B200)."""
import ctypes as C
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "audio-classification-using-a-deep-cnn-combined-with-multi-level-attention_b200")
lib = C.CDLL(os.path.join(PKG, "b200", "libvggish_mla_b200.so"))
lib.vmb_last_error.restype = C.c_char_p
vp, ll, ci = C.c_void_p, C.c_longlong, C.c_int
lib.vmb_linear.argtypes = [vp, vp, vp, vp, ci, ci, ll, ci, ci, vp]
lib.vmb_conv3x3_relu.argtypes = [vp, vp, vp, vp, ll, ci, ci, ci, ci, ci, vp]
lib.vmb_conv1_relu_pool.argtypes = [vp, vp, vp, vp, ll, vp]
lib.vmb_logmel.argtypes = [vp, ll, ll, ll, ll, vp, vp]
lib.vmb_postprocess.argtypes = [vp, vp, vp, vp, vp, ll, vp]

dev = torch.device("cuda:0")
print("device:", torch.cuda.get_device_name(0), "arch", lib.vmb_device_arch(0), flush=True)


def st():
    return torch.cuda.current_stream().cuda_stream


def chk(rc, what):
    if rc != 0:
        print(f"!! {what} failed: {lib.vmb_last_error().decode()}", flush=True)
        return False
    torch.cuda.synchronize()
    return True


def report(name, got, ref):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-12
    cos = F.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
    bad = (err > 0.02 * denom).sum().item()
    print(f"{name}: max_abs_err={err.max().item():.4g} ref_max={denom:.4g} rel={err.max().item()/denom:.3g} "
          f"cos={cos:.6f} n_bad={bad}/{err.numel()} nan={torch.isnan(got).sum().item()}", flush=True)
    return err.max().item() / denom


def test_linear(M, N, K, out_f32, relu=1):
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = (torch.randn(M, K, generator=g) * 0.5).to(dev).bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev).bfloat16()
    b = torch.randn(N, generator=g).to(dev)
    out = torch.full((M, N), float("nan"), device=dev, dtype=torch.float32 if out_f32 else torch.bfloat16)
    ok = chk(lib.vmb_linear(a.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), out_f32, relu, M, N, K, st()),
             f"linear {M}x{N}x{K}")
    if not ok:
        return
    ref = a.float() @ w.float().t() + b
    if relu:
        ref = ref.relu()
    r = report(f"linear M={M} N={N} K={K} f32={out_f32}", out, ref)
    if r > 0.02:
        e = (out.float() - ref).abs()
        rows = (e.max(dim=1).values > 0.02 * ref.abs().max()).nonzero().flatten()
        cols = (e.max(dim=0).values > 0.02 * ref.abs().max()).nonzero().flatten()
        print("   bad rows (first 16):", rows[:16].tolist(), " count", rows.numel())
        print("   bad cols (first 16):", cols[:16].tolist(), " count", cols.numel())
        print("   got[0,:8]", out[0, :8].float().tolist())
        print("   ref[0,:8]", ref[0, :8].tolist())


def test_conv(n, H, W, Cin, Cout, pool):
    g = torch.Generator(device="cpu").manual_seed(H * 7 + Cin)
    x = torch.randn(n, Cin, H, W, generator=g).to(dev).bfloat16()          # NCHW reference layout
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * (2.0 / (9 * Cin)) ** 0.5).to(dev).bfloat16()
    b = (torch.randn(Cout, generator=g) * 0.1).to(dev)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    w_k = w.permute(0, 2, 3, 1).contiguous().reshape(Cout, 9 * Cin)         # (kh, kw, c)
    Ho, Wo = (H // 2, W // 2) if pool else (H, W)
    out = torch.full((n, Ho, Wo, Cout), float("nan"), device=dev, dtype=torch.bfloat16)
    ok = chk(lib.vmb_conv3x3_relu(x_nhwc.data_ptr(), w_k.data_ptr(), b.data_ptr(), out.data_ptr(), n, H, W, Cin, Cout,
                                  pool, st()), f"conv n={n} {H}x{W} {Cin}->{Cout} pool={pool}")
    if not ok:
        return
    ref = F.relu(F.conv2d(x.float(), w.float(), b, padding=1))
    if pool:
        ref = F.max_pool2d(ref, 2, 2)
    ref = ref.permute(0, 2, 3, 1)
    r = report(f"conv n={n} {H}x{W} {Cin}->{Cout} pool={pool}", out, ref)
    if r > 0.02:
        e = (out.float() - ref).abs().amax(dim=3)  # n, Ho, Wo
        print("   err map img0 (rows x cols, >thr marked):")
        thr = 0.02 * ref.abs().max()
        for y in range(min(Ho, 12)):
            print("   ", "".join("X" if e[0, y, xx] > thr else "." for xx in range(Wo)))
        print("   per-image max err:", e.amax(dim=(1, 2))[:8].tolist())


def test_conv1(n):
    g = torch.Generator(device="cpu").manual_seed(5)
    x = torch.randn(n, 96, 64, generator=g).to(dev)
    w = (torch.randn(64, 1, 3, 3, generator=g) * 0.47).to(dev)
    b = (torch.randn(64, generator=g) * 0.1).to(dev)
    out = torch.full((n, 48, 32, 64), float("nan"), device=dev, dtype=torch.bfloat16)
    if not chk(lib.vmb_conv1_relu_pool(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), n, st()), "conv1"):
        return
    ref = F.max_pool2d(F.relu(F.conv2d(x[:, None], w, b, padding=1)), 2, 2).permute(0, 2, 3, 1)
    report(f"conv1 n={n}", out, ref)


def test_logmel():
    import numpy as np
    sys.path.insert(0, ROOT)
    from oracle import frontend_np
    rng = np.random.default_rng(0)
    n_clips, ns = 3, 160000
    t = np.arange(ns) / 16000.0
    waves = np.stack([rng.uniform(-1, 1, ns),
                      0.5 * np.sin(2 * np.pi * (200 + 300 * t) * t) + 1e-2 * rng.standard_normal(ns),
                      rng.normal(0, 0.1, ns) * (1 + np.sin(2 * np.pi * 2 * t))]).astype(np.float32)
    wd = torch.from_numpy(waves).to(dev)
    nf = 998
    out = torch.full((n_clips, nf, 64), float("nan"), device=dev)
    if not chk(lib.vmb_logmel(wd.data_ptr(), n_clips, ns, ns, nf, out.data_ptr(), st()), "logmel"):
        return
    for i in range(n_clips):
        ref = frontend_np.log_mel_spectrogram(waves[i].astype(np.float64))
        err = np.abs(out[i].cpu().numpy().astype(np.float64) - ref)
        print(f"logmel clip {i}: max_abs_err={err.max():.3g} mean={err.mean():.3g} (target <= 1e-4)", flush=True)


def bench_conv(n, H, W, Cin, Cout, pool, iters=10):
    x = torch.randn(n, H, W, Cin, device=dev).bfloat16()
    w = (torch.randn(Cout, 9 * Cin, device=dev) * 0.02).bfloat16()
    b = torch.zeros(Cout, device=dev)
    Ho, Wo = (H // 2, W // 2) if pool else (H, W)
    out = torch.empty((n, Ho, Wo, Cout), device=dev, dtype=torch.bfloat16)
    args = (x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), n, H, W, Cin, Cout, pool, st())
    for _ in range(3):
        lib.vmb_conv3x3_relu(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        lib.vmb_conv3x3_relu(*args)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 2.0 * n * H * W * Cout * 9 * Cin
    print(f"bench conv n={n} {H}x{W} {Cin}->{Cout} pool={pool}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s", flush=True)


def bench_linear(M, N, K, iters=10):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    b = torch.zeros(N, device=dev)
    out = torch.empty((M, N), device=dev, dtype=torch.bfloat16)
    args = (a.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), 0, 1, M, N, K, st())
    for _ in range(3):
        lib.vmb_linear(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        lib.vmb_linear(*args)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"bench linear {M}x{N}x{K}: {ms:.3f} ms  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["linear", "conv", "conv1", "logmel", "bench"]
    if "linear" in which:
        test_linear(128, 128, 64, 1)
        test_linear(128, 256, 64, 1)
        test_linear(128, 128, 256, 0)
        test_linear(200, 256, 512, 0)
        test_linear(2560, 4096, 12288, 0)
        test_linear(2560, 128, 4096, 1)
    if "conv" in which:
        test_conv(2, 48, 32, 64, 128, 1)
        test_conv(2, 24, 16, 128, 256, 0)
        test_conv(2, 24, 16, 256, 256, 1)
        test_conv(5, 12, 8, 256, 512, 0)
        test_conv(5, 12, 8, 512, 512, 1)
        test_conv(40, 48, 32, 64, 128, 1)
    if "conv1" in which:
        test_conv1(7)
    if "logmel" in which:
        test_logmel()
    if "bench" in which:
        bench_conv(2560, 48, 32, 64, 128, 1)
        bench_conv(2560, 24, 16, 128, 256, 0)
        bench_conv(2560, 24, 16, 256, 256, 1)
        bench_conv(2560, 12, 8, 256, 512, 0)
        bench_conv(2560, 12, 8, 512, 512, 1)
        bench_linear(2560, 4096, 12288)
        bench_linear(2560, 4096, 4096)
    print("probe done", flush=True)
