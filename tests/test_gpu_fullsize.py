"""Parity at BASELINE.json's full sizes (256 clips = 2 560 examples per step; a 1-hour stream) through
size-independent properties — the CPU oracle cannot reach these sizes in test time:
   * shard / permutation invariance of the whole path (bit-identical)
   * every tensor-core kernel against its independent first-generation CUDA-core implementation on the device
   * spot checks of a few clips of the full batch against the CPU oracle
   * the 1-hour stream: example count, chunking invariance, accuracy-mode uint8 vs the oracle on a sample"""
import numpy as np
import pytest
import torch

from b200 import _lib, engine, sharding, stream, synth
from oracle import frontend_np, model_torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
B = 256


@pytest.fixture(scope="module")
def handles(vgg_sd, head_sd):
    v = engine.VggishHandle(vgg_sd, DEV)
    h = engine.MlaHandle(head_sd, (2, 1), 128, 600, 527, 10, DEV)
    yield v, h
    v.close()
    h.close()


@pytest.fixture(scope="module")
def batch():
    return synth.fast_clips(0, B).to(DEV)


def test_full_batch_shard_and_permutation_invariance(handles, batch):
    vgg, head = handles
    pipe = engine.Pipeline(vgg, head)
    whole, emb = pipe.forward(batch, want_embeddings=True)
    assert whole.shape == (B, 527) and torch.isfinite(whole).all() and emb.shape == (B * 10, 128)
    for world in (2, 8):
        parts = [pipe.forward(batch[slice(*sharding.shard_bounds(B, r, world))]).clone() for r in range(world)]
        assert torch.equal(torch.cat(parts), whole)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0)).to(DEV)
    assert torch.equal(pipe.forward(batch[perm].contiguous()), whole[perm])
    # the four signal families give four distinct behaviours, not a constant
    assert whole.std(dim=0).mean() > 1e-3


def test_full_batch_tensor_core_kernels_vs_cuda_core_kernels(handles, batch, vgg_sd):
    vgg, head = handles
    L = _lib.lib()
    ex = engine.examples_from_wave(batch)                                            # tcgen05 DFT
    ex_cc = engine.logmel_cudacore(batch)[:, :960].reshape(-1, 96, 64)               # fp32 CUDA cores
    assert (ex - ex_cc).abs().max().item() <= 2e-4
    w, b = vgg_sd["features.0.weight"].to(DEV).contiguous(), vgg_sd["features.0.bias"].to(DEV)
    o_tc = torch.empty(B * 10, 48, 32, 64, device=DEV, dtype=torch.bfloat16)
    o_cc = torch.empty_like(o_tc)
    engine.check(L.vmb_conv1_relu_pool(ex.data_ptr(), w.data_ptr(), b.data_ptr(), o_tc.data_ptr(), B * 10,
                                       engine.stream_ptr()), "conv1 tc")
    engine.check(L.vmb_conv1_relu_pool_cudacore(ex.data_ptr(), w.data_ptr(), b.data_ptr(), o_cc.data_ptr(), B * 10,
                                                engine.stream_ptr()), "conv1 cc")
    d = (o_tc.float() - o_cc.float()).abs().max().item() / o_cc.float().abs().max().item()
    assert d <= 8e-3                                                                  # one bf16 ulp of the max
    emb = vgg.forward(ex).reshape(B, 10, 128)
    s_tc = head.forward(emb)
    s_cc = head.forward(emb, fp32_crosscheck=True)
    assert (s_tc - s_cc).abs().max().item() <= 2e-5


def test_full_batch_spot_checks_vs_oracle(handles, vgg_sd, head_sd):
    vgg, head = handles
    pipe = engine.Pipeline(vgg, head)
    waves = synth.fast_clips(0, B)
    got = pipe.forward(waves.to(DEV)).cpu()
    idx = [0, 85, 170, 255]
    ex = np.concatenate([frontend_np.waveform_to_examples(waves[i].numpy().astype(np.float64)) for i in idx])
    with torch.no_grad():
        e = model_torch.vgg_forward(vgg_sd, torch.from_numpy(ex).float()[:, None])
        want = model_torch.mla_forward(head_sd, e.reshape(len(idx), 10, 128), (2, 1))
    assert (got[idx] - want).abs().max().item() <= 2e-2


def test_one_hour_stream(vgg_sd):
    n = 3600 * 16000
    g = torch.Generator().manual_seed(5)
    t = torch.arange(n, dtype=torch.float32) / 16000.0
    wave = (0.2 * torch.sin(2 * np.pi * 220.0 * t) + 0.05 * torch.randn(n, generator=g)).pin_memory()
    eig, means = synth.pca_params(1)
    vgg = engine.VggishHandle(vgg_sd, DEV, precision="split")
    emb, q = stream.embed_stream(vgg, wave, eig, means, examples_per_chunk=2048)
    assert emb.shape == (3749, 128) and q.shape == (3749, 128) and q.dtype == torch.uint8       # F10: 3 749, not 3 750
    emb2, q2 = stream.embed_stream(vgg, wave, eig, means, examples_per_chunk=500)                # other chunking
    assert torch.equal(q, q2) and torch.equal(emb, emb2)
    parts = [stream.embed_stream(vgg, wave, eig, means, 1024, rank=r, world=8)[1] for r in range(8)]   # 8-way shard
    assert torch.equal(torch.cat(parts), q)
    m = 64                                                                                        # oracle on the last minute
    s0, s1 = sharding.stream_sample_range(3749 - m, 3749)
    ex = frontend_np.waveform_to_examples(wave[s0:s1].numpy().astype(np.float64)).astype(np.float32)
    with torch.no_grad():
        ref = model_torch.postprocess(eig, means, model_torch.vgg_forward(vgg_sd, torch.from_numpy(ex)[:, None]))
    d = (q[-m:].cpu().float() - ref).abs()
    print("1-hour stream, accuracy mode, last 64 examples: uint8 LSB histogram", np.bincount(d.numpy().astype(np.int64).ravel()).tolist())
    assert d.max() <= 1 and (d > 0).float().mean() <= 0.02
    vgg.close()
    # the default (fp16) body on the same stream: every value within +-1 LSB of the reference too (north_star's bar),
    # with more values sitting on the other side of a quantisation boundary than in the split mode
    v16 = engine.VggishHandle(vgg_sd, DEV, precision="fp16")
    emb16, q16 = stream.embed_stream(v16, wave, eig, means, examples_per_chunk=2048)
    v16.check_saturation()
    d16 = (q16[-m:].cpu().float() - ref).abs()
    dq = (q16.int() - q.int()).abs()
    print("1-hour stream, fp16 mode, last 64 examples: uint8 LSB histogram vs the oracle",
          np.bincount(d16.numpy().astype(np.int64).ravel()).tolist(), "; all 3 749 examples vs the split mode",
          torch.bincount(dq.flatten()).tolist())
    assert d16.max() <= 1 and (d16 > 0).float().mean() <= 0.08 and dq.max() <= 1
    v16.close()


def test_sanitize_case_runs_clean():
    """tools/sanitize_case.py touches every kernel family on tiny inputs; it is what would run under
    `compute-sanitizer --tool memcheck` (the tool is refused on the pool's boxes).  Here it must at least exit 0 in a
    fresh process, and under the sanitizer when the box allows it."""
    import os
    import shutil
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = os.path.join(root, "tools", "sanitize_case.py")
    r = subprocess.run([sys.executable, script], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "sanitize case ok" in r.stdout
    cs = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if os.path.exists(cs) and os.environ.get("VMB_RUN_SANITIZER") == "1":
        r = subprocess.run([cs, "--tool", "memcheck", "--error-exitcode", "3", sys.executable, script],
                           capture_output=True, text=True, timeout=3000)
        assert r.returncode == 0, r.stdout[-3000:]
