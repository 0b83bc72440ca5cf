"""Tiny end-to-end exercise of every kernel family, meant to be run under `compute-sanitizer --tool memcheck`
(after the same command has exited 0 without it): 2-clip pipeline in the three precisions, chunked stream +
postprocessor, ill-conditioned frames (float64 log-mel kernel), stft_magnitude, the standalone head levels, fp32
cross-check kernels, one head training step.  tests/test_gpu_fullsize.py runs it as a plain subprocess on every GPU
test run (compute-sanitizer itself is refused on the pool's boxes)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-classification-using-a-deep-cnn-combined-with-multi-level-attention_b200"))
from b200 import _lib, engine, stream, synth, training  # noqa: E402

dev = torch.device("cuda:0")
vsd = synth.vggish_state_dict(0)
msd = synth.mla_state_dict((2, 1), 128, 600, 527, 10, seed=2)
waves = torch.from_numpy(synth.make_clips(0, 2)).to(dev)
head = engine.MlaHandle(msd, (2, 1), 128, 600, 527, 10, dev)
for prec in ("fp16", "bf16", "split"):
    vgg = engine.VggishHandle(vsd, dev, precision=prec)
    pipe = engine.Pipeline(vgg, head)
    s = pipe.forward(waves)
    h = pipe.forward_host(waves.cpu().pin_memory(), clips_per_batch=1)
    assert torch.equal(h, s.cpu()) and torch.isfinite(s).all()
    eig, means = synth.pca_params(1)
    emb, q = stream.embed_stream(vgg, waves[0, :15360 * 4 + 15600].contiguous(), eig, means, examples_per_chunk=2)
    assert q.shape == (5, 128)
    vgg.close()
engine.logmel_cudacore(waves[:1, :32000].contiguous())
tone = torch.sin(torch.arange(20000, device=dev) * 0.4) * 0.5
assert torch.isfinite(engine.logmel(tone)).all() and torch.isfinite(engine.logmel(tone[1:])).all()   # exact-kernel path
assert engine.stft_magnitude(tone.double()).shape == (123, 257)
assert head.attention(0, head.embedded_mapping(0, torch.randn(3, 10, 128, device=dev).abs())).shape == (3, 527)
head.forward(torch.randn(3, 10, 128, device=dev).abs(), fp32_crosscheck=True)
tr = training.HeadTrainer((2, 1), 128, 600, 527, 10, 8, dev, dropout_p=0.4)
tr.load_state_dict(msd)
loss = tr.step(torch.randn(6, 10, 128), torch.randint(0, 527, (6,)))
torch.cuda.synchronize()
print("sanitize case ok: launches", _lib.lib().vmb_launch_count(), "loss", float(loss))
