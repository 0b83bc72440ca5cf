"""Bring-up helper: compare internal gradient buffers of the trainer with autograd intermediates of the oracle."""
import ctypes as C
import os
os.environ["VMB_TRAIN_DEBUG"] = "1"
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-classification-using-a-deep-cnn-combined-with-multi-level-attention_b200"))
sys.path.insert(0, ROOT)
from b200 import _lib, synth, training  # noqa: E402
from oracle import model_torch  # noqa: E402

DEV = torch.device("cuda:0")
K, conf, batch = 10, (2,), 32
tr = training.HeadTrainer(conf, 128, 600, K, 10, 128, DEV, dropout_p=0.0)
sd = synth.mla_state_dict(conf, 128, 600, K, 10, seed=7)
tr.load_state_dict(sd)
g = torch.Generator().manual_seed(batch)
x = torch.randn(batch, 10, 128, generator=g)
labels = torch.randint(0, K, (batch,), generator=g)
tr.forward_backward(x, labels)
torch.cuda.synchronize()

L = _lib.lib()
fn = L.vmb_mla_trainer_debug_buffer
fn.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong)]


def buf(name, l=0, j=0, rows=batch * 10, cols=600):
    p, ld = C.c_void_p(), C.c_longlong()
    assert fn(tr._h, name.encode(), l, j, C.byref(p), C.byref(ld)) == 0
    t2 = torch.empty(rows * ld.value, device=DEV)
    rt = C.CDLL("libcudart.so.12")
    rt.cudaMemcpy(C.c_void_p(t2.data_ptr()), p, C.c_size_t(rows * ld.value * 4), 3)
    return t2.view(rows, ld.value)[:, :cols].cpu()


# oracle intermediates with autograd
p = {k: v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v for k, v in sd.items()}
h0 = model_torch._bn_time(p, "embedded_mappings.0.norm0", x, True)
u0 = F.linear(h0, p["embedded_mappings.0.fc.0.weight"], p["embedded_mappings.0.fc.0.bias"]); u0.retain_grad()
a0 = F.relu(model_torch._bn_time(p, "embedded_mappings.0.norms.0", u0, True)); a0.retain_grad()
u1 = F.linear(a0, p["embedded_mappings.0.fc.1.weight"], p["embedded_mappings.0.fc.1.bias"]); u1.retain_grad()
v1 = model_torch._bn_time(p, "embedded_mappings.0.norms.1", u1, True); e0 = F.relu(v1); e0.retain_grad()
z = F.linear(e0, p["attention_modules.0.fcv.weight"], p["attention_modules.0.fcv.bias"]); z.retain_grad()
att = F.softmax(model_torch._bn_time(p, "attention_modules.0.normv", z, True), dim=2)
cla = torch.sigmoid(model_torch._bn_time(p, "attention_modules.0.normf", z, True))
y = torch.sum(cla * (att / att.sum(dim=1)[:, None, :]), dim=1); y.retain_grad()
o = F.linear(y, p["fc.weight"], p["fc.bias"]); o.retain_grad()
out = torch.sigmoid(F.batch_norm(o, None, None, p["norm.weight"], p["norm.bias"], True, 0.1, 1e-5))
loss = F.cross_entropy(out, labels)
loss.backward()


def cmp(name, got, ref):
    ref = ref.reshape(got.shape)
    print(f"{name:8s} max|ref| {ref.abs().max():.3e}  max-abs-err {(got - ref).abs().max():.3e}  rel {(got - ref).abs().max() / ref.abs().max():.2e}")


cmp("U00", buf("U", 0, 0), u0.detach())
cmp("U01", buf("U", 0, 1), u1.detach())
cmp("E0", buf("E", 0), e0.detach())
cmp("Z", buf("Z", 0, cols=K), z.detach())
cmp("Y", buf("Y", rows=batch, cols=K), y.detach())
cmp("dO", buf("dO", rows=batch, cols=K), o.grad)
cmp("dY", buf("dY", rows=batch, cols=K), y.grad)
# in conf (2,): dA = d E0 (attention), then dB = d a0, then dA = d n0
cmp("GF=dE0", buf("GF"), e0.grad)
cmp("GV=dU1", buf("GV"), u1.grad)
cmp("dB=da0", buf("dB"), a0.grad)
cmp("dA(end)", buf("dA", cols=128), h0.grad if h0.grad is not None else torch.zeros(batch, 10, 128))

got = buf("GV"); ref = u1.grad.reshape(got.shape)
err = (got - ref).abs()
bad = (err > 2e-6).nonzero()
print("bad count", bad.shape[0], "of", err.numel())
print("bad rows histogram (r % 10):", torch.bincount(bad[:, 0] % 10, minlength=10).tolist())
print("bad rows (first 20):", sorted(set(bad[:, 0].tolist()))[:20])
print("bad cols (first 20):", sorted(set(bad[:, 1].tolist()))[:20], "max col", bad[:, 1].max().item())
for r, c in bad[:8].tolist():
    print(r, c, got[r, c].item(), ref[r, c].item(), "e0.grad", e0.grad.reshape(got.shape)[r, c].item(), "v1", v1.detach().reshape(got.shape)[r, c].item())
