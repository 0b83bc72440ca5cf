"""Head training on the B200: the multi-level attention head's train-mode forward,
backward and Adam update run in the library (csrc/mla_train.cu); across GPUs the only communication is ONE NCCL
all-reduce of the flat gradient bucket per step (data parallel, BatchNorm statistics stay local to each rank, which
is what stock DistributedDataParallel computes — SURVEY §7 H6).

Two entry levels:
  * `HeadTrainer` — flat fp32 parameter / gradient / Adam-moment buffers, `step(x, labels)` = one library call for
    forward+backward, one `dist.all_reduce`, one Adam kernel.  Used by bench_train.py and the tests.
  * `head_train_forward(module, x)` — autograd bridge for the reference-named `model.MultiLevelAttention` in
    train mode, so the reference's own loop (`criterion(outputs, labels); loss.backward(); optimizer.step()`,
    train.py:130-138) keeps working on top of the library kernels.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import B200Error, check, ptr, stream_ptr


def train_layout(model_conf: Sequence[int], emb_in: int, hidden: int, n_classes: int, t_steps: int):
    """(params, running): lists of (state_dict key, shape, offset) in the library's flat order — the reference
    module's named_parameters() order with `fcf` left out (it never receives a gradient)."""
    params: List[Tuple[str, Tuple[int, ...], int]] = []
    running: List[Tuple[str, Tuple[int, ...], int]] = []
    po = ro = 0

    def add_p(key, shape):
        nonlocal po
        n = 1
        for s in shape:
            n *= s
        params.append((key, tuple(shape), po))
        po += n

    def add_bn(prefix, n):
        nonlocal ro
        add_p(prefix + ".weight", (n,))
        add_p(prefix + ".bias", (n,))
        running.append((prefix + ".running_mean", (n,), ro))
        running.append((prefix + ".running_var", (n,), ro + n))
        ro += 2 * n

    for lvl, n_fc in enumerate(model_conf):
        p = f"embedded_mappings.{lvl}"
        add_bn(p + ".norm0", t_steps)
        for j in range(n_fc):
            fin = emb_in if (lvl == 0 and j == 0) else hidden
            add_p(f"{p}.fc.{j}.weight", (hidden, fin))
            add_p(f"{p}.fc.{j}.bias", (hidden,))
        for j in range(n_fc):
            add_bn(f"{p}.norms.{j}", t_steps)
    for lvl in range(len(model_conf)):
        p = f"attention_modules.{lvl}"
        add_p(p + ".fcv.weight", (n_classes, hidden))
        add_p(p + ".fcv.bias", (n_classes,))
        add_bn(p + ".normv", t_steps)
        add_bn(p + ".normf", t_steps)
    add_p("fc.weight", (n_classes, len(model_conf) * n_classes))
    add_p("fc.bias", (n_classes,))
    add_bn("norm", n_classes)
    return params, running, po, ro


class HeadTrainer:
    """Owns the library trainer handle and the flat buffers of one data-parallel rank."""

    def __init__(self, model_conf: Sequence[int], emb_in: int, hidden: int, n_classes: int, t_steps: int,
                 max_batch: int, device: torch.device, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 dropout_p: float = 0.4, seed: int = 0, process_group=None):
        from . import engine
        engine.require_b200(device)
        self.device, self.model_conf = device, tuple(int(v) for v in model_conf)
        self.dims = (emb_in, hidden, n_classes, t_steps)
        self.lr, self.betas, self.eps, self.dropout_p, self.seed = lr, betas, eps, dropout_p, seed
        self.group = process_group
        self.p_layout, self.r_layout, n_p, n_r = train_layout(self.model_conf, emb_in, hidden, n_classes, t_steps)
        conf = (C.c_int * len(self.model_conf))(*self.model_conf)
        n_run = C.c_longlong(0)
        expect = _lib.lib().vmb_mla_train_param_count(len(self.model_conf), conf, emb_in, hidden, n_classes, t_steps,
                                                      C.byref(n_run))
        if expect != n_p or n_run.value != n_r:
            raise B200Error(f"flat layout mismatch: python {n_p}/{n_r} vs library {expect}/{n_run.value}")
        self.n_params = n_p
        with torch.cuda.device(device):
            self.params = torch.zeros(n_p, device=device)
            self.grads = torch.zeros(n_p, device=device)
            self.exp_avg = torch.zeros(n_p, device=device)
            self.exp_avg_sq = torch.zeros(n_p, device=device)
            self.running = torch.zeros(n_r, device=device)
            self.loss = torch.zeros(1, device=device)
            self._h = C.c_void_p()
            check(_lib.lib().vmb_mla_trainer_create(C.byref(self._h), len(self.model_conf), conf, emb_in, hidden,
                                                    n_classes, t_steps, int(max_batch), stream_ptr()),
                  "vmb_mla_trainer_create")
        self.max_batch = int(max_batch)
        self.step_count = 0
        self.num_batches_tracked = 0
        self.generation = 0      # bumped by every train-mode forward through the autograd bridge (one step in flight)
        self.fcf: Dict[str, torch.Tensor] = {}      # carried through state_dict round trips untouched

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value and _lib is not None:     # _lib is None at interpreter exit
            _lib.lib().vmb_mla_trainer_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    # -- reference-format state_dict in / out
    def load_state_dict(self, sd: dict) -> None:
        for key, shape, off in self.p_layout:
            self.view(self.params, key).copy_(sd[key].detach().to(self.device, torch.float32))
        for key, shape, off in self.r_layout:
            self.running[off:off + shape[0]].copy_(sd[key].detach().to(self.device, torch.float32))
        self.fcf = {k: v.detach().clone() for k, v in sd.items() if ".fcf." in k}

    def view(self, flat: torch.Tensor, key: str) -> torch.Tensor:
        for k, shape, off in self.p_layout:
            if k == key:
                n = 1
                for s in shape:
                    n *= s
                return flat[off:off + n].view(shape)
        raise KeyError(key)

    def state_dict(self) -> dict:
        out = {}
        for key, shape, off in self.p_layout:
            out[key] = self.view(self.params, key).clone()
        for key, shape, off in self.r_layout:
            out[key] = self.running[off:off + shape[0]].clone()
            if key.endswith("running_var"):
                out[key[:-len("running_var")] + "num_batches_tracked"] = torch.tensor(self.num_batches_tracked)
        out.update(self.fcf)
        return out

    # -- one optimisation step
    def forward_backward(self, x: torch.Tensor, labels: torch.Tensor, want_scores: bool = False):
        """grads <- d(mean CE-on-sigmoid loss)/d(params) of THIS rank's batch; returns (loss tensor, scores|None)."""
        if x.dim() != 3 or x.shape[1] != self.dims[3] or x.shape[2] != self.dims[0]:
            raise ValueError(f"expected (B, {self.dims[3]}, {self.dims[0]}), got {tuple(x.shape)}")
        x = x.to(self.device, torch.float32).contiguous()
        labels = labels.to(self.device, torch.int64).contiguous()
        b = x.shape[0]
        scores = torch.empty(b, self.dims[2], device=self.device) if want_scores else None
        with torch.cuda.device(self.device):
            check(_lib.lib().vmb_mla_train_step(self._h, ptr(self.params), ptr(self.running), ptr(x), ptr(labels), b,
                                                float(self.dropout_p), int(self.seed + self.step_count) & (2 ** 63 - 1),
                                                ptr(self.grads), ptr(self.loss), ptr(scores), stream_ptr()),
                  "vmb_mla_train_step")
        self.num_batches_tracked += 1
        return self.loss, scores

    def all_reduce_grads(self) -> int:
        """One SUM all-reduce of the flat gradient bucket; returns the world size (the Adam kernel divides)."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return 1
        world = dist.get_world_size(self.group)
        if world > 1:
            dist.all_reduce(self.grads, op=dist.ReduceOp.SUM, group=self.group)
        return world

    def adam(self, world: int = 1) -> None:
        self.step_count += 1
        with torch.cuda.device(self.device):
            check(_lib.lib().vmb_adam_step(ptr(self.params), ptr(self.grads), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                                           self.n_params, float(self.lr), float(self.betas[0]), float(self.betas[1]),
                                           float(self.eps), 0.0, self.step_count, 1.0 / world, stream_ptr()),
                  "vmb_adam_step")

    def step(self, x: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        loss, _ = self.forward_backward(x, labels)
        self.adam(self.all_reduce_grads())
        return loss


# ---------------------------------------------------------------------------------------------- nn.Module bridge
class _HeadTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, *params):
        st = module._b200_train_state()
        tr: HeadTrainer = st["trainer"]
        x = x.detach().to(tr.device, torch.float32).contiguous()
        if x.shape[0] > tr.max_batch:
            raise B200Error(f"batch {x.shape[0]} > trainer capacity {tr.max_batch}")
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())        # follows torch's global RNG like nn.Dropout does
        scores = torch.empty(x.shape[0], tr.dims[2], device=tr.device)
        with torch.cuda.device(tr.device):
            check(_lib.lib().vmb_mla_train_forward(tr._h, ptr(tr.params), ptr(tr.running), ptr(x), x.shape[0],
                                                   float(tr.dropout_p), seed, ptr(scores), stream_ptr()),
                  "vmb_mla_train_forward")
        # the activations, BatchNorm statistics and dropout masks the backward pass needs live in the trainer handle,
        # which holds ONE step: remember which forward this graph belongs to
        tr.generation += 1
        ctx.module, ctx.seed, ctx.x, ctx.generation = module, seed, x, tr.generation
        ctx.param_key = tuple((p.data_ptr(), p._version) for p in params)
        return scores

    @staticmethod
    def backward(ctx, dscores):
        st = ctx.module._b200_train_state()
        tr: HeadTrainer = st["trainer"]
        if ctx.generation != tr.generation:
            raise B200Error("backward of a stale head forward: the library trainer keeps the activations of ONE train-mode "
                            "forward per module, and another train-mode forward ran before this backward (e.g. two "
                            "micro-batches summed into one loss); call backward() before the next forward")
        if ctx.param_key != tuple((p.data_ptr(), p._version) for _, p in ctx.module.named_parameters()):
            raise B200Error("the head's parameters changed between forward and backward")
        d = dscores.to(tr.device, torch.float32).contiguous()
        with torch.cuda.device(tr.device):
            check(_lib.lib().vmb_mla_train_backward(tr._h, ptr(tr.params), ptr(ctx.x), ptr(d), ctx.x.shape[0],
                                                    float(tr.dropout_p), ctx.seed, ptr(tr.grads), stream_ptr()),
                  "vmb_mla_train_backward")
        grads = []
        for name in st["param_names"]:
            grads.append(None if ".fcf." in name else tr.view(tr.grads, name).clone())
        return (None, None, *grads)


def head_train_forward(module, x: torch.Tensor) -> torch.Tensor:
    """Train-mode forward of the reference-named head through the library (autograd-aware)."""
    if isinstance(x, torch.Tensor) and x.requires_grad and torch.is_grad_enabled():
        raise NotImplementedError("the head's backward pass produces parameter gradients only: the gradient with respect "
                                  "to the embeddings is outside this path (the reference freezes the CNN, "
                                  "model.py:159-160); pass x.detach()")
    st = module._b200_train_state()
    st["sync_in"]()
    params = [p for _, p in module.named_parameters()]
    out = _HeadTrainFn.apply(module, x, *params)
    st["sync_out"]()
    return out
