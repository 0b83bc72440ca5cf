"""Drop-in for the reference's model.py (vggish branch): `Ensemble`, `Input`, `CNN`, `CnnFlatten`,
`EmbeddedMapping`, `AttentionModule`, `MultiLevelAttention`, `set_requires_grad` with the reference's constructor
signatures, state_dict keys and quirks (model.py:12-280), computing on the B200 library.

As in the reference, `T`, `H`, `DR`, `K` are module globals star-imported from params.py and read when a module
is CONSTRUCTED: set `model.K = 527` before building the head for AudioSet.

The nn.Linear / nn.BatchNorm1d children own the parameters (and running statistics) under the reference's key
names; their forward methods are never used.  Eval mode runs the tensor-core head (csrc/mla_tc.cu); training mode
runs the library's forward/backward kernels through a torch.autograd.Function (csrc/mla_train.cu) so that
`loss.backward()` and `torch.optim.Adam` from the reference's train loop keep working.  The ResNet50 branch is
third-party torchvision code outside this path: asking for it raises.
"""
from typing import Dict, List, Union

import torch
from torch import nn

from b200 import engine as _engine
from b200._lib import B200Error
from params import *  # noqa: F401,F403  (T, H, DR, K, M_VGGISH, ... bound here like in the reference)
from torchvggish.vggish import B200HandleMixin, VGGish, VGGishFeatures

_RESNET_MSG = ("the ResNet50 branch is third-party torchvision code outside the B200 waveform->VGGish->MLA path "
               "(SURVEY.md §2)")


class Ensemble(nn.Module):
    """CNN feature extractor + multi-level attention head (model.py:12-62).

    cnn_conf keys: cnn_type ('vggish'), num_classes, use_pretrained, just_bottlenecks, cnn_trainable,
    first_cnn_layer_trainable, in_channels — passed through to CNN like the reference does."""

    def __init__(self, input_conf: str, cnn_conf: Dict[str, Union[str, int]], model_conf: List[int], device):
        super().__init__()
        self.cnn_type = cnn_conf["cnn_type"]
        self.just_bottlenecks = cnn_conf["just_bottlenecks"]
        self.num_classes = cnn_conf["num_classes"]
        if self.cnn_type == "vggish":
            self.emb_input_size = M_VGGISH_JB if self.just_bottlenecks else M_VGGISH
        elif self.cnn_type == "resnet":
            raise NotImplementedError(_RESNET_MSG)
        else:
            raise Exception("CNN type is not valid.")
        self.input = Input(input_conf=input_conf, cnn_type=self.cnn_type, device=device)
        self.mla = MultiLevelAttention(model_conf, self.emb_input_size)
        self.cnn = CNN(**cnn_conf)

    def forward(self, x):
        """x (B, T, 1, 96, 64) examples -> (B, K) scores."""
        features = self.cnn(self.input(x))
        return self.mla(features.reshape(-1, T, self.emb_input_size))

    def forward_waveform(self, wave):
        """Fast path that also replaces the front end: wave (B, 160000) fp32 CUDA 16 kHz -> (B, K) scores with the
        log-mel, VGGish and head kernels enqueued back to back inside one library call (vmb_pipeline_forward)."""
        if self.just_bottlenecks or self.training:
            raise NotImplementedError("forward_waveform covers the eval-mode 128-d embedding configuration")
        pipe = _engine.Pipeline(self.cnn.cnn_model._b200_handle(), self.mla._b200_handle())
        return pipe.forward(wave)


class Input(nn.Module):
    """Reshapes (B, T, 1, 96, 64) to the CNN's (B*T, 1, 96, 64) (model.py:66-103).  No arithmetic."""

    def __init__(self, input_conf, cnn_type, device):
        super().__init__()
        self.conf, self.device, self.cnn_type = input_conf, device, cnn_type

    def forward(self, x):
        if self.cnn_type == "vggish":
            return x.reshape((-1, 1, S_VGGISH_SHAPE[0], S_VGGISH_SHAPE[1]))
        if self.cnn_type == "resnet":
            raise NotImplementedError(_RESNET_MSG)
        raise Exception("CNN type is not valid.")


class CNN(nn.Module):
    """Feature extractor (model.py:106-175), VGGish only.  The body is frozen unless cnn_trainable (training it
    is outside this path and raises at forward time)."""

    def __init__(self, cnn_type="vggish", num_classes=10, use_pretrained=True, just_bottlenecks=False,
                 cnn_trainable=False, first_cnn_layer_trainable=False, in_channels=3):
        super().__init__()
        if cnn_type == "resnet":
            raise NotImplementedError(_RESNET_MSG)
        if cnn_type != "vggish":
            raise Exception("Invalid CNN model name specified.")
        urls = {"vggish": "https://github.com/harritaylor/torchvggish/releases/download/v0.1/vggish-10086976.pth"}
        self.cnn_model = VGGish(urls=urls, pretrained=use_pretrained, preprocess=False, postprocess=False,
                                progress=True)
        self.just_bottlenecks = just_bottlenecks
        if not cnn_trainable:
            set_requires_grad(self.cnn_model, False)
        if just_bottlenecks:
            # the reference keeps only the first child (the conv stack) and re-wraps it with CnnFlatten in an
            # nn.Sequential (model.py:161-166): state_dict keys become cnn_model.0.N.*, the FC stack is gone
            self.cnn_model = VGGishFeatures(list(self.cnn_model.children())[0], CnnFlatten(cnn_type))

    def forward(self, x):
        return self.cnn_model(x)


class CnnFlatten(nn.Module):
    """(N, C, H, W) -> (N, H*W*C) in (h, w, c) order for VGGish (model.py:178-196).  Pure layout."""

    def __init__(self, cnn_type):
        super().__init__()
        self.cnn_type = cnn_type

    def forward(self, x):
        if self.cnn_type == "vggish":
            return x.permute(0, 2, 3, 1).contiguous().view(x.size(0), -1)
        if self.cnn_type == "resnet":
            return torch.flatten(x, 1)
        raise Exception("Invalid CNN model name specified.")


def _bn_identity(sd, prefix, n, dev):
    sd[prefix + ".weight"] = torch.ones(n, device=dev)
    sd[prefix + ".bias"] = torch.zeros(n, device=dev)
    sd[prefix + ".running_mean"] = torch.zeros(n, device=dev)
    sd[prefix + ".running_var"] = torch.ones(n, device=dev)


def _sub_module_handle(mod, build):
    """Library handle of a sub-module used on its own: a one-level head built around the module's parameters."""
    tensors = list(mod.parameters()) + list(mod.buffers())
    dev = tensors[0].device
    if dev.type != "cuda":
        raise B200Error("parameters are on %s: move the module to a CUDA device; this build has no CPU path" % dev)
    key = mod._tensor_key(tensors, dev)
    if mod._sub_handle is None or key != mod._sub_handle_key:
        if mod._sub_handle is not None:
            mod._sub_handle.close()
        mod._sub_handle = build(dev)
        mod._sub_handle_key = key
    return mod._sub_handle


class EmbeddedMapping(B200HandleMixin, nn.Module):
    """One embedding level: norm0, fc[j], norms[j], dropouts[j] (model.py:200-222).  Inside MultiLevelAttention the
    level is part of the head's kernel sequence; called on its own (eval mode) it runs the same kernels through
    vmb_mla_embedded_mapping."""

    def __init__(self, n_fc, is_first, emb_input_size):
        super().__init__()
        self.n_fc = n_fc
        self.norm0 = nn.BatchNorm1d(T)
        widths = [emb_input_size if is_first else H] + [H] * (n_fc - 1)
        self.fc = nn.ModuleList(nn.Linear(w, H) for w in widths)
        self.dropouts = nn.ModuleList(nn.Dropout(p=DR) for _ in range(n_fc))
        self.norms = nn.ModuleList(nn.BatchNorm1d(T) for _ in range(n_fc))
        self._t, self._h = T, H
        self._sub_handle = None
        self._sub_handle_key = None

    def _build(self, dev):
        # a one-level head around this level's parameters; the attention / output parameters are placeholders that
        # vmb_mla_embedded_mapping never touches
        sd = {"embedded_mappings.0." + k: v for k, v in self.state_dict().items()}
        sd["attention_modules.0.fcv.weight"] = torch.zeros(1, self._h, device=dev)
        sd["attention_modules.0.fcv.bias"] = torch.zeros(1, device=dev)
        _bn_identity(sd, "attention_modules.0.normv", self._t, dev)
        _bn_identity(sd, "attention_modules.0.normf", self._t, dev)
        sd["fc.weight"] = torch.zeros(1, 1, device=dev)
        sd["fc.bias"] = torch.zeros(1, device=dev)
        _bn_identity(sd, "norm", 1, dev)
        return _engine.MlaHandle(sd, [self.n_fc], self.fc[0].in_features, self._h, 1, self._t, dev)

    def forward(self, x):
        """(B, T, in) -> (B, T, H): norm0, then Dropout(ReLU(BN_T(Linear))) per fc (model.py:217-222), eval mode."""
        if self.training:
            raise NotImplementedError("train-mode EmbeddedMapping runs inside MultiLevelAttention.forward on this path "
                                      "(csrc/mla_train.cu); call the head, or .eval() to use the level on its own")
        h = _sub_module_handle(self, self._build)
        return h.embedded_mapping(0, x.detach().to(device=h.device, dtype=torch.float32))


class AttentionModule(B200HandleMixin, nn.Module):
    """One attention level: fcv, fcf (constructed but unused, F3), normv, normf (model.py:226-242); on its own (eval
    mode) it runs through vmb_mla_attention."""

    def __init__(self):
        super().__init__()
        self.fcv = nn.Linear(H, K)
        self.fcf = nn.Linear(H, K)
        self.normv = nn.BatchNorm1d(T)
        self.normf = nn.BatchNorm1d(T)
        self._t, self._h, self._k = T, H, K
        self._sub_handle = None
        self._sub_handle_key = None

    def _build(self, dev):
        sd = {"attention_modules.0." + k: v for k, v in self.state_dict().items()}
        _bn_identity(sd, "embedded_mappings.0.norm0", self._t, dev)          # placeholders: never evaluated
        sd["embedded_mappings.0.fc.0.weight"] = torch.zeros(self._h, self._h, device=dev)
        sd["embedded_mappings.0.fc.0.bias"] = torch.zeros(self._h, device=dev)
        _bn_identity(sd, "embedded_mappings.0.norms.0", self._t, dev)
        sd["fc.weight"] = torch.zeros(self._k, self._k, device=dev)
        sd["fc.bias"] = torch.zeros(self._k, device=dev)
        _bn_identity(sd, "norm", self._k, dev)
        return _engine.MlaHandle(sd, [1], self._h, self._h, self._k, self._t, dev)

    def forward(self, h):
        """(B, T, H) -> (B, K): att = softmax_K(normv(fcv h)), cla = sigmoid(normf(fcv h)), sum_t cla att / sum_t att
        (model.py:235-242, with its fcv-twice and softmax-over-classes quirks), eval mode."""
        if self.training:
            raise NotImplementedError("train-mode AttentionModule runs inside MultiLevelAttention.forward on this path "
                                      "(csrc/mla_train.cu); call the head, or .eval() to use the level on its own")
        hd = _sub_module_handle(self, self._build)
        return hd.attention(0, h.detach().to(device=hd.device, dtype=torch.float32))


class MultiLevelAttention(B200HandleMixin, nn.Module):
    """Multi-level attention head (model.py:246-269): (B, T, M) embeddings -> (B, K) sigmoid scores."""

    def __init__(self, model_conf, emb_input_size):
        super().__init__()
        self.model = model_conf
        self.emb_input_size = emb_input_size
        self.embedded_mappings = nn.ModuleList(
            EmbeddedMapping(n, is_first=(i == 0), emb_input_size=emb_input_size) for i, n in enumerate(model_conf))
        self.attention_modules = nn.ModuleList(AttentionModule() for _ in model_conf)
        self.fc = nn.Linear(len(model_conf) * K, K)
        self.norm = nn.BatchNorm1d(K)
        # construction-time values of the star-imported globals
        self._t, self._h, self._k, self._dr = T, H, K, DR
        self._handle = None
        self._handle_key = None
        self._train_state = None

    def _b200_handle(self):
        dev = self.fc.weight.device
        if dev.type != "cuda":
            raise B200Error("head parameters are on %s: move the module to a CUDA device; this build has no CPU "
                            "path" % dev)
        key = self._tensor_key(list(self.parameters()) + list(self.buffers()), dev)
        if self._handle is None or key != self._handle_key:
            if self._handle is not None:
                self._handle.close()
            self._handle = _engine.MlaHandle(self.state_dict(), list(self.model), self.emb_input_size, self._h, self._k,
                                             self._t, dev)
            self._handle_key = key
        return self._handle

    def _b200_train_state(self):
        """Lazily built bridge to the library trainer: flat parameter buffers mirrored from / to this module."""
        dev = self.fc.weight.device
        if dev.type != "cuda":
            raise B200Error("head parameters are on %s: move the module to a CUDA device; this build has no CPU "
                            "path" % dev)
        st = getattr(self, "_train_state", None)
        if st is not None and st["trainer"].device == dev:
            return st
        from b200 import training as _training
        tr = _training.HeadTrainer(list(self.model), self.emb_input_size, self._h, self._k, self._t, max_batch=4096,
                                   device=dev, dropout_p=self._dr)
        names = [n for n, _ in self.named_parameters()]
        pmap = dict(self.named_parameters())
        bmap = dict(self.named_buffers())

        def sync_in():        # module -> flat (parameters may have been changed by the optimiser or load_state_dict)
            for key, shape, off in tr.p_layout:
                tr.view(tr.params, key).copy_(pmap[key].detach())
            for key, shape, off in tr.r_layout:
                tr.running[off:off + shape[0]].copy_(bmap[key])

        def sync_out():       # flat -> module: BatchNorm running statistics updated by the forward pass
            for key, shape, off in tr.r_layout:
                bmap[key].copy_(tr.running[off:off + shape[0]])
                if key.endswith("running_var"):
                    bmap[key[:-len("running_var")] + "num_batches_tracked"].add_(1)

        self._train_state = dict(trainer=tr, param_names=names, sync_in=sync_in, sync_out=sync_out)
        return self._train_state

    def invalidate(self):
        st = self.__dict__.get("_train_state")
        if st is not None:
            st["trainer"].close()
        super().invalidate()

    def forward(self, x):
        if self.training:
            from b200 import training as _training
            return _training.head_train_forward(self, x)
        h = self._b200_handle()
        return h.forward(x.detach().to(device=h.device, dtype=torch.float32))


def set_requires_grad(model, value):
    """Sets requires_grad of every parameter of `model` (model.py:272-280)."""
    for param in model.parameters():
        param.requires_grad = value
