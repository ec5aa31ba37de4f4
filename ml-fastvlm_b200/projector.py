"""mm_projector -- drop-in for `build_vision_projector` (llava/model/multimodal_projector/builder.py:17-35).

Module structure and state-dict keys are the reference's (`nn.Sequential` of Linear / GELU / Linear ->
`0.weight, 0.bias, 2.weight, 2.bias`; bare `nn.Linear` for "linear"); `forward` runs the tcgen05 GEMM with
the bias / GELU epilogue instead of ATen (GELU = erf form evaluated as 0.5x(1+tanh(x P(x^2))), |error| <= 3e-5,
see csrc/ptx.cuh gelu_erf).
"""
import re
from collections import OrderedDict

import torch
import torch.nn as nn

from .lib import FvhdError

_util_engines = {}


def _gemm_engine(device):
    """Plan-less handle for the stand-alone GEMM entry, one per CUDA device: the library's per-device setup
    (max dynamic smem attribute, SM count) is done by whichever device is current at the first call of a handle."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    eng = _util_engines.get(key)
    if eng is None:
        from .engine import Engine
        eng = _util_engines[key] = Engine(64, 0, 2, 1)
    return eng


class _PackedMixin:
    def _init_pack(self):
        self._version_counter = 0
        self._packed = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())

    def _invalidate(self):
        self._version_counter += 1
        self._packed = None

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._invalidate()
        return out

    def _layers(self):
        raise NotImplementedError

    def packed_state_dict(self):
        """Keys `0.weight/0.bias[/2.weight/2.bias]` regardless of projector flavour."""
        sd = OrderedDict()
        for i, lin in enumerate(self._layers()):
            sd[f"{2 * i}.weight"] = lin.weight.detach()
            sd[f"{2 * i}.bias"] = lin.bias.detach()
        return sd

    def _device_pack(self, device):
        if self._packed is None or self._packed[0] != device:
            ws = [(lin.weight.detach().to(device=device, dtype=torch.bfloat16).contiguous(),
                   lin.bias.detach().to(device=device, dtype=torch.float32).contiguous()) for lin in self._layers()]
            self._packed = (device, ws)
        return self._packed[1]

    def _run(self, x):
        if x.device.type != "cuda":
            raise FvhdError(f"mm_projector computes on CUDA (sm_100a) only; input is on {x.device}. No CPU fallback exists.")
        eng = _gemm_engine(x.device)
        shape = x.shape
        a = x.reshape(-1, shape[-1]).to(torch.bfloat16).contiguous()
        ws = self._device_pack(x.device)
        for i, (w, b) in enumerate(ws):
            a = eng.gemm(a, w, bias=b, act=1 if i + 1 < len(ws) else 0)
        return a.reshape(*shape[:-1], a.shape[-1]).to(x.dtype)


class FastVLMProjector(_PackedMixin, nn.Sequential):
    """`mlp{N}x_gelu` (builder.py:23-30)."""

    def __init__(self, mm_hidden_size, hidden_size, depth):
        mods = [nn.Linear(mm_hidden_size, hidden_size)]
        for _ in range(1, depth):
            mods.append(nn.GELU())
            mods.append(nn.Linear(hidden_size, hidden_size))
        super().__init__(*mods)
        if depth > 2:
            raise ValueError("libfastvithd_b200 implements mlp1x/mlp2x_gelu projectors (FastVLM uses mlp2x_gelu)")
        self._init_pack()

    def _layers(self):
        return [m for m in self if isinstance(m, nn.Linear)]

    def forward(self, x):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("libfastvithd_b200 is inference-only: training mm_projector needs the reference projector "
                                      "(call under torch.no_grad() or freeze it)")
        with torch.no_grad():
            return self._run(x)


class FastVLMLinearProjector(_PackedMixin, nn.Linear):
    """`linear` (builder.py:20-21)."""

    def __init__(self, mm_hidden_size, hidden_size):
        super().__init__(mm_hidden_size, hidden_size)
        self._init_pack()

    def _layers(self):
        return [self]

    def forward(self, x):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("libfastvithd_b200 is inference-only: training mm_projector needs the reference projector "
                                      "(call under torch.no_grad() or freeze it)")
        with torch.no_grad():
            return self._run(x)


class IdentityMap(nn.Module):
    """builder.py:5-14."""

    def forward(self, x, *args, **kwargs):
        return x

    @property
    def config(self):
        return {"mm_projector_type": "identity"}


def build_vision_projector(config, delay_load=False, **kwargs):
    projector_type = getattr(config, "mm_projector_type", "linear")
    if projector_type == "linear":
        return FastVLMLinearProjector(config.mm_hidden_size, config.hidden_size)
    m = re.match(r"^mlp(\d+)x_gelu$", projector_type)
    if m:
        return FastVLMProjector(config.mm_hidden_size, config.hidden_size, int(m.group(1)))
    if projector_type == "identity":
        return IdentityMap()
    raise ValueError(f"Unknown projector type: {projector_type}")
