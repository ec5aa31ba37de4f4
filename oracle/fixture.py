"""TEST INFRASTRUCTURE ONLY -- seeded, reference-keyed parity fixture for FastViTHD + mm_projector.

Why not the reference's default init: `layer_scale` starts at 1e-5 (mci.py:756,1058,1132) and
BatchNorm at identity, which makes every ConvFFN / attention branch (~98 % of the MACs)
numerically invisible (SURVEY.md finding 3).  This generator therefore writes *every* tensor of
the reference state-dict itself, with fan-in scaled Gaussians, O(1) layer scales and random BN
statistics, so that each unit on the path moves the output.

The key names and shapes are those of the unmodified reference model
(`MobileCLIPVisionTower.state_dict()`, checked by oracle/gen_golden.py against the live
reference, and by tests/test_oracle.py against the key list stored in tests/golden/).
Values come from numpy's PCG64 so they are reproducible on the GPU box, where
/root/reference does not exist.

Architecture constants follow `fastvithd()` (mci.py:1454-1478).
"""
from collections import OrderedDict

import numpy as np
import torch

LAYERS = (2, 12, 24, 4, 2)                 # mci.py:1457
EMBED_DIMS = (96, 192, 384, 768, 1536)     # mci.py:1458
TOKEN_MIXERS = ("repmixer", "repmixer", "repmixer", "attention", "attention")  # mci.py:1462
MLP_RATIO = 4                              # mci.py:1459
HEAD_DIM = 32                              # mci.py:636
CLS_RATIO = 2                              # mci.py:1329 (conv_exp: 1536 -> 3072)
SE_RD = 0.0625                             # mci.py:49
PROJECTION_DIM = 768                       # mobileclip_l.json:2 (head.proj, off-path)
TOWER_PREFIX = "vision_tower.model."       # MobileCLIPVisionTower.vision_tower (MCi) .model (FastViT)


def network_layout():
    """Yield (network_index, kind, stage) in reference order (mci.py:1357-1398).

    kind in {"stage", "down", "cpe"}.  RepCPE precedes stages 3 and 4 (pos_embs, mci.py:1461).
    """
    out = []
    idx = 0
    for i in range(len(LAYERS)):
        if i >= 3:
            out.append((idx, "cpe", i))
            idx += 1
        out.append((idx, "stage", i))
        idx += 1
        if i < len(LAYERS) - 1:
            out.append((idx, "down", i))
            idx += 1
    return out


class _Gen:
    def __init__(self, seed):
        self.rng = np.random.Generator(np.random.PCG64(seed))

    def normal(self, shape, std, mean=0.0):
        return torch.from_numpy((self.rng.standard_normal(shape, dtype=np.float32) * std + mean).astype(np.float32))

    def uniform(self, shape, lo, hi):
        return torch.from_numpy(self.rng.uniform(lo, hi, size=shape).astype(np.float32))


def _conv(g, sd, name, cout, cin_per_group, k, gain=1.0, bias_std=0.05, identity=False):
    fan_in = cin_per_group * k * k
    w = g.normal((cout, cin_per_group, k, k), gain / np.sqrt(fan_in))
    if identity:  # reparameterised skip: centre tap carries the identity (mci.py:808-859, 1000-1039)
        w[:, 0, k // 2, k // 2] += 1.0
    sd[name + ".weight"] = w
    sd[name + ".bias"] = g.normal((cout,), bias_std)


def _convffn(g, sd, p, c, gain=1.0):
    # ConvFFN (mci.py:862-927): dw7x7 (no bias) -> BN -> fc1 -> GELU -> fc2
    sd[p + ".conv.conv.weight"] = g.normal((c, 1, 7, 7), 1.0 / 7.0)
    sd[p + ".conv.bn.weight"] = g.uniform((c,), 0.5, 1.5)
    sd[p + ".conv.bn.bias"] = g.normal((c,), 0.1)
    sd[p + ".conv.bn.running_mean"] = g.normal((c,), 0.1)
    sd[p + ".conv.bn.running_var"] = g.uniform((c,), 0.5, 1.5)
    sd[p + ".conv.bn.num_batches_tracked"] = torch.zeros((), dtype=torch.int64)
    _conv(g, sd, p + ".fc1", MLP_RATIO * c, c, 1, gain=gain)
    _conv(g, sd, p + ".fc2", c, MLP_RATIO * c, 1, gain=gain)


def tower_state_dict(seed=123, variant="a"):
    """Reference-keyed fp32 state-dict of MobileCLIPVisionTower (629 entries).

    variant "a" (default, all goldens): layer scales U(0.15, 0.8) with fan-in gain 1 on the ConvFFN / attention weights.
    variant "b": the layer-scale range SURVEY.md 8(d) prescribes, U(0.5, 1.5) for every layer_scale*, with the branch weights
    scaled down (gain 0.4) -- with gain-1 weights that range makes the activations grow to std 3e4 and the reference's own
    bf16 run then sits 0.11 from its fp32 run (measured), i.e. it cannot resolve a kernel error.  Variant b keeps the
    activations O(1) at the survey's scales; it is a second, independent parity case (tests/test_gpu_r2.py)."""
    g = _Gen(seed)
    vb = variant == "b"
    ls = (lambda lo, hi: (0.5, 1.5)) if vb else (lambda lo, hi: (lo, hi))
    bg = 0.4 if vb else 1.0
    sd = OrderedDict()
    P = TOWER_PREFIX
    c0 = EMBED_DIMS[0]
    # stem (mci.py:553-603)
    _conv(g, sd, P + "patch_embed.0.reparam_conv", c0, 3, 3, gain=1.6)
    _conv(g, sd, P + "patch_embed.1.reparam_conv", c0, 1, 3, gain=1.6)
    _conv(g, sd, P + "patch_embed.2.reparam_conv", c0, c0, 1, gain=1.6)
    for idx, kind, i in network_layout():
        c = EMBED_DIMS[i]
        n = P + f"network.{idx}"
        if kind == "cpe":       # RepCPE (mci.py:971-980): dw7x7 with folded identity
            _conv(g, sd, n + ".reparam_conv", c, 1, 7, gain=0.5, identity=True)
        elif kind == "down":    # PatchEmbed (mci.py:688-741): dw7x7 s2 (x2 channels) + 1x1
            co = EMBED_DIMS[i + 1]
            _conv(g, sd, n + ".proj.0.lkb_reparam", co, 1, 7, gain=1.6)
            _conv(g, sd, n + ".proj.1.reparam_conv", co, co, 1, gain=1.6)
        else:
            for b in range(LAYERS[i]):
                p = n + f".{b}"
                if TOKEN_MIXERS[i] == "repmixer":   # RepMixerBlock (mci.py:1042-1113)
                    sd[p + ".layer_scale"] = g.uniform((c, 1, 1), *ls(0.15, 0.45))
                    _conv(g, sd, p + ".token_mixer.reparam_conv", c, 1, 3, gain=0.3, identity=True)
                    _convffn(g, sd, p + ".convffn", c, bg)
                else:                               # AttentionBlock (mci.py:1116-1192)
                    sd[p + ".layer_scale_1"] = g.uniform((c, 1, 1), *ls(0.3, 0.8))
                    sd[p + ".layer_scale_2"] = g.uniform((c, 1, 1), *ls(0.2, 0.6))
                    sd[p + ".norm.weight"] = g.uniform((c,), 0.5, 1.5)
                    sd[p + ".norm.bias"] = g.normal((c,), 0.1)
                    sd[p + ".token_mixer.qkv.weight"] = g.normal((3 * c, c), 1.3 / np.sqrt(c))
                    sd[p + ".token_mixer.proj.weight"] = g.normal((c, c), bg / np.sqrt(c))
                    sd[p + ".token_mixer.proj.bias"] = g.normal((c,), 0.05)
                    _convffn(g, sd, p + ".convffn", c, bg)
    # conv_exp + SE (mci.py:1401-1411, 42-81)
    ce = EMBED_DIMS[-1] * CLS_RATIO
    rd = int(ce * SE_RD)
    _conv(g, sd, P + "conv_exp.se.reduce", rd, ce, 1, gain=2.0)
    _conv(g, sd, P + "conv_exp.se.expand", ce, rd, 1, gain=2.0)
    _conv(g, sd, P + "conv_exp.reparam_conv", ce, 1, 3, gain=1.2)
    # GlobalPool2D head (mci.py:1272-1302) -- off-path, present in the state-dict
    sd[P + "head.proj"] = g.normal((ce, PROJECTION_DIM), ce ** -0.5)
    return sd


def projector_state_dict(hidden, mm_hidden=EMBED_DIMS[-1] * CLS_RATIO, depth=2, seed=321):
    """`mlp{depth}x_gelu` projector (multimodal_projector/builder.py:23-30): keys `0`, `2`, ..."""
    g = _Gen(seed)
    sd = OrderedDict()
    fin = mm_hidden
    for d in range(depth):
        sd[f"{2 * d}.weight"] = g.normal((hidden, fin), 1.4 / np.sqrt(fin))
        sd[f"{2 * d}.bias"] = g.normal((hidden,), 0.05)
        fin = hidden
    return sd


def default_init_like_reference(sd):
    """Return a copy with layer scales at 1e-5 and BN at identity -- the reference's *default* init
    for those tensors (mci.py:756,1058,1132).  Used only by the fixture-sensitivity test."""
    out = OrderedDict((k, v.clone()) for k, v in sd.items())
    for k in out:
        if "layer_scale" in k:
            out[k].fill_(1e-5)
        elif k.endswith("bn.weight") or k.endswith("bn.running_var"):
            out[k].fill_(1.0)
        elif k.endswith("bn.bias") or k.endswith("bn.running_mean"):
            out[k].zero_()
    return out


def synthetic_images(batch, res, seed=1):
    """`torch.rand(B,3,R,R)` in [0,1) -- processor output range (mobileclip_encoder.py:45-49,
    model_export/export_vision_encoder.py:72), numpy-seeded so it travels."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy(rng.random((batch, 3, res, res), dtype=np.float32))
