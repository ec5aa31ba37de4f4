"""Parameter layout of the reference FastViTHD tower / projector (names, shapes, buffers).

The tower module must accept released checkpoints unchanged, i.e. expose exactly the state-dict keys
of `MobileCLIPVisionTower` (llava/model/multimodal_encoder/mobileclip_encoder.py:13):
`vision_tower.model.<k>` with <k> from `FastViT.__init__` for `fastvithd()`
(mobileclip/mci.py:1353-1411, 1454-1478).  This module derives them from the architecture constants.
"""
from collections import OrderedDict

import torch

LAYERS = (2, 12, 24, 4, 2)                  # mci.py:1457
EMBED_DIMS = (96, 192, 384, 768, 1536)      # mci.py:1458
TOKEN_MIXERS = ("repmixer", "repmixer", "repmixer", "attention", "attention")   # mci.py:1462
MLP_RATIO = 4
EMBED_DIM = 3072                            # conv_exp output == mobileclip_l.json image_cfg.embed_dim
SE_RD = 192                                 # int(3072 * 0.0625), mci.py:49
PROJECTION_DIM = 768                        # mobileclip_l.json embed_dim (GlobalPool2D head, off-path)
PATCH_SIZE = 64                             # mobileclip_l.json image_cfg.patch_size
MODEL_PREFIX = "vision_tower.model."


def network_layout():
    """[(network index, kind, stage)] with kind in {"cpe","stage","down"} (mci.py:1357-1398)."""
    out, idx = [], 0
    for i in range(5):
        if i >= 3:
            out.append((idx, "cpe", i)); idx += 1
        out.append((idx, "stage", i)); idx += 1
        if i < 4:
            out.append((idx, "down", i)); idx += 1
    return out


def _conv(sp, name, cout, cin_g, k):
    sp[name + ".weight"] = ((cout, cin_g, k, k), torch.float32, False)
    sp[name + ".bias"] = ((cout,), torch.float32, False)


def _convffn(sp, p, c):
    sp[p + ".conv.conv.weight"] = ((c, 1, 7, 7), torch.float32, False)
    sp[p + ".conv.bn.weight"] = ((c,), torch.float32, False)
    sp[p + ".conv.bn.bias"] = ((c,), torch.float32, False)
    sp[p + ".conv.bn.running_mean"] = ((c,), torch.float32, True)
    sp[p + ".conv.bn.running_var"] = ((c,), torch.float32, True)
    sp[p + ".conv.bn.num_batches_tracked"] = ((), torch.int64, True)
    _conv(sp, p + ".fc1", MLP_RATIO * c, c, 1)
    _conv(sp, p + ".fc2", c, MLP_RATIO * c, 1)


def reference_param_specs():
    """OrderedDict key -> (shape, dtype, is_buffer), keys relative to the tower module
    (i.e. starting with "vision_tower.model."), in the reference's registration order."""
    sp = OrderedDict()
    P = MODEL_PREFIX
    c0 = EMBED_DIMS[0]
    _conv(sp, P + "patch_embed.0.reparam_conv", c0, 3, 3)
    _conv(sp, P + "patch_embed.1.reparam_conv", c0, 1, 3)
    _conv(sp, P + "patch_embed.2.reparam_conv", c0, c0, 1)
    for idx, kind, i in network_layout():
        c = EMBED_DIMS[i]
        n = P + f"network.{idx}"
        if kind == "cpe":
            _conv(sp, n + ".reparam_conv", c, 1, 7)
        elif kind == "down":
            co = EMBED_DIMS[i + 1]
            _conv(sp, n + ".proj.0.lkb_reparam", co, 1, 7)
            _conv(sp, n + ".proj.1.reparam_conv", co, co, 1)
        else:
            for b in range(LAYERS[i]):
                p = n + f".{b}"
                if TOKEN_MIXERS[i] == "repmixer":
                    sp[p + ".layer_scale"] = ((c, 1, 1), torch.float32, False)
                    _conv(sp, p + ".token_mixer.reparam_conv", c, 1, 3)
                    _convffn(sp, p + ".convffn", c)
                else:
                    sp[p + ".layer_scale_1"] = ((c, 1, 1), torch.float32, False)
                    sp[p + ".layer_scale_2"] = ((c, 1, 1), torch.float32, False)
                    sp[p + ".norm.weight"] = ((c,), torch.float32, False)
                    sp[p + ".norm.bias"] = ((c,), torch.float32, False)
                    sp[p + ".token_mixer.qkv.weight"] = ((3 * c, c), torch.float32, False)
                    sp[p + ".token_mixer.proj.weight"] = ((c, c), torch.float32, False)
                    sp[p + ".token_mixer.proj.bias"] = ((c,), torch.float32, False)
                    _convffn(sp, p + ".convffn", c)
    _conv(sp, P + "conv_exp.se.reduce", SE_RD, EMBED_DIM, 1)
    _conv(sp, P + "conv_exp.se.expand", EMBED_DIM, SE_RD, 1)
    _conv(sp, P + "conv_exp.reparam_conv", EMBED_DIM, 1, 3)
    sp[P + "head.proj"] = ((EMBED_DIM, PROJECTION_DIM), torch.float32, False)
    return sp


def projector_param_specs(mm_hidden, hidden, depth):
    """`mlp{depth}x_gelu` / `linear` projector keys (multimodal_projector/builder.py:20-30)."""
    sp = OrderedDict()
    fin = mm_hidden
    for d in range(depth):
        sp[f"{2 * d}.weight"] = ((hidden, fin), torch.float32, False)
        sp[f"{2 * d}.bias"] = ((hidden,), torch.float32, False)
        fin = hidden
    return sp
