"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the reference FastViTHD + mm_projector path.

This file is the parity oracle for the CUDA library.  It is NOT a product path: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import it.

Every function restates one reference function with plain `torch.nn.functional` calls on a
reference-keyed state-dict (no nn.Module, no timm) and cites the reference lines it follows
(paths relative to /root/reference/llava/model/).  It is pinned against the *unmodified* reference
executed under oracle/timm_stub.py by oracle/gen_golden.py (outputs committed in tests/golden/),
because the reference itself ships no tests or golden vectors ("parity unpinned" by the
reference's own tests; pinned here by live execution of the reference files -- see DESIGN.md).
"""
import math

import torch
import torch.nn.functional as F

from . import fixture as fx

P = fx.TOWER_PREFIX


def _w(sd, name):
    return sd[name + ".weight"], sd.get(name + ".bias")


def mobileone_infer(x, sd, name, stride, padding, groups, act=True):
    """MobileOneBlock inference path: act(se(reparam_conv(x))) without SE
    (multimodal_encoder/mobileclip/mci.py:197-198)."""
    w, b = _w(sd, name + ".reparam_conv")
    x = F.conv2d(x, w, b, stride=stride, padding=padding, groups=groups)
    return F.gelu(x) if act else x


def convolutional_stem(x, sd):
    """mci.py:553-603: conv3x3 s2 -> dw3x3 s2 -> 1x1, GELU after each."""
    c = fx.EMBED_DIMS[0]
    x = mobileone_infer(x, sd, P + "patch_embed.0", 2, 1, 1)
    x = mobileone_infer(x, sd, P + "patch_embed.1", 2, 1, c)
    x = mobileone_infer(x, sd, P + "patch_embed.2", 1, 0, 1)
    return x


def repmixer(x, sd, name):
    """RepMixer.forward, reparameterised branch (mci.py:808-811)."""
    w, b = _w(sd, name + ".reparam_conv")
    return F.conv2d(x, w, b, stride=1, padding=1, groups=x.shape[1])


def convffn(x, sd, name):
    """ConvFFN.forward (mci.py:920-927): dw7x7 -> BN(eval) -> fc1 -> GELU -> fc2 (dropout p=0)."""
    c = x.shape[1]
    x = F.conv2d(x, sd[name + ".conv.conv.weight"], None, padding=3, groups=c)
    x = F.batch_norm(x, sd[name + ".conv.bn.running_mean"], sd[name + ".conv.bn.running_var"],
                     sd[name + ".conv.bn.weight"], sd[name + ".conv.bn.bias"], training=False, eps=1e-5)
    x = F.conv2d(x, *_w(sd, name + ".fc1"))
    x = F.gelu(x)
    x = F.conv2d(x, *_w(sd, name + ".fc2"))
    return x


def repmixer_block(x, sd, name):
    """RepMixerBlock.forward with layer scale (mci.py:1106-1109)."""
    x = repmixer(x, sd, name + ".token_mixer")
    return x + sd[name + ".layer_scale"] * convffn(x, sd, name + ".convffn")


def patch_embed(x, sd, name):
    """PatchEmbed.forward (mci.py:739-741): ReparamLargeKernelConv (442-451, dw7x7 s2, groups=Cin,
    Cout=2Cin, GELU) then MobileOneBlock 1x1 + GELU."""
    cin = x.shape[1]
    w, b = _w(sd, name + ".proj.0.lkb_reparam")
    x = F.gelu(F.conv2d(x, w, b, stride=2, padding=3, groups=cin))
    return mobileone_infer(x, sd, name + ".proj.1", 1, 0, 1)


def repcpe(x, sd, name):
    """RepCPE.forward, reparameterised (mci.py:992-995)."""
    w, b = _w(sd, name + ".reparam_conv")
    return F.conv2d(x, w, b, stride=1, padding=3, groups=x.shape[1])


def layernorm_channel(x, w, b, eps=1e-5):
    """LayerNormChannel.forward (mci.py:617-623)."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return w[None, :, None, None] * x + b[None, :, None, None]


def mhsa(x, sd, name):
    """MHSA.forward (mci.py:661-685): qkv (no bias) -> heads of 32 -> softmax((q*scale) k^T) v -> proj."""
    B, C, H, W = x.shape
    N = H * W
    hd = fx.HEAD_DIM
    nh = C // hd
    t = torch.flatten(x, start_dim=2).transpose(-2, -1)               # (B, N, C)
    qkv = F.linear(t, sd[name + ".qkv.weight"]).reshape(B, N, 3, nh, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    attn = attn.softmax(dim=-1)
    t = (attn @ v).transpose(1, 2).reshape(B, N, C)
    t = F.linear(t, sd[name + ".proj.weight"], sd[name + ".proj.bias"])
    return t.transpose(-2, -1).reshape(B, C, H, W)


def attention_block(x, sd, name):
    """AttentionBlock.forward with layer scale (mci.py:1185-1188)."""
    n = layernorm_channel(x, sd[name + ".norm.weight"], sd[name + ".norm.bias"])
    x = x + sd[name + ".layer_scale_1"] * mhsa(n, sd, name + ".token_mixer")
    x = x + sd[name + ".layer_scale_2"] * convffn(x, sd, name + ".convffn")
    return x


def se_block(x, sd, name):
    """SEBlock.forward (mci.py:72-81)."""
    b, c, h, w = x.shape
    s = F.avg_pool2d(x, kernel_size=[h, w])
    s = F.relu(F.conv2d(s, *_w(sd, name + ".reduce")))
    s = torch.sigmoid(F.conv2d(s, *_w(sd, name + ".expand")))
    return x * s.view(-1, c, 1, 1)


def conv_exp(x, sd):
    """conv_exp = MobileOneBlock(use_se=True): act(se(conv(x))) (mci.py:1401-1411, 198)."""
    w, b = _w(sd, P + "conv_exp.reparam_conv")
    x = F.conv2d(x, w, b, stride=1, padding=1, groups=x.shape[1])
    x = se_block(x, sd, P + "conv_exp.se")
    return F.gelu(x)


def fastvit_forward(images, sd, collect=None):
    """FastViT.forward(..., return_image_embeddings=True)["image_embeddings"] (mci.py:1436-1451).
    The GlobalPool2D head output ("logits") is discarded by the tower
    (mobileclip_encoder.py:62) and is not computed.  `collect`, if a dict, receives every unit's
    NCHW output keyed "stem", "network.<i>" / "network.<i>.<b>", "conv_exp"."""
    x = convolutional_stem(images, sd)
    if collect is not None:
        collect["stem"] = x
    for idx, kind, i in fx.network_layout():
        n = P + f"network.{idx}"
        if kind == "cpe":
            x = repcpe(x, sd, n)
        elif kind == "down":
            x = patch_embed(x, sd, n)
        else:
            blk = repmixer_block if fx.TOKEN_MIXERS[i] == "repmixer" else attention_block
            for b in range(fx.LAYERS[i]):
                x = blk(x, sd, n + f".{b}")
                if collect is not None:
                    collect[f"network.{idx}.{b}"] = x
        if collect is not None:
            collect[f"network.{idx}"] = x
    x = conv_exp(x, sd)
    if collect is not None:
        collect["conv_exp"] = x
    return x


def feature_select(emb):
    """MobileCLIPVisionTower.feature_select (mobileclip_encoder.py:60-68): NCHW -> [B, HW, C]."""
    B, C, H, W = emb.shape
    return emb.reshape(B, C, H * W).transpose(1, 2)


def tower_forward(images, sd, collect=None):
    """MobileCLIPVisionTower.forward_images for a batched tensor (mobileclip_encoder.py:84-86), fp32."""
    with torch.no_grad():
        return feature_select(fastvit_forward(images.float(), sd, collect)).contiguous()


def mm_projector(feats, psd):
    """mlp{N}x_gelu projector (multimodal_projector/builder.py:23-30)."""
    depth = len([k for k in psd if k.endswith(".weight")])
    x = feats
    for d in range(depth):
        if d > 0:
            x = F.gelu(x)
        x = F.linear(x, psd[f"{2 * d}.weight"], psd[f"{2 * d}.bias"])
    return x


def encode_images(images, sd, psd, collect=None):
    """LlavaMetaForCausalLM.encode_images (llava_arch.py:141-144): mm_projector(tower(images))."""
    with torch.no_grad():
        feats = tower_forward(images, sd, collect)
        if collect is not None:
            collect["tokens"] = feats
        return mm_projector(feats, psd)


def num_tokens(res):
    """(R/64)^2 visual tokens (mobileclip_l.json:7 patch_size 64; mobileclip_encoder.py:111-116)."""
    return (res // 64) ** 2


def gmacs_per_image(res):
    """MACs of conv/Linear/QK^T/PV on the tower path for one RxR image (probe-verified
    243.35 G at 1024).  Elementwise, GELU, softmax and LN excluded."""
    m = 0
    h = res // 2
    m += h * h * 96 * 27
    h //= 2
    m += h * h * 96 * 9 + h * h * 96 * 96
    for i, c in enumerate(fx.EMBED_DIMS):
        px = h * h
        for _ in range(fx.LAYERS[i]):
            m += px * c * 49 + 2 * px * c * 4 * c          # dw7x7 + fc1 + fc2
            if fx.TOKEN_MIXERS[i] == "repmixer":
                m += px * c * 9
            else:
                m += px * c * 3 * c + px * c * c + 2 * px * px * c
        if i >= 3:
            pass
        if i < 4:
            co = fx.EMBED_DIMS[i + 1]
            h //= 2
            m += h * h * co * 49 + h * h * co * co
            if i + 1 >= 3:
                m += h * h * co * 49                        # RepCPE of the next stage
    ce = 3072
    m += h * h * ce * 9 + 2 * ce * 192
    return m / 1e9
