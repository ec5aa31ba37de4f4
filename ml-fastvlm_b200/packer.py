"""Reference state-dict -> packed device weights of libfastvithd_b200 (the checkpoint-ingest step).

Offline folds (exact in fp32, then rounded once to the storage dtype):
  * eval-mode BatchNorm of every ConvFFN.conv (mci.py:885-907) into the dw7x7 weight + a bias,
  * layer scales (mci.py:1107-1109, 1185-1188) into fc2 / proj weights and biases,
  * NCHW conv kernels -> tap-major [k*k][C] (depthwise, fp32) and [N][K] K-major bf16 (1x1 / Linear GEMM B operand).
Input keys are those of SURVEY.md 2.2: `[model.vision_tower.]vision_tower.model.<k>`; the prefix is detected.
"""
from collections import OrderedDict

import torch

from . import arch

BN_EPS = 1e-5   # nn.BatchNorm2d default (mci.py:890)


def _strip_prefix(sd):
    probe = "patch_embed.0.reparam_conv.weight"
    for k in sd:
        if k.endswith(probe):
            pre = k[: -len(probe)]
            return OrderedDict((kk[len(pre):], v) for kk, v in sd.items() if kk.startswith(pre))
    raise KeyError(f"no '*{probe}' key: not a FastViTHD tower state-dict")


def _dw(w):
    """[C,1,k,k] depthwise kernel -> [k*k, C] fp32."""
    c, _, k, _ = w.shape
    return w[:, 0].permute(1, 2, 0).reshape(k * k, c).contiguous().float()


def _pw(w):
    """[N,K,1,1] or [N,K] -> [N,K] bf16 (K-major GEMM B operand)."""
    return w.reshape(w.shape[0], w.shape[1]).contiguous().to(torch.bfloat16)


def _convffn(out, sd, src, dst, layer_scale, f16_fc2=False):
    w = sd[src + ".conv.conv.weight"].float()
    s = sd[src + ".conv.bn.weight"].float() / torch.sqrt(sd[src + ".conv.bn.running_var"].float() + BN_EPS)
    out[dst + "dw.w"] = _dw(w * s[:, None, None, None])
    out[dst + "dw.b"] = (sd[src + ".conv.bn.bias"].float() - sd[src + ".conv.bn.running_mean"].float() * s).contiguous()
    out[dst + "fc1.w"] = _pw(sd[src + ".fc1.weight"].float())
    out[dst + "fc1.b"] = sd[src + ".fc1.bias"].float().contiguous()
    ls = layer_scale.float().reshape(-1)
    w2 = sd[src + ".fc2.weight"].float().reshape(ls.numel(), -1)
    out[dst + "fc2.w"] = _pw(w2 * ls[:, None])
    if f16_fc2:     # RepMixer blocks: the fused ConvFFN kernel keeps the hidden in f16 (packed-half GELU) and multiplies it by an f16 copy
        out[dst + "fc2.wh"] = (w2 * ls[:, None]).contiguous().to(torch.float16)
    out[dst + "fc2.b"] = (sd[src + ".fc2.bias"].float() * ls).contiguous()


def pack_tower(state_dict):
    """-> OrderedDict packed name -> CPU tensor (fp32, bf16 or f16), names as fvhd_weight_spec reports."""
    sd = _strip_prefix(state_dict)
    out = OrderedDict()
    w0 = sd["patch_embed.0.reparam_conv.weight"].float()                 # [96,3,3,3]
    out["stem.w0"] = w0.permute(1, 2, 3, 0).reshape(27, w0.shape[0]).contiguous()
    out["stem.b0"] = sd["patch_embed.0.reparam_conv.bias"].float().contiguous()
    out["stem.w1"] = _dw(sd["patch_embed.1.reparam_conv.weight"])
    out["stem.b1"] = sd["patch_embed.1.reparam_conv.bias"].float().contiguous()
    out["stem.w2"] = _pw(sd["patch_embed.2.reparam_conv.weight"].float())
    out["stem.b2"] = sd["patch_embed.2.reparam_conv.bias"].float().contiguous()
    for idx, kind, i in arch.network_layout():
        n = f"network.{idx}"
        if kind == "cpe":
            out[n + ".dw.w"] = _dw(sd[n + ".reparam_conv.weight"])
            out[n + ".dw.b"] = sd[n + ".reparam_conv.bias"].float().contiguous()
        elif kind == "down":
            out[n + ".dw.w"] = _dw(sd[n + ".proj.0.lkb_reparam.weight"])
            out[n + ".dw.b"] = sd[n + ".proj.0.lkb_reparam.bias"].float().contiguous()
            out[n + ".pw.w"] = _pw(sd[n + ".proj.1.reparam_conv.weight"].float())
            out[n + ".pw.b"] = sd[n + ".proj.1.reparam_conv.bias"].float().contiguous()
        else:
            for b in range(arch.LAYERS[i]):
                p = f"{n}.{b}"
                d = p + "."
                if arch.TOKEN_MIXERS[i] == "repmixer":
                    out[d + "mix.w"] = _dw(sd[p + ".token_mixer.reparam_conv.weight"])
                    out[d + "mix.b"] = sd[p + ".token_mixer.reparam_conv.bias"].float().contiguous()
                    _convffn(out, sd, p + ".convffn", d, sd[p + ".layer_scale"], f16_fc2=True)
                else:
                    out[d + "ln.w"] = sd[p + ".norm.weight"].float().contiguous()
                    out[d + "ln.b"] = sd[p + ".norm.bias"].float().contiguous()
                    out[d + "qkv.w"] = _pw(sd[p + ".token_mixer.qkv.weight"].float())
                    ls1 = sd[p + ".layer_scale_1"].float().reshape(-1)
                    out[d + "proj.w"] = _pw(sd[p + ".token_mixer.proj.weight"].float() * ls1[:, None])
                    out[d + "proj.b"] = (sd[p + ".token_mixer.proj.bias"].float() * ls1).contiguous()
                    _convffn(out, sd, p + ".convffn", d, sd[p + ".layer_scale_2"])
    out["conv_exp.dw.w"] = _dw(sd["conv_exp.reparam_conv.weight"])
    out["conv_exp.dw.b"] = sd["conv_exp.reparam_conv.bias"].float().contiguous()
    out["conv_exp.se.r.w"] = _pw(sd["conv_exp.se.reduce.weight"].float())
    out["conv_exp.se.r.b"] = sd["conv_exp.se.reduce.bias"].float().contiguous()
    out["conv_exp.se.e.w"] = _pw(sd["conv_exp.se.expand.weight"].float())
    out["conv_exp.se.e.b"] = sd["conv_exp.se.expand.bias"].float().contiguous()
    return out


def pack_projector(state_dict):
    """`{0,2}.{weight,bias}` (optionally prefixed, e.g. `model.mm_projector.`) -> projector.* tensors."""
    keys = [k for k in state_dict if k.endswith("0.weight")]
    if not keys:
        raise KeyError("no '*0.weight' key: not an mlp{N}x_gelu / linear projector state-dict")
    pre = keys[0][: -len("0.weight")]
    out = OrderedDict()
    d = 0
    while f"{pre}{2 * d}.weight" in state_dict:
        out[f"projector.{2 * d}.w"] = _pw(state_dict[f"{pre}{2 * d}.weight"].float())
        out[f"projector.{2 * d}.b"] = state_dict[f"{pre}{2 * d}.bias"].float().contiguous()
        d += 1
    return out
