"""TEST INFRASTRUCTURE -- fp32 torch emulation of what the CUDA kernels compute FROM THE PACKED WEIGHTS
(NHWC, tap-major depthwise kernels, K-major GEMM weights, BN / layer-scale folds).  Validates
ml_fastvlm_b200.packer against the oracle on CPU, and documents each kernel's contract."""
import torch
import torch.nn.functional as F

from ml_fastvlm_b200 import arch


def dw(x, w_taps, b, k, stride=1, mult=1):
    """x NHWC [B,H,W,C]; w_taps [k*k, C*mult]; out channel o reads in channel o // mult."""
    B, H, W, C = x.shape
    w = w_taps.float().reshape(k, k, C * mult).permute(2, 0, 1).unsqueeze(1)
    y = F.conv2d(x.permute(0, 3, 1, 2), w, b.float(), stride=stride, padding=k // 2, groups=C)
    return y.permute(0, 2, 3, 1)


def gemm(a, w, b=None, act=False, res=None):
    y = a @ w.float().t()
    if b is not None:
        y = y + b.float()
    if act:
        y = F.gelu(y)
    if res is not None:
        y = y + res
    return y


def convffn(pk, p, x_for_dw, resid):
    c = x_for_dw.shape[-1]
    z = dw(x_for_dw, pk[p + "dw.w"], pk[p + "dw.b"], 7)
    h = gemm(z, pk[p + "fc1.w"], pk[p + "fc1.b"], act=True)
    return gemm(h, pk[p + "fc2.w"], pk[p + "fc2.b"], res=resid)


def forward(images, pk, collect=None):
    """images NCHW fp32 -> tokens [B, HW, 3072] (and projected if projector.* present)."""
    w0 = pk["stem.w0"].reshape(3, 3, 3, 96).permute(3, 0, 1, 2)           # [(ci,ky,kx), co] -> [co,ci,ky,kx]
    x = F.gelu(F.conv2d(images, w0, pk["stem.b0"], stride=2, padding=1)).permute(0, 2, 3, 1)
    x = F.gelu(dw(x, pk["stem.w1"], pk["stem.b1"], 3, stride=2))
    x = gemm(x, pk["stem.w2"], pk["stem.b2"], act=True)
    if collect is not None:
        collect["stem"] = x
    for idx, kind, i in arch.network_layout():
        n = f"network.{idx}"
        if kind == "cpe":
            x = dw(x, pk[n + ".dw.w"], pk[n + ".dw.b"], 7)
        elif kind == "down":
            x = F.gelu(dw(x, pk[n + ".dw.w"], pk[n + ".dw.b"], 7, stride=2, mult=2))
            x = gemm(x, pk[n + ".pw.w"], pk[n + ".pw.b"], act=True)
        else:
            for b in range(arch.LAYERS[i]):
                p = f"{n}.{b}."
                if arch.TOKEN_MIXERS[i] == "repmixer":
                    y = dw(x, pk[p + "mix.w"], pk[p + "mix.b"], 3)
                    x = convffn(pk, p, y, y)
                else:
                    B, H, W, C = x.shape
                    u = x.mean(-1, keepdim=True)
                    s = (x - u).pow(2).mean(-1, keepdim=True)
                    nx = (x - u) / torch.sqrt(s + 1e-5) * pk[p + "ln.w"] + pk[p + "ln.b"]
                    qkv = gemm(nx.reshape(B, H * W, C), pk[p + "qkv.w"]).reshape(B, H * W, 3, C // 32, 32).permute(2, 0, 3, 1, 4)
                    q, k, v = qkv.unbind(0)
                    a = ((q * 32 ** -0.5) @ k.transpose(-2, -1)).softmax(-1)
                    o = (a @ v).transpose(1, 2).reshape(B, H, W, C)
                    x1 = gemm(o, pk[p + "proj.w"], pk[p + "proj.b"], res=x)
                    x = convffn(pk, p, x1, x1)
                if collect is not None:
                    collect[f"{n}.{b}"] = x
        if collect is not None:
            collect[n] = x
    c = dw(x, pk["conv_exp.dw.w"], pk["conv_exp.dw.b"], 3, mult=2)
    B, H, W, C = c.shape
    pooled = c.mean((1, 2))
    r = F.relu(gemm(pooled, pk["conv_exp.se.r.w"], pk["conv_exp.se.r.b"]))
    s = torch.sigmoid(gemm(r, pk["conv_exp.se.e.w"], pk["conv_exp.se.e.b"]))
    tokens = F.gelu(c * s[:, None, None, :]).reshape(B, H * W, C)
    if collect is not None:
        collect["conv_exp"] = tokens
    proj = None
    if "projector.0.w" in pk:
        proj = gemm(tokens, pk["projector.0.w"], pk["projector.0.b"], act="projector.2.w" in pk)
        if "projector.2.w" in pk:
            proj = gemm(proj, pk["projector.2.w"], pk["projector.2.b"])
    return tokens, proj
