"""GPU diagnostic: per-kernel / per-unit parity table against the oracle (writes gpurun_out/diag_<R>.txt)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import ml_fastvlm_b200 as pkg  # noqa: E402
from oracle import fastvithd_oracle as orc  # noqa: E402
from oracle import fixture as fx  # noqa: E402


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    log = open(os.path.join(out_dir, f"diag_{R}.txt"), "w")

    def P(*a):
        s = " ".join(str(x) for x in a)
        print(s, flush=True)
        log.write(s + "\n")
        log.flush()

    dev = torch.device("cuda:0")
    P("device", torch.cuda.get_device_name(0))
    eng0 = pkg.Engine(64, 0, 2, 1)
    g = torch.Generator().manual_seed(0)
    for (M, N, K, act, use_b, use_r) in [(128, 128, 64, 0, 0, 0), (128, 128, 128, 0, 0, 0), (256, 128, 192, 0, 1, 0), (1000, 96, 96, 1, 1, 0),
                                          (4096, 1536, 384, 1, 1, 0), (4096, 384, 1536, 0, 1, 1), (16, 896, 3072, 1, 1, 0), (300, 2304, 768, 0, 0, 0),
                                          (512, 200, 256, 0, 1, 0)]:
        A = (torch.randn(M, K, generator=g)).to(torch.bfloat16).to(dev)
        W = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(dev)
        b = torch.randn(N, generator=g).to(dev) if use_b else None
        r = torch.randn(M, N, generator=g).to(torch.bfloat16).to(dev) if use_r else None
        try:
            D = eng0.gemm(A, W, b, r, act)
            torch.cuda.synchronize()
            ref = A.float() @ W.float().t()
            if b is not None:
                ref = ref + b
            if act:
                ref = torch.nn.functional.gelu(ref)
            if r is not None:
                ref = ref + r.float()
            P(f"gemm M={M} N={N} K={K} act={act} bias={use_b} res={use_r}: rel {rel(D.float(), ref):.3e} maxabs {(D.float() - ref).abs().max().item():.3e}")
        except Exception as e:  # noqa: BLE001
            P(f"gemm M={M} N={N} K={K}: EXC {e}")
            raise

    # determinism / race probe: same GEMM 20x, count distinct results
    for (M, N, K) in [(256, 384, 384), (256, 1536, 384), (256, 384, 1536), (64, 768, 768), (1024, 192, 768), (4096, 1536, 384)]:
        A = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
        W = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(dev)
        b = torch.randn(N, generator=g).to(dev)
        ref = torch.nn.functional.gelu(A.float() @ W.float().t() + b)
        errs = []
        for _ in range(20):
            D = eng0.gemm(A, W, b, None, 1)
            errs.append(rel(D.float(), ref))
        torch.cuda.synchronize()
        P(f"repeat gemm M={M} N={N} K={K}: min {min(errs):.3e} max {max(errs):.3e} bad {sum(e > 5e-3 for e in errs)}/20")

    sd = fx.tower_state_dict()
    psd = fx.projector_state_dict(896)
    x = fx.synthetic_images(1, R)
    col = {}
    t0 = time.time()
    ref = orc.encode_images(x, sd, psd, col)
    P(f"oracle R={R}: {time.time() - t0:.2f}s")
    packed = pkg.pack_tower(sd)
    packed.update(pkg.pack_projector(psd))
    eng = pkg.Engine(R, 896, 2, 1).load(packed, dev)
    units = eng.units()

    def nhwc(t):  # oracle NCHW fp32 -> [1, HW*C] bf16 on device
        return t.permute(0, 2, 3, 1).contiguous().reshape(1, -1).to(torch.bfloat16).to(dev)

    # per-unit, isolated: oracle input -> our output vs oracle output
    prev = None
    for u in units:
        name = u["name"]
        if name == "stem":
            xin = x.to(dev)
            want = col["stem"].permute(0, 2, 3, 1).reshape(1, -1)
        elif name == "conv_exp":
            xin = nhwc(prev)
            want = col["tokens"].reshape(1, -1)
        elif name == "projector":
            xin = col["tokens"].reshape(1, -1).to(torch.bfloat16).to(dev)
            want = ref.reshape(1, -1)
        else:
            xin = nhwc(prev)
            want = col[name].permute(0, 2, 3, 1).reshape(1, -1)
        try:
            got = eng.run_units(u["index"], u["index"], xin, 1)
            torch.cuda.synchronize()
            P(f"unit {u['index']:2d} {name:14s} isolated rel {rel(got.float().cpu(), want):.3e}  finite={bool(torch.isfinite(got.float()).all())}")
        except Exception as e:  # noqa: BLE001
            P(f"unit {u['index']:2d} {name:14s} EXC {e}")
            raise
        prev = col[name] if name in col else None
    # cumulative
    tokens, proj = eng.forward(x.to(dev), True, True)
    torch.cuda.synchronize()
    P(f"e2e tokens rel {rel(tokens.float().cpu(), col['tokens']):.3e}   projected rel {rel(proj.float().cpu(), ref):.3e}")
    ms = eng.profile_units(x.to(dev))
    ms = eng.profile_units(x.to(dev))
    P(f"profile (2nd run) total {sum(ms):.3f} ms")
    for u, m in zip(units, ms):
        P(f"   {u['name']:14s} {m * 1e3:9.1f} us   {u['flops'] / m / 1e9:9.1f} TFLOP/s   {u['min_bytes'] / m / 1e6:9.1f} GB/s(min-bytes)")


if __name__ == "__main__":
    main()
