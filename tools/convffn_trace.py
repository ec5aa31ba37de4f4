"""Per-CTA timeline (globaltimer stamps) of the 4-CTA-cluster ConvFFN kernel at the stage-2 shape (M = 4096, C = 384)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ml_fastvlm_b200 as pkg

NC = 6


def reference(z, w1, b1, w2, b2, resid):
    h = torch.nn.functional.gelu(z.float() @ w1.float().t() + b1).to(torch.bfloat16).float()
    return (h @ w2.float().t() + b2 + resid.float())


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    C = 384
    dev = torch.device("cuda:0")
    eng = pkg.Engine(64, 0, 2, 1)
    g = torch.Generator().manual_seed(0)
    z = torch.randn(M, C, generator=g).to(torch.bfloat16).to(dev)
    resid = torch.randn(M, C, generator=g).to(torch.bfloat16).to(dev)
    w1 = (torch.randn(4 * C, C, generator=g) / C ** 0.5).to(torch.bfloat16).to(dev)
    w2 = (torch.randn(C, 4 * C, generator=g) / (4 * C) ** 0.5).to(torch.bfloat16).to(dev)
    b1 = torch.randn(4 * C, generator=g).to(dev)
    b2 = torch.randn(C, generator=g).to(dev)
    out = eng.convffn(z, w1, b1, w2, b2, resid)
    ref = reference(z, w1, b1, w2, b2, resid)
    err = ((out.float() - ref).norm() / ref.norm()).item()
    print(f"M={M} C={C}: rel_l2 vs torch fp32 = {err:.2e}")
    nct = 148
    buf = torch.zeros(nct * 64, dtype=torch.int64, device=dev)
    for _ in range(3):
        eng.convffn(z, w1, b1, w2, b2, resid)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.convffn(z, w1, b1, w2, b2, resid)
    e1.record()
    torch.cuda.synchronize()
    print(f"back-to-back launches: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us each")
    eng.convffn(z, w1, b1, w2, b2, resid, trace=buf)
    torch.cuda.synchronize()
    t = buf.cpu().numpy().reshape(nct, 64).astype(np.float64)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    rel = (t - t0) / 1e3
    rel[t == 0] = np.nan
    med = np.nanmedian(rel, axis=0)
    mx = np.nanmax(rel, axis=0)

    def row(name, idx):
        print(f"  {name:34s} median " + " ".join(f"{med[i]:6.2f}" for i in idx) + "   | max " + " ".join(f"{mx[i]:6.2f}" for i in idx))

    print(f"{t.shape[0]} CTAs; times in us since the first CTA's entry")
    row("entry, setup done", [0, 1])
    row("z load issued, z landed (MMA)", [2, 3])
    row("W1(j) load issued", range(48, 48 + NC))
    row("W2(j) load issued", range(54, 54 + NC))
    row("MMA1(j) issued", range(8, 8 + NC))
    row("acc1(j) ready (epilogue)", range(24, 24 + NC))
    row("H(j) written", range(32, 32 + NC))
    row("MMA2(j) issued", range(16, 16 + NC))
    row("acc2 final / all ready / sent", [40, 41, 42])
    row("received / stored / exit", [43, 44, 45])


if __name__ == "__main__":
    main()
