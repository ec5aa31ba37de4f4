#!/usr/bin/env python
"""Launch the fused ConvFFN kernel (convffn.cuh) alone at a bench-size problem: for ncu captures and quick timings.
    python tools/convffn_probe.py C M [reps]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

import ml_fastvlm_b200 as pkg

C, M = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
eng = pkg.Engine(64, 0, 2, 1)
g = torch.Generator().manual_seed(C + M)
z = torch.randn(M, C, generator=g).to(torch.bfloat16).to(dev)
w1 = (torch.randn(4 * C, C, generator=g) / C ** 0.5).to(torch.bfloat16).to(dev)
w2 = (torch.randn(C, 4 * C, generator=g) / (4 * C) ** 0.5).to(torch.float16).to(dev)
b1 = torch.randn(4 * C, generator=g).to(dev)
b2 = torch.randn(C, generator=g).to(dev)
r = torch.randn(M, C, generator=g).to(torch.bfloat16).to(dev)
ts = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.convffn2(z, w1, b1, w2, b2, r)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print(f"convffn C={C} M={M}: {min(ts):.1f} us  {4.0 * M * C * 4 * C / min(ts) / 1e6:.0f} TFLOP/s")
