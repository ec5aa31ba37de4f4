"""Per-launch timing of one forward (CUDA events around every kernel): which launches the step time is made of."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ml_fastvlm_b200 as pkg
from oracle import fixture as fx


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    dev = torch.device("cuda:0")
    pk = pkg.pack_tower(fx.tower_state_dict())
    pk.update(pkg.pack_projector(fx.projector_state_dict(896)))
    eng = pkg.Engine(R, 896, 2, B).load(pk, dev)
    x = fx.synthetic_images(B, R).to(dev).to(torch.bfloat16)
    for _ in range(3):
        eng.forward(x, False, True)
    runs = np.array([eng.profile_steps(x) for _ in range(7)])
    ms = np.median(runs, axis=0)
    steps = eng.steps(B)
    units = eng.units() if hasattr(eng, "units") else None
    print(f"R={R} B={B}: {len(steps)} launches, sum {ms.sum():.3f} ms")
    agg = {}
    for s, t in zip(steps, ms):
        key = (s["unit"], s["kernel"])
        a = agg.setdefault(key, [0, 0.0, 0.0])
        a[0] += 1; a[1] += t; a[2] += s["flops"]
    last_unit = None
    for i, (s, t) in enumerate(zip(steps, ms)):
        tf = s["flops"] / (t * 1e-3) / 1e12 if t > 0 else 0
        gb = s["bytes"] / (t * 1e-3) / 1e9 if t > 0 else 0
        print(f"{i:4d} unit {s['unit']:3d} {s['kernel']:32s} {t * 1e3:8.1f} us  {tf:7.1f} TF/s {gb:8.1f} GB/s")


if __name__ == "__main__":
    main()
