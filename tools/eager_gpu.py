"""'Reference on this box' bar (second bar): the reference algorithm (oracle port = the same
torch.nn.functional calls the reference's modules make) in PyTorch eager bf16, channels-last, on the B200
(cuDNN / cuBLAS).  Measurement tool; prints one JSON line per batch size."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import fastvithd_oracle as orc
from oracle import fixture as fx


def eager_time(batch, res=1024, hidden=896, steps=10, warmup=3, dtype=torch.bfloat16, dev="cuda:0"):
    torch.backends.cudnn.benchmark = True
    sd = {k: v.to(dev).to(dtype) for k, v in fx.tower_state_dict().items()}
    for k in list(sd):                                  # conv weights channels-last as the activations
        if sd[k].dim() == 4:
            sd[k] = sd[k].contiguous(memory_format=torch.channels_last)
    psd = {k: v.to(dev).to(dtype) for k, v in fx.projector_state_dict(hidden).items()}
    x = fx.synthetic_images(batch, res).to(dev).to(dtype).contiguous(memory_format=torch.channels_last)

    def fwd():
        with torch.no_grad():
            f = orc.feature_select(orc.fastvit_forward(x, sd)).contiguous()
            return orc.mm_projector(f, psd)

    for _ in range(warmup):
        fwd()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        a.record()
        fwd()
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    med = ms[len(ms) // 2]
    return {"batch": batch, "ms_median": med, "ms_min": ms[0], "images_per_s": batch * 1e3 / med,
            "impl": "torch eager bf16 channels_last (cuDNN/cuBLAS), oracle port of the reference modules"}


if __name__ == "__main__":
    for b in [int(a) for a in sys.argv[1:]] or [1, 8]:
        print(json.dumps(eager_time(b)), flush=True)
