"""Per-CTA timeline of the tcgen05 GEMM (globaltimer stamps) for representative layer shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ml_fastvlm_b200 as pkg

NAMES = ["entry", "setup", "pdl", "ld1", "ldN", "kb0", "mma1", "mmaN", "acc1", "epi1", "epiN", "drain", "exit"]


def main():
    dev = torch.device("cuda:0")
    eng = pkg.Engine(64, 0, 2, 1)
    lib = eng.lib
    g = torch.Generator().manual_seed(0)
    shapes = [("s2.fc1", 4096, 1536, 384, 1), ("s2.fc2", 4096, 384, 1536, 0), ("s1.fc1", 16384, 768, 192, 1), ("s1.fc2", 16384, 192, 768, 0),
              ("s0.fc1", 65536, 384, 96, 1), ("s3.proj", 1024, 768, 768, 0), ("proj2", 256, 896, 896, 0)]
    variants = [(0, 1), (0, 4), (128, 1), (256, 1), (64, 1)]
    if len(sys.argv) > 1 and sys.argv[1] == "llm":      # the four GEMMs of a Qwen2-0.5B layer at 287 tokens + lm_head, cost-model configuration
        shapes = [("qkv", 287, 1152, 896, 0), ("o", 287, 896, 896, 0), ("gate_up", 287, 9728, 896, 0), ("down", 287, 896, 4864, 0), ("lm_head", 1, 151936, 896, 0)]
        variants = [(0, 1), (128, 1), (256, 1)]
    for name, M, N, K, act in shapes:
        A = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
        W = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(dev)
        b = torch.randn(N, generator=g).to(dev)
        for bn, cs in variants:
            if bn and N % bn:
                continue
            buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
            lib.fvhd_debug_gemm_trace(buf.data_ptr(), bn, cs)
            for _ in range(3):
                eng.gemm(A, W, b, None, act)
            torch.cuda.synchronize()
            buf.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.gemm(A, W, b, None, act)
            e1.record()
            torch.cuda.synchronize()
            t = buf.cpu().numpy().reshape(148, 16).astype(np.float64)
            used = t[:, 0] > 0
            t = t[used]
            t0 = t[:, 0].min()
            rel = (t - t0) / 1e3
            rel[t == 0] = np.nan
            med = np.nanmedian(rel, axis=0)
            mx = np.nanmax(rel, axis=0)
            print(f"{name} M={M} N={N} K={K} forceBN={bn} maxCS={cs}: event {e0.elapsed_time(e1) * 1e3:.1f} us, ctas {used.sum()}")
            print("   median us: " + " ".join(f"{n}={v:.1f}" for n, v in zip(NAMES, med[:13])))
            print("   max    us: " + " ".join(f"{n}={v:.1f}" for n, v in zip(NAMES, mx[:13])))
    lib.fvhd_debug_gemm_trace(None, 0, 4)


if __name__ == "__main__":
    main()
