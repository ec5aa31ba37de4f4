#!/usr/bin/env python
"""One Qwen2-0.5B-shaped prefill (287 tokens) on the library, random-init weights: for ncu launch lists and quick timings.
    python tools/llm_probe.py [layers=24] [L=287]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from transformers import Qwen2Config, Qwen2ForCausalLM

import ml_fastvlm_b200 as pkg

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 24
L = int(sys.argv[2]) if len(sys.argv) > 2 else 287
dev = torch.device("cuda:0")
torch.manual_seed(0)
cfg = Qwen2Config(hidden_size=896, num_hidden_layers=layers, num_attention_heads=14, num_key_value_heads=2, intermediate_size=4864,
                  vocab_size=151936, max_position_embeddings=32768, tie_word_embeddings=True)
with torch.device(dev):
    m = Qwen2ForCausalLM(cfg)
m = m.to(torch.bfloat16).eval()
eng = pkg.LlmPrefill.from_hf(m, max_seq=L, device=dev)
eng.input(L).copy_(torch.randn(L, 896, device=dev).to(torch.bfloat16))
for _ in range(3):
    eng.prefill(L)
ts = []
for _ in range(20):
    t0 = time.perf_counter()
    eng.prefill(L)
    ts.append((time.perf_counter() - t0) * 1e3)
ts.sort()
print(f"prefill L={L} layers={layers}: {ts[len(ts) // 2]:.3f} ms median, {ts[0]:.3f} min, {eng.launches(L)} launches")
