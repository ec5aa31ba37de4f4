"""FastVLM-0.5B p50 TTFT on one B200 (BASELINE.json configs[2]) -- encoder from this repo, LLM prefill from stock HF.

TTFT := pinned host uint8 RGB image (1024x1024) -> H2D (3 MB) -> fvhd_preprocess (row f1: resize/crop/1-255 on the GPU,
bit-exact with the reference's PIL path) -> encode_images (FastViTHD + mlp2x_gelu, this
library) -> splice the 256 visual tokens between the text embeddings (llava_arch.py:251-271 semantics) -> Qwen2-0.5B
prefill -> argmax of the last position = first generated token, synchronised.  Mirrors the app's definition
(app/FastVLM App/FastVLMModel.swift:114-138: from before `processor.prepare` to the first token).
The LLM is a random-init Qwen2ForCausalLM with the public Qwen2-0.5B shape (hidden 896, 24 layers, 14 heads / 2 KV heads,
MLP 4864, vocab 151936), bf16, HF eager/SDPA: the prefill is row f3 (out of the rebuilt path) and is reported as is.
"""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import ml_fastvlm_b200 as pkg  # noqa: E402
from oracle import fixture as fx  # noqa: E402


def main():
    from transformers import Qwen2Config, Qwen2ForCausalLM
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    cfg = Qwen2Config(hidden_size=896, num_hidden_layers=24, num_attention_heads=14, num_key_value_heads=2, intermediate_size=4864,
                      vocab_size=151936, max_position_embeddings=32768, tie_word_embeddings=True)
    llm = Qwen2ForCausalLM(cfg).to(device=dev, dtype=torch.bfloat16).eval()
    packed = pkg.pack_tower(fx.tower_state_dict())
    packed.update(pkg.pack_projector(fx.projector_state_dict(896)))
    eng = pkg.Engine(1024, 896, 2, 1).load(packed, dev)
    import numpy as np
    host_u8 = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (1024, 1024, 3), dtype=np.uint8)).pin_memory()
    img = torch.empty(1, 3, 1024, 1024, dtype=torch.float16, device=dev)
    n_pre, n_post = 14, 17                       # qwen_2 template around "<image>\nDescribe the image." (predict.py:34-42,80)
    ids = torch.randint(0, 150000, (1, n_pre + n_post), device=dev)
    embed = llm.get_input_embeddings()

    def one():
        t0 = time.perf_counter()
        pkg.preprocess_into(eng, host_u8, img[0])     # uint8 H2D + resize/crop/scale on the GPU
        x = torch.empty(1, n_pre + 256 + n_post, 896, dtype=torch.bfloat16, device=dev)
        eng.forward_into(img, x, n_pre)             # projector epilogue stores at the <image> position (row f2: no cat)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        txt = embed(ids)
        x[:, :n_pre] = txt[:, :n_pre]
        x[:, n_pre + 256:] = txt[:, n_pre:]
        out = llm(inputs_embeds=x, use_cache=True)
        tok = out.logits[:, -1].argmax(-1)
        tok.item()                                                                   # first token on the host
        t2 = time.perf_counter()
        return (t2 - t0) * 1e3, (t1 - t0) * 1e3, (t2 - t1) * 1e3

    with torch.inference_mode():
        for _ in range(5):
            one()
        runs = [one() for _ in range(50)]
    ttft = [r[0] for r in runs]
    enc = [r[1] for r in runs]
    pre = [r[2] for r in runs]
    res = {"metric": "fastvlm_0.5b_ttft_ms_p50", "value": statistics.median(ttft), "unit": "ms", "n_gpus": 1, "runs": 50, "warmup": 5,
           "encode_ms_p50": statistics.median(enc), "prefill_first_token_ms_p50": statistics.median(pre),
           "ttft_ms_min": min(ttft), "ttft_ms_p90": sorted(ttft)[44], "sequence": n_pre + 256 + n_post,
           "llm": "random-init Qwen2ForCausalLM (0.5B shape), bf16, stock transformers %s" % __import__("transformers").__version__,
           "encoder": "libfastvithd_b200 (this repo), 1024x1024, batch 1", "includes": "uint8 H2D + GPU preprocessing (row f1) + tower + projector (splice store, row f2) + stock-HF prefill"}
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
