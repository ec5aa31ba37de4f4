"""Row f3 -- LLM prefill (llava_qwen.py:57-143 -> transformers Qwen2ForCausalLM.forward): the library's prefill against the stock
Hugging Face model run in fp32 on the same random-init weights and the same input embeddings.

Oracle: transformers' Qwen2ForCausalLM itself (the third-party module the reference's LlavaQwen2ForCausalLM subclasses unchanged --
pinned version: the transformers wheel of this image), fp32, CPU.  Tolerance: bf16 weights / activations through the layers give
~1e-2 rel-L2 on the last-position logits; the HF model's own bf16 run is measured alongside as the floor.
"""
import pytest
import torch

import ml_fastvlm_b200 as pkg

LOGIT_TOL = 3e-2


def rel_l2(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _hf(hidden, layers, heads, kv, inter, vocab, seed=0):
    from transformers import Qwen2Config, Qwen2ForCausalLM
    torch.manual_seed(seed)
    cfg = Qwen2Config(hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads, num_key_value_heads=kv, intermediate_size=inter,
                      vocab_size=vocab, max_position_embeddings=4096, tie_word_embeddings=False)
    m = Qwen2ForCausalLM(cfg).eval()
    with torch.no_grad():                          # random biases / norm weights so that every operand matters
        for n, p in m.named_parameters():
            if n.endswith("bias"):
                p.normal_(0, 0.1)
            elif "layernorm" in n or n.endswith("norm.weight"):
                p.uniform_(0.5, 1.5)
    return m


def test_pack_qwen2_layout_cpu():
    m = _hf(128, 2, 2, 1, 256, 512)
    cfg = dict(layers=2)
    ws = pkg.pack_qwen2(m.state_dict(), cfg, "cpu")
    assert len(ws) == 2 * 7 + 2
    assert ws[1].shape == (128 + 2 * 64, 128) and ws[1].dtype == torch.bfloat16 and ws[2].shape == (256,) and ws[2].dtype == torch.float32
    assert ws[5].shape == (512, 128) and ws[6].shape == (128, 256) and ws[-1].shape == (512, 128) and ws[-2].dtype == torch.float32
    sd = m.state_dict()
    assert torch.equal(ws[1][128:192].float(), sd["model.layers.0.self_attn.k_proj.weight"].to(torch.bfloat16).float())
    assert torch.equal(ws[5][256:].float(), sd["model.layers.0.mlp.up_proj.weight"].to(torch.bfloat16).float())


def test_llm_prefill_needs_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(pkg.FvhdError):
        pkg.LlmPrefill(128, 2, 2, 1, 256, 512, 64, device="cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("shape,lengths", [
    ((128, 2, 2, 1, 256, 512), (1, 5, 33, 100, 287)),            # head_dim 64, ragged lengths incl. one token and > 4 key tiles
    ((256, 2, 2, 2, 512, 1024), (7, 130)),                       # head_dim 128 (the 7B geometry), no grouping
    ((896, 3, 14, 2, 4864, 151936), (287,)),                     # Qwen2-0.5B layer geometry (FastVLM-0.5B), 3 layers, config-3 length
])
def test_prefill_matches_hf_fp32(shape, lengths):
    dev = torch.device("cuda:0")
    m = _hf(*shape)
    H = shape[0]
    eng = pkg.LlmPrefill.from_hf(m, max_seq=max(lengths) + 8, device=dev)
    assert eng.launches(lengths[0]) == shape[1] * 8 + 3
    g = torch.Generator().manual_seed(11)
    for L in lengths:
        x = torch.randn(1, L, H, generator=g)
        xb = x.to(torch.bfloat16)
        with torch.no_grad():
            ref = m(inputs_embeds=xb.float(), use_cache=True)
            lref = ref.logits[0, -1]
            floor = rel_l2(m.to(torch.bfloat16)(inputs_embeds=xb).logits[0, -1].float(), lref)
            m.float()
        eng.input(L).copy_(xb[0].to(dev))
        tok, logits = eng.prefill(L, want_logits=True)
        err = rel_l2(logits.float(), lref)
        print(f"prefill {shape} L={L}: logits rel-L2 {err:.2e} (HF bf16 floor {floor:.2e}), token {tok} vs {int(lref.argmax())}")
        assert torch.isfinite(logits.float()).all()
        assert err < LOGIT_TOL and err < max(2.5 * floor, 1e-2), (L, err, floor)
        assert tok == int(logits.float().argmax())                              # argmax kernel == argmax of the returned logits
        # KV cache of the first and last layer against HF's (K after RoPE)
        kc, vc = eng.kv_cache()
        for li in (0, shape[1] - 1):
            kr = ref.past_key_values.layers[li].keys[0].transpose(0, 1)         # [L, kv, D]
            vr = ref.past_key_values.layers[li].values[0].transpose(0, 1)
            assert rel_l2(kc[li, :L].float(), kr) < 2e-2 and rel_l2(vc[li, :L].float(), vr) < 2e-2, li
        # rerun: bit-identical (graph replay)
        tok2, logits2 = eng.prefill(L, want_logits=True)
        assert tok2 == tok and torch.equal(logits, logits2)


@pytest.mark.gpu
def test_fma_attention_variant_matches_mma():
    """FVHD_LLM_ATTN=f selects the FMA-pipe attention kernel (first version); both kernels against each other on the same prefill."""
    import os
    dev = torch.device("cuda:0")
    m = _hf(128, 2, 2, 1, 256, 512, seed=3)
    x = torch.randn(70, 128, generator=torch.Generator().manual_seed(5)).to(torch.bfloat16)
    outs = []
    for env in (None, "f"):
        if env:
            os.environ["FVHD_LLM_ATTN"] = env
        try:
            eng = pkg.LlmPrefill.from_hf(m, max_seq=80, device=dev)
            eng.input(70).copy_(x.to(dev))
            outs.append(eng.prefill(70, want_logits=True)[1].float().cpu())
        finally:
            os.environ.pop("FVHD_LLM_ATTN", None)
    assert rel_l2(outs[0], outs[1]) < 8e-3 and not torch.equal(outs[0], outs[1])      # different kernels, same result up to bf16 rounding of P
