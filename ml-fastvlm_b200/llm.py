"""LLM prefill behind the tower: the first forward of `LlavaQwen2ForCausalLM.generate` (llava_qwen.py:57-143 ->
transformers `Qwen2ForCausalLM.forward`, predict.py:58-65) for one spliced embedding sequence, on the C library.

Host code is plumbing: it packs a Hugging Face Qwen2 state-dict into the operand layout of `fvhd_llm_load` (fused q/k/v and gate/up
weights, fp32 norm weights and biases) and hands raw device pointers to the C ABI.  No fallback: without the CUDA library or a GPU the
constructor raises.
"""
import ctypes as C

import torch

from . import lib as L


class _DeviceView:
    """`__cuda_array_interface__` wrapper so torch can view a library-owned device buffer without copying."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": "<i2", "version": 3}


def pack_qwen2(state_dict, cfg, device):
    """HF Qwen2 state-dict (keys `model.layers.N....`, `model.norm.weight`, `lm_head.weight` / tied `model.embed_tokens.weight`)
    -> the 7 * layers + 2 device tensors fvhd_llm_load expects."""
    sd = state_dict
    out = []
    bf = lambda t: t.detach().to(device=device, dtype=torch.bfloat16).contiguous()
    f32 = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
    for i in range(cfg["layers"]):
        p = f"model.layers.{i}."
        out.append(f32(sd[p + "input_layernorm.weight"]))
        out.append(bf(torch.cat([sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.k_proj.weight"], sd[p + "self_attn.v_proj.weight"]], 0)))
        out.append(f32(torch.cat([sd[p + "self_attn.q_proj.bias"], sd[p + "self_attn.k_proj.bias"], sd[p + "self_attn.v_proj.bias"]], 0)))
        out.append(bf(sd[p + "self_attn.o_proj.weight"]))
        out.append(f32(sd[p + "post_attention_layernorm.weight"]))
        out.append(bf(torch.cat([sd[p + "mlp.gate_proj.weight"], sd[p + "mlp.up_proj.weight"]], 0)))
        out.append(bf(sd[p + "mlp.down_proj.weight"]))
    out.append(f32(sd["model.norm.weight"]))
    head = sd["lm_head.weight"] if "lm_head.weight" in sd else sd["model.embed_tokens.weight"]
    out.append(bf(head))
    return out


class LlmPrefill:
    """Qwen2 prefill engine: `input(L)` is the [L, hidden] bf16 device buffer to fill with the spliced sequence (the projector epilogue
    can store the visual tokens straight into it), `prefill(L)` runs the pass and returns the first token."""

    def __init__(self, hidden, layers, heads, kv_heads, intermediate, vocab, max_seq, rope_theta=1e6, rms_eps=1e-6, device="cuda"):
        self.lib = L.load_library()
        dev = torch.device(device)
        if dev.type != "cuda":
            raise L.FvhdError("LlmPrefill runs on a CUDA device only (no CPU fallback)")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self.cfg = dict(hidden=hidden, layers=layers, heads=heads, kv_heads=kv_heads, head_dim=hidden // heads, intermediate=intermediate,
                        vocab=vocab, max_seq=max_seq, rope_theta=float(rope_theta), rms_eps=float(rms_eps))
        hcfg = L.FvhdConfig(64, 0, 2, 1)                      # the handle only hosts the LLM; its (unused) tower plan is the smallest one
        self.handle = C.c_void_p()
        L.check(self.lib.fvhd_create(C.byref(hcfg), C.byref(self.handle)))
        self._weights = None
        self._tok = torch.zeros(1, dtype=torch.int32).pin_memory()

    @classmethod
    def from_hf(cls, model, max_seq, device="cuda"):
        """Build from a transformers Qwen2ForCausalLM (or LlavaQwen2ForCausalLM: same decoder keys)."""
        c = model.config
        rp = getattr(c, "rope_parameters", None) or {}
        theta = rp.get("rope_theta") or getattr(c, "rope_theta", None) or 10000.0
        self = cls(c.hidden_size, c.num_hidden_layers, c.num_attention_heads, c.num_key_value_heads, c.intermediate_size, c.vocab_size, max_seq,
                   rope_theta=theta, rms_eps=c.rms_norm_eps, device=device)
        self.load(model.state_dict())
        return self

    def load(self, state_dict):
        ws = pack_qwen2(state_dict, self.cfg, self.device)
        arr = (C.c_void_p * len(ws))(*[w.data_ptr() for w in ws])
        cfg = L.FvhdLlmConfig(self.cfg["hidden"], self.cfg["layers"], self.cfg["heads"], self.cfg["kv_heads"], self.cfg["head_dim"],
                              self.cfg["intermediate"], self.cfg["vocab"], self.cfg["max_seq"], self.cfg["rope_theta"], self.cfg["rms_eps"])
        with torch.cuda.device(self.device):
            L.check(self.lib.fvhd_llm_load(self.handle, C.byref(cfg), arr, len(ws)), self.handle)
        self._weights = ws                                    # caller-owned operands: keep them alive
        ptr = self.lib.fvhd_llm_input(self.handle)
        self._x = torch.as_tensor(_DeviceView(ptr, (self.cfg["max_seq"], self.cfg["hidden"])), device=self.device).view(torch.bfloat16)
        return self

    def input(self, length):
        return self._x[:length]

    def launches(self, length):
        return int(self.lib.fvhd_llm_launches(self.handle, int(length)))

    def prefill(self, length, want_logits=False, sync=True):
        """Run the prefill over input(length); returns (first token id, logits [V] bf16 or None)."""
        if self._weights is None:
            raise L.FvhdError("LlmPrefill.load() has not been called")
        logits = torch.empty(self.cfg["vocab"], dtype=torch.bfloat16, device=self.device) if want_logits else None
        with torch.cuda.device(self.device):
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            L.check(self.lib.fvhd_llm_prefill(self.handle, st, int(length), logits.data_ptr() if want_logits else None,
                                              C.c_void_p(self._tok.data_ptr())), self.handle)
            if sync:
                torch.cuda.current_stream(self.device).synchronize()
        return (int(self._tok.item()) if sync else None), logits

    def kv_cache(self):
        """(K, V) views [layers, max_seq, kv_heads, head_dim] bf16 written by the last prefill (K after RoPE)."""
        k, v = C.c_void_p(), C.c_void_p()
        L.check(self.lib.fvhd_llm_kv_cache(self.handle, C.byref(k), C.byref(v)), self.handle)
        shape = (self.cfg["layers"], self.cfg["max_seq"], self.cfg["kv_heads"], self.cfg["head_dim"])
        mk = lambda p: torch.as_tensor(_DeviceView(p.value, shape), device=self.device).view(torch.bfloat16)
        return mk(k), mk(v)

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.handle.value:
                self.lib.fvhd_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:  # noqa: BLE001
            pass
