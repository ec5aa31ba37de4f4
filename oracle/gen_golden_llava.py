"""TEST INFRASTRUCTURE ONLY -- golden vectors of the reference's multimodal splice, from the UNMODIFIED reference tree.

Run in the build container (needs /root/reference):   python -m oracle.gen_golden_llava

Builds a tiny random-init `LlavaQwen2ForCausalLM` (llava/model/language_model/llava_qwen.py:37-56) with
mm_vision_tower = "mobileclip_l_256", mm_projector_type = "mlp2x_gelu" under oracle/timm_stub.py, loads the seeded fixture
into tower and projector by the reference's own key names, and records what
`prepare_inputs_labels_for_multimodal` (llava/model/llava_arch.py:146-332) returns for a ragged batch:
    sample 0: text(5) <image> text(7)                 one image
    sample 1: text(3) <image> text(2) <image> text(4) two images
    sample 2: text(9)                                 no image token (consumes one image's zero-length slice, :239-246)
Stored in tests/golden/llava_splice.npz: input_ids, the embed_tokens table, image seeds, `new_input_embeds`
[3, Lmax, H] (fp32, right-padded with zeros), attention_mask and position_ids -- what tests/test_splice.py requires
the B200 path (glue.prepare_inputs_embeds + fvhd_forward_scatter) to reproduce.
"""
import os
import sys

import numpy as np
import torch

from . import fixture as fx
from . import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
HID = 128            # tiny LLM hidden size (projector output)
VOCAB = 512
IMAGE_TOKEN_INDEX = -200     # llava/constants.py:8


def build_reference_model():
    ref_loader._prepare()
    from llava.model.language_model.llava_qwen import LlavaConfig, LlavaQwen2ForCausalLM
    cfg = LlavaConfig(hidden_size=HID, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                      vocab_size=VOCAB, max_position_embeddings=512)
    cfg.mm_vision_tower = "mobileclip_l_256"
    cfg.mm_projector_type = "mlp2x_gelu"
    cfg.mm_hidden_size = 3072
    cfg.unfreeze_mm_vision_tower = True      # materialise the tower weights despite delay_load (mobileclip_encoder.py:23-26)
    cfg.tokenizer_padding_side = "right"
    torch.manual_seed(0)
    model = LlavaQwen2ForCausalLM(cfg).eval()
    return model


def sample_inputs():
    g = torch.Generator().manual_seed(77)
    def txt(n):
        return torch.randint(1, VOCAB, (n,), generator=g)
    img = torch.tensor([IMAGE_TOKEN_INDEX])
    rows = [torch.cat([txt(5), img, txt(7)]), torch.cat([txt(3), img, txt(2), img, txt(4)]), txt(9)]
    L = max(r.numel() for r in rows)
    ids = torch.zeros(len(rows), L, dtype=torch.long)
    mask = torch.zeros(len(rows), L, dtype=torch.bool)
    for i, r in enumerate(rows):
        ids[i, :r.numel()] = r
        mask[i, :r.numel()] = True
    return ids, mask


def main():
    torch.set_grad_enabled(False)
    model = build_reference_model()
    sd = fx.tower_state_dict()
    psd = fx.projector_state_dict(HID)
    model.get_model().get_vision_tower().load_state_dict(sd, strict=True)
    model.get_model().mm_projector.load_state_dict(psd, strict=True)
    ids, mask = sample_inputs()
    images = fx.synthetic_images(4, 256, seed=55)         # 1 + 2 + 1 (the image-less sample still consumes one, llava_arch.py:239-246)
    out = model.prepare_inputs_labels_for_multimodal(ids, None, mask, None, None, images)
    _, position_ids, attention_mask, _, new_embeds, _ = out
    feats = model.encode_images(images)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "llava_splice.npz"),
                        input_ids=ids.numpy(), attention_mask_in=mask.numpy(),
                        embed_tokens=model.get_model().embed_tokens.weight.detach().numpy().astype(np.float32),
                        image_seed=np.array(55), n_images=np.array(4), hidden=np.array(HID),
                        new_input_embeds=new_embeds.numpy().astype(np.float32),
                        attention_mask=attention_mask.numpy(), image_features=feats.numpy().astype(np.float32))
    print("new_input_embeds", tuple(new_embeds.shape), "attention_mask sums", attention_mask.sum(1).tolist(), "position_ids", position_ids)
    # state-dict key names of the whole reference model that belong to the path (for the drop-in test)
    keys = [k for k in model.state_dict().keys() if "vision_tower" in k or "mm_projector" in k]
    with open(os.path.join(OUT, "llava_model_keys.txt"), "w") as f:
        f.write("\n".join(keys) + "\n")
    print(len(keys), "path keys, e.g.", keys[0], keys[-1])


if __name__ == "__main__":
    sys.exit(main())
