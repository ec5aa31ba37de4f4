"""Drop-in exercised through the REAL reference tree (llava_arch.py:29-41,141-144; llava_qwen.py:37-56).

Build container only (skipped where /root/reference is absent, i.e. on the GPU box): `patch_llava()` swaps the two factories
and encode_images inside the importable reference package; a tiny random-init LlavaQwen2ForCausalLM is then built BY THE
REFERENCE'S OWN constructors, the fixture is loaded by the reference's key names, and encode_images must dispatch to the
library (which, on a CPU-only host, fails loudly -- there is no fallback)."""
import os

import pytest
import torch

import ml_fastvlm_b200 as pkg
from oracle import fixture as fx
from oracle import ref_loader


def test_model_state_dict_keys_match_reference_model(golden_dir, tower_sd, proj_sd):
    """Runs anywhere: the key names a released checkpoint uses for the path (recorded from the reference model by
    oracle/gen_golden_llava.py) == `model.vision_tower.` + our tower keys, `model.mm_projector.` + our projector keys."""
    want = [l.strip() for l in open(os.path.join(golden_dir, "llava_model_keys.txt")) if l.strip()]

    class Args:
        mm_vision_tower = "mobileclip_l_256"
        unfreeze_mm_vision_tower = False
        mm_projector_type = "mlp2x_gelu"
        mm_hidden_size = 3072
        hidden_size = 128
    tower = pkg.build_vision_tower(Args())
    proj = pkg.build_vision_projector(Args())
    got = ["model.vision_tower." + k for k in tower.state_dict().keys()] + ["model.mm_projector." + k for k in proj.state_dict().keys()]
    assert sorted(got) == sorted(want)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
def test_patch_llava_builds_reference_model_with_b200_modules(tower_sd):
    ref_loader._prepare()
    llava_arch = pkg.patch_llava()
    from llava.model.language_model.llava_qwen import LlavaConfig, LlavaQwen2ForCausalLM
    cfg = LlavaConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                      vocab_size=512, max_position_embeddings=512)
    cfg.mm_vision_tower = "mobileclip_l_256"
    cfg.mm_projector_type = "mlp2x_gelu"
    cfg.mm_hidden_size = 3072
    cfg.unfreeze_mm_vision_tower = True          # materialise the tower despite delay_load (mobileclip_encoder.py:23-26)
    torch.manual_seed(0)
    model = LlavaQwen2ForCausalLM(cfg).eval()
    tower = model.get_model().get_vision_tower()
    proj = model.get_model().mm_projector
    assert isinstance(tower, pkg.FastViTHDVisionTower) and isinstance(proj, pkg.FastVLMProjector)
    # the reference's own key names load strictly
    sd = {"model.vision_tower." + k: v for k, v in tower_sd.items()}
    sd.update({"model.mm_projector." + k: v for k, v in fx.projector_state_dict(128).items()})
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all("vision_tower" not in k and "mm_projector" not in k for k in missing)
    assert tower.hidden_size == 3072 and tower.num_patches == 16 and tower.num_patches_per_side == 4
    assert llava_arch.LlavaMetaForCausalLM.encode_images is pkg.EncodeImagesMixin.encode_images
    x = fx.synthetic_images(2, 256)
    with torch.inference_mode(), pytest.raises(pkg.FvhdError):      # dispatches to the library: CPU host -> loud failure
        model.encode_images(x)
    if torch.cuda.is_available():                                   # (a GPU lease that ships a reference copy)
        model.to(device="cuda", dtype=torch.bfloat16)
        with torch.inference_mode():
            feats = model.encode_images(x.to("cuda"))
        assert tuple(feats.shape) == (2, 16, 128)
