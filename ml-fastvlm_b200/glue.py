"""encode_images -- the single call site of tower + projector (llava/model/llava_arch.py:141-144).

`encode_images(model, images)` keeps the reference semantics `mm_projector(vision_tower(images))`
but, when both modules are the B200 ones, issues ONE library call (tower + projector in one plan,
visual tokens never round-trip through PyTorch).  `patch_llava()` installs the three replacements
into an importable reference tree so `prepare_inputs_labels_for_multimodal` / `generate` /
a device-parameterised `predict.py` run unchanged.
"""
import torch

from .projector import FastVLMLinearProjector, FastVLMProjector, build_vision_projector
from .tower import FastViTHDVisionTower, build_vision_tower


def encode_images(model, images):
    tower = model.get_model().get_vision_tower()
    projector = model.get_model().mm_projector
    if (isinstance(tower, FastViTHDVisionTower) and isinstance(projector, (FastVLMProjector, FastVLMLinearProjector))
            and torch.is_tensor(images)):
        return tower.encode_with_projector(images, projector)
    image_features = tower(images)
    return projector(image_features)


def splice_visual_tokens(model, input_embeds, images, position):
    """Row f2: place the projected visual tokens of image b at input_embeds[b, position:position+N] without an
    intermediate tensor -- the projector GEMM's epilogue stores straight into the LLM embedding buffer
    (the single-image case of the splice loop in prepare_inputs_labels_for_multimodal, llava_arch.py:251-271)."""
    tower = model.get_model().get_vision_tower()
    projector = model.get_model().mm_projector
    with torch.no_grad():
        eng = tower.fused_engine(projector)
        x = images.to(device=tower.device, dtype=tower.dtype)
        if input_embeds.dtype == torch.bfloat16 and input_embeds.is_contiguous():
            return eng.forward_into(x, input_embeds, position)
        _, proj = eng.forward(x, want_tokens=False, want_projected=True)     # other dtypes: one cast-copy, as the reference does
        input_embeds[:, position:position + proj.shape[1]] = proj.to(input_embeds.dtype)
        return input_embeds


class EncodeImagesMixin:
    """Mix into a `LlavaMetaForCausalLM` subclass to override `encode_images` (llava_arch.py:141-144)."""

    def encode_images(self, images):
        return encode_images(self, images)


def patch_llava():
    """Swap the FastVLM tower / projector factories and encode_images inside an importable `llava` package."""
    import llava.model.llava_arch as llava_arch
    import llava.model.multimodal_encoder.builder as enc_builder
    import llava.model.multimodal_projector.builder as proj_builder

    enc_builder.build_vision_tower = build_vision_tower
    proj_builder.build_vision_projector = build_vision_projector
    llava_arch.build_vision_tower = build_vision_tower
    llava_arch.build_vision_projector = build_vision_projector
    llava_arch.LlavaMetaForCausalLM.encode_images = EncodeImagesMixin.encode_images
    return llava_arch
