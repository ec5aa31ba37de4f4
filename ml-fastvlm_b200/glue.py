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


IMAGE_TOKEN_INDEX = -200      # llava/constants.py:8


def splice_layout(input_ids, attention_mask, num_tokens, max_length=None):
    """Host-side geometry of prepare_inputs_labels_for_multimodal (llava_arch.py:233-317) for tensor `images` (one feature
    block of `num_tokens` rows per <image> token, padding removed by attention_mask, right padding): returns
      lengths        new sequence length per sample,
      text_src/dst   flat index pairs: row text_dst of the [B*Lmax] output takes embed_tokens(ids_flat[text_src]),
      image_dst      [(image_index, sample, position)] for every <image> token, in consumption order
                     (a sample WITHOUT an <image> token still consumes one image index, llava_arch.py:239-246),
      Lmax.
    Pure function of integer tensors -- tested against the reference's own output (tests/test_splice.py)."""
    ids = input_ids.cpu()
    B = ids.shape[0]
    mask = torch.ones_like(ids, dtype=torch.bool) if attention_mask is None else attention_mask.cpu().bool()
    lengths, text_src, text_dst, image_dst = [], [], [], []
    img = 0
    rows = []
    for b in range(B):
        cols = torch.nonzero(mask[b], as_tuple=False).flatten().tolist()
        pos, items = 0, []
        had_image = False
        for c in cols:
            if int(ids[b, c]) == IMAGE_TOKEN_INDEX:
                items.append(("img", img, pos))
                img += 1
                pos += num_tokens
                had_image = True
            else:
                items.append(("txt", b * ids.shape[1] + c, pos))
                pos += 1
        if not had_image:
            img += 1
        if max_length is not None:
            pos = min(pos, max_length)
        lengths.append(pos)
        rows.append(items)
    Lmax = max(lengths) if lengths else 0
    for b, items in enumerate(rows):
        for kind, src, pos in items:
            if kind == "txt":
                if pos < lengths[b]:
                    text_src.append(src)
                    text_dst.append(b * Lmax + pos)
            elif pos + num_tokens <= lengths[b]:
                image_dst.append((src, b, pos))
            elif pos < lengths[b]:
                raise NotImplementedError("tokenizer_model_max_length truncates inside an image block")
    return dict(lengths=lengths, text_src=text_src, text_dst=text_dst, image_dst=image_dst, Lmax=Lmax, images_consumed=img)


def prepare_inputs_embeds(model, input_ids, attention_mask, images):
    """Row f2, general case: the `new_input_embeds` of prepare_inputs_labels_for_multimodal (llava_arch.py:146-332, tensor
    `images`, `flat` features) built WITHOUT materialising image features: text rows come from `embed_tokens`, and every
    <image> block is written by the projector GEMM's epilogue straight into its slot of the [B, Lmax, H] buffer
    (fvhd_forward_scatter).  Returns (inputs_embeds bf16, attention_mask bool, position_ids) like the reference
    (position_ids for right padding; None when attention_mask was None, llava_arch.py:323-330)."""
    tower = model.get_model().get_vision_tower()
    projector = model.get_model().mm_projector
    embed = model.get_model().embed_tokens
    with torch.no_grad():
        eng = tower.fused_engine(projector)
        dev = tower.device
        lay = splice_layout(input_ids, attention_mask, eng.num_tokens, getattr(getattr(model, "config", None), "tokenizer_model_max_length", None))
        B, Lmax, H = input_ids.shape[0], lay["Lmax"], eng.hidden
        out = torch.zeros(B, Lmax, H, dtype=torch.bfloat16, device=dev)
        if lay["text_src"]:
            src = torch.tensor(lay["text_src"], device=input_ids.device)
            tok = input_ids.reshape(-1)[src].to(embed.weight.device)
            rows = embed(tok).to(device=dev, dtype=torch.bfloat16)
            out.view(B * Lmax, H).index_copy_(0, torch.tensor(lay["text_dst"], device=dev), rows)
        if lay["image_dst"]:
            which = [i for i, _, _ in lay["image_dst"]]
            x = images[which].to(device=dev, dtype=tower.dtype)
            base = out.data_ptr()
            ptrs = [base + ((b * Lmax + pos) * H) * 2 for _, b, pos in lay["image_dst"]]
            eng.forward_scatter(x, ptrs)
        amask = torch.zeros(B, Lmax, dtype=torch.bool, device=dev)
        pos_ids = torch.zeros(B, Lmax, dtype=torch.long, device=dev)
        for b, n in enumerate(lay["lengths"]):
            amask[b, :n] = True
            pos_ids[b, :n] = torch.arange(n, device=dev)
        return out, (amask if attention_mask is not None else None), (pos_ids if attention_mask is not None else None)


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
