"""Row f2 against the reference's OWN prepare_inputs_labels_for_multimodal (llava_arch.py:146-332): golden vectors generated
by oracle/gen_golden_llava.py from the unmodified reference tree (tiny random-init LlavaQwen2ForCausalLM, fixture tower +
projector, ragged batch with one / two / zero <image> tokens)."""
import os

import numpy as np
import pytest
import torch

import ml_fastvlm_b200 as pkg
from oracle import fixture as fx


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "llava_splice.npz"))


def test_splice_layout_matches_reference_geometry(gold):
    """CPU: lengths, padding, text rows and image slots of the host layout function == the reference output."""
    ids = torch.from_numpy(gold["input_ids"])
    mask = torch.from_numpy(gold["attention_mask_in"])
    ref = torch.from_numpy(gold["new_input_embeds"])
    table = torch.from_numpy(gold["embed_tokens"])
    feats = torch.from_numpy(gold["image_features"])
    N = feats.shape[1]
    lay = pkg.splice_layout(ids, mask, N)
    assert lay["Lmax"] == ref.shape[1]
    assert lay["lengths"] == torch.from_numpy(gold["attention_mask"]).sum(1).tolist() == [28, 41, 9]
    assert lay["images_consumed"] == int(gold["n_images"])
    assert [(i, b) for i, b, _ in lay["image_dst"]] == [(0, 0), (1, 1), (2, 1)]      # image 3 is consumed by the image-less sample
    out = torch.zeros(ref.shape[0] * ref.shape[1], ref.shape[2])
    out[torch.tensor(lay["text_dst"])] = table[ids.reshape(-1)[torch.tensor(lay["text_src"])]]
    out = out.view_as(ref)
    for i, b, pos in lay["image_dst"]:
        out[b, pos:pos + N] = feats[i]
    assert torch.equal(out, ref)                                                      # bit-exact: pure data movement


@pytest.mark.gpu
def test_prepare_inputs_embeds_vs_reference_golden(gold, tower_sd):
    """GPU: text rows bit-equal to bf16(embed_tokens), image blocks written by the projector epilogue (fvhd_forward_scatter)
    within the bf16 tolerance of the reference's fp32 features, padding rows zero."""
    dev = torch.device("cuda:0")
    H = int(gold["hidden"])

    class Args:
        mm_vision_tower = "mobileclip_l_256"
        unfreeze_mm_vision_tower = False
        mm_projector_type = "mlp2x_gelu"
        mm_hidden_size = 3072
        hidden_size = H
    tower = pkg.build_vision_tower(Args())
    tower.load_state_dict(tower_sd, strict=True)
    proj = pkg.build_vision_projector(Args())
    proj.load_state_dict(fx.projector_state_dict(H), strict=True)
    tower.to(device=dev, dtype=torch.bfloat16)
    proj.to(device=dev, dtype=torch.bfloat16)
    embed = torch.nn.Embedding.from_pretrained(torch.from_numpy(gold["embed_tokens"])).to(dev)

    class Inner:
        mm_projector = proj
        embed_tokens = embed
        def get_vision_tower(self):
            return tower

    class Model:
        config = None
        def get_model(self):
            return Inner()
    ids = torch.from_numpy(gold["input_ids"]).to(dev)
    mask = torch.from_numpy(gold["attention_mask_in"]).to(dev)
    images = fx.synthetic_images(int(gold["n_images"]), 256, seed=int(gold["image_seed"])).to(dev)
    out, amask, pos_ids = pkg.prepare_inputs_embeds(Model(), ids, mask, images)
    ref = torch.from_numpy(gold["new_input_embeds"])
    assert tuple(out.shape) == tuple(ref.shape) and out.dtype == torch.bfloat16
    assert torch.equal(amask.cpu(), torch.from_numpy(gold["attention_mask"]))
    lay = pkg.splice_layout(ids, mask, tower.num_patches)
    img_rows = torch.zeros(ref.shape[0], ref.shape[1], dtype=torch.bool)
    for _, b, pos in lay["image_dst"]:
        img_rows[b, pos:pos + tower.num_patches] = True
    o = out.float().cpu()
    assert torch.equal(o[~img_rows], ref.to(torch.bfloat16).float()[~img_rows])       # text + padding rows: exact
    err = ((o[img_rows] - ref[img_rows]).norm() / ref[img_rows].norm()).item()
    assert err < 5e-2, err
    for n, (b, row) in enumerate(zip(range(3), pos_ids.cpu())):
        L = lay["lengths"][b]
        assert row[:L].tolist() == list(range(L)) and (row[L:] == 0).all()
