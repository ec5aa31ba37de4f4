"""Row f1: preprocessing parity.  CPU: the oracle restatement of Pillow's resampler is pinned bit-for-bit against PIL itself,
and the library's host-side coefficient tables against the oracle's.  GPU: fvhd_preprocess == oracle, bit-exact."""
import numpy as np
import pytest
import torch
from PIL import Image

import ml_fastvlm_b200 as pkg
from oracle import preprocess_oracle as po

CASES = [(256, 256, 1024), (37, 53, 64), (480, 640, 256), (1500, 1000, 256), (300, 200, 1024), (64, 64, 64), (333, 517, 192), (719, 1280, 512)]


def _img(h, w, seed):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("h,w,res", CASES)
def test_oracle_resize_is_bit_exact_with_pil(h, w, res):
    img = _img(h, w, h * 7 + w)
    oh, ow = po.resize_output_size(h, w, res)
    ref = np.asarray(Image.fromarray(img).resize((ow, oh), resample=Image.BICUBIC))
    assert np.array_equal(po.resize_bicubic_u8(img, ow, oh), ref)


def test_oracle_pipeline_matches_pil_pipeline():
    """resize -> centre crop -> x 1/255 (float64 multiply, float32 result) -> CHW, and the 'pad' (expand2square) variant."""
    img = _img(200, 320, 3)
    res = 128
    for pad in (False, True):
        pil = Image.fromarray(img)
        if pad:
            sq = Image.new("RGB", (320, 320), (0, 0, 0))
            sq.paste(pil, (0, (320 - 200) // 2))
            pil = sq
        w, h = pil.size
        oh, ow = po.resize_output_size(h, w, res)
        r = np.asarray(pil.resize((ow, oh), resample=Image.BICUBIC))
        top, left = (oh - res) // 2, (ow - res) // 2
        r = r[top:top + res, left:left + res]
        want = (r.astype(np.float64) * (1 / 255)).astype(np.float32).transpose(2, 0, 1)
        assert np.array_equal(po.preprocess(img, res, pad=pad), want)


@pytest.mark.parametrize("n_in,n_out", [(256, 1024), (1000, 256), (53, 91), (64, 64), (3000, 768), (17, 5)])
def test_library_coefficient_tables_match_oracle(n_in, n_out):
    b, k = pkg.resample_coeffs(n_in, n_out)
    ob, ok = po.precompute_coeffs(n_in, n_out)
    assert k.shape == ok.shape and np.array_equal(b, ob) and np.array_equal(k, ok)
    assert int(np.abs(k.sum(1) - (1 << 22)).max()) <= k.shape[1]          # rows are normalised in 22-bit fixed point


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,res", [(256, 256, 1024), (480, 640, 256), (1500, 1000, 256), (300, 200, 1024), (719, 1280, 512)])
def test_gpu_preprocess_bit_exact(h, w, res):
    dev = torch.device("cuda:0")
    eng = pkg.Engine(res, 0, 2, 1)
    eng.device = dev                                       # preprocessing needs no weights
    img = _img(h, w, h + w)
    for pad in (False, True):
        want = torch.from_numpy(po.preprocess(img, res, pad=pad))
        out32 = torch.empty(3, res, res, dtype=torch.float32, device=dev)
        pkg.preprocess_into(eng, img, out32, pad=pad)
        assert torch.equal(out32.cpu(), want)                                          # uint8 pipeline + LUT: exact
        out16 = torch.empty(3, res, res, dtype=torch.float16, device=dev)
        pkg.preprocess_into(eng, torch.from_numpy(img).to(dev), out16, pad=pad)        # device-resident source
        assert torch.equal(out16.cpu(), want.half())
        outbf = torch.empty(3, res, res, dtype=torch.bfloat16, device=dev)
        pkg.preprocess_into(eng, Image.fromarray(img), outbf, pad=pad)                 # PIL input
        assert torch.equal(outbf.cpu(), want.bfloat16())


@pytest.mark.gpu
def test_gpu_process_images_feeds_the_tower(tower_sd):
    """config 1 flow: a 256x256 uint8 image through process_images (upsampled to the tower's R) and the tower."""
    dev = torch.device("cuda:0")

    class Args:
        mm_vision_tower = "mobileclip_l_256"
        unfreeze_mm_vision_tower = False
    tower = pkg.build_vision_tower(Args())
    tower.load_state_dict(tower_sd, strict=True)
    tower.to(device=dev, dtype=torch.float16)
    imgs = [_img(300, 400, 1), _img(256, 256, 2)]
    x = pkg.process_images(imgs, tower)
    assert tuple(x.shape) == (2, 3, 256, 256) and x.dtype == torch.float16 and x.is_cuda
    want = torch.stack([torch.from_numpy(po.preprocess(im, 256)) for im in imgs]).half()
    assert torch.equal(x.cpu(), want)
    feats = tower(x)
    assert tuple(feats.shape) == (2, 16, 3072) and torch.isfinite(feats.float()).all()


# ------------------------------------------------------------------ image_aspect_ratio == "anyres" (mm_utils.py:14-147)
def _anyres_cases():
    import json
    import os
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    cases = json.load(open(os.path.join(gdir, "anyres_cases.json")))
    gold = np.load(os.path.join(gdir, "anyres_ref.npz"))
    return [(h, w, R, [tuple(p) for p in pins], seed, gold[f"case{i}"], tuple(int(v) for v in gold[f"grid{i}"]))
            for i, (h, w, R, pins, seed) in enumerate(cases)]


def test_anyres_oracle_matches_reference_golden():
    """tests/golden/anyres_ref.npz was produced by the unmodified reference (oracle/gen_golden_anyres.py): global view + tiles,
    including a canvas that is not a multiple of the patch size.  The numpy restatement must reproduce it bit for bit, and the
    host-side geometry helpers must agree with the reference's get_anyres_image_grid_shape."""
    for h, w, R, pins, seed, k, grid in _anyres_cases():
        img = np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
        got = po.preprocess_anyres(img, R, pins)
        assert got.shape == k.shape and got.dtype == np.float32
        assert np.array_equal(got, po.RESCALE_LUT[k])
        assert pkg.get_anyres_image_grid_shape((w, h), pins, R) == grid == pkg.get_anyres_image_grid_shape((w, h), str(pins), R)
        assert pkg.select_best_resolution((w, h), pins) == po.select_best_resolution((w, h), pins)


def test_anyres_host_geometry_matches_oracle_on_random_sizes():
    """The host arithmetic that parameterises fvhd_preprocess_tiles (resize target, paste offset, tile grid) against the oracle's
    restatement of resize_and_pad_image / divide_to_patches, over random image sizes and pin sets (incl. non-multiples of R)."""
    rng = np.random.default_rng(7)
    for _ in range(500):
        R = int(rng.choice([64, 128, 336, 1024]))
        w, h = int(rng.integers(8, 4000)), int(rng.integers(8, 4000))
        pins = [(int(R * a), int(R * b)) for a, b in rng.choice([1, 2, 3, 1.5], size=(int(rng.integers(1, 6)), 2))]
        cw, ch, nw, nh, px, py = po.anyres_geometry(w, h, pins)
        got = pkg.anyres_geometry(w, h, R, pins)
        assert got == (nh, nw, py, px, -(-ch // R), -(-cw // R)), (w, h, R, pins)
        assert 1 <= nh <= ch and 1 <= nw <= cw and py >= 0 and px >= 0


@pytest.mark.gpu
def test_gpu_anyres_bit_exact():
    dev = torch.device("cuda:0")
    for h, w, R, pins, seed, k, grid in _anyres_cases():
        eng = pkg.Engine(R, 0, 2, 1)
        eng.device = dev                                   # preprocessing needs no weights
        img = np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = torch.from_numpy(po.RESCALE_LUT[k])         # the reference's own output
        out = pkg.process_anyres_image(img, eng, pins, torch.float32)
        assert tuple(out.shape) == tuple(want.shape) and torch.equal(out.cpu(), want)
        out16 = pkg.process_anyres_image(torch.from_numpy(img).to(dev), eng, str(pins), torch.float16)      # device source, str pins
        assert torch.equal(out16.cpu(), want.half())

    class Cfg:
        image_aspect_ratio = "anyres"
        image_grid_pinpoints = [(64, 128), (128, 64), (128, 128)]
    eng = pkg.Engine(64, 0, 2, 1)
    eng.device = dev
    imgs = [np.random.default_rng(s).integers(0, 256, (100, 150, 3), dtype=np.uint8) for s in (1, 2)]
    x = pkg.process_images(imgs, eng, Cfg(), dtype=torch.float32)
    wants = [torch.from_numpy(po.preprocess_anyres(im, 64, Cfg.image_grid_pinpoints)) for im in imgs]
    assert tuple(x.shape) == (2,) + tuple(wants[0].shape) == (2, 5, 3, 64, 64)     # same tile count -> stacked, as the reference does
    for i in range(2):
        assert torch.equal(x[i].cpu(), wants[i])
