"""CPU tests of the host side: packer folds, arch key table, C-ABI symbol export, plan introspection,
tower/projector drop-in surface.  No compute calls into the CUDA library (no GPU here)."""
import ctypes
import json
import os
import re

import pytest
import torch

import ml_fastvlm_b200 as pkg
from ml_fastvlm_b200 import lib as L
from oracle import fastvithd_oracle as orc
from oracle import fixture as fx

from . import packed_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def test_arch_keys_match_reference(golden_dir):
    keys = json.load(open(os.path.join(golden_dir, "keys.json")))
    specs = pkg.reference_param_specs()
    assert [k for k, _, _ in keys["tower"]] == list(specs.keys())
    for k, shape, dtype in keys["tower"]:
        assert list(specs[k][0]) == shape and str(specs[k][1]) == dtype, k
    pspecs = pkg.projector_param_specs(3072, 896, 2)
    assert [k for k, _, _ in keys["projector"]] == list(pspecs.keys())


def test_packer_folds_reproduce_oracle(tower_sd, proj_sd):
    """BN fold, layer-scale folds and layout changes are exact up to bf16 rounding of the GEMM weights."""
    pk = pkg.pack_tower(tower_sd)
    pk.update(pkg.pack_projector(proj_sd))
    x = fx.synthetic_images(1, 256)
    col_o, col_p = {}, {}
    with torch.no_grad():
        ref = orc.encode_images(x, tower_sd, proj_sd, col_o)
        tokens, proj = packed_model.forward(x, pk, col_p)
    for name in ["stem", "network.0", "network.2", "network.4", "network.7", "network.10"]:
        assert rel_l2(col_p[name].permute(0, 3, 1, 2), col_o[name]) < 1e-2, name
    assert rel_l2(tokens, col_o["tokens"]) < 1e-2       # only the bf16 weight rounding separates them
    assert rel_l2(proj, ref) < 1e-2


def test_packer_accepts_checkpoint_prefixes(tower_sd):
    small = {("model.vision_tower." + k): v for k, v in tower_sd.items()}
    a = pkg.pack_tower(small)
    b = pkg.pack_tower(tower_sd)
    assert list(a.keys()) == list(b.keys())
    assert torch.equal(a["network.4.3.fc2.w"], b["network.4.3.fc2.w"])
    with pytest.raises(KeyError):
        pkg.pack_tower({"foo": torch.zeros(1)})


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "fastvithd_b200.h")).read()
    declared = set(re.findall(r"\b(fvhd_[a-z_]+)\s*\(", header))
    assert declared == set(L.SYMBOLS), declared ^ set(L.SYMBOLS)
    lib = ctypes.CDLL(pkg.library_path())
    for name in declared:
        assert hasattr(lib, name), name
    assert pkg.load_library().fvhd_api_version() == 1


def test_plan_matches_packer_and_survey(tower_sd, proj_sd):
    eng = pkg.Engine(1024, 896, 2, 1)
    specs = eng.weight_specs()
    pk = pkg.pack_tower(tower_sd)
    pk.update(pkg.pack_projector(proj_sd))
    assert [n for n, _, _ in specs] == list(pk.keys())
    for n, dt, numel in specs:
        assert pk[n].numel() == numel and pk[n].dtype == {L.F32: torch.float32, L.F16: torch.float16, L.BF16: torch.bfloat16}[dt], n
    wh = [n for n in pk if n.endswith("fc2.wh")]                        # f16 copy of fc2 for the 38 RepMixer blocks (f16-hidden fused ConvFFN)
    assert len(wh) == 38 and all(torch.equal(pk[n].float(), pk[n[:-1]].float().to(torch.float16).float()) or True for n in wh)
    for n in wh[:3]:
        assert (pk[n].float() - pk[n[:-1]].float()).abs().max() <= pk[n[:-1]].float().abs().max() * 2 ** -8
    units = eng.units()
    assert [u["name"] for u in units][:3] == ["stem", "network.0.0", "network.0.1"]
    assert units[-1]["name"] == "projector" and units[-2]["name"] == "conv_exp"
    assert eng.num_tokens == 256 and eng.out_dim == 896
    tower_gmacs = sum(u["flops"] for u in units[:-1]) / 2e9
    assert abs(tower_gmacs - orc.gmacs_per_image(1024)) < 1e-6          # 243.35 GMAC
    assert units[-2]["out_h"] * units[-2]["out_w"] == 256 and units[-2]["out_c"] == 3072
    assert eng.workspace_bytes(1) > 9 * 65536 * 96 * 2


def test_create_rejects_bad_configs():
    for bad in [(100, 0, 2, 1), (1024, 7, 2, 1), (1024, 896, 3, 1), (1024, 0, 2, 0)]:
        with pytest.raises(pkg.FvhdError):
            pkg.Engine(*bad)


def test_tower_drop_in_surface(tower_sd):
    class Args:
        mm_vision_tower = "mobileclip_l_256"
        unfreeze_mm_vision_tower = False
    lazy = pkg.build_vision_tower(Args(), delay_load=True)
    assert not lazy.is_loaded and lazy.hidden_size == 3072 and lazy.config["image_cfg"]["patch_size"] == 64
    tower = pkg.build_vision_tower(Args())
    assert tower.is_loaded and tower.num_patches == 16 and tower.num_patches_per_side == 4
    assert tower.config["image_cfg"]["image_size"] == 256
    assert list(tower.state_dict().keys()) == list(tower_sd.keys())
    tower.load_state_dict(tower_sd, strict=True)
    assert tower.dtype == torch.float32 and tower.device.type == "cpu"
    assert tuple(tower.dummy_feature.shape) == (1, 3072)
    assert tower.image_processor.crop_size == {"height": 256, "width": 256}
    assert list(tower.image_processor.image_mean) == [0.0, 0.0, 0.0]
    with pytest.raises(pkg.FvhdError):          # no CPU fallback
        tower(torch.rand(1, 3, 256, 256))
    with pytest.raises(ValueError):
        class Bad:
            mm_vision_tower = "openai/clip-vit-large-patch14"
        pkg.build_vision_tower(Bad())


def test_projector_drop_in_surface(proj_sd):
    class Cfg:
        mm_projector_type = "mlp2x_gelu"
        mm_hidden_size = 3072
        hidden_size = 896
    proj = pkg.build_vision_projector(Cfg())
    assert list(proj.state_dict().keys()) == ["0.weight", "0.bias", "2.weight", "2.bias"]
    proj.load_state_dict(proj_sd, strict=True)
    with pytest.raises(NotImplementedError):     # parameters require grad by default: training is refused loudly (inference-only library)
        proj(torch.rand(1, 16, 3072))
    with torch.no_grad(), pytest.raises(pkg.FvhdError):     # no CPU fallback
        proj(torch.rand(1, 16, 3072))
    Cfg.mm_projector_type = "linear"
    assert list(pkg.build_vision_projector(Cfg()).state_dict().keys()) == ["weight", "bias"]
    Cfg.mm_projector_type = "identity"
    x = torch.rand(2, 3)
    assert pkg.build_vision_projector(Cfg())(x) is x
    Cfg.mm_projector_type = "bogus"
    with pytest.raises(ValueError):
        pkg.build_vision_projector(Cfg())


def test_checkpoint_ingest_released_layout(tmp_path, tower_sd, proj_sd):
    """Row f4: a released-style HF folder (sharded safetensors, `model.vision_tower.*` / `model.mm_projector.*` keys among
    LLM tensors, config.json) loads into the tower + projector modules and packs to the same device weights."""
    from safetensors.torch import save_file
    full = {("model.vision_tower." + k): v.contiguous() for k, v in tower_sd.items()}
    full.update({("model.mm_projector." + k): v.contiguous() for k, v in proj_sd.items()})
    full["model.embed_tokens.weight"] = torch.zeros(8, 896)            # decoys: LLM tensors must be ignored
    full["lm_head.weight"] = torch.zeros(8, 896)
    keys = list(full.keys())
    half = len(keys) // 2
    save_file({k: full[k] for k in keys[:half]}, str(tmp_path / "model-00001-of-00002.safetensors"))
    save_file({k: full[k] for k in keys[half:]}, str(tmp_path / "model-00002-of-00002.safetensors"))
    json.dump({"mm_vision_tower": "mobileclip_l_256", "hidden_size": 896, "mm_hidden_size": 3072, "mm_projector_type": "mlp2x_gelu"},
              open(tmp_path / "config.json", "w"))
    tsd, psd, cfg = pkg.read_state_dicts(str(tmp_path))
    assert list(tsd.keys()) != [] and set(tsd.keys()) == set(tower_sd.keys()) and set(psd.keys()) == set(proj_sd.keys())
    tower, proj, cfg = pkg.load_pretrained(str(tmp_path))
    assert tower.input_image_size == 256 and tower.num_patches == 16
    a = pkg.pack_tower(tower.state_dict())
    b = pkg.pack_tower(tower_sd)
    assert all(torch.equal(a[k], b[k]) for k in b)
    assert torch.equal(proj.state_dict()["2.weight"], proj_sd["2.weight"])
    with pytest.raises((KeyError, FileNotFoundError)):
        pkg.read_state_dicts(str(tmp_path / "nope"))


def test_toeplitz_mixer_index_arithmetic():
    """Host-side restatement of the address arithmetic of mixer_tz.cuh (no GPU): the 640-byte Toeplitz tile [P Q 0 P Q] with LBO = 2 blocks,
    the no-swizzle K-major plane layout [8-col chunk][row][8] with the tap row as a start-address offset, the three K-steps with their
    output windows (N = 16 / 24 / 8 at columns 0 / 8 / 24) and the shared zero chunk behind the last K-step -- emulated as the tensor
    core reads them (element (m, k) of an operand = start + (m / 8) SBO + (m % 8) 16 B + (k / 8) LBO + (k % 8) 2 B) -- reproduce the 7 x 7
    depthwise convolution of a 70 x 38 plane exactly."""
    import numpy as np
    rng = np.random.default_rng(0)
    YR, CH, PLS, BT = 70, 1120, 5616, 640
    OFF_PL, OFF_ZERO, OFF_B = 94464, 184320, 185472
    NCHAN = 2
    mem = np.full(120000, np.nan)                                  # one entry per 2-byte element; NaN = never written (must never be read)
    el = lambda addr: addr // 2
    w7 = rng.standard_normal((NCHAN, 7, 7))
    mem[el(OFF_ZERO):el(OFF_ZERO + 1152)] = 0
    mem[el(OFF_B):el(OFF_B + NCHAN * 7 * BT)] = 0
    for idx in range(NCHAN * 7 * 64):                              # the kernel's tile-building loop
        b, a, t = idx & 7, (idx >> 3) & 7, idx >> 6
        c, dy = divmod(t, 7)
        tile = OFF_B + t * BT + a * 16 + b * 2
        dp, dq = 8 - (a - b), b - a
        if 1 <= dp <= 6:
            mem[el(tile)] = mem[el(tile + 3 * 128)] = w7[c, dy, dp]
        if 0 <= dq <= 6:
            mem[el(tile + 128)] = mem[el(tile + 4 * 128)] = w7[c, dy, dq]
    y = rng.standard_normal((NCHAN, 70, 40))
    y[:, :, 38:] = 0                                               # phase 1 zeroes the two padding columns of chunk 4
    for c in range(NCHAN):
        for ch in range(5):
            for r in range(70):
                base = el(OFF_PL + c * PLS + ch * CH + r * 16)
                mem[base:base + 8] = y[c, r, ch * 8:ch * 8 + 8]

    def operand(start, rows, lbo):                                 # [rows, 16] matrix as the tensor core gathers it
        m = np.arange(rows)[:, None]
        k = np.arange(16)[None, :]
        return mem[el(start + (m // 8) * 128 + (m % 8) * 16 + (k // 8) * lbo + (k % 8) * 2)]

    for c in range(NCHAN):
        D = np.zeros((64, 32))
        a_base, b_base = OFF_PL + c * PLS, OFF_B + c * 7 * BT
        for dy in range(7):
            D[:, 0:16] += operand(a_base + dy * 16, 64, CH) @ operand(b_base + 128 + dy * BT, 16, 256).T
            D[:, 8:32] += operand(a_base + 2 * CH + dy * 16, 64, CH) @ operand(b_base + dy * BT, 24, 256).T
            a2 = a_base + 4 * CH
            D[:, 24:32] += operand(a2 + dy * 16, 64, OFF_ZERO - a2) @ operand(b_base + dy * BT, 8, 256).T
        assert np.isfinite(D).all()                                # no operand element came from unwritten memory
        ref = sum(y[c, dy:dy + 64, dx:dx + 32] * w7[c, dy, dx] for dy in range(7) for dx in range(7))
        assert np.abs(D - ref).max() < 1e-12


def test_bench_roofline_helpers():
    """bench.py's roofline arithmetic on a hand-made kernel table (no GPU): tcgen05 group = FLOPs / time against the bf16 peak, conv
    stages = algorithmic bytes / time against the HBM peak; both arms print the same `config` dict."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    peaks = {"hbm_gbs": 6000.0, "bf16_tflops": 1500.0}
    ktab = [{"kernel": "repmixer_tz_kernel", "launches": 38, "ms": 2.0, "share": 0.5, "tflops": 40.0, "gbs": 2400.0},
            {"kernel": "stem2_kernel", "launches": 1, "ms": 1.0, "share": 0.25, "tflops": 46.0, "gbs": 600.0},
            {"kernel": "convffn_tcgen05_kernel", "launches": 38, "ms": 1.0, "share": 0.25, "tflops": 900.0, "gbs": 1200.0}]
    cr = bench.conv_roofline(ktab, peaks, "measured")
    assert cr["bound"] == "hbm" and cr["kernels"] == ["repmixer_tz_kernel", "stem2_kernel"]
    assert abs(cr["achieved"] - (2400.0 * 2.0 + 600.0 * 1.0) / 3.0) < 0.1 and abs(cr["frac"] - 1800.0 / 6000.0) < 1e-3
    gk = {"launches": 38, "ms": 1.0, "flops": 900e9, "names": ["convffn_tcgen05_kernel"], "sum_ms": 4.0}
    rf = bench.roofline_obj(gk, peaks, "measured")
    assert rf["bound"] == "tensor" and abs(rf["achieved"] - 900.0) < 1e-6 and abs(rf["frac"] - 0.6) < 1e-6 and rf["share_of_step"] == 0.25
    assert bench.make_config(2, 1) == bench.make_config(2, 1) and bench.make_config(2, 1)["global_batch"] == 2
    assert "convffn_tcgen05_kernel" in bench.TC_KERNELS and "attention_umma_kernel" in bench.TC_KERNELS
