"""TEST INFRASTRUCTURE ONLY -- run the UNMODIFIED reference (apple/ml-fastvlm) in this container.

Only usable where /root/reference exists (the build container, not the GPU box).  Used by
oracle/gen_golden.py to pin oracle/fastvithd_oracle.py and to produce tests/golden/*.
Nothing under tests -m gpu, smoke() or bench.py imports this module.
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("FVHD_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "llava"))


def _prepare():
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    from . import timm_stub
    timm_stub.install()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


class _Args:
    unfreeze_mm_vision_tower = False


def reference_tower(res, state_dict=None):
    """MobileCLIPVisionTower("mobileclip_l_<res>") (mobileclip_encoder.py:13-58), fp32, eval."""
    _prepare()
    from llava.model.multimodal_encoder.builder import build_vision_tower

    class Cfg(_Args):
        mm_vision_tower = f"mobileclip_l_{res}"
    tower = build_vision_tower(Cfg(), delay_load=False)
    if state_dict is not None:
        tower.load_state_dict(state_dict, strict=True)
    return tower.eval()


def reference_projector(hidden, mm_hidden=3072, ptype="mlp2x_gelu", state_dict=None):
    """build_vision_projector (multimodal_projector/builder.py:17-35)."""
    _prepare()
    from llava.model.multimodal_projector.builder import build_vision_projector

    class Cfg:
        mm_projector_type = ptype
        mm_hidden_size = mm_hidden
        hidden_size = hidden
    proj = build_vision_projector(Cfg())
    if state_dict is not None:
        proj.load_state_dict(state_dict, strict=True)
    return proj.eval()
