"""FastViTHDVisionTower -- drop-in for `MobileCLIPVisionTower`
(llava/model/multimodal_encoder/mobileclip_encoder.py:13-116) backed by libfastvithd_b200.

Same constructor, same attributes read by the LLaVA scaffolding (`is_loaded`, `load_model`,
`image_processor`, `hidden_size`, `num_patches`, `num_patches_per_side`, `config`, `dtype`, `device`,
`dummy_feature`), same state-dict keys (`vision_tower.model.*`), same call contract
(`tower(images)`: tensor [B,3,R,R] or list of [3,R,R] -> [B,(R/64)^2,3072] in `images.dtype`).
The arithmetic is one C-ABI call into the sm_100a library; there is no PyTorch/CPU compute path.
"""
import torch
import torch.nn as nn

from . import arch
from .engine import Engine
from .lib import FvhdError
from .packer import pack_projector, pack_tower


def load_model_config(model_name):
    """mobileclip.load_model_config (mobileclip/__init__.py:16-31) for the one config the reference ships
    (mobileclip/configs/mobileclip_l.json)."""
    base = "_".join(model_name.split("_")[0:2])
    if base != "mobileclip_l":
        raise ValueError(f"Unsupported model name: {base}")
    return {
        "embed_dim": arch.PROJECTION_DIM,
        "image_cfg": {"image_size": 1024, "model_name": "fastvithd", "embed_dim": arch.EMBED_DIM, "patch_size": arch.PATCH_SIZE},
    }


class _Node(nn.Module):
    """Parameter container: reproduces the reference's module nesting so state-dict keys match."""


def _build_param_tree(root, specs):
    gen = torch.Generator().manual_seed(0)
    for key, (shape, dtype, is_buffer) in specs.items():
        parts = key.split(".")
        mod = root
        for p in parts[:-1]:
            if not hasattr(mod, p):
                mod.add_module(p, _Node())
            mod = getattr(mod, p)
        leaf = parts[-1]
        if dtype == torch.int64:
            t = torch.zeros(shape, dtype=dtype)
        elif "layer_scale" in key:
            t = torch.full(shape, 1e-5)                       # mci.py:1058,1132
        elif key.endswith(("bn.weight", "running_var", "norm.weight")):
            t = torch.ones(shape)
        elif key.endswith(("bias", "running_mean")):
            t = torch.zeros(shape)
        else:
            t = torch.randn(shape, generator=gen) * 0.02      # trunc_normal_/normal_(std=0.02) class of inits
        if is_buffer:
            mod.register_buffer(leaf, t)
        else:
            mod.register_parameter(leaf, nn.Parameter(t, requires_grad=False))


class FastViTHDVisionTower(nn.Module):
    def __init__(self, vision_tower, args, delay_load=False, max_batch=8):
        super().__init__()
        self.is_loaded = False
        self.vision_tower_name = vision_tower
        self.tune_vision_tower = getattr(args, "unfreeze_mm_vision_tower", False)
        self.input_image_size = int(vision_tower.split("_")[-1])          # mobileclip_encoder.py:20
        if self.input_image_size % arch.PATCH_SIZE:
            raise ValueError(f"image size {self.input_image_size} is not a multiple of {arch.PATCH_SIZE}")
        self.max_batch = max_batch
        self._engine = None            # tower-only plan
        self._fused = {}               # id(projector) -> (Engine with projector, projector version)
        self._version = 0
        if not delay_load or getattr(args, "unfreeze_mm_vision_tower", False):
            self.load_model()
        else:
            self.cfg_only = load_model_config(self.vision_tower_name)

    # ------------------------------------------------------------------ construction
    def load_model(self, device_map=None):
        if self.is_loaded:
            print("{} is already loaded, `load_model` called again, skipping.".format(self.vision_tower_name))
            return
        model_cfg = load_model_config(self.vision_tower_name)
        model_cfg["image_cfg"]["image_size"] = self.input_image_size
        self.cfg_only = model_cfg
        from transformers import CLIPImageProcessor
        sz = model_cfg["image_cfg"]["image_size"]
        self.image_processor = CLIPImageProcessor(crop_size={"height": sz, "width": sz}, image_mean=[0.0, 0.0, 0.0],
                                                  image_std=[1.0, 1.0, 1.0], size={"shortest_edge": sz})
        # parameter tree with the reference's names: self.vision_tower (MCi) . model (FastViT) . <k>
        _build_param_tree(self, arch.reference_param_specs())
        self.requires_grad_(False)
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())
        self.is_loaded = True

    def _invalidate(self):
        self._version += 1
        self._engine = None
        self._fused = {}

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._invalidate()
        return out

    def refresh_weights(self):
        """Re-pack after mutating parameters in place (the library holds its own folded copy)."""
        self._invalidate()

    # ------------------------------------------------------------------ engines
    def _check_runnable(self):
        if not self.is_loaded:
            raise FvhdError("vision tower weights are not loaded (delay_load); call load_model()")
        if self.tune_vision_tower and torch.is_grad_enabled():
            raise NotImplementedError("libfastvithd_b200 is inference-only: encoder backward (unfreeze_mm_vision_tower) is out of scope")
        if self.device.type != "cuda":
            raise FvhdError(f"FastViTHDVisionTower computes on CUDA (sm_100a) only; module is on {self.device}. No CPU fallback exists.")

    def engine(self):
        self._check_runnable()
        if self._engine is None:
            eng = Engine(self.input_image_size, 0, 2, self.max_batch)
            eng.load(pack_tower(self.state_dict()), self.device)
            self._engine = eng
        return self._engine

    def fused_engine(self, projector):
        """Tower + projector in one plan (one library call == encode_images)."""
        self._check_runnable()
        key = id(projector)
        ver = getattr(projector, "_version_counter", 0)
        hit = self._fused.get(key)
        if hit is None or hit[1] != ver:
            psd = projector.packed_state_dict()
            depth = len([k for k in psd if k.endswith(".weight")])
            hidden = psd["0.weight"].shape[0]
            eng = Engine(self.input_image_size, hidden, depth, self.max_batch)
            packed = pack_tower(self.state_dict())
            packed.update(pack_projector(psd))
            eng.load(packed, self.device)
            hit = (eng, ver)
            self._fused = {key: hit}
        return hit[0]

    # ------------------------------------------------------------------ forward (mobileclip_encoder.py:70-88)
    def feature_select(self, image_forward_outs):
        """The library already emits [B, HW, C]; kept for API parity (mobileclip_encoder.py:60-68)."""
        return image_forward_outs["image_embeddings"]

    def _refuse_training(self, projector=None):
        """The reference runs the tower under grad when `tune_vision_tower` is set (mobileclip_encoder.py:70-75).  This library has
        no backward: fail loudly BEFORE entering no_grad instead of silently returning detached features."""
        if not torch.is_grad_enabled():
            return
        if self.tune_vision_tower or any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("libfastvithd_b200 is inference-only: encoder backward (unfreeze_mm_vision_tower) is out of scope; "
                                      "call under torch.no_grad() / freeze the tower")
        if projector is not None and any(p.requires_grad for p in projector.parameters()):
            raise NotImplementedError("libfastvithd_b200 is inference-only: mm_projector training needs the reference projector; "
                                      "call under torch.no_grad() / freeze the projector")

    def forward(self, images):
        self._refuse_training()
        with torch.no_grad():
            return self.forward_images(images)

    def forward_images(self, images):
        eng = self.engine()
        if type(images) is list:
            # the reference loops one image at a time (mobileclip_encoder.py:78-83); images are independent through the tower,
            # so the list goes through ONE batched call and is split back into the reference's list of [1, N, C]
            if len(images) == 0:
                return []
            x = torch.stack([image.to(device=self.device, dtype=self.dtype) for image in images], 0)
            tokens, _ = eng.forward(x, want_tokens=True, want_projected=False)
            return [tokens[i:i + 1].to(image.dtype) for i, image in enumerate(images)]
        x = images.to(device=self.device, dtype=self.dtype)
        tokens, _ = eng.forward(x, want_tokens=True, want_projected=False)
        return tokens.to(images.dtype)

    def encode_with_projector(self, images, projector):
        """mm_projector(tower(images)) as ONE call (llava_arch.py:141-144)."""
        self._refuse_training(projector)
        with torch.no_grad():
            eng = self.fused_engine(projector)
            x = images.to(device=self.device, dtype=self.dtype)
            _, proj = eng.forward(x, want_tokens=False, want_projected=True)
            return proj.to(images.dtype)

    # ------------------------------------------------------------------ attributes (mobileclip_encoder.py:90-116)
    @property
    def dummy_feature(self):
        return torch.zeros(1, self.hidden_size, device=self.device, dtype=self.dtype)

    @property
    def dtype(self):
        return next(self.vision_tower.parameters()).dtype

    @property
    def device(self):
        return next(self.vision_tower.parameters()).device

    @property
    def config(self):
        return self.cfg_only

    @property
    def hidden_size(self):
        return self.config["image_cfg"]["embed_dim"]

    @property
    def num_patches_per_side(self):
        return self.config["image_cfg"]["image_size"] // self.config["image_cfg"]["patch_size"]

    @property
    def num_patches(self):
        return (self.config["image_cfg"]["image_size"] // self.config["image_cfg"]["patch_size"]) ** 2


def build_vision_tower(vision_tower_cfg, **kwargs):
    """multimodal_encoder/builder.py:6-19 for the FastVLM tower: dispatch on the `mm_vision_tower` string.
    CLIP / CLIP-S2 towers are not FastVLM's encoder and are out of scope -> ValueError like the reference's
    fall-through (builder.py:19)."""
    vision_tower = getattr(vision_tower_cfg, "mm_vision_tower", getattr(vision_tower_cfg, "vision_tower", None))
    if vision_tower is not None and "mobileclip" in vision_tower.lower():
        return FastViTHDVisionTower(vision_tower, args=vision_tower_cfg, **kwargs)
    raise ValueError(f"Unknown vision tower: {vision_tower}")
