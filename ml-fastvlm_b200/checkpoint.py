"""Checkpoint ingest for released FastVLM weights (row f4).

A FastVLM checkpoint directory (get_models.sh:8-13) is an HF folder: `config.json` (`mm_vision_tower`, `hidden_size`,
`mm_projector_type`, ...) plus `model*.safetensors` / `pytorch_model*.bin` shards whose tower / projector tensors are keyed
`model.vision_tower.vision_tower.model.<k>` and `model.mm_projector.{0,2}.{weight,bias}` (loaded by
llava/model/builder.py:131 through `from_pretrained`).  This module reads only those tensors -- the LLM weights stay with
whatever loads the language model -- and returns modules of this package with the weights in place.
"""
import glob
import json
import os
from collections import OrderedDict

import torch

TOWER_KEY = "model.vision_tower."
PROJ_KEY = "model.mm_projector."


def _iter_shards(path):
    st = sorted(glob.glob(os.path.join(path, "*.safetensors")))
    if st:
        from safetensors import safe_open
        for f in st:
            with safe_open(f, framework="pt", device="cpu") as sf:
                for k in sf.keys():
                    if k.startswith(TOWER_KEY) or k.startswith(PROJ_KEY):
                        yield k, sf.get_tensor(k)
        return
    bins = sorted(glob.glob(os.path.join(path, "pytorch_model*.bin"))) + sorted(glob.glob(os.path.join(path, "mm_projector.bin")))
    if not bins:
        raise FileNotFoundError(f"no *.safetensors / pytorch_model*.bin under {path}")
    for f in bins:
        sd = torch.load(f, map_location="cpu", weights_only=True)
        for k, v in sd.items():
            if k.startswith(TOWER_KEY) or k.startswith(PROJ_KEY):
                yield k, v


def read_state_dicts(path):
    """-> (tower_sd keyed `vision_tower.model.*`, projector_sd keyed `0.weight` ..., config dict)."""
    cfg_path = os.path.join(path, "config.json")
    config = json.load(open(cfg_path)) if os.path.exists(cfg_path) else {}
    tower, proj = OrderedDict(), OrderedDict()
    for k, v in _iter_shards(path):
        if k.startswith(TOWER_KEY):
            tower[k[len(TOWER_KEY):]] = v
        else:
            proj[k[len(PROJ_KEY):]] = v
    if not tower:
        raise KeyError(f"{path}: no '{TOWER_KEY}*' tensors (was the tower saved? unfreeze_mm_vision_tower checkpoints carry it)")
    return tower, proj, config


def load_pretrained(path, device=None, dtype=None, max_batch=8):
    """Build FastViTHDVisionTower + projector from a released checkpoint folder.  Returns (tower, projector, config)."""
    from .projector import build_vision_projector
    from .tower import FastViTHDVisionTower
    tower_sd, proj_sd, config = read_state_dicts(path)

    class Args:
        mm_vision_tower = config.get("mm_vision_tower", "mobileclip_l_1024")
        unfreeze_mm_vision_tower = False
        mm_projector_type = config.get("mm_projector_type", "mlp2x_gelu")
        mm_hidden_size = config.get("mm_hidden_size", 3072)
        hidden_size = config.get("hidden_size", proj_sd["0.weight"].shape[0] if "0.weight" in proj_sd else 0)
    tower = FastViTHDVisionTower(Args.mm_vision_tower, Args(), delay_load=False, max_batch=max_batch)
    missing = set(tower.state_dict().keys()) - set(tower_sd.keys())
    if missing:
        raise KeyError(f"{path}: tower tensors missing from the checkpoint, e.g. {sorted(missing)[:3]}")
    tower.load_state_dict(tower_sd, strict=True)
    projector = None
    if proj_sd:
        projector = build_vision_projector(Args())
        projector.load_state_dict(proj_sd, strict=True)
    if device is not None or dtype is not None:
        tower.to(device=device, dtype=dtype)
        if projector is not None:
            projector.to(device=device, dtype=dtype)
    return tower, projector, config
