"""process_images on the GPU (row f1) -- mirror of llava/mm_utils.py:168-184 for the FastVLM image processor.

The reference preprocesses on the CPU with PIL (CLIPImageProcessor of the pinned transformers 4.48.3, configured by
mobileclip_encoder.py:45-49: resize shortest edge -> R, BICUBIC; centre crop; x 1/255; mean 0 / std 1).  Here the uint8 RGB
image is uploaded as is (3 B/pixel instead of 12) and `fvhd_preprocess` reproduces Pillow's fixed-point resampler bit for bit,
writing straight into the tower's NCHW input batch.
"""
import ctypes as C

import numpy as np
import torch

from . import lib as L
from .engine import _DT


def _as_u8_hwc(image):
    if isinstance(image, torch.Tensor):
        t = image
        if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
            raise L.FvhdError(f"expected uint8 HWC RGB, got {t.dtype} {tuple(t.shape)}")
        return t.contiguous()
    if isinstance(image, np.ndarray):
        if image.dtype != np.uint8 or image.ndim != 3 or image.shape[2] != 3:
            raise L.FvhdError(f"expected uint8 HWC RGB, got {image.dtype} {image.shape}")
        return torch.from_numpy(np.ascontiguousarray(image))
    if hasattr(image, "convert"):                                     # PIL.Image: do_convert_rgb
        return torch.from_numpy(np.asarray(image.convert("RGB")).copy())
    raise L.FvhdError(f"unsupported image type {type(image)}")


def preprocess_into(engine, image, out, pad=False):
    """One image -> out ([3,R,R] CUDA tensor slice of the tower input batch, fp32/fp16/bf16)."""
    t = _as_u8_hwc(image)
    H, W = int(t.shape[0]), int(t.shape[1])
    if out.dtype not in _DT or tuple(out.shape) != (3, engine.image_size, engine.image_size) or not out.is_contiguous():
        raise L.FvhdError(f"out must be a contiguous [3,{engine.image_size},{engine.image_size}] fp32/fp16/bf16 CUDA tensor")
    dev = out.device
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        on_host = 0 if t.is_cuda else 1
        if on_host and not t.is_pinned():
            t = t.pin_memory()
        L.check(engine.lib.fvhd_preprocess(engine.handle, stream, t.data_ptr(), on_host, H, W, 1 if pad else 0, out.data_ptr(), _DT[out.dtype]),
                engine.handle)
        if on_host:
            torch.cuda.current_stream(dev).synchronize()              # the pinned staging tensor must outlive the async copy
    return out


def process_images(images, tower, model_cfg=None, dtype=None):
    """mm_utils.process_images for `image_aspect_ratio` None / 'pad' -> CUDA tensor [B,3,R,R] in the tower's dtype.
    ('anyres' builds multi-patch inputs; it is out of scope here and raises.)"""
    aspect = getattr(model_cfg, "image_aspect_ratio", None) if model_cfg is not None else None
    if aspect == "anyres":
        raise NotImplementedError("image_aspect_ratio='anyres' (process_anyres_image) is not built on the GPU path")
    eng = tower.engine() if hasattr(tower, "engine") else tower
    dev = eng.device
    R = eng.image_size
    dt = dtype or (tower.dtype if hasattr(tower, "dtype") else torch.float16)
    if dt not in _DT:
        dt = torch.float16
    out = torch.empty(len(images), 3, R, R, dtype=dt, device=dev)
    for i, im in enumerate(images):
        preprocess_into(eng, im, out[i], pad=(aspect == "pad"))
    return out


def resample_coeffs(in_size, out_size):
    """Pillow's fixed-point bicubic table from the library (host code) -> (bounds [out,2], kk [out,ksize]) int32 arrays."""
    lib = L.load_library()
    ksize = lib.fvhd_resample_coeffs(in_size, out_size, None, None, 0)
    if ksize < 0:
        raise L.FvhdError("fvhd_resample_coeffs failed")
    b = np.zeros((out_size, 2), dtype=np.int32)
    k = np.zeros((out_size, ksize), dtype=np.int32)
    rc = lib.fvhd_resample_coeffs(in_size, out_size, b.ctypes.data_as(C.POINTER(C.c_int)), k.ctypes.data_as(C.POINTER(C.c_int)), k.size)
    if rc < 0:
        raise L.FvhdError("fvhd_resample_coeffs failed")
    return b, k
