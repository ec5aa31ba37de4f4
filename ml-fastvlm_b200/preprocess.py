"""process_images on the GPU (row f1) -- mirror of llava/mm_utils.py:168-184 (and :14-147 for 'anyres') for the FastVLM image processor.

The reference preprocesses on the CPU with PIL (CLIPImageProcessor of the pinned transformers 4.48.3, configured by
mobileclip_encoder.py:45-49: resize shortest edge -> R, BICUBIC; centre crop; x 1/255; mean 0 / std 1).  Here the uint8 RGB
image is uploaded as is (3 B/pixel instead of 12) and `fvhd_preprocess` reproduces Pillow's fixed-point resampler bit for bit,
writing straight into the tower's NCHW input batch.
"""
import ast
import ctypes as C
import math

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


def select_best_resolution(original_size, possible_resolutions):
    """mm_utils.py:14-43: (width, height) pin that keeps the most pixels of the aspect-preserving fit, then wastes least."""
    ow, oh = original_size
    best, max_eff, min_waste = None, 0, float("inf")
    for w, h in possible_resolutions:
        scale = min(w / ow, h / oh)
        dw, dh = int(ow * scale), int(oh * scale)
        eff = min(dw * dh, ow * oh)
        waste = w * h - eff
        if eff > max_eff or (eff == max_eff and waste < min_waste):
            max_eff, min_waste, best = eff, waste, (w, h)
    return best


def get_anyres_image_grid_shape(image_size, grid_pinpoints, patch_size):
    """mm_utils.py:100-118: (tiles_x, tiles_y) of the best-fit canvas for an image of (width, height)."""
    pins = grid_pinpoints if isinstance(grid_pinpoints, list) else ast.literal_eval(grid_pinpoints)
    w, h = select_best_resolution(image_size, [tuple(p) for p in pins])
    return w // patch_size, h // patch_size


def anyres_geometry(width, height, patch_size, grid_pinpoints):
    """Host arithmetic of resize_and_pad_image + divide_to_patches (mm_utils.py:46-98) for an image of width x height:
    -> (new_h, new_w, pad_y, pad_x, tiles_y, tiles_x): PIL-resize target, paste offset on the best-fit canvas, tile grid."""
    pins = grid_pinpoints if isinstance(grid_pinpoints, list) else ast.literal_eval(grid_pinpoints)
    cw, ch = select_best_resolution((width, height), [tuple(p) for p in pins])
    sw, sh = cw / width, ch / height                                   # resize_and_pad_image, mm_utils.py:58-68
    if sw < sh:
        nw, nh = cw, min(math.ceil(height * sw), ch)
    else:
        nh, nw = ch, min(math.ceil(width * sh), cw)
    R = patch_size                                                     # divide_to_patches: crops past the canvas are black too
    return nh, nw, (ch - nh) // 2, (cw - nw) // 2, (ch + R - 1) // R, (cw + R - 1) // R


def process_anyres_image(image, engine, grid_pinpoints, dtype=torch.float16):
    """mm_utils.py:121-147 on the GPU: [global view resized to R x R] + the R x R tiles of the image resized (aspect kept) and
    centred on the black best-fit canvas -> CUDA tensor [1 + tiles, 3, R, R].  Two library calls, bit-exact with the PIL path."""
    t = _as_u8_hwc(image)
    H, W = int(t.shape[0]), int(t.shape[1])
    R = engine.image_size
    nh, nw, py, px, ty, tx = anyres_geometry(W, H, R, grid_pinpoints)
    dev = engine.device
    out = torch.empty(1 + ty * tx, 3, R, R, dtype=dtype, device=dev)
    with torch.cuda.device(dev):
        if not t.is_cuda:                                              # one upload (3 B/pixel) serves both calls
            t = (t if t.is_pinned() else t.pin_memory()).to(dev, non_blocking=True)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        L.check(engine.lib.fvhd_preprocess_tiles(engine.handle, stream, t.data_ptr(), 0, H, W, R, R, 0, 0, 1, 1, out[0].data_ptr(), _DT[dtype]),
                engine.handle)
        L.check(engine.lib.fvhd_preprocess_tiles(engine.handle, stream, t.data_ptr(), 0, H, W, nh, nw, py, px, ty, tx, out[1:].data_ptr(), _DT[dtype]),
                engine.handle)
        t.record_stream(torch.cuda.current_stream(dev))               # the staging tensor is read asynchronously
    return out


def process_images(images, tower, model_cfg=None, dtype=None):
    """mm_utils.process_images -> CUDA tensor(s) in the tower's dtype: [B,3,R,R] for `image_aspect_ratio` None / 'pad';
    for 'anyres' one [1 + tiles, 3, R, R] tensor per image, stacked when all images have the same tile count (as the reference)."""
    aspect = getattr(model_cfg, "image_aspect_ratio", None) if model_cfg is not None else None
    eng = tower.engine() if hasattr(tower, "engine") else tower
    dev = eng.device
    R = eng.image_size
    dt = dtype or (tower.dtype if hasattr(tower, "dtype") else torch.float16)
    if dt not in _DT:
        dt = torch.float16
    if aspect == "anyres":
        views = [process_anyres_image(im, eng, model_cfg.image_grid_pinpoints, dt) for im in images]
        if all(v.shape == views[0].shape for v in views):
            return torch.stack(views, dim=0)
        return views
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
