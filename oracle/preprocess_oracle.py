"""TEST INFRASTRUCTURE ONLY -- restatement of the reference's CPU image preprocessing (row f1).

Reference behaviour (llava/mm_utils.py:168-184 -> CLIPImageProcessor of the PINNED transformers 4.48.3, configured by
mobileclip_encoder.py:45-49): convert RGB -> resize shortest edge to R with PIL BICUBIC (aspect kept,
long = int(R * long / short)) -> centre crop R x R -> x 1/255 (float64 multiply, cast to float32) -> normalise with
mean 0 / std 1 (identity) -> CHW float32.  `image_aspect_ratio == 'pad'` first pastes the image centred on a square
canvas of colour int(255 * mean) = 0 (expand2square, mm_utils.py:154-165).

The arithmetic that matters is Pillow's 8-bit resampler (third-party, Pillow src/libImaging/Resample.c, version 12.2.0 in
this image): separable, horizontal pass then vertical pass, uint8 intermediate, coefficients in fixed point
(PRECISION_BITS = 22).  `resize_bicubic_u8` restates it; tests pin it bit-for-bit against PIL itself.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size, out_size):
    """Resample.c:precompute_coeffs + normalize_coeffs_8bpc for box (0, in_size).  -> (bounds [out,2] int, kk [out,ksize] int32)."""
    support_f = 2.0
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = support_f * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ww = 0.0
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [0.0] * ksize
        for x in range(xmax):
            w = _bicubic((x + xmin - center + 0.5) * ss)
            k[x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
        for x in range(ksize):
            v = k[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis(img, out_size):
    """img uint8 [N, in_size, C] filtered along axis 1 -> uint8 [N, out_size, C]."""
    in_size = img.shape[1]
    bounds, kk = precompute_coeffs(in_size, out_size)
    out = np.empty((img.shape[0], out_size, img.shape[2]), dtype=np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        xmin, xmax = bounds[xx]
        acc = np.full((img.shape[0], img.shape[2]), 1 << (PRECISION_BITS - 1), dtype=np.int64)
        acc += (src[:, xmin:xmin + xmax, :] * kk[xx, :xmax][None, :, None]).sum(1)
        out[:, xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_bicubic_u8(img, out_w, out_h):
    """PIL Image.resize((out_w, out_h), BICUBIC) for an RGB uint8 HWC array: horizontal pass, then vertical pass."""
    h, w, _ = img.shape
    x = img
    if out_w != w:
        x = _resample_axis(x, out_w)
    if out_h != h:
        x = _resample_axis(x.transpose(1, 0, 2), out_h).transpose(1, 0, 2)
    return np.ascontiguousarray(x)


def resize_output_size(h, w, shortest_edge):
    """transformers get_resize_output_image_size(size=int, default_to_square=False): short -> R, long -> int(R*long/short)."""
    short, long_ = (w, h) if w <= h else (h, w)
    new_short, new_long = shortest_edge, int(shortest_edge * long_ / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)      # (out_h, out_w)


def expand2square(img, fill=0):
    """mm_utils.py:154-165 on an HWC uint8 array."""
    h, w, c = img.shape
    if w == h:
        return img
    s = max(w, h)
    out = np.full((s, s, c), fill, dtype=np.uint8)
    if w > h:
        out[(w - h) // 2:(w - h) // 2 + h, :, :] = img
    else:
        out[:, (h - w) // 2:(h - w) // 2 + w, :] = img
    return out


RESCALE_LUT = (np.arange(256, dtype=np.float64) * (1 / 255)).astype(np.float32)      # transformers rescale(): float64 multiply


def preprocess(img, res, pad=False):
    """uint8 HWC RGB -> float32 CHW [3, res, res], the tensor the tower receives."""
    if pad:
        img = expand2square(img, 0)
    h, w, _ = img.shape
    oh, ow = resize_output_size(h, w, res)
    x = resize_bicubic_u8(img, ow, oh)
    top, left = (oh - res) // 2, (ow - res) // 2
    x = x[top:top + res, left:left + res, :]
    return np.ascontiguousarray(RESCALE_LUT[x].transpose(2, 0, 1))


# ------------------------------------------------------------------ image_aspect_ratio == "anyres"
def select_best_resolution(original_size, possible_resolutions):
    """mm_utils.py:14-43.  original_size = (width, height); returns the (width, height) pin that keeps the most pixels of the
    aspect-preserving downscale and, among those, wastes the least canvas."""
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


def anyres_geometry(w, h, grid_pinpoints):
    """mm_utils.py:46-76 (resize_and_pad_image) without the pixels: canvas (cw, ch), resized size (nw, nh), paste offset (px, py)."""
    import math
    cw, ch = select_best_resolution((w, h), [tuple(p) for p in grid_pinpoints])
    sw, sh = cw / w, ch / h
    if sw < sh:
        nw, nh = cw, min(math.ceil(h * sw), ch)
    else:
        nh, nw = ch, min(math.ceil(w * sh), cw)
    return cw, ch, nw, nh, (cw - nw) // 2, (ch - nh) // 2


def preprocess_anyres(img, res, grid_pinpoints):
    """mm_utils.py:121-147 (process_anyres_image) for the FastVLM processor (crop = shortest_edge = res, mean 0, std 1):
    [global view resized to res x res] + [res x res patches of the resized image pasted on the black best-fit canvas], each
    x 1/255 -> float32 [1 + n, 3, res, res].  Patches already have the processor's size, so its resize / crop are no-ops."""
    h, w, _ = img.shape
    cw, ch, nw, nh, px, py = anyres_geometry(w, h, grid_pinpoints)
    canvas = np.zeros((ch, cw, 3), dtype=np.uint8)
    canvas[py:py + nh, px:px + nw] = resize_bicubic_u8(img, nw, nh)
    views = [resize_bicubic_u8(img, res, res)]
    for i in range(0, ch, res):
        for j in range(0, cw, res):
            patch = np.zeros((res, res, 3), dtype=np.uint8)          # PIL crop beyond the canvas pads with black
            sub = canvas[i:i + res, j:j + res]
            patch[:sub.shape[0], :sub.shape[1]] = sub
            views.append(patch)
    return np.ascontiguousarray(RESCALE_LUT[np.stack(views)].transpose(0, 3, 1, 2))
