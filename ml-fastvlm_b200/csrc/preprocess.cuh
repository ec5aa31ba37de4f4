// Row f1: the reference's CPU image preprocessing on the GPU, bit-exact.
//
// Reference: llava/mm_utils.py:168-184 (process_images) -> CLIPImageProcessor of the pinned transformers 4.48.3, configured by
// mobileclip_encoder.py:45-49: resize shortest edge to R with PIL BICUBIC -> centre crop R x R -> x 1/255 -> (mean 0, std 1)
// -> CHW.  'pad' mode pastes the image centred on a zero square first (expand2square, mm_utils.py:154-165); 'anyres'
// (process_anyres_image, mm_utils.py:121-147) resizes to a computed size, pastes on a black best-fit canvas and cuts R x R tiles.
// The arithmetic is Pillow's 8-bit separable resampler (src/libImaging/Resample.c): horizontal pass, uint8 intermediate,
// vertical pass, coefficients in 22-bit fixed point.  The tables are computed on the host with the same double arithmetic
// (resample_coeffs below), the passes are HBM-bound byte kernels: one thread per output pixel (3 channels), int32 accumulate.
#pragma once
#include <cmath>
#include <vector>

#include "ptx.cuh"

namespace fvhd {

constexpr int RS_PRECISION_BITS = 32 - 8 - 2;

inline double rs_bicubic(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for the box (0, in_size).  bounds: [out][2] = (xmin, count); kk: [out][ksize].
inline int resample_coeffs(int in_size, int out_size, std::vector<int>& bounds, std::vector<int>& kk) {
    const double support_f = 2.0;
    const double scale = (double)in_size / out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = support_f * filterscale;
    const int ksize = (int)std::ceil(support) * 2 + 1;
    bounds.assign((size_t)out_size * 2, 0);
    kk.assign((size_t)out_size * ksize, 0);
    std::vector<double> k(ksize);
    const double ss = 1.0 / filterscale;
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = (xx + 0.5) * scale;
        double ww = 0.0;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < ksize; ++x) k[x] = 0.0;
        for (int x = 0; x < xmax; ++x) {
            const double w = rs_bicubic((x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        for (int x = 0; x < xmax; ++x)
            if (ww != 0.0) k[x] /= ww;
        for (int x = 0; x < ksize; ++x) {
            const double v = k[x];
            kk[(size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << RS_PRECISION_BITS)) : (int)(0.5 + v * (1 << RS_PRECISION_BITS));
        }
        bounds[(size_t)xx * 2] = xmin;
        bounds[(size_t)xx * 2 + 1] = xmax;
    }
    return ksize;
}

__device__ __forceinline__ int rs_clip8(int v) {
    v >>= RS_PRECISION_BITS;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// Horizontal pass over the (virtually zero-padded) source: src [H, W, 3] u8 placed at (offy, offx) inside an Hs x Ws canvas.
// tmp [Hs, ow, 3] u8.  grid (ceil(ow/128), Hs), 128 threads.
__global__ void __launch_bounds__(128)
resample_h_kernel(const uint8_t* __restrict__ src, int H, int W, int offy, int offx, int Ws, uint8_t* __restrict__ tmp, int ow,
                  const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
    pdl_launch_dependents();
    pdl_wait();
    const int x = blockIdx.x * 128 + threadIdx.x;
    const int ys = blockIdx.y;
    if (x >= ow) return;
    const int y = ys - offy;
    int a0 = 1 << (RS_PRECISION_BITS - 1), a1 = a0, a2 = a0;
    if (y >= 0 && y < H) {
        const int xmin = bounds[2 * x], cnt = bounds[2 * x + 1];
        const int* kp = kk + (size_t)x * ksize;
        const uint8_t* row = src + (size_t)y * W * 3;
        for (int k = 0; k < cnt; ++k) {
            const int xs = xmin + k - offx;
            if (xs >= 0 && xs < W) {
                const int c = kp[k];
                a0 += row[xs * 3 + 0] * c;
                a1 += row[xs * 3 + 1] * c;
                a2 += row[xs * 3 + 2] * c;
            }
        }
    }
    (void)Ws;
    uint8_t* o = tmp + ((size_t)ys * ow + x) * 3;
    o[0] = (uint8_t)rs_clip8(a0); o[1] = (uint8_t)rs_clip8(a1); o[2] = (uint8_t)rs_clip8(a2);
}

template <typename T> __device__ __forceinline__ T rs_cast(float v);
template <> __device__ __forceinline__ float rs_cast<float>(float v) { return v; }
template <> __device__ __forceinline__ __half rs_cast<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ bf16 rs_cast<bf16>(float v) { return __float2bfloat16_rn(v); }

// Vertical pass + centre crop + x 1/255 (256-entry LUT, computed in double on the host) + NCHW store: the window lies inside
// the resized image.  tmp [Hs, ow, 3] u8 -> out [3, R, R] of T.  grid (ceil(R/128), R), 128 threads.
template <typename T>
__global__ void __launch_bounds__(128)
resample_v_crop_kernel(const uint8_t* __restrict__ tmp, int ow, T* __restrict__ out, int R, int top, int left,
                       const int* __restrict__ bounds, const int* __restrict__ kk, int ksize, const float* __restrict__ lut) {
    pdl_launch_dependents();
    pdl_wait();
    const int xo = blockIdx.x * 128 + threadIdx.x;
    const int yo = blockIdx.y;
    if (xo >= R) return;
    const int yy = yo + top, xx = xo + left;
    const int ymin = bounds[2 * yy], cnt = bounds[2 * yy + 1];
    const int* kp = kk + (size_t)yy * ksize;
    int a0 = 1 << (RS_PRECISION_BITS - 1), a1 = a0, a2 = a0;
    for (int k = 0; k < cnt; ++k) {
        const uint8_t* p = tmp + ((size_t)(ymin + k) * ow + xx) * 3;
        const int c = kp[k];
        a0 += p[0] * c; a1 += p[1] * c; a2 += p[2] * c;
    }
    const size_t plane = (size_t)R * R;
    const size_t o = (size_t)yo * R + xo;
    out[o] = rs_cast<T>(__ldg(lut + rs_clip8(a0)));
    out[plane + o] = rs_cast<T>(__ldg(lut + rs_clip8(a1)));
    out[2 * plane + o] = rs_cast<T>(__ldg(lut + rs_clip8(a2)));
}

// Tiled variant: vertical pass + window + x 1/255 (256-entry LUT, computed in double on the host) + NCHW store.
// tmp [Hs, ow, 3] u8 (rows of the horizontally resized image) -> out [tiles, 3, R, R] of T.  Tile t = (ty, tx) = (t / tiles_x,
// t % tiles_x) shows rows [top + ty R, +R) x cols [left + tx R, +R) of the vertically resized image (oh x ow); pixels outside
// it are 0 (the black canvas of 'pad' / 'anyres').  The plain centre crop is one tile with top, left >= 0 inside the image.
// grid (ceil(R/128), R, tiles), 128 threads.
template <typename T>
__global__ void __launch_bounds__(128)
resample_v_tiles_kernel(const uint8_t* __restrict__ tmp, int ow, int oh, T* __restrict__ out, int R, int top, int left, int tiles_x,
                       const int* __restrict__ bounds, const int* __restrict__ kk, int ksize, const float* __restrict__ lut) {
    pdl_launch_dependents();
    pdl_wait();
    const int xo = blockIdx.x * 128 + threadIdx.x;
    const int yo = blockIdx.y;
    if (xo >= R) return;
    const int t = blockIdx.z;
    const int yy = yo + top + (t / tiles_x) * R, xx = xo + left + (t % tiles_x) * R;
    int a0 = 0, a1 = 0, a2 = 0;                                  // outside the image: black
    if (yy >= 0 && yy < oh && xx >= 0 && xx < ow) {
        const int ymin = bounds[2 * yy], cnt = bounds[2 * yy + 1];
        const int* kp = kk + (size_t)yy * ksize;
        a0 = a1 = a2 = 1 << (RS_PRECISION_BITS - 1);
        for (int k = 0; k < cnt; ++k) {
            const uint8_t* p = tmp + ((size_t)(ymin + k) * ow + xx) * 3;
            const int c = kp[k];
            a0 += p[0] * c; a1 += p[1] * c; a2 += p[2] * c;
        }
    }
    const size_t plane = (size_t)R * R;
    T* o = out + (size_t)t * 3 * plane + (size_t)yo * R + xo;
    o[0] = rs_cast<T>(__ldg(lut + rs_clip8(a0)));
    o[plane] = rs_cast<T>(__ldg(lut + rs_clip8(a1)));
    o[2 * plane] = rs_cast<T>(__ldg(lut + rs_clip8(a2)));
}

// No resampling needed along an axis (out == in): PIL skips that pass.  Copy kernels keep the code path uniform.
__global__ void __launch_bounds__(128)
pad_copy_kernel(const uint8_t* __restrict__ src, int H, int W, int offy, int offx, uint8_t* __restrict__ tmp, int ow) {
    pdl_launch_dependents();
    pdl_wait();
    const int x = blockIdx.x * 128 + threadIdx.x;
    const int ys = blockIdx.y;
    if (x >= ow) return;
    const int y = ys - offy, xs = x - offx;
    uint8_t* o = tmp + ((size_t)ys * ow + x) * 3;
    if (y >= 0 && y < H && xs >= 0 && xs < W) {
        const uint8_t* p = src + ((size_t)y * W + xs) * 3;
        o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
    } else {
        o[0] = o[1] = o[2] = 0;
    }
}

}  // namespace fvhd
