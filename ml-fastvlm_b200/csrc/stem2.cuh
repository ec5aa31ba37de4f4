// Stem, second generation: convolutional_stem blocks 0 and 1 fused (mci.py:567-590):
//     NCHW image (fp32 / fp16 / bf16) -> conv3x3 s2 (3 -> 96) + GELU -> dw3x3 s2 + GELU -> NHWC bf16 [B, R/4, R/4, 96]
// Same decomposition as stem_kernel (stem_attn_se.cuh: an 8 x 8 output tile needs a 17 x 17 conv0 patch, conv0 is an implicit GEMM on
// mma.sync m16n8k16 with f16 inputs; block 2, the 1x1 + GELU, is a tcgen05 GEMM launch), rebuilt around what ncu showed in the first
// version (39.6 k warp-instructions per tile at IPC 1.5, 2 CTAs / SM):
//   * persistent CTAs, three per SM (B fragments of w0 live in a lane-major smem table instead of 48 registers): weights are staged
//     once per CTA, not once per tile;
//   * the input patch is fetched as 16-byte row chunks (the old kernel issued one 2-byte load and ~20 index instructions per pixel);
//   * both GELUs run as packed half2 (one tanh.approx.f16x2 per two channels), the conv0 patch is kept as f16x2 (3 more mantissa
//     bits than the bf16 the first version stored);
//   * the depthwise stage slides a 3 x 9 register window over four outputs (27 smem reads for 4 outputs instead of 36) and
//     accumulates both channels of a pair with one FFMA2.
#pragma once
#include "convffn.cuh"        // gelu_f16x2
#include "mixer_tz.cuh"       // ffma2
#include "stem_attn_se.cuh"

namespace fvhd {

struct Stem2 {
    static constexpr int C = 96, TO = 8, MID = 17, IN = 35;
    static constexpr int INP = 40;                 // input patch pitch (halves): 5 chunks of 8; patch column xx lives at xx + 5
    static constexpr int MIDP = 49;                // words per conv0 pixel (48 channel pairs + 1 pad)
    static constexpr int NPIX = MID * MID;         // 289
    static constexpr int MT = (NPIX + 15) / 16;    // 19 m-tiles
    static constexpr int THREADS = 256;
    static constexpr int BT_WORDS = 12 * 2 * 2 * 32;                       // B-fragment table: [n-tile][k-step][half][lane]
    static constexpr int OFF_W1 = BT_WORDS * 4;                            // fp32 [9][96]
    static constexpr int OFF_B0 = OFF_W1 + 9 * C * 4;
    static constexpr int OFF_B1 = OFF_B0 + C * 4;
    static constexpr int OFF_IN = OFF_B1 + C * 4;                          // fp16 [3][35][40]
    static constexpr int OFF_S1 = OFF_IN + ((3 * IN * INP * 2 + 15) / 16) * 16;
    static constexpr size_t SMEM = (size_t)OFF_S1 + (size_t)NPIX * MIDP * 4;
    static_assert(OFF_IN % 16 == 0, "16-byte patch stores");
};

template <typename T> __device__ __forceinline__ uint4 stem2_chunk_f16(const T* p);          // 8 consecutive pixels -> 8 halves
template <> __device__ __forceinline__ uint4 stem2_chunk_f16<__half>(const __half* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
template <> __device__ __forceinline__ uint4 stem2_chunk_f16<bf16>(const bf16* p) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    uint4 r;
    __half2 h;
    h = __floats2half2_rn(a.x, a.y); r.x = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2half2_rn(b.x, b.y); r.y = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2half2_rn(c.x, c.y); r.z = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2half2_rn(d.x, d.y); r.w = *reinterpret_cast<uint32_t*>(&h);
    return r;
}
template <> __device__ __forceinline__ uint4 stem2_chunk_f16<float>(const float* p) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    uint4 r;
    __half2 h;
    h = __floats2half2_rn(a.x, a.y); r.x = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2half2_rn(a.z, a.w); r.y = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2half2_rn(b.x, b.y); r.z = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2half2_rn(b.z, b.w); r.w = *reinterpret_cast<uint32_t*>(&h);
    return r;
}

// grid: persistent, <= 3 CTAs per SM; tile index -> (image, tile row, tile column).  R % 8 == 0 (the engine requires R % 64 == 0).
template <typename T>
__global__ void __launch_bounds__(Stem2::THREADS, 3)
stem2_kernel(const IoBlock* __restrict__ io, bf16* __restrict__ out, const float* __restrict__ w0 /*[27][96], k=(ci*3+ky)*3+kx*/,
             const float* __restrict__ b0, const float* __restrict__ w1 /*[9][96]*/, const float* __restrict__ b1, int R, int tiles_x,
             int n_tiles /*batch * tiles_x * tiles_x*/) {
    using S = Stem2;
    extern __shared__ __align__(16) uint8_t stem2_smem[];
    uint32_t* bt = reinterpret_cast<uint32_t*>(stem2_smem);
    float* w1s = reinterpret_cast<float*>(stem2_smem + S::OFF_W1);
    float* b0s = reinterpret_cast<float*>(stem2_smem + S::OFF_B0);
    float* b1s = reinterpret_cast<float*>(stem2_smem + S::OFF_B1);
    __half* sin = reinterpret_cast<__half*>(stem2_smem + S::OFF_IN);
    uint32_t* s1 = reinterpret_cast<uint32_t*>(stem2_smem + S::OFF_S1);

    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int R2 = R / 2, R4 = R / 4;

    // ---- constants, once per CTA (never written by a kernel of the forward): before the PDL wait
    for (int i = threadIdx.x; i < S::BT_WORDS; i += S::THREADS) {
        const int ln = i & 31, h2 = (i >> 5) & 1, s = (i >> 6) & 1, nt = i >> 7;
        const int k = 16 * s + 2 * (ln & 3) + 8 * h2, n = nt * 8 + (ln >> 2);
        const float lo = k < 27 ? __ldg(w0 + k * S::C + n) : 0.f;
        const float hi = k + 1 < 27 ? __ldg(w0 + (k + 1) * S::C + n) : 0.f;
        const __half2 h = __floats2half2_rn(lo, hi);
        bt[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    for (int i = threadIdx.x; i < 9 * S::C; i += S::THREADS) w1s[i] = __ldg(w1 + i);
    if (threadIdx.x < S::C) { b0s[threadIdx.x] = __ldg(b0 + threadIdx.x); b1s[threadIdx.x] = __ldg(b1 + threadIdx.x); }
    // patch offsets of this thread's 8 K indices: k -> (ci, ky, kx) -> (ci*35 + ky)*40 + kx + 5  (k >= 27: the weight is zero)
    int koff[2][2][2];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = 16 * s + 2 * t + 8 * h2 + e;
                const int kk = k < 27 ? k : 0;
                const int ci = kk / 9, r9 = kk - ci * 9, ky = r9 / 3, kx = r9 - ky * 3;
                koff[s][h2][e] = (ci * S::IN + ky) * S::INP + kx + 5;
            }
    pdl_wait();                                                          // io block is written by set_io_kernel
    const T* __restrict__ img = reinterpret_cast<const T*>(io->images);
    const int tiles_img = tiles_x * tiles_x;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b = tile / tiles_img, tl = tile - b * tiles_img;
        const int ty0 = (tl / tiles_x) * S::TO, tx0 = (tl % tiles_x) * S::TO;
        const int iy0 = 4 * ty0 - 3;                          // input patch origin (row); column xx <-> gx = 4 tx0 - 8 + (xx + 5)
        const int cy0 = 2 * ty0 - 1, cx0 = 2 * tx0 - 1;       // conv0 patch origin
        __syncthreads();                                      // previous tile's phase 2 has finished with s1 / sin (and the constants are staged)

        // ---- input patch: 3 x 35 rows x 5 chunks of 8 pixels, zero outside the image
        for (int i = threadIdx.x; i < 3 * S::IN * 5; i += S::THREADS) {
            const int ch5 = i % 5, row = i / 5;               // row = ci * 35 + yy
            const int ci = row / S::IN, yy = row - ci * S::IN;
            const int gy = iy0 + yy, gx = 4 * tx0 - 8 + ch5 * 8;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (gy >= 0 && gy < R && gx >= 0 && gx + 8 <= R) v = stem2_chunk_f16<T>(img + (((size_t)b * 3 + ci) * R + gy) * R + gx);
            *reinterpret_cast<uint4*>(sin + row * S::INP + ch5 * 8) = v;
        }
        __syncthreads();

        // ---- phase 1: conv0 3x3 s2 + GELU on the 17 x 17 patch, implicit GEMM on mma.sync (M = 289 px, N = 96, K = 27 -> 32)
        {
            const uint16_t* sinu = reinterpret_cast<const uint16_t*>(sin);
            for (int mt = warp; mt < S::MT; mt += S::THREADS / 32) {
                int pix[2], base[2];
                bool inside[2];
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    int pp = mt * 16 + g + 8 * rr;
                    pix[rr] = pp;
                    if (pp > S::NPIX - 1) pp = S::NPIX - 1;              // padded rows: computed, never stored
                    const int py = pp / S::MID, px = pp - py * S::MID;
                    base[rr] = (2 * py) * S::INP + 2 * px;
                    const int cy = cy0 + py, cx = cx0 + px;
                    inside[rr] = cy >= 0 && cy < R2 && cx >= 0 && cx < R2;   // outside: zero padding of the depthwise conv
                }
                uint32_t af[2][4];
#pragma unroll
                for (int s = 0; s < 2; ++s)
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
                        for (int rr = 0; rr < 2; ++rr)
                            af[s][h2 * 2 + rr] = (uint32_t)sinu[base[rr] + koff[s][h2][0]] | ((uint32_t)sinu[base[rr] + koff[s][h2][1]] << 16);
#pragma unroll
                for (int half = 0; half < 2; ++half) {                   // 6 n-tiles at a time: 24 accumulator registers
                    float d[6][4];
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        const int nt = half * 6 + q;
                        const uint32_t* bq = bt + nt * 128 + lane;
                        d[q][0] = d[q][1] = d[q][2] = d[q][3] = 0.f;
                        mma_f16_16816(d[q], af[0], bq[0], bq[32]);
                        mma_f16_16816(d[q], af[1], bq[64], bq[96]);
                    }
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        if (pix[rr] >= S::NPIX) continue;
                        uint32_t* dst = s1 + pix[rr] * S::MIDP + half * 24 + t;
#pragma unroll
                        for (int q = 0; q < 6; ++q) {
                            const float2 bb = *reinterpret_cast<const float2*>(b0s + (half * 6 + q) * 8 + 2 * t);
                            const uint32_t v = gelu_f16x2(d[q][2 * rr + 0] + bb.x, d[q][2 * rr + 1] + bb.y);
                            dst[q * 4] = inside[rr] ? v : 0u;
                        }
                    }
                }
            }
        }
        __syncthreads();

        // ---- phase 2: depthwise 3x3 s2 + GELU; item = (channel pair, output row, strip of 4 outputs)
        for (int it = threadIdx.x; it < 48 * S::TO * 2; it += S::THREADS) {
            const int cp = it % 48, rest = it / 48;
            const int oy = rest >> 1, ox0 = (rest & 1) * 4;
            const float2 bb = *reinterpret_cast<const float2*>(b1s + 2 * cp);
            float2 acc[4] = {bb, bb, bb, bb};
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const uint32_t* rowp = s1 + ((2 * oy + ky) * S::MID + 2 * ox0) * S::MIDP + cp;
                float2 v[9];
#pragma unroll
                for (int j = 0; j < 9; ++j) {
                    const uint32_t u = rowp[j * S::MIDP];
                    v[j] = __half22float2(*reinterpret_cast<const __half2*>(&u));
                }
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float2 w = *reinterpret_cast<const float2*>(w1s + (ky * 3 + kx) * S::C + 2 * cp);
#pragma unroll
                    for (int o = 0; o < 4; ++o) ffma2(acc[o], v[2 * o + kx], w);
                }
            }
            const int gy = ty0 + oy;
            if (gy < R4) {
                bf16* orow = out + (((size_t)b * R4 + gy) * R4 + tx0 + ox0) * S::C + 2 * cp;
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    if (tx0 + ox0 + o < R4) {
                        const uint32_t hg = gelu_f16x2(acc[o].x, acc[o].y);
                        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hg));
                        *reinterpret_cast<uint32_t*>(orow + (size_t)o * S::C) = pack_bf16x2(f.x, f.y);
                    }
                }
            }
        }
    }
}

}  // namespace fvhd
