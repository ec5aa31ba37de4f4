// Depthwise convolutions on NHWC bf16 activations (fp32 accumulate, fp32 weights in smem).
//
//   dwconv_kernel<KS,S,MULT,ACT>  one depthwise conv (+bias, +GELU):
//       RepCPE dw7x7 (mci.py:992-995), ConvFFN dw7x7+BN folded (mci.py:885-907,921),
//       PatchEmbed dw7x7 s2 with channel multiplier 2 + GELU (mci.py:442-451),
//       conv_exp dw3x3 multiplier 2 (mci.py:1401-1411).
//   repmixer_dw_kernel            RepMixer dw3x3 (mci.py:808-811) fused with the ConvFFN dw7x7+BN of the
//       same block (mci.py:921): x -> y (block residual, written once) -> z (fc1 input); y never
//       leaves shared memory between the two convs.
//
// Tiling: a CTA owns a TOHxTOW output tile of ONE 32-channel group; the input tile (+halo) is staged
// once in shared memory with 16-B vector loads; a thread then owns one bf16x2 channel pair and slides
// a register window along a row strip (SW outputs), so each staged input is read ~KS/SW.. times from
// smem instead of KS*KS times.  Row pitch is an odd number of pixels so the two half-warps (which
// work on adjacent rows) hit disjoint banks.
#pragma once
#include "ptx.cuh"

namespace fvhd {

constexpr int DW_CG = 32;          // input channels per CTA
constexpr int DW_THREADS = 256;

template <int V>
struct OddUp { static constexpr int value = (V & 1) ? V : V + 1; };

// Stage an (IH x IW) pixel window x 32 channels into smem (zero outside the image).
template <int IH, int IW, int IWP>
__device__ __forceinline__ void dw_stage_tile(uint32_t* tile, const bf16* __restrict__ in, int b, int H, int W, int C,
                                              int iy0, int ix0, int c0) {
    constexpr int CHUNKS = IH * IW * 4;               // 16-B chunks (8 channels each)
    for (int i = threadIdx.x; i < CHUNKS; i += DW_THREADS) {
        const int ch = i & 3;
        const int px = i >> 2;
        const int ty = px / IW, tx = px - ty * IW;
        const int gy = iy0 + ty, gx = ix0 + tx;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (gy >= 0 && gy < H && gx >= 0 && gx < W)
            v = __ldg(reinterpret_cast<const uint4*>(in + (((size_t)b * H + gy) * W + gx) * C + c0 + ch * 8));
        *reinterpret_cast<uint4*>(tile + (ty * IWP + tx) * 16 + ch * 4) = v;
    }
}

// One strip: SW consecutive outputs of one row, one channel pair (-> 2*MULT output channels).
template <int KS, int S, int MULT, int SW>
__device__ __forceinline__ void dw_strip(const uint32_t* __restrict__ tile_row0, int pitch_words,
                                         const float* __restrict__ wsm, int cp, float (&acc)[SW][2 * MULT]) {
    constexpr int NIN = (SW - 1) * S + KS;
#pragma unroll
    for (int ky = 0; ky < KS; ++ky) {
        const uint32_t* rowp = tile_row0 + ky * pitch_words;
        float2 xin[NIN];
#pragma unroll
        for (int j = 0; j < NIN; ++j) xin[j] = unpack_bf16x2(rowp[j * 16]);
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
            const float* wp = wsm + (ky * KS + kx) * (DW_CG * MULT) + cp * 2 * MULT;
            if (MULT == 1) {
                const float2 w = *reinterpret_cast<const float2*>(wp);
#pragma unroll
                for (int j = 0; j < SW; ++j) {
                    acc[j][0] = fmaf(xin[j * S + kx].x, w.x, acc[j][0]);
                    acc[j][1] = fmaf(xin[j * S + kx].y, w.y, acc[j][1]);
                }
            } else {
                const float4 w = *reinterpret_cast<const float4*>(wp);
#pragma unroll
                for (int j = 0; j < SW; ++j) {
                    acc[j][0] = fmaf(xin[j * S + kx].x, w.x, acc[j][0]);
                    acc[j][1] = fmaf(xin[j * S + kx].x, w.y, acc[j][1]);
                    acc[j][2 * MULT - 2] = fmaf(xin[j * S + kx].y, w.z, acc[j][2 * MULT - 2]);
                    acc[j][2 * MULT - 1] = fmaf(xin[j * S + kx].y, w.w, acc[j][2 * MULT - 1]);
                }
            }
        }
    }
}

template <int KS, int S, int MULT, int TOH, int TOW>
struct DwCfg {
    static constexpr int IH = (TOH - 1) * S + KS;
    static constexpr int IW = (TOW - 1) * S + KS;
    static constexpr int IWP = OddUp<IW>::value;
    static constexpr int TILE_WORDS = IH * IWP * 16;
    static constexpr int W_FLOATS = KS * KS * DW_CG * MULT;
    static constexpr int B_FLOATS = DW_CG * MULT;
    static constexpr size_t SMEM = (size_t)(TILE_WORDS + W_FLOATS + B_FLOATS) * 4;
};

// grid: x = tiles_x * tiles_y, y = C/32, z = B
template <int KS, int S, int MULT, int ACT, int TOH, int TOW, int SW>
__global__ void __launch_bounds__(DW_THREADS, 2)
dwconv_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, const float* __restrict__ w /*[KS*KS][C*MULT]*/,
              const float* __restrict__ bias /*[C*MULT]*/, int H, int W, int C, int Ho, int Wo, int tiles_x) {
    using Cfg = DwCfg<KS, S, MULT, TOH, TOW>;
    extern __shared__ __align__(16) uint32_t dw_smem[];
    uint32_t* tile = dw_smem;
    float* wsm = reinterpret_cast<float*>(dw_smem + Cfg::TILE_WORDS);
    float* bsm = wsm + Cfg::W_FLOATS;

    const int b = blockIdx.z;
    const int c0 = blockIdx.y * DW_CG;
    const int ty0 = (blockIdx.x / tiles_x) * TOH;
    const int tx0 = (blockIdx.x % tiles_x) * TOW;
    const int Cout = C * MULT;

    for (int i = threadIdx.x; i < Cfg::W_FLOATS; i += DW_THREADS) {
        const int tap = i / (DW_CG * MULT), o = i - tap * (DW_CG * MULT);
        wsm[i] = __ldg(w + (size_t)tap * Cout + c0 * MULT + o);
    }
    for (int i = threadIdx.x; i < Cfg::B_FLOATS; i += DW_THREADS) bsm[i] = __ldg(bias + c0 * MULT + i);
    dw_stage_tile<Cfg::IH, Cfg::IW, Cfg::IWP>(tile, in, b, H, W, C, ty0 * S - KS / 2, tx0 * S - KS / 2, c0);
    __syncthreads();

    constexpr int STRIPS = TOW / SW;
    constexpr int ITEMS = 16 * TOH * STRIPS;
    static_assert(TOW % SW == 0 && TOH % 2 == 0, "tile shape");
    for (int it = threadIdx.x; it < ITEMS; it += DW_THREADS) {
        const int cp = it & 15;
        const int t = it >> 4;
        const int sub = t & 1;
        const int u = t >> 1;
        const int strip = u % STRIPS;
        const int oy = (u / STRIPS) * 2 + sub;
        const int ox0 = strip * SW;
        float acc[SW][2 * MULT];
#pragma unroll
        for (int j = 0; j < SW; ++j)
#pragma unroll
            for (int m = 0; m < 2 * MULT; ++m) acc[j][m] = bsm[cp * 2 * MULT + m];
        dw_strip<KS, S, MULT, SW>(tile + ((oy * S) * Cfg::IWP + ox0 * S) * 16 + cp, Cfg::IWP * 16, wsm, cp, acc);
        const int gy = ty0 + oy;
        if (gy >= Ho) continue;
#pragma unroll
        for (int j = 0; j < SW; ++j) {
            const int gx = tx0 + ox0 + j;
            if (gx >= Wo) break;
            float v[2 * MULT];
#pragma unroll
            for (int m = 0; m < 2 * MULT; ++m) v[m] = ACT ? gelu_erf(acc[j][m]) : acc[j][m];
            bf16* op = out + (((size_t)b * Ho + gy) * Wo + gx) * Cout + (c0 + cp * 2) * MULT;
            if (MULT == 1) {
                *reinterpret_cast<uint32_t*>(op) = pack_bf16x2(v[0], v[1]);
            } else {
                uint2 o2;
                o2.x = pack_bf16x2(v[0], v[1]);
                o2.y = pack_bf16x2(v[2 * MULT - 2], v[2 * MULT - 1]);
                *reinterpret_cast<uint2*>(op) = o2;
            }
        }
    }
}

// ---------------------------------------------------------------- fused RepMixer dw3x3 -> ConvFFN dw7x7(+BN)
struct MixCfg {
    static constexpr int TO = 16;                  // output tile 16x16
    static constexpr int YH = TO + 6;              // y region (3-px halo for the 7x7)
    static constexpr int XH = TO + 8;              // x region (+1 more for the 3x3)
    static constexpr int XP = OddUp<XH>::value;    // 25
    static constexpr int YP = OddUp<YH>::value;    // 23
    static constexpr int X_WORDS = XH * XP * 16;
    static constexpr int Y_WORDS = YH * YP * 16;
    static constexpr int W_FLOATS = (9 + 49) * DW_CG + 2 * DW_CG;
    static constexpr size_t SMEM = (size_t)(X_WORDS + Y_WORDS + W_FLOATS) * 4;
};

__global__ void __launch_bounds__(DW_THREADS, 2)
repmixer_dw_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, bf16* __restrict__ z,
                   const float* __restrict__ w3 /*[9][C]*/, const float* __restrict__ b3,
                   const float* __restrict__ w7 /*[49][C], BN folded*/, const float* __restrict__ b7,
                   int H, int W, int C, int tiles_x) {
    extern __shared__ __align__(16) uint32_t dw_smem[];
    uint32_t* sx = dw_smem;
    uint32_t* sy = sx + MixCfg::X_WORDS;
    float* w3s = reinterpret_cast<float*>(sy + MixCfg::Y_WORDS);
    float* w7s = w3s + 9 * DW_CG;
    float* b3s = w7s + 49 * DW_CG;
    float* b7s = b3s + DW_CG;

    const int b = blockIdx.z;
    const int c0 = blockIdx.y * DW_CG;
    const int ty0 = (blockIdx.x / tiles_x) * MixCfg::TO;
    const int tx0 = (blockIdx.x % tiles_x) * MixCfg::TO;

    for (int i = threadIdx.x; i < 9 * DW_CG; i += DW_THREADS) w3s[i] = __ldg(w3 + (size_t)(i / DW_CG) * C + c0 + (i % DW_CG));
    for (int i = threadIdx.x; i < 49 * DW_CG; i += DW_THREADS) w7s[i] = __ldg(w7 + (size_t)(i / DW_CG) * C + c0 + (i % DW_CG));
    if (threadIdx.x < DW_CG) {
        b3s[threadIdx.x] = __ldg(b3 + c0 + threadIdx.x);
        b7s[threadIdx.x] = __ldg(b7 + c0 + threadIdx.x);
    }
    dw_stage_tile<MixCfg::XH, MixCfg::XH, MixCfg::XP>(sx, x, b, H, W, C, ty0 - 4, tx0 - 4, c0);
    __syncthreads();

    // phase 1: y = dw3x3(x) + b on the (TO+6)^2 region; zero outside the image (the 7x7's zero padding)
    {
        constexpr int SW = 11, STRIPS = MixCfg::YH / SW;     // 22 = 2 x 11
        constexpr int ITEMS = 16 * MixCfg::YH * STRIPS;
        for (int it = threadIdx.x; it < ITEMS; it += DW_THREADS) {
            const int cp = it & 15;
            const int t = it >> 4;
            const int sub = t & 1;
            const int u = t >> 1;
            const int strip = u % STRIPS;
            const int ry = (u / STRIPS) * 2 + sub;            // row in the y region
            const int rx0 = strip * SW;
            float acc[SW][2];
#pragma unroll
            for (int j = 0; j < SW; ++j) { acc[j][0] = b3s[cp * 2]; acc[j][1] = b3s[cp * 2 + 1]; }
            dw_strip<3, 1, 1, SW>(sx + (ry * MixCfg::XP + rx0) * 16 + cp, MixCfg::XP * 16, w3s, cp, acc);
            const int gy = ty0 - 3 + ry;
            const bool row_in = gy >= 0 && gy < H;
            const bool row_center = ry >= 3 && ry < 3 + MixCfg::TO;
#pragma unroll
            for (int j = 0; j < SW; ++j) {
                const int rx = rx0 + j;
                const int gx = tx0 - 3 + rx;
                const bool in_img = row_in && gx >= 0 && gx < W;
                const uint32_t pk = in_img ? pack_bf16x2(acc[j][0], acc[j][1]) : 0u;
                sy[(ry * MixCfg::YP + rx) * 16 + cp] = pk;
                if (in_img && row_center && rx >= 3 && rx < 3 + MixCfg::TO)
                    *reinterpret_cast<uint32_t*>(y + (((size_t)b * H + gy) * W + gx) * C + c0 + cp * 2) = pk;
            }
        }
    }
    __syncthreads();

    // phase 2: z = dw7x7(y) (BN folded) on the 16x16 tile
    {
        constexpr int SW = 8, STRIPS = MixCfg::TO / SW;
        constexpr int ITEMS = 16 * MixCfg::TO * STRIPS;
        for (int it = threadIdx.x; it < ITEMS; it += DW_THREADS) {
            const int cp = it & 15;
            const int t = it >> 4;
            const int sub = t & 1;
            const int u = t >> 1;
            const int strip = u % STRIPS;
            const int oy = (u / STRIPS) * 2 + sub;
            const int ox0 = strip * SW;
            float acc[SW][2];
#pragma unroll
            for (int j = 0; j < SW; ++j) { acc[j][0] = b7s[cp * 2]; acc[j][1] = b7s[cp * 2 + 1]; }
            dw_strip<7, 1, 1, SW>(sy + (oy * MixCfg::YP + ox0) * 16 + cp, MixCfg::YP * 16, w7s, cp, acc);
            const int gy = ty0 + oy;
            if (gy >= H) continue;
#pragma unroll
            for (int j = 0; j < SW; ++j) {
                const int gx = tx0 + ox0 + j;
                if (gx >= W) break;
                *reinterpret_cast<uint32_t*>(z + (((size_t)b * H + gy) * W + gx) * C + c0 + cp * 2) = pack_bf16x2(acc[j][0], acc[j][1]);
            }
        }
    }
}

}  // namespace fvhd
