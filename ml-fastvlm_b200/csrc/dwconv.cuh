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
// once in shared memory by ONE TMA 4-D box load (cp.async.bulk.tensor, coords {channel, x, y, image}; the
// out-of-bounds zero fill is the conv's zero padding, and the box is one pixel wider than needed so that the
// row pitch is odd); the per-channel kernels are staged with 16-B vector loads before griddepcontrol.wait; a thread then owns one bf16x2 channel pair and slides
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

// One strip: SW consecutive outputs of one row, one channel pair (-> 2*MULT output channels).
template <int KS, int S, int MULT, int SW, int KYU = KS>
__device__ __forceinline__ void dw_strip(const uint32_t* __restrict__ tile_row0, int pitch_words,
                                         const float* __restrict__ wsm, int cp, float (&acc)[SW][2 * MULT]) {
    constexpr int NIN = (SW - 1) * S + KS;
#pragma unroll KYU
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

// Stage TAPS rows of NCH consecutive per-channel coefficients (row pitch `ldw` floats in global) with 16-B loads,
// all loads of a thread issued before its stores.
template <int TAPS, int NCH, int NT>
__device__ __forceinline__ void dw_stage_weights(float* __restrict__ wsm, const float* __restrict__ w, int ldw, int ch0) {
    constexpr int V = TAPS * NCH / 4;                 // float4 pieces
    constexpr int PER = (V + NT - 1) / NT;
    float4 v[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int i = threadIdx.x + k * NT;
        if (i < V) {
            const int tap = i / (NCH / 4), q = i - tap * (NCH / 4);
            v[k] = __ldg(reinterpret_cast<const float4*>(w + (size_t)tap * ldw + ch0) + q);
        }
    }
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int i = threadIdx.x + k * NT;
        if (i < V) reinterpret_cast<float4*>(wsm)[i] = v[k];
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
    static constexpr size_t SMEM = (size_t)(TILE_WORDS + W_FLOATS + B_FLOATS) * 4 + 16 /*mbarrier*/;
};

// grid: x = tiles_x * tiles_y, y = C/32, z = B
template <int KS, int S, int MULT, int ACT, int TOH, int TOW, int SW>
__global__ void __launch_bounds__(DW_THREADS, 2)
dwconv_kernel(const __grid_constant__ CUtensorMap tmX /*in: NHWC, box {32, IWP, IH, 1}*/, bf16* __restrict__ out,
              const float* __restrict__ w /*[KS*KS][C*MULT]*/, const float* __restrict__ bias /*[C*MULT]*/,
              int H, int W, int C, int Ho, int Wo, int tiles_x) {
    using Cfg = DwCfg<KS, S, MULT, TOH, TOW>;
    extern __shared__ __align__(128) uint32_t dw_smem[];
    uint32_t* tile = dw_smem;
    float* wsm = reinterpret_cast<float*>(dw_smem + Cfg::TILE_WORDS);
    float* bsm = wsm + Cfg::W_FLOATS;
    uint64_t* bar = reinterpret_cast<uint64_t*>(bsm + Cfg::B_FLOATS);

    pdl_launch_dependents();
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * DW_CG;
    const int ty0 = (blockIdx.x / tiles_x) * TOH;
    const int tx0 = (blockIdx.x % tiles_x) * TOW;
    const int Cout = C * MULT;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmX);
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    dw_stage_weights<KS * KS, DW_CG * MULT, DW_THREADS>(wsm, w, Cout, c0 * MULT);
    for (int i = threadIdx.x; i < Cfg::B_FLOATS; i += DW_THREADS) bsm[i] = __ldg(bias + c0 * MULT + i);
    __syncthreads();                  // barrier initialised, weights staged
    if (threadIdx.x == 0) {
        pdl_wait();                   // weights above are constants; the activation tile is the predecessor's output
        mbar_expect_tx(bar, Cfg::TILE_WORDS * 4);
        tma_load_4d(tile, &tmX, c0, tx0 * S - KS / 2, ty0 * S - KS / 2, b, bar);
    }
    mbar_wait(bar, 0);
    pdl_wait();                       // (returns at once here) orders this thread's later global writes after the predecessor

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
// Tile shape / CTA size are template parameters: large maps use 16x16 tiles with 256 threads, small maps
// (<= 64x64, where a 16x16 tiling gives < 2 CTAs per SM at batch 1) use 8x16 tiles with 128 threads so that four
// independent CTAs are resident per SM and one CTA's global->smem staging overlaps another's FMA phase.
template <int TOH, int TOW>
struct MixCfgT {
    static constexpr int YH = TOH + 6, YW = TOW + 6;      // y region (3-px halo for the 7x7)
    static constexpr int XH = TOH + 8, XW = TOW + 8;      // x region (+1 more for the 3x3)
    static constexpr int XP = OddUp<XW>::value;
    static constexpr int YP = OddUp<YW>::value;
    static constexpr int X_WORDS = XH * XP * 16;
    static constexpr int Y_WORDS = YH * YP * 16;
    static constexpr int W_FLOATS = (9 + 49) * DW_CG + 2 * DW_CG;
    static constexpr size_t SMEM = (size_t)(X_WORDS + Y_WORDS + W_FLOATS) * 4 + 16 /*mbarrier*/;
};

// debug timeline (fvhd_debug_mixer_trace): 8 globaltimer stamps per CTA of the most recent mixer launch
__device__ unsigned long long* d_mix_trace = nullptr;
#define MIX_TRACE(i) do { unsigned long long* t_ = d_mix_trace; if (t_) { unsigned long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); \
    t_[(size_t)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) * 8 + (i)] = g_; } } while (0)

template <int TOH, int TOW, int NT, int SW1 = 11, int SW2 = 8, int MINB = 512 / NT>
__global__ void __launch_bounds__(NT, MINB)
repmixer_dw_kernel(const __grid_constant__ CUtensorMap tmX /*x: NHWC, box {32, XP, XH, 1}*/, bf16* __restrict__ y, bf16* __restrict__ z,
                   const float* __restrict__ w3 /*[9][C]*/, const float* __restrict__ b3,
                   const float* __restrict__ w7 /*[49][C], BN folded*/, const float* __restrict__ b7,
                   int H, int W, int C, int tiles_x) {
    using Cfg = MixCfgT<TOH, TOW>;
    extern __shared__ __align__(128) uint32_t dw_smem[];
    uint32_t* sx = dw_smem;
    uint32_t* sy = sx + Cfg::X_WORDS;
    float* w3s = reinterpret_cast<float*>(sy + Cfg::Y_WORDS);
    float* w7s = w3s + 9 * DW_CG;
    float* b3s = w7s + 49 * DW_CG;
    float* b7s = b3s + DW_CG;
    uint64_t* bar = reinterpret_cast<uint64_t*>(b7s + DW_CG);

    pdl_launch_dependents();
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * DW_CG;
    const int ty0 = (blockIdx.x / tiles_x) * TOH;
    const int tx0 = (blockIdx.x % tiles_x) * TOW;

    if (threadIdx.x == 0) {
        MIX_TRACE(0);
        tma_prefetch_desc(&tmX);
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    dw_stage_weights<9, DW_CG, NT>(w3s, w3, C, c0);
    dw_stage_weights<49, DW_CG, NT>(w7s, w7, C, c0);
    if (threadIdx.x < DW_CG) {
        b3s[threadIdx.x] = __ldg(b3 + c0 + threadIdx.x);
        b7s[threadIdx.x] = __ldg(b7 + c0 + threadIdx.x);
    }
    __syncthreads();                  // barrier initialised, weights staged
    if (threadIdx.x == 0) {
        MIX_TRACE(1);
        pdl_wait();                   // x is the predecessor's output
        MIX_TRACE(2);
        mbar_expect_tx(bar, Cfg::X_WORDS * 4);
        tma_load_4d(sx, &tmX, c0, tx0 - 4, ty0 - 4, b, bar);
    }
    mbar_wait(bar, 0);
    pdl_wait();                       // orders this thread's global writes (y, z) after the predecessor
    if (threadIdx.x == 0) MIX_TRACE(3);

    // phase 1: y = dw3x3(x) + b on the (TO+6)^2 region; zero outside the image (the 7x7's zero padding)
    {
        // strips of SW columns; when SW does not divide the region width the last strip is shifted left and recomputes a few
        // columns of its neighbour (same values, benign duplicate stores)
        constexpr int SW = SW1, STRIPS = (Cfg::YW + SW - 1) / SW;
        static_assert(Cfg::YH % 2 == 0 && SW <= Cfg::YW, "phase-1 strip shape");
        constexpr int ITEMS = 16 * Cfg::YH * STRIPS;
        for (int it = threadIdx.x; it < ITEMS; it += NT) {
            const int cp = it & 15;
            const int t = it >> 4;
            const int sub = t & 1;
            const int u = t >> 1;
            const int strip = u % STRIPS;
            const int ry = (u / STRIPS) * 2 + sub;            // row in the y region
            const int rx0 = (strip * SW + SW <= Cfg::YW) ? strip * SW : Cfg::YW - SW;
            float acc[SW][2];
#pragma unroll
            for (int j = 0; j < SW; ++j) { acc[j][0] = b3s[cp * 2]; acc[j][1] = b3s[cp * 2 + 1]; }
            dw_strip<3, 1, 1, SW>(sx + (ry * Cfg::XP + rx0) * 16 + cp, Cfg::XP * 16, w3s, cp, acc);
            const int gy = ty0 - 3 + ry;
            const bool row_in = gy >= 0 && gy < H;
            const bool row_center = ry >= 3 && ry < 3 + TOH;
#pragma unroll
            for (int j = 0; j < SW; ++j) {
                const int rx = rx0 + j;
                const int gx = tx0 - 3 + rx;
                const bool in_img = row_in && gx >= 0 && gx < W;
                const uint32_t pk = in_img ? pack_bf16x2(acc[j][0], acc[j][1]) : 0u;
                sy[(ry * Cfg::YP + rx) * 16 + cp] = pk;
                if (in_img && row_center && rx >= 3 && rx < 3 + TOW)
                    *reinterpret_cast<uint32_t*>(y + (((size_t)b * H + gy) * W + gx) * C + c0 + cp * 2) = pk;
            }
        }
    }
    if (threadIdx.x == 0) MIX_TRACE(4);
    __syncthreads();
    if (threadIdx.x == 0) MIX_TRACE(5);

    // phase 2: z = dw7x7(y) (BN folded) on the output tile
    {
        constexpr int SW = SW2, STRIPS = TOW / SW;
        static_assert(TOW % SW == 0 && TOH % 2 == 0, "phase-2 strip shape");
        constexpr int ITEMS = 16 * TOH * STRIPS;
        for (int it = threadIdx.x; it < ITEMS; it += NT) {
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
            dw_strip<7, 1, 1, SW, (MINB * NT > 512 ? 1 : 7)>(sy + (oy * Cfg::YP + ox0) * 16 + cp, Cfg::YP * 16, w7s, cp, acc);
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
    if (threadIdx.x == 0) MIX_TRACE(6);
    if (threadIdx.x == NT - 1) MIX_TRACE(7);
}

}  // namespace fvhd
