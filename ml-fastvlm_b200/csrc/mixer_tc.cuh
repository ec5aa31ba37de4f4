// Fused RepMixer depthwise pair with the 7x7 on the tensor cores (mci.py:806-853 RepMixer, :922-926 ConvFFN.conv):
//     y = dw3x3(x) + b3        (RepMixer, identity + BN branches folded by the packer)     -> global (block residual)
//     z = dw7x7(y) + b7        (ConvFFN.conv, BN folded)                                    -> global (fc1's operand)
//
// The 7x7 is 84 % of the multiply-adds and, on the FMA pipes, ~2300 instructions per thread.  A depthwise convolution has no
// channel reduction, but along one image row it IS a matrix product with a banded Toeplitz matrix:
//
//     z[c][y][x] = sum_ky  sum_k  Y_c[y + ky][k] * T_{c,ky}[k][x],      T_{c,ky}[k][x] = w7[c][ky][k - x]  (0 <= k - x <= 6, else 0)
//
// so per channel and ky the 16 x 16 output tile is one  [16 rows x 24 cols] x [24 x 16]  product on mma.sync.m16n8k8
// (f16 in, fp32 accumulate).  Of its (k8-block, n-tile) pairs only the four with k0 - x0 in {0, 8} meet the 7-wide band; their
// B fragments are TWO table words per lane and ky (the band is shift-invariant: a fragment register is the pair
// (w[d], w[d+1]) with d fixed by the lane).  Per channel: 14 ldmatrix.x4 + 14 predicated LDS + 28 mma.m16n8k8 instead of ~800
// FFMA-pipe instructions (the tensor pipe does 3x the useful MACs).
//
//   x tile (24 x 24 px x 32 ch, NHWC)  --TMA 4-D box, zero OOB fill-->  smem
//   phase 1 (FMA pipes, as dwconv.cuh): y = dw3x3(x) on the 22 x 22 region; bf16 y -> global (the block's residual), and f16 y
//            (saturating; 3 more mantissa bits than bf16) -> per-channel PLANES [c][row][col] in smem
//   phase 2 (tensor cores): warp w owns channels w, w+8, w+16, w+24 of the group; accumulators -> +b7 -> bf16 -> NHWC staging
//   write-out: staging -> global, 64 contiguous bytes per pixel
#pragma once
#include "dwconv.cuh"
#include "stem_attn_se.cuh"

namespace fvhd {

struct MixTc {
    static constexpr int TOH = 16, TOW = 16, NT = 256;
    static constexpr int YH = TOH + 6, YW = TOW + 6;
    static constexpr int XH = TOH + 8, XW = TOW + 8;
    static constexpr int XP = OddUp<XW>::value;           // 25: pixel pitch of a tile row (TMA box width)
    static constexpr int X_WORDS = XH * XP * 16;          // 9600
    static constexpr int YROW = 40;                       // halfs per plane row: 32 (K of the product, cols >= 22 are zero) + 8 pad
                                                          // -> 80-B rows: 16-B aligned and conflict-free for ldmatrix
    static constexpr int PLANE = YH * YROW + 8;           // 888 halfs = 1776 B per channel (16-B multiple; +8 spreads the planes over banks)
    static constexpr int YP_WORDS = DW_CG * PLANE / 2;    // 14208
    static constexpr int ZPIX = 17;                       // words per pixel of the z staging tile (aliases the x tile)
    static constexpr int PTAB_PITCH = DW_CG + 1;          // pair table [ky][i][c]: (w[i-1], w[i]) as f16x2, zero outside taps 0..6;
    static constexpr int PTAB_WORDS = 7 * 8 * PTAB_PITCH; // pitch 33 keeps both the build (lanes = c) and the reads (lanes = i) conflict-free
    static constexpr int W_FLOATS = 9 * DW_CG + 2 * DW_CG;
    static constexpr size_t SMEM = (size_t)(X_WORDS + YP_WORDS + PTAB_WORDS + W_FLOATS) * 4 + 16;
    static_assert(TOH * TOW * ZPIX <= X_WORDS, "z staging must fit in the x tile");
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_f16_1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ unsigned short f16_bits_sat(float v) {   // fp32 -> f16 bits, saturating instead of inf
    unsigned short h;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
    return h;
}

__global__ void __launch_bounds__(MixTc::NT, 2)
repmixer_tc_kernel(const __grid_constant__ CUtensorMap tmX /*x: NHWC, box {32, XP, XH, 1}*/, bf16* __restrict__ y, bf16* __restrict__ z,
                   const float* __restrict__ w3 /*[9][C]*/, const float* __restrict__ b3,
                   const float* __restrict__ w7 /*[49][C], BN folded*/, const float* __restrict__ b7,
                   int H, int W, int C, int tiles_x, int batch) {
    using Cfg = MixTc;
    constexpr int NT = Cfg::NT, TOH = Cfg::TOH, TOW = Cfg::TOW;
    extern __shared__ __align__(128) uint32_t dw_smem[];
    uint32_t* sx = dw_smem;                                            // x tile; later the z staging tile
    __half* yp = reinterpret_cast<__half*>(sx + Cfg::X_WORDS);        // y planes
    uint32_t* ptab = sx + Cfg::X_WORDS + Cfg::YP_WORDS;
    float* w3s = reinterpret_cast<float*>(ptab + Cfg::PTAB_WORDS);
    float* b3s = w3s + 9 * DW_CG;
    float* b7s = b3s + DW_CG;
    uint64_t* bar = reinterpret_cast<uint64_t*>(b7s + DW_CG);

    pdl_launch_dependents();
    // persistent over the spatial tiles of ONE 32-channel group: the constants (3x3 taps, biases, 7x7 pair table) are staged once
    // per CTA; blockIdx.x walks items (image, tile) = blockIdx.x, blockIdx.x + gridDim.x, ...
    const int c0 = blockIdx.y * DW_CG;
    const int tiles_per_img = tiles_x * ((H + TOH - 1) / TOH);
    const int n_items = tiles_per_img * batch;

    if (threadIdx.x == 0) {
        MIX_TRACE(0);
        tma_prefetch_desc(&tmX);
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    // constants (weights, never written by a kernel of the forward): staged before the PDL wait
    dw_stage_weights<9, DW_CG, NT>(w3s, w3, C, c0);
    if (threadIdx.x < DW_CG) {
        b3s[threadIdx.x] = __ldg(b3 + c0 + threadIdx.x);
        b7s[threadIdx.x] = __ldg(b7 + c0 + threadIdx.x);
    }
    float* w7s = reinterpret_cast<float*>(sx);                        // fp32 taps parked in the x tile until the TMA load is issued
    dw_stage_weights<49, DW_CG, NT>(w7s, w7, C, c0);
    {   // zero the planes: the K padding (cols 22..31) must be finite for the zero band entries it meets
        uint4* p4 = reinterpret_cast<uint4*>(yp);
        for (int i = threadIdx.x; i < Cfg::YP_WORDS / 4; i += NT) p4[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();                  // barrier initialised, constants staged
    for (int idx = threadIdx.x; idx < 56 * DW_CG; idx += NT) {
        const int c = idx & (DW_CG - 1), r = idx / DW_CG, i = r & 7, ky = r >> 3;
        const float lo = i >= 1 ? w7s[(ky * 7 + i - 1) * DW_CG + c] : 0.f;
        const float hi = i <= 6 ? w7s[(ky * 7 + i) * DW_CG + c] : 0.f;
        ptab[r * Cfg::PTAB_PITCH + c] = pack_f16x2_sat(lo, hi);
    }
    __syncthreads();                  // pair table built; the x tile region may be overwritten by the TMA load
    pdl_wait();                       // x is the predecessor's output; also orders this thread's global writes (y, z) after it
    uint32_t xphase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int b = item / tiles_per_img;
    const int tl_ = item - b * tiles_per_img;
    const int ty0 = (tl_ / tiles_x) * TOH;
    const int tx0 = (tl_ % tiles_x) * TOW;
    if (threadIdx.x == 0) {
        MIX_TRACE(2);
        mbar_expect_tx(bar, Cfg::X_WORDS * 4);
        tma_load_4d(sx, &tmX, c0, tx0 - 4, ty0 - 4, b, bar);
    }
    mbar_wait(bar, xphase);
    xphase ^= 1u;
    if (threadIdx.x == 0) MIX_TRACE(3);

    // phase 1: y = dw3x3(x) + b on the 22 x 22 region; zero outside the image (the 7x7's zero padding)
    {
        constexpr int SW = 11, STRIPS = Cfg::YW / SW;
        constexpr int ITEMS = 16 * Cfg::YH * STRIPS;
        for (int it = threadIdx.x; it < ITEMS; it += NT) {
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
            dw_strip<3, 1, 1, SW>(sx + (ry * Cfg::XP + rx0) * 16 + cp, Cfg::XP * 16, w3s, cp, acc);
            const int gy = ty0 - 3 + ry;
            const bool row_in = gy >= 0 && gy < H;
            const bool row_center = ry >= 3 && ry < 3 + TOH;
            // y leaves in two forms: bf16 (rounded, NHWC) to global for the centre pixels -- the block's residual -- and f16 into
            // the per-channel planes for the 7x7 (f16 keeps 3 more mantissa bits than the bf16 the FMA-pipe kernel fed its 7x7)
            float m[SW];
#pragma unroll
            for (int j = 0; j < SW; ++j) {
                const int gx = tx0 - 3 + rx0 + j;
                m[j] = (row_in && gx >= 0 && gx < W) ? 1.f : 0.f;
            }
            if (row_in && row_center) {
#pragma unroll
                for (int j = 0; j < SW; ++j) {
                    const int rx = rx0 + j;
                    if (m[j] != 0.f && rx >= 3 && rx < 3 + TOW)
                        *reinterpret_cast<uint32_t*>(y + (((size_t)b * H + gy) * W + (tx0 - 3 + rx)) * C + c0 + cp * 2) = pack_bf16x2(acc[j][0], acc[j][1]);
                }
            }
            unsigned short* p0 = reinterpret_cast<unsigned short*>(yp) + (2 * cp) * Cfg::PLANE + ry * Cfg::YROW + rx0;
            unsigned short* p1 = p0 + Cfg::PLANE;
            // strip 0 covers cols 0..10, strip 1 cols 11..21: pairs start at even columns; cols 10 / 11 are single halfword stores
            if (strip == 0) {
#pragma unroll
                for (int j = 0; j < 10; j += 2) {
                    *reinterpret_cast<uint32_t*>(p0 + j) = pack_f16x2_sat(acc[j][0] * m[j], acc[j + 1][0] * m[j + 1]);
                    *reinterpret_cast<uint32_t*>(p1 + j) = pack_f16x2_sat(acc[j][1] * m[j], acc[j + 1][1] * m[j + 1]);
                }
                p0[10] = f16_bits_sat(acc[10][0] * m[10]);
                p1[10] = f16_bits_sat(acc[10][1] * m[10]);
            } else {
                p0[0] = f16_bits_sat(acc[0][0] * m[0]);
                p1[0] = f16_bits_sat(acc[0][1] * m[0]);
#pragma unroll
                for (int j = 1; j < 11; j += 2) {
                    *reinterpret_cast<uint32_t*>(p0 + j) = pack_f16x2_sat(acc[j][0] * m[j], acc[j + 1][0] * m[j + 1]);
                    *reinterpret_cast<uint32_t*>(p1 + j) = pack_f16x2_sat(acc[j][1] * m[j], acc[j + 1][1] * m[j + 1]);
                }
            }
        }
    }
    if (threadIdx.x == 0) MIX_TRACE(4);
    __syncthreads();                  // planes complete; the x tile is dead from here on
    if (threadIdx.x == 0) MIX_TRACE(5);

    // phase 2: z = dw7x7(y) + b7 on the tensor cores, one channel per warp at a time
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int t = lane & 3, g = lane >> 2;
        // ldmatrix.x4 row addresses: matrices (rows 0-7, k 0-7), (rows 8-15, k 0-7), (rows 0-7, k 8-15), (rows 8-15, k 8-15)
        const int a_off = ((lane & 7) + ((lane >> 3) & 1) * 8) * Cfg::YROW + (lane >> 4) * 8;
        const int i1 = 2 * t - g + 1;          // table index of the pair starting at d = 2t - g      (valid: 0..7)
        const int i2 = i1 + 8;                 //                               ... at d = 2t - g + 8
        const bool v1ok = i1 >= 0, v2ok = i2 <= 7;
        unsigned short* zs = reinterpret_cast<unsigned short*>(sx);
#pragma unroll 2
        for (int cc = 0; cc < DW_CG / 8; ++cc) {
            const int c = warp + cc * 8;
            const __half* plane = yp + c * Cfg::PLANE + a_off;
            const uint32_t* tab = ptab + c;
            float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};     // output cols 0-7 / 8-15
#pragma unroll
            for (int ky = 0; ky < 7; ++ky) {
                uint32_t a0[4], a1[4];
                ldmatrix_x4(a0, plane + ky * Cfg::YROW);          // k = y-region cols 0..15  (regs 0,1: k 0-7; regs 2,3: k 8-15)
                ldmatrix_x4(a1, plane + ky * Cfg::YROW + 16);     // k = 16..31               (only k 16-23 meets a non-zero band block)
                const uint32_t v1 = v1ok ? tab[(ky * 8 + i1) * Cfg::PTAB_PITCH] : 0u;   // B block with k0 - x0 = 0
                const uint32_t v2 = v2ok ? tab[(ky * 8 + i2) * Cfg::PTAB_PITCH] : 0u;   // B block with k0 - x0 = 8
                // k8 blocks: (k0, x0) with k0 - x0 in {0, 8} are the only ones the 7-wide band touches
                mma_f16_1688(acc0, a0[0], a0[1], v1);             // k 0-7,   x 0-7
                mma_f16_1688(acc0, a0[2], a0[3], v2);             // k 8-15,  x 0-7
                mma_f16_1688(acc1, a0[2], a0[3], v1);             // k 8-15,  x 8-15
                mma_f16_1688(acc1, a1[0], a1[1], v2);             // k 16-23, x 8-15
            }
            const float bias = b7s[c];
            // accumulator (row g / g+8, cols 2t, 2t+1 [+8]) -> bf16 -> staging [pixel][channel], pixel pitch 17 words
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
                const int row = g + hrow * 8;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = 2 * t + e;
                    zs[((row * TOW + col) * Cfg::ZPIX) * 2 + c] = __bfloat16_as_ushort(__float2bfloat16_rn(acc0[hrow * 2 + e] + bias));
                    zs[((row * TOW + col + 8) * Cfg::ZPIX) * 2 + c] = __bfloat16_as_ushort(__float2bfloat16_rn(acc1[hrow * 2 + e] + bias));
                }
            }
        }
    }
    __syncthreads();
    // write-out: 16 channel-pair words per pixel, two pixels per warp instruction
    for (int i = threadIdx.x; i < TOH * TOW * 16; i += NT) {
        const int pix = i >> 4, cpw = i & 15;
        const int gy = ty0 + (pix >> 4), gx = tx0 + (pix & 15);
        if (gy < H && gx < W)
            *reinterpret_cast<uint32_t*>(z + (((size_t)b * H + gy) * W + gx) * C + c0 + cpw * 2) = sx[pix * Cfg::ZPIX + cpw];
    }
    if (threadIdx.x == 0) MIX_TRACE(6);
    if (threadIdx.x == NT - 1) MIX_TRACE(7);
    __syncthreads();                  // the staging tile (aliasing the x tile) has been read: the next item's TMA load may land
    }   // item loop
}

}  // namespace fvhd
