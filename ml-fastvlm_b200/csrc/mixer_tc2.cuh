// RepMixer depthwise pair, third implementation: BOTH convolutions on the tensor cores (mma.sync m16n8k16, f16 in, fp32
// accumulate), 16 channels per CTA (mci.py:806-853 RepMixer, :921 ConvFFN.conv):
//     y = dw3x3(x) + b3   -> global (block residual)          z = dw7x7(y) + b7   -> global (fc1's operand)
//
// Why: mixer_tc.cuh ran the 3x3 on the FMA pipes and spent 22 k warp-instructions per 16 x 16 x 32 tile (ncu: issue-bound, 30 %
// integer/address work, 2-byte staging stores); the tcgen05 variant (mixer_umma.cuh) is bound by the tensor core's smem
// A-operand reads.  Here a row of a depthwise conv is a product with a banded Toeplitz matrix (as in mixer_tc.cuh) for BOTH
// convs, so per channel the work is 18 + 14 HMMA.16816 and as many ldmatrix.x4; everything else is layout:
//   x tile (24 x 24 px x 16 ch, NHWC bf16)  --TMA 4-D box, zero OOB fill = the 3x3's zero padding-->  smem
//   transpose: NHWC bf16 -> per-channel PLANES [c][row][col] f16 (two pixels x two channels per 4 instructions; bf16 -> f16 is
//              exact inside the f16 range, saturating outside)
//   phase 1: warp w owns channels 2w, 2w+1.  y = dw3x3(x) on the 22 x 22 region as 2 x 3 tiles of 16 x 8: per tile and ky one
//            ldmatrix.x4 (A = 16 rows x 16 input columns of the x plane) + one mma (B = two table words per lane: the band is
//            shift-invariant).  Accumulators -> +b3, zero outside the image -> f16 y plane (the 7x7's operand) and, for the
//            centre, a bf16 staging tile [channel pair][row][col] with BOTH channels of the warp packed into one 32-bit store.
//   phase 2: z = dw7x7(y): 2 tiles x 7 ky, same scheme, reads only planes this warp wrote (no block barrier in between).
//   write-out: the two staging tiles are re-read four channel pairs at a time and leave as 16-byte stores (32 contiguous
//            bytes per pixel).
#pragma once
#include "mixer_tc.cuh"

namespace fvhd {

struct MixT2 {
    static constexpr int TOH = 16, TOW = 16, NT = 256, CG = 16;
    static constexpr int XH = TOH + 8, XW = TOW + 8;             // 24 x 24 input tile (halo 4)
    static constexpr int XP = 25;                                 // TMA box width (odd pitch, as the other dw kernels)
    static constexpr int YH = TOH + 6;                            // 22 rows / cols of y
    static constexpr int X_BYTES = XH * XP * CG * 2;              // 19200: NHWC tile; later the y / z staging tiles (2 x 8 KB)
    static constexpr int ROW = 40;                                // halfs per plane row (80 B: 16-B aligned, conflict-free ldmatrix)
    static constexpr int XPL = XH * ROW + 8;                      // halfs per x plane (1936 B; +8 spreads the planes over banks)
    static constexpr int YPL = YH * ROW + 8;                      // halfs per y plane (1776 B)
    static constexpr int TPITCH = CG + 1;                         // pair tables [ky][i][c], pitch 17 words
    static constexpr int T7_WORDS = 7 * 8 * TPITCH, T3_WORDS = 3 * 8 * TPITCH;
    static constexpr size_t SMEM = (size_t)X_BYTES + (size_t)CG * (XPL + YPL) * 2 + (size_t)(T7_WORDS + T3_WORDS) * 4 + 2 * CG * 4 + 16 + 128;
    static constexpr int SRP = 18;                                // staging row pitch in words: [channel pair][row][col], 2-way conflicts at most
    static constexpr int SPL = TOH * SRP;                         // words per channel-pair plane of a staging tile
    static constexpr int STILE = 8 * SPL;                         // words per staging tile (y or z)
    static_assert(2 * STILE * 4 <= X_BYTES, "y and z staging tiles must fit in the x tile");
    static_assert((XPL * 2) % 16 == 0 && (YPL * 2) % 16 == 0 && X_BYTES % 128 == 0, "plane alignment for ldmatrix");
};

__device__ __forceinline__ void mma_f16_16816_acc(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(MixT2::NT, 2)
repmixer_tc2_kernel(const __grid_constant__ CUtensorMap tmX /*x: NHWC, box {16, XP, XH, 1}*/, bf16* __restrict__ y, bf16* __restrict__ z,
                    const float* __restrict__ w3 /*[9][C]*/, const float* __restrict__ b3,
                    const float* __restrict__ w7 /*[49][C], BN folded*/, const float* __restrict__ b7,
                    int H, int W, int C, int tiles_x) {
    using Cfg = MixT2;
    constexpr int NT = Cfg::NT, TOH = Cfg::TOH, TOW = Cfg::TOW, CG = Cfg::CG, ROW = Cfg::ROW;
    extern __shared__ __align__(128) uint8_t mt2_smem[];
    uint32_t* sx = reinterpret_cast<uint32_t*>(mt2_smem);                       // x tile [pix][8 words]; later staging
    __half* xp = reinterpret_cast<__half*>(mt2_smem + Cfg::X_BYTES);            // x planes
    __half* yp = xp + CG * Cfg::XPL;                                            // y planes
    uint32_t* tab7 = reinterpret_cast<uint32_t*>(yp + CG * Cfg::YPL);
    uint32_t* tab3 = tab7 + Cfg::T7_WORDS;
    float* b3s = reinterpret_cast<float*>(tab3 + Cfg::T3_WORDS);
    float* b7s = b3s + CG;
    uint64_t* bar = reinterpret_cast<uint64_t*>(b7s + CG);

    pdl_launch_dependents();
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * CG;
    const int ty0 = (blockIdx.x / tiles_x) * TOH;
    const int tx0 = (blockIdx.x % tiles_x) * TOW;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmX);
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    // ---- constants (weights are never written by a kernel of the forward): pair tables + biases, before the PDL wait.
    // entry i of a table row holds (w[i-1], w[i]) as f16x2, zero outside the taps: the B fragment of lane (g, t) is entry 2t-g+1 (+8)
    for (int idx = threadIdx.x; idx < 7 * 8 * CG; idx += NT) {
        const int c = idx % CG, r = idx / CG, i = r & 7, ky = r >> 3;
        const float lo = i >= 1 ? __ldg(w7 + (size_t)(ky * 7 + i - 1) * C + c0 + c) : 0.f;
        const float hi = i <= 6 ? __ldg(w7 + (size_t)(ky * 7 + i) * C + c0 + c) : 0.f;
        tab7[r * Cfg::TPITCH + c] = pack_f16x2_sat(lo, hi);
    }
    for (int idx = threadIdx.x; idx < 3 * 8 * CG; idx += NT) {
        const int c = idx % CG, r = idx / CG, i = r & 7, ky = r >> 3;
        const float lo = (i >= 1 && i <= 3) ? __ldg(w3 + (size_t)(ky * 3 + i - 1) * C + c0 + c) : 0.f;
        const float hi = i <= 2 ? __ldg(w3 + (size_t)(ky * 3 + i) * C + c0 + c) : 0.f;
        tab3[r * Cfg::TPITCH + c] = pack_f16x2_sat(lo, hi);
    }
    if (threadIdx.x < CG) {
        b3s[threadIdx.x] = __ldg(b3 + c0 + threadIdx.x);
        b7s[threadIdx.x] = __ldg(b7 + c0 + threadIdx.x);
    }
    // x-plane columns 24..31 are multiplied by zero band entries: they must be finite -> zero them (16 B per row and channel)
    for (int idx = threadIdx.x; idx < CG * Cfg::XH; idx += NT) {
        const int c = idx / Cfg::XH, r = idx - c * Cfg::XH;
        *reinterpret_cast<uint4*>(xp + c * Cfg::XPL + r * ROW + 24) = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();                  // barrier initialised, constants staged
    if (threadIdx.x == 0) {
        pdl_wait();                   // x is the predecessor's output
        mbar_expect_tx(bar, Cfg::X_BYTES);
        tma_load_4d(sx, &tmX, c0, tx0 - 4, ty0 - 4, b, bar);
    }
    mbar_wait(bar, 0);
    pdl_wait();                       // orders this thread's global writes (y, z) after the predecessor

    // ---- transpose: NHWC bf16 tile -> f16 planes.  item = (row, column pair, channel half): 2 x LDS.128 -> 4 channel pairs
    for (int it = threadIdx.x; it < Cfg::XH * 12 * 2; it += NT) {
        const int hf = it & 1, t2 = it >> 1, j = t2 % 12, r = t2 / 12;
        const uint4 p0 = *reinterpret_cast<const uint4*>(sx + (r * Cfg::XP + 2 * j) * 8 + hf * 4);
        const uint4 p1 = *reinterpret_cast<const uint4*>(sx + (r * Cfg::XP + 2 * j + 1) * 8 + hf * 4);
        const uint32_t w0[4] = {p0.x, p0.y, p0.z, p0.w}, w1[4] = {p1.x, p1.y, p1.z, p1.w};
        __half* dst = xp + (hf * 8) * Cfg::XPL + r * ROW + 2 * j;
#pragma unroll
        for (int cp = 0; cp < 4; ++cp) {
            const float2 a = unpack_bf16x2(w0[cp]), c2 = unpack_bf16x2(w1[cp]);
            *reinterpret_cast<uint32_t*>(dst + (2 * cp) * Cfg::XPL) = pack_f16x2_sat(a.x, c2.x);        // channel 2cp:   pixels 2j, 2j+1
            *reinterpret_cast<uint32_t*>(dst + (2 * cp + 1) * Cfg::XPL) = pack_f16x2_sat(a.y, c2.y);    // channel 2cp+1
        }
    }
    __syncthreads();                  // planes complete (every thread wrote parts of every channel); the NHWC tile is dead

    const int t = lane & 3, g = lane >> 2;
    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lcol = (lane >> 4) * 8;       // ldmatrix.x4 row / column of this lane
    const int i1 = 2 * t - g + 1, i2 = i1 + 8;
    const bool v1ok = i1 >= 0, v2ok = i2 <= 7;
    const int ca = 2 * warp, cb = 2 * warp + 1;                                          // this warp's two channels
    uint32_t* ystage = sx + warp * Cfg::SPL;                                             // staging tiles [channel pair = warp][row][col], bf16x2 words
    uint32_t* zstage = sx + Cfg::STILE + warp * Cfg::SPL;
    const __half* xa = xp + ca * Cfg::XPL;
    const __half* xb = xp + cb * Cfg::XPL;
    __half* ya = yp + ca * Cfg::YPL;
    __half* yb = yp + cb * Cfg::YPL;

    // ---- phase 1: y = dw3x3(x) + b3 on the 22 x 22 region (2 x 3 tiles of 16 rows x 8 cols), both channels of the warp
    {
        uint32_t ba[3][2], bb[3][2];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            ba[ky][0] = v1ok ? tab3[(ky * 8 + i1) * Cfg::TPITCH + ca] : 0u;
            ba[ky][1] = v2ok ? tab3[(ky * 8 + i2) * Cfg::TPITCH + ca] : 0u;
            bb[ky][0] = v1ok ? tab3[(ky * 8 + i1) * Cfg::TPITCH + cb] : 0u;
            bb[ky][1] = v2ok ? tab3[(ky * 8 + i2) * Cfg::TPITCH + cb] : 0u;
        }
        const float biasa = b3s[ca], biasb = b3s[cb];
#pragma unroll 1
        for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
                float da[4] = {0.f, 0.f, 0.f, 0.f}, db[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    int row = 16 * mt + ky + lrow;
                    row = row < Cfg::XH ? row : Cfg::XH - 1;                 // rows past the tile feed discarded outputs only
                    const int off = row * ROW + 8 * nt + lcol;
                    uint32_t fa[4], fb[4];
                    ldmatrix_x4(fa, xa + off);
                    ldmatrix_x4(fb, xb + off);
                    mma_f16_16816_acc(da, fa, ba[ky][0], ba[ky][1]);
                    mma_f16_16816_acc(db, fb, bb[ky][0], bb[ky][1]);
                }
                // accumulator (rows g, g+8; cols 2t, 2t+1) of this tile
#pragma unroll
                for (int hr = 0; hr < 2; ++hr) {
                    const int yr = 16 * mt + g + 8 * hr;
                    const int yc = 8 * nt + 2 * t;
                    if (yr < Cfg::YH) {
                        const int gy = ty0 - 3 + yr, gx = tx0 - 3 + yc;
                        const bool rin = gy >= 0 && gy < H;
                        const float m0 = (rin && gx >= 0 && gx < W) ? 1.f : 0.f;          // y is ZERO outside the image (the 7x7's padding)
                        const float m1 = (rin && gx + 1 >= 0 && gx + 1 < W) ? 1.f : 0.f;
                        const float a0 = (da[2 * hr] + biasa) * m0, a1 = (da[2 * hr + 1] + biasa) * m1;
                        const float c0v = (db[2 * hr] + biasb) * m0, c1v = (db[2 * hr + 1] + biasb) * m1;
                        *reinterpret_cast<uint32_t*>(ya + yr * ROW + yc) = pack_f16x2_sat(a0, a1);
                        *reinterpret_cast<uint32_t*>(yb + yr * ROW + yc) = pack_f16x2_sat(c0v, c1v);
                        if (yr >= 3 && yr < 3 + TOH) {                                   // centre: bf16, both channels in one word
                            const int pr = (yr - 3) * Cfg::SRP;
                            if (yc >= 3 && yc < 3 + TOW) ystage[pr + yc - 3] = pack_bf16x2(a0, c0v);
                            if (yc + 1 >= 3 && yc + 1 < 3 + TOW) ystage[pr + yc - 2] = pack_bf16x2(a1, c1v);
                        }
                    }
                }
            }
        }
    }
    __syncwarp();                     // phase 2 reads only the y planes THIS warp wrote

    // ---- phase 2: z = dw7x7(y) + b7 on the 16 x 16 tile (2 tiles of 16 x 8)
    {
        const float biasa = b7s[ca], biasb = b7s[cb];
        float za[2][4], zb[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) { za[nt][e] = 0.f; zb[nt][e] = 0.f; }
#pragma unroll
        for (int ky = 0; ky < 7; ++ky) {
            const uint32_t a1w = v1ok ? tab7[(ky * 8 + i1) * Cfg::TPITCH + ca] : 0u;
            const uint32_t a2w = v2ok ? tab7[(ky * 8 + i2) * Cfg::TPITCH + ca] : 0u;
            const uint32_t b1w = v1ok ? tab7[(ky * 8 + i1) * Cfg::TPITCH + cb] : 0u;
            const uint32_t b2w = v2ok ? tab7[(ky * 8 + i2) * Cfg::TPITCH + cb] : 0u;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const int off = (ky + lrow) * ROW + 8 * nt + lcol;
                uint32_t fa[4], fb[4];
                ldmatrix_x4(fa, ya + off);
                ldmatrix_x4(fb, yb + off);
                mma_f16_16816_acc(za[nt], fa, a1w, a2w);
                mma_f16_16816_acc(zb[nt], fb, b1w, b2w);
            }
        }
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int hr = 0; hr < 2; ++hr) {
                const int pix = (g + 8 * hr) * Cfg::SRP + 8 * nt + 2 * t;
                zstage[pix] = pack_bf16x2(za[nt][2 * hr] + biasa, zb[nt][2 * hr] + biasb);
                zstage[pix + 1] = pack_bf16x2(za[nt][2 * hr + 1] + biasa, zb[nt][2 * hr + 1] + biasb);
            }
    }
    __syncthreads();
    // ---- write-out: 2 tensors x 256 px x 32 B as 16-byte stores
    for (int i = threadIdx.x; i < 2 * TOH * TOW * 2; i += NT) {
        const int hf = i & 1, pix = (i >> 1) & (TOH * TOW - 1), which = i >> 9;
        const int gy = ty0 + (pix >> 4), gx = tx0 + (pix & 15);
        if (gy < H && gx < W) {
            const uint32_t* src = sx + which * Cfg::STILE + (hf * 4) * Cfg::SPL + (pix >> 4) * Cfg::SRP + (pix & 15);
            const uint4 v = make_uint4(src[0], src[Cfg::SPL], src[2 * Cfg::SPL], src[3 * Cfg::SPL]);
            bf16* dst = (which ? z : y) + (((size_t)b * H + gy) * W + gx) * C + c0 + hf * 8;
            *reinterpret_cast<uint4*>(dst) = v;
        }
    }
}

}  // namespace fvhd
