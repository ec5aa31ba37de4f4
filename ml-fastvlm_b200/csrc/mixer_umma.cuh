// RepMixer depthwise pair on the 5th-gen tensor cores (tcgen05 / TMEM), mci.py:806-853 (RepMixer) + :921 (ConvFFN.conv):
//     y = dw3x3(x) + b3        (identity + BN branches folded by the packer)     -> global (block residual)
//     z = dw7x7(y) + b7        (BN folded)                                        -> global (fc1's operand)
//
// A depthwise conv has no channel reduction, but ONE TAP of it is a GEMM with a diagonal matrix:
//     out[p, c'] += sum_c  in[p + shift(tap), c] * diag(w[tap])[c, c']
// i.e. tcgen05.mma with  A = 128 consecutive pixels x 16 channels (K-major: the NHWC tile itself, read at a shifted start
// address -- no im2col),  B = 16 x 16 diagonal of that tap's 16 weights,  D = 128 x 16 fp32 accumulator in TMEM.  15/16 of the
// multiplies hit zeros, yet at 4096 MAC/clk/SM the 58 taps cost 58 x 8 = 464 tensor cycles per 2048 pixel-channels -- 4x
// faster than the 128 FMA lanes can do the 58 useful MACs, with the CUDA cores left free for the epilogues.
//
// Layout trick: the tile lives in smem as [8-channel chunk][pixel (row-major, pitch P = 40)][8 ch] -- the canonical no-swizzle
// K-major core-matrix layout (8 rows x 16 B contiguous, SBO = 128 B between 8-pixel groups, LBO = chunk-plane stride).  A tap
// (ky, kx) is then a start-address offset of (ky * P + kx) * 16 B, and an M-tile is ANY 128 consecutive pixels of the linearised
// tile: the conv is evaluated on the flattened tile, columns >= P - (k - 1) of each row are garbage and never stored.
//   x tile 24 x 40 px (+4 halo)  --TMA 5-D box {8 ch, 40, 24, 2 chunks, 1}, zero OOB fill = the 3x3's zero padding
//   y  = 7 M-tiles x 9 taps  -> TMEM -> +b3, zero outside the image (the 7x7's zero padding) -> bf16 -> smem (same layout) and,
//        for the 16 x 32 centre, global (the block residual)
//   z  = 5 M-tiles x 49 taps -> TMEM -> +b7 -> bf16 -> global
// One work item = 16 output rows x 32 output cols x 16 channels: 308 MMAs (2464 tensor cycles).  Persistent CTAs (6 warps:
// TMA, MMA, 4 epilogue), two per SM; a CTA keeps one 16-channel group (its 58 diagonal B tiles are built once, before the PDL
// wait) and walks that group's spatial tiles; the x load of item i+1 overlaps everything after the y-MMAs of item i.
#pragma once
#include "gemm_tcgen05.cuh"

namespace fvhd {

struct MixU {
    static constexpr int TOH = 16, TOW = 32;            // output tile
    static constexpr int P = 40;                        // pixel pitch of the linearised tile (= TOW + 8)
    static constexpr int XH = TOH + 8;                  // 24 input rows
    static constexpr int CG = 16;                       // channels per work item (one MMA: K = N = 16)
    static constexpr int YT = 7;                        // y M-tiles: positions [0, 896) cover rows 0..21 (22 * 40 = 880)
    static constexpr int ZT = 5;                        // z M-tiles: positions [0, 640) = 16 rows * 40
    static constexpr int X_PLANE = XH * P * 16;         // 15360 B per 8-channel chunk (dense TMA box)
    static constexpr int X_BYTES = 2 * X_PLANE + 512;   // + slack: the last y M-tile reads up to position 977 of each plane
    static constexpr int Y_PLANE = YT * 128 * 16;       // 14336 B
    static constexpr int Y_BYTES = 2 * Y_PLANE;
    static constexpr int B_TAP = 512;                   // one 16 x 16 bf16 diagonal tile (4 core matrices)
    static constexpr int B_BYTES = 58 * B_TAP;          // taps 0..8: 3x3, 9..57: 7x7
    static constexpr int THREADS = 192;
    static constexpr int TMEM_COLS = 256;               // y acc at [0, 112), z acc at [128, 208)
    static constexpr int ZACC_COL = 128;
    static constexpr size_t SMEM = (size_t)X_BYTES + Y_BYTES + B_BYTES + 2 * CG * 4 + 128 /*barriers*/ + 1024 /*align*/;
    static_assert((YT * 128 - 1) + 2 * P + 2 < 2 * XH * P + 32, "y M-tiles read inside the x planes + slack");
    static_assert((ZT * 128 - 1) + 6 * P + 6 < YT * 128, "z M-tiles read inside the y planes");
};

// No-swizzle K-major smem descriptor: 8-row x 16-B core matrices; LBO = byte distance between the two K core matrices of one
// MMA (K = 16 bf16 = 2 x 16 B), SBO = byte distance between consecutive 8-row groups.
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// 5-D tiled load: coords {ch-in-chunk, x, y, chunk, image}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

struct MixUParams {
    bf16* y;                 // [B, H, W, C]
    bf16* z;
    const float* w3;         // [9][C]
    const float* b3;         // [C]
    const float* w7;         // [49][C]
    const float* b7;
    int B, H, W, C;
    int tiles_x, tiles_y;    // spatial tiles per image
    int groups;              // C / 16
    int ctas_per_group;      // gridDim.x / groups
};

__global__ void __launch_bounds__(MixU::THREADS, 2)
repmixer_umma_kernel(const __grid_constant__ CUtensorMap tmX, const MixUParams p) {
    using U = MixU;
    extern __shared__ uint8_t mixu_smem_raw[];
    const uint32_t raw_addr = smem_u32(mixu_smem_raw);
    uint8_t* smem = mixu_smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* sx = smem;
    uint8_t* sy = sx + U::X_BYTES;
    uint8_t* sb = sy + U::Y_BYTES;
    float* sb3 = reinterpret_cast<float*>(sb + U::B_BYTES);
    float* sb7 = sb3 + U::CG;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sb7 + U::CG);
    uint64_t* x_full = bars;          // TMA -> MMA
    uint64_t* x_free = bars + 1;      // y-MMAs retired -> TMA may overwrite x
    uint64_t* yacc_full = bars + 2;   // y-MMAs retired -> epilogue
    uint64_t* yacc_free = bars + 3;   // epilogue drained y acc (4 warps)
    uint64_t* ysm_full = bars + 4;    // epilogue wrote y smem (4 warps) -> MMA
    uint64_t* zacc_full = bars + 5;   // z-MMAs retired -> epilogue (y smem is free again, too)
    uint64_t* zacc_free = bars + 6;   // epilogue drained z acc (4 warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

    pdl_launch_dependents();
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);     // broadcast: lets ptxas treat role branches as warp-uniform
    const int lane = threadIdx.x & 31;
    const int grp = (int)blockIdx.x % p.groups;                 // this CTA's 16-channel group
    const int cta_in_grp = (int)blockIdx.x / p.groups;
    const int c0 = grp * U::CG;
    const int tiles_img = p.tiles_x * p.tiles_y;
    const int n_sp = p.B * tiles_img;                           // spatial tiles of the whole batch

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmX);
        mbar_init(x_full, 1); mbar_init(x_free, 1);
        mbar_init(yacc_full, 1); mbar_init(yacc_free, 4);
        mbar_init(ysm_full, 4);
        mbar_init(zacc_full, 1); mbar_init(zacc_free, 4);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, U::TMEM_COLS);
        tmem_relinquish();
    }
    // ---- constants (never written by a kernel of the forward): diagonal B tiles of the 58 taps, biases; before the PDL wait
    {
        uint4* b4 = reinterpret_cast<uint4*>(sb);
        for (int i = threadIdx.x; i < U::B_BYTES / 16; i += U::THREADS) b4[i] = make_uint4(0, 0, 0, 0);
        if (threadIdx.x < U::CG) {
            sb3[threadIdx.x] = __ldg(p.b3 + c0 + threadIdx.x);
            sb7[threadIdx.x] = __ldg(p.b7 + c0 + threadIdx.x);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 58 * U::CG; i += U::THREADS) {
        const int tap = i / U::CG, c = i % U::CG;
        const float w = tap < 9 ? __ldg(p.w3 + (size_t)tap * p.C + c0 + c) : __ldg(p.w7 + (size_t)(tap - 9) * p.C + c0 + c);
        // B[n = c][k = c] of the K-major 16 x 16 tile: n-group (c >> 3) * 256 B, k-chunk (c >> 3) * 128 B, row (c & 7) * 16 B, elem (c & 7) * 2 B
        const int off = tap * U::B_TAP + (c >> 3) * 256 + (c >> 3) * 128 + (c & 7) * 16 + (c & 7) * 2;
        *reinterpret_cast<bf16*>(sb + off) = __float2bfloat16_rn(w);
    }
    fence_proxy_async_smem();                    // generic-proxy writes of B -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer: one 5-D box per item (both 8-channel chunk planes)
            pdl_wait();                          // x is the predecessor's output
            int it = 0;
            for (int s = cta_in_grp; s < n_sp; s += p.ctas_per_group, ++it) {
                const int b = s / tiles_img, t = s - b * tiles_img;
                const int ty0 = (t / p.tiles_x) * U::TOH, tx0 = (t % p.tiles_x) * U::TOW;
                mbar_wait(x_free, ((uint32_t)it & 1u) ^ 1u);
                mbar_expect_tx(x_full, 2 * U::X_PLANE);
                tma_load_5d(sx, &tmX, 0, tx0 - 4, ty0 - 4, grp * 2, b, x_full);
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer: the WHOLE warp runs the loop (converged, uniform operands); one elected lane issues
        const uint32_t idesc = umma_idesc_bf16(128, U::CG);
        const uint32_t sx_a = smem_u32(sx), sy_a = smem_u32(sy), sb_a = smem_u32(sb);
        const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint64_t hi_x = ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);          // SBO = 128 B, version 1, no swizzle
        const uint64_t hi_b = ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);          // SBO = 256 B
        const uint64_t dx0 = hi_x | ((uint64_t)(U::X_PLANE >> 4) << 16) | (uint64_t)(sx_a >> 4);
        const uint64_t dy0 = hi_x | ((uint64_t)(U::Y_PLANE >> 4) << 16) | (uint64_t)(sy_a >> 4);
        const uint64_t db0 = hi_b | ((uint64_t)(128 >> 4) << 16) | (uint64_t)(sb_a >> 4);
        int it = 0;
        for (int s = cta_in_grp; s < n_sp; s += p.ctas_per_group, ++it) {
            const uint32_t ph = (uint32_t)it & 1u;
            mbar_wait(x_full, ph);
            mbar_wait(yacc_free, ph ^ 1u);
            tc_fence_after();
#pragma unroll 1
            for (int m = 0; m < U::YT; ++m) {
                const uint64_t da_m = dx0 + (uint64_t)(128 * m);                         // start address advances in 16-B units
                const uint32_t acc = tm + (uint32_t)(m * U::CG);
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const int ky = tap / 3, kx = tap % 3;
                    if (elect_one())
                        umma_bf16(acc, da_m + (uint64_t)(ky * U::P + kx), db0 + (uint64_t)(tap * (U::B_TAP >> 4)), idesc, tap != 0 ? 1u : 0u);
                }
            }
            if (elect_one()) {
                umma_commit(x_free);             // x may be overwritten by the next item's load
                umma_commit(yacc_full);
            }
            __syncwarp();
            mbar_wait(ysm_full, ph);             // epilogue has written y (bf16) into smem
            mbar_wait(zacc_free, ph ^ 1u);
            tc_fence_after();
#pragma unroll 1
            for (int m = 0; m < U::ZT; ++m) {
                const uint64_t da_m = dy0 + (uint64_t)(128 * m);
                const uint32_t acc = tm + (uint32_t)(U::ZACC_COL + m * U::CG);
#pragma unroll 1
                for (int ky = 0; ky < 7; ++ky) {
#pragma unroll
                    for (int kx = 0; kx < 7; ++kx) {
                        if (elect_one())
                            umma_bf16(acc, da_m + (uint64_t)(ky * U::P + kx), db0 + (uint64_t)((9 + ky * 7 + kx) * (U::B_TAP >> 4)), idesc,
                                      (ky | kx) != 0 ? 1u : 0u);
                    }
                }
            }
            if (elect_one()) umma_commit(zacc_full);
            __syncwarp();
        }
    } else {
        // ---------------- epilogue warps: TMEM lane quarter q <-> positions 32 q + lane of every M-tile
        const int q = warp & 3;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        pdl_wait();                              // global writes (y, z) come after the predecessor
        int it = 0;
        for (int s = cta_in_grp; s < n_sp; s += p.ctas_per_group, ++it) {
            const uint32_t ph = (uint32_t)it & 1u;
            const int b = s / tiles_img, t = s - b * tiles_img;
            const int ty0 = (t / p.tiles_x) * U::TOH, tx0 = (t % p.tiles_x) * U::TOW;
            const size_t img_off = (size_t)b * p.H * p.W * p.C;
            // ---- y: +b3, zero outside the image, bf16 -> smem planes (7x7 operand) and centre -> global (residual)
            mbar_wait(yacc_full, ph);
            tc_fence_after();
#pragma unroll 1
            for (int m = 0; m < U::YT; ++m) {
                uint32_t r[16];
                tmem_ld16(lane_base + (uint32_t)(m * U::CG), r);
                tmem_ld_wait();
                const int pos = 128 * m + 32 * q + lane;
                const int yr = pos / U::P, yc = pos - yr * U::P;
                const int gy = ty0 - 3 + yr, gx = tx0 - 3 + yc;
                const bool in_img = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
                uint32_t o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float v0 = in_img ? __uint_as_float(r[2 * j]) + sb3[2 * j] : 0.f;
                    const float v1 = in_img ? __uint_as_float(r[2 * j + 1]) + sb3[2 * j + 1] : 0.f;
                    o[j] = pack_bf16x2(v0, v1);
                }
                *reinterpret_cast<uint4*>(sy + (size_t)pos * 16) = make_uint4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<uint4*>(sy + U::Y_PLANE + (size_t)pos * 16) = make_uint4(o[4], o[5], o[6], o[7]);
                if (in_img && yr >= 3 && yr < 3 + U::TOH && yc >= 3 && yc < 3 + U::TOW) {
                    uint4* dst = reinterpret_cast<uint4*>(p.y + img_off + ((size_t)gy * p.W + gx) * p.C + c0);
                    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
                    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
                }
            }
            fence_proxy_async_smem();            // y planes are read by the tensor core
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(ysm_full); mbar_arrive(yacc_free); }
            // ---- z: +b7 -> bf16 -> global
            mbar_wait(zacc_full, ph);
            tc_fence_after();
#pragma unroll 1
            for (int m = 0; m < U::ZT; ++m) {
                uint32_t r[16];
                tmem_ld16(lane_base + (uint32_t)(U::ZACC_COL + m * U::CG), r);
                tmem_ld_wait();
                const int pos = 128 * m + 32 * q + lane;
                const int zr = pos / U::P, zc = pos - zr * U::P;
                const int gy = ty0 + zr, gx = tx0 + zc;
                if (zc < U::TOW && gy < p.H && gx < p.W) {
                    uint32_t o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        o[j] = pack_bf16x2(__uint_as_float(r[2 * j]) + sb7[2 * j], __uint_as_float(r[2 * j + 1]) + sb7[2 * j + 1]);
                    uint4* dst = reinterpret_cast<uint4*>(p.z + img_off + ((size_t)gy * p.W + gx) * p.C + c0);
                    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
                    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(zacc_free);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, U::TMEM_COLS);
    }
}

}  // namespace fvhd
