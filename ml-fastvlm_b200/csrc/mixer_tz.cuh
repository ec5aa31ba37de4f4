// RepMixer depthwise pair with the 7x7 as banded-Toeplitz products on the 5th-gen tensor cores (tcgen05 / TMEM),
// mci.py:806-853 (RepMixer) + :921 (ConvFFN.conv):
//     y = dw3x3(x) + b3        (identity + BN branches folded by the packer)     -> global (block residual)
//     z = dw7x7(y) + b7        (BN folded)                                        -> global (fc1's operand)
//
// Along one image row a depthwise conv IS a matrix product with a banded Toeplitz matrix.  Per channel c and tap row dy:
//     Z_c[h][w] += sum_k  Y_c[h + dy][k] * T_{c,dy}[k][w],       T_{c,dy}[k][w] = w7[c][dy][k - w]   (0 <= k - w <= 6, else 0)
// with Y_c the channel's PLANE (rows h, columns k contiguous).  On tcgen05 that is, per channel, 3 K-steps x 7 tap rows of
//     D[64 rows x N cols] (+)= A[64 rows x 16 k] * B[16 k x N]
//   A = the plane in the canonical no-swizzle K-major layout [8-column chunk][row][8 cols]: every row is 16 B apart, so the
//       tap-row shift dy is a START-ADDRESS offset of dy * 16 B -- no im2col, no copies;
//   B = the Toeplitz slice.  The band is shift-invariant, so ONE 640-byte tile per (channel, dy) serves every K-step: with the
//       output window starting 8 columns left of the K-step, B[n][k] = w7[dy][k - n + 8]; of its 3 x 2 core matrices only two
//       distinct ones (P, Q) are non-zero and LBO = 2 blocks lets the zero block be shared: memory = [P Q 0 P Q];
//   D = fp32 accumulator in TMEM.  M = 64 keeps the plane small enough to double-buffer; two channels share the 32-lane
//       sub-partitions (M = 64 occupies lanes 0-15 of each; the partner channel's accumulator sits at lane offset 16), so every
//       epilogue thread drains a useful row.
// An N <= 24, K = 16 MMA is bound by its shared-memory A read (~16 clk for 64 rows); 168 of them per 64 x 32 x 8 item.  The
// tensor core replaces 49 of the 58 multiply-adds per output (the 3x3 stays on the FMA pipes, as packed FFMA2).
//
// One CTA per SM, persistent over the (image, tile) items of ONE 8-channel group; warp-specialised, everything double-buffered:
//   warp 0      TMA: x tile (72 x 41 px x 8 ch, NHWC, zero OOB fill = the 3x3's zero padding)
//   warps 6-16  phase 1 (FMA pipes): y = dw3x3(x) + b3 on 70 x 38; bf16 y -> global (centre) and -> the 8 planes
//               (zero outside the image = the 7x7's zero padding)
//   warp 1      MMA issue: 8 channels x 3 K-steps x 7 tap rows, accumulate-only (the epilogue re-zeroes what it drains)
//   warps 2-5   epilogue: TMEM -> +b7 -> bf16 -> staging tile [col][row][8 ch] (reuses the plane buffer the MMAs just released;
//               channels c and c + 4 share a sub-partition, so a thread packs 4 channels = one 8-byte store, and the column-major
//               tile makes those stores conflict-free) -> one TMA store per tile column (64 rows x 16 B, no LSU traffic)
//
// Every global access of this kernel is a 16-byte piece (8 channels) of a 32-byte sector whose other half belongs to the neighbouring
// channel group, i.e. to ANOTHER CTA.  Measured: whenever the two drift apart (staggered CTA starts under programmatic dependent launch
// were enough), half-written / half-read sectors leave the L2 and the kernel -- and its successor -- slow down by up to 1.6x.  The two
// sibling groups are therefore launched as a 2-CTA cluster and their TMA warps handshake once per item (two alternating mbarriers,
// remote arrives; a final cluster barrier keeps a CTA alive while its sibling may still arrive on it): stage 0 at batch 32 went from
// 770 to 493 us per launch.  Clusters of 4 / 8 lose more resident CTAs to GPC placement than they gain.
#pragma once
#include "gemm_tcgen05.cuh"

namespace fvhd {

struct MixZ {
    static constexpr int CG = 8;                        // channels per CTA: one 16-byte NHWC vector
    static constexpr int ZR = 64, ZC = 32;              // z tile
    static constexpr int YR = ZR + 6;                   // 70 plane rows
    static constexpr int YC = ZC + 6;                   // 38 used plane columns (K is padded to 48 with a shared zero chunk)
    static constexpr int XR = ZR + 8;                   // 72 x rows
    static constexpr int XP = 41;                       // x pixel pitch = TMA box width (>= ZC + 8; odd -> conflict-free phase-1 reads)
    static constexpr int X_BYTES = XR * XP * 16;        // 47232
    static constexpr int NCH = 5;                       // 8-column chunks stored per plane (columns 0..39)
    static constexpr int CH_BYTES = YR * 16;            // 1120: one chunk = 70 rows x 16 B; = LBO between the K core matrices
    static constexpr int PL_STRIDE = NCH * CH_BYTES + 16;   // 5616: +16 spreads the planes over the banks for the 16-B plane stores
    static constexpr int PL_BYTES = CG * PL_STRIDE;     // 44928 per buffer
    static constexpr int ZERO_BYTES = 1152;             // the shared all-zero K chunk (>= 70 rows x 16 B)
    static constexpr int BT_BYTES = 640;                // one Toeplitz tile: [P Q 0 P Q] core matrices
    static constexpr int B_BYTES = CG * 7 * BT_BYTES;   // 35840
    static constexpr int ST_COL = ZR * 16;              // 1024: one z staging column = 64 rows x 16 B (aliases a plane buffer)
    static constexpr int NPROD = 11;                    // phase-1 warps: 1400 units / 352 threads = 3.98
    static constexpr int THREADS = (6 + NPROD) * 32;    // 544
    static constexpr int TMEM_COLS = 256;               // 2 buffers x 4 channel pairs x 32 columns
    static constexpr int OFF_X = 0;
    static constexpr int OFF_PL = 2 * X_BYTES;                   // 94464
    static constexpr int OFF_ZERO = OFF_PL + 2 * PL_BYTES;       // 184320
    static constexpr int OFF_B = OFF_ZERO + ZERO_BYTES;          // 185472
    static constexpr int OFF_CONST = OFF_B + B_BYTES;            // 221312: w3 [9][8], b3 [8], b7 [8] floats
    static constexpr int OFF_BAR = OFF_CONST + 512;
    static constexpr size_t SMEM = (size_t)OFF_BAR + 256 + 1024 /*align*/;
    static_assert(ZC * ST_COL <= PL_BYTES, "z staging must fit in one plane buffer");
    static_assert(X_BYTES % 128 == 0 && PL_BYTES % 128 == 0 && OFF_B % 16 == 0, "alignment");
    static_assert(SMEM <= 232448, "shared memory budget");
};

// No-swizzle K-major smem descriptor: 8-row x 16-B core matrices; LBO = byte distance between the two K core matrices of one
// MMA (K = 16 bf16 = 2 x 16 B), SBO = byte distance between consecutive 8-row groups.
__device__ __forceinline__ uint64_t umma_desc_k_nosw(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// registers -> TMEM: zero this warp's 32 lanes x 32 consecutive columns
__device__ __forceinline__ void tmem_zero32(uint32_t taddr) {
    asm volatile(
        "{\n\t.reg .b32 z;\n\tmov.b32 z, 0;\n\t"
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {z, z, z, z, z, z, z, z, z, z, z, z, z, z, z, z, z, z, z, z, z, z, z, z, z, z, z, z, z, z, z, z};\n\t}"
        ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// TMEM -> registers: this warp's 32 lanes x 8 consecutive fp32 columns
__device__ __forceinline__ void tmem_ldx8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
// 4-D tiled store shared -> global (bulk async group), coords {channel, x, y, image}; clips what lies outside the tensor
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

// d += a * b on both halves (Blackwell packed fp32 FMA: SASS FFMA2)
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
    uint64_t dd, aa, bb;
    asm("mov.b64 %0, {%1, %2};" : "=l"(dd) : "f"(d.x), "f"(d.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b.x), "f"(b.y));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(dd));
}

struct MixZParams {
    bf16* y;                 // [B, H, W, C]
    bf16* z;
    const float* w3;         // [9][C]
    const float* b3;         // [C]
    const float* w7;         // [49][C]
    const float* b7;
    int B, H, W, C;
    int tiles_x, tiles_y;    // spatial tiles per image
    int groups;              // C / 8
    int ctas_per_group;      // gridDim.x / groups
    int pair_sync;           // cluster size CS (0 / 1: none): launched as CS-CTA clusters of sibling channel groups (CS = 2: the two halves
                             //    of every 32-byte sector of x / y / z; 4: a 64-byte pair of sectors); their TMA warps handshake once per
                             //    item so all pieces of a sector are read and written within one item time
    int pdl_trigger;         // 1: griddepcontrol.launch_dependents at kernel start (the successor's CTAs may become resident early)
    int dbg;                 // FVHD_TZ_SKIP bits (timing experiments only): 1 no MMA issue, 2 no phase-1 math, 4 no epilogue work, 8 no TMA load
};

__global__ void __launch_bounds__(MixZ::THREADS, 1)
repmixer_tz_kernel(const __grid_constant__ CUtensorMap tmX /*x: NHWC, box {8, XP, XR, 1}*/,
                   const __grid_constant__ CUtensorMap tmZ /*z: NHWC, box {8, 1, ZR, 1}*/, const MixZParams p) {
    using U = MixZ;
    extern __shared__ uint8_t mixz_smem_raw[];
    const uint32_t raw_addr = smem_u32(mixz_smem_raw);
    uint8_t* smem = mixz_smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* sx = smem + U::OFF_X;
    uint8_t* spl = smem + U::OFF_PL;
    uint8_t* szero = smem + U::OFF_ZERO;
    uint8_t* sb = smem + U::OFF_B;
    float* w3s = reinterpret_cast<float*>(smem + U::OFF_CONST);      // [9][8]
    float* b3s = w3s + 72;
    float* b7s = b3s + 8;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + U::OFF_BAR);
    uint64_t* x_full = bars;            // [2] TMA -> phase 1
    uint64_t* x_free = bars + 2;        // [2] phase-1 warps (8) -> TMA
    uint64_t* pl_full = bars + 4;       // [2] phase-1 warps (8) -> MMA
    uint64_t* pl_free = bars + 6;       // [2] epilogue warps (4): staging read out -> phase 1 may rewrite the planes
    uint64_t* acc_full = bars + 8;      // [2] MMA commit -> epilogue
    uint64_t* acc_free = bars + 10;     // [2] epilogue warps (4): accumulator drained and re-zeroed -> MMA
    uint64_t* pair_bar = bars + 12;     // [2] sibling-CTA handshake (pair_sync), alternating so a sibling one item ahead cannot wrap a phase
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    if (p.pdl_trigger) pdl_launch_dependents();
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);     // broadcast: role branches are provably warp-uniform
    const int lane = threadIdx.x & 31;
    const int grp = (int)blockIdx.x % p.groups;
    const int cta_in_grp = (int)blockIdx.x / p.groups;
    const int c0 = grp * U::CG;
    const int tiles_img = p.tiles_x * p.tiles_y;
    const int n_items = p.B * tiles_img;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmZ);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&x_full[i], 1); mbar_init(&x_free[i], U::NPROD);
            mbar_init(&pl_full[i], U::NPROD); mbar_init(&pl_free[i], 4);
            mbar_init(&acc_full[i], 1); mbar_init(&acc_free[i], 4);
        }
        { const uint32_t others = p.pair_sync > 1 ? (uint32_t)p.pair_sync - 1u : 1u; mbar_init(&pair_bar[0], others); mbar_init(&pair_bar[1], others); }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, U::TMEM_COLS);
        tmem_relinquish();
    }
    // ---- constants (never written by a kernel of the forward): Toeplitz tiles, 3x3 taps, biases -- before the PDL wait
    {
        uint4* z4 = reinterpret_cast<uint4*>(szero);                       // zero chunk + all B tiles are contiguous
        for (int i = threadIdx.x; i < (U::ZERO_BYTES + U::B_BYTES) / 16; i += U::THREADS) z4[i] = make_uint4(0, 0, 0, 0);
        if (threadIdx.x < 72) w3s[threadIdx.x] = __ldg(p.w3 + (size_t)(threadIdx.x >> 3) * p.C + c0 + (threadIdx.x & 7));
        if (threadIdx.x < 8) {
            b3s[threadIdx.x] = __ldg(p.b3 + c0 + threadIdx.x);
            b7s[threadIdx.x] = __ldg(p.b7 + c0 + threadIdx.x);
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < U::CG * 7 * 64; idx += U::THREADS) {
        const int b = idx & 7, a = (idx >> 3) & 7, t = idx >> 6;          // t = c * 7 + dy
        const int c = t / 7, dy = t - c * 7;
        const float* wrow = p.w7 + (size_t)(dy * 7) * p.C + c0 + c;
        uint8_t* tile = sb + t * U::BT_BYTES + a * 16 + b * 2;
        const int dp = 8 - (a - b);        // P[a][b] = w[dy][8 - (a - b)],  a - b in [2, 7]
        const int dq = b - a;              // Q[a][b] = w[dy][b - a],        b - a in [0, 6]
        if (dp >= 1 && dp <= 6) {
            const bf16 v = __float2bfloat16_rn(__ldg(wrow + (size_t)dp * p.C));
            *reinterpret_cast<bf16*>(tile) = v;
            *reinterpret_cast<bf16*>(tile + 3 * 128) = v;
        }
        if (dq >= 0 && dq <= 6) {
            const bf16 v = __float2bfloat16_rn(__ldg(wrow + (size_t)dq * p.C));
            *reinterpret_cast<bf16*>(tile + 128) = v;
            *reinterpret_cast<bf16*>(tile + 4 * 128) = v;
        }
    }
    fence_proxy_async_smem();                    // generic-proxy writes of B / zero chunk -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (p.pair_sync > 1) cluster_sync_all();     // the siblings' barriers are initialised before anyone arrives on them

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer
            pdl_wait();                          // x is the predecessor's output
            int i = 0;
            for (int it = cta_in_grp; it < n_items; it += p.ctas_per_group, ++i) {
                const int b = it / tiles_img, t = it - b * tiles_img;
                const int ty0 = (t / p.tiles_x) * U::ZR, tx0 = (t % p.tiles_x) * U::ZC;
                const int xb = i & 1;
                mbar_wait(&x_free[xb], (((uint32_t)i >> 1) & 1u) ^ 1u);
                if (p.pair_sync > 1) {           // every sibling is ready to fetch item i
                    const uint32_t me = cluster_ctarank();
                    for (uint32_t r = 0; r < (uint32_t)p.pair_sync; ++r)
                        if (r != me) mbar_arrive_cluster(cluster_map(smem_u32(&pair_bar[xb]), r));
                    mbar_wait_cluster(&pair_bar[xb], ((uint32_t)i >> 1) & 1u);
                }
                if (p.dbg & 8) { mbar_arrive(&x_full[xb]); continue; }
                mbar_expect_tx(&x_full[xb], U::X_BYTES);
                tma_load_4d(sx + xb * U::X_BYTES, &tmX, c0, tx0 - 4, ty0 - 4, b, &x_full[xb]);
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer: the whole warp runs the loop (converged, uniform operands); one elected lane issues
        const uint32_t idesc0 = umma_idesc_bf16(64, 16), idesc1 = umma_idesc_bf16(64, 24), idesc2 = umma_idesc_bf16(64, 8);
        // descriptor start addresses are 18-bit CTA-relative offsets: in a cluster launch the shared-window address carries the CTA's
        // rank in its high bits, which would spill into the LBO field
        const uint32_t pl_a = smem_u32(spl) & 0x3FFFFu, zero_a = smem_u32(szero) & 0x3FFFFu, sb_a = smem_u32(sb) & 0x3FFFFu;
        const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint64_t hi_sbo = ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);        // SBO = 128 B, version 1, no swizzle
        const uint64_t hi_b = hi_sbo | ((uint64_t)(256 >> 4) << 16);                       // B: LBO = 2 blocks
        const uint64_t hi_a = hi_sbo | ((uint64_t)(U::CH_BYTES >> 4) << 16);               // A: LBO = chunk stride
        int i = 0;
        for (int it = cta_in_grp; it < n_items; it += p.ctas_per_group, ++i) {
            const int buf = i & 1;
            const uint32_t ph = ((uint32_t)i >> 1) & 1u;
            mbar_wait(&pl_full[buf], ph);
            mbar_wait(&acc_free[buf], ph);       // completion 0 = the initial zero fill
            tc_fence_after();
#pragma unroll 1
            for (int c = (p.dbg & 1) ? U::CG : 0; c < U::CG; ++c) {
                const uint32_t dacc = tm + ((uint32_t)(16 * (c >> 2)) << 16) + (uint32_t)(buf * 128 + (c & 3) * 32);   // channels c, c + 4 share lanes
                const uint32_t a_base = pl_a + (uint32_t)(buf * U::PL_BYTES + c * U::PL_STRIDE);
                const uint32_t b_base = sb_a + (uint32_t)(c * 7 * U::BT_BYTES);
                // K-step 0: plane cols 0..15 -> outputs 0..15 (tile rows 8..23: start one block in);  K-step 1: cols 16..31 ->
                // outputs 8..31;  K-step 2: cols 32..47 (40..47 = the zero chunk) -> outputs 24..31 (tile rows 0..7)
                const uint64_t da0 = hi_a | (uint64_t)(a_base >> 4);
                const uint64_t da1 = hi_a | (uint64_t)((a_base + 2 * U::CH_BYTES) >> 4);
                const uint32_t a2 = a_base + 4 * U::CH_BYTES;
                const uint64_t da2 = hi_sbo | ((uint64_t)(((zero_a - a2) >> 4) & 0x3FFF) << 16) | (uint64_t)(a2 >> 4);
                const uint64_t db0 = hi_b | (uint64_t)((b_base + 128) >> 4);
                const uint64_t db1 = hi_b | (uint64_t)(b_base >> 4);
#pragma unroll
                for (int dy = 0; dy < 7; ++dy)
                    if (elect_one()) umma_bf16(dacc, da0 + (uint64_t)dy, db0 + (uint64_t)(dy * (U::BT_BYTES >> 4)), idesc0, 1u);
#pragma unroll
                for (int dy = 0; dy < 7; ++dy)
                    if (elect_one()) umma_bf16(dacc + 8, da1 + (uint64_t)dy, db1 + (uint64_t)(dy * (U::BT_BYTES >> 4)), idesc1, 1u);
#pragma unroll
                for (int dy = 0; dy < 7; ++dy)
                    if (elect_one()) umma_bf16(dacc + 24, da2 + (uint64_t)dy, db1 + (uint64_t)(dy * (U::BT_BYTES >> 4)), idesc2, 1u);
            }
            if (elect_one()) umma_commit(&acc_full[buf]);
            __syncwarp();
        }
    } else if (warp < 6) {
        // ---------------- epilogue warps: TMEM sub-partition q; lanes 0-15 = rows 16q.. of channels 0-3, lanes 16-31 = channels 4-7
        const int q = warp & 3;
        const int row = q * 16 + (lane & 15), mem = lane >> 4;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int k = 0; k < U::TMEM_COLS / 32; ++k) tmem_zero32(lane_base + (uint32_t)(k * 32));
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&acc_free[0]); mbar_arrive(&acc_free[1]); }
        const float4 bias = *reinterpret_cast<const float4*>(b7s + 4 * mem);
        pdl_wait();                              // global writes (z) come after the predecessor
        int i = 0;
        for (int it = cta_in_grp; it < n_items; it += p.ctas_per_group, ++i) {
            const int buf = i & 1;
            const uint32_t ph = ((uint32_t)i >> 1) & 1u;
            const int b = it / tiles_img, t = it - b * tiles_img;
            const int ty0 = (t / p.tiles_x) * U::ZR, tx0 = (t % p.tiles_x) * U::ZC;
            uint8_t* stg = spl + buf * U::PL_BYTES;                   // the planes of this item: free once acc_full has fired
            mbar_wait(&acc_full[buf], ph);
            tc_fence_after();
            if (!(p.dbg & 4)) {
#pragma unroll 1
                for (int g = 0; g < 4; ++g) {    // 8 tile columns at a time: 4 channels x 8 columns per thread
                    uint32_t r[4][8];
#pragma unroll
                    for (int pr = 0; pr < 4; ++pr) tmem_ldx8(lane_base + (uint32_t)(buf * 128 + pr * 32 + g * 8), r[pr]);
                    tmem_ld_wait();
                    uint8_t* dst = stg + (g * 8) * U::ST_COL + row * 16 + mem * 8;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        uint2 v;
                        v.x = pack_bf16x2(__uint_as_float(r[0][j]) + bias.x, __uint_as_float(r[1][j]) + bias.y);
                        v.y = pack_bf16x2(__uint_as_float(r[2][j]) + bias.z, __uint_as_float(r[3][j]) + bias.w);
                        *reinterpret_cast<uint2*>(dst + j * U::ST_COL) = v;
                    }
                }
#pragma unroll
                for (int pr = 0; pr < 4; ++pr) tmem_zero32(lane_base + (uint32_t)(buf * 128 + pr * 32));   // accumulate-only MMAs: hand the columns back zeroed
                tmem_st_wait();
            }
            fence_proxy_async_smem();            // staging is read by the TMA store
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_free[buf]);
            named_bar_sync(1, 128);              // staging tile complete (4 epilogue warps)
            if (lane == 0) {
                if (!(p.dbg & 4)) {
#pragma unroll 1
                    for (int k = 0; k < U::ZC / 4; ++k) {             // this warp's 8 tile columns
                        const int zc = q * (U::ZC / 4) + k;
                        if (tx0 + zc < p.W) tma_store_4d(&tmZ, stg + zc * U::ST_COL, c0, tx0 + zc, ty0, b);
                    }
                    tma_store_commit();
                    tma_store_wait_read0();      // staging has been read: the planes may be rewritten
                }
                mbar_arrive(&pl_free[buf]);
            }
            __syncwarp();
        }
        if (lane == 0) tma_store_wait_all();
    } else {
        // ---------------- phase 1: y = dw3x3(x) + b3; one unit = one plane row x one 8-column chunk x one channel pair
        const int pt = (int)threadIdx.x - 6 * 32;                        // 0..351
        const int cp = pt & 3;                                          // this thread's channel pair (NPROD * 32 is a multiple of 4)
        float2 wk[9];                                                    // its 3x3 taps and bias live in registers
#pragma unroll
        for (int k = 0; k < 9; ++k) wk[k] = *reinterpret_cast<const float2*>(w3s + k * 8 + 2 * cp);
        const float2 bb = *reinterpret_cast<const float2*>(b3s + 2 * cp);
        constexpr int UNITS = 4 * U::YR * U::NCH;                        // 1400
        pdl_wait();                              // global writes (y) come after the predecessor
        int i = 0;
        for (int it = cta_in_grp; it < n_items; it += p.ctas_per_group, ++i) {
            const int buf = i & 1;
            const uint32_t ph = ((uint32_t)i >> 1) & 1u;
            const int b = it / tiles_img, t = it - b * tiles_img;
            const int ty0 = (t / p.tiles_x) * U::ZR, tx0 = (t % p.tiles_x) * U::ZC;
            const uint32_t* xt = reinterpret_cast<const uint32_t*>(sx + buf * U::X_BYTES);
            uint8_t* pl = spl + buf * U::PL_BYTES;
            const size_t img_off = (size_t)b * p.H * p.W * p.C;
            mbar_wait(&x_full[buf], ph);
            mbar_wait(&pl_free[buf], ph ^ 1u);
#pragma unroll 1
            for (int u = (p.dbg & 2) ? UNITS : pt; u < UNITS; u += U::NPROD * 32) {
                const int rr = u >> 2;
                const int chunk = rr / U::YR, r = rr - chunk * U::YR;
                float2 acc[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = bb;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const uint32_t* rowp = xt + ((r + ky) * U::XP + chunk * 8) * 4 + cp;
                    float2 xin[10];
#pragma unroll
                    for (int j = 0; j < 10; ++j) xin[j] = unpack_bf16x2(rowp[j * 4]);
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) ffma2(acc[j], xin[j + kx], wk[ky * 3 + kx]);
                    }
                }
                const int gy = ty0 - 3 + r;
                const int gx0 = tx0 - 3 + chunk * 8;
                const bool row_in = gy >= 0 && gy < p.H;
                const bool row_c = r >= 3 && r < 3 + U::ZR;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int gx = gx0 + j;
                    const bool in = row_in && gx >= 0 && gx < p.W && (chunk * 8 + j) < U::YC;
                    acc[j].x = in ? acc[j].x : 0.f;              // zero outside the image: the 7x7's zero padding
                    acc[j].y = in ? acc[j].y : 0.f;
                    const int pc = chunk * 8 + j;
                    if (in && row_c && pc >= 3 && pc < 3 + U::ZC)
                        *reinterpret_cast<uint32_t*>(p.y + img_off + ((size_t)gy * p.W + gx) * p.C + c0 + 2 * cp) = pack_bf16x2(acc[j].x, acc[j].y);
                }
                uint4 v0, v1;
                v0.x = pack_bf16x2(acc[0].x, acc[1].x); v0.y = pack_bf16x2(acc[2].x, acc[3].x);
                v0.z = pack_bf16x2(acc[4].x, acc[5].x); v0.w = pack_bf16x2(acc[6].x, acc[7].x);
                v1.x = pack_bf16x2(acc[0].y, acc[1].y); v1.y = pack_bf16x2(acc[2].y, acc[3].y);
                v1.z = pack_bf16x2(acc[4].y, acc[5].y); v1.w = pack_bf16x2(acc[6].y, acc[7].y);
                uint8_t* d0 = pl + (2 * cp) * U::PL_STRIDE + chunk * U::CH_BYTES + r * 16;
                *reinterpret_cast<uint4*>(d0) = v0;
                *reinterpret_cast<uint4*>(d0 + U::PL_STRIDE) = v1;
            }
            fence_proxy_async_smem();            // the planes are read by the tensor core
            __syncwarp();
            if (lane == 0) { mbar_arrive(&pl_full[buf]); mbar_arrive(&x_free[buf]); }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, U::TMEM_COLS);
    }
    if (p.pair_sync > 1) cluster_sync_all();     // no CTA exits while a sibling may still arrive on its barrier
}

}  // namespace fvhd
