// Fused ConvFFN (mci.py:922-926), second generation, for every RepMixer stage (C in {96, 192, 384}) at any batch:
//     out = resid + fc2( GELU( fc1(z) + b1 ) ) + b2            (layer scale folded into fc2)
//
// Same dataflow as mlp_fused.cuh (z tile resident in smem, 64-wide hidden chunks: MMA1 -> TMEM acc1 -> GELU -> H chunk in
// swizzled smem -> MMA2 accumulates acc2 in TMEM; the [M, 4C] hidden never exists), rebuilt around what ncu showed for
// the first generation at batch 8 (profiles/r02_*): tensor pipe 23 % active, 54 k warp-instructions per 128-row tile,
// issue-bound at IPC 0.36 per scheduler with only two epilogue warps to pick from.  Changes:
//   * SIXTEEN epilogue warps (four per scheduler, one 16-column slice of the chunk each) instead of eight;
//   * GELU in packed half precision (f16x2: HMUL2 / HFMA2 / one tanh.approx.f16x2 per TWO elements) and H kept as f16 -- 11
//     significand bits, three more than the bf16 H of the first generation; MMA2 multiplies it by an f16 copy of W2 (packer:
//     `fc2.wh`; a mixed f16 x bf16 tcgen05.mma raised an illegal-instruction error on B200);
//   * the MMA warp runs warp-uniformly (role index broadcast by shfl, `elect_one` only on the tcgen05 instructions), so the
//     descriptors stay in uniform registers and the issue loop is back-to-back UTCHMMA;
//   * C = 384 in ONE CTA: acc2 = 384 TMEM columns as two N = 192 halves, half-chunk weight slots (as the cluster kernel),
//     three 24-KB slots -- this is the large-batch path of stage 2 (two plain GEMMs and a 12.6-MB hidden per image before);
//   * results leave through 16-B global stores straight from registers (no staging buffer: the smem goes to z / W).
//   * CS = 2 (optional, FVHD_CONVFFN_CS): launched as 2-CTA clusters that SHARE THE WEIGHT STREAM.  The two CTAs of a cluster work on
//     neighbouring tiles and walk the same ring of weight slots; every slot is fetched ONCE per cluster -- even slots by rank 0, odd
//     slots by rank 1 -- as a TMA multicast into both CTAs' rings (each CTA arms its own full barrier; a slot is reused when BOTH MMA
//     warps have released it: tcgen05.commit multicast onto both empty barriers).  Per-SM weight ingest halves; results are
//     bit-identical.  Measured neutral at batch 32 (the kernel is bound by its MMA1 -> GELU -> MMA2 chain there, not by ingest).
// Two restructurings were also measured at batch 32 and NOT kept (profiles/r02_findings.md): MMA1 running ahead of MMA2 across tile
// boundaries with the accumulator drain on 8 dedicated warps (C = 192: 415 vs 383 us per launch), and the 16 GELU warps split into
// two groups alternating over the chunks (412 us).
#pragma once
#include "mlp_fused.cuh"

namespace fvhd {

constexpr int CF_EPI_WARPS = 16;
constexpr int CF_THREADS = 32 * (2 + CF_EPI_WARPS);          // 576
constexpr int CF_SLOT = 24576;                               // one weight ring slot

template <int C> struct CfCfg {
    static constexpr int KB = (C + 63) / 64;                 // k-blocks of z / W1 (C = 96: the second one is half zero-filled)
    static constexpr int NC = C / 16;                        // hidden chunks (4C / 64)
    static constexpr int W1U = C == 384 ? 2 : 1;             // slots per W1 chunk
    static constexpr int W2U = C == 384 ? 2 : 1;             // slots per W2 chunk == N halves of MMA2
    static constexpr int N2 = C / W2U;                       // N of one MMA2
    static constexpr int KB1 = KB / W1U;                     // k-blocks of W1 per slot
    static constexpr int NSLOT = C == 384 ? 3 : 4;
    // acc1 / H buffers: MMA1 runs NB - 1 chunks ahead of MMA2, so the GELU epilogue's latency (a1_full -> h_full, ~1.5-2 k cycles
    // measured) is hidden behind NB - 1 chunks of tensor work instead of one.  C = 384 has neither the TMEM columns nor the smem.
    static constexpr int NB = C == 384 ? 2 : 4;
    static constexpr int ACC1_COL = C == 96 ? 128 : C;       // TMEM: acc2 at [0, C), acc1[b] at ACC1_COL + 64 b
    static constexpr int Z_BYTES = KB * GEMM_A_STAGE_BYTES;
    static constexpr size_t SMEM = (size_t)Z_BYTES + (size_t)NB * GEMM_A_STAGE_BYTES + (size_t)NSLOT * CF_SLOT + (size_t)5 * C * 4 + 256 + 1024;
    static_assert(KB1 * 64 * 128 <= CF_SLOT && N2 * 128 <= CF_SLOT, "weight slot size");
    static_assert(SMEM <= 227 * 1024, "ConvFFN smem");
    static_assert(ACC1_COL + NB * 64 <= 512, "TMEM columns");
};

__host__ __device__ __forceinline__ uint32_t umma_idesc_f16bf16(uint32_t a_bf16, uint32_t b_bf16, uint32_t M, uint32_t N) {
    return (1u << 4) | (a_bf16 << 7) | (b_bf16 << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}

// GELU of two pre-activations in packed half precision (same fit as gelu_erf in ptx.cuh: erf(x/sqrt2) = tanh(x P(x^2))).
// x2 saturates to +inf for |x| > 255 and is clamped to 50 by the min; for x << 0 tanh rounds to exactly -1 in f16, so
// hx * t + hx is exactly 0 (no select needed); for x >> 0 it is 2 hx = x.
__device__ __forceinline__ uint32_t gelu_f16x2(float a, float b) {
    uint32_t x, r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(x) : "f"(b), "f"(a));
    asm("{\n\t.reg .b32 x2, p, u, t, hx;\n\t"
        "mul.f16x2 x2, %1, %1;\n\t"
        "min.f16x2 x2, x2, %2;\n\t"
        "fma.rn.f16x2 p, x2, %3, %4;\n\t"
        "fma.rn.f16x2 p, p, x2, %5;\n\t"
        "mul.f16x2 u, %1, p;\n\t"
        "tanh.approx.f16x2 t, u;\n\t"
        "mul.f16x2 hx, %1, %6;\n\t"
        "fma.rn.f16x2 %0, hx, t, hx;\n\t}"
        : "=r"(r)
        : "r"(x), "r"(0x52405240u) /*50*/, "r"(0x8DE18DE1u) /*-3.5873e-4*/, "r"(0x28BE28BEu) /*3.70503e-2*/, "r"(0x3A613A61u) /*0.797458*/,
          "r"(0x38003800u) /*0.5*/);
    return r;
}

template <int C, int CS = 1>
__global__ void __launch_bounds__(CF_THREADS, 1)
convffn_tcgen05_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmW1,
                       const __grid_constant__ CUtensorMap tmW2, const MlpParams p) {
    static_assert(CS == 1 || CS == 2, "cluster size");
    using Cfg = CfCfg<C>;
    constexpr int KB = Cfg::KB, NC = Cfg::NC, NSLOT = Cfg::NSLOT, W1U = Cfg::W1U, W2U = Cfg::W2U, N2 = Cfg::N2, KB1 = Cfg::KB1;
    constexpr int NB = Cfg::NB, LA = NB - 1, ACC1 = Cfg::ACC1_COL;
    extern __shared__ uint8_t cf_smem_raw[];
    const uint32_t raw_addr = smem_u32(cf_smem_raw);
    uint8_t* smem = cf_smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* smemZ = smem;
    uint8_t* smemH = smemZ + Cfg::Z_BYTES;
    uint8_t* smemW = smemH + NB * GEMM_A_STAGE_BYTES;
    float* sb1 = reinterpret_cast<float*>(smemW + (size_t)NSLOT * CF_SLOT);
    float* sb2 = sb1 + 4 * C;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sb2 + C);
    uint64_t* z_full = bars;                 // [1]
    uint64_t* z_empty = bars + 1;            // [1]
    uint64_t* w_full = bars + 2;             // [NSLOT]
    uint64_t* w_empty = w_full + 4;          // [NSLOT]
    uint64_t* a1_full = w_empty + 4;         // [NB <= 4]
    uint64_t* a1_empty = a1_full + 4;        // [NB]
    uint64_t* h_full = a1_empty + 4;         // [NB]
    uint64_t* h_empty = h_full + 4;          // [NB]
    uint64_t* a2_full = h_empty + 4;         // [1]
    uint64_t* a2_empty = a2_full + 1;        // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a2_empty + 1);

    pdl_launch_dependents();
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);      // broadcast: role branches are warp-uniform for ptxas
    const int lane = threadIdx.x & 31;
    // tile walk: cluster `cl` of `ncl` takes tiles CS * (cl + k * ncl) + rank, k = 0, 1, ...; every CTA of a cluster runs the same
    // number of iterations (with CS = 2 and an odd tile count the last tile index of rank 1 is tiles_m: its z box is out of bounds
    // = zeros, its rows fail row_ok, nothing is stored), so the shared weight stream never loses a consumer
    const int rank = CS == 1 ? 0 : (int)(blockIdx.x % CS);
    const int cl = (int)blockIdx.x / CS, ncl = (int)gridDim.x / CS;
    const int tile0 = CS * cl + rank, tstep = CS * ncl;
    const int tile_end = p.tiles_m + rank;         // tile < tile_end  <=>  CS * (cl + k * ncl) < tiles_m

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmZ); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
        mbar_init(z_full, 1); mbar_init(z_empty, 1);
        for (int s = 0; s < NSLOT; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], CS); }
        for (int b = 0; b < NB; ++b) {
            mbar_init(&a1_full[b], 1); mbar_init(&a1_empty[b], CF_EPI_WARPS);
            mbar_init(&h_full[b], CF_EPI_WARPS); mbar_init(&h_empty[b], 1);
        }
        mbar_init(a2_full, 1); mbar_init(a2_empty, CF_EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < 4 * C; i += CF_THREADS) sb1[i] = __ldg(p.b1 + i);       // biases are constants
    for (int i = threadIdx.x; i < C; i += CF_THREADS) sb2[i] = __ldg(p.b2 + i);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (CS > 1) cluster_sync_all();              // the sibling's barriers exist before anything is multicast or committed onto them

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer: z tile, then the weight (half-)chunk boxes in the MMA warp's consumption order
            int wit = 0;
            auto load_w = [&](bool is_w2, int j, int u) {
                const int s = wit % NSLOT;
                const uint32_t ph = (uint32_t)(wit / NSLOT) & 1u;
                const bool mine = CS == 1 || (wit & (CS - 1)) == rank;           // CS = 2: this slot-load is fetched by one CTA for both
                ++wit;
                mbar_wait(&w_empty[s], ph ^ 1u);                                 // (CS = 2: released by BOTH MMA warps)
                uint8_t* dst = smemW + (size_t)s * CF_SLOT;
                if (!is_w2) {           // W1 rows [64 j, +64), k-blocks [KB1 u, +KB1): KB1 boxes of 64 x 64
                    mbar_expect_tx(&w_full[s], (uint32_t)KB1 * 64 * 128);
                    if (mine)
                        for (int kk = 0; kk < KB1; ++kk) {
                            if (CS == 1) tma_load_2d(dst + (size_t)kk * 64 * 128, &tmW1, (KB1 * u + kk) * 64, j * MLP_NH, &w_full[s]);
                            else tma_load_2d_mc(dst + (size_t)kk * 64 * 128, &tmW1, (KB1 * u + kk) * 64, j * MLP_NH, &w_full[s], (uint16_t)((1u << CS) - 1u));
                        }
                } else {                // W2 rows [N2 u, +N2), K columns [64 j, +64): one N2 x 64 box
                    mbar_expect_tx(&w_full[s], (uint32_t)N2 * 128);
                    if (mine) {
                        if (CS == 1) tma_load_2d(dst, &tmW2, j * MLP_NH, u * N2, &w_full[s]);
                        else tma_load_2d_mc(dst, &tmW2, j * MLP_NH, u * N2, &w_full[s], (uint16_t)((1u << CS) - 1u));
                    }
                }
            };
            // a tile's weight stream in the MMA warp's consumption order: step t issues W1[t] (t < NC) then W2[t - LA] (t >= LA);
            // slot-loads [from, to) of that sequence
            constexpr int TOT = NC * (W1U + W2U);
            auto run_tile_loads = [&](int from, int to) {
                int q = 0;
                for (int t = 0; t < NC + LA; ++t) {
                    if (t < NC)
                        for (int u = 0; u < W1U; ++u, ++q) if (q >= from && q < to) load_w(false, t, u);
                    if (t >= LA)
                        for (int u = 0; u < W2U; ++u, ++q) if (q >= from && q < to) load_w(true, t - LA, u);
                }
            };
            int q0 = 0;                                  // weights are constants: first ring fill before the PDL wait
            if (tile0 < tile_end) { q0 = NSLOT < TOT ? NSLOT : TOT; run_tile_loads(0, q0); }
            pdl_wait();
            int ti = 0;
            for (int tile = tile0; tile < tile_end; tile += tstep, ++ti) {
                mbar_wait(z_empty, ((uint32_t)ti & 1u) ^ 1u);
                mbar_expect_tx(z_full, (uint32_t)Cfg::Z_BYTES);
                for (int kb = 0; kb < KB; ++kb) tma_load_2d(smemZ + (size_t)kb * GEMM_A_STAGE_BYTES, &tmZ, kb * 64, tile * GEMM_BM, z_full);
                run_tile_loads(ti == 0 ? q0 : 0, TOT);
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (whole warp, uniform; one elected lane issues)
        const uint32_t idesc1 = umma_idesc_f16bf16(1, 1, GEMM_BM, MLP_NH);           // z (bf16) x W1 (bf16)
        const uint32_t idesc2 = umma_idesc_f16bf16(0, p.w2_f16 ? 0u : 1u, GEMM_BM, (uint32_t)N2);     // H (f16) x W2 (f16 copy; bf16 = mixed formats, test only)
        const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint64_t dz0 = umma_desc_sw128(smem_u32(smemZ));
        const uint64_t dh0 = umma_desc_sw128(smem_u32(smemH));
        const uint64_t dw0 = umma_desc_sw128(smem_u32(smemW));
        int wit = 0, ti = 0;
        auto take_slot = [&]() -> int {
            const int s = wit % NSLOT;
            const uint32_t ph = (uint32_t)(wit / NSLOT) & 1u;
            ++wit;
            mbar_wait(&w_full[s], ph);
            return s;
        };
        for (int tile = tile0; tile < tile_end; tile += tstep, ++ti) {
            mbar_wait(z_full, (uint32_t)ti & 1u);
#pragma unroll 1
            for (int t = 0; t < NC + LA; ++t) {
                if (t < NC) {        // ---- MMA1(t): acc1[b] = z . W1[t]^T
                    const int gj = ti * NC + t, b = gj % NB;
                    mbar_wait(&a1_empty[b], ((uint32_t)(gj / NB) & 1u) ^ 1u);    // epilogue has drained acc1[b]
                    const uint32_t acc1 = tm + ACC1 + (uint32_t)(b * MLP_NH);
#pragma unroll
                    for (int u = 0; u < W1U; ++u) {
                        const int s = take_slot();
                        tc_fence_after();
                        const uint64_t db = dw0 + (uint64_t)(s * (CF_SLOT >> 4));
#pragma unroll
                        for (int kk = 0; kk < KB1; ++kk) {
                            const int kb = KB1 * u + kk;
                            const int ksteps = (C - kb * 64) >= 64 ? 4 : (C - kb * 64) / 16;     // C = 96: the second k-block has 2 steps
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (k < ksteps && elect_one())
                                    umma_bf16(acc1, dz0 + (uint64_t)(kb * (GEMM_A_STAGE_BYTES >> 4) + 2 * k), db + (uint64_t)(kk * (64 * 128 >> 4) + 2 * k), idesc1,
                                              (u | kk | k) != 0 ? 1u : 0u);
                            }
                        }
                        if (elect_one()) { if (CS == 1) umma_commit(&w_empty[s]); else umma_commit_mc(&w_empty[s], (uint16_t)((1u << CS) - 1u)); }
                        __syncwarp();
                    }
                    if (elect_one()) {
                        umma_commit(&a1_full[b]);
                        if (t == NC - 1) umma_commit(z_empty);                   // z tile no longer needed
                    }
                    __syncwarp();
                }
                if (t >= LA) {       // ---- MMA2(j): acc2 (+)= H[b] . W2[:, chunk j]^T, LA chunks behind MMA1
                    const int j = t - LA;
                    const int gj = ti * NC + j, b = gj % NB;
                    mbar_wait(&h_full[b], (uint32_t)(gj / NB) & 1u);
                    if (j == 0) mbar_wait(a2_empty, ((uint32_t)ti & 1u) ^ 1u);   // previous tile's epilogue drained acc2
                    const uint64_t da = dh0 + (uint64_t)(b * (GEMM_A_STAGE_BYTES >> 4));
#pragma unroll
                    for (int u = 0; u < W2U; ++u) {
                        const int s = take_slot();
                        tc_fence_after();
                        const uint64_t db = dw0 + (uint64_t)(s * (CF_SLOT >> 4));
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (elect_one()) umma_bf16(tm + (uint32_t)(u * N2), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc2, (j | k) != 0 ? 1u : 0u);
                        if (elect_one()) { if (CS == 1) umma_commit(&w_empty[s]); else umma_commit_mc(&w_empty[s], (uint16_t)((1u << CS) - 1u)); }
                        __syncwarp();
                    }
                    if (elect_one()) umma_commit(&h_empty[b]);
                    __syncwarp();
                }
            }
            if (elect_one()) umma_commit(a2_full);
            __syncwarp();
        }
    } else {
        // ---------------- 16 epilogue warps: lane quarter q == 32 tile rows (hardware: warp % 4); k4 = 16-column slice of a chunk
        const int q = warp & 3;
        const int k4 = (warp - 2) >> 2;
        constexpr int CW = C / 4;                      // acc2 columns finished by this warp: [k4 * CW, +CW)
        const int row_in_tile = q * 32 + lane;
        const uint32_t sw = (uint32_t)(lane & 7);
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        pdl_wait();
        int ti = 0;
        for (int tile = tile0; tile < tile_end; tile += tstep, ++ti) {
            const int row = tile * GEMM_BM + row_in_tile;
            const bool row_ok = row < p.M;
            // ---- epilogue 1 per hidden chunk: columns [16 k4, +16) of the 64-wide chunk -> GELU -> f16 -> H[b]
#pragma unroll 1
            for (int j = 0; j < NC; ++j) {
                const int gj = ti * NC + j, b = gj % NB;
                const uint32_t use = (uint32_t)(gj / NB);
                const float4* bb = reinterpret_cast<const float4*>(sb1 + j * MLP_NH + k4 * 16);
                const float4 bv0 = bb[0], bv1 = bb[1], bv2 = bb[2], bv3 = bb[3];
                mbar_wait(&a1_full[b], use & 1u);
                tc_fence_after();
                uint32_t r[16];
                tmem_ld16(lane_base + ACC1 + (uint32_t)(b * MLP_NH + k4 * 16), r);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a1_empty[b]);                        // acc1[b] slice is in registers now
                uint4 o0, o1;
                o0.x = gelu_f16x2(__uint_as_float(r[0]) + bv0.x, __uint_as_float(r[1]) + bv0.y);
                o0.y = gelu_f16x2(__uint_as_float(r[2]) + bv0.z, __uint_as_float(r[3]) + bv0.w);
                o0.z = gelu_f16x2(__uint_as_float(r[4]) + bv1.x, __uint_as_float(r[5]) + bv1.y);
                o0.w = gelu_f16x2(__uint_as_float(r[6]) + bv1.z, __uint_as_float(r[7]) + bv1.w);
                o1.x = gelu_f16x2(__uint_as_float(r[8]) + bv2.x, __uint_as_float(r[9]) + bv2.y);
                o1.y = gelu_f16x2(__uint_as_float(r[10]) + bv2.z, __uint_as_float(r[11]) + bv2.w);
                o1.z = gelu_f16x2(__uint_as_float(r[12]) + bv3.x, __uint_as_float(r[13]) + bv3.y);
                o1.w = gelu_f16x2(__uint_as_float(r[14]) + bv3.z, __uint_as_float(r[15]) + bv3.w);
                mbar_wait(&h_empty[b], (use & 1u) ^ 1u);                         // MMA2 of the previous use has read H[b]
                uint8_t* hrow = smemH + (size_t)b * GEMM_A_STAGE_BYTES + (size_t)row_in_tile * 128;
                *reinterpret_cast<uint4*>(hrow + ((((uint32_t)(2 * k4)) ^ sw) << 4)) = o0;
                *reinterpret_cast<uint4*>(hrow + ((((uint32_t)(2 * k4 + 1)) ^ sw) << 4)) = o1;
                fence_proxy_async_smem();                                        // visible to the tensor core's smem reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&h_full[b]);
            }
            // ---- epilogue 2: acc2 columns [k4 CW, +CW) -> +b2 -> +resid -> bf16 -> global (16-B stores from registers)
            const bf16* rrow = p.resid + (size_t)row * C + k4 * CW;
            bf16* drow = p.D + (size_t)row * C + k4 * CW;
            constexpr int NU = CW / 8;                                           // 3 / 6 / 12 eight-column units
            constexpr int PRE = NU < 6 ? NU : 6;
            uint4 rpre[PRE];
#pragma unroll
            for (int i = 0; i < PRE; ++i) rpre[i] = row_ok ? *reinterpret_cast<const uint4*>(rrow + i * 8) : make_uint4(0, 0, 0, 0);
            mbar_wait(a2_full, (uint32_t)ti & 1u);
            tc_fence_after();
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                uint32_t r[8];
                tmem_ld8(lane_base + (uint32_t)(k4 * CW + i * 8), r);
                uint4 rv;
                if (i < PRE) rv = rpre[i];
                else rv = row_ok ? *reinterpret_cast<const uint4*>(rrow + i * 8) : make_uint4(0, 0, 0, 0);
                const float4 b0 = *reinterpret_cast<const float4*>(sb2 + k4 * CW + i * 8);
                const float4 b1v = *reinterpret_cast<const float4*>(sb2 + k4 * CW + i * 8 + 4);
                tmem_ld_wait();
                const float2 r0 = unpack_bf16x2(rv.x), r1 = unpack_bf16x2(rv.y), r2 = unpack_bf16x2(rv.z), r3 = unpack_bf16x2(rv.w);
                uint4 o;
                o.x = pack_bf16x2(__uint_as_float(r[0]) + b0.x + r0.x, __uint_as_float(r[1]) + b0.y + r0.y);
                o.y = pack_bf16x2(__uint_as_float(r[2]) + b0.z + r1.x, __uint_as_float(r[3]) + b0.w + r1.y);
                o.z = pack_bf16x2(__uint_as_float(r[4]) + b1v.x + r2.x, __uint_as_float(r[5]) + b1v.y + r2.y);
                o.w = pack_bf16x2(__uint_as_float(r[6]) + b1v.z + r3.x, __uint_as_float(r[7]) + b1v.w + r3.y);
                if (row_ok) *reinterpret_cast<uint4*>(drow + i * 8) = o;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a2_empty);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
    if (CS > 1) cluster_sync_all();              // no CTA exits while its sibling may still multicast into its ring / commit onto its barriers
}

}  // namespace fvhd
