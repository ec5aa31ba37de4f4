// Fused ConvFFN (mci.py:922-926) for C in {96, 192}:   out = resid + fc2( GELU( fc1(z) + b1 ) ) + b2
// (layer scale folded into fc2).  One CTA owns a 128-pixel tile; the 4C-wide hidden never leaves the SM:
//
//   z tile [128 x C]  --TMA-->  smem (resident for the tile)
//   for each 64-wide hidden chunk j (4C/64 chunks):
//       MMA1   acc1[j&1] (TMEM, 64 cols)  = z . W1[j]^T                      (tcgen05, K = C)
//       epi1   acc1 -> +b1 -> GELU -> bf16 -> H[j&1]  (smem, 128 x 64, 128-B swizzle = K-major A operand)
//       MMA2   acc2 (TMEM, C cols)       += H[j&1] . W2[:, j]^T              (tcgen05, K = 64)
//   epi2   acc2 -> +b2 -> +resid -> bf16 -> swizzled staging -> TMA store
//
// MMA1(j+1) and MMA2(j-1) are issued while the epilogue warps run GELU on chunk j (two acc1 / two H buffers), and the
// W1 / W2 chunk boxes stream through one mbarrier ring in exactly the order the MMA thread consumes them.
// The fc2 accumulation order equals the unfused GEMM's (k-block j == hidden chunk j), so results are bit-identical to
// gemm(fc1)+gemm(fc2) while the [M, 4C] hidden (50 / 25 MB per image and block in stages 0 / 1) is never written.
#pragma once
#include "gemm_tcgen05.cuh"

namespace fvhd {

constexpr int MLP_NH = 64;                       // hidden chunk == one 128-B swizzle k-block of fc2
constexpr int MLP_SLOTS = 4;
constexpr int MLP_THREADS = GEMM_THREADS;        // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int MLP_ACC1_COL = 256;                // TMEM: acc2 at columns [0, C), acc1[b] at 256 + 64 b

struct MlpParams {
    int M, C;
    int tiles_m;
    const float* b1;           // [4C]
    const float* b2;           // [C]
    const bf16* resid;         // [M, C]
    bf16* D;                   // [M, C]
    unsigned long long* trace; // cluster kernel, debug: 64 globaltimer stamps per CTA (nullptr = off)
    int w2_f16;                // convffn.cuh: fc2 weights are f16 (same format as the f16 hidden); 0 = bf16 (mixed-format MMA, test only)
};

__host__ __device__ inline int mlp_kb(int C) { return (C + 63) / 64; }
__host__ __device__ inline int mlp_slot_bytes(int C) { return mlp_kb(C) * 64 * 128; }     // >= C * 128 (W2 chunk)
__host__ inline size_t mlp_smem_bytes(int C) {
    return (size_t)mlp_kb(C) * GEMM_A_STAGE_BYTES + 2 * GEMM_A_STAGE_BYTES + (size_t)MLP_SLOTS * mlp_slot_bytes(C) +
           GEMM_EPI_WARPS * GEMM_EPI_STAGE_BYTES + (size_t)5 * C * 4 + 1024 + 256;
}

__global__ void __launch_bounds__(MLP_THREADS, 1)
mlp_fused_tcgen05_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmW1,
                         const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmD, const MlpParams p) {
    extern __shared__ uint8_t mlp_smem_raw[];
    const uint32_t raw_addr = smem_u32(mlp_smem_raw);
    uint8_t* smem = mlp_smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

    const int C = p.C;
    const int KB = mlp_kb(C);                    // k-blocks of z / W1 (the last one may be half zero-filled: C = 96)
    const int NC = C / 16;                       // hidden chunks: 4C / 64
    const int ks1 = C / 16;                      // K = 16 steps of MMA1
    const uint32_t slot_bytes = (uint32_t)mlp_slot_bytes(C);
    uint8_t* smemZ = smem;
    uint8_t* smemH = smemZ + (size_t)KB * GEMM_A_STAGE_BYTES;
    uint8_t* smemW = smemH + 2 * GEMM_A_STAGE_BYTES;
    uint8_t* smemE = smemW + (size_t)MLP_SLOTS * slot_bytes;
    float* sb1 = reinterpret_cast<float*>(smemE + GEMM_EPI_WARPS * GEMM_EPI_STAGE_BYTES);
    float* sb2 = sb1 + 4 * C;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sb2 + C);
    uint64_t* z_full = bars;                 // [1]
    uint64_t* z_empty = bars + 1;            // [1]
    uint64_t* w_full = bars + 2;             // [MLP_SLOTS]
    uint64_t* w_empty = w_full + MLP_SLOTS;  // [MLP_SLOTS]
    uint64_t* a1_full = w_empty + MLP_SLOTS; // [2]
    uint64_t* a1_empty = a1_full + 2;        // [2]
    uint64_t* h_full = a1_empty + 2;         // [2]
    uint64_t* h_empty = h_full + 2;          // [2]
    uint64_t* a2_full = h_empty + 2;         // [1]
    uint64_t* a2_empty = a2_full + 1;        // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a2_empty + 1);

    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmZ); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2); tma_prefetch_desc(&tmD);
        mbar_init(z_full, 1); mbar_init(z_empty, 1);
        for (int s = 0; s < MLP_SLOTS; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&a1_full[b], 1); mbar_init(&a1_empty[b], GEMM_EPI_WARPS);
            mbar_init(&h_full[b], GEMM_EPI_WARPS); mbar_init(&h_empty[b], 1);
        }
        mbar_init(a2_full, 1); mbar_init(a2_empty, GEMM_EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < 4 * C; i += MLP_THREADS) sb1[i] = __ldg(p.b1 + i);       // biases are constants
    for (int i = threadIdx.x; i < C; i += MLP_THREADS) sb2[i] = __ldg(p.b2 + i);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer: z tile, then the W1/W2 chunk boxes in the MMA thread's consumption order
            int wit = 0;
            auto load_w = [&](bool is_w2, int j) {
                const int s = wit % MLP_SLOTS;
                const uint32_t ph = (uint32_t)(wit / MLP_SLOTS) & 1u;
                ++wit;
                mbar_wait(&w_empty[s], ph ^ 1u);
                uint8_t* dst = smemW + (size_t)s * slot_bytes;
                if (!is_w2) {           // W1 rows [64 j, 64 j + 64), all of K = C: KB boxes of 64 x 64
                    mbar_expect_tx(&w_full[s], (uint32_t)KB * 64 * 128);
                    for (int kb = 0; kb < KB; ++kb) tma_load_2d(dst + (size_t)kb * 64 * 128, &tmW1, kb * 64, j * MLP_NH, &w_full[s]);
                } else {                // W2 all C rows, K columns [64 j, 64 j + 64): one box of C x 64
                    mbar_expect_tx(&w_full[s], (uint32_t)C * 128);
                    tma_load_2d(dst, &tmW2, j * MLP_NH, 0, &w_full[s]);
                }
            };
            // the weight stream of one tile, in the MMA thread's consumption order: W1[0], (W1[j], W2[j-1]) j = 1..NC-1, W2[NC-1]
            auto load_seq = [&](int q) {
                if (q == 0) load_w(false, 0);
                else if (q == 2 * NC - 1) load_w(true, NC - 1);
                else if (q & 1) load_w(false, (q + 1) >> 1);
                else load_w(true, (q >> 1) - 1);
            };
            // weights are constants of the forward: the first ring fill is issued BEFORE the PDL wait and overlaps the
            // predecessor's tail; only z (the predecessor's output) has to wait
            int q0 = 0;
            if ((int)blockIdx.x < p.tiles_m)
                for (; q0 < MLP_SLOTS && q0 < 2 * NC; ++q0) load_seq(q0);
            pdl_wait();
            int ti = 0;
            for (int tile = blockIdx.x; tile < p.tiles_m; tile += gridDim.x, ++ti) {
                mbar_wait(z_empty, ((uint32_t)ti & 1u) ^ 1u);
                mbar_expect_tx(z_full, (uint32_t)KB * GEMM_A_STAGE_BYTES);
                for (int kb = 0; kb < KB; ++kb) tma_load_2d(smemZ + (size_t)kb * GEMM_A_STAGE_BYTES, &tmZ, kb * 64, tile * GEMM_BM, z_full);
                for (int q = ti == 0 ? q0 : 0; q < 2 * NC; ++q) load_seq(q);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------- MMA issuer
            const uint32_t idesc1 = umma_idesc_bf16(GEMM_BM, MLP_NH);
            const uint32_t idesc2 = umma_idesc_bf16(GEMM_BM, (uint32_t)C);
            const int half_nc = NC / 2;          // uses of each acc1 / H buffer per tile (NC is even for C = 96, 192, 384)
            int wit = 0, ti = 0;
            auto mma2 = [&](int j, int ti_) {    // acc2 (+)= H[j&1] . W2[:, j]^T
                const int b = j & 1;
                const uint32_t use = (uint32_t)(ti_ * half_nc + (j >> 1));
                mbar_wait(&h_full[b], use & 1u);
                const int s = wit % MLP_SLOTS;
                const uint32_t ph = (uint32_t)(wit / MLP_SLOTS) & 1u;
                ++wit;
                mbar_wait(&w_full[s], ph);
                if (j == 0) mbar_wait(a2_empty, ((uint32_t)ti_ & 1u) ^ 1u);     // previous tile's epilogue drained acc2
                tc_fence_after();
                const uint64_t da = umma_desc_sw128(smem_u32(smemH + (size_t)b * GEMM_A_STAGE_BYTES));
                const uint64_t db = umma_desc_sw128(smem_u32(smemW + (size_t)s * slot_bytes));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc2, (j | k) != 0 ? 1u : 0u);
                umma_commit(&w_empty[s]);
                umma_commit(&h_empty[b]);
            };
            for (int tile = blockIdx.x; tile < p.tiles_m; tile += gridDim.x, ++ti) {
                mbar_wait(z_full, (uint32_t)ti & 1u);
                for (int j = 0; j < NC; ++j) {
                    const int b = j & 1;
                    const uint32_t use = (uint32_t)(ti * half_nc + (j >> 1));
                    mbar_wait(&a1_empty[b], (use & 1u) ^ 1u);                    // epilogue has drained acc1[b]
                    const int s = wit % MLP_SLOTS;
                    const uint32_t ph = (uint32_t)(wit / MLP_SLOTS) & 1u;
                    ++wit;
                    mbar_wait(&w_full[s], ph);
                    tc_fence_after();
                    const uint32_t acc1 = tmem_base + MLP_ACC1_COL + (uint32_t)b * MLP_NH;
                    for (int ks = 0; ks < ks1; ++ks) {
                        const int kb = ks >> 2, k = ks & 3;
                        const uint64_t da = umma_desc_sw128(smem_u32(smemZ + (size_t)kb * GEMM_A_STAGE_BYTES)) + (uint64_t)(2 * k);
                        const uint64_t db = umma_desc_sw128(smem_u32(smemW + (size_t)s * slot_bytes + (size_t)kb * 64 * 128)) + (uint64_t)(2 * k);
                        umma_bf16(acc1, da, db, idesc1, ks != 0 ? 1u : 0u);
                    }
                    umma_commit(&w_empty[s]);
                    umma_commit(&a1_full[b]);
                    if (j == NC - 1) umma_commit(z_empty);                       // z tile no longer needed
                    if (j >= 1) mma2(j - 1, ti);
                }
                mma2(NC - 1, ti);
                umma_commit(a2_full);
            }
        }
    } else {
        // ---------------- epilogue warps: lane quarter q == 32 tile rows; warps 2-5 / 6-9 split the columns
        const int q = warp & 3;
        const int hh = (warp - 2) >> 2;
        const int half_nc = NC / 2;
        const int row_in_tile = q * 32 + lane;
        uint8_t* stage = smemE + (size_t)(warp - 2) * GEMM_EPI_STAGE_BYTES;
        const uint32_t sw = (uint32_t)(lane & 7);
        bool store_pending = false;
        pdl_wait();
        int ti = 0;
        for (int tile = blockIdx.x; tile < p.tiles_m; tile += gridDim.x, ++ti) {
            const int m0 = tile * GEMM_BM;
            const int row = m0 + row_in_tile;
            const bool row_ok = row < p.M;
            const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
            // ---- epilogue 1 per hidden chunk: this warp converts columns [32 hh, 32 hh + 32) of the 64-wide chunk
            for (int j = 0; j < NC; ++j) {
                const int b = j & 1;
                const uint32_t use = (uint32_t)(ti * half_nc + (j >> 1));
                mbar_wait(&a1_full[b], use & 1u);
                tc_fence_after();
                uint32_t r[32];
                tmem_ld32(lane_base + MLP_ACC1_COL + (uint32_t)(b * MLP_NH + hh * 32), r);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a1_empty[b]);                        // acc1[b] is in registers now
                const float* bb = sb1 + j * MLP_NH + hh * 32;
                uint4 o[4];
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {
                    float v[8];
                    const float4 b0 = *reinterpret_cast<const float4*>(bb + g8 * 8);
                    const float4 b1v = *reinterpret_cast<const float4*>(bb + g8 * 8 + 4);
                    v[0] = gelu_erf(__uint_as_float(r[g8 * 8 + 0]) + b0.x); v[1] = gelu_erf(__uint_as_float(r[g8 * 8 + 1]) + b0.y);
                    v[2] = gelu_erf(__uint_as_float(r[g8 * 8 + 2]) + b0.z); v[3] = gelu_erf(__uint_as_float(r[g8 * 8 + 3]) + b0.w);
                    v[4] = gelu_erf(__uint_as_float(r[g8 * 8 + 4]) + b1v.x); v[5] = gelu_erf(__uint_as_float(r[g8 * 8 + 5]) + b1v.y);
                    v[6] = gelu_erf(__uint_as_float(r[g8 * 8 + 6]) + b1v.z); v[7] = gelu_erf(__uint_as_float(r[g8 * 8 + 7]) + b1v.w);
                    o[g8].x = pack_bf16x2(v[0], v[1]); o[g8].y = pack_bf16x2(v[2], v[3]);
                    o[g8].z = pack_bf16x2(v[4], v[5]); o[g8].w = pack_bf16x2(v[6], v[7]);
                }
                mbar_wait(&h_empty[b], (use & 1u) ^ 1u);                         // MMA2 of the previous use has read H[b]
                uint8_t* hrow = smemH + (size_t)b * GEMM_A_STAGE_BYTES + (size_t)row_in_tile * 128;
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) *reinterpret_cast<uint4*>(hrow + ((((uint32_t)(hh * 4 + g8)) ^ sw) << 4)) = o[g8];
                fence_proxy_async_smem();                                        // visible to the tensor core's smem reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&h_full[b]);
            }
            // ---- epilogue 2: acc2 -> +b2 -> +resid -> bf16 -> D; 64-column groups alternate between the two warps of a quarter
            const bf16* rrow = p.resid + (size_t)row * C;
            bf16* drow = p.D + (size_t)row * C;
            const int ngroups = (C + 63) / 64;
            uint4 rpre[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int col = hh * 64 + jj * 8;
                rpre[jj] = (row_ok && col + 8 <= C) ? *reinterpret_cast<const uint4*>(rrow + col) : make_uint4(0, 0, 0, 0);
            }
            mbar_wait(a2_full, (uint32_t)ti & 1u);
            tc_fence_after();
            int gk = 0;
            for (int g = hh; g < ngroups; g += 2, ++gk) {
                const int gcol = g * 64;
                const int gw = (C - gcol) < 64 ? (C - gcol) : 64;               // 32 for the tail group of C = 96
                const bool via_tma = gw == 64;
                uint32_t r[2][32];
                tmem_ld32(lane_base + (uint32_t)gcol, r[0]);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (c * 32 >= gw) break;
                    tmem_ld_wait();
                    if (c == 0 && gw == 64) tmem_ld32(lane_base + (uint32_t)(gcol + 32), r[1]);
#pragma unroll
                    for (int g8 = 0; g8 < 4; ++g8) {
                        const int col = gcol + c * 32 + g8 * 8;
                        float v[8];
                        const float4 b0 = *reinterpret_cast<const float4*>(sb2 + col);
                        const float4 b1v = *reinterpret_cast<const float4*>(sb2 + col + 4);
                        v[0] = __uint_as_float(r[c][g8 * 8 + 0]) + b0.x; v[1] = __uint_as_float(r[c][g8 * 8 + 1]) + b0.y;
                        v[2] = __uint_as_float(r[c][g8 * 8 + 2]) + b0.z; v[3] = __uint_as_float(r[c][g8 * 8 + 3]) + b0.w;
                        v[4] = __uint_as_float(r[c][g8 * 8 + 4]) + b1v.x; v[5] = __uint_as_float(r[c][g8 * 8 + 5]) + b1v.y;
                        v[6] = __uint_as_float(r[c][g8 * 8 + 6]) + b1v.z; v[7] = __uint_as_float(r[c][g8 * 8 + 7]) + b1v.w;
                        uint4 rv;
                        if (gk == 0) rv = rpre[c * 4 + g8];
                        else rv = row_ok ? *reinterpret_cast<const uint4*>(rrow + col) : make_uint4(0, 0, 0, 0);
                        const float2 r0 = unpack_bf16x2(rv.x), r1 = unpack_bf16x2(rv.y), r2 = unpack_bf16x2(rv.z), r3 = unpack_bf16x2(rv.w);
                        v[0] += r0.x; v[1] += r0.y; v[2] += r1.x; v[3] += r1.y;
                        v[4] += r2.x; v[5] += r2.y; v[6] += r3.x; v[7] += r3.y;
                        uint4 o;
                        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
                        o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
                        if (via_tma) {
                            if (c == 0 && g8 == 0) {
                                if (store_pending) { tma_store_wait_read0(); store_pending = false; }
                                __syncwarp();
                            }
                            *reinterpret_cast<uint4*>(stage + lane * 128 + ((((uint32_t)(c * 4 + g8)) ^ sw) << 4)) = o;
                        } else if (row_ok) {
                            *reinterpret_cast<uint4*>(drow + col) = o;
                        }
                    }
                }
                if (via_tma) {
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmD, stage, gcol, m0 + q * 32);
                        tma_store_commit();
                    }
                    store_pending = true;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a2_empty);
        }
        if (store_pending) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace fvhd
