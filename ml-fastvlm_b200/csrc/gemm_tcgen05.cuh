// D[M,N] = epilogue( A[M,K] . W[N,K]^T ) on the 5th-gen tensor cores.
//
// Every pointwise (1x1) conv, Linear and projector layer of the FastViTHD path is this kernel:
// NHWC activations ARE the row-major [pixels, C] A operand, so no im2col / transpose exists.
//   ConvFFN.fc1/fc2 (mci.py:922-926), MobileOneBlock 1x1 (mci.py:591-602, 727-737),
//   MHSA.qkv / MHSA.proj (mci.py:669-681), mm_projector (multimodal_projector/builder.py:23-30).
//
// Persistent, warp-specialised (320 threads, one CTA per SM, CTA c walks tiles c, c+G, ...):
//   warp 0      : TMA producer  -- cp.async.bulk.tensor 2D, 128-B swizzled 128x64 (A) / BNx64 (W) bf16 boxes
//                 into a `stages`-deep smem ring (full/empty mbarriers)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (M=128, N=BN, K=16 per instruction);
//                 tcgen05.commit frees ring slots and publishes finished accumulators
//   warps 2..9  : epilogue -- two accumulator buffers in TMEM, so the epilogue of tile i (tcgen05.ld ->
//                 +bias -> GELU -> +residual -> bf16 -> global) overlaps the MMAs of tile i+1.  A warp may
//                 only touch TMEM lanes 32*(warp%4)..+31, so warps 2-5 and 6-9 split the columns.
#pragma once
#include "ptx.cuh"

namespace fvhd {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_THREADS = 32 * (2 + GEMM_EPI_WARPS);
constexpr int GEMM_MAX_STAGES = 8;
constexpr int GEMM_A_STAGE_BYTES = GEMM_BM * GEMM_BK * 2;   // 16 KiB
constexpr int GEMM_SMEM_BUDGET = 200 * 1024;

struct GemmParams {
    int M, N, K;
    int BN;            // N tile: multiple of 32, 32..256
    int stages;        // 1..GEMM_MAX_STAGES
    int tiles_m, tiles_n;
    bf16* D;           // [M, ldd] bf16; nullptr => io->final_out (caller memory)
    const IoBlock* io;
    int ldd;
    const float* bias;       // [N] fp32 or nullptr
    const bf16* residual;    // [M, ldr] bf16 or nullptr (added after activation)
    int ldr;
    int act;                 // 0 = identity, 1 = exact-erf GELU
};

__host__ __device__ inline int gemm_acc_stride(int bn) {     // TMEM columns per accumulator buffer (power of 2)
    int c = 32;
    while (c < bn) c <<= 1;
    return c;
}
__host__ inline size_t gemm_smem_bytes(int bn, int stages) {
    return (size_t)stages * (GEMM_A_STAGE_BYTES + (size_t)bn * GEMM_BK * 2) + 1024 /*align slack*/ + 256 /*barriers*/;
}
__host__ inline int gemm_pick_stages(int bn, int num_kb) {
    int s = GEMM_SMEM_BUDGET / (GEMM_A_STAGE_BYTES + bn * GEMM_BK * 2);
    if (s > GEMM_MAX_STAGES) s = GEMM_MAX_STAGES;
    (void)num_kb;
    return s < 1 ? 1 : s;
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    extern __shared__ uint8_t gemm_smem_raw[];
    const uint32_t raw_addr = smem_u32(gemm_smem_raw);
    const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
    uint8_t* smem = gemm_smem_raw + pad;                       // 1024-B aligned (SWIZZLE_128B atoms)

    const int stages = p.stages;
    const int BN = p.BN;
    const uint32_t b_stage_bytes = (uint32_t)BN * GEMM_BK * 2;
    uint8_t* smemA = smem;
    uint8_t* smemB = smem + (size_t)stages * GEMM_A_STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smemB + (size_t)stages * b_stage_bytes);
    uint64_t* empty_bar = full_bar + GEMM_MAX_STAGES;
    uint64_t* tfull_bar = empty_bar + GEMM_MAX_STAGES;          // [2] accumulator ready
    uint64_t* tempty_bar = tfull_bar + 2;                       // [2] accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
    const int num_tiles = p.tiles_m * p.tiles_n;
    const uint32_t acc_stride = gemm_acc_stride(BN);
    const uint32_t tmem_cols = 2 * acc_stride;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], GEMM_EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / p.tiles_n) * GEMM_BM;
                const int n0 = (tile % p.tiles_n) * BN;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % stages;
                    const uint32_t ph = (uint32_t)(it / stages) & 1u;
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    mbar_expect_tx(&full_bar[s], GEMM_A_STAGE_BYTES + b_stage_bytes);
                    tma_load_2d(smemA + (size_t)s * GEMM_A_STAGE_BYTES, &tmA, kb * GEMM_BK, m0, &full_bar[s]);
                    tma_load_2d(smemB + (size_t)s * b_stage_bytes, &tmB, kb * GEMM_BK, n0, &full_bar[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------- MMA issuer (one thread drives the tensor core for the whole CTA)
            const uint32_t idesc = umma_idesc_bf16(GEMM_BM, (uint32_t)BN);
            int it = 0, tl = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
                const int a = tl & 1;
                const uint32_t aph = (uint32_t)(tl >> 1) & 1u;
                mbar_wait(&tempty_bar[a], aph ^ 1u);            // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)a * acc_stride;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % stages;
                    const uint32_t ph = (uint32_t)(it / stages) & 1u;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128(smem_u32(smemA + (size_t)s * GEMM_A_STAGE_BYTES));
                    const uint64_t db = umma_desc_sw128(smem_u32(smemB + (size_t)s * b_stage_bytes));
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        // advance 16 bf16 = 32 B along K inside the 128-B swizzle row: +2 in (addr >> 4) units
                        umma_bf16(acc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[s]);      // slot reusable once these MMAs have read it
                }
                umma_commit(&tfull_bar[a]);          // accumulator complete
            }
        }
    } else {
        // ---------------- epilogue: warp w owns TMEM lanes [32*(w%4), +32) == tile rows; warps 2-5 take the first
        // half of the 32-column chunks, warps 6-9 the rest.
        const int q = warp & 3;
        const int hh = (warp - 2) >> 2;
        const int nchunks = BN / 32;
        const int c_begin = hh == 0 ? 0 : (nchunks + 1) / 2;
        const int c_end = hh == 0 ? (nchunks + 1) / 2 : nchunks;
        int tl = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
            const int a = tl & 1;
            const uint32_t aph = (uint32_t)(tl >> 1) & 1u;
            const int m0 = (tile / p.tiles_n) * GEMM_BM;
            const int n0 = (tile % p.tiles_n) * BN;
            const int row = m0 + q * 32 + lane;
            const bool row_ok = row < p.M;
            bf16* dbase = p.D ? p.D : reinterpret_cast<bf16*>(p.io->final_out);
            bf16* drow = dbase + (size_t)row * p.ldd;
            const bf16* rrow = p.residual ? p.residual + (size_t)row * p.ldr : nullptr;
            mbar_wait(&tfull_bar[a], aph);
            tc_fence_after();
            const uint32_t lane_addr = tmem_base + (uint32_t)a * acc_stride + ((uint32_t)(q * 32) << 16);
            for (int c = c_begin; c < c_end; ++c) {
                uint32_t r[32];
                tmem_ld32(lane_addr + (uint32_t)(c * 32), r);
                tmem_ld_wait();
                const int col0 = n0 + c * 32;
                if (!row_ok || col0 >= p.N) continue;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int col = col0 + g * 8;
                    if (col + 8 > p.N) break;                  // N % 8 == 0 is enforced on the host
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]);
                    if (p.bias) {
                        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
                        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
                    }
                    if (p.act == 1) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = gelu_erf(v[j]);
                    }
                    if (rrow) {
                        const uint4 rv = *reinterpret_cast<const uint4*>(rrow + col);
                        const float2 r0 = unpack_bf16x2(rv.x), r1 = unpack_bf16x2(rv.y), r2 = unpack_bf16x2(rv.z), r3 = unpack_bf16x2(rv.w);
                        v[0] += r0.x; v[1] += r0.y; v[2] += r1.x; v[3] += r1.y;
                        v[4] += r2.x; v[5] += r2.y; v[6] += r3.x; v[7] += r3.y;
                    }
                    uint4 o;
                    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
                    o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
                    *reinterpret_cast<uint4*>(drow + col) = o;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[a]);     // this warp is done reading accumulator a
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

}  // namespace fvhd
