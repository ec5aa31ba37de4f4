// MHSA core on the 5th-gen tensor cores (north_star (ii)): softmax((q * 32^-1/2) k^T) v per head of 32 (mci.py:661-685),
// flash-style, S and P.V accumulators in TMEM.
//
// The mma.sync kernel (stem_attn_se.cuh) sits on the legacy HMMA roof of the B200 (~330 MAC/clk/SM measured: 184 TFLOP/s at
// batch 32).  Here one CTA owns TWO 128-query tiles of one (image, head) -- two independent softmax warpgroups that share every
// K / V tile:
//   warp 0      TMA producer: Q (2 x [128 x 32]), then K_j / V_j tiles ([128 keys x 32], rows of 64 B, SWIZZLE_64B) through a 3-deep ring
//   warp 1      MMA issuer (warp-uniform): S_g(j) = Q_g . K_j^T  (M = 128, N = 128, K = 32: two tcgen05.mma) into TMEM, and
//               PV_g(j) = P_g(j) . V_j (M = 128, N = 32, K = 128: eight MMAs; A = P from 128-B-swizzled smem, B = V used IN PLACE as
//               an MN-major operand -- no transpose of V anywhere) into a double-buffered 32-column TMEM tile
//   warps 2-5   softmax warpgroup 0 (thread = query row of tile 0), warps 6-9 warpgroup 1: tcgen05.ld S (128 columns), running
//               max / sum in registers, P = exp2(S * c - m) -> bf16 -> smem (the A operand of P.V), O kept in REGISTERS:
//               O = (O + PV(j-1)) * 2^(m_{j-1} - m_j), so TMEM is never read-modified-written.
// Per key tile and CTA: 4 + 16 MMAs, 2 x 16384 exponentials -- the kernel is bound by the SFU (16 ex2/clk/SM), ~3.4x the
// mma.sync kernel's rate.
#pragma once
#include "gemm_tcgen05.cuh"

namespace fvhd {

struct AttU {
    static constexpr int HD = 32;                     // head dim (mci.py:636)
    static constexpr int QT = 128;                    // queries per warpgroup tile
    static constexpr int KT = 128;                    // keys per tile
    static constexpr int NST = 3;                     // K/V ring depth
    static constexpr int TILE_B = KT * HD * 2;        // 8192 B: [128 rows x 64 B]
    static constexpr int P_B = QT * KT * 2;           // 32768 B per P buffer (two SW128 blocks of 64 keys)
    static constexpr int THREADS = 320;
    static constexpr int TMEM_COLS = 512;             // S0 [0,128) S1 [128,256) PV0 2x32 [256,320) PV1 2x32 [320,384)
    static constexpr size_t SMEM = 2 * TILE_B + (size_t)NST * 2 * TILE_B + 2 * P_B + 256 + 1024;
};

// K-major operand with 64-byte rows (32 bf16), SWIZZLE_64B: 8-row atoms of 512 B.
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;                           // layout type SWIZZLE_64B
    return d;
}
// idesc with explicit majorness: bit 15 = A is MN-major, bit 16 = B is MN-major
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16_major(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ float ex2_approx(float x) {      // 2^x on the SFU; ex2(-inf) = +0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct AttUParams {
    bf16* out;               // [B*N, C]
    int N, C;                // tokens per image, channels
    int qpairs;              // ceil(N / 256)
    float scale_log2e;       // 32^-0.5 * log2(e)
};

__global__ void __launch_bounds__(AttU::THREADS, 1)
attention_umma_kernel(const __grid_constant__ CUtensorMap tmQKV /*[B*N, 3C] bf16, box {32, 128}, SWIZZLE_64B*/, const AttUParams p) {
    using A = AttU;
    extern __shared__ uint8_t att_smem_raw[];
    const uint32_t raw_addr = smem_u32(att_smem_raw);
    uint8_t* smem = att_smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* sQ = smem;                                   // 2 x 8 KB
    uint8_t* sK = sQ + 2 * A::TILE_B;                     // NST x 8 KB
    uint8_t* sV = sK + A::NST * A::TILE_B;                // NST x 8 KB
    uint8_t* sP = sV + A::NST * A::TILE_B;                // 2 x 32 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * A::P_B);
    uint64_t* q_full = bars;                  // [1]
    uint64_t* kv_full = bars + 1;             // [NST]
    uint64_t* kv_free = kv_full + A::NST;     // [NST]  both warpgroups' S and PV MMAs of the tile retired (2 commits)
    uint64_t* s_full = kv_free + A::NST;      // [2]    per warpgroup
    uint64_t* s_free = s_full + 2;            // [2]    (4 warps)
    uint64_t* p_full = s_free + 2;            // [2]    (4 warps)
    uint64_t* pv_full = p_full + 2;           // [2][2] per warpgroup, per PV buffer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_full + 4);

    pdl_launch_dependents();
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int qp = blockIdx.x, head = blockIdx.y, img = blockIdx.z;
    const int N = p.N, C = p.C;
    const int ntiles = (N + A::KT - 1) / A::KT;
    const int row0 = img * N;                                 // first row of this image in the [B*N, 3C] matrix
    const int q0 = qp * 2 * A::QT;                            // first query of this CTA

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        mbar_init(q_full, 1);
        for (int s = 0; s < A::NST; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_free[s], 2); }
        for (int g = 0; g < 2; ++g) {
            mbar_init(&s_full[g], 1); mbar_init(&s_free[g], 4); mbar_init(&p_full[g], 4);
            mbar_init(&pv_full[2 * g], 1); mbar_init(&pv_full[2 * g + 1], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, A::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            pdl_wait();                                       // qkv is the predecessor's output
            mbar_expect_tx(q_full, 2 * A::TILE_B);
            // rows past the image's last token belong to the next image (or are OOB): their scores are never stored
            tma_load_2d(sQ, &tmQKV, head * A::HD, row0 + q0, q_full);
            tma_load_2d(sQ + A::TILE_B, &tmQKV, head * A::HD, row0 + q0 + A::QT, q_full);
            for (int j = 0; j < ntiles; ++j) {
                const int s = j % A::NST;
                mbar_wait(&kv_free[s], ((uint32_t)(j / A::NST) & 1u) ^ 1u);
                mbar_expect_tx(&kv_full[s], 2 * A::TILE_B);
                tma_load_2d(sK + s * A::TILE_B, &tmQKV, C + head * A::HD, row0 + j * A::KT, &kv_full[s]);
                tma_load_2d(sV + s * A::TILE_B, &tmQKV, 2 * C + head * A::HD, row0 + j * A::KT, &kv_full[s]);
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (whole warp, uniform)
        const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t idesc_s = umma_idesc_bf16_major(A::QT, A::KT, 0, 0);        // Q (K-major) x K (K-major)
        const uint32_t idesc_pv = umma_idesc_bf16_major(A::QT, A::HD, 0, 1);       // P (K-major) x V (MN-major: [key][dim] as stored)
        const uint64_t dq0 = umma_desc_sw64(smem_u32(sQ));
        const uint64_t dk0 = umma_desc_sw64(smem_u32(sK));
        const uint64_t dv0 = umma_desc_sw64(smem_u32(sV));
        const uint64_t dp0 = umma_desc_sw128(smem_u32(sP));
        mbar_wait(q_full, 0);
        auto issue_s = [&](int j) {           // S_g(j) = Q_g . K_j^T for both warpgroups
            const int s = j % A::NST;
            mbar_wait(&kv_full[s], (uint32_t)(j / A::NST) & 1u);
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                mbar_wait(&s_free[g], ((uint32_t)j & 1u) ^ 1u);                     // softmax g has read S_g(j-1)
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 2; ++k)                                         // K = 32 dims = two k16 steps (+32 B inside the 64-B row)
                    if (elect_one())
                        umma_bf16(tm + (uint32_t)(g * A::KT), dq0 + (uint64_t)(g * (A::TILE_B >> 4) + 2 * k), dk0 + (uint64_t)(s * (A::TILE_B >> 4) + 2 * k),
                                  idesc_s, k != 0 ? 1u : 0u);
                if (elect_one()) umma_commit(&s_full[g]);
                __syncwarp();
            }
        };
        auto issue_pv = [&](int j) {          // PV_g(j) = P_g(j) . V_j for both warpgroups; releases the K/V slot
            const int s = j % A::NST;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                mbar_wait(&p_full[g], (uint32_t)j & 1u);                            // P_g(j) is in smem (and PV_g(j-2) has been read)
                tc_fence_after();
                const uint32_t acc = tm + 256u + (uint32_t)(g * 64 + (j & 1) * 32);
#pragma unroll
                for (int k = 0; k < 8; ++k)                                         // 128 keys = 8 k16 steps: P block k/4 (+32 B per step), V +16 keys (1024 B)
                    if (elect_one())
                        umma_bf16(acc, dp0 + (uint64_t)(g * (A::P_B >> 4) + (k >> 2) * (16384 >> 4) + 2 * (k & 3)),
                                  dv0 + (uint64_t)(s * (A::TILE_B >> 4) + k * (1024 >> 4)), idesc_pv, k != 0 ? 1u : 0u);
                if (elect_one()) { umma_commit(&pv_full[2 * g + (j & 1)]); umma_commit(&kv_free[s]); }
                __syncwarp();
            }
        };
        issue_s(0);
        for (int j = 0; j < ntiles; ++j) {
            if (j + 1 < ntiles) issue_s(j + 1);
            issue_pv(j);
        }
    } else {
        // ---------------- softmax warpgroups: g = 0 (warps 2-5) / 1 (warps 6-9); thread = one query row
        const int g = (warp - 2) >> 2;
        const int q = warp & 3;                                   // TMEM lane quarter (hardware: warp % 4)
        const int rloc = q * 32 + lane;                           // row inside the 128-query tile
        const int qrow = q0 + g * A::QT + rloc;                   // query index inside the image
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t s_col = (uint32_t)(g * A::KT), pv_col = 256u + (uint32_t)(g * 64);
        uint8_t* prow = sP + (size_t)g * A::P_B + (size_t)rloc * 128;
        const uint32_t sw = (uint32_t)(lane & 7);
        float o[A::HD];
#pragma unroll
        for (int i = 0; i < A::HD; ++i) o[i] = 0.f;
        float m = -INFINITY, l = 0.f;
        const float c = p.scale_log2e;
        for (int j = 0; j < ntiles; ++j) {
            // fold in PV(j-1) (accumulated relative to the previous max) before the scores occupy the registers
            if (j > 0) {
                mbar_wait(&pv_full[2 * g + ((j - 1) & 1)], (uint32_t)((j - 1) >> 1) & 1u);
                tc_fence_after();
                uint32_t pv[32];
                tmem_ld32(lane_base + pv_col + (uint32_t)(((j - 1) & 1) * 32), pv);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < A::HD; ++i) o[i] += __uint_as_float(pv[i]);
            }
            mbar_wait(&s_full[g], (uint32_t)j & 1u);
            tc_fence_after();
            uint32_t sv[128];
            tmem_ld32(lane_base + s_col, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
            tmem_ld32(lane_base + s_col + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
            tmem_ld32(lane_base + s_col + 64, *reinterpret_cast<uint32_t(*)[32]>(&sv[64]));
            tmem_ld32(lane_base + s_col + 96, *reinterpret_cast<uint32_t(*)[32]>(&sv[96]));
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[g]);               // S_g is in registers: the next tile's scores may overwrite it
            const int kvalid = N - j * A::KT;                     // keys of this tile that exist (>= 128 except for the last tile)
            if (kvalid < A::KT) {                                 // warp-uniform: only the last tile can be ragged
#pragma unroll
                for (int i = 0; i < 128; ++i) sv[i] = i < kvalid ? sv[i] : 0xff800000u;      // -inf
            }
            // running max on the RAW scores (the scale c > 0 commutes with max); four independent chains
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int i = 0; i < 128; i += 4) {
                mx0 = fmaxf(mx0, __uint_as_float(sv[i])); mx1 = fmaxf(mx1, __uint_as_float(sv[i + 1]));
                mx2 = fmaxf(mx2, __uint_as_float(sv[i + 2])); mx3 = fmaxf(mx3, __uint_as_float(sv[i + 3]));
            }
            const float mx = fmaxf(m, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * c);      // m, mx: scaled (log2 domain)
            const float alpha = ex2_approx(m - mx);               // 0 for the first tile (m = -inf); rescale O to the new max
            m = mx;
#pragma unroll
            for (int i = 0; i < A::HD; ++i) o[i] *= alpha;
            // P = exp2(s c - m) -> bf16 -> smem: key kk of the tile lives in SW128 block kk / 64, 16-B chunk (kk % 64) / 8 ^ (row & 7)
            float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
            const float nmx = -mx;
#pragma unroll
            for (int c8 = 0; c8 < 16; ++c8) {
                float pe[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) pe[e] = ex2_approx(fmaf(__uint_as_float(sv[c8 * 8 + e]), c, nmx));
                rs0 += pe[0] + pe[4]; rs1 += pe[1] + pe[5]; rs2 += pe[2] + pe[6]; rs3 += pe[3] + pe[7];
                *reinterpret_cast<uint4*>(prow + (c8 >> 3) * 16384 + ((((uint32_t)(c8 & 7)) ^ sw) << 4)) =
                    make_uint4(pack_bf16x2(pe[0], pe[1]), pack_bf16x2(pe[2], pe[3]), pack_bf16x2(pe[4], pe[5]), pack_bf16x2(pe[6], pe[7]));
            }
            const float rs = (rs0 + rs1) + (rs2 + rs3);
            l = l * alpha + rs;
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[g]);
        }
        {   // last tile's P.V
            const int j = ntiles - 1;
            mbar_wait(&pv_full[2 * g + (j & 1)], (uint32_t)(j >> 1) & 1u);
            tc_fence_after();
            uint32_t pv[32];
            tmem_ld32(lane_base + pv_col + (uint32_t)((j & 1) * 32), pv);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < A::HD; ++i) o[i] += __uint_as_float(pv[i]);
        }
        pdl_wait();                                               // global writes after the predecessor
        if (qrow < N) {
            const float inv = 1.0f / l;
            uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)row0 + qrow) * C + head * A::HD);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                dst[i] = make_uint4(pack_bf16x2(o[8 * i] * inv, o[8 * i + 1] * inv), pack_bf16x2(o[8 * i + 2] * inv, o[8 * i + 3] * inv),
                                    pack_bf16x2(o[8 * i + 4] * inv, o[8 * i + 5] * inv), pack_bf16x2(o[8 * i + 6] * inv, o[8 * i + 7] * inv));
        }
        tc_fence_before();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, A::TMEM_COLS);
    }
}

}  // namespace fvhd
