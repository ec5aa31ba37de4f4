// Fused ConvFFN (mci.py:922-926) for C = 384 (stage 2, 24 blocks = half of the forward at batch 1):
//     out = resid + fc2( GELU( fc1(z) + b1 ) ) + b2            (layer scale folded into fc2)
//
// A 128-pixel tile has only 32 CTAs' worth of rows at 64 x 64 pixels, and one CTA cannot hold the z tile, a W chunk ring and
// the 1536-wide hidden.  So a CLUSTER OF 4 CTAs owns one tile and splits the HIDDEN dimension: CTA r computes hidden slice
// [384 r, 384 r + 384) with the same chunk pipeline as mlp_fused.cuh,
//
//   z tile [128 x 384] --TMA--> smem (resident, 96 KB)                      (each CTA of the cluster loads it: L2 hit)
//   for each 64-wide chunk j of the slice (6 chunks):
//       MMA1  acc1[j&1] (TMEM 64 cols)  = z . W1[384 r + 64 j ..]^T          (K = 384, two 24-KB half-chunk slots)
//       epi1  acc1 -> +b1 -> GELU -> bf16 -> H[j&1] (smem, K-major SW128)
//       MMA2  acc2 (TMEM 384 cols)     += H[j&1] . W2[:, 384 r + 64 j ..]^T  (two N = 192 halves, one slot each)
//
// and the four partial fc2 accumulators (fp32, TMEM) are reduce-scattered THROUGH DISTRIBUTED SHARED MEMORY: CTA q owns output
// columns [96 q, 96 q + 96); every CTA converts its partial of a peer's columns to f16 (saturating), stages it in the layout the
// receiver reads ([8-column group][row] x 16 B, overlaid on the then-dead z / H smem) and sends it with ONE 24-KB
// cp.async.bulk.shared::cluster per peer that completes on the peer's mbarrier; q then adds its own fp32 partial, the three
// received ones, b2 and the residual and stores bf16.  Per CTA this moves 96 KB (z) + 2 x 288 KB (W1, W2 slices) through TMA and
// 72 KB each way through DSMEM, instead of the two GEMM launches' 196 KB / 786 KB per tile over 2.6 / 1 waves -- and the
// 12.6 MB hidden never exists.  (fp32 partials over st.shared::cluster + fence.acq_rel.cluster were measured first: 9.7 us of
// reduction tail against 6.1 us for this scheme; the DSMEM path itself moves ~13 B/clk per SM.)
//
// Synchronisation across the cluster uses mbarriers with remote arrives only (no barrier.cluster after start-up, so the
// single-lane TMA / MMA warps never have to take part):
//   ready_bar (count 4, remote arrives): "CTA x has retired all its MMAs of this tile" -> x's smem may be overwritten
//   recv_bar  (count 1 + 72 KB of complete_tx): own expect_tx + the three peers' bulk copies have landed
//   ack_bar   (count 3, remote arrives): the peers have received what this CTA staged -> staging may be recycled / CTA may exit
//   tile_done (count 8 warps, local): receive buffer consumed -> the producer may load the next tile's z / W
#pragma once
#include "mlp_fused.cuh"

namespace fvhd {

constexpr int MLPC_C = 384;
constexpr int MLPC_CS = 4;                                   // CTAs per cluster == hidden slices
constexpr int MLPC_KB = MLPC_C / 64;                         // 6 k-blocks of z / W1
constexpr int MLPC_NC = 4 * MLPC_C / MLPC_CS / MLP_NH;       // 6 hidden chunks per CTA
constexpr int MLPC_OWN = MLPC_C / MLPC_CS;                   // 96 output columns reduced + stored by each CTA
constexpr int MLPC_NHALF = MLPC_C / 2;                       // 192: N of one MMA2 half
constexpr int MLPC_SLOTS = 4;
constexpr int MLPC_SLOT_BYTES = 24576;                       // half a W1 chunk (3 k-blocks x 64 rows) == half a W2 chunk (192 rows)
constexpr int MLPC_ACC1_COL = MLPC_C;                        // TMEM: acc2 at [0, 384), acc1[b] at 384 + 64 b
constexpr int MLPC_THREADS = MLP_THREADS;
constexpr int MLPC_Z_BYTES = MLPC_KB * GEMM_A_STAGE_BYTES;   // 98304
constexpr int MLPC_H_BYTES = 2 * GEMM_A_STAGE_BYTES;         // 32768
constexpr int MLPC_W_BYTES = MLPC_SLOTS * MLPC_SLOT_BYTES;   // 98304
constexpr int MLPC_PEER_BYTES = GEMM_BM * MLPC_OWN * 2;      // 24576: one peer's f16 partial of the 96 columns a CTA owns
constexpr int MLPC_RECV_BYTES = (MLPC_CS - 1) * MLPC_PEER_BYTES;   // 73728 received, overlays the (dead) z tile ...
constexpr int MLPC_STAGE_OFF = MLPC_RECV_BYTES;              // ... followed by 73728 staged for the peers (z / H)
constexpr size_t MLPC_SMEM = (size_t)MLPC_Z_BYTES + MLPC_H_BYTES + MLPC_W_BYTES + MLPC_OWN * 4 + 256 + 1024;
static_assert(MLPC_SMEM <= 227 * 1024, "cluster ConvFFN smem");
static_assert(2 * MLPC_RECV_BYTES <= MLPC_Z_BYTES + MLPC_H_BYTES + MLPC_W_BYTES, "receive + staging overlay");
static_assert(MLPC_NC % 2 == 0, "acc1 / H double buffering assumes an even chunk count");

// debug timeline (fvhd_convffn with a trace buffer): stamp i of this CTA
#define MLPC_TRACE(i) do { if (p.trace) p.trace[(size_t)blockIdx.x * 64 + (i)] = gtime(); } while (0)

__global__ void __launch_bounds__(MLPC_THREADS, 1)
mlp_cluster_tcgen05_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmW1,
                           const __grid_constant__ CUtensorMap tmW2, const MlpParams p) {
    extern __shared__ uint8_t mlpc_smem_raw[];
    const uint32_t raw_addr = smem_u32(mlpc_smem_raw);
    uint8_t* smem = mlpc_smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);       // same offset in every CTA of the cluster

    constexpr int C = MLPC_C;
    uint8_t* smemZ = smem;
    uint8_t* smemH = smemZ + MLPC_Z_BYTES;
    uint8_t* smemW = smemH + MLPC_H_BYTES;
    float* sb2 = reinterpret_cast<float*>(smemW + MLPC_W_BYTES);                  // b2 of the 96 columns this CTA owns
    uint64_t* bars = reinterpret_cast<uint64_t*>(sb2 + MLPC_OWN);
    uint64_t* z_full = bars;                    // [1]
    uint64_t* tile_done = bars + 1;             // [1]
    uint64_t* w_full = bars + 2;                // [MLPC_SLOTS]
    uint64_t* w_empty = w_full + MLPC_SLOTS;    // [MLPC_SLOTS]
    uint64_t* a1_full = w_empty + MLPC_SLOTS;   // [2]
    uint64_t* a1_empty = a1_full + 2;           // [2]
    uint64_t* h_full = a1_empty + 2;            // [2]
    uint64_t* h_empty = h_full + 2;             // [2]
    uint64_t* a2_full = h_empty + 2;            // [1]
    uint64_t* a2_empty = a2_full + 1;           // [1]
    uint64_t* ready_bar = a2_empty + 1;         // [1]  remote arrives
    uint64_t* recv_bar = ready_bar + 1;         // [1]  own expect_tx + the peers' bulk copies (complete_tx)
    uint64_t* ack_bar = recv_bar + 1;           // [1]  remote arrives: the peers have received what I staged
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ack_bar + 1);

    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) MLPC_TRACE(0);
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = (int)(blockIdx.x / MLPC_CS);
    const int num_clusters = (int)(gridDim.x / MLPC_CS);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmZ); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
        mbar_init(z_full, 1); mbar_init(tile_done, GEMM_EPI_WARPS);
        for (int s = 0; s < MLPC_SLOTS; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&a1_full[b], 1); mbar_init(&a1_empty[b], GEMM_EPI_WARPS);
            mbar_init(&h_full[b], GEMM_EPI_WARPS); mbar_init(&h_empty[b], 1);
        }
        mbar_init(a2_full, 1); mbar_init(a2_empty, GEMM_EPI_WARPS);
        mbar_init(ready_bar, MLPC_CS);
        mbar_init(recv_bar, 1);
        mbar_init(ack_bar, MLPC_CS - 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < MLPC_OWN; i += MLPC_THREADS) sb2[i] = __ldg(p.b2 + rank * MLPC_OWN + i);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                          // every CTA's barriers exist before any remote arrive / store
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) MLPC_TRACE(1);
    const int hid0 = (int)rank * (MLPC_NC * MLP_NH);                              // first hidden unit of this CTA's slice

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer: z tile, then half-chunk W boxes in the MMA thread's consumption order
            int wit = 0;
            auto load_w = [&](bool is_w2, int j, int u) {
                const int s = wit % MLPC_SLOTS;
                const uint32_t ph = (uint32_t)(wit / MLPC_SLOTS) & 1u;
                ++wit;
                mbar_wait(&w_empty[s], ph ^ 1u);
                uint8_t* dst = smemW + (size_t)s * MLPC_SLOT_BYTES;
                if (u == 0) MLPC_TRACE((is_w2 ? 54 : 48) + j);                     // slot free -> load issued
                mbar_expect_tx(&w_full[s], MLPC_SLOT_BYTES);
                if (!is_w2) {           // W1 rows [hid0 + 64 j, +64), k-blocks 3u .. 3u+2: three 64 x 64 boxes
                    for (int kk = 0; kk < 3; ++kk)
                        tma_load_2d(dst + (size_t)kk * 64 * 128, &tmW1, (3 * u + kk) * 64, hid0 + j * MLP_NH, &w_full[s]);
                } else {                // W2 rows [192 u, +192), K columns [hid0 + 64 j, +64): one 192 x 64 box
                    tma_load_2d(dst, &tmW2, hid0 + j * MLP_NH, u * MLPC_NHALF, &w_full[s]);
                }
            };
            // half-chunk q of a tile's weight stream, in consumption order: W1[0] x2, then (W1[j] x2, W2[j-1] x2) j = 1..NC-1, W2[NC-1] x2
            auto load_seq = [&](int q) {
                if (q < 2) { load_w(false, 0, q); return; }
                if (q >= 4 * MLPC_NC - 2) { load_w(true, MLPC_NC - 1, q - (4 * MLPC_NC - 2)); return; }
                const int g = (q - 2) >> 2, r = (q - 2) & 3;
                if (r < 2) load_w(false, g + 1, r); else load_w(true, g, r - 2);
            };
            // weights are constants of the forward: the first ring fill precedes the PDL wait (only z depends on the predecessor)
            int q0 = 0;
            if (cluster_id < p.tiles_m)
                for (; q0 < MLPC_SLOTS; ++q0) load_seq(q0);
            pdl_wait();
            int ti = 0;
            for (int tile = cluster_id; tile < p.tiles_m; tile += num_clusters, ++ti) {
                mbar_wait(tile_done, ((uint32_t)ti & 1u) ^ 1u);                   // previous tile's receive buffer (z overlay) consumed ...
                mbar_wait_cluster(ack_bar, ((uint32_t)ti & 1u) ^ 1u);             // ... and its staged partials fetched by the peers
                MLPC_TRACE(2);
                mbar_expect_tx(z_full, MLPC_Z_BYTES);
                for (int kb = 0; kb < MLPC_KB; ++kb) tma_load_2d(smemZ + (size_t)kb * GEMM_A_STAGE_BYTES, &tmZ, kb * 64, tile * GEMM_BM, z_full);
                for (int q = ti == 0 ? q0 : 0; q < 4 * MLPC_NC; ++q) load_seq(q);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------- MMA issuer
            const uint32_t idesc1 = umma_idesc_bf16(GEMM_BM, MLP_NH);
            const uint32_t idesc2 = umma_idesc_bf16(GEMM_BM, MLPC_NHALF);
            constexpr int half_nc = MLPC_NC / 2;
            int wit = 0, ti = 0;
            auto take_slot = [&]() -> uint32_t {
                const int s = wit % MLPC_SLOTS;
                const uint32_t ph = (uint32_t)(wit / MLPC_SLOTS) & 1u;
                ++wit;
                mbar_wait(&w_full[s], ph);
                return (uint32_t)s;
            };
            auto mma2 = [&](int j, int ti_) {    // acc2 (+)= H[j&1] . W2[:, chunk j]^T, as two N = 192 halves
                const int b = j & 1;
                const uint32_t use = (uint32_t)(ti_ * half_nc + (j >> 1));
                mbar_wait(&h_full[b], use & 1u);
                if (j == 0) mbar_wait(a2_empty, ((uint32_t)ti_ & 1u) ^ 1u);     // previous tile's reduction drained acc2
                const uint64_t da = umma_desc_sw128(smem_u32(smemH + (size_t)b * GEMM_A_STAGE_BYTES));
                for (int u = 0; u < 2; ++u) {
                    const uint32_t s = take_slot();
                    tc_fence_after();
                    const uint64_t db = umma_desc_sw128(smem_u32(smemW + (size_t)s * MLPC_SLOT_BYTES));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_base + (uint32_t)(u * MLPC_NHALF), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc2, (j | k) != 0 ? 1u : 0u);
                    umma_commit(&w_empty[s]);
                }
                umma_commit(&h_empty[b]);
                MLPC_TRACE(16 + j);                                              // MMA2(j) issued
            };
            for (int tile = cluster_id; tile < p.tiles_m; tile += num_clusters, ++ti) {
                mbar_wait(z_full, (uint32_t)ti & 1u);
                MLPC_TRACE(3);
                for (int j = 0; j < MLPC_NC; ++j) {
                    const int b = j & 1;
                    const uint32_t use = (uint32_t)(ti * half_nc + (j >> 1));
                    mbar_wait(&a1_empty[b], (use & 1u) ^ 1u);                    // epilogue has drained acc1[b]
                    const uint32_t acc1 = tmem_base + MLPC_ACC1_COL + (uint32_t)b * MLP_NH;
                    for (int u = 0; u < 2; ++u) {
                        const uint32_t s = take_slot();
                        tc_fence_after();
#pragma unroll
                        for (int kk = 0; kk < 12; ++kk) {
                            const int kb = 3 * u + (kk >> 2), k = kk & 3;
                            const uint64_t da = umma_desc_sw128(smem_u32(smemZ + (size_t)kb * GEMM_A_STAGE_BYTES)) + (uint64_t)(2 * k);
                            const uint64_t db = umma_desc_sw128(smem_u32(smemW + (size_t)s * MLPC_SLOT_BYTES + (size_t)(kk >> 2) * 64 * 128)) + (uint64_t)(2 * k);
                            umma_bf16(acc1, da, db, idesc1, (u | kk) != 0 ? 1u : 0u);
                        }
                        umma_commit(&w_empty[s]);
                    }
                    umma_commit(&a1_full[b]);
                    MLPC_TRACE(8 + j);                                           // MMA1(j) issued
                    if (j >= 1) mma2(j - 1, ti);
                }
                mma2(MLPC_NC - 1, ti);
                umma_commit(a2_full);
            }
        }
    } else {
        // ---------------- epilogue warps: lane quarter q == 32 tile rows; warps 2-5 / 6-9 split the columns
        const int q = warp & 3;
        const int hh = (warp - 2) >> 2;
        constexpr int half_nc = MLPC_NC / 2;
        const int row_in_tile = q * 32 + lane;
        const uint32_t sw = (uint32_t)(lane & 7);
        const uint32_t recv_local = smem_u32(smem);
        pdl_wait();
        int ti = 0;
        for (int tile = cluster_id; tile < p.tiles_m; tile += num_clusters, ++ti) {
            const int row = tile * GEMM_BM + row_in_tile;
            const bool row_ok = row < p.M;
            const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
            // ---- epilogue 1 per hidden chunk: this warp converts columns [32 hh, 32 hh + 32) of the 64-wide chunk
            for (int j = 0; j < MLPC_NC; ++j) {
                const int b = j & 1;
                const uint32_t use = (uint32_t)(ti * half_nc + (j >> 1));
                const float4* bb = reinterpret_cast<const float4*>(p.b1 + hid0 + j * MLP_NH + hh * 32);
                float4 bv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) bv[i] = __ldg(bb + i);                // constants: issued before the wait
                mbar_wait(&a1_full[b], use & 1u);
                if (warp == 2 && lane == 0) MLPC_TRACE(24 + j);                  // acc1(j) ready
                tc_fence_after();
                uint32_t r[32];
                tmem_ld32(lane_base + MLPC_ACC1_COL + (uint32_t)(b * MLP_NH + hh * 32), r);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a1_empty[b]);                        // acc1[b] is in registers now
                uint4 o[4];
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {
                    float v[8];
                    const float4 b0 = bv[2 * g8], b1v = bv[2 * g8 + 1];
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
                if (warp == 2 && lane == 0) MLPC_TRACE(32 + j);                  // H(j) written
            }
            // ---- cluster reduction of the four partial acc2.  Residual of the columns this warp will finish: prefetch now.
            const bf16* rrow = p.resid + (size_t)row * C + rank * MLPC_OWN;
            bf16* drow = p.D + (size_t)row * C + rank * MLPC_OWN;
            uint4 rpre[6];                                                        // this warp finishes own columns [48 hh, 48 hh + 48)
#pragma unroll
            for (int i = 0; i < 6; ++i) rpre[i] = row_ok ? *reinterpret_cast<const uint4*>(rrow + hh * 48 + i * 8) : make_uint4(0, 0, 0, 0);
            mbar_wait(a2_full, (uint32_t)ti & 1u);                               // all MMAs of this CTA retired: acc2 final, smem dead
            tc_fence_after();
            if (warp == 2 && lane == 0) { MLPC_TRACE(40); mbar_expect_tx(recv_bar, MLPC_RECV_BYTES); }   // 3 peers x 24 KB will land here
            if (warp == 2 && lane < MLPC_CS) mbar_arrive_cluster(cluster_map(smem_u32(ready_bar), (uint32_t)lane));
            mbar_wait_cluster(ready_bar, (uint32_t)ti & 1u);                     // ... and the same holds for all four CTAs
            if (warp == 2 && lane == 0) MLPC_TRACE(41);
            // stage: 32-column group g belongs to CTA g / 3; f16 partials in the receiver's layout
            //        [peer slot (3)][8-column group c8 (12)][row (128)] x 16 B  -> one 24-KB bulk copy per peer
            for (int g = hh; g < C / 32; g += 2) {
                const uint32_t dst = (uint32_t)g / 3u;
                if (dst == rank) continue;
                uint32_t r[32];
                tmem_ld32(lane_base + (uint32_t)(g * 32), r);
                tmem_ld_wait();
                const uint32_t d = dst < rank ? dst : dst - 1;                    // my staging buffer for that peer
                uint4* srow = reinterpret_cast<uint4*>(smem + MLPC_STAGE_OFF) + (d * 12u + (uint32_t)(g % 3) * 4u) * GEMM_BM + row_in_tile;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint4 o;
                    o.x = pack_f16x2_sat(__uint_as_float(r[8 * i + 0]), __uint_as_float(r[8 * i + 1]));
                    o.y = pack_f16x2_sat(__uint_as_float(r[8 * i + 2]), __uint_as_float(r[8 * i + 3]));
                    o.z = pack_f16x2_sat(__uint_as_float(r[8 * i + 4]), __uint_as_float(r[8 * i + 5]));
                    o.w = pack_f16x2_sat(__uint_as_float(r[8 * i + 6]), __uint_as_float(r[8 * i + 7]));
                    srow[i * GEMM_BM] = o;
                }
            }
            fence_proxy_async_smem();                                            // staging is read by the bulk-copy engine
            named_bar_sync(1, GEMM_EPI_WARPS * 32);                              // all eight epilogue warps have staged
            if (warp == 2 && lane < MLPC_CS && (uint32_t)lane != rank) {
                const uint32_t dst = (uint32_t)lane;
                const uint32_t d = dst < rank ? dst : dst - 1;
                const uint32_t slot = rank < dst ? rank : rank - 1;              // where the peer expects my partial
                bulk_copy_to_cluster(cluster_map(recv_local + slot * MLPC_PEER_BYTES, dst), smem + MLPC_STAGE_OFF + d * MLPC_PEER_BYTES,
                                     MLPC_PEER_BYTES, cluster_map(smem_u32(recv_bar), dst));
            }
            if (warp == 2 && lane == 0) MLPC_TRACE(42);
            mbar_wait(recv_bar, (uint32_t)ti & 1u);                              // the three peers' partials of my columns are here
            if (warp == 2 && lane == 0) MLPC_TRACE(43);
            // the peers may now recycle the staging buffers they sent from
            if (warp == 2 && lane < MLPC_CS && (uint32_t)lane != rank) mbar_arrive_cluster(cluster_map(smem_u32(ack_bar), (uint32_t)lane));
            // finish the columns this CTA owns: own fp32 partial + 3 received + b2 + resid -> bf16
            const uint4* recv = reinterpret_cast<const uint4*>(smem);
#pragma unroll
            for (int u = 0; u < 3; ++u) {                                         // three 16-column units per warp
                const int col = hh * 48 + u * 16;                                 // within the 96 own columns
                uint32_t r[16];
                tmem_ld16(lane_base + (uint32_t)((int)rank * MLPC_OWN + col), r);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c8 = (col >> 3) + e;
                    const uint4 x0 = recv[(0 * 12 + c8) * GEMM_BM + row_in_tile];
                    const uint4 x1 = recv[(1 * 12 + c8) * GEMM_BM + row_in_tile];
                    const uint4 x2 = recv[(2 * 12 + c8) * GEMM_BM + row_in_tile];
                    const float4 bq0 = *reinterpret_cast<const float4*>(sb2 + c8 * 8);
                    const float4 bq1 = *reinterpret_cast<const float4*>(sb2 + c8 * 8 + 4);
                    const uint4 rv = rpre[u * 2 + e];
                    float v[8];
#define MLPC_SUM2(k, fld, ba, bb)                                                                                         \
                    {                                                                                                     \
                        const float2 a0 = unpack_f16x2(x0.fld), a1 = unpack_f16x2(x1.fld), a2 = unpack_f16x2(x2.fld);    \
                        const float2 rr = unpack_bf16x2(rv.fld);                                                          \
                        v[k] = (((__uint_as_float(r[e * 8 + k]) + a0.x) + (a1.x + a2.x)) + ba) + rr.x;                    \
                        v[k + 1] = (((__uint_as_float(r[e * 8 + k + 1]) + a0.y) + (a1.y + a2.y)) + bb) + rr.y;            \
                    }
                    MLPC_SUM2(0, x, bq0.x, bq0.y) MLPC_SUM2(2, y, bq0.z, bq0.w) MLPC_SUM2(4, z, bq1.x, bq1.y) MLPC_SUM2(6, w, bq1.z, bq1.w)
#undef MLPC_SUM2
                    uint4 o;
                    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
                    o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
                    if (row_ok) *reinterpret_cast<uint4*>(drow + col + e * 8) = o;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(a2_empty); mbar_arrive(tile_done); }
            if (warp == 2 && lane == 0) MLPC_TRACE(44);
        }
        // the bulk copies read THIS CTA's smem: stay resident until the peers have acknowledged the last tile
        if (warp == 2 && ti > 0) mbar_wait_cluster(ack_bar, (uint32_t)(ti - 1) & 1u);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
    if (threadIdx.x == 0) MLPC_TRACE(45);
}

}  // namespace fvhd
