// D[M,N] = epilogue( A[M,K] . W[N,K]^T ) on the 5th-gen tensor cores.
//
// Every pointwise (1x1) conv, Linear and projector layer of the FastViTHD path is this kernel:
// NHWC activations ARE the row-major [pixels, C] A operand, so no im2col / transpose exists.
//   ConvFFN.fc1/fc2 (mci.py:922-926), MobileOneBlock 1x1 (mci.py:591-602, 727-737),
//   MHSA.qkv / MHSA.proj (mci.py:669-681), mm_projector (multimodal_projector/builder.py:23-30).
//
// Persistent, warp-specialised (320 threads, one CTA per SM, CTA c walks tiles c, c+G, ...):
//   warp 0      : TMA producer  -- cp.async.bulk.tensor 2D, 128-B swizzled 128x64 (A) / BNx64 (W) bf16 boxes
//                 into a `stages`-deep smem ring (full/empty mbarriers).  With a thread-block cluster of CS CTAs the
//                 operand the CTAs have in common (A when the cluster runs along N, W when it runs along M) is
//                 fetched ONCE per cluster: every CTA loads 1/CS of the box and multicasts it to all peers, which
//                 divides the L2->SM traffic of that operand by CS (these GEMMs are L2-ingest bound at batch 1).
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (M=128, N=BN, K=16 per instruction);
//                 tcgen05.commit frees ring slots and publishes finished accumulators
//   warps 2..9  : epilogue -- two accumulator buffers in TMEM, so the epilogue of tile i overlaps the MMAs of
//                 tile i+1.  A warp may only touch TMEM lanes 32*(warp%4)..+31, so a warp owns 32 rows and walks
//                 64-column groups: tcgen05.ld (next 32-column chunk in flight while the current one is
//                 processed) -> +bias -> GELU -> +residual -> bf16 -> 128-B-swizzled smem staging (32 rows x 128 B)
//                 -> one TMA bulk store per group (full 128-B lines, M/N tails clipped by the tensor map).
#pragma once
#include "ptx.cuh"

namespace fvhd {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_THREADS = 32 * (2 + GEMM_EPI_WARPS);
constexpr int GEMM_MAX_STAGES = 8;
constexpr int GEMM_A_STAGE_BYTES = GEMM_BM * GEMM_BK * 2;   // 16 KiB
constexpr int GEMM_EPI_STAGE_BYTES = 32 * 128;                  // per epilogue warp: 32 rows x 64 bf16, SWIZZLE_128B
constexpr int GEMM_EPI_BIAS_BYTES = 2 * 64 * 4;                 // per epilogue warp: bias of its (up to) two 64-column groups
constexpr int kSplitKDoneOfs = 2048;                              // counters[tile] = arrivals, counters[2048 + tile] = finished reducers
constexpr int GEMM_SMEM_BUDGET = 200 * 1024 - GEMM_EPI_WARPS * (GEMM_EPI_STAGE_BYTES + GEMM_EPI_BIAS_BYTES);

struct GemmParams {
    int M, N, K;
    int BN;            // N tile: multiple of 32, 32..256
    int stages;        // 1..GEMM_MAX_STAGES
    int tiles_m, tiles_n;
    bf16* D;           // [M, ldd] bf16; nullptr => io->final_out (caller memory)
    const IoBlock* io;
    int rows_per_image;      // D == nullptr only: rows (tokens) per image, images io->final_image_stride elements apart
    int ldd;
    const float* bias;       // [N] fp32 or nullptr
    const bf16* residual;    // [M, ldr] bf16 or nullptr (added after activation)
    int ldr;
    int act;                 // 0 = identity, 1 = exact-erf GELU
    int cs;                  // cluster size 1 / 2 / 4
    int share_b;             // cluster runs along M and shares the W box (else along N, sharing the A box)
    int ctiles;              // cluster-tiles = work items of one cluster
    unsigned long long* trace;   // debug: 16 globaltimer stamps per CTA (nullptr = off)
    int tma_store;           // 1: tmD is valid (D is library memory) -> full 64-column groups leave through TMA
    // Split-K (small-M, weight-streaming GEMMs of stages 3-4 / projector at small batch: a 2 x 12-tile grid cannot keep 148 SMs
    // streaming weights).  Work item = (tile, k-slice), ONE per CTA (all co-resident): every CTA accumulates its slice, writes the
    // fp32 partial tile to `ws`, waits for the tile's other slices, then reduces its own band of 128 / split_k rows (slice order:
    // deterministic) with the bias / GELU / residual epilogue.  Counters are self-resetting.  cs must be 1, D must be non-null.
    int split_k;             // 1 = off
    int kb_per_split;        // k-blocks per slice
    float* ws;               // [tiles][split_k][128][BN] fp32
    int* counters;           // [tiles], zero before the first launch
};

__host__ __device__ inline int gemm_acc_stride(int bn) {     // TMEM columns per accumulator buffer (power of 2)
    int c = 32;
    while (c < bn) c <<= 1;
    return c;
}
__host__ inline size_t gemm_smem_bytes(int bn, int stages) {
    return (size_t)stages * (GEMM_A_STAGE_BYTES + (size_t)bn * GEMM_BK * 2) + GEMM_EPI_WARPS * (GEMM_EPI_STAGE_BYTES + GEMM_EPI_BIAS_BYTES) +
           1024 /*align slack*/ + 256 /*barriers*/;
}
__host__ inline int gemm_pick_stages(int bn, int num_kb) {
    int s = GEMM_SMEM_BUDGET / (GEMM_A_STAGE_BYTES + bn * GEMM_BK * 2);
    if (s > GEMM_MAX_STAGES) s = GEMM_MAX_STAGES;
    (void)num_kb;
    return s < 1 ? 1 : s;
}

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define GEMM_TRACE(i) do { if (p.trace) p.trace[(size_t)blockIdx.x * 16 + (i)] = gtime(); } while (0)

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmD, const GemmParams p) {
    extern __shared__ uint8_t gemm_smem_raw[];
    const uint32_t raw_addr = smem_u32(gemm_smem_raw);
    const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
    uint8_t* smem = gemm_smem_raw + pad;                       // 1024-B aligned (SWIZZLE_128B atoms)

    const int stages = p.stages;
    const int BN = p.BN;
    const uint32_t b_stage_bytes = (uint32_t)BN * GEMM_BK * 2;
    uint8_t* smemA = smem;
    uint8_t* smemB = smem + (size_t)stages * GEMM_A_STAGE_BYTES;
    uint8_t* smemE = smemB + (size_t)stages * b_stage_bytes;    // epilogue staging, 8 x 4 KiB (1024-B aligned)
    float* smemBias = reinterpret_cast<float*>(smemE + GEMM_EPI_WARPS * GEMM_EPI_STAGE_BYTES);   // 8 x 128 floats
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smemE + GEMM_EPI_WARPS * (GEMM_EPI_STAGE_BYTES + GEMM_EPI_BIAS_BYTES));
    uint64_t* empty_bar = full_bar + GEMM_MAX_STAGES;
    uint64_t* tfull_bar = empty_bar + GEMM_MAX_STAGES;          // [2] accumulator ready
    uint64_t* tempty_bar = tfull_bar + 2;                       // [2] accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    pdl_launch_dependents();
    if (threadIdx.x == 0) GEMM_TRACE(0);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
    // work decomposition: cluster `cid` of `nclusters` walks cluster-tiles cid, cid + nclusters, ...; the CTA of
    // rank r inside it takes the r-th tile of the cluster-tile along the clustered dimension.
    const int CS = p.cs;
    const int rank = CS > 1 ? (int)cluster_ctarank() : 0;
    const int cid = (int)blockIdx.x / CS;
    const int nclusters = (int)gridDim.x / CS;
    const int gdim_n = p.share_b ? p.tiles_n : (p.tiles_n + CS - 1) / CS;     // cluster-tile grid width
    const uint16_t mc_mask = (uint16_t)((1u << CS) - 1u);
    auto tile_origin = [&](int ct, int& m0, int& n0) {
        const int gm = ct / gdim_n, gn = ct % gdim_n;
        m0 = (p.share_b ? gm * CS + rank : gm) * GEMM_BM;
        n0 = (p.share_b ? gn : gn * CS + rank) * BN;
    };
    const uint32_t acc_stride = gemm_acc_stride(BN);
    const uint32_t tmem_cols = 2 * acc_stride;
    // split-K: work item ct = tile * S + slice; slice ks covers k-blocks [ks * kbs, min(.., num_kb))
    const int S = p.split_k > 1 ? p.split_k : 1;
    const int kbs = S > 1 ? p.kb_per_split : num_kb;
    auto item_kb = [&](int ct, int& tile, int& kb0, int& kb1) {
        tile = ct / S;
        kb0 = (ct - tile * S) * kbs;
        kb1 = kb0 + kbs < num_kb ? kb0 + kbs : num_kb;
    };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (p.tma_store) tma_prefetch_desc(&tmD);
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CS);               // every consumer of the cluster releases the slot
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
    if (CS > 1) cluster_sync_all(); else __syncthreads();      // peers' barriers must exist before any multicast lands
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) GEMM_TRACE(1);

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer
            const uint32_t a_slice = GEMM_A_STAGE_BYTES / (uint32_t)CS, b_slice = b_stage_bytes / (uint32_t)CS;
            // W (the B operand) is a constant of the forward: the first ring fill of weight boxes is issued BEFORE the PDL wait,
            // so it overlaps the predecessor's tail; only the A boxes (the predecessor's output) wait for it.
            int pre = 0;
            if (CS == 1) {
                for (int ct = cid; ct < p.ctiles && pre < stages; ct += nclusters) {
                    int m0, n0, tile, kb0, kb1;
                    item_kb(ct, tile, kb0, kb1);
                    tile_origin(tile, m0, n0);
                    for (int kb = kb0; kb < kb1 && pre < stages; ++kb, ++pre) {
                        mbar_expect_tx(&full_bar[pre], GEMM_A_STAGE_BYTES + b_stage_bytes);
                        tma_load_2d(smemB + (size_t)pre * b_stage_bytes, &tmB, kb * GEMM_BK, n0, &full_bar[pre]);
                    }
                }
            }
            pdl_wait();                                     // A is the previous kernel's output
            GEMM_TRACE(2);
            int it = 0;
            for (int ct = cid; ct < p.ctiles; ct += nclusters) {
                int m0, n0, tile, kb0, kb1;
                item_kb(ct, tile, kb0, kb1);
                tile_origin(tile, m0, n0);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % stages;
                    const uint32_t ph = (uint32_t)(it / stages) & 1u;
                    uint8_t* sa = smemA + (size_t)s * GEMM_A_STAGE_BYTES;
                    uint8_t* sb = smemB + (size_t)s * b_stage_bytes;
                    if (it < pre) {                          // slot armed and its W box already in flight
                        tma_load_2d(sa, &tmA, kb * GEMM_BK, m0, &full_bar[s]);
                        continue;
                    }
                    mbar_wait(&empty_bar[s], ph ^ 1u);       // slot s is free in EVERY CTA of the cluster
                    mbar_expect_tx(&full_bar[s], GEMM_A_STAGE_BYTES + b_stage_bytes);
                    if (CS == 1) {
                        tma_load_2d(sa, &tmA, kb * GEMM_BK, m0, &full_bar[s]);
                        tma_load_2d(sb, &tmB, kb * GEMM_BK, n0, &full_bar[s]);
                    } else if (p.share_b) {                  // private A box; my 1/CS of the shared W box -> all peers
                        tma_load_2d(sa, &tmA, kb * GEMM_BK, m0, &full_bar[s]);
                        tma_load_2d_mc(sb + (size_t)rank * b_slice, &tmB, kb * GEMM_BK, n0 + rank * (BN / CS), &full_bar[s], mc_mask);
                    } else {                                 // my 1/CS of the shared A box -> all peers; private W box
                        tma_load_2d_mc(sa + (size_t)rank * a_slice, &tmA, kb * GEMM_BK, m0 + rank * (GEMM_BM / CS), &full_bar[s], mc_mask);
                        tma_load_2d(sb, &tmB, kb * GEMM_BK, n0, &full_bar[s]);
                    }
                }
                if (ct == cid) GEMM_TRACE(3);               // first tile fully issued
            }
            GEMM_TRACE(4);                                  // all loads issued
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------- MMA issuer (one thread drives the tensor core for the whole CTA)
            const uint32_t idesc = umma_idesc_bf16(GEMM_BM, (uint32_t)BN);
            int it = 0, tl = 0;
            for (int ct = cid; ct < p.ctiles; ct += nclusters, ++tl) {
                const int a = tl & 1;
                const uint32_t aph = (uint32_t)(tl >> 1) & 1u;
                mbar_wait(&tempty_bar[a], aph ^ 1u);            // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)a * acc_stride;
                int tile_, kb0, kb1;
                item_kb(ct, tile_, kb0, kb1);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % stages;
                    const uint32_t ph = (uint32_t)(it / stages) & 1u;
                    mbar_wait(&full_bar[s], ph);
                    if (it == 0) GEMM_TRACE(5);             // first k-block landed
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128(smem_u32(smemA + (size_t)s * GEMM_A_STAGE_BYTES));
                    const uint64_t db = umma_desc_sw128(smem_u32(smemB + (size_t)s * b_stage_bytes));
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        // advance 16 bf16 = 32 B along K inside the 128-B swizzle row: +2 in (addr >> 4) units
                        umma_bf16(acc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
                    }
                    if (CS == 1) umma_commit(&empty_bar[s]);          // slot reusable once these MMAs have read it
                    else umma_commit_mc(&empty_bar[s], mc_mask);      // ... told to every producer that writes into it
                }
                umma_commit(&tfull_bar[a]);          // accumulator complete
                if (ct == cid) GEMM_TRACE(6);        // first tile's MMAs issued
            }
            GEMM_TRACE(7);                           // all MMAs issued
        }
    } else {
        // ---------------- epilogue: warp w owns TMEM lanes [32*(w%4), +32) == 32 tile rows; the two warps of a
        // lane quarter (w and w+4) alternate over the 64-column groups of the tile.
        const int q = warp & 3;
        const int hh = (warp - 2) >> 2;
        const int ngroups = (BN + 63) / 64;
        uint8_t* stage = smemE + (size_t)(warp - 2) * GEMM_EPI_STAGE_BYTES;
        float* bstage = smemBias + (warp - 2) * 128;
        const uint32_t sw = (uint32_t)(lane & 7);               // 128-B swizzle: 16-B chunk j of row r lives at chunk j ^ (r & 7)
        uint8_t* srow = stage + lane * 128;
        bool store_pending = false;
        int tl = 0;
        pdl_wait();                                         // residual / io reads and every global write come after this
        int* sflag = reinterpret_cast<int*>(tmem_slot + 2);   // split-K: "this CTA is the last slice of its tile"
        for (int ct = cid; ct < p.ctiles; ct += nclusters, ++tl) {
            const int a = tl & 1;
            const uint32_t aph = (uint32_t)(tl >> 1) & 1u;
            int m0, n0, tile, kb0_, kb1_;
            item_kb(ct, tile, kb0_, kb1_);
            tile_origin(tile, m0, n0);
            const int row = m0 + q * 32 + lane;
            const bool row_ok = row < p.M;
            bf16* drow;
            size_t doff = 0;        // element offset of this row in the caller's buffer (and in every peer's gathered buffer)
            int npeers = 0;
            if (p.D) {
                drow = p.D + (size_t)row * p.ldd;
            } else {        // caller memory: image b's tokens start at final_out + b * final_image_stride
                const int bi = row / p.rows_per_image;
                if (p.io->scatter_n > 0) {     // every image has its own destination block (ragged / multi-<image> splice)
                    drow = reinterpret_cast<bf16*>(p.io->scatter[bi < p.io->scatter_n ? bi : 0]) + (size_t)(row - bi * p.rows_per_image) * p.ldd;
                } else {
                    doff = (size_t)bi * (size_t)p.io->final_image_stride + (size_t)(row - bi * p.rows_per_image) * p.ldd;
                    drow = reinterpret_cast<bf16*>(p.io->final_out) + doff;
                    npeers = p.io->n_peers;
                }
            }
            const bf16* rrow = p.residual ? p.residual + (size_t)row * p.ldr : nullptr;
            // ---- everything that does not depend on the accumulator is fetched BEFORE waiting for it: the bias of this
            // warp's column groups (coalesced, staged in smem for broadcast reads) and the residual row segment of the
            // first group (registers).  Their global-load latency used to sit on the critical path of every chunk.
            if (p.bias) {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int g = hh + 2 * k;
                    if (g < ngroups) {
                        const int col = n0 + g * 64 + 2 * lane;
                        float2 bv = make_float2(0.f, 0.f);
                        if (col + 2 <= p.N) bv = __ldg(reinterpret_cast<const float2*>(p.bias + col));
                        *reinterpret_cast<float2*>(bstage + k * 64 + 2 * lane) = bv;
                    }
                }
            }
            uint4 rpre[8];
            if (rrow) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int col = n0 + hh * 64 + j * 8;
                    rpre[j] = (row_ok && hh < ngroups && col + 8 <= p.N) ? *reinterpret_cast<const uint4*>(rrow + col) : make_uint4(0, 0, 0, 0);
                }
            }
            __syncwarp();
            mbar_wait(&tfull_bar[a], aph);
            if (warp == 2 && lane == 0 && tl == 0) GEMM_TRACE(8);    // first accumulator ready
            tc_fence_after();
            const uint32_t lane_addr = tmem_base + (uint32_t)a * acc_stride + ((uint32_t)(q * 32) << 16);
            bool do_epi = true;
            if (S > 1) {
                // ---- split-K.  (1) publish this slice's fp32 partial tile; (2) wait until all S slices of the tile have published
                // (the S CTAs are co-resident: the host launches exactly one work item per CTA); (3) every slice CTA reduces a band of
                // 128 / S rows of the tile -- S coalesced float4 streams, summed in slice order (deterministic) -- applies bias / GELU /
                // residual and stores bf16.  The reduction is spread over all CTAs instead of serialising on the last arrival.
                do_epi = false;
                const int ks = ct - tile * S;
                float* wsp = p.ws + ((size_t)ct * GEMM_BM + (size_t)(q * 32 + lane)) * (size_t)BN;
                for (int g = hh; g < ngroups; g += 2) {
                    const int gcol = g * 64;
                    const int gw = (BN - gcol) < 64 ? (BN - gcol) : 64;
                    for (int c = 0; c * 32 < gw; ++c) {
                        uint32_t r[32];
                        tmem_ld32(lane_addr + (uint32_t)(gcol + c * 32), r);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            *reinterpret_cast<uint4*>(wsp + gcol + c * 32 + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                    }
                }
                __threadfence();
                named_bar_sync(1, GEMM_EPI_WARPS * 32);
                if (warp == 2 && lane == 0) {
                    atomicAdd(p.counters + tile, 1);
                    while (*reinterpret_cast<volatile int*>(p.counters + tile) < S) { }
                    __threadfence();
                }
                named_bar_sync(1, GEMM_EPI_WARPS * 32);
                const int band = (GEMM_BM + S - 1) / S;
                const int r_lo = ks * band, r_hi = (r_lo + band) < GEMM_BM ? (r_lo + band) : GEMM_BM;
                const int c4n = BN >> 2;                                  // float4 per tile row
                const int et = (int)threadIdx.x - 64;                     // 0..255 over the epilogue warps
                const float* wt = p.ws + (size_t)tile * S * GEMM_BM * BN;
                for (int idx = et; idx < (r_hi - r_lo) * c4n; idx += GEMM_EPI_WARPS * 32) {
                    const int rr = r_lo + idx / c4n, c4 = idx - (idx / c4n) * c4n;
                    const int grow = m0 + rr, gcol = n0 + 4 * c4;
                    float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int sl = 0; sl < S; ++sl) {
                        const float4 v4 = __ldcg(reinterpret_cast<const float4*>(wt + ((size_t)sl * GEMM_BM + rr) * BN) + c4);
                        acc4.x += v4.x; acc4.y += v4.y; acc4.z += v4.z; acc4.w += v4.w;
                    }
                    if (grow < p.M && gcol + 4 <= p.N) {
                        if (p.bias) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + gcol));
                            acc4.x += b4.x; acc4.y += b4.y; acc4.z += b4.z; acc4.w += b4.w;
                        }
                        if (p.act == 1) { acc4.x = gelu_erf(acc4.x); acc4.y = gelu_erf(acc4.y); acc4.z = gelu_erf(acc4.z); acc4.w = gelu_erf(acc4.w); }
                        if (p.residual) {
                            const uint2 rv = *reinterpret_cast<const uint2*>(p.residual + (size_t)grow * p.ldr + gcol);
                            const float2 r0 = unpack_bf16x2(rv.x), r1 = unpack_bf16x2(rv.y);
                            acc4.x += r0.x; acc4.y += r0.y; acc4.z += r1.x; acc4.w += r1.y;
                        }
                        *reinterpret_cast<uint2*>(p.D + (size_t)grow * p.ldd + gcol) = make_uint2(pack_bf16x2(acc4.x, acc4.y), pack_bf16x2(acc4.z, acc4.w));
                    }
                }
                // the last slice to finish re-arms the tile's counters for the next launch
                named_bar_sync(1, GEMM_EPI_WARPS * 32);
                if (warp == 2 && lane == 0) {
                    const int done = atomicAdd(p.counters + kSplitKDoneOfs + tile, 1);
                    if (done == S - 1) { p.counters[tile] = 0; p.counters[kSplitKDoneOfs + tile] = 0; }
                }
            }
            auto load_chunk = [&](int colofs, uint32_t (&dst)[32]) { tmem_ld32(lane_addr + (uint32_t)colofs, dst); };
            int gk = 0;
            for (int g = hh; do_epi && g < ngroups; g += 2, ++gk) {
                const int gcol = g * 64;                         // first tile column of the group
                const int gw = (BN - gcol) < 64 ? (BN - gcol) : 64;   // 64, or 32 for the tail group of BN = 32 / 96
                const bool via_tma = p.tma_store && gw == 64;
                const float* bgrp = bstage + (gk & 1) * 64;
                uint32_t r[2][32];
                load_chunk(gcol, r[0]);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (c * 32 >= gw) break;
                    tmem_ld_wait();
                    if (c == 0 && gw == 64) load_chunk(gcol + 32, r[1]);   // next chunk in flight
                    const int col0 = n0 + gcol + c * 32;
#pragma unroll
                    for (int g8 = 0; g8 < 4; ++g8) {
                        const int col = col0 + g8 * 8;
                        const bool col_ok = col + 8 <= p.N;      // N % 8 == 0 is enforced on the host
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[c][g8 * 8 + j]);
                        if (p.bias) {
                            const float4 b0 = *reinterpret_cast<const float4*>(bgrp + c * 32 + g8 * 8);
                            const float4 b1 = *reinterpret_cast<const float4*>(bgrp + c * 32 + g8 * 8 + 4);
                            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                            v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
                        }
                        if (p.act == 1) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = gelu_erf(v[j]);
                        }
                        if (rrow) {
                            uint4 rv;
                            if (gk == 0) rv = rpre[c * 4 + g8];                         // prefetched before the accumulator wait
                            else rv = (row_ok && col_ok) ? *reinterpret_cast<const uint4*>(rrow + col) : make_uint4(0, 0, 0, 0);
                            const float2 r0 = unpack_bf16x2(rv.x), r1 = unpack_bf16x2(rv.y), r2 = unpack_bf16x2(rv.z), r3 = unpack_bf16x2(rv.w);
                            v[0] += r0.x; v[1] += r0.y; v[2] += r1.x; v[3] += r1.y;
                            v[4] += r2.x; v[5] += r2.y; v[6] += r3.x; v[7] += r3.y;
                        }
                        uint4 o;
                        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
                        o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
                        if (via_tma) {
                            if (c == 0 && g8 == 0) {             // first smem write of the group: previous TMA store must have read the staging
                                if (store_pending) { tma_store_wait_read0(); store_pending = false; }
                                __syncwarp();
                            }
                            *reinterpret_cast<uint4*>(srow + ((((uint32_t)(c * 4 + g8)) ^ sw) << 4)) = o;
                        } else if (row_ok && col_ok) {
                            *reinterpret_cast<uint4*>(drow + col) = o;
                            // fused all-gather: the same 16 bytes go to this rank's slot in every peer GPU's gathered buffer
                            for (int q = 0; q < npeers; ++q)
                                *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.io->peer_out[q]) + doff + col) = o;
                        }
                    }
                }
                if (via_tma) {
                    fence_proxy_async_smem();                    // generic-proxy smem writes -> visible to the TMA engine
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmD, stage, n0 + gcol, m0 + q * 32);
                        tma_store_commit();
                    }
                    store_pending = true;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[a]);     // this warp is done reading accumulator a
            if (warp == 2 && lane == 0 && tl == 0) GEMM_TRACE(9);    // first tile's epilogue done
        }
        if (warp == 2 && lane == 0) GEMM_TRACE(10);         // all epilogues done
        if (store_pending) tma_store_wait_all();            // global writes complete before the CTA retires
        if (warp == 2 && lane == 0) GEMM_TRACE(11);         // stores drained
    }

    tc_fence_before();
    if (CS > 1) cluster_sync_all(); else __syncthreads();      // no CTA may retire while peers can still write into it
    if (threadIdx.x == 0) GEMM_TRACE(12);
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

}  // namespace fvhd
