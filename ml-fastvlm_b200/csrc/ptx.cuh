// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM).
// Hand-written for this library; descriptor bit layouts follow the PTX ISA "tcgen05 matrix
// descriptor" / "instruction descriptor" tables (K-major operands, SWIZZLE_128B).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fvhd {

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

constexpr int FVHD_MAX_PEERS = 8;
constexpr int FVHD_MAX_SCATTER = 64;

// Caller-owned pointers of one forward call, kept in device memory so that the launch sequence itself is
// static (replayable as a CUDA graph): written by set_io_kernel, read by the first and last kernels.
struct IoBlock {
    const void* images;     // NCHW input of the stem
    void* final_out;        // destination of the last unit's output when it is caller memory
    void* tokens_out;       // optional copy-out of the tower tokens when a projector follows
    long long final_image_stride;   // elements between consecutive images in final_out (N*H when dense; L*H when the
                                    // destination is a [B, L, H] LLM embedding buffer -- the token splice, llava_arch.py:251-271)
    // Fused all-gather: the projector epilogue ALSO stores every output vector at the same element offset in
    // `n_peers` other GPUs' gathered buffers (peer-mapped device pointers: NVLink P2P stores issued by the kernel that
    // runs the tcgen05 tiles), so no separate collective pass follows the projector.
    int n_peers;
    void* peer_out[FVHD_MAX_PEERS];
    // Scatter (multi-<image> / ragged splice, llava_arch.py:233-271): when scatter_n > 0 image b's token block goes to
    // scatter[b] (dense [N, H] rows) instead of final_out + b * final_image_stride.
    int scatter_n;
    void* scatter[FVHD_MAX_SCATTER];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------ programmatic dependent launch
// Every kernel of the forward is launched with programmaticStreamSerialization: it may start while its
// predecessor is still draining.  Before pdl_wait() a kernel may only touch its own smem/TMEM and the
// (constant) weights; everything produced by earlier kernels is read -- and any global write is issued --
// strictly after pdl_wait(), which returns once all prerequisite grids have completed and flushed.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load global -> shared, completion on an mbarrier (complete_tx::bytes).
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// 4-D tiled load (NHWC activation tile + halo: coords {channel, x, y, image}); out-of-bounds elements are zero-filled,
// which IS the convolution's zero padding.
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// Same load, multicast: the box lands at the same CTA-relative smem offset in every CTA of `cta_mask`, and each
// destination CTA's mbarrier (same offset) receives the complete_tx.
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
        : "memory");
}

// 2-D tiled store shared -> global (bulk async group); clips rows / columns outside the tensor.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------ tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 x bf16 -> fp32, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// One lane of a fully converged warp (warp-uniform predicate: lets the compiler keep descriptors in uniform registers and
// predicate the tcgen05 instruction instead of emitting a per-instruction uniformisation loop).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Same, arriving on the mbarrier at this offset in every CTA of `cta_mask` (cluster multicast pipelines).
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {      // every thread of every CTA in the cluster
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- distributed shared memory (cluster): address of `local_smem_addr` in CTA `cta`, remote arrive / store
__device__ __forceinline__ uint32_t cluster_map(uint32_t local_smem_addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar_addr) {    // release at cluster scope
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {  // acquire at cluster scope
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// Bulk copy local smem -> smem of another CTA of the cluster; the DESTINATION CTA's mbarrier receives complete_tx(bytes).
__device__ __forceinline__ void bulk_copy_to_cluster(uint32_t dst_cluster_addr, const void* src_local, uint32_t bytes, uint32_t dst_cluster_bar) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_cluster_addr), "r"(smem_u32(src_local)), "r"(bytes), "r"(dst_cluster_bar) : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// two fp32 -> packed f16x2 (lo in the low half), saturating to +-65504 instead of overflowing to inf
__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t v) {
    return __half22float2(*reinterpret_cast<const __half2*>(&v));
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (thread t <- lane base+t).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {      // same, 16 columns
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major operand, SWIZZLE_128B, rows of 128 bytes (64 bf16):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (=1, unused for swizzled K-major) |
//   [32,46) SBO >> 4 (= 8 rows * 128 B = 1024 B) | [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor, kind::f16: D=f32 (bit4), A=B=bf16 (bits 7,10), both K-major,
// N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(uint32_t M, uint32_t N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ------------------------------------------------------------------ misc
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    bf162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
    float2 f;
    f.x = __uint_as_float(u << 16);
    f.y = __uint_as_float(u & 0xFFFF0000u);
    return f;
}

// Exact-erf GELU (nn.GELU() default, mci.py:108,387,870): 0.5 x (1 + erf(x / sqrt 2)).
// Reference-accuracy version: erf by Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7), one ex2 + one rcp on the SFU.
__device__ __forceinline__ float gelu_erf_as(float x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    p *= t;
    const float e = p * exp2f(-1.4426950408889634f * z * z);   // 1 - erf(z)
    const float half_erfc = 0.5f * e;
    return x >= 0.f ? x * (1.0f - half_erfc) : x * half_erfc;
}
// Epilogue version: erf(x/sqrt2) == tanh(atanh(erf(x/sqrt2))) and atanh(erf(x/sqrt2)) = x * P(x^2) with a
// 3-term least-squares P (|GELU error| <= 3.0e-5 over all x, 16x tighter than the textbook tanh-GELU and
// two orders below bf16 output rounding).  One MUFU (tanh.approx, rel. error 2^-11) + 6 FMA-pipe ops: the
// 2-MUFU form above costs more SFU cycles per 128x128 tile than the tensor core needs for K <= 384.
__device__ __forceinline__ float gelu_erf(float x) {
    // P peaks at x^2 = 51.6 and would change sign near |x| = 11: clamp there -- x*P(50) = 1.75x >= 12 for |x| >= 7.07,
    // where tanh is exactly +-1 in fp32, so the clamp is invisible in the result.
    const float x2 = fminf(x * x, 50.0f);
    float p = fmaf(-3.5873236112e-04f, x2, 3.7050345100e-02f);
    p = fmaf(p, x2, 7.9745847078e-01f);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * p));
    const float hx = 0.5f * x;
    // far negative tail: 1 + tanh cancels to the MUFU's 2^-11 relative error; the true value is |GELU(x)| < 1e-6 there
    return x < -5.5f ? 0.0f : fmaf(hx, t, hx);
}

// (A packed f16x2 variant -- HFMA2 pipe, one tanh.approx.f16x2 per two elements -- was measured on B200: parity fine,
// but no faster than this f32 form once the f32<->f16 conversions are counted; not kept.)

}  // namespace fvhd
