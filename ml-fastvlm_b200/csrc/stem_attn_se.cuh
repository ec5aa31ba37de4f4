// Stem, channel LayerNorm, MHSA core and conv_exp squeeze-excite kernels (NHWC bf16 activations).
#pragma once
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace fvhd {

// ====================================================================== stem
// convolutional_stem blocks 0 and 1 fused (mci.py:567-590): NCHW image (fp32/fp16/bf16) ->
// conv3x3 s2 (3->96) + GELU -> dw3x3 s2 + GELU -> NHWC bf16 [B, R/4, R/4, 96].
// The R/2 x R/2 x 96 intermediate (50 MB / image at R=1024) never reaches HBM: a CTA computes the 17x17 conv0 patch its
// 8x8 output tile needs into shared memory.  conv0 is an implicit GEMM on the tensor cores (mma.sync m16n8k16, f16 inputs --
// 11-bit significands hold both the k/255 pixel grid and the weights more precisely than bf16 -- fp32 accumulate):
// M = 289 patch pixels (19 m-tiles), N = 96, K = 27 taps padded to 32; A fragments are gathered directly from the staged
// fp16 input patch (no im2col buffer), the 48 B-fragment registers hold all of w0 for the CTA's lifetime.
// Block 2 (1x1 + GELU) is a tcgen05 GEMM launch.
constexpr int STEM_C = 96;
constexpr int STEM_TO = 8;                       // output tile (at R/4)
constexpr int STEM_MID = 2 * STEM_TO + 1;        // 17: conv0 patch edge
constexpr int STEM_IN = 2 * STEM_MID + 1;        // 35: input patch edge
constexpr int STEM_INP = STEM_IN + 1;            // 36: padded pitch
constexpr int STEM_MIDP = STEM_C / 2 + 1;        // 49 words per conv0 pixel (bf16 pairs, +1 pad)
constexpr int STEM_THREADS = 256;
constexpr int STEM_NPIX = STEM_MID * STEM_MID;   // 289
constexpr int STEM_MT = (STEM_NPIX + 15) / 16;   // 19 m-tiles
constexpr size_t STEM_SMEM = (size_t)32 * STEM_C * 2 /*w0 fp16 [32][96]*/ + (size_t)(9 * STEM_C + 2 * STEM_C) * 4 /*w1, b0, b1*/ +
                             (size_t)3 * STEM_IN * STEM_INP * 2 /*input patch fp16*/ + 16 + (size_t)STEM_NPIX * STEM_MIDP * 4 /*conv0 out*/;

template <typename T> __device__ __forceinline__ float img_ld(const T* p);
template <> __device__ __forceinline__ float img_ld<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float img_ld<__half>(const __half* p) { return __half2float(*p); }
template <> __device__ __forceinline__ float img_ld<bf16>(const bf16* p) { return __bfloat162float(*p); }

__device__ __forceinline__ void mma_f16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <typename T>
__global__ void __launch_bounds__(STEM_THREADS)
stem_kernel(const IoBlock* __restrict__ io, bf16* __restrict__ out, const float* __restrict__ w0 /*[27][96], k=(ci*3+ky)*3+kx*/,
            const float* __restrict__ b0, const float* __restrict__ w1 /*[9][96]*/, const float* __restrict__ b1, int R, int tiles_x) {
    extern __shared__ __align__(16) uint8_t stem_smem_raw[];
    __half* w0h = reinterpret_cast<__half*>(stem_smem_raw);                       // [32][96], rows 27..31 zero
    float* w1s = reinterpret_cast<float*>(w0h + 32 * STEM_C);                     // [9][96]
    float* b0s = w1s + 9 * STEM_C;
    float* b1s = b0s + STEM_C;
    __half* sin = reinterpret_cast<__half*>(b1s + STEM_C);                        // [3][35][36]
    uint32_t* s1 = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(sin) + ((3 * STEM_IN * STEM_INP * 2 + 15) / 16) * 16);   // [289][49]

    pdl_launch_dependents();
    const int b = blockIdx.z;
    const int ty0 = (blockIdx.x / tiles_x) * STEM_TO, tx0 = (blockIdx.x % tiles_x) * STEM_TO;
    const int R2 = R / 2, R4 = R / 4;
    const int iy0 = 4 * ty0 - 3, ix0 = 4 * tx0 - 3;     // input patch origin
    const int cy0 = 2 * ty0 - 1, cx0 = 2 * tx0 - 1;     // conv0 patch origin

    for (int i = threadIdx.x; i < 32 * STEM_C; i += STEM_THREADS) w0h[i] = __float2half_rn(i < 27 * STEM_C ? __ldg(w0 + i) : 0.f);
    for (int i = threadIdx.x; i < 9 * STEM_C; i += STEM_THREADS) w1s[i] = __ldg(w1 + i);
    if (threadIdx.x < STEM_C) { b0s[threadIdx.x] = __ldg(b0 + threadIdx.x); b1s[threadIdx.x] = __ldg(b1 + threadIdx.x); }
    pdl_wait();                                                          // io block is written by set_io_kernel
    const T* __restrict__ img = reinterpret_cast<const T*>(io->images);
    for (int i = threadIdx.x; i < 3 * STEM_IN * STEM_IN; i += STEM_THREADS) {
        const int ci = i / (STEM_IN * STEM_IN);
        const int rem = i - ci * STEM_IN * STEM_IN;
        const int yy = rem / STEM_IN, xx = rem - yy * STEM_IN;
        const int gy = iy0 + yy, gx = ix0 + xx;
        float v = 0.f;
        if (gy >= 0 && gy < R && gx >= 0 && gx < R) v = img_ld<T>(img + (((size_t)b * 3 + ci) * R + gy) * R + gx);
        sin[(ci * STEM_IN + yy) * STEM_INP + xx] = __float2half_rn(v);
    }
    __syncthreads();

    // phase 1: conv0 3x3 s2 + GELU on the 17x17 patch as an implicit GEMM on the tensor cores
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int g = lane >> 2, t = lane & 3;
        const uint16_t* w0u = reinterpret_cast<const uint16_t*>(w0h);
        const uint16_t* sinu = reinterpret_cast<const uint16_t*>(sin);
        // B fragments of all 12 n-tiles x 2 k-steps: {w0[k][n], w0[k+1][n]}, {w0[k+8][n], w0[k+9][n]}, n = 8 nt + g
        uint32_t bf[12][2][2];
#pragma unroll
        for (int nt = 0; nt < 12; ++nt)
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int k = 16 * s + 2 * t + 8 * h2, n = nt * 8 + g;
                    bf[nt][s][h2] = (uint32_t)w0u[k * STEM_C + n] | ((uint32_t)w0u[(k + 1) * STEM_C + n] << 16);
                }
        // patch offsets of this thread's 8 K indices: k -> (ci, ky, kx) -> (ci*35 + ky)*36 + kx  (k >= 27: weight is zero)
        int koff[2][2][2];
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = 16 * s + 2 * t + 8 * h2 + e;
                    const int kk = k < 27 ? k : 0;
                    const int ci = kk / 9, r9 = kk - ci * 9, ky = r9 / 3, kx = r9 - ky * 3;
                    koff[s][h2][e] = (ci * STEM_IN + ky) * STEM_INP + kx;
                }
        for (int mt = warp; mt < STEM_MT; mt += STEM_THREADS / 32) {
            int pix[2], base[2];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                int pp = mt * 16 + g + 8 * rr;
                pix[rr] = pp;
                if (pp > STEM_NPIX - 1) pp = STEM_NPIX - 1;          // padded rows: computed, never stored
                const int py = pp / STEM_MID, px = pp - py * STEM_MID;
                base[rr] = (2 * py) * STEM_INP + 2 * px;
            }
            uint32_t af[2][4];
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr)
                        af[s][h2 * 2 + rr] = (uint32_t)sinu[base[rr] + koff[s][h2][0]] | ((uint32_t)sinu[base[rr] + koff[s][h2][1]] << 16);
            float d[12][4];
#pragma unroll
            for (int nt = 0; nt < 12; ++nt) {
                d[nt][0] = d[nt][1] = d[nt][2] = d[nt][3] = 0.f;
                mma_f16_16816(d[nt], af[0], bf[nt][0][0], bf[nt][0][1]);
                mma_f16_16816(d[nt], af[1], bf[nt][1][0], bf[nt][1][1]);
            }
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int pp = pix[rr];
                if (pp >= STEM_NPIX) continue;
                const int py = pp / STEM_MID, px = pp - py * STEM_MID;
                const int cy = cy0 + py, cx = cx0 + px;
                const bool inside = cy >= 0 && cy < R2 && cx >= 0 && cx < R2;      // outside: zero padding of the depthwise conv
                uint32_t* dst = s1 + pp * STEM_MIDP + t;
#pragma unroll
                for (int nt = 0; nt < 12; ++nt) {
                    const int ch = nt * 8 + 2 * t;
                    const float v0 = gelu_erf(d[nt][2 * rr + 0] + b0s[ch]);
                    const float v1 = gelu_erf(d[nt][2 * rr + 1] + b0s[ch + 1]);
                    dst[nt * 4] = inside ? pack_bf16x2(v0, v1) : 0u;
                }
            }
        }
    }
    __syncthreads();

    // phase 2: depthwise 3x3 s2 + GELU; item = (output pixel, channel pair)
    for (int it = threadIdx.x; it < STEM_TO * STEM_TO * (STEM_C / 2); it += STEM_THREADS) {
        const int op = it / (STEM_C / 2);
        const int cp = it - op * (STEM_C / 2);
        const int oy = op / STEM_TO, ox = op - oy * STEM_TO;
        float a0 = b1s[2 * cp], a1 = b1s[2 * cp + 1];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float2 v = unpack_bf16x2(s1[((2 * oy + ky) * STEM_MID + 2 * ox + kx) * STEM_MIDP + cp]);
                const float2 w = *reinterpret_cast<const float2*>(w1s + (ky * 3 + kx) * STEM_C + 2 * cp);
                a0 = fmaf(v.x, w.x, a0);
                a1 = fmaf(v.y, w.y, a1);
            }
        const int gy = ty0 + oy, gx = tx0 + ox;
        if (gy < R4 && gx < R4)
            *reinterpret_cast<uint32_t*>(out + (((size_t)b * R4 + gy) * R4 + gx) * STEM_C + 2 * cp) = pack_bf16x2(gelu_erf(a0), gelu_erf(a1));
    }
}

// ====================================================================== LayerNormChannel
// mci.py:617-623: per-pixel LayerNorm over channels (eps 1e-5), fp32 statistics, two-pass in registers.
// One warp per pixel row of the [M, C] activation matrix.  NV = C / 256.
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_channel_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, const float* __restrict__ gamma,
                         const float* __restrict__ beta, int M, float eps) {
    constexpr int C = NV * 256;
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    float v[NV * 8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (size_t)row * C + (i * 32 + lane) * 8));
        const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
        v[i * 8 + 0] = f0.x; v[i * 8 + 1] = f0.y; v[i * 8 + 2] = f1.x; v[i * 8 + 3] = f1.y;
        v[i * 8 + 4] = f2.x; v[i * 8 + 5] = f2.y; v[i * 8 + 6] = f3.x; v[i * 8 + 7] = f3.y;
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += v[i * 8 + j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * (1.0f / C);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV * 8; ++i) { const float d = v[i] - mean; sq = fmaf(d, d, sq); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * (1.0f / C) + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 8;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
        uint4 o;
        o.x = pack_bf16x2((v[i * 8 + 0] - mean) * rstd * g0.x + b0.x, (v[i * 8 + 1] - mean) * rstd * g0.y + b0.y);
        o.y = pack_bf16x2((v[i * 8 + 2] - mean) * rstd * g0.z + b0.z, (v[i * 8 + 3] - mean) * rstd * g0.w + b0.w);
        o.z = pack_bf16x2((v[i * 8 + 4] - mean) * rstd * g1.x + b1.x, (v[i * 8 + 5] - mean) * rstd * g1.y + b1.y);
        o.w = pack_bf16x2((v[i * 8 + 6] - mean) * rstd * g1.z + b1.z, (v[i * 8 + 7] - mean) * rstd * g1.w + b1.w);
        *reinterpret_cast<uint4*>(y + (size_t)row * C + c) = o;
    }
}

// ====================================================================== MHSA core
// softmax((q * 32^-1/2) k^T) v per head of dim 32 (mci.py:675-679), flash-style: the [N,N] score
// matrix the reference materialises never exists; scores live in mma.sync accumulators, the online
// softmax runs in fp32 with exp2.  qkv is the row-major [B*N, 3C] output of the qkv GEMM
// (q | k | v, each split into heads of 32 -- the reshape(B,N,3,h,32) of mci.py:669-672).
// grid (ceil(N/64), heads, B), 128 threads: a warp owns 16 query rows; K/V stream through a
// double-buffered cp.async ring in chunks of 64 keys.
constexpr int ATT_KC = 64;
constexpr int ATT_PITCH = 40;      // bf16 per smem row (32 + 8 pad): conflict-free LDS.32 / ldmatrix

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPEND> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NPEND) : "memory"); }

__global__ void __launch_bounds__(128)
attention_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int N, int C, float scale_log2e) {
    __shared__ __align__(16) bf16 Ks[2][ATT_KC * ATT_PITCH];
    __shared__ __align__(16) bf16 Vs[2][ATT_KC * ATT_PITCH];
    pdl_launch_dependents();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int head = blockIdx.y, b = blockIdx.z;
    const size_t ld = (size_t)3 * C;
    const bf16* base = qkv + (size_t)b * N * ld + head * 32;
    const int q0 = blockIdx.x * 64 + warp * 16;

    auto load_chunk = [&](int chunk, int buf) {
        // 64 keys x (K 64 B + V 64 B) = 512 x 16-B pieces, 4 per thread
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int piece = threadIdx.x + i * 128;
            const int which = piece >> 8;            // 0 = K, 1 = V
            const int key = (piece & 255) >> 2;
            const int part = piece & 3;
            const int gk = chunk * ATT_KC + key;
            bf16* dst = (which ? Vs[buf] : Ks[buf]) + key * ATT_PITCH + part * 8;
            if (gk < N) cp_async16(dst, base + (size_t)gk * ld + (which + 1) * C + part * 8);
            else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
        }
        cp_async_commit();
    };

    uint32_t qa[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        const int r0 = q0 + g, r1 = q0 + g + 8;
        const int c = ks * 16 + t * 2;
        qa[ks][0] = r0 < N ? *reinterpret_cast<const uint32_t*>(base + (size_t)r0 * ld + c) : 0u;
        qa[ks][1] = r1 < N ? *reinterpret_cast<const uint32_t*>(base + (size_t)r1 * ld + c) : 0u;
        qa[ks][2] = r0 < N ? *reinterpret_cast<const uint32_t*>(base + (size_t)r0 * ld + c + 8) : 0u;
        qa[ks][3] = r1 < N ? *reinterpret_cast<const uint32_t*>(base + (size_t)r1 * ld + c + 8) : 0u;
    }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }

    const int nchunks = (N + ATT_KC - 1) / ATT_KC;
    load_chunk(0, 0);
    for (int ch = 0; ch < nchunks; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < nchunks) { load_chunk(ch + 1, buf ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const bf16* Kb = Ks[buf];
        const bf16* Vb = Vs[buf];

        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const bf16* kp = Kb + (j * 8 + g) * ATT_PITCH + ks * 16 + t * 2;
                mma_bf16_16816(s[j], qa[ks], *reinterpret_cast<const uint32_t*>(kp), *reinterpret_cast<const uint32_t*>(kp + 8));
            }
        }
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int key = ch * ATT_KC + j * 8 + t * 2;
            s[j][0] = key < N ? s[j][0] * scale_log2e : -INFINITY;
            s[j][1] = key + 1 < N ? s[j][1] * scale_log2e : -INFINITY;
            s[j][2] = key < N ? s[j][2] * scale_log2e : -INFINITY;
            s[j][3] = key + 1 < N ? s[j][3] * scale_log2e : -INFINITY;
            mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
            mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float a0 = exp2f(m0 - mn0), a1 = exp2f(m1 - mn1);
        m0 = mn0; m1 = mn1;
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j][0] = exp2f(s[j][0] - mn0); s[j][1] = exp2f(s[j][1] - mn0);
            s[j][2] = exp2f(s[j][2] - mn1); s[j][3] = exp2f(s[j][3] - mn1);
            rs0 += s[j][0] + s[j][1];
            rs1 += s[j][2] + s[j][3];
        }
        l0 = l0 * a0 + rs0;
        l1 = l1 * a1 + rs1;
#pragma unroll
        for (int i = 0; i < 4; ++i) { o[i][0] *= a0; o[i][1] *= a0; o[i][2] *= a1; o[i][3] *= a1; }

#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {          // 16 keys per step
            uint32_t pa[4];
            pa[0] = pack_bf16x2(s[2 * k2][0], s[2 * k2][1]);
            pa[1] = pack_bf16x2(s[2 * k2][2], s[2 * k2][3]);
            pa[2] = pack_bf16x2(s[2 * k2 + 1][0], s[2 * k2 + 1][1]);
            pa[3] = pack_bf16x2(s[2 * k2 + 1][2], s[2 * k2 + 1][3]);
#pragma unroll
            for (int nt = 0; nt < 4; nt += 2) {   // two 8-wide dim tiles per ldmatrix.x4
                const int mi = lane >> 3, r = lane & 7;
                const int key = k2 * 16 + (mi & 1) * 8 + r;
                const int dim0 = (nt + (mi >> 1)) * 8;
                uint32_t vb[4];
                ldmatrix_x4_trans(vb, Vb + key * ATT_PITCH + dim0);
                mma_bf16_16816(o[nt], pa, vb[0], vb[1]);
                mma_bf16_16816(o[nt + 1], pa, vb[2], vb[3]);
            }
        }
        __syncthreads();
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    const int r0 = q0 + g, r1 = q0 + g + 8;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const int c = head * 32 + nt * 8 + t * 2;
        if (r0 < N) *reinterpret_cast<uint32_t*>(out + ((size_t)b * N + r0) * C + c) = pack_bf16x2(o[nt][0] * inv0, o[nt][1] * inv0);
        if (r1 < N) *reinterpret_cast<uint32_t*>(out + ((size_t)b * N + r1) * C + c) = pack_bf16x2(o[nt][2] * inv1, o[nt][3] * inv1);
    }
}

// ====================================================================== conv_exp squeeze-excite
// SEBlock (mci.py:72-81) on the dw3x3 output c [B, HW, C]:  s = sigmoid(We relu(Wr mean_hw(c) + br) + be),
// tokens = GELU(c * s) (order act(se(conv(x))), mci.py:198), emitted directly as [B, HW, 3072].
// grid (C/64, B): per-channel mean over the HW pixels of one image.
__global__ void __launch_bounds__(256)
se_pool_kernel(const bf16* __restrict__ c, float* __restrict__ pooled, int HW, int C) {
    __shared__ float red[8][64];
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y, c0 = blockIdx.x * 64;
    const int cp = threadIdx.x & 31, sl = threadIdx.x >> 5;
    float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
    for (int p = sl; p < HW; p += 8) {          // unrolled: 8 independent loads in flight per thread (the loop is pure load latency)
        const float2 v = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(c + ((size_t)b * HW + p) * C + c0 + 2 * cp)));
        a0 += v.x; a1 += v.y;
    }
    red[sl][2 * cp] = a0; red[sl][2 * cp + 1] = a1;
    __syncthreads();
    if (threadIdx.x < 64) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
        pooled[(size_t)b * C + c0 + threadIdx.x] = s / (float)HW;
    }
}
// grid (RD/8, B): warp per reduced channel.
__global__ void __launch_bounds__(256)
se_reduce_kernel(const float* __restrict__ pooled, const bf16* __restrict__ wr /*[RD][C]*/, const float* __restrict__ br,
                 float* __restrict__ r, int C, int RD) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y, j = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= RD) return;
    float a = 0.f;
#pragma unroll 4
    for (int c = lane * 8; c < C; c += 256) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(wr + (size_t)j * C + c));
        const float4 p0 = __ldg(reinterpret_cast<const float4*>(pooled + (size_t)b * C + c));
        const float4 p1 = __ldg(reinterpret_cast<const float4*>(pooled + (size_t)b * C + c + 4));
        const float2 w0 = unpack_bf16x2(u.x), w1 = unpack_bf16x2(u.y), w2 = unpack_bf16x2(u.z), w3 = unpack_bf16x2(u.w);
        a += w0.x * p0.x + w0.y * p0.y + w1.x * p0.z + w1.y * p0.w + w2.x * p1.x + w2.y * p1.y + w3.x * p1.z + w3.y * p1.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) r[(size_t)b * RD + j] = fmaxf(a + __ldg(br + j), 0.f);
}
// grid (C/64, B): expand + sigmoid for a 64-channel slice (one warp per channel: 192-term dot product across the lanes),
// then scale + GELU all HW pixels of that slice.
__global__ void __launch_bounds__(256)
se_expand_scale_gelu_kernel(const bf16* __restrict__ c, const float* __restrict__ r, const bf16* __restrict__ we /*[C][RD]*/,
                            const float* __restrict__ be, bf16* __restrict__ tokens_or_null, const IoBlock* __restrict__ io, int HW, int C, int RD) {
    pdl_launch_dependents();
    pdl_wait();
    bf16* __restrict__ tokens = tokens_or_null ? tokens_or_null : reinterpret_cast<bf16*>(io->final_out);
    const size_t out_img_stride = tokens_or_null ? (size_t)HW * C : (size_t)io->final_image_stride;
    __shared__ float rs[256];
    __shared__ float ss[64];
    const int b = blockIdx.y, c0 = blockIdx.x * 64;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < RD; i += 256) rs[i] = r[(size_t)b * RD + i];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) {                       // 8 warps x 8 channels; unrolled so the 8 channels' weight loads overlap
        const int ch = c0 + warp * 8 + k;
        float a = 0.f;
#pragma unroll 4
        for (int j = lane * 2; j < RD; j += 64) {
            const float2 w = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(we + (size_t)ch * RD + j)));
            a = fmaf(w.x, rs[j], a);
            a = fmaf(w.y, rs[j + 1], a);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) ss[warp * 8 + k] = 1.0f / (1.0f + __expf(-(a + __ldg(be + ch))));
    }
    __syncthreads();
    const int cp = threadIdx.x & 31;
    const float s0 = ss[2 * cp], s1 = ss[2 * cp + 1];
    // 8 pixels per iteration, all loads issued before the first store (the stores go through a pointer the compiler cannot
    // prove disjoint from `c`, so a plain loop serialises one load latency per pixel: 22 us of the 31 at batch 1)
    for (int p0 = threadIdx.x >> 5; p0 < HW; p0 += 64) {
        uint32_t u[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int p = p0 + 8 * k;
            u[k] = p < HW ? __ldg(reinterpret_cast<const uint32_t*>(c + ((size_t)b * HW + p) * C + c0 + 2 * cp)) : 0u;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int p = p0 + 8 * k;
            if (p < HW) {
                const float2 v = unpack_bf16x2(u[k]);
                *reinterpret_cast<uint32_t*>(tokens + (size_t)b * out_img_stride + (size_t)p * C + c0 + 2 * cp) = pack_bf16x2(gelu_erf(v.x * s0), gelu_erf(v.y * s1));
            }
        }
    }
}

// ====================================================================== IO plumbing
struct PeerList { int n; void* p[FVHD_MAX_PEERS]; };
struct ScatterList { int n; void* p[FVHD_MAX_SCATTER]; };
__global__ void set_io_kernel(IoBlock* io, const void* images, void* final_out, void* tokens_out, long long final_image_stride, const PeerList peers,
                              const ScatterList sc) {
    pdl_launch_dependents();
    pdl_wait();                 // the previous forward's last kernels may still be reading the block
    if (threadIdx.x == 0) {
        io->images = images;
        io->final_out = final_out;
        io->tokens_out = tokens_out;
        io->final_image_stride = final_image_stride;
        io->n_peers = peers.n;
        for (int i = 0; i < FVHD_MAX_PEERS; ++i) io->peer_out[i] = i < peers.n ? peers.p[i] : nullptr;
        io->scatter_n = sc.n;
    }
    if (threadIdx.x < FVHD_MAX_SCATTER) io->scatter[threadIdx.x] = (int)threadIdx.x < sc.n ? sc.p[threadIdx.x] : nullptr;
}
// tokens (workspace) -> caller buffer, 16 B per thread-iteration
__global__ void __launch_bounds__(256)
copy_tokens_kernel(const uint4* __restrict__ src, const IoBlock* __restrict__ io, size_t n16) {
    pdl_launch_dependents();
    pdl_wait();
    uint4* __restrict__ dst = reinterpret_cast<uint4*>(io->tokens_out);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

}  // namespace fvhd
