// LLM prefill (row f3: the caller of encode_images, llava_qwen.py:57-143 -> transformers Qwen2ForCausalLM.forward on the spliced
// [1, L, H] embedding sequence; TTFT = this pass + the first token, FastVLMModel.swift:114-138).  The four GEMMs of a decoder layer
// run on the tcgen05 GEMM kernel of the tower (gemm_tcgen05.cuh: bias / residual epilogues, split-K at small M); this header holds
// the glue kernels a Qwen2 layer needs around them:
//     rmsnorm_kernel        Qwen2RMSNorm: y = w * x * rsqrt(mean(x^2) + eps)                        (modeling_qwen2.py Qwen2RMSNorm.forward)
//     rope_table_kernel     cos / sin of pos * theta^(-2d/D), fp32, built once per handle          (Qwen2RotaryEmbedding)
//     causal_attn_kernel<D> rotate_half RoPE on q and k (table lookups while the tiles are staged), K (post-RoPE) / V rows -> KV cache,
//                           causal GQA attention, flash-style over 64-key tiles, fp32 softmax / accumulation on the FMA pipes
//                           (L = 287 tokens x 14 heads: 0.3 GFLOP per layer -- latency, not throughput, matters here)
//     silu_mul_kernel       SwiGLU gate: h = silu(gate) * up                                         (Qwen2MLP.forward)
//     argmax_kernel         first token = argmax of the last position's logits
#pragma once
#include "mixer_tz.cuh"       // ffma2
#include "stem_attn_se.cuh"   // mma_bf16_16816, ldmatrix_x4_trans

namespace fvhd {

// one warp per row; H % 8 == 0
__global__ void __launch_bounds__(256)
rmsnorm_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, const float* __restrict__ w, int rows, int H, float eps) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)row * H);
    const int nv = H / 8;
    float ss = 0.f;
    for (int v = lane; v < nv; v += 32) {
        const uint4 u = __ldg(xr + v);
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float r = rsqrtf(ss / (float)H + eps);
    uint4* yr = reinterpret_cast<uint4*>(y + (size_t)row * H);
    for (int v = lane; v < nv; v += 32) {
        const uint4 u = __ldg(xr + v);
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(w) + 2 * v), w1 = __ldg(reinterpret_cast<const float4*>(w) + 2 * v + 1);
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        uint4 o;
        o.x = pack_bf16x2(a.x * r * w0.x, a.y * r * w0.y);
        o.y = pack_bf16x2(b.x * r * w0.z, b.y * r * w0.w);
        o.z = pack_bf16x2(c.x * r * w1.x, c.y * r * w1.y);
        o.w = pack_bf16x2(d.x * r * w1.z, d.y * r * w1.w);
        yr[v] = o;
    }
}

// table[pos][d] = (cos, sin)(pos * theta^(-2 d / D)), d < D / 2
__global__ void rope_table_kernel(float2* __restrict__ table, int max_pos, int half, float theta) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= max_pos * half) return;
    const int pos = i / half, d = i - pos * half;
    const float inv = exp2f(-(float)d / (float)half * log2f(theta));          // theta^(-2d/D), fp32 as in the reference
    float s, c;
    sincosf((float)pos * inv, &s, &c);
    table[i] = make_float2(c, s);
}

constexpr int LLM_ATTN_QB = 16;                 // queries per CTA: 4 warps x 4 queries (287 tokens x 14 heads -> 252 CTAs, two per SM)
constexpr int LLM_ATTN_THREADS = 128;
template <int D> struct LlmAttnSmem { static constexpr size_t BYTES = (size_t)2 * 64 * (D + 8) * 2 + (size_t)LLM_ATTN_QB * D * 4; };

// Causal grouped-query attention over the fused qkv rows [L, (heads + 2 kv) * D] (q heads, k heads, v heads), positions 0..L-1.
// RoPE (rotate_half: pairs (d, d + D/2), Qwen2 apply_rotary_pos_emb) is applied while q and the K tiles are staged; the rotated K is
// rounded to bf16 as the reference's bf16 tensors are.  The CTAs of the LAST query block of each group's first q head walk every key
// tile: they also write K (post-RoPE) and V into the KV cache.  grid (ceil(L / 16), heads); 4 warps x 4 queries; 64-key tiles in smem;
// the dot products and the P V accumulation run as packed FFMA2 (two keys / two output dims per instruction).
// out [L, heads * D] bf16.  scale_log2 = D^-0.5 * log2(e).
template <int D>
__global__ void __launch_bounds__(LLM_ATTN_THREADS)
causal_attn_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, const float2* __restrict__ rope, bf16* __restrict__ k_cache,
                   bf16* __restrict__ v_cache, int L, int heads, int kv_heads, float scale_log2) {
    constexpr int TK = 64, QB = LLM_ATTN_QB, NQ = 4, NT = LLM_ATTN_THREADS, KP = D + 8;   // key pitch (halves): +16 B keeps the 16-B row reads conflict-free
    constexpr int DW = D / 64;                                   // bf16x2 words of V / o per lane
    extern __shared__ __align__(16) uint8_t attn_smem[];         // LlmAttnSmem<D>::BYTES
    bf16* Ks = reinterpret_cast<bf16*>(attn_smem);
    bf16* Vs = Ks + TK * KP;
    float* Qs = reinterpret_cast<float*>(Vs + TK * KP);
    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hq = blockIdx.y, hk = hq / (heads / kv_heads);
    const int q0 = blockIdx.x * QB;
    const int ld = (heads + 2 * kv_heads) * D;
    constexpr int HALF = D / 2;
    const bool write_cache = k_cache != nullptr && blockIdx.x == gridDim.x - 1 && hq % (heads / kv_heads) == 0;
    pdl_wait();
    // this CTA's queries: RoPE, pre-scaled, fp32
    for (int i = threadIdx.x; i < QB * HALF; i += NT) {
        const int qi = i / HALF, d = i - qi * HALF;
        float x1 = 0.f, x2 = 0.f;
        float2 cs = make_float2(1.f, 0.f);
        if (q0 + qi < L) {
            const bf16* qp = qkv + (size_t)(q0 + qi) * ld + (size_t)hq * D + d;
            x1 = __bfloat162float(qp[0]); x2 = __bfloat162float(qp[HALF]);
            cs = rope[(size_t)(q0 + qi) * HALF + d];
        }
        Qs[qi * D + d] = (x1 * cs.x - x2 * cs.y) * scale_log2;
        Qs[qi * D + d + HALF] = (x2 * cs.x + x1 * cs.y) * scale_log2;
    }
    float m[NQ], l[NQ];
    float2 o[NQ][DW];                                             // output dims 2 (lane + 32 e), +1
#pragma unroll
    for (int t = 0; t < NQ; ++t) {
        m[t] = -1e30f; l[t] = 0.f;
#pragma unroll
        for (int e = 0; e < DW; ++e) o[t][e] = make_float2(0.f, 0.f);
    }
    const int q_last = min(q0 + QB, L) - 1;
    const int qw = q0 + warp * NQ;                               // this warp's first query
    for (int k0 = 0; k0 <= q_last; k0 += TK) {
        __syncthreads();                                         // previous tile consumed (and Qs written)
        for (int i = threadIdx.x; i < TK * D / 16; i += NT) {        // one thread: the 8-wide chunk c and its rotation partner c + D/16
            const int kj = i / (D / 16), c = i - kj * (D / 16);
            uint4 ka = make_uint4(0, 0, 0, 0), kb = ka, va = ka, vb = ka;
            if (k0 + kj < L) {
                const bf16* row = qkv + (size_t)(k0 + kj) * ld;
                const bf16* kp = row + (size_t)(heads + hk) * D + c * 8;
                const bf16* vp = row + (size_t)(heads + kv_heads + hk) * D + c * 8;
                const uint4 k1 = *reinterpret_cast<const uint4*>(kp), k2 = *reinterpret_cast<const uint4*>(kp + HALF);
                va = *reinterpret_cast<const uint4*>(vp); vb = *reinterpret_cast<const uint4*>(vp + HALF);
                const float4* tb = reinterpret_cast<const float4*>(rope + (size_t)(k0 + kj) * HALF + c * 8);     // 8 (cos, sin) pairs
                const uint32_t w1[4] = {k1.x, k1.y, k1.z, k1.w}, w2[4] = {k2.x, k2.y, k2.z, k2.w};
                uint32_t o1[4], o2[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 t = tb[e];                                   // (cos, sin) of dims 2e and 2e + 1 of the chunk
                    const float2 a = unpack_bf16x2(w1[e]), b = unpack_bf16x2(w2[e]);
                    o1[e] = pack_bf16x2(a.x * t.x - b.x * t.y, a.y * t.z - b.y * t.w);
                    o2[e] = pack_bf16x2(b.x * t.x + a.x * t.y, b.y * t.z + a.y * t.w);
                }
                ka = make_uint4(o1[0], o1[1], o1[2], o1[3]); kb = make_uint4(o2[0], o2[1], o2[2], o2[3]);
                if (write_cache) {
                    bf16* kc = k_cache + ((size_t)(k0 + kj) * kv_heads + hk) * D + c * 8;
                    bf16* vc = v_cache + ((size_t)(k0 + kj) * kv_heads + hk) * D + c * 8;
                    *reinterpret_cast<uint4*>(kc) = ka; *reinterpret_cast<uint4*>(kc + HALF) = kb;
                    *reinterpret_cast<uint4*>(vc) = va; *reinterpret_cast<uint4*>(vc + HALF) = vb;
                }
            }
            *reinterpret_cast<uint4*>(Ks + kj * KP + c * 8) = ka; *reinterpret_cast<uint4*>(Ks + kj * KP + c * 8 + HALF) = kb;
            *reinterpret_cast<uint4*>(Vs + kj * KP + c * 8) = va; *reinterpret_cast<uint4*>(Vs + kj * KP + c * 8 + HALF) = vb;
        }
        __syncthreads();
        if (k0 > qw + NQ - 1) continue;                          // every key of the tile is in this warp's future (barriers stay uniform)
        // ---- scores: lane owns keys `lane` and `lane + 32` of the tile; one FFMA2 per (query, dim) covers both keys
        float2 s[NQ];
#pragma unroll
        for (int t = 0; t < NQ; ++t) s[t] = make_float2(0.f, 0.f);
#pragma unroll 2
        for (int c = 0; c < D / 8; ++c) {
            const uint4 ka = *reinterpret_cast<const uint4*>(Ks + lane * KP + c * 8);
            const uint4 kb = *reinterpret_cast<const uint4*>(Ks + (lane + 32) * KP + c * 8);
            float2 kk[8];                                        // (key lane, key lane + 32) for the chunk's 8 dims
            { const float2 a0 = unpack_bf16x2(ka.x), a1 = unpack_bf16x2(ka.y), a2 = unpack_bf16x2(ka.z), a3 = unpack_bf16x2(ka.w);
              const float2 b0 = unpack_bf16x2(kb.x), b1 = unpack_bf16x2(kb.y), b2 = unpack_bf16x2(kb.z), b3 = unpack_bf16x2(kb.w);
              kk[0] = make_float2(a0.x, b0.x); kk[1] = make_float2(a0.y, b0.y); kk[2] = make_float2(a1.x, b1.x); kk[3] = make_float2(a1.y, b1.y);
              kk[4] = make_float2(a2.x, b2.x); kk[5] = make_float2(a2.y, b2.y); kk[6] = make_float2(a3.x, b3.x); kk[7] = make_float2(a3.y, b3.y); }
#pragma unroll
            for (int t = 0; t < NQ; ++t) {
                const float4 qa = *reinterpret_cast<const float4*>(Qs + (warp * NQ + t) * D + c * 8);
                const float4 qb = *reinterpret_cast<const float4*>(Qs + (warp * NQ + t) * D + c * 8 + 4);
                ffma2(s[t], kk[0], make_float2(qa.x, qa.x)); ffma2(s[t], kk[1], make_float2(qa.y, qa.y));
                ffma2(s[t], kk[2], make_float2(qa.z, qa.z)); ffma2(s[t], kk[3], make_float2(qa.w, qa.w));
                ffma2(s[t], kk[4], make_float2(qb.x, qb.x)); ffma2(s[t], kk[5], make_float2(qb.y, qb.y));
                ffma2(s[t], kk[6], make_float2(qb.z, qb.z)); ffma2(s[t], kk[7], make_float2(qb.w, qb.w));
            }
        }
        // ---- online softmax per query, then o += P V
#pragma unroll
        for (int t = 0; t < NQ; ++t) {
            const int qi = qw + t;
            const bool a_ok = k0 + lane <= qi && k0 + lane < L, b_ok = k0 + lane + 32 <= qi && k0 + lane + 32 < L;
            const float sa = a_ok ? s[t].x : -1e30f, sb = b_ok ? s[t].y : -1e30f;
            float mx = fmaxf(sa, sb);
#pragma unroll
            for (int of = 16; of > 0; of >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, of));
            const float mn = fmaxf(m[t], mx);
            const float pa = a_ok ? exp2f(sa - mn) : 0.f, pb = b_ok ? exp2f(sb - mn) : 0.f;
            float ps = pa + pb;
#pragma unroll
            for (int of = 16; of > 0; of >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, of);
            const float corr = exp2f(m[t] - mn);
            m[t] = mn;
            l[t] = l[t] * corr + ps;
#pragma unroll
            for (int e = 0; e < DW; ++e) { o[t][e].x *= corr; o[t][e].y *= corr; }
            s[t] = make_float2(pa, pb);
        }
#pragma unroll 4
        for (int kj = 0; kj < 32; ++kj) {
            float2 va[DW], vb[DW];
#pragma unroll
            for (int e = 0; e < DW; ++e) {
                va[e] = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(Vs + kj * KP + 2 * (lane + 32 * e)));
                vb[e] = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(Vs + (kj + 32) * KP + 2 * (lane + 32 * e)));
            }
#pragma unroll
            for (int t = 0; t < NQ; ++t) {
                const float pa = __shfl_sync(0xffffffffu, s[t].x, kj), pb = __shfl_sync(0xffffffffu, s[t].y, kj);
#pragma unroll
                for (int e = 0; e < DW; ++e) {
                    ffma2(o[t][e], va[e], make_float2(pa, pa));
                    ffma2(o[t][e], vb[e], make_float2(pb, pb));
                }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < NQ; ++t) {
        const int qi = qw + t;
        if (qi < L) {
            const float inv = 1.f / l[t];
#pragma unroll
            for (int e = 0; e < DW; ++e)
                *reinterpret_cast<uint32_t*>(out + (size_t)qi * heads * D + (size_t)hq * D + 2 * (lane + 32 * e)) = pack_bf16x2(o[t][e].x * inv, o[t][e].y * inv);
        }
    }
}

// The same attention on the tensor cores (mma.sync m16n8k16 bf16, fp32 accumulate; flash-attention-2 register layout as in the tower's
// attention_kernel): one warp = 16 queries, CTA = 2 warps = 32 queries x one head, 64-key tiles.  Q (RoPE'd, bf16) is staged once and
// held as A fragments; S = Q K^T uses K rows as the col-major B operand, P V reads V through ldmatrix.trans.  ~27x fewer instructions
// per query than the FMA-pipe kernel above (which stays selectable: FVHD_LLM_ATTN=f).
constexpr int LLM_MMA_QB = 32, LLM_MMA_THREADS = 64;
template <int D> struct LlmAttnMmaSmem { static constexpr size_t BYTES = (size_t)(2 * 64 + LLM_MMA_QB) * (D + 8) * 2; };

template <int D>
__global__ void __launch_bounds__(LLM_MMA_THREADS)
causal_attn_mma_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, const float2* __restrict__ rope, bf16* __restrict__ k_cache,
                       bf16* __restrict__ v_cache, int L, int heads, int kv_heads, float scale_log2) {
    constexpr int TK = 64, QB = LLM_MMA_QB, NT = LLM_MMA_THREADS, KP = D + 8, HALF = D / 2, KS = D / 16, ND = D / 8;
    extern __shared__ __align__(16) uint8_t attn_mma_smem[];
    bf16* Ks = reinterpret_cast<bf16*>(attn_mma_smem);
    bf16* Vs = Ks + TK * KP;
    bf16* Qs = Vs + TK * KP;
    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int hq = blockIdx.y, hk = hq / (heads / kv_heads);
    const int q0 = blockIdx.x * QB;
    const int ld = (heads + 2 * kv_heads) * D;
    const bool write_cache = k_cache != nullptr && blockIdx.x == gridDim.x - 1 && hq % (heads / kv_heads) == 0;
    pdl_wait();
    // ---- Q tile: RoPE, bf16 (as the reference's bf16 q), rows beyond L zero
    for (int i = threadIdx.x; i < QB * HALF; i += NT) {
        const int qi = i / HALF, d = i - qi * HALF;
        float o1 = 0.f, o2 = 0.f;
        if (q0 + qi < L) {
            const bf16* qp = qkv + (size_t)(q0 + qi) * ld + (size_t)hq * D + d;
            const float x1 = __bfloat162float(qp[0]), x2 = __bfloat162float(qp[HALF]);
            const float2 cs = rope[(size_t)(q0 + qi) * HALF + d];
            o1 = x1 * cs.x - x2 * cs.y; o2 = x2 * cs.x + x1 * cs.y;
        }
        Qs[qi * KP + d] = __float2bfloat16_rn(o1);
        Qs[qi * KP + d + HALF] = __float2bfloat16_rn(o2);
    }
    __syncthreads();
    const int qw = q0 + warp * 16;                               // this warp's first query
    const int r0 = qw + g, r1 = qw + g + 8;
    uint32_t qa[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        const bf16* qp = Qs + (warp * 16 + g) * KP + ks * 16 + t * 2;
        qa[ks][0] = *reinterpret_cast<const uint32_t*>(qp);
        qa[ks][1] = *reinterpret_cast<const uint32_t*>(qp + 8 * KP);
        qa[ks][2] = *reinterpret_cast<const uint32_t*>(qp + 8);
        qa[ks][3] = *reinterpret_cast<const uint32_t*>(qp + 8 * KP + 8);
    }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float o[ND][4];
#pragma unroll
    for (int i = 0; i < ND; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }

    const int q_last = min(q0 + QB, L) - 1;
    for (int k0 = 0; k0 <= q_last; k0 += TK) {
        __syncthreads();                                         // previous tile consumed
        for (int i = threadIdx.x; i < TK * D / 16; i += NT) {        // one thread: the 8-wide chunk c and its rotation partner c + D/16
            const int kj = i / (D / 16), c = i - kj * (D / 16);
            uint4 ka = make_uint4(0, 0, 0, 0), kb = ka, va = ka, vb = ka;
            if (k0 + kj < L) {
                const bf16* row = qkv + (size_t)(k0 + kj) * ld;
                const bf16* kp = row + (size_t)(heads + hk) * D + c * 8;
                const bf16* vp = row + (size_t)(heads + kv_heads + hk) * D + c * 8;
                const uint4 k1 = *reinterpret_cast<const uint4*>(kp), k2 = *reinterpret_cast<const uint4*>(kp + HALF);
                va = *reinterpret_cast<const uint4*>(vp); vb = *reinterpret_cast<const uint4*>(vp + HALF);
                const float4* tb = reinterpret_cast<const float4*>(rope + (size_t)(k0 + kj) * HALF + c * 8);     // 8 (cos, sin) pairs
                const uint32_t w1[4] = {k1.x, k1.y, k1.z, k1.w}, w2[4] = {k2.x, k2.y, k2.z, k2.w};
                uint32_t o1[4], o2[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 tt = tb[e];
                    const float2 a = unpack_bf16x2(w1[e]), b = unpack_bf16x2(w2[e]);
                    o1[e] = pack_bf16x2(a.x * tt.x - b.x * tt.y, a.y * tt.z - b.y * tt.w);
                    o2[e] = pack_bf16x2(b.x * tt.x + a.x * tt.y, b.y * tt.z + a.y * tt.w);
                }
                ka = make_uint4(o1[0], o1[1], o1[2], o1[3]); kb = make_uint4(o2[0], o2[1], o2[2], o2[3]);
                if (write_cache) {
                    bf16* kc = k_cache + ((size_t)(k0 + kj) * kv_heads + hk) * D + c * 8;
                    bf16* vc = v_cache + ((size_t)(k0 + kj) * kv_heads + hk) * D + c * 8;
                    *reinterpret_cast<uint4*>(kc) = ka; *reinterpret_cast<uint4*>(kc + HALF) = kb;
                    *reinterpret_cast<uint4*>(vc) = va; *reinterpret_cast<uint4*>(vc + HALF) = vb;
                }
            }
            *reinterpret_cast<uint4*>(Ks + kj * KP + c * 8) = ka; *reinterpret_cast<uint4*>(Ks + kj * KP + c * 8 + HALF) = kb;
            *reinterpret_cast<uint4*>(Vs + kj * KP + c * 8) = va; *reinterpret_cast<uint4*>(Vs + kj * KP + c * 8 + HALF) = vb;
        }
        __syncthreads();
        if (k0 > qw + 15) continue;                              // every key of the tile lies in this warp's future (barriers stay uniform)
        // ---- S = Q K^T for 64 keys
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const bf16* kp = Ks + (j * 8 + g) * KP + ks * 16 + t * 2;
                mma_bf16_16816(s[j], qa[ks], *reinterpret_cast<const uint32_t*>(kp), *reinterpret_cast<const uint32_t*>(kp + 8));
            }
        }
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int key = k0 + j * 8 + t * 2;                  // causal mask: key <= query row (and < L)
            s[j][0] = (key <= r0 && key < L) ? s[j][0] * scale_log2 : -INFINITY;
            s[j][1] = (key + 1 <= r0 && key + 1 < L) ? s[j][1] * scale_log2 : -INFINITY;
            s[j][2] = (key <= r1 && key < L) ? s[j][2] * scale_log2 : -INFINITY;
            s[j][3] = (key + 1 <= r1 && key + 1 < L) ? s[j][3] * scale_log2 : -INFINITY;
            mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
            mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        // a row whose keys are all masked in this tile (k0 > row) keeps its state: its first tile (k0 = 0) always has key 0 unmasked
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float a0 = mn0 == -INFINITY ? 1.f : exp2f(m0 - mn0), a1 = mn1 == -INFINITY ? 1.f : exp2f(m1 - mn1);
        m0 = mn0; m1 = mn1;
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j][0] = mn0 == -INFINITY ? 0.f : exp2f(s[j][0] - mn0); s[j][1] = mn0 == -INFINITY ? 0.f : exp2f(s[j][1] - mn0);
            s[j][2] = mn1 == -INFINITY ? 0.f : exp2f(s[j][2] - mn1); s[j][3] = mn1 == -INFINITY ? 0.f : exp2f(s[j][3] - mn1);
            rs0 += s[j][0] + s[j][1];
            rs1 += s[j][2] + s[j][3];
        }
        l0 = l0 * a0 + rs0;
        l1 = l1 * a1 + rs1;
#pragma unroll
        for (int i = 0; i < ND; ++i) { o[i][0] *= a0; o[i][1] *= a0; o[i][2] *= a1; o[i][3] *= a1; }
        // ---- O += P V
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {          // 16 keys per step
            uint32_t pa[4];
            pa[0] = pack_bf16x2(s[2 * k2][0], s[2 * k2][1]);
            pa[1] = pack_bf16x2(s[2 * k2][2], s[2 * k2][3]);
            pa[2] = pack_bf16x2(s[2 * k2 + 1][0], s[2 * k2 + 1][1]);
            pa[3] = pack_bf16x2(s[2 * k2 + 1][2], s[2 * k2 + 1][3]);
#pragma unroll
            for (int nt = 0; nt < ND; nt += 2) {   // two 8-wide dim tiles per ldmatrix.x4
                const int mi = lane >> 3, r = lane & 7;
                const int key = k2 * 16 + (mi & 1) * 8 + r;
                const int dim0 = (nt + (mi >> 1)) * 8;
                uint32_t vb[4];
                ldmatrix_x4_trans(vb, Vs + key * KP + dim0);
                mma_bf16_16816(o[nt], pa, vb[0], vb[1]);
                mma_bf16_16816(o[nt + 1], pa, vb[2], vb[3]);
            }
        }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
#pragma unroll
    for (int nt = 0; nt < ND; ++nt) {
        const size_t c = (size_t)hq * D + nt * 8 + t * 2;
        if (r0 < L) *reinterpret_cast<uint32_t*>(out + (size_t)r0 * heads * D + c) = pack_bf16x2(o[nt][0] * inv0, o[nt][1] * inv0);
        if (r1 < L) *reinterpret_cast<uint32_t*>(out + (size_t)r1 * heads * D + c) = pack_bf16x2(o[nt][2] * inv1, o[nt][3] * inv1);
    }
}

// gu [rows, 2 I] (gate columns then up columns) -> h [rows, I] = silu(gate) * up
__global__ void __launch_bounds__(256)
silu_mul_kernel(const bf16* __restrict__ gu, bf16* __restrict__ hm, int rows, int I) {
    pdl_launch_dependents();
    pdl_wait();
    const int nv = I / 8;
    const long total = (long)rows * nv;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int r = (int)(i / nv), v = (int)(i - (long)r * nv);
        const uint4 g = *reinterpret_cast<const uint4*>(gu + (size_t)r * 2 * I + v * 8);
        const uint4 u = *reinterpret_cast<const uint4*>(gu + (size_t)r * 2 * I + I + v * 8);
        const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, uw[4] = {u.x, u.y, u.z, u.w};
        uint32_t ow[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 a = unpack_bf16x2(gw[k]), b = unpack_bf16x2(uw[k]);
            ow[k] = pack_bf16x2(a.x / (1.f + __expf(-a.x)) * b.x, a.y / (1.f + __expf(-a.y)) * b.y);
        }
        *reinterpret_cast<uint4*>(hm + (size_t)r * I + v * 8) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
}

// one CTA: index of the largest of n bf16 logits (lowest index on ties, as torch.argmax)
__global__ void __launch_bounds__(1024)
argmax_kernel(const bf16* __restrict__ logits, int n, int* __restrict__ out) {
    __shared__ float bv[32];
    __shared__ int bi[32];
    pdl_launch_dependents();
    pdl_wait();
    float best = -INFINITY;
    int idx = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += 1024) {
        const float v = __bfloat162float(logits[i]);
        if (v > best || (v == best && i < idx)) { best = v; idx = i; }
    }
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, of);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, of);
        if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
    }
    if ((threadIdx.x & 31) == 0) { bv[threadIdx.x >> 5] = best; bi[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x < 32) {
        best = bv[threadIdx.x]; idx = bi[threadIdx.x];
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, of);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, of);
            if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
        }
        if (threadIdx.x == 0) *out = idx;
    }
}

}  // namespace fvhd
