// Bandwidth-bound helper kernels of the SR3 step (everything that is not a tensor-core tile):
// GroupNorm apply (+SiLU) with channel concat, fp32->bf16 cast / nearest 2x upsample, row softmax,
// noise-level embedding MLP + FiLM projections, weight packing, layout conversion at the API boundary.
#pragma once
#include "gemm_tcgen05.cuh"

namespace sr3 {

__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

__device__ __forceinline__ uint2 pack_bf16x4(float a, float b, float c, float d) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    return u;
}
// precise mode: the part of x that bf16(x) loses, itself rounded to bf16 (x ~ hi + lo to ~2^-17 relative)
__device__ __forceinline__ float bf16_residual(float x) { return x - __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ uint2 pack_bf16x4_residual(float a, float b, float c, float d) {
    return pack_bf16x4(bf16_residual(a), bf16_residual(b), bf16_residual(c), bf16_residual(d));
}
// One 4-channel vector of a bf16 operand tensor: `o` = element offset of the high halves; precise mode (lo_off != 0) also stores the low halves
__device__ __forceinline__ void store_operand4(__nv_bfloat16* base, long long o, long long lo_off, float a, float b, float c, float d) {
    *reinterpret_cast<uint2*>(base + o) = pack_bf16x4(a, b, c, d);
    if (lo_off) *reinterpret_cast<uint2*>(base + o + lo_off) = pack_bf16x4_residual(a, b, c, d);
}

// ------------------------------------------------------------------------------------------------ dropout (unet.py:86, block2 only)
// keep-mask of element (b, c, pixel) of a [B][HW][C] tensor: either an injected mask (tests: the reference's own masks, uint8 NCHW) or
// Philox4x32-10 keyed by (seed; vector index, layer).  Forward and backward evaluate the same function.
struct DropSpec {
    const unsigned char* mask;     // optional [B][C][HW] (1 = keep)
    float p;                       // drop probability; 0 = no dropout
    unsigned int layer;
    unsigned long long seed;
};
__device__ __forceinline__ void drop_scale4(const DropSpec& d, int b, int c, int pix, int C, int HW, float (&s)[4]) {
    const float keep_scale = 1.0f / (1.0f - d.p);
    if (d.mask) {
#pragma unroll
        for (int j = 0; j < 4; ++j) s[j] = d.mask[(static_cast<long long>(b) * C + c + j) * HW + pix] ? keep_scale : 0.f;
        return;
    }
    const unsigned long long vec = (static_cast<unsigned long long>(b) * HW + pix) * (C >> 2) + (c >> 2);
    uint32_t ctr[4] = {static_cast<uint32_t>(vec), static_cast<uint32_t>(vec >> 32), d.layer, 0x5d0u};
    philox4x32_10(ctr, static_cast<uint32_t>(d.seed), static_cast<uint32_t>(d.seed >> 32));
#pragma unroll
    for (int j = 0; j < 4; ++j) s[j] = (static_cast<float>(ctr[j]) * 2.3283064365386963e-10f >= d.p) ? keep_scale : 0.f;
}

// ---------------------------------------------------------------------------------------------
// GroupNorm apply: a = [silu]( (x - mean_g) * rstd_g * gamma_c + beta_c ), x = concat(src0, src1) along channels.
// Statistics come from the per-(image, channel) sums the producing GEMM epilogues accumulated.
// reference: nn.GroupNorm(groups, C, eps=1e-5) + Swish of Block (unet.py:80-91), torch.cat of unet.py:255.
struct PrepParams {
    const float* src0; const float* src1;
    const double* st0; const double* st1;   // [B][C0][2], [B][C1][2]: fp64 (sum, sumsq) of the producing epilogues
    int C0, C1;
    const float* gamma; const float* beta;
    int groups, HW, pix_per_block, silu;
    float eps;
    __nv_bfloat16* out_a;                   // [B][HW][C0+C1]
    __nv_bfloat16* out_raw;                 // optional bf16(x), same shape
    int B, items_per_image;                 // persistent step kernel: work items = B x items_per_image blocks of pix_per_block pixels
    int precise;                            // 1: operands are (hi | lo) pairs -- out rows are 2 (C0+C1) wide, low halves C0+C1 elements behind
    const DropSpec* drop;                   // training-mode forward only: Dropout after the SiLU (unet.py:86); nullptr otherwise
    float* save_mr;                         // training-mode forward only: [B][groups][2] (mean, rstd) kept for the backward; nullptr otherwise
};

// Per-(image, channel) scale / shift of a GroupNorm from the fp64 channel sums: y = x * sc[c] + sh[c].
// sc / sh: [C] floats each, CONTIGUOUS ([2C] floats = [C] doubles of scratch) in shared memory; gm / gr: [groups] each.
// All channel sums are fetched in ONE parallel round trip (a thread per channel), then reduced per group from shared memory.
// Ends with a __syncthreads().
__device__ __forceinline__ void groupnorm_scale_shift(const PrepParams& p, int b, float* sc, float* sh, float* gm, float* gr) {
    const int C = p.C0 + p.C1;
    const int gs = C / p.groups;
    // The [2C]-float sc/sh area holds C doubles: first all the sums, then (after the group means are known) all the sums of squares.  The two
    // rounds re-read the statistics from L2 instead of caching (sum, sumsq) pairs in registers: the streaming loop that follows decides the
    // kernel's register count, and 40 extra registers here cost the 128x128-level launches a third of their resident blocks (24 vs 15 us).
    double* scratch = reinterpret_cast<double*>(sc);
    const double inv = 1.0 / (static_cast<double>(gs) * static_cast<double>(p.HW));
    for (int c = threadIdx.x; c < C; c += blockDim.x)
        scratch[c] = (c < p.C0) ? __ldcg(p.st0 + (static_cast<long long>(b) * p.C0 + c) * 2) : __ldcg(p.st1 + (static_cast<long long>(b) * p.C1 + (c - p.C0)) * 2);
    __syncthreads();
    double gmean = 0.0;
    for (int g = threadIdx.x; g < p.groups; g += blockDim.x) {        // groups <= blockDim.x: one iteration
        double s = 0.0;
        for (int j = 0; j < gs; ++j) s += scratch[g * gs + j];
        gmean = s * inv;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x)
        scratch[c] = (c < p.C0) ? __ldcg(p.st0 + (static_cast<long long>(b) * p.C0 + c) * 2 + 1) : __ldcg(p.st1 + (static_cast<long long>(b) * p.C1 + (c - p.C0)) * 2 + 1);
    __syncthreads();
    for (int g = threadIdx.x; g < p.groups; g += blockDim.x) {
        double q = 0.0;
        for (int j = 0; j < gs; ++j) q += scratch[g * gs + j];
        double var = q * inv - gmean * gmean;               // fp64: no cancellation problem for |mean| >> std
        if (var < 0.0) var = 0.0;
        gm[g] = static_cast<float>(gmean);
        gr[g] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.eps)));
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / gs;
        const float k = gr[g] * __ldg(&p.gamma[c]);
        sc[c] = k; sh[c] = __ldg(&p.beta[c]) - gm[g] * k;
    }
    __syncthreads();
}

// Block size = (C/4) * k threads: every thread owns ONE 4-channel column for the whole kernel (scale / shift live in
// registers, no shared-memory or integer-division traffic in the streaming loop) and walks pixels k at a time.
// DROP: training-mode forward of a block2 (Dropout after the SiLU); a separate instantiation keeps the Philox code out of the sampler's kernel
template <bool DROP>
__global__ void __launch_bounds__(512) prep_kernel(const PrepParams p) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float sm[];
    const int C = p.C0 + p.C1;
    float* sc = sm;              // [C] scale
    float* sh = sm + C;          // [C] shift
    float* gm = sm + 2 * C;      // [groups] mean
    float* gr = gm + p.groups;   // [groups] rstd
    const int b = blockIdx.y;
    groupnorm_scale_shift(p, b, sc, sh, gm, gr);
    if (p.save_mr != nullptr && blockIdx.x == 0) {
        for (int g = threadIdx.x; g < p.groups; g += blockDim.x) {
            p.save_mr[(static_cast<long long>(b) * p.groups + g) * 2] = gm[g];
            p.save_mr[(static_cast<long long>(b) * p.groups + g) * 2 + 1] = gr[g];
        }
    }
    const int vpp = C >> 2;                       // 4-channel vectors per pixel
    const int kpix = blockDim.x / vpp;            // pixels covered by the block per step
    const int c = (threadIdx.x % vpp) << 2;       // this thread's channels (constant)
    const int lp = threadIdx.x / vpp;             // this thread's pixel lane
    float k4[4], s4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { k4[j] = sc[c + j]; s4[j] = sh[c + j]; }
    const bool from0 = c < p.C0;
    const float* src = from0 ? p.src0 + c : p.src1 + (c - p.C0);
    const int cs = from0 ? p.C0 : p.C1;
    const int pix0 = blockIdx.x * p.pix_per_block;
    const int pix1 = min(pix0 + p.pix_per_block, p.HW);
    const long long img = static_cast<long long>(b) * p.HW;
    constexpr int U = 4;                          // independent 16-byte loads in flight per thread
    for (int pix = pix0 + lp; pix < pix1; pix += kpix * U) {
        float4 x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int pp = pix + u * kpix;
            if (pp < pix1) x[u] = __ldg(reinterpret_cast<const float4*>(src + (img + pp) * cs));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int pp = pix + u * kpix;
            if (pp < pix1) {
                float y0 = x[u].x * k4[0] + s4[0], y1 = x[u].y * k4[1] + s4[1], y2 = x[u].z * k4[2] + s4[2], y3 = x[u].w * k4[3] + s4[3];
                if (p.silu) { y0 = silu_f(y0); y1 = silu_f(y1); y2 = silu_f(y2); y3 = silu_f(y3); }
                if (DROP && p.drop->p > 0.f) {
                    float ds[4];
                    drop_scale4(*p.drop, b, c, pp, C, p.HW, ds);
                    y0 *= ds[0]; y1 *= ds[1]; y2 *= ds[2]; y3 *= ds[3];
                }
                const long long o = (img + pp) * (p.precise ? 2 * C : C) + c;
                const long long lo_off = p.precise ? C : 0;
                store_operand4(p.out_a, o, lo_off, y0, y1, y2, y3);
                if (p.out_raw) store_operand4(p.out_raw, o, lo_off, x[u].x, x[u].y, x[u].z, x[u].w);
            }
        }
    }
}

// fp32 NHWC -> bf16 NHWC, optionally nearest-2x upsampled (nn.Upsample(scale_factor=2,'nearest'), unet.py:58-65)
__global__ void __launch_bounds__(256) cast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int B, int H, int W,
                                                   int C, int up) {
    pdl_launch_dependents();
    pdl_wait();
    const int OH = H * up, OW = W * up;
    const long long total = static_cast<long long>(B) * OH * OW * (C >> 2);
    const int vpp = C >> 2;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int v = static_cast<int>(i % vpp);
        long long pix = i / vpp;
        const int ow = static_cast<int>(pix % OW); pix /= OW;
        const int oh = static_cast<int>(pix % OH);
        const int b = static_cast<int>(pix / OH);
        const long long sp = (static_cast<long long>(b) * H + oh / up) * W + ow / up;
        const float4 x = __ldg(reinterpret_cast<const float4*>(src + sp * C + (v << 2)));
        *reinterpret_cast<uint2*>(dst + i * 4) = pack_bf16x4(x.x, x.y, x.z, x.w);
    }
}

// Row softmax over keys (torch.softmax(attn, -1), unet.py:136): S fp32 [rows][L] -> P bf16 [rows][L].
// Rows are grouped in segments of `seg` tokens; a row only attends to the keys of its own segment (two 64-token images share
// one 128-row attention batch); keys outside get probability 0.
__global__ void __launch_bounds__(256) softmax_kernel(const float* __restrict__ S, __nv_bfloat16* __restrict__ P, long long rows, int L,
                                                      int seg, int precise) {
    pdl_launch_dependents();
    pdl_wait();
    const long long row = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const int r_in = static_cast<int>(row % L);
    const int k0 = (r_in / seg) * seg;
    const float* s = S + row * L;
    __nv_bfloat16* pr = P + row * (precise ? 2 * L : L);
    float m = -INFINITY;
    for (int k = k0 + lane; k < k0 + seg; k += 32) m = fmaxf(m, s[k]);
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int k = k0 + lane; k < k0 + seg; k += 32) sum += expf(s[k] - m);
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    for (int k = lane; k < L; k += 32) {
        const float v = (k >= k0 && k < k0 + seg) ? expf(s[k] - m) * inv : 0.f;
        pr[k] = __float2bfloat16_rn(v);
        if (precise) pr[L + k] = __float2bfloat16_rn(bf16_residual(v));
    }
}

// Start of a step: clear the GroupNorm statistics arena and advance the device-side timestep.
__global__ void __launch_bounds__(256) step_begin_kernel(float4* stats, long long n4, StepCtl* ctl) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x)
        stats[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const int t = ctl->t_next;
        ctl->t_cur = t;
        ctl->t_next = t - 1;
    }
}

// PositionalEncoding + noise_level_mlp (unet.py:18-31, 177-184): one block per image -> tau[b][inner].
struct EmbedParams {
    const StepCtl* ctl;
    const float* nl_table;   // fp32(sqrt_alphas_cumprod_prev) [T+1]
    const float* nl_buf;     // [B]
    const float* w1; const float* b1;   // [4*inner][inner]
    const float* w2; const float* b2;   // [inner][4*inner]
    float* tau;              // [B][inner]
    int inner;
};
__global__ void __launch_bounds__(256) embed_kernel(const EmbedParams p) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float sm[];
    const int inner = p.inner, hid = 4 * inner;
    float* pe = sm;            // [inner]
    float* h = sm + inner;     // [hid]
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const float nl = p.ctl->nl_from_table ? p.nl_table[p.ctl->t_cur + 1] : p.nl_buf[b];
    const int count = inner / 2;
    for (int j = threadIdx.x; j < inner; j += blockDim.x) {
        const int jj = j < count ? j : j - count;
        const float step = static_cast<float>(jj) / static_cast<float>(count);
        const float e = nl * expf(-9.210340371976184f * step);
        pe[j] = j < count ? sinf(e) : cosf(e);
    }
    __syncthreads();
    // one warp per output row (coalesced weights + shuffle reduction); 8 rows are in flight per warp so the L2 latency of the
    // weight loads is paid once per batch of rows, not once per row
    constexpr int R = 8;
    for (int j0 = warp * R; j0 < hid; j0 += nwarp * R) {
        float a[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            a[r] = 0.f;
            if (j0 + r < hid)
                for (int i = lane; i < inner; i += 32) a[r] += __ldg(&p.w1[(j0 + r) * inner + i]) * pe[i];
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int o = 16; o; o >>= 1) a[r] += __shfl_xor_sync(0xffffffffu, a[r], o);
            if (lane == 0 && j0 + r < hid) { const float t = a[r] + p.b1[j0 + r]; h[j0 + r] = t / (1.0f + expf(-t)); }
        }
    }
    __syncthreads();
    for (int j0 = warp * R; j0 < inner; j0 += nwarp * R) {
        float a[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            a[r] = 0.f;
            if (j0 + r < inner)
                for (int i = lane; i < hid; i += 32) a[r] += __ldg(&p.w2[(j0 + r) * hid + i]) * h[i];
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int o = 16; o; o >>= 1) a[r] += __shfl_xor_sync(0xffffffffu, a[r], o);
            if (lane == 0 && j0 + r < inner) p.tau[b * inner + j0 + r] = a[r] + p.b2[j0 + r];
        }
    }
}

// All FeatureWiseAffine projections at once (unet.py:34-50, bias-only form) + the block1 conv bias folded in:
// film[b][j] = Wf[j] . tau[b] + bf[j] + cbias[j],  j over the concatenated Cout of every ResnetBlock.
// A block owns 64 outputs: their weight rows are staged (coalesced) in padded smem together with tau of every image.
__global__ void __launch_bounds__(256) film_kernel(const float* __restrict__ wf, const float* __restrict__ bf, const float* __restrict__ cbias,
                                                   const float* __restrict__ tau, float* __restrict__ film, int F, int inner, int B) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float sm[];
    float* ws = sm;                          // [64][inner + 1]
    float* ts = sm + 64 * (inner + 1);       // [B][inner]
    const int j0 = blockIdx.x * 64;
    for (int i = threadIdx.x; i < 64 * inner; i += blockDim.x) {
        const int r = i / inner, c = i % inner;
        ws[r * (inner + 1) + c] = (j0 + r < F) ? __ldg(&wf[static_cast<long long>(j0 + r) * inner + c]) : 0.f;
    }
    for (int i = threadIdx.x; i < B * inner; i += blockDim.x) ts[i] = __ldg(&tau[i]);
    __syncthreads();
    const int jl = threadIdx.x & 63, bq = threadIdx.x >> 6;
    const int j = j0 + jl;
    if (j >= F) return;
    const float base = bf[j] + cbias[j];
    const float* w = ws + jl * (inner + 1);
    for (int b = bq; b < B; b += 4) {
        float a = base;
        const float* t = ts + b * inner;
        for (int i = 0; i < inner; ++i) a += w[i] * t[i];
        film[static_cast<long long>(b) * F + j] = a;
    }
}

// OIHW fp32 conv weight -> K-major bf16 GEMM operand: dst[o][k_off + (r*KW+s)*cin_pad + c]
// precise mode (lo_off != 0): the row is [hi (ktot) | lo (ktot)], lo_off = ktot; `ld` = row length in elements.
// One thread per (o, c): it reads the KH*KW contiguous taps of that pair (the warp reads one contiguous span) and writes tap by tap (for a
// fixed tap the warp's c are contiguous in dst): both sides coalesced.  Runs after every optimizer step of the training loop.
__global__ void __launch_bounds__(256) pack_conv_weight_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int Cout, int Cin,
                                                               int KH, int KW, int ld, int k_off, int cin_pad, int lo_off) {
    const long long total = static_cast<long long>(Cout) * Cin;
    const int taps = KH * KW;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % Cin);
        const int o = static_cast<int>(i / Cin);
        const float* sp = src + i * taps;
        for (int t = 0; t < taps; ++t) {
            const float v = sp[t];
            const long long di = static_cast<long long>(o) * ld + k_off + t * cin_pad + c;
            dst[di] = __float2bfloat16_rn(v);
            if (lo_off) dst[di + lo_off] = __float2bfloat16_rn(bf16_residual(v));
        }
    }
}

// Weights of the four phase convs of a folded (nearest-2x -> conv3x3): for output parity py the kernel rows that land on low-res
// row offset a are R(0,0)={0}, R(0,1)={1,2}, R(1,0)={0,1}, R(1,1)={2} (same for columns); summed in fp32, rounded once to bf16.
// dst[phase][o][(a*2+b)*Cin + c]
__global__ void __launch_bounds__(256) fold_upsample_weight_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ d0,
                                                                   __nv_bfloat16* __restrict__ d1, __nv_bfloat16* __restrict__ d2,
                                                                   __nv_bfloat16* __restrict__ d3, int Cout, int Cin, int ld, int lo_off) {
    const long long total = 4LL * Cout * Cin * 4;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        long long r = i;
        const int c = static_cast<int>(r % Cin); r /= Cin;
        const int ab = static_cast<int>(r % 4); r /= 4;
        const int o = static_cast<int>(r % Cout);
        const int ph = static_cast<int>(r / Cout);
        const int py = ph >> 1, px = ph & 1, a = ab >> 1, b = ab & 1;
        const int r0 = (py == 0) ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2), r1 = (py == 0) ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2);
        const int s0 = (px == 0) ? (b == 0 ? 0 : 1) : (b == 0 ? 0 : 2), s1 = (px == 0) ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : 2);
        float acc = 0.f;
        for (int rr = r0; rr <= r1; ++rr)
            for (int ss = s0; ss <= s1; ++ss) acc += src[((static_cast<long long>(o) * Cin + c) * 3 + rr) * 3 + ss];
        __nv_bfloat16* d = ph == 0 ? d0 : (ph == 1 ? d1 : (ph == 2 ? d2 : d3));
        const long long di = static_cast<long long>(o) * ld + ab * Cin + c;
        d[di] = __float2bfloat16_rn(acc);
        if (lo_off) d[di + lo_off] = __float2bfloat16_rn(bf16_residual(acc));
    }
}

__global__ void add_vec_kernel(const float* a, const float* b, float* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + (b ? b[i] : 0.f);
}

// API boundary: NCHW fp32 (reference layout) -> bf16 NHWC channel slice of the UNet input buffer (+ optional fp32 NCHW copy).
__global__ void __launch_bounds__(256) load_nchw_kernel(const float* __restrict__ src, int B, int C, int H, int W, __nv_bfloat16* __restrict__ in_buf,
                                                        int in_C, int coff, float* __restrict__ copy, int lo_off) {
    const long long total = static_cast<long long>(B) * C * H * W;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        long long r = i;
        const int w = static_cast<int>(r % W); r /= W;
        const int h = static_cast<int>(r % H); r /= H;
        const int c = static_cast<int>(r % C);
        const int b = static_cast<int>(r / C);
        const float v = src[i];
        const long long di = ((static_cast<long long>(b) * H + h) * W + w) * in_C + coff + c;
        in_buf[di] = __float2bfloat16_rn(v);
        if (lo_off) in_buf[di + lo_off] = __float2bfloat16_rn(bf16_residual(v));
        if (copy) copy[i] = v;
    }
}

// q_sample (diffusion.py:212-219) fused with the input load: x_noisy = g * x0 + sqrt(1 - g^2) * noise, written as bf16 into the
// UNet input buffer (channels [coff, coff+C)); g = continuous_sqrt_alpha_cumprod of the image.
__global__ void __launch_bounds__(256) q_sample_load_kernel(const float* __restrict__ x0, const float* __restrict__ noise, const float* __restrict__ gamma,
                                                            int B, int C, int H, int W, __nv_bfloat16* __restrict__ in_buf, int in_C, int coff, int lo_off) {
    const long long total = static_cast<long long>(B) * C * H * W;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        long long r = i;
        const int w = static_cast<int>(r % W); r /= W;
        const int h = static_cast<int>(r % H); r /= H;
        const int c = static_cast<int>(r % C);
        const int b = static_cast<int>(r / C);
        const float g = gamma[b];
        const float v = __fadd_rn(__fmul_rn(g, x0[i]), __fmul_rn(sqrtf(__fsub_rn(1.0f, __fmul_rn(g, g))), noise[i]));
        const long long di = ((static_cast<long long>(b) * H + h) * W + w) * in_C + coff + c;
        in_buf[di] = __float2bfloat16_rn(v);
        if (lo_off) in_buf[di + lo_off] = __float2bfloat16_rn(bf16_residual(v));
    }
}

// L1Loss / MSELoss with reduction='sum' (diffusion.py:84-90, 245): loss += sum |noise - eps| (or squared), double accumulation
__global__ void __launch_bounds__(256) loss_sum_kernel(const float* __restrict__ noise, const float* __restrict__ eps, long long n, int l2,
                                                       double* __restrict__ out) {
    double acc = 0.0;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float d = noise[i] - eps[i];
        acc += l2 ? static_cast<double>(d) * d : static_cast<double>(fabsf(d));
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += ws[i];
        atomicAdd(out, t);
    }
}

// tensor2img of the reference (core/metrics.py:8-34) on the device: clamp to [lo, hi] -> (x - lo) / (hi - lo) -> * 255 -> round half to even
// -> uint8, laid out HWC.  `n` images [n][C][H][W] are tiled like torchvision.utils.make_grid(nrow, padding=2, pad_value=0) when n > 1
// (grid of ncol x nrow cells of (H+2) x (W+2) pixels plus a 2-pixel border); n == 1: plain [H][W][C].
__global__ void __launch_bounds__(256) tensor2img_kernel(const float* __restrict__ src, unsigned char* __restrict__ dst, int n, int C, int H, int W,
                                                         int ncol, int GH, int GW, float lo, float hi) {
    const long long total = static_cast<long long>(GH) * GW * C;
    const int pad = n > 1 ? 2 : 0;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        const int gx = static_cast<int>((i / C) % GW);
        const int gy = static_cast<int>(i / (static_cast<long long>(C) * GW));
        float v = 0.f;                                               // pad_value 0 (already in [0, 1] units)
        bool inside = true;
        int img = 0, y = gy, x = gx;
        if (n > 1) {
            const int cy = (gy - pad) / (H + pad), cx = (gx - pad) / (W + pad);
            y = (gy - pad) - cy * (H + pad); x = (gx - pad) - cx * (W + pad);
            img = cy * ncol + cx;
            inside = gy >= pad && gx >= pad && y < H && x < W && img < n && cx < ncol;
        }
        if (inside) {
            float t = src[((static_cast<long long>(img) * C + c) * H + y) * W + x];
            t = fminf(fmaxf(t, lo), hi);
            v = __fdiv_rn(__fsub_rn(t, lo), __fsub_rn(hi, lo));
        }
        dst[i] = static_cast<unsigned char>(rintf(__fmul_rn(v, 255.0f)));
    }
}

// sum of squared differences of two uint8 images (calculate_psnr, core/metrics.py:42-50): exact in integers
__global__ void __launch_bounds__(256) ssd_u8_kernel(const unsigned char* __restrict__ a, const unsigned char* __restrict__ b, long long n,
                                                     unsigned long long* __restrict__ out) {
    unsigned long long acc = 0;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int d = static_cast<int>(a[i]) - static_cast<int>(b[i]);
        acc += static_cast<unsigned long long>(d * d);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

// One pass of Pillow's 8-bit resampler (src/libImaging/Resample.c, ImagingResampleHorizontal_8bpc / Vertical_8bpc) with host-built
// integer coefficient tables: out = clip8((2^21 + sum_k in[first + k] * coef[k]) >> 22).  `axis_stride` / `line_stride` / `img_stride` are in
// elements of the uint8 input [img][line][axis][C]; one thread per output element (img, line, o, c).
// Output: uint8 [img][..][C] with the same structure (dst_u8) and / or float CHW planes (dst_f32: ToTensor /255, optional flip along W,
// * (max - min) + min -- data/util.py:74-83) -- only meaningful for the LAST pass (vertical), where line = x and o = y.
__global__ void __launch_bounds__(256) resample_u8_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst_u8, float* __restrict__ dst_f32,
                                                          int n_img, int n_line, int n_out, int C, long long in_axis_stride, long long in_line_stride,
                                                          long long in_img_stride, long long out_axis_stride, long long out_line_stride, long long out_img_stride,
                                                          const int* __restrict__ bounds, const int* __restrict__ coef, int ksize, int flip, float vmin, float vmax) {
    const long long total = static_cast<long long>(n_img) * n_line * n_out * C;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        const int line = static_cast<int>((i / C) % n_line);
        const int o = static_cast<int>((i / (static_cast<long long>(C) * n_line)) % n_out);
        const int img = static_cast<int>(i / (static_cast<long long>(C) * n_line * n_out));
        const int first = bounds[2 * o], n = bounds[2 * o + 1];
        const unsigned char* sp = src + img * in_img_stride + line * in_line_stride + first * in_axis_stride + c;
        int acc = 1 << 21;
        for (int k = 0; k < n; ++k) acc += static_cast<int>(sp[k * in_axis_stride]) * coef[o * ksize + k];
        int v = acc >> 22;
        v = v < 0 ? 0 : (v > 255 ? 255 : v);
        if (dst_u8) dst_u8[img * out_img_stride + line * out_line_stride + o * out_axis_stride + c] = static_cast<unsigned char>(v);
        if (dst_f32) {                                 // last (vertical) pass: o = y, line = x; CHW planes of n_out x n_line
            const int x = flip ? (n_line - 1 - line) : line;
            const float t = __fdiv_rn(static_cast<float>(v), 255.0f);
            dst_f32[((static_cast<long long>(img) * C + c) * n_out + o) * n_line + x] = __fadd_rn(__fmul_rn(t, __fsub_rn(vmax, vmin)), vmin);
        }
    }
}

}  // namespace sr3
