// attn_tcgen05.cuh -- the core of SelfAttention (reference model/sr3_modules/unet.py:129-139) as ONE tensor-core kernel:
//
//   S = q k^T / sqrt(C)   (tcgen05.mma, accumulator in TMEM)   -> row softmax in registers (unnormalised exp, bf16) -> shared memory
//   O = P v               (tcgen05.mma, P is the A operand straight from shared memory)                -> O / rowsum -> bf16 [token][C]
//
// One CTA = (attention batch z, 128 query rows, DN of the C output channels); it recomputes S for its query rows (0.27 GFLOP per
// image at 16x16: cheaper than a second launch) so that 2 * C/DN CTAs per image run instead of 2.  n_head = 1 in every reference
// config, so the head dimension is C (512): S needs all of it (K loop over C), O = 128 x C fp32 would not fit TMEM next to S, hence
// the split of O's columns across CTAs.  Key count Lt <= 256 (16x16 = 256 tokens; two 8x8 images share a 128-token batch with a
// block-diagonal mask, as in softmax_kernel); longer sequences (32x32 mid block of the 64->512 config) keep the three-launch path.
//
// Operands (both produced by tile-kernel launches): qk [nz*Lt][2C] bf16 (q = columns [0,C), k = [C,2C)), vT [nz*C][Lt] bf16.
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2..5 = softmax / epilogue (one TMEM lane quadrant each).
#pragma once
#include <cuda_bf16.h>
#include <cuda.h>
#include "gemm_tcgen05.cuh"

namespace sr3 {

constexpr int ATTN_THREADS = 192;
constexpr int ATTN_STAGES = 3;
constexpr int ATTN_STAGE_BYTES = 16384 + 32768;          // A: 128 rows x 64 | B: up to 256 rows x 64 (bf16, 128B-swizzled)
constexpr int ATTN_P_BYTES = 65536;                      // P: 128 rows x up to 256 keys, as K chunks of 64
constexpr int ATTN_SMEM_BYTES = 1024 + GEMM_HDR_BYTES + ATTN_STAGES * ATTN_STAGE_BYTES + ATTN_P_BYTES;   // header: barriers + TMEM slot
constexpr uint32_t ATTN_O_COL = 256;                     // TMEM: S in columns [0, Lt), O in [256, 256 + DN)

struct AttnParams {
    CUtensorMap qk_map;      // 2-D bf16 [nz*Lt rows][2C], box {64, 128}
    CUtensorMap vt_map;      // 2-D bf16 [nz*C rows][Lt], box {64, 128}
    __nv_bfloat16* out;      // [nz*Lt][C]
    int C, Lt, HW, dn, nz;
    float scale_log2e;       // log2(e) / sqrt(C)
};

// One (attention batch z, query tile qt, channel slice dc) unit.  MEGA = false: the body of attn_kernel (one unit per CTA, the kernel owns
// barriers and TMEM).  MEGA = true: called by step_kernel for every unit of this CTA; `pm` is the global-memory copy of the parameters
// (TMA descriptors), TMEM was allocated by the caller; warps >= 6 only take part in the block-wide barriers.
template <bool MEGA>
__device__ __forceinline__ void attn_unit(const AttnParams& p, const AttnParams* pm, const uint32_t base_in, uint8_t* base_ptr_in,
                                          const uint32_t tmem_base_in, const int qt, const int dc, const int z) {
    const uint32_t bar_base = base_in;                                 // header (gemm_tcgen05.cuh)
    const uint32_t base = base_in + GEMM_HDR_BYTES;
    uint8_t* base_ptr = base_ptr_in + GEMM_HDR_BYTES;
    const uint32_t p_base = base + ATTN_STAGES * ATTN_STAGE_BYTES;
    uint8_t* p_ptr = base_ptr + ATTN_STAGES * ATTN_STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (ATTN_STAGES + s); };
    const uint32_t s_full = bar_base + 8u * (2 * ATTN_STAGES);
    const uint32_t p_ready = s_full + 8u;
    const uint32_t o_full = s_full + 16u;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr_in + HDR_TMEM_SLOT);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int kc1 = p.C / 64;        // K chunks of S = q k^T
    const int kc3 = p.Lt / 64;       // K chunks of O = P v

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&pm->qk_map);
        tma_prefetch_desc(&pm->vt_map);
        if constexpr (MEGA) {
            for (int i = 0; i < HDR_NUM_BARS; ++i) mbar_inval(bar_base + 8u * i);
        }
        for (int s = 0; s < ATTN_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(s_full, 1);
        mbar_init(p_ready, 4);       // one arrive per softmax warp
        mbar_init(o_full, 1);
        fence_mbar_init();
    }
    if constexpr (!MEGA) {
        if (warp == 1) {
            tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = MEGA ? tmem_base_in : *tmem_slot;
    if constexpr (!MEGA) {
        pdl_launch_dependents();
        pdl_wait();                  // q, k, vT come from the two preceding launches
    }

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < kc1 + kc3; ++it) {
            mbar_wait(empty_bar(s), ph ^ 1u, 11);
            if (elect_one_sync()) {
                const uint32_t dst = base + s * ATTN_STAGE_BYTES;
                if (it < kc1) {
                    mbar_arrive_expect_tx(full_bar(s), 16384 + p.Lt * 128);
                    tma_load_2d(dst, &pm->qk_map, full_bar(s), it * 64, z * p.Lt + qt * 128);
                    for (int j = 0; j < p.Lt / 128; ++j)
                        tma_load_2d(dst + 16384 + j * 16384, &pm->qk_map, full_bar(s), p.C + it * 64, z * p.Lt + j * 128);
                } else {
                    mbar_arrive_expect_tx(full_bar(s), p.dn * 128);
                    for (int j = 0; j < p.dn / 128; ++j)
                        tma_load_2d(dst + 16384 + j * 16384, &pm->vt_map, full_bar(s), (it - kc1) * 64, z * p.C + dc * p.dn + j * 128);
                }
            }
            __syncwarp();
            if (++s == ATTN_STAGES) { s = 0; ph ^= 1u; }
        }
        if constexpr (MEGA) {        // tail: no commit arrival may be in flight when the barriers are recycled
            const int total = kc1 + kc3, n_wait = total < ATTN_STAGES ? total : ATTN_STAGES;
            for (int i = 0; i < n_wait; ++i) {
                mbar_wait(empty_bar(s), ph ^ 1u, 16);
                if (++s == ATTN_STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        const uint64_t desc_hi = umma_desc_kmajor_sw128(0, 1024) & 0xFFFFFFFF00000000ull;
        const uint32_t desc_lo0 = static_cast<uint32_t>(umma_desc_kmajor_sw128(base, 1024) & 0xFFFFFFFFull);
        const uint32_t idesc_s = umma_idesc_bf16(128, p.Lt);
        const uint32_t idesc_o = umma_idesc_bf16(128, p.dn);
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < kc1 + kc3; ++it) {
            if (it == kc1) {                               // P is in shared memory (generic-proxy writes fenced by the softmax warps)
                mbar_wait(p_ready, 0, 12);
                tc_fence_after();
            }
            mbar_wait(full_bar(s), ph, 13);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint32_t st_lo = desc_lo0 + ((s * ATTN_STAGE_BYTES) >> 4);
                const uint64_t bdesc = desc_hi | (st_lo + (16384 >> 4));
                if (it < kc1) {
                    const uint64_t adesc = desc_hi | st_lo;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tmem_base, adesc + 2 * kk, bdesc + 2 * kk, idesc_s, (it | kk) != 0);
                } else {
                    const int kc = it - kc1;
                    const uint64_t adesc = desc_hi | (desc_lo0 + ((ATTN_STAGES * ATTN_STAGE_BYTES + kc * 16384) >> 4));
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tmem_base + ATTN_O_COL, adesc + 2 * kk, bdesc + 2 * kk, idesc_o, (kc | kk) != 0);
                }
                umma_commit(empty_bar(s));
                if (it == kc1 - 1) umma_commit(s_full);
                if (it == kc1 + kc3 - 1) umma_commit(o_full);
            }
            __syncwarp();
            if (++s == ATTN_STAGES) { s = 0; ph ^= 1u; }
        }
    } else if (warp < 6) {
        // ------------------------------------------------------------ softmax + epilogue: thread = one query row
        const int q = warp & 3;                            // TMEM lane quadrant this warp may access
        const int row = q * 32 + lane;
        const int g = qt * 128 + row;                      // token index inside the attention batch
        const int seg = g / p.HW;                          // image inside the batch (block-diagonal mask)
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const int nch = p.Lt / 32;
        mbar_wait(s_full, 0, 14);
        tc_fence_after();
        float mx = -3.0e38f;
#pragma unroll 1
        for (int ch = 0; ch < nch; ++ch) {
            uint32_t v[32];
            tmem_ld_32x32(t_row + ch * 32, v);
            tmem_ld_wait();
            if ((ch * 32) / p.HW == seg && (ch * 32 + 31) / p.HW == seg) {
#pragma unroll
                for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if ((ch * 32 + j) / p.HW == seg) mx = fmaxf(mx, __uint_as_float(v[j]));
            }
        }
        float sum = 0.f;
        const float mxs = mx * p.scale_log2e;
        uint8_t* prow = p_ptr + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll 1
        for (int ch = 0; ch < nch; ++ch) {
            uint32_t v[32];
            tmem_ld_32x32(t_row + ch * 32, v);
            tmem_ld_wait();
            float e[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const bool ok = (ch * 32 + j) / p.HW == seg;
                const float x = exp2f(fmaf(__uint_as_float(v[j]), p.scale_log2e, -mxs));
                e[j] = ok ? x : 0.f;
                sum += e[j];
            }
            // 32 keys = four 16-byte units of the K chunk (64 keys, 128 B per row); units are XOR-swizzled with the row (128B swizzle)
            uint8_t* pc = prow + (ch >> 1) * 16384;
            const int u0 = (ch & 1) * 4;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                __nv_bfloat162 h0 = __floats2bfloat162_rn(e[8 * u], e[8 * u + 1]);
                __nv_bfloat162 h1 = __floats2bfloat162_rn(e[8 * u + 2], e[8 * u + 3]);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(e[8 * u + 4], e[8 * u + 5]);
                __nv_bfloat162 h3 = __floats2bfloat162_rn(e[8 * u + 6], e[8 * u + 7]);
                uint4 w;
                w.x = *reinterpret_cast<uint32_t*>(&h0); w.y = *reinterpret_cast<uint32_t*>(&h1);
                w.z = *reinterpret_cast<uint32_t*>(&h2); w.w = *reinterpret_cast<uint32_t*>(&h3);
                *reinterpret_cast<uint4*>(pc + (((u0 + u) ^ (row & 7)) << 4)) = w;
            }
        }
        fence_proxy_async_smem();                          // P was written through the generic proxy, the MMA reads it through the async proxy
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready);

        mbar_wait(o_full, 0, 15);
        tc_fence_after();
        const float inv = 1.0f / sum;
        __nv_bfloat16* orow = p.out + (static_cast<long long>(z) * p.Lt + g) * p.C + dc * p.dn;
#pragma unroll 1
        for (int ch = 0; ch < p.dn / 32; ++ch) {
            uint32_t v[32];
            tmem_ld_32x32(t_row + ATTN_O_COL + ch * 32, v);
            tmem_ld_wait();
            uint4* o4 = reinterpret_cast<uint4*>(orow + ch * 32);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(v[8 * u]) * inv, __uint_as_float(v[8 * u + 1]) * inv);
                __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(v[8 * u + 2]) * inv, __uint_as_float(v[8 * u + 3]) * inv);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(v[8 * u + 4]) * inv, __uint_as_float(v[8 * u + 5]) * inv);
                __nv_bfloat162 h3 = __floats2bfloat162_rn(__uint_as_float(v[8 * u + 6]) * inv, __uint_as_float(v[8 * u + 7]) * inv);
                uint4 w;
                w.x = *reinterpret_cast<uint32_t*>(&h0); w.y = *reinterpret_cast<uint32_t*>(&h1);
                w.z = *reinterpret_cast<uint32_t*>(&h2); w.w = *reinterpret_cast<uint32_t*>(&h3);
                o4[u] = w;
            }
        }
    }
    if constexpr (MEGA) fence_proxy_async_all();      // the next unit's TMA loads overwrite shared memory this unit wrote generically
    tc_fence_before();
    __syncthreads();
    if constexpr (!MEGA) {
        if (warp == 1) tmem_dealloc(tmem_base, 512);
    }
}

__global__ void __launch_bounds__(ATTN_THREADS, 1) attn_kernel(const __grid_constant__ AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const int n_dc = p.C / p.dn;
    attn_unit<false>(p, &p, base, smem_raw + (base - raw), 0u, blockIdx.x / n_dc, blockIdx.x % n_dc, blockIdx.y);
}

}  // namespace sr3
