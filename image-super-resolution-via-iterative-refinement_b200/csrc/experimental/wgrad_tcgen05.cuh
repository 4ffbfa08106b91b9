// wgrad_tcgen05.cuh -- EXPERIMENTAL (training row; not on any product path yet, not yet run on a B200):
// weight gradient of a stride-1 conv3x3,
//
//     dW[co][tap][ci] = sum over pixels p of  dY[p][co] * X[p + tap][ci],
//
// as a tcgen05 GEMM whose contraction runs over PIXELS.  With NHWC activations both operands are MN-major (channels contiguous):
// a TMA box {64 channels, 8 x 8 pixels} with the 128-byte swizzle IS the canonical MN-major layout of the UMMA shared-memory
// descriptor -- in 16-byte units  Swizzle<3,4,3> o ((8,n),(8,k)) : ((1,LBO),(8,SBO))  (CUTLASS cute/atom/mma_traits_sm100.hpp,
// make_umma_desc<Major::MN>): one 128 B row = 64 channels of one pixel, 8 pixels = one 1024 B atom (SBO), the next 64-channel
// panel LBO bytes further; instruction-descriptor bits 15 / 16 select MN-major A / B.  A tap is the X box shifted by (dh, dw) with
// TMA zero fill, exactly as in the forward kernel.
//
// One CTA = (128 output channels, 64 input channels, one kernel row dh, a slice of the batch): three accumulators (dw = -1, 0, +1)
// of 128 x 64 fp32 in TMEM; partial sums of different batch slices meet through fp32 atomics in a pre-zeroed dW (a tuned version
// would use the deterministic split-K reduction of gemm_tcgen05.cuh).
#pragma once
#include <cuda_bf16.h>
#include <cuda.h>
#include "ptx.cuh"

namespace sr3 {

constexpr int WGRAD_THREADS = 192;
constexpr int WGRAD_STAGES = 4;
constexpr int WGRAD_STAGE_BYTES = 16384 + 3 * 8192;     // dY: 2 panels of 64 co x 64 px | X: 3 taps x (64 px x 64 ci)
constexpr int WGRAD_SMEM_BYTES = 1024 + WGRAD_STAGES * WGRAD_STAGE_BYTES + 256;

struct WgradParams {
    CUtensorMap dy_map;      // 5-D bf16 (Cout, W, 1, H, B), box {64, 8, 1, 8, 1}
    CUtensorMap x_map;       // 5-D bf16 (Cin,  W, 1, H, B), box {64, 8, 1, 8, 1}
    float* dw;               // [Cout][9][Cin], pre-zeroed
    int Cin, Cout, H, W, B;
    int b_per_cta;           // images per batch slice
};

// MN-major operand, 128-byte swizzle: [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
    d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int m, int n) {      // as umma_idesc_bf16, both operands MN-major
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

__global__ void __launch_bounds__(WGRAD_THREADS, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - raw);
    const uint32_t bar_base = base + WGRAD_STAGES * WGRAD_STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (WGRAD_STAGES + s); };
    const uint32_t acc_full = bar_base + 8u * (2 * WGRAD_STAGES);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + WGRAD_STAGES * WGRAD_STAGE_BYTES + 8 * (2 * WGRAD_STAGES + 1));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_ci = p.Cin / 64;
    const int co0 = (blockIdx.x / n_ci) * 128, ci0 = (blockIdx.x % n_ci) * 64;
    const int dh = static_cast<int>(blockIdx.y) - 1;
    const int b_begin = blockIdx.z * p.b_per_cta;
    const int b_end = min(p.B, b_begin + p.b_per_cta);
    const int tiles_w = p.W / 8, tiles_h = p.H / 8;
    const int iters = (b_end - b_begin) * tiles_w * tiles_h;          // K chunks of 64 pixels

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.dy_map);
        tma_prefetch_desc(&p.x_map);
        for (int s = 0; s < WGRAD_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(acc_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < iters; ++it) {
            const int b = b_begin + it / (tiles_w * tiles_h);
            const int r = it % (tiles_w * tiles_h);
            const int w0 = (r % tiles_w) * 8, h0 = (r / tiles_w) * 8;
            mbar_wait(empty_bar(s), ph ^ 1u, 21);
            if (elect_one_sync()) {
                const uint32_t dst = base + s * WGRAD_STAGE_BYTES;
                mbar_arrive_expect_tx(full_bar(s), WGRAD_STAGE_BYTES);
                tma_load_5d(dst, &p.dy_map, full_bar(s), co0, w0, 0, h0, b);
                tma_load_5d(dst + 8192, &p.dy_map, full_bar(s), co0 + 64, w0, 0, h0, b);
                for (int t = 0; t < 3; ++t)      // X shifted by (dh, dw = t - 1); out-of-image pixels arrive as zeros (= padding)
                    tma_load_5d(dst + 16384 + t * 8192, &p.x_map, full_bar(s), ci0, w0 + t - 1, 0, h0 + dh, b);
            }
            __syncwarp();
            if (++s == WGRAD_STAGES) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 1) {
        constexpr uint32_t IDESC = umma_idesc_bf16_mn(128, 64);
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(full_bar(s), ph, 22);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint32_t st = base + s * WGRAD_STAGE_BYTES;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {                       // 16 pixels = two 8-pixel atoms per UMMA
                    const uint64_t adesc = umma_desc_mnmajor_sw128(st + kk * 2048, 8192, 1024);
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        const uint64_t bdesc = umma_desc_mnmajor_sw128(st + 16384 + t * 8192 + kk * 2048, 8192, 1024);
                        umma_bf16_ss(tmem_base + t * 64, adesc, bdesc, IDESC, (it | kk) != 0);
                    }
                }
                umma_commit(empty_bar(s));
                if (it == iters - 1) umma_commit(acc_full);
            }
            __syncwarp();
            if (++s == WGRAD_STAGES) { s = 0; ph ^= 1u; }
        }
    } else if (iters > 0) {
        const int q = warp & 3;
        const int co = co0 + q * 32 + lane;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        mbar_wait(acc_full, 0, 23);
        tc_fence_after();
#pragma unroll 1
        for (int t = 0; t < 3; ++t) {
#pragma unroll 1
            for (int ch = 0; ch < 2; ++ch) {
                uint32_t v[32];
                tmem_ld_32x32(t_row + t * 64 + ch * 32, v);
                tmem_ld_wait();
                if (co < p.Cout) {
                    float* dst = p.dw + (static_cast<long long>(co) * 9 + (dh + 1) * 3 + t) * p.Cin + ci0 + ch * 32;
#pragma unroll
                    for (int j = 0; j < 32; ++j) atomicAdd(dst + j, __uint_as_float(v[j]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 256);
}

}  // namespace sr3
