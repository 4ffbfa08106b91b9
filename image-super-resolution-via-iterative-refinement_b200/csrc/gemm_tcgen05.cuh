// Implicit-GEMM tile kernel for sm_100a (tcgen05 + TMEM + TMA), the one tensor-core kernel of the SR3 path.
//
//   D[(MH x 128) pixels x BLOCK_N] = sum over stages, taps   A_tap[128 x 64] * B_tap[BLOCK_N x 64]^T     (bf16 -> fp32 in TMEM)
//
// * A pipeline stage holds ONE activation tile fetched by one 5-D TMA box from an NHWC bf16 tensor (out-of-image pixels are
//   zero-filled by TMA = conv padding; no im2col buffer exists) plus 1..3 weight tiles.  For a 3x3 stride-1 conv the
//   box is a "tall halo": 8 pixels wide, (rows + 2) high, shifted horizontally by the tap column dw.  The three vertical
//   taps of that column are then just three UMMA descriptors whose start address is shifted by whole 8-pixel rows
//   (1024 B = one 128-byte-swizzle atom), so every activation byte that enters shared memory feeds 3 taps x BLOCK_N outputs,
//   and with MH = 2 two 128-row accumulators share the halo and the weights: the SM ingest rate (~64 B/clk), not the
//   tensor pipe, is what bounds a 128x64 or 128x128 tile fed tap by tap.
// * The generic form (MH = 1, one tap per stage, box = any 128 rows of a 5-D view) covers stride-2 convs (row / column parity
//   folded into the view), 1x1 convs, the 8x8 levels and the attention matrix products.
// * B rows are output channels (or keys / head-dim for attention) of a K-major bf16 matrix, 2-D TMA box {64, BLOCK_N}.
// * The stage table is built on the host (engine.cu): 3x3 taps, the 1x1 residual conv accumulated into the same TMEM
//   tile and channel concats are all just more stages.
// * Persistent CTAs (one per SM) walk a contiguous range of tiles; accumulators are double buffered in TMEM so the epilogue
//   of tile i overlaps the MMAs of tile i+1.  Warp 0 = TMA producer, warp 1 = TMEM owner + tcgen05.mma issuer (converged
//   warps, one elected lane issues), warps 2..5 = epilogue: tcgen05.ld -> bias / FiLM -> residual (TMA prefetched into smem)
//   -> fp32 tile staged in swizzled smem and written by TMA store (and/or bf16 direct stores) -> per-(image, channel)
//   GroupNorm partial sums by a transposing warp-shuffle reduction, accumulated in registers across tiles.
// * The last UNet conv (Cout = 3) uses the posterior epilogue: eps -> x0 -> clamp -> posterior mean -> + sigma_t * z
//   (model/sr3_modules/diffusion.py:141-174 of the reference) and writes x_{t-1} straight into the next step's input.
#pragma once
#include <cuda.h>
#include "ptx.cuh"

namespace sr3 {

struct OutSpec {
    long long sZ, sB, sH, sW, off;   // element strides: gemm-batch, image, row, column; constant offset
};

// Device-resident control block: everything that changes between launches of the (captured) step graph.
struct StepCtl {
    int t_cur;            // timestep of the running step
    int t_next;           // timestep the next step_begin will load
    int nl_from_table;    // 1: noise level = sqrt_alphas_cumprod_prev[t+1] (sampling); 0: read nl_buf[b]
    int out_mode;         // 0: write eps (UNet.forward); 1: posterior update (p_sample)
    int use_noise_buf;    // 1: z from noise_buf; 0: Philox
    int write_mean;       // 1: also store the posterior mean (p_mean_variance)
    int clip;             // clip_denoised
    int update_state;     // 1: write x_{t-1} into x_state and the UNet input buffer
    unsigned long long seed;
    unsigned long long sample_offset;   // global index of image 0 (multi-GPU sharding keeps streams rank independent)
};

struct PostParams {
    const float* tab;         // [5][T]: sqrt_recip_ac, sqrt_recipm1_ac, post_coef1, post_coef2, post_logvar_clipped
    int T;
    int H, W, C;              // image geometry, C = sample channels (3)
    float* x_state;           // [B,C,H,W] fp32 NCHW
    float* eps_out;           // [B,C,H,W]
    float* mean_out;          // [B,C,H,W]
    const float* noise_buf;   // [B,C,H,W]
    __nv_bfloat16* in_buf;    // NHWC bf16 UNet input, channel stride in_C, x_t lives at channels [in_coff, in_coff+C)
    int in_C, in_coff;
    int in_lo_off;            // precise mode: the low halves (x - bf16(x)) live in_lo_off channels further; 0 = bf16 mode
};

// One pipeline stage of the K loop (host-built table, copied to shared memory by the kernel): up to three K slabs
// (64 channels each).  Tall-halo stages load ONE activation box that serves all taps (a_multi = 0); generic stages load one
// box per slab (a_multi = 1) -- grouping slabs only amortises the mbarrier handshake.
struct StageTap {
    int a_chan;      // channel coordinate (dim 0) of the A box
    int dw, dh, p;   // box shift along W', H' and the parity coordinate
    int b_col;       // B column (K coordinate)
    int a_off;       // byte offset of this tap's first row inside the A stage buffer (multiple of 1024)
};
struct StageDesc {
    int ntaps;       // K slabs in this stage (1..3)
    int a_multi;     // 1: every tap has its own A box (loaded at a_off), 0: one box (tap 0's coordinates) shared by all taps
    int a_sel;       // which A tensor map
    int pad0;
    StageTap tap[3];
    int pad1, pad2;  // 96 bytes = 6 x int4
};

struct GemmParams {
    CUtensorMap a_map[2];
    CUtensorMap b_map;
    CUtensorMap out_map;     // fp32 output, 5-D {N, W, 1, H, B|Z}, box {32, w_sub, 1, h_sub, 1}: one warp's 32 rows x 32 columns
    CUtensorMap res_map;     // fp32 residual, same geometry
    const StageDesc* ktab;   // [num_k]
    int num_k;
    int a_stage_bytes;       // bytes reserved for activation boxes per stage (multiple of 1024)
    int a_box_bytes;         // bytes of ONE activation box (what a single TMA load delivers)
    int a_half_off;          // byte offset between the two 128-row halves inside the A buffer (MH = 2)
    int b_taps;              // weight tiles reserved per stage (max ntaps)
    int tiles_w, tiles_h, tiles_b;
    int n_tiles, nz;         // persistent schedule: tile = m_tile + tiles_m * (n_tile + n_tiles * z); CTA c owns a contiguous range
    int tma_epi;             // 1: fp32 output / residual go through smem + TMA (out_map / res_map)
    int epi_c4_is_z;         // 5th coordinate of out_map / res_map: gemm-batch z (1) or image index (0)
    int ksplit;              // split-K factor: ksplit CTAs share one output tile, partial sums meet in `ws` (fp32, same addressing as out_f32)
    float* ws;               // zero between launches (the finalising CTA clears what it reads)
    unsigned int* counters;  // [tiles] arrival counters (monotonic: +ksplit per launch)
    const void* pf_ptr;      // weights of the NEXT tile-kernel launch: pulled into L2 while this launch runs (they would otherwise be
    long long pf_bytes;      // first-touch HBM reads on the critical path of every pipeline stage of that launch)
    int dbg;                 // SR3_DBG bit mask (timing experiments only): 1 skip epilogue body, 2 skip stats, 4 skip out store,
                             // 8 skip A loads, 16 skip B loads, 32 skip MMAs
    int w_box, h_box, b_box; // pixel patch of one tile: w_box * h_box * b_box == MH * 128 (all powers of two)
    int w_shift, h_shift;    // log2(w_box), log2(h_box): the epilogue splits a tile row into (w, h, image) with shifts, not divisions
    int a_zstep, b_zrows;
    int stages;
    // epilogue
    int mode;                // 0 normal, 1 final conv (eps / posterior)
    int OW, OH, OB, n_valid;
    float scale;
    const float* bias;
    const float* bias2;      // per-image bias (FiLM + conv bias), bias2[img * bias2_stride + n]
    int bias2_stride;
    const float* resid;
    OutSpec rs;
    float* out_f32;
    OutSpec os;
    __nv_bfloat16* out_bf16;
    OutSpec hs;
    // columns >= t_col0 are stored TRANSPOSED instead: out_t[(img / t_per) * t_rows + (n - t_col0)][(img % t_per) * OH*OW + oh*OW + ow]
    // (the v third of the qkv projection lands directly as the K-major B operand v^T of O = P v)
    __nv_bfloat16* out_t;
    int t_col0, t_rows, t_ld, t_per;
    double* stats;           // [B][stats_C][2] (sum, sumsq) in fp64, channel offset stats_coff (see stats_quantize)
    int stats_C, stats_coff;
    const StepCtl* ctl;
    PostParams post;
    // z_phase = 1: the gemm-batch index z = 2*py + px is the output phase of a folded (nearest-2x -> conv3x3): taps shift by (py, px) input
    // pixels, weights of phase z start at row z * b_zrows, the tile lands at output pixel (2h + py, 2w + px) -- out_map is then
    // {2N (px, n), W, 2 (py), H, B} and plain stores add (py * z_off_hi + px * z_off_lo) elements
    int z_phase;
    long long z_off_hi, z_off_lo;
    // Precise mode (fp32-level accuracy on the bf16 tensor cores): every operand is a pair hi = bf16(x), lo = bf16(x - hi) and a product is
    // hi*hi + hi*lo + lo*hi (the dropped lo*lo term is ~2^-18 relative).  The K loop runs `passes` = 3 times over the SAME stage table:
    // pass 1 reads the low weights (B column + lo_b_col), pass 2 the low activations (A channel + lo_a_chan[source]).  Low halves of
    // bf16 outputs go lo_out_off elements (lo_t_off for the transposed store) behind the high halves.  passes = 1: plain bf16.
    int passes, lo_b_col, lo_a_chan[2];
    long long lo_out_off, lo_t_off;
    int t_fixed;             // >= 0: timestep of the running step (persistent step kernel: ctl->t_cur is not used there); -1: read ctl->t_cur
};

constexpr int GEMM_THREADS = 320;           // warp 0 producer, warp 1 MMA, warps 2..9 epilogue (two groups of four)
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_MAX_STAGES = 8;
// per epilogue warp: 4 KB output staging (fp32 chunk for the TMA store) [+ 2 x 4 KB residual staging when the layer has one]
__host__ __device__ constexpr int gemm_epi_warp_bytes(bool resid) { return resid ? 12288 : 4096; }
__host__ __device__ constexpr int gemm_epi_bytes(bool resid) { return GEMM_EPI_WARPS * gemm_epi_warp_bytes(resid); }
constexpr int GEMM_MAX_K = 160;              // stages per tile (<= 3 K slabs each); the table is sized per launch
// Shared-memory header (first 2 KB of the 1024-aligned region, same place for every op of the persistent step kernel):
//   [0, 512) mbarriers | [512, 516) TMEM base slot | [520, 528) step scalars | [768, 2048) parameter block of the running op
constexpr int GEMM_HDR_BYTES = 2048;
constexpr int HDR_TMEM_SLOT = 512;
constexpr int HDR_PARAMS = 768;
constexpr int HDR_NUM_BARS = 64;
__host__ __device__ constexpr int gemm_aux_bytes(int num_k) {
    return GEMM_HDR_BYTES + ((num_k * 96 + 127) / 128) * 128 /*stage table*/ + GEMM_EPI_WARPS * 2 * 32 * 4 /*per-warp bias staging*/;
}

__host__ __device__ constexpr int gemm_stage_bytes(int block_n, int a_stage_bytes, int b_taps) { return a_stage_bytes + b_taps * block_n * 128; }
__host__ __device__ constexpr int gemm_smem_bytes(int block_n, int a_stage_bytes, int b_taps, int stages, bool resid, int num_k) {
    return stages * gemm_stage_bytes(block_n, a_stage_bytes, b_taps) + gemm_epi_bytes(resid) + 1024 /*align slack*/ + gemm_aux_bytes(num_k);
}

// ---------------------------------------------------------------- Philox4x32-10 + Box-Muller
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
    const float u1 = (static_cast<float>(a) + 1.0f) * 2.3283064365386963e-10f;   // (0, 1]
    const float u2 = static_cast<float>(b) * 2.3283064365386963e-10f;            // [0, 1)
    const float r = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    z0 = r * c; z1 = r * s;
}

// GroupNorm statistics are accumulated with fp64 atomics.  Every contribution is first rounded to a multiple of 2^-20, so as long as
// a sum stays below 2^33 (|x|_rms < 180 over a 512x512 channel) every addition is EXACT: the result does not depend on the order in
// which the CTAs arrive -- repeat runs are bit identical -- and E[x^2] - mean^2 is evaluated in fp64 by the consumer (no fp32
// cancellation for |mean| >> std).  The rounding itself is unbiased and ~1e-6 absolute per contribution, far below eps = 1e-5.
__device__ __forceinline__ double stats_quantize(double v) { return rint(v * 1048576.0) * (1.0 / 1048576.0); }

// Transposing warp reduction: every lane holds v[0..31] (one row, 32 columns); afterwards lane l returns the sum over the
// 32 lanes (rows) of column l.  31 shuffles instead of 32 x 5.
__device__ __forceinline__ float warp_column_sums(float (&v)[32]) {
    const uint32_t lane = lane_id();
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int j = 0; j < half; ++j) {
            const float keep = up ? v[j + half] : v[j];
            const float send = up ? v[j] : v[j + half];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}

__device__ __forceinline__ long long out_index(const OutSpec& s, int z, int img, int oh, int ow) {
    return s.off + static_cast<long long>(z) * s.sZ + static_cast<long long>(img) * s.sB + static_cast<long long>(oh) * s.sH +
           static_cast<long long>(ow) * s.sW;
}

// reference: predict_start_from_noise + clamp + q_posterior + p_sample noise add (diffusion.py:141-174)
__device__ __forceinline__ void final_epilogue(const GemmParams& p, const float (&eps)[4], int img, int oh, int ow) {
    const PostParams& q = p.post;
    const StepCtl ctl = *p.ctl;
    const long long plane = static_cast<long long>(q.H) * q.W;
    const long long pix = static_cast<long long>(oh) * q.W + ow;
    if (ctl.out_mode == 0) {
        for (int c = 0; c < q.C; ++c) q.eps_out[(static_cast<long long>(img) * q.C + c) * plane + pix] = eps[c];
        return;
    }
    const int t = p.t_fixed >= 0 ? p.t_fixed : ctl.t_cur;
    const float c1 = q.tab[t], c2 = q.tab[q.T + t], pc1 = q.tab[2 * q.T + t], pc2 = q.tab[3 * q.T + t];
    const float sigma = (t > 0) ? expf(0.5f * q.tab[4 * q.T + t]) : 0.0f;
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    if (t > 0) {
        if (ctl.use_noise_buf) {
            for (int c = 0; c < q.C; ++c) z[c] = q.noise_buf[(static_cast<long long>(img) * q.C + c) * plane + pix];
        } else {
            uint32_t ctr[4] = {static_cast<uint32_t>(pix), static_cast<uint32_t>(ctl.sample_offset + img), static_cast<uint32_t>(t),
                               static_cast<uint32_t>((ctl.sample_offset + img) >> 32)};
            philox4x32_10(ctr, static_cast<uint32_t>(ctl.seed), static_cast<uint32_t>(ctl.seed >> 32));
            box_muller(ctr[0], ctr[1], z[0], z[1]);
            box_muller(ctr[2], ctr[3], z[2], z[3]);
        }
    }
    for (int c = 0; c < q.C; ++c) {
        const long long idx = (static_cast<long long>(img) * q.C + c) * plane + pix;
        const float xt = q.x_state[idx];
        float x0 = __fsub_rn(__fmul_rn(c1, xt), __fmul_rn(c2, eps[c]));
        if (ctl.clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
        const float mean = __fadd_rn(__fmul_rn(pc1, x0), __fmul_rn(pc2, xt));
        if (ctl.write_mean) q.mean_out[idx] = mean;
        const float xn = __fadd_rn(mean, __fmul_rn(z[c], sigma));
        if (ctl.update_state) {
            q.x_state[idx] = xn;
            __nv_bfloat16* ib = q.in_buf + (static_cast<long long>(img) * plane + pix) * q.in_C + q.in_coff + c;
            const __nv_bfloat16 hi = __float2bfloat16_rn(xn);
            *ib = hi;
            if (q.in_lo_off) ib[q.in_lo_off] = __float2bfloat16_rn(xn - __bfloat162float(hi));
        }
    }
}

static_assert(sizeof(GemmParams) <= GEMM_HDR_BYTES - HDR_PARAMS, "GemmParams must fit the shared-memory header");

// CTA-local set-up of one tile-kernel op: stage table -> shared memory, TMA descriptor prefetch, mbarrier (re-)initialisation.  Nothing here
// depends on data produced by other CTAs, so the persistent step kernel runs it BEFORE waiting at the grid barrier in front of the op.
// Must be followed by a block-wide barrier.
__device__ __forceinline__ void gemm_stage_setup(const GemmParams& p, const GemmParams* pm, const uint32_t base, uint8_t* base_ptr, const int block_n,
                                                 const bool recycle) {
    const int stages = p.stages;
    const int stage_bytes = p.a_stage_bytes + p.b_taps * block_n * 128;
    const bool use_res_tma = p.tma_epi && p.resid != nullptr && p.ksplit <= 1;
    const int epi_bytes = GEMM_EPI_WARPS * gemm_epi_warp_bytes(use_res_tma);
    uint8_t* aux_ptr = base_ptr + GEMM_HDR_BYTES + stages * stage_bytes + epi_bytes;
    {
        const int4* src = reinterpret_cast<const int4*>(p.ktab);
        int4* dst = reinterpret_cast<int4*>(aux_ptr);
        for (int i = threadIdx.x; i < p.num_k * 6; i += blockDim.x) dst[i] = __ldg(&src[i]);
    }
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&pm->a_map[0]);
        tma_prefetch_desc(&pm->a_map[1]);
        tma_prefetch_desc(&pm->b_map);
        if (p.tma_epi) { tma_prefetch_desc(&pm->out_map); tma_prefetch_desc(&pm->res_map); }
        const uint32_t bar_base = base;
        if (recycle) {                           // the previous op's barriers (all quiescent: see the end of gemm_tile_body) are recycled
            for (int i = 0; i < HDR_NUM_BARS; ++i) mbar_inval(bar_base + 8u * i);
        }
        for (int s = 0; s < stages; ++s) {
            mbar_init(bar_base + 8u * s, 1);                                     // full
            mbar_init(bar_base + 8u * (GEMM_MAX_STAGES + s), 1);                 // empty
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_base + 8u * (2 * GEMM_MAX_STAGES + a), 1);             // tmem full
            mbar_init(bar_base + 8u * (2 * GEMM_MAX_STAGES + 2 + a), GEMM_EPI_WARPS);   // tmem empty: one arrive per epilogue warp
        }
        for (int w = 0; w < 2 * GEMM_EPI_WARPS; ++w) mbar_init(bar_base + 8u * (2 * GEMM_MAX_STAGES + 4 + w), 1);   // residual tiles
        fence_mbar_init();
    }
}

// Persistent, warp-specialised tile loop (see the file header).  Two callers:
//   MEGA = false: gemm_tile_kernel, one launch per layer; `p` lives in the kernel parameter space, barriers / TMEM are set up here.
//   MEGA = true : step_kernel (step_megakernel.cuh), the whole reverse step in ONE cooperative launch; `p` is the shared-memory copy
//                 of the op's parameter block, `pm` its global-memory original (TMA descriptors must not live in shared memory),
//                 TMEM was allocated once by the caller, `cta / ncta` replace blockIdx / gridDim.
// `base` / `base_ptr`: 1024-aligned start of the CTA's dynamic shared memory (header first).
template <int BLOCK_N, int MH, bool MEGA>
__device__ __forceinline__ void gemm_tile_body(const GemmParams& p, const GemmParams* pm, const uint32_t base, uint8_t* base_ptr,
                                               const uint32_t tmem_base_in, const int cta, const int ncta) {
    constexpr int B_BYTES = BLOCK_N * 128;
    constexpr uint32_t ACC_COLS = MH * BLOCK_N;                              // columns per accumulator buffer
    constexpr uint32_t TMEM_COLS = (2 * ACC_COLS) < 32 ? 32 : 2 * ACC_COLS;  // two buffers, power of two >= 32
    static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM budget");
    constexpr uint32_t IDESC = umma_idesc_bf16(128, BLOCK_N);

    const int stages = p.stages;
    const int stage_bytes = p.a_stage_bytes + p.b_taps * B_BYTES;            // multiple of 1024
    const bool use_res_tma = p.tma_epi && p.resid != nullptr && p.ksplit <= 1;
    const int epi_warp_bytes = gemm_epi_warp_bytes(use_res_tma);
    const int epi_bytes = GEMM_EPI_WARPS * epi_warp_bytes;
    const uint32_t bar_base = base;                                          // header
    const uint32_t stage_base = base + GEMM_HDR_BYTES;
    uint8_t* stage_ptr = base_ptr + GEMM_HDR_BYTES;
    const uint32_t epi_base = stage_base + stages * stage_bytes;
    // barriers: full[8] empty[8] tmem_full[2] tmem_empty[2] res_full[4 warps]; then the TMEM slot
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (GEMM_MAX_STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * GEMM_MAX_STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * GEMM_MAX_STAGES + 2 + a); };
    auto res_bar = [&](int w, int b) { return bar_base + 8u * (2 * GEMM_MAX_STAGES + 4 + 2 * w + b); };   // w in [0, 8)
    uint8_t* aux_ptr = stage_ptr + stages * stage_bytes + epi_bytes;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + HDR_TMEM_SLOT);
    // with ~227 KB of shared memory per CTA there is no L1 left: anything re-read per iteration must live in smem
    StageDesc* ktab_s = reinterpret_cast<StageDesc*>(aux_ptr);
    float* bias_s = reinterpret_cast<float*>(aux_ptr + ((p.num_k * 96 + 127) / 128) * 128);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_b;
    const int ksplit = p.ksplit > 1 ? p.ksplit : 1;
    const int total_tiles = tiles_m * p.n_tiles * p.nz * ksplit;      // split index fastest: the CTAs of one output tile run together
    const int num_kt = p.num_k * (p.passes > 1 ? p.passes : 1);       // stages per tile (precise mode: three passes over the table)
    if constexpr (!MEGA) gemm_stage_setup(p, pm, base, base_ptr, BLOCK_N, false);     // (the step kernel did this before its grid barrier)
    if constexpr (!MEGA) {
        if (warp == 1) {
            tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), TMEM_COLS);
            tmem_relinquish();
        }
    }
    if constexpr (!MEGA) {
        tc_fence_before();
        __syncthreads();
    }
    tc_fence_after();
    const uint32_t tmem_base = MEGA ? tmem_base_in : *tmem_slot;
    if (warp == 2 && lane == 0 && p.pf_bytes > 0) {                   // L2 prefetch of this CTA's slice of the next layer's weights
        long long chunk = ((p.pf_bytes + ncta - 1) / ncta + 15) & ~15ll;
        const long long off = chunk * cta;
        if (off < p.pf_bytes) {
            if (off + chunk > p.pf_bytes) chunk = (p.pf_bytes - off) & ~15ll;
            const char* src = static_cast<const char*>(p.pf_ptr) + off;
            for (long long done = 0; done < chunk; done += 65536) {
                const unsigned int n = static_cast<unsigned int>(chunk - done < 65536 ? chunk - done : 65536);
                if (n >= 16) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + done), "r"(n) : "memory");
            }
        }
    }
    if constexpr (!MEGA) {
        pdl_launch_dependents();  // the next kernel may be scheduled onto SMs as our CTAs retire ...
        pdl_wait();               // ... and we touch upstream activations / statistics only after the previous kernel completed
    }

    auto decode = [&](int tile_s, int& w0, int& h0, int& b0, int& n0, int& z) {
        const int tile = tile_s / ksplit;
        const int tm = tile % tiles_m;          // m fastest: a CTA's consecutive tiles share the weight slab and mostly the image
        const int r = tile / tiles_m;
        const int nt = r % p.n_tiles;
        z = r / p.n_tiles;
        const int tw = tm % p.tiles_w;
        const int th = (tm / p.tiles_w) % p.tiles_h;
        const int tb = tm / (p.tiles_w * p.tiles_h);
        w0 = tw * p.w_box; h0 = th * p.h_box; b0 = tb * p.b_box + z * p.a_zstep; n0 = nt * BLOCK_N;
    };
    // contiguous tile range of this CTA (balanced to +-1 tile)
    const int tile_begin = static_cast<int>((static_cast<long long>(total_tiles) * cta) / ncta);
    const int tile_end = static_cast<int>((static_cast<long long>(total_tiles) * (cta + 1)) / ncta);

    if (warp == 0) {
        // ---------------------------------------------------- TMA producer warp (converged; one elected lane issues)
        int s = 0;
        uint32_t ph = 0;
        for (int tile = tile_begin; tile < tile_end; ++tile) {
            int w0, h0, b0, n0, z;
            decode(tile, w0, h0, b0, n0, z);
            const int brow = n0 + z * p.b_zrows;
            const int zdw = p.z_phase ? (z & 1) : 0, zdh = p.z_phase ? (z >> 1) : 0;
            const int sp = tile % ksplit;
            const int k0 = (num_kt * sp) / ksplit, k1 = (num_kt * (sp + 1)) / ksplit;
            for (int k = k0; k < k1; ++k) {
                mbar_wait(empty_bar(s), ph ^ 1u, 1);
                if (elect_one_sync()) {
                    const int pass = k / p.num_k;
                    const StageDesc& e = ktab_s[k - pass * p.num_k];
                    const int a_lo = (pass == 2) ? p.lo_a_chan[e.a_sel] : 0, b_lo = (pass == 1) ? p.lo_b_col : 0;
                    const uint32_t a_dst = stage_base + s * stage_bytes;
                    const int na = e.a_multi ? e.ntaps : 1;
                    mbar_arrive_expect_tx(full_bar(s), ((p.dbg & 8) ? 0 : na * p.a_box_bytes) + ((p.dbg & 16) ? 0 : e.ntaps * B_BYTES));
                    if (!(p.dbg & 8)) {
                        for (int t = 0; t < na; ++t)
                            tma_load_5d(a_dst + (e.a_multi ? e.tap[t].a_off : 0), &pm->a_map[e.a_sel], full_bar(s), e.tap[t].a_chan + a_lo, w0 + e.tap[t].dw + zdw, e.tap[t].p,
                                        h0 + e.tap[t].dh + (e.a_multi ? zdh : 0), b0);     // tall halo boxes always start one row above the tile
                    }
                    if (!(p.dbg & 16)) {
                        for (int t = 0; t < e.ntaps; ++t)
                            tma_load_2d(a_dst + p.a_stage_bytes + t * B_BYTES, &pm->b_map, full_bar(s), e.tap[t].b_col + b_lo, brow);
                    }
                }
                __syncwarp();
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
        }
        if constexpr (MEGA) {
            // tail: every stage this CTA filled has been released by the MMA warp's commit -> no arrival is in flight when the
            // barriers are recycled by the next op
            int filled = 0;
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                const int sp = tile % ksplit;
                filled += (num_kt * (sp + 1)) / ksplit - (num_kt * sp) / ksplit;
            }
            const int n_wait = filled < stages ? filled : stages;
            for (int i = 0; i < n_wait; ++i) {
                mbar_wait(empty_bar(s), ph ^ 1u, 6);
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------- MMA issuer warp (converged; one elected lane issues)
        // descriptor low word = start address >> 4 (+ fixed LBO); byte offsets are added as (bytes >> 4)
        const uint64_t desc_hi = umma_desc_kmajor_sw128(0, 1024) & 0xFFFFFFFF00000000ull;
        const uint32_t desc_lo0 = static_cast<uint32_t>(umma_desc_kmajor_sw128(stage_base, 1024) & 0xFFFFFFFFull);
        int s = 0, ti = 0;
        uint32_t ph = 0;
        for (int tile = tile_begin; tile < tile_end; ++tile, ++ti) {
            const int acc = ti & 1;
            mbar_wait(tempty_bar(acc), ((ti >> 1) & 1) ^ 1u, 4);        // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
            int zrow_off = 0;                                           // folded-upsample phase py: vertical taps one halo row further down
            if (p.z_phase) {
                const int zz = (tile / ksplit) / (tiles_m * p.n_tiles);
                zrow_off = (zz >> 1) * 1024;
            }
            const int sp = tile % ksplit;
            const int k0 = (num_kt * sp) / ksplit, k1 = (num_kt * (sp + 1)) / ksplit;
            for (int k = k0; k < k1; ++k) {
                mbar_wait(full_bar(s), ph, 2);
                tc_fence_after();
                if (elect_one_sync()) {
                    const StageDesc& e = ktab_s[k % p.num_k];
                    if (!(p.dbg & 32)) {
                        const uint32_t a_lo = desc_lo0 + ((s * stage_bytes) >> 4);
                        const uint32_t b_lo = desc_lo0 + ((s * stage_bytes + p.a_stage_bytes) >> 4);
#pragma unroll
                        for (int half = 0; half < MH; ++half) {
                            for (int t = 0; t < e.ntaps; ++t) {
                                const uint64_t adesc = desc_hi | (a_lo + ((half * p.a_half_off + e.tap[t].a_off + (e.a_multi ? 0 : zrow_off)) >> 4));
                                const uint64_t bdesc = desc_hi | (b_lo + ((t * B_BYTES) >> 4));
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)   // 4 x UMMA_K(16) = 64 channels; +32 B inside the 128 B swizzle row
                                    umma_bf16_ss(d_tmem + half * BLOCK_N, adesc + 2 * kk, bdesc + 2 * kk, IDESC, ((k - k0) | t | kk) != 0);
                            }
                        }
                    }
                    umma_commit(empty_bar(s));
                    if (k == k1 - 1) umma_commit(tfull_bar(acc));
                }
                __syncwarp();
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // ---------------------------------------------------- epilogue: 2 groups x 4 warps; a warp owns one TMEM lane quadrant and
        // every second 32-column chunk of it, so eight warps hide each other's tcgen05.ld / shuffle / fence latencies
        const int q = warp & 3;
        const int ew = warp - 2;                                            // 0..7
        const int grp = ew >> 2;
        const uint32_t out_smem = epi_base + ew * epi_warp_bytes;           // 4 KB
        const uint32_t res_smem = out_smem + 4096;                          // 2 x 4 KB (layers with a residual only)
        uint8_t* out_ptr = stage_ptr + stages * stage_bytes + ew * epi_warp_bytes;
        uint8_t* res_ptr = out_ptr + 4096;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const bool use_out_tma = p.tma_epi && p.out_f32 != nullptr;
        uint32_t res_phase = 0;        // bit b = parity of res_bar(q, b)
        uint32_t res_count = 0;        // residual chunks consumed so far
        bool out_pending = false;      // a bulk store from the staging buffer may still be reading it
        // GroupNorm partial sums of this lane's column, kept in registers across tiles of the same (image, column block)
        constexpr int NCHS = BLOCK_N >= 32 ? BLOCK_N / 32 : 1;
        double st_sum[NCHS], st_sq[NCHS];
#pragma unroll
        for (int c = 0; c < NCHS; ++c) { st_sum[c] = 0.0; st_sq[c] = 0.0; }
        int st_img = -1, st_n0 = -1;
        auto flush_stats = [&]() {
            if (st_img >= 0 && st_img < p.OB) {
#pragma unroll
                for (int c = 0; c < NCHS; ++c) {
                    const int n = st_n0 + c * 32 + lane;
                    if (n < p.n_valid) {
                        double* st = p.stats + (static_cast<long long>(st_img) * p.stats_C + p.stats_coff + n) * 2;
                        red_add_f64_global(st, stats_quantize(st_sum[c]));
                        red_add_f64_global(st + 1, stats_quantize(st_sq[c]));
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < NCHS; ++c) { st_sum[c] = 0.0; st_sq[c] = 0.0; }
        };
        constexpr int NCH = BLOCK_N >= 32 ? BLOCK_N / 32 : 1;      // 32-column chunks per 128-row half
        constexpr int NITEMS = MH * NCH;                            // work items (half, chunk) per tile; this warp takes item % 2 == grp
        int ti = 0;
        for (int tile = tile_begin; tile < tile_end; ++tile, ++ti) {
            int w0, h0, b0, n0, z;
            decode(tile, w0, h0, b0, n0, z);
            const int acc = ti & 1;
            // geometry of a work item: rows of `half`, columns of `ch`
            auto item_geom = [&](int item, int qq, int& half, int& ch, int& sw, int& sh, int& c4) {
                half = item / NCH; ch = item % NCH;
                const int r0 = half * 128 + qq * 32;
                sw = r0 & (p.w_box - 1);
                sh = (r0 >> p.w_shift) & (p.h_box - 1);
                const int sb = r0 >> (p.w_shift + p.h_shift);
                c4 = p.epi_c4_is_z ? z : (b0 + sb);
            };
            auto request_resid = [&](int item) {                   // lane 0 only
                int half, ch, sw, sh, c4;
                item_geom(item, q, half, ch, sw, sh, c4);
                const uint32_t b = res_count & 1;
                mbar_arrive_expect_tx(res_bar(ew, b), 4096);
                tma_load_5d(res_smem + b * 4096, &pm->res_map, res_bar(ew, b), n0 + ch * 32, w0 + sw, 0, h0 + sh, c4);
            };
            const bool has_work = grp < NITEMS;
            if (use_res_tma && has_work && lane == 0) request_resid(grp);     // overlaps the main loop
            // bias (+ per-image FiLM bias) of this lane's column for every work item of this warp: fetched now (there is no L1, an
            // L2 round trip per item would sit on the critical path), broadcast through smem when the item is processed
            constexpr int MAXI = (NITEMS + 1) / 2;
            float bvs[MAXI];
            if constexpr (BLOCK_N != 16) {
#pragma unroll
                for (int ii = 0; ii < MAXI; ++ii) {
                    const int item = grp + 2 * ii;
                    float bv = 0.f;
                    if (item < NITEMS) {
                        const int half = item / NCH, ch = item % NCH;
                        const int bb = (half * 128 + q * 32) >> (p.w_shift + p.h_shift);
                        const int img = b0 + bb;
                        const int n = n0 + ch * 32 + lane;
                        if (n < p.n_valid) {
                            if (p.bias) bv += __ldg(&p.bias[n]);
                            if (p.bias2) bv += __ldcg(&p.bias2[static_cast<long long>(img < p.OB ? img : 0) * p.bias2_stride + n]);   // written earlier in the same launch (step kernel): L2 only
                        }
                    }
                    bvs[ii] = bv;
                }
            }
            mbar_wait(tfull_bar(acc), (ti >> 1) & 1, 3);
            tc_fence_after();
            if (!has_work) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(acc));
            }
            // pass 0 reads the accumulator from TMEM.  With split-K it only stores this CTA's partial tile into its slice of `ws`;
            // once all `ksplit` CTAs of the output tile have arrived at the tile's counter, pass 1 runs in EVERY one of them on a
            // 1/ksplit share of the tile's 32x32 units: sum the partials (fixed order: deterministic) and do the real epilogue.
            // (All CTAs of the grid are resident -- at most one (tile, split) pair per SM -- so the spin wait cannot deadlock.)
            const int npass = ksplit > 1 ? 2 : 1;
            const int sp = tile % ksplit;
            constexpr long long SLICE = static_cast<long long>(MH) * 128 * BLOCK_N;       // floats per partial tile
#pragma unroll 1
            for (int pass = 0; pass < npass; ++pass) {
            if (pass == 1) {
                __threadfence();
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (ew == 0 && lane == 0) {
                    unsigned int* ctr = p.counters + tile / ksplit;
                    const unsigned int old = atomicAdd(ctr, 1u);
                    const unsigned int target = (old / ksplit + 1u) * ksplit;              // the counter only ever grows
                    uint64_t t0 = 0;
                    for (uint32_t spins = 0;; ++spins) {
                        unsigned int v;
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
                        if (static_cast<int>(v - target) >= 0) break;
                        if ((spins & 1023u) == 1023u) {
                            const uint64_t now = globaltimer_ns();
                            if (t0 == 0) t0 = now;
                            if (now - t0 > 4000000000ull) {
                                printf("sr3: split-K wait timeout cta=%d tile=%d counter=%u target=%u\n", cta, tile, v, target);
                                __trap();
                            }
                        }
                    }
                    __threadfence();
                }
                __syncwarp();
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            const bool from_ws = (pass == 1);
            const bool to_ws = (ksplit > 1 && pass == 0);
            // iteration space: pass 0 / no split -- this warp's (half, chunk) items of its own TMEM lane quadrant;
            // pass 1 -- units u = item * 4 + quadrant with u % ksplit == sp, dealt round-robin to the eight warps
            const int it0 = from_ws ? sp + ksplit * ew : grp;
            const int itn = from_ws ? NITEMS * 4 : (has_work ? NITEMS : 0);
            const int its = from_ws ? ksplit * GEMM_EPI_WARPS : 2;
#pragma unroll 1
            for (int it = it0; it < itn; it += its) {
                const int item = from_ws ? (it >> 2) : it;
                const int qq = from_ws ? (it & 3) : q;
                int half, ch, sw, sh, c4;
                item_geom(item, qq, half, ch, sw, sh, c4);
                const bool last_item = !from_ws && (item + 2 >= NITEMS);
                const int row = half * 128 + qq * 32 + lane;
                const int w = row & (p.w_box - 1);
                const int h = (row >> p.w_shift) & (p.h_box - 1);
                const int bb = row >> (p.w_shift + p.h_shift);
                const int ow = w0 + w, oh = h0 + h, img = b0 + bb;
                const bool row_ok = (ow < p.OW) && (oh < p.OH) && (img < p.OB);
                if (p.stats && !(p.dbg & 2) && !to_ws) {
                    const int img0 = __shfl_sync(0xffffffffu, img, 0);      // all rows of a warp belong to one image
                    if (img0 != st_img || n0 != st_n0) { flush_stats(); st_img = img0; st_n0 = n0; }
                }
                const uint32_t t_acc = t_lane + acc * ACC_COLS + half * BLOCK_N;

                if constexpr (BLOCK_N == 16) {
                    uint32_t v[16];
                    tmem_ld_32x16(t_acc, v);
                    tmem_ld_wait();
                    if (last_item) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty_bar(acc));
                    }
                    if (row_ok) {
                        float eps[4] = {0.f, 0.f, 0.f, 0.f};
                        for (int c = 0; c < p.post.C; ++c) eps[c] = __uint_as_float(v[c]) + __ldg(&p.bias[c]);
                        final_epilogue(p, eps, img, oh, ow);
                    }
                } else {
                    const int nb = n0 + ch * 32;
                    float bv = 0.f;
                    if (from_ws) {
                        const int n = nb + lane;
                        if (n < p.n_valid) {
                            if (p.bias) bv += __ldg(&p.bias[n]);
                            if (p.bias2) {
                                const int img0 = b0 + ((half * 128 + qq * 32) >> (p.w_shift + p.h_shift));
                                bv += __ldcg(&p.bias2[static_cast<long long>(img0 < p.OB ? img0 : 0) * p.bias2_stride + n]);
                            }
                        }
                    } else {
#pragma unroll
                        for (int ii = 0; ii < MAXI; ++ii)
                            if (item == grp + 2 * ii) bv = bvs[ii];
                    }
                    float* bs = bias_s + (ew * 2 + ((from_ws ? (it / its) : (item >> 1)) & 1)) * 32;
                    bs[lane] = bv;
                    __syncwarp();
                    if (use_res_tma) {
                        ++res_count;                             // this item is residual request number res_count
                        if (!last_item && lane == 0) request_resid(item + 2);   // other buffer: freed one item ago
                    }
                    uint32_t v[32];
                    // this thread's eight float4 inside a partial tile: laid out [chunk][j][row] so that the 32 lanes (consecutive rows)
                    // of one store / load instruction touch 512 contiguous bytes
                    const long long wrow = (static_cast<long long>(ch) * 8 * (MH * 128) + row) * 4;
                    constexpr int WJ = MH * 128;                  // float4 stride between the j-th and (j+1)-th quad of a row
                    if (!from_ws) {
                        tmem_ld_32x32(t_acc + ch * 32, v);
                        tmem_ld_wait();
                        if (last_item) {                         // this warp's share of the accumulator is read: hand it back
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(tempty_bar(acc));
                        }
                        if (to_ws) {                             // split-K: park the partial sums
                            float4* dst = reinterpret_cast<float4*>(p.ws + static_cast<long long>(tile) * SLICE + wrow);
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                __stcg(dst + j * WJ, make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                                            __uint_as_float(v[4 * j + 3])));
                            continue;
                        }
                    } else {                                     // sum the ksplit partial tiles (L2 resident), split 0 first
                        float acc_f[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc_f[j] = 0.f;
                        const float* src0 = p.ws + static_cast<long long>(tile - sp) * SLICE + wrow;
                        for (int s2 = 0; s2 < ksplit; ++s2) {
                            const float4* src = reinterpret_cast<const float4*>(src0 + s2 * SLICE);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 r = __ldcg(src + j * WJ);
                                acc_f[4 * j] += r.x; acc_f[4 * j + 1] += r.y; acc_f[4 * j + 2] += r.z; acc_f[4 * j + 3] += r.w;
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(acc_f[j]);
                    }
                    const float ep_scale = p.scale;
                    if (p.dbg & 1) {
                        if (use_res_tma) {
                            const uint32_t b = (res_count - 1) & 1;
                            mbar_wait(res_bar(ew, b), (res_phase >> b) & 1u, 5);
                            res_phase ^= (1u << b);
                            __syncwarp();
                        }
                        continue;
                    }
                    float f[32];
                    const bool full = (nb + 32 <= p.n_valid);
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {               // bias broadcast: 8 x LDS.128 (the slot is 128 B aligned)
                        const float4 bq = reinterpret_cast<const float4*>(bs)[j4];
                        const float bb4[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int j = 4 * j4 + jj;
                            float x = __uint_as_float(v[j]) * ep_scale + bb4[jj];
                            if (!full && nb + j >= p.n_valid) x = 0.f;
                            f[j] = x;
                        }
                    }
                    if (use_res_tma) {
                        const uint32_t b = (res_count - 1) & 1;
                        mbar_wait(res_bar(ew, b), (res_phase >> b) & 1u, 5);
                        res_phase ^= (1u << b);
                        const uint8_t* rp = res_ptr + b * 4096 + lane * 128;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 r = *reinterpret_cast<const float4*>(rp + ((j ^ (lane & 7)) << 4));
                            f[4 * j] += r.x; f[4 * j + 1] += r.y; f[4 * j + 2] += r.z; f[4 * j + 3] += r.w;
                        }
                        __syncwarp();                            // everyone is done with the buffer before it is re-requested
                    } else if (row_ok && p.resid) {
                        const long long ro = out_index(p.rs, z, img, oh, ow);
                        for (int j = 0; j < 32; ++j)
                            if (nb + j < p.n_valid) f[j] += __ldcg(&p.resid[ro + nb + j]);
                    }
                    if (p.dbg & 4) {
                    } else if (use_out_tma) {
                        if (out_pending) {                       // the previous bulk store must have finished reading the staging buffer
                            if (lane == 0) tma_store_wait_read<0>();
                            __syncwarp();
                        }
                        uint8_t* op = out_ptr + lane * 128;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            *reinterpret_cast<float4*>(op + ((j ^ (lane & 7)) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            if (p.z_phase) tma_store_5d(&pm->out_map, out_smem, (z & 1) * p.n_valid + nb, w0 + sw, z >> 1, h0 + sh, c4);
                            else tma_store_5d(&pm->out_map, out_smem, nb, w0 + sw, 0, h0 + sh, c4);
                            tma_store_commit();
                        }
                        out_pending = true;
                    } else if (row_ok && p.out_f32) {
                        const long long oo = out_index(p.os, z, img, oh, ow) + (p.z_phase ? (z >> 1) * p.z_off_hi + (z & 1) * p.z_off_lo : 0);
                        if (full) {
                            float4* o4 = reinterpret_cast<float4*>(p.out_f32 + oo + nb);
#pragma unroll
                            for (int j = 0; j < 8; ++j) o4[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                        } else {
                            for (int j = 0; j < 32; ++j)
                                if (nb + j < p.n_valid) p.out_f32[oo + nb + j] = f[j];
                        }
                    }
                    if (p.out_t && nb >= p.t_col0) {
                        if (ow < p.OW && oh < p.OH && nb + 32 <= p.n_valid) {          // padded images too: their (finite) values are multiplied by P = 0 later
                            __nv_bfloat16* tp = p.out_t + (static_cast<long long>(img / p.t_per) * p.t_rows + (nb - p.t_col0)) * p.t_ld +
                                                (img % p.t_per) * (p.OH * p.OW) + oh * p.OW + ow;
#pragma unroll
                            for (int j = 0; j < 32; ++j) {                                                              // lanes = consecutive tokens
                                const __nv_bfloat16 hi = __float2bfloat16_rn(f[j]);
                                tp[static_cast<long long>(j) * p.t_ld] = hi;
                                if (p.lo_t_off) tp[static_cast<long long>(j) * p.t_ld + p.lo_t_off] = __float2bfloat16_rn(f[j] - __bfloat162float(hi));
                            }
                        }
                    } else if (row_ok && p.out_bf16) {
                        const long long ho = out_index(p.hs, z, img, oh, ow);
                        if (full) {
                            uint4* o4 = reinterpret_cast<uint4*>(p.out_bf16 + ho + nb);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                __nv_bfloat162 h0v = __floats2bfloat162_rn(f[8 * j], f[8 * j + 1]);
                                __nv_bfloat162 h1v = __floats2bfloat162_rn(f[8 * j + 2], f[8 * j + 3]);
                                __nv_bfloat162 h2v = __floats2bfloat162_rn(f[8 * j + 4], f[8 * j + 5]);
                                __nv_bfloat162 h3v = __floats2bfloat162_rn(f[8 * j + 6], f[8 * j + 7]);
                                uint4 uu;
                                uu.x = *reinterpret_cast<uint32_t*>(&h0v); uu.y = *reinterpret_cast<uint32_t*>(&h1v);
                                uu.z = *reinterpret_cast<uint32_t*>(&h2v); uu.w = *reinterpret_cast<uint32_t*>(&h3v);
                                o4[j] = uu;
                            }
                        } else {
                            for (int j = 0; j < 32; ++j)
                                if (nb + j < p.n_valid) p.out_bf16[ho + nb + j] = __float2bfloat16_rn(f[j]);
                        }
                        if (p.lo_out_off) {                       // precise mode: low halves
                            __nv_bfloat16* lp = p.out_bf16 + ho + nb + p.lo_out_off;
                            if (full) {
                                uint4* o4 = reinterpret_cast<uint4*>(lp);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    float r[8];
#pragma unroll
                                    for (int u = 0; u < 8; ++u) r[u] = f[8 * j + u] - __bfloat162float(__float2bfloat16_rn(f[8 * j + u]));
                                    __nv_bfloat162 h0v = __floats2bfloat162_rn(r[0], r[1]), h1v = __floats2bfloat162_rn(r[2], r[3]);
                                    __nv_bfloat162 h2v = __floats2bfloat162_rn(r[4], r[5]), h3v = __floats2bfloat162_rn(r[6], r[7]);
                                    uint4 uu;
                                    uu.x = *reinterpret_cast<uint32_t*>(&h0v); uu.y = *reinterpret_cast<uint32_t*>(&h1v);
                                    uu.z = *reinterpret_cast<uint32_t*>(&h2v); uu.w = *reinterpret_cast<uint32_t*>(&h3v);
                                    o4[j] = uu;
                                }
                            } else {
                                for (int j = 0; j < 32; ++j)
                                    if (nb + j < p.n_valid) lp[j] = __float2bfloat16_rn(f[j] - __bfloat162float(__float2bfloat16_rn(f[j])));
                            }
                        }
                    }
                    if (p.stats && !(p.dbg & 2)) {
                        double cs, cq;
                        if (use_out_tma && !(p.dbg & 4)) {
                            // the 32x32 tile sits in the (swizzled) staging buffer: lane c walks down column c -- conflict free, and a
                            // third of the instructions of the shuffle transposition below
                            const unsigned int okm = __ballot_sync(0xffffffffu, row_ok);
                            const uint8_t* colp = out_ptr + ((lane & 3) << 2);
                            const int cq4 = lane >> 2;
                            // shifted sums: the fp32 partial sums run over x - x[row 0] (values of the size of the spread, not of the
                            // mean), the shift is put back in fp64:  sum x = a + n s,  sum x^2 = q + 2 s a + n s^2
                            const float sft = *reinterpret_cast<const float*>(colp + (cq4 << 4));
                            float a0 = 0.f, a1 = 0.f, q0 = 0.f, q1 = 0.f;
                            if (okm == 0xffffffffu) {            // (warp-uniform) the usual case: no padded pixel / image in this chunk
#pragma unroll
                                for (int r = 0; r < 32; r += 2) {
                                    const float x0 = *reinterpret_cast<const float*>(colp + r * 128 + ((cq4 ^ (r & 7)) << 4)) - sft;
                                    const float x1 = *reinterpret_cast<const float*>(colp + (r + 1) * 128 + ((cq4 ^ ((r + 1) & 7)) << 4)) - sft;
                                    a0 += x0; q0 = fmaf(x0, x0, q0);
                                    a1 += x1; q1 = fmaf(x1, x1, q1);
                                }
                            } else {
#pragma unroll
                                for (int r = 0; r < 32; r += 2) {
                                    float x0 = *reinterpret_cast<const float*>(colp + r * 128 + ((cq4 ^ (r & 7)) << 4)) - sft;
                                    float x1 = *reinterpret_cast<const float*>(colp + (r + 1) * 128 + ((cq4 ^ ((r + 1) & 7)) << 4)) - sft;
                                    if (!((okm >> r) & 1u)) x0 = 0.f;
                                    if (!((okm >> (r + 1)) & 1u)) x1 = 0.f;
                                    a0 += x0; q0 = fmaf(x0, x0, q0);
                                    a1 += x1; q1 = fmaf(x1, x1, q1);
                                }
                            }
                            const double n_ok = static_cast<double>(__popc(okm)), ds = static_cast<double>(sft);
                            const double da = static_cast<double>(a0 + a1), dq = static_cast<double>(q0 + q1);
                            cs = da + n_ok * ds;
                            cq = dq + 2.0 * ds * da + n_ok * ds * ds;
                        } else {
                            float s2[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const float x = row_ok ? f[j] : 0.f;
                                f[j] = x; s2[j] = x * x;
                            }
                            cs = static_cast<double>(warp_column_sums(f));
                            cq = static_cast<double>(warp_column_sums(s2));
                        }
#pragma unroll
                        for (int c = 0; c < NCHS; ++c)
                            if (c == ch) { st_sum[c] += cs; st_sq[c] += cq; }
                    }
                }
            }       // items
            }       // pass
        }           // tiles
        if (p.stats && !(p.dbg & 2)) flush_stats();
        if constexpr (MEGA) {
            // the consumer is a later op of the SAME launch (other CTAs, after a grid barrier): the bulk stores must be complete in
            // global memory, not merely done reading shared memory
            if (use_out_tma && out_pending && lane == 0) tma_store_wait_all<0>();
        } else {
            if (use_out_tma && out_pending && lane == 0) tma_store_wait_read<0>();     // smem must outlive the reads of the last bulk stores
        }
    }
    tc_fence_before();
    if constexpr (!MEGA) __syncthreads();      // (the step kernel fences and synchronises after every op itself)
    if constexpr (!MEGA) {
        if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int BLOCK_N, int MH>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tile_kernel(const __grid_constant__ GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    gemm_tile_body<BLOCK_N, MH, false>(p, &p, base, smem_raw + (base - raw), 0u, blockIdx.x, gridDim.x);
}

}  // namespace sr3
