// Host side of libsr3_b200.so: builds the per-step kernel plan of the SR3 UNet + posterior update for one
// (config, batch), owns device buffers / packed weights / TMA descriptors, captures the step as a CUDA graph and exposes
// the C ABI declared in include/sr3_b200.h.  Reference call sites are cited in the header next to each entry point.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/sr3_b200.h"
#include "aux_kernels.cuh"
#include "attn_tcgen05.cuh"
#include "step_megakernel.cuh"
#include "train_kernels.cuh"

using namespace sr3;
typedef __nv_bfloat16 bf16;

namespace {

thread_local std::string g_err;

std::string fmt(const char* f, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, f);
    vsnprintf(buf, sizeof(buf), f, ap);
    va_end(ap);
    return std::string(buf);
}
#define CK(expr)                                                                                                  \
    do {                                                                                                          \
        cudaError_t e_ = (expr);                                                                                  \
        if (e_ != cudaSuccess) throw std::runtime_error(fmt("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_))); \
    } while (0)
#define REQUIRE(cond, ...)                                                 \
    do {                                                                   \
        if (!(cond)) throw std::runtime_error(fmt(__VA_ARGS__));           \
    } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    REQUIRE(p != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

CUtensorMap encode_map(int rank, const void* ptr, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                       CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
    CUtensorMap m;
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
    // a box may be larger than the tensor extent (halo rows of small images): TMA zero-fills / clips the out-of-range part
    for (int i = 0; i < rank; ++i) REQUIRE(box[i] >= 1 && box[i] <= 256, "TMA box %u out of range (axis %d)", box[i], i);
    REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA base not 16B aligned");
    for (int i = 0; i < rank - 1; ++i) REQUIRE(strides_bytes[i] % 16 == 0 && strides_bytes[i] > 0, "TMA stride %llu (axis %d) must be a positive multiple of 16", (unsigned long long)strides_bytes[i], i + 1);
    CUresult r = get_encode_fn()(&m, dtype, rank, const_cast<void*>(ptr), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rank=%d dims=%llu,%llu box=%u,%u", (int)r, rank,
            (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return m;
}

// 5-D source view of an A operand: dims (C', W', P, H', Bn), byte strides of dims 1..4
struct ASrc {
    const void* ptr = nullptr;
    int C = 0, W = 1, P = 1, H = 1, Bn = 1;
    long long sW = 0, sP = 0, sH = 0, sB = 0;
};
ASrc nhwc_src(const void* ptr, int Bn, int H, int W, int C) {
    ASrc a; a.ptr = ptr; a.C = C; a.W = W; a.P = 1; a.H = H; a.Bn = Bn;
    a.sW = 2LL * C; a.sP = 2LL * W * C; a.sH = 2LL * W * C; a.sB = 2LL * H * W * C;
    return a;
}
// stride-2 view of an NHWC tensor: (2C, W/2, 2, H/2, B)
ASrc nhwc_stride2_src(const void* ptr, int Bn, int H, int W, int C) {
    ASrc a; a.ptr = ptr; a.C = 2 * C; a.W = W / 2; a.P = 2; a.H = H / 2; a.Bn = Bn;
    a.sW = 4LL * C; a.sP = 2LL * W * C; a.sH = 4LL * W * C; a.sB = 2LL * H * W * C;
    return a;
}
// batched row-major matrices [nb][rows][K] (row stride ld elements)
ASrc matrix_src(const void* ptr, int nb, int rows, int K, long long ld, long long batch_stride_elems) {
    ASrc a; a.ptr = ptr; a.C = K; a.W = rows; a.P = 1; a.H = 1; a.Bn = nb;
    a.sW = 2LL * ld; a.sP = 2LL * ld * rows; a.sH = 2LL * ld * rows; a.sB = 2LL * batch_stride_elems;
    if (nb == 1) a.sB = a.sH;
    return a;
}

struct KSlab { int a_sel, a_chan, dw, dh, p, b_col; };

struct GemmDesc {
    ASrc a[2];
    int n_a = 1;
    const void* b_ptr = nullptr;
    long long b_rows = 0; int b_K = 0;            // B matrix [b_rows][b_K] bf16 row-major
    std::vector<KSlab> slabs;
    int block_n = 128;
    int w_box = 16, h_box = 8, b_box = 1;         // output pixel patch of one tile (mh * 128 rows)
    int mh = 1;                                   // 128-row accumulator halves per tile
    int tall = 0;                                 // 3x3 stride-1 "tall halo" mode: A box = 8 x (rows + 2) pixels, vertical taps share it
    int a_box_w = 0, a_box_h = 0, a_box_b = 0;    // TMA box of the A operand (0: same as the tile patch)
    int a_half_off = 0;
    bool b_is_param = false;                      // B operand is a packed weight matrix (constant within a step): eligible for L2 prefetch
    int ksplit_max = 1;                           // split-K allowed up to this factor (image convs with few output tiles)
    int tiles_w = 1, tiles_h = 1, tiles_b = 1, n_tiles = 1, nz = 1;
    int a_zstep = 0, b_zrows = 0;
    int z_phase = 0; long long z_off_hi = 0, z_off_lo = 0;   // nz = 4 output phases of a folded upsample conv (gemm_tcgen05.cuh)
    int passes = 1, lo_b_col = 0, lo_a_chan[2] = {0, 0};      // precise mode: three passes over the stage table (gemm_tcgen05.cuh)
    long long lo_out_off = 0, lo_t_off = 0;
    // epilogue
    int mode = 0, OW = 0, OH = 1, OB = 1, n_valid = 0;
    float scale = 1.f;
    const float* bias = nullptr; const float* bias2 = nullptr; int bias2_stride = 0;
    const float* resid = nullptr; OutSpec rs{};
    float* out_f32 = nullptr; OutSpec os{};
    bf16* out_bf16 = nullptr; OutSpec hs{};
    bf16* out_t = nullptr; int t_col0 = 0, t_rows = 0, t_ld = 0, t_per = 1;   // transposed bf16 store of columns >= t_col0
    double* stats = nullptr; int stats_C = 0, stats_coff = 0;
    const StepCtl* ctl = nullptr; PostParams post{};
};

OutSpec nhwc_out(int H, int W, int C, long long off = 0) {
    OutSpec s; s.sZ = 0; s.sB = 1LL * H * W * C; s.sH = 1LL * W * C; s.sW = C; s.off = off; return s;
}

constexpr int SMEM_LIMIT = 232448;   // 227 KB opt-in maximum per CTA on sm_100
int pick_stages(int block_n, int a_stage_bytes, int b_taps, bool resid, int num_k) {
    int s = GEMM_MAX_STAGES;
    if (const char* e = getenv("SR3_STAGES")) s = atoi(e);
    if (s > GEMM_MAX_STAGES) s = GEMM_MAX_STAGES;
    while (s > 1 && gemm_smem_bytes(block_n, a_stage_bytes, b_taps, s, resid, num_k) > SMEM_LIMIT) --s;
    if (s < 1) s = 1;
    return s;
}
int current_device() { int dev = 0; CK(cudaGetDevice(&dev)); return dev; }
int num_sms() {
    static std::map<int, int> cache;               // per device: a process may hold engines on several GPUs
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    const int dev = current_device();
    auto it = cache.find(dev);
    if (it != cache.end()) return it->second;
    int n = 0;
    CK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    cache[dev] = n;
    return n;
}
// cudaFuncSetAttribute is per device: remember which devices already have the opt-in shared-memory size of a kernel family
bool first_use_on_device(std::vector<int>& seen) {
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    const int dev = current_device();
    for (int d : seen) if (d == dev) return false;
    seen.push_back(dev);
    return true;
}
// Every kernel of the step is launched with programmatic stream serialization (PDL): its prologue overlaps the tail of the
// previous kernel; the kernels call griddepcontrol.wait before touching upstream data.
bool use_pdl() { static int v = -1; if (v < 0) v = getenv("SR3_NO_PDL") ? 0 : 1; return v == 1; }
template <typename... KArgs, typename... Args>
void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = use_pdl() ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
}
// Split-K layers: the `ksplit` CTAs that share an output tile are consecutive blocks and wait for each other inside the kernel.  They are
// launched as ONE THREAD-BLOCK CLUSTER (cluster dimension = ksplit): the hardware gang-schedules a cluster, so the partners are co-resident
// by construction -- no assumption about what else occupies the device (other engines / streams, NCCL, library kernels), inside or outside
// graph capture, and the launch keeps its programmatic-dependent-launch edge.  SR3_NO_CLUSTER=1: the round-1 form (plain launch inside
// graphs, cooperative launch outside).
bool use_cluster_split() { static int v = -1; if (v < 0) v = getenv("SR3_NO_CLUSTER") ? 0 : 1; return v == 1; }
constexpr int MAX_CLUSTER_SPLIT = 8;               // portable cluster size limit
template <int BN, int MH>
void launch_gemm_bn(const GemmParams& p, dim3 grid, int smem, cudaStream_t st) {
    if (p.ksplit > 1 && use_cluster_split()) {
        REQUIRE(p.ksplit <= MAX_CLUSTER_SPLIT && grid.x % p.ksplit == 0, "split-K factor %d does not form clusters of grid %u", p.ksplit, grid.x);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = grid; cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)p.ksplit; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = use_pdl() ? 2 : 1;
        CK(cudaLaunchKernelEx(&cfg, gemm_tile_kernel<BN, MH>, p));
        return;
    }
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (p.ksplit > 1) CK(cudaStreamIsCapturing(st, &cap));
    if (p.ksplit > 1 && cap == cudaStreamCaptureStatusNone && getenv("SR3_NO_COOP") == nullptr) {
        // (SR3_NO_CLUSTER) split-K CTAs wait for their partners inside the kernel: launch cooperatively so that the runtime guarantees
        // co-residency (or fails the launch) even when another stream / engine / library kernel holds SMs.  No PDL overlap for these launches.
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = grid; cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, gemm_tile_kernel<BN, MH>, p));
        return;
    }
    launch_k(gemm_tile_kernel<BN, MH>, grid, dim3(GEMM_THREADS), (size_t)smem, st, p);
}
void init_gemm_attrs() {
    static std::vector<int> seen;
    if (!first_use_on_device(seen)) return;
    CK(cudaFuncSetAttribute(gemm_tile_kernel<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    CK(cudaFuncSetAttribute(gemm_tile_kernel<16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    CK(cudaFuncSetAttribute(gemm_tile_kernel<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    CK(cudaFuncSetAttribute(gemm_tile_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    CK(cudaFuncSetAttribute(gemm_tile_kernel<64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    CK(cudaFuncSetAttribute(gemm_tile_kernel<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    CK(cudaFuncSetAttribute(gemm_tile_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    CK(cudaFuncSetAttribute(gemm_tile_kernel<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
}

// How many clusters of `c` tile-kernel CTAs (one CTA per SM: they use the whole shared memory) the device holds at once; a cluster lives
// inside one GPC, so this is less than SMs / c.  Split-K factors are chosen so that all clusters of a layer run in one wave.
int cluster_capacity(int c) {
    static std::map<std::pair<int, int>, int> cache;
    static std::mutex mu;
    if (c <= 1) return num_sms();
    init_gemm_attrs();
    std::lock_guard<std::mutex> lock(mu);
    const std::pair<int, int> key(current_device(), c);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(c * 32); cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = SMEM_LIMIT;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)c; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm_tile_kernel<64, 1>, &cfg) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        n = 8 * ((num_sms() / 8) / c);               // eight GPCs of equal size, conservatively
    }
    if (const char* e = getenv("SR3_CLUSTER_DEBUG")) { if (atoi(e)) fprintf(stderr, "sr3: cluster_capacity(%d) = %d\n", c, n); }
    cache[key] = n;
    return n;
}
// largest split-K factor <= want whose clusters all fit the device together with `tiles` output tiles
int fit_split(int want, long long tiles) {
    if (!use_cluster_split()) return want;
    if (want > MAX_CLUSTER_SPLIT) want = MAX_CLUSTER_SPLIT;
    while (want > 1 && cluster_capacity(want) < tiles) --want;
    return want;
}

struct DevAllocs {
    std::vector<void*> ptrs;
    long long bytes = 0;
    void* alloc(size_t n, bool zero = true) {
        void* p = nullptr;
        if (n == 0) n = 16;
        CK(cudaMalloc(&p, n));
        if (zero) CK(cudaMemset(p, 0, n));
        ptrs.push_back(p);
        bytes += (long long)n;
        return p;
    }
    ~DevAllocs() { for (void* p : ptrs) cudaFree(p); }
};

typedef std::function<void(cudaStream_t)> Op;
struct GemmHandle { std::shared_ptr<GemmParams> p; const void* w_ptr = nullptr; long long w_bytes = 0; bool w_is_param = false; int bn = 0, mh = 0; };
// One op of the persistent step kernel (step_megakernel.cuh), recorded next to the per-layer launch it replaces.
struct MegaRec { int type = 0, variant = 0; std::shared_ptr<GemmParams> gp; std::vector<uint8_t> raw; };
static thread_local std::vector<MegaRec>* g_mega_registry = nullptr;
template <typename T>
void mega_record(int type, const T& params) {
    if (!g_mega_registry) return;
    MegaRec r; r.type = type;
    r.raw.resize((sizeof(T) + 3) & ~size_t(3));
    memcpy(r.raw.data(), &params, sizeof(T));
    g_mega_registry->push_back(std::move(r));
}
static thread_local std::vector<GemmHandle>* g_gemm_registry = nullptr;   // set by the engine while it builds its plan

// Turns a GemmDesc into a launchable op (encodes the TMA maps, uploads the K-slab table).
Op make_gemm_op(const GemmDesc& d, DevAllocs& mem) {
    REQUIRE(d.w_box * d.h_box * d.b_box == 128 * d.mh, "tile box must cover %d rows", 128 * d.mh);
    REQUIRE(!d.slabs.empty(), "gemm without K slabs");
    GemmParams p;
    memset(&p, 0, sizeof(p));
    const int abw = d.a_box_w ? d.a_box_w : d.w_box, abh = d.a_box_h ? d.a_box_h : d.h_box, abb = d.a_box_b ? d.a_box_b : d.b_box;
    p.a_stage_bytes = abw * abh * abb * 128;
    REQUIRE(p.a_stage_bytes % 1024 == 0, "A box must be a whole number of swizzle atoms");
    for (int i = 0; i < 2; ++i) {
        const ASrc& a = d.a[i < d.n_a ? i : 0];
        REQUIRE(a.C % 64 == 0, "A channels (%d) must be a multiple of 64", a.C);
        const uint64_t dims[5] = {(uint64_t)a.C, (uint64_t)a.W, (uint64_t)a.P, (uint64_t)a.H, (uint64_t)a.Bn};
        const uint64_t str[4] = {(uint64_t)a.sW, (uint64_t)a.sP, (uint64_t)a.sH, (uint64_t)a.sB};
        const uint32_t box[5] = {64u, (uint32_t)abw, 1u, (uint32_t)abh, (uint32_t)abb};
        p.a_map[i] = encode_map(5, a.ptr, dims, str, box);
    }
    {
        REQUIRE(d.b_K % 64 == 0, "B K extent must be a multiple of 64");
        const uint64_t dims[2] = {(uint64_t)d.b_K, (uint64_t)d.b_rows};
        const uint64_t str[1] = {2ull * d.b_K};
        const uint32_t box[2] = {64u, (uint32_t)d.block_n};
        REQUIRE(d.b_rows >= d.block_n, "B rows %lld < block_n %d", d.b_rows, d.block_n);
        p.b_map = encode_map(2, d.b_ptr, dims, str, box);
    }
    // stage table.  Tall mode: the three vertical taps (dh = -1, 0, +1) of one (source, channel chunk, dw) share one halo box.
    // Generic mode: up to three consecutive K slabs of the same source are grouped into a stage (one box each).
    std::vector<StageDesc> tab;
    int b_taps = 1, a_boxes = 1;
    const int group_max = getenv("SR3_GROUP") ? atoi(getenv("SR3_GROUP")) : 3;
    for (size_t i = 0; i < d.slabs.size(); ++i) {
        const KSlab& k = d.slabs[i];
        REQUIRE(k.a_sel < d.n_a, "slab refers to missing A source");
        StageDesc e; memset(&e, 0, sizeof(e));
        e.a_sel = k.a_sel; e.ntaps = 1;
        e.tap[0].a_chan = k.a_chan; e.tap[0].dw = k.dw; e.tap[0].dh = k.dh; e.tap[0].p = k.p; e.tap[0].b_col = k.b_col;
        if (d.tall) {
            // all vertical taps (dh) of this (source, channel chunk, dw) share the halo box: 3 for a 3x3 conv, 2 for a phase of a
            // folded upsample conv, 1 for the 1x1 shortcut
            bool first = true;
            int nt = 0;
            for (const KSlab& o : d.slabs) {
                if (o.a_sel != k.a_sel || o.a_chan != k.a_chan || o.dw != k.dw) continue;
                if (&o < &k) { first = false; break; }
                REQUIRE(nt < 3 && o.p == 0 && o.dh >= -1 && o.dh <= 1, "bad tall tap group");
                e.tap[nt] = e.tap[0];
                e.tap[nt].dh = -1;                        // the box always starts one row above the tile
                e.tap[nt].b_col = o.b_col;
                e.tap[nt].a_off = (o.dh + 1) * 1024;
                ++nt;
            }
            if (!first) continue;                         // folded into the stage of the group's first slab
            e.ntaps = nt;
            if (nt > b_taps) b_taps = nt;
        } else {
            e.a_multi = 1;
            int n = 1;
            while (n < group_max && i + n < d.slabs.size() && d.slabs[i + n].a_sel == k.a_sel) ++n;
            e.ntaps = n;
            for (int t = 0; t < n; ++t) {
                const KSlab& o = d.slabs[i + t];
                e.tap[t].a_chan = o.a_chan; e.tap[t].dw = o.dw; e.tap[t].dh = o.dh; e.tap[t].p = o.p; e.tap[t].b_col = o.b_col;
                e.tap[t].a_off = t * p.a_stage_bytes;     // boxes back to back (p.a_stage_bytes still holds ONE box here)
            }
            if (n > b_taps) b_taps = n;
            if (n > a_boxes) a_boxes = n;
            i += n - 1;
        }
        tab.push_back(e);
    }
    p.a_box_bytes = p.a_stage_bytes;
    p.a_stage_bytes = p.a_box_bytes * a_boxes;
    REQUIRE((int)tab.size() <= GEMM_MAX_K, "gemm with %d stages per tile (max %d)", (int)tab.size(), GEMM_MAX_K);
    StageDesc* dtab = static_cast<StageDesc*>(mem.alloc(tab.size() * sizeof(StageDesc), false));
    CK(cudaMemcpy(dtab, tab.data(), tab.size() * sizeof(StageDesc), cudaMemcpyHostToDevice));
    p.ktab = dtab; p.num_k = (int)tab.size();
    p.b_taps = b_taps; p.a_half_off = d.a_half_off;
    p.tiles_w = d.tiles_w; p.tiles_h = d.tiles_h; p.tiles_b = d.tiles_b;
    p.w_box = d.w_box; p.h_box = d.h_box; p.b_box = d.b_box;
    {
        auto lg = [](int v) { int l = 0; while ((1 << l) < v) ++l; return l; };
        REQUIRE((d.w_box & (d.w_box - 1)) == 0 && (d.h_box & (d.h_box - 1)) == 0, "tile box %dx%d must be powers of two", d.w_box, d.h_box);
        p.w_shift = lg(d.w_box); p.h_shift = lg(d.h_box);
    }
    p.a_zstep = d.a_zstep; p.b_zrows = d.b_zrows;
    p.dbg = getenv("SR3_DBG") ? atoi(getenv("SR3_DBG")) : 0;
    p.n_tiles = d.n_tiles; p.nz = d.nz;
    p.mode = d.mode; p.OW = d.OW; p.OH = d.OH; p.OB = d.OB; p.n_valid = d.n_valid; p.scale = d.scale;
    p.bias = d.bias; p.bias2 = d.bias2; p.bias2_stride = d.bias2_stride;
    p.resid = d.resid; p.rs = d.rs; p.out_f32 = d.out_f32; p.os = d.os; p.out_bf16 = d.out_bf16; p.hs = d.hs;
    p.out_t = d.out_t; p.t_col0 = d.t_col0; p.t_rows = d.t_rows; p.t_ld = d.t_ld; p.t_per = d.t_per > 0 ? d.t_per : 1;
    p.stats = d.stats; p.stats_C = d.stats_C; p.stats_coff = d.stats_coff; p.ctl = d.ctl; p.post = d.post;
    p.t_fixed = -1;
    p.z_phase = d.z_phase; p.z_off_hi = d.z_off_hi; p.z_off_lo = d.z_off_lo;
    p.passes = d.passes; p.lo_b_col = d.lo_b_col; p.lo_a_chan[0] = d.lo_a_chan[0]; p.lo_a_chan[1] = d.lo_a_chan[1];
    p.lo_out_off = d.lo_out_off; p.lo_t_off = d.lo_t_off;
    if (d.stats) REQUIRE((d.w_box * d.h_box) % 32 == 0, "stats need whole warps per image");
    // fp32 output / residual through smem + TMA: one 32-row x 32-column box per epilogue warp
    p.tma_epi = (d.mode == 0 && getenv("SR3_NO_TMA_EPI") == nullptr && (d.out_f32 || d.resid)) ? 1 : 0;
    if (p.tma_epi) {
        const int w_sub = d.w_box < 32 ? d.w_box : 32, h_sub = 32 / w_sub;
        REQUIRE(d.w_box % w_sub == 0 && (d.h_box % h_sub == 0 || d.h_box == 1) && (h_sub == 1 || (128 / w_sub) % h_sub == 0), "tile box %dx%d cannot be split into per-warp boxes", d.w_box, d.h_box);
        auto mk = [&](const float* ptr, const OutSpec& o, bool& c4z) {
            c4z = (o.sB == 0 && o.sZ != 0);
            const long long s4 = c4z ? o.sZ : o.sB;
            const uint64_t n4 = c4z ? (uint64_t)d.nz : (uint64_t)(d.tiles_b * d.b_box > d.OB ? d.tiles_b * d.b_box : d.OB);
            uint64_t dims[5] = {(uint64_t)d.n_valid, (uint64_t)d.OW, 1ull, (uint64_t)d.OH, n4};
            uint64_t str[4];
            str[0] = (uint64_t)o.sW * 4;
            str[1] = str[0] * d.OW;
            str[2] = d.OH > 1 ? (uint64_t)o.sH * 4 : str[1];
            str[3] = s4 != 0 ? (uint64_t)s4 * 4 : str[2] * d.OH;
            if (d.z_phase) {      // {2N (px, n), W, 2 (py), H, B}: phase (py, px) is coordinate 2 and an offset of px * N in coordinate 0
                REQUIRE(d.z_off_lo == d.n_valid && o.sW == 2 * d.n_valid, "phase-batched output must be NHWC with 2x the width");
                dims[0] = 2ull * d.n_valid; dims[2] = 2ull; str[1] = (uint64_t)d.z_off_hi * 4;
            }
            const uint32_t box[5] = {32u, (uint32_t)w_sub, 1u, (uint32_t)h_sub, 1u};
            REQUIRE(d.n_valid >= 32, "TMA epilogue needs at least 32 output columns");
            return encode_map(5, ptr + o.off, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
        };
        bool zo = false, zr = false;
        if (d.out_f32) p.out_map = mk(d.out_f32, d.os, zo);
        if (d.resid) p.res_map = mk(d.resid, d.rs, zr);
        if (d.out_f32 && d.resid) REQUIRE(zo == zr, "output and residual must share the batch coordinate");
        if (!d.out_f32) p.out_map = p.res_map;
        if (!d.resid) p.res_map = p.out_map;
        p.epi_c4_is_z = (d.out_f32 ? zo : zr) ? 1 : 0;
    } else {
        p.out_map = p.b_map; p.res_map = p.b_map;
    }
    // split-K: when the output has too few tiles to occupy the SMs, `ksplit` CTAs share a tile, each streams a slice of K and parks its
    // partial tile in `ws`; every one of them then finalises a 1/ksplit share of the tile (gemm_tcgen05.cuh, epilogue pass 1).
    // One (tile, split) pair per SM at most: all CTAs are co-resident, which the in-kernel wait relies on.
    p.ksplit = 1;
    {
        const int tiles = d.tiles_w * d.tiles_h * d.tiles_b * d.n_tiles * d.nz;
        const int ks = d.ksplit_max;
        if (ks > 1 && d.mode == 0 && d.out_f32 && d.block_n >= 32 && d.n_valid % 32 == 0) {
            int want = num_sms() / tiles;                  // CTAs per tile that still fit one wave
            if (want > ks) want = ks;
            if (want > p.num_k * d.passes / 2) want = p.num_k * d.passes / 2;    // at least two stages per slice
            const int units = d.mh * (d.block_n / 32) * 4; // 32x32 units of a tile: every split finalises at least one
            if (want > units) want = units;
            want = fit_split(want, tiles);
            if (want > 1) {
                p.ksplit = want;
                p.ws = static_cast<float*>(mem.alloc((size_t)tiles * want * d.mh * 128 * d.block_n * sizeof(float), false));
                p.counters = static_cast<unsigned int*>(mem.alloc((size_t)tiles * sizeof(unsigned int)));
            }
        }
    }
    const int total_tiles = d.tiles_w * d.tiles_h * d.tiles_b * d.n_tiles * d.nz * p.ksplit;
    int ctas = total_tiles < num_sms() ? total_tiles : num_sms();
    if (const char* e = getenv("SR3_MAX_CTAS")) { int v = atoi(e); if (v > 0 && v < ctas) ctas = v; }
    const dim3 grid(ctas, 1, 1);
    const int bn = d.block_n;
    const int mh = d.mh;
    const bool res_smem = p.tma_epi && d.resid != nullptr && p.ksplit <= 1;
    p.stages = pick_stages(d.block_n, p.a_stage_bytes, p.b_taps, res_smem, p.num_k);
    const int smem = gemm_smem_bytes(bn, p.a_stage_bytes, p.b_taps, p.stages, res_smem, p.num_k);
    REQUIRE(smem <= SMEM_LIMIT, "gemm shared memory %d exceeds the limit", smem);
    init_gemm_attrs();
    REQUIRE((bn == 16 || bn == 32 || bn == 64 || bn == 128 || bn == 256) && (mh == 1 || (mh == 2 && bn <= 128 && bn != 32)), "unsupported tile %dx%d", 128 * mh, bn);
    std::shared_ptr<GemmParams> sp = std::make_shared<GemmParams>(p);
    if (g_gemm_registry) {
        GemmHandle h; h.p = sp; h.w_ptr = d.b_ptr; h.w_bytes = 2LL * d.b_rows * d.b_K; h.w_is_param = d.b_is_param; h.bn = bn; h.mh = mh;
        g_gemm_registry->push_back(h);
    }
    if (g_mega_registry) {
        MegaRec r; r.type = MOP_GEMM; r.variant = bn | (mh << 16); r.gp = sp;
        g_mega_registry->push_back(std::move(r));
    }
    return [sp, grid, bn, mh, smem](cudaStream_t st) {
        const GemmParams& p = *sp;
        switch (bn) {
            case 16: if (mh == 2) launch_gemm_bn<16, 2>(p, grid, smem, st); else launch_gemm_bn<16, 1>(p, grid, smem, st); break;
            case 32: launch_gemm_bn<32, 1>(p, grid, smem, st); break;
            case 64: if (mh == 2) launch_gemm_bn<64, 2>(p, grid, smem, st); else launch_gemm_bn<64, 1>(p, grid, smem, st); break;
            case 128: if (mh == 2) launch_gemm_bn<128, 2>(p, grid, smem, st); else launch_gemm_bn<128, 1>(p, grid, smem, st); break;
            default: launch_gemm_bn<256, 1>(p, grid, smem, st); break;
        }
    };
}

// Fused attention core (attn_tcgen05.cuh): S = q k^T / sqrt(C), softmax, O = P v in one launch.  qk [nz*Lt][2C], vT [nz*C][Lt], out [nz*Lt][C].
bool attn_fusable(int Lt, int C) { return getenv("SR3_NO_FUSED_ATTN") == nullptr && (Lt == 128 || Lt == 256) && C % 128 == 0 && C >= 128; }

Op make_attn_op(const bf16* qk, const bf16* vT, bf16* out, int nz, int Lt, int HW, int C) {
    REQUIRE(attn_fusable(Lt, C) && Lt % HW == 0, "attention shape Lt=%d HW=%d C=%d is not supported by the fused kernel", Lt, HW, C);
    AttnParams p;
    memset(&p, 0, sizeof(p));
    {
        const uint64_t dims[2] = {(uint64_t)2 * C, (uint64_t)nz * Lt};
        const uint64_t str[1] = {(uint64_t)2 * C * 2};
        const uint32_t box[2] = {64u, 128u};
        p.qk_map = encode_map(2, qk, dims, str, box);
    }
    {
        const uint64_t dims[2] = {(uint64_t)Lt, (uint64_t)nz * C};
        const uint64_t str[1] = {(uint64_t)Lt * 2};
        const uint32_t box[2] = {64u, 128u};
        p.vt_map = encode_map(2, vT, dims, str, box);
    }
    p.out = out; p.C = C; p.Lt = Lt; p.HW = HW; p.nz = nz;
    p.dn = (C % 256 == 0) ? 256 : 128;
    p.scale_log2e = 1.4426950408889634f / sqrtf((float)C);
    static std::vector<int> seen;
    if (first_use_on_device(seen)) CK(cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_SMEM_BYTES));
    const dim3 grid((Lt / 128) * (C / p.dn), nz, 1);
    mega_record(MOP_ATTN, p);
    return [p, grid](cudaStream_t st) { launch_k(attn_kernel, grid, dim3(ATTN_THREADS), ATTN_SMEM_BYTES, st, p); };
}

void pick_image_box(int W, int H, int& w_box, int& h_box, int& b_box);
int pick_block_n(int cout);

// Geometry of an image conv: the "tall halo" form for 3x3 stride-1 convs at >= 16x16, else a plain 128-pixel patch per tap.
// The tile shape (rows x BLOCK_N) and the split-K factor are chosen by a byte model of the per-CTA critical path: a CTA ingests
// stages x (A box + B boxes) through TMA at a fixed ~47 B/clk, runs ceil(tiles * split / SMs) waves, and a split tile costs an extra
// partial-tile store + reload + a grid-level handshake.  (Measured on B200: tools/gpu_splitk_sweep.py)
void conv_geometry(GemmDesc& d, int OW, int OH, int Bp, int cout, bool has_resid = false, int nz = 1) {
    const int npass = d.passes > 1 ? d.passes : 1;
    bool tall_ok = getenv("SR3_NO_TALL") == nullptr && OW >= 8 && OH >= 16, has3 = false;
    for (const KSlab& k : d.slabs) { if (k.p != 0) tall_ok = false; if (k.dh != 0) has3 = true; }
    tall_ok = tall_ok && has3 && (OH >= 32 || Bp % 2 == 0);
    const bool allow_split = getenv("SR3_NO_KSPLIT") == nullptr;
    const int sms = num_sms();
    // cost in bytes of the slowest CTA; `split` returns the factor the cost was computed for
    auto model = [&](long long tiles, int nstage, long long stage_bytes, int rows, int bn, int& split) -> double {
        int smax = 1;
        if (allow_split && bn >= 32 && cout % 32 == 0) {
            smax = (int)(sms / tiles);
            const int units = (rows / 128) * (bn / 32) * 4;
            if (smax > units) smax = units;
            if (smax > nstage / 2) smax = nstage / 2;
            if (smax > 16) smax = 16;
            if (smax < 1) smax = 1;
            smax = fit_split(smax, tiles);
        }
        split = smax;
        const long long waves = (tiles * smax + sms - 1) / sms;
        double c = (double)waves * ((nstage + smax - 1) / smax) * (double)stage_bytes;
        // a split tile: partial tile out (TMEM -> registers -> L2) and back, weighted 2x against streamed TMA bytes, plus ~1.4 us of
        // grid-level handshake; a residual is then read with plain loads instead of TMA
        if (smax > 1) c += 4.0 * rows * bn * 4 + 131072.0 + (has_resid ? 2.0 * rows * bn * 4 : 0.0);
        return c;
    };
    d.ksplit_max = 1;
    if (tall_ok) {
        d.tall = 1; d.w_box = 8;
        struct Cand { int mh, bn; };
        const Cand cands[4] = {{2, 128}, {2, 64}, {1, 64}, {1, 32}};
        auto geom = [&](int mh, int& h_box, int& b_box) {
            if (mh == 2 && OH >= 32) { h_box = 32; b_box = 1; }
            else if (mh == 2) { h_box = 16; b_box = 2; }
            else { h_box = 16; b_box = 1; }
        };
        // stages per tile: one per (source, 64-channel chunk, dw) group of vertical taps
        int nstage = 0;
        for (size_t i = 0; i < d.slabs.size(); ++i) {
            bool first = true;
            for (size_t j = 0; j < i; ++j)
                if (d.slabs[j].a_sel == d.slabs[i].a_sel && d.slabs[j].a_chan == d.slabs[i].a_chan && d.slabs[j].dw == d.slabs[i].dw) { first = false; break; }
            if (first) ++nstage;
        }
        int mh = 2, bn = 16, split = 1;               // Cout = 3 (final conv) keeps the 16-wide tile
        if (cout % 32 == 0) {
            double best = 1e300;
            for (int i = 0; i < 4; ++i) {
                const Cand& c = cands[i];
                if (cout % c.bn != 0) continue;
                int hb, bb; geom(c.mh, hb, bb);
                const long long tiles = (long long)(OW / 8) * (OH / hb) * (Bp / bb) * (cout / c.bn) * nz;
                const long long stage_bytes = (c.mh == 2 ? 36864 : 18432) + 3ll * c.bn * 128;
                int sp = 1;
                const double cost = model(tiles, nstage * npass, stage_bytes, c.mh * 128, c.bn, sp);
                // the residual is staged through smem (8 warps x 8 KB) unless the tile is split: a 256x128 tile would be left with one stage
                if (c.bn == 128 && has_resid && sp <= 1) continue;
                if (cost < best) { best = cost; mh = c.mh; bn = c.bn; split = sp; }
            }
        }
        if (const char* e = getenv("SR3_TALL_BN")) { int v = atoi(e); if ((v == 32 || v == 64 || v == 128) && cout % v == 0) { bn = v; split = 16; } }
        if (const char* e = getenv("SR3_TALL_MH")) { int v = atoi(e); if (v == 1 || v == 2) { mh = v; split = 16; } }
        if (mh == 2 && bn == 32) bn = 64;
        d.mh = mh; d.block_n = bn; d.ksplit_max = allow_split ? split : 1;
        geom(mh, d.h_box, d.b_box);
        d.a_box_w = 8; d.a_box_b = d.b_box;
        if (mh == 2 && d.b_box == 1) { d.a_box_h = 34; d.a_half_off = 16 * 1024; }
        else if (mh == 2) { d.a_box_h = 18; d.a_half_off = 18 * 1024; }
        else { d.a_box_h = 18; d.a_half_off = 0; }
    } else {
        d.tall = 0; d.mh = 1;
        pick_image_box(OW, OH, d.w_box, d.h_box, d.b_box);
        d.block_n = pick_block_n(cout);
        const long long mt = (long long)(OW / d.w_box) * (OH / d.h_box) * (Bp / d.b_box) * nz;
        if (getenv("SR3_BLOCK_N") == nullptr && cout % 32 == 0) {
            // generic stages group up to three K slabs (one A box + one B box each)
            const int nstage = (((int)d.slabs.size() + 2) / 3) * npass;
            double best = 1e300;
            for (int bn = 128; bn >= 32; bn >>= 1) {
                if (cout % bn != 0) continue;
                int sp = 1;
                const double cost = model(mt * (cout / bn), nstage, 3ll * (16384 + bn * 128), 128, bn, sp);
                if (bn == 128 && sp > 1) continue;      // 96 KB stages leave a 2-deep pipeline: measured slower than 64-wide split tiles
                if (cost < best) { best = cost; d.block_n = bn; d.ksplit_max = sp; }
            }
        }
    }
    if (const char* e = getenv("SR3_KSPLIT")) d.ksplit_max = atoi(e);
    d.tiles_w = OW / d.w_box; d.tiles_h = OH / d.h_box; d.tiles_b = Bp / d.b_box;
}

void pick_image_box(int W, int H, int& w_box, int& h_box, int& b_box) {
    w_box = W < 16 ? W : 16;
    h_box = 128 / w_box;
    if (h_box > H) h_box = H;
    b_box = 128 / (w_box * h_box);
}

int pick_block_n(int cout) {
    if (const char* e = getenv("SR3_BLOCK_N")) { int v = atoi(e); if (v == 32 || v == 64 || v == 128 || v == 256) if (cout % v == 0) return v; }
    if (cout % 128 == 0) return 128;
    if (cout % 64 == 0) return 64;
    return 16;
}

// `row` = channels per pixel of the source tensor (2 * cin in precise mode: [hi | lo]); only the stride-2 view needs it
void add_conv_slabs(std::vector<KSlab>& slabs, int a_sel, int cin, int ksize, int stride, int b_col0, int row = 0) {
    if (row == 0) row = cin;
    if (ksize == 1) {
        for (int c = 0; c < cin; c += 64) slabs.push_back({a_sel, c, 0, 0, 0, b_col0 + c});
        return;
    }
    for (int r = 0; r < 3; ++r)
        for (int s = 0; s < 3; ++s)
            for (int c = 0; c < cin; c += 64) {
                KSlab k; k.a_sel = a_sel; k.b_col = b_col0 + (r * 3 + s) * cin + c;
                if (stride == 1) { k.dh = r - 1; k.dw = s - 1; k.p = 0; k.a_chan = c; }
                else {   // input row 2*oh + r - 1, column 2*ow + s - 1 in the (2C, W/2, 2, H/2, B) view
                    k.dh = (r == 0) ? -1 : 0; k.p = (r == 1) ? 0 : 1;
                    k.dw = (s == 0) ? -1 : 0; k.a_chan = ((s == 1) ? 0 : row) + c;
                }
                slabs.push_back(k);
            }
}

// ------------------------------------------------------------------------------------------------ engine
struct Act {
    float* p = nullptr; double* stats = nullptr;
    int C = 0, H = 0, W = 0;
    // training plan only: gradient of the loss with respect to this tensor (fp32 NHWC), its bf16 copy (operand of the data / weight gradient
    // GEMMs) and its per-(image, channel) sums (bias gradients), all complete when the producing layer's backward runs
    float* g = nullptr; bf16* gb = nullptr; float* gsum = nullptr;
};

struct ParamEntry {
    std::string name;
    std::vector<int64_t> shape;
    int64_t numel = 0;
    std::function<void(const float*, cudaStream_t)> load;
    std::vector<std::function<void(const float*, cudaStream_t)>> hooks;   // training plan: further packed copies of the same tensor (data-gradient weights)
    bool loaded = false;
};

struct LayerSpec { std::string name; int kind; int cin, cout; bool attn; int res; };   // kind: 0 conv, 1 res, 2 down, 3 up

}  // namespace

struct sr3_engine {
    sr3_unet_config cfg{};
    int B = 0, Bp = 0, dev = 0;
    int H = 0, W = 0, inner = 0, cond_c = 0, in_C = 64;
    bool precise = false; int PW = 1;       // precise mode: every bf16 operand tensor is PW = 2 times as wide ([hi | lo] per pixel / row)
    DevAllocs mem;
    std::vector<ParamEntry> params;
    std::map<std::string, int> pindex;
    std::vector<Op> ops, finalize_ops;
    std::vector<GemmHandle> gemms;          // tile-kernel launches of the step, in order (for next-layer weight prefetch)
    struct OpInfo { int kind; double flops; double bytes; };   // kind: 0 gemm, 1 groupnorm-apply, 2 cast/upsample, 3 softmax, 4 other
    std::vector<OpInfo> op_info;
    std::map<std::string, Act> taps;
    std::map<std::string, size_t> role_max;
    std::map<std::string, void*> role_ptr;
    bool dry = true;
    // ---- training plan (sr3_engine_create_train): every scratch tensor of the forward is kept for the backward, which is recorded layer by
    // layer while the forward plan is built and replayed in reverse order
    bool train = false; float drop_p = 0.f;
    std::vector<Op>* bwd_sink = nullptr;                   // where push() records while a layer's backward is being described
    std::vector<std::vector<Op>> bwd_blocks;               // one op list per forward layer, executed last to first
    std::vector<std::vector<int>> bwd_kinds;               // op kinds (profiling): 0 data-gradient tile kernel, 1 GroupNorm / elementwise, 4 other, 6 weight gradient, 7 attention GEMMs
    // single-launch re-pack (bf16 precision): one PackDesc per packed copy, its source = parameter `pack_src[i]` (second source of a fused
    // bias: `pack_src2[i]`); device table rebuilt only when the parameter pointers change
    std::vector<PackDesc> pack_descs; std::vector<int> pack_src, pack_src2;
    PackDesc* pack_dev = nullptr; int* pack_ends_dev = nullptr; int pack_blocks = 0; std::vector<const float*> pack_last_ptrs;
    void add_pack(const std::string& pname, PackDesc d, const std::string& pname2 = "") {
        if (dry || precise) return;
        pack_descs.push_back(d); pack_src.push_back(pindex.at(pname)); pack_src2.push_back(pname2.empty() ? -1 : pindex.at(pname2));
    }
    std::vector<float*> grad_dst;                          // per parameter (state_dict order): where the running backward writes its gradient
    float gscale = 1.f;                                    // d(total) / d(summed loss) of the running backward (1 / (b c h w), model.py:50-53)
    float* zero_arena = nullptr; size_t zero_cap = 0, zero_used = 0;    // everything the backward accumulates into (cleared at its start)
    DropSpec* drop_dev = nullptr; std::vector<DropSpec> drop_host; std::vector<std::string> drop_names;
    bf16* last_xraw = nullptr;
    bf16* deps_b = nullptr; float* fin_bias_sum = nullptr; float* dfilm = nullptr; float* dtau = nullptr;
    float *dwf_all = nullptr, *dbf_all = nullptr, *dcb_all = nullptr;
    float *hr_buf = nullptr;
    int loss_type_cur = 1;

    StepCtl* ctl_dev = nullptr;
    StepCtl ctl{};
    double* stats_arena = nullptr; size_t stats_cap = 0, stats_used = 0;
    // persistent step kernel (one cooperative launch per reverse step)
    std::vector<MegaRec> mega;
    bool use_mega = false;
    MegaOp* mega_ops_dev = nullptr; uint8_t* mega_blob = nullptr; unsigned long long* mega_bar = nullptr; unsigned long long* mega_prof = nullptr;
    std::vector<int> mega_types;
    bf16* in_buf = nullptr;
    float *x_state = nullptr, *eps_buf = nullptr, *mean_buf = nullptr, *noise_buf = nullptr, *nl_buf = nullptr, *io_a = nullptr, *io_b = nullptr;
    float *nl_table = nullptr, *post_tab = nullptr;
    int T = 0, T_cap = 0;
    std::vector<float> logvar_host;
    double* loss_dev = nullptr;
    float *tau = nullptr, *film = nullptr, *film_w = nullptr, *film_b = nullptr, *film_cb = nullptr;
    float *mlp_w1 = nullptr, *mlp_b1 = nullptr, *mlp_w2 = nullptr, *mlp_b2 = nullptr;
    int F = 0;
    cudaGraphExec_t graph = nullptr;
    cudaStream_t cap_stream = nullptr;
    cudaStream_t side_stream = nullptr;            // graph capture only: the noise-level embedding + FiLM projections run beside the first conv
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int side_begin = -1, side_end = -1, side_join = -1;   // ops [side_begin, side_end) on the side branch, joined before op side_join
    bool use_graph = true;
    uint64_t seed = 0, first_index = 0;
    bool have_cond = false;

    ~sr3_engine() {
        if (graph) cudaGraphExecDestroy(graph);
        if (cap_stream) cudaStreamDestroy(cap_stream);
        if (side_stream) cudaStreamDestroy(side_stream);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
    }

    // ---- buffers shared between layers of the same role (stream order makes reuse safe)
    void* role(const std::string& r, size_t bytes) {
        // training plan: forward scratch is read again by the backward -> one allocation per use ("g_*" = backward scratch stays shared)
        if (train && r.compare(0, 2, "g_") != 0) {
            if (dry) return reinterpret_cast<void*>(0x1000);
            return mem.alloc(bytes);
        }
        if (dry) { size_t& m = role_max[r]; if (bytes > m) m = bytes; return reinterpret_cast<void*>(0x1000); }
        REQUIRE(role_ptr.count(r) && role_max[r] >= bytes, "role buffer %s too small", r.c_str());
        return role_ptr[r];
    }
    double* new_stats(int C) {
        const size_t n = (size_t)Bp * C * 2;
        if (dry) { stats_used += n; return nullptr; }
        REQUIRE(stats_used + n <= stats_cap, "stats arena overflow");
        double* p = stats_arena + stats_used;
        stats_used += n;
        return p;
    }
    Act new_act(int C, int Hh, int Ww, const std::string& tap_name = "") {
        Act a; a.C = C; a.H = Hh; a.W = Ww;
        a.stats = new_stats(C);
        if (train) a.gsum = new_zero((size_t)Bp * C);
        if (!dry) {
            a.p = static_cast<float*>(mem.alloc((size_t)Bp * Hh * Ww * C * sizeof(float)));
            if (train) {
                a.g = static_cast<float*>(mem.alloc((size_t)Bp * Hh * Ww * C * sizeof(float)));
                a.gb = static_cast<bf16*>(mem.alloc((size_t)Bp * Hh * Ww * C * sizeof(bf16)));
            }
            if (!tap_name.empty()) taps[tap_name] = a;
        }
        return a;
    }
    // floats from the arena the backward clears before it starts (sums it accumulates into with atomics)
    float* new_zero(size_t n) {
        n = (n + 3) & ~size_t(3);
        if (dry) { zero_used += n; return nullptr; }
        REQUIRE(zero_used + n <= zero_cap, "zero arena overflow");
        float* p = zero_arena + zero_used;
        zero_used += n;
        return p;
    }
    void add_param_hook(const std::string& name, std::function<void(const float*, cudaStream_t)> fn) {
        if (dry) return;
        params[pindex.at(name)].hooks.push_back(std::move(fn));
    }
    void add_param(const std::string& name, std::vector<int64_t> shape, std::function<void(const float*, cudaStream_t)> load) {
        if (dry) return;
        ParamEntry e; e.name = name; e.shape = shape; e.numel = 1;
        for (auto s : shape) e.numel *= s;
        e.load = std::move(load);
        pindex[name] = (int)params.size();
        params.push_back(std::move(e));
    }
    float* f32_param(const std::string& name, std::vector<int64_t> shape) {
        if (dry) return nullptr;
        int64_t n = 1; for (auto s : shape) n *= s;
        float* dst = static_cast<float*>(mem.alloc(n * sizeof(float)));
        add_param(name, shape, [dst, n](const float* src, cudaStream_t st) { CK(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st)); });
        { PackDesc d{}; d.type = 0; d.dst = dst; d.n = n; add_pack(name, d); }
        return dst;
    }
    // f32 parameter stored into a slice of a bigger array
    void f32_param_into(const std::string& name, std::vector<int64_t> shape, float* dst) {
        if (dry) return;
        int64_t n = 1; for (auto s : shape) n *= s;
        add_param(name, shape, [dst, n](const float* src, cudaStream_t st) { CK(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st)); });
        { PackDesc d{}; d.type = 0; d.dst = dst; d.n = n; add_pack(name, d); }
    }
    // conv weight packed into rows [0,Cout) of a [rows_pad][ktot] bf16 matrix at column k_off
    void conv_weight_param(const std::string& name, bf16* dst, int Cout, int Cin, int k, int ktot, int k_off, int cin_pad) {
        if (dry) return;
        const int ld = PW * ktot, lo_off = precise ? ktot : 0;
        add_param(name, {Cout, Cin, k, k}, [=](const float* src, cudaStream_t st) {
            const long long total = 1LL * Cout * Cin;
            const int blocks = (int)std::min<long long>((total + 255) / 256, 4096);
            pack_conv_weight_kernel<<<blocks, 256, 0, st>>>(src, dst, Cout, Cin, k, k, ld, k_off, cin_pad, lo_off);
            CK(cudaGetLastError());
        });
        { PackDesc d{}; d.type = 1; d.dst = dst; d.Cout = Cout; d.Cin = Cin; d.k = k; d.ld = ld; d.k_off = k_off; d.cin_pad = cin_pad; add_pack(name, d); }
    }
    bf16* new_weight(int rows, int ktot, int block_n) {      // [rows_pad][PW * ktot]: precise mode appends the low halves of every row
        if (dry) return nullptr;
        (void)block_n;
        const int rows_pad = ((rows + 127) / 128) * 128;
        return static_cast<bf16*>(mem.alloc((size_t)rows_pad * PW * ktot * sizeof(bf16)));
    }
    // precise-mode fields of an image conv whose A sources have c0 (c1) channels and whose weight rows hold ktot (high) columns
    void set_precise(GemmDesc& d, int c0, int c1, int ktot) {
        if (!precise) return;
        d.passes = 3; d.lo_a_chan[0] = c0; d.lo_a_chan[1] = c1; d.lo_b_col = ktot;
    }
    void push(Op op, int kind = 4, double flops = 0, double bytes = 0) {
        if (dry) return;
        if (bwd_sink) { bwd_sink->push_back(std::move(op)); bwd_kinds.back().push_back(kind); return; }
        ops.push_back(std::move(op));
        op_info.push_back({kind, flops, bytes});
    }
    // executed work of a gemm op: 2*M*N*K flops; bytes = A read once per tap set + B once + outputs
    void push_gemm(const GemmDesc& d) {
        const double M = (double)d.OW * d.OH * d.OB * (d.nz > 1 && d.a_zstep == 0 ? d.nz : 1);
        const double K = 64.0 * d.slabs.size() * (d.passes > 1 ? d.passes : 1);      // executed MACs (precise mode: three passes)
        const double N = d.n_valid;
        double bytes = N * K * 2;
        if (d.out_f32) bytes += M * N * 4;
        if (d.out_bf16) bytes += M * N * 2;
        if (d.resid) bytes += M * N * 4;
        double a_elems = 0;
        for (int i = 0; i < d.n_a; ++i) a_elems += (double)d.a[i].C * d.a[i].W * d.a[i].P * d.a[i].H * d.a[i].Bn;
        bytes += a_elems * 2;
        push(make_gemm_op(d, mem), 0, 2.0 * M * N * K, bytes);
    }

    // ---- layer builders -------------------------------------------------------------------------
    float* last_mr = nullptr;              // training plan: (mean, rstd) buffer written by the most recent add_prep, read by its backward
    void add_prep(const Act& s0, const Act* s1, const float* gamma, const float* beta, int groups, bool silu, bf16* out_a, bf16* out_raw, const DropSpec* drop = nullptr) {
        if (dry) return;
        PrepParams p{};
        p.drop = drop;
        if (train) { last_mr = static_cast<float*>(mem.alloc((size_t)Bp * groups * 2 * sizeof(float))); p.save_mr = last_mr; }
        p.src0 = s0.p; p.st0 = s0.stats; p.C0 = s0.C;
        p.src1 = s1 ? s1->p : nullptr; p.st1 = s1 ? s1->stats : nullptr; p.C1 = s1 ? s1->C : 0;
        p.gamma = gamma; p.beta = beta; p.groups = groups; p.HW = s0.H * s0.W; p.silu = silu ? 1 : 0; p.eps = 1e-5f;
        p.out_a = out_a; p.out_raw = out_raw; p.precise = precise ? 1 : 0;
        const int C = p.C0 + p.C1;
        REQUIRE(C % groups == 0 && C % 4 == 0 && p.C0 % 4 == 0, "bad GroupNorm geometry C=%d groups=%d", C, groups);
        const int vpp = C / 4;
        REQUIRE(vpp <= 512, "GroupNorm over %d channels is not supported", C);
        const int kpix = vpp >= 256 ? 1 : 256 / vpp;
        const int threads = vpp * kpix;                                // <= 512, every thread owns one 4-channel column
        // ONE wave of blocks: the kernel uses 64 registers per thread, i.e. 4 resident 256-thread blocks (2 of 512) per SM; a grid sized
        // for 8 per SM (round 1, 32-register version) runs as two waves and pays the statistics set-up twice (24 vs 15 us on the 128x128
        // level).  SR3_PREP_SLOTS overrides the number of resident blocks per SM assumed here.
        int ppb;
        {
            const int per_sm = getenv("SR3_PREP_SLOTS") ? atoi(getenv("SR3_PREP_SLOTS")) : (threads > 256 ? 2 : 4);
            int bpi = per_sm * num_sms() / B; if (bpi < 1) bpi = 1;    // blocks per image
            const int q = kpix * 4;                                    // 4 loads in flight per thread
            ppb = (p.HW + bpi - 1) / bpi;
            ppb = ((ppb + q - 1) / q) * q;
            if (ppb < q) ppb = q;
            if (ppb > p.HW) ppb = p.HW;
        }
        p.pix_per_block = ppb;
        const dim3 grid((p.HW + ppb - 1) / ppb, B);
        const int smem = (2 * C + 2 * groups) * sizeof(float);
        {   // the same op inside the persistent step kernel: B x items_per_image work items dealt contiguously to the CTAs
            PrepParams m = p;
            const int nth = GEMM_THREADS;
            const int kp = vpp > nth ? 1 : nth / vpp;
            int ipi = (4 * num_sms() + B / 2) / B; if (ipi < 1) ipi = 1;
            int mp = (p.HW + ipi - 1) / ipi;
            mp = ((mp + kp - 1) / kp) * kp;
            if (mp > p.HW) mp = p.HW;
            m.pix_per_block = mp; m.items_per_image = (p.HW + mp - 1) / mp; m.B = B;
            mega_record(MOP_PREP, m);
        }
        push([p, grid, smem, threads](cudaStream_t st) {
            if (p.drop) launch_k(prep_kernel<true>, grid, dim3(threads), (size_t)smem, st, p);
            else launch_k(prep_kernel<false>, grid, dim3(threads), (size_t)smem, st, p);
        }, 1, 0, (double)B * p.HW * C * (4.0 + 2.0 + (out_raw ? 2.0 : 0.0)));
    }
    void add_cast(const Act& s, bf16* dst, int up) {
        if (dry) return;
        const long long total = 1LL * B * s.H * up * s.W * up * (s.C / 4);
        const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
        const float* src = s.p; const int Bn = B, Hh = s.H, Ww = s.W, C = s.C;
        push([=](cudaStream_t st) { launch_k(cast_kernel, dim3(blocks), dim3(256), 0, st, src, dst, Bn, Hh, Ww, C, up); }, 2, 0, (double)Bn * Hh * Ww * C * (4.0 + 2.0 * up * up));
    }

    // generic image conv: A sources already bf16; out fp32 NHWC (+stats)
    struct ConvArgs {
        ASrc a[2]; int n_a = 1;
        std::vector<KSlab> slabs;
        const bf16* w = nullptr; int ktot = 0; int cout = 0;
        int OH = 0, OW = 0;
        const float* bias = nullptr; const float* bias2 = nullptr; int bias2_stride = 0;
        const float* resid = nullptr;
        Act out;
        bf16* raw_out = nullptr;          // also store bf16(out) (input of a following Down / Upsample conv): no separate cast pass
        bool custom_os = false; OutSpec os{};   // output addressing other than plain NHWC (phase of a folded upsample conv)
        int nz = 1, b_zrows = 0, z_phase = 0; long long z_off_hi = 0, z_off_lo = 0;   // the four phases of a folded upsample conv in ONE launch
        int c0 = 0, c1 = 0;               // channels of the A sources (precise mode: where their low halves start)
    };
    void add_conv(const ConvArgs& c) {
        if (dry) return;
        GemmDesc d;
        d.n_a = c.n_a; d.a[0] = c.a[0]; d.a[1] = c.a[1];
        d.slabs = c.slabs;
        REQUIRE(!precise || c.c0 > 0, "precise mode: conv without source channel counts");
        set_precise(d, c.c0, c.c1, c.ktot);
        conv_geometry(d, c.OW, c.OH, Bp, c.cout, c.resid != nullptr, c.nz);
        d.b_ptr = c.w; d.b_K = PW * c.ktot; d.b_rows = (long long)c.nz * (((c.cout + 127) / 128) * 128);      // weights are padded to 128 rows (new_weight)
        d.b_is_param = true;
        d.n_tiles = (c.cout + d.block_n - 1) / d.block_n; d.nz = c.nz; d.a_zstep = 0; d.b_zrows = c.b_zrows;
        d.z_phase = c.z_phase; d.z_off_hi = c.z_off_hi; d.z_off_lo = c.z_off_lo;
        d.OW = c.OW; d.OH = c.OH; d.OB = B; d.n_valid = c.cout;
        d.bias = c.bias; d.bias2 = c.bias2; d.bias2_stride = c.bias2_stride;
        d.resid = c.resid; d.rs = nhwc_out(c.OH, c.OW, c.cout);
        d.out_f32 = c.out.p; d.os = c.custom_os ? c.os : nhwc_out(c.OH, c.OW, c.cout);
        if (c.raw_out) { d.out_bf16 = c.raw_out; d.hs = nhwc_out(c.OH, c.OW, PW * c.cout); d.lo_out_off = precise ? c.cout : 0; }
        d.stats = c.out.stats; d.stats_C = c.cout; d.stats_coff = 0;
        push_gemm(d);
    }

    // ResnetBlock (+ optional SelfAttention): reference unet.py:94-158
    Act add_res_block(const LayerSpec& L, const Act& x, const Act* skip, int& film_off, bf16* raw_out = nullptr, bool x_has_skip = false) {
        const int cin = x.C + (skip ? skip->C : 0), cout = L.cout, Hh = x.H, Ww = x.W, G = cfg.norm_groups;
        REQUIRE(cin == L.cin, "%s: cin mismatch %d vs %d", L.name.c_str(), cin, L.cin);
        const std::string p = L.name + ".res_block";
        const bool has_res = cin != cout;
        // parameters in the reference's registration order
        const int foff = film_off; film_off += cout;
        f32_param_into(p + ".noise_func.noise_func.0.weight", {cout, inner}, dry ? nullptr : film_w + (size_t)foff * inner);
        f32_param_into(p + ".noise_func.noise_func.0.bias", {cout}, dry ? nullptr : film_b + foff);
        float* g1 = f32_param(p + ".block1.block.0.weight", {cin});
        float* b1 = f32_param(p + ".block1.block.0.bias", {cin});
        bf16* w1 = new_weight(cout, 9 * cin, pick_block_n(cout));
        conv_weight_param(p + ".block1.block.3.weight", w1, cout, cin, 3, 9 * cin, 0, cin);
        f32_param_into(p + ".block1.block.3.bias", {cout}, dry ? nullptr : film_cb + foff);
        float* g2 = f32_param(p + ".block2.block.0.weight", {cout});
        float* b2 = f32_param(p + ".block2.block.0.bias", {cout});
        const int k2 = 9 * cout + (has_res ? cin : 0);
        bf16* w2 = new_weight(cout, k2, pick_block_n(cout));
        conv_weight_param(p + ".block2.block.3.weight", w2, cout, cout, 3, k2, 0, cout);
        float* cb2 = f32_param(p + ".block2.block.3.bias", {cout});
        float* bias_total = cb2;
        if (has_res) {
            conv_weight_param(p + ".res_conv.weight", w2, cout, cin, 1, k2, 9 * cout, cin);
            float* cbr = f32_param(p + ".res_conv.bias", {cout});
            if (!dry) {
                bias_total = static_cast<float*>(mem.alloc(cout * sizeof(float)));
                float* bt = bias_total;
                finalize_ops.push_back([=](cudaStream_t st) { add_vec_kernel<<<(cout + 255) / 256, 256, 0, st>>>(cb2, cbr, bt, cout); CK(cudaGetLastError()); });
                { PackDesc d{}; d.type = 6; d.dst = bt; d.n = cout; add_pack(p + ".block2.block.3.bias", d, p + ".res_conv.bias"); }
            }
        }
        // scratch
        bf16* a1 = static_cast<bf16*>(role("a1", (size_t)Bp * Hh * Ww * cin * 2 * PW));
        bf16* raw = has_res ? static_cast<bf16*>(role("raw", (size_t)Bp * Hh * Ww * cin * 2 * PW)) : nullptr;
        bf16* a2 = static_cast<bf16*>(role("a2", (size_t)Bp * Hh * Ww * cout * 2 * PW));
        Act h; h.C = cout; h.H = Hh; h.W = Ww; h.stats = new_stats(cout);
        h.p = static_cast<float*>(role("h", (size_t)Bp * Hh * Ww * cout * 4));
        Act y = new_act(cout, Hh, Ww, L.attn ? "" : L.name);

        add_prep(x, skip, g1, b1, G, true, a1, raw);
        float* mr1 = last_mr;
        {
            ConvArgs c; c.n_a = 1; c.a[0] = nhwc_src(a1, Bp, Hh, Ww, cin * PW); c.c0 = cin;
            add_conv_slabs(c.slabs, 0, cin, 3, 1, 0);
            c.w = w1; c.ktot = 9 * cin; c.cout = cout; c.OH = Hh; c.OW = Ww;
            c.bias2 = dry ? nullptr : film + foff; c.bias2_stride = F;
            c.out = h;
            add_conv(c);
        }
        const DropSpec* drop = nullptr;
        if (train && drop_p > 0.f) {          // Dropout sits in block2 only (unet.py:100-101)
            if (!dry) {
                REQUIRE(drop_host.size() < 256, "too many dropout layers");
                DropSpec ds{}; ds.mask = nullptr; ds.p = drop_p; ds.layer = (unsigned)drop_host.size(); ds.seed = 0;
                drop = drop_dev + drop_host.size();
                drop_host.push_back(ds); drop_names.push_back(p + ".block2");
            }
        }
        add_prep(h, nullptr, g2, b2, G, true, a2, nullptr, drop);
        float* mr2 = last_mr;
        {
            ConvArgs c; c.n_a = has_res ? 2 : 1; c.a[0] = nhwc_src(a2, Bp, Hh, Ww, cout * PW); c.c0 = cout;
            add_conv_slabs(c.slabs, 0, cout, 3, 1, 0);
            if (has_res) { c.a[1] = nhwc_src(raw, Bp, Hh, Ww, cin * PW); c.c1 = cin; add_conv_slabs(c.slabs, 1, cin, 1, 1, 9 * cout); }
            c.w = w2; c.ktot = k2; c.cout = cout; c.OH = Hh; c.OW = Ww;
            c.bias = bias_total; c.resid = has_res ? nullptr : x.p;
            c.out = y;
            if (!L.attn) c.raw_out = raw_out;
            add_conv(c);
        }
        if (train) {
            if (!dry) film_slices.push_back({foff, cout, pid(p + ".noise_func.noise_func.0.weight"), pid(p + ".noise_func.noise_func.0.bias"), pid(p + ".block1.block.3.bias")});
            ResBwdCtx c; c.p = p; c.x = x; c.skip = skip; c.h = h; c.y = y; c.cin = cin; c.cout = cout; c.Hh = Hh; c.Ww = Ww; c.foff = foff;
            c.has_res = has_res; c.x_acc = x_has_skip; c.a1 = a1; c.raw = raw; c.a2 = a2; c.g1 = g1; c.b1 = b1; c.g2 = g2; c.b2 = b2; c.drop = drop; c.mr1 = mr1; c.mr2 = mr2;
            bwd_res_block(c);
        }
        return L.attn ? add_attention(L, y, raw_out) : y;      // (its backward is recorded after the block's: it runs first)
    }

    // SelfAttention (reference unet.py:113-142): GN -> qkv 1x1 (no bias) -> softmax(q k^T / sqrt(C)) v -> out 1x1 + bias + x
    Act add_attention(const LayerSpec& L, const Act& x, bf16* raw_out = nullptr) {
        const int C = x.C, Hh = x.H, Ww = x.W, HW = Hh * Ww, G = cfg.norm_groups;
        const std::string p = L.name + ".attn";
        const int Lt = HW >= 128 ? HW : 128;            // tokens per attention batch (two 8x8 images share one)
        const int per = Lt / HW;                         // images per attention batch
        REQUIRE(Lt % 128 == 0 && (Bp % per) == 0, "attention geometry HW=%d", HW);
        REQUIRE(C % 128 == 0, "%s: self-attention over %d channels is not supported (the q/k/v and P.v tiles are 128 columns wide; C must be a multiple of 128)", L.name.c_str(), C);
        const int nz = Bp / per;
        float* gn_w = f32_param(p + ".norm.weight", {C});
        float* gn_b = f32_param(p + ".norm.bias", {C});
        bf16* wqkv = new_weight(3 * C, C, 128);
        conv_weight_param(p + ".qkv.weight", wqkv, 3 * C, C, 1, C, 0, C);
        bf16* wout = new_weight(C, C, 128);
        conv_weight_param(p + ".out.weight", wout, C, C, 1, C, 0, C);
        float* bout = f32_param(p + ".out.bias", {C});
        bf16* n = static_cast<bf16*>(role("a1", (size_t)Bp * HW * C * 2 * PW));
        bf16* qk = static_cast<bf16*>(role("qk", (size_t)Bp * HW * 2 * C * 2 * PW));
        bf16* vT = static_cast<bf16*>(role("vT", (size_t)nz * C * Lt * 2 * PW));
        float* S = static_cast<float*>(role("S", (size_t)nz * Lt * Lt * 4));
        bf16* P = static_cast<bf16*>(role("P", (size_t)nz * Lt * Lt * 2 * PW));
        bf16* O = static_cast<bf16*>(role("O", (size_t)Bp * HW * C * 2 * PW));
        Act y = new_act(C, Hh, Ww, L.name);
        add_prep(x, nullptr, gn_w, gn_b, G, false, n, nullptr);
        float* mr_attn = last_mr;
        if (dry) {
            if (train) { AttnBwdCtx c; c.p = p; c.x = x; c.y = y; c.C = C; c.Hh = Hh; c.Ww = Ww; c.Lt = Lt; c.per = per; c.nz = nz; c.n = n; c.qk = qk; c.vT = vT; c.P = P; c.O = O; c.gn_w = gn_w; c.gn_b = gn_b; c.mr = mr_attn; bwd_attention(c); }
            return y;
        }
        const bool merged_qkv = (getenv("SR3_NO_MERGED_QKV") == nullptr || precise) && C % 128 == 0;     // whole 128-column tiles on either side of 2C
        {   // q,k (,v) = Wqkv n : [Bp*HW tokens] x [2C (3C)]; the v columns are stored transposed as vT[z][d][token]
            GemmDesc d; d.n_a = 1; d.a[0] = nhwc_src(n, Bp, Hh, Ww, C * PW);
            add_conv_slabs(d.slabs, 0, C, 1, 1, 0);
            set_precise(d, C, 0, C);
            const int ncol = merged_qkv ? 3 * C : 2 * C;
            d.block_n = 128; d.b_ptr = wqkv; d.b_K = C * PW; d.b_rows = ncol; d.b_is_param = true;
            pick_image_box(Ww, Hh, d.w_box, d.h_box, d.b_box);
            d.tiles_w = Ww / d.w_box; d.tiles_h = Hh / d.h_box; d.tiles_b = Bp / d.b_box; d.n_tiles = ncol / 128;
            d.OW = Ww; d.OH = Hh; d.OB = Bp; d.n_valid = ncol;
            d.out_bf16 = qk; d.hs = nhwc_out(Hh, Ww, 2 * C * PW); d.lo_out_off = precise ? 2 * C : 0;      // rows [q_hi | k_hi | q_lo | k_lo]
            if (merged_qkv) { d.out_t = vT; d.t_col0 = 2 * C; d.t_rows = C; d.t_ld = Lt * PW; d.t_per = per; d.lo_t_off = precise ? Lt : 0; }
            push_gemm(d);
        }
        if (!merged_qkv) {   // vT[z][d][token] = Wv[d,:] . n[token,:]  (weights are the A operand, tokens the B operand)
            GemmDesc d; d.n_a = 1; d.a[0] = matrix_src(wqkv + (size_t)2 * C * C, 1, C, C, C, 0);
            for (int c = 0; c < C; c += 64) d.slabs.push_back({0, c, 0, 0, 0, c});
            d.block_n = 128; d.b_ptr = n; d.b_K = C; d.b_rows = (long long)Bp * HW;
            d.w_box = 128; d.h_box = 1; d.b_box = 1; d.tiles_w = C / 128; d.tiles_h = 1; d.tiles_b = 1;
            d.n_tiles = Lt / 128; d.nz = nz; d.a_zstep = 0; d.b_zrows = Lt;
            d.OW = C; d.OH = 1; d.OB = 1; d.n_valid = Lt;
            d.out_bf16 = vT; d.hs = OutSpec{(long long)C * Lt, 0, 0, Lt, 0};
            push_gemm(d);
        }
        if (attn_fusable(Lt, C) && !precise && !train) {
            // S = q k^T / sqrt(C), softmax over the keys of the same image, O = P v: one launch (attn_tcgen05.cuh)
            const double fl = 4.0 * nz * (double)Lt * Lt * C;
            push(make_attn_op(qk, vT, O, nz, Lt, HW, C), 5, fl, (double)nz * Lt * C * 2 * 4);
        } else {
            {   // S[z] = q k^T / sqrt(C)
                GemmDesc d; d.n_a = 1; d.a[0] = matrix_src(qk, nz, Lt, 2 * C * PW, 2 * C * PW, (long long)Lt * 2 * C * PW);
                for (int c = 0; c < C; c += 64) d.slabs.push_back({0, c, 0, 0, 0, C + c});
                set_precise(d, 2 * C, 0, 2 * C);
                d.block_n = 128; d.b_ptr = qk; d.b_K = 2 * C * PW; d.b_rows = (long long)Bp * HW;
                d.w_box = 128; d.h_box = 1; d.b_box = 1; d.tiles_w = Lt / 128; d.tiles_h = 1; d.tiles_b = 1;
                d.n_tiles = Lt / 128; d.nz = nz; d.a_zstep = 1; d.b_zrows = Lt;
                d.OW = Lt; d.OH = 1; d.OB = nz; d.n_valid = Lt; d.scale = 1.0f / sqrtf((float)C);
                d.out_f32 = S; d.os = OutSpec{0, (long long)Lt * Lt, 0, Lt, 0};
                push_gemm(d);
            }
            {
                const long long rows = (long long)nz * Lt;
                const int blocks = (int)((rows + 7) / 8);
                const int prec = precise ? 1 : 0;
                push([=](cudaStream_t st) { launch_k(softmax_kernel, dim3(blocks), dim3(256), 0, st, (const float*)S, P, rows, Lt, HW, prec); }, 3, 0, (double)rows * Lt * 6.0);
                SoftmaxParams sp{}; sp.S = S; sp.P = P; sp.rows = rows; sp.L = Lt; sp.seg = HW; sp.precise = prec;
                mega_record(MOP_SOFTMAX, sp);
            }
            {   // O[z] = P v : rows = queries, N = head dim, K = keys
                GemmDesc d; d.n_a = 1; d.a[0] = matrix_src(P, nz, Lt, Lt * PW, Lt * PW, (long long)Lt * Lt * PW);
                for (int c = 0; c < Lt; c += 64) d.slabs.push_back({0, c, 0, 0, 0, c});
                set_precise(d, Lt, 0, Lt);
                d.block_n = 128; d.b_ptr = vT; d.b_K = Lt * PW; d.b_rows = (long long)nz * C;
                d.w_box = 128; d.h_box = 1; d.b_box = 1; d.tiles_w = Lt / 128; d.tiles_h = 1; d.tiles_b = 1;
                d.n_tiles = C / 128; d.nz = nz; d.a_zstep = 1; d.b_zrows = C;
                d.OW = Lt; d.OH = 1; d.OB = nz; d.n_valid = C;
                d.out_bf16 = O; d.hs = OutSpec{0, (long long)Lt * C * PW, 0, (long long)C * PW, 0}; d.lo_out_off = precise ? C : 0;
                push_gemm(d);
            }
        }
        {   // out projection + bias + residual (un-normalised input)
            ConvArgs c; c.n_a = 1; c.a[0] = nhwc_src(O, Bp, Hh, Ww, C * PW); c.c0 = C;
            add_conv_slabs(c.slabs, 0, C, 1, 1, 0);
            c.w = wout; c.ktot = C; c.cout = C; c.OH = Hh; c.OW = Ww; c.bias = bout; c.resid = x.p; c.out = y; c.raw_out = raw_out;
            add_conv(c);
        }
        if (train) {
            AttnBwdCtx c; c.p = p; c.x = x; c.y = y; c.C = C; c.Hh = Hh; c.Ww = Ww; c.Lt = Lt; c.per = per; c.nz = nz; c.n = n; c.qk = qk; c.vT = vT; c.P = P; c.O = O;
            c.gn_w = gn_w; c.gn_b = gn_b; c.mr = mr_attn;
            bwd_attention(c);
        }
        return y;
    }

#include "train_plan.inc"

    void build_plan() {
        // topology: reference unet.py:186-231
        std::vector<LayerSpec> downs, mid, ups;
        std::vector<int> feat;
        int pre = inner, res = cfg.image_size;
        feat.push_back(pre);
        downs.push_back({"downs.0", 0, cfg.in_channel, inner, false, res});
        for (int ind = 0; ind < cfg.n_mults; ++ind) {
            const bool last = ind == cfg.n_mults - 1;
            bool use_attn = false;
            for (int i = 0; i < cfg.n_attn_res; ++i) use_attn |= (cfg.attn_res[i] == res);
            const int ch = inner * cfg.channel_mults[ind];
            for (int r = 0; r < cfg.res_blocks; ++r) {
                downs.push_back({"downs." + std::to_string(downs.size()), 1, pre, ch, use_attn, res});
                feat.push_back(ch); pre = ch;
            }
            if (!last) { downs.push_back({"downs." + std::to_string(downs.size()), 2, pre, pre, false, res}); feat.push_back(pre); res /= 2; }
        }
        mid.push_back({"mid.0", 1, pre, pre, true, res});
        mid.push_back({"mid.1", 1, pre, pre, false, res});
        for (int ind = cfg.n_mults - 1; ind >= 0; --ind) {
            const bool last = ind < 1;
            bool use_attn = false;
            for (int i = 0; i < cfg.n_attn_res; ++i) use_attn |= (cfg.attn_res[i] == res);
            const int ch = inner * cfg.channel_mults[ind];
            for (int r = 0; r < cfg.res_blocks + 1; ++r) {
                ups.push_back({"ups." + std::to_string(ups.size()), 1, pre + feat.back(), ch, use_attn, res});
                feat.pop_back(); pre = ch;
            }
            if (!last) { ups.push_back({"ups." + std::to_string(ups.size()), 3, pre, pre, false, res}); res *= 2; }
        }
        int F_total = 0;
        for (auto* v : {&downs, &mid, &ups}) for (auto& L : *v) if (L.kind == 1) F_total += L.cout;
        F = F_total;
        int min_res = cfg.image_size;
        for (int i = 1; i < cfg.n_mults; ++i) min_res /= 2;
        REQUIRE(min_res >= 8, "lowest UNet resolution %d < 8 is not supported", min_res);
        if (train) {
            fin_bias_sum = new_zero(4); dfilm = new_zero((size_t)Bp * F); dtau = new_zero((size_t)Bp * inner);
            if (!dry) {
                deps_b = static_cast<bf16*>(mem.alloc((size_t)Bp * H * W * 64 * sizeof(bf16)));
                dwf_all = static_cast<float*>(mem.alloc((size_t)F * inner * 4)); dbf_all = static_cast<float*>(mem.alloc((size_t)F * 4));
                dcb_all = static_cast<float*>(mem.alloc((size_t)F * 4));
                drop_dev = static_cast<DropSpec*>(mem.alloc(256 * sizeof(DropSpec)));
            }
        }

        if (!dry) {
            mlp_w1 = f32_param("noise_level_mlp.1.weight", {4 * inner, inner});
            mlp_b1 = f32_param("noise_level_mlp.1.bias", {4 * inner});
            mlp_w2 = f32_param("noise_level_mlp.3.weight", {inner, 4 * inner});
            mlp_b2 = f32_param("noise_level_mlp.3.bias", {inner});
            film_w = static_cast<float*>(mem.alloc((size_t)F * inner * 4));
            film_b = static_cast<float*>(mem.alloc((size_t)F * 4));
            film_cb = static_cast<float*>(mem.alloc((size_t)F * 4));
            film = static_cast<float*>(mem.alloc((size_t)Bp * F * 4));
            tau = static_cast<float*>(mem.alloc((size_t)Bp * inner * 4));
            // step prologue
            float4* sa = reinterpret_cast<float4*>(stats_arena);
            const long long n4 = (long long)(stats_cap * sizeof(double) / 16);
            StepCtl* c = ctl_dev;
            push([=](cudaStream_t st) { launch_k(step_begin_kernel, dim3((int)std::min<long long>((n4 + 255) / 256, 592)), dim3(256), 0, st, sa, n4, c); });
            EmbedParams ep{}; ep.ctl = ctl_dev; ep.nl_table = nl_table; ep.nl_buf = nl_buf; ep.w1 = mlp_w1; ep.b1 = mlp_b1; ep.w2 = mlp_w2; ep.b2 = mlp_b2;
            ep.tau = tau; ep.inner = inner;
            const int Bn = B; const int esm = 5 * inner * 4;
            side_begin = (int)ops.size();
            push([=](cudaStream_t st) { launch_k(embed_kernel, dim3(Bn), dim3(256), (size_t)esm, st, ep); });
            float *fw = film_w, *fb = film_b, *fc = film_cb, *ta = tau, *fi = film; const int Fn = F, inn = inner;
            push([=](cudaStream_t st) { launch_k(film_kernel, dim3((Fn + 63) / 64), dim3(256), (size_t)((64 * (inn + 1) + Bn * inn) * 4), st, (const float*)fw, (const float*)fb, (const float*)fc, (const float*)ta, fi, Fn, inn, Bn); });
            side_end = (int)ops.size();
            EmbedFilmParams fp{}; fp.e = ep; fp.wf = film_w; fp.bf = film_b; fp.cbias = film_cb; fp.film = film; fp.F = F; fp.B = B;
            mega_record(MOP_EMBED_FILM, fp);
        }

        int film_off = 0;
        std::vector<Act> feats;
        Act x;
        const bool fuse_cast = getenv("SR3_NO_FUSE_CAST") == nullptr;     // producers also emit the bf16 copy Down / Upsample convs read
        const bool fold_up = getenv("SR3_NO_FOLD_UP") == nullptr;         // nearest-2x + conv3x3 as four 2x2-tap phase convs on the low-res input
        for (size_t li = 0; li < downs.size(); ++li) {
            auto& L = downs[li];
            const bool next_is_down = li + 1 < downs.size() && downs[li + 1].kind == 2;
            if (L.kind == 0) {          // first conv on the (zero-padded to 64 ch) input buffer
                bf16* w = new_weight(inner, 9 * in_C, pick_block_n(inner));
                conv_weight_param(L.name + ".weight", w, inner, cfg.in_channel, 3, 9 * in_C, 0, in_C);
                float* b = f32_param(L.name + ".bias", {inner});
                x = new_act(inner, H, W, L.name);
                ConvArgs c; c.n_a = 1; c.a[0] = nhwc_src(in_buf, Bp, H, W, in_C * PW); c.c0 = in_C;
                add_conv_slabs(c.slabs, 0, in_C, 3, 1, 0);
                c.w = w; c.ktot = 9 * in_C; c.cout = inner; c.OH = H; c.OW = W; c.bias = b; c.out = x;
                add_conv(c);
                if (train) bwd_first_conv(L.name, x);
                if (!dry) side_join = (int)ops.size();      // the FiLM biases are first read by the next block's conv1 epilogue
            } else if (L.kind == 1) {
                bf16* xr = (fuse_cast && next_is_down) ? static_cast<bf16*>(role("xraw", (size_t)Bp * x.H * x.W * L.cout * 2 * PW)) : nullptr;
                last_xraw = xr;
                x = add_res_block(L, x, nullptr, film_off, xr, /*x_has_skip=*/true);
            } else {                    // Downsample: conv3x3 stride 2 on the raw stream (unet.py:68-74)
                const int C = x.C;
                bf16* w = new_weight(C, 9 * C, pick_block_n(C));
                conv_weight_param(L.name + ".conv.weight", w, C, C, 3, 9 * C, 0, C);
                float* b = f32_param(L.name + ".conv.bias", {C});
                bf16* raw = (train && fuse_cast) ? last_xraw : static_cast<bf16*>(role(fuse_cast ? "xraw" : "raw", (size_t)Bp * x.H * x.W * C * 2 * PW));
                if (!fuse_cast) add_cast(x, raw, 1);
                Act y = new_act(C, x.H / 2, x.W / 2, L.name);
                ConvArgs c; c.n_a = 1; c.a[0] = nhwc_stride2_src(raw, Bp, x.H, x.W, C * PW); c.c0 = C;
                add_conv_slabs(c.slabs, 0, C, 3, 2, 0, C * PW);
                c.w = w; c.ktot = 9 * C; c.cout = C; c.OH = y.H; c.OW = y.W; c.bias = b; c.out = y;
                add_conv(c);
                if (train) bwd_downsample(L.name, x, y, raw);
                x = y;
            }
            feats.push_back(x);
        }
        for (size_t mi = 0; mi < mid.size(); ++mi) x = add_res_block(mid[mi], x, nullptr, film_off, nullptr, /*x_has_skip=*/mi == 0);
        for (size_t li = 0; li < ups.size(); ++li) {
            auto& L = ups[li];
            const bool next_is_up = li + 1 < ups.size() && ups[li + 1].kind == 3;
            if (L.kind == 1) {
                Act skip = feats.back(); feats.pop_back();
                bf16* xr = (fuse_cast && fold_up && next_is_up) ? static_cast<bf16*>(role("xraw", (size_t)Bp * x.H * x.W * L.cout * 2 * PW)) : nullptr;
                last_xraw = xr;
                x = add_res_block(L, x, &skip, film_off, xr);
            } else if (!fold_up) {      // Upsample: nearest 2x then conv3x3 (unet.py:58-65), materialised
                const int C = x.C;
                bf16* w = new_weight(C, 9 * C, pick_block_n(C));
                conv_weight_param(L.name + ".conv.weight", w, C, C, 3, 9 * C, 0, C);
                float* b = f32_param(L.name + ".conv.bias", {C});
                bf16* upb = static_cast<bf16*>(role("raw", (size_t)Bp * x.H * 2 * x.W * 2 * C * 2));
                add_cast(x, upb, 2);
                Act y = new_act(C, x.H * 2, x.W * 2, L.name);
                ConvArgs c; c.n_a = 1; c.a[0] = nhwc_src(upb, Bp, y.H, y.W, C);
                add_conv_slabs(c.slabs, 0, C, 3, 1, 0);
                c.w = w; c.ktot = 9 * C; c.cout = C; c.OH = y.H; c.OW = y.W; c.bias = b; c.out = y;
                add_conv(c);
                x = y;
            } else {
                // Upsample folded: output pixel (2i+py, 2j+px) only sees a 2x2 neighbourhood of the low-res input, with the
                // 3x3 taps that alias onto the same low-res pixel summed into one weight (exact in real arithmetic, 2.25x fewer
                // MACs, no 4x-sized intermediate).  One GEMM per phase, each writing its quarter of the NHWC output.
                const int C = x.C, Hl = x.H, Wl = x.W;
                const int rows_pad = ((C + 127) / 128) * 128;
                const bool merge = getenv("SR3_NO_MERGE_UP") == nullptr;        // all four phases in one launch / op (gemm-batch z = phase)
                bf16* wf[4];
                if (dry) { for (int ph = 0; ph < 4; ++ph) wf[ph] = nullptr; }
                else {
                    bf16* wall = static_cast<bf16*>(mem.alloc((size_t)4 * rows_pad * 4 * C * PW * sizeof(bf16)));   // [phase][rows_pad][PW * 4C]
                    for (int ph = 0; ph < 4; ++ph) wf[ph] = wall + (size_t)ph * rows_pad * 4 * C * PW;
                }
                if (!dry) {
                    bf16* w0 = wf[0]; bf16* w1 = wf[1]; bf16* w2 = wf[2]; bf16* w3 = wf[3];
                    const int ldw = 4 * C * PW, low = precise ? 4 * C : 0;
                    add_param(L.name + ".conv.weight", {C, C, 3, 3}, [=](const float* src, cudaStream_t st) {
                        const long long total = 4LL * C * C * 4;
                        fold_upsample_weight_kernel<<<(int)std::min<long long>((total + 255) / 256, 4096), 256, 0, st>>>(src, w0, w1, w2, w3, C, C, ldw, low);
                        CK(cudaGetLastError());
                    });
                    { PackDesc d{}; d.type = 5; d.dst = w0; d.Cout = C; d.Cin = C; d.ld = ldw; d.n = (long long)rows_pad * 4 * C * PW; add_pack(L.name + ".conv.weight", d); }
                }
                float* b = f32_param(L.name + ".conv.bias", {C});
                bf16* raw = (train && fuse_cast) ? last_xraw : static_cast<bf16*>(role(fuse_cast ? "xraw" : "raw", (size_t)Bp * Hl * Wl * C * 2 * PW));
                if (!fuse_cast) add_cast(x, raw, 1);
                bf16* upb = nullptr;
                if (train) {      // the weight gradient contracts dY with the nearest-2x upsampled input: keep a bf16 copy of it
                    upb = static_cast<bf16*>(role("up_x", (size_t)Bp * Hl * 2 * Wl * 2 * C * 2));
                    add_cast(x, upb, 2);
                }
                Act y = new_act(C, Hl * 2, Wl * 2, L.name);
                for (int ph = 0; ph < (merge ? 1 : 4); ++ph) {
                    const int py = ph >> 1, px = ph & 1;
                    ConvArgs c; c.n_a = 1; c.a[0] = nhwc_src(raw, Bp, Hl, Wl, C * PW); c.c0 = C;
                    for (int a = 0; a < 2; ++a)
                        for (int bb = 0; bb < 2; ++bb)
                            for (int ch = 0; ch < C; ch += 64) {
                                KSlab k; k.a_sel = 0; k.a_chan = ch; k.dh = py - 1 + a; k.dw = px - 1 + bb; k.p = 0; k.b_col = (a * 2 + bb) * C + ch;
                                c.slabs.push_back(k);
                            }
                    c.w = wf[ph]; c.ktot = 4 * C; c.cout = C; c.OH = Hl; c.OW = Wl; c.bias = b; c.out = y;
                    c.custom_os = true;
                    c.os.sZ = 0; c.os.sB = 4LL * Hl * Wl * C; c.os.sH = 4LL * Wl * C; c.os.sW = 2LL * C; c.os.off = (long long)py * 2 * Wl * C + (long long)px * C;
                    if (merge) { c.nz = 4; c.b_zrows = rows_pad; c.z_phase = 1; c.z_off_hi = 2LL * Wl * C; c.z_off_lo = C; }
                    add_conv(c);
                }
                if (train) bwd_upsample(L.name, x, y, upb);
                x = y;
            }
        }
        REQUIRE(film_off == F, "film bookkeeping");
        {   // final Block (GN -> SiLU -> conv 64 -> out_channel) with the eps / posterior epilogue
            const int C = x.C, co = cfg.out_channel;
            REQUIRE(co <= 4 && co == cfg.channels, "out_channel must equal diffusion channels (<=4)");
            float* g = f32_param("final_conv.block.0.weight", {C});
            float* be = f32_param("final_conv.block.0.bias", {C});
            bf16* w = new_weight(co, 9 * C, 16);
            conv_weight_param("final_conv.block.3.weight", w, co, C, 3, 9 * C, 0, C);
            float* b = f32_param("final_conv.block.3.bias", {co});
            bf16* a = static_cast<bf16*>(role("a1", (size_t)Bp * H * W * C * 2 * PW));
            add_prep(x, nullptr, g, be, cfg.norm_groups, true, a, nullptr);
            if (train) bwd_final(x, a, g, be, last_mr);
            if (!dry) {
                GemmDesc d; d.n_a = 1; d.a[0] = nhwc_src(a, Bp, H, W, C * PW);
                add_conv_slabs(d.slabs, 0, C, 3, 1, 0);
                set_precise(d, C, 0, 9 * C);
                conv_geometry(d, W, H, Bp, 16);
                d.block_n = 16; d.b_ptr = w; d.b_K = 9 * C * PW; d.b_rows = 128; d.n_tiles = 1; d.b_is_param = true;
                d.mode = 1; d.OW = W; d.OH = H; d.OB = B; d.n_valid = co; d.bias = b; d.ctl = ctl_dev;
                d.post.tab = post_tab; d.post.T = T_cap; d.post.H = H; d.post.W = W; d.post.C = co;
                d.post.x_state = x_state; d.post.eps_out = eps_buf; d.post.mean_out = mean_buf; d.post.noise_buf = noise_buf;
                d.post.in_buf = in_buf; d.post.in_C = in_C * PW; d.post.in_coff = cond_c; d.post.in_lo_off = precise ? in_C : 0;
                push_gemm(d);
                // the statistics arena is cleared for the NEXT step once nobody reads it any more (every GroupNorm apply has passed
                // the grid barrier in front of the final conv)
                ZeroParams zp{}; zp.ptr = reinterpret_cast<float4*>(stats_arena); zp.n4 = (long long)(stats_cap * sizeof(double) / 16);
                mega_record(MOP_ZERO, zp);
            }
        }
    }

    void init(const sr3_unet_config& c, int batch, int device) {
        cfg = c; B = batch; dev = device;
        CK(cudaSetDevice(dev));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, dev));
        REQUIRE(prop.major == 10, "sr3_b200 needs an sm_100 class GPU (found sm_%d%d); there is no fallback path", prop.major, prop.minor);
        REQUIRE(B >= 1, "batch must be >= 1");
        REQUIRE(cfg.inner_channel % 64 == 0, "inner_channel must be a multiple of 64 (got %d)", cfg.inner_channel);
        REQUIRE(cfg.in_channel <= 64, "in_channel must be <= 64");
        REQUIRE(cfg.n_mults >= 1 && cfg.n_mults <= SR3_MAX_LEVELS, "bad n_mults");
        for (int i = 0; i < cfg.n_mults; ++i) REQUIRE((cfg.inner_channel * cfg.channel_mults[i]) % (2 * cfg.norm_groups) == 0 || (cfg.inner_channel * cfg.channel_mults[i]) % cfg.norm_groups == 0, "norm_groups must divide the channel counts");
        inner = cfg.inner_channel; H = W = cfg.image_size;
        REQUIRE(cfg.precision == 0 || cfg.precision == 1, "precision must be 0 (bf16) or 1 (precise)");
        precise = cfg.precision == 1; PW = precise ? 2 : 1;
        cond_c = cfg.conditional ? cfg.in_channel - cfg.channels : 0;
        Bp = (B + 1) & ~1;                         // 8x8 levels tile two images per CTA
        use_graph = getenv("SR3_NO_GRAPH") == nullptr;
        T_cap = 4096;
        REQUIRE(!(train && precise), "the training plan supports the bf16 precision only");
        // pass 1: sizes
        dry = true; stats_used = 0; zero_used = 0;
        build_plan();
        bwd_blocks.clear(); bwd_kinds.clear(); bwd_block_params.clear(); film_slices.clear(); drop_host.clear(); drop_names.clear();
        if (train) {
            zero_cap = zero_used + 4;
            zero_arena = static_cast<float*>(mem.alloc(zero_cap * sizeof(float)));
            loss_dev = static_cast<double*>(mem.alloc(sizeof(double)));
        }
        // allocate
        stats_cap = (stats_used + 3) & ~size_t(3);
        stats_arena = static_cast<double*>(mem.alloc(stats_cap * sizeof(double)));
        for (auto& kv : role_max) role_ptr[kv.first] = mem.alloc(kv.second);
        ctl_dev = static_cast<StepCtl*>(mem.alloc(sizeof(StepCtl)));
        in_buf = static_cast<bf16*>(mem.alloc((size_t)Bp * H * W * in_C * 2 * PW));
        const size_t img = (size_t)Bp * cfg.channels * H * W * 4;
        x_state = static_cast<float*>(mem.alloc(img)); eps_buf = static_cast<float*>(mem.alloc(img));
        mean_buf = static_cast<float*>(mem.alloc(img)); noise_buf = static_cast<float*>(mem.alloc(img));
        io_a = static_cast<float*>(mem.alloc(img)); io_b = static_cast<float*>(mem.alloc(img));
        nl_buf = static_cast<float*>(mem.alloc(Bp * 4));
        nl_table = static_cast<float*>(mem.alloc((T_cap + 1) * 4));
        post_tab = static_cast<float*>(mem.alloc((size_t)5 * T_cap * 4));
        // pass 2: real plan
        dry = false; stats_used = 0; zero_used = 0;
        g_gemm_registry = &gemms;
        g_mega_registry = &mega;
        try { build_plan(); } catch (...) { g_gemm_registry = nullptr; g_mega_registry = nullptr; throw; }
        g_gemm_registry = nullptr;
        g_mega_registry = nullptr;
        if (getenv("SR3_NO_PREFETCH") == nullptr && !train) {
            // every tile kernel pulls the weights of the next one into L2 (the last one those of the next step's first)
            for (size_t i = 0; i < gemms.size(); ++i) {
                const GemmHandle& nx = gemms[(i + 1) % gemms.size()];
                // attention "weights" (B operands that are activations produced by an earlier kernel of this step) are skipped
                bool is_param = false;
                is_param = nx.w_is_param;
                if (is_param) { gemms[i].p->pf_ptr = nx.w_ptr; gemms[i].p->pf_bytes = nx.w_bytes & ~15ll; }
            }
        }
        CK(cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking));
        build_mega();
        CK(cudaDeviceSynchronize());
    }

    // ---- persistent step kernel: serialise the recorded ops (parameter blocks 128-byte aligned) and upload them
    void build_mega() {
        // Off by default: measured on the B200 (profiles/r02_step_kernel.md) the grid barrier + per-op fill / drain costs as much as a
        // launch inside a CUDA graph, and the 320-thread GroupNorm apply runs at half the bandwidth of the stand-alone kernel.
        // SR3_MEGA=1 selects it (bit-identical results).
        use_mega = !train && getenv("SR3_MEGA") != nullptr && atoi(getenv("SR3_MEGA")) != 0 && getenv("SR3_NO_FUSE_CAST") == nullptr && getenv("SR3_NO_FOLD_UP") == nullptr;
        if (!use_mega) return;
        std::vector<MegaOp> host_ops;
        std::vector<uint8_t> blob;
        for (size_t i = 0; i < mega.size(); ++i) {
            MegaRec& r = mega[i];
            MegaOp o{}; o.type = r.type; o.variant = r.variant;
            const uint8_t* src = r.raw.data(); size_t n = r.raw.size();
            if (r.type == MOP_GEMM) {
                const int bn = r.variant & 0xffff, mh = r.variant >> 16;
                if (!((bn == 16 || bn == 32 || bn == 64 || bn == 128) && (mh == 1 || mh == 2) && !(bn == 32 && mh == 2))) { use_mega = false; return; }
                src = reinterpret_cast<const uint8_t*>(r.gp.get()); n = sizeof(GemmParams);
            }
            REQUIRE(n % 4 == 0 && n <= (size_t)(GEMM_HDR_BYTES - HDR_PARAMS), "step kernel: parameter block of %zu bytes does not fit the header", n);
            // ops 0 (embedding + FiLM + nothing upstream) and 1 (first conv: reads the input buffer of the previous launch) and the final
            // clear need no barrier; every other op consumes what all CTAs of its predecessor produced
            o.sync_before = (i >= 2 && r.type != MOP_ZERO) ? 1 : 0;
            o.param_bytes = (int)n;
            blob.resize((blob.size() + 127) & ~size_t(127));
            o.param_off = (long long)blob.size();
            blob.insert(blob.end(), src, src + n);
            host_ops.push_back(o);
            mega_types.push_back(r.type);
        }
        REQUIRE(host_ops.size() >= 3 && mega[0].type == MOP_EMBED_FILM && mega[1].type == MOP_GEMM, "step kernel: unexpected plan head");
        mega_ops_dev = static_cast<MegaOp*>(mem.alloc(host_ops.size() * sizeof(MegaOp), false));
        mega_blob = static_cast<uint8_t*>(mem.alloc(blob.size(), false));
        mega_bar = static_cast<unsigned long long*>(mem.alloc(128));
        mega_prof = static_cast<unsigned long long*>(mem.alloc((host_ops.size() + 1) * 4 * sizeof(unsigned long long)));
        REQUIRE((reinterpret_cast<uintptr_t>(mega_blob) & 127) == 0, "blob not 128B aligned");
        CK(cudaMemcpy(mega_ops_dev, host_ops.data(), host_ops.size() * sizeof(MegaOp), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(mega_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
        static std::vector<int> seen;
        if (first_use_on_device(seen)) CK(cudaFuncSetAttribute(step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_kernel, GEMM_THREADS, SMEM_LIMIT));
        REQUIRE(per_sm >= 1, "step kernel does not fit an SM");
    }
    void launch_mega(cudaStream_t st) {
        MegaParams mp{};
        mp.ops = mega_ops_dev; mp.n_ops = (int)mega_types.size(); mp.blob = mega_blob; mp.bar = mega_bar; mp.prof = mega_prof; mp.ctl = ctl_dev;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(num_sms()); cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = SMEM_LIMIT; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;            // all CTAs co-resident: the grid barriers (and split-K meets) cannot deadlock
        attr[0].val.cooperative = 1;
        cfg.attrs = attr; cfg.numAttrs = getenv("SR3_NO_COOP") ? 0 : 1;
        CK(cudaLaunchKernelEx(&cfg, step_kernel, mp));
    }

    void run_step(cudaStream_t st) {
        if (use_mega) { launch_mega(st); return; }
        if (!use_graph) { for (auto& op : ops) op(st); return; }
        if (!graph) {
            cudaGraph_t g;
            const bool fork = getenv("SR3_NO_FORK") == nullptr && side_begin > 0 && side_end > side_begin && side_join >= side_end;
            if (fork && !side_stream) {
                CK(cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking));
                CK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
            }
            CK(cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal));
            for (int i = 0; i < (int)ops.size(); ++i) {
                if (fork && i == side_begin) {              // branch: embed + film depend on the step prologue only
                    CK(cudaEventRecord(ev_fork, cap_stream));
                    CK(cudaStreamWaitEvent(side_stream, ev_fork, 0));
                }
                if (fork && i == side_join) {
                    CK(cudaEventRecord(ev_join, side_stream));
                    CK(cudaStreamWaitEvent(cap_stream, ev_join, 0));
                }
                ops[i](fork && i >= side_begin && i < side_end ? side_stream : cap_stream);
            }
            CK(cudaStreamEndCapture(cap_stream, &g));
            CK(cudaGraphInstantiate(&graph, g, 0));
            CK(cudaGraphDestroy(g));
        }
        CK(cudaGraphLaunch(graph, st));
    }
    void push_ctl(cudaStream_t st) { CK(cudaMemcpyAsync(ctl_dev, &ctl, sizeof(StepCtl), cudaMemcpyHostToDevice, st)); }
    void load_nchw(const float* src, int C, int coff, float* copy, cudaStream_t st) {
        const long long total = 1LL * B * C * H * W;
        const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 8);
        load_nchw_kernel<<<blocks, 256, 0, st>>>(src, B, C, H, W, in_buf, in_C * PW, coff, copy, precise ? in_C : 0);
        CK(cudaGetLastError());
    }
    size_t img_bytes() const { return (size_t)B * cfg.channels * H * W * 4; }
    void check_params() {
        for (auto& p : params) REQUIRE(p.loaded, "parameter %s was never loaded", p.name.c_str());
    }
    void load_inputs(const float* x, const float* cond, cudaStream_t st) {
        if (cfg.conditional) { REQUIRE(cond != nullptr, "condition_x is required by a conditional model"); load_nchw(cond, cond_c, 0, nullptr, st); }
        else REQUIRE(cond == nullptr, "condition_x given to an unconditional model");
        load_nchw(x, cfg.channels, cond_c, x_state, st);
    }
};

// ------------------------------------------------------------------------------------------------ C ABI
#define API_BEGIN try {
#define API_END                      \
    }                                \
    catch (const std::exception& e) { \
        g_err = e.what();            \
        return 1;                    \
    }                                \
    return 0;

extern "C" {

const char* sr3_last_error(void) { return g_err.c_str(); }
int sr3_abi_version(void) { return 3; }

int sr3_engine_create(const sr3_unet_config* cfg, int batch, int device, sr3_engine** out) {
    API_BEGIN
    REQUIRE(cfg && out, "null argument");
    std::unique_ptr<sr3_engine> e(new sr3_engine());
    e->init(*cfg, batch, device);
    *out = e.release();
    API_END
}
int sr3_engine_create_train(const sr3_unet_config* cfg, int batch, int device, float dropout, sr3_engine** out) {
    API_BEGIN
    REQUIRE(cfg && out, "null argument");
    REQUIRE(dropout >= 0.f && dropout < 1.f, "dropout %f out of range", dropout);
    std::unique_ptr<sr3_engine> e(new sr3_engine());
    e->train = true; e->drop_p = dropout;
    e->init(*cfg, batch, device);
    *out = e.release();
    API_END
}
void sr3_engine_destroy(sr3_engine* e) { delete e; }

int sr3_train_forward(sr3_engine* e, const float* hr, const float* sr, const float* gamma, const float* noise, int loss_type, uint64_t dropout_seed,
                      double* loss_host, void* stream) {
    API_BEGIN
    REQUIRE(e && hr && gamma && noise, "null argument");
    REQUIRE(loss_type == 1 || loss_type == 2, "loss_type must be 1 (l1) or 2 (l2)");
    CK(cudaSetDevice(e->dev));
    e->check_params();
    e->train_forward(hr, sr, gamma, noise, loss_type, dropout_seed, loss_host, static_cast<cudaStream_t>(stream));
    API_END
}
int sr3_train_backward(sr3_engine* e, float grad_scale, float* const* grads, int n_grads, void* stream) {
    API_BEGIN
    REQUIRE(e && grads, "null argument");
    REQUIRE(n_grads == (int)e->params.size(), "expected %d gradient pointers, got %d", (int)e->params.size(), n_grads);
    CK(cudaSetDevice(e->dev));
    e->train_backward(grad_scale, grads, static_cast<cudaStream_t>(stream));
    API_END
}
/* the same backward, layer by layer (last layer first), so that a caller can overlap the gradient all-reduce of finished layers */
int sr3_train_num_backward_blocks(const sr3_engine* e) { return e ? (int)e->bwd_blocks.size() : 0; }
int sr3_train_backward_begin(sr3_engine* e, float grad_scale, float* const* grads, int n_grads) {
    API_BEGIN
    REQUIRE(e && grads && n_grads == (int)e->params.size(), "bad argument");
    e->train_backward_begin(grad_scale, grads);
    API_END
}
int sr3_train_backward_block(sr3_engine* e, int block, void* stream) {
    API_BEGIN
    REQUIRE(e, "null engine");
    CK(cudaSetDevice(e->dev));
    e->train_backward_block(block, static_cast<cudaStream_t>(stream));
    API_END
}
/* Reduce the weight-gradient partial tiles of the layers run since the last flush (one launch): call it before handing a finished bucket of
 * gradients to the all-reduce.  sr3_train_backward / _finish flush what is left themselves. */
int sr3_train_backward_flush(sr3_engine* e, void* stream) {
    API_BEGIN
    REQUIRE(e && !e->grad_dst.empty(), "sr3_train_backward_begin has not been called");
    CK(cudaSetDevice(e->dev));
    e->reduce_flush(static_cast<cudaStream_t>(stream));
    API_END
}
int sr3_train_backward_finish(sr3_engine* e, void* stream) {
    API_BEGIN
    REQUIRE(e && !e->grad_dst.empty(), "sr3_train_backward_begin has not been called");
    CK(cudaSetDevice(e->dev));
    e->bwd_film_and_embed(static_cast<cudaStream_t>(stream));
    API_END
}
int sr3_train_block_params(const sr3_engine* e, int block, int* indices, int cap, int* n) {
    API_BEGIN
    REQUIRE(e && block >= 0 && block < (int)e->bwd_block_params.size() && n, "bad argument");
    const std::vector<int>& v = e->bwd_block_params[block];
    *n = (int)v.size();
    for (int i = 0; i < (int)v.size() && i < cap; ++i) indices[i] = v[i];
    API_END
}
/* Profiling: the backward of the forward that just ran, with CUDA events around every op; ms_by_kind[8] receives the summed device time per
 * op kind (0 data-gradient tile kernel, 1 GroupNorm / elementwise, 4 other, 6 weight gradient + slice reduction, 7 attention GEMMs). */
int sr3_train_backward_profile(sr3_engine* e, float grad_scale, float* const* grads, int n_grads, float* ms_by_kind, void* stream) {
    API_BEGIN
    REQUIRE(e && grads && ms_by_kind && n_grads == (int)e->params.size(), "bad argument");
    CK(cudaSetDevice(e->dev));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    e->train_backward_begin(grad_scale, grads);
    e->reduce_prepare(st); e->reduce_begun = true;
    std::vector<cudaEvent_t> evs;
    std::vector<int> kinds;
    auto mark = [&]() { cudaEvent_t ev; CK(cudaEventCreate(&ev)); CK(cudaEventRecord(ev, st)); evs.push_back(ev); };
    mark();
    int blocks_since_flush = 0;
    for (size_t i = e->bwd_blocks.size(); i-- > 0;) {
        for (size_t j = 0; j < e->bwd_blocks[i].size(); ++j) { e->bwd_blocks[i][j](st); kinds.push_back(e->bwd_kinds[i][j]); mark(); }
        if (++blocks_since_flush == 6) { e->reduce_flush(st); kinds.push_back(5); mark(); blocks_since_flush = 0; }     // about one flush per gradient bucket
    }
    e->reduce_flush(st); kinds.push_back(5); mark();
    e->bwd_film_and_embed(st); kinds.push_back(4); mark();
    CK(cudaStreamSynchronize(st));
    for (int k = 0; k < 8; ++k) ms_by_kind[k] = 0.f;
    for (size_t i = 0; i < kinds.size(); ++i) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, evs[i], evs[i + 1]));
        ms_by_kind[kinds[i] & 7] += ms;
    }
    for (auto ev : evs) cudaEventDestroy(ev);
    API_END
}
int sr3_train_set_dropout_mask(sr3_engine* e, const char* block_name, const unsigned char* mask_nchw) {
    API_BEGIN
    REQUIRE(e && block_name, "null argument");
    REQUIRE(e->train, "engine was not created with sr3_engine_create_train");
    bool found = false;
    for (size_t i = 0; i < e->drop_names.size(); ++i)
        if (e->drop_names[i] == block_name) { e->drop_host[i].mask = mask_nchw; found = true; }
    REQUIRE(found, "no dropout layer named %s", block_name);
    API_END
}
int sr3_train_num_dropout_layers(const sr3_engine* e) { return e ? (int)e->drop_names.size() : 0; }
int sr3_train_dropout_layer_name(const sr3_engine* e, int index, char* name, int name_cap) {
    API_BEGIN
    REQUIRE(e && index >= 0 && index < (int)e->drop_names.size() && name && name_cap > 0, "bad argument");
    strncpy(name, e->drop_names[index].c_str(), name_cap - 1); name[name_cap - 1] = 0;
    API_END
}
int sr3_adam_step(const void* table_dev, int n_tensors, float lr, float beta1, float beta2, float eps, int step, float grad_scale, void* stream) {
    API_BEGIN
    REQUIRE(table_dev && n_tensors > 0 && step >= 1, "bad argument");
    const float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
    const int ny = n_tensors < 65535 ? n_tensors : 65535;
    adam_kernel<<<dim3(64, ny), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const AdamTensor*>(table_dev), n_tensors, lr, beta1, beta2, eps, bc1, sqrtf(bc2), grad_scale);
    CK(cudaGetLastError());
    API_END
}

int sr3_engine_num_params(const sr3_engine* e) { return e ? (int)e->params.size() : 0; }
int sr3_engine_param_info(const sr3_engine* e, int index, char* name, int name_cap, int64_t shape[4], int* ndim) {
    API_BEGIN
    REQUIRE(e && index >= 0 && index < (int)e->params.size(), "bad param index");
    const ParamEntry& p = e->params[index];
    if (name && name_cap > 0) { strncpy(name, p.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
    for (int i = 0; i < 4; ++i) shape[i] = i < (int)p.shape.size() ? p.shape[i] : 1;
    if (ndim) *ndim = (int)p.shape.size();
    API_END
}
int sr3_engine_load_param(sr3_engine* e, const char* name, const float* src, int64_t numel, void* stream) {
    API_BEGIN
    REQUIRE(e && name && src, "null argument");
    auto it = e->pindex.find(name);
    REQUIRE(it != e->pindex.end(), "unexpected key in state_dict: %s", name);
    ParamEntry& p = e->params[it->second];
    REQUIRE(p.numel == numel, "size mismatch for %s: expected %lld elements, got %lld", name, (long long)p.numel, (long long)numel);
    CK(cudaSetDevice(e->dev));
    p.load(src, static_cast<cudaStream_t>(stream));
    for (auto& h : p.hooks) h(src, static_cast<cudaStream_t>(stream));
    p.loaded = true;
    API_END
}
/* load_state_dict in one call: srcs[i] = DEVICE fp32 pointer of parameter i (sr3_engine_param_info order); re-packs everything and runs the
 * finalisation.  Asynchronous on `stream` (the training loop calls it after every optimizer step). */
int sr3_engine_load_all_params(sr3_engine* e, const float* const* srcs, int n, void* stream) {
    API_BEGIN
    REQUIRE(e && srcs && n == (int)e->params.size(), "expected %d parameter pointers", e ? (int)e->params.size() : 0);
    CK(cudaSetDevice(e->dev));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (int i = 0; i < n; ++i) REQUIRE(srcs[i] != nullptr, "null pointer for %s", e->params[i].name.c_str());
    if (!e->pack_descs.empty() && !e->precise && getenv("SR3_NO_PACK_TABLE") == nullptr) {
        // one launch over the descriptor table (rebuilt only when a parameter moved)
        const size_t nd = e->pack_descs.size();
        if (e->pack_last_ptrs.size() != (size_t)n || memcmp(e->pack_last_ptrs.data(), srcs, n * sizeof(float*)) != 0 || !e->pack_dev) {
            std::vector<PackDesc> tab = e->pack_descs;
            for (size_t i = 0; i < nd; ++i) { tab[i].src = srcs[e->pack_src[i]]; tab[i].src2 = e->pack_src2[i] >= 0 ? srcs[e->pack_src2[i]] : nullptr; }
            if (!e->pack_dev) {
                e->pack_dev = static_cast<PackDesc*>(e->mem.alloc(nd * sizeof(PackDesc)));
                e->pack_ends_dev = static_cast<int*>(e->mem.alloc(nd * sizeof(int)));
                // blocks in proportion to the work of an entry: ~8 (o, c) pairs (x k*k taps) or 32 plain elements per thread
                std::vector<int> ends(nd);
                int acc = 0;
                for (size_t i = 0; i < nd; ++i) {
                    const PackDesc& d = tab[i];
                    long long items;
                    switch (d.type) {
                        case 0: case 6: items = (d.n + 3) / 4; break;
                        case 1: case 2: items = 1LL * d.Cout * d.Cin; break;
                        case 3: items = 2LL * d.Cin * d.Cout; break;
                        case 4: items = 4LL * d.Cin * d.Cout; break;
                        default: items = 4LL * d.Cin * d.Cout; break;
                    }
                    long long nb = (items + 256 * 8 - 1) / (256 * 8);
                    if (nb < 1) nb = 1;
                    if (nb > 4096) nb = 4096;
                    acc += (int)nb; ends[i] = acc;
                }
                e->pack_blocks = acc;
                CK(cudaMemcpy(e->pack_ends_dev, ends.data(), nd * sizeof(int), cudaMemcpyHostToDevice));
            }
            CK(cudaMemcpyAsync(e->pack_dev, tab.data(), nd * sizeof(PackDesc), cudaMemcpyHostToDevice, st));
            CK(cudaStreamSynchronize(st));                 // `tab` is pageable host memory
            e->pack_last_ptrs.assign(srcs, srcs + n);
        }
        pack_all_kernel<<<dim3((unsigned)e->pack_blocks), 256, 0, st>>>(e->pack_dev, e->pack_ends_dev, (int)nd);
        CK(cudaGetLastError());
        for (auto& p : e->params) p.loaded = true;
        return 0;
    }
    for (int i = 0; i < n; ++i) {
        ParamEntry& p = e->params[i];
        p.load(srcs[i], st);
        for (auto& h : p.hooks) h(srcs[i], st);
        p.loaded = true;
    }
    for (auto& op : e->finalize_ops) op(st);
    API_END
}
int sr3_engine_finalize_params(sr3_engine* e, void* stream) {
    API_BEGIN
    REQUIRE(e, "null engine");
    e->check_params();
    for (auto& op : e->finalize_ops) op(static_cast<cudaStream_t>(stream));
    API_END
}

int sr3_engine_set_schedule(sr3_engine* e, int T, const float* a, const float* b, const float* c1, const float* c2, const float* lv,
                            const double* sqrt_ac_prev, void* stream) {
    API_BEGIN
    REQUIRE(e && a && b && c1 && c2 && lv && sqrt_ac_prev, "null argument");
    REQUIRE(T >= 1 && T <= e->T_cap, "n_timestep %d out of range (max %d)", T, e->T_cap);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(e->dev));
    e->T = T;
    std::vector<float> tab((size_t)5 * e->T_cap, 0.f), nl(T + 1);
    const float* srcs[5] = {a, b, c1, c2, lv};
    for (int k = 0; k < 5; ++k) memcpy(tab.data() + (size_t)k * e->T_cap, srcs[k], T * sizeof(float));
    for (int i = 0; i <= T; ++i) nl[i] = static_cast<float>(sqrt_ac_prev[i]);     // FloatTensor([f64]) rounding, diffusion.py:153
    e->logvar_host.assign(lv, lv + T);
    CK(cudaMemcpyAsync(e->post_tab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(e->nl_table, nl.data(), nl.size() * 4, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    API_END
}

int sr3_unet_forward(sr3_engine* e, const float* x, const float* noise_level, float* eps, void* stream) {
    API_BEGIN
    REQUIRE(e && x && noise_level && eps, "null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(e->dev));
    e->load_nchw(x, e->cfg.in_channel, 0, nullptr, st);
    CK(cudaMemcpyAsync(e->nl_buf, noise_level, e->B * 4, cudaMemcpyDeviceToDevice, st));
    StepCtl& c = e->ctl; memset(&c, 0, sizeof(c));
    c.nl_from_table = 0; c.out_mode = 0; c.t_next = 0;
    e->push_ctl(st);
    e->run_step(st);
    CK(cudaMemcpyAsync(eps, e->eps_buf, e->img_bytes(), cudaMemcpyDeviceToDevice, st));
    API_END
}

int sr3_p_mean_variance(sr3_engine* e, const float* x, const float* cond, int t, int clip, float* mean, float* log_variance, void* stream) {
    API_BEGIN
    REQUIRE(e && x && mean, "null argument");
    REQUIRE(e->T > 0, "set_new_noise_schedule has not been called");
    REQUIRE(t >= 0 && t < e->T, "t=%d out of range [0,%d)", t, e->T);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(e->dev));
    e->load_inputs(x, cond, st);
    StepCtl& c = e->ctl; memset(&c, 0, sizeof(c));
    c.nl_from_table = 1; c.out_mode = 1; c.write_mean = 1; c.update_state = 0; c.use_noise_buf = 1; c.clip = clip; c.t_next = t;
    CK(cudaMemsetAsync(e->noise_buf, 0, e->img_bytes(), st));
    e->push_ctl(st);
    e->run_step(st);
    CK(cudaMemcpyAsync(mean, e->mean_buf, e->img_bytes(), cudaMemcpyDeviceToDevice, st));
    if (log_variance) *log_variance = e->logvar_host[t];
    API_END
}

int sr3_p_sample(sr3_engine* e, const float* x, const float* cond, int t, const float* noise, uint64_t seed, uint64_t first_index,
                 float* x_prev, void* stream) {
    API_BEGIN
    REQUIRE(e && x && x_prev, "null argument");
    REQUIRE(e->T > 0, "set_new_noise_schedule has not been called");
    REQUIRE(t >= 0 && t < e->T, "t=%d out of range [0,%d)", t, e->T);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(e->dev));
    e->load_inputs(x, cond, st);
    StepCtl& c = e->ctl; memset(&c, 0, sizeof(c));
    c.nl_from_table = 1; c.out_mode = 1; c.update_state = 1; c.clip = 1; c.t_next = t; c.seed = seed; c.sample_offset = first_index;
    if (noise) { c.use_noise_buf = 1; CK(cudaMemcpyAsync(e->noise_buf, noise, e->img_bytes(), cudaMemcpyDeviceToDevice, st)); }
    e->push_ctl(st);
    e->run_step(st);
    CK(cudaMemcpyAsync(x_prev, e->x_state, e->img_bytes(), cudaMemcpyDeviceToDevice, st));
    API_END
}

int sr3_p_losses(sr3_engine* e, const float* hr, const float* sr, const float* gamma, const float* noise, int loss_type, double* loss_host,
                 void* stream) {
    API_BEGIN
    REQUIRE(e && hr && gamma && noise && loss_host, "null argument");
    REQUIRE(loss_type == 1 || loss_type == 2, "loss_type must be 1 (l1) or 2 (l2)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(e->dev));
    if (e->cfg.conditional) { REQUIRE(sr != nullptr, "x_in['SR'] is required by a conditional model"); e->load_nchw(sr, e->cond_c, 0, nullptr, st); }
    const long long total = 1LL * e->B * e->cfg.channels * e->H * e->W;
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 8);
    q_sample_load_kernel<<<blocks, 256, 0, st>>>(hr, noise, gamma, e->B, e->cfg.channels, e->H, e->W, e->in_buf, e->in_C * e->PW, e->cond_c,
                                                 e->precise ? e->in_C : 0);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(e->nl_buf, gamma, e->B * 4, cudaMemcpyDeviceToDevice, st));
    StepCtl& c = e->ctl; memset(&c, 0, sizeof(c));
    c.nl_from_table = 0; c.out_mode = 0; c.t_next = 0;
    e->push_ctl(st);
    e->run_step(st);
    if (!e->loss_dev) e->loss_dev = static_cast<double*>(e->mem.alloc(sizeof(double)));
    CK(cudaMemsetAsync(e->loss_dev, 0, sizeof(double), st));
    loss_sum_kernel<<<blocks, 256, 0, st>>>(noise, e->eps_buf, total, loss_type == 2 ? 1 : 0, e->loss_dev);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(loss_host, e->loss_dev, sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    API_END
}

int sr3_p_sample_loop_begin(sr3_engine* e, const float* cond, const float* x_T, uint64_t seed, uint64_t first_index, void* stream) {
    API_BEGIN
    REQUIRE(e && x_T, "null argument");
    REQUIRE(e->T > 0, "set_new_noise_schedule has not been called");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(e->dev));
    e->load_inputs(x_T, cond, st);
    e->seed = seed; e->first_index = first_index;
    API_END
}
int sr3_p_sample_steps(sr3_engine* e, int t_start, int steps, void* stream) {
    API_BEGIN
    REQUIRE(e, "null engine");
    REQUIRE(t_start < e->T && steps >= 0 && t_start - steps + 1 >= 0, "bad step range t_start=%d steps=%d T=%d", t_start, steps, e->T);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(e->dev));
    StepCtl& c = e->ctl; memset(&c, 0, sizeof(c));
    c.nl_from_table = 1; c.out_mode = 1; c.update_state = 1; c.clip = 1; c.t_next = t_start; c.seed = e->seed; c.sample_offset = e->first_index;
    e->push_ctl(st);
    for (int i = 0; i < steps; ++i) e->run_step(st);
    API_END
}
int sr3_read_state(sr3_engine* e, float* x_out, void* stream) {
    API_BEGIN
    REQUIRE(e && x_out, "null argument");
    CK(cudaMemcpyAsync(x_out, e->x_state, e->img_bytes(), cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    API_END
}

int sr3_p_sample_loop(sr3_engine* e, const float* cond, const float* x_T, const float* noises, uint64_t seed, uint64_t first_index,
                      float* final, float* snapshots, int snapshot_cap, int* n_snapshots, void* stream) {
    API_BEGIN
    REQUIRE(e && x_T, "null argument");
    REQUIRE(e->T > 0, "set_new_noise_schedule has not been called");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(e->dev));
    e->load_inputs(x_T, cond, st);
    const int T = e->T;
    StepCtl& c = e->ctl; memset(&c, 0, sizeof(c));
    c.nl_from_table = 1; c.out_mode = 1; c.update_state = 1; c.clip = 1; c.t_next = T - 1; c.seed = seed; c.sample_offset = first_index;
    c.use_noise_buf = noises ? 1 : 0;
    e->push_ctl(st);
    const int inter = 1 | (T / 10);                     // diffusion.py:179
    const size_t ib = e->img_bytes();
    int ns = 0;
    for (int i = T - 1; i >= 0; --i) {
        if (noises) CK(cudaMemcpyAsync(e->noise_buf, noises + (size_t)i * (ib / 4), ib, cudaMemcpyDeviceToDevice, st));
        e->run_step(st);
        if (i % inter == 0) {
            if (snapshots) {
                REQUIRE(ns < snapshot_cap, "snapshot buffer too small (%d)", snapshot_cap);
                CK(cudaMemcpyAsync(snapshots + (size_t)ns * (ib / 4), e->x_state, ib, cudaMemcpyDeviceToDevice, st));
            }
            ++ns;
        }
    }
    if (final) CK(cudaMemcpyAsync(final, e->x_state, ib, cudaMemcpyDeviceToDevice, st));
    if (n_snapshots) *n_snapshots = ns;
    API_END
}

int sr3_super_resolution_host(sr3_engine* e, const float* cond_host, const float* x_T_host, uint64_t seed, uint64_t first_index,
                              float* final_host, void* stream) {
    API_BEGIN
    REQUIRE(e && x_T_host && final_host, "null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(e->dev));
    const size_t ib = e->img_bytes();
    if (cond_host) CK(cudaMemcpyAsync(e->io_a, cond_host, ib, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(e->io_b, x_T_host, ib, cudaMemcpyHostToDevice, st));
    int ns = 0;
    int rc = sr3_p_sample_loop(e, cond_host ? e->io_a : nullptr, e->io_b, nullptr, seed, first_index, nullptr, nullptr, 0, &ns, stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(final_host, e->x_state, ib, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    API_END
}

int sr3_engine_profile_step(sr3_engine* e, int t, int reps, int cap, int* kinds, float* ms, double* flops, double* bytes, int* n_ops, void* stream) {
    API_BEGIN
    REQUIRE(e && kinds && ms && flops && bytes && n_ops, "null argument");
    REQUIRE(e->T > 0 && t >= 0 && t < e->T, "bad t");
    const int n = (int)e->ops.size();
    REQUIRE(cap >= n, "profile buffers too small (%d ops)", n);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(e->dev));
    std::vector<cudaEvent_t> ev(n + 1);
    for (auto& x : ev) CK(cudaEventCreate(&x));
    std::vector<double> acc(n, 0.0);
    for (int r = 0; r < reps + 1; ++r) {          // first repetition is a warm-up
        StepCtl& c = e->ctl; memset(&c, 0, sizeof(c));
        c.nl_from_table = 1; c.out_mode = 1; c.update_state = 0; c.clip = 1; c.t_next = t; c.seed = 1;
        e->push_ctl(st);
        for (int i = 0; i < n; ++i) { CK(cudaEventRecord(ev[i], st)); e->ops[i](st); }
        CK(cudaEventRecord(ev[n], st));
        CK(cudaStreamSynchronize(st));
        if (r == 0) continue;
        for (int i = 0; i < n; ++i) { float m = 0; CK(cudaEventElapsedTime(&m, ev[i], ev[i + 1])); acc[i] += m; }
    }
    for (int i = 0; i < n; ++i) {
        kinds[i] = e->op_info[i].kind; ms[i] = (float)(acc[i] / (reps > 0 ? reps : 1));
        flops[i] = e->op_info[i].flops; bytes[i] = e->op_info[i].bytes;
    }
    *n_ops = n;
    for (auto& x : ev) cudaEventDestroy(x);
    // the per-layer path clears the statistics arena at the START of a step, the step kernel at the END of one: leave it clean
    CK(cudaMemsetAsync(e->stats_arena, 0, e->stats_cap * sizeof(double), st));
    CK(cudaStreamSynchronize(st));
    API_END
}

int sr3_engine_uses_step_kernel(const sr3_engine* e) { return (e && e->use_mega) ? 1 : 0; }

// Per-op device time of the most recent step-kernel launch (globaltimer stamps taken by CTA 0 after each grid barrier).
int sr3_engine_step_kernel_profile(sr3_engine* e, int cap, int* types, double* us, double* phases, int* n_ops, void* stream) {
    API_BEGIN
    REQUIRE(e && types && us && n_ops, "null argument");
    REQUIRE(e->use_mega, "this engine runs the per-layer path (set SR3_MEGA=1 for the step kernel), there is nothing to profile here");
    const int n = (int)e->mega_types.size();
    REQUIRE(cap >= n, "profile buffers too small (%d ops)", n);
    CK(cudaSetDevice(e->dev));
    CK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    std::vector<unsigned long long> ts((size_t)(n + 1) * 4);
    CK(cudaMemcpy(ts.data(), e->mega_prof, ts.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) {
        types[i] = e->mega_types[i];
        us[i] = (double)(ts[4 * (i + 1)] - ts[4 * i]) * 1e-3;
        if (phases) {      // [set-up (arrive + parameter / stage-table copy), barrier wait, body, end-of-op fence]
            phases[4 * i + 0] = (double)(ts[4 * i + 1] - ts[4 * i]) * 1e-3;
            phases[4 * i + 1] = (double)(ts[4 * i + 2] - ts[4 * i + 1]) * 1e-3;
            phases[4 * i + 2] = (double)(ts[4 * i + 3] - ts[4 * i + 2]) * 1e-3;
            phases[4 * i + 3] = (double)(ts[4 * (i + 1)] - ts[4 * i + 3]) * 1e-3;
        }
    }
    *n_ops = n;
    API_END
}

int sr3_engine_num_launches_per_step(const sr3_engine* e) { return e ? (e->use_mega ? 1 : (int)e->ops.size()) : 0; }
int sr3_engine_num_ops_per_step(const sr3_engine* e) { return e ? (int)e->ops.size() : 0; }
int64_t sr3_engine_workspace_bytes(const sr3_engine* e) { return e ? e->mem.bytes : 0; }

int sr3_engine_read_activation(sr3_engine* e, const char* name, float* dst, int64_t cap, int64_t* numel, int shape_bhwc[4], void* stream) {
    API_BEGIN
    REQUIRE(e && name, "null argument");
    auto it = e->taps.find(name);
    REQUIRE(it != e->taps.end(), "no activation tap named %s", name);
    const Act& a = it->second;
    const int64_t n = (int64_t)e->B * a.H * a.W * a.C;
    if (numel) *numel = n;
    if (shape_bhwc) { shape_bhwc[0] = e->B; shape_bhwc[1] = a.H; shape_bhwc[2] = a.W; shape_bhwc[3] = a.C; }
    if (dst) {
        REQUIRE(cap >= n, "destination too small");
        CK(cudaMemcpyAsync(dst, a.p, n * 4, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    }
    API_END
}

int sr3_test_gemm(const void* a, const void* b, float* dptr, int M, int N, int K, int block_n, void* stream) {
    API_BEGIN
    REQUIRE(M % 128 == 0 && K % 64 == 0 && N % block_n == 0, "bad test gemm shape");
    DevAllocs mem;
    GemmDesc d; d.n_a = 1; d.a[0] = matrix_src(a, 1, M, K, K, 0);
    for (int c = 0; c < K; c += 64) d.slabs.push_back({0, c, 0, 0, 0, c});
    d.block_n = block_n; d.b_ptr = b; d.b_K = K; d.b_rows = N;
    d.w_box = 128; d.h_box = 1; d.b_box = 1; d.tiles_w = M / 128; d.tiles_h = 1; d.tiles_b = 1; d.n_tiles = N / block_n; d.nz = 1;
    d.OW = M; d.OH = 1; d.OB = 1; d.n_valid = N;
    d.out_f32 = dptr; d.os = OutSpec{0, 0, 0, N, 0};
    Op op = make_gemm_op(d, mem);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    op(st);
    CK(cudaStreamSynchronize(st));
    API_END
}

int sr3_test_attention(const void* qk, const void* vT, void* out, int nz, int Lt, int HW, int C, void* stream) {
    API_BEGIN
    Op op = make_attn_op(static_cast<const bf16*>(qk), static_cast<const bf16*>(vT), static_cast<bf16*>(out), nz, Lt, HW, C);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    op(st);
    CK(cudaStreamSynchronize(st));
    API_END
}

int sr3_bench_conv(int B, int H, int W, int Cin, int Cout, int ksize, int stride, int with_resid, int with_stats, int reps, float* ms_out) {
    API_BEGIN
    REQUIRE(ms_out && reps > 0, "bad arguments");
    DevAllocs mem;
    const int ktot = ksize * ksize * Cin;
    const int OH = H / stride, OW = W / stride;
    bf16* x = static_cast<bf16*>(mem.alloc((size_t)B * H * W * Cin * 2));
    bf16* wp = static_cast<bf16*>(mem.alloc((size_t)Cout * ktot * 2));
    float* y = static_cast<float*>(mem.alloc((size_t)B * OH * OW * Cout * 4));
    float* r = with_resid ? static_cast<float*>(mem.alloc((size_t)B * OH * OW * Cout * 4)) : nullptr;
    double* st = with_stats ? static_cast<double*>(mem.alloc((size_t)B * Cout * 2 * 8)) : nullptr;
    float* bias = static_cast<float*>(mem.alloc((size_t)Cout * 4));
    GemmDesc d; d.n_a = 1;
    d.a[0] = stride == 1 ? nhwc_src(x, B, H, W, Cin) : nhwc_stride2_src(x, B, H, W, Cin);
    add_conv_slabs(d.slabs, 0, Cin, ksize, stride, 0);
    conv_geometry(d, OW, OH, B, Cout, with_resid != 0);
    d.b_ptr = wp; d.b_K = ktot; d.b_rows = Cout;
    REQUIRE(B % d.b_box == 0, "batch must be a multiple of %d at this resolution", d.b_box);
    REQUIRE(Cout >= d.block_n, "Cout smaller than the tile");
    d.n_tiles = Cout / d.block_n;
    d.OW = OW; d.OH = OH; d.OB = B; d.n_valid = Cout; d.bias = bias;
    d.out_f32 = y; d.os = nhwc_out(OH, OW, Cout);
    d.resid = r; d.rs = nhwc_out(OH, OW, Cout);
    d.stats = st; d.stats_C = Cout;
    Op op = make_gemm_op(d, mem);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    // timed as ONE captured graph of `reps` launches: no CPU launch overhead in the number (as inside the step graph)
    cudaStream_t cs;
    CK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    for (int i = 0; i < 3; ++i) op(cs);
    CK(cudaStreamSynchronize(cs));
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < reps; ++i) op(cs);
    CK(cudaStreamEndCapture(cs, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    CK(cudaGraphLaunch(ge, cs));
    CK(cudaStreamSynchronize(cs));
    CK(cudaEventRecord(e0, cs));
    CK(cudaGraphLaunch(ge, cs));
    CK(cudaEventRecord(e1, cs));
    CK(cudaEventSynchronize(e1));
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaGraphExecDestroy(ge); cudaGraphDestroy(g); cudaStreamDestroy(cs);
    *ms_out = ms / reps;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    API_END
}

int sr3_test_conv(const void* x, const float* w_oihw, const float* bias, float* y, double* stats, int B, int H, int W, int Cin, int Cout,
                  int ksize, int stride, void* stream) {
    API_BEGIN
    REQUIRE((ksize == 1 || ksize == 3) && (stride == 1 || stride == 2) && Cin % 64 == 0 && Cout % 64 == 0, "bad test conv shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DevAllocs mem;
    const int ktot = ksize * ksize * Cin;
    bf16* wp = static_cast<bf16*>(mem.alloc((size_t)Cout * ktot * 2));
    pack_conv_weight_kernel<<<1024, 256, 0, st>>>(w_oihw, wp, Cout, Cin, ksize, ksize, ktot, 0, Cin, 0);
    CK(cudaGetLastError());
    const int OH = H / stride, OW = W / stride;
    GemmDesc d; d.n_a = 1;
    d.a[0] = stride == 1 ? nhwc_src(x, B, H, W, Cin) : nhwc_stride2_src(x, B, H, W, Cin);
    add_conv_slabs(d.slabs, 0, Cin, ksize, stride, 0);
    conv_geometry(d, OW, OH, B, Cout);
    d.b_ptr = wp; d.b_K = ktot; d.b_rows = Cout;
    REQUIRE(B % d.b_box == 0, "batch must be a multiple of %d at this resolution", d.b_box);
    REQUIRE(Cout >= d.block_n, "Cout smaller than the tile");
    d.n_tiles = Cout / d.block_n;
    d.OW = OW; d.OH = OH; d.OB = B; d.n_valid = Cout; d.bias = bias;
    d.out_f32 = y; d.os = nhwc_out(OH, OW, Cout);
    d.stats = stats; d.stats_C = Cout;
    Op op = make_gemm_op(d, mem);
    op(st);
    CK(cudaStreamSynchronize(st));
    API_END
}

// core/metrics.py:8-34 tensor2img on the device (see tensor2img_kernel).  src fp32 [n][C][H][W] DEVICE, dst uint8 DEVICE [GH][GW][C] with
// n == 1: GH = H, GW = W;  n > 1: make_grid geometry, nrow images per row: GH = rows * (H + 2) + 2, GW = ncol * (W + 2) + 2.
int sr3_tensor2img(const float* src, unsigned char* dst, int n, int C, int H, int W, int nrow, float min_v, float max_v, void* stream) {
    API_BEGIN
    REQUIRE(src && dst && n >= 1 && C >= 1 && H >= 1 && W >= 1 && max_v > min_v, "bad tensor2img arguments");
    int ncol = 1, GH = H, GW = W;
    if (n > 1) {
        REQUIRE(nrow >= 1, "nrow must be >= 1");
        ncol = nrow < n ? nrow : n;                       // make_grid: xmaps = min(nrow, nmaps), ymaps = ceil(nmaps / xmaps)
        const int rows = (n + ncol - 1) / ncol;
        GH = rows * (H + 2) + 2; GW = ncol * (W + 2) + 2;
    }
    const long long total = 1LL * GH * GW * C;
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 8);
    tensor2img_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, n, C, H, W, ncol, GH, GW, min_v, max_v);
    CK(cudaGetLastError());
    API_END
}

}  // extern "C"

namespace {
// Pillow's Resample.c: precompute_coeffs (bicubic filter, a = -0.5, support 2) + normalize_coeffs_8bpc (PRECISION_BITS = 22), whole input range
double pil_bicubic_filter(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}
void pil_bicubic_tables(int in_size, int out_size, std::vector<int>& bounds, std::vector<int>& coef, int& ksize) {
    const double scale = (double)in_size / out_size;
    double filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 2.0 * filterscale;
    ksize = (int)ceil(support) * 2 + 1;
    bounds.assign((size_t)out_size * 2, 0);
    coef.assign((size_t)out_size * ksize, 0);
    const double ss = 1.0 / filterscale;
    std::vector<double> w((size_t)ksize);
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = 0.0 + (xx + 0.5) * scale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) { w[x] = pil_bicubic_filter((x + xmin - center + 0.5) * ss); ww += w[x]; }
        for (int x = 0; x < xmax; ++x) {
            const double v = (ww != 0.0) ? w[x] / ww : w[x];
            coef[(size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << 22)) : (int)(0.5 + v * (1 << 22));
        }
        bounds[2 * xx] = xmin; bounds[2 * xx + 1] = xmax;
    }
}
}  // namespace

extern "C" {

// Host-only: the integer coefficient tables of one resampling pass (needs no GPU; checked against Pillow's by the CPU test-suite).
int sr3_pil_bicubic_tables(int in_size, int out_size, int* bounds, int* coef, int coef_cap, int* ksize) {
    API_BEGIN
    REQUIRE(in_size >= 1 && out_size >= 1 && bounds && coef && ksize, "bad arguments");
    std::vector<int> b, c;
    int ks = 0;
    pil_bicubic_tables(in_size, out_size, b, c, ks);
    REQUIRE((int)c.size() <= coef_cap, "coefficient buffer too small (%d needed)", (int)c.size());
    memcpy(bounds, b.data(), b.size() * sizeof(int));
    memcpy(coef, c.data(), c.size() * sizeof(int));
    *ksize = ks;
    API_END
}

// data/prepare_data.py:17-40 (`trans_fn.resize(img, size, Image.BICUBIC)`: Pillow's two-pass fixed-point bicubic resampler) and
// data/util.py:74-83 (ToTensor, optional horizontal flip, range mapping) on the device.  src uint8 DEVICE [B][h][w][C] (HWC, as PIL hands
// it over); dst_u8 (optional) uint8 DEVICE [B][H][W][C]; dst_f32 (optional) fp32 DEVICE [B][C][H][W] = (resized / 255) * (max - min) + min,
// mirrored along W when flip != 0.  Integer-exact against Pillow.
int sr3_resize_bicubic_u8(const unsigned char* src, unsigned char* dst_u8, float* dst_f32, int B, int h, int w, int C, int H, int W, int flip,
                          float min_v, float max_v, void* stream) {
    API_BEGIN
    REQUIRE(src && (dst_u8 || dst_f32) && B >= 1 && h >= 1 && w >= 1 && C >= 1 && C <= 4 && H >= 1 && W >= 1, "bad resize arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DevAllocs mem;
    std::vector<int> bh, ch, bv, cv;
    int ksh = 0, ksv = 0;
    pil_bicubic_tables(w, W, bh, ch, ksh);
    pil_bicubic_tables(h, H, bv, cv, ksv);
    auto up = [&](const std::vector<int>& v) {
        int* d = static_cast<int*>(mem.alloc(v.size() * sizeof(int), false));
        CK(cudaMemcpyAsync(d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        return d;
    };
    int *dbh = up(bh), *dch = up(ch), *dbv = up(bv), *dcv = up(cv);
    // horizontal pass first (Pillow's order): [B][h][w][C] -> tmp [B][h][W][C]; then vertical: -> [B][H][W][C] (+ float planes)
    unsigned char* tmp = static_cast<unsigned char*>(mem.alloc((size_t)B * h * W * C, false));
    {
        const long long total = 1LL * B * h * W * C;
        resample_u8_kernel<<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, st>>>(
            src, tmp, nullptr, B, /*lines*/ h, /*out*/ W, C, /*in axis*/ C, /*in line*/ 1LL * w * C, /*in img*/ 1LL * h * w * C,
            /*out axis*/ C, /*out line*/ 1LL * W * C, /*out img*/ 1LL * h * W * C, dbh, dch, ksh, 0, 0.f, 1.f);
        CK(cudaGetLastError());
    }
    {
        const long long total = 1LL * B * W * H * C;
        resample_u8_kernel<<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, st>>>(
            tmp, dst_u8, dst_f32, B, /*lines = x*/ W, /*out = y*/ H, C, /*in axis (y)*/ 1LL * W * C, /*in line (x)*/ C, /*in img*/ 1LL * h * W * C,
            /*out axis*/ 1LL * W * C, /*out line*/ C, /*out img*/ 1LL * H * W * C, dbv, dcv, ksv, flip, min_v, max_v);
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(st));             // the tables / intermediate die with this call
    API_END
}

// calculate_psnr (core/metrics.py:42-50): returns the exact integer sum of squared differences of two uint8 DEVICE images through *ssd_host.
int sr3_ssd_u8(const unsigned char* a, const unsigned char* b, int64_t n, unsigned long long* ssd_host, void* stream) {
    API_BEGIN
    REQUIRE(a && b && ssd_host && n >= 1, "bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DevAllocs mem;
    unsigned long long* d = static_cast<unsigned long long*>(mem.alloc(sizeof(unsigned long long)));
    ssd_u8_kernel<<<(int)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, st>>>(a, b, (long long)n, d);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ssd_host, d, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    API_END
}

// Test hook: conv (tile kernel, statistics in its epilogue) followed by the GroupNorm(+SiLU) apply pass, i.e. one Block of the
// reference (GN -> Swish -> conv, unet.py:80-91) seen from the GN's side: y = conv(x) + bias, a = [silu](GN(y; gamma, beta)).
int sr3_test_conv_groupnorm(const void* x, const float* w_oihw, const float* bias, const float* gamma, const float* beta, int groups, int silu,
                            float* y, void* a_bf16, int B, int H, int W, int Cin, int Cout, int ksize, void* stream) {
    API_BEGIN
    REQUIRE((ksize == 1 || ksize == 3) && Cin % 64 == 0 && Cout % 64 == 0 && Cout % groups == 0, "bad test shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DevAllocs mem;
    double* stats = static_cast<double*>(mem.alloc((size_t)B * Cout * 2 * sizeof(double)));
    int rc = sr3_test_conv(x, w_oihw, bias, y, stats, B, H, W, Cin, Cout, ksize, 1, stream);
    if (rc) return rc;
    PrepParams p{};
    p.src0 = y; p.st0 = stats; p.C0 = Cout; p.C1 = 0;
    p.gamma = gamma; p.beta = beta; p.groups = groups; p.HW = H * W; p.silu = silu; p.eps = 1e-5f;
    p.out_a = static_cast<bf16*>(a_bf16); p.out_raw = nullptr;
    const int vpp = Cout / 4;
    REQUIRE(vpp <= 512, "too many channels");
    const int kpix = vpp >= 256 ? 1 : 256 / vpp;
    p.pix_per_block = kpix * 4; p.B = B; p.items_per_image = (p.HW + p.pix_per_block - 1) / p.pix_per_block;
    const dim3 grid((p.HW + p.pix_per_block - 1) / p.pix_per_block, B);
    launch_k(prep_kernel<false>, grid, dim3(vpp * kpix), (size_t)((2 * Cout + 2 * groups) * sizeof(float)), st, p);
    CK(cudaStreamSynchronize(st));
    API_END
}

}  // extern "C"
