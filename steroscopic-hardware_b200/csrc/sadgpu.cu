// sadgpu.cu — C-ABI shim of libsadgpu.so (include/sadgpu.h): context, per-stream pinned and
// device buffers, the launch planner and the kernel dispatch.  No CPU fallback exists: every
// compute entry point ends in a kernel launch or an error code.
#include "../../include/sadgpu.h"
#include "sad_kernels.cuh"
#include "sad_fast.cuh"
#include "sad_ws.cuh"
#include "sad_wide.cuh"
#include "sad_vh.cuh"
#include "sad_ring.cuh"
#include "gray_kernels.cuh"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

using namespace sadgpu;

namespace {

constexpr int kSmemBudget = 232448;      // 227 KB opt-in dynamic shared memory per CTA on sm_100
constexpr int kGT = 5;                   // disparity groups per phase-B thread
constexpr int kMaxThreadsPerColumn = 4;  // K: phase-B threads per pixel column
constexpr int kMaxDevices = 64;

// Developer builds (-DSADGPU_DEV_H0=7 -DSADGPU_DEV_H1=15) instantiate the kernels of two half-windows only: seconds, not minutes.
constexpr bool dev_on(int half)
{
#ifdef SADGPU_DEV_H0
    return half == SADGPU_DEV_H0 || half == SADGPU_DEV_H1;
#else
    (void)half; return true;
#endif
}

struct Plan {
    SadArgs a;
    dim3 grid;
    int nthreads;
    size_t smem;
    int half;
    int launches;
};

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

int validate(int w, int h, int B, int D, int y0, int y1)
{
    if (w <= 0 || h <= 0) return SADGPU_EINVAL;
    if (B < 1 || B > SADGPU_MAX_BLOCK_SIZE || D < 1 || D > SADGPU_MAX_DISPARITY) return SADGPU_EINVAL;
    if (y0 < 0 || y1 > h || y0 > y1) return SADGPU_ERANGE;
    return SADGPU_OK;
}

// Chooses tile geometry.  All quantities are documented in DESIGN.md §3.
int make_plan(int w, int h, int B, int D, int y0, int y1, const sadgpu_tuning* t, int sm_count, Plan* p)
{
    int rc = validate(w, h, B, D, y0, y1);
    if (rc) return rc;
    SadArgs& a = p->a;
    memset(&a, 0, sizeof(a));
    const int half = B / 2, WIN = 2 * half + 1;
    p->half = half;
    a.W = w; a.H = h; a.y0 = y0; a.y1 = y1; a.D = D;
    a.NG = (D + 4) / 4;                                   // groups of 4 disparities covering 0..D
    int maxg = kGT * kMaxThreadsPerColumn;
    if (t && t->groups_per_chunk > 0) maxg = std::min(maxg, t->groups_per_chunk);
    a.NSTEP = 64;
    a.TW = a.NSTEP - 2 * half;
    a.TWp = round_up(a.TW, 32);
    int RBmax = 0;
    for (;; --maxg) {
        if (maxg < 1) return SADGPU_EINVAL;
        a.NC = ceil_div(a.NG, maxg);
        a.NGc = ceil_div(a.NG, a.NC);
        a.K = ceil_div(a.NGc, kGT);
        a.NGP = a.NGc | 1;                                // odd stride: conflict-free 64-bit column reads
        a.RW = a.NGc + a.NSTEP / 4;
        const int ringrow = a.TW * a.NGP * 8;
        const int perrow = ringrow + a.NSTEP * 4 + a.RW * 4 + a.K * a.TW * 4;
        const int fixed = WIN * ringrow + round_up(4 * a.NG, 16) + 64;
        RBmax = (kSmemBudget - fixed) / perrow;
        if (RBmax >= 4) break;
    }
    a.RB = std::min(RBmax, 16);
    if (t && t->rows_per_batch > 0) a.RB = std::min(RBmax, t->rows_per_batch);
    a.NR = a.RB + WIN;
    const int ringrow = a.TW * a.NGP * 8;
    a.offL = round_up(a.NR * ringrow, 16);
    a.offR = a.offL + a.RB * a.NSTEP * 4;
    a.offPk = a.offR + a.RB * a.RW * 4;
    a.offLut = a.offPk + a.RB * a.K * a.TW * 4;
    p->smem = (size_t)a.offLut + round_up(4 * a.NG, 16);
    if (p->smem > (size_t)kSmemBudget) return SADGPU_EINVAL;
    p->nthreads = 256;
    if (a.TWp * a.K > p->nthreads) return SADGPU_EINVAL;

    const int rows = y1 - y0;
    const int nstrips = ceil_div(w, a.TW);
    int nbands = 1;
    if (t && t->band_rows > 0) {
        a.BH = std::min(std::max(1, t->band_rows), std::max(rows, 1));
        nbands = ceil_div(std::max(rows, 1), a.BH);
    } else {
        // minimise waves x rows-per-CTA (one CTA per SM: the H ring fills shared memory)
        long best_cost = -1;
        for (int nb = 1; nb <= std::max(1, std::min(rows, 256)); ++nb) {
            const int bh = ceil_div(std::max(rows, 1), nb);
            const long ctas = (long)nstrips * nb * a.NC;
            const long waves = (ctas + sm_count - 1) / sm_count;
            const long cost = waves * (round_up(bh + 2 * half, a.RB) + 2);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; nbands = nb; }
        }
        a.BH = ceil_div(std::max(rows, 1), nbands);
        nbands = ceil_div(std::max(rows, 1), a.BH);
    }
    p->grid = dim3(nstrips, nbands, a.NC);
    p->launches = a.NC == 1 ? 1 : 3;
    return SADGPU_OK;
}

template <int HALF>
cudaError_t launch_generic(const Plan& p, cudaStream_t s, bool* attr_done)
{
    auto k = sad_generic_kernel<HALF, kGT>;
    if (!*attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
        if (e != cudaSuccess) return e;
        *attr_done = true;
    }
    k<<<p.grid, p.nthreads, p.smem, s>>>(p.a);
    return cudaGetLastError();
}

typedef cudaError_t (*launch_fn)(const Plan&, cudaStream_t, bool*);
template <int HALF> constexpr launch_fn generic_entry()
{
    if constexpr (dev_on(HALF)) return launch_generic<HALF>; else return nullptr;
}
const launch_fn kLaunchGeneric[16] = {
    generic_entry<0>(), generic_entry<1>(), generic_entry<2>(), generic_entry<3>(),
    generic_entry<4>(), generic_entry<5>(), generic_entry<6>(), generic_entry<7>(),
    generic_entry<8>(), generic_entry<9>(), generic_entry<10>(), generic_entry<11>(),
    generic_entry<12>(), generic_entry<13>(), generic_entry<14>(), generic_entry<15>()};


// ---------------------------------------------------------------------------------------
// Fast path (sad_fast.cuh): block_size <= 15.  Template instance = (h, groups per chunk).
// ---------------------------------------------------------------------------------------
struct FastPlan {
    FastArgs a;
    dim3 grid;
    int nthreads;
    size_t smem;
    int half, ngc, rb, tw;
    int launches;
};

template <int HALF, int NGC>
cudaError_t launch_fast(const FastPlan& p, cudaStream_t s, bool* attr_done)
{
    using C = FastCfg<HALF, NGC>;
    static_assert(C::SMEM <= kSmemBudget, "fast kernel does not fit shared memory");
    auto k = sad_fast_kernel<HALF, NGC>;
    if (!*attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) return e;
        *attr_done = true;
    }
    k<<<p.grid, C::NT, C::SMEM, s>>>(p.a);
    return cudaGetLastError();
}

template <int HALF>
cudaError_t launch_ws(const FastPlan& p, cudaStream_t s, bool* attr_done)
{
    using C = WsCfg<HALF>;
    static_assert(C::SMEM <= kSmemBudget, "warp-specialised kernel does not fit shared memory");
    if (!*attr_done) {
        cudaError_t e = cudaFuncSetAttribute(sad_ws_kernel<HALF, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(sad_ws_kernel<HALF, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) return e;
        *attr_done = true;
    }
    if (p.a.use_tma) sad_ws_kernel<HALF, true><<<p.grid, C::NT, C::SMEM, s>>>(p.a);
    else             sad_ws_kernel<HALF, false><<<p.grid, C::NT, C::SMEM, s>>>(p.a);
    return cudaGetLastError();
}

template <int HALF>
cudaError_t launch_wide(const FastPlan& p, cudaStream_t s, bool* attr_done)
{
    using C = WideCfg<HALF>;
    static_assert(C::SMEM <= kSmemBudget, "wide kernel does not fit shared memory");
    auto k = sad_wide_kernel<HALF>;
    if (!*attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) return e;
        *attr_done = true;
    }
    k<<<p.grid, C::NT, C::SMEM, s>>>(p.a);
    return cudaGetLastError();
}

template <int HALF>
cudaError_t launch_vh(const FastPlan& p, cudaStream_t s, bool* attr_done)
{
    using C = VhCfg<HALF>;
    static_assert(C::SMEM <= kSmemBudget, "vertical-first kernel does not fit shared memory");
    auto k = sad_vh_kernel<HALF>;
    if (!*attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) return e;
        *attr_done = true;
    }
    k<<<p.grid, C::NT, C::SMEM, s>>>(p.a);
    return cudaGetLastError();
}

template <int HALF>
cudaError_t launch_ring(const FastPlan& p, cudaStream_t s, bool* attr_done)
{
    using C = RingCfg<HALF>;
    static_assert(C::SMEM <= kSmemBudget, "ring kernel does not fit shared memory");
    auto k = sad_ring_kernel<HALF>;
    if (!*attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) return e;
        *attr_done = true;
    }
    k<<<p.grid, C::NT, C::SMEM, s>>>(p.a);
    return cudaGetLastError();
}

typedef cudaError_t (*fast_fn)(const FastPlan&, cudaStream_t, bool*);
struct FastEntry { fast_fn fn; int nt, smem, rb; };
template <int HALF, int NGC> constexpr FastEntry fast_entry()
{
    if constexpr (dev_on(HALF)) return FastEntry{launch_fast<HALF, NGC>, FastCfg<HALF, NGC>::NT, FastCfg<HALF, NGC>::SMEM, FastCfg<HALF, NGC>::RB};
    else return FastEntry{nullptr, FastCfg<HALF, NGC>::NT, FastCfg<HALF, NGC>::SMEM, FastCfg<HALF, NGC>::RB};
}
template <int HALF> constexpr FastEntry ws_entry()
{
    if constexpr (dev_on(HALF)) return FastEntry{launch_ws<HALF>, WsCfg<HALF>::NT, WsCfg<HALF>::SMEM, WsCfg<HALF>::RB};
    else return FastEntry{nullptr, WsCfg<HALF>::NT, WsCfg<HALF>::SMEM, WsCfg<HALF>::RB};
}
template <int HALF> constexpr FastEntry wide_entry()
{
    if constexpr (dev_on(HALF)) return FastEntry{launch_wide<HALF>, WideCfg<HALF>::NT, WideCfg<HALF>::SMEM, WideCfg<HALF>::RB};
    else return FastEntry{nullptr, WideCfg<HALF>::NT, WideCfg<HALF>::SMEM, WideCfg<HALF>::RB};
}
template <int HALF> constexpr FastEntry ring_entry()
{
    if constexpr (dev_on(HALF)) return FastEntry{launch_ring<HALF>, RingCfg<HALF>::NT, RingCfg<HALF>::SMEM, 4};
    else return FastEntry{nullptr, RingCfg<HALF>::NT, RingCfg<HALF>::SMEM, 4};
}
template <int HALF> constexpr FastEntry vh_entry()
{
    if constexpr (dev_on(HALF)) return FastEntry{launch_vh<HALF>, VhCfg<HALF>::NT, VhCfg<HALF>::SMEM, VhCfg<HALF>::CROWS};
    else return FastEntry{nullptr, VhCfg<HALF>::NT, VhCfg<HALF>::SMEM, VhCfg<HALF>::CROWS};
}
// index [half][slot], slot 0/1/2 = 9/17/33 groups per chunk; h >= 5 has no 33-group instance (shared memory)
const int kFastNgc[3] = {9, 17, 33};       // h >= 5 uses 18 in slot 1 (3 groups per phase-B thread, 384 threads: no register spills)
const FastEntry kFast[8][3] = {
    {fast_entry<0, 9>(), fast_entry<0, 17>(), fast_entry<0, 33>()},
    {fast_entry<1, 9>(), fast_entry<1, 17>(), fast_entry<1, 33>()},
    {fast_entry<2, 9>(), fast_entry<2, 17>(), fast_entry<2, 33>()},
    {fast_entry<3, 9>(), fast_entry<3, 17>(), fast_entry<3, 33>()},
    {fast_entry<4, 9>(), fast_entry<4, 17>(), fast_entry<4, 33>()},
    {fast_entry<5, 9>(), fast_entry<5, 18>(), FastEntry{nullptr, 0, 0, 0}},
    {fast_entry<6, 9>(), fast_entry<6, 18>(), FastEntry{nullptr, 0, 0, 0}},
    {fast_entry<7, 9>(), fast_entry<7, 18>(), FastEntry{nullptr, 0, 0, 0}}};

// slot 3 = warp-specialised kernel (sad_ws.cuh): h <= 4, 33-group chunks, 32-column strips
const FastEntry kWs[5] = {ws_entry<0>(), ws_entry<1>(), ws_entry<2>(), ws_entry<3>(), ws_entry<4>()};

// slot 4 = large-window kernel (sad_wide.cuh): h = 8..15, chunks of 8 groups
#define WIDE_ENTRY(H) wide_entry<H>()
const FastEntry kWide[8] = {WIDE_ENTRY(8), WIDE_ENTRY(9), WIDE_ENTRY(10), WIDE_ENTRY(11), WIDE_ENTRY(12), WIDE_ENTRY(13), WIDE_ENTRY(14), WIDE_ENTRY(15)};

// slot 5 = vertical-first warp-specialised kernel (sad_vh.cuh): h = 5..15, 33-group chunks
#define VH_ENTRY(H) vh_entry<H>()
const FastEntry kVh[11] = {VH_ENTRY(5), VH_ENTRY(6), VH_ENTRY(7), VH_ENTRY(8), VH_ENTRY(9), VH_ENTRY(10), VH_ENTRY(11), VH_ENTRY(12),
                           VH_ENTRY(13), VH_ENTRY(14), VH_ENTRY(15)};
const int kVhTw[11] = {VhCfg<5>::TW, VhCfg<6>::TW, VhCfg<7>::TW, VhCfg<8>::TW, VhCfg<9>::TW, VhCfg<10>::TW, VhCfg<11>::TW, VhCfg<12>::TW,
                       VhCfg<13>::TW, VhCfg<14>::TW, VhCfg<15>::TW};
bool vh_supported(int B) { return B / 2 >= 5 && B / 2 <= 15; }
// slot 6 = shared-memory-ring warp-specialised kernel (sad_ring.cuh): h = 5..15, 32-column strips, chunks of 33 groups (h <= 7) or 17
const FastEntry kRing[11] = {ring_entry<5>(), ring_entry<6>(), ring_entry<7>(), ring_entry<8>(), ring_entry<9>(), ring_entry<10>(),
                             ring_entry<11>(), ring_entry<12>(), ring_entry<13>(), ring_entry<14>(), ring_entry<15>()};
bool ring_supported(int B) { return B / 2 >= 5 && B / 2 <= 15; }
// Planner default, from the measured variant sweep (profiles/r01_variant_sweep.json, inputs streaming from HBM): a ring pass over
// 33 groups costs about 1.7x a pass of the phase-alternating kernel over 18 groups, a ring pass over
// 17 groups 1.95x a pass of the wide kernel over 8 groups; the ring kernel is chosen whenever its passes are cheaper in total.
bool ring_auto(int B, int D)
{
    if (!ring_supported(B)) return false;
    const int ng = (D + 4) / 4;
    if (B / 2 <= 7) return 172 * ((ng + 32) / 33) < 100 * ((ng + 17) / 18);
    return 195 * ((ng + 16) / 17) < 100 * ((ng + 7) / 8);
}

bool vh_auto(int B, int D) { (void)B; (void)D; return false; }      // never the fastest variant (profiles/r01_variant_sweep.json)

bool fast_supported(int B) { return B / 2 <= 7; }
bool wide_supported(int B) { return B / 2 >= 8 && B / 2 <= 15; }
bool ws_supported(int B, int D) { return B / 2 <= 4 && (D + 4) / 4 > 17; }

int make_fast_plan(int w, int h, int B, int D, int y0, int y1, int n_frames, const sadgpu_tuning* t, int sm_count,
                   FastPlan* p, int* slot_out, bool want_ws = false, bool want_vh = false, bool want_ring = false)
{
    int rc = validate(w, h, B, D, y0, y1);
    if (rc) return rc;
    if ((!fast_supported(B) && !wide_supported(B)) || n_frames < 1) return SADGPU_EINVAL;
    const int half = B / 2;
    const bool wide = wide_supported(B);
    FastArgs& a = p->a;
    memset(&a, 0, sizeof(a));
    a.W = w; a.H = h; a.y0 = y0; a.y1 = y1; a.D = D;
    a.NG = (D + 4) / 4;
    int slot = a.NG <= 9 ? 0 : a.NG <= (half >= 5 ? 18 : 17) ? 1 : 2;
    if (!wide && !kFast[half][slot].fn) slot = 1;
    if (t && t->groups_per_chunk > 0) {                     // tests: force smaller chunks
        slot = t->groups_per_chunk <= 9 ? 0 : t->groups_per_chunk <= 17 ? 1 : slot;
    }
    const bool ws = !wide && want_ws && ws_supported(B, D);
    const bool vh = want_vh && vh_supported(B);
    if (ws) slot = 3;
    if (wide) slot = 4;
    const bool ring = want_ring && ring_supported(B);
    if (vh) slot = 5;
    if (ring) slot = 6;
    const FastEntry& fe = ring ? kRing[half - 5] : vh ? kVh[half - 5] : wide ? kWide[half - 8] : ws ? kWs[half] : kFast[half][slot];
    const int tw = ring ? 32 : vh ? kVhTw[half - 5] : ws ? 32 : 64;
    p->tw = tw;
    p->half = half; p->ngc = ring ? (half >= 8 ? 17 : 33) : vh ? 33 : wide ? 8 : ws ? 33 : (slot == 1 && half >= 5) ? 18 : kFastNgc[slot]; p->rb = fe.rb;
    a.NC = ceil_div(a.NG, p->ngc);
    p->nthreads = fe.nt; p->smem = fe.smem;
    const int rows = std::max(1, y1 - y0);
    const int nstrips = ceil_div(w, tw);
    int nbands = 1;
    if (t && t->band_rows > 0) {
        a.BH = std::min(std::max(1, t->band_rows), rows);
    } else {
        long best_cost = -1;
        for (int nb = 1; nb <= std::min(rows, 64); ++nb) {
            const int bh = ceil_div(rows, nb);
            const long ctas = (long)nstrips * nb * a.NC * n_frames;
            const long waves = (ctas + sm_count - 1) / sm_count;
            // rows a CTA spends on pipeline fill and drain, beyond its band and the window halo
            const int fill = ws ? 2 * fe.rb : (ring || vh) ? 12 : fe.rb / 2 + 2;
            const long cost = waves * (round_up(bh + 2 * half, fe.rb) + fill);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; nbands = nb; }
        }
        a.BH = ceil_div(rows, nbands);
    }
    nbands = ceil_div(rows, a.BH);
    p->grid = dim3(nstrips, nbands, a.NC * n_frames);
    p->launches = a.NC == 1 ? 1 : 3;
    *slot_out = slot;
    return SADGPU_OK;
}

// ---- TMA descriptors for the warp-specialised kernel (driver entry point fetched through the runtime) ----
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
encode_tiled_fn get_encode_tiled()
{
    static encode_tiled_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess)
            fn = (encode_tiled_fn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

// 3-D uint8 tensor (x, y, frame) with a (box_w x box_h x 1) box, zero fill outside.  Returns false when TMA cannot be used.
bool make_tmap(CUtensorMap* m, const uint8_t* base, int w, int h, size_t pitch, long long frame_stride, int n_frames, int box_w, int box_h)
{
    encode_tiled_fn enc = get_encode_tiled();
    if (!enc) return false;
    const unsigned long long fs = (n_frames > 1 && frame_stride > 0) ? (unsigned long long)frame_stride : (unsigned long long)pitch * h;
    if ((uintptr_t)base % 16 || pitch % 16 || fs % 16 || box_w % 16 || box_w > 256 || box_h > 256) return false;
    cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n_frames};
    cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)fs};
    cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int HALF> void ws_box(int* lbox, int* rbox, int* rb) { *lbox = WsCfg<HALF>::LBOX; *rbox = WsCfg<HALF>::RWT * 4; *rb = WsCfg<HALF>::RB; }

// Host-side staging copies (pageable caller memory <-> pinned buffers) are memory-bound single-thread memcpys of several
// megabytes per frame; a few helper threads cut them to a fraction.  The pool is created with the context.
class CopyPool {
public:
    explicit CopyPool(int helpers)
    {
        for (int i = 0; i < helpers; ++i) th_.emplace_back([this] { run(); });
    }
    ~CopyPool()
    {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    // rows x width bytes, row pitches dp / sp; contiguous when dp == sp == width.  Small copies stay on the caller.
    void copy(uint8_t* dst, size_t dp, const uint8_t* src, size_t sp, size_t width, size_t rows)
    {
        if (rows == 0 || width == 0) return;
        const bool contig = dp == width && sp == width;
        const size_t total = width * rows;
        // waking a sleeping helper costs tens of microseconds on this class of host: only copies of 4 MB and more are split
        // (measured: 1080p planes got slower when split three ways, 4K planes 20 % faster)
        const int parts = total < (4u << 20) ? 1 : (int)std::min<size_t>(th_.size() + 1, total / (2u << 20));
        if (parts <= 1) { one(dst, dp, src, sp, width, rows, contig); return; }
        Latch latch{parts - 1};
        for (int p = 1; p < parts; ++p) {
            const size_t r0 = rows * p / parts, r1 = rows * (p + 1) / parts;
            const size_t b0 = total * p / parts, b1 = total * (p + 1) / parts;
            push([=, &latch] {
                if (contig) memcpy(dst + b0, src + b0, b1 - b0);
                else one(dst + r0 * dp, dp, src + r0 * sp, sp, width, r1 - r0, false);
                latch.done();
            });
        }
        if (contig) memcpy(dst, src, total / parts);
        else one(dst, dp, src, sp, width, rows / parts, false);
        latch.wait();
    }
private:
    struct Latch {                      // lives on the caller's stack: the worker's last access is the decrement
        std::atomic<int> n;
        explicit Latch(int k) : n(k) {}
        void done() { n.fetch_sub(1, std::memory_order_release); }
        void wait() { while (n.load(std::memory_order_acquire) > 0) std::this_thread::yield(); }   // tens of microseconds
    };
    static void one(uint8_t* dst, size_t dp, const uint8_t* src, size_t sp, size_t width, size_t rows, bool contig)
    {
        if (contig) { memcpy(dst, src, width * rows); return; }
        for (size_t y = 0; y < rows; ++y) memcpy(dst + y * dp, src + y * sp, width);
    }
    void push(std::function<void()> f)
    {
        { std::lock_guard<std::mutex> g(m_); q_.push_back(std::move(f)); }
        cv_.notify_one();
    }
    void run()
    {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                f = std::move(q_.front()); q_.erase(q_.begin());
            }
            f();
        }
    }
    std::vector<std::thread> th_;
    std::vector<std::function<void()>> q_;
    std::mutex m_;
    std::condition_variable cv_;
    bool stop_ = false;
};

struct Slot {
    int dev_index = 0, device = 0;
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;
    uint8_t *hL = nullptr, *hR = nullptr, *hOut = nullptr;     // pinned staging
    uint8_t *dL = nullptr, *dR = nullptr, *dOut = nullptr;     // device, pitched
    uint32_t* gkey = nullptr;
    size_t pitch = 0;                                          // pitch of the frame in flight = round_up(w, 4): contiguous copies when the host stride matches
    std::mutex mu;
    bool busy = false;
    uint64_t seq = 0;
    int w = 0, h = 0, y0 = 0, y1 = 0;                          // frame in flight
    bool out_direct = false;
    uint8_t *hRGBA = nullptr, *dRGBA = nullptr;                // staging of the interleaved colour pair (sadgpu_compute_nrgba), lazily allocated
    size_t rgba_bytes = 0;
    int cap = 1;                                               // frame pairs the buffers hold (sadgpu_reserve_batch)
    int nfr = 1;                                               // frames in flight
};

}  // namespace

struct sadgpu_ctx {
    std::vector<int> devices;
    std::vector<int> sm_count;
    std::vector<Slot*> slots;
    int max_w = 0, max_h = 0;
    std::atomic<int> last_launches{0};
    std::mutex pool_mu;
    std::vector<std::pair<uint8_t*, size_t>> pool;
    bool attr_done[kMaxDevices][16];
    bool fast_attr_done[kMaxDevices][8][5];
    bool vh_attr_done[kMaxDevices][11];
    bool ring_attr_done[kMaxDevices][11];
    CopyPool* copier = nullptr;
    std::vector<size_t> dev_gkey_bytes;
    std::vector<uint32_t*> dev_gkey;       // per device scratch for sadgpu_compute_device
    std::mutex dev_mu;
};

namespace {

bool in_pool(sadgpu_ctx* c, const void* p, size_t bytes)
{
    std::lock_guard<std::mutex> g(c->pool_mu);
    const uint8_t* q = static_cast<const uint8_t*>(p);
    for (auto& r : c->pool)
        if (q >= r.first && q + bytes <= r.first + r.second) return true;
    return false;
}

struct Job {
    const uint8_t* dL; size_t pitchL; long long frameL;
    const uint8_t* dR; size_t pitchR; long long frameR;
    uint8_t* dOut; size_t pitchOut; long long frameOut;
    int n_frames, w, h, B, D, y0, y1;
};

// Grows the per-device key-map scratch (only needed when the disparity range is chunked).
int ensure_gkey(sadgpu_ctx* c, int dev_index, size_t bytes, uint32_t** out)
{
    std::lock_guard<std::mutex> g(c->dev_mu);
    if (c->dev_gkey_bytes[dev_index] < bytes) {
        if (c->dev_gkey[dev_index]) { cudaDeviceSynchronize(); cudaFree(c->dev_gkey[dev_index]); c->dev_gkey[dev_index] = nullptr; }
        cudaError_t e = cudaMalloc((void**)&c->dev_gkey[dev_index], bytes);
        if (e != cudaSuccess) { c->dev_gkey_bytes[dev_index] = 0; return (int)e; }
        c->dev_gkey_bytes[dev_index] = bytes;
    }
    *out = c->dev_gkey[dev_index];
    return SADGPU_OK;
}

// Enqueue one job (one frame, or a batch of frames) on stream s.  The device must be current.
// Chooses the fast path (block_size <= 15) or the generic kernel; never a CPU path.
int run_job(sadgpu_ctx* c, int dev_index, const Job& j, const sadgpu_tuning* t, uint32_t* slot_gkey, cudaStream_t s, int slot_gkey_frames = 1)
{
    const int variant = t ? t->kernel_variant : 0;
    if (variant < 0 || variant > 6) return SADGPU_EINVAL;
    if (variant == 5 && !vh_supported(j.B)) return SADGPU_EINVAL;
    if (variant == 6 && !ring_supported(j.B)) return SADGPU_EINVAL;
    if (variant == 2 && !fast_supported(j.B)) return SADGPU_EINVAL;
    if (variant == 3 && !ws_supported(j.B, j.D)) return SADGPU_EINVAL;
    if (variant == 4 && !wide_supported(j.B)) return SADGPU_EINVAL;
    const bool use_fast = variant >= 2 || (variant == 0 && (fast_supported(j.B) || wide_supported(j.B)));
    const bool want_ws = variant == 3 || variant == 0;
    int launches = 0;
    if (use_fast) {
        FastPlan p; int slot = 0;
        int rc = make_fast_plan(j.w, j.h, j.B, j.D, j.y0, j.y1, j.n_frames, t, c->sm_count[dev_index], &p, &slot, want_ws,
                                variant == 5 || (variant == 0 && vh_auto(j.B, j.D)), variant == 6 || (variant == 0 && ring_auto(j.B, j.D)));
        if (rc) return rc;
        if (j.y1 == j.y0) return SADGPU_OK;
        FastArgs& a = p.a;
        a.L = j.dL; a.R = j.dR; a.out = j.dOut;
        a.pitchL = (int)j.pitchL; a.pitchR = (int)j.pitchR; a.pitchOut = (int)j.pitchOut;
        a.frameL = j.frameL; a.frameR = j.frameR; a.frameOut = j.frameOut;
        a.k65536 = 65536u;
        a.debug_skip = t ? t->reserved[1] : 0;
        a.use_tma = 0;
        if (slot == 3 && !(t && t->reserved[2] == 1)) {                 // reserved[2] == 1: force the non-TMA loader (tests)
            int lbox = 0, rbox = 0, rb = 0;
            switch (p.half) { case 0: ws_box<0>(&lbox, &rbox, &rb); break; case 1: ws_box<1>(&lbox, &rbox, &rb); break;
                              case 2: ws_box<2>(&lbox, &rbox, &rb); break; case 3: ws_box<3>(&lbox, &rbox, &rb); break;
                              default: ws_box<4>(&lbox, &rbox, &rb); }
            if (make_tmap(&a.tmapL, j.dL, j.w, j.h, j.pitchL, j.frameL, j.n_frames, lbox, rb) &&
                make_tmap(&a.tmapR, j.dR, j.w, j.h, j.pitchR, j.frameR, j.n_frames, rbox, rb))
                a.use_tma = 1;
        }
        a.aligned = ((uintptr_t)j.dR % 4 == 0 && j.pitchR % 4 == 0 && j.frameR % 4 == 0) ? 1 : 0;
        if (a.debug_skip & 4) { uint32_t* gk = nullptr; rc = ensure_gkey(c, dev_index, 4096, &gk); if (rc) return rc; a.gkey = gk; }
        if (a.NC > 1) {
            const size_t n = (size_t)j.n_frames * j.w * j.h;
            uint32_t* gk = slot_gkey;
            if (!gk || j.n_frames > slot_gkey_frames) { rc = ensure_gkey(c, dev_index, n * sizeof(uint32_t), &gk); if (rc) return rc; }
            a.gkey = gk;
            sad_fill_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 8192), 256, 0, s>>>(gk, n, 0xFFFFFFFFu);
        }
        const FastEntry& fe = slot == 6 ? kRing[p.half - 5] : slot == 5 ? kVh[p.half - 5] : slot == 4 ? kWide[p.half - 8] : slot == 3 ? kWs[p.half] : kFast[p.half][slot];
        bool* done = slot == 6 ? &c->ring_attr_done[dev_index][p.half - 5] : slot == 5 ? &c->vh_attr_done[dev_index][p.half - 5]
                   : slot == 4 ? &c->fast_attr_done[dev_index][p.half - 8][4] : &c->fast_attr_done[dev_index][p.half][slot];
        if (!fe.fn) return SADGPU_EINVAL;                              // developer build without this instance
        cudaError_t e = fe.fn(p, s, done);
        if (e != cudaSuccess) return (int)e;
        if (a.NC > 1) {
            dim3 g(ceil_div(j.w, 256), j.y1 - j.y0, j.n_frames);
            sad_finalize_kernel<<<g, 256, 0, s>>>(a.gkey, j.dOut, j.w, j.h, j.y0, j.y1, (int)j.pitchOut, j.frameOut, j.D);
            if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
        }
        launches = p.launches;
    } else {
        Plan p;
        int rc = make_plan(j.w, j.h, j.B, j.D, j.y0, j.y1, t, c->sm_count[dev_index], &p);
        if (rc) return rc;
        if (j.y1 == j.y0) return SADGPU_OK;
        uint32_t* gk = slot_gkey;
        if (p.a.NC > 1 && !gk) { rc = ensure_gkey(c, dev_index, (size_t)j.w * j.h * sizeof(uint32_t), &gk); if (rc) return rc; }
        for (int f = 0; f < j.n_frames; ++f) {
            p.a.L = j.dL + f * j.frameL; p.a.R = j.dR + f * j.frameR; p.a.out = j.dOut + f * j.frameOut;
            p.a.pitchL = (int)j.pitchL; p.a.pitchR = (int)j.pitchR; p.a.pitchOut = (int)j.pitchOut;
            if (p.a.NC > 1) {
                p.a.gkey = gk;
                const size_t n = (size_t)j.w * j.h;
                sad_fill_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 4096), 256, 0, s>>>(gk, n, 0xFFFFFFFFu);
            }
            if (!kLaunchGeneric[p.half]) return SADGPU_EINVAL;         // developer build without this instance
            cudaError_t e = kLaunchGeneric[p.half](p, s, &c->attr_done[dev_index][p.half]);
            if (e != cudaSuccess) return (int)e;
            if (p.a.NC > 1) {
                dim3 g(ceil_div(j.w, 256), j.y1 - j.y0, 1);
                sad_finalize_kernel<<<g, 256, 0, s>>>(gk, p.a.out, j.w, j.h, j.y0, j.y1, (int)j.pitchOut, 0, j.D);
                if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
            }
            launches += p.launches;
        }
    }
    c->last_launches.store(launches);
    return SADGPU_OK;
}

int check_io(sadgpu_ctx* c, int stream, const uint8_t* l, int ls, const uint8_t* r, int rs, int w, int h)
{
    if (!c || !l || !r) return SADGPU_EINVAL;
    if (stream < 0 || stream >= (int)c->slots.size()) return SADGPU_ERANGE;
    if (w <= 0 || h <= 0 || ls < w || rs < w) return SADGPU_EINVAL;
    if (w > c->max_w || h > c->max_h) return SADGPU_ERANGE;
    return SADGPU_OK;
}

// Stage rows [ys,ye) of a host image and enqueue the H2D copy.  Pinned pool memory is uploaded in place.
int upload(sadgpu_ctx* c, Slot* s, const uint8_t* src, int stride, uint8_t* pinned, uint8_t* dev, int w, int ys, int ye)
{
    const int n = ye - ys;
    if (n <= 0) return SADGPU_OK;
    const uint8_t* from = src + (size_t)ys * stride;
    size_t from_pitch = (size_t)stride;
    if (!in_pool(c, from, (size_t)(n - 1) * stride + w)) {
        uint8_t* st = pinned + (size_t)ys * s->pitch;
        if ((size_t)stride == s->pitch && s->pitch == (size_t)w) c->copier->copy(st, (size_t)w, from, (size_t)w, (size_t)w, (size_t)n);
        else c->copier->copy(st, s->pitch, from, (size_t)stride, (size_t)w, (size_t)n);
        from = st; from_pitch = s->pitch;
    }
    cudaError_t e;
    if (from_pitch == s->pitch && s->pitch == (size_t)w)        // one contiguous DMA
        e = cudaMemcpyAsync(dev + (size_t)ys * s->pitch, from, (size_t)n * w, cudaMemcpyHostToDevice, s->st);
    else
        e = cudaMemcpy2DAsync(dev + (size_t)ys * s->pitch, s->pitch, from, from_pitch, (size_t)w, (size_t)n,
                              cudaMemcpyHostToDevice, s->st);
    return e == cudaSuccess ? SADGPU_OK : (int)e;
}

int finish_submit(sadgpu_ctx* c, Slot* s, int w, int h, int B, int D, int y0, int y1, uint8_t* direct_out, int direct_stride);

int submit_locked(sadgpu_ctx* c, Slot* s, const uint8_t* l, int ls, const uint8_t* r, int rs,
                  int w, int h, int B, int D, int y0, int y1, uint8_t* direct_out, int direct_stride)
{
    cudaError_t e = cudaSetDevice(s->device);
    if (e != cudaSuccess) return (int)e;
    int rc = validate(w, h, B, D, y0, y1);
    if (rc) return rc;
    const int half = B / 2;
    const int ys = std::max(0, y0 - half), ye = std::min(h, y1 + half);
    s->pitch = (size_t)round_up(w, 4);
    s->dR = s->dL + s->pitch * (size_t)h;                      // right image directly behind the left one
    s->hR = s->hL + s->pitch * (size_t)h;
    const size_t img = s->pitch * (size_t)h;
    const bool whole = ys == 0 && ye == h && s->pitch == (size_t)w && ls == w && rs == w;
    if (whole && r == l + img && in_pool(c, l, 2 * img)) {     // caller's pinned pair is contiguous: one DMA, zero staging
        cudaError_t e1 = cudaMemcpyAsync(s->dL, l, 2 * img, cudaMemcpyHostToDevice, s->st);
        if (e1 != cudaSuccess) return (int)e1;
    } else if (whole && !in_pool(c, l, img) && !in_pool(c, r, img)) {   // pageable pair: stage both, one DMA
        // pageable pair: stage the left plane, start its DMA, stage the right plane meanwhile
        c->copier->copy(s->hL, img, l, img, img, 1);
        cudaError_t e1 = cudaMemcpyAsync(s->dL, s->hL, img, cudaMemcpyHostToDevice, s->st);
        if (e1 != cudaSuccess) return (int)e1;
        c->copier->copy(s->hR, img, r, img, img, 1);
        e1 = cudaMemcpyAsync(s->dR, s->hR, img, cudaMemcpyHostToDevice, s->st);
        if (e1 != cudaSuccess) return (int)e1;
    } else {
        if ((rc = upload(c, s, l, ls, s->hL, s->dL, w, ys, ye))) return rc;
        if ((rc = upload(c, s, r, rs, s->hR, s->dR, w, ys, ye))) return rc;
    }
    return finish_submit(c, s, w, h, B, D, y0, y1, direct_out, direct_stride);
}

// Second half of a submit: the frame pair is (being) uploaded to s->dL / s->dR on the slot's stream; enqueue the kernels
// and the download of rows [y0,y1).
int finish_submit(sadgpu_ctx* c, Slot* s, int w, int h, int B, int D, int y0, int y1, uint8_t* direct_out, int direct_stride)
{
    cudaError_t e = cudaSuccess;
    int rc = SADGPU_OK;
    s->out_direct = direct_out != nullptr;
    if (y1 > y0) {
        Job j{s->dL, s->pitch, 0, s->dR, s->pitch, 0, s->dOut, s->pitch, 0, 1, w, h, B, D, y0, y1};
        if ((rc = run_job(c, s->dev_index, j, nullptr, s->gkey, s->st))) return rc;
        uint8_t* dst = direct_out ? direct_out + (size_t)y0 * direct_stride : s->hOut + (size_t)y0 * s->pitch;
        const size_t dpitch = direct_out ? (size_t)direct_stride : s->pitch;
        if (dpitch == s->pitch && s->pitch == (size_t)w)
            e = cudaMemcpyAsync(dst, s->dOut + (size_t)y0 * s->pitch, (size_t)(y1 - y0) * w, cudaMemcpyDeviceToHost, s->st);
        else
            e = cudaMemcpy2DAsync(dst, dpitch, s->dOut + (size_t)y0 * s->pitch, s->pitch, (size_t)w, (size_t)(y1 - y0),
                                  cudaMemcpyDeviceToHost, s->st);
        if (e != cudaSuccess) return (int)e;
    }
    e = cudaEventRecord(s->done, s->st);
    if (e != cudaSuccess) return (int)e;
    s->busy = true; s->w = w; s->h = h; s->y0 = y0; s->y1 = y1;
    return SADGPU_OK;
}

int wait_locked(sadgpu_ctx* c, Slot* s, uint8_t* out, int out_stride)
{
    cudaError_t e = cudaSetDevice(s->device);
    if (e == cudaSuccess) e = cudaEventSynchronize(s->done);
    s->busy = false;
    if (e != cudaSuccess) return (int)e;
    if (!s->out_direct) {
        if (!out || out_stride < s->w) return SADGPU_EINVAL;
        if ((size_t)out_stride == s->pitch && s->pitch == (size_t)s->w)          // one contiguous block
            c->copier->copy(out + (size_t)s->y0 * out_stride, (size_t)s->w, s->hOut + (size_t)s->y0 * s->pitch, (size_t)s->w,
                            (size_t)s->w, (size_t)(s->y1 - s->y0));
        else
            c->copier->copy(out + (size_t)s->y0 * out_stride, (size_t)out_stride, s->hOut + (size_t)s->y0 * s->pitch, s->pitch,
                            (size_t)s->w, (size_t)(s->y1 - s->y0));
    }
    return SADGPU_OK;
}

// (Re)allocates the pinned and device buffers of a slot for `cap` frame pairs of max_w x max_h.
cudaError_t alloc_slot_buffers(Slot* s, int max_w, int max_h, int cap)
{
    const size_t img = (size_t)round_up(max_w, 256) * (size_t)max_h;
    cudaFreeHost(s->hL); cudaFreeHost(s->hOut); cudaFree(s->dL); cudaFree(s->dOut); cudaFree(s->gkey);
    s->hL = s->hR = s->hOut = s->dL = s->dR = s->dOut = nullptr; s->gkey = nullptr;
    cudaGetLastError();
    // left and right live back to back (pinned and device) so that a whole frame pair is ONE DMA
    cudaError_t e = cudaHostAlloc((void**)&s->hL, 2 * img * cap, cudaHostAllocPortable);
    if (e == cudaSuccess) s->hR = s->hL + img;
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&s->hOut, img * cap, cudaHostAllocPortable);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->dL, 2 * img * cap);
    if (e == cudaSuccess) s->dR = s->dL + img;
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->dOut, img * cap);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->gkey, (size_t)max_w * max_h * sizeof(uint32_t) * cap);
    if (e == cudaSuccess) s->cap = cap;
    return e;
}

void free_slot(Slot* s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->st) { cudaStreamSynchronize(s->st); cudaStreamDestroy(s->st); }
    if (s->done) cudaEventDestroy(s->done);
    cudaFreeHost(s->hL); cudaFreeHost(s->hOut); cudaFreeHost(s->hRGBA);
    cudaFree(s->dL); cudaFree(s->dOut); cudaFree(s->gkey); cudaFree(s->dRGBA);
    delete s;
}

}  // namespace

extern "C" {

int sadgpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int sadgpu_create(const int* devices, int n_devices, int max_w, int max_h, int n_streams, sadgpu_ctx** out)
{
    if (!out || n_devices < 1 || n_devices > kMaxDevices || max_w <= 0 || max_h <= 0 || n_streams < 1)
        return SADGPU_EINVAL;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess) { cudaGetLastError(); return ndev == 0 ? SADGPU_ENODEV : (int)e; }
    if (ndev == 0) return SADGPU_ENODEV;
    sadgpu_ctx* c = new (std::nothrow) sadgpu_ctx();
    if (!c) return SADGPU_ENOMEM;
    memset(c->attr_done, 0, sizeof(c->attr_done));
    memset(c->fast_attr_done, 0, sizeof(c->fast_attr_done));
    memset(c->vh_attr_done, 0, sizeof(c->vh_attr_done));
    memset(c->ring_attr_done, 0, sizeof(c->ring_attr_done));
    c->max_w = max_w; c->max_h = max_h;
    for (int i = 0; i < n_devices; ++i) {
        const int d = devices ? devices[i] : i;
        if (d < 0 || d >= ndev) { delete c; return SADGPU_ERANGE; }
        cudaDeviceProp prop;
        if ((e = cudaGetDeviceProperties(&prop, d)) != cudaSuccess) { delete c; return (int)e; }
        if (prop.major != 10) { delete c; return SADGPU_ENODEV; }     // sm_100a cubin only, no fallback
        c->devices.push_back(d);
        c->sm_count.push_back(prop.multiProcessorCount);
    }
    c->copier = new (std::nothrow) CopyPool((int)std::min(3u, std::max(1u, std::thread::hardware_concurrency() / 4)));
    if (!c->copier) { delete c; return SADGPU_ENOMEM; }
    c->dev_gkey.assign(n_devices, nullptr);
    c->dev_gkey_bytes.assign(n_devices, 0);
    const size_t pitch = (size_t)round_up(max_w, 256);
    for (int i = 0; i < n_streams; ++i) {
        Slot* s = new (std::nothrow) Slot();
        if (!s) { sadgpu_destroy(c); return SADGPU_ENOMEM; }
        c->slots.push_back(s);
        s->dev_index = i % n_devices; s->device = c->devices[s->dev_index]; s->pitch = pitch;   // re-set per frame
        e = cudaSetDevice(s->device);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = alloc_slot_buffers(s, max_w, max_h, 1);
        if (e != cudaSuccess) { sadgpu_destroy(c); return (int)e; }
    }
    *out = c;
    return SADGPU_OK;
}

void sadgpu_destroy(sadgpu_ctx* c)
{
    if (!c) return;
    for (Slot* s : c->slots) free_slot(s);
    for (size_t i = 0; i < c->dev_gkey.size(); ++i)
        if (c->dev_gkey[i]) { cudaSetDevice(c->devices[i]); cudaFree(c->dev_gkey[i]); }
    for (auto& r : c->pool) cudaFreeHost(r.first);
    delete c->copier;
    delete c;
}

int sadgpu_compute(sadgpu_ctx* c, int stream, const uint8_t* l, int ls, const uint8_t* r, int rs,
                   int w, int h, int B, int D, int y0, int y1, uint8_t* out, int out_stride)
{
    int rc = check_io(c, stream, l, ls, r, rs, w, h);
    if (rc) return rc;
    if (!out || out_stride < w) return SADGPU_EINVAL;
    if ((rc = validate(w, h, B, D, y0, y1))) return rc;
    Slot* s = c->slots[stream];
    std::lock_guard<std::mutex> g(s->mu);
    if (s->busy) return SADGPU_EBUSY;
    uint8_t* direct = nullptr;
    if (y1 > y0 && in_pool(c, out + (size_t)y0 * out_stride, (size_t)(y1 - y0 - 1) * out_stride + w)) direct = out;
    rc = submit_locked(c, s, l, ls, r, rs, w, h, B, D, y0, y1, direct, out_stride);
    if (rc) { cudaStreamSynchronize(s->st); s->busy = false; return rc; }
    return wait_locked(c, s, out, out_stride);
}

int sadgpu_compute_nrgba(sadgpu_ctx* c, int stream, const uint8_t* l, int ls, const uint8_t* r, int rs,
                         int w, int h, int B, int D, uint8_t* out, int out_stride)
{
    if (!c || !l || !r || !out) return SADGPU_EINVAL;
    if (stream < 0 || stream >= (int)c->slots.size()) return SADGPU_ERANGE;
    if (w <= 0 || h <= 0 || ls < 4 * w || rs < 4 * w || out_stride < w) return SADGPU_EINVAL;
    if (w > c->max_w || h > c->max_h) return SADGPU_ERANGE;
    int rc = validate(w, h, B, D, 0, h);
    if (rc) return rc;
    Slot* s = c->slots[stream];
    std::lock_guard<std::mutex> g(s->mu);
    if (s->busy) return SADGPU_EBUSY;
    cudaError_t e = cudaSetDevice(s->device);
    if (e != cudaSuccess) return (int)e;
    const size_t cpitch = (size_t)round_up(4 * w, 16), plane = cpitch * (size_t)h;     // 16-byte rows: vector loads in the luma kernel
    if (s->rgba_bytes < 2 * plane) {
        cudaStreamSynchronize(s->st);
        cudaFreeHost(s->hRGBA); cudaFree(s->dRGBA); s->hRGBA = s->dRGBA = nullptr; s->rgba_bytes = 0;
        e = cudaHostAlloc((void**)&s->hRGBA, 2 * plane, cudaHostAllocPortable);
        if (e == cudaSuccess) e = cudaMalloc((void**)&s->dRGBA, 2 * plane);
        if (e != cudaSuccess) return (int)e;
        s->rgba_bytes = 2 * plane;
    }
    s->pitch = (size_t)round_up(w, 4);
    s->dR = s->dL + s->pitch * (size_t)h;
    const uint8_t* src[2] = {l, r};
    const int stride[2] = {ls, rs};
    uint8_t* dgray[2] = {s->dL, s->dR};
    for (int i = 0; i < 2; ++i) {                                  // stage (or take from the pinned pool), upload, luma on the device
        const uint8_t* from = src[i];
        size_t fpitch = (size_t)stride[i];
        if (!in_pool(c, from, (size_t)(h - 1) * stride[i] + 4 * (size_t)w)) {
            c->copier->copy(s->hRGBA + i * plane, cpitch, from, (size_t)stride[i], 4 * (size_t)w, (size_t)h);
            from = s->hRGBA + i * plane; fpitch = cpitch;
        }
        e = cudaMemcpy2DAsync(s->dRGBA + i * plane, cpitch, from, fpitch, 4 * (size_t)w, (size_t)h, cudaMemcpyHostToDevice, s->st);
        if (e != cudaSuccess) return (int)e;
        dim3 block(128), grid(ceil_div(ceil_div(w, 4), 128), ceil_div(h, kGrayRows));
        gray_kernel<GRAY_NRGBA8, 4><<<grid, block, 0, s->st>>>(s->dRGBA + i * plane, cpitch, dgray[i], s->pitch, w, h, 1);
        if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
    }
    uint8_t* direct = in_pool(c, out, (size_t)(h - 1) * out_stride + w) ? out : nullptr;
    rc = finish_submit(c, s, w, h, B, D, 0, h, direct, out_stride);
    if (rc) { cudaStreamSynchronize(s->st); s->busy = false; return rc; }
    return wait_locked(c, s, out, out_stride);
}

int sadgpu_submit(sadgpu_ctx* c, int stream, const uint8_t* l, int ls, const uint8_t* r, int rs,
                  int w, int h, int B, int D, int y0, int y1, uint64_t* ticket)
{
    int rc = check_io(c, stream, l, ls, r, rs, w, h);
    if (rc) return rc;
    if (!ticket) return SADGPU_EINVAL;
    if ((rc = validate(w, h, B, D, y0, y1))) return rc;
    Slot* s = c->slots[stream];
    std::lock_guard<std::mutex> g(s->mu);
    if (s->busy) return SADGPU_EBUSY;
    rc = submit_locked(c, s, l, ls, r, rs, w, h, B, D, y0, y1, nullptr, 0);
    if (rc) { cudaStreamSynchronize(s->st); s->busy = false; return rc; }
    s->seq++;
    *ticket = (s->seq << 16) | (uint64_t)stream;
    return SADGPU_OK;
}

int sadgpu_submit_into(sadgpu_ctx* c, int stream, const uint8_t* l, int ls, const uint8_t* r, int rs,
                       int w, int h, int B, int D, int y0, int y1, uint8_t* out, int out_stride, uint64_t* ticket)
{
    int rc = check_io(c, stream, l, ls, r, rs, w, h);
    if (rc) return rc;
    if (!ticket || !out || out_stride < w) return SADGPU_EINVAL;
    if ((rc = validate(w, h, B, D, y0, y1))) return rc;
    // the destination is retained until sadgpu_wait: only the context's own pinned memory qualifies (cgo pointer rule)
    if (y1 > y0 && !in_pool(c, out + (size_t)y0 * out_stride, (size_t)(y1 - y0 - 1) * out_stride + w)) return SADGPU_EINVAL;
    Slot* s = c->slots[stream];
    std::lock_guard<std::mutex> g(s->mu);
    if (s->busy) return SADGPU_EBUSY;
    rc = submit_locked(c, s, l, ls, r, rs, w, h, B, D, y0, y1, out, out_stride);
    if (rc) { cudaStreamSynchronize(s->st); s->busy = false; return rc; }
    s->seq++;
    *ticket = (s->seq << 16) | (uint64_t)stream;
    return SADGPU_OK;
}

int sadgpu_reserve_batch(sadgpu_ctx* c, int max_frames)
{
    if (!c || max_frames < 1 || max_frames > 256) return SADGPU_EINVAL;
    for (Slot* s : c->slots) {
        std::lock_guard<std::mutex> g(s->mu);
        if (s->busy) return SADGPU_EBUSY;
        if (s->cap >= max_frames) continue;
        cudaError_t e = cudaSetDevice(s->device);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s->st);
        if (e == cudaSuccess) e = alloc_slot_buffers(s, c->max_w, c->max_h, max_frames);
        if (e != cudaSuccess) return (int)e;
    }
    return SADGPU_OK;
}

int sadgpu_submit_batch_into(sadgpu_ctx* c, int stream, int n_frames, const uint8_t* pairs, int w, int h, int B, int D,
                             uint8_t* out, uint64_t* ticket)
{
    if (!c || !pairs || !out || !ticket || n_frames < 1) return SADGPU_EINVAL;
    if (stream < 0 || stream >= (int)c->slots.size()) return SADGPU_ERANGE;
    if (w <= 0 || h <= 0 || w % 4) return SADGPU_EINVAL;           // contiguous frames: the pitch is w
    if (w > c->max_w || h > c->max_h) return SADGPU_ERANGE;
    int rc = validate(w, h, B, D, 0, h);
    if (rc) return rc;
    const size_t img = (size_t)w * h;
    if (!in_pool(c, out, img * n_frames)) return SADGPU_EINVAL;    // the destination is retained until sadgpu_wait (cgo pointer rule)
    Slot* s = c->slots[stream];
    std::lock_guard<std::mutex> g(s->mu);
    if (s->busy) return SADGPU_EBUSY;
    if (n_frames > 256) return SADGPU_ERANGE;
    cudaError_t e = cudaSetDevice(s->device);
    if (e != cudaSuccess) return (int)e;
    if (n_frames > s->cap) {                                       // not pre-sized by sadgpu_reserve_batch: grow this stream's buffers now
        e = cudaStreamSynchronize(s->st);
        if (e == cudaSuccess) e = alloc_slot_buffers(s, c->max_w, c->max_h, n_frames);
        if (e != cudaSuccess) return (int)e;
    }
    const uint8_t* src = pairs;
    if (!in_pool(c, pairs, 2 * img * n_frames)) {                  // pageable source: staged
        c->copier->copy(s->hL, 2 * img * n_frames, pairs, 2 * img * n_frames, 2 * img * n_frames, 1);
        src = s->hL;
    }
    e = cudaMemcpyAsync(s->dL, src, 2 * img * n_frames, cudaMemcpyHostToDevice, s->st);                     // ONE DMA for the batch
    if (e != cudaSuccess) return (int)e;
    Job j{s->dL, (size_t)w, (long long)(2 * img), s->dL + img, (size_t)w, (long long)(2 * img), s->dOut, (size_t)w, (long long)img,
          n_frames, w, h, B, D, 0, h};
    if ((rc = run_job(c, s->dev_index, j, nullptr, s->gkey, s->st, s->cap))) { cudaStreamSynchronize(s->st); return rc; }
    e = cudaMemcpyAsync(out, s->dOut, img * n_frames, cudaMemcpyDeviceToHost, s->st);
    if (e == cudaSuccess) e = cudaEventRecord(s->done, s->st);
    if (e != cudaSuccess) { cudaStreamSynchronize(s->st); return (int)e; }
    s->busy = true; s->out_direct = true; s->w = w; s->h = h; s->y0 = 0; s->y1 = h; s->nfr = n_frames;
    s->seq++;
    *ticket = (s->seq << 16) | (uint64_t)stream;
    return SADGPU_OK;
}

int sadgpu_wait(sadgpu_ctx* c, uint64_t ticket, uint8_t* out, int out_stride)
{
    if (!c) return SADGPU_EINVAL;
    const int stream = (int)(ticket & 0xFFFF);
    if (stream >= (int)c->slots.size()) return SADGPU_EBUSY;
    Slot* s = c->slots[stream];
    std::lock_guard<std::mutex> g(s->mu);
    if (!s->busy || s->seq != (ticket >> 16)) return SADGPU_EBUSY;
    return wait_locked(c, s, out, out_stride);
}

int sadgpu_compute_sharded(sadgpu_ctx* c, const uint8_t* l, int ls, const uint8_t* r, int rs,
                           int w, int h, int B, int D, uint8_t* out, int out_stride)
{
    if (!c) return SADGPU_EINVAL;
    // one band per device; with spare streams up to four bands per device, so that on each device the upload of a band
    // overlaps the kernel of the previous one (stream s lives on device s % n_devices)
    const int n = (int)std::min(c->slots.size(), c->devices.size() * 4);
    int rc = check_io(c, 0, l, ls, r, rs, w, h);
    if (rc) return rc;
    if (!out || out_stride < w) return SADGPU_EINVAL;
    if ((rc = validate(w, h, B, D, 0, h))) return rc;
    const bool direct = in_pool(c, out, (size_t)(h - 1) * out_stride + w);      // pinned destination: the D2H copies land in it
    std::vector<int> started(n, 0);
    int first_err = 0;
    for (int i = 0; i < n; ++i) {                       // row band i -> stream i, halo handled by the upload
        const int y0 = (int)((long)h * i / n), y1 = (int)((long)h * (i + 1) / n);
        Slot* s = c->slots[i];
        std::lock_guard<std::mutex> g(s->mu);
        if (s->busy) { first_err = SADGPU_EBUSY; break; }
        rc = submit_locked(c, s, l, ls, r, rs, w, h, B, D, y0, y1, direct ? out : nullptr, out_stride);
        if (rc) { cudaStreamSynchronize(s->st); s->busy = false; first_err = rc; break; }
        started[i] = 1;
    }
    for (int i = 0; i < n; ++i) {                       // host-side gather: disjoint rows of one Pix
        if (!started[i]) continue;
        Slot* s = c->slots[i];
        std::lock_guard<std::mutex> g(s->mu);
        rc = wait_locked(c, s, out, out_stride);
        if (rc && !first_err) first_err = rc;
    }
    return first_err;
}

int sadgpu_compute_device_batch(sadgpu_ctx* c, int device, int n_frames,
                                const uint8_t* dL, size_t pitch_l, size_t frame_stride_l,
                                const uint8_t* dR, size_t pitch_r, size_t frame_stride_r,
                                int w, int h, int B, int D, int y0, int y1,
                                uint8_t* dOut, size_t pitch_out, size_t frame_stride_out,
                                void* cuda_stream, const sadgpu_tuning* tuning)
{
    if (!c || !dL || !dR || !dOut || n_frames < 1) return SADGPU_EINVAL;
    if (device < 0 || device >= (int)c->devices.size()) return SADGPU_ERANGE;
    if (pitch_l < (size_t)w || pitch_r < (size_t)w || pitch_out < (size_t)w) return SADGPU_EINVAL;
    int rc = validate(w, h, B, D, y0, y1);
    if (rc) return rc;
    cudaError_t e = cudaSetDevice(c->devices[device]);
    if (e != cudaSuccess) return (int)e;
    Job j{dL, pitch_l, (long long)frame_stride_l, dR, pitch_r, (long long)frame_stride_r,
          dOut, pitch_out, (long long)frame_stride_out, n_frames, w, h, B, D, y0, y1};
    return run_job(c, device, j, tuning, nullptr, (cudaStream_t)cuda_stream);
}

int sadgpu_compute_device(sadgpu_ctx* c, int device, const uint8_t* dL, size_t pitch_l, const uint8_t* dR, size_t pitch_r,
                          int w, int h, int B, int D, int y0, int y1, uint8_t* dOut, size_t pitch_out,
                          void* cuda_stream, const sadgpu_tuning* tuning)
{
    return sadgpu_compute_device_batch(c, device, 1, dL, pitch_l, 0, dR, pitch_r, 0, w, h, B, D, y0, y1,
                                       dOut, pitch_out, 0, cuda_stream, tuning);
}

int sadgpu_gray_device(sadgpu_ctx* c, int device, const uint8_t* dSrc, size_t src_pitch, int channels, int mode,
                       int w, int h, uint8_t* dGray, size_t gray_pitch, void* cuda_stream)
{
    if (!c || !dSrc || !dGray || w <= 0 || h <= 0) return SADGPU_EINVAL;
    if (device < 0 || device >= (int)c->devices.size()) return SADGPU_ERANGE;
    if ((channels != 3 && channels != 4) || mode < 0 || mode > 2) return SADGPU_EINVAL;
    if (mode == GRAY_RGB8_INTENDED && channels != 3) return SADGPU_EINVAL;
    if (src_pitch < (size_t)w * channels || gray_pitch < (size_t)w) return SADGPU_EINVAL;
    cudaError_t e = cudaSetDevice(c->devices[device]);
    if (e != cudaSuccess) return (int)e;
    cudaStream_t s = (cudaStream_t)cuda_stream;
    const int vec_ok = ((uintptr_t)dSrc % 16 == 0 && src_pitch % 16 == 0 && (uintptr_t)dGray % 4 == 0 && gray_pitch % 4 == 0) ? 1 : 0;
    dim3 block(128), grid(ceil_div(ceil_div(w, 4), 128), ceil_div(h, kGrayRows));
    if (channels == 4 && mode == GRAY_NRGBA8)             gray_kernel<GRAY_NRGBA8, 4><<<grid, block, 0, s>>>(dSrc, src_pitch, dGray, gray_pitch, w, h, vec_ok);
    else if (channels == 3 && mode == GRAY_NRGBA8)        gray_kernel<GRAY_NRGBA8, 3><<<grid, block, 0, s>>>(dSrc, src_pitch, dGray, gray_pitch, w, h, vec_ok);
    else if (channels == 3 && mode == GRAY_RGB8_INTENDED) gray_kernel<GRAY_RGB8_INTENDED, 3><<<grid, block, 0, s>>>(dSrc, src_pitch, dGray, gray_pitch, w, h, vec_ok);
    else if (channels == 4)                               gray_kernel<GRAY_RGBX8_LOADPNG, 4><<<grid, block, 0, s>>>(dSrc, src_pitch, dGray, gray_pitch, w, h, vec_ok);
    else                                                  gray_kernel<GRAY_RGBX8_LOADPNG, 3><<<grid, block, 0, s>>>(dSrc, src_pitch, dGray, gray_pitch, w, h, vec_ok);
    e = cudaGetLastError();
    return e == cudaSuccess ? SADGPU_OK : (int)e;
}

void* sadgpu_host_alloc(sadgpu_ctx* c, size_t bytes)
{
    if (!c || bytes == 0) return nullptr;
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    std::lock_guard<std::mutex> g(c->pool_mu);
    c->pool.emplace_back(static_cast<uint8_t*>(p), bytes);
    return p;
}

void sadgpu_host_free(sadgpu_ctx* c, void* p)
{
    if (!c || !p) return;
    std::lock_guard<std::mutex> g(c->pool_mu);
    for (size_t i = 0; i < c->pool.size(); ++i)
        if (c->pool[i].first == p) { cudaFreeHost(p); c->pool.erase(c->pool.begin() + i); return; }
}

int sadgpu_debug_read(sadgpu_ctx* c, int device, uint32_t* host, int n_words)
{
    if (!c || !host || device < 0 || device >= (int)c->devices.size() || n_words < 0) return SADGPU_EINVAL;
    if (!c->dev_gkey[device] || c->dev_gkey_bytes[device] < (size_t)n_words * 4) return SADGPU_ERANGE;
    cudaError_t e = cudaSetDevice(c->devices[device]);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(host, c->dev_gkey[device], (size_t)n_words * 4, cudaMemcpyDeviceToHost);
    return e == cudaSuccess ? SADGPU_OK : (int)e;
}

int sadgpu_last_launch_count(sadgpu_ctx* c) { return c ? c->last_launches.load() : 0; }

int sadgpu_plan_describe(int w, int h, int B, int D, int y0, int y1, const sadgpu_tuning* t, char* buf, size_t buflen)
{
    const int variant = t ? t->kernel_variant : 0;
    if (variant > 6 || (variant == 2 && !fast_supported(B)) || (variant == 4 && !wide_supported(B)) || (variant == 5 && !vh_supported(B)) ||
        (variant == 6 && !ring_supported(B)))
        return SADGPU_EINVAL;
    if (variant >= 2 || (variant == 0 && (fast_supported(B) || wide_supported(B)))) {
        FastPlan p; int slot = 0;
        const int nf = t && t->reserved[0] > 0 ? t->reserved[0] : 1;       // reserved[0]: frames per launch (describe only)
        int rc = make_fast_plan(w, h, B, D, y0, y1, nf, t, 148, &p, &slot, variant == 3 || variant == 0,
                                variant == 5 || (variant == 0 && vh_auto(B, D)), variant == 6 || (variant == 0 && ring_auto(B, D)));
        if (rc) return rc;
        if (variant == 3 && slot != 3) return SADGPU_EINVAL;
        if (buf && buflen)
            snprintf(buf, buflen,
                     "{\"variant\":\"%s\",\"half\":%d,\"NG\":%d,\"NC\":%d,\"NGc\":%d,\"TW\":%d,\"RB\":%d,\"BH\":%d,"
                     "\"grid\":[%u,%u,%u],\"threads\":%d,\"smem\":%zu,\"launches\":%d,\"frames_per_launch\":%d}",
                     slot == 6 ? "ring" : slot == 5 ? "vertical-first" : slot == 4 ? "wide" : slot == 3 ? "warp-specialised" : "fast", p.half, p.a.NG, p.a.NC, p.ngc, p.tw, p.rb, p.a.BH,
                     p.grid.x, p.grid.y, p.grid.z, p.nthreads, p.smem, p.launches, nf);
        return SADGPU_OK;
    }
    Plan p;
    int rc = make_plan(w, h, B, D, y0, y1, t, 148, &p);
    if (rc) return rc;
    if (buf && buflen)
        snprintf(buf, buflen,
                 "{\"variant\":\"generic\",\"half\":%d,\"NG\":%d,\"NC\":%d,\"NGc\":%d,\"K\":%d,\"TW\":%d,\"NSTEP\":%d,"
                 "\"RB\":%d,\"NR\":%d,\"BH\":%d,\"grid\":[%u,%u,%u],\"threads\":%d,\"smem\":%zu,\"launches\":%d}",
                 p.half, p.a.NG, p.a.NC, p.a.NGc, p.a.K, p.a.TW, p.a.NSTEP, p.a.RB, p.a.NR, p.a.BH,
                 p.grid.x, p.grid.y, p.grid.z, p.nthreads, p.smem, p.launches);
    return SADGPU_OK;
}

const char* sadgpu_strerror(int code)
{
    switch (code) {
        case SADGPU_OK: return "ok";
        case SADGPU_EINVAL: return "invalid argument (null pointer, stride < width, block_size not in 1..31 or max_disparity not in 1..256)";
        case SADGPU_ERANGE: return "argument out of range (image larger than the context, bad stream/device index or row range)";
        case SADGPU_ENOMEM: return "host allocation failed";
        case SADGPU_EBUSY: return "stream slot busy or stale ticket";
        case SADGPU_ENODEV: return "no usable sm_100 CUDA device (there is no CPU fallback)";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown sadgpu error";
    }
}

const char* sadgpu_version(void) { return "sadgpu 0.1 (sm_100a)"; }

}  // extern "C"
