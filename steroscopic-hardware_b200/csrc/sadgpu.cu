// sadgpu.cu — C-ABI shim of libsadgpu.so (include/sadgpu.h): context, per-stream pinned and
// device buffers, the launch planner and the kernel dispatch.  No CPU fallback exists: every
// compute entry point ends in a kernel launch or an error code.
#include "../../include/sadgpu.h"
#include "sad_common.cuh"
#include "sad_fast.cuh"
#include "sad_ws.cuh"
#include "sad_wide.cuh"
#include "sad_ring.cuh"
#include "sad_wsr.cuh"
#ifdef SADGPU_DEV_VARIANTS          // developer builds only: the first correct kernel and the vertical-first experiment (cross-checks)
#include "sad_kernels.cuh"
#include "sad_vh.cuh"
#endif
#include "gray_kernels.cuh"
#include "post_kernels.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

using namespace sadgpu;

namespace {

constexpr int kSmemBudget = 232448;      // 227 KB opt-in dynamic shared memory per CTA on sm_100
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

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }
// Row pitch of the shim's own device / pinned planes: a multiple of 16 bytes, so that TMA can address every frame that came through
// a host entry point whatever its width (the planes are allocated for round_up(max_w, 256) bytes per row).
inline int shim_pitch(int w) { return round_up(w, 16); }

int validate(int w, int h, int B, int D, int y0, int y1)
{
    if (w <= 0 || h <= 0) return SADGPU_EINVAL;
    if (B < 1 || B > SADGPU_MAX_BLOCK_SIZE || D < 1 || D > SADGPU_MAX_DISPARITY) return SADGPU_EINVAL;
    if (y0 < 0 || y1 > h || y0 > y1) return SADGPU_ERANGE;
    return SADGPU_OK;
}

// kernel_variant values of sadgpu_tuning
enum { V_AUTO = 0, V_GENERIC = 1, V_FAST = 2, V_WS = 3, V_WIDE = 4, V_VH = 5, V_RING = 6, V_WSR = 7 };

// ---------------------------------------------------------------------------------------
// Kernel table.  Every production kernel takes FastArgs and a grid (column strips, row bands, frames x chunks).
// ---------------------------------------------------------------------------------------
struct FastPlan {
    FastArgs a;
    dim3 grid;
    int nthreads;
    size_t smem;
    int half, ngc, rb, tw;
    int launches;
    int variant, mode;
    const struct FastEntry* fe;
};

typedef cudaError_t (*fast_fn)(const FastPlan&, cudaStream_t);
struct FastEntry { fast_fn fn; int nt, smem, rb, tw, ngc, lbox, rbox; };
constexpr FastEntry kNoEntry{nullptr, 0, 0, 0, 0, 0, 0, 0};

// cudaFuncSetAttribute is idempotent and cheap; a per-kernel once-flag keeps it off the per-frame path (one flag per
// device would be needed if the attribute were per device, but it belongs to the function in the primary context of each
// device, so the flag is indexed by device).
template <class K> cudaError_t ensure_smem(K k, int bytes, std::atomic<unsigned long long>* done_mask)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (done_mask->load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done_mask->fetch_or(bit, std::memory_order_release);
    return e;
}

template <int HALF, int NGC>
cudaError_t launch_fast(const FastPlan& p, cudaStream_t s)
{
    using C = FastCfg<HALF, NGC>;
    static_assert(C::SMEM <= kSmemBudget, "fast kernel does not fit shared memory");
    static std::atomic<unsigned long long> done{0};
    auto k = sad_fast_kernel<HALF, NGC>;
    cudaError_t e = ensure_smem(k, C::SMEM, &done);
    if (e != cudaSuccess) return e;
    k<<<p.grid, C::NT, C::SMEM, s>>>(p.a);
    return cudaGetLastError();
}

template <int HALF, int MODE>
cudaError_t launch_ws(const FastPlan& p, cudaStream_t s)
{
    using C = WsCfg<HALF, MODE>;
    static std::atomic<unsigned long long> done_tma{0}, done_plain{0};
    cudaError_t e = ensure_smem(sad_ws_kernel<HALF, MODE, true>, C::SMEM, &done_tma);
    if (e == cudaSuccess) e = ensure_smem(sad_ws_kernel<HALF, MODE, false>, C::SMEM, &done_plain);
    if (e != cudaSuccess) return e;
    if (p.a.use_tma) sad_ws_kernel<HALF, MODE, true><<<p.grid, C::NT, C::SMEM, s>>>(p.a);
    else             sad_ws_kernel<HALF, MODE, false><<<p.grid, C::NT, C::SMEM, s>>>(p.a);
    return cudaGetLastError();
}

template <int HALF>
cudaError_t launch_wide(const FastPlan& p, cudaStream_t s)
{
    using C = WideCfg<HALF>;
    static_assert(C::SMEM <= kSmemBudget, "wide kernel does not fit shared memory");
    static std::atomic<unsigned long long> done{0};
    auto k = sad_wide_kernel<HALF>;
    cudaError_t e = ensure_smem(k, C::SMEM, &done);
    if (e != cudaSuccess) return e;
    k<<<p.grid, C::NT, C::SMEM, s>>>(p.a);
    return cudaGetLastError();
}

template <int HALF>
cudaError_t launch_ring(const FastPlan& p, cudaStream_t s)
{
    using C = RingCfg<HALF>;
    static_assert(C::SMEM <= kSmemBudget, "ring kernel does not fit shared memory");
    static std::atomic<unsigned long long> done{0};
    auto k = sad_ring_kernel<HALF>;
    cudaError_t e = ensure_smem(k, C::SMEM, &done);
    if (e != cudaSuccess) return e;
    k<<<p.grid, C::NT, C::SMEM, s>>>(p.a);
    return cudaGetLastError();
}

template <int HALF, int NGC>
cudaError_t launch_wsr(const FastPlan& p, cudaStream_t s)
{
    using C = WsrCfg<HALF, NGC>;
    static_assert(C::SMEM <= kSmemBudget, "shared-memory-ring warp-specialised kernel does not fit shared memory");
    static std::atomic<unsigned long long> done{0};
    auto k = sad_wsr_kernel<HALF, NGC>;
    cudaError_t e = ensure_smem(k, C::SMEM, &done);
    if (e != cudaSuccess) return e;
    if (!p.a.use_tma) return cudaErrorInvalidValue;            // run_job re-plans onto another kernel when TMA cannot be used
    k<<<p.grid, C::NT, C::SMEM, s>>>(p.a);
    return cudaGetLastError();
}

template <int HALF, int NGC> constexpr FastEntry fast_entry()
{
    using C = FastCfg<HALF, NGC>;
    if constexpr (dev_on(HALF)) return FastEntry{launch_fast<HALF, NGC>, C::NT, C::SMEM, C::RB, C::TW, NGC, 0, 0};
    else return kNoEntry;
}
template <int HALF, int MODE> constexpr FastEntry ws_entry()
{
    using C = WsCfg<HALF, MODE>;
    if constexpr (dev_on(HALF)) return FastEntry{launch_ws<HALF, MODE>, C::NT, C::SMEM, C::RB, C::CW, C::NGC, C::LBOX, C::RWT * 4};
    else return kNoEntry;
}
template <int HALF> constexpr FastEntry wide_entry()
{
    using C = WideCfg<HALF>;
    if constexpr (dev_on(HALF)) return FastEntry{launch_wide<HALF>, C::NT, C::SMEM, C::RB, C::TW, C::NGC, 0, 0};
    else return kNoEntry;
}
template <int HALF> constexpr FastEntry ring_entry()
{
    using C = RingCfg<HALF>;
    if constexpr (dev_on(HALF)) return FastEntry{launch_ring<HALF>, C::NT, C::SMEM, 4, C::TW, C::NGC, 0, 0};
    else return kNoEntry;
}

template <int HALF, int NGC> constexpr FastEntry wsr_entry()
{
    using C = WsrCfg<HALF, NGC>;
    if constexpr (dev_on(HALF)) return FastEntry{launch_wsr<HALF, NGC>, C::NT, C::SMEM, C::RB, C::CW, C::NGC, C::LBOX, C::RWT * 4};
    else return kNoEntry;
}

// phase-alternating kernel (sad_fast.cuh): h <= 7; slot 0/1/2 = 9 / 17 (18 for h >= 5) / 33 groups per chunk; h >= 5 has no
// 33-group instance (shared memory)
const FastEntry kFast[8][3] = {
    {fast_entry<0, 9>(), fast_entry<0, 17>(), fast_entry<0, 33>()},
    {fast_entry<1, 9>(), fast_entry<1, 17>(), fast_entry<1, 33>()},
    {fast_entry<2, 9>(), fast_entry<2, 17>(), fast_entry<2, 33>()},
    {fast_entry<3, 9>(), fast_entry<3, 17>(), fast_entry<3, 33>()},
    {fast_entry<4, 9>(), fast_entry<4, 17>(), fast_entry<4, 33>()},
    {fast_entry<5, 9>(), fast_entry<5, 18>(), kNoEntry},
    {fast_entry<6, 9>(), fast_entry<6, 18>(), kNoEntry},
    {fast_entry<7, 9>(), fast_entry<7, 18>(), kNoEntry}};

// warp-specialised kernel (sad_ws.cuh): h <= 4; mode 0..3 = chunks of 33 / 17 / 9 / 5 groups on 1 / 2 / 3 / 6 strips per CTA
// h = 5..8: mode 1 only (17 groups on one strip, two rows per walker warp)
#define WS_ROW(H) {ws_entry<H, 0>(), ws_entry<H, 1>(), ws_entry<H, 2>(), ws_entry<H, 3>()}
#define WS_ROW17(H) {kNoEntry, ws_entry<H, 1>(), ws_entry<H, 2>(), kNoEntry}
#define WS_ROW17W(H) {kNoEntry, ws_entry<H, 1>(), kNoEntry, kNoEntry}
const FastEntry kWs[9][4] = {WS_ROW(0), WS_ROW(1), WS_ROW(2), WS_ROW(3), WS_ROW(4), WS_ROW17(5), WS_ROW17(6), WS_ROW17(7), WS_ROW17W(8)};

// large-window phase-alternating kernel (sad_wide.cuh): h = 8..15, chunks of 8 groups
const FastEntry kWide[8] = {wide_entry<8>(), wide_entry<9>(), wide_entry<10>(), wide_entry<11>(), wide_entry<12>(), wide_entry<13>(),
                            wide_entry<14>(), wide_entry<15>()};

// shared-memory-ring warp-specialised kernel (sad_ring.cuh): h = 5..15, 32-column strips, chunks of 33 groups (h <= 7) or 17
const FastEntry kRing[11] = {ring_entry<5>(), ring_entry<6>(), ring_entry<7>(), ring_entry<8>(), ring_entry<9>(), ring_entry<10>(),
                             ring_entry<11>(), ring_entry<12>(), ring_entry<13>(), ring_entry<14>(), ring_entry<15>()};

// warp-specialised kernel with a shared-memory ring (sad_wsr.cuh): h = 5..15, 32-column strips, chunks of 9 groups, and for
// h >= 9 also of 13 groups (7-row batches); needs TMA
#define WSR_ROW(H) {wsr_entry<H, 9>(), kNoEntry}
#define WSR_ROW2(H) {wsr_entry<H, 9>(), wsr_entry<H, 13>()}
const FastEntry kWsr[11][2] = {WSR_ROW(5), WSR_ROW(6), WSR_ROW(7), WSR_ROW(8), WSR_ROW2(9), WSR_ROW2(10), WSR_ROW2(11), WSR_ROW2(12),
                               WSR_ROW2(13), WSR_ROW2(14), WSR_ROW2(15)};

#ifdef SADGPU_DEV_VARIANTS
template <int HALF>
cudaError_t launch_vh(const FastPlan& p, cudaStream_t s)
{
    using C = VhCfg<HALF>;
    static_assert(C::SMEM <= kSmemBudget, "vertical-first kernel does not fit shared memory");
    static std::atomic<unsigned long long> done{0};
    auto k = sad_vh_kernel<HALF>;
    cudaError_t e = ensure_smem(k, C::SMEM, &done);
    if (e != cudaSuccess) return e;
    k<<<p.grid, C::NT, C::SMEM, s>>>(p.a);
    return cudaGetLastError();
}
template <int HALF> constexpr FastEntry vh_entry()
{
    using C = VhCfg<HALF>;
    if constexpr (dev_on(HALF)) return FastEntry{launch_vh<HALF>, C::NT, C::SMEM, C::CROWS, C::TW, 33, 0, 0};
    else return kNoEntry;
}
const FastEntry kVh[11] = {vh_entry<5>(), vh_entry<6>(), vh_entry<7>(), vh_entry<8>(), vh_entry<9>(), vh_entry<10>(), vh_entry<11>(),
                           vh_entry<12>(), vh_entry<13>(), vh_entry<14>(), vh_entry<15>()};
bool vh_supported(int B) { return B / 2 >= 5 && B / 2 <= 15; }
#else
bool vh_supported(int) { return false; }
#endif

bool ring_supported(int B) { return B / 2 >= 5 && B / 2 <= 15; }
bool fast_supported(int B) { return B / 2 <= 7; }
bool wide_supported(int B) { return B / 2 >= 8 && B / 2 <= 15; }
bool ws_supported(int B) { return B / 2 <= 8; }
bool wsr_supported(int B) { return B / 2 >= 5 && B / 2 <= 15; }
// relative cost of one pass of the shared-memory-ring kernel over a chunk of 9 / 13 groups (measured at block 21 and 31, 1080p:
// 36.8..39.5 against 51.5..57.1 us for one chunk, 31.1..31.5 against 45.7..48.9 us per chunk of a chunked range)
const int wsr_pass_cost[2] = {100, 150};

// Planner default for block_size >= 10, from the measured variant sweep (profiles/r01_variant_sweep.json, inputs streaming
// from HBM): a ring pass over 33 groups costs about 1.7x a pass of the phase-alternating kernel over 18 groups, a ring pass
// over 17 groups 1.95x a pass of the wide kernel over 8 groups; the ring kernel is chosen whenever its passes are cheaper in total.
bool ring_auto(int B, int D)
{
    if (!ring_supported(B)) return false;
    const int ng = (D + 4) / 4;
    if (B / 2 <= 7) return 172 * ((ng + 32) / 33) < 100 * ((ng + 17) / 18);
    return 195 * ((ng + 16) / 17) < 100 * ((ng + 7) / 8);
}

// Smallest warp-specialised mode whose chunk holds all ng groups (mode 0 chunks the range).
int ws_mode_for(int ng) { return ng <= 5 ? 3 : ng <= 9 ? 2 : ng <= 17 ? 1 : 0; }

// Resolves (block size, disparity range, tuning) to a kernel: variant, table entry and (for the warp-specialised kernel) mode.
int choose_kernel(int B, int D, const sadgpu_tuning* t, int* variant_out, int* mode_out, const FastEntry** fe_out)
{
    const int half = B / 2, ng = (D + 4) / 4;
    int variant = t ? t->kernel_variant : V_AUTO;
    const int gpc = t ? t->groups_per_chunk : 0;               // tests: force smaller chunks
    if (variant < 0 || variant > 7) return SADGPU_EINVAL;
    if (variant == V_AUTO) {
        // measured (profiles/r02_*): the warp-specialised kernel wins for every block_size <= 17 except where another kernel
        // covers the whole range in ONE pass that two 17-group chunks cannot beat: 18 groups (max_disparity 65..68) at block_size
        // 11..15 (phase-alternating kernel, 18-group instance) and <= 8 groups (max_disparity <= 28) at block_size 16, 17
        // block_size >= 18, and <= 9 groups at block_size 16, 17: the warp-specialised kernel with the shared-memory ring (it falls
        // back to the mbarrier-pipelined ring kernel in run_job when TMA cannot address the images)
        if (half <= 4) variant = V_WS;
        else if (half <= 7 && ng != 18) variant = V_WS;
        else if (half == 8 && ng > 9) variant = V_WS;
        else if (half >= 8) variant = V_WSR;
        else if (ring_auto(B, D)) variant = V_RING;
        else variant = fast_supported(B) ? V_FAST : V_WIDE;
    }
    int mode = 0;
    const FastEntry* fe = nullptr;
    switch (variant) {
    case V_WS:
        if (!ws_supported(B)) return SADGPU_EINVAL;
        mode = ws_mode_for(ng);
        if (gpc > 0) { const int forced = gpc >= 33 ? 0 : gpc >= 17 ? 1 : gpc >= 9 ? 2 : 3; mode = std::max(mode, forced); }
        if (half >= 5) mode = (half <= 7 && ng <= 9 && !(gpc >= 17)) ? 2 : 1;    // block_size 11..17: chunks of 17 groups, or 2 strips x 9 groups
        fe = &kWs[half][mode];
        break;
    case V_FAST: {
        if (!fast_supported(B)) return SADGPU_EINVAL;
        int slot = ng <= 9 ? 0 : ng <= (half >= 5 ? 18 : 17) ? 1 : 2;
        if (gpc > 0) slot = gpc <= 9 ? 0 : gpc <= 17 ? 1 : slot;
        if (!kFast[half][slot].fn && slot == 2) slot = 1;
        mode = slot;
        fe = &kFast[half][slot];
        break;
    }
    case V_WIDE:
        if (!wide_supported(B)) return SADGPU_EINVAL;
        fe = &kWide[half - 8];
        break;
    case V_RING:
        if (!ring_supported(B)) return SADGPU_EINVAL;
        fe = &kRing[half - 5];
        break;
    case V_WSR: {
        if (!wsr_supported(B)) return SADGPU_EINVAL;
        // chunks of 9 or 13 groups: the cheaper cover of the range by the measured cost of a pass (wsr_pass_cost); 13 wins for
        // 10..13, 37..39 and 65 groups (max_disparity 36..48, 144..152, 256)
        int slot = 0;
        if (gpc > 0) slot = gpc <= 9 ? 0 : 1;
        else if (kWsr[half - 5][1].fn && wsr_pass_cost[1] * ((ng + 12) / 13) < wsr_pass_cost[0] * ((ng + 8) / 9)) slot = 1;
        if (!kWsr[half - 5][slot].fn) slot = 0;
        mode = slot;
        fe = &kWsr[half - 5][slot];
        break;
    }
#ifdef SADGPU_DEV_VARIANTS
    case V_VH:
        if (!vh_supported(B)) return SADGPU_EINVAL;
        fe = &kVh[half - 5];
        break;
#endif
    default:
        return SADGPU_EINVAL;                                  // variants 1 and 5 exist in developer builds only
    }
    if (!fe->fn) return SADGPU_EINVAL;                         // developer build without this instance
    *variant_out = variant; *mode_out = mode; *fe_out = fe;
    return SADGPU_OK;
}

const char* variant_name(int v)
{
    switch (v) { case V_GENERIC: return "generic"; case V_FAST: return "fast"; case V_WS: return "warp-specialised"; case V_WIDE: return "wide";
                 case V_VH: return "vertical-first"; case V_RING: return "ring"; case V_WSR: return "warp-specialised, shared-memory ring"; default: return "?"; }
}

int make_fast_plan(int w, int h, int B, int D, int y0, int y1, int n_frames, const sadgpu_tuning* t, int sm_count, FastPlan* p)
{
    int rc = validate(w, h, B, D, y0, y1);
    if (rc) return rc;
    if (n_frames < 1) return SADGPU_EINVAL;
    const FastEntry* fe = nullptr;
    if ((rc = choose_kernel(B, D, t, &p->variant, &p->mode, &fe))) return rc;
    p->fe = fe;
    const int half = B / 2;
    FastArgs& a = p->a;
    memset(&a, 0, sizeof(a));
    a.W = w; a.H = h; a.y0 = y0; a.y1 = y1; a.D = D;
    a.NG = (D + 4) / 4;
    p->tw = fe->tw; p->half = half; p->ngc = fe->ngc; p->rb = fe->rb;
    a.NC = ceil_div(a.NG, p->ngc);
    p->nthreads = fe->nt; p->smem = fe->smem;
    const int rows = std::max(1, y1 - y0);
    const int nstrips = ceil_div(w, p->tw);
    int nbands = 1;
    if (t && t->band_rows > 0) {
        a.BH = std::min(std::max(1, t->band_rows), rows);
    } else {
        // minimise waves x rows-per-CTA (one CTA per SM: the shared-memory rings fill it)
        long best_cost = -1;
        for (int nb = 1; nb <= std::min(rows, 64); ++nb) {
            const int bh = ceil_div(rows, nb);
            const long ctas = (long)nstrips * nb * a.NC * n_frames;
            const long waves = (ctas + sm_count - 1) / sm_count;
            // rows a CTA spends on pipeline fill and drain, beyond its band and the window halo
            const int fill = (p->variant == V_WS || p->variant == V_WSR) ? 2 * fe->rb : (p->variant == V_RING || p->variant == V_VH) ? 12 : fe->rb / 2 + 2;
            const long cost = waves * (round_up(bh + 2 * half, fe->rb) + fill);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; nbands = nb; }
        }
        a.BH = ceil_div(rows, nbands);
    }
    nbands = ceil_div(rows, a.BH);
    p->grid = dim3(nstrips, nbands, a.NC * n_frames);
    p->launches = a.NC == 1 ? 1 : 3;
    return SADGPU_OK;
}

// ---- TMA descriptors for the warp-specialised kernel (driver entry point fetched through the runtime) ----
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
encode_tiled_fn lookup_encode_tiled()
{
    void* p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess)
        return (encode_tiled_fn)p;
    cudaGetLastError();
    return nullptr;
}
encode_tiled_fn get_encode_tiled()
{
    static const encode_tiled_fn fn = lookup_encode_tiled();     // initialised once, thread-safe
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

// Host-side staging copies (pageable caller memory <-> pinned buffers) are memory-bound single-thread memcpys of several
// megabytes per frame; a few helper threads cut them to a fraction.  The pool is created with the context.
class CopyPool {
public:
    explicit CopyPool(int helpers) : w_(helpers > 0 ? new Worker[helpers] : nullptr), n_(helpers)
    {
        for (int i = 0; i < n_; ++i) w_[i].th = std::thread([this, i] { run(w_[i]); });
    }
    ~CopyPool()
    {
        stop_.store(true);
        for (int i = 0; i < n_; ++i) {
            { std::lock_guard<std::mutex> g(w_[i].m); }
            w_[i].cv.notify_all();
            w_[i].th.join();
        }
        delete[] w_;
    }
    // rows x width bytes, row pitches dp / sp; contiguous when dp == sp == width.  Copies of 512 KB and more are shared with the
    // helpers that are idle right now (a helper that another caller holds is simply not used); small copies stay on the caller.
    void copy(uint8_t* dst, size_t dp, const uint8_t* src, size_t sp, size_t width, size_t rows)
    {
        if (rows == 0 || width == 0) return;
        const bool contig = dp == width && sp == width;
        const size_t total = width * rows;
        const int want = total < (512u << 10) ? 0 : (int)std::min<size_t>((size_t)n_, total / (256u << 10) - 1);
        Worker* got[8];
        int ngot = 0;
        for (int i = 0; i < n_ && ngot < want && ngot < 8; ++i) {
            int idle = IDLE;
            if (w_[i].state.compare_exchange_strong(idle, CLAIMED)) got[ngot++] = &w_[i];
        }
        if (ngot == 0) { one(dst, dp, src, sp, width, rows, contig); return; }
        const int parts = ngot + 1;
        std::atomic<int> latch{ngot};                       // lives on this stack: a helper's last access is the decrement
        for (int p = 1; p < parts; ++p) {
            Worker& w = *got[p - 1];
            if (contig) {
                const size_t b0 = total * p / parts, b1 = total * (p + 1) / parts;
                w.task = Task{dst + b0, src + b0, b1 - b0, b1 - b0, b1 - b0, 1, true};
            } else {
                const size_t r0 = rows * p / parts, r1 = rows * (p + 1) / parts;
                w.task = Task{dst + r0 * dp, src + r0 * sp, dp, sp, width, r1 - r0, false};
            }
            w.latch = &latch;
            w.state.store(READY);                            // seq_cst: ordered against the helper's `asleep` flag
            if (w.asleep.load()) { { std::lock_guard<std::mutex> g(w.m); } w.cv.notify_one(); }
        }
        if (contig) memcpy(dst, src, total / parts);
        else one(dst, dp, src, sp, width, rows / parts, false);
        for (int spin = 0; latch.load(std::memory_order_acquire) > 0; ++spin)
            if (spin > 2000) std::this_thread::yield();
    }
private:
    enum { IDLE = 0, CLAIMED = 1, READY = 2 };
    struct Task { uint8_t* dst; const uint8_t* src; size_t dp, sp, width, rows; bool contig; };
    struct alignas(64) Worker {
        std::atomic<int> state{IDLE};
        std::atomic<bool> asleep{false};
        Task task{};
        std::atomic<int>* latch = nullptr;
        std::mutex m;
        std::condition_variable cv;
        std::thread th;
    };
    static void one(uint8_t* dst, size_t dp, const uint8_t* src, size_t sp, size_t width, size_t rows, bool contig)
    {
        if (contig) { memcpy(dst, src, width * rows); return; }
        for (size_t y = 0; y < rows; ++y) memcpy(dst + y * dp, src + y * sp, width);
    }
    // A helper polls its mailbox for a while after each task (a frame is several copies a few hundred microseconds apart; waking a
    // sleeping thread costs tens of microseconds on the GPU hosts), then sleeps.
    void run(Worker& w)
    {
        for (;;) {
            bool ready = false;
            const auto t0 = std::chrono::steady_clock::now();
            for (int spin = 0; !ready && !stop_.load(std::memory_order_relaxed); ++spin) {
                ready = w.state.load(std::memory_order_acquire) == READY;
                if (!ready && (spin & 255) == 255 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(300)) break;
            }
            if (!ready) {
                std::unique_lock<std::mutex> l(w.m);
                w.asleep.store(true);
                w.cv.wait(l, [&] { return stop_.load() || w.state.load() == READY; });
                w.asleep.store(false);
                if (w.state.load() != READY) return;           // stop
            }
            const Task t = w.task;
            std::atomic<int>* latch = w.latch;
            one(t.dst, t.dp, t.src, t.sp, t.width, t.rows, t.contig);
            w.state.store(IDLE, std::memory_order_release);
            latch->fetch_sub(1, std::memory_order_release);
        }
    }
    Worker* w_;
    int n_;
    std::atomic<bool> stop_{false};
};

// A few dozen nanoseconds of critical section, taken ~300 times per frame by up to 32 threads: a contended std::mutex costs a
// system call per hand-over (1.6 us measured on the GPU hosts), a test-and-test-and-set lock 0.1 us.
class SpinLock {
public:
    void lock()
    {
        while (b_.exchange(true, std::memory_order_acquire))
            while (b_.load(std::memory_order_relaxed)) {
#if defined(__x86_64__) || defined(__i386__)
                __builtin_ia32_pause();
#endif
            }
    }
    void unlock() { b_.store(false, std::memory_order_release); }
private:
    std::atomic<bool> b_{false};
};

struct Slot {
    int dev_index = 0, device = 0;
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;
    cudaEvent_t uploaded = nullptr;                            // recorded behind the H2D copies of a submit: the caller's pinned source is free again
    uint8_t *hL = nullptr, *hR = nullptr, *hOut = nullptr;     // pinned staging
    uint8_t *dL = nullptr, *dR = nullptr, *dOut = nullptr;     // device, pitched
    uint32_t* gkey = nullptr;
    size_t pitch = 0;                                          // pitch of the frame in flight = shim_pitch(w): contiguous copies when the host stride matches
    std::mutex mu;
    bool busy = false;
    uint64_t seq = 0;
    int w = 0, h = 0, y0 = 0, y1 = 0;                          // frame in flight
    bool out_direct = false;
    uint8_t *hRGBA = nullptr, *dRGBA = nullptr;                // staging of the interleaved colour pair (sadgpu_compute_nrgba), lazily allocated
    size_t rgba_bytes = 0;
    uint8_t* dPost = nullptr;                                  // five scratch planes of the post-processing hooks (sadgpu_compute_checked), lazily allocated
    size_t post_bytes = 0;
    int cap = 1;                                               // frame pairs the buffers hold (sadgpu_reserve_batch)
    int nfr = 1;                                               // frames in flight
};

struct FrameEntry;

}  // namespace

struct sadgpu_ctx {
    std::vector<int> devices;
    std::vector<int> sm_count;
    std::vector<Slot*> slots;
    int max_w = 0, max_h = 0;
    std::atomic<int> last_launches{0};
    std::mutex pool_mu;
    std::vector<std::pair<uint8_t*, size_t>> pool;
    CopyPool* copier = nullptr;
    std::mutex dev_mu;
    std::vector<uint32_t*> dev_dbg;        // per device: 4 KB of developer counters (profile builds), allocated on first use
    // frame cache of sadgpu_compute_region: chunks of one frame pair share one whole-frame GPU pass
    SpinLock cache_mu;                     // protects the entries' bookkeeping (never held across a copy, a CUDA call or a sleep)
    std::mutex cache_sleep_mu;             // sleepers only
    std::condition_variable cache_cv;
    std::atomic<int> cache_sleepers{0};    // threads asleep on cache_cv (a wake-up call is only made for them)
    std::once_flag cache_once;
    int cache_rc = 0;
    std::atomic<int> cache_spinners{0};    // threads polling an entry's state
    std::vector<FrameEntry*> cache;
    uint64_t cache_tick = 0;
    std::atomic<long long> region_calls{0}, region_frames{0}, region_stale{0};
    std::atomic<uint64_t> verify_clock{1};  // orders chunk calls and snapshot verifications (rows_match_snapshot)
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

// Developer counters (profile builds): 4 KB per device, allocated once, never resized.
int ensure_dbg(sadgpu_ctx* c, int dev_index, uint32_t** out)
{
    std::lock_guard<std::mutex> g(c->dev_mu);
    if (!c->dev_dbg[dev_index]) {
        cudaError_t e = cudaMalloc((void**)&c->dev_dbg[dev_index], 4096);
        if (e != cudaSuccess) return (int)e;
    }
    *out = c->dev_dbg[dev_index];
    return SADGPU_OK;
}

// Enqueue one job (one frame, or a batch of frames) on stream s.  The device must be current.  Never a CPU path.
// A chunked disparity range meets in a key map: the slot's own (host entry points: one per stream slot) or, for the
// device-resident entry points, a stream-ordered allocation that lives exactly as long as this call's kernels — two calls
// on different streams of one device never share scratch.
int run_job(sadgpu_ctx* c, int dev_index, const Job& j, const sadgpu_tuning* t, uint32_t* slot_gkey, cudaStream_t s, int slot_gkey_frames = 1)
{
    FastPlan p;
    int rc = make_fast_plan(j.w, j.h, j.B, j.D, j.y0, j.y1, j.n_frames, t, c->sm_count[dev_index], &p);
    if (rc) return rc;
    if (j.y1 == j.y0) return SADGPU_OK;
    bool tma_ok = false;
    if ((p.variant == V_WS || p.variant == V_WSR) && !(t && t->reserved[2] == 1))      // reserved[2] == 1: force the non-TMA loader (tests)
        tma_ok = make_tmap(&p.a.tmapL, j.dL, j.w, j.h, j.pitchL, j.frameL, j.n_frames, p.fe->lbox, p.fe->rb) &&
                 make_tmap(&p.a.tmapR, j.dR, j.w, j.h, j.pitchR, j.frameR, j.n_frames, p.fe->rbox, p.fe->rb);
    if (p.variant == V_WSR && !tma_ok && t && t->reserved[2] == 2) return SADGPU_EINVAL;     // tests: no silent substitution
    if (p.variant == V_WSR && !tma_ok) {
        // the shared-memory-ring warp-specialised kernel has no plain-load path: images TMA cannot address (base, pitch or frame
        // stride not a multiple of 16 bytes) take the mbarrier-pipelined ring kernel
        sadgpu_tuning t2{};
        if (t) t2 = *t;
        // what the planner chose before this kernel existed: ring or wide by the measured pass costs
        t2.kernel_variant = ring_auto(j.B, j.D) ? V_RING : fast_supported(j.B) ? V_FAST : V_WIDE; t2.groups_per_chunk = 0;
        if ((rc = make_fast_plan(j.w, j.h, j.B, j.D, j.y0, j.y1, j.n_frames, &t2, c->sm_count[dev_index], &p))) return rc;
    }
    const FastEntry* fe = p.fe;
    const int variant = p.variant;
    FastArgs& a = p.a;
    a.L = j.dL; a.R = j.dR; a.out = j.dOut;
    a.pitchL = (int)j.pitchL; a.pitchR = (int)j.pitchR; a.pitchOut = (int)j.pitchOut;
    a.frameL = j.frameL; a.frameR = j.frameR; a.frameOut = j.frameOut;
    a.k65536 = 65536u;
    a.debug_skip = t ? t->reserved[1] : 0;
    a.use_tma = tma_ok ? 1 : 0;
    a.aligned = ((uintptr_t)j.dR % 4 == 0 && j.pitchR % 4 == 0 && j.frameR % 4 == 0) ? 1 : 0;
    if (a.debug_skip & 4) { uint32_t* gk = nullptr; rc = ensure_dbg(c, dev_index, &gk); if (rc) return rc; a.gkey = gk; }
    bool async_scratch = false;
    if (a.NC > 1) {
        const size_t n = (size_t)j.n_frames * j.w * j.h;
        uint32_t* gk = slot_gkey;
        if (!gk || j.n_frames > slot_gkey_frames) {
            cudaError_t e = cudaMallocAsync((void**)&gk, n * sizeof(uint32_t), s);
            if (e != cudaSuccess) return (int)e;
            async_scratch = true;
        }
        a.gkey = gk;
        sad_fill_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 8192), 256, 0, s>>>(gk, n, 0xFFFFFFFFu);
    }
    cudaError_t e = fe->fn(p, s);
    if (e == cudaSuccess && a.NC > 1) {
        const int vec4 = (j.w % 4 == 0 && (uintptr_t)j.dOut % 4 == 0 && j.pitchOut % 4 == 0 && j.frameOut % 4 == 0 && (uintptr_t)a.gkey % 16 == 0) ? 1 : 0;
        dim3 g(ceil_div(j.w, vec4 ? 1024 : 256), j.y1 - j.y0, j.n_frames);
        const uint32_t magic = j.D > 1 ? (uint32_t)((1ull << 32) / (unsigned)j.D + 1ull) : 0u;
        sad_finalize_kernel<<<g, 256, 0, s>>>(a.gkey, j.dOut, j.w, j.h, j.y0, j.y1, (int)j.pitchOut, j.frameOut, j.D, magic, vec4);
        e = cudaGetLastError();
    }
    if (async_scratch) { cudaError_t e2 = cudaFreeAsync(a.gkey, s); if (e == cudaSuccess) e = e2; }   // stream-ordered: after the kernels above
    if (e != cudaSuccess) return (int)e;
    c->last_launches.store(p.launches);
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
    s->pitch = (size_t)shim_pitch(w);
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
    cudaError_t e = cudaEventRecord(s->uploaded, s->st);     // every H2D copy of this frame is in the stream by now
    if (e != cudaSuccess) return (int)e;
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

// (Re)allocates the pinned and device buffers of a slot for `cap` frame pairs of max_w x max_h.  The new buffers are
// allocated first and swapped in only when every allocation succeeded: a failed grow leaves the slot as it was.
cudaError_t alloc_slot_buffers(Slot* s, int max_w, int max_h, int cap)
{
    const size_t img = (size_t)round_up(max_w, 256) * (size_t)max_h;
    uint8_t *hL = nullptr, *hOut = nullptr, *dL = nullptr, *dOut = nullptr;
    uint32_t* gkey = nullptr;
    // left and right live back to back (pinned and device) so that a whole frame pair is ONE DMA
    cudaError_t e = cudaHostAlloc((void**)&hL, 2 * img * cap, cudaHostAllocPortable);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&hOut, img * cap, cudaHostAllocPortable);
    if (e == cudaSuccess) e = cudaMalloc((void**)&dL, 2 * img * cap);
    if (e == cudaSuccess) e = cudaMalloc((void**)&dOut, img * cap);
    if (e == cudaSuccess) e = cudaMalloc((void**)&gkey, (size_t)max_w * max_h * sizeof(uint32_t) * cap);
    if (e != cudaSuccess) {
        cudaFreeHost(hL); cudaFreeHost(hOut); cudaFree(dL); cudaFree(dOut); cudaFree(gkey);
        cudaGetLastError();
        return e;
    }
    cudaFreeHost(s->hL); cudaFreeHost(s->hOut); cudaFree(s->dL); cudaFree(s->dOut); cudaFree(s->gkey);
    cudaGetLastError();
    s->hL = hL; s->hR = hL + img; s->hOut = hOut; s->dL = dL; s->dR = dL + img; s->dOut = dOut; s->gkey = gkey;
    s->cap = cap;
    return cudaSuccess;
}

void free_slot(Slot* s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->st) { cudaStreamSynchronize(s->st); cudaStreamDestroy(s->st); }
    if (s->done) cudaEventDestroy(s->done);
    if (s->uploaded) cudaEventDestroy(s->uploaded);
    cudaFreeHost(s->hL); cudaFreeHost(s->hOut); cudaFreeHost(s->hRGBA);
    cudaFree(s->dL); cudaFree(s->dOut); cudaFree(s->gkey); cudaFree(s->dRGBA); cudaFree(s->dPost);
    delete s;
}


// ---------------------------------------------------------------------------------------
// Frame cache of sadgpu_compute_region.  The reference's caller cuts one frame pair into up to 160 InputChunks that all
// point at the same two images (pkg/camera/output.go:172-187, pkg/despair/sad.go:135-165).  The first chunk of a pair
// to arrive becomes the producer: every caller that has arrived by then helps copying the pair into a pinned snapshot
// (blocks of rows handed out by an atomic counter), the producer runs ONE whole-frame GPU pass, and every chunk slices its
// rectangle out of the pinned result.  A chunk is served from an entry only after the rows of ITS images that influence
// its region compared equal to the snapshot, so a caller that reuses an image object for new pixels can never see an old
// frame's result.
// ---------------------------------------------------------------------------------------
constexpr int kCacheEntries = 4;
constexpr int kStageRows = 64;          // rows of one plane per staging block

struct FrameEntry {
    Slot* slot = nullptr;
    const uint8_t *l = nullptr, *r = nullptr;                  // key: addresses are compared, never dereferenced outside the caller's own call
    int ls = 0, rs = 0, w = 0, h = 0, B = 0, D = 0;
    enum State { EMPTY, STAGING, COMPUTING, READY, FAILED };
    std::atomic<int> state{EMPTY};                             // written under cache_mu; waiters poll it before they sleep
    int rc = 0;
    int users = 0;                                             // calls currently attached
    long long served = 0;                                      // pixels handed out
    uint64_t tick = 0;
    bool stale = false;                                        // a caller's pixels differed from the snapshot: no new caller may attach
    int nblocks = 0;
    std::atomic<int> next_block{0}, done_blocks{0};
    // per plane and block of kVerifyRows rows: (clock at the start of the latest comparison with the snapshot) << 2 | differs << 1 | done
    std::atomic<uint64_t>* vstate = nullptr;
    int vblocks = 0;                                           // blocks per plane
    ~FrameEntry() { delete[] vstate; }
};

void stage_blocks(sadgpu_ctx*, FrameEntry* e, const uint8_t* l, const uint8_t* r)
{
    Slot* s = e->slot;
    const int per_plane = ceil_div(e->h, kStageRows);
    for (;;) {
        const int b = e->next_block.fetch_add(1, std::memory_order_relaxed);
        if (b >= e->nblocks) break;
        const int plane = b / per_plane, y0 = (b - plane * per_plane) * kStageRows, y1 = std::min(e->h, y0 + kStageRows);
        const uint8_t* src = (plane ? r : l) + (size_t)y0 * (plane ? e->rs : e->ls);
        uint8_t* dst = (plane ? s->hR : s->hL) + (size_t)y0 * s->pitch;
        const size_t sp = (size_t)(plane ? e->rs : e->ls);
        if (sp == s->pitch && s->pitch == (size_t)e->w) memcpy(dst, src, (size_t)(y1 - y0) * e->w);
        else for (int y = y0; y < y1; ++y) memcpy(dst + (size_t)(y - y0) * s->pitch, src + (size_t)(y - y0) * sp, (size_t)e->w);
        e->done_blocks.fetch_add(1, std::memory_order_release);
    }
}

// Producer: snapshot complete -> one DMA up, one whole-frame pass, one DMA down, synchronise.
int produce_frame(sadgpu_ctx* c, FrameEntry* e)
{
    Slot* s = e->slot;
    cudaError_t err = cudaSetDevice(s->device);
    if (err != cudaSuccess) return (int)err;
    const size_t img = s->pitch * (size_t)e->h;
    err = cudaMemcpyAsync(s->dL, s->hL, 2 * img, cudaMemcpyHostToDevice, s->st);            // left and right planes are adjacent
    if (err != cudaSuccess) return (int)err;
    Job j{s->dL, s->pitch, 0, s->dR, s->pitch, 0, s->dOut, s->pitch, 0, 1, e->w, e->h, e->B, e->D, 0, e->h};
    int rc = run_job(c, s->dev_index, j, nullptr, s->gkey, s->st);
    if (rc) { cudaStreamSynchronize(s->st); return rc; }
    err = cudaMemcpyAsync(s->hOut, s->dOut, img, cudaMemcpyDeviceToHost, s->st);
    if (err == cudaSuccess) err = cudaStreamSynchronize(s->st);
    return err == cudaSuccess ? SADGPU_OK : (int)err;
}

// Do the caller's rows [ya, yb) still equal the snapshot the entry was computed from?  Compared in blocks of kVerifyRows rows, and a
// block that another chunk call compared AFTER this call began (`ticket` was drawn from the same clock at its start) is not
// compared again: the images may not change during a call, so what held at any moment of it holds for it.  Neighbouring chunks
// overlap by 2h rows (three times a chunk's own rows for the 8-row bands of a 1080p frame), which made these comparisons 30 % of
// the OutputCamera call pattern.  Blocks are supersets of [ya, yb): a difference outside the rows that matter only costs a recompute.
constexpr int kVerifyRows = 8;
bool rows_match_snapshot(sadgpu_ctx* c, FrameEntry* e, const uint8_t* l, const uint8_t* r, int ya, int yb, uint64_t ticket)
{
#ifdef SADGPU_DEV_NOVERIFY
    return true;                                     // developer timing experiment only: UNSAFE
#endif
    const Slot* s = e->slot;
    const int b0 = ya / kVerifyRows, b1 = (yb - 1) / kVerifyRows;
    for (int plane = 0; plane < 2; ++plane) {
        const uint8_t* src = plane ? r : l;
        const size_t sp = (size_t)(plane ? e->rs : e->ls);
        const uint8_t* snap = plane ? s->hR : s->hL;
        for (int b = b0; b <= b1; ++b) {
            std::atomic<uint64_t>& st = e->vstate[plane * e->vblocks + b];
            for (;;) {
                uint64_t v = st.load(std::memory_order_acquire);
                if ((v >> 2) >= ticket) {                          // compared (or being compared) after this call began
                    if (v & 1) { if (v & 2) return false; break; }
#if defined(__x86_64__) || defined(__i386__)
                    __builtin_ia32_pause();
#endif
                    continue;
                }
                const uint64_t now = c->verify_clock.fetch_add(1, std::memory_order_acq_rel);
                if (!st.compare_exchange_strong(v, now << 2, std::memory_order_acq_rel)) continue;
                const int r0 = b * kVerifyRows, r1 = std::min(e->h, r0 + kVerifyRows);
                bool same = true;
                if (sp == s->pitch && s->pitch == (size_t)e->w) same = memcmp(src + (size_t)r0 * sp, snap + (size_t)r0 * s->pitch, (size_t)(r1 - r0) * e->w) == 0;
                else
                    for (int y = r0; y < r1 && same; ++y) same = memcmp(src + (size_t)y * sp, snap + (size_t)y * s->pitch, (size_t)e->w) == 0;
                st.store((now << 2) | (same ? 1u : 3u), std::memory_order_release);
                if (!same) return false;
                break;
            }
        }
    }
    return true;
}

// Detach from an entry; an idle entry whose every pixel has been handed out (or that is stale / failed) is retired.
void release_entry(sadgpu_ctx* c, FrameEntry* e, long long area)
{
    {
        std::lock_guard<SpinLock> g(c->cache_mu);
        e->served += area;
        if (--e->users == 0 && (e->stale || e->state == FrameEntry::FAILED || e->served >= (long long)e->w * e->h)) {
            e->state = FrameEntry::EMPTY; e->l = e->r = nullptr;
        }
    }
    if (c->cache_sleepers.load(std::memory_order_seq_cst) > 0) { std::lock_guard<std::mutex> g(c->cache_sleep_mu); c->cache_cv.notify_all(); }
}

int new_cache_entry(sadgpu_ctx* c, FrameEntry** out)
{
    FrameEntry* e = new (std::nothrow) FrameEntry();
    Slot* s = new (std::nothrow) Slot();
    if (!e || !s) { delete e; delete s; return SADGPU_ENOMEM; }
    s->dev_index = (int)(c->cache.size() % c->devices.size()); s->device = c->devices[s->dev_index];
    cudaError_t err = cudaSetDevice(s->device);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&s->done, cudaEventDisableTiming);
    if (err == cudaSuccess) err = alloc_slot_buffers(s, c->max_w, c->max_h, 1);
    if (err != cudaSuccess) { free_slot(s); delete e; return (int)err; }
    e->slot = s;
    e->vblocks = ceil_div(c->max_h, kVerifyRows);
    e->vstate = new (std::nothrow) std::atomic<uint64_t>[2 * (size_t)e->vblocks];
    if (!e->vstate) { free_slot(s); delete e; return SADGPU_ENOMEM; }
    for (int i = 0; i < 2 * e->vblocks; ++i) e->vstate[i].store(0, std::memory_order_relaxed);
    *out = e;
    return SADGPU_OK;
}

}  // namespace

extern "C" {

int sadgpu_compute_region(sadgpu_ctx* c, const uint8_t* l, int ls, const uint8_t* r, int rs, int w, int h, int B, int D,
                          int x0, int y0, int x1, int y1, uint8_t* out, int out_stride)
{
    if (!c || !l || !r) return SADGPU_EINVAL;
    if (w <= 0 || h <= 0 || ls < w || rs < w) return SADGPU_EINVAL;
    if (w > c->max_w || h > c->max_h) return SADGPU_ERANGE;
    int rc = validate(w, h, B, D, y0, y1);
    if (rc) return rc;
    if (x0 < 0 || x1 > w || x0 > x1) return SADGPU_ERANGE;
    if (x1 == x0 || y1 == y0) return SADGPU_OK;                  // empty region: nothing to write (sad.go:48-50 allocates 0 bytes)
    if (!out || out_stride < x1 - x0) return SADGPU_EINVAL;
    c->region_calls.fetch_add(1, std::memory_order_relaxed);
    std::call_once(c->cache_once, [c] {                          // the entries' pinned and device buffers: once, outside every lock
        for (int i = 0; i < kCacheEntries && !c->cache_rc; ++i) {
            FrameEntry* fresh = nullptr;
            if ((c->cache_rc = new_cache_entry(c, &fresh)) == SADGPU_OK) c->cache.push_back(fresh);
        }
    });
    if (c->cache.empty()) return c->cache_rc ? c->cache_rc : SADGPU_ENOMEM;
    const int half = B / 2;
    const uint64_t ticket = c->verify_clock.fetch_add(1, std::memory_order_acq_rel);      // this call began here
    for (int attempt = 0;; ++attempt) {
        FrameEntry* e = nullptr;
        bool producer = false;
        {
            std::unique_lock<SpinLock> lk(c->cache_mu);
            if (attempt < 3)                                     // a frame that keeps changing under its caller is computed privately
                for (FrameEntry* q : c->cache)
                    if (q->state != FrameEntry::EMPTY && q->state != FrameEntry::FAILED && !q->stale && q->l == l && q->r == r &&
                        q->ls == ls && q->rs == rs && q->w == w && q->h == h && q->B == B && q->D == D) { e = q; break; }
            if (e) {
                ++e->users;
            } else {
                for (;;) {                                       // an idle entry: empty first, then the least recently used
                    for (FrameEntry* q : c->cache)
                        if (q->users == 0 && q->state == FrameEntry::EMPTY) { e = q; break; }
                    if (!e)
                        for (FrameEntry* q : c->cache)
                            if (q->users == 0 && (!e || q->tick < e->tick)) e = q;
                    if (e) break;
                    lk.unlock();                                 // every entry is attached to a frame in flight: doze until one is released
                    {
                        std::unique_lock<std::mutex> sl(c->cache_sleep_mu);
                        c->cache_sleepers.fetch_add(1, std::memory_order_seq_cst);
                        c->cache_cv.wait_for(sl, std::chrono::microseconds(100));
                        c->cache_sleepers.fetch_sub(1, std::memory_order_seq_cst);
                    }
                    lk.lock();
                }
                producer = true;
                Slot* s = e->slot;
                e->l = l; e->r = r; e->ls = ls; e->rs = rs; e->w = w; e->h = h; e->B = B; e->D = D;
                e->state = FrameEntry::STAGING; e->rc = 0; e->users = 1; e->served = 0; e->stale = attempt >= 3;
                for (int i = 0; i < 2 * e->vblocks; ++i) e->vstate[i].store(0, std::memory_order_relaxed);      // no comparison with this snapshot yet
                s->pitch = (size_t)shim_pitch(w);
                s->dR = s->dL + s->pitch * (size_t)h; s->hR = s->hL + s->pitch * (size_t)h;
                e->nblocks = 2 * ceil_div(h, kStageRows);
                e->done_blocks.store(0); e->next_block.store(0, std::memory_order_release);
                c->region_frames.fetch_add(1, std::memory_order_relaxed);
            }
            e->tick = ++c->cache_tick;
        }
        int state;
        const int ya = std::max(0, y0 - half), yb = std::min(h, y1 + half);
        if (producer) {
            stage_blocks(c, e, l, r);
            while (e->done_blocks.load(std::memory_order_acquire) < e->nblocks) std::this_thread::yield();      // helpers finishing their last block
            e->state.store(FrameEntry::COMPUTING, std::memory_order_release);            // the snapshot is complete: waiters may compare
            rc = produce_frame(c, e);
            {
                std::lock_guard<SpinLock> g(c->cache_mu);
                e->rc = rc; e->state = rc ? FrameEntry::FAILED : FrameEntry::READY; state = e->state;
            }
            if (c->cache_sleepers.load(std::memory_order_seq_cst) > 0) { std::lock_guard<std::mutex> g(c->cache_sleep_mu); c->cache_cv.notify_all(); }
        } else {
            // the snapshot is still being taken: copy blocks of it from this caller's (identical) images
            if (e->state.load(std::memory_order_acquire) == FrameEntry::STAGING) stage_blocks(c, e, l, r);
            // A whole-frame pass takes 0.1 .. 3 ms: poll for a while (waking sleeping workers costs more than the pass of a small
            // frame), then sleep.  Two stops: snapshot complete (the rows this chunk depends on are compared with it while the
            // GPU works), then result ready.
            auto wait_for = [&](auto reached) {
                if (reached()) return;
                static const int spin_limit = std::max(1, (int)std::thread::hardware_concurrency() / 2 - 1);   // never more pollers than half the cores
                if (c->cache_spinners.fetch_add(1, std::memory_order_relaxed) < spin_limit) {
                    const auto t0 = std::chrono::steady_clock::now();
                    while (!reached() && std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(400)) {
                        for (int i = 0; i < 64 && !reached(); ++i) {
#if defined(__x86_64__) || defined(__i386__)
                            __builtin_ia32_pause();
#endif
                        }
                    }
                }
                c->cache_spinners.fetch_sub(1, std::memory_order_relaxed);
                if (reached()) return;
                std::unique_lock<std::mutex> lk(c->cache_sleep_mu);
                // the producer leaves STAGING without a wake-up call: sleepers re-check every 50 us until the result is there
                c->cache_sleepers.fetch_add(1, std::memory_order_seq_cst);
                while (!reached()) c->cache_cv.wait_for(lk, std::chrono::microseconds(50));
                c->cache_sleepers.fetch_sub(1, std::memory_order_seq_cst);
            };
            wait_for([&] { return e->state.load(std::memory_order_acquire) != FrameEntry::STAGING; });
            bool same = true;
            if (e->state.load(std::memory_order_acquire) != FrameEntry::FAILED) same = rows_match_snapshot(c, e, l, r, ya, yb, ticket);
            if (same) wait_for([&] { const int st = e->state.load(std::memory_order_acquire); return st == FrameEntry::READY || st == FrameEntry::FAILED; });
            { std::lock_guard<SpinLock> g(c->cache_mu); state = e->state; rc = e->rc; if (!same) e->stale = true; }
            if (!same) {
                release_entry(c, e, 0);
                c->region_stale.fetch_add(1, std::memory_order_relaxed);
                continue;                                                              // recompute from the caller's current pixels
            }
        }
        if (state == FrameEntry::FAILED) { release_entry(c, e, 0); return rc; }       // every chunk of the frame reports the error
        const Slot* s = e->slot;
        for (int y = y0; y < y1; ++y)                                                  // region-local rows: OutputChunk.DisparityData (sad.go:91)
            memcpy(out + (size_t)(y - y0) * out_stride, s->hOut + (size_t)y * s->pitch + x0, (size_t)(x1 - x0));
        release_entry(c, e, (long long)(x1 - x0) * (y1 - y0));
        return SADGPU_OK;
    }
}

int sadgpu_region_stats(sadgpu_ctx* c, long long* calls, long long* frames, long long* stale)
{
    if (!c) return SADGPU_EINVAL;
    if (calls) *calls = c->region_calls.load();
    if (frames) *frames = c->region_frames.load();
    if (stale) *stale = c->region_stale.load();
    return SADGPU_OK;
}

int sadgpu_wait_uploaded(sadgpu_ctx* c, uint64_t ticket)
{
    if (!c) return SADGPU_EINVAL;
    const int stream = (int)(ticket & 0xFFFF);
    if (stream >= (int)c->slots.size()) return SADGPU_EBUSY;
    Slot* s = c->slots[stream];
    std::lock_guard<std::mutex> g(s->mu);
    if (!s->busy || s->seq != (ticket >> 16)) return SADGPU_EBUSY;
    cudaError_t e = cudaSetDevice(s->device);
    if (e == cudaSuccess) e = cudaEventSynchronize(s->uploaded);
    return e == cudaSuccess ? SADGPU_OK : (int)e;
}


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
    c->max_w = max_w; c->max_h = max_h;
    // devices == NULL: 0..n_devices-1, or the first n_devices entries of the comma-separated list in SADGPU_DEVICES
    std::vector<int> env_list;
    if (!devices) {
        if (const char* env = getenv("SADGPU_DEVICES")) {
            for (const char* q = env; *q;) {
                char* end = nullptr;
                const long v = strtol(q, &end, 10);
                if (end == q) break;
                env_list.push_back((int)v);
                q = *end == ',' ? end + 1 : end;
                if (*end && *end != ',') break;
            }
            if ((int)env_list.size() < n_devices) { delete c; return SADGPU_ERANGE; }
        }
    }
    for (int i = 0; i < n_devices; ++i) {
        const int d = devices ? devices[i] : env_list.empty() ? i : env_list[i];
        if (d < 0 || d >= ndev) { delete c; return SADGPU_ERANGE; }
        cudaDeviceProp prop;
        if ((e = cudaGetDeviceProperties(&prop, d)) != cudaSuccess) { delete c; return (int)e; }
        if (prop.major != 10) { delete c; return SADGPU_ENODEV; }     // sm_100a cubin only, no fallback
        c->devices.push_back(d);
        c->sm_count.push_back(prop.multiProcessorCount);
        // key-map scratch of the device-resident entry points comes from the stream-ordered pool: keep freed blocks cached
        cudaMemPool_t mp;
        if (cudaDeviceGetDefaultMemPool(&mp, d) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    c->copier = new (std::nothrow) CopyPool((int)std::min(3u, std::max(1u, std::thread::hardware_concurrency() / 4)));
    if (!c->copier) { delete c; return SADGPU_ENOMEM; }
    c->dev_dbg.assign(n_devices, nullptr);
    const size_t pitch = (size_t)round_up(max_w, 256);
    for (int i = 0; i < n_streams; ++i) {
        Slot* s = new (std::nothrow) Slot();
        if (!s) { sadgpu_destroy(c); return SADGPU_ENOMEM; }
        c->slots.push_back(s);
        s->dev_index = i % n_devices; s->device = c->devices[s->dev_index]; s->pitch = pitch;   // re-set per frame
        e = cudaSetDevice(s->device);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->uploaded, cudaEventDisableTiming);
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
    for (FrameEntry* e : c->cache) { free_slot(e->slot); delete e; }
    for (size_t i = 0; i < c->dev_dbg.size(); ++i)
        if (c->dev_dbg[i]) { cudaSetDevice(c->devices[i]); cudaFree(c->dev_dbg[i]); }
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
    s->pitch = (size_t)shim_pitch(w);
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
    if ((e = cudaEventRecord(s->uploaded, s->st)) != cudaSuccess) return (int)e;
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

// ---- post-processing hooks: additive, never on the bit-exact path ----
int sadgpu_median3_device(sadgpu_ctx* c, int device, const uint8_t* dSrc, size_t src_pitch, int w, int h, uint8_t* dDst, size_t dst_pitch,
                          void* cuda_stream)
{
    if (!c || !dSrc || !dDst || dSrc == dDst || w <= 0 || h <= 0 || src_pitch < (size_t)w || dst_pitch < (size_t)w) return SADGPU_EINVAL;
    if (device < 0 || device >= (int)c->devices.size()) return SADGPU_ERANGE;
    cudaError_t e = cudaSetDevice(c->devices[device]);
    if (e != cudaSuccess) return (int)e;
    median3_kernel<<<dim3(ceil_div(w, 256), h), 256, 0, (cudaStream_t)cuda_stream>>>(dSrc, src_pitch, dDst, dst_pitch, w, h);
    e = cudaGetLastError();
    return e == cudaSuccess ? SADGPU_OK : (int)e;
}

int sadgpu_lrcheck_device(sadgpu_ctx* c, int device, const uint8_t* dLeftMap, size_t left_pitch, const uint8_t* dRightMap, size_t right_pitch,
                          int w, int h, int max_disparity, int tolerance, int invalid_value, uint8_t* dDst, size_t dst_pitch, void* cuda_stream)
{
    if (!c || !dLeftMap || !dRightMap || !dDst || w <= 0 || h <= 0 || left_pitch < (size_t)w || right_pitch < (size_t)w || dst_pitch < (size_t)w)
        return SADGPU_EINVAL;
    if (max_disparity < 1 || max_disparity > SADGPU_MAX_DISPARITY || tolerance < 0 || invalid_value < 0 || invalid_value > 255) return SADGPU_EINVAL;
    if (device < 0 || device >= (int)c->devices.size()) return SADGPU_ERANGE;
    cudaError_t e = cudaSetDevice(c->devices[device]);
    if (e != cudaSuccess) return (int)e;
    lrcheck_kernel<<<dim3(ceil_div(w, 256), h), 256, 0, (cudaStream_t)cuda_stream>>>(dLeftMap, left_pitch, dRightMap, right_pitch, dDst, dst_pitch,
                                                                                   w, h, max_disparity, tolerance, invalid_value);
    e = cudaGetLastError();
    return e == cudaSuccess ? SADGPU_OK : (int)e;
}

int sadgpu_compute_checked(sadgpu_ctx* c, int stream, const uint8_t* l, int ls, const uint8_t* r, int rs, int w, int h, int B, int D,
                           int tolerance, int invalid_value, int median, uint8_t* out, int out_stride)
{
    int rc = check_io(c, stream, l, ls, r, rs, w, h);
    if (rc) return rc;
    if (!out || out_stride < w || tolerance < 0 || invalid_value < 0 || invalid_value > 255) return SADGPU_EINVAL;
    if ((rc = validate(w, h, B, D, 0, h))) return rc;
    Slot* s = c->slots[stream];
    std::lock_guard<std::mutex> g(s->mu);
    if (s->busy) return SADGPU_EBUSY;
    cudaError_t e = cudaSetDevice(s->device);
    if (e != cudaSuccess) return (int)e;
    const size_t pitch = (size_t)shim_pitch(w), plane = pitch * (size_t)h;
    if (s->post_bytes < 5 * plane) {
        cudaStreamSynchronize(s->st);
        cudaFree(s->dPost); s->dPost = nullptr; s->post_bytes = 0;
        if ((e = cudaMalloc((void**)&s->dPost, 5 * plane)) != cudaSuccess) return (int)e;
        s->post_bytes = 5 * plane;
    }
    s->pitch = pitch; s->dR = s->dL + plane; s->hR = s->hL + plane;
    if ((rc = upload(c, s, l, ls, s->hL, s->dL, w, 0, h))) return rc;
    if ((rc = upload(c, s, r, rs, s->hR, s->dR, w, 0, h))) return rc;
    uint8_t *mL = s->dPost, *mR = mL + plane, *mapRm = mR + plane, *mapR = mapRm + plane, *tmp = mapR + plane;
    const dim3 grid(ceil_div(w, 256), h);
    // left-referenced map (the bit-exact path), then the right-referenced one: the same path on the mirrored pair with the roles swapped
    Job jl{s->dL, pitch, 0, s->dR, pitch, 0, s->dOut, pitch, 0, 1, w, h, B, D, 0, h};
    if ((rc = run_job(c, s->dev_index, jl, nullptr, s->gkey, s->st))) { cudaStreamSynchronize(s->st); return rc; }
    mirror_kernel<<<grid, 256, 0, s->st>>>(s->dR, pitch, mL, pitch, w, h);
    mirror_kernel<<<grid, 256, 0, s->st>>>(s->dL, pitch, mR, pitch, w, h);
    Job jr{mL, pitch, 0, mR, pitch, 0, mapRm, pitch, 0, 1, w, h, B, D, 0, h};
    if ((rc = run_job(c, s->dev_index, jr, nullptr, s->gkey, s->st))) { cudaStreamSynchronize(s->st); return rc; }
    mirror_kernel<<<grid, 256, 0, s->st>>>(mapRm, pitch, mapR, pitch, w, h);
    lrcheck_kernel<<<grid, 256, 0, s->st>>>(s->dOut, pitch, mapR, pitch, tmp, pitch, w, h, D, tolerance, invalid_value);
    const uint8_t* result = tmp;
    if (median) { median3_kernel<<<grid, 256, 0, s->st>>>(tmp, pitch, mapRm, pitch, w, h); result = mapRm; }
    if ((e = cudaGetLastError()) != cudaSuccess) { cudaStreamSynchronize(s->st); return (int)e; }
    const bool direct = in_pool(c, out, (size_t)(h - 1) * out_stride + w);
    uint8_t* dst = direct ? out : s->hOut;
    const size_t dpitch = direct ? (size_t)out_stride : pitch;
    e = cudaMemcpy2DAsync(dst, dpitch, result, pitch, (size_t)w, (size_t)h, cudaMemcpyDeviceToHost, s->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->st);
    if (e != cudaSuccess) return (int)e;
    if (!direct) c->copier->copy(out, (size_t)out_stride, s->hOut, pitch, (size_t)w, (size_t)h);
    return SADGPU_OK;
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
    if (!c->dev_dbg[device] || n_words > 1024) return SADGPU_ERANGE;
    cudaError_t e = cudaSetDevice(c->devices[device]);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(host, c->dev_dbg[device], (size_t)n_words * 4, cudaMemcpyDeviceToHost);
    return e == cudaSuccess ? SADGPU_OK : (int)e;
}

int sadgpu_last_launch_count(sadgpu_ctx* c) { return c ? c->last_launches.load() : 0; }

int sadgpu_plan_describe(int w, int h, int B, int D, int y0, int y1, const sadgpu_tuning* t, char* buf, size_t buflen)
{
    FastPlan p;
    const int nf = t && t->reserved[0] > 0 ? t->reserved[0] : 1;       // reserved[0]: frames per launch (describe only)
    int rc = make_fast_plan(w, h, B, D, y0, y1, nf, t, 148, &p);
    if (rc) return rc;
    if (buf && buflen)
        snprintf(buf, buflen,
                 "{\"variant\":\"%s\",\"mode\":%d,\"half\":%d,\"NG\":%d,\"NC\":%d,\"NGc\":%d,\"TW\":%d,\"RB\":%d,\"BH\":%d,"
                 "\"grid\":[%u,%u,%u],\"threads\":%d,\"smem\":%zu,\"launches\":%d,\"frames_per_launch\":%d}",
                 variant_name(p.variant), p.mode, p.half, p.a.NG, p.a.NC, p.ngc, p.tw, p.rb, p.a.BH,
                 p.grid.x, p.grid.y, p.grid.z, p.nthreads, p.smem, p.launches, nf);
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
