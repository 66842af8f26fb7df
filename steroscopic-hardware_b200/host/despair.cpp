// despair.cpp — see despair.hpp.  Host threads + the sadgpu C ABI; no disparity arithmetic lives here.
#include "despair.hpp"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>

#include "../../include/sadgpu.h"

namespace despair {

namespace {
std::mutex g_params_mu;                                   // params.go:8-11
Parameters g_params{16, 64};                              // params.go:13-18

std::mutex g_backend_mu;
sadgpu_ctx* g_ctx = nullptr;
int g_max_w = 4096, g_max_h = 2304, g_streams = 32;

sadgpu_ctx* backend()
{
    std::lock_guard<std::mutex> l(g_backend_mu);
    if (!g_ctx) {
        int dev = 0;
        int rc = sadgpu_create(&dev, 1, g_max_w, g_max_h, g_streams, &g_ctx);
        if (rc != SADGPU_OK) throw std::runtime_error(std::string("sadgpu_create: ") + sadgpu_strerror(rc));
    }
    return g_ctx;
}

void check_pair(const Gray* l, const Gray* r)
{
    if (!l || !r) throw std::runtime_error("nil image");
    if (l->Rect_.MinX || l->Rect_.MinY || r->Rect_.MinX || r->Rect_.MinY)
        throw std::runtime_error("Rect.Min must be (0,0)");
    if (l->Rect_.Dx() != r->Rect_.Dx() || l->Rect_.Dy() != r->Rect_.Dy())
        throw std::runtime_error("left.Rect != right.Rect");
}

// One chunk on stream slot `slot`: rows of the region through the C ABI, then the region's columns.
OutputChunk process_chunk(const InputChunk& c, Parameters p, int slot)
{
    check_pair(c.Left, c.Right);
    const int w = c.Left->Rect_.Dx(), h = c.Left->Rect_.Dy();
    const Rectangle& g = c.Region;
    OutputChunk o;
    o.Region = g;
    o.DisparityData.assign((size_t)std::max(0, g.Dx()) * std::max(0, g.Dy()), 0);   // sad.go:48-50
    if (g.Dx() <= 0 || g.Dy() <= 0) return o;
    std::vector<uint8_t> rows((size_t)w * g.Dy());
    // full-map addressing: row y lands at base + y*w, so shift the base by MinY rows
    int rc = sadgpu_compute(backend(), slot % g_streams, c.Left->Pix.data(), c.Left->Stride, c.Right->Pix.data(),
                            c.Right->Stride, w, h, p.BlockSize, p.MaxDisparity, g.MinY, g.MaxY,
                            rows.data() - (ptrdiff_t)g.MinY * w, w);
    if (rc != SADGPU_OK) throw std::runtime_error(std::string("sadgpu_compute: ") + sadgpu_strerror(rc));
    for (int y = 0; y < g.Dy(); ++y)
        memcpy(&o.DisparityData[(size_t)y * g.Dx()], &rows[(size_t)y * w + g.MinX], (size_t)g.Dx());
    return o;
}
}  // namespace

Gray NewGray(Rectangle r)
{
    Gray g;
    g.Rect_ = r;
    g.Stride = r.Dx();
    g.Pix.assign((size_t)std::max(0, r.Dx()) * std::max(0, r.Dy()), 0);
    return g;
}

void SetDefaultParams(Parameters p) { std::lock_guard<std::mutex> l(g_params_mu); g_params = p; }
Parameters DefaultParams() { std::lock_guard<std::mutex> l(g_params_mu); return g_params; }

void ConfigureBackend(int max_w, int max_h, int n_streams)
{
    std::lock_guard<std::mutex> l(g_backend_mu);
    if (g_ctx) { sadgpu_destroy(g_ctx); g_ctx = nullptr; }
    g_max_w = max_w; g_max_h = max_h; g_streams = std::max(1, n_streams);
}

void ShutdownBackend()
{
    std::lock_guard<std::mutex> l(g_backend_mu);
    if (g_ctx) { sadgpu_destroy(g_ctx); g_ctx = nullptr; }
}

Pipeline SetupConcurrentSAD(int numWorkers)
{
    if (numWorkers <= 0) numWorkers = (int)std::max(1u, std::thread::hardware_concurrency()) * 4;   // sad.go:32-34
    numWorkers = std::min(numWorkers, g_streams);          // one CUDA stream slot per worker
    Pipeline p{std::make_shared<Chan<InputChunk>>((size_t)numWorkers * 2),
               std::make_shared<Chan<OutputChunk>>((size_t)numWorkers * 2)};                          // :36-37
    auto live = std::make_shared<std::atomic<int>>(numWorkers);
    for (int w = 0; w < numWorkers; ++w) {
        std::thread([in = p.In, out = p.Out, live, w] {
            InputChunk chunk;
            while (in->Recv(chunk)) {                      // for chunk := range inputChan  (:47)
                Parameters params = DefaultParams();       // snapshot per chunk (:51-53)
                try {
                    out->Send(process_chunk(chunk, params, w));                                       // :98-101
                } catch (const std::exception&) {
                    OutputChunk dead; dead.Region = chunk.Region;                                    // error: empty chunk, caller's retry loop applies
                    try { out->Send(std::move(dead)); } catch (...) {}
                }
            }
            if (live->fetch_sub(1) == 1) out->Close();     // wg.Wait(); close(outputChan)  (:107-110)
        }).detach();
    }
    return p;
}

std::vector<Rectangle> RunSadChunks(Rectangle dims, int numCPU)
{
    const int numWorkers = numCPU * 4, numChunks = numWorkers * 4;                                    // :128-129
    const int W = dims.Dx(), H = dims.Dy();
    if ((W * H) / numChunks <= 0) throw std::runtime_error("integer divide by zero (sad.go:143)");
    int chunkWidth = (int)std::sqrt((double)((W * H) / numChunks));                                   // :138-142
    const int horChunks = std::max(1, W / chunkWidth);                                                // :143
    const int verChunks = std::max(1, numChunks / horChunks);                                         // :144
    chunkWidth = W / horChunks;                                                                       // :145
    const int chunkHeight = H / verChunks;                                                            // :146
    if (chunkHeight <= 0) throw std::runtime_error("chunkHeight == 0: the reference loops forever (sad.go:147)");
    std::vector<Rectangle> chunks;
    for (int y = dims.MinY; y < dims.MaxY; y += chunkHeight)                                          // :147-153
        for (int x = dims.MinX; x < dims.MaxX; x += chunkWidth)
            chunks.push_back(Rect(x, y, std::min(x + chunkWidth, dims.MaxX), std::min(y + chunkHeight, dims.MaxY)));
    return chunks;
}

Gray AssembleDisparityMap(Chan<OutputChunk>& outputChan, Rectangle dimensions, int chunks, bool faithful_drop)
{
    Gray map = NewGray(dimensions);                                                                   // :177
    int i = 0;
    OutputChunk c;
    while (i < chunks && outputChan.Recv(c)) {
        ++i;
        if (faithful_drop && i >= chunks) break;           // sad.go:179-184: the reference drops the chunk that arrives last
        const int width = c.Region.Dx();
        if ((int)c.DisparityData.size() != width * c.Region.Dy()) continue;     // failed chunk: leave zeros
        for (int y = 0; y < c.Region.Dy(); ++y)                                                       // :186-197
            memcpy(&map.Pix[(size_t)(c.Region.MinY + y - dimensions.MinY) * map.Stride + (c.Region.MinX - dimensions.MinX)],
                   &c.DisparityData[(size_t)y * width], (size_t)width);
    }
    return map;
}

Gray RunSad(const Gray& left, const Gray& right, int blockSize, int maxDisparity)
{
    if (maxDisparity == 0) throw std::runtime_error("integer divide by zero (sad.go:92)");
    SetDefaultParams(Parameters{blockSize, maxDisparity});                                            // :123-126 (global side effect kept)
    const int numCPU = (int)std::max(1u, std::thread::hardware_concurrency());
    Pipeline p = SetupConcurrentSAD(numCPU * 4);                                                      // :128-132
    std::vector<Rectangle> chunks = RunSadChunks(left.Rect_, numCPU);                                 // :135-153
    std::thread feeder([&] {                                                                          // :156-165
        for (const Rectangle& r : chunks) p.In->Send(InputChunk{&left, &right, r});
        p.In->Close();
    });
    Gray out = AssembleDisparityMap(*p.Out, left.Rect_, (int)chunks.size());                          // :168
    feeder.join();
    return out;
}

}  // namespace despair

// ---- C hooks so that the Python test-suite can drive the C++ mirror (tests only) -------------------
namespace despair {

static PinnedFrames new_pinned(int n, int planes, int w, int h)
{
    PinnedFrames f;
    f.base = static_cast<uint8_t*>(sadgpu_host_alloc(backend(), (size_t)n * planes * w * h));
    if (!f.base) throw std::runtime_error("sadgpu_host_alloc failed");
    f.n = n; f.planes = planes; f.w = w; f.h = h;
    return f;
}
PinnedFrames NewPinnedPairs(int n, int w, int h) { return new_pinned(n, 2, w, h); }
PinnedFrames NewPinnedMaps(int n, int w, int h) { return new_pinned(n, 1, w, h); }
void FreePinned(PinnedFrames& f) { if (f.base) sadgpu_host_free(backend(), f.base); f = PinnedFrames{}; }

void StreamSad(const PinnedFrames& pairs, PinnedFrames& maps, int batch, int depth)
{
    if (pairs.planes != 2 || maps.planes != 1 || pairs.n != maps.n || pairs.w != maps.w || pairs.h != maps.h || batch < 1)
        throw std::runtime_error("StreamSad: mismatching frame sets");
    sadgpu_ctx* g = backend();
    const Parameters p = DefaultParams();                       // one snapshot per call (documented deviation: per chunk in Go)
    depth = std::max(1, std::min(depth, g_streams));
    int rc = SADGPU_OK;                                         // the `depth` streams used here grow their buffers on their first batch
    std::vector<uint64_t> ticket(depth, 0);
    std::vector<char> busy(depth, 0);
    int call = 0, first_err = SADGPU_OK;
    for (int i = 0; i < pairs.n; i += batch, ++call) {
        const int s = call % depth, n = std::min(batch, pairs.n - i);
        if (busy[s]) { rc = sadgpu_wait(g, ticket[s], nullptr, 0); busy[s] = 0; if (rc && !first_err) first_err = rc; }
        rc = sadgpu_submit_batch_into(g, s, n, pairs.Left(i), pairs.w, pairs.h, p.BlockSize, p.MaxDisparity, maps.Map(i), &ticket[s]);
        if (rc) { if (!first_err) first_err = rc; break; }
        busy[s] = 1;
    }
    for (int s = 0; s < depth; ++s)
        if (busy[s]) { rc = sadgpu_wait(g, ticket[s], nullptr, 0); if (rc && !first_err) first_err = rc; }
    if (first_err) throw std::runtime_error(std::string("StreamSad: ") + sadgpu_strerror(first_err));
}

}  // namespace despair

extern "C" {

// Video path: n frame pairs through StreamSad (pinned pairs in, pinned maps out), results copied back for the test.
int despair_host_stream(const uint8_t* left, const uint8_t* right, int n, int w, int h, int block_size, int max_disparity,
                        int batch, uint8_t* out)
{
    try {
        despair::SetDefaultParams(despair::Parameters{block_size, max_disparity});
        despair::PinnedFrames pairs = despair::NewPinnedPairs(n, w, h), maps = despair::NewPinnedMaps(n, w, h);
        for (int i = 0; i < n; ++i) {
            memcpy(pairs.Left(i), left + (size_t)i * w * h, (size_t)w * h);
            memcpy(pairs.Right(i), right + (size_t)i * w * h, (size_t)w * h);
        }
        despair::StreamSad(pairs, maps, batch);
        memcpy(out, maps.base, (size_t)n * w * h);
        despair::FreePinned(pairs); despair::FreePinned(maps);
        return 0;
    } catch (const std::exception&) { return -1; }
}

int despair_host_run_sad(const uint8_t* left, const uint8_t* right, int w, int h, int block_size, int max_disparity,
                         uint8_t* out)
{
    try {
        despair::Gray l = despair::NewGray(despair::Rect(0, 0, w, h)), r = l;
        memcpy(l.Pix.data(), left, (size_t)w * h);
        memcpy(r.Pix.data(), right, (size_t)w * h);
        despair::Gray o = despair::RunSad(l, r, block_size, max_disparity);
        memcpy(out, o.Pix.data(), (size_t)w * h);
        return 0;
    } catch (const std::exception&) { return -1; }
}

// OutputCamera-style use: SetupConcurrentSAD(workers), row bands of chunk_rows rows (output.go:172-187).
int despair_host_pipeline(const uint8_t* left, const uint8_t* right, int w, int h, int block_size, int max_disparity,
                          int workers, int chunk_rows, int faithful_drop, uint8_t* out, int* params_seen)
{
    try {
        despair::SetDefaultParams(despair::Parameters{block_size, max_disparity});
        despair::Gray l = despair::NewGray(despair::Rect(0, 0, w, h)), r = l;
        memcpy(l.Pix.data(), left, (size_t)w * h);
        memcpy(r.Pix.data(), right, (size_t)w * h);
        despair::Pipeline p = despair::SetupConcurrentSAD(workers);
        const int n = (h + chunk_rows - 1) / chunk_rows;
        std::thread feeder([&] {
            for (int y = 0; y < h; y += chunk_rows)
                p.In->Send(despair::InputChunk{&l, &r, despair::Rect(0, y, w, std::min(y + chunk_rows, h))});
        });
        despair::Gray o = despair::AssembleDisparityMap(*p.Out, l.Rect_, n, faithful_drop != 0);
        feeder.join();
        p.In->Close();
        memcpy(out, o.Pix.data(), (size_t)w * h);
        if (params_seen) { params_seen[0] = despair::DefaultParams().BlockSize; params_seen[1] = despair::DefaultParams().MaxDisparity; }
        return 0;
    } catch (const std::exception&) { return -1; }
}

int despair_host_run_sad_chunks(int w, int h, int num_cpu, int* rects, int max_rects)
{
    try {
        std::vector<despair::Rectangle> c = despair::RunSadChunks(despair::Rect(0, 0, w, h), num_cpu);
        const int n = (int)std::min<size_t>(c.size(), (size_t)max_rects);
        for (int i = 0; i < n; ++i) { rects[4 * i] = c[i].MinX; rects[4 * i + 1] = c[i].MinY; rects[4 * i + 2] = c[i].MaxX; rects[4 * i + 3] = c[i].MaxY; }
        return (int)c.size();
    } catch (const std::exception&) { return -1; }
}

void despair_host_shutdown(void) { despair::ShutdownBackend(); }

}  // extern "C"
