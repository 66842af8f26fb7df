// despair.cpp — see despair.hpp.  Host threads + the sadgpu C ABI; no disparity arithmetic lives here.
#include "despair.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <deque>
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
int g_max_w = 4096, g_max_h = 2304, g_streams = 4;
std::atomic<unsigned> g_next_slot{0};

sadgpu_ctx* backend()
{
    std::lock_guard<std::mutex> l(g_backend_mu);
    if (!g_ctx) {
        // devices: SADGPU_DEVICES (comma-separated ordinals) or device 0
        int rc = sadgpu_create(nullptr, 1, g_max_w, g_max_h, g_streams, &g_ctx);
        if (rc != SADGPU_OK) throw std::runtime_error(std::string("sadgpu_create: ") + sadgpu_strerror(rc));
    }
    return g_ctx;
}

[[noreturn]] void fail(const char* what, int rc) { throw std::runtime_error(std::string(what) + ": " + sadgpu_strerror(rc)); }

template <class Img> void check_pair(const Img* l, const Img* r)
{
    if (!l || !r) throw std::runtime_error("nil image");
    if (l->Rect_.MinX || l->Rect_.MinY || r->Rect_.MinX || r->Rect_.MinY)
        throw std::runtime_error("Rect.Min must be (0,0)");
    if (l->Rect_.Dx() != r->Rect_.Dx() || l->Rect_.Dy() != r->Rect_.Dy())
        throw std::runtime_error("left.Rect != right.Rect");
}

// One chunk = one call: the region-local result lands directly in OutputChunk.DisparityData (sad.go:48-50, :91); the chunks of
// a frame pair share one GPU pass inside the library (sadgpu_compute_region).
OutputChunk process_chunk(const InputChunk& c, Parameters p)
{
    check_pair(c.Left, c.Right);
    const int w = c.Left->Rect_.Dx(), h = c.Left->Rect_.Dy();
    const Rectangle& g = c.Region;
    OutputChunk o;
    o.Region = g;
    o.DisparityData.assign((size_t)std::max(0, g.Dx()) * std::max(0, g.Dy()), 0);   // sad.go:48-50
    if (g.Dx() <= 0 || g.Dy() <= 0) return o;
    int rc = sadgpu_compute_region(backend(), c.Left->Pix.data(), c.Left->Stride, c.Right->Pix.data(), c.Right->Stride, w, h,
                                   p.BlockSize, p.MaxDisparity, g.MinX, g.MinY, g.MaxX, g.MaxY, o.DisparityData.data(), g.Dx());
    if (rc != SADGPU_OK) fail("sadgpu_compute_region", rc);
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

NRGBA NewNRGBA(Rectangle r)
{
    NRGBA g;
    g.Rect_ = r;
    g.Stride = 4 * r.Dx();
    g.Pix.assign((size_t)4 * std::max(0, r.Dx()) * std::max(0, r.Dy()), 0);
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
    const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
    if (numWorkers <= 0) numWorkers = hw * 4;                                                        // sad.go:32-34
    Pipeline p{std::make_shared<Chan<InputChunk>>((size_t)numWorkers * 2),
               std::make_shared<Chan<OutputChunk>>((size_t)numWorkers * 2)};                          // :36-37
    const int threads = std::max(1, std::min(numWorkers, hw / 2));           // OS threads that carry the logical workers
    const int quota = (numWorkers + threads - 1) / threads;                  // chunks one thread may hold: threads * quota >= numWorkers
    auto live = std::make_shared<std::atomic<int>>(threads);
    for (int t = 0; t < threads; ++t) {
        std::thread([in = p.In, out = p.Out, live, quota] {
            std::deque<OutputChunk> held;                  // finished chunks the output channel had no room for yet
            auto work = [&](const InputChunk& chunk) {     // the worker body, sad.go:47-102
                Parameters params = DefaultParams();       // snapshot per chunk (:51-53)
                OutputChunk result;
                try {
                    result = process_chunk(chunk, params);
                } catch (const std::exception& e) {        // one OutputChunk per InputChunk all the same (:98-101); the error travels beside it
                    out->Fail(e.what());
                    result = OutputChunk{};
                    result.Region = chunk.Region;
                }
                held.push_back(std::move(result));
            };
            try {
                for (bool open = true; open || !held.empty();) {
                    bool progress = false;
                    while (!held.empty() && out->TrySend(held.front())) { held.pop_front(); progress = true; }
                    if (open && (int)held.size() < quota) {
                        InputChunk chunk;
                        // nothing held: sleep in the channel like `for chunk := range inputChan` (:47); otherwise only look
                        const int got = held.empty() ? (in->Recv(chunk) ? 1 : -1) : in->TryRecv(chunk);
                        if (got > 0) { work(chunk); progress = true; }
                        else if (got < 0) open = false;
                    }
                    if (!progress && !held.empty()) {
                        if (!open || (int)held.size() >= quota) { out->Send(std::move(held.front())); held.pop_front(); continue; }   // nothing else to do
                        // room in the quota: wait for a place in the output OR a chunk in the input — poll both for a while, then doze
                        const auto t0 = std::chrono::steady_clock::now();
                        bool ready = false;
                        while (!ready && std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(100)) {
                            for (int i = 0; i < 64 && !ready; ++i) {
                                ready = out->Size() < out->Cap() || in->Size() > 0;
#if defined(__x86_64__) || defined(__i386__)
                                if (!ready) __builtin_ia32_pause();
#endif
                            }
                        }
                        if (!ready) out->WaitNotFullFor(std::chrono::microseconds(50));
                    }
                }
            } catch (...) {}                               // send on a closed output channel: the pipeline is being torn down
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
        if ((int)c.DisparityData.size() != width * c.Region.Dy()) continue;     // failed chunk: reported below
        for (int y = 0; y < c.Region.Dy(); ++y)                                                       // :186-197
            memcpy(&map.Pix[(size_t)(c.Region.MinY + y - dimensions.MinY) * map.Stride + (c.Region.MinX - dimensions.MinX)],
                   &c.DisparityData[(size_t)y * width], (size_t)width);
    }
    const std::string err = outputChan.TakeError();
    if (!err.empty()) throw std::runtime_error("AssembleDisparityMap: a chunk failed: " + err);
    return map;
}

Gray RunSad(const Gray& left, const Gray& right, int blockSize, int maxDisparity)
{
    if (maxDisparity == 0) throw std::runtime_error("integer divide by zero (sad.go:92)");
    check_pair(&left, &right);
    // the backend's parameter surface (include/sadgpu.h) is checked before any chunk is sent
    if (blockSize < 1 || blockSize > SADGPU_MAX_BLOCK_SIZE || maxDisparity < 1 || maxDisparity > SADGPU_MAX_DISPARITY)
        throw std::runtime_error("RunSad: blockSize must be 1..31 and maxDisparity 1..256 on the CUDA backend");
    if (left.Rect_.Dx() > g_max_w || left.Rect_.Dy() > g_max_h)
        throw std::runtime_error("RunSad: image larger than the backend (ConfigureBackend)");
    SetDefaultParams(Parameters{blockSize, maxDisparity});                                            // :123-126 (global side effect kept)
    const int numCPU = (int)std::max(1u, std::thread::hardware_concurrency());
    Pipeline p = SetupConcurrentSAD(numCPU * 4);                                                      // :128-132
    std::vector<Rectangle> chunks = RunSadChunks(left.Rect_, numCPU);                                 // :135-153
    std::thread feeder([&] {                                                                          // :156-165
        for (const Rectangle& r : chunks) p.In->Send(InputChunk{&left, &right, r});
        p.In->Close();
    });
    Gray out;
    try {
        out = AssembleDisparityMap(*p.Out, left.Rect_, (int)chunks.size());                           // :168
    } catch (...) { feeder.join(); throw; }
    feeder.join();
    return out;
}

Gray ProcessDepthMap(const NRGBA& left, const NRGBA& right)
{
    check_pair(&left, &right);
    const Parameters p = DefaultParams();                                                             // output.go:169
    Gray map = NewGray(left.Rect_);
    const int w = left.Rect_.Dx(), h = left.Rect_.Dy();
    if (w <= 0 || h <= 0) return map;
    const int slot = (int)(g_next_slot.fetch_add(1) % (unsigned)g_streams);
    int rc = sadgpu_compute_nrgba(backend(), slot, left.Pix.data(), left.Stride, right.Pix.data(), right.Stride, w, h,
                                  p.BlockSize, p.MaxDisparity, map.Pix.data(), map.Stride);
    if (rc != SADGPU_OK) fail("sadgpu_compute_nrgba", rc);
    return map;
}

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
    if (first_err) fail("StreamSad", first_err);
}

bool ReadFrame(Port& port, uint8_t* pix, int w, int h)
{
    const size_t expected = (size_t)w * h;                      // serial.go:248
    size_t have = 0;
    while (have < expected) {                                   // :274-295, reads of at most 1024 bytes, straight into the frame
        const int n = port.Read(pix + have, (int)std::min<size_t>(1024, expected - have));
        if (n <= 0) return false;                               // error reading from serial port (:277-281)
        have += (size_t)n;
    }
    return true;
}

int StreamSerialPairs(Port& left, Port& right, int w, int h, int max_frames,
                      const std::function<void(int, const uint8_t*)>& sink)
{
    sadgpu_ctx* g = backend();
    if (g_streams < 2) throw std::runtime_error("StreamSerialPairs needs two stream slots (ConfigureBackend)");
    PinnedFrames pairs = NewPinnedPairs(2, w, h), maps = NewPinnedMaps(2, w, h);
    uint64_t ticket[2] = {0, 0};
    bool busy[2] = {false, false};
    int submitted = 0, delivered = 0, err = SADGPU_OK;
    auto retire = [&](int b) {                                  // frame in slot b is done: hand its map over, the pair is free again
        int rc = sadgpu_wait(g, ticket[b], nullptr, 0);
        busy[b] = false;
        if (rc) { if (!err) err = rc; return; }
        sink(delivered++, maps.Map(b));
    };
    for (int k = 0; k < max_frames && !err; ++k) {
        const int b = k & 1;
        if (busy[b]) retire(b);                                 // frame k-2; frame k-1 stays on the GPU while frame k is read
        if (err) break;
        if (!ReadFrame(left, pairs.Left(b), w, h) || !ReadFrame(right, pairs.Right(b), w, h)) break;
        const Parameters p = DefaultParams();
        int rc = sadgpu_submit_into(g, b, pairs.Left(b), w, pairs.Right(b), w, w, h, p.BlockSize, p.MaxDisparity, 0, h,
                                    maps.Map(b), w, &ticket[b]);
        if (rc) { err = rc; break; }
        busy[b] = true; ++submitted;
    }
    for (int i = 0; i < 2; ++i) {                               // drain in submission order
        const int b = (submitted + i) & 1;
        if (busy[b]) retire(b);
    }
    FreePinned(pairs); FreePinned(maps);
    if (err) fail("StreamSerialPairs", err);
    return delivered;
}

}  // namespace despair

// ---- C hooks so that the Python test-suite and bench.py can drive the C++ mirror --------------------------------
namespace {

struct BufferPort : despair::Port {      // a "serial port" over a memory buffer that answers reads in ragged pieces
    const uint8_t* p; size_t n, pos = 0; unsigned state;
    BufferPort(const uint8_t* p_, size_t n_, unsigned seed) : p(p_), n(n_), state(seed * 2654435761u + 1u) {}
    int Read(uint8_t* buf, int want) override
    {
        if (pos >= n) return 0;
        state = state * 1664525u + 1013904223u;
        size_t k = 1 + (state >> 8) % 1024;                     // 1..1024 bytes, like a UART driver
        k = std::min(k, std::min((size_t)want, n - pos));
        memcpy(buf, p + pos, k);
        pos += k;
        return (int)k;
    }
};

}  // namespace

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
        int rc = 0;
        despair::Gray o;
        try { o = despair::AssembleDisparityMap(*p.Out, l.Rect_, n, faithful_drop != 0); } catch (const std::exception&) { rc = -2; }
        feeder.join();
        p.In->Close();
        if (rc) return rc;
        memcpy(out, o.Pix.data(), (size_t)w * h);
        if (params_seen) { params_seen[0] = despair::DefaultParams().BlockSize; params_seen[1] = despair::DefaultParams().MaxDisparity; }
        return 0;
    } catch (const std::exception&) { return -1; }
}

// The OutputCamera frame loop as the reference runs it (pkg/camera/output.go:129-210 minus the PNG files): ONE pipeline of `workers`
// workers for the whole run, every frame new image objects in pageable memory (`n_distinct` frame pairs of w*h bytes at
// left / right are cycled; each is copied into a fresh Gray, like image.NewGray + the conversion loop of :141-162), row bands of
// max(1, h / (workers*4)) rows (:172), AssembleDisparityMap (:190).  Times `iters` frames after `warm` untimed ones and returns
// the map of the last frame.  reuse_objects != 0 refills the SAME two Gray objects in place every frame instead (the stale-
// snapshot path of the frame cache).
int despair_host_output_camera_loop(const uint8_t* left, const uint8_t* right, int n_distinct, int w, int h, int block_size,
                                    int max_disparity, int workers, int warm, int iters, int reuse_objects, uint8_t* out_last,
                                    double* us_per_frame)
{
    try {
        despair::SetDefaultParams(despair::Parameters{block_size, max_disparity});
        despair::Pipeline p = despair::SetupConcurrentSAD(workers);
        const int chunk = std::max(1, h / (workers * 4));                                             // output.go:172
        const int n = (h + chunk - 1) / chunk;                                                        // :173
        const size_t img = (size_t)w * h;
        despair::Gray keepL = despair::NewGray(despair::Rect(0, 0, w, h)), keepR = keepL, last;
        std::chrono::steady_clock::time_point t0;
        double total = 0;
        for (int f = 0; f < warm + iters; ++f) {
            const uint8_t* sl = left + (size_t)(f % n_distinct) * img;
            const uint8_t* sr = right + (size_t)(f % n_distinct) * img;
            despair::Gray freshL, freshR;
            if (!reuse_objects) { freshL = despair::NewGray(despair::Rect(0, 0, w, h)); freshR = freshL; }
            despair::Gray& l = reuse_objects ? keepL : freshL;
            despair::Gray& r = reuse_objects ? keepR : freshR;
            memcpy(l.Pix.data(), sl, img); memcpy(r.Pix.data(), sr, img);     // stands for PNG decode + gray conversion: outside the timed region
            t0 = std::chrono::steady_clock::now();                            // output.go:166 startTime
            for (int y = 0; y < h; y += chunk)                                // :176-187 (the reference sends from the same goroutine that assembles)
                p.In->Send(despair::InputChunk{&l, &r, despair::Rect(0, y, w, std::min(y + chunk, h))});
            last = despair::AssembleDisparityMap(*p.Out, l.Rect_, n);         // :190
            if (f >= warm) total += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        }
        p.In->Close();
        if (out_last) memcpy(out_last, last.Pix.data(), img);
        if (us_per_frame) *us_per_frame = iters > 0 ? total / iters : 0.0;
        return 0;
    } catch (const std::exception&) { return -1; }
}

// N2: decoded NRGBA pair -> disparity map, luma on the device (pkg/camera/output.go:129-210 without PNG files).
int despair_host_process_depth_map(const uint8_t* left_rgba, const uint8_t* right_rgba, int w, int h, int block_size,
                                   int max_disparity, uint8_t* out)
{
    try {
        despair::SetDefaultParams(despair::Parameters{block_size, max_disparity});
        despair::NRGBA l = despair::NewNRGBA(despair::Rect(0, 0, w, h)), r = l;
        memcpy(l.Pix.data(), left_rgba, (size_t)4 * w * h);
        memcpy(r.Pix.data(), right_rgba, (size_t)4 * w * h);
        despair::Gray o = despair::ProcessDepthMap(l, r);
        memcpy(out, o.Pix.data(), (size_t)w * h);
        return 0;
    } catch (const std::exception&) { return -1; }
}

// N3: two raw-gray byte streams ("serial ports" answering in ragged reads of 1..1024 bytes) -> n disparity maps.
int despair_host_serial_stream(const uint8_t* left_bytes, const uint8_t* right_bytes, size_t n_bytes_each, int w, int h,
                               int block_size, int max_disparity, int max_frames, unsigned seed, uint8_t* out)
{
    try {
        despair::SetDefaultParams(despair::Parameters{block_size, max_disparity});
        BufferPort pl(left_bytes, n_bytes_each, seed), pr(right_bytes, n_bytes_each, seed + 17);
        const size_t img = (size_t)w * h;
        return despair::StreamSerialPairs(pl, pr, w, h, max_frames,
                                          [&](int k, const uint8_t* map) { memcpy(out + (size_t)k * img, map, img); });
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

void despair_host_configure(int max_w, int max_h, int n_streams) { despair::ConfigureBackend(max_w, max_h, n_streams); }
void despair_host_shutdown(void) { despair::ShutdownBackend(); }

}  // extern "C"
