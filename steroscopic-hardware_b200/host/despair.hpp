// despair.hpp — C++ host-side mirror of the reference Go package pkg/despair over the sadgpu C ABI.
//
// The reference's host language (Go 1.24) is not available in the build image, so the layer that a
// `//go:build cuda` file would provide is written here in C++ with the same names,
// argument meaning and semantics, one to one:
//
//   Parameters / SetDefaultParams / DefaultParams      pkg/despair/params.go:8-37
//   InputChunk / OutputChunk                           pkg/despair/sad.go:12-21
//   SetupConcurrentSAD(numWorkers)                     pkg/despair/sad.go:29-113
//   RunSad(left, right, blockSize, maxDisparity)       pkg/despair/sad.go:119-169
//   AssembleDisparityMap(out, dimensions, chunks)      pkg/despair/sad.go:172-202
// and, for the callers either side of the path:
//   ProcessDepthMap(left, right)                       pkg/camera/output.go:129-210 without the PNG round trips
//   ReadFrame(port, frame) / StreamSerialPairs(...)    pkg/camera/serial.go:238-326 straight into pinned frames
//
// Chan<T> is a bounded MPMC queue with Go channel semantics (blocking send/recv, close drains).
// The workers do not compute anything themselves: every chunk ends in sadgpu_compute_region (CUDA);
// the chunks of one frame pair share one whole-frame GPU pass inside the library.  There is no CPU
// path.  Deviations from the reference are the documented ones: all chunks are written
// by AssembleDisparityMap (the dropped-last-chunk bug of sad.go:179-184 is available behind
// `faithful_drop` for comparison tests only), parameters are read once per chunk exactly as in
// sad.go:51-53, and an error of the backend (block size > 31, image larger than the backend, no device)
// surfaces as std::runtime_error from AssembleDisparityMap / RunSad where Go would panic — never as a
// silently black map.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

struct sadgpu_ctx;

namespace despair {

struct Rectangle {                      // image.Rectangle
    int MinX = 0, MinY = 0, MaxX = 0, MaxY = 0;
    int Dx() const { return MaxX - MinX; }
    int Dy() const { return MaxY - MinY; }
};
inline Rectangle Rect(int x0, int y0, int x1, int y1) { return Rectangle{x0, y0, x1, y1}; }

struct Gray {                           // image.Gray: Pix, Stride, Rect (Rect.Min must be (0,0), as sad.go:228 assumes)
    std::vector<uint8_t> Pix;
    int Stride = 0;
    Rectangle Rect_;
};
Gray NewGray(Rectangle r);

struct NRGBA {                          // image.NRGBA: what image/png yields for an 8-bit RGBA file (pkg/camera/output.go:138)
    std::vector<uint8_t> Pix;           // 4 bytes per pixel, non-premultiplied
    int Stride = 0;
    Rectangle Rect_;
};
NRGBA NewNRGBA(Rectangle r);

struct Parameters {                     // pkg/despair/params.go:34-37
    int BlockSize;
    int MaxDisparity;
};
void SetDefaultParams(Parameters p);    // params.go:21-25
Parameters DefaultParams();             // params.go:28-30 (initial value {16, 64}, params.go:13-18)

struct InputChunk {                     // sad.go:12-15
    const Gray* Left = nullptr;
    const Gray* Right = nullptr;
    Rectangle Region;
};
struct OutputChunk {                    // sad.go:18-21
    std::vector<uint8_t> DisparityData;
    Rectangle Region;
};

// Buffered Go channel.  OutputCamera moves 160 chunks per frame through two of these (pkg/camera/output.go:176-190), so the
// hot path must cost what a goroutine channel costs: the queue is a bounded lock-free ring (per-cell sequence numbers,
// D. Vyukov's MPMC queue) — a contended std::mutex costs 1.6 us per operation on the GPU hosts, the ring 0.1 us — and
// blocking is poll-then-sleep: receivers and senders poll for some tens of microseconds before they sleep (a goroutine parks
// in ~100 ns, a futex sleep / wake pair costs ~10 us), at most half the cores poll at a time, and a wake-up system call is
// only made when somebody is asleep and more items are queued than there are pollers to take them.
// The ring holds the next power of two >= cap items (more buffering than Go's exact cap never blocks a correct program).
template <class T> class Chan {
public:
    explicit Chan(size_t cap) : cap_(cap ? cap : 1) {
        size_t n = 1;
        while (n < cap_) n <<= 1;
        mask_ = n - 1;
        cells_.reset(new Cell[n]);
        for (size_t i = 0; i < n; ++i) cells_[i].seq.store(i, std::memory_order_relaxed);
    }
    // Non-blocking forms (SetupConcurrentSAD runs several logical workers on one thread).
    bool TrySend(T& v) {                // moves from v on success
        if (closed_.load(std::memory_order_acquire)) throw std::runtime_error("send on closed channel");
        size_t pos = enq_.load(std::memory_order_relaxed);
        Cell* c;
        for (;;) {
            c = &cells_[pos & mask_];
            const intptr_t dif = (intptr_t)c->seq.load(std::memory_order_acquire) - (intptr_t)pos;
            if (dif == 0) { if (enq_.compare_exchange_weak(pos, pos + 1, std::memory_order_relaxed)) break; }
            else if (dif < 0) return false;                                    // full
            else pos = enq_.load(std::memory_order_relaxed);
        }
        c->data = std::move(v);
        c->seq.store(pos + 1, std::memory_order_release);
        if (recv_sleepers_.load(std::memory_order_seq_cst) > 0 && (int)Size() > spinners_.load(std::memory_order_relaxed)) {
            std::lock_guard<std::mutex> l(sm_);
            not_empty_.notify_one();
        }
        return true;
    }
    int TryRecv(T& out) {               // 1 = received, 0 = empty, -1 = closed and drained
        for (int attempt = 0;; ++attempt) {
            size_t pos = deq_.load(std::memory_order_relaxed);
            Cell* c;
            bool got = false;
            for (;;) {
                c = &cells_[pos & mask_];
                const intptr_t dif = (intptr_t)c->seq.load(std::memory_order_acquire) - (intptr_t)(pos + 1);
                if (dif == 0) { if (deq_.compare_exchange_weak(pos, pos + 1, std::memory_order_relaxed)) { got = true; break; } }
                else if (dif < 0) break;                                       // empty
                else pos = deq_.load(std::memory_order_relaxed);
            }
            if (got) {
                out = std::move(c->data);
                c->seq.store(pos + mask_ + 1, std::memory_order_release);
                // hysteresis: one wake-up for many free places, not one per place (senders never sleep without a time limit)
                if (send_sleepers_.load(std::memory_order_seq_cst) > 0 && Size() <= (mask_ + 1) / 2) {
                    std::lock_guard<std::mutex> l(sm_);
                    not_full_.notify_all();
                }
                return 1;
            }
            if (!closed_.load(std::memory_order_acquire)) return 0;
            if (attempt) return -1;                                            // closed: one more look for an item sent before Close
        }
    }
    void Send(T v) {
        for (;;) {
            if (TrySend(v)) return;
            if (spin_until([&] { return Size() <= mask_ || closed_.load(std::memory_order_relaxed); })) continue;
            std::unique_lock<std::mutex> l(sm_);
            send_sleepers_.fetch_add(1, std::memory_order_seq_cst);
            if (Size() > mask_ && !closed_.load(std::memory_order_relaxed)) not_full_.wait_for(l, std::chrono::microseconds(100));
            send_sleepers_.fetch_sub(1, std::memory_order_seq_cst);
        }
    }
    bool Recv(T& out) {                 // false once the channel is closed and drained (Go: v, ok := <-ch)
        for (;;) {
            const int r = TryRecv(out);
            if (r) return r > 0;
            if (spin_until([&] { return Size() > 0 || closed_.load(std::memory_order_relaxed); })) continue;
            std::unique_lock<std::mutex> l(sm_);
            recv_sleepers_.fetch_add(1, std::memory_order_seq_cst);
            if (Size() == 0 && !closed_.load(std::memory_order_relaxed)) not_empty_.wait_for(l, std::chrono::milliseconds(2));
            recv_sleepers_.fetch_sub(1, std::memory_order_seq_cst);
        }
    }
    void WaitNotFullFor(std::chrono::microseconds d) {
        std::unique_lock<std::mutex> l(sm_);
        send_sleepers_.fetch_add(1, std::memory_order_seq_cst);
        if (Size() > mask_ && !closed_.load(std::memory_order_relaxed)) not_full_.wait_for(l, d);
        send_sleepers_.fetch_sub(1, std::memory_order_seq_cst);
    }
    void Close() {
        closed_.store(true, std::memory_order_release);
        std::lock_guard<std::mutex> l(sm_);
        not_empty_.notify_all();
        not_full_.notify_all();
    }
    size_t Cap() const { return mask_ + 1; }
    size_t Size() const {               // a glance (pollers): exact when nobody is in the middle of an operation
        const size_t e = enq_.load(std::memory_order_relaxed), d = deq_.load(std::memory_order_relaxed);
        return e > d ? e - d : 0;
    }
    // Error side band (the Go structs have no error field): the first backend failure of a worker is recorded on the output
    // channel; AssembleDisparityMap reports it instead of returning a map with holes.
    void Fail(const std::string& what) { std::lock_guard<std::mutex> l(sm_); if (err_.empty()) err_ = what; }
    std::string TakeError() { std::lock_guard<std::mutex> l(sm_); std::string e; e.swap(err_); return e; }
private:
    struct Cell { std::atomic<size_t> seq; T data; };
    template <class Pred> bool spin_until(Pred ready) {              // true: the condition held before the spin budget ran out
        if (ready()) return true;
        static const int limit = std::max(1, (int)std::thread::hardware_concurrency() / 2 + 1);
        if (spinners_.fetch_add(1, std::memory_order_relaxed) >= limit) { spinners_.fetch_sub(1, std::memory_order_relaxed); return ready(); }
        const auto t0 = std::chrono::steady_clock::now();
        bool ok = false;
        for (;;) {
            for (int i = 0; i < 64 && !ok; ++i) {
                ok = ready();
#if defined(__x86_64__) || defined(__i386__)
                if (!ok) __builtin_ia32_pause();
#endif
            }
            if (ok || std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(50)) break;
        }
        spinners_.fetch_sub(1, std::memory_order_relaxed);
        return ok || ready();
    }
    size_t cap_, mask_ = 0;
    std::unique_ptr<Cell[]> cells_;
    alignas(64) std::atomic<size_t> enq_{0};
    alignas(64) std::atomic<size_t> deq_{0};
    alignas(64) std::atomic<bool> closed_{false};
    std::atomic<int> spinners_{0}, recv_sleepers_{0}, send_sleepers_{0};
    std::mutex sm_;                      // sleepers only
    std::condition_variable not_empty_, not_full_;
    std::string err_;
};

struct Pipeline {                       // the (chan<- InputChunk, <-chan OutputChunk) pair of SetupConcurrentSAD
    std::shared_ptr<Chan<InputChunk>> In;
    std::shared_ptr<Chan<OutputChunk>> Out;
};

// numWorkers <= 0 => hardware_concurrency*4 (sad.go:32-34); channels buffered 2*numWorkers (:36-37).
// Go runs its numWorkers goroutines on GOMAXPROCS threads; here the numWorkers logical workers run on at most half the
// cores' worth of threads, each of which may hold several finished chunks, so that the pipeline keeps the reference's capacity
// of 2n + n + 2n chunks in flight (OutputCamera sends all 160 chunks of a frame before it starts receiving,
// pkg/camera/output.go:176-190: with fewer than n chunks held by workers it would deadlock).
Pipeline SetupConcurrentSAD(int numWorkers);
Gray RunSad(const Gray& left, const Gray& right, int blockSize, int maxDisparity);
Gray AssembleDisparityMap(Chan<OutputChunk>& outputChan, Rectangle dimensions, int chunks, bool faithful_drop = false);

// OutputCamera.processDepthMap (pkg/camera/output.go:129-210) without the PNG files: the decoded colour pair goes to the GPU
// as it is, the per-pixel color.GrayModel.Convert loop (:143-147, :158-162) runs there with Go's exact arithmetic
// (sadgpu_compute_nrgba), parameters are the current DefaultParams() (:169).
Gray ProcessDepthMap(const NRGBA& left, const NRGBA& right);

// Video path (examples/run.stream.go:33-67 without the per-frame channel traffic): frames live in the
// backend's pinned pool, n pairs travel per GPU call (sadgpu_submit_batch_into), up to `depth` calls are in flight.
// PinnedFrames owns [n][2][h][w] input pairs or [n][h][w] maps in pinned memory; Left(i)/Right(i)/Map(i) address one plane.
struct PinnedFrames {
    uint8_t* base = nullptr;
    int n = 0, planes = 0, w = 0, h = 0;
    uint8_t* Left(int i) const { return base + (size_t)i * planes * w * h; }
    uint8_t* Right(int i) const { return Left(i) + (size_t)w * h; }
    uint8_t* Map(int i) const { return base + (size_t)i * w * h; }
};
PinnedFrames NewPinnedPairs(int n, int w, int h);
PinnedFrames NewPinnedMaps(int n, int w, int h);
void FreePinned(PinnedFrames& f);
// Disparity maps of `pairs` into `maps` with the current DefaultParams(), `batch` pairs per call.
void StreamSad(const PinnedFrames& pairs, PinnedFrames& maps, int batch, int depth = 4);

// Serial ingest (pkg/camera/serial.go:238-326): a camera delivers width*height raw gray bytes per frame in reads of at most
// 1024 bytes (:275-276).  ReadFrame lands them directly in a pinned plane — no intermediate buffer, no per-pixel SetGray
// (:304-311).  Port::Read returns the number of bytes read (> 0) or <= 0 on error, like serial.Port.Read.
struct Port {
    virtual ~Port() = default;
    virtual int Read(uint8_t* buf, int n) = 0;
};
bool ReadFrame(Port& port, uint8_t* pix, int w, int h);
// Two cameras -> disparity maps: frame k+1 is read into the second pinned pair while frame k is on the GPU (the pair in
// flight is borrowed by the upload, include/sadgpu.h).  sink(k, map) sees each w*h map in pinned memory, in order.
// Returns the number of frames delivered (a short read on either port ends the stream).
int StreamSerialPairs(Port& left, Port& right, int w, int h, int max_frames,
                      const std::function<void(int, const uint8_t*)>& sink);

// Tile planner of RunSad (sad.go:128-153), exposed for tests.
std::vector<Rectangle> RunSadChunks(Rectangle dims, int numCPU);

// Configuration of the backing context (additive; the reference has no equivalent).
void ConfigureBackend(int max_w, int max_h, int n_streams);
void ShutdownBackend();

}  // namespace despair
