/*
 * sadgpu.h — C ABI of libsadgpu.so: the B200 (sm_100a) drop-in for the SAD block-matching
 * disparity path of conneroisu/steroscopic-hardware, pkg/despair.
 *
 * This is the boundary a cgo build of pkg/despair binds (see INTEGRATION.md for the Go
 * stub).  Plain pointers and sizes only; no C++ types, no exceptions, no torch types.
 *
 * What each entry point replaces in the reference:
 *   sadgpu_compute / sadgpu_submit+sadgpu_wait
 *       the worker body of SetupConcurrentSAD            pkg/despair/sad.go:47-102
 *       (per-pixel scan :55-95 and SumAbsoluteDifferences :205-244), for the rows
 *       [y0,y1) of one left/right *image.Gray pair — i.e. one or many InputChunk regions
 *       (pkg/despair/sad.go:12-15) of the same frame — and the copy loop of
 *       AssembleDisparityMap                              pkg/despair/sad.go:186-197
 *   sadgpu_compute_region
 *       the same worker body for ONE InputChunk (sad.go:12-15) with a region-local result buffer, exactly
 *       OutputChunk.DisparityData (sad.go:18-21, :48-50, :91); the chunks of one frame pair share one GPU pass
 *   sadgpu_compute_sharded
 *       the row-band fan-out of OutputCamera.processDepthMap pkg/camera/output.go:172-190,
 *       with GPUs in place of goroutines
 *   block_size / max_disparity arguments
 *       Parameters{BlockSize, MaxDisparity}               pkg/despair/params.go:34-37
 *   sadgpu_compute_device
 *       same computation on device-resident buffers (kernel-only timing, bench.py)
 *
 * Semantics are bit-exact with the reference for every pixel:
 * window side 2*(block_size/2)+1, d = 0..max_disparity inclusive, candidate d skipped when
 * X-d < 0, windows clamped as in sad.go:212-218, strict '<' argmin in ascending d (lowest d
 * wins ties), out = uint8(best*255/max_disparity).  Documented deviations: all rows are
 * written (the dropped-last-chunk bug of sad.go:179-184 is not reproduced); parameters are
 * read once per call; invalid arguments return an error instead of panicking.
 *
 * There is NO CPU fallback: every compute entry point fails with a CUDA error code when no
 * sm_100 device is usable.
 */
#ifndef SADGPU_H
#define SADGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sadgpu_ctx sadgpu_ctx;

/* Error codes: 0 = OK, negative = argument / state errors, positive = cudaError_t passthrough. */
#define SADGPU_OK          0
#define SADGPU_EINVAL    (-1)   /* null pointer, bad size/stride, block_size or max_disparity out of range */
#define SADGPU_ERANGE    (-2)   /* w/h exceed the context's max_w/max_h, bad stream/device index, bad rows */
#define SADGPU_ENOMEM    (-3)   /* host allocation failed */
#define SADGPU_EBUSY     (-4)   /* stream slot already has a frame in flight / bad ticket */
#define SADGPU_ENODEV    (-5)   /* no usable CUDA device */

/* Supported parameter surface (a superset of the UI range B 3..31, D 16..256,
 * cmd/handlers/params.go:37,51; the library itself accepts any ints, params.go:21-25). */
#define SADGPU_MAX_BLOCK_SIZE    31
#define SADGPU_MAX_DISPARITY     256

/* Optional tuning overrides for tests/benchmarks; 0 = let the planner choose. */
typedef struct sadgpu_tuning {
    int rows_per_batch;     /* RB: rows staged through shared memory per iteration */
    int band_rows;          /* BH: output rows per CTA band                           */
    int groups_per_chunk;   /* disparity groups (4 disparities each) per CTA chunk     */
    int kernel_variant;     /* 0 = auto, 2 = phase-alternating register-ring kernel (block_size <= 15),
                               3 = warp-specialised kernel (block_size <= 9; the default there: chunks of 33 / 17 / 9 / 5 disparity groups
                                   on 1 / 2 / 3 / 6 column strips per CTA, by max_disparity),
                               4 = phase-alternating large-window kernel (block_size 16..31),
                               6 = H-ring mbarrier-pipelined kernel (block_size 10..31),
                               7 = warp-specialised kernel with a shared-memory ring (block_size 10..31; the default for block_size >= 18 and
                                   for block_size 16, 17 with max_disparity <= 32).  It loads its tiles with TMA only: images whose base,
                                   pitch or frame stride is not a multiple of 16 bytes take variant 6 or 4 instead.
                               1 and 5 (removed kernels) are rejected. */
    int reserved[4];        /* [0]: frames per launch (sadgpu_plan_describe only); [1]: developer flags (profile builds of the ring kernel);
                               [2] = 1: do not use TMA tile loads in the warp-specialised kernel; = 2: fail with SADGPU_EINVAL
                               instead of substituting another kernel when variant 7 cannot use TMA */
} sadgpu_tuning;

int  sadgpu_device_count(void);

/* devices == NULL  => devices 0..n_devices-1, or — when the environment variable SADGPU_DEVICES holds a comma-separated list
 * of CUDA device ordinals — the first n_devices entries of that list (SADGPU_ERANGE if it is shorter).  n_streams logical camera streams are created;
 * stream s lives on devices[s % n_devices] and owns a CUDA stream, pinned upload/download
 * buffers and device buffers sized for max_w x max_h.  Nothing is allocated per frame. */
int  sadgpu_create(const int *devices, int n_devices, int max_w, int max_h, int n_streams,
                   sadgpu_ctx **out);
void sadgpu_destroy(sadgpu_ctx *ctx);

/* Synchronous: stage -> H2D -> kernel -> D2H -> copy rows [y0,y1) into out.  `out` addresses
 * the FULL map: row y is written at out + y*out_stride, so several calls with disjoint row
 * ranges assemble one Pix slice exactly like AssembleDisparityMap does (sad.go:186-197).
 * Full frame: y0 = 0, y1 = h.  Only rows [y0-h, y1+h) of left/right are read. */
int  sadgpu_compute(sadgpu_ctx *ctx, int stream,
                    const uint8_t *left, int left_stride, const uint8_t *right, int right_stride,
                    int w, int h, int block_size, int max_disparity, int y0, int y1,
                    uint8_t *out, int out_stride);

/* ONE InputChunk of a frame pair (pkg/despair/sad.go:12-15): region [x0,x1) x [y0,y1) in image coordinates; `out` is REGION-LOCAL,
 * row r of the region at out + r*out_stride (out_stride >= x1-x0) — with out_stride = x1-x0 exactly OutputChunk.DisparityData
 * (sad.go:48-50, :91).  No stream argument: calls that name the same frame pair (same left / right addresses, strides, size
 * and parameters) share ONE whole-frame GPU pass — the first call snapshots the pair into pinned memory (callers that arrive
 * meanwhile help copying), uploads and launches; the others wait for it and slice their rectangle out of the pinned result.
 * The reference's callers send up to 160 three-row chunks per frame (pkg/camera/output.go:172-187): this is the entry point
 * that makes the UNCHANGED SetupConcurrentSAD worker a real drop-in.  A chunk is only served from a shared pass after the
 * rows of its own images that influence its region ([y0-h, y1+h)) compared equal to the snapshot; otherwise the frame is
 * recomputed from the caller's current pixels — reusing an image object for new pixels is safe.  Thread-safe; no caller
 * pointer is retained (the addresses are remembered as keys and only ever dereferenced inside a call that passed them). */
int  sadgpu_compute_region(sadgpu_ctx *ctx,
                           const uint8_t *left, int left_stride, const uint8_t *right, int right_stride,
                           int w, int h, int block_size, int max_disparity,
                           int x0, int y0, int x1, int y1, uint8_t *out, int out_stride);
/* Counters of the frame cache: region calls, whole-frame GPU passes they cost, chunks that met a stale snapshot. */
int  sadgpu_region_stats(sadgpu_ctx *ctx, long long *calls, long long *frames, long long *stale);

/* sadgpu_compute for a colour pair as Go decodes it: left / right are 8-bit NON-premultiplied RGBA planes,
 * 4 bytes per pixel — the Pix of an *image.NRGBA, what image/png returns for an 8-bit RGBA PNG — and the 8-bit luma is taken
 * on the device with the exact arithmetic of color.GrayModel.Convert (pkg/camera/output.go:143-147, :158-162; the same values
 * as convertGenericToGray, pkg/despair/gray.go:43-58), so the per-pixel conversion loop of processDepthMap disappears.
 * Whole frame, synchronous; strides in bytes (>= 4*w). */
int  sadgpu_compute_nrgba(sadgpu_ctx *ctx, int stream,
                          const uint8_t *left_rgba, int left_stride, const uint8_t *right_rgba, int right_stride,
                          int w, int h, int block_size, int max_disparity, uint8_t *out, int out_stride);

/* Asynchronous pair for per-camera pipelining.  submit enqueues H2D + kernel + D2H on the stream; wait blocks until that
 * frame is done and copies rows [y0,y1) of the result to `out` (full-map addressing).
 * Ownership of the inputs: caller memory that is NOT from sadgpu_host_alloc (a Go Pix slice) is copied into the stream's
 * pinned staging buffer before submit returns — no caller pointer is retained (cgo rule).  Inputs that DO live in
 * sadgpu_host_alloc memory are uploaded in place and are therefore BORROWED until the upload has run: do not overwrite them
 * before sadgpu_wait_uploaded(ticket) (or sadgpu_wait) has returned — a camera that refills one pinned frame per capture
 * must double-buffer or wait for the upload first.  A stream holds one frame in flight: a second submit
 * (or sadgpu_compute) on a busy stream returns SADGPU_EBUSY; concurrent sadgpu_compute calls on one stream serialise. */
int  sadgpu_submit(sadgpu_ctx *ctx, int stream,
                   const uint8_t *left, int left_stride, const uint8_t *right, int right_stride,
                   int w, int h, int block_size, int max_disparity, int y0, int y1,
                   uint64_t *ticket);
int  sadgpu_wait(sadgpu_ctx *ctx, uint64_t ticket, uint8_t *out, int out_stride);
/* Blocks until the H2D copies of the submitted frame (or batch) have run: pinned-pool inputs may be overwritten again. */
int  sadgpu_wait_uploaded(sadgpu_ctx *ctx, uint64_t ticket);

/* sadgpu_submit with the destination named up front: `out` (full-map addressing, rows [y0,y1) written) must lie in
 * memory from sadgpu_host_alloc, so the D2H copy lands in it directly and sadgpu_wait(ctx, ticket, NULL, 0) only
 * synchronises — no host-side copy on either side of the frame (the AssembleDisparityMap copy loop,
 * pkg/despair/sad.go:186-197, disappears; a Go image.Gray can wrap the pinned block).
 * SADGPU_EINVAL when `out` is not pool memory. */
int  sadgpu_submit_into(sadgpu_ctx *ctx, int stream,
                        const uint8_t *left, int left_stride, const uint8_t *right, int right_stride,
                        int w, int h, int block_size, int max_disparity, int y0, int y1,
                        uint8_t *out, int out_stride, uint64_t *ticket);

/* Video streams (BASELINE configs[4], examples/run.stream.go:33-67): n_frames frame pairs per call.  `pairs` is
 * [n_frames][2][h][w] bytes (left plane, right plane, next pair ...; w a multiple of 4), `out` is [n_frames][h][w] and must lie
 * in memory from sadgpu_host_alloc.  The batch travels as ONE H2D copy, ONE kernel launch that is many waves deep, and ONE
 * D2H copy; sadgpu_wait(ctx, ticket, NULL, 0) synchronises.  sadgpu_reserve_batch sizes every stream's buffers for
 * max_frames pairs up front (call it while no frame is in flight); a stream that meets a larger batch grows its own
 * buffers on that call (n_frames <= 256). */
int  sadgpu_reserve_batch(sadgpu_ctx *ctx, int max_frames);
int  sadgpu_submit_batch_into(sadgpu_ctx *ctx, int stream, int n_frames, const uint8_t *pairs, int w, int h,
                              int block_size, int max_disparity, uint8_t *out, uint64_t *ticket);

/* One frame split into row bands with a block_size/2 halo over ALL devices of the context, host-side gather into out.
 * Band i runs on stream i (device i % n_devices); a context with spare streams uses up to four bands per device, so that
 * the upload of a band overlaps the kernel of the previous one — also the lowest-latency way to run ONE large frame on ONE
 * device (streams 0..n-1 must be idle). */
int  sadgpu_compute_sharded(sadgpu_ctx *ctx,
                            const uint8_t *left, int left_stride, const uint8_t *right, int right_stride,
                            int w, int h, int block_size, int max_disparity,
                            uint8_t *out, int out_stride);

/* Device-resident: dL/dR/dOut are device pointers on `device` (index into the context's
 * device list); rows [y0,y1) of dOut are written; enqueued on cuda_stream (a cudaStream_t,
 * NULL = the legacy default stream) without synchronising.  tuning may be NULL.  Calls on different cuda_streams of one
 * device are independent: scratch for a chunked disparity range (max_disparity > 128 at block_size <= 9, ...) is a
 * stream-ordered allocation (cudaMallocAsync / cudaFreeAsync on cuda_stream) private to the call. */
int  sadgpu_compute_device(sadgpu_ctx *ctx, int device,
                           const uint8_t *dL, size_t pitch_l, const uint8_t *dR, size_t pitch_r,
                           int w, int h, int block_size, int max_disparity, int y0, int y1,
                           uint8_t *dOut, size_t pitch_out, void *cuda_stream,
                           const sadgpu_tuning *tuning);

/* Batch of n_frames frame pairs in ONE launch (video streams, BASELINE configs[4]): frame f lives at
 * base + f*frame_stride (bytes).  Grid = strips x bands x frames, so the launch is many waves deep. */
int  sadgpu_compute_device_batch(sadgpu_ctx *ctx, int device, int n_frames,
                                 const uint8_t *dL, size_t pitch_l, size_t frame_stride_l,
                                 const uint8_t *dR, size_t pitch_r, size_t frame_stride_r,
                                 int w, int h, int block_size, int max_disparity, int y0, int y1,
                                 uint8_t *dOut, size_t pitch_out, size_t frame_stride_out,
                                 void *cuda_stream, const sadgpu_tuning *tuning);

/* Go-exact 8-bit luma of an interleaved 8-bit pixel plane on the device — what the reference
 * does to a decoded PNG before the SAD path: mode 0 = NRGBA (pkg/despair/gray.go:43-58 generic path ==
 * color.GrayModel.Convert, pkg/camera/output.go:145,160; channels 4, or 3 = opaque), mode 1 = opaque RGB with the
 * intended 16-bit formula, mode 2 = convertRGBAToGray as written (gray.go:20-40: 8-bit values >> 24, always 0). */
#define SADGPU_GRAY_NRGBA8         0
#define SADGPU_GRAY_RGB8_INTENDED  1
#define SADGPU_GRAY_RGBX8_LOADPNG  2
int  sadgpu_gray_device(sadgpu_ctx *ctx, int device, const uint8_t *dSrc, size_t src_pitch, int channels, int mode,
                        int w, int h, uint8_t *dGray, size_t gray_pitch, void *cuda_stream);

/* Post-processing hooks.  Nothing in the reference corresponds to them (cmd/handlers/stream.go:14-37 only serves
 * the files OutputCamera writes): they are strictly additive, run only when called, and leave every entry point above bit-exact.
 *   sadgpu_median3_device   3x3 median of a map (window clamped to the image); dDst != dSrc
 *   sadgpu_lrcheck_device   left-right consistency: a left-map pixel (x, y) with decoded disparity d = round(v*D/255) survives iff the
 *                           right-referenced map at (x - d, y) decodes to within `tolerance` of d; otherwise (or when x - d < 0) it
 *                           becomes invalid_value
 *   sadgpu_compute_checked  host call: left-referenced map (the bit-exact path), right-referenced map (the same path on the mirrored
 *                           pair with the roles swapped), consistency check, optional median; synchronous, whole frame */
int  sadgpu_median3_device(sadgpu_ctx *ctx, int device, const uint8_t *dSrc, size_t src_pitch, int w, int h,
                           uint8_t *dDst, size_t dst_pitch, void *cuda_stream);
int  sadgpu_lrcheck_device(sadgpu_ctx *ctx, int device, const uint8_t *dLeftMap, size_t left_pitch,
                           const uint8_t *dRightMap, size_t right_pitch, int w, int h, int max_disparity, int tolerance,
                           int invalid_value, uint8_t *dDst, size_t dst_pitch, void *cuda_stream);
int  sadgpu_compute_checked(sadgpu_ctx *ctx, int stream,
                            const uint8_t *left, int left_stride, const uint8_t *right, int right_stride,
                            int w, int h, int block_size, int max_disparity, int tolerance, int invalid_value, int median,
                            uint8_t *out, int out_stride);

/* Pinned host memory from the context's pool: frames that already live here are uploaded
 * without the staging memcpy (cameras write straight into it). */
void *sadgpu_host_alloc(sadgpu_ctx *ctx, size_t bytes);
void  sadgpu_host_free(sadgpu_ctx *ctx, void *p);

/* Introspection used by bench.py / tests. */
int  sadgpu_debug_read(sadgpu_ctx *ctx, int device, uint32_t *host, int n_words);   /* developer: per-role cycle counters (library built with -DRING_PROFILE, tuning.reserved[1] & 4) */
int  sadgpu_last_launch_count(sadgpu_ctx *ctx);    /* kernels launched by the last compute call */
int  sadgpu_plan_describe(int w, int h, int block_size, int max_disparity, int y0, int y1,
                          const sadgpu_tuning *tuning, char *buf, size_t buflen);
const char *sadgpu_strerror(int code);
const char *sadgpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SADGPU_H */
