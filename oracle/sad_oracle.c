/*
 * sad_oracle.c — CPU ORACLE for the SAD block-matching disparity path of
 * conneroisu/steroscopic-hardware (pkg/despair/sad.go).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (libsadgpu.so) never links, loads or calls anything in this directory.
 *
 * The Go toolchain is absent from the build image, so this file is a
 * restatement of the reference algorithm in plain C:
 *
 *   oracle_sum_abs_diff      <- pkg/despair/sad.go:205-244  SumAbsoluteDifferences
 *   oracle_region_literal    <- pkg/despair/sad.go:55-95    worker body (per-pixel scan)
 *   oracle_frame_literal_mt  <- pkg/camera/output.go:172-187 row-band chunking
 *                               + pkg/despair/sad.go:41-104 worker pool (pthreads)
 *   oracle_frame_box         <- closed form of the same function:
 *                               zero-padded separable box filter, O(D) per pixel;
 *                               proven equal to the literal form by tests/test_oracle.py
 *
 * Parity pin status: the reference has no Go test for this path.  The
 * oracle is pinned by (1) the FPGA golden vectors hardware/mems/exp_disp_p.mem and
 * hardware/exp_disp.mem on the interior rectangle where the conventions coincide,
 * (2) the reference's own C golden generator hardware/sad.c compiled unmodified into
 * oracle/_ref/hw_sad and run on random patches, (3) SHA-256 pins of SURVEY.md §8(c).
 * It has NOT been compared against a real Go binary ("parity pinned by golden vectors
 * and an independent restatement, not by the Go executable").
 *
 * All images are 8-bit, row-major, Rect.Min == (0,0) (the reference indexes
 * ly*Stride+x, sad.go:228-229, so it assumes the same).
 */
#include <limits.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* pkg/despair/sad.go:205-244 — literal.  Variable names follow the Go source. */
ORACLE_API int oracle_sum_abs_diff(const uint8_t *leftPix, int leftStride, int leftMaxX, int leftMaxY,
                                   const uint8_t *rightPix, int rightStride, int rightMaxX, int rightMaxY,
                                   int leftX, int leftY, int rightX, int rightY, int blockSize)
{
    int halfSize = blockSize / 2;                                  /* :209 */
    int leftMinY = leftY - halfSize > 0 ? leftY - halfSize : 0;    /* :212 */
    int leftMaxYw = leftY + halfSize + 1 < leftMaxY ? leftY + halfSize + 1 : leftMaxY; /* :213 */
    int leftMinX = leftX - halfSize > 0 ? leftX - halfSize : 0;    /* :214 */
    int leftMaxXw = leftX + halfSize + 1 < leftMaxX ? leftX + halfSize + 1 : leftMaxX; /* :215 */
    int rightMinY = rightY - halfSize > 0 ? rightY - halfSize : 0; /* :217 */
    int rightMinX = rightX - halfSize > 0 ? rightX - halfSize : 0; /* :218 */
    int sad = 0, lx;
    for (int ly = leftMinY; ly < leftMaxYw; ly++) {                /* :224 */
        if (rightMinY + (ly - leftMinY) >= rightMaxY)              /* :225 */
            break;
        int leftRowStart = ly * leftStride + leftMinX;             /* :228 */
        int rightRowStart = (rightMinY + (ly - leftMinY)) * rightStride + rightMinX; /* :229 */
        for (lx = leftMinX; lx < leftMaxXw; lx++) {                /* :230 */
            if (rightMinX + (lx - leftMinX) >= rightMaxX)          /* :231 */
                break;
            int diff = (int)leftPix[leftRowStart + lx - leftMinX] -
                       (int)rightPix[rightRowStart + (rightMinX + (lx - leftMinX)) - rightMinX]; /* :234-235 */
            if (diff < 0)
                diff = -diff;
            sad += diff;                                           /* :239 */
        }
    }
    return sad;
}

/* pkg/despair/sad.go:55-95 — literal worker body for one chunk region
 * [rx0,rx1) x [ry0,ry1).  out is region-local row-major, length Dx*Dy (:48-50, :91).
 * early_exit != 0 keeps the `break` on sad == 0 (:84-86); it never changes the result. */
ORACLE_API void oracle_region_literal(const uint8_t *left, int leftStride, const uint8_t *right, int rightStride,
                                      int w, int h, int rx0, int ry0, int rx1, int ry1,
                                      int blockSize, int maxDisparity, int early_exit, uint8_t *out)
{
    int dx = rx1 - rx0, dy = ry1 - ry0;
    for (int y = 0; y < dy; y++) {                                 /* :55 */
        int globalY = ry0 + y;                                     /* :56 */
        for (int x = 0; x < dx; x++) {                             /* :57 */
            int minSAD = INT_MAX;                                  /* :59 math.MaxInt32 */
            int bestDisparity = 0;                                 /* :60 */
            for (int d = 0; d <= maxDisparity; d++) {              /* :62 */
                if ((rx0 + x) - d < 0)                             /* :64-67 Rect.Min.X == 0 */
                    continue;
                int sad = oracle_sum_abs_diff(left, leftStride, w, h, right, rightStride, w, h,
                                              rx0 + x, globalY, rx0 + x - d, globalY, blockSize); /* :69-77 */
                if (sad < minSAD) {                                /* :79 strict */
                    minSAD = sad;
                    bestDisparity = d;
                    if (early_exit && sad == 0)                    /* :84-86 */
                        break;
                }
            }
            out[y * dx + x] = (uint8_t)((bestDisparity * 255) / maxDisparity); /* :91-93 */
        }
    }
}

/* Closed form: AD_d(x,y) = |L(x,y) - R(x-d,y)| for 0 <= x-d, zero
 * outside the image; S_d = (2h+1)^2 zero-padded box sum of AD_d; candidates {0} if X < h,
 * else d in [0, min(D, X-h)]; lowest d wins ties; out = best*255/D.
 * out has its own stride so a row range [y0,y1) can be written into a larger map. */
ORACLE_API int oracle_frame_box(const uint8_t *left, int leftStride, const uint8_t *right, int rightStride,
                                int w, int h, int blockSize, int maxDisparity, int y0, int y1,
                                uint8_t *out, int outStride)
{
    if (w <= 0 || h <= 0 || blockSize < 1 || maxDisparity < 1 || y0 < 0 || y1 > h || y0 > y1)
        return -1;
    const int half = blockSize / 2;
    int32_t *col = (int32_t *)malloc(sizeof(int32_t) * (size_t)w);
    int32_t *best = (int32_t *)malloc(sizeof(int32_t) * (size_t)w * (size_t)(y1 - y0));
    int32_t *bestd = (int32_t *)calloc((size_t)w * (size_t)(y1 - y0), sizeof(int32_t));
    if (!col || !best || !bestd) { free(col); free(best); free(bestd); return -2; }
    for (size_t i = 0; i < (size_t)w * (size_t)(y1 - y0); i++) best[i] = INT32_MAX;
    for (int d = 0; d <= maxDisparity; d++) {
        /* column sums for output row y0: rows [y0-half, y0+half] clipped */
        memset(col, 0, sizeof(int32_t) * (size_t)w);
        for (int yy = y0 - half; yy <= y0 + half; yy++) {
            if (yy < 0 || yy >= h) continue;
            for (int x = d; x < w; x++) {
                int v = (int)left[yy * leftStride + x] - (int)right[yy * rightStride + x - d];
                col[x] += v < 0 ? -v : v;
            }
        }
        for (int y = y0; y < y1; y++) {
            if (y > y0) {   /* slide the column sums down one row */
                int add = y + half, sub = y - half - 1;
                if (add < h)
                    for (int x = d; x < w; x++) {
                        int v = (int)left[add * leftStride + x] - (int)right[add * rightStride + x - d];
                        col[x] += v < 0 ? -v : v;
                    }
                if (sub >= 0)
                    for (int x = d; x < w; x++) {
                        int v = (int)left[sub * leftStride + x] - (int)right[sub * rightStride + x - d];
                        col[x] -= v < 0 ? -v : v;
                    }
            }
            /* candidates exist only for X >= half + d */
            int xs = half + d;
            if (xs >= w) continue;
            int32_t s = 0;
            for (int xx = xs - half; xx <= xs + half && xx < w; xx++) s += col[xx];
            int32_t *b = best + (size_t)(y - y0) * w, *bd = bestd + (size_t)(y - y0) * w;
            for (int x = xs; x < w; x++) {
                if (x > xs) {
                    if (x + half < w) s += col[x + half];
                    s -= col[x - half - 1];
                }
                if (s < b[x]) { b[x] = s; bd[x] = d; }
            }
        }
    }
    for (int y = y0; y < y1; y++)
        for (int x = 0; x < w; x++)
            out[(size_t)(y - y0) * outStride + x] =
                (uint8_t)((bestd[(size_t)(y - y0) * w + x] * 255) / maxDisparity);
    free(col); free(best); free(bestd);
    return 0;
}

/* ---- threaded driver: the reference's production parallelisation ------------------
 * pkg/camera/output.go:172-187: chunkSize = max(1, H/(32*4)) rows, full-width bands;
 * pkg/despair/sad.go:41-104: a pool of workers pulls chunks.  Used as the CPU baseline. */
typedef struct {
    const uint8_t *left, *right;
    int leftStride, rightStride, w, h, blockSize, maxDisparity, chunkRows, y0, y1, early_exit;
    uint8_t *out; int outStride;
    int next;                /* next band index, guarded by mu */
    pthread_mutex_t mu;
} job_t;

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    uint8_t *tmp = (uint8_t *)malloc((size_t)j->w * (size_t)j->chunkRows);
    for (;;) {
        pthread_mutex_lock(&j->mu);
        int band = j->next++;
        pthread_mutex_unlock(&j->mu);
        int ys = j->y0 + band * j->chunkRows;
        if (ys >= j->y1) break;
        int ye = ys + j->chunkRows < j->y1 ? ys + j->chunkRows : j->y1;
        oracle_region_literal(j->left, j->leftStride, j->right, j->rightStride, j->w, j->h,
                              0, ys, j->w, ye, j->blockSize, j->maxDisparity, j->early_exit, tmp);
        for (int y = ys; y < ye; y++)   /* AssembleDisparityMap without the dropped chunk */
            memcpy(j->out + (size_t)(y - j->y0) * j->outStride, tmp + (size_t)(y - ys) * j->w, (size_t)j->w);
    }
    free(tmp);
    return NULL;
}

/* Rows [y0,y1) of the literal algorithm on `threads` pthreads (threads <= 0 -> 1). */
ORACLE_API int oracle_frame_literal_mt(const uint8_t *left, int leftStride, const uint8_t *right, int rightStride,
                                       int w, int h, int blockSize, int maxDisparity, int y0, int y1,
                                       int threads, int early_exit, uint8_t *out, int outStride)
{
    if (w <= 0 || h <= 0 || blockSize < 1 || maxDisparity < 1 || y0 < 0 || y1 > h || y0 > y1)
        return -1;
    if (threads <= 0) threads = 1;
    job_t j;
    j.left = left; j.right = right; j.leftStride = leftStride; j.rightStride = rightStride;
    j.w = w; j.h = h; j.blockSize = blockSize; j.maxDisparity = maxDisparity;
    j.chunkRows = h / (32 * 4) > 1 ? h / (32 * 4) : 1;            /* output.go:172 */
    j.y0 = y0; j.y1 = y1; j.early_exit = early_exit; j.out = out; j.outStride = outStride; j.next = 0;
    pthread_mutex_init(&j.mu, NULL);
    pthread_t *t = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int i = 0; i < threads; i++) pthread_create(&t[i], NULL, worker, &j);
    for (int i = 0; i < threads; i++) pthread_join(t[i], NULL);
    free(t);
    pthread_mutex_destroy(&j.mu);
    return 0;
}
