// mock of the sadgpu C ABI: compute_region copies bytes from left (no GPU) — measures the host pipeline alone
#include "../../../include/sadgpu.h"
#include <cstring>
#include <cstdlib>
struct sadgpu_ctx { int x; };
extern "C" {
int sadgpu_device_count(void) { return 1; }
int sadgpu_create(const int*, int, int, int, int, sadgpu_ctx** out) { *out = new sadgpu_ctx(); return 0; }
void sadgpu_destroy(sadgpu_ctx* c) { delete c; }
int sadgpu_compute_region(sadgpu_ctx*, const uint8_t* l, int ls, const uint8_t*, int, int w, int h, int, int, int x0, int y0, int x1, int y1, uint8_t* out, int os)
{ for (int y = y0; y < y1; ++y) memcpy(out + (size_t)(y - y0) * os, l + (size_t)y * ls + x0, x1 - x0); return 0; }
int sadgpu_compute_nrgba(sadgpu_ctx*, int, const uint8_t*, int, const uint8_t*, int, int, int, int, int, uint8_t*, int) { return 0; }
int sadgpu_submit_batch_into(sadgpu_ctx*, int, int, const uint8_t*, int, int, int, int, uint8_t*, uint64_t*) { return 0; }
int sadgpu_submit_into(sadgpu_ctx*, int, const uint8_t*, int, const uint8_t*, int, int, int, int, int, int, int, uint8_t*, int, uint64_t*) { return 0; }
int sadgpu_wait(sadgpu_ctx*, uint64_t, uint8_t*, int) { return 0; }
void* sadgpu_host_alloc(sadgpu_ctx*, size_t n) { return malloc(n); }
void sadgpu_host_free(sadgpu_ctx*, void* p) { free(p); }
const char* sadgpu_strerror(int) { return "mock"; }
}
