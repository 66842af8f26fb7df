#include <cstdio>
#include <cstdint>
#include <vector>
extern "C" int despair_host_output_camera_loop(const uint8_t*, const uint8_t*, int, int, int, int, int, int, int, int, int, uint8_t*, double*);
int main(int argc, char** argv) {
    int w = 640, h = 480, workers = 32;
    if (argc > 1) { w = 1920; h = 1080; }
    std::vector<uint8_t> L((size_t)w * h * 2, 7), R(L), out((size_t)w * h);
    for (int rep = 0; rep < 3; ++rep) {
        double us = 0;
        int rc = despair_host_output_camera_loop(L.data(), R.data(), 2, w, h, 16, 64, workers, 20, 400, 0, out.data(), &us);
        printf("rc=%d %dx%d workers=%d: %.1f us/frame (host pipeline only, mock backend)\n", rc, w, h, workers, us);
    }
}
