#include "../../../steroscopic-hardware_b200/host/despair.hpp"
#include <cstdio>
#include <thread>
using namespace despair;
using clk = std::chrono::steady_clock;
static double us(clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); }
int main() {
    for (int idle : {0, 30, 300}) {
        Chan<OutputChunk> c(64); Chan<int> back(4);
        std::atomic<long long> t_first{0}, t_last{0};
        std::thread cons([&] { OutputChunk o; int n = 0; while (c.Recv(o)) { if (n % 160 == 0) t_first = clk::now().time_since_epoch().count(); if (++n % 160 == 0) { t_last = clk::now().time_since_epoch().count(); back.Send(1); } } });
        double a = 0, b = 0, d = 0, e = 0; int reps = 300;
        for (int rep = 0; rep < reps; ++rep) {
            if (idle) std::this_thread::sleep_for(std::chrono::microseconds(idle));
            auto t0 = clk::now();
            for (int i = 0; i < 160; ++i) { OutputChunk o; o.DisparityData.assign(1920, 0); c.Send(std::move(o)); }
            auto t1 = clk::now();
            int x; back.Recv(x);
            auto t2 = clk::now();
            a += us(t0, t1); b += us(t0, t2);
            d += (t_first.load() - t0.time_since_epoch().count()) / 1e3; e += (t_last.load() - t0.time_since_epoch().count()) / 1e3;
        }
        c.Close(); cons.join();
        printf("idle %3d us: sends done %.1f us, first recv at %.1f, last recv at %.1f, ack seen at %.1f\n", idle, a / reps, d / reps, e / reps, b / reps);
    }
    auto t0 = clk::now(); for (int i = 0; i < 100000; ++i) (void)clk::now(); printf("steady_clock::now(): %.3f us\n", us(t0, clk::now()) / 100000);
}
