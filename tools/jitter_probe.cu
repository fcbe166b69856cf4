// Where do the slow steps come from?  Two probes run side by side for a few seconds:
//   (1) launch + stream-synchronize round trips of an empty kernel (host -> GPU -> host path: doorbell, command fetch, completion write);
//   (2) a busy loop on another host thread that looks for gaps in its own clock readings (a gap = the vCPU was taken away).
// Prints, per 100 ms window, the round-trip count / mean / max and the largest clock gap.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 tools/jitter_probe.cu -o tools/jitter_probe
#include <cstdio>
#include <chrono>
#include <thread>
#include <atomic>
#include <vector>
#include <cuda_runtime.h>
__global__ void k_empty() {}
using clk = std::chrono::steady_clock;
static double now_ms() { return std::chrono::duration<double, std::milli>(clk::now().time_since_epoch()).count(); }
int main(int argc, char **argv) {
    const double secs = argc > 1 ? atof(argv[1]) : 6.0; const int spin = argc > 2 ? atoi(argv[2]) : 0;
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    for (int i = 0; i < 100; i++) { k_empty<<<1, 32, 0, s>>>(); cudaStreamSynchronize(s); }
    const int W = (int)(secs * 10);
    std::vector<double> gap(W, 0.0); std::atomic<bool> stop{false};
    const double t0 = now_ms();
    std::thread spinner([&] { double last = now_ms(); while (!stop.load(std::memory_order_relaxed)) { double t = now_ms(); int w = (int)((t - t0) / 100.0); if (w >= 0 && w < W && t - last > gap[w]) gap[w] = t - last; last = t; } });
    std::vector<int> cnt(W, 0); std::vector<double> sum(W, 0.0), mx(W, 0.0);
    for (;;) {
        double a = now_ms(); int w = (int)((a - t0) / 100.0); if (w >= W) break;
        k_empty<<<1, 32, 0, s>>>();
        if (spin) { while (cudaStreamQuery(s) == cudaErrorNotReady) {} } else cudaStreamSynchronize(s);
        double d = now_ms() - a; cnt[w]++; sum[w] += d; if (d > mx[w]) mx[w] = d;
    }
    stop = true; spinner.join();
    printf("window(100ms): roundtrips mean_us max_us | max clock gap of the busy thread (us)\n");
    double worst = 0, worstgap = 0; int slow = 0;
    for (int w = 0; w < W; w++) {
        double mean = cnt[w] ? sum[w] / cnt[w] * 1e3 : 0;
        if (mx[w] * 1e3 > 300 || gap[w] * 1e3 > 300 || mean > 20.0) { printf("%3d: %6d %8.1f %9.1f | %9.1f\n", w, cnt[w], mean, mx[w] * 1e3, gap[w] * 1e3); slow++; }
        if (mx[w] > worst) worst = mx[w]; if (gap[w] > worstgap) worstgap = gap[w];
    }
    long total = 0; for (int c : cnt) total += c;
    printf("summary: %ld round trips in %.1f s (%s), windows with a >300 us event: %d/%d, worst round trip %.1f us, worst clock gap %.1f us\n", total, secs, spin ? "polling" : "blocking", slow, W, worst * 1e3, worstgap * 1e3);
    return 0;
}
