// tools/pack_bench.cpp -- host-only throughput of the staging pool's clip-and-pack (csrc/pcf_stager.hpp) on C2-shaped clouds:
//   g++ -O2 -std=c++17 -pthread -I high-fidelity-pointcloud-fusion_b200/csrc tools/pack_bench.cpp -o /tmp/pack_bench && /tmp/pack_bench [threads] [frames]
// Prints input GB/s and points/s.  No GPU needed.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "pcf_stager.hpp"

int main(int argc, char** argv) {
    int threads = argc > 1 ? atoi(argv[1]) : 8, frames = argc > 2 ? atoi(argv[2]) : 64;
    const uint32_t n = 307200;
    std::vector<std::vector<float>> src(frames, std::vector<float>((size_t)n * 4));
    for (int f = 0; f < frames; f++)
        for (uint32_t i = 0; i < n; i++) {
            uint32_t u = i % 640, v = i / 640;
            bool hit = (u - 320.f) * (u - 320.f) + (v - 240.f) * (v - 240.f) < 212.f * 212.f;       // ~46 % of the image
            float* p = &src[f][(size_t)i * 4];
            p[0] = hit ? 0.001f * u : NAN; p[1] = hit ? 0.001f * v : NAN; p[2] = hit ? 0.3f + 1e-4f * (i % 1000) : NAN; p[3] = 0.f;
        }
    // every thread cycles through 8 destination slots (like the pinned slot ring: the packed clouds do not stay in the cache)
    const size_t slot_floats = ((size_t)n * 3 + 16 + 15) / 16 * 16;
    std::vector<float*> dst(threads);
    for (int t = 0; t < threads; t++) dst[t] = static_cast<float*>(aligned_alloc(4096, 8 * slot_floats * sizeof(float)));
    for (int rep = 0; rep < 3; rep++) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> pool;
        std::vector<uint64_t> kept(threads, 0);
        const int passes = 8;
        for (int t = 0; t < threads; t++)
            pool.emplace_back([&, t] {
                for (int pass = 0; pass < passes; pass++)
                    for (int f = t; f < frames; f += threads) {
                        pcf::StageJob j;
                        j.data = reinterpret_cast<const uint8_t*>(src[f].data());
                        j.rows = 1; j.cols = n; j.point_step = 16; j.x_offset = 0;
                        kept[t] += pcf::clip_pack(j, 0.28f, 0.6f, dst[t] + (size_t)((f / threads) & 7) * slot_floats);
                    }
            });
        for (auto& th : pool) th.join();
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        uint64_t k = 0;
        for (auto x : kept) k += x;
        double pts = (double)passes * frames * n;
        printf("threads %d: %.2f G points/s, %.1f GB/s read, kept %.3f\n", threads, pts / s / 1e9, pts * 16 / s / 1e9, k / pts);
    }
    return 0;
}
