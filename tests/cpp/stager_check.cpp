// tests/cpp/stager_check.cpp -- CPU-only check of the staging pool's concurrency contract (csrc/pcf_stager.hpp) with fake GPU hooks:
//   * clouds are handed over strictly in submission order, whatever the mix of packers and raw lanes and however long a pack takes;
//   * every cloud is handed over exactly once, with the packed payload of ITS OWN source (slots are not recycled too early);
//   * drop_queued() (= clouds_.clear(), node.cpp:356) only removes clouds no thread has taken; drain() returns when all are out.
// Prints "OK <n>" per scenario; exit code 1 on the first violation.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../high-fidelity-pointcloud-fusion_b200/csrc/pcf_stager.hpp"

using namespace pcf;

struct Cloud { std::vector<float> xyzw; double pose[16]; };

static int run(int threads, int raw_lanes, int n_clouds, bool pinned, int drop_after, unsigned seed) {
    std::mt19937 rng(seed);
    std::vector<Cloud> clouds(n_clouds);
    for (int i = 0; i < n_clouds; i++) {
        const int n = 64 + (int)(rng() % 4000);
        clouds[i].xyzw.resize((size_t)n * 4);
        for (int p = 0; p < n; p++) {
            float* q = &clouds[i].xyzw[(size_t)p * 4];
            q[0] = (float)i; q[1] = (float)p; q[2] = (rng() % 3) ? 0.4f : 0.9f; q[3] = 0.f;      // x = cloud id, y = point id; a third is clipped
        }
        for (int k = 0; k < 16; k++) clouds[i].pose[k] = i;
    }
    std::vector<uint32_t> order;                     // frame_idx in hand-over order (written under the pool's lock: pushes are serialised)
    std::atomic<int> bad{0};
    Stager::Hooks h;
    h.alloc_pinned = [](size_t b) { return malloc(b); };
    h.free_pinned = [](void* p) { free(p); };
    h.thread_init = [](int) {};
    h.slot_wait = [](int) {};
    h.raw_wait = [](int) {};
    h.raw_ok = [pinned](const StageJob& j) { return pinned && j.point_step == 16; };
    h.push = [&](int, const float* xyz, uint32_t n_staged, uint32_t n_offered, const double* pose, uint32_t frame_idx) {
        order.push_back(frame_idx);
        if (drop_after >= 0) std::this_thread::sleep_for(std::chrono::microseconds(100));   // a slow GPU hand-over: clouds pile up in the queue
        const Cloud& c = clouds[frame_idx];
        if (pose[5] != (double)frame_idx || n_offered != c.xyzw.size() / 4 || n_staged % 4) bad++;
        uint32_t k = 0;                               // the packed payload must be this cloud's kept points, in order
        for (size_t p = 0; p < c.xyzw.size() / 4; p++) {
            if (!(c.xyzw[p * 4 + 2] > 0.28f && c.xyzw[p * 4 + 2] < 0.6f)) continue;
            if (k >= n_staged || xyz[3 * k] != (float)frame_idx || xyz[3 * k + 1] != (float)p) { bad++; break; }
            k++;
        }
        if (n_staged - k > 3) bad++;
        return 0;
    };
    h.push_raw = [&](int, const StageJob& j) {
        order.push_back(j.frame_idx);
        if (j.data != reinterpret_cast<const uint8_t*>(clouds[j.frame_idx].xyzw.data())) bad++;
        return 0;
    };
    size_t dropped = 0;
    {
        Stager st(threads, raw_lanes, 0.28f, 0.6f, std::move(h));
        for (int i = 0; i < n_clouds; i++) {
            StageJob j;
            j.data = reinterpret_cast<const uint8_t*>(clouds[i].xyzw.data());
            j.rows = 1; j.cols = (uint32_t)(clouds[i].xyzw.size() / 4); j.point_step = 16; j.x_offset = 0;
            for (int k = 0; k < 16; k++) j.pose[k] = clouds[i].pose[k];
            j.frame_idx = (uint32_t)i;
            st.submit(j);
            if (i == drop_after) dropped = st.drop_queued();
        }
        if (st.drain() != 0) bad++;
        if (st.staged() + dropped != (uint64_t)n_clouds) bad++;
    }
    for (size_t i = 1; i < order.size(); i++) if (order[i] <= order[i - 1]) bad++;       // submission order, nothing twice
    if (order.size() + dropped != (size_t)n_clouds) bad++;
    if (drop_after < 0 && dropped != 0) bad++;
    if (bad) { printf("FAILED threads=%d lanes=%d pinned=%d: %d violations\n", threads, raw_lanes, (int)pinned, bad.load()); return 1; }
    printf("OK %zu handed over, %zu dropped (threads=%d lanes=%d pinned=%d)\n", order.size(), dropped, threads, raw_lanes, (int)pinned);
    return 0;
}

int main() {
    int rc = 0;
    rc |= run(1, 0, 200, false, -1, 1);
    rc |= run(4, 0, 400, false, -1, 2);
    rc |= run(7, 0, 300, false, 150, 3);
    rc |= run(3, 2, 400, true, -1, 4);
    rc |= run(1, 12, 400, true, -1, 5);          // upload mode
    rc |= run(2, 3, 300, true, 100, 6);
    rc |= run(2, 3, 300, false, -1, 7);          // raw lanes that have to pack (clouds not pinned)
    return rc;
}
