// host/sharded_fusion.cpp -- see sharded_fusion.hpp.  Plain C++17 over the C ABI.
#include "sharded_fusion.hpp"

#include <algorithm>
#include <chrono>
#include <cstring>

namespace pcfusion {

pcf_result ShardedResult::view() const {
    pcf_result r;
    r.n = hash.size();
    r.hash = hash.data(); r.centroid = centroid.data(); r.normal = normal.data(); r.sd = sd.data();
    r.mean_dist = mean_dist.data(); r.sd_dist = sd_dist.data(); r.count = count.data();
    return r;
}

void frame_block(uint32_t n_frames, uint32_t r, uint32_t world, uint32_t& lo, uint32_t& hi) {
    uint32_t base = n_frames / world, rem = n_frames % world;
    lo = r * base + std::min(r, rem);
    hi = lo + base + (r < rem ? 1u : 0u);
}

namespace {
// x-plane boundaries with ~equal numbers of log records per slab (sharded.py::slab_bounds_from_points)
std::vector<int32_t> slab_bounds(const std::vector<uint64_t>& plane_points, uint32_t world) {
    const size_t n_planes = plane_points.size();
    std::vector<uint64_t> cum(n_planes + 1, 0);
    for (size_t i = 0; i < n_planes; i++) cum[i + 1] = cum[i] + plane_points[i];
    std::vector<int32_t> b{0};
    for (uint32_t r = 1; r < world; r++) {
        uint64_t target = cum[n_planes] * r / world;
        size_t x = (size_t)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
        b.push_back((int32_t)std::min<size_t>(std::max<size_t>(x, (size_t)b.back()), n_planes));
    }
    b.push_back((int32_t)n_planes);
    return b;
}
template <class T>
void append(std::vector<T>& dst, const T* src, size_t n) { dst.insert(dst.end(), src, src + n); }
}  // namespace

bool merge_and_extract(const std::vector<pcf_ctx*>& ranks, const std::vector<uint32_t>& first_frame,
                       const std::vector<uint32_t>& n_frames, ShardedResult& out, std::string& err, double* exchange_ms, double* slab_ms) {
    const uint32_t R = (uint32_t)ranks.size();
    auto fail = [&](pcf_ctx* c, const char* what) { err = std::string(what) + ": " + pcf_last_error(c); return false; };
    auto t0 = std::chrono::steady_clock::now();
    int32_t dims[3];
    pcf_dims(ranks[0], dims);
    const size_t n_planes = (size_t)dims[0] + 1;
    // 1. plane histogram, summed over ranks -> slab bounds
    std::vector<uint64_t> plane(n_planes, 0);
    std::vector<uint32_t> tmp(n_planes);
    for (pcf_ctx* c : ranks) {
        if (pcf_sync(c) != PCF_OK || pcf_plane_point_counts(c, tmp.data()) != PCF_OK) return fail(c, "plane histogram");
        for (size_t i = 0; i < n_planes; i++) plane[i] += tmp[i];
    }
    const std::vector<int32_t> bounds = slab_bounds(plane, R);
    // 2. every rank learns every frame's viewpoint (disjoint rows)
    for (uint32_t s = 0; s < R; s++) {
        if (!n_frames[s]) continue;
        std::vector<float> rows((size_t)n_frames[s] * 4);
        if (pcf_get_viewpoints(ranks[s], rows.data(), first_frame[s], n_frames[s]) != PCF_OK) return fail(ranks[s], "viewpoints");
        for (uint32_t d = 0; d < R; d++)
            if (d != s && pcf_set_viewpoints(ranks[d], rows.data(), first_frame[s], n_frames[s]) != PCF_OK) return fail(ranks[d], "viewpoints");
    }
    // 3. count matrix -> offsets of (source, destination) blocks: source-rank order = frame order = arrival order
    std::vector<std::vector<uint64_t>> cnt(R, std::vector<uint64_t>(R, 0)), off(R, std::vector<uint64_t>(R, 0));
    for (uint32_t s = 0; s < R; s++)
        if (pcf_exchange_counts(ranks[s], bounds.data(), (int32_t)R, cnt[s].data()) != PCF_OK) return fail(ranks[s], "exchange counts");
    std::vector<uint64_t> total(R, 0);
    for (uint32_t d = 0; d < R; d++)
        for (uint32_t s = 0; s < R; s++) { off[s][d] = total[d]; total[d] += cnt[s][d]; }
    // 4. receive buffers; ONE kernel per rank compacts and stores into all of them
    std::vector<void*> bufs(R, nullptr);
    for (uint32_t d = 0; d < R; d++)
        if (pcf_recv_buffer(ranks[d], total[d], &bufs[d]) != PCF_OK) return fail(ranks[d], "receive buffer");
    for (uint32_t s = 0; s < R; s++)
        if (pcf_exchange_scatter(ranks[s], bufs.data(), off[s].data()) != PCF_OK) return fail(ranks[s], "exchange scatter");   // synchronises its stream
    auto t1 = std::chrono::steady_clock::now();
    // 5. install + slab work; concatenate in rank order
    out = ShardedResult();
    for (uint32_t r = 0; r < R; r++) {
        pcf_ctx* c = ranks[r];
        pcf_result res;
        if (pcf_install_records(c, bufs[r], total[r]) != PCF_OK || pcf_set_slab(c, bounds[r], bounds[r + 1]) != PCF_OK ||
            pcf_update(c) != PCF_OK || pcf_extract(c, &res) != PCF_OK)
            return fail(c, "slab process");
        const size_t n = (size_t)res.n;
        append(out.hash, res.hash, n);
        append(out.centroid, res.centroid, 3 * n); append(out.normal, res.normal, 3 * n); append(out.sd, res.sd, 3 * n);
        append(out.mean_dist, res.mean_dist, n); append(out.sd_dist, res.sd_dist, n); append(out.count, res.count, n);
    }
    auto t2 = std::chrono::steady_clock::now();
    if (exchange_ms) *exchange_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (slab_ms) *slab_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
    return true;
}

}  // namespace pcfusion
