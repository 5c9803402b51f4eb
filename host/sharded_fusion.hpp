// host/sharded_fusion.hpp -- frame-sharded fusion driven from ONE C++ process: N contexts (one per GPU, or several on
// one GPU), frames in contiguous frame_idx blocks per context, no communication while frames are integrated, and at
// process() the "exchange v2" of include/pcfusion.h: every context's ONE compaction+scatter kernel stores its records
// straight into the other contexts' receive buffers (peer access over NVLink), then every context installs, updates and
// extracts its x-slab; the slab results concatenated in context order are the reference's x-major scan (OG.hpp:463-465),
// byte-identical to a single context fed every frame.  No NCCL, no MPI: the three tiny "collectives" (plane histogram,
// count matrix, viewpoint rows) are host-side sums because all contexts live in this process.
// Python counterpart for one process per GPU: high-fidelity-pointcloud-fusion_b200/sharded.py::merge_and_extract_v2.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../include/pcfusion.h"

namespace pcfusion {

struct ShardedResult {            // concatenated extraction, x-major (owned by the object that returned it)
    std::vector<uint64_t> hash;
    std::vector<float> centroid, normal, sd, mean_dist, sd_dist;
    std::vector<int32_t> count;
    pcf_result view() const;
};

// [lo, hi) of `n_frames` owned by rank `r` of `world` (same rule as sharded.py::frame_block)
void frame_block(uint32_t n_frames, uint32_t r, uint32_t world, uint32_t& lo, uint32_t& hi);

// ranks[r] has integrated the frames [first_frame[r], first_frame[r] + n_frames[r]) (global indices, contiguous blocks in
// rank order).  Runs exchange v2 + update + extract on every rank and returns the merged result.  false + err on failure.
bool merge_and_extract(const std::vector<pcf_ctx*>& ranks, const std::vector<uint32_t>& first_frame,
                       const std::vector<uint32_t>& n_frames, ShardedResult& out, std::string& err, double* exchange_ms = nullptr,
                       double* slab_ms = nullptr);

}  // namespace pcfusion
