// host/pcf_replay.cpp -- offline replay driver: replaces `roslaunch pointcloud_fusion pointcloud_fusion_node.launch`
// + `rosbag play` + the start / stop / process service calls of the reference (launch:1-10, node.cpp:442-460).
//   pcf_replay <sequence.bin> --out <dir> [--update-every k] [--device d] [--slots n] [--reset-after f] [--quiet]
// Reads a PCFSEQ1 sequence (host/sequence.hpp), stages every cloud straight into pinned memory, and drives
// PointcloudFusion: start -> frames -> stop -> process.  Prints one JSON line with the measured rates.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "pointcloud_fusion.hpp"
#include "sequence.hpp"

int main(int argc, char** argv) {
    std::string seq, out = ".";
    pcfusion::PointcloudFusion::Params p;
    long reset_after = -1;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--out") out = next();
        else if (a == "--update-every") p.update_every = atoi(next());
        else if (a == "--device") p.device = atoi(next());
        else if (a == "--slots") p.staging_slots = (uint32_t)atoi(next());
        else if (a == "--reset-after") reset_after = atol(next());
        else if (a[0] != '-') seq = a;
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    if (seq.empty()) { fprintf(stderr, "usage: pcf_replay <sequence.bin> --out <dir> [--update-every k] [--device d] [--slots n]\n"); return 2; }
    FILE* f = fopen(seq.c_str(), "rb");
    pcfusion::SeqHeader h;
    if (!f || !pcfusion::read_header(f, h)) { fprintf(stderr, "pcf_replay: cannot read %s\n", seq.c_str()); return 2; }
    std::memcpy(p.box, h.box, sizeof p.box);
    p.res = h.res[0];
    p.clip_zmin = h.clip_zmin;
    p.clip_zmax = h.clip_zmax;
    p.directory_name = out;
    p.log_capacity_hint = (uint64_t)h.n_frames * h.points_per_frame;

    pcfusion::PointcloudFusion node(p);
    if (!node.ok()) { fprintf(stderr, "pcf_replay: %s\n", node.last_error().c_str()); return 3; }   // no CUDA device: no fallback
    node.start();
    const size_t floats = (size_t)h.points_per_frame * h.stride_floats;
    auto t0 = std::chrono::steady_clock::now();
    for (uint32_t i = 0; i < h.n_frames; i++) {
        float* slot = node.acquire(floats);
        double pose[16];
        if (!slot || !pcfusion::read_frame(f, h, pose, slot)) { fprintf(stderr, "pcf_replay: short read at frame %u\n", i); return 2; }
        node.submit(slot, h.points_per_frame, h.stride_floats, pose);
        if (reset_after >= 0 && (long)i == reset_after) node.reset();
    }
    fclose(f);
    node.stop();
    node.drain();
    double ingest_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    bool ok = node.getFusedCloud();
    auto c = node.counters();
    pcf_stats st;
    pcf_get_stats(node.handle(), &st);
    printf("{\"frames\": %llu, \"integrated\": %llu, \"dropped\": %llu, \"discarded_by_reset\": %llu, \"updates\": %llu, "
           "\"points\": %llu, \"ingest_s\": %.6f, \"points_per_s\": %.1f, \"process_ms\": %.3f, \"kernel_launches\": %llu, \"ok\": %s}\n",
           (unsigned long long)c.received, (unsigned long long)c.integrated, (unsigned long long)c.dropped,
           (unsigned long long)c.discarded_by_reset, (unsigned long long)c.updates,
           (unsigned long long)c.integrated * h.points_per_frame, ingest_s,
           ingest_s > 0 ? (double)c.integrated * h.points_per_frame / ingest_s : 0.0, node.last_process_ms(),
           (unsigned long long)st.kernel_launches, ok ? "true" : "false");
    return ok ? 0 : 4;
}
