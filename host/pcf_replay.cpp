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

#include <vector>

#include "pointcloud_fusion.hpp"
#include "sequence.hpp"
#include "sharded_fusion.hpp"

// --gpus N [--devices a,b,..]: frames sharded in contiguous blocks over N contexts of this process (host/sharded_fusion.hpp)
static int replay_sharded(const std::string& seq, const std::string& out, const std::vector<int>& devices) {
    FILE* f = fopen(seq.c_str(), "rb");
    pcfusion::SeqHeader h;
    if (!f || !pcfusion::read_header(f, h)) { fprintf(stderr, "pcf_replay: cannot read %s\n", seq.c_str()); return 2; }
    const uint32_t R = (uint32_t)devices.size();
    std::vector<pcf_ctx*> ranks(R, nullptr);
    std::vector<uint32_t> first(R), count(R);
    pcf_config cfg;
    pcf_default_config(&cfg);
    std::memcpy(cfg.box, h.box, sizeof cfg.box);
    std::memcpy(cfg.res, h.res, sizeof cfg.res);
    cfg.clip_zmin = h.clip_zmin;
    cfg.clip_zmax = h.clip_zmax;
    for (uint32_t r = 0; r < R; r++) {
        uint32_t lo, hi;
        pcfusion::frame_block(h.n_frames, r, R, lo, hi);
        first[r] = lo; count[r] = hi - lo;
        cfg.device = devices[r];
        cfg.log_capacity_hint = (uint64_t)(hi - lo + 1) * h.points_per_frame;
        if (pcf_create(&cfg, &ranks[r]) != PCF_OK) { fprintf(stderr, "pcf_replay: %s\n", pcf_last_error(nullptr)); return 3; }
        pcf_start(ranks[r]);
    }
    for (uint32_t a = 0; a < R; a++)
        for (uint32_t b = 0; b < R; b++)
            if (devices[a] != devices[b] && pcf_enable_peer_access(ranks[a], devices[b]) != PCF_OK) { fprintf(stderr, "pcf_replay: %s\n", pcf_last_error(ranks[a])); return 3; }
    const size_t floats = (size_t)h.points_per_frame * h.stride_floats;
    float* stage[2] = {static_cast<float*>(pcf_host_alloc(floats * 4)), static_cast<float*>(pcf_host_alloc(floats * 4))};
    uint64_t ticket[2] = {0, 0};
    pcf_ctx* owner[2] = {nullptr, nullptr};
    auto t0 = std::chrono::steady_clock::now();
    for (uint32_t i = 0, r = 0; i < h.n_frames; i++) {
        while (i >= first[r] + count[r]) r++;
        const int s = (int)(i & 1);
        if (owner[s]) pcf_wait_upload(owner[s], ticket[s]);          // the copy that last read this staging buffer is done
        double pose[16];
        if (!pcfusion::read_frame(f, h, pose, stage[s])) { fprintf(stderr, "pcf_replay: short read at frame %u\n", i); return 2; }
        if (pcf_push_frame(ranks[r], stage[s], h.points_per_frame, h.stride_floats, pose, i) < 0) { fprintf(stderr, "pcf_replay: %s\n", pcf_last_error(ranks[r])); return 4; }
        owner[s] = ranks[r];
        pcf_upload_ticket(ranks[r], &ticket[s]);
    }
    fclose(f);
    for (pcf_ctx* c : ranks) pcf_sync(c);
    double ingest_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    pcfusion::ShardedResult res;
    std::string err;
    double ex_ms = 0, slab_ms = 0;
    bool ok = pcfusion::merge_and_extract(ranks, first, count, res, err, &ex_ms, &slab_ms);
    if (!ok) fprintf(stderr, "pcf_replay: %s\n", err.c_str());
    if (ok) {
        pcf_result v = res.view();
        ok = pcf_write_result(&v, (out + "/test_cloud.pcd").c_str(), (out + "/meta.csv").c_str()) == PCF_OK;
    }
    uint64_t launches = 0;
    for (pcf_ctx* c : ranks) { pcf_stats st; pcf_get_stats(c, &st); launches += st.kernel_launches; pcf_clear(c); }
    printf("{\"gpus\": %u, \"frames\": %u, \"points\": %llu, \"ingest_s\": %.6f, \"points_per_s\": %.1f, \"exchange_ms\": %.3f, "
           "\"slab_process_ms\": %.3f, \"voxels\": %zu, \"kernel_launches\": %llu, \"ok\": %s}\n",
           R, h.n_frames, (unsigned long long)h.n_frames * h.points_per_frame, ingest_s,
           ingest_s > 0 ? (double)h.n_frames * h.points_per_frame / ingest_s : 0.0, ex_ms, slab_ms, res.hash.size(),
           (unsigned long long)launches, ok ? "true" : "false");
    pcf_host_free(stage[0]); pcf_host_free(stage[1]);
    for (pcf_ctx* c : ranks) pcf_destroy(c);
    return ok ? 0 : 4;
}

int main(int argc, char** argv) {
    std::string seq, out = ".";
    pcfusion::PointcloudFusion::Params p;
    long reset_after = -1;
    int gpus = 1;
    std::vector<int> devices;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--out") out = next();
        else if (a == "--update-every") p.update_every = atoi(next());
        else if (a == "--device") p.device = atoi(next());
        else if (a == "--slots") p.staging_slots = (uint32_t)atoi(next());
        else if (a == "--reset-after") reset_after = atol(next());
        else if (a == "--gpus") gpus = atoi(next());
        else if (a == "--devices") { for (const char* p = next(); *p;) { devices.push_back(atoi(p)); while (*p && *p != ',') p++; if (*p) p++; } }
        else if (a[0] != '-') seq = a;
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    if (seq.empty()) { fprintf(stderr, "usage: pcf_replay <sequence.bin> --out <dir> [--update-every k] [--device d] [--slots n] [--gpus N [--devices a,b,..]]\n"); return 2; }
    if (gpus > 1 || devices.size() > 1) {
        if (devices.empty()) for (int d = 0; d < gpus; d++) devices.push_back(d);
        return replay_sharded(seq, out, devices);
    }
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
