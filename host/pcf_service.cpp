// host/pcf_service.cpp -- line-oriented service shell around PointcloudFusion: the offline stand-in for the node's four
// std_srvs/Trigger services + its subscriber (node.cpp:152-157, 327-440).  Commands on stdin (or any pipe / socket
// redirected to it, e.g. `nc -l 9000 | pcf_service seq.bin --out dir`):
//   start | stop | reset | process        the four services (node.cpp:351-440); `process` prints success=<0|1> (D8)
//   play <first> <count>                  publish `count` recorded clouds starting at frame `first` (dropped unless started)
//   drain                                 wait until every queued cloud is integrated
//   stats                                 one JSON line with the counters
//   quit
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>

#include "pointcloud_fusion.hpp"
#include "sequence.hpp"

int main(int argc, char** argv) {
    std::string seq, out = ".";
    pcfusion::PointcloudFusion::Params p;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--out") out = next();
        else if (a == "--update-every") p.update_every = atoi(next());
        else if (a == "--device") p.device = atoi(next());
        else if (a[0] != '-') seq = a;
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    FILE* f = seq.empty() ? nullptr : fopen(seq.c_str(), "rb");
    pcfusion::SeqHeader h;
    if (!f || !pcfusion::read_header(f, h)) { fprintf(stderr, "usage: pcf_service <sequence.bin> --out <dir> [--update-every k] [--device d]\n"); return 2; }
    std::memcpy(p.box, h.box, sizeof p.box);
    p.res = h.res[0];
    p.clip_zmin = h.clip_zmin;
    p.clip_zmax = h.clip_zmax;
    p.directory_name = out;
    pcfusion::PointcloudFusion node(p);
    if (!node.ok()) { fprintf(stderr, "pcf_service: %s\n", node.last_error().c_str()); return 3; }
    const size_t floats = (size_t)h.points_per_frame * h.stride_floats;
    const size_t frame_bytes = 16 * sizeof(double) + floats * sizeof(float);
    std::string line;
    while (std::getline(std::cin, line)) {
        std::istringstream is(line);
        std::string cmd;
        is >> cmd;
        if (cmd.empty() || cmd[0] == '#') continue;
        if (cmd == "start") node.start();
        else if (cmd == "stop") node.stop();
        else if (cmd == "reset") node.reset();
        else if (cmd == "drain") node.drain();
        else if (cmd == "process") {
            const bool ok = node.getFusedCloud();
            std::cout << "success=" << (ok ? 1 : 0) << std::endl;
        }
        else if (cmd == "play") {
            long first = 0, count = 0;
            is >> first >> count;
            for (long i = first; i < first + count && i < (long)h.n_frames; i++) {
                float* slot = node.acquire(floats);
                double pose[16];
                if (!slot || fseek(f, (long)(sizeof h + (size_t)i * frame_bytes), SEEK_SET) != 0 || !pcfusion::read_frame(f, h, pose, slot)) {
                    fprintf(stderr, "pcf_service: cannot read frame %ld\n", i);
                    return 2;
                }
                node.submit(slot, h.points_per_frame, h.stride_floats, pose);
            }
        } else if (cmd == "stats") {
            auto c = node.counters();
            std::cout << "{\"received\": " << c.received << ", \"dropped\": " << c.dropped << ", \"integrated\": " << c.integrated
                      << ", \"discarded_by_reset\": " << c.discarded_by_reset << ", \"updates\": " << c.updates << "}" << std::endl;
        } else if (cmd == "quit") break;
        else std::cerr << "unknown command: " << cmd << std::endl;
    }
    fclose(f);
    return 0;
}
