// host/sequence.hpp -- on-disk format of a recorded replay sequence ("PCFSEQ1"): what a rosbag of
// `input_point_cloud` messages + tf lookups (node.cpp:327-349) boils down to for the offline replay driver.
//   header (SeqHeader, 104 bytes), then per frame: double pose[16] (row-major fusion<-camera) + float pts[n * stride]
// Written by high-fidelity-pointcloud-fusion_b200/synth.py::write_sequence, read by host/pcf_replay and the tests.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>

namespace pcfusion {

struct SeqHeader {
    char magic[8];              // "PCFSEQ1\0"
    uint32_t n_frames;
    uint32_t points_per_frame;  // width * height of the organized cloud
    uint32_t stride_floats;     // floats per point (4 = x y z pad)
    uint32_t reserved;
    double box[6];              // xmin xmax ymin ymax zmin zmax (launch:8 order)
    float res[3];
    float pad;
    double clip_zmin, clip_zmax;
};
static_assert(sizeof(SeqHeader) == 104, "SeqHeader layout");

inline bool read_header(FILE* f, SeqHeader& h) {
    return fread(&h, sizeof h, 1, f) == 1 && std::memcmp(h.magic, "PCFSEQ1", 8) == 0 && h.stride_floats >= 3;
}
// reads one frame into caller memory (pts must hold points_per_frame * stride_floats floats)
inline bool read_frame(FILE* f, const SeqHeader& h, double pose[16], float* pts) {
    size_t n = (size_t)h.points_per_frame * h.stride_floats;
    return fread(pose, sizeof(double), 16, f) == 16 && fread(pts, sizeof(float), n, f) == n;
}

}  // namespace pcfusion
