// tests/cpp/dropin_driver.cpp -- ONE driver source, compiled twice:
//   -DUSE_REFERENCE_GRID : against the reference's own utilities/OccupancyGrid.hpp (-> oracle/_ref/dropin_ref)
//   (default)            : against include/pcfusion/OccupancyGrid.hpp + libpcfusion.so (-> tests/_build/dropin_b200)
// It does what PointcloudFusion's three threads do for every message (node.cpp:218-325), in the canonical schedule
// (D4): z clip (node.cpp:251-255) -> transformPointCloud (node.cpp:289) -> grid.addPoints<6>(cloud, viewpoint)
// (node.cpp:290-295) ... updateThicknessVectors<6,3>() every `update_every` frames and once at the end
// (node.cpp:311) -> downloadData(dir/test_cloud.pcd, dir/meta.csv) (node.cpp:395-398) -> clearVoxels() (node.cpp:438).
// The drop-in claim is that both builds write byte-identical files; tests/test_dropin_gpu.py checks exactly that.
// TEST INFRASTRUCTURE: the PCL / Eigen types come from the stand-in headers under oracle/ref_shim/.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#ifdef USE_REFERENCE_GRID
#include <cstddef>
#include <new>
// D1 pin applied from outside the reference source: zero-filling operator new (VoxelInfo::mean_dist starts at 0)
void* operator new(std::size_t n) { void* p = std::calloc(1, n ? n : 1); if (!p) throw std::bad_alloc(); return p; }
void* operator new[](std::size_t n) { void* p = std::calloc(1, n ? n : 1); if (!p) throw std::bad_alloc(); return p; }
void operator delete(void* p) noexcept { std::free(p); }
void operator delete[](void* p) noexcept { std::free(p); }
void operator delete(void* p, std::size_t) noexcept { std::free(p); }
void operator delete[](void* p, std::size_t) noexcept { std::free(p); }
#include "utilities/OccupancyGrid.hpp"
#else
#include <Eigen/Core>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include "pcfusion/OccupancyGrid.hpp"
using pcfusion::OccupancyGrid;
#endif
#include "../../host/sequence.hpp"

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s sequence.bin out_dir [update_every]\n", argv[0]); return 2; }
    const std::string dir = argv[2];
    const int update_every = argc > 3 ? atoi(argv[3]) : 0;
    FILE* f = fopen(argv[1], "rb");
    pcfusion::SeqHeader h;
    if (!f || !pcfusion::read_header(f, h)) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }

    OccupancyGrid grid;
    grid.setResolution(h.res[0], h.res[1], h.res[2]);                                      // node.cpp:161
    grid.setDimensions(h.box[0], h.box[1], h.box[2], h.box[3], h.box[4], h.box[5]);        // node.cpp:162
    grid.setK(2);                                                                          // node.cpp:163
    if (!grid.construct()) {                                                               // node.cpp:164
#ifndef USE_REFERENCE_GRID
        fprintf(stderr, "construct failed: %s\n", grid.last_error().c_str());
#endif
        return 3;
    }
    std::vector<float> pts((size_t)h.points_per_frame * h.stride_floats);
    double T[16];
    for (uint32_t i = 0; i < h.n_frames; i++) {
        if (!pcfusion::read_frame(f, h, T, pts.data())) { fprintf(stderr, "short read at frame %u\n", i); return 2; }
        pcl::PointCloud<pcl::PointXYZRGB>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZRGB>);
        for (uint32_t p = 0; p < h.points_per_frame; p++) {
            const float* q = &pts[(size_t)p * h.stride_floats];
            if (q[2] < h.clip_zmax && q[2] > h.clip_zmin && std::isfinite(q[0]) && std::isfinite(q[1])) {   // node.cpp:251 (+D11)
                pcl::PointXYZRGB w;                                                        // pcl::transformPointCloud, node.cpp:289
                const double dx = q[0], dy = q[1], dz = q[2];
                w.x = static_cast<float>(T[0] * dx + T[1] * dy + T[2] * dz + T[3]);
                w.y = static_cast<float>(T[4] * dx + T[5] * dy + T[6] * dz + T[7]);
                w.z = static_cast<float>(T[8] * dx + T[9] * dy + T[10] * dz + T[11]);
                cloud->points.push_back(w);
            }
        }
        Eigen::Vector3f vp(T[3], T[7], T[11]);                                             // node.cpp:290
        if (!grid.addPoints<6>(cloud, vp)) return 4;                                       // node.cpp:292-295
        if (update_every > 0 && (i + 1) % update_every == 0) grid.updateThicknessVectors<6, 3>();   // node.cpp:311
    }
    fclose(f);
    grid.updateThicknessVectors<6, 3>();
    grid.downloadData(dir + "/test_cloud.pcd", dir + "/meta.csv");                         // node.cpp:395-398
    // the download variants the node keeps behind `#if 0` (node.cpp:399-437): dumped with their exact float bits
    {
        pcl::PointCloud<pcl::PointXYZRGB>::Ptr hq(new pcl::PointCloud<pcl::PointXYZRGB>), cls(new pcl::PointCloud<pcl::PointXYZRGB>),
            plain(new pcl::PointCloud<pcl::PointXYZRGB>);
        pcl::PointCloud<pcl::PointXYZRGBNormal>::Ptr nrm(new pcl::PointCloud<pcl::PointXYZRGBNormal>);
        grid.downloadHQ(hq, 3.0);                                                          // OG.hpp:545-575
        grid.downloadClassified(cls);                                                      // OG.hpp:514-543
        grid.download(plain);                                                              // OG.hpp:491-512
        grid.download(nrm);                                                                // OG.hpp:577-601
        FILE* o = fopen((dir + "/variants.txt").c_str(), "w");
        auto bits = [](float v) { uint32_t u; memcpy(&u, &v, 4); return u; };
        fprintf(o, "hq %zu %u %u\n", hq->points.size(), hq->width, hq->height);
        for (auto& p : hq->points) fprintf(o, "%08x %08x %08x %d %d %d\n", bits(p.x), bits(p.y), bits(p.z), p.r, p.g, p.b);
        fprintf(o, "classified %zu\n", cls->points.size());
        for (auto& p : cls->points) fprintf(o, "%08x %08x %08x %d %d %d\n", bits(p.x), bits(p.y), bits(p.z), p.r, p.g, p.b);
        fprintf(o, "plain %zu\n", plain->points.size());
        for (auto& p : plain->points) fprintf(o, "%08x %08x %08x\n", bits(p.x), bits(p.y), bits(p.z));
        fprintf(o, "normals %zu\n", nrm->points.size());
        for (auto& p : nrm->points) fprintf(o, "%08x %08x %08x %08x %08x %08x\n", bits(p.x), bits(p.y), bits(p.z), bits(p.normal[0]), bits(p.normal[1]), bits(p.normal[2]));
        fclose(o);
    }
    grid.clearVoxels();                                                                    // node.cpp:438
    return 0;
}
