// tests/cpp/helpers_check.cpp -- CPU only.  ONE source compiled twice (reference header / drop-in header): prints the results
// of the grid's coordinate helpers (OG.hpp:131-135,151-165,630-650) for seeded random and adversarial inputs, one line each,
// with exact float bits.  tests/test_host_cpp.py compares the two outputs byte for byte.  No device is touched: the drop-in
// build never calls construct().
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#ifdef USE_REFERENCE_GRID
#include "utilities/OccupancyGrid.hpp"
#else
#include <Eigen/Core>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include "pcfusion/OccupancyGrid.hpp"
using pcfusion::OccupancyGrid;
#endif

static uint32_t bits(float v) { uint32_t u; memcpy(&u, &v, 4); return u; }

int main() {
    const double boxes[3][6] = {{-0.25, 0.25, -0.25, 0.25, -0.25, 0.25}, {-0.8, 1.8, -1.5, 1.5, 0.0, 1.0}, {-0.21, 0.24, -0.2, 0.26, 0.0, 0.23}};
    const float res[3][3] = {{0.001f, 0.001f, 0.001f}, {0.005f, 0.005f, 0.005f}, {0.005f, 0.004f, 0.006f}};
    std::mt19937_64 rng(99);
    for (int c = 0; c < 3; c++) {
        OccupancyGrid g;
        g.setResolution(res[c][0], res[c][1], res[c][2]);
        g.setDimensions(boxes[c][0], boxes[c][1], boxes[c][2], boxes[c][3], boxes[c][4], boxes[c][5]);
#ifdef USE_REFERENCE_GRID
        g.xdim_ = (int)((g.xmax_ - g.xmin_) / g.xres_); g.ydim_ = (int)((g.ymax_ - g.ymin_) / g.yres_); g.zdim_ = (int)((g.zmax_ - g.zmin_) / g.zres_);
#else
        g.computeDims();
#endif
        printf("dims %d %d %d\n", g.xdim_, g.ydim_, g.zdim_);
        std::uniform_real_distribution<double> U(-0.05, 1.05);
        for (int i = 0; i < 20000; i++) {
            double t[3] = {U(rng), U(rng), U(rng)};
            Eigen::Vector3f p((float)(boxes[c][0] + t[0] * (boxes[c][1] - boxes[c][0])), (float)(boxes[c][2] + t[1] * (boxes[c][3] - boxes[c][2])),
                              (float)(boxes[c][4] + t[2] * (boxes[c][5] - boxes[c][4])));
            if (i % 5 == 0) {     // exactly on a cell border / box face
                int k = (int)(rng() % 600);
                p = Eigen::Vector3f((float)(boxes[c][0] + (double)res[c][0] * k), p(1), (float)boxes[c][5]);
            }
            bool valid = g.validPoints(p);
            int x = 0, y = 0, z = 0;
            if (valid) {
                auto ijk = g.getVoxelCoords(p);
                x = std::get<0>(ijk); y = std::get<1>(ijk); z = std::get<2>(ijk);
            }
            unsigned long long h = g.getHashId(x, y, z);
            auto back = g.getVoxelCoords(h);
#ifdef USE_REFERENCE_GRID
            Eigen::Vector3f ctr = g.getVoxelCenter(x, y, z);
#else
            Eigen::Vector3f ctr = g.getVoxelCenter<Eigen::Vector3f>(x, y, z);
#endif
            printf("%d %d %d %d %llu %d %d %d %d %08x %08x %08x\n", (int)valid, x, y, z, h, std::get<0>(back), std::get<1>(back), std::get<2>(back),
                   (int)g.validCoord(x, y, z), bits(ctr(0)), bits(ctr(1)), bits(ctr(2)));
        }
    }
    return 0;
}
