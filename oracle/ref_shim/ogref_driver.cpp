// oracle/ref_shim/ogref_driver.cpp -- TEST INFRASTRUCTURE.
// C entry points around the reference's UNMODIFIED OccupancyGrid.hpp, which is compiled from where it lies
// (/root/reference/pointcloud_fusion/pointcloud_fusion/include/utilities/OccupancyGrid.hpp) against the
// shim Eigen/PCL headers in this directory.  Output goes to oracle/_ref/ (git-ignored).  The node-side
// stages that need ROS (z clip node.cpp:248-255, transformPointCloud node.cpp:288-290) are restated here.
//
// Pins applied from OUTSIDE the reference source:
//   D1  global operator new zero-fills, so VoxelInfo::mean_dist starts at 0
//   D2  built with -fsanitize=unreachable -fno-sanitize=return so the reference's value-returning
//       functions that fall off their end simply return (GCC would otherwise mark them unreachable)
//   D3  -DOGREF_ORDERED_KEYS swaps the two work lists for an ascending-order set (second build variant)
#include <cstdlib>
#include <new>
#ifndef OGREF_STOCK_NEW
#define OGREF_LOCAL __attribute__((visibility("hidden")))
OGREF_LOCAL void* operator new(std::size_t n) { void* p = std::calloc(1, n ? n : 1); if (!p) throw std::bad_alloc(); return p; }
OGREF_LOCAL void* operator new[](std::size_t n) { void* p = std::calloc(1, n ? n : 1); if (!p) throw std::bad_alloc(); return p; }
OGREF_LOCAL void operator delete(void* p) noexcept { std::free(p); }
OGREF_LOCAL void operator delete[](void* p) noexcept { std::free(p); }
OGREF_LOCAL void operator delete(void* p, std::size_t) noexcept { std::free(p); }
OGREF_LOCAL void operator delete[](void* p, std::size_t) noexcept { std::free(p); }
#endif  // OGREF_STOCK_NEW (timing build: the reference's own malloc behaviour, D1 unpinned)

#include <cstdint>
#include "utilities/OccupancyGrid.hpp"

namespace {
struct Ref {
    OccupancyGrid grid;
    double clip_zmin, clip_zmax;
    pcl::PointCloud<pcl::PointXYZRGBNormal>::Ptr out;
    std::vector<uint64_t> hash;
    std::vector<float> sd, mean_dist, sd_dist;
    std::vector<int32_t> count;
};
}  // namespace

extern "C" {
void* ref_create(const double* box, const float* res, double clip_zmin, double clip_zmax, int /*reserve*/) {
    Ref* r = new Ref();
    r->clip_zmin = clip_zmin; r->clip_zmax = clip_zmax;
    r->grid.setResolution(res[0], res[1], res[2]);                              // node.cpp:161
    r->grid.setDimensions(box[0], box[1], box[2], box[3], box[4], box[5]);      // node.cpp:162
    r->grid.setK(2);                                                           // node.cpp:163
    r->grid.construct();                                                       // node.cpp:164
    return r;
}
void ref_destroy(void* h) { delete (Ref*)h; }   // leaks VoxelInfo like the reference does
void ref_dims(void* h, int* d) { Ref* r = (Ref*)h; d[0] = r->grid.xdim_; d[1] = r->grid.ydim_; d[2] = r->grid.zdim_; }

int64_t ref_add_points_world(void* h, const float* xyz, int stride, int64_t n, const float* vp) {
    Ref* r = (Ref*)h;
    pcl::PointCloud<pcl::PointXYZRGB>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZRGB>);
    cloud->points.resize((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        cloud->points[i].x = xyz[i * stride]; cloud->points[i].y = xyz[i * stride + 1]; cloud->points[i].z = xyz[i * stride + 2];
    }
    r->grid.addPoints<6>(cloud, Eigen::Vector3f(vp[0], vp[1], vp[2]));          // node.cpp:292-295
    return n;
}
int64_t ref_add_frame(void* h, const float* pts, int stride, int64_t n, const double* T) {
    Ref* r = (Ref*)h;
    pcl::PointCloud<pcl::PointXYZRGB>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZRGB>);
    for (int64_t i = 0; i < n; i++) {
        float x = pts[i * stride], y = pts[i * stride + 1], z = pts[i * stride + 2];
        if (z < r->clip_zmax && z > r->clip_zmin) {                             // node.cpp:251-255
            // D11: the reference would index the grid with garbage for a non-finite x/y; callers of this
            // driver never pass such points (the oracle drops them), so nothing to do here.
            pcl::PointXYZRGB q;
            double dx = x, dy = y, dz = z;                                      // PCL transformPointCloud, A.3
            q.x = static_cast<float>(T[0] * dx + T[1] * dy + T[2] * dz + T[3]);
            q.y = static_cast<float>(T[4] * dx + T[5] * dy + T[6] * dz + T[7]);
            q.z = static_cast<float>(T[8] * dx + T[9] * dy + T[10] * dz + T[11]);
            cloud->points.push_back(q);
        }
    }
    Eigen::Vector3f vp(T[3], T[7], T[11]);                                      // node.cpp:290
    r->grid.addPoints<6>(cloud, vp);
    return (int64_t)cloud->points.size();
}
void ref_update(void* h) { ((Ref*)h)->grid.updateThicknessVectors<6, 3>(); }    // node.cpp:311

int64_t ref_download(void* h) {
    Ref* r = (Ref*)h;
    r->out.reset(new pcl::PointCloud<pcl::PointXYZRGBNormal>);
    r->grid.download(r->out);                                                   // OG.hpp:577-601 (same scan as downloadData)
    r->hash.clear(); r->sd.clear(); r->mean_dist.clear(); r->sd_dist.clear(); r->count.clear();
    OccupancyGrid& g = r->grid;
    for (int x = 0; x < g.xdim_; x++)
        for (int y = 0; y < g.ydim_; y++)
            for (int z = 0; z < g.zdim_; z++)
                if (g.voxels_[x][y][z].occupied) {
                    VoxelInfo* d = reinterpret_cast<VoxelInfo*>(g.voxels_[x][y][z].data);
                    if (!d->normal_found) continue;
                    r->hash.push_back(((uint64_t)x << 40) ^ ((uint64_t)y << 20) ^ (uint64_t)z);
                    r->sd.push_back(d->sd(0)); r->sd.push_back(d->sd(1)); r->sd.push_back(d->sd(2));
                    r->mean_dist.push_back(d->mean_dist); r->sd_dist.push_back(d->sd_dist); r->count.push_back(d->count);
                }
    return (int64_t)r->out->points.size();
}
void ref_get_result(void* h, uint64_t* hash, float* centroid, float* normal, float* sd, float* mean_dist, float* sd_dist, int32_t* count) {
    Ref* r = (Ref*)h;
    size_t n = r->hash.size();
    for (size_t i = 0; i < n; i++) {
        const pcl::PointXYZRGBNormal& p = r->out->points[i];
        if (hash) hash[i] = r->hash[i];
        if (centroid) { centroid[3 * i] = p.x; centroid[3 * i + 1] = p.y; centroid[3 * i + 2] = p.z; }
        if (normal) { normal[3 * i] = p.normal[0]; normal[3 * i + 1] = p.normal[1]; normal[3 * i + 2] = p.normal[2]; }
        if (sd) { sd[3 * i] = r->sd[3 * i]; sd[3 * i + 1] = r->sd[3 * i + 1]; sd[3 * i + 2] = r->sd[3 * i + 2]; }
        if (mean_dist) mean_dist[i] = r->mean_dist[i];
        if (sd_dist) sd_dist[i] = r->sd_dist[i];
        if (count) count[i] = r->count[i];
    }
}
int ref_download_data(void* h, const char* cloud_path, const char* meta_path) {  // OG.hpp:456-488
    return ((Ref*)h)->grid.downloadData(cloud_path, meta_path) ? 0 : 0;
}
int64_t ref_state_size(void* h) {
    OccupancyGrid& g = ((Ref*)h)->grid;
    int64_t n = 0;
    for (auto& a : g.voxels_) for (auto& b : a) for (auto& v : b) n += v.occupied ? 1 : 0;
    return n;
}
void ref_get_state(void* h, uint64_t* hash, int32_t* buffer_len, uint8_t* normal_found, int32_t* count, float* normal, float* viewpoint) {
    OccupancyGrid& g = ((Ref*)h)->grid;
    int64_t n = 0;
    for (int x = 0; x <= g.xdim_; x++)
        for (int y = 0; y <= g.ydim_; y++)
            for (int z = 0; z <= g.zdim_; z++) {
                if (!g.voxels_[x][y][z].occupied) continue;
                VoxelInfo* d = reinterpret_cast<VoxelInfo*>(g.voxels_[x][y][z].data);
                hash[n] = ((uint64_t)x << 40) ^ ((uint64_t)y << 20) ^ (uint64_t)z;
                buffer_len[n] = (int32_t)d->buffer.size();
                normal_found[n] = d->normal_found ? 1 : 0;
                count[n] = d->count;
                if (normal) { normal[3 * n] = d->normal(0); normal[3 * n + 1] = d->normal(1); normal[3 * n + 2] = d->normal(2); }
                if (viewpoint) { viewpoint[3 * n] = d->viewpoint(0); viewpoint[3 * n + 1] = d->viewpoint(1); viewpoint[3 * n + 2] = d->viewpoint(2); }
                n++;
            }
}
}  // extern "C"
