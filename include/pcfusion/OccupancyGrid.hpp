// pcfusion/OccupancyGrid.hpp -- C++ drop-in for the reference's `class OccupancyGrid`
// (pointcloud_fusion/pointcloud_fusion/include/utilities/OccupancyGrid.hpp:99-136), backed by libpcfusion.so
// (B200, sm_100a) through the C ABI in pcfusion.h.  Same method names, argument meaning and return conventions, so
// that `PointcloudFusion` (src/pointcloud_fusion_and_filter.cpp) compiles against it unchanged apart from the include
// line -- see INTEGRATION.md.  Header-only; link with -lpcfusion.
//
// What is the same:   setResolution / setDimensions / setK / construct (node.cpp:161-164), addPoints<N>(cloud, viewpoint)
//                     (node.cpp:292-295), updateThicknessVectors<N,K>() (node.cpp:311,317), downloadData(cloud, meta)
//                     (node.cpp:395-398), clearVoxels() (node.cpp:438), download / downloadHQ / downloadClassified
//                     (node.cpp:399-437, compiled out there), state_changed, the public geometry fields and helpers.
// What is different:  no `voxels_` / work-list members (the grid lives in HBM); every `bool` function really returns
//                     (reference D2); the declared deviations D1-D12 of SURVEY.md section 9 apply; errors are
//                     reported through last_error() instead of being undefined behaviour.
// The cloud type is a template parameter so this header does not itself depend on PCL: anything whose `->points`
// is a contiguous array of structs starting with float x, y, z (every pcl::Point* type) works.
#ifndef PCFUSION_OCCUPANCYGRID_HPP
#define PCFUSION_OCCUPANCYGRID_HPP

#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <string>
#include <tuple>
#include <type_traits>

#include "../pcfusion.h"

namespace pcfusion {

constexpr double kGoodPointsThreshold = 100;   // OG.hpp:34
constexpr double kBballRadius = 0.015;         // OG.hpp:35
constexpr double kCylinderRadius = 0.001;      // OG.hpp:36

class OccupancyGrid {
   public:
    double xmin_, xmax_, ymin_, ymax_, zmin_, zmax_;
    double xres_, yres_, zres_;
    int xdim_, ydim_, zdim_;
    int k_;
    int counter;
    bool state_changed;

    OccupancyGrid() : xmin_(0), xmax_(0), ymin_(0), ymax_(0), zmin_(0), zmax_(0), xres_(0), yres_(0), zres_(0),
                      xdim_(0), ydim_(0), zdim_(0), k_(2), counter(0), state_changed(false) {
        pcf_default_config(&cfg_);
        cfg_.clip_zmin = -INFINITY;    // addPoints() receives clouds the caller has already clipped (node.cpp:251)
        cfg_.clip_zmax = INFINITY;
    }
    ~OccupancyGrid() { pcf_destroy(ctx_); }
    OccupancyGrid(const OccupancyGrid&) = delete;
    OccupancyGrid& operator=(const OccupancyGrid&) = delete;

    // OG.hpp:604-612
    void setDimensions(double xmin, double xmax, double ymin, double ymax, double zmin, double zmax) {
        xmin_ = xmin; xmax_ = xmax; ymin_ = ymin; ymax_ = ymax; zmin_ = zmin; zmax_ = zmax;
    }
    // OG.hpp:614-619: float arguments on purpose (the double members hold float-rounded values)
    void setResolution(float x, float y, float z) {
        xres_ = x; yres_ = y; zres_ = z;
        res_f_[0] = x; res_f_[1] = y; res_f_[2] = z;
    }
    // OG.hpp:138-149.  Only k = 2 is meaningful: the reference's scan hard-codes 125 probes (OG.hpp:334, D9).
    bool setK(int k) { k_ = k; return true; }
    // Extra knobs the reference fixes at compile time (node.cpp:91-93, OG.hpp:35-36); call before construct().
    void setDevice(int device) { cfg_.device = device; }
    void setDepthClip(double zmin, double zmax) { cfg_.clip_zmin = zmin; cfg_.clip_zmax = zmax; }
    void setWalkK(int K) { cfg_.walk_k = K; }

    // OG.hpp:621-628: allocates the grid (here: in HBM).  false + last_error() on failure (no CUDA device, ...).
    bool construct() {
        pcf_destroy(ctx_);
        ctx_ = nullptr;
        const double box[6] = {xmin_, xmax_, ymin_, ymax_, zmin_, zmax_};
        std::memcpy(cfg_.box, box, sizeof box);
        std::memcpy(cfg_.res, res_f_, sizeof res_f_);
        cfg_.k_neighbourhood = k_;
        if (pcf_create(&cfg_, &ctx_) != PCF_OK) { err_ = pcf_last_error(nullptr); return false; }
        int32_t d[3];
        pcf_dims(ctx_, d);
        xdim_ = d[0]; ydim_ = d[1]; zdim_ = d[2];
        pcf_start(ctx_);
        frame_ = 0;
        return true;
    }

    // ---- coordinate helpers, same arithmetic as OG.hpp:630-650,151-165,131-135 (host side, for callers) ----
    template <class Vec3>
    std::tuple<int, int, int> getVoxelCoords(const Vec3& p) const {
        return std::make_tuple((int)std::floor((double(p(0)) - xmin_) / xres_), (int)std::floor((double(p(1)) - ymin_) / yres_),
                               (int)std::floor((double(p(2)) - zmin_) / zres_));
    }
    std::tuple<int, int, int> getVoxelCoords(unsigned long long hash) const {
        return std::make_tuple((int)(hash >> 40), (int)((hash >> 20) & 0xFFFFF), (int)(hash & 0xFFFFF));
    }
    unsigned long long getHashId(int x, int y, int z) const {
        return ((unsigned long long)x << 40) ^ ((unsigned long long)y << 20) ^ (unsigned long long)z;   // D7: 64-bit shifts
    }
    template <class Vec3>
    bool validPoints(const Vec3& p) const {
        return !(p(0) >= xmax_ || p(1) >= ymax_ || p(2) >= zmax_ || p(0) <= xmin_ || p(1) <= ymin_ || p(2) <= zmin_);
    }
    bool validCoord(int x, int y, int z) const { return x >= 0 && y >= 0 && z >= 0 && x < xdim_ && y < ydim_ && z < zdim_; }
    // OG.hpp:131-135: double arithmetic narrowed to float; returns any 3-float type constructible from (x, y, z)
    template <class Vec3>
    Vec3 getVoxelCenter(int x, int y, int z) const {
        return Vec3(xmin_ + xres_ * (x) + xres_ / 2.0, ymin_ + yres_ * (y) + yres_ / 2.0, zmin_ + zres_ * (z) + zres_ / 2.0);
    }
    // dims without a device (construct() needs one): xdim_ = (int)((xmax_ - xmin_) / xres_) etc., OG.hpp:623-625
    void computeDims() {
        xdim_ = (int)((xmax_ - xmin_) / xres_); ydim_ = (int)((ymax_ - ymin_) / yres_); zdim_ = (int)((zmax_ - zmin_) / zres_);
    }

    // OG.hpp:185-280.  `cloud` is in the fusion frame (the node transformed it, node.cpp:289); `viewpoint` is the
    // camera position.  N (OpenMP threads in the reference, whose pragmas are commented out) is ignored.
    template <int N, class CloudPtr, class Vec3>
    bool addPoints(const CloudPtr& cloud, const Vec3& viewpoint) {
        if (cloud == nullptr) return false;                      // OG.hpp:187-188
        if (!ctx_) return fail("construct() has not been called");
        const float vp[3] = {(float)viewpoint(0), (float)viewpoint(1), (float)viewpoint(2)};
        const size_t n = cloud->points.size();
        typedef typename std::remove_reference<decltype(cloud->points[0])>::type PointT;
        static_assert(sizeof(PointT) % 4 == 0, "point type must be a multiple of 4 bytes");
        const float* xyz = n ? reinterpret_cast<const float*>(&cloud->points[0]) : &vp[0];
        int rc = pcf_add_points(ctx_, xyz, (uint32_t)n, (uint32_t)(sizeof(PointT) / 4), vp, frame_++);
        if (rc < 0) return fail(pcf_last_error(ctx_));
        state_changed = true;                                    // OG.hpp:279
        return true;
    }
    template <int N, class CloudPtr>
    bool addPoints(const CloudPtr& cloud) {                      // default viewpoint {0,0,0}, OG.hpp:121
        struct Zero { float operator()(int) const { return 0.f; } } z;
        return addPoints<N>(cloud, z);
    }
    // Faster route for new callers: camera-frame cloud + fusion<-camera pose; the depth clip (setDepthClip) and the
    // FP64 transform run on the GPU (replaces node.cpp:248-255 + 288-290 + OG.hpp:185).  pose: row-major 4x4.
    bool integrateFrame(const float* xyz, uint32_t n, uint32_t stride_floats, const double pose[16]) {
        if (!ctx_) return fail("construct() has not been called");
        int rc = pcf_push_frame(ctx_, xyz, n, stride_floats, pose, frame_++);
        if (rc < 0) return fail(pcf_last_error(ctx_));
        state_changed = true;
        return true;
    }

    // OG.hpp:311-454.  K must equal the walk length given to setWalkK (default 3 = node.cpp:311,317).
    template <int N, int K>
    bool updateThicknessVectors() {
        if (!ctx_) return fail("construct() has not been called");
        if (K != cfg_.walk_k) return fail("updateThicknessVectors<N,K>: K differs from setWalkK()");
        if (pcf_update(ctx_) < 0) return fail(pcf_last_error(ctx_));
        pcf_stats st;
        pcf_get_stats(ctx_, &st);
        std::cout << "Total Voxels: " << st.occupied_voxels << std::endl;    // OG.hpp:317
        state_changed = false;                                               // OG.hpp:452
        return true;
    }

    // OG.hpp:456-488: ASCII PCD of PointXYZRGBNormal + CSV metadata, x-major order.  Does not clear the grid.
    bool downloadData(std::string cloud_location, std::string metadata) {
        pcf_result r;
        if (!extract(r)) return false;
        std::cout << "Copying the pointcloud..." << std::endl;
        if (pcf_write_result(&r, cloud_location.c_str(), metadata.c_str()) < 0) return fail("could not write the cloud / metadata files");
        std::cout << "Saved the pointcloud..." << std::endl << "Points: " << r.n << std::endl << "Saving the metadata..." << std::endl;
        return true;
    }
    // OG.hpp:491-512 / 577-601: append centroid (+ normal when the point type has one) of every voxel with a normal
    template <class CloudPtr>
    bool download(const CloudPtr& cloud) { return downloadIf(cloud, 0.0, false); }
    // OG.hpp:545-575: only voxels with count >= threshold; white points; sets width / height
    template <class CloudPtr>
    bool downloadHQ(const CloudPtr& cloud, double threshold = kGoodPointsThreshold) {
        if (!downloadIf(cloud, threshold, true)) return false;
        cloud->height = 1;
        cloud->width = (uint32_t)cloud->points.size();
        return true;
    }
    // OG.hpp:514-543: white, red where count > kGoodPointsThreshold
    template <class CloudPtr>
    bool downloadClassified(const CloudPtr& cloud) {
        if (cloud == nullptr) return false;
        pcf_result r;
        if (!extract(r)) return false;
        typedef typename std::remove_reference<decltype(cloud->points[0])>::type PointT;
        for (uint64_t i = 0; i < r.n; i++) {
            PointT pt;
            pt.x = r.centroid[3 * i]; pt.y = r.centroid[3 * i + 1]; pt.z = r.centroid[3 * i + 2];
            pt.r = 255; pt.g = 255; pt.b = 255;
            if (r.count[i] > kGoodPointsThreshold) { pt.g = 0; pt.b = 0; }
            cloud->points.push_back(pt);
        }
        std::cout << "Points: " << cloud->points.size() << std::endl;
        return true;
    }
    // OG.hpp:167-183 (D5: a full reset, the reference leaves stale holders behind)
    bool clearVoxels() {
        if (!ctx_) return fail("construct() has not been called");
        if (pcf_clear(ctx_) < 0) return fail(pcf_last_error(ctx_));
        frame_ = 0;
        return true;
    }

    const std::string& last_error() const { return err_; }
    pcf_ctx* handle() const { return ctx_; }   // for callers that want the C ABI directly (stats, timings, multi-GPU hooks)

   private:
    pcf_config cfg_;
    pcf_ctx* ctx_ = nullptr;
    float res_f_[3] = {0.f, 0.f, 0.f};
    uint32_t frame_ = 0;
    std::string err_;

    bool fail(const char* msg) { err_ = msg ? msg : "unknown error"; return false; }
    bool extract(pcf_result& r) {
        if (!ctx_) return fail("construct() has not been called");
        if (pcf_extract(ctx_, &r) < 0) return fail(pcf_last_error(ctx_));
        return true;
    }
    template <class PointT>
    static auto set_normal(PointT& pt, const float* n, int) -> decltype(pt.normal[0], void()) {
        pt.normal[0] = n[0]; pt.normal[1] = n[1]; pt.normal[2] = n[2];
    }
    template <class PointT>
    static void set_normal(PointT&, const float*, long) {}
    template <class PointT>
    static auto set_white(PointT& pt, int) -> decltype(pt.r, void()) { pt.r = 255; pt.g = 255; pt.b = 255; }
    template <class PointT>
    static void set_white(PointT&, long) {}
    template <class CloudPtr>
    bool downloadIf(const CloudPtr& cloud, double threshold, bool hq) {
        if (cloud == nullptr) return false;                      // OG.hpp:493-494
        pcf_result r;
        if (!extract(r)) return false;
        typedef typename std::remove_reference<decltype(cloud->points[0])>::type PointT;
        for (uint64_t i = 0; i < r.n; i++) {
            if (hq && r.count[i] < threshold) continue;          // OG.hpp:561
            PointT pt;
            pt.x = r.centroid[3 * i]; pt.y = r.centroid[3 * i + 1]; pt.z = r.centroid[3 * i + 2];
            if (hq) set_white(pt, 0);
            else set_normal(pt, r.normal + 3 * i, 0);
            cloud->points.push_back(pt);
        }
        std::cout << "Points: " << cloud->points.size() << std::endl;
        return true;
    }
};

}  // namespace pcfusion
#endif  // PCFUSION_OCCUPANCYGRID_HPP
