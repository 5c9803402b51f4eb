// =====================================================================================
// oracle/occupancy_grid_oracle.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Dependency-free CPU restatement of the reference's frame-integration + process() path:
//   OG.hpp   = /root/reference/pointcloud_fusion/pointcloud_fusion/include/utilities/OccupancyGrid.hpp
//   node.cpp = /root/reference/pointcloud_fusion/pointcloud_fusion/src/pointcloud_fusion_and_filter.cpp
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this file's shared object.  The product path (libpcfusion.so) never links or calls it.
//
// PARITY STATUS: "parity unpinned" for the third-party arithmetic.  The reference ships no
// tests, golden vectors or fixtures and needs ROS + PCL + Eigen, none of
// which exist here.  What IS pinned: the control flow of OG.hpp itself, because oracle/Makefile
// compiles the unmodified OG.hpp against small shim headers (oracle/ref_shim/) into
// oracle/_ref/libogref.so and tests/test_oracle_vs_ref.py compares this restatement against it.
// The PCL / Eigen arithmetic (covariance, eigen33, transformPointCloud, Vector3f op order) is
// restated from upstream PCL 1.8-1.10 / Eigen 3.3 behaviour; the
// reference does not pin a version (CMakeLists.txt:213 `find_package(PCL 1.7 REQUIRED)`).
//
// Floating point: build with -O2 -ffp-contract=off, x86-64 baseline (SSE2, no FMA), matching
// the reference's `-std=c++17 -O3` without -march / -ffast-math (CMakeLists.txt:5).
//
// Declared deviations from reference undefined behaviour:
//   D1  VoxelInfo::mean_dist starts at 0 (reference leaves it uninitialised, OG.hpp:68,74-81)
//   D3  update() visits unprocessed keys in ascending hash order (reference: libstdc++
//       unordered_set order, OG.hpp:314-316)
//   D5  clear() resets all state (reference leaks holders and work lists, OG.hpp:167-183)
//   D7  hash computed in 64 bit (reference shifts a 32-bit int, OG.hpp:154)
//   D11 points whose transformed coordinates are not finite are dropped (reference indexes
//       the grid with INT_MIN and crashes); a NaN walk position is skipped (reference relies
//       on x86 cvttsd2si returning INT_MIN, OG.hpp:405-411)
//   D12 computeRoots' atan2/cos/sin are evaluated in double and rounded to float, i.e. the
//       correctly rounded float result (reference: whatever libm's float routines return)
// =====================================================================================
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <unordered_set>
#include <utility>
#include <vector>

namespace ora {

// ---- mini Vector3f with Eigen 3.3's non-vectorised op order ----------
struct V3 {
    float x, y, z;
};
static inline V3 mk(float x, float y, float z) { return V3{x, y, z}; }
static inline V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
static inline V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
static inline V3 operator/(V3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
// 3-element reduction unrolls as x0 + (x1 + x2)
static inline float dot(V3 a, V3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
static inline float sqnorm(V3 a) { return dot(a, a); }
static inline float norm(V3 a) { return std::sqrt(sqnorm(a)); }
static inline V3 normalized(V3 a) {
    float z = sqnorm(a);
    if (z > 0.0f) return a / std::sqrt(z);
    return a;
}
static inline V3 cross(V3 a, V3 b) {
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

static const double kCylinderRadius = 0.001;  // OG.hpp:36
static const double kBallRadius = 0.015;      // OG.hpp:35

// OG.hpp:40-49
static inline V3 project_point_to_vector(V3 pt, V3 axis_pt, V3 n) {
    V3 d = n * (float)kBallRadius;
    V3 a = axis_pt - d;
    V3 b = axis_pt + d;
    V3 ap = a - pt;
    V3 ab = a - b;
    float t = dot(ap, ab) / dot(ab, ab);
    return a - t * ab;
}

// ---- PCL restatements ---------------------------------------
// pcl::computeMeanAndCovarianceMatrix (float accumulators, single pass, called at OG.hpp:302)
static void mean_and_covariance(const std::vector<V3>& pts, float cov[9], float centroid[3]) {
    float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (const V3& p : pts) {
        accu[0] += p.x * p.x;
        accu[1] += p.x * p.y;
        accu[2] += p.x * p.z;
        accu[3] += p.y * p.y;
        accu[4] += p.y * p.z;
        accu[5] += p.z * p.z;
        accu[6] += p.x;
        accu[7] += p.y;
        accu[8] += p.z;
    }
    float n = (float)pts.size();
    for (int i = 0; i < 9; i++) accu[i] /= n;
    centroid[0] = accu[6];
    centroid[1] = accu[7];
    centroid[2] = accu[8];
    cov[0] = accu[0] - accu[6] * accu[6];
    cov[1] = accu[1] - accu[6] * accu[7];
    cov[2] = accu[2] - accu[6] * accu[8];
    cov[4] = accu[3] - accu[7] * accu[7];
    cov[5] = accu[4] - accu[7] * accu[8];
    cov[8] = accu[5] - accu[8] * accu[8];
    cov[3] = cov[1];
    cov[6] = cov[2];
    cov[7] = cov[5];
}

static void compute_roots2(float b, float c, float roots[3]) {
    roots[0] = 0.0f;
    float d = (float)((double)(b * b) - 4.0 * (double)c);
    if (d < 0.0f) d = 0.0f;
    float sd = std::sqrt(d);
    roots[2] = 0.5f * (b + sd);
    roots[1] = 0.5f * (b - sd);
}

// m is the symmetric 3x3 (row major, m[3*r+c])
static void compute_roots(const float m[9], float roots[3]) {
    const float m00 = m[0], m01 = m[1], m02 = m[2], m11 = m[4], m12 = m[5], m22 = m[8];
    float c0 = m00 * m11 * m22 + 2.0f * m01 * m02 * m12 - m00 * m12 * m12 - m11 * m02 * m02 -
               m22 * m01 * m01;
    float c1 = m00 * m11 - m01 * m01 + m00 * m22 - m02 * m02 + m11 * m22 - m12 * m12;
    float c2 = m00 + m11 + m22;
    if (std::fabs(c0) < FLT_EPSILON) {
        compute_roots2(c2, c1, roots);
        return;
    }
    const float s_inv3 = (float)(1.0 / 3.0);
    const float s_sqrt3 = std::sqrt(3.0f);
    float c2_over_3 = c2 * s_inv3;
    float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
    if (a_over_3 > 0.0f) a_over_3 = 0.0f;
    float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
    float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
    if (q > 0.0f) q = 0.0f;
    float rho = std::sqrt(-a_over_3);
    // D12: correctly rounded float trig (double evaluation, then narrowing)
    float theta = (float)std::atan2((double)std::sqrt(-q), (double)half_b) * s_inv3;
    float cos_theta = (float)std::cos((double)theta);
    float sin_theta = (float)std::sin((double)theta);
    roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
    roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
    roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
    if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
    if (roots[1] >= roots[2]) {
        std::swap(roots[1], roots[2]);
        if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
    }
    if (roots[0] <= 0.0f) compute_roots2(c2, c1, roots);
}

// pcl::eigen33(mat, eigenvalue, eigenvector): smallest eigenpair (called via OG.hpp:289)
static V3 eigen33_smallest(const float mat[9], float* eigenvalue) {
    float scale = 0.0f;
    for (int i = 0; i < 9; i++) scale = std::max(scale, std::fabs(mat[i]));
    if (scale <= FLT_MIN) scale = 1.0f;
    float s[9];
    for (int i = 0; i < 9; i++) s[i] = mat[i] / scale;
    float roots[3];
    compute_roots(s, roots);
    if (eigenvalue) *eigenvalue = roots[0] * scale;
    s[0] -= roots[0];
    s[4] -= roots[0];
    s[8] -= roots[0];
    V3 r0 = mk(s[0], s[1], s[2]), r1 = mk(s[3], s[4], s[5]), r2 = mk(s[6], s[7], s[8]);
    V3 v1 = cross(r0, r1), v2 = cross(r0, r2), v3 = cross(r1, r2);
    float l1 = sqnorm(v1), l2 = sqnorm(v2), l3 = sqnorm(v3);
    if (l1 >= l2 && l1 >= l3) return v1 / std::sqrt(l1);
    if (l2 >= l1 && l2 >= l3) return v2 / std::sqrt(l2);
    return v3 / std::sqrt(l3);
}

// ---- data model (OG.hpp:51-82) --------------------------------------------------------------
struct VoxelInfo {
    V3 centroid{0, 0, 0};
    V3 normal{0, 0, 0};
    V3 sd{0, 0, 0};
    float sd_dist = 0;
    float mean_dist = 0;  // D1
    V3 viewpoint{0, 0, 0};
    std::vector<std::pair<V3, V3>> buffer;
    std::vector<uint64_t> dependants;
    bool normal_found = false;
    int count = 0;
};

struct Voxel {
    bool occupied = false;
    VoxelInfo* data = nullptr;
};

// The reference's grid is a dense vector<vector<vector<Voxel>>> (OG.hpp:108,626): 16 bytes per cell, 16 GB for the
// 1000^3 grids of C3/C4/C5.  The restatement keeps the same cell semantics (a never-touched cell reads as
// {occupied=false, data=nullptr}) in 8x8x8-cell pages allocated on first WRITE, so that full-size runs fit any host.
// Container choice only: tests/test_oracle_golden.py pins it against the reference's own dense header build.
class PagedVoxels {
   public:
    void init(int nx, int ny, int nz) {
        clear();
        nx_ = nx; ny_ = ny; nz_ = nz;
        px_ = (nx + 7) >> 3; py_ = (ny + 7) >> 3; pz_ = (nz + 7) >> 3;
        table_.assign((size_t)px_ * py_ * pz_, nullptr);
    }
    ~PagedVoxels() { clear(); }
    Voxel& at(int x, int y, int z) {
        Voxel*& pg = table_[page_of(x, y, z)];
        if (!pg) pg = new Voxel[512]();
        return pg[cell_of(x, y, z)];
    }
    const Voxel& peek(int x, int y, int z) const {
        const Voxel* pg = table_[page_of(x, y, z)];
        return pg ? pg[cell_of(x, y, z)] : empty_;
    }
    bool page_empty(int x, int y, int z) const { return table_[page_of(x, y, z)] == nullptr; }
    void clear() {
        for (Voxel* p : table_) delete[] p;
        table_.clear();
    }
    int nx() const { return nx_; }
    int ny() const { return ny_; }
    int nz() const { return nz_; }

   private:
    size_t page_of(int x, int y, int z) const { return ((size_t)(x >> 3) * py_ + (size_t)(y >> 3)) * pz_ + (size_t)(z >> 3); }
    static int cell_of(int x, int y, int z) { return ((x & 7) << 6) | ((y & 7) << 3) | (z & 7); }
    int nx_ = 0, ny_ = 0, nz_ = 0, px_ = 0, py_ = 0, pz_ = 0;
    std::vector<Voxel*> table_;
    Voxel empty_{};
};

struct Result {
    std::vector<uint64_t> hash;
    std::vector<float> centroid, normal, sd;  // 3 per entry
    std::vector<float> mean_dist, sd_dist;
    std::vector<int32_t> count;
};

struct Grid {
    double xmin, xmax, ymin, ymax, zmin, zmax;
    double xres, yres, zres;
    int xdim = 0, ydim = 0, zdim = 0;
    double clip_zmin, clip_zmax;
    int reserve_hint = 0;
    int walk_k = 3;          // K of updateThicknessVectors<N,K> (node.cpp:311)
    int min_neighbours = 20; // OG.hpp:352
    PagedVoxels voxels;   // (xdim+1) x (ydim+1) x (zdim+1) cells, OG.hpp:626
    std::unordered_set<uint64_t> unprocessed, processed;
    std::vector<VoxelInfo*> all_infos;  // ownership (lets clear() free holders too, D5)
    int dx[125], dy[125], dz[125];
    Result result;
    bool state_changed = false;

    // OG.hpp:604-628 (+ setK OG.hpp:138-149)
    Grid(const double box[6], const float res[3], double czmin, double czmax, int reserve) {
        xmin = box[0]; xmax = box[1]; ymin = box[2]; ymax = box[3]; zmin = box[4]; zmax = box[5];
        xres = res[0]; yres = res[1]; zres = res[2];  // float -> double, OG.hpp:614-619
        clip_zmin = czmin; clip_zmax = czmax; reserve_hint = reserve;
        int d = 0;
        for (int i = -2; i <= 2; i++)
            for (int j = -2; j <= 2; j++)
                for (int k = -2; k <= 2; k++) { dx[d] = i; dy[d] = j; dz[d] = k; d++; }
        xdim = (int)((xmax - xmin) / xres);
        ydim = (int)((ymax - ymin) / yres);
        zdim = (int)((zmax - zmin) / zres);
        voxels.init(xdim + 1, ydim + 1, zdim + 1);
    }
    ~Grid() { for (VoxelInfo* p : all_infos) delete p; }

    VoxelInfo* new_info() { VoxelInfo* p = new VoxelInfo(); all_infos.push_back(p); return p; }

    // OG.hpp:630-637
    void voxel_coords(V3 p, int& x, int& y, int& z) const {
        x = (int)std::floor(((double)p.x - xmin) / xres);
        y = (int)std::floor(((double)p.y - ymin) / yres);
        z = (int)std::floor(((double)p.z - zmin) / zres);
    }
    // OG.hpp:639-645
    bool valid_point(V3 p) const {
        return !((double)p.x >= xmax || (double)p.y >= ymax || (double)p.z >= zmax ||
                 (double)p.x <= xmin || (double)p.y <= ymin || (double)p.z <= zmin);
    }
    // OG.hpp:647-650
    bool valid_coord(int x, int y, int z) const {
        return x >= 0 && y >= 0 && z >= 0 && x < xdim && y < ydim && z < zdim;
    }
    // OG.hpp:151-156 (D7)
    static uint64_t hash_id(int x, int y, int z) {
        return ((uint64_t)x << 40) ^ ((uint64_t)y << 20) ^ (uint64_t)z;
    }
    // OG.hpp:158-165
    static void hash_coords(uint64_t id, int& x, int& y, int& z) {
        const uint64_t mask = (1u << 20) - 1;
        x = (int)(id >> 40); y = (int)((id >> 20) & mask); z = (int)(id & mask);
    }
    // OG.hpp:131-135
    V3 voxel_center(int x, int y, int z) const {
        return mk((float)(xmin + xres * x + xres / 2.0), (float)(ymin + yres * y + yres / 2.0),
                  (float)(zmin + zres * z + zres / 2.0));
    }

    // cylinder test + Welford update (OG.hpp:260-274 and OG.hpp:424-439 are the same arithmetic)
    static void score(VoxelInfo* v, V3 pt, V3 axis_pt) {
        V3 proj = project_point_to_vector(pt, axis_pt, v->normal);
        double dist = (double)norm(pt - proj);
        if (dist < kCylinderRadius) {
            v->count++;
            V3 old_mean = v->centroid;
            float c = (float)v->count;
            v->centroid = v->centroid + (proj - v->centroid) / c;
            v->sd.x = v->sd.x + ((proj.x - v->centroid.x) * (proj.x - old_mean.x) - v->sd.x) / c;
            v->sd.y = v->sd.y + ((proj.y - v->centroid.y) * (proj.y - old_mean.y) - v->sd.y) / c;
            v->sd.z = v->sd.z + ((proj.z - v->centroid.z) * (proj.z - old_mean.z) - v->sd.z) / c;
            float old_md = v->mean_dist;
            v->mean_dist = (float)((double)v->mean_dist + (dist - (double)v->mean_dist) / (double)v->count);
            v->sd_dist = (float)((double)v->sd_dist +
                                 ((dist - (double)v->mean_dist) * (dist - (double)old_md) - (double)v->sd_dist) /
                                     (double)v->count);
        }
    }

    // OG.hpp:185-280.  pts are already in the fusion frame.
    int64_t add_points(const float* xyz, int stride, int64_t n, V3 vp) {
        state_changed = true;
        int64_t inserted = 0;
        for (int64_t p = 0; p < n; p++) {
            V3 pt = mk(xyz[p * stride], xyz[p * stride + 1], xyz[p * stride + 2]);
            if (!valid_point(pt)) continue;
            if (!(std::isfinite(pt.x) && std::isfinite(pt.y) && std::isfinite(pt.z))) continue;  // D11
            int x, y, z;
            voxel_coords(pt, x, y, z);
            uint64_t hash = hash_id(x, y, z);
            Voxel& vox = voxels.at(x, y, z);
            inserted++;
            if (vox.occupied) {
                VoxelInfo* d = vox.data;
                if (!d->normal_found) d->buffer.push_back(std::make_pair(pt, vp));
                else unprocessed.erase(hash);
            } else {
                vox.occupied = true;
                unprocessed.insert(hash);
                if (vox.data == nullptr) vox.data = new_info();
                VoxelInfo* d = vox.data;
                if (reserve_hint > 0) d->buffer.reserve(reserve_hint);  // OG.hpp:228
                d->viewpoint = vp;
                d->buffer.push_back(std::make_pair(pt, vp));
            }
            // incremental scoring of registered dependants (OG.hpp:244-277)
            VoxelInfo* d = vox.data;
            size_t nd = d->dependants.size();
            for (size_t i = 0; i < nd; i++) {
                int xx, yy, zz;
                hash_coords(d->dependants[i], xx, yy, zz);
                VoxelInfo* dep = voxels.peek(xx, yy, zz).data;
                score(dep, pt, voxel_center(xx, yy, zz));
            }
        }
        return inserted;
    }

    // node.cpp:248-255 (z clip), node.cpp:288-290 + PCL transformPointCloud (appendix A.3)
    int64_t add_frame(const float* pts, int stride, int64_t n, const double T[16]) {
        std::vector<float> world;
        world.reserve((size_t)n * 3 / 2);
        for (int64_t i = 0; i < n; i++) {
            float x = pts[i * stride], y = pts[i * stride + 1], z = pts[i * stride + 2];
            if ((double)z < clip_zmax && (double)z > clip_zmin) {
                double dx_ = (double)x, dy_ = (double)y, dz_ = (double)z;
                world.push_back((float)(T[0] * dx_ + T[1] * dy_ + T[2] * dz_ + T[3]));
                world.push_back((float)(T[4] * dx_ + T[5] * dy_ + T[6] * dz_ + T[7]));
                world.push_back((float)(T[8] * dx_ + T[9] * dy_ + T[10] * dz_ + T[11]));
            }
        }
        V3 vp = mk((float)T[3], (float)T[7], (float)T[11]);
        return add_points(world.data(), 3, (int64_t)(world.size() / 3), vp);
    }

    // OG.hpp:311-454
    void update() {
        state_changed = false;
        std::vector<uint64_t> keys(unprocessed.begin(), unprocessed.end());
        std::sort(keys.begin(), keys.end());  // D3
        std::vector<V3> centres;
        for (uint64_t key : keys) {
            int x, y, z;
            hash_coords(key, x, y, z);
            const Voxel& vox = voxels.peek(x, y, z);
            if (!vox.occupied) continue;
            VoxelInfo* data = vox.data;
            int total = 0;
            centres.clear();
            for (int d = 0; d < 125; d++) {
                int i = dx[d], j = dy[d], k = dz[d];
                if (valid_coord(x + i, y + j, z + k) && voxels.peek(x + i, y + j, z + k).occupied) {
                    centres.push_back(voxel_center(x + i, y + j, z + k));
                    total++;
                }
            }
            if (!(total > min_neighbours && !data->normal_found)) continue;
            float cov[9], mean[3];
            mean_and_covariance(centres, cov, mean);
            V3 normal = eigen33_smallest(cov, nullptr);
            V3 centre = voxel_center(x, y, z);
            V3 dir = normalized(data->viewpoint - centre);
            if (dot(dir, normal) < 0.0f) normal = normal * -1.0f;
            data->normal = normal;
            data->normal_found = true;
            uint64_t hash = hash_id(x, y, z);
            processed.insert(hash);
            for (int i = -walk_k; i <= walk_k; i++) {
                V3 q = centre + (float)((double)i * xres) * data->normal;
                if (!(std::isfinite(q.x) && std::isfinite(q.y) && std::isfinite(q.z))) continue;  // D11
                if (!valid_point(q)) continue;
                int xx, yy, zz;
                voxel_coords(q, xx, yy, zz);
                if (!valid_coord(xx, yy, zz)) continue;
                Voxel& nb = voxels.at(xx, yy, zz);
                if (nb.occupied) {
                    nb.data->dependants.push_back(hash);
                    // copy: the reference iterates `for(auto pt: neighbor_data->buffer)` and the
                    // walk can land on the voxel itself; the buffer is not modified meanwhile.
                    for (const auto& pv : nb.data->buffer) score(data, pv.first, centre);
                } else {
                    VoxelInfo* holder = new_info();  // replaces any previous holder (OG.hpp:445-448)
                    holder->dependants.push_back(hash);
                    nb.data = holder;
                }
            }
        }
    }

    // OG.hpp:456-488 (scan order and filter; file writing is separate)
    int64_t download() {
        result = Result();
        for (int x = 0; x < xdim; x++)
            for (int y = 0; y < ydim; y++)
                for (int z = 0; z < zdim; z++) {
                    if ((z & 7) == 0 && voxels.page_empty(x, y, z)) { z += 7; continue; }
                    const Voxel& v = voxels.peek(x, y, z);
                    if (!v.occupied || !v.data->normal_found) continue;
                    const VoxelInfo* d = v.data;
                    result.hash.push_back(hash_id(x, y, z));
                    result.centroid.insert(result.centroid.end(), {d->centroid.x, d->centroid.y, d->centroid.z});
                    result.normal.insert(result.normal.end(), {d->normal.x, d->normal.y, d->normal.z});
                    result.sd.insert(result.sd.end(), {d->sd.x, d->sd.y, d->sd.z});
                    result.mean_dist.push_back(d->mean_dist);
                    result.sd_dist.push_back(d->sd_dist);
                    result.count.push_back(d->count);
                }
        return (int64_t)result.hash.size();
    }

    // OG.hpp:167-183 + D5
    void clear() {
        voxels.init(xdim + 1, ydim + 1, zdim + 1);
        for (VoxelInfo* p : all_infos) delete p;
        all_infos.clear();
        unprocessed.clear();
        processed.clear();
        result = Result();
        state_changed = true;
    }
};

}  // namespace ora

// ================================ C API (ctypes) ================================================
using ora::Grid;
using ora::V3;

extern "C" {

void* ora_create(const double* box, const float* res, double clip_zmin, double clip_zmax, int reserve_hint) {
    return new Grid(box, res, clip_zmin, clip_zmax, reserve_hint);
}
void ora_destroy(void* g) { delete (Grid*)g; }
void ora_dims(void* g, int* dims) {
    Grid* G = (Grid*)g; dims[0] = G->xdim; dims[1] = G->ydim; dims[2] = G->zdim;
}
int64_t ora_add_frame(void* g, const float* pts, int stride, int64_t n, const double* pose16) {
    return ((Grid*)g)->add_frame(pts, stride, n, pose16);
}
int64_t ora_add_points_world(void* g, const float* xyz, int stride, int64_t n, const float* vp) {
    return ((Grid*)g)->add_points(xyz, stride, n, ora::mk(vp[0], vp[1], vp[2]));
}
void ora_update(void* g) { ((Grid*)g)->update(); }
int64_t ora_download(void* g) { return ((Grid*)g)->download(); }
void ora_get_result(void* g, uint64_t* hash, float* centroid, float* normal, float* sd, float* mean_dist,
                    float* sd_dist, int32_t* count) {
    const ora::Result& r = ((Grid*)g)->result;
    size_t n = r.hash.size();
    if (!n) return;
    if (hash) memcpy(hash, r.hash.data(), n * 8);
    if (centroid) memcpy(centroid, r.centroid.data(), n * 12);
    if (normal) memcpy(normal, r.normal.data(), n * 12);
    if (sd) memcpy(sd, r.sd.data(), n * 12);
    if (mean_dist) memcpy(mean_dist, r.mean_dist.data(), n * 4);
    if (sd_dist) memcpy(sd_dist, r.sd_dist.data(), n * 4);
    if (count) memcpy(count, r.count.data(), n * 4);
}
void ora_clear(void* g) { ((Grid*)g)->clear(); }

// Canonical state dump: every occupied cell (pad cells included) in x-major order.
int64_t ora_state_size(void* g) {
    Grid* G = (Grid*)g;
    int64_t n = 0;
    for (int x = 0; x <= G->xdim; x++)
        for (int y = 0; y <= G->ydim; y++)
            for (int z = 0; z <= G->zdim; z++) {
                if ((z & 7) == 0 && G->voxels.page_empty(x, y, z)) { z += 7; continue; }
                n += G->voxels.peek(x, y, z).occupied ? 1 : 0;
            }
    return n;
}
void ora_get_state(void* g, uint64_t* hash, int32_t* buffer_len, uint8_t* normal_found, int32_t* count,
                   float* normal, float* viewpoint) {
    Grid* G = (Grid*)g;
    int64_t n = 0;
    for (int x = 0; x <= G->xdim; x++)
        for (int y = 0; y <= G->ydim; y++)
            for (int z = 0; z <= G->zdim; z++) {
                if ((z & 7) == 0 && G->voxels.page_empty(x, y, z)) { z += 7; continue; }
                const ora::Voxel& v = G->voxels.peek(x, y, z);
                if (!v.occupied) continue;
                hash[n] = Grid::hash_id(x, y, z);
                buffer_len[n] = (int32_t)v.data->buffer.size();
                normal_found[n] = v.data->normal_found ? 1 : 0;
                count[n] = v.data->count;
                if (normal) { normal[3 * n] = v.data->normal.x; normal[3 * n + 1] = v.data->normal.y; normal[3 * n + 2] = v.data->normal.z; }
                if (viewpoint) { viewpoint[3 * n] = v.data->viewpoint.x; viewpoint[3 * n + 1] = v.data->viewpoint.y; viewpoint[3 * n + 2] = v.data->viewpoint.z; }
                n++;
            }
}

// ---- file formats (OG.hpp:460-462,478 CSV; PCL savePCDFileASCII appendix A.4) ----------------
int ora_write_csv(void* g, const char* path) {
    const ora::Result& r = ((Grid*)g)->result;
    std::ofstream f(path);
    if (!f) return -1;
    f << "Id,sdx,sdy,sdz,mean distance from normal, distance from normal sd, points in cylinder" << std::endl;
    for (size_t i = 0; i < r.hash.size(); i++)
        f << i << "," << r.sd[3 * i] << "," << r.sd[3 * i + 1] << "," << r.sd[3 * i + 2] << "," << r.mean_dist[i]
          << "," << r.sd_dist[i] << "," << r.count[i] << "\n";
    return 0;
}
int ora_write_pcd(void* g, const char* path) {
    const ora::Result& r = ((Grid*)g)->result;
    std::ofstream f(path);
    if (!f) return -1;
    size_t n = r.hash.size();
    f << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\n"
      << "FIELDS x y z rgb normal_x normal_y normal_z curvature\nSIZE 4 4 4 4 4 4 4 4\n"
      << "TYPE F F F F F F F F\nCOUNT 1 1 1 1 1 1 1 1\nWIDTH " << n << "\nHEIGHT 1\n"
      << "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA ascii\n";
    std::ostringstream s;
    s.precision(8);
    s.imbue(std::locale::classic());
    auto put = [&](float v) { if (std::isnan(v)) s << "nan"; else s << v; };
    for (size_t i = 0; i < n; i++) {
        s.str("");
        put(r.centroid[3 * i]); s << " "; put(r.centroid[3 * i + 1]); s << " "; put(r.centroid[3 * i + 2]); s << " ";
        s << 4278190080u << " ";  // default PointXYZRGBNormal colour r=g=b=0,a=255 printed as uint32
        put(r.normal[3 * i]); s << " "; put(r.normal[3 * i + 1]); s << " "; put(r.normal[3 * i + 2]); s << " ";
        put(0.0f);
        f << s.str() << "\n";
    }
    return 0;
}

// ---- known-answer helpers: one reference formula each, for bit-compare against device code ----
void ora_kat_transform(const double* pose16, const float* in, int stride, int64_t n, float* out3) {
    const double* T = pose16;
    for (int64_t i = 0; i < n; i++) {
        double x = in[i * stride], y = in[i * stride + 1], z = in[i * stride + 2];
        out3[3 * i] = (float)(T[0] * x + T[1] * y + T[2] * z + T[3]);
        out3[3 * i + 1] = (float)(T[4] * x + T[5] * y + T[6] * z + T[7]);
        out3[3 * i + 2] = (float)(T[8] * x + T[9] * y + T[10] * z + T[11]);
    }
}
// voxel index + box validity for world points
void ora_kat_voxel(void* g, const float* xyz, int64_t n, int32_t* ijk, uint8_t* valid) {
    Grid* G = (Grid*)g;
    for (int64_t i = 0; i < n; i++) {
        V3 p = ora::mk(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        bool ok = G->valid_point(p) && std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z);
        valid[i] = ok;
        int x = -1, y = -1, z = -1;
        if (ok) G->voxel_coords(p, x, y, z);
        ijk[3 * i] = x; ijk[3 * i + 1] = y; ijk[3 * i + 2] = z;
    }
}
void ora_kat_center(void* g, const int32_t* ijk, int64_t n, float* out3) {
    Grid* G = (Grid*)g;
    for (int64_t i = 0; i < n; i++) {
        V3 c = G->voxel_center(ijk[3 * i], ijk[3 * i + 1], ijk[3 * i + 2]);
        out3[3 * i] = c.x; out3[3 * i + 1] = c.y; out3[3 * i + 2] = c.z;
    }
}
void ora_kat_project(const float* pt, const float* axis_pt, const float* nrm, int64_t n, float* out3, double* dist) {
    for (int64_t i = 0; i < n; i++) {
        V3 p = ora::mk(pt[3 * i], pt[3 * i + 1], pt[3 * i + 2]);
        V3 r = ora::project_point_to_vector(p, ora::mk(axis_pt[3 * i], axis_pt[3 * i + 1], axis_pt[3 * i + 2]),
                                            ora::mk(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]));
        out3[3 * i] = r.x; out3[3 * i + 1] = r.y; out3[3 * i + 2] = r.z;
        dist[i] = (double)ora::norm(p - r);
    }
}
// PCA normal of one neighbourhood (n points, xyz packed): covariance + eigen33
void ora_kat_normal(const float* xyz, int64_t n, float* cov9, float* normal3, float* eigenvalue) {
    std::vector<V3> pts((size_t)n);
    for (int64_t i = 0; i < n; i++) pts[i] = ora::mk(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    float mean[3];
    ora::mean_and_covariance(pts, cov9, mean);
    V3 v = ora::eigen33_smallest(cov9, eigenvalue);
    normal3[0] = v.x; normal3[1] = v.y; normal3[2] = v.z;
}
void ora_kat_eigen33(const float* cov9, float* normal3, float* eigenvalue) {
    V3 v = ora::eigen33_smallest(cov9, eigenvalue);
    normal3[0] = v.x; normal3[1] = v.y; normal3[2] = v.z;
}
// sequential Welford over a point list against one axis: the OG.hpp:426-439 recurrence
void ora_kat_score(const float* pts, int64_t n, const float* axis_pt, const float* nrm, float* centroid3,
                   float* sd3, float* mean_dist, float* sd_dist, int32_t* count) {
    ora::VoxelInfo v;
    v.normal = ora::mk(nrm[0], nrm[1], nrm[2]);
    V3 c = ora::mk(axis_pt[0], axis_pt[1], axis_pt[2]);
    for (int64_t i = 0; i < n; i++) Grid::score(&v, ora::mk(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), c);
    centroid3[0] = v.centroid.x; centroid3[1] = v.centroid.y; centroid3[2] = v.centroid.z;
    sd3[0] = v.sd.x; sd3[1] = v.sd.y; sd3[2] = v.sd.z;
    *mean_dist = v.mean_dist; *sd_dist = v.sd_dist; *count = v.count;
}

}  // extern "C"
