// pcf_device.cuh -- device-side arithmetic of the fusion path.  Compiled with -fmad=false: every float and
// double operation below is a separately rounded IEEE op in the order written, because the reference's
// integer outputs (cylinder counts, walk cells) are thresholded functions of these floats.
//
// Reference formulas (OG.hpp = .../include/utilities/OccupancyGrid.hpp, node.cpp = .../src/pointcloud_fusion_and_filter.cpp):
//   transform        PCL transformPointCloud(Affine3d) as called at node.cpp:289
//   voxel index      OG.hpp:630-637, validity OG.hpp:639-650, centre OG.hpp:131-135, hash OG.hpp:151-156
//   projection       OG.hpp:40-49, cylinder test + Welford OG.hpp:260-274 / 424-439
//   PCA normal       pcl::computeMeanAndCovarianceMatrix + pcl::eigen33 as called at OG.hpp:289,302
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace pcf {

constexpr uint32_t kEmpty = 0x7FFFFFFFu;   // first_frame value of an unoccupied cell (INT32_MAX: a signed min-reduce across GPUs works too)
constexpr uint32_t kNone = 0xFFFFFFFFu;

// Everything a kernel needs to know about the grid.  Passed by value (fits the 4 KB param space).
struct GridParams {
    double min[3];          // xmin_, ymin_, zmin_
    double res[3];          // double(float res), OG.hpp:614-619
    double inv_res[3];      // 1.0 / res (fast path of the voxel index; exact division near cell borders)
    double half_res[3];     // res / 2.0
    float lo[3], hi[3];     // float thresholds equivalent to the double box compares of OG.hpp:644
    float clip_lo, clip_hi; // same for the camera-frame depth clip, node.cpp:251
    int32_t dim[3];         // xdim_, ydim_, zdim_
    uint32_t n1[3];         // dim + 1 (allocated cells per axis, OG.hpp:626)
    uint64_t cells;         // n1[0]*n1[1]*n1[2]
    float walk_step[7];     // float(i * xres_), i = -3..3  (OG.hpp:405, scalar narrowed before the product)
    int32_t walk_k;
    int32_t min_neighbours;
    float ball_radius_f;    // float(kBballRadius), OG.hpp:42
    double cylinder_radius; // OG.hpp:36
};

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ V3 operator/(V3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
// Eigen's 3-term reduction order: x0 + (x1 + x2)
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
__device__ __forceinline__ float sqnorm(V3 a) { return dot(a, a); }
__device__ __forceinline__ V3 normalized(V3 a) {
    float z = sqnorm(a);
    if (z > 0.0f) return a / sqrtf(z);
    return a;
}
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ bool finite3(V3 a) { return isfinite(a.x) && isfinite(a.y) && isfinite(a.z); }

// ---- rigid transform: out = float(T(r,0)*x + T(r,1)*y + T(r,2)*z + T(r,3)), double, left to right ----------
__device__ __forceinline__ V3 transform_point(const double* __restrict__ T, float x, float y, float z) {
    double dx = (double)x, dy = (double)y, dz = (double)z;
    V3 o;
    o.x = (float)(T[0] * dx + T[1] * dy + T[2] * dz + T[3]);
    o.y = (float)(T[4] * dx + T[5] * dy + T[6] * dz + T[7]);
    o.z = (float)(T[8] * dx + T[9] * dy + T[10] * dz + T[11]);
    return o;
}

// strict box test; the float thresholds make it identical to the reference's double compares, NaN fails (D11)
__device__ __forceinline__ bool valid_point(const GridParams& g, V3 p) {
    return p.x > g.lo[0] && p.x < g.hi[0] && p.y > g.lo[1] && p.y < g.hi[1] && p.z > g.lo[2] && p.z < g.hi[2];
}

// floor((double(p) - min) / res).  a * (1/res) is within 2 ulp of the true quotient; only when it lands
// within 1e-6 of an integer can floor() of the correctly rounded quotient differ, and then we divide.
__device__ __forceinline__ int voxel_axis(double a, double res, double inv_res) {
    double q = a * inv_res;
    double fq = floor(q);
    double fr = q - fq;
    if (fr < 1e-6 || fr > 1.0 - 1e-6) fq = floor(__ddiv_rn(a, res));
    return (int)fq;
}
__device__ __forceinline__ void voxel_coords(const GridParams& g, V3 p, int& x, int& y, int& z) {
    x = voxel_axis((double)p.x - g.min[0], g.res[0], g.inv_res[0]);
    y = voxel_axis((double)p.y - g.min[1], g.res[1], g.inv_res[1]);
    z = voxel_axis((double)p.z - g.min[2], g.res[2], g.inv_res[2]);
}
__device__ __forceinline__ bool valid_coord(const GridParams& g, int x, int y, int z) {
    return x >= 0 && y >= 0 && z >= 0 && x < g.dim[0] && y < g.dim[1] && z < g.dim[2];
}
__device__ __forceinline__ uint32_t cell_index(const GridParams& g, int x, int y, int z) {
    return ((uint32_t)x * g.n1[1] + (uint32_t)y) * g.n1[2] + (uint32_t)z;
}
__device__ __forceinline__ void cell_coords(const GridParams& g, uint32_t cell, int& x, int& y, int& z) {
    uint32_t xy = cell / g.n1[2];
    z = (int)(cell - xy * g.n1[2]);
    x = (int)(xy / g.n1[1]);
    y = (int)(xy - (uint32_t)x * g.n1[1]);
}
__device__ __forceinline__ uint64_t hash_id(int x, int y, int z) {
    return ((uint64_t)x << 40) ^ ((uint64_t)y << 20) ^ (uint64_t)z;
}
__device__ __forceinline__ float center_axis(const GridParams& g, int axis, int i) {
    return (float)(g.min[axis] + g.res[axis] * (double)i + g.half_res[axis]);
}
__device__ __forceinline__ V3 voxel_center(const GridParams& g, int x, int y, int z) {
    return mk(center_axis(g, 0, x), center_axis(g, 1, y), center_axis(g, 2, z));
}

// ---- projection onto the normal axis and cylinder scoring --------------------------------------------------
struct Axis {      // per-voxel invariants of projectPointToVector(pt, centre, normal)
    V3 a, ab;
    float ab_ab;
};
__device__ __forceinline__ Axis make_axis(const GridParams& g, V3 centre, V3 n) {
    V3 d = g.ball_radius_f * n;      // n * float(kBballRadius): commutative per component
    Axis ax;
    ax.a = centre - d;
    V3 b = centre + d;
    ax.ab = ax.a - b;
    ax.ab_ab = dot(ax.ab, ax.ab);
    return ax;
}
__device__ __forceinline__ V3 project(const Axis& ax, V3 pt) {
    V3 ap = ax.a - pt;
    float t = dot(ap, ax.ab) / ax.ab_ab;
    return ax.a - t * ax.ab;
}

struct Stats {     // the scored part of VoxelInfo (OG.hpp:64-68,73); mean_dist starts at 0 (D1)
    V3 centroid, sd;
    float sd_dist, mean_dist;
    int count;
};
__device__ __forceinline__ void stats_init(Stats& s) {
    s.centroid = mk(0, 0, 0); s.sd = mk(0, 0, 0); s.sd_dist = 0.f; s.mean_dist = 0.f; s.count = 0;
}
__device__ __forceinline__ void score_point(const GridParams& g, const Axis& ax, Stats& s, V3 pt) {
    V3 proj = project(ax, pt);
    V3 diff = pt - proj;
    double dist = (double)sqrtf(sqnorm(diff));
    if (dist < g.cylinder_radius) {
        s.count++;
        V3 old_mean = s.centroid;
        float c = (float)s.count;
        s.centroid = s.centroid + (proj - s.centroid) / c;
        s.sd.x = s.sd.x + ((proj.x - s.centroid.x) * (proj.x - old_mean.x) - s.sd.x) / c;
        s.sd.y = s.sd.y + ((proj.y - s.centroid.y) * (proj.y - old_mean.y) - s.sd.y) / c;
        s.sd.z = s.sd.z + ((proj.z - s.centroid.z) * (proj.z - old_mean.z) - s.sd.z) / c;
        float old_md = s.mean_dist;
        double dc = (double)s.count;
        s.mean_dist = (float)((double)s.mean_dist + (dist - (double)s.mean_dist) / dc);
        s.sd_dist = (float)((double)s.sd_dist +
                            ((dist - (double)s.mean_dist) * (dist - (double)old_md) - (double)s.sd_dist) / dc);
    }
}

// ---- PCA normal ---------------------------------------------------------------------------------------------
struct CovAccum { float a[9]; };
__device__ __forceinline__ void cov_init(CovAccum& c) {
#pragma unroll
    for (int i = 0; i < 9; i++) c.a[i] = 0.f;
}
__device__ __forceinline__ void cov_add(CovAccum& c, float x, float y, float z) {
    c.a[0] += x * x; c.a[1] += x * y; c.a[2] += x * z;
    c.a[3] += y * y; c.a[4] += y * z; c.a[5] += z * z;
    c.a[6] += x; c.a[7] += y; c.a[8] += z;
}
// -> symmetric covariance m00,m01,m02,m11,m12,m22
__device__ __forceinline__ void cov_finish(CovAccum& c, int n, float m[6]) {
    float fn = (float)n;
#pragma unroll
    for (int i = 0; i < 9; i++) c.a[i] = c.a[i] / fn;
    m[0] = c.a[0] - c.a[6] * c.a[6];
    m[1] = c.a[1] - c.a[6] * c.a[7];
    m[2] = c.a[2] - c.a[6] * c.a[8];
    m[3] = c.a[3] - c.a[7] * c.a[7];
    m[4] = c.a[4] - c.a[7] * c.a[8];
    m[5] = c.a[5] - c.a[8] * c.a[8];
}

__device__ __forceinline__ float roots2_smallest() { return 0.0f; }   // computeRoots2 sets roots(0) = 0

// smallest root of the characteristic polynomial of the scaled matrix (pcl::computeRoots)
__device__ __forceinline__ float smallest_root(float m00, float m01, float m02, float m11, float m12, float m22) {
    float c0 = m00 * m11 * m22 + 2.0f * m01 * m02 * m12 - m00 * m12 * m12 - m11 * m02 * m02 - m22 * m01 * m01;
    float c1 = m00 * m11 - m01 * m01 + m00 * m22 - m02 * m02 + m11 * m22 - m12 * m12;
    float c2 = m00 + m11 + m22;
    if (fabsf(c0) < FLT_EPSILON) return roots2_smallest();
    const float s_inv3 = (float)(1.0 / 3.0);
    const float s_sqrt3 = 1.7320508075688772f;   // sqrtf(3.0f)
    float c2_over_3 = c2 * s_inv3;
    float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
    if (a_over_3 > 0.0f) a_over_3 = 0.0f;
    float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
    float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
    if (q > 0.0f) q = 0.0f;
    float rho = sqrtf(-a_over_3);
    // D12: correctly rounded float trig = double evaluation, then narrowing
    float theta = (float)atan2((double)sqrtf(-q), (double)half_b) * s_inv3;
    float cos_theta = (float)cos((double)theta);
    float sin_theta = (float)sin((double)theta);
    float r0 = c2_over_3 + 2.0f * rho * cos_theta;
    float r1 = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
    float r2 = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
    float t;
    if (r0 >= r1) { t = r0; r0 = r1; r1 = t; }
    if (r1 >= r2) {
        t = r1; r1 = r2; r2 = t;
        if (r0 >= r1) { t = r0; r0 = r1; r1 = t; }
    }
    if (r0 <= 0.0f) return roots2_smallest();
    return r0;
}

// pcl::eigen33(mat, eigenvalue, eigenvector): eigenvector of the smallest eigenvalue
__device__ __forceinline__ V3 eigen33_smallest(const float m[6]) {
    float scale = fmaxf(fmaxf(fmaxf(fabsf(m[0]), fabsf(m[1])), fmaxf(fabsf(m[2]), fabsf(m[3]))),
                        fmaxf(fabsf(m[4]), fabsf(m[5])));
    if (scale <= FLT_MIN) scale = 1.0f;
    float s00 = m[0] / scale, s01 = m[1] / scale, s02 = m[2] / scale;
    float s11 = m[3] / scale, s12 = m[4] / scale, s22 = m[5] / scale;
    float ev = smallest_root(s00, s01, s02, s11, s12, s22);
    s00 -= ev; s11 -= ev; s22 -= ev;
    V3 r0 = mk(s00, s01, s02), r1 = mk(s01, s11, s12), r2 = mk(s02, s12, s22);
    V3 v1 = cross(r0, r1), v2 = cross(r0, r2), v3 = cross(r1, r2);
    float l1 = sqnorm(v1), l2 = sqnorm(v2), l3 = sqnorm(v3);
    if (l1 >= l2 && l1 >= l3) return v1 / sqrtf(l1);
    if (l2 >= l1 && l2 >= l3) return v2 / sqrtf(l2);
    return v3 / sqrtf(l3);
}

}  // namespace pcf
