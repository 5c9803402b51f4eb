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
#include <cstring>
#include <cmath>
#include <cuda_runtime.h>

#define PCF_HD __host__ __device__ __forceinline__

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
    uint32_t nzp;           // z stride of the LOGICAL cell index: n1[2] rounded up to 32 (a bitmap word never straddles rows)
    uint32_t plane_cells;   // n1[1] * nzp: logical cells per x plane
    uint32_t nb[3];         // 64^3 bricks per axis of the PHYSICAL first-frame grid
    uint64_t cells;         // logical cells = n1[0] * n1[1] * nzp  (sort keys, bitmaps, x-major order)
    uint64_t phys_cells;    // nb[0]*nb[1]*nb[2] * 64^3 (allocated first_frame entries)
    float walk_step[7];     // float(i * xres_), i = -3..3  (OG.hpp:405, scalar narrowed before the product)
    int32_t walk_k;
    int32_t min_neighbours;
    float ball_radius_f;    // float(kBballRadius), OG.hpp:42
    float cylinder_thr;     // float threshold equivalent to the double compare of OG.hpp:262:  (double)d < radius  <=>  d < cylinder_thr
    double cylinder_radius; // OG.hpp:36
};

struct V3 { float x, y, z; };
PCF_HD V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
PCF_HD V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
PCF_HD V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
PCF_HD V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
PCF_HD V3 operator/(V3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
// Eigen's 3-term reduction order: x0 + (x1 + x2)
PCF_HD float dot(V3 a, V3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
PCF_HD float sqnorm(V3 a) { return dot(a, a); }
PCF_HD V3 normalized(V3 a) {
    float z = sqnorm(a);
    if (z > 0.0f) return a / sqrtf(z);
    return a;
}
PCF_HD V3 cross(V3 a, V3 b) {
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
PCF_HD bool finite3(V3 a) { return isfinite(a.x) && isfinite(a.y) && isfinite(a.z); }

// ---- exact conversions without the XU pipe ------------------------------------------------------------------
// F2F / FRND / F2I run on the 16-lane XU pipe; at 15 of them per point the integration kernel was XU-bound
// (ncu r01: pipe_xu 45 %, fp64 19 %).  The three helpers below produce bit-identical results with integer and
// FP64-pipe instructions only; values outside their fast range take the hardware conversion (rare branch).
PCF_HD uint32_t d_hi(double d) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__double2hiint(d);
#else
    uint64_t u; memcpy(&u, &d, 8); return (uint32_t)(u >> 32);
#endif
}
PCF_HD uint32_t d_lo(double d) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__double2loint(d);
#else
    uint64_t u; memcpy(&u, &d, 8); return (uint32_t)u;
#endif
}
PCF_HD double mk_double(uint32_t hi, uint32_t lo) {
#ifdef __CUDA_ARCH__
    return __hiloint2double((int)hi, (int)lo);
#else
    uint64_t u = ((uint64_t)hi << 32) | lo; double d; memcpy(&d, &u, 8); return d;
#endif
}
PCF_HD uint32_t f_bits(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
PCF_HD float bits_f(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

// double(f), exact.  Normal floats: re-bias the exponent (+896) and shift the mantissa by 3 bits.
PCF_HD double f2d_exact(float f) {
    uint32_t b = f_bits(f);
    uint32_t e = b & 0x7f800000u;
    if (e == 0u || e == 0x7f800000u) return (double)f;        // zero, denormal, inf, NaN
    return mk_double((b & 0x80000000u) | (((b & 0x7fffffffu) >> 3) + 0x38000000u), b << 29);
}

// f = float(d) (round to nearest even) and r = double(f).  For |d| in [2^-126, 2^127): adding M = sign(d)*2^(e+29)
// (e = exponent of d) makes the FP64 adder round d, to nearest even, to a multiple of 2^(e-23) -- the float
// grid at that exponent (the sum's significand is 2^52 + k with k the float significand, so "even" agrees);
// subtracting M again is exact.  The float bit pattern is then a pure re-packing of r.
PCF_HD void narrow_f32(double d, double& r, float& f) {
    uint32_t hi = d_hi(d);
    uint32_t e = hi & 0x7ff00000u;
    if (e - (897u << 20) <= ((1149u - 897u) << 20)) {
        double M = mk_double((hi & 0xfff00000u) + (29u << 20), 0u);
        r = (d + M) - M;
        uint32_t rh = d_hi(r), rl = d_lo(r);
        uint32_t m = (rh & 0x7fffffffu) - 0x38000000u;
        f = bits_f((rh & 0x80000000u) | (m << 3) | (rl >> 29));
    } else {
        f = (float)d;
        r = (double)f;
    }
}

// ---- rigid transform: out = float(T(r,0)*x + T(r,1)*y + T(r,2)*z + T(r,3)), double, left to right ----------
// (PCL transformPointCloud(Affine3d), node.cpp:289).  wd receives double(out) for the voxel index.
template <bool HW = false>
PCF_HD V3 transform_point(const double* __restrict__ T, float x, float y, float z, double wd[3]) {
    V3 o;
    if (HW) {      // hardware conversions (F2F on the XU pipe): the plain statement of the formula
        double dx = (double)x, dy = (double)y, dz = (double)z;
        o.x = (float)(T[0] * dx + T[1] * dy + T[2] * dz + T[3]);
        o.y = (float)(T[4] * dx + T[5] * dy + T[6] * dz + T[7]);
        o.z = (float)(T[8] * dx + T[9] * dy + T[10] * dz + T[11]);
        wd[0] = (double)o.x; wd[1] = (double)o.y; wd[2] = (double)o.z;
        return o;
    }
    double dx = f2d_exact(x), dy = f2d_exact(y), dz = f2d_exact(z);
    narrow_f32(T[0] * dx + T[1] * dy + T[2] * dz + T[3], wd[0], o.x);
    narrow_f32(T[4] * dx + T[5] * dy + T[6] * dz + T[7], wd[1], o.y);
    narrow_f32(T[8] * dx + T[9] * dy + T[10] * dz + T[11], wd[2], o.z);
    return o;
}

// strict box test; the float thresholds make it identical to the reference's double compares, NaN fails (D11)
PCF_HD bool valid_point(const GridParams& g, V3 p) {
    return p.x > g.lo[0] && p.x < g.hi[0] && p.y > g.lo[1] && p.y < g.hi[1] && p.z > g.lo[2] && p.z < g.hi[2];
}

// floor(a / res) for a point strictly inside the box (0 < a, a/res < 2^20), OG.hpp:633-635.
// q = a * (1/res) is within 2 ulp of the true quotient.  One FP64 add exposes floor(q) and the distance to the
// next cell border as integer fields; `near` asks the caller for voxel_axis_exact (kept out of line so that the
// hot path stays branch-free and needs no FRND / F2I on the XU pipe).
PCF_HD int voxel_axis_fast(double a, double inv_res, bool& near) {
    // t in [2^32, 2^33): ulp(t) = 2^-20, so the significand holds rint(q * 2^20): integer part of q in bits 51..20,
    // fraction in bits 19..0.  A non-zero fraction field proves q, and therefore the correctly rounded a / res
    // (|q - a/res| < 2^-31), lies strictly inside (n, n + 1): floor is n.  Fraction field 0 (p = 2^-20): divide.
    double t = a * inv_res + 4294967296.0;
    uint32_t hi = d_hi(t), lo = d_lo(t);
    near = (lo & 0xFFFFFu) == 0u;
    return (int)((hi << 12) | (lo >> 20));
}
PCF_HD int voxel_axis_exact(double a, double res) {
#ifdef __CUDA_ARCH__
    return (int)floor(__ddiv_rn(a, res));
#else
    return (int)floor(a / res);
#endif
}
PCF_HD int voxel_axis(double a, double res, double inv_res) {
    bool near;
    int i = voxel_axis_fast(a, inv_res, near);
    if (near) i = voxel_axis_exact(a, res);
    return i;
}
// pd = double(p) per axis (p passed valid_point)
PCF_HD void voxel_coords_d(const GridParams& g, const double pd[3], int& x, int& y, int& z) {
    x = voxel_axis(pd[0] - g.min[0], g.res[0], g.inv_res[0]);
    y = voxel_axis(pd[1] - g.min[1], g.res[1], g.inv_res[1]);
    z = voxel_axis(pd[2] - g.min[2], g.res[2], g.inv_res[2]);
}
PCF_HD void voxel_coords(const GridParams& g, V3 p, int& x, int& y, int& z) {
    double pd[3] = {f2d_exact(p.x), f2d_exact(p.y), f2d_exact(p.z)};
    voxel_coords_d(g, pd, x, y, z);
}
PCF_HD bool valid_coord(const GridParams& g, int x, int y, int z) {
    return x >= 0 && y >= 0 && z >= 0 && x < g.dim[0] && y < g.dim[1] && z < g.dim[2];
}
// LOGICAL cell index: lexicographic in (x, y, z) -- the reference's scan order (OG.hpp:463-465); sort key, bitmap index
PCF_HD uint32_t cell_index(const GridParams& g, int x, int y, int z) {
    return ((uint32_t)x * g.n1[1] + (uint32_t)y) * g.nzp + (uint32_t)z;
}
PCF_HD void cell_coords(const GridParams& g, uint32_t cell, int& x, int& y, int& z) {
    uint32_t xy = cell / g.nzp;
    z = (int)(cell - xy * g.nzp);
    x = (int)(xy / g.n1[1]);
    y = (int)(xy - (uint32_t)x * g.n1[1]);
}
// PHYSICAL index into first_frame: 64^3-cell bricks (1 MB each, z fastest inside a brick).  A surface patch touches a
// handful of 2 MB pages instead of one page per x plane: with the plain x-major grid the integration kernel ran 2.2x
// slower on a plate lying in the x-y plane than on the same plate in the y-z plane (TLB misses behind every probe).
PCF_HD uint32_t phys_index(const GridParams& g, int x, int y, int z) {
    uint32_t brick = (((uint32_t)x >> 6) * g.nb[1] + ((uint32_t)y >> 6)) * g.nb[2] + ((uint32_t)z >> 6);
    return (brick << 18) | (((uint32_t)x & 63u) << 12) | (((uint32_t)y & 63u) << 6) | ((uint32_t)z & 63u);
}
PCF_HD uint64_t hash_id(int x, int y, int z) {
    return ((uint64_t)x << 40) ^ ((uint64_t)y << 20) ^ (uint64_t)z;
}
PCF_HD float center_axis(const GridParams& g, int axis, int i) {
    return (float)(g.min[axis] + g.res[axis] * (double)i + g.half_res[axis]);
}
PCF_HD V3 voxel_center(const GridParams& g, int x, int y, int z) {
    return mk(center_axis(g, 0, x), center_axis(g, 1, y), center_axis(g, 2, z));
}

// ---- projection onto the normal axis and cylinder scoring --------------------------------------------------
struct Axis {      // per-voxel invariants of projectPointToVector(pt, centre, normal)
    V3 a, ab;
    float ab_ab;
};
PCF_HD Axis make_axis(const GridParams& g, V3 centre, V3 n) {
    V3 d = g.ball_radius_f * n;      // n * float(kBballRadius): commutative per component
    Axis ax;
    ax.a = centre - d;
    V3 b = centre + d;
    ax.ab = ax.a - b;
    ax.ab_ab = dot(ax.ab, ax.ab);
    return ax;
}
PCF_HD V3 project(const Axis& ax, V3 pt) {
    V3 ap = ax.a - pt;
    float t = dot(ap, ax.ab) / ax.ab_ab;
    return ax.a - t * ax.ab;
}

struct Stats {     // the scored part of VoxelInfo (OG.hpp:64-68,73); mean_dist starts at 0 (D1)
    V3 centroid, sd;
    float sd_dist, mean_dist;
    int count;
};
PCF_HD void stats_init(Stats& s) {
    s.centroid = mk(0, 0, 0); s.sd = mk(0, 0, 0); s.sd_dist = 0.f; s.mean_dist = 0.f; s.count = 0;
}
// The cylinder test of one point is independent of every other point; only the fold of the in-cylinder points into
// the running statistics is an ordered recurrence.  score_test / score_fold split the two so that the scoring kernel
// can evaluate several tests back to back (independent instruction chains) before folding them in buffer order.
PCF_HD float score_test(const Axis& ax, V3 pt, V3& proj) {
    proj = project(ax, pt);
    V3 diff = pt - proj;
    return sqrtf(sqnorm(diff));
}
// fold of one IN-CYLINDER point (OG.hpp:264-273 / 428-438)
PCF_HD void score_apply(Stats& s, V3 proj, float dist_f) {
    double dist = (double)dist_f;
    {
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
PCF_HD void score_fold(const GridParams& g, Stats& s, V3 proj, float dist_f) {
    if ((double)dist_f < g.cylinder_radius) score_apply(s, proj, dist_f);      // OG.hpp:262 / 426
}
PCF_HD void score_point(const GridParams& g, const Axis& ax, Stats& s, V3 pt) {
    V3 proj;
    float d = score_test(ax, pt, proj);
    score_fold(g, s, proj, d);
}

// ---- PCA normal ---------------------------------------------------------------------------------------------
struct CovAccum { float a[9]; };
PCF_HD void cov_init(CovAccum& c) {
#pragma unroll
    for (int i = 0; i < 9; i++) c.a[i] = 0.f;
}
PCF_HD void cov_add(CovAccum& c, float x, float y, float z) {
    c.a[0] += x * x; c.a[1] += x * y; c.a[2] += x * z;
    c.a[3] += y * y; c.a[4] += y * z; c.a[5] += z * z;
    c.a[6] += x; c.a[7] += y; c.a[8] += z;
}
// -> symmetric covariance m00,m01,m02,m11,m12,m22
PCF_HD void cov_finish(CovAccum& c, int n, float m[6]) {
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

PCF_HD float roots2_smallest() { return 0.0f; }   // computeRoots2 sets roots(0) = 0

// smallest root of the characteristic polynomial of the scaled matrix (pcl::computeRoots)
PCF_HD float smallest_root(float m00, float m01, float m02, float m11, float m12, float m22) {
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
PCF_HD V3 eigen33_smallest(const float m[6]) {
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
