// oracle/ref_shim/pcl/point_cloud.h -- TEST INFRASTRUCTURE: stand-in for pcl::PointCloud and the three PCL
// routines OG.hpp calls (restated from upstream PCL 1.8-1.10).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <string>
#include <tuple>
#include <unordered_set>
#include <vector>
#include <Eigen/Core>
#include <pcl/point_types.h>
#ifdef OGREF_ORDERED_KEYS
#include <set>
// D3 pin: a drop-in for the two unordered_set work lists that iterates in ascending key order.
template <class K> struct ogref_ordered_set : std::set<K> {};
#define unordered_set ogref_ordered_set
#endif
namespace pcl {
template <class PointT> struct PointCloud {
    typedef std::shared_ptr<PointCloud<PointT>> Ptr;
    std::vector<PointT> points;
    uint32_t width = 0, height = 0;
    bool is_dense = true;
    size_t size() const { return points.size(); }
};

// pcl::computeMeanAndCovarianceMatrix, float, dense path
inline unsigned computeMeanAndCovarianceMatrix(const PointCloud<PointXYZ>& cloud, Eigen::Matrix3f& cov, Eigen::Vector4f& centroid) {
    float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (size_t i = 0; i < cloud.points.size(); ++i) {
        const PointXYZ& p = cloud.points[i];
        accu[0] += p.x * p.x; accu[1] += p.x * p.y; accu[2] += p.x * p.z;
        accu[3] += p.y * p.y; accu[4] += p.y * p.z; accu[5] += p.z * p.z;
        accu[6] += p.x; accu[7] += p.y; accu[8] += p.z;
    }
    unsigned n = (unsigned)cloud.points.size();
    if (n == 0) return 0;
    for (int i = 0; i < 9; i++) accu[i] /= (float)n;
    centroid[0] = accu[6]; centroid[1] = accu[7]; centroid[2] = accu[8]; centroid[3] = 1;
    cov(0, 0) = accu[0] - accu[6] * accu[6];
    cov(0, 1) = accu[1] - accu[6] * accu[7];
    cov(0, 2) = accu[2] - accu[6] * accu[8];
    cov(1, 1) = accu[3] - accu[7] * accu[7];
    cov(1, 2) = accu[4] - accu[7] * accu[8];
    cov(2, 2) = accu[5] - accu[8] * accu[8];
    cov(1, 0) = cov(0, 1); cov(2, 0) = cov(0, 2); cov(2, 1) = cov(1, 2);
    return n;
}

inline void computeRoots2(float b, float c, float* roots) {
    roots[0] = 0.f;
    float d = float(b * b - 4.0 * c);
    if (d < 0.0) d = 0.0;
    float sd = ::std::sqrt(d);
    roots[2] = 0.5f * (b + sd);
    roots[1] = 0.5f * (b - sd);
}
inline void computeRoots(const Eigen::Matrix3f& m, float* roots) {
    float c0 = m(0, 0) * m(1, 1) * m(2, 2) + 2.f * m(0, 1) * m(0, 2) * m(1, 2) - m(0, 0) * m(1, 2) * m(1, 2) -
               m(1, 1) * m(0, 2) * m(0, 2) - m(2, 2) * m(0, 1) * m(0, 1);
    float c1 = m(0, 0) * m(1, 1) - m(0, 1) * m(0, 1) + m(0, 0) * m(2, 2) - m(0, 2) * m(0, 2) + m(1, 1) * m(2, 2) -
               m(1, 2) * m(1, 2);
    float c2 = m(0, 0) + m(1, 1) + m(2, 2);
    if (std::fabs(c0) < FLT_EPSILON) { computeRoots2(c2, c1, roots); return; }
    const float s_inv3 = float(1.0 / 3.0);
    const float s_sqrt3 = std::sqrt(float(3.0));
    float c2_over_3 = c2 * s_inv3;
    float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
    if (a_over_3 > 0.f) a_over_3 = 0.f;
    float half_b = 0.5f * (c0 + c2_over_3 * (2.f * c2_over_3 * c2_over_3 - c1));
    float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
    if (q > 0.f) q = 0.f;
    float rho = std::sqrt(-a_over_3);
#ifdef OGREF_LIBM_FLOAT_TRIG
    float theta = std::atan2(std::sqrt(-q), half_b) * s_inv3;   // stock: libm float routines
    float cos_theta = std::cos(theta), sin_theta = std::sin(theta);
#else
    float theta = (float)std::atan2((double)std::sqrt(-q), (double)half_b) * s_inv3;  // D12 pin
    float cos_theta = (float)std::cos((double)theta), sin_theta = (float)std::sin((double)theta);
#endif
    roots[0] = c2_over_3 + 2.f * rho * cos_theta;
    roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
    roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
    if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
    if (roots[1] >= roots[2]) {
        std::swap(roots[1], roots[2]);
        if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
    }
    if (roots[0] <= 0) computeRoots2(c2, c1, roots);
}
inline void eigen33(const Eigen::Matrix3f& mat, float& eigenvalue, Eigen::Vector3f& eigenvector) {
    float scale = 0.f;
    for (int i = 0; i < 9; i++) scale = std::fabs(mat.m[i]) > scale ? std::fabs(mat.m[i]) : scale;
    if (scale <= FLT_MIN) scale = 1.f;
    Eigen::Matrix3f s;
    for (int i = 0; i < 9; i++) s.m[i] = mat.m[i] / scale;
    float ev[3];
    computeRoots(s, ev);
    eigenvalue = ev[0] * scale;
    s(0, 0) -= ev[0]; s(1, 1) -= ev[0]; s(2, 2) -= ev[0];
    Eigen::Vector3f r0(s(0, 0), s(0, 1), s(0, 2)), r1(s(1, 0), s(1, 1), s(1, 2)), r2(s(2, 0), s(2, 1), s(2, 2));
    Eigen::Vector3f vec1 = r0.cross(r1), vec2 = r0.cross(r2), vec3 = r1.cross(r2);
    float len1 = vec1.squaredNorm(), len2 = vec2.squaredNorm(), len3 = vec3.squaredNorm();
    if (len1 >= len2 && len1 >= len3) eigenvector = vec1 / std::sqrt(len1);
    else if (len2 >= len1 && len2 >= len3) eigenvector = vec2 / std::sqrt(len2);
    else eigenvector = vec3 / std::sqrt(len3);
}

namespace io {
// pcl::io::savePCDFileASCII<PointXYZRGBNormal>, precision 8
inline int savePCDFileASCII(const std::string& path, const PointCloud<PointXYZRGBNormal>& cloud) {
    std::ofstream f(path.c_str());
    if (!f) return -1;
    size_t n = cloud.points.size();
    f << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgb normal_x normal_y normal_z curvature\n"
      << "SIZE 4 4 4 4 4 4 4 4\nTYPE F F F F F F F F\nCOUNT 1 1 1 1 1 1 1 1\nWIDTH " << cloud.width << "\nHEIGHT "
      << cloud.height << "\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA ascii\n";
    std::ostringstream s;
    s.precision(8);
    s.imbue(std::locale::classic());
    for (size_t i = 0; i < n; i++) {
        const PointXYZRGBNormal& p = cloud.points[i];
        s.str("");
        const float vals[3] = {p.x, p.y, p.z};
        for (float v : vals) { if (std::isnan(v)) s << "nan"; else s << v; s << " "; }
        s << p.rgba << " ";
        const float nv[4] = {p.normal[0], p.normal[1], p.normal[2], p.curvature};
        for (int k = 0; k < 4; k++) { if (std::isnan(nv[k])) s << "nan"; else s << nv[k]; if (k < 3) s << " "; }
        f << s.str() << "\n";
    }
    return 0;
}
}  // namespace io
}  // namespace pcl
