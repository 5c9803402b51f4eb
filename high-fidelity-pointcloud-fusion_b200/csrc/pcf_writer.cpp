// pcf_writer.cpp -- on-disk output of process(): test_cloud.pcd + meta.csv, byte-compatible with what the
// reference writes (node.cpp:395-398 -> downloadData, OG.hpp:456-488):
//   meta.csv   header OG.hpp:462, one row per exported voxel OG.hpp:478 (default ostream float format = %g)
//   cloud      pcl::io::savePCDFileASCII<PointXYZRGBNormal> (OG.hpp:485): PCD v0.7 ASCII, precision 8,
//              rgb printed as the uint32 bit pattern of a default-constructed point (r=g=b=0, a=255), curvature 0
#include <cmath>
#include <cstdio>
#include <vector>

#include "../../include/pcfusion.h"

namespace {
inline int put_float(char* p, float v, const char* fmt) {
    if (std::isnan(v)) { p[0] = 'n'; p[1] = 'a'; p[2] = 'n'; return 3; }
    return snprintf(p, 32, fmt, (double)v);
}
}  // namespace

extern "C" int pcf_write_result(const pcf_result* r, const char* cloud_path, const char* meta_path) {
    if (!r) return PCF_ERR_INVALID;
    const size_t n = (size_t)r->n;
    std::vector<char> buf(1 << 20);
    if (meta_path) {
        FILE* f = fopen(meta_path, "w");
        if (!f) return PCF_ERR_IO;
        fputs("Id,sdx,sdy,sdz,mean distance from normal, distance from normal sd, points in cylinder\n", f);
        size_t used = 0;
        for (size_t i = 0; i < n; i++) {
            if (used + 256 > buf.size()) { fwrite(buf.data(), 1, used, f); used = 0; }
            char* p = buf.data() + used;
            p += snprintf(p, 32, "%zu", i);
            const float vals[5] = {r->sd[3 * i], r->sd[3 * i + 1], r->sd[3 * i + 2], r->mean_dist[i], r->sd_dist[i]};
            for (float v : vals) { *p++ = ','; p += put_float(p, v, "%g"); }
            p += snprintf(p, 32, ",%d\n", r->count[i]);
            used = (size_t)(p - buf.data());
        }
        fwrite(buf.data(), 1, used, f);
        if (fclose(f) != 0) return PCF_ERR_IO;
    }
    if (cloud_path) {
        FILE* f = fopen(cloud_path, "w");
        if (!f) return PCF_ERR_IO;
        fprintf(f,
                "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\n"
                "FIELDS x y z rgb normal_x normal_y normal_z curvature\nSIZE 4 4 4 4 4 4 4 4\n"
                "TYPE F F F F F F F F\nCOUNT 1 1 1 1 1 1 1 1\nWIDTH %zu\nHEIGHT 1\n"
                "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS %zu\nDATA ascii\n",
                n, n);
        size_t used = 0;
        for (size_t i = 0; i < n; i++) {
            if (used + 512 > buf.size()) { fwrite(buf.data(), 1, used, f); used = 0; }
            char* p = buf.data() + used;
            for (int k = 0; k < 3; k++) { p += put_float(p, r->centroid[3 * i + k], "%.8g"); *p++ = ' '; }
            p += snprintf(p, 32, "4278190080 ");
            for (int k = 0; k < 3; k++) { p += put_float(p, r->normal[3 * i + k], "%.8g"); *p++ = ' '; }
            *p++ = '0';
            *p++ = '\n';
            used = (size_t)(p - buf.data());
        }
        fwrite(buf.data(), 1, used, f);
        if (fclose(f) != 0) return PCF_ERR_IO;
    }
    return PCF_OK;
}
