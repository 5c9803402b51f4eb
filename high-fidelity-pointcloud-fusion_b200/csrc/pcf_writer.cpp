// pcf_writer.cpp -- on-disk output of process(): test_cloud.pcd + meta.csv, byte-compatible with what the
// reference writes (node.cpp:395-398 -> downloadData, OG.hpp:456-488):
//   meta.csv   header OG.hpp:462, one row per exported voxel OG.hpp:478 (default ostream float format = %g)
//   cloud      pcl::io::savePCDFileASCII<PointXYZRGBNormal> (OG.hpp:485): PCD v0.7 ASCII, precision 8,
//              rgb printed as the uint32 bit pattern of a default-constructed point (r=g=b=0, a=255), curvature 0
// The reference formats one row at a time through iostreams (and flushes every CSV row, OG.hpp:478); at 6e5 voxels that
// is seconds, three orders of magnitude more than the GPU extraction in front of it.  Here rows are formatted with
// std::to_chars (shortest-path printf equivalents: chars_format::general with an explicit precision == %g / %.8g) by
// all host threads into per-thread blocks that are then written in order.
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pcfusion.h"

namespace {
inline char* put_float(char* p, float v, int precision) {      // == snprintf(p, ., "%.<precision>g", (double)v); NaN -> "nan"
    if (std::isnan(v)) { memcpy(p, "nan", 3); return p + 3; }
    return std::to_chars(p, p + 48, (double)v, std::chars_format::general, precision).ptr;
}
inline char* put_uint(char* p, unsigned long long v) { return std::to_chars(p, p + 24, v).ptr; }
inline char* put_int(char* p, int v) { return std::to_chars(p, p + 16, v).ptr; }

// rows [lo, hi) of either file into `out`
void format_meta(const pcf_result* r, size_t lo, size_t hi, std::string& out) {
    out.resize((hi - lo) * 112);
    char* p = &out[0];
    for (size_t i = lo; i < hi; i++) {
        p = put_uint(p, i);
        const float vals[5] = {r->sd[3 * i], r->sd[3 * i + 1], r->sd[3 * i + 2], r->mean_dist[i], r->sd_dist[i]};
        for (float v : vals) { *p++ = ','; p = put_float(p, v, 6); }
        *p++ = ',';
        p = put_int(p, r->count[i]);
        *p++ = '\n';
    }
    out.resize((size_t)(p - out.data()));
}
void format_cloud(const pcf_result* r, size_t lo, size_t hi, std::string& out) {
    out.resize((hi - lo) * 128);
    char* p = &out[0];
    for (size_t i = lo; i < hi; i++) {
        for (int k = 0; k < 3; k++) { p = put_float(p, r->centroid[3 * i + k], 8); *p++ = ' '; }
        memcpy(p, "4278190080 ", 11);
        p += 11;
        for (int k = 0; k < 3; k++) { p = put_float(p, r->normal[3 * i + k], 8); *p++ = ' '; }
        *p++ = '0';
        *p++ = '\n';
    }
    out.resize((size_t)(p - out.data()));
}

template <class F>
int write_rows(FILE* f, size_t n, F&& format) {
    // blocks of 64 K rows, formatted by up to 16 threads, written in order while later blocks are still being formatted
    const size_t block = 1 << 16, n_blocks = (n + block - 1) / block;
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    for (size_t b0 = 0; b0 < n_blocks; b0 += hw) {
        const size_t nb = std::min<size_t>(hw, n_blocks - b0);
        std::vector<std::string> out(nb);
        std::vector<std::thread> th;
        for (size_t k = 1; k < nb; k++)
            th.emplace_back([&, k] { format((b0 + k) * block, std::min(n, (b0 + k + 1) * block), out[k]); });
        format(b0 * block, std::min(n, (b0 + 1) * block), out[0]);
        for (auto& t : th) t.join();
        for (auto& s : out)
            if (fwrite(s.data(), 1, s.size(), f) != s.size()) return PCF_ERR_IO;
    }
    return PCF_OK;
}
}  // namespace

extern "C" int pcf_write_result(const pcf_result* r, const char* cloud_path, const char* meta_path) {
    if (!r) return PCF_ERR_INVALID;
    const size_t n = (size_t)r->n;
    if (meta_path) {
        FILE* f = fopen(meta_path, "w");
        if (!f) return PCF_ERR_IO;
        fputs("Id,sdx,sdy,sdz,mean distance from normal, distance from normal sd, points in cylinder\n", f);
        int rc = write_rows(f, n, [&](size_t lo, size_t hi, std::string& out) { format_meta(r, lo, hi, out); });
        if (fclose(f) != 0 || rc != PCF_OK) return PCF_ERR_IO;
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
        int rc = write_rows(f, n, [&](size_t lo, size_t hi, std::string& out) { format_cloud(r, lo, hi, out); });
        if (fclose(f) != 0 || rc != PCF_OK) return PCF_ERR_IO;
    }
    return PCF_OK;
}

// known-answer hook for the tests: format one float the way the two files do (precision 6 = CSV, 8 = PCD)
extern "C" int pcf_kat_format_float(float v, int precision, char* out32) {
    char* e = put_float(out32, v, precision);
    *e = 0;
    return (int)(e - out32);
}
