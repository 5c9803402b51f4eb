// oracle/ref_shim/pcl/point_types.h -- TEST INFRASTRUCTURE: stand-in for the PCL point structs OG.hpp touches.
#pragma once
#include <cstdint>
namespace pcl {
struct PointXYZ { float x = 0, y = 0, z = 0; };
struct PointXYZRGB {
    float x = 0, y = 0, z = 0;
    union { struct { uint8_t b, g, r, a; }; float rgb; uint32_t rgba; };
    PointXYZRGB() { r = g = b = 0; a = 255; }
};
struct PointXYZRGBNormal {
    float x = 0, y = 0, z = 0;
    union { struct { uint8_t b, g, r, a; }; float rgb; uint32_t rgba; };
    union { float normal[3]; struct { float normal_x, normal_y, normal_z; }; };
    float curvature = 0;
    PointXYZRGBNormal() { r = g = b = 0; a = 255; normal[0] = normal[1] = normal[2] = 0; }
};
}  // namespace pcl
