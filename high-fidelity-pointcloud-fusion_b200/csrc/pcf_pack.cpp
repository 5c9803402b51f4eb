// pcf_pack.cpp -- the staging pool's clip-and-pack inner loop (see pcf_stager.hpp), host code only.
//
// One pass over a cloud: keep the points with clip_lo < z < clip_hi (node.cpp:251; NaN fails), write them in order as
// packed xyz.  The loop is a pure streaming filter, so it is written to run at memory speed: with AVX-512 one 64-byte
// load holds 4 float4 points (or 2 points of a 32-byte PointCloud2 layout), one compare gives the keep bits, one
// VCOMPRESSPS squeezes the surviving x,y,z lanes to the front and one unaligned store appends them.  AVX2 and scalar
// versions are selected at run time (__builtin_cpu_supports); all three produce identical bytes.
#include <immintrin.h>

#include <cstring>
#include <limits>

#include "pcf_stager.hpp"

namespace pcf {
namespace {

// bytes of software prefetch distance: a single core's demand misses cover only ~5 GB/s of DRAM stream (line fill
// buffers x latency); prefetching well ahead, across 4 KB page boundaries where the hardware streamer stops, keeps more
// lines in flight per core
#ifndef PCF_PREFETCH_AHEAD
#define PCF_PREFETCH_AHEAD 4096
#endif
constexpr int kPrefetchAhead = PCF_PREFETCH_AHEAD;

uint32_t pack_row_scalar(const uint8_t* p, uint32_t cols, uint32_t step, float lo, float hi, float* out, uint32_t k) {
    for (uint32_t c = 0; c < cols; c++, p += step) {
        __builtin_prefetch(p + kPrefetchAhead);
        float v[3];
        std::memcpy(v, p, 12);
        out[3 * (size_t)k] = v[0];
        out[3 * (size_t)k + 1] = v[1];
        out[3 * (size_t)k + 2] = v[2];
        k += (v[2] > lo && v[2] < hi) ? 1u : 0u;
    }
    return k;
}

// P = floats per point (4 or 8), xo = float offset of x inside the point.  `out` has 16 floats of slack past the last point.
template <int P>
__attribute__((target("avx512f,avx512vl,popcnt"))) uint32_t pack_row_avx512(const uint8_t* p, uint32_t cols, uint32_t xo, float lo, float hi,
                                                                            float* out, uint32_t k) {
    constexpr int PPV = 16 / P;                      // points per 512-bit vector
    const __m512 vlo = _mm512_set1_ps(lo), vhi = _mm512_set1_ps(hi);
    __mmask16 zmask = 0;
    for (int j = 0; j < PPV; j++) zmask |= (__mmask16)(1u << (j * P + xo + 2));
    const float* src = reinterpret_cast<const float*>(p);
    uint32_t c = 0;
    for (; c + PPV <= cols; c += PPV, src += 16) {
        _mm_prefetch(reinterpret_cast<const char*>(src) + kPrefetchAhead, _MM_HINT_T0);
        __m512 v = _mm512_loadu_ps(src);
        __mmask16 m = _mm512_cmp_ps_mask(v, vlo, _CMP_GT_OQ) & _mm512_cmp_ps_mask(v, vhi, _CMP_LT_OQ) & zmask;
        __mmask16 m3 = (__mmask16)((uint32_t)(m >> 2) * 7u);          // the x, y, z lanes of every kept point
        _mm512_storeu_ps(out + 3 * (size_t)k, _mm512_maskz_compress_ps(m3, v));
        k += (uint32_t)_mm_popcnt_u32(m);
    }
    if (c < cols) k = pack_row_scalar(reinterpret_cast<const uint8_t*>(src) + xo * 4, cols - c, P * 4, lo, hi, out, k);
    return k;
}

// (A variant with non-temporal output -- compressed lanes collected in an aligned line buffer and streamed out with VMOVNTPS,
//  to spare the read-for-ownership of every destination line -- was measured on the GPU host: 5.7-5.9 vs 6.4 G points/s for
//  the loop alone, no difference end to end.  Not kept.)
// float4 points, two per 256-bit vector
__attribute__((target("avx2,popcnt"))) uint32_t pack_row_avx2(const uint8_t* p, uint32_t cols, float lo, float hi, float* out, uint32_t k) {
    alignas(32) static const int32_t kIdx[4][8] = {{0, 1, 2, 4, 5, 6, 3, 7}, {0, 1, 2, 4, 5, 6, 3, 7}, {4, 5, 6, 0, 1, 2, 3, 7}, {0, 1, 2, 4, 5, 6, 3, 7}};
    const __m256 vlo = _mm256_set1_ps(lo), vhi = _mm256_set1_ps(hi);
    const float* src = reinterpret_cast<const float*>(p);
    uint32_t c = 0;
    for (; c + 2 <= cols; c += 2, src += 8) {
        _mm_prefetch(reinterpret_cast<const char*>(src) + kPrefetchAhead, _MM_HINT_T0);
        __m256 v = _mm256_loadu_ps(src);
        int m = _mm256_movemask_ps(_mm256_and_ps(_mm256_cmp_ps(v, vlo, _CMP_GT_OQ), _mm256_cmp_ps(v, vhi, _CMP_LT_OQ)));
        int sel = ((m >> 2) & 1) | ((m >> 5) & 2);                     // bit 0: point 0 kept, bit 1: point 1 kept
        _mm256_storeu_ps(out + 3 * (size_t)k, _mm256_permutevar8x32_ps(v, _mm256_load_si256(reinterpret_cast<const __m256i*>(kIdx[sel]))));
        k += (uint32_t)_mm_popcnt_u32((unsigned)sel);
    }
    if (c < cols) k = pack_row_scalar(reinterpret_cast<const uint8_t*>(src), cols - c, 16, lo, hi, out, k);
    return k;
}

int detect_isa() {
    int forced = -1;
    if (const char* e = std::getenv("PCF_PACK_ISA")) forced = std::atoi(e);        // 0 scalar, 1 avx2, 2 avx512 (tests)
    __builtin_cpu_init();
    int best = 0;
    if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("popcnt")) best = 1;
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl")) best = 2;
    return forced >= 0 && forced <= best ? forced : best;
}

}  // namespace

int clip_pack_isa() {
    static const int isa = detect_isa();
    return isa;
}

uint32_t clip_pack(const StageJob& j, float clip_lo, float clip_hi, float* out) {
    return clip_pack_with(clip_pack_isa(), j, clip_lo, clip_hi, out);
}

uint32_t clip_pack_with(int isa, const StageJob& j, float clip_lo, float clip_hi, float* out) {
    if (isa > clip_pack_isa() && isa > 0) {      // never run an instruction set the CPU lacks
        __builtin_cpu_init();
        if ((isa == 2 && !(__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl"))) ||
            (isa >= 1 && !__builtin_cpu_supports("avx2")))
            isa = clip_pack_isa();
    }
    uint32_t k = 0;
    for (uint32_t r = 0; r < j.rows; r++) {
        const uint8_t* row = j.data + (size_t)r * j.row_step;
        const uint32_t xo = j.x_offset / 4;
        if (isa == 2 && j.point_step == 16 && xo <= 1) k = pack_row_avx512<4>(row, j.cols, xo, clip_lo, clip_hi, out, k);
        else if (isa == 2 && j.point_step == 32 && xo <= 5) k = pack_row_avx512<8>(row, j.cols, xo, clip_lo, clip_hi, out, k);
        else if (isa >= 1 && j.point_step == 16 && xo == 0) k = pack_row_avx2(row, j.cols, clip_lo, clip_hi, out, k);
        else k = pack_row_scalar(row + j.x_offset, j.cols, j.point_step, clip_lo, clip_hi, out, k);
    }
    const float nan = std::numeric_limits<float>::quiet_NaN();
    while (k & 3u) { out[3 * (size_t)k] = 0.f; out[3 * (size_t)k + 1] = 0.f; out[3 * (size_t)k + 2] = nan; k++; }
    return k;
}

}  // namespace pcf
