// Host-side known-answer test of the conversion helpers in csrc/pcf_device.cuh (they are __host__ __device__):
// the integer / FP64-adder replacements for F2F, FRND and F2I must equal the plain C conversions bit for bit.
// Built and run by tests/test_host_kat.py (no GPU needed).  Exit code 0 = all equal.
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../high-fidelity-pointcloud-fusion_b200/csrc/pcf_device.cuh"

using namespace pcf;

static uint64_t dbits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static long fails = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (fails < 20) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } fails++; } } while (0)

static void check_f2d(float f) {
    double a = f2d_exact(f), b = (double)f;
    CHECK(dbits(a) == dbits(b) || (a != a && b != b), "f2d_exact(%a) = %a want %a", f, a, b);
}
static void check_narrow(double d) {
    double r; float f;
    narrow_f32(d, r, f);
    float wf = (float)d; double wr = (double)wf;
    CHECK(f_bits(f) == f_bits(wf) || (f != f && wf != wf), "narrow_f32(%a): f = %a want %a", d, f, wf);
    CHECK(dbits(r) == dbits(wr) || (r != r && wr != wr), "narrow_f32(%a): r = %a want %a", d, r, wr);
}
static void check_voxel(double a, double res) {
    int got = voxel_axis(a, res, 1.0 / res);
    int want = (int)floor(a / res);
    CHECK(got == want, "voxel_axis(%a, %a) = %d want %d", a, res, got, want);
}

int main(int argc, char** argv) {
    long n = argc > 1 ? atol(argv[1]) : 20000000;
    std::mt19937_64 rng(12345);
    // ---- float -> double: random bit patterns + every exponent with extreme mantissas
    for (long i = 0; i < n; i++) check_f2d(bits_f((uint32_t)rng()));
    for (uint32_t e = 0; e < 256; e++)
        for (uint32_t m : {0u, 1u, 2u, 0x400000u, 0x7ffffeu, 0x7fffffu})
            for (uint32_t s : {0u, 0x80000000u}) check_f2d(bits_f(s | (e << 23) | m));
    // ---- double -> float (RN-even) + back: random mantissas over the float exponent range and beyond,
    //      exact ties (29 low bits = 100...0), near-ties, range boundaries, specials
    for (long i = 0; i < n; i++) {
        uint64_t u = rng();
        int e = (int)(rng() % 300) - 150;                    // 2^-150 .. 2^149
        uint64_t bits = (u & 0x800fffffffffffffull) | ((uint64_t)(1023 + e) << 52);
        double d; memcpy(&d, &bits, 8);
        check_narrow(d);
        uint64_t tie = (bits & ~0x1fffffffull) | 0x10000000ull;
        memcpy(&d, &tie, 8); check_narrow(d);
        tie += 1; memcpy(&d, &tie, 8); check_narrow(d);
        tie -= 2; memcpy(&d, &tie, 8); check_narrow(d);
        uint64_t top = bits | 0xfffffffffffffull;            // rounds up into the next exponent
        memcpy(&d, &top, 8); check_narrow(d);
    }
    for (double d : {0.0, -0.0, 1.0, -1.0, 0x1p-126, 0x1.fffffffffffffp-127, 0x1p-127, 0x1p-149, 0x1p-150, 0x1.8p-150,
                     0x1.fffffep127, 0x1.ffffffp127, 0x1.fffffefffffffp127, 0x1p128, 1e300, -1e300, (double)INFINITY,
                     -(double)INFINITY, (double)NAN, 0x1.fffffffffffffp126, 0x1p127, 0x1.ffffffffffffep-126})
        check_narrow(d), check_narrow(-d);
    // ---- voxel index: random interior points and points on / next to every cell border
    for (float resf : {0.001f, 0.0005f, 0.005f, 0.002f, 0.015f, 0.25f}) {
        double res = (double)resf;
        std::uniform_real_distribution<double> U(0.0, 1.0);
        for (long i = 0; i < n / 4; i++) {
            double a = U(rng) * 1.0 + 1e-300;
            check_voxel((double)(float)a + 0.25 - 0.25, res);
            check_voxel(a, res);
        }
        for (int k = 0; k < 2100; k++) {
            double b = res * k;
            for (int s = -3; s <= 3; s++) {
                double a = b;
                for (int j = 0; j < (s < 0 ? -s : s); j++) a = nextafter(a, s < 0 ? -1.0 : 2.0);
                if (a > 0) check_voxel(a, res);
            }
            // the doubles (float(x) - min) the kernels actually produce near this border
            float fx = (float)(b - 0.25);
            for (int s = -2; s <= 2; s++) {
                float g = fx;
                for (int j = 0; j < (s < 0 ? -s : s); j++) g = nextafterf(g, s < 0 ? -1.f : 2.f);
                double a = (double)g - (-0.25);
                if (a > 0) check_voxel(a, res);
            }
        }
    }
    // ---- the two statements of the transform agree
    {
        std::uniform_real_distribution<double> U(-1.0, 1.0);
        for (long i = 0; i < n / 4; i++) {
            double T[12];
            for (double& t : T) t = U(rng);
            if (i % 7 == 0) { T[1] = 0; T[2] = 0; T[3] = 0; }
            float x = (float)U(rng), y = (float)U(rng), z = (float)(0.3 + 0.3 * U(rng));
            if (i % 11 == 0) x = 0.f;
            if (i % 13 == 0) y = bits_f((uint32_t)rng() & 0x807fffffu);   // denormal
            double wa[3], wb[3];
            V3 a = transform_point<false>(T, x, y, z, wa), b = transform_point<true>(T, x, y, z, wb);
            CHECK(f_bits(a.x) == f_bits(b.x) && f_bits(a.y) == f_bits(b.y) && f_bits(a.z) == f_bits(b.z), "transform floats differ");
            CHECK(dbits(wa[0]) == dbits(wb[0]) && dbits(wa[1]) == dbits(wb[1]) && dbits(wa[2]) == dbits(wb[2]), "transform doubles differ");
        }
    }
    printf("host_kat: %ld failures\n", fails);
    return fails ? 1 : 0;
}
