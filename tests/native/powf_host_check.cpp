// CPU check of the powf restatement used by the device Phong term against the host libm
// (the libm the oracle and the reference link).  usage: powf_host_check <n> <seed>
// prints: "<n> <mismatches>"
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>
#include "../../580-raytracer_b200/csrc/powf_glibc.cuh"

static uint64_t s[2];
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t next() { uint64_t s0 = s[0], s1 = s[1], r = s0 + s1; s1 ^= s0; s[0] = rotl(s0, 24) ^ s1 ^ (s1 << 16); s[1] = rotl(s1, 37); return r; }

int main(int argc, char** argv) {
    long n = argc > 1 ? atol(argv[1]) : 1000000;
    s[0] = argc > 2 ? strtoull(argv[2], 0, 10) : 580; s[1] = 0x9E3779B97F4A7C15ull;
    static const float ys[] = { 2.0f, 5.0f, 10.0f, 32.0f, 700.0f, 900.0f, 1.0f, 0.5f, 3.0f, 64.0f, 128.0f, 1.0f / 2.2f };
    long bad = 0;
    for (long i = 0; i < n; i++) {
        uint64_t r = next();
        float x, y;
        switch (r & 3) {
        case 0: x = (float)((r >> 8) & 0xffffff) / 16777216.0f; break;                   // uniform value in [0,1)
        case 1: { uint32_t b = (uint32_t)(r >> 8) % 0x3f800001u; memcpy(&x, &b, 4); } break;   // uniform bits in [0,1]
        case 2: x = 1.0f - (float)((r >> 8) & 0xffff) / 16777216.0f; break;               // just below 1 (specular lobe)
        default: x = (float)((r >> 8) & 0xff) / 255.0f; break;                            // gamma inputs c/255
        }
        uint64_t r2 = next();
        if (r2 & 1) y = ys[(r2 >> 1) % (sizeof ys / sizeof ys[0])];
        else y = (float)((r2 >> 8) & 0xffffff) / 16777216.0f * 1000.0f;
        float a = rt580::powf_glibc(x, y), b = powf(x, y);
        if (memcmp(&a, &b, 4) != 0) { if (bad < 10) fprintf(stderr, "x=%a y=%a mine=%a libm=%a\n", x, y, a, b); bad++; }
    }
    printf("%ld %ld\n", n, bad);
    return 0;
}
