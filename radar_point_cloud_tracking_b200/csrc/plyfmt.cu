// Host-side ASCII PLY body writer (SURVEY.md section 8 f, rank 4): the lines np.savetxt(fh, data, fmt="%.4f %.4f %.4f %d %d %d")
// produces in write_ply_fast (5_gain_fusion_ply_builder.py:370-403; the same bytes as the per-point f-string loop of
// write_ply, T5:345-367, and as the ASCII branch of PointCloudWorkF/stdbscan_denoising_pipeline.py:828-851), for
// float32 coordinates and uint8 colours - without building the float64 table and formatting it row by row in Python.
//
// "%.4f" of a float32 value, exactly: v * 10^4 is exact in float64 (24-bit significand times 10^4 = 625 * 2^4 needs at
// most 34 bits), so printf's correctly rounded, ties-to-even result is nearbyint(v * 1e4) split into integer part and
// four fraction digits; the sign comes from the sign bit (printf writes "-0.0000" for -0.0 and for negatives that
// round to zero). Non-finite and huge values take snprintf itself ("nan" without a sign, as Python prints it).
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace {

inline char* put_uint(char* p, unsigned long long v) {
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}

inline char* put_fixed4(char* p, float f) {
    if (isnan(f)) { memcpy(p, "nan", 3); return p + 3; }
    if (!isfinite(f) || fabsf(f) >= 1e14f) return p + snprintf(p, 64, "%.4f", (double)f);
    const double r = nearbyint((double)f * 10000.0);               // exact product, ties to even
    if (signbit(f)) *p++ = '-';
    const unsigned long long u = (unsigned long long)fabs(r);
    p = put_uint(p, u / 10000ull);
    unsigned frac = (unsigned)(u % 10000ull);
    *p++ = '.';
    p[3] = (char)('0' + frac % 10); frac /= 10;
    p[2] = (char)('0' + frac % 10); frac /= 10;
    p[1] = (char)('0' + frac % 10); frac /= 10;
    p[0] = (char)('0' + frac);
    return p + 4;
}

}  // namespace

extern "C" int rb_ply_append_ascii(const char* path, const float* x, const float* y, const float* z, const uint8_t* rgb,
                                   int64_t n) {
    if (!path || n < 0 || (n && (!x || !y || !z || !rgb))) {
        rb_set_error("rb_ply_append_ascii: bad arguments");
        return RB_ERR_ARG;
    }
    FILE* fh = fopen(path, "ab");
    if (!fh) {
        rb_set_error("rb_ply_append_ascii: cannot open %s: %s", path, strerror(errno));
        return RB_ERR_ARG;
    }
    constexpr int64_t CHUNK = 4096;                                 // lines per write
    constexpr int PLY_LINE_CAP = 3 * 48 + 3 * 4 + 8;                    // three "%.4f" of up to 45 characters, three "%d", separators
    std::vector<char> buf((size_t)(CHUNK * PLY_LINE_CAP));
    int rc = RB_OK;
    for (int64_t i0 = 0; i0 < n && rc == RB_OK; i0 += CHUNK) {
        const int64_t i1 = i0 + CHUNK < n ? i0 + CHUNK : n;
        char* p = buf.data();
        for (int64_t i = i0; i < i1; ++i) {
            p = put_fixed4(p, x[i]); *p++ = ' ';
            p = put_fixed4(p, y[i]); *p++ = ' ';
            p = put_fixed4(p, z[i]); *p++ = ' ';
            p = put_uint(p, rgb[3 * i]); *p++ = ' ';
            p = put_uint(p, rgb[3 * i + 1]); *p++ = ' ';
            p = put_uint(p, rgb[3 * i + 2]); *p++ = '\n';
        }
        const size_t len = (size_t)(p - buf.data());
        if (fwrite(buf.data(), 1, len, fh) != len) {
            rb_set_error("rb_ply_append_ascii: write to %s failed: %s", path, strerror(errno));
            rc = RB_ERR_ARG;
        }
    }
    if (fclose(fh) != 0 && rc == RB_OK) {
        rb_set_error("rb_ply_append_ascii: closing %s failed: %s", path, strerror(errno));
        rc = RB_ERR_ARG;
    }
    return rc;
}
