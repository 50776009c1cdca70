// formats_host.cpp — the reference's on-disk artefacts as first-class I/O, host only:
//   matchPairs.match   int32 n + n x 40-byte MatchPointPairs        WriteMatchPairs / LoadMatchPairs  M/MosaicWithoutPos.cpp:4736-4749, :4774-4797
//   matchPairs.txt     "imgA xA yA fixedA imgB xB yB fixedB" lines  WriteMatchPairs_ASC2              :4751-4772
//   tran0.txt          rows for images 1..N-1: m0..m7 fixed         OutTransform                      :2798-2818
//   (import format)    count, then 9 floats per image               ImportTransform                   :2820-2843
//   keypoint_%d.key    int32 n + n x 28-byte cv::KeyPoint           WriteSurfKeyPoints / LoadSurfKeyPoints :4682-4734
//   discriptor_%d.xml  OpenCV FileStorage node "descriptor" (Mat)   same functions
// Text numbers are written the way the reference's `ofstream << float` does under its MSVC runtime (6 significant digits,
// ties away from zero, 3-digit exponents), so files written here are byte-identical to files written by the original tool.  Everything returns UAVM_OK / UAVM_EINVAL / UAVM_EFAIL.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "internal.h"

namespace {
struct File {
    FILE* f;
    File(const char* path, const char* mode) : f(path ? fopen(path, mode) : nullptr) {}
    ~File() { if (f) fclose(f); }
};
// `ofstream << float` of the reference's runtime (MSVC 9 CRT): %g with 6 significant digits, but decimal ties are rounded
// away from zero (436.5625 -> "436.563"; glibc's exact round-half-even gives "436.562") and exponents have three digits.
// Built from the exact decimal expansion of the float so the tie rule is applied to the true value.
std::string fmt_float(float v)
{
    if (v != v) return "nan";
    if (v == 0.0f) return "0";
    char e[80];
    snprintf(e, sizeof(e), "%.40e", (double)fabsf(v));          // d.dddd...e+XX, exact digits (glibc prints any precision exactly)
    if (e[0] < '0' || e[0] > '9') return v < 0 ? "-inf" : "inf";
    int exp10 = atoi(strchr(e, 'e') + 1);
    int dig[42]; int nd = 0;
    for (const char* p = e; *p && *p != 'e'; p++) if (*p >= '0' && *p <= '9') dig[nd++] = *p - '0';
    if (dig[6] >= 5) {                                           // round half up at the 6th significant digit
        int i = 5;
        while (i >= 0 && ++dig[i] == 10) { dig[i] = 0; i--; }
        if (i < 0) { for (int k = 5; k > 0; k--) dig[k] = dig[k - 1]; dig[0] = 1; exp10++; }
    }
    int last = 5;
    while (last > 0 && dig[last] == 0) last--;                   // %g strips trailing zeros
    std::string out = v < 0 ? "-" : "";
    if (exp10 < -4 || exp10 >= 6) {
        out += (char)('0' + dig[0]);
        if (last > 0) { out += '.'; for (int i = 1; i <= last; i++) out += (char)('0' + dig[i]); }
        char ex[16]; snprintf(ex, sizeof(ex), "e%c%03d", exp10 < 0 ? '-' : '+', exp10 < 0 ? -exp10 : exp10);
        out += ex;
    } else if (exp10 >= 0) {
        for (int i = 0; i <= exp10; i++) out += (char)('0' + (i <= 5 ? dig[i] : 0));
        if (last > exp10) { out += '.'; for (int i = exp10 + 1; i <= last; i++) out += (char)('0' + dig[i]); }
    } else {
        out += "0.";
        for (int i = 0; i < -exp10 - 1; i++) out += '0';
        for (int i = 0; i <= last; i++) out += (char)('0' + dig[i]);
    }
    return out;
}

bool read_all(const char* path, std::string& out)
{
    File fp(path, "rb");
    if (!fp.f) return false;
    char buf[1 << 16]; size_t n;
    while ((n = fread(buf, 1, sizeof(buf), fp.f)) > 0) out.append(buf, n);
    return true;
}
}  // namespace

// ---- matchPairs.match ------------------------------------------------------------------------------------------------
extern "C" int uavm_match_file_count(const char* path, int* n_out)
{
    if (!path || !n_out) return UAVM_EINVAL;
    File fp(path, "rb");
    int32_t n = 0;
    if (!fp.f || fread(&n, 4, 1, fp.f) != 1 || n < 0) return UAVM_EFAIL;       // LoadMatchPairs returns -1 when the count cannot be read
    *n_out = n;
    return UAVM_OK;
}
extern "C" int uavm_match_file_read(const char* path, uavm_matchpointpairs* out, int cap, int* n_out)
{
    if (!path || !n_out || (cap > 0 && !out)) return UAVM_EINVAL;
    File fp(path, "rb");
    int32_t n = 0;
    if (!fp.f || fread(&n, 4, 1, fp.f) != 1 || n < 0) return UAVM_EFAIL;
    *n_out = n;
    if (n > cap) return UAVM_EINVAL;
    if (n > 0 && fread(out, sizeof(uavm_matchpointpairs), (size_t)n, fp.f) != (size_t)n) return UAVM_EFAIL;
    return UAVM_OK;
}
extern "C" int uavm_match_file_write(const char* path, const uavm_matchpointpairs* pairs, int n)
{
    if (!path || n < 0 || (n > 0 && !pairs)) return UAVM_EINVAL;
    if (n == 0) return UAVM_OK;                                                  // the reference writes nothing for an empty list (:4739)
    File fp(path, "wb");
    const int32_t n32 = n;
    if (!fp.f || fwrite(&n32, 4, 1, fp.f) != 1 || fwrite(pairs, sizeof(uavm_matchpointpairs), (size_t)n, fp.f) != (size_t)n) return UAVM_EFAIL;
    return UAVM_OK;
}

// ---- matchPairs.txt ----------------------------------------------------------------------------------------------------
extern "C" int uavm_match_text_write(const char* path, const uavm_matchpointpairs* pairs, int n)
{
    if (!path || n < 0 || (n > 0 && !pairs)) return UAVM_EINVAL;
    File fp(path, "w");
    if (!fp.f) return UAVM_EFAIL;
    for (int i = 0; i < n; i++) {
        const uavm_matchpointpairs& m = pairs[i];
        if (fprintf(fp.f, "%d %s %s %d %d %s %s %d\n", m.ptA_i, fmt_float(m.ptA.x).c_str(), fmt_float(m.ptA.y).c_str(), m.ptA_Fixed,
                    m.ptB_i, fmt_float(m.ptB.x).c_str(), fmt_float(m.ptB.y).c_str(), m.ptB_Fixed) < 0) return UAVM_EFAIL;
    }
    return UAVM_OK;
}
// point ids are not part of the text format: they are returned as -1
extern "C" int uavm_match_text_read(const char* path, uavm_matchpointpairs* out, int cap, int* n_out)
{
    if (!path || !n_out || (cap > 0 && !out)) return UAVM_EINVAL;
    File fp(path, "r");
    if (!fp.f) return UAVM_EFAIL;
    int n = 0;
    for (;;) {
        int ia, fa, ib, fb; float xa, ya, xb, yb;
        const int got = fscanf(fp.f, "%d %f %f %d %d %f %f %d", &ia, &xa, &ya, &fa, &ib, &xb, &yb, &fb);
        if (got == EOF) break;
        if (got != 8) return UAVM_EFAIL;
        if (n < cap) {
            uavm_matchpointpairs& m = out[n];
            m.ptA.x = xa; m.ptA.y = ya; m.ptA.id = -1; m.ptA_i = ia; m.ptA_Fixed = fa;
            m.ptB.x = xb; m.ptB.y = yb; m.ptB.id = -1; m.ptB_i = ib; m.ptB_Fixed = fb;
        }
        n++;
    }
    *n_out = n;
    return n <= cap ? UAVM_OK : UAVM_EINVAL;
}

// ---- tran0.txt ---------------------------------------------------------------------------------------------------------
extern "C" int uavm_transform_file_write(const char* path, const uavm_imagetransform* t, int n_images)
{
    if (!path || !t || n_images < 1) return UAVM_EINVAL;
    File fp(path, "w");
    if (!fp.f) return UAVM_EFAIL;
    for (int i = 1; i < n_images; i++) {                                         // image 0 (the reference frame) is not written (:2804)
        for (int j = 0; j < 8; j++) if (fprintf(fp.f, "%s ", fmt_float(t[i].h.m[j]).c_str()) < 0) return UAVM_EFAIL;
        if (fprintf(fp.f, "%d\n", t[i].fixed) < 0) return UAVM_EFAIL;
    }
    return UAVM_OK;
}
// inverse of OutTransform: image 0 = identity, fixed; rows give m0..m7 and the fixed flag, m8 = 1
extern "C" int uavm_transform_file_read(const char* path, uavm_imagetransform* out, int cap, int* n_images_out)
{
    if (!path || !n_images_out || (cap > 0 && !out)) return UAVM_EINVAL;
    File fp(path, "r");
    if (!fp.f) return UAVM_EFAIL;
    int n = 1;
    if (cap > 0) { memset(&out[0], 0, sizeof(out[0])); out[0].h.m[0] = out[0].h.m[4] = out[0].h.m[8] = 1.0f; out[0].fixed = 1; }
    for (;;) {
        float m[8]; int fixed;
        int got = 0;
        for (int j = 0; j < 8; j++) { const int g = fscanf(fp.f, "%f", &m[j]); if (g != 1) { got = (j == 0 && g == EOF) ? EOF : -2; break; } got++; }
        if (got == EOF) break;
        if (got != 8 || fscanf(fp.f, "%d", &fixed) != 1) return UAVM_EFAIL;
        if (n < cap) { memcpy(out[n].h.m, m, sizeof(m)); out[n].h.m[8] = 1.0f; out[n].fixed = fixed; }
        n++;
    }
    *n_images_out = n;
    return n <= cap ? UAVM_OK : UAVM_EINVAL;
}
// ImportTransform's own format (:2820-2843): count, then 9 floats per image; image 0 is marked fixed
extern "C" int uavm_transform_import(const char* path, uavm_imagetransform* out, int cap, int* n_images_out)
{
    if (!path || !n_images_out || (cap > 0 && !out)) return UAVM_EINVAL;
    File fp(path, "r");
    int n = 0;
    if (!fp.f || fscanf(fp.f, "%d", &n) != 1 || n < 0) return UAVM_EFAIL;
    *n_images_out = n;
    if (n > cap) return UAVM_EINVAL;
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < 9; j++) if (fscanf(fp.f, "%f", &out[i].h.m[j]) != 1) return UAVM_EFAIL;
        out[i].fixed = (i == 0) ? 1 : 0;
    }
    return UAVM_OK;
}

// ---- keypoint_%d.key ---------------------------------------------------------------------------------------------------
extern "C" int uavm_key_file_read(const char* path, uavm_keypoint* out, int cap, int* n_out)
{
    if (!path || !n_out || (cap > 0 && !out)) return UAVM_EINVAL;
    File fp(path, "rb");
    int32_t n = 0;
    if (!fp.f || fread(&n, 4, 1, fp.f) != 1 || n < 0) return UAVM_EFAIL;
    *n_out = n;
    if (n > cap) return UAVM_EINVAL;
    if (n > 0 && fread(out, sizeof(uavm_keypoint), (size_t)n, fp.f) != (size_t)n) return UAVM_EFAIL;
    return UAVM_OK;
}
extern "C" int uavm_key_file_write(const char* path, const uavm_keypoint* kp, int n)
{
    if (!path || n < 0 || (n > 0 && !kp)) return UAVM_EINVAL;
    if (n == 0) return UAVM_OK;                                                  // the reference skips empty keypoint lists (:4693)
    File fp(path, "wb");
    const int32_t n32 = n;
    if (!fp.f || fwrite(&n32, 4, 1, fp.f) != 1 || fwrite(kp, sizeof(uavm_keypoint), (size_t)n, fp.f) != (size_t)n) return UAVM_EFAIL;
    return UAVM_OK;
}

// ---- discriptor_%d.xml (OpenCV FileStorage, node "descriptor", CV_32F matrix) ----------------------------------------------
static bool xml_int(const std::string& s, size_t from, const char* tag, long& v)
{
    const std::string open = std::string("<") + tag + ">";
    const size_t p = s.find(open, from);
    if (p == std::string::npos) return false;
    v = strtol(s.c_str() + p + open.size(), nullptr, 10);
    return true;
}
extern "C" int uavm_descriptor_xml_size(const char* path, int* rows, int* cols)
{
    if (!path || !rows || !cols) return UAVM_EINVAL;
    std::string s;
    if (!read_all(path, s)) return UAVM_EFAIL;
    const size_t node = s.find("<descriptor");
    long r = 0, c = 0;
    if (node == std::string::npos || !xml_int(s, node, "rows", r) || !xml_int(s, node, "cols", c) || r < 0 || c < 0) return UAVM_EFAIL;
    *rows = (int)r; *cols = (int)c;
    return UAVM_OK;
}
extern "C" int uavm_descriptor_xml_read(const char* path, float* out, int cap_floats, int* rows, int* cols)
{
    if (!path || !rows || !cols || (cap_floats > 0 && !out)) return UAVM_EINVAL;
    std::string s;
    if (!read_all(path, s)) return UAVM_EFAIL;
    const size_t node = s.find("<descriptor");
    long r = 0, c = 0;
    if (node == std::string::npos || !xml_int(s, node, "rows", r) || !xml_int(s, node, "cols", c) || r < 0 || c < 0) return UAVM_EFAIL;
    *rows = (int)r; *cols = (int)c;
    const size_t dt = s.find("<dt>", node);
    if (dt == std::string::npos || s.compare(dt + 4, 1, "f") != 0) return UAVM_EFAIL;                 // the reference stores CV_32F descriptors
    if ((long long)r * c > cap_floats) return UAVM_EINVAL;
    size_t p = s.find("<data>", node);
    if (p == std::string::npos) return UAVM_EFAIL;
    const char* q = s.c_str() + p + 6;
    for (long long i = 0; i < (long long)r * c; i++) {
        char* end = nullptr;
        out[i] = strtof(q, &end);
        if (end == q) return UAVM_EFAIL;
        q = end;
    }
    return UAVM_OK;
}
extern "C" int uavm_descriptor_xml_write(const char* path, const float* desc, int rows, int cols)
{
    if (!path || rows < 0 || cols < 0 || ((long long)rows * cols > 0 && !desc)) return UAVM_EINVAL;
    File fp(path, "w");
    if (!fp.f) return UAVM_EFAIL;
    fprintf(fp.f, "<?xml version=\"1.0\"?>\n<opencv_storage>\n<descriptor type_id=\"opencv-matrix\">\n  <rows>%d</rows>\n  <cols>%d</cols>\n  <dt>f</dt>\n  <data>\n", rows, cols);
    for (long long i = 0; i < (long long)rows * cols; i++) {
        if (i % 4 == 0) fputs("    ", fp.f);
        fprintf(fp.f, "%.8e%s", (double)desc[i], (i % 4 == 3 || i + 1 == (long long)rows * cols) ? "\n" : " ");
    }
    if (fputs("  </data></descriptor>\n</opencv_storage>\n", fp.f) < 0) return UAVM_EFAIL;
    return UAVM_OK;
}
