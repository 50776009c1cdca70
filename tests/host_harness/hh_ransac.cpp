// Host-compiled copy of the DEVICE math header (imagemosaicing_b200/csrc/ransac_math.cuh) for CPU unit
// tests.  This is a test harness: it lets `pytest -m "not gpu"` check the kernel's arithmetic against
// the oracle without a GPU.  It is never loaded by the product library.
#include "../../imagemosaicing_b200/csrc/ransac_math.cuh"
extern "C" {
int hh_hypothesis(const float* x1, const float* y1, const float* x2, const float* y2, float* h, int* took_slow) {
    bool slow = false;
    int st = uavm::rmath::hypothesis(x1, y1, x2, y2, h, &slow);
    if (took_slow) *took_slow = slow ? 1 : 0;
    return st;
}
int hh_hypothesis_slow(const float* x1, const float* y1, const float* x2, const float* y2, float* h) {
    return uavm::rmath::hypothesis_slow(x1, y1, x2, y2, h);
}
int hh_draw_group(uint32_t seed, uint32_t g, int n, int* idx) { return uavm::rmath::draw_group(seed, g, n, idx) ? 1 : 0; }
}
extern "C" int hh_inverse8_generic_reg(const float* src, float eps, float* dst) {
    float N[8][8], out[8][8];
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) { N[i][j] = src[i * 8 + j]; out[i][j] = dst[i * 8 + j]; }
    int rc = uavm::rmath::inverse8_generic_reg(N, eps, out);
    if (rc == 1) for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) dst[i * 8 + j] = out[i][j];
    return rc;
}
