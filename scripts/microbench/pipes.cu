// pipes.cu — instruction-throughput microbenchmark for the packed-f32x2 / PRMT / IMAD.WIDE mix of the K5 warp kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
// Prints warp-instructions per clock per SM for each instruction kind (4.0 = every SMSP issues every cycle).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define NACC 8

typedef unsigned long long u64;

#define KERNEL(name, DECL, BODY)                                                     \
    __global__ void __launch_bounds__(256) name(u64* out, u64 seed, float one) {     \
        DECL                                                                         \
        for (int it = 0; it < ITERS; it++) {                                         \
            _Pragma("unroll") for (int k = 0; k < NACC; k++) { BODY }                \
        }                                                                            \
        u64 acc = 0;                                                                 \
        _Pragma("unroll") for (int k = 0; k < NACC; k++) acc ^= (u64)x[k] ^ (u64)y[k]; \
        if (acc == 0x1234567) out[threadIdx.x] = acc;                                \
    }

#define DECL64 u64 x[NACC], y[NACC]; u64 a = seed | 0x3f8000013f800001ull, b = seed ^ 0x3a8000013a800001ull; \
    _Pragma("unroll") for (int k = 0; k < NACC; k++) { x[k] = seed + k * 0x0000100000001000ull + threadIdx.x; y[k] = 0; }
#define DECL32 uint32_t x[NACC], y[NACC]; uint32_t a = (uint32_t)seed | 0x3f800001u, b = (uint32_t)seed ^ 0x3a800001u; \
    _Pragma("unroll") for (int k = 0; k < NACC; k++) { x[k] = (uint32_t)seed + k * 0x1000u + threadIdx.x; y[k] = 0; }

KERNEL(k_ffma,   DECL32, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(a), "r"(b));)
KERNEL(k_fmul,   DECL32, asm volatile("mul.rn.f32 %0, %0, %1;" : "+r"(x[k]) : "r"(a));)
KERNEL(k_fadd,   DECL32, asm volatile("add.rn.f32 %0, %0, %1;" : "+r"(x[k]) : "r"(a));)
KERNEL(k_faddrz, DECL32, asm volatile("add.rz.f32 %0, %0, %1;" : "+r"(x[k]) : "r"(a));)
KERNEL(k_ffma2,  DECL64, asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[k]) : "l"(a), "l"(b));)
KERNEL(k_fmul2,  DECL64, asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x[k]) : "l"(a));)
KERNEL(k_fadd2,  DECL64, asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x[k]) : "l"(a));)
KERNEL(k_fadd2rz,DECL64, asm volatile("add.rz.f32x2 %0, %0, %1;" : "+l"(x[k]) : "l"(a));)
KERNEL(k_prmt,   DECL32, asm volatile("prmt.b32 %0, %0, %1, 0x7540;" : "+r"(x[k]) : "r"(a));)
KERNEL(k_lop3,   DECL32, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(a), "r"(b));)
KERNEL(k_imadw,  DECL64, asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[k]) : "r"((uint32_t)a), "r"((uint32_t)b));)
KERNEL(k_iadd,   DECL32, asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(a));)
KERNEL(k_f2i,    DECL32, asm volatile("cvt.rzi.s32.f32 %0, %0;" : "+r"(x[k]));)
// mixes (two instructions per body)
KERNEL(k_ffma_prmt,  DECL32, asm volatile("fma.rn.f32 %0, %0, %2, %3;\n prmt.b32 %1, %1, %2, 0x7540;" : "+r"(x[k]), "+r"(y[k]) : "r"(a), "r"(b));)
KERNEL(k_ffma2_prmt, DECL64, asm volatile("fma.rn.f32x2 %0, %0, %2, %3;\n {.reg .b32 lo, hi; mov.b64 {lo, hi}, %1; prmt.b32 lo, lo, hi, 0x7540; mov.b64 %1, {lo, hi};}" : "+l"(x[k]), "+l"(y[k]) : "l"(a), "l"(b));)
KERNEL(k_ffma2_ffma, DECL64, asm volatile("fma.rn.f32x2 %0, %0, %2, %3;\n {.reg .b32 lo, hi; mov.b64 {lo, hi}, %1; fma.rn.f32 lo, lo, hi, hi; mov.b64 %1, {lo, hi};}" : "+l"(x[k]), "+l"(y[k]) : "l"(a), "l"(b));)
KERNEL(k_ffma_fadd,  DECL32, asm volatile("fma.rn.f32 %0, %0, %2, %3;\n add.rn.f32 %1, %1, %2;" : "+r"(x[k]), "+r"(y[k]) : "r"(a), "r"(b));)
KERNEL(k_fmul_fadd,  DECL32, asm volatile("mul.rn.f32 %0, %0, %2;\n add.rn.f32 %1, %1, %3;" : "+r"(x[k]), "+r"(y[k]) : "r"(a), "r"(b));)

template <typename K>
static void run(const char* name, K kern, int per_body, int sms, double mhz)
{
    u64* out; cudaMalloc(&out, 4096);
    const int ctas = sms * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<ctas, 256>>>(out, 3, 1.0f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); kern<<<ctas, 256>>>(out, 3, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double warp_inst = (double)ctas * 8 * ITERS * NACC * per_body;
    const double clk = best * 1e-3 * mhz * 1e6;
    printf("%-14s %8.3f ms  %6.3f warp-inst/clk/SM (at %.0f MHz)\n", name, best, warp_inst / clk / sms, mhz);
    cudaFree(out);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    printf("%s, %d SMs, %.0f MHz nominal\n", p.name, p.multiProcessorCount, mhz);
    const int sms = p.multiProcessorCount;
#define RUN(k, n) run(#k, k, n, sms, mhz)
    RUN(k_ffma, 1); RUN(k_fmul, 1); RUN(k_fadd, 1); RUN(k_faddrz, 1);
    RUN(k_ffma2, 1); RUN(k_fmul2, 1); RUN(k_fadd2, 1); RUN(k_fadd2rz, 1);
    RUN(k_prmt, 1); RUN(k_lop3, 1); RUN(k_imadw, 1); RUN(k_iadd, 1); RUN(k_f2i, 1);
    RUN(k_ffma_prmt, 2); RUN(k_ffma2_prmt, 2); RUN(k_ffma2_ffma, 2); RUN(k_ffma_fadd, 2); RUN(k_fmul_fadd, 2);
    return 0;
}
