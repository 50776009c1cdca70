// ldtm.cu — TMEM -> register drain microbenchmark (tcgen05.ld), the ceiling of K2's fused arg-min epilogue.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm ldtm.cu && ./ldtm
// One CTA per SM allocates all 512 TMEM columns and W warps read them back with tcgen05.ld.32x32b.xN in a loop (a warp can
// only read the 32-lane quarter selected by warp_id % 4, so 4 warps cover the 128 lanes once and 8 warps cover every quarter
// twice, on different column halves).  Prints 32-bit words per clock per SM for W = 4, 8, 16 and x16 / x32 / x64 / x128
// shapes.  K2 needs one s32 per 128 MACs, i.e. 16 words / clk / SM at the 2048 MAC / clk / SM of kind::i8 M=128 — the number
// printed here is the roofline of that epilogue.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int N> struct Ld;
#define LD_BODY(N, REGS, ...)                                                                                          \
    template <> struct Ld<N> {                                                                                         \
        static __device__ __forceinline__ uint32_t go(uint32_t taddr) {                                                \
            uint32_t v[N];                                                                                             \
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x" #N ".b32 {" REGS "}, [%" #N "];" : __VA_ARGS__ : "r"(taddr) : "memory"); \
            uint32_t a = 0;                                                                                            \
            _Pragma("unroll") for (int k = 0; k < N; k++) a ^= v[k];                                                   \
            return a;                                                                                                  \
        }                                                                                                              \
    };
#define R4(b) "%" #b
#define O(i) "=r"(v[i])
LD_BODY(16, "%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15",
        O(0), O(1), O(2), O(3), O(4), O(5), O(6), O(7), O(8), O(9), O(10), O(11), O(12), O(13), O(14), O(15))
LD_BODY(32, "%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31",
        O(0), O(1), O(2), O(3), O(4), O(5), O(6), O(7), O(8), O(9), O(10), O(11), O(12), O(13), O(14), O(15), O(16), O(17), O(18), O(19), O(20),
        O(21), O(22), O(23), O(24), O(25), O(26), O(27), O(28), O(29), O(30), O(31))
LD_BODY(64, "%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
            "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63",
        O(0), O(1), O(2), O(3), O(4), O(5), O(6), O(7), O(8), O(9), O(10), O(11), O(12), O(13), O(14), O(15), O(16), O(17), O(18), O(19), O(20),
        O(21), O(22), O(23), O(24), O(25), O(26), O(27), O(28), O(29), O(30), O(31), O(32), O(33), O(34), O(35), O(36), O(37), O(38), O(39), O(40),
        O(41), O(42), O(43), O(44), O(45), O(46), O(47), O(48), O(49), O(50), O(51), O(52), O(53), O(54), O(55), O(56), O(57), O(58), O(59), O(60),
        O(61), O(62), O(63))

// LOADS_IN_FLIGHT tcgen05.ld are issued before one tcgen05.wait::ld (K2 issues 4 x32 loads per wait)
template <int N, int INFLIGHT>
__global__ void __launch_bounds__(512) k_ldtm(uint32_t* out, long long* cycles, int iters)
{
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t lane_base = (uint32_t)(warp & 3) * 32u;
    // warps sharing a lane quarter start at different columns
    const uint32_t col0 = (uint32_t)((warp >> 2) * N * INFLIGHT) % 512u;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int f = 0; f < INFLIGHT; f++) {
            const uint32_t col = (col0 + (uint32_t)(f * N) + (uint32_t)it * 32u) % (512u - (uint32_t)N + 1u) & ~31u;
            acc ^= Ld<N>::go(tbase + (lane_base << 16) + col);
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) out[threadIdx.x] = acc;
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
}

template <int N, int INFLIGHT>
static void run(int warps, int sms)
{
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, sizeof(long long) * sms);
    const int iters = 20000;
    k_ldtm<N, INFLIGHT><<<sms, warps * 32>>>(out, cyc, 100);
    cudaDeviceSynchronize();
    k_ldtm<N, INFLIGHT><<<sms, warps * 32>>>(out, cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("x%-3d warps %2d: %s\n", N, warps, cudaGetErrorString(e)); return; }
    long long h[256]; cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < sms; i++) c += (double)h[i]; c /= sms;
    const double words = (double)iters * INFLIGHT * N * 32.0 * warps;          // 32 lanes x N columns per warp-instruction
    printf("tcgen05.ld.32x32b.x%-3d  %2d warps, %d loads per wait: %7.2f words/clk/SM  (%6.1f B/clk/SM, %.0f cycles)\n", N, warps, INFLIGHT, words / c, 4.0 * words / c, c);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    const int sms = p.multiProcessorCount;
    for (int w : {4, 8, 16}) {
        run<16, 4>(w, sms); run<32, 1>(w, sms); run<32, 4>(w, sms); run<64, 2>(w, sms);
    }
    return 0;
}
