// ransac.cu — K4: RANSAC homography, bit-exact with Ransac2D (M/mosaicimage.h:1729-2035).
//
// The reference loop is sequential and data dependent, but its 4-index draws depend only on the RNG
// and on ptNum (the "4 distinct indices" rejection at :1801-1813 never looks at the data, and redraws
// always consume 4 rand() calls).  So draw group g = rand() calls 4g..4g+3 of the per-pair LCG stream
// is either a valid 4-tuple or skipped, and every group can be evaluated independently:
//   k4_ransac_eval      one thread per draw group: 4-point solve + 5 px gate + Gauss-Newton refine
//                       (ransac_math.cuh) + support count over all candidates (smem broadcast reads)
//   k4_ransac_finalize  one CTA per pair: replays the reference's sequential rule over the per-group
//                       results with prefix scans (t counts accepted tuples, realSamTimes counts valid
//                       tuples, first strict maximum wins, 0.99 early exit), evaluates further groups
//                       itself if the first pass did not reach sampleTimes, then the final inlier pass
//                       (division form, :1922-1944) and the final refine on all inliers (:1979-1986)
//                       with the reference's summation order.
#include <stdlib.h>
#include "internal.h"
#include "ransac_math.cuh"
#include "ransac_warp.cuh"

using namespace uavm::rmath;

namespace {

constexpr int kEvalThreads = 128;
constexpr int kFinThreads = 256;
constexpr uint32_t RES_VALID = 0x80000000u, RES_REJ = 0x40000000u, RES_DEFER = 0x20000000u, RES_SUP = 0x0000ffffu;
constexpr int kMaxGroups = 1 << 20;

// per-thread evaluation of draw group g with the fast path only; candidates are staged in shared memory
// (x1,y1,x2,y2 per point).  Tuples the fast path cannot reproduce come back as RES_VALID | RES_DEFER.
__device__ __forceinline__ uint32_t eval_group_fast(const float4* __restrict__ pts, int n, uint32_t seed, uint32_t g,
                                                    float thr2, float h[9])
{
    int idx[4];
    if (!draw_group(seed, g, n, idx)) return 0u;
    float x1[4], y1[4], x2[4], y2[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { float4 p = pts[idx[i]]; x1[i] = p.x; y1[i] = p.y; x2[i] = p.z; y2[i] = p.w; }
    const int st = hypothesis_fast(x1, y1, x2, y2, h);
    if (st == TUPLE_NEED_SLOW) return RES_VALID | RES_DEFER;
    if (st == TUPLE_REJECTED) return RES_VALID | RES_REJ;
    int sup = 0;
    for (int i = 0; i < n; i++) {
        const float4 p = pts[i];
        float xb, yb;
        project_mul(p.z, p.w, h, xb, yb);
        if (dist2(xb, yb, p.x, p.y) < thr2) sup++;
    }
    return RES_VALID | (uint32_t)sup;
}

// Warp-cooperative clean-up: every lane whose result is RES_DEFER gets its tuple evaluated by the whole warp
// with the generic algorithm (ransac_warp.cuh).  Must be called by all 32 lanes.
__device__ __forceinline__ void warp_resolve_deferred(const float4* __restrict__ pts, int n, uint32_t seed, uint32_t g,
                                                      float thr2, uint32_t& res, float h[9])
{
    const int lane = threadIdx.x & 31;
    unsigned m = __ballot_sync(0xffffffffu, (res & RES_DEFER) != 0u);
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t gs = __shfl_sync(0xffffffffu, g, src);
        int idx[4];
        draw_group(seed, gs, n, idx);
        float x1[4], y1[4], x2[4], y2[4], hh[9];
#pragma unroll
        for (int i = 0; i < 4; i++) { float4 p = pts[idx[i]]; x1[i] = p.x; y1[i] = p.y; x2[i] = p.z; y2[i] = p.w; }
        const int st = uavm::rwarp::warp_hypothesis_generic(x1, y1, x2, y2, hh, lane);
        uint32_t r = RES_VALID | RES_REJ;
        if (st != TUPLE_REJECTED) {
            int sup = 0;
            for (int i = lane; i < n; i += 32) {
                const float4 p = pts[i];
                float xb, yb;
                project_mul(p.z, p.w, hh, xb, yb);
                if (dist2(xb, yb, p.x, p.y) < thr2) sup++;
            }
            sup = __reduce_add_sync(0xffffffffu, sup);
            r = RES_VALID | (uint32_t)sup;
        }
        if (lane == src) {
            res = r;
#pragma unroll
            for (int i = 0; i < 9; i++) h[i] = hh[i];
        }
    }
}

__device__ __forceinline__ void stage_points(float4* pts, const float* xy1, const float* xy2, int n, int tid, int nthreads) {
    for (int i = tid; i < n; i += nthreads)
        pts[i] = make_float4(xy1[2 * i], xy1[2 * i + 1], xy2[2 * i], xy2[2 * i + 1]);
}

// support of hypothesis h over all candidates (scoring loop, :1889-1904): reciprocal-multiply projection
__device__ __forceinline__ int count_support(const float4* __restrict__ pts, int n, const float* h, float thr2)
{
    int sup = 0;
    for (int i = 0; i < n; i++) {
        const float4 p = pts[i];
        float xb, yb;
        project_mul(p.z, p.w, h, xb, yb);
        if (dist2(xb, yb, p.x, p.y) < thr2) sup++;
    }
    return sup;
}

// First pass over the draw groups of every pair.  One thread per group for the cheap part (draw, 4-point
// direct solve, 5 px gate; kGroupsPerThread groups per thread); the ~half of the tuples that survive the
// gate are COMPACTED in shared memory so that the expensive parts — the 15-iteration Gauss-Newton refine
// and the support count over all candidates — run on full, converged warps.
constexpr int kGroupsPerThread = 2;
constexpr int kGroupsPerBlock = kEvalThreads * kGroupsPerThread;

__global__ void __launch_bounds__(kEvalThreads, 3)
k4_ransac_eval(const PairDesc* __restrict__ pairs, const float* __restrict__ cand_xy1, const float* __restrict__ cand_xy2,
               const int32_t* __restrict__ cand_n, float thr2, int groups, uint32_t* __restrict__ tuple_res,
               float* __restrict__ tuple_h)
{
    __shared__ float4 pts[UAVM_CAND_SLOTS];
    __shared__ float samp[kGroupsPerBlock][17];   // x1[4] y1[4] x2[4] y2[4] of the tuples to refine (+1 pad)
    __shared__ float hs[kGroupsPerBlock][9];      // hypotheses (indexed by the tuple's position in the block)
    __shared__ short refine_src[kGroupsPerBlock], score_src[kGroupsPerBlock];
    __shared__ int n_refine, n_score;
    const int p = blockIdx.y;
    const int n = cand_n[p];
    const size_t base = (size_t)p * UAVM_CAND_SLOTS;
    const int tid = threadIdx.x;
    const uint32_t g0 = blockIdx.x * kGroupsPerBlock;
    uint32_t* res_out = tuple_res + (size_t)p * UAVM_RANSAC_MAX_TUPLES_FIRST + g0;
    float* h_out = tuple_h + ((size_t)p * UAVM_RANSAC_MAX_TUPLES_FIRST + g0) * 9;
    if (n < 4) {
        for (int k = tid; k < kGroupsPerBlock; k += kEvalThreads) if (g0 + k < (uint32_t)groups) res_out[k] = 0u;
        return;
    }
    stage_points(pts, cand_xy1 + base * 2, cand_xy2 + base * 2, n, tid, kEvalThreads);
    if (tid == 0) { n_refine = 0; n_score = 0; }
    __syncthreads();
    const uint32_t seed = pairs[p].seed;

    // ---- phase A: draw + direct solve + gate ----
    for (int k = tid; k < kGroupsPerBlock; k += kEvalThreads) {
        if (g0 + k >= (uint32_t)groups) break;
        int idx[4];
        uint32_t res = 0u;
        if (draw_group(seed, g0 + k, n, idx)) {
            float x1[4], y1[4], x2[4], y2[4], h[9];
#pragma unroll
            for (int i = 0; i < 4; i++) { const float4 q = pts[idx[i]]; x1[i] = q.x; y1[i] = q.y; x2[i] = q.z; y2[i] = q.w; }
            const int st = dlt_fast(x1, y1, x2, y2, h);
            res = RES_VALID;
            if (st == TUPLE_NEED_SLOW) res = RES_VALID | RES_DEFER;      // rare: evaluated by the finalize kernel (tiers 2/3)
            else if (st == TUPLE_REJECTED) res = RES_VALID | RES_REJ;
            else if (st == TUPLE_NEED_REFINE) {
                const int slot = atomicAdd(&n_refine, 1);
                refine_src[slot] = (short)k;
#pragma unroll
                for (int i = 0; i < 4; i++) { samp[slot][i] = x1[i]; samp[slot][4 + i] = y1[i]; samp[slot][8 + i] = x2[i]; samp[slot][12 + i] = y2[i]; }
#pragma unroll
                for (int i = 0; i < 9; i++) hs[k][i] = h[i];
            } else {
                score_src[atomicAdd(&n_score, 1)] = (short)k;
#pragma unroll
                for (int i = 0; i < 9; i++) hs[k][i] = h[i];
            }
        }
        res_out[k] = res;
    }
    __syncthreads();

    // ---- phase B: Gauss-Newton refine on the compacted survivors ----
    const int nr = n_refine;
    for (int t = tid; t < nr; t += kEvalThreads) {
        const int src = refine_src[t];
        float x1[4], y1[4], x2[4], y2[4], h[9];
#pragma unroll
        for (int i = 0; i < 4; i++) { x1[i] = samp[t][i]; y1[i] = samp[t][4 + i]; x2[i] = samp[t][8 + i]; y2[i] = samp[t][12 + i]; }
#pragma unroll
        for (int i = 0; i < 9; i++) h[i] = hs[src][i];
        const int st = refine_fast(x1, y1, x2, y2, h);
        if (st == TUPLE_NEED_SLOW) res_out[src] = RES_VALID | RES_DEFER;
        else {
#pragma unroll
            for (int i = 0; i < 9; i++) hs[src][i] = h[i];
            score_src[atomicAdd(&n_score, 1)] = (short)src;
        }
    }
    __syncthreads();

    // ---- phase C: support count on the compacted accepted hypotheses ----
    const int ns = n_score;
    for (int t = tid; t < ns; t += kEvalThreads) {
        const int src = score_src[t];
        float h[9];
#pragma unroll
        for (int i = 0; i < 9; i++) h[i] = hs[src][i];
        res_out[src] = RES_VALID | (uint32_t)count_support(pts, n, h, thr2);
#pragma unroll
        for (int i = 0; i < 9; i++) h_out[src * 9 + i] = h[i];
    }
}

// ---- block-wide helpers (kFinThreads = 256 threads, 8 warps) ----
__device__ __forceinline__ int block_excl_scan(int v, int* warp_tot /*[8]*/, int* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    __syncthreads();
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    int off = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kFinThreads / 32; w++) { int t = warp_tot[w]; if (w < wid) off += t; tot += t; }
    *total = tot;
    return off + inc - v;
}
__device__ __forceinline__ int block_reduce_max(int v, int* scratch /*[8]*/) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = __reduce_max_sync(0xffffffffu, v);
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    int m = scratch[0];
#pragma unroll
    for (int w = 1; w < kFinThreads / 32; w++) m = max(m, scratch[w]);
    return m;
}
__device__ __forceinline__ int block_reduce_min(int v, int* scratch) { return -block_reduce_max(-v, scratch); }

struct FinSmem {
    float4 pts[UAVM_CAND_SLOTS];
    float in_x1[UAVM_CAND_SLOTS], in_y1[UAVM_CAND_SLOTS], in_x2[UAVM_CAND_SLOTS], in_y2[UAVM_CAND_SLOTS];
    float C[2 * UAVM_CAND_SLOTS];
    float N1[64], N2[64];
    float w[8], dx[8];
    float hbest[9];
    int scratch[8];
    int flag;
};

__global__ void __launch_bounds__(kFinThreads, 1)
k4_ransac_finalize(const PairDesc* __restrict__ pairs, const float* __restrict__ cand_xy1, const float* __restrict__ cand_xy2,
                   const int32_t* __restrict__ cand_n, float thr2, int sample_times, int groups_first,
                   const uint32_t* __restrict__ tuple_res, float* __restrict__ tuple_h,
                   uint8_t* __restrict__ inlier, uavm_ransac_result* __restrict__ results)
{
    extern __shared__ __align__(16) uint8_t fin_raw[];
    FinSmem& S = *reinterpret_cast<FinSmem*>(fin_raw);
    float* J = reinterpret_cast<float*>(fin_raw + sizeof(FinSmem));        // [2N][8]
    float* L = J + 2 * UAVM_CAND_SLOTS * 8;                                // [8][2N]

    const int p = blockIdx.x, tid = threadIdx.x;
    const int n = cand_n[p];
    const size_t base = (size_t)p * UAVM_CAND_SLOTS;
    const uint32_t seed = pairs[p].seed;
    uavm_ransac_result R;
    R.ok = 0; R.n_inliers = 0; R.max_support = 0; R.best_tuple = -1; R.n_tuples = 0; R.n_counted = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) R.H[i] = 0.0f;
    for (int i = tid; i < UAVM_CAND_SLOTS; i += kFinThreads) inlier[base + i] = 0;
    if (n < 4) { if (tid == 0) results[p] = R; return; }
    stage_points(S.pts, cand_xy1 + base * 2, cand_xy2 + base * 2, n, tid, kFinThreads);
    __syncthreads();

    // ---------------- replay of the sequential sampling loop (:1785-1920) ----------------
    const float inv_n = 1.0f / (float)n;
    int k_valid = 0, t_acc = 0;
    int best_sup = 0, best_g = -1, best_k = -1, first_acc_g = -1;
    int n_tuples = 0, n_counted = 0;
    bool done = false;
    for (int g0 = 0; !done && g0 < kMaxGroups; g0 += kFinThreads) {
        const int g = g0 + tid;
        uint32_t res;
        float hd[9];
        bool fresh = false;                                      // evaluated here (not by the first pass)
        if (g < groups_first) res = tuple_res[(size_t)p * UAVM_RANSAC_MAX_TUPLES_FIRST + g];
        else { res = eval_group_fast(S.pts, n, seed, (uint32_t)g, thr2, hd); fresh = true; }
        if (res & RES_DEFER) fresh = true;
        warp_resolve_deferred(S.pts, n, seed, (uint32_t)g, thr2, res, hd);   // rare tuples: generic algorithm, whole warp
        if (fresh && g < UAVM_RANSAC_MAX_TUPLES_FIRST && (res & RES_VALID) && !(res & RES_REJ)) {
#pragma unroll
            for (int i = 0; i < 9; i++) tuple_h[((size_t)p * UAVM_RANSAC_MAX_TUPLES_FIRST + g) * 9 + i] = hd[i];
        }
        const int v = (res & RES_VALID) ? 1 : 0;
        const int a = (v && !(res & RES_REJ)) ? 1 : 0;
        const int sup = (int)(res & RES_SUP);
        int tot_v, tot_a;
        const int kx = k_valid + block_excl_scan(v, S.scratch, &tot_v);
        const int tx = t_acc + block_excl_scan(a, S.scratch, &tot_a);
        const bool processed = v && (kx + 1 < 5000) && (tx < sample_times);
        const bool cand = processed && a;
        const bool early = cand && ((float)sup * inv_n > 0.99f);
        const int e_tid = block_reduce_min(early ? tid : kFinThreads, S.scratch);     // first early exit in chunk
        const bool in_range = tid <= e_tid;                                           // tuples after the exit never run
        // first accepted tuple overall (its matrix sits in pProjectMat[0])
        const int fa = block_reduce_min((cand && in_range) ? tid : kFinThreads, S.scratch);
        if (first_acc_g < 0 && fa < kFinThreads) first_acc_g = g0 + fa;
        // first strict maximum: max support, ties -> lowest index
        const int key = (cand && in_range) ? ((sup << 8) | (kFinThreads - 1 - tid)) : -1;
        const int bk = block_reduce_max(key, S.scratch);
        if (bk >= 0) {
            const int bsup = bk >> 8, btid = kFinThreads - 1 - (bk & 255);
            if (bsup > best_sup) {
                best_sup = bsup; best_g = g0 + btid;
                // valid-order index of that tuple: broadcast kx of thread btid
                __syncthreads();
                if (tid == btid) S.flag = kx;
                __syncthreads();
                best_k = S.flag;
            }
        }
        // bookkeeping of realSamTimes / t at loop end
        const int proc_cnt_tot = block_reduce_max((processed && in_range) ? (kx + 1) : 0, S.scratch);
        const int acc_cnt_tot = block_reduce_max((cand && in_range) ? (tx + 1) : 0, S.scratch);
        if (proc_cnt_tot > n_tuples) n_tuples = proc_cnt_tot;
        if (acc_cnt_tot > n_counted) n_counted = acc_cnt_tot;
        k_valid += tot_v; t_acc += tot_a;
        if (e_tid < kFinThreads) done = true;
        if (t_acc >= sample_times || k_valid + 1 >= 5000) done = true;
    }
    R.max_support = best_sup; R.best_tuple = best_k; R.n_tuples = n_tuples; R.n_counted = n_counted;

    // ---------------- winning hypothesis matrix ----------------
    const int hg = best_g >= 0 ? best_g : first_acc_g;
    __syncthreads();                                    // tuple_h written above by other threads of this CTA
    if (tid < 32) {
        float h[9];
#pragma unroll
        for (int i = 0; i < 9; i++) h[i] = 0.0f;
        if (hg >= 0 && hg < UAVM_RANSAC_MAX_TUPLES_FIRST) {          // stored by the first pass or by the replay above
            if (tid == 0) {
#pragma unroll
                for (int i = 0; i < 9; i++) h[i] = tuple_h[((size_t)p * UAVM_RANSAC_MAX_TUPLES_FIRST + hg) * 9 + i];
            }
        } else if (hg >= 0) {                                        // beyond the stored range: evaluate again
            uint32_t r = (tid == 0) ? eval_group_fast(S.pts, n, seed, (uint32_t)hg, -1.0f, h) : 0u;
            warp_resolve_deferred(S.pts, n, seed, (uint32_t)hg, -1.0f, r, h);
        }
        if (tid == 0) {
#pragma unroll
            for (int i = 0; i < 9; i++) S.hbest[i] = h[i];
        }
    }
    __syncthreads();

    // ---------------- final inlier pass, division form (:1922-1944) ----------------
    int cnt = 0;
    for (int i0 = 0; i0 < n; i0 += kFinThreads) {
        const int i = i0 + tid;
        int f = 0;
        float4 pt = make_float4(0, 0, 0, 0);
        if (i < n) {
            pt = S.pts[i];
            float xb, yb;
            project_div(pt.z, pt.w, S.hbest, xb, yb);
            f = dist2(xb, yb, pt.x, pt.y) < thr2 ? 1 : 0;
            inlier[base + i] = (uint8_t)f;
        }
        int tot;
        const int pos = cnt + block_excl_scan(f, S.scratch, &tot);
        if (f) { S.in_x1[pos] = pt.x; S.in_y1[pos] = pt.y; S.in_x2[pos] = pt.z; S.in_y2[pos] = pt.w; }
        cnt += tot;
    }
    __syncthreads();
    R.n_inliers = cnt;
    R.ok = cnt >= 4 ? 1 : 0;
    if (cnt < 4) { if (tid == 0) results[p] = R; return; }

    // ---------------- final refine on all inliers (:1979-1986 -> M/LeastSquare.h:353-531) --------
    // (the SolveHomographyMatrix result of :1958 is overwritten by this refine and is not computed)
    const int rows = 2 * cnt;
    if (tid < 8) S.w[tid] = S.hbest[tid];
    if (tid < 64) S.N2[tid] = 0.0f;      // reference: uninitialised; defined as zeros (see oracle.c)
    __syncthreads();
    for (int it = 0; it < 15; it++) {
        for (int i = tid; i < cnt; i += kFinThreads) {
            const float xs = S.in_x2[i], ys = S.in_y2[i];
            const float d = S.w[6] * xs + S.w[7] * ys + 1.0f;
            const float u = S.w[0] * xs + S.w[1] * ys + S.w[2];
            const float v = S.w[3] * xs + S.w[4] * ys + S.w[5];
            float* j0 = J + (2 * i) * 8; float* j1 = j0 + 8;
            const float a = xs / d, b = ys / d, c = 1.0f / d;
            j0[0] = a; j0[1] = b; j0[2] = c; j0[3] = 0.0f; j0[4] = 0.0f; j0[5] = 0.0f;
            j0[6] = -xs * u / (d * d); j0[7] = -ys * u / (d * d);
            j1[0] = 0.0f; j1[1] = 0.0f; j1[2] = 0.0f; j1[3] = a; j1[4] = b; j1[5] = c;
            j1[6] = -xs * v / (d * d); j1[7] = -ys * v / (d * d);
            S.C[2 * i] = S.in_x1[i] - u / d;
            S.C[2 * i + 1] = S.in_y1[i] - v / d;
        }
        __syncthreads();
        if (tid < 64) {                      // N1 = J^T J, ascending-k accumulation (M/matrix.h:94-120)
            const int r = tid >> 3, c = tid & 7;
            float acc = 0.0f;
            for (int k = 0; k < rows; k++) acc += J[k * 8 + r] * J[k * 8 + c];
            S.N1[tid] = acc;
        }
        __syncthreads();
        if (tid == 0) {                      // InverseMatrix, eps 1e-6, return value ignored (:451)
            float M[8][8];
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int c = 0; c < 8; c++) M[r][c] = S.N1[r * 8 + c];
            if (inverse8_fast(M, 1e-6f)) {
#pragma unroll
                for (int r = 0; r < 8; r++)
#pragma unroll
                    for (int c = 0; c < 8; c++) S.N2[r * 8 + c] = M[r][c];
            } else {
                float tmp[64];
                if (inverse8_generic(S.N1, tmp, 1e-6f) == 1)
                    for (int i = 0; i < 64; i++) S.N2[i] = tmp[i];
            }
        }
        __syncthreads();
        for (int e = tid; e < 8 * rows; e += kFinThreads) {   // L = N2 * J^T  (8 x rows)
            const int r = e / rows, k = e - r * rows;
            float acc = 0.0f;
#pragma unroll
            for (int m = 0; m < 8; m++) acc += S.N2[r * 8 + m] * J[k * 8 + m];
            L[e] = acc;
        }
        __syncthreads();
        if (tid < 8) {                       // delta = L * C, ascending k
            float acc = 0.0f;
            const float* Lr = L + tid * rows;
            for (int k = 0; k < rows; k++) acc += Lr[k] * S.C[k];
            S.dx[tid] = acc;
        }
        __syncthreads();
        if (tid == 0) {
            bool small = true;
            for (int i = 0; i < 8; i++) { S.w[i] += S.dx[i]; if (!(fabsf(S.dx[i]) < 1e-10f)) small = false; }
            S.flag = small ? 1 : 0;
        }
        __syncthreads();
        if (S.flag) break;
    }
    // out[8] = max residual (float, reciprocal-multiply projection, sqrtf)  (:501-519)
    float emax = 0.0f;
    for (int i = tid; i < cnt; i += kFinThreads) {
        float xb, yb;
        project_mul(S.in_x2[i], S.in_y2[i], S.w, xb, yb);
        const float ex = S.in_x1[i] - xb, ey = S.in_y1[i] - yb;
        const float dist = sqrtf(ex * ex + ey * ey);
        if (dist > emax) emax = dist;
    }
    const int emax_bits = block_reduce_max(__float_as_int(emax), S.scratch);   // emax >= 0: int order == float order
    if (tid == 0) {
        for (int i = 0; i < 8; i++) R.H[i] = S.w[i];
        R.H[8] = __int_as_float(emax_bits);
        results[p] = R;
    }
}

constexpr size_t kFinSmemBytes = sizeof(FinSmem) + (size_t)2 * UAVM_CAND_SLOTS * 8 * 4 * 2;

}  // namespace

int uavm_launch_ransac(uavm_ctx* ctx, uavm_pairbatch* pb, float dist, int sample_times)
{
    if (sample_times > 5000) sample_times = 5000;            // maxTimes clamp (:1765-1769)
    if (sample_times < 0) sample_times = 0;
    const float thr2 = dist * dist;                          // fRansacDistSquare (:1757)
    // first pass: enough draw groups that, at the ~35 % gate-rejection rate seen on 4000x3000 data, the
    // 1000th accepted tuple is usually inside it; the finalize kernel continues past it when needed.
    int groups = sample_times * 9 / 4 + 64;     // ~2115 +- 50 groups are consumed per 1000 counted tuples at 4000x3000
    if (groups > UAVM_RANSAC_MAX_TUPLES_FIRST) groups = UAVM_RANSAC_MAX_TUPLES_FIRST;
    groups = ((groups + kGroupsPerBlock - 1) / kGroupsPerBlock) * kGroupsPerBlock;
    if (groups > UAVM_RANSAC_MAX_TUPLES_FIRST) groups = UAVM_RANSAC_MAX_TUPLES_FIRST;
    UAVM_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->k4_fin_attr_set) {                          // function attributes are per device: tracked per context
        UAVM_CUDA(ctx, cudaFuncSetAttribute(k4_ransac_finalize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFinSmemBytes));
        ctx->k4_fin_attr_set = true;
    }
    if (groups > 0) {
        dim3 grid(groups / kGroupsPerBlock, pb->n_pairs);
        // Residency knob (blocks per SM, via dummy dynamic shared memory).  Measured on B200 with RANSAC on the
        // high-priority side stream next to the warp kernel: 3 blocks/SM (no padding) gives the best step time
        // (3.20 ms vs 3.50 at 2 and 3.53 at 1), so the default stays 3; the knob is kept for experiments.
        size_t pad_smem = 0;
        int bpsm = 3;
        if (getenv("UAVM_RANSAC_EVAL_BPSM")) bpsm = atoi(getenv("UAVM_RANSAC_EVAL_BPSM"));
        if (bpsm == 1) pad_smem = 100 * 1024; else if (bpsm == 2) pad_smem = 60 * 1024;
        if (pad_smem > ctx->k4_eval_attr_smem) {
            UAVM_CUDA(ctx, cudaFuncSetAttribute(k4_ransac_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad_smem));
            ctx->k4_eval_attr_smem = pad_smem;
        }
        k4_ransac_eval<<<grid, kEvalThreads, pad_smem, ctx->stream>>>(pb->d_pairs, pb->d_cand_xy1, pb->d_cand_xy2, pb->d_cand_n,
                                                               thr2, groups, pb->d_tuple_res, pb->d_tuple_h);
        UAVM_CHECK_LAUNCH(ctx);
    }
    k4_ransac_finalize<<<pb->n_pairs, kFinThreads, kFinSmemBytes, ctx->stream>>>(
        pb->d_pairs, pb->d_cand_xy1, pb->d_cand_xy2, pb->d_cand_n, thr2, sample_times, groups, pb->d_tuple_res,
        pb->d_tuple_h, pb->d_inlier, pb->d_res);
    UAVM_CHECK_LAUNCH(ctx);
    return UAVM_OK;
}
