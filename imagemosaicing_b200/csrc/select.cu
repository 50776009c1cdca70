// select.cu — K3: candidate selection.  Replaces std::sort(matches) + nMatch = Min(400, 0.3*size) +
// SelectMatchPairs (M/MosaicWithoutPos.cpp:5111, :5146-5153, :4977-5028).
//
// The reference walks ALL matches in ascending distance order and keeps a match while the 3x3 grid
// cell of its image-1 keypoint holds fewer than quota = int(nMatch/9) matches.  That is: every cell
// keeps its `quota` best matches, and the output is those matches in global key order.  So instead of
// sorting 8k matches we (1) radix-select, for every cell at once, the quota-th smallest key,
// (2) gather the <= cells*quota survivors, (3) bitonic-sort only the survivors.
// Order key = (d2 << 20) | queryIdx  — the (distance, queryIdx) order fixed by the oracle (std::sort on
// equal distances is unspecified in the reference).  Grid-cell quirk (label index may alias / overrun,
// :4993-5010) is kept: counters exist for every index gridX*nY+nX the walk can produce.
// One CTA per pair.
#include "internal.h"

namespace {

constexpr int kSelThreads = 1024;
constexpr int kMaxCells = 32;
constexpr int kMaxSel = 1024;          // survivors sorted in shared memory
constexpr int kIdxBits = 20;

struct SelParams {
    int width, height, grid_x, grid_y, n_label;
    int step_x, step_y;
    int max_num; double frac;
};

__device__ __forceinline__ int cell_of(float x, float y, const SelParams& P) {
    // int nX = temp1.x / stepX (float / int -> float division, truncation), :5006-5007
    int nx = (int)__fdiv_rn(x, (float)P.step_x);
    int ny = (int)__fdiv_rn(y, (float)P.step_y);
    return P.grid_x * ny + nx;
}

__global__ void __launch_bounds__(kSelThreads, 1)
k3_select(const PairDesc* __restrict__ pairs, const int32_t* __restrict__ train_idx, const int32_t* __restrict__ d2,
          const float* __restrict__ kp, SelParams P,
          float* __restrict__ cand_xy1, float* __restrict__ cand_xy2, int32_t* __restrict__ cand_id1,
          int32_t* __restrict__ cand_id2, int32_t* __restrict__ cand_n)
{
    __shared__ uint32_t hist[kMaxCells][256];
    __shared__ unsigned long long prefix[kMaxCells];
    __shared__ int remaining[kMaxCells];
    __shared__ int cell_count[kMaxCells];
    __shared__ unsigned long long thresh[kMaxCells];
    __shared__ unsigned long long sel[kMaxSel];
    __shared__ int n_sel;

    const PairDesc pd = pairs[blockIdx.x];
    const int n = pd.nq;
    const int tid = threadIdx.x;
    // an empty train image has no matches (trainIdx stays -1): the reference's matcher returns an empty list and the pair
    // yields no candidates (M/MosaicWithoutPos.cpp:5108-5111)
    if (pd.nt <= 0 || n <= 0) { if (tid == 0) cand_n[blockIdx.x] = 0; return; }
    const int32_t* tix = train_idx + pd.match_off;
    const int32_t* dd = d2 + pd.match_off;
    const float* kp1 = kp + (size_t)pd.q_row * 2;
    const float* kp2 = kp + (size_t)pd.t_row * 2;

    // nMatch = Min(maxNum, frac*n) (double, then int); quota = int((float)nMatch / nGrids)
    const double nm_d = P.frac * (double)n;
    const int n_match = (int)(((double)P.max_num < nm_d) ? (double)P.max_num : nm_d);
    const int quota = (int)((float)n_match / (float)(P.grid_x * P.grid_y));

    if (tid < kMaxCells) { cell_count[tid] = 0; prefix[tid] = 0ull; }
    if (tid == 0) n_sel = 0;
    __syncthreads();
    for (int q = tid; q < n; q += kSelThreads) {
        int c = cell_of(kp1[2 * q], kp1[2 * q + 1], P);
        if (c >= 0 && c < P.n_label) atomicAdd(&cell_count[c], 1);
    }
    __syncthreads();
    if (tid < kMaxCells) {
        remaining[tid] = quota;                     // rank (1-based) of the threshold element
        // cells holding <= quota matches keep everything (threshold = max key)
        thresh[tid] = (tid < P.n_label && cell_count[tid] > quota && quota > 0) ? 0ull : ~0ull;
        if (quota <= 0) thresh[tid] = 0ull;         // nothing is ever taken; handled below
    }
    __syncthreads();

    // radix select, 8 bits per pass, most significant digit first (key < 2^(23+20) -> 6 passes)
    for (int shift = 40; shift >= 0; shift -= 8) {
        for (int i = tid; i < kMaxCells * 256; i += kSelThreads) (&hist[0][0])[i] = 0;
        __syncthreads();
        for (int q = tid; q < n; q += kSelThreads) {
            int c = cell_of(kp1[2 * q], kp1[2 * q + 1], P);
            if (c < 0 || c >= P.n_label) continue;
            if (thresh[c] != 0ull || quota <= 0) continue;              // cell not under selection
            unsigned long long key = ((unsigned long long)(uint32_t)dd[q] << kIdxBits) | (unsigned long long)q;
            if ((key >> (shift + 8)) == prefix[c]) atomicAdd(&hist[c][(key >> shift) & 255], 1u);
        }
        __syncthreads();
        if (tid < P.n_label && thresh[tid] == 0ull && quota > 0) {
            int rem = remaining[tid];
            int digit = 0; uint32_t cum = 0;
            for (int b = 0; b < 256; b++) {
                uint32_t h = hist[tid][b];
                if (cum + h >= (uint32_t)rem) { digit = b; break; }
                cum += h;
            }
            remaining[tid] = rem - (int)cum;
            prefix[tid] = (prefix[tid] << 8) | (unsigned long long)digit;
        }
        __syncthreads();
    }
    if (tid < P.n_label && thresh[tid] == 0ull && quota > 0) thresh[tid] = prefix[tid];
    __syncthreads();

    // gather survivors
    if (quota > 0) {
        for (int q = tid; q < n; q += kSelThreads) {
            int c = cell_of(kp1[2 * q], kp1[2 * q + 1], P);
            if (c < 0 || c >= P.n_label) continue;
            unsigned long long key = ((unsigned long long)(uint32_t)dd[q] << kIdxBits) | (unsigned long long)q;
            if (key <= thresh[c]) {
                int slot = atomicAdd(&n_sel, 1);
                if (slot < kMaxSel) sel[slot] = key;
            }
        }
    }
    __syncthreads();
    int ns = n_sel < kMaxSel ? n_sel : kMaxSel;
    for (int i = tid; i < kMaxSel; i += kSelThreads) if (i >= ns) sel[i] = ~0ull;
    __syncthreads();
    // bitonic sort of kMaxSel keys, one element pair per thread per step
    for (int k = 2; k <= kMaxSel; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            int i = tid;
            int ixj = i ^ j;
            if (ixj > i && i < kMaxSel) {
                unsigned long long a = sel[i], b = sel[ixj];
                bool up = ((i & k) == 0);
                if ((a > b) == up) { sel[i] = b; sel[ixj] = a; }
            }
            __syncthreads();
        }
    }
    const int n_out = ns < UAVM_CAND_SLOTS ? ns : UAVM_CAND_SLOTS;
    if (tid == 0) cand_n[blockIdx.x] = n_out;
    const size_t base = (size_t)blockIdx.x * UAVM_CAND_SLOTS;
    for (int r = tid; r < n_out; r += kSelThreads) {
        int q = (int)(sel[r] & ((1ull << kIdxBits) - 1));
        int t = tix[q];
        cand_xy1[(base + r) * 2] = kp1[2 * q]; cand_xy1[(base + r) * 2 + 1] = kp1[2 * q + 1];
        cand_xy2[(base + r) * 2] = kp2[2 * t]; cand_xy2[(base + r) * 2 + 1] = kp2[2 * t + 1];
        cand_id1[base + r] = q; cand_id2[base + r] = t;
    }
}

}  // namespace

int uavm_launch_select(uavm_ctx* ctx, uavm_pairbatch* pb, int width, int height, int gx, int gy, int max_num, double frac)
{
    if (gx <= 0 || gy <= 0 || width < gx || height < gy || max_num <= 0) return UAVM_EINVAL;
    SelParams P;
    P.width = width; P.height = height; P.grid_x = gx; P.grid_y = gy;
    P.n_label = gx * (gy + 1) + gx + 1;
    P.step_x = width / gx; P.step_y = height / gy;
    P.max_num = max_num; P.frac = frac;
    if (P.n_label > kMaxCells) { UAVM_SET_ERR(ctx, "grid too large"); return UAVM_EINVAL; }
    if ((int64_t)P.n_label * (max_num / (gx * gy)) > UAVM_CAND_SLOTS) { UAVM_SET_ERR(ctx, "max_num too large for the candidate slots"); return UAVM_EINVAL; }
    if (pb->max_nq >= (1 << kIdxBits)) { UAVM_SET_ERR(ctx, "too many keypoints per image"); return UAVM_EINVAL; }
    k3_select<<<pb->n_pairs, kSelThreads, 0, ctx->stream>>>(pb->d_pairs, pb->d_train_idx, pb->d_d2, pb->fs->d_kp, P,
                                                            pb->d_cand_xy1, pb->d_cand_xy2, pb->d_cand_id1,
                                                            pb->d_cand_id2, pb->d_cand_n);
    UAVM_CHECK_LAUNCH(ctx);
    return UAVM_OK;
}
