// align_host.cpp — host-side stages between RANSAC and warp:
//   uavm_connected_images  <- Select_Connected_Matched_Images + ClusterMatchNode (M/MosaicWithoutPos.cpp:2673-2796)
//   uavm_align_affine      <- BundleAdjustmentSparse (:6971-7202) + SolveSparseSystem2 (M/test_cholmod.cpp:180-262)
// Both are tiny (<= 6(N-1) unknowns) and stay on the host, as SURVEY §8e says ("replicas only"/host).
// The reference solves x = (A^T A)^-1 A^T b with CHOLMOD in double; here the normal equations are
// accumulated directly (double) and solved with a dense Cholesky — same system, same precision class.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <thread>
#include <vector>
#include "internal.h"

extern "C" int uavm_connected_images(const uavm_matchpointpairs* pairs, int n_pairs, int n_images, int32_t* label)
{
    if (!label || n_images <= 0 || n_pairs < 0 || (n_pairs > 0 && !pairs)) return UAVM_EINVAL;
    for (int i = 0; i < n_images; i++) label[i] = 0;
    // adjacency (r = min, c = max) scanned row-major -> edge list (:2756-2785)
    std::vector<uint8_t> node((size_t)n_images * n_images, 0);
    for (int i = 0; i < n_pairs; i++) {
        int a = pairs[i].ptA_i, b = pairs[i].ptB_i;
        if (a < 0 || a >= n_images || b < 0 || b >= n_images) return UAVM_EINVAL;
        int r = a < b ? a : b, c = a < b ? b : a;
        node[(size_t)r * n_images + c] = 255;
    }
    const int words = (n_images + 63) / 64;
    std::vector<std::vector<uint64_t>> clusters;
    for (int r = 0; r < n_images; r++)
        for (int c = 0; c < n_images; c++)
            if (node[(size_t)r * n_images + c]) {
                std::vector<uint64_t> s(words, 0);
                s[r >> 6] |= 1ull << (r & 63); s[c >> 6] |= 1ull << (c & 63);
                clusters.push_back(s);
            }
    if (clusters.empty()) return UAVM_OK;
    // ClusterMatchNode's merge loop with its exact control flow (merge j into i, move the last cluster
    // into slot j, repeat passes until nothing merges); sets as bitmasks instead of index lists.
    bool over;
    do {
        over = true;
        for (size_t i = 0; i < clusters.size(); i++) {
            for (size_t j = i + 1; j < clusters.size();) {
                bool share = false;
                for (int w = 0; w < words; w++) if (clusters[i][w] & clusters[j][w]) { share = true; break; }
                if (share) {
                    over = false;
                    for (int w = 0; w < words; w++) clusters[i][w] |= clusters[j][w];
                    clusters[j] = clusters.back();
                    clusters.pop_back();
                } else j++;
            }
        }
    } while (!over);
    size_t imax = 0; int maxval = 0;
    for (size_t i = 0; i < clusters.size(); i++) {
        int cnt = 0;
        for (int w = 0; w < words; w++) cnt += __builtin_popcountll(clusters[i][w]);
        if (cnt > maxval) { maxval = cnt; imax = i; }
    }
    for (int k = 0; k < n_images; k++)
        if (clusters[imax][k >> 6] >> (k & 63) & 1ull) label[k] = 1;
    return UAVM_OK;
}

// Dense double Cholesky restricted to the ENVELOPE of the normal matrix (first non-zero column of every row): the pair graph
// of a strip or block couples an image only with nearby images, so rows are short (a 200-image block: 1194 unknowns, rows of
// <= ~150 entries) and the factorisation costs n * row^2 instead of n^3 / 3.  Fill-in stays inside the envelope and every
// skipped term is an exact zero, so the result is bit-identical to the full dense factorisation.
// The affine model decouples: x' = a x + b y + e and y' = c x + d y + f have the SAME normal matrix (built from the same
// [x y 1] rows, M/MosaicWithoutPos.cpp:7035-7146) and no cross terms, so one factorisation of order 3 (N - 1) serves both
// right-hand sides.  In the interleaved 6 (N - 1) system the cross entries are exact zeros and every non-zero term is
// accumulated and eliminated in the same order, so this is again bit-identical — at a quarter of the work.
static int cholesky_solve2(std::vector<double>& N, std::vector<double>& b1, std::vector<double>& b2, int n)
{
    std::vector<int> first(n);
    for (int i = 0; i < n; i++) {
        int f = 0;
        while (f < i && N[(size_t)i * n + f] == 0.0) f++;
        first[i] = f;
    }
    for (int j = 0; j < n; j++) {
        double s = N[(size_t)j * n + j];
        for (int k = first[j]; k < j; k++) s -= N[(size_t)j * n + k] * N[(size_t)j * n + k];
        if (!(s > 0)) return -1;
        const double l = sqrt(s);
        N[(size_t)j * n + j] = l;
        for (int i = j + 1; i < n; i++) {
            if (first[i] > j) continue;                                   // N[i][j] is and stays 0
            double t = N[(size_t)i * n + j];
            for (int k = first[i] > first[j] ? first[i] : first[j]; k < j; k++) t -= N[(size_t)i * n + k] * N[(size_t)j * n + k];
            N[(size_t)i * n + j] = t / l;
        }
    }
    for (std::vector<double>* pb : {&b1, &b2}) {
        std::vector<double>& b = *pb;
        for (int i = 0; i < n; i++) {
            double t = b[i];
            for (int k = first[i]; k < i; k++) t -= N[(size_t)i * n + k] * b[k];
            b[i] = t / N[(size_t)i * n + i];
        }
        for (int i = n - 1; i >= 0; i--) {
            double t = b[i];
            for (int k = i + 1; k < n; k++) if (first[k] <= i) t -= N[(size_t)k * n + i] * b[k];
            b[i] = t / N[(size_t)i * n + i];
        }
    }
    return 0;
}

extern "C" int uavm_align_affine(const uavm_matchpointpairs* pairs, int n_pairs, const uavm_imagetransform* init, int n_images,
                                 int n_fixed, uavm_imagetransform* out)
{
    if (!pairs || n_pairs <= 0 || !init || !out) return UAVM_EINVAL;
    if (n_images <= 1) return UAVM_EFAIL;                 // reference returns -2 (:6978-6981)
    std::vector<int> acc_fixed(n_images, 0);
    int nf = 0;
    for (int i = 0; i < n_images; i++) { acc_fixed[i] = nf; if (init[i].fixed == 1) nf++; }
    if (n_fixed >= 0 && n_fixed != nf) return UAVM_EINVAL;
    const int nu = 3 * (n_images - nf);                   // per coordinate: (a, b, e) resp. (c, d, f) of every free image
    if (nu <= 0) return UAVM_EFAIL;
    if (nu > 12000) return UAVM_EFAIL;                    // dense normal matrix limit (1.1 GB)
    const auto t_begin = std::chrono::steady_clock::now();
    std::vector<double> N((size_t)nu * nu, 0.0), gx(nu, 0.0), gy(nu, 0.0);
    // The matches of one image pair arrive as a run (uavm_pairbatch_collect / _allgather emit pair after pair) and all touch the
    // same <= 6 x 6 block of N.  Runs are independent: each is summed (match by match, in order) into 21 + 12 local accumulators,
    // several runs at a time on host threads, and the blocks are then added to N in run order — the result does not depend on
    // the number of threads.  Summation order differs from adding match by match into N, so the result equals the dense oracle
    // to rounding (1e-12 relative), not bit for bit.
    struct Run { int begin, end, nc; int cols[6]; double loc[6][6], gx[6], gy[6]; };
    auto classify = [&](const uavm_matchpointpairs& m, int* cols) -> int {       // columns of the match's unknown blocks; -1: contributes nothing
        if (m.ptA_Fixed == 0 && m.ptB_Fixed == 0) {
            const int ca = 3 * (m.ptA_i - acc_fixed[m.ptA_i]), cb = 3 * (m.ptB_i - acc_fixed[m.ptB_i]);
            cols[0] = cols[1] = cols[2] = ca; cols[3] = cols[4] = cols[5] = cb; return 6;
        }
        if (m.ptA_Fixed == 1 && m.ptB_Fixed == 0) { cols[0] = cols[1] = cols[2] = 3 * (m.ptB_i - acc_fixed[m.ptB_i]); return 3; }
        if (m.ptA_Fixed == 0 && m.ptB_Fixed == 1) { cols[0] = cols[1] = cols[2] = 3 * (m.ptA_i - acc_fixed[m.ptA_i]); return 3; }
        return -1;
    };
    // fixed chunks of the match list (a run that crosses a chunk edge is summed as two blocks — at the same places whatever the
    // number of threads), one pass per chunk: find the runs and sum them
    const auto t_alloc = std::chrono::steady_clock::now();
    constexpr int kChunk = 8192;
    const int n_chunks = (n_pairs + kChunk - 1) / kChunk;
    std::vector<std::vector<Run>> chunk_runs(n_chunks);
    std::vector<int> chunk_bad(n_chunks, 0);
    auto do_chunk = [&](int c) {
        std::vector<Run>& runs = chunk_runs[c];
        const int n0 = c * kChunk, n1 = std::min(n_pairs, n0 + kChunk);
        for (int n = n0; n < n1; n++) {
            const uavm_matchpointpairs& m = pairs[n];
            if (m.ptA_i < 0 || m.ptA_i >= n_images || m.ptB_i < 0 || m.ptB_i >= n_images) { chunk_bad[c] = 1; return; }
            int cols[6] = {0, 0, 0, 0, 0, 0};
            const int nc = classify(m, cols);
            if (nc < 0) continue;
            for (int a = 0; a < nc; a++)
                if (cols[a] < 0 || cols[a] + 2 >= nu) { chunk_bad[c] = 1; return; }      // a "free" point on a fixed image
            bool same = false;
            if (!runs.empty()) {
                const Run& r = runs.back();
                const uavm_matchpointpairs& p = pairs[r.begin];
                same = r.end == n && r.nc == nc && p.ptA_i == m.ptA_i && p.ptB_i == m.ptB_i && p.ptA_Fixed == m.ptA_Fixed && p.ptB_Fixed == m.ptB_Fixed;
            }
            if (!same) {
                Run r; memset(&r, 0, sizeof(r));
                r.begin = n; r.nc = nc; memcpy(r.cols, cols, sizeof(cols));
                runs.push_back(r);
            }
            Run& r = runs.back();
            r.end = n + 1;
            double vals[6]; double rx = 0, ry = 0;
            const double xa = m.ptA.x, ya = m.ptA.y, xb = m.ptB.x, yb = m.ptB.y;
            if (nc == 6) { vals[0] = xa; vals[1] = ya; vals[2] = 1; vals[3] = -xb; vals[4] = -yb; vals[5] = -1; }
            else if (m.ptA_Fixed == 1) {
                vals[0] = xb; vals[1] = yb; vals[2] = 1;
                double h[9]; for (int t = 0; t < 9; t++) h[t] = init[m.ptA_i].h.m[t];
                rx = (h[0] * xa + h[1] * ya + h[2]) / (h[6] * xa + h[7] * ya + h[8]);      // ApplyProject9 (M/MosaicWithoutPos.h:331-336)
                ry = (h[3] * xa + h[4] * ya + h[5]) / (h[6] * xa + h[7] * ya + h[8]);
            } else {
                vals[0] = xa; vals[1] = ya; vals[2] = 1;
                double h[9]; for (int t = 0; t < 9; t++) h[t] = init[m.ptB_i].h.m[t];
                rx = (h[0] * xb + h[1] * yb + h[2]) / (h[6] * xb + h[7] * yb + h[8]);
                ry = (h[3] * xb + h[4] * yb + h[5]) / (h[6] * xb + h[7] * yb + h[8]);
            }
            for (int a = 0; a < nc; a++) {
                for (int b = 0; b <= a; b++) r.loc[a][b] += vals[a] * vals[b];          // symmetric: lower triangle of the block only
                r.gx[a] += vals[a] * rx;
                r.gy[a] += vals[a] * ry;
            }
        }
    };
    {
        unsigned hw = std::thread::hardware_concurrency();
        int T = std::min<int>(std::min<unsigned>(hw ? hw : 1, 8), n_chunks);
        if (getenv("UAVM_ALIGN_THREADS")) T = std::max(1, std::min(atoi(getenv("UAVM_ALIGN_THREADS")), n_chunks));
        if (T <= 1) { for (int c = 0; c < n_chunks; c++) do_chunk(c); }
        else {
            std::vector<std::thread> th;
            for (int t = 0; t < T; t++)
                th.emplace_back([&, t]() { for (int c = t; c < n_chunks; c += T) do_chunk(c); });
            for (auto& x : th) x.join();
        }
    }
    const auto t_par = std::chrono::steady_clock::now();
    for (int c = 0; c < n_chunks; c++) if (chunk_bad[c]) return UAVM_EINVAL;
    for (int c = 0; c < n_chunks; c++)                                            // blocks into N in list order: independent of T
        for (const Run& r : chunk_runs[c])
            for (int a = 0; a < r.nc; a++) {
                const int ar = r.cols[a] + a % 3;
                for (int b = 0; b < r.nc; b++) {
                    const int bc = r.cols[b] + b % 3;
                    if (bc <= ar) N[(size_t)ar * nu + bc] += b <= a ? r.loc[a][b] : r.loc[b][a];   // the factorisation reads the lower triangle only
                }
                gx[ar] += r.gx[a]; gy[ar] += r.gy[a];
            }
    const auto t_acc = std::chrono::steady_clock::now();
    if (cholesky_solve2(N, gx, gy, nu) != 0) return UAVM_EFAIL;
    if (getenv("UAVM_ALIGN_TIMING")) fprintf(stderr, "align: alloc %.3f ms, runs %.3f ms, accumulate %.3f ms, solve %.3f ms (nu = %d)\n", std::chrono::duration<double, std::milli>(t_alloc - t_begin).count(), std::chrono::duration<double, std::milli>(t_par - t_alloc).count(), std::chrono::duration<double, std::milli>(t_acc - t_begin).count(), std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_acc).count(), nu);
    int k = 0;
    for (int i = 0; i < n_images; i++) {
        if (init[i].fixed == 0) {
            uavm_imagetransform t; memset(&t, 0, sizeof(t));
            t.fixed = 0;
            t.h.m[0] = (float)gx[3 * k + 0]; t.h.m[1] = (float)gx[3 * k + 1]; t.h.m[2] = (float)gx[3 * k + 2];
            t.h.m[3] = (float)gy[3 * k + 0]; t.h.m[4] = (float)gy[3 * k + 1]; t.h.m[5] = (float)gy[3 * k + 2];
            t.h.m[6] = 0; t.h.m[7] = 0; t.h.m[8] = 1;
            out[i] = t; k++;
        } else out[i] = init[i];
    }
    return UAVM_OK;
}


// ------------------------------------------------------------------------------------------------------------------------------
// f3: the constrained variants of the global alignment (dead code in the shipped reference — `methodType = 0`,
// M/MosaicWithoutPos.cpp:4588 — but part of the paper's method and of SURVEY §8 f3):
//   uavm_align_affine_constrained  <- BundleAdjustmentSparseConstraint (:6032-6300): the affine least squares plus, per free image,
//                                     the soft similarity constraints nC (a - d) = 0 and nC (b + c) = 0, nC = number of its
//                                     non-fixed match points (the reference pushes the same row nC times, which sums to nC)
//   uavm_align_affine_rot          <- SparseAffineRotConstraint (:6302-6808): starts from the result above and runs 10 Gauss-Newton
//                                     steps on the match residuals plus, per free image, w (a b + c d), w (a^2 + c^2 - 1),
//                                     w (b^2 + d^2 - 1) with w = int(nC * weight) (Jacobian rows and residuals both scaled by w,
//                                     as in the reference), each step one sparse Cholesky solve (SolveSparseSystem2).
// Unknown order per free image: a b c d e f with x' = a x + b y + e, y' = c x + d y + f (:6390-6400).  The systems couple a..d, so
// the x / y decoupling of uavm_align_affine does not apply; the envelope Cholesky above is used on 6 (N - nFixed) unknowns.
namespace {

struct MatchSystem {
    int nu = 0;
    std::vector<double> N, g;            // A^T A (full symmetric storage) and A^T b of the match rows
    std::vector<int> acc_fixed, freq;    // images fixed before i; non-fixed match points per image
};

int build_match_system(const uavm_matchpointpairs* pairs, int n_pairs, const uavm_imagetransform* init, int n_images, int n_fixed, MatchSystem& S)
{
    S.acc_fixed.assign(n_images, 0); S.freq.assign(n_images, 0);
    int nf = 0;
    for (int i = 0; i < n_images; i++) { S.acc_fixed[i] = nf; if (init[i].fixed == 1) nf++; }
    if (n_fixed >= 0 && n_fixed != nf) return UAVM_EINVAL;
    S.nu = 6 * (n_images - nf);
    if (S.nu <= 0 || S.nu > 12000) return UAVM_EFAIL;
    const int nu = S.nu;
    S.N.assign((size_t)nu * nu, 0.0); S.g.assign(nu, 0.0);
    static const int slot_x[3] = {0, 1, 4}, slot_y[3] = {2, 3, 5};
    for (int n = 0; n < n_pairs; n++) {
        const uavm_matchpointpairs& m = pairs[n];
        if (m.ptA_i < 0 || m.ptA_i >= n_images || m.ptB_i < 0 || m.ptB_i >= n_images) return UAVM_EINVAL;
        if (m.ptA_Fixed == 0) S.freq[m.ptA_i]++;
        if (m.ptB_Fixed == 0) S.freq[m.ptB_i]++;
        int cols[6]; double vals[6]; int nc = 0; double rx = 0, ry = 0;
        const double xa = m.ptA.x, ya = m.ptA.y, xb = m.ptB.x, yb = m.ptB.y;
        if (m.ptA_Fixed == 0 && m.ptB_Fixed == 0) {
            const int ca = 6 * (m.ptA_i - S.acc_fixed[m.ptA_i]), cb = 6 * (m.ptB_i - S.acc_fixed[m.ptB_i]);
            cols[0] = cols[1] = cols[2] = ca; vals[0] = xa; vals[1] = ya; vals[2] = 1;
            cols[3] = cols[4] = cols[5] = cb; vals[3] = -xb; vals[4] = -yb; vals[5] = -1; nc = 6;
        } else if (m.ptA_Fixed == 1 && m.ptB_Fixed == 0) {
            const int cb = 6 * (m.ptB_i - S.acc_fixed[m.ptB_i]);
            cols[0] = cols[1] = cols[2] = cb; vals[0] = xb; vals[1] = yb; vals[2] = 1; nc = 3;
            double h[9]; for (int t = 0; t < 9; t++) h[t] = init[m.ptA_i].h.m[t];
            rx = (h[0] * xa + h[1] * ya + h[2]) / (h[6] * xa + h[7] * ya + h[8]);
            ry = (h[3] * xa + h[4] * ya + h[5]) / (h[6] * xa + h[7] * ya + h[8]);
        } else if (m.ptA_Fixed == 0 && m.ptB_Fixed == 1) {
            const int ca = 6 * (m.ptA_i - S.acc_fixed[m.ptA_i]);
            cols[0] = cols[1] = cols[2] = ca; vals[0] = xa; vals[1] = ya; vals[2] = 1; nc = 3;
            double h[9]; for (int t = 0; t < 9; t++) h[t] = init[m.ptB_i].h.m[t];
            rx = (h[0] * xb + h[1] * yb + h[2]) / (h[6] * xb + h[7] * yb + h[8]);
            ry = (h[3] * xb + h[4] * yb + h[5]) / (h[6] * xb + h[7] * yb + h[8]);
        } else continue;
        for (int a = 0; a < nc; a++) {
            if (cols[a] < 0 || cols[a] + 5 >= nu) return UAVM_EINVAL;
            const int ax = cols[a] + slot_x[a % 3], ay = cols[a] + slot_y[a % 3];
            for (int b = 0; b < nc; b++) {
                S.N[(size_t)ax * nu + cols[b] + slot_x[b % 3]] += vals[a] * vals[b];
                S.N[(size_t)ay * nu + cols[b] + slot_y[b % 3]] += vals[a] * vals[b];
            }
            S.g[ax] += vals[a] * rx;
            S.g[ay] += vals[a] * ry;
        }
    }
    return UAVM_OK;
}

void unpack6(const std::vector<double>& x, const uavm_imagetransform* init, int n_images, uavm_imagetransform* out)
{
    int k = 0;
    for (int i = 0; i < n_images; i++) {
        if (init[i].fixed == 0) {
            uavm_imagetransform t; memset(&t, 0, sizeof(t));
            t.h.m[0] = (float)x[6 * k + 0]; t.h.m[1] = (float)x[6 * k + 1]; t.h.m[3] = (float)x[6 * k + 2]; t.h.m[4] = (float)x[6 * k + 3];
            t.h.m[2] = (float)x[6 * k + 4]; t.h.m[5] = (float)x[6 * k + 5]; t.h.m[8] = 1;
            out[i] = t; k++;
        } else out[i] = init[i];
    }
}

int solve_sym(std::vector<double>& N, std::vector<double>& b, int n)
{
    std::vector<double> dummy(n, 0.0);
    return cholesky_solve2(N, b, dummy, n);
}

}  // namespace

extern "C" int uavm_align_affine_constrained(const uavm_matchpointpairs* pairs, int n_pairs, const uavm_imagetransform* init, int n_images,
                                             int n_fixed, uavm_imagetransform* out)
{
    if (!pairs || n_pairs <= 0 || !init || !out) return UAVM_EINVAL;
    if (n_images <= 1) return UAVM_EFAIL;
    MatchSystem S;
    int rc = build_match_system(pairs, n_pairs, init, n_images, n_fixed, S);
    if (rc != UAVM_OK) return rc;
    const int nu = S.nu;
    for (int i = 0; i < n_images; i++) {                  // nC (a - d) = 0, nC (b + c) = 0 (:6228-6241)
        if (init[i].fixed == 1) continue;
        const int u = 6 * (i - S.acc_fixed[i]);
        const double w2 = (double)S.freq[i] * (double)S.freq[i];
        S.N[(size_t)(u + 0) * nu + u + 0] += w2; S.N[(size_t)(u + 3) * nu + u + 3] += w2;
        S.N[(size_t)(u + 0) * nu + u + 3] -= w2; S.N[(size_t)(u + 3) * nu + u + 0] -= w2;
        S.N[(size_t)(u + 1) * nu + u + 1] += w2; S.N[(size_t)(u + 2) * nu + u + 2] += w2;
        S.N[(size_t)(u + 1) * nu + u + 2] += w2; S.N[(size_t)(u + 2) * nu + u + 1] += w2;
    }
    std::vector<double> x(S.g);
    if (solve_sym(S.N, x, nu) != 0) return UAVM_EFAIL;
    unpack6(x, init, n_images, out);
    return UAVM_OK;
}

extern "C" int uavm_align_affine_rot(const uavm_matchpointpairs* pairs, int n_pairs, const uavm_imagetransform* init, int n_images,
                                     int n_fixed, float weight, int iterations, uavm_imagetransform* out)
{
    if (!pairs || n_pairs <= 0 || !init || !out || iterations < 0) return UAVM_EINVAL;
    if (n_images <= 1) return UAVM_EFAIL;
    int rc = uavm_align_affine_constrained(pairs, n_pairs, init, n_images, n_fixed, out);        // initial value (:6374)
    if (rc != UAVM_OK) return rc;
    MatchSystem S;
    rc = build_match_system(pairs, n_pairs, init, n_images, n_fixed, S);
    if (rc != UAVM_OK) return rc;
    const int nu = S.nu;
    std::vector<double> X(nu, 0.0);
    std::vector<int> img_of(nu / 6, 0);
    for (int i = 0, k = 0; i < n_images; i++)
        if (init[i].fixed == 0) {                          // pX from the float transforms (:6380-6393)
            X[6 * k + 0] = out[i].h.m[0]; X[6 * k + 1] = out[i].h.m[1]; X[6 * k + 2] = out[i].h.m[3]; X[6 * k + 3] = out[i].h.m[4];
            X[6 * k + 4] = out[i].h.m[2]; X[6 * k + 5] = out[i].h.m[5];
            img_of[k++] = i;
        }
    std::vector<double> Nt, rhs(nu);
    for (int it = 0; it < iterations; it++) {              // MAX_ITER = 10 (:6404), no convergence test
        Nt = S.N;
        for (int r = 0; r < nu; r++) {                     // A^T (A X - b)
            double acc = -S.g[r];
            const double* row = &S.N[(size_t)r * nu];
            for (int c = 0; c < nu; c++) acc += row[c] * X[c];
            rhs[r] = acc;
        }
        for (int k = 0; k < nu / 6; k++) {
            const int u = 6 * k;
            const double a = X[u], b = X[u + 1], c = X[u + 2], d = X[u + 3];
            const double w = (double)(int)((float)S.freq[img_of[k]] * weight);            // freqMatch[n] *= weight (int), wRot = nC (:6334-6338, :6637)
            const double J[3][4] = {{w * b, w * a, w * d, w * c}, {w * 2 * a, 0, w * 2 * c, 0}, {0, w * 2 * b, 0, w * 2 * d}};
            const double res[3] = {w * (a * b + c * d), w * (a * a + c * c - 1), w * (b * b + d * d - 1)};
            for (int q = 0; q < 3; q++)
                for (int i = 0; i < 4; i++) {
                    rhs[u + i] += J[q][i] * res[q];
                    for (int j = 0; j < 4; j++) Nt[(size_t)(u + i) * nu + u + j] += J[q][i] * J[q][j];
                }
        }
        if (solve_sym(Nt, rhs, nu) != 0) return UAVM_EFAIL;
        for (int r = 0; r < nu; r++) X[r] -= rhs[r];       // pX = pX - pDX (:6655-6656)
    }
    unpack6(X, init, n_images, out);
    return UAVM_OK;
}
