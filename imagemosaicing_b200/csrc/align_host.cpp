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
