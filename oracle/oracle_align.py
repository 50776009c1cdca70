"""CPU restatement (numpy, dense) of the reference's constrained global-alignment variants.  TEST INFRASTRUCTURE ONLY.

  align_affine_constrained  <- BundleAdjustmentSparseConstraint, M/MosaicWithoutPos.cpp:6032-6300
  align_affine_rot          <- SparseAffineRotConstraint,        M/MosaicWithoutPos.cpp:6302-6808

Both are dead code in the shipped reference (methodType = 0, :4588) and no fixture exercises them: parity is UNPINNED beyond this
restatement, which builds the reference's rows one by one (same unknown order a b c d e f, same constraint rows and weights) and
solves the normal equations densely in double (the reference: CHOLMOD on A^T A, M/test_cholmod.cpp:180-262).
pairs: (P, 8) [imgA xA yA fixedA imgB xB yB fixedB]; fixed_img: (N,) 0/1; T0: (N, 9) initial transforms (used for fixed images)."""
import numpy as np


def _rows(pairs, fixed_img, T0):
    n = len(fixed_img)
    acc = np.concatenate([[0], np.cumsum(fixed_img)[:-1]]).astype(int)
    nu = 6 * int(n - fixed_img.sum())
    A = []; b = []
    freq = np.zeros(n, int)
    for p in pairs:
        ia, xa, ya, fa, ib, xb, yb, fb = int(p[0]), p[1], p[2], int(p[3]), int(p[4]), p[5], p[6], int(p[7])
        if fa == 0: freq[ia] += 1
        if fb == 0: freq[ib] += 1
        r1 = np.zeros(nu); r2 = np.zeros(nu); t1 = t2 = 0.0
        def put(img, x, y, s):
            c = 6 * (img - acc[img])
            r1[c + 0] += s * x; r1[c + 1] += s * y; r1[c + 4] += s
            r2[c + 2] += s * x; r2[c + 3] += s * y; r2[c + 5] += s
        def proj(img, x, y):
            h = T0[img].astype(np.float64)
            d = h[6] * x + h[7] * y + h[8]
            return (h[0] * x + h[1] * y + h[2]) / d, (h[3] * x + h[4] * y + h[5]) / d
        if fa == 0 and fb == 0:
            put(ia, xa, ya, 1.0); put(ib, xb, yb, -1.0)
        elif fa == 1 and fb == 0:
            put(ib, xb, yb, 1.0); t1, t2 = proj(ia, xa, ya)
        elif fa == 0 and fb == 1:
            put(ia, xa, ya, 1.0); t1, t2 = proj(ib, xb, yb)
        else:
            continue
        A.append(r1); A.append(r2); b.append(t1); b.append(t2)
    return np.array(A), np.array(b), acc, freq, nu


def _unpack(x, fixed_img, T0):
    n = len(fixed_img); out = np.zeros((n, 9), np.float32); k = 0
    for i in range(n):
        if fixed_img[i]:
            out[i] = T0[i]
        else:
            a, b, c, d, e, f = x[6 * k:6 * k + 6]; k += 1
            out[i] = [a, b, e, c, d, f, 0, 0, 1]
    return out


def align_affine_constrained(pairs, fixed_img, T0):
    pairs = np.asarray(pairs, np.float64); fixed_img = np.asarray(fixed_img, int)
    A, b, acc, freq, nu = _rows(pairs, fixed_img, T0)
    extra = []
    for i in range(len(fixed_img)):
        if fixed_img[i]:
            continue
        c = 6 * (i - acc[i]); nC = float(freq[i])
        r = np.zeros(nu); r[c + 0] = nC; r[c + 3] = -nC; extra.append(r)          # nC identical rows "a - d = 0" sum to nC (a - d)
        r = np.zeros(nu); r[c + 1] = nC; r[c + 2] = nC; extra.append(r)
    A2 = np.vstack([A, np.array(extra)]); b2 = np.concatenate([b, np.zeros(len(extra))])
    x = np.linalg.solve(A2.T @ A2, A2.T @ b2)
    return _unpack(x, fixed_img, T0)


def align_affine_rot(pairs, fixed_img, T0, weight=1.0, iterations=10):
    pairs = np.asarray(pairs, np.float64); fixed_img = np.asarray(fixed_img, int)
    T = align_affine_constrained(pairs, fixed_img, T0)                               # initial value, as float32 (:6374-6393)
    A, b, acc, freq, nu = _rows(pairs, fixed_img, T0)
    free = [i for i in range(len(fixed_img)) if not fixed_img[i]]
    X = np.zeros(nu)
    for k, i in enumerate(free):
        X[6 * k:6 * k + 6] = [T[i][0], T[i][1], T[i][3], T[i][4], T[i][2], T[i][5]]
    for _ in range(iterations):
        J = [A]; r = [A @ X - b]
        for k, i in enumerate(free):
            a, bb, c, d = X[6 * k:6 * k + 4]
            w = float(int(np.float32(freq[i]) * np.float32(weight)))
            rows = np.zeros((3, nu))
            rows[0, 6 * k:6 * k + 4] = [w * bb, w * a, w * d, w * c]
            rows[1, 6 * k + 0] = w * 2 * a; rows[1, 6 * k + 2] = w * 2 * c
            rows[2, 6 * k + 1] = w * 2 * bb; rows[2, 6 * k + 3] = w * 2 * d
            J.append(rows); r.append(np.array([w * (a * bb + c * d), w * (a * a + c * c - 1), w * (bb * bb + d * d - 1)]))
        J = np.vstack(J); r = np.concatenate(r)
        X = X - np.linalg.solve(J.T @ J, J.T @ r)
    return _unpack(X, fixed_img, T0)
