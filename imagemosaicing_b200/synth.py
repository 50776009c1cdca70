"""Seeded synthetic inputs for the hot path: SIFT-like u8 descriptors, keypoints,
pair geometry with known homographies, band-limited BGR textures.  Used by tests/ and bench.py."""
import numpy as np

SEED_BASE = 20160308


def sift_like_descriptors(rng, n):
    """gamma(1,30) -> L2 normalise -> clip 0.2 -> renormalise -> x512 -> floor -> saturate 255 (u8)."""
    d = rng.gamma(1.0, 30.0, size=(n, 128)).astype(np.float32)
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-12)
    d = np.minimum(d, 0.2)
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-12)
    return np.minimum(np.floor(d * 512.0), 255).astype(np.uint8)


def random_keypoints(rng, n, w, h):
    return np.stack([rng.uniform(0, w - 1, n), rng.uniform(0, h - 1, n)], 1).astype(np.float32)


def pair_homography(rng, w, h, overlap=(0.6, 0.8), max_rot_deg=5.0, scale=(0.95, 1.05), proj=1e-5):
    """H maps image-2 points to image-1 points (the reference's convention, M/matrix.h:783)."""
    th = np.deg2rad(rng.uniform(-max_rot_deg, max_rot_deg))
    s = rng.uniform(*scale)
    ov = rng.uniform(*overlap)
    tx = rng.uniform(-0.05, 0.05) * w
    ty = (1.0 - ov) * h * (1 if rng.random() < 0.5 else -1)
    c, si = np.cos(th) * s, np.sin(th) * s
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    H = np.array([[c, -si, cx - c * cx + si * cy + tx],
                  [si, c, cy - si * cx - c * cy + ty],
                  [rng.uniform(-proj, proj), rng.uniform(-proj, proj), 1.0]])
    return H


def apply_h(H, xy):
    q = np.c_[xy, np.ones(len(xy))] @ H.T
    return q[:, :2] / q[:, 2:3]


def strip_poses(rng, n_images, w, h, overlap=(0.6, 0.8), max_rot_deg=5.0, scale=(0.95, 1.05)):
    """Absolute poses T_k (image k -> mosaic frame): rotation +-5 deg and scale 0.95..1.05 about the image
    centre, along-track step (1 - overlap) * h, small cross-track jitter.  T_0 = identity."""
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    T = [np.eye(3)]
    ty = 0.0
    for k in range(1, n_images):
        th = np.deg2rad(rng.uniform(-max_rot_deg, max_rot_deg)); s = rng.uniform(*scale)
        ty += (1.0 - rng.uniform(*overlap)) * h
        tx = rng.uniform(-0.05, 0.05) * w
        c, si = np.cos(th) * s, np.sin(th) * s
        T.append(np.array([[c, -si, cx - c * cx + si * cy + tx], [si, c, cy - si * cx - c * cy + ty], [0, 0, 1.0]]))
    return T


def make_strip(n_images, w, h, n_kp, seed=SEED_BASE, true_frac=0.5, desc_noise=4.0, pos_noise=0.5,
               geom_outlier_frac=0.5, proj=1e-5):
    """Sequential-overlap strip: image k+1 overlaps image k.  Returns (descs[u8], kps[f32], Hs) where
    Hs[k] maps image k+1 -> image k (the reference's convention, M/matrix.h:783) and is
    inv(T_k) T_{k+1} of the absolute poses plus projective terms <= 1e-5.  A fraction
    `geom_outlier_frac` of the descriptor-true matches is given a random position (repeated-structure
    mismatches), so the <=396 candidates that reach RANSAC hold ~50 % inliers like the reference's real
    runs (32..275 inliers of <=396 in matchPairs.match) and the loop really runs its 1000 counted
    hypotheses instead of taking the 0.99 early exit."""
    descs, kps, Hs = [], [], []
    rng = np.random.default_rng(seed)
    poses = strip_poses(np.random.default_rng(seed + 999983), n_images, w, h)
    descs.append(sift_like_descriptors(rng, n_kp)); kps.append(random_keypoints(rng, n_kp, w, h))
    for k in range(1, n_images):
        rng = np.random.default_rng(seed + k)
        H = np.linalg.inv(poses[k - 1]) @ poses[k]
        H[2, 0] += rng.uniform(-proj, proj); H[2, 1] += rng.uniform(-proj, proj)
        Hi = np.linalg.inv(H)
        p_prev = kps[k - 1].astype(np.float64)
        p2 = apply_h(Hi, p_prev) + rng.normal(0, pos_noise, size=p_prev.shape)
        inside = (p2[:, 0] >= 0) & (p2[:, 0] <= w - 1) & (p2[:, 1] >= 0) & (p2[:, 1] <= h - 1)
        take = inside & (rng.random(len(p2)) < true_frac)
        idx = np.nonzero(take)[0]
        bad = rng.random(len(idx)) < geom_outlier_frac
        p2[idx[bad]] = random_keypoints(rng, int(bad.sum()), w, h)
        d_true = np.clip(np.rint(descs[k - 1][idx].astype(np.float32) + rng.normal(0, desc_noise, size=(len(idx), 128))), 0, 255).astype(np.uint8)
        n_rand = n_kp - len(idx)
        d = np.concatenate([d_true, sift_like_descriptors(rng, n_rand)], 0)
        p = np.concatenate([p2[idx].astype(np.float32), random_keypoints(rng, n_rand, w, h)], 0)
        perm = rng.permutation(n_kp)
        descs.append(np.ascontiguousarray(d[perm])); kps.append(np.ascontiguousarray(p[perm])); Hs.append(H)
    return descs, kps, Hs


def make_candidates(rng, n, w, h, inlier_frac=0.5, noise=0.5):
    """Candidate point pairs for RANSAC: (xy1, xy2, H) with xy1 ~ H xy2 for inliers."""
    xy2 = random_keypoints(rng, n, w, h)
    H = pair_homography(rng, w, h)
    xy1 = apply_h(H, xy2.astype(np.float64)) + rng.normal(0, noise, size=(n, 2))
    out = rng.random(n) > inlier_frac
    xy1[out] = random_keypoints(rng, int(out.sum()), w, h)
    return xy1.astype(np.float32), xy2, H


def texture_image(rng, w, h, n_waves=24):
    """Band-limited BGR u8 texture: a sum of random sinusoids per channel (deterministic per rng)."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((h, w, 3), np.float32)
    for c in range(3):
        acc = np.zeros((h, w), np.float32)
        for _ in range(n_waves):
            fx, fy = rng.uniform(-0.15, 0.15, 2)
            ph = rng.uniform(0, 2 * np.pi)
            acc += np.sin(xx * fx + yy * fy + ph).astype(np.float32)
        acc = (acc - acc.min()) / max(float(acc.max() - acc.min()), 1e-6)
        img[..., c] = acc * 255.0
    return img.astype(np.uint8)


def block_poses(rng, rows, cols, w, h, along=0.30, cross=0.70, max_rot_deg=5.0, scale=(0.95, 1.05)):
    """Absolute poses T_k (image k -> mosaic frame) of a UAV block of `rows` strips x `cols` frames (BASELINE configs[2]):
    along-track step `along`*w (70 % overlap), cross-track step `cross`*h (30 % overlap), rotation / scale jitter about the
    image centre, a few px of position jitter.  T_0 = identity."""
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    T = []
    for r in range(rows):
        for c in range(cols):
            if r == 0 and c == 0:
                T.append(np.eye(3)); continue
            th = np.deg2rad(rng.uniform(-max_rot_deg, max_rot_deg)); s = rng.uniform(*scale)
            tx = c * along * w + rng.uniform(-0.02, 0.02) * w; ty = r * cross * h + rng.uniform(-0.02, 0.02) * h
            co, si = np.cos(th) * s, np.sin(th) * s
            T.append(np.array([[co, -si, cx - co * cx + si * cy + tx], [si, co, cy - si * cx - co * cy + ty], [0, 0, 1.0]]))
    return T


def make_block(rows, cols, w, h, n_kp, seed=SEED_BASE, true_frac=0.5, desc_noise=4.0, pos_noise=0.5):
    """2-D block with a shared WORLD point model: world points (each with one SIFT-like descriptor) are scattered over the
    block so that an image sees ~true_frac * n_kp of them; every image observes the world points inside its footprint
    (position noise, descriptor noise) and fills the rest of its n_kp keypoints with clutter.  Any two overlapping images
    therefore share true correspondences, along and across strips, and the pair graph has loops.
    Returns (descs, kps, poses, pairs): poses = absolute ground-truth T_k, pairs = every (i, j), i < j, whose footprints'
    bounding boxes intersect (the 'all pairs in overlap' list of configs[2])."""
    rng = np.random.default_rng(seed)
    poses = block_poses(np.random.default_rng(seed + 424243), rows, cols, w, h)
    n = rows * cols
    corners = np.array([[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]], np.float64)
    boxes = []
    for T in poses:
        q = apply_h(T, corners); boxes.append((q[:, 0].min(), q[:, 1].min(), q[:, 0].max(), q[:, 1].max()))
    x0 = min(b[0] for b in boxes); y0 = min(b[1] for b in boxes); x1 = max(b[2] for b in boxes); y1 = max(b[3] for b in boxes)
    n_world = int(true_frac * n_kp * (x1 - x0) * (y1 - y0) / (w * h))
    wp = np.stack([rng.uniform(x0, x1, n_world), rng.uniform(y0, y1, n_world)], 1)
    wd = sift_like_descriptors(rng, n_world)
    descs, kps = [], []
    for k, T in enumerate(poses):
        r = np.random.default_rng(seed + 7919 * (k + 1))
        b = boxes[k]
        cand = np.nonzero((wp[:, 0] >= b[0]) & (wp[:, 0] <= b[2]) & (wp[:, 1] >= b[1]) & (wp[:, 1] <= b[3]))[0]
        p = apply_h(np.linalg.inv(T), wp[cand]) + r.normal(0, pos_noise, size=(len(cand), 2))
        ok = (p[:, 0] >= 0) & (p[:, 0] <= w - 1) & (p[:, 1] >= 0) & (p[:, 1] <= h - 1)
        cand = cand[ok][:n_kp]; p = p[ok][:n_kp]
        d_true = np.clip(np.rint(wd[cand].astype(np.float32) + r.normal(0, desc_noise, size=(len(cand), 128))), 0, 255).astype(np.uint8)
        n_rand = n_kp - len(cand)
        d = np.concatenate([d_true, sift_like_descriptors(r, n_rand)], 0)
        q = np.concatenate([p.astype(np.float32), random_keypoints(r, n_rand, w, h)], 0)
        perm = r.permutation(n_kp)
        descs.append(np.ascontiguousarray(d[perm])); kps.append(np.ascontiguousarray(q[perm]))
    pairs = [(i, j) for i in range(n) for j in range(i + 1, n)
             if boxes[i][0] < boxes[j][2] and boxes[j][0] < boxes[i][2] and boxes[i][1] < boxes[j][3] and boxes[j][1] < boxes[i][3]]
    return descs, kps, poses, np.array(pairs, np.int32).reshape(-1, 2)
