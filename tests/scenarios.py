"""Synthetic matcher scenarios shared by the CPU and GPU tests (inputs only, no expected values)."""
import numpy as np

from viorb_b200 import synth

KITTI_FX, KITTI_BF = 718.856, 386.1448          # Examples/Stereo/KITTI00-02.yaml


def flip_bits(desc, rng, max_flips):
    out = desc.copy()
    for i in range(len(out)):
        for b in rng.integers(0, 256, size=int(rng.integers(0, max_flips + 1))):
            out[i, b >> 3] ^= np.uint8(1 << (b & 7))
    return out


def snap_projection(u, v, invz):
    """(u, v, 1/z) that a camera with identity pose and unit intrinsics reproduces exactly from a float world point:
    Z = float(1/invz), 1/z = float(1/Z), X = float(u / (1/z)), u = float(X * (1/z)).  The adapters that drive the reference's
    own SearchByProjection overloads (oracle/refbuild/ref_matcher_capi.cpp) can only be given such projections."""
    f32 = np.float32
    Z = (1.0 / invz.astype(np.float64)).astype(f32)
    iz = (1.0 / Z.astype(np.float64)).astype(f32)
    X, Y = (u.astype(f32) / iz).astype(f32), (v.astype(f32) / iz).astype(f32)
    return (X * iz).astype(f32), (Y * iz).astype(f32), iz


def projection_scenario(kps, desc, scale_factors, seed, n_mp=600, conflicts=60):
    """A current frame (kps/desc) and a set of map points projecting near its keypoints.
    Returns dict of arrays for SearchByProjection (local) and (frame) variants."""
    rng = np.random.default_rng(seed)
    n = len(kps)
    pick = rng.integers(0, n, n_mp)
    pick[:conflicts] = pick[conflicts:2 * conflicts]          # several map points compete for the same keypoint
    order = rng.permutation(n_mp)
    pick = pick[order]
    px = kps["x"][pick] + rng.normal(0, 2.0, n_mp).astype(np.float32)
    py = kps["y"][pick] + rng.normal(0, 2.0, n_mp).astype(np.float32)
    mp_desc = flip_bits(desc[pick], rng, 60)
    dup = rng.integers(0, n_mp, 40)
    mp_desc[dup] = desc[pick[dup]]                            # exact copies -> distance ties between competitors
    octave = kps["octave"][pick]
    pred = np.clip(octave + rng.integers(-1, 2, n_mp), 0, len(scale_factors) - 1).astype(np.int32)
    u_right = np.where(rng.random(n) < 0.5, kps["x"] - rng.uniform(2, 60, n), -1).astype(np.float32)
    proj_xr = (px - (kps["x"][pick] - u_right[pick]) + rng.normal(0, 3.0, n_mp)).astype(np.float32)
    obs0 = (rng.random(n) < 0.1).astype(np.int32) * 2
    valid = (rng.random(n_mp) < 0.9).astype(np.uint8)
    nobs = rng.integers(0, 4, n_mp).astype(np.int32)
    invz = rng.uniform(0.02, 0.5, n_mp).astype(np.float32)
    px, py, invz = snap_projection(px.astype(np.float32), py.astype(np.float32), invz)
    return dict(
        proj_x=px, proj_y=py, proj_xr=proj_xr,
        pred_level=pred, view_cos=rng.uniform(0.99, 1.0, n_mp).astype(np.float32),
        valid=valid, nobs=nobs, mp_desc=mp_desc, u_right=u_right, obs0=obs0, invz=invz, last_octave=octave.astype(np.int32),
        last_angle=(kps["angle"][pick] + rng.choice([0.0, 0.0, 0.0, 90.0, 200.0], n_mp) +
                    rng.normal(0, 3, n_mp)).astype(np.float32) % np.float32(360.0))


def feature_vector(kps, nodes_of):
    """DBoW2::FeatureVector stand-in: {node id: [keypoint indices]} flattened to (node_ids, node_ptr, idx)."""
    node = nodes_of(kps)
    ids = np.unique(node)
    ptr, idx = [0], []
    for i in ids:
        members = np.nonzero(node == i)[0]
        idx.extend(members.tolist())
        ptr.append(len(idx))
    return ids.astype(np.int32), np.array(ptr, np.int32), np.array(idx, np.int32)


def row_band_nodes(band=24, drop_every=7):
    def f(kps):
        node = (kps["y"] // band).astype(np.int32) * 3 + 5
        node[node % drop_every == 0] += 100000          # some nodes exist in one key frame only
        return node
    return f


RECTIFIED_F12 = np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0]], np.float32)   # l = x1' F12 = [0, 1, -y1]


def bow_inputs(k1, d1, k2, d2, seed):
    """two keypoint sets with feature vectors; some descriptors of set 1 are planted into set 2 (exact ties)"""
    rng = np.random.default_rng(seed)
    k2, d2 = k2.copy(), d2.copy()
    fv1 = feature_vector(k1, row_band_nodes())
    fv2 = feature_vector(k2, row_band_nodes(drop_every=5))
    # plant near-copies of set-1 descriptors into set 2 inside the same node band, several per source (competition)
    for _ in range(300):
        i = int(rng.integers(0, len(k1)))
        band = np.nonzero((k2["y"] // 24) == (k1["y"][i] // 24))[0]
        if len(band) == 0:
            continue
        j = int(band[rng.integers(0, len(band))])
        d2[j] = flip_bits(d1[i:i + 1], rng, 12)[0]
        k2["angle"][j] = (k1["angle"][i] + rng.choice([0.0, 0.0, 0.0, 0.0, 100.0, 250.0])) % 360.0
    v1 = (rng.random(len(k1)) < 0.8).astype(np.uint8)
    v2 = (rng.random(len(k2)) < 0.8).astype(np.uint8)
    return k1, d1, v1, k2, d2, v2, fv1, fv2


def window_queries(k_src, d_src, k_dst, seed, disp):
    """map points of `src` keypoints projected near the corresponding place in `dst` (stereo pair: shift by disparity)"""
    rng = np.random.default_rng(seed)
    n = len(k_src)
    u = (k_src["x"] - disp + rng.normal(0, 1.5, n)).astype(np.float32)
    v = (k_src["y"] + rng.normal(0, 1.0, n)).astype(np.float32)
    level = np.clip(k_src["octave"] + rng.integers(0, 2, n), 0, 7).astype(np.int32)
    valid = (rng.random(n) < 0.85).astype(np.uint8)
    return u, v, level, valid, flip_bits(d_src, rng, 20)


def distinctive_batch(seed, nmp=300, max_obs=40, big=(0, 1, 2, 33, 257)):
    """CSR batch of map-point observation descriptors for MapPoint::ComputeDistinctiveDescriptors: clusters of noisy
    copies of a base descriptor (so medians tie between rows), plus points with 0/1/2 and many observations."""
    rng = np.random.default_rng(seed)
    counts = list(big) + [int(c) for c in rng.integers(1, max_obs, nmp - len(big))]
    rows, ptr = [], [0]
    for c in counts:
        base = rng.integers(0, 256, 32).astype(np.uint8)
        d = np.repeat(base[None], c, 0)
        flips = rng.integers(0, 256, (c, 32)).astype(np.uint8) & rng.integers(0, 256, (c, 32)).astype(np.uint8) \
            & rng.integers(0, 256, (c, 32)).astype(np.uint8)
        d ^= flips
        if c > 3:
            d[c // 2] = d[0]                      # exact duplicates -> equal medians, first row must win
        rows.append(d)
        ptr.append(ptr[-1] + c)
    return np.concatenate(rows).astype(np.uint8), np.asarray(ptr, np.int32)


def vocabulary(seed, k=6, L=4, unbalanced=True):
    """Synthetic DBoW2 tree in loadFromTextFile form: (parent, desc, weight) per node, node 0 = root.  Children are
    noisy copies of their parent (so the descent is meaningful) with some exact duplicates among siblings (ties: the
    first child must win); with `unbalanced` some inner nodes are leaves above depth L; ~5% of the words are stopped
    (weight 0)."""
    rng = np.random.default_rng(seed)
    parent, desc, depth = [0], [np.zeros(32, np.uint8)], [0]
    frontier = [0]
    while frontier:
        nxt = []
        for p in frontier:
            if depth[p] >= L or (unbalanced and depth[p] >= 1 and rng.random() < 0.15):
                continue
            nch = k if not unbalanced else int(rng.integers(2, k + 1))
            base = desc[p] if p else rng.integers(0, 256, 32).astype(np.uint8)
            for c in range(nch):
                d = base ^ (rng.integers(0, 256, 32).astype(np.uint8) & rng.integers(0, 256, 32).astype(np.uint8)
                            & (rng.integers(0, 256, 32).astype(np.uint8) if depth[p] else 0xff))
                if c == nch - 1 and nch > 2 and rng.random() < 0.3:
                    d = desc[len(desc) - 1].copy()          # duplicate of the previous sibling
                parent.append(p); desc.append(d.astype(np.uint8)); depth.append(depth[p] + 1)
                nxt.append(len(parent) - 1)
        frontier = nxt
    n = len(parent)
    children = np.bincount(np.asarray(parent[1:]), minlength=n)
    weight = np.zeros(n)
    leaves = [i for i in range(1, n) if children[i] == 0]
    weight[leaves] = -np.log(rng.uniform(0.001, 0.9, len(leaves)))
    stopped = rng.random(len(leaves)) < 0.05
    weight[np.asarray(leaves)[stopped]] = 0.0
    return np.asarray(parent, np.int32), np.stack(desc).astype(np.uint8), weight


def vocabulary_features(seed, voc, n=700):
    """descriptors to transform: noisy copies of random leaves (many features per word) plus fresh random ones"""
    parent, desc, weight = voc
    rng = np.random.default_rng(seed)
    children = np.bincount(parent[1:], minlength=len(parent))
    leaves = np.nonzero(children[1:] == 0)[0] + 1
    pick = desc[rng.choice(leaves[:max(len(leaves) // 3, 1)], n)]
    noise = rng.integers(0, 256, (n, 32)).astype(np.uint8) & rng.integers(0, 256, (n, 32)).astype(np.uint8) \
        & rng.integers(0, 256, (n, 32)).astype(np.uint8) & rng.integers(0, 256, (n, 32)).astype(np.uint8)
    out = pick ^ noise
    out[::7] = rng.integers(0, 256, (len(out[::7]), 32)).astype(np.uint8)
    return out.astype(np.uint8)
