"""Independent Python restatement of ORBextractor::operator() on OpenCV (cv2) primitives.

Used only to PIN the C++ oracle: every arithmetic primitive here is the real OpenCV one (cv2.resize,
cv2.copyMakeBorder, cv2.FastFeatureDetector per cell, cv2.GaussianBlur, cv2.fastAtan2) and libm's
sincosf, glued together exactly as /root/reference/src/ORBextractor.cc does (line numbers inline).
The only non-OpenCV logic is the quadtree (DistributeOctTree), restated here a second time,
independently of oracle/orb_oracle.cpp, with the same documented tie rule (node creation order).
"""
import ctypes
import math

import cv2
import numpy as np

f32 = np.float32
_libm = ctypes.CDLL("libm.so.6")
_libm.sincosf.argtypes = [ctypes.c_float, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]


def cv_round(v):
    return int(np.rint(v))


class Params:
    def __init__(self, nfeatures, scale_factor, nlevels, ini_th, min_th):
        # ORBextractor.cc:410-470
        self.nfeatures, self.nlevels, self.ini_th, self.min_th = nfeatures, nlevels, ini_th, min_th
        sfd = float(f32(scale_factor))          # double scaleFactor member initialised from a float
        self.scale = [f32(1.0)]
        for _ in range(1, nlevels):
            self.scale.append(f32(float(self.scale[-1]) * sfd))
        self.inv_scale = [f32(1.0) / s for s in self.scale]
        factor = f32(1.0 / sfd)
        nd = f32(f32(nfeatures) * (f32(1) - factor)) / (f32(1) - f32(math.pow(float(factor), float(nlevels))))
        self.quota, tot = [], 0
        for _ in range(nlevels - 1):
            self.quota.append(cv_round(nd))
            tot += self.quota[-1]
            nd = f32(nd * factor)
        self.quota.append(max(nfeatures - tot, 0))
        hp = 15
        vmax = int(math.floor(f32(f32(hp) * f32(math.sqrt(2.0))) / f32(2) + f32(1)))
        vmin = int(math.ceil(f32(f32(hp) * f32(math.sqrt(2.0))) / f32(2)))
        self.umax = [0] * (hp + 1)
        for v in range(vmax + 1):
            self.umax[v] = cv_round(math.sqrt(hp * hp - v * v))
        v0 = 0
        for v in range(hp, vmin - 1, -1):
            while self.umax[v0] == self.umax[v0 + 1]:
                v0 += 1
            self.umax[v] = v0
            v0 += 1


def compute_pyramid(p, image):
    # ORBextractor.cc:1107-1132 ; returns padded levels (border 19)
    rows, cols = image.shape
    levels = []
    for l in range(p.nlevels):
        w = cv_round(f32(cols) * p.inv_scale[l])
        h = cv_round(f32(rows) * p.inv_scale[l])
        if l == 0:
            levels.append(cv2.copyMakeBorder(image, 19, 19, 19, 19, cv2.BORDER_REFLECT_101))
        else:
            prev = levels[-1][19:-19, 19:-19]
            r = cv2.resize(prev, (w, h), interpolation=cv2.INTER_LINEAR)
            levels.append(cv2.copyMakeBorder(r, 19, 19, 19, 19, cv2.BORDER_REFLECT_101))
    return levels


class _Node:
    __slots__ = ("keys", "ul", "ur", "bl", "br", "nomore", "seq", "alive")


def _divide(n, new_seq):
    # ExtractorNode::DivideNode ORBextractor.cc:481-537
    halfx = int(math.ceil(f32(n.ur[0] - n.ul[0]) / f32(2)))
    halfy = int(math.ceil(f32(n.br[1] - n.ul[1]) / f32(2)))
    c = [_Node() for _ in range(4)]
    c[0].ul = n.ul; c[0].ur = (n.ul[0] + halfx, n.ul[1]); c[0].bl = (n.ul[0], n.ul[1] + halfy)
    c[0].br = (n.ul[0] + halfx, n.ul[1] + halfy)
    c[1].ul = c[0].ur; c[1].ur = n.ur; c[1].bl = c[0].br; c[1].br = (n.ur[0], n.ul[1] + halfy)
    c[2].ul = c[0].bl; c[2].ur = c[0].br; c[2].bl = n.bl; c[2].br = (c[0].br[0], n.bl[1])
    c[3].ul = c[2].ur; c[3].ur = c[1].br; c[3].bl = c[2].br; c[3].br = n.br
    for ch in c:
        ch.keys = []
        ch.nomore = False
        ch.alive = True
    for k in n.keys:
        if k[0] < c[0].ur[0]:
            (c[0] if k[1] < c[0].br[1] else c[2]).keys.append(k)
        elif k[1] < c[0].br[1]:
            c[1].keys.append(k)
        else:
            c[3].keys.append(k)
    for ch in c:
        if len(ch.keys) == 1:
            ch.nomore = True
    return c


def distribute_octree(keys, min_x, max_x, min_y, max_y, N):
    """keys: list of (x, y, response, tag).  Returns selected keys in reference list order.
    The std::list is modelled by a python list kept in *reverse* (front == end of the list)."""
    n_ini = int(np.round(f32(max_x - min_x) / f32(max_y - min_y)))   # round() half away; ratio > 0
    n_ini = int(math.floor(float(f32(max_x - min_x) / f32(max_y - min_y)) + 0.5))
    hx = f32(max_x - min_x) / f32(n_ini)
    seq = [0]

    def nseq():
        seq[0] += 1
        return seq[0] - 1

    roots = []
    for i in range(n_ini):
        n = _Node()
        n.ul = (int(hx * f32(i)), 0); n.ur = (int(hx * f32(i + 1)), 0)
        n.bl = (n.ul[0], max_y - min_y); n.br = (n.ur[0], max_y - min_y)
        n.keys = []; n.nomore = False; n.alive = True; n.seq = nseq()
        roots.append(n)
    for k in keys:
        roots[int(f32(k[0]) / hx)].keys.append(k)
    # list front..back order
    lst = []
    for n in roots:
        if len(n.keys) == 1:
            n.nomore = True
        if n.keys:
            lst.append(n)
    finish = False
    size_ptr = []

    def push_children(ch, front, count):
        for c in ch:
            if c.keys:
                c.seq = nseq()
                front.append(c)            # push_front (front list is reversed at the end)
                if len(c.keys) > 1:
                    count[0] += 1
                    size_ptr.append(c)

    while not finish:
        prev = len(lst)
        n_expand = [0]
        size_ptr = []
        front = []
        keep = []
        for n in lst:
            if n.nomore:
                keep.append(n)
                continue
            push_children(_divide(n, None), front, n_expand)
        lst = front[::-1] + keep
        if len(lst) >= N or len(lst) == prev:
            finish = True
        elif len(lst) + n_expand[0] * 3 > N:
            while not finish:
                prev = len(lst)
                vprev = sorted(size_ptr, key=lambda n: (len(n.keys), n.seq))
                size_ptr = []
                front = []
                cur = len(lst)
                dummy = [0]
                for n in reversed(vprev):
                    before = len(front)
                    push_children(_divide(n, None), front, dummy)
                    n.alive = False
                    cur += len(front) - before - 1
                    if cur >= N:
                        break
                lst = front[::-1] + [n for n in lst if n.alive]
                if len(lst) >= N or len(lst) == prev:
                    finish = True
    out = []
    for n in lst:
        best = n.keys[0]
        for k in n.keys[1:]:
            if k[2] > best[2]:
                best = k
        out.append(best)
    return out


def detect_level(p, padded, level):
    # ORBextractor::ComputeKeyPointsOctTree ORBextractor.cc:765-853 for one level
    roi = padded[19:-19, 19:-19]
    h, w = roi.shape
    min_bx = min_by = 16
    max_bx, max_by = w - 16, h - 16
    width, height = f32(max_bx - min_bx), f32(max_by - min_by)
    ncols, nrows = int(width / f32(30)), int(height / f32(30))
    wcell = int(math.ceil(width / f32(ncols)))
    hcell = int(math.ceil(height / f32(nrows)))
    det = {t: cv2.FastFeatureDetector_create(t, True) for t in (p.ini_th, p.min_th)}
    cand = []
    for i in range(nrows):
        ini_y = min_by + i * hcell
        max_y = ini_y + hcell + 6
        if ini_y >= max_by - 3:
            continue
        max_y = min(max_y, max_by)
        for j in range(ncols):
            ini_x = min_bx + j * wcell
            max_x = ini_x + wcell + 6
            if ini_x >= max_bx - 6:
                continue
            max_x = min(max_x, max_bx)
            cell = np.ascontiguousarray(roi[ini_y:max_y, ini_x:max_x])
            k = det[p.ini_th].detect(cell)
            if not k:
                k = det[p.min_th].detect(cell)
            for kp in k:
                cand.append((int(kp.pt[0]) + j * wcell, int(kp.pt[1]) + i * hcell, int(kp.response), len(cand)))
    sel = distribute_octree(cand, min_bx, max_bx, min_by, max_by, p.quota[level]) if cand else []
    return cand, [(k[0] + min_bx, k[1] + min_by, k[2]) for k in sel]


def ic_angle(p, padded, x, y):
    # IC_Angle ORBextractor.cc:77-104
    roi = padded.astype(np.int64)
    cy, cx = y + 19, x + 19
    m10 = m01 = 0
    for u in range(-15, 16):
        m10 += u * int(roi[cy, cx + u])
    for v in range(1, 16):
        d = p.umax[v]
        vs = 0
        for u in range(-d, d + 1):
            a, b = int(roi[cy + v, cx + u]), int(roi[cy - v, cx + u])
            vs += a - b
            m10 += u * (a + b)
        m01 += v * vs
    return f32(cv2.fastAtan2(float(m01), float(m10)))


def descriptor(pattern, blurred_padded, x, y, angle_deg):
    # computeOrbDescriptor ORBextractor.cc:108-147
    factor_pi = f32(math.pi / float(f32(180.0)))
    ang = f32(f32(angle_deg) * factor_pi)
    s, c = ctypes.c_float(), ctypes.c_float()
    _libm.sincosf(float(ang), ctypes.byref(s), ctypes.byref(c))
    a, b = f32(c.value), f32(s.value)
    px = pattern[:, 0].astype(np.float32)
    py = pattern[:, 1].astype(np.float32)
    rr = np.rint(px * b + py * a).astype(np.int64)       # float32 mul, float32 add (no FMA in numpy)
    cc = np.rint(px * a - py * b).astype(np.int64)
    vals = blurred_padded[y + 19 + rr, x + 19 + cc].astype(np.int32)
    bits = (vals[0::2] < vals[1::2]).astype(np.uint8)
    return np.packbits(bits, bitorder="little")


def extract(p, image, pattern):
    """Returns (levels, per-level candidates, keypoints list of dicts, descriptors uint8 [N,32])."""
    levels = compute_pyramid(p, image)
    all_cand, kps, descs = [], [], []
    per_level = []
    for l in range(p.nlevels):
        cand, sel = detect_level(p, levels[l], l)
        all_cand.append(cand)
        per_level.append(sel)
    for l in range(p.nlevels):
        if not per_level[l]:
            continue
        roi = np.ascontiguousarray(levels[l][19:-19, 19:-19])
        bl = cv2.GaussianBlur(roi, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        blp = cv2.copyMakeBorder(bl, 19, 19, 19, 19, cv2.BORDER_CONSTANT, value=0)
        for (x, y, resp) in per_level[l]:
            ang = ic_angle(p, levels[l], x, y)
            descs.append(descriptor(pattern, blp, x, y, ang))
            s = p.scale[l] if l else f32(1)
            kps.append(dict(x=f32(x) * s if l else f32(x), y=f32(y) * s if l else f32(y),
                            size=f32(int(f32(31) * p.scale[l])), angle=ang, response=f32(resp), octave=l))
    return levels, all_cand, kps, (np.stack(descs) if descs else np.zeros((0, 32), np.uint8))
