"""Small end-to-end run of every kernel for compute-sanitizer (memcheck / racecheck): one odd-sized frame through
the extractor, a stereo pair, the windowed searches and both top-2 kernels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from viorb_b200 import api, synth  # noqa: E402
import scenarios as S  # noqa: E402

api.lib()
ctx = api.Context(0)
for (h, w, nf) in ((360, 640, 777), (97, 211, 200)):
    ex = api.ORBextractor(nf, 1.2, 3 if h < 100 else 8, 20, 7, ctx=ctx)
    k, d = ex(synth.frame(h, w, 5))
    print("extract", h, w, len(k))
    imgs = synth.frames(3, h, w, seed0=9)
    ex.configure(chunk_frames=2)
    ex.extract_batch(imgs)
    ex.set_describe_mode(2)             # blur_levels_kernel + describe_blurred_kernel on the same inputs
    k2, d2 = ex(synth.frame(h, w, 5))
    assert k2.tobytes() == k.tobytes() and (d2 == d).all()
    ex.extract_batch(imgs)
    ex.set_describe_mode(0)
    ex.configure(chunk_frames=12)
    ex.extract_batch(synth.frames(12, h, w, seed0=20))      # more than 8 frames per pass: whole levels by default
    ex.close()
left, right, disp = synth.stereo_pair(376, 620, 7)
exl, exr = api.ORBextractor(800, 1.2, 8, 20, 7, ctx=ctx), api.ORBextractor(800, 1.2, 8, 20, 7, ctx=ctx)
kl, dl = exl(left)
kr, dr = exr(right)
ur, dp = api.ComputeStereoMatches(exl, exr, kl, dl, kr, dr, S.KITTI_BF, S.KITTI_BF / S.KITTI_FX)
print("stereo", int((ur >= 0).sum()))
sf = exl.GetScaleFactors()
sc = S.projection_scenario(kl, dl, sf, seed=3, n_mp=300, conflicts=30)
fi = api.FrameIndex(ctx, kl, dl, sc["u_right"], (0.0, 620.0, 0.0, 376.0), sf)
m = api.ORBmatcher(0.8, True, ctx=ctx)
print("local", m.SearchByProjectionLocal(fi, sc["obs0"], sc["proj_x"], sc["proj_y"], sc["proj_xr"], sc["pred_level"],
                                         sc["view_cos"], sc["valid"], sc["nobs"], sc["mp_desc"], 3.0)[0])
for mode in (0, 1, 2, 8, 11):
    print("frame", mode, m.SearchByProjectionFrame(fi, sc["obs0"], sc["proj_x"], sc["proj_y"], sc["invz"], sc["last_octave"],
                                                   sc["last_angle"], sc["valid"], sc["nobs"], sc["mp_desc"], 10.0, S.KITTI_BF,
                                                   mode, 100)[0])
fi.GetFeaturesInArea(100.0, 100.0, 30.0, 0, 3)
fv1 = S.feature_vector(kl, S.row_band_nodes())
fv2 = S.feature_vector(kr, S.row_band_nodes(drop_every=5))
none1, none2 = np.full(len(kl), -1, np.float32), np.full(len(kr), -1, np.float32)
z1, z2 = np.zeros(len(kl), np.uint8), np.zeros(len(kr), np.uint8)
print("tri", m.SearchForTriangulation(kl, dl, none1, z1, kr, dr, none2, z2, fv1, fv2, S.RECTIFIED_F12, 600.0, 180.0, sf,
                                      (sf * sf).astype(np.float32))[0])
dmap = synth.descriptor_map(30011, seed=1)
for q in (1, 3, 8, 200):
    m.hamming_top2(synth.queries_from_map(dmap, q, seed=2), dmap)
print(m.DescriptorDistance(dl[:5], dl[5:10]))
print("sanitize smoke done")
