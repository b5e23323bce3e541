#!/usr/bin/env python3
"""Single-call latency of the drop-in entry points (BASELINE configs 0 and 1): one 752x480 frame through viorb_extract,
and one KITTI-shape stereo pair (2 x extract + ComputeStereoMatches), host buffers in and out, wall clock."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from viorb_b200 import api, synth  # noqa: E402


def timeit(fn, n=200, warm=20):
    for _ in range(warm):
        fn()
    t = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        t.append((time.perf_counter() - t0) * 1e3)
    t = np.sort(np.asarray(t))
    return {"median_ms": float(np.median(t)), "p10_ms": float(t[len(t) // 10]), "p90_ms": float(t[9 * len(t) // 10])}


def main():
    ctx = api.Context(0)
    out = {}
    img = api.pinned_empty((480, 752), np.uint8)
    img[:] = synth.frame(480, 752, 0)
    ex = api.ORBextractor(1000, 1.2, 8, 20, 7, ctx=ctx)
    out["euroc_frame_extract"] = timeit(lambda: ex(img))
    ex.profile(True)
    for _ in range(50):
        ex(img)
    st, passes = ex.stage_ms()
    ex.profile(False)
    out["euroc_frame_stage_us"] = {k: 1e3 * v / max(passes, 1) for k, v in st.items()}
    left, right, _ = synth.stereo_pair(376, 1241, 7)
    exl, exr = api.ORBextractor(2000, 1.2, 8, 20, 7, ctx=ctx), api.ORBextractor(2000, 1.2, 8, 20, 7, ctx=ctx)

    def stereo():
        kl, dl = exl(left)
        kr, dr = exr(right)
        return api.ComputeStereoMatches(exl, exr, kl, dl, kr, dr, 386.1448, 386.1448 / 718.856)

    out["kitti_stereo_pair"] = timeit(stereo, n=100)
    # the reference extracts left and right in two threads (src/Frame.cc:258-261): one context per thread
    import threading
    ctx2 = api.Context(0)
    exr2 = api.ORBextractor(2000, 1.2, 8, 20, 7, ctx=ctx2)
    res = [None, None]

    def stereo_threads():
        def right_job():
            res[1] = exr2(right)
        t = threading.Thread(target=right_job)
        t.start()
        res[0] = exl(left)
        t.join()
        return api.ComputeStereoMatches(exl, exr2, res[0][0], res[0][1], res[1][0], res[1][1], 386.1448, 386.1448 / 718.856)

    out["kitti_stereo_pair_two_threads"] = timeit(stereo_threads, n=100)
    kl, dl = exl(left)
    kr, dr = exr(right)
    out["kitti_compute_stereo_matches"] = timeit(lambda: api.ComputeStereoMatches(exl, exr, kl, dl, kr, dr, 386.1448, 386.1448 / 718.856), n=100)
    big = synth.frame(1080, 1920, 1)
    exb = api.ORBextractor(5000, 1.2, 8, 20, 7, ctx=ctx)
    out["hd1080_frame_extract"] = timeit(lambda: exb(big), n=50, warm=5)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
