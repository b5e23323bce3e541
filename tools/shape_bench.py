#!/usr/bin/env python3
"""Device-resident extraction throughput on the other BASELINE shapes (configs[1], configs[3]) by frames per pass.
usage: tools/shape_bench.py            one line of JSON: {shape: {frames_per_pass: frames/s}}
The pass sizes bracket the L2-resident size (126 MB / pyramid bytes per frame) and the default (constant pixels per pass)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from viorb_b200 import api, synth  # noqa: E402

SHAPES = {"kitti 1241x376 nf2000": (376, 1241, 2000, 512, (24, 32, 48, 64, 99, 128)),
          "hd 1920x1080 nf5000": (1080, 1920, 5000, 128, (6, 8, 12, 16, 22, 32, 48)),
          "uhd 3840x2160 nf5000": (2160, 3840, 5000, 32, (1, 2, 3, 4, 5, 8))}


def main():
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = api.Context(0, stream.cuda_stream)
    out = {}
    for name, (h, w, nf, B, chunks) in SHAPES.items():
        uniq = min(B, 16)
        imgs = synth.frames(uniq, h, w, seed0=100)
        d = torch.from_numpy(np.ascontiguousarray(np.tile(imgs, (B // uniq, 1, 1)))).to(dev)
        out[name] = {}
        for chunk in chunks:
            ex = api.ORBextractor(nf, 1.2, 8, 20, 7, ctx=ctx)
            cap = ex.cap
            d_kps = torch.empty((B, cap, 7), dtype=torch.float32, device=dev)
            d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
            d_cnt = torch.zeros((B,), dtype=torch.int32, device=dev)
            ex.configure(chunk_frames=chunk)
            for _ in range(2):
                ex.extract_batch_device(d, B, h, w, d_kps, d_desc, d_cnt)
            ex.check()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            steps = 3
            torch.cuda.synchronize()
            e0.record()
            for _ in range(steps):
                ex.extract_batch_device(d, B, h, w, d_kps, d_desc, d_cnt)
            e1.record()
            torch.cuda.synchronize()
            ex.check()
            ms = e0.elapsed_time(e1) / steps
            out[name][chunk] = round(B / (ms * 1e-3), 1)
            ex.close()
            del d_kps, d_desc, d_cnt
            torch.cuda.empty_cache()
        del d
    print(json.dumps(out))


if __name__ == "__main__":
    main()
