#!/usr/bin/env python3
"""Device-resident extraction throughput on the other BASELINE shapes (configs[1], configs[3]); not the bench line."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from viorb_b200 import api, synth  # noqa: E402

SHAPES = {"kitti 1241x376 nf2000": (376, 1241, 2000, 256), "hd 1920x1080 nf5000": (1080, 1920, 5000, 128),
          "uhd 3840x2160 nf5000": (2160, 3840, 5000, 32)}


def main():
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = api.Context(0, stream.cuda_stream)
    out = {}
    for name, (h, w, nf, B) in SHAPES.items():
        ex = api.ORBextractor(nf, 1.2, 8, 20, 7, ctx=ctx)
        uniq = min(B, 16)
        imgs = synth.frames(uniq, h, w, seed0=100)
        d = torch.from_numpy(np.ascontiguousarray(np.tile(imgs, (B // uniq, 1, 1)))).to(dev)
        cap = ex.cap
        d_kps = torch.empty((B, cap, 7), dtype=torch.float32, device=dev)
        d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
        d_cnt = torch.zeros((B,), dtype=torch.int32, device=dev)
        ex.configure(chunk_frames=max(1, min(128, (128 * 752 * 480) // (h * w))))
        for _ in range(2):
            ex.extract_batch_device(d, B, h, w, d_kps, d_desc, d_cnt)
        ex.check()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 3
        e0.record()
        for _ in range(steps):
            ex.extract_batch_device(d, B, h, w, d_kps, d_desc, d_cnt)
        e1.record()
        torch.cuda.synchronize()
        ex.check()
        ms = e0.elapsed_time(e1) / steps
        out[name] = {"frames_per_s": B / (ms * 1e-3), "ms_per_frame": ms / B, "keypoints_per_frame": float(d_cnt.float().mean().item()),
                     "mpix_per_s": B * h * w / (ms * 1e-3) / 1e6}
        ex.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
