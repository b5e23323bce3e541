#!/usr/bin/env python3
"""Host<->device copy rates of this box for the e2e bound: 1.48 GB in (pinned), 0.28 GB out, alone and together."""
import torch

dev = torch.device("cuda:0")
nin, nout = 4096 * 360960, 4096 * 1128 * 60
h_in = torch.empty(nin, dtype=torch.uint8).pin_memory()
d_in = torch.empty(nin, dtype=torch.uint8, device=dev)
d_out = torch.empty(nout, dtype=torch.uint8, device=dev)
h_out = torch.empty(nout, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(do_in, do_out, chunk_frames):
    ci, co = chunk_frames * 360960, chunk_frames * 1128 * 60
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_event(a); s2.wait_event(a)
    for k in range(4096 // chunk_frames):
        if do_in:
            with torch.cuda.stream(s1):
                d_in[k * ci:(k + 1) * ci].copy_(h_in[k * ci:(k + 1) * ci], non_blocking=True)
        if do_out:
            with torch.cuda.stream(s2):
                h_out[k * co:(k + 1) * co].copy_(d_out[k * co:(k + 1) * co], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b)


for cf in (128, 32):
    for name, i, o in (("H2D only", 1, 0), ("D2H only", 0, 1), ("both", 1, 1)):
        run(i, o, cf)
        ms = min(run(i, o, cf) for _ in range(3))
        print("chunk %3d %-9s %.2f ms  in %.1f GB/s out %.1f GB/s  -> %.0f frames/s bound" %
              (cf, name, ms, i * nin / ms / 1e6, o * nout / ms / 1e6, 4096 / ms * 1e3))
