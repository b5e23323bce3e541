#!/usr/bin/env python3
"""bench.py -- ORB extraction throughput (BASELINE.json metric) on 1..8 B200s, plus the Hamming matcher.

  python bench.py --gpus N --steps K --warmup W          (N > 1: launched by torch.distributed.run)
  python bench.py --impl reference ...                    the reference's own ORBextractor (oracle/_ref: src/ORBextractor.cc
                                                          compiled unmodified) on the box's host cores, same config and metric

A "step" is one pass of the hot path -- ORBextractor::operator() -- over one batch of synthetic frames:
BASELINE.json configs[2], 4096 EuRoC-shape (752x480, 1000 features, 8 levels, 1.2, FAST 20/7) frames,
sharded contiguously by frame across the ranks (no data-path collective).  `value` is frames/s with the
frames resident in HBM; `e2e` is the same batch through the host-buffer C-ABI call (viorb_extract_batch:
pinned host frames in, keypoints+descriptors out, copies inside the timed region) with the box's measured
copy bound beside it.  The JSON line also carries

  parity      the timed workload is a parity workload: sha256 of every frame's result (device path and host path, every
              rank) against digests written by the reference itself (tests/golden/ref_extract_batch4096.json), and the
              N-rank merged top-2 of configs[4] against the CPU oracle on the full 1000 x 10 M
  roofline    dominant kernel, live CUDA-event timing
  cpu_baseline  the reference (all host cores, one frame per thread) + the oracle port + a cv2-primitives column
  shapes      configs[1] (KITTI) and configs[3] (1080p, 4K): throughput, step roofline, CPU reference
  matchers    per-call wall time of every ORBmatcher / Frame call surface through the C++ drop-in classes, beside the
              reference's own classes on the same scenario
  matcher     configs[4]: brute-force Hamming top-2, map sharded across ranks, NCCL all-gather of the per-shard top-2
              records + merge kernel
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

EUROC = dict(rows=480, cols=752, nfeatures=1000, scale=1.2, levels=8, ini=20, min=7)
# compulsory bytes per frame = W*H + sum_l (w_l+38)(h_l+38) + 60*N
BYTES_PER_FRAME_EUROC = 1765453
SHAPES = {      # name: rows, cols, nfeatures, frames in the device-resident batch, compulsory bytes per frame
    "configs[1] KITTI 1241x376 nf2000": (376, 1241, 2000, 512, 2325175),
    "configs[3] 1920x1080 nf5000": (1080, 1920, 5000, 128, 9329405),
    "configs[3] 3840x2160 nf5000": (2160, 3840, 5000, 32, 35333518),
}
L2_BYTES = 126 * 1024 * 1024


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=4096, help="frames in the whole job (all ranks)")
    ap.add_argument("--chunk", type=int, default=0, help="frames per device pass (0 = library default)")
    ap.add_argument("--map", type=int, default=10_000_000, help="map descriptors in the whole job (matcher)")
    ap.add_argument("--queries", type=int, default=1000)
    ap.add_argument("--no-matcher", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-latency", action="store_true", help="skip the per-call latency / call-surface blocks (profiling runs)")
    ap.add_argument("--no-shapes", action="store_true", help="skip the configs[1]/configs[3] block")
    ap.add_argument("--no-bind", action="store_true", help="do not bind the rank to the NUMA node of its GPU")
    ap.add_argument("--duplex", action="store_true", help="keep the default (two-stream) copy order of viorb_extract_batch")
    ap.add_argument("--cpu-frames", type=int, default=2048, help="frames of the CPU-baseline sample (about 30 CPU-seconds)")
    return ap.parse_args()


def workload_config(args):
    """identical in both arms"""
    return {"workload": "configs[2]: %d synthetic 752x480 frames, ORBextractor 1000/1.2/8/20/7, sharded by frame"
                        % args.frames,
            "frames": args.frames, "shape": [EUROC["rows"], EUROC["cols"]], "nfeatures": EUROC["nfeatures"],
            "l2": "inputs (%.2f GB per job) larger than the 126 MB L2; no flush needed"
                  % (args.frames * EUROC["rows"] * EUROC["cols"] / 1e9),
            "parallelism": "frame-sharded x%d" % args.gpus}


def frame_digest(kps, desc):
    """sha256 over the keypoint records (28 B each) followed by the descriptor rows (tests/util.py extraction_digest)"""
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(kps).tobytes())
    h.update(np.ascontiguousarray(desc, np.uint8).tobytes())
    return h.hexdigest()


# ------------------------------------------------------------------------------------------------ CPU arms
def load_oracle_native():
    """oracle port rebuilt with -march=native (the reference's own flags, CMakeLists.txt:10-11) for timing"""
    from oracle import oracle_py
    from viorb_b200 import build
    try:
        path = build.build_oracle(march="native", outdir="_build_native")
        return oracle_py, oracle_py.lib(path)
    except Exception:
        return oracle_py, oracle_py.lib()


def reference_available():
    try:
        from oracle import ref_py
        return os.path.exists(ref_py.PATH) or ref_py.available()
    except Exception:
        return False


def cpu_extract_fps(frames, nthreads, kind, params=None):
    """one frame per thread (the reference uses one thread per image, Frame.cc:258-261).
    kind "reference": oracle/_ref, the reference's ORBextractor::operator() compiled unmodified (process allocator);
    kind "port": the oracle restatement.  -> frames/s, wall s, per-stage seconds of thread 0 (port only)"""
    p = params or (EUROC["nfeatures"], EUROC["scale"], EUROC["levels"], EUROC["ini"], EUROC["min"])
    n = len(frames)
    if kind == "reference":
        from oracle import ref_py
        exs = [ref_py.PlainExtractor(*p) for _ in range(nthreads)]
    else:
        O, L = load_oracle_native()
        exs = [O.Extractor(*p, _lib=L) for _ in range(nthreads)]
    exs[0](frames[0])

    def work(t):
        for i in range(t, n, nthreads):
            exs[t](frames[i])

    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
    [t.start() for t in th]
    [t.join() for t in th]
    dt = time.perf_counter() - t0
    stage = exs[0].stage_seconds() if kind == "port" else None
    if kind == "reference":
        [e.close() for e in exs]
    return n / dt, dt, stage


def cpu_match_rate(q, dmap, nthreads):
    O, L = load_oracle_native()
    t0 = time.perf_counter()
    O.hamming_top2(q, dmap, popcnt=False, nthreads=nthreads, _lib=L)     # the reference's bit-hack distance
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    O.hamming_top2(q, dmap, popcnt=True, nthreads=nthreads, _lib=L)
    dt2 = time.perf_counter() - t0
    return len(q) * len(dmap) / dt, len(q) * len(dmap) / dt2


def cv2_primitives_ms(frame):
    """third column: OpenCV's own SIMD primitives (python cv2, one thread) on the stages the reference
    delegates to OpenCV -- pyramid (resize + copyMakeBorder), cv::FAST per 30-px cell incl. the minThFAST retry, GaussianBlur
    per level -- glued as src/ORBextractor.cc does (tests/cv2_restatement.py)."""
    try:
        import cv2
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import cv2_restatement as R
    except Exception as e:          # pragma: no cover
        return {"unavailable": str(e)}
    cv2.setNumThreads(1)
    p = R.Params(EUROC["nfeatures"], EUROC["scale"], EUROC["levels"], EUROC["ini"], EUROC["min"])

    def best(fn, n=5):
        ts = []
        for _ in range(n):
            t0 = time.perf_counter()
            fn()
            ts.append((time.perf_counter() - t0) * 1e3)
        return min(ts)

    levels = R.compute_pyramid(p, frame)
    det = {t: cv2.FastFeatureDetector_create(t, True) for t in (p.ini_th, p.min_th)}

    def fast_all():
        for padded in levels:
            roi = padded[19:-19, 19:-19]
            h, w = roi.shape
            width, height = w - 32, h - 32
            ncols, nrows = int(width / 30), int(height / 30)
            wc, hc = -(-width // ncols), -(-height // nrows)
            for i in range(nrows):
                y0 = 16 + i * hc
                if y0 >= h - 16 - 3:
                    continue
                y1 = min(y0 + hc + 6, h - 16)
                for j in range(ncols):
                    x0 = 16 + j * wc
                    if x0 >= w - 16 - 6:
                        continue
                    cell = roi[y0:y1, x0:min(x0 + wc + 6, w - 16)]
                    if not det[p.ini_th].detect(cell):
                        det[p.min_th].detect(cell)

    def blur_all():
        for padded in levels:
            cv2.GaussianBlur(padded[19:-19, 19:-19], (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)

    def fast_levels():
        for padded in levels:
            det[p.ini_th].detect(padded[19 + 13:-19 - 13, 19 + 13:-19 - 13])

    return {"pyramid": best(lambda: R.compute_pyramid(p, frame)), "fast_per_cell": best(fast_all, 3),
            "fast_whole_levels_iniTh": best(fast_levels, 3), "blur": best(blur_all), "opencv": cv2.__version__, "threads": 1,
            "note": "fast_per_cell makes ~1100 python calls per frame (about 8 us of interpreter per cell); "
                    "fast_whole_levels_iniTh is one cv::FAST call per level at iniThFAST, the SIMD floor of that stage"}


def run_reference(args):
    """--impl reference: the reference's own ORBextractor (oracle/_ref/libviorb_ref.so = /root/reference/src/ORBextractor.cc
    compiled unmodified against stand-in OpenCV headers whose primitives are bit-equal scalar restatements of OpenCV's) on
    all host cores, one frame per thread; the oracle port when that library is absent.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from viorb_b200 import synth
    cores = os.cpu_count() or 1
    kind = "reference" if reference_available() else "port"
    nsample = max(cores, min(args.cpu_frames, args.frames))
    frames = synth.frames(nsample, EUROC["rows"], EUROC["cols"], seed0=0)
    fps_all = []
    for _ in range(args.warmup):
        cpu_extract_fps(frames[:cores], cores, kind)
    t_ms = []
    for _ in range(args.steps):
        fps, dt, _ = cpu_extract_fps(frames, cores, kind)
        fps_all.append(fps)
        t_ms.append(dt * 1e3)
    fps = float(np.mean(fps_all))
    sample = "%d of %d frames per step (seeds 0..%d), one frame per thread" % (nsample, args.frames, nsample - 1)
    line = {
        "impl": "reference", "metric": "orb_frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(t_ms)), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample,
                         "what": REFERENCE_NOTE if kind == "reference" else "oracle port (oracle/_ref absent)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


REFERENCE_NOTE = ("oracle/_ref/libviorb_ref.so: the reference's src/ORBextractor.cc compiled unmodified (-O3 -march=x86-64-v3, "
                  "glibc malloc); its cv::resize/FAST/GaussianBlur are this repo's scalar stand-ins, bit-equal to OpenCV but "
                  "without OpenCV's SIMD (see cpu_baseline.cv2_primitives_ms for what those stages cost in OpenCV)")


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = "/tmp/viorb_clocks_%d.csv" % os.getpid()

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
        sm, mx, reasons = [], [], set()
        for ln in open(self.path):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v == "Active":
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # samples under load = upper half
        return {"sm_mhz": float(np.median(sorted(sm)[len(sm) // 2:])), "sm_max_mhz": float(max(mx)),
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa(local):
    """Pin this rank's threads (and therefore the first-touch placement of its pinned staging memory) to the NUMA node its
    GPU hangs off, before anything is allocated.  Returns a description for the JSON line."""
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True,
                             text=True, timeout=20).stdout.split()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else local
        bus = out[idx].strip().lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        if node < 0 or len(nodes) < 2:
            return {"bound": False, "why": "numa_node=%d, %d node(s) visible" % (node, len(nodes))}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"bound": False, "why": "node %d has no allowed cpus" % node}
        os.sched_setaffinity(0, cpus)
        return {"bound": True, "node": node, "cpus": len(cpus), "nodes": len(nodes)}
    except Exception as e:
        return {"bound": False, "why": "%s: %s" % (type(e).__name__, e)}


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = {"bound": False, "why": "--no-bind"} if args.no_bind else bind_to_gpu_numa(local)

    import torch
    import torch.distributed as dist
    from viorb_b200 import api, sharding, synth

    saved_stdout = None
    if world > 1:
        # NCCL prints its version banner on stdout at the first collective; keep stdout for the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    api.lib()
    # a dedicated (non-default) torch stream: the library launches on it and torch.cuda.Event times it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = api.Context(local, stream.cuda_stream)
    ex = api.ORBextractor(EUROC["nfeatures"], EUROC["scale"], EUROC["levels"], EUROC["ini"], EUROC["min"], ctx=ctx)
    if args.chunk:
        ex.configure(chunk_frames=args.chunk)
    rows, cols, cap = EUROC["rows"], EUROC["cols"], ex.cap

    # ---- synthetic frames of this rank's shard (seeds are global frame indices) ----
    f0, f1 = sharding.shard_range(args.frames, rank, world)
    nloc = f1 - f0
    h_imgs = api.pinned_empty((nloc, rows, cols), np.uint8)
    synth.frames(nloc, rows, cols, seed0=f0, out=h_imgs)
    h_kps = api.pinned_empty((nloc, cap), api.KEYPOINT)
    h_desc = api.pinned_empty((nloc, cap, 32), np.uint8)
    h_cnt = api.pinned_empty((nloc,), np.int32)
    d_imgs = torch.empty((nloc, rows, cols), dtype=torch.uint8, device=dev)
    d_imgs.copy_(torch.from_numpy(h_imgs))
    d_kps = torch.empty((nloc, cap, 7), dtype=torch.float32, device=dev)
    d_desc = torch.empty((nloc, cap, 32), dtype=torch.uint8, device=dev)
    d_cnt = torch.zeros((nloc,), dtype=torch.int32, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return reduce(x, dist.ReduceOp.MAX)

    def sum_over_ranks(x):
        return reduce(x, dist.ReduceOp.SUM)

    def step_device():
        ex.extract_batch_device(d_imgs, nloc, rows, cols, d_kps, d_desc, d_cnt)

    # ---- device-resident throughput (value) ----
    clocks = ClockSampler(local)
    clocks.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
    ex.check()
    barrier()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launch_count() - l0
    ex.check()
    # stage timing: the same K steps once more with the per-stage CUDA-event timers on.  With the timers on the
    # library runs its passes on one lane (no inter-pass overlap), so each stage's events bracket only its kernels.
    ex.profile(True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    ex.profile(False)
    prof_ms = e0.elapsed_time(e1)
    stage_ms, passes = ex.stage_ms()
    ex.check()
    clk = clocks.stop()
    value = args.frames * args.steps / (dev_ms * 1e-3)
    total_kp = int(d_cnt.sum().item())

    # ---- parity of the timed workload, device path: every frame's result against the reference's own digest ----
    golden = None
    try:
        golden = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_extract_batch4096.json")))
    except Exception:
        pass

    def digest_mismatches(kps, desc, cnt):
        """frames of this shard whose sha256(keypoints || descriptors) differs from the reference's"""
        bad = 0
        for i in range(nloc):
            g = f0 + i
            if g >= golden["frames"]:
                break
            n = int(cnt[i])
            if n != golden["n"][g] or frame_digest(kps[i, :n], desc[i, :n])[:16] != golden["digest16"][g]:
                bad += 1
        return bad

    parity = {}
    if golden is not None:
        k_host = d_kps.cpu().numpy().view(np.uint8).reshape(nloc, cap, 28).view(api.KEYPOINT).reshape(nloc, cap)
        bad_dev = int(sum_over_ranks(digest_mismatches(k_host, d_desc.cpu().numpy(), d_cnt.cpu().numpy())))
        del k_host
        parity["extraction_device_path"] = {"frames_compared": min(args.frames, golden["frames"]), "mismatching_frames": bad_dev}

    # ---- copy bound of this box at this N: the step's bytes with no kernel in between -- every rank copies its shard in
    #      (pinned -> device) and its results out (device -> pinned), all ranks concurrently; each direction alone, and both
    #      at once on two streams.  Where both at once take longer than one after the other (hosts whose device-to-host
    #      writes disturb the host-to-device reads), the library's serial copy order is the faster one and is selected. ----
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    t_imgs = torch.from_numpy(h_imgs)
    t_out = torch.from_numpy(h_desc)
    t_outk = torch.from_numpy(h_kps.view(np.uint8).reshape(nloc, cap, 28))
    d_kb = d_kps.view(torch.uint8).reshape(nloc, cap, 28)

    def copy_pass(do_in, do_out, pieces=8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record()
        s_in.wait_event(a)
        s_out.wait_event(a)
        step = (nloc + pieces - 1) // pieces
        for k in range(0, nloc, step):
            if do_in:
                with torch.cuda.stream(s_in):
                    d_imgs[k:k + step].copy_(t_imgs[k:k + step], non_blocking=True)
            if do_out:
                with torch.cuda.stream(s_out):
                    t_out[k:k + step].copy_(d_desc[k:k + step], non_blocking=True)
                    t_outk[k:k + step].copy_(d_kb[k:k + step], non_blocking=True)
        stream.wait_stream(s_in)
        stream.wait_stream(s_out)
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b))

    copy_pass(True, True)
    cb_both = min(copy_pass(True, True) for _ in range(4))
    cb_in = min(copy_pass(True, False) for _ in range(4))
    cb_out = min(copy_pass(False, True) for _ in range(4))
    serial = cb_both > 0.97 * (cb_in + cb_out) and not args.duplex
    cb = min(cb_both, cb_in + cb_out)
    h2d = nloc * rows * cols
    d2h = nloc * cap * 60 + nloc * 4
    copy_bound = {"value": args.frames / (cb * 1e-3), "unit": "frames/s", "ms_per_step": cb,
                  "h2d_and_d2h_concurrent_ms": cb_both, "h2d_only_ms": cb_in, "d2h_only_ms": cb_out,
                  "h2d_gb_per_s_all_ranks": h2d * world / (cb_in * 1e-3) / 1e9,
                  "copy_mode": "serial (viorb_extractor_set_copy_mode 1)" if serial else "duplex (default)",
                  "how": "the step's H2D and D2H bytes copied with cudaMemcpyAsync from / to pinned memory, no kernels, all %d rank(s) "
                         "at once, max over ranks, best of 4: each direction alone and both at once on two streams; the bound is the "
                         "faster of `both at once` and `one after the other`, and the library's copy order is chosen accordingly" % world}
    ex.set_copy_mode(1 if serial else 0)

    # ---- end to end through the host-buffer C ABI (e2e) ----
    for _ in range(2):
        ex.extract_batch(h_imgs, h_kps, h_desc, h_cnt)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        ex.extract_batch(h_imgs, h_kps, h_desc, h_cnt)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
    e2e = args.frames * args.steps / (e2e_ms * 1e-3)
    assert int(h_cnt.sum()) == total_kp, "host and device paths disagree"
    if golden is not None:
        bad_host = int(sum_over_ranks(digest_mismatches(h_kps, h_desc, h_cnt)))
        parity["extraction_host_path"] = {"frames_compared": min(args.frames, golden["frames"]), "mismatching_frames": bad_host}
        parity["extraction_golden"] = ("tests/golden/ref_extract_batch4096.json: sha256(keypoints || descriptors) per frame, written "
                                       "by the reference's ORBextractor.cc compiled unmodified (tests/golden/make_ref_batch_golden.py)")
        assert bad_dev == 0 and bad_host == 0, "extraction differs from the reference on %d / %d frames" % (bad_dev, bad_host)
    ex.set_copy_mode(0)

    # ---- per-call latency of the C-ABI entry points (configs[0]: one EuRoC frame; configs[1]: KITTI stereo pair) ----
    latency = None
    if rank == 0 and not args.no_latency:
        def med_ms(fn, n, warm):
            for _ in range(warm):
                fn()
            ts = []
            for _ in range(n):
                t0 = time.perf_counter()
                fn()
                ts.append((time.perf_counter() - t0) * 1e3)
            return float(np.median(ts))
        one = h_imgs[0]
        lat_frame = med_ms(lambda: ex(one), 100, 10)
        left, right, _ = synth.stereo_pair(376, 1241, 7)
        exl = api.ORBextractor(2000, 1.2, 8, 20, 7, ctx=ctx)
        exr = api.ORBextractor(2000, 1.2, 8, 20, 7, ctx=ctx)

        def stereo_pair():
            kl, dl = exl(left)
            kr, dr = exr(right)
            api.ComputeStereoMatches(exl, exr, kl, dl, kr, dr, 386.1448, 386.1448 / 718.856)

        lat_pair = med_ms(stereo_pair, 30, 5)
        exl.close()
        exr.close()
        latency = {"configs[0] one 752x480 frame, viorb_extract host->host, median ms": lat_frame,
                   "configs[1] KITTI 1241x376 stereo pair, 2 x viorb_extract + viorb_stereo_match, median ms": lat_pair,
                   "note": "C ABI from python; the C++ drop-in classes (ORBextractor::operator() with and without the "
                           "mvImagePyramid read, every matcher surface) are timed in `matchers`"}

    # ---- roofline of the dominant kernel (live CUDA-event stage timing inside the timed region) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    dom = max(stage_ms, key=stage_ms.get)
    kernel_names = {"pyramid": "pyr_level0_kernel+pyr_resize_kernel x7", "fast": "fast_cells_kernel",
                    "octree": "octree_kernel", "describe": "blur_levels_kernel+describe_blurred_kernel"}
    frames_timed = nloc * args.steps
    # per-kernel algorithmic bytes per frame (DESIGN.md "Kernels"): pyramid = read input + write padded pyramid;
    # fast = read every level ROI once + 4 B per candidate; octree = 6 B per candidate + 4 B per selected;
    # describe = blur (read the padded levels, write the blurred ROI pixels) + per keypoint the 31x31 level patch, the
    # 37x37 blurred patch and the 60 B record
    P, Ppad, ncand, nkp = 1117367, 1344493, 6085, 1006
    alg = {"pyramid": rows * cols + Ppad, "fast": P + 4 * ncand, "octree": 6 * ncand + 4 * nkp,
           "describe": Ppad + P + nkp * (31 * 31 + 37 * 37 + 60)}
    dom_ms_per_launch = stage_ms[dom] / max(passes, 1)
    frames_per_pass = frames_timed / max(passes, 1)
    achieved = alg[dom] * frames_per_pass / (dom_ms_per_launch * 1e-3) / 1e9
    step_achieved = BYTES_PER_FRAME_EUROC * value / world / 1e9
    # dram bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/traffic.json,
    # written by tools/ncu_summary.py; same launch shape: 128 frames per pass)
    traffic, traffic_src, pipes = None, None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        names = {"pyramid": ["pyr_level0_kernel", "pyr_resize_kernel"], "fast": ["fast_cells_kernel"],
                 "octree": ["octree_kernel"], "describe": ["blur_levels_kernel", "describe_blurred_kernel"]}[dom]
        if all(n in tj for n in names) and int(round(frames_per_pass)) == 128:
            traffic = float(sum(l["dram_bytes"] for n in names for l in tj[n]["launches"]))
            l0 = tj[names[0]]["launches"][0]
            if l0.get("alu_pipe_pct") is not None:
                pipes = {"alu_pipe_pct_of_peak": l0["alu_pipe_pct"], "issue_active_pct": l0["issue_active_pct"],
                         "dram_pct_of_peak": l0.get("dram_pct"),
                         "note": "same ncu capture: the kernel is bound by instruction issue on the integer (64 lanes/clk/SM "
                                 "measured, tools/ubench/pipes.cu) and shared-memory pipes, not by HBM"}
            traffic_src = "ncu --set full, %s (dram__bytes_read.sum + dram__bytes_write.sum per 128-frame launch)" % tj[names[0]]["source"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kernel_names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "pipes": pipes, "peak_source": peak_src,
                "alg_bytes_per_launch": alg[dom] * frames_per_pass, "avg_launch_ms": dom_ms_per_launch,
                "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
                "stage_timing": "separate pass of the same %d steps with per-stage CUDA events, single lane: %.2f ms/step "
                                "(the timed `value` region overlaps consecutive passes on four lanes)" % (args.steps, prof_ms / args.steps),
                "step": {"achieved": step_achieved, "frac": step_achieved / peak,
                         "bytes_per_frame": BYTES_PER_FRAME_EUROC,
                         "note": "whole hot path per GPU by SURVEY 8(d) compulsory bytes; latency/INT-bound by design"}}

    # the big batch buffers are not needed any more
    del d_imgs, d_kps, d_desc, d_kb, t_imgs, t_out, t_outk
    cpu_frames_sample = np.array(h_imgs[:min(args.cpu_frames, nloc)])
    for a in (h_imgs, h_kps, h_desc, h_cnt):
        api.pinned_free(a)
    torch.cuda.empty_cache()

    # ---- configs[1] / configs[3]: the other BASELINE shapes, device resident, rank 0 ----
    shapes = None
    if rank == 0 and not args.no_shapes:
        shapes = {}
        cores = os.cpu_count() or 1
        cpu_kind = "reference" if reference_available() else "port"
        for name, (h, w, nf, B, bytes_pf) in SHAPES.items():
            exs = api.ORBextractor(nf, 1.2, 8, 20, 7, ctx=ctx)
            uniq = min(B, 16)
            imgs = synth.frames(uniq, h, w, seed0=100)
            d = torch.from_numpy(np.ascontiguousarray(np.tile(imgs, (B // uniq, 1, 1)))).to(dev)
            scap = exs.cap
            s_kps = torch.empty((B, scap, 7), dtype=torch.float32, device=dev)
            s_desc = torch.empty((B, scap, 32), dtype=torch.uint8, device=dev)
            s_cnt = torch.zeros((B,), dtype=torch.int32, device=dev)
            chunk = max(1, min(128, (192 * 752 * 480 + h * w // 2) // (h * w)))      # the library's default pass size
            for _ in range(3):
                exs.extract_batch_device(d, B, h, w, s_kps, s_desc, s_cnt)
            exs.check()
            torch.cuda.synchronize()
            nst = 3
            e0.record()
            for _ in range(nst):
                exs.extract_batch_device(d, B, h, w, s_kps, s_desc, s_cnt)
            e1.record()
            torch.cuda.synchronize()
            exs.check()
            ms = e0.elapsed_time(e1) / nst
            fps = B / (ms * 1e-3)
            row = {"frames_per_s": fps, "frames_per_batch": B, "frames_per_pass": chunk, "ms_per_frame": ms / B,
                   "keypoints_per_frame": float(s_cnt.float().mean().item()), "mpix_per_s": fps * h * w / 1e6,
                   "roofline_step": {"bytes_per_frame": bytes_pf, "achieved": fps * bytes_pf / 1e9, "peak": peak, "unit": "GB/s",
                                     "frac": fps * bytes_pf / 1e9 / peak},
                   "l2": "batch input %.0f MB > 126 MB L2; no flush" % (B * h * w / 1e6)}
            if world == 1 and not args.no_cpu:
                ncpu = min(uniq, cores)
                cfps, cdt, _ = cpu_extract_fps(imgs[:ncpu], ncpu, cpu_kind, (nf, 1.2, 8, 20, 7))
                row["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": ncpu, "kind": cpu_kind,
                                       "sample": "%d frames, one per thread, %.2f s wall" % (ncpu, cdt)}
            shapes[name] = row
            exs.close()
            del d, s_kps, s_desc, s_cnt
            torch.cuda.empty_cache()
        if latency is not None:
            shapes["configs[1] KITTI 1241x376 nf2000"]["stereo_pair_ms"] = \
                latency["configs[1] KITTI 1241x376 stereo pair, 2 x viorb_extract + viorb_stereo_match, median ms"]

    # ---- matcher: config 5 exactly as SURVEY 8(d): map = default_rng(1234) bytes; 500 queries = map rows with <= 40 bit flips,
    #      500 fresh.  Map sharded by rows across ranks, NCCL all-gather of the top-2 records + merge kernel; the merged result
    #      is compared with the CPU oracle on the whole Q x M ----
    matcher = None
    if not args.no_matcher:
        M, Q = args.map, args.queries
        m0, m1 = sharding.shard_range(M, rank, world)
        h_map = np.random.default_rng(1234).integers(0, 256, (M, 32), dtype=np.uint8)      # the same on every rank
        rq = np.random.default_rng(1235)
        h_q = rq.integers(0, 256, (Q, 32), dtype=np.uint8)
        src_rows = rq.integers(0, M, Q // 2)
        for i, r in enumerate(src_rows):
            dsc = h_map[r].copy()
            for b in rq.integers(0, 256, int(rq.integers(0, 41))):
                dsc[b >> 3] ^= np.uint8(1 << (b & 7))
            h_q[i] = dsc
        d_map = torch.from_numpy(h_map[m0:m1]).to(dev)
        d_q = torch.from_numpy(h_q).to(dev)
        if rank != 0:
            del h_map
        mt = api.ORBmatcher(ctx=ctx)
        d_part = torch.zeros((Q, 4), dtype=torch.int32, device=dev)
        d_all = torch.zeros((world, Q, 4), dtype=torch.int32, device=dev)
        d_out = torch.zeros((Q, 4), dtype=torch.int32, device=dev)
        shard_bytes = 32 * (m1 - m0)
        flush = shard_bytes < 2 * L2_BYTES                 # a shard the L2 could keep between passes: flush it
        d_flush = torch.empty(2 * L2_BYTES, dtype=torch.uint8, device=dev) if flush else None

        def match_step(q=Q):
            mt.hamming_top2_device(d_q, q, d_map, m1 - m0, m0, d_part)
            if world > 1:
                dist.all_gather_into_tensor(d_all, d_part)
                mt.top2_merge_device(d_all, world, Q, d_out)

        def time_passes(fn, n):
            """ms per pass; with `flush` every pass is timed on its own after the L2 was overwritten"""
            if not flush:
                barrier()
                e0.record()
                for _ in range(n):
                    fn()
                e1.record()
                barrier()
                return max_over_ranks(e0.elapsed_time(e1)) / n
            tot = 0.0
            for _ in range(n):
                d_flush.fill_(1)
                barrier()
                e0.record()
                fn()
                e1.record()
                barrier()
                tot += max_over_ranks(e0.elapsed_time(e1))
            return tot / n

        for _ in range(3):
            match_step()
        m_ms = time_passes(match_step, max(args.steps, 3))
        pairs = Q * M / (m_ms * 1e-3)
        popc_peak = 148 * 16 * (clk.get("sm_max_mhz") or 1965.0) * 1e6
        l2_note = ("map shard %.0f MB per rank: fits the 126 MB L2, so 252 MB are written between passes and every pass is timed "
                   "on its own" % (shard_bytes / 1e6)) if flush else \
                  ("map shard %.0f MB per rank, streamed once per pass: larger than the 126 MB L2, no flush" % (shard_bytes / 1e6))
        matcher = {"workload": "configs[4]: %d queries (half = map rows with <= 40 bit flips, half fresh) x %d map descriptors "
                               "(default_rng(1234)), map sharded x%d" % (Q, M, world),
                   "value": pairs, "unit": "descriptor pairs/s", "ms_per_pass": m_ms,
                   "queries_per_s": Q / (m_ms * 1e-3), "l2": l2_note,
                   "hbm": {"achieved": (32 * M + 48 * Q) / world / (m_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                           "frac": (32 * M + 48 * Q) / world / (m_ms * 1e-3) / 1e9 / peak},
                   "popc": {"achieved": 8 * pairs / world, "peak": popc_peak, "unit": "popc32-equivalent/s per GPU",
                            "frac": 8 * pairs / world / popc_peak, "executed_frac": 5 * pairs / world / popc_peak,
                            "note": "algorithmic 8 popc32 per descriptor pair against 148 SMs x 16 POPC/clk x max SM clock; the "
                                    "kernel executes 5 POPC per pair (Harley-Seal carry-save compression), so frac may exceed 1 "
                                    "while executed_frac is the real POPC-pipe utilisation (INT-bound at Q=1000)"}}
        # parity at full size: the merged top-2 (device) against the CPU oracle over the whole map
        match_step()
        torch.cuda.synchronize()
        got = (d_out if world > 1 else d_part).cpu().numpy()
        if rank == 0:
            from oracle import oracle_py as O
            t0 = time.perf_counter()
            _, L = load_oracle_native()
            want = O.hamming_top2(h_q, h_map, popcnt=True, nthreads=os.cpu_count() or 1, _lib=L)
            chk_s = time.perf_counter() - t0
            ties = int((want["d1"] == want["d2"]).sum())
            bad = int(sum((got[:, j] != want[f]).sum() for j, f in enumerate(("d1", "i1", "d2", "i2"))))
            parity["matcher_top2"] = {"queries": Q, "map": M, "ranks": world, "mismatching_fields": bad,
                                      "queries_with_d1_eq_d2": ties,
                                      "checker": "CPU oracle orc_hamming_top2 over the whole map, %.1f s" % chk_s}
            assert bad == 0, "merged top-2 differs from the oracle in %d fields" % bad
            del h_map
        # the HBM-bound regime of the same search: a few queries per pass over the resident map shard
        small = []
        for qs in (1, 2, 4):
            for _ in range(2):
                mt.hamming_top2_device(d_q, qs, d_map, m1 - m0, m0, d_part)
            ms_q = time_passes(lambda: mt.hamming_top2_device(d_q, qs, d_map, m1 - m0, m0, d_part), 10)
            gbs = (32 * (m1 - m0) + 48 * qs) / (ms_q * 1e-3) / 1e9
            small.append({"queries": qs, "ms_per_pass": ms_q, "achieved": gbs, "unit": "GB/s per GPU", "frac": gbs / peak,
                          "pairs_per_s": qs * M / (ms_q * 1e-3)})
        matcher["hbm_bound_small_q"] = {"kernel": "hamming_top2_smallq_kernel", "peak": peak, "passes": small, "l2": l2_note}
        del d_map, d_flush
        torch.cuda.empty_cache()

    # ---- CPU baseline (rank 0, N == 1): the reference on the host cores, bounded sample; port and cv2 columns ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        ns = len(cpu_frames_sample)
        kind = "reference" if reference_available() else "port"
        fps, dt, _ = cpu_extract_fps(cpu_frames_sample, cores, kind)
        cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
               "sample": "%d of the %d frames (seeds 0..%d), one frame per thread, %.1f s wall" % (ns, args.frames, ns - 1, dt),
               "what": REFERENCE_NOTE if kind == "reference" else "oracle port (oracle/_ref absent)"}
        if kind == "reference":
            r1, _, _ = cpu_extract_fps(cpu_frames_sample[:max(8, ns // 32)], 1, "reference")
            cpu["single_thread_frames_per_s"] = r1
        pfps, pdt, _ = cpu_extract_fps(cpu_frames_sample[:max(cores, ns // 2)], cores, "port")
        fps1, dt1, st1 = cpu_extract_fps(cpu_frames_sample[:max(8, ns // 32)], 1, "port")
        cpu["port"] = {"value": pfps, "unit": "frames/s", "cores": cores, "kind": "port", "march": "native",
                       "single_thread_frames_per_s": fps1,
                       "stage_ms_single_thread": {k: v * 1e3 for k, v in st1.items()}}
        cpu["cv2_primitives_ms"] = cv2_primitives_ms(cpu_frames_sample[0])
        if matcher is not None:
            qs = synth.descriptor_map(64, seed=3)
            ms_ = synth.descriptor_map(1_000_000, seed=4)
            hack, pop = cpu_match_rate(qs, ms_, cores)
            matcher["cpu_baseline"] = {"value": hack, "unit": "descriptor pairs/s", "cores": cores, "kind": "port",
                                       "sample": "64 queries x 1M map descriptors (bit-hack distance of ORBmatcher.cc:1648)",
                                       "popcnt_variant": pop}

    # ---- ORBmatcher / Frame call surfaces through the C++ drop-in classes, beside the reference's own classes ----
    matchers = None
    if rank == 0 and world == 1 and not args.no_latency:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import matcher_bench
            matchers = {"rows": matcher_bench.measure(30),
                        "how": "tests/cpp/matcher_bench.cc built twice: against libviorb_b200.so (host containers in and out, every "
                               "copy and synchronisation inside the call) and against the reference's own sources "
                               "(oracle/_ref/matcher_bench_ref, one thread); KITTI-shape scenario, median of 30 calls, microseconds"}
        except Exception as e:
            matchers = {"unavailable": "%s: %s" % (type(e).__name__, e)}

    if rank == 0:
        line = {
            "metric": "orb_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args),
            "frames_per_pass": passes and int(round(frames_per_pass)),
            "clocks": clk,
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "ms_per_step": e2e_ms / args.steps, "copy_bound": copy_bound,
                    "frac_of_copy_bound": e2e / copy_bound["value"], "numa": numa},
            "gpu_launches": int(launches),
            "keypoints_per_frame": total_kp / nloc,
            "parity": parity,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "latency": latency,
            "shapes": shapes,
            "matchers": matchers,
            "matcher": matcher,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
