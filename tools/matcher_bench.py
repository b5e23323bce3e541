#!/usr/bin/env python3
"""Per-call wall time of the ORBextractor / ORBmatcher / Frame call surfaces through the C++ drop-in classes, beside the
reference's own classes (compiled unmodified, CPU) on the same KITTI-shape scenario: builds tests/cpp/matcher_bench.cc
against libviorb_b200.so, runs it and the prebuilt oracle/_ref/matcher_bench_ref, prints one JSON document.
usage: tools/matcher_bench.py [reps]"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

EXE = os.path.join(ROOT, "tests", "cpp", "_build", "matcher_bench")
REF = os.path.join(ROOT, "oracle", "_ref", "matcher_bench_ref")


def build_product():
    from viorb_b200 import build
    lib = build.build_cuda()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    libdir = os.path.dirname(lib)
    deps = [lib, os.path.join(ROOT, "tests", "cpp", "matcher_bench.cc"), os.path.join(ROOT, "tests", "cpp", "pose_scenarios.h")]
    hostdir = os.path.join(ROOT, "viorb_b200", "host")
    deps += [os.path.join(hostdir, f) for f in os.listdir(hostdir)]
    if os.path.exists(EXE) and all(os.path.getmtime(EXE) >= os.path.getmtime(d) for d in deps):
        return EXE
    subprocess.check_call([build.CXX, "-O2", "-std=gnu++17", "-I", os.path.join(ROOT, "viorb_b200", "host"), "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "matcher_bench.cc"), "-o", EXE, "-L", libdir, "-lviorb_b200",
                           "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64"])
    return EXE


def write_pair(path):
    from viorb_b200 import synth
    left, right, _ = synth.stereo_pair(376, 1241, 7)
    with open(path, "wb") as f:
        f.write(left.tobytes())
        f.write(right.tobytes())


def run(exe, pair, reps):
    out = subprocess.run([exe, pair, str(reps)], capture_output=True, text=True, timeout=1200)
    if out.returncode != 0:
        raise RuntimeError("%s failed: %s" % (exe, out.stderr[-2000:]))
    return [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]


def measure(reps=30, product=True):
    """-> list of {surface, gpu_us, cpu_reference_us, speedup, n_gpu, n_cpu}; cpu only where the reference binary exists"""
    with tempfile.TemporaryDirectory() as d:
        pair = os.path.join(d, "pair.bin")
        write_pair(pair)
        gpu = run(build_product(), pair, reps) if product else []
        cpu = run(REF, pair, max(3, reps // 3)) if os.path.exists(REF) else []
    by = {r["surface"]: r for r in cpu}
    rows = []
    for r in gpu or cpu:
        c = by.get(r["surface"])
        g = r if gpu else None
        rows.append({"surface": r["surface"],
                     "gpu_us": g["us_median"] if g else None, "gpu_us_min": g["us_min"] if g else None,
                     "cpu_reference_us": c["us_median"] if c else None,
                     "speedup": round(c["us_median"] / g["us_median"], 2) if (c and g) else None,
                     "n_gpu": g["n"] if g else None, "n_cpu": c["n"] if c else None})
    return rows


if __name__ == "__main__":
    rows = measure(int(sys.argv[1]) if len(sys.argv) > 1 else 30, product="--cpu-only" not in sys.argv)
    print(json.dumps({"matchers": rows}, indent=1))
