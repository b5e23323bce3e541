"""In-tree native builds.

  libviorb_b200.so   -- the product: hand-written sm_100a CUDA kernels + the extern "C" boundary
                        (include/viorb_gpu.h) + the C++ host shims.  nvcc cross-compiles without a GPU.
  libviorb_synth.so  -- host-only synthetic input generator (bench / test input, g++).

The oracle (oracle/) is test infrastructure and is built by oracle/Makefile, not here.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "viorb_b200", "csrc")
HOST = os.path.join(ROOT, "viorb_b200", "host")
LIBDIR = os.path.join(ROOT, "viorb_b200", "lib")

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
CXX = os.environ.get("CXX") or shutil.which("g++") or "g++"

CUDA_SOURCES = ["extractor_kernels.cu", "matcher_kernels.cu", "c_api.cu", "c_api_match.cu", "search_kernels.cu", "bow_kernels.cu"]
HOST_SOURCES = ["ORBextractor.cc", "ORBmatcher.cc", "ORBVocabulary.cc"]   # C++ shims compiled into the same library

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                       # float steps must round like the reference (no FMA)
    "-Xcompiler", "-fPIC,-O3,-ffp-contract=off,-Wall",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-I", HOST,
    "--shared", "-lcudart",
]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _deps(folder):
    out = []
    for base, _, files in os.walk(folder):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".hpp", ".cpp", ".cc", ".inc")):
                out.append(os.path.join(base, f))
    return out


def build_cuda(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    target = os.path.join(LIBDIR, "libviorb_b200.so")
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    srcs += [os.path.join(HOST, s) for s in HOST_SOURCES if os.path.exists(os.path.join(HOST, s))]
    deps = _deps(CSRC) + _deps(HOST) + _deps(os.path.join(ROOT, "include"))
    if force or _stale(target, deps):
        cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", target] + srcs
        print("[viorb build]", " ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
    return target


def build_synth(force=False):
    os.makedirs(LIBDIR, exist_ok=True)
    target = os.path.join(LIBDIR, "libviorb_synth.so")
    src = os.path.join(CSRC, "synth.cpp")
    if force or _stale(target, [src]):
        cmd = [CXX, "-O3", "-fPIC", "-shared", "-std=c++14", "-pthread", "-o", target, src]
        print("[viorb build]", " ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
    return target


def build_oracle(force=False, march=None, outdir=None):
    """Builds the CPU oracle -- the checker, test infrastructure only."""
    odir = os.path.join(ROOT, "oracle")
    cmd = ["make", "-C", odir] + (["-B"] if force else [])
    if march:
        cmd.append("MARCH=%s" % march)
    if outdir:
        cmd.append("OUT=%s" % outdir)
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return os.path.join(odir, outdir or "_build", "liborb_oracle.so")


def build_reference_oracle(reference_root="/root/reference"):
    """oracle/_ref/libviorb_ref.so: the reference's own sources compiled unmodified (oracle/refbuild/Makefile) -- the
    checker's checker.  Only where the reference tree exists (the build container); returns None elsewhere."""
    if not os.path.isdir(os.path.join(reference_root, "src")):
        return None
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle", "refbuild"), "REF=" + reference_root], stdout=subprocess.DEVNULL)
    return os.path.join(ROOT, "oracle", "_ref", "libviorb_ref.so")


def build_all(force=False, verbose=False):
    return build_cuda(force, verbose), build_synth(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
