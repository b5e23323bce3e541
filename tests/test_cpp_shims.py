"""The C++ drop-in classes (viorb_b200/host: ORB_SLAM2::ORBextractor, ORBmatcher, Frame::ComputeStereoMatches)
compile against the compat headers everywhere, and on a GPU box reproduce the oracle through the same calls
the reference's Frame / Tracking / LocalMapping make (tests/cpp/test_shims.cc)."""
import os
import subprocess

import pytest

from util import ROOT

EXE = os.path.join(ROOT, "tests", "cpp", "_build", "test_shims")


def build_exe():
    from viorb_b200 import build
    lib = build.build_cuda()
    build.build_synth()
    olib = build.build_oracle()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "test_shims.cc")
    libdir, odir = os.path.dirname(lib), os.path.dirname(olib)
    cmd = [build.CXX, "-O1", "-std=gnu++17", "-I", os.path.join(ROOT, "viorb_b200", "host"), "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "oracle"), src, "-o", EXE, "-L", libdir, "-lviorb_b200", "-lviorb_synth", "-L", odir,
           "-lorb_oracle", "-Wl,-rpath," + libdir, "-Wl,-rpath," + odir, "-Wl,-rpath,/usr/local/cuda/lib64"]
    subprocess.check_call(cmd)
    return EXE


def test_shims_compile_and_link():
    exe = build_exe()
    assert os.path.exists(exe)
    # the library exports the reference's C++ entry points (mangled names of SURVEY.md section 8(b), modulo cv::Mat
    # standing in for cv::_InputArray when built against the compat header)
    out = subprocess.check_output("nm -D --defined-only %s | c++filt" % os.path.join(ROOT, "viorb_b200", "lib", "libviorb_b200.so"),
                                  shell=True, text=True)
    for sym in ("ORB_SLAM2::ORBextractor::ORBextractor(int, float, int, int, int)", "ORB_SLAM2::ORBextractor::operator()",
                "ORB_SLAM2::ORBmatcher::ORBmatcher(float, bool)", "ORB_SLAM2::ORBmatcher::DescriptorDistance(cv::Mat const&, cv::Mat const&)",
                "ORB_SLAM2::ORBmatcher::SearchByProjection(ORB_SLAM2::Frame&, std::vector<ORB_SLAM2::MapPoint*",
                "ORB_SLAM2::ORBmatcher::SearchByProjection(ORB_SLAM2::Frame&, ORB_SLAM2::Frame const&, float, bool)",
                "ORB_SLAM2::ORBmatcher::SearchForTriangulation(ORB_SLAM2::KeyFrame*, ORB_SLAM2::KeyFrame*, cv::Mat",
                "ORB_SLAM2::ORBmatcher::SearchByBoW(ORB_SLAM2::KeyFrame*, ORB_SLAM2::Frame&, std::vector<ORB_SLAM2::MapPoint*",
                "ORB_SLAM2::ORBmatcher::SearchByBoW(ORB_SLAM2::KeyFrame*, ORB_SLAM2::KeyFrame*, std::vector<ORB_SLAM2::MapPoint*",
                "ORB_SLAM2::ORBmatcher::SearchForInitialization(ORB_SLAM2::Frame&, ORB_SLAM2::Frame&, std::vector<cv::Point2",
                "ORB_SLAM2::ORBmatcher::SearchBySim3(ORB_SLAM2::KeyFrame*, ORB_SLAM2::KeyFrame*, std::vector<ORB_SLAM2::MapPoint*",
                "ORB_SLAM2::ORBmatcher::Fuse(ORB_SLAM2::KeyFrame*, std::vector<ORB_SLAM2::MapPoint*",
                "ORB_SLAM2::ORBmatcher::Fuse(ORB_SLAM2::KeyFrame*, cv::Mat, std::vector<ORB_SLAM2::MapPoint*",
                "ORB_SLAM2::ORBVocabulary::loadFromTextFile(", "ORB_SLAM2::ORBVocabulary::transform(",
                "ORB_SLAM2::Frame::ComputeStereoMatches()"):
        assert sym in out, sym


@pytest.mark.gpu
def test_shims_match_oracle_on_gpu():
    exe = build_exe()
    p = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "ALL SHIM CHECKS PASSED" in p.stdout, p.stdout[-3000:] + p.stderr[-2000:]
