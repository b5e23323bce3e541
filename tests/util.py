"""Shared helpers for the test-suite."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_pattern():
    """512 sampling points (x, y) as int array [512, 2] parsed from include/viorb_orb_pattern.h"""
    txt = open(os.path.join(ROOT, "include", "viorb_orb_pattern.h")).read()
    body = txt.split("VIORB_ORB_PATTERN_INIT {", 1)[1].split("}", 1)[0]
    nums = np.array([int(v) for v in re.findall(r"-?\d+", body)], np.int8)
    assert nums.size == 1024
    return nums.reshape(512, 2).astype(np.int32), nums


CONFIGS = {
    # name: (h, w, nfeatures, scale, levels, iniTh, minTh)   -- BASELINE.json configs / reference YAMLs
    "euroc": (480, 752, 1000, 1.2, 8, 20, 7),     # Examples/Monocular/EuRoC.yaml:29-42
    "kitti": (376, 1241, 2000, 1.2, 8, 20, 7),    # Examples/Stereo/KITTI00-02.yaml:38-51
    "hd": (1080, 1920, 5000, 1.2, 8, 20, 7),
    "uhd": (2160, 3840, 5000, 1.2, 8, 20, 7),
    "odd": (360, 640, 777, 1.2, 8, 20, 7),
    "kitti12": (376, 1241, 2000, 1.2, 8, 12, 7),  # Examples/Stereo/KITTI04-12.yaml:50
}
