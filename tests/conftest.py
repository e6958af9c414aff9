"""pytest configuration: the `gpu` marker, repo-root imports, shared fixtures."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
FRAMES = os.path.join(GOLDEN, "frames")
CAL = os.path.join(GOLDEN, "calibration")
CANOPY_TS = ["2025-11-14T143013", "2025-11-14T143028", "2025-11-14T143037", "2025-11-14T143042", "2025-12-05T152733"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "reference_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def rs720():
    with open(os.path.join(CAL, "factory_color_intrinsics_1280_720.json")) as f:
        d = json.load(f)
    return dict(fx=d["fx"], fy=d["fy"], cx=d["ppx"], cy=d["ppy"], width=d["width"], height=d["height"])


def load_frame(ts):
    import cv2
    color = cv2.imread(os.path.join(FRAMES, f"canopy_capture_{ts}_HD.png"), cv2.IMREAD_COLOR)
    depth = cv2.imread(os.path.join(FRAMES, f"depth_snapshot_{ts}_HD.png"), cv2.IMREAD_UNCHANGED)
    assert color is not None and depth is not None and depth.dtype == np.uint16
    return color, depth


def blob_mask(h, w, seed):
    """Same generator as tests/golden/make_golden.py (kept in sync by test_oracle_golden)."""
    import cv2
    rng = np.random.default_rng(seed)
    m = np.zeros((h, w), np.uint8)
    for _ in range(4):
        c = (int(rng.integers(0, w)), int(rng.integers(0, h)))
        ax = (int(rng.integers(40, 220)), int(rng.integers(40, 160)))
        cv2.ellipse(m, c, ax, float(rng.uniform(0, 180)), 0, 360, 255, -1)
    return m
