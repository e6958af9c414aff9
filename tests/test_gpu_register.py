"""K2 parity on the B200: depth -> colour registration (z-buffer scatter) through the C ABI against the C oracle
(oracle/oracle.c, librealsense align semantics).  Integer result: bit-exact depth AND winners."""
import json
import os

import numpy as np
import pytest

from conftest import CAL
from synth import synth_depth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rv():
    import torch
    assert torch.cuda.is_available()
    import repas_vision_b200 as rv
    return rv


def _femto(rv):
    c = rv.load_camera(os.path.join(CAL, "factory_color_intrinsics_2025-09-08T143506.json"))
    d = rv.load_camera(os.path.join(CAL, "factory_depth_intrinsics_2025-09-08T143506.json"))
    return d, c


def _pose(pitch_deg, baseline):
    a = np.deg2rad(pitch_deg)
    R = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    return R, np.array([baseline, -0.002, 0.004])


def _check(rv, depth, dcam, ccam, R, t, units=0.001):
    from oracle import oracle_c
    single = depth.ndim == 2
    out, win = rv.register_depth_to_color(depth, dcam, ccam, R, t, depth_units=units, return_winner=True)
    frames = depth[None] if single else depth
    outs = out[None] if single else out
    wins = win[None] if single else win
    for b in range(frames.shape[0]):
        ro, rw = oracle_c.register_depth_to_color(frames[b], dcam.as_dict(), ccam.as_dict(), np.asarray(R).T.reshape(9), t, units)
        assert np.array_equal(outs[b], ro), f"frame {b}: aligned depth differs in {(outs[b] != ro).sum()} pixels"
        assert np.array_equal(wins[b], rw), f"frame {b}: winners differ in {(wins[b] != rw).sum()} pixels"
    return outs


def test_femto_576_into_720p_identity_extrinsics(rv):
    """What the on-disk Femto calibration describes: 640x576 ToF -> 1280x720 colour, R = I, t = 0."""
    dcam, ccam = _femto(rv)
    R, t = rv.load_extrinsics(os.path.join(CAL, "factory_extrinsics_d2c_2025-09-08T143506.json"))
    depth = synth_depth(576, 640, 3)
    out = _check(rv, depth, dcam, ccam, R, t)
    assert (out > 0).mean() > 0.2


def test_femto_480_crop_into_720p_with_baseline(rv):
    """BASELINE config 3: 640x480 depth (centre crop of the 640x576 intrinsics: cy - 48) with a 32 mm baseline, 6 deg pitch."""
    dcam, ccam = _femto(rv)
    dcam = rv.Camera(dcam.fx, dcam.fy, dcam.cx, dcam.cy - 48.0, 640, 480)
    R, t = _pose(6.0, 0.032)
    depth = np.stack([synth_depth(480, 640, 10 + i) for i in range(3)])
    _check(rv, depth, dcam, ccam, R, t)


def test_realsense_factory_extrinsics_same_resolution(rv):
    """RealSense 640x480 -> 640x480 with the factory d2c extrinsics (15 mm baseline)."""
    cam = rv.load_camera(os.path.join(CAL, "factory_color_intrinsics_640_480.json"))
    R, t = rv.read_depth_to_color_extrinsics(os.path.join(CAL, "factory_d2c_extrinsics.json"))
    depth = synth_depth(480, 640, 5)
    _check(rv, depth, cam, cam, R, t)


def test_distorted_colour_camera_and_far_near_collisions(rv):
    """Forward distortion on projection + a scene built so that many depth pixels land on the same colour pixel."""
    dcam, _ = _femto(rv)
    dist = (0.09217283086787045, -0.11526566137629435, 0.0013528051910911107, 0.001999155652236615, 0.04589787600037745)
    ccam = rv.Camera(765.924059488859, 765.4664588620507, 646.6240000261831, 365.8248519689665, 1280, 720, dist,
                     "modified_brown_conrady")
    rng = np.random.default_rng(9)
    depth = np.where(rng.random((576, 640)) < 0.5, 500, 3000).astype(np.uint16)
    depth[rng.random((576, 640)) < 0.1] = 0
    R, t = _pose(-4.0, 0.05)
    _check(rv, depth, dcam, ccam, R, t)
    inv = rv.Camera(dcam.fx, dcam.fy, dcam.cx, dcam.cy, 640, 576, (0.01, -0.02, 0.001, -0.001, 0.003), "inverse_brown_conrady")
    _check(rv, depth, inv, ccam, R, t)
    bc = rv.Camera(dcam.fx, dcam.fy, dcam.cx, dcam.cy, 640, 576, (0.01, -0.02, 0.001, -0.001, 0.003), "brown_conrady")
    _check(rv, depth, bc, rv.Camera(ccam.fx, ccam.fy, ccam.cx, ccam.cy, 1280, 720, dist, "brown_conrady"), R, t)


def test_edge_cases(rv):
    dcam, ccam = _femto(rv)
    R, t = np.eye(3), np.zeros(3)
    z = np.zeros((576, 640), np.uint16)
    out, win = rv.register_depth_to_color(z, dcam, ccam, R, t, return_winner=True)
    assert not out.any() and (win == -1).all()
    sat = np.full((576, 640), 65535, np.uint16)
    _check(rv, sat, dcam, ccam, R, np.array([0.0, 0.0, -70.0]))  # everything behind / at the camera plane
    _check(rv, sat, dcam, ccam, R, t)
    # up-scaling (splat rectangles > 1 px) and down-scaling targets
    small = rv.Camera(ccam.fx / 4, ccam.fy / 4, ccam.cx / 4, ccam.cy / 4, 320, 180)
    _check(rv, synth_depth(576, 640, 77), dcam, small, R, t)
    big = rv.Camera(ccam.fx * 1.5, ccam.fy * 1.5, ccam.cx * 1.5, ccam.cy * 1.5, 1920, 1080)
    _check(rv, synth_depth(576, 640, 78), dcam, big, R, t)


def test_batch_larger_than_one_chunk_and_idempotence(rv):
    import torch
    dcam, ccam = _femto(rv)
    dcam = rv.Camera(dcam.fx, dcam.fy, dcam.cx, dcam.cy - 48.0, 640, 480)
    R, t = _pose(2.0, 0.02)
    frames = np.stack([synth_depth(480, 640, 100 + i) for i in range(4)])
    depth = np.concatenate([frames] * 5)  # 20 frames > the 8-frame L2 chunk
    d = torch.from_numpy(depth).cuda()
    out = rv.register_depth_to_color(d, dcam, ccam, R, t)
    assert out.is_cuda and tuple(out.shape) == (20, 720, 1280)
    o = out.cpu().numpy()
    from oracle import oracle_c
    for b in range(4):
        ro, _ = oracle_c.register_depth_to_color(frames[b], dcam.as_dict(), ccam.as_dict(), R.T.reshape(9), t)
        for rep in range(5):
            assert np.array_equal(o[b + 4 * rep], ro)
    # a second pass over an already aligned image (equal cameras, identity extrinsics) still matches the oracle
    _check(rv, ro, ccam, ccam, np.eye(3), np.zeros(3))


def test_full_size_batch_properties(rv):
    """BASELINE configs[2] at batch size (64 x 640x480 -> 1280x720) through size-independent properties: every filled
    colour pixel holds exactly the depth of the source pixel named as its winner, empty pixels have no winner, the
    depth-only call equals the call with winners, and a frame gives the same image wherever it sits in the batch."""
    import torch
    dcam, ccam = _femto(rv)
    dcam = rv.Camera(dcam.fx, dcam.fy, dcam.cx, dcam.cy - 48.0, 640, 480)
    R, t = _pose(6.0, 0.032)
    base = np.stack([synth_depth(480, 640, 300 + i) for i in range(8)])
    depth = torch.from_numpy(np.concatenate([base] * 8)).cuda()  # 64 frames: four 16-frame chunks
    out, win = rv.register_depth_to_color(depth, dcam, ccam, R, t, return_winner=True)
    assert out.is_cuda and tuple(out.shape) == (64, 720, 1280)
    filled = win >= 0
    assert torch.equal(filled, out != 0)
    src = depth.view(64, -1).to(torch.int32)
    taken = torch.gather(src, 1, win.view(64, -1).clamp(min=0).to(torch.int64)).view(64, 720, 1280)
    assert torch.equal(torch.where(filled, taken, torch.zeros_like(taken)), out.to(torch.int32))
    assert float(filled.float().mean()) > 0.2
    only = rv.register_depth_to_color(depth, dcam, ccam, R, t)
    assert torch.equal(only, out)
    for rep in range(1, 8):
        assert torch.equal(out[rep * 8:(rep + 1) * 8], out[:8]) and torch.equal(win[rep * 8:(rep + 1) * 8], win[:8])
