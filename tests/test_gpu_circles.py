"""GPU front end of the circle finder (csrc/circles.cu) against the reference's own OpenCV / NumPy
calls (oracle/circles.py): bit-exact at every stage."""
import numpy as np
import pytest
import torch

from oracle import circles as oc
from test_circles_host import synthetic_discs

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


def dev(a, device):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32, np.float64])
def test_to_uint8(cuda_device, dtype):
    from magnify_b200 import circles as mc

    rng = np.random.default_rng(1)
    cases = [rng.integers(3, 250, (37, 91)).astype(dtype), (rng.random((64, 64)) * 5000 + 100).astype(dtype),
             np.full((5, 7), 9).astype(dtype), np.zeros((0, 4), dtype)]
    if np.dtype(dtype).kind == "f":
        cases.append((rng.standard_normal((33, 17)) * 1e3).astype(dtype))
    for arr in cases:
        got = mc.to_uint8(dev(arr, cuda_device)).cpu().numpy()
        np.testing.assert_array_equal(got, oc.to_uint8(arr))


SHAPES = [(1, 1), (1, 9), (2, 2), (3, 5), (4, 33), (31, 32), (33, 65), (72, 72), (100, 100), (257, 511), (700, 900)]


@pytest.mark.parametrize("shape", SHAPES)
def test_gradients_quantiles_canny_match_opencv(cuda_device, shape):
    from magnify_b200 import circles as mc

    h, w = shape
    rng = np.random.default_rng(h * 1000 + w)
    discs = [(int(rng.integers(0, h)), int(rng.integers(0, w)), int(rng.integers(2, 30))) for _ in range(6)]
    img = oc.to_uint8(synthetic_discs(h, w, discs, seed=h + w)) if h * w > 1 else np.array([[7]], np.uint8)
    for low_q, high_q in ((0.1, 0.9), (0.5, 0.99), (0.0, 1.0), (0.9, 0.1)):
        want = oc.edge_stages(img, low_q, high_q)
        dx, dy = mc.edge_gradients(dev(img, cuda_device))
        np.testing.assert_array_equal(dx.cpu().numpy(), want["dx"])         # float32 holding exact integers
        np.testing.assert_array_equal(dy.cpu().numpy(), want["dy"])
        low, high = mc.gradient_quantiles(dx, dy, (low_q, high_q))
        assert low == want["low"] and high == want["high"], (low, want["low"], high, want["high"])
        assert low.dtype == want["low"].dtype
        edges, sweeps = mc.canny(dx, dy, low, high, return_sweeps=True)
        np.testing.assert_array_equal(edges.cpu().numpy(), want["edges"])
        assert sweeps >= 1
        e2, _, _ = mc.find_edges(dev(img, cuda_device), low_q, high_q)
        np.testing.assert_array_equal(e2.cpu().numpy(), want["edges"])


def test_canny_long_spiral_needs_many_sweeps(cuda_device):
    """Hysteresis must follow a weak edge across many tiles from a single strong seed."""
    from magnify_b200 import circles as mc

    h = w = 400
    img = np.zeros((h, w), np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    rr = np.hypot(yy - 200, xx - 200)
    theta = np.arctan2(yy - 200, xx - 200)
    spiral = np.abs(rr - (20 + 25 * (theta + np.pi) / (2 * np.pi) + 25 * np.round((rr - 20) / 25 - (theta + np.pi) / (2 * np.pi)))) < 2.5
    img[spiral] = 60
    img[195:205, 225:235][spiral[195:205, 225:235]] = 255          # one bright stretch
    want = oc.edge_stages(img, 0.5, 0.999)
    dx, dy = mc.edge_gradients(dev(img, cuda_device))
    edges, sweeps = mc.canny(dx, dy, want["low"], want["high"], return_sweeps=True)
    np.testing.assert_array_equal(edges.cpu().numpy(), want["edges"])
    assert want["edges"].sum() > 1000 and sweeps > 2


def test_full_size_tile(cuda_device):
    from magnify_b200 import circles as mc

    rng = np.random.default_rng(5)
    discs = [(int(rng.integers(0, 2048)), int(rng.integers(0, 2048)), int(rng.integers(8, 26))) for _ in range(300)]
    raw = synthetic_discs(2048, 2048, discs, seed=3)
    u8 = mc.to_uint8(dev(raw, cuda_device))
    np.testing.assert_array_equal(u8.cpu().numpy(), oc.to_uint8(raw))
    want = oc.edge_stages(oc.to_uint8(raw), 0.1, 0.9)
    edges, dx, dy = mc.find_edges(u8, 0.1, 0.9)
    np.testing.assert_array_equal(edges.cpu().numpy(), want["edges"])
