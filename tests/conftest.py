import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the engine library and the checker exist (a no-op when already built)."""
    from rayaccel_b200 import build as engine_build
    engine_build.build()
    import oracle
    if not os.path.exists(oracle.ORACLE_SO):
        oracle.build(ref=False)
    if not oracle.have_ref() and os.path.isdir("/root/reference/RayAccelerator"):
        oracle.build(ref=True)


@pytest.fixture(scope="session")
def battlefield():
    import rayaccel_b200 as rb
    return rb.load_scene()


@pytest.fixture(scope="session")
def battlefield_images(battlefield):
    """Host-built scene images of battlefield.bin wrapped for the oracle (with the light probe)."""
    import oracle
    import rayaccel_b200 as rb
    img = rb.HostImages(battlefield.vertices, battlefield.indices)
    return oracle.SceneImages(img.nodes, img.pairs, img.remap, battlefield.environment)


def make_rays(origins, dirs, tmin=0.0, tmax=1e6):
    import oracle
    origins = np.asarray(origins, np.float32).reshape(-1, 3)
    dirs = np.asarray(dirs, np.float32).reshape(-1, 3)
    r = np.zeros(origins.shape[0], dtype=oracle.RAY_DTYPE)
    r["origin"], r["dir"], r["minT"], r["maxT"] = origins, dirs, tmin, tmax
    return r


def random_rays(n, lo, hi, seed):
    rng = np.random.default_rng(seed)
    o = rng.uniform(lo, hi, size=(n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return make_rays(o, d)


def primary_rays_numpy(cam, width, height):
    """Pixel-centre primary rays, same formula as csrc/raygen.cu (not bit-identical; only used on CPU)."""
    ys, xs = np.meshgrid(np.arange(height, dtype=np.float32) + 0.5, np.arange(width, dtype=np.float32) + 0.5, indexing="ij")
    d = cam.view[None, None, :] + cam.up[None, None, :] * ys[..., None] + cam.right[None, None, :] * xs[..., None]
    d = d / np.linalg.norm(d, axis=2, keepdims=True)
    return make_rays(np.broadcast_to(cam.origin, (width * height, 3)), d.reshape(-1, 3).astype(np.float32))
