"""Device-side Whitted renderer (rayaccel_b200/csrc/whitted.cu) through the C-ABI (racc_cuda_whitted_trace) against its
checker oracle_whitted_trace -- bit for bit (32.32 fixed-point accumulation is order-independent) -- and against the image
of the reference's UNMODIFIED WhittedRenderer (tests/golden/ref_whitted_tiles.npz). All tests need a GPU. The file sorts
last among the GPU tests on purpose: it is the newest device code."""
import os

import numpy as np
import pytest

import oracle
import rayaccel_b200 as rb
from test_render_oracle import camera_for, synthetic_shading_case, tile_means

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def world(battlefield):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    torch.cuda.set_device(0)
    rb.init(0)
    scene = rb.create_scene(battlefield.vertices, battlefield.indices)
    env = rb.create_environment(battlefield.environment)
    shading = rb.create_shading(battlefield.normals, battlefield.triangle_normals, battlefield.materials)
    nodes, pairs, remap = scene.download()
    images = oracle.SceneImages(nodes, pairs, remap, battlefield.environment)
    sh = oracle.Shading(battlefield.indices, battlefield.normals, battlefield.triangle_normals, battlefield.materials)
    yield scene, env, shading, images, sh
    shading.destroy(); env.destroy(); scene.destroy()


@pytest.mark.parametrize("width,height,spp,depth,seed,batch", [
    (256, 128, 2, 8, 11, 0),  # the reference's settings for this renderer (depth 8)
    (256, 128, 3, 8, 11, 1),  # one sample per batch, three batches
    (200, 96, 2, 2, 3, 0),    # cut by the depth limit, sizes that are no multiple of the CTA
    (64, 64, 1, 0, 0, 0),     # depth 0, pixel centres: the light probe seen directly
    (33, 17, 1, 1, 9, 0),
])
def test_device_whitted_equals_oracle(world, battlefield, width, height, spp, depth, seed, batch):
    scene, env, shading, images, sh = world
    cam = camera_for(battlefield, width, height)
    want, want_waves = oracle.whitted_trace(images, sh, cam, width, height, spp, depth, seed)
    got, waves = rb.whitted_trace(scene, env, shading, cam, width, height, spp, depth, seed, batch_spp=batch)
    assert waves == [int(x) for x in want_waves], "rays traced per bounce differ"
    bad = np.flatnonzero((got.view(np.uint32) != want.view(np.uint32)).reshape(-1, 4).any(axis=1))
    assert bad.size == 0, f"{bad.size} of {width * height} pixels differ, first {bad[:5]}: {got.reshape(-1, 4)[bad[:3]]} vs {want.reshape(-1, 4)[bad[:3]]}"


def test_device_whitted_on_synthetic_scene(world):
    """Skewed and degenerate shading normals, triangles seen from both sides (both eta branches, total internal reflection)."""
    verts, indices, normals, tri_normals, tri_materials, materials, env_img, cam = synthetic_shading_case()
    scene = rb.create_scene(verts, indices)
    env = rb.create_environment(env_img)
    shading = rb.create_shading(normals, tri_normals, tri_materials, materials)
    nodes, pairs, remap = scene.download()
    images = oracle.SceneImages(nodes, pairs, remap, env_img)
    sh = oracle.Shading(indices, normals, tri_normals, tri_materials, materials)
    want, want_waves = oracle.whitted_trace(images, sh, cam, 96, 64, 2, 8, 5)
    got, waves = rb.whitted_trace(scene, env, shading, cam, 96, 64, 2, 8, 5)
    assert waves == [int(x) for x in want_waves]
    assert got.tobytes() == want.tobytes()
    shading.destroy(); env.destroy(); scene.destroy()


def test_device_whitted_matches_reference_renderer_image_and_device_framebuffer(world, battlefield):
    scene, env, shading, _, _ = world
    g = np.load(os.path.join(GOLDEN, "ref_whitted_tiles.npz"))
    w, h, tile, depth = int(g["width"]), int(g["height"]), int(g["tile"]), int(g["max_depth"])
    spp = 32
    cam = camera_for(battlefield, w, h)
    fb, waves = rb.whitted_trace(scene, env, shading, cam, w, h, spp, depth, seed=11)
    got, ref = tile_means(fb, spp, tile), g["tiles"].astype(np.float64)
    assert abs(got.mean() / ref.mean() - 1.0) < 1e-3
    assert (np.abs(got - ref) / np.maximum(ref, 1e-2)).max() < 0.02
    assert abs(sum(waves) / spp / float(g["rays_per_frame"]) - 1.0) < 1e-3
    dev = torch.zeros(h * w * 4, dtype=torch.float32, device="cuda")
    _, waves2 = rb.whitted_trace(scene, env, shading, cam, w, h, spp, depth, seed=11, framebuffer_ptr=dev.data_ptr(), batch_spp=5)
    rb.sync()
    assert waves2 == waves and dev.cpu().numpy().tobytes() == fb.tobytes(), "device framebuffer / other batch split differs"


@pytest.mark.parametrize("arena,combine", [(0, 0), (0, 1), (1, 1)])
def test_device_whitted_options_keep_the_bits(world, battlefield, arena, combine):
    """Grow-only wave buffers (key 15) and per-warp radiance sums before the atomics (key 16) change scheduling only:
    same framebuffer bytes and wave sizes as the checker, over several frames of different sizes so the buffers grow,
    get reused while too large, and are handed over between calls."""
    scene, env, shading, images, sh = world
    rb.set_tuning(whitted_arena=arena, whitted_combine=combine)
    try:
        for width, height, spp, depth, seed, batch in [(64, 48, 1, 8, 3, 0), (256, 128, 3, 8, 11, 1), (200, 96, 2, 2, 3, 0),
                                                       (33, 17, 1, 1, 9, 0), (256, 128, 2, 8, 12, 0)]:
            cam = camera_for(battlefield, width, height)
            want, want_waves = oracle.whitted_trace(images, sh, cam, width, height, spp, depth, seed)
            got, waves = rb.whitted_trace(scene, env, shading, cam, width, height, spp, depth, seed, batch_spp=batch)
            assert waves == [int(x) for x in want_waves], "rays traced per bounce differ"
            assert got.tobytes() == want.tobytes(), f"{width}x{height}x{spp} depth {depth}: framebuffer differs"
    finally:
        rb.set_tuning(whitted_arena=1, whitted_combine=0)
