"""Device-side wavefront path tracer (rayaccel_b200/csrc/pathtrace.cu) through the C-ABI
(racc_cuda_path_trace) against its checker oracle_path_trace -- bit for bit: framebuffer and rays per bounce -- and
against the image of the reference's UNMODIFIED PathTracingRenderer (tests/golden/ref_render_tiles.npz, made by
tests/golden/make_render_golden.py; Monte-Carlo tolerances, see tests/test_render_oracle.py). All tests need a GPU."""
import os

import numpy as np
import pytest

import oracle
import rayaccel_b200 as rb
from test_render_oracle import camera_for, check_against_reference_image, synthetic_shading_case

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gpu():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    torch.cuda.set_device(0)
    rb.init(0)
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def scene(gpu, battlefield):
    s = rb.create_scene(battlefield.vertices, battlefield.indices)
    yield s
    s.destroy()


@pytest.fixture(scope="module")
def env(gpu, battlefield):
    e = rb.create_environment(battlefield.environment)
    yield e
    e.destroy()


@pytest.fixture(scope="module")
def shading(gpu, battlefield):
    s = rb.create_shading(battlefield.normals, battlefield.triangle_normals, battlefield.materials)
    yield s
    s.destroy()


@pytest.fixture(scope="module")
def checker(scene, battlefield):
    nodes, pairs, remap = scene.download()  # the images the kernel walks
    return (oracle.SceneImages(nodes, pairs, remap, battlefield.environment),
            oracle.Shading(battlefield.indices, battlefield.normals, battlefield.triangle_normals, battlefield.materials))


@pytest.mark.parametrize("width,height,spp,depth,seed,batch", [
    (256, 128, 4, 3, 11, 0),   # the reference's battlefield settings (maxDepth 3)
    (256, 128, 4, 3, 11, 1),   # one sample per batch: same image
    (256, 128, 5, 3, 11, 2),   # ragged last batch
    (200, 96, 2, 8, 3, 0),     # deep paths, sizes that are no multiple of the CTA
    (64, 64, 3, 0, 0, 0),      # depth 0, pixel centres: the light probe seen directly
    (33, 17, 1, 1, 9, 0),
])
def test_device_path_trace_equals_oracle(scene, env, shading, checker, battlefield, width, height, spp, depth, seed, batch):
    images, sh = checker
    cam = camera_for(battlefield, width, height)
    want, want_waves = oracle.path_trace(images, sh, cam, width, height, spp, depth, seed)
    got, waves = rb.path_trace(scene, env, shading, cam, width, height, spp, depth, seed, batch_spp=batch)
    assert waves == [int(x) for x in want_waves], "rays traced per bounce differ"
    bad = np.flatnonzero((got.view(np.uint32) != want.view(np.uint32)).reshape(-1, 4).any(axis=1))
    assert bad.size == 0, f"{bad.size} of {width * height} pixels differ, first {bad[:5]}: {got.reshape(-1, 4)[bad[:3]]} vs {want.reshape(-1, 4)[bad[:3]]}"


def test_device_path_trace_equals_oracle_on_synthetic_scene(gpu):
    """Branches battlefield never takes (tests/test_render_oracle.py::synthetic_shading_case): coloured materials,
    eta >= 1, clamped material ids, shading normals far from the geometric ones, degenerate normals."""
    verts, indices, normals, tri_normals, tri_materials, materials, env_img, cam = synthetic_shading_case()
    scene = rb.create_scene(verts, indices)
    env = rb.create_environment(env_img)
    shading = rb.create_shading(normals, tri_normals, tri_materials, materials)
    nodes, pairs, remap = scene.download()
    images = oracle.SceneImages(nodes, pairs, remap, env_img)
    sh = oracle.Shading(indices, normals, tri_normals, tri_materials, materials)
    for spp, depth, seed in ((8, 6, 2), (3, 12, 0)):
        want, want_waves = oracle.path_trace(images, sh, cam, 96, 64, spp, depth, seed)
        got, waves = rb.path_trace(scene, env, shading, cam, 96, 64, spp, depth, seed)
        assert waves == [int(x) for x in want_waves]
        assert got.tobytes() == want.tobytes(), f"{(got.view(np.uint32) != want.view(np.uint32)).reshape(-1, 4).any(axis=1).sum()} pixels differ"
    shading.destroy(); env.destroy(); scene.destroy()


def test_device_framebuffer_accumulates_and_splits_by_sample(scene, env, shading, battlefield):
    """A device framebuffer is added to in place; samples 0..5 at once == samples 0..1 then 2..5 (how the ranks of a
    multi-GPU job split a frame, rayaccel_b200/sharding.py sample_range) == the host-framebuffer path."""
    w, h = 320, 192
    cam = camera_for(battlefield, w, h)
    whole, waves = rb.path_trace(scene, env, shading, cam, w, h, 6, 3, seed=21)
    fb = torch.zeros(h * w * 4, dtype=torch.float32, device="cuda")
    _, w0 = rb.path_trace(scene, env, shading, cam, w, h, 2, 3, seed=21, framebuffer_ptr=fb.data_ptr())
    _, w1 = rb.path_trace(scene, env, shading, cam, w, h, 4, 3, seed=21, framebuffer_ptr=fb.data_ptr(), sample_base=2)
    rb.sync()
    assert fb.cpu().numpy().tobytes() == whole.tobytes()
    assert [a + b for a, b in zip(w0, w1)] == waves
    again, _ = rb.path_trace(scene, env, shading, cam, w, h, 6, 3, seed=21)
    assert again.tobytes() == whole.tobytes(), "two runs of the same frame differ"
    # with the ray re-binning on (the order rays are visited in must not matter)
    rb.set_tuning(sort=1)
    try:
        sorted_fb, sorted_waves = rb.path_trace(scene, env, shading, cam, w, h, 6, 3, seed=21)
    finally:
        rb.set_tuning(sort=2)
    assert sorted_fb.tobytes() == whole.tobytes() and sorted_waves == waves


@pytest.fixture
def streamed():
    """The streamed form of racc_cuda_path_trace (tuning key 19, pathstream.cu) for the duration of a test."""
    rb.set_tuning(path_stream=1)
    yield
    rb.set_tuning(path_stream=0)


@pytest.mark.parametrize("width,height,spp,depth,seed,batch", [
    (256, 128, 4, 3, 11, 0),
    (256, 128, 5, 3, 11, 2),   # ragged last batch; epochs of the earlier launches in the flag words
    (200, 96, 2, 8, 3, 0),     # deep paths
    (64, 64, 3, 0, 0, 0),      # depth 0: no queue at all
    (33, 17, 1, 1, 9, 0),      # fewer paths than one SM's lanes
    (640, 360, 4, 3, 5, 0),    # 0.9 M paths: every CTA of the persistent grid takes part in the queue
])
def test_streamed_path_trace_equals_oracle(scene, env, shading, checker, battlefield, streamed, width, height, spp, depth, seed, batch):
    """One persistent kernel per batch that traces, shades and queues the paths' next rays itself: same framebuffer bits and
    the same rays per bounce as oracle.path_trace (and so as the wavefront form)."""
    images, sh = checker
    cam = camera_for(battlefield, width, height)
    want, want_waves = oracle.path_trace(images, sh, cam, width, height, spp, depth, seed)
    got, waves = rb.path_trace(scene, env, shading, cam, width, height, spp, depth, seed, batch_spp=batch)
    assert waves == [int(x) for x in want_waves], "rays traced per bounce differ"
    bad = np.flatnonzero((got.view(np.uint32) != want.view(np.uint32)).reshape(-1, 4).any(axis=1))
    assert bad.size == 0, f"{bad.size} of {width * height} pixels differ, first {bad[:5]}: {got.reshape(-1, 4)[bad[:3]]} vs {want.reshape(-1, 4)[bad[:3]]}"


def test_streamed_path_trace_synthetic_scene_and_stack_tops(gpu, streamed):
    verts, indices, normals, tri_normals, tri_materials, materials, env_img, cam = synthetic_shading_case()
    scene = rb.create_scene(verts, indices)
    env = rb.create_environment(env_img)
    shading = rb.create_shading(normals, tri_normals, tri_materials, materials)
    nodes, pairs, remap = scene.download()
    images = oracle.SceneImages(nodes, pairs, remap, env_img)
    sh = oracle.Shading(indices, normals, tri_normals, tri_materials, materials)
    try:
        for spp, depth, seed, smem_stack in ((8, 6, 2, -1), (3, 12, 0, 16)):
            rb.set_tuning(smem_stack=smem_stack)
            want, want_waves = oracle.path_trace(images, sh, cam, 96, 64, spp, depth, seed)
            got, waves = rb.path_trace(scene, env, shading, cam, 96, 64, spp, depth, seed)
            assert waves == [int(x) for x in want_waves]
            assert got.tobytes() == want.tobytes(), f"{(got.view(np.uint32) != want.view(np.uint32)).reshape(-1, 4).any(axis=1).sum()} pixels differ"
    finally:
        rb.set_tuning(smem_stack=-1)
        shading.destroy(); env.destroy(); scene.destroy()


def test_streamed_full_frame_equals_wavefront_form(scene, env, shading, battlefield):
    """BASELINE.json's 1920x1080 x 4 spp frame, and 16 spp in one batch (33 M paths): the streamed form's framebuffer and
    rays per bounce are the wavefront form's, bit for bit; device framebuffers accumulate the same way."""
    w, h = 1920, 1080
    cam = camera_for(battlefield, w, h)
    for spp in (4, 16):
        fb, waves = rb.path_trace(scene, env, shading, cam, w, h, spp, 3, seed=1)
        rb.set_tuning(path_stream=1)
        try:
            got, waves2 = rb.path_trace(scene, env, shading, cam, w, h, spp, 3, seed=1)
            dev = torch.zeros(h * w * 4, dtype=torch.float32, device="cuda")
            rb.path_trace(scene, env, shading, cam, w, h, spp // 2, 3, seed=1, framebuffer_ptr=dev.data_ptr())
            rb.path_trace(scene, env, shading, cam, w, h, spp - spp // 2, 3, seed=1, framebuffer_ptr=dev.data_ptr(), sample_base=spp // 2)
            rb.sync()
        finally:
            rb.set_tuning(path_stream=0)
        assert waves2 == waves
        assert got.tobytes() == fb.tobytes()
        assert dev.cpu().numpy().tobytes() == fb.tobytes()


def test_device_path_trace_matches_reference_renderer_image(scene, env, shading, battlefield):
    g = np.load(os.path.join(GOLDEN, "ref_render_tiles.npz"))
    w, h, tile = int(g["width"]), int(g["height"]), int(g["tile"])
    spp = 64
    fb, waves = rb.path_trace(scene, env, shading, camera_for(battlefield, w, h), w, h, spp, int(g["max_depth"]), seed=11)
    check_against_reference_image(fb, spp, g["tiles"], tile)
    assert abs(sum(waves) / spp / float(g["rays_per_frame"]) - 1.0) < 0.01


def test_full_frame_properties(scene, env, shading, battlefield):
    """BASELINE.json's 1920x1080 x 4 spp frame, too large for the oracle in a test: every primary ray is traced, waves
    shrink, the image is finite and non-negative, its mean agrees with the small reference image (same camera
    framing), and a second run reproduces it bit for bit."""
    w, h, spp = 1920, 1080, 4
    cam = camera_for(battlefield, w, h)
    fb, waves = rb.path_trace(scene, env, shading, cam, w, h, spp, 3, seed=1)
    assert waves[0] == w * h * spp and all(waves[k] > waves[k + 1] > 0 for k in range(3))
    assert np.isfinite(fb).all() and (fb >= 0).all() and (fb[..., 3] == 0).all()
    again, waves2 = rb.path_trace(scene, env, shading, cam, w, h, spp, 3, seed=1, batch_spp=1)
    assert again.tobytes() == fb.tobytes() and waves2 == waves
    mean = float(fb[..., :3].astype(np.float64).mean() / spp)
    assert 0.3 < mean < 1.0


def test_path_trace_argument_errors(scene, env, shading, battlefield, checker):
    cam = camera_for(battlefield, 64, 64)
    other = rb.create_shading(battlefield.normals, battlefield.triangle_normals[:100], battlefield.materials[:100])
    with pytest.raises(rb.EngineError, match="does not match"):
        rb.path_trace(scene, env, other, cam, 64, 64, 1, 3, 1)
    other.destroy()
    images, _ = checker
    from_images = rb.create_scene_from_images(images.nodes, images.pairs, images.remap)
    with pytest.raises(rb.EngineError, match="created from images"):
        rb.path_trace(from_images, env, shading, cam, 64, 64, 1, 3, 1)
    from_images.destroy()
    with pytest.raises(rb.EngineError, match="max_depth"):
        rb.path_trace(scene, env, shading, cam, 64, 64, 1, 63, 1)
    # no light probe: nothing can contribute
    fb, waves = rb.path_trace(scene, None, shading, cam, 64, 64, 2, 3, 1)
    assert not fb.any() and waves[0] == 64 * 64 * 2
    # nothing to do
    fb, waves = rb.path_trace(scene, env, shading, cam, 64, 64, 0, 3, 1)
    assert not fb.any() and sum(waves) == 0
