"""Parity tests proper: the CUDA path, called through the C-ABI, against the CPU oracle on the same
bytes. Bar: bit-exact triangle ids AND bit-exact t/u/v (hits) and r/g/b (misses) -- the kernel and
the oracle use the same pinned fp32 operation sequence -- which implies the north_star's
"|dt|/t <= 1e-4". All tests here need a GPU."""
import os

import numpy as np
import pytest

import oracle
import rayaccel_b200 as rb
from conftest import assert_tolerance_parity, make_rays, random_rays

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

VARIANTS = [
    dict(variant=0, block=256, smem_nodes=-1, ctas_per_sm=0),
    dict(variant=0, block=256, smem_nodes=0, fetch_threshold=12),
    dict(variant=2, block=256, smem_nodes=-1, ctas_per_sm=0),
    dict(variant=2, inner_bail=32, leaf_bail=32, fetch_threshold=1),
    dict(variant=2, inner_bail=0, leaf_bail=0, block=512, ctas_per_sm=0),
    dict(variant=2, inner_bail=20, leaf_bail=1, block=128, ctas_per_sm=0, smem_nodes=100),
    dict(variant=0, block=1024, smem_nodes=-1),
    dict(variant=0, block=128, smem_nodes=64, fetch_threshold=1),
    dict(variant=0, block=512, smem_nodes=-1, fetch_threshold=32),
    dict(variant=1),
    dict(variant=2),
    dict(variant=3, ctas_per_sm=4, fetch_threshold=1, inner_bail=32, leaf_bail=32),
    dict(variant=3, block=128, ctas_per_sm=10, inner_bail=0, leaf_bail=0),
    dict(variant=3, block=512, ctas_per_sm=2, inner_bail=20, leaf_bail=1, fetch_threshold=32),
    dict(variant=3, smem_stack=16),
    dict(variant=3, smem_stack=8, fetch_threshold=4),
    dict(variant=3, sort=1),
    dict(variant=3, sort=1, sort_origin_bits=10, sort_dir_bits=0),
    dict(variant=3, sort=1, sort_origin_bits=0, sort_dir_bits=4, sort_dir_major=1),
    dict(variant=3, sort=1, sort_origin_bits=6, sort_dir_bits=4, sort_dir_major=1, fetch_threshold=8),
]
DEFAULT = dict(variant=3, block=256, ctas_per_sm=5, smem_nodes=0, fetch_threshold=16, inner_bail=8, leaf_bail=4,
               sort=0, sort_origin_bits=5, sort_dir_bits=0, sort_dir_major=0, smem_stack=0)


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
def images(scene, battlefield):
    nodes, pairs, remap = scene.download()  # what the kernel sees
    return oracle.SceneImages(nodes, pairs, remap, battlefield.environment)


def to_device(rays):
    return torch.from_numpy(rays.view(np.float32).reshape(-1).copy()).cuda()


def trace_dev(scene, env, rays_np):
    d_rays = to_device(rays_np)
    n = rays_np.shape[0]
    d_res = torch.full((max(n, 1) * 4,), 7.0, dtype=torch.float32, device="cuda")
    rb.trace_device(scene, env, [(d_rays.data_ptr(), d_res.data_ptr(), n)])
    torch.cuda.synchronize()
    return d_res[: n * 4].cpu().numpy().view(np.uint32).reshape(-1, 4)


def assert_bit_exact(got_u32, want_struct, what):
    want = want_struct.view(np.uint32).reshape(-1, 4)
    if np.array_equal(got_u32, want):
        return
    bad = np.nonzero((got_u32 != want).any(axis=1))[0]
    ids = (got_u32[bad, 0] != want[bad, 0]).sum()
    raise AssertionError(f"{what}: {bad.size}/{want.shape[0]} results differ ({ids} in the triangle id); first at ray {bad[0]}: "
                         f"got {got_u32[bad[0]]} want {want[bad[0]]}")


def device_primary(battlefield, width, height, spp=1, seed=0):
    cam = rb.Camera.for_scene(battlefield, width, height)
    n = width * height * spp
    d = torch.empty(n * 8, dtype=torch.float32, device="cuda")
    rb.generate_primary(cam, width, height, spp, seed, d.data_ptr())
    torch.cuda.synchronize()
    return d


def device_bounce(scene, d_rays, d_res, n, seed):
    out = torch.empty(max(n, 1) * 8, dtype=torch.float32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    rb.generate_bounce(scene, d_rays.data_ptr(), d_res.data_ptr(), n, seed, out.data_ptr(), cnt.data_ptr())
    torch.cuda.synchronize()
    m = int(cnt.item())
    return out[: m * 8], m


def test_scene_on_device_matches_host_build(scene, battlefield_images):
    nodes, pairs, remap = scene.download()
    assert np.array_equal(nodes.view(np.uint32), battlefield_images.nodes.view(np.uint32))
    assert np.array_equal(pairs.view(np.uint32), battlefield_images.pairs.view(np.uint32))
    assert np.array_equal(remap, battlefield_images.remap)


@pytest.mark.parametrize("tuning", VARIANTS, ids=lambda t: "-".join(f"{k}{v}" for k, v in t.items()))
def test_primary_and_bounces_bit_exact(tuning, scene, env, images, battlefield):
    """C2/C3 shaped input at reduced size: 480x270 primaries, then 3 diffuse bounces."""
    rb.set_tuning(**{**DEFAULT, **tuning})
    try:
        w, h = 480, 270
        d_rays = device_primary(battlefield, w, h, 1, seed=1)
        n = w * h
        for bounce in range(4):
            rays_np = d_rays[: n * 8].cpu().numpy().view(oracle.RAY_DTYPE)
            d_res = torch.empty(n * 4, dtype=torch.float32, device="cuda")
            rb.trace_device(scene, env, [(d_rays.data_ptr(), d_res.data_ptr(), n)])
            torch.cuda.synchronize()
            got = d_res.cpu().numpy().view(np.uint32).reshape(-1, 4)
            assert_bit_exact(got, oracle.traverse(images, rays_np), f"bounce {bounce}")
            d_rays, n = device_bounce(scene, d_rays, d_res, n, seed=2 + bounce)
            assert n > 0
    finally:
        rb.set_tuning(**DEFAULT)


def test_random_rays_bit_exact_and_brute_force(scene, env, images, battlefield):
    lo = battlefield.vertices[:, :3].min(0)
    hi = battlefield.vertices[:, :3].max(0)
    rays = random_rays(300_000, lo, hi, seed=8)
    got = trace_dev(scene, env, rays)
    want = oracle.traverse(images, rays)
    assert_bit_exact(got, want, "uniform random rays")
    # independent arbiter on a subset: id in the fp64 tie set, t within 1e-4 relative
    m = 3000
    t64, id64 = oracle.brute_f64(battlefield.vertices, battlefield.indices, rays[:m])
    ids = got[:m, 0]
    hit = ids != rb.INVALID_TRIANGLE
    assert np.array_equal(hit, np.isfinite(t64))
    t_gpu = got[:m, 1].view(np.float32)
    assert np.all(np.abs(t_gpu[hit] - t64[hit]) <= 1e-4 * t64[hit])
    t_of_id = oracle.tri_t_f64(battlefield.vertices, battlefield.indices, rays[:m], np.where(hit, ids, rb.INVALID_TRIANGLE))
    assert np.all(t_of_id[hit] <= t64[hit] * (1 + 1e-6) + 1e-9)


def test_counters_match_oracle(scene, env, images, battlefield):
    lo = battlefield.vertices[:, :3].min(0)
    hi = battlefield.vertices[:, :3].max(0)
    rays = random_rays(100_000, lo, hi, seed=3)
    _, cnt = oracle.traverse(images, rays, counters=True)
    for variant in (0, 1, 2, 3):
        rb.set_tuning(**{**DEFAULT, "variant": variant})
        d_rays = to_device(rays)
        d_res = torch.empty(rays.shape[0] * 4, dtype=torch.float32, device="cuda")
        d_cnt = torch.zeros(8, dtype=torch.int64, device="cuda")
        rb.trace_device(scene, env, [(d_rays.data_ptr(), d_res.data_ptr(), rays.shape[0])], counters_ptr=d_cnt.data_ptr())
        torch.cuda.synchronize()
        c = d_cnt.cpu().numpy()
        assert c[0] == rays.shape[0]
        assert c[1] == int(cnt["hit"].sum())
        assert c[2] == int(cnt["inner"].astype(np.int64).sum())
        assert c[3] == int(cnt["pairs"].astype(np.int64).sum())
        if variant == 3:
            assert c[4] == int(cnt["pushes"].astype(np.int64).sum()) and c[5] == int(cnt["leaves"].astype(np.int64).sum())
    rb.set_tuning(**DEFAULT)


def test_host_stream_path(scene, env, images, battlefield):
    """The reference-facing call: host buffers in, host results out (H2D + trace + D2H)."""
    lo = battlefield.vertices[:, :3].min(0)
    hi = battlefield.vertices[:, :3].max(0)
    rays = random_rays(50_000, lo, hi, seed=4)
    res = rb.trace_host(scene, env, rays)
    assert_bit_exact(res.view(np.uint32).reshape(-1, 4), oracle.traverse(images, rays), "host stream")


def test_many_streams_one_launch_ragged(scene, env, images, battlefield):
    """Several ray streams of ragged sizes (including empty ones) in a single launch."""
    lo = battlefield.vertices[:, :3].min(0)
    hi = battlefield.vertices[:, :3].max(0)
    sizes = [1, 0, 31, 32, 33, 11264, 0, 27648, 7, 1000]
    rays = [random_rays(s, lo, hi, seed=10 + k) for k, s in enumerate(sizes)]
    d_rays = [to_device(r) if r.shape[0] else torch.empty(8, device="cuda") for r in rays]
    d_res = [torch.full((max(s, 1) * 4,), 7.0, dtype=torch.float32, device="cuda") for s in sizes]
    rb.trace_device(scene, env, [(a.data_ptr(), b.data_ptr(), s) for a, b, s in zip(d_rays, d_res, sizes)])
    torch.cuda.synchronize()
    for k, s in enumerate(sizes):
        if s:
            got = d_res[k][: s * 4].cpu().numpy().view(np.uint32).reshape(-1, 4)
            assert_bit_exact(got, oracle.traverse(images, rays[k]), f"stream {k}")
        else:
            assert float(d_res[k][0]) == 7.0  # untouched


def test_empty_launch_is_a_noop(scene, env):
    rb.trace_device(scene, env, [])
    rb.trace_device(scene, env, [(0, 0, 0)])
    torch.cuda.synchronize()


def test_no_environment_misses_return_zero(scene, images, battlefield):
    rays = make_rays([[0, 500, 0]] * 64, [[0, 1, 0]] * 64)
    got = trace_dev(scene, None, rays)
    assert np.all(got[:, 0] == rb.INVALID_TRIANGLE)
    assert np.all(got[:, 1:] == 0)


def test_reference_built_images_give_identical_hits(gpu, env, battlefield):
    """Trace the UNMODIFIED reference builder's own images with our kernel: results must equal the
    ones from our build (the images are structurally identical, only the numbering differs)."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    ref_img = oracle.ref_build_scene(battlefield.vertices, battlefield.indices)
    s_ref = rb.create_scene_from_images(ref_img.nodes, ref_img.pairs, ref_img.remap)
    s_own = rb.create_scene(battlefield.vertices, battlefield.indices)
    lo = battlefield.vertices[:, :3].min(0)
    hi = battlefield.vertices[:, :3].max(0)
    rays = random_rays(200_000, lo, hi, seed=5)
    a = trace_dev(s_ref, env, rays)
    b = trace_dev(s_own, env, rays)
    assert np.array_equal(a, b)
    ref_img.env = battlefield.environment
    assert_bit_exact(a, oracle.traverse(ref_img, rays), "reference-built images")
    s_ref.destroy()
    s_own.destroy()


def test_full_size_properties(scene, env, battlefield):
    """BASELINE config 2 at full size (1920x1080x4 spp = 8.3 M rays): size-independent properties
    instead of an oracle run -- idempotence (two launch shapes, same bits), every hit id in range,
    t inside (minT, maxT], barycentrics in the triangle, and a shortened ray (maxT just past t)
    re-hits the same triangle at the same t."""
    w, h, spp = 1920, 1080, 4
    n = w * h * spp
    d_rays = device_primary(battlefield, w, h, spp, seed=1)
    d_a = torch.empty(n * 4, dtype=torch.float32, device="cuda")
    d_b = torch.empty(n * 4, dtype=torch.float32, device="cuda")
    rb.set_tuning(**DEFAULT)
    rb.trace_device(scene, env, [(d_rays.data_ptr(), d_a.data_ptr(), n)])
    rb.set_tuning(**{**DEFAULT, "variant": 1})
    rb.trace_device(scene, env, [(d_rays.data_ptr(), d_b.data_ptr(), n)])
    rb.set_tuning(**DEFAULT)
    torch.cuda.synchronize()
    assert torch.equal(d_a.view(torch.int32), d_b.view(torch.int32))
    # re-binned visiting order (raysort.cu) and the reference-format bail-out kernel: same bits again
    for other in (dict(sort=1, sort_origin_bits=6, sort_dir_bits=4), dict(variant=2)):
        d_b.zero_()
        rb.set_tuning(**{**DEFAULT, **other})
        rb.trace_device(scene, env, [(d_rays.data_ptr(), d_b.data_ptr(), n)])
        rb.set_tuning(**DEFAULT)
        torch.cuda.synchronize()
        assert torch.equal(d_a.view(torch.int32), d_b.view(torch.int32)), other
    res = d_a.view(-1, 4)
    ids = res[:, 0].view(torch.int32)
    hit = ids != -1
    assert 0.5 < float(hit.float().mean()) < 0.95
    assert int(ids[hit].max()) < battlefield.triangle_count and int(ids[hit].min()) >= 0
    t, u, v = res[hit, 1], res[hit, 2], res[hit, 3]
    assert bool((t > 0).all()) and bool((t <= 1e6).all())
    assert bool((u >= -1e-6).all()) and bool((v >= -1e-6).all()) and bool((u + v <= 1 + 1e-5).all())
    # shorten every hit ray to just past its hit: the closest hit must not change
    rays2 = d_rays.view(-1, 8).clone()
    rays2[hit, 7] = t * (1.0 + 1e-5)
    d_c = torch.empty(n * 4, dtype=torch.float32, device="cuda")
    rb.trace_device(scene, env, [(rays2.data_ptr(), d_c.data_ptr(), n)])
    torch.cuda.synchronize()
    res2 = d_c.view(-1, 4)
    same_id = res2[:, 0].view(torch.int32) == ids
    frac = float(same_id.float().mean())
    assert frac > 0.9999, frac  # a different id is only legal when two triangles tie at exactly t
    assert bool((res2[~hit, 0].view(torch.int32) == -1).all())
    both = same_id & hit
    assert torch.equal(res2[both, 1], res[both, 1])


def test_kat_scenes_on_gpu(gpu):
    """Every hand-built tie / boundary case of tests/kat_scenes.py, on the CUDA path: equal to the
    oracle bit for bit, and to the analytic answers."""
    from kat_scenes import KAT_CASES, build_kat_scene
    for case in KAT_CASES:
        images, rays = build_kat_scene(case)
        pad = (-images.pairs.shape[0] * 3) % 32 // 3 + (1 if (images.pairs.shape[0] * 3) % 32 == 0 else 0)
        pairs = np.concatenate([images.pairs, np.repeat(images.pairs[:1], max(pad, 1), axis=0)])
        s = rb.create_scene_from_images(images.nodes, pairs, images.remap)
        for variant in (0, 1, 2, 3):
            rb.set_tuning(**{**DEFAULT, "variant": variant})
            got = trace_dev(s, None, rays)
            assert_bit_exact(got, oracle.traverse(images, rays), f"{case['name']} variant {variant}")
            for k, want in enumerate(case["expect"]):
                if want is None:
                    assert got[k, 0] == rb.INVALID_TRIANGLE, (case["name"], k)
                else:
                    assert got[k, 0] == want["triangle"], (case["name"], k)
                    assert abs(float(got[k, 1:2].view(np.float32)[0]) - want["t"]) <= 1e-4 * want["t"], (case["name"], k)
        rb.set_tuning(**DEFAULT)
        s.destroy()


def test_golden_rays_on_gpu(scene, env):
    """The committed golden vectors (tests/golden/battlefield_rays.npz: oracle results on the
    REFERENCE-built images + fp64 arbiter): bit-exact ids and t/u/v/r/g/b on the CUDA path."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "battlefield_rays.npz"))
    rays = np.ascontiguousarray(g["rays"]).view(oracle.RAY_DTYPE).reshape(-1)
    got = trace_dev(scene, env, rays)
    assert np.array_equal(got, g["results"])
    both = (got[:, 0] != rb.INVALID_TRIANGLE) & np.isfinite(g["t64"])
    t = got[:, 1].copy().view(np.float32)
    assert np.all(np.abs(t[both] - g["t64"][both]) <= 1e-4 * g["t64"][both])  # north_star: t within 1e-4 relative


def test_engine_matches_reference_cpu_path(scene, env, battlefield):
    """north_star's parity statement on the CUDA path: results equal the reference's CPU query path on the same rays --
    executeRayQueryCPU (Scene.cpp:374-484) run from its own source over the stand-in for its binary-only Embree
    (oracle/ref_shim/mini_embree.cpp). Committed golden vectors always; when oracle/_ref travelled, also live on 1 M rays
    of the full-size bench streams (every 8th primary ray of 1920x1080x4spp... and its first bounce)."""
    from conftest import assert_matches_reference_cpu_path
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "battlefield_rays.npz"))
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_cpu_path.npz"))["results"]
    rays = np.ascontiguousarray(g["rays"]).view(oracle.RAY_DTYPE).reshape(-1)
    assert_matches_reference_cpu_path(trace_dev(scene, env, rays), ref, rays, battlefield.vertices, battlefield.indices, "golden rays")
    if not oracle.have_ref_cpu_query():
        return
    w, h, spp = 1920, 1080, 4
    n = w * h * spp
    d_rays = device_primary(battlefield, w, h, spp, seed=1)
    d_res = torch.empty(n * 4, dtype=torch.float32, device="cuda")
    rb.trace_device(scene, env, [(d_rays.data_ptr(), d_res.data_ptr(), n)])
    torch.cuda.synchronize()
    d_b, nb = device_bounce(scene, d_rays, d_res, n, seed=2)
    d_bres = torch.empty(max(nb, 1) * 4, dtype=torch.float32, device="cuda")
    rb.trace_device(scene, env, [(d_b.data_ptr(), d_bres.data_ptr(), nb)])
    torch.cuda.synchronize()
    for name, dr, do, count, stride in (("primary", d_rays, d_res, n, 13), ("bounce", d_b, d_bres, nb, 11)):
        sub = dr.view(-1, 8)[:count][::stride].contiguous().cpu().numpy().reshape(-1).view(oracle.RAY_DTYPE)
        got = do.view(-1, 4)[:count][::stride].contiguous().cpu().numpy()
        want = oracle.ref_cpu_query(battlefield.vertices, battlefield.indices, battlefield.environment, sub)
        differ, rel = assert_matches_reference_cpu_path(got, want, sub, battlefield.vertices, battlefield.indices, name)
        print(f"engine vs reference CPU path, {name}: {len(sub)} rays, {differ} primIDs differ (ties), max |dt|/t {rel:.2e}")


def test_host_streams_packed_into_shared_launches(scene, env, images, battlefield):
    """Many small HOST streams (the sizes racc::render() submits) in one call: the engine packs them
    into shared staging chunks; every stream's results must land in its own buffer, bit-exact."""
    lo = battlefield.vertices[:, :3].min(0)
    hi = battlefield.vertices[:, :3].max(0)
    sizes = [11264, 27648, 1, 65535, 300, 49152, 7, 49152, 49152, 1000, 600000, 13]
    rays = [random_rays(s, lo, hi, seed=40 + k) for k, s in enumerate(sizes)]
    pinned_r = [torch.from_numpy(r.view(np.float32).reshape(-1).copy()).pin_memory() for r in rays]
    pinned_o = [torch.full((s * 4,), 7.0, dtype=torch.float32).pin_memory() for s in sizes]
    before = rb.launch_count()
    rb.trace_host_ptrs(scene, env, [(a.data_ptr(), b.data_ptr(), s) for a, b, s in zip(pinned_r, pinned_o, sizes)])
    rb.sync()
    launches = rb.launch_count() - before
    assert launches <= 3, f"{launches} launches for {sum(sizes)} rays: small streams were not packed"
    for k, s in enumerate(sizes):
        got = pinned_o[k].numpy().view(np.uint32).reshape(-1, 4)
        assert_bit_exact(got, oracle.traverse(images, rays[k]), f"host stream {k}")


def test_host_streams_zero_copy(scene, env, images, battlefield):
    """HOST streams in pinned memory read and written by the kernel itself (host_zero_copy=1): one launch, same bits;
    pageable host memory still goes through the staging pipeline."""
    lo = battlefield.vertices[:, :3].min(0)
    hi = battlefield.vertices[:, :3].max(0)
    sizes = [27648, 1, 65535, 300000]
    rays = [random_rays(s, lo, hi, seed=60 + k) for k, s in enumerate(sizes)]
    pinned_r = [torch.from_numpy(r.view(np.float32).reshape(-1).copy()).pin_memory() for r in rays]
    pinned_o = [torch.full((s * 4,), 7.0, dtype=torch.float32).pin_memory() for s in sizes]
    rb.set_tuning(host_zero_copy=1)
    try:
        before = rb.launch_count()
        rb.trace_host_ptrs(scene, env, [(a.data_ptr(), b.data_ptr(), s) for a, b, s in zip(pinned_r, pinned_o, sizes)])
        rb.sync()
        assert rb.launch_count() - before == 1
        for k, s in enumerate(sizes):
            assert_bit_exact(pinned_o[k].numpy().view(np.uint32).reshape(-1, 4), oracle.traverse(images, rays[k]), f"zero-copy stream {k}")
        pageable = random_rays(5000, lo, hi, seed=70)
        res = rb.trace_host(scene, env, pageable)
        assert_bit_exact(res.view(np.uint32).reshape(-1, 4), oracle.traverse(images, pageable), "pageable host stream")
    finally:
        rb.set_tuning(host_zero_copy=0)


def test_synthetic_soup_scene_bit_exact(gpu):
    """BASELINE.json configs[4] at reduced size: random triangle soup + uniform random rays."""
    v, i = rb.synthetic_triangles(200_000, seed=7, extent=1000.0, edge=2.0)
    s = rb.create_scene(v, i)
    nodes, pairs, remap = s.download()
    img = oracle.SceneImages(nodes, pairs, remap)
    rays = random_rays(400_000, np.zeros(3), np.full(3, 1000.0), seed=8)
    got = trace_dev(s, None, rays)
    assert_bit_exact(got, oracle.traverse(img, rays), "soup")
    assert 0.001 < (got[:, 0] != rb.INVALID_TRIANGLE).mean() < 0.9
    s.destroy()


@pytest.mark.parametrize("tuning", VARIANTS, ids=lambda t: "-".join(f"{k}{v}" for k, v in t.items()))
def test_deep_stack_scene_bit_exact(tuning, gpu):
    """tests/kat_scenes.py::deep_stack_scene on every launch shape: traversal stacks of up to 60 entries (Kernels.h:166
    allows 64), lanes of one warp at different depths. For `smem_stack` 16 / 8 everything past the shared-memory entries
    goes through HybridStack's spill branch (traverse_packed.cu) -- no battlefield ray gets there (deepest stack 12)."""
    from kat_scenes import deep_stack_scene
    images, rays, expect = deep_stack_scene(n_rays=40_000)
    s = rb.create_scene_from_images(images.nodes, images.pairs, images.remap)
    rb.set_tuning(**{**DEFAULT, **tuning})
    try:
        got = trace_dev(s, None, rays)
        want, cnt = oracle.traverse(images, rays, counters=True)
        assert int(cnt["max_stack"].max()) == 60
        assert_bit_exact(got, want, "deep stack")
        sure = expect != 0xFFFFFFFE
        assert np.array_equal(got[sure, 0], expect[sure])
    finally:
        rb.set_tuning(**DEFAULT)
        s.destroy()


def test_scene_larger_than_l2_auto_path_bit_exact(gpu):
    """BASELINE.json configs[4] at a size where the AUTO path engages (3 M triangles: node + pair images 279 MB > 2 x L2):
    rays re-binned by origin (raysort.cu) and stack tops in shared memory -- what the config-5 numbers are measured with --
    against the oracle on the downloaded images, 1 M uniform random rays, every bit. Then the same rays with each of the
    two features alone and with neither."""
    v, i = rb.synthetic_triangles(3_000_000, seed=7, extent=1000.0, edge=2.0)
    s = rb.create_scene(v, i)
    assert (s.info["node_count"] + s.info["pair_count"]) * 64 > 256 << 20, "the scene no longer exceeds the auto threshold"
    assert s.info["depth"] > 16
    nodes, pairs, remap = s.download()
    img = oracle.SceneImages(nodes, pairs, remap)
    rays = random_rays(1_000_000, np.zeros(3), np.full(3, 1000.0), seed=8)
    want, cnt = oracle.traverse(img, rays, counters=True)
    assert 0.3 < (want["triangle"] != oracle.INVALID).mean() < 0.95
    try:
        for tuning in (dict(sort=2, smem_stack=-1), dict(sort=1, smem_stack=0), dict(sort=0, smem_stack=16), dict(sort=0, smem_stack=0),
                       dict(sort=1, smem_stack=8, sort_origin_bits=8, sort_dir_bits=2)):
            rb.set_tuning(**{**DEFAULT, **tuning})
            d_rays = to_device(rays)
            d_res = torch.full((rays.shape[0] * 4,), 7.0, dtype=torch.float32, device="cuda")
            d_cnt = torch.zeros(8, dtype=torch.int64, device="cuda")
            rb.trace_device(s, None, [(d_rays.data_ptr(), d_res.data_ptr(), rays.shape[0])], counters_ptr=d_cnt.data_ptr())
            torch.cuda.synchronize()
            assert_bit_exact(d_res.cpu().numpy().view(np.uint32).reshape(-1, 4), want, f"3 M-triangle soup {tuning}")
            c = d_cnt.cpu().numpy()
            assert c[2] == int(cnt["inner"].astype(np.int64).sum()) and c[3] == int(cnt["pairs"].astype(np.int64).sum())
    finally:
        rb.set_tuning(**{**DEFAULT, "sort": 2, "smem_stack": -1})
        s.destroy()


# ---------------------------------------------------------------------------------------------
# variant 4: 32-byte quantised nodes. NOT bit-exact by design: north_star's bar, arbitrated by fp64.

@pytest.mark.parametrize("tuning", [dict(variant=4), dict(variant=4, smem_stack=16), dict(variant=4, sort=1, smem_stack=16)],
                         ids=["quantised", "quantised_smem16", "quantised_rebinned_smem16"])
def test_quantised_nodes_battlefield_tolerance_parity(tuning, scene, env, images, battlefield):
    """480x270 primaries + 3 bounces through the quantised-node kernel: every result equal to the oracle's except rays whose
    id the fp64 brute force confirms as a tie; same id => same bits (the pair test is the exact one)."""
    rb.set_tuning(**{**DEFAULT, **tuning})
    try:
        w, h = 480, 270
        d_rays = device_primary(battlefield, w, h, 1, seed=1)
        n = w * h
        differing = 0
        for bounce in range(4):
            rays_np = d_rays[: n * 8].cpu().numpy().view(oracle.RAY_DTYPE)
            d_res = torch.empty(n * 4, dtype=torch.float32, device="cuda")
            rb.trace_device(scene, env, [(d_rays.data_ptr(), d_res.data_ptr(), n)])
            torch.cuda.synchronize()
            got = d_res.cpu().numpy().view(np.uint32).reshape(-1, 4)
            differing += assert_tolerance_parity(got, oracle.traverse(images, rays_np), rays_np, battlefield.vertices, battlefield.indices, f"bounce {bounce}")
            d_rays, n = device_bounce(scene, d_rays, d_res, n, seed=2 + bounce)
        print(f"quantised nodes {tuning}: {differing} rays differ from the oracle over 4 waves")
    finally:
        rb.set_tuning(**DEFAULT)


def test_quantised_nodes_full_size_and_soup(scene, env, battlefield):
    """Variant 4 at size, against variant 3 on the same rays (which the tests above pin to the oracle): the full 8.3 M-ray
    primary stream of the bench and 1 M random rays in a 3 M-triangle soup (the config the format is meant for). Ids may
    differ only for a vanishing fraction of rays, and where they agree every bit agrees."""
    w, h, spp = 1920, 1080, 4
    n = w * h * spp
    d_rays = device_primary(battlefield, w, h, spp, seed=1)
    cases = [("battlefield primaries", scene, env, d_rays, n, battlefield.vertices, battlefield.indices)]
    v, i = rb.synthetic_triangles(3_000_000, seed=7, extent=1000.0, edge=2.0)
    soup = rb.create_scene(v, i)
    soup_rays = to_device(random_rays(1_000_000, np.zeros(3), np.full(3, 1000.0), seed=8))
    cases.append(("3 M-triangle soup", soup, None, soup_rays, 1_000_000, v, i))
    try:
        for name, sc, en, rays, count, verts, idx in cases:
            out = []
            for variant in (3, 4):
                rb.set_tuning(**{**DEFAULT, "variant": variant, "sort": 2, "smem_stack": -1})
                d_res = torch.empty(count * 4, dtype=torch.float32, device="cuda")
                rb.trace_device(sc, en, [(rays.data_ptr(), d_res.data_ptr(), count)])
                torch.cuda.synchronize()
                out.append(d_res.view(-1, 4).view(torch.int32))
            exact, quant = out
            same = (exact[:, 0] == quant[:, 0])
            differing = int((~same).sum())
            assert differing <= 2e-5 * count + 3, f"{name}: {differing} of {count} ids differ"
            assert torch.equal(exact[same], quant[same]), f"{name}: same triangle, different bits"
            if differing:
                bad = torch.nonzero(~same).flatten().cpu().numpy()
                sub = rays.view(-1, 8)[torch.from_numpy(bad).cuda()].cpu().numpy().reshape(-1).view(oracle.RAY_DTYPE)
                got = quant[torch.from_numpy(bad).cuda()].cpu().numpy().view(np.uint32)
                want = exact[torch.from_numpy(bad).cuda()].cpu().numpy().view(np.uint32)
                # route the differing rays through the fp64 arbiter (reusing the helper: `want` plays the oracle)
                assert_tolerance_parity(got, want.view(oracle.RESULT_DTYPE).reshape(-1), sub, verts, idx, name, max_id_mismatch=1.0)
            print(f"quantised vs exact, {name}: {differing} of {count} ids differ")
    finally:
        rb.set_tuning(**{**DEFAULT, "sort": 2, "smem_stack": -1})
        soup.destroy()


def test_quantised_nodes_keep_hits_at_the_scene_bounds(gpu):
    """kat_scenes.bound_vertex_case: rays through the vertices that are the scene's own bounds. Conservative boxes may find
    MORE than the exact ones, never less: no ray may miss where the checker hits (three did while cell 0 of the grid sat on
    the lower bound), and where the triangle agrees so do all the words."""
    from kat_scenes import bound_vertex_case
    verts, indices, rays = bound_vertex_case()
    scene = rb.create_scene(verts, indices)
    nodes, pairs, remap = scene.download()
    want = oracle.traverse(oracle.SceneImages(nodes, pairs, remap), rays).view(np.uint32).reshape(-1, 4)
    rb.set_tuning(variant=4)
    try:
        got = trace_dev(scene, None, rays)
    finally:
        rb.set_tuning(variant=3)
        scene.destroy()
    assert not ((got[:, 0] == 0xFFFFFFFF) & (want[:, 0] != 0xFFFFFFFF)).any(), "a quantised box lost a hit"
    same = got[:, 0] == want[:, 0]
    assert same.mean() > 0.99 and np.array_equal(got[same], want[same])


def test_quantised_nodes_hand_built_scenes(gpu):
    from kat_scenes import KAT_CASES, build_kat_scene, deep_stack_scene
    rb.set_tuning(**{**DEFAULT, "variant": 4})
    try:
        for case in KAT_CASES:
            if case["name"] == "coplanar_duplicates_later_pair_wins":
                continue  # an exact tie by construction: the one thing the format leaves open
            images, rays = build_kat_scene(case)
            pairs = np.concatenate([images.pairs, np.repeat(images.pairs[:1], (-images.pairs.shape[0]) % 32 or 32, axis=0)])
            s = rb.create_scene_from_images(images.nodes, pairs, images.remap)
            assert_bit_exact(trace_dev(s, None, rays), oracle.traverse(images, rays), case["name"])
            s.destroy()
        images, rays, _ = deep_stack_scene(n_rays=20_000)
        s = rb.create_scene_from_images(images.nodes, images.pairs, images.remap)
        assert_bit_exact(trace_dev(s, None, rays), oracle.traverse(images, rays), "deep stack, quantised nodes")
        s.destroy()
    finally:
        rb.set_tuning(**DEFAULT)


def test_eight_bounces_bit_exact(scene, env, images, battlefield):
    """BASELINE.json configs[2]'s depth at reduced size: 480x270 primaries followed by EIGHT diffuse bounces, every wave
    against the oracle (deep bounces are the short, incoherent streams; the waves shrink to a few thousand rays)."""
    w, h = 480, 270
    d_rays = device_primary(battlefield, w, h, 1, seed=3)
    n = w * h
    sizes = []
    for bounce in range(9):
        rays_np = d_rays[: n * 8].cpu().numpy().view(oracle.RAY_DTYPE)
        d_res = torch.empty(n * 4, dtype=torch.float32, device="cuda")
        rb.trace_device(scene, env, [(d_rays.data_ptr(), d_res.data_ptr(), n)])
        torch.cuda.synchronize()
        assert_bit_exact(d_res.cpu().numpy().view(np.uint32).reshape(-1, 4), oracle.traverse(images, rays_np), f"bounce {bounce}")
        sizes.append(n)
        d_rays, n = device_bounce(scene, d_rays, d_res, n, seed=20 + bounce)
        assert n > 0
    assert sizes[8] < sizes[0] // 50, sizes


def test_hundreds_of_tiny_host_streams_in_one_call(scene, env, images, battlefield):
    """A flush of many partially filled API streams: 700 HOST streams of 1..3000 rays (and a few empty ones) in ONE
    racc_cuda_trace call. They share staging chunks -- a chunk takes as many stream segments as fit, not 64 -- so the
    call stays a handful of launches, and every stream's results land in its own buffer, bit-exact."""
    lo = battlefield.vertices[:, :3].min(0)
    hi = battlefield.vertices[:, :3].max(0)
    rng = np.random.default_rng(123)
    sizes = [int(x) for x in rng.integers(1, 3000, size=700)]
    for k in (5, 77, 300):
        sizes[k] = 0
    total = sum(sizes)
    all_rays = random_rays(total, lo, hi, seed=321)
    want = oracle.traverse(images, all_rays)
    pin_r = torch.from_numpy(all_rays.view(np.float32).reshape(-1).copy()).pin_memory()
    pin_o = torch.full((total * 4 + 4,), 7.0, dtype=torch.float32).pin_memory()
    descs, off = [], 0
    for n in sizes:
        descs.append((pin_r.data_ptr() + off * 32, pin_o.data_ptr() + off * 16, n))
        off += n
    before = rb.launch_count()
    rb.trace_host_ptrs(scene, env, descs)
    rb.sync()
    launches = rb.launch_count() - before
    assert launches <= 3, f"{launches} launches for {total} rays in {len(sizes)} streams"
    got = pin_o.numpy()[: total * 4].view(np.uint32).reshape(-1, 4)
    assert_bit_exact(got, want, "tiny host streams")
    assert float(pin_o[total * 4]) == 7.0  # nothing written past the last stream


def test_shared_edge_zero_keeps_its_sign_on_gpu(gpu):
    """kat_scenes.shared_edge_mesh_case (the round on which tests/fuzz/fuzz_gpu.py caught the checker following gcc's fnmsub fold):
    rays through shared edges whose edge function cancels to an exact zero -- known answers and the checker's bits on all rays,
    every kernel family, both scene builders."""
    from kat_scenes import shared_edge_mesh_case
    verts, indices, rays, known = shared_edge_mesh_case()
    for build in (0, 2):
        rb.set_tuning(build_device=build)
        s = rb.create_scene(verts, indices)
        rb.set_tuning(build_device=3)
        nodes, pairs, remap = s.download()
        want = oracle.traverse(oracle.SceneImages(nodes, pairs, remap), rays)
        for tuning in (dict(), dict(variant=1), dict(variant=2), dict(variant=0), dict(smem_stack=16), dict(sort=1)):
            rb.set_tuning(**{**DEFAULT, **tuning})
            try:
                got = trace_dev(s, None, rays)
            finally:
                rb.set_tuning(**DEFAULT)
            for k, w in known.items():
                assert tuple(int(x) for x in got[k]) == w, (build, tuning, k, got[k], w)
            assert_bit_exact(got, want, f"shared-edge mesh, build {build}, {tuning}")
        s.destroy()


def test_trees_deeper_than_the_stack_are_refused(gpu):
    from test_library_on_cpu import deep_chain_images
    nodes, pairs, remap = deep_chain_images(70)
    with pytest.raises(rb.EngineError, match="levels deep"):
        rb.create_scene_from_images(nodes, pairs, remap)


def test_tiny_scenes_trace(gpu):
    """Scenes whose root stays a leaf: a cube and a single triangle now get a synthetic root and trace."""
    c = np.array([[x, y, z, 1] for x in (0, 1) for y in (0, 1) for z in (0, 1)], np.float32)
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    cube_i = np.array([[a, b, cc, a, cc, d] for a, b, cc, d in quads], np.uint32).ravel()
    tri_v = np.array([[0, 0, 0, 1], [1, 0, 0, 1], [0, 1, 0, 1]], np.float32)
    for v, i, name in ((c, cube_i, "cube"), (tri_v, np.arange(3, dtype=np.uint32), "one triangle")):
        for build in (0, 2):
            rb.set_tuning(build_device=build)
            try:
                s = rb.create_scene(v, i)
            finally:
                rb.set_tuning(build_device=3)
            nodes, pairs, remap = s.download()
            rays = random_rays(20_000, v[:, :3].min(0) - 1, v[:, :3].max(0) + 1, seed=9)
            got = trace_dev(s, None, rays)
            assert_bit_exact(got, oracle.traverse(oracle.SceneImages(nodes, pairs, remap), rays), name)
            assert (got[:, 0] != rb.INVALID_TRIANGLE).any()
            s.destroy()


# ---------------------------------------------------------------------------------------------
# device SAH build (bvh_build.cu): byte equality of the device images with the host build, which the
# CPU suite pins to the unmodified reference builder (tests/test_scene_build.py)

def _images_with(build_device, v, i):
    rb.set_tuning(build_device=build_device)
    try:
        s = rb.create_scene(v, i)
        nodes, pairs, remap = s.download()
        return nodes.view(np.uint32).copy(), pairs.view(np.uint32).copy(), remap.copy(), s.info
    finally:
        rb.set_tuning(build_device=3)


def _assert_same_images(v, i, what):
    """mode 1: SAH tree on the device, packing on the host; mode 2: everything on the device."""
    hn, hp, hr, hinfo = _images_with(0, v, i)
    for mode in (1, 2):
        dn, dp, dr, dinfo = _images_with(mode, v, i)
        assert hinfo == dinfo, f"{what} mode {mode}: scene info differs: {hinfo} vs {dinfo}"
        assert np.array_equal(hn, dn), f"{what} mode {mode}: node images differ"
        assert np.array_equal(hp, dp), f"{what} mode {mode}: pair images differ"
        assert np.array_equal(hr, dr), f"{what} mode {mode}: remap tables differ"


def test_device_build_battlefield_identical_to_host(battlefield):
    _assert_same_images(battlefield.vertices, battlefield.indices, "battlefield")


def test_device_build_synthetic_identical_to_host():
    import sys as _sys
    _sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden import synthetic_meshes
    for name, (v, i) in sorted(synthetic_meshes().items()):
        _assert_same_images(v, i, name)
    # soups of awkward sizes: below one 8-block, around the warp/CTA hand-over, forced median splits
    for n, seed, extent, edge in ((3, 1, 10.0, 1.0), (9, 2, 10.0, 1.0), (17, 3, 10.0, 3.0), (64, 4, 20.0, 2.0), (65, 5, 20.0, 2.0),
                                  (127, 6, 0.0, 1.0), (300, 7, 0.0, 1.0), (1000, 8, 30.0, 5.0), (4097, 9, 100.0, 3.0), (50000, 10, 200.0, 2.0)):
        v, i = rb.synthetic_triangles(n, seed=seed, extent=extent, edge=edge)
        _assert_same_images(v, i, f"soup n={n}")


def test_device_build_stress_meshes_identical_to_host():
    """Meshes that lean on the order-dependent parts of the sweep: a large regular grid (thousands of equal centre keys
    and equal costs: tie rules, wide kernel), thousands of coincident triangles (zero-extent parents: forced median
    splits down to 126-triangle leaves), huge coordinates with needle triangles, and degenerate (zero-area) triangles."""
    g = 300
    xs, zs = np.meshgrid(np.arange(g + 1, dtype=np.float32), np.arange(g + 1, dtype=np.float32))
    v = np.ones(((g + 1) * (g + 1), 4), np.float32)
    v[:, 0], v[:, 1], v[:, 2] = xs.ravel(), 0.0, zs.ravel()
    a = (np.arange(g)[:, None] * (g + 1) + np.arange(g)[None, :]).ravel().astype(np.uint32)
    idx = np.stack([a, a + g + 1, a + 1, a + 1, a + g + 2 - 1, a + g + 2], axis=1).ravel().astype(np.uint32)
    _assert_same_images(v, idx, "grid 300x300")

    one = np.array([[0, 0, 0, 1], [1, 0, 0, 1], [0, 1, 0, 1]], np.float32)
    n = 5000
    _assert_same_images(np.tile(one, (n, 1)), np.arange(3 * n, dtype=np.uint32), "5000 coincident triangles")

    rng = np.random.default_rng(99)
    n = 30000
    c = rng.uniform(-1e6, 1e6, size=(n, 1, 3))
    d = rng.normal(size=(n, 1, 3)) * 1e4
    t = np.concatenate([c, c + d, c + d * 1.0001 + rng.normal(size=(n, 1, 3)) * 1e-2], axis=1).astype(np.float32)
    v = np.ones((3 * n, 4), np.float32)
    v[:, :3] = t.reshape(-1, 3)
    _assert_same_images(v, np.arange(3 * n, dtype=np.uint32), "needles at 1e6")

    v, i = rb.synthetic_triangles(20000, seed=31, extent=50.0, edge=2.0)
    v = v.copy()
    v[3 * 5000 + 1: 3 * 9000: 3, :3] = v[3 * 5000: 3 * 9000: 3, :3]  # 4000 triangles with two equal vertices
    v[2::9, 1] = 0.0
    _assert_same_images(v, i, "soup with degenerate triangles")


@pytest.mark.slow
def test_device_build_large_soup_identical_to_host():
    v, i = rb.synthetic_triangles(1_000_000, seed=7, extent=1000.0, edge=2.0)
    _assert_same_images(v, i, "soup n=1e6")


def test_device_built_scene_traces_bit_exact(battlefield, env, images):
    """A scene whose images never existed on the host (build mode 2) gives the oracle's bits."""
    rb.set_tuning(build_device=2)
    try:
        s = rb.create_scene(battlefield.vertices, battlefield.indices)
    finally:
        rb.set_tuning(build_device=3)
    rays = random_rays(200_000, np.array(s.info["bounds_min"]) - 20, np.array(s.info["bounds_max"]) + 20, seed=77)
    d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1)).cuda()
    d_res = torch.empty(rays.shape[0] * 4, dtype=torch.float32, device="cuda")
    rb.trace_device(s, env, [(d_rays.data_ptr(), d_res.data_ptr(), rays.shape[0])])
    torch.cuda.synchronize()
    got = d_res.cpu().numpy().view(np.uint32).reshape(-1, 4)
    assert_bit_exact(got, oracle.traverse(images, rays), "device-built scene")


# ---------------------------------------------------------------------------------------------
# independent arbiter at scale: brute-force fp64 Moller-Trumbore over ALL original triangles (no BVH, no pairs),
# evaluated with plain torch ops on the GPU -- shares nothing with the engine or the oracle but the input bytes

def _brute_f64_torch(vertices, indices, rays8, chunk=2048):
    """rays8: (N,8) float32 cuda tensor. Returns t_min (N,) float64 (inf = miss) and the fp64 t of every ray
    against one given triangle per ray via the returned closure."""
    v = torch.from_numpy(np.ascontiguousarray(vertices[:, :3])).cuda().double()
    idx = torch.from_numpy(indices.reshape(-1, 3).astype(np.int64)).cuda()
    p0, p1, p2 = v[idx[:, 0]], v[idx[:, 1]], v[idx[:, 2]]
    e1, e2 = p1 - p0, p2 - p0

    def tri_t(o, d, tmin, tmax, p0s, e1s, e2s):
        # shapes broadcast: rays (n,1,3) against triangles (1,T,3), or row-wise (n,3) against (n,3)
        p = torch.cross(d, e2s, dim=-1)
        det = (e1s * p).sum(-1)
        inv = 1.0 / det
        tv = o - p0s
        u = (tv * p).sum(-1) * inv
        q = torch.cross(tv, e1s, dim=-1)
        w = (d * q).sum(-1) * inv
        t = (e2s * q).sum(-1) * inv
        ok = (det != 0) & (u >= 0) & (u <= 1) & (w >= 0) & (u + w <= 1) & (t > tmin) & (t <= tmax)
        return torch.where(ok, t, torch.full_like(t, float("inf")))

    n = rays8.shape[0]
    t_min = torch.empty(n, dtype=torch.float64, device="cuda")
    for b in range(0, n, chunk):
        r = rays8[b:b + chunk].double()
        o, d = r[:, None, 0:3], r[:, None, 4:7]
        t = tri_t(o.expand(-1, p0.shape[0], -1), d.expand(-1, p0.shape[0], -1), r[:, None, 3], r[:, None, 7], p0[None], e1[None], e2[None])
        t_min[b:b + chunk] = t.min(dim=1).values

    def t_of(tri_ids):
        r = rays8.double()
        k = tri_ids.clamp(min=0).long()
        return tri_t(r[:, 0:3], r[:, 4:7], r[:, 3], r[:, 7], p0[k], e1[k], e2[k])
    return t_min, t_of


def test_fp64_brute_force_arbiter_at_scale(scene, env, battlefield):
    """400 K rays of the full-size bench streams (every 41st primary ray of 1920x1080x4spp and every 43rd of its
    first bounce) against an fp64 brute force over all 64 256 triangles: hit/miss agrees, |t - t64| <= 1e-4 t64,
    and the engine's triangle is in the fp64 tie set (north_star's parity bar, checked by an independent algorithm).
    Rays that graze an edge may legitimately flip between fp32 and fp64; their fraction is bounded and reported."""
    w, h, spp = 1920, 1080, 4
    n = w * h * spp
    d_rays = device_primary(battlefield, w, h, spp, seed=1)
    d_res = torch.empty(n * 4, dtype=torch.float32, device="cuda")
    rb.trace_device(scene, env, [(d_rays.data_ptr(), d_res.data_ptr(), n)])
    torch.cuda.synchronize()
    d_b, nb = device_bounce(scene, d_rays, d_res, n, seed=2)
    d_bres = torch.empty(max(nb, 1) * 4, dtype=torch.float32, device="cuda")
    rb.trace_device(scene, env, [(d_b.data_ptr(), d_bres.data_ptr(), nb)])
    torch.cuda.synchronize()
    rays = torch.cat([d_rays.view(-1, 8)[::41], d_b.view(-1, 8)[:nb][::43]]).contiguous()
    res = torch.cat([d_res.view(-1, 4)[::41], d_bres.view(-1, 4)[:nb][::43]]).contiguous()
    assert rays.shape[0] > 300_000
    t_min, t_of = _brute_f64_torch(battlefield.vertices, battlefield.indices, rays)
    ids = res[:, 0].view(torch.int32)
    hit = ids != -1
    hit64 = torch.isfinite(t_min)
    flips = int((hit != hit64).sum())
    assert flips <= 1e-5 * rays.shape[0] + 2, f"{flips} hit/miss disagreements with fp64 out of {rays.shape[0]}"
    both = hit & hit64
    t = res[:, 1].double()
    rel = ((t - t_min).abs() / t_min)[both]
    assert float(rel.max()) <= 1e-4, float(rel.max())
    t_id = t_of(ids)
    outside = both & ~(t_id <= t_min * (1 + 1e-6) + 1e-9)
    assert int(outside.sum()) <= 1e-5 * rays.shape[0] + 2, f"{int(outside.sum())} engine ids outside the fp64 tie set"
    print(f"fp64 arbiter: {rays.shape[0]} rays, {int(both.sum())} hits, {flips} grazing flips, {int(outside.sum())} ids outside the tie set, "
          f"max |dt|/t = {float(rel.max()):.3e}")


@pytest.mark.timeout(120)
def test_pathological_rays_terminate_and_match(scene, env, images, battlefield):
    """Rays a client can produce by accident: zero and axis-parallel directions (epsilon clamp, Kernels.h:149-157),
    empty intervals (minT >= maxT), huge and tiny magnitudes, and non-finite components. Finite rays must match the
    oracle bit for bit; non-finite ones must simply come back (every variant), never hang the persistent kernel."""
    lo = battlefield.vertices[:, :3].min(0)
    hi = battlefield.vertices[:, :3].max(0)
    base = random_rays(4096, lo, hi, seed=91)
    finite = base.copy()
    finite["dir"][0:256] = 0.0
    finite["dir"][256:512, 0] = 0.0
    finite["dir"][512:768, 1:] = 0.0
    finite["minT"][768:1024] = 10.0
    finite["maxT"][768:1024] = 10.0
    finite["maxT"][1024:1280] = -1.0
    finite["origin"][1280:1536] *= 1e20
    finite["dir"][1536:1792] *= 1e-30
    finite["dir"][1792:2048] *= 1e30
    finite["maxT"][2048:2304] = 3.0e38
    got = trace_dev(scene, env, finite)
    assert_bit_exact(got, oracle.traverse(images, finite), "pathological finite rays")
    weird = base.copy()
    weird["origin"][0:512, 0] = np.nan
    weird["dir"][512:1024, 1] = np.nan
    weird["dir"][1024:1536, 2] = np.inf
    weird["origin"][1536:2048] = -np.inf
    weird["maxT"][2048:2560] = np.inf
    weird["minT"][2560:3072] = np.nan
    for tuning in (dict(), dict(variant=2), dict(variant=0, smem_nodes=-1), dict(variant=1), dict(sort=1, smem_stack=16)):
        rb.set_tuning(**{**DEFAULT, **tuning})
        try:
            out = trace_dev(scene, env, weird)
        finally:
            rb.set_tuning(**DEFAULT)
        assert out.shape == (4096, 4)
        # the untouched tail of the batch is still right
        assert np.array_equal(out[3072:], oracle.traverse(images, weird[3072:]).view(np.uint32).reshape(-1, 4))
