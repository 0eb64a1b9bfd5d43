"""Scene build (SAH BVH2 + greedy pair merge + node packing) of the engine's host builder against the
UNMODIFIED reference builder: live when oracle/_ref is present, and always against the committed
digests that were produced by executing the reference (tests/golden/ref_scene_digests.json).
CPU only -- racc_cuda_build_images makes no CUDA call."""
import json
import os
import sys

import numpy as np
import pytest

import oracle
import rayaccel_b200 as rb

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLDEN)
from make_golden import synthetic_meshes  # noqa: E402
from conftest import random_rays  # noqa: E402


def own_images(v, i):
    h = rb.HostImages(v, i)
    return oracle.SceneImages(h.nodes, h.pairs, h.remap), h.info


@pytest.fixture(scope="module")
def golden_digests():
    with open(os.path.join(GOLDEN, "ref_scene_digests.json")) as f:
        return json.load(f)


def test_battlefield_digest_matches_reference_golden(battlefield, golden_digests):
    img, info = own_images(battlefield.vertices, battlefield.indices)
    assert img.digest() == golden_digests["battlefield"]
    # the survey's structural facts
    assert info["node_count"] == 19192 and info["real_pair_count"] == 35823 and info["depth"] == 21
    assert info["pair_count"] == 35840 and info["remap_count"] == 71646


@pytest.mark.parametrize("name", sorted(synthetic_meshes().keys()))
def test_synthetic_digest_matches_reference_golden(name, golden_digests):
    v, i = synthetic_meshes()[name]
    img, _ = own_images(v, i)
    assert img.digest() == golden_digests[name]


def test_digest_matches_reference_live(battlefield):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    img, _ = own_images(battlefield.vertices, battlefield.indices)
    assert img.digest() == oracle.ref_build_scene(battlefield.vertices, battlefield.indices).digest()
    for seed in (21, 22):
        v, i = rb.synthetic_triangles(5000 + seed, seed=seed, extent=50.0, edge=4.0)
        assert own_images(v, i)[0].digest() == oracle.ref_build_scene(v, i).digest()


def test_image_invariants(battlefield):
    img, info = own_images(battlefield.vertices, battlefield.indices)
    nodes_u = img.nodes.view(np.uint32)
    refs = nodes_u[:, 2:4].ravel()
    inner = refs[(refs & 0x80000000) != 0] & 0x7FFFFFFF
    leaves = refs[(refs & 0x80000000) == 0]
    # every inner node except the root is referenced exactly once; parents precede children (hot-first order)
    assert np.array_equal(np.sort(inner), np.arange(1, info["node_count"]))
    parent_of = np.repeat(np.arange(info["node_count"]), 2)[(refs & 0x80000000) != 0]
    assert np.all(parent_of < inner)
    # leaves tile the pair array without gaps or overlap
    first, count = leaves & 0xFFFFFF, leaves >> 24
    order = np.argsort(first)
    assert first[order][0] == 0 and np.all(first[order][1:] == (first[order] + count[order])[:-1])
    assert (first + count).max() == info["real_pair_count"] and count.min() >= 1 and count.max() <= 127
    # remap: every original triangle appears exactly once; unused second slots hold 0
    words = img.remap
    singles = np.isclose(img.pairs[: info["real_pair_count"], [3, 7, 11]], -img.pairs[: info["real_pair_count"], [0, 1, 2]]).all(axis=1)
    used = np.ones(words.shape[0], bool)
    used[1::2] = ~singles
    ids = words[used] & 0x3FFFFFFF
    assert np.array_equal(np.sort(ids), np.arange(battlefield.triangle_count))
    assert np.all(words[~used] == 0)
    # pair array padded to a multiple of 32 float4 (Scene.cpp:335-338)
    assert (info["pair_count"] * 3) % 32 == 0 and info["pair_count"] > info["real_pair_count"]
    # child boxes: parent box contains both children's boxes
    lo = np.minimum(img.nodes[:, 4:7], img.nodes[:, 10:13])
    hi = np.maximum(img.nodes[:, 7:10], img.nodes[:, 13:16])
    kids = np.nonzero((nodes_u[:, 2] & 0x80000000) != 0)[0]
    c = nodes_u[kids, 2] & 0x7FFFFFFF
    assert np.all(lo[c] >= img.nodes[kids, 4:7]) and np.all(hi[c] <= img.nodes[kids, 7:10])


def test_pair_geometry_reconstructs_triangles(battlefield):
    """Each pair triangle, rebuilt from (e1,e2,e3,p0) and un-rotated by its edge code, is the original
    triangle (Scene.cpp:122-181)."""
    img, info = own_images(battlefield.vertices, battlefield.indices)
    n = info["real_pair_count"]
    p = img.pairs[:n]
    e1, e2, p0 = p[:, 0:3], p[:, 4:7], p[:, 8:11]
    e3 = p[:, [3, 7, 11]]
    p1, p2, p3 = p0 - e1, p0 + e2, p0 + e3
    v = battlefield.vertices[:, :3]
    tri = battlefield.indices.reshape(-1, 3)
    for slot, (a, b, c) in enumerate(((p0, p1, p2), (p0, p3, p1))):
        w = img.remap[slot::2][:n]
        valid = np.ones(n, bool) if slot == 0 else ~np.isclose(p3, p1).all(axis=1)
        ids, code = w & 0x3FFFFFFF, w >> 30
        orig = v[tri[ids]]  # (n, 3 verts, 3)
        # code k: pair verts (a,b,c) = original (v[k%3], v[(k+1)%3], v[(k+2)%3])  (reorder(), Scene.cpp:100-107)
        perm = {0: (0, 1, 2), 3: (0, 1, 2), 1: (1, 2, 0), 2: (2, 0, 1)}  # code 3 = rotation by 3 = none (Scene.cpp:133)
        for k, (i0, i1, i2) in perm.items():
            m = valid & (code == k)
            if not m.any():
                continue
            assert np.allclose(a[m], orig[m, i0], atol=2e-4) and np.allclose(b[m], orig[m, i1], atol=2e-4) and np.allclose(c[m], orig[m, i2], atol=2e-4), (slot, k)


def test_invalid_inputs_are_rejected():
    v, i = rb.synthetic_triangles(10, seed=1)
    with pytest.raises(ValueError):
        rb.create_scene(v, i[:-1])  # indexCount % 3 (Scene.cpp:186)
    with pytest.raises(rb.EngineError, match="multiple of 3"):
        rb.HostImages(v, i[:-1])
    with pytest.raises(rb.EngineError, match="out of range"):
        rb.HostImages(v[:5], i)
    with pytest.raises(rb.EngineError, match="no triangles"):
        rb.HostImages(v, i[:0])


def _brute_hits(v, i, rays):
    t64, _ = oracle.brute_f64(v, i, rays)
    return np.isfinite(t64), t64


@pytest.mark.parametrize("name", ["one", "two", "cube", "five_large", "coincident"])
def test_scenes_whose_root_stays_a_leaf(name):
    """The SAH test may decline to split 1..126 triangles; the reference then uploads an empty node image (Scene.cpp:274-342:
    `&nodes[0]` of an empty vector) and cannot trace the scene. Here the image gets one synthetic inner root whose two
    children are that leaf, and the traversal gives the brute-force answer."""
    rng = np.random.default_rng(5)
    if name == "cube":
        c = np.array([[x, y, z, 1] for x in (0, 1) for y in (0, 1) for z in (0, 1)], np.float32)
        quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
        v, i = c, np.array([[a, b, cc, a, cc, d] for a, b, cc, d in quads], np.uint32).ravel()
    elif name == "coincident":
        one = np.array([[0, 0, 0, 1], [1, 0, 0, 1], [0, 1, 0, 1]], np.float32)
        v, i = np.tile(one, (40, 1)), np.arange(120, dtype=np.uint32)
    else:
        n = {"one": 1, "two": 2, "five_large": 5}[name]
        v = np.ones((3 * n, 4), np.float32)
        v[:, :3] = rng.uniform(-5, 5, size=(3 * n, 3))
        i = np.arange(3 * n, dtype=np.uint32)
    h = rb.HostImages(v, i)
    assert h.info["node_count"] >= 1 and h.info["triangle_count"] == i.shape[0] // 3
    img = oracle.SceneImages(h.nodes, h.pairs, h.remap)
    lo, hi = v[:, :3].min(0) - 1, v[:, :3].max(0) + 1
    rays = random_rays(4000, lo, hi, seed=3)
    res = oracle.traverse(img, rays)
    hit64, t64 = _brute_hits(v, i, rays)
    hit = res["triangle"] != oracle.INVALID
    assert (hit != hit64).sum() <= 2  # grazing rays may flip between fp32 and fp64
    both = hit & hit64
    assert both.any() or name == "one"
    assert np.all(np.abs(res["a"][both] - t64[both]) <= 1e-4 * np.abs(t64[both]) + 1e-6)


def test_build_is_deterministic_and_thread_count_independent(battlefield):
    a, _ = own_images(battlefield.vertices, battlefield.indices)
    os.environ["RACC_B200_BUILD_THREADS"] = "1"
    try:
        b, _ = own_images(battlefield.vertices, battlefield.indices)
    finally:
        del os.environ["RACC_B200_BUILD_THREADS"]
    assert np.array_equal(a.nodes.view(np.uint32), b.nodes.view(np.uint32))
    assert np.array_equal(a.pairs.view(np.uint32), b.pairs.view(np.uint32))
    assert np.array_equal(a.remap, b.remap)


def test_rcpss_table_model_of_the_device_builder():
    """The device scene builder replays the reference's _mm_rcp_ss leaf-cost test (Bvh2.cpp:462-467) through a
    table of this host's RCPSS values; check the table is a sane 12-bit reciprocal and that the model holds here."""
    import ctypes
    from rayaccel_b200 import _lib
    table = (ctypes.c_float * 2048)()
    rc = _lib.load().racc_cuda_debug_rcp_table(table)
    t = np.frombuffer(table, dtype=np.float32)
    x = 1.0 + np.arange(2048) / 2048.0
    assert np.all(np.abs(t * x - 1.0) <= 1.5 * 2.0 ** -12 + 2.0 ** -11), "not a reciprocal approximation"
    assert np.all(np.diff(t) <= 0), "RCPSS table must be monotone"
    assert rc == 0, "this CPU's RCPSS does not follow the table model: the device builder will decline (host build is used)"


def test_interleaved_blocks_partition():
    from rayaccel_b200 import sharding
    for total, world, block in ((10, 3, 4), (4097, 2, 500), (8294400, 8, 16384), (5, 8, 16)):
        seen = np.zeros(total, dtype=np.int32)
        for r in range(world):
            for b, e in sharding.interleaved_blocks(total, r, world, block):
                assert 0 <= b < e <= total and e - b <= block
                seen[b:e] += 1
        assert np.all(seen == 1)
