"""Hand-built scenes in the reference's GPU image format (Scene.cpp:73-87,132-133,294-312) and the
known answers of the rules they probe. Shared by the CPU oracle tests and the GPU parity tests."""
import numpy as np

import oracle
from conftest import make_rays

INNER = 0x80000000


def pair(p0, p1, p2, p3=None):
    """TrianglePair (Scene.cpp:80-87): tri0 = (p0,p1,p2), tri1 = (p0,p3,p1); p3=None -> singleton (p3=p1)."""
    p0, p1, p2 = (np.asarray(p, np.float32) for p in (p0, p1, p2))
    p3 = p1 if p3 is None else np.asarray(p3, np.float32)
    e1, e2, e3 = p0 - p1, p2 - p0, p3 - p0
    return np.array([e1[0], e1[1], e1[2], e3[0], e2[0], e2[1], e2[2], e3[1], p0[0], p0[1], p0[2], e3[2]], np.float32)


def node(first, last, lmin, lmax, rmin, rmax):
    n = np.zeros(16, np.float32)
    u = n.view(np.uint32)
    u[0], u[1], u[2], u[3] = 1, 0xFFFFFFFF, first, last
    n[4:7], n[7:10], n[10:13], n[13:16] = lmin, lmax, rmin, rmax
    return n


def leaf(first_pair, count):
    return (count << 24) | first_pair


def two_leaf_scene(pairs_left, pairs_right, remap, box_l, box_r):
    pairs = np.stack(list(pairs_left) + list(pairs_right))
    nodes = node(leaf(0, len(pairs_left)), leaf(len(pairs_left), len(pairs_right)), box_l[0], box_l[1], box_r[0], box_r[1])[None, :]
    return oracle.SceneImages(nodes, pairs, np.asarray(remap, np.uint32))


BIG = ([-100, -100, -100], [100, 100, 100])
FAR_AWAY = pair([500, 500, 500], [501, 500, 500], [500, 501, 500])

# unit quad in the z=0 plane split along the diagonal p0-p1: tri0=(p0,p1,p2) tri1=(p0,p3,p1)
P0, P1, P2, P3 = [0, 0, 0], [1, 1, 0], [0, 1, 0], [1, 0, 0]
QUAD = pair(P0, P1, P2, P3)

KAT_CASES = [
    dict(name="interior_hits_and_bary",
         # remap: pair-triangle 0 -> original 7, pair-triangle 1 -> original 9
         scene=lambda: two_leaf_scene([QUAD], [FAR_AWAY], [7, 9, 3, 0], BIG, ([499, 499, 499], [502, 502, 502])),
         rays=lambda: make_rays([[0.25, 0.75, 5], [0.75, 0.25, 5], [3, 3, 5]], [[0, 0, -1]] * 3),
         # tri0 = (p0,p1,p2): point = p0 + u*(p1-p0) + v*(p2-p0) -> (0.25,0.75): u=0.25, v=0.5
         # tri1 = (p0,p3,p1): (0.75,0.25): u (weight of p3)=0.5, v (weight of p1)=0.25
         expect=[dict(triangle=7, t=5.0, u=0.25, v=0.5), dict(triangle=9, t=5.0, u=0.5, v=0.25), None]),
    dict(name="shared_edge_tie_goes_to_first_triangle",  # Kernels.h:97
         scene=lambda: two_leaf_scene([QUAD], [FAR_AWAY], [0, 1, 2, 0], BIG, ([499, 499, 499], [502, 502, 502])),
         rays=lambda: make_rays([[0.5, 0.5, 2]], [[0, 0, -1]]),
         expect=[dict(triangle=0, t=2.0)]),
    dict(name="coplanar_duplicates_later_pair_wins",  # accept iff T <= det*tMax (Kernels.h:88-89)
         scene=lambda: two_leaf_scene([pair(P0, P1, P2), pair(P0, P1, P2)], [FAR_AWAY], [4, 0, 5, 0, 6, 0], BIG, ([499, 499, 499], [502, 502, 502])),
         rays=lambda: make_rays([[0.25, 0.75, 1]], [[0, 0, -1]]),
         expect=[dict(triangle=5, t=1.0)]),
    dict(name="t_equal_minT_rejected_t_equal_maxT_accepted",  # T > det*tNear && T <= det*tMax
         scene=lambda: two_leaf_scene([pair(P0, P1, P2)], [FAR_AWAY], [0, 0, 1, 0], BIG, ([499, 499, 499], [502, 502, 502])),
         rays=lambda: np.concatenate([make_rays([[0.25, 0.75, 2]], [[0, 0, -1]], tmin=2.0, tmax=10.0),
                                      make_rays([[0.25, 0.75, 2]], [[0, 0, -1]], tmin=0.0, tmax=2.0),
                                      make_rays([[0.25, 0.75, 2]], [[0, 0, -1]], tmin=0.0, tmax=1.9999)]),
         expect=[None, dict(triangle=0, t=2.0), None]),
    dict(name="axis_parallel_direction_epsilon_clamp",  # Kernels.h:149-157
         scene=lambda: two_leaf_scene([pair(P0, P1, P2)], [FAR_AWAY], [0, 0, 1, 0], ([0, 0, 0], [1, 1, 0]), ([499, 499, 499], [502, 502, 502])),
         rays=lambda: make_rays([[0.25, 0.75, 3], [0.25, 0.75, -3]], [[0, 0, -1], [-0.0, 0.0, 1]]),
         expect=[dict(triangle=0, t=3.0), dict(triangle=0, t=3.0)]),
    dict(name="origin_inside_box",
         scene=lambda: two_leaf_scene([pair(P0, P1, P2)], [FAR_AWAY], [0, 0, 1, 0], ([-5, -5, -5], [5, 5, 5]), ([499, 499, 499], [502, 502, 502])),
         rays=lambda: make_rays([[0.25, 0.75, 1]], [[0, 0, -1]]),
         expect=[dict(triangle=0, t=1.0)]),
    dict(name="singleton_second_triangle_never_hits",  # p3 = p1 -> n2 = 0 (Scene.cpp:146-151)
         scene=lambda: two_leaf_scene([pair(P0, P1, P2)], [FAR_AWAY], [3, 0, 1, 0], BIG, ([499, 499, 499], [502, 502, 502])),
         rays=lambda: make_rays([[0.75, 0.25, 1], [0.25, 0.75, 1]], [[0, 0, -1]] * 2),
         expect=[None, dict(triangle=3, t=1.0)]),
    dict(name="edge_rotation_codes",  # Kernels.h:227-235: code 1 -> barys.zxy, code 2 -> barys.yzx
         # the same geometric triangle stored as pair triangle (p0,p1,p2) but whose ORIGINAL vertex order
         # was rotated: code 1 means original = (p2,p0,p1)... expressed through the remap word only
         scene=lambda: two_leaf_scene([pair(P0, P1, P2)], [pair([10, 0, 0], [11, 1, 0], [10, 1, 0])],
                                      [(1 << 30) | 11, 0, (2 << 30) | 12, 0], ([0, 0, 0], [1, 1, 0]), ([10, 0, 0], [11, 1, 0])),
         rays=lambda: make_rays([[0.25, 0.75, 1], [10.25, 0.75, 1]], [[0, 0, -1]] * 2),
         # pair barycentrics (u,v) = (0.25, 0.5), w = 0.25.  code 1: (u,v) <- (w,u) = (0.25,0.25); code 2: (u,v) <- (v,w) = (0.5,0.25)
         expect=[dict(triangle=11, t=1.0, u=0.25, v=0.25), dict(triangle=12, t=1.0, u=0.5, v=0.25)]),
    dict(name="nearer_of_two_leaves_and_far_child_pruned",
         scene=lambda: two_leaf_scene([pair([0, 0, -4], [1, 1, -4], [0, 1, -4])], [pair(P0, P1, P2)], [20, 0, 21, 0],
                                      ([0, 0, -4], [1, 1, -4]), ([0, 0, 0], [1, 1, 0])),
         rays=lambda: make_rays([[0.25, 0.75, 3], [0.25, 0.75, -7]], [[0, 0, -1], [0, 0, 1]]),
         expect=[dict(triangle=21, t=3.0), dict(triangle=20, t=3.0)]),
    dict(name="backface_hits_count",  # sign trick: both orientations intersect (Kernels.h:60-72)
         scene=lambda: two_leaf_scene([QUAD], [FAR_AWAY], [0, 1, 2, 0], BIG, ([499, 499, 499], [502, 502, 502])),
         rays=lambda: make_rays([[0.25, 0.75, -2], [0.75, 0.25, -2]], [[0, 0, 1]] * 2),
         expect=[dict(triangle=0, t=2.0), dict(triangle=1, t=2.0)]),
]


def build_kat_scene(case):
    return case["scene"](), case["rays"]()


def deep_stack_scene(levels=60, n_rays=4096, seed=17):
    """A chain of `levels` inner nodes built so that a ray travelling along +z pushes one stack entry per level: at level
    k the FIRST child is a far leaf (one triangle at z = 100 + k) and the LAST child is the next inner node, whose box
    starts at z = 0 -- nearer, so the ray descends into the chain and the far leaf goes onto the stack (Kernels.h:192-197).
    The chain's boxes narrow in x with k, so rays with different |x| leave the chain at different levels (lanes of a warp
    hold stacks of different depths), and each triangle covers only y < its own threshold, so some popped leaves hit and
    some miss. With levels > 17 the stack outgrows the 16 entries the packed kernel may keep in shared memory
    (HybridStack's spill branch), with 60 it comes close to the reference's 64 (Kernels.h:166). Leaves come off the stack
    farthest first, so every hit replaces the previous one: the answer is the triangle of the lowest level whose
    threshold the ray is under. Returns (SceneImages, rays, expected triangle per ray)."""
    rng = np.random.default_rng(seed)
    half = float(levels + 4)
    nodes, pairs, remap = [], [], []
    thresholds = rng.uniform(-0.8 * half, 0.8 * half, size=levels + 1)
    for k in range(levels + 1):  # leaf k: one triangle in the plane z = 100 + k covering y < thresholds[k]
        z = 100.0 + k
        pairs.append(pair([-4 * half, thresholds[k], z], [4 * half, thresholds[k], z], [0.0, thresholds[k] - 8 * half, z]))
        remap += [1000 + k, 0]
    for k in range(levels):
        w = half - k * (half - 2.0) / levels  # chain box half-width at level k
        leaf_box = ([-half, -half, 100.0 + k], [half, half, 100.0 + k])
        if k + 1 < levels:
            w_next = half - (k + 1) * (half - 2.0) / levels
            nodes.append(node(leaf(k, 1), INNER | (k + 1), leaf_box[0], leaf_box[1], [-w_next, -half, 0.0], [w_next, half, 50.0]))
        else:  # the end of the chain: two leaves
            nodes.append(node(leaf(k, 1), leaf(k + 1, 1), leaf_box[0], leaf_box[1], [-w, -half, 100.0 + k + 1], [w, half, 100.0 + k + 1]))
    pairs = np.stack(pairs)
    pairs = np.concatenate([pairs, np.repeat(pairs[:1], (-pairs.shape[0]) % 32 or 32, axis=0)])  # tail padding (Scene.cpp:335-338)
    images = oracle.SceneImages(np.stack(nodes), pairs, np.asarray(remap, np.uint32))
    x = rng.uniform(-half, half, size=n_rays)
    y = rng.uniform(-half, half, size=n_rays)
    rays = make_rays(np.stack([x, y, np.full(n_rays, -1.0)], axis=1), [[0.0, 0.0, 1.0]] * n_rays)
    # reachable leaves for |x|: 0..m where m = number of chain boxes entered; the last node adds leaf `levels` within its box
    expect = np.full(n_rays, 0xFFFFFFFF, np.uint32)
    for r in range(n_rays):
        reach = [0]
        for k in range(1, levels):
            if abs(x[r]) <= half - k * (half - 2.0) / levels:
                reach.append(k)
            else:
                break
        else:
            if abs(x[r]) <= half - (levels - 1) * (half - 2.0) / levels:
                reach.append(levels)
        for k in reach:
            if y[r] < thresholds[k] - 1e-3 and abs(x[r]) < 4 * half * (1 - (thresholds[k] - y[r]) / (8 * half)) - 1e-3:
                expect[r] = 1000 + k
                break
            if abs(y[r] - thresholds[k]) <= 1e-3:
                expect[r] = 0xFFFFFFFE  # too close to an edge to call by hand
                break
    return images, rays, expect


def shared_edge_mesh_case():
    """The scene and rays on which the randomised campaign (tests/fuzz/fuzz_gpu.py, round 2) caught the CHECKER leaving the
    source's semantics: an integer-grid height field (15 x 6 cells, 180 triangles) and 17 193 rays aimed at its vertices,
    edge midpoints and interiors. Five of them pass through a shared edge such that dot(R, e3) of Kernels.h:66 cancels to an
    exact +0; its negation is -0 (sign bit set: outside the pair's second triangle, the neighbour owns the edge). gcc folds
    that negation into the dot product's last fma (vfnmsub), which yields +0 -- the other triangle. Returns (vertices
    (N,4) f32, indices u32, rays, {ray index: (triangle, t bits, u bits, v bits)} as the source's semantics give them."""
    import numpy as np
    import oracle
    rng = np.random.default_rng(1368281523)
    rng.choice(["soup", "mesh", "coincident", "slivers", "quads", "huge", "tiny", "offset", "few"])
    w, h = int(rng.integers(2, 60)), int(rng.integers(2, 60))
    xs, ys = np.meshgrid(np.arange(w + 1, dtype=np.float64), np.arange(h + 1, dtype=np.float64))
    z = rng.normal(0, rng.uniform(0.0, 2.0), xs.shape)
    v = np.stack([xs, z, ys], -1).reshape(-1, 3)
    q = (np.arange(h)[:, None] * (w + 1) + np.arange(w)[None, :]).reshape(-1)
    indices = np.stack([q, q + 1, q + w + 2, q, q + w + 2, q + w + 1], -1).reshape(-1).astype(np.uint32)
    verts = np.zeros((v.shape[0], 4), np.float32)
    verts[:, :3] = v.astype(np.float32)
    rng.choice([0, 2, 3])
    if rng.random() < 0.6:
        rng.random((int(rng.integers(1, 9)), int(rng.integers(1, 9)), 4))
    rng.random()
    n = int(rng.integers(1, 20000))
    # the ray recipe of tests/fuzz/fuzz_gpu.py: rays_for()
    vv = verts[:, :3].astype(np.float64)
    tri = vv[indices.reshape(-1, 3).astype(np.int64)]
    lo, hi = vv.min(0), vv.max(0)
    ext = np.maximum(hi - lo, 1e-30)
    pick = tri[rng.integers(0, tri.shape[0], n)]
    bary = rng.dirichlet([1, 1, 1], n)
    mode = rng.integers(0, 6, n)
    target = (pick * bary[:, :, None]).sum(1)
    target[mode == 1] = pick[mode == 1, rng.integers(0, 3)]
    target[mode == 2] = 0.5 * (pick[mode == 2, 0] + pick[mode == 2, 1])
    origin = lo + rng.uniform(-0.5, 1.5, (n, 3)) * ext
    on = mode == 3
    origin[on] = target[on]
    target[on] = lo + rng.uniform(0, 1, (int(on.sum()), 3)) * ext
    d = target - origin
    rnd = mode >= 4
    d[rnd] = rng.normal(size=(int(rnd.sum()), 3))
    ln = np.linalg.norm(d, axis=1, keepdims=True)
    d = np.where(ln > 0, d / np.maximum(ln, 1e-300), [[1.0, 0.0, 0.0]])
    axis = rng.random(n) < 0.08
    k = rng.integers(0, 3, n)
    d[axis] = 0.0
    d[axis, k[axis]] = rng.choice([-1.0, 1.0], int(axis.sum()))
    rays = np.zeros(n, dtype=oracle.RAY_DTYPE)
    rays["origin"], rays["dir"] = origin.astype(np.float32), d.astype(np.float32)
    rays["minT"] = np.where(rng.random(n) < 0.2, rng.choice([1e-3, 1e-6, 1.0]) * float(ext.max()), 0.0).astype(np.float32)
    rays["maxT"] = np.float32(1e6) * np.float32(max(1.0, float(ext.max())))
    short = rng.random(n) < 0.1
    rays["maxT"][short] = (np.linalg.norm(ln[short], axis=1) * rng.uniform(0.3, 1.3, int(short.sum()))).astype(np.float32)
    empty = rng.random(n) < 0.02
    rays["maxT"][empty] = rays["minT"][empty]
    # what the source's semantics (and the B200) give; the folded build said triangles 23, 153, 53, MISS (a crack at a vertex), 142
    known = {523: (22, 1107818876, 0, 1056964610), 1917: (142, 1103844780, 0, 1065353216), 6883: (55, 1102174747, 1065353212, 0),
             11445: (142, 1091516065, 884998144, 1065353210), 17187: (145, 1103608929, 869358250, 1065353212)}
    return verts, indices, rays, known


def bound_vertex_case(seed=7):
    """Four random triangles and rays aimed exactly at their vertices -- among them the vertices that ARE the scene's lower
    and upper bounds, where the quantised-node grid (variant 4) once had no margin: a ray through such a vertex slipped past
    the root box (tests/fuzz/fuzz_gpu.py, round 2). Returns (vertices (N,4) f32, indices u32, rays)."""
    import numpy as np
    import oracle
    rng = np.random.default_rng(seed)
    v = rng.uniform(-5, 5, (12, 3)).astype(np.float32)
    verts = np.zeros((12, 4), np.float32)
    verts[:, :3] = v
    indices = np.arange(12, dtype=np.uint32)
    n = 12 * 600
    target = np.repeat(v.astype(np.float64), 600, axis=0)
    origin = rng.uniform(-12, 12, (n, 3))
    d = target - origin
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros(n, dtype=oracle.RAY_DTYPE)
    rays["origin"], rays["dir"], rays["minT"], rays["maxT"] = origin.astype(np.float32), d.astype(np.float32), 0.0, 1e6
    return verts, indices, rays
