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
