"""CPU-only campaign: our host scene builder (racc_cuda_build_images, csrc/scene_build.cpp) against the UNMODIFIED reference builder
compiled from /root/reference (oracle/_ref/libracc_ref.so) on the random scene families of tests/fuzz/fuzz_gpu.py: the
numbering-independent digest of nodes, pairs and remap must be equal. Needs oracle/_ref (this container).

    python tests/fuzz/fuzz_builder_cpu.py [--seconds 120] [--seed 1]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "fuzz"))
import oracle  # noqa: E402
import rayaccel_b200 as rb  # noqa: E402
from fuzz_gpu import scene_family  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120.0)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    assert oracle.have_ref(), "oracle/_ref/libracc_ref.so is not built (needs /root/reference)"
    rng = np.random.default_rng(args.seed)
    t0 = time.time()
    rounds, kinds, failures, skipped = 0, {}, [], 0
    while time.time() - t0 < args.seconds and len(failures) < 5:
        seed = int(rng.integers(0, 2 ** 31))
        kind, verts, indices = scene_family(np.random.default_rng(seed))
        what = f"round {rounds} seed {seed} {kind} {len(indices) // 3} triangles"
        try:
            h = rb.HostImages(verts, indices)
            ours = oracle.SceneImages(h.nodes, h.pairs, h.remap)
        except Exception as e:  # noqa: BLE001
            failures.append(f"{what}: our builder failed: {e}")
            rounds += 1
            continue
        try:
            ref = oracle.ref_build_scene(verts, indices)
        except Exception:  # noqa: BLE001
            ref = None  # the reference builder refuses the scene (its root stays a leaf: ours adds a synthetic root)
        if ref is None:
            skipped += 1
        else:
            try:
                b = ref.digest()
            except Exception:  # noqa: BLE001
                b = None  # a leaf-root scene: the reference's image has no inner node to start from; ours adds a synthetic root
            if b is None:
                skipped += 1
            else:
                a = ours.digest()
                if a != b:
                    failures.append(f"{what}: digests differ: ours {a} reference {b}")
        kinds[kind] = kinds.get(kind, 0) + 1
        rounds += 1
    print(f"builder fuzz: {rounds} scenes in {time.time() - t0:.0f} s, {skipped} leaf-root scenes (the reference has no inner node there), families { {str(k): v for k, v in kinds.items()} }")
    for f in failures:
        print("FAIL", f)
    print("builder fuzz: ok" if not failures else f"builder fuzz: {len(failures)} failures")
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
