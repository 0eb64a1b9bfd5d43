"""Generates tests/golden/ref_cpu_path.npz: what the REFERENCE's CPU query path returns for the golden rays.

The path is executeRayQueryCPU (RayAccelerator/Scene.cpp:374-484) executed from its own source (oracle/_ref/libracc_ref.so,
built by oracle/Makefile where /root/reference is present) over oracle/ref_shim/mini_embree.cpp, the stand-in for the
binary-only Embree 2.7 it calls. Rays: the 6144 golden rays of battlefield_rays.npz (primary, first bounce, random).
Stored beside the results: the fp64 brute-force answer (t64, id64, already in battlefield_rays.npz) is reused by the tests.
Run here (the GPU box has no /root/reference); the committed file lets the GPU tests compare without oracle/_ref.
usage: python tests/golden/make_golden_cpu_path.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from rayaccel_b200 import scene_io  # noqa: E402


def main():
    assert oracle.have_ref_cpu_query(), "build oracle/_ref first (make -C oracle ref)"
    sf = scene_io.load_scene()
    g = np.load(os.path.join(HERE, "battlefield_rays.npz"))
    rays = np.ascontiguousarray(g["rays"]).view(oracle.RAY_DTYPE).reshape(-1)
    res = oracle.ref_cpu_query(sf.vertices, sf.indices, sf.environment, rays, threads=1)
    again = oracle.ref_cpu_query(sf.vertices, sf.indices, sf.environment, rays, threads=5)  # other slice boundaries, scalar tail
    assert res.tobytes() == again.tobytes(), "the reference CPU path is not deterministic across slicings"
    np.savez_compressed(os.path.join(HERE, "ref_cpu_path.npz"), results=res.view(np.uint32).reshape(-1, 4),
                        results_by=np.array("reference executeRayQueryCPU (Scene.cpp:374-484) from its own source over oracle/ref_shim/mini_embree.cpp"))
    hits = int((res["triangle"] != oracle.INVALID).sum())
    print(f"ref_cpu_path.npz: {rays.shape[0]} rays, {hits} hits, {os.path.getsize(os.path.join(HERE, 'ref_cpu_path.npz'))} bytes")


if __name__ == "__main__":
    main()
