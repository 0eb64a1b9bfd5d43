#!/usr/bin/env python
"""Generates tests/golden/ref_render_tiles.npz and ref_whitted_tiles.npz: the images of the reference's UNMODIFIED example
path tracer and Whitted renderer
(Renderer/PathTracingRenderer.cpp + Materials.cpp + Camera.cpp + LightPath.cpp + TiledRenderer.cpp, compiled from
/root/reference by `make -C oracle renderer` into oracle/_ref/racc_render_cpu and driven by
tests/harness/render_headless.cpp) on data/battlefield.bin, reduced to 16x16-pixel tile means. The traversal
under it is the oracle's (tests/harness/fake_capi.cpp), which tests/test_oracle_kat.py pins to the reference's
own kernel. Run in the build container:

    python tests/golden/make_render_golden.py

The reference's random numbers are seeded from libc rand() per call and its work is spread over threads, so the
image is a Monte-Carlo estimate, not a fixed vector: 512 frames (one sample per pixel each) bring the tile means
to a few tenths of a percent. tests/test_render_oracle.py and tests/test_gpu_render.py compare against it with
tolerances derived from that noise."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
WIDTH, HEIGHT, FRAMES, TILE = 256, 128, 512, 16


def render(exe, frames, extra, out_name, made_by):
    with tempfile.TemporaryDirectory() as tmp:
        dump = os.path.join(tmp, "fb.f32")
        out = subprocess.check_output([exe, "--width", str(WIDTH), "--height", str(HEIGHT), "--frames", str(frames), "--dump", dump,
                                       "--scene", os.path.join(ROOT, "data", "battlefield.bin")] + extra, cwd=ROOT)
        info = json.loads(out.decode().strip().splitlines()[-1])
        fb = np.fromfile(dump, dtype=np.float32).reshape(HEIGHT, WIDTH, 4)
    mean = fb[..., :3].astype(np.float64) / frames
    tiles = mean.reshape(HEIGHT // TILE, TILE, WIDTH // TILE, TILE, 3).mean(axis=(1, 3)).astype(np.float32)
    rays_per_frame = (info["rays_first_frame"] + info["rays_timed"]) / frames
    np.savez_compressed(os.path.join(HERE, out_name), tiles=tiles, width=WIDTH, height=HEIGHT, frames=frames, tile=TILE,
                        max_depth=info["max_depth"], rays_per_frame=rays_per_frame, mean=np.float64(mean.mean()),
                        made_by=made_by % frames)
    print(out_name, "tiles", tiles.shape, "mean radiance", mean.mean(), "rays per frame", rays_per_frame, "max_depth", info["max_depth"])


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "racc_render_cpu")
    assert os.path.exists(exe), "build oracle/_ref first (make -C oracle renderer)"
    which = sys.argv[1:] or ["path", "whitted"]
    if "path" in which:
        render(exe, FRAMES, [], "ref_render_tiles.npz", "reference PathTracingRenderer (oracle/_ref/racc_render_cpu), %d frames")
    if "whitted" in which:
        # deterministic apart from the pixel jitter: 128 frames are plenty; depth 8 as Renderer/main.cpp:346 forces for this renderer
        render(exe, 128, ["--whitted"], "ref_whitted_tiles.npz", "reference WhittedRenderer (oracle/_ref/racc_render_cpu --whitted), %d frames")


if __name__ == "__main__":
    sys.exit(main())
