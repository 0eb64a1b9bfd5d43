"""Device-side wavefront path tracer (racc_cuda_path_trace) on BASELINE.json's 1920x1080 frame: rays per second with
rays, hits and path state resident in HBM, beside the reference's unchanged PathTracingRenderer driven through the
drop-in API (host shading, every bounce over PCIe; oracle/_ref/racc_render_gpu when it is there). Development tool.

    python tools/render_bench.py [--width 1920 --height 1080 --spp 16 --depth 3 --reps 5]

Appends a JSON line to gpurun_out/render_bench.jsonl."""
import argparse
import json
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rayaccel_b200 as rb  # noqa: E402
from rayaccel_b200 import scene_io  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--depth", type=int, default=3)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-api", action="store_true")
    ap.add_argument("--whitted", action="store_true", help="the Whitted renderer (racc_cuda_whitted_trace) instead of the path tracer")
    ap.add_argument("--tuning", default="", help="rb.set_tuning keys, e.g. whitted_arena=1,whitted_combine=1")
    ap.add_argument("--same-seed", action="store_true", help="every repetition renders the same frame (same wave sizes)")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    rb.init(0)
    tuning = {k: int(v) for k, v in (kv.split("=") for kv in args.tuning.split(",") if kv)}
    if tuning:
        rb.set_tuning(**tuning)
    sf = rb.load_scene()
    scene = rb.create_scene(sf.vertices, sf.indices)
    env = rb.create_environment(sf.environment)
    shading = rb.create_shading(sf.normals, sf.triangle_normals, sf.materials)
    cam = scene_io.Camera.look_at(sf.cam_origin, sf.cam_target, sf.cam_up, sf.cam_fov, args.width, args.height)
    fb = torch.zeros(args.width * args.height * 4, dtype=torch.float32, device="cuda")
    times, waves = [], None
    for rep in range(args.reps + 2):
        fb.zero_()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        _, waves = (rb.whitted_trace if args.whitted else rb.path_trace)(scene, env, shading, cam, args.width, args.height, args.spp, args.depth,
                                                                         seed=1 if args.same_seed else 1 + rep, framebuffer_ptr=fb.data_ptr(), batch_spp=args.batch)
        b.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if rep >= 2:
            times.append((a.elapsed_time(b) * 1e-3, wall))
    rays = sum(waves)
    best = min(t[0] for t in times)
    med = sorted(t[0] for t in times)[len(times) // 2]
    line = {"what": "device-side Whitted renderer, battlefield" if args.whitted else "device-side wavefront path tracer, battlefield", "width": args.width, "height": args.height, "spp": args.spp,
            "batch_spp": args.batch, "max_depth": args.depth, "tuning": tuning, "same_seed": args.same_seed, "rays_per_frame_set": rays, "waves": waves,
            "ms_all": [round(t[0] * 1e3, 3) for t in times],
            "ms_best": round(best * 1e3, 3), "ms_median": round(med * 1e3, 3), "wall_ms_best": round(min(t[1] for t in times) * 1e3, 3),
            "mrays_best": round(rays / best / 1e6, 1), "mrays_median": round(rays / med / 1e6, 1),
            "mean_radiance": float(fb.view(-1, 4)[:, :3].double().mean().item() / args.spp)}
    exe = os.path.join(ROOT, "oracle", "_ref", "racc_render_gpu")
    if not args.no_api and os.path.exists(exe):
        try:
            out = subprocess.check_output([exe, "--width", str(args.width), "--height", str(args.height), "--frames", "6",
                                           "--scene", os.path.join(ROOT, "data", "battlefield.bin")] + (["--whitted"] if args.whitted else []),
                                          cwd=ROOT, timeout=300)
            info = json.loads(out.decode().strip().splitlines()[-1])
            line["api_host_shading"] = {"what": "reference renderer unchanged through racc::render (host callbacks)",
                                        "mrps": info["mrps"], "mrps_best_frame": info["mrps_best_frame"],
                                        "callback_threads": info["callback_threads"], "mean_radiance": info["mean_luminance"] / 3}
        except Exception as e:  # noqa: BLE001
            line["api_host_shading"] = {"error": str(e)[:200]}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "render_bench.jsonl"), "a") as f:
        f.write(json.dumps(line) + "\n")
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
