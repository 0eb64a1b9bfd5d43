#!/usr/bin/env python
"""Headline benchmark: Mrays/s (primary + incoherent secondary) on battlefield.bin.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): path-tracer wavefront on battlefield.bin, 1920x1080, 4 spp,
depth 3 (the scene header's maxDepth: primary rays + 3 diffuse bounces, Renderer/
PathTracingRenderer.cpp:100,120). A "step" is ONE pass of the hot path over that batch: all four ray
streams (~17 M rays, 0.8 GB of rays+results, larger than L2) traced by one launch of the
persistent traversal kernel, device-resident. `value` = rays/s over all ranks with inputs in HBM;
`e2e` = the same batch through the reference-facing C-ABI with pinned HOST buffers (H2D + trace +
D2H inside the timed region). Multi-GPU is weak scaling: the scene is replicated, every rank traces
its own batch (different jitter seed), and the only collective is the per-frame hit-count
all-reduce (NCCL).

The oracle (oracle/) is used here only as the checker / CPU baseline: `cpu_baseline` times it on a
bounded sample of the same rays on rank 0 at N=1, and `--impl reference` times it alone (the
reference's own CPU path is Embree 2.7, binary-only for macOS/Windows, so the port stands in).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Mrays/s (primary + incoherent secondary) on battlefield.bin"
WIDTH, HEIGHT, SPP, BOUNCES = 1920, 1080, 4, 3
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


_JSON_OUT = None


def _claim_stdout() -> None:
    """Keep stdout for the ONE JSON line: native libraries (NCCL prints its version banner to stdout)
    write to fd 1, so fd 1 is pointed at stderr for the run and the line goes to the saved descriptor."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def bind_near_gpu(local_rank: int):
    """Multi-rank runs: pin this process to the CPUs of the NUMA node its GPU hangs off BEFORE any pinned host
    buffer is allocated (first touch then places the staging pages next to the GPU's PCIe root). Without this the
    8 ranks' H2D/D2H streams all cross the socket interconnect. Best effort: returns a note for the JSON line."""
    try:
        import torch
        props = torch.cuda.get_device_properties(local_rank)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return "numa node unknown"
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return f"numa node {node}: none of its CPUs are available to this process"
        os.sched_setaffinity(0, allowed)
        return f"numa node {node}, {len(allowed)} CPUs"
    except Exception as e:  # no sysfs, no permission, older torch: run unbound
        return f"unbound ({type(e).__name__})"


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this
    same command (profiles/ncu_bench_traffic.json, written by tools/ncu_traffic.py); None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_bench_traffic.json")) as f:
            t = json.load(f)
        return float(t["dram_bytes_per_launch"]), t.get("source", "profiles/ncu_bench_traffic.json"), t.get("limiters")
    except Exception:
        return None, None, None


def workload_name():
    return (f"path tracer wavefront, battlefield.bin, {WIDTH}x{HEIGHT}, {SPP} spp, depth {BOUNCES} "
            f"(primary + {BOUNCES} diffuse bounces)")


# --------------------------------------------------------------------------------------------
# clocks


class ClockSampler:
    """nvidia-smi sampled every 200 ms during the timed region (B200_PROFILING.md recipe)."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="racc_clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.device_index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx = float(f[2])
                except ValueError:
                    continue
                for name, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


# --------------------------------------------------------------------------------------------
# reference arm: the CPU port, all host threads, bounded sample; never touches the engine


def sample_rows(target_rays: int) -> np.ndarray:
    """Evenly strided pixel rows so that rows*WIDTH*SPP*(1+bounce mix) is about target_rays."""
    per_row = WIDTH * SPP * 2.05  # primary + ~1.05 secondary rays per primary on battlefield
    n_rows = int(max(1, min(HEIGHT, round(target_rays / per_row))))
    return np.unique(np.linspace(0, HEIGHT - 1, n_rows).astype(np.int64))


def build_cpu_sample(rows: np.ndarray, seed: int):
    """The bench batch restricted to `rows`, generated and traced on the CPU only."""
    import oracle
    from oracle import raygen
    from rayaccel_b200 import scene_io  # file loader only (no engine call)
    sf = scene_io.load_scene()
    if oracle.have_ref():
        images = oracle.ref_build_scene(sf.vertices, sf.indices)  # the unmodified reference builder
        images.env = sf.environment
        built_by = "unmodified reference builder (oracle/_ref)"
    else:
        # oracle/_ref is built where /root/reference exists and travels as a prebuilt .so; without it
        # the host-only half of the engine's scene build (no CUDA) supplies the structurally
        # identical images (tests/test_scene_build.py).
        from rayaccel_b200 import HostImages
        h = HostImages(sf.vertices, sf.indices)
        images = oracle.SceneImages(h.nodes, h.pairs, h.remap, sf.environment)
        built_by = "engine host builder (oracle/_ref absent)"
    cam = raygen.look_at(sf.cam_origin, sf.cam_target, sf.cam_up, sf.cam_fov, WIDTH, HEIGHT)
    streams = [raygen.primary_rays(cam, WIDTH, HEIGHT, SPP, seed, rows=rows)]
    for b in range(BOUNCES):
        res = oracle.traverse(images, streams[-1])
        streams.append(raygen.bounce_rays(sf.vertices, sf.indices, streams[-1], res, seed + 1 + b))
    return images, streams, built_by


def run_reference(args, rank: int, world: int) -> None:
    if rank != 0:
        return  # rank 0 alone runs the reference arm
    import oracle
    try:
        oracle.build(ref=os.path.isdir("/root/reference/RayAccelerator"))
    except Exception:
        pass  # prebuilt liboracle.so / _ref travel with the snapshot
    cores = os.cpu_count() or 1
    # calibrate, then size the per-step sample so the whole run stays within ~2 minutes
    images, streams, built_by = build_cpu_sample(sample_rows(60_000), seed=1)
    rays = np.concatenate(streams)
    t0 = time.perf_counter()
    oracle.traverse_avx2(images, rays)
    rate = rays.shape[0] / (time.perf_counter() - t0)
    budget_s = 100.0
    target = int(min(WIDTH * HEIGHT * SPP * 2.05, max(60_000, rate * budget_s / max(1, args.steps + args.warmup))))
    rows = sample_rows(target)
    images, streams, built_by = build_cpu_sample(rows, seed=1)
    rays = np.concatenate(streams)
    n = rays.shape[0]
    for _ in range(args.warmup):
        oracle.traverse_avx2(images, rays)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.traverse_avx2(images, rays)
    dt = time.perf_counter() - t0
    mrays = n * args.steps / dt / 1e6
    # beside it, the reference's own kernel SOURCE (Kernels.h compiled over oracle/ref_shim/opencl_c.h into oracle/_ref
    # in the build container), all host threads, on a prefix of the same rays; reported, not the arm's value: the AVX2
    # port above is the faster CPU implementation and therefore the stricter baseline
    ref_kernel = None
    if oracle.have_ref_kernel():
        m = min(n, max(50_000, int(rate * 3.0)))
        got = oracle.traverse_avx2(images, rays[:m])
        t1 = time.perf_counter()
        want = oracle.ref_kernel_traverse(images, rays[:m], threads=0)
        ref_kernel = {"value": round(m / (time.perf_counter() - t1) / 1e6, 3), "unit": "Mrays/s", "cores": cores, "kind": "reference",
                      "sample": f"first {m} rays of the step", "what": "reference traversal kernel source text run on the CPU through an "
                      "OpenCL C shim, one work-item at a time per thread", "port_results_bit_identical": bool(got.tobytes() == want.tobytes())}
    sample = (f"{len(rows)} of {HEIGHT} pixel rows (evenly strided) of the {WIDTH}x{HEIGHT}x{SPP}spp batch, primary + {BOUNCES} "
              f"bounces = {n} rays per step; scene images by the {built_by}")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(mrays, 3), "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(), "rays_per_step": n, "note": "reference CPU path is Embree 2.7 (binary-only, macOS/Windows); "
                   "timed here: the CPU restatement of the reference's own traversal kernel (oracle/racc_oracle.c), AVX2 node test (1 ray x 2 "
                   "boxes, SURVEY.md 8d), one ray at a time per thread, all host threads; bit-identical to the scalar checker"},
        "cpu_baseline": {"value": round(mrays, 3), "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(mrays, 3), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if ref_kernel is not None:
        line["reference_kernel_source"] = ref_kernel
    emit(line)


# --------------------------------------------------------------------------------------------
# our arm


def run_engine(args, rank: int, local_rank: int, world: int) -> None:
    import torch
    import torch.distributed as dist

    import rayaccel_b200 as rb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    host_binding = bind_near_gpu(local_rank) if world > 1 else "single rank: unbound"
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    rb.init(local_rank)
    stream = torch.cuda.current_stream()

    # ---- workload: scene replicated per rank, rays generated on the device (seed differs per rank)
    sf = rb.load_scene()
    t_build = time.perf_counter()
    scene = rb.create_scene(sf.vertices, sf.indices)
    env = rb.create_environment(sf.environment)
    t_build = time.perf_counter() - t_build
    cam = rb.Camera.for_scene(sf, WIDTH, HEIGHT)
    seed = 1 + rank
    streams = []  # (rays tensor, results tensor, count)
    n = WIDTH * HEIGHT * SPP
    rays = torch.empty(n * 8, dtype=torch.float32, device="cuda")
    rb.generate_primary(cam, WIDTH, HEIGHT, SPP, seed, rays.data_ptr(), stream=stream)
    alg_bytes = 0
    visit = []
    for depth in range(BOUNCES + 1):
        res = torch.empty(max(n, 1) * 4, dtype=torch.float32, device="cuda")
        cnt = torch.zeros(4, dtype=torch.int64, device="cuda")
        rb.trace_device(scene, env, [(rays.data_ptr(), res.data_ptr(), n)], stream=stream, counters_ptr=cnt.data_ptr(), detail=True)
        torch.cuda.synchronize()
        c = [int(x) for x in cnt.cpu().tolist()]
        # B_ray = 32 + 16 + 64*N_inner + 48*N_pair + 4*[hit] + 64*[miss]
        b = 48 * n + 64 * c[2] + 48 * c[3] + 4 * c[1] + 64 * (n - c[1])
        alg_bytes += b
        visit.append({"stream": "primary" if depth == 0 else f"bounce{depth}", "rays": n, "hit_rate": round(c[1] / max(n, 1), 4),
                      "inner_per_ray": round(c[2] / max(n, 1), 3), "pairs_per_ray": round(c[3] / max(n, 1), 3), "alg_bytes_per_ray": round(b / max(n, 1), 1)})
        streams.append((rays, res, n))
        if depth == BOUNCES:
            break
        nxt = torch.empty(max(c[1], 1) * 8, dtype=torch.float32, device="cuda")
        k = torch.zeros(1, dtype=torch.int32, device="cuda")
        rb.generate_bounce(scene, rays.data_ptr(), res.data_ptr(), n, seed + 1 + depth, nxt.data_ptr(), k.data_ptr(), stream=stream)
        torch.cuda.synchronize()
        rays, n = nxt, int(k.item())
    rays_per_step = sum(s[2] for s in streams)
    descs = [(r.data_ptr(), o.data_ptr(), c) for r, o, c in streams]
    frame_stats = torch.zeros(4, dtype=torch.int64, device="cuda")

    def step():
        """One pass of the hot path over the batch: ONE traversal launch over the four streams; with
        N>1 followed by the per-frame hit reduction (the only collective on the path)."""
        frame_stats.zero_()
        rb.trace_device(scene, env, descs, stream=stream, counters_ptr=frame_stats.data_ptr(), detail=False)
        if world > 1:
            dist.all_reduce(frame_stats)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = rb.launch_count()
    kernel_events = []
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        frame_stats.zero_()
        a.record(stream)
        rb.trace_device(scene, env, descs, stream=stream, counters_ptr=frame_stats.data_ptr(), detail=False)
        b.record(stream)
        if world > 1:
            dist.all_reduce(frame_stats)
        kernel_events.append((a, b))
    ev1.record(stream)
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    launches = rb.launch_count() - launches0
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kernel_events]))
    clocks = sampler.stop() if rank == 0 else {}
    hits_all_ranks = int(frame_stats[1].item())

    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    ms_per_step = total_ms_max / args.steps
    value = world * rays_per_step / (ms_per_step * 1e-3) / 1e6

    # ---- breakdown (untimed section): primary and secondary streams launched separately
    def time_launch(d, iters=5):
        best = 1e30
        for _ in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            rb.trace_device(scene, env, d, stream=stream)
            b.record(stream)
            b.synchronize()
            best = min(best, a.elapsed_time(b))
        return best
    ms_primary = time_launch(descs[:1])
    ms_secondary = time_launch(descs[1:])
    n_secondary = rays_per_step - streams[0][2]

    # ---- end to end through the C-ABI with pinned HOST buffers (H2D + trace + D2H timed)
    e2e_steps = max(1, min(args.steps, 5))
    host = []
    for r, o, c in streams:
        hr = torch.empty(c * 8, dtype=torch.float32).pin_memory()
        hr.copy_(r[: c * 8])
        ho = torch.empty(c * 4, dtype=torch.float32).pin_memory()
        host.append((hr, ho, c))
    hdescs = [(hr.data_ptr(), ho.data_ptr(), c) for hr, ho, c in host]
    for _ in range(2):
        rb.trace_host_ptrs(scene, env, hdescs, stream=stream)
        rb.sync(stream)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record(stream)
    for _ in range(e2e_steps):
        rb.trace_host_ptrs(scene, env, hdescs, stream=stream)
        rb.sync(stream)  # results are in host memory here: the call a user makes returns them
    e1.record(stream)
    barrier()
    e2e_wall_ms = (time.perf_counter() - w0) * 1e3
    e2e_ms = max(e0.elapsed_time(e1), 0.0)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item()) / e2e_steps
    e2e_value = world * rays_per_step / (e2e_ms * 1e-3) / 1e6
    # the D2H copy is real: results on the host equal the device-resident run
    e2e_ok = all(bool(torch.equal(ho.view(torch.int32), o[: c * 4].cpu().view(torch.int32))) for (hr, ho, c), (r, o, _) in zip(host, streams))

    # ---- beside the contract's numbers: the whole frame rendered on the device -- the reference's
    # path-tracing estimator with the shading in CUDA (csrc/pathtrace.cu), rays / hits / path state never leaving HBM.
    # Every rank renders SPP samples of all pixels (weak scaling), the framebuffers are summed by one NCCL all-reduce.
    device_render = None
    from rayaccel_b200 import sharding
    shading = fb = None
    render_error = ""
    try:  # rank-local part only: a rank that fails here must still reach the collectives below
        shading = rb.create_shading(sf.normals, sf.triangle_normals, sf.materials)
        fb = torch.zeros(WIDTH * HEIGHT * 4, dtype=torch.float32, device="cuda")
        rb.path_trace(scene, env, shading, cam, WIDTH, HEIGHT, SPP, sf.max_depth, seed=1, framebuffer_ptr=fb.data_ptr(),
                      sample_base=rank * SPP, stream=stream)  # untimed: pool growth, first launches
        rb.sync(stream)
    except Exception as e:  # noqa: BLE001 -- an extra, never allowed to take the contract's line down
        render_error = str(e)[:300]
    ok = torch.tensor([0 if render_error else 1], dtype=torch.int32, device="cuda")
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()):
        best_ms, waves = 1e30, None
        for rep in range(3):
            fb.zero_()
            barrier()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record(stream)
            _, waves = rb.path_trace(scene, env, shading, cam, WIDTH, HEIGHT, SPP, sf.max_depth, seed=1, framebuffer_ptr=fb.data_ptr(),
                                     sample_base=rank * SPP, stream=stream)
            if world > 1:
                with torch.cuda.stream(stream):
                    sharding.reduce_framebuffer(fb)
            r1.record(stream)
            r1.synchronize()
            t = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best_ms = min(best_ms, float(t.item()))
        n = torch.tensor([sum(waves)], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(n)
        device_render = {"value": round(int(n.item()) / best_ms / 1e3, 1), "unit": "Mrays/s", "ms": round(best_ms, 3), "rays": int(n.item()),
                         "rays_per_depth_rank0": waves, "spp_per_gpu": SPP, "max_depth": int(sf.max_depth),
                         "mean_radiance": round(float(fb.view(-1, 4)[:, :3].double().mean().item()) / (SPP * world), 5),
                         "what": "racc_cuda_path_trace: camera rays, traversal, material sampling, compaction and framebuffer accumulation on "
                                 "the device, one host round trip (the next wave's size) per bounce; every rank renders spp_per_gpu samples of "
                                 "all pixels, framebuffers summed by one NCCL all-reduce inside the timed region; bit-identical to "
                                 "oracle_path_trace (tests/test_gpu_render.py); best of 3, max over ranks"}
    else:
        device_render = {"error": render_error or "another rank failed"}
    if shading is not None:
        shading.destroy()
    del fb

    # ---- CPU baseline on a bounded sample of the SAME rays (rank 0, N=1 only) + parity spot check
    cpu_baseline = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle
        nodes, pairs, remap = scene.download()
        images = oracle.SceneImages(nodes, pairs, remap, sf.environment)
        frac = min(1.0, args.cpu_sample_rays / rays_per_step)
        parts, gpu_parts = [], []
        for r, o, c in streams:
            m = max(1, int(c * frac))
            parts.append(r[: m * 8].cpu().numpy().view(oracle.RAY_DTYPE))
            gpu_parts.append(o[: m * 4].cpu().numpy().view(np.uint32).reshape(-1, 4))
        sample = np.concatenate(parts)
        oracle.traverse_avx2(images, sample[: 20000])  # warm the library and the caches
        t0 = time.perf_counter()
        want = oracle.traverse_avx2(images, sample)
        dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        checker = oracle.traverse(images, sample[: 2_000_000])  # the scalar parity checker on a prefix, for the record
        dt_checker = time.perf_counter() - t0
        assert np.array_equal(checker.view(np.uint32), want[: checker.shape[0]].view(np.uint32)), "AVX2 baseline differs from the scalar checker"
        cpu_baseline = {"value": round(sample.shape[0] / dt / 1e6, 3), "unit": "Mrays/s", "cores": os.cpu_count() or 1, "kind": "port",
                        "sample": f"first {frac * 100:.1f}% of each of the {len(streams)} ray streams of the batch ({sample.shape[0]} rays), "
                                  f"C restatement of the reference kernel with an AVX2 node test (1 ray x 2 boxes), all host threads, {dt:.1f} s; "
                                  f"the scalar parity checker runs at {min(sample.shape[0], 2_000_000) / dt_checker / 1e6:.1f} Mrays/s on the same cores"}
        parity = bool(np.array_equal(np.concatenate(gpu_parts), want.view(np.uint32).reshape(-1, 4)))

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        traffic, traffic_src, limiters = ncu_traffic()
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(), "rays_per_step_per_gpu": rays_per_step, "streams": visit,
                       "l2": "inputs larger than L2 (rays+results 0.8 GB per step vs 126 MB); the 3.2 MB scene is L2-resident by nature",
                       "parallelism": f"ray-sharded x{world}, scene replicated, per-frame hit all-reduce" if world > 1 else "single GPU",
                       "scene_build_s": round(t_build, 3)},
            "breakdown": {"primary_mrays": round(streams[0][2] / ms_primary / 1e3, 1), "secondary_mrays": round(n_secondary / ms_secondary / 1e3, 1),
                          "primary_rays": streams[0][2], "secondary_rays": n_secondary, "note": "per GPU, separate launches, best of 5"},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": traffic, "kernel": "tracePackedKernel", "kernel_ms": round(kernel_ms, 4),
                         "algorithmic_bytes_per_launch": alg_bytes, "compulsory_bytes_per_launch": 48 * rays_per_step,
                         "peak_source": peak_src, "traffic_source": traffic_src, "ncu_limiters": limiters,
                         "note": "algorithmic bytes = sum over rays of 32+16+64*N_inner+48*N_pair+4*[hit]+64*[miss] (SURVEY.md 8d), visits counted "
                                 "by the kernel itself. frac > 1 is expected here: battlefield's 3.2 MB scene is L1/L2-resident, so of the algorithmic "
                                 "bytes only the compulsory 48 B/ray (rays in, results out) reach DRAM -- `traffic` (ncu dram read+write per launch) "
                                 "equals compulsory_bytes_per_launch, i.e. no wasted re-reads. What bounds the kernel is the L1 data pipe (ncu_limiters, from "
                                 "the committed ncu --set full capture of this command: l1tex data-pipe wavefronts ~80 % of peak, issue slots ~66 %), "
                                 "not HBM (profiles/r01_ncu_bench_launch_packed.txt)"},
            "e2e": {"value": round(e2e_value, 2), "unit": "Mrays/s", "h2d_bytes_per_step": 32 * rays_per_step, "d2h_bytes_per_step": 16 * rays_per_step,
                    "ms_per_step": round(e2e_ms, 3), "wall_ms_per_step": round(e2e_wall_ms / e2e_steps, 3), "steps": e2e_steps,
                    "results_match_device_run": e2e_ok, "host_binding": host_binding,
                    "path": "racc_cuda_trace with RACC_CUDA_STREAM_HOST descriptors, pinned host memory",
                    "pcie_gbs": {"h2d": round(32 * world * rays_per_step / (e2e_ms * 1e-3) / 1e9 / world, 1), "d2h": round(16 * world * rays_per_step / (e2e_ms * 1e-3) / 1e9 / world, 1),
                                 "note": "per GPU; plain pinned copies on this pool reach 52.6 + 26.3 GB/s with both directions busy (profiles/r01_pcie_copy_bandwidth.txt)"}},
            "gpu_launches": int(launches), "clocks": clocks, "frame_hits_all_ranks": hits_all_ranks,
        }
        if device_render is not None:
            line["device_render"] = device_render
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
            line["parity_sample_bit_exact"] = parity
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["engine", "reference"], default="engine")
    ap.add_argument("--cpu-sample-rays", type=int, default=20_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    _claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit(f"bench.py: --gpus {args.gpus} needs torchrun (python -m torch.distributed.run --nproc-per-node {args.gpus} ...)")
    run_engine(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
