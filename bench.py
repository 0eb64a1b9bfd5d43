#!/usr/bin/env python
"""Headline benchmark: Mrays/s (primary + incoherent secondary) on battlefield.bin.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): path-tracer wavefront on battlefield.bin, 1920x1080, 4 spp,
depth 3 (the scene header's maxDepth: primary rays + 3 diffuse bounces, Renderer/
PathTracingRenderer.cpp:100,120). The batch is generated on the HOST by bench_rays.py -- the same
module, seed and code path in both arms, so the engine arm and `--impl reference` trace the same ray
bytes (`config.ray_digest`). A "step" is ONE pass of the hot path over that batch: all four ray
streams (~17 M rays, 0.8 GB of rays+results, larger than L2) traced by one launch of the
persistent traversal kernel, device-resident, followed by the per-frame hit reduction
(racc_cuda_frame_reduce: NCCL, called from the engine library). `value` = rays/s over all ranks with
inputs in HBM; `e2e` = the same batch through the reference-facing C-ABI with pinned HOST buffers
(H2D + trace + D2H inside the timed region). Multi-GPU is weak scaling: the scene is replicated,
every rank traces its own batch (different jitter seed).

Beside the contract's keys the line carries the records VERDICT r01 asked for: `roofline` is the L1
gather roofline that actually bounds the kernel on this L2-resident scene (the algorithmic-HBM figure
stays inside it, labelled cache-served); `c5` = BASELINE configs[4] (10 M-triangle soup, the one
HBM-bound config) with its own HBM roofline and an oracle-checked parity sample; `c3` =
configs[2] (64 spp, 8 bounces) on the device renderer; `c4` = configs[3] (3840x2160x16 spp) as a
STRONG-scaling record (samples split over the ranks); `e2e_one_process` (N>1) = the HOST-stream path
with ONE process driving all N GPUs through the library's own multi-device staging.

The oracle (oracle/) is used here only as the checker / CPU baseline: `cpu_baseline` times it on a
bounded sample of the same rays on rank 0 at N=1, and `--impl reference` times it alone (the
reference's own CPU path is Embree 2.7, binary-only for macOS/Windows, so the port stands in).
"""
from __future__ import annotations

import argparse
import hashlib
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

import bench_rays  # noqa: E402  (neutral: numpy only)

METRIC = "Mrays/s (primary + incoherent secondary) on battlefield.bin"
WIDTH, HEIGHT, SPP, BOUNCES = 1920, 1080, 4, 3
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
L1_WAVEFRONTS_PER_CLK_PER_SM = 1.0  # measured: profiles/r01_l1_wavefront_microbench.md, r02_call1_open_questions.md


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


def note(msg: str) -> None:
    sys.stderr.write(f"[bench {time.strftime('%H:%M:%S')}] {msg}\n")
    sys.stderr.flush()


def numa_report(local_rank: int) -> dict:
    """What the box says about NUMA placement: the GPU's node from sysfs, the nodes the kernel
    exposes, and where this process may run. On this pool's VMs sysfs reports -1 for every GPU and a single node, i.e.
    there is nothing to bind to; the evidence goes into the JSON line instead of a guess."""
    out = {}
    try:
        import torch
        props = torch.cuda.get_device_properties(local_rank)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            out["gpu_numa_node"] = int(f.read().strip())
    except Exception as e:  # noqa: BLE001
        out["gpu_numa_node"] = f"unreadable ({type(e).__name__})"
    try:
        out["host_numa_nodes"] = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
    except Exception:  # noqa: BLE001
        out["host_numa_nodes"] = "unreadable"
    out["cpus_allowed"] = len(os.sched_getaffinity(0))
    node = out.get("gpu_numa_node")
    if isinstance(node, int) and node >= 0 and isinstance(out["host_numa_nodes"], list) and len(out["host_numa_nodes"]) > 1:
        try:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                spec = f.read().strip()
            cpus = set()
            for part in spec.split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            allowed = os.sched_getaffinity(0) & cpus
            if allowed:
                os.sched_setaffinity(0, allowed)
                out["bound_to"] = f"node {node}, {len(allowed)} CPUs"
        except Exception as e:  # noqa: BLE001
            out["bound_to"] = f"failed ({type(e).__name__})"
    else:
        out["bound_to"] = "nothing to bind to: the VM exposes one NUMA node / no GPU affinity"
    return out


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def kernel_source_sha() -> str:
    """Identity of the traversal kernel's source: a committed ncu capture is only quoted while it still describes this code."""
    h = hashlib.sha256()
    for name in ("traverse_packed.cu", "traverse_packed.cuh", "traverse_common.cuh", "engine.h"):
        with open(os.path.join(ROOT, "rayaccel_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def ncu_capture(name: str):
    """DRAM bytes and L1 data-pipe wavefronts per launch + limiter percentages from a committed `ncu --set full` capture
    (profiles/<name>.json, written by tools/ncu_traffic.py). Returns (traffic, source, limiters, wavefronts); traffic and
    wavefronts are None -- with the reason as the source -- when the file is absent or was captured from different kernel
    source (no blind replay)."""
    try:
        with open(os.path.join(ROOT, "profiles", name + ".json")) as f:
            t = json.load(f)
    except Exception:
        return None, f"no committed capture (profiles/{name}.json)", None, None
    if t.get("kernel_source_sha16") != kernel_source_sha():
        return None, (f"profiles/{name}.json is stale: captured from kernel source {t.get('kernel_source_sha16')}, this build is "
                      f"{kernel_source_sha()}; re-run tools/profile.sh"), None, None
    return float(t["dram_bytes_per_launch"]), t.get("source"), t.get("limiters"), t.get("l1_wavefronts_per_launch")


def workload_config(rays_per_step: int, ray_digest: str) -> dict:
    """`config` of the JSON line, the same dict in both arms: what is traced, how many rays, a digest of their bytes, and the
    cache policy between timed iterations (the inputs of one step are far larger than L2, nothing is flushed)."""
    return {"workload": workload_name(), "rays_per_step": rays_per_step, "ray_digest": ray_digest,
            "l2": "inputs larger than L2: 0.8 GB of rays + results per step vs 126 MB; the 3.2 MB scene is cache-resident by nature"}


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
# reference arm: the CPU port, all host threads; never touches the engine


def reference_images(sf):
    """Scene images for the CPU path: by the unmodified reference builder when oracle/_ref travelled, else by the engine's
    host-only builder (structurally identical, tests/test_scene_build.py)."""
    import oracle
    if oracle.have_ref():
        images = oracle.ref_build_scene(sf.vertices, sf.indices)  # the unmodified reference builder
        images.env = sf.environment
        return images, "unmodified reference builder (oracle/_ref)"
    from rayaccel_b200 import HostImages
    h = HostImages(sf.vertices, sf.indices)
    return oracle.SceneImages(h.nodes, h.pairs, h.remap, sf.environment), "engine host builder (oracle/_ref absent)"


def run_reference(args, rank: int, world: int) -> None:
    if rank != 0:
        return  # rank 0 alone runs the reference arm
    import oracle
    from rayaccel_b200 import scene_io  # file loader only (no engine call)
    try:
        oracle.build(ref=os.path.isdir("/root/reference/RayAccelerator"))
    except Exception:
        pass  # prebuilt liboracle.so / _ref travel with the snapshot
    cores = os.cpu_count() or 1
    sf = scene_io.load_scene()
    images, built_by = reference_images(sf)
    # the engine arm's batch, byte for byte: same generator, same seed; the hits that drive the bounces come from the CPU path
    t0 = time.perf_counter()
    streams, _ = bench_rays.wavefront(sf, WIDTH, HEIGHT, SPP, BOUNCES, 1, lambda r: oracle.traverse_avx2(images, r))
    gen_s = time.perf_counter() - t0
    rays = np.concatenate(streams)
    n_full = rays.shape[0]
    ray_digest = bench_rays.digest(streams)
    # one pass over the full batch is ~0.4-1 s on a 16-32 core box; if the box is slower than that, time a strided sample
    t0 = time.perf_counter()
    oracle.traverse_avx2(images, rays[: min(n_full, 2_000_000)])
    rate = min(n_full, 2_000_000) / (time.perf_counter() - t0)
    budget_s = 100.0
    per_step = max(200_000, int(rate * budget_s / max(1, args.steps + args.warmup)))
    if per_step >= n_full:
        sample, what = rays, f"the whole batch ({n_full} rays per step)"
    else:
        stride = -(-n_full // per_step)
        sample = np.ascontiguousarray(rays[::stride])
        what = f"every {stride}th ray of the batch ({sample.shape[0]} of {n_full} rays per step)"
    n = sample.shape[0]
    for _ in range(args.warmup):
        oracle.traverse_avx2(images, sample)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.traverse_avx2(images, sample)
    dt = time.perf_counter() - t0
    mrays = n * args.steps / dt / 1e6
    # beside it, the reference's own kernel SOURCE (Kernels.h compiled over oracle/ref_shim/opencl_c.h into oracle/_ref
    # in the build container), all host threads, on a prefix of the same rays; reported, not the arm's value: the AVX2
    # port above is the faster CPU implementation and therefore the stricter baseline
    ref_kernel = None
    if oracle.have_ref_kernel():
        m = min(n, max(50_000, int(rate * 3.0)))
        got = oracle.traverse_avx2(images, sample[:m])
        t1 = time.perf_counter()
        want = oracle.ref_kernel_traverse(images, sample[:m], threads=0)
        ref_kernel = {"value": round(m / (time.perf_counter() - t1) / 1e6, 3), "unit": "Mrays/s", "cores": cores, "kind": "reference",
                      "sample": f"first {m} rays of the step", "what": "reference traversal kernel source text run on the CPU through an "
                      "OpenCL C shim, one work-item at a time per thread", "port_results_bit_identical": bool(got.tobytes() == want.tobytes())}
    # ... and the reference's CPU query path, executeRayQueryCPU (Scene.cpp:374-484), from its own source over the stand-in for
    # its binary-only Embree 2.7 (oracle/ref_shim/mini_embree.cpp: an unoptimised 8-wide packet traversal, so this rate says
    # nothing about Embree's speed; it is here for the agreement check)
    ref_cpu_path = None
    if oracle.have_ref_cpu_query():
        m = min(n, 2_000_000)
        t1 = time.perf_counter()
        got = oracle.ref_cpu_query(sf.vertices, sf.indices, sf.environment, sample[:m], threads=0)
        dt1 = time.perf_counter() - t1
        want = oracle.traverse_avx2(images, sample[:m])
        hit_a, hit_b = got["triangle"] != oracle.INVALID, want["triangle"] != oracle.INVALID
        same = hit_a & hit_b & (got["triangle"] == want["triangle"])
        ref_cpu_path = {"value": round(m / dt1 / 1e6, 3), "unit": "Mrays/s", "cores": cores, "kind": "reference",
                        "sample": f"first {m} rays of the step (scene set-up included in the time)",
                        "what": "racc_internal::executeRayQueryCPU run from its own source; rtcIntersect8 served by oracle/ref_shim/mini_embree.cpp",
                        "hit_miss_disagreements_with_port": int((hit_a != hit_b).sum()),
                        "prim_ids_differing_from_port": int((hit_a & hit_b & ~same).sum()),
                        "max_rel_dt": float(np.max(np.abs(got["a"][same] - want["a"][same]) / np.abs(want["a"][same]))) if same.any() else 0.0}
    line = {
        "impl": "reference", "metric": METRIC, "value": round(mrays, 3), "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(n_full, ray_digest),
        "cpu_baseline": {"value": round(mrays, 3), "unit": "Mrays/s", "cores": cores, "kind": "port",
                         "sample": f"{what}; scene images by the {built_by}; batch generated in {gen_s:.1f} s (untimed)"},
        "e2e": {"value": round(mrays, 3), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference CPU path is Embree 2.7 (binary-only, macOS/Windows); timed here: the CPU restatement of the reference's own "
                "traversal kernel (oracle/racc_oracle.c), AVX2 node test (1 ray x 2 boxes, SURVEY.md 8d), one ray at a time per thread, all "
                "host threads; bit-identical to the scalar checker",
    }
    if ref_kernel is not None:
        line["reference_kernel_source"] = ref_kernel
    if ref_cpu_path is not None:
        line["reference_cpu_query_path"] = ref_cpu_path
    emit(line)


# --------------------------------------------------------------------------------------------
# our arm


def l1_gather_instructions(c, n, hits):
    """Per-lane gather instructions of a traced stream from the kernel's own counters: 2 x LDG.256 per inner node and per
    pair, one LDL/STL per stack entry pushed and popped, the ray (2 x LDG.128), the result (1 x STG.128), the remap word of
    a hit, the two texel-pair rows of a miss."""
    return 2 * c[2] + 2 * c[3] + 2 * c[4] + 3 * n + hits + 2 * (n - hits)


def run_engine(args, rank: int, local_rank: int, world: int) -> None:
    import torch
    import torch.distributed as dist

    import rayaccel_b200 as rb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = numa_report(local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    rb.init(local_rank)
    stream = torch.cuda.current_stream()
    if world > 1:
        # the engine's own communicator for the per-frame hit reduction: rank 0's NCCL id goes round through torch's store
        box = [rb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        rb.comm_init_rank(box[0], rank, world)

    def gmax(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gsum(x: int) -> int:
        t = torch.tensor([x], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(t)
        return int(t.item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- workload: scene replicated per rank; the batch comes from the host generator (seed differs per rank), the hits
    # that drive its bounces from the engine
    sf = rb.load_scene()
    t_build = time.perf_counter()
    scene = rb.create_scene(sf.vertices, sf.indices)
    env = rb.create_environment(sf.environment)
    t_build = time.perf_counter() - t_build
    seed = 1 + rank
    dev_streams = []   # (rays tensor, results tensor, count, gather instructions)
    visit = []
    totals = {"alg_bytes": 0, "gathers": 0}

    def trace_wave(rays_np):
        n = rays_np.shape[0]
        d_rays = torch.from_numpy(rays_np.view(np.float32).reshape(-1)).cuda()
        d_res = torch.empty(max(n, 1) * 4, dtype=torch.float32, device="cuda")
        cnt = torch.zeros(8, dtype=torch.int64, device="cuda")
        rb.trace_device(scene, env, [(d_rays.data_ptr(), d_res.data_ptr(), n)], stream=stream, counters_ptr=cnt.data_ptr(), detail=True)
        torch.cuda.synchronize()
        c = [int(x) for x in cnt.cpu().tolist()]
        # B_ray = 32 + 16 + 64*N_inner + 48*N_pair + 4*[hit] + 64*[miss]
        b = 48 * n + 64 * c[2] + 48 * c[3] + 4 * c[1] + 64 * (n - c[1])
        g = l1_gather_instructions(c, n, c[1])
        totals["alg_bytes"] += b
        totals["gathers"] += g
        depth = len(dev_streams)
        visit.append({"stream": "primary" if depth == 0 else f"bounce{depth}", "rays": n, "hit_rate": round(c[1] / max(n, 1), 4),
                      "inner_per_ray": round(c[2] / max(n, 1), 3), "pairs_per_ray": round(c[3] / max(n, 1), 3),
                      "stack_pushes_per_ray": round(c[4] / max(n, 1), 3), "leaves_per_ray": round(c[5] / max(n, 1), 3),
                      "alg_bytes_per_ray": round(b / max(n, 1), 1), "gather_instructions_per_ray": round(g / max(n, 1), 2)})
        dev_streams.append((d_rays, d_res, n, g))
        return d_res[: n * 4].cpu().numpy().view(bench_rays.RESULT_DTYPE)

    t_gen = time.perf_counter()
    host_streams, host_results = bench_rays.wavefront(sf, WIDTH, HEIGHT, SPP, BOUNCES, seed, trace_wave)
    t_gen = time.perf_counter() - t_gen
    alg_bytes, gathers = totals["alg_bytes"], totals["gathers"]
    ray_digest = bench_rays.digest(host_streams)
    rays_per_step = sum(s[2] for s in dev_streams)
    note(f"rank {rank}: batch of {rays_per_step} rays generated in {t_gen:.1f} s, digest {ray_digest}")
    descs = [(r.data_ptr(), o.data_ptr(), c) for r, o, c, _ in dev_streams]

    def step():
        """One pass of the hot path over the batch: ONE traversal launch over the four streams, then the per-frame hit
        reduction -- the only collective on the path, called from the engine library (NCCL all-reduce over the ranks at N>1)."""
        rb.trace_device(scene, env, descs, stream=stream)
        rb.frame_reduce(stream, wait=False)

    rb.frame_reduce(stream)  # the accounting launches above are not part of any frame
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
        a.record(stream)
        rb.trace_device(scene, env, descs, stream=stream)
        b.record(stream)
        rb.frame_reduce(stream, wait=False)
        kernel_events.append((a, b))
    ev1.record(stream)
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    launches = rb.launch_count() - launches0
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kernel_events]))
    clocks = sampler.stop() if rank == 0 else {}
    total_ms_max = gmax(total_ms)
    ms_per_step = total_ms_max / args.steps
    rays_all_ranks = gsum(rays_per_step)
    value = rays_all_ranks / (ms_per_step * 1e-3) / 1e6

    # the hit reduction is real: one more frame, reduced with a wait, equals the sum of the ranks' own counts
    rb.trace_device(scene, env, descs, stream=stream)
    frame = rb.frame_reduce(stream)
    hits_mine = sum(int((r["triangle"] != bench_rays.INVALID).sum()) for r in host_results)
    frame_ok = frame["rays"] == rays_all_ranks and frame["hits"] == gsum(hits_mine)

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
    n_secondary = rays_per_step - dev_streams[0][2]
    g_primary, g_secondary = dev_streams[0][3], sum(s[3] for s in dev_streams[1:])
    # ... and once more on the quantised node image (variant 4, opt-in), results into scratch so the exact ones stay
    scratch = [torch.empty_like(o) for _, o, _, _ in dev_streams]
    qdescs = [(r.data_ptr(), q.data_ptr(), c) for (r, _, c, _), q in zip(dev_streams, scratch)]
    rb.set_tuning(variant=4)
    try:
        ms_primary_q = time_launch(qdescs[:1])
        ms_secondary_q = time_launch(qdescs[1:])
    finally:
        rb.set_tuning(variant=3)
    quant_ids_differing = sum(int((q.view(-1, 4).view(torch.int32)[:c, 0] != o.view(-1, 4).view(torch.int32)[:c, 0]).sum())
                              for (_, o, c, _), q in zip(dev_streams, scratch))
    del scratch
    rb.frame_reduce(stream)

    # ---- end to end through the C-ABI with pinned HOST buffers (H2D + trace + D2H timed)
    e2e_steps = max(1, min(args.steps, 5))
    host = []
    for (r, o, c, _), h in zip(dev_streams, host_streams):
        hr = torch.from_numpy(h.view(np.float32).reshape(-1)).pin_memory()
        ho = torch.empty(c * 4, dtype=torch.float32).pin_memory()
        host.append((hr, ho, c))
    hdescs = [(hr.data_ptr(), ho.data_ptr(), c) for hr, ho, c in host]

    # what plain pinned copies reach on this box with every rank busy: the ceiling e2e is to be read against
    def pcie_ceiling():
        n = 1 << 28  # 256 MiB in, 128 MiB out: 2:1 like the path (32 B in, 16 B out per ray)
        hin = torch.empty(n, dtype=torch.uint8).pin_memory()
        hout = torch.empty(n // 2, dtype=torch.uint8).pin_memory()
        din = torch.empty(n, dtype=torch.uint8, device="cuda")
        dout = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        best = 1e30
        for rep in range(4):
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            s1.wait_stream(stream)
            s2.wait_stream(stream)
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)
            stream.wait_stream(s1)
            stream.wait_stream(s2)
            b.record(stream)
            b.synchronize()
            t = gmax(a.elapsed_time(b))
            if rep:
                best = min(best, t)
        return {"h2d_gbs_per_gpu": round(n / best / 1e6, 1), "d2h_gbs_per_gpu": round(n / 2 / best / 1e6, 1),
                "mrays_ceiling_all_ranks": round(world * (n / 32) / best / 1e3, 1),
                "what": "256 MiB H2D + 128 MiB D2H of pinned memory on two streams, all ranks at once, max over ranks, best of 3"}
    ceiling = pcie_ceiling()

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
    e2e_ms = gmax(max(e0.elapsed_time(e1), 0.0)) / e2e_steps
    e2e_value = rays_all_ranks / (e2e_ms * 1e-3) / 1e6
    # the D2H copy is real: results on the host equal the device-resident run
    e2e_ok = all(bool(torch.equal(ho.view(torch.int32), o[: c * 4].cpu().view(torch.int32))) for (hr, ho, c), (r, o, _, _) in zip(host, dev_streams))
    rb.frame_reduce(stream)

    # ---- N>1: the same HOST batch with ONE process driving every GPU of the box through the library's own multi-device
    # staging (racc_cuda_init over all devices: scene replicated peer to peer, chunks dealt over the devices). The other
    # ranks hold still on the store (no NCCL barrier: that would spin a kernel on their GPUs).
    one_process = None
    if world > 1:
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            try:
                if torch.cuda.device_count() < world:
                    raise RuntimeError(f"only {torch.cuda.device_count()} devices visible to rank 0")
                rb.init(list(range(world)))
                scene_all = rb.create_scene(sf.vertices, sf.indices)
                env_all = rb.create_environment(sf.environment)
                outs = [torch.empty(c * 4, dtype=torch.float32).pin_memory() for _, _, c in host]
                d_all = [(hr.data_ptr(), o.data_ptr(), c) for (hr, _, c), o in zip(host, outs)]
                for _ in range(2):
                    rb.trace_host_ptrs(scene_all, env_all, d_all, stream=stream)
                    rb.sync(stream)
                rb.frame_reduce(stream)
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    rb.trace_host_ptrs(scene_all, env_all, d_all, stream=stream)
                    rb.sync(stream)
                dt = (time.perf_counter() - t0) / e2e_steps
                tot = rb.frame_reduce(stream)  # NCCL all-reduce over the devices, inside the library, one process
                same = all(bool(torch.equal(o.view(torch.int32), ho.view(torch.int32))) for o, (_, ho, _) in zip(outs, host))
                one_process = {"value": round(rays_per_step / dt / 1e6, 2), "unit": "Mrays/s", "devices": world, "ms_per_step": round(dt * 1e3, 3),
                               "results_match_device_run": same, "frame_rays_reduced_in_library": tot["rays"] == rays_per_step * e2e_steps,
                               "what": "ONE batch (rank 0's), one process, racc_cuda_trace over a device set of all GPUs: strong scaling of the "
                                       "HOST-stream path; wall clock around the calls"}
                env_all.destroy()
                scene_all.destroy()
                rb.thread_release()
            except Exception as e:  # noqa: BLE001 -- an extra, never allowed to take the contract's line down
                one_process = {"error": str(e)[:300]}
            rb.init(local_rank)
            store.set("racc_one_process_done", "1")
        else:
            try:
                import datetime
                store.wait(["racc_one_process_done"], datetime.timedelta(seconds=240))
            except Exception as e:  # noqa: BLE001 -- rank 0's extra took too long: go on, the barrier below still meets it
                note(f"rank {rank}: still waiting for rank 0's one-process measurement ({type(e).__name__})")
        barrier()

    # ---- beside the contract's numbers: frames rendered on the device -- the reference's
    # path-tracing estimator with the shading in CUDA (csrc/pathtrace.cu), rays / hits / path state never leaving HBM.
    from rayaccel_b200 import sharding
    shading = None
    render_error = ""
    try:  # rank-local part only: a rank that fails here must still reach the collectives below
        shading = rb.create_shading(sf.normals, sf.triangle_normals, sf.materials)
    except Exception as e:  # noqa: BLE001
        render_error = str(e)[:300]
    ok = gsum(0 if render_error else 1) == world

    def render(width, height, spp_total, depth, strong: bool, what: str):
        """spp_total samples of every pixel. strong: the samples are SPLIT over the ranks (total work fixed); else every rank
        renders spp_total of its own. Framebuffers summed by one NCCL all-reduce inside the timed region."""
        camr = rb.Camera.for_scene(sf, width, height)
        if strong:
            base, spp = sharding.sample_range(spp_total, rank, world)  # (first sample, count)
        else:
            spp, base = spp_total, rank * spp_total
        fb = torch.zeros(width * height * 4, dtype=torch.float32, device="cuda")
        if spp:
            rb.path_trace(scene, env, shading, camr, width, height, spp, depth, seed=1, framebuffer_ptr=fb.data_ptr(), sample_base=base, stream=stream)
        rb.sync(stream)  # untimed: pool growth, first launches
        best_ms, best_render_ms, waves = 1e30, 0.0, [0] * (depth + 1)
        for rep in range(3):
            fb.zero_()
            barrier()
            r0, rm, r1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            r0.record(stream)
            if spp:
                _, waves = rb.path_trace(scene, env, shading, camr, width, height, spp, depth, seed=1, framebuffer_ptr=fb.data_ptr(),
                                         sample_base=base, stream=stream)
            rm.record(stream)
            if world > 1:
                with torch.cuda.stream(stream):
                    sharding.reduce_framebuffer(fb)
            r1.record(stream)
            r1.synchronize()
            t_all, t_render = gmax(r0.elapsed_time(r1)), gmax(r0.elapsed_time(rm))
            if t_all < best_ms:
                best_ms, best_render_ms = t_all, t_render
        n_all = gsum(sum(waves))
        n_secondary_all = gsum(sum(waves[1:]))
        rec = {"value": round(n_all / best_ms / 1e3, 1), "unit": "Mrays/s", "ms": round(best_ms, 3), "render_ms_slowest_rank": round(best_render_ms, 3),
               "framebuffer_allreduce_bytes": width * height * 16 if world > 1 else 0, "rays": n_all,
               "secondary_rays": n_secondary_all, "rays_per_depth_rank0": waves, "scaling": "strong" if strong else "weak",
               "spp_this_rank": spp, "max_depth": int(depth),
               "mean_radiance": round(float(fb.view(-1, 4)[:, :3].double().mean().item()) / max(1, spp_total * (1 if strong else world)), 5),
               "what": what}
        del fb
        return rec

    device_render = c3 = c4 = None
    if ok:
        device_render = render(WIDTH, HEIGHT, SPP, int(sf.max_depth), False,
                               "racc_cuda_path_trace: camera rays, traversal, material sampling, compaction and framebuffer accumulation on the "
                               "device, no host round trip (wave sizes stay on the device); every rank renders 4 spp of all pixels, framebuffers "
                               "summed by one NCCL all-reduce inside the timed region; bit-identical to oracle_path_trace "
                               "(tests/test_gpu_render.py); best of 3, max over ranks")
        c3 = render(WIDTH, HEIGHT, 64, 8, False,
                    "BASELINE configs[2]: 1920x1080, 64 spp, 8 bounces, per GPU, on the device renderer; `secondary_rays` are the incoherent "
                    "stress (every bounce >= 1)")
        c3["secondary_share"] = round(c3["secondary_rays"] / max(1, c3["rays"]), 3)
        c4 = render(3840, 2160, 16, int(sf.max_depth), True,
                    "BASELINE configs[3]: 3840x2160, 16 spp, the 16 samples of every pixel SPLIT over the ranks (sharding.sample_range), "
                    "framebuffers summed by one NCCL all-reduce: strong scaling, value = all rays / max-over-ranks time")
    else:
        device_render = {"error": render_error or "another rank failed"}
    if shading is not None:
        shading.destroy()

    # ---- BASELINE configs[4]: 10 M-triangle soup, 100 M uniform random rays (split over the ranks) -- the HBM-bound config
    c5 = None
    c5_error = ""
    try:
        c5 = run_c5(args, rb, torch, rank, world, stream, gmax, gsum, barrier)
    except Exception as e:  # noqa: BLE001
        c5_error = str(e)[:300]
        note(f"rank {rank}: c5 failed: {c5_error}")
    if c5_error:
        c5 = {"error": c5_error}

    # ---- CPU baseline on a bounded sample of the SAME rays (rank 0, N=1 only) + parity spot check
    cpu_baseline = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle
        nodes, pairs, remap = scene.download()
        images = oracle.SceneImages(nodes, pairs, remap, sf.environment)
        frac = min(1.0, args.cpu_sample_rays / rays_per_step)
        parts, gpu_parts = [], []
        for h, (r, o, c, _) in zip(host_streams, dev_streams):
            m = max(1, int(c * frac))
            parts.append(h[:m])
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
                        "sample": f"first {frac * 100:.1f}% of each of the {len(dev_streams)} ray streams of the batch ({sample.shape[0]} rays), "
                                  f"C restatement of the reference kernel with an AVX2 node test (1 ray x 2 boxes), all host threads, {dt:.1f} s; "
                                  f"the scalar parity checker runs at {min(sample.shape[0], 2_000_000) / dt_checker / 1e6:.1f} Mrays/s on the same cores"}
        parity = bool(np.array_equal(np.concatenate(gpu_parts), want.view(np.uint32).reshape(-1, 4)))

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        achieved_hbm = alg_bytes / (kernel_ms * 1e-3) / 1e9
        traffic, traffic_src, limiters, wavefronts = ncu_capture("ncu_bench_traffic")
        sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
        sm_mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
        l1_peak = sm_count * sm_mhz * 1e6 * L1_WAVEFRONTS_PER_CLK_PER_SM / 1e9   # G wavefronts/s
        l1_algorithmic = gathers / (kernel_ms * 1e-3) / 1e9
        # achieved: the data-pipe wavefronts ncu counted for this very launch (same kernel source, same batch) over the time
        # measured live in this run; without a matching capture, the algorithmic count (which ignores that lanes share sectors)
        l1_achieved = wavefronts / (kernel_ms * 1e-3) / 1e9 if wavefronts else l1_algorithmic

        def l1(g, ms):
            return round(g / (ms * 1e-3) / 1e9 / l1_peak, 4)
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": workload_config(rays_per_step, ray_digest),
            "workload_detail": {"streams": visit, "rays_per_step_all_ranks": rays_all_ranks,
                                "l2": "inputs larger than L2 (rays+results 0.8 GB per step vs 126 MB); the 3.2 MB scene is L2-resident by nature",
                                "parallelism": (f"ray-sharded x{world}, scene replicated, per-frame hit reduction in the engine library (NCCL)"
                                                if world > 1 else "single GPU"),
                                "scene_build_s": round(t_build, 3), "batch_generation_s": round(t_gen, 1),
                                "generator": "bench_rays.py on the host (numpy), the same module and seed as --impl reference: equal ray_digest = equal ray bytes"},
            "breakdown": {"primary_mrays": round(dev_streams[0][2] / ms_primary / 1e3, 1), "secondary_mrays": round(n_secondary / ms_secondary / 1e3, 1),
                          "primary_rays": dev_streams[0][2], "secondary_rays": n_secondary,
                          "primary_l1_frac": l1(g_primary, ms_primary), "secondary_l1_frac": l1(g_secondary, ms_secondary),
                          "quantised_nodes": {"primary_mrays": round(dev_streams[0][2] / ms_primary_q / 1e3, 1),
                                              "secondary_mrays": round(n_secondary / ms_secondary_q / 1e3, 1), "ids_differing_from_exact": quant_ids_differing,
                                              "what": "tuning variant 4 (32-byte quantised nodes, opt-in, not bit-exact at ties) on the same streams"},
                          "note": "per GPU, separate launches, best of 5; l1_frac as in `roofline`"},
            "roofline": {"bound": "l1", "achieved": round(l1_achieved, 1), "peak": round(l1_peak, 1), "unit": "Gwavefronts/s",
                         "frac": round(l1_achieved / l1_peak, 4), "traffic": traffic, "kernel": "tracePackedKernel", "kernel_ms": round(kernel_ms, 4),
                         "achieved_is": ("l1tex__data_pipe_lsu_wavefronts.sum per launch (committed ncu capture of this kernel source and batch) / live kernel time"
                                         if wavefronts else "algorithmic gather instructions / live kernel time (no matching ncu capture): an upper bound, lanes share sectors"),
                         "l1_wavefronts_per_launch": wavefronts, "gather_instructions_per_launch": gathers,
                         "algorithmic_frac": round(l1_algorithmic / l1_peak, 4),
                         "peak_source": f"{sm_count} SMs x {sm_mhz:.0f} MHz (sampled during the run) x 1 L1 data-pipe wavefront per clock per SM "
                                        "(measured: tools/micro/l1_wavefronts.cu, profiles/r01_l1_wavefront_microbench.md)",
                         "traffic_source": traffic_src, "ncu_limiters": limiters,
                         "note": "What bounds this kernel on battlefield is the L1 data pipe, not HBM: the 3.2 MB scene is L1/L2-resident and every "
                                 "divergent load instruction costs about one data-pipe wavefront per lane whatever its width. frac = measured "
                                 "wavefronts / (SMs x clock x time): <= 1 by construction. algorithmic_frac uses the per-lane gather instructions the "
                                 "kernel counts itself (2 per inner node, 2 per pair, 1 per stack push and pop, 3 for ray in / result out, 1 remap "
                                 "word per hit, 2 probe rows per miss); it passes 1 where lanes of a warp fetch the same sector (primary rays: "
                                 "breakdown.primary_l1_frac) and sits near 1 for the incoherent bounce streams (breakdown.secondary_l1_frac). "
                                 "`traffic` = ncu DRAM bytes per launch, quoted only while the committed capture matches this kernel source.",
                         "hbm_cache_served": {"bound": "hbm", "achieved": round(achieved_hbm, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved_hbm / peak, 4),
                                              "algorithmic_bytes_per_launch": alg_bytes, "compulsory_bytes_per_launch": 48 * rays_per_step,
                                              "peak_source": peak_src,
                                              "note": "SURVEY 8d's algorithmic bytes (32+16+64*N_inner+48*N_pair+4*[hit]+64*[miss]) over kernel time: NOT a "
                                                      "roofline fraction on this scene -- the bytes are served by L1/L2, only the compulsory 48 B/ray reach DRAM "
                                                      "(see c5 for the config where HBM is the bound)"}},
            "e2e": {"value": round(e2e_value, 2), "unit": "Mrays/s", "h2d_bytes_per_step": 32 * rays_per_step, "d2h_bytes_per_step": 16 * rays_per_step,
                    "ms_per_step": round(e2e_ms, 3), "wall_ms_per_step": round(e2e_wall_ms / e2e_steps, 3), "steps": e2e_steps,
                    "results_match_device_run": e2e_ok, "numa": numa,
                    "path": "racc_cuda_trace with RACC_CUDA_STREAM_HOST descriptors, pinned host memory",
                    "pcie_gbs": {"h2d": round(32 * rays_all_ranks / (e2e_ms * 1e-3) / 1e9 / world, 1), "d2h": round(16 * rays_all_ranks / (e2e_ms * 1e-3) / 1e9 / world, 1),
                                 "note": "per GPU"},
                    "pcie_ceiling": ceiling, "frac_of_pcie_ceiling": round(e2e_value / ceiling["mrays_ceiling_all_ranks"], 3)},
            "gpu_launches": int(launches), "clocks": clocks,
            "frame_reduce": {"rays": frame["rays"], "hits": frame["hits"], "equals_sum_over_ranks": frame_ok,
                             "by": ("racc_cuda_frame_reduce (ncclAllReduce inside libracc_b200.so)" if world > 1
                                    else "racc_cuda_frame_reduce (one device: no collective)")},
        }
        if one_process is not None:
            line["e2e_one_process"] = one_process
        line["device_render"] = device_render
        if c3 is not None:
            line["c3"] = c3
        if c4 is not None:
            line["c4"] = c4
        if c5 is not None:
            line["c5"] = c5
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
            line["parity_sample_bit_exact"] = parity
        emit(line)
    if world > 1:
        rb.comm_destroy()
        dist.destroy_process_group()


def run_c5(args, rb, torch, rank, world, stream, gmax, gsum, barrier):
    """BASELINE configs[4]: 10 M-triangle random soup (0.86 GB of node + pair images: far larger than L2), 100 M uniform random
    rays split by index over the ranks, the engine's default (auto) path: rays re-binned by origin, stack tops in shared
    memory. Reported with its own HBM roofline; on rank 0 at N=1 an oracle-checked parity sample of >= 1 M rays."""
    tris, total_rays = args.c5_triangles, args.c5_rays
    if tris <= 0 or total_rays <= 0:
        return None
    v, i = rb.synthetic_triangles(tris, seed=7, extent=1000.0, edge=2.0)
    t0 = time.perf_counter()
    scene = rb.create_scene(v, i)
    build_s = time.perf_counter() - t0
    del v, i
    lo, hi = (total_rays * rank) // world, (total_rays * (rank + 1)) // world
    n = hi - lo
    # uniform origins in the scene box, uniform directions on the sphere, generated on the device per rank
    g = torch.Generator(device="cuda")
    g.manual_seed(8 + rank)
    rays = torch.empty(n, 8, dtype=torch.float32, device="cuda")
    rays[:, 0:3] = torch.rand(n, 3, generator=g, device="cuda") * 1000.0
    d = torch.randn(n, 3, generator=g, device="cuda")
    rays[:, 4:7] = d / d.norm(dim=1, keepdim=True)
    rays[:, 3] = 0.0
    rays[:, 7] = 1e6
    del d
    res = torch.empty(n * 4, dtype=torch.float32, device="cuda")
    cnt = torch.zeros(8, dtype=torch.int64, device="cuda")
    desc = [(rays.data_ptr(), res.data_ptr(), n)]
    rb.trace_device(scene, None, desc, stream=stream, counters_ptr=cnt.data_ptr(), detail=True)
    torch.cuda.synchronize()
    c = [int(x) for x in cnt.cpu().tolist()]
    alg = 48 * n + 64 * c[2] + 48 * c[3] + 4 * c[1]
    best = 1e30
    launches0 = rb.launch_count()
    for rep in range(4):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        rb.trace_device(scene, None, desc, stream=stream)
        b.record(stream)
        b.synchronize()
        t = gmax(a.elapsed_time(b))
        if rep:
            best = min(best, t)
    launches = (rb.launch_count() - launches0) // 4
    # the same pass on the 32-byte quantised node image (tuning variant 4, opt-in: not bit-exact at ties)
    res_q = torch.empty(n * 4, dtype=torch.float32, device="cuda")
    desc_q = [(rays.data_ptr(), res_q.data_ptr(), n)]
    best_q = 1e30
    rb.set_tuning(variant=4)
    try:
        for rep in range(4):
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            rb.trace_device(scene, None, desc_q, stream=stream)
            b.record(stream)
            b.synchronize()
            t = gmax(a.elapsed_time(b))
            if rep:
                best_q = min(best_q, t)
    finally:
        rb.set_tuning(variant=3)
    ri, qi = res.view(-1, 4).view(torch.int32), res_q.view(-1, 4).view(torch.int32)
    same_id = ri[:, 0] == qi[:, 0]
    ids_differing = gsum(int((~same_id).sum()))
    bits_equal_where_ids_agree = gsum(0 if torch.equal(ri[same_id], qi[same_id]) else 1) == 0
    del res_q, ri, qi, same_id
    rb.frame_reduce(stream)
    rec = None
    alg_all = gsum(alg)
    peak, peak_src = measured_hbm_peak()
    if rank == 0:
        achieved = alg_all / (best * 1e-3) / 1e9 / world  # per GPU
        traffic, traffic_src, limiters, _ = ncu_capture("ncu_c5_traffic")
        rec = {"value": round(total_rays / best / 1e3, 1), "unit": "Mrays/s", "ms": round(best, 3), "rays": total_rays, "triangles": tris,
               "scaling": "strong", "scene_bytes": (scene.info["node_count"] + scene.info["pair_count"]) * 64, "tree_depth": scene.info["depth"],
               "build_s": round(build_s, 3), "kernels_per_pass": launches, "hit_rate": round(c[1] / n, 4),
               "inner_per_ray": round(c[2] / n, 2), "pairs_per_ray": round(c[3] / n, 2),
               "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                            "algorithmic_bytes_per_pass_per_gpu": alg_all // world, "peak_source": peak_src, "traffic_source": traffic_src, "ncu_limiters": limiters,
                            "note": "algorithmic bytes (SURVEY 8d) over the time of the whole pass (re-binning sort + traversal), per GPU. The scene is 7x "
                                    "L2, so HBM is the real bound for rays in arrival order; re-binned, ~4/5 of the algorithmic bytes are served by L2 "
                                    "and the limiter moves to the L1 data pipe (profiles/r01_ncu_c5_rebinned_launch.txt) -- a frac near 1 here means "
                                    "'as fast as if every algorithmic byte came from HBM at copy speed', not 'HBM busy'"},
               "quantised_nodes": {"value": round(total_rays / best_q / 1e3, 1), "unit": "Mrays/s", "ms": round(best_q, 3), "speedup": round(best / best_q, 3),
                                   "ids_differing_from_exact": ids_differing, "bits_equal_where_ids_agree": bits_equal_where_ids_agree,
                                   "what": "the same pass with tuning variant 4: 32-byte nodes, child boxes as 16-bit grid coordinates rounded outwards "
                                           "(one gather per node visit, node image halved). Opt-in: exact ties in t may resolve to the other triangle "
                                           "(north_star's bar, tests/test_gpu_parity.py::test_quantised_nodes_*); the default stays bit-exact"},
               "what": "BASELINE configs[4]: random soup, uniform random rays split by index over the ranks (strong scaling), engine defaults "
                       "(auto: origin re-binning + stack tops in shared memory); best of 3 passes, max over ranks"}
        if world == 1 and not args.no_cpu_baseline:
            import oracle
            m = min(n, args.c5_parity_rays)
            nodes, pairs, remap = scene.download()
            img = oracle.SceneImages(nodes, pairs, remap)
            idx = torch.arange(0, n, max(1, n // m), device="cuda")[:m]
            sample = rays[idx].cpu().numpy().reshape(-1).view(oracle.RAY_DTYPE)
            got = res.view(-1, 4)[idx].cpu().numpy().view(np.uint32)
            t0 = time.perf_counter()
            want = oracle.traverse(img, sample).view(np.uint32).reshape(-1, 4)
            rec["parity_sample_bit_exact"] = bool(np.array_equal(got, want))
            rec["parity_sample"] = (f"{sample.shape[0]} rays (every {max(1, n // m)}th) against the CPU oracle on the downloaded images, "
                                    f"{time.perf_counter() - t0:.1f} s")
            del img, nodes, pairs, remap
    scene.destroy()
    return rec


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["engine", "reference"], default="engine")
    ap.add_argument("--cpu-sample-rays", type=int, default=20_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--c5-triangles", type=int, default=10_000_000)
    ap.add_argument("--c5-rays", type=int, default=100_000_000)
    ap.add_argument("--c5-parity-rays", type=int, default=1_000_000)
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
