"""The host scheduler behind include/RayAccelerator.h (rayaccel_b200/csrc/racc_api.cpp).

CPU tests link racc_api.cpp against tests/harness/fake_capi.cpp (an oracle-backed test double of the
C-ABI) and check the callback contract of SURVEY.md section 8b plus result routing. GPU tests link
the same client against the product library. The reference's UNMODIFIED example renderers are run
through the same API when oracle/_ref/racc_render_{cpu,gpu} exist (built where /root/reference is
present; BASELINE.json configs[0] = the Whitted 512x512 1-bounce plumbing case)."""
import json
import os
import subprocess

import pytest

import oracle
from rayaccel_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "tests", "harness")
BUILD = os.path.join(HARNESS, "_build")
CXX = ["g++", "-std=c++17", "-O2", "-mavx2", "-mfma", "-ffp-contract=off", "-pthread", "-I", os.path.join(ROOT, "include")]


def _stale(target, sources):
    return not os.path.exists(target) or any(os.path.getmtime(s) > os.path.getmtime(target) for s in sources)


@pytest.fixture(scope="module")
def plumbing_cpu():
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "plumbing_cpu")
    src = [os.path.join(HARNESS, "plumbing_client.cpp"), os.path.join(HARNESS, "fake_capi.cpp"),
           os.path.join(ROOT, "rayaccel_b200", "csrc", "racc_api.cpp"), os.path.join(ROOT, "rayaccel_b200", "csrc", "scene_build.cpp")]
    if _stale(exe, src + [oracle.ORACLE_SO]):
        subprocess.run(CXX + ["-DRACC_FAKE_CAPI"] + src + ["-L", os.path.dirname(oracle.ORACLE_SO), "-loracle", "-Wl,-rpath," + os.path.dirname(oracle.ORACLE_SO), "-o", exe], check=True)
    return exe


@pytest.fixture(scope="module")
def plumbing_gpu():
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "plumbing_gpu")
    src = [os.path.join(HARNESS, "plumbing_client.cpp"), os.path.join(ROOT, "rayaccel_b200", "csrc", "scene_build.cpp")]
    libdir = os.path.dirname(_lib.LIB_PATH)
    if _stale(exe, src + [oracle.ORACLE_SO, _lib.LIB_PATH]):
        subprocess.run(CXX + src + ["-L", libdir, "-lracc_b200", "-Wl,-rpath," + libdir,
                                    "-L", os.path.dirname(oracle.ORACLE_SO), "-loracle", "-Wl,-rpath," + os.path.dirname(oracle.ORACLE_SO), "-o", exe], check=True)
    return exe


def run_json(cmd, timeout=600):
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert lines, f"no JSON from {cmd}: rc={p.returncode}\n{p.stdout}\n{p.stderr}"
    return p.returncode, json.loads(lines[-1])


CONFIGS = [
    [],                                                                               # defaults of the client
    ["--threads", "1", "--submitters", "1", "--rays", "50000"],                      # fully serial
    ["--threads", "8", "--submitters", "3", "--rays", "300000", "--frames", "3"],    # contended
    ["--spawn", "16384", "--shade", "8192", "--batch", "11264", "--inflight", "262144"],  # the reference's defaults (RayAccelerator.cpp:429-446)
    ["--spawn", "1000", "--shade", "333", "--batch", "777", "--inflight", "5000", "--rays", "40001"],  # ragged everything
    ["--rays", "1"],
    ["--spawn", "16384", "--shade", "8192", "--batch", "49152", "--inflight", "2097152", "--rays", "400000"],  # the B200 defaults
]


@pytest.mark.parametrize("args", CONFIGS, ids=lambda a: "_".join(a).replace("--", "") or "default")
def test_scheduler_contract_cpu(plumbing_cpu, args):
    rc, out = run_json([plumbing_cpu] + args)
    assert rc == 0 and out["ok"], out
    assert out["rays_traced"] == out["rays_expected"] and out["mismatches"] == 0 and out["missing"] == 0 and out["violations"] == 0


@pytest.mark.parametrize("devices", [2, 4])
def test_one_context_over_several_devices_deals_streams_and_reduces_cpu(plumbing_cpu, devices):
    """racc::cudaDevices(0, n): one context, n pretend devices behind the test double. Every device gets its own submitters
    and a share of the ready streams, every result still comes back to the right ray (checked against the oracle by the
    client), and Stats.raysTraced of each frame is the sum of the devices' frame records (one reduction per frame). The
    submitters release their per-thread device scratch when the context is destroyed."""
    frames, submitters = 3, 2
    p = subprocess.run([plumbing_cpu, "--devices", str(devices), "--threads", "8", "--submitters", str(submitters), "--rays", "400000",
                        "--frames", str(frames), "--inflight", "1048576", "--batch", "8192"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env={**os.environ, "FAKE_CAPI_DEVICES": str(devices)})
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert p.returncode == 0 and len(lines) == 2, p.stdout + p.stderr
    fake, out = lines
    assert out["ok"] and out["rays_traced"] == out["rays_expected"] and out["mismatches"] == 0 and out["missing"] == 0
    assert fake["reduces"] == frames and fake["thread_releases"] == devices * submitters
    assert sum(fake["device_rays"]) == out["rays_traced"]
    assert all(r > 0.02 * out["rays_traced"] for r in fake["device_rays"]), f"a device was starved: {fake}"
    assert "devices counted" not in p.stderr


def test_reference_whitted_renderer_unchanged_cpu_plumbing():
    """BASELINE.json configs[0]: Whitted renderer, battlefield.bin, 512x512, 1 bounce, through the API."""
    exe = os.path.join(ROOT, "oracle", "_ref", "racc_render_cpu")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/racc_render_cpu not built (needs /root/reference; `make -C oracle renderer`)")
    rc, out = run_json([exe, "--whitted", "--width", "512", "--height", "512", "--depth", "1", "--frames", "2", "--threads", "4"])
    assert rc == 0
    assert out["rendered_width"] == 512 and out["rendered_height"] == 512
    # 262 144 primaries + one reflect/refract generation; jitter is rand()-seeded so only bounds are stable
    assert 262144 < out["rays_first_frame"] < 3 * 262144
    assert out["nonblack_fraction"] > 0.9 and out["not_finite"] == 0 and 0.2 < out["mean_luminance"] < 5.0


def test_reference_path_tracer_unchanged_cpu_plumbing():
    exe = os.path.join(ROOT, "oracle", "_ref", "racc_render_cpu")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/racc_render_cpu not built")
    rc, out = run_json([exe, "--width", "384", "--height", "256", "--frames", "2", "--threads", "3"])
    assert rc == 0
    assert out["max_depth"] == 3 and 384 * 256 < out["rays_first_frame"] < 4 * 384 * 256
    assert out["nonblack_fraction"] > 0.9 and out["not_finite"] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("args", [CONFIGS[0], CONFIGS[2], CONFIGS[4], CONFIGS[6]], ids=["default", "contended", "ragged", "b200_defaults"])
def test_scheduler_contract_gpu(plumbing_gpu, args):
    """Same client, product library: results routed through racc::render() equal the oracle's bit for bit."""
    rc, out = run_json([plumbing_gpu] + args)
    assert rc == 0 and out["ok"], out


@pytest.mark.gpu
def test_reference_renderers_unchanged_gpu():
    """The reference's path tracer (1920x1080 -> 1920x1024 tiles, BASELINE.json configs[1] through the
    API) and Whitted renderer, unmodified, on the CUDA engine."""
    exe = os.path.join(ROOT, "oracle", "_ref", "racc_render_gpu")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/racc_render_gpu not built")
    rc, out = run_json([exe, "--width", "1920", "--height", "1080", "--frames", "4"])
    assert rc == 0, out
    assert out["rendered_width"] == 1920 and out["rendered_height"] == 1024
    assert 1920 * 1024 < out["rays_first_frame"] < 4 * 1920 * 1024
    assert out["nonblack_fraction"] > 0.9 and out["not_finite"] == 0
    rc, out = run_json([exe, "--whitted", "--width", "512", "--height", "512", "--depth", "1", "--frames", "2"])
    assert rc == 0 and out["nonblack_fraction"] > 0.9 and out["not_finite"] == 0


def test_scheduler_under_thread_sanitizer(tmp_path):
    """The scheduler state machine (racc_api.cpp) under ThreadSanitizer, contended configuration: no data race reports.
    (The reference guards all of its scheduler state with one mutex, RayAccelerator.cpp:48-415, and was never checked.)"""
    exe = str(tmp_path / "plumbing_tsan")
    src = [os.path.join(HARNESS, "plumbing_client.cpp"), os.path.join(HARNESS, "fake_capi.cpp"),
           os.path.join(ROOT, "rayaccel_b200", "csrc", "racc_api.cpp"), os.path.join(ROOT, "rayaccel_b200", "csrc", "scene_build.cpp")]
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-mavx2", "-mfma", "-ffp-contract=off", "-pthread", "-I", os.path.join(ROOT, "include")] + src + \
          ["-L", os.path.dirname(oracle.ORACLE_SO), "-loracle", "-Wl,-rpath," + os.path.dirname(oracle.ORACLE_SO), "-o", exe]
    built = subprocess.run(cmd, capture_output=True, text=True)
    if built.returncode != 0:
        pytest.skip("this toolchain cannot build with -fsanitize=thread: " + built.stderr[-200:])
    for args in (["--threads", "8", "--submitters", "3", "--rays", "200000", "--frames", "2"],
                 ["--spawn", "1000", "--shade", "333", "--batch", "777", "--inflight", "5000", "--rays", "40001"]):
        p = subprocess.run([exe] + args, capture_output=True, text=True, timeout=600, cwd=ROOT, env={**os.environ, "TSAN_OPTIONS": "halt_on_error=0"})
        assert "WARNING: ThreadSanitizer" not in p.stderr, p.stderr[:2000]
        lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
        assert p.returncode == 0 and lines and json.loads(lines[-1])["ok"]
