"""The drop-in boundary on a machine without a GPU: the C-ABI library loads, exports every symbol
include/racc_b200.h declares plus the C++ racc:: API of include/RayAccelerator.h, and fails loudly
(no CPU fallback) when asked to compute without a CUDA device."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import rayaccel_b200 as rb
from rayaccel_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "racc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(racc_cuda_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"libracc_b200.so does not export {n}"
    assert sorted(_lib.SYMBOLS) == names, "rayaccel_b200/_lib.py SYMBOLS out of sync with include/racc_b200.h"
    assert lib.racc_cuda_abi_version() == 2


def test_cpp_api_is_exported():
    out = subprocess.run(["nm", "-D", "--defined-only", "-C", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    for sym in ["racc::init()", "racc::deinit()", "racc::defaultConfiguration(", "racc::createContext(", "racc::destroy(racc::Context*)",
                "racc::info(", "racc::createScene(", "racc::destroy(racc::Scene*)", "racc::createEnvironment(",
                "racc::destroy(racc::Environment*)", "racc::render("]:
        assert sym in out, f"{sym} missing from libracc_b200.so"


def test_header_layouts_match_reference_abi(tmp_path):
    """sizeof/offsetof of every POD in include/RayAccelerator.h equal the reference's (RayAccelerator.h:32-93)."""
    src = tmp_path / "layout.cpp"
    src.write_text(r'''
#include <RayAccelerator.h>
#include <cstddef>
#include <cstdio>
int main() {
	using namespace racc;
	printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(Configuration), sizeof(ContextInfo), sizeof(Vertex), sizeof(Color), sizeof(Ray),
	       sizeof(Result), sizeof(RayStream), sizeof(Stats), sizeof(RenderCallbacks));
	printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", offsetof(Configuration, allowCpuTracing), offsetof(Configuration, cpuThreads),
	       offsetof(Configuration, gpuSubmissionThreads), offsetof(Configuration, maxRaysInFlight), offsetof(Configuration, maxRaysPerSpawn),
	       offsetof(Configuration, cpuTestBatch), offsetof(Configuration, cpuShadeBatch), offsetof(Configuration, rayStreamBatchSize));
	printf("%zu %zu %zu %zu %zu\n", alignof(Ray), alignof(Result), alignof(Vertex), offsetof(Ray, dir), offsetof(Result, hit.t));
	return invalidTriangle == 0xffffffffu ? 0 : 1;
}''')
    exe = tmp_path / "layout"
    subprocess.run(["g++", "-std=c++17", "-mavx2", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    assert out[0].split() == ["24", "12", "16", "16", "32", "16", "24", "8", "24"]
    assert out[1].split() == ["8", "9", "10", "12", "16", "18", "20", "22"]
    assert out[2].split() == ["32", "16", "16", "16", "4"]


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="needs a machine WITHOUT a GPU")
def test_no_gpu_means_loud_failure_not_fallback(battlefield):
    lib = _lib.load()
    assert lib.racc_cuda_device_count() <= 0
    with pytest.raises(rb.EngineError, match="no CPU fallback|no CUDA|insufficient|CUDA"):
        rb.init(0)
    with pytest.raises(rb.EngineError):
        rb.create_scene(battlefield.vertices, battlefield.indices)
    with pytest.raises(rb.EngineError):
        rb.create_environment(battlefield.environment)
    with pytest.raises(rb.EngineError):  # the device-side renderer has no host path either
        rb.create_shading(battlefield.normals, battlefield.triangle_normals, battlefield.materials)
    assert lib.racc_cuda_path_trace(None, None, None, None, None, None, None, None) != 0
    assert b"null argument" in lib.racc_cuda_last_error()
    with pytest.raises(ValueError):
        rb.create_shading(battlefield.normals[:, :3], battlefield.triangle_normals, battlefield.materials)
    with pytest.raises(ValueError):
        rb.create_shading(battlefield.normals, battlefield.triangle_normals[:10], battlefield.materials)
    rays = np.zeros(4, rb.RAY_DTYPE)
    assert lib.racc_cuda_trace(None, None, None, 0, None) != 0
    assert b"null scene" in lib.racc_cuda_last_error()
    assert not lib.racc_cuda_host_alloc(4096)
    assert rays.shape == (4,)


def test_host_only_entry_points_work_without_gpu(battlefield):
    h = rb.HostImages(battlefield.vertices, battlefield.indices)
    assert h.nodes.shape == (h.info["node_count"], 16) and h.pairs.shape == (h.info["pair_count"], 12)
    assert np.allclose(h.info["bounds_min"], battlefield.vertices[:, :3].min(0), atol=1e-3)
    assert np.allclose(h.info["bounds_max"], battlefield.vertices[:, :3].max(0), atol=1e-3)


def test_tuning_keys_documented_and_settable_without_gpu():
    """racc_cuda_set_tuning is host-only: every key the header documents is accepted and returns the previous value, the
    defaults are the measured ones (profiles/r02_call1_open_questions.md: 15 on, 16 off, 17 = 256 K rays), an unknown key is an
    error with a message."""
    lib = _lib.load()
    lib.racc_cuda_set_tuning.restype = ctypes.c_int
    header = open(os.path.join(ROOT, "include", "racc_b200.h")).read()
    for env in ["_WHITTED_ARENA", "_WHITTED_COMBINE", "_HOST_TAPER", "_SMEM_STACK", "_HOST_ZERO_COPY"]:
        assert env in header
    defaults = {15: 1, 16: 0, 17: 256, 18: 0, 19: 0, 20: 4}
    for key in range(21):
        prev = lib.racc_cuda_set_tuning(key, 1)
        assert lib.racc_cuda_set_tuning(key, prev) == 1, f"key {key} did not keep the value"
        if key in defaults and not os.environ.get("RACC_B200_HOST_TAPER"):
            assert prev == defaults[key], f"key {key}: default {prev}, documented {defaults[key]}"
    assert lib.racc_cuda_set_tuning(99, 1) == -1
    assert b"unknown tuning key" in lib.racc_cuda_last_error()


def test_product_never_touches_the_oracle():
    """Static check: nothing under rayaccel_b200/, include/ imports, links or loads oracle/."""
    bad = []
    for base in ("rayaccel_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".cpp", ".h")):
                    text = open(os.path.join(dirpath, fn), errors="replace").read()
                    if re.search(r"import\s+oracle|from\s+oracle|liboracle|racc_oracle\.h|oracle_traverse|oracle_path_trace\s*\(|oracle_material_sample\s*\(", text):
                        bad.append(os.path.join(dirpath, fn))
    assert not bad, bad
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
