#!/usr/bin/env python
"""AddressSanitizer / UBSan over the CPU test build of the engine library: the kernels' and the
host code's own source, compiled by g++ with -fsanitize=..., driven through the C-ABI -- host staging (both tail
policies), every traversal variant, re-binning, both device renderers (with and without tuning keys 15/16) and the device
scene builder. compute-sanitizer does this on the B200 (tools/sanitize.sh); this is the same question asked where there is
no GPU.   python tests/harness/sanitize_cpu_build.py address|undefined [seconds of tests/fuzz/fuzz_gpu.py on top]
First run `python -m pytest tests/test_library_on_cpu.py -k builds` so that the rewritten sources exist."""
import subprocess
SAN = (__import__("sys").argv[1:] or ["address"])[0]
LIBDIR = "/root/repo/tests/harness/_build/library"
OUT = f"/tmp/libracc_{SAN}.so"
FLAGS = ["-fsanitize=address"] if SAN == "address" else ["-fsanitize=undefined", "-fno-sanitize=alignment,vptr"]
if __import__("os").environ.get("RACC_SANITIZE_CHILD") != "1":
    import glob, os, sys
    srcs = sorted(glob.glob(LIBDIR + "/*_on_cpu.cpp")) + ["/root/repo/rayaccel_b200/csrc/scene_build.cpp", "/root/repo/rayaccel_b200/csrc/racc_api.cpp"]
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fno-omit-frame-pointer", "-mavx2", "-mfma", "-ffp-contract=off", "-fno-fast-math", "-fPIC",
                    "-shared", "-w", "-pthread", "-DCUDA_ON_CPU_UCONTEXT"] + FLAGS + ["-I", "/root/repo/tests/harness/cuda_on_cpu", "-I", LIBDIR, "-I", "/root/repo/rayaccel_b200/csrc"] +
                   srcs + ["-o", OUT], check=True)
    rt = subprocess.check_output(["gcc", "-print-file-name=" + ("libasan.so" if SAN == "address" else "libubsan.so")], text=True).strip()
    env = dict(os.environ, RACC_SANITIZE_CHILD="1", LD_PRELOAD=rt, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1")
    sys.exit(subprocess.run([sys.executable, __file__, SAN] + sys.argv[2:], env=env).returncode)
import sys, os, ctypes
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
os.environ["RACC_B200_BUILD_DEVICE"] = "0"; os.environ["RACC_B200_HOST_CHUNK"] = "1024"
import numpy as np
import oracle, rayaccel_b200 as rb
from rayaccel_b200 import _lib
from oracle import raygen
from test_render_oracle import camera_for
lib = ctypes.CDLL(OUT)
for name, (restype, argtypes) in _lib.SYMBOLS.items():
    fn = getattr(lib, name); fn.restype, fn.argtypes = restype, argtypes
_lib._lib = lib
rb.init(0)
sf = rb.load_scene()
scene = rb.create_scene(sf.vertices, sf.indices); env = rb.create_environment(sf.environment)
nodes, pairs, remap = scene.download(); images = oracle.SceneImages(nodes, pairs, remap, sf.environment)
cam = raygen.look_at(sf.cam_origin, sf.cam_target, sf.cam_up, sf.cam_fov, 80, 45)
primary = raygen.primary_rays(cam, 80, 45)
bounce = raygen.bounce_rays(sf.vertices, sf.indices, primary, oracle.traverse(images, primary), seed=2)
print("host streams", flush=True)
for taper in (0, 1):
    rb.set_tuning(host_taper=taper)
    got = rb.trace_host(scene, env, np.concatenate([primary, bounce]))
rb.set_tuning(host_taper=0)
print("variants", flush=True)
for variant, sort, ss in [(3,0,0),(3,1,0),(3,1,16),(0,0,0),(1,0,0),(2,0,0)]:
    rb.set_tuning(variant=variant, sort=sort, smem_stack=ss)
    s = np.ascontiguousarray(np.concatenate([primary, bounce])); o = np.zeros(len(s), dtype=oracle.RESULT_DTYPE); c = np.zeros(8, np.uint64)
    rb.trace_device(scene, env, [(s.ctypes.data, o.ctypes.data, len(s))], counters_ptr=c.ctypes.data); rb.sync()
rb.set_tuning(variant=3, sort=2, smem_stack=-1)
print("renderers", flush=True)
sh = rb.create_shading(sf.normals, sf.triangle_normals, sf.materials)
camr = camera_for(sf, 48, 32)
rb.path_trace(scene, env, sh, camr, 48, 32, 3, 3, 1, batch_spp=1)
for a, c in [(0,0),(1,1)]:
    rb.set_tuning(whitted_arena=a, whitted_combine=c)
    rb.whitted_trace(scene, env, sh, camr, 48, 32, 2, 8, 11)
    rb.whitted_trace(scene, env, sh, camr, 33, 17, 1, 8, 12)
rb.set_tuning(whitted_arena=0, whitted_combine=0)
print("device build", flush=True)
v, i = rb.synthetic_triangles(300, seed=11, extent=50.0, edge=4.0)
rb.set_tuning(build_device=2); s2 = rb.create_scene(v, i); s2.destroy(); rb.set_tuning(build_device=0)
sh.destroy(); env.destroy(); scene.destroy()
if len(sys.argv) > 2:  # ... and N seconds of the randomised campaign (tests/fuzz/fuzz_gpu.py) through the same sanitised library:
    print("randomised campaign", flush=True)  # degenerate scenes, ragged streams, quantised nodes, both renderer forms, Whitted
    sys.path.insert(0, "/root/repo/tests/fuzz")
    import fuzz_gpu
    sys.argv = ["fuzz", "--seconds", sys.argv[2], "--rays", "300", "--seed", "31"]
    if fuzz_gpu.main():
        sys.exit(1)
print(f"sanitizer run complete ({SAN}): no report above means clean", flush=True)
