// racc_on_cpu_stubs.cpp -- TEST INFRASTRUCTURE ONLY. The CPU test build of the engine library (tests/test_library_on_cpu.py)
// compiles every .cu file of rayaccel_b200/csrc over tests/harness/cuda_on_cpu/cuda_runtime.h except bvh_build.cu, whose
// cooperative (grid-synchronised) kernels need all blocks resident at once; these two entry points decline, and
// racc_cuda_scene_create then builds the same images with the host builder (scene_build.cpp), as it does on a machine
// whose RCPSS differs from the table model.
#include "scene_build.h"

namespace racc_b200 {

bool buildBvh2Device(const float*, uint32_t, const uint32_t*, uint32_t, std::vector<BuildNode>*, std::vector<uint32_t>*, const char** error) {
	if (error) *error = "no device builder in the CPU test build";
	return false;
}

bool buildSceneImagesDevice(const float*, uint32_t, const uint32_t*, uint32_t, DeviceSceneImages*, const char** error) {
	if (error) *error = "no device builder in the CPU test build";
	return false;
}

} // namespace racc_b200
