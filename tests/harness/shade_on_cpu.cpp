// shade_on_cpu.cpp -- TEST INFRASTRUCTURE ONLY. C entry points over the engine's shading-kernel launchers
// (rayaccel_b200/csrc/{pathtrace,whitted}.cu) compiled for the CPU over tests/harness/cuda_on_cpu/cuda_runtime.h.
// tests/test_kernels_on_cpu.py builds this together with the two .cu files (their <<<...>>> launches rewritten to
// cuda_on_cpu::launch, nothing else touched) and drives the wave loops with the checker's traversal in between.
#include "engine.h"

#include <pmmintrin.h>
#include <xmmintrin.h>

using namespace racc_b200;

namespace {
// the device runs with flush-to-zero; so do the kernels here
struct FlushToZero {
	unsigned saved;
	FlushToZero() : saved(_mm_getcsr()) {
		_MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
		_MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
	}
	~FlushToZero() { _mm_setcsr(saved); }
};
} // namespace

extern "C" {

int cpu_path_primary(const float* camera12, uint32_t width, uint32_t height, uint32_t sampleBase, uint32_t firstPath, uint32_t count,
                     uint32_t seed, void* rays, void* states) {
	FlushToZero ftz;
	return launchPathPrimary(camera12, width, height, sampleBase, firstPath, count, seed, static_cast<DevRay*>(rays), static_cast<float4*>(states), nullptr, nullptr);
}

int cpu_path_shade(const void* rays, const void* results, const void* states, uint32_t count, uint32_t depth, uint32_t maxDepth, uint32_t seed,
                   uint32_t pixels, uint32_t sampleBase, const uint32_t* indices, const void* normals, const void* triangleNormals,
                   const uint16_t* triangleMaterials, const void* materials, uint32_t triangleCount, uint32_t materialCount, void* outRays,
                   void* outStates, uint32_t* outCount, void* radiance) {
	FlushToZero ftz;
	PathShadeParams p{};
	p.rays = static_cast<const DevRay*>(rays); p.results = static_cast<const float4*>(results); p.states = static_cast<const float4*>(states);
	p.count = count; p.depth = depth; p.maxDepth = maxDepth; p.seed = seed; p.pixels = pixels; p.sampleBase = sampleBase;
	p.indices = indices; p.normals = static_cast<const float4*>(normals); p.triangleNormals = static_cast<const float4*>(triangleNormals);
	p.triangleMaterials = triangleMaterials; p.materials = static_cast<const float4*>(materials);
	p.triangleCount = triangleCount; p.materialCount = materialCount;
	p.outRays = static_cast<DevRay*>(outRays); p.outStates = static_cast<float4*>(outStates); p.outCount = outCount;
	p.radiance = static_cast<float4*>(radiance);
	return launchPathShade(p, nullptr, nullptr);
}

int cpu_path_accumulate(const void* radiance, uint32_t pixels, uint32_t spp, void* framebuffer) {
	FlushToZero ftz;
	return launchPathAccumulate(static_cast<const float4*>(radiance), pixels, spp, static_cast<float4*>(framebuffer), nullptr, nullptr);
}

int cpu_whitted_primary(const float* camera12, uint32_t width, uint32_t height, uint32_t sampleBase, uint32_t firstPath, uint32_t count,
                        uint32_t seed, void* rays, void* states) {
	FlushToZero ftz;
	return launchWhittedPrimary(camera12, width, height, sampleBase, firstPath, count, seed, static_cast<DevRay*>(rays), static_cast<float4*>(states), nullptr, nullptr);
}

int cpu_whitted_shade(const void* rays, const void* results, const void* states, uint32_t count, uint32_t depth, uint32_t maxDepth,
                      const uint32_t* indices, const void* normals, const void* triangleNormals, uint32_t triangleCount, void* outRays,
                      void* outStates, uint32_t* outCount, unsigned long long* accumulators, int combine) {
	FlushToZero ftz;
	WhittedShadeParams p{};
	p.rays = static_cast<const DevRay*>(rays); p.results = static_cast<const float4*>(results); p.states = static_cast<const float4*>(states);
	p.count = count; p.depth = depth; p.maxDepth = maxDepth;
	p.indices = indices; p.normals = static_cast<const float4*>(normals); p.triangleNormals = static_cast<const float4*>(triangleNormals);
	p.triangleCount = triangleCount;
	p.outRays = static_cast<DevRay*>(outRays); p.outStates = static_cast<float4*>(outStates); p.outCount = outCount;
	p.accumulators = accumulators;
	p.combine = combine != 0;
	return launchWhittedShade(p, nullptr, nullptr);
}

int cpu_whitted_finish(const unsigned long long* accumulators, uint32_t pixels, void* framebuffer) {
	FlushToZero ftz;
	return launchWhittedFinish(accumulators, pixels, static_cast<float4*>(framebuffer), nullptr, nullptr);
}

} // extern "C"
