// kernel_host.cpp -- TEST INFRASTRUCTURE ONLY: runs the reference's OpenCL `traversal` kernel, compiled from its own
// source text (oracle/_ref/traversal_kernel.cl, extracted from /root/reference/RayAccelerator/Kernels.h by
// oracle/Makefile), on the CPU: one work-item per ray, as clEnqueueNDRangeKernel would (RayAccelerator.cpp:380-403).
#include "opencl_c.h"

#include <pmmintrin.h>
#include <xmmintrin.h>

namespace ocl {
thread_local int g_globalId = 0;
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Wattributes"
#pragma GCC diagnostic ignored "-Wunused-variable"
#pragma GCC diagnostic ignored "-Wsign-compare"
#include "traversal_kernel.cl"
#pragma GCC diagnostic pop
} // namespace ocl

// rays: count x 8 floats; nodes: 16 floats each; pairs: 12 floats each; remap: words; env: RGBA32F or null;
// hits: count x 4 floats. Threads of the reference run with FTZ + DAZ (Threading.h:77-79).
extern "C" int ref_kernel_traverse(const float* nodes, const float* pairs, const uint32_t* remap, const float* env, uint32_t env_width,
                                   uint32_t env_height, const float* rays, uint32_t count, float* hits) {
	const unsigned saved = _mm_getcsr();
	_MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
	_MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
	ocl::image2d img{env, (int)env_width, (int)env_height};
	for (uint32_t i = 0; i < count; ++i) {
		ocl::g_globalId = (int)i;
		ocl::traversal(reinterpret_cast<ocl::float8*>(const_cast<float*>(rays)), reinterpret_cast<ocl::float4*>(const_cast<float*>(nodes)),
		               reinterpret_cast<ocl::float4*>(const_cast<float*>(pairs)), const_cast<unsigned*>(remap), reinterpret_cast<ocl::float4*>(hits),
		               (int)count, &img);
	}
	_mm_setcsr(saved);
	return 0;
}
