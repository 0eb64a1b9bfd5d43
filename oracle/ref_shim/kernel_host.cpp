// kernel_host.cpp -- TEST INFRASTRUCTURE ONLY: runs the reference's OpenCL `traversal` kernel, compiled from its own
// source text (oracle/_ref/traversal_kernel.cl, extracted from /root/reference/RayAccelerator/Kernels.h by
// oracle/Makefile), on the CPU: one work-item per ray, as clEnqueueNDRangeKernel would (RayAccelerator.cpp:380-403).
#include "opencl_c.h"

#include <atomic>
#include <pmmintrin.h>
#include <thread>
#include <vector>
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

namespace {

struct Job {
	const float *nodes, *pairs;
	const uint32_t* remap;
	ocl::image2d img;
	const float* rays;
	uint32_t count;
	float* hits;
};

// work-items [begin, end) on the calling thread, which runs with FTZ + DAZ as the reference's threads do (Threading.h:77-79)
void runItems(const Job& j, uint32_t begin, uint32_t end) {
	const unsigned saved = _mm_getcsr();
	_MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
	_MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
	for (uint32_t i = begin; i < end; ++i) {
		ocl::g_globalId = (int)i;
		ocl::traversal(reinterpret_cast<ocl::float8*>(const_cast<float*>(j.rays)), reinterpret_cast<ocl::float4*>(const_cast<float*>(j.nodes)),
		               reinterpret_cast<ocl::float4*>(const_cast<float*>(j.pairs)), const_cast<unsigned*>(j.remap),
		               reinterpret_cast<ocl::float4*>(j.hits), (int)j.count, &j.img);
	}
	_mm_setcsr(saved);
}

} // namespace

// rays: count x 8 floats; nodes: 16 floats each; pairs: 12 floats each; remap: words; env: RGBA32F or null;
// hits: count x 4 floats. threads <= 1: the calling thread; otherwise work-items are handed out in batches of 1024
// (the reference's cpuTestBatch, RayAccelerator.cpp:438) to that many threads.
extern "C" int ref_kernel_traverse_mt(const float* nodes, const float* pairs, const uint32_t* remap, const float* env, uint32_t env_width,
                                      uint32_t env_height, const float* rays, uint32_t count, float* hits, int threads) {
	const Job job{nodes, pairs, remap, ocl::image2d{env, (int)env_width, (int)env_height}, rays, count, hits};
	if (threads <= 1) {
		runItems(job, 0, count);
		return 0;
	}
	std::atomic<uint32_t> next{0};
	std::vector<std::thread> pool;
	for (int t = 0; t < threads; ++t)
		pool.emplace_back([&] {
			for (;;) {
				const uint32_t b = next.fetch_add(1024u);
				if (b >= count) break;
				runItems(job, b, b + 1024u < count ? b + 1024u : count);
			}
		});
	for (auto& t : pool) t.join();
	return 0;
}

extern "C" int ref_kernel_traverse(const float* nodes, const float* pairs, const uint32_t* remap, const float* env, uint32_t env_width,
                                   uint32_t env_height, const float* rays, uint32_t count, float* hits) {
	return ref_kernel_traverse_mt(nodes, pairs, remap, env, env_width, env_height, rays, count, hits, 1);
}
