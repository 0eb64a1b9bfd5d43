// trace_on_cpu.cpp -- TEST INFRASTRUCTURE ONLY. C entry point over the default traversal kernel's launchers
// (rayaccel_b200/csrc/traverse_packed.cu: packNodesKernel, packPairsKernel, packEnvKernel, tracePackedKernel) compiled for
// the CPU over tests/harness/cuda_on_cpu/cuda_runtime.h after the two mechanical rewrites of cuda_on_cpu/rewrite.py
// (kernel launches, inline PTX). tests/test_kernels_on_cpu.py compares its results with the checker bit for bit.
#include "engine.h"

#include <pmmintrin.h>
#include <xmmintrin.h>

#include <vector>

using namespace racc_b200;

namespace {
struct FlushToZero {
	unsigned saved;
	FlushToZero() : saved(_mm_getcsr()) {
		_MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
		_MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
	}
	~FlushToZero() { _mm_setcsr(saved); }
};

template <typename T> T* aligned(size_t bytes) { return static_cast<T*>(aligned_alloc(64, (bytes + 127) & ~(size_t)63)); }
} // namespace

extern "C" {

// What capi.cu does for a launch over device-resident streams: packed copies of the images (once per scene there),
// one StreamRef per stream, the work cursor, then launchTracePacked. tuning = {block threads, CTAs per SM, refill
// threshold, leaf bail-out, inner bail-out, stack entries in shared memory}; counters4 may be null.
int cpu_trace_packed(const void* nodes, uint32_t nodeCount, const void* pairs, uint32_t pairCount, const uint32_t* remap, const void* env,
                     uint32_t envWidth, uint32_t envHeight, const void* const* rays, void* const* results, const uint32_t* counts,
                     uint32_t nstreams, unsigned long long* counters4, int counterMode, const int* tuning, int smCount, const uint32_t* perm) {
	FlushToZero ftz;
	float4* tnodes = aligned<float4>((size_t)nodeCount * 64 + 64);
	float4* tpairs = aligned<float4>((size_t)pairCount * 64 + 64);
	float4* envPairs = env ? aligned<float4>(((size_t)envWidth + 1) * envHeight * 32) : nullptr;
	int launches = 0;
	int rc = launchPackImages(static_cast<const float4*>(nodes), nodeCount, static_cast<const float4*>(pairs), pairCount, tnodes, tpairs, nullptr, &launches);
	if (!rc && env) rc = launchPackEnv(static_cast<const float4*>(env), envWidth, envHeight, envPairs, nullptr, &launches);

	std::vector<StreamRef> refs;
	uint32_t total = 0;
	for (uint32_t i = 0; i < nstreams; ++i) {
		if (!counts[i]) continue;
		refs.push_back(StreamRef{static_cast<const DevRay*>(rays[i]), static_cast<float4*>(results[i]), total, counts[i]});
		total += counts[i];
	}
	uint32_t cursor = 0;
	if (!rc && total) {
		TraceParams p{};
		p.nodes = static_cast<const float4*>(nodes); p.pairs = static_cast<const float4*>(pairs); p.remap = remap;
		p.env = static_cast<const float4*>(env); p.envWidth = envWidth; p.envHeight = envHeight; p.nodeCount = nodeCount;
		p.streams = refs.data(); p.nstreams = (uint32_t)refs.size(); p.total = total; p.single = refs[0];
		p.cursor = &cursor; p.counters = counters4; p.smemNodes = 0;
		p.tnodes = tnodes; p.tpairs = tpairs; p.perm = perm; p.envPairs = envPairs;
		Tuning t;
		t.blockThreads = tuning[0]; t.ctasPerSm = tuning[1]; t.fetchThreshold = tuning[2]; t.leafBail = tuning[3]; t.innerBail = tuning[4];
		t.smemStack = tuning[5];
		rc = launchTracePacked(p, t, counterMode, smCount, nullptr, &launches);
	}
	free(tnodes); free(tpairs); free(envPairs);
	return rc;
}

} // extern "C"
