// trace_on_cpu.cpp -- TEST INFRASTRUCTURE ONLY. C entry point over the default traversal kernel's launchers
// (rayaccel_b200/csrc/traverse_packed.cu: packNodesKernel, packPairsKernel, packEnvKernel, tracePackedKernel) compiled for
// the CPU over tests/harness/cuda_on_cpu/cuda_runtime.h after the two mechanical rewrites of cuda_on_cpu/rewrite.py
// (kernel launches, inline PTX). tests/test_kernels_on_cpu.py compares its results with the checker bit for bit.
#include "engine.h"

#include <pmmintrin.h>
#include <xmmintrin.h>

#include <algorithm>
#include <map>
#include <tuple>
#include <vector>

using namespace racc_b200;

// ---- L1 gather model (tests/harness/l1_model.py) -------------------------------------------------------------------------------
// profiles/r01_l1_wavefront_microbench.md: a divergent 256-bit load costs the L1 data pipe about 1.07 wavefronts per
// DISTINCT 32-byte sector it touches and never less than 0.27 per participating lane (register write-back), one
// wavefront per clock per SM. The sink below groups the traced loads of a launch into warp-level instruction instances
// (same block, warp, vote epoch and load ordinal within the epoch) and sums that cost.
namespace {
struct L1Model {
	std::map<std::tuple<uint32_t, uint32_t, uint32_t, uint32_t>, std::vector<unsigned long long>> instances; // address | lane in the low 5 bits
	uint32_t epochOf[32][32];
	uint32_t ordinal[32][32];
	uint32_t block = 0xffffffffu;
	double wavefronts = 0, laneLoads = 0, count = 0, wavefrontsByQuarter = 0, wavefrontsByFour = 0, wavefrontsByPair = 0;
	static double price(std::vector<unsigned long long>& a) {
		if (a.empty()) return 0.0;
		const double lanes = (double)a.size();
		std::sort(a.begin(), a.end());
		const double d = (double)(std::unique(a.begin(), a.end()) - a.begin());
		return std::max(0.266 * lanes, 1.065 * d);
	}
	void flush() {
		for (auto& kv : instances) {
			std::vector<unsigned long long> all, quarter[4];
			for (unsigned long long v : kv.second) {
				all.push_back(v & ~31ull);
				quarter[(v & 31u) / 8].push_back(v & ~31ull);
			}
			laneLoads += (double)all.size();
			wavefronts += price(all);
			for (int q = 0; q < 4; ++q) wavefrontsByQuarter += price(quarter[q]);
			std::vector<unsigned long long> four[8], pair[16];
			for (unsigned long long v : kv.second) {
				four[(v & 31u) / 4].push_back(v & ~31ull);
				pair[(v & 31u) / 2].push_back(v & ~31ull);
			}
			for (int q = 0; q < 8; ++q) wavefrontsByFour += price(four[q]);
			for (int q = 0; q < 16; ++q) wavefrontsByPair += price(pair[q]);
			count += 1;
		}
		instances.clear();
	}
} g_model;

void modelSink(uint32_t block, uint32_t warp, uint32_t epoch, uint32_t lane, unsigned long long address, uint32_t) {
	if (block != g_model.block) {
		g_model.flush();
		g_model.block = block;
		memset(g_model.epochOf, 0xff, sizeof(g_model.epochOf));
	}
	if (g_model.epochOf[warp][lane] != epoch) { g_model.epochOf[warp][lane] = epoch; g_model.ordinal[warp][lane] = 0; }
	g_model.instances[std::make_tuple(warp, epoch, g_model.ordinal[warp][lane]++, 0u)].push_back((address & ~31ull) | lane);
	if (g_model.instances.size() > (1u << 20)) g_model.flush(); // epochs only grow: old instances are complete
}
} // namespace

extern "C" void cpu_l1_model_begin() {
	g_model = L1Model();
	::cuda_on_cpu::load_sink = modelSink;
}
// out = {modelled wavefronts (sharing across the whole warp), lane-level loads, warp-level instruction instances,
//        modelled wavefronts when only lanes of the same quarter-warp can share a sector}
extern "C" void cpu_l1_model_end(double* out4) { // out4 has room for 6 values: ..., sharing within 4 lanes, within 2 lanes
	::cuda_on_cpu::load_sink = nullptr;
	g_model.flush();
	out4[0] = g_model.wavefronts; out4[1] = g_model.laneLoads; out4[2] = g_model.count; out4[3] = g_model.wavefrontsByQuarter; out4[4] = g_model.wavefrontsByFour; out4[5] = g_model.wavefrontsByPair;
}

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
