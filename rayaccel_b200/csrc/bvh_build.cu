// bvh_build.cu -- the full-sweep SAH BVH2 build on the GPU.
//
// WHAT it builds is fixed by the reference builder (/root/reference/RayAccelerator/Bvh2.cpp:257-535,
// 772-907) as restated by the host builder in scene_build.cpp: the same tree, decision for decision --
// the same per-axis sweep with its 8-wide block / scalar-tail arithmetic split (Bvh2.cpp:339,354-357,
// 405,438-450), the same pruning against the best cost so far (:346-351,417-418), the same tie rules,
// the same leaf-cost test through RCPSS (:462-467), the same forced median split at >= 127 triangles
// (:468-480) and the same stable three-list partition (:217-253). The parity test is byte equality of
// the resulting device images with the host build (tests/test_gpu_parity.py::test_device_build_*).
//
// HOW is GPU-first. The reference recurses depth-first with a task pool; here the tree grows level
// by level: nodes above 131 072 triangles are built one at a time by the WHOLE GPU (cooperative launch,
// grid-wide barriers between chunk-parallel phases), the others by ONE CTA PER NODE (512 threads x 8
// positions above 4096 triangles, 256 threads above 64, one warp below). The sequential sweep becomes
//   * prefix / suffix unions of the triangle boxes by chunked CTA-wide max-scans (max is exact and
//     associative, so every position sees the very box the sequential loop would hold),
//   * per-position cost in the arithmetic variant its position selects (block region vs scalar tail),
//   * the order-dependent part -- running best, first-lane-wins minima, early termination -- replayed
//     by one thread over per-8-block summaries (min, argmin, max right cost), a few dozen steps per chunk.
// Node slots follow the host builder's deterministic numbering (subtree over n triangles owns
// 2n-1 consecutive slots), so no atomic counter decides the layout. Triangle centres are sorted with
// the radix sort of raysort.cu (stable; ties broken like the reference's packed key order).
#include "engine.h"
#include "scene_build.h"

#include <cooperative_groups.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>
#include <vector>

namespace racc_b200 {
namespace {

constexpr uint32_t kNone = 0xffffffffu;
constexpr uint32_t kSmallNode = 64;  // nodes up to this many triangles are built by one warp
constexpr uint32_t kWideNode = 4096; // nodes above this are built by 512 threads x 8 positions each
constexpr int kWideThreads = 512;
constexpr uint32_t kHugeNode = 131072; // nodes above this are built by the whole GPU, one node at a time (cooperative launch)

__constant__ float c_rcpTable[2048];

struct DevBuild {
	uint32_t n;
	const float4* tb;   // per triangle: {-min.xyzw}, {max.xyzw}
	uint32_t* sorted[3];
	uint32_t* scratch;
	float* leftSah;
	uint8_t* goesLeft;
	BuildNode* nodes;
};

// Work lists of the next level, one per node class (index: 0 mid, 1 small, 2 wide, 3 huge).
struct NextLists {
	uint32_t* list[4];
	uint32_t* counts;
};
__device__ __forceinline__ int nodeClass(uint32_t triangles);
__device__ __forceinline__ void emitChild(const NextLists& nx, uint32_t slot, uint32_t triangles) {
	const int c = nodeClass(triangles);
	nx.list[c][atomicAdd(&nx.counts[c], 1u)] = slot;
}

struct Box6 { float v[6]; }; // -min.xyz, max.xyz

__device__ __forceinline__ Box6 emptyBox() {
	Box6 b;
#pragma unroll
	for (int k = 0; k < 6; ++k) b.v[k] = -INFINITY;
	return b;
}
__device__ __forceinline__ Box6 loadBox(const float4* tb, uint32_t tri) {
	const float4 a = __ldg(tb + 2 * (size_t)tri), c = __ldg(tb + 2 * (size_t)tri + 1);
	Box6 b;
	b.v[0] = a.x; b.v[1] = a.y; b.v[2] = a.z; b.v[3] = c.x; b.v[4] = c.y; b.v[5] = c.z;
	return b;
}
__device__ __forceinline__ void boxMax(Box6& a, const Box6& b) {
#pragma unroll
	for (int k = 0; k < 6; ++k) a.v[k] = fmaxf(a.v[k], b.v[k]);
}
// Bvh2.cpp:76-80 surfaceArea(): (d0*d1 + d1*d2) + d0*d2 with d = (-min) + max
__device__ __forceinline__ float areaScalar(const Box6& b) {
	const float d0 = __fadd_rn(b.v[0], b.v[3]), d1 = __fadd_rn(b.v[1], b.v[4]), d2 = __fadd_rn(b.v[2], b.v[5]);
	return __fadd_rn(__fadd_rn(__fmul_rn(d0, d1), __fmul_rn(d1, d2)), __fmul_rn(d0, d2));
}
// Bvh2.cpp:339,405: the 8-wide blocks use fma(d0,d1, fma(d0,d2, d1*d2))
__device__ __forceinline__ float areaBlock(const Box6& b) {
	const float d0 = __fadd_rn(b.v[0], b.v[3]), d1 = __fadd_rn(b.v[1], b.v[4]), d2 = __fadd_rn(b.v[2], b.v[5]);
	return __fmaf_rn(d0, d1, __fmaf_rn(d0, d2, __fmul_rn(d1, d2)));
}
// _mm_rcp_ss of the host CPU: table on the top 11 mantissa bits, exact power-of-two scaling
__device__ __forceinline__ float rcpSS(float x) {
	const uint32_t u = __float_as_uint(x), e = (u >> 23) & 0xffu, m = u & 0x7fffffu;
	if (e < 3u || e > 251u) return __frcp_rn(x);
	return __fmul_rn(c_rcpTable[m >> 12], __uint_as_float((254u - e) << 23));
}

// Inclusive max-scan of one box per thread in thread order; `total` = union over the CTA.
// sWarp: [T/32][6] floats of shared memory (unused when T == 32). Contains __syncthreads for T > 32.
template <int T>
__device__ __forceinline__ void ctaScanMax(Box6& b, Box6& total, float (*sWarp)[6]) {
	const unsigned lane = threadIdx.x & 31;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
		for (int k = 0; k < 6; ++k) {
			const float v = __shfl_up_sync(0xffffffffu, b.v[k], o);
			if ((int)lane >= o) b.v[k] = fmaxf(b.v[k], v);
		}
	}
	if (T == 32) {
#pragma unroll
		for (int k = 0; k < 6; ++k) total.v[k] = __shfl_sync(0xffffffffu, b.v[k], 31);
		return;
	}
	const unsigned warp = threadIdx.x >> 5;
	__syncthreads(); // previous users of sWarp are done
	if (lane == 31) {
#pragma unroll
		for (int k = 0; k < 6; ++k) sWarp[warp][k] = b.v[k];
	}
	__syncthreads();
	total = emptyBox();
	for (unsigned w = 0; w < T / 32; ++w) {
		Box6 x;
#pragma unroll
		for (int k = 0; k < 6; ++k) x.v[k] = sWarp[w][k];
		if (w < warp) boxMax(b, x);
		boxMax(total, x);
	}
}

template <int T>
__device__ __forceinline__ Box6 ctaReduceMax(Box6 b, float (*sWarp)[6]) {
	Box6 total;
	ctaScanMax<T>(b, total, sWarp);
	return total;
}

// CTA-wide exclusive count of `flag` over threads in thread order; returns my rank, *total = CTA sum.
template <int T>
__device__ __forceinline__ uint32_t ctaRank(bool flag, uint32_t* total, uint32_t* sCount) {
	const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const unsigned ballot = __ballot_sync(0xffffffffu, flag);
	const uint32_t inWarp = __popc(ballot & ((1u << lane) - 1u));
	if (T == 32) {
		*total = __popc(ballot);
		return inWarp;
	}
	__syncthreads();
	if (lane == 0) sCount[warp] = __popc(ballot);
	__syncthreads();
	uint32_t before = 0, all = 0;
	for (unsigned w = 0; w < T / 32; ++w) {
		const uint32_t c = sCount[w];
		if (w < warp) before += c;
		all += c;
	}
	*total = all;
	return before + inWarp;
}

template <int T>
struct Shared {
	float warpBox[T / 32][6];
	uint32_t warpCount[T / 32];
	float sah[T];
	float blkMin[T / 8];
	float blkMaxR[T / 8];
	int blkArg[T / 8];
	uint32_t prune;
	float best;
	uint32_t pivot;
	int done;
};

// One axis of the sweep (scene_build.cpp sweepAxis / Bvh2.cpp:287-460), CTA-parallel. Updates bestSah
// (uniform across the CTA) and returns the best pivot found on this axis, kNone if none beat bestSah.
template <int T>
__device__ uint32_t sweepAxis(const DevBuild& ctx, const uint32_t* __restrict__ ids, uint32_t first, uint32_t last, float& bestSah, Shared<T>& sh) {
	const uint32_t tid = threadIdx.x;
	const uint32_t count = last - first;
	const uint32_t nb = count > 8 ? (count - 1) / 8 : 0; // iterations of `for (i = first; i < last - 8; i += 8)`
	const uint32_t blockEnd = first + 8 * nb;

	// ---- left to right: leftSah[k] = area(union ids[first..k]) * (k - first + 1), k <= last - 2
	if (tid == 0) sh.prune = kNone;
	__syncthreads();
	Box6 carry = emptyBox();
	uint32_t prune = kNone;
	for (uint32_t base = first; base + 1 < last; base += T) {
		const uint32_t pos = base + tid;
		const bool valid = pos + 1 < last;
		Box6 b = valid ? loadBox(ctx.tb, ids[pos]) : emptyBox();
		Box6 total;
		ctaScanMax<T>(b, total, sh.warpBox);
		boxMax(b, carry);
		boxMax(carry, total);
		if (valid) {
			const bool inBlock = pos < blockEnd;
			const float sah = __fmul_rn(inBlock ? areaBlock(b) : areaScalar(b), (float)(pos - first + 1));
			ctx.leftSah[pos] = sah;
			if (inBlock && sah > bestSah) atomicMin(&sh.prune, pos); // Bvh2.cpp:346-351: first position that is worse
		}
		__syncthreads();
		prune = sh.prune;
		if (prune != kNone) break;
	}

	// ---- right to left from i0: pivot q splits into [first,q) | [q,last)
	const uint32_t i0 = prune != kNone ? prune : last - 1;
	Box6 run;
	if (prune == kNone) {
		run = loadBox(ctx.tb, ids[last - 1]);
	}
	else {
		Box6 acc = emptyBox();
		for (uint32_t pos = i0 + tid; pos < last; pos += T) boxMax(acc, loadBox(ctx.tb, ids[pos]));
		run = ctaReduceMax<T>(acc, sh.warpBox);
	}
	const uint32_t span = i0 - first;      // pivots q = i0 - o, o in [0, span)
	const uint32_t blockSpan = (span / 8) * 8; // offsets handled by `for (; i > first + 7; i -= 8)`
	uint32_t bestPivot = kNone;
	bool done = false;
	carry = run;
	for (uint32_t off = 0; off < span && !done; off += T) {
		const uint32_t o = off + tid;
		const bool valid = o < span;
		const uint32_t q = i0 - o;
		Box6 b = valid ? loadBox(ctx.tb, ids[q]) : emptyBox();
		Box6 total;
		ctaScanMax<T>(b, total, sh.warpBox);
		boxMax(b, carry);
		boxMax(carry, total);
		const bool inBlock = o < blockSpan;
		float rightSah = 0.0f, sah = INFINITY;
		if (valid) {
			rightSah = __fmul_rn(inBlock ? areaBlock(b) : areaScalar(b), (float)(last - q));
			sah = __fadd_rn(ctx.leftSah[q - 1], rightSah);
		}
		// summaries of each 8-position block: minimum (first lane wins ties), its lane, max right cost
		float m = (valid && inBlock) ? sah : INFINITY;
		float mr = (valid && inBlock) ? rightSah : -INFINITY;
		int arg = (int)(tid & 7u);
#pragma unroll
		for (int d = 1; d < 8; d <<= 1) {
			const float om = __shfl_xor_sync(0xffffffffu, m, d);
			const int oa = __shfl_xor_sync(0xffffffffu, arg, d);
			const float omr = __shfl_xor_sync(0xffffffffu, mr, d);
			if (om < m || (om == m && oa < arg)) { m = om; arg = oa; }
			mr = fmaxf(mr, omr);
		}
		__syncthreads();
		if ((tid & 7u) == 0) {
			sh.blkMin[tid >> 3] = m;
			sh.blkArg[tid >> 3] = arg;
			sh.blkMaxR[tid >> 3] = mr;
		}
		sh.sah[tid] = sah;
		__syncthreads();
		if (tid == 0) {
			// the order-dependent part, replayed sequentially over the chunk's blocks
			float best = bestSah;
			uint32_t bp = bestPivot;
			int fin = 0;
			for (uint32_t g = 0; g < T / 8; ++g) {
				const uint32_t ob = off + 8 * g;
				if (ob >= span) break;
				if (ob >= blockSpan) {
					// scalar tail (Bvh2.cpp:438-450): at most 7 pivots, strict improvement only
					for (uint32_t j = 0; j < 8 && ob + j < span; ++j) {
						const float s = sh.sah[8 * g + j];
						if (s < best) { best = s; bp = i0 - (ob + j); }
					}
					fin = 1;
					break;
				}
				const float mb = sh.blkMin[g];
				const bool better = mb < best;            // Bvh2.cpp:417
				const bool worse = sh.blkMaxR[g] > best;  // Bvh2.cpp:418, against the not-yet-updated best
				if (mb < best) best = mb;
				if (better) bp = (i0 - ob) - (uint32_t)sh.blkArg[g]; // Bvh2.cpp:428-429
				if (worse) { fin = 1; break; }
			}
			sh.best = best;
			sh.pivot = bp;
			sh.done = fin;
		}
		__syncthreads();
		bestSah = sh.best;
		bestPivot = sh.pivot;
		done = sh.done != 0;
	}
	__syncthreads();
	return bestPivot;
}

// Stable split of one sorted list by goesLeft[] (Bvh2.cpp:217-240): lefts keep their order in
// [first,pivot), rights in [pivot,last).
template <int T>
__device__ void splitList(const DevBuild& ctx, uint32_t* __restrict__ ids, uint32_t first, uint32_t pivot, uint32_t last, Shared<T>& sh) {
	const uint32_t tid = threadIdx.x;
	uint32_t nL = 0, nR = 0;
	for (uint32_t base = first; base < last; base += T) {
		const uint32_t pos = base + tid;
		const bool valid = pos < last;
		const uint32_t t = valid ? ids[pos] : 0u;
		const bool f = valid && ctx.goesLeft[t] != 0;
		uint32_t totalL;
		const uint32_t rankL = ctaRank<T>(f, &totalL, sh.warpCount);
		const uint32_t chunk = min((uint32_t)T, last - base);
		if (valid) ctx.scratch[f ? first + nL + rankL : pivot + nR + (tid - rankL)] = t;
		nL += totalL;
		nR += chunk - totalL;
	}
	__syncthreads();
	for (uint32_t pos = first + tid; pos < last; pos += T) ids[pos] = ctx.scratch[pos];
	__syncthreads();
}


__device__ __forceinline__ int nodeClass(uint32_t triangles) {
	return triangles > kHugeNode ? 3 : (triangles > kWideNode ? 2 : (triangles > kSmallNode ? 0 : 1));
}

// ---- wide variants: every thread owns 8 consecutive positions = exactly one 8-block of the sweep -----

// Exclusive max-scan of one box per thread: returns the union of all earlier threads' boxes.
template <int T>
__device__ __forceinline__ Box6 ctaExclusiveScanMax(Box6 b, Box6& total, float (*sWarp)[6]) {
	const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
		for (int k = 0; k < 6; ++k) {
			const float v = __shfl_up_sync(0xffffffffu, b.v[k], o);
			if ((int)lane >= o) b.v[k] = fmaxf(b.v[k], v);
		}
	}
	Box6 ex;
#pragma unroll
	for (int k = 0; k < 6; ++k) {
		const float v = __shfl_up_sync(0xffffffffu, b.v[k], 1);
		ex.v[k] = lane ? v : -INFINITY;
	}
	__syncthreads();
	if (lane == 31) {
#pragma unroll
		for (int k = 0; k < 6; ++k) sWarp[warp][k] = b.v[k];
	}
	__syncthreads();
	total = emptyBox();
	for (unsigned w = 0; w < T / 32; ++w) {
		Box6 x;
#pragma unroll
		for (int k = 0; k < 6; ++k) x.v[k] = sWarp[w][k];
		if (w < warp) boxMax(ex, x);
		boxMax(total, x);
	}
	return ex;
}

// Exclusive min-scan of one float per thread (identity +inf).
template <int T>
__device__ __forceinline__ float ctaExclusiveScanMin(float v, float* sWarp) {
	const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const float x = __shfl_up_sync(0xffffffffu, v, o);
		if ((int)lane >= o) v = fminf(v, x);
	}
	float ex = __shfl_up_sync(0xffffffffu, v, 1);
	if (!lane) ex = INFINITY;
	__syncthreads();
	if (lane == 31) sWarp[warp] = v;
	__syncthreads();
	for (unsigned w = 0; w < warp; ++w) ex = fminf(ex, sWarp[w]);
	return ex;
}

// Exclusive sum-scan of one count per thread.
template <int T>
__device__ __forceinline__ uint32_t ctaExclusiveScanSum(uint32_t v, uint32_t* total, uint32_t* sWarp) {
	const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t incl = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
		if ((int)lane >= o) incl += x;
	}
	__syncthreads();
	if (lane == 31) sWarp[warp] = incl;
	__syncthreads();
	uint32_t before = 0, all = 0;
	for (unsigned w = 0; w < T / 32; ++w) {
		const uint32_t c = sWarp[w];
		if (w < warp) before += c;
		all += c;
	}
	*total = all;
	return before + incl - v;
}

template <int T>
struct SharedWide {
	float warpBox[T / 32][6];
	float warpMin[T / 32];
	uint32_t warpCount[T / 32];
	uint32_t prune;
	uint32_t term;
	uint32_t lastBetter;
	float best;
	uint32_t pivot;
	int done;
};

// sweepAxis with 8 positions per thread. Chunks start at `first` (forward) / at offset 0 from i0
// (backward), so a thread's 8 positions are one block of the reference's 8-wide loops, and the
// order-dependent replay runs in parallel: running best = exclusive prefix-min over the block minima,
// termination = first block whose largest right cost exceeds it, pivot = last block at or before
// the terminating one that strictly improves the running best.
template <int T>
__device__ uint32_t sweepAxisWide(const DevBuild& ctx, const uint32_t* __restrict__ ids, uint32_t first, uint32_t last, float& bestSah, SharedWide<T>& sh) {
	constexpr uint32_t C = T * 8;
	const uint32_t tid = threadIdx.x;
	const uint32_t count = last - first;
	const uint32_t nb = count > 8 ? (count - 1) / 8 : 0;
	const uint32_t blockEnd = first + 8 * nb;

	if (tid == 0) sh.prune = kNone;
	__syncthreads();
	Box6 carry = emptyBox();
	uint32_t prune = kNone;
	for (uint32_t base = first; base + 1 < last; base += C) {
		const uint32_t p0 = base + 8 * tid;
		Box6 run[8];
		Box6 acc = emptyBox();
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			if (p0 + j + 1 < last) boxMax(acc, loadBox(ctx.tb, ids[p0 + j]));
			run[j] = acc;
		}
		Box6 total;
		Box6 pre = ctaExclusiveScanMax<T>(acc, total, sh.warpBox);
		boxMax(pre, carry);
		boxMax(carry, total);
		const bool inBlock = p0 < blockEnd;
		uint32_t myPrune = kNone;
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			const uint32_t pos = p0 + j;
			if (pos + 1 < last) {
				Box6 b = run[j];
				boxMax(b, pre);
				const float sah = __fmul_rn(inBlock ? areaBlock(b) : areaScalar(b), (float)(pos - first + 1));
				ctx.leftSah[pos] = sah;
				if (inBlock && sah > bestSah && myPrune == kNone) myPrune = pos;
			}
		}
		if (myPrune != kNone) atomicMin(&sh.prune, myPrune);
		__syncthreads();
		prune = sh.prune;
		if (prune != kNone) break;
	}

	const uint32_t i0 = prune != kNone ? prune : last - 1;
	Box6 run0;
	if (prune == kNone) {
		run0 = loadBox(ctx.tb, ids[last - 1]);
	}
	else {
		Box6 acc = emptyBox();
		for (uint32_t pos = i0 + tid; pos < last; pos += T) boxMax(acc, loadBox(ctx.tb, ids[pos]));
		run0 = ctaReduceMax<T>(acc, sh.warpBox);
	}
	const uint32_t span = i0 - first;
	const uint32_t blockSpan = (span / 8) * 8;
	uint32_t bestPivot = kNone;
	bool done = false;
	carry = run0;
	for (uint32_t off = 0; off < span && !done; off += C) {
		const uint32_t o0 = off + 8 * tid;
		Box6 run[8];
		Box6 acc = emptyBox();
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			if (o0 + j < span) boxMax(acc, loadBox(ctx.tb, ids[i0 - (o0 + j)]));
			run[j] = acc;
		}
		Box6 total;
		Box6 pre = ctaExclusiveScanMax<T>(acc, total, sh.warpBox);
		boxMax(pre, carry);
		boxMax(carry, total);
		const bool isBlock = o0 < blockSpan;          // 8 valid pivots of the 8-wide loop
		const bool isTail = !isBlock && o0 < span;    // the one thread holding the <= 7 scalar-tail pivots
		float sah[8];
		float m = INFINITY, mr = -INFINITY;
		int arg = 0;
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			sah[j] = INFINITY;
			if (o0 + j < span) {
				const uint32_t q = i0 - (o0 + j);
				Box6 b = run[j];
				boxMax(b, pre);
				const float r = __fmul_rn(isBlock ? areaBlock(b) : areaScalar(b), (float)(last - q));
				const float sv = __fadd_rn(ctx.leftSah[q - 1], r);
				sah[j] = sv;
				if (isBlock) {
					if (sv < m) { m = sv; arg = j; } // first lane holding the minimum
					mr = fmaxf(mr, r);
				}
			}
		}
		if (tid == 0) { sh.term = kNone; sh.lastBetter = 0; sh.best = bestSah; sh.pivot = bestPivot; }
		const float before = fminf(ctaExclusiveScanMin<T>(m, sh.warpMin), bestSah); // running best when this block starts
		if (isBlock && mr > before) atomicMin(&sh.term, tid); // Bvh2.cpp:418
		__syncthreads();
		const uint32_t term = sh.term;
		if (isBlock && m < before && tid <= term) atomicMax(&sh.lastBetter, tid + 1); // Bvh2.cpp:417,428
		__syncthreads();
		const uint32_t lb = sh.lastBetter;
		if (lb && tid == lb - 1) { sh.best = m; sh.pivot = (i0 - o0) - (uint32_t)arg; }
		__syncthreads();
		if (isTail && term == kNone) {
			// scalar tail (Bvh2.cpp:438-450): strict improvement, in order
			float best = before;
			uint32_t bp = sh.pivot;
			bool changed = false;
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				if (o0 + j < span && sah[j] < best) { best = sah[j]; bp = i0 - (o0 + j); changed = true; }
			}
			if (changed) { sh.best = best; sh.pivot = bp; }
		}
		if (tid == 0) sh.done = (term != kNone || off + C >= span) ? 1 : 0;
		__syncthreads();
		bestSah = sh.best;
		bestPivot = sh.pivot;
		done = sh.done != 0;
		__syncthreads();
	}
	__syncthreads();
	return bestPivot;
}

template <int T>
__device__ void splitListWide(const DevBuild& ctx, uint32_t* __restrict__ ids, uint32_t first, uint32_t pivot, uint32_t last, SharedWide<T>& sh) {
	constexpr uint32_t C = T * 8;
	const uint32_t tid = threadIdx.x;
	uint32_t nL = 0, nR = 0;
	for (uint32_t base = first; base < last; base += C) {
		const uint32_t p0 = base + 8 * tid;
		uint32_t t[8];
		unsigned flags = 0;
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			t[j] = 0;
			if (p0 + j < last) {
				t[j] = ids[p0 + j];
				if (ctx.goesLeft[t[j]]) flags |= 1u << j;
			}
		}
		uint32_t totalL;
		const uint32_t exL = ctaExclusiveScanSum<T>(__popc(flags), &totalL, sh.warpCount);
		const uint32_t chunk = min(C, last - base);
		uint32_t dl = first + nL + exL;
		uint32_t dr = pivot + nR + (8 * tid - exL);
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			if (p0 + j < last) {
				if (flags & (1u << j)) ctx.scratch[dl++] = t[j];
				else ctx.scratch[dr++] = t[j];
			}
		}
		nL += totalL;
		nR += chunk - totalL;
	}
	__syncthreads();
	for (uint32_t pos = first + tid; pos < last; pos += T) ids[pos] = ctx.scratch[pos];
	__syncthreads();
}


// ---- huge nodes: the whole GPU on one node ------------------------------------------------------------
// With one CTA per node the first levels of a large tree run on 1, 2, 4 ... SMs. Nodes above kHugeNode
// triangles are instead built one after the other by ALL resident CTAs of a cooperative launch. The sweep is
// the wide one (8 positions per thread = one 8-block, 4096 positions per chunk), cut in chunk-parallel phases
// separated by grid-wide barriers: chunk unions -> scan of the chunk unions by CTA 0 -> chunks again with their
// carry; the order-dependent replay runs on per-block summaries kept in global memory (prefix-min over chunk
// minima by CTA 0, then first terminating block / last improving block by atomics). Same values at every
// position as the sequential sweep, so the same tree.
namespace cg = cooperative_groups;

struct CoopScratch {
	float* chunkBox;      // [chunks][6] union of each chunk
	float* chunkCarry;    // [chunks][6] union of everything before each chunk (incl. the initial box)
	float* chunkMin;      // [chunks] minimum block cost of each chunk
	float* chunkBefore;   // [chunks] running best when each chunk starts
	uint32_t* chunkCount; // [2][chunks] lefts per chunk of the two lists being partitioned, then their prefix
	float* partBox;       // [gridDim][6]
	float* blkM;          // per 8-block: minimum cost
	float* blkMr;         //              maximum right cost
	float* blkBefore;     //              running best when the block starts
	int* blkArg;          //              lane of the minimum
	float* tailSah;       // [8] costs of the scalar-tail pivots
	uint32_t* scal;       // 0 prune, 1 term, 2 lastBetter, 3 pivot, 4 best (float bits)
	uint32_t maxChunks;
};

template <int T>
__device__ Box6 gridUnion(cg::grid_group& grid, Box6 acc, const CoopScratch& sc, float (*sWarp)[6]) {
	const Box6 mine = ctaReduceMax<T>(acc, sWarp);
	if (threadIdx.x == 0) {
#pragma unroll
		for (int k = 0; k < 6; ++k) sc.partBox[6 * blockIdx.x + k] = mine.v[k];
	}
	grid.sync();
	Box6 all = emptyBox();
	for (unsigned g = 0; g < gridDim.x; ++g) {
#pragma unroll
		for (int k = 0; k < 6; ++k) all.v[k] = fmaxf(all.v[k], sc.partBox[6 * g + k]);
	}
	grid.sync();
	return all;
}

// CTA 0: chunkCarry[c] = init U chunkBox[0..c)
template <int T>
__device__ void scanChunkBoxes(const CoopScratch& sc, uint32_t chunks, Box6 init, float (*sWarp)[6]) {
	if (blockIdx.x != 0)
		return;
	Box6 carry = init;
	for (uint32_t base = 0; base < chunks; base += T) {
		const uint32_t c = base + threadIdx.x;
		Box6 b = emptyBox();
		if (c < chunks) {
#pragma unroll
			for (int k = 0; k < 6; ++k) b.v[k] = sc.chunkBox[6 * (size_t)c + k];
		}
		Box6 total;
		Box6 ex = ctaExclusiveScanMax<T>(b, total, sWarp);
		boxMax(ex, carry);
		boxMax(carry, total);
		if (c < chunks) {
#pragma unroll
			for (int k = 0; k < 6; ++k) sc.chunkCarry[6 * (size_t)c + k] = ex.v[k];
		}
	}
}
// CTA 0: chunkBefore[c] = min(init, chunkMin[0..c)); returns min over everything (uniform in CTA 0)
template <int T>
__device__ void scanChunkMins(const CoopScratch& sc, uint32_t chunks, float init, float* sWarpMin) {
	if (blockIdx.x != 0)
		return;
	__shared__ float carryS;
	if (threadIdx.x == 0) carryS = init;
	__syncthreads();
	for (uint32_t base = 0; base < chunks; base += T) {
		const uint32_t c = base + threadIdx.x;
		const float v = c < chunks ? sc.chunkMin[c] : INFINITY;
		const float carry = carryS;
		const float ex = fminf(ctaExclusiveScanMin<T>(v, sWarpMin), carry);
		if (c < chunks) sc.chunkBefore[c] = ex;
		__syncthreads();
		if (threadIdx.x == T - 1) carryS = fminf(ex, v);
		__syncthreads();
	}
}
// CTA 0: in-place exclusive prefix sum of counts[0..chunks)
template <int T>
__device__ void scanChunkCounts(uint32_t* counts, uint32_t chunks, uint32_t* sWarp) {
	if (blockIdx.x != 0)
		return;
	__shared__ uint32_t carryS;
	if (threadIdx.x == 0) carryS = 0;
	__syncthreads();
	for (uint32_t base = 0; base < chunks; base += T) {
		const uint32_t c = base + threadIdx.x;
		const uint32_t v = c < chunks ? counts[c] : 0u;
		uint32_t total;
		const uint32_t ex = ctaExclusiveScanSum<T>(v, &total, sWarp);
		const uint32_t carry = carryS;
		if (c < chunks) counts[c] = carry + ex;
		__syncthreads();
		if (threadIdx.x == 0) carryS = carry + total;
		__syncthreads();
	}
}

// One axis of the sweep with the whole grid. bestSah / returned pivot are uniform over the grid.
template <int T>
__device__ uint32_t sweepAxisCoop(cg::grid_group& grid, const DevBuild& ctx, const CoopScratch& sc, const uint32_t* __restrict__ ids,
                                  uint32_t first, uint32_t last, float& bestSah, SharedWide<T>& sh) {
	constexpr uint32_t C = T * 8;
	const uint32_t tid = threadIdx.x;
	const uint32_t count = last - first;
	const uint32_t nb = count > 8 ? (count - 1) / 8 : 0;
	const uint32_t blockEnd = first + 8 * nb;
	const uint32_t chunksF = (count - 1 + C - 1) / C; // positions first .. last-2

	// ---- forward ------------------------------------------------------------------------------------
	for (uint32_t c = blockIdx.x; c < chunksF; c += gridDim.x) {
		const uint32_t p0 = first + c * C + 8 * tid;
		Box6 acc = emptyBox();
#pragma unroll
		for (int j = 0; j < 8; ++j)
			if (p0 + j + 1 < last) boxMax(acc, loadBox(ctx.tb, ids[p0 + j]));
		const Box6 u = ctaReduceMax<T>(acc, sh.warpBox);
		if (tid == 0) {
#pragma unroll
			for (int k = 0; k < 6; ++k) sc.chunkBox[6 * (size_t)c + k] = u.v[k];
		}
	}
	if (blockIdx.x == 0 && tid == 0) { sc.scal[0] = kNone; sc.scal[1] = kNone; sc.scal[2] = 0u; }
	grid.sync();
	scanChunkBoxes<T>(sc, chunksF, emptyBox(), sh.warpBox);
	grid.sync();
	for (uint32_t c = blockIdx.x; c < chunksF; c += gridDim.x) {
		const uint32_t p0 = first + c * C + 8 * tid;
		Box6 run[8];
		Box6 acc = emptyBox();
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			if (p0 + j + 1 < last) boxMax(acc, loadBox(ctx.tb, ids[p0 + j]));
			run[j] = acc;
		}
		Box6 total;
		Box6 pre = ctaExclusiveScanMax<T>(acc, total, sh.warpBox);
		Box6 carry;
#pragma unroll
		for (int k = 0; k < 6; ++k) carry.v[k] = sc.chunkCarry[6 * (size_t)c + k];
		boxMax(pre, carry);
		const bool inBlock = p0 < blockEnd;
		uint32_t myPrune = kNone;
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			const uint32_t pos = p0 + j;
			if (pos + 1 < last) {
				Box6 b = run[j];
				boxMax(b, pre);
				const float sah = __fmul_rn(inBlock ? areaBlock(b) : areaScalar(b), (float)(pos - first + 1));
				ctx.leftSah[pos] = sah;
				if (inBlock && sah > bestSah && myPrune == kNone) myPrune = pos;
			}
		}
		if (myPrune != kNone) atomicMin(&sc.scal[0], myPrune);
	}
	grid.sync();
	const uint32_t prune = sc.scal[0];

	// ---- backward -----------------------------------------------------------------------------------
	const uint32_t i0 = prune != kNone ? prune : last - 1;
	Box6 run0;
	if (prune == kNone) {
		run0 = loadBox(ctx.tb, ids[last - 1]);
	}
	else {
		Box6 acc = emptyBox();
		for (uint32_t pos = i0 + blockIdx.x * T + tid; pos < last; pos += gridDim.x * T) boxMax(acc, loadBox(ctx.tb, ids[pos]));
		run0 = gridUnion<T>(grid, acc, sc, sh.warpBox);
	}
	const uint32_t span = i0 - first;
	const uint32_t blockSpan = (span / 8) * 8;
	const uint32_t chunksB = (span + C - 1) / C;
	uint32_t bestPivot = kNone;
	if (span == 0) {
		grid.sync();
		return kNone;
	}
	for (uint32_t d = blockIdx.x; d < chunksB; d += gridDim.x) {
		const uint32_t o0 = d * C + 8 * tid;
		Box6 acc = emptyBox();
#pragma unroll
		for (int j = 0; j < 8; ++j)
			if (o0 + j < span) boxMax(acc, loadBox(ctx.tb, ids[i0 - (o0 + j)]));
		const Box6 u = ctaReduceMax<T>(acc, sh.warpBox);
		if (tid == 0) {
#pragma unroll
			for (int k = 0; k < 6; ++k) sc.chunkBox[6 * (size_t)d + k] = u.v[k];
		}
	}
	grid.sync();
	scanChunkBoxes<T>(sc, chunksB, run0, sh.warpBox);
	grid.sync();
	for (uint32_t d = blockIdx.x; d < chunksB; d += gridDim.x) {
		const uint32_t o0 = d * C + 8 * tid;
		Box6 run[8];
		Box6 acc = emptyBox();
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			if (o0 + j < span) boxMax(acc, loadBox(ctx.tb, ids[i0 - (o0 + j)]));
			run[j] = acc;
		}
		Box6 total;
		Box6 pre = ctaExclusiveScanMax<T>(acc, total, sh.warpBox);
		Box6 carry;
#pragma unroll
		for (int k = 0; k < 6; ++k) carry.v[k] = sc.chunkCarry[6 * (size_t)d + k];
		boxMax(pre, carry);
		const bool isBlock = o0 < blockSpan;
		const bool isTail = !isBlock && o0 < span;
		float m = INFINITY, mr = -INFINITY;
		int arg = 0;
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			float sv = INFINITY;
			if (o0 + j < span) {
				const uint32_t q = i0 - (o0 + j);
				Box6 b = run[j];
				boxMax(b, pre);
				const float r = __fmul_rn(isBlock ? areaBlock(b) : areaScalar(b), (float)(last - q));
				sv = __fadd_rn(ctx.leftSah[q - 1], r);
				if (isBlock) {
					if (sv < m) { m = sv; arg = j; }
					mr = fmaxf(mr, r);
				}
			}
			if (isTail) sc.tailSah[j] = sv;
		}
		const uint32_t gb = o0 / 8;
		if (o0 < span) {
			sc.blkM[gb] = m;
			sc.blkMr[gb] = mr;
			sc.blkArg[gb] = arg;
		}
		// minimum block cost of the chunk
		float cm = m;
#pragma unroll
		for (int o = 16; o; o >>= 1) cm = fminf(cm, __shfl_xor_sync(0xffffffffu, cm, o));
		__syncthreads();
		if ((tid & 31) == 0) sh.warpMin[tid >> 5] = cm;
		__syncthreads();
		if (tid == 0) {
			float x = INFINITY;
			for (int w = 0; w < T / 32; ++w) x = fminf(x, sh.warpMin[w]);
			sc.chunkMin[d] = x;
		}
	}
	grid.sync();
	scanChunkMins<T>(sc, chunksB, bestSah, sh.warpMin);
	grid.sync();
	for (uint32_t d = blockIdx.x; d < chunksB; d += gridDim.x) {
		const uint32_t o0 = d * C + 8 * tid;
		const uint32_t gb = o0 / 8;
		const bool isBlock = o0 < blockSpan;
		const float m = isBlock ? sc.blkM[gb] : INFINITY;
		const float before = fminf(ctaExclusiveScanMin<T>(m, sh.warpMin), sc.chunkBefore[d]);
		if (o0 < span) sc.blkBefore[gb] = before;
		if (isBlock && sc.blkMr[gb] > before) atomicMin(&sc.scal[1], gb);
	}
	grid.sync();
	const uint32_t term = sc.scal[1];
	for (uint32_t gb = blockIdx.x * T + tid; gb * 8 < blockSpan; gb += gridDim.x * T) {
		if (gb <= term && sc.blkM[gb] < sc.blkBefore[gb]) atomicMax(&sc.scal[2], gb + 1);
	}
	grid.sync();
	if (blockIdx.x == 0 && tid == 0) {
		const uint32_t lb = sc.scal[2];
		float best = bestSah;
		uint32_t bp = kNone;
		if (lb) {
			best = sc.blkM[lb - 1];
			bp = (i0 - 8 * (lb - 1)) - (uint32_t)sc.blkArg[lb - 1];
		}
		if (term == kNone && blockSpan < span) {
			// scalar tail: its running best is the minimum over all blocks
			const uint32_t tb = blockSpan / 8;
			float before = sc.blkBefore[tb];
			if (before < best) best = before; // (equal by construction when any block improved)
			for (uint32_t j = 0; j < 8 && blockSpan + j < span; ++j) {
				const float sv = sc.tailSah[j];
				if (sv < best) { best = sv; bp = i0 - (blockSpan + j); }
			}
		}
		sc.scal[3] = bp;
		sc.scal[4] = __float_as_uint(best);
	}
	grid.sync();
	bestPivot = sc.scal[3];
	bestSah = __uint_as_float(sc.scal[4]);
	grid.sync();
	return bestPivot;
}

template <int T>
__global__ void __launch_bounds__(T) buildHugeLevelKernel(const DevBuild ctx, const uint32_t* __restrict__ cur, uint32_t curCount, const NextLists next,
                                                         const CoopScratch sc) {
	cg::grid_group grid = cg::this_grid();
	__shared__ SharedWide<T> sh;
	constexpr uint32_t C = T * 8;
	const uint32_t tid = threadIdx.x;
	for (uint32_t item = 0; item < curCount; ++item) {
		const uint32_t slot = cur[item];
		BuildNode* node = ctx.nodes + slot;
		const uint32_t first = node->first, last = node->last;
		const uint32_t count = last - first;

		Box6 acc = emptyBox();
		for (uint32_t pos = first + blockIdx.x * T + tid; pos < last; pos += gridDim.x * T) boxMax(acc, loadBox(ctx.tb, ctx.sorted[0][pos]));
		const Box6 bounds = gridUnion<T>(grid, acc, sc, sh.warpBox);
		if (blockIdx.x == 0 && tid == 0) {
			node->bounds[0] = bounds.v[0]; node->bounds[1] = bounds.v[1]; node->bounds[2] = bounds.v[2]; node->bounds[3] = 0.0f;
			node->bounds[4] = bounds.v[3]; node->bounds[5] = bounds.v[4]; node->bounds[6] = bounds.v[5]; node->bounds[7] = 0.0f;
		}
		const float parentArea = areaScalar(bounds);
		uint32_t bestDim = kNone, pivot = kNone;
		bool split = false;
		if (parentArea > 0.0f) {
			float bestSah = INFINITY;
			for (uint32_t dim = 0; dim < 3; ++dim) {
				const uint32_t p = sweepAxisCoop<T>(grid, ctx, sc, ctx.sorted[dim], first, last, bestSah, sh);
				if (p != kNone) {
					pivot = p;
					bestDim = dim;
				}
			}
			const float cost = __fadd_rn(2.0f, __fmul_rn(__fmul_rn(1.0f, rcpSS(parentArea)), bestSah));
			split = !(cost > (float)(int)count) && pivot != kNone;
		}
		if (!split) { // count > kHugeNode >= 127: forced median split (Bvh2.cpp:468-471,478-480)
			bestDim = 0;
			pivot = (first + last) >> 1;
		}

		// partition the two other lists
		const uint32_t* ref = ctx.sorted[bestDim];
		for (uint32_t pos = first + blockIdx.x * T + tid; pos < last; pos += gridDim.x * T) ctx.goesLeft[ref[pos]] = pos < pivot ? 1 : 0;
		grid.sync();
		uint32_t* lists[2] = {ctx.sorted[(bestDim + 1) % 3], ctx.sorted[(bestDim + 2) % 3]};
		const uint32_t chunks = (count + C - 1) / C;
		for (int l = 0; l < 2; ++l) {
			uint32_t* cnt = sc.chunkCount + (size_t)l * sc.maxChunks;
			for (uint32_t c = blockIdx.x; c < chunks; c += gridDim.x) {
				const uint32_t p0 = first + c * C + 8 * tid;
				uint32_t mine = 0;
#pragma unroll
				for (int j = 0; j < 8; ++j)
					if (p0 + j < last && ctx.goesLeft[lists[l][p0 + j]]) ++mine;
				uint32_t total;
				ctaExclusiveScanSum<T>(mine, &total, sh.warpCount);
				if (tid == 0) cnt[c] = total;
			}
		}
		grid.sync();
		scanChunkCounts<T>(sc.chunkCount, chunks, sh.warpCount);
		scanChunkCounts<T>(sc.chunkCount + sc.maxChunks, chunks, sh.warpCount);
		grid.sync();
		for (int l = 0; l < 2; ++l) {
			const uint32_t* cnt = sc.chunkCount + (size_t)l * sc.maxChunks;
			uint32_t* out = l == 0 ? ctx.scratch : reinterpret_cast<uint32_t*>(ctx.leftSah); // leftSah is free again: second scratch
			for (uint32_t c = blockIdx.x; c < chunks; c += gridDim.x) {
				const uint32_t p0 = first + c * C + 8 * tid;
				uint32_t t[8];
				unsigned flags = 0;
#pragma unroll
				for (int j = 0; j < 8; ++j) {
					t[j] = 0;
					if (p0 + j < last) {
						t[j] = lists[l][p0 + j];
						if (ctx.goesLeft[t[j]]) flags |= 1u << j;
					}
				}
				uint32_t totalL;
				const uint32_t exL = ctaExclusiveScanSum<T>(__popc(flags), &totalL, sh.warpCount);
				const uint32_t nL = cnt[c];
				const uint32_t nR = c * C - nL;
				uint32_t dl = first + nL + exL;
				uint32_t dr = pivot + nR + (8 * tid - exL);
#pragma unroll
				for (int j = 0; j < 8; ++j) {
					if (p0 + j < last) {
						if (flags & (1u << j)) out[dl++] = t[j];
						else out[dr++] = t[j];
					}
				}
			}
		}
		grid.sync();
		for (int l = 0; l < 2; ++l) {
			const uint32_t* in = l == 0 ? ctx.scratch : reinterpret_cast<const uint32_t*>(ctx.leftSah);
			for (uint32_t pos = first + blockIdx.x * T + tid; pos < last; pos += gridDim.x * T) lists[l][pos] = in[pos];
		}
		if (blockIdx.x == 0 && tid == 0) {
			const uint32_t left = slot + 1;
			const uint32_t right = slot + 2 * (pivot - first);
			node->kind = bestDim + 1;
			node->left = left;
			node->right = right;
			BuildNode l{}, r{};
			l.kind = 0; l.parent = slot; l.first = first; l.last = pivot; l.left = kNone; l.right = kNone;
			r.kind = 0; r.parent = slot; r.first = pivot; r.last = last; r.left = kNone; r.right = kNone;
			ctx.nodes[left] = l;
			ctx.nodes[right] = r;
			emitChild(next, left, pivot - first);
			emitChild(next, right, last - pivot);
		}
		grid.sync();
	}
}

template <int T, int E>
__global__ void __launch_bounds__(T) buildLevelKernel(const DevBuild ctx, const uint32_t* __restrict__ cur, const NextLists next) {
	__shared__ typename std::conditional<E == 8, SharedWide<T>, Shared<T>>::type sh;
	const uint32_t tid = threadIdx.x;
	const uint32_t slot = cur[blockIdx.x];
	BuildNode* node = ctx.nodes + slot;
	const uint32_t first = node->first, last = node->last;
	const uint32_t count = last - first;

	// bounds of the node (scene_build.cpp build(): rangeBounds over the x-sorted list)
	Box6 acc = emptyBox();
	for (uint32_t pos = first + tid; pos < last; pos += T) boxMax(acc, loadBox(ctx.tb, ctx.sorted[0][pos]));
	const Box6 bounds = ctaReduceMax<T>(acc, sh.warpBox);
	if (tid == 0) {
		node->bounds[0] = bounds.v[0]; node->bounds[1] = bounds.v[1]; node->bounds[2] = bounds.v[2]; node->bounds[3] = 0.0f;
		node->bounds[4] = bounds.v[3]; node->bounds[5] = bounds.v[4]; node->bounds[6] = bounds.v[5]; node->bounds[7] = 0.0f;
	}
	if (count <= 2) // Bvh2.cpp:272
		return;

	const float parentArea = areaScalar(bounds);
	uint32_t bestDim = kNone, pivot = kNone;
	bool split = false;
	if (parentArea > 0.0f) {
		float bestSah = INFINITY;
		for (uint32_t dim = 0; dim < 3; ++dim) {
			uint32_t p;
			if constexpr (E == 8) p = sweepAxisWide<T>(ctx, ctx.sorted[dim], first, last, bestSah, sh);
			else p = sweepAxis<T>(ctx, ctx.sorted[dim], first, last, bestSah, sh);
			if (p != kNone) {
				pivot = p;
				bestDim = dim;
			}
		}
		// Bvh2.cpp:462-467: cost = 2 + rcp_ss(area) * bestSah against the triangle count
		const float cost = __fadd_rn(2.0f, __fmul_rn(__fmul_rn(1.0f, rcpSS(parentArea)), bestSah));
		split = !(cost > (float)(int)count) && pivot != kNone;
	}
	if (!split) {
		if (count >= 127) { // Bvh2.cpp:468-471,478-480: leaf references hold a 7-bit count
			bestDim = 0;
			pivot = (first + last) >> 1;
		}
		else {
			return;
		}
	}

	// partition: the chosen axis is already split at pivot; re-split the other two stably
	{
		const uint32_t* ref = ctx.sorted[bestDim];
		for (uint32_t pos = first + tid; pos < last; pos += T) ctx.goesLeft[ref[pos]] = pos < pivot ? 1 : 0;
		__syncthreads();
		if constexpr (E == 8) {
			splitListWide<T>(ctx, ctx.sorted[(bestDim + 1) % 3], first, pivot, last, sh);
			splitListWide<T>(ctx, ctx.sorted[(bestDim + 2) % 3], first, pivot, last, sh);
		}
		else {
			splitList<T>(ctx, ctx.sorted[(bestDim + 1) % 3], first, pivot, last, sh);
			splitList<T>(ctx, ctx.sorted[(bestDim + 2) % 3], first, pivot, last, sh);
		}
	}

	if (tid == 0) {
		const uint32_t left = slot + 1;
		const uint32_t right = slot + 2 * (pivot - first);
		node->kind = bestDim + 1;
		node->left = left;
		node->right = right;
		BuildNode l{}, r{};
		l.kind = 0; l.parent = slot; l.first = first; l.last = pivot; l.left = kNone; l.right = kNone;
		r.kind = 0; r.parent = slot; r.first = pivot; r.last = last; r.left = kNone; r.right = kNone;
		ctx.nodes[left] = l;
		ctx.nodes[right] = r;
		emitChild(next, left, pivot - first);
		emitChild(next, right, last - pivot);
	}
}

__device__ __forceinline__ uint32_t sortKeyOf(float mid) {
	// order-preserving float -> uint map (Bvh2.cpp:661-669,743-745)
	const uint32_t u = __float_as_uint(mid);
	return u ^ ((u & 0x80000000u) ? 0xffffffffu : 0x80000000u);
}
// Bvh2.cpp:671-687: the reference's packed key order inside each group of eight (0,1,4,5,2,3,6,7);
// triangles past the last multiple of 32 keep natural order. Self-inverse.
__device__ __forceinline__ uint32_t tiePosition(uint32_t i, uint32_t n) {
	if (i >= (n & ~31u)) return i;
	const uint32_t k = i & 7u;
	const uint32_t swapped = (k >= 2 && k <= 5) ? (k ^ 6u) : k;
	return (i & ~7u) | swapped;
}

// Per-triangle bounds and centre keys (Bvh2.cpp:537-753). Element j of the key/value arrays is the
// triangle the reference would have WRITTEN j-th, so that a stable sort breaks ties its way.
__global__ void boundsKeysKernel(const float4* __restrict__ verts, const uint32_t* __restrict__ indices, uint32_t n, float4* __restrict__ tb,
                                 uint32_t* __restrict__ k0, uint32_t* __restrict__ k1, uint32_t* __restrict__ k2, uint32_t* __restrict__ v0,
                                 uint32_t* __restrict__ v1, uint32_t* __restrict__ v2) {
	const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n)
		return;
	const uint32_t t = tiePosition(j, n);
	const float4 p0 = __ldg(verts + indices[3 * (size_t)t]), p1 = __ldg(verts + indices[3 * (size_t)t + 1]), p2 = __ldg(verts + indices[3 * (size_t)t + 2]);
	const float4 mn = make_float4(fminf(fminf(p0.x, p1.x), p2.x), fminf(fminf(p0.y, p1.y), p2.y), fminf(fminf(p0.z, p1.z), p2.z), fminf(fminf(p0.w, p1.w), p2.w));
	const float4 mx = make_float4(fmaxf(fmaxf(p0.x, p1.x), p2.x), fmaxf(fmaxf(p0.y, p1.y), p2.y), fmaxf(fmaxf(p0.z, p1.z), p2.z), fmaxf(fmaxf(p0.w, p1.w), p2.w));
	const uint32_t sign = 0x80000000u;
	tb[2 * (size_t)t] = make_float4(__uint_as_float(__float_as_uint(mn.x) ^ sign), __uint_as_float(__float_as_uint(mn.y) ^ sign),
	                                __uint_as_float(__float_as_uint(mn.z) ^ sign), __uint_as_float(__float_as_uint(mn.w) ^ sign));
	tb[2 * (size_t)t + 1] = mx;
	k0[j] = sortKeyOf(__fmul_rn(__fadd_rn(mn.x, mx.x), 0.5f));
	k1[j] = sortKeyOf(__fmul_rn(__fadd_rn(mn.y, mx.y), 0.5f));
	k2[j] = sortKeyOf(__fmul_rn(__fadd_rn(mn.z, mx.z), 0.5f));
	v0[j] = t; v1[j] = t; v2[j] = t;
}

// Scratch for one build, stream-ordered (default stream) so that repeated builds recycle the pool
// instead of paying cudaMalloc/cudaFree each time.
struct DeviceBuffers {
	std::vector<void*> ptrs;
	~DeviceBuffers() { for (void* p : ptrs) cudaFreeAsync(p, nullptr); }
	template <typename Tp>
	bool alloc(Tp** out, size_t count) {
		void* p = nullptr;
		if (cudaMallocAsync(&p, count * sizeof(Tp) + 256, nullptr) != cudaSuccess) return false;
		ptrs.push_back(p);
		*out = static_cast<Tp*>(p);
		return true;
	}
};

std::once_flag g_rcpOnce;
bool g_rcpOk = false;
float g_rcpTable[2048];


// Everything the SAH build leaves on the device.
struct SahOnDevice {
	DeviceBuffers buf;
	float4* verts = nullptr;
	uint32_t* indices = nullptr;
	BuildNode* nodes = nullptr;  // 2n slots; unused slots have kind == kNone
	uint32_t* sorted0 = nullptr; // x-sorted triangle list, partitioned by the tree
	uint32_t* hist = nullptr;    // radix-sort histogram scratch
	uint32_t n = 0;
	int levels = 0;              // = depth of the deepest leaf, root = 1
	int smCount = 148;
	double msPrepare = 0, msLevels = 0;
};

const char* const kErrRcp = "device build unavailable: this CPU's RCPSS does not follow the table model";
const char* const kErrMem = "device build failed: out of device memory";
const char* const kErrCuda = "device build failed: CUDA error";

double nowSeconds() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

bool runSahOnDevice(const float* vertices4, uint32_t vertexCount, const uint32_t* indices, uint32_t triangleCount, SahOnDevice& out, const char** error) {
	std::call_once(g_rcpOnce, [] { g_rcpOk = fillRcpTable(g_rcpTable); });
	if (!g_rcpOk) { if (error) *error = kErrRcp; return false; }
	const uint32_t n = triangleCount;
	const double t0 = nowSeconds();
	int device = 0, smCount = 148;
	cudaGetDevice(&device);
	cudaDeviceGetAttribute(&smCount, cudaDevAttrMultiProcessorCount, device);
	if (cudaMemcpyToSymbol(c_rcpTable, g_rcpTable, sizeof(g_rcpTable)) != cudaSuccess) { if (error) *error = kErrCuda; return false; }

	DeviceBuffers& buf = out.buf;
	float4 *dVerts, *dTb;
	uint32_t *dIdx, *keys[3], *vals[3], *keysTmp, *valsTmp, *hist, *scratch, *lists[8], *counts;
	float* leftSah;
	uint8_t* goesLeft;
	BuildNode* dNodes;
	bool ok = buf.alloc(&dVerts, vertexCount) && buf.alloc(&dIdx, (size_t)n * 3) && buf.alloc(&dTb, (size_t)n * 2) && buf.alloc(&keysTmp, n) &&
	          buf.alloc(&valsTmp, n) && buf.alloc(&hist, radixSortHistWords()) && buf.alloc(&scratch, n) && buf.alloc(&leftSah, n) &&
	          buf.alloc(&goesLeft, n) && buf.alloc(&dNodes, (size_t)n * 2) && buf.alloc(&counts, 4);
	for (int d = 0; d < 3 && ok; ++d) ok = buf.alloc(&keys[d], n) && buf.alloc(&vals[d], n);
	for (int l = 0; l < 8 && ok; ++l) ok = buf.alloc(&lists[l], n);
	if (!ok) { if (error) *error = kErrMem; return false; }

	cudaError_t e = cudaMemcpy(dVerts, vertices4, (size_t)vertexCount * 16, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = cudaMemcpy(dIdx, indices, (size_t)n * 12, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = cudaMemsetAsync(dNodes, 0xff, (size_t)n * 2 * sizeof(BuildNode)); // kind == kNone: slot unused
	if (e != cudaSuccess) { if (error) *error = kErrCuda; return false; }

	boundsKeysKernel<<<(n + 255u) / 256u, 256>>>(dVerts, dIdx, n, dTb, keys[0], keys[1], keys[2], vals[0], vals[1], vals[2]);
	DevBuild ctx{};
	ctx.n = n;
	ctx.tb = dTb;
	ctx.scratch = scratch;
	ctx.leftSah = leftSah;
	ctx.goesLeft = goesLeft;
	ctx.nodes = dNodes;
	for (int d = 0; d < 3; ++d) {
		uint32_t* sortedVals = nullptr;
		e = launchRadixSort(keys[d], vals[d], keysTmp, valsTmp, hist, n, 32, smCount, nullptr, nullptr, &sortedVals, nullptr);
		if (e != cudaSuccess) { if (error) *error = kErrCuda; return false; }
		ctx.sorted[d] = sortedVals; // four passes: the result is back in vals[d]
		if (sortedVals != vals[d]) { // defensive: keep the three lists in distinct buffers
			e = cudaMemcpyAsync(vals[d], sortedVals, (size_t)n * 4, cudaMemcpyDeviceToDevice);
			ctx.sorted[d] = vals[d];
		}
	}
	cudaDeviceSynchronize();
	const double t1 = nowSeconds();

	BuildNode root{};
	root.kind = 0; root.parent = kNone; root.first = 0; root.last = n; root.left = kNone; root.right = kNone;
	e = cudaMemcpy(dNodes, &root, sizeof(root), cudaMemcpyHostToDevice);
	// level by level: lists[4*cur + {0,1,2,3}] = mid / small / wide / huge nodes of this level, the other four of the next
	auto classOf = [](uint32_t t) { return t > kHugeNode ? 3 : (t > kWideNode ? 2 : (t > kSmallNode ? 0 : 1)); };
	uint32_t have[4] = {0, 0, 0, 0};
	have[classOf(n)] = 1;
	const uint32_t zero = 0;
	if (e == cudaSuccess) e = cudaMemcpy(lists[classOf(n)], &zero, 4, cudaMemcpyHostToDevice);
	if (e != cudaSuccess) { if (error) *error = kErrCuda; return false; }

	// the cooperative kernel for huge nodes: as many CTAs as are co-resident
	CoopScratch sc{};
	int coopGrid = 0;
	if (n > kHugeNode) {
		int perSm = 0;
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, buildHugeLevelKernel<kWideThreads>, kWideThreads, 0);
		coopGrid = perSm * smCount;
		sc.maxChunks = n / (kWideThreads * 8) + 2;
		const size_t blocks8 = (size_t)n / 8 + 8;
		bool okc = coopGrid > 0 && buf.alloc(&sc.chunkBox, (size_t)sc.maxChunks * 6) && buf.alloc(&sc.chunkCarry, (size_t)sc.maxChunks * 6) &&
		           buf.alloc(&sc.chunkMin, sc.maxChunks) && buf.alloc(&sc.chunkBefore, sc.maxChunks) && buf.alloc(&sc.chunkCount, (size_t)sc.maxChunks * 2) &&
		           buf.alloc(&sc.partBox, (size_t)coopGrid * 6) && buf.alloc(&sc.blkM, blocks8) && buf.alloc(&sc.blkMr, blocks8) &&
		           buf.alloc(&sc.blkBefore, blocks8) && buf.alloc(&sc.blkArg, blocks8) && buf.alloc(&sc.tailSah, 8) && buf.alloc(&sc.scal, 16);
		if (!okc) { if (error) *error = kErrMem; return false; }
	}

	int cur = 0, levels = 0;
	for (int level = 0; level < 4096 && (have[0] || have[1] || have[2] || have[3]); ++level, ++levels) {
		cudaMemsetAsync(counts, 0, 16);
		uint32_t** in = lists + 4 * cur;
		NextLists nx;
		for (int c = 0; c < 4; ++c) nx.list[c] = lists[4 * (1 - cur) + c];
		nx.counts = counts;
		if (have[3]) {
			// Whole GPU per node, one node after the other (~0.35 ms of grid barriers each), or one CTA per node
			// all at once (~15 ns per triangle of the largest node)? Measured on B200 (profiles/r01_build_levels_10M.txt):
			// the cooperative kernel wins while the level has few, large nodes.
			const double coopMs = have[3] * 0.35 + 1.5, wideMs = 1.5e-5 * (double)(n >> (level < 31 ? level : 31));
			if (coopMs < wideMs) {
				const uint32_t* list = in[3];
				uint32_t cnt = have[3];
				void* args[] = {(void*)&ctx, (void*)&list, (void*)&cnt, (void*)&nx, (void*)&sc};
				e = cudaLaunchCooperativeKernel((void*)buildHugeLevelKernel<kWideThreads>, dim3(coopGrid), dim3(kWideThreads), args, 0, nullptr);
				if (e != cudaSuccess) { if (error) *error = kErrCuda; return false; }
			}
			else {
				buildLevelKernel<kWideThreads, 8><<<have[3], kWideThreads>>>(ctx, in[3], nx);
			}
		}
		if (have[2]) buildLevelKernel<kWideThreads, 8><<<have[2], kWideThreads>>>(ctx, in[2], nx);
		if (have[0]) buildLevelKernel<256, 1><<<have[0], 256>>>(ctx, in[0], nx);
		if (have[1]) buildLevelKernel<32, 1><<<have[1], 32>>>(ctx, in[1], nx);
		const uint32_t had[4] = {have[0], have[1], have[2], have[3]};
		const double tl = nowSeconds();
		e = cudaMemcpy(have, counts, 16, cudaMemcpyDeviceToHost);
		if (e != cudaSuccess) { if (error) *error = kErrCuda; return false; }
		if (getenv("RACC_B200_BUILD_VERBOSE") && atoi(getenv("RACC_B200_BUILD_VERBOSE")) > 1)
			fprintf(stderr, "racc device build: level %d: %u huge + %u wide + %u mid + %u small nodes, %.2f ms (wait)\n", level, had[3], had[2], had[0], had[1],
			        (nowSeconds() - tl) * 1e3);
		cur = 1 - cur;
	}
	out.verts = dVerts;
	out.indices = dIdx;
	out.nodes = dNodes;
	out.sorted0 = ctx.sorted[0];
	out.hist = hist;
	out.n = n;
	out.levels = levels;
	out.smCount = smCount;
	out.msPrepare = (t1 - t0) * 1e3;
	out.msLevels = (nowSeconds() - t1) * 1e3;
	return true;
}

// ---------------------------------------------------------------------------------------------
// From the sparse tree to the three device images, on the device (scene_build.cpp buildSceneImages /
// Scene.cpp:223-338): inner nodes ordered by surface area (largest first, ties by tree order), leaves
// packed in the order their parents appear, greedy edge-sharing pair merge per leaf, remap words.

// exclusive prefix sum of n words: 8192 per CTA, then the CTA totals by one CTA (n <= 8192 * 8192)
__global__ void __launch_bounds__(1024) scanBlocksKernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n, uint32_t* __restrict__ blockSums) {
	__shared__ uint32_t warpSums[32];
	const uint32_t base = blockIdx.x * 8192u + threadIdx.x * 8u;
	uint32_t v[8], sum = 0;
#pragma unroll
	for (int j = 0; j < 8; ++j) { v[j] = base + j < n ? in[base + j] : 0u; sum += v[j]; }
	uint32_t total;
	const uint32_t ex = ctaExclusiveScanSum<1024>(sum, &total, warpSums);
	uint32_t run = ex;
#pragma unroll
	for (int j = 0; j < 8; ++j) { if (base + j < n) out[base + j] = run; run += v[j]; }
	if (threadIdx.x == 0) blockSums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) scanSumsKernel(uint32_t* __restrict__ blockSums, uint32_t blocks, uint32_t* __restrict__ grandTotal) {
	__shared__ uint32_t warpSums[32];
	const uint32_t base = threadIdx.x * 8u;
	uint32_t v[8], sum = 0;
#pragma unroll
	for (int j = 0; j < 8; ++j) { v[j] = base + j < blocks ? blockSums[base + j] : 0u; sum += v[j]; }
	uint32_t total;
	const uint32_t ex = ctaExclusiveScanSum<1024>(sum, &total, warpSums);
	uint32_t run = ex;
#pragma unroll
	for (int j = 0; j < 8; ++j) { if (base + j < blocks) blockSums[base + j] = run; run += v[j]; }
	if (threadIdx.x == 0) *grandTotal = total;
}
__global__ void __launch_bounds__(1024) scanAddKernel(uint32_t* __restrict__ out, uint32_t n, const uint32_t* __restrict__ blockSums) {
	const uint32_t base = blockIdx.x * 8192u + threadIdx.x * 8u;
	const uint32_t add = blockSums[blockIdx.x];
#pragma unroll
	for (int j = 0; j < 8; ++j) if (base + j < n) out[base + j] += add;
}
// out[i] = sum of in[0..i); *grandTotal (device) = sum of all. blockSums: >= ceil(n/8192) words.
void exclusiveScan(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* blockSums, uint32_t* grandTotal) {
	const uint32_t blocks = (n + 8191u) / 8192u;
	scanBlocksKernel<<<blocks, 1024>>>(in, out, n, blockSums);
	scanSumsKernel<<<1, 1024>>>(blockSums, blocks, grandTotal);
	scanAddKernel<<<blocks, 1024>>>(out, n, blockSums);
}

__global__ void innerFlagKernel(const BuildNode* __restrict__ nodes, uint32_t slots, uint32_t* __restrict__ flag) {
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s < slots) {
		const uint32_t kind = nodes[s].kind;
		flag[s] = (kind != kNone && kind != 0u) ? 1u : 0u;
	}
}
// key = surface area, larger first (scene_build.cpp boxArea: dx*dy + dy*dz + dx*dz); root first
__global__ void innerKeyKernel(const BuildNode* __restrict__ nodes, uint32_t slots, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ rank,
                               uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= slots || !flag[s])
		return;
	const BuildNode& nd = nodes[s];
	// bbMax - bbMin with bbMin = -bounds[k]
	const float dx = __fsub_rn(nd.bounds[4], -nd.bounds[0]), dy = __fsub_rn(nd.bounds[5], -nd.bounds[1]), dz = __fsub_rn(nd.bounds[6], -nd.bounds[2]);
	const float area = __fadd_rn(__fadd_rn(__fmul_rn(dx, dy), __fmul_rn(dy, dz)), __fmul_rn(dx, dz));
	// areas are >= 0: the bit pattern orders like the value; invert for "largest first"
	keys[rank[s]] = s == 0 ? 0u : ~__float_as_uint(area);
	vals[rank[s]] = s;
}
__global__ void deviceOfKernel(const uint32_t* __restrict__ order, uint32_t count, uint32_t* __restrict__ deviceOf) {
	const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
	if (d < count) deviceOf[order[d]] = d;
}
__global__ void leafChildCountKernel(const BuildNode* __restrict__ nodes, const uint32_t* __restrict__ order, uint32_t count, uint32_t* __restrict__ leafChildren) {
	const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
	if (d >= count)
		return;
	const BuildNode& nd = nodes[order[d]];
	leafChildren[d] = (nodes[nd.left].kind == 0u ? 1u : 0u) + (nodes[nd.right].kind == 0u ? 1u : 0u);
}

// Directed edge (a[e0], a[e0+1]) of triangle a equals the reversed edge of b (Scene.cpp:109-120).
__device__ __forceinline__ bool sharedEdgeDev(const uint32_t* a, const uint32_t* b, unsigned& e0, unsigned& e1) {
	for (unsigned i = 0; i < 3; ++i)
		for (unsigned j = 0; j < 3; ++j)
			if (a[i] == b[(j + 1) % 3] && a[(i + 1) % 3] == b[j]) {
				e0 = i;
				e1 = j;
				return true;
			}
	return false;
}
__device__ __forceinline__ void writePair(float4* pairs, uint32_t at, float4 p0, float4 p1, float4 p2, float4 p3) {
	// e1 = p0 - p1, e2 = p2 - p0, e3 = p3 - p0 (Scene.cpp:149-153); layout Scene.cpp:80-87
	pairs[3 * (size_t)at + 0] = make_float4(__fsub_rn(p0.x, p1.x), __fsub_rn(p0.y, p1.y), __fsub_rn(p0.z, p1.z), __fsub_rn(p3.x, p0.x));
	pairs[3 * (size_t)at + 1] = make_float4(__fsub_rn(p2.x, p0.x), __fsub_rn(p2.y, p0.y), __fsub_rn(p2.z, p0.z), __fsub_rn(p3.y, p0.y));
	pairs[3 * (size_t)at + 2] = make_float4(p0.x, p0.y, p0.z, __fsub_rn(p3.z, p0.z));
}

// Greedy pairing of one leaf's triangles in list order (Scene.cpp:122-181,251-256). One thread per
// leaf. kWrite == false: only counts the pairs; kWrite == true: writes pairs and remap words at the
// leaf's offset. Leaf l is the l-th leaf child in device order (first child before last child).
template <bool kWrite>
__global__ void leafMergeKernel(const BuildNode* __restrict__ nodes, const uint32_t* __restrict__ order, uint32_t innerCount,
                                const uint32_t* __restrict__ leafBase, const uint32_t* __restrict__ sorted0, const uint32_t* __restrict__ indices,
                                const float4* __restrict__ verts, uint32_t* __restrict__ pairCount, const uint32_t* __restrict__ pairStart,
                                float4* __restrict__ pairs, uint32_t* __restrict__ remap) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= 2 * innerCount)
		return;
	const uint32_t d = t >> 1, c = t & 1u;
	const BuildNode& nd = nodes[order[d]];
	const uint32_t childSlot = c ? nd.right : nd.left;
	const BuildNode& leaf = nodes[childSlot];
	if (leaf.kind != 0u)
		return;
	const uint32_t l = leafBase[d] + ((c && nodes[nd.left].kind == 0u) ? 1u : 0u);
	uint32_t cand[128];
	uint32_t m = leaf.last - leaf.first;
	for (uint32_t i = 0; i < m; ++i) cand[i] = sorted0[leaf.first + i];
	uint32_t at = kWrite ? pairStart[l] : 0u, made = 0;
	while (m) {
		const uint32_t a = cand[0];
		--m;
		for (uint32_t i = 0; i < m; ++i) cand[i] = cand[i + 1];
		uint32_t ta[3] = {indices[3 * (size_t)a], indices[3 * (size_t)a + 1], indices[3 * (size_t)a + 2]};
		bool merged = false;
		for (uint32_t i = 0; i < m; ++i) {
			const uint32_t b = cand[i];
			uint32_t tbv[3] = {indices[3 * (size_t)b], indices[3 * (size_t)b + 1], indices[3 * (size_t)b + 2]};
			unsigned e0, e1;
			if (!sharedEdgeDev(ta, tbv, e0, e1)) continue;
			if (kWrite) {
				remap[2 * (size_t)at] = a | (e0 << 30);
				remap[2 * (size_t)at + 1] = b | ((e1 + 1) << 30); // 3 == "no rotation", same as 0 (Kernels.h:232-235)
				writePair(pairs, at, verts[ta[e0]], verts[ta[(e0 + 1) % 3]], verts[ta[(e0 + 2) % 3]], verts[tbv[(e1 + 2) % 3]]);
			}
			++at; ++made;
			--m;
			for (uint32_t k = i; k < m; ++k) cand[k] = cand[k + 1];
			merged = true;
			break;
		}
		if (!merged) {
			// singleton: second triangle degenerates to (p0, p1, p1), never hit (Scene.cpp:160-180)
			if (kWrite) {
				remap[2 * (size_t)at] = a;
				remap[2 * (size_t)at + 1] = 0u;
				const float4 p1 = verts[ta[1]];
				writePair(pairs, at, verts[ta[0]], p1, verts[ta[2]], p1);
			}
			++at; ++made;
		}
	}
	if (!kWrite) pairCount[l] = made;
}

// 64-byte inner node, both child boxes inline (Scene.cpp:73-78,274-332)
__global__ void emitNodesKernel(const BuildNode* __restrict__ nodes, const uint32_t* __restrict__ order, uint32_t innerCount,
                                const uint32_t* __restrict__ deviceOf, const uint32_t* __restrict__ leafBase, const uint32_t* __restrict__ pairCount,
                                const uint32_t* __restrict__ pairStart, float4* __restrict__ out, uint32_t* __restrict__ overflow) {
	const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
	if (d >= innerCount)
		return;
	const BuildNode& nd = nodes[order[d]];
	const BuildNode& l = nodes[nd.left];
	const BuildNode& r = nodes[nd.right];
	uint32_t ref[2];
	uint32_t leafIndex = leafBase[d];
	const BuildNode* child[2] = {&l, &r};
	const uint32_t childSlot[2] = {nd.left, nd.right};
	for (int c = 0; c < 2; ++c) {
		if (child[c]->kind != 0u) {
			ref[c] = 0x80000000u | deviceOf[childSlot[c]];
		}
		else {
			const uint32_t start = pairStart[leafIndex], count = pairCount[leafIndex];
			if (start + count > (1u << 24)) *overflow = 1u;
			ref[c] = (count << 24) | start; // Scene.cpp:298,308
			++leafIndex;
		}
	}
	const uint32_t parent = nd.parent == kNone ? kNone : deviceOf[nd.parent];
	out[4 * (size_t)d + 0] = make_float4(__uint_as_float(nd.kind), __uint_as_float(parent), __uint_as_float(ref[0]), __uint_as_float(ref[1]));
	out[4 * (size_t)d + 1] = make_float4(-l.bounds[0], -l.bounds[1], -l.bounds[2], l.bounds[4]);
	out[4 * (size_t)d + 2] = make_float4(l.bounds[5], l.bounds[6], -r.bounds[0], -r.bounds[1]);
	out[4 * (size_t)d + 3] = make_float4(-r.bounds[2], r.bounds[4], r.bounds[5], r.bounds[6]);
}
// tail padding: copies of pair 0 (Scene.cpp:335-338)
__global__ void padPairsKernel(float4* __restrict__ pairs, uint32_t realPairs, uint32_t pairCount) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < (pairCount - realPairs) * 3u) pairs[3 * (size_t)realPairs + i] = pairs[i % 3u];
}

} // namespace

// DeviceBvhBuilder (scene_build.h): SAH tree on the GPU, handed back to the host for packing.
bool buildBvh2Device(const float* vertices4, uint32_t vertexCount, const uint32_t* indices, uint32_t triangleCount,
                     std::vector<BuildNode>* outNodes, std::vector<uint32_t>* outSorted0, const char** error) {
	SahOnDevice sah;
	if (!runSahOnDevice(vertices4, vertexCount, indices, triangleCount, sah, error))
		return false;
	const uint32_t n = triangleCount;
	const double t2 = nowSeconds();
	outNodes->resize((size_t)n * 2);
	outSorted0->resize(n);
	cudaError_t e = cudaMemcpy(outNodes->data(), sah.nodes, (size_t)n * 2 * sizeof(BuildNode), cudaMemcpyDeviceToHost);
	if (e == cudaSuccess) e = cudaMemcpy(outSorted0->data(), sah.sorted0, (size_t)n * 4, cudaMemcpyDeviceToHost);
	if (e != cudaSuccess) { if (error) *error = kErrCuda; return false; }
	if (getenv("RACC_B200_BUILD_VERBOSE"))
		fprintf(stderr, "racc device build: %u triangles: upload+bounds+3 sorts %.1f ms, %d levels %.1f ms, download %.1f ms\n", n,
		        sah.msPrepare, sah.levels, sah.msLevels, (nowSeconds() - t2) * 1e3);
	return true;
}

// The whole scene build on the device: SAH tree, node order, pair merge, node packing. The three
// images stay in device memory (caller owns them, cudaFree); nothing but counters crosses PCIe.
bool buildSceneImagesDevice(const float* vertices4, uint32_t vertexCount, const uint32_t* indices, uint32_t indexCount,
                            DeviceSceneImages* out, const char** error) {
	static const char* kMod3 = "index count is not a multiple of 3";
	static const char* kEmpty = "scene has no triangles";
	static const char* kIndex = "triangle index out of range";
	static const char* kTooBig = "scene exceeds 2^30 triangles (remap word holds 30 index bits)";
	static const char* kTiny = "the root is a leaf (tiny scene): the host builder makes the synthetic root";
	static const char* kPairs = "scene exceeds 2^24 triangle pairs (leaf reference holds 24 index bits)";
	if (indexCount % 3) { if (error) *error = kMod3; return false; }
	const uint32_t n = indexCount / 3;
	if (!n) { if (error) *error = kEmpty; return false; }
	if (n >= (1u << 30)) { if (error) *error = kTooBig; return false; }
	for (size_t i = 0; i < (size_t)indexCount; ++i)
		if (indices[i] >= vertexCount) { if (error) *error = kIndex; return false; }
	if (n < 3) { if (error) *error = kTiny; return false; }

	SahOnDevice sah;
	if (!runSahOnDevice(vertices4, vertexCount, indices, n, sah, error))
		return false;
	const double t0 = nowSeconds();
	const uint32_t slots = 2 * n;
	DeviceBuffers& buf = sah.buf;
	uint32_t *flag, *rank, *blockSums, *totals, *keys, *vals, *keysTmp, *valsTmp, *deviceOf, *leafChildren, *leafBase, *pairCount, *pairStart, *overflow;
	bool ok = buf.alloc(&flag, slots) && buf.alloc(&rank, slots) && buf.alloc(&blockSums, 8192 + slots / 8192) && buf.alloc(&totals, 8) &&
	          buf.alloc(&keys, n) && buf.alloc(&vals, n) && buf.alloc(&keysTmp, n) && buf.alloc(&valsTmp, n) && buf.alloc(&deviceOf, slots) &&
	          buf.alloc(&leafChildren, n) && buf.alloc(&leafBase, n) && buf.alloc(&pairCount, n + 1) && buf.alloc(&pairStart, n + 1) && buf.alloc(&overflow, 1);
	if (!ok) { if (error) *error = kErrMem; return false; }
	cudaMemsetAsync(overflow, 0, 4);

	// inner nodes in tree (slot) order, then ordered by area with a stable sort
	innerFlagKernel<<<(slots + 255u) / 256u, 256>>>(sah.nodes, slots, flag);
	exclusiveScan(flag, rank, slots, blockSums, totals + 0);
	uint32_t innerCount = 0;
	cudaError_t e = cudaMemcpy(&innerCount, totals + 0, 4, cudaMemcpyDeviceToHost);
	if (e != cudaSuccess) { if (error) *error = kErrCuda; return false; }
	if (!innerCount) { if (error) *error = kTiny; return false; } // root is a leaf
	innerKeyKernel<<<(slots + 255u) / 256u, 256>>>(sah.nodes, slots, flag, rank, keys, vals);
	uint32_t* order = nullptr;
	e = launchRadixSort(keys, vals, keysTmp, valsTmp, sah.hist, innerCount, 32, sah.smCount, nullptr, nullptr, &order, nullptr);
	if (e != cudaSuccess) { if (error) *error = kErrCuda; return false; }
	deviceOfKernel<<<(innerCount + 255u) / 256u, 256>>>(order, innerCount, deviceOf);

	// leaves in the order their parents appear, first child before last child
	leafChildCountKernel<<<(innerCount + 255u) / 256u, 256>>>(sah.nodes, order, innerCount, leafChildren);
	exclusiveScan(leafChildren, leafBase, innerCount, blockSums, totals + 1);
	uint32_t leafCount = 0;
	e = cudaMemcpy(&leafCount, totals + 1, 4, cudaMemcpyDeviceToHost);
	if (e != cudaSuccess) { if (error) *error = kErrCuda; return false; }
	leafMergeKernel<false><<<(2 * innerCount + 127u) / 128u, 128>>>(sah.nodes, order, innerCount, leafBase, sah.sorted0, sah.indices, sah.verts,
	                                                              pairCount, nullptr, nullptr, nullptr);
	exclusiveScan(pairCount, pairStart, leafCount, blockSums, totals + 2);
	uint32_t realPairs = 0;
	e = cudaMemcpy(&realPairs, totals + 2, 4, cudaMemcpyDeviceToHost);
	if (e != cudaSuccess) { if (error) *error = kErrCuda; return false; }
	if (realPairs > (1u << 24)) { if (error) *error = kPairs; return false; }
	const uint32_t pad = 32u - (realPairs % 32u); // at least one pad pair, total a multiple of 32 float4 (Scene.cpp:335-338)
	const uint32_t pairTotal = realPairs + pad;

	float4 *dNodes = nullptr, *dPairs = nullptr, *dVertsOut = nullptr;
	uint32_t *dRemap = nullptr, *dIndicesOut = nullptr;
	if (cudaMalloc(reinterpret_cast<void**>(&dNodes), (size_t)innerCount * 64) != cudaSuccess ||
	    cudaMalloc(reinterpret_cast<void**>(&dPairs), (size_t)pairTotal * 48) != cudaSuccess ||
	    cudaMalloc(reinterpret_cast<void**>(&dRemap), (size_t)realPairs * 8 + 16) != cudaSuccess ||
	    cudaMalloc(reinterpret_cast<void**>(&dVertsOut), (size_t)vertexCount * 16) != cudaSuccess ||
	    cudaMalloc(reinterpret_cast<void**>(&dIndicesOut), (size_t)indexCount * 4) != cudaSuccess) {
		cudaFree(dNodes); cudaFree(dPairs); cudaFree(dRemap); cudaFree(dVertsOut); cudaFree(dIndicesOut);
		if (error) *error = kErrMem;
		return false;
	}
	cudaMemcpyAsync(dVertsOut, sah.verts, (size_t)vertexCount * 16, cudaMemcpyDeviceToDevice);
	cudaMemcpyAsync(dIndicesOut, sah.indices, (size_t)indexCount * 4, cudaMemcpyDeviceToDevice);
	leafMergeKernel<true><<<(2 * innerCount + 127u) / 128u, 128>>>(sah.nodes, order, innerCount, leafBase, sah.sorted0, sah.indices, sah.verts,
	                                                             pairCount, pairStart, dPairs, dRemap);
	emitNodesKernel<<<(innerCount + 255u) / 256u, 256>>>(sah.nodes, order, innerCount, deviceOf, leafBase, pairCount, pairStart, dNodes, overflow);
	padPairsKernel<<<(pad * 3u + 255u) / 256u, 256>>>(dPairs, realPairs, pairTotal);
	uint32_t over = 0;
	BuildNode root{};
	e = cudaMemcpy(&over, overflow, 4, cudaMemcpyDeviceToHost);
	if (e == cudaSuccess) e = cudaMemcpy(&root, sah.nodes, sizeof(root), cudaMemcpyDeviceToHost);
	if (e == cudaSuccess) e = cudaDeviceSynchronize();
	if (e != cudaSuccess || over) {
		cudaFree(dNodes); cudaFree(dPairs); cudaFree(dRemap); cudaFree(dVertsOut); cudaFree(dIndicesOut);
		if (error) *error = over ? kPairs : kErrCuda;
		return false;
	}
	out->verts = dVertsOut;
	out->indices = dIndicesOut;
	out->nodes = dNodes;
	out->pairs = dPairs;
	out->remap = dRemap;
	out->nodeCount = innerCount;
	out->pairCount = pairTotal;
	out->realPairs = realPairs;
	out->remapCount = 2 * realPairs;
	out->depth = (uint32_t)sah.levels;
	for (int k = 0; k < 3; ++k) {
		out->boundsMin[k] = -root.bounds[k];
		out->boundsMax[k] = root.bounds[4 + k];
	}
	if (getenv("RACC_B200_BUILD_VERBOSE"))
		fprintf(stderr, "racc device scene build: %u triangles: upload+bounds+3 sorts %.1f ms, %d levels %.1f ms, order+merge+packing %.1f ms -> %u nodes, %u pairs\n",
		        n, sah.msPrepare, sah.levels, sah.msLevels, (nowSeconds() - t0) * 1e3, innerCount, realPairs);
	return true;
}

} // namespace racc_b200
