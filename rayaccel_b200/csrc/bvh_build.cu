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
// by level, one kernel launch per level, ONE CTA PER NODE of the level (256 threads for large nodes,
// one warp for nodes of <= 64 triangles). Inside a CTA the sequential sweep becomes
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

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace racc_b200 {
namespace {

constexpr uint32_t kNone = 0xffffffffu;
constexpr uint32_t kSmallNode = 64; // nodes up to this many triangles are built by one warp

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

template <int T>
__global__ void __launch_bounds__(T) buildLevelKernel(const DevBuild ctx, const uint32_t* __restrict__ cur, uint32_t* __restrict__ nextBig,
                                                     uint32_t* __restrict__ nextSmall, uint32_t* __restrict__ counts) {
	__shared__ Shared<T> sh;
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
			const uint32_t p = sweepAxis<T>(ctx, ctx.sorted[dim], first, last, bestSah, sh);
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
		splitList<T>(ctx, ctx.sorted[(bestDim + 1) % 3], first, pivot, last, sh);
		splitList<T>(ctx, ctx.sorted[(bestDim + 2) % 3], first, pivot, last, sh);
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
		const uint32_t child[2] = {left, right};
		const uint32_t size[2] = {pivot - first, last - pivot};
		for (int c = 0; c < 2; ++c) {
			if (size[c] > kSmallNode) nextBig[atomicAdd(&counts[0], 1u)] = child[c];
			else nextSmall[atomicAdd(&counts[1], 1u)] = child[c];
		}
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

struct DeviceBuffers {
	std::vector<void*> ptrs;
	~DeviceBuffers() { for (void* p : ptrs) cudaFree(p); }
	template <typename Tp>
	bool alloc(Tp** out, size_t count) {
		void* p = nullptr;
		if (cudaMalloc(&p, count * sizeof(Tp) + 256) != cudaSuccess) return false;
		ptrs.push_back(p);
		*out = static_cast<Tp*>(p);
		return true;
	}
};

std::once_flag g_rcpOnce;
bool g_rcpOk = false;
float g_rcpTable[2048];

} // namespace

// DeviceBvhBuilder (scene_build.h). Uses the current CUDA device.
bool buildBvh2Device(const float* vertices4, uint32_t vertexCount, const uint32_t* indices, uint32_t triangleCount,
                     std::vector<BuildNode>* outNodes, std::vector<uint32_t>* outSorted0, const char** error) {
	static const char* kRcp = "device build unavailable: this CPU's RCPSS does not follow the table model";
	static const char* kMem = "device build failed: out of device memory";
	static const char* kCuda = "device build failed: CUDA error";
	std::call_once(g_rcpOnce, [] { g_rcpOk = fillRcpTable(g_rcpTable); });
	if (!g_rcpOk) { if (error) *error = kRcp; return false; }
	const uint32_t n = triangleCount;
	const bool verbose = getenv("RACC_B200_BUILD_VERBOSE") != nullptr;
	auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	const double t0 = now();
	int device = 0, smCount = 148;
	cudaGetDevice(&device);
	cudaDeviceGetAttribute(&smCount, cudaDevAttrMultiProcessorCount, device);
	if (cudaMemcpyToSymbol(c_rcpTable, g_rcpTable, sizeof(g_rcpTable)) != cudaSuccess) { if (error) *error = kCuda; return false; }

	DeviceBuffers buf;
	float4 *dVerts, *dTb;
	uint32_t *dIdx, *keys[3], *vals[3], *keysTmp, *valsTmp, *hist, *scratch, *lists[4], *counts;
	float* leftSah;
	uint8_t* goesLeft;
	BuildNode* dNodes;
	bool ok = buf.alloc(&dVerts, vertexCount) && buf.alloc(&dIdx, (size_t)n * 3) && buf.alloc(&dTb, (size_t)n * 2) && buf.alloc(&keysTmp, n) &&
	          buf.alloc(&valsTmp, n) && buf.alloc(&hist, radixSortHistWords()) && buf.alloc(&scratch, n) && buf.alloc(&leftSah, n) &&
	          buf.alloc(&goesLeft, n) && buf.alloc(&dNodes, (size_t)n * 2) && buf.alloc(&counts, 2);
	for (int d = 0; d < 3 && ok; ++d) ok = buf.alloc(&keys[d], n) && buf.alloc(&vals[d], n);
	for (int l = 0; l < 4 && ok; ++l) ok = buf.alloc(&lists[l], n);
	if (!ok) { if (error) *error = kMem; return false; }

	cudaError_t e = cudaMemcpy(dVerts, vertices4, (size_t)vertexCount * 16, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = cudaMemcpy(dIdx, indices, (size_t)n * 12, cudaMemcpyHostToDevice);
	if (e != cudaSuccess) { if (error) *error = kCuda; return false; }

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
		if (e != cudaSuccess) { if (error) *error = kCuda; return false; }
		// four passes: the result is back in vals[d]
		ctx.sorted[d] = sortedVals;
		if (sortedVals != vals[d]) { // defensive: keep the three lists in distinct buffers
			e = cudaMemcpyAsync(vals[d], sortedVals, (size_t)n * 4, cudaMemcpyDeviceToDevice);
			ctx.sorted[d] = vals[d];
		}
	}

	if (verbose) cudaDeviceSynchronize();
	const double t1 = now();
	BuildNode root{};
	root.kind = 0; root.parent = kNone; root.first = 0; root.last = n; root.left = kNone; root.right = kNone;
	e = cudaMemcpy(dNodes, &root, sizeof(root), cudaMemcpyHostToDevice);
	const uint32_t zero = 0;
	if (e == cudaSuccess) e = cudaMemcpy(lists[n > kSmallNode ? 0 : 1], &zero, 4, cudaMemcpyHostToDevice);
	if (e != cudaSuccess) { if (error) *error = kCuda; return false; }

	// level by level: lists[0]/[1] = big/small nodes of this level, lists[2]/[3] = of the next
	uint32_t have[2] = {n > kSmallNode ? 1u : 0u, n > kSmallNode ? 0u : 1u};
	int cur = 0, levels = 0;
	for (int level = 0; level < 4096 && (have[0] || have[1]); ++level, ++levels) {
		cudaMemsetAsync(counts, 0, 8);
		uint32_t* big = lists[2 * cur], * small = lists[2 * cur + 1];
		uint32_t* nextBig = lists[2 * (1 - cur)], * nextSmall = lists[2 * (1 - cur) + 1];
		if (have[0]) buildLevelKernel<256><<<have[0], 256>>>(ctx, big, nextBig, nextSmall, counts);
		if (have[1]) buildLevelKernel<32><<<have[1], 32>>>(ctx, small, nextBig, nextSmall, counts);
		e = cudaMemcpy(have, counts, 8, cudaMemcpyDeviceToHost);
		if (e != cudaSuccess) { if (error) *error = kCuda; return false; }
		cur = 1 - cur;
	}

	const double t2 = now();
	outNodes->resize((size_t)n * 2);
	outSorted0->resize(n);
	e = cudaMemcpy(outNodes->data(), dNodes, (size_t)n * 2 * sizeof(BuildNode), cudaMemcpyDeviceToHost);
	if (e == cudaSuccess) e = cudaMemcpy(outSorted0->data(), ctx.sorted[0], (size_t)n * 4, cudaMemcpyDeviceToHost);
	if (e != cudaSuccess) { if (error) *error = kCuda; return false; }
	if (verbose)
		fprintf(stderr, "racc device build: %u triangles: upload+bounds+3 sorts %.1f ms, %d levels %.1f ms, download %.1f ms\n", n,
		        (t1 - t0) * 1e3, levels, (t2 - t1) * 1e3, (now() - t2) * 1e3);
	return true;
}

} // namespace racc_b200
