// raysort.cu -- ray re-binning before traversal: builds a visiting
// order for the rays of one launch so that the 32 rays a warp holds start close together and point
// the same way. The traversal kernel (traverse_packed.cu) reads rays through this permutation and
// still writes results index-parallel to the rays, so nothing changes at the boundary
// (racc::Result i belongs to racc::Ray i, RayAccelerator.h:59-83) and every ray performs exactly
// the same tests; only which rays share a warp changes.
//
// The reference has no such stage: its iGPU kernel runs 8-wide work-groups over streams in arrival
// order (RayAccelerator.cpp:380-403). The nearest idiom is the per-material radix sort its path
// tracer does before shading (Renderer/PathTracingRenderer.cpp:16-51,124).
//
//   key   = Morton code of the ray origin quantised inside the scene bounds (originBits per axis),
//           combined with a Morton code of the direction on the unit cube (dirBits per axis);
//   order = stable LSD radix sort of (key, ray index), 8 bits per pass, hand-written: every warp owns
//           a contiguous segment; pass = per-warp digit histogram -> row scans -> ranked scatter
//           (ranks inside a 32-element chunk by match.any, running bases in shared memory).
#include "engine.h"

namespace racc_b200 {
namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kSortBlock = 256;              // 8 warps per CTA
constexpr int kWarpsPerBlock = kSortBlock / 32;

__device__ __forceinline__ uint32_t expand3(uint32_t v) { // 10 bits -> every third bit
	v &= 0x3ffu;
	v = (v | (v << 16)) & 0x030000ffu;
	v = (v | (v << 8)) & 0x0300f00fu;
	v = (v | (v << 4)) & 0x030c30c3u;
	v = (v | (v << 2)) & 0x09249249u;
	return v;
}

__device__ __forceinline__ uint32_t quantise(float x, int bits) { // x in [0,1] (anything else is clamped; NaN -> 0)
	const float s = (float)(1u << bits);
	float q = x * s;
	if (!(q > 0.0f)) q = 0.0f;
	if (q > s - 1.0f) q = s - 1.0f;
	return (uint32_t)q;
}

struct KeyArgs {
	float bmin[3];
	float invExtent[3];
	int originBits; // per axis, 0..10
	int dirBits;    // per axis, 0..10
	int dirMajor;   // 0: origin bits above direction bits, 1: direction bits above origin bits
};

__global__ void rayKeyKernel(const TraceParams p, const KeyArgs a, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
	const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= p.total)
		return;
	const DevRay* rays = p.single.rays;
	uint32_t local = idx;
	if (p.nstreams > 1) {
		uint32_t lo = 0, hi = p.nstreams - 1;
		while (lo < hi) {
			const uint32_t mid = (lo + hi + 1) >> 1;
			if (__ldg(&p.streams[mid].begin) <= idx) lo = mid; else hi = mid - 1;
		}
		rays = p.streams[lo].rays;
		local = idx - p.streams[lo].begin;
	}
	const float4 o = __ldg(&rays[local].a);
	const float4 d = __ldg(&rays[local].b);
	uint32_t ko = 0, kd = 0;
	if (a.originBits) {
		const uint32_t qx = quantise((o.x - a.bmin[0]) * a.invExtent[0], a.originBits);
		const uint32_t qy = quantise((o.y - a.bmin[1]) * a.invExtent[1], a.originBits);
		const uint32_t qz = quantise((o.z - a.bmin[2]) * a.invExtent[2], a.originBits);
		ko = expand3(qx) | (expand3(qy) << 1) | (expand3(qz) << 2);
	}
	if (a.dirBits) {
		const float m = fmaxf(fmaxf(fabsf(d.x), fabsf(d.y)), fabsf(d.z));
		const float s = m > 0.0f ? 0.5f / m : 0.0f;
		const uint32_t qx = quantise(fmaf(d.x, s, 0.5f), a.dirBits);
		const uint32_t qy = quantise(fmaf(d.y, s, 0.5f), a.dirBits);
		const uint32_t qz = quantise(fmaf(d.z, s, 0.5f), a.dirBits);
		kd = expand3(qx) | (expand3(qy) << 1) | (expand3(qz) << 2);
	}
	keys[idx] = a.dirMajor ? (kd << (3 * a.originBits)) | ko : (ko << (3 * a.dirBits)) | kd;
	vals[idx] = idx;
}

// ---- one radix pass -----------------------------------------------------------------------------
// Warp w owns elements [w*seg, min((w+1)*seg, total)). hist[digit*warps + w] = its count of `digit`.

__global__ void __launch_bounds__(kSortBlock) radixHistKernel(const uint32_t* __restrict__ keys, uint32_t total, uint32_t seg, uint32_t warps,
                                                             int shift, uint32_t* __restrict__ hist) {
	__shared__ uint32_t counts[kWarpsPerBlock][256];
	const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const uint32_t w = blockIdx.x * kWarpsPerBlock + wib;
	for (int i = lane; i < 256; i += 32) counts[wib][i] = 0;
	__syncwarp();
	if (w < warps) {
		const uint32_t begin = w * seg;
		const uint32_t end = min(begin + seg, total);
		for (uint32_t c = begin; c < end; c += 32) {
			const uint32_t i = c + lane;
			const bool valid = i < end;
			const unsigned mask = __ballot_sync(kFull, valid);
			if (valid) {
				const uint32_t digit = (__ldg(keys + i) >> shift) & 255u;
				const unsigned peers = __match_any_sync(mask, digit);
				if (lane == (unsigned)(__ffs(peers) - 1)) counts[wib][digit] += __popc(peers);
			}
			__syncwarp();
		}
		for (int i = lane; i < 256; i += 32) hist[(size_t)i * warps + w] = counts[wib][i];
	}
}

// Exclusive scan of each digit row in place; rowTotal[digit] = row sum. One CTA per digit.
__global__ void __launch_bounds__(1024) radixRowScanKernel(uint32_t* __restrict__ hist, uint32_t warps, uint32_t* __restrict__ rowTotal) {
	__shared__ uint32_t warpSums[32];
	__shared__ uint32_t carry, chunkTotal;
	uint32_t* row = hist + (size_t)blockIdx.x * warps;
	const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (uint32_t start = 0; start < warps; start += 1024) {
		const uint32_t i = start + threadIdx.x;
		const uint32_t v = i < warps ? row[i] : 0;
		uint32_t incl = v;
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t n = __shfl_up_sync(kFull, incl, o);
			if ((int)lane >= o) incl += n;
		}
		if (lane == 31) warpSums[wib] = incl;
		__syncthreads();
		if (wib == 0) {
			const uint32_t s = warpSums[lane];
			uint32_t sIncl = s;
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t n = __shfl_up_sync(kFull, sIncl, o);
				if ((int)lane >= o) sIncl += n;
			}
			warpSums[lane] = sIncl - s; // exclusive prefix over the 32 warps
			if (lane == 31) chunkTotal = sIncl;
		}
		__syncthreads();
		if (i < warps) row[i] = carry + warpSums[wib] + incl - v;
		__syncthreads();
		if (threadIdx.x == 0) carry += chunkTotal;
		__syncthreads();
	}
	if (threadIdx.x == 0) rowTotal[blockIdx.x] = carry;
}

// Exclusive scan of the 256 row totals (one CTA of 256 threads).
__global__ void __launch_bounds__(256) radixDigitScanKernel(const uint32_t* __restrict__ rowTotal, uint32_t* __restrict__ rowBase) {
	__shared__ uint32_t warpSums[8];
	const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const uint32_t v = rowTotal[threadIdx.x];
	uint32_t incl = v;
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t n = __shfl_up_sync(kFull, incl, o);
		if ((int)lane >= o) incl += n;
	}
	if (lane == 31) warpSums[wib] = incl;
	__syncthreads();
	uint32_t base = 0;
	for (unsigned k = 0; k < wib; ++k) base += warpSums[k];
	rowBase[threadIdx.x] = base + incl - v;
}

__global__ void __launch_bounds__(kSortBlock) radixScatterKernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint32_t total,
                                                                uint32_t seg, uint32_t warps, int shift, const uint32_t* __restrict__ hist,
                                                                const uint32_t* __restrict__ rowBase, uint32_t* __restrict__ keysOut,
                                                                uint32_t* __restrict__ valsOut) {
	__shared__ uint32_t bases[kWarpsPerBlock][256];
	const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const unsigned lt = (1u << lane) - 1u;
	const uint32_t w = blockIdx.x * kWarpsPerBlock + wib;
	if (w >= warps)
		return;
	for (int i = lane; i < 256; i += 32) bases[wib][i] = rowBase[i] + hist[(size_t)i * warps + w];
	__syncwarp();
	const uint32_t begin = w * seg;
	const uint32_t end = min(begin + seg, total);
	for (uint32_t c = begin; c < end; c += 32) {
		const uint32_t i = c + lane;
		const bool valid = i < end;
		const unsigned mask = __ballot_sync(kFull, valid);
		if (valid) {
			const uint32_t key = __ldg(keys + i);
			const uint32_t val = __ldg(vals + i);
			const uint32_t digit = (key >> shift) & 255u;
			const unsigned peers = __match_any_sync(mask, digit);
			const uint32_t pos = bases[wib][digit] + __popc(peers & lt);
			if (keysOut) keysOut[pos] = key;
			valsOut[pos] = val;
			__syncwarp(mask);
			if (lane == (unsigned)(__ffs(peers) - 1)) bases[wib][digit] += __popc(peers);
		}
		__syncwarp();
	}
}

} // namespace

size_t radixSortHistWords() { return (size_t)256 * 148 * 16 + 512; }

// Stable LSD radix sort of (key, value) pairs on the low `keyBits` bits, 8 bits per pass. keys/vals
// and keysTmp/valsTmp ping-pong; on return *keysOut / *valsOut point at the buffers holding the
// sorted sequence (keysOut may be null; the last pass then skips writing keys). hist needs
// radixSortHistWords() words.
cudaError_t launchRadixSort(uint32_t* keys, uint32_t* vals, uint32_t* keysTmp, uint32_t* valsTmp, uint32_t* hist, uint32_t total,
                            int keyBits, int smCount, cudaStream_t stream, uint32_t** keysOut, uint32_t** valsOut, int* launches) {
	uint32_t warps = (uint32_t)smCount * 16u;
	if (warps > 148u * 16u) warps = 148u * 16u;
	uint32_t seg = (total + warps - 1) / warps;
	seg = (seg + 31u) & ~31u;
	if (seg < 512u) seg = 512u;
	warps = (total + seg - 1) / seg;
	uint32_t* rowTotal = hist + (size_t)256 * 148 * 16;
	uint32_t* rowBase = rowTotal + 256;
	const uint32_t blocks = (warps + kWarpsPerBlock - 1) / kWarpsPerBlock;
	const int passes = (keyBits + 7) / 8;
	for (int pass = 0; pass < passes; ++pass) {
		const int shift = 8 * pass;
		const bool last = pass == passes - 1;
		radixHistKernel<<<blocks, kSortBlock, 0, stream>>>(keys, total, seg, warps, shift, hist);
		radixRowScanKernel<<<256, 1024, 0, stream>>>(hist, warps, rowTotal);
		radixDigitScanKernel<<<1, 256, 0, stream>>>(rowTotal, rowBase);
		radixScatterKernel<<<blocks, kSortBlock, 0, stream>>>(keys, vals, total, seg, warps, shift, hist, rowBase, (last && !keysOut) ? nullptr : keysTmp, valsTmp);
		if (launches) *launches += 4;
		uint32_t* t = keys; keys = keysTmp; keysTmp = t;
		t = vals; vals = valsTmp; valsTmp = t;
	}
	if (keysOut) *keysOut = keys;
	*valsOut = vals;
	return cudaGetLastError();
}

size_t raySortScratchBytes(uint32_t total) {
	// keys x2, vals x2, histogram (256 x warps), row totals and bases
	const size_t n = ((size_t)total + 63) & ~(size_t)63;
	return n * 4 * 4 + radixSortHistWords() * 4 + 4096;
}

// Builds the visiting order of the launch described by `p` into scratch memory and returns it in
// *perm (a pointer into scratch, valid until scratch is reused). All work is enqueued on `stream`.
cudaError_t launchRaySort(const TraceParams& p, const float boundsMin[3], const float boundsMax[3], int originBits, int dirBits,
                          int dirMajor, void* scratch, int smCount, cudaStream_t stream, const uint32_t** perm, int* launches) {
	const uint32_t total = p.total;
	if (originBits < 0) originBits = 0;
	if (dirBits < 0) dirBits = 0;
	if (originBits > 10) originBits = 10;
	if (dirBits > 10) dirBits = 10;
	while (3 * (originBits + dirBits) > 32) { if (originBits > dirBits) --originBits; else --dirBits; }
	const int keyBits = 3 * (originBits + dirBits);
	const size_t n = ((size_t)total + 63) & ~(size_t)63;
	uint32_t* keysA = static_cast<uint32_t*>(scratch);
	uint32_t* keysB = keysA + n;
	uint32_t* valsA = keysB + n;
	uint32_t* valsB = valsA + n;
	uint32_t* hist = valsB + n;

	KeyArgs a;
	for (int k = 0; k < 3; ++k) {
		a.bmin[k] = boundsMin[k];
		const float e = boundsMax[k] - boundsMin[k];
		a.invExtent[k] = e > 0.0f ? 1.0f / e : 0.0f;
	}
	a.originBits = originBits;
	a.dirBits = dirBits;
	a.dirMajor = dirMajor;
	rayKeyKernel<<<(total + 255u) / 256u, 256, 0, stream>>>(p, a, keysA, valsA);
	if (launches) *launches += 1;
	uint32_t* sorted = nullptr;
	cudaError_t e = launchRadixSort(keysA, valsA, keysB, valsB, hist, total, keyBits, smCount, stream, nullptr, &sorted, launches);
	*perm = sorted;
	return e;
}

} // namespace racc_b200
