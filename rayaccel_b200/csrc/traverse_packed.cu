// traverse_packed.cu -- the default traversal kernel: persistent while-while with bail-out over a
// device-private "packed" copy of the scene images, written for sm_100a.
//
// WHAT it computes is unchanged (Kernels.h:36-242 in the pinned arithmetic of DESIGN.md section 3);
// every ray sees exactly the same sequence of box and pair tests as in the reference-format kernels
// of traverse.cu, so results stay bit-identical to the oracle. HOW differs in three places:
//
//   * node layout. The reference's 64-byte node (Scene.cpp:73-78) stores {lMin.xyz,lMax.x |
//     lMax.yz,rMin.xy | rMin.z,rMax.xyz}. The packed node stores each axis' (min,max) side by side,
//     {lx lX ly lY lz lZ rx rX | ry rY rz rZ first last - -}, so that one Blackwell packed-fp32
//     instruction (fma.rn.ftz.f32x2 -> FFMA2 with the ray's invDir/OoD broadcast) evaluates the near
//     and far plane of an axis at once: 6 FFMA2 per node instead of 12 FFMA. Same IEEE fma per
//     component, so the same bits.
//   * pair layout. 64 bytes instead of 48: {e1.xyz,e3.x | e2.xyz,e3.y | p0.xyz,e3.z | n1.xyz,-},
//     where n1 = e1 x e2 is computed once per scene by packPairsKernel with the very instruction
//     sequence the per-ray code used (mad_cross, Kernels.h:23-25). A pair is two aligned 256-bit
//     loads that never straddle a cache line (the 48-byte stride does every third pair).
//   * control. Child references keep the reference's encoding (bit 31 = inner), which makes the
//     state tests single signed compares (<0 inner, >0 leaf, 0 none) and the node address one
//     shift-add from a pre-biased base held in uniform registers; the push is a predicated STL.
//
// An optional permutation (TraceParams::perm, built by raysort.cu) makes the kernel visit the rays
// in a coherence-improving order; results are still written index-parallel to the rays.
//
// kQuant (tuning variant 4; SURVEY.md section 8f rank 4, "compressed node format"): the same kernel walking 32-byte
// QUANTISED inner nodes -- the same BVH2, child references and visiting rule, but both child boxes as 12 16-bit
// coordinates on a global grid over the scene bounds, rounded outwards with a quarter cell of margin. A node visit is then ONE
// 256-bit gather instead of two (the L1 data pipe, not HBM, bounds this kernel: DESIGN.md section 5.1) and the node
// image is half as large. The boxes are conservative, so a ray visits a superset of the leaves the exact boxes would
// let it into and finds the same closest hit with the same t, u, v (the pair test is unchanged); what can differ from
// the reference's result is only which of two triangles wins an EXACT tie in t (the visiting order of near-equal
// children may flip) and rays that graze a box the fp32 slab test of the reference misses by rounding. That is
// north_star's bar (ids equal except fp ties, |dt|/t <= 1e-4), not the bit-exact bar of the default format, so the
// exact format stays the default and this one is opt-in.
#include "traverse_packed.cuh"

namespace racc_b200 {
namespace {

// ---------------------------------------------------------------------------------------------
// reference-format images -> packed images (once per scene, on the device so that n1 is produced by
// the same arithmetic the per-ray code would use)

__global__ void packNodesKernel(const float4* __restrict__ nodes, uint32_t count, float4* __restrict__ out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count)
		return;
	const float4 d0 = nodes[4 * (size_t)i], d1 = nodes[4 * (size_t)i + 1], d2 = nodes[4 * (size_t)i + 2], d3 = nodes[4 * (size_t)i + 3];
	// left box: min (d1.x,d1.y,d1.z) max (d1.w,d2.x,d2.y); right box: min (d2.z,d2.w,d3.x) max (d3.y,d3.z,d3.w)
	out[4 * (size_t)i + 0] = make_float4(d1.x, d1.w, d1.y, d2.x);
	out[4 * (size_t)i + 1] = make_float4(d1.z, d2.y, d2.z, d3.y);
	out[4 * (size_t)i + 2] = make_float4(d2.w, d3.z, d3.x, d3.w);
	out[4 * (size_t)i + 3] = make_float4(d0.z, d0.w, 0.0f, 0.0f);
}

__global__ void packPairsKernel(const float4* __restrict__ pairs, uint32_t count, float4* __restrict__ out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count)
		return;
	const float4 t0 = pairs[3 * (size_t)i], t1 = pairs[3 * (size_t)i + 1], t2 = pairs[3 * (size_t)i + 2];
	RACC_CROSS(n1x, n1y, n1z, t0.x, t0.y, t0.z, t1.x, t1.y, t1.z)
	out[4 * (size_t)i + 0] = t0;
	out[4 * (size_t)i + 1] = t1;
	out[4 * (size_t)i + 2] = t2;
	out[4 * (size_t)i + 3] = make_float4(n1x, n1y, n1z, 0.0f);
}

// Packed node image -> 32-byte quantised nodes: {Lx Ly Lz Rx Ry Rz | first last}, every box word = (min | max << 16) in
// cells of a 16-bit grid over the scene bounds (plus a few cells of padding, launchQuantiseNodes). min rounds down and max rounds up after moving a quarter cell outwards:
// the quantised box contains the fp32 box with a margin of at least 0.25 cell, ~15x what the kernel's grid-space slab
// arithmetic can differ from the reference's (both err by about 2^-22 of the scene extent = 0.016 cell). A whole extra
// cell of padding was measured and dropped: rays that START on flat, axis-aligned geometry (battlefield's ground: its
// exact boxes have zero thickness, so a bounce ray leaving the ground never enters them) would then begin inside every
// such box on their way -- +23 % node visits and +82 % pair tests on the first bounce (profiles/r02_quantised_nodes.md).
// Arithmetic in double: done once per scene.
struct QuantGrid { float origin[3]; float cell[3]; double inverse[3]; double margin[3]; };

__global__ void quantiseNodesKernel(const float4* __restrict__ nodes, uint32_t count, QuantGrid g, uint4* __restrict__ out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count)
		return;
	const float4 d0 = nodes[4 * (size_t)i], d1 = nodes[4 * (size_t)i + 1], d2 = nodes[4 * (size_t)i + 2], d3 = nodes[4 * (size_t)i + 3];
	// reference layout (Scene.cpp:73-78): left min (d1.x,d1.y,d1.z) max (d1.w,d2.x,d2.y); right min (d2.z,d2.w,d3.x) max (d3.y,d3.z,d3.w)
	const float mn[6] = {d1.x, d1.y, d1.z, d2.z, d2.w, d3.x};
	const float mx[6] = {d1.w, d2.x, d2.y, d3.y, d3.z, d3.w};
	uint32_t w[6];
	for (int k = 0; k < 6; ++k) {
		const int a = k % 3;
		double lo = floor(((double)mn[k] - (double)g.origin[a]) * g.inverse[a] - g.margin[a]);
		double hi = ceil(((double)mx[k] - (double)g.origin[a]) * g.inverse[a] + g.margin[a]);
		if (!(lo > 0.0)) lo = 0.0;          // also NaN
		if (lo > 65535.0) lo = 65535.0;     // +inf: the synthetic root's unreachable child (scene_build.cpp)
		if (!(hi > 0.0)) hi = 0.0;
		if (hi > 65535.0) hi = 65535.0;
		w[k] = (uint32_t)lo | ((uint32_t)hi << 16);
	}
	out[2 * (size_t)i + 0] = make_uint4(w[0], w[1], w[2], w[3]);
	out[2 * (size_t)i + 1] = make_uint4(w[4], w[5], __float_as_uint(d0.z), __float_as_uint(d0.w));
}

// Light probe as texel pairs: entry k of row j holds {texel(clamp(k-1)), texel(clamp(k))}, k in [0, width], so the two
// horizontally adjacent texels of a bilinear footprint -- clamp-to-edge included -- are ONE aligned 256-bit load.
__global__ void packEnvKernel(const float4* __restrict__ texels, uint32_t width, uint32_t height, float4* __restrict__ out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (width + 1) * height)
		return;
	const uint32_t j = i / (width + 1), k = i % (width + 1);
	const uint32_t x0 = k ? k - 1 : 0, x1 = k < width ? k : width - 1;
	out[2 * (size_t)i + 0] = texels[(size_t)j * width + x0];
	out[2 * (size_t)i + 1] = texels[(size_t)j * width + x1];
}

template <bool kCount, int kBlock, int kMinBlocks, int kSmStack, bool kQuant>
__global__ void __launch_bounds__(kBlock, kMinBlocks) tracePackedKernel(const TraceParams p, const int fetchThreshold, const int innerBail, const int leafBail) {
	const unsigned lane = threadIdx.x & 31;
	const unsigned ltMask = (1u << lane) - 1u;
	// both bases are made opaque so that they stay in registers (otherwise they are re-derived from
	// the constant bank, several instructions, at every step)
	u64 nodeBase, pairBase;
	asm volatile("mov.b64 %0, %1;" : "=l"(nodeBase) : "l"(kQuant ? reinterpret_cast<u64>(p.qnodes) - (0x80000000ull << 5)
	                                                                    : reinterpret_cast<u64>(p.tnodes) - (0x80000000ull << 6)));
	asm volatile("mov.b64 %0, %1;" : "=l"(pairBase) : "l"(reinterpret_cast<u64>(p.tpairs)));

	const uint32_t total = p.totalPtr ? min(__ldg(p.totalPtr), p.total) : p.total;
	bool exhausted = false; // warp-uniform: the cursor has run past the last ray
	RayState r; HitState h;
	__shared__ uint32_t smStack[kSmStack ? kSmStack : 1][kSmStack ? kBlock : 1];
	uint32_t stackStorage[kStackSize - kSmStack];
	typename std::conditional<(kSmStack > 0), HybridStack<(kSmStack > 0 ? kSmStack : 1), kBlock>, PlainStack>::type stack;
	if constexpr (kSmStack > 0) stack.attach(smStack, stackStorage);
	else stack.attach(stackStorage);
	uint32_t node = 0;         // 0: no ray in flight on this lane; bit 31: at an inner node; else at a leaf
	float4* outPtr = nullptr;  // non-null: the lane holds a ray (in flight, or finished and not yet written)
	unsigned long long cInner = 0, cPairs = 0;
	unsigned cRays = 0, cHits = 0, cPushes = 0, cLeaves = 0; // the last two only in the kCount build

	for (;;) {
		// ---- retire finished lanes and refill idle ones, warp-wide -----------------------------
		unsigned idle = __ballot_sync(kFullMask, node == 0);
		if (idle == kFullMask || __popc(idle) >= (exhausted ? 32 : fetchThreshold)) {
			if (node == 0 && outPtr) {
				*outPtr = finishRayPacked(p, r, h);
				++cRays;
				cHits += h.index != kMiss;
				outPtr = nullptr;
			}
			if (!exhausted) {
				const int want = __popc(idle);
				const int leader = __ffs(idle) - 1;
				uint32_t base = 0;
				if ((int)lane == leader) base = atomicAdd(p.cursor, (uint32_t)want);
				base = __shfl_sync(kFullMask, base, leader);
				if (node == 0) {
					uint32_t idx = base + __popc(idle & ltMask);
					if (idx < total) {
						if (p.perm) idx = __ldg(p.perm + idx);
						const DevRay* rays; uint32_t local;
						locate(p, idx, rays, outPtr, local);
						outPtr += local;
						initRay(rays, local, r, h);
						if (kQuant) quantRay(p, r);
						stack.reset();
						node = kInnerBit;
					}
				}
				exhausted = base + (uint32_t)want >= total;
			}
			idle = __ballot_sync(kFullMask, node == 0);
			if (idle == kFullMask)
				break;
		}
		const unsigned liveMask = ~idle;

		// ---- inner phase: warp-uniform loop, left once too few lanes still descend -------------
		{
			unsigned innerMask = __ballot_sync(kFullMask, (int)node < 0);
			int descending = __popc(innerMask);
			bool go = innerMask == liveMask || descending >= innerBail || 2 * descending > __popc(liveMask);
			go = go && innerMask;
			while (go) {
				if ((int)node < 0) {
					if (kCount) ++cInner;
					node = kQuant ? innerStepQuant<kCount>(nodeBase, node, r, stack, cPushes) : innerStepPacked<kCount>(nodeBase, node, r, stack, cPushes);
					if (kCount) cLeaves += (int)node > 0;
				}
				innerMask = __ballot_sync(kFullMask, (int)node < 0);
				go = innerMask == liveMask || __popc(innerMask) >= innerBail;
			}
		}
		// ---- leaf phase: one pair per iteration, left once too few lanes still have pairs while
		// others wait at inner nodes -----------------------------------------------------------
		{
			bool go = __ballot_sync(kFullMask, (int)node > 0) != 0;
			while (go) {
				if ((int)node > 0) {
					pairTestPacked(pairBase, node & 0xffffffu, r, h);
					if (kCount) ++cPairs;
					// one pair of this leaf done: (count << 24 | first) -> (count-1 << 24 | first+1)
					const bool lastPair = node < 0x2000000u;
					node = !lastPair ? node + 1u - 0x1000000u : (stack.empty() ? 0u : stack.pop());
					if (kCount) cLeaves += lastPair && (int)node > 0; // the popped entry is another leaf
				}
				const unsigned leafMask = __ballot_sync(kFullMask, (int)node > 0);
				go = leafMask != 0;
				if (go && __popc(leafMask) < leafBail)
					go = __ballot_sync(kFullMask, (int)node < 0) == 0;
			}
		}
	}

	// Frame statistics (rays, hits): one atomic per warp, always on when a counter record is given;
	// this is the value the multi-GPU hit reduction sums. Visit counters only in the kCount build.
	if (p.counters) {
		unsigned long long rays = cRays, hits = cHits, pushes = cPushes, leaves = cLeaves;
		for (int o = 16; o; o >>= 1) {
			rays += __shfl_xor_sync(kFullMask, rays, o);
			hits += __shfl_xor_sync(kFullMask, hits, o);
			if (kCount) {
				cInner += __shfl_xor_sync(kFullMask, cInner, o);
				cPairs += __shfl_xor_sync(kFullMask, cPairs, o);
				pushes += __shfl_xor_sync(kFullMask, pushes, o);
				leaves += __shfl_xor_sync(kFullMask, leaves, o);
			}
		}
		if (lane == 0) {
			atomicAdd(p.counters + 0, rays);
			atomicAdd(p.counters + 1, hits);
			if (kCount) {
				atomicAdd(p.counters + 2, cInner);
				atomicAdd(p.counters + 3, cPairs);
				atomicAdd(p.counters + 4, pushes);
				atomicAdd(p.counters + 5, leaves);
			}
		}
	}
}

template <bool kCount, int kBlock, int kMinBlocks, int kSmStack = 0, bool kQuant = false>
cudaError_t launchPacked(const TraceParams& p, const Tuning& t, int smCount, cudaStream_t stream) {
	auto kernel = tracePackedKernel<kCount, kBlock, kMinBlocks, kSmStack, kQuant>;
	static thread_local int plannedDevice = -1, plannedCarveout = -2, resident = 1;
	int device = 0;
	cudaGetDevice(&device);
	cudaError_t err;
	if (plannedDevice != device || plannedCarveout != t.carveout) {
		// shared memory only for the (optional) stack tops: the rest of the 228 KB of each SM serves as L1
		int carve = 0;
		if (kSmStack > 0) {
			int smPerSm = 233472;
			cudaDeviceGetAttribute(&smPerSm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
			carve = (int)(((size_t)(kSmStack * kBlock * 4 + 1024) * kMinBlocks * 100 + smPerSm - 1) / smPerSm);
			if (carve > 100) carve = 100;
		}
		err = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, t.carveout >= 0 ? t.carveout : carve);
		if (err != cudaSuccess) return err;
		err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kernel, kBlock, 0);
		if (err != cudaSuccess) return err;
		if (resident < 1) resident = 1;
		plannedDevice = device;
		plannedCarveout = t.carveout;
	}
	int ctas = resident;
	if (t.ctasPerSm > 0 && ctas > t.ctasPerSm) ctas = t.ctasPerSm;
	if (t.gridCtasPerSm > 0 && ctas > t.gridCtasPerSm) ctas = t.gridCtasPerSm; // same instantiation, fewer resident CTAs
	long long grid = (long long)smCount * ctas;
	const long long needed = ((long long)p.total + kBlock - 1) / kBlock;
	if (grid > needed) grid = needed > 0 ? needed : 1;
	err = cudaMemsetAsync(p.cursor, 0, sizeof(uint32_t), stream);
	if (err != cudaSuccess) return err;
	// innerBail 0 and 1 mean the same (stay while any lane descends); the loop test relies on >= 1
	kernel<<<(unsigned)grid, kBlock, 0, stream>>>(p, t.fetchThreshold, t.innerBail < 1 ? 1 : t.innerBail, t.leafBail);
	return cudaGetLastError();
}

template <bool kCount>
cudaError_t dispatchPacked(const TraceParams& p, const Tuning& t, int smCount, cudaStream_t stream) {
	if (t.variant == 4) { // quantised nodes: the default launch shape, stack all-local or its tops in shared memory
		if (t.smemStack > 0) return launchPacked<kCount, 256, 5, 16, true>(p, t, smCount, stream);
		return launchPacked<kCount, 256, 5, 0, true>(p, t, smCount, stream);
	}
	if (t.smemStack > 0 && t.blockThreads == 256) { // stack tops in shared memory: 256 x 5 only
		if (t.smemStack <= 8) return launchPacked<kCount, 256, 5, 8>(p, t, smCount, stream);
		return launchPacked<kCount, 256, 5, 16>(p, t, smCount, stream);
	}
	switch (t.blockThreads * 100 + t.ctasPerSm) {
	case 12800 + 8: return launchPacked<kCount, 128, 8>(p, t, smCount, stream);
	case 12800 + 10: return launchPacked<kCount, 128, 10>(p, t, smCount, stream);
	case 12800 + 12: return launchPacked<kCount, 128, 12>(p, t, smCount, stream);
	case 25600 + 4: return launchPacked<kCount, 256, 4>(p, t, smCount, stream);
	case 25600 + 6: return launchPacked<kCount, 256, 6>(p, t, smCount, stream);
	case 51200 + 2: return launchPacked<kCount, 512, 2>(p, t, smCount, stream);
	case 51200 + 3: return launchPacked<kCount, 512, 3>(p, t, smCount, stream);
	default: return launchPacked<kCount, 256, 5>(p, t, smCount, stream);
	}
}

} // namespace

cudaError_t launchPackImages(const float4* nodes, uint32_t nodeCount, const float4* pairs, uint32_t pairCount,
                             float4* tnodes, float4* tpairs, cudaStream_t stream, int* launches) {
	if (nodeCount) {
		packNodesKernel<<<(nodeCount + 255u) / 256u, 256, 0, stream>>>(nodes, nodeCount, tnodes);
		if (launches) *launches += 1;
	}
	if (pairCount) {
		packPairsKernel<<<(pairCount + 255u) / 256u, 256, 0, stream>>>(pairs, pairCount, tpairs);
		if (launches) *launches += 1;
	}
	return cudaGetLastError();
}

cudaError_t launchQuantiseNodes(const float4* nodes, uint32_t nodeCount, const float boundsMin[3], const float boundsMax[3], void* qnodes,
                                float qOrigin[3], float qCell[3], cudaStream_t stream, int* launches) {
	QuantGrid g;
	for (int a = 0; a < 3; ++a) {
		// The margin (quantiseNodesKernel) is a quarter cell, or more where fp32 cannot resolve a quarter cell: a scene far from
		// the origin has coordinates whose ulp exceeds the cell, and the kernel's grid-space slab arithmetic errs by a few of
		// those ulps. The grid starts `pad` cells below the scene's lower bound so that the margin exists on that side too
		// (with cell 0 at the bound itself a ray through a vertex ON the bound could slip past the root box: found by
		// tests/fuzz/fuzz_gpu.py on a four-triangle scene).
		const double extent = (double)boundsMax[a] - (double)boundsMin[a];
		const bool usable = extent > 0.0 && extent < 3.0e38;
		const double reach = fmax(fabs((double)boundsMin[a]), fabs((double)boundsMax[a]));
		const double ulp = reach * 1.1920928955078125e-7;
		double margin = 0.25, pad = 1.0; // the common case: bounds at cells 1 and 65534, planes rounded out to 0 and 65535
		if (usable) {
			margin = fmax(0.25, 4.0 * ulp / (extent / 65533.0));
			if (margin > 8192.0) margin = 8192.0;
			pad = ceil(margin + 0.75);
		}
		const double cells = 65535.0 - 2.0 * pad;
		const double cell = usable ? extent / cells : 0.0;
		g.margin[a] = margin;
		g.cell[a] = (float)cell;
		g.origin[a] = (float)((double)boundsMin[a] - pad * cell);
		g.inverse[a] = usable ? cells / extent : 0.0;
		qOrigin[a] = g.origin[a];
		qCell[a] = g.cell[a];
	}
	if (nodeCount) {
		quantiseNodesKernel<<<(nodeCount + 255u) / 256u, 256, 0, stream>>>(nodes, nodeCount, g, static_cast<uint4*>(qnodes));
		if (launches) *launches += 1;
	}
	return cudaGetLastError();
}

cudaError_t launchPackEnv(const float4* texels, uint32_t width, uint32_t height, float4* pairsOut, cudaStream_t stream, int* launches) {
	const uint32_t n = (width + 1) * height;
	packEnvKernel<<<(n + 255u) / 256u, 256, 0, stream>>>(texels, width, height, pairsOut);
	if (launches) *launches += 1;
	return cudaGetLastError();
}

cudaError_t launchTracePacked(const TraceParams& p, const Tuning& t, int counterMode, int smCount, cudaStream_t stream, int* launches) {
	if (!p.total)
		return cudaSuccess;
	if (launches) *launches += 1;
	return counterMode == 2 ? dispatchPacked<true>(p, t, smCount, stream) : dispatchPacked<false>(p, t, smCount, stream);
}

} // namespace racc_b200
