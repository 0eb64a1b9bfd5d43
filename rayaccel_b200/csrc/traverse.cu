// traverse.cu -- the hot path: closest-hit BVH2 traversal + triangle-pair intersection over packed
// ray streams, hand-written for sm_100a.
//
// WHAT it computes is fixed by the reference's OpenCL `traversal` kernel
// (/root/reference/RayAccelerator/Kernels.h:36-242): same slab test, same triangle-pair test with
// its accept/tie rules, same near-first order and push rule, same miss (light-probe) and hit
// (remap + barycentric rotation) epilogues, in the pinned fp32 arithmetic of DESIGN.md section 3
// (mad -> fmaf, everything else separately rounded, IEEE reciprocal, FTZ). Compiled with
// -fmad=false -ftz=true -prec-div=true -prec-sqrt=true so the compiler adds or removes no rounding.
//
// HOW is B200-first, not the reference's one-work-item-per-ray NDRange:
//   * persistent CTAs (grid = SMs x resident CTAs) pulling rays from a global cursor, one warp-
//     aggregated atomic per refill, refilling idle lanes when enough of a warp has retired;
//   * while-while traversal: every lane descends inner nodes until it holds a leaf, then the warp
//     reconverges and tests triangle pairs, so the two instruction streams are not interleaved;
//   * the hottest inner nodes (the scene builder orders nodes by surface area) are staged once per
//     CTA into shared memory with a TMA bulk copy (cp.async.bulk + mbarrier);
//   * rays are read as 2 x LDG.128, nodes as 4 x 128-bit loads, pairs as 3 x LDG.128 through the
//     read-only path; results leave as one STG.128 per ray, index-parallel to the rays;
//   * no tensor cores: this is branchy scalar fp32.
#include "traverse_common.cuh"

namespace racc_b200 {

// Diagnostics of the counted (kCount) instantiations only: warp-level loop trip counts, for
// lane-utilisation analysis (racc_cuda_debug_warp_stats). [0] outer rounds, [1] inner-loop
// iterations, [2] leaf-loop iterations, [3] lanes active summed over inner iterations, [4] lanes
// active summed over leaf iterations, [5] refills.
__device__ unsigned long long g_warpStats[8];

namespace {


// The four float4s of one inner node, all requested up front (the child references travel with the
// boxes instead of being fetched after the hit test). kGlobal: read-only global path; else generic
// (the node may live in shared memory).
template <bool kGlobal>
__device__ __forceinline__ void loadNode(const float4* np, float4& d0, float4& d1, float4& d2, float4& d3) {
	if (kGlobal) {
		// Blackwell 256-bit loads (LDG.E.ENL2.256): a 64-byte node is two load instructions instead of
		// four, which halves the L1 wavefronts of a divergent node fetch (one wavefront per distinct
		// line per instruction, whatever the width).
		asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		             : "=f"(d0.x), "=f"(d0.y), "=f"(d0.z), "=f"(d0.w), "=f"(d1.x), "=f"(d1.y), "=f"(d1.z), "=f"(d1.w) : "l"(np));
		asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8+32];"
		             : "=f"(d2.x), "=f"(d2.y), "=f"(d2.z), "=f"(d2.w), "=f"(d3.x), "=f"(d3.y), "=f"(d3.z), "=f"(d3.w) : "l"(np));
	}
	else {
		asm volatile("ld.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(d0.x), "=f"(d0.y), "=f"(d0.z), "=f"(d0.w) : "l"(np));
		asm volatile("ld.v4.f32 {%0,%1,%2,%3}, [%4+16];" : "=f"(d1.x), "=f"(d1.y), "=f"(d1.z), "=f"(d1.w) : "l"(np));
		asm volatile("ld.v4.f32 {%0,%1,%2,%3}, [%4+32];" : "=f"(d2.x), "=f"(d2.y), "=f"(d2.z), "=f"(d2.w) : "l"(np));
		asm volatile("ld.v4.f32 {%0,%1,%2,%3}, [%4+48];" : "=f"(d3.x), "=f"(d3.y), "=f"(d3.z), "=f"(d3.w) : "l"(np));
	}
}

// One inner-node step (Kernels.h:170-199). Returns the next node reference, or 0 when neither child
// is hit and the stack is empty.
template <bool kGlobal>
__device__ __forceinline__ uint32_t innerStep(const float4* np, const RayState& r, LocalStack& stack) {
	float4 d0, d1, d2, d3;
	loadNode<kGlobal>(np, d0, d1, d2, d3);
	const float tRay = r.tFar;
	const float tFirst = slab(d1.x, d1.y, d1.z, d1.w, d2.x, d2.y, r);
	const float tLast = slab(d2.z, d2.w, d3.x, d3.y, d3.z, d3.w, r);
	const float firstDiff = tRay - tFirst;
	const float lastDiff = tRay - tLast;
	if (firstDiff + lastDiff != 0.0f) {
		const bool sgn = (__float_as_uint(tLast - tFirst) >> 31) != 0;
		const uint32_t cf = __float_as_uint(d0.z), cl = __float_as_uint(d0.w);
		if (fmaxf(tFirst, tLast) != tRay)
			stack.push(sgn ? cf : cl);
		return sgn ? cl : cf;
	}
	return stack.empty() ? 0u : stack.pop();
}

// ---------------------------------------------------------------------------------------------
// variant 1: one thread per ray, the reference's launch shape (kept as the simple A/B baseline)

template <bool kCount>
__global__ void __launch_bounds__(256) traceSimpleKernel(const TraceParams p) {
	const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= (p.totalPtr ? min(__ldg(p.totalPtr), p.total) : p.total))
		return;
	const DevRay* rays; float4* out; uint32_t local;
	locate(p, idx, rays, out, local);
	RayState r; HitState h;
	initRay(rays, local, r, h);
	uint32_t stackStorage[kStackSize];
	LocalStack stack;
	stack.attach(stackStorage);
	uint32_t node = kInnerBit;
	unsigned nInner = 0, nPairs = 0;
	for (;;) {
		if (node & kInnerBit) {
			if (kCount) ++nInner;
			node = innerStep<true>(p.nodes + 4 * (size_t)(node & ~kInnerBit), r, stack);
			if (!node) break;
			continue;
		}
		const uint32_t first = node & 0xffffffu, last = first + (node >> 24);
		for (uint32_t i = first; i < last; ++i) {
			pairTest(p.pairs, i, r, h);
			if (kCount) ++nPairs;
		}
		if (stack.empty()) break;
		node = stack.pop();
	}
	out[local] = finishRay(p, r, h);
	if (p.counters) {
		// warp-aggregated: one atomic per counter per warp
		const unsigned active = __activemask();
		const unsigned hits = __popc(__ballot_sync(active, h.index != kMiss));
		unsigned long long sumInner = nInner, sumPairs = nPairs;
		if (kCount) {
			for (int o = 16; o; o >>= 1) {
				sumInner += __shfl_xor_sync(active, sumInner, o);
				sumPairs += __shfl_xor_sync(active, sumPairs, o);
			}
		}
		if ((threadIdx.x & 31) == (unsigned)(__ffs(active) - 1)) {
			atomicAdd(p.counters + 0, (unsigned long long)__popc(active));
			atomicAdd(p.counters + 1, (unsigned long long)hits);
			if (kCount) {
				atomicAdd(p.counters + 2, sumInner);
				atomicAdd(p.counters + 3, sumPairs);
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------
// variant 0: persistent while-while kernel with TMA-staged tree top

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// TMA bulk copy of the first `bytes` of the node array into shared memory, completion on an
// mbarrier (cp.async.bulk -> UBLKCP in SASS). One thread issues, all threads wait.
__device__ __forceinline__ void stageNodes(float4* dst, const float4* src, uint32_t bytes, uint64_t* bar) {
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smemAddr(bar)));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
		const uint32_t kChunk = 32768;
		for (uint32_t off = 0; off < bytes; off += kChunk) {
			const uint32_t n = min(kChunk, bytes - off);
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
			             ::"r"(smemAddr(reinterpret_cast<char*>(dst) + off)), "l"(reinterpret_cast<const char*>(src) + off), "r"(n), "r"(smemAddr(bar))
			             : "memory");
		}
	}
	uint32_t done = 0;
	while (!done) {
		asm volatile(
		    "{\n\t.reg .pred p;\n\t"
		    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
		    "selp.u32 %0, 1, 0, p;\n\t}"
		    : "=r"(done)
		    : "r"(smemAddr(bar))
		    : "memory");
	}
}

// The four float4s of inner node n. The hottest nodes live in shared memory (staged by TMA), the
// rest in global memory behind L1/L2. Both bases are generic 64-bit addresses held in registers so
// that picking one is two selects and the address one IMAD.WIDE.
struct NodeBases {
	const char* shared;
	const char* global;
	uint32_t sharedCount;
};

template <bool kStage>
__device__ __forceinline__ const float4* nodeAddress(const NodeBases& nb, uint32_t n) {
	const char* base = (kStage && n < nb.sharedCount) ? nb.shared : nb.global;
	unsigned long long a;
	asm("mad.wide.u32 %0, %1, 64, %2;" : "=l"(a) : "r"(n), "l"(reinterpret_cast<unsigned long long>(base)));
	return reinterpret_cast<const float4*>(a);
}

// kMode 0: while-while (every lane walks inner nodes until it holds a leaf; then all leaves).
// kMode 1: while-while with bail-out (the inner loop stops once fewer than `innerBail` lanes still
//          descend while others wait; the leaf loop tests one pair per iteration and stops once
//          fewer than `leafBail` lanes still have pairs while others wait at inner nodes).
//
// kStage: the first p.smemNodes inner nodes (the hottest: the builder orders nodes by surface area)
// are served from shared memory, staged once per CTA by a TMA bulk copy; the un-staged
// instantiation reads every node through L1/L2 and saves the per-step address select.
template <bool kCount, int kBlock, int kMinBlocks, int kMode, bool kStage>
__global__ void __launch_bounds__(kBlock, kMinBlocks) tracePersistentKernel(const TraceParams p, const int fetchThreshold, const int innerBail, const int leafBail) {
	extern __shared__ __align__(128) unsigned char smemRaw[];
	__shared__ uint64_t stageBar;
	float4* sNodes = reinterpret_cast<float4*>(smemRaw);
	if (kStage && p.smemNodes)
		stageNodes(sNodes, p.nodes, p.smemNodes * 64u, &stageBar);

	NodeBases nb;
	{
		// materialise the generic address of the staging area once (otherwise the conversion is
		// redone, S2R and all, in the traversal loop)
		unsigned long long g;
		asm volatile("cvta.shared.u64 %0, %1;" : "=l"(g) : "l"((unsigned long long)smemAddr(sNodes)));
		nb.shared = reinterpret_cast<const char*>(g);
		nb.global = reinterpret_cast<const char*>(p.nodes);
		nb.sharedCount = kStage ? p.smemNodes : 0u;
	}

	const unsigned lane = threadIdx.x & 31;
	const unsigned ltMask = (1u << lane) - 1u;

	enum { kEmpty = 0, kTraversing = 1, kFinished = 2 };
	int state = kEmpty;
	const uint32_t total = p.totalPtr ? min(__ldg(p.totalPtr), p.total) : p.total;
	bool exhausted = false; // warp-uniform: the cursor has run past the last ray

	RayState r; HitState h;
	uint32_t stackStorage[kStackSize];
	LocalStack stack;
	stack.attach(stackStorage);
	uint32_t node = 0;
	float4* outPtr = nullptr;
	unsigned long long cInner = 0, cPairs = 0;
	unsigned cRays = 0, cHits = 0;

	for (;;) {
		// ---- retire finished lanes and refill idle ones, warp-wide -----------------------------
		const unsigned idle = __ballot_sync(kFullMask, state != kTraversing);
		if (idle == kFullMask || __popc(idle) >= (exhausted ? 32 : fetchThreshold)) {
			if (state == kFinished) {
				*outPtr = finishRay(p, r, h);
				++cRays;
				cHits += h.index != kMiss;
				state = kEmpty;
			}
			if (!exhausted) {
				const int want = __popc(idle);
				const int leader = __ffs(idle) - 1;
				uint32_t base = 0;
				if ((int)lane == leader) base = atomicAdd(p.cursor, (uint32_t)want);
				base = __shfl_sync(kFullMask, base, leader);
				if (state == kEmpty) {
					const uint32_t idx = base + __popc(idle & ltMask);
					if (idx < total) {
						const DevRay* rays; uint32_t local;
						locate(p, idx, rays, outPtr, local);
						outPtr += local;
						initRay(rays, local, r, h);
						stack.reset();
						node = kInnerBit;
						state = kTraversing;
					}
				}
				exhausted = base + (uint32_t)want >= total;
			}
			if (!__ballot_sync(kFullMask, state == kTraversing))
				break;
		}

		if (kMode == 0) {
			// ---- while-while ---------------------------------------------------------------------
			if (state == kTraversing) {
				while (node & kInnerBit) {
					if (kCount) ++cInner;
					node = innerStep<!kStage>(nodeAddress<kStage>(nb, node & ~kInnerBit), r, stack);
				}
			}
			__syncwarp();
			if (state == kTraversing) {
				if (node) {
					const uint32_t first = node & 0xffffffu, last = first + (node >> 24);
					for (uint32_t i = first; i < last; ++i) {
						pairTest(p.pairs, i, r, h);
						if (kCount) ++cPairs;
					}
					node = stack.empty() ? 0u : stack.pop();
				}
				if (!node)
					state = kFinished;
			}
			__syncwarp();
		}
		else {
			// ---- while-while with bail-out ----------------------------------------------------------
			// Both phases run as warp-uniform loops that stop early once too few lanes still take part:
			// in plain while-while a phase lasts as long as its slowest lane (7-8 inner steps while the
			// median lane needs 2), which on incoherent rays leaves ~60 % of the issue slots masked
			// off. A lane that is cut short keeps its place (node / remaining pairs) and simply
			// continues in the next round, so every ray still sees exactly the same sequence of node
			// and pair tests -- only the interleaving between lanes changes.
			const bool live = state == kTraversing;
			const unsigned liveMask = __ballot_sync(kFullMask, live);
			if (kCount && lane == 0) atomicAdd(&g_warpStats[0], 1ull);
			for (int steps = 0;; ++steps) {
				const bool atInner = live && (node & kInnerBit);
				const unsigned innerMask = __ballot_sync(kFullMask, atInner);
				if (!innerMask)
					break;
				const int descending = __popc(innerMask);
				if (descending < innerBail && innerMask != liveMask && (steps || 2 * descending <= __popc(liveMask)))
					break;
				if (kCount && lane == 0) { atomicAdd(&g_warpStats[1], 1ull); atomicAdd(&g_warpStats[3], (unsigned long long)descending); }
				if (atInner) {
					if (kCount) ++cInner;
					node = innerStep<!kStage>(nodeAddress<kStage>(nb, node & ~kInnerBit), r, stack);
				}
			}
			for (int steps = 0;; ++steps) {
				const bool atLeaf = live && node && !(node & kInnerBit);
				const unsigned leafMask = __ballot_sync(kFullMask, atLeaf);
				if (!leafMask)
					break;
				if (steps && __popc(leafMask) < leafBail && __ballot_sync(kFullMask, live && (node & kInnerBit)))
					break;
				if (kCount && lane == 0) { atomicAdd(&g_warpStats[2], 1ull); atomicAdd(&g_warpStats[4], (unsigned long long)__popc(leafMask)); }
				if (atLeaf) {
					pairTest(p.pairs, node & 0xffffffu, r, h);
					if (kCount) ++cPairs;
					// one pair of this leaf done: (count << 24 | first) -> (count-1 << 24 | first+1)
					node = (node >> 24) > 1u ? node + 1u - 0x1000000u : (stack.empty() ? 0u : stack.pop());
				}
			}
			if (live && !node)
				state = kFinished;
		}
	}

	// Frame statistics (rays, hits): one atomic per warp, always on when a counter record is given;
	// this is the value the multi-GPU hit reduction sums. Visit counters only in the kCount build.
	if (p.counters) {
		unsigned long long rays = cRays, hits = cHits;
		for (int o = 16; o; o >>= 1) {
			rays += __shfl_xor_sync(kFullMask, rays, o);
			hits += __shfl_xor_sync(kFullMask, hits, o);
			if (kCount) {
				cInner += __shfl_xor_sync(kFullMask, cInner, o);
				cPairs += __shfl_xor_sync(kFullMask, cPairs, o);
			}
		}
		if (lane == 0) {
			atomicAdd(p.counters + 0, rays);
			atomicAdd(p.counters + 1, hits);
			if (kCount) {
				atomicAdd(p.counters + 2, cInner);
				atomicAdd(p.counters + 3, cPairs);
			}
		}
	}
}

// Launch shape of one kernel instantiation for one (device, tuning) pair; computed once.
struct LaunchPlan {
	bool valid = false;
	int device = -1, ctasPerSm = -2, smemNodesWanted = -2, carveout = -2;
	uint32_t nodeCount = 0;
	uint32_t smemNodes = 0;
	int resident = 1;
};

template <bool kCount, int kBlock, int kMinBlocks, int kMode, bool kStage>
cudaError_t launchPersistent(const TraceParams& p, const Tuning& t, int smCount, cudaStream_t stream) {
	auto kernel = tracePersistentKernel<kCount, kBlock, kMinBlocks, kMode, kStage>;
	static thread_local LaunchPlan plan;
	cudaError_t err;
	int device = 0;
	cudaGetDevice(&device);
	if (!plan.valid || plan.device != device || plan.ctasPerSm != t.ctasPerSm || plan.smemNodesWanted != t.smemNodes || plan.nodeCount != p.nodeCount || plan.carveout != t.carveout) {
		// Shared-memory budget: stage as many hot nodes as fit while keeping the requested residency.
		int maxOptin = 0, smPerSm = 0;
		cudaDeviceGetAttribute(&maxOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
		cudaDeviceGetAttribute(&smPerSm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
		const int ctas = t.ctasPerSm > 0 ? t.ctasPerSm : kMinBlocks;
		uint32_t wantNodes = !kStage ? 0u : (t.smemNodes > 0 ? (uint32_t)t.smemNodes : 0xffffffffu);
		if (wantNodes > p.nodeCount) wantNodes = p.nodeCount;
		size_t budget = (size_t)smPerSm / (size_t)ctas;
		budget = budget > 2048 ? budget - 2048 : 0; // static smem + the per-CTA system reservation
		if (budget > (size_t)maxOptin - 1024) budget = (size_t)maxOptin - 1024;
		if ((size_t)wantNodes * 64 > budget) wantNodes = (uint32_t)(budget / 64);
		const size_t smemBytes = (size_t)wantNodes * 64;
		err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes);
		if (err != cudaSuccess) return err;
		// Give shared memory exactly what the resident CTAs need; the rest of the 228 KB stays L1,
		// which caches the un-staged nodes, the triangle pairs and the per-thread traversal stacks.
		{
			const size_t perSm = (smemBytes + 1024 + 256) * (size_t)ctas;
			int carveout = (int)((perSm * 100 + (size_t)smPerSm - 1) / (size_t)smPerSm);
			if (t.carveout >= 0) carveout = t.carveout;
			if (carveout > 100) carveout = 100;
			err = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
			if (err != cudaSuccess) return err;
		}
		int resident = 0;
		err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kernel, kBlock, smemBytes);
		if (err != cudaSuccess) return err;
		if (resident < 1) resident = 1;
		if (t.ctasPerSm > 0 && resident > t.ctasPerSm) resident = t.ctasPerSm;
		plan.valid = true;
		plan.device = device;
		plan.ctasPerSm = t.ctasPerSm;
		plan.smemNodesWanted = t.smemNodes;
		plan.nodeCount = p.nodeCount;
		plan.smemNodes = wantNodes;
		plan.resident = resident;
		plan.carveout = t.carveout;
	}
	TraceParams q = p;
	q.smemNodes = plan.smemNodes;

	// Persistent grid: SMs x resident CTAs, but no more CTAs than there is work for.
	long long grid = (long long)smCount * plan.resident;
	const long long needed = ((long long)p.total + kBlock - 1) / kBlock;
	if (grid > needed) grid = needed > 0 ? needed : 1;

	err = cudaMemsetAsync(p.cursor, 0, sizeof(uint32_t), stream);
	if (err != cudaSuccess) return err;
	kernel<<<(unsigned)grid, kBlock, (size_t)plan.smemNodes * 64, stream>>>(q, t.fetchThreshold, t.innerBail, t.leafBail);
	return cudaGetLastError();
}

template <bool kCount>
cudaError_t dispatch(const TraceParams& p, const Tuning& t, int smCount, cudaStream_t stream) {
	if (t.variant == 1) {
		const unsigned grid = (p.total + 255u) / 256u;
		traceSimpleKernel<kCount><<<grid, 256, 0, stream>>>(p);
		return cudaGetLastError();
	}
	const int mode = t.variant == 2 ? 1 : 0;
	const bool stage = t.smemNodes != 0; // -1 = as many as fit, 0 = none (un-staged instantiation)
#define RACC_LAUNCH(B, M)                                                                                 \
	(mode ? (stage ? launchPersistent<kCount, B, M, 1, true>(p, t, smCount, stream) : launchPersistent<kCount, B, M, 1, false>(p, t, smCount, stream)) \
	      : (stage ? launchPersistent<kCount, B, M, 0, true>(p, t, smCount, stream) : launchPersistent<kCount, B, M, 0, false>(p, t, smCount, stream)))
	// the reference-format kernels are kept for A/B only: three launch shapes
	switch (t.blockThreads * 100 + t.ctasPerSm) {
	case 12800 + 8: case 12800: return RACC_LAUNCH(128, 8);
	case 25600 + 4: return RACC_LAUNCH(256, 4);
	default:
		if (t.blockThreads == 128) return RACC_LAUNCH(128, 8);
		return RACC_LAUNCH(256, 5);
	}
#undef RACC_LAUNCH
}

} // namespace

cudaError_t readWarpStats(unsigned long long* out8, bool reset) {
	cudaError_t e = cudaMemcpyFromSymbol(out8, g_warpStats, 8 * sizeof(unsigned long long));
	if (e == cudaSuccess && reset) {
		const unsigned long long zeros[8] = {};
		e = cudaMemcpyToSymbol(g_warpStats, zeros, sizeof(zeros));
	}
	return e;
}

cudaError_t launchTrace(const TraceParams& p, const Tuning& t, int counterMode, int smCount, cudaStream_t stream, int* launches) {
	if (!p.total)
		return cudaSuccess;
	if (launches) *launches += 1;
	return counterMode == 2 ? dispatch<true>(p, t, smCount, stream) : dispatch<false>(p, t, smCount, stream);
}

} // namespace racc_b200
