// pathstream.cu -- the device-side path tracer as ONE persistent kernel per batch lane:
// the traversal loop of traverse_packed.cu with the example renderer's shading (pathshade.cuh) at its retire point and
// a ray queue in HBM that the same kernel fills and drains.
//
// Why: the wavefront form (pathtrace.cu + one traversal launch per bounce) spends a fifth of a frame outside traversal --
// a shading pass that re-reads every ray and result from HBM and ends every bounce with a launch whose tail leaves most
// SMs idle, the worse the smaller the wave (deep bounces, small frames, one GPU's share of a strong-scaled frame). Here
// a lane that finishes a ray shades it on the spot, from registers, and the path's next ray goes to the queue; lanes take
// primary rays (generated in the kernel from the path index, Camera.cpp:55-114) until there are none left, then queue
// entries in arrival order. Nothing waits for a wave to end: the kernel runs until every path of the launch has ended.
//
// Same image: a ray's Result does not depend on which lane traces it or when (same node steps and pair tests as
// tracePackedKernel, from the same functions), a path's shading depends only on (pixel, sample, depth, seed), and every
// path owns its radiance slot, summed per pixel in sample order by pathAccumulateKernel -- so the framebuffer has the
// bits of the wavefront form's and of oracle_path_trace's. Primaries stay together (consecutive path indices per warp
// refill), which keeps the coherence the first wave has in the wavefront form.
//
// The queue. Entry = {origin.xyz,minT | dir.xyz,maxT | weight.rgb, path + (depth << 26)}, 48 bytes, at most
// paths * maxDepth of them (every path adds at most one per bounce), never reused within a launch. A producer reserves
// slots with one atomic per warp on `tail`, writes its entry, fences, then stores the launch's epoch number to the
// slot's flag word; a consumer takes TICKETS -- as many as it sees unclaimed entries, one atomic per warp on `head` -- and
// polls its slot's flag once per pass through the refill block while the other lanes of its warp keep traversing. Warps
// racing for the same entries can push `head` past `tail`: such a ticket simply waits for a ray that a still-running
// path will produce, or for `ended` to reach the number of paths (a path counts as ended only after its last ray was
// shaded, so nothing is outstanding then). A warp with no ray and no ticket leaves as soon as the primaries are gone and
// the queue has nothing unclaimed -- a ray queued later belongs to a running warp, which looks at the queue again
// itself -- so CTAs retire while the last paths finish and the next launch moves in. Nobody waits for a CTA that has not
// started: a waiting lane's producer holds a ray, so it is running. Flags carry epochs instead of being cleared per
// launch (capi_render.cu keeps the array between calls).
#include "traverse_packed.cuh"
#include "pathshade.cuh"

namespace racc_b200 {
namespace {

constexpr uint32_t kNoTicket = 0xffffffffu;
constexpr uint32_t kPathBits = 26; // path index within the batch; the bounce number sits above
constexpr int kDepthSlots = 64;

struct StreamArgs {
	const float4* tnodes;
	const float4* tpairs;
	const uint32_t* remap;
	const float4* envPairs;  // null: no light probe (escaping paths add nothing)
	uint32_t envWidth, envHeight;
	ShadeScene scene;
	CameraArgs cam;
	uint32_t width, pixels, sampleBase, seed, maxDepth;
	uint32_t firstPath, paths;
	float4* queue;
	uint32_t* flags;
	uint32_t capacity, epoch;
	uint32_t* ctrl;          // [0] next primary, [1] queue head (tickets), [2] queue tail, [3] paths ended
	float4* radiance;
	unsigned long long* depthRays;
	unsigned long long* counters;
};

// Per-lane path state that is written when a ray is taken and read when it is shaded: thread-private memory, so that it
// costs the traversal loop no registers. Every lane uses the same offsets: a warp-wide access is one L1 wavefront.
struct Parked {
	uint32_t base;
	__device__ __forceinline__ void attach(uint32_t* storage) { base = (uint32_t)__cvta_generic_to_local(storage); }
	__device__ __forceinline__ void put(int k, uint32_t v) { asm volatile("st.local.u32 [%0], %1;" ::"l"((u64)(base + 4u * k)), "r"(v) : "memory"); }
	__device__ __forceinline__ uint32_t get(int k) {
		uint32_t v;
		asm volatile("ld.local.u32 %0, [%1];" : "=r"(v) : "l"((u64)(base + 4u * k)) : "memory");
		return v;
	}
	__device__ __forceinline__ void putf(int k, float v) { put(k, __float_as_uint(v)); }
	__device__ __forceinline__ float getf(int k) { return __uint_as_float(get(k)); }
};

template <int kBlock, int kMinBlocks, int kSmStack>
__global__ void __launch_bounds__(kBlock, kMinBlocks) pathStreamKernel(const StreamArgs a, const int fetchThreshold, const int innerBail, const int leafBail) {
	const unsigned lane = threadIdx.x & 31;
	const unsigned ltMask = (1u << lane) - 1u;
	u64 nodeBase, pairBase;
	asm volatile("mov.b64 %0, %1;" : "=l"(nodeBase) : "l"(reinterpret_cast<u64>(a.tnodes) - (0x80000000ull << 6)));
	asm volatile("mov.b64 %0, %1;" : "=l"(pairBase) : "l"(reinterpret_cast<u64>(a.tpairs)));

	__shared__ uint32_t depthCount[kDepthSlots];
	if (threadIdx.x < kDepthSlots) depthCount[threadIdx.x] = 0;
	__syncthreads();

	bool exhausted = false; // warp-uniform: no primary rays left
	RayState r; HitState h;
	__shared__ uint32_t smStack[kSmStack ? kSmStack : 1][kSmStack ? kBlock : 1];
	uint32_t stackStorage[kStackSize - kSmStack];
	typename std::conditional<(kSmStack > 0), HybridStack<(kSmStack > 0 ? kSmStack : 1), kBlock>, PlainStack>::type stack;
	if constexpr (kSmStack > 0) stack.attach(smStack, stackStorage);
	else stack.attach(stackStorage);
	uint32_t parkStorage[8];
	Parked park;
	park.attach(parkStorage);
	uint32_t node = 0;           // 0: nothing to traverse on this lane; bit 31: at an inner node; else at a leaf
	bool holding = false;        // the lane has a ray: in flight, or finished and not yet shaded
	uint32_t ticket = kNoTicket; // queue slot this lane waits for
	unsigned cRays = 0, cHits = 0;

	for (;;) {
		unsigned idle = __ballot_sync(kFullMask, node == 0);
		if (__popc(idle) >= fetchThreshold) {
			// ---- shade the finished rays ---------------------------------------------------------------------------
			bool ended = false, push = false;
			DevRay next;
			float4 nextState;
			if (node == 0 && holding) {
				holding = false;
				++cRays;
				const uint32_t pd = park.get(6);
				const uint32_t path = pd & ((1u << kPathBits) - 1u), depth = pd >> kPathBits;
				float weight[3] = {park.getf(3), park.getf(4), park.getf(5)};
				atomicAdd(&depthCount[depth], 1u);
				if (h.index == kMiss) {
					// PathTracingRenderer.cpp:468-566: the light probe's radiance times the path weight
					if (a.envPairs) {
						const float4 m = missRadiancePairs(a.envPairs, a.envWidth, a.envHeight, r);
						a.radiance[path] = make_float4(m.y * weight[0], m.z * weight[1], m.w * weight[2], 0.0f);
					}
					ended = true;
				}
				else {
					++cHits;
					const float4 res = hitResult(a.remap, h);
					const uint32_t tri = __float_as_uint(res.x);
					if (tri < a.scene.triangleCount && depth < a.maxDepth) {
						const float ro[3] = {r.ox, r.oy, r.oz};
						const float rd[3] = {park.getf(0), park.getf(1), park.getf(2)}; // as submitted: r.d* went through the epsilon clamp
						push = shadeHit(a.scene, tri, res.y, res.z, res.w, ro, rd, path % a.pixels, a.sampleBase + path / a.pixels, a.seed, depth, weight, next);
						nextState = make_float4(weight[0], weight[1], weight[2], __uint_as_float(path | ((depth + 1u) << kPathBits)));
					}
					ended = !push;
				}
			}
			// ---- the paths that go on: one slot reservation per warp, entry, fence, flag -------------------------------
			const unsigned pushMask = __ballot_sync(kFullMask, push);
			if (pushMask) {
				const int leader = __ffs(pushMask) - 1;
				uint32_t base = 0;
				if ((int)lane == leader) base = atomicAdd(a.ctrl + 2, (uint32_t)__popc(pushMask));
				base = __shfl_sync(kFullMask, base, leader);
				if (push) {
					const uint32_t slot = base + __popc(pushMask & ltMask);
					float4* e = a.queue + 3 * (size_t)slot;
					e[0] = next.a; e[1] = next.b; e[2] = nextState;
					__threadfence();
					*(volatile uint32_t*)(a.flags + slot) = a.epoch;
				}
			}
			const unsigned endMask = __ballot_sync(kFullMask, ended);
			if (endMask && (int)lane == __ffs(endMask) - 1) atomicAdd(a.ctrl + 3, (uint32_t)__popc(endMask));

			// ---- new work: primary rays while there are any, then queue tickets ----------------------------------------
			unsigned need = __ballot_sync(kFullMask, node == 0 && ticket == kNoTicket);
			if (need && !exhausted) {
				const int want = __popc(need);
				const int leader = __ffs(need) - 1;
				uint32_t base = 0;
				if ((int)lane == leader) base = atomicAdd(a.ctrl + 0, (uint32_t)want);
				base = __shfl_sync(kFullMask, base, leader);
				if (node == 0 && ticket == kNoTicket) {
					const uint32_t idx = base + __popc(need & ltMask);
					if (idx < a.paths) {
						const uint32_t path = a.firstPath + idx;
						const DevRay ray = primaryRay(a.cam, a.width, path % a.pixels, a.sampleBase + path / a.pixels, a.seed);
						initRayFrom(ray.a, ray.b, r, h);
						park.putf(0, ray.b.x); park.putf(1, ray.b.y); park.putf(2, ray.b.z);
						park.putf(3, 1.0f); park.putf(4, 1.0f); park.putf(5, 1.0f);
						park.put(6, path);
						stack.reset();
						node = kInnerBit;
						holding = true;
					}
				}
				exhausted = base + (uint32_t)want >= a.paths;
				need = __ballot_sync(kFullMask, node == 0 && ticket == kNoTicket);
			}
			bool queueEmpty = false; // warp-uniform: a ticket was wanted and the queue had nothing unclaimed
			if (need && exhausted) {
				// as many tickets as there are unclaimed entries right now (several warps may see the same ones: then some
				// tickets run ahead of the tail and wait)
				const int leader = __ffs(need) - 1;
				uint32_t base = 0;
				int n = 0;
				if ((int)lane == leader) {
					const uint32_t tail = *(volatile uint32_t*)(a.ctrl + 2), head = *(volatile uint32_t*)(a.ctrl + 1);
					n = min((int)(tail - head), __popc(need));
					if (n > 0) base = atomicAdd(a.ctrl + 1, (uint32_t)n);
				}
				n = __shfl_sync(kFullMask, n, leader);
				base = __shfl_sync(kFullMask, base, leader);
				if (node == 0 && ticket == kNoTicket && (int)__popc(need & ltMask) < n) ticket = base + __popc(need & ltMask);
				queueEmpty = n <= 0;
			}
			if (node == 0 && ticket < a.capacity) { // kNoTicket and tickets past the queue's end never pass
				if (*(volatile uint32_t*)(a.flags + ticket) == a.epoch) {
					__threadfence();
					const float4* e = a.queue + 3 * (size_t)ticket;
					const float4 ra = __ldcg(e), rb = __ldcg(e + 1), st = __ldcg(e + 2);
					initRayFrom(ra, rb, r, h);
					park.putf(0, rb.x); park.putf(1, rb.y); park.putf(2, rb.z);
					park.putf(3, st.x); park.putf(4, st.y); park.putf(5, st.z);
					park.put(6, __float_as_uint(st.w));
					stack.reset();
					node = kInnerBit;
					holding = true;
					ticket = kNoTicket;
				}
			}
			idle = __ballot_sync(kFullMask, node == 0);
			if (idle == kFullMask) {
				// Nothing to traverse on this warp. Without a ticket it leaves once there are no primaries and no unclaimed
				// entries: whoever queues a ray later is a running warp, which will look at the queue again itself. With a
				// ticket it waits for its ray, or for the last path to end (then its ticket ran ahead of the last entry).
				if (__ballot_sync(kFullMask, ticket != kNoTicket) == 0) {
					if (exhausted && queueEmpty) break;
				}
				else {
					const bool done = *(volatile uint32_t*)(a.ctrl + 3) >= a.paths;
					if (__any_sync(kFullMask, done)) break;
					__nanosleep(200);
				}
				continue;
			}
		}
		const unsigned liveMask = ~idle;

		// ---- inner phase, leaf phase: as in tracePackedKernel (traverse_packed.cu) ------------------------------------
		{
			unsigned innerMask = __ballot_sync(kFullMask, (int)node < 0);
			int descending = __popc(innerMask);
			bool go = innerMask == liveMask || descending >= innerBail || 2 * descending > __popc(liveMask);
			go = go && innerMask;
			unsigned unusedPushes = 0;
			while (go) {
				if ((int)node < 0) node = innerStepPacked<false>(nodeBase, node, r, stack, unusedPushes);
				innerMask = __ballot_sync(kFullMask, (int)node < 0);
				go = innerMask == liveMask || __popc(innerMask) >= innerBail;
			}
		}
		{
			bool go = __ballot_sync(kFullMask, (int)node > 0) != 0;
			while (go) {
				if ((int)node > 0) {
					pairTestPacked(pairBase, node & 0xffffffu, r, h);
					const bool lastPair = node < 0x2000000u;
					node = !lastPair ? node + 1u - 0x1000000u : (stack.empty() ? 0u : stack.pop());
				}
				const unsigned leafMask = __ballot_sync(kFullMask, (int)node > 0);
				go = leafMask != 0;
				if (go && __popc(leafMask) < leafBail)
					go = __ballot_sync(kFullMask, (int)node < 0) == 0;
			}
		}
	}

	// rays traced per bounce (Stats::raysTraced of the equivalent render() calls), frame statistics (rays, hits)
	__syncthreads();
	if (a.depthRays && threadIdx.x < kDepthSlots && depthCount[threadIdx.x]) atomicAdd(a.depthRays + threadIdx.x, (unsigned long long)depthCount[threadIdx.x]);
	if (a.counters) {
		unsigned long long rays = cRays, hits = cHits;
		for (int o = 16; o; o >>= 1) {
			rays += __shfl_xor_sync(kFullMask, rays, o);
			hits += __shfl_xor_sync(kFullMask, hits, o);
		}
		if (lane == 0) {
			atomicAdd(a.counters + 0, rays);
			atomicAdd(a.counters + 1, hits);
		}
	}
}

template <int kSmStack>
cudaError_t launchStream(const StreamArgs& a, const Tuning& t, int smCount, cudaStream_t stream) {
	constexpr int kBlock = 256, kMinBlocks = 5;
	auto kernel = pathStreamKernel<kBlock, kMinBlocks, kSmStack>;
	static thread_local int plannedDevice = -1, resident = 1;
	int device = 0;
	cudaGetDevice(&device);
	cudaError_t err;
	if (plannedDevice != device) {
		int carve = 0;
		int smPerSm = 233472;
		cudaDeviceGetAttribute(&smPerSm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
		carve = (int)(((size_t)(kSmStack * kBlock * 4 + kDepthSlots * 4 + 1024) * kMinBlocks * 100 + smPerSm - 1) / smPerSm);
		if (carve > 100) carve = 100;
		err = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
		if (err != cudaSuccess) return err;
		err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kernel, kBlock, 0);
		if (err != cudaSuccess) return err;
		if (resident < 1) resident = 1;
		plannedDevice = device;
	}
	long long grid = (long long)smCount * resident;
	const long long needed = ((long long)a.paths + kBlock - 1) / kBlock;
	if (grid > needed) grid = needed;
	err = cudaMemsetAsync(a.ctrl, 0, 4 * sizeof(uint32_t), stream);
	if (err != cudaSuccess) return err;
	kernel<<<(unsigned)grid, kBlock, 0, stream>>>(a, t.pathStreamThreshold < 1 ? 1 : (t.pathStreamThreshold > 32 ? 32 : t.pathStreamThreshold), t.innerBail < 1 ? 1 : t.innerBail, t.leafBail);
	return cudaGetLastError();
}

} // namespace

uint32_t pathStreamMaxPaths() { return 1u << kPathBits; }

cudaError_t launchPathStream(const PathStreamParams& p, const Tuning& t, int smCount, cudaStream_t stream, int* launches) {
	if (!p.paths) return cudaSuccess;
	StreamArgs a;
	a.tnodes = p.tnodes; a.tpairs = p.tpairs; a.remap = p.remap; a.envPairs = p.envPairs; a.envWidth = p.envWidth; a.envHeight = p.envHeight;
	a.scene.indices = p.indices; a.scene.normals = p.normals; a.scene.triangleNormals = p.triangleNormals;
	a.scene.triangleMaterials = p.triangleMaterials; a.scene.materials = p.materials;
	a.scene.triangleCount = p.triangleCount; a.scene.materialCount = p.materialCount;
	a.cam = cameraArgs(p.camera12);
	a.width = p.width; a.pixels = p.pixels; a.sampleBase = p.sampleBase; a.seed = p.seed; a.maxDepth = p.maxDepth;
	a.firstPath = p.firstPath; a.paths = p.paths;
	a.queue = p.queue; a.flags = p.flags; a.capacity = p.capacity; a.epoch = p.epoch; a.ctrl = p.ctrl;
	a.radiance = p.radiance; a.depthRays = p.depthRays; a.counters = p.counters;
	if (launches) *launches += 1;
	return p.smemStack > 0 ? launchStream<16>(a, t, smCount, stream) : launchStream<0>(a, t, smCount, stream);
}

} // namespace racc_b200
