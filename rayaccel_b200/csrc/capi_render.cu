// capi_render.cu -- the C-ABI's device-side renderers: racc_cuda_shading_*, racc_cuda_path_trace (pathtrace.cu) and
// racc_cuda_whitted_trace (whitted.cu). They are additions beside the drop-in boundary (the renderers are client code in
// the reference, Renderer/*.cpp); nothing in capi.cu depends on them. Shared state: capi_internal.h.
#include "capi_internal.h"

#include <cstdlib>

using namespace racc_b200;

namespace racc_b200 {
namespace {

int envIntRender(const char* name, int fallback) {
	const char* v = getenv(name);
	return v && *v ? atoi(v) : fallback;
}

template <typename T>
T* perDevice(std::vector<std::unique_ptr<T>>& v, int ordinal) {
	for (auto& p : v)
		if (p->device == ordinal) return p.get();
	v.emplace_back(new T());
	v.back()->device = ordinal;
	return v.back().get();
}

// Internal streams of racc_cuda_path_trace (one set per calling host thread and device), see there.
struct PathLanes {
	static constexpr int kMax = 4;
	static constexpr int kDepths = 64; // wave sizes per lane (max_depth <= 62)
	bool ready = false;
	int device = -1;
	int count = 2;
	cudaStream_t stream[kMax] = {};
	cudaEvent_t done[kMax] = {};
	cudaEvent_t fork = nullptr;
	uint32_t* hostCounts = nullptr; // pinned: the wave sizes of each lane

	int init() {
		if (ready) return 0;
		release();
		count = envIntRender("RACC_B200_PATH_LANES", 2);
		if (count < 1) count = 1;
		if (count > kMax) count = kMax;
		RACC_CUDA_CHECK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
		RACC_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&hostCounts), kMax * kDepths * sizeof(uint32_t), cudaHostAllocPortable));
		for (int l = 0; l < kMax; ++l) {
			RACC_CUDA_CHECK(cudaStreamCreateWithFlags(&stream[l], cudaStreamNonBlocking));
			RACC_CUDA_CHECK(cudaEventCreateWithFlags(&done[l], cudaEventDisableTiming));
		}
		ready = true;
		return 0;
	}
	void release() {
		if (device >= 0) cudaSetDevice(device);
		for (int l = 0; l < kMax; ++l) {
			if (stream[l]) { cudaStreamSynchronize(stream[l]); cudaStreamDestroy(stream[l]); }
			if (done[l]) cudaEventDestroy(done[l]);
			stream[l] = nullptr; done[l] = nullptr;
		}
		if (fork) cudaEventDestroy(fork);
		if (hostCounts) cudaFreeHost(hostCounts);
		fork = nullptr; hostCounts = nullptr;
		ready = false;
	}
	~PathLanes() { /* see HostPipeline: released by racc_cuda_thread_release, else leaked on purpose */ }
};

thread_local std::vector<std::unique_ptr<PathLanes>> t_pathLanes;

// Wave buffers of racc_cuda_whitted_trace kept between waves, batches and calls (Tuning::whittedArena), one set per
// calling host thread and device. A wave's size is only known after the previous one was shaded and differs from frame to
// frame, so per-wave stream-ordered allocations keep asking the pool for sizes it has no block for (measured: 24-31 ms
// per 1920x1080x4spp frame against 11.2 ms with these buffers); they only ever grow (by a quarter more than asked).
// Slots: 0/1 rays (ping-pong), 2/3 states, 4 results. A buffer is grown only while it holds nothing live: the results
// before a wave is traced, the next wave's rays/states before they are written.
struct WhittedArena {
	static constexpr int kSlots = 5;
	int device = -1;
	void* p[kSlots] = {};
	size_t cap[kSlots] = {};
	cudaEvent_t idle = nullptr; // end of the previous call's work on these buffers
	bool idleRecorded = false;

	// a later call may come on another CUDA stream: it waits for the previous call's kernels before touching the buffers
	int begin(cudaStream_t stream) {
		if (!idle) RACC_CUDA_CHECK(cudaEventCreateWithFlags(&idle, cudaEventDisableTiming));
		if (idleRecorded) RACC_CUDA_CHECK(cudaStreamWaitEvent(stream, idle, 0));
		return 0;
	}
	int end(cudaStream_t stream) {
		RACC_CUDA_CHECK(cudaEventRecord(idle, stream));
		idleRecorded = true;
		return 0;
	}
	int ensure(int k, size_t bytes, cudaStream_t stream) {
		if (cap[k] >= bytes && p[k]) return 0;
		if (p[k]) RACC_CUDA_CHECK(cudaFreeAsync(p[k], stream));
		p[k] = nullptr;
		cap[k] = 0;
		const size_t want = ((bytes + bytes / 4 + (2u << 20)) >> 21) << 21; // a quarter of headroom, whole 2 MiB pages
		RACC_CUDA_CHECK(cudaMallocAsync(&p[k], want, stream));
		cap[k] = want;
		return 0;
	}
	void release() {
		if (device >= 0) cudaSetDevice(device);
		if (idle) { cudaEventSynchronize(idle); cudaEventDestroy(idle); }
		for (int k = 0; k < kSlots; ++k) {
			if (p[k]) cudaFree(p[k]);
			p[k] = nullptr; cap[k] = 0;
		}
		idle = nullptr;
		idleRecorded = false;
	}
	~WhittedArena() { /* see HostPipeline */ }
};

thread_local std::vector<std::unique_ptr<WhittedArena>> t_whittedArenas;

// What the streamed form of racc_cuda_path_trace (pathstream.cu) keeps between calls, one set per calling host thread and
// device: the queue's flag words -- they hold the epoch numbers of earlier launches, so that no launch has to clear them --
// its four control words and the per-bounce ray counters.
struct StreamArena {
	int device = -1;
	uint32_t* flags = nullptr;
	size_t flagWords = 0;
	uint32_t epoch = 0;
	uint32_t* ctrl = nullptr;                 // device, 4 words
	unsigned long long* depthRays = nullptr;  // device, PathLanes::kDepths
	unsigned long long* hostDepthRays = nullptr; // pinned
	cudaEvent_t idle = nullptr;               // end of the previous call's work
	bool idleRecorded = false;

	int begin(cudaStream_t stream) {
		if (!idle) {
			RACC_CUDA_CHECK(cudaEventCreateWithFlags(&idle, cudaEventDisableTiming));
			RACC_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&ctrl), 4 * sizeof(uint32_t)));
			RACC_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&depthRays), PathLanes::kDepths * sizeof(unsigned long long)));
			RACC_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&hostDepthRays), PathLanes::kDepths * sizeof(unsigned long long), cudaHostAllocPortable));
		}
		if (idleRecorded) RACC_CUDA_CHECK(cudaStreamWaitEvent(stream, idle, 0));
		return 0;
	}
	int end(cudaStream_t stream) {
		RACC_CUDA_CHECK(cudaEventRecord(idle, stream));
		idleRecorded = true;
		return 0;
	}
	// grows only; new words are zero, which no epoch equals
	int ensureFlags(size_t words) {
		if (words <= flagWords) return 0;
		if (idleRecorded) RACC_CUDA_CHECK(cudaEventSynchronize(idle));
		if (flags) RACC_CUDA_CHECK(cudaFree(flags));
		flags = nullptr;
		flagWords = 0;
		const size_t want = ((words + words / 4 + (1u << 19)) >> 19) << 19;
		RACC_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&flags), want * sizeof(uint32_t)));
		RACC_CUDA_CHECK(cudaMemset(flags, 0, want * sizeof(uint32_t)));
		flagWords = want;
		epoch = 0;
		return 0;
	}
	uint32_t nextEpoch() {
		if (++epoch == 0) { // wrapped: every stale value could collide again
			cudaMemset(flags, 0, flagWords * sizeof(uint32_t));
			epoch = 1;
		}
		return epoch;
	}
	void release() {
		if (device >= 0) cudaSetDevice(device);
		if (idle) { cudaEventSynchronize(idle); cudaEventDestroy(idle); }
		if (flags) cudaFree(flags);
		if (ctrl) cudaFree(ctrl);
		if (depthRays) cudaFree(depthRays);
		if (hostDepthRays) cudaFreeHost(hostDepthRays);
		flags = nullptr; ctrl = nullptr; depthRays = nullptr; hostDepthRays = nullptr; idle = nullptr;
		flagWords = 0;
		idleRecorded = false;
	}
	~StreamArena() { /* see HostPipeline */ }
};

thread_local std::vector<std::unique_ptr<StreamArena>> t_streamArenas;

// racc_cuda_path_trace as one persistent kernel per batch (pathstream.cu): arguments checked by the caller.
int pathTraceStreamed(DeviceState* dev, const Tuning& tuning, racc_cuda_scene* s, racc_cuda_env* env, const SceneReplica* rep, const EnvReplica* erep,
                      const racc_cuda_shading* sh, const ShadingReplica* shr, const racc_cuda_camera* camera, const racc_cuda_path_desc* d,
                      float* framebuffer4, uint64_t* wave_rays, cudaStream_t stream, uint64_t maxBatchPaths) {
	const uint64_t pixels = (uint64_t)d->width * d->height;
	StreamArena& ar = *perDevice(t_streamArenas, dev->ordinal);
	if (ar.begin(stream)) return -1;
	uint64_t batchSpp = d->batch_spp ? d->batch_spp : maxBatchPaths / pixels;
	if (batchSpp > maxBatchPaths / pixels) batchSpp = maxBatchPaths / pixels;
	if (batchSpp > d->spp) batchSpp = d->spp;
	if (batchSpp < 1) batchSpp = 1;
	const size_t paths = (size_t)(pixels * batchSpp);
	const size_t capacity = paths * d->max_depth;
	if (ar.ensureFlags(capacity)) return -1;
	const bool hostFb = (d->flags & RACC_CUDA_FRAMEBUFFER_HOST) != 0;

	struct Buffers {
		cudaStream_t stream;
		void* p[3] = {};
		~Buffers() { for (void* q : p) if (q) cudaFreeAsync(q, stream); }
	} buf;
	buf.stream = stream;
	const size_t sizes[3] = {capacity * 48 + 48, paths * 16, hostFb ? (size_t)pixels * 16 : 0};
	for (int k = 0; k < 3; ++k)
		if (sizes[k]) RACC_CUDA_CHECK(cudaMallocAsync(&buf.p[k], sizes[k], stream));
	float4* radiance = static_cast<float4*>(buf.p[1]);
	float4* fb = hostFb ? static_cast<float4*>(buf.p[2]) : reinterpret_cast<float4*>(framebuffer4);
	if (hostFb) RACC_CUDA_CHECK(cudaMemcpyAsync(fb, framebuffer4, (size_t)pixels * 16, cudaMemcpyHostToDevice, stream));
	if (wave_rays) RACC_CUDA_CHECK(cudaMemsetAsync(ar.depthRays, 0, PathLanes::kDepths * sizeof(unsigned long long), stream));

	Tuning t = tuning;
	if (t.smemStack < 0) t.smemStack = sceneExceedsL2(s) ? 16 : 0;
	int launches = 0;
	for (uint32_t done = 0; done < d->spp; done += (uint32_t)batchSpp) {
		const uint32_t spp = d->spp - done < batchSpp ? d->spp - done : (uint32_t)batchSpp;
		const uint32_t batchPaths = (uint32_t)(pixels * spp);
		RACC_CUDA_CHECK(cudaMemsetAsync(radiance, 0, (size_t)batchPaths * 16, stream));
		PathStreamParams p{};
		p.tnodes = rep->dTNodes; p.tpairs = rep->dTPairs; p.remap = rep->dRemap;
		p.envPairs = erep ? erep->dTexelPairs : nullptr; p.envWidth = env ? env->width : 0; p.envHeight = env ? env->height : 0;
		p.indices = rep->dIndices; p.normals = shr->dNormals; p.triangleNormals = shr->dTriangleNormals;
		p.triangleMaterials = shr->dTriangleMaterials; p.materials = shr->dMaterials;
		p.triangleCount = sh->triangleCount; p.materialCount = sh->materialCount;
		p.camera12 = camera->origin;
		p.width = d->width; p.pixels = (uint32_t)pixels; p.sampleBase = d->sample_base + done; p.seed = d->seed; p.maxDepth = d->max_depth;
		p.firstPath = 0; p.paths = batchPaths;
		p.queue = static_cast<float4*>(buf.p[0]); p.flags = ar.flags; p.capacity = (uint32_t)capacity; p.epoch = ar.nextEpoch(); p.ctrl = ar.ctrl;
		p.radiance = radiance; p.depthRays = wave_rays ? ar.depthRays : nullptr; p.counters = dev->dFrame;
		p.smemStack = t.smemStack;
		RACC_CUDA_CHECK(launchPathStream(p, t, dev->smCount, stream, &launches));
		RACC_CUDA_CHECK(launchPathAccumulate(radiance, (uint32_t)pixels, spp, fb, stream, &launches));
	}
	countLaunches(launches);
	if (wave_rays) {
		// the one wait of the call, and only because the caller asked for the ray counts
		RACC_CUDA_CHECK(cudaMemcpyAsync(ar.hostDepthRays, ar.depthRays, (d->max_depth + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
		RACC_CUDA_CHECK(cudaStreamSynchronize(stream));
		for (uint32_t k = 0; k <= d->max_depth; ++k) wave_rays[k] += ar.hostDepthRays[k];
	}
	if (hostFb) {
		RACC_CUDA_CHECK(cudaMemcpyAsync(framebuffer4, fb, (size_t)pixels * 16, cudaMemcpyDeviceToHost, stream));
		RACC_CUDA_CHECK(cudaStreamSynchronize(stream));
	}
	return ar.end(stream);
}

} // namespace

void releaseRenderScratch() {
	for (auto& p : t_pathLanes) p->release();
	t_pathLanes.clear();
	for (auto& p : t_whittedArenas) p->release();
	t_whittedArenas.clear();
	for (auto& p : t_streamArenas) p->release();
	t_streamArenas.clear();
}

} // namespace racc_b200

extern "C" {

// ---- device-side wavefront path tracer (pathtrace.cu) ----

racc_cuda_shading* racc_cuda_shading_create(const racc_cuda_shading_desc* d) {
	if (!d || !d->normals4 || !d->triangle_normals4 || !d->triangle_materials || !d->materials_ke4) {
		fail("racc_cuda_shading_create: null input");
		return nullptr;
	}
	if (!d->material_count) { fail("racc_cuda_shading_create: no materials"); return nullptr; }
	if (!currentDevice()) return nullptr;
	const std::vector<int> set = currentDeviceSet();
	racc_cuda_shading* sh = new racc_cuda_shading();
	sh->vertexCount = d->vertex_count;
	sh->triangleCount = d->triangle_count;
	sh->materialCount = d->material_count;
	for (int ordinal : set) {
		if (!useDevice(ordinal)) { racc_cuda_shading_destroy(sh); return nullptr; }
		sh->replicas.emplace_back(new ShadingReplica());
		ShadingReplica* r = sh->replicas.back().get();
		r->device = ordinal;
		cudaError_t e;
#define UP(dst, src, bytes)                                                                                 \
		if ((e = cudaMalloc(reinterpret_cast<void**>(&dst), (bytes) != 0 ? (bytes) : 16)) != cudaSuccess ||     \
		    (e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice)) != cudaSuccess) {                         \
			fail("racc_cuda_shading_create: upload failed: %s", cudaGetErrorString(e));                         \
			racc_cuda_shading_destroy(sh);                                                                      \
			cudaSetDevice(set[0]);                                                                              \
			return nullptr;                                                                                     \
		}
		UP(r->dNormals, d->normals4, (size_t)d->vertex_count * 16)
		UP(r->dTriangleNormals, d->triangle_normals4, (size_t)d->triangle_count * 16)
		UP(r->dTriangleMaterials, d->triangle_materials, (size_t)d->triangle_count * 2)
		UP(r->dMaterials, d->materials_ke4, (size_t)d->material_count * 16)
#undef UP
	}
	cudaSetDevice(set[0]);
	return sh;
}

void racc_cuda_shading_destroy(racc_cuda_shading* sh) {
	if (!sh) return;
	for (auto& r : sh->replicas) {
		cudaFree(r->dNormals);
		cudaFree(r->dTriangleNormals);
		cudaFree(r->dTriangleMaterials);
		cudaFree(r->dMaterials);
	}
	delete sh;
}

int racc_cuda_path_trace(racc_cuda_scene* s, racc_cuda_env* env, const racc_cuda_shading* sh, const racc_cuda_camera* camera,
                         const racc_cuda_path_desc* d, float* framebuffer4, uint64_t* wave_rays, void* cuda_stream) {
	if (!s || !sh || !camera || !d || !framebuffer4) return fail("racc_cuda_path_trace: null argument");
	if (!s->vertexCount) return fail("racc_cuda_path_trace: scene was created from images and has no index data");
	if (sh->triangleCount != s->triangleCount || sh->vertexCount < s->vertexCount)
		return fail("racc_cuda_path_trace: shading data (%u triangles, %u vertices) does not match the scene (%u, %u)", sh->triangleCount,
		            sh->vertexCount, s->triangleCount, s->vertexCount);
	if (d->max_depth > 62) return fail("racc_cuda_path_trace: max_depth %u > 62", d->max_depth);
	DeviceState* dev = currentDevice();
	if (!dev) return -1;
	const Tuning tuning = tuningSnapshot();
	SceneReplica* rep = s->on(dev->ordinal);
	const ShadingReplica* shr = sh->on(dev->ordinal);
	if (!rep || !shr) return fail("racc_cuda_path_trace: the scene or the shading data has no copy on CUDA device %d", dev->ordinal);
	if (!rep->dIndices) return fail("racc_cuda_path_trace: scene was created from images and has no index data");
	const uint64_t pixels = (uint64_t)d->width * d->height;
	if (!pixels || !d->spp) return 0;
	if (pixels > (1ull << 28)) return fail("racc_cuda_path_trace: viewport too large");
	cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
	// paths per batch: whole samples, about 128 M paths (128 B of device memory each: 16 GB of a B200's 180; never more than an
	// eighth of the device's memory -- sized from the total noted at device set-up: cudaMemGetInfo costs ~8 ms per call on a
	// 180 GB part with a warm memory pool, measured) unless the caller says otherwise: the deeper waves of a batch are a tenth to a hundredth of its size,
	// and a persistent launch over < 1 M rays is mostly tail (1920x1080, 16 spp: 4.4 / 5.8 / 6.3 Gray/s at 2 M / 8 M / 33 M
	// paths per batch, profiles/r01_render_device.md; 64 spp x 8 bounces: 5.91 / 6.06 / 6.14 Gray/s at 33 M / 66 M / 133 M)
	uint64_t autoPaths = 128ull << 20;
	if (dev->totalBytes / 1024 < autoPaths) autoPaths = dev->totalBytes / 1024; // at most an eighth of the device's memory
	uint32_t batchSpp = d->batch_spp ? d->batch_spp : (uint32_t)(autoPaths / pixels);
	if (batchSpp < 1) batchSpp = 1;
	if (batchSpp > d->spp) batchSpp = d->spp;
	if (pixels * batchSpp > 0x7fffffffull) return fail("racc_cuda_path_trace: batch of %u samples is too large", batchSpp);
	const size_t paths = (size_t)pixels * batchSpp;
	const bool hostFb = (d->flags & RACC_CUDA_FRAMEBUFFER_HOST) != 0;
	// Wave sizes. A wave's size is the compaction counter the previous shading kernel leaves in device memory. The
	// traversal kernel is persistent (its grid does not depend on the ray count) and the shading kernel loops over tiles,
	// so both can be launched from an upper bound -- the lane's path count -- and read the real size on the device: the
	// whole batch is enqueued without one host round trip. The host-synchronised scheme of round 1 (wait for each wave's
	// size, launch exactly that) remains for launches that need the size on the host: re-binned traversal (scenes far
	// larger than L2) and tuning key 18.
	const bool rebinned = (tuning.variant == 3 || tuning.variant == 4) && (tuning.sortMode == 1 || (tuning.sortMode == 2 && sceneExceedsL2(s)));
	const bool hostSizes = tuning.pathSync != 0 || rebinned;
	// The streamed form (Tuning::pathStream, pathstream.cu): one persistent kernel per batch traces, shades and queues. It
	// walks the exact packed images in arrival order, so the other forms keep the wavefront scheme below. A batch needs
	// 16 B per path and 48 B per path and bounce; it stays within an eighth of the device's memory.
	if (tuning.pathStream != 0 && tuning.variant == 3 && !hostSizes) {
		uint64_t maxBatchPaths = (dev->totalBytes / 8) / (16 + 48ull * d->max_depth);
		if (maxBatchPaths > pathStreamMaxPaths()) maxBatchPaths = pathStreamMaxPaths();
		if (d->max_depth && maxBatchPaths > 0xfffffff0ull / d->max_depth) maxBatchPaths = 0xfffffff0ull / d->max_depth;
		const EnvReplica* erep = env ? env->on(dev->ordinal) : nullptr;
		if (env && !erep) return fail("racc_cuda_path_trace: the environment has no copy on CUDA device %d", dev->ordinal);
		if (pixels <= maxBatchPaths)
			return pathTraceStreamed(dev, tuning, s, env, rep, erep, sh, shr, camera, d, framebuffer4, wave_rays, stream, maxBatchPaths);
	}

	// Lanes: a batch is cut into contiguous path ranges that advance bounce by bounce on their own streams, so that the
	// tail of one lane's traversal launch (few long paths left, most SMs idle) is filled by the other lane's kernels.
	PathLanes& lanes = *perDevice(t_pathLanes, dev->ordinal);
	if (lanes.init()) return -1;
	const int nlanes = paths >= (size_t)lanes.count * 65536 ? lanes.count : 1;
	const size_t lanePaths = (paths + nlanes - 1) / nlanes;

	struct Buffers {
		cudaStream_t stream;
		void* p[8] = {};
		bool joined = false; // every lane's work is ordered before `stream`; false on an error return
		~Buffers() {
			if (!joined) cudaDeviceSynchronize(); // lanes may still be using the buffers
			for (void* q : p) if (q) cudaFreeAsync(q, stream);
		}
	} buf;
	buf.stream = stream;
	const size_t sizes[8] = {lanePaths * nlanes * 32, lanePaths * nlanes * 32, lanePaths * nlanes * 16, lanePaths * nlanes * 16,
	                         lanePaths * nlanes * 16, paths * 16, (size_t)PathLanes::kMax * PathLanes::kDepths * sizeof(uint32_t), hostFb ? (size_t)pixels * 16 : 0};
	for (int k = 0; k < 8; ++k)
		if (sizes[k]) RACC_CUDA_CHECK(cudaMallocAsync(&buf.p[k], sizes[k], stream));
	float4* radiance = static_cast<float4*>(buf.p[5]);
	uint32_t* counts = static_cast<uint32_t*>(buf.p[6]);
	float4* fb = hostFb ? static_cast<float4*>(buf.p[7]) : reinterpret_cast<float4*>(framebuffer4);
	if (hostFb) RACC_CUDA_CHECK(cudaMemcpyAsync(fb, framebuffer4, (size_t)pixels * 16, cudaMemcpyHostToDevice, stream));

	struct Lane {
		DevRay* rays[2];
		float4* states[2];
		float4* results;
		uint32_t* counts; // device, one per depth: counts[k] = size of wave k+1
		cudaStream_t stream;
		uint32_t first;   // size of wave 0 (known on the host)
		uint32_t count, depth;
		int cur;
		bool busy;
	} lane[PathLanes::kMax];
	for (int l = 0; l < nlanes; ++l) {
		lane[l].rays[0] = static_cast<DevRay*>(buf.p[0]) + (size_t)l * lanePaths;
		lane[l].rays[1] = static_cast<DevRay*>(buf.p[1]) + (size_t)l * lanePaths;
		lane[l].states[0] = static_cast<float4*>(buf.p[2]) + (size_t)l * lanePaths;
		lane[l].states[1] = static_cast<float4*>(buf.p[3]) + (size_t)l * lanePaths;
		lane[l].results = static_cast<float4*>(buf.p[4]) + (size_t)l * lanePaths;
		lane[l].counts = counts + (size_t)l * PathLanes::kDepths;
		lane[l].stream = nlanes > 1 ? lanes.stream[l] : stream;
	}

	int launches = 0;
	for (uint32_t done = 0; done < d->spp; done += batchSpp) {
		const uint32_t spp = d->spp - done < batchSpp ? d->spp - done : batchSpp;
		const uint32_t sampleBase = d->sample_base + done;
		const uint32_t batchPaths = (uint32_t)(pixels * spp);
		RACC_CUDA_CHECK(cudaMemsetAsync(radiance, 0, (size_t)batchPaths * 16, stream));
		RACC_CUDA_CHECK(cudaMemsetAsync(counts, 0, (size_t)PathLanes::kMax * PathLanes::kDepths * sizeof(uint32_t), stream));
		if (nlanes > 1) {
			RACC_CUDA_CHECK(cudaEventRecord(lanes.fork, stream));
			for (int l = 0; l < nlanes; ++l) RACC_CUDA_CHECK(cudaStreamWaitEvent(lane[l].stream, lanes.fork, 0));
		}
		// one wave of one lane: trace, then shade + compact. upper: how many rays the launches are sized for; size: where the
		// wave's real size is on the device (null: it is `upper`)
		auto enqueue = [&](Lane& ln, uint32_t upper, const uint32_t* size) -> int {
			racc_cuda_stream_desc sd{};
			sd.rays = ln.rays[ln.cur];
			sd.results = ln.results;
			sd.count = upper;
			sd.flags = 0;
			if (traceImpl(s, env, &sd, 1, ln.stream, nullptr, false, size, nlanes > 1 ? tuning.pathTraceCtas : 0)) return -1;
			PathShadeParams p{};
			p.rays = ln.rays[ln.cur]; p.results = ln.results; p.states = ln.states[ln.cur]; p.count = upper; p.countPtr = size;
			p.gridLimit = (uint32_t)dev->smCount * 8u;
			p.depth = ln.depth; p.maxDepth = d->max_depth; p.seed = d->seed; p.pixels = (uint32_t)pixels; p.sampleBase = sampleBase;
			p.indices = rep->dIndices; p.normals = shr->dNormals; p.triangleNormals = shr->dTriangleNormals;
			p.triangleMaterials = shr->dTriangleMaterials; p.materials = shr->dMaterials;
			p.triangleCount = sh->triangleCount; p.materialCount = sh->materialCount;
			p.outRays = ln.rays[ln.cur ^ 1]; p.outStates = ln.states[ln.cur ^ 1]; p.outCount = ln.counts + ln.depth; p.radiance = radiance;
			RACC_CUDA_CHECK(launchPathShade(p, ln.stream, &launches));
			return 0;
		};
		int active = 0;
		for (int l = 0; l < nlanes; ++l) {
			Lane& ln = lane[l];
			const size_t first = (size_t)l * lanePaths;
			ln.first = ln.count = first < batchPaths ? (uint32_t)(batchPaths - first < lanePaths ? batchPaths - first : lanePaths) : 0;
			ln.depth = 0;
			ln.cur = 0;
			ln.busy = ln.count != 0;
			if (!ln.busy) continue;
			RACC_CUDA_CHECK(launchPathPrimary(camera->origin, d->width, d->height, sampleBase, (uint32_t)first, ln.count, d->seed, ln.rays[0],
			                                  ln.states[0], ln.stream, &launches));
			if (!hostSizes) {
				// every wave of the lane, back to back; a wave can only shrink, so the lane's path count bounds them all
				for (ln.depth = 0; ln.depth <= d->max_depth; ++ln.depth, ln.cur ^= 1)
					if (enqueue(ln, ln.first, ln.depth ? ln.counts + ln.depth - 1 : nullptr)) return -1;
				if (wave_rays && d->max_depth)
					RACC_CUDA_CHECK(cudaMemcpyAsync(lanes.hostCounts + (size_t)l * PathLanes::kDepths, ln.counts, d->max_depth * sizeof(uint32_t),
					                                cudaMemcpyDeviceToHost, ln.stream));
				continue;
			}
			if (wave_rays) wave_rays[0] += ln.count;
			if (enqueue(ln, ln.count, nullptr)) return -1;
			if (ln.depth == d->max_depth) ln.busy = false; // nothing is extended past the last bounce
			else RACC_CUDA_CHECK(cudaMemcpyAsync(lanes.hostCounts + (size_t)l * PathLanes::kDepths, ln.counts + ln.depth, sizeof(uint32_t), cudaMemcpyDeviceToHost, ln.stream));
			active += ln.busy;
		}
		// host-sized waves, round robin: the size of a lane's next wave decides its launch -- one host round trip per bounce and lane
		for (int l = 0; hostSizes && active; l = (l + 1) % nlanes) {
			Lane& ln = lane[l];
			if (!ln.busy) continue;
			RACC_CUDA_CHECK(cudaStreamSynchronize(ln.stream));
			ln.count = lanes.hostCounts[(size_t)l * PathLanes::kDepths];
			ln.depth += 1;
			ln.cur ^= 1;
			if (!ln.count) { ln.busy = false; --active; continue; }
			if (wave_rays) wave_rays[ln.depth] += ln.count;
			if (enqueue(ln, ln.count, nullptr)) return -1;
			if (ln.depth == d->max_depth) { ln.busy = false; --active; continue; }
			RACC_CUDA_CHECK(cudaMemcpyAsync(lanes.hostCounts + (size_t)l * PathLanes::kDepths, ln.counts + ln.depth, sizeof(uint32_t), cudaMemcpyDeviceToHost, ln.stream));
		}
		if (nlanes > 1)
			for (int l = 0; l < nlanes; ++l) {
				RACC_CUDA_CHECK(cudaEventRecord(lanes.done[l], lane[l].stream));
				RACC_CUDA_CHECK(cudaStreamWaitEvent(stream, lanes.done[l], 0));
			}
		RACC_CUDA_CHECK(launchPathAccumulate(radiance, (uint32_t)pixels, spp, fb, stream, &launches));
		if (!hostSizes && wave_rays) {
			// the only wait of the batch, and only because the caller asked for the ray counts
			RACC_CUDA_CHECK(cudaStreamSynchronize(stream));
			for (int l = 0; l < nlanes; ++l) {
				if (!lane[l].first) continue;
				wave_rays[0] += lane[l].first;
				for (uint32_t k = 0; k < d->max_depth; ++k) wave_rays[k + 1] += lanes.hostCounts[(size_t)l * PathLanes::kDepths + k];
			}
		}
	}
	buf.joined = true;
	countLaunches(launches);
	if (hostFb) {
		RACC_CUDA_CHECK(cudaMemcpyAsync(framebuffer4, fb, (size_t)pixels * 16, cudaMemcpyDeviceToHost, stream));
		RACC_CUDA_CHECK(cudaStreamSynchronize(stream));
	}
	return 0;
}

// The reference's Whitted renderer with the shading on the device (whitted.cu). Same descriptor and framebuffer
// meaning as racc_cuda_path_trace; desc->batch_spp 0 = about 4 M primary rays per batch (a hit spawns up to two rays,
// so waves grow before the 0.3-per-bounce weight ends them).
int racc_cuda_whitted_trace(racc_cuda_scene* s, racc_cuda_env* env, const racc_cuda_shading* sh, const racc_cuda_camera* camera,
                            const racc_cuda_path_desc* d, float* framebuffer4, uint64_t* wave_rays, void* cuda_stream) {
	if (!s || !sh || !camera || !d || !framebuffer4) return fail("racc_cuda_whitted_trace: null argument");
	if (!s->vertexCount) return fail("racc_cuda_whitted_trace: scene was created from images and has no index data");
	if (sh->triangleCount != s->triangleCount || sh->vertexCount < s->vertexCount)
		return fail("racc_cuda_whitted_trace: shading data (%u triangles, %u vertices) does not match the scene (%u, %u)", sh->triangleCount,
		            sh->vertexCount, s->triangleCount, s->vertexCount);
	if (d->max_depth > 62) return fail("racc_cuda_whitted_trace: max_depth %u > 62", d->max_depth);
	DeviceState* dev = currentDevice();
	if (!dev) return -1;
	const Tuning tuning = tuningSnapshot();
	SceneReplica* rep = s->on(dev->ordinal);
	const ShadingReplica* shr = sh->on(dev->ordinal);
	if (!rep || !shr) return fail("racc_cuda_whitted_trace: the scene or the shading data has no copy on CUDA device %d", dev->ordinal);
	if (!rep->dIndices) return fail("racc_cuda_whitted_trace: scene was created from images and has no index data");
	const uint64_t pixels = (uint64_t)d->width * d->height;
	if (!pixels || !d->spp) return 0;
	if (pixels > (1ull << 24)) return fail("racc_cuda_whitted_trace: viewport too large");
	cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
	uint32_t batchSpp = d->batch_spp ? d->batch_spp : (uint32_t)((4ull << 20) / pixels);
	if (batchSpp < 1) batchSpp = 1;
	if (batchSpp > d->spp) batchSpp = d->spp;
	if (pixels * batchSpp > (1ull << 28)) return fail("racc_cuda_whitted_trace: batch of %u samples is too large", batchSpp);
	const bool hostFb = (d->flags & RACC_CUDA_FRAMEBUFFER_HOST) != 0;

	// stream-ordered scratch; everything still held is released on any return
	struct Scratch {
		cudaStream_t stream;
		std::vector<void*> held;
		int get(void** p, size_t bytes) {
			RACC_CUDA_CHECK(cudaMallocAsync(p, bytes ? bytes : 16, stream));
			held.push_back(*p);
			return 0;
		}
		void release(void* p) {
			for (size_t k = 0; k < held.size(); ++k)
				if (held[k] == p) { held.erase(held.begin() + (long)k); cudaFreeAsync(p, stream); return; }
		}
		~Scratch() { for (void* q : held) cudaFreeAsync(q, stream); }
	} scratch;
	scratch.stream = stream;

	unsigned long long* acc = nullptr;
	uint32_t* counts = nullptr;
	float4* fb = reinterpret_cast<float4*>(framebuffer4);
	if (scratch.get(reinterpret_cast<void**>(&acc), (size_t)pixels * 3 * sizeof(unsigned long long))) return -1;
	if (scratch.get(reinterpret_cast<void**>(&counts), 64 * sizeof(uint32_t))) return -1;
	if (hostFb) {
		if (scratch.get(reinterpret_cast<void**>(&fb), (size_t)pixels * 16)) return -1;
		RACC_CUDA_CHECK(cudaMemcpyAsync(fb, framebuffer4, (size_t)pixels * 16, cudaMemcpyHostToDevice, stream));
	}
	RACC_CUDA_CHECK(cudaMemsetAsync(acc, 0, (size_t)pixels * 3 * sizeof(unsigned long long), stream));

	// wave buffers: stream-ordered allocations per wave (default) or the calling thread's grow-only arena
	const bool useArena = tuning.whittedArena != 0;
	WhittedArena& arena = *perDevice(t_whittedArenas, dev->ordinal);
	if (useArena && arena.begin(stream)) return -1;
	int cur = 0; // arena: which half of the ping-pong holds the current wave
	enum { kRays = 0, kStates = 2, kResults = 4 };

	int launches = 0;
	for (uint32_t done = 0; done < d->spp; done += batchSpp) {
		const uint32_t spp = d->spp - done < batchSpp ? d->spp - done : batchSpp;
		uint32_t count = (uint32_t)(pixels * spp);
		RACC_CUDA_CHECK(cudaMemsetAsync(counts, 0, 64 * sizeof(uint32_t), stream));
		DevRay* rays = nullptr;
		float4* states = nullptr;
		if (useArena) {
			if (arena.ensure(kRays + cur, (size_t)count * 32, stream) || arena.ensure(kStates + cur, (size_t)count * 16, stream)) return -1;
			rays = static_cast<DevRay*>(arena.p[kRays + cur]);
			states = static_cast<float4*>(arena.p[kStates + cur]);
		}
		else if (scratch.get(reinterpret_cast<void**>(&rays), (size_t)count * 32) || scratch.get(reinterpret_cast<void**>(&states), (size_t)count * 16)) return -1;
		RACC_CUDA_CHECK(launchWhittedPrimary(camera->origin, d->width, d->height, d->sample_base + done, 0, count, d->seed, rays, states, stream, &launches));
		for (uint32_t depth = 0; depth <= d->max_depth && count; ++depth) {
			if (wave_rays) wave_rays[depth] += count;
			if (count > 0x3fffffffu) return fail("racc_cuda_whitted_trace: a wave of %u rays is too large; lower batch_spp", count);
			float4* results = nullptr;
			DevRay* nextRays = nullptr;
			float4* nextStates = nullptr;
			const bool last = depth == d->max_depth; // nothing is extended past the last bounce
			if (useArena) {
				if (arena.ensure(kResults, (size_t)count * 16, stream)) return -1;
				results = static_cast<float4*>(arena.p[kResults]);
				if (!last) {
					if (arena.ensure(kRays + (cur ^ 1), (size_t)count * 2 * 32, stream) || arena.ensure(kStates + (cur ^ 1), (size_t)count * 2 * 16, stream)) return -1;
					nextRays = static_cast<DevRay*>(arena.p[kRays + (cur ^ 1)]);
					nextStates = static_cast<float4*>(arena.p[kStates + (cur ^ 1)]);
				}
			}
			else {
				if (scratch.get(reinterpret_cast<void**>(&results), (size_t)count * 16)) return -1;
				if (!last && (scratch.get(reinterpret_cast<void**>(&nextRays), (size_t)count * 2 * 32) ||
				              scratch.get(reinterpret_cast<void**>(&nextStates), (size_t)count * 2 * 16))) return -1;
			}
			racc_cuda_stream_desc sd{};
			sd.rays = rays;
			sd.results = results;
			sd.count = count;
			sd.flags = 0;
			if (traceImpl(s, env, &sd, 1, stream, nullptr, false)) return -1;
			WhittedShadeParams p{};
			p.rays = rays; p.results = results; p.states = states; p.count = count; p.depth = depth; p.maxDepth = d->max_depth;
			p.indices = rep->dIndices; p.normals = shr->dNormals; p.triangleNormals = shr->dTriangleNormals; p.triangleCount = sh->triangleCount;
			p.outRays = nextRays; p.outStates = nextStates; p.outCount = counts + depth; p.accumulators = acc;
			p.combine = tuning.whittedCombine != 0;
			RACC_CUDA_CHECK(launchWhittedShade(p, stream, &launches));
			uint32_t next = 0;
			if (!last) {
				// the size of the next wave decides its launch and its buffers: the one host round trip per bounce
				RACC_CUDA_CHECK(cudaMemcpyAsync(&next, counts + depth, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
				RACC_CUDA_CHECK(cudaStreamSynchronize(stream));
			}
			if (useArena) cur ^= 1;
			else {
				scratch.release(results);
				scratch.release(rays);
				scratch.release(states);
			}
			rays = nextRays;
			states = nextStates;
			count = next;
		}
		if (!useArena) {
			if (rays) scratch.release(rays);
			if (states) scratch.release(states);
		}
	}
	if (useArena && arena.end(stream)) return -1;
	RACC_CUDA_CHECK(launchWhittedFinish(acc, (uint32_t)pixels, fb, stream, &launches));
	countLaunches(launches);
	if (hostFb) {
		RACC_CUDA_CHECK(cudaMemcpyAsync(framebuffer4, fb, (size_t)pixels * 16, cudaMemcpyDeviceToHost, stream));
		RACC_CUDA_CHECK(cudaStreamSynchronize(stream));
	}
	return 0;
}

} // extern "C"
