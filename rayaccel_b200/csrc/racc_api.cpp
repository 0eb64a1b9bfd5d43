// racc_api.cpp -- host orchestration behind include/RayAccelerator.h: Context, the ray-stream pool
// and the spawn -> intersect -> shade scheduler. Pure C++17 and a client of the C-ABI in
// include/racc_b200.h only (no CUDA headers here): every device action is a racc_cuda_* call.
//
// WHAT it has to do is the contract of /root/reference/RayAccelerator/RayAccelerator.cpp:48-415,
// 429-788 as seen by a client: streams owned by the library, `spawn` may add
// up to maxRaysPerSpawn rays, `shade` sees [start,end) of a tested stream and may add up to
// end-start rays to an output stream, callbacks run without the scheduler lock on distinct
// `thread` values, render() blocks until no rays are left, Stats counts rays handed to a tester.
//
// HOW differs from the reference, because the device does:
//   * the reference launches ONE stream of <= 27 648 rays per clEnqueueNDRangeKernel + clFinish on
//     zero-copy memory (RayAccelerator.cpp:393-403). A B200 needs ~200 K resident rays to fill 148
//     SMs and sits behind PCIe, so a submitter here drains EVERY ready stream into one
//     racc_cuda_trace() call; the engine packs them into few large staged launches whose H2D,
//     traversal and D2H overlap.
//   * partially filled streams are flushed only when nothing else can make progress (all callback
//     threads idle, nothing being tested), instead of whenever a submitter is idle (:360-363).
//   * there is no CPU tester (:158-244): no CpuTestJob slots, no cpuThreadsTesting.
//   * stream memory is one pinned slab (racc_cuda_host_alloc) instead of CL_MEM_USE_HOST_PTR.
//   * a context may own SEVERAL B200s (racc::cudaDevices): scene and environment are replicated on each, every device
//     gets its own `gpuSubmissionThreads` submitters (one CUDA stream each), ready streams are dealt over the devices
//     that have an idle submitter, and Stats.raysTraced of a frame is the NCCL sum of the devices' frame records
//     (racc_cuda_frame_reduce) -- the "per-frame hit reduction". No ray, node or result crosses between GPUs.
#include "../../include/RayAccelerator.h"
#include "../../include/racc_b200.h"

#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <thread>
#include <vector>

#include <pmmintrin.h>
#include <xmmintrin.h>

namespace racc {

struct Scene {
	racc_cuda_scene* handle;
};

struct Environment {
	racc_cuda_env* handle;
};

struct Context {
	Configuration configuration{};
	std::vector<int> devices;      // CUDA ordinals this context drives; scenes are replicated on all of them
	uint32_t idleSubmitters = 0;   // submitters asleep on `work`

	// stream pool
	std::vector<RayStream> streams;
	uint32_t streamCapacity = 0;
	void* slab = nullptr;

	// The four places a stream can rest in (Context.h:29-32), plus "held by a thread".
	std::vector<uint32_t> idle;     // count == 0
	std::vector<uint32_t> partial;  // 0 < count < rayStreamBatchSize
	std::vector<uint32_t> ready;    // count >= rayStreamBatchSize: wants intersection
	std::vector<uint32_t> tested;   // results valid: wants shading

	std::mutex mutex;
	std::condition_variable work;  // callback threads and submitters sleep here
	std::condition_variable done;  // render() sleeps here

	Scene* scene = nullptr;
	Environment* environment = nullptr;
	RenderCallbacks callbacks{};

	bool quit = false;
	bool frameOpen = false;       // spawn has not returned false yet (moreRaysExist)
	uint32_t busyCallbacks = 0;   // callback threads currently inside spawn/shade
	uint32_t launchesInFlight = 0; // submitters currently inside racc_cuda_trace/sync
	uint32_t raysInFlight = 0;
	uint64_t raysTraced = 0;

	std::vector<std::thread> threads;
};

} // namespace racc

namespace {

using racc::Context;
using racc::RayStream;

void setFlushToZero() {
	// the reference runs every thread with FTZ+DAZ (Threading.h:77-79); shaders inherit it
	_MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
	_MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
}

size_t align4k(size_t v) { return (v + 4095) & ~(size_t)4095; }

// Files a stream by its fill level. Caller holds the lock.
void park(Context* c, uint32_t index) {
	const uint32_t count = c->streams[index].count;
	if (count >= c->configuration.rayStreamBatchSize)
		c->ready.push_back(index);
	else if (count == 0)
		c->idle.push_back(index);
	else
		c->partial.push_back(index);
}

// A stream to append rays to: prefer topping up a partial one. Caller holds the lock.
bool takeOutputStream(Context* c, uint32_t* index) {
	if (!c->partial.empty()) {
		*index = c->partial.back();
		c->partial.pop_back();
		return true;
	}
	if (!c->idle.empty()) {
		*index = c->idle.back();
		c->idle.pop_back();
		return true;
	}
	return false;
}

bool frameFinished(const Context* c) { return !c->frameOpen && c->raysInFlight == 0; }

// spawn step (contract of RayAccelerator.cpp:48-90). Returns false if nothing could be spawned.
bool spawnStep(Context* c, unsigned thread, std::unique_lock<std::mutex>& lock) {
	const uint32_t quota = c->configuration.maxRaysPerSpawn;
	if (!c->frameOpen || c->raysInFlight + quota > c->configuration.maxRaysInFlight)
		return false;
	uint32_t index;
	if (!takeOutputStream(c, &index))
		return false;
	RayStream* stream = &c->streams[index];
	const racc::RenderCallbacks cb = c->callbacks;
	c->raysInFlight += quota; // reserve the worst case while the callback runs unlocked
	++c->busyCallbacks;
	lock.unlock();

	const uint32_t before = stream->count;
	const bool more = cb.spawn(cb.data, thread, stream);
	const uint32_t added = stream->count - before;
	if (added > quota || stream->count > c->streamCapacity) {
		fprintf(stderr, "RayAccelerator: spawn callback added %u rays (limit %u)\n", added, quota);
		abort();
	}

	lock.lock();
	--c->busyCallbacks;
	c->raysInFlight -= quota - added;
	park(c, index);
	if (!more)
		c->frameOpen = false;
	c->work.notify_all();
	if (frameFinished(c))
		c->done.notify_all();
	return true;
}

// shade step (contract of RayAccelerator.cpp:92-156).
bool shadeStep(Context* c, unsigned thread, std::unique_lock<std::mutex>& lock) {
	if (c->tested.empty())
		return false;
	const uint32_t index = c->tested.back();
	c->tested.pop_back();
	RayStream* input = &c->streams[index];
	const racc::RenderCallbacks cb = c->callbacks;
	const uint32_t slice = c->configuration.cpuShadeBatch;
	const uint32_t count = input->count;
	++c->busyCallbacks;

	for (uint32_t start = 0; start < count; start += slice) {
		const uint32_t end = std::min(count, start + slice);
		uint32_t outIndex;
		while (!takeOutputStream(c, &outIndex))
			c->work.wait(lock); // cannot happen with the pool sizing of createContext; be safe
		RayStream* output = &c->streams[outIndex];
		lock.unlock();

		const uint32_t before = output->count;
		cb.shade(cb.data, thread, input, start, end, output);
		const uint32_t added = output->count - before;
		if (added > end - start || output->count > c->streamCapacity) {
			fprintf(stderr, "RayAccelerator: shade callback added %u rays for %u inputs\n", added, end - start);
			abort();
		}

		lock.lock();
		c->raysInFlight += added;
		park(c, outIndex);
		c->work.notify_all();
	}

	--c->busyCallbacks;
	c->raysInFlight -= count;
	input->count = 0;
	c->idle.push_back(index);
	c->work.notify_all();
	if (frameFinished(c))
		c->done.notify_all();
	return true;
}

void callbackThread(Context* c, unsigned thread) {
	setFlushToZero();
	std::unique_lock<std::mutex> lock(c->mutex);
	while (!c->quit) {
		// hybrid order of the reference (RayAccelerator.cpp:272-279): keep the device fed first
		if (spawnStep(c, thread, lock))
			continue;
		if (shadeStep(c, thread, lock))
			continue;
		c->work.wait(lock);
	}
}

// How many of `ready` full streams one submitter takes: with several devices a burst is split so that every device
// with a sleeping submitter gets a share, but never into pieces too small to fill 148 SMs (four streams ~ 200 K rays).
size_t submitterShare(const Context* c, size_t ready) {
	const size_t others = std::min<size_t>(c->idleSubmitters, c->devices.size() - 1);
	const size_t share = std::max<size_t>(4, (ready + others) / (others + 1));
	return std::min(ready, share);
}

// The device boundary (replaces gpuWorkerThread, RayAccelerator.cpp:335-414). One or more per device of the context.
void submitterThread(Context* c, int device) {
	setFlushToZero();
	if (racc_cuda_init(&device, 1)) {
		fprintf(stderr, "RayAccelerator: %s\n", racc_cuda_last_error());
		abort();
	}
	void* cudaStream = racc_cuda_stream_create();
	if (!cudaStream) {
		fprintf(stderr, "RayAccelerator: %s\n", racc_cuda_last_error());
		abort();
	}
	std::vector<uint32_t> batch;
	std::vector<racc_cuda_stream_desc> descs;

	std::unique_lock<std::mutex> lock(c->mutex);
	while (!c->quit) {
		batch.clear();
		if (!c->ready.empty()) {
			// full streams go into one aggregated launch; with several devices, this one's share of them
			const size_t take = c->devices.size() > 1 ? submitterShare(c, c->ready.size()) : c->ready.size();
			batch.assign(c->ready.end() - (std::ptrdiff_t)take, c->ready.end());
			c->ready.resize(c->ready.size() - take);
		}
		else if (!c->partial.empty() && c->busyCallbacks == 0 && c->tested.empty() && c->launchesInFlight == 0) {
			batch.swap(c->partial); // nothing else can make progress: flush the stragglers
		}
		else {
			++c->idleSubmitters;
			c->work.wait(lock);
			--c->idleSubmitters;
			continue;
		}

		descs.clear();
		uint64_t rays = 0;
		for (uint32_t index : batch) {
			const RayStream& s = c->streams[index];
			descs.push_back(racc_cuda_stream_desc{s.rays, s.results, s.count, RACC_CUDA_STREAM_HOST});
			rays += s.count;
		}
		c->raysTraced += rays;
		racc_cuda_scene* scene = c->scene ? c->scene->handle : nullptr;
		racc_cuda_env* env = c->environment ? c->environment->handle : nullptr;
		++c->launchesInFlight;
		lock.unlock();

		if (racc_cuda_trace(scene, env, descs.data(), (uint32_t)descs.size(), cudaStream) || racc_cuda_sync(cudaStream)) {
			// render() cannot report failure (RayAccelerator.h:115) and wrong hits must not pass silently
			fprintf(stderr, "RayAccelerator: %s\n", racc_cuda_last_error());
			abort();
		}

		lock.lock();
		--c->launchesInFlight;
		c->tested.insert(c->tested.end(), batch.begin(), batch.end());
		c->work.notify_all();
	}
	lock.unlock();
	racc_cuda_stream_destroy(cudaStream);
	racc_cuda_thread_release(); // this thread's staging pipeline: three lanes of device buffers (clReleaseCommandQueue, :774-779)
}

// racc::cudaDevice / racc::cudaDevices tokens: bits 0-7 first ordinal + 1, bits 8-15 device count (0 = one)
std::vector<int> devicesOf(cl_context token) {
	const uintptr_t v = reinterpret_cast<uintptr_t>(token);
	const int first = (int)(v & 0xff) - 1;
	const int count = (int)((v >> 8) & 0xff);
	std::vector<int> out;
	for (int k = 0; k < (count ? count : 1); ++k) out.push_back(first + k);
	return out;
}

int bindAll(const Context* c) { return racc_cuda_init(c->devices.data(), (int)c->devices.size()); }

} // namespace

void racc::init() {
	setFlushToZero(); // RayAccelerator.cpp:417-420; the Embree start-up (:422) has no counterpart
}

void racc::deinit() {}

racc::Configuration racc::defaultConfiguration(cl_context gpuContext) {
	// Sized for a B200 behind PCIe Gen5 rather than for an 8 960-work-item iGPU
	// (RayAccelerator.cpp:429-446): streams as large as the uint16 fields allow and 2 M rays alive,
	// so that one aggregated launch carries several hundred thousand rays.
	Configuration cfg = {};
	cfg.gpuContext = gpuContext;
	cfg.allowCpuTracing = false;
	unsigned cores = std::thread::hardware_concurrency();
	if (cores == 0) cores = 1;
	cfg.gpuSubmissionThreads = 2;
	unsigned callbackThreads = cores > cfg.gpuSubmissionThreads ? cores - cfg.gpuSubmissionThreads : 1;
	cfg.cpuThreads = (uint8_t)std::min(callbackThreads, 255u);
	cfg.maxRaysInFlight = 128 * 128 * 128;
	cfg.maxRaysPerSpawn = 128 * 128; // one 128x128 tile of the example renderers
	cfg.cpuTestBatch = 1024;
	cfg.cpuShadeBatch = 8 * 1024;
	cfg.rayStreamBatchSize = 48 * 1024;
	return cfg;
}

racc::Context* racc::createContext(Configuration cfg) {
	if (!cfg.gpuContext) {
		fprintf(stderr, "RayAccelerator: no GPU context given; this engine has no CPU intersection path.\n");
		return nullptr;
	}
	if (!cfg.cpuThreads || !cfg.gpuSubmissionThreads || !cfg.maxRaysPerSpawn || !cfg.cpuShadeBatch || !cfg.rayStreamBatchSize ||
	    cfg.maxRaysInFlight < cfg.maxRaysPerSpawn) {
		fprintf(stderr, "RayAccelerator: invalid configuration.\n");
		return nullptr;
	}
	const std::vector<int> devices = devicesOf(cfg.gpuContext);
	const int available = racc_cuda_device_count();
	if (available <= 0 || devices.front() < 0 || devices.back() >= available) {
		fprintf(stderr, "RayAccelerator: CUDA device %d is not available (%s).\n", devices.front() < 0 ? devices.front() : devices.back(),
		        available < 0 ? racc_cuda_last_error() : "no such device");
		return nullptr;
	}
	if (racc_cuda_init(devices.data(), (int)devices.size())) {
		fprintf(stderr, "RayAccelerator: %s\n", racc_cuda_last_error());
		return nullptr;
	}

	// Pool sizing as the reference's clients expect it (RayAccelerator.cpp:516-521): a stream holds a
	// full batch plus whatever one more callback may add; enough streams that every thread can hold
	// two (one in, one out) on top of the rays-in-flight budget.
	const uint32_t capacity = (uint32_t)cfg.rayStreamBatchSize + std::max<uint32_t>(cfg.maxRaysPerSpawn, cfg.cpuShadeBatch);
	const uint32_t held = (uint32_t)cfg.gpuSubmissionThreads * (uint32_t)devices.size() + 2u * cfg.cpuThreads;
	const uint32_t count = held + (cfg.maxRaysInFlight + cfg.rayStreamBatchSize - 1u) / cfg.rayStreamBatchSize;
	if (count > 65535u) {
		fprintf(stderr, "RayAccelerator: configuration needs %u ray streams (limit 65535).\n", count);
		return nullptr;
	}
	const size_t rayBytes = align4k(sizeof(Ray) * (size_t)capacity);
	const size_t resultBytes = align4k(sizeof(Result) * (size_t)capacity);
	void* slab = racc_cuda_host_alloc((rayBytes + resultBytes) * count);
	if (!slab) {
		fprintf(stderr, "RayAccelerator: Unable to allocate memory (%s).\n", racc_cuda_last_error());
		return nullptr;
	}
	memset(slab, 0, (rayBytes + resultBytes) * count);

	Context* c = new Context();
	c->configuration = cfg;
	c->devices = devices;
	c->slab = slab;
	c->streamCapacity = capacity;
	c->streams.resize(count);
	char* p = static_cast<char*>(slab);
	for (uint32_t i = 0; i < count; ++i) {
		RayStream& s = c->streams[i];
		s.index = i;
		s.count = 0;
		s.rays = reinterpret_cast<Ray*>(p);
		p += rayBytes;
		s.results = reinterpret_cast<Result*>(p);
		p += resultBytes;
		c->idle.push_back(count - 1 - i); // stream 0 is handed out first
	}
	for (unsigned i = 0; i < cfg.cpuThreads; ++i)
		c->threads.emplace_back(callbackThread, c, i);
	for (int device : devices)
		for (unsigned i = 0; i < cfg.gpuSubmissionThreads; ++i)
			c->threads.emplace_back(submitterThread, c, device);
	return c;
}

void racc::destroy(Context* c) {
	if (!c) return;
	{
		std::lock_guard<std::mutex> lock(c->mutex);
		c->quit = true;
	}
	c->work.notify_all();
	for (std::thread& t : c->threads)
		t.join();
	racc_cuda_host_free(c->slab);
	delete c;
}

racc::ContextInfo racc::info(Context* c) {
	ContextInfo i = {};
	i.threadCount = c->configuration.cpuThreads;
	i.rayStreamCount = (uint16_t)c->streams.size();
	i.rayStreamSize = c->streamCapacity;
	i.maxRaysInFlight = c->configuration.maxRaysInFlight;
	return i;
}

racc::Scene* racc::createScene(Context* context, const Vertex* vertices, unsigned vertexCount, const uint32_t* indices, unsigned indexCount) {
	if (!context || !vertices || !indices || indexCount % 3 != 0 || (reinterpret_cast<uintptr_t>(vertices) & 15)) {
		fprintf(stderr, "RayAccelerator: Invalid scene input.\n"); // the reference asserts (Scene.cpp:186-187)
		return nullptr;
	}
	bindAll(context); // the scene is replicated on every device of the context
	racc_cuda_scene* handle = racc_cuda_scene_create(&vertices->x, vertexCount, indices, indexCount);
	if (!handle) {
		fprintf(stderr, "RayAccelerator: Unable to create scene (%s).\n", racc_cuda_last_error());
		return nullptr;
	}
	return new Scene{handle};
}

void racc::destroy(Scene* scene) {
	if (!scene) return;
	racc_cuda_scene_destroy(scene->handle);
	delete scene;
}

racc::Environment* racc::createEnvironment(Context* context, const Color* colors, unsigned width, unsigned height) {
	if (!context || !colors || !width || !height) {
		fprintf(stderr, "RayAccelerator: Invalid environment input.\n");
		return nullptr;
	}
	bindAll(context);
	racc_cuda_env* handle = racc_cuda_env_create(&colors->r, width, height);
	if (!handle) {
		fprintf(stderr, "RayAccelerator: Unable to create environment image (%s).\n", racc_cuda_last_error());
		return nullptr;
	}
	return new Environment{handle};
}

void racc::destroy(Environment* environment) {
	if (!environment) return;
	racc_cuda_env_destroy(environment->handle);
	delete environment;
}

racc::Stats racc::render(Context* c, Scene* scene, Environment* environment, RenderCallbacks callbacks) {
	std::unique_lock<std::mutex> lock(c->mutex);
	c->scene = scene;
	c->environment = environment;
	c->callbacks = callbacks;
	c->raysTraced = 0;
	c->frameOpen = true;
	c->work.notify_all();
	c->done.wait(lock, [c] { return frameFinished(c); });
	Stats stats = {};
	stats.raysTraced = c->raysTraced; // what the submitters handed to the device (RayAccelerator.cpp:372)
	lock.unlock();
	// The per-frame hit reduction: every device counted the rays its launches traced; their NCCL sum is the frame's total.
	// All launches have completed (a submitter files its streams as tested only after racc_cuda_sync). The frame records
	// are per device, not per context: when another context traced on the same devices meanwhile the sum is larger, and
	// the host count above stands.
	racc_cuda_counters totals = {};
	if (c->devices.size() > 1 && bindAll(c) == 0 && racc_cuda_frame_reduce(&totals, nullptr) == 0) {
		if (totals.rays != stats.raysTraced)
			fprintf(stderr, "RayAccelerator: devices counted %llu rays, submitters %llu (another context on the same devices?)\n",
			        (unsigned long long)totals.rays, (unsigned long long)stats.raysTraced);
		else
			stats.raysTraced = totals.rays;
	}
	return stats;
}
