// capi.cu -- implementation of the C-ABI declared in include/racc_b200.h: devices, scene / environment upload and
// replication, ray-stream staging and kernel launch. (The device-side renderers live in capi_render.cu, the per-frame
// hit reduction in comm.cu.) No CPU fallback: every compute entry point fails with an error string when CUDA is
// unavailable.
#include "capi_internal.h"

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>

using namespace racc_b200;

namespace racc_b200 {

namespace {

thread_local char t_error[512] = "";
thread_local std::vector<int> t_deviceSet; // CUDA ordinals, bound device first

std::mutex g_mutex; // device table, tuning
DeviceState g_devices[kMaxDevices];
Tuning g_tuning;
bool g_tuningFromEnv = false;
std::atomic<uint64_t> g_launches{0};

int envInt(const char* name, int fallback) {
	const char* v = getenv(name);
	return v && *v ? atoi(v) : fallback;
}

// caller holds g_mutex
void readTuningFromEnvironment() {
	if (g_tuningFromEnv) return;
	g_tuningFromEnv = true;
	Tuning& t = g_tuning;
	t.variant = envInt("RACC_B200_VARIANT", t.variant);
	t.blockThreads = envInt("RACC_B200_BLOCK", t.blockThreads);
	t.ctasPerSm = envInt("RACC_B200_CTAS_PER_SM", t.ctasPerSm);
	t.smemNodes = envInt("RACC_B200_SMEM_NODES", t.smemNodes);
	t.fetchThreshold = envInt("RACC_B200_FETCH_THRESHOLD", t.fetchThreshold);
	t.leafBail = envInt("RACC_B200_LEAF_BAIL", t.leafBail);
	t.innerBail = envInt("RACC_B200_INNER_BAIL", t.innerBail);
	t.carveout = envInt("RACC_B200_CARVEOUT", t.carveout);
	t.sortMode = envInt("RACC_B200_SORT", t.sortMode);
	t.sortOriginBits = envInt("RACC_B200_SORT_ORIGIN_BITS", t.sortOriginBits);
	t.sortDirBits = envInt("RACC_B200_SORT_DIR_BITS", t.sortDirBits);
	t.sortDirMajor = envInt("RACC_B200_SORT_DIR_MAJOR", t.sortDirMajor);
	t.buildDevice = envInt("RACC_B200_BUILD_DEVICE", t.buildDevice);
	t.smemStack = envInt("RACC_B200_SMEM_STACK", t.smemStack);
	t.hostZeroCopy = envInt("RACC_B200_HOST_ZERO_COPY", t.hostZeroCopy);
	t.hostTaper = envInt("RACC_B200_HOST_TAPER", t.hostTaper);
	t.whittedArena = envInt("RACC_B200_WHITTED_ARENA", t.whittedArena);
	t.whittedCombine = envInt("RACC_B200_WHITTED_COMBINE", t.whittedCombine);
	t.pathSync = envInt("RACC_B200_PATH_SYNC", t.pathSync);
	t.pathStream = envInt("RACC_B200_PATH_STREAM", t.pathStream);
	t.pathTraceCtas = envInt("RACC_B200_PATH_TRACE_CTAS", t.pathTraceCtas);
	t.pathStreamThreshold = envInt("RACC_B200_PATH_STREAM_THRESHOLD", t.pathStreamThreshold);
}

} // namespace

int fail(const char* fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(t_error, sizeof(t_error), fmt, ap);
	va_end(ap);
	return -1;
}

Tuning tuningSnapshot() {
	std::lock_guard<std::mutex> lock(g_mutex);
	readTuningFromEnvironment();
	return g_tuning;
}

void countLaunches(int launches) { g_launches.fetch_add((uint64_t)launches); }

DeviceState* useDevice(int ordinal) {
	if (ordinal < 0 || ordinal >= kMaxDevices) {
		fail("CUDA device %d out of range", ordinal);
		return nullptr;
	}
	cudaError_t e = cudaSetDevice(ordinal);
	if (e != cudaSuccess) {
		fail("cudaSetDevice(%d) failed: %s; the engine has no CPU fallback", ordinal, cudaGetErrorString(e));
		return nullptr;
	}
	DeviceState& d = g_devices[ordinal];
	std::lock_guard<std::mutex> lock(g_mutex);
	if (d.ready) return &d;
	readTuningFromEnvironment();
	d.ordinal = ordinal;
	if ((e = cudaDeviceGetAttribute(&d.smCount, cudaDevAttrMultiProcessorCount, ordinal)) != cudaSuccess) {
		fail("cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
		return nullptr;
	}
	{
		size_t freeBytes = 0;
		if (cudaMemGetInfo(&freeBytes, &d.totalBytes) != cudaSuccess) d.totalBytes = (size_t)16 << 30;
	}
	{
		// stream-ordered scratch (stream tables, re-binning buffers, renderer waves) is recycled instead of being handed
		// back to the driver at every synchronisation
		cudaMemPool_t pool = nullptr;
		if (cudaDeviceGetDefaultMemPool(&pool, ordinal) == cudaSuccess && pool) {
			unsigned long long keep = ~0ull;
			cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
		}
	}
	if ((e = cudaMalloc(reinterpret_cast<void**>(&d.dFrame), 16 * sizeof(unsigned long long))) != cudaSuccess ||
	    (e = cudaMemset(d.dFrame, 0, 16 * sizeof(unsigned long long))) != cudaSuccess ||
	    (e = cudaStreamCreateWithFlags(&d.reduceStream, cudaStreamNonBlocking)) != cudaSuccess ||
	    (e = cudaEventCreateWithFlags(&d.reduceReady, cudaEventDisableTiming)) != cudaSuccess ||
	    (e = cudaEventCreateWithFlags(&d.reduceDone, cudaEventDisableTiming)) != cudaSuccess) {
		fail("device %d: frame counters: %s", ordinal, cudaGetErrorString(e));
		return nullptr;
	}
	d.dFrameTotal = d.dFrame + 8;
	d.ready = true;
	return &d;
}

DeviceState* currentDevice() {
	if (t_deviceSet.empty()) {
		int count = 0;
		cudaError_t e = cudaGetDeviceCount(&count);
		if (e != cudaSuccess || count <= 0) {
			fail("no CUDA device available (%s); the engine has no CPU fallback", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
			return nullptr;
		}
		int ordinal = 0;
		if ((e = cudaGetDevice(&ordinal)) != cudaSuccess) {
			fail("cudaGetDevice failed: %s", cudaGetErrorString(e));
			return nullptr;
		}
		DeviceState* d = useDevice(ordinal);
		if (d) t_deviceSet.assign(1, ordinal);
		return d;
	}
	return useDevice(t_deviceSet[0]);
}

const std::vector<int>& currentDeviceSet() { return t_deviceSet; }

namespace {

void fillInfo(const SceneImages& h, uint32_t triangleCount, racc_cuda_scene_info* info) {
	info->node_count = (uint32_t)h.nodes.size();
	info->pair_count = (uint32_t)h.pairs.size();
	info->real_pair_count = h.realPairs;
	info->remap_count = (uint32_t)h.remap.size();
	info->depth = h.depth;
	info->triangle_count = triangleCount;
	for (int k = 0; k < 3; ++k) {
		info->bounds_min[k] = h.boundsMin[k];
		info->bounds_max[k] = h.boundsMax[k];
	}
}

// Every traversal kernel keeps the reference's 64-entry stack (Kernels.h:166). A ray pushes at most one far child per inner
// level, so a tree of `depth` levels (leaves included) needs depth-1 entries; the reference overruns its private array
// silently on a deeper tree, here such a scene is refused when it is created.
constexpr uint32_t kTraversalStackEntries = 64;

bool depthFitsStack(uint32_t depth) {
	if (depth <= kTraversalStackEntries + 1) return true;
	fail("scene tree is %u levels deep; the traversal stack holds %u entries (Kernels.h:166), i.e. trees up to %u levels", depth,
	     kTraversalStackEntries, kTraversalStackEntries + 1);
	return false;
}

// Levels of a node image (leaves included), or 0 when the image is not a tree (a reference cycle or a shared node).
uint32_t imageDepth(const std::vector<GpuNode>& nodes) {
	std::vector<std::pair<uint32_t, uint32_t>> todo;
	todo.emplace_back(0u, 1u);
	size_t visited = 0;
	uint32_t depth = 0;
	while (!todo.empty()) {
		const auto [index, level] = todo.back();
		todo.pop_back();
		if (++visited > nodes.size()) return 0;
		const uint32_t refs[2] = {nodes[index].first, nodes[index].last};
		for (uint32_t r : refs) {
			if (r & 0x80000000u) todo.emplace_back(r & 0x7fffffffu, level + 1);
			else if (level + 1 > depth) depth = level + 1;
		}
	}
	return depth;
}

void freeReplica(SceneReplica& r) {
	if (r.device >= 0) cudaSetDevice(r.device);
	cudaFree(r.dNodes);
	cudaFree(r.dPairs);
	cudaFree(r.dRemap);
	cudaFree(r.dTNodes);
	cudaFree(r.dTPairs);
	cudaFree(r.dQNodes);
	cudaFree(r.dVerts);
	cudaFree(r.dIndices);
	cudaFree(r.dCursors);
	cudaFree(r.dBounceScratch);
}

// device-private packed copies of the node and pair images, derived on the device the replica lives on (current)
bool packReplica(racc_cuda_scene* s, SceneReplica* r) {
	cudaError_t e;
	int launches = 0;
	if ((e = cudaMalloc(reinterpret_cast<void**>(&r->dCursors), kCursorRing * sizeof(uint32_t))) != cudaSuccess ||
	    (e = cudaMalloc(reinterpret_cast<void**>(&r->dTNodes), (size_t)s->info.node_count * 64 + 64)) != cudaSuccess ||
	    (e = cudaMalloc(reinterpret_cast<void**>(&r->dTPairs), (size_t)s->info.pair_count * 64 + 64)) != cudaSuccess ||
	    (e = launchPackImages(r->dNodes, s->info.node_count, r->dPairs, s->info.pair_count, r->dTNodes, r->dTPairs, nullptr, &launches)) != cudaSuccess ||
	    (e = cudaMalloc(&r->dQNodes, (size_t)s->info.node_count * 32 + 32)) != cudaSuccess ||
	    (e = launchQuantiseNodes(r->dNodes, s->info.node_count, s->info.bounds_min, s->info.bounds_max, r->dQNodes, s->qOrigin, s->qCell, nullptr,
	                             &launches)) != cudaSuccess ||
	    (e = cudaDeviceSynchronize()) != cudaSuccess) {
		fail("scene packing failed: %s", cudaGetErrorString(e));
		return false;
	}
	countLaunches(launches);
	return true;
}

// dst (on device `to`, allocated here) = src (on device `from`)
template <typename T>
bool cloneBuffer(T** dst, int to, const T* src, int from, size_t bytes) {
	*dst = nullptr;
	if (!src) return true;
	cudaError_t e;
	if ((e = cudaSetDevice(to)) != cudaSuccess || (e = cudaMalloc(reinterpret_cast<void**>(dst), bytes ? bytes : 16)) != cudaSuccess ||
	    (bytes && (e = cudaMemcpyPeer(*dst, to, src, from, bytes)) != cudaSuccess)) {
		fail("replicating to CUDA device %d failed: %s", to, cudaGetErrorString(e));
		return false;
	}
	return true;
}

// The scene exists on the first device of the calling thread's set; copy it to the others (peer copies over NVLink: the
// build, the pair merge and the packing run once).
racc_cuda_scene* finishScene(racc_cuda_scene* s) {
	const std::vector<int> set = currentDeviceSet();
	SceneReplica* first = s->replicas[0].get();
	const racc_cuda_scene_info& in = s->info;
	for (size_t k = 1; k < set.size(); ++k) {
		if (!useDevice(set[k])) { racc_cuda_scene_destroy(s); return nullptr; }
		std::unique_ptr<SceneReplica> r(new SceneReplica());
		r->device = set[k];
		const bool ok = cloneBuffer(&r->dNodes, r->device, first->dNodes, first->device, (size_t)in.node_count * 64) &&
		                cloneBuffer(&r->dPairs, r->device, first->dPairs, first->device, (size_t)in.pair_count * 48) &&
		                cloneBuffer(&r->dRemap, r->device, first->dRemap, first->device, (size_t)in.remap_count * 4) &&
		                cloneBuffer(&r->dTNodes, r->device, first->dTNodes, first->device, (size_t)in.node_count * 64 + 64) &&
		                cloneBuffer(&r->dTPairs, r->device, first->dTPairs, first->device, (size_t)in.pair_count * 64 + 64) &&
		                cloneBuffer(reinterpret_cast<char**>(&r->dQNodes), r->device, static_cast<const char*>(first->dQNodes), first->device, (size_t)in.node_count * 32 + 32) &&
		                cloneBuffer(&r->dVerts, r->device, first->dVerts, first->device, (size_t)s->vertexCount * 16) &&
		                cloneBuffer(&r->dIndices, r->device, first->dIndices, first->device, (size_t)s->triangleCount * 12);
		cudaError_t e = cudaSuccess;
		if (ok) e = cudaMalloc(reinterpret_cast<void**>(&r->dCursors), kCursorRing * sizeof(uint32_t));
		s->replicas.push_back(std::move(r));
		if (!ok || e != cudaSuccess) {
			if (ok) fail("replicating to CUDA device %d failed: %s", set[k], cudaGetErrorString(e));
			racc_cuda_scene_destroy(s);
			return nullptr;
		}
	}
	cudaSetDevice(set[0]);
	return s;
}

racc_cuda_scene* uploadScene(racc_cuda_scene* s, const float* verts4, uint32_t nverts, const uint32_t* indices, uint32_t nindices) {
	const SceneImages& h = s->host;
	DeviceState* dev = currentDevice();
	if (!dev) { delete s; return nullptr; }
	auto bail = [&]() -> racc_cuda_scene* { racc_cuda_scene_destroy(s); return nullptr; };
	cudaError_t e;
	fillInfo(h, s->triangleCount, &s->info);
	if (!depthFitsStack(s->info.depth)) return bail();
	s->replicas.emplace_back(new SceneReplica());
	SceneReplica* r = s->replicas[0].get();
	r->device = dev->ordinal;
#define UP(dst, src, bytes)                                                                  \
	if ((e = cudaMalloc(reinterpret_cast<void**>(&dst), (bytes) != 0 ? (bytes) : 16)) != cudaSuccess || \
	    (e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice)) != cudaSuccess) {            \
		fail("scene upload failed: %s", cudaGetErrorString(e));                               \
		return bail();                                                                        \
	}
	UP(r->dNodes, h.nodes.data(), h.nodes.size() * sizeof(GpuNode))
	UP(r->dPairs, h.pairs.data(), h.pairs.size() * sizeof(GpuPair))
	UP(r->dRemap, h.remap.data(), h.remap.size() * sizeof(uint32_t))
	if (verts4 && indices) {
		UP(r->dVerts, verts4, (size_t)nverts * 16)
		UP(r->dIndices, indices, (size_t)nindices * 4)
	}
#undef UP
	if (!packReplica(s, r)) return bail();
	return finishScene(s);
}

} // namespace

bool sceneExceedsL2(const racc_cuda_scene* s) {
	return ((size_t)s->info.node_count + s->info.pair_count) * 64 > kAutoSortSceneBytes;
}

} // namespace racc_b200

extern "C" {

int racc_cuda_abi_version(void) { return RACC_CUDA_ABI_VERSION; }

const char* racc_cuda_last_error(void) { return t_error; }

int racc_cuda_device_count(void) {
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess) {
		fail("cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
		return -1;
	}
	return count;
}

int racc_cuda_init(const int* devices, int n) {
	if (!devices || n <= 0) {
		t_deviceSet.clear();
		return currentDevice() ? 0 : -1;
	}
	std::vector<int> set;
	for (int k = 0; k < n; ++k) {
		for (int seen : set)
			if (seen == devices[k]) return fail("racc_cuda_init: CUDA device %d named twice", devices[k]);
		if (!useDevice(devices[k])) return -1;
		set.push_back(devices[k]);
	}
	t_deviceSet = set;
	return currentDevice() ? 0 : -1;
}

int racc_cuda_current_devices(int* devices, int capacity) {
	if (!currentDevice()) return -1;
	const std::vector<int>& set = currentDeviceSet();
	for (int k = 0; k < (int)set.size() && k < capacity; ++k)
		if (devices) devices[k] = set[k];
	return (int)set.size();
}

int racc_cuda_set_variant(int variant) { return racc_cuda_set_tuning(0, variant); }

// Extended tuning access for the benchmark sweeps (keys in include/racc_b200.h). Returns the previous value.
int racc_cuda_set_tuning(int key, int value) {
	std::lock_guard<std::mutex> lock(g_mutex);
	readTuningFromEnvironment();
	int* slot = nullptr;
	switch (key) {
	case 0: slot = &g_tuning.variant; break;
	case 1: slot = &g_tuning.blockThreads; break;
	case 2: slot = &g_tuning.ctasPerSm; break;
	case 3: slot = &g_tuning.smemNodes; break;
	case 4: slot = &g_tuning.fetchThreshold; break;
	case 5: slot = &g_tuning.leafBail; break;
	case 7: slot = &g_tuning.innerBail; break;
	case 6: slot = &g_tuning.carveout; break;
	case 8: slot = &g_tuning.sortMode; break;
	case 9: slot = &g_tuning.sortOriginBits; break;
	case 10: slot = &g_tuning.sortDirBits; break;
	case 11: slot = &g_tuning.sortDirMajor; break;
	case 12: slot = &g_tuning.buildDevice; break;
	case 13: slot = &g_tuning.smemStack; break;
	case 14: slot = &g_tuning.hostZeroCopy; break;
	case 15: slot = &g_tuning.whittedArena; break;
	case 16: slot = &g_tuning.whittedCombine; break;
	case 17: slot = &g_tuning.hostTaper; break;
	case 18: slot = &g_tuning.pathSync; break;
	case 19: slot = &g_tuning.pathStream; break;
	case 20: slot = &g_tuning.pathTraceCtas; break;
	default: return fail("unknown tuning key %d", key);
	}
	const int previous = *slot;
	*slot = value;
	return previous;
}

uint64_t racc_cuda_launch_count(void) { return g_launches.load(); }

int racc_cuda_debug_rcp_table(float* out2048) {
	if (!out2048) return fail("racc_cuda_debug_rcp_table: null argument");
	return fillRcpTable(out2048) ? 0 : 1;
}

int racc_cuda_debug_warp_stats(uint64_t* out8, int reset) {
	if (!out8) return fail("racc_cuda_debug_warp_stats: null argument");
	if (!currentDevice()) return -1;
	RACC_CUDA_CHECK(cudaDeviceSynchronize());
	RACC_CUDA_CHECK(readWarpStats(reinterpret_cast<unsigned long long*>(out8), reset != 0));
	return 0;
}

racc_cuda_scene* racc_cuda_scene_create(const float* verts4, uint32_t nverts, const uint32_t* indices, uint32_t nindices) {
	if (!verts4 || !indices) { fail("racc_cuda_scene_create: null input"); return nullptr; }
	DeviceState* dev = currentDevice();
	if (!dev) return nullptr;
	const Tuning tuning = tuningSnapshot();
	racc_cuda_scene* s = new racc_cuda_scene();
	const char* why = "";
	bool built = false;
	s->triangleCount = nindices / 3;
	s->vertexCount = nverts;
	if (tuning.buildDevice == 2 || (tuning.buildDevice == 3 && nindices / 3 >= kAutoDeviceBuildTriangles)) {
		// the whole build on the device: the images never exist on the host
		DeviceSceneImages img;
		if (buildSceneImagesDevice(verts4, nverts, indices, nindices, &img, &why)) {
			s->replicas.emplace_back(new SceneReplica());
			SceneReplica* r = s->replicas[0].get();
			r->device = dev->ordinal;
			r->dNodes = static_cast<float4*>(img.nodes);
			r->dPairs = static_cast<float4*>(img.pairs);
			r->dRemap = img.remap;
			r->dVerts = static_cast<float4*>(img.verts);
			r->dIndices = img.indices;
			s->info.node_count = img.nodeCount;
			s->info.pair_count = img.pairCount;
			s->info.real_pair_count = img.realPairs;
			s->info.remap_count = img.remapCount;
			s->info.depth = img.depth;
			s->info.triangle_count = s->triangleCount;
			for (int k = 0; k < 3; ++k) { s->info.bounds_min[k] = img.boundsMin[k]; s->info.bounds_max[k] = img.boundsMax[k]; }
			if (!depthFitsStack(s->info.depth) || !packReplica(s, r)) {
				racc_cuda_scene_destroy(s);
				return nullptr;
			}
			return finishScene(s);
		}
		if (tuning.buildDevice == 2) fprintf(stderr, "RayAccelerator: device scene build declined (%s); building on the host\n", why);
	}
	else if (tuning.buildDevice == 1) {
		built = buildSceneImages(verts4, nverts, indices, nindices, 0, &s->host, &why, buildBvh2Device);
		if (!built) fprintf(stderr, "RayAccelerator: device SAH build declined (%s); building on the host\n", why);
	}
	if (!built && !buildSceneImages(verts4, nverts, indices, nindices, envInt("RACC_B200_BUILD_THREADS", 0), &s->host, &why)) {
		fail("racc_cuda_scene_create: %s", why);
		delete s;
		return nullptr;
	}
	return uploadScene(s, verts4, nverts, indices, nindices);
}

racc_cuda_host_images* racc_cuda_build_images(const float* verts4, uint32_t nverts, const uint32_t* indices, uint32_t nindices) {
	if (!verts4 || !indices) { fail("racc_cuda_build_images: null input"); return nullptr; }
	racc_cuda_host_images* img = new racc_cuda_host_images();
	const char* why = "";
	if (!buildSceneImages(verts4, nverts, indices, nindices, envInt("RACC_B200_BUILD_THREADS", 0), &img->images, &why)) {
		fail("racc_cuda_build_images: %s", why);
		delete img;
		return nullptr;
	}
	img->triangleCount = nindices / 3;
	return img;
}

int racc_cuda_host_images_get_info(const racc_cuda_host_images* img, racc_cuda_scene_info* info) {
	if (!img || !info) return fail("racc_cuda_host_images_get_info: null argument");
	fillInfo(img->images, img->triangleCount, info);
	return 0;
}

int racc_cuda_host_images_copy(const racc_cuda_host_images* img, void* nodes, void* pairs, uint32_t* remap) {
	if (!img) return fail("racc_cuda_host_images_copy: null argument");
	if (nodes) memcpy(nodes, img->images.nodes.data(), img->images.nodes.size() * sizeof(GpuNode));
	if (pairs) memcpy(pairs, img->images.pairs.data(), img->images.pairs.size() * sizeof(GpuPair));
	if (remap) memcpy(remap, img->images.remap.data(), img->images.remap.size() * sizeof(uint32_t));
	return 0;
}

void racc_cuda_host_images_destroy(racc_cuda_host_images* img) { delete img; }

racc_cuda_scene* racc_cuda_scene_create_from_images(const void* nodes, uint32_t node_count, const void* pairs,
                                                    uint32_t pair_count, const uint32_t* remap, uint32_t remap_count) {
	if (!nodes || !pairs || !remap || !node_count || !pair_count) { fail("racc_cuda_scene_create_from_images: null or empty image"); return nullptr; }
	if (!currentDevice()) return nullptr;
	racc_cuda_scene* s = new racc_cuda_scene();
	s->host.nodes.assign(static_cast<const GpuNode*>(nodes), static_cast<const GpuNode*>(nodes) + node_count);
	s->host.pairs.assign(static_cast<const GpuPair*>(pairs), static_cast<const GpuPair*>(pairs) + pair_count);
	s->host.remap.assign(remap, remap + remap_count);
	s->host.realPairs = remap_count / 2;
	{
		// scene bounds = union of the root's two child boxes (used only by the ray re-binning keys); a synthetic root's
		// second box sits at +infinity (scene_build.cpp) and is left out
		const GpuNode& root = s->host.nodes[0];
		for (int k = 0; k < 3; ++k) {
			const bool both = root.rightMin[k] <= root.rightMax[k] && root.rightMin[k] < 3.0e38f;
			s->host.boundsMin[k] = both && root.rightMin[k] < root.leftMin[k] ? root.rightMin[k] : root.leftMin[k];
			s->host.boundsMax[k] = both && root.rightMax[k] > root.leftMax[k] ? root.rightMax[k] : root.leftMax[k];
		}
	}
	// validate references so a malformed image cannot send the kernel out of bounds
	for (const GpuNode& n : s->host.nodes) {
		const uint32_t refs[2] = {n.first, n.last};
		for (uint32_t r : refs) {
			const bool ok = (r & 0x80000000u) ? (r & 0x7fffffffu) < node_count
			                                  : ((r >> 24) > 0 && (r & 0xffffffu) + (r >> 24) <= pair_count && 2 * ((r & 0xffffffu) + (r >> 24)) <= remap_count);
			if (!ok) {
				fail("racc_cuda_scene_create_from_images: child reference 0x%08x out of range", r);
				delete s;
				return nullptr;
			}
		}
	}
	s->host.depth = imageDepth(s->host.nodes);
	if (!s->host.depth) {
		fail("racc_cuda_scene_create_from_images: the node image is not a tree (a node is referenced twice or lies on a cycle)");
		delete s;
		return nullptr;
	}
	return uploadScene(s, nullptr, 0, nullptr, 0);
}

void racc_cuda_scene_destroy(racc_cuda_scene* s) {
	if (!s) return;
	int before = -1;
	cudaGetDevice(&before);
	for (auto& r : s->replicas) freeReplica(*r);
	if (before >= 0) cudaSetDevice(before);
	delete s;
}

int racc_cuda_scene_get_info(const racc_cuda_scene* s, racc_cuda_scene_info* info) {
	if (!s || !info) return fail("racc_cuda_scene_get_info: null argument");
	*info = s->info;
	return 0;
}

int racc_cuda_scene_download(const racc_cuda_scene* s, void* nodes, void* pairs, uint32_t* remap) {
	if (!s) return fail("racc_cuda_scene_download: null scene");
	DeviceState* dev = currentDevice();
	if (!dev) return -1;
	// read back from the device so the test sees what the kernel sees
	const SceneReplica* r = s->on(dev->ordinal);
	if (!r) {
		r = s->replicas[0].get();
		RACC_CUDA_CHECK(cudaSetDevice(r->device));
	}
	if (nodes) RACC_CUDA_CHECK(cudaMemcpy(nodes, r->dNodes, (size_t)s->info.node_count * sizeof(GpuNode), cudaMemcpyDeviceToHost));
	if (pairs) RACC_CUDA_CHECK(cudaMemcpy(pairs, r->dPairs, (size_t)s->info.pair_count * sizeof(GpuPair), cudaMemcpyDeviceToHost));
	if (remap) RACC_CUDA_CHECK(cudaMemcpy(remap, r->dRemap, (size_t)s->info.remap_count * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	RACC_CUDA_CHECK(cudaSetDevice(dev->ordinal));
	return 0;
}

racc_cuda_env* racc_cuda_env_create(const float* rgba, uint32_t width, uint32_t height) {
	if (!rgba || !width || !height) { fail("racc_cuda_env_create: null or empty image"); return nullptr; }
	if (!currentDevice()) return nullptr;
	const std::vector<int> set = currentDeviceSet();
	racc_cuda_env* env = new racc_cuda_env();
	env->width = width;
	env->height = height;
	const size_t bytes = (size_t)width * height * 16;
	for (int ordinal : set) {
		if (!useDevice(ordinal)) { racc_cuda_env_destroy(env); return nullptr; }
		env->replicas.emplace_back(new EnvReplica());
		EnvReplica* r = env->replicas.back().get();
		r->device = ordinal;
		cudaError_t e;
		int launches = 0;
		if ((e = cudaMalloc(reinterpret_cast<void**>(&r->dTexels), bytes)) != cudaSuccess ||
		    (e = cudaMemcpy(r->dTexels, rgba, bytes, cudaMemcpyHostToDevice)) != cudaSuccess ||
		    (e = cudaMalloc(reinterpret_cast<void**>(&r->dTexelPairs), ((size_t)width + 1) * height * 32)) != cudaSuccess ||
		    (e = launchPackEnv(r->dTexels, width, height, r->dTexelPairs, nullptr, &launches)) != cudaSuccess ||
		    (e = cudaDeviceSynchronize()) != cudaSuccess) {
			fail("environment upload failed: %s", cudaGetErrorString(e));
			racc_cuda_env_destroy(env);
			cudaSetDevice(set[0]);
			return nullptr;
		}
		countLaunches(launches);
	}
	cudaSetDevice(set[0]);
	return env;
}

void racc_cuda_env_destroy(racc_cuda_env* env) {
	if (!env) return;
	int before = -1;
	cudaGetDevice(&before);
	for (auto& r : env->replicas) {
		cudaSetDevice(r->device);
		cudaFree(r->dTexels);
		cudaFree(r->dTexelPairs);
	}
	if (before >= 0) cudaSetDevice(before);
	delete env;
}

} // extern "C"

namespace racc_b200 {
namespace {

// Staging pipeline for HOST ray streams: chunks alternate over a few internal CUDA streams so that
// the H2D copy of chunk k+1, the traversal of chunk k and the D2H copy of chunk k-1 overlap (PCIe
// is full duplex). One pipeline per calling host thread and device, so concurrent submitters never share
// staging buffers. Replaces the reference's zero-copy CL_MEM_USE_HOST_PTR streams
// (RayAccelerator.cpp:643-644), which a discrete GPU does not have. Released by racc_cuda_thread_release
// (racc_api.cpp calls it when a submitter thread ends).
struct HostPipeline {
	static constexpr int kLanes = 3;
	bool ready = false;
	int device = -1;
	uint32_t chunkRays = 0;
	cudaStream_t lane[kLanes] = {};
	cudaEvent_t done[kLanes] = {};
	cudaEvent_t fork = nullptr;
	DevRay* dRays[kLanes] = {};
	float4* dResults[kLanes] = {};

	// the device is current
	int init(int ordinal) {
		if (ready) return 0;
		release(); // whatever an earlier, failed init left behind
		device = ordinal;
		chunkRays = (uint32_t)envInt("RACC_B200_HOST_CHUNK", 1 << 20); // 32 MB of rays per H2D copy: best of 256K..4M (profiles/r01_e2e_chunk_sweep.txt)
		if (chunkRays < 1024) chunkRays = 1024;
		RACC_CUDA_CHECK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
		for (int l = 0; l < kLanes; ++l) {
			RACC_CUDA_CHECK(cudaStreamCreateWithFlags(&lane[l], cudaStreamNonBlocking));
			RACC_CUDA_CHECK(cudaEventCreateWithFlags(&done[l], cudaEventDisableTiming));
			RACC_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&dRays[l]), (size_t)chunkRays * sizeof(DevRay)));
			RACC_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&dResults[l]), (size_t)chunkRays * sizeof(float4)));
		}
		ready = true;
		return 0;
	}

	// frees whatever exists, also after an init() that failed part-way
	void release() {
		if (device >= 0) cudaSetDevice(device);
		for (int l = 0; l < kLanes; ++l) {
			if (lane[l]) { cudaStreamSynchronize(lane[l]); cudaStreamDestroy(lane[l]); }
			if (done[l]) cudaEventDestroy(done[l]);
			if (dRays[l]) cudaFree(dRays[l]);
			if (dResults[l]) cudaFree(dResults[l]);
			lane[l] = nullptr; done[l] = nullptr; dRays[l] = nullptr; dResults[l] = nullptr;
		}
		if (fork) cudaEventDestroy(fork);
		fork = nullptr;
		ready = false;
	}

	~HostPipeline() { /* thread exit without racc_cuda_thread_release, or process teardown (the context may be gone): leak on purpose */ }
};

thread_local std::vector<std::unique_ptr<HostPipeline>> t_pipelines;

HostPipeline* pipelineOn(int ordinal) {
	for (auto& p : t_pipelines)
		if (p->device == ordinal) return p.get();
	t_pipelines.emplace_back(new HostPipeline());
	t_pipelines.back()->device = ordinal;
	return t_pipelines.back().get();
}

void fillSceneParams(TraceParams& p, const racc_cuda_scene* s, const SceneReplica* r, const racc_cuda_env* env, const EnvReplica* er,
                     unsigned long long* counters) {
	p.nodes = r->dNodes;
	p.pairs = r->dPairs;
	p.remap = r->dRemap;
	p.env = er ? er->dTexels : nullptr;
	p.envWidth = env ? env->width : 0;
	p.envHeight = env ? env->height : 0;
	p.nodeCount = s->info.node_count;
	p.counters = counters;
	p.tnodes = r->dTNodes;
	p.tpairs = r->dTPairs;
	p.perm = nullptr;
	p.envPairs = er ? er->dTexelPairs : nullptr;
	p.totalPtr = nullptr;
	p.qnodes = r->dQNodes;
	for (int a = 0; a < 3; ++a) { p.qOrigin[a] = s->qOrigin[a]; p.qCell[a] = s->qCell[a]; }
}

cudaError_t launchAny(const racc_cuda_scene* s, const Tuning& tuning, const DeviceState* dev, const TraceParams& p, int counterMode,
                      cudaStream_t stream, int* launches) {
	if (tuning.variant != 3 && tuning.variant != 4)
		return launchTrace(p, tuning, counterMode, dev->smCount, stream, launches);
	Tuning t = tuning;
	if (t.smemStack < 0) t.smemStack = sceneExceedsL2(s) ? 16 : 0;
	return launchTracePacked(p, t, counterMode, dev->smCount, stream, launches);
}

} // namespace

int traceImpl(racc_cuda_scene* s, racc_cuda_env* env, const racc_cuda_stream_desc* streams, uint32_t nstreams, void* cuda_stream,
              void* device_counters, bool fullCounters, const uint32_t* deviceTotal, int gridCtasPerSm) {
	if (!s) return fail("racc_cuda_trace: null scene");
	if (!streams && nstreams) return fail("racc_cuda_trace: null stream list");
	DeviceState* dev = currentDevice();
	if (!dev) return -1;
	Tuning tuning = tuningSnapshot();
	tuning.gridCtasPerSm = gridCtasPerSm;
	cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
	SceneReplica* rep = s->on(dev->ordinal);
	if (!rep) return fail("racc_cuda_trace: the scene has no copy on CUDA device %d (it was created for another device set)", dev->ordinal);
	EnvReplica* erep = env ? env->on(dev->ordinal) : nullptr;
	if (env && !erep) return fail("racc_cuda_trace: the environment has no copy on CUDA device %d", dev->ordinal);
	// Without a caller's record every launch adds rays + hits to its device's frame record (racc_cuda_frame_reduce)
	const int counterMode = device_counters ? (fullCounters ? 2 : 1) : 1;

	std::vector<StreamRef> refs; // device-resident streams: one launch for all of them
	std::vector<const racc_cuda_stream_desc*> hostStreams;
	uint64_t total = 0;
	bool zeroCopy = false; // some stream of the launch lives in host memory: do not re-bin (random gathers over PCIe)
	for (uint32_t i = 0; i < nstreams; ++i) {
		const racc_cuda_stream_desc& d = streams[i];
		if (!d.count) continue;
		if (!d.rays || !d.results) return fail("racc_cuda_trace: stream %u has null buffers", i);
		const void* rays = d.rays;
		void* results = d.results;
		if (d.flags & RACC_CUDA_STREAM_HOST) {
			// Pinned, mapped host memory (racc_cuda_host_alloc, cudaHostAlloc, torch pin_memory) can be read and
			// written by the kernel itself over PCIe: no staging copies, one launch, H2D and D2H traffic interleaved
			// ray by ray. Anything else goes through the staging pipeline below.
			bool direct = false;
			if (tuning.hostZeroCopy && !((reinterpret_cast<uintptr_t>(d.rays) | reinterpret_cast<uintptr_t>(d.results)) & 15)) {
				cudaPointerAttributes ar{}, ao{};
				if (cudaPointerGetAttributes(&ar, d.rays) == cudaSuccess && cudaPointerGetAttributes(&ao, d.results) == cudaSuccess &&
				    ar.type == cudaMemoryTypeHost && ao.type == cudaMemoryTypeHost && ar.devicePointer && ao.devicePointer) {
					rays = ar.devicePointer;
					results = ao.devicePointer;
					direct = true;
					zeroCopy = true;
				}
				else {
					cudaGetLastError(); // unregistered host memory is reported as an error by older runtimes
				}
			}
			if (!direct) {
				hostStreams.push_back(&d);
				continue;
			}
		}
		if ((reinterpret_cast<uintptr_t>(rays) & 15) || (reinterpret_cast<uintptr_t>(results) & 15))
			return fail("racc_cuda_trace: stream %u buffers must be 16-byte aligned", i);
		StreamRef ref;
		ref.begin = (uint32_t)total;
		ref.count = d.count;
		ref.rays = static_cast<const DevRay*>(rays);
		ref.results = static_cast<float4*>(results);
		refs.push_back(ref);
		total += d.count;
		if (total > 0x7fffffffull) return fail("racc_cuda_trace: more than 2^31-1 rays in one launch");
	}
	if (deviceTotal && (refs.size() != 1 || !hostStreams.empty())) return fail("racc_cuda_trace: a device-side ray count needs exactly one DEVICE stream");

	int launches = 0;
	if (total) {
		TraceParams p{};
		fillSceneParams(p, s, rep, env, erep, device_counters ? static_cast<unsigned long long*>(device_counters) : dev->dFrame);
		p.nstreams = (uint32_t)refs.size();
		p.total = (uint32_t)total;
		p.totalPtr = deviceTotal;
		p.single = refs[0];
		p.cursor = rep->dCursors + (rep->nextCursor.fetch_add(1) % kCursorRing);
		void* dRefs = nullptr;
		if (refs.size() > 1) {
			RACC_CUDA_CHECK(cudaMallocAsync(&dRefs, refs.size() * sizeof(StreamRef), stream));
			RACC_CUDA_CHECK(cudaMemcpyAsync(dRefs, refs.data(), refs.size() * sizeof(StreamRef), cudaMemcpyHostToDevice, stream));
			p.streams = static_cast<const StreamRef*>(dRefs);
		}
		void* sortScratch = nullptr;
		const bool rebin = (tuning.variant == 3 || tuning.variant == 4) && total >= 4096 && !zeroCopy && !deviceTotal &&
		                   (tuning.sortMode == 1 || (tuning.sortMode == 2 && sceneExceedsL2(s) && total >= (1u << 18)));
		if (rebin) {
			// re-bin the launch: visiting order by origin/direction key, results stay index-parallel
			// an optimisation only: when the scratch does not fit, trace in arrival order
			if (cudaMallocAsync(&sortScratch, raySortScratchBytes(p.total), stream) != cudaSuccess) {
				cudaGetLastError();
				sortScratch = nullptr;
			}
			else {
				RACC_CUDA_CHECK(launchRaySort(p, s->info.bounds_min, s->info.bounds_max, tuning.sortOriginBits, tuning.sortDirBits,
				                              tuning.sortDirMajor, sortScratch, dev->smCount, stream, &p.perm, &launches));
			}
		}
		RACC_CUDA_CHECK(launchAny(s, tuning, dev, p, counterMode, stream, &launches));
		if (sortScratch) RACC_CUDA_CHECK(cudaFreeAsync(sortScratch, stream));
		if (dRefs) RACC_CUDA_CHECK(cudaFreeAsync(dRefs, stream));
	}

	if (!hostStreams.empty()) {
		// Devices of this call: the calling thread's whole set, on which the scene (and environment) have a copy. Chunks
		// are dealt round-robin over them -- the analogue of the reference's gpuSubmissionThreads sharing one device's
		// queue (RayAccelerator.cpp:335-414), for a box where every GPU hangs off its own PCIe link.
		struct Target { DeviceState* dev; SceneReplica* rep; EnvReplica* erep; HostPipeline* pipe; unsigned chunks; };
		std::vector<Target> targets;
		for (int ordinal : currentDeviceSet()) {
			SceneReplica* r = s->on(ordinal);
			EnvReplica* er = env ? env->on(ordinal) : nullptr;
			if (!r || (env && !er)) continue;
			DeviceState* d = useDevice(ordinal);
			if (!d) return -1;
			HostPipeline* pipe = pipelineOn(ordinal);
			if (pipe->init(ordinal)) return -1;
			targets.push_back(Target{d, r, er, pipe, 0u});
		}
		RACC_CUDA_CHECK(cudaSetDevice(dev->ordinal));
		HostPipeline& first = *targets[0].pipe;
		RACC_CUDA_CHECK(cudaEventRecord(first.fork, stream));
		for (Target& t : targets)
			for (int l = 0; l < HostPipeline::kLanes; ++l)
				RACC_CUDA_CHECK(cudaStreamWaitEvent(t.pipe->lane[l], first.fork, 0));
		// Pack host streams into staging chunks: a chunk takes whole streams or slices of them until it
		// holds chunkRays rays, so that many small API streams (<= 65 535 rays each under the
		// RayAccelerator.h Configuration) still become ONE launch that fills the machine.
		unsigned chunk = 0;
		size_t si = 0;
		uint32_t sBegin = 0;
		// Tuning::hostTaper: what follows the call's last H2D copy -- one traversal and one D2H copy -- overlaps nothing,
		// so the last chunks halve (remaining / 2, never below hostTaper K rays) instead of ending on a full-size one.
		uint64_t remaining = 0;
		for (const racc_cuda_stream_desc* d : hostStreams) remaining += d->count;
		const uint64_t taperFloor = tuning.hostTaper > 0 ? (uint64_t)tuning.hostTaper << 10 : 0;
		struct Segment { char* hResults; uint32_t offset, n; };
		std::vector<Segment> segs; // a chunk of many tiny streams has as many segments as it needs
		while (si < hostStreams.size()) {
			Target& t = targets[chunk % targets.size()];
			HostPipeline& pipe = *t.pipe;
			const int l = (int)(t.chunks % HostPipeline::kLanes);
			if (targets.size() > 1) RACC_CUDA_CHECK(cudaSetDevice(t.dev->ordinal));
			segs.clear();
			uint32_t filled = 0;
			uint32_t capacity = pipe.chunkRays;
			if (taperFloor && remaining < 2ull * pipe.chunkRays * targets.size()) {
				uint64_t half = (remaining / targets.size() + 1) / 2;
				if (half < taperFloor) half = taperFloor;
				if (half < capacity) capacity = (uint32_t)half;
			}
			while (si < hostStreams.size() && filled < capacity) {
				const racc_cuda_stream_desc* d = hostStreams[si];
				const uint32_t left = d->count - sBegin;
				const uint32_t room = capacity - filled;
				const uint32_t n = left < room ? left : room;
				RACC_CUDA_CHECK(cudaMemcpyAsync(pipe.dRays[l] + filled, static_cast<const char*>(d->rays) + (size_t)sBegin * 32, (size_t)n * 32,
				                                cudaMemcpyHostToDevice, pipe.lane[l]));
				segs.push_back(Segment{static_cast<char*>(d->results) + (size_t)sBegin * 16, filled, n});
				filled += n;
				sBegin += n;
				remaining -= n;
				if (sBegin == d->count) { ++si; sBegin = 0; }
			}
			TraceParams p{};
			fillSceneParams(p, s, t.rep, env, t.erep, device_counters && t.dev == dev ? static_cast<unsigned long long*>(device_counters) : t.dev->dFrame);
			p.nstreams = 1;
			p.total = filled;
			p.single.rays = pipe.dRays[l];
			p.single.results = pipe.dResults[l];
			p.single.begin = 0;
			p.single.count = filled;
			p.cursor = t.rep->dCursors + (t.rep->nextCursor.fetch_add(1) % kCursorRing);
			RACC_CUDA_CHECK(launchAny(s, tuning, t.dev, p, device_counters && t.dev == dev ? counterMode : 1, pipe.lane[l], &launches));
			// results go home: adjacent segments of one host buffer (a stream cut by nothing) are already one copy;
			// segments of different streams are separate copies of at least one stream each
			for (const Segment& g : segs)
				RACC_CUDA_CHECK(cudaMemcpyAsync(g.hResults, pipe.dResults[l] + g.offset, (size_t)g.n * 16, cudaMemcpyDeviceToHost, pipe.lane[l]));
			++t.chunks;
			++chunk;
		}
		for (Target& t : targets) {
			const int used = t.chunks < (unsigned)HostPipeline::kLanes ? (int)t.chunks : HostPipeline::kLanes;
			if (targets.size() > 1) RACC_CUDA_CHECK(cudaSetDevice(t.dev->ordinal));
			for (int l = 0; l < used; ++l)
				RACC_CUDA_CHECK(cudaEventRecord(t.pipe->done[l], t.pipe->lane[l]));
		}
		RACC_CUDA_CHECK(cudaSetDevice(dev->ordinal));
		for (Target& t : targets) {
			const int used = t.chunks < (unsigned)HostPipeline::kLanes ? (int)t.chunks : HostPipeline::kLanes;
			for (int l = 0; l < used; ++l)
				RACC_CUDA_CHECK(cudaStreamWaitEvent(stream, t.pipe->done[l], 0));
		}
	}
	countLaunches(launches);
	return 0;
}

} // namespace racc_b200

extern "C" {

int racc_cuda_trace(racc_cuda_scene* scene, racc_cuda_env* env, const racc_cuda_stream_desc* streams, uint32_t nstreams, void* cuda_stream) {
	return traceImpl(scene, env, streams, nstreams, cuda_stream, nullptr, false);
}

int racc_cuda_trace_counted(racc_cuda_scene* scene, racc_cuda_env* env, const racc_cuda_stream_desc* streams, uint32_t nstreams,
                            void* cuda_stream, void* device_counters, int detail) {
	if (!device_counters) return fail("racc_cuda_trace_counted: null counter buffer");
	return traceImpl(scene, env, streams, nstreams, cuda_stream, device_counters, detail != 0);
}

int racc_cuda_frame_reduce(racc_cuda_counters* totals, void* cuda_stream) {
	if (!currentDevice()) return -1;
	return commFrameReduce(totals, static_cast<cudaStream_t>(cuda_stream));
}

// Pinned host memory for ray streams (replaces the 4 KiB-aligned slab of RayAccelerator.cpp:532-568
// that the reference wraps in CL_MEM_USE_HOST_PTR buffers, :643-644).
void* racc_cuda_host_alloc(size_t bytes) {
	if (!currentDevice()) return nullptr;
	void* p = nullptr;
	cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable);
	if (e != cudaSuccess) {
		fail("cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
		return nullptr;
	}
	return p;
}

void racc_cuda_host_free(void* p) {
	if (p) cudaFreeHost(p);
}

// One CUDA stream per submitter (replaces the per-thread cl_command_queue, RayAccelerator.cpp:711-717).
void* racc_cuda_stream_create(void) {
	if (!currentDevice()) return nullptr;
	cudaStream_t s = nullptr;
	cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
	if (e != cudaSuccess) {
		fail("cudaStreamCreate failed: %s", cudaGetErrorString(e));
		return nullptr;
	}
	return s;
}

void racc_cuda_stream_destroy(void* cuda_stream) {
	if (cuda_stream) cudaStreamDestroy(static_cast<cudaStream_t>(cuda_stream));
}

int racc_cuda_sync(void* cuda_stream) {
	if (!currentDevice()) return -1;
	RACC_CUDA_CHECK(cudaStreamSynchronize(static_cast<cudaStream_t>(cuda_stream)));
	return 0;
}

void racc_cuda_thread_release(void) {
	int before = -1;
	cudaGetDevice(&before);
	for (auto& p : t_pipelines) p->release();
	t_pipelines.clear();
	releaseRenderScratch();
	if (before >= 0) cudaSetDevice(before);
}

int racc_cuda_generate_primary(const racc_cuda_camera* camera, uint32_t width, uint32_t height, uint32_t spp,
                               uint32_t jitter_seed, void* device_rays, void* cuda_stream) {
	if (!camera || !device_rays) return fail("racc_cuda_generate_primary: null argument");
	if (!currentDevice()) return -1;
	if ((uint64_t)width * height * spp > 0x7fffffffull) return fail("racc_cuda_generate_primary: too many rays");
	int launches = 0;
	RACC_CUDA_CHECK(launchGeneratePrimary(camera->origin, width, height, spp, jitter_seed, static_cast<DevRay*>(device_rays),
	                                      static_cast<cudaStream_t>(cuda_stream), &launches));
	countLaunches(launches);
	return 0;
}

int racc_cuda_generate_bounce(const racc_cuda_scene* scene_, const void* device_rays, const void* device_results, uint32_t count,
                              uint32_t seed, void* device_out_rays, uint32_t* device_out_count, void* cuda_stream) {
	racc_cuda_scene* s = const_cast<racc_cuda_scene*>(scene_);
	if (!s || !device_rays || !device_results || !device_out_rays || !device_out_count) return fail("racc_cuda_generate_bounce: null argument");
	DeviceState* dev = currentDevice();
	if (!dev) return -1;
	SceneReplica* r = s->on(dev->ordinal);
	if (!r) return fail("racc_cuda_generate_bounce: the scene has no copy on CUDA device %d", dev->ordinal);
	if (!r->dVerts) return fail("racc_cuda_generate_bounce: scene was created from images and has no vertex data");
	const size_t words = bounceScratchWords(count);
	if (words > r->bounceScratchWords) {
		RACC_CUDA_CHECK(cudaDeviceSynchronize());
		cudaFree(r->dBounceScratch);
		r->dBounceScratch = nullptr;
		RACC_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&r->dBounceScratch), words * sizeof(uint32_t)));
		r->bounceScratchWords = words;
	}
	int launches = 0;
	RACC_CUDA_CHECK(launchGenerateBounce(r->dVerts, r->dIndices, static_cast<const DevRay*>(device_rays),
	                                     static_cast<const float4*>(device_results), count, seed, static_cast<DevRay*>(device_out_rays),
	                                     device_out_count, r->dBounceScratch, static_cast<cudaStream_t>(cuda_stream), &launches));
	countLaunches(launches);
	return 0;
}

} // extern "C"
