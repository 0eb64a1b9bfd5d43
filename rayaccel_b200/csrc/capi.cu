// capi.cu -- implementation of the C-ABI declared in include/racc_b200.h.
// Device management, scene/environment upload, stream staging and kernel launch. No CPU fallback:
// every compute entry point fails with an error string when CUDA is unavailable.
#include "../../include/racc_b200.h"

#include "engine.h"
#include "scene_build.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

using namespace racc_b200;

namespace {

thread_local char g_error[512] = "";
std::mutex g_initMutex;
bool g_initialised = false;
int g_device = 0;
int g_smCount = 0;
Tuning g_tuning;
std::atomic<uint64_t> g_launches{0};

int fail(const char* fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_error, sizeof(g_error), fmt, ap);
	va_end(ap);
	return -1;
}

#define RACC_CUDA_CHECK(call)                                                                              \
	do {                                                                                                   \
		cudaError_t e_ = (call);                                                                           \
		if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while (0)

#define RACC_CUDA_CHECK_NULL(call)                                                                         \
	do {                                                                                                   \
		cudaError_t e_ = (call);                                                                           \
		if (e_ != cudaSuccess) {                                                                           \
			fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);              \
			return nullptr;                                                                                \
		}                                                                                                  \
	} while (0)

int envInt(const char* name, int fallback) {
	const char* v = getenv(name);
	return v && *v ? atoi(v) : fallback;
}

int ensureInit() {
	std::lock_guard<std::mutex> lock(g_initMutex);
	if (g_initialised) {
		cudaError_t e = cudaSetDevice(g_device);
		return e == cudaSuccess ? 0 : fail("cudaSetDevice(%d) failed: %s", g_device, cudaGetErrorString(e));
	}
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count <= 0)
		return fail("no CUDA device available (%s); the engine has no CPU fallback", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
	RACC_CUDA_CHECK(cudaGetDevice(&g_device));
	RACC_CUDA_CHECK(cudaDeviceGetAttribute(&g_smCount, cudaDevAttrMultiProcessorCount, g_device));
	{
		// stream-ordered scratch (stream tables, re-binning buffers) is recycled instead of being handed
		// back to the driver at every synchronisation
		cudaMemPool_t pool = nullptr;
		if (cudaDeviceGetDefaultMemPool(&pool, g_device) == cudaSuccess && pool) {
			unsigned long long keep = ~0ull;
			cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
		}
	}
	g_tuning.variant = envInt("RACC_B200_VARIANT", g_tuning.variant);
	g_tuning.blockThreads = envInt("RACC_B200_BLOCK", g_tuning.blockThreads);
	g_tuning.ctasPerSm = envInt("RACC_B200_CTAS_PER_SM", g_tuning.ctasPerSm);
	g_tuning.smemNodes = envInt("RACC_B200_SMEM_NODES", g_tuning.smemNodes);
	g_tuning.fetchThreshold = envInt("RACC_B200_FETCH_THRESHOLD", g_tuning.fetchThreshold);
	g_tuning.leafBail = envInt("RACC_B200_LEAF_BAIL", g_tuning.leafBail);
	g_tuning.innerBail = envInt("RACC_B200_INNER_BAIL", g_tuning.innerBail);
	g_tuning.carveout = envInt("RACC_B200_CARVEOUT", g_tuning.carveout);
	g_tuning.sortMode = envInt("RACC_B200_SORT", g_tuning.sortMode);
	g_tuning.sortOriginBits = envInt("RACC_B200_SORT_ORIGIN_BITS", g_tuning.sortOriginBits);
	g_tuning.sortDirBits = envInt("RACC_B200_SORT_DIR_BITS", g_tuning.sortDirBits);
	g_tuning.sortDirMajor = envInt("RACC_B200_SORT_DIR_MAJOR", g_tuning.sortDirMajor);
	g_tuning.buildDevice = envInt("RACC_B200_BUILD_DEVICE", g_tuning.buildDevice);
	g_tuning.smemStack = envInt("RACC_B200_SMEM_STACK", g_tuning.smemStack);
	g_tuning.hostZeroCopy = envInt("RACC_B200_HOST_ZERO_COPY", g_tuning.hostZeroCopy);
	g_tuning.hostTaper = envInt("RACC_B200_HOST_TAPER", g_tuning.hostTaper);
	g_tuning.whittedArena = envInt("RACC_B200_WHITTED_ARENA", g_tuning.whittedArena);
	g_tuning.whittedCombine = envInt("RACC_B200_WHITTED_COMBINE", g_tuning.whittedCombine);
	g_initialised = true;
	return 0;
}

constexpr int kCursorRing = 256;
constexpr size_t kAutoSortSceneBytes = 256u << 20; // twice the 126 MB L2
constexpr uint32_t kAutoDeviceBuildTriangles = 4096; // tiny scenes: not worth a dozen kernel launches

} // namespace

struct racc_cuda_scene {
	SceneImages host;            // host images; left empty when the scene was built on the device
	racc_cuda_scene_info info{}; // counts, depth and bounds of what is on the device
	uint32_t triangleCount = 0;
	uint32_t vertexCount = 0;    // of dVerts (0 when the scene was created from images)
	float4* dNodes = nullptr;
	float4* dPairs = nullptr;
	uint32_t* dRemap = nullptr;
	float4* dTNodes = nullptr;   // packed images walked by the default kernel (traverse_packed.cu)
	float4* dTPairs = nullptr;
	float4* dVerts = nullptr;    // for the synthetic bounce generator and the device-side renderer (indices)
	uint32_t* dIndices = nullptr;
	uint32_t* dCursors = nullptr;
	std::atomic<uint32_t> nextCursor{0};
	uint32_t* dBounceScratch = nullptr;
	size_t bounceScratchWords = 0;
};

struct racc_cuda_host_images {
	SceneImages images;
	uint32_t triangleCount = 0;
};

struct racc_cuda_env {
	float4* dTexels = nullptr;
	float4* dTexelPairs = nullptr; // (width+1) x height pairs of horizontally adjacent texels (traverse_packed.cu)
	uint32_t width = 0, height = 0;
};

// What the reference's example path tracer shades with (Renderer/SceneData.h), resident on the device.
struct racc_cuda_shading {
	float4* dNormals = nullptr;
	float4* dTriangleNormals = nullptr;
	uint16_t* dTriangleMaterials = nullptr;
	float4* dMaterials = nullptr;
	uint32_t vertexCount = 0, triangleCount = 0, materialCount = 0;
};

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

// device-private packed copies of the node and pair images, derived on the device
racc_cuda_scene* packScene(racc_cuda_scene* s) {
	cudaError_t e;
	int launches = 0;
	if ((e = cudaMalloc(reinterpret_cast<void**>(&s->dTNodes), (size_t)s->info.node_count * 64 + 64)) != cudaSuccess ||
	    (e = cudaMalloc(reinterpret_cast<void**>(&s->dTPairs), (size_t)s->info.pair_count * 64 + 64)) != cudaSuccess ||
	    (e = launchPackImages(s->dNodes, s->info.node_count, s->dPairs, s->info.pair_count, s->dTNodes, s->dTPairs, nullptr, &launches)) != cudaSuccess ||
	    (e = cudaDeviceSynchronize()) != cudaSuccess) {
		fail("scene packing failed: %s", cudaGetErrorString(e));
		racc_cuda_scene_destroy(s);
		return nullptr;
	}
	g_launches.fetch_add((uint64_t)launches);
	return s;
}

racc_cuda_scene* uploadScene(racc_cuda_scene* s, const float* verts4, uint32_t nverts, const uint32_t* indices, uint32_t nindices) {
	const SceneImages& h = s->host;
	auto bail = [&]() -> racc_cuda_scene* { racc_cuda_scene_destroy(s); return nullptr; };
	cudaError_t e;
	fillInfo(h, s->triangleCount, &s->info);
#define UP(dst, src, bytes)                                                                  \
	if ((e = cudaMalloc(reinterpret_cast<void**>(&dst), (bytes) != 0 ? (bytes) : 16)) != cudaSuccess || \
	    (e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice)) != cudaSuccess) {            \
		fail("scene upload failed: %s", cudaGetErrorString(e));                               \
		return bail();                                                                        \
	}
	UP(s->dNodes, h.nodes.data(), h.nodes.size() * sizeof(GpuNode))
	UP(s->dPairs, h.pairs.data(), h.pairs.size() * sizeof(GpuPair))
	UP(s->dRemap, h.remap.data(), h.remap.size() * sizeof(uint32_t))
	if (verts4 && indices) {
		UP(s->dVerts, verts4, (size_t)nverts * 16)
		UP(s->dIndices, indices, (size_t)nindices * 4)
	}
#undef UP
	if ((e = cudaMalloc(reinterpret_cast<void**>(&s->dCursors), kCursorRing * sizeof(uint32_t))) != cudaSuccess) {
		fail("scene upload failed: %s", cudaGetErrorString(e));
		return bail();
	}
	return packScene(s);
}

} // namespace

extern "C" {

int racc_cuda_abi_version(void) { return RACC_CUDA_ABI_VERSION; }

const char* racc_cuda_last_error(void) { return g_error; }

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
	if (devices && n > 0) {
		cudaError_t e = cudaSetDevice(devices[0]);
		if (e != cudaSuccess)
			return fail("cudaSetDevice(%d) failed: %s; the engine has no CPU fallback", devices[0], cudaGetErrorString(e));
		std::lock_guard<std::mutex> lock(g_initMutex);
		if (g_initialised && g_device != devices[0])
			g_initialised = false; // re-bind the process (tuning is re-read from the environment)
	}
	return ensureInit();
}

int racc_cuda_set_variant(int variant) {
	const int previous = g_tuning.variant;
	g_tuning.variant = variant;
	return previous;
}

// Extended tuning access for the benchmark sweeps: key 0 variant, 1 block threads, 2 CTAs/SM,
// 3 staged nodes, 4 fetch threshold. Returns the previous value.
int racc_cuda_set_tuning(int key, int value) {
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
	if (ensureInit()) return -1;
	RACC_CUDA_CHECK(cudaDeviceSynchronize());
	RACC_CUDA_CHECK(readWarpStats(reinterpret_cast<unsigned long long*>(out8), reset != 0));
	return 0;
}

racc_cuda_scene* racc_cuda_scene_create(const float* verts4, uint32_t nverts, const uint32_t* indices, uint32_t nindices) {
	if (!verts4 || !indices) { fail("racc_cuda_scene_create: null input"); return nullptr; }
	if (ensureInit()) return nullptr;
	racc_cuda_scene* s = new racc_cuda_scene();
	const char* why = "";
	bool built = false;
	s->triangleCount = nindices / 3;
	s->vertexCount = nverts;
	if (g_tuning.buildDevice == 2 || (g_tuning.buildDevice == 3 && nindices / 3 >= kAutoDeviceBuildTriangles)) {
		// the whole build on the device: the images never exist on the host
		DeviceSceneImages img;
		if (buildSceneImagesDevice(verts4, nverts, indices, nindices, &img, &why)) {
			s->dNodes = static_cast<float4*>(img.nodes);
			s->dPairs = static_cast<float4*>(img.pairs);
			s->dRemap = img.remap;
			s->info.node_count = img.nodeCount;
			s->info.pair_count = img.pairCount;
			s->info.real_pair_count = img.realPairs;
			s->info.remap_count = img.remapCount;
			s->info.depth = img.depth;
			s->info.triangle_count = s->triangleCount;
			for (int k = 0; k < 3; ++k) { s->info.bounds_min[k] = img.boundsMin[k]; s->info.bounds_max[k] = img.boundsMax[k]; }
			s->dVerts = static_cast<float4*>(img.verts);
			s->dIndices = img.indices;
			cudaError_t e;
			if ((e = cudaMalloc(reinterpret_cast<void**>(&s->dCursors), kCursorRing * sizeof(uint32_t))) != cudaSuccess) {
				fail("scene upload failed: %s", cudaGetErrorString(e));
				racc_cuda_scene_destroy(s);
				return nullptr;
			}
			return packScene(s);
		}
		fprintf(stderr, "RayAccelerator: device scene build declined (%s); building on the host\n", why);
	}
	else if (g_tuning.buildDevice == 1) {
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
	if (ensureInit()) return nullptr;
	racc_cuda_scene* s = new racc_cuda_scene();
	s->host.nodes.assign(static_cast<const GpuNode*>(nodes), static_cast<const GpuNode*>(nodes) + node_count);
	s->host.pairs.assign(static_cast<const GpuPair*>(pairs), static_cast<const GpuPair*>(pairs) + pair_count);
	s->host.remap.assign(remap, remap + remap_count);
	s->host.realPairs = remap_count / 2;
	{
		// scene bounds = union of the root's two child boxes (used only by the ray re-binning keys)
		const GpuNode& root = s->host.nodes[0];
		for (int k = 0; k < 3; ++k) {
			s->host.boundsMin[k] = root.leftMin[k] < root.rightMin[k] ? root.leftMin[k] : root.rightMin[k];
			s->host.boundsMax[k] = root.leftMax[k] > root.rightMax[k] ? root.leftMax[k] : root.rightMax[k];
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
	return uploadScene(s, nullptr, 0, nullptr, 0);
}

void racc_cuda_scene_destroy(racc_cuda_scene* s) {
	if (!s) return;
	cudaFree(s->dNodes);
	cudaFree(s->dPairs);
	cudaFree(s->dRemap);
	cudaFree(s->dTNodes);
	cudaFree(s->dTPairs);
	cudaFree(s->dVerts);
	cudaFree(s->dIndices);
	cudaFree(s->dCursors);
	cudaFree(s->dBounceScratch);
	delete s;
}

int racc_cuda_scene_get_info(const racc_cuda_scene* s, racc_cuda_scene_info* info) {
	if (!s || !info) return fail("racc_cuda_scene_get_info: null argument");
	*info = s->info;
	return 0;
}

int racc_cuda_scene_download(const racc_cuda_scene* s, void* nodes, void* pairs, uint32_t* remap) {
	if (!s) return fail("racc_cuda_scene_download: null scene");
	// read back from the device so the test sees what the kernel sees
	if (nodes) RACC_CUDA_CHECK(cudaMemcpy(nodes, s->dNodes, (size_t)s->info.node_count * sizeof(GpuNode), cudaMemcpyDeviceToHost));
	if (pairs) RACC_CUDA_CHECK(cudaMemcpy(pairs, s->dPairs, (size_t)s->info.pair_count * sizeof(GpuPair), cudaMemcpyDeviceToHost));
	if (remap) RACC_CUDA_CHECK(cudaMemcpy(remap, s->dRemap, (size_t)s->info.remap_count * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	return 0;
}

racc_cuda_env* racc_cuda_env_create(const float* rgba, uint32_t width, uint32_t height) {
	if (!rgba || !width || !height) { fail("racc_cuda_env_create: null or empty image"); return nullptr; }
	if (ensureInit()) return nullptr;
	racc_cuda_env* env = new racc_cuda_env();
	env->width = width;
	env->height = height;
	const size_t bytes = (size_t)width * height * 16;
	cudaError_t e;
	if ((e = cudaMalloc(reinterpret_cast<void**>(&env->dTexels), bytes)) != cudaSuccess ||
	    (e = cudaMemcpy(env->dTexels, rgba, bytes, cudaMemcpyHostToDevice)) != cudaSuccess) {
		fail("environment upload failed: %s", cudaGetErrorString(e));
		cudaFree(env->dTexels);
		delete env;
		return nullptr;
	}
	int launches = 0;
	if ((e = cudaMalloc(reinterpret_cast<void**>(&env->dTexelPairs), ((size_t)width + 1) * height * 32)) != cudaSuccess ||
	    (e = launchPackEnv(env->dTexels, width, height, env->dTexelPairs, nullptr, &launches)) != cudaSuccess ||
	    (e = cudaDeviceSynchronize()) != cudaSuccess) {
		fail("environment packing failed: %s", cudaGetErrorString(e));
		racc_cuda_env_destroy(env);
		return nullptr;
	}
	g_launches.fetch_add((uint64_t)launches);
	return env;
}

void racc_cuda_env_destroy(racc_cuda_env* env) {
	if (!env) return;
	cudaFree(env->dTexels);
	cudaFree(env->dTexelPairs);
	delete env;
}

namespace {

// Staging pipeline for HOST ray streams: chunks alternate over a few internal CUDA streams so that
// the H2D copy of chunk k+1, the traversal of chunk k and the D2H copy of chunk k-1 overlap (PCIe
// is full duplex). One pipeline per calling host thread, so concurrent submitters never share
// staging buffers. Replaces the reference's zero-copy CL_MEM_USE_HOST_PTR streams
// (RayAccelerator.cpp:643-644), which a discrete GPU does not have.
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

	int init() {
		if (ready && device == g_device) return 0;
		release();
		device = g_device;
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

	void release() {
		if (!ready) return;
		for (int l = 0; l < kLanes; ++l) {
			if (lane[l]) { cudaStreamSynchronize(lane[l]); cudaStreamDestroy(lane[l]); }
			if (done[l]) cudaEventDestroy(done[l]);
			cudaFree(dRays[l]);
			cudaFree(dResults[l]);
			lane[l] = nullptr; done[l] = nullptr; dRays[l] = nullptr; dResults[l] = nullptr;
		}
		if (fork) cudaEventDestroy(fork);
		fork = nullptr;
		ready = false;
	}

	~HostPipeline() { /* process teardown: the context may already be gone, leak on purpose */ }
};

thread_local HostPipeline t_pipeline;

// Internal streams of racc_cuda_path_trace (one set per calling host thread), see there.
struct PathLanes {
	static constexpr int kMax = 4;
	bool ready = false;
	int device = -1;
	int count = 2;
	cudaStream_t stream[kMax] = {};
	cudaEvent_t done[kMax] = {};
	cudaEvent_t fork = nullptr;
	uint32_t* hostCounts = nullptr; // pinned: the next wave's size of each lane

	int init() {
		if (ready && device == g_device) return 0;
		device = g_device;
		count = envInt("RACC_B200_PATH_LANES", 2);
		if (count < 1) count = 1;
		if (count > kMax) count = kMax;
		RACC_CUDA_CHECK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
		RACC_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&hostCounts), kMax * sizeof(uint32_t), cudaHostAllocPortable));
		for (int l = 0; l < kMax; ++l) {
			RACC_CUDA_CHECK(cudaStreamCreateWithFlags(&stream[l], cudaStreamNonBlocking));
			RACC_CUDA_CHECK(cudaEventCreateWithFlags(&done[l], cudaEventDisableTiming));
		}
		ready = true;
		return 0;
	}
	~PathLanes() { /* process teardown: the context may already be gone, leak on purpose */ }
};

thread_local PathLanes t_pathLanes;

// Wave buffers of racc_cuda_whitted_trace kept between waves, batches and calls (Tuning::whittedArena), one set per
// calling host thread. A wave's size is only known after the previous one was shaded and differs from frame to frame, so
// per-wave stream-ordered allocations keep asking the pool for sizes it has no block for; these buffers only ever grow
// (by a quarter more than asked). Slots: 0/1 rays (ping-pong), 2/3 states, 4 results. A buffer is grown only while it
// holds nothing live: the results before a wave is traced, the next wave's rays/states before they are written.
struct WhittedArena {
	static constexpr int kSlots = 5;
	int device = -1;
	void* p[kSlots] = {};
	size_t cap[kSlots] = {};
	cudaEvent_t idle = nullptr; // end of the previous call's work on these buffers
	bool idleRecorded = false;

	// a later call may come on another CUDA stream: it waits for the previous call's kernels before touching the buffers
	int begin(cudaStream_t stream) {
		if (device != g_device) {
			for (int k = 0; k < kSlots; ++k) { p[k] = nullptr; cap[k] = 0; } // another device's pointers: dropped, not freed here
			idle = nullptr;
			idleRecorded = false;
			device = g_device;
		}
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
	~WhittedArena() { /* process teardown: the context may already be gone, leak on purpose */ }
};

thread_local WhittedArena t_whittedArena;

void fillSceneParams(TraceParams& p, racc_cuda_scene* s, racc_cuda_env* env, void* device_counters) {
	p.nodes = s->dNodes;
	p.pairs = s->dPairs;
	p.remap = s->dRemap;
	p.env = env ? env->dTexels : nullptr;
	p.envWidth = env ? env->width : 0;
	p.envHeight = env ? env->height : 0;
	p.nodeCount = s->info.node_count;
	p.counters = static_cast<unsigned long long*>(device_counters);
	p.tnodes = s->dTNodes;
	p.tpairs = s->dTPairs;
	p.perm = nullptr;
	p.envPairs = env ? env->dTexelPairs : nullptr;
}

bool sceneExceedsL2(const racc_cuda_scene* s) {
	return ((size_t)s->info.node_count + s->info.pair_count) * 64 > kAutoSortSceneBytes;
}

cudaError_t launchAny(const racc_cuda_scene* s, const TraceParams& p, int counterMode, cudaStream_t stream, int* launches) {
	if (g_tuning.variant != 3)
		return launchTrace(p, g_tuning, counterMode, g_smCount, stream, launches);
	Tuning t = g_tuning;
	if (t.smemStack < 0) t.smemStack = sceneExceedsL2(s) ? 16 : 0;
	return launchTracePacked(p, t, counterMode, g_smCount, stream, launches);
}

int traceImpl(racc_cuda_scene* s, racc_cuda_env* env, const racc_cuda_stream_desc* streams, uint32_t nstreams,
              void* cuda_stream, void* device_counters, bool fullCounters) {
	if (!s) return fail("racc_cuda_trace: null scene");
	if (!streams && nstreams) return fail("racc_cuda_trace: null stream list");
	if (ensureInit()) return -1;
	cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);

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
			if (g_tuning.hostZeroCopy && !((reinterpret_cast<uintptr_t>(d.rays) | reinterpret_cast<uintptr_t>(d.results)) & 15)) {
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

	int launches = 0;
	if (total) {
		TraceParams p{};
		fillSceneParams(p, s, env, device_counters);
		p.nstreams = (uint32_t)refs.size();
		p.total = (uint32_t)total;
		p.single = refs[0];
		p.cursor = s->dCursors + (s->nextCursor.fetch_add(1) % kCursorRing);
		void* dRefs = nullptr;
		if (refs.size() > 1) {
			RACC_CUDA_CHECK(cudaMallocAsync(&dRefs, refs.size() * sizeof(StreamRef), stream));
			RACC_CUDA_CHECK(cudaMemcpyAsync(dRefs, refs.data(), refs.size() * sizeof(StreamRef), cudaMemcpyHostToDevice, stream));
			p.streams = static_cast<const StreamRef*>(dRefs);
		}
		void* sortScratch = nullptr;
		const bool rebin = g_tuning.variant == 3 && total >= 4096 && !zeroCopy &&
		                   (g_tuning.sortMode == 1 || (g_tuning.sortMode == 2 && sceneExceedsL2(s) && total >= (1u << 18)));
		if (rebin) {
			// re-bin the launch: visiting order by origin/direction key, results stay index-parallel
			// an optimisation only: when the scratch does not fit, trace in arrival order
			if (cudaMallocAsync(&sortScratch, raySortScratchBytes(p.total), stream) != cudaSuccess) {
				cudaGetLastError();
				sortScratch = nullptr;
			}
			else {
				RACC_CUDA_CHECK(launchRaySort(p, s->info.bounds_min, s->info.bounds_max, g_tuning.sortOriginBits, g_tuning.sortDirBits,
				                              g_tuning.sortDirMajor, sortScratch, g_smCount, stream, &p.perm, &launches));
			}
		}
		RACC_CUDA_CHECK(launchAny(s, p, device_counters ? (fullCounters ? 2 : 1) : 0, stream, &launches));
		if (sortScratch) RACC_CUDA_CHECK(cudaFreeAsync(sortScratch, stream));
		if (dRefs) RACC_CUDA_CHECK(cudaFreeAsync(dRefs, stream));
	}

	if (!hostStreams.empty()) {
		HostPipeline& pipe = t_pipeline;
		if (pipe.init()) return -1;
		RACC_CUDA_CHECK(cudaEventRecord(pipe.fork, stream));
		for (int l = 0; l < HostPipeline::kLanes; ++l)
			RACC_CUDA_CHECK(cudaStreamWaitEvent(pipe.lane[l], pipe.fork, 0));
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
		const uint64_t taperFloor = g_tuning.hostTaper > 0 ? (uint64_t)g_tuning.hostTaper << 10 : 0;
		while (si < hostStreams.size()) {
			const int l = (int)(chunk % HostPipeline::kLanes);
			struct Segment { char* hResults; uint32_t offset, n; };
			Segment segs[64];
			int nsegs = 0;
			uint32_t filled = 0;
			uint32_t capacity = pipe.chunkRays;
			if (taperFloor && remaining < 2ull * pipe.chunkRays) {
				uint64_t half = (remaining + 1) / 2;
				if (half < taperFloor) half = taperFloor;
				if (half < capacity) capacity = (uint32_t)half;
			}
			while (si < hostStreams.size() && filled < capacity && nsegs < 64) {
				const racc_cuda_stream_desc* d = hostStreams[si];
				const uint32_t left = d->count - sBegin;
				const uint32_t room = capacity - filled;
				const uint32_t n = left < room ? left : room;
				RACC_CUDA_CHECK(cudaMemcpyAsync(pipe.dRays[l] + filled, static_cast<const char*>(d->rays) + (size_t)sBegin * 32, (size_t)n * 32,
				                                cudaMemcpyHostToDevice, pipe.lane[l]));
				segs[nsegs++] = Segment{static_cast<char*>(d->results) + (size_t)sBegin * 16, filled, n};
				filled += n;
				sBegin += n;
				remaining -= n;
				if (sBegin == d->count) { ++si; sBegin = 0; }
			}
			TraceParams p{};
			fillSceneParams(p, s, env, device_counters);
			p.nstreams = 1;
			p.total = filled;
			p.single.rays = pipe.dRays[l];
			p.single.results = pipe.dResults[l];
			p.single.begin = 0;
			p.single.count = filled;
			p.cursor = s->dCursors + (s->nextCursor.fetch_add(1) % kCursorRing);
			RACC_CUDA_CHECK(launchAny(s, p, device_counters ? (fullCounters ? 2 : 1) : 0, pipe.lane[l], &launches));
			for (int k = 0; k < nsegs; ++k)
				RACC_CUDA_CHECK(cudaMemcpyAsync(segs[k].hResults, pipe.dResults[l] + segs[k].offset, (size_t)segs[k].n * 16, cudaMemcpyDeviceToHost, pipe.lane[l]));
			++chunk;
		}
		const int used = chunk < (unsigned)HostPipeline::kLanes ? (int)chunk : HostPipeline::kLanes;
		for (int l = 0; l < used; ++l) {
			RACC_CUDA_CHECK(cudaEventRecord(pipe.done[l], pipe.lane[l]));
			RACC_CUDA_CHECK(cudaStreamWaitEvent(stream, pipe.done[l], 0));
		}
	}
	g_launches.fetch_add((uint64_t)launches);
	return 0;
}

} // namespace

int racc_cuda_trace(racc_cuda_scene* scene, racc_cuda_env* env, const racc_cuda_stream_desc* streams, uint32_t nstreams, void* cuda_stream) {
	return traceImpl(scene, env, streams, nstreams, cuda_stream, nullptr, false);
}

int racc_cuda_trace_counted(racc_cuda_scene* scene, racc_cuda_env* env, const racc_cuda_stream_desc* streams, uint32_t nstreams,
                            void* cuda_stream, void* device_counters, int detail) {
	if (!device_counters) return fail("racc_cuda_trace_counted: null counter buffer");
	return traceImpl(scene, env, streams, nstreams, cuda_stream, device_counters, detail != 0);
}

// Pinned host memory for ray streams (replaces the 4 KiB-aligned slab of RayAccelerator.cpp:532-568
// that the reference wraps in CL_MEM_USE_HOST_PTR buffers, :643-644).
void* racc_cuda_host_alloc(size_t bytes) {
	if (ensureInit()) return nullptr;
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
	if (ensureInit()) return nullptr;
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
	if (ensureInit()) return -1;
	RACC_CUDA_CHECK(cudaStreamSynchronize(static_cast<cudaStream_t>(cuda_stream)));
	return 0;
}

int racc_cuda_generate_primary(const racc_cuda_camera* camera, uint32_t width, uint32_t height, uint32_t spp,
                               uint32_t jitter_seed, void* device_rays, void* cuda_stream) {
	if (!camera || !device_rays) return fail("racc_cuda_generate_primary: null argument");
	if (ensureInit()) return -1;
	if ((uint64_t)width * height * spp > 0x7fffffffull) return fail("racc_cuda_generate_primary: too many rays");
	int launches = 0;
	RACC_CUDA_CHECK(launchGeneratePrimary(camera->origin, width, height, spp, jitter_seed, static_cast<DevRay*>(device_rays),
	                                      static_cast<cudaStream_t>(cuda_stream), &launches));
	g_launches.fetch_add((uint64_t)launches);
	return 0;
}

int racc_cuda_generate_bounce(const racc_cuda_scene* scene_, const void* device_rays, const void* device_results, uint32_t count,
                              uint32_t seed, void* device_out_rays, uint32_t* device_out_count, void* cuda_stream) {
	racc_cuda_scene* s = const_cast<racc_cuda_scene*>(scene_);
	if (!s || !device_rays || !device_results || !device_out_rays || !device_out_count) return fail("racc_cuda_generate_bounce: null argument");
	if (!s->dVerts) return fail("racc_cuda_generate_bounce: scene was created from images and has no vertex data");
	if (ensureInit()) return -1;
	const size_t words = bounceScratchWords(count);
	if (words > s->bounceScratchWords) {
		RACC_CUDA_CHECK(cudaDeviceSynchronize());
		cudaFree(s->dBounceScratch);
		s->dBounceScratch = nullptr;
		RACC_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&s->dBounceScratch), words * sizeof(uint32_t)));
		s->bounceScratchWords = words;
	}
	int launches = 0;
	RACC_CUDA_CHECK(launchGenerateBounce(s->dVerts, s->dIndices, static_cast<const DevRay*>(device_rays),
	                                     static_cast<const float4*>(device_results), count, seed, static_cast<DevRay*>(device_out_rays),
	                                     device_out_count, s->dBounceScratch, static_cast<cudaStream_t>(cuda_stream), &launches));
	g_launches.fetch_add((uint64_t)launches);
	return 0;
}

// ---- device-side wavefront path tracer (pathtrace.cu) ----

racc_cuda_shading* racc_cuda_shading_create(const racc_cuda_shading_desc* d) {
	if (!d || !d->normals4 || !d->triangle_normals4 || !d->triangle_materials || !d->materials_ke4) {
		fail("racc_cuda_shading_create: null input");
		return nullptr;
	}
	if (!d->material_count) { fail("racc_cuda_shading_create: no materials"); return nullptr; }
	if (ensureInit()) return nullptr;
	racc_cuda_shading* sh = new racc_cuda_shading();
	sh->vertexCount = d->vertex_count;
	sh->triangleCount = d->triangle_count;
	sh->materialCount = d->material_count;
	cudaError_t e;
#define UP(dst, src, bytes)                                                                                 \
	if ((e = cudaMalloc(reinterpret_cast<void**>(&dst), (bytes) != 0 ? (bytes) : 16)) != cudaSuccess ||     \
	    (e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice)) != cudaSuccess) {                         \
		fail("racc_cuda_shading_create: upload failed: %s", cudaGetErrorString(e));                         \
		racc_cuda_shading_destroy(sh);                                                                      \
		return nullptr;                                                                                     \
	}
	UP(sh->dNormals, d->normals4, (size_t)d->vertex_count * 16)
	UP(sh->dTriangleNormals, d->triangle_normals4, (size_t)d->triangle_count * 16)
	UP(sh->dTriangleMaterials, d->triangle_materials, (size_t)d->triangle_count * 2)
	UP(sh->dMaterials, d->materials_ke4, (size_t)d->material_count * 16)
#undef UP
	return sh;
}

void racc_cuda_shading_destroy(racc_cuda_shading* sh) {
	if (!sh) return;
	cudaFree(sh->dNormals);
	cudaFree(sh->dTriangleNormals);
	cudaFree(sh->dTriangleMaterials);
	cudaFree(sh->dMaterials);
	delete sh;
}

int racc_cuda_path_trace(racc_cuda_scene* s, racc_cuda_env* env, const racc_cuda_shading* sh, const racc_cuda_camera* camera,
                         const racc_cuda_path_desc* d, float* framebuffer4, uint64_t* wave_rays, void* cuda_stream) {
	if (!s || !sh || !camera || !d || !framebuffer4) return fail("racc_cuda_path_trace: null argument");
	if (!s->dIndices) return fail("racc_cuda_path_trace: scene was created from images and has no index data");
	if (sh->triangleCount != s->triangleCount || sh->vertexCount < s->vertexCount)
		return fail("racc_cuda_path_trace: shading data (%u triangles, %u vertices) does not match the scene (%u, %u)", sh->triangleCount,
		            sh->vertexCount, s->triangleCount, s->vertexCount);
	if (d->max_depth > 62) return fail("racc_cuda_path_trace: max_depth %u > 62", d->max_depth);
	if (ensureInit()) return -1;
	const uint64_t pixels = (uint64_t)d->width * d->height;
	if (!pixels || !d->spp) return 0;
	if (pixels > (1ull << 28)) return fail("racc_cuda_path_trace: viewport too large");
	cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
	// paths per batch: whole samples, about 32 M paths (128 B of device memory each, 4 GB) unless the caller says otherwise:
	// the deeper waves of a batch are a tenth of its size, and a persistent launch over < 1 M rays is mostly tail
	// (1920x1080, 16 spp: 4.4 / 5.8 / 6.3 Gray/s at 2 M / 8 M / 33 M paths per batch, profiles/r01_render_device.md)
	uint32_t batchSpp = d->batch_spp ? d->batch_spp : (uint32_t)((32ull << 20) / pixels);
	if (batchSpp < 1) batchSpp = 1;
	if (batchSpp > d->spp) batchSpp = d->spp;
	if (pixels * batchSpp > 0x7fffffffull) return fail("racc_cuda_path_trace: batch of %u samples is too large", batchSpp);
	const size_t paths = (size_t)pixels * batchSpp;
	const bool hostFb = (d->flags & RACC_CUDA_FRAMEBUFFER_HOST) != 0;

	// Lanes: a batch is cut into contiguous path ranges that advance bounce by bounce on their own streams. While the host
	// waits for one lane's wave size, the other lane's launches are queued, and the tail of a traversal launch (few long
	// paths left, most SMs idle) is filled by the other lane's kernels.
	PathLanes& lanes = t_pathLanes;
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
	                         lanePaths * nlanes * 16, paths * 16, (size_t)PathLanes::kMax * 64 * sizeof(uint32_t), hostFb ? (size_t)pixels * 16 : 0};
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
		uint32_t* counts; // device, one per depth
		cudaStream_t stream;
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
		lane[l].counts = counts + (size_t)l * 64;
		lane[l].stream = nlanes > 1 ? lanes.stream[l] : stream;
	}

	int launches = 0;
	for (uint32_t done = 0; done < d->spp; done += batchSpp) {
		const uint32_t spp = d->spp - done < batchSpp ? d->spp - done : batchSpp;
		const uint32_t sampleBase = d->sample_base + done;
		const uint32_t batchPaths = (uint32_t)(pixels * spp);
		RACC_CUDA_CHECK(cudaMemsetAsync(radiance, 0, (size_t)batchPaths * 16, stream));
		RACC_CUDA_CHECK(cudaMemsetAsync(counts, 0, (size_t)PathLanes::kMax * 64 * sizeof(uint32_t), stream));
		if (nlanes > 1) {
			RACC_CUDA_CHECK(cudaEventRecord(lanes.fork, stream));
			for (int l = 0; l < nlanes; ++l) RACC_CUDA_CHECK(cudaStreamWaitEvent(lane[l].stream, lanes.fork, 0));
		}
		// one wave of one lane: trace, shade + compact, and (unless it was the last bounce) ask for the next wave's size
		auto enqueue = [&](Lane& ln, int l) -> int {
			if (wave_rays) wave_rays[ln.depth] += ln.count;
			racc_cuda_stream_desc sd{};
			sd.rays = ln.rays[ln.cur];
			sd.results = ln.results;
			sd.count = ln.count;
			sd.flags = 0;
			if (traceImpl(s, env, &sd, 1, ln.stream, nullptr, false)) return -1;
			PathShadeParams p{};
			p.rays = ln.rays[ln.cur]; p.results = ln.results; p.states = ln.states[ln.cur]; p.count = ln.count;
			p.depth = ln.depth; p.maxDepth = d->max_depth; p.seed = d->seed; p.pixels = (uint32_t)pixels; p.sampleBase = sampleBase;
			p.indices = s->dIndices; p.normals = sh->dNormals; p.triangleNormals = sh->dTriangleNormals;
			p.triangleMaterials = sh->dTriangleMaterials; p.materials = sh->dMaterials;
			p.triangleCount = sh->triangleCount; p.materialCount = sh->materialCount;
			p.outRays = ln.rays[ln.cur ^ 1]; p.outStates = ln.states[ln.cur ^ 1]; p.outCount = ln.counts + ln.depth; p.radiance = radiance;
			RACC_CUDA_CHECK(launchPathShade(p, ln.stream, &launches));
			if (ln.depth == d->max_depth) { ln.busy = false; return 0; } // nothing is extended past the last bounce
			RACC_CUDA_CHECK(cudaMemcpyAsync(lanes.hostCounts + l, ln.counts + ln.depth, sizeof(uint32_t), cudaMemcpyDeviceToHost, ln.stream));
			return 0;
		};
		int active = 0;
		for (int l = 0; l < nlanes; ++l) {
			Lane& ln = lane[l];
			const size_t first = (size_t)l * lanePaths;
			ln.count = first < batchPaths ? (uint32_t)(batchPaths - first < lanePaths ? batchPaths - first : lanePaths) : 0;
			ln.depth = 0;
			ln.cur = 0;
			ln.busy = ln.count != 0;
			if (!ln.busy) continue;
			RACC_CUDA_CHECK(launchPathPrimary(camera->origin, d->width, d->height, sampleBase, (uint32_t)first, ln.count, d->seed, ln.rays[0],
			                                  ln.states[0], ln.stream, &launches));
			if (enqueue(ln, l)) return -1;
			active += ln.busy;
		}
		// round robin: the size of a lane's next wave decides its launch -- the one host round trip per bounce and lane
		for (int l = 0; active; l = (l + 1) % nlanes) {
			Lane& ln = lane[l];
			if (!ln.busy) continue;
			RACC_CUDA_CHECK(cudaStreamSynchronize(ln.stream));
			ln.count = lanes.hostCounts[l];
			ln.depth += 1;
			ln.cur ^= 1;
			if (!ln.count) { ln.busy = false; --active; continue; }
			if (enqueue(ln, l)) return -1;
			if (!ln.busy) --active;
		}
		if (nlanes > 1)
			for (int l = 0; l < nlanes; ++l) {
				RACC_CUDA_CHECK(cudaEventRecord(lanes.done[l], lane[l].stream));
				RACC_CUDA_CHECK(cudaStreamWaitEvent(stream, lanes.done[l], 0));
			}
		RACC_CUDA_CHECK(launchPathAccumulate(radiance, (uint32_t)pixels, spp, fb, stream, &launches));
	}
	buf.joined = true;
	g_launches.fetch_add((uint64_t)launches);
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
	if (!s->dIndices) return fail("racc_cuda_whitted_trace: scene was created from images and has no index data");
	if (sh->triangleCount != s->triangleCount || sh->vertexCount < s->vertexCount)
		return fail("racc_cuda_whitted_trace: shading data (%u triangles, %u vertices) does not match the scene (%u, %u)", sh->triangleCount,
		            sh->vertexCount, s->triangleCount, s->vertexCount);
	if (d->max_depth > 62) return fail("racc_cuda_whitted_trace: max_depth %u > 62", d->max_depth);
	if (ensureInit()) return -1;
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
	const bool useArena = g_tuning.whittedArena != 0;
	WhittedArena& arena = t_whittedArena;
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
			p.indices = s->dIndices; p.normals = sh->dNormals; p.triangleNormals = sh->dTriangleNormals; p.triangleCount = sh->triangleCount;
			p.outRays = nextRays; p.outStates = nextStates; p.outCount = counts + depth; p.accumulators = acc;
			p.combine = g_tuning.whittedCombine != 0;
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
	g_launches.fetch_add((uint64_t)launches);
	if (hostFb) {
		RACC_CUDA_CHECK(cudaMemcpyAsync(framebuffer4, fb, (size_t)pixels * 16, cudaMemcpyDeviceToHost, stream));
		RACC_CUDA_CHECK(cudaStreamSynchronize(stream));
	}
	return 0;
}

} // extern "C"
