// capi_internal.h -- state shared by the files that implement the C-ABI (capi.cu: devices, scenes, ray streams;
// capi_render.cu: the device-side renderers; comm.cu: the per-frame hit reduction over NCCL). Nothing here crosses the
// boundary: include/racc_b200.h sees only opaque handles.
//
// Devices. The reference owns ONE OpenCL device per racc::Context and feeds it from `gpuSubmissionThreads` host threads,
// each with its own command queue (RayAccelerator.cpp:335-414,711-717). Here a process may drive several B200s: every
// CUDA device the process has named in racc_cuda_init() gets a DeviceState; a host thread works on its *device set*
// (the list it passed to racc_cuda_init last; the first entry is the device its DEVICE streams and cudaStream_t handles
// belong to), scenes / environments / shading data are replicated on every device of the creating thread's set, and
// per-thread scratch (staging pipelines, renderer lanes) exists once per thread and device. There is no process-wide
// "current device": two contexts on different devices do not disturb each other.
#pragma once

#include "../../include/racc_b200.h"

#include "engine.h"
#include "scene_build.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <memory>
#include <vector>

namespace racc_b200 {

constexpr int kMaxDevices = 64;
constexpr int kCursorRing = 256;
constexpr size_t kAutoSortSceneBytes = 256u << 20;   // twice the 126 MB L2
constexpr uint32_t kAutoDeviceBuildTriangles = 4096; // tiny scenes: not worth a dozen kernel launches

struct DeviceState {
	bool ready = false;
	int ordinal = -1;
	int smCount = 0;
	size_t totalBytes = 0; // device memory, noted once
	// racc_cuda_counters accumulated by every traversal launch on this device since the last racc_cuda_frame_reduce
	// (rays + hits: one atomic per warp at kernel exit), and the reduced record
	unsigned long long* dFrame = nullptr;
	unsigned long long* dFrameTotal = nullptr;
	cudaStream_t reduceStream = nullptr;
	cudaEvent_t reduceReady = nullptr, reduceDone = nullptr;
};

// thread-local error text + return -1
int fail(const char* fmt, ...);

// The calling thread's bound device (first of its set), made current. nullptr (error set) without a CUDA device.
DeviceState* currentDevice();
// The calling thread's device set, bound device first. Valid after currentDevice() succeeded.
const std::vector<int>& currentDeviceSet();
// Makes `ordinal` the current CUDA device and returns its state (initialising it on first use); nullptr on failure.
DeviceState* useDevice(int ordinal);
Tuning tuningSnapshot();
void countLaunches(int launches);

#define RACC_CUDA_CHECK(call)                                                                              \
	do {                                                                                                   \
		cudaError_t e_ = (call);                                                                           \
		if (e_ != cudaSuccess) return ::racc_b200::fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while (0)

#define RACC_CUDA_CHECK_NULL(call)                                                                         \
	do {                                                                                                   \
		cudaError_t e_ = (call);                                                                           \
		if (e_ != cudaSuccess) {                                                                           \
			::racc_b200::fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);  \
			return nullptr;                                                                                \
		}                                                                                                  \
	} while (0)

// One device's copy of a scene: the reference-format images, the packed images the default kernel walks, the optional
// compressed node image (traverse_wide.cu), the mesh (renderers, bounce generator) and the work cursors.
struct SceneReplica {
	int device = -1;
	float4* dNodes = nullptr;
	float4* dPairs = nullptr;
	uint32_t* dRemap = nullptr;
	float4* dTNodes = nullptr;   // packed images walked by the default kernel (traverse_packed.cu)
	float4* dTPairs = nullptr;
	void* dQNodes = nullptr;     // 32-byte quantised nodes (traverse_packed.cu, variant 4)
	float4* dVerts = nullptr;    // for the synthetic bounce generator and the device-side renderers
	uint32_t* dIndices = nullptr;
	uint32_t* dCursors = nullptr;
	std::atomic<uint32_t> nextCursor{0};
	uint32_t* dBounceScratch = nullptr;
	size_t bounceScratchWords = 0;
};

struct EnvReplica {
	int device = -1;
	float4* dTexels = nullptr;
	float4* dTexelPairs = nullptr; // (width+1) x height pairs of horizontally adjacent texels (traverse_packed.cu)
};

struct ShadingReplica {
	int device = -1;
	float4* dNormals = nullptr;
	float4* dTriangleNormals = nullptr;
	uint16_t* dTriangleMaterials = nullptr;
	float4* dMaterials = nullptr;
};

template <typename R>
R* replicaOn(const std::vector<std::unique_ptr<R>>& v, int device) {
	for (const auto& r : v)
		if (r->device == device) return r.get();
	return nullptr;
}

} // namespace racc_b200

struct racc_cuda_scene {
	racc_b200::SceneImages host;  // host images; left empty when the scene was built on the device
	racc_cuda_scene_info info{};  // counts, depth and bounds of what is on the devices
	uint32_t triangleCount = 0;
	uint32_t vertexCount = 0;     // of dVerts (0 when the scene was created from images)
	float qOrigin[3] = {0, 0, 0}, qCell[3] = {0, 0, 0}; // grid of the quantised node image
	std::vector<std::unique_ptr<racc_b200::SceneReplica>> replicas;
	racc_b200::SceneReplica* on(int device) const { return racc_b200::replicaOn(replicas, device); }
};

struct racc_cuda_host_images {
	racc_b200::SceneImages images;
	uint32_t triangleCount = 0;
};

struct racc_cuda_env {
	uint32_t width = 0, height = 0;
	std::vector<std::unique_ptr<racc_b200::EnvReplica>> replicas;
	racc_b200::EnvReplica* on(int device) const { return racc_b200::replicaOn(replicas, device); }
};

// What the reference's example path tracer shades with (Renderer/SceneData.h), resident on the devices.
struct racc_cuda_shading {
	uint32_t vertexCount = 0, triangleCount = 0, materialCount = 0;
	std::vector<std::unique_ptr<racc_b200::ShadingReplica>> replicas;
	racc_b200::ShadingReplica* on(int device) const { return racc_b200::replicaOn(replicas, device); }
};

namespace racc_b200 {

// One traversal of `streams` on the calling thread's bound device (DEVICE streams) / dealt over its device set (HOST
// streams); capi.cu. deviceTotal: see TraceParams::totalPtr (single DEVICE stream only).
int traceImpl(racc_cuda_scene* s, racc_cuda_env* env, const racc_cuda_stream_desc* streams, uint32_t nstreams, void* cuda_stream,
              void* device_counters, bool fullCounters, const uint32_t* deviceTotal = nullptr, int gridCtasPerSm = 0);

bool sceneExceedsL2(const racc_cuda_scene* s);

// Per-thread, per-device scratch of the renderers (capi_render.cu); released by racc_cuda_thread_release.
void releaseRenderScratch();

// comm.cu
int commFrameReduce(racc_cuda_counters* totals, cudaStream_t stream);
int commAllGather(const void* send, void* recv, size_t bytesPerRank, cudaStream_t stream);
int commRanks(int* rank);
void commShutdown();

} // namespace racc_b200
