// RayAccelerator.h -- the public C++ API of the B200 engine.
//
// Same namespace, type names, field order/widths and function signatures as the reference's public
// header (/root/reference/RayAccelerator/RayAccelerator.h:25-116), so that clients written against
// the reference -- its example path tracer and Whitted renderer included -- compile and link
// without edits. The reference header refuses to compile on Linux (:21-23) and pulls in OpenCL
// only for the one `cl_context gpuContext` field (:33); here `cl_context` is an opaque token that
// names a CUDA device (see racc::cudaDevice below). Everything behind these functions is
// rayaccel_b200/csrc/racc_api.cpp, a client of the C-ABI in racc_b200.h.
//
// Differences a client can observe:
//   * there is no CPU intersection back-end: createContext() fails (returns null after printing
//     "RayAccelerator: ...") when configuration.gpuContext is null or no CUDA device is present;
//     `allowCpuTracing` and `cpuTestBatch` are accepted and ignored;
//   * ray streams live in pinned host memory and are moved to the GPU in aggregated launches
//     (several streams per launch) instead of one zero-copy launch per stream.
#ifndef RACC_B200_RAYACCELERATOR_H
#define RACC_B200_RAYACCELERATOR_H

#include <stdint.h>
#include <immintrin.h>

// the reference's clients rely on these arriving through this header (its OpenCL include did that)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#if defined(_WIN32)
#define RACC_ALIGNED(n) __declspec(align(n))
#else
#define RACC_ALIGNED(n) __attribute__((aligned(n)))
#endif

#ifndef RACC_B200_NO_CL_CONTEXT_TYPEDEF
typedef struct _cl_context* cl_context; // opaque: non-null = "use the GPU engine", see cudaDevice()
#endif

namespace racc {
	static const uint32_t invalidTriangle = ~(uint32_t)0; // Result.triangle of a miss

	struct Context;
	struct Scene;
	struct Environment;

	struct Configuration {
		cl_context gpuContext;         // token from cudaDevice(); null is rejected (no CPU back-end)
		bool allowCpuTracing;          // ignored: intersection always runs on the GPU
		uint8_t cpuThreads;            // host threads running the spawn/shade callbacks
		uint8_t gpuSubmissionThreads;  // host threads that gather ready streams and launch them
		uint32_t maxRaysInFlight;      // cap on rays alive in the system
		uint16_t maxRaysPerSpawn;      // most rays one spawn callback may append
		uint16_t cpuTestBatch;         // ignored (was: CPU intersection slice)
		uint16_t cpuShadeBatch;        // rays handed to one shade callback
		uint16_t rayStreamBatchSize;   // fill level at which a stream is queued for intersection
	};

	struct ContextInfo {
		uint16_t threadCount;     // callbacks receive thread in [0, threadCount)
		uint16_t rayStreamCount;  // RayStream.index < rayStreamCount
		uint32_t rayStreamSize;   // RayStream.count <= rayStreamSize
		uint32_t maxRaysInFlight;
	};

	struct RACC_ALIGNED(16) Vertex {
		float x, y, z, w;
	};

	struct RACC_ALIGNED(16) Color {
		float r, g, b, a;
	};

	struct RACC_ALIGNED(32) Ray {
		float origin[3];
		float minT;
		float dir[3];
		float maxT;
	};

	struct RACC_ALIGNED(16) Result {
		uint32_t triangle; // original triangle index, or invalidTriangle
		union {
			struct {
				float t, u, v; // u, v weight the triangle's 2nd and 3rd vertex
			} hit;
			struct {
				float r, g, b; // light-probe radiance along the ray
			} miss;
		};
	};

	struct RayStream {
		uint32_t index;
		uint32_t count;
		Ray* rays;
		Result* results; // index-parallel to rays, valid in shade()
	};

	struct Stats {
		uint64_t raysTraced;
	};

	struct RenderCallbacks {
		void* data;
		bool (*spawn)(void* data, unsigned thread, RayStream* output);
		void (*shade)(void* data, unsigned thread, const RayStream* input, unsigned start, unsigned end, RayStream* output);
	};

	// Process-wide set-up. Sets flush-to-zero / denormals-are-zero on the calling thread, as the reference
	// does for every thread it starts (RayAccelerator.cpp:417-420, Threading.h:77-79): the shaders of a client
	// inherit that. No device is touched before createContext().
	void init();

	void deinit();

	// Defaults sized for one B200, not for the reference's integrated GPU
	// (RayAccelerator.cpp:429-446): 65 536-ray streams, 2 M rays in flight, 2 submitter threads.
	Configuration defaultConfiguration(cl_context gpuContext);

	// Allocates the stream slab in pinned host memory and starts the callback and submitter threads
	// (RayAccelerator.cpp:448-736). Returns null after printing "RayAccelerator: ..." on failure.
	Context* createContext(Configuration configuration);

	// Joins the threads and releases the slab (RayAccelerator.cpp:761-788). Scenes and environments created
	// from the context must be destroyed first.
	void destroy(Context* context);

	// Sizes a client needs for its per-thread and per-stream side arrays (RayAccelerator.cpp:790-797).
	ContextInfo info(Context* context);

	// Builds the scene (full-sweep SAH BVH2, triangle pairs) and uploads it; vertices and indices are copied
	// (Scene.cpp:183-357). vertices must be 16-byte aligned, indexCount a multiple of 3. Null on failure.
	Scene* createScene(Context* context, const Vertex* vertices, unsigned vertexCount, const uint32_t* indices, unsigned indexCount);

	void destroy(Scene* scene);

	// An angular-map light probe of width x height RGBA32F texels, copied (Environment.cpp:13-60). Misses
	// return its bilinearly filtered radiance in Result.miss.
	Environment* createEnvironment(Context* context, const Color* colors, unsigned width, unsigned height);

	void destroy(Environment* environment);

	// One frame: runs spawn() until it returns false and shade() on every tested stream until no ray is left,
	// then returns (RayAccelerator.cpp:738-759). Callbacks run concurrently, without any library lock held, each
	// with a `thread` value no other running callback has. spawn appends at most maxRaysPerSpawn rays to
	// `output`; shade reads input->rays/results[start, end) and appends at most end - start rays to `output`,
	// which may already hold rays. One render() per context at a time.
	Stats render(Context* context, Scene* scene, Environment* environment, RenderCallbacks callbacks);

	// --- additions (not in the reference) -------------------------------------------------------

	// The token to put in Configuration.gpuContext: selects CUDA device `ordinal` of this process.
	// Stands in for the OpenCL context the reference's main.cpp creates (Renderer/main.cpp:68-115).
	inline cl_context cudaDevice(int ordinal = 0) {
		return reinterpret_cast<cl_context>(static_cast<uintptr_t>(ordinal) + 1);
	}

	// A token for `count` consecutive CUDA devices starting at `firstOrdinal` (at most 255 each): the context then drives
	// all of them -- scene and environment replicated on each, gpuSubmissionThreads submitters per device, ready ray
	// streams dealt over the devices, Stats.raysTraced summed over them with NCCL at the end of every frame. Ray
	// streams shard by index and need nothing from another GPU, so this is the reference's gpuSubmissionThreads idea
	// (RayAccelerator.cpp:711-717) taken from "several queues of one device" to "several devices".
	inline cl_context cudaDevices(int firstOrdinal, int count) {
		return reinterpret_cast<cl_context>((static_cast<uintptr_t>(firstOrdinal) + 1) | (static_cast<uintptr_t>(count) << 8));
	}
}

#endif
