// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or shipped with the product.
//
// Thin extern "C" driver around the UNMODIFIED reference sources, which are compiled from
// where they lie under /root/reference (see oracle/Makefile, target `ref`) into
// oracle/_ref/libracc_ref.so. It lets the tests obtain:
//   * the exact node / triangle-pair / remap byte images racc::createScene() would upload to
//     the reference's GPU back-end (RayAccelerator/Scene.cpp:183-357, Bvh2.cpp:772-907);
//   * the raw Bvh2 the reference builder produces (RayAccelerator/Bvh2.h:15-33);
//   * the reference's CPU light-probe lookup racc_internal::sample()
//     (RayAccelerator/Environment.h:27-82).
//   * the reference's CPU query path executeRayQueryCPU (RayAccelerator/Scene.cpp:374-484), run from its own
//     source: the AoS -> RTCRay8 transposes, the primID / tfar / u / v scatter, the light-probe lookup of misses
//     and the scalar tail are the reference's. What it calls -- rtcIntersect8 / rtcIntersect -- lives inside
//     Embree 2.7 (binary-only, macOS/Windows); ref_shim/mini_embree.cpp stands in for that binary.
#include "Scene.h"
#include "Bvh2.h"
#include "Context.h"
#include "Environment.h"
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <xmmintrin.h>
#include <pmmintrin.h>
#include <thread>
#include <vector>

namespace {
	struct FtzScope {
		unsigned saved;
		FtzScope() : saved(_mm_getcsr()) {
			// Same state racc::init() establishes (RayAccelerator.cpp:417-420).
			_MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
			_MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
		}
		~FtzScope() { _mm_setcsr(saved); }
	};

	racc::Context* fakeContext(bool gpu) {
		racc::Context* c = static_cast<racc::Context*>(_mm_malloc(sizeof(racc::Context), 64));
		memset(c, 0, sizeof(racc::Context));
		c->configuration.gpuContext = gpu ? reinterpret_cast<cl_context>(1) : 0;
		return c;
	}

	void* dup(const void* p, size_t n) {
		void* r = malloc(n ? n : 1);
		memcpy(r, p, n);
		return r;
	}
}

extern "C" {

struct ref_scene_images {
	void* nodes;  uint64_t nodes_bytes;   // 64 B per inner node
	void* pairs;  uint64_t pairs_bytes;   // 48 B per triangle pair (incl. padding pairs)
	void* remap;  uint64_t remap_bytes;   // 4 B per pair-triangle slot
};

int ref_build_scene(const float* verts4, uint32_t nverts, const uint32_t* indices, uint32_t nindices, ref_scene_images* out) {
	FtzScope ftz;
	racc::Vertex* v = static_cast<racc::Vertex*>(_mm_malloc(sizeof(racc::Vertex) * (size_t)nverts + 64, 64));
	memcpy(v, verts4, sizeof(racc::Vertex) * (size_t)nverts);
	racc::Context* ctx = fakeContext(true);
	racc::Scene* scene = racc::createScene(ctx, v, nverts, indices, nindices);
	int rc = -1;
	if (scene && scene->gpuNodes) {
		out->nodes = dup(scene->gpuNodes->data, scene->gpuNodes->size);
		out->nodes_bytes = scene->gpuNodes->size;
		out->pairs = dup(scene->gpuTriangles->data, scene->gpuTriangles->size);
		out->pairs_bytes = scene->gpuTriangles->size;
		out->remap = dup(scene->gpuTriangleIndices->data, scene->gpuTriangleIndices->size);
		out->remap_bytes = scene->gpuTriangleIndices->size;
		rc = 0;
	}
	if (scene)
		racc::destroy(scene);
	_mm_free(ctx);
	_mm_free(v);
	return rc;
}

void ref_free_scene_images(ref_scene_images* img) {
	free(img->nodes);
	free(img->pairs);
	free(img->remap);
	memset(img, 0, sizeof(*img));
}

// Raw builder output: nodes_out must hold 2*ntris Bvh2Node (48 B each), tris_out ntris uint32.
// Returns the node count, or -1.
int ref_build_bvh2(const float* verts4, uint32_t nverts, const uint32_t* indices, uint32_t ntris, void* nodes_out, uint32_t* tris_out) {
	FtzScope ftz;
	racc::Vertex* v = static_cast<racc::Vertex*>(_mm_malloc(sizeof(racc::Vertex) * (size_t)nverts + 64, 64));
	memcpy(v, verts4, sizeof(racc::Vertex) * (size_t)nverts);
	racc_internal::Bvh2* bvh = racc_internal::createBvh2(v, nverts, indices, ntris);
	int n = -1;
	if (bvh) {
		n = (int)bvh->nodeCount;
		memcpy(nodes_out, bvh->nodes, sizeof(racc_internal::Bvh2Node) * (size_t)n);
		memcpy(tris_out, bvh->triangles, sizeof(uint32_t) * (size_t)ntris);
		racc_internal::destroy(bvh);
	}
	_mm_free(v);
	return n;
}

// results[i] = what racc's CPU back-end returns for rays[i] (32 B racc::Ray -> 16 B racc::Result): the reference's own
// executeRayQueryCPU over [0, count), cut into slices of a multiple of 8 rays for `threads` host threads (the reference
// hands it slices of cpuTestBatch rays, RayAccelerator.cpp:158-244). rgba may be null (misses then keep r=g=b=0... the
// reference dereferences the environment unconditionally, so a 1x1 black probe is created for it).
int ref_cpu_query(const float* verts4, uint32_t nverts, const uint32_t* indices, uint32_t nindices, const float* rgba, uint32_t width,
                  uint32_t height, const void* rays, uint32_t count, void* results, int threads) {
	FtzScope ftz;
	racc::Vertex* v = static_cast<racc::Vertex*>(_mm_malloc(sizeof(racc::Vertex) * (size_t)nverts + 64, 64));
	memcpy(v, verts4, sizeof(racc::Vertex) * (size_t)nverts);
	racc::Context* ctx = fakeContext(false);
	racc::Scene* scene = racc::createScene(ctx, v, nverts, indices, nindices);
	const float black[4] = {0, 0, 0, 0};
	racc::Environment* env = rgba ? racc::createEnvironment(ctx, reinterpret_cast<const racc::Color*>(rgba), width, height)
	                              : racc::createEnvironment(ctx, reinterpret_cast<const racc::Color*>(black), 1, 1);
	int rc = -1;
	if (scene && env) {
		// the reference's streams are 4 KiB-aligned slabs (RayAccelerator.cpp:532-568); executeRayQueryCPU uses aligned loads
		racc::Ray* r = static_cast<racc::Ray*>(_mm_malloc(sizeof(racc::Ray) * (size_t)count + 64, 4096));
		racc::Result* o = static_cast<racc::Result*>(_mm_malloc(sizeof(racc::Result) * (size_t)count + 64, 4096));
		memcpy(r, rays, sizeof(racc::Ray) * (size_t)count);
		memset(o, 0, sizeof(racc::Result) * (size_t)count);
		racc::RayStream stream = {};
		stream.index = 0;
		stream.count = count;
		stream.rays = r;
		stream.results = o;
		if (threads < 1) threads = (int)std::thread::hardware_concurrency();
		if (threads < 1) threads = 1;
		const uint32_t slice = ((count / (uint32_t)threads + 8) / 8) * 8;
		std::vector<std::thread> pool;
		for (uint32_t begin = 0; begin < count; begin += slice) {
			const uint32_t end = begin + slice < count ? begin + slice : count;
			pool.emplace_back([=, &stream] {
				FtzScope inner; // worker threads of the reference run with FTZ + DAZ (Threading.h:77-79)
				racc_internal::executeRayQueryCPU(scene, &stream, env, begin, end);
			});
		}
		for (std::thread& t : pool) t.join();
		memcpy(results, o, sizeof(racc::Result) * (size_t)count);
		_mm_free(r);
		_mm_free(o);
		rc = 0;
	}
	if (env) racc::destroy(env);
	if (scene) racc::destroy(scene);
	_mm_free(ctx);
	_mm_free(v);
	return rc;
}

// out4[i] = sample(environment, dirs4[i]); dirs4 is n x 4 floats (xyz used).
int ref_env_sample(const float* rgba, uint32_t width, uint32_t height, const float* dirs4, uint32_t n, float* out4) {
	FtzScope ftz;
	racc::Context* ctx = fakeContext(false);
	racc::Environment* env = racc::createEnvironment(ctx, reinterpret_cast<const racc::Color*>(rgba), width, height);
	if (!env) {
		_mm_free(ctx);
		return -1;
	}
	for (uint32_t i = 0; i < n; ++i) {
		__m128 d = _mm_loadu_ps(dirs4 + 4 * (size_t)i);
		_mm_storeu_ps(out4 + 4 * (size_t)i, racc_internal::sample(env, d));
	}
	racc::destroy(env);
	_mm_free(ctx);
	return 0;
}

} // extern "C"
