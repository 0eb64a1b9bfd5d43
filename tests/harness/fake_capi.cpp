// fake_capi.cpp -- TEST DOUBLE of the C-ABI (include/racc_b200.h) for CPU-only plumbing tests.
//
// TEST INFRASTRUCTURE ONLY. Linked -- instead of libracc_b200.so -- with rayaccel_b200/csrc/
// racc_api.cpp so that the host scheduler behind RayAccelerator.h can be exercised on a machine
// without a GPU: scenes are built by the engine's host-only builder (scene_build.cpp, no CUDA) and
// "traced" by the CPU oracle (oracle/racc_oracle.c). Nothing in the product links or loads this
// file; the product library has no such switch.
#include "../../include/racc_b200.h"
#include "../../oracle/racc_oracle.h"
#include "../../rayaccel_b200/csrc/scene_build.h"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace racc_b200;

struct racc_cuda_scene {
	SceneImages images;
};

struct racc_cuda_env {
	std::vector<float> texels;
	uint32_t width, height;
};

static std::atomic<uint64_t> g_calls{0};
static std::atomic<uint64_t> g_streams{0};
static std::atomic<uint64_t> g_rays{0};
static std::atomic<uint64_t> g_maxRaysPerCall{0};
// several pretend devices (FAKE_CAPI_DEVICES=n): which thread is bound to which, what each traced, per-device frame records
static constexpr int kFakeDevices = 16;
static std::atomic<uint64_t> g_deviceCalls[kFakeDevices];
static std::atomic<uint64_t> g_deviceRays[kFakeDevices];
static std::atomic<uint64_t> g_frameRays[kFakeDevices];
static std::atomic<uint64_t> g_reduces{0};
static std::atomic<uint64_t> g_threadReleases{0};
static thread_local int t_bound = 0;
static thread_local int t_setSize = 1;
static thread_local int t_set[kFakeDevices] = {0};
static int fakeDeviceCount() { const char* v = getenv("FAKE_CAPI_DEVICES"); const int n = v ? atoi(v) : 1; return n < 1 ? 1 : (n > kFakeDevices ? kFakeDevices : n); }

extern "C" {

// test-only introspection
uint64_t fake_capi_trace_calls(void) { return g_calls.load(); }
uint64_t fake_capi_streams(void) { return g_streams.load(); }
uint64_t fake_capi_rays(void) { return g_rays.load(); }
uint64_t fake_capi_max_rays_per_call(void) { return g_maxRaysPerCall.load(); }

uint64_t fake_capi_device_calls(int d) { return g_deviceCalls[d].load(); }
uint64_t fake_capi_device_rays(int d) { return g_deviceRays[d].load(); }
uint64_t fake_capi_reduces(void) { return g_reduces.load(); }
uint64_t fake_capi_thread_releases(void) { return g_threadReleases.load(); }

int racc_cuda_init(const int* devices, int n) {
	if (!devices || n <= 0) { t_bound = 0; t_setSize = 1; t_set[0] = 0; return 0; }
	for (int k = 0; k < n; ++k) {
		if (devices[k] < 0 || devices[k] >= fakeDeviceCount()) return -1;
		t_set[k] = devices[k];
	}
	t_setSize = n;
	t_bound = devices[0];
	return 0;
}
int racc_cuda_device_count(void) { return fakeDeviceCount(); }
void racc_cuda_thread_release(void) { g_threadReleases.fetch_add(1); }
int racc_cuda_frame_reduce(racc_cuda_counters* totals, void*) {
	uint64_t rays = 0;
	for (int k = 0; k < t_setSize; ++k) rays += g_frameRays[t_set[k]].exchange(0);
	if (totals) { totals->rays = rays; totals->hits = 0; totals->inner_nodes = 0; totals->pairs_tested = 0; }
	g_reduces.fetch_add(1);
	return 0;
}
int racc_cuda_abi_version(void) { return RACC_CUDA_ABI_VERSION; }
const char* racc_cuda_last_error(void) { return "fake C-ABI"; }

racc_cuda_scene* racc_cuda_scene_create(const float* verts4, uint32_t nverts, const uint32_t* indices, uint32_t nindices) {
	racc_cuda_scene* s = new racc_cuda_scene();
	const char* why = "";
	if (!buildSceneImages(verts4, nverts, indices, nindices, 0, &s->images, &why)) {
		delete s;
		return nullptr;
	}
	return s;
}

void racc_cuda_scene_destroy(racc_cuda_scene* s) { delete s; }

racc_cuda_env* racc_cuda_env_create(const float* rgba, uint32_t width, uint32_t height) {
	racc_cuda_env* e = new racc_cuda_env();
	e->texels.assign(rgba, rgba + (size_t)width * height * 4);
	e->width = width;
	e->height = height;
	return e;
}

void racc_cuda_env_destroy(racc_cuda_env* e) { delete e; }

void* racc_cuda_host_alloc(size_t bytes) {
	void* p = nullptr;
	return posix_memalign(&p, 4096, bytes ? bytes : 1) ? nullptr : p;
}

void racc_cuda_host_free(void* p) { free(p); }

void* racc_cuda_stream_create(void) { return malloc(1); }
void racc_cuda_stream_destroy(void* s) { free(s); }

int racc_cuda_trace(racc_cuda_scene* s, racc_cuda_env* env, const racc_cuda_stream_desc* streams, uint32_t nstreams, void*) {
	if (!s) return -1;
	oracle_scene sc{};
	sc.nodes = reinterpret_cast<const float*>(s->images.nodes.data());
	sc.node_count = (uint32_t)s->images.nodes.size();
	sc.pairs = reinterpret_cast<const float*>(s->images.pairs.data());
	sc.pair_count = (uint32_t)s->images.pairs.size();
	sc.remap = s->images.remap.data();
	sc.remap_count = (uint32_t)s->images.remap.size();
	if (env) {
		sc.env = env->texels.data();
		sc.env_width = env->width;
		sc.env_height = env->height;
	}
	uint64_t rays = 0;
	for (uint32_t i = 0; i < nstreams; ++i) {
		if (!streams[i].count) continue;
		if (oracle_traverse(&sc, static_cast<const oracle_ray*>(streams[i].rays), streams[i].count,
		                    static_cast<oracle_result*>(streams[i].results), nullptr, 2))
			return -1;
		rays += streams[i].count;
	}
	g_calls.fetch_add(1);
	g_deviceCalls[t_bound].fetch_add(1);
	g_deviceRays[t_bound].fetch_add(rays);
	g_frameRays[t_bound].fetch_add(rays);
	g_streams.fetch_add(nstreams);
	g_rays.fetch_add(rays);
	uint64_t prev = g_maxRaysPerCall.load();
	while (rays > prev && !g_maxRaysPerCall.compare_exchange_weak(prev, rays)) {}
	return 0;
}

int racc_cuda_sync(void*) { return 0; }

} // extern "C"
