// plumbing_client.cpp -- TEST INFRASTRUCTURE ONLY. A small client of include/RayAccelerator.h, written
// the way the reference's example renderers use the API (side arrays indexed by
// stream.index * rayStreamSize + i, tiles handed out by an atomic counter, bounce rays appended in
// shade()), that checks the scheduler contract of SURVEY.md section 8b and compares every result
// that came back through the callbacks with the CPU oracle run directly on the same rays.
//
// Built two ways by tests/test_api_plumbing.py:
//   CPU : racc_api.cpp + tests/harness/fake_capi.cpp + scene_build.cpp + liboracle   (no GPU needed)
//   GPU : linked against rayaccel_b200/libracc_b200.so (the product), checker still liboracle
// Prints one JSON object; exit code 0 iff every check passed.
#include <RayAccelerator.h>

#include "../../oracle/racc_oracle.h"
#include "../../rayaccel_b200/csrc/scene_build.h"

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifdef RACC_FAKE_CAPI
extern "C" {
uint64_t fake_capi_device_calls(int d);
uint64_t fake_capi_device_rays(int d);
uint64_t fake_capi_reduces(void);
uint64_t fake_capi_thread_releases(void);
}
#endif

namespace {

struct Tag {
	uint32_t id;    // primary ray id
	uint32_t depth; // 0 primary, 1 bounce
};

struct Client {
	racc::ContextInfo info;
	uint32_t totalPrimary = 0, tile = 0, tiles = 0;
	std::atomic<uint32_t> nextTile{0};
	std::vector<Tag> tags; // rayStreamCount * rayStreamSize
	std::vector<racc::Ray> primaryRays, bounceRays;
	std::vector<racc::Result> primaryResults, bounceResults;
	std::vector<uint8_t> bounceSeen, primarySeen;
	std::vector<std::atomic<int>> threadBusy;
	std::atomic<uint64_t> violations{0}, shaded{0}, spawned{0};
	float lo[3], hi[3];

	explicit Client(size_t threads) : threadBusy(threads) {}
};

uint32_t hash32(uint32_t v) {
	v ^= v >> 16; v *= 0x7feb352du; v ^= v >> 15; v *= 0x846ca68bu; v ^= v >> 16;
	return v;
}
float unit(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }

racc::Ray primaryRay(const Client& c, uint32_t id) {
	racc::Ray r;
	const uint32_t h = hash32(id * 3 + 1);
	// origins above the height field, directions mostly downwards, some grazing or upwards (misses)
	r.origin[0] = c.lo[0] + (c.hi[0] - c.lo[0]) * unit(hash32(h));
	r.origin[1] = c.hi[1] + 5.0f;
	r.origin[2] = c.lo[2] + (c.hi[2] - c.lo[2]) * unit(hash32(h + 1));
	float d[3] = {unit(hash32(h + 2)) - 0.5f, -unit(hash32(h + 3)) + 0.15f, unit(hash32(h + 4)) - 0.5f};
	const float l = 1.0f / std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
	for (int k = 0; k < 3; ++k) r.dir[k] = d[k] * l;
	r.minT = 0.0f;
	r.maxT = 1e6f;
	return r;
}

racc::Ray bounceRay(const racc::Ray& in, const racc::Result& res, uint32_t id) {
	racc::Ray r;
	for (int k = 0; k < 3; ++k) r.origin[k] = in.origin[k] + in.dir[k] * res.hit.t;
	r.origin[1] += 1e-3f;
	const uint32_t h = hash32(id * 7 + 5);
	float d[3] = {unit(hash32(h)) - 0.5f, unit(hash32(h + 1)) * 0.5f + 0.01f, unit(hash32(h + 2)) - 0.5f};
	const float l = 1.0f / std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
	for (int k = 0; k < 3; ++k) r.dir[k] = d[k] * l;
	r.minT = 1e-3f;
	r.maxT = 1e6f;
	return r;
}

struct BusyGuard {
	Client& c; unsigned thread;
	BusyGuard(Client& c_, unsigned t) : c(c_), thread(t) {
		if (t >= c.threadBusy.size() || c.threadBusy[t].fetch_add(1) != 0) c.violations.fetch_add(1);
	}
	~BusyGuard() { if (thread < c.threadBusy.size()) c.threadBusy[thread].fetch_sub(1); }
};

bool spawnCb(void* data, unsigned thread, racc::RayStream* out) {
	Client& c = *static_cast<Client*>(data);
	BusyGuard guard(c, thread);
	if (out->index >= c.info.rayStreamCount) c.violations.fetch_add(1);
	const uint32_t t = c.nextTile.fetch_add(1);
	if (t >= c.tiles) return false;
	const uint32_t begin = t * c.tile, end = begin + c.tile < c.totalPrimary ? begin + c.tile : c.totalPrimary;
	if (out->count + (end - begin) > c.info.rayStreamSize) { c.violations.fetch_add(1); return false; }
	Tag* tags = c.tags.data() + (size_t)out->index * c.info.rayStreamSize;
	for (uint32_t id = begin; id < end; ++id) {
		out->rays[out->count] = c.primaryRays[id];
		tags[out->count] = Tag{id, 0};
		++out->count;
	}
	c.spawned.fetch_add(end - begin);
	return t + 1 < c.tiles;
}

void shadeCb(void* data, unsigned thread, const racc::RayStream* in, unsigned start, unsigned end, racc::RayStream* out) {
	Client& c = *static_cast<Client*>(data);
	BusyGuard guard(c, thread);
	if (in->index >= c.info.rayStreamCount || out->index >= c.info.rayStreamCount || in->index == out->index || end > in->count || start >= end)
		c.violations.fetch_add(1);
	const Tag* inTags = c.tags.data() + (size_t)in->index * c.info.rayStreamSize;
	Tag* outTags = c.tags.data() + (size_t)out->index * c.info.rayStreamSize;
	for (unsigned i = start; i < end; ++i) {
		const Tag tag = inTags[i];
		const racc::Result& res = in->results[i];
		if (tag.depth == 0) {
			if (c.primarySeen[tag.id]++) c.violations.fetch_add(1);
			c.primaryResults[tag.id] = res;
			if (res.triangle != racc::invalidTriangle) {
				if (out->count >= c.info.rayStreamSize) { c.violations.fetch_add(1); continue; }
				const racc::Ray r = bounceRay(in->rays[i], res, tag.id);
				c.bounceRays[tag.id] = r;
				out->rays[out->count] = r;
				outTags[out->count] = Tag{tag.id, 1};
				++out->count;
			}
		}
		else {
			if (c.bounceSeen[tag.id]++) c.violations.fetch_add(1);
			c.bounceResults[tag.id] = res;
		}
	}
	c.shaded.fetch_add(end - start);
}

} // namespace

int main(int argc, char** argv) {
	uint32_t totalPrimary = 200000, grid = 96, frames = 2, devices = 1;
	racc::init();
	racc::Configuration cfg = racc::defaultConfiguration(racc::cudaDevice(0));
	cfg.cpuThreads = 4;
	cfg.gpuSubmissionThreads = 2;
	cfg.maxRaysPerSpawn = 4096;
	cfg.cpuShadeBatch = 2048;
	cfg.rayStreamBatchSize = 11 * 1024;
	cfg.maxRaysInFlight = 96 * 1024;
	for (int i = 1; i + 1 < argc; i += 2) {
		const long v = atol(argv[i + 1]);
		if (!strcmp(argv[i], "--rays")) totalPrimary = (uint32_t)v;
		else if (!strcmp(argv[i], "--grid")) grid = (uint32_t)v;
		else if (!strcmp(argv[i], "--frames")) frames = (uint32_t)v;
		else if (!strcmp(argv[i], "--threads")) cfg.cpuThreads = (uint8_t)v;
		else if (!strcmp(argv[i], "--submitters")) cfg.gpuSubmissionThreads = (uint8_t)v;
		else if (!strcmp(argv[i], "--spawn")) cfg.maxRaysPerSpawn = (uint16_t)v;
		else if (!strcmp(argv[i], "--shade")) cfg.cpuShadeBatch = (uint16_t)v;
		else if (!strcmp(argv[i], "--batch")) cfg.rayStreamBatchSize = (uint16_t)v;
		else if (!strcmp(argv[i], "--inflight")) cfg.maxRaysInFlight = (uint32_t)v;
		else if (!strcmp(argv[i], "--devices")) devices = (uint32_t)v;
	}
	if (devices > 1) cfg.gpuContext = racc::cudaDevices(0, (int)devices); // one context over several GPUs

	// a bumpy height field of 2*grid*grid triangles
	std::vector<racc::Vertex> verts((size_t)(grid + 1) * (grid + 1));
	std::vector<uint32_t> indices;
	for (uint32_t z = 0; z <= grid; ++z)
		for (uint32_t x = 0; x <= grid; ++x) {
			racc::Vertex& v = verts[(size_t)z * (grid + 1) + x];
			v.x = (float)x; v.z = (float)z; v.w = 1.0f;
			v.y = 3.0f * unit(hash32(z * 7919u + x)) + 2.0f * std::sin(0.2f * x) * std::cos(0.15f * z);
		}
	for (uint32_t z = 0; z < grid; ++z)
		for (uint32_t x = 0; x < grid; ++x) {
			const uint32_t a = z * (grid + 1) + x, b = a + 1, c2 = a + grid + 1, d = c2 + 1;
			const uint32_t quad[6] = {a, c2, b, b, c2, d};
			indices.insert(indices.end(), quad, quad + 6);
		}
	std::vector<racc::Color> envPixels(32 * 16);
	for (size_t i = 0; i < envPixels.size(); ++i)
		envPixels[i] = racc::Color{unit(hash32((uint32_t)i)), unit(hash32((uint32_t)i + 99)), unit(hash32((uint32_t)i + 777)), 1.0f};

	racc::Context* context = racc::createContext(cfg);
	if (!context) { printf("{\"ok\": false, \"error\": \"createContext failed\"}\n"); return 2; }
	racc::Scene* scene = racc::createScene(context, verts.data(), (unsigned)verts.size(), indices.data(), (unsigned)indices.size());
	racc::Environment* env = racc::createEnvironment(context, envPixels.data(), 32, 16);
	if (!scene || !env) { printf("{\"ok\": false, \"error\": \"createScene/createEnvironment failed\"}\n"); return 2; }

	Client client(cfg.cpuThreads);
	client.info = racc::info(context);
	client.totalPrimary = totalPrimary;
	client.tile = cfg.maxRaysPerSpawn;
	client.tiles = (totalPrimary + client.tile - 1) / client.tile;
	client.tags.resize((size_t)client.info.rayStreamCount * client.info.rayStreamSize);
	client.lo[0] = 0; client.lo[1] = -2; client.lo[2] = 0;
	client.hi[0] = (float)grid; client.hi[1] = 5; client.hi[2] = (float)grid;
	client.primaryRays.resize(totalPrimary);
	for (uint32_t id = 0; id < totalPrimary; ++id) client.primaryRays[id] = primaryRay(client, id);

	// checker: the oracle on the engine's host-built images
	racc_b200::SceneImages images;
	const char* why = "";
	if (!racc_b200::buildSceneImages(&verts[0].x, (uint32_t)verts.size(), indices.data(), (uint32_t)indices.size(), 0, &images, &why)) {
		printf("{\"ok\": false, \"error\": \"%s\"}\n", why);
		return 2;
	}
	oracle_scene sc{};
	sc.nodes = reinterpret_cast<const float*>(images.nodes.data()); sc.node_count = (uint32_t)images.nodes.size();
	sc.pairs = reinterpret_cast<const float*>(images.pairs.data()); sc.pair_count = (uint32_t)images.pairs.size();
	sc.remap = images.remap.data(); sc.remap_count = (uint32_t)images.remap.size();
	sc.env = &envPixels[0].r; sc.env_width = 32; sc.env_height = 16;

	uint64_t mismatches = 0, missing = 0, tracedTotal = 0, expectedTotal = 0, hitsTotal = 0;
	for (uint32_t f = 0; f < frames; ++f) {
		client.nextTile = 0;
		client.primaryResults.assign(totalPrimary, racc::Result{});
		client.bounceResults.assign(totalPrimary, racc::Result{});
		client.bounceRays.assign(totalPrimary, racc::Ray{});
		client.primarySeen.assign(totalPrimary, 0);
		client.bounceSeen.assign(totalPrimary, 0);
		const racc::RenderCallbacks cb = {&client, spawnCb, shadeCb};
		const racc::Stats stats = racc::render(context, scene, env, cb);
		tracedTotal += stats.raysTraced;

		std::vector<oracle_result> want(totalPrimary), wantBounce;
		oracle_traverse(&sc, reinterpret_cast<const oracle_ray*>(client.primaryRays.data()), totalPrimary, want.data(), nullptr, 0);
		std::vector<racc::Ray> bounce;
		std::vector<uint32_t> bounceIds;
		for (uint32_t id = 0; id < totalPrimary; ++id) {
			if (!client.primarySeen[id]) { ++missing; continue; }
			if (memcmp(&want[id], &client.primaryResults[id], 16)) ++mismatches;
			if (want[id].triangle != ORACLE_INVALID_TRIANGLE) {
				racc::Result r; memcpy(&r, &want[id], 16);
				bounce.push_back(bounceRay(client.primaryRays[id], r, id));
				bounceIds.push_back(id);
			}
		}
		wantBounce.resize(bounce.size());
		oracle_traverse(&sc, reinterpret_cast<const oracle_ray*>(bounce.data()), (uint32_t)bounce.size(), wantBounce.data(), nullptr, 0);
		for (size_t k = 0; k < bounceIds.size(); ++k) {
			const uint32_t id = bounceIds[k];
			if (!client.bounceSeen[id]) { ++missing; continue; }
			if (memcmp(&bounce[k], &client.bounceRays[id], 32) || memcmp(&wantBounce[k], &client.bounceResults[id], 16)) ++mismatches;
		}
		hitsTotal += bounceIds.size();
		expectedTotal += totalPrimary + bounceIds.size();
	}

	racc::destroy(env);
	racc::destroy(scene);
	racc::destroy(context);
	racc::deinit();

	const bool ok = !mismatches && !missing && !client.violations.load() && tracedTotal == expectedTotal;
#ifdef RACC_FAKE_CAPI
	// the test double's bookkeeping: what every pretend device traced, frame reductions, submitter teardown
	printf("{\"fake\": true, \"reduces\": %llu, \"thread_releases\": %llu, \"device_rays\": [", (unsigned long long)fake_capi_reduces(),
	       (unsigned long long)fake_capi_thread_releases());
	for (uint32_t d = 0; d < devices; ++d) printf("%s%llu", d ? ", " : "", (unsigned long long)fake_capi_device_rays((int)d));
	printf("], \"device_calls\": [");
	for (uint32_t d = 0; d < devices; ++d) printf("%s%llu", d ? ", " : "", (unsigned long long)fake_capi_device_calls((int)d));
	printf("]}\n");
#endif
	printf("{\"ok\": %s, \"frames\": %u, \"rays_traced\": %llu, \"rays_expected\": %llu, \"primary_hits\": %llu, \"mismatches\": %llu, "
	       "\"missing\": %llu, \"violations\": %llu, \"stream_count\": %u, \"stream_size\": %u, \"thread_count\": %u}\n",
	       ok ? "true" : "false", frames, (unsigned long long)tracedTotal, (unsigned long long)expectedTotal, (unsigned long long)hitsTotal,
	       (unsigned long long)mismatches, (unsigned long long)missing, (unsigned long long)client.violations.load(),
	       client.info.rayStreamCount, client.info.rayStreamSize, client.info.threadCount);
	return ok ? 0 : 1;
}
