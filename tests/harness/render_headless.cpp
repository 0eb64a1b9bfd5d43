// render_headless.cpp -- TEST INFRASTRUCTURE ONLY. Headless stand-in for the reference's GLUT
// application (Renderer/main.cpp, out of scope: needs GLUT, CoreAudio/Win32 and an Intel OpenCL
// device). It drives the reference's UNMODIFIED example renderers -- Renderer/{TiledRenderer,
// PathTracingRenderer,WhittedRenderer,Camera,LightPath,Materials}.cpp, compiled from where they lie
// under /root/reference by oracle/Makefile (target `renderer`) -- through include/RayAccelerator.h,
// which is the proof that those clients "link unchanged" against this engine.
//
// Scene loading restates the container layout of Renderer/main.cpp:117-191; the throughput figure
// is the reference's own: Stats.raysTraced / microseconds per render() (main.cpp:208-231).
//
//   racc_render_{gpu,cpu} [--whitted] [--width W --height H] [--frames N] [--depth D]
//                         [--threads T] [--scene path] [--out image.ppm] [--dump framebuffer.f32]
// --dump writes the raw float4 framebuffer (radiance sums over the frames, width*height*4 floats).
// Prints one JSON line. `_gpu` links libracc_b200.so; `_cpu` links tests/harness/fake_capi.cpp
// (oracle-backed, for plumbing checks on machines without a GPU -- BASELINE.json configs[0]).
#include "Camera.h"
#include "Materials.h"
#include "PathTracingRenderer.h"
#include "SceneData.h"
#include "WhittedRenderer.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

static TiledRenderer* g_renderer;

static bool spawnCb(void*, unsigned thread, racc::RayStream* out) { return g_renderer->spawnPrimary(thread, out); }
static void shadeCb(void*, unsigned thread, const racc::RayStream* in, unsigned start, unsigned end, racc::RayStream* out) {
	g_renderer->shade(thread, in, start, end, out);
}

template <class T>
static T* alignedArray(size_t n) { return static_cast<T*>(_mm_malloc(n * sizeof(T) + 64, 64)); }

int main(int argc, char** argv) {
	bool whitted = false;
	int width = 0, height = 0, frames = 4, depth = -1, threads = 0, device = 0, devices = 1;
	std::string scenePath = "data/battlefield.bin", outPath, dumpPath;
	for (int i = 1; i < argc; ++i) {
		auto next = [&]() { return i + 1 < argc ? argv[++i] : "0"; };
		if (!strcmp(argv[i], "--whitted")) whitted = true;
		else if (!strcmp(argv[i], "--width")) width = atoi(next());
		else if (!strcmp(argv[i], "--height")) height = atoi(next());
		else if (!strcmp(argv[i], "--frames")) frames = atoi(next());
		else if (!strcmp(argv[i], "--depth")) depth = atoi(next());
		else if (!strcmp(argv[i], "--threads")) threads = atoi(next());
		else if (!strcmp(argv[i], "--device")) device = atoi(next());
		else if (!strcmp(argv[i], "--devices")) devices = atoi(next());
		else if (!strcmp(argv[i], "--scene")) scenePath = next();
		else if (!strcmp(argv[i], "--out")) outPath = next();
		else if (!strcmp(argv[i], "--dump")) dumpPath = next();
		else { fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
	}

	FILE* f = fopen(scenePath.c_str(), "rb");
	if (!f) { fprintf(stderr, "cannot open %s\n", scenePath.c_str()); return 2; }
#pragma pack(push, 1)
	struct Header {
		uint32_t maxDepth, vertexCount, triangleCount;
		uint16_t viewportWidth, viewportHeight, envWidth, envHeight;
		float origin[3], target[3], up[3], fov;
	} h;
#pragma pack(pop)
	static_assert(sizeof(Header) == 60, "battlefield.bin header is 60 bytes");
	if (fread(&h, sizeof(h), 1, f) != 1) return 2;

	SceneData sd = {};
	sd.maxDepth = (uint16_t)(depth >= 0 ? depth : (whitted ? 8 : (int)h.maxDepth));
	sd.vertexCount = h.vertexCount;
	sd.triangleCount = h.triangleCount;
	sd.viewportWidth = (uint16_t)(width > 0 ? width : h.viewportWidth);
	sd.viewportHeight = (uint16_t)(height > 0 ? height : h.viewportHeight);
	sd.indices = alignedArray<uint32_t>((size_t)h.triangleCount * 3);
	sd.triangleMaterials = alignedArray<uint16_t>(h.triangleCount);
	sd.triangleNormals = alignedArray<float4>(h.triangleCount);
	racc::Vertex* vertices = alignedArray<racc::Vertex>(h.vertexCount);
	sd.normals = alignedArray<float4>(h.vertexCount);
	sd.texcoords = alignedArray<float2>(h.vertexCount);
	racc::Color* envPixels = alignedArray<racc::Color>((size_t)h.envWidth * h.envHeight);
	bool ok = true;
	ok &= fread(sd.indices, sizeof(uint32_t) * 3, h.triangleCount, f) == h.triangleCount;
	ok &= fread(sd.triangleMaterials, sizeof(uint16_t), h.triangleCount, f) == h.triangleCount;
	ok &= fread(sd.triangleNormals, sizeof(float4), h.triangleCount, f) == h.triangleCount;
	ok &= fread(vertices, sizeof(racc::Vertex), h.vertexCount, f) == h.vertexCount;
	ok &= fread(sd.normals, sizeof(float4), h.vertexCount, f) == h.vertexCount;
	ok &= fread(sd.texcoords, sizeof(float2), h.vertexCount, f) == h.vertexCount;
	ok &= fread(envPixels, sizeof(racc::Color), (size_t)h.envWidth * h.envHeight, f) == (size_t)h.envWidth * h.envHeight;
	fclose(f);
	if (!ok) { fprintf(stderr, "%s: truncated scene file\n", scenePath.c_str()); return 2; }

	// the four materials main.cpp:160-165 assigns to battlefield
	const float albedo[4] = {0.8f, 0.1f, 0.6f, 0.3f};
	const float eta[4] = {1.0f / 1.4f, 1.0f / 1.4f, 1.0f / 1.2f, 1.0f / 1.2f};
	sd.materials = new Material*[4];
	for (int m = 0; m < 4; ++m)
		sd.materials[m] = new (_mm_malloc(sizeof(ReflectiveDiffuseMaterial), 64)) ReflectiveDiffuseMaterial(make_float3(albedo[m]), eta[m]);

	Camera camera = {};
	camera.lookAt(make_float3(h.origin[0], h.origin[1], h.origin[2]), make_float3(h.target[0], h.target[1], h.target[2]),
	              make_float3(h.up[0], h.up[1], h.up[2]), h.fov, 1e-3f, 1e+6f, sd.viewportWidth, sd.viewportHeight);

	racc::init();
	racc::Configuration cfg = racc::defaultConfiguration(devices > 1 ? racc::cudaDevices(device, devices) : racc::cudaDevice(device));
	if (threads > 0) cfg.cpuThreads = (uint8_t)threads;
	racc::Context* context = racc::createContext(cfg);
	if (!context) return 3;
	racc::Scene* scene = racc::createScene(context, vertices, sd.vertexCount, sd.indices, sd.triangleCount * 3);
	racc::Environment* environment = racc::createEnvironment(context, envPixels, h.envWidth, h.envHeight);
	if (!scene || !environment) return 3;

	if (whitted) g_renderer = new (_mm_malloc(sizeof(WhittedRenderer), 64)) WhittedRenderer(context, camera, sd);
	else g_renderer = new (_mm_malloc(sizeof(PathTracingRenderer), 64)) PathTracingRenderer(context, camera, sd);

	const racc::RenderCallbacks callbacks = {nullptr, spawnCb, shadeCb};
	uint64_t rays = 0, firstFrameRays = 0;
	double seconds = 0, best = 0;
	for (int frame = 0; frame < frames; ++frame) {
		const auto t0 = std::chrono::steady_clock::now();
		const racc::Stats stats = racc::render(context, scene, environment, callbacks);
		g_renderer->endFrame();
		const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		if (frame == 0) firstFrameRays = stats.raysTraced;
		if (frame > 0 || frames == 1) { rays += stats.raysTraced; seconds += dt; } // frame 0 warms up
		best = std::max(best, stats.raysTraced / dt / 1e6);
	}

	// framebuffer statistics over the rendered tile grid (TiledRenderer floors to whole 128-px tiles)
	const unsigned tw = sd.viewportWidth / 128 * 128, th = sd.viewportHeight / 128 * 128;
	double sum = 0;
	size_t nonBlack = 0, notFinite = 0;
	for (unsigned y = 0; y < th; ++y)
		for (unsigned x = 0; x < tw; ++x) {
			const float4 p = g_renderer->frameBuffer[(size_t)y * sd.viewportWidth + x];
			const float l = p.x + p.y + p.z;
			if (!std::isfinite(l)) { ++notFinite; continue; }
			sum += l;
			nonBlack += l > 0;
		}
	if (!outPath.empty()) {
		FILE* o = fopen(outPath.c_str(), "wb");
		if (o) {
			fprintf(o, "P6\n%u %u\n255\n", (unsigned)sd.viewportWidth, (unsigned)sd.viewportHeight);
			const float scale = 255.0f / (float)frames; // DisplayBuffer's tonemap: 255/spp
			for (size_t i = 0; i < (size_t)sd.viewportWidth * sd.viewportHeight; ++i) {
				const float4 p = g_renderer->frameBuffer[i];
				const float c[3] = {p.x, p.y, p.z};
				for (int k = 0; k < 3; ++k) fputc((int)std::fmin(255.0f, std::fmax(0.0f, c[k] * scale)), o);
			}
			fclose(o);
		}
	}
	if (!dumpPath.empty()) {
		FILE* o = fopen(dumpPath.c_str(), "wb");
		if (o) {
			fwrite(g_renderer->frameBuffer, sizeof(float4), (size_t)sd.viewportWidth * sd.viewportHeight, o);
			fclose(o);
		}
	}
	const racc::ContextInfo info = racc::info(context);
	printf("{\"renderer\": \"%s\", \"width\": %u, \"height\": %u, \"rendered_width\": %u, \"rendered_height\": %u, \"frames\": %d, \"max_depth\": %u, "
	       "\"rays_first_frame\": %llu, \"rays_timed\": %llu, \"seconds_timed\": %.6f, \"mrps\": %.3f, \"mrps_best_frame\": %.3f, "
	       "\"mean_luminance\": %.6f, \"nonblack_fraction\": %.6f, \"not_finite\": %zu, \"callback_threads\": %u, \"stream_count\": %u, \"stream_size\": %u}\n",
	       whitted ? "whitted" : "path", (unsigned)sd.viewportWidth, (unsigned)sd.viewportHeight, tw, th, frames, (unsigned)sd.maxDepth,
	       (unsigned long long)firstFrameRays, (unsigned long long)rays, seconds, seconds > 0 ? rays / seconds / 1e6 : 0.0, best,
	       tw * th ? sum / ((double)tw * th * frames) : 0.0, tw * th ? (double)nonBlack / ((double)tw * th) : 0.0, notFinite,
	       (unsigned)info.threadCount, (unsigned)info.rayStreamCount, info.rayStreamSize);

	racc::destroy(environment);
	racc::destroy(scene);
	racc::destroy(context);
	racc::deinit();
	return 0;
}
