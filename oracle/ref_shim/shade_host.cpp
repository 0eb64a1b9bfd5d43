// shade_host.cpp -- TEST INFRASTRUCTURE ONLY: calls the reference's UNMODIFIED ReflectiveDiffuseMaterial::sample8,
// Camera::lookAt, generateTileRays and generateTileLightPaths (compiled from /root/reference/Renderer/{Materials,Camera,
// LightPath}.cpp by oracle/Makefile, target `renderer`) so that the checker's material, camera and primary rays can be
// pinned against them (tests/test_render_oracle.py).
#include "Materials.h"

#include <cstdint>
#include <cstring>
#include <new>

// rnd / normal / wo / wi / color: count x 3 floats (array of structures); lanes are padded to a multiple of eight
extern "C" int ref_material_sample(const float* ke4, const float* rnd, const float* normal, const float* wo, uint32_t count, float* wi,
                                   float* color) {
	void* memory = _mm_malloc(sizeof(ReflectiveDiffuseMaterial), 64);
	ReflectiveDiffuseMaterial* material = new (memory) ReflectiveDiffuseMaterial(make_float3(ke4[0], ke4[1], ke4[2]), ke4[3]);
	ALIGNED(32) float3_8 r, n, uvt, o, i, c;
	std::memset(&uvt, 0, sizeof uvt);
	for (uint32_t base = 0; base < count; base += 8) {
		for (uint32_t l = 0; l < 8; ++l) {
			const uint32_t s = base + l < count ? base + l : count - 1;
			r.x.x[l] = rnd[3 * s]; r.y.x[l] = rnd[3 * s + 1]; r.z.x[l] = rnd[3 * s + 2];
			n.x.x[l] = normal[3 * s]; n.y.x[l] = normal[3 * s + 1]; n.z.x[l] = normal[3 * s + 2];
			o.x.x[l] = wo[3 * s]; o.y.x[l] = wo[3 * s + 1]; o.z.x[l] = wo[3 * s + 2];
		}
		unsigned transmitted = 0;
		material->sample8(&r, &n, &uvt, &o, &i, &c, &transmitted);
		if (transmitted) return -1;
		for (uint32_t l = 0; l < 8 && base + l < count; ++l) {
			const uint32_t s = base + l;
			wi[3 * s] = i.x.x[l]; wi[3 * s + 1] = i.y.x[l]; wi[3 * s + 2] = i.z.x[l];
			color[3 * s] = c.x.x[l]; color[3 * s + 1] = c.y.x[l]; color[3 * s + 2] = c.z.x[l];
		}
	}
	material->~ReflectiveDiffuseMaterial();
	_mm_free(memory);
	return 0;
}

// ---- the reference's camera, from /root/reference/Renderer/Camera.cpp (linked in unmodified) ----
#include "Camera.h"
#include "LightPath.h"

// camera12 out: origin, view, right, up of Camera::lookAt (Camera.cpp:13-25)
extern "C" void ref_camera_look_at(const float* origin, const float* target, const float* up, float fov, int width, int height, float* camera12) {
	Camera c = {};
	c.lookAt(make_float3(origin[0], origin[1], origin[2]), make_float3(target[0], target[1], target[2]), make_float3(up[0], up[1], up[2]), fov,
	         1e-3f, 1e+6f, width, height);
	const float3 v[4] = {c.origin, c.view, c.right, c.up};
	for (int k = 0; k < 4; ++k) { camera12[3 * k] = v[k].x; camera12[3 * k + 1] = v[k].y; camera12[3 * k + 2] = v[k].z; }
}

// generateTileRays (Camera.cpp:55-114) and generateTileLightPaths (LightPath.cpp:11-39) for one tile: tileSize^2 rays
// (8 floats each) and light paths (weight rgb + pixel index, 4 words each). libc rand() seeds the jitter, as in the reference.
extern "C" void ref_generate_tile(const float* camera12, unsigned tileX, unsigned tileY, unsigned tileSize, unsigned viewportWidth, float* rays8,
                                  float* lightPaths4) {
	Camera c = {};
	c.origin = make_float3(camera12[0], camera12[1], camera12[2]);
	c.view = make_float3(camera12[3], camera12[4], camera12[5]);
	c.right = make_float3(camera12[6], camera12[7], camera12[8]);
	c.up = make_float3(camera12[9], camera12[10], camera12[11]);
	const size_t n = (size_t)tileSize * tileSize;
	racc::Ray* rays = static_cast<racc::Ray*>(_mm_malloc(n * sizeof(racc::Ray), 64));
	LightPath* paths = static_cast<LightPath*>(_mm_malloc(n * sizeof(LightPath), 64));
	generateTileRays(rays, c, tileX, tileY, tileSize);
	generateTileLightPaths(paths, viewportWidth, tileX, tileY, tileSize);
	std::memcpy(rays8, rays, n * sizeof(racc::Ray));
	std::memcpy(lightPaths4, paths, n * sizeof(LightPath));
	_mm_free(rays);
	_mm_free(paths);
}
