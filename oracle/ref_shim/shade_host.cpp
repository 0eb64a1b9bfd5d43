// shade_host.cpp -- TEST INFRASTRUCTURE ONLY: calls the reference's UNMODIFIED ReflectiveDiffuseMaterial::sample8
// (compiled from /root/reference/Renderer/Materials.cpp by oracle/Makefile, target `renderer`) for arrays of lanes, so
// that oracle_material_sample (oracle/racc_oracle.c) can be pinned against it (tests/test_render_oracle.py).
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
