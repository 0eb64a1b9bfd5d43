// pathshade.cuh -- the example path tracer's shading of one hit, shared by the wavefront shading kernel (pathtrace.cu)
// and the streamed path tracer (pathstream.cu): both extend a path with exactly these operations, so both produce the
// framebuffer of oracle_path_trace. What is restated, and how the arithmetic is pinned: see pathtrace.cu.
#pragma once

#include "raygen.cuh"

namespace racc_b200 {
namespace {

// scene data of Renderer/SceneData.h on the device
struct ShadeScene {
	const uint32_t* indices;
	const float4* normals;
	const float4* triangleNormals;
	const uint16_t* triangleMaterials;
	const float4* materials;  // {r, g, b, eta} per material
	uint32_t triangleCount, materialCount;
};

__device__ __forceinline__ float xorSign(float x, uint32_t signBit) { return __uint_as_float(__float_as_uint(x) ^ signBit); }

// Materials.cpp:11-22: parabola through sin(2 pi x), x in [0,1]
__device__ __forceinline__ float sinApprox(float x) {
	const float y = fmaf(-16.0f, x, 8.0f);
	const bool gt = x >= 0.5f;
	float xy = x * y;
	if (gt) xy = -xy;
	return xy + (gt ? y : 0.0f);
}

// Materials.cpp:24-28
__device__ __forceinline__ float cosApprox(float x) {
	const float y = x - 0.75f;
	x = (__float_as_uint(y) & 0x80000000u) ? x + 0.25f : y;
	return sinApprox(x);
}

// Materials.cpp:39-151, one lane. ke = {r, g, b, eta}
__device__ __forceinline__ void materialSample(const float4 ke, const float rnd[3], const float n[3], const float wo[3], float wi[3],
                                               float color[3]) {
	const float nx = n[0], ny = n[1], nz = n[2];
	const float eta = ke.w;
	// reflection vector and fresnel term
	float cosi = fmaf(nz, wo[2], fmaf(ny, wo[1], nx * wo[0]));
	cosi = cosi > 0.0f ? cosi : 0.0f;
	const float c2 = 2.0f * cosi;
	const float rx = fmaf(c2, nx, -wo[0]), ry = fmaf(c2, ny, -wo[1]), rz = fmaf(c2, nz, -wo[2]);
	const float cosi2m1 = fmaf(cosi, cosi, -1.0f);
	const float eta2 = eta * eta;
	const float k = fmaf(eta2, cosi2m1, 1.0f);
	const float cost = sqrtf(k);
	const float rper = fmaf(eta, cosi, -cost) * (1.0f / fmaf(eta, cosi, cost));
	const float rpar = -(fmaf(eta, cost, -cosi) * (1.0f / fmaf(eta, cost, cosi)));
	float fresnel = 0.5f * fmaf(rpar, rpar, rper * rper);
	if (__float_as_uint(k) & 0x80000000u) fresnel = 1.0f;
	// diffuse direction: cosine-weighted about n in the basis (u, v, n)
	const bool wide = !(fabsf(nx) <= 0.1f);
	float ux = wide ? -nz : 0.0f, uy = wide ? 0.0f : -nz, uz = wide ? nx : ny;
	const float fb = 1.0f / sqrtf(fmaf(uz, uz, fmaf(uy, uy, ux * ux)));
	ux *= fb; uy *= fb; uz *= fb;
	const float vx = fmaf(ny, uz, -(nz * uy)), vy = fmaf(nz, ux, -(nx * uz)), vz = fmaf(nx, uy, -(ny * ux));
	const float sinx = sinApprox(rnd[0]), cosx = cosApprox(rnd[0]);
	const float r2s = sqrtf(rnd[1]);
	const float sq = sqrtf(1.0f - rnd[1]);
	float dx = fmaf(nx, sq, fmaf(ux, cosx, vx * sinx) * r2s);
	float dy = fmaf(ny, sq, fmaf(uy, cosx, vy * sinx) * r2s);
	float dz = fmaf(nz, sq, fmaf(uz, cosx, vz * sinx) * r2s);
	const float fd = 1.0f / sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
	dx *= fd; dy *= fd; dz *= fd;
	// reflection with probability 3 F / (3 F + r + g + b), else diffuse; the weight keeps the estimator unbiased
	const float s0 = fresnel * 3.0f;
	const float s1 = ke.z + (ke.x + ke.y);
	const float sum = s0 + s1;
	const float uniform = rnd[2] * sum;
	const bool diffuse = uniform >= s0;
	wi[0] = diffuse ? dx : rx; wi[1] = diffuse ? dy : ry; wi[2] = diffuse ? dz : rz;
	const float r = diffuse ? ke.x : fresnel, g = diffuse ? ke.y : fresnel, b = diffuse ? ke.z : fresnel;
	const float scale = sum * (1.0f / (b + (r + g)));
	color[0] = r * scale; color[1] = g * scale; color[2] = b * scale;
}

// One hit of path `path` (pixel, sample) at bounce `depth`: PathTracingRenderer.cpp:231-300 (shading normal), :376-466
// (sample, weight, continuation, next ray). ro / rd: the ray as it was submitted; t, u, v, tri: its Result. weight is
// multiplied by the sample's colour. Returns whether the path goes on, with `next` its next ray.
__device__ __forceinline__ bool shadeHit(const ShadeScene& a, uint32_t tri, float t, float u, float v, const float ro[3], const float rd[3],
                                         uint32_t pixel, uint32_t sample, uint32_t seed, uint32_t depth, float weight[3], DevRay& next) {
	const uint32_t i0 = __ldg(&a.indices[3 * (size_t)tri]), i1 = __ldg(&a.indices[3 * (size_t)tri + 1]), i2 = __ldg(&a.indices[3 * (size_t)tri + 2]);
	const float4 n0 = __ldg(&a.normals[i0]), n1 = __ldg(&a.normals[i1]), n2 = __ldg(&a.normals[i2]);
	const float4 gn4 = __ldg(&a.triangleNormals[tri]);
	uint32_t m = __ldg(&a.triangleMaterials[tri]);
	if (m >= a.materialCount) m = 0;
	const float4 ke = __ldg(&a.materials[m]);
	const float w = 1.0f - (u + v);
	float n[3] = {fmaf(n2.x, v, fmaf(n1.x, u, n0.x * w)), fmaf(n2.y, v, fmaf(n1.y, u, n0.y * w)), fmaf(n2.z, v, fmaf(n1.z, u, n0.z * w))};
	const float fn = 1.0f / sqrtf(fmaf(n[2], n[2], fmaf(n[1], n[1], n[0] * n[0])));
	const float gn[3] = {gn4.x, gn4.y, gn4.z};
	const float rdgn = fmaf(rd[2], gn[2], fmaf(rd[1], gn[1], rd[0] * gn[0]));
	const uint32_t sgn0 = __float_as_uint(rdgn) & 0x80000000u;
	float wo[3], pos[3];
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		n[k] = xorSign(n[k] * fn, sgn0);
		wo[k] = -rd[k];
		pos[k] = fmaf(rd[k], t, ro[k]);
	}
	uint32_t h = pcg(pixel ^ pcg(sample ^ pcg(seed ^ (0x9e3779b9u * (depth + 1u)))));
	float rnd[3];
	rnd[0] = unitFloat(h); h = pcg(h);
	rnd[1] = unitFloat(h); h = pcg(h);
	rnd[2] = unitFloat(h);
	float wi[3], color[3];
	materialSample(ke, rnd, n, wo, wi, color);
	weight[0] *= color[0]; weight[1] *= color[1]; weight[2] *= color[2];
	bool go = weight[0] > 0.01f || weight[1] > 0.01f || weight[2] > 0.01f;
	const float sgn1 = fmaf(wi[2], gn[2], fmaf(wi[1], gn[1], wi[0] * gn[0]));
	go = go && ((__float_as_uint(sgn1) ^ sgn0) >> 31) != 0; // leaves on the side it arrived from
	const uint32_t flip = __float_as_uint(sgn1) & 0x80000000u;
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		pos[k] = fmaf(xorSign(gn[k], flip), 1e-4f, pos[k]);
		go = go && pos[k] == pos[k] && wi[k] == wi[k];
	}
	next.a = make_float4(pos[0], pos[1], pos[2], 1e-3f);
	next.b = make_float4(wi[0], wi[1], wi[2], 1e+6f);
	return go;
}

} // namespace
} // namespace racc_b200
