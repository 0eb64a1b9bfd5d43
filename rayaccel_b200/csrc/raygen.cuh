// raygen.cuh -- device helpers shared by raygen.cu (synthetic streams) and pathtrace.cu (device-side renderer):
// the counter-based hash both draw their random numbers from, and the primary ray of the reference's camera
// model (/root/reference/Renderer/Camera.cpp:55-114).
#pragma once

#include "engine.h"

namespace racc_b200 {

__device__ __forceinline__ uint32_t pcg(uint32_t v) {
	uint32_t s = v * 747796405u + 2891336453u;
	uint32_t w = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
	return (w >> 22u) ^ w;
}

__device__ __forceinline__ float unitFloat(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }

struct CameraArgs { float origin[3], view[3], right[3], up[3]; };

inline CameraArgs cameraArgs(const float* camera12) {
	CameraArgs cam;
	for (int k = 0; k < 3; ++k) {
		cam.origin[k] = camera12[k];
		cam.view[k] = camera12[3 + k];
		cam.right[k] = camera12[6 + k];
		cam.up[k] = camera12[9 + k];
	}
	return cam;
}

// Camera.cpp:62-84 for one pixel sample: jittered position on the image plane (pixel centre when seed == 0),
// direction = normalize(view + up*py + right*px), minT 0, maxT 1e6
__device__ __forceinline__ DevRay primaryRay(const CameraArgs& cam, uint32_t width, uint32_t pixel, uint32_t sample, uint32_t seed) {
	const uint32_t x = pixel % width, y = pixel / width;
	float jx = 0.5f, jy = 0.5f;
	if (seed) {
		const uint32_t h = pcg(pixel ^ pcg(sample ^ pcg(seed)));
		jx = unitFloat(h);
		jy = unitFloat(pcg(h));
	}
	const float px = (float)x + jx, py = (float)y + jy;
	const float dx = fmaf(cam.right[0], px, fmaf(cam.up[0], py, cam.view[0]));
	const float dy = fmaf(cam.right[1], px, fmaf(cam.up[1], py, cam.view[1]));
	const float dz = fmaf(cam.right[2], px, fmaf(cam.up[2], py, cam.view[2]));
	const float scale = 1.0f / sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
	DevRay r;
	r.a = make_float4(cam.origin[0], cam.origin[1], cam.origin[2], 0.0f);
	r.b = make_float4(dx * scale, dy * scale, dz * scale, 1e+6f);
	return r;
}

} // namespace racc_b200
