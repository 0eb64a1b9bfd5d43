// pathtrace.cu -- device-side wavefront shading for the reference's example path tracer (SURVEY.md section 8f,
// rank 2): rays, hit records and path state never leave HBM between bounces.
//
// What is restated (one GPU thread per path instead of eight AVX lanes per host thread):
//   /root/reference/Renderer/Camera.cpp:55-114            generateTileRays   -> pathPrimaryKernel
//   /root/reference/Renderer/LightPath.cpp:11-39          generateTileLightPaths (weight 1, pixel) -> pathPrimaryKernel
//   /root/reference/Renderer/PathTracingRenderer.cpp:72-566  shade: who is shaded (:116-127), shading normal (:231-300),
//                                                         sample / weight / continuation / next ray (:376-466),
//                                                         radiance of escaping paths (:468-566)    -> pathShadeKernel
//   /root/reference/Renderer/Materials.cpp:11-151         ReflectiveDiffuseMaterial::sample8       -> materialSample
//
// Data in HBM per batch of R = width*height*spp_batch paths: two ray buffers (32 B/path), two path-state buffers
// (16 B: weight rgb + path index), one result buffer (16 B), one radiance buffer (16 B, written at most once per path:
// a path contributes only when it escapes to the light probe) -- 128 B per path. Compaction between bounces is one
// atomic per CTA; paths keep their arrival order inside a CTA, CTAs land in completion order. A batch is cut into lanes
// (contiguous path ranges) that advance bounce by bounce on their own CUDA streams, so that the tail of one lane's
// traversal launch is filled by the other lane's kernels (capi.cu: racc_cuda_path_trace). The image does not depend
// on any of that order: a ray's result does not depend on its neighbours, and every path owns its radiance slot, which
// pathAccumulateKernel adds to the framebuffer sample by sample in ascending order (so there are no float atomics and
// the framebuffer is bit-reproducible, and equal to oracle_path_trace's).
//
// Random numbers: the reference seeds an MWC generator from libc rand() per call (SimdRandom.h:20-56), which is not
// reproducible; here every draw is a counter-based hash of (pixel, sample, depth, seed). Its _mm256_rsqrt_ps /
// _mm256_rcp_ps approximations are exact 1/sqrt and 1/x here. Arithmetic is pinned like the traversal's (DESIGN.md
// section 3): explicit fmaf, no contraction, IEEE division and square root, FTZ.
#include "raygen.cuh"

namespace racc_b200 {
namespace {

constexpr int kShadeBlock = 256;

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

// paths [firstPath, firstPath + count) of a batch; path i is pixel i % pixels of sample sampleBase + i / pixels
__global__ void pathPrimaryKernel(CameraArgs cam, uint32_t width, uint32_t pixels, uint32_t sampleBase, uint32_t firstPath, uint32_t count,
                                  uint32_t seed, DevRay* rays, float4* states) {
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= count) return;
	const uint32_t i = firstPath + k;
	const uint32_t pixel = i % pixels, sample = sampleBase + i / pixels;
	rays[k] = primaryRay(cam, width, pixel, sample, seed);
	states[k] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(i));
}

struct ShadeArgs {
	const DevRay* rays;      // the wave just traced
	const float4* results;
	const float4* states;    // weight rgb, path index (within the batch) in .w
	uint32_t count;
	const uint32_t* countPtr; // non-null: the wave's size lives on the device (<= count); see TraceParams::totalPtr
	uint32_t depth, maxDepth; // this wave's bounce number; paths are extended while depth < maxDepth
	uint32_t seed, pixels, sampleBase;
	const uint32_t* indices;  // scene data of Renderer/SceneData.h
	const float4* normals;
	const float4* triangleNormals;
	const uint16_t* triangleMaterials;
	const float4* materials;
	uint32_t triangleCount, materialCount;
	DevRay* outRays;          // next wave, compacted
	float4* outStates;
	uint32_t* outCount;       // zeroed by the caller
	float4* radiance;         // per path of the batch, zeroed by the caller
};

__global__ void __launch_bounds__(kShadeBlock) pathShadeKernel(const ShadeArgs a) {
	__shared__ uint32_t warpCount[kShadeBlock / 32];
	__shared__ uint32_t ctaBase;
	// tiles of kShadeBlock rays dealt round-robin to the CTAs: the grid is sized on the host from an upper bound of the
	// wave's size, the size itself may still be on the device
	const uint32_t count = a.countPtr ? min(__ldg(a.countPtr), a.count) : a.count;
	for (uint32_t tile = blockIdx.x; (size_t)tile * kShadeBlock < count; tile += gridDim.x) {
	const uint32_t i = tile * kShadeBlock + threadIdx.x;
	bool go = false;
	DevRay next;
	float4 state = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	if (i < count) {
		const float4 res = a.results[i];
		state = a.states[i];
		const uint32_t tri = __float_as_uint(res.x);
		if (tri == 0xffffffffu) {
			// PathTracingRenderer.cpp:468-566: the light probe's radiance times the path weight
			a.radiance[__float_as_uint(state.w)] = make_float4(res.y * state.x, res.z * state.y, res.w * state.z, 0.0f);
		}
		else if (tri < a.triangleCount && a.depth < a.maxDepth) {
			const DevRay ray = a.rays[i];
			const float t = res.y, u = res.z, v = res.w;
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
			const float rd[3] = {ray.b.x, ray.b.y, ray.b.z};
			const float ro[3] = {ray.a.x, ray.a.y, ray.a.z};
			const float rdgn = fmaf(rd[2], gn[2], fmaf(rd[1], gn[1], rd[0] * gn[0]));
			const uint32_t sgn0 = __float_as_uint(rdgn) & 0x80000000u;
			float wo[3], pos[3];
#pragma unroll
			for (int k = 0; k < 3; ++k) {
				n[k] = xorSign(n[k] * fn, sgn0);
				wo[k] = -rd[k];
				pos[k] = fmaf(rd[k], t, ro[k]);
			}
			const uint32_t path = __float_as_uint(state.w);
			const uint32_t pixel = path % a.pixels, sample = a.sampleBase + path / a.pixels;
			uint32_t h = pcg(pixel ^ pcg(sample ^ pcg(a.seed ^ (0x9e3779b9u * (a.depth + 1u)))));
			float rnd[3];
			rnd[0] = unitFloat(h); h = pcg(h);
			rnd[1] = unitFloat(h); h = pcg(h);
			rnd[2] = unitFloat(h);
			float wi[3], color[3];
			materialSample(ke, rnd, n, wo, wi, color);
			state.x *= color[0]; state.y *= color[1]; state.z *= color[2];
			go = state.x > 0.01f || state.y > 0.01f || state.z > 0.01f;
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
		}
	}
	// compaction: one atomic per CTA, arrival order kept inside the CTA
	const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t ballot = __ballot_sync(0xffffffffu, go);
	if (lane == 0) warpCount[warp] = __popc(ballot);
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t total = 0;
		for (int wv = 0; wv < kShadeBlock / 32; ++wv) {
			const uint32_t c = warpCount[wv];
			warpCount[wv] = total;
			total += c;
		}
		ctaBase = total ? atomicAdd(a.outCount, total) : 0;
	}
	__syncthreads();
	if (go) {
		const uint32_t slot = ctaBase + warpCount[warp] + __popc(ballot & ((1u << lane) - 1u));
		a.outRays[slot] = next;
		a.outStates[slot] = state;
	}
	__syncthreads(); // warpCount / ctaBase are reused by the next tile
	}
}

// framebuffer[p] += radiance of sample 0, 1, ... of the batch, in that order (PathTracingRenderer.cpp:540-543 adds one
// sample per frame)
__global__ void pathAccumulateKernel(const float4* radiance, uint32_t pixels, uint32_t spp, float4* framebuffer) {
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= pixels) return;
	float4 acc = framebuffer[p];
	for (uint32_t s = 0; s < spp; ++s) {
		const float4 r = radiance[(size_t)s * pixels + p];
		acc.x += r.x; acc.y += r.y; acc.z += r.z;
	}
	framebuffer[p] = acc;
}

} // namespace

cudaError_t launchPathPrimary(const float* camera12, uint32_t width, uint32_t height, uint32_t sampleBase, uint32_t firstPath, uint32_t count,
                              uint32_t seed, DevRay* rays, float4* states, cudaStream_t stream, int* launches) {
	if (!count) return cudaSuccess;
	pathPrimaryKernel<<<(count + 255) / 256, 256, 0, stream>>>(cameraArgs(camera12), width, width * height, sampleBase, firstPath, count, seed,
	                                                           rays, states);
	if (launches) *launches += 1;
	return cudaGetLastError();
}

cudaError_t launchPathShade(const PathShadeParams& p, cudaStream_t stream, int* launches) {
	if (!p.count) return cudaSuccess;
	ShadeArgs a;
	a.rays = p.rays; a.results = p.results; a.states = p.states; a.count = p.count; a.countPtr = p.countPtr;
	a.depth = p.depth; a.maxDepth = p.maxDepth; a.seed = p.seed; a.pixels = p.pixels; a.sampleBase = p.sampleBase;
	a.indices = p.indices; a.normals = p.normals; a.triangleNormals = p.triangleNormals; a.triangleMaterials = p.triangleMaterials;
	a.materials = p.materials; a.triangleCount = p.triangleCount; a.materialCount = p.materialCount;
	a.outRays = p.outRays; a.outStates = p.outStates; a.outCount = p.outCount; a.radiance = p.radiance;
	uint32_t grid = (p.count + kShadeBlock - 1) / kShadeBlock;
	if (p.gridLimit && grid > p.gridLimit) grid = p.gridLimit;
	pathShadeKernel<<<grid, kShadeBlock, 0, stream>>>(a);
	if (launches) *launches += 1;
	return cudaGetLastError();
}

cudaError_t launchPathAccumulate(const float4* radiance, uint32_t pixels, uint32_t spp, float4* framebuffer, cudaStream_t stream,
                                 int* launches) {
	if (!pixels || !spp) return cudaSuccess;
	pathAccumulateKernel<<<(pixels + 255) / 256, 256, 0, stream>>>(radiance, pixels, spp, framebuffer);
	if (launches) *launches += 1;
	return cudaGetLastError();
}

} // namespace racc_b200
