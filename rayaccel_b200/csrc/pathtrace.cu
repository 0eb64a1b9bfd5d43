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
#include "pathshade.cuh"

namespace racc_b200 {
namespace {

constexpr int kShadeBlock = 256;

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
	ShadeScene scene;
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
		else if (tri < a.scene.triangleCount && a.depth < a.maxDepth) {
			const DevRay ray = a.rays[i];
			const float rd[3] = {ray.b.x, ray.b.y, ray.b.z};
			const float ro[3] = {ray.a.x, ray.a.y, ray.a.z};
			const uint32_t path = __float_as_uint(state.w);
			float weight[3] = {state.x, state.y, state.z};
			go = shadeHit(a.scene, tri, res.y, res.z, res.w, ro, rd, path % a.pixels, a.sampleBase + path / a.pixels, a.seed, a.depth, weight, next);
			state.x = weight[0]; state.y = weight[1]; state.z = weight[2];
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
	a.scene.indices = p.indices; a.scene.normals = p.normals; a.scene.triangleNormals = p.triangleNormals; a.scene.triangleMaterials = p.triangleMaterials;
	a.scene.materials = p.materials; a.scene.triangleCount = p.triangleCount; a.scene.materialCount = p.materialCount;
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
