// raygen.cu -- synthetic ray-stream generators for the benchmark and tests.
// NOT on the hot path and not part of the parity contract: whatever bytes these kernels write are
// handed unchanged to both the CUDA traversal and the CPU oracle.
//
//   primary rays : the camera model of /root/reference/Renderer/Camera.cpp:13-25,55-114
//   bounce rays  : the ray construction of /root/reference/Renderer/PathTracingRenderer.cpp:405-422
//                  (origin = hit + 1e-4 * n_g, minT 1e-3, maxT 1e6) with a cosine-hemisphere
//                  direction about the flipped geometric normal, i.e. a diffuse path-tracer bounce.
#include "raygen.cuh"

namespace racc_b200 {
namespace {

__global__ void primaryKernel(CameraArgs cam, uint32_t width, uint32_t height, uint32_t spp, uint32_t seed, DevRay* rays) {
	const uint64_t total = (uint64_t)width * height * spp;
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= total) return;
	const uint32_t pixel = (uint32_t)(i % ((uint64_t)width * height));
	const uint32_t sample = (uint32_t)(i / ((uint64_t)width * height));
	rays[i] = primaryRay(cam, width, pixel, sample, seed);
}

constexpr int kTile = 1024; // rays per compaction tile (256 threads x 4)

__device__ __forceinline__ bool isHit(const float4* results, uint32_t i, uint32_t count) {
	return i < count && __float_as_uint(__ldg(&results[i]).x) != 0xffffffffu;
}

__global__ void bounceCountKernel(const float4* results, uint32_t count, uint32_t* tileCounts) {
	__shared__ uint32_t warpSums[8];
	const uint32_t base = blockIdx.x * kTile + threadIdx.x * 4;
	uint32_t n = 0;
	for (int k = 0; k < 4; ++k) n += isHit(results, base + k, count);
	for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
	if ((threadIdx.x & 31) == 0) warpSums[threadIdx.x >> 5] = n;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t s = 0;
		for (int w = 0; w < 8; ++w) s += warpSums[w];
		tileCounts[blockIdx.x] = s;
	}
}

// single-CTA exclusive scan over the tile counts (a few thousand to ~130 K entries)
__global__ void bounceScanKernel(uint32_t* tileCounts, uint32_t tiles, uint32_t* outCount) {
	__shared__ uint32_t partial[1024];
	__shared__ uint32_t carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (uint32_t start = 0; start < tiles; start += 1024) {
		const uint32_t i = start + threadIdx.x;
		const uint32_t v = i < tiles ? tileCounts[i] : 0;
		partial[threadIdx.x] = v;
		__syncthreads();
		for (uint32_t o = 1; o < 1024; o <<= 1) {
			const uint32_t add = threadIdx.x >= o ? partial[threadIdx.x - o] : 0;
			__syncthreads();
			partial[threadIdx.x] += add;
			__syncthreads();
		}
		const uint32_t inclusive = partial[threadIdx.x];
		if (i < tiles) tileCounts[i] = carry + inclusive - v;
		__syncthreads();
		if (threadIdx.x == 1023) carry += inclusive;
		__syncthreads();
	}
	if (threadIdx.x == 0) *outCount = carry;
}

__global__ void bounceWriteKernel(const float4* verts, const uint32_t* indices, const DevRay* rays, const float4* results,
                                  uint32_t count, uint32_t seed, const uint32_t* tileOffsets, DevRay* outRays) {
	__shared__ uint32_t warpSums[8];
	const uint32_t base = blockIdx.x * kTile + threadIdx.x * 4;
	bool hit[4];
	uint32_t n = 0;
	for (int k = 0; k < 4; ++k) { hit[k] = isHit(results, base + k, count); n += hit[k]; }
	// exclusive scan of n across the CTA
	uint32_t incl = n;
	const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= (unsigned)o) incl += t;
	}
	if (lane == 31) warpSums[warp] = incl;
	__syncthreads();
	uint32_t offset = tileOffsets[blockIdx.x] + incl - n;
	for (unsigned w = 0; w < warp; ++w) offset += warpSums[w];

	for (int k = 0; k < 4; ++k) {
		if (!hit[k]) continue;
		const uint32_t i = base + k;
		const float4 res = __ldg(&results[i]);
		const DevRay in = rays[i];
		const uint32_t tri = __float_as_uint(res.x);
		const float t = res.y;
		const float4 p0 = __ldg(&verts[__ldg(&indices[3 * (size_t)tri])]);
		const float4 p1 = __ldg(&verts[__ldg(&indices[3 * (size_t)tri + 1])]);
		const float4 p2 = __ldg(&verts[__ldg(&indices[3 * (size_t)tri + 2])]);
		const float ax = p1.x - p0.x, ay = p1.y - p0.y, az = p1.z - p0.z;
		const float bx = p2.x - p0.x, by = p2.y - p0.y, bz = p2.z - p0.z;
		float nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
		const float nl = 1.0f / sqrtf(fmaxf(nx * nx + ny * ny + nz * nz, 1e-30f));
		nx *= nl; ny *= nl; nz *= nl;
		if (nx * in.b.x + ny * in.b.y + nz * in.b.z > 0.0f) { nx = -nx; ny = -ny; nz = -nz; }
		const float hx = fmaf(in.b.x, t, in.a.x), hy = fmaf(in.b.y, t, in.a.y), hz = fmaf(in.b.z, t, in.a.z);
		// cosine-weighted hemisphere about n
		const uint32_t h0 = pcg(i ^ pcg(seed));
		const float u1 = unitFloat(h0), u2 = unitFloat(pcg(h0));
		const float rr = sqrtf(u1), phi = 6.28318530718f * u2;
		const float lx = rr * cosf(phi), ly = rr * sinf(phi), lz = sqrtf(fmaxf(0.0f, 1.0f - u1));
		// orthonormal basis (branchless, Duff et al.)
		const float sgn = copysignf(1.0f, nz);
		const float a = -1.0f / (sgn + nz);
		const float b = nx * ny * a;
		const float t1x = 1.0f + sgn * nx * nx * a, t1y = sgn * b, t1z = -sgn * nx;
		const float t2x = b, t2y = sgn + ny * ny * a, t2z = -ny;
		float dx = lx * t1x + ly * t2x + lz * nx;
		float dy = lx * t1y + ly * t2y + lz * ny;
		float dz = lx * t1z + ly * t2z + lz * nz;
		const float dl = 1.0f / sqrtf(fmaxf(dx * dx + dy * dy + dz * dz, 1e-30f));
		DevRay o;
		o.a = make_float4(fmaf(nx, 1e-4f, hx), fmaf(ny, 1e-4f, hy), fmaf(nz, 1e-4f, hz), 1e-3f);
		o.b = make_float4(dx * dl, dy * dl, dz * dl, 1e+6f);
		outRays[offset++] = o;
	}
}

} // namespace

cudaError_t launchGeneratePrimary(const float* camera12, uint32_t width, uint32_t height, uint32_t spp, uint32_t seed,
                                  DevRay* rays, cudaStream_t stream, int* launches) {
	const CameraArgs cam = cameraArgs(camera12);
	const uint64_t total = (uint64_t)width * height * spp;
	if (!total) return cudaSuccess;
	const unsigned grid = (unsigned)((total + 255) / 256);
	primaryKernel<<<grid, 256, 0, stream>>>(cam, width, height, spp, seed, rays);
	if (launches) *launches += 1;
	return cudaGetLastError();
}

size_t bounceScratchWords(uint32_t count) { return ((size_t)count + kTile - 1) / kTile + 1; }

cudaError_t launchGenerateBounce(const float4* verts, const uint32_t* indices, const DevRay* rays, const float4* results,
                                 uint32_t count, uint32_t seed, DevRay* outRays, uint32_t* outCount, uint32_t* scratch,
                                 cudaStream_t stream, int* launches) {
	if (!count) return cudaMemsetAsync(outCount, 0, sizeof(uint32_t), stream);
	const uint32_t tiles = (count + kTile - 1) / kTile;
	bounceCountKernel<<<tiles, 256, 0, stream>>>(results, count, scratch);
	bounceScanKernel<<<1, 1024, 0, stream>>>(scratch, tiles, outCount);
	bounceWriteKernel<<<tiles, 256, 0, stream>>>(verts, indices, rays, results, count, seed, scratch, outRays);
	if (launches) *launches += 3;
	return cudaGetLastError();
}

} // namespace racc_b200
