// whitted.cu -- device-side shading for the reference's second example client, the Whitted renderer
// (/root/reference/Renderer/WhittedRenderer.cpp:136-676); companion of pathtrace.cu, same data conventions.
//
// Per hit with depth < maxDepth (WhittedRenderer.cpp): shading normal interpolated and flipped to the side the ray
// arrived on (:231-300,349-357), direct light weight*0.3*max(n.L, 0) for the fixed direction L = (0.57, 0.57, 0.57)
// (:343-372), weight *= 0.3, and while a weight channel exceeds 0.01 (:404-413) a mirror reflection AND a refraction
// (eta 1/1.1 entering, 1.1 leaving; :415-437), each subject to a side test against the geometric normal (:440-444) and
// a NaN test (:465-472); both children carry the parent's weight. Misses add probe radiance * weight (:578-660).
//
// The reference keeps each pixel's ray tree depth-first in a linked list with a free list under a mutex (:19-134) to
// bound host memory; a GPU wants the opposite: the tree is walked breadth-first, one wave per depth, every hit writing
// up to two rays into the next wave (compaction: one atomic per CTA). Many rays of a wave belong to the same pixel, so
// radiance goes to per-pixel 32.32 fixed-point accumulators with integer atomics: integer addition is associative, so
// the image is bit-reproducible and equal to the depth-first checker's (oracle_whitted_trace), which floating-point
// atomics would not give. (The reference's own read-modify-write of the framebuffer from several threads is unsynchronised.)
#include "raygen.cuh"

namespace racc_b200 {
namespace {

constexpr int kBlock = 256;
constexpr float kFixedOne = 4294967296.0f;

__device__ __forceinline__ float xorSign(float x, uint32_t signBit) { return __uint_as_float(__float_as_uint(x) ^ signBit); }

__device__ __forceinline__ unsigned long long toFixed(float c) {
	if (!(c > 0.0f)) return 0ull; // also NaN
	if (c > 1048576.0f) c = 1048576.0f;
	return __float2ull_rn(c * kFixedOne);
}

__device__ __forceinline__ void addRadiance(unsigned long long* acc, uint32_t pixel, float r, float g, float b) {
	const unsigned long long qr = toFixed(r), qg = toFixed(g), qb = toFixed(b);
	unsigned long long* p = acc + 3 * (size_t)pixel;
	if (qr) atomicAdd(p, qr);
	if (qg) atomicAdd(p + 1, qg);
	if (qb) atomicAdd(p + 2, qb);
}

// paths [firstPath, firstPath + count) of a batch; the state carries the PIXEL (children of a path share it)
__global__ void whittedPrimaryKernel(CameraArgs cam, uint32_t width, uint32_t pixels, uint32_t sampleBase, uint32_t firstPath, uint32_t count,
                                     uint32_t seed, DevRay* rays, float4* states) {
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= count) return;
	const uint32_t i = firstPath + k;
	const uint32_t pixel = i % pixels, sample = sampleBase + i / pixels;
	rays[k] = primaryRay(cam, width, pixel, sample, seed);
	states[k] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(pixel));
}

struct WhittedArgs {
	const DevRay* rays;
	const float4* results;
	const float4* states; // weight rgb, pixel in .w
	uint32_t count, depth, maxDepth;
	const uint32_t* indices;
	const float4* normals;
	const float4* triangleNormals;
	uint32_t triangleCount;
	DevRay* outRays;      // capacity 2 * count
	float4* outStates;
	uint32_t* outCount;   // zeroed by the caller
	unsigned long long* accumulators; // 3 per pixel, 32.32 fixed point
};

// kCombine (Tuning::whittedCombine): the rays of a pixel's tree stay adjacent through the order-preserving compaction, so
// a warp holds runs of lanes with the same pixel. Each lane turns its contribution into fixed point as before; the runs
// are then summed with a segmented shuffle reduction and only the first lane of a run issues the (up to three) atomics.
// Integer sums are associative and cannot overflow here (32 terms < 2^52 each), so the accumulators receive the same
// totals bit for bit. Runs are told apart by a run number, not by the pixel, so two separate runs of one pixel inside a
// warp are simply two atomics.
template <bool kCombine>
__global__ void __launch_bounds__(kBlock) whittedShadeKernel(const WhittedArgs a) {
	__shared__ uint32_t warpBase[kBlock / 32];
	__shared__ uint32_t ctaBase;
	const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
	bool reflect = false, refract = false;
	DevRay rl, rr;
	float4 state = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	uint32_t runKey = 0xffffffffu;                // kCombine: the lane's pixel
	unsigned long long qr = 0ull, qg = 0ull, qb = 0ull; // kCombine: the lane's contribution, fixed point
	if (i < a.count) {
		const float4 res = a.results[i];
		state = a.states[i];
		const uint32_t pixel = __float_as_uint(state.w);
		const uint32_t tri = __float_as_uint(res.x);
		runKey = pixel;
		if (tri == 0xffffffffu) {
			if (kCombine) { qr = toFixed(res.y * state.x); qg = toFixed(res.z * state.y); qb = toFixed(res.w * state.z); }
			else addRadiance(a.accumulators, pixel, res.y * state.x, res.z * state.y, res.w * state.z);
		}
		else if (tri < a.triangleCount && a.depth < a.maxDepth) {
			const DevRay ray = a.rays[i];
			const float t = res.y, u = res.z, v = res.w;
			const uint32_t i0 = __ldg(&a.indices[3 * (size_t)tri]), i1 = __ldg(&a.indices[3 * (size_t)tri + 1]), i2 = __ldg(&a.indices[3 * (size_t)tri + 2]);
			const float4 n0 = __ldg(&a.normals[i0]), n1 = __ldg(&a.normals[i1]), n2 = __ldg(&a.normals[i2]);
			const float4 gn4 = __ldg(&a.triangleNormals[tri]);
			const float w = 1.0f - (u + v);
			float n[3] = {fmaf(n2.x, v, fmaf(n1.x, u, n0.x * w)), fmaf(n2.y, v, fmaf(n1.y, u, n0.y * w)), fmaf(n2.z, v, fmaf(n1.z, u, n0.z * w))};
			const float fn = 1.0f / sqrtf(fmaf(n[2], n[2], fmaf(n[1], n[1], n[0] * n[0])));
			const float gn[3] = {gn4.x, gn4.y, gn4.z};
			const float rd[3] = {ray.b.x, ray.b.y, ray.b.z};
			const float ro[3] = {ray.a.x, ray.a.y, ray.a.z};
			const float rdgn = fmaf(rd[2], gn[2], fmaf(rd[1], gn[1], rd[0] * gn[0]));
			const uint32_t sgn0 = __float_as_uint(rdgn) & 0x80000000u;
#pragma unroll
			for (int k = 0; k < 3; ++k) n[k] = xorSign(n[k] * fn, sgn0);
			// direct light
			float light = fmaf(n[2], 0.57f, fmaf(n[1], 0.57f, n[0] * 0.57f));
			light = light > 0.0f ? light : 0.0f;
			state.x *= 0.3f; state.y *= 0.3f; state.z *= 0.3f;
			if (kCombine) { qr = toFixed(state.x * light); qg = toFixed(state.y * light); qb = toFixed(state.z * light); }
			else addRadiance(a.accumulators, pixel, state.x * light, state.y * light, state.z * light);
			if (!(state.x <= 0.01f) || !(state.y <= 0.01f) || !(state.z <= 0.01f)) {
				// reflection and refraction
				const float ddn = fmaf(rd[2], n[2], fmaf(rd[1], n[1], rd[0] * n[0]));
				const float cosi = ddn * -2.0f;
				const float eta = sgn0 ? 1.1f : 1.0f / 1.1f;
				const float r = 1.0f - (eta * eta) * (1.0f - ddn * ddn);
				const float mu = fmaf(eta, ddn, sqrtf(r));
				float dl[3], dr[3], pos[3];
#pragma unroll
				for (int k = 0; k < 3; ++k) {
					dl[k] = fmaf(cosi, n[k], rd[k]);
					dr[k] = fmaf(eta, rd[k], -(mu * n[k]));
					pos[k] = fmaf(rd[k], t, ro[k]);
				}
				const float sl = fmaf(dl[2], gn[2], fmaf(dl[1], gn[1], dl[0] * gn[0]));
				const float sr = fmaf(dr[2], gn[2], fmaf(dr[1], gn[1], dr[0] * gn[0]));
				reflect = ((__float_as_uint(sl) ^ sgn0) >> 31) != 0;                 // back to the side it came from
				refract = r > 0.0f && ((__float_as_uint(sr) ^ sgn0) >> 31) == 0;     // through the surface
				float pl[3], pr[3];
#pragma unroll
				for (int k = 0; k < 3; ++k) {
					pl[k] = fmaf(xorSign(gn[k], __float_as_uint(sl) & 0x80000000u), 1e-4f, pos[k]);
					pr[k] = fmaf(xorSign(gn[k], __float_as_uint(sr) & 0x80000000u), 1e-4f, pos[k]);
					reflect = reflect && pl[k] == pl[k] && dl[k] == dl[k];
					refract = refract && pr[k] == pr[k] && dr[k] == dr[k];
				}
				rl.a = make_float4(pl[0], pl[1], pl[2], 1e-3f);
				rl.b = make_float4(dl[0], dl[1], dl[2], 1e+6f);
				rr.a = make_float4(pr[0], pr[1], pr[2], 1e-3f);
				rr.b = make_float4(dr[0], dr[1], dr[2], 1e+6f);
			}
		}
	}
	const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (kCombine) {
		// every lane of the warp is here (no early exits above); lanes past the end carry key 0xffffffff and zeros
		const uint32_t before = __shfl_up_sync(0xffffffffu, runKey, 1);
		const bool head = lane == 0 || before != runKey;
		const uint32_t run = __popc(__ballot_sync(0xffffffffu, head) & (0xffffffffu >> (31 - lane))); // heads at or below this lane
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t otherRun = __shfl_down_sync(0xffffffffu, run, o);
			const unsigned long long ur = __shfl_down_sync(0xffffffffu, qr, o);
			const unsigned long long ug = __shfl_down_sync(0xffffffffu, qg, o);
			const unsigned long long ub = __shfl_down_sync(0xffffffffu, qb, o);
			if (lane + o < 32 && otherRun == run) { qr += ur; qg += ug; qb += ub; }
		}
		if (head && runKey != 0xffffffffu) {
			unsigned long long* p = a.accumulators + 3 * (size_t)runKey;
			if (qr) atomicAdd(p, qr);
			if (qg) atomicAdd(p + 1, qg);
			if (qb) atomicAdd(p + 2, qb);
		}
	}
	// compaction of 0..2 rays per thread: warp scan, one atomic per CTA; reflection before refraction
	const uint32_t mine = (reflect ? 1u : 0u) + (refract ? 1u : 0u);
	uint32_t incl = mine;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= (unsigned)o) incl += up;
	}
	if (lane == 31) warpBase[warp] = incl;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t total = 0;
		for (int wv = 0; wv < kBlock / 32; ++wv) {
			const uint32_t c = warpBase[wv];
			warpBase[wv] = total;
			total += c;
		}
		ctaBase = total ? atomicAdd(a.outCount, total) : 0;
	}
	__syncthreads();
	uint32_t slot = ctaBase + warpBase[warp] + incl - mine;
	if (reflect) {
		a.outRays[slot] = rl;
		a.outStates[slot] = state;
		++slot;
	}
	if (refract) {
		a.outRays[slot] = rr;
		a.outStates[slot] = state;
	}
}

// framebuffer += the call's radiance sums (fixed point -> float)
__global__ void whittedFinishKernel(const unsigned long long* acc, uint32_t pixels, float4* framebuffer) {
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= pixels) return;
	float4 f = framebuffer[p];
	f.x += __ull2float_rn(acc[3 * (size_t)p]) * (1.0f / kFixedOne);
	f.y += __ull2float_rn(acc[3 * (size_t)p + 1]) * (1.0f / kFixedOne);
	f.z += __ull2float_rn(acc[3 * (size_t)p + 2]) * (1.0f / kFixedOne);
	framebuffer[p] = f;
}

} // namespace

cudaError_t launchWhittedPrimary(const float* camera12, uint32_t width, uint32_t height, uint32_t sampleBase, uint32_t firstPath, uint32_t count,
                                 uint32_t seed, DevRay* rays, float4* states, cudaStream_t stream, int* launches) {
	if (!count) return cudaSuccess;
	whittedPrimaryKernel<<<(count + 255) / 256, 256, 0, stream>>>(cameraArgs(camera12), width, width * height, sampleBase, firstPath, count, seed,
	                                                              rays, states);
	if (launches) *launches += 1;
	return cudaGetLastError();
}

cudaError_t launchWhittedShade(const WhittedShadeParams& p, cudaStream_t stream, int* launches) {
	if (!p.count) return cudaSuccess;
	WhittedArgs a;
	a.rays = p.rays; a.results = p.results; a.states = p.states; a.count = p.count; a.depth = p.depth; a.maxDepth = p.maxDepth;
	a.indices = p.indices; a.normals = p.normals; a.triangleNormals = p.triangleNormals; a.triangleCount = p.triangleCount;
	a.outRays = p.outRays; a.outStates = p.outStates; a.outCount = p.outCount; a.accumulators = p.accumulators;
	if (p.combine) whittedShadeKernel<true><<<(p.count + kBlock - 1) / kBlock, kBlock, 0, stream>>>(a);
	else whittedShadeKernel<false><<<(p.count + kBlock - 1) / kBlock, kBlock, 0, stream>>>(a);
	if (launches) *launches += 1;
	return cudaGetLastError();
}

cudaError_t launchWhittedFinish(const unsigned long long* accumulators, uint32_t pixels, float4* framebuffer, cudaStream_t stream, int* launches) {
	if (!pixels) return cudaSuccess;
	whittedFinishKernel<<<(pixels + 255) / 256, 256, 0, stream>>>(accumulators, pixels, framebuffer);
	if (launches) *launches += 1;
	return cudaGetLastError();
}

} // namespace racc_b200
