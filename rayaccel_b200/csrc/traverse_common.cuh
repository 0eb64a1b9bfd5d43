// traverse_common.cuh -- device functions shared by the traversal kernels (traverse.cu: the
// reference-format variants; traverse_packed.cu: the packed-format kernel). Everything here restates
// /root/reference/RayAccelerator/Kernels.h in the pinned fp32 arithmetic of DESIGN.md section 3.
#pragma once

#include "engine.h"

namespace racc_b200 {
namespace {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr uint32_t kInnerBit = 0x80000000u;
constexpr uint32_t kMiss = 0xffffffffu;
constexpr int kStackSize = 64; // Kernels.h:166


struct RayState {
	float ox, oy, oz;
	float dx, dy, dz;      // after the epsilon clamp (Kernels.h:149-157)
	float ix, iy, iz;      // invDir (Kernels.h:159)
	float px, py, pz;      // OoD = -origin * invDir (Kernels.h:160)
	float tNear, tFar;
};

struct HitState {
	uint32_t index; // pair-triangle index, kMiss if none (Kernels.h:162)
	float t, u, v;
};

__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
	return fmaf(az, bz, fmaf(ay, by, ax * bx));
}

// mad_cross (Kernels.h:23-25)
#define RACC_CROSS(rx, ry, rz, ax, ay, az, bx, by, bz) \
	float rx = fmaf(ay, bz, -(az * by));               \
	float ry = fmaf(az, bx, -(ax * bz));               \
	float rz = fmaf(ax, by, -(ay * bx));

// a = origin.xyz, minT; b = direction.xyz, maxT
__device__ __forceinline__ void initRayFrom(const float4 a, const float4 b, RayState& r, HitState& h) {
	r.ox = a.x; r.oy = a.y; r.oz = a.z; r.tNear = a.w;
	r.dx = b.x; r.dy = b.y; r.dz = b.z; r.tFar = b.w;
	const float epsilon = 1e-10f;
	if (fabsf(r.dx) < epsilon) r.dx = copysignf(epsilon, r.dx);
	if (fabsf(r.dy) < epsilon) r.dy = copysignf(epsilon, r.dy);
	if (fabsf(r.dz) < epsilon) r.dz = copysignf(epsilon, r.dz);
	r.ix = __frcp_rn(r.dx); r.iy = __frcp_rn(r.dy); r.iz = __frcp_rn(r.dz);
	r.px = -r.ox * r.ix; r.py = -r.oy * r.iy; r.pz = -r.oz * r.iz;
	h.index = kMiss; h.t = r.tFar; h.u = 0.0f; h.v = 0.0f;
}

__device__ __forceinline__ void initRay(const DevRay* rays, uint32_t i, RayState& r, HitState& h) {
	initRayFrom(__ldg(&rays[i].a), __ldg(&rays[i].b), r, h);
}

// Three-input min / max (FMNMX3, new on sm_100): exact operations, so max3(a, b, c) has the bits of fmaxf(fmaxf(a, b), c) --
// minNum / maxNum are associative, NaN operands are skipped either way and -0 < +0 is a total order. One ALU-pipe
// instruction instead of two in the slab tests, on a kernel whose ALU pipe runs at 61-74 %.
__device__ __forceinline__ float max3(float a, float b, float c) {
	float d;
	asm("max.ftz.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
	return d;
}
__device__ __forceinline__ float min3(float a, float b, float c) {
	float d;
	asm("min.ftz.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
	return d;
}

// aabbIntersect (Kernels.h:117-135): entry distance, or tFar as the miss sentinel.
__device__ __forceinline__ float slab(float mnx, float mny, float mnz, float mxx, float mxy, float mxz, const RayState& r) {
	const float nx = fmaf(mnx, r.ix, r.px), ny = fmaf(mny, r.iy, r.py), nz = fmaf(mnz, r.iz, r.pz);
	const float fx = fmaf(mxx, r.ix, r.px), fy = fmaf(mxy, r.iy, r.py), fz = fmaf(mxz, r.iz, r.pz);
	const float t0 = fmaxf(fmaxf(r.tNear, fminf(nx, fx)), fmaxf(fminf(ny, fy), fminf(nz, fz)));
	const float t1 = fminf(fminf(r.tFar, fmaxf(nx, fx)), fminf(fmaxf(ny, fy), fmaxf(nz, fz)));
	return t0 > t1 ? r.tFar : t0;
}

// trianglePairIntersect (Kernels.h:36-115). Updates r.tFar and h on an accepted hit.
__device__ __forceinline__ void pairTest(const float4* __restrict__ pairs, uint32_t index, RayState& r, HitState& h) {
	const float4 t0 = __ldg(pairs + 3 * (size_t)index);
	const float4 t1 = __ldg(pairs + 3 * (size_t)index + 1);
	const float4 t2 = __ldg(pairs + 3 * (size_t)index + 2);
	// e1 = t0.xyz, e2 = t1.xyz, e3 = (t0.w,t1.w,t2.w), v0 = t2.xyz
	RACC_CROSS(n1x, n1y, n1z, t0.x, t0.y, t0.z, t1.x, t1.y, t1.z)
	RACC_CROSS(n2x, n2y, n2z, t0.w, t1.w, t2.w, t0.x, t0.y, t0.z)
	const float cx = t2.x - r.ox, cy = t2.y - r.oy, cz = t2.z - r.oz;
	RACC_CROSS(Rx, Ry, Rz, r.dx, r.dy, r.dz, cx, cy, cz)

	const float det1 = dot3(n1x, n1y, n1z, r.dx, r.dy, r.dz);
	const float det2 = dot3(n2x, n2y, n2z, r.dx, r.dy, r.dz);
	const uint32_t s1 = __float_as_uint(det1) & 0x80000000u;
	const uint32_t s2 = __float_as_uint(det2) & 0x80000000u;

	const float dRe1 = dot3(Rx, Ry, Rz, t0.x, t0.y, t0.z);
	const int iU1 = (int)(__float_as_uint(dot3(Rx, Ry, Rz, t1.x, t1.y, t1.z)) ^ s1);
	const int iV1 = (int)(__float_as_uint(dRe1) ^ s1);
	// The two negations of Kernels.h:65-66 as sign-bit flips. Written as -fma(...), a host compiler (gcc, for the CPU builds of
	// this source and of the checker) folds the minus into ONE fnmsub; when the products cancel exactly that instruction
	// returns +0 where -(+0) is -0, and the sign of that zero decides which of two triangles owns their shared edge
	// (found by tests/fuzz/fuzz_gpu.py on integer-grid meshes). The flip has one meaning everywhere.
	const uint32_t flip2 = s2 ^ 0x80000000u;
	const int iU2 = (int)(__float_as_uint(dRe1) ^ flip2);
	const int iV2 = (int)(__float_as_uint(dot3(Rx, Ry, Rz, t0.w, t1.w, t2.w)) ^ flip2);

	if (((iU1 | iV1) & (iU2 | iV2)) < 0)
		return;

	bool out1 = (iU1 | iV1) < 0;
	bool out2 = (iU2 | iV2) < 0;
	float U1 = __int_as_float(iU1), V1 = __int_as_float(iV1);
	const float U2 = __int_as_float(iU2), V2 = __int_as_float(iV2);
	float a1 = fabsf(det1);
	const float a2 = fabsf(det2);
	const float W1 = (a1 - U1) - V1;
	const float W2 = (a2 - U2) - V2;
	float T1 = __uint_as_float(__float_as_uint(dot3(n1x, n1y, n1z, cx, cy, cz)) ^ s1);
	const float T2 = __uint_as_float(__float_as_uint(dot3(n2x, n2y, n2z, cx, cy, cz)) ^ s2);

	out1 = out1 || (W1 < 0.0f || T1 <= a1 * r.tNear || T1 > a1 * r.tFar);
	out2 = out2 || (W2 < 0.0f || T2 <= a2 * r.tNear || T2 > a2 * r.tFar);
	if (out1 && out2)
		return;

	index *= 2;
	if ((!out2 && out1) || (!out1 && !out2 && T1 * a2 > T2 * a1)) {
		a1 = a2; T1 = T2; U1 = U2; V1 = V2;
		++index;
	}
	const float rcp = __frcp_rn(a1); // native_recip pinned to the IEEE reciprocal
	const float t = T1 * rcp;
	h.index = index;
	h.t = t;
	h.u = U1 * rcp;
	h.v = V1 * rcp;
	r.tFar = t;
}

// pinned acos of the miss path; identical operation sequence to oracle_acosf (oracle/racc_oracle.c)
__device__ __forceinline__ float acosPinned(float x) {
	const float pio2 = 1.57079637050628662109375f;
	const float pi = 3.1415927410125732421875f;
	const float pS0 = 1.6666586697e-01f, pS1 = -4.2743422091e-02f, pS2 = -8.6563630030e-03f, qS1 = -7.0662963390e-01f;
	const float ax = fabsf(x);
	if (!(ax < 1.0f)) {
		if (x != x) return x;
		return x > 0.0f ? 0.0f : pi;
	}
	if (ax <= 0.5f) {
		const float z = x * x;
		const float p = z * fmaf(z, fmaf(z, pS2, pS1), pS0);
		const float q = fmaf(z, qS1, 1.0f);
		const float rr = __fdiv_rn(p, q);
		return pio2 - fmaf(x, rr, x);
	}
	const float z = (1.0f - ax) * 0.5f;
	const float s = __fsqrt_rn(z);
	const float p = z * fmaf(z, fmaf(z, pS2, pS1), pS0);
	const float q = fmaf(z, qS1, 1.0f);
	const float rr = __fdiv_rn(p, q);
	const float w = 2.0f * fmaf(s, rr, s);
	return x > 0.0f ? w : pi - w;
}

__device__ __forceinline__ int texelFloor(float x, float& frac) {
	const float f = floorf(x);
	frac = x - f;
	if (!(f > -4.0f)) {
		if (f != f) { frac = 0.0f; return 0; }
		return -4;
	}
	if (f > 1.0e9f) return 1000000000;
	return (int)f;
}

// Miss epilogue (Kernels.h:213-221): angular-map lookup, bilinear, clamp-to-edge, written out with
// the OpenCL 1.2 linear-filter formula so that the oracle can follow it bit for bit.
__device__ __forceinline__ float4 missRadiance(const float4* __restrict__ env, uint32_t w, uint32_t hgt, const RayState& r) {
	if (!env)
		return make_float4(__uint_as_float(kMiss), 0.0f, 0.0f, 0.0f);
	const float s = r.dy * r.dy + r.dz * r.dz;
	const float rlen = __frcp_rn(__fsqrt_rn(s));
	const float inv2pi = 1.0f / (2.0f * 3.141593f);
	const float rr = (rlen > 1e+6f) ? 0.0f : (acosPinned(-r.dx) * inv2pi) * rlen;
	const float u = 0.5f - rr * r.dz;
	const float v = 0.5f - rr * r.dy;
	const float fu = u * (float)(int)w - 0.5f;
	const float fv = v * (float)(int)hgt - 0.5f;
	float a, b;
	int i0 = texelFloor(fu, a);
	int j0 = texelFloor(fv, b);
	const int i1 = min(max(i0 + 1, 0), (int)w - 1);
	const int j1 = min(max(j0 + 1, 0), (int)hgt - 1);
	i0 = min(max(i0, 0), (int)w - 1);
	j0 = min(max(j0, 0), (int)hgt - 1);
	const float4 t00 = __ldg(env + (size_t)j0 * w + i0);
	const float4 t10 = __ldg(env + (size_t)j0 * w + i1);
	const float4 t01 = __ldg(env + (size_t)j1 * w + i0);
	const float4 t11 = __ldg(env + (size_t)j1 * w + i1);
	const float na = 1.0f - a, nb = 1.0f - b;
	const float w00 = na * nb, w10 = a * nb, w01 = na * b, w11 = a * b;
	float4 o;
	o.x = __uint_as_float(kMiss);
	o.y = ((w00 * t00.x + w10 * t10.x) + w01 * t01.x) + w11 * t11.x;
	o.z = ((w00 * t00.y + w10 * t10.y) + w01 * t01.y) + w11 * t11.y;
	o.w = ((w00 * t00.z + w10 * t10.z) + w01 * t01.z) + w11 * t11.z;
	return o;
}

// Hit epilogue (Kernels.h:223-239): original index + barycentric rotation by the edge code.
__device__ __forceinline__ float4 hitResult(const uint32_t* __restrict__ remap, const HitState& h) {
	uint32_t index = __ldg(remap + h.index);
	const uint32_t edge = index >> 30;
	index &= 0x3fffffffu;
	const float bz = (1.0f - h.u) - h.v;
	float u = h.u, v = h.v;
	if (edge == 1) { u = bz; v = h.u; }
	else if (edge == 2) { u = h.v; v = bz; }
	return make_float4(__uint_as_float(index), h.t, u, v);
}

__device__ __forceinline__ float4 finishRay(const TraceParams& p, const RayState& r, const HitState& h) {
	return h.index == kMiss ? missRadiance(p.env, p.envWidth, p.envHeight, r) : hitResult(p.remap, h);
}

// Launch-wide ray index -> (rays, results) of its stream.
__device__ __forceinline__ void locate(const TraceParams& p, uint32_t idx, const DevRay*& rays, float4*& out, uint32_t& local) {
	if (p.nstreams == 1) {
		rays = p.single.rays; out = p.single.results; local = idx;
		return;
	}
	uint32_t lo = 0, hi = p.nstreams - 1;
	while (lo < hi) {
		const uint32_t mid = (lo + hi + 1) >> 1;
		if (__ldg(&p.streams[mid].begin) <= idx) lo = mid; else hi = mid - 1;
	}
	rays = p.streams[lo].rays; out = p.streams[lo].results; local = idx - p.streams[lo].begin;
}

// Per-ray traversal stack (Kernels.h:166: 64 entries) in local memory, addressed through a 32-bit
// local-window address so that a push is one STL and a pop one LDL with no index arithmetic.
struct LocalStack {
	uint32_t base, top; // local-space byte addresses; top == base when empty
	__device__ __forceinline__ void attach(uint32_t* storage) {
		base = top = (uint32_t)__cvta_generic_to_local(storage);
	}
	__device__ __forceinline__ void reset() { top = base; }
	__device__ __forceinline__ bool empty() const { return top == base; }
	__device__ __forceinline__ void push(uint32_t v) {
		asm volatile("st.local.u32 [%0], %1;" ::"l"((unsigned long long)top), "r"(v) : "memory");
		top += 4;
	}
	__device__ __forceinline__ uint32_t pop() {
		top -= 4;
		uint32_t v;
		asm volatile("ld.local.u32 %0, [%1];" : "=r"(v) : "l"((unsigned long long)top) : "memory");
		return v;
	}
};

} // namespace
} // namespace racc_b200
