// traverse_packed.cuh -- device functions of the packed-format traversal (see traverse_packed.cu for what the format is and
// why): the node step, the pair test, the miss epilogue on the texel-pair table and the traversal stacks. Shared by the
// traversal kernel (traverse_packed.cu) and by the streamed path tracer (pathstream.cu), which walks the same images
// with the same instruction sequence, so that both produce the same bits.
#pragma once

#include "traverse_common.cuh"

#include <type_traits>

namespace racc_b200 {
namespace {

typedef unsigned long long u64;

// (v,v): declared volatile so that the broadcast is not hoisted into a loop-invariant register pair;
// ptxas folds it into the scalar-broadcast operand form of FFMA2 (Rn.F32) instead.
__device__ __forceinline__ u64 splat2(float v) {
	u64 r;
	asm volatile("mov.b64 %0, {%1,%1};" : "=l"(r) : "f"(v));
	return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
	asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// (a.lo*b.lo+c.lo, a.hi*b.hi+c.hi), each an IEEE fma with flush-to-zero: FFMA2
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
	u64 d;
	asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
	return d;
}

// Miss epilogue (Kernels.h:213-221) on the texel-pair table: the arithmetic of missRadiance() of
// traverse_common.cuh, two gathers instead of four.
__device__ __forceinline__ float4 missRadiancePairs(const float4* __restrict__ envPairs, uint32_t w, uint32_t hgt, const RayState& r) {
	const float s = r.dy * r.dy + r.dz * r.dz;
	const float rlen = __frcp_rn(__fsqrt_rn(s));
	const float inv2pi = 1.0f / (2.0f * 3.141593f);
	const float rr = (rlen > 1e+6f) ? 0.0f : (acosPinned(-r.dx) * inv2pi) * rlen;
	const float u = 0.5f - rr * r.dz;
	const float v = 0.5f - rr * r.dy;
	const float fu = u * (float)(int)w - 0.5f;
	const float fv = v * (float)(int)hgt - 0.5f;
	float a, b;
	const int i0 = texelFloor(fu, a);
	int j0 = texelFloor(fv, b);
	const int j1 = min(max(j0 + 1, 0), (int)hgt - 1);
	j0 = min(max(j0, 0), (int)hgt - 1);
	const uint32_t k = (uint32_t)(min(max(i0, -1), (int)w - 1) + 1);
	const float4* row0 = envPairs + 2 * ((size_t)j0 * (w + 1) + k);
	const float4* row1 = envPairs + 2 * ((size_t)j1 * (w + 1) + k);
	float4 t00, t10, t01, t11;
	asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=f"(t00.x), "=f"(t00.y), "=f"(t00.z), "=f"(t00.w), "=f"(t10.x), "=f"(t10.y), "=f"(t10.z), "=f"(t10.w) : "l"(row0));
	asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=f"(t01.x), "=f"(t01.y), "=f"(t01.z), "=f"(t01.w), "=f"(t11.x), "=f"(t11.y), "=f"(t11.z), "=f"(t11.w) : "l"(row1));
	const float na = 1.0f - a, nb = 1.0f - b;
	const float w00 = na * nb, w10 = a * nb, w01 = na * b, w11 = a * b;
	float4 o;
	o.x = __uint_as_float(kMiss);
	o.y = ((w00 * t00.x + w10 * t10.x) + w01 * t01.x) + w11 * t11.x;
	o.z = ((w00 * t00.y + w10 * t10.y) + w01 * t01.y) + w11 * t11.y;
	o.w = ((w00 * t00.z + w10 * t10.z) + w01 * t01.z) + w11 * t11.z;
	return o;
}

__device__ __forceinline__ float4 finishRayPacked(const TraceParams& p, const RayState& r, const HitState& h) {
	if (h.index != kMiss) return hitResult(p.remap, h);
	return p.envPairs ? missRadiancePairs(p.envPairs, p.envWidth, p.envHeight, r) : missRadiance(p.env, p.envWidth, p.envHeight, r);
}

// ---------------------------------------------------------------------------------------------

// trianglePairIntersect (Kernels.h:36-115) on a packed pair; same operations as pairTest() of
// traverse_common.cuh except that n1 arrives precomputed.
__device__ __forceinline__ void pairTestPacked(u64 pairBase, uint32_t index, RayState& r, HitState& h) {
	u64 a;
	asm("mad.wide.u32 %0, %1, 64, %2;" : "=l"(a) : "r"(index), "l"(pairBase));
	float4 t0, t1, t2, t3;
	asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=f"(t0.x), "=f"(t0.y), "=f"(t0.z), "=f"(t0.w), "=f"(t1.x), "=f"(t1.y), "=f"(t1.z), "=f"(t1.w) : "l"(a));
	asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8+32];"
	             : "=f"(t2.x), "=f"(t2.y), "=f"(t2.z), "=f"(t2.w), "=f"(t3.x), "=f"(t3.y), "=f"(t3.z), "=f"(t3.w) : "l"(a));
	const float n1x = t3.x, n1y = t3.y, n1z = t3.z;
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

// Traversal stack whose first kSm entries live in shared memory, laid out [entry][thread] so that the
// 32 lanes of a warp always hit 32 different banks whatever their depths: a push or pop is ONE data-pipe
// wavefront per warp, where the thread-interleaved local-memory stack costs one per lane once the lanes'
// depths differ (profiles/r01_l1_wavefront_microbench.md). Deeper entries spill to local memory
// (Kernels.h:166 allows 64 in total).
template <int kSm, int kBlock>
struct HybridStack {
	uint32_t sp;          // entries on the stack
	uint32_t smAddr;      // shared-space byte address of entry 0 of this thread
	uint32_t localAddr;   // local-space byte address of the first spilled entry
	__device__ __forceinline__ void attach(uint32_t (*sm)[kBlock], uint32_t* spill) {
		sp = 0;
		smAddr = (uint32_t)__cvta_generic_to_shared(&sm[0][threadIdx.x]);
		localAddr = (uint32_t)__cvta_generic_to_local(spill);
	}
	__device__ __forceinline__ void reset() { sp = 0; }
	__device__ __forceinline__ bool empty() const { return sp == 0; }
	__device__ __forceinline__ void pushIf(bool pred, uint32_t v) {
		if (pred) {
			if (sp < (uint32_t)kSm) asm volatile("st.shared.u32 [%0], %1;" ::"r"(smAddr + sp * (uint32_t)(kBlock * 4)), "r"(v) : "memory");
			else asm volatile("st.local.u32 [%0], %1;" ::"l"((u64)(localAddr + (sp - kSm) * 4u)), "r"(v) : "memory");
			++sp;
		}
	}
	__device__ __forceinline__ uint32_t pop() {
		--sp;
		uint32_t v;
		if (sp < (uint32_t)kSm) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(smAddr + sp * (uint32_t)(kBlock * 4)) : "memory");
		else asm volatile("ld.local.u32 %0, [%1];" : "=r"(v) : "l"((u64)(localAddr + (sp - kSm) * 4u)) : "memory");
		return v;
	}
};

// The all-local stack with the same interface (predicated STL, no branch around it).
struct PlainStack : LocalStack {
	__device__ __forceinline__ void pushIf(bool pred, uint32_t v) {
		asm volatile(
		    "{\n\t.reg .pred pu;\n\t"
		    "setp.ne.u32 pu, %2, 0;\n\t"
		    "@pu st.local.u32 [%1], %3;\n\t"
		    "@pu add.u32 %0, %0, 4;\n\t}"
		    : "+r"(top)
		    : "l"((u64)top), "r"((uint32_t)pred), "r"(v)
		    : "memory");
	}
};

// One inner-node step (Kernels.h:170-199) on a packed node. `node` has bit 31 set. Returns the next
// reference: the nearer hit child, else the popped entry, else 0.
template <bool kCount, typename Stack>
__device__ __forceinline__ uint32_t innerStepPacked(u64 nodeBase, uint32_t node, const RayState& r, Stack& stack, unsigned& pushes) {
	u64 a;
	asm("mad.wide.u32 %0, %1, 64, %2;" : "=l"(a) : "r"(node), "l"(nodeBase)); // nodeBase is biased by -(2^31 * 64)
	u64 lx, ly, lz, rx, ry, rz, refs, unused;
	asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(lx), "=l"(ly), "=l"(lz), "=l"(rx) : "l"(a));
	asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4+32];" : "=l"(ry), "=l"(rz), "=l"(refs), "=l"(unused) : "l"(a));
	const u64 ix2 = splat2(r.ix), iy2 = splat2(r.iy), iz2 = splat2(r.iz);
	const u64 px2 = splat2(r.px), py2 = splat2(r.py), pz2 = splat2(r.pz);
	float n0x, f0x, n0y, f0y, n0z, f0z, n1x, f1x, n1y, f1y, n1z, f1z;
	unpack2(fma2(lx, ix2, px2), n0x, f0x);
	unpack2(fma2(ly, iy2, py2), n0y, f0y);
	unpack2(fma2(lz, iz2, pz2), n0z, f0z);
	unpack2(fma2(rx, ix2, px2), n1x, f1x);
	unpack2(fma2(ry, iy2, py2), n1y, f1y);
	unpack2(fma2(rz, iz2, pz2), n1z, f1z);
	const float tRay = r.tFar;
	// aabbIntersect (Kernels.h:117-135), twice
	// same value as max(max(tNear, min(nx,fx)), max(min(ny,fy), min(nz,fz))) of Kernels.h:128-131, one instruction fewer per line
	const float a0 = max3(fmaxf(r.tNear, fminf(n0x, f0x)), fminf(n0y, f0y), fminf(n0z, f0z));
	const float b0 = min3(fminf(tRay, fmaxf(n0x, f0x)), fmaxf(n0y, f0y), fmaxf(n0z, f0z));
	const float a1 = max3(fmaxf(r.tNear, fminf(n1x, f1x)), fminf(n1y, f1y), fminf(n1z, f1z));
	const float b1 = min3(fminf(tRay, fmaxf(n1x, f1x)), fmaxf(n1y, f1y), fmaxf(n1z, f1z));
	const float tFirst = a0 > b0 ? tRay : a0;
	const float tLast = a1 > b1 ? tRay : a1;
	const float firstDiff = tRay - tFirst;
	const float lastDiff = tRay - tLast;
	const bool any = firstDiff + lastDiff != 0.0f;
	const bool sgn = (int)__float_as_uint(tLast - tFirst) < 0;
	const bool both = any && fmaxf(tFirst, tLast) != tRay;
	uint32_t cf, cl;
	asm("mov.b64 {%0,%1}, %2;" : "=r"(cf), "=r"(cl) : "l"(refs));
	const uint32_t nearRef = sgn ? cl : cf, farRef = sgn ? cf : cl;
	// The push is predicated (no branch around one STL); the pop is a real branch. A predicated pop
	// measured 2x slower on DRAM-bound scenes (profiles/r01_c5_predicated_pop.md): an LDL issued with
	// most or all lanes off still sits in the load pipeline behind the warp's outstanding misses.
	uint32_t next = any ? nearRef : 0u;
	stack.pushIf(both, farRef);
	if (kCount) pushes += both;
	if (!any && !stack.empty())
		next = stack.pop();
	return next;
}

// The same step on a 32-byte quantised node (kQuant). The ray was moved into grid space once (quantRay): r.ix = cell *
// invDir, r.px = (gridOrigin - origin) * invDir, so a plane's distance is fma(q, r.ix, r.px) with q the plane's 16-bit
// cell index as a float: I2F.U16 with a half-word selector, 12 per node, on the XU pipe (quarter rate). Measured against
// the alternative that keeps the XU idle -- a 15-bit index dropped into the mantissa of 2^15 by one PRMT, bias folded
// into r.px -- the conversions on the ALU pipe cost more than they save (that pipe already runs at 64 %) and the
// coarser grid lets bounce rays into 12 % more nodes: battlefield 6880 vs 7418 Mrays/s, config 5 899 vs 928
// (profiles/r02_quantised_nodes.md). So: 16 bits, I2F.
template <bool kCount, typename Stack>
__device__ __forceinline__ uint32_t innerStepQuant(u64 nodeBase, uint32_t node, const RayState& r, Stack& stack, unsigned& pushes) {
	u64 a;
	asm("mad.wide.u32 %0, %1, 32, %2;" : "=l"(a) : "r"(node), "l"(nodeBase)); // nodeBase is biased by -(2^31 * 32)
	uint32_t w0, w1, w2, w3, w4, w5, cf, cl;
	asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3), "=r"(w4), "=r"(w5), "=r"(cf), "=r"(cl) : "l"(a));
	const float tRay = r.tFar;
	float n0x, f0x, n0y, f0y, n0z, f0z, n1x, f1x, n1y, f1y, n1z, f1z;
#define RACC_QPLANES(w, i, p, lo, hi)                                                                                  \
	{                                                                                                                  \
		u64 q;                                                                                                         \
		asm("mov.b64 %0, {%1,%2};" : "=l"(q) : "f"((float)(unsigned short)(w)), "f"((float)(unsigned short)((w) >> 16))); \
		unpack2(fma2(q, splat2(i), splat2(p)), lo, hi);                                                                \
	}
	RACC_QPLANES(w0, r.ix, r.px, n0x, f0x)
	RACC_QPLANES(w1, r.iy, r.py, n0y, f0y)
	RACC_QPLANES(w2, r.iz, r.pz, n0z, f0z)
	RACC_QPLANES(w3, r.ix, r.px, n1x, f1x)
	RACC_QPLANES(w4, r.iy, r.py, n1y, f1y)
	RACC_QPLANES(w5, r.iz, r.pz, n1z, f1z)
#undef RACC_QPLANES
	// same value as max(max(tNear, min(nx,fx)), max(min(ny,fy), min(nz,fz))) of Kernels.h:128-131, one instruction fewer per line
	const float a0 = max3(fmaxf(r.tNear, fminf(n0x, f0x)), fminf(n0y, f0y), fminf(n0z, f0z));
	const float b0 = min3(fminf(tRay, fmaxf(n0x, f0x)), fmaxf(n0y, f0y), fmaxf(n0z, f0z));
	const float a1 = max3(fmaxf(r.tNear, fminf(n1x, f1x)), fminf(n1y, f1y), fminf(n1z, f1z));
	const float b1 = min3(fminf(tRay, fmaxf(n1x, f1x)), fmaxf(n1y, f1y), fmaxf(n1z, f1z));
	const float tFirst = a0 > b0 ? tRay : a0;
	const float tLast = a1 > b1 ? tRay : a1;
	const float firstDiff = tRay - tFirst;
	const float lastDiff = tRay - tLast;
	const bool any = firstDiff + lastDiff != 0.0f;
	const bool sgn = (int)__float_as_uint(tLast - tFirst) < 0;
	const bool both = any && fmaxf(tFirst, tLast) != tRay;
	const uint32_t nearRef = sgn ? cl : cf, farRef = sgn ? cf : cl;
	uint32_t next = any ? nearRef : 0u;
	stack.pushIf(both, farRef);
	if (kCount) pushes += both;
	if (!any && !stack.empty())
		next = stack.pop();
	return next;
}

// Moves a freshly initialised ray into the quantisation grid (see innerStepQuant). Only ix..pz change; the pair test uses
// the origin and the direction, the miss epilogue the direction.
__device__ __forceinline__ void quantRay(const TraceParams& p, RayState& r) {
	r.px = (p.qOrigin[0] - r.ox) * r.ix; r.py = (p.qOrigin[1] - r.oy) * r.iy; r.pz = (p.qOrigin[2] - r.oz) * r.iz;
	r.ix = p.qCell[0] * r.ix; r.iy = p.qCell[1] * r.iy; r.iz = p.qCell[2] * r.iz;
}

} // namespace
} // namespace racc_b200
