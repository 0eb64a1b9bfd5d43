// traverse_packed.cu -- the default traversal kernel: persistent while-while with bail-out over a
// device-private "packed" copy of the scene images, written for sm_100a.
//
// WHAT it computes is unchanged (Kernels.h:36-242 in the pinned arithmetic of DESIGN.md section 3);
// every ray sees exactly the same sequence of box and pair tests as in the reference-format kernels
// of traverse.cu, so results stay bit-identical to the oracle. HOW differs in three places:
//
//   * node layout. The reference's 64-byte node (Scene.cpp:73-78) stores {lMin.xyz,lMax.x |
//     lMax.yz,rMin.xy | rMin.z,rMax.xyz}. The packed node stores each axis' (min,max) side by side,
//     {lx lX ly lY lz lZ rx rX | ry rY rz rZ first last - -}, so that one Blackwell packed-fp32
//     instruction (fma.rn.ftz.f32x2 -> FFMA2 with the ray's invDir/OoD broadcast) evaluates the near
//     and far plane of an axis at once: 6 FFMA2 per node instead of 12 FFMA. Same IEEE fma per
//     component, so the same bits.
//   * pair layout. 64 bytes instead of 48: {e1.xyz,e3.x | e2.xyz,e3.y | p0.xyz,e3.z | n1.xyz,-},
//     where n1 = e1 x e2 is computed once per scene by packPairsKernel with the very instruction
//     sequence the per-ray code used (mad_cross, Kernels.h:23-25). A pair is two aligned 256-bit
//     loads that never straddle a cache line (the 48-byte stride does every third pair).
//   * control. Child references keep the reference's encoding (bit 31 = inner), which makes the
//     state tests single signed compares (<0 inner, >0 leaf, 0 none) and the node address one
//     shift-add from a pre-biased base held in uniform registers; the push is a predicated STL.
//
// An optional permutation (TraceParams::perm, built by raysort.cu) makes the kernel visit the rays
// in a coherence-improving order; results are still written index-parallel to the rays.
//
// kQuant (tuning variant 4; SURVEY.md section 8f rank 4, "compressed node format"): the same kernel walking 32-byte
// QUANTISED inner nodes -- the same BVH2, child references and visiting rule, but both child boxes as 12 16-bit
// coordinates on a global grid over the scene bounds, rounded outwards with a quarter cell of margin. A node visit is then ONE
// 256-bit gather instead of two (the L1 data pipe, not HBM, bounds this kernel: DESIGN.md section 5.1) and the node
// image is half as large. The boxes are conservative, so a ray visits a superset of the leaves the exact boxes would
// let it into and finds the same closest hit with the same t, u, v (the pair test is unchanged); what can differ from
// the reference's result is only which of two triangles wins an EXACT tie in t (the visiting order of near-equal
// children may flip) and rays that graze a box the fp32 slab test of the reference misses by rounding. That is
// north_star's bar (ids equal except fp ties, |dt|/t <= 1e-4), not the bit-exact bar of the default format, so the
// exact format stays the default and this one is opt-in.
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

// ---------------------------------------------------------------------------------------------
// reference-format images -> packed images (once per scene, on the device so that n1 is produced by
// the same arithmetic the per-ray code would use)

__global__ void packNodesKernel(const float4* __restrict__ nodes, uint32_t count, float4* __restrict__ out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count)
		return;
	const float4 d0 = nodes[4 * (size_t)i], d1 = nodes[4 * (size_t)i + 1], d2 = nodes[4 * (size_t)i + 2], d3 = nodes[4 * (size_t)i + 3];
	// left box: min (d1.x,d1.y,d1.z) max (d1.w,d2.x,d2.y); right box: min (d2.z,d2.w,d3.x) max (d3.y,d3.z,d3.w)
	out[4 * (size_t)i + 0] = make_float4(d1.x, d1.w, d1.y, d2.x);
	out[4 * (size_t)i + 1] = make_float4(d1.z, d2.y, d2.z, d3.y);
	out[4 * (size_t)i + 2] = make_float4(d2.w, d3.z, d3.x, d3.w);
	out[4 * (size_t)i + 3] = make_float4(d0.z, d0.w, 0.0f, 0.0f);
}

__global__ void packPairsKernel(const float4* __restrict__ pairs, uint32_t count, float4* __restrict__ out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count)
		return;
	const float4 t0 = pairs[3 * (size_t)i], t1 = pairs[3 * (size_t)i + 1], t2 = pairs[3 * (size_t)i + 2];
	RACC_CROSS(n1x, n1y, n1z, t0.x, t0.y, t0.z, t1.x, t1.y, t1.z)
	out[4 * (size_t)i + 0] = t0;
	out[4 * (size_t)i + 1] = t1;
	out[4 * (size_t)i + 2] = t2;
	out[4 * (size_t)i + 3] = make_float4(n1x, n1y, n1z, 0.0f);
}

// Packed node image -> 32-byte quantised nodes: {Lx Ly Lz Rx Ry Rz | first last}, every box word = (min | max << 16) in
// cells of a 16-bit grid over the scene bounds. min rounds down and max rounds up after moving a quarter cell outwards:
// the quantised box contains the fp32 box with a margin of at least 0.25 cell, ~15x what the kernel's grid-space slab
// arithmetic can differ from the reference's (both err by about 2^-22 of the scene extent = 0.016 cell). A whole extra
// cell of padding was measured and dropped: rays that START on flat, axis-aligned geometry (battlefield's ground: its
// exact boxes have zero thickness, so a bounce ray leaving the ground never enters them) would then begin inside every
// such box on their way -- +23 % node visits and +82 % pair tests on the first bounce (profiles/r02_quantised_nodes.md).
// Arithmetic in double: done once per scene.
struct QuantGrid { float origin[3]; float cell[3]; double inverse[3]; };

__global__ void quantiseNodesKernel(const float4* __restrict__ nodes, uint32_t count, QuantGrid g, uint4* __restrict__ out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count)
		return;
	const float4 d0 = nodes[4 * (size_t)i], d1 = nodes[4 * (size_t)i + 1], d2 = nodes[4 * (size_t)i + 2], d3 = nodes[4 * (size_t)i + 3];
	// reference layout (Scene.cpp:73-78): left min (d1.x,d1.y,d1.z) max (d1.w,d2.x,d2.y); right min (d2.z,d2.w,d3.x) max (d3.y,d3.z,d3.w)
	const float mn[6] = {d1.x, d1.y, d1.z, d2.z, d2.w, d3.x};
	const float mx[6] = {d1.w, d2.x, d2.y, d3.y, d3.z, d3.w};
	uint32_t w[6];
	for (int k = 0; k < 6; ++k) {
		const int a = k % 3;
		double lo = floor(((double)mn[k] - (double)g.origin[a]) * g.inverse[a] - 0.25);
		double hi = ceil(((double)mx[k] - (double)g.origin[a]) * g.inverse[a] + 0.25);
		if (!(lo > 0.0)) lo = 0.0;          // also NaN
		if (lo > 65535.0) lo = 65535.0;     // +inf: the synthetic root's unreachable child (scene_build.cpp)
		if (!(hi > 0.0)) hi = 0.0;
		if (hi > 65535.0) hi = 65535.0;
		w[k] = (uint32_t)lo | ((uint32_t)hi << 16);
	}
	out[2 * (size_t)i + 0] = make_uint4(w[0], w[1], w[2], w[3]);
	out[2 * (size_t)i + 1] = make_uint4(w[4], w[5], __float_as_uint(d0.z), __float_as_uint(d0.w));
}

// Light probe as texel pairs: entry k of row j holds {texel(clamp(k-1)), texel(clamp(k))}, k in [0, width], so the two
// horizontally adjacent texels of a bilinear footprint -- clamp-to-edge included -- are ONE aligned 256-bit load.
__global__ void packEnvKernel(const float4* __restrict__ texels, uint32_t width, uint32_t height, float4* __restrict__ out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (width + 1) * height)
		return;
	const uint32_t j = i / (width + 1), k = i % (width + 1);
	const uint32_t x0 = k ? k - 1 : 0, x1 = k < width ? k : width - 1;
	out[2 * (size_t)i + 0] = texels[(size_t)j * width + x0];
	out[2 * (size_t)i + 1] = texels[(size_t)j * width + x1];
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
	const int iU2 = (int)(__float_as_uint(-dRe1) ^ s2);
	const int iV2 = (int)(__float_as_uint(-dot3(Rx, Ry, Rz, t0.w, t1.w, t2.w)) ^ s2);

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

template <bool kCount, int kBlock, int kMinBlocks, int kSmStack, bool kQuant>
__global__ void __launch_bounds__(kBlock, kMinBlocks) tracePackedKernel(const TraceParams p, const int fetchThreshold, const int innerBail, const int leafBail) {
	const unsigned lane = threadIdx.x & 31;
	const unsigned ltMask = (1u << lane) - 1u;
	// both bases are made opaque so that they stay in registers (otherwise they are re-derived from
	// the constant bank, several instructions, at every step)
	u64 nodeBase, pairBase;
	asm volatile("mov.b64 %0, %1;" : "=l"(nodeBase) : "l"(kQuant ? reinterpret_cast<u64>(p.qnodes) - (0x80000000ull << 5)
	                                                                    : reinterpret_cast<u64>(p.tnodes) - (0x80000000ull << 6)));
	asm volatile("mov.b64 %0, %1;" : "=l"(pairBase) : "l"(reinterpret_cast<u64>(p.tpairs)));

	const uint32_t total = p.totalPtr ? min(__ldg(p.totalPtr), p.total) : p.total;
	bool exhausted = false; // warp-uniform: the cursor has run past the last ray
	RayState r; HitState h;
	__shared__ uint32_t smStack[kSmStack ? kSmStack : 1][kSmStack ? kBlock : 1];
	uint32_t stackStorage[kStackSize - kSmStack];
	typename std::conditional<(kSmStack > 0), HybridStack<(kSmStack > 0 ? kSmStack : 1), kBlock>, PlainStack>::type stack;
	if constexpr (kSmStack > 0) stack.attach(smStack, stackStorage);
	else stack.attach(stackStorage);
	uint32_t node = 0;         // 0: no ray in flight on this lane; bit 31: at an inner node; else at a leaf
	float4* outPtr = nullptr;  // non-null: the lane holds a ray (in flight, or finished and not yet written)
	unsigned long long cInner = 0, cPairs = 0;
	unsigned cRays = 0, cHits = 0, cPushes = 0, cLeaves = 0; // the last two only in the kCount build

	for (;;) {
		// ---- retire finished lanes and refill idle ones, warp-wide -----------------------------
		unsigned idle = __ballot_sync(kFullMask, node == 0);
		if (idle == kFullMask || __popc(idle) >= (exhausted ? 32 : fetchThreshold)) {
			if (node == 0 && outPtr) {
				*outPtr = finishRayPacked(p, r, h);
				++cRays;
				cHits += h.index != kMiss;
				outPtr = nullptr;
			}
			if (!exhausted) {
				const int want = __popc(idle);
				const int leader = __ffs(idle) - 1;
				uint32_t base = 0;
				if ((int)lane == leader) base = atomicAdd(p.cursor, (uint32_t)want);
				base = __shfl_sync(kFullMask, base, leader);
				if (node == 0) {
					uint32_t idx = base + __popc(idle & ltMask);
					if (idx < total) {
						if (p.perm) idx = __ldg(p.perm + idx);
						const DevRay* rays; uint32_t local;
						locate(p, idx, rays, outPtr, local);
						outPtr += local;
						initRay(rays, local, r, h);
						if (kQuant) quantRay(p, r);
						stack.reset();
						node = kInnerBit;
					}
				}
				exhausted = base + (uint32_t)want >= total;
			}
			idle = __ballot_sync(kFullMask, node == 0);
			if (idle == kFullMask)
				break;
		}
		const unsigned liveMask = ~idle;

		// ---- inner phase: warp-uniform loop, left once too few lanes still descend -------------
		{
			unsigned innerMask = __ballot_sync(kFullMask, (int)node < 0);
			int descending = __popc(innerMask);
			bool go = innerMask == liveMask || descending >= innerBail || 2 * descending > __popc(liveMask);
			go = go && innerMask;
			while (go) {
				if ((int)node < 0) {
					if (kCount) ++cInner;
					node = kQuant ? innerStepQuant<kCount>(nodeBase, node, r, stack, cPushes) : innerStepPacked<kCount>(nodeBase, node, r, stack, cPushes);
					if (kCount) cLeaves += (int)node > 0;
				}
				innerMask = __ballot_sync(kFullMask, (int)node < 0);
				go = innerMask == liveMask || __popc(innerMask) >= innerBail;
			}
		}
		// ---- leaf phase: one pair per iteration, left once too few lanes still have pairs while
		// others wait at inner nodes -----------------------------------------------------------
		{
			bool go = __ballot_sync(kFullMask, (int)node > 0) != 0;
			while (go) {
				if ((int)node > 0) {
					pairTestPacked(pairBase, node & 0xffffffu, r, h);
					if (kCount) ++cPairs;
					// one pair of this leaf done: (count << 24 | first) -> (count-1 << 24 | first+1)
					const bool lastPair = node < 0x2000000u;
					node = !lastPair ? node + 1u - 0x1000000u : (stack.empty() ? 0u : stack.pop());
					if (kCount) cLeaves += lastPair && (int)node > 0; // the popped entry is another leaf
				}
				const unsigned leafMask = __ballot_sync(kFullMask, (int)node > 0);
				go = leafMask != 0;
				if (go && __popc(leafMask) < leafBail)
					go = __ballot_sync(kFullMask, (int)node < 0) == 0;
			}
		}
	}

	// Frame statistics (rays, hits): one atomic per warp, always on when a counter record is given;
	// this is the value the multi-GPU hit reduction sums. Visit counters only in the kCount build.
	if (p.counters) {
		unsigned long long rays = cRays, hits = cHits, pushes = cPushes, leaves = cLeaves;
		for (int o = 16; o; o >>= 1) {
			rays += __shfl_xor_sync(kFullMask, rays, o);
			hits += __shfl_xor_sync(kFullMask, hits, o);
			if (kCount) {
				cInner += __shfl_xor_sync(kFullMask, cInner, o);
				cPairs += __shfl_xor_sync(kFullMask, cPairs, o);
				pushes += __shfl_xor_sync(kFullMask, pushes, o);
				leaves += __shfl_xor_sync(kFullMask, leaves, o);
			}
		}
		if (lane == 0) {
			atomicAdd(p.counters + 0, rays);
			atomicAdd(p.counters + 1, hits);
			if (kCount) {
				atomicAdd(p.counters + 2, cInner);
				atomicAdd(p.counters + 3, cPairs);
				atomicAdd(p.counters + 4, pushes);
				atomicAdd(p.counters + 5, leaves);
			}
		}
	}
}

template <bool kCount, int kBlock, int kMinBlocks, int kSmStack = 0, bool kQuant = false>
cudaError_t launchPacked(const TraceParams& p, const Tuning& t, int smCount, cudaStream_t stream) {
	auto kernel = tracePackedKernel<kCount, kBlock, kMinBlocks, kSmStack, kQuant>;
	static thread_local int plannedDevice = -1, plannedCarveout = -2, resident = 1;
	int device = 0;
	cudaGetDevice(&device);
	cudaError_t err;
	if (plannedDevice != device || plannedCarveout != t.carveout) {
		// shared memory only for the (optional) stack tops: the rest of the 228 KB of each SM serves as L1
		int carve = 0;
		if (kSmStack > 0) {
			int smPerSm = 233472;
			cudaDeviceGetAttribute(&smPerSm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
			carve = (int)(((size_t)(kSmStack * kBlock * 4 + 1024) * kMinBlocks * 100 + smPerSm - 1) / smPerSm);
			if (carve > 100) carve = 100;
		}
		err = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, t.carveout >= 0 ? t.carveout : carve);
		if (err != cudaSuccess) return err;
		err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kernel, kBlock, 0);
		if (err != cudaSuccess) return err;
		if (resident < 1) resident = 1;
		plannedDevice = device;
		plannedCarveout = t.carveout;
	}
	int ctas = resident;
	if (t.ctasPerSm > 0 && ctas > t.ctasPerSm) ctas = t.ctasPerSm;
	long long grid = (long long)smCount * ctas;
	const long long needed = ((long long)p.total + kBlock - 1) / kBlock;
	if (grid > needed) grid = needed > 0 ? needed : 1;
	err = cudaMemsetAsync(p.cursor, 0, sizeof(uint32_t), stream);
	if (err != cudaSuccess) return err;
	// innerBail 0 and 1 mean the same (stay while any lane descends); the loop test relies on >= 1
	kernel<<<(unsigned)grid, kBlock, 0, stream>>>(p, t.fetchThreshold, t.innerBail < 1 ? 1 : t.innerBail, t.leafBail);
	return cudaGetLastError();
}

template <bool kCount>
cudaError_t dispatchPacked(const TraceParams& p, const Tuning& t, int smCount, cudaStream_t stream) {
	if (t.variant == 4) { // quantised nodes: the default launch shape, stack all-local or its tops in shared memory
		if (t.smemStack > 0) return launchPacked<kCount, 256, 5, 16, true>(p, t, smCount, stream);
		return launchPacked<kCount, 256, 5, 0, true>(p, t, smCount, stream);
	}
	if (t.smemStack > 0 && t.blockThreads == 256) { // stack tops in shared memory: 256 x 5 only
		if (t.smemStack <= 8) return launchPacked<kCount, 256, 5, 8>(p, t, smCount, stream);
		return launchPacked<kCount, 256, 5, 16>(p, t, smCount, stream);
	}
	switch (t.blockThreads * 100 + t.ctasPerSm) {
	case 12800 + 8: return launchPacked<kCount, 128, 8>(p, t, smCount, stream);
	case 12800 + 10: return launchPacked<kCount, 128, 10>(p, t, smCount, stream);
	case 12800 + 12: return launchPacked<kCount, 128, 12>(p, t, smCount, stream);
	case 25600 + 4: return launchPacked<kCount, 256, 4>(p, t, smCount, stream);
	case 25600 + 6: return launchPacked<kCount, 256, 6>(p, t, smCount, stream);
	case 51200 + 2: return launchPacked<kCount, 512, 2>(p, t, smCount, stream);
	case 51200 + 3: return launchPacked<kCount, 512, 3>(p, t, smCount, stream);
	default: return launchPacked<kCount, 256, 5>(p, t, smCount, stream);
	}
}

} // namespace

cudaError_t launchPackImages(const float4* nodes, uint32_t nodeCount, const float4* pairs, uint32_t pairCount,
                             float4* tnodes, float4* tpairs, cudaStream_t stream, int* launches) {
	if (nodeCount) {
		packNodesKernel<<<(nodeCount + 255u) / 256u, 256, 0, stream>>>(nodes, nodeCount, tnodes);
		if (launches) *launches += 1;
	}
	if (pairCount) {
		packPairsKernel<<<(pairCount + 255u) / 256u, 256, 0, stream>>>(pairs, pairCount, tpairs);
		if (launches) *launches += 1;
	}
	return cudaGetLastError();
}

cudaError_t launchQuantiseNodes(const float4* nodes, uint32_t nodeCount, const float boundsMin[3], const float boundsMax[3], void* qnodes,
                                float qOrigin[3], float qCell[3], cudaStream_t stream, int* launches) {
	QuantGrid g;
	for (int a = 0; a < 3; ++a) {
		const double extent = (double)boundsMax[a] - (double)boundsMin[a];
		const bool usable = extent > 0.0 && extent < 3.0e38;
		g.origin[a] = boundsMin[a];
		g.cell[a] = usable ? (float)(extent / 65533.0) : 0.0f;   // head room for the outward rounding
		g.inverse[a] = usable ? 65533.0 / extent : 0.0;
		qOrigin[a] = g.origin[a];
		qCell[a] = g.cell[a];
	}
	if (nodeCount) {
		quantiseNodesKernel<<<(nodeCount + 255u) / 256u, 256, 0, stream>>>(nodes, nodeCount, g, static_cast<uint4*>(qnodes));
		if (launches) *launches += 1;
	}
	return cudaGetLastError();
}

cudaError_t launchPackEnv(const float4* texels, uint32_t width, uint32_t height, float4* pairsOut, cudaStream_t stream, int* launches) {
	const uint32_t n = (width + 1) * height;
	packEnvKernel<<<(n + 255u) / 256u, 256, 0, stream>>>(texels, width, height, pairsOut);
	if (launches) *launches += 1;
	return cudaGetLastError();
}

cudaError_t launchTracePacked(const TraceParams& p, const Tuning& t, int counterMode, int smCount, cudaStream_t stream, int* launches) {
	if (!p.total)
		return cudaSuccess;
	if (launches) *launches += 1;
	return counterMode == 2 ? dispatchPacked<true>(p, t, smCount, stream) : dispatchPacked<false>(p, t, smCount, stream);
}

} // namespace racc_b200
