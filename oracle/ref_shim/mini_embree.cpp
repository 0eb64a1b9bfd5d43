// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or shipped with the product.
//
// A stand-in for the nine Embree 2.x entry points the reference's Scene.cpp uses (Scene.cpp:198-211 scene set-up, :360
// rtcDeleteScene, :416 rtcIntersect8, :466 rtcIntersect). Embree 2.7.0 ships with the reference only as
// Mac/libembree.2.dylib and Win/embree.dll, so the reference's CPU query path -- executeRayQueryCPU, Scene.cpp:374-484,
// the parity reference north_star names -- could not run here. With this file linked in place of the binary, that
// function runs from ITS OWN SOURCE: the AoS -> RTCRay8 transposes, the call, the primID / tfar / u / v scatter, the
// light-probe lookup for misses and the scalar tail are the reference's; only what happens inside rtcIntersect[8] is ours.
//
// What is implemented is what the Embree 2 API documents for these calls (include/embree2/rtcore*.h, under
// /root/reference/include): a static scene with one triangle mesh (vertex buffer of 16-byte x,y,z,pad records, index
// buffer of three 32-bit indices per triangle); rtcIntersect finds the closest hit with tnear < t < tfar and writes
// tfar = t, u, v (hit point = (1-u-v) v0 + u v1 + v v2), Ng (unnormalised), geomID = mesh id, primID = triangle index;
// rays that miss keep geomID = RTC_INVALID_GEOMETRY_ID; rtcIntersect8 does that for the lanes whose `valid` word is -1.
//
// It is written independently of the engine and of the oracle on purpose -- it is the SECOND opinion: its own tree (a
// median-split BVH over single triangles, four per leaf; no pairs, no SAH), a packet traversal that keeps eight rays
// together, the plain Moeller-Trumbore test with edge vectors (Embree's published triangle test: den = Ng . dir,
// U = (dir x O) . e2, V = (dir x O) . e1, T = Ng . O, all compared after multiplying by sign(den)), and strict
// interval ends. So it can differ from the GPU kernel exactly where north_star allows a difference: which of two
// triangles wins an exact tie, and the last ulps of t, u, v.
#include <embree2/rtcore.h>
#include <embree2/rtcore_ray.h>

#include <immintrin.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace {

struct Node {
	float lo[3], hi[3];
	uint32_t left;   // inner: index of the left child (right = left + 1); leaf: first triangle slot
	uint32_t count;  // 0 = inner, else triangles in the leaf
};

struct Mesh {
	float* vertices = nullptr;   // nverts x {x, y, z, pad}
	uint32_t* indices = nullptr; // ntris x 3
	size_t nverts = 0, ntris = 0;
	std::vector<Node> nodes;
	std::vector<uint32_t> order; // leaf slots -> triangle index
};

Mesh* asMesh(RTCScene s) { return reinterpret_cast<Mesh*>(s); }

void triangleBounds(const Mesh& m, uint32_t tri, float lo[3], float hi[3]) {
	for (int k = 0; k < 3; ++k) { lo[k] = 3.0e38f; hi[k] = -3.0e38f; }
	for (int c = 0; c < 3; ++c) {
		const float* v = m.vertices + 4 * (size_t)m.indices[3 * (size_t)tri + c];
		for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], v[k]); hi[k] = std::max(hi[k], v[k]); }
	}
}

void build(Mesh& m, uint32_t node, uint32_t first, uint32_t count) {
	Node n{};
	float clo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, chi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
	for (int k = 0; k < 3; ++k) { n.lo[k] = 3.0e38f; n.hi[k] = -3.0e38f; }
	for (uint32_t i = first; i < first + count; ++i) {
		float lo[3], hi[3];
		triangleBounds(m, m.order[i], lo, hi);
		for (int k = 0; k < 3; ++k) {
			n.lo[k] = std::min(n.lo[k], lo[k]); n.hi[k] = std::max(n.hi[k], hi[k]);
			const float c = 0.5f * (lo[k] + hi[k]);
			clo[k] = std::min(clo[k], c); chi[k] = std::max(chi[k], c);
		}
	}
	if (count <= 4) {
		n.left = first;
		n.count = count;
		m.nodes[node] = n;
		return;
	}
	int axis = 0;
	for (int k = 1; k < 3; ++k) if (chi[k] - clo[k] > chi[axis] - clo[axis]) axis = k;
	const uint32_t mid = first + count / 2;
	std::nth_element(m.order.begin() + first, m.order.begin() + mid, m.order.begin() + first + count, [&](uint32_t a, uint32_t b) {
		float la[3], ha[3], lb[3], hb[3];
		triangleBounds(m, a, la, ha);
		triangleBounds(m, b, lb, hb);
		const float ca = la[axis] + ha[axis], cb = lb[axis] + hb[axis];
		return ca < cb || (ca == cb && a < b);
	});
	n.left = (uint32_t)m.nodes.size();
	n.count = 0;
	m.nodes.push_back(Node{});
	m.nodes.push_back(Node{});
	m.nodes[node] = n;
	build(m, n.left, first, mid - first);
	build(m, n.left + 1, mid, first + count - mid);
}

// Eight rays against one triangle; updates the hit lanes of `ray` in place.
inline void intersectTriangle8(const Mesh& m, uint32_t tri, __m256 active, RTCRay8& ray) {
	const uint32_t* ix = m.indices + 3 * (size_t)tri;
	const float* p0 = m.vertices + 4 * (size_t)ix[0];
	const float* p1 = m.vertices + 4 * (size_t)ix[1];
	const float* p2 = m.vertices + 4 * (size_t)ix[2];
	// e1 = v0 - v1, e2 = v2 - v0, Ng = e1 x e2 (Embree's triangle record)
	const float e1[3] = {p0[0] - p1[0], p0[1] - p1[1], p0[2] - p1[2]};
	const float e2[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
	const float ng[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
	const __m256 ox = _mm256_sub_ps(_mm256_set1_ps(p0[0]), _mm256_load_ps(ray.orgx));
	const __m256 oy = _mm256_sub_ps(_mm256_set1_ps(p0[1]), _mm256_load_ps(ray.orgy));
	const __m256 oz = _mm256_sub_ps(_mm256_set1_ps(p0[2]), _mm256_load_ps(ray.orgz));
	const __m256 dx = _mm256_load_ps(ray.dirx), dy = _mm256_load_ps(ray.diry), dz = _mm256_load_ps(ray.dirz);
	const __m256 den = _mm256_add_ps(_mm256_add_ps(_mm256_mul_ps(dx, _mm256_set1_ps(ng[0])), _mm256_mul_ps(dy, _mm256_set1_ps(ng[1]))),
	                                 _mm256_mul_ps(dz, _mm256_set1_ps(ng[2])));
	const __m256 signMask = _mm256_set1_ps(-0.0f);
	const __m256 sgn = _mm256_and_ps(den, signMask);
	const __m256 absDen = _mm256_andnot_ps(signMask, den);
	// R = dir x O
	const __m256 rx = _mm256_sub_ps(_mm256_mul_ps(dy, oz), _mm256_mul_ps(dz, oy));
	const __m256 ry = _mm256_sub_ps(_mm256_mul_ps(dz, ox), _mm256_mul_ps(dx, oz));
	const __m256 rz = _mm256_sub_ps(_mm256_mul_ps(dx, oy), _mm256_mul_ps(dy, ox));
	auto dot = [](__m256 ax, __m256 ay, __m256 az, const float b[3]) {
		return _mm256_add_ps(_mm256_add_ps(_mm256_mul_ps(ax, _mm256_set1_ps(b[0])), _mm256_mul_ps(ay, _mm256_set1_ps(b[1]))),
		                     _mm256_mul_ps(az, _mm256_set1_ps(b[2])));
	};
	const __m256 U = _mm256_xor_ps(dot(rx, ry, rz, e2), sgn);
	const __m256 V = _mm256_xor_ps(dot(rx, ry, rz, e1), sgn);
	const __m256 T = _mm256_xor_ps(dot(ox, oy, oz, ng), sgn);
	const __m256 zero = _mm256_setzero_ps();
	__m256 hit = _mm256_and_ps(active, _mm256_cmp_ps(den, zero, _CMP_NEQ_OQ));
	hit = _mm256_and_ps(hit, _mm256_cmp_ps(U, zero, _CMP_GE_OQ));
	hit = _mm256_and_ps(hit, _mm256_cmp_ps(V, zero, _CMP_GE_OQ));
	hit = _mm256_and_ps(hit, _mm256_cmp_ps(_mm256_add_ps(U, V), absDen, _CMP_LE_OQ));
	hit = _mm256_and_ps(hit, _mm256_cmp_ps(T, _mm256_mul_ps(absDen, _mm256_load_ps(ray.tnear)), _CMP_GT_OQ));
	hit = _mm256_and_ps(hit, _mm256_cmp_ps(T, _mm256_mul_ps(absDen, _mm256_load_ps(ray.tfar)), _CMP_LT_OQ));
	if (_mm256_movemask_ps(hit) == 0)
		return;
	const __m256 rcp = _mm256_div_ps(_mm256_set1_ps(1.0f), absDen);
	_mm256_maskstore_ps(ray.tfar, _mm256_castps_si256(hit), _mm256_mul_ps(T, rcp));
	_mm256_maskstore_ps(ray.u, _mm256_castps_si256(hit), _mm256_mul_ps(U, rcp));
	_mm256_maskstore_ps(ray.v, _mm256_castps_si256(hit), _mm256_mul_ps(V, rcp));
	_mm256_maskstore_ps(ray.Ngx, _mm256_castps_si256(hit), _mm256_set1_ps(ng[0]));
	_mm256_maskstore_ps(ray.Ngy, _mm256_castps_si256(hit), _mm256_set1_ps(ng[1]));
	_mm256_maskstore_ps(ray.Ngz, _mm256_castps_si256(hit), _mm256_set1_ps(ng[2]));
	_mm256_maskstore_epi32(ray.geomID, _mm256_castps_si256(hit), _mm256_set1_epi32(0));
	_mm256_maskstore_epi32(ray.primID, _mm256_castps_si256(hit), _mm256_set1_epi32((int)tri));
}

// lanes of `active` whose ray segment [tnear, tfar] overlaps the box
inline __m256 intersectBox8(const Node& n, __m256 active, const RTCRay8& ray, const __m256 inv[3]) {
	const float* org[3] = {ray.orgx, ray.orgy, ray.orgz};
	__m256 t0 = _mm256_load_ps(ray.tnear), t1 = _mm256_load_ps(ray.tfar);
	for (int k = 0; k < 3; ++k) {
		const __m256 o = _mm256_load_ps(org[k]);
		const __m256 a = _mm256_mul_ps(_mm256_sub_ps(_mm256_set1_ps(n.lo[k]), o), inv[k]);
		const __m256 b = _mm256_mul_ps(_mm256_sub_ps(_mm256_set1_ps(n.hi[k]), o), inv[k]);
		t0 = _mm256_max_ps(t0, _mm256_min_ps(a, b));
		t1 = _mm256_min_ps(t1, _mm256_max_ps(a, b));
	}
	// a little slack on both ends: this test only has to be conservative, the triangle test decides
	const __m256 slack = _mm256_set1_ps(1.0f + 4.0e-6f);
	return _mm256_and_ps(active, _mm256_cmp_ps(t0, _mm256_mul_ps(t1, slack), _CMP_LE_OQ));
}

void intersect8(const Mesh& m, __m256 valid, RTCRay8& ray) {
	if (m.nodes.empty())
		return;
	__m256 inv[3];
	const float* dir[3] = {ray.dirx, ray.diry, ray.dirz};
	for (int k = 0; k < 3; ++k) {
		// a zero component gives +-inf, which the min / max of the slab test handle (0 * inf cannot occur for origins outside the planes' exact positions; NaN lanes fail the compare and are caught by the slack-free triangle test anyway)
		inv[k] = _mm256_div_ps(_mm256_set1_ps(1.0f), _mm256_load_ps(dir[k]));
	}
	struct Entry { uint32_t node; int mask; };
	Entry stack[128];
	int sp = 0;
	stack[sp++] = Entry{0u, _mm256_movemask_ps(valid)};
	static const int bit[8] = {1, 2, 4, 8, 16, 32, 64, 128};
	while (sp) {
		const Entry e = stack[--sp];
		const __m256i lanes = _mm256_cmpeq_epi32(_mm256_and_si256(_mm256_set1_epi32(e.mask), _mm256_loadu_si256(reinterpret_cast<const __m256i*>(bit))),
		                                         _mm256_loadu_si256(reinterpret_cast<const __m256i*>(bit)));
		const Node& n = m.nodes[e.node];
		const __m256 in = intersectBox8(n, _mm256_castsi256_ps(lanes), ray, inv);
		const int mask = _mm256_movemask_ps(in);
		if (!mask)
			continue;
		if (n.count) {
			for (uint32_t i = 0; i < n.count; ++i)
				intersectTriangle8(m, m.order[n.left + i], in, ray);
		}
		else {
			stack[sp++] = Entry{n.left + 1, mask};
			stack[sp++] = Entry{n.left, mask};
		}
	}
}

} // namespace

RTCScene rtcNewScene(RTCSceneFlags, RTCAlgorithmFlags) { return reinterpret_cast<RTCScene>(new Mesh()); }

unsigned rtcNewTriangleMesh(RTCScene s, RTCGeometryFlags, size_t triangles, size_t vertices, size_t) {
	Mesh* m = asMesh(s);
	m->nverts = vertices;
	m->ntris = triangles;
	m->vertices = static_cast<float*>(malloc(vertices * 16 + 64));
	m->indices = static_cast<uint32_t*>(malloc(triangles * 12 + 64));
	return 0;
}

void* rtcMapBuffer(RTCScene s, unsigned, RTCBufferType type) {
	return type == RTC_VERTEX_BUFFER ? static_cast<void*>(asMesh(s)->vertices) : static_cast<void*>(asMesh(s)->indices);
}

void rtcUnmapBuffer(RTCScene, unsigned, RTCBufferType) {}
void rtcSetMask(RTCScene, unsigned, int) {}

void rtcCommit(RTCScene s) {
	Mesh& m = *asMesh(s);
	m.order.resize(m.ntris);
	for (size_t i = 0; i < m.ntris; ++i) m.order[i] = (uint32_t)i;
	m.nodes.clear();
	if (!m.ntris)
		return;
	m.nodes.reserve(m.ntris);
	m.nodes.push_back(Node{});
	build(m, 0, 0, (uint32_t)m.ntris);
}

void rtcDeleteScene(RTCScene s) {
	Mesh* m = asMesh(s);
	free(m->vertices);
	free(m->indices);
	delete m;
}

void rtcIntersect8(const void* valid, RTCScene s, RTCRay8& ray) {
	intersect8(*asMesh(s), _mm256_castsi256_ps(_mm256_load_si256(static_cast<const __m256i*>(valid))), ray);
}

void rtcIntersect(RTCScene s, RTCRay& ray) {
	// the single-ray entry point through the same code: lane 0 of a packet
	alignas(32) RTCRay8 p;
	memset(&p, 0, sizeof(p));
	p.orgx[0] = ray.org[0]; p.orgy[0] = ray.org[1]; p.orgz[0] = ray.org[2];
	p.dirx[0] = ray.dir[0]; p.diry[0] = ray.dir[1]; p.dirz[0] = ray.dir[2];
	p.tnear[0] = ray.tnear; p.tfar[0] = ray.tfar;
	for (int k = 0; k < 8; ++k) { p.geomID[k] = p.primID[k] = p.instID[k] = (int)RTC_INVALID_GEOMETRY_ID; p.dirx[k] = k ? 1.0f : p.dirx[0]; }
	alignas(32) int valid[8] = {-1, 0, 0, 0, 0, 0, 0, 0};
	intersect8(*asMesh(s), _mm256_castsi256_ps(_mm256_load_si256(reinterpret_cast<const __m256i*>(valid))), p);
	if (p.geomID[0] != (int)RTC_INVALID_GEOMETRY_ID) {
		ray.tfar = p.tfar[0]; ray.u = p.u[0]; ray.v = p.v[0];
		ray.Ng[0] = p.Ngx[0]; ray.Ng[1] = p.Ngy[0]; ray.Ng[2] = p.Ngz[0];
		ray.geomID = p.geomID[0]; ray.primID = p.primID[0];
	}
}
