// scene_build.cpp -- full-sweep SAH BVH2 + triangle-pair packing for the B200 traversal kernel.
//
// Own implementation; decision-equivalent to the reference (see scene_build.h). The places where
// bit-level agreement with the reference's arithmetic decides the tree are marked "PIN:" and cite
// the reference line they must agree with. Compiled with -ffp-contract=off: only written FMAs exist.
#include "scene_build.h"

#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <queue>
#include <thread>

namespace racc_b200 {
namespace {

constexpr uint32_t kNone = 0xffffffffu;

template <typename T>
struct AlignedBuffer {
	T* p = nullptr;
	explicit AlignedBuffer(size_t n) { p = static_cast<T*>(_mm_malloc(std::max<size_t>(n, 1) * sizeof(T) + 64, 64)); }
	~AlignedBuffer() { _mm_free(p); }
	AlignedBuffer(const AlignedBuffer&) = delete;
	AlignedBuffer& operator=(const AlignedBuffer&) = delete;
	T& operator[](size_t i) { return p[i]; }
	const T& operator[](size_t i) const { return p[i]; }
};

// The reference runs its builder threads with FTZ+DAZ (Threading.h:77-79).
struct FtzScope {
	unsigned saved;
	FtzScope() : saved(_mm_getcsr()) {
		_MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
		_MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
	}
	~FtzScope() { _mm_setcsr(saved); }
};

struct Builder {
	uint32_t n;
	const float* tb;            // per-triangle bounds, 8 floats each
	uint32_t* sorted[3];        // triangle ids sorted by box centre along x / y / z
	uint32_t* scratch;          // partition scratch, indexed by position
	float* leftSah;             // prefix SAH, indexed by position (Bvh2.cpp:23)
	uint8_t* goesLeft;          // indexed by triangle id
	BuildNode* nodes;
	std::atomic<int> spareThreads{0};

	static inline __m256 box(const float* tb, uint32_t tri) { return _mm256_load_ps(tb + 8 * (size_t)tri); }

	// PIN Bvh2.cpp:76-80  surfaceArea(): (d0*d1 + d1*d2) + d0*d2 with d = (-min) + max
	static inline float areaScalar(__m256 b) {
		__m128 d = _mm_add_ps(_mm256_castps256_ps128(b), _mm256_extractf128_ps(b, 1));
		float e[4];
		_mm_storeu_ps(e, d);
		return (e[0] * e[1] + e[1] * e[2]) + e[0] * e[2];
	}

	// PIN Bvh2.cpp:339,405  the 8-wide blocks use fma(d0,d1, fma(d0,d2, d1*d2)) instead
	static inline float areaBlock(__m256 b) {
		__m128 d = _mm_add_ps(_mm256_castps256_ps128(b), _mm256_extractf128_ps(b, 1));
		float e[4];
		_mm_storeu_ps(e, d);
		return __builtin_fmaf(e[0], e[1], __builtin_fmaf(e[0], e[2], e[1] * e[2]));
	}

	__m256 rangeBounds(const uint32_t* ids, uint32_t first, uint32_t last) const {
		__m256 b = box(tb, ids[first]);
		for (uint32_t i = first + 1; i < last; ++i)
			b = _mm256_max_ps(b, box(tb, ids[i]));
		return b;
	}

	// Stable split of one sorted list by goesLeft[] (Bvh2.cpp:217-240).
	void splitList(uint32_t* ids, uint32_t first, uint32_t last) {
		uint32_t l = first, r = first;
		for (uint32_t i = first; i < last; ++i) {
			uint32_t t = ids[i];
			if (goesLeft[t]) ids[l++] = t;
			else scratch[r++] = t;
		}
		std::memcpy(ids + l, scratch + first, sizeof(uint32_t) * (size_t)(r - first));
	}

	// Evaluates every split position of one axis. Updates bestSah / returns the best pivot found
	// on this axis (kNone if none beat bestSah). Mirrors the evaluation order, the block/scalar
	// arithmetic split and the pruning of Bvh2.cpp:287-460 so that ties and roundings agree.
	uint32_t sweepAxis(const uint32_t* ids, uint32_t first, uint32_t last, float& bestSah) {
		int i = (int)first;
		const int iFirst = (int)first, iLast = (int)last;
		__m256 run = box(tb, ids[first]);
		bool pruned = false;

		// left-to-right: leftSah[k] = area(union ids[first..k]) * (k - first + 1)
		for (; i < iLast - 8; i += 8) {
			int worse = -1;
			for (int j = 0; j < 8; ++j) {
				run = _mm256_max_ps(box(tb, ids[i + j]), run);
				float sah = areaBlock(run) * (float)(i - iFirst + 1 + j);
				leftSah[i + j] = sah;
				if (worse < 0 && sah > bestSah) worse = j; // PIN Bvh2.cpp:346-351
			}
			if (worse >= 0) {
				i += worse;
				pruned = true;
				break;
			}
		}
		if (!pruned) {
			for (; i < iLast - 1; ++i) { // PIN Bvh2.cpp:354-357 scalar tail
				run = _mm256_max_ps(run, box(tb, ids[i]));
				leftSah[i] = areaScalar(run) * (float)(i - iFirst + 1);
			}
		}

		// right-to-left from position i: pivot p splits into [first,p) | [p,last)
		run = rangeBounds(ids, (uint32_t)i, last);
		uint32_t bestPivot = kNone;

		for (; i > iFirst + 7; i -= 8) {
			float sah[8];
			bool better = false, worseRhs = false;
			for (int j = 0; j < 8; ++j) {
				run = _mm256_max_ps(box(tb, ids[i - j]), run);
				float rightSah = areaBlock(run) * (float)(iLast - i + j);
				sah[j] = leftSah[i - 1 - j] + rightSah;
				better |= sah[j] < bestSah;      // PIN Bvh2.cpp:417
				worseRhs |= rightSah > bestSah;  // PIN Bvh2.cpp:418 (against the not-yet-updated best)
			}
			float minSah = sah[0];
			int arg = 0;
			for (int j = 1; j < 8; ++j)
				if (sah[j] < minSah) { minSah = sah[j]; arg = j; } // first lane holding the minimum
			if (minSah < bestSah) bestSah = minSah;
			if (better) bestPivot = (uint32_t)(i - arg); // PIN Bvh2.cpp:428-429
			if (worseRhs) return bestPivot;
		}
		for (; i > iFirst; --i) { // PIN Bvh2.cpp:438-450 scalar tail
			run = _mm256_max_ps(run, box(tb, ids[i]));
			float sah = leftSah[i - 1] + areaScalar(run) * (float)(iLast - i);
			if (sah < bestSah) {
				bestSah = sah;
				bestPivot = (uint32_t)i;
			}
		}
		return bestPivot;
	}

	void build(uint32_t slot, bool isRoot) {
		BuildNode& node = nodes[slot];
		const uint32_t first = node.first, last = node.last;

		if (!isRoot)
			_mm256_storeu_ps(node.bounds, rangeBounds(sorted[0], first, last));

		if (last - first <= 2) // Bvh2.cpp:272
			return;

		const float parentArea = areaScalar(_mm256_loadu_ps(node.bounds));
		uint32_t bestDim = kNone, pivot = kNone;
		bool split = false;

		if (parentArea > 0.0f) {
			float bestSah = std::numeric_limits<float>::infinity();
			for (uint32_t dim = 0; dim < 3; ++dim) {
				uint32_t p = sweepAxis(sorted[dim], first, last, bestSah);
				if (p != kNone) {
					pivot = p;
					bestDim = dim;
				}
			}
			// PIN Bvh2.cpp:462-467  cost = 2 + rcp_ss(area) * bestSah  vs  triangle count
			const float rcpArea = _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(parentArea)));
			const float cost = 2.0f + (1.0f * rcpArea) * bestSah;
			split = !(cost > (float)(int)(last - first) * 1.0f) && pivot != kNone;
		}
		if (!split) {
			if (last - first >= 127) { // Bvh2.cpp:468-471,478-480: leaf refs hold a 7-bit count
				bestDim = 0;
				pivot = (first + last) >> 1;
			}
			else {
				return;
			}
		}

		// partition: the chosen axis is already split at pivot; re-split the other two stably
		{
			const uint32_t* ref = sorted[bestDim];
			for (uint32_t i = first; i < pivot; ++i) goesLeft[ref[i]] = 1;
			for (uint32_t i = pivot; i < last; ++i) goesLeft[ref[i]] = 0;
			splitList(sorted[(bestDim + 1) % 3], first, last);
			splitList(sorted[(bestDim + 2) % 3], first, last);
		}

		const uint32_t left = slot + 1;
		const uint32_t right = slot + 2 * (pivot - first);
		node.kind = bestDim + 1;
		node.left = left;
		node.right = right;

		nodes[left] = BuildNode{0, slot, first, pivot, kNone, kNone, {}};
		nodes[right] = BuildNode{0, slot, pivot, last, kNone, kNone, {}};

		// Subtrees touch disjoint position ranges / triangle ids / slot blocks: fork when big.
		if (pivot - first > 16384 && last - pivot > 16384 && spareThreads.fetch_sub(1) > 0) {
			std::thread t([this, left] { FtzScope ftz; build(left, false); });
			build(right, false);
			t.join();
			spareThreads.fetch_add(1);
		}
		else {
			if (pivot - first > 16384 && last - pivot > 16384) spareThreads.fetch_add(1); // undo failed claim
			build(left, false);
			build(right, false);
		}
	}
};

inline uint32_t sortKey(float mid) {
	// order-preserving float -> uint map (Bvh2.cpp:661-669,743-745)
	uint32_t u;
	std::memcpy(&u, &mid, 4);
	return u ^ ((u & 0x80000000u) ? 0xffffffffu : 0x80000000u);
}

// PIN Bvh2.cpp:671-687: the reference's stable radix sort breaks ties between equal centres by the
// position a key was WRITTEN to, and its 8-wide key packing (unpacklo/unpackhi per 128-bit lane)
// writes each group of eight triangles in the order 0,1,4,5,2,3,6,7. Triangles past the last
// multiple of 32 go through the scalar path (Bvh2.cpp:717-750) in natural order. The map is its
// own inverse.
inline uint32_t tiePosition(uint32_t i, uint32_t n) {
	if (i >= (n & ~31u)) return i;
	const uint32_t k = i & 7u;
	const uint32_t swapped = (k >= 2 && k <= 5) ? (k ^ 6u) : k; // 2<->4, 3<->5
	return (i & ~7u) | swapped;
}

} // namespace

namespace {

// compact: pre-order walk, parents before children, first child before last child
void compactBvh2(const BuildNode* nodes, const uint32_t* sorted0, uint32_t n, Bvh2* out) {
	out->nodes.clear();
	out->nodes.reserve((size_t)n);
	out->triangles.assign(sorted0, sorted0 + n);
	std::vector<std::pair<uint32_t, uint32_t>> stack; // (slot, compact parent)
	stack.emplace_back(0u, kNone);
	while (!stack.empty()) {
		auto [slot, parent] = stack.back();
		stack.pop_back();
		const BuildNode& bn = nodes[slot];
		const uint32_t me = (uint32_t)out->nodes.size();
		Bvh2::Node nd{};
		nd.kind = bn.kind;
		nd.parent = parent;
		nd.first = bn.first;
		nd.last = bn.last;
		for (int k = 0; k < 3; ++k) {
			nd.bbMin[k] = -bn.bounds[k];
			nd.bbMax[k] = bn.bounds[4 + k];
		}
		out->nodes.push_back(nd);
		if (parent != kNone) {
			Bvh2::Node& pn = out->nodes[parent];
			// children are visited first-then-last; the first to arrive fills `first`
			if (pn.first == kNone) pn.first = me; else pn.last = me;
		}
		if (bn.kind) {
			out->nodes[me].first = kNone;
			out->nodes[me].last = kNone;
			stack.emplace_back(bn.right, me);
			stack.emplace_back(bn.left, me);
		}
	}
}

} // namespace

bool fillRcpTable(float table[2048]) {
	auto rcp = [](float x) { return _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(x))); };
	auto fromBits = [](uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; };
	for (uint32_t i = 0; i < 2048; ++i)
		table[i] = rcp(fromBits(0x3f800000u | (i << 12)));
	// the two properties the device relies on, sampled
	uint32_t s = 12345u;
	for (int k = 0; k < 200000; ++k) {
		s = s * 1664525u + 1013904223u;
		const uint32_t m = s >> 9, e = 1u + ((s >> 3) % 253u);
		const float got = rcp(fromBits((e << 23) | m));
		const float want = table[m >> 12] * fromBits((254u - e) << 23);
		if (std::memcmp(&got, &want, 4) != 0 && e > 2 && e < 252)
			return false;
	}
	return true;
}

bool buildBvh2(const float* vertices4, uint32_t vertexCount, const uint32_t* indices, uint32_t triangleCount,
               int threads, Bvh2* out, const char** error, DeviceBvhBuilder deviceBuilder) {
	static const char* kEmpty = "scene has no triangles";
	static const char* kIndex = "triangle index out of range";
	static const char* kTooBig = "scene exceeds 2^30 triangles (remap word holds 30 index bits)";
	if (!triangleCount) { if (error) *error = kEmpty; return false; }
	if (triangleCount >= (1u << 30)) { if (error) *error = kTooBig; return false; }
	for (size_t i = 0; i < (size_t)triangleCount * 3; ++i)
		if (indices[i] >= vertexCount) { if (error) *error = kIndex; return false; }

	if (deviceBuilder) {
		std::vector<BuildNode> devNodes;
		std::vector<uint32_t> devSorted;
		if (!deviceBuilder(vertices4, vertexCount, indices, triangleCount, &devNodes, &devSorted, error))
			return false;
		compactBvh2(devNodes.data(), devSorted.data(), triangleCount, out);
		return true;
	}

	FtzScope ftz;
	const uint32_t n = triangleCount;
	if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());

	AlignedBuffer<float> tb((size_t)n * 8);
	AlignedBuffer<uint32_t> s0(n), s1(n), s2(n), scratch(n);
	AlignedBuffer<float> leftSah(n);
	AlignedBuffer<uint8_t> goesLeft(n);
	std::vector<uint64_t> keys[3];
	for (auto& k : keys) k.resize(n);

	// per-triangle bounds + sort keys (Bvh2.cpp:537-753)
	__m256 sceneBounds = _mm256_set1_ps(-std::numeric_limits<float>::infinity());
	{
		const __m128 neg = _mm_set1_ps(-0.0f), half = _mm_set1_ps(0.5f);
		for (uint32_t i = 0; i < n; ++i) {
			const uint32_t* t = indices + 3 * (size_t)i;
			__m128 p0 = _mm_loadu_ps(vertices4 + 4 * (size_t)t[0]);
			__m128 p1 = _mm_loadu_ps(vertices4 + 4 * (size_t)t[1]);
			__m128 p2 = _mm_loadu_ps(vertices4 + 4 * (size_t)t[2]);
			__m128 mn = _mm_min_ps(_mm_min_ps(p0, p1), p2);
			__m128 mx = _mm_max_ps(_mm_max_ps(p0, p1), p2);
			float mid[4];
			_mm_storeu_ps(mid, _mm_mul_ps(_mm_add_ps(mn, mx), half));
			__m256 b = _mm256_insertf128_ps(_mm256_castps128_ps256(_mm_xor_ps(mn, neg)), mx, 1);
			_mm256_store_ps(&tb[(size_t)i * 8], b);
			sceneBounds = _mm256_max_ps(sceneBounds, b);
			for (int d = 0; d < 3; ++d)
				keys[d][i] = ((uint64_t)sortKey(mid[d]) << 32) | tiePosition(i, n);
		}
	}

	// three centre-sorted lists; ties keep triangle order (LSD radix on the high word, Bvh2.cpp:128-184)
	uint32_t* sorted[3] = {s0.p, s1.p, s2.p};
	{
		auto sortOne = [&](int d) {
			std::sort(keys[d].begin(), keys[d].end());
			for (uint32_t i = 0; i < n; ++i) sorted[d][i] = tiePosition((uint32_t)keys[d][i], n); // self-inverse
			std::vector<uint64_t>().swap(keys[d]);
		};
		if (threads > 1 && n > 100000) {
			std::thread a(sortOne, 0), b(sortOne, 1);
			sortOne(2);
			a.join();
			b.join();
		}
		else {
			for (int d = 0; d < 3; ++d) sortOne(d);
		}
	}

	std::vector<BuildNode> nodes((size_t)n * 2);
	nodes[0] = BuildNode{0, kNone, 0, n, kNone, kNone, {}};
	_mm256_storeu_ps(nodes[0].bounds, sceneBounds);

	Builder b;
	b.n = n;
	b.tb = tb.p;
	b.sorted[0] = sorted[0]; b.sorted[1] = sorted[1]; b.sorted[2] = sorted[2];
	b.scratch = scratch.p;
	b.leftSah = leftSah.p;
	b.goesLeft = goesLeft.p;
	b.nodes = nodes.data();
	b.spareThreads.store(threads - 1);
	b.build(0, true);

	compactBvh2(nodes.data(), sorted[0], n, out);
	return true;
}

namespace {

struct Float3 { float x, y, z; };

inline Float3 vertexAt(const float* v4, uint32_t i) { return Float3{v4[4 * (size_t)i], v4[4 * (size_t)i + 1], v4[4 * (size_t)i + 2]}; }

// Directed edge (a[e0], a[e0+1]) of triangle a equals the reversed edge (b[e1+1], b[e1]) of b
// (Scene.cpp:109-120). Scan order e0 outer, e1 inner decides which edge wins.
inline bool sharedEdge(const uint32_t* a, const uint32_t* b, unsigned& e0, unsigned& e1) {
	for (unsigned i = 0; i < 3; ++i)
		for (unsigned j = 0; j < 3; ++j)
			if (a[i] == b[(j + 1) % 3] && a[(i + 1) % 3] == b[j]) {
				e0 = i;
				e1 = j;
				return true;
			}
	return false;
}

inline GpuPair makePair(Float3 p0, Float3 p1, Float3 p2, Float3 p3) {
	// e1 = p0 - p1, e2 = p2 - p0, e3 = p3 - p0 (Scene.cpp:149-153)
	GpuPair p;
	p.e1[0] = p0.x - p1.x; p.e1[1] = p0.y - p1.y; p.e1[2] = p0.z - p1.z;
	p.e2[0] = p2.x - p0.x; p.e2[1] = p2.y - p0.y; p.e2[2] = p2.z - p0.z;
	p.p0[0] = p0.x; p.p0[1] = p0.y; p.p0[2] = p0.z;
	p.e3x = p3.x - p0.x; p.e3y = p3.y - p0.y; p.e3z = p3.z - p0.z;
	return p;
}

// Greedy pairing of one leaf's triangles, in list order (Scene.cpp:122-181,251-256).
void packLeaf(const uint32_t* tris, uint32_t count, const float* v4, const uint32_t* indices,
              std::vector<GpuPair>& pairs, std::vector<uint32_t>& remap) {
	uint32_t cand[128];
	uint32_t m = count;
	for (uint32_t i = 0; i < count; ++i) cand[i] = tris[i];
	while (m) {
		const uint32_t a = cand[0];
		--m;
		for (uint32_t i = 0; i < m; ++i) cand[i] = cand[i + 1];
		const uint32_t* ta = indices + 3 * (size_t)a;
		bool merged = false;
		for (uint32_t i = 0; i < m; ++i) {
			const uint32_t b = cand[i];
			const uint32_t* tbv = indices + 3 * (size_t)b;
			unsigned e0, e1;
			if (!sharedEdge(ta, tbv, e0, e1)) continue;
			remap.push_back(a | (e0 << 30));
			remap.push_back(b | ((e1 + 1) << 30)); // 3 == "no rotation", same as 0 (Kernels.h:232-235)
			Float3 p0 = vertexAt(v4, ta[e0]);
			Float3 p1 = vertexAt(v4, ta[(e0 + 1) % 3]);
			Float3 p2 = vertexAt(v4, ta[(e0 + 2) % 3]);
			Float3 p3 = vertexAt(v4, tbv[(e1 + 2) % 3]);
			pairs.push_back(makePair(p0, p1, p2, p3));
			--m;
			for (uint32_t k = i; k < m; ++k) cand[k] = cand[k + 1];
			merged = true;
			break;
		}
		if (!merged) {
			// singleton: second triangle degenerates to (p0, p1, p1), never hit (Scene.cpp:160-180)
			remap.push_back(a);
			remap.push_back(0);
			Float3 p0 = vertexAt(v4, ta[0]), p1 = vertexAt(v4, ta[1]), p2 = vertexAt(v4, ta[2]);
			pairs.push_back(makePair(p0, p1, p2, p1));
		}
	}
}

inline float boxArea(const Bvh2::Node& n) {
	float dx = n.bbMax[0] - n.bbMin[0], dy = n.bbMax[1] - n.bbMin[1], dz = n.bbMax[2] - n.bbMin[2];
	return dx * dy + dy * dz + dx * dz;
}

} // namespace

bool buildSceneImages(const float* vertices4, uint32_t vertexCount, const uint32_t* indices, uint32_t indexCount,
                      int threads, SceneImages* out, const char** error, DeviceBvhBuilder deviceBuilder) {
	static const char* kMod3 = "index count is not a multiple of 3";
	static const char* kEmpty = "scene has no triangles";
	static const char* kPairs = "scene exceeds 2^24 triangle pairs (leaf reference holds 24 index bits)";
	if (indexCount % 3) { if (error) *error = kMod3; return false; }
	const bool verbose = getenv("RACC_B200_BUILD_VERBOSE") != nullptr;
	auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	const double t0 = now();
	Bvh2 bvh;
	if (!buildBvh2(vertices4, vertexCount, indices, indexCount / 3, threads, &bvh, error, deviceBuilder))
		return false;
	if (bvh.nodes.empty() || indexCount == 0) { if (error) *error = kEmpty; return false; }
	if (!bvh.nodes[0].kind) {
		// The SAH test left the root a leaf (it may for 1..126 triangles). The traversal starts at inner node 0
		// (Kernels.h:168; the reference uploads an EMPTY node image in this case, Scene.cpp:274-342, and cannot trace
		// the scene at all), so the image gets one synthetic inner root: first child = that leaf, last child = a box
		// at +infinity that no finite ray enters (every slab distance is +-inf, so t0 > t1). Its reference names the
		// same leaf, so even a ray with maxT = inf that "enters" it only re-tests pairs it has already tested.
		const Bvh2::Node& leaf = bvh.nodes[0];
		out->nodes.assign(1, GpuNode{});
		out->pairs.clear();
		out->remap.clear();
		packLeaf(bvh.triangles.data() + leaf.first, leaf.last - leaf.first, vertices4, indices, out->pairs, out->remap);
		GpuNode& g = out->nodes[0];
		g.kind = 1;
		g.parent = kNone;
		g.first = g.last = ((uint32_t)out->pairs.size() << 24) | 0u;
		for (int k = 0; k < 3; ++k) {
			g.leftMin[k] = leaf.bbMin[k];
			g.leftMax[k] = leaf.bbMax[k];
			g.rightMin[k] = g.rightMax[k] = std::numeric_limits<float>::infinity();
			out->boundsMin[k] = leaf.bbMin[k];
			out->boundsMax[k] = leaf.bbMax[k];
		}
		out->realPairs = (uint32_t)out->pairs.size();
		do {
			out->pairs.push_back(out->pairs[0]);
		} while ((out->pairs.size() * 3) % 32 != 0);
		out->depth = 2;
		return true;
	}

	const uint32_t nodeCount = (uint32_t)bvh.nodes.size();
	const double t1 = now();

	// Device order of inner nodes: largest surface area first. A child's box lies inside its
	// parent's, so parents precede children and node 0 is the root; the first K nodes are the K
	// most-likely-visited ones, which is what the kernel stages in shared memory.
	std::vector<uint32_t> order;      // device index -> bvh node
	std::vector<uint32_t> deviceOf(nodeCount, kNone);
	{
		using Item = std::pair<float, uint32_t>;
		auto cmp = [](const Item& a, const Item& b) { return a.first < b.first || (a.first == b.first && a.second > b.second); };
		std::priority_queue<Item, std::vector<Item>, decltype(cmp)> heap(cmp);
		heap.emplace(std::numeric_limits<float>::infinity(), 0u);
		while (!heap.empty()) {
			uint32_t i = heap.top().second;
			heap.pop();
			deviceOf[i] = (uint32_t)order.size();
			order.push_back(i);
			const Bvh2::Node& nd = bvh.nodes[i];
			if (bvh.nodes[nd.first].kind) heap.emplace(boxArea(bvh.nodes[nd.first]), nd.first);
			if (bvh.nodes[nd.last].kind) heap.emplace(boxArea(bvh.nodes[nd.last]), nd.last);
		}
	}

	const double t2 = now();
	out->nodes.resize(order.size());
	out->pairs.clear();
	out->remap.clear();
	out->pairs.reserve(indexCount / 3 + 32);
	out->remap.reserve((size_t)(indexCount / 3) * 2);

	// Leaves are packed in the order their parents appear on the device, first child then last
	// child, so a hot node's leaves sit next to each other.
	for (uint32_t d = 0; d < (uint32_t)order.size(); ++d) {
		const Bvh2::Node& nd = bvh.nodes[order[d]];
		GpuNode g{};
		g.kind = nd.kind;
		g.parent = nd.parent == kNone ? kNone : deviceOf[nd.parent];
		const uint32_t child[2] = {nd.first, nd.last};
		uint32_t ref[2];
		for (int c = 0; c < 2; ++c) {
			const Bvh2::Node& cn = bvh.nodes[child[c]];
			if (cn.kind) {
				ref[c] = 0x80000000u | deviceOf[child[c]];
			}
			else {
				const uint32_t start = (uint32_t)out->pairs.size();
				packLeaf(bvh.triangles.data() + cn.first, cn.last - cn.first, vertices4, indices, out->pairs, out->remap);
				const uint32_t count = (uint32_t)out->pairs.size() - start;
				if (start + count > (1u << 24)) { if (error) *error = kPairs; return false; }
				ref[c] = (count << 24) | start; // Scene.cpp:298,308
			}
		}
		g.first = ref[0];
		g.last = ref[1];
		const Bvh2::Node& l = bvh.nodes[nd.first];
		const Bvh2::Node& r = bvh.nodes[nd.last];
		for (int k = 0; k < 3; ++k) {
			g.leftMin[k] = l.bbMin[k]; g.leftMax[k] = l.bbMax[k];
			g.rightMin[k] = r.bbMin[k]; g.rightMax[k] = r.bbMax[k];
		}
		out->nodes[d] = g;
	}
	out->realPairs = (uint32_t)out->pairs.size();

	// tail padding to a multiple of 32 float4, at least one pair (Scene.cpp:335-338)
	do {
		out->pairs.push_back(out->pairs[0]);
	} while ((out->pairs.size() * 3) % 32 != 0);

	const double t3 = now();
	// depth + bounds
	{
		uint32_t depth = 0;
		std::vector<std::pair<uint32_t, uint32_t>> st;
		st.emplace_back(0u, 1u);
		while (!st.empty()) {
			auto [i, dep] = st.back();
			st.pop_back();
			depth = std::max(depth, dep);
			const Bvh2::Node& nd = bvh.nodes[i];
			if (nd.kind) {
				st.emplace_back(nd.first, dep + 1);
				st.emplace_back(nd.last, dep + 1);
			}
		}
		out->depth = depth;
		for (int k = 0; k < 3; ++k) {
			out->boundsMin[k] = bvh.nodes[0].bbMin[k];
			out->boundsMax[k] = bvh.nodes[0].bbMax[k];
		}
	}
	if (verbose)
		fprintf(stderr, "racc scene build: SAH tree (%s) %.1f ms, node order %.1f ms, pair merge + packing %.1f ms, depth %.1f ms\n",
		        deviceBuilder ? "device, incl. compaction" : "host", (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (now() - t3) * 1e3);
	return true;
}

} // namespace racc_b200
