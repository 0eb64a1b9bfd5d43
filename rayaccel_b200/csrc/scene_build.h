// scene_build.h -- host-side scene construction for the B200 engine.
//
// Produces the three device images the traversal kernel walks (inner-node array, triangle-pair
// array, pair-triangle -> original-triangle remap) from an indexed triangle mesh. The images are
// decision-for-decision equivalent to what the reference builds in
//   /root/reference/RayAccelerator/Bvh2.cpp:257-535,772-907  (full-sweep SAH BVH2)
//   /root/reference/RayAccelerator/Scene.cpp:122-181,223-338 (greedy pair merge, node packing)
// i.e. the same tree topology, the same child boxes, the same pairs in the same leaf order and the
// same remap words -- only the *numbering* (which the reference leaves to thread timing,
// Bvh2.cpp:489) is ours: nodes are laid out hottest-first so the top of the tree can be staged in
// shared memory by the kernel.
#pragma once

#include <cstdint>
#include <vector>

namespace racc_b200 {

// Output of the SAH builder: a binary tree over a permuted triangle list.
struct Bvh2 {
	struct Node {
		uint32_t kind;        // 0 = leaf, else split axis + 1            (Bvh2.h:16)
		uint32_t parent;      // index into nodes, 0xffffffff for the root
		uint32_t first, last; // leaf: triangle range; inner: child node indices (Bvh2.h:17)
		float bbMin[3];
		float bbMax[3];
	};
	std::vector<Node> nodes;         // compact, root at 0, parents before children
	std::vector<uint32_t> triangles; // leaf ranges index into this
};

// 64-byte inner node, both child boxes inline (Scene.cpp:73-78; Kernels.h:173-186).
struct GpuNode {
	uint32_t kind, parent;
	uint32_t first, last; // bit31: inner-node index; else (pairCount << 24) | firstPair
	float leftMin[3], leftMax[3];
	float rightMin[3], rightMax[3];
};
static_assert(sizeof(GpuNode) == 64, "node must be 64 bytes");

// 48-byte triangle pair sharing edge e1 (Scene.cpp:80-87).
struct GpuPair {
	float e1[3], e3x;
	float e2[3], e3y;
	float p0[3], e3z;
};
static_assert(sizeof(GpuPair) == 48, "pair must be 48 bytes");

struct SceneImages {
	std::vector<GpuNode> nodes;
	std::vector<GpuPair> pairs;   // includes the reference's tail padding (Scene.cpp:335-338)
	std::vector<uint32_t> remap;  // 2 words per real pair
	uint32_t realPairs = 0;       // pairs referenced by leaves (without padding)
	uint32_t depth = 0;           // deepest leaf, root = 1
	float boundsMin[3] = {0, 0, 0};
	float boundsMax[3] = {0, 0, 0};
};

// Sparse node as produced during the build (host and device builders share it). A subtree over n
// triangles owns the slot block [slot, slot + 2n - 1): node at slot, left subtree right after it,
// right subtree after that. This numbers nodes without an atomic counter, so the build is
// deterministic under any scheduling.
struct BuildNode {
	uint32_t kind;
	uint32_t parent;
	uint32_t first, last;   // triangle range (always kept, also for inner nodes)
	uint32_t left, right;   // child slots for inner nodes
	float bounds[8];        // {-min.xyzw, max.xyzw}: one max() unions a box (Bvh2.cpp:82-126)
};
static_assert(sizeof(BuildNode) == 56, "BuildNode is shared with the device builder");

// Device SAH builder hook (bvh_build.cu): fills the sparse node array (2n slots) and the x-sorted
// triangle list exactly as the host builder would. Returns false and sets *error on failure.
typedef bool (*DeviceBvhBuilder)(const float* vertices4, uint32_t vertexCount, const uint32_t* indices, uint32_t triangleCount,
                                 std::vector<BuildNode>* nodes, std::vector<uint32_t>* sorted0, const char** error);

// The device builder itself (bvh_build.cu); uses the current CUDA device.
bool buildBvh2Device(const float* vertices4, uint32_t vertexCount, const uint32_t* indices, uint32_t triangleCount,
                     std::vector<BuildNode>* nodes, std::vector<uint32_t>* sorted0, const char** error);

// The whole scene build on the device (bvh_build.cu): the three images are produced in device
// memory in the reference's byte formats; the caller owns the pointers (cudaFree).
struct DeviceSceneImages {
	void* nodes = nullptr;   // nodeCount x 64 B
	void* pairs = nullptr;   // pairCount x 48 B (incl. tail padding)
	uint32_t* remap = nullptr;
	void* verts = nullptr;   // device copies of the input mesh (the build uploaded them anyway)
	uint32_t* indices = nullptr;
	uint32_t nodeCount = 0, pairCount = 0, realPairs = 0, remapCount = 0, depth = 0;
	float boundsMin[3] = {0, 0, 0};
	float boundsMax[3] = {0, 0, 0};
};
bool buildSceneImagesDevice(const float* vertices4, uint32_t vertexCount, const uint32_t* indices, uint32_t indexCount,
                            DeviceSceneImages* out, const char** error);

// The x86 RCPSS approximation of this host for the 2048 leading-mantissa patterns of [1,2): the
// reference's leaf-cost test uses _mm_rcp_ss (Bvh2.cpp:462-467), whose bits the device builder
// reproduces through this table (RCPSS depends on the top 11 mantissa bits only and scales exactly
// with the exponent; checked at table-build time, false if this CPU behaves differently).
bool fillRcpTable(float table[2048]);

// vertices: float4 per vertex (w ignored for intersection but, as in the reference, it takes part in
// the per-triangle bounds reduction and is harmless). indexCount must be a multiple of 3.
// threads <= 0: use all hardware threads. Returns false (and sets *error) on invalid input.
// deviceBuilder != nullptr: the SAH tree is built by it (on the GPU) instead of by the host threads.
bool buildBvh2(const float* vertices4, uint32_t vertexCount, const uint32_t* indices, uint32_t triangleCount,
               int threads, Bvh2* out, const char** error, DeviceBvhBuilder deviceBuilder = nullptr);

bool buildSceneImages(const float* vertices4, uint32_t vertexCount, const uint32_t* indices, uint32_t indexCount,
                      int threads, SceneImages* out, const char** error, DeviceBvhBuilder deviceBuilder = nullptr);

} // namespace racc_b200
