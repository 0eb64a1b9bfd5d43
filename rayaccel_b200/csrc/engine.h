// engine.h -- internal declarations shared by capi.cu, traverse.cu and raygen.cu.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace racc_b200 {

// racc::Ray (RayAccelerator.h:59-64) as two 16-byte halves, and racc::Result (:66-76).
struct DevRay { float4 a; float4 b; };          // a = origin.xyz,minT   b = dir.xyz,maxT
static_assert(sizeof(DevRay) == 32, "ray is 32 bytes");

// One ray stream of a launch; `begin` is its first ray's index in the launch-wide numbering.
struct StreamRef {
	const DevRay* rays;
	float4* results;
	uint32_t begin;
	uint32_t count;
};

struct TraceParams {
	const float4* nodes;     // 4 x float4 per inner node
	const float4* pairs;     // 3 x float4 per triangle pair
	const uint32_t* remap;
	const float4* env;       // RGBA32F texels, may be null
	uint32_t envWidth, envHeight;
	uint32_t nodeCount;
	const StreamRef* streams; // device array, nstreams entries (unused when nstreams == 1)
	uint32_t nstreams;
	uint32_t total;           // rays in the launch
	StreamRef single;         // the stream when nstreams == 1 (no indirection)
	uint32_t* cursor;         // work cursor for the persistent kernels (zeroed before launch)
	unsigned long long* counters; // racc_cuda_counters (8 x u64: rays, hits, inner, pairs, pushes, leaves, -, -) or null
	uint32_t smemNodes;       // inner nodes staged in shared memory by the persistent kernel
	const float4* tnodes;     // packed images (traverse_packed.cu): 4 x float4 per inner node,
	const float4* tpairs;     //   4 x float4 per triangle pair
	const uint32_t* perm;     // optional visiting order (launch-wide ray indices), null = arrival order
	const float4* envPairs;   // light probe as horizontally adjacent texel PAIRS: (envWidth+1) x envHeight x 32 B, or null
	const void* qnodes;       // quantised node image (variant 4): 32 B per inner node, see traverse_packed.cu
	float qOrigin[3], qCell[3]; //   its grid: plane = qOrigin + q * qCell
	const uint32_t* totalPtr; // non-null: the number of rays is read from DEVICE memory at kernel start (<= total, which then
	                          // only sizes the grid); lets a wavefront renderer enqueue wave k+1 before wave k's size is known
};

// Tunables (racc_cuda_set_variant / RACC_B200_* environment variables), see DESIGN.md section 5.
struct Tuning {
	int variant = 3;        // 3 packed-format persistent while-while with bail-out (default, traverse_packed.cu); 4 the same
	                        // kernel on 32-byte quantised nodes (one gather per node; not bit-exact at ties, opt-in);
	                        // reference-format kernels of traverse.cu kept for A/B: 0 persistent while-while,
	                        // 1 one-thread-per-ray, 2 persistent while-while with bail-out
	int blockThreads = 256; // threads per CTA
	int ctasPerSm = 5;      // 0 = as many as fit
	int smemNodes = 0;      // inner nodes staged in shared memory by TMA: -1 = as many as fit, 0 = none. On
	                        // battlefield (3.2 MB, L1/L2-resident) the un-staged instantiation measures ~4 %
	                        // faster (profiles/r01_sweep_c_unstaged_256bit.jsonl), so it is the default.
	int fetchThreshold = 16; // refill a warp when at least this many lanes are idle
	int leafBail = 4;        // variants 2, 3: leave the leaf loop when fewer lanes than this still have pairs to test
	int innerBail = 8;       // variants 2, 3: leave the inner loop when fewer lanes than this still descend
	int carveout = -1;       // shared-memory carveout percent, -1 = exactly what the CTAs need
	// ray re-binning before traversal (raysort.cu; variant 3 only)
	int sortMode = 2;        // 0 arrival order, 1 re-bin every device-resident launch, 2 auto: re-bin when the scene
	                         // is far larger than L2 (traversal is then DRAM-bound and coherence pays for the sort)
	int sortOriginBits = 5;  // Morton bits per axis of the ray origin inside the scene bounds
	int sortDirBits = 0;     // Morton bits per axis of the direction on the unit cube (0: origin only, measured best on config 5)
	int sortDirMajor = 0;    // 0 origin-major key, 1 direction-major key
	int smemStack = -1;      // variant 3: entries of every traversal stack kept in shared memory: 0, 8, 16, or -1 = auto
	                         // (16 when the scene is far larger than L2: the stacks then stop competing with the scene for
	                         // L1 lines, +10 % on config 5; on L2-resident scenes the all-local stack is 4 % faster)
	int hostZeroCopy = 0;    // HOST streams in pinned, mapped memory: 1 = the kernel reads rays / writes results over PCIe
	                         // itself (one launch, no staging copies), 0 = staged H2D / trace / D2H pipeline
	int buildDevice = 3;     // scene build (same images either way): 0 host threads; 1 SAH tree on the GPU, packing on the
	                         // host; 2 everything on the GPU (bvh_build.cu); 3 auto = 2 from kAutoDeviceBuildTriangles up
	int hostTaper = 256;     // staged HOST streams: > 0 = the last chunks of a call shrink geometrically down to this many K rays,
	                         // so that the traversal and D2H copy left exposed after the last H2D copy are short; measured on the
	                         // bench batch: e2e 1498 (off) / 1582 (64 K) / 1596 (256 K) Mrays/s (profiles/r02_call1_open_questions.md)
	// device-side Whitted renderer (whitted.cu)
	int whittedArena = 1;    // 1 = wave buffers kept and grown per calling thread instead of stream-ordered allocations per wave:
	                         // 1920x1080x4spp depth 8 in 11.2 ms instead of 24-31 ms (same file)
	int whittedCombine = 0;  // 1 = a warp sums its rays' fixed-point radiance per pixel run before the atomics (same bits);
	                         // measured neutral with the arena (11.13 vs 11.16 ms) and erratic without it: stays off
	int pathSync = 0;        // racc_cuda_path_trace: 1 = the host waits for every wave's size (round 1's scheme); 0 = wave sizes stay on
	                         // the device and the whole batch is enqueued without a host round trip
	int pathTraceCtas = 4;   // racc_cuda_path_trace, wavefront form: its traversal launches use at most this many CTAs per SM
	                         // (0 = all that fit), which leaves registers for one CTA of the OTHER lane's shading kernel beside
	                         // them: the HBM-bound shading pass then overlaps the L1-bound traversal instead of waiting for it
	int gridCtasPerSm = 0;   // internal (not a tuning key): cap of a persistent launch's grid, set per call by the renderer
	int pathStream = 0;      // racc_cuda_path_trace: 1 = one persistent kernel per batch that traces, shades and queues the paths'
	                         // next rays itself (pathstream.cu); 0 = a traversal launch and a shading launch per bounce. Same
	                         // framebuffer bits; measured 31 % slower on the 1920x1080 x 4 spp frame (3.88 vs 2.95 ms,
	                         // profiles/r02_streamed_path_tracer.md), so it stays opt-in
	int pathStreamThreshold = 28; // its refill threshold: the idle lanes of a warp also SHADE together, so fuller is better
	                         // (8 / 16 / 24 / 28 / 32 idle lanes: 5.66 / 4.69 / 4.10 / 3.88 / 3.89 ms)
};

// counterMode: 0 none, 1 rays+hits only, 2 rays, hits, inner-node and pair visits
cudaError_t launchTrace(const TraceParams& p, const Tuning& t, int counterMode, int smCount, cudaStream_t stream, int* launches);

// variant 3 (traverse_packed.cu)
cudaError_t launchTracePacked(const TraceParams& p, const Tuning& t, int counterMode, int smCount, cudaStream_t stream, int* launches);

// light probe -> texel-pair table for the packed kernel (once per environment)
cudaError_t launchPackEnv(const float4* texels, uint32_t width, uint32_t height, float4* pairsOut, cudaStream_t stream, int* launches);

// reference-format node image -> 32-byte quantised nodes on a 16-bit grid over the scene bounds (once per scene)
cudaError_t launchQuantiseNodes(const float4* nodes, uint32_t nodeCount, const float boundsMin[3], const float boundsMax[3], void* qnodes,
                                float qOrigin[3], float qCell[3], cudaStream_t stream, int* launches);

// reference-format images -> packed images, on the device (once per scene)
cudaError_t launchPackImages(const float4* nodes, uint32_t nodeCount, const float4* pairs, uint32_t pairCount,
                             float4* tnodes, float4* tpairs, cudaStream_t stream, int* launches);

// ray re-binning (raysort.cu): builds a visiting order in scratch memory (raySortScratchBytes), see there
size_t raySortScratchBytes(uint32_t total);
cudaError_t launchRaySort(const TraceParams& p, const float boundsMin[3], const float boundsMax[3], int originBits, int dirBits,
                          int dirMajor, void* scratch, int smCount, cudaStream_t stream, const uint32_t** perm, int* launches);

// stable LSD radix sort of (key, value) pairs, 8 bits per pass (raysort.cu); see there
size_t radixSortHistWords();
cudaError_t launchRadixSort(uint32_t* keys, uint32_t* vals, uint32_t* keysTmp, uint32_t* valsTmp, uint32_t* hist, uint32_t total,
                            int keyBits, int smCount, cudaStream_t stream, uint32_t** keysOut, uint32_t** valsOut, int* launches);

cudaError_t readWarpStats(unsigned long long* out8, bool reset);

cudaError_t launchGeneratePrimary(const float* camera12, uint32_t width, uint32_t height, uint32_t spp, uint32_t seed,
                                  DevRay* rays, cudaStream_t stream, int* launches);

cudaError_t launchGenerateBounce(const float4* verts, const uint32_t* indices, const DevRay* rays, const float4* results,
                                 uint32_t count, uint32_t seed, DevRay* outRays, uint32_t* outCount, uint32_t* scratch,
                                 cudaStream_t stream, int* launches);

size_t bounceScratchWords(uint32_t count);

// device-side wavefront path tracer (pathtrace.cu), see there
struct PathShadeParams {
	const DevRay* rays;       // the wave just traced: rays, results, path states (weight rgb + path index in .w)
	const float4* results;
	const float4* states;
	uint32_t count;
	const uint32_t* countPtr; // non-null: the wave's size is read from device memory (<= count, which then sizes the grid)
	uint32_t gridLimit;       // 0 = one CTA per 256 rays of `count`; else at most this many CTAs, each looping over tiles
	uint32_t depth, maxDepth; // bounce number of this wave; hits are extended while depth < maxDepth
	uint32_t seed, pixels, sampleBase;
	const uint32_t* indices;
	const float4* normals;
	const float4* triangleNormals;
	const uint16_t* triangleMaterials;
	const float4* materials;  // {r, g, b, eta} per material
	uint32_t triangleCount, materialCount;
	DevRay* outRays;          // next wave (compacted), its size in *outCount (zeroed by the caller)
	float4* outStates;
	uint32_t* outCount;
	float4* radiance;         // one slot per path of the batch (zeroed by the caller), written when the path escapes
};

cudaError_t launchPathPrimary(const float* camera12, uint32_t width, uint32_t height, uint32_t sampleBase, uint32_t firstPath, uint32_t count,
                              uint32_t seed, DevRay* rays, float4* states, cudaStream_t stream, int* launches);
cudaError_t launchPathShade(const PathShadeParams& p, cudaStream_t stream, int* launches);
cudaError_t launchPathAccumulate(const float4* radiance, uint32_t pixels, uint32_t spp, float4* framebuffer, cudaStream_t stream,
                                 int* launches);

// the same path tracer as one persistent kernel per launch (pathstream.cu), see there
struct PathStreamParams {
	const float4* tnodes;     // packed scene images and light probe of the traversal (TraceParams)
	const float4* tpairs;
	const uint32_t* remap;
	const float4* envPairs;
	uint32_t envWidth, envHeight;
	const uint32_t* indices;  // shading data (PathShadeParams)
	const float4* normals;
	const float4* triangleNormals;
	const uint16_t* triangleMaterials;
	const float4* materials;
	uint32_t triangleCount, materialCount;
	const float* camera12;
	uint32_t width, pixels, sampleBase, seed, maxDepth;
	uint32_t firstPath, paths; // this launch traces paths firstPath .. firstPath + paths - 1 of the batch (< pathStreamMaxPaths())
	float4* queue;            // capacity x 48 B
	uint32_t* flags;          // capacity words holding epochs of earlier launches (or zero)
	uint32_t capacity;        // >= paths * maxDepth
	uint32_t epoch;           // differs from every value in flags
	uint32_t* ctrl;           // 4 words, zeroed by the launch
	float4* radiance;         // one slot per path of the batch (zeroed by the caller)
	unsigned long long* depthRays; // [maxDepth + 1] += rays traced per bounce, or null
	unsigned long long* counters;  // racc_cuda_counters: += rays, hits; or null
	int smemStack;            // > 0: the tops of the traversal stacks in shared memory (scenes far larger than L2)
};

uint32_t pathStreamMaxPaths();
cudaError_t launchPathStream(const PathStreamParams& p, const Tuning& t, int smCount, cudaStream_t stream, int* launches);

// device-side Whitted renderer (whitted.cu), see there
struct WhittedShadeParams {
	const DevRay* rays;       // the wave just traced: rays, results, states (weight rgb + pixel in .w)
	const float4* results;
	const float4* states;
	uint32_t count, depth, maxDepth;
	const uint32_t* indices;
	const float4* normals;
	const float4* triangleNormals;
	uint32_t triangleCount;
	DevRay* outRays;          // next wave (capacity 2 * count), its size in *outCount (zeroed by the caller)
	float4* outStates;
	uint32_t* outCount;
	unsigned long long* accumulators; // 3 per pixel, 32.32 fixed-point radiance sums
	bool combine;             // Tuning::whittedCombine
};

cudaError_t launchWhittedPrimary(const float* camera12, uint32_t width, uint32_t height, uint32_t sampleBase, uint32_t firstPath, uint32_t count,
                                 uint32_t seed, DevRay* rays, float4* states, cudaStream_t stream, int* launches);
cudaError_t launchWhittedShade(const WhittedShadeParams& p, cudaStream_t stream, int* launches);
cudaError_t launchWhittedFinish(const unsigned long long* accumulators, uint32_t pixels, float4* framebuffer, cudaStream_t stream, int* launches);

} // namespace racc_b200
