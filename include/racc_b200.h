/* racc_b200.h -- the drop-in boundary: C-ABI of the B200 ray-intersection engine.
 *
 * Plain C types only (pointers, sizes, opaque handles); no C++ or torch types cross it. The
 * C++ `racc::` API in include/RayAccelerator.h (same names and layouts as the reference's
 * /root/reference/RayAccelerator/RayAccelerator.h:25-116) is a thin caller of these entry points,
 * and so are the Python mirror in rayaccel_b200/ and bench.py. Every entry point below names the
 * reference interface it replaces. 0 = success unless stated; on failure racc_cuda_last_error()
 * returns a thread-local message. Nothing here falls back to the CPU: without a CUDA device every
 * compute entry point fails.
 *
 * Implemented in rayaccel_b200/csrc/ (capi.cu, traverse.cu, scene_build.cpp); built into
 * rayaccel_b200/libracc_b200.so by __graft_entry__.build().
 */
#ifndef RACC_B200_H
#define RACC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RACC_CUDA_ABI_VERSION 2

typedef struct racc_cuda_scene racc_cuda_scene;   /* replaces racc::Scene        (Scene.h:15-21) */
typedef struct racc_cuda_env racc_cuda_env;       /* replaces racc::Environment  (Environment.h:16-24) */

/* One ray stream handed to the tester: replaces racc_internal::GpuRayStream
 * (RayAccelerator.cpp:36-40) = racc::RayStream + its two cl_mem views. rays: count x 32 B
 * racc::Ray; results: count x 16 B racc::Result, index-parallel to rays. */
typedef struct {
	const void* rays;
	void* results;
	uint32_t count;
	uint32_t flags; /* RACC_CUDA_STREAM_* */
} racc_cuda_stream_desc;

#define RACC_CUDA_STREAM_DEVICE 0u /* rays/results are device pointers on the current device */
#define RACC_CUDA_STREAM_HOST 1u   /* host pointers; the engine stages them (H2D, trace, D2H) */

typedef struct {
	uint32_t node_count;      /* inner nodes, 64 B each */
	uint32_t pair_count;      /* triangle pairs incl. tail padding, 48 B each */
	uint32_t real_pair_count; /* pairs referenced by leaves */
	uint32_t remap_count;     /* 4 B words */
	uint32_t depth;
	uint32_t triangle_count;
	float bounds_min[3];
	float bounds_max[3];
} racc_cuda_scene_info;

/* Aggregate visit counters of a traced batch: the inputs of the algorithmic-bytes roofline
 * (B_ray = 32 + 16 + 64*inner + 48*pairs + 4*[hit] + 64*[miss]). */
typedef struct {
	uint64_t rays;
	uint64_t hits;
	uint64_t inner_nodes;
	uint64_t pairs_tested;
	uint64_t stack_pushes;  /* far children pushed (= entries popped): with the two above, the gather instructions a ray costs */
	uint64_t leaf_visits;
	uint64_t reserved[2];   /* 64 bytes in all: what a caller of racc_cuda_trace_counted allocates and zeroes */
} racc_cuda_counters;

/* replaces racc::init() + the OpenCL device pick (RayAccelerator.cpp:417-423,463-478;
 * Renderer/main.cpp:68-115). Sets the calling THREAD's device set to the n CUDA devices named (devices == NULL or
 * n == 0: the current CUDA device), initialising each on first use. The first device is the one the thread is bound to:
 * its DEVICE streams, cudaStream_t handles and renderer calls live there. Scenes, environments and shading data created
 * by the thread are replicated on every device of its set (built once, copied peer to peer), and the HOST streams of
 * one racc_cuda_trace call are dealt over all of them. A thread that never calls this is bound to the current CUDA
 * device. Nothing is process-global: threads (and racc::Contexts) on different devices do not disturb each other; a
 * scene can be traced on any device it has a copy on. */
int racc_cuda_init(const int* devices, int n);

/* The calling thread's device set (bound device first): writes up to `capacity` CUDA ordinals, returns the set's size. */
int racc_cuda_current_devices(int* devices, int capacity);

/* Releases the calling thread's per-device scratch (staging pipelines of HOST streams: three lanes of device buffers and
 * streams per device; renderer lanes and wave buffers). Call it when a thread that traced HOST streams or rendered ends --
 * racc::destroy(Context*) does for its submitter threads; the counterpart of clReleaseCommandQueue at
 * RayAccelerator.cpp:774-779. */
void racc_cuda_thread_release(void);

/* CUDA devices visible, or -1 if the runtime cannot be initialised (no fallback exists). */
int racc_cuda_device_count(void);

int racc_cuda_abi_version(void);

const char* racc_cuda_last_error(void);

/* replaces racc::createScene()'s GPU branch (Scene.cpp:216-349): SAH build, pair merge, node
 * packing -- on the device by default (bvh_build.cu), on the host threads on request; same images. verts4: nverts x float4 (w ignored); indices: nindices (multiple of 3).
 * The caller keeps ownership of its arrays. NULL on failure. */
racc_cuda_scene* racc_cuda_scene_create(const float* verts4, uint32_t nverts, const uint32_t* indices, uint32_t nindices);

/* Upload prebuilt images (same byte formats) instead of building; used to trace the reference's
 * own images and by the multi-GPU path to replicate one build. */
racc_cuda_scene* racc_cuda_scene_create_from_images(const void* nodes, uint32_t node_count, const void* pairs,
                                                    uint32_t pair_count, const uint32_t* remap, uint32_t remap_count);

/* Host-only half of racc_cuda_scene_create (no CUDA call is made): builds the three images in
 * host memory. Lets the scene build be checked against the reference builder on a machine without
 * a GPU, and lets one build be uploaded to several devices. */
typedef struct racc_cuda_host_images racc_cuda_host_images;
racc_cuda_host_images* racc_cuda_build_images(const float* verts4, uint32_t nverts, const uint32_t* indices, uint32_t nindices);
int racc_cuda_host_images_get_info(const racc_cuda_host_images* images, racc_cuda_scene_info* info);
int racc_cuda_host_images_copy(const racc_cuda_host_images* images, void* nodes, void* pairs, uint32_t* remap);
void racc_cuda_host_images_destroy(racc_cuda_host_images* images);

/* replaces racc::destroy(Scene*) (Scene.cpp:359-372) */
void racc_cuda_scene_destroy(racc_cuda_scene* scene);

int racc_cuda_scene_get_info(const racc_cuda_scene* scene, racc_cuda_scene_info* info);

/* Copies the host images out (any pointer may be NULL): what the reference keeps in
 * scene->gpuNodes / gpuTriangles / gpuTriangleIndices (Scene.cpp:342-346). */
int racc_cuda_scene_download(const racc_cuda_scene* scene, void* nodes, void* pairs, uint32_t* remap);

/* replaces racc::createEnvironment() (Environment.cpp:13-60): rgba = width*height RGBA32F. */
racc_cuda_env* racc_cuda_env_create(const float* rgba, uint32_t width, uint32_t height);

/* replaces racc::destroy(Environment*) (Environment.cpp:62-67) */
void racc_cuda_env_destroy(racc_cuda_env* env);

/* replaces the body of gpuWorkerThread: 7x clSetKernelArg + clEnqueueNDRangeKernel
 * (RayAccelerator.cpp:377-401) for one OR MORE ray streams in a single launch. env may be NULL
 * (misses then return r=g=b=0). cuda_stream: a cudaStream_t (NULL = default stream).
 * Asynchronous. DEVICE streams of one call are traced by ONE launch. HOST streams
 * (pinned memory recommended) are packed into staging chunks (several small streams share one
 * chunk and one launch) that alternate over internal CUDA streams so H2D, traversal and D2H overlap; cuda_stream waits for them, so everything is complete when
 * racc_cuda_sync(cuda_stream) returns. */
int racc_cuda_trace(racc_cuda_scene* scene, racc_cuda_env* env, const racc_cuda_stream_desc* streams,
                    uint32_t nstreams, void* cuda_stream);

/* Same, additionally accumulating into *device_counters (a device pointer to a racc_cuda_counters
 * that the caller zeroed). detail == 0: rays and hits only (free: one atomic per warp; this is the
 * per-frame hit count the multi-GPU reduction sums, and what racc::Stats reports). detail != 0:
 * also inner-node and pair visits (slower; used for the roofline accounting only) and, from the default
 * kernel (variant 3), stack pushes and leaf visits. */
int racc_cuda_trace_counted(racc_cuda_scene* scene, racc_cuda_env* env, const racc_cuda_stream_desc* streams,
                            uint32_t nstreams, void* cuda_stream, void* device_counters, int detail);

/* The per-frame hit reduction -- the one collective on the path. Every traversal launch without a
 * caller-supplied counter record adds its rays and hits to its device's frame record. This call sums the records over
 * the calling thread's device set (ncclAllReduce, one communicator per device, this process) or -- for a thread whose set
 * is the one device the process joined a rank communicator with (below) -- over all ranks, in which case EVERY rank must
 * call; zeroes them for the next frame; and returns the totals -- what
 * racc::Stats.raysTraced is on the reference's single host (RayAccelerator.cpp:200,372,755-758). totals == NULL: no
 * wait, the sums stay on the device, ordered on cuda_stream. Launches to be counted must have completed or have been
 * enqueued on cuda_stream. NCCL is bound at run time (dlopen) and only when more than one GPU takes part. */
int racc_cuda_frame_reduce(racc_cuda_counters* totals, void* cuda_stream);

/* Multi-process jobs (one process per GPU, e.g. under torchrun): rank 0 makes an id (128 bytes, ncclGetUniqueId), the
 * application hands it to every rank, every rank joins. racc_cuda_frame_reduce then spans all ranks. */
int racc_cuda_comm_unique_id(void* id128);
int racc_cuda_comm_init_rank(const void* id128, int rank, int nranks);
void racc_cuda_comm_destroy(void);
/* Ranks of the communicator joined (1 without one); *rank (may be NULL) receives this process' rank. */
int racc_cuda_comm_ranks(int* rank);

/* A ray-sharded frame's hit buffer made whole on every rank: every rank passes the racc::Result slice
 * of its contiguous ray range -- rays_per_rank records of 16 bytes, the same count on every rank (pad the last slice) -- and
 * receives all slices in rank order in device_all_results (nranks * rays_per_rank records; may alias the slice of this rank in
 * place). ncclAllGather over the rank communicator, enqueued on cuda_stream; a copy when there is only one rank. Results stay
 * index-parallel to the frame's rays, as the reference's clients expect (RayAccelerator.h:66-83). */
int racc_cuda_gather_results(const void* device_results, uint32_t rays_per_rank, void* device_all_results, void* cuda_stream);

/* Pinned (page-locked) host memory for ray streams: replaces the 4 KiB-aligned stream slab that the
 * reference wraps in zero-copy CL_MEM_USE_HOST_PTR buffers (RayAccelerator.cpp:532-568,636-645).
 * A discrete GPU has no zero-copy path; pinned memory is what lets HOST streams move at PCIe rate.
 * NULL on failure. */
void* racc_cuda_host_alloc(size_t bytes);
void racc_cuda_host_free(void* ptr);

/* One CUDA stream per submitter thread: replaces the per-thread cl_command_queue
 * (RayAccelerator.cpp:711-717, released :774-779). Returns a cudaStream_t, NULL on failure. */
void* racc_cuda_stream_create(void);
void racc_cuda_stream_destroy(void* cuda_stream);

/* replaces clFinish(queue) (RayAccelerator.cpp:403) */
int racc_cuda_sync(void* cuda_stream);

/* Number of engine kernels launched by this process so far (bench.py's gpu_launches). */
uint64_t racc_cuda_launch_count(void);

/* Diagnostics (not in the reference): warp-level loop statistics accumulated by the COUNTED traversal
 * launches (racc_cuda_trace_counted with detail != 0) of the bail-out kernel variant: out8[0] outer
 * rounds, [1] inner-loop iterations, [2] leaf-loop iterations, [3]/[4] lanes active summed over those
 * iterations, [5] refills. Synchronises the device. reset != 0 zeroes them afterwards. */
int racc_cuda_debug_warp_stats(uint64_t* out8, int reset);

/* Diagnostics (not in the reference): the 2048-entry table through which the device scene builder
 * reproduces the host's _mm_rcp_ss (the reference's leaf-cost test, Bvh2.cpp:462-467): out2048[i] =
 * RCPSS(1 + i/2048). Host only, no CUDA call. Returns 0 when this CPU's RCPSS follows the table
 * model (result depends on the top 11 mantissa bits, scales exactly with the exponent), else 1 --
 * the device builder then declines and scenes are built on the host threads. */
int racc_cuda_debug_rcp_table(float* out2048);

/* Engine tuning knob (not in the reference): which traversal kernel variant racc_cuda_trace uses.
 * 0 = default. See DESIGN.md section 5. Returns the previous value. */
int racc_cuda_set_variant(int variant);

/* Launch-shape knobs for benchmark sweeps: key 0 variant, 1 threads per CTA (128/256/512/1024),
 * 2 CTAs per SM (0 = as many as fit), 3 inner nodes staged in shared memory (-1 = as many as
 * fit, 0 = none), 4 refill threshold (idle lanes per warp), 5 leaf-loop bail-out, 6 shared-memory
 * carve-out percent, 7 inner-loop bail-out, 8 ray re-binning (0 off, 1 on, 2 auto: scenes far larger than L2), 9 / 10
 * Morton bits per axis of the re-binning key (origin / direction), 11 direction-major key, 12 scene build (0 host,
 * 1 SAH tree on the device + host packing, 2 all on the device, 3 auto), 13 traversal-stack entries kept in shared memory
 * (0, 8, 16, -1 auto: 16 for scenes far larger than L2), 14 HOST streams in pinned memory read by the kernel itself
 * instead of being staged (0 staged = default, 1 zero-copy), 15 / 16 racc_cuda_whitted_trace only: 15 wave buffers kept
 * and grown per calling thread instead of allocated per wave (default 1: 11 ms instead of 24-31 ms per 1080p x 4 spp frame),
 * 16 a warp sums its rays' fixed-point radiance per pixel before the atomics (same bits; default 0: neutral); 17 staged HOST streams: the
 * last chunks of a call shrink geometrically down to this many K rays (0 = off, default 256); 18 racc_cuda_path_trace waits
 * for every wave's size on the host (1) instead of leaving the sizes on the device (0, default); 19 racc_cuda_path_trace as one
 * persistent kernel per batch that traces, shades and queues the paths' next rays itself (1) instead of one traversal and one
 * shading launch per bounce (0): same framebuffer bits either way; 20 racc_cuda_path_trace's traversal launches use at most this
 * many CTAs per SM (0 = all that fit; default 4 of 5: the other lane's shading pass then runs beside them). Returns the previous value. Also settable through
 * RACC_B200_VARIANT / _BLOCK / _CTAS_PER_SM / _SMEM_NODES / _FETCH_THRESHOLD / _LEAF_BAIL / _INNER_BAIL / _SORT /
 * _SORT_ORIGIN_BITS / _SORT_DIR_BITS / _SORT_DIR_MAJOR / _BUILD_DEVICE / _SMEM_STACK / _HOST_ZERO_COPY / _WHITTED_ARENA / _WHITTED_COMBINE / _HOST_TAPER / _PATH_SYNC / _PATH_STREAM / _PATH_TRACE_CTAS. Variant 3 (default) is the packed-format kernel;
 * 0-2 are the reference-format kernels kept for A/B. */
int racc_cuda_set_tuning(int key, int value);

/* ---- synthetic ray streams for the benchmark; not on the hot path ---- */

/* Camera as Renderer/Camera.cpp:13-25 builds it; rays as Camera.cpp:55-114 (minT 0, maxT 1e6). */
typedef struct {
	float origin[3];
	float view[3];
	float right[3];
	float up[3];
} racc_cuda_camera;

/* Primary rays for a width x height grid, `spp` samples per pixel, written to device memory
 * (count = width*height*spp, sample-major). jitter_seed == 0: pixel centres. */
int racc_cuda_generate_primary(const racc_cuda_camera* camera, uint32_t width, uint32_t height, uint32_t spp,
                               uint32_t jitter_seed, void* device_rays, void* cuda_stream);

/* One diffuse bounce (Renderer/PathTracingRenderer.cpp:405-422): for every HIT in results[],
 * a cosine-hemisphere ray about the geometric normal flipped against the incoming direction,
 * origin = hit + 1e-4*n, minT 1e-3, maxT 1e6; compacted, in arrival order, into device_out_rays.
 * *device_out_count (device uint32, zeroed by the caller) receives the number written. */
int racc_cuda_generate_bounce(const racc_cuda_scene* scene, const void* device_rays, const void* device_results,
                              uint32_t count, uint32_t seed, void* device_out_rays, uint32_t* device_out_count,
                              void* cuda_stream);

/* ---- device-side wavefront path tracer ----
 * The reference's example client (Renderer/PathTracingRenderer.cpp) shades on the host: every bounce crosses the
 * spawn/shade callbacks of RayAccelerator.h:85-93, i.e. PCIe on a discrete GPU. These entry points run the same
 * estimator with the shading on the device (rayaccel_b200/csrc/pathtrace.cu), so rays, hit records and path state stay
 * in HBM from the camera to the framebuffer. They are an addition beside the drop-in API, not a replacement of it. */

typedef struct racc_cuda_shading racc_cuda_shading;

/* Host arrays of Renderer/SceneData.h:13-30 as main.cpp:154-180 loads them; copied to the device. */
typedef struct {
	const float* normals4;              /* per vertex, 4 floats (SceneData::normals) */
	uint32_t vertex_count;
	const float* triangle_normals4;     /* per triangle, 4 floats (SceneData::triangleNormals) */
	const uint16_t* triangle_materials; /* per triangle (SceneData::triangleMaterials) */
	uint32_t triangle_count;
	const float* materials_ke4;         /* per material {r, g, b, eta}: ReflectiveDiffuseMaterial(k, eta), Materials.cpp:32-37 */
	uint32_t material_count;
} racc_cuda_shading_desc;

/* NULL on failure (racc_cuda_last_error). */
racc_cuda_shading* racc_cuda_shading_create(const racc_cuda_shading_desc* desc);
void racc_cuda_shading_destroy(racc_cuda_shading* shading);

#define RACC_CUDA_FRAMEBUFFER_HOST 1u /* framebuffer4 is a host pointer (copied in, accumulated, copied back, synchronised) */

typedef struct {
	uint32_t width, height;  /* every pixel is rendered (the reference's TiledRenderer covers whole 128-pixel tiles only) */
	uint32_t sample_base;    /* index of the first sample: samples sample_base .. sample_base+spp-1 (ranks of a multi-GPU job
	                            take disjoint ranges and sum their framebuffers) */
	uint32_t spp;            /* paths per pixel; the reference renders one per frame and accumulates (TiledRenderer.cpp:40-48) */
	uint32_t max_depth;      /* SceneData::maxDepth: a path is extended while its depth < max_depth (PathTracingRenderer.cpp:126) */
	uint32_t seed;           /* 0 = primary rays through pixel centres */
	uint32_t batch_spp;      /* samples traced together, 0 = about 128 M paths (128 B of device memory per path; at most an eighth of the device's memory) */
	uint32_t flags;
} racc_cuda_path_desc;

/* Adds the radiance sums of `spp` paths per pixel to framebuffer4 (width*height x {r,g,b,unused} floats; a DEVICE
 * pointer unless RACC_CUDA_FRAMEBUFFER_HOST), like the reference's frameBuffer (PathTracingRenderer.cpp:540-543): sums,
 * not means. wave_rays (host, may be NULL): [max_depth+1] counters, += rays traced at each depth; their total is
 * Stats::raysTraced of the equivalent render() calls. The scene must have been created from triangles (not from images).
 * Work is enqueued on cuda_stream without a host round trip (wave sizes stay on the device); the call waits -- once per
 * batch -- only when wave_rays is given, and a device framebuffer is complete once the stream has been synchronised. Results are bit-reproducible and equal oracle_path_trace's. 0 on success. */
int racc_cuda_path_trace(racc_cuda_scene* scene, racc_cuda_env* env, const racc_cuda_shading* shading, const racc_cuda_camera* camera,
                         const racc_cuda_path_desc* desc, float* framebuffer4, uint64_t* wave_rays, void* cuda_stream);

/* The reference's second example client, the Whitted renderer (Renderer/WhittedRenderer.cpp:136-676), with the shading on
 * the device (rayaccel_b200/csrc/whitted.cu): same arguments and framebuffer meaning as racc_cuda_path_trace. Every hit
 * adds direct light and spawns a reflection and a refraction ray; the ray tree is walked breadth-first, radiance is summed
 * per pixel in 32.32 fixed point (integer atomics: order-independent), then added to framebuffer4. Uses the normals and
 * triangle normals of `shading` (the renderer has one hard-wired material). desc->batch_spp 0 = about 4 M primary rays per
 * batch. Results are bit-reproducible and equal oracle_whitted_trace's. 0 on success. */
int racc_cuda_whitted_trace(racc_cuda_scene* scene, racc_cuda_env* env, const racc_cuda_shading* shading, const racc_cuda_camera* camera,
                            const racc_cuda_path_desc* desc, float* framebuffer4, uint64_t* wave_rays, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif
