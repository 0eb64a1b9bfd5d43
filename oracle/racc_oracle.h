/* racc_oracle.h -- TEST INFRASTRUCTURE ONLY (the parity checker).
 *
 * CPU restatement, in pinned IEEE-754 binary32 arithmetic, of the reference's
 * ray-intersection hot path: the OpenCL `traversal` kernel in
 * /root/reference/RayAccelerator/Kernels.h:9-242 operating on the scene images
 * built by RayAccelerator/Scene.cpp:223-346.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library. The product (rayaccel_b200/) never does.
 *
 * PARITY PIN STATUS
 *   - scene build (nodes / pairs / remap):  pinned against the UNMODIFIED reference builder
 *     compiled into oracle/_ref/libracc_ref.so (tests/test_scene_build.py).
 *   - light-probe lookup: pinned (float tolerance) against the reference's own CPU
 *     racc_internal::sample() (Environment.h:27-82) via oracle/_ref.
 *   - BVH traversal + triangle-pair test: the reference has NO golden vectors or tests, no
 *     OpenCL runtime exists here, and its CPU path is inside Embree 2.7.0 (binary-only,
 *     macOS/Windows). What CAN run here is the reference's own kernel SOURCE: oracle/Makefile
 *     target `kernel` extracts the `traversal` text from Kernels.h:9-242 at build time, compiles
 *     it unmodified as C++ over a small OpenCL C language shim (oracle/ref_shim/opencl_c.h:
 *     vector types, swizzles, built-ins) into oracle/_ref/libkernel_ref.so and runs it one
 *     work-item at a time. The restatement is pinned, bit for bit (ids, t, u, v, r, g, b),
 *     against that on the KAT scenes, the golden rays, random and edge-case rays and on
 *     reference-BUILT images (tests/test_oracle_kat.py::test_oracle_matches_reference_kernel_*;
 *     tests/golden/battlefield_rays.npz holds that library's outputs). So control flow, traversal
 *     order, tie rules, edge codes and the operation sequence are the reference's own; what stays
 *     a MODEL is only the bit behaviour of the built-ins the reference leaves implementation-
 *     defined (-cl-fast-relaxed-math, native_*, read_imagef filtering; listed below), which the
 *     shim and this file define identically. Additional anchors: (a) an independent brute-force
 *     fp64 Moller-Trumbore arbiter over the ORIGINAL triangles (oracle_brute_f64), (b) hand-built
 *     known-answer cases for every tie / boundary rule (tests/test_oracle_kat.py).
 *
 * PINNED ARITHMETIC (the "ideal reading" of the OpenCL source; the CUDA kernel uses the
 * identical operation sequence, so kernel-vs-oracle is bit-exact in IDs *and* t,u,v,r,g,b):
 *   - every OpenCL `mad(a,b,c)`            -> fmaf(a,b,c)  (single rounding)
 *   - every other written  a*b, a+b, a-b   -> separately rounded (compiled -ffp-contract=off)
 *   - dot(a,b)                             -> fmaf(a.z,b.z, fmaf(a.y,b.y, a.x*b.x))
 *   - 1.0f/x, native_recip(x)              -> IEEE division 1.0f/x, round-to-nearest-even
 *   - native_rsqrt(x)                      -> 1.0f / sqrtf(x), both correctly rounded
 *   - acos                                 -> oracle_acosf() below, a fixed fdlibm-style
 *                                             polynomial evaluated in binary32 (libm acosf is
 *                                             not bit-reproducible across CPU and GPU)
 *   - read_imagef(CLK_FILTER_LINEAR | CLK_ADDRESS_CLAMP_TO_EDGE | NORMALIZED) -> the OpenCL 1.2
 *     spec formula (section 8.2) in binary32, see oracle_env_sample()
 *   - flush-to-zero / denormals-are-zero ON (RayAccelerator.cpp:417-420, Threading.h:77-79)
 *   - fmin/fmax: IEEE minNum/maxNum with -0 < +0 (what CUDA's FMNMX implements)
 *   - unary minus on a float (Kernels.h:65-66: -dot(R,e1), -dot(R,e3)) -> a flip of the sign bit, AFTER the dot product was
 *     rounded. gcc folds -fma(a,b,c) into one fnmsub, whose exact-cancellation zero is +0 where -(+0) is -0; that sign decides
 *     who owns a shared edge. nvcc keeps the negation, so the B200 had it right and this file did not until the randomised
 *     campaign (tests/fuzz/fuzz_gpu.py) met a ray through an edge of an integer-grid mesh.
 */
#ifndef RACC_ORACLE_H
#define RACC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* racc::Ray, RayAccelerator.h:59-64 */
typedef struct {
	float origin[3];
	float minT;
	float dir[3];
	float maxT;
} oracle_ray;

/* racc::Result, RayAccelerator.h:66-76. a,b,c = {t,u,v} on a hit, {r,g,b} on a miss. */
typedef struct {
	uint32_t triangle;
	float a, b, c;
} oracle_result;

#define ORACLE_INVALID_TRIANGLE 0xffffffffu

/* The three scene images of Scene.cpp:342-346 plus the light probe of Environment.cpp:36-50. */
typedef struct {
	const float* nodes;      /* 16 floats (64 B) per inner node, Scene.cpp:73-78 */
	uint32_t node_count;
	const float* pairs;      /* 12 floats (48 B) per triangle pair, Scene.cpp:83-87 */
	uint32_t pair_count;
	const uint32_t* remap;   /* pair-triangle -> original index | edge<<30, Scene.cpp:132-133 */
	uint32_t remap_count;
	const float* env;        /* RGBA32F, may be NULL (then misses return r=g=b=0) */
	uint32_t env_width, env_height;
} oracle_scene;

/* Per-ray visit counters: the inputs of the algorithmic-bytes formula. */
typedef struct {
	uint16_t inner;     /* inner nodes fetched */
	uint16_t pairs;     /* triangle pairs tested */
	uint16_t max_stack; /* deepest stack head reached */
	uint16_t hit;       /* 1 = hit, 0 = miss */
	uint16_t pushes;    /* far children pushed (= entries popped): the stack traffic of the ray */
	uint16_t leaves;    /* leaves visited */
} oracle_counters;

/* Kernels.h:141-242 for rays[0..count). counters may be NULL. threads<=0 -> all cores. */
int oracle_traverse(const oracle_scene* scene, const oracle_ray* rays, uint32_t count,
                    oracle_result* results, oracle_counters* counters, int threads);

/* CPU BASELINE only (not the checker): the same traversal with an AVX2 node test (1 ray x 2 boxes),
 * multi-threaded; bit-identical to oracle_traverse for rays with minT >= 0. */
int oracle_traverse_avx2(const oracle_scene* scene, const oracle_ray* rays, uint32_t count, oracle_result* results, int threads);

/* online cores, as used when threads<=0 */
int oracle_hardware_threads(void);

/* Kernels.h:213-221 alone: radiance for direction d (already epsilon-clamped by the caller). */
void oracle_env_sample(const float* env, uint32_t width, uint32_t height, const float d[3], float rgb[3]);

/* pinned acos used by the miss path */
float oracle_acosf(float x);

/* Independent arbiter: brute-force double-precision Moller-Trumbore over the original
 * triangles, no BVH, no pairs. For each ray: t_min (or +inf on miss) and one argmin id. */
int oracle_brute_f64(const float* verts4, uint32_t nverts, const uint32_t* indices, uint32_t ntris,
                     const oracle_ray* rays, uint32_t count, double* t_min, uint32_t* tri_min, int threads);

/* fp64 hit distance of ONE given triangle for each ray (+inf if that triangle is missed);
 * used to decide whether a reported id lies in the tie set {tri : t <= t_min*(1+eps)}. */
int oracle_tri_t_f64(const float* verts4, const uint32_t* indices, const oracle_ray* rays,
                     const uint32_t* tri_ids, uint32_t count, double* t_out);

/* Structural digest of a node/pair/remap image, independent of node and pair numbering
 * (the reference's numbering depends on thread timing, Bvh2.cpp:489,511-532). Walks from inner
 * node 0. digest[0]=inner nodes reached, [1]=leaves, [2]=pairs in leaves, [3]=max depth,
 * [4]=singleton pairs, [5..11]=leaf histogram for 1..7 pairs, [12]=order-dependent 64-bit
 * FNV-1a hash over (DFS order) child boxes, leaf pair bytes and remap words. Returns 0. */
int oracle_scene_digest(const oracle_scene* scene, uint64_t digest[16]);

/* ---- wavefront path tracer; checker for rayaccel_b200/csrc/pathtrace.cu ---- */

/* What the reference's example renderer shades with (Renderer/SceneData.h:13-30, main.cpp:160-180). */
typedef struct {
	const uint32_t* indices;            /* 3 per triangle */
	uint32_t triangle_count;
	const float* normals4;              /* per vertex, 4 floats */
	const float* triangle_normals4;     /* per triangle, 4 floats */
	const uint16_t* triangle_materials; /* per triangle */
	const float* materials_ke4;         /* per material {r,g,b,eta}: ReflectiveDiffuseMaterial, Materials.cpp:32-37 */
	uint32_t material_count;
} oracle_shading;

/* Camera::lookAt's outputs (Camera.cpp:13-25) */
typedef struct { float origin[3], view[3], right[3], up[3]; } oracle_camera;

/* ReflectiveDiffuseMaterial::sample8 (Materials.cpp:39-151) for one lane: rnd in [0,1]^3, shading normal,
 * wo = -ray direction -> sampled direction wi and the path-weight factor. Exact 1/x and 1/sqrt where the
 * reference uses the 12-bit _mm256_rcp_ps / _mm256_rsqrt_ps. */
void oracle_material_sample(const float ke[4], const float rnd[3], const float normal[3], const float wo[3], float wi[3], float color[3]);

/* `spp` paths per pixel (samples sample_base .. sample_base+spp-1) of at most max_depth bounces each
 * (PathTracingRenderer.cpp:72-566), radiance of escaping paths ADDED to framebuffer4 (width*height x 4
 * floats, .w untouched) sample by sample in ascending order. wave_rays (may be NULL): [max_depth+1]
 * counters, += rays traced at each depth. Random numbers: counter-based hash of (pixel, sample, depth,
 * seed) -- see racc_oracle.c. seed 0 = pixel centres for the primary rays. */
int oracle_path_trace(const oracle_scene* scene, const oracle_shading* shading, const oracle_camera* camera, uint32_t width,
                      uint32_t height, uint32_t sample_base, uint32_t spp, uint32_t max_depth, uint32_t seed, float* framebuffer4,
                      uint64_t* wave_rays, int threads);

/* The reference's Whitted renderer (Renderer/WhittedRenderer.cpp:136-676) per pixel sample: direct light from a
 * fixed direction at every hit, weight * 0.3 per bounce, reflection + refraction children while a weight channel
 * exceeds 0.01 and depth < max_depth, probe radiance on misses. Uses shading->indices / normals4 /
 * triangle_normals4 only (the renderer has one hard-wired material). The radiance of a call is summed in 32.32
 * fixed point per pixel (order-independent) and then ADDED to framebuffer4. wave_rays as oracle_path_trace. */
int oracle_whitted_trace(const oracle_scene* scene, const oracle_shading* shading, const oracle_camera* camera, uint32_t width,
                         uint32_t height, uint32_t sample_base, uint32_t spp, uint32_t max_depth, uint32_t seed, float* framebuffer4,
                         uint64_t* wave_rays, int threads);

#ifdef __cplusplus
}
#endif
#endif
