/* racc_oracle.c -- TEST INFRASTRUCTURE ONLY (the parity checker). See racc_oracle.h for the
 * scope, the pin status (traversal pinned to the reference's own kernel source run on the CPU, oracle/_ref/libkernel_ref.so) and
 * the pinned-arithmetic rules. Every function cites the reference lines it restates; paths are
 * relative to /root/reference/.
 *
 * Build: oracle/Makefile (gcc -O2 -mfma -ffp-contract=off, pthreads). */
#include "racc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <immintrin.h>
#include <xmmintrin.h>
#include <pmmintrin.h>
#include <pthread.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------ */
/* bit helpers                                                                                 */

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* Reference threads run with FTZ+DAZ (Threading.h:77-79, RayAccelerator.cpp:417-420). */
static inline unsigned ftz_on(void) {
	unsigned saved = _mm_getcsr();
	_MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
	_MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
	return saved;
}

static inline float flush(float x) {
	uint32_t u = f2u(x);
	return (u & 0x7f800000u) ? x : u2f(u & 0x80000000u);
}

/* minNum / maxNum with -0 < +0 and subnormal inputs flushed: the semantics of the GPU's
 * min.ftz.f32 / max.ftz.f32, which OpenCL fmin/fmax (Kernels.h:125-129) map onto. */
static inline float pmin(float a, float b) {
	a = flush(a); b = flush(b);
	if (a != a) return b;
	if (b != b) return a;
	if (a < b) return a;
	if (b < a) return b;
	return u2f(f2u(a) | f2u(b)); /* equal: only +-0 can differ, prefer -0 */
}
static inline float pmax(float a, float b) {
	a = flush(a); b = flush(b);
	if (a != a) return b;
	if (b != b) return a;
	if (a > b) return a;
	if (b > a) return b;
	return u2f(f2u(a) & f2u(b)); /* equal: prefer +0 */
}

typedef struct { float x, y, z; } v3;

/* dot(a,b), association pinned (racc_oracle.h) */
static inline float dot3(v3 a, v3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }

/* mad_cross, Kernels.h:23-25:  mad(e1.y,e2.z, -e1.z*e2.y), ... */
static inline v3 mad_cross(v3 e1, v3 e2) {
	v3 r;
	r.x = fmaf(e1.y, e2.z, -(e1.z * e2.y));
	r.y = fmaf(e1.z, e2.x, -(e1.x * e2.z));
	r.z = fmaf(e1.x, e2.y, -(e1.y * e2.x));
	return r;
}

/* ------------------------------------------------------------------------------------------ */
/* pinned acos                                                                                 */

/* acos(x) for the light-probe angular map (Kernels.h:217). fdlibm-style: a rational
 * approximation R(z) ~ (asin(sqrt z)/sqrt z - 1) on z in [0, 0.25], evaluated with a fixed
 * sequence of binary32 operations so that CPU and GPU agree bit for bit. |error| < 2 ulp
 * (checked against double acos in tests/test_oracle_kat.py). */
float oracle_acosf(float x) {
	const float pio2 = 1.57079637050628662109375f;  /* 0x3fc90fdb */
	const float pi   = 3.1415927410125732421875f;   /* 0x40490fdb */
	const float pS0 =  1.6666586697e-01f;
	const float pS1 = -4.2743422091e-02f;
	const float pS2 = -8.6563630030e-03f;
	const float qS1 = -7.0662963390e-01f;
	float ax = fabsf(x);
	if (!(ax < 1.0f)) {
		if (x != x) return x;
		return x > 0.0f ? 0.0f : pi;
	}
	if (ax <= 0.5f) {
		float z = x * x;
		float p = z * fmaf(z, fmaf(z, pS2, pS1), pS0);
		float q = fmaf(z, qS1, 1.0f);
		float r = p / q;
		return pio2 - fmaf(x, r, x);
	}
	{
		float z = (1.0f - ax) * 0.5f;
		float s = sqrtf(z);
		float p = z * fmaf(z, fmaf(z, pS2, pS1), pS0);
		float q = fmaf(z, qS1, 1.0f);
		float r = p / q;
		float w = 2.0f * fmaf(s, r, s);
		return x > 0.0f ? w : pi - w;
	}
}

/* ------------------------------------------------------------------------------------------ */
/* light probe, Kernels.h:137,213-221                                                          */

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* floor(x) to int for the texel index, saturated well outside the image so the int conversion
 * is defined for any finite or infinite x. NaN maps to 0. */
static inline int texel_floor(float x, float* frac) {
	float f = floorf(x);
	*frac = x - f;
	if (!(f > -4.0f)) { if (f != f) { *frac = 0.0f; return 0; } return -4; }
	if (f > 1.0e9f) return 1000000000;
	return (int)f;
}

void oracle_env_sample(const float* env, uint32_t width, uint32_t height, const float d[3], float rgb[3]) {
	unsigned saved = ftz_on();
	/* Kernels.h:216  rlen = native_rsqrt(d.y*d.y + d.z*d.z) */
	float s = d[1] * d[1] + d[2] * d[2];
	float rlen = 1.0f / sqrtf(s);
	/* Kernels.h:217  r = (rlen > 1e+6f) ? 0 : acos(-d.x)*(1/(2*3.141593f))*rlen */
	const float inv2pi = 1.0f / (2.0f * 3.141593f);
	float r = (rlen > 1e+6f) ? 0.0f : (oracle_acosf(-d[0]) * inv2pi) * rlen;
	/* Kernels.h:218-219 */
	float u = 0.5f - r * d[2];
	float v = 0.5f - r * d[1];
	/* read_imagef, normalized coords, CLK_ADDRESS_CLAMP_TO_EDGE, CLK_FILTER_LINEAR
	 * (OpenCL 1.2 spec 8.2): i0 = floor(u*w - 0.5), a = frac(u*w - 0.5), clamp indices. */
	float fu = u * (float)(int)width - 0.5f;
	float fv = v * (float)(int)height - 0.5f;
	float a, b;
	int i0 = texel_floor(fu, &a);
	int j0 = texel_floor(fv, &b);
	int i1 = clampi(i0 + 1, 0, (int)width - 1);
	int j1 = clampi(j0 + 1, 0, (int)height - 1);
	i0 = clampi(i0, 0, (int)width - 1);
	j0 = clampi(j0, 0, (int)height - 1);
	const float* t00 = env + 4 * ((size_t)j0 * width + (size_t)i0);
	const float* t10 = env + 4 * ((size_t)j0 * width + (size_t)i1);
	const float* t01 = env + 4 * ((size_t)j1 * width + (size_t)i0);
	const float* t11 = env + 4 * ((size_t)j1 * width + (size_t)i1);
	float na = 1.0f - a, nb = 1.0f - b;
	float w00 = na * nb, w10 = a * nb, w01 = na * b, w11 = a * b;
	for (int c = 0; c < 3; ++c)
		rgb[c] = ((w00 * t00[c] + w10 * t10[c]) + w01 * t01[c]) + w11 * t11[c];
	_mm_setcsr(saved);
}

/* ------------------------------------------------------------------------------------------ */
/* intersection primitives                                                                     */

typedef struct {
	v3 o, d;       /* origin; direction after the epsilon clamp */
	float tNear;   /* RAY_NEAR */
	float tFar;    /* RAY_FAR: shrinks as hits are found */
} ray_state;

typedef struct { uint32_t index; float t, u, v; } hit_state;

/* trianglePairIntersect, Kernels.h:36-115. Returns the new RAY_FAR. */
static inline float pair_intersect(const float* pairs, uint32_t index, const ray_state* ray, hit_state* hit) {
	const float* t0 = pairs + 12 * (size_t)index;
	const float* t1 = t0 + 4;
	const float* t2 = t0 + 8;

	float tNear = ray->tNear;
	float tMax = ray->tFar;
	v3 ro = ray->o, rd = ray->d;

	v3 e1 = { t0[0], t0[1], t0[2] };
	v3 e2 = { t1[0], t1[1], t1[2] };
	v3 e3 = { t0[3], t1[3], t2[3] };
	v3 v0 = { t2[0], t2[1], t2[2] };

	v3 n1 = mad_cross(e1, e2);
	v3 n2 = mad_cross(e3, e1);

	v3 C = { v0.x - ro.x, v0.y - ro.y, v0.z - ro.z };
	v3 R = mad_cross(rd, C);

	float det1 = dot3(n1, rd);
	float det2 = dot3(n2, rd);

	uint32_t sgnDet1 = f2u(det1) & 0x80000000u;
	uint32_t sgnDet2 = f2u(det2) & 0x80000000u;

	float dRe1 = dot3(R, e1);
	int32_t iU1 = (int32_t)(f2u(dot3(R, e2)) ^ sgnDet1);
	int32_t iV1 = (int32_t)(f2u(dRe1) ^ sgnDet1);
	/* Kernels.h:65-66 negate two dot products. The minus is a sign-bit flip here: written as -dot3(...), gcc folds it into the
	 * last fma of the dot product (vfnmsub), and when the products cancel exactly that instruction returns +0 where the
	 * source's -(+0) is -0 -- the sign of that zero decides which of two triangles owns their shared edge (a ray aimed at an
	 * edge of an integer-grid mesh; found by tests/fuzz/fuzz_gpu.py, where the B200 followed the source and this file did not). */
	uint32_t flip2 = sgnDet2 ^ 0x80000000u;
	int32_t iU2 = (int32_t)(f2u(dRe1) ^ flip2);
	int32_t iV2 = (int32_t)(f2u(dot3(R, e3)) ^ flip2);

	if (((iU1 | iV1) & (iU2 | iV2)) < 0)
		return tMax;

	int outside1 = (iU1 | iV1) < 0;
	int outside2 = (iU2 | iV2) < 0;

	float U1 = u2f((uint32_t)iU1), V1 = u2f((uint32_t)iV1);
	float U2 = u2f((uint32_t)iU2), V2 = u2f((uint32_t)iV2);

	float absDet1 = fabsf(det1);
	float absDet2 = fabsf(det2);

	float W1 = (absDet1 - U1) - V1;
	float W2 = (absDet2 - U2) - V2;

	float T1 = u2f(f2u(dot3(n1, C)) ^ sgnDet1);
	float T2 = u2f(f2u(dot3(n2, C)) ^ sgnDet2);

	outside1 = outside1 || (W1 < 0.0f || T1 <= absDet1 * tNear || T1 > absDet1 * tMax);
	outside2 = outside2 || (W2 < 0.0f || T2 <= absDet2 * tNear || T2 > absDet2 * tMax);

	if (outside1 && outside2)
		return tMax;

	index = index * 2;

	if ((!outside2 && outside1) || (!outside1 && !outside2 && T1 * absDet2 > T2 * absDet1)) {
		absDet1 = absDet2;
		T1 = T2;
		U1 = U2;
		V1 = V2;
		++index;
	}

	float rcpAbsDet1 = 1.0f / absDet1; /* native_recip, pinned to IEEE division */
	float t = T1 * rcpAbsDet1;
	float u = U1 * rcpAbsDet1;
	float v = V1 * rcpAbsDet1;

	hit->index = index;
	hit->t = t;
	hit->u = u;
	hit->v = v;
	return t;
}

/* aabbIntersect, Kernels.h:117-135 */
static inline float aabb_intersect(v3 mn, v3 mx, float t0, float t1, v3 invDir, v3 OoD) {
	float rayFar = t1;
	float nx = fmaf(mn.x, invDir.x, OoD.x), ny = fmaf(mn.y, invDir.y, OoD.y), nz = fmaf(mn.z, invDir.z, OoD.z);
	float fx = fmaf(mx.x, invDir.x, OoD.x), fy = fmaf(mx.y, invDir.y, OoD.y), fz = fmaf(mx.z, invDir.z, OoD.z);
	float minx = pmin(nx, fx), miny = pmin(ny, fy), minz = pmin(nz, fz);
	float maxx = pmax(nx, fx), maxy = pmax(ny, fy), maxz = pmax(nz, fz);
	t0 = pmax(pmax(t0, minx), pmax(miny, minz));
	t1 = pmin(pmin(t1, maxx), pmin(maxy, maxz));
	if (t0 > t1)
		return rayFar;
	return t0;
}

/* miss (Kernels.h:213-222) and hit (Kernels.h:223-239) epilogues */
static int finish_ray(const oracle_scene* sc, const ray_state* rayp, const hit_state* hitp, oracle_result* out) {
	const ray_state ray = *rayp;
	const hit_state hit = *hitp;
	if (hit.index == 0xffffffffu) { /* Kernels.h:213-222 */
		float d[3] = { ray.d.x, ray.d.y, ray.d.z };
		float rgb[3] = { 0.0f, 0.0f, 0.0f };
		if (sc->env)
			oracle_env_sample(sc->env, sc->env_width, sc->env_height, d, rgb);
		out->triangle = ORACLE_INVALID_TRIANGLE;
		out->a = rgb[0]; out->b = rgb[1]; out->c = rgb[2];
	}
	else { /* Kernels.h:223-239 */
		if (hit.index >= sc->remap_count) return -1;
		uint32_t index = sc->remap[hit.index];
		uint32_t edge = index >> 30;
		index &= 0x3fffffffu;
		float bx = hit.u, by = hit.v, bz = (1.0f - hit.u) - hit.v;
		float u = bx, v = by;
		if (edge == 1) { u = bz; v = bx; }      /* barys.zxy */
		else if (edge == 2) { u = by; v = bz; } /* barys.yzx */
		out->triangle = index;
		out->a = hit.t; out->b = u; out->c = v;
	}
	return 0;
}

#define ORACLE_STACK 64 /* Kernels.h:166 */

/* traversal, Kernels.h:141-242, one ray. Returns 0, or -1 on stack overflow / bad reference. */
static int traverse_one(const oracle_scene* sc, const oracle_ray* in, oracle_result* out, oracle_counters* cnt) {
	ray_state ray;
	ray.o.x = in->origin[0]; ray.o.y = in->origin[1]; ray.o.z = in->origin[2];
	ray.d.x = in->dir[0]; ray.d.y = in->dir[1]; ray.d.z = in->dir[2];
	ray.tNear = in->minT;
	ray.tFar = in->maxT;

	/* Kernels.h:149-157 */
	const float epsilon = 1e-10f;
	if (fabsf(ray.d.x) < epsilon) ray.d.x = copysignf(epsilon, ray.d.x);
	if (fabsf(ray.d.y) < epsilon) ray.d.y = copysignf(epsilon, ray.d.y);
	if (fabsf(ray.d.z) < epsilon) ray.d.z = copysignf(epsilon, ray.d.z);

	/* Kernels.h:159-160 */
	v3 invDir = { 1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z };
	v3 OoD = { -ray.o.x * invDir.x, -ray.o.y * invDir.y, -ray.o.z * invDir.z };

	hit_state hit = { 0xffffffffu, ray.tFar, 0.0f, 0.0f }; /* Kernels.h:162 */

	uint32_t node = 0x80000000u; /* Kernels.h:164 */
	uint32_t stack[ORACLE_STACK];
	unsigned stackHead = 0;
	unsigned nInner = 0, nPairs = 0, maxStack = 0, nPushes = 0, nLeaves = 0;
	const uint32_t* nodesU = (const uint32_t*)sc->nodes;

	for (;;) {
		if (node & 0x80000000u) {
			node &= ~0x80000000u;
			if (node >= sc->node_count) return -1;
			const float* d = sc->nodes + 16 * (size_t)node;
			uint32_t childFirst = nodesU[16 * (size_t)node + 2];
			uint32_t childLast = nodesU[16 * (size_t)node + 3];
			++nInner;

			v3 firstMin = { d[4], d[5], d[6] };
			v3 firstMax = { d[7], d[8], d[9] };
			v3 lastMin = { d[10], d[11], d[12] };
			v3 lastMax = { d[13], d[14], d[15] };

			float tRay = ray.tFar;
			float tFirst = aabb_intersect(firstMin, firstMax, ray.tNear, ray.tFar, invDir, OoD);
			float tLast = aabb_intersect(lastMin, lastMax, ray.tNear, ray.tFar, invDir, OoD);

			float firstDiff = tRay - tFirst;
			float lastDiff = tRay - tLast;
			if (firstDiff + lastDiff != 0.0f) { /* Kernels.h:192 */
				int sgn = (f2u(tLast - tFirst) >> 31) != 0; /* signbit, Kernels.h:193 */
				if (pmax(tFirst, tLast) != tRay) { /* both hit: push the far one, Kernels.h:194-195 */
					if (stackHead >= ORACLE_STACK) return -1;
					stack[stackHead++] = sgn ? childFirst : childLast;
					++nPushes;
					if (stackHead > maxStack) maxStack = stackHead;
				}
				node = sgn ? childLast : childFirst; /* Kernels.h:196 */
				continue;
			}
		}
		else {
			uint32_t first = node & 0xffffffu; /* Kernels.h:201-204 */
			uint32_t last = first + (node >> 24);
			if (last > sc->pair_count) return -1;
			++nLeaves;
			for (uint32_t i = first; i < last; ++i) {
				ray.tFar = pair_intersect(sc->pairs, i, &ray, &hit);
				++nPairs;
			}
		}
		if (!stackHead)
			break;
		node = stack[--stackHead];
	}

	if (finish_ray(sc, &ray, &hit, out))
		return -1;
	if (cnt) {
		cnt->inner = (uint16_t)(nInner > 65535 ? 65535 : nInner);
		cnt->pairs = (uint16_t)(nPairs > 65535 ? 65535 : nPairs);
		cnt->max_stack = (uint16_t)maxStack;
		cnt->hit = hit.index != 0xffffffffu;
		cnt->pushes = (uint16_t)(nPushes > 65535 ? 65535 : nPushes);
		cnt->leaves = (uint16_t)(nLeaves > 65535 ? 65535 : nLeaves);
	}
	return 0;
}

/* minimal fork-join helper (pthreads; no OpenMP dependency) */
typedef struct {
	void (*body)(void* ctx, int64_t chunk);
	void* ctx;
	int64_t nchunks;
	volatile int64_t next;
} pfor_job;

static void* pfor_worker(void* p) {
	pfor_job* job = (pfor_job*)p;
	unsigned saved = ftz_on();
	for (;;) {
		int64_t c = __sync_fetch_and_add(&job->next, 1);
		if (c >= job->nchunks) break;
		job->body(job->ctx, c);
	}
	_mm_setcsr(saved);
	return 0;
}

static void parallel_for(int threads, int64_t nchunks, void (*body)(void*, int64_t), void* ctx) {
	if (threads <= 0) threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
	if (threads < 1) threads = 1;
	if (threads > 256) threads = 256;
	if ((int64_t)threads > nchunks) threads = nchunks > 0 ? (int)nchunks : 1;
	pfor_job job = { body, ctx, nchunks, 0 };
	pthread_t tid[256];
	for (int i = 1; i < threads; ++i) pthread_create(&tid[i], 0, pfor_worker, &job);
	pfor_worker(&job);
	for (int i = 1; i < threads; ++i) pthread_join(tid[i], 0);
}

int oracle_hardware_threads(void) { return (int)sysconf(_SC_NPROCESSORS_ONLN); }

typedef struct {
	const oracle_scene* scene; const oracle_ray* rays; uint32_t count;
	oracle_result* results; oracle_counters* counters; volatile int err;
} trav_ctx;

/* chunks of 1024 rays, the reference's cpuTestBatch (RayAccelerator.cpp:438) */
#define TRAV_CHUNK 1024

static void trav_body(void* p, int64_t c) {
	trav_ctx* t = (trav_ctx*)p;
	int64_t lo = c * TRAV_CHUNK, hi = lo + TRAV_CHUNK < (int64_t)t->count ? lo + TRAV_CHUNK : (int64_t)t->count;
	for (int64_t i = lo; i < hi; ++i)
		if (traverse_one(t->scene, t->rays + i, t->results + i, t->counters ? t->counters + i : 0))
			t->err = -1;
}

int oracle_traverse(const oracle_scene* scene, const oracle_ray* rays, uint32_t count,
                    oracle_result* results, oracle_counters* counters, int threads) {
	trav_ctx t = { scene, rays, count, results, counters, 0 };
	parallel_for(threads, ((int64_t)count + TRAV_CHUNK - 1) / TRAV_CHUNK, trav_body, &t);
	return t.err;
}


/* ------------------------------------------------------------------------------------------ */
/* CPU BASELINE (bench.py cpu_baseline / --impl reference only; SURVEY.md 8d "AVX2 restatement of the
 * reference's own algorithm, 1 ray x 2 boxes"): the same traversal with both child boxes of a node
 * tested in one AVX2 pass. Same fused multiply-adds, exact min/max; the only liberty is which zero
 * (+0/-0) an intermediate min/max returns, which cannot reach the result while minT >= 0 (the final
 * max takes minT as the tie winner). tests/test_oracle_kat.py checks it bit for bit against
 * oracle_traverse on the golden rays. Not used as the parity checker. */
typedef struct { __m128 inv1, inv2, inv3, ood1, ood2, ood3; } ray_simd;

static inline void node_test_avx2(const float* d, const ray_simd* rs, float tNear, float tFar, float* tFirst, float* tLast) {
	/* node floats 4..15 = lmn.xyz lmx.x | lmx.yz rmn.xy | rmn.z rmx.xyz; ray constants pre-permuted to match */
	const __m128 t1 = _mm_fmadd_ps(_mm_loadu_ps(d + 4), rs->inv1, rs->ood1);
	const __m128 t2 = _mm_fmadd_ps(_mm_loadu_ps(d + 8), rs->inv2, rs->ood2);
	const __m128 t3 = _mm_fmadd_ps(_mm_loadu_ps(d + 12), rs->inv3, rs->ood3);
	const __m128 al = _mm_shuffle_ps(t1, t1, _MM_SHUFFLE(2, 2, 1, 0));                 /* lmn.x lmn.y lmn.z lmn.z */
	const __m128 tmp = _mm_shuffle_ps(t1, t2, _MM_SHUFFLE(1, 0, 3, 3));                /* lmx.x lmx.x lmx.y lmx.z */
	const __m128 bl = _mm_shuffle_ps(tmp, tmp, _MM_SHUFFLE(3, 3, 2, 0));               /* lmx.x lmx.y lmx.z lmx.z */
	const __m128 ar = _mm_shuffle_ps(t2, t3, _MM_SHUFFLE(0, 0, 3, 2));                 /* rmn.x rmn.y rmn.z rmn.z */
	const __m128 br = _mm_shuffle_ps(t3, t3, _MM_SHUFFLE(3, 3, 2, 1));                 /* rmx.x rmx.y rmx.z rmx.z */
	const __m256 a = _mm256_set_m128(ar, al), b = _mm256_set_m128(br, bl);
	__m256 lo = _mm256_min_ps(a, b), hi = _mm256_max_ps(a, b);
	lo = _mm256_max_ps(lo, _mm256_permute_ps(lo, _MM_SHUFFLE(1, 0, 3, 2)));
	hi = _mm256_min_ps(hi, _mm256_permute_ps(hi, _MM_SHUFFLE(1, 0, 3, 2)));
	lo = _mm256_max_ps(lo, _mm256_permute_ps(lo, _MM_SHUFFLE(2, 3, 0, 1)));
	hi = _mm256_min_ps(hi, _mm256_permute_ps(hi, _MM_SHUFFLE(2, 3, 0, 1)));
	const float e0l = _mm_cvtss_f32(_mm256_castps256_ps128(lo)), e0r = _mm_cvtss_f32(_mm256_extractf128_ps(lo, 1));
	const float x1l = _mm_cvtss_f32(_mm256_castps256_ps128(hi)), x1r = _mm_cvtss_f32(_mm256_extractf128_ps(hi, 1));
	const float t0l = e0l > tNear ? e0l : tNear, t0r = e0r > tNear ? e0r : tNear;
	const float t1l = x1l < tFar ? x1l : tFar, t1r = x1r < tFar ? x1r : tFar;
	*tFirst = t0l > t1l ? tFar : t0l;
	*tLast = t0r > t1r ? tFar : t0r;
}

static int traverse_one_avx2(const oracle_scene* sc, const oracle_ray* in, oracle_result* out) {
	ray_state ray;
	ray.o.x = in->origin[0]; ray.o.y = in->origin[1]; ray.o.z = in->origin[2];
	ray.d.x = in->dir[0]; ray.d.y = in->dir[1]; ray.d.z = in->dir[2];
	ray.tNear = in->minT;
	ray.tFar = in->maxT;
	const float epsilon = 1e-10f;
	if (fabsf(ray.d.x) < epsilon) ray.d.x = copysignf(epsilon, ray.d.x);
	if (fabsf(ray.d.y) < epsilon) ray.d.y = copysignf(epsilon, ray.d.y);
	if (fabsf(ray.d.z) < epsilon) ray.d.z = copysignf(epsilon, ray.d.z);
	const float ix = 1.0f / ray.d.x, iy = 1.0f / ray.d.y, iz = 1.0f / ray.d.z;
	const float px = -ray.o.x * ix, py = -ray.o.y * iy, pz = -ray.o.z * iz;
	ray_simd rs;
	rs.inv1 = _mm_setr_ps(ix, iy, iz, ix); rs.ood1 = _mm_setr_ps(px, py, pz, px);
	rs.inv2 = _mm_setr_ps(iy, iz, ix, iy); rs.ood2 = _mm_setr_ps(py, pz, px, py);
	rs.inv3 = _mm_setr_ps(iz, ix, iy, iz); rs.ood3 = _mm_setr_ps(pz, px, py, pz);
	hit_state hit = { 0xffffffffu, ray.tFar, 0.0f, 0.0f };
	uint32_t node = 0x80000000u;
	uint32_t stack[ORACLE_STACK];
	unsigned stackHead = 0;
	const uint32_t* nodesU = (const uint32_t*)sc->nodes;
	for (;;) {
		if (node & 0x80000000u) {
			node &= ~0x80000000u;
			if (node >= sc->node_count) return -1;
			const float* d = sc->nodes + 16 * (size_t)node;
			const uint32_t childFirst = nodesU[16 * (size_t)node + 2], childLast = nodesU[16 * (size_t)node + 3];
			const float tRay = ray.tFar;
			float tFirst, tLast;
			node_test_avx2(d, &rs, ray.tNear, tRay, &tFirst, &tLast);
			if ((tRay - tFirst) + (tRay - tLast) != 0.0f) {
				const int sgn = (f2u(tLast - tFirst) >> 31) != 0;
				if ((tFirst > tLast ? tFirst : tLast) != tRay) {
					if (stackHead >= ORACLE_STACK) return -1;
					stack[stackHead++] = sgn ? childFirst : childLast;
				}
				node = sgn ? childLast : childFirst;
				continue;
			}
		}
		else {
			const uint32_t first = node & 0xffffffu, last = first + (node >> 24);
			if (last > sc->pair_count) return -1;
			for (uint32_t i = first; i < last; ++i)
				ray.tFar = pair_intersect(sc->pairs, i, &ray, &hit);
		}
		if (!stackHead)
			break;
		node = stack[--stackHead];
	}
	return finish_ray(sc, &ray, &hit, out);
}

static void trav_body_avx2(void* p, int64_t c) {
	trav_ctx* t = (trav_ctx*)p;
	unsigned saved = ftz_on(); /* the reference's threads run FTZ+DAZ; min/max then treat subnormals as zero */
	int64_t lo = c * TRAV_CHUNK, hi = lo + TRAV_CHUNK < (int64_t)t->count ? lo + TRAV_CHUNK : (int64_t)t->count;
	for (int64_t i = lo; i < hi; ++i)
		if (traverse_one_avx2(t->scene, t->rays + i, t->results + i))
			t->err = -1;
	_mm_setcsr(saved);
}

int oracle_traverse_avx2(const oracle_scene* scene, const oracle_ray* rays, uint32_t count, oracle_result* results, int threads) {
	trav_ctx t = { scene, rays, count, results, 0, 0 };
	parallel_for(threads, ((int64_t)count + TRAV_CHUNK - 1) / TRAV_CHUNK, trav_body_avx2, &t);
	return t.err;
}

/* ------------------------------------------------------------------------------------------ */
/* independent fp64 arbiter (textbook Moller-Trumbore, no relation to the pair formulation)    */

static inline double tri_t_f64(const float* verts4, const uint32_t* tri, const oracle_ray* r) {
	const float* p0 = verts4 + 4 * (size_t)tri[0];
	const float* p1 = verts4 + 4 * (size_t)tri[1];
	const float* p2 = verts4 + 4 * (size_t)tri[2];
	double e1x = (double)p1[0] - p0[0], e1y = (double)p1[1] - p0[1], e1z = (double)p1[2] - p0[2];
	double e2x = (double)p2[0] - p0[0], e2y = (double)p2[1] - p0[1], e2z = (double)p2[2] - p0[2];
	double dx = r->dir[0], dy = r->dir[1], dz = r->dir[2];
	double px = dy * e2z - dz * e2y, py = dz * e2x - dx * e2z, pz = dx * e2y - dy * e2x;
	double det = e1x * px + e1y * py + e1z * pz;
	if (det == 0.0) return INFINITY;
	double inv = 1.0 / det;
	double tx = (double)r->origin[0] - p0[0], ty = (double)r->origin[1] - p0[1], tz = (double)r->origin[2] - p0[2];
	double u = (tx * px + ty * py + tz * pz) * inv;
	if (u < 0.0 || u > 1.0) return INFINITY;
	double qx = ty * e1z - tz * e1y, qy = tz * e1x - tx * e1z, qz = tx * e1y - ty * e1x;
	double v = (dx * qx + dy * qy + dz * qz) * inv;
	if (v < 0.0 || u + v > 1.0) return INFINITY;
	double t = (e2x * qx + e2y * qy + e2z * qz) * inv;
	if (!(t > (double)r->minT) || !(t <= (double)r->maxT)) return INFINITY; /* Kernels.h:88 interval */
	return t;
}

typedef struct {
	const float* verts4; const uint32_t* indices; uint32_t ntris;
	const oracle_ray* rays; uint32_t count; double* t_min; uint32_t* tri_min;
} brute_ctx;

#define BRUTE_CHUNK 16

static void brute_body(void* p, int64_t c) {
	brute_ctx* b = (brute_ctx*)p;
	int64_t lo = c * BRUTE_CHUNK, hi = lo + BRUTE_CHUNK < (int64_t)b->count ? lo + BRUTE_CHUNK : (int64_t)b->count;
	for (int64_t i = lo; i < hi; ++i) {
		double best = INFINITY;
		uint32_t id = ORACLE_INVALID_TRIANGLE;
		for (uint32_t k = 0; k < b->ntris; ++k) {
			double t = tri_t_f64(b->verts4, b->indices + 3 * (size_t)k, b->rays + i);
			if (t < best) { best = t; id = k; }
		}
		b->t_min[i] = best;
		b->tri_min[i] = id;
	}
}

int oracle_brute_f64(const float* verts4, uint32_t nverts, const uint32_t* indices, uint32_t ntris,
                     const oracle_ray* rays, uint32_t count, double* t_min, uint32_t* tri_min, int threads) {
	(void)nverts;
	brute_ctx b = { verts4, indices, ntris, rays, count, t_min, tri_min };
	parallel_for(threads, ((int64_t)count + BRUTE_CHUNK - 1) / BRUTE_CHUNK, brute_body, &b);
	return 0;
}

int oracle_tri_t_f64(const float* verts4, const uint32_t* indices, const oracle_ray* rays,
                     const uint32_t* tri_ids, uint32_t count, double* t_out) {
	for (uint32_t i = 0; i < count; ++i)
		t_out[i] = tri_ids[i] == ORACLE_INVALID_TRIANGLE ? INFINITY
		         : tri_t_f64(verts4, indices + 3 * (size_t)tri_ids[i], rays + i);
	return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* numbering-independent structural digest                                                     */

static inline uint64_t fnv(uint64_t h, const void* p, size_t n) {
	const unsigned char* b = (const unsigned char*)p;
	for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
	return h;
}

int oracle_scene_digest(const oracle_scene* sc, uint64_t digest[16]) {
	memset(digest, 0, 16 * sizeof(uint64_t));
	if (!sc->node_count) return -1;
	typedef struct { uint32_t ref; uint32_t depth; } item;
	size_t cap = (size_t)sc->node_count * 2 + 16;
	item* st = (item*)malloc(cap * sizeof(item));
	size_t top = 0;
	uint64_t h = 14695981039346656037ull;
	const uint32_t* nodesU = (const uint32_t*)sc->nodes;
	st[top].ref = 0x80000000u; st[top].depth = 1; ++top;
	int rc = 0;
	while (top) {
		item it = st[--top];
		if (it.depth > digest[3]) digest[3] = it.depth;
		if (it.ref & 0x80000000u) {
			uint32_t n = it.ref & 0x7fffffffu;
			if (n >= sc->node_count || top + 2 > cap) { rc = -1; break; }
			digest[0]++;
			h = fnv(h, sc->nodes + 16 * (size_t)n + 4, 48); /* the two child boxes */
			/* visit first child before last child */
			st[top].ref = nodesU[16 * (size_t)n + 3]; st[top].depth = it.depth + 1; ++top;
			st[top].ref = nodesU[16 * (size_t)n + 2]; st[top].depth = it.depth + 1; ++top;
		}
		else {
			uint32_t first = it.ref & 0xffffffu, cnt = it.ref >> 24;
			if (first + cnt > sc->pair_count) { rc = -1; break; }
			digest[1]++;
			digest[2] += cnt;
			if (cnt >= 1 && cnt <= 7) digest[4 + cnt]++;
			h = fnv(h, &cnt, 4);
			for (uint32_t i = first; i < first + cnt; ++i) {
				const float* p = sc->pairs + 12 * (size_t)i;
				if (p[3] == -p[0] && p[7] == -p[1] && p[11] == -p[2]) digest[4]++; /* p3 == p1 */
				h = fnv(h, p, 48);
				h = fnv(h, sc->remap + 2 * (size_t)i, 8);
			}
		}
	}
	digest[12] = h;
	free(st);
	return rc;
}


/* ------------------------------------------------------------------------------------------ */
/* Wavefront path tracer: CPU restatement of what the reference's example
 * client computes per path -- Renderer/PathTracingRenderer.cpp:72-566 (shade),
 * Renderer/Materials.cpp:11-151 (ReflectiveDiffuseMaterial::sample8 and its sin/cos parabolas),
 * Renderer/Camera.cpp:55-114 (generateTileRays), Renderer/LightPath.cpp:11-39 -- one path at a
 * time, in the arithmetic the CUDA renderer (rayaccel_b200/csrc/pathtrace.cu) uses, so that the
 * two agree bit for bit. Departures from the reference, all of them where the reference is not
 * reproducible itself: its random numbers come from an MWC generator seeded with libc rand() per
 * call (SimdRandom.h:20-56, PathTracingRenderer.cpp:107, Camera.cpp:58), here from a counter-based
 * hash of (pixel, sample, depth, seed); its _mm256_rsqrt_ps / _mm256_rcp_ps 12-bit approximations
 * are exact 1/sqrt and 1/x here. The estimator (what is sampled with which probability and
 * weight, when a path ends) is the reference's. */

static inline uint32_t pcg_hash(uint32_t v) {
	uint32_t s = v * 747796405u + 2891336453u;
	uint32_t w = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
	return (w >> 22u) ^ w;
}
static inline float unit_float(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }
static inline float xor_sign(float x, uint32_t signbit) { return u2f(f2u(x) ^ signbit); }

/* Materials.cpp:11-22: parabola through sin(2 pi x), x in [0,1] */
static inline float sin_approx(float x) {
	float y = fmaf(-16.0f, x, 8.0f);
	int gt = x >= 0.5f;
	float xy = x * y;
	if (gt) xy = -xy;
	return xy + (gt ? y : 0.0f);
}
/* Materials.cpp:24-28 */
static inline float cos_approx(float x) {
	float y = x - 0.75f;
	x = (f2u(y) & 0x80000000u) ? x + 0.25f : y;
	return sin_approx(x);
}

/* Materials.cpp:39-151, one lane */
void oracle_material_sample(const float ke[4], const float rnd[3], const float normal[3], const float wo[3], float wi[3], float color[3]) {
	const float nx = normal[0], ny = normal[1], nz = normal[2];
	const float eta = ke[3];
	/* reflection vector and fresnel term, :57-80 */
	float cosi = fmaf(nz, wo[2], fmaf(ny, wo[1], nx * wo[0]));
	cosi = cosi > 0.0f ? cosi : 0.0f;
	const float c2 = 2.0f * cosi;
	const float rx = fmaf(c2, nx, -wo[0]), ry = fmaf(c2, ny, -wo[1]), rz = fmaf(c2, nz, -wo[2]);
	const float cosi2m1 = fmaf(cosi, cosi, -1.0f);
	const float eta2 = eta * eta;
	const float k = fmaf(eta2, cosi2m1, 1.0f);
	const float cost = sqrtf(k);
	const float rper = fmaf(eta, cosi, -cost) * (1.0f / fmaf(eta, cosi, cost));
	const float rpar = -(fmaf(eta, cost, -cosi) * (1.0f / fmaf(eta, cost, cosi)));
	float fresnel = 0.5f * fmaf(rpar, rpar, rper * rper);
	if (f2u(k) & 0x80000000u) fresnel = 1.0f;
	/* diffuse direction, :82-120 */
	const int wide = !(fabsf(nx) <= 0.1f);
	float ux = wide ? -nz : 0.0f, uy = wide ? 0.0f : -nz, uz = wide ? nx : ny;
	const float fb = 1.0f / sqrtf(fmaf(uz, uz, fmaf(uy, uy, ux * ux)));
	ux *= fb; uy *= fb; uz *= fb;
	const float vx = fmaf(ny, uz, -(nz * uy)), vy = fmaf(nz, ux, -(nx * uz)), vz = fmaf(nx, uy, -(ny * ux));
	const float sinx = sin_approx(rnd[0]), cosx = cos_approx(rnd[0]);
	const float r2s = sqrtf(rnd[1]);
	const float sq = sqrtf(1.0f - rnd[1]);
	float dx = fmaf(nx, sq, fmaf(ux, cosx, vx * sinx) * r2s);
	float dy = fmaf(ny, sq, fmaf(uy, cosx, vy * sinx) * r2s);
	float dz = fmaf(nz, sq, fmaf(uz, cosx, vz * sinx) * r2s);
	const float fd = 1.0f / sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
	dx *= fd; dy *= fd; dz *= fd;
	/* reflection or diffuse, :122-150 */
	const float s0 = fresnel * 3.0f;
	const float s1 = ke[2] + (ke[0] + ke[1]);
	const float sum = s0 + s1;
	const float uniform = rnd[2] * sum;
	const int diffuse = uniform >= s0;
	wi[0] = diffuse ? dx : rx; wi[1] = diffuse ? dy : ry; wi[2] = diffuse ? dz : rz;
	const float r = diffuse ? ke[0] : fresnel, g = diffuse ? ke[1] : fresnel, b = diffuse ? ke[2] : fresnel;
	const float scale = sum * (1.0f / (b + (r + g)));
	color[0] = r * scale; color[1] = g * scale; color[2] = b * scale;
}

/* One bounce of one path, PathTracingRenderer.cpp:124-127 (who is shaded), :231-246 (shading normal),
 * :376-466 (sample, weight, continuation test, next ray). Returns 1 when the path goes on. */
static int shade_hit(const oracle_shading* sh, const oracle_ray* ray, const oracle_result* res, const float rnd[3], float weight[3],
                     oracle_ray* next) {
	const uint32_t tri = res->triangle;
	const float t = res->a, u = res->b, v = res->c;
	const uint32_t* idx = sh->indices + 3 * (size_t)tri;
	const float* n0 = sh->normals4 + 4 * (size_t)idx[0];
	const float* n1 = sh->normals4 + 4 * (size_t)idx[1];
	const float* n2 = sh->normals4 + 4 * (size_t)idx[2];
	const float w = 1.0f - (u + v);
	float n[3];
	for (int k = 0; k < 3; ++k) n[k] = fmaf(n2[k], v, fmaf(n1[k], u, n0[k] * w));
	const float fn = 1.0f / sqrtf(fmaf(n[2], n[2], fmaf(n[1], n[1], n[0] * n[0])));
	const float* gn = sh->triangle_normals4 + 4 * (size_t)tri;
	const float rdgn = fmaf(ray->dir[2], gn[2], fmaf(ray->dir[1], gn[1], ray->dir[0] * gn[0]));
	const uint32_t sgn0 = f2u(rdgn) & 0x80000000u;
	float wo[3], pos[3];
	for (int k = 0; k < 3; ++k) {
		n[k] = xor_sign(n[k] * fn, sgn0);
		wo[k] = -ray->dir[k];
		pos[k] = fmaf(ray->dir[k], t, ray->origin[k]);
	}
	uint32_t m = sh->triangle_materials[tri];
	if (m >= sh->material_count) m = 0;
	float wi[3], color[3];
	oracle_material_sample(sh->materials_ke4 + 4 * (size_t)m, rnd, n, wo, wi, color);
	for (int k = 0; k < 3; ++k) weight[k] *= color[k];
	int go = weight[0] > 0.01f || weight[1] > 0.01f || weight[2] > 0.01f;
	const float sgn1 = fmaf(wi[2], gn[2], fmaf(wi[1], gn[1], wi[0] * gn[0]));
	go &= ((f2u(sgn1) ^ sgn0) >> 31) != 0; /* leaves on the side it arrived from (no transmission in this material) */
	const uint32_t flip = f2u(sgn1) & 0x80000000u;
	for (int k = 0; k < 3; ++k) {
		pos[k] = fmaf(xor_sign(gn[k], flip), 1e-4f, pos[k]);
		go &= pos[k] == pos[k] && wi[k] == wi[k];
		next->origin[k] = pos[k];
		next->dir[k] = wi[k];
	}
	next->minT = 1e-3f;
	next->maxT = 1e+6f;
	return go;
}

typedef struct {
	const oracle_scene* scene; const oracle_shading* sh; const oracle_camera* cam;
	uint32_t width, height, sample_base, spp, max_depth, seed;
	float* fb; uint64_t* waves; volatile int err;
} pt_ctx;

#define PT_CHUNK 256 /* pixels per task */

static void pt_body(void* p, int64_t c) {
	pt_ctx* x = (pt_ctx*)p;
	const uint64_t pixels = (uint64_t)x->width * x->height;
	uint64_t lo = (uint64_t)c * PT_CHUNK, hi = lo + PT_CHUNK < pixels ? lo + PT_CHUNK : pixels;
	uint64_t waves[64] = { 0 };
	for (uint64_t pixel = lo; pixel < hi; ++pixel) {
		const uint32_t px_i = (uint32_t)(pixel % x->width), py_i = (uint32_t)(pixel / x->width);
		float acc[3] = { x->fb[4 * pixel], x->fb[4 * pixel + 1], x->fb[4 * pixel + 2] };
		for (uint32_t s = 0; s < x->spp; ++s) {
			const uint32_t sample = x->sample_base + s;
			/* primary ray: Camera.cpp:55-114 */
			float jx = 0.5f, jy = 0.5f;
			if (x->seed) {
				uint32_t h = pcg_hash((uint32_t)pixel ^ pcg_hash(sample ^ pcg_hash(x->seed)));
				jx = unit_float(h);
				jy = unit_float(pcg_hash(h));
			}
			const float fx = (float)px_i + jx, fy = (float)py_i + jy;
			oracle_ray ray;
			float d[3];
			for (int k = 0; k < 3; ++k) d[k] = fmaf(x->cam->right[k], fx, fmaf(x->cam->up[k], fy, x->cam->view[k]));
			const float scale = 1.0f / sqrtf(fmaf(d[2], d[2], fmaf(d[1], d[1], d[0] * d[0])));
			for (int k = 0; k < 3; ++k) { ray.origin[k] = x->cam->origin[k]; ray.dir[k] = d[k] * scale; }
			ray.minT = 0.0f;
			ray.maxT = 1e+6f;
			float weight[3] = { 1.0f, 1.0f, 1.0f };
			float contrib[3] = { 0.0f, 0.0f, 0.0f };
			for (uint32_t depth = 0; depth <= x->max_depth; ++depth) {
				oracle_result res;
				if (traverse_one(x->scene, &ray, &res, 0)) { x->err = -1; break; }
				waves[depth < 64 ? depth : 63]++;
				if (res.triangle == ORACLE_INVALID_TRIANGLE) { /* PathTracingRenderer.cpp:468-566 */
					contrib[0] = res.a * weight[0]; contrib[1] = res.b * weight[1]; contrib[2] = res.c * weight[2];
					break;
				}
				if (!(res.triangle < x->sh->triangle_count && depth < x->max_depth)) break;
				uint32_t h = pcg_hash((uint32_t)pixel ^ pcg_hash(sample ^ pcg_hash(x->seed ^ (0x9e3779b9u * (depth + 1u)))));
				float rnd[3];
				rnd[0] = unit_float(h); h = pcg_hash(h);
				rnd[1] = unit_float(h); h = pcg_hash(h);
				rnd[2] = unit_float(h);
				oracle_ray next;
				if (!shade_hit(x->sh, &ray, &res, rnd, weight, &next)) break;
				ray = next;
			}
			for (int k = 0; k < 3; ++k) acc[k] += contrib[k];
		}
		for (int k = 0; k < 3; ++k) x->fb[4 * pixel + k] = acc[k];
	}
	if (x->waves)
		for (uint32_t dpt = 0; dpt <= x->max_depth && dpt < 64; ++dpt) __sync_fetch_and_add(&x->waves[dpt], waves[dpt]);
}

int oracle_path_trace(const oracle_scene* scene, const oracle_shading* shading, const oracle_camera* camera, uint32_t width,
                      uint32_t height, uint32_t sample_base, uint32_t spp, uint32_t max_depth, uint32_t seed, float* framebuffer4,
                      uint64_t* wave_rays, int threads) {
	if (max_depth > 62) return -1;
	pt_ctx x = { scene, shading, camera, width, height, sample_base, spp, max_depth, seed, framebuffer4, wave_rays, 0 };
	parallel_for(threads, ((int64_t)width * height + PT_CHUNK - 1) / PT_CHUNK, pt_body, &x);
	return x.err;
}


/* ------------------------------------------------------------------------------------------ */
/* Whitted renderer (the reference's second example client, Renderer/WhittedRenderer.cpp:136-676):
 * at every hit with depth < max_depth the path adds weight*0.3*max(n.L, 0) for the fixed light
 * direction L = (0.57, 0.57, 0.57) (:343-372), scales its weight by 0.3 and, while a weight channel
 * exceeds 0.01 (:404-413), continues as a mirror reflection AND a refraction (eta 1/1.1 entering,
 * 1.1 leaving, :415-437), each subject to a side test against the geometric normal (:440-444) and a
 * NaN test (:465-472); both children carry the parent's weight. Misses add probe radiance * weight
 * (:578-660). The reference keeps the tree depth-first in a linked list to bound memory (:19-134);
 * that is scheduling and is not restated -- the sum over the tree is. Contributions are summed in
 * 32.32 fixed point so that the result does not depend on the order the tree is walked in (the CUDA
 * renderer walks it breadth-first with atomics). */

#define FIXED_ONE 4294967296.0f
static inline uint64_t to_fixed(float c) {
	if (!(c > 0.0f)) return 0; /* also NaN */
	if (c > 1048576.0f) c = 1048576.0f;
	return (uint64_t)llrintf(c * FIXED_ONE);
}

typedef struct { oracle_ray ray; float weight[3]; uint32_t depth; } whitted_item;

typedef struct {
	const oracle_scene* scene; const oracle_shading* sh; const oracle_camera* cam;
	uint32_t width, height, sample_base, spp, max_depth, seed;
	float* fb; uint64_t* waves; volatile int err;
} wt_ctx;

static void wt_body(void* p, int64_t c) {
	wt_ctx* x = (wt_ctx*)p;
	const oracle_shading* sh = x->sh;
	const uint64_t pixels = (uint64_t)x->width * x->height;
	uint64_t lo = (uint64_t)c * PT_CHUNK, hi = lo + PT_CHUNK < pixels ? lo + PT_CHUNK : pixels;
	uint64_t waves[64] = { 0 };
	whitted_item stack[72]; /* depth-first: at most one pending sibling per level */
	for (uint64_t pixel = lo; pixel < hi; ++pixel) {
		uint64_t acc[3] = { 0, 0, 0 };
		for (uint32_t s = 0; s < x->spp; ++s) {
			const uint32_t sample = x->sample_base + s;
			float jx = 0.5f, jy = 0.5f;
			if (x->seed) {
				uint32_t h = pcg_hash((uint32_t)pixel ^ pcg_hash(sample ^ pcg_hash(x->seed)));
				jx = unit_float(h);
				jy = unit_float(pcg_hash(h));
			}
			const float fx = (float)(uint32_t)(pixel % x->width) + jx, fy = (float)(uint32_t)(pixel / x->width) + jy;
			float d[3];
			for (int k = 0; k < 3; ++k) d[k] = fmaf(x->cam->right[k], fx, fmaf(x->cam->up[k], fy, x->cam->view[k]));
			const float scale = 1.0f / sqrtf(fmaf(d[2], d[2], fmaf(d[1], d[1], d[0] * d[0])));
			int top = 0;
			for (int k = 0; k < 3; ++k) { stack[0].ray.origin[k] = x->cam->origin[k]; stack[0].ray.dir[k] = d[k] * scale; stack[0].weight[k] = 1.0f; }
			stack[0].ray.minT = 0.0f; stack[0].ray.maxT = 1e+6f; stack[0].depth = 0;
			top = 1;
			while (top) {
				whitted_item it = stack[--top];
				oracle_result res;
				if (traverse_one(x->scene, &it.ray, &res, 0)) { x->err = -1; break; }
				waves[it.depth < 64 ? it.depth : 63]++;
				if (res.triangle == ORACLE_INVALID_TRIANGLE) {
					acc[0] += to_fixed(res.a * it.weight[0]); acc[1] += to_fixed(res.b * it.weight[1]); acc[2] += to_fixed(res.c * it.weight[2]);
					continue;
				}
				if (!(res.triangle < sh->triangle_count && it.depth < x->max_depth)) continue;
				const uint32_t tri = res.triangle;
				const float t = res.a, u = res.b, v = res.c;
				const uint32_t* idx = sh->indices + 3 * (size_t)tri;
				const float* n0 = sh->normals4 + 4 * (size_t)idx[0];
				const float* n1 = sh->normals4 + 4 * (size_t)idx[1];
				const float* n2 = sh->normals4 + 4 * (size_t)idx[2];
				const float w = 1.0f - (u + v);
				float n[3];
				for (int k = 0; k < 3; ++k) n[k] = fmaf(n2[k], v, fmaf(n1[k], u, n0[k] * w));
				const float fn = 1.0f / sqrtf(fmaf(n[2], n[2], fmaf(n[1], n[1], n[0] * n[0])));
				const float* gn = sh->triangle_normals4 + 4 * (size_t)tri;
				const float* rd = it.ray.dir;
				const float rdgn = fmaf(rd[2], gn[2], fmaf(rd[1], gn[1], rd[0] * gn[0]));
				const uint32_t sgn0 = f2u(rdgn) & 0x80000000u;
				for (int k = 0; k < 3; ++k) n[k] = xor_sign(n[k] * fn, sgn0);
				/* direct light, :349-372 */
				float light = fmaf(n[2], 0.57f, fmaf(n[1], 0.57f, n[0] * 0.57f));
				light = light > 0.0f ? light : 0.0f;
				float weight[3];
				for (int k = 0; k < 3; ++k) {
					weight[k] = it.weight[k] * 0.3f;
					acc[k] += to_fixed(weight[k] * light);
				}
				if (!(!(weight[0] <= 0.01f) || !(weight[1] <= 0.01f) || !(weight[2] <= 0.01f))) continue; /* _CMP_NLE_US */
				/* reflection and refraction, :413-437 */
				const float ddn = fmaf(rd[2], n[2], fmaf(rd[1], n[1], rd[0] * n[0]));
				const float cosi = ddn * -2.0f;
				const float eta = sgn0 ? 1.1f : 1.0f / 1.1f;
				const float r = 1.0f - (eta * eta) * (1.0f - ddn * ddn);
				const float mu = fmaf(eta, ddn, sqrtf(r));
				float rl[3], rr[3], pos[3];
				for (int k = 0; k < 3; ++k) {
					rl[k] = fmaf(cosi, n[k], rd[k]);
					rr[k] = fmaf(eta, rd[k], -(mu * n[k]));
					pos[k] = fmaf(rd[k], t, it.ray.origin[k]);
				}
				const float sl = fmaf(rl[2], gn[2], fmaf(rl[1], gn[1], rl[0] * gn[0]));
				const float sr = fmaf(rr[2], gn[2], fmaf(rr[1], gn[1], rr[0] * gn[0]));
				int reflect = ((f2u(sl) ^ sgn0) >> 31) != 0;       /* back to the side it came from */
				int refract = r > 0.0f && ((f2u(sr) ^ sgn0) >> 31) == 0; /* through the surface */
				whitted_item a, b;
				for (int k = 0; k < 3; ++k) {
					a.ray.origin[k] = fmaf(xor_sign(gn[k], f2u(sl) & 0x80000000u), 1e-4f, pos[k]);
					b.ray.origin[k] = fmaf(xor_sign(gn[k], f2u(sr) & 0x80000000u), 1e-4f, pos[k]);
					a.ray.dir[k] = rl[k]; b.ray.dir[k] = rr[k];
					reflect &= a.ray.origin[k] == a.ray.origin[k] && rl[k] == rl[k];
					refract &= b.ray.origin[k] == b.ray.origin[k] && rr[k] == rr[k];
					a.weight[k] = b.weight[k] = weight[k];
				}
				a.ray.minT = b.ray.minT = 1e-3f; a.ray.maxT = b.ray.maxT = 1e+6f;
				a.depth = b.depth = it.depth + 1;
				if (top + 2 > 72) { x->err = -1; break; }
				if (refract) stack[top++] = b;
				if (reflect) stack[top++] = a;
			}
		}
		for (int k = 0; k < 3; ++k) x->fb[4 * pixel + k] += (float)acc[k] * (1.0f / FIXED_ONE);
	}
	if (x->waves)
		for (uint32_t dpt = 0; dpt <= x->max_depth && dpt < 64; ++dpt) __sync_fetch_and_add(&x->waves[dpt], waves[dpt]);
}

int oracle_whitted_trace(const oracle_scene* scene, const oracle_shading* shading, const oracle_camera* camera, uint32_t width,
                         uint32_t height, uint32_t sample_base, uint32_t spp, uint32_t max_depth, uint32_t seed, float* framebuffer4,
                         uint64_t* wave_rays, int threads) {
	if (max_depth > 62) return -1;
	wt_ctx x = { scene, shading, camera, width, height, sample_base, spp, max_depth, seed, framebuffer4, wave_rays, 0 };
	parallel_for(threads, ((int64_t)width * height + PT_CHUNK - 1) / PT_CHUNK, wt_body, &x);
	return x.err;
}
