// opencl_c.h -- TEST INFRASTRUCTURE ONLY: just enough of OpenCL C 1.2, as C++17, to compile the reference's
// traversal kernel FROM ITS OWN SOURCE TEXT (the string literal in /root/reference/RayAccelerator/Kernels.h:9-242,
// extracted at build time into oracle/_ref/traversal_kernel.cl by oracle/Makefile) and run it on the CPU, one
// work-item after the other. Nothing of the reference is copied into this repository: this header only supplies the
// language (vector types with swizzles, address-space keywords, built-ins).
//
// The reference builds its kernel with -cl-fast-relaxed-math and native_* built-ins (RayAccelerator.cpp:489), so its
// bits are implementation-defined. The built-in model chosen here is the oracle's pinned reading (racc_oracle.h):
//   mad(a,b,c)      = fmaf(a,b,c)                         dot(a,b) = fmaf(a.z,b.z, fmaf(a.y,b.y, a.x*b.x))
//   fmin/fmax       = minNum/maxNum, -0 < +0, subnormals flushed
//   native_recip(x) = 1.0f/x    native_rsqrt(x) = 1.0f/sqrtf(x)    acos = oracle_acosf
//   read_imagef     = OpenCL 1.2 section 8.2 linear filter, normalized coordinates, clamp to edge
// With that model the kernel's own control and data flow decide everything else, which is what the comparison with
// oracle_traverse() pins (tests/test_oracle_kat.py::test_oracle_matches_reference_kernel_source).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <xmmintrin.h>

extern "C" float oracle_acosf(float x); // oracle/racc_oracle.c

namespace ocl {

// ---- scalar and vector types -------------------------------------------------------------------------
template <int N, int A>
struct Sw1 { // one component of an N-float vector, as an l-value
	float d[N];
	operator float() const { return d[A]; }
	Sw1& operator=(float f) { d[A] = f; return *this; }
	Sw1& operator=(const Sw1& o) { d[A] = o.d[A]; return *this; }
};

struct float2;
struct float3;

template <int N, int A, int B>
struct Sw2 {
	union { float d[N]; Sw1<N, A> x; Sw1<N, B> y; };
	inline operator float2() const;
	inline Sw2& operator=(const float2& v);
	template <int M, int P, int Q> Sw2& operator=(const Sw2<M, P, Q>& o) { const float a = o.d[P], b = o.d[Q]; d[A] = a; d[B] = b; return *this; }
	Sw2& operator=(const Sw2& o) { const float a = o.d[A], b = o.d[B]; d[A] = a; d[B] = b; return *this; }
};

template <int N, int A, int B, int C>
struct Sw3 {
	union { float d[N]; Sw1<N, A> x; Sw1<N, B> y; Sw1<N, C> z; };
	inline operator float3() const;
	inline Sw3& operator=(const float3& v);
	template <int M, int P, int Q, int R> Sw3& operator=(const Sw3<M, P, Q, R>& o) {
		const float a = o.d[P], b = o.d[Q], c = o.d[R];
		d[A] = a; d[B] = b; d[C] = c;
		return *this;
	}
	Sw3& operator=(const Sw3& o) { const float a = o.d[A], b = o.d[B], c = o.d[C]; d[A] = a; d[B] = b; d[C] = c; return *this; }
	inline float3 operator-() const;
};

struct float2 {
	union { float d[2]; Sw1<2, 0> x; Sw1<2, 1> y; };
	float2() { d[0] = d[1] = 0.0f; }
	float2(float a, float b) { d[0] = a; d[1] = b; }
	float2(const float2& o) { std::memcpy(d, o.d, sizeof d); }
	float2& operator=(const float2& o) { std::memcpy(d, o.d, sizeof d); return *this; }
};

struct float3 {
	union {
		float d[3];
		Sw1<3, 0> x; Sw1<3, 1> y; Sw1<3, 2> z;
		Sw2<3, 0, 1> xy;
		Sw3<3, 2, 0, 1> zxy;
		Sw3<3, 1, 2, 0> yzx;
	};
	float3() { d[0] = d[1] = d[2] = 0.0f; }
	float3(float a, float b, float c) { d[0] = a; d[1] = b; d[2] = c; }
	float3(const float3& o) { std::memcpy(d, o.d, sizeof d); }
	float3& operator=(const float3& o) { std::memcpy(d, o.d, sizeof d); return *this; }
};

struct float4 {
	union {
		float d[4];
		Sw1<4, 0> x; Sw1<4, 1> y; Sw1<4, 2> z; Sw1<4, 3> w;
		Sw3<4, 0, 1, 2> xyz;
		Sw3<4, 1, 2, 3> yzw;
		Sw2<4, 2, 3> zw;
	};
	float4() { d[0] = d[1] = d[2] = d[3] = 0.0f; }
	float4(float a, float b, float c, float e) { d[0] = a; d[1] = b; d[2] = c; d[3] = e; }
	float4(const float4& o) { std::memcpy(d, o.d, sizeof d); }
	float4& operator=(const float4& o) { std::memcpy(d, o.d, sizeof d); return *this; }
};

struct float8 {
	union {
		float d[8];
		Sw1<8, 3> w;
		Sw1<8, 7> s7;
		Sw3<8, 0, 1, 2> xyz;
		Sw3<8, 4, 5, 6> s456;
	};
	float8() { std::memset(d, 0, sizeof d); }
	float8(const float8& o) { std::memcpy(d, o.d, sizeof d); }
	float8& operator=(const float8& o) { std::memcpy(d, o.d, sizeof d); return *this; }
};

template <int N, int A, int B> inline Sw2<N, A, B>::operator float2() const { return float2(d[A], d[B]); }
template <int N, int A, int B> inline Sw2<N, A, B>& Sw2<N, A, B>::operator=(const float2& v) { d[A] = v.d[0]; d[B] = v.d[1]; return *this; }
template <int N, int A, int B, int C> inline Sw3<N, A, B, C>::operator float3() const { return float3(d[A], d[B], d[C]); }
template <int N, int A, int B, int C> inline Sw3<N, A, B, C>& Sw3<N, A, B, C>::operator=(const float3& v) {
	d[A] = v.d[0]; d[B] = v.d[1]; d[C] = v.d[2];
	return *this;
}
template <int N, int A, int B, int C> inline float3 Sw3<N, A, B, C>::operator-() const { return float3(-d[A], -d[B], -d[C]); }

inline float2 make_float2(float x, float y) { return float2(x, y); }
inline float3 make_float3(float x, float y, float z) { return float3(x, y, z); }
inline float4 make_float4(float x, float y, float z, float w) { return float4(x, y, z, w); }

inline float3 operator-(const float3& a, const float3& b) { return float3(a.d[0] - b.d[0], a.d[1] - b.d[1], a.d[2] - b.d[2]); }
inline float3 operator*(const float3& a, const float3& b) { return float3(a.d[0] * b.d[0], a.d[1] * b.d[1], a.d[2] * b.d[2]); }
inline float3 operator-(const float3& a) { return float3(-a.d[0], -a.d[1], -a.d[2]); }

// ---- built-ins (the pinned model, see the header comment) --------------------------------------------
inline int as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
inline float as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
inline float as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }

inline float flushed(float x) {
	uint32_t u;
	std::memcpy(&u, &x, 4);
	if (u & 0x7f800000u) return x;
	u &= 0x80000000u;
	std::memcpy(&x, &u, 4);
	return x;
}
inline float fmin(float a, float b) {
	a = flushed(a); b = flushed(b);
	if (a != a) return b;
	if (b != b) return a;
	if (a < b) return a;
	if (b < a) return b;
	return as_float(as_int(a) | as_int(b));
}
inline float fmax(float a, float b) {
	a = flushed(a); b = flushed(b);
	if (a != a) return b;
	if (b != b) return a;
	if (a > b) return a;
	if (b > a) return b;
	return as_float(as_int(a) & as_int(b));
}
inline float3 fmin(const float3& a, const float3& b) { return float3(fmin(a.d[0], b.d[0]), fmin(a.d[1], b.d[1]), fmin(a.d[2], b.d[2])); }
inline float3 fmax(const float3& a, const float3& b) { return float3(fmax(a.d[0], b.d[0]), fmax(a.d[1], b.d[1]), fmax(a.d[2], b.d[2])); }
#ifndef RACC_SHIM_RELAXED
inline float mad(float a, float b, float c) { return ::fmaf(a, b, c); }
// The result passes through an empty asm so that a unary minus applied to it (Kernels.h:65-66: -dot(R, e1), -dot(R, e3)) stays the
// separate IEEE negation the OpenCL C source says: g++ otherwise folds it into the last fma (vfnmsub), whose exact-cancellation
// zero is +0 where -(+0) is -0 -- and that sign decides which of two triangles owns their shared edge.
inline float dot(const float3& a, const float3& b) {
	float r = ::fmaf(a.d[2], b.d[2], ::fmaf(a.d[1], b.d[1], a.d[0] * b.d[0]));
	__asm__("" : "+x"(r));
	return r;
}
#else
// A second, equally legal reading of what -cl-fast-relaxed-math leaves open, used only to measure how far another
// implementation's results can lie from the pinned ones (tests/test_oracle_kat.py::test_relaxed_builtin_model_*):
// mad as a separately rounded multiply and add, dot summed left to right without fusion, native_recip / native_rsqrt
// from the 12-bit hardware estimates refined by one Newton step (about 22 bits, not correctly rounded).
inline float mad(float a, float b, float c) { return a * b + c; }
inline float dot(const float3& a, const float3& b) { return (a.d[0] * b.d[0] + a.d[1] * b.d[1]) + a.d[2] * b.d[2]; }
#endif
inline float3 mad(const float3& a, const float3& b, const float3& c) {
	return float3(mad(a.d[0], b.d[0], c.d[0]), mad(a.d[1], b.d[1], c.d[1]), mad(a.d[2], b.d[2], c.d[2]));
}
inline float fabs(float x) { return ::fabsf(x); }
inline float copysign(float a, float b) { return ::copysignf(a, b); }
inline int signbit(float x) { return as_int(x) < 0 ? 1 : 0; }
inline int select(int a, int b, int c) { return c ? b : a; }
#ifndef RACC_SHIM_RELAXED
inline float native_recip(float x) { return 1.0f / x; }
inline float native_rsqrt(float x) { return 1.0f / ::sqrtf(x); }
#else
inline float native_recip(float x) {
	const float e = _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(x)));
	return e * (2.0f - x * e);
}
inline float native_rsqrt(float x) {
	const float e = _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(x)));
	return e * (1.5f - 0.5f * x * e * e);
}
#endif
inline float acos(float x) { return oracle_acosf(x); }

// ---- images ------------------------------------------------------------------------------------------
typedef int sampler_t;
enum { CLK_NORMALIZED_COORDS_TRUE = 1, CLK_ADDRESS_CLAMP_TO_EDGE = 2, CLK_FILTER_LINEAR = 4 };
struct image2d { const float* texels; int width, height; }; // RGBA32F
typedef const image2d* image2d_t;

inline int texelFloor(float x, float* frac) {
	const float f = ::floorf(x);
	*frac = x - f;
	if (!(f > -4.0f)) { if (f != f) { *frac = 0.0f; return 0; } return -4; }
	if (f > 1.0e9f) return 1000000000;
	return (int)f;
}
inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
// OpenCL 1.2 section 8.2, CLK_FILTER_LINEAR with normalized coordinates and CLK_ADDRESS_CLAMP_TO_EDGE
inline float4 read_imagef(image2d_t img, sampler_t, const float2& uv) {
	if (!img || !img->texels) return float4(0.0f, 0.0f, 0.0f, 0.0f);
	const float fu = uv.d[0] * (float)img->width - 0.5f, fv = uv.d[1] * (float)img->height - 0.5f;
	float a, b;
	int i0 = texelFloor(fu, &a), j0 = texelFloor(fv, &b);
	const int i1 = clampi(i0 + 1, 0, img->width - 1), j1 = clampi(j0 + 1, 0, img->height - 1);
	i0 = clampi(i0, 0, img->width - 1);
	j0 = clampi(j0, 0, img->height - 1);
	const float* t00 = img->texels + 4 * ((size_t)j0 * img->width + i0);
	const float* t10 = img->texels + 4 * ((size_t)j0 * img->width + i1);
	const float* t01 = img->texels + 4 * ((size_t)j1 * img->width + i0);
	const float* t11 = img->texels + 4 * ((size_t)j1 * img->width + i1);
	const float na = 1.0f - a, nb = 1.0f - b;
	const float w00 = na * nb, w10 = a * nb, w01 = na * b, w11 = a * b;
	float4 o;
	for (int k = 0; k < 4; ++k) o.d[k] = ((w00 * t00[k] + w10 * t10[k]) + w01 * t01[k]) + w11 * t11[k];
	return o;
}

// ---- work-items --------------------------------------------------------------------------------------
extern thread_local int g_globalId;
inline int get_global_id(int) { return g_globalId; }

} // namespace ocl

#define kernel
#define global
#define __read_only
