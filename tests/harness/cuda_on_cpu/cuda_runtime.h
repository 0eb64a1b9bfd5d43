// cuda_runtime.h -- TEST INFRASTRUCTURE ONLY: a stand-in for the CUDA runtime header that lets g++ compile the
// engine's *shading* kernels (rayaccel_b200/csrc/pathtrace.cu, whitted.cu) unchanged and run them on the CPU, so that
// the CPU-only test suite executes the kernels' own source against the checker (tests/test_kernels_on_cpu.py) -- the
// same idea as oracle/ref_shim/opencl_c.h for the reference's OpenCL kernel. Nothing in the product includes this.
//
// Execution model: a launch runs its blocks one after the other; the threads of a block are user-level fibers
// (ucontext) scheduled round-robin on the calling OS thread, so warp collectives (__ballot_sync, __shfl_*_sync) and
// __syncthreads are real rendezvous points and divergence is whatever the kernel's control flow makes it. Only what the
// two files use is provided. Floating point: the harness is compiled -ffp-contract=off -mfma and runs with FTZ/DAZ set,
// which is the arithmetic nvcc is held to by -fmad=false -ftz=true -prec-div=true -prec-sqrt=true.
#pragma once

#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 v = {x, y, z, w}; return v; }

typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef void* cudaStream_t;
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }

#define __global__ static
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static

namespace cuda_on_cpu {

struct Dim { unsigned x, y, z; };

struct Rendezvous {
	unsigned alive = 0, arrived = 0, generation = 0;
};

struct Fiber {
	ucontext_t context;
	Dim tid = {0, 0, 0};
	bool done = false;
	int parity = 0; // which half of the warp's exchange slots the next collective uses
};

struct Warp {
	Rendezvous rendezvous;
	uint64_t slot[2][32];
};

struct State {
	ucontext_t scheduler;
	std::vector<Fiber> fibers;
	std::vector<char> stacks;
	std::vector<Warp> warps;
	Rendezvous block;
	Fiber* current = nullptr;
	Dim blockIdx_ = {0, 0, 0}, blockDim_ = {1, 1, 1}, gridDim_ = {1, 1, 1};
	const std::function<void()>* body = nullptr;
};

inline State g;

inline void yield() { swapcontext(&g.current->context, &g.scheduler); }

inline void release_if_complete(Rendezvous& r) {
	if (r.alive && r.arrived == r.alive) { r.arrived = 0; ++r.generation; }
}

inline void rendezvous(Rendezvous& r) {
	const unsigned generation = r.generation;
	++r.arrived;
	release_if_complete(r);
	while (r.generation == generation) yield();
}

inline void fiber_main() {
	Fiber* self = g.current;
	(*g.body)();
	self->done = true;
	// a thread that has left the kernel no longer takes part in barriers (CUDA: exited threads count as arrived)
	Rendezvous& w = g.warps[self->tid.x / 32].rendezvous;
	--w.alive; release_if_complete(w);
	--g.block.alive; release_if_complete(g.block);
	swapcontext(&self->context, &g.scheduler);
}

// kernel<<<grid, block>>>(args) -> launch(grid, block, [=] { kernel(args); })
inline void launch(unsigned grid, unsigned block, const std::function<void()>& body) {
	const size_t kStack = 256 * 1024;
	if (block == 0 || block % 32 != 0 || block > 1024) abort();
	g.fibers.assign(block, Fiber());
	g.stacks.resize(kStack * block);
	g.warps.assign(block / 32, Warp());
	g.body = &body;
	g.blockDim_ = {block, 1, 1};
	g.gridDim_ = {grid, 1, 1};
	for (unsigned b = 0; b < grid; ++b) {
		g.blockIdx_ = {b, 0, 0};
		g.block = Rendezvous();
		g.block.alive = block;
		for (Warp& w : g.warps) { w.rendezvous = Rendezvous(); w.rendezvous.alive = 32; }
		for (unsigned t = 0; t < block; ++t) {
			Fiber& f = g.fibers[t];
			f.tid = {t, 0, 0};
			f.done = false;
			f.parity = 0;
			getcontext(&f.context);
			f.context.uc_stack.ss_sp = g.stacks.data() + kStack * t;
			f.context.uc_stack.ss_size = kStack;
			f.context.uc_link = &g.scheduler;
			makecontext(&f.context, fiber_main, 0);
		}
		unsigned remaining = block;
		while (remaining) {
			for (unsigned t = 0; t < block; ++t) {
				Fiber& f = g.fibers[t];
				if (f.done) continue;
				g.current = &f;
				swapcontext(&g.scheduler, &f.context);
				if (f.done) --remaining;
			}
		}
	}
	g.current = nullptr;
	g.body = nullptr;
}

// every lane of the warp deposits a value, then reads any lane's; one rendezvous per collective (slots alternate)
inline const uint64_t* exchange(uint64_t mine) {
	Fiber* self = g.current;
	Warp& w = g.warps[self->tid.x / 32];
	const int parity = self->parity;
	self->parity ^= 1;
	w.slot[parity][self->tid.x % 32] = mine;
	rendezvous(w.rendezvous);
	return w.slot[parity];
}

template <typename T> inline uint64_t to_bits(T v) { static_assert(sizeof(T) <= 8, ""); uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

} // namespace cuda_on_cpu

#define threadIdx (::cuda_on_cpu::g.current->tid)
#define blockIdx (::cuda_on_cpu::g.blockIdx_)
#define blockDim (::cuda_on_cpu::g.blockDim_)
#define gridDim (::cuda_on_cpu::g.gridDim_)

static inline void __syncthreads() { ::cuda_on_cpu::rendezvous(::cuda_on_cpu::g.block); }

// Only full-mask collectives are supported (all the two files use); anything else aborts instead of guessing.
static inline unsigned __ballot_sync(unsigned mask, bool predicate) {
	if (mask != 0xffffffffu) abort();
	const uint64_t* s = ::cuda_on_cpu::exchange(predicate ? 1u : 0u);
	unsigned r = 0;
	for (int l = 0; l < 32; ++l) r |= (unsigned)(s[l] & 1u) << l;
	return r;
}
template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src) {
	if (mask != 0xffffffffu) abort();
	const uint64_t* s = ::cuda_on_cpu::exchange(::cuda_on_cpu::to_bits(v));
	return ::cuda_on_cpu::from_bits<T>(s[src & 31]);
}
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta) {
	if (mask != 0xffffffffu) abort();
	const int lane = (int)(threadIdx.x % 32);
	const uint64_t* s = ::cuda_on_cpu::exchange(::cuda_on_cpu::to_bits(v));
	return lane - (int)delta >= 0 ? ::cuda_on_cpu::from_bits<T>(s[lane - (int)delta]) : v;
}
template <typename T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta) {
	if (mask != 0xffffffffu) abort();
	const int lane = (int)(threadIdx.x % 32);
	const uint64_t* s = ::cuda_on_cpu::exchange(::cuda_on_cpu::to_bits(v));
	return lane + (int)delta < 32 ? ::cuda_on_cpu::from_bits<T>(s[lane + (int)delta]) : v;
}

static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned long long __float2ull_rn(float f) { return (unsigned long long)llrintf(f); } // callers pass 0 <= f < 2^63
static inline float __ull2float_rn(unsigned long long v) { return (float)v; }
template <typename T> static inline T __ldg(const T* p) { return *p; }

// one OS thread runs every fiber, and fibers only switch inside collectives: plain read-modify-write is atomic here
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { const uint32_t old = *p; *p = old + v; return old; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long old = *p; *p = old + v; return old; }

// ---- what the traversal kernels need on top of the shading kernels ------------------------------------------------
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int laneMask) {
	if (mask != 0xffffffffu) abort();
	const int lane = (int)(threadIdx.x % 32);
	const uint64_t* s = ::cuda_on_cpu::exchange(::cuda_on_cpu::to_bits(v));
	return ::cuda_on_cpu::from_bits<T>(s[(lane ^ laneMask) & 31]);
}
// IEEE round-to-nearest reciprocal, square root and division (the process runs FTZ/DAZ like the device code)
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline float __fsqrt_rn(float x) { return sqrtf(x); }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned v) { const unsigned long long old = *p; *p = old + v; return old; }
template <typename T> static inline T min(T a, T b) { return b < a ? b : a; }
template <typename T> static inline T max(T a, T b) { return a < b ? b : a; }

namespace cuda_on_cpu {
// FMNMX: minNum / maxNum, subnormal inputs flushed, -0 < +0 (what fminf/fmaxf compile to under -ftz=true)
inline float flush_subnormal(float x) {
	uint32_t u; memcpy(&u, &x, 4);
	if (u & 0x7f800000u) return x;
	u &= 0x80000000u; memcpy(&x, &u, 4);
	return x;
}
inline float device_fminf(float a, float b) {
	a = flush_subnormal(a); b = flush_subnormal(b);
	if (a != a) return b;
	if (b != b) return a;
	if (a < b) return a;
	if (b < a) return b;
	uint32_t ua, ub; memcpy(&ua, &a, 4); memcpy(&ub, &b, 4); ua |= ub; memcpy(&a, &ua, 4);
	return a;
}
inline float device_fmaxf(float a, float b) {
	a = flush_subnormal(a); b = flush_subnormal(b);
	if (a != a) return b;
	if (b != b) return a;
	if (a > b) return a;
	if (b > a) return b;
	uint32_t ua, ub; memcpy(&ua, &a, 4); memcpy(&ub, &b, 4); ua &= ub; memcpy(&a, &ua, 4);
	return a;
}

// Address windows. Kernel-local arrays live on the fibers' stacks, __shared__ arrays are statics of this library:
// a 32-bit "local" / "shared" address is the distance from a fixed anchor of that region.
inline char shared_anchor;
inline uint32_t local_address(const void* p) { return (uint32_t)((uintptr_t)p - (uintptr_t)g.stacks.data()); }
inline uint32_t shared_address(const void* p) { return (uint32_t)((intptr_t)p - (intptr_t)&shared_anchor); }

namespace ptx {
inline uint32_t* local_word(unsigned long long address) { return reinterpret_cast<uint32_t*>(g.stacks.data() + (uint32_t)address); }
inline uint32_t* shared_word(unsigned long long address) { return reinterpret_cast<uint32_t*>(&shared_anchor + (int32_t)(uint32_t)address); }
inline void pack2(unsigned long long& out, float lo, float hi) {
	uint32_t a, b; memcpy(&a, &lo, 4); memcpy(&b, &hi, 4);
	out = (unsigned long long)a | ((unsigned long long)b << 32);
}
inline void unpack2(float& lo, float& hi, unsigned long long v) {
	const uint32_t a = (uint32_t)v, b = (uint32_t)(v >> 32); memcpy(&lo, &a, 4); memcpy(&hi, &b, 4);
}
inline void unpack2(uint32_t& lo, uint32_t& hi, unsigned long long v) { lo = (uint32_t)v; hi = (uint32_t)(v >> 32); }
// fma.rn.ftz.f32x2: one IEEE fma per half, inputs and results flushed (MXCSR FTZ/DAZ does that to fmaf here)
inline void fma2(unsigned long long& out, unsigned long long a, unsigned long long b, unsigned long long c) {
	float al, ah, bl, bh, cl, ch;
	unpack2(al, ah, a); unpack2(bl, bh, b); unpack2(cl, ch, c);
	pack2(out, fmaf(al, bl, cl), fmaf(ah, bh, ch));
}
inline void ld8f(unsigned long long address, float& a, float& b, float& c, float& d, float& e, float& f, float& g_, float& h) {
	if (address & 31) abort(); // a 256-bit load must be 32-byte aligned
	float v[8]; memcpy(v, reinterpret_cast<const void*>((uintptr_t)address), 32);
	a = v[0]; b = v[1]; c = v[2]; d = v[3]; e = v[4]; f = v[5]; g_ = v[6]; h = v[7];
}
inline void ld4q(unsigned long long address, unsigned long long& a, unsigned long long& b, unsigned long long& c, unsigned long long& d) {
	if (address & 31) abort();
	unsigned long long v[4]; memcpy(v, reinterpret_cast<const void*>((uintptr_t)address), 32);
	a = v[0]; b = v[1]; c = v[2]; d = v[3];
}
[[noreturn]] inline void unsupported(const char* what) { fprintf(stderr, "cuda_on_cpu: PTX not available on the CPU: %s\n", what); abort(); }
} // namespace ptx
} // namespace cuda_on_cpu

#define fminf(a, b) ::cuda_on_cpu::device_fminf((a), (b))
#define fmaxf(a, b) ::cuda_on_cpu::device_fmaxf((a), (b))
static inline size_t __cvta_generic_to_local(const void* p) { return ::cuda_on_cpu::local_address(p); }
static inline size_t __cvta_generic_to_shared(const void* p) { return ::cuda_on_cpu::shared_address(p); }

// the few runtime calls the launchers make
enum cudaDeviceAttr { cudaDevAttrMaxSharedMemoryPerMultiprocessor = 81, cudaDevAttrMultiProcessorCount = 16 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
static inline cudaError_t cudaGetDevice(int* device) { *device = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* value, cudaDeviceAttr attr, int) { *value = attr == cudaDevAttrMultiProcessorCount ? 2 : 233472; return cudaSuccess; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <typename F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* blocks, F, int, size_t) { *blocks = 2; return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int value, size_t bytes, cudaStream_t) { memset(p, value, bytes); return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "cuda_on_cpu"; }
