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
};

// One warp-level rendezvous per distinct member mask (a kernel may have lanes inside `if (valid) __match_any_sync(mask, ..)`
// while the others already wait in a full-mask __syncwarp()). Values are exchanged through two slot sets that alternate
// with the generation, so one rendezvous per collective is enough.
struct Collective {
	unsigned mask = 0, arrived = 0, generation = 0;
	unsigned participants[2] = {0, 0}; // who took part in the generation that used this slot set
	uint64_t slot[2][32];
};

struct Warp {
	unsigned alive = 0xffffffffu; // lanes that have not left the kernel
	unsigned used = 0;
	Collective collective[16];
};

struct State {
	ucontext_t scheduler;
	std::vector<Fiber> fibers;
	std::vector<char> stacks;
	std::vector<Warp> warps;
	std::vector<char> dynamicShared;
	Rendezvous block;
	Fiber* current = nullptr;
	Dim blockIdx_ = {0, 0, 0}, blockDim_ = {1, 1, 1}, gridDim_ = {1, 1, 1};
	const std::function<void()>* body = nullptr;
};

// per OS thread: host threads (the API's submitters) may launch concurrently, each runs its own fibers
inline thread_local State g;

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

inline void release_if_complete(const Warp& w, Collective& c) {
	const unsigned expected = c.mask & w.alive;
	if (c.arrived && (c.arrived & w.alive) == expected) {
		c.participants[c.generation & 1] = expected;
		c.arrived = 0;
		++c.generation;
	}
}

inline void fiber_main() {
	Fiber* self = g.current;
	(*g.body)();
	self->done = true;
	// a thread that has left the kernel no longer takes part in barriers (CUDA: exited threads count as arrived)
	Warp& w = g.warps[self->tid.x / 32];
	w.alive &= ~(1u << (self->tid.x % 32));
	for (unsigned k = 0; k < w.used; ++k) release_if_complete(w, w.collective[k]);
	--g.block.alive; release_if_complete(g.block);
	swapcontext(&self->context, &g.scheduler);
}

// kernel<<<grid, block>>>(args) -> launch(grid, block, [=] { kernel(args); })
inline void launch(unsigned grid, unsigned block, const std::function<void()>& body, size_t dynamicSharedBytes = 0) {
	const size_t kStack = 256 * 1024;
	if (block == 0 || block % 32 != 0 || block > 1024) abort();
	g.fibers.assign(block, Fiber());
	g.stacks.resize(kStack * block);
	g.warps.resize(block / 32);
	g.dynamicShared.assign(dynamicSharedBytes + 128, 0);
	g.body = &body;
	g.blockDim_ = {block, 1, 1};
	g.gridDim_ = {grid, 1, 1};
	for (unsigned b = 0; b < grid; ++b) {
		g.blockIdx_ = {b, 0, 0};
		g.block = Rendezvous();
		g.block.alive = block;
		for (Warp& w : g.warps) { w.alive = 0xffffffffu; w.used = 0; }
		for (unsigned t = 0; t < block; ++t) {
			Fiber& f = g.fibers[t];
			f.tid = {t, 0, 0};
			f.done = false;
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

// every lane named in `mask` deposits a value, then reads any member's; returns the slot set and who took part
struct Exchanged { const uint64_t* slot; unsigned participants; };
inline Exchanged exchange(unsigned mask, uint64_t mine) {
	Fiber* self = g.current;
	Warp& w = g.warps[self->tid.x / 32];
	const unsigned lane = self->tid.x % 32;
	if (!(mask >> lane & 1u)) abort(); // a lane must name itself
	unsigned k = 0;
	while (k < w.used && w.collective[k].mask != mask) ++k;
	if (k == w.used) {
		if (w.used == 16) abort();
		w.collective[k] = Collective();
		w.collective[k].mask = mask;
		++w.used;
	}
	Collective& c = w.collective[k];
	const unsigned generation = c.generation;
	c.slot[generation & 1][lane] = mine;
	c.arrived |= 1u << lane;
	release_if_complete(w, c);
	while (c.generation == generation) yield();
	return Exchanged{c.slot[generation & 1], c.participants[generation & 1]};
}

inline void* dynamic_shared() { return reinterpret_cast<void*>(((uintptr_t)g.dynamicShared.data() + 127) & ~(uintptr_t)127); }

template <typename T> inline uint64_t to_bits(T v) { static_assert(sizeof(T) <= 8, ""); uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

} // namespace cuda_on_cpu

#define threadIdx (::cuda_on_cpu::g.current->tid)
#define blockIdx (::cuda_on_cpu::g.blockIdx_)
#define blockDim (::cuda_on_cpu::g.blockDim_)
#define gridDim (::cuda_on_cpu::g.gridDim_)

static inline void __syncthreads() { ::cuda_on_cpu::rendezvous(::cuda_on_cpu::g.block); }

static inline unsigned __activemask() { return ::cuda_on_cpu::g.warps[threadIdx.x / 32].alive; } // lanes still in the kernel
static inline void __syncwarp(unsigned mask = 0xffffffffu) { ::cuda_on_cpu::exchange(mask, 0); }
static inline unsigned __ballot_sync(unsigned mask, bool predicate) {
	const ::cuda_on_cpu::Exchanged e = ::cuda_on_cpu::exchange(mask, predicate ? 1u : 0u);
	unsigned r = 0;
	for (int l = 0; l < 32; ++l) if (e.participants >> l & 1u) r |= (unsigned)(e.slot[l] & 1u) << l;
	return r;
}
template <typename T> static inline unsigned __match_any_sync(unsigned mask, T v) {
	const uint64_t mine = ::cuda_on_cpu::to_bits(v);
	const ::cuda_on_cpu::Exchanged e = ::cuda_on_cpu::exchange(mask, mine);
	unsigned r = 0;
	for (int l = 0; l < 32; ++l) if ((e.participants >> l & 1u) && e.slot[l] == mine) r |= 1u << l;
	return r;
}
template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src) {
	const ::cuda_on_cpu::Exchanged e = ::cuda_on_cpu::exchange(mask, ::cuda_on_cpu::to_bits(v));
	return ::cuda_on_cpu::from_bits<T>(e.slot[src & 31]);
}
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta) {
	const int lane = (int)(threadIdx.x % 32);
	const ::cuda_on_cpu::Exchanged e = ::cuda_on_cpu::exchange(mask, ::cuda_on_cpu::to_bits(v));
	return lane - (int)delta >= 0 ? ::cuda_on_cpu::from_bits<T>(e.slot[lane - (int)delta]) : v;
}
template <typename T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta) {
	const int lane = (int)(threadIdx.x % 32);
	const ::cuda_on_cpu::Exchanged e = ::cuda_on_cpu::exchange(mask, ::cuda_on_cpu::to_bits(v));
	return lane + (int)delta < 32 ? ::cuda_on_cpu::from_bits<T>(e.slot[lane + (int)delta]) : v;
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
	const int lane = (int)(threadIdx.x % 32);
	const ::cuda_on_cpu::Exchanged e = ::cuda_on_cpu::exchange(mask, ::cuda_on_cpu::to_bits(v));
	return ::cuda_on_cpu::from_bits<T>(e.slot[(lane ^ laneMask) & 31]);
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
