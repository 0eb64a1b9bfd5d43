// cuda_runtime.h -- TEST INFRASTRUCTURE ONLY: a stand-in for the CUDA runtime header that lets g++ compile the engine's
// .cu files (rayaccel_b200/csrc: the traversal kernels, the shading kernels, the radix sort, the device scene builder and
// capi.cu itself) as they are and run them on the CPU, so that the CPU-only test suite executes the kernels' OWN SOURCE
// against the checker (tests/test_kernels_on_cpu.py, tests/test_library_on_cpu.py) -- the same idea
// as oracle/ref_shim/opencl_c.h for the reference's OpenCL kernel. Nothing in the product includes this.
//
// Execution model: a launch runs its blocks one after the other; the threads of a block are user-level fibers scheduled
// round-robin on the calling OS thread, so warp collectives (__ballot_sync, __shfl_*_sync, __match_any_sync,
// __syncwarp(mask): one rendezvous per member mask) and __syncthreads are real rendezvous points and divergence is
// whatever the kernel's control flow makes it. __shared__ is a per-host-thread static; 32-bit local / shared "window"
// addresses map onto the fibers' stacks and those statics. The runtime API is a synchronous host equivalent (device
// memory is host memory). Inline PTX and <<<...>>> launches are rewritten by rewrite.py into the functions of namespace
// cuda_on_cpu(::ptx) below. Only what the engine's files use is provided; anything else aborts loudly.
// Floating point: built -ffp-contract=off -mfma and run with FTZ/DAZ set, fminf/fmaxf = FMNMX -- the arithmetic nvcc is
// held to by -fmad=false -ftz=true -prec-div=true -prec-sqrt=true.
#pragma once

#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <ucontext.h>

#include <functional>
#include <vector>

struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 v = {x, y, z, w}; return v; }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned selector) { // PRMT, default mode
	const unsigned long long both = ((unsigned long long)b << 32) | a;
	unsigned r = 0;
	for (int k = 0; k < 4; ++k) r |= (unsigned)((both >> (8 * ((selector >> (4 * k)) & 7))) & 0xff) << (8 * k);
	return r;
}
struct uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 v = {x, y, z, w}; return v; }

typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef void* cudaStream_t;
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }

#define __global__ static
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
// per host thread: the API's submitters launch concurrently, each must see its own block's shared memory
#define __shared__ static thread_local

// Fiber switch. By default a dozen instructions of x86-64 assembly (callee-saved registers + stack pointer; every fiber
// runs with the same MXCSR / x87 control words, so those are not switched): barrier-heavy kernels spend their time here,
// and swapcontext() costs two system calls per switch. -DCUDA_ON_CPU_UCONTEXT selects swapcontext (what the sanitizer build
// uses: AddressSanitizer understands it).
#if !defined(CUDA_ON_CPU_UCONTEXT) && defined(__x86_64__)
#define CUDA_ON_CPU_ASM_SWITCH 1
extern "C" void cuda_on_cpu_switch(void** saveStackPointer, void* loadStackPointer);
__asm__(".text\n"
        ".weak cuda_on_cpu_switch\n"
        ".type cuda_on_cpu_switch,@function\n"
        "cuda_on_cpu_switch:\n"
        "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
        "  movq %rsp, (%rdi)\n"
        "  movq %rsi, %rsp\n"
        "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
        "  ret\n"
        ".size cuda_on_cpu_switch,.-cuda_on_cpu_switch\n");
#endif

namespace cuda_on_cpu {

struct Dim { unsigned x, y, z; };

struct Rendezvous {
	unsigned alive = 0, arrived = 0, generation = 0;
};

struct Fiber {
#ifdef CUDA_ON_CPU_ASM_SWITCH
	void* stackPointer = nullptr;
#else
	ucontext_t context;
#endif
	Dim tid = {0, 0, 0};
	bool done = false;
	uint32_t collectives = 0; // warp collectives this lane has taken part in: lanes of a warp between the same two votes agree
};

// Optional load tracing for the L1 gather model (tests/harness/l1_model.py): every 256-bit gather of a kernel is reported with the
// warp, the lane's collective count (lanes between the same two warp votes execute the same instruction instance) and
// its address. Off unless a sink is installed.
typedef void (*LoadSink)(uint32_t block, uint32_t warp, uint32_t epoch, uint32_t lane, unsigned long long address, uint32_t bytes);
inline LoadSink load_sink = nullptr;

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
#ifdef CUDA_ON_CPU_ASM_SWITCH
	void* scheduler = nullptr; // the scheduler's saved stack pointer while a fiber runs
#else
	ucontext_t scheduler;
#endif
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

#ifdef CUDA_ON_CPU_ASM_SWITCH
inline void yield() { cuda_on_cpu_switch(&g.current->stackPointer, g.scheduler); }
#else
inline void yield() { swapcontext(&g.current->context, &g.scheduler); }
#endif

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
	for (;;) yield(); // never resumed: the scheduler skips fibers that are done
}

// kernel<<<grid, block>>>(args) -> launch(grid, block, [=] { kernel(args); })
inline bool trace_launches() { static const bool on = getenv("CUDA_ON_CPU_TRACE") != nullptr; return on; } // per-launch wall time on stderr
inline void launch(unsigned grid, unsigned block, const std::function<void()>& body, size_t dynamicSharedBytes = 0) {
	timespec t0; clock_gettime(CLOCK_MONOTONIC, &t0);
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
			f.collectives = 0;
#ifdef CUDA_ON_CPU_ASM_SWITCH
			// first switch "returns" into fiber_main: [top-16] = return address (16-byte aligned slot, so that the stack is
			// aligned as after a call), six zeroed callee-saved registers below it
			char* top = g.stacks.data() + kStack * (t + 1);
			top -= (uintptr_t)top & 15;
			void** slot = reinterpret_cast<void**>(top - 16);
			slot[0] = reinterpret_cast<void*>(&fiber_main);
			for (int r = 1; r <= 6; ++r) slot[-r] = nullptr;
			f.stackPointer = slot - 6;
#else
			getcontext(&f.context);
			f.context.uc_stack.ss_sp = g.stacks.data() + kStack * t;
			f.context.uc_stack.ss_size = kStack;
			f.context.uc_link = &g.scheduler;
			makecontext(&f.context, fiber_main, 0);
#endif
		}
		unsigned remaining = block;
		while (remaining) {
			for (unsigned t = 0; t < block; ++t) {
				Fiber& f = g.fibers[t];
				if (f.done) continue;
				g.current = &f;
#ifdef CUDA_ON_CPU_ASM_SWITCH
				cuda_on_cpu_switch(&g.scheduler, f.stackPointer);
#else
				swapcontext(&g.scheduler, &f.context);
#endif
				if (f.done) --remaining;
			}
		}
	}
	g.current = nullptr;
	g.body = nullptr;
	if (trace_launches()) {
		timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1);
		fprintf(stderr, "cuda_on_cpu: launch %u x %u: %.1f ms\n", grid, block, (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6);
	}
}

// every lane named in `mask` deposits a value, then reads any member's; returns the slot set and who took part
struct Exchanged { const uint64_t* slot; unsigned participants; };
inline Exchanged exchange(unsigned mask, uint64_t mine) {
	Fiber* self = g.current;
	Warp& w = g.warps[self->tid.x / 32];
	const unsigned lane = self->tid.x % 32;
	if (!(mask >> lane & 1u)) { fprintf(stderr, "cuda_on_cpu: lane %u calls a collective with mask %08x that does not name it\n", lane, mask); abort(); }
	unsigned k = 0;
	while (k < w.used && w.collective[k].mask != mask) ++k;
	if (k == w.used) {
		if (w.used == 16) { fprintf(stderr, "cuda_on_cpu: more than 16 distinct member masks in one warp\n"); abort(); }
		w.collective[k] = Collective();
		w.collective[k].mask = mask;
		++w.used;
	}
	Collective& c = w.collective[k];
	++self->collectives;
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

// lanes still in the kernel. The other fibers get one turn first, so lanes that leave the kernel without waiting for
// anybody (`if (i >= n) return;`) have left -- as they have on hardware by the time a working lane asks.
static inline unsigned __activemask() { ::cuda_on_cpu::yield(); return ::cuda_on_cpu::g.warps[threadIdx.x / 32].alive; }
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
inline uint32_t local_address(const void* p) { return (uint32_t)((uintptr_t)p - (uintptr_t)g.stacks.data()); }
// shared window: a 32-bit "shared address" is (segment << 28 | offset). Segment 8 and up is the launch's dynamic block;
// segments 1-7 are 256 MB regions of this process registered the first time a __shared__ static inside them is converted
// (the statics of the libraries built over this header may lie anywhere in the address space).
inline thread_local uintptr_t shared_segment[7];
inline thread_local unsigned shared_segments = 0;
inline uint32_t shared_address(const void* p) {
	const char* c = static_cast<const char*>(p);
	const char* dyn = static_cast<const char*>(dynamic_shared());
	if (c >= dyn && c < dyn + g.dynamicShared.size()) return 0x80000000u + (uint32_t)(c - dyn);
	const uintptr_t a = (uintptr_t)c;
	for (unsigned k = 0; k < shared_segments; ++k)
		if (a >= shared_segment[k] + 0x00100000u && a < shared_segment[k] + 0x0ff00000u) return ((k + 1) << 28) | (uint32_t)(a - shared_segment[k]);
	if (shared_segments == 7) { fprintf(stderr, "cuda_on_cpu: too many shared-memory regions\n"); abort(); }
	shared_segment[shared_segments] = (a - 0x04000000u) & ~(uintptr_t)0xfffff;
	++shared_segments;
	return (shared_segments << 28) | (uint32_t)(a - shared_segment[shared_segments - 1]);
}

namespace ptx {
inline uint32_t* local_word(unsigned long long address) { return reinterpret_cast<uint32_t*>(g.stacks.data() + (uint32_t)address); }
inline char* shared_pointer(unsigned long long address) {
	const uint32_t a = (uint32_t)address;
	if (a & 0x80000000u) return static_cast<char*>(dynamic_shared()) + (a & 0x7fffffffu);
	const unsigned segment = a >> 28;
	if (segment == 0 || segment > shared_segments) { fprintf(stderr, "cuda_on_cpu: %08x is not a shared address\n", a); abort(); }
	return reinterpret_cast<char*>(shared_segment[segment - 1] + (a & 0x0fffffffu));
}
inline uint32_t* shared_word(unsigned long long address) { return reinterpret_cast<uint32_t*>(shared_pointer(address)); }
inline void pack2(unsigned long long& out, float lo, float hi) {
	uint32_t a, b; memcpy(&a, &lo, 4); memcpy(&b, &hi, 4);
	out = (unsigned long long)a | ((unsigned long long)b << 32);
}
inline void pack2(unsigned long long& out, unsigned lo, unsigned hi) { out = (unsigned long long)lo | ((unsigned long long)hi << 32); }
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
	if (address & 31) { fprintf(stderr, "cuda_on_cpu: misaligned 256-bit load\n"); abort(); }
	if (load_sink) load_sink(g.blockIdx_.x, g.current->tid.x / 32, g.current->collectives, g.current->tid.x % 32, address, 32);
	float v[8]; memcpy(v, reinterpret_cast<const void*>((uintptr_t)address), 32);
	a = v[0]; b = v[1]; c = v[2]; d = v[3]; e = v[4]; f = v[5]; g_ = v[6]; h = v[7];
}
inline void ld8u(unsigned long long address, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d, uint32_t& e, uint32_t& f, uint32_t& g_, uint32_t& h) {
	if (address & 31) { fprintf(stderr, "cuda_on_cpu: misaligned 256-bit load\n"); abort(); }
	if (load_sink) load_sink(g.blockIdx_.x, g.current->tid.x / 32, g.current->collectives, g.current->tid.x % 32, address, 32);
	uint32_t v[8]; memcpy(v, reinterpret_cast<const void*>((uintptr_t)address), 32);
	a = v[0]; b = v[1]; c = v[2]; d = v[3]; e = v[4]; f = v[5]; g_ = v[6]; h = v[7];
}
inline void ld4f(unsigned long long address, float& a, float& b, float& c, float& d) {
	if (address & 15) { fprintf(stderr, "cuda_on_cpu: misaligned 128-bit load\n"); abort(); }
	float v[4]; memcpy(v, reinterpret_cast<const void*>((uintptr_t)address), 16);
	a = v[0]; b = v[1]; c = v[2]; d = v[3];
}
inline void ld4q(unsigned long long address, unsigned long long& a, unsigned long long& b, unsigned long long& c, unsigned long long& d) {
	if (address & 31) { fprintf(stderr, "cuda_on_cpu: misaligned 256-bit load\n"); abort(); }
	if (load_sink) load_sink(g.blockIdx_.x, g.current->tid.x / 32, g.current->collectives, g.current->tid.x % 32, address, 32);
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
enum cudaDeviceAttr { cudaDevAttrMaxSharedMemoryPerMultiprocessor = 81, cudaDevAttrMultiProcessorCount = 16, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
// Several "devices" on request (CUDA_ON_CPU_DEVICES=n): they share the host's memory, which is all the multi-device host
// logic of capi.cu needs (replicas, dealing HOST streams over a device set, the frame reduction over a stand-in NCCL).
namespace cuda_on_cpu {
inline int deviceCount() { const char* v = getenv("CUDA_ON_CPU_DEVICES"); const int n = v ? atoi(v) : 1; return n < 1 ? 1 : n; }
inline int& currentDevice() { static thread_local int d = 0; return d; }
} // namespace cuda_on_cpu
static inline cudaError_t cudaGetDevice(int* device) { *device = ::cuda_on_cpu::currentDevice(); return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* value, cudaDeviceAttr attr, int) { *value = attr == cudaDevAttrMultiProcessorCount ? 2 : 233472; return cudaSuccess; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <typename F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* blocks, F, int, size_t) { *blocks = 2; return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int value, size_t bytes, cudaStream_t = nullptr) { memset(p, value, bytes); return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "cuda_on_cpu"; }

// ---- a synchronous stand-in for the runtime calls capi.cu makes (device memory is host memory, streams run in order
// because every call completes before it returns) -- enough to compile the WHOLE engine library for the CPU test build
#define __constant__
#define __align__(n) __attribute__((aligned(n)))
typedef void* cudaEvent_t;
typedef void* cudaMemPool_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
enum cudaMemPoolAttr { cudaMemPoolAttrReleaseThreshold = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocPortable = 1, cudaHostAllocMapped = 2 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };

namespace cuda_on_cpu {
inline cudaError_t allocate(void** p, size_t bytes) {
	*p = aligned_alloc(256, (bytes + 255) & ~(size_t)255);
	return *p ? cudaSuccess : 2;
}
} // namespace cuda_on_cpu
static inline cudaError_t cudaMalloc(void** p, size_t bytes) { return ::cuda_on_cpu::allocate(p, bytes ? bytes : 1); }
template <typename T> static inline cudaError_t cudaMalloc(T** p, size_t bytes) { return cudaMalloc(reinterpret_cast<void**>(p), bytes); }
static inline cudaError_t cudaMallocAsync(void** p, size_t bytes, cudaStream_t) { return cudaMalloc(p, bytes); }
template <typename T> static inline cudaError_t cudaMallocAsync(T** p, size_t bytes, cudaStream_t s) { return cudaMallocAsync(reinterpret_cast<void**>(p), bytes, s); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned) { return ::cuda_on_cpu::allocate(p, bytes ? bytes : 1); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind) { if (bytes) memmove(dst, src, bytes); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind k, cudaStream_t = nullptr) { return cudaMemcpy(dst, src, bytes, k); }
static inline cudaError_t cudaMemset(void* p, int value, size_t bytes) { memset(p, value, bytes); return cudaSuccess; }
template <typename T> static inline cudaError_t cudaMemcpyToSymbol(T& symbol, const void* src, size_t bytes) { memcpy(&symbol, src, bytes); return cudaSuccess; }
template <typename T> static inline cudaError_t cudaMemcpyFromSymbol(void* dst, const T& symbol, size_t bytes) { memcpy(dst, &symbol, bytes); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = malloc(8); return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { return cudaStreamCreateWithFlags(s, 0); }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = malloc(8); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = ::cuda_on_cpu::deviceCount(); return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int device) {
	if (device < 0 || device >= ::cuda_on_cpu::deviceCount()) return 101;
	::cuda_on_cpu::currentDevice() = device;
	return cudaSuccess;
}
static inline cudaError_t cudaMemcpyPeer(void* dst, int, const void* src, int, size_t bytes) { if (bytes) memmove(dst, src, bytes); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaMemGetInfo(size_t* freeBytes, size_t* totalBytes) { *freeBytes = (size_t)4 << 30; *totalBytes = (size_t)8 << 30; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* pool, int) { *pool = nullptr; return cudaSuccess; }
static inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void*) { return cudaSuccess; }
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
	a->type = cudaMemoryTypeHost; a->device = 0; a->devicePointer = const_cast<void*>(p); a->hostPointer = const_cast<void*>(p);
	return cudaSuccess;
}

// ---- what bvh_build.cu needs on top ---------------------------------------------------------------------------------
struct dim3 { unsigned x, y, z; dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };
static inline cudaError_t cudaLaunchCooperativeKernel(const void*, dim3, dim3, void**, size_t, cudaStream_t) { return 720; } // cooperative launch too large
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline uint32_t atomicMin(uint32_t* p, uint32_t v) { const uint32_t old = *p; if (v < old) *p = v; return old; }
static inline uint32_t atomicMax(uint32_t* p, uint32_t v) { const uint32_t old = *p; if (v > old) *p = v; return old; }
static inline int atomicAdd(int* p, int v) { const int old = *p; *p = old + v; return old; }

// memory-ordering and back-off intrinsics of the streamed path tracer (pathstream.cu): fibers of one CTA share one host
// thread, so a fence has nothing to order; a sleeping warp lets the other fibers run
static inline void __threadfence() {}
static inline void __nanosleep(unsigned) { ::cuda_on_cpu::yield(); }
template <typename T>
static inline T __ldcg(const T* p) { return *p; }
static inline bool __any_sync(unsigned mask, bool predicate) { return __ballot_sync(mask, predicate) != 0; }
