// Micro-benchmark: L1 data-pipe wavefronts of divergent read-only loads by width and sharing pattern.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1wf l1_wavefronts.cu ; run under
// ncu --metrics l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_tex_wavefronts.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,gpu__time_duration.sum
// Modes 8-10 (added after round 1's last GPU call, not yet run): the same 64-byte node fetched through the TEXTURE path
// (tex1Dfetch, 16 bytes per lane per instruction), alone and mixed with LDG.256 -- the question for round 2 is whether
// texture fetches draw on a wavefront budget of their own (then the pair fetches of the traversal kernel could move there)
// or on the same data pipe as LSU loads (then nothing is gained).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// Modes 13-15 (same status): WHICH lanes may share a sector for free? tests/harness/l1_model.py (the traversal kernel's own source
// run on the CPU with its gathers traced) reproduces the measured bounce-stream throughput only if sharing is exploited
// among very few neighbouring lanes. 13: lanes with equal (lane & 7) share (8 nodes per warp, every quarter-warp all
// distinct): 8.5 wavefronts per instruction if the whole warp coalesces, 34 if only a quarter-warp does. 14: lanes l and
// l^4 share (16 nodes per warp, 4 per quarter-warp, every aligned group of 4 all distinct): 17 if groups of 8 or more
// coalesce, 34 if only groups of 4. 15: lanes l and l^2 share: 17 if groups of 4 coalesce, 34 if only pairs.
// Mode 11 (same status): the node read from __constant__ memory with a per-lane index -- the constant path serialises
// distinct addresses, but it is a pipe of its own; if it sustains a useful rate beside LDG traffic, the top of the tree
// (512 nodes = 32 KB take roughly half of all visits) could be served from it.
__constant__ float4 cnodes[2048];

// Each lane reads `BYTES` bytes at idx[lane-th]*64 (a 64-byte "node"); pattern decides how many lanes share a node.
template <int MODE>
__global__ void k(const float4* __restrict__ data, const uint32_t* __restrict__ idx, int iters, float* out, cudaTextureObject_t tex) {
	const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
	float acc = 0.f;
	uint32_t n = idx[gid];
	for (int it = 0; it < iters; ++it) {
		const float4* p = data + 4 * (size_t)n;
		if (MODE == 0) { // 4 x LDG.128
			float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
			acc += a.x + b.y + c.z + d.w;
			n = (__float_as_uint(d.w) + n * 1664525u + 1013904223u) & 0xffffu;
		}
		else if (MODE == 1) { // 2 x LDG.256
			float v[16];
			asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
			asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8+32];" : "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]) : "l"(p));
			acc += v[0] + v[5] + v[10] + v[15];
			n = (__float_as_uint(v[15]) + n * 1664525u + 1013904223u) & 0xffffu;
		}
		else if (MODE == 2) { // 1 x LDG.256 (first 32 bytes only)
			float v[8];
			asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
			acc += v[0] + v[5];
			n = (__float_as_uint(v[7]) + n * 1664525u + 1013904223u) & 0xffffu;
		}
		else if (MODE == 3) { // 1 x LDG.128
			float4 a = __ldg(p);
			acc += a.x;
			n = (__float_as_uint(a.w) + n * 1664525u + 1013904223u) & 0xffffu;
		}
		else if (MODE == 4) { // 1 x LDG.32
			float a = __ldg(reinterpret_cast<const float*>(p));
			acc += a;
			n = (__float_as_uint(a) + n * 1664525u + 1013904223u) & 0xffffu;
		}
		else if (MODE == 5) { // 2 x LDG.256, all lanes of a warp on the SAME node (broadcast)
			const float4* q = data + 4 * (size_t)__shfl_sync(0xffffffffu, n, 0);
			float v[16];
			asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(q));
			asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8+32];" : "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]) : "l"(q));
			acc += v[0] + v[5] + v[10] + v[15];
			n = (__float_as_uint(v[15]) + n * 1664525u + 1013904223u) & 0xffffu;
		}
		else if (MODE == 6) { // 2 x LDG.256, lanes in groups of 4 share a node
			const float4* q = data + 4 * (size_t)__shfl_sync(0xffffffffu, n, threadIdx.x & 28);
			float v[16];
			asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(q));
			asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8+32];" : "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]) : "l"(q));
			acc += v[0] + v[5] + v[10] + v[15];
			n = (__float_as_uint(v[15]) + n * 1664525u + 1013904223u) & 0xffffu;
		}
		else if (MODE == 8) { // 4 x TEX (tex1Dfetch float4), every lane its own node
			const int e = 4 * (int)n;
			float4 a = tex1Dfetch<float4>(tex, e), b = tex1Dfetch<float4>(tex, e + 1), c = tex1Dfetch<float4>(tex, e + 2), d = tex1Dfetch<float4>(tex, e + 3);
			acc += a.x + b.y + c.z + d.w;
			n = (__float_as_uint(d.w) + n * 1664525u + 1013904223u) & 0xffffu;
		}
		else if (MODE == 9) { // 1 x LDG.256 (first half) + 2 x TEX (second half): a node split over both paths
			float v[8];
			asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
			const int e = 4 * (int)n;
			float4 c = tex1Dfetch<float4>(tex, e + 2), d = tex1Dfetch<float4>(tex, e + 3);
			acc += v[0] + v[5] + c.z + d.w;
			n = (__float_as_uint(d.w) + n * 1664525u + 1013904223u) & 0xffffu;
		}
		else if (MODE == 10) { // 1 x TEX (16 bytes)
			float4 a = tex1Dfetch<float4>(tex, 4 * (int)n);
			acc += a.x;
			n = (__float_as_uint(a.w) + n * 1664525u + 1013904223u) & 0xffffu;
		}
		else if (MODE == 11) { // 4 x LDC.128 with a divergent index (512 nodes in constant memory)
			const int e = 4 * (int)(n & 511u);
			float4 a = cnodes[e], b = cnodes[e + 1], c = cnodes[e + 2], d = cnodes[e + 3];
			acc += (a.x + a.y + a.z + a.w) + (b.x + b.y + b.z + b.w) + (c.x + c.y + c.z + c.w) + (d.x + d.y + d.z + d.w); // all 64 bytes are needed
			n = (__float_as_uint(d.w) + n * 1664525u + 1013904223u) & 0xffffu;
		}
		else if (MODE == 12) { // every other visit from constant memory, the others 2 x LDG.256: do the two paths overlap?
			if (it & 1) {
				const int e = 4 * (int)(n & 511u);
				float4 a = cnodes[e], b = cnodes[e + 1], c = cnodes[e + 2], d = cnodes[e + 3];
				acc += (a.x + a.y + a.z + a.w) + (b.x + b.y + b.z + b.w) + (c.x + c.y + c.z + c.w) + (d.x + d.y + d.z + d.w); // all 64 bytes are needed
				n = (__float_as_uint(d.w) + n * 1664525u + 1013904223u) & 0xffffu;
			}
			else {
				float v[16];
				asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
				asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8+32];" : "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]) : "l"(p));
				acc += v[0] + v[5] + v[10] + v[15];
				n = (__float_as_uint(v[15]) + n * 1664525u + 1013904223u) & 0xffffu;
			}
		}
		else if (MODE == 13 || MODE == 14 || MODE == 15) { // 2 x LDG.256, sharing lanes far apart / 4 apart / 2 apart
			const int src = MODE == 13 ? (threadIdx.x & 7) : MODE == 14 ? (threadIdx.x & 27) : (threadIdx.x & 29);
			const float4* q = data + 4 * (size_t)__shfl_sync(0xffffffffu, n, src);
			float v[16];
			asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(q));
			asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8+32];" : "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]) : "l"(q));
			acc += v[0] + v[5] + v[10] + v[15];
			n = (__float_as_uint(v[15]) + n * 1664525u + 1013904223u) & 0xffffu;
		}
		else if (MODE == 7) { // 1 x LDG.64
			float2 a = __ldg(reinterpret_cast<const float2*>(p));
			acc += a.x;
			n = (__float_as_uint(a.y) + n * 1664525u + 1013904223u) & 0xffffu;
		}
	}
	out[gid] = acc;
}

int main() {
	const int nodes = 65536, threads = 148 * 5 * 256, iters = 256; // 4 MB of nodes: L2-resident, not L1-resident
	float4* data; uint32_t* idx; float* out;
	cudaMalloc(&data, (size_t)nodes * 64); cudaMalloc(&idx, threads * 4); cudaMalloc(&out, threads * 4);
	cudaMemset(data, 0, (size_t)nodes * 64);
	uint32_t* h = new uint32_t[threads];
	uint32_t s = 1; for (int i = 0; i < threads; ++i) { s = s * 1664525u + 1013904223u; h[i] = (s >> 8) & 0xffffu; }
	cudaMemcpy(idx, h, threads * 4, cudaMemcpyHostToDevice);
	cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = data;
	rd.res.linear.desc = cudaCreateChannelDesc<float4>(); rd.res.linear.sizeInBytes = (size_t)nodes * 64;
	cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
	cudaTextureObject_t tex = 0;
	if (cudaCreateTextureObject(&tex, &rd, &td, nullptr) != cudaSuccess) { printf("texture object failed\n"); return 1; }
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
#define RUN(M) { k<M><<<threads / 256, 256>>>(data, idx, iters, out, tex); cudaEventRecord(a); k<M><<<threads / 256, 256>>>(data, idx, iters, out, tex); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); \
	printf("mode %d: %.3f ms, %.2f G lane-loads(of a node)/s\n", M, ms, (double)threads * iters / ms / 1e6); }
	RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13) RUN(14) RUN(15)
	if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel error\n"); return 1; }
	return 0;
}
