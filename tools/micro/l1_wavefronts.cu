// Micro-benchmark: L1 data-pipe wavefronts of divergent read-only loads by width and sharing pattern.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1wf l1_wavefronts.cu ; run under
// ncu --metrics l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,gpu__time_duration.sum
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// Each lane reads `BYTES` bytes at idx[lane-th]*64 (a 64-byte "node"); pattern decides how many lanes share a node.
template <int MODE>
__global__ void k(const float4* __restrict__ data, const uint32_t* __restrict__ idx, int iters, float* out) {
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
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
#define RUN(M) { k<M><<<threads / 256, 256>>>(data, idx, iters, out); cudaEventRecord(a); k<M><<<threads / 256, 256>>>(data, idx, iters, out); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); \
	printf("mode %d: %.3f ms, %.2f G lane-loads(of a node)/s\n", M, ms, (double)threads * iters / ms / 1e6); }
	RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7)
	return 0;
}
