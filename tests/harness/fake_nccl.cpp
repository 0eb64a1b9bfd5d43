// fake_nccl.cpp -- TEST INFRASTRUCTURE ONLY. A stand-in for libnccl.so.2 over host memory, for the CPU test build of the
// engine library (tests/test_library_on_cpu.py), where "device memory" is host memory and every stream call completes
// before it returns. It implements exactly the entry points rayaccel_b200/csrc/comm.cu binds with dlsym: communicators
// of ONE process (ncclCommInitAll, or ncclCommInitRank with nranks == 1) and a grouped sum all-reduce of 64-bit counters.
// The engine loads it only because the test names it in RACC_B200_NCCL_LIB; the product never ships or links it.
#include <stdint.h>
#include <string.h>

#include <vector>

namespace {
struct Comm { int clique; int rank; int size; };
struct Posted { const void* send; void* recv; size_t count; Comm* comm; };
int g_cliques = 0;
int g_groupDepth = 0;
std::vector<Posted> g_posted;
int g_allreduces = 0;

int flush() {
	// every clique must have posted one call per member
	while (!g_posted.empty()) {
		const int clique = g_posted[0].comm->clique, size = g_posted[0].comm->size;
		std::vector<Posted> mine;
		for (size_t k = 0; k < g_posted.size();)
			if (g_posted[k].comm->clique == clique) { mine.push_back(g_posted[k]); g_posted.erase(g_posted.begin() + (long)k); }
			else ++k;
		if ((int)mine.size() != size) return 5; // ncclInvalidUsage: a member did not take part
		std::vector<uint64_t> sum(mine[0].count, 0);
		for (const Posted& p : mine)
			for (size_t i = 0; i < p.count; ++i) sum[i] += static_cast<const uint64_t*>(p.send)[i];
		for (const Posted& p : mine) memcpy(p.recv, sum.data(), p.count * sizeof(uint64_t));
		++g_allreduces;
	}
	return 0;
}
} // namespace

extern "C" {
struct ncclUniqueId { char internal[128]; };
int fake_nccl_allreduces(void) { return g_allreduces; }
int ncclGetUniqueId(ncclUniqueId* id) { memset(id, 7, sizeof(*id)); return 0; }
int ncclCommInitRank(void** comm, int nranks, ncclUniqueId, int rank) {
	if (nranks != 1 || rank != 0) return 5;
	*comm = new Comm{g_cliques++, 0, 1};
	return 0;
}
int ncclCommInitAll(void** comms, int n, const int*) {
	const int clique = g_cliques++;
	for (int k = 0; k < n; ++k) comms[k] = new Comm{clique, k, n};
	return 0;
}
int ncclCommDestroy(void* comm) { delete static_cast<Comm*>(comm); return 0; }
int ncclGroupStart(void) { ++g_groupDepth; return 0; }
int ncclGroupEnd(void) { return --g_groupDepth == 0 ? flush() : 0; }
int ncclAllReduce(const void* send, void* recv, size_t count, int datatype, int op, void* comm, void*) {
	if (datatype != 5 || op != 0) return 4; // ncclUint64, ncclSum only
	g_posted.push_back(Posted{send, recv, count, static_cast<Comm*>(comm)});
	return g_groupDepth ? 0 : flush();
}
int ncclAllGather(const void* send, void* recv, size_t count, int datatype, void* comm, void*) {
	if (datatype != 1 || static_cast<Comm*>(comm)->size != 1) return 4; // ncclUint8, one-rank communicators only
	if (send != recv) memmove(recv, send, count);
	return 0;
}
const char* ncclGetErrorString(int) { return "fake nccl error"; }
}
