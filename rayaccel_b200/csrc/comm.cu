// comm.cu -- the one collective on the path: the per-frame hit reduction (north_star: "NCCL over NVLink only for the
// per-frame hit reduction"). Rays shard by index and the scene is replicated, so GPUs exchange no
// ray, node or result; what a frame needs from all of them is the sum of their counters -- Stats.raysTraced and the
// hit count (RayAccelerator.cpp:200,372,755-758 keep that sum on the single host the reference runs on).
//
// Every traversal launch adds {rays, hits} to its device's frame record (DeviceState::dFrame, one atomic per warp at
// kernel exit). racc_cuda_frame_reduce sums the records with ncclAllReduce
//   * over the devices of the calling thread's device set, one communicator per device in ONE process
//     (ncclCommInitAll) -- the library drives several B200s itself, e.g. behind racc::render(); or
//   * over the ranks of a multi-process job, one device per process (racc_cuda_comm_init_rank: the host application
//     broadcasts the unique id however it likes -- bench.py uses its torch.distributed store). The rank communicator
//     is used by callers whose device set is that one device; a thread that drives a multi-device set reduces over its
//     set only (a rank-wide collective needs every rank to call, which only the per-rank frame loop guarantees);
// and zeroes them for the next frame. NCCL is bound at run time (dlopen "libnccl.so.2": in a process that already
// loaded one -- PyTorch brings its own -- that very library is used, so there is a single NCCL in the address space);
// a single device without a rank communicator needs no NCCL at all.
#include "capi_internal.h"

#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace racc_b200 {
namespace {

// the handful of NCCL declarations used (nccl.h of NCCL 2.x; ABI-stable since 2.0)
typedef void* ncclComm_t;
typedef int ncclResult_t; // ncclSuccess = 0
struct ncclUniqueId { char internal[128]; };
constexpr int kNcclUint64 = 5; // ncclDataType_t: ncclInt8 0, ncclUint8 1, ncclInt32 2, ncclUint32 3, ncclInt64 4, ncclUint64 5
constexpr int kNcclSum = 0;    // ncclRedOp_t

struct Nccl {
	void* lib = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

std::mutex g_commMutex;
Nccl g_nccl;
// single process, several devices: one communicator per device of the set the communicators were made for
std::vector<int> g_setDevices;
std::vector<ncclComm_t> g_setComms;
// several processes, one device each
ncclComm_t g_rankComm = nullptr;
int g_rankDevice = -1, g_rank = 0, g_ranks = 1;

// caller holds g_commMutex
int loadNccl() {
	if (g_nccl.lib) return 0;
	const char* names[] = {getenv("RACC_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
	void* lib = nullptr;
	for (const char* name : names)
		if (name && *name && (lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL))) break;
	if (!lib) return fail("NCCL is needed for a frame reduction over more than one GPU and libnccl.so.2 cannot be loaded (%s)", dlerror());
	Nccl n;
	n.lib = lib;
#define BIND(field, symbol)                                                            \
	*reinterpret_cast<void**>(&n.field) = dlsym(lib, symbol);                           \
	if (!n.field) return fail("libnccl: symbol %s not found", symbol);
	BIND(GetUniqueId, "ncclGetUniqueId")
	BIND(CommInitRank, "ncclCommInitRank")
	BIND(CommInitAll, "ncclCommInitAll")
	BIND(CommDestroy, "ncclCommDestroy")
	BIND(AllReduce, "ncclAllReduce")
	BIND(AllGather, "ncclAllGather")
	BIND(GroupStart, "ncclGroupStart")
	BIND(GroupEnd, "ncclGroupEnd")
	BIND(GetErrorString, "ncclGetErrorString")
#undef BIND
	g_nccl = n;
	return 0;
}

#define RACC_NCCL_CHECK(call)                                                                                   \
	do {                                                                                                        \
		ncclResult_t r_ = (call);                                                                               \
		if (r_ != 0) return fail("%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(r_), __FILE__, __LINE__); \
	} while (0)

// caller holds g_commMutex
int ensureSetComms(const std::vector<int>& set) {
	if (g_setDevices == set) return 0;
	for (ncclComm_t c : g_setComms) g_nccl.CommDestroy(c);
	g_setComms.assign(set.size(), nullptr);
	g_setDevices.clear();
	RACC_NCCL_CHECK(g_nccl.CommInitAll(g_setComms.data(), (int)set.size(), set.data()));
	g_setDevices = set;
	return 0;
}

} // namespace

// totals (host, may be null): receives the sums and makes the call wait for them; null leaves them in the bound device's
// DeviceState::dFrameTotal, ordered on `stream`. Every launch to be counted must have completed, or have been enqueued on
// `stream`, before the call (launches on the set's other devices run on other streams: they must have completed).
int commFrameReduce(racc_cuda_counters* totals, cudaStream_t stream) {
	DeviceState* dev = currentDevice();
	if (!dev) return -1;
	const std::vector<int> set = currentDeviceSet();
	std::lock_guard<std::mutex> lock(g_commMutex);
	const bool ranks = g_rankComm != nullptr && set.size() == 1 && g_rankDevice == dev->ordinal;
	if (set.size() > 1) {
		if (loadNccl() || ensureSetComms(set)) return -1;
		// bound device: the reduction follows the caller's stream; the others reduce on their own streams
		RACC_CUDA_CHECK(cudaEventRecord(dev->reduceReady, stream));
		RACC_CUDA_CHECK(cudaStreamWaitEvent(dev->reduceStream, dev->reduceReady, 0));
		RACC_NCCL_CHECK(g_nccl.GroupStart());
		for (size_t k = 0; k < set.size(); ++k) {
			DeviceState* d = useDevice(set[k]);
			if (!d) { g_nccl.GroupEnd(); return -1; }
			RACC_NCCL_CHECK(g_nccl.AllReduce(d->dFrame, d->dFrameTotal, 8, kNcclUint64, kNcclSum, g_setComms[k], d->reduceStream));
		}
		RACC_NCCL_CHECK(g_nccl.GroupEnd());
		for (size_t k = 0; k < set.size(); ++k) {
			DeviceState* d = useDevice(set[k]);
			if (!d) return -1;
			RACC_CUDA_CHECK(cudaMemsetAsync(d->dFrame, 0, 8 * sizeof(unsigned long long), d->reduceStream));
			if (k) RACC_CUDA_CHECK(cudaStreamSynchronize(d->reduceStream)); // later launches there start from a zeroed record
		}
		RACC_CUDA_CHECK(cudaSetDevice(dev->ordinal));
		RACC_CUDA_CHECK(cudaEventRecord(dev->reduceDone, dev->reduceStream));
		RACC_CUDA_CHECK(cudaStreamWaitEvent(stream, dev->reduceDone, 0));
	}
	else {
		RACC_CUDA_CHECK(cudaMemcpyAsync(dev->dFrameTotal, dev->dFrame, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, stream));
		RACC_CUDA_CHECK(cudaMemsetAsync(dev->dFrame, 0, 8 * sizeof(unsigned long long), stream));
	}
	if (ranks) // this process' record, summed over the processes
		RACC_NCCL_CHECK(g_nccl.AllReduce(dev->dFrameTotal, dev->dFrameTotal, 8, kNcclUint64, kNcclSum, g_rankComm, stream));
	if (totals) {
		static_assert(sizeof(racc_cuda_counters) == 8 * sizeof(unsigned long long), "counter record is 8 x u64");
		RACC_CUDA_CHECK(cudaMemcpyAsync(totals, dev->dFrameTotal, sizeof(*totals), cudaMemcpyDeviceToHost, stream));
		RACC_CUDA_CHECK(cudaStreamSynchronize(stream));
	}
	return 0;
}

// Result slices of a ray-sharded frame gathered so that every rank holds the full, index-parallel hit buffer (SURVEY.md
// section 8e): rank r's `bytes_per_rank` bytes at recv + r * bytes_per_rank. Ranks of the multi-process communicator only (a
// single process already sees every device's results in its own address space).
int commAllGather(const void* send, void* recv, size_t bytesPerRank, cudaStream_t stream) {
	DeviceState* dev = currentDevice();
	if (!dev) return -1;
	std::lock_guard<std::mutex> lock(g_commMutex);
	if (!g_rankComm) {
		// one rank: the gather is a copy
		if (send != recv) RACC_CUDA_CHECK(cudaMemcpyAsync(recv, send, bytesPerRank, cudaMemcpyDeviceToDevice, stream));
		return 0;
	}
	if (g_rankDevice != dev->ordinal)
		return fail("racc_cuda_gather_results: the rank communicator belongs to CUDA device %d, the calling thread is bound to %d", g_rankDevice, dev->ordinal);
	RACC_NCCL_CHECK(g_nccl.AllGather(send, recv, bytesPerRank, /*ncclUint8*/ 1, g_rankComm, stream));
	return 0;
}

int commRanks(int* rank) {
	std::lock_guard<std::mutex> lock(g_commMutex);
	if (rank) *rank = g_rank;
	return g_ranks;
}

void commShutdown() {
	std::lock_guard<std::mutex> lock(g_commMutex);
	if (!g_nccl.lib) return;
	for (ncclComm_t c : g_setComms) g_nccl.CommDestroy(c);
	g_setComms.clear();
	g_setDevices.clear();
	if (g_rankComm) g_nccl.CommDestroy(g_rankComm);
	g_rankComm = nullptr;
	g_rankDevice = -1;
	g_rank = 0;
	g_ranks = 1;
}

} // namespace racc_b200

using namespace racc_b200;

extern "C" {

int racc_cuda_comm_unique_id(void* id128) {
	if (!id128) return fail("racc_cuda_comm_unique_id: null argument");
	std::lock_guard<std::mutex> lock(g_commMutex);
	if (loadNccl()) return -1;
	ncclUniqueId id;
	RACC_NCCL_CHECK(g_nccl.GetUniqueId(&id));
	memcpy(id128, id.internal, sizeof(id.internal));
	return 0;
}

int racc_cuda_comm_init_rank(const void* id128, int rank, int nranks) {
	if (!id128 || rank < 0 || nranks < 1 || rank >= nranks) return fail("racc_cuda_comm_init_rank: bad arguments");
	DeviceState* dev = currentDevice();
	if (!dev) return -1;
	std::lock_guard<std::mutex> lock(g_commMutex);
	if (g_rankComm) return fail("racc_cuda_comm_init_rank: this process already has a rank communicator (racc_cuda_comm_destroy first)");
	if (loadNccl()) return -1;
	ncclUniqueId id;
	memcpy(id.internal, id128, sizeof(id.internal));
	RACC_NCCL_CHECK(g_nccl.CommInitRank(&g_rankComm, nranks, id, rank));
	g_rankDevice = dev->ordinal;
	g_rank = rank;
	g_ranks = nranks;
	return 0;
}

void racc_cuda_comm_destroy(void) { commShutdown(); }

int racc_cuda_comm_ranks(int* rank) { return commRanks(rank); }

int racc_cuda_gather_results(const void* device_results, uint32_t rays_per_rank, void* device_all_results, void* cuda_stream) {
	if (!device_results || !device_all_results) return fail("racc_cuda_gather_results: null argument");
	return commAllGather(device_results, device_all_results, (size_t)rays_per_rank * 16, static_cast<cudaStream_t>(cuda_stream));
}

} // extern "C"
