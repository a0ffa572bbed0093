// comm.cu — the exchange steps of the sharded path between the B200s of a node (NVLink 5 / NVSwitch).
// Bulk data (seed records by key range, hits by diagonal, position-ordered keys) travels peer-to-peer: every rank
// owns exchange windows that all peers map through CUDA IPC, an exchange is DMA copies into the peers' windows
// plus a barrier.  NCCL carries the small collectives (histogram / count all-gathers, the one-word barrier, the IPC
// handles) and is the fallback transport where IPC is not available (all-to-all-v from grouped ncclSend/ncclRecv).
// libnccl is bound at run time with dlopen: inside a torch process that is the NCCL torch already loaded, in a
// plain C++ host the system libnccl.so.2.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "mems_b200.h"

namespace mems {

namespace {
struct NcclApi {
	void* lib = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, ncclConfig_t*) = nullptr;  // optional (NCCL >= 2.18)
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi& nccl() {
	static NcclApi api;
	if (api.lib) return api;
	void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // the copy torch (or the host) already loaded
	if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
	if (!h) throw Error(MEMS_ERR_NCCL, std::string("cannot load libnccl.so.2: ") + dlerror());
#define MEMS_NCCL_SYM(name)                                                                    \
	api.name = reinterpret_cast<decltype(api.name)>(dlsym(h, "nccl" #name));                  \
	if (!api.name) throw Error(MEMS_ERR_NCCL, "libnccl lacks nccl" #name);
	MEMS_NCCL_SYM(GetUniqueId)
	MEMS_NCCL_SYM(CommInitRank)
	MEMS_NCCL_SYM(CommDestroy)
	MEMS_NCCL_SYM(GroupStart)
	MEMS_NCCL_SYM(GroupEnd)
	MEMS_NCCL_SYM(Send)
	MEMS_NCCL_SYM(Recv)
	MEMS_NCCL_SYM(AllReduce)
	MEMS_NCCL_SYM(AllGather)
	MEMS_NCCL_SYM(Broadcast)
	MEMS_NCCL_SYM(GetErrorString)
#undef MEMS_NCCL_SYM
	api.CommSplit = reinterpret_cast<decltype(api.CommSplit)>(dlsym(h, "ncclCommSplit"));
	api.lib = h;
	return api;
}

void check(ncclResult_t r, const char* what) {
	if (r != ncclSuccess) throw Error(MEMS_ERR_NCCL, std::string(what) + ": " + nccl().GetErrorString(r));
}
}  // namespace

struct Comm {
	std::shared_ptr<Ctx> ctx;
	ncclComm_t comm = nullptr;
	// second communicator + stream: the all-gather of the position-ordered keys runs beside the seed-range
	// exchange and the local sort instead of in front of them
	ncclComm_t side_comm = nullptr;
	cudaStream_t side_stream = nullptr;
	cudaEvent_t side_ready = nullptr, side_done = nullptr;
	int rank = 0, world = 1;
	// exchange windows: device buffers every peer can write straight into over NVLink (CUDA IPC mappings);
	// the all-to-all steps are plain peer-to-peer copies into them plus one barrier.  Two windows, because a peer
	// may already be sending hits (window 1) while this rank still reads the seed records it received (window 0).
	struct Window {
		void* local = nullptr;
		size_t bytes = 0;
		std::vector<void*> peer;  // peer[p] = rank p's window mapped here (peer[rank] = local)
	};
	Window win[3];  // 0: seed records, 1: hits, 2: position-ordered keys of all sequences
	bool windows_ok = true;     // false once IPC mapping failed on any rank: the NCCL send/recv path is used
	uint32_t* d_barrier = nullptr;
	// one copy stream per peer: the DMA copies of an exchange run side by side over the NVLink ports instead of
	// one peer after the other
	std::vector<cudaStream_t> peer_stream;
	std::vector<cudaEvent_t> peer_done;
	cudaEvent_t fork = nullptr;
	void ensure_peer_streams() {
		if (!peer_stream.empty() || world == 1) return;
		peer_stream.assign(world, nullptr);
		peer_done.assign(world, nullptr);
		for (int p = 0; p < world; ++p) {
			if (p == rank) continue;
			MEMS_CUDA(cudaStreamCreateWithFlags(&peer_stream[p], cudaStreamNonBlocking));
			MEMS_CUDA(cudaEventCreateWithFlags(&peer_done[p], cudaEventDisableTiming));
		}
		MEMS_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
	}
	// CUDA IPC contract: exported memory may be freed only after every importer has closed its mapping.  Releasing a
	// window is therefore two-phase and collective: all ranks close their mappings of the peers' buffers, a barrier
	// proves it, then every rank frees its own buffer.
	void close_peer_mappings(int w) {
		Window& x = win[w];
		for (int p = 0; p < (int)x.peer.size(); ++p)
			if (p != rank && x.peer[p]) cudaIpcCloseMemHandle(x.peer[p]);
		x.peer.clear();
	}
	void free_local(int w) {
		Window& x = win[w];
		if (x.local) cudaFree(x.local);
		x.local = nullptr;
		x.bytes = 0;
	}
	void host_barrier() {  // every rank has reached this point (and its earlier stream work is done)
		if (world == 1 || !comm || !d_barrier) return;
		if (nccl().AllReduce(d_barrier, d_barrier + 8, 1, ncclUint32, ncclSum, comm, ctx->stream) == ncclSuccess)
			cudaStreamSynchronize(ctx->stream);
	}
	~Comm() {
		// collective (mems_comm_destroy): close, barrier, free
		bool any = false;
		for (int w = 0; w < 3; ++w) {
			any = any || win[w].local != nullptr;
			close_peer_mappings(w);
		}
		if (any) host_barrier();
		for (int w = 0; w < 3; ++w) free_local(w);
		if (d_barrier) cudaFree(d_barrier);
		for (cudaStream_t st : peer_stream)
			if (st) cudaStreamDestroy(st);
		for (cudaEvent_t ev : peer_done)
			if (ev) cudaEventDestroy(ev);
		if (fork) cudaEventDestroy(fork);
		if (side_comm) nccl().CommDestroy(side_comm);
		if (comm) nccl().CommDestroy(comm);
		if (side_stream) cudaStreamDestroy(side_stream);
		if (side_ready) cudaEventDestroy(side_ready);
		if (side_done) cudaEventDestroy(side_done);
	}
};

int Comm_rank(const Comm* c) { return c->rank; }
int Comm_world(const Comm* c) { return c->world; }

void comm_unique_id(char* id128) {
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
	ncclUniqueId id;
	check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
	memcpy(id128, &id, 128);
}

Comm* comm_create(std::shared_ptr<Ctx> ctx, const char* id128, int rank, int world) {
	if (world < 1 || rank < 0 || rank >= world) throw Error(MEMS_ERR_INVALID, "bad rank / world size");
	MEMS_CUDA(cudaSetDevice(ctx->device));
	auto* c = new Comm();
	c->ctx = ctx;
	c->rank = rank;
	c->world = world;
	ncclUniqueId id;
	memcpy(&id, id128, 128);
	try {
		check(nccl().CommInitRank(&c->comm, world, id, rank), "ncclCommInitRank");
		if (world > 1 && nccl().CommSplit && !getenv("MEMS_NO_SIDE_COMM")) {
			check(nccl().CommSplit(c->comm, 0, rank, &c->side_comm, nullptr), "ncclCommSplit");
			MEMS_CUDA(cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking));
			MEMS_CUDA(cudaEventCreateWithFlags(&c->side_ready, cudaEventDisableTiming));
			MEMS_CUDA(cudaEventCreateWithFlags(&c->side_done, cudaEventDisableTiming));
		}
	} catch (...) {
		delete c;
		throw;
	}
	return c;
}

void comm_destroy(Comm* c) { delete c; }

// every rank contributes n u64 values; d_recv receives world * n (rank-major)
void comm_all_gather_u64(Comm* c, const uint64_t* d_send, uint64_t* d_recv, size_t n) {
	if (c->world == 1) {
		MEMS_CUDA(cudaMemcpyAsync(d_recv, d_send, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, c->ctx->stream));
		return;
	}
	check(nccl().AllGather(d_send, d_recv, n, ncclUint64, c->comm, c->ctx->stream), "ncclAllGather");
}

// all-to-all with per-peer element counts; buffers are laid out peer-major (offsets = prefix sums of the counts).
// Several arrays that share the counts (keys and values of the same records) go out in ONE NCCL group, i.e. one
// fused send/recv kernel instead of one per array, and the rank's own slice is a plain device copy instead of
// an NCCL send to itself.
void comm_all_to_all_v_multi(Comm* c, int n_arrays, const void* const* d_send, void* const* d_recv, const size_t* elem_bytes,
                             const uint64_t* send_counts, const uint64_t* recv_counts) {
	if (c->world > 1) check(nccl().GroupStart(), "ncclGroupStart");
	for (int a = 0; a < n_arrays; ++a) {
		const char* s = static_cast<const char*>(d_send[a]);
		char* r = static_cast<char*>(d_recv[a]);
		const size_t eb = elem_bytes[a];
		size_t so = 0, ro = 0;
		for (int p = 0; p < c->world; ++p) {
			if (p == c->rank) {
				if (send_counts[p])
					MEMS_CUDA(cudaMemcpyAsync(r + ro, s + so, send_counts[p] * eb, cudaMemcpyDeviceToDevice, c->ctx->stream));
			} else {
				if (send_counts[p]) check(nccl().Send(s + so, send_counts[p] * eb, ncclChar, p, c->comm, c->ctx->stream), "ncclSend");
				if (recv_counts[p]) check(nccl().Recv(r + ro, recv_counts[p] * eb, ncclChar, p, c->comm, c->ctx->stream), "ncclRecv");
			}
			so += send_counts[p] * eb;
			ro += recv_counts[p] * eb;
		}
	}
	if (c->world > 1) check(nccl().GroupEnd(), "ncclGroupEnd");
}

void comm_all_to_all_v(Comm* c, const void* d_send, const uint64_t* send_counts, void* d_recv, const uint64_t* recv_counts,
                       size_t elem_bytes) {
	comm_all_to_all_v_multi(c, 1, &d_send, &d_recv, &elem_bytes, send_counts, recv_counts);
}

// ---- exchange windows ------------------------------------------------------------------------------
// Make window w hold at least `bytes` on EVERY rank (all ranks call this with the same value: it is computed
// from counts every rank already has, so growing needs no negotiation).  Growing is collective: allocate,
// all-gather the IPC handles through NCCL, map the peers' buffers.  Returns false if the windows cannot be used
// (IPC refused on some rank); the caller then falls back to NCCL send/recv.
bool comm_window_reserve(Comm* c, int w, size_t bytes) {
	if (!c->windows_ok) return false;
	Comm::Window& x = c->win[w];
	if (x.local && x.bytes >= bytes) return true;
	MEMS_CUDA(cudaStreamSynchronize(c->ctx->stream));
	if (c->world > 1 && !c->d_barrier) {
		MEMS_CUDA(cudaMalloc(&c->d_barrier, 256));
		MEMS_CUDA(cudaMemset(c->d_barrier, 0, 256));
	}
	if (x.local) {  // growing: every rank takes this branch in the same call (the size is derived identically everywhere)
		c->close_peer_mappings(w);
		c->host_barrier();
		c->free_local(w);
	}
	const size_t want = (bytes + bytes / 4 + (1u << 20)) & ~(size_t)0xfffff;  // head room, 1 MiB granules
	MEMS_CUDA(cudaMalloc(&x.local, want));
	x.bytes = want;
	x.peer.assign(c->world, nullptr);
	x.peer[c->rank] = x.local;
	if (c->world == 1) return true;
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
	struct Slot {
		cudaIpcMemHandle_t h;
		uint64_t ok;
	};
	Slot mine;
	memset(&mine, 0, sizeof mine);
	mine.ok = cudaIpcGetMemHandle(&mine.h, x.local) == cudaSuccess ? 1 : 0;
	if (!mine.ok) cudaGetLastError();
	DevBuf<uint64_t> d_slots(c->ctx.get(), sizeof(Slot) / 8 * (size_t)(c->world + 1));
	MEMS_CUDA(cudaMemcpyAsync(d_slots.p, &mine, sizeof mine, cudaMemcpyHostToDevice, c->ctx->stream));
	check(nccl().AllGather(d_slots.p, d_slots.p + sizeof(Slot) / 8, sizeof(Slot) / 8, ncclUint64, c->comm, c->ctx->stream),
	      "ncclAllGather");
	std::vector<Slot> all(c->world);
	MEMS_CUDA(cudaMemcpyAsync(all.data(), d_slots.p + sizeof(Slot) / 8, sizeof(Slot) * c->world, cudaMemcpyDeviceToHost,
	                          c->ctx->stream));
	MEMS_CUDA(cudaStreamSynchronize(c->ctx->stream));
	uint64_t ok = 1;
	for (int p = 0; p < c->world; ++p) ok &= all[p].ok;
	for (int p = 0; p < c->world && ok; ++p) {
		if (p == c->rank) continue;
		if (cudaIpcOpenMemHandle(&x.peer[p], all[p].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
			cudaGetLastError();
			x.peer[p] = nullptr;
			ok = 0;
		}
	}
	// every rank must take the same path: agree on the outcome
	DevBuf<uint64_t> d_ok(c->ctx.get(), 1);
	MEMS_CUDA(cudaMemcpyAsync(d_ok.p, &ok, 8, cudaMemcpyHostToDevice, c->ctx->stream));
	check(nccl().AllReduce(d_ok.p, d_ok.p, 1, ncclUint64, ncclMin, c->comm, c->ctx->stream), "ncclAllReduce");
	MEMS_CUDA(cudaMemcpyAsync(&ok, d_ok.p, 8, cudaMemcpyDeviceToHost, c->ctx->stream));
	MEMS_CUDA(cudaStreamSynchronize(c->ctx->stream));
	if (!ok) {  // agreed by all ranks: the same two-phase release
		c->close_peer_mappings(w);
		c->host_barrier();
		c->free_local(w);
		c->windows_ok = false;
		return false;
	}
	return true;
}

void* comm_window_local(Comm* c, int w) { return c->win[w].local; }
void* comm_window_peer(Comm* c, int w, int p) { return c->win[w].peer[p]; }

// All ranks' writes into the windows that were queued before this call are complete on every rank once the
// calling stream passes it (a one-word all-reduce: it cannot finish before every rank's stream has reached it).
void comm_window_barrier(Comm* c) {
	if (c->world == 1) return;
	check(nccl().AllReduce(c->d_barrier, c->d_barrier + 8, 1, ncclUint32, ncclSum, c->comm, c->ctx->stream), "ncclAllReduce");
}

// all-to-all-v into window w: array a of every rank lands in the window region that starts at region_off[a]
// (the same layout on every rank: regions are sized for the largest receiver), source q's slice at element
// offset sum_{q' < q} counts[q' * world + p] inside the region of destination p.  counts is the full
// world x world matrix (row = sender), identical on every rank.  Copies are peer-to-peer DMA over NVLink.
void comm_window_all_to_all(Comm* c, int w, int n_arrays, const void* const* d_send, const size_t* elem_bytes,
                            const size_t* region_off, const uint64_t* counts, bool barrier) {
	const int W = c->world, R = c->rank;
	c->ensure_peer_streams();
	if (W > 1) MEMS_CUDA(cudaEventRecord(c->fork, c->ctx->stream));
	for (int i = 0; i < W; ++i) {
		const int p = (R + i) % W;  // own slice first (main stream), then the peers, each on its own stream
		cudaStream_t stream = p == R ? c->ctx->stream : c->peer_stream[p];
		if (p != R) MEMS_CUDA(cudaStreamWaitEvent(stream, c->fork, 0));
		uint64_t before = 0, so_elems = 0;
		for (int q = 0; q < R; ++q) before += counts[(size_t)q * W + p];
		for (int q = 0; q < p; ++q) so_elems += counts[(size_t)R * W + q];
		const uint64_t n = counts[(size_t)R * W + p];
		for (int a = 0; a < n_arrays && n; ++a) {
			const char* src = static_cast<const char*>(d_send[a]) + so_elems * elem_bytes[a];
			char* dst = static_cast<char*>(c->win[w].peer[p]) + region_off[a] + before * elem_bytes[a];
			MEMS_CUDA(cudaMemcpyAsync(dst, src, n * elem_bytes[a], cudaMemcpyDefault, stream));
		}
		if (p != R) MEMS_CUDA(cudaEventRecord(c->peer_done[p], stream));
	}
	for (int p = 0; p < W; ++p)
		if (p != R) MEMS_CUDA(cudaStreamWaitEvent(c->ctx->stream, c->peer_done[p], 0));
	if (barrier) comm_window_barrier(c);
}

// all-gather into window w: this rank's bytes go to byte offset `offset` of EVERY rank's window (its own included),
// queued on the side stream behind whatever the main stream holds so far, so the caller carries on at once.
// comm_all_gather_v_wait() + comm_window_barrier() on the main stream make the gathered data readable.
void comm_window_all_gather(Comm* c, int w, const void* d_send, size_t bytes, size_t offset) {
	cudaStream_t stream = c->side_stream ? c->side_stream : c->ctx->stream;
	if (c->side_stream) {
		MEMS_CUDA(cudaEventRecord(c->side_ready, c->ctx->stream));
		MEMS_CUDA(cudaStreamWaitEvent(c->side_stream, c->side_ready, 0));
	}
	if (bytes)
		for (int i = 0; i < c->world; ++i) {
			const int p = (c->rank + 1 + i) % c->world;  // start with the next rank: the ranks do not all hit rank 0 first
			MEMS_CUDA(cudaMemcpyAsync(static_cast<char*>(c->win[w].peer[p]) + offset, d_send, bytes, cudaMemcpyDefault, stream));
		}
	if (c->side_stream) MEMS_CUDA(cudaEventRecord(c->side_done, c->side_stream));
}

// all-gather with per-rank byte counts: rank p's bytes land at d_recv + offsets[p] on every rank.
// With a side communicator the transfer is queued on its own stream (after everything queued on the main stream so
// far) and the caller continues; comm_all_gather_v_wait() makes the main stream wait for it.
void comm_all_gather_v(Comm* c, const void* d_send, void* d_recv, const uint64_t* byte_counts, const uint64_t* byte_offsets) {
	char* r = static_cast<char*>(d_recv);
	if (c->world == 1) {
		if (byte_counts[0])
			MEMS_CUDA(cudaMemcpyAsync(r + byte_offsets[0], d_send, byte_counts[0], cudaMemcpyDeviceToDevice, c->ctx->stream));
		return;
	}
	ncclComm_t comm = c->side_comm ? c->side_comm : c->comm;
	cudaStream_t stream = c->side_comm ? c->side_stream : c->ctx->stream;
	if (c->side_comm) {
		MEMS_CUDA(cudaEventRecord(c->side_ready, c->ctx->stream));
		MEMS_CUDA(cudaStreamWaitEvent(c->side_stream, c->side_ready, 0));
	}
	// every rank sends its slice to every peer and receives theirs: all NVLink ports busy in both directions
	// (a ring/tree broadcast per rank reached only ~300 GB/s here)
	check(nccl().GroupStart(), "ncclGroupStart");
	for (int p = 0; p < c->world; ++p) {
		if (p == c->rank) continue;
		if (byte_counts[c->rank]) check(nccl().Send(d_send, byte_counts[c->rank], ncclChar, p, comm, stream), "ncclSend");
		if (byte_counts[p]) check(nccl().Recv(r + byte_offsets[p], byte_counts[p], ncclChar, p, comm, stream), "ncclRecv");
	}
	check(nccl().GroupEnd(), "ncclGroupEnd");
	if (byte_counts[c->rank])
		MEMS_CUDA(cudaMemcpyAsync(r + byte_offsets[c->rank], d_send, byte_counts[c->rank], cudaMemcpyDeviceToDevice, stream));
	if (c->side_comm) MEMS_CUDA(cudaEventRecord(c->side_done, c->side_stream));
}

// the main stream continues only after the gathered data has arrived (no-op without a side communicator)
void comm_all_gather_v_wait(Comm* c) {
	if (c->world > 1 && c->side_stream) MEMS_CUDA(cudaStreamWaitEvent(c->ctx->stream, c->side_done, 0));
}

void comm_side_synchronize(Comm* c) {
	if (c->side_stream) cudaStreamSynchronize(c->side_stream);
}

}  // namespace mems
