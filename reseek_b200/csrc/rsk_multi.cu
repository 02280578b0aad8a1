// rsk_multi.cu - multi-GPU plumbing of libreseek_b200: one communicator per context (= per GPU), NCCL over NVLink / NVSwitch.
//
// The search shards by the streamed (-db) side (the loop being sharded is runquery.cpp:82-125): every rank holds a
// contiguous block of DB chains, queries and parameters are replicated, the DP needs no exchange.  The only data-path
// collectives are
//   * the hit gather: every rank's device-compacted hit records + path bytes (sink_kernel.cu) go to the root rank as
//     EXACT-SIZE point-to-point transfers (one grouped ncclSend/ncclRecv round; sizes from a 32-byte all-gather) - no padding, no
//     copy of a rank's hits to ranks that do not need them, no staging through host memory on the sending side;
//   * the `-fast -db` bag merge: the (target, query, score) triples of every block are all-gathered in rank order (= stream
//     order of the unsharded DB) so that every rank replays the same RankedScoresBag stream (rankedscoresbag.cpp:34-51).
// NCCL is bound at run time (dlopen of libnccl.so.2: inside a torch process that is torch's own copy, otherwise the system
// one), so the library still loads on a machine without NCCL; the multi-GPU calls then fail with RSK_ERR_CUDA.
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "rsk_host.cuh"
#include "rsk_multi.cuh"

namespace {

struct NcclApi {
	void *handle = nullptr;
	ncclResult_t (*GetVersion)(int *) = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
	std::string err;
};

NcclApi g_nccl;
std::once_flag g_nccl_once;

void nccl_load()
{
	const char *names[] = {getenv("RSK_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
	for (const char *n : names) {
		if (!n)
			continue;
		g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
		if (g_nccl.handle)
			break;
	}
	if (!g_nccl.handle) {
		g_nccl.err = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?");
		return;
	}
#define SYM(field, name)                                                         \
	do {                                                                         \
		*(void **)(&g_nccl.field) = dlsym(g_nccl.handle, name);                  \
		if (!g_nccl.field && g_nccl.err.empty())                                 \
			g_nccl.err = std::string("libnccl lacks ") + name;                   \
	} while (0)
	SYM(GetVersion, "ncclGetVersion");
	SYM(GetUniqueId, "ncclGetUniqueId");
	SYM(CommInitRank, "ncclCommInitRank");
	SYM(CommInitAll, "ncclCommInitAll");
	SYM(CommDestroy, "ncclCommDestroy");
	SYM(AllGather, "ncclAllGather");
	SYM(Broadcast, "ncclBroadcast");
	SYM(Send, "ncclSend");
	SYM(Recv, "ncclRecv");
	SYM(GroupStart, "ncclGroupStart");
	SYM(GroupEnd, "ncclGroupEnd");
	SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
}

int nccl_ready()
{
	std::call_once(g_nccl_once, nccl_load);
	if (!g_nccl.err.empty())
		return fail(RSK_ERR_CUDA, "NCCL unavailable: %s", g_nccl.err.c_str());
	return RSK_OK;
}

#define NK(call)                                                                                            \
	do {                                                                                                    \
		ncclResult_t r_ = (call);                                                                           \
		if (r_ != ncclSuccess)                                                                              \
			return fail(RSK_ERR_CUDA, "%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(r_), __FILE__, __LINE__); \
	} while (0)

}  // namespace

struct rsk_comm {
	rsk_ctx *ctx = nullptr;
	ncclComm_t comm = nullptr;
	int nranks = 1, rank = 0;
	unsigned long long *d_counts = nullptr;   // [kCountWords * (nranks + 1)]: own words, then everybody's
	unsigned long long *h_counts = nullptr;   // pinned copy
	DevBuf<unsigned char> gbuf[3];            // grow-only receive buffers at the gathering rank(s)
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	rsk_comm_stats stats;
};

static constexpr int kCountWords = 4;

extern "C" int rsk_comm_unique_id(void *id)
{
	if (!id)
		return fail(RSK_ERR_ARG, "rsk_comm_unique_id: null argument");
	int rc = nccl_ready();
	if (rc)
		return rc;
	ncclUniqueId u;
	NK(g_nccl.GetUniqueId(&u));
	static_assert(RSK_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
	memcpy(id, u.internal, RSK_COMM_ID_BYTES);
	return RSK_OK;
}

static int comm_finish_init(rsk_comm *c)
{
	CK(cudaSetDevice(c->ctx->device));
	const size_t words = (size_t)kCountWords * (c->nranks + 1);
	CK(cudaMalloc((void **)&c->d_counts, words * sizeof(unsigned long long)));
	CK(cudaHostAlloc((void **)&c->h_counts, words * sizeof(unsigned long long), cudaHostAllocDefault));
	CK(cudaEventCreate(&c->ev0));
	CK(cudaEventCreate(&c->ev1));
	memset(&c->stats, 0, sizeof(c->stats));
	return RSK_OK;
}

extern "C" int rsk_comm_create(rsk_ctx *ctx, int nranks, int rank, const void *id, rsk_comm **out)
{
	if (!ctx || !out || nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !id))
		return fail(RSK_ERR_ARG, "rsk_comm_create: bad argument (nranks=%d rank=%d)", nranks, rank);
	*out = nullptr;
	rsk_comm *c = new rsk_comm();
	c->ctx = ctx;
	c->nranks = nranks;
	c->rank = rank;
	if (nranks > 1) {
		int rc = nccl_ready();
		if (rc) {
			delete c;
			return rc;
		}
		ncclUniqueId u;
		memcpy(u.internal, id, RSK_COMM_ID_BYTES);
		cudaSetDevice(ctx->device);
		ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, u, rank);
		if (r != ncclSuccess) {
			delete c;
			return fail(RSK_ERR_CUDA, "ncclCommInitRank(rank %d of %d) failed: %s", rank, nranks, g_nccl.GetErrorString(r));
		}
	}
	int rc = comm_finish_init(c);
	if (rc) {
		rsk_comm_destroy(c);
		return rc;
	}
	*out = c;
	return RSK_OK;
}

// One process driving several GPUs (DBSearcher in rsk_host_demo): ncclCommInitAll over the contexts' devices.
extern "C" int rsk_comm_create_all(rsk_ctx *const *ctxs, int n, rsk_comm **out)
{
	if (!ctxs || !out || n < 1)
		return fail(RSK_ERR_ARG, "rsk_comm_create_all: bad argument");
	std::vector<ncclComm_t> comms(n, nullptr);
	if (n > 1) {
		int rc = nccl_ready();
		if (rc)
			return rc;
		std::vector<int> devs(n);
		for (int k = 0; k < n; ++k) {
			if (!ctxs[k])
				return fail(RSK_ERR_ARG, "rsk_comm_create_all: null context %d", k);
			devs[k] = ctxs[k]->device;
			for (int j = 0; j < k; ++j)
				if (devs[j] == devs[k])
					return fail(RSK_ERR_ARG, "rsk_comm_create_all: contexts %d and %d share device %d (one rank per GPU)", j, k, devs[k]);
		}
		NK(g_nccl.CommInitAll(comms.data(), n, devs.data()));
	}
	for (int k = 0; k < n; ++k) {
		rsk_comm *c = new rsk_comm();
		c->ctx = ctxs[k];
		c->nranks = n;
		c->rank = k;
		c->comm = comms[k];
		out[k] = c;
		int rc = comm_finish_init(c);
		if (rc)
			return rc;
	}
	return RSK_OK;
}

extern "C" void rsk_comm_destroy(rsk_comm *c)
{
	if (!c)
		return;
	cudaSetDevice(c->ctx->device);
	if (c->comm && g_nccl.CommDestroy)
		g_nccl.CommDestroy(c->comm);
	if (c->d_counts) cudaFree(c->d_counts);
	if (c->h_counts) cudaFreeHost(c->h_counts);
	for (auto &b : c->gbuf)
		b.release();
	if (c->ev0) cudaEventDestroy(c->ev0);
	if (c->ev1) cudaEventDestroy(c->ev1);
	delete c;
}

extern "C" int rsk_comm_rank(const rsk_comm *c) { return c ? c->rank : 0; }
extern "C" int rsk_comm_nranks(const rsk_comm *c) { return c ? c->nranks : 1; }
extern "C" int rsk_comm_get_stats(const rsk_comm *c, rsk_comm_stats *out)
{
	if (!c || !out)
		return fail(RSK_ERR_ARG, "rsk_comm_get_stats: null argument");
	*out = c->stats;
	return RSK_OK;
}
extern "C" int rsk_comm_reset_stats(rsk_comm *c)
{
	if (!c)
		return fail(RSK_ERR_ARG, "rsk_comm_reset_stats: null argument");
	memset(&c->stats, 0, sizeof(c->stats));
	return RSK_OK;
}

// Contiguous chain ranges with (nearly) equal residue totals: bounds[r] .. bounds[r+1] is rank r's block (balance
// sum L, not chain count).  The same rule as reseek_b200/shard.py::partition_by_residues.
extern "C" int rsk_partition_by_residues(const uint32_t *len, uint32_t n, int nranks, uint32_t *bounds)
{
	if ((n && !len) || nranks < 1 || !bounds)
		return fail(RSK_ERR_ARG, "rsk_partition_by_residues: bad argument");
	uint64_t total = 0;
	for (uint32_t i = 0; i < n; ++i)
		total += len[i];
	bounds[0] = 0;
	uint64_t cum = 0;
	uint32_t k = 0;
	for (int r = 1; r < nranks; ++r) {
		// first k with cum(k) >= total * r / nranks (cum(k) = residues of chains [0, k)); exact integer form of the comparison
		while (k < n && (unsigned __int128)cum * (unsigned)nranks < (unsigned __int128)total * (unsigned)r)
			cum += len[k++];
		bounds[r] = k;
	}
	bounds[nranks] = n;
	return RSK_OK;
}

// ---- internal: collectives used by the search drivers (declared in rsk_multi.cuh) ----
namespace rsk {

// every rank contributes kCountWords words; afterwards all[r*kCountWords + w] is rank r's word w (host memory, valid until the next call)
int comm_exchange_counts(rsk_comm *c, const unsigned long long *mine, const unsigned long long **all)
{
	cudaStream_t st = c->ctx->stream;
	unsigned long long *h = c->h_counts;
	for (int w = 0; w < kCountWords; ++w)
		h[w] = mine[w];
	if (c->nranks == 1) {
		for (int w = 0; w < kCountWords; ++w)
			h[kCountWords + w] = mine[w];
		*all = h + kCountWords;
		return RSK_OK;
	}
	CK(cudaMemcpyAsync(c->d_counts, h, sizeof(unsigned long long) * kCountWords, cudaMemcpyHostToDevice, st));
	NK(g_nccl.AllGather(c->d_counts, c->d_counts + kCountWords, kCountWords, ncclUint64, c->comm, st));
	CK(cudaMemcpyAsync(h + kCountWords, c->d_counts + kCountWords, sizeof(unsigned long long) * kCountWords * c->nranks,
			cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	c->stats.collectives += 1;
	c->stats.bytes_sent += sizeof(unsigned long long) * kCountWords * (c->nranks - 1);
	c->stats.bytes_recv += sizeof(unsigned long long) * kCountWords * (c->nranks - 1);
	*all = h + kCountWords;
	return RSK_OK;
}

// Gather `nparts` device arrays of every rank on `root` (part p of rank r = bytes[r][p] bytes; the caller got the sizes from
// comm_exchange_counts).  On the root, out[p] points at a device buffer holding part p of rank 0, rank 1, ... back to back.
// Exact-size transfers: one grouped round of ncclSend / ncclRecv.
int comm_gather_parts(rsk_comm *c, int root, int nparts, const void *const *src, const unsigned long long *bytes /* [nranks][nparts] */,
		unsigned char **out, unsigned long long *out_total)
{
	if (nparts > 3)
		return fail(RSK_ERR_ARG, "comm_gather_parts: at most 3 parts");
	CK(cudaSetDevice(c->ctx->device));
	cudaStream_t st = c->ctx->stream;
	const int N = c->nranks, me = c->rank;
	CK(cudaEventRecord(c->ev0, st));
	if (me == root) {
		for (int p = 0; p < nparts; ++p) {
			unsigned long long tot = 0;
			for (int r = 0; r < N; ++r)
				tot += bytes[(size_t)r * nparts + p];
			if (c->gbuf[p].ensure((size_t)tot + 16)) {
				cudaGetLastError();
				return fail(RSK_ERR_NOMEM, "gather buffer of %llu bytes on the root rank", tot);
			}
			out[p] = c->gbuf[p].p;
			out_total[p] = tot;
		}
	}
	unsigned long long moved = 0;
	if (N > 1)
		NK(g_nccl.GroupStart());
	for (int p = 0; p < nparts; ++p) {
		if (me == root) {
			unsigned long long off = 0;
			for (int r = 0; r < N; ++r) {
				const unsigned long long nb = bytes[(size_t)r * nparts + p];
				if (nb) {
					if (r == me)
						CK(cudaMemcpyAsync(c->gbuf[p].p + off, src[p], nb, cudaMemcpyDeviceToDevice, st));
					else {
						NK(g_nccl.Recv(c->gbuf[p].p + off, nb, ncclUint8, r, c->comm, st));
						moved += nb;
					}
				}
				off += nb;
			}
		} else {
			const unsigned long long nb = bytes[(size_t)me * nparts + p];
			if (nb) {
				NK(g_nccl.Send(src[p], nb, ncclUint8, root, c->comm, st));
				moved += nb;
			}
		}
	}
	if (N > 1)
		NK(g_nccl.GroupEnd());
	CK(cudaEventRecord(c->ev1, st));
	CK(cudaEventSynchronize(c->ev1));
	float ms = 0;
	CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
	c->stats.collective_ms += ms;
	c->stats.collectives += 1;
	if (me == root)
		c->stats.bytes_recv += moved;
	else
		c->stats.bytes_sent += moved;
	return RSK_OK;
}

// All-gather of one variable-size device array per rank, in rank order, into a device buffer on EVERY rank (broadcast rounds
// of exact size: every rank needs every block - the bag merge of `-fast -db`).
int comm_allgather_parts(rsk_comm *c, int nparts, const void *const *src, const unsigned long long *bytes /* [nranks][nparts] */,
		unsigned char **out, unsigned long long *out_total)
{
	if (nparts > 3)
		return fail(RSK_ERR_ARG, "comm_allgather_parts: at most 3 parts");
	CK(cudaSetDevice(c->ctx->device));
	cudaStream_t st = c->ctx->stream;
	const int N = c->nranks, me = c->rank;
	CK(cudaEventRecord(c->ev0, st));
	for (int p = 0; p < nparts; ++p) {
		unsigned long long tot = 0;
		for (int r = 0; r < N; ++r)
			tot += bytes[(size_t)r * nparts + p];
		if (c->gbuf[p].ensure((size_t)tot + 16)) {
			cudaGetLastError();
			return fail(RSK_ERR_NOMEM, "all-gather buffer of %llu bytes", tot);
		}
		out[p] = c->gbuf[p].p;
		out_total[p] = tot;
	}
	unsigned long long sent = 0, recv = 0;
	if (N > 1)
		NK(g_nccl.GroupStart());
	for (int p = 0; p < nparts; ++p) {
		unsigned long long off = 0;
		for (int r = 0; r < N; ++r) {
			const unsigned long long nb = bytes[(size_t)r * nparts + p];
			if (nb) {
				if (N == 1 || r == me) {
					CK(cudaMemcpyAsync(c->gbuf[p].p + off, src[p], nb, cudaMemcpyDeviceToDevice, st));
					if (N > 1)
						for (int d = 0; d < N; ++d)
							if (d != me) {
								NK(g_nccl.Send(src[p], nb, ncclUint8, d, c->comm, st));
								sent += nb;
							}
				} else {
					NK(g_nccl.Recv(c->gbuf[p].p + off, nb, ncclUint8, r, c->comm, st));
					recv += nb;
				}
			}
			off += nb;
		}
	}
	if (N > 1)
		NK(g_nccl.GroupEnd());
	CK(cudaEventRecord(c->ev1, st));
	CK(cudaEventSynchronize(c->ev1));
	float ms = 0;
	CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
	c->stats.collective_ms += ms;
	c->stats.collectives += 1;
	c->stats.bytes_sent += sent;
	c->stats.bytes_recv += recv;
	return RSK_OK;
}

int comm_rank(const rsk_comm *c) { return c ? c->rank : 0; }
int comm_nranks(const rsk_comm *c) { return c ? c->nranks : 1; }
rsk_ctx *comm_ctx(const rsk_comm *c) { return c ? c->ctx : nullptr; }

}  // namespace rsk
