// rsk_host.cuh - host-side objects shared by the API translation units (context, chain sets, grow-only buffers).
#pragma once

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <utility>
#include <vector>

#include "rsk_internal.cuh"

int rsk_fail(int code, const char *fmt, ...);
struct rsk_ctx;
int rsk_h2d(rsk_ctx *ctx, void *dst, const void *src, size_t bytes, bool *direct);
#define fail rsk_fail

#define CK(call)                                                                                     \
	do {                                                                                             \
		cudaError_t e_ = (call);                                                                     \
		if (e_ != cudaSuccess)                                                                       \
			return fail(RSK_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while (0)

using namespace rsk;

// ------------------------------------------------------------------------------------------------
// objects
// ------------------------------------------------------------------------------------------------
struct rsk_chainset {
	rsk_ctx *ctx = nullptr;  // identity only (never dereferenced by rsk_chainset_free: the context may be gone)
	int device = 0;
	DevChains d;               // pointers into `slab`
	void *slab = nullptr;      // the set's device memory (one allocation, recycled through the context)
	size_t slab_bytes = 0;
	std::vector<uint32_t> hlen;
	std::vector<uint64_t> hoff;
	uint32_t maxlen = 0;
	bool has_mu = false;
};

template <typename T>
struct DevBuf {
	T *p = nullptr;
	size_t cap = 0;  // elements
	int ensure(size_t n)
	{
		if (n <= cap)
			return 0;
		if (p)
			cudaFree(p);
		p = nullptr;
		cap = 0;
		size_t want = n + n / 8 + 16;
		if (cudaMalloc((void **)&p, want * sizeof(T)) != cudaSuccess) {
			cudaGetLastError();
			if (cudaMalloc((void **)&p, n * sizeof(T)) != cudaSuccess)
				return -1;
			want = n;
		}
		cap = want;
		return 0;
	}
	void release()
	{
		if (p)
			cudaFree(p);
		p = nullptr;
		cap = 0;
	}
};

template <typename T>
struct PinBuf {
	T *p = nullptr;
	size_t cap = 0;
	int ensure(size_t n)
	{
		if (n <= cap)
			return 0;
		if (p)
			cudaFreeHost(p);
		p = nullptr;
		cap = 0;
		size_t want = n + n / 8 + 16;
		if (cudaHostAlloc((void **)&p, want * sizeof(T), cudaHostAllocDefault) != cudaSuccess)
			return -1;
		cap = want;
		return 0;
	}
	void release()
	{
		if (p)
			cudaFreeHost(p);
		p = nullptr;
		cap = 0;
	}
};

// Device hit sink of a search call (sink_kernel.cu): compacted records + paths of all batches, grow-only, content preserved
// when it grows.
struct HitSink {
	SinkRec *rec = nullptr; size_t rec_cap = 0;    // elements
	uint8_t *pool = nullptr; size_t pool_cap = 0;  // bytes
	DevBuf<uint32_t> keep, plen, keep_scan;
	DevBuf<unsigned long long> plen_scan;
	DevBuf<uint8_t> tmp;
	unsigned long long *d_tot = nullptr;           // [4] records, path bytes, pairs with E-value, Mu-rejected pairs
	unsigned long long *h_tot = nullptr;           // pinned [2][4]: totals after the batches of either batch set
	void release()
	{
		if (rec) cudaFree(rec);
		if (pool) cudaFree(pool);
		rec = nullptr; pool = nullptr; rec_cap = pool_cap = 0;
		keep.release(); plen.release(); keep_scan.release(); plen_scan.release(); tmp.release();
		if (d_tot) cudaFree(d_tot);
		if (h_tot) cudaFreeHost(h_tot);
		d_tot = nullptr; h_tot = nullptr;
	}
};

struct rsk_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;
	bool own_stream = false;
	int num_sms = 0;
	rsk_params params;
	float *d_tables = nullptr;
	unsigned long long *d_pool_cursor = nullptr;
	cudaEvent_t ev[8] = {};
	// grow-only scratch
	void *pf_scratch = nullptr;                    // prefilter scratch (rsk_prefilter.cu), freed through pf_scratch_free
	void (*pf_scratch_free)(void *) = nullptr;
	std::vector<std::pair<void *, size_t>> slabs;  // freed chain-set slabs waiting for reuse (guarded by the context registry lock)
	DevBuf<uint8_t> upload_stage;                  // plane-major profile bytes of the upload in flight
	DevBuf<float4> ckpt;
	DevBuf<unsigned long long> tile;
	DevBuf<float4> best;
	DevBuf<float2> bnd;
	DevBuf<uint8_t> stage;
	DevBuf<PairRec> rec;
	DevBuf<uint8_t> pool;
	DevBuf<uint32_t> blist, bslot, task_a, task_begin, task_cnt, pair_a, pair_b;
	DevBuf<uint32_t> run_a, run_begin, run_cnt;  // explicit pair lists: runs of the same row chain
	DevBuf<float> lddt_scratch;  // column buffers of the LDDT kernel for alignments that do not fit shared memory
	// -global (K9): pair lists, records, path pool, per-warp trace matrices and boundary rows; kept between calls
	DevBuf<uint32_t> gl_a, gl_b, gl_order, gl_cnt;
	DevBuf<uint8_t> gl_skip, gl_tb;
	DevBuf<unsigned long long> gl_poff;
	DevBuf<char> gl_pool;
	DevBuf<GlobalRec> gl_rec;
	DevBuf<float> gl_bnd;
	// Mu filter (K3) state
	int *d_mu_mx = nullptr;                 // IntScoreMx_Mu widened to int32
	float *d_mu_f32 = nullptr;              // ScoreMx_Mu
	DevBuf<uint8_t> keep;
	DevBuf<int2> mu_bnd;
	DevBuf<uint32_t> c_blist, c_bslot, c_task_a, c_task_begin, c_task_cnt;  // compacted survivors
	size_t filt_explicit_pairs = 0; uint64_t filt_explicit_cells = 0; uint32_t filt_explicit_tasks = 0; bool batch_cross = true;
	struct Counters { uint32_t task_count[kSwClasses]; uint32_t sw_task_counter[kSwClasses]; uint32_t sat_count, mu_task_counter; unsigned long long pair_count, cell_count; };
	DevBuf<uint32_t> rowlist, colsort;
	// long-chain path (K4)
	DevBuf<uint32_t> mk_a, mk_b, mk_slot, mk_hash, mk_hchain;
	DevBuf<unsigned long long> mk_off;
	DevBuf<uint16_t> mk_ht;
	DevBuf<MkfSeed> mk_seed;
	DevBuf<MkfXdrop> mk_x;
	DevBuf<uint32_t> mk_work, mk_cnt;
	DevBuf<unsigned char> mk_scratch;
	Counters *d_counters = nullptr;
	// Second set of the per-batch device outputs: batch i+1 computes into one set while the records and paths of batch i
	// are copied out of the other on copy_stream.  swap_batch_set() exchanges the set the members above refer to.
	struct BatchSet {
		DevBuf<PairRec> rec;
		DevBuf<uint8_t> pool;
		unsigned long long *d_pool_cursor = nullptr;
		Counters *d_counters = nullptr;
		cudaEvent_t ev[5] = {};
		cudaEvent_t done = nullptr;
		bool batch_filtered = false, batch_cross = true;
		size_t filt_explicit_pairs = 0;
		uint64_t filt_explicit_cells = 0;
	} alt;
	cudaEvent_t done = nullptr;    // all kernels of the batch in the current set have been queued before this event
	cudaStream_t copy_stream = nullptr;
	void swap_batch_set()
	{
		std::swap(rec, alt.rec);
		std::swap(pool, alt.pool);
		std::swap(d_pool_cursor, alt.d_pool_cursor);
		std::swap(d_counters, alt.d_counters);
		for (int k = 0; k < 5; ++k)
			std::swap(ev[k], alt.ev[k]);
		std::swap(done, alt.done);
		std::swap(batch_filtered, alt.batch_filtered);
		std::swap(batch_cross, alt.batch_cross);
		std::swap(filt_explicit_pairs, alt.filt_explicit_pairs);
		std::swap(filt_explicit_cells, alt.filt_explicit_cells);
	}
	// DSS on the device (K10): trained tables and per-residue scratch
	DssTables *d_dss_tables = nullptr;
	DevBuf<uint8_t> dss_ss, dss_conf, dss_aa;
	DevBuf<double> dss_dens;
	DevBuf<uint32_t> dss_helix;
	PinBuf<char> h_up[2];       // pinned staging of uploads from pageable caller memory (rsk_h2d)
	cudaEvent_t ev_up[2] = {nullptr, nullptr};
	bool up_busy[2] = {false, false};
	HitSink sink;
	PinBuf<SinkRec> h_sink[2];  // pinned staging of the sink read-out (chunked, double-buffered)
	PinBuf<PairRec> h_rec[2];   // double-buffered: batch i is converted on the host while batch i+1 runs on the GPU
	PinBuf<uint8_t> h_pool[2];
	int host_threads = 1;
	rsk_stats stats;
	bool batch_filtered = false;
	size_t max_batch_pairs = 2u << 20;
	size_t scratch_budget = (size_t)24 << 30;
};

