// rsk_api.cu - C ABI of libreseek_b200: context, device chain store, batched search drivers, results.
// See include/reseek_b200.h for the reference interfaces each entry point replaces.
#include <float.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <numeric>
#include <thread>
#include <string>
#include <vector>

#include "rsk_host.cuh"
#include "rsk_multi.cuh"
#include "score_tables_data.inc"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

int rsk_fail(int code, const char *fmt, ...)
{
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	g_err = buf;
	return code;
}

// Large host blocks (hit arrays, path pools) are recycled through a small process-wide cache: a 10^7-hit result is
// ~1 GB, and handing such blocks back to the OS and faulting fresh ones in on every search call costs more host time
// than converting the records.
namespace {
struct HostBlockCache {
	static constexpr size_t kMinBytes = (size_t)16 << 20, kMaxBlocks = 6, kMaxTotal = (size_t)6 << 30;
	std::mutex mu;
	std::vector<std::pair<void *, size_t>> blocks;
	size_t total = 0;
	void *get(size_t bytes, size_t &cap)
	{
		if (bytes >= kMinBytes) {
			std::lock_guard<std::mutex> g(mu);
			int best = -1;
			for (int k = 0; k < (int)blocks.size(); ++k)
				if (blocks[k].second >= bytes && blocks[k].second <= 2 * bytes + kMinBytes && (best < 0 || blocks[k].second < blocks[best].second))
					best = k;
			if (best >= 0) {
				void *p = blocks[best].first;
				cap = blocks[best].second;
				total -= cap;
				blocks.erase(blocks.begin() + best);
				return p;
			}
		}
		cap = std::max<size_t>(bytes, 64);
		return malloc(cap);
	}
	void put(void *p, size_t cap)
	{
		if (!p)
			return;
		if (cap >= kMinBytes) {
			std::lock_guard<std::mutex> g(mu);
			if (blocks.size() < kMaxBlocks && total + cap <= kMaxTotal) {
				blocks.push_back({p, cap});
				total += cap;
				return;
			}
		}
		free(p);
	}
	~HostBlockCache()
	{
		for (auto &b : blocks)
			free(b.first);
	}
};
HostBlockCache g_blocks;
}  // namespace

struct rsk_results {
	rsk_hit *hits = nullptr;   // never value-initialised (filled by the conversion workers)
	uint64_t nhits = 0;
	size_t hits_cap = 0;       // bytes
	char *paths = nullptr;
	uint64_t npath = 0;
	size_t paths_cap = 0;      // bytes
	~rsk_results()
	{
		g_blocks.put(hits, hits_cap);
		g_blocks.put(paths, paths_cap);
	}
};

// ------------------------------------------------------------------------------------------------
// library / params
// ------------------------------------------------------------------------------------------------
extern "C" const char *rsk_version(void) { return "reseek_b200 0.1 (sm_100a)"; }
extern "C" const char *rsk_last_error(void) { return g_err.c_str(); }

extern "C" int rsk_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

extern "C" int rsk_feature_alpha(int f) { return (f >= 0 && f < RSK_NFEAT) ? rsk_tbl_feat_alpha[f] : -1; }
extern "C" int rsk_feature_offset(int f) { return (f >= 0 && f < RSK_NFEAT) ? feat_table_off(f) : -1; }
extern "C" const float *rsk_feature_bgfreq(int f)
{
	if (f < 0 || f >= RSK_NFEAT)
		return nullptr;
	int off = 0;
	for (int k = 0; k < f; ++k)
		off += rsk_tbl_feat_alpha[k];
	return rsk_tbl_bgfreq + off;
}
extern "C" const int8_t *rsk_mu_matrix_i8(void) { return (const int8_t *)rsk_tbl_mu_i8; }
extern "C" const int8_t *rsk_mu_kmer_matrix_i8(void) { return (const int8_t *)rsk_tbl_mu_kmer_i8; }
extern "C" const float *rsk_mu_matrix_f32(void) { return rsk_tbl_mu_f32; }

// DSSParams::SetDefaults (namedparams.cpp:32-53) + SetDSSParams presets (dssparams.cpp:52-81) + ApplyWeights
extern "C" int rsk_params_preset(rsk_params *p, int mode)
{
	if (!p)
		return fail(RSK_ERR_ARG, "rsk_params_preset: null params");
	memset(p, 0, sizeof(*p));
	p->gap_open = rsk_tbl_gap_open;
	p->gap_ext = rsk_tbl_gap_ext;
	p->min_fwd_score = 7.0f;
	p->mu_gap_open = 2;
	p->mu_gap_ext = 1;
	p->max_evalue = 10;
	switch (mode) {
	case RSK_MODE_FAST:
		p->omega = 22; p->omega_fwd = 50; p->mkfl = 500;
		p->mkf_x1 = 8; p->mkf_x2 = 8; p->mkf_min_hsp_score = 50; p->mkf_min_mega_hsp_score = -4;
		break;
	case RSK_MODE_SENSITIVE:
		p->omega = 12; p->omega_fwd = 20; p->mkfl = 600;
		p->mkf_x1 = 8; p->mkf_x2 = 8; p->mkf_min_hsp_score = 50; p->mkf_min_mega_hsp_score = -4;
		break;
	case RSK_MODE_VERYSENSITIVE:
		p->omega = 0; p->omega_fwd = 0; p->mkfl = 99999;
		p->mkf_x1 = 99999; p->mkf_x2 = 99999; p->mkf_min_hsp_score = 0; p->mkf_min_mega_hsp_score = -99999;
		p->min_fwd_score = 0;
		p->max_evalue = DBL_MAX;  // dbsearcher.cpp:79-80
		break;
	default:
		return fail(RSK_ERR_ARG, "rsk_params_preset: unknown mode %d (the reference dies: 'Must set -fast, -sensitive or -verysensitive')", mode);
	}
	for (int f = 0; f < RSK_NFEAT; ++f) {
		const float w = rsk_tbl_feat_weight[f];
		p->weights[f] = w;
		const int off = feat_table_off(f), n = feat_alpha(f) * feat_alpha(f);
		for (int k = 0; k < n; ++k)
			p->tables[off + k] = w * rsk_tbl_logodds[off + k];
	}
	return RSK_OK;
}

// statsig.cpp:27-50, statsig.h:8-23
extern "C" double rsk_pvalue(double ts)
{
	const double l = (ts < 0.11) ? (-80.0 * ts + -0.58) : (-52.0 * ts + -3.7);
	double P = pow(10, l);
	return P > 1 ? 1 : P;
}
extern "C" double rsk_evalue(double ts) { return rsk_pvalue(ts) * 8340; }
extern "C" double rsk_qual(double ts)
{
	const double logE = 5.0 + -40.0 * ts;
	if (logE < -20)
		return 1;
	return 1 / (1 + pow(10, logE / 10) / 2);
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
// RSK_TIMING=1 accumulators of host-side phases (developer aid; single search thread assumed)
static double g_t_plan = 0, g_t_tasks = 0, g_t_explicit = 0, g_t_keepwait = 0;
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// live contexts (a chain set may outlive its context; it may only hand memory back to a context that still exists)
static std::mutex g_ctx_mu;
static std::vector<rsk_ctx *> g_live_ctx;

static int check_params(const rsk_params *p)
{
	if (!p)
		return fail(RSK_ERR_ARG, "null params");
	if (p->gap_open > 0 || p->gap_ext > 0)
		return fail(RSK_ERR_ARG, "open=%.3g ext=%.3g, gap penalties must be >= 0", -p->gap_open, -p->gap_ext);  // dssparams.cpp:106-108
	return RSK_OK;
}

extern "C" int rsk_ctx_set_params(rsk_ctx *ctx, const rsk_params *params)
{
	if (!ctx)
		return fail(RSK_ERR_ARG, "null context");
	int rc = check_params(params);
	if (rc)
		return rc;
	ctx->params = *params;
	CK(cudaSetDevice(ctx->device));
	CK(cudaMemcpyAsync(ctx->d_tables, ctx->params.tables, sizeof(float) * RSK_TABLE_FLOATS, cudaMemcpyHostToDevice, ctx->stream));
	int mx[36 * 36];
	for (int k = 0; k < 36 * 36; ++k)
		mx[k] = rsk_tbl_mu_i8[k];
	CK(cudaMemcpyAsync(ctx->d_mu_mx, mx, sizeof(mx), cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaMemcpyAsync(ctx->d_mu_f32, rsk_tbl_mu_f32, sizeof(float) * 36 * 36, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return RSK_OK;
}

extern "C" int rsk_ctx_get_params(const rsk_ctx *ctx, rsk_params *out)
{
	if (!ctx || !out)
		return fail(RSK_ERR_ARG, "rsk_ctx_get_params: null argument");
	*out = ctx->params;
	return RSK_OK;
}

extern "C" int rsk_ctx_create(int device, const rsk_params *params, void *cuda_stream, rsk_ctx **out)
{
	if (!out)
		return fail(RSK_ERR_ARG, "rsk_ctx_create: null out");
	*out = nullptr;
	int rc = check_params(params);
	if (rc)
		return rc;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
		cudaGetLastError();
		return fail(RSK_ERR_CUDA, "no CUDA device available: libreseek_b200 has no CPU fallback");
	}
	if (device < 0 || device >= ndev)
		return fail(RSK_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
	CK(cudaSetDevice(device));
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10)
		return fail(RSK_ERR_CUDA, "device %d is sm_%d%d; this library carries sm_100a kernels only", device, prop.major, prop.minor);
	rsk_ctx *ctx = new rsk_ctx();
	ctx->device = device;
	ctx->num_sms = prop.multiProcessorCount;
	if (cuda_stream) {
		ctx->stream = (cudaStream_t)cuda_stream;
	} else {
		if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
			delete ctx;
			return fail(RSK_ERR_CUDA, "cudaStreamCreate failed");
		}
		ctx->own_stream = true;
	}
	memset(&ctx->stats, 0, sizeof(ctx->stats));
	if (cudaMalloc((void **)&ctx->d_tables, sizeof(float) * RSK_TABLE_FLOATS) != cudaSuccess ||
		cudaMalloc((void **)&ctx->d_pool_cursor, sizeof(unsigned long long)) != cudaSuccess ||
		cudaMalloc((void **)&ctx->d_mu_mx, sizeof(int) * 36 * 36) != cudaSuccess ||
		cudaMalloc((void **)&ctx->d_mu_f32, sizeof(float) * 36 * 36) != cudaSuccess ||
		cudaMalloc((void **)&ctx->d_counters, sizeof(rsk_ctx::Counters)) != cudaSuccess) {
		rsk_ctx_destroy(ctx);
		return fail(RSK_ERR_NOMEM, "cudaMalloc failed in rsk_ctx_create");
	}
	for (auto &e : ctx->ev)
		cudaEventCreate(&e);
	for (auto &e : ctx->alt.ev)
		cudaEventCreate(&e);
	cudaEventCreateWithFlags(&ctx->done, cudaEventDisableTiming);
	cudaEventCreateWithFlags(&ctx->alt.done, cudaEventDisableTiming);
	if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
		cudaMalloc((void **)&ctx->alt.d_pool_cursor, sizeof(unsigned long long)) != cudaSuccess ||
		cudaMalloc((void **)&ctx->alt.d_counters, sizeof(rsk_ctx::Counters)) != cudaSuccess) {
		rsk_ctx_destroy(ctx);
		return fail(RSK_ERR_NOMEM, "rsk_ctx_create: second batch set");
	}
	ctx->host_threads = (int)std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
	if (const char *s = getenv("RSK_HOST_THREADS")) {
		const int v = atoi(s);
		if (v > 0)
			ctx->host_threads = v;
	}
	if (const char *s = getenv("RSK_BATCH_PAIRS")) {
		const long long v = atoll(s);
		if (v > 0)
			ctx->max_batch_pairs = (size_t)v;
	}
	rc = rsk_ctx_set_params(ctx, params);
	if (rc) {
		rsk_ctx_destroy(ctx);
		return rc;
	}
	{
		std::lock_guard<std::mutex> g(g_ctx_mu);
		g_live_ctx.push_back(ctx);
	}
	*out = ctx;
	return RSK_OK;
}

extern "C" void rsk_ctx_destroy(rsk_ctx *ctx)
{
	if (!ctx)
		return;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	{
		std::lock_guard<std::mutex> g(g_ctx_mu);
		g_live_ctx.erase(std::remove(g_live_ctx.begin(), g_live_ctx.end(), ctx), g_live_ctx.end());
	}
	for (auto &sl : ctx->slabs)
		cudaFree(sl.first);
	ctx->slabs.clear();
	if (ctx->pf_scratch && ctx->pf_scratch_free)
		ctx->pf_scratch_free(ctx->pf_scratch);
	ctx->pf_scratch = nullptr;
	ctx->upload_stage.release();
	ctx->sink.release(); ctx->h_sink[0].release(); ctx->h_sink[1].release();
	ctx->h_up[0].release(); ctx->h_up[1].release();
	for (auto &e : ctx->ev_up)
		if (e) cudaEventDestroy(e);
	ctx->dss_ss.release(); ctx->dss_conf.release(); ctx->dss_aa.release(); ctx->dss_dens.release(); ctx->dss_helix.release();
	if (ctx->d_dss_tables) cudaFree(ctx->d_dss_tables);
	ctx->ckpt.release(); ctx->tile.release(); ctx->best.release(); ctx->bnd.release(); ctx->stage.release(); ctx->rec.release(); ctx->pool.release();
	ctx->blist.release(); ctx->bslot.release(); ctx->task_a.release(); ctx->task_begin.release();
	ctx->task_cnt.release(); ctx->pair_a.release(); ctx->pair_b.release();
	ctx->h_rec[0].release(); ctx->h_rec[1].release(); ctx->h_pool[0].release(); ctx->h_pool[1].release();
	ctx->keep.release(); ctx->mu_bnd.release(); ctx->c_blist.release(); ctx->c_bslot.release();
	ctx->c_task_a.release(); ctx->c_task_begin.release(); ctx->c_task_cnt.release(); ctx->rowlist.release(); ctx->colsort.release();
	ctx->run_a.release(); ctx->run_begin.release(); ctx->run_cnt.release(); ctx->lddt_scratch.release();
	ctx->gl_a.release(); ctx->gl_b.release(); ctx->gl_order.release(); ctx->gl_cnt.release(); ctx->gl_skip.release(); ctx->gl_tb.release();
	ctx->gl_poff.release(); ctx->gl_pool.release(); ctx->gl_rec.release(); ctx->gl_bnd.release();
	ctx->mk_a.release(); ctx->mk_b.release(); ctx->mk_slot.release(); ctx->mk_hash.release(); ctx->mk_hchain.release();
	ctx->mk_off.release(); ctx->mk_work.release(); ctx->mk_cnt.release(); ctx->mk_ht.release(); ctx->mk_seed.release(); ctx->mk_x.release(); ctx->mk_scratch.release();
	if (ctx->d_mu_mx) cudaFree(ctx->d_mu_mx);
	if (ctx->d_mu_f32) cudaFree(ctx->d_mu_f32);
	if (ctx->d_counters) cudaFree(ctx->d_counters);
	if (ctx->d_tables) cudaFree(ctx->d_tables);
	if (ctx->d_pool_cursor) cudaFree(ctx->d_pool_cursor);
	for (auto &e : ctx->ev)
		if (e) cudaEventDestroy(e);
	for (auto &e : ctx->alt.ev)
		if (e) cudaEventDestroy(e);
	if (ctx->done) cudaEventDestroy(ctx->done);
	if (ctx->alt.done) cudaEventDestroy(ctx->alt.done);
	if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
	ctx->alt.rec.release(); ctx->alt.pool.release();
	if (ctx->alt.d_counters) cudaFree(ctx->alt.d_counters);
	if (ctx->alt.d_pool_cursor) cudaFree(ctx->alt.d_pool_cursor);
	if (ctx->own_stream && ctx->stream)
		cudaStreamDestroy(ctx->stream);
	delete ctx;
}

extern "C" int rsk_ctx_stats(const rsk_ctx *ctx, rsk_stats *out)
{
	if (!ctx || !out)
		return fail(RSK_ERR_ARG, "rsk_ctx_stats: null argument");
	*out = ctx->stats;
	return RSK_OK;
}

extern "C" int rsk_ctx_sync(rsk_ctx *ctx)
{
	if (!ctx)
		return fail(RSK_ERR_ARG, "null context");
	CK(cudaStreamSynchronize(ctx->stream));
	return RSK_OK;
}

// ------------------------------------------------------------------------------------------------
// chain sets
// ------------------------------------------------------------------------------------------------
// Device memory of a chain set is ONE slab, recycled through its context: streaming a database through
// upload -> search -> free (DBSearcher::RunQuery's blocks) would otherwise pay a dozen cudaMalloc / cudaFree round trips
// (each cudaFree synchronises the device) per block.  A freed slab goes back to the context that made it, if that
// context is still alive, and is handed to the next upload that fits.
namespace {
constexpr size_t kSlabCacheMax = 3;

void *slab_get(rsk_ctx *ctx, size_t bytes, size_t &cap)
{
	{
		std::lock_guard<std::mutex> g(g_ctx_mu);
		int best = -1;
		for (int k = 0; k < (int)ctx->slabs.size(); ++k)
			if (ctx->slabs[k].second >= bytes && ctx->slabs[k].second <= bytes + bytes / 4 + (1 << 20) &&
				(best < 0 || ctx->slabs[k].second < ctx->slabs[best].second))
				best = k;
		if (best >= 0) {
			void *p = ctx->slabs[best].first;
			cap = ctx->slabs[best].second;
			ctx->slabs.erase(ctx->slabs.begin() + best);
			return p;
		}
	}
	void *p = nullptr;
	cap = bytes;
	if (cudaMalloc(&p, bytes) != cudaSuccess) {
		cudaGetLastError();
		// make room: drop the cached slabs and try once more
		std::vector<std::pair<void *, size_t>> drop;
		{
			std::lock_guard<std::mutex> g(g_ctx_mu);
			drop.swap(ctx->slabs);
		}
		for (auto &sl : drop)
			cudaFree(sl.first);
		if (cudaMalloc(&p, bytes) != cudaSuccess) {
			cudaGetLastError();
			return nullptr;
		}
	}
	return p;
}

void slab_put(rsk_ctx *ctx, void *p, size_t cap);
}  // namespace
void *rsk_slab_get(rsk_ctx *ctx, size_t bytes, size_t &cap) { return slab_get(ctx, bytes, cap); }  // rsk_dss.cu
namespace {
void slab_put(rsk_ctx *ctx, void *p, size_t cap)
{
	if (!p)
		return;
	{
		std::lock_guard<std::mutex> g(g_ctx_mu);
		if (std::find(g_live_ctx.begin(), g_live_ctx.end(), ctx) != g_live_ctx.end() && ctx->slabs.size() < kSlabCacheMax) {
			ctx->slabs.push_back({p, cap});
			return;
		}
	}
	cudaFree(p);
}
}  // namespace

extern "C" void rsk_chainset_free(rsk_chainset *cs)
{
	if (!cs)
		return;
	cudaSetDevice(cs->device);
	slab_put(cs->ctx, cs->slab, cs->slab_bytes);  // cs->ctx is only compared against the live contexts, never dereferenced if gone
	delete cs;
}

extern "C" uint32_t rsk_chainset_count(const rsk_chainset *cs) { return cs ? cs->d.n : 0; }
extern "C" uint64_t rsk_chainset_residues(const rsk_chainset *cs) { return cs ? cs->d.total : 0; }

// Host -> device copy of caller memory on the context stream.  Pinned (or registered) sources go straight to the copy engine.
// Pageable sources are staged through the context's two pinned buffers in 32 MB chunks: worker threads fill buffer i+1 while
// buffer i is on the bus, so a caller with ordinary malloc'ed arrays (the reference's vectors, a numpy array) gets the full
// PCIe rate instead of the driver's single-threaded bounce path.  On return the caller's memory is no longer needed unless
// *direct is set (the source was pinned and the DMA may still be reading it).
int rsk_h2d(rsk_ctx *ctx, void *dst, const void *src, size_t bytes, bool *direct)
{
	if (bytes == 0)
		return RSK_OK;
	cudaStream_t st = ctx->stream;
	cudaPointerAttributes at;
	const bool known = cudaPointerGetAttributes(&at, src) == cudaSuccess;
	cudaGetLastError();
	const bool pinned = known && (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged);
	if (pinned || bytes < ((size_t)1 << 20)) {
		CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
		if (pinned && direct)
			*direct = true;
		return RSK_OK;
	}
	const size_t chunk = (size_t)32 << 20;
	for (int k = 0; k < 2; ++k) {
		if (ctx->h_up[k].ensure(std::min(bytes, chunk)))
			return fail(RSK_ERR_NOMEM, "pinned upload staging");
		if (!ctx->ev_up[k])
			CK(cudaEventCreateWithFlags(&ctx->ev_up[k], cudaEventDisableTiming));
	}
	const size_t nchunks = (bytes + chunk - 1) / chunk;
	for (size_t c = 0; c < nchunks; ++c) {
		const size_t off = c * chunk, n = std::min(chunk, bytes - off);
		const int buf = (int)(c & 1);
		if (ctx->up_busy[buf])
			CK(cudaEventSynchronize(ctx->ev_up[buf]));  // the DMA that last read this buffer
		const int T = (int)std::max<size_t>(1, std::min<size_t>((size_t)ctx->host_threads, n >> 21));
		if (T == 1) {
			memcpy(ctx->h_up[buf].p, (const char *)src + off, n);
		} else {
			std::vector<std::thread> th;
			for (int t = 0; t < T; ++t)
				th.emplace_back([=]() {
					const size_t b0 = n * (size_t)t / T, b1 = n * (size_t)(t + 1) / T;
					memcpy(ctx->h_up[buf].p + b0, (const char *)src + off + b0, b1 - b0);
				});
			for (auto &t : th)
				t.join();
		}
		CK(cudaMemcpyAsync((char *)dst + off, ctx->h_up[buf].p, n, cudaMemcpyHostToDevice, st));
		CK(cudaEventRecord(ctx->ev_up[buf], st));
		ctx->up_busy[buf] = true;
	}
	return RSK_OK;
}

static int chainset_upload_impl(rsk_ctx *ctx, const rsk_chains_host *h, rsk_chainset **out, bool async);

extern "C" int rsk_chainset_upload(rsk_ctx *ctx, const rsk_chains_host *h, rsk_chainset **out)
{
	return chainset_upload_impl(ctx, h, out, false);
}

extern "C" int rsk_chainset_upload_async(rsk_ctx *ctx, const rsk_chains_host *h, rsk_chainset **out)
{
	return chainset_upload_impl(ctx, h, out, true);
}

static int chainset_upload_impl(rsk_ctx *ctx, const rsk_chains_host *h, rsk_chainset **out, bool async)
{
	if (!ctx || !h || !out)
		return fail(RSK_ERR_ARG, "rsk_chainset_upload: null argument");
	*out = nullptr;
	if (h->n == 0 || !h->len || !h->prof || !h->xyz)
		return fail(RSK_ERR_ARG, "rsk_chainset_upload: empty chain set or missing len/prof/xyz");
	CK(cudaSetDevice(ctx->device));
	rsk_chainset *cs = new rsk_chainset();
	cs->ctx = ctx;
	cs->device = ctx->device;
	cs->hlen.assign(h->len, h->len + h->n);
	cs->hoff.resize(h->n);
	uint64_t tot = 0;
	for (uint32_t i = 0; i < h->n; ++i) {
		if (h->len[i] == 0) {
			delete cs;
			return fail(RSK_ERR_ARG, "chain %u has length 0 (the reference's reader skips such chains, chainreader2.cpp:103-107)", i);
		}
		cs->hoff[i] = tot;
		tot += h->len[i];
		cs->maxlen = std::max(cs->maxlen, h->len[i]);
	}
	if (tot != h->total) {
		delete cs;
		return fail(RSK_ERR_ARG, "rsk_chainset_upload: total=%llu but sum(len)=%llu", (unsigned long long)h->total, (unsigned long long)tot);
	}
	DevChains &d = cs->d;
	d.n = h->n;
	d.total = tot;
	cs->has_mu = h->mu != nullptr;
	// slab layout: every array on a 256-byte boundary
	auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
	size_t o_len = 0, o_off = o_len + up(sizeof(uint32_t) * d.n), o_prof = o_off + up(sizeof(uint64_t) * d.n),
		   o_x = o_prof + up(sizeof(uint64_t) * tot), o_y = o_x + up(sizeof(float) * tot), o_z = o_y + up(sizeof(float) * tot),
		   o_sr = o_z + up(sizeof(float) * tot), o_mu = o_sr + up(sizeof(float) * d.n), o_end = o_mu + (h->mu ? up(tot) : 0);
	cs->slab = slab_get(ctx, o_end, cs->slab_bytes);
	if (!cs->slab || ctx->upload_stage.ensure((size_t)RSK_NFEAT * tot)) {
		cudaGetLastError();
		rsk_chainset_free(cs);
		return fail(RSK_ERR_NOMEM, "rsk_chainset_upload: device memory for %llu residues", (unsigned long long)tot);
	}
	unsigned char *base = (unsigned char *)cs->slab;
	d.len = (uint32_t *)(base + o_len); d.off = (uint64_t *)(base + o_off); d.prof8 = (uint64_t *)(base + o_prof);
	d.x = (float *)(base + o_x); d.y = (float *)(base + o_y); d.z = (float *)(base + o_z); d.selfrev = (float *)(base + o_sr);
	d.mu = h->mu ? (uint8_t *)(base + o_mu) : nullptr;
	uint8_t *d_planes = ctx->upload_stage.p;
	cudaStream_t st = ctx->stream;
	std::vector<float> sr(d.n, FLT_MAX);
	if (h->selfrev)
		sr.assign(h->selfrev, h->selfrev + d.n);
	cudaError_t e = cudaSuccess;
	bool direct = false;  // some caller buffer is pinned and read by the DMA itself
	auto cp = [&](void *dst, const void *src, size_t bytes) {
		if (e == cudaSuccess && rsk_h2d(ctx, dst, src, bytes, &direct) != RSK_OK)
			e = cudaErrorUnknown;
	};
	cp(d.len, h->len, sizeof(uint32_t) * d.n);
	if (e == cudaSuccess)
		e = cudaMemcpyAsync(d.off, cs->hoff.data(), sizeof(uint64_t) * d.n, cudaMemcpyHostToDevice, st);  // lives in the set
	cp(d_planes, h->prof, (size_t)RSK_NFEAT * tot);
	cp(d.x, h->xyz, sizeof(float) * tot);
	cp(d.y, h->xyz + tot, sizeof(float) * tot);
	cp(d.z, h->xyz + 2 * tot, sizeof(float) * tot);
	if (h->mu)
		cp(d.mu, h->mu, tot);
	if (e == cudaSuccess && h->selfrev)
		cp(d.selfrev, h->selfrev, sizeof(float) * d.n);
	else if (e == cudaSuccess) {
		e = cudaMemcpyAsync(d.selfrev, sr.data(), sizeof(float) * d.n, cudaMemcpyHostToDevice, st);  // pageable: staged before it returns
	}
	int nl = 0;
	if (e == cudaSuccess) {
		nl = launch_pack_profiles(d_planes, tot, d.prof8, st);
		if (nl < 0)
			e = cudaErrorLaunchFailure;
	}
	// Everything is queued on the context stream, in order before any later search call.  The synchronous entry waits (the
	// caller may free or overwrite its arrays at once); the asynchronous one only promises that PAGEABLE sources have been
	// read - pinned sources must stay untouched until rsk_ctx_sync() or the next call that returns results.
	if (e == cudaSuccess && !async)
		e = cudaStreamSynchronize(st);
	(void)direct;
	if (e != cudaSuccess) {
		rsk_chainset_free(cs);
		return fail(RSK_ERR_CUDA, "rsk_chainset_upload: %s", cudaGetErrorString(e));
	}
	ctx->stats.h2d_bytes += (uint64_t)tot * (RSK_NFEAT + 12 + (h->mu ? 1 : 0)) + (uint64_t)d.n * 16;
	ctx->stats.kernel_launches += nl;
	*out = cs;
	return RSK_OK;
}

// ------------------------------------------------------------------------------------------------
// search drivers
// ------------------------------------------------------------------------------------------------
namespace {

struct Batch {
	// cross mode: A range [a0,a1) x all B
	bool cross = true;
	uint32_t a0 = 0, a1 = 0;
	// explicit mode: sorted pair range [k0,k1)
	size_t k0 = 0, k1 = 0;
	size_t npairs = 0;
	uint32_t ntasks = 0, nruns = 0, sw_task_cap = 0;
	uint32_t maxLA = 0, maxLB = 0;
	uint64_t cells = 0;
	uint64_t pool_bound = 0;
};

// run fn(lo, hi, t) over [0, n) cut into T contiguous ranges, on T host threads (T = 1: inline)
template <typename F>
static void parallel_ranges(uint64_t n, int T, F fn)
{
	T = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)std::max(1, T), n >> 16));
	if (T == 1) {
		fn((uint64_t)0, n, 0);
		return;
	}
	std::vector<std::thread> th;
	for (int t = 0; t < T; ++t)
		th.emplace_back([&, t]() { fn(n * t / T, n * (t + 1) / T, t); });
	for (auto &x : th)
		x.join();
}

struct SearchPlan {
	const rsk_chainset *A = nullptr, *B = nullptr;
	bool cross = true;
	// cross: B indices sorted by length
	std::vector<uint32_t> border;
	// explicit: pairs sorted by (a, lenB); perm[k] = original pair index
	std::vector<uint32_t> sa, sb;
	std::vector<uint64_t> perm;
	uint64_t npairs = 0;
};

int ensure_scratch(rsk_ctx *ctx, uint32_t maxRow, uint32_t maxCol, int &grid, uint64_t &ckpt_stride, uint64_t &bnd_stride,
		uint32_t &bnd_pass_stride, uint32_t &stage_stride, uint32_t &stage_chain_stride)
{
	int npass, R;
	sw_geometry(maxRow, npass, R);
	// every pair of the batch has npass(rows) <= npass(maxRow) and columns <= maxCol; a warp sweeps up to kSwChain column
	// chains as one concatenated wavefront
	const uint64_t cols = (uint64_t)maxCol * kSwChainMax;
	ckpt_stride = sw_ckpt_units(npass, cols);  // float4 units
	bnd_pass_stride = (uint32_t)(((cols + 3) & ~(uint64_t)3) + 4);
	bnd_stride = (uint64_t)bnd_pass_stride * (uint64_t)npass;
	stage_chain_stride = ((maxRow + maxCol + 16) + 15) & ~15u;
	stage_stride = stage_chain_stride * kSwChainMax;
	const uint64_t per_cta = (ckpt_stride * 16 + bnd_stride * 8 + stage_stride + kSwStripSteps * 32 * 8 + kSwChainMax * 32 * 16) * kSwMaxWarps;
	grid = ctx->num_sms;
	if (per_cta * (uint64_t)grid > ctx->scratch_budget)
		grid = (int)std::max<uint64_t>(1, ctx->scratch_budget / per_cta);
	const size_t warps = (size_t)grid * kSwMaxWarps;
	if (ctx->ckpt.ensure(ckpt_stride * warps) || ctx->bnd.ensure((size_t)bnd_stride * warps) ||
		ctx->tile.ensure((size_t)kSwStripSteps * 32 * warps) || ctx->stage.ensure((size_t)stage_stride * warps) ||
		ctx->best.ensure((size_t)kSwChainMax * 32 * warps)) {
		cudaGetLastError();
		return fail(RSK_ERR_NOMEM, "SW scratch allocation failed (rows=%u cols=%u)", maxRow, maxCol);
	}
	return RSK_OK;
}

template <typename T>
int upload_vec(rsk_ctx *ctx, DevBuf<T> &dst, const std::vector<T> &src)
{
	if (dst.ensure(std::max<size_t>(1, src.size())))
		return fail(RSK_ERR_NOMEM, "device task buffer (%zu elements)", src.size());
	if (!src.empty())
		CK(cudaMemcpyAsync(dst.p, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
	ctx->stats.h2d_bytes += src.size() * sizeof(T);
	return RSK_OK;
}

// Long-chain pairs of a batch through K4 (hash tables, seeds/chain/re-score, banded x-drop, merge), in chunks that
// respect the scratch budget.  Pair order is A-major so that one hash table serves consecutive pairs.
int run_mkf(rsk_ctx *ctx, const SearchPlan &plan, const Batch &b)
{
	const rsk_chainset *A = plan.A, *B = plan.B;
	cudaStream_t st = ctx->stream;
	const uint32_t mkfl = ctx->params.mkfl;
	std::vector<uint32_t> pa, pb, ps;
	auto is_mkf = [&](uint32_t la, uint32_t lb) { return la >= 3 && lb >= 3 && (la >= mkfl || lb >= mkfl); };
	if (b.cross) {
		const uint32_t nB = B->d.n;
		std::vector<uint32_t> longB;
		for (uint32_t j = 0; j < nB; ++j)
			if (B->hlen[j] >= mkfl)
				longB.push_back(j);
		for (uint32_t a = b.a0; a < b.a1; ++a) {
			const uint32_t la = A->hlen[a];
			if (la < 3)
				continue;
			if (la >= mkfl) {
				for (uint32_t j = 0; j < nB; ++j)
					if (B->hlen[j] >= 3) { pa.push_back(a); pb.push_back(j); ps.push_back((a - b.a0) * nB + j); }
			} else {
				for (uint32_t j : longB) { pa.push_back(a); pb.push_back(j); ps.push_back((a - b.a0) * nB + j); }
			}
		}
	} else {
		for (size_t k = 0; k < b.npairs; ++k) {
			const uint32_t a = plan.sa[b.k0 + k], j = plan.sb[b.k0 + k];
			if (is_mkf(A->hlen[a], B->hlen[j])) { pa.push_back(a); pb.push_back(j); ps.push_back((uint32_t)k); }
		}
	}
	const size_t n = pa.size();
	if (n == 0)
		return RSK_OK;
	ctx->stats.mkf_pairs += n;
	const bool timing = getenv("RSK_TIMING") != nullptr;
	auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	double t_prep = 0, t_sync = 0, t_pre = 0;
	size_t nchunks = 0;
	if (timing) {
		const double t0 = now();
		cudaStreamSynchronize(st);  // what was queued before (Mu filter, SW): keep it out of the long-chain figures
		t_pre = now() - t0;
	}
	const size_t scratch_cap = (size_t)40 << 30;  // trace matrices of the pairs in flight
	const size_t max_hash = 4096;
	size_t k0 = 0;
	while (k0 < n) {
		const double tc0 = now();
		++nchunks;
		std::vector<uint32_t> hidx, hchain;
		std::vector<unsigned long long> off;
		size_t bytes = 0, k1 = k0;
		while (k1 < n) {
			const size_t need = (xdrop_pair_bytes(A->hlen[pa[k1]], B->hlen[pb[k1]]) + 255) & ~(size_t)255;
			const bool newchain = hchain.empty() || hchain.back() != pa[k1];
			if (k1 > k0 && (bytes + need > scratch_cap || (newchain && hchain.size() >= max_hash)))
				break;
			if (newchain)
				hchain.push_back(pa[k1]);
			hidx.push_back((uint32_t)hchain.size() - 1);
			off.push_back(bytes);
			bytes += need;
			++k1;
		}
		const size_t m = k1 - k0;
		if (ctx->mk_a.ensure(m) || ctx->mk_b.ensure(m) || ctx->mk_slot.ensure(m) || ctx->mk_hash.ensure(m) ||
			ctx->mk_hchain.ensure(hchain.size()) || ctx->mk_off.ensure(m) || ctx->mk_seed.ensure(m) || ctx->mk_x.ensure(2 * m) || ctx->mk_work.ensure(2 * m) || ctx->mk_cnt.ensure(16) ||
			ctx->mk_ht.ensure(hchain.size() * (mkf_hash_bytes() / 2)) || ctx->mk_scratch.ensure(bytes + 256)) {
			cudaGetLastError();
			return fail(RSK_ERR_NOMEM, "long-chain path buffers (%zu pairs, %zu MB scratch)", m, bytes >> 20);
		}
		CK(cudaMemcpyAsync(ctx->mk_a.p, pa.data() + k0, 4 * m, cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(ctx->mk_b.p, pb.data() + k0, 4 * m, cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(ctx->mk_slot.p, ps.data() + k0, 4 * m, cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(ctx->mk_hash.p, hidx.data(), 4 * m, cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(ctx->mk_hchain.p, hchain.data(), 4 * hchain.size(), cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(ctx->mk_off.p, off.data(), 8 * m, cudaMemcpyHostToDevice, st));
		CK(cudaMemsetAsync(ctx->mk_scratch.p, 0, bytes, st));  // trace matrices start zeroed
		ctx->stats.h2d_bytes += 24 * m + 4 * hchain.size();
		MkfArgs ma;
		memset(&ma, 0, sizeof(ma));
		ma.muA = A->d.mu; ma.profA = A->d.prof8; ma.offA = A->d.off; ma.lenA = A->d.len;
		ma.muB = B->d.mu; ma.profB = B->d.prof8; ma.offB = B->d.off; ma.lenB = B->d.len;
		ma.npairs = (uint32_t)m;
		ma.pair_a = ctx->mk_a.p; ma.pair_b = ctx->mk_b.p; ma.pair_slot = ctx->mk_slot.p; ma.pair_hash = ctx->mk_hash.p;
		ma.hash_chain = ctx->mk_hchain.p; ma.hash = ctx->mk_ht.p;
		ma.seeds = ctx->mk_seed.p; ma.xres = ctx->mk_x.p; ma.xwork = ctx->mk_work.p; ma.xcnt = ctx->mk_cnt.p;
		ma.scratch = ctx->mk_scratch.p; ma.scratch_off = ctx->mk_off.p;
		ma.rec = ctx->rec.p; ma.pool = ctx->pool.p; ma.pool_cursor = ctx->d_pool_cursor;
		ma.mu_mx = ctx->d_mu_mx; ma.tables = ctx->d_tables;
		ma.x1 = ctx->params.mkf_x1; ma.min_hsp_score = ctx->params.mkf_min_hsp_score;
		ma.x2 = (float)ctx->params.mkf_x2; ma.min_mega_hsp_score = ctx->params.mkf_min_mega_hsp_score;
		ma.open = ctx->params.gap_open; ma.ext = ctx->params.gap_ext;
		// x-drop grid: the lanes balance their load by pulling items from the work list, which needs several items per
		// thread; too few threads on the other hand cannot hide the memory latency of the DP rows
		static const int xgrid_env = getenv("RSK_XGRID") ? std::max(1, atoi(getenv("RSK_XGRID"))) : 0;  // blocks per SM (A/B switch)
		// the warp-per-item kernel (80 registers, 4 warps per CTA) fits 6 CTAs per SM; its warps pull items from the work list
		int xblocks = ctx->num_sms * (xgrid_env ? xgrid_env : 6);
		int nl = launch_mkf(ma, (uint32_t)hchain.size(), xblocks, st);
		if (nl < 0)
			return fail(RSK_ERR_CUDA, "long-chain kernels failed to launch: %s", cudaGetErrorString(cudaGetLastError()));
		ctx->stats.kernel_launches += nl;
		const double ts0 = now();
		t_prep += ts0 - tc0;
		CK(cudaStreamSynchronize(st));  // the host vectors of this chunk are reused
		t_sync += now() - ts0;
		k0 = k1;
	}
	if (timing)
		fprintf(stderr, "[rsk run_mkf] %zu pairs in %zu chunks: earlier kernels %.1f ms, host prep + launches %.1f ms, waiting for the long-chain kernels %.1f ms\n",
				n, nchunks, t_pre, t_prep, t_sync);
	return RSK_OK;
}

// Run one batch on the device: (Mu filter + compaction), SW+traceback, then LDDT/TS.  Records land in ctx->rec[0..npairs).
// Task model: a task is one "row" chain (its score table is staged in shared memory by the CTA) against up to W
// "column" chains.  In cross mode the side with fewer chains supplies the rows (tr = rows are the reference's B).
int run_batch(rsk_ctx *ctx, const SearchPlan &plan, const Batch &b, const rsk_search_opts &opts)
{
	const rsk_chainset *A = plan.A, *B = plan.B;
	cudaStream_t st = ctx->stream;
	const uint32_t nA_b = b.cross ? (b.a1 - b.a0) : 0;
	const bool tr = b.cross && B->d.n <= nA_b;
	const rsk_chainset *Rw = tr ? B : A, *Cl = tr ? A : B;
	const uint32_t maxRow = tr ? b.maxLB : b.maxLA, maxCol = tr ? b.maxLA : b.maxLB;
	int grid;
	uint64_t ckpt_stride, bnd_stride;
	uint32_t bnd_pass_stride, stage_stride, stage_chain_stride;
	int rc = ensure_scratch(ctx, maxRow, maxCol, grid, ckpt_stride, bnd_stride, bnd_pass_stride, stage_stride, stage_chain_stride);
	if (rc)
		return rc;
	if (ctx->rec.ensure(b.npairs) || ctx->pool.ensure((size_t)b.pool_bound + 64)) {
		cudaGetLastError();
		return fail(RSK_ERR_NOMEM, "result buffers: %zu pairs, %llu path bytes", b.npairs, (unsigned long long)b.pool_bound);
	}
	CK(cudaMemsetAsync(ctx->rec.p, 0, b.npairs * sizeof(PairRec), st));
	CK(cudaMemsetAsync(ctx->d_pool_cursor, 0, sizeof(unsigned long long), st));
	CK(cudaMemsetAsync(ctx->d_counters, 0, sizeof(rsk_ctx::Counters), st));

	const bool filter = ctx->params.omega > 0 && A->has_mu && B->has_mu;

	// ---- row lists per kernel class and the column list (cross mode) ----
	std::vector<uint32_t> rowlist;            // rows grouped by class
	uint32_t row_off[kSwClasses + 1] = {};
	uint32_t ncols = 0;
	if (b.cross) {
		std::vector<uint32_t> byclass[kSwClasses];
		const uint32_t r0 = tr ? 0 : b.a0, r1 = tr ? B->d.n : b.a1;
		for (uint32_t r = r0; r < r1; ++r)
			byclass[sw_class_of_len(Rw->hlen[r])].push_back(r);
		for (int c = 0; c < kSwClasses; ++c) {
			row_off[c] = (uint32_t)rowlist.size();
			rowlist.insert(rowlist.end(), byclass[c].begin(), byclass[c].end());
		}
		row_off[kSwClasses] = (uint32_t)rowlist.size();
		if ((rc = upload_vec(ctx, ctx->rowlist, rowlist)))
			return rc;
		if (tr) {
			// columns = this batch's A range, longest first so that the chains of a task have similar lengths
			std::vector<uint32_t> cl(nA_b);
			std::iota(cl.begin(), cl.end(), b.a0);
			std::stable_sort(cl.begin(), cl.end(), [&](uint32_t x, uint32_t y) { return A->hlen[x] > A->hlen[y]; });
			if ((rc = upload_vec(ctx, ctx->colsort, cl)))
				return rc;
			ncols = nA_b;
		} else {
			ncols = B->d.n;  // ctx->blist already holds B sorted by length (uploaded once per search call)
		}
	}

	// ---- Mu filter (K3) + survivor compaction: only when the reference would run MuFilter (dssaligner.cpp:819-829) ----
	CK(cudaEventRecord(ctx->ev[3], st));
	uint32_t task_cap = 0;
	if (filter) {
		const size_t warps = (size_t)ctx->num_sms * 2 * kSwWarps;
		const uint32_t mu_bnd_stride = ((maxCol + 3) & ~3u) + 4;
		if (ctx->keep.ensure(b.npairs) || ctx->mu_bnd.ensure((size_t)mu_bnd_stride * warps)) {
			cudaGetLastError();
			return fail(RSK_ERR_NOMEM, "Mu filter buffers for %zu pairs", b.npairs);
		}
		MuArgs ma;
		memset(&ma, 0, sizeof(ma));
		ma.mu_row = Rw->d.mu; ma.off_row = Rw->d.off; ma.len_row = Rw->d.len;
		ma.mu_col = Cl->d.mu; ma.off_col = Cl->d.off; ma.len_col = Cl->d.len;
		ma.tr = tr ? 1 : 0;
		ma.cross = b.cross ? 1 : 0;
		ma.a_begin = b.a0; ma.nB = B->d.n;
		// packed 16-bit lanes (two column chains per warp) unless a pair could exceed their range
		const bool mu16 = std::min(b.maxLA, b.maxLB) <= kMu16MaxLen && !getenv("RSK_MU32");
		// cross mode numbers the packed kernel's segments in units of kMuTaskCols / 2 columns (a short row chain takes two at a time)
		const uint32_t task_cols = mu16 ? (uint32_t)kMuTaskCols / 2 : (uint32_t)kSwWarps;
		if (b.cross) {
			ma.nseg = (ncols + task_cols - 1) / task_cols;
			ma.ncols = ncols;
			ma.rowlist = ctx->rowlist.p;
			ma.ntasks = (uint32_t)rowlist.size() * ma.nseg;
			ma.clist = tr ? ctx->colsort.p : ctx->blist.p;
		} else {
			ma.ntasks = b.ntasks;
			ma.task_row = ctx->task_a.p; ma.task_begin = ctx->task_begin.p; ma.task_cnt = ctx->task_cnt.p;
			ma.clist = ctx->blist.p; ma.cslot = ctx->bslot.p;
		}
		ma.bnd = ctx->mu_bnd.p; ma.bnd_stride = mu_bnd_stride;
		ma.rec = ctx->rec.p; ma.keep = ctx->keep.p;
		ma.task_counter = &ctx->d_counters->mu_task_counter;
		ma.sat_counter = &ctx->d_counters->sat_count;
		ma.mu_mx = ctx->d_mu_mx;
		ma.open = ctx->params.mu_gap_open; ma.ext = ctx->params.mu_gap_ext;
		ma.omega = ctx->params.omega; ma.omega_fwd = ctx->params.omega_fwd;
		ma.mkfl = ctx->params.mkfl;
		const int mu_grid = (int)std::min<uint64_t>((uint64_t)ctx->num_sms * 2, std::max<uint32_t>(1, ma.ntasks));
		int nlm = mu16 ? launch_mu_filter16(ma, mu_grid, st) : launch_mu_filter(ma, mu_grid, st);
		if (nlm < 0)
			return fail(RSK_ERR_CUDA, "Mu filter kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
		ctx->stats.kernel_launches += nlm;
		{
			// DoMKF() pairs return from AlignQueryTarget before ++m_MuFilterInputCount (dssaligner.cpp:811-820)
			uint64_t nmkf = 0;
			const uint32_t mkfl = ctx->params.mkfl;
			if (b.maxLA >= mkfl || b.maxLB >= mkfl) {
				if (b.cross) {
					uint64_t longB = 0, okB = 0;
					for (uint32_t j = 0; j < B->d.n; ++j) {
						longB += B->hlen[j] >= mkfl && B->hlen[j] >= 3;
						okB += B->hlen[j] >= 3;
					}
					for (uint32_t a = b.a0; a < b.a1; ++a)
						if (A->hlen[a] >= 3)
							nmkf += A->hlen[a] >= mkfl ? okB : longB;
				} else {
					std::vector<uint64_t> cnt((size_t)std::max(1, ctx->host_threads) * 8, 0);
					parallel_ranges(b.npairs, ctx->host_threads, [&](uint64_t lo, uint64_t hi, int t) {
						uint64_t c = 0;
						for (uint64_t k = lo; k < hi; ++k) {
							const uint32_t la = A->hlen[plan.sa[b.k0 + k]], lb = B->hlen[plan.sb[b.k0 + k]];
							c += la >= 3 && lb >= 3 && (la >= mkfl || lb >= mkfl);
						}
						cnt[(size_t)t * 8] = c;
					});
					for (uint64_t c : cnt)
						nmkf += c;
				}
			}
			ctx->stats.mu_filter_in += b.npairs - nmkf;
		}
		if (b.cross) {
			const uint32_t nrows = (uint32_t)rowlist.size();
			task_cap = nrows * (ncols / (16 * kSwChain) + 10);  // sw_task_chains: at most 9 small tasks per row chain, or full lists of >= 16 * kSwChain
			if (ctx->c_blist.ensure(b.npairs) || ctx->c_bslot.ensure(b.npairs) || ctx->c_task_a.ensure((size_t)task_cap * kSwClasses) ||
				ctx->c_task_begin.ensure((size_t)task_cap * kSwClasses) || ctx->c_task_cnt.ensure((size_t)task_cap * kSwClasses)) {
				cudaGetLastError();
				return fail(RSK_ERR_NOMEM, "survivor task buffers for %zu pairs", b.npairs);
			}
			CompactArgs ca;
			memset(&ca, 0, sizeof(ca));
			ca.tr = tr ? 1 : 0; ca.a_begin = b.a0; ca.nB = B->d.n;
			ca.rowlist = ctx->rowlist.p; ca.ncols = ncols; ca.clist = tr ? ctx->colsort.p : ctx->blist.p; ca.keep = ctx->keep.p;
			ca.len_row = Rw->d.len; ca.len_col = Cl->d.len;
			ca.out_clist = ctx->c_blist.p; ca.out_cslot = ctx->c_bslot.p;
			ca.task_row = ctx->c_task_a.p; ca.task_begin = ctx->c_task_begin.p; ca.task_cnt = ctx->c_task_cnt.p;
			ca.task_cap = task_cap;
			ca.task_count = ctx->d_counters->task_count;
			ca.pair_count = &ctx->d_counters->pair_count;
			ca.cell_count = &ctx->d_counters->cell_count;
			nlm = launch_compact_survivors(ca, nrows, st);
			if (nlm < 0)
				return fail(RSK_ERR_CUDA, "compaction kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
			ctx->stats.kernel_launches += nlm;
		}
	}

	// ---- explicit pair lists: tasks per class are built on the host (lists are small: PostMuFilter, -alignpair, self) ----
	std::vector<uint32_t> e_row[kSwClasses], e_begin[kSwClasses], e_cnt[kSwClasses];
	std::vector<uint32_t> e_clist, e_cslot;
	uint32_t e_off[kSwClasses + 1] = {};
	const bool dev_tasks = !b.cross && filter && !getenv("RSK_HOST_TASKS");  // SW tasks of explicit lists built on the device
	if (dev_tasks) {
		task_cap = b.sw_task_cap;
		if (ctx->c_blist.ensure(b.npairs) || ctx->c_bslot.ensure(b.npairs) || ctx->c_task_a.ensure((size_t)task_cap * kSwClasses) ||
			ctx->c_task_begin.ensure((size_t)task_cap * kSwClasses) || ctx->c_task_cnt.ensure((size_t)task_cap * kSwClasses)) {
			cudaGetLastError();
			return fail(RSK_ERR_NOMEM, "survivor task buffers for %zu pairs", b.npairs);
		}
		CompactArgs ca;
		memset(&ca, 0, sizeof(ca));
		ca.run_row = ctx->run_a.p; ca.run_begin = ctx->run_begin.p; ca.run_cnt = ctx->run_cnt.p;
		ca.clist = ctx->blist.p; ca.keep = ctx->keep.p;
		ca.len_row = Rw->d.len; ca.len_col = Cl->d.len;
		ca.out_clist = ctx->c_blist.p; ca.out_cslot = ctx->c_bslot.p;
		ca.task_row = ctx->c_task_a.p; ca.task_begin = ctx->c_task_begin.p; ca.task_cnt = ctx->c_task_cnt.p;
		ca.task_cap = task_cap;
		ca.task_count = ctx->d_counters->task_count;
		ca.pair_count = &ctx->d_counters->pair_count;
		ca.cell_count = &ctx->d_counters->cell_count;
		const int nlc = launch_compact_explicit(ca, b.nruns, st);
		if (nlc < 0)
			return fail(RSK_ERR_CUDA, "compaction kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
		ctx->stats.kernel_launches += nlc;
	}
	if (!b.cross && !dev_tasks) {
		std::vector<uint8_t> keep;
		const double tk0 = now_ms();
		if (filter) {
			keep.resize(b.npairs);
			CK(cudaMemcpyAsync(keep.data(), ctx->keep.p, b.npairs, cudaMemcpyDeviceToHost, st));
			CK(cudaStreamSynchronize(st));
			ctx->stats.d2h_bytes += b.npairs;
		}
		const double tk1 = now_ms();
		g_t_keepwait += tk1 - tk0;
		size_t k = 0;
		uint64_t cells = 0;
		while (k < b.npairs) {
			const uint32_t a0 = plan.sa[b.k0 + k];
			const int cls = sw_class_of_len(A->hlen[a0]);
			const uint32_t W = (uint32_t)kClassWarps[cls] * sw_class_chains(cls);  // column chains per SW task
			const uint32_t begin = (uint32_t)e_clist.size();
			uint32_t cnt = 0;
			while (k < b.npairs && plan.sa[b.k0 + k] == a0 && cnt < W) {
				if (!filter || keep[k]) {
					e_clist.push_back(plan.sb[b.k0 + k]);
					e_cslot.push_back((uint32_t)k);
					cells += (uint64_t)A->hlen[a0] * B->hlen[plan.sb[b.k0 + k]];
					++cnt;
				}
				++k;
			}
			if (cnt) {
				e_row[cls].push_back(a0); e_begin[cls].push_back(begin); e_cnt[cls].push_back(cnt);
			}
		}
		std::vector<uint32_t> trow, tbegin, tcnt;
		for (int c = 0; c < kSwClasses; ++c) {
			e_off[c] = (uint32_t)trow.size();
			trow.insert(trow.end(), e_row[c].begin(), e_row[c].end());
			tbegin.insert(tbegin.end(), e_begin[c].begin(), e_begin[c].end());
			tcnt.insert(tcnt.end(), e_cnt[c].begin(), e_cnt[c].end());
		}
		e_off[kSwClasses] = (uint32_t)trow.size();
		if ((rc = upload_vec(ctx, ctx->c_blist, e_clist)) || (rc = upload_vec(ctx, ctx->c_bslot, e_cslot)) ||
			(rc = upload_vec(ctx, ctx->c_task_a, trow)) || (rc = upload_vec(ctx, ctx->c_task_begin, tbegin)) ||
			(rc = upload_vec(ctx, ctx->c_task_cnt, tcnt)))
			return rc;
		CK(cudaStreamSynchronize(st));  // the host vectors die at the end of this scope
		ctx->filt_explicit_pairs = e_clist.size();
		ctx->filt_explicit_cells = cells;
		g_t_explicit += now_ms() - tk1;
	}

	// ---- K1: one launch per kernel class ----
	SwArgs sa;
	memset(&sa, 0, sizeof(sa));
	sa.prof_row = Rw->d.prof8; sa.off_row = Rw->d.off; sa.len_row = Rw->d.len;
	sa.prof_col = Cl->d.prof8; sa.off_col = Cl->d.off; sa.len_col = Cl->d.len;
	sa.tr = tr ? 1 : 0;
	sa.a_begin = b.a0; sa.nB = B->d.n;
	sa.ckpt = ctx->ckpt.p; sa.ckpt_stride = ckpt_stride;
	sa.tile = ctx->tile.p;
	sa.bnd = ctx->bnd.p; sa.bnd_stride = bnd_stride; sa.bnd_pass_stride = bnd_pass_stride;
	sa.stage = ctx->stage.p; sa.stage_stride = stage_stride; sa.stage_chain_stride = stage_chain_stride;
	sa.best = ctx->best.p;
	sa.rec = ctx->rec.p;
	sa.pool = ctx->pool.p; sa.pool_cursor = ctx->d_pool_cursor;
	sa.tables = ctx->d_tables;
	sa.open = ctx->params.gap_open; sa.ext = ctx->params.gap_ext;

	CK(cudaEventRecord(ctx->ev[0], st));
	for (int c = 0; c < kSwClasses; ++c) {
		SwArgs sc = sa;
		sc.task_counter = &ctx->d_counters->sw_task_counter[c];
		int g = grid;
		if (b.cross && !filter) {
			const uint32_t nrows = row_off[c + 1] - row_off[c];
			if (nrows == 0)
				continue;
			sc.cross = 1;
			sc.rowlist = ctx->rowlist.p + row_off[c];
			sc.ncols = ncols;
			sc.nseg = (ncols + kClassWarps[c] * sw_class_chains(c) - 1) / (kClassWarps[c] * sw_class_chains(c));
			sc.clist = tr ? ctx->colsort.p : ctx->blist.p;
			sc.ntasks = nrows * sc.nseg;
			g = (int)std::min<uint64_t>((uint64_t)grid, sc.ntasks);
		} else if (b.cross || dev_tasks) {
			if (b.cross && row_off[c + 1] == row_off[c])
				continue;  // no row chain of this class: its task list stays empty
			sc.cross = 0;
			sc.task_row = ctx->c_task_a.p + (size_t)c * task_cap;
			sc.task_begin = ctx->c_task_begin.p + (size_t)c * task_cap;
			sc.task_cnt = ctx->c_task_cnt.p + (size_t)c * task_cap;
			sc.clist = ctx->c_blist.p; sc.cslot = ctx->c_bslot.p;
			sc.ntasks_dev = &ctx->d_counters->task_count[c];
		} else {
			const uint32_t nt = e_off[c + 1] - e_off[c];
			if (nt == 0)
				continue;
			sc.cross = 0;
			sc.task_row = ctx->c_task_a.p + e_off[c];
			sc.task_begin = ctx->c_task_begin.p + e_off[c];
			sc.task_cnt = ctx->c_task_cnt.p + e_off[c];
			sc.clist = ctx->c_blist.p; sc.cslot = ctx->c_bslot.p;
			sc.ntasks = nt;
			g = (int)std::min<uint32_t>((uint32_t)grid, nt);
		}
		int nl = launch_sw(sc, c, g, st);
		if (nl < 0)
			return fail(RSK_ERR_CUDA, "SW kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
		ctx->stats.kernel_launches += nl;
		ctx->stats.sw_kernel_launches += (uint32_t)nl;
	}
	CK(cudaEventRecord(ctx->ev[1], st));

	// ---- K4: pairs with a chain >= MKFL take the k-mer / x-drop path instead (dssaligner.cpp:811-815, 715-732) ----
	if (A->has_mu && B->has_mu && (b.maxLA >= ctx->params.mkfl || b.maxLB >= ctx->params.mkfl)) {
		rc = run_mkf(ctx, plan, b);
		if (rc)
			return rc;
	}
	CK(cudaEventRecord(ctx->ev[4], st));

	if (!opts.skip_evalue) {
		LddtArgs la;
		memset(&la, 0, sizeof(la));
		la.lenA = A->d.len; la.offA = A->d.off; la.xA = A->d.x; la.yA = A->d.y; la.zA = A->d.z; la.selfrevA = A->d.selfrev;
		la.lenB = B->d.len; la.offB = B->d.off; la.xB = B->d.x; la.yB = B->d.y; la.zB = B->d.z; la.selfrevB = B->d.selfrev;
		la.npairs = (uint32_t)b.npairs;
		la.cross = b.cross ? 1 : 0;
		la.a_begin = b.a0;
		la.nB = B->d.n;
		la.pair_a = ctx->pair_a.p; la.pair_b = ctx->pair_b.p;
		la.rec = ctx->rec.p;
		la.pool = ctx->pool.p;
		la.min_fwd_score = ctx->params.min_fwd_score;
		la.maxcols = std::max(1u, std::min(b.maxLA, b.maxLB));
		if (const size_t nf = lddt_scratch_floats(la.maxcols)) {
			if (ctx->lddt_scratch.ensure(nf))
				return fail(RSK_ERR_NOMEM, "LDDT column buffers for alignments of %u columns", la.maxcols);
			la.scratch = ctx->lddt_scratch.p;
		}
		int nl = launch_lddt(la, st);
		if (nl < 0)
			return fail(RSK_ERR_CUDA, "LDDT kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
		ctx->stats.kernel_launches += nl;
	}
	CK(cudaEventRecord(ctx->ev[2], st));
	if (!filter) {
		ctx->stats.sw_pairs += b.npairs;
		ctx->stats.sw_cells += b.cells;
	}
	ctx->batch_filtered = filter;
	ctx->batch_cross = b.cross || dev_tasks;  // SW pair/cell counts come from the device counters
	return RSK_OK;
}

int finish_batch_timing(rsk_ctx *ctx)
{
	float ms = 0;
	if (ctx->batch_filtered) {
		rsk_ctx::Counters c;
		CK(cudaMemcpy(&c, ctx->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
		if (ctx->batch_cross) {
			ctx->stats.sw_pairs += c.pair_count;
			ctx->stats.sw_cells += c.cell_count;
		} else {
			ctx->stats.sw_pairs += ctx->filt_explicit_pairs;
			ctx->stats.sw_cells += ctx->filt_explicit_cells;
		}
		ctx->stats.mu_saturated += c.sat_count;
		CK(cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[0]));
		ctx->stats.mu_kernel_ms += ms;
	}
	CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
	ctx->stats.sw_kernel_ms += ms;
	CK(cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[2]));
	ctx->stats.lddt_kernel_ms += ms;
	CK(cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[4]));
	ctx->stats.mkf_kernel_ms += ms;
	return RSK_OK;
}

void fill_hit(const rsk_params &P, const PairRec &r, uint32_t a, uint32_t b, uint64_t pool_base, rsk_hit &h)
{
	h.a = a; h.b = b;
	h.score = r.score;
	h.lo_a = r.lo_a; h.lo_b = r.lo_b;
	h.mu_fwd = r.mu_fwd; h.mu_rev = r.mu_rev;
	h.mu_score = (r.mu_fwd != 0 && !((float)r.mu_fwd < P.omega_fwd)) ? (float)r.mu_fwd - (float)r.mu_rev : 0.0f;
	h.flags = r.flags;
	h.path_len = r.path_len;
	h.path_off = pool_base + r.path_off;
	if (r.flags & RSK_HIT_HAS_EVALUE) {
		h.hi_a = r.hi_a; h.hi_b = r.hi_b; h.ids = r.ids; h.gaps = r.gaps;
		h.lddt = r.lddt; h.ts = r.ts;
		const double P = rsk_pvalue(r.ts);    // dssaligner.cpp:891-893: (float) of the double results
		h.pvalue = (float)P;
		h.qual = (float)rsk_qual(r.ts);
		h.evalue = (float)(P * 8340);         // statsig.cpp: E = P * DBSize, the same double product as rsk_evalue()
	} else {
		// ClearAlign values (dssaligner.cpp:906-927)
		h.hi_a = h.hi_b = h.ids = h.gaps = 0xffffffffu;
		h.lddt = 0; h.ts = -FLT_MAX;
		h.pvalue = h.evalue = h.qual = FLT_MAX;
		if (r.path_len == 0) { h.lo_a = h.lo_b = 0xffffffffu; }
	}
	// runquery.cpp:72-73 + DBSearcher::Reject (dbsearcher.cpp:258-265)
	if (r.path_len > 0 && !((double)h.evalue > P.max_evalue))
		h.flags |= RSK_HIT_REPORTED;
}

// Explicit pair lists: build and upload the Mu filter's tasks for the batch's sorted pairs [k0,k1) - runs of equal A, chunks
// of one task's column count - plus the pair index arrays the LDDT kernel and the hit sink read.
int upload_explicit_tasks(rsk_ctx *ctx, const SearchPlan &plan, Batch &b)
{
	cudaStream_t st = ctx->stream;
	rsk_stats &S = ctx->stats;
	static thread_local std::vector<uint32_t> t_a, t_begin, t_cnt, slots, r_a, r_begin, r_cnt;
	// (cudaMemcpyAsync from pageable memory returns once the source has been staged, so the vectors can be reused per batch)
	const bool mu16 = std::min(b.maxLA, b.maxLB) <= kMu16MaxLen && !getenv("RSK_MU32");
	// column chains per Mu task: the packed kernel takes 64 for row chains of <= 192 residues (half-warp wavefronts), 32 otherwise
	auto task_cols_of = [&](uint32_t a) -> size_t {
		return !mu16 ? (size_t)kSwWarps : plan.A->hlen[a] <= 192 ? (size_t)kMuTaskCols : (size_t)kMuTaskCols / 2;
	};
	t_a.clear(); t_begin.clear(); t_cnt.clear();
	const size_t n = b.k1 - b.k0;
	slots.resize(n);
	for (size_t k = 0; k < n; ++k)
		slots[k] = (uint32_t)k;
	size_t k = 0;
	while (k < n) {
		size_t e = k + 1;
		const size_t task_cols = task_cols_of(plan.sa[b.k0 + k]);
		while (e < n && e - k < task_cols && plan.sa[b.k0 + e] == plan.sa[b.k0 + k])
			++e;
		t_a.push_back(plan.sa[b.k0 + k]);
		t_begin.push_back((uint32_t)k);
		t_cnt.push_back((uint32_t)(e - k));
		k = e;
	}
	b.ntasks = (uint32_t)t_a.size();
	// runs of the same row chain (the Mu tasks above are pieces of them): what compact_explicit_kernel cuts into SW tasks
	r_a.clear(); r_begin.clear(); r_cnt.clear();
	uint64_t sw_tasks_max = 0;
	for (size_t t = 0; t < t_a.size();) {
		size_t e = t;
		uint32_t cnt = 0;
		while (e < t_a.size() && t_a[e] == t_a[t] && t_begin[e] == t_begin[t] + cnt)
			cnt += t_cnt[e++];
		r_a.push_back(t_a[t]); r_begin.push_back(t_begin[t]); r_cnt.push_back(cnt);
		sw_tasks_max += cnt / (16 * kSwChain) + 10;  // sw_task_chains: at most 9 small tasks per row chain, or full lists of >= 16 * kSwChain
		t = e;
	}
	b.nruns = (uint32_t)r_a.size();
	b.sw_task_cap = (uint32_t)std::min<uint64_t>(sw_tasks_max, 0xffffffffull / kSwClasses);
	if (ctx->run_a.ensure(b.nruns) || ctx->run_begin.ensure(b.nruns) || ctx->run_cnt.ensure(b.nruns))
		return fail(RSK_ERR_NOMEM, "task buffers");
	CK(cudaMemcpyAsync(ctx->run_a.p, r_a.data(), 4 * (size_t)b.nruns, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(ctx->run_begin.p, r_begin.data(), 4 * (size_t)b.nruns, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(ctx->run_cnt.p, r_cnt.data(), 4 * (size_t)b.nruns, cudaMemcpyHostToDevice, st));
	if (ctx->blist.ensure(n) || ctx->bslot.ensure(n) || ctx->pair_a.ensure(n) || ctx->pair_b.ensure(n) ||
		ctx->task_a.ensure(b.ntasks) || ctx->task_begin.ensure(b.ntasks) || ctx->task_cnt.ensure(b.ntasks)) {
		return fail(RSK_ERR_NOMEM, "task buffers");
	}
	CK(cudaMemcpyAsync(ctx->blist.p, plan.sb.data() + b.k0, 4 * n, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(ctx->bslot.p, slots.data(), 4 * n, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(ctx->pair_a.p, plan.sa.data() + b.k0, 4 * n, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(ctx->pair_b.p, plan.sb.data() + b.k0, 4 * n, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(ctx->task_a.p, t_a.data(), 4 * (size_t)b.ntasks, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(ctx->task_begin.p, t_begin.data(), 4 * (size_t)b.ntasks, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(ctx->task_cnt.p, t_cnt.data(), 4 * (size_t)b.ntasks, cudaMemcpyHostToDevice, st));
	S.h2d_bytes += 16 * n + 12 * (uint64_t)b.ntasks + 12 * (uint64_t)b.nruns;
	return RSK_OK;
}

struct SinkOpts { uint32_t a_base = 0, b_base = 0; bool append = false; };  // append: keep what earlier calls put into the sink
int run_to_sink(rsk_ctx *ctx, SearchPlan &plan, std::vector<Batch> &batches, const rsk_search_opts &opts, const SinkOpts &so);

int search_impl(rsk_ctx *ctx, SearchPlan &plan, const rsk_search_opts *opts_in, rsk_results **out, bool device_only, const SinkOpts *sink = nullptr)
{
	rsk_search_opts opts;
	memset(&opts, 0, sizeof(opts));
	if (opts_in)
		opts = *opts_in;
	const rsk_chainset *A = plan.A, *B = plan.B;
	if (A->ctx != ctx || B->ctx != ctx)
		return fail(RSK_ERR_ARG, "chain sets belong to a different context");
	CK(cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	rsk_stats &S = ctx->stats;
	const uint64_t launches0 = S.kernel_launches;
	memset(&S, 0, sizeof(S));
	S.kernel_launches = 0;
	(void)launches0;
	CK(cudaEventRecord(ctx->ev[6], st));

	// ---- batches ----
	std::vector<Batch> batches;
	if (plan.cross) {
		const uint32_t nA = A->d.n, nB = B->d.n;
		plan.npairs = (uint64_t)nA * nB;
		uint64_t sumLB = B->d.total;
		uint32_t maxLB = B->maxlen;
		const uint32_t per = (uint32_t)std::max<uint64_t>(1, ctx->max_batch_pairs / nB);
		// The host converts the records of a batch while the next one computes; nothing hides the conversion of the LAST
		// batch, so the tail of a multi-batch search is cut into a short final batch (a quarter of a full one).
		std::vector<uint32_t> cuts;
		for (uint32_t a0 = 0; a0 < nA; a0 += per)
			cuts.push_back(a0);
		if (cuts.size() >= 2 && !device_only) {
			const uint32_t last0 = cuts.back(), rows = nA - last0, tail = std::max(1u, per / 4);
			if (rows > tail)
				cuts.push_back(nA - tail);
		}
		cuts.push_back(nA);
		for (size_t ci = 0; ci + 1 < cuts.size(); ++ci) {
			const uint32_t a0 = cuts[ci];
			Batch b;
			b.cross = true;
			b.a0 = a0;
			b.a1 = cuts[ci + 1];
			b.npairs = (size_t)(b.a1 - b.a0) * nB;
			b.ntasks = (b.a1 - b.a0) * ((nB + kSwWarps - 1) / kSwWarps);
			uint64_t sumLA = 0;
			for (uint32_t a = b.a0; a < b.a1; ++a) {
				sumLA += A->hlen[a];
				b.maxLA = std::max(b.maxLA, A->hlen[a]);
			}
			b.maxLB = maxLB;
			b.cells = sumLA * sumLB;
			b.pool_bound = sumLA * nB + (uint64_t)(b.a1 - b.a0) * sumLB;
			batches.push_back(b);
		}
		// B order: by length so that the 16 chains of a task have similar lengths
		plan.border.resize(nB);
		std::iota(plan.border.begin(), plan.border.end(), 0u);
		std::stable_sort(plan.border.begin(), plan.border.end(),
				[&](uint32_t x, uint32_t y) { return B->hlen[x] > B->hlen[y]; });
		if (ctx->blist.ensure(nB))
			return fail(RSK_ERR_NOMEM, "blist");
		CK(cudaMemcpyAsync(ctx->blist.p, plan.border.data(), sizeof(uint32_t) * nB, cudaMemcpyHostToDevice, st));
		S.h2d_bytes += sizeof(uint32_t) * nB;
	} else {
		const uint64_t np = plan.npairs;
		for (size_t k0 = 0; k0 < np; ) {
			// a batch ends on a task boundary: cut at max_batch_pairs, then extend to the end of that A run
			size_t k1 = std::min<size_t>(np, k0 + ctx->max_batch_pairs);
			Batch b;
			b.cross = false;
			b.k0 = k0; b.k1 = k1;
			b.npairs = k1 - k0;
			{
				struct Part { uint32_t maxLA = 0, maxLB = 0; uint64_t cells = 0, pool = 0; char pad[40]; };
				std::vector<Part> parts((size_t)std::max(1, ctx->host_threads));
				parallel_ranges(k1 - k0, ctx->host_threads, [&](uint64_t lo, uint64_t hi, int t) {
					Part pt;
					for (uint64_t k = k0 + lo; k < k0 + hi; ++k) {
						const uint32_t la = A->hlen[plan.sa[k]], lb = B->hlen[plan.sb[k]];
						pt.maxLA = std::max(pt.maxLA, la);
						pt.maxLB = std::max(pt.maxLB, lb);
						pt.cells += (uint64_t)la * lb;
						pt.pool += (uint64_t)la + lb;
					}
					parts[t] = pt;
				});
				for (const Part &pt : parts) {
					b.maxLA = std::max(b.maxLA, pt.maxLA);
					b.maxLB = std::max(b.maxLB, pt.maxLB);
					b.cells += pt.cells;
					b.pool_bound += pt.pool;
				}
			}
			batches.push_back(b);
			k0 = k1;
		}
	}
	S.pairs = plan.npairs;

	if (sink)
		return run_to_sink(ctx, plan, batches, opts, *sink);

	rsk_results *res = nullptr;
	struct ResGuard {
		rsk_results *&r;
		~ResGuard() { delete r; r = nullptr; }
	} res_guard{res};  // `res` is handed to the caller by setting it to nullptr after copying the pointer out
	const bool keep_all = opts.keep == RSK_KEEP_ALL;
	if (!device_only) {
		res = new rsk_results();
		if (keep_all) {
			res->hits = (rsk_hit *)g_blocks.get(std::max<uint64_t>(1, plan.npairs) * sizeof(rsk_hit), res->hits_cap);
			if (!res->hits)
				return fail(RSK_ERR_NOMEM, "host memory for %llu hit records", (unsigned long long)plan.npairs);
			res->nhits = plan.npairs;
		}
	}

	// Host conversion jobs (PairRec -> rsk_hit incl. the libm P/E/Qual of statsig.cpp) run on worker threads while the
	// next batch occupies the GPU.  Path bytes go straight from the pinned staging pool to their final place; the hits of
	// RSK_KEEP_HITS are gathered per worker and appended to the result array when the job is collected (also overlapped,
	// except for the last batch).
	// Declared after `res_guard`, so that on ANY early return (CK, NOMEM) the workers are joined first and the result they
	// write into is deleted afterwards; a joinable std::thread must never reach its destructor.
	struct Job {
		std::vector<std::thread> threads;
		std::vector<std::vector<rsk_hit>> kept;  // per thread (KEEP_HITS)
		std::vector<uint64_t> n_eval, n_hit, n_rej;
		size_t npairs = 0;
		bool active = false;
		~Job()
		{
			for (auto &t : threads)
				if (t.joinable())
					t.join();
		}
	};
	Job jobs[2];
	const bool timing = getenv("RSK_TIMING") != nullptr;  // developer aid: host-side phases of the call on stderr
	double t_wait = 0, t_d2h = 0, t_launch = 0;
	auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	const double t_call0 = now();
	double keep_ratio = (ctx->params.omega > 0) ? 0.02 : 1.0;  // expected share of reported pairs, refined per batch
	uint64_t batches_left = batches.size();
	bool nomem = false;
	auto parallel_copy = [&](char *dst, const char *src, size_t bytes) {
		const int T = (int)std::max<size_t>(1, std::min<size_t>((size_t)ctx->host_threads, bytes >> 22));
		if (T == 1) {
			memcpy(dst, src, bytes);
			return;
		}
		std::vector<std::thread> th;
		for (int t = 0; t < T; ++t)
			th.emplace_back([=]() {
				const size_t b0 = bytes * (size_t)t / T, b1 = bytes * (size_t)(t + 1) / T;
				memcpy(dst + b0, src + b0, b1 - b0);
			});
		for (auto &t : th)
			t.join();
	};
	// grow-only result storage; growing moves the block, so no worker may be writing into it at that moment
	auto ensure_hits = [&](uint64_t need, uint64_t hint) -> bool {
		if (need * sizeof(rsk_hit) <= res->hits_cap)
			return true;
		size_t cap = 0;
		const uint64_t want = std::max<uint64_t>(std::max<uint64_t>(need, hint), 2 * (res->hits_cap / sizeof(rsk_hit)));
		rsk_hit *nh = (rsk_hit *)g_blocks.get(want * sizeof(rsk_hit), cap);
		if (!nh)
			return false;
		if (res->nhits)
			parallel_copy((char *)nh, (const char *)res->hits, res->nhits * sizeof(rsk_hit));
		g_blocks.put(res->hits, res->hits_cap);
		res->hits = nh;
		res->hits_cap = cap;
		return true;
	};
	auto wait_job = [&](Job &J) {
		if (!J.active)
			return;
		for (auto &t : J.threads)
			t.join();
		J.threads.clear();
		uint64_t kept_n = 0;
		for (size_t t = 0; t < J.n_eval.size(); ++t) {
			S.evalue_pairs += J.n_eval[t];
			S.hits += J.n_hit[t];
			S.mu_filter_rejected += J.n_rej[t];
		}
		for (auto &v : J.kept)
			kept_n += v.size();
		--batches_left;
		if (!keep_all) {
			if (J.npairs)
				keep_ratio = std::min(1.0, (double)kept_n / (double)J.npairs * 1.05 + 1e-3);
			// capacity hint: what is there plus this batch's yield for every batch still to come
			if (kept_n && !ensure_hits(res->nhits + kept_n, res->nhits + kept_n * (batches_left + 1) + 1024)) {
				nomem = true;
			} else if (kept_n) {
				std::vector<std::thread> th;
				uint64_t off = res->nhits;
				for (auto &v : J.kept) {
					if (v.empty())
						continue;
					rsk_hit *dst = res->hits + off;
					const std::vector<rsk_hit> *src = &v;
					off += v.size();
					if (v.size() < 16384)
						memcpy(dst, src->data(), src->size() * sizeof(rsk_hit));
					else
						th.emplace_back([dst, src]() { memcpy(dst, src->data(), src->size() * sizeof(rsk_hit)); });
				}
				for (auto &t : th)
					t.join();
				res->nhits = off;
			}
		}
		J.kept.clear();
		J.active = false;
	};
	auto abort_all = [&]() {  // explicit form of what the guards do on any other early return
		wait_job(jobs[0]);
		wait_job(jobs[1]);
	};

	uint64_t pool_total = 0;
	size_t bi = 0, collected = 0;
	// D2H of one finished batch (the batch set currently swapped in) and the launch of its host conversion job
	Batch pending;
	int pending_buf = 0;
	bool have_pending = false;
	auto collect = [&](const Batch &b, const int buf) -> int {
		int rc = RSK_OK;
		cudaStream_t cst = ctx->copy_stream;
		CK(cudaStreamWaitEvent(cst, ctx->done, 0));
		// the GPU is busy with this batch: now make sure the host buffer we are about to overwrite is free
		Job &J = jobs[buf];
		const double tw0 = now();
		wait_job(J);
		t_wait += now() - tw0;
		if (nomem) {
			abort_all();
			return fail(RSK_ERR_NOMEM, "host memory for the hit records");
		}
		// ---- D2H: records (+ paths) of this batch ----
		if (ctx->h_rec[buf].ensure(b.npairs)) {
			abort_all();
			return fail(RSK_ERR_NOMEM, "pinned record buffer");
		}
		unsigned long long pool_used = 0;
		const double td0 = now();
		CK(cudaMemcpyAsync(ctx->h_rec[buf].p, ctx->rec.p, b.npairs * sizeof(PairRec), cudaMemcpyDeviceToHost, cst));
		CK(cudaMemcpyAsync(&pool_used, ctx->d_pool_cursor, sizeof(pool_used), cudaMemcpyDeviceToHost, cst));
		CK(cudaStreamSynchronize(cst));
		S.d2h_bytes += b.npairs * sizeof(PairRec) + 8;
		const uint64_t pool_base = pool_total;
		const bool copy_paths = opts.want_paths && pool_used > 0;
		if (copy_paths) {
			if (ctx->h_pool[buf].ensure(pool_used)) {
				abort_all();
				return fail(RSK_ERR_NOMEM, "pinned path buffer");
			}
			CK(cudaMemcpyAsync(ctx->h_pool[buf].p, ctx->pool.p, pool_used, cudaMemcpyDeviceToHost, cst));
			CK(cudaStreamSynchronize(cst));
			S.d2h_bytes += pool_used;
			if (pool_total + pool_used > res->paths_cap) {
				// grow the path pool (rare: the first batch sizes it for the whole call); nobody may be writing into it
				wait_job(jobs[buf ^ 1]);
				size_t cap = 0;
				const uint64_t want = std::max<uint64_t>(pool_total + pool_used * (batches.size() - collected) + (pool_used >> 3) + 4096,
						2 * (uint64_t)res->paths_cap);
				char *np = (char *)g_blocks.get(want, cap);
				if (!np) {
					abort_all();
					return fail(RSK_ERR_NOMEM, "host memory for %llu path bytes", (unsigned long long)want);
				}
				if (pool_total)
					parallel_copy(np, res->paths, pool_total);
				g_blocks.put(res->paths, res->paths_cap);
				res->paths = np;
				res->paths_cap = cap;
			}
			pool_total += pool_used;
			res->npath = pool_total;
		}
		++collected;
		t_d2h += now() - td0;  // includes waiting for the batch's kernels
		rc = finish_batch_timing(ctx);
		if (rc) {
			abort_all();
			return rc;
		}
		// ---- records -> rsk_hit on worker threads (overlaps the next batch's kernels) ----
		const int T = (int)std::max<size_t>(1, std::min<size_t>((size_t)ctx->host_threads, b.npairs / 4096 + 1));
		J.active = true;
		J.npairs = b.npairs;
		J.kept.assign(T, {});
		J.n_eval.assign(T, 0); J.n_hit.assign(T, 0); J.n_rej.assign(T, 0);
		const PairRec *hrec = ctx->h_rec[buf].p;
		const uint8_t *hpool = ctx->h_pool[buf].p;
		char *path_dst = copy_paths ? res->paths + pool_base : nullptr;
		const uint64_t npath = copy_paths ? pool_used : 0;
		const rsk_params *P = &ctx->params;
		const uint32_t nB = B->d.n;
		rsk_hit *all = keep_all ? res->hits : nullptr;
		const SearchPlan *pl = &plan;
		const double ratio = keep_ratio;
		for (int t = 0; t < T; ++t) {
			J.threads.emplace_back([=, &J]() {
				const size_t k0 = b.npairs * (size_t)t / T, k1 = b.npairs * (size_t)(t + 1) / T;
				if (path_dst) {  // path bytes: each worker copies its share of the pinned pool to its final place
					const uint64_t p0 = npath * (uint64_t)t / T, p1 = npath * (uint64_t)(t + 1) / T;
					memcpy(path_dst + p0, hpool + p0, p1 - p0);
				}
				uint64_t ne = 0, nh = 0, nr = 0;
				std::vector<rsk_hit> &kept = J.kept[t];
				if (!all)
					kept.reserve((size_t)((double)(k1 - k0) * ratio) + 64);
				for (size_t k = k0; k < k1; ++k) {
					uint32_t a, bb;
					uint64_t orig;
					if (b.cross) {
						a = b.a0 + (uint32_t)(k / nB);
						bb = (uint32_t)(k % nB);
						orig = (uint64_t)a * nB + bb;
					} else {
						a = pl->sa[b.k0 + k];
						bb = pl->sb[b.k0 + k];
						orig = pl->perm[b.k0 + k];
					}
					rsk_hit h;
					fill_hit(*P, hrec[k], a, bb, pool_base, h);
					ne += (h.flags & RSK_HIT_HAS_EVALUE) != 0;
					nh += (h.flags & RSK_HIT_REPORTED) != 0;
					nr += (h.flags & RSK_HIT_MU_REJECTED) != 0;
					if (all)
						all[orig] = h;
					else if (h.flags & RSK_HIT_REPORTED)
						kept.push_back(h);
				}
				J.n_eval[t] = ne; J.n_hit[t] = nh; J.n_rej[t] = nr;
			});
		}
			return RSK_OK;
	};

	for (const Batch &b0 : batches) {
		Batch b = b0;
		const int buf = (int)(bi++ & 1);
		const double tt_0 = now_ms();
		if (!b.cross) {
			const int rce = upload_explicit_tasks(ctx, plan, b);
			if (rce)
				return rce;
		}
		g_t_tasks += now_ms() - tt_0;
		const double tl0 = now();
		int rc = run_batch(ctx, plan, b, opts);
		t_launch += now() - tl0;
		if (rc) {
			if (!device_only)
				abort_all();
			return rc;
		}
		if (device_only) {
			CK(cudaStreamSynchronize(st));
			rc = finish_batch_timing(ctx);
			if (rc)
				return rc;
			continue;
		}
		// the GPU is busy with this batch: collect the previous one (its device outputs live in the other batch set; the copies
		// run on copy_stream next to this batch's kernels), then make this batch's set the pending one
		CK(cudaEventRecord(ctx->done, st));
		ctx->swap_batch_set();
		if (have_pending) {
			rc = collect(pending, pending_buf);
			if (rc)
				return rc;
		}
		pending = b;
		pending_buf = buf;
		have_pending = true;
	}
	if (!device_only && have_pending) {
		ctx->swap_batch_set();
		int rc = collect(pending, pending_buf);
		ctx->swap_batch_set();
		if (rc)
			return rc;
	}
	if (!device_only) {
		// explicit-mode KEEP_HITS hits arrive in sorted-pair order; cross-mode hits in (a, b) order
		const double tt0 = now();
		wait_job(jobs[bi & 1]);
		wait_job(jobs[(bi + 1) & 1]);
		if (timing) {
			fprintf(stderr, "[rsk_search] host phases: task lists + H2D %.1f ms, waiting for the Mu filter (keep flags) %.1f ms, SW task lists + H2D %.1f ms, plan %.1f ms\n",
					g_t_tasks, g_t_keepwait, g_t_explicit, g_t_plan);
			g_t_tasks = g_t_keepwait = g_t_explicit = g_t_plan = 0;
		}
		if (timing)
			fprintf(stderr, "[rsk_search] %zu batches: launch %.1f ms, kernels+D2H %.1f ms, waiting for host conversion %.1f ms, tail %.1f ms, call %.1f ms\n",
					batches.size(), t_launch, t_d2h, t_wait, now() - tt0, now() - t_call0);
		if (nomem)
			return fail(RSK_ERR_NOMEM, "host memory for the hit records");
	}
	CK(cudaEventRecord(ctx->ev[7], st));
	CK(cudaEventSynchronize(ctx->ev[7]));
	CK(cudaEventElapsedTime(&S.total_ms, ctx->ev[6], ctx->ev[7]));
	if (out)
		*out = res;
	res = nullptr;  // ownership passed to the caller (or there was no result object)
	return RSK_OK;
}


// ------------------------------------------------------------------------------------------------
// device hit sink: search -> compacted hits on the device -> (gather over NVLink) -> host
// ------------------------------------------------------------------------------------------------
// Emit threshold on the test statistic.  E(ts) = (float)(P(ts) * 8340) (statsig.cpp:27-50, dssaligner.cpp:891-893) does not
// increase with ts, so "E <= MaxEvalue" is "ts >= t*"; t* is found by bisection over the ordered float bit patterns with the
// very functions the host applies afterwards, then lowered by a margin far above libm's error.  The device keeps a superset,
// the exact test runs on what arrives.
float sink_ts_threshold(double max_evalue)
{
	auto rejected = [&](float ts) { return (double)(float)(rsk_pvalue(ts) * 8340) > max_evalue; };
	if (!rejected(-FLT_MAX))
		return -FLT_MAX;
	if (rejected(FLT_MAX))
		return FLT_MAX;
	auto f2o = [](float f) { int32_t i; memcpy(&i, &f, 4); return i >= 0 ? (int64_t)i : -(int64_t)(i & 0x7fffffff); };
	auto o2f = [](int64_t o) { int32_t i = o >= 0 ? (int32_t)o : (int32_t)(0x80000000u | (uint32_t)(-o)); float f; memcpy(&f, &i, 4); return f; };
	int64_t lo = f2o(-FLT_MAX), hi = f2o(FLT_MAX);  // lo rejected, hi accepted
	while (hi - lo > 1) {
		const int64_t mid = lo + (hi - lo) / 2;
		if (rejected(o2f(mid)))
			lo = mid;
		else
			hi = mid;
	}
	const float t = o2f(hi);
	return t - fabsf(t) * 1e-5f - 1e-6f;
}

// capacity for `need_rec` records / `need_pool` path bytes; growing keeps what the finished batches wrote
int sink_reserve(rsk_ctx *ctx, size_t need_rec, size_t need_pool)
{
	HitSink &K = ctx->sink;
	if (need_rec <= K.rec_cap && need_pool <= K.pool_cap)
		return RSK_OK;
	cudaStream_t st = ctx->stream;
	CK(cudaStreamSynchronize(st));
	unsigned long long tot[4] = {0, 0, 0, 0};
	CK(cudaMemcpy(tot, K.d_tot, sizeof(tot), cudaMemcpyDeviceToHost));
	if (need_rec > K.rec_cap) {
		const size_t want = std::max(need_rec, 2 * K.rec_cap) + 1024;
		SinkRec *p = nullptr;
		if (cudaMalloc((void **)&p, want * sizeof(SinkRec)) != cudaSuccess) {
			cudaGetLastError();
			return fail(RSK_ERR_NOMEM, "device hit sink: %zu records", want);
		}
		if (K.rec && tot[0])
			CK(cudaMemcpy(p, K.rec, tot[0] * sizeof(SinkRec), cudaMemcpyDeviceToDevice));
		if (K.rec)
			cudaFree(K.rec);
		K.rec = p;
		K.rec_cap = want;
	}
	if (need_pool > K.pool_cap) {
		const size_t want = std::max(need_pool, 2 * K.pool_cap) + 4096;
		uint8_t *p = nullptr;
		if (cudaMalloc((void **)&p, want) != cudaSuccess) {
			cudaGetLastError();
			return fail(RSK_ERR_NOMEM, "device hit sink: %zu path bytes", want);
		}
		if (K.pool && tot[1])
			CK(cudaMemcpy(p, K.pool, tot[1], cudaMemcpyDeviceToDevice));
		if (K.pool)
			cudaFree(K.pool);
		K.pool = p;
		K.pool_cap = want;
	}
	return RSK_OK;
}

// The batches of a search call, their emitted records compacted into ctx->sink.  No D2H of per-pair records and no host
// conversion per batch: the host only runs two batches ahead of the device (it needs the exact sink totals to bound the space
// the next batch may take).  On return the stream is idle and ctx->sink.h_tot[0..3] holds the totals.
int run_to_sink(rsk_ctx *ctx, SearchPlan &plan, std::vector<Batch> &batches, const rsk_search_opts &opts, const SinkOpts &so)
{
	HitSink &K = ctx->sink;
	cudaStream_t st = ctx->stream;
	rsk_stats &S = ctx->stats;
	if (!K.d_tot) {
		CK(cudaMalloc((void **)&K.d_tot, 4 * sizeof(unsigned long long)));
		CK(cudaHostAlloc((void **)&K.h_tot, 8 * sizeof(unsigned long long), cudaHostAllocDefault));
	}
	if (!so.append)
		CK(cudaMemsetAsync(K.d_tot, 0, 4 * sizeof(unsigned long long), st));
	const float ts_lo = sink_ts_threshold(ctx->params.max_evalue);
	const uint32_t report_no_evalue = !((double)FLT_MAX > ctx->params.max_evalue) ? 1u : 0u;
	// exact totals so far (an appending call starts from what the previous call left; every call ends synchronised)
	unsigned long long conf_rec = so.append ? K.h_tot[0] : 0, conf_pool = so.append ? K.h_tot[1] : 0;
	size_t unconf_rec[2] = {0, 0}, unconf_pool[2] = {0, 0};
	const size_t nb = batches.size();
	int rc;
	for (size_t i = 0; i < nb; ++i) {
		Batch b = batches[i];
		const int set = (int)(i & 1);
		if (i >= 2) {  // batch i-2 ran in this batch set: its timing events and the exact totals after it
			CK(cudaEventSynchronize(ctx->done));
			if ((rc = finish_batch_timing(ctx)))
				return rc;
			conf_rec = K.h_tot[set * 4 + 0];
			conf_pool = K.h_tot[set * 4 + 1];
			unconf_rec[set] = unconf_pool[set] = 0;
		}
		if (!b.cross && (rc = upload_explicit_tasks(ctx, plan, b)))
			return rc;
		unconf_rec[set] = b.npairs;
		unconf_pool[set] = opts.want_paths ? (size_t)b.pool_bound : 0;
		if ((rc = sink_reserve(ctx, conf_rec + unconf_rec[0] + unconf_rec[1], conf_pool + unconf_pool[0] + unconf_pool[1] + 64)))
			return rc;
		if ((rc = run_batch(ctx, plan, b, opts)))
			return rc;
		if (b.npairs > 0x7fffffffull)
			return fail(RSK_ERR_LIMIT, "hit sink: batch of %zu pairs", b.npairs);
		const size_t tmp_bytes = sink_scan_tmp_bytes((uint32_t)b.npairs);
		if (K.keep.ensure(b.npairs) || K.plen.ensure(b.npairs) || K.keep_scan.ensure(b.npairs) || K.plen_scan.ensure(b.npairs) ||
			K.tmp.ensure(tmp_bytes)) {
			cudaGetLastError();
			return fail(RSK_ERR_NOMEM, "hit sink scratch for %zu pairs", b.npairs);
		}
		SinkArgs a;
		memset(&a, 0, sizeof(a));
		a.rec = ctx->rec.p; a.npairs = (uint32_t)b.npairs; a.pool = ctx->pool.p;
		a.cross = b.cross ? 1 : 0; a.a_begin = b.a0; a.nB = plan.B->d.n;
		a.pair_a = ctx->pair_a.p; a.pair_b = ctx->pair_b.p;
		a.a_base = so.a_base; a.b_base = so.b_base;
		a.keep_all = opts.keep == RSK_KEEP_ALL ? 1 : 0;
		a.want_paths = opts.want_paths ? 1 : 0;
		a.report_no_evalue = report_no_evalue;
		a.ts_lo = ts_lo;
		a.keep = K.keep.p; a.plen = K.plen.p; a.keep_scan = K.keep_scan.p; a.plen_scan = K.plen_scan.p;
		a.out_rec = K.rec; a.out_pool = K.pool; a.totals = K.d_tot;
		const int nl = launch_sink_append(a, K.tmp.p, tmp_bytes, st);
		if (nl < 0)
			return fail(RSK_ERR_CUDA, "hit sink kernels failed to launch: %s", cudaGetErrorString(cudaGetLastError()));
		S.kernel_launches += nl;
		CK(cudaMemcpyAsync(K.h_tot + set * 4, K.d_tot, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
		CK(cudaEventRecord(ctx->done, st));
		ctx->swap_batch_set();
	}
	if (nb >= 2) {  // batch nb-2 lives in the current set
		CK(cudaEventSynchronize(ctx->done));
		if ((rc = finish_batch_timing(ctx)))
			return rc;
	}
	ctx->swap_batch_set();
	if (nb >= 1) {
		CK(cudaEventSynchronize(ctx->done));
		if ((rc = finish_batch_timing(ctx)))
			return rc;
	}
	CK(cudaMemcpyAsync(K.h_tot, K.d_tot, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
	CK(cudaEventRecord(ctx->ev[7], st));
	CK(cudaEventSynchronize(ctx->ev[7]));
	CK(cudaEventElapsedTime(&S.total_ms, ctx->ev[6], ctx->ev[7]));
	S.evalue_pairs = K.h_tot[2];
	S.mu_filter_rejected = K.h_tot[3];
	return RSK_OK;
}

// Device -> host through the double-buffered pinned path staging of the context, the memcpy out of one buffer overlapping
// the DMA into the other.
int d2h_staged(rsk_ctx *ctx, char *dst, const unsigned char *src, size_t bytes)
{
	if (bytes == 0)
		return RSK_OK;
	const size_t chunk = (size_t)64 << 20;
	cudaStream_t cst = ctx->copy_stream;
	for (int k = 0; k < 2; ++k)
		if (ctx->h_pool[k].ensure(std::min(bytes, chunk)))
			return fail(RSK_ERR_NOMEM, "pinned staging buffer");
	const size_t nchunks = (bytes + chunk - 1) / chunk;
	std::thread copier;
	struct Joiner {
		std::thread &t;
		~Joiner() { if (t.joinable()) t.join(); }
	} joiner{copier};
	for (size_t c = 0; c < nchunks; ++c) {
		const size_t off = c * chunk, n = std::min(chunk, bytes - off);
		const int buf = (int)(c & 1);
		// the memcpy out of this buffer (chunk c-2) was joined one iteration ago; chunk c-1's runs next to this DMA
		CK(cudaMemcpyAsync(ctx->h_pool[buf].p, src + off, n, cudaMemcpyDeviceToHost, cst));
		CK(cudaStreamSynchronize(cst));
		if (copier.joinable())
			copier.join();
		const uint8_t *hp = ctx->h_pool[buf].p;
		char *d = dst + off;
		copier = std::thread([hp, d, n]() { memcpy(d, hp, n); });
	}
	if (copier.joinable())
		copier.join();
	ctx->stats.d2h_bytes += bytes;
	return RSK_OK;
}

// Read `nrec` sink records (device) and their path pool out to a result object: chunked D2H into pinned staging, conversion
// to rsk_hit (libm P/E/Qual + the exact emit test) on worker threads while the next chunk is in flight.  pool_base[k] is
// added to the path offsets of the records of segment k (segments = the ranks' blocks inside a gathered array).
int sink_to_results(rsk_ctx *ctx, const SinkRec *d_rec, uint64_t nrec, const unsigned char *d_pool, uint64_t npool,
		const std::vector<uint64_t> &seg_end, const std::vector<uint64_t> &pool_base, const rsk_search_opts &opts, rsk_results **out)
{
	rsk_results *res = new rsk_results();
	struct ResGuard {
		rsk_results *&r;
		~ResGuard() { delete r; }
	} guard{res};
	const bool keep_all = opts.keep == RSK_KEEP_ALL;
	res->hits = (rsk_hit *)g_blocks.get(std::max<uint64_t>(1, nrec) * sizeof(rsk_hit), res->hits_cap);
	if (!res->hits)
		return fail(RSK_ERR_NOMEM, "host memory for %llu hit records", (unsigned long long)nrec);
	if (opts.want_paths && npool) {
		res->paths = (char *)g_blocks.get(npool, res->paths_cap);
		if (!res->paths)
			return fail(RSK_ERR_NOMEM, "host memory for %llu path bytes", (unsigned long long)npool);
	}
	const size_t chunk = (size_t)1 << 20;  // records per chunk (72 MB)
	cudaStream_t cst = ctx->copy_stream;
	std::vector<std::thread> workers;
	struct Joiner {
		std::vector<std::thread> &t;
		~Joiner() { for (auto &x : t) if (x.joinable()) x.join(); }
	} joiner{workers};
	std::atomic<uint64_t> dropped{0};
	const rsk_params *P = &ctx->params;
	const size_t nchunks = (nrec + chunk - 1) / chunk;
	for (size_t c = 0; c < nchunks; ++c) {
		const uint64_t k0 = c * chunk, n = std::min<uint64_t>(chunk, nrec - k0);
		const int buf = (int)(c & 1);
		if (ctx->h_sink[buf].ensure(std::min<uint64_t>(chunk, nrec)))
			return fail(RSK_ERR_NOMEM, "pinned record staging");
		// the workers of chunk c-1 read the other buffer; those of chunk c-2 (this buffer) were joined one iteration ago
		CK(cudaMemcpyAsync(ctx->h_sink[buf].p, d_rec + k0, n * sizeof(SinkRec), cudaMemcpyDeviceToHost, cst));
		CK(cudaStreamSynchronize(cst));
		for (auto &t : workers)
			t.join();
		workers.clear();
		const SinkRec *hr = ctx->h_sink[buf].p;
		rsk_hit *hits = res->hits;
		const int T = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)ctx->host_threads, n / 8192 + 1));
		for (int t = 0; t < T; ++t)
			workers.emplace_back([=, &dropped, &seg_end, &pool_base]() {
				const uint64_t i0 = n * (uint64_t)t / T, i1 = n * (uint64_t)(t + 1) / T;
				size_t seg = 0;
				uint64_t nd = 0;
				for (uint64_t i = i0; i < i1; ++i) {
					const uint64_t k = k0 + i;
					while (seg + 1 < seg_end.size() && k >= seg_end[seg])
						++seg;
					rsk_hit h;
					fill_hit(*P, hr[i].r, hr[i].a, hr[i].b, pool_base[seg], h);
					if (!keep_all && !(h.flags & RSK_HIT_REPORTED)) {
						h.path_len = 0xffffffffu;  // marks a record the exact emit test drops (compacted away below)
						++nd;
					}
					hits[k] = h;
				}
				if (nd)
					dropped += nd;
			});
	}
	for (auto &t : workers)
		t.join();
	workers.clear();
	ctx->stats.d2h_bytes += nrec * sizeof(SinkRec);
	res->nhits = nrec;
	if (dropped.load()) {
		uint64_t w = 0;
		for (uint64_t k = 0; k < nrec; ++k)
			if (res->hits[k].path_len != 0xffffffffu)
				res->hits[w++] = res->hits[k];
		res->nhits = w;
	}
	if (opts.want_paths && npool) {
		int rc = d2h_staged(ctx, res->paths, d_pool, npool);
		if (rc)
			return rc;
		res->npath = npool;
	}
	*out = res;
	res = nullptr;
	return RSK_OK;
}

uint64_t mix64(uint64_t x)
{
	x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
	x ^= x >> 27; x *= 0x94d049bb133111ebull;
	x ^= x >> 31;
	return x;
}
}  // namespace

extern "C" int rsk_search_cross(rsk_ctx *ctx, const rsk_chainset *A, const rsk_chainset *B, const rsk_search_opts *opts, rsk_results **out)
{
	if (!ctx || !A || !B || !out)
		return fail(RSK_ERR_ARG, "rsk_search_cross: null argument");
	*out = nullptr;
	SearchPlan plan;
	plan.A = A; plan.B = B; plan.cross = true;
	return search_impl(ctx, plan, opts, out, false);
}

extern "C" int rsk_search_cross_device(rsk_ctx *ctx, const rsk_chainset *A, const rsk_chainset *B, const rsk_search_opts *opts)
{
	if (!ctx || !A || !B)
		return fail(RSK_ERR_ARG, "rsk_search_cross_device: null argument");
	SearchPlan plan;
	plan.A = A; plan.B = B; plan.cross = true;
	return search_impl(ctx, plan, opts, nullptr, true);
}

static int build_explicit_plan(SearchPlan &plan, const rsk_chainset *A, const rsk_chainset *B, uint64_t npairs,
		const uint32_t *ia, const uint32_t *ib, int nthreads, uint32_t mkfl);

static int sink_reset_empty(rsk_ctx *ctx)
{
	memset(&ctx->stats, 0, sizeof(ctx->stats));
	CK(cudaSetDevice(ctx->device));
	if (!ctx->sink.h_tot)
		CK(cudaHostAlloc((void **)&ctx->sink.h_tot, 8 * sizeof(unsigned long long), cudaHostAllocDefault));
	memset(ctx->sink.h_tot, 0, 8 * sizeof(unsigned long long));
	return RSK_OK;
}

// Collect the context's hit sink: gather over the communicator (if any) and read out on the root.
static int sink_finish(rsk_ctx *ctx, rsk_comm *comm, int root, const rsk_search_opts &opts, rsk_results **out)
{
	HitSink &K = ctx->sink;
	const uint64_t nrec = K.h_tot ? K.h_tot[0] : 0, npool = (K.h_tot && opts.want_paths) ? K.h_tot[1] : 0;
	const int N = comm_nranks(comm), me = comm_rank(comm);
	if (N == 1) {
		std::vector<uint64_t> seg_end{nrec}, pool_base{0};
		int rc = sink_to_results(ctx, K.rec, nrec, K.pool, npool, seg_end, pool_base, opts, out);
		if (!rc)
			ctx->stats.hits = (*out)->nhits;
		return rc;
	}
	const unsigned long long mine[kCommCountWords] = {nrec * sizeof(SinkRec), npool, 0, 0};
	const unsigned long long *all = nullptr;
	int rc = comm_exchange_counts(comm, mine, &all);
	if (rc)
		return rc;
	std::vector<unsigned long long> bytes((size_t)N * 2);
	for (int r = 0; r < N; ++r) {
		bytes[(size_t)r * 2 + 0] = all[(size_t)r * kCommCountWords + 0];
		bytes[(size_t)r * 2 + 1] = all[(size_t)r * kCommCountWords + 1];
	}
	const void *src[2] = {K.rec, K.pool};
	unsigned char *g[2] = {nullptr, nullptr};
	unsigned long long gtot[2] = {0, 0};
	if ((rc = comm_gather_parts(comm, root, 2, src, bytes.data(), g, gtot)))
		return rc;
	ctx->stats.hits = nrec;  // this rank's device-compacted records (the exact count is the root's)
	if (me != root) {
		*out = nullptr;
		return RSK_OK;
	}
	std::vector<uint64_t> seg_end(N), pool_base(N);
	uint64_t racc = 0, pacc = 0;
	for (int r = 0; r < N; ++r) {
		pool_base[r] = pacc;
		racc += bytes[(size_t)r * 2 + 0] / sizeof(SinkRec);
		pacc += bytes[(size_t)r * 2 + 1];
		seg_end[r] = racc;
	}
	rc = sink_to_results(ctx, (const SinkRec *)g[0], racc, g[1], pacc, seg_end, pool_base, opts, out);
	if (!rc)
		ctx->stats.hits = (*out)->nhits;
	return rc;
}

extern "C" int rsk_search_cross_sharded(rsk_ctx *ctx, rsk_comm *comm, const rsk_chainset *A_local, const rsk_chainset *B, uint32_t a_base,
		const rsk_search_opts *opts_in, int root, rsk_results **out)
{
	if (!ctx || !B || !out)
		return fail(RSK_ERR_ARG, "rsk_search_cross_sharded: null argument");
	*out = nullptr;
	if (comm && comm_ctx(comm) != ctx)
		return fail(RSK_ERR_ARG, "rsk_search_cross_sharded: the communicator belongs to a different context");
	if (root < 0 || root >= comm_nranks(comm))
		return fail(RSK_ERR_ARG, "rsk_search_cross_sharded: root %d of %d ranks", root, comm_nranks(comm));
	rsk_search_opts opts;
	memset(&opts, 0, sizeof(opts));
	if (opts_in)
		opts = *opts_in;
	if (A_local && A_local->d.n) {
		SearchPlan plan;
		plan.A = A_local; plan.B = B; plan.cross = true;
		SinkOpts so;
		so.a_base = a_base;
		int rc = search_impl(ctx, plan, &opts, nullptr, false, &so);
		if (rc)
			return rc;  // NB: the other ranks will wait in the gather; a failing rank is fatal for the job, as a Die() is
	} else {
		// an empty block (more ranks than chains): nothing to search, but the rank still takes part in the collective
		int rc = sink_reset_empty(ctx);
		if (rc)
			return rc;
	}
	return sink_finish(ctx, comm, root, opts, out);
}

// Explicit pair list through the hit sink and the gather: PostMuFilter's scan over this rank's share of the candidate list
// (postmufilter.cpp:185-195).  Hits arrive rank by rank, inside a rank in schedule order (A-major).
int rsk::search_pairs_sharded(rsk_ctx *ctx, rsk_comm *comm, const rsk_chainset *A, const rsk_chainset *B, uint64_t npairs,
		const uint32_t *ia, const uint32_t *ib, uint32_t a_base, uint32_t b_base, const rsk_search_opts *opts_in, int root, rsk_results **out)
{
	*out = nullptr;
	rsk_search_opts opts;
	memset(&opts, 0, sizeof(opts));
	if (opts_in)
		opts = *opts_in;
	int rc;
	if (npairs && A && B) {
		SearchPlan plan;
		if ((rc = build_explicit_plan(plan, A, B, npairs, ia, ib, ctx->host_threads, ctx->params.mkfl)))
			return rc;
		SinkOpts so;
		so.a_base = a_base; so.b_base = b_base;
		if ((rc = search_impl(ctx, plan, &opts, nullptr, false, &so)))
			return rc;
	} else if ((rc = sink_reset_empty(ctx))) {
		return rc;
	}
	return sink_finish(ctx, comm, root, opts, out);
}

// Order-independent digest: the sum over hits of a hash of the record fields and the path bytes.
extern "C" uint64_t rsk_results_digest(const rsk_results *r)
{
	if (!r || !r->nhits)
		return 0;
	const unsigned T = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(std::min(32u, std::max(1u, std::thread::hardware_concurrency())), r->nhits >> 14));
	std::vector<uint64_t> part(T, 0);
	auto work = [&](unsigned t) {
		uint64_t acc = 0;
		const uint64_t k0 = r->nhits * t / T, k1 = r->nhits * (uint64_t)(t + 1) / T;
		for (uint64_t k = k0; k < k1; ++k) {
			const rsk_hit &h = r->hits[k];
			uint32_t w[19];
			memcpy(w, &h, sizeof(uint32_t) * 19);  // every field up to and including path_len (path_off depends on the layout)
			uint64_t x = 0x9e3779b97f4a7c15ull;
			for (int i = 0; i < 19; ++i)
				x = mix64(x ^ w[i]) + 0x9e3779b97f4a7c15ull * (uint64_t)(i + 1);
			if (r->paths && h.path_len && h.path_off + h.path_len <= r->npath) {
				const unsigned char *p = (const unsigned char *)r->paths + h.path_off;
				uint64_t y = 1469598103934665603ull;
				for (uint32_t i = 0; i < h.path_len; ++i)
					y = (y ^ p[i]) * 1099511628211ull;
				x = mix64(x ^ y);
			}
			acc += x;
		}
		part[t] = acc;
	};
	if (T == 1) {
		work(0);
	} else {
		std::vector<std::thread> th;
		for (unsigned t = 0; t < T; ++t)
			th.emplace_back(work, t);
		for (auto &t : th)
			t.join();
	}
	uint64_t d = 0;
	for (uint64_t v : part)
		d += v;
	return d;
}

// Explicit pair lists are scheduled a-major with the longest B first inside every run of equal A (the chains of one task
// then have similar lengths).  perm[k] = caller's index of the k-th scheduled pair.  A counting sort by A followed by an
// independent (stable) sort of every run by B length, runs spread over the host threads: an all-vs-all of 11 k chains is
// 6.3e7 pairs, for which one comparison sort over the whole list took longer than all kernels together.
//
// Long-chain pairs (DoMKF: a chain >= mkfl, Mu letters on both sides) are scheduled after all other pairs, again a-major:
// their banded x-drop DPs are sequential per pair, so a batch must hold as many of them as possible to keep the GPU
// busy behind the longest one; sprinkled over all batches (3 % of each) every batch would wait for its own straggler.
static int build_explicit_plan(SearchPlan &plan, const rsk_chainset *A, const rsk_chainset *B, uint64_t npairs,
		const uint32_t *ia, const uint32_t *ib, int nthreads, uint32_t mkfl)
{
	plan.A = A; plan.B = B; plan.cross = false; plan.npairs = npairs;
	const uint32_t nA = A->d.n, nB = B->d.n;
	const bool mkf_possible = A->has_mu && B->has_mu;
	auto is_mkf = [&](uint32_t a, uint32_t b) {
		const uint32_t la = A->hlen[a], lb = B->hlen[b];
		return mkf_possible && la >= 3 && lb >= 3 && (la >= mkfl || lb >= mkfl);
	};
	// runs: key a for ordinary pairs, nA + a for long-chain pairs
	const size_t nruns = (size_t)2 * nA;
	std::vector<uint64_t> start(nruns + 1, 0);
	for (uint64_t k = 0; k < npairs; ++k) {
		if (ia[k] >= nA || ib[k] >= nB)
			return fail(RSK_ERR_ARG, "pair %llu: chain index out of range (%u,%u)", (unsigned long long)k, ia[k], ib[k]);
		++start[(is_mkf(ia[k], ib[k]) ? (size_t)nA : 0) + ia[k] + 1];
	}
	for (size_t r = 0; r < nruns; ++r)
		start[r + 1] += start[r];
	plan.perm.resize(npairs);
	{
		std::vector<uint64_t> cur(start.begin(), start.end() - 1);
		for (uint64_t k = 0; k < npairs; ++k)
			plan.perm[cur[(is_mkf(ia[k], ib[k]) ? (size_t)nA : 0) + ia[k]]++] = k;  // stable: caller order inside a run
	}
	plan.sa.resize(npairs);
	plan.sb.resize(npairs);
	const uint32_t maxlen = B->maxlen;
	const int T = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)std::max(1, nthreads), npairs >> 16));
	std::atomic<uint32_t> next_a{0};
	auto work = [&]() {
		std::vector<uint32_t> bucket;
		std::vector<uint64_t> tmp;
		for (;;) {
			const uint32_t r0 = next_a.fetch_add(64);
			if (r0 >= nruns)
				break;
			for (uint32_t r = r0; r < std::min<size_t>(nruns, (size_t)r0 + 64); ++r) {
				const uint32_t a = r >= nA ? r - nA : r;
				const uint64_t lo = start[r], hi = start[r + 1], n = hi - lo;
				if (n == 0)
					continue;
				uint64_t *pp = plan.perm.data() + lo;
				if (n > 1) {
					if (n >= 4096 && n >= (uint64_t)maxlen / 4) {
						// stable counting sort by B length, descending
						bucket.assign((size_t)maxlen + 2, 0);
						for (uint64_t k = 0; k < n; ++k)
							++bucket[maxlen - B->hlen[ib[pp[k]]] + 1];
						for (uint32_t l = 0; l <= maxlen; ++l)
							bucket[l + 1] += bucket[l];
						tmp.resize(n);
						for (uint64_t k = 0; k < n; ++k)
							tmp[bucket[maxlen - B->hlen[ib[pp[k]]]]++] = pp[k];
						memcpy(pp, tmp.data(), n * sizeof(uint64_t));
					} else {
						std::stable_sort(pp, pp + n, [&](uint64_t x, uint64_t y) { return B->hlen[ib[x]] > B->hlen[ib[y]]; });
					}
				}
				for (uint64_t k = 0; k < n; ++k) {
					plan.sa[lo + k] = a;
					plan.sb[lo + k] = ib[pp[k]];
				}
			}
		}
	};
	if (T == 1) {
		work();
	} else {
		std::vector<std::thread> th;
		for (int t = 0; t < T; ++t)
			th.emplace_back(work);
		for (auto &t : th)
			t.join();
	}
	return RSK_OK;
}

// The plan of RunSelf's pair triangle for rows i0, i0 + step, ... < i1 (pairs (i, j >= i) in that enumeration order), built
// directly: the same layout build_explicit_plan gives for that list - ordinary runs per row, then the long-chain runs per row,
// every run ordered by the partner's length (longest first, ties by index) - without materialising the 2 x 4 B/pair index
// lists and without the counting/permutation passes over them (1.2 s of single-threaded work for SCOP40's 6.3e7 pairs).
// A run is the filtered copy of ONE length-sorted list of all chains, so rows are independent and go to the host threads.
static int build_self_plan(SearchPlan &plan, const rsk_chainset *S, uint64_t i0, uint64_t i1, uint64_t step, int nthreads, uint32_t mkfl)
{
	plan.A = S; plan.B = S; plan.cross = false;
	const uint32_t n = S->d.n;
	const std::vector<uint32_t> &len = S->hlen;
	std::vector<uint32_t> order(n);
	std::iota(order.begin(), order.end(), 0u);
	std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return len[x] > len[y]; });
	const bool mkf_possible = S->has_mu;
	auto is_mkf = [&](uint32_t a, uint32_t b) {
		const uint32_t la = len[a], lb = len[b];
		return mkf_possible && la >= 3 && lb >= 3 && (la >= mkfl || lb >= mkfl);
	};
	// partners j >= i that make a long-chain pair with i: all of length >= 3 if i itself is long, else the long ones
	std::vector<uint32_t> ge3(n + 1, 0), gel(n + 1, 0);
	for (uint32_t j = n; j-- > 0;) {
		ge3[j] = ge3[j + 1] + (len[j] >= 3);
		gel[j] = gel[j + 1] + (len[j] >= 3 && len[j] >= mkfl);
	}
	std::vector<uint64_t> rows;
	for (uint64_t i = i0; i < i1; i += step)
		rows.push_back(i);
	const size_t nr = rows.size();
	std::vector<uint64_t> ord_start(nr + 1, 0), mkf_start(nr + 1, 0), enum_start(nr + 1, 0);
	for (size_t r = 0; r < nr; ++r) {
		const uint32_t a = (uint32_t)rows[r];
		const uint64_t all = n - a;
		const uint64_t m = !(mkf_possible && len[a] >= 3) ? 0 : len[a] >= mkfl ? ge3[a] : gel[a];
		ord_start[r + 1] = ord_start[r] + (all - m);
		mkf_start[r + 1] = mkf_start[r] + m;
		enum_start[r + 1] = enum_start[r] + all;
	}
	const uint64_t nord = ord_start[nr], np = enum_start[nr];
	plan.npairs = np;
	plan.perm.resize(np);
	plan.sa.resize(np);
	plan.sb.resize(np);
	std::atomic<size_t> next{0};
	auto work = [&]() {
		for (;;) {
			const size_t r0 = next.fetch_add(8);
			if (r0 >= nr)
				break;
			for (size_t r = r0; r < std::min(nr, r0 + 8); ++r) {
				const uint32_t a = (uint32_t)rows[r];
				uint64_t po = ord_start[r], pm = nord + mkf_start[r];
				const uint64_t e0 = enum_start[r];
				for (uint32_t t = 0; t < n; ++t) {
					const uint32_t j = order[t];
					if (j < a)
						continue;
					const uint64_t dst = is_mkf(a, j) ? pm++ : po++;
					plan.perm[dst] = e0 + (j - a);
					plan.sa[dst] = a;
					plan.sb[dst] = j;
				}
			}
		}
	};
	const int T = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)std::max(1, nthreads), np >> 16));
	if (T == 1) {
		work();
	} else {
		std::vector<std::thread> th;
		for (int t = 0; t < T; ++t)
			th.emplace_back(work);
		for (auto &t : th)
			t.join();
	}
	return RSK_OK;
}

extern "C" int rsk_search_pairs(rsk_ctx *ctx, const rsk_chainset *A, const rsk_chainset *B, uint64_t npairs,
		const uint32_t *ia, const uint32_t *ib, const rsk_search_opts *opts, rsk_results **out)
{
	if (!ctx || !A || !B || !out || (npairs && (!ia || !ib)))
		return fail(RSK_ERR_ARG, "rsk_search_pairs: null argument");
	*out = nullptr;
	if (npairs == 0) {
		*out = new rsk_results();
		return RSK_OK;
	}
	SearchPlan plan;
	int rc = build_explicit_plan(plan, A, B, npairs, ia, ib, ctx->host_threads, ctx->params.mkfl);
	if (rc)
		return rc;
	return search_impl(ctx, plan, opts, out, false);
}

static void add_stats(rsk_stats &acc, const rsk_stats &s)
{
	acc.pairs += s.pairs; acc.mu_filter_in += s.mu_filter_in; acc.mu_filter_rejected += s.mu_filter_rejected;
	acc.mu_saturated += s.mu_saturated; acc.sw_pairs += s.sw_pairs; acc.sw_cells += s.sw_cells;
	acc.evalue_pairs += s.evalue_pairs; acc.hits += s.hits; acc.kernel_launches += s.kernel_launches;
	acc.h2d_bytes += s.h2d_bytes; acc.d2h_bytes += s.d2h_bytes; acc.sw_kernel_ms += s.sw_kernel_ms;
	acc.mu_kernel_ms += s.mu_kernel_ms; acc.lddt_kernel_ms += s.lddt_kernel_ms; acc.total_ms += s.total_ms;
	acc.mkf_kernel_ms += s.mkf_kernel_ms; acc.sw_kernel_launches += s.sw_kernel_launches; acc.mkf_pairs += s.mkf_pairs;
}

// append the hits and paths of `src` to `dst` (path offsets re-based); src is consumed
static int append_results(rsk_results *dst, rsk_results *src)
{
	const uint64_t nh = dst->nhits + src->nhits, np = dst->npath + src->npath;
	if (nh * sizeof(rsk_hit) > dst->hits_cap) {
		size_t cap = 0;
		rsk_hit *p = (rsk_hit *)g_blocks.get(std::max<uint64_t>(nh, 2 * dst->nhits) * sizeof(rsk_hit), cap);
		if (!p) {
			delete src;
			return fail(RSK_ERR_NOMEM, "host memory for %llu hit records", (unsigned long long)nh);
		}
		if (dst->nhits)
			memcpy(p, dst->hits, dst->nhits * sizeof(rsk_hit));
		g_blocks.put(dst->hits, dst->hits_cap);
		dst->hits = p;
		dst->hits_cap = cap;
	}
	if (np > dst->paths_cap) {
		size_t cap = 0;
		char *p = (char *)g_blocks.get(std::max<uint64_t>(np, 2 * dst->npath), cap);
		if (!p) {
			delete src;
			return fail(RSK_ERR_NOMEM, "host memory for %llu path bytes", (unsigned long long)np);
		}
		if (dst->npath)
			memcpy(p, dst->paths, dst->npath);
		g_blocks.put(dst->paths, dst->paths_cap);
		dst->paths = p;
		dst->paths_cap = cap;
	}
	for (uint64_t k = 0; k < src->nhits; ++k) {
		rsk_hit h = src->hits[k];
		h.path_off += dst->npath;
		dst->hits[dst->nhits + k] = h;
	}
	if (src->npath)
		memcpy(dst->paths + dst->npath, src->paths, src->npath);
	dst->nhits = nh;
	dst->npath = np;
	delete src;
	return RSK_OK;
}

// runself.cpp:72-99: pairs (i, j >= i), A = chain i, B = chain j.  The pair list of a large set does not fit the host
// (1e5 chains = 5e9 pairs), so the rows are processed in chunks of at most RSK_SELF_CHUNK_PAIRS pairs (default 2^26) and the
// chunks' results are concatenated; the order of RSK_KEEP_ALL records is the enumeration order either way.
extern "C" int rsk_search_self(rsk_ctx *ctx, const rsk_chainset *Sx, const rsk_search_opts *opts, rsk_results **out)
{
	if (!ctx || !Sx || !out)
		return fail(RSK_ERR_ARG, "rsk_search_self: null argument");
	*out = nullptr;
	const uint64_t n = Sx->d.n;
	uint64_t chunk_pairs = (uint64_t)1 << 26;
	if (const char *e = getenv("RSK_SELF_CHUNK_PAIRS")) {
		const long long v = atoll(e);
		if (v > 0)
			chunk_pairs = (uint64_t)v;
	}
	rsk_results *total = nullptr;
	rsk_stats acc;
	memset(&acc, 0, sizeof(acc));
	for (uint64_t i0 = 0; i0 < n;) {
		uint64_t i1 = i0, np = 0;
		while (i1 < n && (i1 == i0 || np + (n - i1) <= chunk_pairs))
			np += n - i1++;
		SearchPlan plan;
		const double tp0 = now_ms();
		int rc = build_self_plan(plan, Sx, i0, i1, 1, ctx->host_threads, ctx->params.mkfl);
		g_t_plan += now_ms() - tp0;
		rsk_results *part = nullptr;
		if (!rc)
			rc = search_impl(ctx, plan, opts, &part, false);
		if (!rc) {
			add_stats(acc, ctx->stats);
			if (!total)
				total = part;
			else
				rc = append_results(total, part);
		}
		if (rc) {
			delete total;
			return rc;
		}
		i0 = i1;
	}
	ctx->stats = acc;
	*out = total ? total : new rsk_results();
	return RSK_OK;
}

// RunSelf on several GPUs: every rank holds the whole set; rank r takes the rows i = r, r + N, r + 2N ... of the pair triangle
// (row i has n - i pairs, so interleaved rows balance the ranks), hits are compacted on the device, gathered on the root and
// put back into the order of the unsharded search (a, then b).
extern "C" int rsk_search_self_sharded(rsk_ctx *ctx, rsk_comm *comm, const rsk_chainset *Sx, const rsk_search_opts *opts_in, int root,
		rsk_results **out)
{
	if (!ctx || !Sx || !out)
		return fail(RSK_ERR_ARG, "rsk_search_self_sharded: null argument");
	*out = nullptr;
	if (comm && comm_ctx(comm) != ctx)
		return fail(RSK_ERR_ARG, "rsk_search_self_sharded: the communicator belongs to a different context");
	const int N = comm_nranks(comm), me = comm_rank(comm);
	if (root < 0 || root >= N)
		return fail(RSK_ERR_ARG, "rsk_search_self_sharded: root %d of %d ranks", root, N);
	rsk_search_opts opts;
	memset(&opts, 0, sizeof(opts));
	if (opts_in)
		opts = *opts_in;
	const uint64_t n = Sx->d.n;
	uint64_t chunk_pairs = (uint64_t)1 << 26;
	if (const char *e = getenv("RSK_SELF_CHUNK_PAIRS")) {
		const long long v = atoll(e);
		if (v > 0)
			chunk_pairs = (uint64_t)v;
	}
	rsk_stats acc;
	memset(&acc, 0, sizeof(acc));
	int rc = sink_reset_empty(ctx);
	if (rc)
		return rc;
	bool first = true;
	for (uint64_t i0 = (uint64_t)me; i0 < n;) {
		uint64_t i1 = i0, np = 0;
		while (i1 < n && (i1 == i0 || np + (n - i1) <= chunk_pairs)) {
			np += n - i1;
			i1 += (uint64_t)N;
		}
		SearchPlan plan;
		rc = build_self_plan(plan, Sx, i0, i1, (uint64_t)N, ctx->host_threads, ctx->params.mkfl);
		if (rc)
			return rc;
		SinkOpts so;
		so.append = !first;
		if ((rc = search_impl(ctx, plan, &opts, nullptr, false, &so)))
			return rc;
		add_stats(acc, ctx->stats);
		first = false;
		i0 = i1;
	}
	ctx->stats = acc;
	rc = sink_finish(ctx, comm, root, opts, out);
	if (rc || !*out)
		return rc;
	// unsharded order: a ascending, b ascending (rows are interleaved over the ranks, and a rank schedules its pairs by length)
	rsk_results *res = *out;
	std::sort(res->hits, res->hits + res->nhits, [](const rsk_hit &x, const rsk_hit &y) { return x.a != y.a ? x.a < y.a : x.b < y.b; });
	return RSK_OK;
}

extern "C" int rsk_mu_gapless_scores(rsk_ctx *ctx, const rsk_chainset *A, const rsk_chainset *B, uint64_t npairs,
		const uint32_t *ia, const uint32_t *ib, float *out_profb, int32_t *out_int)
{
	if (!ctx || !A || !B || (npairs && (!ia || !ib)))
		return fail(RSK_ERR_ARG, "rsk_mu_gapless_scores: null argument");
	if (!A->has_mu || !B->has_mu)
		return fail(RSK_ERR_ARG, "rsk_mu_gapless_scores: both chain sets need Mu letters");
	if (npairs == 0)
		return RSK_OK;
	if (npairs > 0xffffffffull)
		return fail(RSK_ERR_LIMIT, "rsk_mu_gapless_scores: too many pairs");
	for (uint64_t k = 0; k < npairs; ++k)
		if (ia[k] >= A->d.n || ib[k] >= B->d.n)
			return fail(RSK_ERR_ARG, "pair %llu: chain index out of range", (unsigned long long)k);
	CK(cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	uint32_t *d_a = nullptr, *d_b = nullptr;
	float *d_f = nullptr;
	int *d_i = nullptr;
	auto cleanup = [&]() { cudaFree(d_a); cudaFree(d_b); cudaFree(d_f); cudaFree(d_i); };
	if (cudaMalloc((void **)&d_a, 4 * npairs) != cudaSuccess || cudaMalloc((void **)&d_b, 4 * npairs) != cudaSuccess ||
		cudaMalloc((void **)&d_f, 4 * npairs) != cudaSuccess || cudaMalloc((void **)&d_i, 4 * npairs) != cudaSuccess) {
		cudaGetLastError();
		cleanup();
		return fail(RSK_ERR_NOMEM, "rsk_mu_gapless_scores: device buffers");
	}
	cudaMemcpyAsync(d_a, ia, 4 * npairs, cudaMemcpyHostToDevice, st);
	cudaMemcpyAsync(d_b, ib, 4 * npairs, cudaMemcpyHostToDevice, st);
	GaplessArgs ga;
	memset(&ga, 0, sizeof(ga));
	ga.muA = A->d.mu; ga.offA = A->d.off; ga.lenA = A->d.len;
	ga.muB = B->d.mu; ga.offB = B->d.off; ga.lenB = B->d.len;
	ga.npairs = (uint32_t)npairs; ga.pair_a = d_a; ga.pair_b = d_b;
	ga.mu_f32 = ctx->d_mu_f32; ga.mu_i32 = ctx->d_mu_mx;
	ga.out_f = d_f; ga.out_i = d_i;
	const int nl = launch_mu_gapless(ga, st);
	cudaError_t e = nl < 0 ? cudaErrorLaunchFailure : cudaSuccess;
	if (e == cudaSuccess && out_profb)
		e = cudaMemcpyAsync(out_profb, d_f, 4 * npairs, cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess && out_int)
		e = cudaMemcpyAsync(out_int, d_i, 4 * npairs, cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess)
		e = cudaStreamSynchronize(st);
	cleanup();
	if (e != cudaSuccess)
		return fail(RSK_ERR_CUDA, "rsk_mu_gapless_scores: %s", cudaGetErrorString(e));
	ctx->stats.kernel_launches += nl;
	return RSK_OK;
}

// -global: DSSAligner::AlignQueryTarget_Global (global.cpp:7-33) for explicit pairs.  One record per pair, in pair order.
extern "C" int rsk_align_global(rsk_ctx *ctx, const rsk_chainset *A, const rsk_chainset *B, uint64_t npairs,
		const uint32_t *ia, const uint32_t *ib, rsk_results **out)
{
	if (!ctx || !A || !B || !out || (npairs && (!ia || !ib)))
		return fail(RSK_ERR_ARG, "rsk_align_global: null argument");
	*out = nullptr;
	if (npairs > 0xffffffffull)
		return fail(RSK_ERR_LIMIT, "rsk_align_global: too many pairs");
	uint32_t maxLA = 1, maxLB = 1;
	for (uint64_t k = 0; k < npairs; ++k) {
		if (ia[k] >= A->d.n || ib[k] >= B->d.n)
			return fail(RSK_ERR_ARG, "pair %llu: chain index out of range", (unsigned long long)k);
		const uint64_t la = A->hlen[ia[k]], lb = B->hlen[ib[k]];
		if (la * lb > 100ull * 1000 * 1000)  // viterbifastmem.cpp:36-37 dies here
			return fail(RSK_ERR_LIMIT, "rsk_align_global: pair %llu too long (LA=%llu, LB=%llu)", (unsigned long long)k,
					(unsigned long long)la, (unsigned long long)lb);
		maxLA = std::max<uint32_t>(maxLA, (uint32_t)la);
		maxLB = std::max<uint32_t>(maxLB, (uint32_t)lb);
	}
	rsk_results *res = new rsk_results();
	if (npairs == 0) {
		*out = res;
		return RSK_OK;
	}
	// Mu filter first when it is on (global.cpp:11-22).  MuFilter() does not look at the chain lengths, so the k-mer path
	// of the local search must not divert long chains here: the filter stage of the pair search runs with mkfl = infinity.
	std::vector<uint8_t> skip;
	std::vector<rsk_hit> filt;
	const bool filter = ctx->params.omega > 0 && A->has_mu && B->has_mu;
	if (filter) {
		const rsk_params saved = ctx->params;
		rsk_params p = saved;
		p.mkfl = 0xffffffffu;
		int rc = rsk_ctx_set_params(ctx, &p);
		rsk_results *fr = nullptr;
		if (rc == RSK_OK) {
			rsk_search_opts o;
			memset(&o, 0, sizeof(o));
			o.keep = RSK_KEEP_ALL;
			o.skip_evalue = 1;
			rc = rsk_search_pairs(ctx, A, B, npairs, ia, ib, &o, &fr);
		}
		const int rc2 = rsk_ctx_set_params(ctx, &saved);
		if (rc != RSK_OK || rc2 != RSK_OK || !fr || fr->nhits != npairs) {
			delete fr;
			delete res;
			return rc != RSK_OK ? rc : fail(RSK_ERR_CUDA, "rsk_align_global: filter stage failed");
		}
		skip.resize(npairs);
		filt.assign(fr->hits, fr->hits + npairs);
		for (uint64_t k = 0; k < npairs; ++k)
			skip[k] = (filt[k].flags & RSK_HIT_MU_REJECTED) ? 1 : 0;
		delete fr;
	}
	CK(cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	// work order: largest matrices first; path slots of LA+LB bytes
	std::vector<uint32_t> order(npairs);
	std::iota(order.begin(), order.end(), 0u);
	std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
		return (uint64_t)A->hlen[ia[x]] * B->hlen[ib[x]] > (uint64_t)A->hlen[ia[y]] * B->hlen[ib[y]];
	});
	std::vector<unsigned long long> poff(npairs);
	unsigned long long pool_bytes = 0;
	for (uint64_t k = 0; k < npairs; ++k) {
		poff[k] = pool_bytes;
		pool_bytes += (unsigned long long)A->hlen[ia[k]] + B->hlen[ib[k]];
	}
	const size_t tb_stride = (global_tb_bytes(maxLA, maxLB) + 15) & ~(size_t)15;
	const uint32_t bnd_stride = (maxLB + 1 + 3) & ~3u;
	const int wpb = global_warps_per_block();
	size_t warps = std::min<size_t>((size_t)ctx->num_sms * 2 * wpb, (size_t)npairs);  // 128 registers: two 8-warp CTAs per SM
	warps = std::min<size_t>(warps, std::max<size_t>(1, ((size_t)8 << 30) / tb_stride));  // at most 8 GB of trace scratch
	const int blocks = (int)((warps + wpb - 1) / wpb);
	warps = (size_t)blocks * wpb;
	// device buffers live in the context and are reused by the next call (the trace scratch alone is up to 8 GB: allocating and
	// freeing it per call cost more than the kernel for batches of a few hundred thousand pairs)
	if (ctx->gl_a.ensure(npairs) || ctx->gl_b.ensure(npairs) || ctx->gl_order.ensure(npairs) || ctx->gl_skip.ensure(npairs) ||
		ctx->gl_poff.ensure(npairs) || ctx->gl_pool.ensure(pool_bytes + 16) || ctx->gl_rec.ensure(npairs) ||
		ctx->gl_tb.ensure(tb_stride * warps) || ctx->gl_bnd.ensure((size_t)2 * bnd_stride * warps) || ctx->gl_cnt.ensure(1)) {
		cudaGetLastError();
		ctx->gl_tb.release();
		delete res;
		return fail(RSK_ERR_NOMEM, "rsk_align_global: device buffers (%zu warps x %zu trace bytes)", warps, tb_stride);
	}
	uint32_t *d_a = ctx->gl_a.p, *d_b = ctx->gl_b.p, *d_order = ctx->gl_order.p;
	uint8_t *d_skip = ctx->gl_skip.p, *d_tb = ctx->gl_tb.p;
	unsigned long long *d_poff = ctx->gl_poff.p;
	char *d_pool = ctx->gl_pool.p;
	GlobalRec *d_rec = ctx->gl_rec.p;
	float *d_bnd = ctx->gl_bnd.p;
	unsigned int *d_cnt = ctx->gl_cnt.p;
	auto cleanup = [&]() {};
	cudaMemcpyAsync(d_a, ia, 4 * npairs, cudaMemcpyHostToDevice, st);
	cudaMemcpyAsync(d_b, ib, 4 * npairs, cudaMemcpyHostToDevice, st);
	cudaMemcpyAsync(d_order, order.data(), 4 * npairs, cudaMemcpyHostToDevice, st);
	cudaMemcpyAsync(d_poff, poff.data(), 8 * npairs, cudaMemcpyHostToDevice, st);
	if (filter)
		cudaMemcpyAsync(d_skip, skip.data(), npairs, cudaMemcpyHostToDevice, st);
	cudaMemsetAsync(d_cnt, 0, sizeof(unsigned int), st);
	GlobalArgs ga;
	memset(&ga, 0, sizeof(ga));
	ga.profA = A->d.prof8; ga.offA = A->d.off; ga.lenA = A->d.len;
	ga.profB = B->d.prof8; ga.offB = B->d.off; ga.lenB = B->d.len;
	ga.npairs = (uint32_t)npairs;
	ga.pair_a = d_a; ga.pair_b = d_b; ga.order = d_order;
	ga.skip = filter ? d_skip : nullptr;
	ga.path_off = d_poff; ga.pool = d_pool; ga.rec = d_rec;
	ga.tb = d_tb; ga.tb_stride = tb_stride;
	ga.bnd = d_bnd; ga.bnd_stride = bnd_stride;
	ga.counter = d_cnt;
	ga.tables = ctx->d_tables;
	cudaEventRecord(ctx->ev[0], st);
	const int nl = launch_global(ga, blocks, st);
	cudaEventRecord(ctx->ev[1], st);
	std::vector<GlobalRec> recs(npairs);
	size_t cap = 0;
	res->paths = (char *)g_blocks.get(pool_bytes + 16, cap);
	res->paths_cap = cap;
	res->hits = (rsk_hit *)g_blocks.get(npairs * sizeof(rsk_hit), cap);
	res->hits_cap = cap;
	cudaError_t e = nl < 0 ? cudaErrorLaunchFailure : cudaSuccess;
	if (e == cudaSuccess)
		e = cudaMemcpyAsync(recs.data(), d_rec, sizeof(GlobalRec) * npairs, cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess)
		e = cudaMemcpyAsync(res->paths, d_pool, pool_bytes, cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess)
		e = cudaStreamSynchronize(st);
	cleanup();
	if (e != cudaSuccess || !res->paths || !res->hits) {
		delete res;
		return fail(RSK_ERR_CUDA, "rsk_align_global: %s", cudaGetErrorString(e));
	}
	for (uint64_t k = 0; k < npairs; ++k) {
		rsk_hit &h = res->hits[k];
		memset(&h, 0, sizeof(h));
		h.a = ia[k]; h.b = ib[k];
		// ClearAlign values (dssaligner.cpp:906-927) for everything the global path does not set
		h.hi_a = h.hi_b = h.ids = h.gaps = 0xffffffffu;
		h.ts = -FLT_MAX;
		h.pvalue = h.evalue = h.qual = FLT_MAX;
		h.flags = RSK_HIT_GLOBAL;
		if (filter) {
			h.mu_fwd = filt[k].mu_fwd; h.mu_rev = filt[k].mu_rev; h.mu_score = filt[k].mu_score;
		}
		h.score = recs[k].score;  // m_GlobalScore
		if (filter && skip[k]) {
			h.flags |= RSK_HIT_MU_REJECTED;
			h.lo_a = h.lo_b = 0xffffffffu;
			continue;
		}
		h.lo_a = h.lo_b = 0;      // global.cpp:30-31
		h.path_len = recs[k].path_len;
		h.path_off = poff[k];
	}
	res->nhits = npairs;
	res->npath = pool_bytes;
	ctx->stats.kernel_launches += nl;
	ctx->stats.pairs = npairs;
	{
		float ms = 0;
		if (cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]) == cudaSuccess)
			ctx->stats.sw_kernel_ms = ms;  // the global kernel's device time
	}
	*out = res;
	return RSK_OK;
}

extern "C" int rsk_chainset_selfrev(rsk_ctx *ctx, rsk_chainset *S, const rsk_chainset *Srev, float *scores_out)
{
	if (!ctx || !S || !Srev)
		return fail(RSK_ERR_ARG, "rsk_chainset_selfrev: null argument");
	if (S->d.n != Srev->d.n)
		return fail(RSK_ERR_ARG, "rsk_chainset_selfrev: %u chains but %u reversed chains", S->d.n, Srev->d.n);
	for (uint32_t i = 0; i < S->d.n; ++i)
		if (S->hlen[i] != Srev->hlen[i])
			return fail(RSK_ERR_ARG, "rsk_chainset_selfrev: chain %u has length %u, its reverse %u", i, S->hlen[i], Srev->hlen[i]);
	const uint32_t n = S->d.n;
	std::vector<uint32_t> idx(n);
	std::iota(idx.begin(), idx.end(), 0u);
	rsk_search_opts o;
	memset(&o, 0, sizeof(o));
	o.keep = RSK_KEEP_ALL;
	o.skip_evalue = 1;  // only m_AlnFwdScore is used (alignpair.cpp:24)
	rsk_results *res = nullptr;
	int rc = rsk_search_pairs(ctx, S, Srev, n, idx.data(), idx.data(), &o, &res);
	if (rc)
		return rc;
	std::vector<float> sr(n);
	for (uint32_t i = 0; i < n; ++i)
		sr[i] = res->hits[i].score;
	rsk_results_free(res);
	CK(cudaSetDevice(ctx->device));
	CK(cudaMemcpyAsync(S->d.selfrev, sr.data(), sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	if (scores_out)
		memcpy(scores_out, sr.data(), sizeof(float) * n);
	return RSK_OK;
}

// ------------------------------------------------------------------------------------------------
// results
// ------------------------------------------------------------------------------------------------
extern "C" uint64_t rsk_results_count(const rsk_results *r) { return r ? r->nhits : 0; }
extern "C" const rsk_hit *rsk_results_hits(const rsk_results *r) { return (r && r->nhits) ? r->hits : nullptr; }
extern "C" const char *rsk_results_paths(const rsk_results *r) { return (r && r->npath) ? r->paths : nullptr; }
extern "C" uint64_t rsk_results_paths_bytes(const rsk_results *r) { return r ? r->npath : 0; }
extern "C" void rsk_results_free(rsk_results *r) { delete r; }
