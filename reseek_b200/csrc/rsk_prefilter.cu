// rsk_prefilter.cu - host side of the `-search Q -db DB -fast` pipeline (search.cpp:76-111):
//   rsk_prefilter      = MuPreFilter (muprefilter.cpp:64-133): GPU index build / probe / two-hit extension (prefilter_kernel.cu)
//                        + the per-query top-B bag RankedScoresBag (rankedscoresbag.cpp), which is order dependent and tiny
//                        and therefore runs here on the host over the (target, query, score) triples the GPU produced.
//   rsk_postfilter     = PostMuFilter's scan loop (postmufilter.cpp:116-208) as one explicit-pair batch of the SW path.
//   rsk_search_fast_db = both.
// No CPU fallback: the triples only ever come from the kernels.
#include <algorithm>
#include <chrono>
#include <map>
#include <memory>
#include <thread>

#include "rsk_host.cuh"
#include "rsk_multi.cuh"

struct rsk_prefilter_result {
	std::vector<uint32_t> t, q;
	std::vector<uint16_t> s;
	uint64_t raw = 0;
	uint32_t ntargets = 0;
};

namespace {

// ---- RankedScoresBag (rankedscoresbag.cpp:5-51): per query keep the top B targets; lazy truncation at 2B ----
struct Bag {
	std::vector<uint32_t> t;
	std::vector<uint16_t> s;
	uint16_t lo = 0;
};

// QuickSortOrderDesc (sort.h:71-108): Hoare partition around the middle element on an index array.  Not stable; ties at
// the cut-off are resolved by exactly this recursion, so it is restated rather than replaced by std::sort.
void order_desc(const uint16_t *v, int left, int right, uint32_t *order)
{
	int i = left, j = right;
	const uint16_t pivot = v[order[(left + right) / 2]];
	while (i <= j) {
		while (v[order[i]] > pivot)
			i++;
		while (v[order[j]] < pivot)
			j--;
		if (i <= j) {
			std::swap(order[i], order[j]);
			i++;
			j--;
		}
	}
	if (left < j)
		order_desc(v, left, j, order);
	if (i < right)
		order_desc(v, i, right, order);
}

void bag_truncate(Bag &b, uint32_t B)  // TruncateVecs
{
	const uint32_t n = (uint32_t)b.s.size();
	if (n < B)
		return;
	std::vector<uint32_t> order(n);
	for (uint32_t i = 0; i < n; ++i)
		order[i] = i;
	order_desc(b.s.data(), 0, (int)n - 1, order.data());
	std::vector<uint32_t> nt(B);
	std::vector<uint16_t> ns(B);
	for (uint32_t k = 0; k < B; ++k) {
		nt[k] = b.t[order[k]];
		ns[k] = b.s[order[k]];
	}
	b.t.swap(nt);
	b.s.swap(ns);
	b.lo = b.s[B - 1];
}

void bag_add(Bag &b, uint32_t B, uint32_t t, uint16_t score)  // AddScore
{
	if (score < b.lo)
		return;
	b.s.push_back(score);
	b.t.push_back(t);
	if (b.s.size() >= (size_t)2 * B)
		bag_truncate(b, B);
}

// A query's bag only ever sees that query's triples, in stream order: the bags are independent, so the stream is replayed by
// several host threads, each feeding the queries q % T == t.  Results are those of the single-threaded loop.
void feed_bags(std::vector<Bag> &bags, uint32_t B, const uint32_t *t, const uint32_t *q, const uint16_t *s, uint64_t n, int nthreads)
{
	const int T = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)std::max(1, nthreads), std::min<uint64_t>(bags.size(), n >> 14)));
	if (T == 1) {
		for (uint64_t k = 0; k < n; ++k)
			bag_add(bags[q[k]], B, t[k], s[k]);
		return;
	}
	std::vector<std::thread> th;
	for (int w = 0; w < T; ++w)
		th.emplace_back([&, w]() {
			for (uint64_t k = 0; k < n; ++k)
				if ((int)(q[k] % (uint32_t)T) == w)
					bag_add(bags[q[k]], B, t[k], s[k]);
		});
	for (auto &x : th)
		x.join();
}

// RankedScoresBag::ToTsv (rankedscoresbag.cpp:185-232): final truncation, then target -> queries (ascending)
void finish_bags(std::vector<Bag> &bags, uint32_t B, rsk_prefilter_result &res, int nthreads)
{
	const uint32_t nQ = (uint32_t)bags.size();
	const int T = (int)std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)std::max(1, nthreads), nQ / 4));
	if (T == 1) {
		for (uint32_t q = 0; q < nQ; ++q)
			bag_truncate(bags[q], B);
	} else {
		std::vector<std::thread> th;
		for (int w = 0; w < T; ++w)
			th.emplace_back([&, w]() {
				for (uint32_t q = (uint32_t)w; q < nQ; q += (uint32_t)T)
					bag_truncate(bags[q], B);
			});
		for (auto &x : th)
			x.join();
	}
	// invert: (target, query) ascending; a (target, query) pair occurs at most once
	std::vector<uint64_t> key;
	std::vector<uint16_t> sc;
	size_t tot = 0;
	for (auto &b : bags)
		tot += b.t.size();
	key.reserve(tot);
	for (uint32_t q = 0; q < nQ; ++q)
		for (size_t k = 0; k < bags[q].t.size(); ++k)
			key.push_back(((uint64_t)bags[q].t[k] << 32) | ((uint64_t)q << 16) | bags[q].s[k]);
	std::sort(key.begin(), key.end());
	res.t.reserve(tot); res.q.reserve(tot); res.s.reserve(tot);
	uint32_t nt = 0;
	for (size_t k = 0; k < key.size(); ++k) {
		const uint32_t tt = (uint32_t)(key[k] >> 32);
		nt += (k == 0 || tt != (uint32_t)(key[k - 1] >> 32));
		res.t.push_back(tt);
		res.q.push_back((uint32_t)(key[k] >> 16) & 0xffffu);
		res.s.push_back((uint16_t)(key[k] & 0xffffu));
	}
	res.ntargets = nt;
}

struct PfScratch {
	DevBuf<uint8_t> muq, tmp;
	DevBuf<uint32_t> qk_off, qk_code, qk_val, nb_count, key_a, key_b, val_a, val_b;
	DevBuf<uint2> row, qinfo;
	DevBuf<uint32_t> fuse_counter;
	DevBuf<unsigned long long> nb_off, hit_count, hit_off, cand_off;
	DevBuf<uint32_t> hit_key, hit_sorted, cand_count, cand_t, cand_q;
	DevBuf<unsigned> best;
	DevBuf<uint16_t> cand_s;
	DevBuf<uint32_t> raw_q, srt_q, bag_n;
	DevBuf<unsigned long long> raw_v, srt_v, seg_begin, seg_end, bag_key, bag_off, dense_a, dense_b;
	int *kmer_mx = nullptr;
	uint8_t *nb_tab = nullptr;
	~PfScratch()
	{
		muq.release(); tmp.release(); qk_off.release(); qk_code.release(); qk_val.release(); nb_count.release();
		key_a.release(); key_b.release(); val_a.release(); val_b.release(); row.release(); qinfo.release(); fuse_counter.release();
		nb_off.release(); hit_count.release(); hit_off.release(); cand_off.release(); hit_key.release();
		hit_sorted.release(); cand_count.release(); cand_t.release(); cand_q.release(); best.release(); cand_s.release();
		raw_q.release(); srt_q.release(); bag_n.release(); raw_v.release(); srt_v.release(); seg_begin.release(); seg_end.release();
		bag_key.release(); bag_off.release(); dense_a.release(); dense_b.release();
		if (kmer_mx)
			cudaFree(kmer_mx);
		if (nb_tab)
			cudaFree(nb_tab);
	}
};

constexpr uint32_t kDict = 36u * 36 * 36 * 36 * 36;  // DICT_SIZE (prefiltermuparams.h)

// RSK_TIMING=1: wall-clock phases of rsk_prefilter on stderr (developer aid)
struct PhaseTimer {
	bool on = getenv("RSK_TIMING") != nullptr;
	std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
	cudaStream_t st = nullptr;
	void mark(const char *what)
	{
		if (!on)
			return;
		cudaStreamSynchronize(st);
		const auto t1 = std::chrono::steady_clock::now();
		fprintf(stderr, "[rsk_prefilter] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
		t0 = t1;
	}
};

#define PFL(call)                                                                 \
	do {                                                                          \
		const int n_ = (call);                                                    \
		if (n_ < 0)                                                               \
			return fail(RSK_ERR_CUDA, "%s: launch failed: %s", #call, cudaGetErrorString(cudaGetLastError())); \
		launches += (uint64_t)n_;                                                 \
	} while (0)
#define NOMEM(x)                                                                  \
	do {                                                                          \
		if ((x) != 0)                                                             \
			return fail(RSK_ERR_NOMEM, "rsk_prefilter: out of device memory (%s)", #x); \
	} while (0)

// grow a device array to `need` elements keeping its first `used` elements (stream-ordered copy, then the old block is freed)
template <typename T>
int grow_keep(DevBuf<T> &b, size_t need, size_t used, cudaStream_t st)
{
	if (need <= b.cap)
		return 0;
	const size_t want = std::max(need, 2 * b.cap) + 1024;
	T *p = nullptr;
	if (cudaMalloc((void **)&p, want * sizeof(T)) != cudaSuccess) {
		cudaGetLastError();
		return -1;
	}
	if (b.p && used) {
		if (cudaMemcpyAsync(p, b.p, used * sizeof(T), cudaMemcpyDeviceToDevice, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
			cudaFree(p);
			return -1;
		}
	}
	if (b.p)
		cudaFree(b.p);
	b.p = p;
	b.cap = want;
	return 0;
}

PfScratch &pf_scratch(rsk_ctx *ctx)
{
	// grow-only device scratch kept in the context: a streamed database calls this once per block of targets
	if (!ctx->pf_scratch) {
		ctx->pf_scratch = new PfScratch();
		ctx->pf_scratch_free = [](void *p) { delete static_cast<PfScratch *>(p); };
	}
	return *static_cast<PfScratch *>(ctx->pf_scratch);
}

// K6..K8 over all targets of T.  The (target, query, score) triples with a two-hit diagonal stay ON THE DEVICE, in stream
// order (targets ascending), as S.raw_q[k] / S.raw_v[k] = target<<16 | score with target numbered from t_base.
int prefilter_raw_device(rsk_ctx *ctx, const rsk_chainset *Q, const rsk_chainset *T, const rsk_prefilter_opts &o, uint32_t t_base,
		unsigned long long &nraw, uint64_t &launches)
{
	nraw = 0;
	const uint32_t nQ = Q->d.n, nT = T ? T->d.n : 0;
	if (nQ >= (1u << 16))
		return fail(RSK_ERR_LIMIT, "rsk_prefilter: at most 65535 queries per call (got %u); split the query set", nQ);
	if (Q->maxlen > 0xffffu)
		return fail(RSK_ERR_LIMIT, "rsk_prefilter: query chains are limited to 65535 residues (uint16 position, mudex.h:19-49)");
	if ((uint64_t)t_base + nT > 0xffffffffull)
		return fail(RSK_ERR_LIMIT, "rsk_prefilter: target index exceeds 32 bits");
	if (nQ == 0 || nT == 0)
		return RSK_OK;
	// <= 100 queries: neighbourhoods go into the query index and exact k-mers are therefore entered twice (mudex.cpp:146-174)
	const bool qhood = o.index_mode == 1 || (o.index_mode == 0 && nQ <= 100);
	CK(cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	PfScratch &S = pf_scratch(ctx);
	PfArgs a = {};
	PhaseTimer tm;
	tm.st = st;
	if (!S.kmer_mx) {
		std::vector<int> mx(36 * 36);
		const int8_t *m8 = rsk_mu_kmer_matrix_i8();
		for (int k = 0; k < 36 * 36; ++k)
			mx[k] = m8[k];
		// per letter x: the 36 letters y by S[x][y] descending (ties by letter), those scores, ge[t] = #{y : S[x][y] >= t - 32}
		std::vector<uint8_t> nb(36 * 136);
		for (int x = 0; x < 36; ++x) {
			int order[36];
			for (int y = 0; y < 36; ++y)
				order[y] = y;
			std::stable_sort(order, order + 36, [&](int u, int v) { return m8[36 * x + u] > m8[36 * x + v]; });
			for (int r = 0; r < 36; ++r) {
				nb[136 * x + r] = (uint8_t)order[r];
				nb[136 * x + 36 + r] = (uint8_t)m8[36 * x + order[r]];
			}
			for (int t = 0; t < 64; ++t) {
				int c = 0;
				for (int y = 0; y < 36; ++y)
					c += m8[36 * x + y] >= t - 32;
				nb[136 * x + 72 + t] = (uint8_t)c;
			}
			if (m8[36 * x + order[35]] < -32 || m8[36 * x + order[0]] > 30)
				return fail(RSK_ERR_LIMIT, "rsk_prefilter: k-mer score outside the neighbourhood tables' range");
		}
		CK(cudaMalloc((void **)&S.kmer_mx, sizeof(int) * 36 * 36));
		CK(cudaMalloc((void **)&S.nb_tab, nb.size()));
		CK(cudaMemcpyAsync(S.kmer_mx, mx.data(), sizeof(int) * 36 * 36, cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(S.nb_tab, nb.data(), nb.size(), cudaMemcpyHostToDevice, st));
		CK(cudaStreamSynchronize(st));
	}
	a.kmer_mx = S.kmer_mx;
	a.nb_tab = S.nb_tab;
	a.nQ = nQ;
	a.offQ = Q->d.off; a.lenQ = Q->d.len;
	a.muT = T->d.mu; a.offT = T->d.off; a.lenT = T->d.len;
	a.exact_twice = qhood ? 1u : 0u;
	a.t_base = t_base;
	// query letters, K/L exchanged unless told otherwise
	if (!o.no_kl_swap) {
		NOMEM(S.muq.ensure(Q->d.total));
		PFL(pf_launch_swap_kl(Q->d.mu, S.muq.p, Q->d.total, st));
		a.muQ = S.muq.p;
	} else {
		a.muQ = Q->d.mu;
	}
	tm.mark("setup");
	// ---- K6: query index ----
	std::vector<uint32_t> qk_off(nQ + 1, 0);
	for (uint32_t q = 0; q < nQ; ++q)
		qk_off[q + 1] = qk_off[q] + (Q->hlen[q] >= 7 ? Q->hlen[q] - 6 : 0);
	const uint32_t nqk = qk_off[nQ];
	a.nqk = nqk;
	unsigned long long nindex = 0;
	NOMEM(S.row.ensure(kDict));
	CK(cudaMemsetAsync(S.row.p, 0, sizeof(uint2) * kDict, st));
	if (nqk) {
		NOMEM(S.qk_off.ensure(nQ + 1));
		NOMEM(S.qk_code.ensure(nqk));
		NOMEM(S.qk_val.ensure(nqk));
		NOMEM(S.nb_count.ensure(nqk));
		NOMEM(S.nb_off.ensure(nqk + 1));
		CK(cudaMemcpyAsync(S.qk_off.p, qk_off.data(), sizeof(uint32_t) * (nQ + 1), cudaMemcpyHostToDevice, st));
		a.qk_off = S.qk_off.p; a.qk_code = S.qk_code.p; a.qk_val = S.qk_val.p; a.nb_count = S.nb_count.p; a.nb_off = S.nb_off.p;
		PFL(pf_launch_query_kmers(a, st));
		PFL(pf_launch_neighborhood(a, false, st));
		// sizes -> offsets (host scan: nqk is the query residue count, small)
		std::vector<uint32_t> cnt(nqk);
		CK(cudaMemcpyAsync(cnt.data(), S.nb_count.p, sizeof(uint32_t) * nqk, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		std::vector<unsigned long long> off(nqk + 1, 0);
		for (uint32_t i = 0; i < nqk; ++i)
			off[i + 1] = off[i] + cnt[i];
		nindex = off[nqk];
		if (nindex >= 0xffffffffull)
			return fail(RSK_ERR_LIMIT, "rsk_prefilter: query index of %llu entries exceeds 2^32; split the query set", nindex);
		CK(cudaMemcpyAsync(S.nb_off.p, off.data(), sizeof(unsigned long long) * (nqk + 1), cudaMemcpyHostToDevice, st));
		if (nindex) {
			NOMEM(S.key_a.ensure(nindex));
			NOMEM(S.key_b.ensure(nindex));
			NOMEM(S.val_a.ensure(nindex));
			NOMEM(S.val_b.ensure(nindex));
			a.ix_key = S.key_a.p; a.ix_val = S.val_a.p;
			PFL(pf_launch_neighborhood(a, true, st));
			size_t tb = 0;
			if (pf_sort_pairs(S.key_a.p, S.key_b.p, S.val_a.p, S.val_b.p, nindex, nullptr, tb, st))
				return fail(RSK_ERR_CUDA, "rsk_prefilter: cub sort sizing failed");
			NOMEM(S.tmp.ensure(tb));
			if (pf_sort_pairs(S.key_a.p, S.key_b.p, S.val_a.p, S.val_b.p, nindex, S.tmp.p, tb, st))
				return fail(RSK_ERR_CUDA, "rsk_prefilter: cub sort failed: %s", cudaGetErrorString(cudaGetLastError()));
			launches += 4;
			PFL(pf_launch_mark_rows(S.key_b.p, nindex, S.row.p, st));
			a.ix_key = S.key_b.p; a.ix_val = S.val_b.p;
		}
	}
	a.row = S.row.p;
	a.diag_safe = (uint64_t)Q->maxlen + T->maxlen <= 0x4000u ? 1u : 0u;
	tm.mark("K6 query index");
	if (!nindex)
		return RSK_OK;

	// ---- K7/K8: targets whose (query, diagonal) pairs fit two shared-memory bitmaps go through the fused kernel; the others
	// (long targets against large query blocks) through global memory: count pass, keys per target, segmented sort, run scan ----
	std::vector<uint2> qinfo(nQ);
	uint64_t sumLQ = 0;
	for (uint32_t q = 0; q < nQ; ++q) {
		qinfo[q] = make_uint2(Q->hlen[q], (uint32_t)sumLQ);
		sumLQ += Q->hlen[q];
	}
	if (sumLQ >= (1ull << 31))
		return fail(RSK_ERR_LIMIT, "rsk_prefilter: more than 2^31 query residues; split the query set");
	NOMEM(S.qinfo.ensure(nQ));
	CK(cudaMemcpyAsync(S.qinfo.p, qinfo.data(), sizeof(uint2) * nQ, cudaMemcpyHostToDevice, st));
	a.qinfo = S.qinfo.p;
	a.sum_lenQ = (uint32_t)sumLQ;
	if (const char *e = getenv("RSK_PF_QUEUE"))
		a.queue_cap = (uint32_t)atoi(e);
	a.no_stage = getenv("RSK_PF_NOSTAGE") ? 1u : 0u;
	NOMEM(S.fuse_counter.ensure(1));
	a.fuse_counter = S.fuse_counter.p;
	// RSK_PF_NOFUSE=1 sends every target through global memory (the parity tests run both paths)
	const bool nofuse = getenv("RSK_PF_NOFUSE") != nullptr;
	const unsigned long long bits_small = nofuse ? 0 : pf_fuse_max_bits(0), bits_max = nofuse ? 0 : pf_fuse_max_bits(1);
	auto fuse_bits = [&](uint32_t t) -> unsigned long long { return T->hlen[t] < 7 ? 0 : sumLQ + (unsigned long long)nQ * (T->hlen[t] - 1); };
	bool any_global = false;
	for (uint32_t t = 0; t < nT && !any_global; ++t)
		any_global = fuse_bits(t) > bits_max;
	std::vector<unsigned long long> hcnt(nT, 0);
	if (any_global) {
		NOMEM(S.hit_count.ensure(nT));
		a.t_begin = 0;
		a.hit_count = S.hit_count.p;
		PFL(pf_launch_probe(a, nT, false, st));
		CK(cudaMemcpyAsync(hcnt.data(), S.hit_count.p, sizeof(unsigned long long) * nT, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		for (uint32_t t = 0; t < nT; ++t)
			if (fuse_bits(t) <= bits_max)
				hcnt[t] = 0;  // taken by the fused kernel
		tm.mark("K7 count pass");
	}
	const unsigned long long kMaxHits = 1ull << 28;              // 1 GB of keys + 1 GB sorted per batch
	const uint32_t kMaxT = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(1u << 16, ((uint64_t)1 << 26) / nQ));
	std::vector<unsigned long long> hoff, coff;
	std::vector<uint32_t> ccnt;
	for (uint32_t t0 = 0; t0 < nT;) {
		uint32_t t1 = t0;
		unsigned long long big = 0;
		while (t1 < nT && t1 - t0 < kMaxT && (t1 == t0 || big + hcnt[t1] <= kMaxHits))
			big += hcnt[t1++];
		const uint32_t ntl = t1 - t0;
		if (big >= (1ull << 32))
			return fail(RSK_ERR_LIMIT, "rsk_prefilter: target %u alone produces %llu index hits", t0, big);
		hoff.assign(ntl + 1, 0);
		int which = 0;
		for (uint32_t k = 0; k < ntl; ++k) {
			hoff[k + 1] = hoff[k] + hcnt[t0 + k];
			const unsigned long long b = fuse_bits(t0 + k);
			which |= b == 0 || nofuse ? 0 : b <= bits_small ? 1 : b <= bits_max ? 2 : 0;
		}
		if (!which && !big) {  // nothing to extend in this batch
			t0 = t1;
			continue;
		}
		NOMEM(S.hit_off.ensure(ntl + 1));
		NOMEM(S.best.ensure((size_t)ntl * nQ));
		NOMEM(S.cand_count.ensure(ntl));
		NOMEM(S.cand_off.ensure(ntl + 1));
		CK(cudaMemsetAsync(S.best.p, 0, sizeof(unsigned) * (size_t)ntl * nQ, st));
		a.t_begin = t0;
		a.best = S.best.p;
		a.cand_count = S.cand_count.p; a.cand_off = S.cand_off.p;
		if (which) {
			CK(cudaMemsetAsync(S.fuse_counter.p, 0, sizeof(uint32_t), st));
			PFL(pf_launch_probe_extend(a, ntl, which, st));
		}
		tm.mark("K7+K8 in shared memory");
		if (big) {
			NOMEM(S.hit_key.ensure(big));
			NOMEM(S.hit_sorted.ensure(big));
			CK(cudaMemcpyAsync(S.hit_off.p, hoff.data(), sizeof(unsigned long long) * (ntl + 1), cudaMemcpyHostToDevice, st));
			a.hit_off = S.hit_off.p; a.hit_key = S.hit_key.p; a.hit_sorted = S.hit_sorted.p;
			PFL(pf_launch_probe(a, ntl, true, st));
			size_t tb = 0;
			if (pf_segmented_sort(S.hit_key.p, S.hit_sorted.p, big, ntl, S.hit_off.p, nullptr, tb, st))
				return fail(RSK_ERR_CUDA, "rsk_prefilter: cub segmented sort sizing failed");
			NOMEM(S.tmp.ensure(tb));
			if (pf_segmented_sort(S.hit_key.p, S.hit_sorted.p, big, ntl, S.hit_off.p, S.tmp.p, tb, st))
				return fail(RSK_ERR_CUDA, "rsk_prefilter: cub segmented sort failed: %s", cudaGetErrorString(cudaGetLastError()));
			launches += 3;
			tm.mark("K7 probe+sort");
			PFL(pf_launch_extend(a, ntl, st));
		}
		PFL(pf_launch_cands(a, ntl, false, st));
		tm.mark("K8 extend+count");
		ccnt.resize(ntl);
		CK(cudaMemcpyAsync(ccnt.data(), S.cand_count.p, sizeof(uint32_t) * ntl, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		coff.assign(ntl + 1, 0);
		for (uint32_t k = 0; k < ntl; ++k)
			coff[k + 1] = coff[k] + ccnt[k];
		const unsigned long long nc = coff[ntl];
		if (nc) {
			// stream order of the reference at -threads 1: targets ascending (AddTwoHitDiag -> AddScore, prefiltermu.cpp:288-313)
			NOMEM(grow_keep(S.raw_q, nraw + nc, nraw, st));
			NOMEM(grow_keep(S.raw_v, nraw + nc, nraw, st));
			CK(cudaMemcpyAsync(S.cand_off.p, coff.data(), sizeof(unsigned long long) * (ntl + 1), cudaMemcpyHostToDevice, st));
			a.raw_base = nraw; a.raw_q = S.raw_q.p; a.raw_v = S.raw_v.p;
			PFL(pf_launch_cands(a, ntl, true, st));
			CK(cudaStreamSynchronize(st));  // coff is reused by the next batch
			nraw += nc;
		}
		tm.mark("triples");
		t0 = t1;
	}
	return RSK_OK;
}

// RankedScoresBag over device triples in stream order (pf_bag_kernel): stable sort by query, one warp per query replays
// AddScore / TruncateVecs, the surviving entries are sorted (target, query) ascending and only that candidate list is read
// back (rankedscoresbag.cpp:185-232).
int bag_device(rsk_ctx *ctx, uint32_t nQ, uint32_t B, const uint32_t *d_q, const unsigned long long *d_v, unsigned long long n,
		rsk_prefilter_result &res, uint64_t &launches)
{
	res.raw = n;
	if (n == 0 || nQ == 0)
		return RSK_OK;
	if (pf_bag_smem_bytes(B) > 220 * 1024)
		return fail(RSK_ERR_LIMIT, "rsk_prefilter: -rsb_size %u exceeds the bag kernel's shared memory (max ~6000)", B);
	CK(cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	PfScratch &S = pf_scratch(ctx);
	PhaseTimer tm;
	tm.st = st;
	NOMEM(S.srt_q.ensure(n));
	NOMEM(S.srt_v.ensure(n));
	NOMEM(S.seg_begin.ensure(nQ));
	NOMEM(S.seg_end.ensure(nQ));
	NOMEM(S.bag_key.ensure((size_t)nQ * B));
	NOMEM(S.bag_n.ensure(nQ));
	NOMEM(S.bag_off.ensure(nQ + 1));
	size_t tb = 0;
	if (pf_sort_by_query(d_q, S.srt_q.p, d_v, S.srt_v.p, n, nullptr, tb, st))
		return fail(RSK_ERR_CUDA, "rsk_prefilter: cub sort sizing failed");
	NOMEM(S.tmp.ensure(tb));
	if (pf_sort_by_query(d_q, S.srt_q.p, d_v, S.srt_v.p, n, S.tmp.p, tb, st))
		return fail(RSK_ERR_CUDA, "rsk_prefilter: sort by query failed: %s", cudaGetErrorString(cudaGetLastError()));
	launches += 3;
	CK(cudaMemsetAsync(S.seg_begin.p, 0, sizeof(unsigned long long) * nQ, st));
	CK(cudaMemsetAsync(S.seg_end.p, 0, sizeof(unsigned long long) * nQ, st));
	PFL(pf_launch_mark_segments(S.srt_q.p, n, S.seg_begin.p, S.seg_end.p, st));
	PFL(pf_launch_bag(S.srt_v.p, S.seg_begin.p, S.seg_end.p, nQ, B, S.bag_key.p, S.bag_n.p, st));
	std::vector<uint32_t> bn(nQ);
	CK(cudaMemcpyAsync(bn.data(), S.bag_n.p, sizeof(uint32_t) * nQ, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	tm.mark("bag (device)");
	std::vector<unsigned long long> off(nQ + 1, 0);
	for (uint32_t q = 0; q < nQ; ++q)
		off[q + 1] = off[q] + bn[q];
	const unsigned long long tot = off[nQ];
	if (tot == 0)
		return RSK_OK;
	NOMEM(S.dense_a.ensure(tot));
	NOMEM(S.dense_b.ensure(tot));
	NOMEM(S.cand_t.ensure(tot));
	NOMEM(S.cand_q.ensure(tot));
	NOMEM(S.cand_s.ensure(tot));
	CK(cudaMemcpyAsync(S.bag_off.p, off.data(), sizeof(unsigned long long) * (nQ + 1), cudaMemcpyHostToDevice, st));
	PFL(pf_launch_bag_compact(S.bag_key.p, S.bag_n.p, S.bag_off.p, nQ, B, S.dense_a.p, st));
	tb = 0;
	if (pf_sort_keys64(S.dense_a.p, S.dense_b.p, tot, nullptr, tb, st))
		return fail(RSK_ERR_CUDA, "rsk_prefilter: cub sort sizing failed");
	NOMEM(S.tmp.ensure(tb));
	if (pf_sort_keys64(S.dense_a.p, S.dense_b.p, tot, S.tmp.p, tb, st))
		return fail(RSK_ERR_CUDA, "rsk_prefilter: candidate sort failed: %s", cudaGetErrorString(cudaGetLastError()));
	launches += 8;
	PFL(pf_launch_unpack_keys(S.dense_b.p, tot, S.cand_t.p, S.cand_q.p, S.cand_s.p, st));
	res.t.resize(tot); res.q.resize(tot); res.s.resize(tot);
	CK(cudaMemcpyAsync(res.t.data(), S.cand_t.p, sizeof(uint32_t) * tot, cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(res.q.data(), S.cand_q.p, sizeof(uint32_t) * tot, cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(res.s.data(), S.cand_s.p, sizeof(uint16_t) * tot, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	ctx->stats.d2h_bytes += tot * 10 + nQ * 4;
	uint32_t nt = 0;
	for (size_t k = 0; k < res.t.size(); ++k)
		nt += (k == 0 || res.t[k] != res.t[k - 1]);
	res.ntargets = nt;
	tm.mark("candidate list D2H");
	return RSK_OK;
}

}  // namespace

extern "C" int rsk_prefilter(rsk_ctx *ctx, const rsk_chainset *Q, const rsk_chainset *T, const rsk_prefilter_opts *opts_in,
		rsk_prefilter_result **out)
{
	if (!ctx || !Q || !T || !out)
		return fail(RSK_ERR_ARG, "rsk_prefilter: null argument");
	*out = nullptr;
	if (Q->ctx != ctx || T->ctx != ctx)
		return fail(RSK_ERR_ARG, "rsk_prefilter: chain sets belong to a different context");
	if (!Q->has_mu || !T->has_mu)
		return fail(RSK_ERR_ARG, "rsk_prefilter: both chain sets need Mu letters");
	rsk_prefilter_opts o = {};
	if (opts_in)
		o = *opts_in;
	const uint32_t B = o.rsb_size ? o.rsb_size : 1500u;
	uint64_t launches = 0;
	std::unique_ptr<rsk_prefilter_result> res(new rsk_prefilter_result());
	unsigned long long nraw = 0;
	int rc = prefilter_raw_device(ctx, Q, T, o, 0, nraw, launches);
	if (rc)
		return rc;
	PfScratch &S = pf_scratch(ctx);
	cudaStream_t st = ctx->stream;
	if (o.raw_only) {
		// every triple, stream order (tests, and callers that merge blocks themselves)
		res->raw = nraw;
		if (nraw) {
			NOMEM(S.cand_t.ensure(nraw));
			NOMEM(S.cand_q.ensure(nraw));
			NOMEM(S.cand_s.ensure(nraw));
			PFL(pf_launch_unpack_triples(S.raw_q.p, S.raw_v.p, nraw, S.cand_t.p, S.cand_q.p, S.cand_s.p, st));
			res->t.resize(nraw); res->q.resize(nraw); res->s.resize(nraw);
			CK(cudaMemcpyAsync(res->t.data(), S.cand_t.p, sizeof(uint32_t) * nraw, cudaMemcpyDeviceToHost, st));
			CK(cudaMemcpyAsync(res->q.data(), S.cand_q.p, sizeof(uint32_t) * nraw, cudaMemcpyDeviceToHost, st));
			CK(cudaMemcpyAsync(res->s.data(), S.cand_s.p, sizeof(uint16_t) * nraw, cudaMemcpyDeviceToHost, st));
			CK(cudaStreamSynchronize(st));
			ctx->stats.d2h_bytes += nraw * 10;
		}
		uint32_t nt = 0;
		for (size_t k = 0; k < res->t.size(); ++k)
			nt += (k == 0 || res->t[k] != res->t[k - 1]);
		res->ntargets = nt;
	} else {
		rc = bag_device(ctx, Q->d.n, B, S.raw_q.p, S.raw_v.p, nraw, *res, launches);
		if (rc)
			return rc;
	}
	ctx->stats.kernel_launches += launches;
	*out = res.release();
	return RSK_OK;
}

// The bag alone, over (target, query, score) triples in stream order: the merge step of a DB-sharded prefilter, where
// every rank produced the triples of its own target block (raw_only) and the blocks were concatenated in rank order.
extern "C" int rsk_prefilter_bag(uint32_t nq, uint64_t n, const uint32_t *t, const uint32_t *q, const uint16_t *s, uint32_t rsb_size,
		rsk_prefilter_result **out)
{
	if (!out || (n && (!t || !q || !s)))
		return fail(RSK_ERR_ARG, "rsk_prefilter_bag: null argument");
	*out = nullptr;
	const uint32_t B = rsb_size ? rsb_size : 1500u;
	if (nq >= (1u << 16))
		return fail(RSK_ERR_LIMIT, "rsk_prefilter_bag: at most 65535 queries (got %u)", nq);
	std::vector<Bag> bags(nq);
	for (uint64_t k = 0; k < n; ++k)
		if (q[k] >= nq)
			return fail(RSK_ERR_ARG, "rsk_prefilter_bag: triple %llu names query %u of %u", (unsigned long long)k, q[k], nq);
	const int nthreads = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
	feed_bags(bags, B, t, q, s, n, nthreads);
	auto *res = new rsk_prefilter_result();
	res->raw = n;
	finish_bags(bags, B, *res, nthreads);
	*out = res;
	return RSK_OK;
}

// The same bag on the device (pf_bag_kernel) over caller-supplied triples in stream order: what the sharded search runs on the
// all-gathered stream, exposed for callers that collect the triples themselves and for the parity tests against the host form.
extern "C" int rsk_prefilter_bag_device(rsk_ctx *ctx, uint32_t nq, uint64_t n, const uint32_t *t, const uint32_t *q, const uint16_t *s,
		uint32_t rsb_size, rsk_prefilter_result **out)
{
	if (!ctx || !out || (n && (!t || !q || !s)))
		return fail(RSK_ERR_ARG, "rsk_prefilter_bag_device: null argument");
	*out = nullptr;
	const uint32_t B = rsb_size ? rsb_size : 1500u;
	if (nq >= (1u << 16))
		return fail(RSK_ERR_LIMIT, "rsk_prefilter_bag_device: at most 65535 queries (got %u)", nq);
	std::vector<unsigned long long> v(n);
	for (uint64_t k = 0; k < n; ++k) {
		if (q[k] >= nq)
			return fail(RSK_ERR_ARG, "rsk_prefilter_bag_device: triple %llu names query %u of %u", (unsigned long long)k, q[k], nq);
		v[k] = ((unsigned long long)t[k] << 16) | s[k];
	}
	CK(cudaSetDevice(ctx->device));
	PfScratch &S = pf_scratch(ctx);
	NOMEM(S.raw_q.ensure(n + 1));
	NOMEM(S.raw_v.ensure(n + 1));
	if (n) {
		CK(cudaMemcpyAsync(S.raw_q.p, q, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaMemcpyAsync(S.raw_v.p, v.data(), sizeof(unsigned long long) * n, cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
	}
	std::unique_ptr<rsk_prefilter_result> res(new rsk_prefilter_result());
	uint64_t launches = 0;
	int rc = bag_device(ctx, nq, B, S.raw_q.p, S.raw_v.p, n, *res, launches);
	if (rc)
		return rc;
	ctx->stats.kernel_launches += launches;
	*out = res.release();
	return RSK_OK;
}

// Candidates whose target lies in [t_lo, t_hi), re-based to t_lo: the share of a merged candidate list that one rank of a
// DB-sharded search post-filters against its own block.
extern "C" int rsk_prefilter_select(const rsk_prefilter_result *r, uint32_t t_lo, uint32_t t_hi, rsk_prefilter_result **out)
{
	if (!r || !out || t_hi < t_lo)
		return fail(RSK_ERR_ARG, "rsk_prefilter_select: bad argument");
	auto *res = new rsk_prefilter_result();
	for (size_t k = 0; k < r->t.size(); ++k) {
		if (r->t[k] < t_lo || r->t[k] >= t_hi)
			continue;
		res->ntargets += (res->t.empty() || res->t.back() != r->t[k] - t_lo);
		res->t.push_back(r->t[k] - t_lo);
		res->q.push_back(r->q[k]);
		res->s.push_back(r->s[k]);
	}
	res->raw = res->t.size();
	*out = res;
	return RSK_OK;
}

extern "C" uint64_t rsk_prefilter_count(const rsk_prefilter_result *r) { return r ? r->t.size() : 0; }
extern "C" const uint32_t *rsk_prefilter_targets(const rsk_prefilter_result *r) { return (r && !r->t.empty()) ? r->t.data() : nullptr; }
extern "C" const uint32_t *rsk_prefilter_queries(const rsk_prefilter_result *r) { return (r && !r->q.empty()) ? r->q.data() : nullptr; }
extern "C" const uint16_t *rsk_prefilter_scores(const rsk_prefilter_result *r) { return (r && !r->s.empty()) ? r->s.data() : nullptr; }
extern "C" uint64_t rsk_prefilter_raw_count(const rsk_prefilter_result *r) { return r ? r->raw : 0; }
extern "C" void rsk_prefilter_free(rsk_prefilter_result *r) { delete r; }

extern "C" long long rsk_prefilter_to_tsv(const rsk_prefilter_result *r, char *out, size_t cap)
{
	if (!r)
		return RSK_ERR_ARG;
	std::string s = "prefilter\t" + std::to_string(r->ntargets) + "\n";
	for (size_t k = 0; k < r->t.size();) {
		size_t e = k;
		while (e < r->t.size() && r->t[e] == r->t[k])
			++e;
		s += std::to_string(r->t[k]) + "\t" + std::to_string(e - k);
		for (size_t i = k; i < e; ++i)
			s += "\t" + std::to_string(r->q[i]);
		s += "\n";
		k = e;
	}
	if (!out || cap < s.size() + 1)
		return -(long long)s.size() - 1;
	memcpy(out, s.c_str(), s.size() + 1);
	return (long long)s.size();
}

extern "C" int rsk_postfilter(rsk_ctx *ctx, const rsk_chainset *Q, const rsk_chainset *T, const rsk_prefilter_result *cands,
		const rsk_search_opts *opts, rsk_results **out)
{
	if (!ctx || !Q || !T || !cands || !out)
		return fail(RSK_ERR_ARG, "rsk_postfilter: null argument");
	// DM_AlwaysSensitive (search.cpp:106-108); everything the caller customised (tables, weights, gaps) is kept
	const rsk_params saved = ctx->params;
	rsk_params sens;
	int rc = rsk_params_preset(&sens, RSK_MODE_SENSITIVE);
	if (rc)
		return rc;
	rsk_params p = saved;
	p.omega = sens.omega; p.omega_fwd = sens.omega_fwd; p.mkfl = sens.mkfl; p.min_fwd_score = sens.min_fwd_score;
	p.mkf_x1 = sens.mkf_x1; p.mkf_x2 = sens.mkf_x2; p.mkf_min_hsp_score = sens.mkf_min_hsp_score;
	p.mkf_min_mega_hsp_score = sens.mkf_min_mega_hsp_score;  // max_evalue stays the caller's: -evalue is honoured (postmufilter.cpp:217-220)
	rc = rsk_ctx_set_params(ctx, &p);
	if (rc)
		return rc;
	// A = query bag, B = DB chain (postmufilter.cpp:190-194); line order of the TSV
	rc = rsk_search_pairs(ctx, Q, T, cands->q.size(), cands->q.data(), cands->t.data(), opts, out);
	const int rc2 = rsk_ctx_set_params(ctx, &saved);
	return rc ? rc : rc2;
}

extern "C" int rsk_search_fast_db(rsk_ctx *ctx, const rsk_chainset *Q, const rsk_chainset *T, const rsk_prefilter_opts *popts,
		const rsk_search_opts *opts, rsk_results **out)
{
	if (!out)
		return fail(RSK_ERR_ARG, "rsk_search_fast_db: null argument");
	*out = nullptr;
	rsk_prefilter_result *pf = nullptr;
	int rc = rsk_prefilter(ctx, Q, T, popts, &pf);
	if (rc)
		return rc;
	const uint64_t launches = ctx->stats.kernel_launches;
	rc = rsk_postfilter(ctx, Q, T, pf, opts, out);
	ctx->stats.kernel_launches += launches;  // the search call restarts the counters; keep the prefilter's launches in the sum
	rsk_prefilter_free(pf);
	return rc;
}

// `-search Q -db DB -fast` on a block-partitioned DB (include/reseek_b200.h).  Exchange steps: the triples of the blocks are
// all-gathered in rank order over NCCL (device to device), every rank replays the bag on the merged stream (same result on
// every rank, equal to the unsharded one), post-filters the candidates of its own block and the hits are gathered on root.
extern "C" int rsk_search_fast_db_sharded(rsk_ctx *ctx, rsk_comm *comm, const rsk_chainset *Q, const rsk_chainset *T_local, uint32_t t_base,
		const rsk_prefilter_opts *popts, const rsk_search_opts *opts, int root, rsk_results **out, rsk_prefilter_result **cands_out)
{
	if (!ctx || !Q || !out)
		return fail(RSK_ERR_ARG, "rsk_search_fast_db_sharded: null argument");
	*out = nullptr;
	if (cands_out)
		*cands_out = nullptr;
	if (comm && comm_ctx(comm) != ctx)
		return fail(RSK_ERR_ARG, "rsk_search_fast_db_sharded: the communicator belongs to a different context");
	if (Q->ctx != ctx || (T_local && T_local->ctx != ctx))
		return fail(RSK_ERR_ARG, "rsk_search_fast_db_sharded: chain sets belong to a different context");
	if (!Q->has_mu || (T_local && !T_local->has_mu))
		return fail(RSK_ERR_ARG, "rsk_search_fast_db_sharded: both chain sets need Mu letters");
	rsk_prefilter_opts o = {};
	if (popts)
		o = *popts;
	const uint32_t B = o.rsb_size ? o.rsb_size : 1500u;
	const uint32_t nT = T_local ? T_local->d.n : 0;
	uint64_t launches = 0;
	unsigned long long nraw = 0;
	int rc = prefilter_raw_device(ctx, Q, T_local, o, t_base, nraw, launches);
	if (rc)
		return rc;
	PfScratch &S = pf_scratch(ctx);
	std::unique_ptr<rsk_prefilter_result> merged(new rsk_prefilter_result());
	const int N = comm_nranks(comm);
	if (N == 1) {
		rc = bag_device(ctx, Q->d.n, B, S.raw_q.p, S.raw_v.p, nraw, *merged, launches);
	} else {
		const unsigned long long mine[kCommCountWords] = {nraw * sizeof(uint32_t), nraw * sizeof(unsigned long long), 0, 0};
		const unsigned long long *all = nullptr;
		if ((rc = comm_exchange_counts(comm, mine, &all)))
			return rc;
		std::vector<unsigned long long> bytes((size_t)N * 2);
		for (int r = 0; r < N; ++r) {
			bytes[(size_t)r * 2 + 0] = all[(size_t)r * kCommCountWords + 0];
			bytes[(size_t)r * 2 + 1] = all[(size_t)r * kCommCountWords + 1];
		}
		const void *src[2] = {S.raw_q.p, S.raw_v.p};
		unsigned char *g[2] = {nullptr, nullptr};
		unsigned long long gtot[2] = {0, 0};
		if ((rc = comm_allgather_parts(comm, 2, src, bytes.data(), g, gtot)))
			return rc;
		rc = bag_device(ctx, Q->d.n, B, (const uint32_t *)g[0], (const unsigned long long *)g[1], gtot[0] / sizeof(uint32_t), *merged, launches);
	}
	if (rc)
		return rc;
	// this rank's share of the candidate lines, targets re-based to the block
	std::vector<uint32_t> lq, lt;
	for (size_t k = 0; k < merged->t.size(); ++k)
		if (merged->t[k] >= t_base && merged->t[k] - t_base < nT) {
			lq.push_back(merged->q[k]);
			lt.push_back(merged->t[k] - t_base);
		}
	// DM_AlwaysSensitive (search.cpp:106-108), as in rsk_postfilter
	const rsk_params saved = ctx->params;
	rsk_params sens;
	if ((rc = rsk_params_preset(&sens, RSK_MODE_SENSITIVE)))
		return rc;
	rsk_params p = saved;
	p.omega = sens.omega; p.omega_fwd = sens.omega_fwd; p.mkfl = sens.mkfl; p.min_fwd_score = sens.min_fwd_score;
	p.mkf_x1 = sens.mkf_x1; p.mkf_x2 = sens.mkf_x2; p.mkf_min_hsp_score = sens.mkf_min_hsp_score;
	p.mkf_min_mega_hsp_score = sens.mkf_min_mega_hsp_score;
	if ((rc = rsk_ctx_set_params(ctx, &p)))
		return rc;
	rc = search_pairs_sharded(ctx, comm, Q, T_local, lq.size(), lq.data(), lt.data(), 0, t_base, opts, root, out);
	const int rc2 = rsk_ctx_set_params(ctx, &saved);
	ctx->stats.kernel_launches += launches;
	if (rc || rc2) {
		if (*out) {
			rsk_results_free(*out);
			*out = nullptr;
		}
		return rc ? rc : rc2;
	}
	if (cands_out)
		*cands_out = merged.release();
	return RSK_OK;
}
