// prefilter_kernel.cu - K6..K8: the `-fast -db` Mu 5-mer prefilter on the GPU.
//
// Replaces MuPreFilter (muprefilter.cpp:64-133): MuDex::FromSeqDB incl. the k-mer neighbourhoods
// (mudex.cpp:386-442, :130-180; MerMx::GetHighScoring5mers mermx.cpp:484-584), PrefilterMu::Search
// (prefiltermu.cpp:382-393) = Search_TargetKmer (:213-261) + TwoHitDiag::Add/SetDupes (twohitdiag.cpp:47/389)
// + FindHSP (:12-48) + AddTwoHitDiag (:288-313).  The per-query top-B bag (RankedScoresBag) is order dependent and
// tiny, so it runs on the host (rsk_api.cu) over the (target, query, score) triples this file produces.
//
//   K6  query index: every unmasked query 5-mer X and all 5-mers Y with pair score >= 36 (its neighbourhood, which
//       contains X itself; in query-neighbourhood mode X is entered a second time, exactly like the reference's
//       Put + neighbourhood loop) -> (key = Y, value = query<<16 | position) pairs, radix-sorted by key, with a
//       dense row table over the 36^5 dictionary.
//   K7  probe: one CTA per target; every unmasked target 5-mer reads its index row (coalesced 4-byte values) and
//       emits key = query<<14 | diagonal for every hit (diagonals > 16383 dropped, prefiltermu.cpp:254).
//       The keys of a target are sorted (CUB segmented radix sort); a key that occurs twice is a two-hit diagonal.
//   K8  extend: one thread per two-hit diagonal runs the reference's Kadane scan over the whole diagonal
//       (int adds in diagonal order) and atomicMax-es the per-(target, query) best score.
//
// Bound: HBM/L2 latency of the index probe (random 8-byte row lookups + short coalesced rows); algorithmic bytes per
// target k-mer: 8 (row start/end) + 4*rowsize, per hit 4 B written + sorted, per diagonal 2*len letter bytes.
#include <cub/cub.cuh>

#include "rsk_internal.cuh"

namespace rsk {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kMinPair = 36;      // MIN_KMER_PAIR_SCORE, also the self-score mask (prefiltermuparams.h)
constexpr uint32_t kMasked = 0xffffffffu;

__device__ __forceinline__ uint32_t kmer5(const uint8_t *w, const int *S, bool &masked)
{
	// spaced pattern 1110011: offsets 0,1,2,5,6 (prefiltermuparams.h:7-10), base-36 big-endian (mudex.cpp:517-538)
	const int o[5] = {0, 1, 2, 5, 6};
	uint32_t k = 0;
	int self = 0;
#pragma unroll
	for (int c = 0; c < 5; ++c) {
		const int x = w[o[c]];
		k = k * 36 + x;
		self += S[36 * x + x];
	}
	masked = self < kMinPair;
	return k;
}

// ---- K6a: list the query 5-mers (position-major per query) ----
__global__ void pf_query_kmers_kernel(const PfArgs a)
{
	__shared__ int S[36 * 36];
	for (int k = threadIdx.x; k < 36 * 36; k += blockDim.x)
		S[k] = a.kmer_mx[k];
	__syncthreads();
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= a.nqk)
		return;
	// binary search the query that owns global k-mer slot i
	uint32_t lo = 0, hi = a.nQ;
	while (hi - lo > 1) {
		const uint32_t mid = (lo + hi) >> 1;
		if (a.qk_off[mid] <= i) lo = mid; else hi = mid;
	}
	const uint32_t q = lo, pos = i - a.qk_off[q];
	bool masked;
	const uint32_t k = kmer5(a.muQ + a.offQ[q] + pos, S, masked);
	a.qk_code[i] = masked ? kMasked : k;
	a.qk_val[i] = (q << 16) | pos;
}

// ---- K6b/c: neighbourhoods.  One warp per query 5-mer; lanes split the first two letters, DFS with bounds. ----
template <bool FILL>
__global__ void __launch_bounds__(128) pf_neighborhood_kernel(const PfArgs a)
{
	__shared__ int S[36 * 36];
	__shared__ int rowmax[36];
	__shared__ unsigned s_cursor[4];
	for (int k = threadIdx.x; k < 36 * 36; k += blockDim.x)
		S[k] = a.kmer_mx[k];
	__syncthreads();
	if (threadIdx.x < 36) {
		int m = -128;
		for (int y = 0; y < 36; ++y)
			m = max(m, S[36 * threadIdx.x + y]);
		rowmax[threadIdx.x] = m;
	}
	__syncthreads();
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t i = blockIdx.x * 4 + warp;
	if (i >= a.nqk)
		return;
	const uint32_t code = a.qk_code[i];
	if (code == kMasked) {
		if (!FILL && lane == 0)
			a.nb_count[i] = 0;
		return;
	}
	int x[5];
	{
		uint32_t c = code;
		for (int k = 4; k >= 0; --k) { x[k] = (int)(c % 36); c /= 36; }
	}
	const int *r0 = S + 36 * x[0], *r1 = S + 36 * x[1], *r2 = S + 36 * x[2], *r3 = S + 36 * x[3], *r4 = S + 36 * x[4];
	const int m4 = rowmax[x[4]], m34 = rowmax[x[3]] + m4, m234 = rowmax[x[2]] + m34;
	const uint32_t val = a.qk_val[i];
	unsigned long long base = 0;
	if (FILL) {
		base = a.nb_off[i];
		if (lane == 0)
			s_cursor[warp] = 0;
		__syncwarp();
	}
	unsigned cnt = 0;
	for (int yy = lane; yy < 36 * 36; yy += 32) {
		const int y0 = yy / 36, y1 = yy - 36 * y0;
		const int p2 = r0[y0] + r1[y1];
		if (p2 + m234 < kMinPair)
			continue;
		for (int y2 = 0; y2 < 36; ++y2) {
			const int p3 = p2 + r2[y2];
			if (p3 + m34 < kMinPair)
				continue;
			for (int y3 = 0; y3 < 36; ++y3) {
				const int p4 = p3 + r3[y3];
				if (p4 + m4 < kMinPair)
					continue;
				for (int y4 = 0; y4 < 36; ++y4) {
					if (p4 + r4[y4] >= kMinPair) {
						if (FILL) {
							const unsigned slot = atomicAdd(&s_cursor[warp], 1u);
							const uint32_t y = (((uint32_t)(y0 * 36 + y1) * 36 + y2) * 36 + y3) * 36 + y4;
							a.ix_key[base + slot] = y;
							a.ix_val[base + slot] = val;
						} else {
							++cnt;
						}
					}
				}
			}
		}
	}
	if (FILL) {
		__syncwarp();
		if (a.exact_twice && lane == 0) {  // the k-mer itself, entered before its neighbourhood (mudex.cpp:146-174)
			const unsigned slot = s_cursor[warp];
			a.ix_key[base + slot] = code;
			a.ix_val[base + slot] = val;
		}
	} else {
#pragma unroll
		for (int o = 16; o >= 1; o >>= 1)
			cnt += __shfl_xor_sync(kFull, cnt, o);
		if (lane == 0)
			a.nb_count[i] = cnt + (a.exact_twice ? 1u : 0u);
	}
}

// ---- K6d: dense row table over the dictionary from the sorted keys ----
__global__ void pf_mark_rows_kernel(const uint32_t *__restrict__ key, unsigned long long n, uint32_t *row_start, uint32_t *row_end)
{
	const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	const uint32_t k = key[i];
	if (i == 0 || key[i - 1] != k)
		row_start[k] = (uint32_t)i;
	if (i + 1 == n || key[i + 1] != k)
		row_end[k] = (uint32_t)(i + 1);
}

// ---- K7: probe.  COUNT pass sizes each target's hit segment, FILL pass writes the (query, diagonal) keys. ----
template <bool FILL>
__global__ void __launch_bounds__(128) pf_probe_kernel(const PfArgs a)
{
	__shared__ int S[36 * 36];
	__shared__ unsigned long long s_cursor;
	for (int k = threadIdx.x; k < 36 * 36; k += blockDim.x)
		S[k] = a.kmer_mx[k];
	if (threadIdx.x == 0)
		s_cursor = 0;
	__syncthreads();
	const uint32_t tl = blockIdx.x;  // target within the batch
	const uint32_t t = a.t_begin + tl;
	const uint32_t LT = a.lenT[t];
	const uint8_t *T = a.muT + a.offT[t];
	unsigned long long cnt = 0;
	const unsigned long long seg = FILL ? a.hit_off[tl] : 0;
	if (LT >= 7) {
		for (uint32_t tpos = threadIdx.x; tpos + 7 <= LT; tpos += blockDim.x) {
			bool masked;
			const uint32_t y = kmer5(T + tpos, S, masked);
			if (masked)
				continue;
			const uint32_t rs = a.row_start[y], re = a.row_end[y];
			for (uint32_t e = rs; e < re; ++e) {
				const uint32_t v = a.ix_val[e];
				const uint32_t q = v >> 16, qpos = v & 0xffffu;
				const uint32_t diag = a.lenQ[q] + tpos - qpos - 1;  // diag.h:22-25
				if (diag > 0x3fffu)
					continue;  // prefiltermu.cpp:254
				if (FILL) {
					const unsigned long long slot = atomicAdd(&s_cursor, 1ull);
					a.hit_key[seg + slot] = (q << 14) | diag;
				} else {
					++cnt;
				}
			}
		}
	}
	if (!FILL) {
		typedef cub::BlockReduce<unsigned long long, 128> BR;
		__shared__ typename BR::TempStorage tmp;
		const unsigned long long tot = BR(tmp).Sum(cnt);
		if (threadIdx.x == 0)
			a.hit_count[tl] = tot;
	}
}

// ---- K8: two-hit diagonals -> FindHSP -> best score per (target, query) ----
// Two phases per CTA (= per target): the sorted hit keys are scanned for the first element of every run of length >= 2 (a
// two-hit diagonal) and those are queued in shared memory; then every thread of the CTA pulls diagonals from the queue and
// runs the Kadane scan.  Scanning and walking in one loop left most threads idle (only run starts walk).
constexpr int kPfQueue = 2048;
__global__ void __launch_bounds__(128) pf_extend_kernel(const PfArgs a)
{
	__shared__ int S[36 * 36];
	__shared__ uint32_t s_queue[kPfQueue];
	__shared__ unsigned s_n, s_next;
	for (int k = threadIdx.x; k < 36 * 36; k += blockDim.x)
		S[k] = a.kmer_mx[k];
	const uint32_t tl = blockIdx.x;
	const uint32_t t = a.t_begin + tl;
	const unsigned long long s0 = a.hit_off[tl], s1 = a.hit_off[tl + 1];
	const uint32_t LT = a.lenT[t];
	const uint8_t *T = a.muT + a.offT[t];
	// the scan advances in windows; a window ends early when the queue is full
	unsigned long long base = s0;
	while (base + 1 < s1) {
		if (threadIdx.x == 0) { s_n = 0; s_next = 0; }
		__syncthreads();
		const unsigned long long wend = min(s1 - 1, base + (unsigned long long)kPfQueue);  // at most kPfQueue candidates per window
		for (unsigned long long i = base + threadIdx.x; i < wend; i += blockDim.x) {
			const uint32_t k = a.hit_sorted[i];
			// first element of a run of length >= 2
			if (a.hit_sorted[i + 1] == k && !(i > s0 && a.hit_sorted[i - 1] == k))
				s_queue[atomicAdd(&s_n, 1u)] = k;
		}
		__syncthreads();
		const unsigned n = s_n;
		for (;;) {
			const unsigned w = atomicAdd(&s_next, 1u);
			if (w >= n)
				break;
			const uint32_t k = s_queue[w];
			const uint32_t q = k >> 14;
			const int d = (int)(k & 0x3fffu);
			const uint32_t LQ = a.lenQ[q];
			const uint8_t *Q = a.muQ + a.offQ[q];
			int qi = (int)LQ - d - 1, tj = 0;
			if (qi < 0) { tj = -qi; qi = 0; }
			int B = 0, F = 0;
			for (; qi < (int)LQ && tj < (int)LT; ++qi, ++tj) {  // prefiltermu.cpp:27-46
				F += S[36 * Q[qi] + T[tj]];
				if (F > B) B = F;
				else if (F < 0) F = 0;
			}
			if (B > 0) {
				if (B >= 65535) B = 65534;  // prefiltermu.cpp:294-295
				atomicMax(&a.best[(size_t)tl * a.nQ + q], (unsigned)B);
			}
		}
		__syncthreads();
		base = wend;
	}
}

// ---- compaction of best > 0 into (target, query, score) triples, target-major ----
__global__ void pf_count_cands_kernel(const PfArgs a, uint32_t ntl)
{
	const uint32_t tl = blockIdx.x * blockDim.x + threadIdx.x;
	if (tl >= ntl)
		return;
	uint32_t n = 0;
	for (uint32_t q = 0; q < a.nQ; ++q)
		n += a.best[(size_t)tl * a.nQ + q] != 0;
	a.cand_count[tl] = n;
}

__global__ void pf_write_cands_kernel(const PfArgs a, uint32_t ntl)
{
	const uint32_t tl = blockIdx.x * blockDim.x + threadIdx.x;
	if (tl >= ntl)
		return;
	unsigned long long o = a.cand_off[tl];
	for (uint32_t q = 0; q < a.nQ; ++q) {
		const unsigned b = a.best[(size_t)tl * a.nQ + q];
		if (b) {
			a.cand_t[o] = a.t_begin + tl;
			a.cand_q[o] = q;
			a.cand_s[o] = (uint16_t)b;
			++o;
		}
	}
}

// K/L swap of the query letters (the query side goes through g_CharToLetterMu, alpha.cpp:3291)
__global__ void pf_swap_kl_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, uint64_t n)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	const uint8_t c = in[i];
	out[i] = c == 10 ? 11 : c == 11 ? 10 : c;
}

}  // namespace

int pf_launch_swap_kl(const uint8_t *in, uint8_t *out, uint64_t n, cudaStream_t st)
{
	if (n == 0)
		return 0;
	pf_swap_kl_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, n);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int pf_launch_query_kmers(const PfArgs &a, cudaStream_t st)
{
	if (a.nqk == 0)
		return 0;
	pf_query_kmers_kernel<<<(a.nqk + 255) / 256, 256, 0, st>>>(a);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int pf_launch_neighborhood(const PfArgs &a, bool fill, cudaStream_t st)
{
	if (a.nqk == 0)
		return 0;
	if (fill)
		pf_neighborhood_kernel<true><<<(a.nqk + 3) / 4, 128, 0, st>>>(a);
	else
		pf_neighborhood_kernel<false><<<(a.nqk + 3) / 4, 128, 0, st>>>(a);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int pf_launch_mark_rows(const uint32_t *key, unsigned long long n, uint32_t *row_start, uint32_t *row_end, cudaStream_t st)
{
	if (n == 0)
		return 0;
	pf_mark_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(key, n, row_start, row_end);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int pf_launch_probe(const PfArgs &a, uint32_t ntl, bool fill, cudaStream_t st)
{
	if (ntl == 0)
		return 0;
	if (fill)
		pf_probe_kernel<true><<<ntl, 128, 0, st>>>(a);
	else
		pf_probe_kernel<false><<<ntl, 128, 0, st>>>(a);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int pf_launch_extend(const PfArgs &a, uint32_t ntl, cudaStream_t st)
{
	if (ntl == 0)
		return 0;
	pf_extend_kernel<<<ntl, 128, 0, st>>>(a);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int pf_launch_cands(const PfArgs &a, uint32_t ntl, bool write, cudaStream_t st)
{
	if (ntl == 0)
		return 0;
	if (write)
		pf_write_cands_kernel<<<(ntl + 127) / 128, 128, 0, st>>>(a, ntl);
	else
		pf_count_cands_kernel<<<(ntl + 127) / 128, 128, 0, st>>>(a, ntl);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ---- CUB plumbing (sorts); temp storage is provided by the caller ----
int pf_sort_pairs(const uint32_t *kin, uint32_t *kout, const uint32_t *vin, uint32_t *vout, unsigned long long n, void *tmp,
		size_t &tmp_bytes, cudaStream_t st)
{
	return cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kin, kout, vin, vout, (long long)n, 0, 26, st) == cudaSuccess ? 0 : -1;
}

int pf_segmented_sort(const uint32_t *kin, uint32_t *kout, unsigned long long n, uint32_t nseg, const unsigned long long *off,
		void *tmp, size_t &tmp_bytes, cudaStream_t st)
{
	return cub::DeviceSegmentedSort::SortKeys(tmp, tmp_bytes, kin, kout, (long long)n, (long long)nseg, off, off + 1, st) == cudaSuccess ? 0 : -1;
}

}  // namespace rsk
