// prefilter_kernel.cu - K6..K8: the `-fast -db` Mu 5-mer prefilter on the GPU.
//
// Replaces MuPreFilter (muprefilter.cpp:64-133): MuDex::FromSeqDB incl. the k-mer neighbourhoods
// (mudex.cpp:386-442, :130-180; MerMx::GetHighScoring5mers mermx.cpp:484-584), PrefilterMu::Search
// (prefiltermu.cpp:382-393) = Search_TargetKmer (:213-261) + TwoHitDiag::Add/SetDupes (twohitdiag.cpp:47/389)
// + FindHSP (:12-48) + AddTwoHitDiag (:288-313), and RankedScoresBag (rankedscoresbag.cpp:5-51, 185-232) - the per-query top-B
// bag, order dependent down to the swap sequence of the reference's unstable quicksort - as pf_bag_kernel.
//
//   K6  query index: every unmasked query 5-mer X and all 5-mers Y with pair score >= 36 (its neighbourhood, which
//       contains X itself; in query-neighbourhood mode X is entered a second time, exactly like the reference's
//       Put + neighbourhood loop) -> (key = Y, value = query<<16 | position) pairs, radix-sorted by key, with a
//       dense row table over the 36^5 dictionary.  The neighbourhood is enumerated work-efficiently (sorted letter lists +
//       count tables: no candidate is tested and rejected).
//   K7+K8 fused (pf_probe_extend*_kernel): one CTA per target; every unmasked target 5-mer reads its index row; a hit sets a bit
//       per (query, diagonal) in shared-memory bitmaps (diagonals > 16383 dropped, prefiltermu.cpp:254), the second hit on a
//       diagonal queues it, and the CTA walks the queued diagonals with the reference's Kadane scan (int adds in diagonal
//       order), atomicMax-ing the per-(target, query) best score.  Targets whose bitmaps do not fit shared memory take the
//       global-memory form: K7 probe -> key = query<<14 | diagonal per hit, CUB segmented sort, K8 extend from the runs.
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

template <typename TS>
__device__ __forceinline__ uint32_t kmer5(const uint8_t *w, const TS *S, bool &masked)
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

// ---- K6b/c: neighbourhoods.  One warp per query 5-mer X: every 5-mer Y with score(X, Y) >= 36. ----
// Branch and bound over the letters of Y, breadth first and warp-wide, and work-efficient: per letter x the host supplies the 36
// letters y sorted by S[x][y] descending, those scores, and ge[t] = #{y : S[x][y] >= t}.  A prefix of depth k with partial score p
// has exactly ge_k[36 - p - (best achievable remainder)] children with a completion, and they are the first that many of the
// sorted list - so a level is expanded by (1) one table look-up per parent, (2) a warp prefix sum, (3) rounds of 32 children,
// each lane finding its parent by binary search over the 32 prefix sums.  No candidate is ever tested and rejected; the earlier
// form scanned all 36^3 three-letter prefixes per k-mer and tested 36 letters per queued prefix (11 % of them hits).
// Queues per level live in shared memory; a full queue is expanded (recursively, depth first) before its producer continues.
// The order of the neighbours inside one k-mer's block is arbitrary: the index is sorted by neighbour code afterwards, and
// within one block all codes are distinct (the k-mer's own second entry in query-neighbourhood mode carries the same value).
constexpr int kNbWarps = 4, kNbCap = 256, kNbCap1 = 64;
constexpr int kNbRow = 136;  // per letter: 36 sorted letters, 36 sorted scores (int8), 64 ge counts (threshold + 32)
constexpr int kNbWarpWords = 32 + kNbCap1 + 3 * kNbCap + 5 * 64;  // Q0 | Q1 | Q2..Q4 | per level: 32 prefix sums + 32 parents

struct NbState {
	const uint8_t *tab[5];  // letter tables of x0..x4 (shared memory)
	int rem[5];             // best achievable score of the letters after position k
	uint32_t *Q[5];         // entries: prefix code << 8 | (partial score + 128)
	uint32_t *scratch;
	unsigned n[5];
	unsigned acc;           // counting pass: this lane's share of the neighbourhood size
	unsigned nout;          // fill pass: neighbours written so far
	unsigned long long base;
	uint32_t val;
	uint32_t *ix_key, *ix_val;
};

template <int K, bool FILL>
__device__ __forceinline__ void nb_expand(NbState &s, const unsigned lane)
{
	const uint8_t *tb = s.tab[K];
	const uint32_t *Qk = s.Q[K];
	const unsigned nk = s.n[K];
	s.n[K] = 0;
	const int need = kMinPair - s.rem[K] + 32;  // ge index of a parent with partial score p: need - p
	if (K == 4 && !FILL) {
		for (unsigned e = lane; e < nk; e += 32)
			s.acc += tb[72 + min(max(need - ((int)(Qk[e] & 0xffu) - 128), 0), 63)];
		__syncwarp();
		return;
	}
	uint32_t *pre = s.scratch + 64 * K, *par = pre + 32;
	unsigned e0 = 0, total = 0, o0 = 0;  // next chunk of parents, children of the current chunk, next child
	for (;;) {
		bool done = false;
		while (o0 >= total) {
			if (e0 >= nk) {
				done = true;
				break;
			}
			__syncwarp();  // the previous rounds are through with pre/par
			const unsigned e = e0 + lane;
			uint32_t q = 0;
			unsigned cnt = 0;
			if (e < nk) {
				q = Qk[e];
				cnt = tb[72 + min(max(need - ((int)(q & 0xffu) - 128), 0), 63)];
			}
			unsigned incl = cnt;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const unsigned t = __shfl_up_sync(kFull, incl, o);
				if ((int)lane >= o)
					incl += t;
			}
			pre[lane] = incl;
			par[lane] = q;
			__syncwarp();
			total = __shfl_sync(kFull, incl, 31);
			o0 = 0;
			e0 += 32;
		}
		if (!done) {
			const unsigned o = o0 + lane;
			const unsigned nround = min(32u, total - o0);
			if (lane < nround) {
				unsigned j = 0;  // number of parents whose children all come before child o
#pragma unroll
				for (int step = 16; step >= 1; step >>= 1)
					if (pre[j + step - 1] <= o)
						j += step;
				const unsigned rank = o - (j ? pre[j - 1] : 0u);
				const uint32_t q = par[j];
				const uint32_t code = (q >> 8) * 36u + tb[rank];
				if (K == 4) {
					s.ix_key[s.base + s.nout + lane] = code;
					s.ix_val[s.base + s.nout + lane] = s.val;
				} else {
					s.Q[K < 4 ? K + 1 : 4][s.n[K < 4 ? K + 1 : 4] + lane] = (code << 8) | (uint32_t)((int)(q & 0xffu) + (int)(int8_t)tb[36 + rank]);
				}
			}
			if (K == 4)
				s.nout += nround;
			else
				s.n[K < 4 ? K + 1 : 4] += nround;
			o0 += 32;
		}
		if constexpr (K < 4) {
			if ((done && s.n[K + 1]) || s.n[K + 1] + 32 > (unsigned)(K + 1 == 1 ? kNbCap1 : kNbCap)) {
				__syncwarp();
				nb_expand<K + 1, FILL>(s, lane);
			}
		}
		if (done)
			break;
	}
	__syncwarp();
}

template <bool FILL>
__global__ void __launch_bounds__(kNbWarps * 32) pf_neighborhood_kernel(const PfArgs a)
{
	__shared__ __align__(16) uint8_t s_tab[36 * kNbRow];
	__shared__ uint32_t s_warp[kNbWarps][kNbWarpWords];
	for (int k = threadIdx.x; k < 36 * kNbRow / 4; k += blockDim.x)
		reinterpret_cast<uint32_t *>(s_tab)[k] = reinterpret_cast<const uint32_t *>(a.nb_tab)[k];
	__syncthreads();
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t i = blockIdx.x * kNbWarps + warp;
	if (i >= a.nqk)
		return;
	const uint32_t code = a.qk_code[i];
	if (code == kMasked) {
		if (!FILL && lane == 0)
			a.nb_count[i] = 0;
		return;
	}
	NbState s;
	{
		uint32_t c = code;
		int x[5];
		for (int k = 4; k >= 0; --k) { x[k] = (int)(c % 36); c /= 36; }
		int rem = 0;
		for (int k = 4; k >= 0; --k) {
			s.tab[k] = s_tab + kNbRow * x[k];
			s.rem[k] = rem;
			rem += (int)(int8_t)s.tab[k][36];  // the best score of row x[k]
		}
	}
	uint32_t *w = s_warp[warp];
	s.Q[0] = w;
	s.Q[1] = w + 32;
	s.Q[2] = s.Q[1] + kNbCap1;
	s.Q[3] = s.Q[2] + kNbCap;
	s.Q[4] = s.Q[3] + kNbCap;
	s.scratch = s.Q[4] + kNbCap;
	for (int k = 0; k < 5; ++k)
		s.n[k] = 0;
	s.acc = 0;
	s.nout = 0;
	s.base = FILL ? a.nb_off[i] : 0;
	s.val = a.qk_val[i];
	s.ix_key = a.ix_key;
	s.ix_val = a.ix_val;
	if (lane == 0)
		s.Q[0][0] = 128u;  // the empty prefix, score 0
	s.n[0] = 1;
	__syncwarp();
	nb_expand<0, FILL>(s, lane);
	if (FILL) {
		if (a.exact_twice && lane == 0) {  // the k-mer itself is entered a second time (mudex.cpp:146-174)
			a.ix_key[s.base + s.nout] = code;
			a.ix_val[s.base + s.nout] = s.val;
		}
	} else {
		unsigned acc = s.acc;
#pragma unroll
		for (int o = 16; o >= 1; o >>= 1)
			acc += __shfl_xor_sync(kFull, acc, o);
		if (lane == 0)
			a.nb_count[i] = acc + (a.exact_twice ? 1u : 0u);
	}
}

// ---- K6d: dense row table over the dictionary from the sorted keys ----
__global__ void pf_mark_rows_kernel(const uint32_t *__restrict__ key, unsigned long long n, uint2 *row)
{
	const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	const uint32_t k = key[i];
	if (i == 0 || key[i - 1] != k)
		row[k].x = (uint32_t)i;
	if (i + 1 == n || key[i + 1] != k)
		row[k].y = (uint32_t)(i + 1);
}

// ---- K7: probe.  COUNT pass sizes each target's hit segment, FILL pass writes the (query, diagonal) keys. ----
// One warp per target position: the lanes read the index row of its 5-mer side by side (coalesced 4-byte values); the row
// bounds of the warp's NEXT position are fetched before the current row is walked, so the two dependent misses of a position
// (row bounds, then row entries) overlap with the previous row.
template <typename TS>
__device__ __forceinline__ uint2 probe_row(const PfArgs &a, const uint8_t *T, uint32_t LT, uint32_t tpos, const TS *S)
{
	if (tpos + 7 > LT)
		return make_uint2(0, 0);
	bool masked;
	const uint32_t y = kmer5(T + tpos, S, masked);
	return masked ? make_uint2(0, 0) : __ldg(a.row + y);
}

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
	const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
	unsigned long long cnt = 0;
	const unsigned long long seg = FILL ? a.hit_off[tl] : 0;
	if (FILL && a.hit_off[tl + 1] == seg)
		return;  // no hits, or a target the fused kernel has taken
	if (LT >= 7) {
		uint2 r = probe_row(a, T, LT, warp, S);
		for (uint32_t tpos = warp; tpos + 7 <= LT; tpos += nwarps) {
			const uint2 rn = probe_row(a, T, LT, tpos + nwarps, S);
			if (!FILL && a.diag_safe) {
				if (lane == 0)
					cnt += r.y - r.x;  // no diagonal can exceed 16383: the row length is the count
			} else {
				for (uint32_t e = r.x + lane; e < r.y; e += 32) {
					const uint32_t v = __ldg(a.ix_val + e);
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
			r = rn;
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

// FindHSP (prefiltermu.cpp:12-48) over the whole diagonal `d` of (query q, target T): Kadane scan in diagonal order, the best
// score goes to best[tl][q].  The reference's `F += s; if (F > B) B = F; else if (F < 0) F = 0` is F = max(F + s, 0),
// B = max(B, F) - two DPX instructions (VIADDMNMX, VIMNMX).  Four positions per trip: the letters of both chains come as 32-bit
// words assembled from aligned loads (one new word per chain and trip), the scores from an int8 table in shared memory:
// per position two byte extracts, one IMAD, one LDS, two DPX (the first form - int table, compare-and-select - ran 16).
// SHARED: the letters of both chains were staged in shared memory (plain loads); otherwise they come from global memory.
template <bool SHARED>
__device__ __forceinline__ int walk_core(const int8_t *S, const uint8_t *Q, const uint32_t LQ, const uint8_t *T, const uint32_t LT, const int d)
{
	int qi = (int)LQ - d - 1, tj = 0;
	if (qi < 0) { tj = -qi; qi = 0; }
	int B = 0, F = 0;
	int n = min((int)LQ - qi, (int)LT - tj);
	const uint8_t *qp = Q + qi, *tp = T + tj;
	if (n >= 8) {
		const uint32_t *qw = reinterpret_cast<const uint32_t *>((uintptr_t)qp & ~(uintptr_t)3);
		const uint32_t *tw = reinterpret_cast<const uint32_t *>((uintptr_t)tp & ~(uintptr_t)3);
		const unsigned qs = ((uintptr_t)qp & 3u) * 8u, ts = ((uintptr_t)tp & 3u) * 8u;
		uint32_t q0 = SHARED ? *qw : __ldg(qw), t0 = SHARED ? *tw : __ldg(tw);
		const int groups = n >> 2;
		for (int g = 0; g < groups; ++g) {
			// the next aligned word holds at least one letter of this group whenever the chain's pointer is unaligned
			uint32_t q1 = q0, t1 = t0;
			if (qs || g + 1 < groups) { ++qw; q1 = SHARED ? *qw : __ldg(qw); }
			if (ts || g + 1 < groups) { ++tw; t1 = SHARED ? *tw : __ldg(tw); }
			const uint32_t qv = __funnelshift_r(q0, q1, qs);
			const uint32_t tv = __funnelshift_r(t0, t1, ts);
			q0 = q1; t0 = t1;
#pragma unroll
			for (int b = 0; b < 4; ++b) {
				const int sc = S[36u * __byte_perm(qv, 0, 0x4440 + b) + __byte_perm(tv, 0, 0x4440 + b)];
				F = __viaddmax_s32(F, sc, 0);
				B = max(B, F);
			}
		}
		qp += 4 * groups; tp += 4 * groups;
		n -= 4 * groups;
	}
	for (int i = 0; i < n; ++i) {
		F = __viaddmax_s32(F, (int)S[36 * qp[i] + tp[i]], 0);
		B = max(B, F);
	}
	return B;
}

__device__ __forceinline__ void walk_diagonal(const PfArgs &a, const int8_t *S, const uint8_t *T, uint32_t LT, uint32_t tl, uint32_t k)
{
	const uint32_t q = k >> 14;
	int B = walk_core<false>(S, a.muQ + a.offQ[q], a.lenQ[q], T, LT, (int)(k & 0x3fffu));
	if (B > 0) {
		if (B >= 65535) B = 65534;  // prefiltermu.cpp:294-295
		atomicMax(&a.best[(size_t)tl * a.nQ + q], (unsigned)B);
	}
}

// ---- K7+K8 fused: the hits of a target never leave the SM.  A target sees sum(LQ) + nQ * (LT - 1) distinct (query, diagonal)
// pairs; with a block of 100 queries that is ~45 000 for a 300-residue target, so two BITMAPS over them fit in shared memory:
// `seen` (a hit fell on the diagonal) and `twice` (a second one did - TwoHitDiag::SetDupes, twohitdiag.cpp:389).  The hit that
// sets `twice` queues the diagonal, once; then the CTA walks the queued diagonals, clearing their `twice` bits.  Diagonals that
// found the queue full keep their bit and are collected from the bitmap afterwards, a queue-full at a time.  Which thread sees
// the second hit is timing dependent, the SET of walked diagonals is not, and best[] is a max.  Against the global-memory path
// (4 B/hit key store, segmented radix sort, scan for runs) this saves the sort and ~1.2 GB of traffic per 100 x 20 000 block.
// Two size classes by bitmap size keep the short targets at high occupancy.
constexpr uint32_t kBmSmallBits = 1u << 16, kBmLargeBits = 3u << 17;  // 2 x 8 KB and 2 x 48 KB of bitmaps
constexpr uint32_t kBmSmallQueue = 4096, kBmLargeQueue = 8192;
constexpr size_t bm_smem_bytes(uint32_t bits, uint32_t queue) { return (size_t)bits / 8 * 2 + sizeof(uint32_t) * queue; }

template <uint32_t MINBITS, uint32_t MAXBITS, uint32_t QUEUE, int THREADS>
__global__ void __launch_bounds__(THREADS) pf_probe_extend_kernel(const PfArgs a)
{
	extern __shared__ __align__(16) uint32_t bm_smem[];
	__shared__ int8_t S[36 * 36];
	__shared__ unsigned s_n, s_next;
	const uint32_t tl = blockIdx.x;
	const uint32_t t = a.t_begin + tl;
	const uint32_t LT = a.lenT[t];
	if (LT < 7)
		return;
	const unsigned long long nbits = (unsigned long long)a.sum_lenQ + (unsigned long long)a.nQ * (LT - 1);
	if (nbits <= MINBITS || nbits > MAXBITS)
		return;  // the other size class, or the global-memory path
	const uint32_t words = ((uint32_t)nbits + 31) / 32;
	const uint32_t qcap = a.queue_cap ? min(a.queue_cap, QUEUE) : QUEUE;  // (the tests shrink the queue to exercise the overflow)
	uint32_t *seen = bm_smem, *twice = bm_smem + words, *queue = bm_smem + MAXBITS / 32 * 2;
	for (int k = threadIdx.x; k < 36 * 36; k += blockDim.x)
		S[k] = (int8_t)a.kmer_mx[k];
	for (uint32_t k = threadIdx.x; k < 2 * words; k += blockDim.x)
		bm_smem[k] = 0;
	if (threadIdx.x == 0) { s_n = 0; s_next = 0; }
	__syncthreads();
	const uint8_t *T = a.muT + a.offT[t];
	const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
	uint2 r = probe_row(a, T, LT, warp, S);
	for (uint32_t tpos = warp; tpos + 7 <= LT; tpos += nwarps) {
		const uint2 rn = probe_row(a, T, LT, tpos + nwarps, S);
		for (uint32_t e = r.x + lane; e < r.y; e += 32) {
			const uint32_t v = __ldg(a.ix_val + e);
			const uint32_t q = v >> 16, qpos = v & 0xffffu;
			const uint2 qi = __ldg(a.qinfo + q);  // length, residues of the queries before q
			const uint32_t diag = qi.x + tpos - qpos - 1;  // diag.h:22-25
			if (diag > 0x3fffu)
				continue;  // prefiltermu.cpp:254
			const uint32_t idx = qi.y + q * (LT - 1) + diag;
			const uint32_t w = idx >> 5, m = 1u << (idx & 31u);
			if ((atomicOr(&seen[w], m) & m) && !(atomicOr(&twice[w], m) & m)) {
				const unsigned slot = atomicAdd(&s_n, 1u);
				if (slot < qcap)
					queue[slot] = (q << 14) | diag;
			}
		}
		r = rn;
	}
	__syncthreads();
	const unsigned found = s_n;
	const unsigned n = min(found, qcap);
	for (;;) {
		const unsigned w = atomicAdd(&s_next, 1u);
		if (w >= n)
			break;
		const uint32_t k = queue[w];
		if (found > qcap) {  // leave only the diagonals that missed the queue in `twice`
			const uint32_t q = k >> 14;
			const uint32_t idx = __ldg(a.qinfo + q).y + q * (LT - 1) + (k & 0x3fffu);
			atomicAnd(&twice[idx >> 5], ~(1u << (idx & 31u)));
		}
		walk_diagonal(a, S, T, LT, tl, k);
	}
	if (found <= qcap)
		return;
	// the diagonals that found the queue full: collect them from the bitmap, at most a queue-full per round
	const uint32_t wstep = max(qcap / 32u, 1u);
	for (uint32_t w0 = 0; w0 < words; w0 += wstep) {
		__syncthreads();
		if (threadIdx.x == 0) { s_n = 0; s_next = 0; }
		__syncthreads();
		for (uint32_t w = w0 + threadIdx.x; w < min(words, w0 + wstep); w += blockDim.x) {
			uint32_t bits = twice[w];
			while (bits) {
				const uint32_t idx = w * 32 + (__ffs(bits) - 1);
				bits &= bits - 1;
				// the query whose diagonals contain idx: the last q with qinfo[q].y + q * (LT - 1) <= idx
				uint32_t lo = 0, hi = a.nQ - 1;
				while (lo < hi) {
					const uint32_t mid = (lo + hi + 1) >> 1;
					if (__ldg(a.qinfo + mid).y + mid * (LT - 1) <= idx)
						lo = mid;
					else
						hi = mid - 1;
				}
				const uint32_t diag = idx - (__ldg(a.qinfo + lo).y + lo * (LT - 1));
				if (qcap >= 32)
					queue[atomicAdd(&s_n, 1u)] = (lo << 14) | diag;
				else
					walk_diagonal(a, S, T, LT, tl, (lo << 14) | diag);
			}
		}
		__syncthreads();
		const unsigned n2 = s_n;
		for (;;) {
			const unsigned w = atomicAdd(&s_next, 1u);
			if (w >= n2)
				break;
			walk_diagonal(a, S, T, LT, tl, queue[w]);
		}
	}
}

// The same with the letters in shared memory, for query blocks small enough to stage (<= kStageQBytes residues, <= kStageQMax
// queries): persistent CTAs pull targets from a counter, the query letters and the per-query (length, offset) pairs are staged
// once per CTA, the target's letters once per target.  The diagonal walks then read both chains from shared memory - in the
// unstaged form each lane's letter words were separate uncoalesced L1 requests and kept the L1 data pipe 92 % busy.
constexpr uint32_t kStageQBytes = 40960, kStageQMax = 1024, kStageTBytes = 4096;
__host__ __device__ constexpr size_t bm_stage_bytes(uint32_t nq, uint32_t sumq) { return (size_t)nq * 8 + ((sumq + 3) & ~3u) + 8 + kStageTBytes + 8; }

template <uint32_t MAXBITS, uint32_t QUEUE, int THREADS>
__global__ void __launch_bounds__(THREADS) pf_probe_extend_staged_kernel(const PfArgs a, const uint32_t ntl)
{
	extern __shared__ __align__(16) uint32_t bm_smem[];
	__shared__ int8_t S[36 * 36];
	__shared__ unsigned s_n, s_next, s_target;
	uint32_t *queue = bm_smem + MAXBITS / 32 * 2;
	uint2 *s_qinfo = reinterpret_cast<uint2 *>(queue + QUEUE);
	uint8_t *sQraw = reinterpret_cast<uint8_t *>(s_qinfo + a.nQ);
	uint8_t *sT = sQraw + ((a.sum_lenQ + 3) & ~3u) + 8;
	const unsigned qsh = (unsigned)((uintptr_t)a.muQ & 3u);  // the block's letters are copied from the aligned word at or below them
	const uint8_t *sQ = sQraw + qsh;
	for (int k = threadIdx.x; k < 36 * 36; k += blockDim.x)
		S[k] = (int8_t)a.kmer_mx[k];
	for (uint32_t k = threadIdx.x; k < a.nQ; k += blockDim.x)
		s_qinfo[k] = a.qinfo[k];
	{
		// chain sets are packed (offsets = prefix sums of the lengths), so the block's letters are one contiguous range
		const uint32_t *src = reinterpret_cast<const uint32_t *>(a.muQ - qsh);
		uint32_t *dst = reinterpret_cast<uint32_t *>(sQraw);
		for (uint32_t k = threadIdx.x; k < (a.sum_lenQ + qsh + 3) / 4; k += blockDim.x)
			dst[k] = __ldg(src + k);
	}
	const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
	const uint32_t qcap = a.queue_cap ? min(a.queue_cap, QUEUE) : QUEUE;
	for (;;) {
		__syncthreads();  // the previous target is done with the bitmaps, the queue and sT
		if (threadIdx.x == 0)
			s_target = atomicAdd(a.fuse_counter, 1u);
		__syncthreads();
		const uint32_t tl = s_target;
		if (tl >= ntl)
			break;
		const uint32_t t = a.t_begin + tl;
		const uint32_t LT = a.lenT[t];
		if (LT < 7)
			continue;
		const unsigned long long nbits = (unsigned long long)a.sum_lenQ + (unsigned long long)a.nQ * (LT - 1);
		if (nbits > MAXBITS || LT > kStageTBytes)
			continue;  // the other size class, or the global-memory path
		const uint32_t words = ((uint32_t)nbits + 31) / 32;
		uint32_t *seen = bm_smem, *twice = bm_smem + words;
		for (uint32_t k = threadIdx.x; k < 2 * words; k += blockDim.x)
			bm_smem[k] = 0;
		{
			// the target's letters, from the aligned word at or below its first byte
			const uint8_t *Tg = a.muT + a.offT[t];
			const unsigned sh = (unsigned)((uintptr_t)Tg & 3u);
			const uint32_t *src = reinterpret_cast<const uint32_t *>(Tg - sh);
			uint32_t *dst = reinterpret_cast<uint32_t *>(sT);
			for (uint32_t k = threadIdx.x; k < (LT + sh + 3) / 4; k += blockDim.x)
				dst[k] = __ldg(src + k);
			if (threadIdx.x == 0) { s_n = 0; s_next = 0; }
			__syncthreads();
			const uint8_t *T = sT + sh;
			uint2 r = probe_row(a, T, LT, warp, S);
			for (uint32_t tpos = warp; tpos + 7 <= LT; tpos += nwarps) {
				const uint2 rn = probe_row(a, T, LT, tpos + nwarps, S);
				for (uint32_t e = r.x + lane; e < r.y; e += 32) {
					const uint32_t v = __ldg(a.ix_val + e);
					const uint32_t q = v >> 16, qpos = v & 0xffffu;
					const uint2 qi = s_qinfo[q];  // length, residues of the queries before q
					const uint32_t diag = qi.x + tpos - qpos - 1;  // diag.h:22-25
					if (diag > 0x3fffu)
						continue;  // prefiltermu.cpp:254
					const uint32_t idx = qi.y + q * (LT - 1) + diag;
					const uint32_t w = idx >> 5, m = 1u << (idx & 31u);
					if ((atomicOr(&seen[w], m) & m) && !(atomicOr(&twice[w], m) & m)) {
						const unsigned slot = atomicAdd(&s_n, 1u);
						if (slot < qcap)
							queue[slot] = (q << 14) | diag;
					}
				}
				r = rn;
			}
			__syncthreads();
			auto walk = [&](const uint32_t k) {
				const uint32_t q = k >> 14;
				const uint2 qi = s_qinfo[q];
				int B = walk_core<true>(S, sQ + qi.y, qi.x, T, LT, (int)(k & 0x3fffu));
				if (B > 0) {
					if (B >= 65535) B = 65534;  // prefiltermu.cpp:294-295
					atomicMax(&a.best[(size_t)tl * a.nQ + q], (unsigned)B);
				}
			};
			const unsigned found = s_n;
			const unsigned n = min(found, qcap);
			for (;;) {
				const unsigned w = atomicAdd(&s_next, 1u);
				if (w >= n)
					break;
				const uint32_t k = queue[w];
				if (found > qcap) {  // leave only the diagonals that missed the queue in `twice`
					const uint32_t q = k >> 14;
					const uint32_t idx = s_qinfo[q].y + q * (LT - 1) + (k & 0x3fffu);
					atomicAnd(&twice[idx >> 5], ~(1u << (idx & 31u)));
				}
				walk(k);
			}
			if (found <= qcap)
				continue;
			// the diagonals that found the queue full: collect them from the bitmap, at most a queue-full per round
			const uint32_t wstep = max(qcap / 32u, 1u);
			for (uint32_t w0 = 0; w0 < words; w0 += wstep) {
				__syncthreads();
				if (threadIdx.x == 0) { s_n = 0; s_next = 0; }
				__syncthreads();
				for (uint32_t w = w0 + threadIdx.x; w < min(words, w0 + wstep); w += blockDim.x) {
					uint32_t bits = twice[w];
					while (bits) {
						const uint32_t idx = w * 32 + (__ffs(bits) - 1);
						bits &= bits - 1;
						uint32_t lo = 0, hi = a.nQ - 1;  // the last q with s_qinfo[q].y + q * (LT - 1) <= idx
						while (lo < hi) {
							const uint32_t mid = (lo + hi + 1) >> 1;
							if (s_qinfo[mid].y + mid * (LT - 1) <= idx)
								lo = mid;
							else
								hi = mid - 1;
						}
						const uint32_t diag = idx - (s_qinfo[lo].y + lo * (LT - 1));
						if (qcap >= 32)
							queue[atomicAdd(&s_n, 1u)] = (lo << 14) | diag;
						else
							walk((lo << 14) | diag);
					}
				}
				__syncthreads();
				const unsigned n2 = s_n;
				for (;;) {
					const unsigned w = atomicAdd(&s_next, 1u);
					if (w >= n2)
						break;
					walk(queue[w]);
				}
			}
		}
	}
}

// ---- K8: two-hit diagonals -> FindHSP -> best score per (target, query) ----
// Two phases per CTA (= per target): the sorted hit keys are scanned for the first element of every run of length >= 2 (a
// two-hit diagonal) and those are queued in shared memory; then every thread of the CTA pulls diagonals from the queue and
// runs the Kadane scan.  Scanning and walking in one loop left most threads idle (only run starts walk).
constexpr int kPfQueue = 2048;
__global__ void __launch_bounds__(128) pf_extend_kernel(const PfArgs a)
{
	__shared__ int8_t S[36 * 36];
	__shared__ uint32_t s_queue[kPfQueue];
	__shared__ unsigned s_n, s_next;
	for (int k = threadIdx.x; k < 36 * 36; k += blockDim.x)
		S[k] = (int8_t)a.kmer_mx[k];
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
			walk_diagonal(a, S, T, LT, tl, s_queue[w]);
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

// triples leave as (query, target<<16 | score): the form the bag kernels sort and replay (target index in the numbering of
// the whole DB: t_base = first target of this rank's block)
__global__ void pf_write_cands_kernel(const PfArgs a, uint32_t ntl)
{
	const uint32_t tl = blockIdx.x * blockDim.x + threadIdx.x;
	if (tl >= ntl)
		return;
	unsigned long long o = a.raw_base + a.cand_off[tl];
	for (uint32_t q = 0; q < a.nQ; ++q) {
		const unsigned b = a.best[(size_t)tl * a.nQ + q];
		if (b) {
			a.raw_q[o] = q;
			a.raw_v[o] = ((unsigned long long)(a.t_base + a.t_begin + tl) << 16) | (unsigned long long)b;
			++o;
		}
	}
}

// ---- RankedScoresBag on the device (rankedscoresbag.cpp:5-51, :185-232) ----
// The triples, stably sorted by query, are replayed one warp per query in stream order: AddScore admits `score >= lo` until
// the vectors hold 2B entries, TruncateVecs then keeps the first B of QuickSortOrderDesc (sort.h:71-108).  That quicksort is
// unstable, and which of the tied entries survive at the cut-off - and how the survivors are arranged for the NEXT
// truncation - is decided by its exact swap sequence, so it is restated (Hoare partition around the middle element) and run
// on shared memory: its partition steps by the whole warp with the serial loop's exact permutation (bag_partition_warp), the
// small ranges serially, one lane each; the admission scan, the gather of the survivors and the output are warp-parallel too.
__device__ void bag_quicksort_desc(uint16_t *vs, uint32_t *order, int left, int right)
{
	// vs[i] mirrors Values[Order[i]] and is swapped together with order[i]; sub-ranges are disjoint, so the order in which the
	// recursion visits them is irrelevant: an explicit stack, larger half pushed, bounds the depth by log2(n)
	int stack_l[32], stack_r[32];
	int sp = 0;
	for (;;) {
		while (left < right) {
			int i = left, j = right;
			const uint16_t pivot = vs[(left + right) / 2];
			while (i <= j) {
				while (vs[i] > pivot)
					i++;
				while (vs[j] < pivot)
					j--;
				if (i <= j) {
					const uint16_t tv = vs[i]; vs[i] = vs[j]; vs[j] = tv;
					const uint32_t to = order[i]; order[i] = order[j]; order[j] = to;
					i++;
					j--;
				}
			}
			// ranges (left, j) and (i, right)
			const bool hasL = left < j, hasR = i < right;
			if (hasL && hasR) {
				if (j - left > right - i) {
					stack_l[sp] = left; stack_r[sp] = j; ++sp;
					left = i;
				} else {
					stack_l[sp] = i; stack_r[sp] = right; ++sp;
					right = j;
				}
			} else if (hasL) {
				right = j;
			} else if (hasR) {
				left = i;
			} else {
				break;
			}
		}
		if (sp == 0)
			break;
		--sp;
		left = stack_l[sp]; right = stack_r[sp];
	}
}

// One partition step of that quicksort on [left, right], by the whole warp, with the permutation the serial loop produces.
// The serial loop swaps the k-th position from the left whose value is <= pivot with the k-th position from the right whose
// value is >= pivot, k = 0, 1, ... while the left one is not beyond the right one; both sequences can be read off the
// UNCHANGED array (the scanning indices never re-enter swapped territory before they cross), the swapped positions are
// pairwise distinct, and where the two indices end follows from the first unswapped stop on either side.
__device__ void bag_partition_warp(uint16_t *vs, uint32_t *order, const int left, const int right, uint16_t *lpos, uint16_t *rpos,
		const int lane, int &iout, int &jout)
{
	const unsigned lt = (1u << lane) - 1u;
	const uint16_t pivot = vs[(left + right) / 2];
	int nL = 0, nR = 0;
	for (int b = left; b <= right; b += 32) {
		const int p = b + lane;
		const bool ok = p <= right && vs[p] <= pivot;
		const unsigned m = __ballot_sync(kFull, ok);
		if (ok)
			lpos[nL + __popc(m & lt)] = (uint16_t)p;
		nL += __popc(m);
	}
	for (int b = right; b >= left; b -= 32) {
		const int p = b - lane;
		const bool ok = p >= left && vs[p] >= pivot;
		const unsigned m = __ballot_sync(kFull, ok);
		if (ok)
			rpos[nR + __popc(m & lt)] = (uint16_t)p;
		nR += __popc(m);
	}
	__syncwarp();
	const int nmin = min(nL, nR);
	int K = 0;
	for (int k = lane; k < nmin; k += 32)
		K += lpos[k] <= rpos[k];
#pragma unroll
	for (int o = 16; o >= 1; o >>= 1)
		K += __shfl_xor_sync(kFull, K, o);
	// K >= 1: the pivot's own position is in both sequences
	const int a_last = lpos[K - 1], b_last = rpos[K - 1];
	int i = a_last + 1, j = b_last - 1;
	if (i <= j) {
		i = (K < nL && lpos[K] < b_last) ? lpos[K] : b_last;
		j = (K < nR && rpos[K] > a_last) ? rpos[K] : a_last;
	}
	for (int k = lane; k < K; k += 32) {
		const int pa = lpos[k], pb = rpos[k];
		if (pa != pb) {
			const uint16_t tv = vs[pa]; vs[pa] = vs[pb]; vs[pb] = tv;
			const uint32_t to = order[pa]; order[pa] = order[pb]; order[pb] = to;
		}
	}
	__syncwarp();
	iout = i;
	jout = j;
}

// QuickSortOrderDesc over vs/order[0..n): large ranges are partitioned by the warp, ranges below kBagSmall are collected and
// then sorted serially, one lane per range (the ranges are disjoint).
constexpr int kBagSmall = 96, kBagRanges = 512;
__device__ void bag_sort_warp(uint16_t *vs, uint32_t *order, const int n, uint16_t *lpos, uint16_t *rpos, int *ranges, const int lane)
{
	// ranges[]: a stack of large ranges growing up from 0 (2 ints each), the list of small ranges growing down from the end
	int nbig = 0, nsmall = 0;
	int left = 0, right = n - 1;
	bool have = n > 1;
	while (have || nbig > 0) {
		if (!have) {
			--nbig;
			left = ranges[2 * nbig];
			right = ranges[2 * nbig + 1];
		}
		have = false;
		if (right - left + 1 < kBagSmall || nsmall + nbig + 2 >= kBagRanges) {
			if (nsmall + nbig + 1 < kBagRanges) {
				if (lane == 0) {
					ranges[2 * (kBagRanges - 1 - nsmall)] = left;
					ranges[2 * (kBagRanges - 1 - nsmall) + 1] = right;
				}
				++nsmall;
			} else {  // no room to remember it: sort it now
				if (lane == 0)
					bag_quicksort_desc(vs, order, left, right);
				__syncwarp();
			}
			continue;
		}
		int i, j;
		bag_partition_warp(vs, order, left, right, lpos, rpos, lane, i, j);
		const bool hasL = left < j, hasR = i < right;
		if (hasL && hasR) {
			if (lane == 0) {
				ranges[2 * nbig] = i;
				ranges[2 * nbig + 1] = right;
			}
			++nbig;
			right = j;
			have = true;
		} else if (hasL) {
			right = j;
			have = true;
		} else if (hasR) {
			left = i;
			have = true;
		}
		__syncwarp();
	}
	__syncwarp();
	for (int k = lane; k < nsmall; k += 32)
		bag_quicksort_desc(vs, order, ranges[2 * (kBagRanges - 1 - k)], ranges[2 * (kBagRanges - 1 - k) + 1]);
	__syncwarp();
}

__global__ void __launch_bounds__(32) pf_bag_kernel(const unsigned long long *__restrict__ val, const unsigned long long *__restrict__ seg_begin,
		const unsigned long long *__restrict__ seg_end, uint32_t B, unsigned long long *out_key, uint32_t *out_n)
{
	extern __shared__ unsigned char bag_smem[];
	uint32_t *t = (uint32_t *)bag_smem;         // [2B] target of entry i
	uint32_t *order = t + 2 * B;                // [2B]
	uint32_t *t2 = order + 2 * B;               // [B]
	uint16_t *s = (uint16_t *)(t2 + B);         // [2B] score of entry i
	uint16_t *vs = s + 2 * B;                   // [2B] s[order[i]]
	uint16_t *lpos = vs + 2 * B;                // [2B] scratch of the warp partition
	uint16_t *rpos = lpos + 2 * B;              // [2B]
	int *ranges = (int *)(((uintptr_t)(rpos + 2 * B) + 3) & ~(uintptr_t)3);  // [2 * kBagRanges]
	const uint32_t q = blockIdx.x, lane = threadIdx.x;
	unsigned long long pos = seg_begin[q];
	const unsigned long long end = seg_end[q];
	uint32_t n = 0, lo = 0;
	auto truncate = [&]() {  // TruncateVecs
		for (uint32_t k = lane; k < n; k += 32) {
			order[k] = k;
			vs[k] = s[k];
		}
		__syncwarp();
		bag_sort_warp(vs, order, (int)n, lpos, rpos, ranges, (int)lane);
		for (uint32_t k = lane; k < B; k += 32)
			t2[k] = t[order[k]];
		__syncwarp();
		for (uint32_t k = lane; k < B; k += 32) {
			t[k] = t2[k];
			s[k] = vs[k];
		}
		__syncwarp();
		lo = s[B - 1];
		n = B;
	};
	while (pos < end) {
		const unsigned long long i = pos + lane;
		uint32_t sc = 0, tt = 0;
		bool ok = false;
		if (i < end) {
			const unsigned long long v = val[i];
			sc = (uint32_t)(v & 0xffffu);
			tt = (uint32_t)(v >> 16);
			ok = sc >= lo;  // AddScore: Score >= LoScore
		}
		unsigned m = __ballot_sync(kFull, ok);
		const uint32_t room = 2 * B - n;
		uint32_t consumed = 32;
		if ((uint32_t)__popc(m) >= room) {
			unsigned mm = m;  // lane of the room-th admitted entry: the one that fills the vectors
			for (uint32_t r = 1; r < room; ++r)
				mm &= mm - 1;
			const uint32_t last = (uint32_t)__ffs((int)mm) - 1u;
			m &= (last == 31) ? kFull : ((2u << last) - 1u);
			consumed = last + 1;
		}
		if ((m >> lane) & 1u) {
			const uint32_t slot = n + __popc(m & ((1u << lane) - 1u));
			s[slot] = (uint16_t)sc;
			t[slot] = tt;
		}
		n += __popc(m);
		pos += consumed;
		__syncwarp();
		if (n >= 2 * B)
			truncate();
	}
	if (n >= B)  // ToTsv's final TruncateVecs (a vector below B entries is left as it is)
		truncate();
	for (uint32_t k = lane; k < n; k += 32)
		out_key[(size_t)q * B + k] = ((unsigned long long)t[k] << 32) | ((unsigned long long)q << 16) | s[k];
	if (lane == 0)
		out_n[q] = n;
}

__global__ void pf_mark_segments_kernel(const uint32_t *__restrict__ key, unsigned long long n, unsigned long long *seg_begin, unsigned long long *seg_end)
{
	const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	const uint32_t k = key[i];
	if (i == 0 || key[i - 1] != k)
		seg_begin[k] = i;
	if (i + 1 == n || key[i + 1] != k)
		seg_end[k] = i + 1;
}

// gather the bags' entries (at most B per query, out_n[q] valid) into one dense key array
__global__ void pf_bag_compact_kernel(const unsigned long long *__restrict__ key, const uint32_t *__restrict__ out_n,
		const unsigned long long *__restrict__ out_off, uint32_t B, unsigned long long *dense)
{
	const uint32_t q = blockIdx.x;
	const uint32_t n = out_n[q];
	const unsigned long long o = out_off[q];
	for (uint32_t k = threadIdx.x; k < n; k += blockDim.x)
		dense[o + k] = key[(size_t)q * B + k];
}

__global__ void pf_unpack_triples_kernel(const uint32_t *__restrict__ q, const unsigned long long *__restrict__ v, unsigned long long n,
		uint32_t *t_out, uint32_t *q_out, uint16_t *s_out)
{
	const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	const unsigned long long x = v[i];
	t_out[i] = (uint32_t)(x >> 16);
	q_out[i] = q[i];
	s_out[i] = (uint16_t)(x & 0xffffu);
}

__global__ void pf_unpack_keys_kernel(const unsigned long long *__restrict__ key, unsigned long long n, uint32_t *t_out, uint32_t *q_out, uint16_t *s_out)
{
	const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	const unsigned long long x = key[i];
	t_out[i] = (uint32_t)(x >> 32);
	q_out[i] = (uint32_t)(x >> 16) & 0xffffu;
	s_out[i] = (uint16_t)(x & 0xffffu);
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
		pf_neighborhood_kernel<true><<<(a.nqk + kNbWarps - 1) / kNbWarps, kNbWarps * 32, 0, st>>>(a);
	else
		pf_neighborhood_kernel<false><<<(a.nqk + kNbWarps - 1) / kNbWarps, kNbWarps * 32, 0, st>>>(a);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int pf_launch_mark_rows(const uint32_t *key, unsigned long long n, uint2 *row, cudaStream_t st)
{
	if (n == 0)
		return 0;
	pf_mark_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(key, n, row);
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

int pf_launch_probe_extend(const PfArgs &a, uint32_t ntl, int which, cudaStream_t st)
{
	if (ntl == 0)
		return 0;
	auto small = pf_probe_extend_kernel<0, kBmSmallBits, kBmSmallQueue, 256>;
	auto large = pf_probe_extend_kernel<kBmSmallBits, kBmLargeBits, kBmLargeQueue, 1024>;
	auto staged = pf_probe_extend_staged_kernel<kBmSmallBits, kBmSmallQueue, 256>;
	const size_t smem_s = bm_smem_bytes(kBmSmallBits, kBmSmallQueue), smem_l = bm_smem_bytes(kBmLargeBits, kBmLargeQueue);
	const size_t smem_st = smem_s + bm_stage_bytes(a.nQ, a.sum_lenQ);
	// per launch, not once per process: the attribute belongs to the current device, and `-gpus N` drives several devices from
	// the threads of one process
	if (cudaFuncSetAttribute(small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s) != cudaSuccess ||
		cudaFuncSetAttribute(large, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l) != cudaSuccess ||
		cudaFuncSetAttribute(staged, cudaFuncAttributeMaxDynamicSharedMemorySize,
			(int)(smem_s + bm_stage_bytes(kStageQMax, kStageQBytes))) != cudaSuccess)
		return -1;
	int n = 0;
	if (which & 1) {
		// small class: letters staged in shared memory when the query block allows it (and the caller gave a target counter);
		// at least 17 queries bound a small-class target to kStageTBytes residues
		const bool stage = a.fuse_counter && a.nQ >= 17 && a.nQ <= kStageQMax && a.sum_lenQ <= kStageQBytes && !a.no_stage;
		if (stage) {
			int dev = 0, sms = 148;
			cudaGetDevice(&dev);
			cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
			const unsigned per_sm = (unsigned)std::max<size_t>(1, std::min<size_t>(8, (size_t)220 * 1024 / (smem_st + 2048)));
			const unsigned grid = (unsigned)std::min<uint64_t>(ntl, (uint64_t)sms * per_sm);
			staged<<<grid, 256, smem_st, st>>>(a, ntl);
		} else {
			small<<<ntl, 256, smem_s, st>>>(a);
		}
		++n;
	}
	if (which & 2) {
		large<<<ntl, 1024, smem_l, st>>>(a);
		++n;
	}
	return cudaGetLastError() == cudaSuccess ? n : -1;
}

// (query, diagonal) pairs a target may have for the fused kernels: size class 0, size class 1
unsigned long long pf_fuse_max_bits(int size) { return size == 0 ? kBmSmallBits : kBmLargeBits; }

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

size_t pf_bag_smem_bytes(uint32_t B) { return (size_t)B * (4 * 2 + 4 * 2 + 4 + 2 * 2 + 2 * 2 + 2 * 2 + 2 * 2) + 8 * kBagRanges + 32; }

// stable sort of the triples by query (16 key bits), value = target<<16 | score
int pf_sort_by_query(const uint32_t *qin, uint32_t *qout, const unsigned long long *vin, unsigned long long *vout, unsigned long long n,
		void *tmp, size_t &tmp_bytes, cudaStream_t st)
{
	return cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, qin, qout, vin, vout, (long long)n, 0, 16, st) == cudaSuccess ? 0 : -1;
}

int pf_sort_keys64(const unsigned long long *kin, unsigned long long *kout, unsigned long long n, void *tmp, size_t &tmp_bytes, cudaStream_t st)
{
	return cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, kin, kout, (long long)n, 0, 64, st) == cudaSuccess ? 0 : -1;
}

int pf_launch_mark_segments(const uint32_t *key, unsigned long long n, unsigned long long *seg_begin, unsigned long long *seg_end, cudaStream_t st)
{
	if (n == 0)
		return 0;
	pf_mark_segments_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(key, n, seg_begin, seg_end);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int pf_launch_bag(const unsigned long long *val, const unsigned long long *seg_begin, const unsigned long long *seg_end, uint32_t nQ, uint32_t B,
		unsigned long long *out_key, uint32_t *out_n, cudaStream_t st)
{
	if (nQ == 0)
		return 0;
	const size_t smem = pf_bag_smem_bytes(B);
	// per launch: the attribute belongs to the current device (several devices per process under `-gpus N`)
	if (smem > 48 * 1024 &&
		cudaFuncSetAttribute(pf_bag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
		return -1;
	pf_bag_kernel<<<nQ, 32, smem, st>>>(val, seg_begin, seg_end, B, out_key, out_n);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int pf_launch_bag_compact(const unsigned long long *key, const uint32_t *out_n, const unsigned long long *out_off, uint32_t nQ, uint32_t B,
		unsigned long long *dense, cudaStream_t st)
{
	if (nQ == 0)
		return 0;
	pf_bag_compact_kernel<<<nQ, 128, 0, st>>>(key, out_n, out_off, B, dense);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int pf_launch_unpack_triples(const uint32_t *q, const unsigned long long *v, unsigned long long n, uint32_t *t_out, uint32_t *q_out,
		uint16_t *s_out, cudaStream_t st)
{
	if (n == 0)
		return 0;
	pf_unpack_triples_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(q, v, n, t_out, q_out, s_out);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int pf_launch_unpack_keys(const unsigned long long *key, unsigned long long n, uint32_t *t_out, uint32_t *q_out, uint16_t *s_out, cudaStream_t st)
{
	if (n == 0)
		return 0;
	pf_unpack_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(key, n, t_out, q_out, s_out);
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
