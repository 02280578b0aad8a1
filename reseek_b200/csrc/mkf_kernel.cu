// mkf_kernel.cu - K4: the long-chain path (a chain >= MKFL): Mu 3-mer seeds -> ungapped x-drop HSPs with the
// reference's order-dependent gating -> 1-D chaining -> mega-HSP re-scoring -> best 8-mer seed -> banded affine
// x-drop forward and backward with traceback -> merged path.
//
// Replaces DSSAligner::AlignMKF / PostAlignMKF (dssaligner.cpp:1387-1437), MuKmerFilter::SetHashTable / Align /
// MuXDrop / ChainHSPs (mukmerfilter.cpp:208, 316, 105, 391), Chainer::Chain (chainer.cpp:31-194),
// GetMegaHSPScore (dssaligner.cpp:488-527), XDropHSP (xdrophsp.cpp:42-150), XDropFwd (xdropfwd.cpp:71-386),
// XDropBwd (xdropbwd.cpp:28-50), MergeFwdBwd (mergefwdback.cpp:6-50).
//
// These pairs are rare (3 % of SCOP40 pairs) and their DP is inherently sequential: the band of row i+1 and every
// x-drop test depend on the running best score in row-major order (xdropfwd.cpp:226-262).  They are therefore
// mapped as: one CTA per query chain for the hash table, one warp per pair for seeding/chaining/re-scoring, and
// ONE THREAD per (pair, direction) for the banded DP, thousands of pairs in flight.  Every float operation is
// performed in the reference's order, so scores and paths are bit-identical.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "rsk_internal.cuh"

namespace rsk {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kDict = 36 * 36 * 36;       // 3-mer dictionary (pattern "111", dssparams.cpp:88)
constexpr int kHashW = 4;                 // first 4 positions of every 3-mer (mukmerfilter.h:7)
constexpr int kMaxHsp = 80;               // kept HSPs per pair (each must beat the previous best score: never many)
constexpr int kSeedWarps = 4;
constexpr int kCandCap = 256;             // listed hash hits per flush of the seed kernel (a chunk of 32 positions adds at most 128)

// ---- query hash tables: uint16 ht[kDict][4], 0xffff = empty ----
__global__ void __launch_bounds__(256) mkf_hash_kernel(const MkfArgs a)
{
	const uint32_t q = a.hash_chain[blockIdx.x];
	uint16_t *ht = a.hash + (size_t)blockIdx.x * kDict * kHashW;
	uint4 *ht4 = reinterpret_cast<uint4 *>(ht);
	const uint4 ff = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
	for (int k = threadIdx.x; k < kDict * kHashW * 2 / 16; k += 256)
		ht4[k] = ff;
	__syncthreads();
	if (threadIdx.x == 0) {
		const uint8_t *mu = a.muA + a.offA[q];
		const int L = (int)a.lenA[q];
		for (int pos = 0; pos + 3 <= L; ++pos) {  // insertion order = position order (mukmerfilter.cpp:212-227)
			const int k = (mu[pos] * 36 + mu[pos + 1]) * 36 + mu[pos + 2];
			for (int w = 0; w < kHashW; ++w)
				if (ht[kHashW * k + w] == 0xffff) {
					ht[kHashW * k + w] = (uint16_t)pos;
					break;
				}
		}
	}
}

// mukmerfilter.cpp:105-175
__device__ __forceinline__ int mu_xdrop(const int *mx, const uint8_t *__restrict__ Q, int LQ, const uint8_t *__restrict__ T, int LT,
		int PosQ, int PosT, int X, int &Loi, int &Loj, int &Len)
{
	Loi = PosQ;
	Loj = PosT;
	int fwd = 0, bestfwd = 0, fwdlen = 0;
	for (int i = PosQ, j = PosT; i < LQ && j < LT;) {
		fwd += mx[36 * Q[i] + T[j]];
		++i; ++j;
		if (fwd > bestfwd) { bestfwd = fwd; fwdlen = i - PosQ; }
		else if (fwd + X < bestfwd) break;
	}
	int rev = 0, bestrev = 0, revlen = 0;
	for (int i = PosQ - 1, j = PosT - 1; i >= 0 && j >= 0; --i, --j) {
		rev += mx[36 * Q[i] + T[j]];
		if (rev > bestrev) { bestrev = rev; Loi = i; Loj = j; revlen = PosQ - i; }
		else if (rev + X < bestrev) break;
	}
	Len = fwdlen + revlen;
	return bestfwd + bestrev;
}

struct Hsp { int loi, loj, len, score; };

// cell score in the order of xdrophsp.cpp:8-33 (starts from 0, features 0..7)
__device__ __forceinline__ float subst(const float *tab, const uint64_t ea, const uint64_t eb)
{
	float t = 0.0f;
#pragma unroll
	for (int f = 0; f < RSK_NFEAT; ++f) {
		const int a = (int)((ea >> (8 * f)) & 0xff) - feat_base(f);
		const int b = (int)((eb >> (8 * f)) & 0xff) - feat_base(f);
		t += tab[feat_table_off(f) + a * feat_alpha(f) + b];
	}
	return t;
}

// One warp per pair: seeds, gating, chaining, mega-HSP scores, 8-mer seed.
__global__ void __launch_bounds__(kSeedWarps * 32) mkf_seed_kernel(const MkfArgs a)
{
	__shared__ int s_mx[36 * 36];
	__shared__ float s_tab[RSK_TABLE_FLOATS];
	__shared__ Hsp s_cand[kSeedWarps][kCandCap];       // extension results of the listed hits; later the chainer's breakpoints
	__shared__ uint32_t s_cpt[kSeedWarps][kCandCap];   // listed hits: target position ...
	__shared__ uint16_t s_cpq[kSeedWarps][kCandCap];   // ... and query position
	__shared__ Hsp s_hsp[kSeedWarps][kMaxHsp];
	__shared__ int s_chain[kSeedWarps][kMaxHsp];
	__shared__ float s_mega[kSeedWarps][kMaxHsp];
	for (int k = threadIdx.x; k < 36 * 36; k += blockDim.x)
		s_mx[k] = a.mu_mx[k];
	for (int k = threadIdx.x; k < RSK_TABLE_FLOATS; k += blockDim.x)
		s_tab[k] = a.tables[k];
	__syncthreads();
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t pair = blockIdx.x * kSeedWarps + warp;
	if (pair >= a.npairs)
		return;
	const uint32_t qa = a.pair_a[pair], tb = a.pair_b[pair];
	const uint8_t *Q = a.muA + a.offA[qa];
	const uint8_t *T = a.muB + a.offB[tb];
	const int LQ = (int)a.lenA[qa], LT = (int)a.lenB[tb];
	const uint16_t *ht = a.hash + (size_t)a.pair_hash[pair] * kDict * kHashW;
	Hsp *cand = s_cand[warp];
	Hsp *hsp = s_hsp[warp];
	int *chain = s_chain[warp];
	MkfSeed out;
	out.valid = 0; out.lo_a = 0; out.lo_b = 0; out.best_hsp = 0; out.best_chain = 0;

	// ---- seeds + gating (mukmerfilter.cpp:316-389): the order is PosT ascending, slot ascending ----
	// The hash hits of a target are sparse (about one per 32 positions), and an ungapped x-drop extension is a loop of its own
	// length: extending each position's hits where they are found kept 1-2 lanes of the warp busy.  So the hits are first
	// collected, in order, into a list (position, query position); a full list is flushed: every lane extends one hit at a time
	// (all lanes busy), then the gate - which is order dependent, but only for the rare extensions that reach MinHSPScore -
	// walks the flagged results in list order with nh / best / found uniform over the warp.
	int nh = 0, best = 0;
	bool found = false;
	const int nkT = LT - 2;
	uint32_t *cpt = s_cpt[warp];
	uint16_t *cpq = s_cpq[warp];
	int ncand = 0;
	auto flush = [&]() {
		__syncwarp();
		for (int i = lane; i < ncand; i += 32) {
			Hsp h;
			h.score = mu_xdrop(s_mx, Q, LQ, T, LT, (int)cpq[i], (int)cpt[i], a.x1, h.loi, h.loj, h.len);
			cand[i] = h;
		}
		__syncwarp();
		for (int c0 = 0; c0 < ncand; c0 += 32) {
			const int i = c0 + lane;
			const int sc = i < ncand ? cand[i].score : -1;
			unsigned m = __ballot_sync(kFull, sc >= 0 && sc >= a.min_hsp_score);
			if (m)
				found = true;
			while (m) {
				const int src = __ffs(m) - 1;
				m &= m - 1;
				const Hsp h = cand[c0 + src];
				if (h.score > best) {
					best = h.score;
					bool old = false;
					for (int k = lane; k < nh; k += 32)
						old = old || hsp[k].loi == h.loi;
					old = __any_sync(kFull, old);
					if (!old && nh < kMaxHsp) {
						if (lane == 0)
							hsp[nh] = h;
						++nh;
						__syncwarp();
					}
				}
			}
		}
		__syncwarp();
		ncand = 0;
	};
	for (int base = 0; base < nkT; base += 32) {
		const int pt = base + lane;
		uint16_t pq[kHashW];
		unsigned cnt = 0;
#pragma unroll
		for (int w = 0; w < kHashW; ++w)
			pq[w] = 0xffff;
		if (pt < nkT) {
			const int k = (T[pt] * 36 + T[pt + 1]) * 36 + T[pt + 2];
			const ushort4 slots = *reinterpret_cast<const ushort4 *>(ht + kHashW * k);
			pq[0] = slots.x; pq[1] = slots.y; pq[2] = slots.z; pq[3] = slots.w;
#pragma unroll
			for (int w = 0; w < kHashW; ++w)
				cnt += pq[w] != 0xffff;
		}
		unsigned incl = cnt;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const unsigned t = __shfl_up_sync(kFull, incl, o);
			if (lane >= o)
				incl += t;
		}
		const int total = (int)__shfl_sync(kFull, incl, 31);
		if (total == 0)
			continue;
		if (ncand + total > kCandCap)
			flush();
		int pos = ncand + (int)(incl - cnt);
#pragma unroll
		for (int w = 0; w < kHashW; ++w)
			if (pq[w] != 0xffff) {
				cpt[pos] = (uint32_t)pt;
				cpq[pos] = pq[w];
				++pos;
			}
		ncand += total;
	}
	flush();
	out.best_hsp = best;

	// ---- chaining on the query axis (chainer.cpp:31-194); lane 0, N is tiny ----
	int nchain = 0, chain_score = 0;
	if (found && lane == 0 && nh > 0) {
		// breakpoints: (pos, is_lo, index), sorted by pos, starts before ends, stable
		int *bp = reinterpret_cast<int *>(cand);  // reuse the candidate buffer: 2*nh*3 <= 480 ints of its 512
		const int nb = 2 * nh;
		for (int i = 0; i < nh; ++i) {
			bp[3 * (2 * i)] = hsp[i].loi; bp[3 * (2 * i) + 1] = 1; bp[3 * (2 * i) + 2] = i;
			bp[3 * (2 * i + 1)] = hsp[i].loi + hsp[i].len - 1; bp[3 * (2 * i + 1) + 1] = 0; bp[3 * (2 * i + 1) + 2] = i;
		}
		for (int i = 1; i < nb; ++i) {
			const int tp = bp[3 * i], tl = bp[3 * i + 1], ti = bp[3 * i + 2];
			int k = i - 1;
			while (k >= 0 && (bp[3 * k] > tp || (bp[3 * k] == tp && !bp[3 * k + 1] && tl))) {
				bp[3 * (k + 1)] = bp[3 * k]; bp[3 * (k + 1) + 1] = bp[3 * k + 1]; bp[3 * (k + 1) + 2] = bp[3 * k + 2];
				--k;
			}
			bp[3 * (k + 1)] = tp; bp[3 * (k + 1) + 1] = tl; bp[3 * (k + 1) + 2] = ti;
		}
		float *cs = s_mega[warp];      // chain scores
		int *tbk = chain;              // traceback links (overwritten by the chain itself afterwards)
		int best_end = -1;
		for (int i = 0; i < nb; ++i) {
			const int ix = bp[3 * i + 2];
			const float sc = (float)hsp[ix].score;
			if (bp[3 * i + 1]) {
				tbk[ix] = best_end;
				cs[ix] = best_end < 0 ? sc : cs[best_end] + sc;
			} else if (best_end < 0 || cs[ix] > cs[best_end]) {
				best_end = ix;
			}
		}
		float total = 0.0f;
		int tmp[kMaxHsp];
		for (int ix = best_end; ix >= 0; ix = tbk[ix]) {
			total += (float)hsp[ix].score;
			tmp[nchain++] = ix;
		}
		for (int k = 0; k < nchain; ++k)
			chain[k] = tmp[k];
		chain_score = (int)total;
	}
	nchain = __shfl_sync(kFull, nchain, 0);
	chain_score = __shfl_sync(kFull, chain_score, 0);
	out.best_chain = chain_score;
	__syncwarp();

	if (chain_score > 0) {  // PostAlignMKF: dssaligner.cpp:1395-1437
		const uint64_t *PA = a.profA + a.offA[qa];
		const uint64_t *PB = a.profB + a.offB[tb];
		// mega-HSP scores: feature-major running sum (dssaligner.cpp:504-524); one lane per chained HSP
		for (int k = lane; k < nchain; k += 32) {
			const Hsp h = hsp[chain[k]];
			float total = 0.0f;
			for (int f = 0; f < RSK_NFEAT; ++f) {
				const float *tf = s_tab + feat_table_off(f);
				const int al = feat_alpha(f), fb = feat_base(f);
				for (int c = 0; c < h.len; ++c) {
					const int ea = (int)((PA[h.loi + c] >> (8 * f)) & 0xff) - fb;
					const int eb = (int)((PB[h.loj + c] >> (8 * f)) & 0xff) - fb;
					total += tf[ea * al + eb];
				}
			}
			s_mega[warp][k] = total;
		}
		__syncwarp();
		float mega_total = 0.0f, best_mega = 0.0f;
		int best_idx = 0;
		for (int k = 0; k < nchain; ++k) {  // every lane redundantly: sequential order matters for the float sum
			const float ms = s_mega[warp][k];
			if (ms > best_mega) { best_mega = ms; best_idx = k; }
			mega_total += ms;
		}
		if (!(mega_total < a.min_mega_hsp_score)) {
			// XDropHSP: best 8-mer of the best HSP (xdrophsp.cpp:62-98)
			const Hsp h = hsp[chain[best_idx]];
			const int K = 8;
			float bestmer = 0.0f;
			int bestms = -1;
			for (int ms = lane; ms + K <= h.len; ms += 32) {
				float sc = 0.0f;
				for (int k = 0; k < K; ++k)
					sc += subst(s_tab, PA[h.loi + ms + k], PB[h.loj + ms + k]);
				if (sc > bestmer) { bestmer = sc; bestms = ms; }  // first maximum within this lane's stride
			}
			// first maximum overall: highest score, then smallest start
#pragma unroll
			for (int o = 16; o >= 1; o >>= 1) {
				const float os = __shfl_xor_sync(kFull, bestmer, o);
				const int om = __shfl_xor_sync(kFull, bestms, o);
				if (om >= 0 && (os > bestmer || (os == bestmer && (bestms < 0 || om < bestms)))) { bestmer = os; bestms = om; }
			}
			uint32_t LoA = (uint32_t)(h.loi + h.len / 2), LoB = (uint32_t)(h.loj + h.len / 2);
			if (bestms >= 0 && bestmer > 0.0f) { LoA = (uint32_t)(h.loi + bestms); LoB = (uint32_t)(h.loj + bestms); }
			if (min(LoA, LoB) < (uint32_t)(K / 2)) { LoA += K / 2; LoB += K / 2; }
			out.valid = 1;
			out.lo_a = LoA;
			out.lo_b = LoB;
		}
	}
	if (lane == 0)
		a.seeds[pair] = out;
}

enum { XB_DM = 1, XB_IM = 2, XB_MD = 4, XB_MI = 8 };
constexpr uint32_t kNone = 0xffffffffu;

// Work list of the x-drop kernel: only (pair, direction) items whose seed is valid, ordered by the number of DP rows
// (4 bins, longest first) so that the 32 sequential DPs of a warp have similar lengths.  Most pairs of a long chain with
// an unrelated chain have no chain of HSPs at all; without the list their threads idle beside the few that work.
constexpr int kXBins = 4;
__device__ __forceinline__ int xdrop_bin(uint32_t rows) { return rows >= 512 ? 0 : rows >= 192 ? 1 : rows >= 64 ? 2 : 3; }
__device__ __forceinline__ uint32_t xdrop_rows(const MkfArgs &a, uint32_t pair, uint32_t dir, const MkfSeed &sd)
{
	return dir ? sd.lo_a : a.lenA[a.pair_a[pair]] - sd.lo_a;
}

// cnt[0..3] = items per bin, cnt[4..7] = fill cursors
__global__ void __launch_bounds__(256) mkf_bin_kernel(const MkfArgs a, const int fill)
{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= 2 * a.npairs)
		return;
	const uint32_t pair = t >> 1, dir = t & 1;
	const MkfSeed sd = a.seeds[pair];
	if (!fill) {
		a.xres[t].score = 0.0f;
		a.xres[t].path_len = 0;
	}
	if (!sd.valid)
		return;
	const int b = xdrop_bin(xdrop_rows(a, pair, dir, sd));
	if (!fill) {
		atomicAdd(a.xcnt + b, 1u);
	} else {
		uint32_t base = 0;
		for (int k = 0; k < b; ++k)
			base += a.xcnt[k];
		a.xwork[base + atomicAdd(a.xcnt + kXBins + b, 1u)] = t;
	}
}

// One thread per (pair, direction): xdropfwd.cpp:71-386.  dir 0 = forward from the seed, dir 1 = backward
// (XDropBwd: the same DP on mirrored coordinates, xdropbwd.cpp:16-26).
__device__ __forceinline__ void xdrop_item(const MkfArgs &a, const float *s_tab, const uint32_t t);

__global__ void __launch_bounds__(64) mkf_xdrop_kernel(const MkfArgs a)
{
	__shared__ float s_tab[RSK_TABLE_FLOATS];
	for (int k = threadIdx.x; k < RSK_TABLE_FLOATS; k += blockDim.x)
		s_tab[k] = a.tables[k];
	__syncthreads();
	const uint32_t n = a.xcnt[0] + a.xcnt[1] + a.xcnt[2] + a.xcnt[3];
	for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n; w += gridDim.x * blockDim.x)
		xdrop_item(a, s_tab, a.xwork[w]);
}

__device__ __forceinline__ void xdrop_item(const MkfArgs &a, const float *s_tab, const uint32_t t)
{
	const uint32_t pair = t >> 1, dir = t & 1;
	MkfXdrop &res = a.xres[t];
	const MkfSeed sd = a.seeds[pair];
	const uint32_t qa = a.pair_a[pair], tb_ = a.pair_b[pair];
	const uint64_t *PA = a.profA + a.offA[qa];
	const uint64_t *PB = a.profB + a.offB[tb_];
	const uint32_t LAf = a.lenA[qa], LBf = a.lenB[tb_];
	const uint32_t LA = dir ? sd.lo_a : LAf - sd.lo_a;  // rows of this DP
	const uint32_t LB = dir ? sd.lo_b : LBf - sd.lo_b;
	// scratch of this pair: [bwd region | fwd region]; each region = M row, D row, path staging, TB matrix
	unsigned char *base = a.scratch + a.scratch_off[pair];
	if (!dir)
		base += xdrop_region_bytes(sd.lo_a, sd.lo_b);
	float *Mbuf = reinterpret_cast<float *>(base);
	float *M = Mbuf + 1;
	float *Dr = Mbuf + (LB + 4);
	uint8_t *stage = reinterpret_cast<uint8_t *>(Mbuf + 2 * (LB + 4));
	const size_t W = (size_t)LB + 3;
	uint8_t *tbm = stage + (((size_t)LA + LB + 4 + 15) & ~(size_t)15);
	res.stage_off = (unsigned long long)(stage - a.scratch);
	const float open = a.open, ext = a.ext, X = a.x2;
	// mirrored coordinates for the backward pass: position p of the DP is original position lo - 1 - p
#define SUBST(pa, pb) (dir ? subst(s_tab, PA[sd.lo_a - 1 - (pa)], PB[sd.lo_b - 1 - (pb)]) : subst(s_tab, PA[sd.lo_a + (pa)], PB[sd.lo_b + (pb)]))
	if (LA == 1 || LB == 1) {  // xdropfwd.cpp:87-93
		const float sc = SUBST(0, 0);
		if (sc > 0) {
			stage[0] = 'M';
			res.path_len = 1;
		}
		res.score = sc;
		return;
	}
	const float absopen = -open, absext = -ext;
	M[-1] = kNegInf;
	Dr[0] = kNegInf;
	Dr[1] = kNegInf;
	float best = 0.0f;
	uint32_t besti = 0, bestj = 0;
	uint32_t prev_jlo = 0, prev_jhi = 0, jlo = 1, jhi = 1;
	float M0 = best;
	for (uint32_t i = 1; i <= LA; ++i) {
		if (jlo == prev_jlo) {
			M[jlo - 1] = kNegInf;
			Dr[jlo] = kNegInf;
		}
		uint32_t endj = min(prev_jhi + 1, LB);
		for (uint32_t j = endj + 1; j <= min(jhi + 1, LB); ++j) {
			M[j - 1] = kNegInf;
			Dr[j] = kNegInf;
		}
		uint32_t next_jlo = kNone, next_jhi = kNone;
		float I0 = kNegInf;
		uint8_t *row = tbm + (size_t)i * W;
		const uint64_t ea = dir ? PA[sd.lo_a - i] : PA[sd.lo_a + i - 1];
		for (uint32_t j = jlo; j <= jhi; ++j) {
			uint8_t bits = 0;
			const float saved = M0;
			float x = M0;
			const float dj = Dr[j];
			if (dj > x) { x = dj; bits = XB_DM; }
			if (I0 > x) { x = I0; bits = XB_IM; }
			M0 = M[j];
			const uint64_t eb = dir ? PB[sd.lo_b - j] : PB[sd.lo_b + j - 1];
			float s = subst(s_tab, ea, eb);
			s += x;
			M[j] = s;
			float h = s - best + X;
			if (h > 0) { next_jlo = min(next_jlo, j + 1); next_jhi = j + 1; }
			if (h > absopen) next_jlo = min(next_jlo, j);
			if (h > absext && j == jhi && jhi + 1 < LB) {
				++jhi;
				const uint32_t ne = max(min(jhi + 1, LB), endj);
				for (uint32_t j2 = endj + 1; j2 <= ne; ++j2) {
					if (j2 - 1 > j) M[j2 - 1] = kNegInf;
					Dr[j2] = kNegInf;
				}
				endj = ne;
			}
			if (s >= best) { best = s; besti = i; bestj = j; }
			if (j != jlo) {
				const float md = saved + open;
				float dn = Dr[j] + ext;
				if (md >= dn) { dn = md; bits |= XB_MD; }
				Dr[j] = dn;
				h = dn - best + X;
				if (h > 0) { next_jlo = min(next_jlo, j - 1); next_jhi = max(next_jhi, j - 1); }
			}
			const float mi = saved + open;
			I0 += ext;
			if (mi >= I0) { I0 = mi; bits |= XB_MI; }
			h = I0 - best + X;
			if (h > 0) { next_jlo = min(next_jlo, j + 1); next_jhi = max(next_jhi, j + 1); }
			if (h > absext && j == jhi && jhi + 1 < LB) {
				++jhi;
				const uint32_t ne = max(min(jhi + 1, LB), endj);
				for (uint32_t j2 = endj + 1; j2 <= ne; ++j2) {
					M[j2 - 1] = kNegInf;
					Dr[j2] = kNegInf;
				}
				endj = ne;
			}
			row[j] = bits;
		}
		if (jhi < LB) {
			const uint32_t j1 = jhi + 1;
			uint8_t b1 = 0;
			const float md = M0 + open;
			float dn = Dr[j1] + ext;
			if (md >= dn) { dn = md; b1 = XB_MD; }
			Dr[j1] = dn;
			row[j1] = b1;
		}
		if (next_jlo == kNone)
			break;
		prev_jlo = jlo; prev_jhi = jhi;
		jlo = next_jlo; jhi = next_jhi;
		if (jlo > LB) jlo = LB;
		if (jhi > LB) jhi = LB;
		if (jlo == prev_jlo) { M0 = kNegInf; Dr[jlo] = kNegInf; }
		else M0 = M[jlo - 1];
	}
#undef SUBST
	if (!(best > 0.0f))
		return;
	res.score = best;
	uint32_t i = besti, j = bestj, n = 0;
	int st = 0;  // 0 M, 1 D, 2 I
	for (;;) {
		stage[n++] = (uint8_t)(st == 0 ? 'M' : st == 1 ? 'D' : 'I');
		if (i == 1 || j == 1)
			break;
		int nx;
		if (st == 0) {
			const uint8_t c = tbm[(size_t)i * W + j];
			nx = (c & XB_DM) ? 1 : (c & XB_IM) ? 2 : 0;
			--i; --j;
		} else if (st == 1) {
			nx = (tbm[(size_t)i * W + j + 1] & XB_MD) ? 0 : 1;
			--i;
		} else {
			nx = (tbm[(size_t)(i + 1) * W + j] & XB_MI) ? 0 : 2;
			--j;
		}
		st = nx;
	}
	res.path_len = n;  // stage[] holds the path from its end to its start
}

// The same DP as xdrop_item, flattened into a per-lane state machine: every trip of the loop computes ONE cell of the lane's
// current item (or one traceback step, or the row / item bookkeeping), and a lane that has finished its item takes the next
// one from the work list inside the same loop.  The 32 lanes of a warp therefore stay in the cell code together whatever
// the lengths and band widths of their items are, and no warp waits for a straggler: on real SCOP40 data the
// thread-per-item kernel ran with 7.8 of 32 lanes active and 9 % of the warp slots occupied (profiles/r1_mkf_xdrop_ncu.md).
// Cells are visited in exactly the order of xdropfwd.cpp:71-386 with the same operations, so results are unchanged.
__global__ void __launch_bounds__(64) mkf_xdrop_flat_kernel(const MkfArgs a)
{
	__shared__ float s_tab[RSK_TABLE_FLOATS];
	for (int k = threadIdx.x; k < RSK_TABLE_FLOATS; k += blockDim.x)
		s_tab[k] = a.tables[k];
	__syncthreads();
	const uint32_t nwork = a.xcnt[0] + a.xcnt[1] + a.xcnt[2] + a.xcnt[3];
	const float open = a.open, ext = a.ext, X = a.x2;
	const float absopen = -open, absext = -ext;
	enum { P_FETCH, P_ROW_START, P_CELL, P_ROW_END, P_TB_INIT, P_TB_STEP };
	int phase = P_FETCH;
	// item
	uint32_t t = 0, dir = 0, lo_a = 0, lo_b = 0, LA = 0, LB = 0;
	const uint64_t *PA = nullptr, *PB = nullptr;
	float *M = nullptr, *Dr = nullptr;
	uint8_t *stage = nullptr, *tbm = nullptr, *row = nullptr;
	size_t W = 0;
	// DP
	uint32_t i = 0, j = 0, jlo = 0, jhi = 0, prev_jlo = 0, prev_jhi = 0, next_jlo = 0, next_jhi = 0, endj = 0, besti = 0, bestj = 0;
	float best = 0.0f, M0 = 0.0f, I0 = 0.0f;
	uint64_t ea = 0;
	// traceback
	uint32_t ti = 0, tj = 0, tn = 0;
	int st = 0;
	for (;;) {
		if (phase == P_FETCH) {
			const uint32_t w = atomicAdd(a.xcnt + 2 * kXBins, 1u);
			if (w >= nwork)
				break;
			t = a.xwork[w];
			const uint32_t pair = t >> 1;
			dir = t & 1;
			const MkfSeed sd = a.seeds[pair];
			lo_a = sd.lo_a; lo_b = sd.lo_b;
			const uint32_t qa = a.pair_a[pair], tb_ = a.pair_b[pair];
			PA = a.profA + a.offA[qa];
			PB = a.profB + a.offB[tb_];
			LA = dir ? lo_a : a.lenA[qa] - lo_a;
			LB = dir ? lo_b : a.lenB[tb_] - lo_b;
			unsigned char *base = a.scratch + a.scratch_off[pair];
			if (!dir)
				base += xdrop_region_bytes(lo_a, lo_b);
			float *Mbuf = reinterpret_cast<float *>(base);
			M = Mbuf + 1;
			Dr = Mbuf + (LB + 4);
			stage = reinterpret_cast<uint8_t *>(Mbuf + 2 * (LB + 4));
			W = (size_t)LB + 3;
			tbm = stage + (((size_t)LA + LB + 4 + 15) & ~(size_t)15);
			a.xres[t].stage_off = (unsigned long long)(stage - a.scratch);
			if (LA == 1 || LB == 1) {  // xdropfwd.cpp:87-93
				const float sc = dir ? subst(s_tab, PA[lo_a - 1], PB[lo_b - 1]) : subst(s_tab, PA[lo_a], PB[lo_b]);
				if (sc > 0) {
					stage[0] = 'M';
					a.xres[t].path_len = 1;
				}
				a.xres[t].score = sc;
				continue;  // next item
			}
			M[-1] = kNegInf;
			Dr[0] = kNegInf;
			Dr[1] = kNegInf;
			best = 0.0f; besti = 0; bestj = 0;
			prev_jlo = 0; prev_jhi = 0; jlo = 1; jhi = 1;
			M0 = best;
			i = 1;
			phase = P_ROW_START;
		}
		if (phase == P_ROW_START) {
			if (jlo == prev_jlo) {
				M[jlo - 1] = kNegInf;
				Dr[jlo] = kNegInf;
			}
			endj = min(prev_jhi + 1, LB);
			for (uint32_t j2 = endj + 1; j2 <= min(jhi + 1, LB); ++j2) {
				M[j2 - 1] = kNegInf;
				Dr[j2] = kNegInf;
			}
			next_jlo = kNone; next_jhi = kNone;
			I0 = kNegInf;
			row = tbm + (size_t)i * W;
			ea = dir ? PA[lo_a - i] : PA[lo_a + i - 1];
			j = jlo;
			phase = P_CELL;
		}
		if (phase == P_CELL) {
			uint8_t bits = 0;
			const float saved = M0;
			float x = M0;
			const float dj = Dr[j];
			if (dj > x) { x = dj; bits = XB_DM; }
			if (I0 > x) { x = I0; bits = XB_IM; }
			M0 = M[j];
			const uint64_t eb = dir ? PB[lo_b - j] : PB[lo_b + j - 1];
			float s = subst(s_tab, ea, eb);
			s += x;
			M[j] = s;
			float h = s - best + X;
			if (h > 0) { next_jlo = min(next_jlo, j + 1); next_jhi = j + 1; }
			if (h > absopen) next_jlo = min(next_jlo, j);
			if (h > absext && j == jhi && jhi + 1 < LB) {
				++jhi;
				const uint32_t ne = max(min(jhi + 1, LB), endj);
				for (uint32_t j2 = endj + 1; j2 <= ne; ++j2) {
					if (j2 - 1 > j) M[j2 - 1] = kNegInf;
					Dr[j2] = kNegInf;
				}
				endj = ne;
			}
			if (s >= best) { best = s; besti = i; bestj = j; }
			if (j != jlo) {
				const float md = saved + open;
				float dn = Dr[j] + ext;
				if (md >= dn) { dn = md; bits |= XB_MD; }
				Dr[j] = dn;
				h = dn - best + X;
				if (h > 0) { next_jlo = min(next_jlo, j - 1); next_jhi = max(next_jhi, j - 1); }
			}
			const float mi = saved + open;
			I0 += ext;
			if (mi >= I0) { I0 = mi; bits |= XB_MI; }
			h = I0 - best + X;
			if (h > 0) { next_jlo = min(next_jlo, j + 1); next_jhi = max(next_jhi, j + 1); }
			if (h > absext && j == jhi && jhi + 1 < LB) {
				++jhi;
				const uint32_t ne = max(min(jhi + 1, LB), endj);
				for (uint32_t j2 = endj + 1; j2 <= ne; ++j2) {
					M[j2 - 1] = kNegInf;
					Dr[j2] = kNegInf;
				}
				endj = ne;
			}
			row[j] = bits;
			++j;
			if (j > jhi)
				phase = P_ROW_END;
		}
		if (phase == P_ROW_END) {
			if (jhi < LB) {
				const uint32_t j1 = jhi + 1;
				uint8_t b1 = 0;
				const float md = M0 + open;
				float dn = Dr[j1] + ext;
				if (md >= dn) { dn = md; b1 = XB_MD; }
				Dr[j1] = dn;
				row[j1] = b1;
			}
			if (next_jlo == kNone) {
				phase = P_TB_INIT;
			} else {
				prev_jlo = jlo; prev_jhi = jhi;
				jlo = next_jlo; jhi = next_jhi;
				if (jlo > LB) jlo = LB;
				if (jhi > LB) jhi = LB;
				if (jlo == prev_jlo) { M0 = kNegInf; Dr[jlo] = kNegInf; }
				else M0 = M[jlo - 1];
				++i;
				phase = (i > LA) ? P_TB_INIT : P_ROW_START;
			}
		}
		if (phase == P_TB_INIT) {
			if (!(best > 0.0f)) {
				phase = P_FETCH;
			} else {
				a.xres[t].score = best;
				ti = besti; tj = bestj; tn = 0; st = 0;
				phase = P_TB_STEP;
			}
		}
		if (phase == P_TB_STEP) {
			stage[tn++] = (uint8_t)(st == 0 ? 'M' : st == 1 ? 'D' : 'I');
			if (ti == 1 || tj == 1) {
				a.xres[t].path_len = tn;  // stage[] holds the path from its end to its start
				phase = P_FETCH;
			} else {
				int nx;
				if (st == 0) {
					const uint8_t c = tbm[(size_t)ti * W + tj];
					nx = (c & XB_DM) ? 1 : (c & XB_IM) ? 2 : 0;
					--ti; --tj;
				} else if (st == 1) {
					nx = (tbm[(size_t)ti * W + tj + 1] & XB_MD) ? 0 : 1;
					--ti;
				} else {
					nx = (tbm[(size_t)(ti + 1) * W + tj] & XB_MI) ? 0 : 2;
					--tj;
				}
				st = nx;
			}
		}
	}
}

// The same DP with ONE WARP per (pair, direction): the lanes take 32 consecutive band columns of a row at a time.
//
// What makes xdropfwd.cpp sequential is (1) the insert state I0, a chain of float additions of Ext along the row, (2) the
// running best score, which every x-drop test reads in row-major order, and (3) the band itself: a row can grow to the right
// while its last cell is being computed, and the next row's extent is only known when the row is complete.  None of this is
// changed; the cells of a row are only evaluated side by side:
//  * I0 entering column j+1 is max(mi_j, I0_j + Ext) with mi_j = M[i-1][j-1] + Open, which depends on the PREVIOUS row only.
//    f(x) = fl(x + Ext) is monotone, so I0 is a max-plus prefix scan; a Hillis-Steele step of distance 2^d applies f exactly
//    2^d times (sequential float adds, never x + 2^d*Ext), which gives bit-for-bit the value of the serial chain.
//  * the best score before a cell is an (exact) prefix maximum; "s >= best" updates resolve to the last lane that fires.
//  * columns beyond the current right edge are evaluated speculatively (their previous-row inputs are -inf by the reference's
//    own initialisation rule) and the row ends at the first edge cell whose extension test fails.
//  * next_jlo is a minimum; next_jhi is folded as the reference does it: "= j+1" on a match test, max(...) otherwise, with the
//    UINT_MAX start value sticking until a match test fires.
// Previous-row values are read through the extents of the previous row (M valid on [prev_jlo, prev_jhi], D on (prev_jlo,
// prev_jhi+1]); everything else is -inf, which is what the reference's in-place initialisation amounts to.
// The traceback walks runs: 32 lanes fetch the next 32 trace cells along the current direction at once.
constexpr int kXWarps = 4;
__global__ void __launch_bounds__(kXWarps * 32) mkf_xdrop_warp_kernel(const MkfArgs a)
{
	__shared__ float s_tab[RSK_TABLE_FLOATS];
	for (int k = threadIdx.x; k < RSK_TABLE_FLOATS; k += blockDim.x)
		s_tab[k] = a.tables[k];
	__syncthreads();
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t nwork = a.xcnt[0] + a.xcnt[1] + a.xcnt[2] + a.xcnt[3];
	const float open = a.open, ext = a.ext, X = a.x2;
	const float absopen = -open, absext = -ext;
	for (;;) {
		uint32_t w = 0;
		if (lane == 0)
			w = atomicAdd(a.xcnt + 2 * kXBins, 1u);
		w = __shfl_sync(kFull, w, 0);
		if (w >= nwork)
			break;
		const uint32_t t = a.xwork[w];
		const uint32_t pair = t >> 1, dir = t & 1;
		const MkfSeed sd = a.seeds[pair];
		const uint32_t lo_a = sd.lo_a, lo_b = sd.lo_b;
		const uint32_t qa = a.pair_a[pair], tb_ = a.pair_b[pair];
		const uint64_t *PA = a.profA + a.offA[qa];
		const uint64_t *PB = a.profB + a.offB[tb_];
		const uint32_t LA = dir ? lo_a : a.lenA[qa] - lo_a;
		const uint32_t LB = dir ? lo_b : a.lenB[tb_] - lo_b;
		unsigned char *base = a.scratch + a.scratch_off[pair];
		if (!dir)
			base += xdrop_region_bytes(lo_a, lo_b);
		float *Mbuf = reinterpret_cast<float *>(base);
		float *M = Mbuf + 1;
		float *Dr = Mbuf + (LB + 4);
		uint8_t *stage = reinterpret_cast<uint8_t *>(Mbuf + 2 * (LB + 4));
		const size_t W = (size_t)LB + 3;
		uint8_t *tbm = stage + (((size_t)LA + LB + 4 + 15) & ~(size_t)15);
		if (lane == 0)
			a.xres[t].stage_off = (unsigned long long)(stage - a.scratch);
		if (LA == 1 || LB == 1) {  // xdropfwd.cpp:87-93
			if (lane == 0) {
				const float sc = dir ? subst(s_tab, PA[lo_a - 1], PB[lo_b - 1]) : subst(s_tab, PA[lo_a], PB[lo_b]);
				if (sc > 0) {
					stage[0] = 'M';
					a.xres[t].path_len = 1;
				}
				a.xres[t].score = sc;
			}
			continue;
		}
		if (lane == 0) {
			M[-1] = kNegInf;
			Dr[0] = kNegInf;
			Dr[1] = kNegInf;
		}
		__syncwarp();
		float best = 0.0f;
		uint32_t besti = 0, bestj = 0;
		uint32_t prev_jlo = 0, prev_jhi = 0, jlo = 1, jhi = 1;
		float diag0 = 0.0f;  // M0 = BestScore before the first row
		for (uint32_t i = 1; i <= LA; ++i) {
			const uint64_t ea = dir ? PA[lo_a - i] : PA[lo_a + i - 1];
			uint8_t *row = tbm + (size_t)i * W;
			uint32_t next_jlo = kNone, next_jhi = kNone;
			float I_carry = kNegInf, diag_carry = diag0, best_run = best;
			uint32_t jhi_cur = jhi, j0 = jlo, jhi_final = jhi;
			float M0_end = kNegInf;
			// A quirk of the reference that has to be kept: `endj` is not advanced by the row-start initialisation
			// (xdropfwd.cpp:142-147), so when the row starts wider than the previous one and then extends at its last start
			// column jstar, the FIRST extension's initialisation loop (j2 = prev_jhi+2 ...) runs over columns of THIS row that
			// are already done: the match-site extension (:190-205, Mrow guarded) resets Drow[prev_jhi+2 .. jstar] (Drow[jstar] is
			// then recomputed from -inf by the cell's delete part), the insert-site extension (:262-276, no guard) resets
			// Drow[prev_jhi+2 .. jstar] and Mrow[prev_jhi+1 .. jstar].  The next row sees those -inf values.
			const uint32_t jstar = jhi;
			const bool quirk_row = jstar >= prev_jhi + 1;
			bool wipe = false, wipe_site1 = false;
			for (;;) {
				const uint32_t j = j0 + lane;
				const bool inb = j <= LB;
				const float oldM = (inb && j >= prev_jlo && j <= prev_jhi) ? M[j] : kNegInf;
				const float oldD = (inb && j > prev_jlo && j <= prev_jhi + 1) ? Dr[j] : kNegInf;
				float diag = __shfl_up_sync(kFull, oldM, 1);
				if (lane == 0)
					diag = diag_carry;
				const float mi = diag + open;
				// insert state after each cell: prefix scan of max(mi, f(.)), f applied by repeated addition
				float A = mi;
				if (lane == 0) {
					const float ti0 = I_carry + ext;
					A = (mi >= ti0) ? mi : ti0;
				}
#pragma unroll
				for (int d = 0; d < 5; ++d) {
					float v = __shfl_up_sync(kFull, A, 1u << d);
#pragma unroll
					for (int r = 0; r < (1 << d); ++r)
						v += ext;
					if (lane >= (1u << d))
						A = fmaxf(A, v);
				}
				float I0 = __shfl_up_sync(kFull, A, 1);
				if (lane == 0)
					I0 = I_carry;
				const float tI = I0 + ext;
				const bool bMI = mi >= tI;
				const float Inew = bMI ? mi : tI;
				// match state
				float x = diag;
				uint32_t bits = 0;
				if (oldD > x) { x = oldD; bits = XB_DM; }
				if (I0 > x) { x = I0; bits = XB_IM; }
				float s = kNegInf;
				if (inb) {
					const uint64_t eb = dir ? PB[lo_b - j] : PB[lo_b + j - 1];
					s = subst(s_tab, ea, eb);
					s += x;
				}
				// best score before / after each cell, row-major
				float pm = s;
#pragma unroll
				for (int d = 0; d < 5; ++d) {
					const float v = __shfl_up_sync(kFull, pm, 1u << d);
					if (lane >= (1u << d))
						pm = fmaxf(pm, v);
				}
				const float bb = __shfl_up_sync(kFull, pm, 1);
				const float best_before = (lane == 0) ? best_run : fmaxf(best_run, bb);
				const float best_after = fmaxf(best_before, s);
				const float h1 = s - best_before + X;
				// delete state
				const bool hasD = j != jlo;
				const float md = diag + open;
				float dn = oldD + ext;
				bool bMD = false;
				if (md >= dn) { dn = md; bMD = true; }
				float h2 = dn - best_after + X;
				const float h3 = Inew - best_after + X;
				// where does the row end?  an edge cell (j >= current jhi) extends the row iff one of its insert tests fires
				const bool E = (h1 > absext || h3 > absext) && (j + 1 < LB);
				if (quirk_row && jstar >= j0 && jstar < j0 + 32) {
					const uint32_t ls = jstar - j0;
					if (__shfl_sync(kFull, (int)E, ls)) {
						wipe = true;
						wipe_site1 = __shfl_sync(kFull, (int)(h1 > absext), ls) != 0;
						if (wipe_site1 && jstar >= prev_jhi + 2 && lane == ls) {  // Drow[jstar] was reset before the delete part of the cell
							dn = kNegInf + ext;
							bMD = false;
							if (md >= dn) { dn = md; bMD = true; }
							h2 = dn - best_after + X;
						}
					}
				}
				const unsigned stop = __ballot_sync(kFull, !inb || (j >= jhi_cur && !E));
				const bool last_chunk = stop != 0;
				const uint32_t f = last_chunk ? (uint32_t)__ffs((int)stop) - 1u : 31u;
				const bool valid = lane <= f;
				if (valid) {
					M[j] = s;
					uint32_t bt = bits | (bMI ? XB_MI : 0u);
					if (hasD) {
						Dr[j] = dn;
						if (bMD)
							bt |= XB_MD;
					}
					row[j] = (uint8_t)bt;
				}
				const unsigned upd = __ballot_sync(kFull, valid && s >= best_before);
				if (upd) {
					bestj = j0 + (31u - (uint32_t)__clz((int)upd));
					besti = i;
				}
				best_run = fmaxf(best_run, __shfl_sync(kFull, pm, f));
				uint32_t lo_c = kNone, mx = 0;
				if (valid) {
					if (h1 > 0) lo_c = j + 1;
					if (h1 > absopen) lo_c = min(lo_c, j);
					if (hasD && h2 > 0) { lo_c = min(lo_c, j - 1); mx = j - 1; }
					if (h3 > 0) { lo_c = min(lo_c, j + 1); mx = j + 1; }
				}
				next_jlo = min(next_jlo, __reduce_min_sync(kFull, lo_c));
				const unsigned mA = __ballot_sync(kFull, valid && h1 > 0);
				if (mA) {
					const uint32_t lA = 31u - (uint32_t)__clz((int)mA);
					const uint32_t m = __reduce_max_sync(kFull, lane >= lA ? mx : 0u);
					next_jhi = max(j0 + lA + 1, m);
				} else {
					const uint32_t m = __reduce_max_sync(kFull, mx);
					if (m)
						next_jhi = max(next_jhi, m);  // UINT_MAX sticks, as in the reference
				}
				if (last_chunk) {
					jhi_final = j0 + f;
					M0_end = __shfl_sync(kFull, oldM, f);
					break;
				}
				I_carry = __shfl_sync(kFull, Inew, 31);
				diag_carry = __shfl_sync(kFull, oldM, 31);
				if (j0 + 32 > jhi_cur)
					jhi_cur = j0 + 32;
				j0 += 32;
			}
			if (jhi_final < LB && lane == 0) {  // special case for the end of Drow[] (xdropfwd.cpp:289-300)
				const uint32_t j1 = jhi_final + 1;
				const float pd = (j1 > prev_jlo && j1 <= prev_jhi + 1) ? Dr[j1] : kNegInf;
				const float md = M0_end + open;
				float dn = pd + ext;
				uint8_t b1 = 0;
				if (md >= dn) { dn = md; b1 = XB_MD; }
				Dr[j1] = dn;
				row[j1] = b1;
			}
			if (wipe) {
				__syncwarp();  // after the row's own stores
				for (uint32_t jj = prev_jhi + 2 + lane; jj <= (wipe_site1 ? jstar - 1 : jstar); jj += 32)
					Dr[jj] = kNegInf;
				if (!wipe_site1)
					for (uint32_t jj = prev_jhi + 1 + lane; jj <= jstar; jj += 32)
						M[jj] = kNegInf;
			}
			best = best_run;
			if (next_jlo == kNone)
				break;
			prev_jlo = jlo; prev_jhi = jhi_final;
			jlo = min(next_jlo, LB);
			jhi = min(next_jhi, LB);
			__syncwarp();  // this row's stores before the next row's loads
			diag0 = (jlo == prev_jlo) ? kNegInf : M[jlo - 1];
		}
		if (!(best > 0.0f))
			continue;
		__syncwarp();
		// traceback by runs (TraceBack, xdropfwd.cpp:14-67): state st at (ti, tj); the lanes read the next 32 trace cells along
		// the direction of the state and the run ends at the first cell that changes state or touches row/column 1
		uint32_t ti = besti, tj = bestj, tn = 0;
		int st = 0;
		for (;;) {
			const uint32_t di = (st != 2) ? lane : 0u, dj = (st != 1) ? lane : 0u;
			const bool inr = ti > di && tj > dj;
			const uint32_t pi = ti - di, pj = tj - dj;
			const bool boundary = inr && (pi == 1 || pj == 1);
			int nx = st;
			if (inr && !boundary) {
				if (st == 0) {
					const uint8_t c = tbm[(size_t)pi * W + pj];
					nx = (c & XB_DM) ? 1 : (c & XB_IM) ? 2 : 0;
				} else if (st == 1) {
					nx = (tbm[(size_t)pi * W + pj + 1] & XB_MD) ? 0 : 1;
				} else {
					nx = (tbm[(size_t)(pi + 1) * W + pj] & XB_MI) ? 0 : 2;
				}
			}
			const unsigned endm = __ballot_sync(kFull, !inr || boundary || nx != st);
			const uint32_t e = endm ? (uint32_t)__ffs((int)endm) - 1u : 31u;
			if (lane <= e)
				stage[tn + lane] = (uint8_t)(st == 0 ? 'M' : st == 1 ? 'D' : 'I');
			tn += e + 1;
			if (!endm) {  // the run goes on
				ti -= (st != 2) ? 32u : 0u;
				tj -= (st != 1) ? 32u : 0u;
				continue;
			}
			const bool fin = __shfl_sync(kFull, (int)boundary, e) != 0;
			if (fin)
				break;
			const int nst = __shfl_sync(kFull, nx, e);
			const uint32_t ei = ti - ((st != 2) ? e : 0u), ej = tj - ((st != 1) ? e : 0u);
			ti = (st != 2) ? ei - 1 : ei;
			tj = (st != 1) ? ej - 1 : ej;
			st = nst;
		}
		if (lane == 0) {
			a.xres[t].score = best;
			a.xres[t].path_len = tn;  // stage[] holds the path from its end to its start
		}
	}
}

// One warp per pair: total score test, MergeFwdBwd, publish record + path.
__global__ void __launch_bounds__(128) mkf_finish_kernel(const MkfArgs a)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t pair = blockIdx.x * 4 + warp;
	if (pair >= a.npairs)
		return;
	PairRec *rec = a.rec + a.pair_slot[pair];
	const MkfSeed sd = a.seeds[pair];
	if (lane == 0) {
		rec->flags = RSK_HIT_MKF;
		rec->mu_fwd = sd.best_hsp;
		rec->mu_rev = sd.best_chain;
	}
	if (!sd.valid) {  // ClearAlign state: score 0, no path (also undoes a record a non-filtered SW pass may have written)
		if (lane == 0) {
			rec->score = 0.0f;
			rec->path_len = 0;
			rec->lo_a = rec->lo_b = 0xffffffffu;
		}
		return;
	}
	const MkfXdrop f = a.xres[2 * pair], b = a.xres[2 * pair + 1];
	const float total = f.score + b.score;
	if (total < 10) {  // xdrophsp.cpp:109-113
		if (lane == 0) {
			rec->score = 0.0f;
			rec->path_len = 0;
			rec->lo_a = rec->lo_b = 0xffffffffu;
		}
		return;
	}
	const uint8_t *fs = a.scratch + f.stage_off, *bs = a.scratch + b.stage_off;
	const uint32_t nf = f.path_len, nb = b.path_len, n = nf + nb;
	unsigned long long off = 0;
	if (lane == 0)
		off = atomicAdd(a.pool_cursor, (unsigned long long)n);
	off = __shfl_sync(kFull, off, 0);
	// backward DP: its own forward order is reverse(stage); XDropBwd reverses once more -> stage order as is.
	uint32_t bM = 0, bD = 0;
	for (uint32_t k = lane; k < nb; k += 32) {
		const uint8_t c = bs[k];
		a.pool[off + k] = c;
		bM += (c == 'M');
		bD += (c == 'D');
	}
	for (uint32_t k = lane; k < nf; k += 32)
		a.pool[off + nb + k] = fs[nf - 1 - k];
#pragma unroll
	for (int o = 16; o >= 1; o >>= 1) {
		bM += __shfl_xor_sync(kFull, bM, o);
		bD += __shfl_xor_sync(kFull, bD, o);
	}
	if (lane == 0) {
		const uint32_t bI = nb - bM - bD;
		rec->score = total;
		rec->lo_a = nb ? sd.lo_a - (bM + bD) : sd.lo_a;  // mergefwdback.cpp:36-49
		rec->lo_b = nb ? sd.lo_b - (bM + bI) : sd.lo_b;
		rec->path_len = n;
		rec->path_off = off;
	}
}

}  // namespace

int launch_mkf(const MkfArgs &args, uint32_t nhash, int xgrid_blocks, cudaStream_t stream)
{
	if (args.npairs == 0)
		return 0;
	mkf_hash_kernel<<<nhash, 256, 0, stream>>>(args);
	mkf_seed_kernel<<<(args.npairs + kSeedWarps - 1) / kSeedWarps, kSeedWarps * 32, 0, stream>>>(args);
	if (cudaMemsetAsync(args.xcnt, 0, (2 * kXBins + 1) * sizeof(uint32_t), stream) != cudaSuccess)
		return -1;
	mkf_bin_kernel<<<(2 * args.npairs + 255) / 256, 256, 0, stream>>>(args, 0);
	mkf_bin_kernel<<<(2 * args.npairs + 255) / 256, 256, 0, stream>>>(args, 1);
	const unsigned xblocks = (unsigned)std::min<uint64_t>(((uint64_t)2 * args.npairs + 63) / 64, (uint64_t)xgrid_blocks);
	// RSK_XDROP=seq | flat: the round-1 kernels (one thread per item), kept for A/B timing and as a cross-check of the warp kernel
	static const char *variant = getenv("RSK_XDROP");
	if (variant && !strcmp(variant, "seq"))
		mkf_xdrop_kernel<<<xblocks, 64, 0, stream>>>(args);
	else if (variant && !strcmp(variant, "flat"))
		mkf_xdrop_flat_kernel<<<xblocks, 64, 0, stream>>>(args);
	else {
		const unsigned wblocks = (unsigned)std::min<uint64_t>(((uint64_t)2 * args.npairs + kXWarps - 1) / kXWarps, (uint64_t)xgrid_blocks);
		mkf_xdrop_warp_kernel<<<wblocks, kXWarps * 32, 0, stream>>>(args);
	}
	mkf_finish_kernel<<<(args.npairs + 3) / 4, 128, 0, stream>>>(args);
	return cudaGetLastError() == cudaSuccess ? 6 : -1;
}

size_t mkf_hash_bytes() { return (size_t)kDict * kHashW * sizeof(uint16_t); }

}  // namespace rsk
