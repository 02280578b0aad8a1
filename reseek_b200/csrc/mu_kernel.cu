// mu_kernel.cu - K3: Mu-letter int8 Smith-Waterman pre-filter (score only, forward and reversed query),
// plus the survivor compaction that turns the filter's verdicts into SW tasks.
//
// Replaces DSSAligner::MuFilter (dssaligner.cpp:619-630) -> AlignMuQP_Para (parasail_mu.cpp:120-161) ->
// parasail_sw_striped_profile_avx2_256_8 (parasail.cpp:515-797) and the profile construction
// SetMuQP_Para (parasail_mu.cpp:163-181).  The striped int8 kernel with its lazy-F correction computes the
// plain Gotoh local score with every value floored at 0 (the int8 lanes are biased by -128 and saturate
// downwards); "saturated" means the best score exceeded 250 (parasail.cpp:585,728-733).  We evaluate the
// same recurrence as a warp wavefront in 32-bit integers, so no saturation can occur internally and the
// flag is simply best > 250.  Decision rule (parasail_mu.cpp:133-160):
//   fwd = saturated ? 777 : score(A,B);  if fwd < omega_fwd -> 0;
//   rev = saturated ? 255 : score(reverse(A),B);  score = fwd - rev;  keep the pair iff score >= omega.
//
// Layout: like K1 the CTA owns one A chain and each warp one B chain.  For the current pass of 32*R rows
// the CTA stages T[b][lane][r] = IntScoreMx_Mu[a(row)][b] as int32 so a lane gets its R row scores for
// column letter b with two conflict-free LDS.128.
#include "rsk_internal.cuh"

namespace rsk {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kMuR = 8;              // rows per lane
constexpr int kMuRows = 32 * kMuR;   // rows per pass
constexpr int kMuLetters = 36;

// smem: int mx[36*36] | int4 T[36][2][32] | bcast
constexpr size_t kMuSmemMx = 0;
constexpr size_t kMuSmemT = 36 * 36 * 4;                                    // 5184
constexpr size_t kMuSmemBcast = kMuSmemT + (size_t)kMuLetters * 2 * 32 * 16;  // + 36864
constexpr size_t kMuSmemTotal = kMuSmemBcast + 16;

__device__ __forceinline__ int max3(int a, int b, int c) { return max(max(a, b), c); }

// One pass of the Gotoh recurrence: rows [pass*256, pass*256+256) of A (possibly reversed) vs all of B.
// bnd[j] = (H, F) leaving the last row of the pass at column j.
__device__ __forceinline__ int mu_pass(const int4 *__restrict__ T, const int lane, const bool first, const bool last,
		const uint8_t *__restrict__ colB, const int LB, int2 *__restrict__ bnd, const int open, const int ext)
{
	int H[kMuR], E[kMuR];  // H[i][j-1] (previous column), E[i][j]
#pragma unroll
	for (int r = 0; r < kMuR; ++r) {
		H[r] = 0;
		E[r] = 0;
	}
	int best = 0;
	int hdiag_next = 0;  // H[i0-1][j-1] for the lane's first row
	int outH = 0, outF = 0;
	const int nsteps = LB + 31;
	int j = -lane;
	int cb = (j >= 0 && j < LB) ? (int)colB[j] : 0;
	int2 bn = make_int2(0, 0);
	if (lane == 0 && !first)
		bn = bnd[0];
	for (int s = 0; s < nsteps; ++s, ++j) {
		const int inH = __shfl_up_sync(kFull, outH, 1);
		const int inF = __shfl_up_sync(kFull, outF, 1);
		const int jn = j + 1;
		const int cb_next = (jn >= 0 && jn < LB) ? (int)colB[jn] : 0;
		int2 bn_next = bn;
		if (lane == 0 && !first && jn < LB)
			bn_next = bnd[jn];
		if (j >= 0 && j < LB) {
			int f, hd = hdiag_next;
			if (lane == 0) {
				f = first ? 0 : bn.y;           // F entering row i0 at column j
				hdiag_next = first ? 0 : bn.x;  // H[i0-1][j]
			} else {
				f = inF;
				hdiag_next = inH;
			}
			const int4 s0 = T[(cb * 2 + 0) * 32 + lane];
			const int4 s1 = T[(cb * 2 + 1) * 32 + lane];
			const int sc[kMuR] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
			for (int r = 0; r < kMuR; ++r) {
				const int x = max3(hd + sc[r], E[r], 0);
				const int h = max(x, f);
				hd = H[r];  // H[i][j-1] is the diagonal of row i+1
				H[r] = h;
				best = max(best, h);
				const int ho = h - open;
				E[r] = max3(E[r] - ext, ho, 0);
				f = max3(f - ext, ho, 0);
			}
			outH = H[kMuR - 1];
			outF = f;
			if (lane == 31 && !last)
				bnd[j] = make_int2(outH, outF);
		}
		cb = cb_next;
		bn = bn_next;
	}
	return best;
}

__device__ __forceinline__ void build_mu_table(int *T32, const int *mx, const uint8_t *__restrict__ muA, const int LA,
		const int pass, const bool reversed, const bool tr)
{
	// T[b][h][lane][q]: row rr = lane*8 + h*4 + q of this pass
	for (int idx = threadIdx.x; idx < kMuLetters * kMuRows; idx += kSwThreads) {
		const int b = idx / kMuRows;
		const int rr = idx - b * kMuRows;
		const int row = pass * kMuRows + rr;
		int v = -1000;  // rows beyond the chain: never contribute (every value is floored at 0)
		if (row < LA) {
			const int a = muA[reversed ? (LA - 1 - row) : row];
			v = tr ? mx[b * kMuLetters + a] : mx[a * kMuLetters + b];  // matrix[reference A letter][reference B letter]
		}
		const int l = rr >> 3, r = rr & 7;
		T32[(((b * 2 + (r >> 2)) * 32 + l) << 2) + (r & 3)] = v;
	}
}

__global__ void __launch_bounds__(kSwThreads, 2) mu_sw_filter_kernel(const MuArgs a)
{
	extern __shared__ __align__(16) unsigned char smem[];
	int *mx = reinterpret_cast<int *>(smem + kMuSmemMx);
	int *T32 = reinterpret_cast<int *>(smem + kMuSmemT);
	const int4 *T = reinterpret_cast<const int4 *>(smem + kMuSmemT);
	volatile int *bcast = reinterpret_cast<volatile int *>(smem + kMuSmemBcast);
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (int k = threadIdx.x; k < 36 * 36; k += kSwThreads)
		mx[k] = a.mu_mx[k];
	__syncthreads();
	const size_t gw = (size_t)blockIdx.x * kSwWarps + warp;
	int2 *bnd = a.bnd + gw * a.bnd_stride;
	for (;;) {
		if (threadIdx.x == 0)
			bcast[0] = (int)atomicAdd(a.task_counter, 1u);
		__syncthreads();
		const uint32_t task = (uint32_t)bcast[0];
		__syncthreads();
		if (task >= a.ntasks)
			break;
		uint32_t rowchain, begin, cnt;
		if (a.cross) {
			const uint32_t ridx = task / a.nseg;
			const uint32_t seg = task - ridx * a.nseg;
			rowchain = a.rowlist[ridx];
			begin = seg * kSwWarps;
			cnt = min((uint32_t)kSwWarps, a.ncols - begin);
		} else {
			rowchain = a.task_row[task];
			begin = a.task_begin[task];
			cnt = a.task_cnt[task];
		}
		const int LA = (int)a.len_row[rowchain];
		const uint8_t *muA = a.mu_row + a.off_row[rowchain];
		const int npass = (LA + kMuRows - 1) / kMuRows;
		const bool have = (uint32_t)warp < cnt;
		uint32_t cidx = 0;
		int LB = 0;
		const uint8_t *colB = nullptr;
		if (have) {
			cidx = a.clist[begin + warp];
			LB = (int)a.len_col[cidx];
			colB = a.mu_col + a.off_col[cidx];
		}
		int fwd = 0, rev = 0;
		bool need_rev = false;
		// DoMKF() pairs are not the filter's business (dssaligner.cpp:811-815 returns before the filter)
		// (k-mers exist only for chains of >= 3 residues: dss.cpp:659-682)
		const bool mkf_a = false;
		const bool mkf = have && LA >= 3 && LB >= 3 && ((uint32_t)LA >= a.mkfl || (uint32_t)LB >= a.mkfl);
		const bool run = have && !mkf;
		for (int dir = 0; dir < 2 && !mkf_a; ++dir) {
			if (dir == 1) {
				// reversed pass only when some warp of the CTA still needs it
				need_rev = run && !((float)fwd < a.omega_fwd);
				if (!__syncthreads_or(need_rev ? 1 : 0))
					break;
			}
			int best = 0;
			for (int pass = 0; pass < npass; ++pass) {
				__syncthreads();
				build_mu_table(T32, mx, muA, LA, pass, dir == 1, a.tr != 0);
				__syncthreads();
				if (run && (dir == 0 || need_rev))
					best = max(best, mu_pass(T, lane, pass == 0, pass == npass - 1, colB, LB, bnd, a.open, a.ext));
			}
#pragma unroll
			for (int o = 16; o >= 1; o >>= 1)
				best = max(best, __shfl_xor_sync(kFull, best, o));
			if (dir == 0)
				fwd = best > 250 ? 777 : best;  // parasail_mu.cpp:133-137
			else
				rev = best > 250 ? 255 : best;  // value read before the 777 assignment (:149-155)
		}
		if (have && lane == 0) {
			uint32_t slot;
			if (a.cross) {
				const uint32_t ra = a.tr ? cidx : rowchain, rb = a.tr ? rowchain : cidx;
				slot = (ra - a.a_begin) * a.nB + rb;
			} else {
				slot = a.cslot[begin + warp];
			}
			PairRec *rec = a.rec + slot;
			float score = 0.0f;
			int rrev = 0;
			if (need_rev) {
				score = (float)fwd - (float)rev;
				rrev = rev;
			}
			rec->mu_fwd = fwd;
			rec->mu_rev = rrev;
			const bool pass_ = !mkf && !(score < a.omega);  // dssaligner.cpp:627
			rec->flags = mkf ? (uint32_t)RSK_HIT_MKF : pass_ ? 0u : (uint32_t)RSK_HIT_MU_REJECTED;
			a.keep[slot] = pass_ ? 1 : 0;
			if (fwd == 777)
				atomicAdd(a.sat_counter, 1u);
		}
	}
}

// ---------------------------------------------------------------------------------------------------------------
// K3 v2: the same filter on packed 16-bit lanes.  A warp carries TWO column chains, one in each half of every 32-bit
// register, and the recurrence runs on the DPX instructions of sm_100a (VIADDMNMX.S16x2.RELU, VIMNMX3.S16x2,
// VIADD.16x2): 7 instructions per row for two cells instead of ~10 per cell.  Rows per lane are chosen per row chain
// (R = 2..12, 32*R rows per pass) so that a 300-residue chain runs in one pass of 320 rows instead of two of 256.
// Values are bounded by 4*min(LA, LB) (best substitution score +4, mumx_data.cpp:42), so int16 is exact for chains up to
// 8 000 residues; the host routes longer ones to the 32-bit kernel above.  Columns past the end of the shorter chain of a
// warp read a padding letter (score -1000): every state of such a column is bounded by an earlier H minus a gap
// penalty, so the running maximum is unaffected.
// smem: int mx[36*36] | short T[37][2][32][8] (plane 0: rows 0..7 of each lane, plane 1: rows 8..15) | bcast
constexpr int kMu16Letters = 37;  // 36 + padding letter
constexpr size_t kMu16SmemT = 36 * 36 * 4;
constexpr size_t kMu16SmemBcast = kMu16SmemT + (size_t)kMu16Letters * 2 * 32 * 16;
constexpr size_t kMu16SmemTotal = kMu16SmemBcast + 16;

template <int R>
__device__ __forceinline__ unsigned mu16_pass(const uint4 *__restrict__ T, const int lane, const bool first, const bool last,
		const uint8_t *__restrict__ colB0, const int LB0, const uint8_t *__restrict__ colB1, const int LB1,
		uint2 *__restrict__ bnd, const unsigned nopen, const unsigned next)
{
	unsigned H[R], E[R];  // H[i][j-1] (previous column), E[i][j]; pair 0 in the low half, pair 1 in the high half
#pragma unroll
	for (int r = 0; r < R; ++r) {
		H[r] = 0;
		E[r] = 0;
	}
	unsigned best = 0, hdiag_next = 0, outH = 0, outF = 0;
	const int LBm = max(LB0, LB1);
	const int nsteps = LBm + 31;
	int j = -lane;
	int cb0 = (j >= 0 && j < LB0) ? (int)colB0[j] : 36;
	int cb1 = (j >= 0 && j < LB1) ? (int)colB1[j] : 36;
	uint2 bn = make_uint2(0, 0);
	if (lane == 0 && !first)
		bn = bnd[0];
	for (int s = 0; s < nsteps; ++s, ++j) {
		const unsigned inH = __shfl_up_sync(kFull, outH, 1);
		const unsigned inF = __shfl_up_sync(kFull, outF, 1);
		const int jn = j + 1;
		const int cb0n = (jn >= 0 && jn < LB0) ? (int)colB0[jn] : 36;
		const int cb1n = (jn >= 0 && jn < LB1) ? (int)colB1[jn] : 36;
		uint2 bn_next = bn;
		if (lane == 0 && !first && jn < LBm)
			bn_next = bnd[jn];
		if (j >= 0 && j < LBm) {
			unsigned f, hd = hdiag_next;
			if (lane == 0) {
				f = first ? 0u : bn.y;           // F entering row i0 at column j
				hdiag_next = first ? 0u : bn.x;  // H[i0-1][j]
			} else {
				f = inF;
				hdiag_next = inH;
			}
			unsigned sc[R];
			{
				const uint4 a = T[(cb0 * 2 + 0) * 32 + lane];
				const uint4 b = T[(cb1 * 2 + 0) * 32 + lane];
				const unsigned av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
				for (int r = 0; r < R && r < 8; ++r)
					sc[r] = __byte_perm(av[r >> 1], bv[r >> 1], (r & 1) ? 0x7632 : 0x5410);
			}
			if (R > 8) {
				const uint4 a = T[(cb0 * 2 + 1) * 32 + lane];
				const uint4 b = T[(cb1 * 2 + 1) * 32 + lane];
				const unsigned av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
				for (int r = 8; r < R; ++r)
					sc[r] = __byte_perm(av[(r - 8) >> 1], bv[(r - 8) >> 1], (r & 1) ? 0x7632 : 0x5410);
			}
#pragma unroll
			for (int r = 0; r < R; ++r) {
				const unsigned x = __viaddmax_s16x2_relu(hd, sc[r], E[r]);  // max(H[i-1][j-1] + s, E, 0)
				const unsigned h = __vmaxs2(x, f);
				best = __vimax3_s16x2(best, x, f);
				hd = H[r];  // H[i][j-1] is the diagonal of row i+1
				H[r] = h;
				E[r] = __viaddmax_s16x2_relu(h, nopen, __vadd2(E[r], next));  // max(E - ext, h - open, 0)
				f = __viaddmax_s16x2_relu(h, nopen, __vadd2(f, next));
			}
			outH = H[R - 1];
			outF = f;
			if (lane == 31 && !last)
				bnd[j] = make_uint2(outH, outF);
		}
		cb0 = cb0n;
		cb1 = cb1n;
		bn = bn_next;
	}
	return best;
}

__device__ __forceinline__ void build_mu16_table(short *T16, const int *mx, const uint8_t *__restrict__ muA, const int LA,
		const int pass, const int R, const bool reversed, const bool tr)
{
	// T[b][plane][lane][q]: row rr = lane*R + plane*8 + q of this pass, q < 8; unused slots are never read
	const int rows = 32 * R;
	for (int idx = threadIdx.x; idx < kMu16Letters * rows; idx += kSwThreads) {
		const int b = idx / rows;
		const int rr = idx - b * rows;
		const int row = pass * rows + rr;
		int v = -1000;  // rows beyond the chain and the padding letter: never contribute (every value is floored at 0)
		if (row < LA && b < kMuLetters) {
			const int a = muA[reversed ? (LA - 1 - row) : row];
			v = tr ? mx[b * kMuLetters + a] : mx[a * kMuLetters + b];  // matrix[reference A letter][reference B letter]
		}
		const int l = rr / R, r = rr - l * R;
		T16[(((b * 2 + (r >> 3)) * 32 + l) << 3) + (r & 7)] = (short)v;
	}
}

template <int R>
__device__ __forceinline__ unsigned mu16_dir(const MuArgs &a, short *T16, const int *mx, const uint8_t *muA, const int LA,
		const int npass, const bool reversed, const bool run, const int lane, const uint8_t *colB0, const int LB0,
		const uint8_t *colB1, const int LB1, uint2 *bnd)
{
	const uint4 *T = reinterpret_cast<const uint4 *>(T16);
	const unsigned no = (unsigned)(unsigned short)(short)(-a.open), ne = (unsigned)(unsigned short)(short)(-a.ext);
	const unsigned nopen = no | (no << 16), next = ne | (ne << 16);
	unsigned best = 0;
	for (int pass = 0; pass < npass; ++pass) {
		__syncthreads();
		build_mu16_table(T16, mx, muA, LA, pass, R, reversed, a.tr != 0);
		__syncthreads();
		if (run)
			best = __vmaxs2(best, mu16_pass<R>(T, lane, pass == 0, pass == npass - 1, colB0, LB0, colB1, LB1, bnd, nopen, next));
	}
#pragma unroll
	for (int o = 16; o >= 1; o >>= 1)
		best = __vmaxs2(best, __shfl_xor_sync(kFull, best, o));
	return best;
}

// rows per lane and passes for a row chain of LA residues (R even, <= 12: 24 state registers + 12 scores fit the
// 64-register budget of two 16-warp CTAs per SM)
__host__ __device__ inline void mu16_geometry(int LA, int &npass, int &R)
{
	npass = (LA + 383) / 384;
	if (npass < 1) npass = 1;
	R = (LA + 32 * npass - 1) / (32 * npass);
	R = (R + 1) & ~1;
	if (R < 2) R = 2;
}

__global__ void __launch_bounds__(kSwThreads, 2) mu_sw_filter16_kernel(const MuArgs a)
{
	extern __shared__ __align__(16) unsigned char smem[];
	int *mx = reinterpret_cast<int *>(smem + kMuSmemMx);
	short *T16 = reinterpret_cast<short *>(smem + kMu16SmemT);
	volatile int *bcast = reinterpret_cast<volatile int *>(smem + kMu16SmemBcast);
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (int k = threadIdx.x; k < 36 * 36; k += kSwThreads)
		mx[k] = a.mu_mx[k];
	__syncthreads();
	const size_t gw = (size_t)blockIdx.x * kSwWarps + warp;
	uint2 *bnd = reinterpret_cast<uint2 *>(a.bnd + gw * a.bnd_stride);
	for (;;) {
		if (threadIdx.x == 0)
			bcast[0] = (int)atomicAdd(a.task_counter, 1u);
		__syncthreads();
		const uint32_t task = (uint32_t)bcast[0];
		__syncthreads();
		if (task >= a.ntasks)
			break;
		uint32_t rowchain, begin, cnt;
		if (a.cross) {
			const uint32_t ridx = task / a.nseg;
			const uint32_t seg = task - ridx * a.nseg;
			rowchain = a.rowlist[ridx];
			begin = seg * kMuTaskCols;
			cnt = min((uint32_t)kMuTaskCols, a.ncols - begin);
		} else {
			rowchain = a.task_row[task];
			begin = a.task_begin[task];
			cnt = a.task_cnt[task];
		}
		const int LA = (int)a.len_row[rowchain];
		const uint8_t *muA = a.mu_row + a.off_row[rowchain];
		int npass, R;
		mu16_geometry(LA, npass, R);
		// this warp's two column chains
		bool have[2], runp[2], mkf[2];
		uint32_t cidx[2] = {0, 0};
		int LB[2] = {0, 0};
		const uint8_t *colB[2] = {muA, muA};  // never dereferenced when LB = 0
#pragma unroll
		for (int p = 0; p < 2; ++p) {
			have[p] = (uint32_t)(2 * warp + p) < cnt;
			mkf[p] = false;
			if (have[p]) {
				cidx[p] = a.clist[begin + 2 * warp + p];
				LB[p] = (int)a.len_col[cidx[p]];
				colB[p] = a.mu_col + a.off_col[cidx[p]];
				// DoMKF() pairs are not the filter's business (dssaligner.cpp:811-815 returns before the filter)
				mkf[p] = LA >= 3 && LB[p] >= 3 && ((uint32_t)LA >= a.mkfl || (uint32_t)LB[p] >= a.mkfl);
			}
			runp[p] = have[p] && !mkf[p];
		}
		const int LBr[2] = {runp[0] ? LB[0] : 0, runp[1] ? LB[1] : 0};
		int fwd[2] = {0, 0}, rev[2] = {0, 0};
		bool need_rev[2] = {false, false};
		for (int dir = 0; dir < 2; ++dir) {
			bool run = runp[0] || runp[1];
			if (dir == 1) {
				// reversed pass only when some warp of the CTA still needs it
				need_rev[0] = runp[0] && !((float)fwd[0] < a.omega_fwd);
				need_rev[1] = runp[1] && !((float)fwd[1] < a.omega_fwd);
				run = need_rev[0] || need_rev[1];
				if (!__syncthreads_or(run ? 1 : 0))
					break;
			}
			const int L0 = (dir == 0 || need_rev[0]) ? LBr[0] : 0, L1 = (dir == 0 || need_rev[1]) ? LBr[1] : 0;
			unsigned best;
			switch (R) {
			case 2: best = mu16_dir<2>(a, T16, mx, muA, LA, npass, dir == 1, run, lane, colB[0], L0, colB[1], L1, bnd); break;
			case 4: best = mu16_dir<4>(a, T16, mx, muA, LA, npass, dir == 1, run, lane, colB[0], L0, colB[1], L1, bnd); break;
			case 6: best = mu16_dir<6>(a, T16, mx, muA, LA, npass, dir == 1, run, lane, colB[0], L0, colB[1], L1, bnd); break;
			case 8: best = mu16_dir<8>(a, T16, mx, muA, LA, npass, dir == 1, run, lane, colB[0], L0, colB[1], L1, bnd); break;
			case 10: best = mu16_dir<10>(a, T16, mx, muA, LA, npass, dir == 1, run, lane, colB[0], L0, colB[1], L1, bnd); break;
			default: best = mu16_dir<12>(a, T16, mx, muA, LA, npass, dir == 1, run, lane, colB[0], L0, colB[1], L1, bnd); break;
			}
			const int b0 = (int)(short)(best & 0xffffu), b1 = (int)(short)(best >> 16);
			if (dir == 0) {
				fwd[0] = b0 > 250 ? 777 : b0;  // parasail_mu.cpp:133-137
				fwd[1] = b1 > 250 ? 777 : b1;
			} else {
				rev[0] = b0 > 250 ? 255 : b0;  // value read before the 777 assignment (:149-155)
				rev[1] = b1 > 250 ? 255 : b1;
			}
		}
		if (lane < 2 && have[lane]) {
			const int p = lane;
			uint32_t slot;
			if (a.cross) {
				const uint32_t ra = a.tr ? cidx[p] : rowchain, rb = a.tr ? rowchain : cidx[p];
				slot = (ra - a.a_begin) * a.nB + rb;
			} else {
				slot = a.cslot[begin + 2 * warp + p];
			}
			PairRec *rec = a.rec + slot;
			float score = 0.0f;
			int rrev = 0;
			if (need_rev[p]) {
				score = (float)fwd[p] - (float)rev[p];
				rrev = rev[p];
			}
			rec->mu_fwd = fwd[p];
			rec->mu_rev = rrev;
			const bool pass_ = !mkf[p] && !(score < a.omega);  // dssaligner.cpp:627
			rec->flags = mkf[p] ? (uint32_t)RSK_HIT_MKF : pass_ ? 0u : (uint32_t)RSK_HIT_MU_REJECTED;
			a.keep[slot] = pass_ ? 1 : 0;
			if (fwd[p] == 777)
				atomicAdd(a.sat_counter, 1u);
		}
	}
}

// One CTA per row chain of the batch: compact the surviving column chains (in clist order, so lengths stay
// sorted) and emit SW tasks of up to W pairs into the task list of the row chain's kernel class.
__global__ void __launch_bounds__(256) compact_survivors_kernel(const CompactArgs a)
{
	__shared__ uint32_t s_warp[8];
	__shared__ uint32_t s_taskbase;
	const uint32_t ridx = blockIdx.x;
	const uint32_t rowchain = a.rowlist[ridx];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	uint32_t running = 0;
	unsigned long long lsum = 0;
	for (uint32_t k0 = 0; k0 < a.ncols; k0 += 256) {
		const uint32_t k = k0 + tid;
		uint32_t c = 0, keep = 0, slot = 0;
		if (k < a.ncols) {
			c = a.clist[k];
			const uint32_t ra = a.tr ? c : rowchain, rb = a.tr ? rowchain : c;
			slot = (ra - a.a_begin) * a.nB + rb;
			keep = a.keep[slot];
			if (keep)
				lsum += a.len_col[c];
		}
		const unsigned m = __ballot_sync(kFull, keep != 0);
		if (lane == 0)
			s_warp[warp] = __popc(m);
		__syncthreads();
		uint32_t before = 0, tot = 0;
		for (int w = 0; w < 8; ++w) {
			if (w < warp)
				before += s_warp[w];
			tot += s_warp[w];
		}
		if (keep) {
			const uint32_t pos = running + before + __popc(m & ((1u << lane) - 1u));
			a.out_clist[(size_t)ridx * a.ncols + pos] = c;
			a.out_cslot[(size_t)ridx * a.ncols + pos] = slot;
		}
		running += tot;
		__syncthreads();
	}
	const int cls = sw_class_of_len(a.len_row[rowchain]);
	const uint32_t W = (uint32_t)sw_class_warps(cls) * sw_class_chains(cls);  // column chains per SW task
	const uint32_t ntask = (running + W - 1) / W;
#pragma unroll
	for (int o = 16; o >= 1; o >>= 1)
		lsum += __shfl_xor_sync(kFull, lsum, o);
	if (lane == 0 && lsum)
		atomicAdd(a.cell_count, lsum * (unsigned long long)a.len_row[rowchain]);
	if (tid == 0) {
		s_taskbase = ntask ? atomicAdd(a.task_count + cls, ntask) : 0;
		atomicAdd(a.pair_count, (unsigned long long)running);
	}
	__syncthreads();
	const size_t tb = (size_t)cls * a.task_cap + s_taskbase;
	for (uint32_t t = tid; t < ntask; t += 256) {
		a.task_row[tb + t] = rowchain;
		a.task_begin[tb + t] = ridx * a.ncols + t * W;
		a.task_cnt[tb + t] = min(W, running - t * W);
	}
}

// The same for explicit pair lists (PostMuFilter, RunSelf): one CTA per run of pairs with the same row chain (slots
// [begin, begin + cnt) of the batch, column chains in a.clist, longest first); the survivors are compacted inside the run's own
// slot range and cut into SW tasks of up to W pairs.  With this the host never reads the keep flags: the batch's kernels are
// queued without a round trip (the host used to wait for the Mu filter, build the task lists and upload them).
__global__ void __launch_bounds__(256) compact_explicit_kernel(const CompactArgs a)
{
	__shared__ uint32_t s_warp[8];
	__shared__ uint32_t s_taskbase;
	const uint32_t ridx = blockIdx.x;
	const uint32_t rowchain = a.run_row[ridx], begin = a.run_begin[ridx], cnt = a.run_cnt[ridx];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	uint32_t running = 0;
	unsigned long long lsum = 0;
	for (uint32_t k0 = 0; k0 < cnt; k0 += 256) {
		const uint32_t k = k0 + tid;
		uint32_t c = 0, keep = 0;
		if (k < cnt) {
			c = a.clist[begin + k];
			keep = a.keep[begin + k];
			if (keep)
				lsum += a.len_col[c];
		}
		const unsigned m = __ballot_sync(kFull, keep != 0);
		if (lane == 0)
			s_warp[warp] = __popc(m);
		__syncthreads();
		uint32_t before = 0, tot = 0;
		for (int w = 0; w < 8; ++w) {
			if (w < warp)
				before += s_warp[w];
			tot += s_warp[w];
		}
		if (keep) {
			const uint32_t pos = begin + running + before + __popc(m & ((1u << lane) - 1u));
			a.out_clist[pos] = c;
			a.out_cslot[pos] = begin + k;
		}
		running += tot;
		__syncthreads();
	}
	const int cls = sw_class_of_len(a.len_row[rowchain]);
	const uint32_t W = (uint32_t)sw_class_warps(cls) * sw_class_chains(cls);  // column chains per SW task
	const uint32_t ntask = (running + W - 1) / W;
#pragma unroll
	for (int o = 16; o >= 1; o >>= 1)
		lsum += __shfl_xor_sync(kFull, lsum, o);
	if (lane == 0 && lsum)
		atomicAdd(a.cell_count, lsum * (unsigned long long)a.len_row[rowchain]);
	if (tid == 0) {
		s_taskbase = ntask ? atomicAdd(a.task_count + cls, ntask) : 0;
		atomicAdd(a.pair_count, (unsigned long long)running);
	}
	__syncthreads();
	const size_t tb = (size_t)cls * a.task_cap + s_taskbase;
	for (uint32_t t = tid; t < ntask; t += 256) {
		a.task_row[tb + t] = rowchain;
		a.task_begin[tb + t] = begin + t * W;
		a.task_cnt[tb + t] = min(W, running - t * W);
	}
}

}  // namespace

size_t mu_smem_bytes() { return kMuSmemTotal; }

int launch_mu_filter16(const MuArgs &args, int grid, cudaStream_t stream)
{
	if (cudaFuncSetAttribute(mu_sw_filter16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMu16SmemTotal) != cudaSuccess)
		return -1;
	mu_sw_filter16_kernel<<<grid, kSwThreads, kMu16SmemTotal, stream>>>(args);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_mu_filter(const MuArgs &args, int grid, cudaStream_t stream)
{
	if (cudaFuncSetAttribute(mu_sw_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMuSmemTotal) != cudaSuccess)
		return -1;
	mu_sw_filter_kernel<<<grid, kSwThreads, kMuSmemTotal, stream>>>(args);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_compact_explicit(const CompactArgs &args, uint32_t nruns, cudaStream_t stream)
{
	if (nruns == 0)
		return 0;
	compact_explicit_kernel<<<nruns, 256, 0, stream>>>(args);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_compact_survivors(const CompactArgs &args, uint32_t nrows, cudaStream_t stream)
{
	if (nrows == 0)
		return 0;
	compact_survivors_kernel<<<nrows, 256, 0, stream>>>(args);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace rsk
