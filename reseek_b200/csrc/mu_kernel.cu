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

// G = 32: the warp is one wavefront over rows [pass*32R, (pass+1)*32R).  G = 16 (row chains of <= 192 residues, one pass):
// the two half-warps are independent 16-lane wavefronts over the same rows, each with its own packed pair of column chains
// (colB0/colB1/LB0/LB1 are then per half, LBm the warp-wide loop bound) - the ramp is 15 steps instead of 31 and no lane idles
// on rows beyond the chain: a 174-residue chain fills 91 % of the lane-steps instead of 77 %.
template <int R, int G>
__device__ __forceinline__ unsigned mu16_pass(const uint4 *__restrict__ T, const int lane, const bool first, const bool last,
		const uint8_t *__restrict__ colB0, const int LB0, const uint8_t *__restrict__ colB1, const int LB1, const int LBmax,
		uint2 *__restrict__ bnd, const unsigned nopen, const unsigned next)
{
	const int sub = lane & (G - 1);
	unsigned H[R], E[R];  // H[i][j-1] (previous column), E[i][j]; pair 0 in the low half, pair 1 in the high half
#pragma unroll
	for (int r = 0; r < R; ++r) {
		H[r] = 0;
		E[r] = 0;
	}
	unsigned best = 0, hdiag_next = 0, outH = 0, outF = 0;
	const int LBm = max(LB0, LB1);     // this wavefront's columns
	const int nsteps = (G == 32 ? LBm : LBmax) + G - 1;  // warp-uniform (G = 16: the longer of the two half-warps' columns)
	int j = -sub;
	int cb0 = (j >= 0 && j < LB0) ? (int)colB0[j] : 36;
	int cb1 = (j >= 0 && j < LB1) ? (int)colB1[j] : 36;
	uint2 bn = make_uint2(0, 0);
	if (G == 32 && lane == 0 && !first)
		bn = bnd[0];
	for (int s = 0; s < nsteps; ++s, ++j) {
		const unsigned inH = __shfl_up_sync(kFull, outH, 1, G);
		const unsigned inF = __shfl_up_sync(kFull, outF, 1, G);
		const int jn = j + 1;
		const int cb0n = (jn >= 0 && jn < LB0) ? (int)colB0[jn] : 36;
		const int cb1n = (jn >= 0 && jn < LB1) ? (int)colB1[jn] : 36;
		uint2 bn_next = bn;
		if (G == 32 && lane == 0 && !first && jn < LBm)
			bn_next = bnd[jn];
		if (j >= 0 && j < LBm) {
			unsigned f, hd = hdiag_next;
			if (sub == 0) {
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
			if (G == 32 && lane == 31 && !last)
				bnd[j] = make_uint2(outH, outF);
		}
		cb0 = cb0n;
		cb1 = cb1n;
		bn = bn_next;
	}
	return best;
}

__device__ __forceinline__ void build_mu16_table(short *T16, const int *mx, const uint8_t *__restrict__ muA, const int LA,
		const int pass, const int R, const bool reversed, const bool tr, const bool half)
{
	// T[b][plane][lane][q]: row lane*R + plane*8 + q of this pass, q < 8; unused slots are never read.  Half-warp mode:
	// lane l holds the rows of sub-lane l & 15 (both half-warps read the same rows, each lane its own 16 bytes).
	// Division-free: thread (l, g) of the CTA owns lane slot l and the letters b = g, g + 16, g + 32; it reads the letters of the
	// slot's R rows once and writes whole 16-byte slots.  (The first form mapped a flat index to (b, lane, row) with two runtime
	// divisions per 2-byte element - ~7 % of a task's time for 300-residue chains.)
	constexpr int kGroups = kSwThreads / 32;
	const int l = threadIdx.x & 31, g = threadIdx.x >> 5;
	const int row0 = half ? (l & 15) * R : (pass * 32 + l) * R;
	int arow[12];
#pragma unroll
	for (int r = 0; r < 12; ++r) {
		const int row = row0 + r;
		arow[r] = (r < R && row < LA) ? (int)muA[reversed ? (LA - 1 - row) : row] : -1;
	}
	uint4 *T4 = reinterpret_cast<uint4 *>(T16);
	for (int b = g; b < kMu16Letters; b += kGroups) {
		unsigned w[8];
#pragma unroll
		for (int k = 0; k < 6; ++k) {
			int v[2];
#pragma unroll
			for (int h = 0; h < 2; ++h) {
				const int a = arow[2 * k + h];
				// rows beyond the chain and the padding letter never contribute (every value is floored at 0)
				v[h] = (a >= 0 && b < kMuLetters) ? (tr ? mx[b * kMuLetters + a] : mx[a * kMuLetters + b]) : -1000;  // matrix[ref A letter][ref B letter]
			}
			w[k] = ((unsigned)v[0] & 0xffffu) | ((unsigned)v[1] << 16);
		}
		w[6] = w[7] = 0xfc18fc18u;  // -1000, -1000
		T4[(b * 2 + 0) * 32 + l] = make_uint4(w[0], w[1], w[2], w[3]);
		if (R > 8)
			T4[(b * 2 + 1) * 32 + l] = make_uint4(w[4], w[5], w[6], w[7]);
	}
}

// rows per lane and passes for a row chain of LA residues (R even, <= 12: 24 state registers + 12 scores fit the
// 64-register budget of two 16-warp CTAs per SM); chains of <= 192 residues run as two 16-lane wavefronts per warp
__host__ __device__ inline void mu16_geometry(int LA, int &npass, int &R, bool &half)
{
	half = LA <= 16 * 12;
	const int lanes = half ? 16 : 32;
	npass = half ? 1 : (LA + 383) / 384;
	if (npass < 1) npass = 1;
	R = (LA + lanes * npass - 1) / (lanes * npass);
	R = (R + 1) & ~1;
	if (R < 2) R = 2;
}

template <int R, int G>
__device__ __forceinline__ unsigned mu16_run(const uint4 *T, const int lane, const bool first, const bool last, const uint8_t *c0,
		const int L0, const uint8_t *c1, const int L1, const int LBmax, uint2 *bnd, const unsigned nopen, const unsigned next)
{
	return mu16_pass<R, G>(T, lane, first, last, c0, L0, c1, L1, LBmax, bnd, nopen, next);
}

// one pass of one round: dispatch on the rows per lane
template <int G>
__device__ __forceinline__ unsigned mu16_pass_R(const int R, const uint4 *T, const int lane, const bool first, const bool last,
		const uint8_t *c0, const int L0, const uint8_t *c1, const int L1, const int LBmax, uint2 *bnd, const unsigned nopen,
		const unsigned next)
{
	switch (R) {
	case 2: return mu16_run<2, G>(T, lane, first, last, c0, L0, c1, L1, LBmax, bnd, nopen, next);
	case 4: return mu16_run<4, G>(T, lane, first, last, c0, L0, c1, L1, LBmax, bnd, nopen, next);
	case 6: return mu16_run<6, G>(T, lane, first, last, c0, L0, c1, L1, LBmax, bnd, nopen, next);
	case 8: return mu16_run<8, G>(T, lane, first, last, c0, L0, c1, L1, LBmax, bnd, nopen, next);
	case 10: return mu16_run<10, G>(T, lane, first, last, c0, L0, c1, L1, LBmax, bnd, nopen, next);
	default: return mu16_run<12, G>(T, lane, first, last, c0, L0, c1, L1, LBmax, bnd, nopen, next);
	}
}

// A task = one row chain x a segment of the column list.  Row chains of <= 192 residues: 64 column chains per task, warp w
// takes columns 4w .. 4w+3 as two packed pairs that run side by side on its two half-warps.  Longer row chains: 32 column
// chains per task, warp w takes columns 2w, 2w+1 as one packed pair on the whole warp.  (Cross mode numbers its segments in
// units of 32 columns; a short row chain uses the even ones, each covering two units.)
__global__ void __launch_bounds__(kSwThreads, 2) mu_sw_filter16_kernel(const MuArgs a)
{
	extern __shared__ __align__(16) unsigned char smem[];
	int *mx = reinterpret_cast<int *>(smem + kMuSmemMx);
	short *T16 = reinterpret_cast<short *>(smem + kMu16SmemT);
	const uint4 *T = reinterpret_cast<const uint4 *>(T16);
	volatile int *bcast = reinterpret_cast<volatile int *>(smem + kMu16SmemBcast);
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (int k = threadIdx.x; k < 36 * 36; k += kSwThreads)
		mx[k] = a.mu_mx[k];
	__syncthreads();
	const size_t gw = (size_t)blockIdx.x * kSwWarps + warp;
	uint2 *bnd0 = reinterpret_cast<uint2 *>(a.bnd + gw * (size_t)a.bnd_stride);
	const unsigned no = (unsigned)(unsigned short)(short)(-a.open), ne = (unsigned)(unsigned short)(short)(-a.ext);
	const unsigned nopen = no | (no << 16), next = ne | (ne << 16);
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
			begin = seg * (kMuTaskCols / 2);
			cnt = a.ncols - begin;
		} else {
			rowchain = a.task_row[task];
			begin = a.task_begin[task];
			cnt = a.task_cnt[task];
		}
		const int LA = (int)a.len_row[rowchain];
		const uint8_t *muA = a.mu_row + a.off_row[rowchain];
		int npass, R;
		bool half;
		mu16_geometry(LA, npass, R, half);
		if (a.cross) {
			if (half && (task - (task / a.nseg) * a.nseg) % 2 != 0)
				continue;  // covered by the even segment before it
			cnt = min(cnt, (uint32_t)(half ? kMuTaskCols : kMuTaskCols / 2));
		}
		const int per = half ? 4 : 2;  // column chains per warp
		// this warp's column chains: pairs (0, 1) and - half-warp mode - (2, 3)
		bool have[4], runp[4], mkf[4];
		uint32_t cidx[4] = {0, 0, 0, 0};
		int LB[4] = {0, 0, 0, 0};
		const uint8_t *colB[4] = {muA, muA, muA, muA};  // never dereferenced when LB = 0
#pragma unroll
		for (int p = 0; p < 4; ++p) {
			have[p] = p < per && (uint32_t)(per * warp + p) < cnt;
			mkf[p] = false;
			if (have[p]) {
				cidx[p] = a.clist[begin + per * warp + p];
				LB[p] = (int)a.len_col[cidx[p]];
				colB[p] = a.mu_col + a.off_col[cidx[p]];
				// DoMKF() pairs are not the filter's business (dssaligner.cpp:811-815 returns before the filter)
				mkf[p] = LA >= 3 && LB[p] >= 3 && ((uint32_t)LA >= a.mkfl || (uint32_t)LB[p] >= a.mkfl);
			}
			runp[p] = have[p] && !mkf[p];
		}
		int fwd[4] = {0, 0, 0, 0}, rev[4] = {0, 0, 0, 0};
		bool need_rev[4] = {false, false, false, false};
		for (int dir = 0; dir < 2; ++dir) {
			int Ld[4];  // columns of each chain in this direction (0 = not run)
			bool any = false;
#pragma unroll
			for (int p = 0; p < 4; ++p) {
				if (dir == 1)
					need_rev[p] = runp[p] && !((float)fwd[p] < a.omega_fwd);
				Ld[p] = (dir == 0 ? runp[p] : need_rev[p]) ? LB[p] : 0;
				any = any || Ld[p] > 0;
			}
			// reversed pass only when some warp of the CTA still needs it
			if (dir == 1 && !__syncthreads_or(any ? 1 : 0))
				break;
			unsigned best[1] = {0};
			const int LBw = max(max(Ld[0], Ld[1]), max(Ld[2], Ld[3]));
			for (int pass = 0; pass < npass; ++pass) {
				__syncthreads();
				build_mu16_table(T16, mx, muA, LA, pass, R, dir == 1, a.tr != 0, half);
				__syncthreads();
				if (!any)
					continue;
				if (half) {
					const bool h = lane >= 16;  // explicit selects: a runtime index would put the arrays into local memory
					best[0] = mu16_pass_R<16>(R, T, lane, true, true, h ? colB[2] : colB[0], h ? Ld[2] : Ld[0], h ? colB[3] : colB[1],
							h ? Ld[3] : Ld[1], LBw, nullptr, nopen, next);
				} else {
					best[0] = __vmaxs2(best[0], mu16_pass_R<32>(R, T, lane, pass == 0, pass == npass - 1, colB[0], Ld[0], colB[1], Ld[1], LBw,
							bnd0, nopen, next));
				}
			}
			// packed maxima of the two pairs: full mode = reduced over the warp per round, half mode = over each half-warp
			unsigned b01, b23;
			if (half) {
				unsigned b = best[0];
#pragma unroll
				for (int o = 8; o >= 1; o >>= 1)
					b = __vmaxs2(b, __shfl_xor_sync(kFull, b, o));
				b01 = __shfl_sync(kFull, b, 0);
				b23 = __shfl_sync(kFull, b, 16);
			} else {
				b01 = best[0];
				b23 = 0;
#pragma unroll
				for (int o = 16; o >= 1; o >>= 1)
					b01 = __vmaxs2(b01, __shfl_xor_sync(kFull, b01, o));
			}
			const int bv[4] = {(int)(short)(b01 & 0xffffu), (int)(short)(b01 >> 16), (int)(short)(b23 & 0xffffu), (int)(short)(b23 >> 16)};
#pragma unroll
			for (int p = 0; p < 4; ++p) {
				if (dir == 0)
					fwd[p] = bv[p] > 250 ? 777 : bv[p];  // parasail_mu.cpp:133-137
				else
					rev[p] = bv[p] > 250 ? 255 : bv[p];  // value read before the 777 assignment (:149-155)
			}
		}
		if (lane < 4 && have[lane]) {
			const int p = lane;
			uint32_t slot;
			if (a.cross) {
				const uint32_t ra = a.tr ? cidx[p] : rowchain, rb = a.tr ? rowchain : cidx[p];
				slot = (ra - a.a_begin) * a.nB + rb;
			} else {
				slot = a.cslot[begin + per * warp + p];
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
	// column chains per SW task: every warp of the CTA gets one chain before any warp gets a second one, and a row chain with few
	// survivors is cut into more, smaller tasks (one CTA each) - in an all-vs-all most row chains keep a few dozen partners and a
	// launch of a few large tasks leaves most SMs idle; rows with thousands of survivors get full lists (one ramp per list)
	const uint32_t W = (uint32_t)sw_class_warps(cls) * sw_task_chains(running, sw_class_warps(cls), sw_class_chains(cls));
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
	// column chains per SW task: every warp of the CTA gets one chain before any warp gets a second one, and a row chain with few
	// survivors is cut into more, smaller tasks (one CTA each) - in an all-vs-all most row chains keep a few dozen partners and a
	// launch of a few large tasks leaves most SMs idle; rows with thousands of survivors get full lists (one ramp per list)
	const uint32_t W = (uint32_t)sw_class_warps(cls) * sw_task_chains(running, sw_class_warps(cls), sw_class_chains(cls));
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
