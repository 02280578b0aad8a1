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
	const uint32_t W = (uint32_t)sw_class_warps(cls);
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

}  // namespace

size_t mu_smem_bytes() { return kMuSmemTotal; }

int launch_mu_filter(const MuArgs &args, int grid, cudaStream_t stream)
{
	if (cudaFuncSetAttribute(mu_sw_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMuSmemTotal) != cudaSuccess)
		return -1;
	mu_sw_filter_kernel<<<grid, kSwThreads, kMuSmemTotal, stream>>>(args);
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
