// lddt_kernel.cu - K2: per-alignment LDDT, test statistic and coordinate bookkeeping, one CTA per pair.
//
// Replaces DSSAligner::CalcEvalue (dssaligner.cpp:852-904) up to the test statistic: GetPathCounts,
// GetPosABs (dssaligner.cpp:1282), GetLDDT_mu_fast (lddt.cpp:63-124), PDBChain::GetDist2
// (pdbchain.cpp:320-336).  The P/E/Qual values are double pow() of the float test statistic
// (statsig.cpp:27-50) and are evaluated on the host with libm so the last ulp matches the reference.
//
// Bit-exactness: this file must be compiled with -fmad=false (no FMA contraction) and the default IEEE
// sqrt/div; the per-column preserved/considered counts are integers, so their evaluation order is free;
// the per-column scores are summed sequentially in column order exactly like lddt.cpp:110-123.
#include <float.h>

#include <algorithm>

#include "rsk_internal.cuh"

namespace rsk {

namespace {

constexpr int kLddtThreads = 128;
constexpr uint32_t kLddtScratchCtas = 296;  // CTAs of the global-scratch form (two per SM)
constexpr unsigned kFull = 0xffffffffu;

__global__ void __launch_bounds__(kLddtThreads) lddt_ts_kernel(const LddtArgs a)
{
	extern __shared__ __align__(16) float sm[];
	const uint32_t mc = a.maxcols;
	// the aligned columns' coordinates and scores: shared memory, or - alignments of more than ~8 000 columns (titin-sized
	// chains) - this CTA's slice of a global scratch buffer, which L1/L2 serve
	float *base = a.scratch ? a.scratch + (size_t)blockIdx.x * 7 * mc : sm;
	float *xa = base, *ya = xa + mc, *za = ya + mc, *xb = za + mc, *yb = xb + mc, *zb = yb + mc;
	float *colscore = zb + mc;
	__shared__ uint32_t s_counts[3];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

	for (uint32_t pair = blockIdx.x; pair < a.npairs; pair += gridDim.x) {
		PairRec *rec = a.rec + pair;
		const float score = rec->score;
		const uint32_t plen = rec->path_len;
		__syncthreads();  // smem reuse across pairs
		// CalcEvalue is reached by every pair that went through Align_NoAccel, also when SWFast found nothing (empty path,
		// score 0): with MinFwdScore <= 0 (-verysensitive) the reference then still sets Hi = Lo - 1 (= UINT_MAX - 1),
		// Ids = Gaps = 0, LDDT = 0 and a test statistic (dssaligner.cpp:861-904).  Pairs the Mu filter dropped and long-chain
		// pairs without an x-drop alignment never get there (dssaligner.cpp:819-829, 1395-1417).
		const bool no_aln = plen == 0 && (rec->flags & (RSK_HIT_MU_REJECTED | RSK_HIT_MKF)) != 0;
		if (no_aln || score < a.min_fwd_score) {
			// dssaligner.cpp:861-862: nothing is computed; members keep their ClearAlign values
			if (tid == 0) {
				rec->hi_a = rec->hi_b = rec->ids = rec->gaps = 0xffffffffu;
				rec->lddt = 0.0f;
				rec->ts = -FLT_MAX;
			}
			continue;
		}
		uint32_t ai, bi;
		if (a.cross) {
			const uint32_t arel = pair / a.nB;
			ai = a.a_begin + arel;
			bi = pair - arel * a.nB;
		} else {
			ai = a.pair_a[pair];
			bi = a.pair_b[pair];
		}
		const uint64_t oa = a.offA[ai], ob = a.offB[bi];
		// ---- warp 0: walk the path, gather the coordinates of the aligned (M) columns ----
		if (warp == 0) {
			const uint8_t *path = a.pool + rec->path_off;
			uint32_t nM = 0, nD = 0, nI = 0;
			const uint32_t lo_a = rec->lo_a, lo_b = rec->lo_b;
			for (uint32_t k0 = 0; k0 < plen; k0 += 32) {
				const uint32_t k = k0 + lane;
				const uint8_t c = (k < plen) ? path[k] : 0;
				const unsigned mM = __ballot_sync(kFull, c == 'M');
				const unsigned mD = __ballot_sync(kFull, c == 'D');
				const unsigned mI = __ballot_sync(kFull, c == 'I');
				const unsigned below = (1u << lane) - 1u;
				if (c == 'M') {
					const uint32_t idx = nM + __popc(mM & below);
					const uint32_t pa = lo_a + idx + nD + __popc(mD & below);
					const uint32_t pb = lo_b + idx + nI + __popc(mI & below);
					xa[idx] = a.xA[oa + pa]; ya[idx] = a.yA[oa + pa]; za[idx] = a.zA[oa + pa];
					xb[idx] = a.xB[ob + pb]; yb[idx] = a.yB[ob + pb]; zb[idx] = a.zB[ob + pb];
				}
				nM += __popc(mM); nD += __popc(mD); nI += __popc(mI);
			}
			if (lane == 0) { s_counts[0] = nM; s_counts[1] = nD; s_counts[2] = nI; }
		}
		__syncthreads();
		const uint32_t n = s_counts[0];
		// ---- per-column preserved / considered counts (lddt.cpp:74-108; symmetric, so each thread owns a column) ----
		for (uint32_t ci = tid; ci < n; ci += kLddtThreads) {
			const float x1 = xa[ci], y1 = ya[ci], z1 = za[ci];
			const float x2 = xb[ci], y2 = yb[ci], z2 = zb[ci];
			uint32_t cons = 0, pres = 0;
			for (uint32_t cj = 0; cj < n; ++cj) {
				if (cj == ci)
					continue;
				// GetDist2(pos_lo, pos_hi): the reference always has coli < colj; squares make the sign irrelevant
				const float dxa = x1 - xa[cj], dya = y1 - ya[cj], dza = z1 - za[cj];
				const float d1s = dxa * dxa + dya * dya + dza * dza;
				const float dxb = x2 - xb[cj], dyb = y2 - yb[cj], dzb = z2 - zb[cj];
				const float d2s = dxb * dxb + dyb * dyb + dzb * dzb;
				if (d1s > 225.0f && d2s > 225.0f)
					continue;
				const float diff = fabsf(sqrtf(d1s) - sqrtf(d2s));
				pres += (diff <= 0.5f) + (diff <= 1.0f) + (diff <= 2.0f) + (diff <= 4.0f);
				cons += 4;
			}
			float sc = 0.0f;
			if (cons > 0)
				sc = (float)pres / (float)cons;
			colscore[ci] = sc;
		}
		__syncthreads();
		if (tid == 0) {
			float total = 0.0f;
			for (uint32_t c = 0; c < n; ++c)
				total += colscore[c];
			const float lddt = (n == 0) ? 0.0f : total / (float)n;
			const uint32_t nD = s_counts[1], nI = s_counts[2];
			const float sa = a.selfrevA[ai], sb = a.selfrevB[bi];
			float rev = 0.0f;
			if (sa != FLT_MAX && sb != FLT_MAX)
				rev = (sa + sb) / 2;
			const float L = (float)(a.lenA[ai] + a.lenB[bi]) / 2;
			float ts = 0.13f * lddt;
			ts += (1.7f * score - 2.0f * rev) / (L + 250.0f);
			rec->hi_a = rec->lo_a + n + nD - 1;
			rec->hi_b = rec->lo_b + n + nI - 1;
			rec->ids = n;
			rec->gaps = nD + nI;
			rec->lddt = lddt;
			rec->ts = ts;
			rec->flags |= RSK_HIT_HAS_EVALUE;
		}
	}
}

// The same computation with one WARP per pair (8 pairs per CTA, no block barriers): used when every alignment of the batch
// fits kLddtWarpCols columns, which is every batch of chains up to that length.  The typical alignment of a random pair has
// a handful of columns; a 128-thread CTA per pair then spends its time in barriers.
constexpr int kLddtWarpCols = 512;
constexpr int kLddtWarps = 8;
__global__ void __launch_bounds__(kLddtWarps * 32) lddt_ts_warp_kernel(const LddtArgs a)
{
	extern __shared__ __align__(16) float sm[];
	const uint32_t mc = a.maxcols;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	float *xa = sm + (size_t)warp * 7 * mc, *ya = xa + mc, *za = ya + mc, *xb = za + mc, *yb = xb + mc, *zb = yb + mc;
	float *colscore = zb + mc;
	const uint32_t nwarps = gridDim.x * kLddtWarps;
	for (uint32_t pair = blockIdx.x * kLddtWarps + warp; pair < a.npairs; pair += nwarps) {
		PairRec *rec = a.rec + pair;
		const float score = rec->score;
		const uint32_t plen = rec->path_len;
		const bool no_aln = plen == 0 && (rec->flags & (RSK_HIT_MU_REJECTED | RSK_HIT_MKF)) != 0;  // see lddt_ts_kernel
		__syncwarp();  // smem reuse across pairs
		if (no_aln || score < a.min_fwd_score) {
			if (lane == 0) {
				rec->hi_a = rec->hi_b = rec->ids = rec->gaps = 0xffffffffu;
				rec->lddt = 0.0f;
				rec->ts = -FLT_MAX;
			}
			continue;
		}
		uint32_t ai, bi;
		if (a.cross) {
			const uint32_t arel = pair / a.nB;
			ai = a.a_begin + arel;
			bi = pair - arel * a.nB;
		} else {
			ai = a.pair_a[pair];
			bi = a.pair_b[pair];
		}
		const uint64_t oa = a.offA[ai], ob = a.offB[bi];
		const uint8_t *path = a.pool + rec->path_off;
		uint32_t nM = 0, nD = 0, nI = 0;
		const uint32_t lo_a = rec->lo_a, lo_b = rec->lo_b;
		for (uint32_t k0 = 0; k0 < plen; k0 += 32) {
			const uint32_t k = k0 + lane;
			const uint8_t c = (k < plen) ? path[k] : 0;
			const unsigned mM = __ballot_sync(kFull, c == 'M');
			const unsigned mD = __ballot_sync(kFull, c == 'D');
			const unsigned mI = __ballot_sync(kFull, c == 'I');
			const unsigned below = (1u << lane) - 1u;
			if (c == 'M') {
				const uint32_t idx = nM + __popc(mM & below);
				const uint32_t pa = lo_a + idx + nD + __popc(mD & below);
				const uint32_t pb = lo_b + idx + nI + __popc(mI & below);
				xa[idx] = a.xA[oa + pa]; ya[idx] = a.yA[oa + pa]; za[idx] = a.zA[oa + pa];
				xb[idx] = a.xB[ob + pb]; yb[idx] = a.yB[ob + pb]; zb[idx] = a.zB[ob + pb];
			}
			nM += __popc(mM); nD += __popc(mD); nI += __popc(mI);
		}
		__syncwarp();
		const uint32_t n = nM;
		for (uint32_t ci = lane; ci < n; ci += 32) {
			const float x1 = xa[ci], y1 = ya[ci], z1 = za[ci];
			const float x2 = xb[ci], y2 = yb[ci], z2 = zb[ci];
			uint32_t cons = 0, pres = 0;
			for (uint32_t cj = 0; cj < n; ++cj) {
				if (cj == ci)
					continue;
				const float dxa = x1 - xa[cj], dya = y1 - ya[cj], dza = z1 - za[cj];
				const float d1s = dxa * dxa + dya * dya + dza * dza;
				const float dxb = x2 - xb[cj], dyb = y2 - yb[cj], dzb = z2 - zb[cj];
				const float d2s = dxb * dxb + dyb * dyb + dzb * dzb;
				if (d1s > 225.0f && d2s > 225.0f)
					continue;
				const float diff = fabsf(sqrtf(d1s) - sqrtf(d2s));
				pres += (diff <= 0.5f) + (diff <= 1.0f) + (diff <= 2.0f) + (diff <= 4.0f);
				cons += 4;
			}
			float sc = 0.0f;
			if (cons > 0)
				sc = (float)pres / (float)cons;
			colscore[ci] = sc;
		}
		__syncwarp();
		if (lane == 0) {
			float total = 0.0f;
			for (uint32_t c = 0; c < n; ++c)  // sequential, in column order (lddt.cpp:110-123)
				total += colscore[c];
			const float lddt = (n == 0) ? 0.0f : total / (float)n;
			const float sa = a.selfrevA[ai], sb = a.selfrevB[bi];
			float rev = 0.0f;
			if (sa != FLT_MAX && sb != FLT_MAX)
				rev = (sa + sb) / 2;
			const float L = (float)(a.lenA[ai] + a.lenB[bi]) / 2;
			float ts = 0.13f * lddt;
			ts += (1.7f * score - 2.0f * rev) / (L + 250.0f);
			rec->hi_a = rec->lo_a + n + nD - 1;
			rec->hi_b = rec->lo_b + n + nI - 1;
			rec->ids = n;
			rec->gaps = nD + nI;
			rec->lddt = lddt;
			rec->ts = ts;
			rec->flags |= RSK_HIT_HAS_EVALUE;
		}
	}
}

}  // namespace

int launch_lddt(const LddtArgs &args, cudaStream_t stream)
{
	if (args.npairs == 0)
		return 0;
	if (args.maxcols <= (uint32_t)kLddtWarpCols) {
		const size_t wsmem = (size_t)kLddtWarps * args.maxcols * 7 * sizeof(float);
		if (wsmem > 48 * 1024 &&
			cudaFuncSetAttribute(lddt_ts_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem) != cudaSuccess)
			return -1;
		const unsigned wgrid = (unsigned)std::min<uint64_t>(((uint64_t)args.npairs + kLddtWarps - 1) / kLddtWarps, 148u * 64u);
		lddt_ts_warp_kernel<<<wgrid, kLddtWarps * 32, wsmem, stream>>>(args);
		return cudaGetLastError() == cudaSuccess ? 1 : -1;
	}
	if (args.scratch) {  // lddt_scratch_floats() said the columns do not fit shared memory
		lddt_ts_kernel<<<std::min<uint32_t>(args.npairs, kLddtScratchCtas), kLddtThreads, 0, stream>>>(args);
		return cudaGetLastError() == cudaSuccess ? 1 : -1;
	}
	const size_t smem = (size_t)args.maxcols * 7 * sizeof(float);
	if (smem > 48 * 1024) {
		if (cudaFuncSetAttribute(lddt_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
			return -1;
	}
	const unsigned grid = args.npairs < (1u << 20) ? args.npairs : (1u << 20);
	lddt_ts_kernel<<<grid, kLddtThreads, smem, stream>>>(args);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// floats of global scratch launch_lddt needs for alignments of up to maxcols columns (0: they fit shared memory)
size_t lddt_scratch_floats(uint32_t maxcols)
{
	return (size_t)maxcols * 7 * sizeof(float) > 220 * 1024 ? (size_t)kLddtScratchCtas * 7 * maxcols : 0;
}

}  // namespace rsk
