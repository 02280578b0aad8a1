// global_kernel.cu - K9: global alignment of explicit pairs (-global).
// Replaces ViterbiFastMem (viterbifastmem.cpp:33-193) + TraceBackBitMem (tracebackbitmem.cpp:8-69) under
// DSSAligner::AlignQueryTarget_Global (global.cpp:7-33); the cell score is SubstScore (xdrophsp.cpp:8-33).
//
// Three states over prefix lengths: M[i][j], D[i][j] (A residue alone), I[i][j] (B residue alone).  The reference walks the
// matrix row by row; here one warp owns a pair and sweeps 32 rows at a time as a lane-skewed wavefront (lane = row, lane l
// is l columns behind lane l-1), so that cell (i, j) gets M[i][j+1] and D[i][j] from the lane above by shuffle one step after
// they were produced, and keeps I[i][j] in a register.  Between 32-row passes the last lane parks M and D per column in a
// per-warp boundary array (read 31 columns ahead of where it is rewritten).  Column LB (deletions after the last B residue)
// is one more step of the same wavefront; the row after the last A residue (insertions) is folded into the lane that owns
// row LA-1.  Trace bits: one byte per cell as in the reference, four cells of a row per 32-bit store, in a per-warp scratch
// matrix; lane 0 walks it back and the warp reverses the path in place.
// All comparisons keep the reference's strictness (>, >=) and its finite "minus infinity" -9e9f; fp32 adds in its order.
#include "rsk_internal.cuh"

namespace rsk {
namespace {

constexpr float kNeg = -9e9f;                                            // xdpmem.h:6
constexpr float kOpen = -1.0f, kExt = -0.05f, kTermOpen = 0.0f, kTermExt = 0.0f;  // viterbifastmem.cpp:6-9
constexpr unsigned TB_DM = 1, TB_IM = 2, TB_MD = 4, TB_MI = 8;          // tracebit.h
constexpr int kGlobalWarps = 8;
constexpr unsigned kFull = 0xffffffffu;

// xdrophsp.cpp:8-33: starts from 0, features 0..7 in order.  rowbase[f] = index of the table row of this lane's A letter of
// feature f, minus the feature's e-letter base, so that a B e-letter byte indexes it directly.
__device__ __forceinline__ void row_bases(const uint64_t ea, int (&rowbase)[RSK_NFEAT])
{
#pragma unroll
	for (int f = 0; f < RSK_NFEAT; ++f) {
		const int a = (int)((ea >> (8 * f)) & 0xff) - feat_base(f);
		rowbase[f] = feat_table_off(f) + a * feat_alpha(f) - feat_base(f);
	}
}
__device__ __forceinline__ float cell_score(const float *tab, const int (&rowbase)[RSK_NFEAT], const uint64_t eb)
{
	const unsigned lo = (unsigned)eb, hi = (unsigned)(eb >> 32);
	float t = 0.0f;
#pragma unroll
	for (int f = 0; f < RSK_NFEAT; ++f) {
		const unsigned w = f < 4 ? lo : hi;
		t += tab[rowbase[f] + (int)((w >> (8 * (f & 3))) & 0xffu)];
	}
	return t;
}

__global__ void __launch_bounds__(kGlobalWarps * 32) global_viterbi_kernel(const GlobalArgs a)
{
	__shared__ float s_tab[RSK_TABLE_FLOATS];
	for (int k = threadIdx.x; k < RSK_TABLE_FLOATS; k += blockDim.x)
		s_tab[k] = a.tables[k];
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const size_t gw = (size_t)blockIdx.x * kGlobalWarps + (threadIdx.x >> 5);
	uint8_t *tb = a.tb + gw * a.tb_stride;
	float *bndM = a.bnd + gw * 2 * (size_t)a.bnd_stride;
	float *bndD = bndM + a.bnd_stride;
	for (;;) {
		uint32_t w = 0;
		if (lane == 0)
			w = atomicAdd(a.counter, 1u);
		w = __shfl_sync(kFull, w, 0);
		if (w >= a.npairs)
			break;
		const uint32_t k = a.order[w];
		GlobalRec *rec = a.rec + k;
		if (a.skip && a.skip[k]) {  // the Mu filter said no (global.cpp:15-21): m_GlobalScore stays at ClearAlign's -9999
			if (lane == 0) {
				rec->score = -9999.0f;
				rec->path_len = 0;
			}
			continue;
		}
		const uint32_t ca = a.pair_a[k], cb = a.pair_b[k];
		const int LA = (int)a.lenA[ca], LB = (int)a.lenB[cb];
		const uint64_t *PA = a.profA + a.offA[ca], *PB = a.profB + a.offB[cb];
		const int W = (LB + 1 + 3) & ~3;  // bytes per trace row
		for (int j = lane; j <= LB; j += 32) {
			bndM[j] = kNeg;  // M[0][j+1]
			bndD[j] = kNeg;  // D[0][j]
		}
		__syncwarp();
		float lastI = kNeg;               // I[LA][.] of the row after the last A residue
		float finM = kNeg, finD = kNeg;
		const int npass = (LA + 31) >> 5;
		for (int pass = 0; pass < npass; ++pass) {
			const int i = (pass << 5) + lane;
			const bool row = i < LA;
			int rowbase[RSK_NFEAT];
			row_bases(row ? PA[i] : 0, rowbase);
			uint32_t *trow = reinterpret_cast<uint32_t *>(tb + (size_t)i * W);
			float ins = kNeg;                       // I[i][j]
			float mout = kNeg, dout = kNeg;          // M[i+1][j+1], D[i+1][j] of the cell just computed
			float mprev = (i == 0) ? 0.0f : kNeg;   // M[i][j]; column 0: 0 for the first row only
			uint32_t acc = 0;
			const int nsteps = LB + 1 + 31;
			for (int s = 0; s < nsteps; ++s) {
				const int j = s - lane;
				float recvM = __shfl_up_sync(kFull, mout, 1);
				float recvD = __shfl_up_sync(kFull, dout, 1);
				const bool act = row && j >= 0 && j <= LB;
				if (lane == 0 && act) {
					recvM = bndM[j];
					recvD = bndD[j];
				}
				if (act) {
					const float mhere = mprev;
					unsigned bits = 0;
					if (j < LB) {
						float best = mhere;
						if (recvD > best) { best = recvD; bits = TB_DM; }
						if (ins > best) { best = ins; bits = TB_IM; }
						mout = best + cell_score(s_tab, rowbase, PB[j]);
						const float open = j == 0 ? kTermOpen : kOpen, ext = j == 0 ? kTermExt : kExt;
						const float md = mhere + open;
						float dn = recvD + ext;
						if (md >= dn) { dn = md; bits |= TB_MD; }
						dout = dn;
						const float mi = mhere + open;
						ins += ext;
						if (mi >= ins) { ins = mi; bits |= TB_MI; }
						if (i == LA - 1) {
							// the row after the last A residue (viterbifastmem.cpp:151-167): columns 1..LB-1, strict >
							if (j + 1 < LB) {
								const float t = mout + kTermOpen;
								lastI += kTermExt;
								unsigned char lb = 0;
								if (t > lastI) { lastI = t; lb = (unsigned char)TB_MI; }
								tb[(size_t)LA * W + j + 1] = lb;
							}
							if (j == LB - 1)
								finM = mout;
						}
					} else {
						// the column after the last B residue (:129-143)
						const float md = mhere + kTermOpen;
						float dn = recvD + kTermExt;
						if (md >= dn) { dn = md; bits = TB_MD; }
						dout = dn;
						if (i == LA - 1)
							finD = dout;
					}
					acc |= bits << (8 * (j & 3));
					if ((j & 3) == 3 || j == LB) {
						trow[j >> 2] = acc;
						acc = 0;
					}
					if (lane == 31) {
						if (j < LB) bndM[j] = mout;
						bndD[j] = dout;
					}
					mprev = recvM;  // M[i][j+1]
				}
			}
			__syncwarp();
		}
		// final state (viterbifastmem.cpp:169-188): M, then D, then I, strict >
		const int owner = (LA - 1) & 31;
		finM = __shfl_sync(kFull, finM, owner);
		finD = __shfl_sync(kFull, finD, owner);
		lastI = __shfl_sync(kFull, lastI, owner);
		float score = finM;
		char state = 'M';
		if (finD > score) { score = finD; state = 'D'; }
		if (lastI > score) { score = lastI; state = 'I'; }
		__threadfence_block();
		__syncwarp();
		char *path = a.pool + a.path_off[k];
		uint32_t n = 0;
		if (lane == 0) {
			int i = LA, j = LB;
			while (i != 0 || j != 0) {
				path[n++] = state;
				if (state == 'M') {
					const unsigned t = tb[(size_t)(i - 1) * W + (j - 1)];
					state = (t & TB_DM) ? 'D' : (t & TB_IM) ? 'I' : 'M';
					--i; --j;
				} else if (state == 'D') {
					const unsigned t = tb[(size_t)(i - 1) * W + j];
					state = (t & TB_MD) ? 'M' : 'D';
					--i;
				} else {
					const unsigned t = tb[(size_t)i * W + (j - 1)];
					state = (t & TB_MI) ? 'M' : 'I';
					--j;
				}
			}
			rec->score = score;
			rec->path_len = n;
		}
		n = __shfl_sync(kFull, n, 0);
		__syncwarp();
		for (uint32_t x = lane; x < n / 2; x += 32) {
			const char c = path[x];
			path[x] = path[n - 1 - x];
			path[n - 1 - x] = c;
		}
		__syncwarp();
	}
}

}  // namespace

size_t global_tb_bytes(uint32_t maxLA, uint32_t maxLB) { return ((size_t)maxLA + 1) * (((size_t)maxLB + 1 + 3) & ~(size_t)3); }
int global_warps_per_block() { return kGlobalWarps; }

int launch_global(const GlobalArgs &args, int blocks, cudaStream_t stream)
{
	global_viterbi_kernel<<<blocks, kGlobalWarps * 32, 0, stream>>>(args);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace rsk
