// global_kernel.cu - K9: global alignment of explicit pairs (-global).
// Replaces ViterbiFastMem (viterbifastmem.cpp:33-193) + TraceBackBitMem (tracebackbitmem.cpp:8-69) under
// DSSAligner::AlignQueryTarget_Global (global.cpp:7-33); the cell score is SubstScore (xdrophsp.cpp:8-33).
//
// Three states over prefix lengths: M[i][j], D[i][j] (A residue alone), I[i][j] (B residue alone).  The reference walks the
// matrix row by row; here one warp owns a pair and sweeps 32 x R rows at a time as a lane-skewed wavefront (a lane owns R
// consecutive rows, R = 1..6 chosen per pair so that the chain needs as few passes as possible; lane l is l columns behind
// lane l-1).  Within a step a lane walks its R rows of one column top down; the first row gets M[i][j+1] and D[i][j] from the
// lane above by shuffle one step after they were produced, the others from registers; I[i][j] stays in a register.  Between
// passes the last lane parks M and D per column in a per-warp boundary array (read 31 columns ahead of where it is
// rewritten).  Column LB (deletions after the last B residue) is one more step of the same wavefront; the row after the last
// A residue (insertions) is folded into the lane that owns row LA-1.  Trace bits: one byte per cell as in the reference, four
// cells of a row per 32-bit store, in a per-warp scratch matrix; lane 0 walks it back and the warp reverses the path in place.
// (Round 1 had one row per lane: two shuffles, one 8-letter column fetch and a 31-step ramp per 32 rows.)
// All comparisons keep the reference's strictness (>, >=) and its finite "minus infinity" -9e9f; fp32 adds in its order.
#include "rsk_internal.cuh"

namespace rsk {
namespace {

constexpr float kNeg = -9e9f;                                            // xdpmem.h:6
constexpr float kOpen = -1.0f, kExt = -0.05f, kTermOpen = 0.0f, kTermExt = 0.0f;  // viterbifastmem.cpp:6-9
constexpr unsigned TB_DM = 1, TB_IM = 2, TB_MD = 4, TB_MI = 8;          // tracebit.h
constexpr int kGlobalWarps = 8;
constexpr int kGlobalMaxR = 6;  // rows per lane of the wavefront (register budget: 8 table-row bases per row)
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

// rows per lane and passes for a row chain of LA residues: as few passes as R <= kGlobalMaxR allows, then the smallest R
__host__ __device__ inline void global_geometry(int LA, int &npass, int &R)
{
	npass = (LA + 32 * kGlobalMaxR - 1) / (32 * kGlobalMaxR);
	if (npass < 1) npass = 1;
	R = (LA + 32 * npass - 1) / (32 * npass);
	if (R < 1) R = 1;
}

// The forward sweep of one pair with R rows per lane.  Within a step a lane walks its R rows of column j top down: row r takes
// D[i][j] from the row above (a register, or the shuffle for r = 0) and M[i][j] from the register its upper neighbour filled one
// step earlier, so the two shuffles, the fetch and unpacking of the column's letters and the loop overhead are paid once per R
// cells, and a chain of LA rows needs ceil(LA / 32R) passes (each with a 31-step ramp) instead of ceil(LA / 32).
template <int R>
__device__ __forceinline__ void global_forward(const float *s_tab, const int lane, uint8_t *tb, float *bndM, float *bndD, const int LA,
		const int LB, const uint64_t *PA, const uint64_t *PB, const int npass, float &finM, float &finD, float &lastI)
{
	const int W = (LB + 1 + 3) & ~3;  // bytes per trace row
	lastI = kNeg;
	finM = kNeg;
	finD = kNeg;
	for (int pass = 0; pass < npass; ++pass) {
		const int i0 = (pass * 32 + lane) * R;  // this lane's first row
		const int nrows = min(R, LA - i0);      // <= 0: no row of the chain in this lane
		const int lastr = LA - 1 - i0;          // which of the lane's rows is row LA-1 (if any)
		int rowbase[R][RSK_NFEAT];
		float ins[R], mdiag[R];
		uint32_t acc[R];
#pragma unroll
		for (int r = 0; r < R; ++r) {
			row_bases(r < nrows ? PA[i0 + r] : 0, rowbase[r]);
			ins[r] = kNeg;                              // I[i][j]
			mdiag[r] = (i0 + r == 0) ? 0.0f : kNeg;     // M[i][j]; column 0: 0 for the first row only
			acc[r] = 0;
		}
		float outM = kNeg, outD = kNeg;  // M[i+1][j+1], D[i+1][j] of the lane's last row, for the lane below
		const int nsteps = LB + 1 + 31;
		for (int s = 0; s < nsteps; ++s) {
			const int j = s - lane;
			float recvM = __shfl_up_sync(kFull, outM, 1);
			float recvD = __shfl_up_sync(kFull, outD, 1);
			const bool act = nrows > 0 && j >= 0 && j <= LB;
			if (lane == 0 && act) {
				recvM = bndM[j];
				recvD = bndD[j];
			}
			if (act) {
				const uint64_t eb = j < LB ? PB[j] : 0;
				const float open = j == 0 ? kTermOpen : kOpen, ext = j == 0 ? kTermExt : kExt;
				float dIn = recvD;     // D[i][j] of the row being computed
				float mAbove = recvM;  // M[i][j+1]: the row's diagonal at the next column
				float mout = kNeg, dout = kNeg;
#pragma unroll
				for (int r = 0; r < R; ++r) {
					if (r < nrows) {
						const int i = i0 + r;
						const float mhere = mdiag[r];
						unsigned bits = 0;
						if (j < LB) {
							float best = mhere;
							if (dIn > best) { best = dIn; bits = TB_DM; }
							if (ins[r] > best) { best = ins[r]; bits = TB_IM; }
							mout = best + cell_score(s_tab, rowbase[r], eb);
							const float md = mhere + open;
							float dn = dIn + ext;
							if (md >= dn) { dn = md; bits |= TB_MD; }
							dout = dn;
							const float mi = mhere + open;
							float in = ins[r] + ext;
							if (mi >= in) { in = mi; bits |= TB_MI; }
							ins[r] = in;
							if (r == lastr) {
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
							float dn = dIn + kTermExt;
							if (md >= dn) { dn = md; bits = TB_MD; }
							dout = dn;
							if (r == lastr)
								finD = dout;
						}
						acc[r] |= bits << (8 * (j & 3));
						if ((j & 3) == 3 || j == LB) {
							reinterpret_cast<uint32_t *>(tb + (size_t)i * W)[j >> 2] = acc[r];
							acc[r] = 0;
						}
						mdiag[r] = mAbove;  // M[i][j+1]
						mAbove = mout;      // M[i+1][j+1] is the next row's diagonal at the next column
						dIn = dout;         // D[i+1][j] enters the next row
					}
				}
				outM = mout;
				outD = dout;
				if (lane == 31) {
					if (j < LB) bndM[j] = mout;
					bndD[j] = dout;
				}
			}
		}
		__syncwarp();
	}
	// the lane that owns row LA-1 holds the final values
	const int owner = ((LA - 1) / R) & 31;
	finM = __shfl_sync(kFull, finM, owner);
	finD = __shfl_sync(kFull, finD, owner);
	lastI = __shfl_sync(kFull, lastI, owner);
}

__global__ void __launch_bounds__(kGlobalWarps * 32, 2) global_viterbi_kernel(const GlobalArgs a)
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
		float lastI, finM, finD;  // I[LA][.] of the row after the last A residue; M[LA][LB], D[LA][LB]
		int npass, R;
		global_geometry(LA, npass, R);
		switch (R) {
		case 1: global_forward<1>(s_tab, lane, tb, bndM, bndD, LA, LB, PA, PB, npass, finM, finD, lastI); break;
		case 2: global_forward<2>(s_tab, lane, tb, bndM, bndD, LA, LB, PA, PB, npass, finM, finD, lastI); break;
		case 3: global_forward<3>(s_tab, lane, tb, bndM, bndD, LA, LB, PA, PB, npass, finM, finD, lastI); break;
		case 4: global_forward<4>(s_tab, lane, tb, bndM, bndD, LA, LB, PA, PB, npass, finM, finD, lastI); break;
		case 5: global_forward<5>(s_tab, lane, tb, bndM, bndD, LA, LB, PA, PB, npass, finM, finD, lastI); break;
		default: global_forward<6>(s_tab, lane, tb, bndM, bndD, LA, LB, PA, PB, npass, finM, finD, lastI); break;
		}
		// final state (viterbifastmem.cpp:169-188): M, then D, then I, strict >
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
