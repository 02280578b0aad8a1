// sw_kernel.cu - K1: float affine-gap local Smith-Waterman with traceback over the on-the-fly 8-feature
// DSS score, one warp per (A,B) pair, one CTA per A chain (16 pairs at a time).
//
// Replaces, per pair: DSSAligner::SetSMx_NoRev (dssaligner.cpp:529-596) + SWFast (sw.cpp:79-212) +
// TraceBackBitSW (sw.cpp:8-77).  Results are bit-identical: every cell performs exactly the reference's
// fp32 adds and compares (sw.cpp:119-197), only the order in which independent cells are visited differs.
//
// Design (B200, sm_100a):
//  * The score matrix S is never materialised.  For the CTA's A chain a "row table" is staged in shared
//    memory: T[e][h][lane][q] = tab_f[a_f(row)][b] for every (feature,letter) code e = 0..131 and every row
//    of the current pass, laid out so that lane L's 16-byte vector lives in bank group L%8: one conflict-free
//    LDS.128 returns a feature's score for 4 consecutive rows.  The B chain supplies one 8-byte column code
//    per DP column (pre-offset "e-letters"), so a cell's score is 8 table reads + 7 fp32 adds in the
//    reference's summation order.
//  * DP wavefront: lane L owns R consecutive rows (R = 1..8), lanes are skewed by one column per step, the
//    last row's (M,D) travels to the next lane by warp shuffle.  A chain longer than 32*R rows is cut into
//    passes; the bottom row of a pass is parked in a small global (L2-resident) boundary buffer.
//  * Traceback bits: 4 bits per cell (2 bits match-source, MD, MI), 8 rows -> one 32-bit word per lane and
//    column, 4 columns -> one 16-byte store.  The buffer is per-warp scratch that is re-used pair after pair,
//    so it stays in the 126 MB L2; the in-kernel traceback reads it back through a 2 KB shared-memory tile.
//  * Persistent CTAs (one per SM) pull tasks from an atomic counter.
#include "rsk_internal.cuh"

namespace rsk {

namespace {

constexpr int kNLet = RSK_NLETTERS;
constexpr unsigned kFull = 0xffffffffu;

// shared memory carve-up
constexpr size_t kSmemTab = 0;                                        // float[2192]
constexpr size_t kSmemRowTab = 8768;                                  // float4[132*2*32]
constexpr size_t kSmemTiles = kSmemRowTab + (size_t)kNLet * 2 * 32 * 16;  // uint4[kSwWarps][4][32]
constexpr size_t kSmemBcast = kSmemTiles + (size_t)kSwWarps * 4 * 32 * 16;
constexpr size_t kSmemTotal = kSmemBcast + 16;

template <int R>
__device__ __forceinline__ void build_rowtab(float *rt, const float *tab, const uint64_t *__restrict__ profA,
		uint32_t LA, int pass)
{
	constexpr int NH = (R + 3) / 4;
	constexpr int ROWS = 32 * R;
	const uint32_t rowbase = (uint32_t)pass * ROWS;
	for (int idx = threadIdx.x; idx < kNLet * ROWS; idx += kSwThreads) {
		const int e = idx / ROWS;
		const int rr = idx - e * ROWS;
		const int l = rr / R;
		const int r = rr - l * R;
		const uint32_t row = rowbase + rr;
		float v = 0.0f;
		if (row < LA) {
			const int f = e < 20 ? 0 : 1 + ((e - 20) >> 4);
			const int b = e - feat_base(f);
			const uint64_t pa = __ldg(profA + row);
			const int a = (int)((pa >> (8 * f)) & 0xff) - feat_base(f);
			v = tab[feat_table_off(f) + a * feat_alpha(f) + b];
		}
		rt[((e * NH + (r >> 2)) * 32 + l) * 4 + (r & 3)] = v;
	}
}

__device__ __forceinline__ float pick4(const float4 &v, int q)
{
	return q == 0 ? v.x : q == 1 ? v.y : q == 2 ? v.z : v.w;
}

// One pass: rows [pass*32R, (pass+1)*32R) of A against all LB columns of B.
template <int R>
__device__ __forceinline__ void sw_pass(const float4 *__restrict__ rowtab, const int lane, const int pass,
		const int npass, const uint32_t LA, const uint64_t *__restrict__ colB, const int LB, const int LBpad,
		float2 *__restrict__ bnd, uint4 *__restrict__ trace_pass, const float open, const float ext,
		float &lbest, int &lbi, int &lbj)
{
	constexpr int NH = (R + 3) / 4;
	float Mrow[R], Irow[R], best[R];
	int bestj[R];
#pragma unroll
	for (int r = 0; r < R; ++r) {
		Mrow[r] = kNegInf;  // M[i_r+1][0]
		Irow[r] = kNegInf;  // I[i_r][0]
		best[r] = 0.0f;
		bestj[r] = 0;
	}
	const bool first = (pass == 0), last = (pass == npass - 1);
	float mdiag_next = (lane == 0 && first) ? 0.0f : kNegInf;  // M[i0][0]; M[0][0] = 0 (sw.cpp:116)
	float outM = kNegInf, outD = kNegInf;
	uint32_t t0 = 0, t1 = 0, t2 = 0, t3 = 0;
	const int nsteps = LBpad + 31;
	int j = -lane;
	uint64_t cb = (j >= 0 && j < LB) ? __ldg(colB + j) : 0ull;
	float2 bn = make_float2(kNegInf, kNegInf);
	if (lane == 0 && !first)
		bn = bnd[0];
	for (int s = 0; s < nsteps; ++s, ++j) {
		const float inM = __shfl_up_sync(kFull, outM, 1);
		const float inD = __shfl_up_sync(kFull, outD, 1);
		const int jn = j + 1;
		const uint64_t cb_next = (jn >= 0 && jn < LB) ? __ldg(colB + jn) : 0ull;
		float2 bn_next = bn;
		if (lane == 0 && !first && jn < LB)
			bn_next = bnd[jn];
		uint32_t tw = 0;
		if (j >= 0 && j < LB) {
			float d, mdiag = mdiag_next;
			if (lane == 0) {
				d = first ? kNegInf : bn.y;           // D[i0][j]
				mdiag_next = first ? kNegInf : bn.x;  // M[i0][j+1]
			} else {
				d = inD;
				mdiag_next = inM;
			}
			float S[R];
#pragma unroll
			for (int f = 0; f < RSK_NFEAT; ++f) {
				const uint32_t e = (uint32_t)(cb >> (8 * f)) & 0xffu;
				const float4 *p = rowtab + (e * NH) * 32 + lane;
				const float4 v0 = p[0];
				float4 v1 = v0;
				if (NH == 2)
					v1 = p[32];
#pragma unroll
				for (int r = 0; r < R; ++r) {
					const float v = (r < 4) ? pick4(v0, r & 3) : pick4(v1, r & 3);
					S[r] = (f == 0) ? v : S[r] + v;  // feature 0 assigns, 1..7 accumulate (dssaligner.cpp:557-595)
				}
			}
#pragma unroll
			for (int r = 0; r < R; ++r) {
				const float m = mdiag;  // M[i][j]
				mdiag = Mrow[r];        // becomes M[i+1][j] for the next row
				const float ii = Irow[r];
				float x = m;
				uint32_t code = 0;
				if (d > x) { x = d; code = 1; }
				if (ii > x) { x = ii; code = 2; }
				if (0.0f >= x) { x = 0.0f; code = 3; }
				x += S[r];
				if (x > best[r]) { best[r] = x; bestj[r] = j; }
				Mrow[r] = x;  // M[i+1][j+1]
				const float mo = m + open;
				float dn = d + ext;
				if (mo >= dn) { dn = mo; code |= 4u; }
				d = dn;  // D[i+1][j]
				float in2 = ii + ext;
				if (mo >= in2) { in2 = mo; code |= 8u; }
				Irow[r] = in2;  // I[i][j+1]
				tw |= code << (4 * r);
			}
			outM = Mrow[R - 1];
			outD = d;
			if (lane == 31 && !last)
				bnd[j] = make_float2(outM, outD);
		}
		t0 = t1; t1 = t2; t2 = t3; t3 = tw;
		if (j >= 0 && j < LBpad && (j & 3) == 3)
			trace_pass[(j >> 2) * 32 + lane] = make_uint4(t0, t1, t2, t3);
		cb = cb_next;
		bn = bn_next;
	}
	const uint32_t row0 = (uint32_t)pass * 32 * R + (uint32_t)lane * R;
#pragma unroll
	for (int r = 0; r < R; ++r) {
		if (row0 + r < LA && best[r] > lbest) {
			lbest = best[r];
			lbi = (int)(row0 + r);
			lbj = bestj[r];
		}
	}
}

// Warp-cooperative traceback (sw.cpp:8-77) through a 2 KB shared tile of the packed trace.
__device__ __forceinline__ void traceback_and_emit(const SwArgs &a, const int lane, const int R, const int LBpad,
		const uint4 *__restrict__ trace, uint4 *tile, uint8_t *stage, const float score, const int bi, const int bj,
		PairRec *rec)
{
	const int rows_per_pass = 32 * R;
	const int nblk = LBpad >> 2;
	const uint32_t *tile32 = reinterpret_cast<const uint32_t *>(tile);
	int i = bi + 1, j = bj + 1;
	int state = 0;  // 0 = M, 1 = D, 2 = I
	uint32_t n = 0;
	int cur_p = -1, cur_g = -1;
	for (;;) {
		if (lane == 0)
			stage[n] = (uint8_t)(state == 0 ? 'M' : state == 1 ? 'D' : 'I');
		++n;
		const int ci = (state == 2) ? i : i - 1;
		const int cj = (state == 1) ? j : j - 1;
		const int p = ci / rows_per_pass;
		const int rr = ci - p * rows_per_pass;
		const int g = cj >> 4;
		if (p != cur_p || g != cur_g) {
			__syncwarp();
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				const int jb = 4 * g + k;
				if (jb < nblk)
					tile[k * 32 + lane] = trace[((size_t)p * nblk + jb) * 32 + lane];
			}
			__syncwarp();
			cur_p = p;
			cur_g = g;
		}
		const int srcl = rr / R;
		const int r = rr - srcl * R;
		const uint32_t w = tile32[((((cj >> 2) & 3) * 32 + srcl) << 2) + (cj & 3)];
		const uint32_t nib = (w >> (4 * r)) & 15u;
		if (state == 0) {
			--i; --j;
			const uint32_t src = nib & 3u;
			if (src == 3u)
				break;
			state = (int)src;  // 0 M, 1 D, 2 I
		} else if (state == 1) {
			--i;
			state = (nib & 4u) ? 0 : 1;
		} else {
			--j;
			state = (nib & 8u) ? 0 : 2;
		}
	}
	// publish: reserve pool space, copy the staged (reversed) path forward
	unsigned long long off = 0;
	if (lane == 0)
		off = atomicAdd(a.pool_cursor, (unsigned long long)n);
	off = __shfl_sync(kFull, off, 0);
	__syncwarp();
	for (uint32_t k = lane; k < n; k += 32)
		a.pool[off + k] = stage[n - 1 - k];
	if (lane == 0) {
		rec->score = score;
		rec->lo_a = (uint32_t)i;
		rec->lo_b = (uint32_t)j;
		rec->path_len = n;
		rec->path_off = off;
	}
}

template <int R>
__device__ __forceinline__ void process_task(const SwArgs &a, unsigned char *smem, const uint32_t ai,
		const uint32_t begin, const uint32_t cnt, const uint32_t slot_base)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const float *tab = reinterpret_cast<const float *>(smem + kSmemTab);
	float4 *rowtab = reinterpret_cast<float4 *>(smem + kSmemRowTab);
	uint4 *tile = reinterpret_cast<uint4 *>(smem + kSmemTiles) + warp * 128;

	const uint32_t LA = a.lenA[ai];
	const uint64_t *profA = a.profA + a.offA[ai];
	const int npass = (int)((LA + 32 * R - 1) / (32 * R));

	const bool have = (uint32_t)warp < cnt;
	uint32_t bidx = 0, slot = 0;
	int LB = 0, LBpad = 0;
	const uint64_t *colB = nullptr;
	if (have) {
		bidx = a.blist[begin + warp];
		slot = a.cross ? slot_base + bidx : a.bslot[begin + warp];
		LB = (int)a.lenB[bidx];
		LBpad = (LB + 3) & ~3;
		colB = a.profB + a.offB[bidx];
	}
	const size_t gw = (size_t)blockIdx.x * kSwWarps + warp;
	uint4 *trace = a.trace + gw * a.trace_stride;
	float2 *bnd = a.bnd + gw * a.bnd_stride;
	uint8_t *stage = a.stage + gw * a.stage_stride;

	float lbest = 0.0f;
	int lbi = 0x7fffffff, lbj = 0;
	for (int pass = 0; pass < npass; ++pass) {
		__syncthreads();  // every warp is done with the previous pass's row table
		build_rowtab<R>(reinterpret_cast<float *>(rowtab), tab, profA, LA, pass);
		__syncthreads();
		if (have)
			sw_pass<R>(rowtab, lane, pass, npass, LA, colB, LB, LBpad, bnd,
					trace + (size_t)pass * (LBpad >> 2) * 32, a.open, a.ext, lbest, lbi, lbj);
	}
	if (!have)
		return;
	// first maximum in row-major order: max score, ties -> smallest row (a row lives in exactly one lane)
#pragma unroll
	for (int o = 16; o >= 1; o >>= 1) {
		const float os = __shfl_xor_sync(kFull, lbest, o);
		const int oi = __shfl_xor_sync(kFull, lbi, o);
		const int oj = __shfl_xor_sync(kFull, lbj, o);
		if (os > lbest || (os == lbest && oi < lbi)) {
			lbest = os; lbi = oi; lbj = oj;
		}
	}
	PairRec *rec = a.rec + slot;
	if (lbest == 0.0f) {  // sw.cpp:200-201: no positive cell -> score 0, empty path
		if (lane == 0) {
			rec->score = 0.0f;
			rec->lo_a = 0xffffffffu;
			rec->lo_b = 0xffffffffu;
			rec->path_len = 0;
			rec->path_off = 0;
		}
		return;
	}
	__syncwarp();  // trace words written by other lanes of this warp are visible
	traceback_and_emit(a, lane, R, LBpad, trace, tile, stage, lbest, lbi, lbj, rec);
}

__global__ void __launch_bounds__(kSwThreads, 1) sw_affine_f32_tb_kernel(const SwArgs a)
{
	extern __shared__ __align__(16) unsigned char smem[];
	float *tab = reinterpret_cast<float *>(smem + kSmemTab);
	volatile int *bcast = reinterpret_cast<volatile int *>(smem + kSmemBcast);
	for (int k = threadIdx.x; k < RSK_TABLE_FLOATS; k += kSwThreads)
		tab[k] = a.tables[k];
	__syncthreads();
	const uint32_t ntasks = a.ntasks_dev ? *a.ntasks_dev : a.ntasks;
	for (;;) {
		if (threadIdx.x == 0)
			bcast[0] = (int)atomicAdd(a.task_counter, 1u);
		__syncthreads();
		const uint32_t task = (uint32_t)bcast[0];
		__syncthreads();
		if (task >= ntasks)
			break;
		uint32_t ai, begin, cnt, slot_base = 0;
		if (a.cross) {
			const uint32_t arel = task / a.nseg;
			const uint32_t seg = task - arel * a.nseg;
			ai = a.a_begin + arel;
			begin = seg * kSwWarps;
			cnt = min((uint32_t)kSwWarps, a.nB - begin);
			slot_base = arel * a.nB;
		} else {
			ai = a.task_a[task];
			begin = a.task_begin[task];
			cnt = a.task_cnt[task];
		}
		int npass, R;
		sw_geometry(a.lenA[ai], npass, R);
		switch (R) {
		case 1: process_task<1>(a, smem, ai, begin, cnt, slot_base); break;
		case 2: process_task<2>(a, smem, ai, begin, cnt, slot_base); break;
		case 3: process_task<3>(a, smem, ai, begin, cnt, slot_base); break;
		case 4: process_task<4>(a, smem, ai, begin, cnt, slot_base); break;
		case 5: process_task<5>(a, smem, ai, begin, cnt, slot_base); break;
		case 6: process_task<6>(a, smem, ai, begin, cnt, slot_base); break;
		case 7: process_task<7>(a, smem, ai, begin, cnt, slot_base); break;
		default: process_task<8>(a, smem, ai, begin, cnt, slot_base); break;
		}
	}
}

// Plane-major feature letters -> 8 pre-offset e-letter bytes per residue.
__global__ void pack_profiles_kernel(const uint8_t *__restrict__ planes, uint64_t total, uint64_t *__restrict__ prof8)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= total)
		return;
	uint64_t v = 0;
#pragma unroll
	for (int f = 0; f < RSK_NFEAT; ++f) {
		uint32_t b = planes[(uint64_t)f * total + i];
		const uint32_t al = (uint32_t)feat_alpha(f);
		if (b >= al)
			b = al - 1;  // out-of-alphabet letters cannot be produced by the 8 default features (dss.cpp:731)
		v |= (uint64_t)(b + feat_base(f)) << (8 * f);
	}
	prof8[i] = v;
}

}  // namespace

size_t sw_smem_bytes() { return kSmemTotal; }

int launch_sw(const SwArgs &args, int grid, size_t smem, cudaStream_t stream)
{
	static bool attr_set = false;
	if (!attr_set) {
		if (cudaFuncSetAttribute(sw_affine_f32_tb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTotal) != cudaSuccess)
			return -1;
		attr_set = true;
	}
	sw_affine_f32_tb_kernel<<<grid, kSwThreads, smem, stream>>>(args);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_pack_profiles(const uint8_t *planes, uint64_t total, uint64_t *prof8, cudaStream_t stream)
{
	if (total == 0)
		return 0;
	const int threads = 256;
	const unsigned blocks = (unsigned)((total + threads - 1) / threads);
	pack_profiles_kernel<<<blocks, threads, 0, stream>>>(planes, total, prof8);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace rsk
