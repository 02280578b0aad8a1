// sw_kernel.cu - K1: float affine-gap local Smith-Waterman with traceback over the on-the-fly 8-feature
// DSS score, one warp per (A,B) pair, one CTA per A chain (16 pairs at a time).
//
// Replaces, per pair: DSSAligner::SetSMx_NoRev (dssaligner.cpp:529-596) + SWFast (sw.cpp:79-212) +
// TraceBackBitSW (sw.cpp:8-77).  Results are bit-identical: every cell performs exactly the reference's
// fp32 adds and compares (sw.cpp:119-197), only the order in which independent cells are visited differs.
//
// Design (B200, sm_100a):
//  * The score matrix S is never materialised.  For the CTA's A chain a "row table" is staged in shared
//    memory for the current pass of 32*R rows: plane 0 holds rows 0..3 of every lane as float4
//    P0[e][lane], plane 1 rows 4..R-1 as float/float2/float4 P1[e][lane], e = (feature,letter) code 0..131,
//    both with a 512-byte stride per code.  A lane's vector sits in its own bank group, so every LDS is
//    conflict-free.  The B chain supplies, per DP column, 8 ready-made table offsets (32 bytes, prepared
//    once per chain set), so a cell's score costs 8 table reads + 7 fp32 adds in the reference's order.
//  * DP wavefront: lane L owns R consecutive rows (R = 1..8), lanes are skewed by one column per step, the
//    last row's (M,D) travels to the next lane by warp shuffle.  A chain longer than 32*R rows is cut into
//    passes; the bottom row of a pass is parked in a small global (L2-resident) boundary buffer.
//    Steps whose 32 lanes are all inside the matrix run a predicate-free body (the steady state).
//  * Traceback bits: 4 bits per cell (2 bits match-source, MD, MI) -> one 32-bit word per lane and step,
//    four steps -> one fully coalesced 16-byte store per lane ("step-major", i.e. anti-diagonal, layout).
//    The in-kernel traceback reads it back through a 2 KB shared-memory tile.
//  * The running maximum is tracked per lane with one FMNMX3 chain per step; the exact first-maximum rule of
//    the reference (row-major order, strict >) is restored in a rarely taken slow path and in the final
//    (score desc, row asc) warp reduction.
//  * Either chain of a pair can supply the rows (template TR): the scheduler puts the side with fewer chains on
//    the rows so that a CTA's row table is shared by a full set of warps.
//  * Persistent CTAs (one per SM) pull tasks from an atomic counter; one kernel per row-length class so that
//    short row blocks (few registers) run with 24 warps per SM and long ones with 16.
#include <type_traits>

#include "rsk_internal.cuh"

namespace rsk {

namespace {

constexpr int kNLet = RSK_NLETTERS;
constexpr unsigned kFull = 0xffffffffu;

// shared memory carve-up
constexpr size_t kSmemTab = 0;                          // float[2192] weighted tables
constexpr size_t kSmemP0 = 8768;                        // plane 0: float4[132][32], rows 0..3 of each lane
constexpr size_t kPlaneBytes = (size_t)kNLet * 512;     // 67584
constexpr size_t kSmemP1 = kSmemP0 + kPlaneBytes;       // plane 1: rows 4..R-1, same 512-byte stride per code
constexpr size_t kSmemTiles = kSmemP1 + kPlaneBytes;    // uint4[warps][4][32] traceback tiles
constexpr size_t kSmemBcast = kSmemTiles + (size_t)kSwMaxWarps * 4 * 32 * 16;
constexpr size_t kSmemTotal = kSmemBcast + 16;

// floats per lane in plane 1 for R rows per lane (R-4 rounded up to a vector width)
__host__ __device__ constexpr int plane1_width(int R) { return R <= 4 ? 0 : R == 5 ? 1 : R == 6 ? 2 : 4; }

// table value of rows beyond the end of the chain: such cells are hugely negative and can never be a maximum
constexpr float kPadScore = -1e30f;

template <int R, int NTHREADS>
__device__ __forceinline__ void build_rowtab(float *p0, float *p1, const float *tab, const uint64_t *__restrict__ profA,
		uint32_t LA, int pass)
{
	constexpr int W1 = plane1_width(R);
	constexpr int ROWS = 32 * R;
	const uint32_t rowbase = (uint32_t)pass * ROWS;
	for (int idx = threadIdx.x; idx < kNLet * ROWS; idx += NTHREADS) {
		const int e = idx / ROWS;
		const int rr = idx - e * ROWS;
		const int l = rr / R;
		const int r = rr - l * R;
		const uint32_t row = rowbase + rr;
		float v = kPadScore;
		if (row < LA) {
			const int f = e < 20 ? 0 : 1 + ((e - 20) >> 4);
			const int b = e - feat_base(f);
			const uint64_t pa = __ldg(profA + row);
			const int a = (int)((pa >> (8 * f)) & 0xff) - feat_base(f);
			v = tab[feat_table_off(f) + a * feat_alpha(f) + b];
		}
		if (r < 4)
			p0[e * 128 + l * 4 + r] = v;
		else
			p1[e * 128 + l * W1 + (r - 4)] = v;
	}
}

// Packed fp32 add (FADD2 on sm_100a): two independent IEEE round-to-nearest adds in one issue slot.
__device__ __forceinline__ float2 add2(const float2 a, const float2 b)
{
	unsigned long long ra, rb, rd;
	asm("mov.b64 %0, {%1,%2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
	asm("mov.b64 %0, {%1,%2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
	float2 d;
	asm("mov.b64 {%0,%1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
	return d;
}

// out[r] = in[r] + c for r < N, two rows per instruction
template <int N>
__device__ __forceinline__ void add_const(float (&out)[N], const float (&in)[N], const float c)
{
#pragma unroll
	for (int r = 0; r + 1 < N; r += 2) {
		const float2 t = add2(make_float2(in[r], in[r + 1]), make_float2(c, c));
		out[r] = t.x;
		out[r + 1] = t.y;
	}
	if (N & 1)
		out[N - 1] = in[N - 1] + c;
}

// acc[r] += v[r] for r < N, two rows per instruction
template <int N>
__device__ __forceinline__ void add_vec(float (&acc)[N], const float (&v)[8])
{
#pragma unroll
	for (int r = 0; r + 1 < N; r += 2) {
		const float2 t = add2(make_float2(acc[r], acc[r + 1]), make_float2(v[r], v[r + 1]));
		acc[r] = t.x;
		acc[r + 1] = t.y;
	}
	if (N & 1)
		acc[N - 1] = acc[N - 1] + v[N - 1];
}

// One pass: rows [pass*32R, (pass+1)*32R) of A against all LB columns of B.
// colB: two uint4 per column = the column's 8 table offsets in units of 16 bytes (code * 32).
// TR = false: kernel rows are the reference's A chain (index i), columns its B chain (j).
// TR = true : rows are the reference's B chain, columns its A chain.  The recurrence is the same with the
//             roles of the two gap states exchanged: the reference tests D (gap that consumes A) before I
//             (sw.cpp:136-147), and its first-maximum rule prefers the smaller i, then the smaller j.
// lbi/lbj are in kernel coordinates (row, column).
template <int R, bool TR>
__device__ __forceinline__ void sw_pass(const unsigned char *smem_p0, const int lane, const int pass, const int npass,
		const uint32_t LA, const uint4 *__restrict__ colB, const int LB, float2 *__restrict__ bnd,
		uint4 *__restrict__ trace_pass, const float open, const float ext, float &lbest, int &lbi, int &lbj)
{
	constexpr int W1 = plane1_width(R);
	// Per-lane byte offsets into the two planes.  They are made opaque to the optimiser so that "offset*16 + lane
	// term" stays one IMAD; the loads themselves are ordinary shared-memory loads (LDS with the plane base folded into
	// the immediate) which the scheduler may hoist above the warp shuffles of the next step.
	uint32_t base0 = (uint32_t)lane * 16u;
	uint32_t base1 = (uint32_t)lane * (uint32_t)(W1 * 4);
	asm volatile("" : "+r"(base0));
	asm volatile("" : "+r"(base1));
	const unsigned char *const plane0 = smem_p0;
	const unsigned char *const plane1 = smem_p0 + kPlaneBytes;
	float Mrow[R], Irow[R];
#pragma unroll
	for (int r = 0; r < R; ++r) {
		Mrow[r] = kNegInf;  // M[i_r+1][0]
		Irow[r] = kNegInf;  // I[i_r][0]
	}
	const bool first = (pass == 0), last = (pass == npass - 1);
	const bool lane0 = (lane == 0);
	const bool use_bnd = lane0 && !first;
	const bool put_bnd = (lane == 31) && !last;
	float mdiag_next = (lane0 && first) ? 0.0f : kNegInf;  // M[i0][0]; M[0][0] = 0 (sw.cpp:116)
	float outM = kNegInf, outD = kNegInf;
	const uint32_t row0 = (uint32_t)pass * 32 * R + (uint32_t)lane * R;
	const int nsteps = LB + 31;
	const int ngroups = (nsteps + 3) >> 2;
	int j = -lane;
	uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0;
	if (j >= 0 && j < LB) {
		c0 = __ldg(colB + 2 * j);
		c1 = __ldg(colB + 2 * j + 1);
	}
	float2 bn = make_float2(kNegInf, kNegInf);
	if (use_bnd)
		bn = bnd[0];

	// CHECK = false: every lane is inside the matrix at this step and at the next one (no range predicates)
	auto step = [&](auto check_tag) -> uint32_t {
		constexpr bool CHECK = decltype(check_tag)::value;
		const float inM = __shfl_up_sync(kFull, outM, 1);
		const float inD = __shfl_up_sync(kFull, outD, 1);
		const int jn = j + 1;
		uint4 n0 = c0, n1 = c1;
		if (!CHECK || (jn >= 0 && jn < LB)) {
			n0 = __ldg(colB + 2 * jn);
			n1 = __ldg(colB + 2 * jn + 1);
		}
		float2 bn_next = bn;
		if (use_bnd && (!CHECK || jn < LB))
			bn_next = bnd[jn];
		uint32_t tw = 0;
		if (!CHECK || (j >= 0 && j < LB)) {
			float d = lane0 ? bn.y : inD;  // D[i0][j]   (bn = -inf pair in the first pass)
			const float mdiag = mdiag_next;  // M[i0][j]
			mdiag_next = lane0 ? bn.x : inM;  // M[i0][j+1]
			const uint32_t co[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
			float S[R];
#pragma unroll
			for (int f = 0; f < RSK_NFEAT; ++f) {
				const uint32_t a0 = co[f] * 16u + base0;
				const float4 v0 = *reinterpret_cast<const float4 *>(plane0 + a0);
				float v[8];
				v[0] = v0.x; v[1] = v0.y; v[2] = v0.z; v[3] = v0.w;
				v[4] = v[5] = v[6] = v[7] = 0.0f;
				if (W1 == 1) {
					v[4] = *reinterpret_cast<const float *>(plane1 + (co[f] * 16u + base1));
				} else if (W1 == 2) {
					const float2 t = *reinterpret_cast<const float2 *>(plane1 + (co[f] * 16u + base1));
					v[4] = t.x; v[5] = t.y;
				} else if (W1 == 4) {
					const float4 t = *reinterpret_cast<const float4 *>(plane1 + a0);
					v[4] = t.x; v[5] = t.y; v[6] = t.z; v[7] = t.w;
				}
				// feature 0 assigns, 1..7 accumulate in this order (dssaligner.cpp:557-595)
				if (f == 0) {
#pragma unroll
					for (int r = 0; r < R; ++r)
						S[r] = v[r];
				} else {
					add_vec<R>(S, v);
				}
			}
			// the adds that do not depend on this column's gap chain, two rows at a time
			float Mdg[R], Mo[R], Ie[R];  // M[i][j] of each row, M[i][j] + open, horizontal gap state + ext
			Mdg[0] = mdiag;
#pragma unroll
			for (int r = 1; r < R; ++r)
				Mdg[r] = Mrow[r - 1];
			Mo[0] = mdiag + open;
			if (R > 1) {
				float prev[R > 1 ? R - 1 : 1], po[R > 1 ? R - 1 : 1];
#pragma unroll
				for (int r = 0; r + 1 < R; ++r)
					prev[r] = Mrow[r];
				add_const<(R > 1 ? R - 1 : 1)>(po, prev, open);
#pragma unroll
				for (int r = 1; r < R; ++r)
					Mo[r] = po[r - 1];
			}
			add_const<R>(Ie, Irow, ext);
			float xmax = kNegInf;
#pragma unroll
			for (int r = 0; r < R; ++r) {
				const float m = Mdg[r];  // M[i][j]
				const float ii = Irow[r];
				float x = m;
				uint32_t code = 0;
				// code: 1 = reference D state (consumes A), 2 = reference I state; D is tested first
				if (!TR) {
					if (d > x) { x = d; code = 1; }
					if (ii > x) { x = ii; code = 2; }
				} else {
					if (ii > x) { x = ii; code = 1; }
					if (d > x) { x = d; code = 2; }
				}
				if (0.0f >= x) { x = 0.0f; code = 3; }
				x += S[r];
				xmax = fmaxf(xmax, x);
				Mrow[r] = x;  // M[row+1][col+1]
				const float mo = Mo[r];
				float dn = d + ext;
				if (mo >= dn) { dn = mo; code |= TR ? 8u : 4u; }
				d = dn;  // vertical gap state entering the next row
				float in2 = Ie[r];
				if (mo >= in2) { in2 = mo; code |= TR ? 4u : 8u; }
				Irow[r] = in2;  // horizontal gap state entering the next column
				tw |= code << (4 * r);
			}
			outM = Mrow[R - 1];
			outD = d;
			if (put_bnd)
				bnd[j] = make_float2(outM, outD);
			if (TR ? (xmax > lbest) : (xmax >= lbest)) {
				// rare: a cell reached the lane's running maximum.  Keep the reference's first-maximum rule
				// (row-major over (i,j), strict >).  !TR: higher score wins; equal score -> smaller row i; same row ->
				// the earlier column (already stored).  TR: columns are i and are visited in ascending order, rows (j)
				// ascending within a step, so the first cell seen with a score is already the right one: strict > only.
#pragma unroll
				for (int r = 0; r < R; ++r) {
					const float x = Mrow[r];
					const int row = (int)(row0 + r);
					const bool better = TR ? (x > lbest) : (x > lbest || (x == lbest && row < lbi));
					if (better && (uint32_t)row < LA) {
						lbest = x;
						lbi = row;
						lbj = j;
					}
				}
			}
		}
		j = jn;
		c0 = n0;
		c1 = n1;
		bn = bn_next;
		return tw;
	};

	for (int g = 0; g < ngroups; ++g) {
		const int s0 = 4 * g;
		uint32_t t0, t1, t2, t3;
		if (s0 >= 31 && s0 + 4 < LB) {
			t0 = step(std::false_type{});
			t1 = step(std::false_type{});
			t2 = step(std::false_type{});
			t3 = step(std::false_type{});
		} else {
			t0 = step(std::true_type{});
			t1 = step(std::true_type{});
			t2 = step(std::true_type{});
			t3 = step(std::true_type{});
		}
		trace_pass[g * 32 + lane] = make_uint4(t0, t1, t2, t3);
	}
}

// Warp-cooperative traceback (sw.cpp:8-77) through a 2 KB shared tile of the packed trace.
// Trace word of cell (row, col): pass p = row / (32R), lane l = (row % 32R) / R, nibble r = row % R,
// step s = col + l, group g = s / 4, word s % 4 of uint4 trace[(p*ngroups + g)*32 + l].
// bi/bj and the walk are in reference coordinates (i over A, j over B); tr maps them onto the kernel's (row, column).
__device__ __forceinline__ void traceback_and_emit(const SwArgs &a, const int lane, const int R, const int LB, const bool tr,
		const uint4 *__restrict__ trace, uint4 *tile, uint8_t *stage, const float score, const int bi, const int bj,
		PairRec *rec)
{
	const int rows_per_pass = 32 * R;
	const int ngroups = (LB + 31 + 3) >> 2;
	const uint32_t *tile32 = reinterpret_cast<const uint32_t *>(tile);
	int i = bi + 1, j = bj + 1;
	int state = 0;  // 0 = M, 1 = D, 2 = I
	uint32_t n = 0;
	int cur_p = -1, cur_G = -1;
	for (;;) {
		if (lane == 0)
			stage[n] = (uint8_t)(state == 0 ? 'M' : state == 1 ? 'D' : 'I');
		++n;
		const int ci = (state == 2) ? i : i - 1;
		const int cj = (state == 1) ? j : j - 1;
		const int krow = tr ? cj : ci, kcol = tr ? ci : cj;
		const int p = krow / rows_per_pass;
		const int rr = krow - p * rows_per_pass;
		const int srcl = rr / R;
		const int r = rr - srcl * R;
		const int s = kcol + srcl;
		const int g = s >> 2;
		const int G = g >> 2;
		if (p != cur_p || G != cur_G) {
			__syncwarp();
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				const int gg = 4 * G + k;
				if (gg < ngroups)
					tile[k * 32 + lane] = trace[((size_t)p * ngroups + gg) * 32 + lane];
			}
			__syncwarp();
			cur_p = p;
			cur_G = G;
		}
		const uint32_t w = tile32[(((g & 3) * 32 + srcl) << 2) + (s & 3)];
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

template <int R, bool TR, int W>
__device__ __forceinline__ void process_task(const SwArgs &a, unsigned char *smem, const uint32_t rowchain,
		const uint32_t begin, const uint32_t cnt)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const float *tab = reinterpret_cast<const float *>(smem + kSmemTab);
	float *p0 = reinterpret_cast<float *>(smem + kSmemP0);
	float *p1 = reinterpret_cast<float *>(smem + kSmemP1);
	const unsigned char *smem_p0 = smem + kSmemP0;
	uint4 *tile = reinterpret_cast<uint4 *>(smem + kSmemTiles) + warp * 128;

	const uint32_t LA = a.len_row[rowchain];  // kernel rows
	const uint64_t *profA = a.prof_row + a.off_row[rowchain];
	const int npass = (int)((LA + 32 * R - 1) / (32 * R));

	const bool have = (uint32_t)warp < cnt;
	uint32_t cidx = 0, slot = 0;
	int LB = 0;  // kernel columns
	const uint4 *colB = nullptr;
	if (have) {
		cidx = a.clist[begin + warp];
		if (a.cross) {
			const uint32_t ra = TR ? cidx : rowchain, rb = TR ? rowchain : cidx;  // reference (a, b)
			slot = (ra - a.a_begin) * a.nB + rb;
		} else {
			slot = a.cslot[begin + warp];
		}
		LB = (int)a.len_col[cidx];
		colB = a.coloff_col + 2 * a.off_col[cidx];
	}
	const int ngroups = (LB + 31 + 3) >> 2;
	const size_t gw = (size_t)blockIdx.x * W + warp;
	uint4 *trace = a.trace + gw * a.trace_stride;
	float2 *bnd = a.bnd + gw * a.bnd_stride;
	uint8_t *stage = a.stage + gw * a.stage_stride;

	float lbest = 0.0f;
	int lbi = 0x7fffffff, lbj = 0x7fffffff;  // kernel (row, column) of the best cell
	for (int pass = 0; pass < npass; ++pass) {
		__syncthreads();  // every warp is done with the previous pass's row table
		build_rowtab<R, W * 32>(p0, p1, tab, profA, LA, pass);
		__syncthreads();
		if (have)
			sw_pass<R, TR>(smem_p0, lane, pass, npass, LA, colB, LB, bnd, trace + (size_t)pass * ngroups * 32,
					a.open, a.ext, lbest, lbi, lbj);
	}
	if (!have)
		return;
	// first maximum in the reference's row-major (i, j) order: max score, then smallest i, then smallest j
#pragma unroll
	for (int o = 16; o >= 1; o >>= 1) {
		const float os = __shfl_xor_sync(kFull, lbest, o);
		const int oi = __shfl_xor_sync(kFull, lbi, o);
		const int oj = __shfl_xor_sync(kFull, lbj, o);
		bool take;
		if (!TR)  // i = row (unique per lane), j = column
			take = os > lbest || (os == lbest && oi < lbi);
		else      // i = column, j = row
			take = os > lbest || (os == lbest && (oj < lbj || (oj == lbj && oi < lbi)));
		if (take) {
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
	traceback_and_emit(a, lane, R, LB, TR, trace, tile, stage, lbest, TR ? lbj : lbi, TR ? lbi : lbj, rec);
}

// One kernel per (row-length class, orientation).  Class C handles R in [kClassRLo[C], kClassRHi[C]] with
// kClassWarps[C] warps: fewer rows per lane need fewer registers, so more warps fit and hide latency better.
template <int C, bool TR>
__global__ void __launch_bounds__(kClassWarps[C] * 32, 1) sw_affine_f32_tb_kernel(const SwArgs a)
{
	constexpr int W = kClassWarps[C];
	extern __shared__ __align__(16) unsigned char smem[];
	float *tab = reinterpret_cast<float *>(smem + kSmemTab);
	volatile int *bcast = reinterpret_cast<volatile int *>(smem + kSmemBcast);
	for (int k = threadIdx.x; k < RSK_TABLE_FLOATS; k += W * 32)
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
		uint32_t rowchain, begin, cnt;
		if (a.cross) {
			const uint32_t ridx = task / a.nseg;
			const uint32_t seg = task - ridx * a.nseg;
			rowchain = a.rowlist[ridx];
			begin = seg * W;
			cnt = min((uint32_t)W, a.ncols - begin);
		} else {
			rowchain = a.task_row[task];
			begin = a.task_begin[task];
			cnt = a.task_cnt[task];
		}
		int npass, R;
		sw_geometry(a.len_row[rowchain], npass, R);
		if (C == 0) {
			switch (R) {
			case 1: process_task<1, TR, W>(a, smem, rowchain, begin, cnt); break;
			case 2: process_task<2, TR, W>(a, smem, rowchain, begin, cnt); break;
			case 3: process_task<3, TR, W>(a, smem, rowchain, begin, cnt); break;
			case 4: process_task<4, TR, W>(a, smem, rowchain, begin, cnt); break;
			default: process_task<5, TR, W>(a, smem, rowchain, begin, cnt); break;
			}
		} else if (C == 1) {
			process_task<6, TR, W>(a, smem, rowchain, begin, cnt);
		} else {
			if (R == 7)
				process_task<7, TR, W>(a, smem, rowchain, begin, cnt);
			else
				process_task<8, TR, W>(a, smem, rowchain, begin, cnt);
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

// e-letters -> per-column row-table offsets in units of 16 bytes (code * 512 bytes / 16)
__global__ void make_coloff_kernel(const uint64_t *__restrict__ prof8, uint64_t total, uint4 *__restrict__ coloff)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= total)
		return;
	const uint64_t v = prof8[i];
	uint32_t o[8];
#pragma unroll
	for (int f = 0; f < 8; ++f)
		o[f] = (uint32_t)((v >> (8 * f)) & 0xff) * 32u;
	coloff[2 * i] = make_uint4(o[0], o[1], o[2], o[3]);
	coloff[2 * i + 1] = make_uint4(o[4], o[5], o[6], o[7]);
}

}  // namespace

size_t sw_smem_bytes() { return kSmemTotal; }

// uint4 units of packed trace one warp needs for a pair with npass passes and LB columns
uint64_t sw_trace_units(int npass, uint32_t LB) { return (uint64_t)npass * ((LB + 31 + 3) >> 2) * 32; }

template <int C, bool TR>
static int launch_sw_ct(const SwArgs &args, int grid, cudaStream_t stream)
{
	if (cudaFuncSetAttribute(sw_affine_f32_tb_kernel<C, TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTotal) != cudaSuccess)
		return -1;
	sw_affine_f32_tb_kernel<C, TR><<<grid, kClassWarps[C] * 32, kSmemTotal, stream>>>(args);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_sw(const SwArgs &args, int cls, int grid, cudaStream_t stream)
{
	const bool tr = args.tr != 0;
	switch (cls) {
	case 0: return tr ? launch_sw_ct<0, true>(args, grid, stream) : launch_sw_ct<0, false>(args, grid, stream);
	case 1: return tr ? launch_sw_ct<1, true>(args, grid, stream) : launch_sw_ct<1, false>(args, grid, stream);
	default: return tr ? launch_sw_ct<2, true>(args, grid, stream) : launch_sw_ct<2, false>(args, grid, stream);
	}
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

int launch_make_coloff(const uint64_t *prof8, uint64_t total, uint4 *coloff, cudaStream_t stream)
{
	if (total == 0)
		return 0;
	const int threads = 256;
	const unsigned blocks = (unsigned)((total + threads - 1) / threads);
	make_coloff_kernel<<<blocks, threads, 0, stream>>>(prof8, total, coloff);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace rsk
