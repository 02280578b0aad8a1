// sw_kernel.cu - K1: float affine-gap local Smith-Waterman with traceback over the on-the-fly 8-feature
// DSS score, one warp per (A,B) pair, one CTA per row chain (W pairs at a time).
//
// Replaces, per pair: DSSAligner::SetSMx_NoRev (dssaligner.cpp:529-596) + SWFast (sw.cpp:79-212) +
// TraceBackBitSW (sw.cpp:8-77).  Results are bit-identical: every cell performs the reference's fp32 adds
// and maxima (sw.cpp:119-197), only the order in which independent cells are visited differs.
//
// Design (B200, sm_100a), v4 "checkpointed forward, strip-recomputed traceback":
//  * The score matrix S is never materialised.  For the CTA's row chain a "row table" is staged in shared
//    memory for the current pass of 32*R rows (R = 1..12 rows per lane): plane 0 holds rows 0..3 of every lane
//    as float4 P0[e][lane], plane 1 rows 4..7, plane 2 rows 8..11 (float/float2/float4 by need), e =
//    (feature,letter) code 0..131, 512-byte stride per code.  A lane's vector sits in its own bank group, so
//    every LDS is conflict-free.  The column chain supplies, per DP column, its 8 e-letter bytes (one 64-bit
//    load; a PRMT + IMAD per feature turns a byte into a table address), so a cell's score costs 8 table
//    reads + 7 fp32 adds in the reference's order.
//  * DP wavefront: lane L owns R consecutive rows, lanes are skewed by one column per step, the last row's
//    (M,D) travels to the next lane by warp shuffle.  A chain longer than 32*12 rows is cut into passes; the
//    bottom row of a pass is parked in a small global (L2-resident) boundary row per pass.
//  * The forward sweep writes NO per-cell trace.  It keeps values only (FMNMX/FADD, about half the
//    instructions of a trace-producing cell) and every 16 steps parks the wavefront state of the warp
//    (2R+3 floats per lane) as a checkpoint.  The traceback (sw.cpp:8-77) then re-runs just the 16-step
//    strips its path crosses, this time producing the reference's 4 trace bits per cell (match source,
//    MD, MI) into a 4 KB L1-resident tile; random pairs align over a handful of columns, so this is
//    about one strip per pair.  Values are identical in both sweeps (same operations on the same inputs).
//  * The running maximum is tracked per lane with one FMNMX chain per step; the exact first-maximum rule of
//    the reference (row-major order, strict >) is restored in a rarely taken slow path and in the final
//    (score desc, row asc) warp reduction.
//  * Either chain of a pair can supply the rows (template TR): the scheduler puts the side with fewer chains on
//    the rows so that a CTA's row table is shared by a full set of warps.
//  * Persistent CTAs (one per SM) pull tasks from an atomic counter; one kernel per row-length class so that
//    short row blocks (few registers) run with more warps per SM than long ones.
#include <type_traits>

#include "rsk_internal.cuh"

namespace rsk {

namespace {

constexpr int kNLet = RSK_NLETTERS;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kStrip = kSwStripSteps;  // steps between two checkpoints

// shared memory carve-up
constexpr size_t kSmemTab = 0;                          // float[2192] weighted tables
constexpr size_t kSmemP0 = 8768;                        // plane 0: float4[132][32], rows 0..3 of each lane
constexpr size_t kPlaneBytes = (size_t)kNLet * 512;     // 67584
constexpr int kPlaneFloats = kNLet * 128;
__host__ __device__ constexpr int class_planes(int C) { return (C == 3 || C == 5) ? 3 : 2; }  // R > 8 needs the third plane
__host__ __device__ constexpr int planes_of_R(int R) { return R > 8 ? 3 : 2; }
// after the planes: 16 bytes for the task broadcast, then the warps' column-chain lists (kSwMaxWarps x kSwChainMax x 24 B)
__host__ __device__ constexpr size_t smem_bcast_off(int planes) { return kSmemP0 + (size_t)planes * kPlaneBytes; }
__host__ __device__ constexpr size_t smem_chains_off(int planes) { return smem_bcast_off(planes) + 16; }
__host__ __device__ constexpr size_t class_smem_base(int C) { return smem_chains_off(class_planes(C)) + (size_t)kSwMaxWarps * kSwChainMax * 24; }
#ifdef RSK_SW_TMA_CKPT
constexpr int kTmaStageBytes = 3072;  // per-warp staging of one checkpoint (ckpt_words <= 6 float4 per lane)
// the two-plane classes with 16 warps have shared memory to spare for the staging
__host__ __device__ constexpr bool tma_class(int planes, int warps) { return planes <= 2 && warps <= 16; }
#define class_smem(C) (class_smem_base(C) + (tma_class(class_planes(C), kClassWarps[C]) ? (size_t)kClassWarps[C] * kTmaStageBytes : 0))
#else
#define class_smem(C) class_smem_base(C)
#endif

// floats per lane in plane 1 / 2 for n rows living there (rounded up to a vector width)
__host__ __device__ constexpr int plane_width(int n) { return n <= 0 ? 0 : n == 1 ? 1 : n == 2 ? 2 : 4; }

// table value of rows beyond the end of the chain: such cells are hugely negative and can never be a maximum
constexpr float kPadScore = -1e30f;

// G = lanes per wavefront (32, or 16 for half-warp chains: lanes l and l + 16 then hold the same rows, each in its own bank
// group, so that the two half-warps read without bank conflicts)
template <int R, int NTHREADS, int G>
__device__ __forceinline__ void build_rowtab(float *planes, const float *tab, const uint64_t *__restrict__ profA,
		uint32_t LA, int pass)
{
	constexpr int W1 = plane_width(R - 4), W2 = plane_width(R - 8);
	constexpr int ROWS = G * R;
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
#pragma unroll
		for (int ll = l; ll < 32; ll += G) {
			if (r < 4)
				planes[e * 128 + ll * 4 + r] = v;
			else if (r < 8)
				planes[kPlaneFloats + e * 128 + ll * W1 + (r - 4)] = v;
			else
				planes[2 * kPlaneFloats + e * 128 + ll * W2 + (r - 8)] = v;
		}
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
__device__ __forceinline__ void add_vec(float (&acc)[N], const float (&v)[12])
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

// float4 units of one checkpoint per lane: M[R], I[R], M diagonal, (M, D) handed to the next lane
__host__ __device__ constexpr int ckpt_words(int R) { return (2 * R + 3 + 3) / 4; }

// One column chain of a warp's work list (kept in shared memory): a warp aligns up to kSwChain column chains against the
// CTA's row chain back to back, as ONE wavefront over their concatenated columns, so that the 31-step ramp is paid once per
// list instead of once per pair.  `base` = first concatenated column of the chain.
struct ColChain {
	const uint64_t *col;  // the chain's residues, 8 e-letter bytes each (DevChains::prof8)
	int LB;
	int base;
	uint32_t slot;        // record slot of the pair
	uint32_t pad;
};

// One sweep over rows [pass*32R, (pass+1)*32R) of the row chain and the concatenated columns of the warp's chains.
//   TRACE = false: the forward sweep.  Values only; saves a checkpoint every kStrip steps and tracks, per chain, the
//                  lane's first maximum (kernel coordinates), parked in `best` when the lane moves on to the next chain.
//   TRACE = true : re-run of the first `s_last` steps of strip `strip` from its checkpoint, writing one 64-bit
//                  trace word per lane and step (4 bits per row) into `tile[(step % kStrip)*32 + lane]`.
// A lane that has consumed the last column of a chain re-initialises its row state and continues with column 0 of the
// next chain at the very next step; its (M, D) hand-over registers keep serving the lane behind it, which is still one
// column back in the old chain - each chain therefore sees exactly the values of a sweep of its own.
// TR = false: kernel rows are the reference's A chain (index i), columns its B chain (j).
// TR = true : rows are the reference's B chain, columns its A chain.  The recurrence is the same with the
//             roles of the two gap states exchanged: the reference tests D (gap that consumes A) before I
//             (sw.cpp:136-147), and its first-maximum rule prefers the smaller i, then the smaller j.
// bnd_in / bnd_out: boundary row (indexed by concatenated column) written by the previous pass / by this one.
// G = 16: the two half-warps sweep their own chain lists (chs / nch / total are then per half); tmax / tmin = the larger /
// smaller `total` of the warp's wavefronts (loop bound / range of the predicate-free steps).
template <int R, bool TR, bool TRACE, int G>
__device__ __forceinline__ void sw_pass(const unsigned char *smem_p0, const int lane, const bool first, const bool last,
		const uint32_t row0, const uint32_t LA, const ColChain *chs, const int nch, const int total, const int tmax, const int tmin,
		const float2 *__restrict__ bnd_in, float2 *__restrict__ bnd_out, float4 *__restrict__ ck, const float open,
		const float ext, float4 *__restrict__ best, const int strip, const int s_last,
		unsigned long long *__restrict__ tile, float4 *stg = nullptr)
{
	const int sub = lane & (G - 1);
	constexpr int W1 = plane_width(R - 4), W2 = plane_width(R - 8);
	constexpr int NW4 = ckpt_words(R);
	// Per-lane byte offsets into the planes.  They are made opaque to the optimiser so that "offset*16 + lane
	// term" stays one IMAD; the loads themselves are ordinary shared-memory loads (LDS with the plane base folded into
	// the immediate) which the scheduler may hoist above the warp shuffles of the next step.
	uint32_t base0 = (uint32_t)lane * 16u;
	uint32_t base1 = (uint32_t)lane * (uint32_t)(W1 * 4);
	uint32_t base2 = (uint32_t)lane * (uint32_t)(W2 * 4);
	asm volatile("" : "+r"(base0));
	asm volatile("" : "+r"(base1));
	asm volatile("" : "+r"(base2));
	const unsigned char *const plane0 = smem_p0;
	const unsigned char *const plane1 = smem_p0 + kPlaneBytes;
	const unsigned char *const plane2 = smem_p0 + 2 * kPlaneBytes;
	float Mrow[R], Irow[R];
#pragma unroll
	for (int r = 0; r < R; ++r) {
		Mrow[r] = kNegInf;  // M[i_r+1][0]
		Irow[r] = kNegInf;  // I[i_r][0]
	}
	const bool lane0 = (sub == 0);
	const bool use_bnd = lane0 && !first;
	const bool put_bnd = (sub == G - 1) && !last;
	const float mdiag_init = (lane0 && first) ? 0.0f : kNegInf;  // M[i0][0]; M[0][0] = 0 (sw.cpp:116)
	float mdiag_next = mdiag_init;
	float outM = kNegInf, outD = kNegInf;
	const int nsteps = tmax + (G - 1);
	const int ngroups = (nsteps + 3) >> 2;

	auto save = [&](const int k) {
		float t[NW4 * 4];
#pragma unroll
		for (int r = 0; r < R; ++r) {
			t[r] = Mrow[r];
			t[R + r] = Irow[r];
		}
		t[2 * R] = mdiag_next;
		t[2 * R + 1] = outM;
		t[2 * R + 2] = outD;
#pragma unroll
		for (int w = 2 * R + 3; w < NW4 * 4; ++w)
			t[w] = 0.0f;
#ifdef RSK_SW_TMA_CKPT
		// Experiment (profiles/r2_tma_experiment.md): the checkpoint is staged in shared memory and leaves as ONE bulk copy
		// (cp.async.bulk shared -> global, the TMA engine) instead of NW4 STG.128 per lane.
		if (stg) {
			if (lane == 0)
				asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the engine has read the previous checkpoint
			__syncwarp();
#pragma unroll
			for (int w = 0; w < NW4; ++w)
				stg[w * 32 + lane] = make_float4(t[4 * w], t[4 * w + 1], t[4 * w + 2], t[4 * w + 3]);
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			__syncwarp();
			if (lane == 0) {
				const uint32_t src = (uint32_t)__cvta_generic_to_shared(stg);
				asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(ck + (size_t)k * NW4 * 32), "r"(src),
						"r"(NW4 * 512)
						: "memory");
				asm volatile("cp.async.bulk.commit_group;" ::: "memory");
			}
			return;
		}
#endif
#pragma unroll
		for (int w = 0; w < NW4; ++w)
			ck[((size_t)k * NW4 + w) * 32 + lane] = make_float4(t[4 * w], t[4 * w + 1], t[4 * w + 2], t[4 * w + 3]);
	};
	auto load = [&](const int k) {
		float t[NW4 * 4];
#pragma unroll
		for (int w = 0; w < NW4; ++w) {
			const float4 v = ck[((size_t)k * NW4 + w) * 32 + lane];
			t[4 * w] = v.x; t[4 * w + 1] = v.y; t[4 * w + 2] = v.z; t[4 * w + 3] = v.w;
		}
#pragma unroll
		for (int r = 0; r < R; ++r) {
			Mrow[r] = t[r];
			Irow[r] = t[R + r];
		}
		mdiag_next = t[2 * R];
		outM = t[2 * R + 1];
		outD = t[2 * R + 2];
	};

	// position of this lane: concatenated column v, chain c, column j of that chain
	int v = -sub;
	int c = 0;
	if (TRACE) {
		load(strip);
		v = kStrip * strip - sub;
		while (c + 1 < nch && v >= chs[c + 1].base)
			++c;
	}
	// a half-warp without chains (tail of a task list) never enters a matrix
	int LBc = nch > 0 ? chs[c].LB : -0x40000000;
	const uint64_t *colp = nch > 0 ? chs[c].col : nullptr;
	int j = nch > 0 ? v - chs[c].base : v;
	uint64_t cv = 0;
	if (j >= 0 && j < LBc)
		cv = __ldg(colp + j);
	float2 bn = make_float2(kNegInf, kNegInf);
	if (use_bnd && v < total)
		bn = bnd_in[v];
	// running first maximum of the current chain
	float lbest = 0.0f;
	int lbi = 0x7fffffff, lbj = 0x7fffffff;
	if (!TRACE && !first) {
		const float4 b = best[c * 32 + lane];
		lbest = b.x; lbi = __float_as_int(b.y); lbj = __float_as_int(b.z);
	}

	// CHECK = false: every lane is inside its matrix at this step and at the next one (no range predicates)
	auto step = [&](auto check_tag) -> unsigned long long {
		constexpr bool CHECK = decltype(check_tag)::value;
		const float inM = __shfl_up_sync(kFull, outM, 1);
		const float inD = __shfl_up_sync(kFull, outD, 1);
		const int jn = j + 1;
		uint64_t cn = cv;
		if (jn >= 0 && jn < LBc)
			cn = __ldg(colp + jn);
		float2 bn_next = bn;
		if (use_bnd && (!CHECK || v + 1 < total))
			bn_next = bnd_in[v + 1];
		uint32_t tw0 = 0, tw1 = 0;
		if (!CHECK || (j >= 0 && j < LBc)) {
			float d = lane0 ? bn.y : inD;  // D[i0][j]   (bn = -inf pair in the first pass)
			const float mdiag = mdiag_next;  // M[i0][j]
			mdiag_next = lane0 ? bn.x : inM;  // M[i0][j+1]
			const uint32_t clo = (uint32_t)cv, chi = (uint32_t)(cv >> 32);
			uint32_t co[8];  // e-letter of each feature
#pragma unroll
			for (int f = 0; f < RSK_NFEAT; ++f)
				co[f] = __byte_perm(f < 4 ? clo : chi, 0u, 0x4440u | (uint32_t)(f & 3));
			float S[R];
#pragma unroll
			for (int f = 0; f < RSK_NFEAT; ++f) {
				const uint32_t a0 = co[f] * 512u + base0;
				const float4 v0 = *reinterpret_cast<const float4 *>(plane0 + a0);
				float vv[12];
				vv[0] = v0.x; vv[1] = v0.y; vv[2] = v0.z; vv[3] = v0.w;
#pragma unroll
				for (int r = 4; r < 12; ++r)
					vv[r] = 0.0f;
				if (W1 == 1) {
					vv[4] = *reinterpret_cast<const float *>(plane1 + (co[f] * 512u + base1));
				} else if (W1 == 2) {
					const float2 t = *reinterpret_cast<const float2 *>(plane1 + (co[f] * 512u + base1));
					vv[4] = t.x; vv[5] = t.y;
				} else if (W1 == 4) {
					const float4 t = *reinterpret_cast<const float4 *>(plane1 + a0);
					vv[4] = t.x; vv[5] = t.y; vv[6] = t.z; vv[7] = t.w;
				}
				if (W2 == 1) {
					vv[8] = *reinterpret_cast<const float *>(plane2 + (co[f] * 512u + base2));
				} else if (W2 == 2) {
					const float2 t = *reinterpret_cast<const float2 *>(plane2 + (co[f] * 512u + base2));
					vv[8] = t.x; vv[9] = t.y;
				} else if (W2 == 4) {
					const float4 t = *reinterpret_cast<const float4 *>(plane2 + a0);
					vv[8] = t.x; vv[9] = t.y; vv[10] = t.z; vv[11] = t.w;
				}
				// feature 0 assigns, 1..7 accumulate in this order (dssaligner.cpp:557-595)
				if (f == 0) {
#pragma unroll
					for (int r = 0; r < R; ++r)
						S[r] = vv[r];
				} else {
					add_vec<R>(S, vv);
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
				const float mo = Mo[r];
				if (TRACE) {
					// the reference's compare-and-record form (sw.cpp:129-190)
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
					Mrow[r] = x;  // M[row+1][col+1]
					float dn = d + ext;
					if (mo >= dn) { dn = mo; code |= TR ? 8u : 4u; }
					d = dn;  // vertical gap state entering the next row
					float in2 = Ie[r];
					if (mo >= in2) { in2 = mo; code |= TR ? 4u : 8u; }
					Irow[r] = in2;  // horizontal gap state entering the next column
					if (r < 8)
						tw0 |= code << (4 * (r & 7));
					else
						tw1 |= code << (4 * (r & 7));
				} else {
					// same values without the bookkeeping: the selected value of each compare chain is the maximum
					// (no -0.0 can arise: 0 is only ever produced as +0.0 and sums of non-zero terms round to +0.0)
					float x = fmaxf(fmaxf(m, d), fmaxf(ii, 0.0f));
					x += S[r];
					xmax = fmaxf(xmax, x);
					Mrow[r] = x;
					d = fmaxf(mo, d + ext);
					Irow[r] = fmaxf(mo, Ie[r]);
				}
			}
			outM = Mrow[R - 1];
			outD = d;
			if (!TRACE) {
				if (put_bnd)
					bnd_out[v] = make_float2(outM, outD);
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
		}
		j = jn;
		++v;
		cv = cn;
		bn = bn_next;
		if (j == LBc) {
			// this lane has consumed its chain: park the chain's maximum and start the next chain with a fresh row state
			if (!TRACE)
				best[c * 32 + lane] = make_float4(lbest, __int_as_float(lbi), __int_as_float(lbj), 0.0f);
			if (c + 1 < nch) {
				++c;
				LBc = chs[c].LB;
				colp = chs[c].col;
				j = 0;
				cv = __ldg(colp);
#pragma unroll
				for (int r = 0; r < R; ++r) {
					Mrow[r] = kNegInf;
					Irow[r] = kNegInf;
				}
				mdiag_next = mdiag_init;
				if (!TRACE) {
					lbest = 0.0f; lbi = 0x7fffffff; lbj = 0x7fffffff;
					if (!first) {
						const float4 b = best[c * 32 + lane];
						lbest = b.x; lbi = __float_as_int(b.y); lbj = __float_as_int(b.z);
					}
				}
			}
		}
		return (unsigned long long)tw0 | ((unsigned long long)tw1 << 32);
	};

	if (TRACE) {
		// s_last = number of steps to re-run from the start of the strip (warp-uniform; `strip` itself may differ between the
		// two half-warps, each re-running the strip its own walk needs)
		for (int n = 0; n < s_last; ++n) {
			const unsigned long long tw = step(std::true_type{});
			tile[n * 32 + lane] = tw;
		}
	} else {
		for (int g = 0; g < ngroups; ++g) {
			const int s0 = 4 * g;
			if ((g & (kStrip / 4 - 1)) == 0)
				save(g / (kStrip / 4));
			if (s0 >= G - 1 && s0 + 4 < tmin) {
				step(std::false_type{});
				step(std::false_type{});
				step(std::false_type{});
				step(std::false_type{});
			} else {
				step(std::true_type{});
				step(std::true_type{});
				step(std::true_type{});
				step(std::true_type{});
			}
		}
#ifdef RSK_SW_TMA_CKPT
		if (stg) {  // the checkpoints are read back (generic proxy) by the traceback
			if (lane == 0)
				asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
			asm volatile("fence.proxy.async.global;" ::: "memory");
			__syncwarp();
		}
#endif
	}
}

// Resumable warp-cooperative traceback (sw.cpp:8-77).  The walk is in reference coordinates (i over A, j over B); TR maps
// them onto the kernel's (row, column).  Trace nibble of cell (row, col) of the chain whose first concatenated column is
// `cbase`: pass p = row / (32R), lane l = (row % 32R) / R, nibble r = row % R, step s = cbase + col + l, strip k = s / kStrip,
// word tile[(s % kStrip)*32 + l].  Within a pass the step of a walk never increases, so a strip is re-run up to the step
// at which the walk enters it (a later chain of the same warp may need a few more steps of the same strip).
struct TbState {
	int i, j;      // reference DP coordinates of the walk (1-based cell indices as in sw.cpp)
	int state;     // 0 = M, 1 = D, 2 = I
	uint32_t n;    // columns emitted so far
};
struct TileCache {
	int p, k, last;  // pass, strip and last step the tile currently holds
};

// Walks while the path stays in pass `pass` (whose row table is the one in shared memory).  Returns true when the path is
// complete, false when it continues in the pass above.
// G / grp: lanes per wavefront and the half-warp (0 or 1) that swept the chain being walked; chs / nch / total are the
// calling lane's own wavefront's (they feed the strip re-runs, which every lane of the warp takes part in).
template <int R, bool TR, int G>
__device__ __forceinline__ bool traceback_in_pass(const unsigned char *smem_p0, const int lane, const int pass, const int npass,
		const uint32_t LA, const ColChain *chs, const int nch, const int total, const int tmax, const int tmin, const int grp,
		const int cbase, float2 *__restrict__ bnd,
		const uint32_t bnd_pass_stride, float4 *__restrict__ ck, const int nstrips, const float open, const float ext,
		unsigned long long *__restrict__ tile, uint8_t *__restrict__ stage, TbState &t, TileCache &tc)
{
	constexpr int rows_per_pass = G * R;
	for (;;) {
		const int ci = (t.state == 2) ? t.i : t.i - 1;
		const int cj = (t.state == 1) ? t.j : t.j - 1;
		const int krow = TR ? cj : ci, kcol = TR ? ci : cj;
		const int p = krow / rows_per_pass;
		if (p != pass)
			return false;
		if (lane == 0)
			stage[t.n] = (uint8_t)(t.state == 0 ? 'M' : t.state == 1 ? 'D' : 'I');
		++t.n;
		const int rr = krow - p * rows_per_pass;
		const int srcl = rr / R;
		const int r = rr - srcl * R;
		const int s = cbase + kcol + srcl;
		const int k = s / kStrip;
		if (p != tc.p || k != tc.k || s > tc.last) {
			__syncwarp();
			sw_pass<R, TR, true, G>(smem_p0, lane, p == 0, p == npass - 1, (uint32_t)p * rows_per_pass + (uint32_t)(lane & (G - 1)) * R, LA,
					chs, nch, total, tmax, tmin, bnd + (size_t)(p > 0 ? p - 1 : 0) * bnd_pass_stride, nullptr,
					ck + (size_t)p * nstrips * ckpt_words(R) * 32, open, ext, nullptr, k, s - kStrip * k + 1, tile);
			__syncwarp();
			tc.p = p;
			tc.k = k;
			tc.last = s;
		}
		const unsigned long long w = tile[(s & (kStrip - 1)) * 32 + grp * G + srcl];
		const uint32_t nib = (uint32_t)(w >> (4 * r)) & 15u;
		if (t.state == 0) {
			--t.i; --t.j;
			const uint32_t src = nib & 3u;
			if (src == 3u)
				return true;
			t.state = (int)src;  // 0 M, 1 D, 2 I
		} else if (t.state == 1) {
			--t.i;
			t.state = (nib & 4u) ? 0 : 1;
		} else {
			--t.j;
			t.state = (nib & 8u) ? 0 : 2;
		}
	}
}

// Half-warp chains (always one pass): chain k of BOTH wavefronts is walked at the same time, each half-warp following its own
// path.  When a walk needs trace words that its half of the tile does not hold, the warp re-runs strips once for both: each
// half re-runs the strip ITS walk is in (a half that needs nothing repeats the strip it has), for as many steps as the longer
// request - the per-pair strip re-run is the dominant cost of short chains, and pairing halves it.  `t`, `cbase`, `stage`
// and `active` are per half-warp.
template <int R, bool TR>
__device__ __forceinline__ void traceback_halves(const unsigned char *smem_p0, const int lane, const uint32_t LA, const ColChain *chs,
		const int nch, const int total, const int tmax, const int tmin, const int cbase, float2 *__restrict__ bnd,
		float4 *__restrict__ ck, const float open, const float ext, unsigned long long *__restrict__ tile,
		uint8_t *__restrict__ stage, TbState &t, const bool active)
{
	constexpr int G = 16;
	const int sub = lane & (G - 1), grp = lane / G;
	int tk = -1, tlast = -1;  // strip held by this half's part of the tile, last step of it that is valid
	bool done = !active;
	for (;;) {
		int need_k = -1, need_s = 0;
		while (!done) {
			const int ci = (t.state == 2) ? t.i : t.i - 1;
			const int cj = (t.state == 1) ? t.j : t.j - 1;
			const int krow = TR ? cj : ci, kcol = TR ? ci : cj;
			const int srcl = krow / R;
			const int r = krow - srcl * R;
			const int s = cbase + kcol + srcl;
			const int k = s / kStrip;
			if (k != tk || s > tlast) {
				need_k = k;
				need_s = s;
				break;
			}
			if (sub == 0)
				stage[t.n] = (uint8_t)(t.state == 0 ? 'M' : t.state == 1 ? 'D' : 'I');
			++t.n;
			const unsigned long long w = tile[(s & (kStrip - 1)) * 32 + grp * G + srcl];
			const uint32_t nib = (uint32_t)(w >> (4 * r)) & 15u;
			if (t.state == 0) {
				--t.i; --t.j;
				const uint32_t src = nib & 3u;
				if (src == 3u)
					done = true;
				else
					t.state = (int)src;  // 0 M, 1 D, 2 I
			} else if (t.state == 1) {
				--t.i;
				t.state = (nib & 4u) ? 0 : 1;
			} else {
				--t.j;
				t.state = (nib & 8u) ? 0 : 2;
			}
		}
		if (!__ballot_sync(kFull, need_k >= 0))
			break;
		const int k = need_k >= 0 ? need_k : max(tk, 0);
		int cnt = need_k >= 0 ? need_s - kStrip * need_k + 1 : 1;
		cnt = max(cnt, __shfl_xor_sync(kFull, cnt, G));
		__syncwarp();
		sw_pass<R, TR, true, G>(smem_p0, lane, true, true, (uint32_t)sub * R, LA, chs, nch, total, tmax, tmin, bnd, nullptr, ck, open, ext,
				nullptr, k, cnt, tile);
		__syncwarp();
		tlast = (k == tk) ? max(tlast, kStrip * k + cnt - 1) : kStrip * k + cnt - 1;
		tk = k;
	}
}

// publish: reserve pool space, copy the staged (reversed) path forward, fill the record
__device__ __forceinline__ void emit_path(const SwArgs &a, const int lane, const uint8_t *stage, const float score, const TbState &t,
		PairRec *rec)
{
	const uint32_t n = t.n;
	unsigned long long off = 0;
	if (lane == 0)
		off = atomicAdd(a.pool_cursor, (unsigned long long)n);
	off = __shfl_sync(kFull, off, 0);
	__syncwarp();
	for (uint32_t k = lane; k < n; k += 32)
		a.pool[off + k] = stage[n - 1 - k];
	if (lane == 0) {
		rec->score = score;
		rec->lo_a = (uint32_t)t.i;
		rec->lo_b = (uint32_t)t.j;
		rec->path_len = n;
		rec->path_off = off;
	}
}

template <int R, bool TR, int W, int G>
__device__ __forceinline__ void process_task(const SwArgs &a, unsigned char *smem, const uint32_t rowchain,
		const uint32_t begin, const uint32_t cnt)
{
	constexpr int NG = 32 / G;            // wavefronts per warp
	constexpr int CH = G == 16 ? kSwChainHalf : kSwChain;  // column chains per warp (= sw_class_chains of the task's class)
	constexpr int CG = CH / NG;           // column chains per wavefront
	static_assert(CH % NG == 0, "chains per warp must split evenly over the half-warps");
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int grp = lane / G;
	const float *tab = reinterpret_cast<const float *>(smem + kSmemTab);
	float *planes = reinterpret_cast<float *>(smem + kSmemP0);
	const unsigned char *smem_p0 = smem + kSmemP0;
	ColChain *chs = reinterpret_cast<ColChain *>(smem + smem_chains_off(planes_of_R(R))) + warp * kSwChainMax;

	const uint32_t LA = a.len_row[rowchain];  // kernel rows
	const uint64_t *profA = a.prof_row + a.off_row[rowchain];
	const int npass = (int)((LA + G * R - 1) / (G * R));

	// this warp's column chains: entries warp, warp + W, warp + 2W ... of the task's list (the list is sorted by length, so
	// every warp gets a similar total); entry kk goes to wavefront kk % NG as its chain kk / NG
	int nchg[NG], totg[NG];
#pragma unroll
	for (int g = 0; g < NG; ++g)
		nchg[g] = totg[g] = 0;
#pragma unroll
	for (int kk = 0; kk < CH; ++kk) {
		const uint32_t e = (uint32_t)(kk * W + warp);
		if (e < cnt) {
			const int g = kk % NG;
			const uint32_t cidx = a.clist[begin + e];
			const int LB = (int)a.len_col[cidx];
			uint32_t slot;
			if (a.cross) {
				const uint32_t ra = TR ? cidx : rowchain, rb = TR ? rowchain : cidx;  // reference (a, b)
				slot = (ra - a.a_begin) * a.nB + rb;
			} else {
				slot = a.cslot[begin + e];
			}
			if (lane == 0) {
				ColChain &cc = chs[g * CG + nchg[g]];
				cc.col = a.prof_col + a.off_col[cidx];
				cc.LB = LB;
				cc.base = totg[g];
				cc.slot = slot;
			}
			totg[g] += LB;
			++nchg[g];
		}
	}
	__syncwarp();
	int tmax = 0, tmin = 0x7fffffff, nall = 0;
#pragma unroll
	for (int g = 0; g < NG; ++g) {
		tmax = max(tmax, totg[g]);
		tmin = min(tmin, nchg[g] ? totg[g] : 0);
		nall += nchg[g];
	}
	const bool have = nall > 0;
	const ColChain *mychs = chs + grp * CG;
	const int mynch = nchg[NG == 1 ? 0 : grp], mytot = totg[NG == 1 ? 0 : grp];
	const int nstrips = (((tmax + (G - 1) + 3) >> 2) + kStrip / 4 - 1) / (kStrip / 4);
	const size_t gw = (size_t)blockIdx.x * W + warp;
	float4 *ck = a.ckpt + gw * a.ckpt_stride;
	float2 *bnd = a.bnd + gw * a.bnd_stride;
	uint8_t *stage = a.stage + gw * a.stage_stride;
	unsigned long long *tile = a.tile + gw * (kStrip * 32);
	float4 *best = a.best + gw * (kSwChainMax * 32);
	float4 *stg = nullptr;
#ifdef RSK_SW_TMA_CKPT
	if (tma_class(planes_of_R(R), W))
		stg = reinterpret_cast<float4 *>(smem + smem_chains_off(planes_of_R(R)) + (size_t)kSwMaxWarps * kSwChainMax * 24) + (size_t)warp * (kTmaStageBytes / 16);
#endif

	for (int pass = 0; pass < npass; ++pass) {
		__syncthreads();  // every warp is done with the previous row table
		build_rowtab<R, W * 32, G>(planes, tab, profA, LA, pass);
		__syncthreads();
		if (have)
			sw_pass<R, TR, false, G>(smem_p0, lane, pass == 0, pass == npass - 1, (uint32_t)pass * G * R + (uint32_t)(lane & (G - 1)) * R, LA,
					mychs, mynch, mytot, tmax, tmin, bnd + (size_t)(pass > 0 ? pass - 1 : 0) * a.bnd_pass_stride,
					bnd + (size_t)pass * a.bnd_pass_stride, ck + (size_t)pass * nstrips * ckpt_words(R) * 32, a.open, a.ext, best, 0, 0, nullptr,
					stg);
	}
	// per chain: first maximum in the reference's row-major (i, j) order: max score, then smallest i, then smallest j.
	// Chain k of every wavefront is reduced at once (inside its half-warp), then broadcast to the whole warp, which walks the
	// paths one after the other.
	float score[CH];
	TbState tb[CH];
	unsigned active = 0;
	__syncwarp();
#pragma unroll
	for (int k = 0; k < CG; ++k) {
		float lbest = 0.0f;
		int lbi = 0x7fffffff, lbj = 0x7fffffff;
		if (k < mynch) {
			const float4 b = best[k * 32 + lane];
			lbest = b.x; lbi = __float_as_int(b.y); lbj = __float_as_int(b.z);
		}
#pragma unroll
		for (int o = G / 2; o >= 1; o >>= 1) {
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
#pragma unroll
		for (int g = 0; g < NG; ++g) {
			const int kk = g * CG + k;  // index of the chain in the warp's list in shared memory
			score[kk] = 0.0f;
			tb[kk].i = tb[kk].j = 0; tb[kk].state = 0; tb[kk].n = 0;
			if (k < nchg[g]) {
				const float gb = __shfl_sync(kFull, lbest, g * G);
				const int gi = __shfl_sync(kFull, lbi, g * G), gj = __shfl_sync(kFull, lbj, g * G);
				PairRec *rec = a.rec + chs[kk].slot;
				if (gb == 0.0f) {  // sw.cpp:200-201: no positive cell -> score 0, empty path
					if (lane == 0) {
						rec->score = 0.0f;
						rec->lo_a = 0xffffffffu;
						rec->lo_b = 0xffffffffu;
						rec->path_len = 0;
						rec->path_off = 0;
					}
				} else {
					active |= 1u << kk;
					score[kk] = gb;
					tb[kk].i = (TR ? gj : gi) + 1;
					tb[kk].j = (TR ? gi : gj) + 1;
				}
			}
		}
	}
	TileCache tc;
	tc.p = -1; tc.k = -1; tc.last = -1;
	__syncwarp();  // checkpoints and boundary rows written by other lanes of this warp are visible
	if constexpr (G == 16) {
		// one pass; the two wavefronts' chains are walked in pairs
#pragma unroll
		for (int k = 0; k < CG; ++k) {
			if (!(active & ((1u << k) | (1u << (CG + k)))))
				continue;
			const int mykk = grp * CG + k;
			TbState t = grp ? tb[(NG - 1) * CG + k] : tb[k];
			traceback_halves<R, TR>(smem_p0, lane, LA, mychs, mynch, mytot, tmax, tmin, chs[mykk].base, bnd, ck, a.open, a.ext, tile,
					stage + (size_t)mykk * a.stage_chain_stride, t, (active >> mykk) & 1u);
#pragma unroll
			for (int g = 0; g < NG; ++g) {
				const int kk = g * CG + k;
				if (active & (1u << kk)) {
					TbState tg;
					tg.i = __shfl_sync(kFull, t.i, g * G);
					tg.j = __shfl_sync(kFull, t.j, g * G);
					tg.n = __shfl_sync(kFull, t.n, g * G);
					tg.state = 0;
					__syncwarp();
					emit_path(a, lane, stage + (size_t)kk * a.stage_chain_stride, score[kk], tg, a.rec + chs[kk].slot);
				}
			}
		}
	} else {
	for (int p = npass - 1; p >= 0; --p) {
		if (p != npass - 1) {
			// the path of some warp continues above this pass: restage that pass's row table for the whole CTA
			build_rowtab<R, W * 32, G>(planes, tab, profA, LA, p);
			__syncthreads();
		}
#pragma unroll
		for (int kk = 0; kk < CH; ++kk) {
			if (active & (1u << kk)) {
				if (traceback_in_pass<R, TR, G>(smem_p0, lane, p, npass, LA, mychs, mynch, mytot, tmax, tmin, kk / CG, chs[kk].base, bnd,
							a.bnd_pass_stride, ck, nstrips, a.open, a.ext, tile, stage + (size_t)kk * a.stage_chain_stride, tb[kk], tc)) {
					emit_path(a, lane, stage + (size_t)kk * a.stage_chain_stride, score[kk], tb[kk], a.rec + chs[kk].slot);
					active &= ~(1u << kk);
				}
			}
		}
		if (p > 0 && !__syncthreads_or(active != 0 ? 1 : 0))
			break;
	}
	}
	__syncwarp();  // the chain list in shared memory is rewritten by the next task
}

// One kernel per (row-length class, orientation).  Class C handles a range of R with kClassWarps[C] warps: fewer rows
// per lane need fewer registers, so more warps fit and hide latency better.
template <int C, bool TR>
__global__ void __launch_bounds__(kClassWarps[C] * 32, 1) sw_affine_f32_tb_kernel(const SwArgs a)
{
	constexpr int W = kClassWarps[C];
	extern __shared__ __align__(16) unsigned char smem[];
	float *tab = reinterpret_cast<float *>(smem + kSmemTab);
	volatile int *bcast = reinterpret_cast<volatile int *>(smem + smem_bcast_off(class_planes(C)));
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
			begin = seg * (W * sw_class_chains(C));
			cnt = min((uint32_t)(W * sw_class_chains(C)), a.ncols - begin);
		} else {
			rowchain = a.task_row[task];
			begin = a.task_begin[task];
			cnt = a.task_cnt[task];
		}
		int npass, R;
		sw_geometry(a.len_row[rowchain], npass, R);
		// the class fixes the wavefront width: 0, 1, 4, 5 hold half-warp chains (<= 192 residues), 2 and 3 full-warp chains
		if (C == 0) {
			switch (R) {
			case 1: process_task<1, TR, W, 16>(a, smem, rowchain, begin, cnt); break;
			case 2: process_task<2, TR, W, 16>(a, smem, rowchain, begin, cnt); break;
			case 3: process_task<3, TR, W, 16>(a, smem, rowchain, begin, cnt); break;
			case 4: process_task<4, TR, W, 16>(a, smem, rowchain, begin, cnt); break;
			default: process_task<5, TR, W, 16>(a, smem, rowchain, begin, cnt); break;
			}
		} else if (C == 1) {
			process_task<6, TR, W, 16>(a, smem, rowchain, begin, cnt);
		} else if (C == 2 || C == 4) {
			constexpr int G = C == 2 ? 32 : 16;
			if (R == 7)
				process_task<7, TR, W, G>(a, smem, rowchain, begin, cnt);
			else
				process_task<8, TR, W, G>(a, smem, rowchain, begin, cnt);
		} else {
			constexpr int G = C == 3 ? 32 : 16;
			switch (R) {
			case 9: process_task<9, TR, W, G>(a, smem, rowchain, begin, cnt); break;
			case 10: process_task<10, TR, W, G>(a, smem, rowchain, begin, cnt); break;
			case 11: process_task<11, TR, W, G>(a, smem, rowchain, begin, cnt); break;
			default: process_task<12, TR, W, G>(a, smem, rowchain, begin, cnt); break;
			}
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

size_t sw_smem_bytes() { return class_smem(3); }

// float4 units of checkpoints one warp needs for npass passes over `LB` concatenated columns (any R)
uint64_t sw_ckpt_units(int npass, uint64_t LB)
{
	const uint64_t nstrips = (((LB + 31 + 3) >> 2) + kStrip / 4 - 1) / (kStrip / 4);
	return (uint64_t)npass * nstrips * ckpt_words(kMaxRowsPerLane) * 32;
}

template <int C, bool TR>
static int launch_sw_ct(const SwArgs &args, int grid, cudaStream_t stream)
{
	if (cudaFuncSetAttribute(sw_affine_f32_tb_kernel<C, TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)class_smem(C)) != cudaSuccess)
		return -1;
	sw_affine_f32_tb_kernel<C, TR><<<grid, kClassWarps[C] * 32, class_smem(C), stream>>>(args);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_sw(const SwArgs &args, int cls, int grid, cudaStream_t stream)
{
	const bool tr = args.tr != 0;
	switch (cls) {
	case 0: return tr ? launch_sw_ct<0, true>(args, grid, stream) : launch_sw_ct<0, false>(args, grid, stream);
	case 1: return tr ? launch_sw_ct<1, true>(args, grid, stream) : launch_sw_ct<1, false>(args, grid, stream);
	case 2: return tr ? launch_sw_ct<2, true>(args, grid, stream) : launch_sw_ct<2, false>(args, grid, stream);
	case 3: return tr ? launch_sw_ct<3, true>(args, grid, stream) : launch_sw_ct<3, false>(args, grid, stream);
	case 4: return tr ? launch_sw_ct<4, true>(args, grid, stream) : launch_sw_ct<4, false>(args, grid, stream);
	default: return tr ? launch_sw_ct<5, true>(args, grid, stream) : launch_sw_ct<5, false>(args, grid, stream);
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

}  // namespace rsk
