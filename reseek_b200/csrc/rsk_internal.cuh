// rsk_internal.cuh - shared declarations of libreseek_b200 (context, device chain store, kernel launchers).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/reseek_b200.h"

namespace rsk {

constexpr float kNegInf = -9e9f;  // finite "minus infinity" of the reference DP (xdpmem.h:6)

// Feature letters are stored pre-offset ("e-letters"): e = kFeatBase[f] + letter, 0..131, so that one byte
// indexes the 132-entry per-row score table directly.
__host__ __device__ constexpr int feat_base(int f) { return f == 0 ? 0 : 20 + 16 * (f - 1); }
__host__ __device__ constexpr int feat_alpha(int f) { return f == 0 ? 20 : 16; }
__host__ __device__ constexpr int feat_table_off(int f) { return f == 0 ? 0 : 400 + 256 * (f - 1); }

// ---- device-resident chains (SoA) ----
struct DevChains {
	uint32_t n = 0;
	uint64_t total = 0;
	uint32_t *len = nullptr;    // [n]
	uint64_t *off = nullptr;    // [n] residue offset of each chain
	uint64_t *prof8 = nullptr;  // [total] 8 e-letter bytes per residue (byte f = feature f)
	uint8_t *mu = nullptr;      // [total] or null
	float *x = nullptr, *y = nullptr, *z = nullptr;  // [total] each
	float *selfrev = nullptr;   // [n]
};

// One record per scheduled pair, written by the kernels (64 bytes).
struct PairRec {
	float score;
	uint32_t lo_a, lo_b, path_len;
	unsigned long long path_off;
	uint32_t hi_a, hi_b, ids, gaps;
	float lddt, ts;
	int32_t mu_fwd, mu_rev;
	uint32_t flags;
	uint32_t pad;
};
static_assert(sizeof(PairRec) == 64, "PairRec layout");

// Device hit sink (sink_kernel.cu): a kept record with its chain indices in the numbering of the whole (unsharded) search.
struct SinkRec {
	PairRec r;        // path_off re-based to the sink's path pool
	uint32_t a, b;
};
static_assert(sizeof(SinkRec) == 72, "SinkRec layout");
struct SinkArgs {
	const PairRec *rec; uint32_t npairs;       // the batch
	const uint8_t *pool;                       // the batch's path pool
	uint32_t cross, a_begin, nB;               // cross: pair k = (a_begin + k / nB, k % nB); else pair_a/pair_b
	const uint32_t *pair_a, *pair_b;
	uint32_t a_base, b_base;                   // added to the local chain indices (this rank's block of the sharded side)
	uint32_t keep_all, want_paths, report_no_evalue;
	float ts_lo;                               // conservative emit test: TS >= ts_lo (the root re-applies the exact E-value test)
	uint32_t *keep, *plen, *keep_scan;         // [npairs] scratch
	unsigned long long *plen_scan;             // [npairs] scratch
	SinkRec *out_rec; uint8_t *out_pool;
	unsigned long long *totals;                // [4] device: records, path bytes, pairs with E-value, Mu-rejected pairs
};
size_t sink_scan_tmp_bytes(uint32_t npairs);
int launch_sink_append(const SinkArgs &a, void *tmp, size_t tmp_bytes, cudaStream_t st);

// ---- SW kernel geometry ----
constexpr int kSwWarps = 16;                 // warps per CTA of the Mu filter kernel
constexpr int kSwThreads = kSwWarps * 32;
constexpr int kMuTaskCols = 4 * kSwWarps;     // column chains per task of the packed 16-bit Mu filter (two packed pairs per warp)
constexpr uint32_t kMu16MaxLen = 8000;       // 4*L must stay below 2^15 for the 16-bit lanes
constexpr int kMaxRowsPerLane = 12;          // R: DP rows owned by one lane within a pass
constexpr int kRowsPerPassMax = 32 * kMaxRowsPerLane;
#ifndef RSK_SW_STRIP
#define RSK_SW_STRIP 16
#endif
constexpr int kSwStripSteps = RSK_SW_STRIP;  // wavefront steps between two checkpoints of the forward sweep (multiple of 4; A/B: 8, 32)
#ifndef RSK_SW_CHAIN
#define RSK_SW_CHAIN 4
#endif
constexpr int kSwChain = RSK_SW_CHAIN;       // column chains a warp aligns back to back as one wavefront (one ramp per list)
// The half-warp classes (chains <= 192 residues) split a warp's list over two wavefronts, and their sweeps are short, so the ramp
// weighs more: eight chains per warp = four per wavefront (A/B profiles/r2_chain_ab.log: L = 100 +5.5 %; L = 300, a full-warp
// class, loses 1 % with eight and keeps four)
#ifndef RSK_SW_CHAIN_HALF
#define RSK_SW_CHAIN_HALF 8
#endif
constexpr int kSwChainHalf = RSK_SW_CHAIN_HALF;
constexpr int kSwChainMax = kSwChain > kSwChainHalf ? kSwChain : kSwChainHalf;
__host__ __device__ constexpr int sw_class_chains(int cls) { return (cls == 2 || cls == 3) ? kSwChain : kSwChainHalf; }
// chains per warp for the SW tasks of a row chain with `n` surviving partners: aim at eight tasks per row chain before lists grow
__host__ __device__ inline uint32_t sw_task_chains(uint32_t n, int warps, int max_chains)
{
	const uint32_t c = (n + (uint32_t)warps * 8 - 1) / ((uint32_t)warps * 8);
	return c < 1 ? 1u : c > (uint32_t)max_chains ? (uint32_t)max_chains : c;
}
// SW kernel classes (one kernel each, so that every class gets its own register allocation and warps per CTA = pairs per task):
//   half-warp chains (<= 192 residues): 0: R <= 5 | 1: R == 6 | 4: R = 7..8 | 5: R = 9..12
//   full-warp chains:                   2: R = 7..8 | 3: R = 9..12   (a chain above 192 residues never has R < 7)
constexpr int kSwClasses = 6;
#ifndef RSK_CLASS_W0
#define RSK_CLASS_W0 32
#endif
#ifndef RSK_CLASS_W1
#define RSK_CLASS_W1 16
#endif
#ifndef RSK_CLASS_W2
#define RSK_CLASS_W2 16
#endif
#ifndef RSK_CLASS_W3
#define RSK_CLASS_W3 20
#endif
#ifndef RSK_CLASS_W4
#define RSK_CLASS_W4 16
#endif
#ifndef RSK_CLASS_W5
#define RSK_CLASS_W5 20
#endif
constexpr int kClassWarps[kSwClasses] = {RSK_CLASS_W0, RSK_CLASS_W1, RSK_CLASS_W2, RSK_CLASS_W3, RSK_CLASS_W4, RSK_CLASS_W5};
__host__ __device__ constexpr int sw_cmax(int a, int b) { return a > b ? a : b; }
constexpr int kSwMaxWarps = sw_cmax(sw_cmax(sw_cmax(RSK_CLASS_W0, RSK_CLASS_W1), sw_cmax(RSK_CLASS_W2, RSK_CLASS_W3)), sw_cmax(RSK_CLASS_W4, RSK_CLASS_W5));
__host__ __device__ inline int sw_class_of_R(int R, bool half)
{
	if (half)
		return R <= 5 ? 0 : R == 6 ? 1 : R <= 8 ? 4 : 5;
	return R <= 8 ? 2 : 3;
}
__host__ __device__ inline int sw_class_warps(int c)
{
	return c == 0 ? RSK_CLASS_W0 : c == 1 ? RSK_CLASS_W1 : c == 2 ? RSK_CLASS_W2 : c == 3 ? RSK_CLASS_W3 : c == 4 ? RSK_CLASS_W4 : RSK_CLASS_W5;
}

// Row chains of at most 16*12 residues are swept by HALF warps: lanes 0-15 and lanes 16-31 run two independent wavefronts
// (different column chains, same row chain), each lane owning twice as many rows.  The per-step overhead of the wavefront
// (shuffles, column fetch, chain bookkeeping) is then shared by twice as many cells per lane - short chains were issue
// bound at 47 instructions per cell (R = 5) against 29 at R = 10 - the ramp is 15 steps instead of 31, and 100 rows pad to
// 112 instead of 128.
#ifndef RSK_SW_HALF_MAX
#define RSK_SW_HALF_MAX (16 * 12)
#endif
constexpr uint32_t kSwHalfMaxLen = RSK_SW_HALF_MAX;
__host__ __device__ inline bool sw_half(uint32_t LA) { return LA <= kSwHalfMaxLen; }

// How a chain of LA rows is cut into passes of 32*R rows (R rows per lane; 16*R rows in one pass for half-warp chains).
__host__ __device__ inline void sw_geometry(uint32_t LA, int &npass, int &R)
{
	if (sw_half(LA)) {
		npass = 1;
		R = (int)((LA + 15u) / 16u);
		if (R < 1) R = 1;
		return;
	}
	npass = (int)((LA + kRowsPerPassMax - 1) / kRowsPerPassMax);
	if (npass < 1) npass = 1;
	R = (int)((LA + 32u * npass - 1) / (32u * npass));
	if (R < 1) R = 1;
}
__host__ __device__ inline int sw_class_of_len(uint32_t L)
{
	int npass, R;
	sw_geometry(L, npass, R);
	return sw_class_of_R(R, sw_half(L));
}

// K1 arguments.  "row"/"col" are the kernel's view; tr says which reference slot supplies the rows.
struct SwArgs {
	const uint64_t *prof_row; const uint64_t *off_row; const uint32_t *len_row;
	const uint64_t *prof_col; const uint64_t *off_col; const uint32_t *len_col;
	uint32_t tr;             // 0: rows = reference A (DSSAligner query slot), 1: rows = reference B
	// tasks: one task = one row chain x up to W column chains (W = warps of the kernel class)
	uint32_t ntasks;
	const uint32_t *ntasks_dev;  // when non-null the task count is read from device memory (filter -> SW hand-off)
	uint32_t cross;          // 1: task t -> row chain rowlist[t / nseg], columns clist[(t % nseg)*W ...]
	const uint32_t *rowlist;
	uint32_t nseg;           // segments per row chain = ceil(ncols / W)
	uint32_t ncols;          // entries in clist (cross mode)
	const uint32_t *clist;   // column chain indices (cross: sorted by length; explicit: per task)
	const uint32_t *task_row, *task_begin, *task_cnt;  // explicit mode
	const uint32_t *cslot;   // explicit mode: record slot of each clist entry
	uint32_t a_begin, nB;    // cross mode: slot = (a - a_begin)*nB + b with (a,b) the reference indices
	// scratch (per warp of the grid)
	float4 *ckpt; uint64_t ckpt_stride;    // wavefront checkpoints, float4 units per warp
	unsigned long long *tile;              // re-computed trace strip: kSwStripSteps*32 words per warp
	float2 *bnd; uint64_t bnd_stride;      // pass-boundary rows (M, vertical gap) per column: per warp, one row per pass
	uint32_t bnd_pass_stride;
	uint8_t *stage; uint32_t stage_stride; // reversed path staging: per warp, one region of stage_chain_stride per chain
	uint32_t stage_chain_stride;
	float4 *best;                          // per warp kSwChainMax*32 parked (score, row, column) maxima
	// outputs
	PairRec *rec;
	uint8_t *pool; unsigned long long *pool_cursor;
	uint32_t *task_counter;
	const float *tables;   // weighted tables [2192]
	float open, ext;
};

struct LddtArgs {
	const uint32_t *lenA; const uint64_t *offA; const float *xA, *yA, *zA; const float *selfrevA;
	const uint32_t *lenB; const uint64_t *offB; const float *xB, *yB, *zB; const float *selfrevB;
	uint32_t npairs;
	uint32_t cross; uint32_t a_begin; uint32_t nB;
	const uint32_t *pair_a; const uint32_t *pair_b;  // explicit mode
	PairRec *rec;
	const uint8_t *pool;
	float min_fwd_score;
	uint32_t maxcols;
	float *scratch;  // null, or lddt_scratch_floats(maxcols) floats when the columns of an alignment do not fit shared memory
};

// K3: Mu int8 SW filter; same row/column task model as K1 (16 warps per CTA)
struct MuArgs {
	const uint8_t *mu_row; const uint64_t *off_row; const uint32_t *len_row;
	const uint8_t *mu_col; const uint64_t *off_col; const uint32_t *len_col;
	uint32_t tr;
	uint32_t ntasks;
	uint32_t cross;
	const uint32_t *rowlist; uint32_t nseg; uint32_t ncols;
	const uint32_t *clist;
	const uint32_t *task_row, *task_begin, *task_cnt, *cslot;  // explicit mode
	uint32_t a_begin, nB;
	int2 *bnd; uint32_t bnd_stride;
	PairRec *rec;
	uint8_t *keep;              // [batch pairs] 1 = passes the filter
	uint32_t *task_counter;
	uint32_t *sat_counter;
	const int *mu_mx;           // IntScoreMx_Mu as int32 [36*36]
	int open, ext;
	float omega, omega_fwd;
	uint32_t mkfl;              // pairs with a chain >= mkfl belong to the k-mer/x-drop path (dssaligner.cpp:715-732)
};

// survivor compaction: keep flags -> per-class SW task lists (explicit-mode arrays of SwArgs)
struct CompactArgs {
	uint32_t tr, a_begin, nB;
	const uint32_t *rowlist; uint32_t ncols;
	const uint32_t *run_row, *run_begin, *run_cnt;  // explicit pair lists: runs of pairs with the same row chain (compact_explicit_kernel)
	const uint32_t *clist;
	const uint8_t *keep;
	const uint32_t *len_row; const uint32_t *len_col;
	uint32_t *out_clist, *out_cslot;   // [nrows * ncols]
	uint32_t *task_row, *task_begin, *task_cnt;  // [kSwClasses * task_cap]
	uint32_t task_cap;
	uint32_t *task_count;              // [kSwClasses] device counters
	unsigned long long *pair_count;
	unsigned long long *cell_count;
};

// K4: long-chain path (MKF seeds -> chain -> x-drop), see mkf_kernel.cu
struct MkfSeed { uint32_t valid, lo_a, lo_b; int32_t best_hsp, best_chain; };
struct MkfXdrop { float score; uint32_t path_len; unsigned long long stage_off; };
// bytes of scratch one banded DP of la rows x lb columns needs (2 float rows, path staging, trace matrix)
__host__ __device__ inline size_t xdrop_region_bytes(uint32_t la, uint32_t lb)
{
	const size_t rows = (size_t)2 * (lb + 4) * sizeof(float);
	const size_t stage = ((size_t)la + lb + 4 + 15) & ~(size_t)15;
	const size_t tb = (size_t)(la + 3) * (lb + 3);
	return (rows + stage + tb + 15) & ~(size_t)15;
}
// upper bound for the backward + forward regions of one pair, whatever the seed position
__host__ __device__ inline size_t xdrop_pair_bytes(uint32_t LA, uint32_t LB)
{
	return (size_t)(LA + 6) * (LB + 6) + (size_t)8 * (LB + 8) + (size_t)LA + LB + 38 + 64;
}
struct MkfArgs {
	// reference orientation: A = query slot (hash table side), B = target slot
	const uint8_t *muA; const uint64_t *profA; const uint64_t *offA; const uint32_t *lenA;
	const uint8_t *muB; const uint64_t *profB; const uint64_t *offB; const uint32_t *lenB;
	uint32_t npairs;
	const uint32_t *pair_a, *pair_b, *pair_slot, *pair_hash;  // [npairs]; pair_hash = index of the A chain's hash table
	const uint32_t *hash_chain;  // [nhash] A chain of each hash table
	uint16_t *hash;
	MkfSeed *seeds;              // [npairs]
	MkfXdrop *xres;              // [2*npairs]
	uint32_t *xwork;             // [2*npairs] work list of the x-drop kernel (valid items, longest first)
	uint32_t *xcnt;              // [9] items per size bin, fill cursors, work-list cursor of the x-drop kernel
	unsigned char *scratch; const unsigned long long *scratch_off;  // per pair
	PairRec *rec;
	uint8_t *pool; unsigned long long *pool_cursor;
	const int *mu_mx; const float *tables;
	int x1, min_hsp_score; float x2, min_mega_hsp_score;
	float open, ext;
};
int launch_mkf(const MkfArgs &args, uint32_t nhash, int xgrid_blocks, cudaStream_t stream);
size_t mkf_hash_bytes();

// K5: gapless Mu pre-scores
struct GaplessArgs {
	const uint8_t *muA; const uint64_t *offA; const uint32_t *lenA;
	const uint8_t *muB; const uint64_t *offB; const uint32_t *lenB;
	uint32_t npairs; const uint32_t *pair_a, *pair_b;
	const float *mu_f32; const int *mu_i32;
	float *out_f; int *out_i;
};
int launch_mu_gapless(const GaplessArgs &args, cudaStream_t stream);

// K9: global alignment of explicit pairs (-global), see global_kernel.cu
struct GlobalRec { float score; uint32_t path_len; };
struct GlobalArgs {
	const uint64_t *profA; const uint64_t *offA; const uint32_t *lenA;
	const uint64_t *profB; const uint64_t *offB; const uint32_t *lenB;
	uint32_t npairs;
	const uint32_t *pair_a, *pair_b;   // [npairs]
	const uint32_t *order;             // [npairs] work order (largest matrices first)
	const uint8_t *skip;               // [npairs] 1 = rejected by the Mu filter, or null
	const unsigned long long *path_off;  // [npairs] slot of LA+LB bytes in pool
	char *pool;
	GlobalRec *rec;                    // [npairs]
	uint8_t *tb; size_t tb_stride;     // per-warp trace matrix (bytes; stride a multiple of 4)
	float *bnd; uint32_t bnd_stride;   // per-warp 2 x bnd_stride floats
	unsigned int *counter;
	const float *tables;
};
size_t global_tb_bytes(uint32_t maxLA, uint32_t maxLB);
int global_warps_per_block();
int launch_global(const GlobalArgs &args, int blocks, cudaStream_t stream);

// K6..K8: `-fast -db` 5-mer prefilter, see prefilter_kernel.cu
struct PfArgs {
	const int *kmer_mx;          // Mu_S_ij_i8 widened to int32 [36*36]
	const uint8_t *nb_tab;       // [36][136] per letter: letters by descending score, those scores, ge counts (pf_neighborhood_kernel)
	// query side (letters already K/L-swapped when the caller asks for it)
	uint32_t nQ; uint32_t nqk;   // queries, total 5-mer slots (sum of max(L-6, 0))
	const uint8_t *muQ; const uint64_t *offQ; const uint32_t *lenQ;
	const uint32_t *qk_off;      // [nQ+1] first 5-mer slot of each query
	const uint2 *qinfo; uint32_t sum_lenQ;  // [nQ] (length, residues of the queries before it); total query residues
	uint32_t *qk_code, *qk_val;  // [nqk] code (0xffffffff = masked), value = query<<16 | position
	uint32_t *nb_count; const unsigned long long *nb_off;  // [nqk] index entries contributed by each query 5-mer
	uint32_t *ix_key, *ix_val;   // index entries (unsorted while filling; ix_val sorted by key when probing)
	uint32_t exact_twice;        // query-neighbourhood mode: the k-mer itself is entered a second time
	const uint2 *row;            // [36^5] dense row table over the sorted index: (first entry, one past the last)
	uint32_t queue_cap;          // 0, or a smaller two-hit queue for the fused kernel (tests: RSK_PF_QUEUE)
	uint32_t *fuse_counter;      // zeroed target counter of the staged fused kernel (persistent CTAs), or null
	uint32_t no_stage;           // tests: RSK_PF_NOSTAGE keeps the letters in global memory
	uint32_t diag_safe;          // longest query + longest target <= 16384: no diagonal is dropped (prefiltermu.cpp:254)
	// target side
	uint32_t t_begin;            // first target of the batch
	const uint8_t *muT; const uint64_t *offT; const uint32_t *lenT;
	unsigned long long *hit_count; const unsigned long long *hit_off;  // per target of the batch ([ntl], [ntl+1])
	uint32_t *hit_key; const uint32_t *hit_sorted;
	unsigned *best;              // [ntl * nQ] best two-hit diagonal score per (target, query), 0 = none
	uint32_t *cand_count; const unsigned long long *cand_off;
	// (target, query, score) triples with a two-hit diagonal, accumulated over the target batches in stream order
	uint32_t t_base;             // index of this rank's first target in the whole DB
	unsigned long long raw_base; // triples written by the earlier batches
	uint32_t *raw_q; unsigned long long *raw_v;  // query, target<<16 | score
};
int pf_launch_swap_kl(const uint8_t *in, uint8_t *out, uint64_t n, cudaStream_t st);
int pf_launch_query_kmers(const PfArgs &a, cudaStream_t st);
int pf_launch_neighborhood(const PfArgs &a, bool fill, cudaStream_t st);
int pf_launch_mark_rows(const uint32_t *key, unsigned long long n, uint2 *row, cudaStream_t st);
int pf_launch_probe(const PfArgs &a, uint32_t ntl, bool fill, cudaStream_t st);
int pf_launch_extend(const PfArgs &a, uint32_t ntl, cudaStream_t st);
// K7+K8 in shared memory: which & 1 = targets with sum(LQ) + nQ * (LT - 1) <= pf_fuse_max_bits(0), which & 2 = those up to
// pf_fuse_max_bits(1); the others are left to pf_launch_probe / pf_launch_extend
int pf_launch_probe_extend(const PfArgs &a, uint32_t ntl, int which, cudaStream_t st);
unsigned long long pf_fuse_max_bits(int size);
int pf_launch_cands(const PfArgs &a, uint32_t ntl, bool write, cudaStream_t st);
size_t pf_bag_smem_bytes(uint32_t B);
int pf_sort_by_query(const uint32_t *qin, uint32_t *qout, const unsigned long long *vin, unsigned long long *vout, unsigned long long n,
		void *tmp, size_t &tmp_bytes, cudaStream_t st);
int pf_sort_keys64(const unsigned long long *kin, unsigned long long *kout, unsigned long long n, void *tmp, size_t &tmp_bytes, cudaStream_t st);
int pf_launch_mark_segments(const uint32_t *key, unsigned long long n, unsigned long long *seg_begin, unsigned long long *seg_end, cudaStream_t st);
int pf_launch_bag(const unsigned long long *val, const unsigned long long *seg_begin, const unsigned long long *seg_end, uint32_t nQ, uint32_t B,
		unsigned long long *out_key, uint32_t *out_n, cudaStream_t st);
int pf_launch_bag_compact(const unsigned long long *key, const uint32_t *out_n, const unsigned long long *out_off, uint32_t nQ, uint32_t B,
		unsigned long long *dense, cudaStream_t st);
int pf_launch_unpack_triples(const uint32_t *q, const unsigned long long *v, unsigned long long n, uint32_t *t_out, uint32_t *q_out,
		uint16_t *s_out, cudaStream_t st);
int pf_launch_unpack_keys(const unsigned long long *key, unsigned long long n, uint32_t *t_out, uint32_t *q_out, uint16_t *s_out, cudaStream_t st);
int pf_sort_pairs(const uint32_t *kin, uint32_t *kout, const uint32_t *vin, uint32_t *vout, unsigned long long n, void *tmp,
		size_t &tmp_bytes, cudaStream_t st);
int pf_segmented_sort(const uint32_t *kin, uint32_t *kout, unsigned long long n, uint32_t nseg, const unsigned long long *off,
		void *tmp, size_t &tmp_bytes, cudaStream_t st);

// K10: DSS feature extraction, see dss_kernel.cu
struct DssTables { double conf[16][9]; double bins[5][15]; unsigned char amino[256]; };  // bins: NENDist, RENDist, DstNxtHlx, StrandDens, NormDens
struct DssArgs {
	uint32_t n; uint64_t total;
	const uint32_t *len; const uint64_t *off; const float *x, *y, *z;
	const uint8_t *aa_char;    // [total] amino-acid characters, or null when ...
	const uint64_t *aa_prof8;  // ... the AA letters come from byte 0 of an existing set's packed profile
	uint32_t reverse;          // 1: every chain is read back to front (PDBChain::GetReverse, pdbchain.cpp:478)
	uint8_t *ss, *conf; double *dens; uint32_t *helix_mid;  // per-residue scratch [total]
	uint8_t *planes;           // out [8][total]
	uint8_t *mu;               // out [total] or null
	const DssTables *tab;
};
int launch_dss(const DssArgs &a, int grid, cudaStream_t st);
int launch_dss_unpack(const uint64_t *prof8, uint64_t total, uint8_t *planes, cudaStream_t st);

// kernel launchers (each returns the number of kernels it launched, or <0 on error)
int launch_sw(const SwArgs &args, int cls, int grid, cudaStream_t stream);
size_t sw_smem_bytes();
uint64_t sw_ckpt_units(int npass, uint64_t LB);
int launch_lddt(const LddtArgs &args, cudaStream_t stream);
size_t lddt_scratch_floats(uint32_t maxcols);
int launch_mu_filter(const MuArgs &args, int grid, cudaStream_t stream);
int launch_mu_filter16(const MuArgs &args, int grid, cudaStream_t stream);
size_t mu_smem_bytes();
int launch_compact_survivors(const CompactArgs &args, uint32_t nrows, cudaStream_t stream);
int launch_compact_explicit(const CompactArgs &args, uint32_t nruns, cudaStream_t stream);
int launch_pack_profiles(const uint8_t *planes, uint64_t total, uint64_t *prof8, cudaStream_t stream);

}  // namespace rsk
