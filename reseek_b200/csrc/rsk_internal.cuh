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
	uint4 *coloff = nullptr;    // [2*total] per-residue row-table offsets (made on first use as the column side)
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

// ---- SW kernel geometry ----
constexpr int kSwWarps = 16;                 // warps per CTA: each aligns one B chain against the CTA's A chain
constexpr int kSwThreads = kSwWarps * 32;
constexpr int kMaxRowsPerLane = 8;           // R: DP rows owned by one lane within a pass
constexpr int kRowsPerPassMax = 32 * kMaxRowsPerLane;

// How a chain of LA rows is cut into passes of 32*R rows (R rows per lane).
__host__ __device__ inline void sw_geometry(uint32_t LA, int &npass, int &R)
{
	npass = (int)((LA + kRowsPerPassMax - 1) / kRowsPerPassMax);
	if (npass < 1) npass = 1;
	R = (int)((LA + 32u * npass - 1) / (32u * npass));
	if (R < 1) R = 1;
}

struct SwArgs {
	// chains
	const uint64_t *profA; const uint64_t *offA; const uint32_t *lenA;
	const uint4 *coloffB; const uint64_t *offB; const uint32_t *lenB;
	// tasks: one task = one A chain x up to kSwWarps B chains
	uint32_t ntasks;
	uint32_t cross;          // 1: task t -> a = a_begin + t / nseg, B's = blist[(t % nseg)*kSwWarps ...]
	uint32_t a_begin;        // first A index of this batch (cross mode)
	uint32_t nseg;           // segments per A chain (cross mode)
	uint32_t nB;             // entries in blist (cross mode)
	const uint32_t *task_a;      // explicit mode [ntasks]
	const uint32_t *task_begin;  // explicit mode [ntasks] offset into blist/bslot
	const uint32_t *task_cnt;    // explicit mode [ntasks] 1..kSwWarps
	const uint32_t *blist;       // B chain indices
	const uint32_t *bslot;       // explicit mode: record slot of each blist entry; cross: slot = (a-a_begin)*nB + b
	const uint32_t *ntasks_dev;  // when non-null the task count is read from device memory (filter -> SW hand-off)
	// scratch (per warp of the grid)
	uint4 *trace; uint64_t trace_stride;   // uint4 units per warp
	float2 *bnd; uint32_t bnd_stride;      // pass-boundary row (M, D) per column
	uint8_t *stage; uint32_t stage_stride; // reversed path staging
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
};

// K3: Mu int8 SW filter over the full A x B rectangle of a batch
struct MuArgs {
	const uint8_t *muA; const uint64_t *offA; const uint32_t *lenA;
	const uint8_t *muB; const uint64_t *offB; const uint32_t *lenB;
	uint32_t ntasks, a_begin, nseg, nB;
	uint32_t cross;             // 1: rectangle tasks (see SwArgs); 0: explicit task arrays below
	const uint32_t *task_a, *task_begin, *task_cnt, *bslot;
	const uint32_t *blist;      // B indices sorted by length
	int2 *bnd; uint32_t bnd_stride;
	PairRec *rec;
	uint8_t *keep;              // [batch pairs] 1 = passes the filter
	uint32_t *task_counter;
	uint32_t *sat_counter;
	const int *mu_mx;           // IntScoreMx_Mu as int32 [36*36]
	int open, ext;
	float omega, omega_fwd;
	uint32_t mkfl;              // pairs with LA >= mkfl or LB >= mkfl belong to the k-mer/x-drop path (dssaligner.cpp:715-732)
};

// survivor compaction: keep flags -> SW tasks (explicit-mode arrays of SwArgs)
struct CompactArgs {
	uint32_t a_begin, nB;
	const uint32_t *blist;
	const uint8_t *keep;
	const uint32_t *lenA; const uint32_t *lenB;
	uint32_t *out_blist, *out_bslot;   // [nA_batch * nB]
	uint32_t *task_a, *task_begin, *task_cnt;
	uint32_t *task_count;              // device counters
	unsigned long long *pair_count;
	unsigned long long *cell_count;
};

// kernel launchers (each returns the number of kernels it launched, or <0 on error)
int launch_sw(const SwArgs &args, int grid, size_t smem, cudaStream_t stream);
size_t sw_smem_bytes();
uint64_t sw_trace_units(int npass, uint32_t LB);
int launch_make_coloff(const uint64_t *prof8, uint64_t total, uint4 *coloff, cudaStream_t stream);
int launch_lddt(const LddtArgs &args, cudaStream_t stream);
int launch_mu_filter(const MuArgs &args, int grid, cudaStream_t stream);
size_t mu_smem_bytes();
int launch_compact_survivors(const CompactArgs &args, uint32_t nA, cudaStream_t stream);
int launch_pack_profiles(const uint8_t *planes, uint64_t total, uint64_t *prof8, cudaStream_t stream);

}  // namespace rsk
