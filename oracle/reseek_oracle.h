/* reseek_oracle.h - CPU restatement of the Reseek per-pair search hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity checker for the CUDA library in reseek_b200/csrc.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (libreseek_b200.so) never links, loads or calls anything in oracle/.
 *
 * Parity status: PINNED.  Every function below is checked in tests/test_oracle_vs_reference.py against
 * the unmodified reference compiled from /root/reference/src (oracle/_ref/libreseek_ref.so, strict IEEE
 * flags, recipe in oracle/Makefile) and against the committed fixtures in tests/golden/ that the same
 * reference build produced (tools/make_golden.py).
 *
 * Each function cites the reference file:line (under /root/reference/src) whose behaviour it restates.
 * Plain C11, compile with -O2 -ffp-contract=off (no -ffast-math): float results are bit-exact targets.
 */
#ifndef RESEEK_ORACLE_H
#define RESEEK_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NFEAT 8
#define ORC_TABLE_FLOATS 2192 /* 20*20 + 7*16*16 */
#define ORC_MU_ALPHA 36

/* Mirror of the DSSParams scalars the hot path reads (dssparams.h:29-68). */
typedef struct orc_params {
	float gap_open;   /* <= 0, namedparams.cpp:45 */
	float gap_ext;    /* <= 0, namedparams.cpp:46 */
	float min_fwd_score; /* dssparams.cpp:80, namedparams.cpp:48 */
	float omega;      /* Mu filter threshold, dssparams.cpp:52-81 */
	float omega_fwd;
	int mu_gap_open;  /* dssparams.h:45 */
	int mu_gap_ext;   /* dssparams.h:46 */
	uint32_t mkfl;    /* MKF length threshold */
	int mkf_x1, mkf_x2, mkf_min_hsp_score;
	float mkf_min_mega_hsp_score;
	float weights[ORC_NFEAT];
	/* weighted tables: feature f at offset orc_feat_offset(f), alpha x alpha row-major (dssparams.cpp:344-364) */
	float tables[ORC_TABLE_FLOATS];
} orc_params;

enum { ORC_MODE_FAST = 1, ORC_MODE_SENSITIVE = 2, ORC_MODE_VERYSENSITIVE = 3 };

/* dssparams.cpp:44-111 (SetDSSParams presets) + namedparams.cpp:32-53 (SetDefaults) + ApplyWeights */
int orc_params_preset(orc_params *p, int mode);
int orc_feat_offset(int f);
int orc_feat_alpha(int f);
/* raw data accessors (trained parameters; used by the synthetic-data generator) */
const float *orc_bgfreq(int f); /* background letter frequencies, trained_features.cpp X_f_i */
const signed char *orc_mu_i8(void);      /* IntScoreMx_Mu mumx_data.cpp:42 */
const signed char *orc_mu_kmer_i8(void); /* Mu_S_ij_i8 mumx_data.cpp:81 */
const float *orc_mu_f32(void);           /* ScoreMx_Mu mumx_data.cpp:3 */

/* One chain as the aligner sees it (dssaligner.h:109-119 SetQuery/SetTarget arguments). */
typedef struct orc_chain {
	uint32_t L;
	const uint8_t *prof; /* [ORC_NFEAT][L] feature letters (dss.cpp:716 GetProfile) */
	const uint8_t *mu;   /* [L] Mu letters 0..35, may be NULL (dss.cpp:700) */
	const float *x, *y, *z; /* [L] C-alpha coordinates (pdbchain.h:13-17) */
	float selfrev;       /* self-reverse score; FLT_MAX = unset (alignpair.cpp:7-25) */
} orc_chain;

/* Result of one pair; mirrors the DSSAligner public members (dssaligner.h:46-72). */
typedef struct orc_result {
	float score;          /* m_AlnFwdScore */
	uint32_t lo_a, lo_b;  /* m_LoA, m_LoB (UINT32_MAX if no alignment) */
	uint32_t hi_a, hi_b;  /* m_HiA, m_HiB */
	uint32_t ids, gaps;   /* M count, D+I count */
	float lddt;
	float ts;             /* m_NewTestStatisticA */
	float pvalue, evalue, qual; /* (float) casts as in dssaligner.cpp:891-900; FLT_MAX if unset */
	float mu_score;       /* Mu filter score fwd-rev (0 if not run) */
	int32_t mu_fwd, mu_rev;
	int32_t filtered;     /* 1 = rejected by the Mu filter (no SW run) */
	uint32_t path_len;    /* 0 = no alignment */
} orc_result;

/* sw.cpp:79-212 SWFast + sw.cpp:8-77 TraceBackBitSW over the on-the-fly 8-feature score
 * (dssaligner.cpp:529-596 SetSMx_NoRev summation order).  path must hold LA+LB+1 bytes; returns score. */
float orc_sw_align(const orc_params *p, const uint8_t *profA, uint32_t LA, const uint8_t *profB, uint32_t LB,
		uint32_t *lo_a, uint32_t *lo_b, char *path, uint32_t *path_len);

/* same DP over an explicit LA x LB float score matrix (row-major) - sw.cpp:79-212 verbatim semantics */
float orc_swfast_matrix(const float *S, uint32_t LA, uint32_t LB, float open, float ext,
		uint32_t *lo_a, uint32_t *lo_b, char *path, uint32_t *path_len);

/* dssaligner.cpp:557-595: S[i][j] in fp32, feature 0 first then += features 1..7 */
float orc_cell_score(const orc_params *p, const uint8_t *profA, uint32_t LA, uint32_t i,
		const uint8_t *profB, uint32_t LB, uint32_t j);

/* Score-only local affine SW on Mu letters; parasail.cpp:515-797 restated as scalar Gotoh with all values
 * floored at 0.  Returns best H; *saturated = (best > 250) (parasail.cpp:585,728-733). */
int orc_mu_sw_score(const uint8_t *a, uint32_t LA, const uint8_t *b, uint32_t LB, int open, int ext, int *saturated);

/* parasail_mu.cpp:120-161 AlignMuQP_Para: fwd (777 when saturated), early 0 if fwd < omega_fwd,
 * rev on reversed A (255 when saturated), score = fwd - rev. */
float orc_mu_filter_score(const orc_params *p, const uint8_t *a, uint32_t LA, const uint8_t *b, uint32_t LB,
		int *fwd, int *rev);

/* lddt.cpp:63-124 GetLDDT_mu_fast over aligned column pairs */
float orc_lddt(const orc_chain *A, const orc_chain *B, const uint32_t *posA, const uint32_t *posB, uint32_t n);

/* statsig.cpp:27-50, statsig.h:8-23 */
double orc_pvalue(double ts);
double orc_evalue(double ts);
double orc_qual(double ts);

/* dssaligner.cpp:852-904 CalcEvalue (+GetPathCounts, GetPosABs :1282, GetLDDT :1313) given score/lo/path in r */
void orc_calc_evalue(const orc_params *p, const orc_chain *A, const orc_chain *B, const char *path, orc_result *r);

/* dssaligner.cpp:793-831 AlignQueryTarget for pairs below the MKF length threshold:
 * ClearAlign -> (omega>0 && mu present: MuFilter :619) -> Align_NoAccel :929.  path: LA+LB+1 bytes. */
void orc_align_pair(const orc_params *p, const orc_chain *A, const orc_chain *B, orc_result *r, char *path);

/* -global: ViterbiFastMem (viterbifastmem.cpp:33-193) + TraceBackBitMem (tracebackbitmem.cpp:8-69); path needs LA+LB+1 bytes */
float orc_viterbi_global(const orc_params *p, const uint8_t *profA, uint32_t LA, const uint8_t *profB, uint32_t LB,
		char *path, uint32_t *path_len);
/* DSSAligner::AlignQueryTarget_Global (global.cpp:7-33): r->score = m_GlobalScore, lo_a = lo_b = 0 */
void orc_align_pair_global(const orc_params *p, const orc_chain *A, const orc_chain *B, orc_result *r, char *path);

/* ---- alternative gapless Mu pre-scores (not reachable from the reference CLI) ---- */
/* swgaplessprofb.cpp:6-61 SWFastGaplessProfb: best gapless local run on ScoreMx_Mu, forward minus reversed-A */
float orc_mu_gapless_profb(const uint8_t *a, uint32_t LA, const uint8_t *b, uint32_t LB);
/* swfastpinopgapless.cpp:6-47 SWFastPinopGapless on IntScoreMx_Mu rows: best gapless local run, integer */
int orc_mu_gapless_int(const uint8_t *a, uint32_t LA, const uint8_t *b, uint32_t LB);

/* ---- long-chain path: Mu k-mer filter (MKF) + chaining + banded x-drop ---- */
typedef struct orc_hsp { int loi, loj, len, score; } orc_hsp;

/* mukmerfilter.cpp:105-175 MuXDrop: ungapped x-drop extension of a 3-mer seed on IntScoreMx_Mu */
int orc_mu_xdrop(const uint8_t *Q, int LQ, const uint8_t *T, int LT, int PosQ, int PosT, int X, int *Loi, int *Loj, int *Len);

/* mukmerfilter.cpp:208-230 SetHashTable + :316-389 Align + :391-464 ChainHSPs + chainer.cpp:31-194.
 * hsps: the kept HSP list in discovery order (cap entries); chain_idx: indices into hsps, chain end first. */
int orc_mkf_align(const orc_params *p, const uint8_t *muQ, int LQ, const uint8_t *muT, int LT,
		orc_hsp *hsps, int cap, int *nhsp, int *best_hsp_score, int *best_chain_score, int *chain_idx, int *nchain);

/* xdropfwd.cpp:71-386 (forward) / xdropbwd.cpp:28-50 (reverse != 0: coordinates mirrored as RevSubFn does).
 * Aligns A[LoA..aLA) x B[LoB..aLB) starting exactly at (LoA,LoB); path gets the forward-order M/D/I string of the
 * DP's own orientation (the caller reverses it for the backward pass, xdropbwd.cpp:48). */
float orc_xdrop_fwd(const orc_params *p, const uint8_t *profA, uint32_t LAfull, const uint8_t *profB, uint32_t LBfull,
		int reverse, uint32_t revLA, uint32_t revLB, float X, uint32_t LoA, uint32_t aLA, uint32_t LoB, uint32_t aLB,
		char *path, uint32_t *path_len);

/* dssaligner.cpp:488-527 GetMegaHSPScore: feature-major accumulation */
float orc_mega_hsp_score(const orc_params *p, const uint8_t *profA, uint32_t LA, const uint8_t *profB, uint32_t LB,
		uint32_t lo_i, uint32_t lo_j, uint32_t len);

/* dssaligner.cpp:1387-1437 AlignMKF + PostAlignMKF, xdrophsp.cpp:42-150 XDropHSP, mergefwdback.cpp:6-50 */
void orc_align_mkf(const orc_params *p, const orc_chain *A, const orc_chain *B, orc_result *r, char *path,
		int *best_hsp_score, int *best_chain_score, float *xdrop_score);

/* dssaligner.cpp:715-732 DoMKF (k-mers exist iff the chain has >= 3 residues) */
int orc_do_mkf(const orc_params *p, const orc_chain *A, const orc_chain *B);

/* ---- -fast -db prefilter ---- */
/* Spaced 5-mer over offsets {0,1,2,5,6} of a 7-window, base-36 big-endian (mudex.cpp:517-538); self score on
 * Mu_S_ij_i8 (mermx.cpp:725); k-mers with self score < 36 are masked (UINT32_MAX). */
uint32_t orc_kmer5(const uint8_t *window7);
int orc_kmer5_pair_score(uint32_t k1, uint32_t k2);
/* prefiltermu.cpp:12-48 FindHSP: best ungapped segment on a whole diagonal d = LQ + j - i - 1 (diag.h:22-25) */
int orc_find_hsp(const uint8_t *q, uint32_t LQ, const uint8_t *t, uint32_t LT, int diag);
/* PrefilterMu::Search for one target (prefiltermu.cpp:382-393), stated without the index: a seed is a pair of
 * unmasked 5-mers scoring >= 36; in query-neighbourhood mode an exact seed counts twice (the k-mer is indexed as
 * itself and as a member of its own neighbourhood, mudex.cpp:146-174); a (query, diagonal) with >= 2 seeds is
 * extended with FindHSP; best[q] = max over its diagonals, clamped to 65534, 0 = no candidate.
 * Query letters must already carry the K/L swap of the reference. */
void orc_prefilter_target(const uint8_t *const *muQ, const uint32_t *LQ, uint32_t nQ, const uint8_t *muT, uint32_t LT,
		int query_neighborhood, uint16_t *best);
/* RankedScoresBag (rankedscoresbag.cpp:5-51): per query keep the top-B targets, lazy truncation at 2B with the
 * reference's own quicksort (sort.h:71-108), admission score >= lo.  Feed scores in target order. */
typedef struct orc_rsb orc_rsb;
orc_rsb *orc_rsb_new(uint32_t nQ, uint32_t B);
void orc_rsb_add(orc_rsb *r, uint32_t q, uint32_t t, uint16_t score);
void orc_rsb_finish(orc_rsb *r); /* the final TruncateVecs of ToTsv (rankedscoresbag.cpp:190-194) */
uint32_t orc_rsb_count(const orc_rsb *r, uint32_t q);
const uint32_t *orc_rsb_targets(const orc_rsb *r, uint32_t q);
const uint16_t *orc_rsb_scores(const orc_rsb *r, uint32_t q);
void orc_rsb_free(orc_rsb *r);

/* Batch helpers for the CPU baseline timing (scalar port, one thread). */
void orc_align_pairs(const orc_params *p, const orc_chain *chainsA, const orc_chain *chainsB,
		const uint32_t *ia, const uint32_t *ib, size_t npairs, orc_result *out);

#ifdef __cplusplus
}
#endif
#endif
