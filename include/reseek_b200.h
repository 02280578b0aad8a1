/* reseek_b200.h - C ABI of the B200-native Reseek search hot path (libreseek_b200.so).
 *
 * The reference (rcedgar/reseek, C++) has no FFI; its boundary is the class surface its search drivers use:
 *   DSSAligner::SetParams/SetQuery/SetTarget/AlignQueryTarget/AlignBags + result members  dssaligner.h:106-223, :46-72
 *   DBSearcher::LoadDB/Setup/RunQuery/RunSelf/OnAln/Reject                                 dbsearcher.h:83-104
 *   MuPreFilter / PostMuFilter free functions                                              search.cpp:9-18
 * Each entry point below names the reference interface it replaces.  The C++ look-alikes that sit on top of
 * this ABI (same class and method names) live in reseek_b200/csrc/host/; INTEGRATION.md shows the binding a
 * reference maintainer would add.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function returns 0 on success or a
 * negative rsk_status, with a message available from rsk_last_error() (the reference's Die() -> exit(1),
 * myutils.cpp:785, becomes an error code; the C++ shim turns it back into Die).  There is NO CPU fallback:
 * if no CUDA device / sm_100 kernel image is available the call fails with RSK_ERR_CUDA.
 * All "host" pointers are ordinary (preferably pinned) host memory; "dev" pointers are device memory on the
 * context's GPU.  A context is bound to one GPU and one stream and must not be shared between threads
 * (same rule as one DSSAligner per thread, dbsearcher.cpp:98-106).
 */
#ifndef RESEEK_B200_H
#define RESEEK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSK_NFEAT 8            /* AA, NENDist, Conf, NENConf, RENDist, DstNxtHlx, StrandDens, NormDens (namedparams.cpp:36-43) */
#define RSK_TABLE_FLOATS 2192  /* 20*20 + 7*16*16 weighted log-odds entries */
#define RSK_NLETTERS 132       /* 20 + 7*16: one "row table" entry per (feature, letter) */
#define RSK_MU_ALPHA 36

typedef enum rsk_status {
	RSK_OK = 0,
	RSK_ERR_ARG = -1,     /* bad argument (the reference would asserta/Die) */
	RSK_ERR_CUDA = -2,    /* CUDA runtime error, no device, or no sm_100a kernel image */
	RSK_ERR_NOMEM = -3,
	RSK_ERR_LIMIT = -4    /* input exceeds a documented limit (e.g. chain length) */
} rsk_status;

enum { RSK_MODE_FAST = 1, RSK_MODE_SENSITIVE = 2, RSK_MODE_VERYSENSITIVE = 3 };

/* POD copy of the DSSParams scalars + weighted score tables the hot path reads (dssparams.h:29-68).
 * Replaces: DSSParams::SetDSSParams (dssparams.cpp:44-111), SetDefaults (namedparams.cpp:32-53),
 * ApplyWeights (dssparams.cpp:344-364). */
typedef struct rsk_params {
	float gap_open;              /* <= 0 */
	float gap_ext;               /* <= 0 */
	float min_fwd_score;         /* CalcEvalue skipped below this (dssaligner.cpp:861) */
	float omega;                 /* Mu filter threshold; <= 0 disables the filter */
	float omega_fwd;
	int32_t mu_gap_open;         /* int8 Mu SW penalties (dssparams.h:45-46) */
	int32_t mu_gap_ext;
	uint32_t mkfl;               /* chains >= mkfl take the k-mer/x-drop path (dssaligner.cpp:715-732) */
	int32_t mkf_x1, mkf_x2, mkf_min_hsp_score;
	float mkf_min_mega_hsp_score;
	double max_evalue;           /* DBSearcher::m_MaxEvalue (dbsearcher.cpp:75-83) */
	float weights[RSK_NFEAT];
	float tables[RSK_TABLE_FLOATS]; /* weighted: float(w) * float(logodds), feature after feature, row-major */
} rsk_params;

typedef struct rsk_ctx rsk_ctx;           /* one GPU + one stream + scratch; the "DSSAligner pool" of a DBSearcher */
typedef struct rsk_chainset rsk_chainset; /* device-resident chains: DBSearcher::m_DBChains/m_DBProfiles/... (dbsearcher.h:26-33) */
typedef struct rsk_results rsk_results;   /* hits of one search call */

/* Host-side structure-of-arrays view of a set of chains.  Layout mirrors what the reference holds per chain:
 * profile = vector<vector<byte>>[feature][pos] (dss.cpp:716), Mu letters vector<byte> (dss.cpp:700),
 * coordinates PDBChain::m_Xs/m_Ys/m_Zs (pdbchain.h:13-17), self-reverse score (alignpair.cpp:7-25). */
typedef struct rsk_chains_host {
	uint32_t n;            /* number of chains */
	uint64_t total;        /* sum of len[] */
	const uint32_t *len;   /* [n] chain lengths (>= 1) */
	const uint8_t *prof;   /* [RSK_NFEAT][total]: plane-major, chains concatenated in order */
	const uint8_t *mu;     /* [total] Mu letters 0..35, or NULL (then no Mu filter for these chains) */
	const float *xyz;      /* [3][total] x plane, y plane, z plane */
	const float *selfrev;  /* [n] self-reverse scores, or NULL (= FLT_MAX "unset", dssaligner.cpp:876-877) */
} rsk_chains_host;

/* One alignment, fixed width; mirrors the DSSAligner result members (dssaligner.h:46-72).
 * a/b index the A ("query" slot, rows) and B ("target" slot, columns) chain sets of the call. */
typedef struct rsk_hit {
	uint32_t a, b;
	float score;            /* m_AlnFwdScore (fp32 bits identical to the reference) */
	uint32_t lo_a, lo_b;    /* m_LoA, m_LoB (0-based) */
	uint32_t hi_a, hi_b;    /* m_HiA, m_HiB; UINT32_MAX when CalcEvalue was skipped */
	uint32_t ids, gaps;     /* m_Ids (M columns), m_Gaps (D+I columns) */
	float lddt;
	float ts;               /* m_NewTestStatisticA; -FLT_MAX when unset */
	float pvalue, evalue, qual; /* (float) of the double formulas, statsig.cpp:27-50; FLT_MAX when unset */
	float mu_score;         /* Mu filter score fwd - rev when the filter ran, else 0 */
	int32_t mu_fwd, mu_rev;
	uint32_t flags;         /* RSK_HIT_* */
	uint32_t path_len;      /* number of path columns (0 = no alignment) */
	uint64_t path_off;      /* offset of the path in the results' path pool */
} rsk_hit;

enum {
	RSK_HIT_MU_REJECTED = 1u, /* dropped by the Mu filter: no SW was run */
	RSK_HIT_HAS_EVALUE = 2u,  /* CalcEvalue ran (score >= min_fwd_score) */
	RSK_HIT_REPORTED = 4u,    /* passes DBSearcher::Reject (E <= max_evalue) */
	RSK_HIT_MKF = 8u,         /* DoMKF() pair (a chain >= mkfl, dssaligner.cpp:715-732): aligned by the k-mer / x-drop path
	                             (AlignMKF); mu_fwd = m_MKF.m_BestHSPScore, mu_rev = m_MKF.m_BestChainScore */
	RSK_HIT_GLOBAL = 16u      /* record of rsk_align_global: score = m_GlobalScore, path = m_GlobalPath over both whole chains */
};

/* which pairs come back from a search call */
enum {
	RSK_KEEP_HITS = 0,   /* only pairs the reference would emit: non-empty path and !Reject (runquery.cpp:72-73) */
	RSK_KEEP_ALL = 1     /* one record per scheduled pair, in schedule order (parity tests, OnAln subclasses) */
};

typedef struct rsk_search_opts {
	int32_t keep;           /* RSK_KEEP_* */
	int32_t want_paths;     /* 0: records only; 1: also the M/D/I path strings */
	int32_t skip_evalue;    /* 1: stop after SW+traceback (no LDDT/TS) - kernel benchmarking only */
	int32_t reserved;
} rsk_search_opts;

/* Work counters of the last search call (DSSAligner::Stats dssaligner.cpp:1088, DBSearcher::RunStats dbsearcher.cpp:29-56) */
typedef struct rsk_stats {
	uint64_t pairs;          /* scheduled pairs */
	uint64_t mu_filter_in;   /* pairs that went through the Mu filter */
	uint64_t mu_filter_rejected;
	uint64_t mu_saturated;   /* forward int8 saturations (m_ParasailSaturateCount) */
	uint64_t sw_pairs;       /* pairs that reached the float SW */
	uint64_t sw_cells;       /* sum LA*LB over those pairs */
	uint64_t evalue_pairs;   /* pairs for which LDDT/TS were computed */
	uint64_t hits;           /* pairs passing Reject */
	uint64_t kernel_launches;/* CUDA kernels of this library launched by the call */
	uint64_t h2d_bytes, d2h_bytes;
	float sw_kernel_ms;      /* device time of the SW kernel(s) (CUDA events on the context stream) */
	float mu_kernel_ms;
	float lddt_kernel_ms;
	float total_ms;          /* device time of the whole call on the context stream */
	float mkf_kernel_ms;     /* long-chain path kernels */
	uint32_t sw_kernel_launches; /* launches of the SW kernel (one per batch and row-length class) */
	uint64_t mkf_pairs;      /* pairs that took the k-mer / x-drop path (DoMKF) */
} rsk_stats;

/* ---- library ---- */
const char *rsk_version(void);
const char *rsk_last_error(void); /* thread-local message of the last failing call */
int rsk_device_count(void);       /* CUDA devices visible; 0 when there is none (no fallback exists) */

/* ---- parameters (DSSParams) ---- */
int rsk_params_preset(rsk_params *p, int mode);
/* raw trained data, for tools and tests */
int rsk_feature_alpha(int f);
int rsk_feature_offset(int f);            /* offset of feature f in tables[] */
const float *rsk_feature_bgfreq(int f);   /* background letter frequencies (trained_features.cpp X_f_i) */
const int8_t *rsk_mu_matrix_i8(void);     /* IntScoreMx_Mu [36][36] (mumx_data.cpp:42) */
const int8_t *rsk_mu_kmer_matrix_i8(void);/* Mu_S_ij_i8 [36][36] (mumx_data.cpp:81) */
const float *rsk_mu_matrix_f32(void);     /* ScoreMx_Mu [36][36] (mumx_data.cpp:3) */

/* ---- context: DBSearcher::Setup (dbsearcher.cpp:73) + DSSAligner::SetParams (dssaligner.h:106) ---- */
int rsk_ctx_create(int device, const rsk_params *params, void *cuda_stream /* cudaStream_t or NULL */, rsk_ctx **out);
int rsk_ctx_set_params(rsk_ctx *ctx, const rsk_params *params);
int rsk_ctx_get_params(const rsk_ctx *ctx, rsk_params *out);
void rsk_ctx_destroy(rsk_ctx *ctx);
int rsk_ctx_stats(const rsk_ctx *ctx, rsk_stats *out);
int rsk_ctx_sync(rsk_ctx *ctx);

/* ---- chains: DBSearcher::LoadDB / the per-chain vectors (dbsearcher.cpp:242, dbsearcher.h:26-33) ---- */
int rsk_chainset_upload(rsk_ctx *ctx, const rsk_chains_host *chains, rsk_chainset **out);
/* Same, without waiting for the copies: they are queued on the context stream ahead of any later search call.  Arrays in
 * pageable memory have been read when the call returns (they are staged through the context's own pinned buffers, worker
 * threads filling one buffer while the other is on the bus); arrays in PINNED memory are read by the DMA itself and must stay
 * untouched until rsk_ctx_sync() or the next call on this context that returns results.  The streamed side of RunQuery
 * (runquery.cpp:82-125: the next block is read while the current one is aligned) uses this. */
int rsk_chainset_upload_async(rsk_ctx *ctx, const rsk_chains_host *chains, rsk_chainset **out);
uint32_t rsk_chainset_count(const rsk_chainset *cs);
uint64_t rsk_chainset_residues(const rsk_chainset *cs);
void rsk_chainset_free(rsk_chainset *cs);

/* ---- DSS on the device: structure -> feature letters ----
 * rsk_chainset_from_coords replaces, for every chain, DSS::Init + GetProfile + GetMuLetters (dss.cpp:716-741, 700-714 and what
 * they call: getss.cpp:6-63, myss.cpp:142-210, dss.cpp:179-244, 339-440, 78-155, 866-881, valuetoint.cpp) as ProfileLoader's
 * threads run them (profileloader.cpp:17-70): the C-alpha coordinates and amino-acid characters go up (13 bytes per residue),
 * the 8 feature planes and the Mu letters are computed by one CTA per chain and stay on the device.  Letter-exact: same types
 * and operation order as the reference, exp() with the bits of the host's libm.  Self-reverse scores start "unset"
 * (FLT_MAX); rsk_chainset_reversed + rsk_chainset_selfrev fill them. */
typedef struct rsk_coords_host {
	uint32_t n;            /* number of chains */
	uint64_t total;        /* sum of len[] */
	const uint32_t *len;   /* [n] */
	const char *aa;        /* [total] amino-acid characters (PDBChain::m_Seq), chains concatenated */
	const float *xyz;      /* [3][total] */
} rsk_coords_host;
int rsk_chainset_from_coords(rsk_ctx *ctx, const rsk_coords_host *chains, int with_mu, rsk_chainset **out);
/* PDBChain::GetReverse (pdbchain.cpp:478) + DSS of every reversed chain of S, carrying S's FORWARD Mu letters (alignpair.cpp:22):
 * the Srev argument of rsk_chainset_selfrev, made on the device. */
int rsk_chainset_reversed(rsk_ctx *ctx, const rsk_chainset *S, rsk_chainset **out);
/* feature letters of a device chain set back on the host (hit writers, OnAln subclasses, tests): prof [RSK_NFEAT][total]
 * plane-major, mu [total], selfrev [n]; any pointer may be NULL */
int rsk_chainset_download_features(rsk_ctx *ctx, const rsk_chainset *S, uint8_t *prof, uint8_t *mu, float *selfrev);

/* Self-reverse scores: GetSelfRevScore (alignpair.cpp:7-25) for every chain of S.  Srev holds, chain by chain, the
 * profile of the coordinate-reversed chain (PDBChain::GetReverse + DSS, upstream of this library) together with the
 * FORWARD Mu letters (the reference passes the forward letters for the reversed chain, alignpair.cpp:22).  Chain i of S is
 * aligned to chain i of Srev with the context's current parameters (ProfileLoader uses omega = 0, profileloader.cpp:22-26;
 * RunQuery the full search parameters, runquery.cpp:43); the scores are stored in S (host + device) and optionally copied out. */
int rsk_chainset_selfrev(rsk_ctx *ctx, rsk_chainset *S, const rsk_chainset *Srev, float *scores_out);

/* ---- the per-pair hot loop ----
 * rsk_search_cross: every chain of A (the streamed "-db" side, DSSAligner query slot) against every chain of
 *   B (the in-memory side, target slot): DBSearcher::RunQuery / ThreadBodyQuery (runquery.cpp:18-80).
 * rsk_search_self:  pairs i <= j of one set, A = chain i, B = chain j: DBSearcher::RunSelf (runself.cpp:72-145).
 * rsk_search_pairs: explicit (ia[k], ib[k]) list: PostMuFilter's per-line AlignBags loop (postmufilter.cpp:185-195),
 *   -alignpair (alignpair.cpp:107-118) and DSSAligner::AlignQueryTarget itself (batch of one).
 * Per pair the control flow is DSSAligner::AlignQueryTarget (dssaligner.cpp:793-831). */
int rsk_search_cross(rsk_ctx *ctx, const rsk_chainset *A, const rsk_chainset *B, const rsk_search_opts *opts, rsk_results **out);
int rsk_search_self(rsk_ctx *ctx, const rsk_chainset *S, const rsk_search_opts *opts, rsk_results **out);
int rsk_search_pairs(rsk_ctx *ctx, const rsk_chainset *A, const rsk_chainset *B, uint64_t npairs,
		const uint32_t *ia, const uint32_t *ib, const rsk_search_opts *opts, rsk_results **out);

/* Same hot loop with results left on the device (no D2H): used to time the resident-data kernel path. */
int rsk_search_cross_device(rsk_ctx *ctx, const rsk_chainset *A, const rsk_chainset *B, const rsk_search_opts *opts);

/* Alternative gapless Mu pre-scores: for every listed pair, out_profb[k] = SWFastGaplessProfb
 * (swgaplessprofb.cpp:6-61, float ScoreMx_Mu, forward minus reversed-A) and out_int[k] = SWFastPinopGapless
 * (swfastpinopgapless.cpp:6-47, IntScoreMx_Mu).  Either output pointer may be NULL.  Both sets need Mu letters. */
int rsk_mu_gapless_scores(rsk_ctx *ctx, const rsk_chainset *A, const rsk_chainset *B, uint64_t npairs,
		const uint32_t *ia, const uint32_t *ib, float *out_profb, int32_t *out_int);

/* ---- `-search Q -db DB -fast`: 5-mer prefilter + post-filter (search.cpp:76-111) ----
 * rsk_prefilter replaces MuPreFilter (muprefilter.cpp:64-133): query index MuDex::FromSeqDB incl. k-mer neighbourhoods
 * (mudex.cpp:386-442, mermx.cpp:484-584), PrefilterMu::Search (prefiltermu.cpp:382-393: index probe, two-hit diagonals
 * twohitdiag.cpp:47/389, FindHSP :12-48) and the per-query top-B bag RankedScoresBag (rankedscoresbag.cpp:34/5/185).
 * Q = the `-search` chains (query index side), T = the `-db` chains streamed in order (target index = position in T,
 * i.e. the reference at -threads 1).  Both sets need Mu letters.  The result is the content of the reference's
 * temporary candidate TSV: targets ascending, and for each target its queries ascending. */
typedef struct rsk_prefilter_opts {
	int32_t index_mode;   /* 0: as the reference (<= 100 queries: neighbourhoods in the query index, else on the target
	                         side; prefiltermu.cpp / mudex.cpp:146), 1: force -idxq, 2: force -idxt */
	uint32_t rsb_size;    /* per-query bag size B (-rsb_size); 0 = 1500 (prefiltermuparams.h) */
	int32_t no_kl_swap;   /* 0: exchange Mu letters 10 and 11 on the query side, as `-search -fast -db` does through
	                         g_CharToLetterMu (muprefilter.cpp:88, alpha.cpp:3291); 1: use the letters as given */
	int32_t raw_only;     /* 1: skip the per-query bag and return every (target, query, score) triple with a two-hit
	                         diagonal in stream order (DB-sharded searches merge the ranks' triples, then rsk_prefilter_bag) */
} rsk_prefilter_opts;
typedef struct rsk_prefilter_result rsk_prefilter_result;
int rsk_prefilter(rsk_ctx *ctx, const rsk_chainset *Q, const rsk_chainset *T, const rsk_prefilter_opts *opts,
		rsk_prefilter_result **out);
/* RankedScoresBag alone (rankedscoresbag.cpp:34/5/185) over triples in stream order (targets ascending); host only.
 * Multi-GPU: every rank runs rsk_prefilter(raw_only) on its block of the DB, the triples are concatenated in rank order
 * (target indices shifted to the unsharded DB) and every rank applies the bag: the result equals the single-GPU one. */
int rsk_prefilter_bag(uint32_t nq, uint64_t n, const uint32_t *t, const uint32_t *q, const uint16_t *s, uint32_t rsb_size,
		rsk_prefilter_result **out);
/* The same bag replayed on the device over caller-supplied triples (one warp per query; the kernel the sharded search runs on the
 * all-gathered stream).  Result identical to rsk_prefilter_bag. */
int rsk_prefilter_bag_device(rsk_ctx *ctx, uint32_t nq, uint64_t n, const uint32_t *t, const uint32_t *q, const uint16_t *s,
		uint32_t rsb_size, rsk_prefilter_result **out);
/* the candidates with t_lo <= target < t_hi, targets re-based to t_lo (one rank's share of a merged list); host only */
int rsk_prefilter_select(const rsk_prefilter_result *r, uint32_t t_lo, uint32_t t_hi, rsk_prefilter_result **out);
uint64_t rsk_prefilter_count(const rsk_prefilter_result *r);          /* candidate (target, query) pairs */
const uint32_t *rsk_prefilter_targets(const rsk_prefilter_result *r); /* [count] */
const uint32_t *rsk_prefilter_queries(const rsk_prefilter_result *r); /* [count] */
const uint16_t *rsk_prefilter_scores(const rsk_prefilter_result *r);  /* [count] best two-hit diagonal score */
uint64_t rsk_prefilter_raw_count(const rsk_prefilter_result *r);      /* (target, query) pairs with a two-hit diagonal, before the bag */
void rsk_prefilter_free(rsk_prefilter_result *r);
/* The candidate TSV itself (rankedscoresbag.cpp:219-232): "prefilter\t<#targets>\n" then "<t>\t<K>\t<q1>...\n".
 * Returns the length written (excluding the NUL) or, when cap is too small, the length needed as a negative number - 1. */
long long rsk_prefilter_to_tsv(const rsk_prefilter_result *r, char *out, size_t cap);

/* PostMuFilter (postmufilter.cpp:211-301, scan loop :116-208) over a prefilter result: every candidate line is aligned
 * with A = the query chain, B = the DB chain (AlignBags, chainbag.cpp:44-84) under the *sensitive* preset whatever the
 * context's mode (search.cpp:106-108) and reported with Up = true (hit.a = query index, hit.b = target index). */
int rsk_postfilter(rsk_ctx *ctx, const rsk_chainset *Q, const rsk_chainset *T, const rsk_prefilter_result *cands,
		const rsk_search_opts *opts, rsk_results **out);
/* Both stages: what `reseek -search Q -db DB -fast` computes. */
int rsk_search_fast_db(rsk_ctx *ctx, const rsk_chainset *Q, const rsk_chainset *T, const rsk_prefilter_opts *popts,
		const rsk_search_opts *opts, rsk_results **out);

/* ---- multi-GPU: the -db side block-partitioned over the GPUs of one box ----
 * The reference parallelises the search loop over std::threads that pull chains from one reader (runquery.cpp:82-125, body
 * :18-80; `-fast -db`: muprefilter.cpp:64-133 + postmufilter.cpp:116-208) and serialises hit emission under DBSearcher::m_Lock
 * (dbsearcher.cpp:267-278).  Here every rank (one per GPU; one process per GPU under torchrun, or one host thread per GPU inside
 * one process) owns a CONTIGUOUS block of the -db chains, queries and parameters are replicated, and the DP needs no exchange.
 * The data-path collectives are (1) the hit gather to the root rank - device-compacted hit records and path bytes travel as
 * exact-size NCCL point-to-point transfers over NVLink - and (2) for `-fast -db` the all-gather of the per-block prefilter
 * triples in rank order, which IS the stream order of the unsharded DB, so that every rank replays the same RankedScoresBag
 * stream (rankedscoresbag.cpp:34-51) and the merged top-B lists equal the single-GPU ones, ties at the cut-off included.
 * A communicator belongs to one context; all ranks must make the same sharded calls in the same order (collective semantics).
 * comm == NULL means "one rank": the same device-compacted path without any transfer. */
typedef struct rsk_comm rsk_comm;
#define RSK_COMM_ID_BYTES 128
int rsk_comm_unique_id(void *id /* [RSK_COMM_ID_BYTES], made on one rank and given to all (ncclGetUniqueId) */);
int rsk_comm_create(rsk_ctx *ctx, int nranks, int rank, const void *id, rsk_comm **out);
/* one process, several GPUs: communicators for n contexts on n different devices (ncclCommInitAll); out[n] */
int rsk_comm_create_all(rsk_ctx *const *ctxs, int n, rsk_comm **out);
void rsk_comm_destroy(rsk_comm *comm);
int rsk_comm_rank(const rsk_comm *comm);
int rsk_comm_nranks(const rsk_comm *comm);
typedef struct rsk_comm_stats {
	uint64_t bytes_sent, bytes_recv;   /* payload this rank put on / took from NVLink (its own block is a device-local copy) */
	float collective_ms;               /* device time of the gather / all-gather rounds (CUDA events on the context stream) */
	uint32_t collectives;
} rsk_comm_stats;
int rsk_comm_get_stats(const rsk_comm *comm, rsk_comm_stats *out);
int rsk_comm_reset_stats(rsk_comm *comm);
/* contiguous blocks with (nearly) equal residue totals: rank r owns chains bounds[r] .. bounds[r+1]-1; bounds[nranks+1] */
int rsk_partition_by_residues(const uint32_t *len, uint32_t n, int nranks, uint32_t *bounds);

/* DBSearcher::RunQuery over this rank's block of the -db chains (A_local; NULL = empty block) against the replicated in-memory
 * chains B; a_base = index of the block's first chain in the unsharded DB.  The records the reference would emit (opts->keep ==
 * RSK_KEEP_HITS) or all of them (RSK_KEEP_ALL) are compacted on the device, gathered on `root` and returned there in the order
 * of the unsharded search (hit.a = DB chain index in the whole DB); *out is NULL on the other ranks. */
int rsk_search_cross_sharded(rsk_ctx *ctx, rsk_comm *comm, const rsk_chainset *A_local, const rsk_chainset *B, uint32_t a_base,
		const rsk_search_opts *opts, int root, rsk_results **out);
/* DBSearcher::RunSelf (runself.cpp:72-145) over several GPUs: every rank holds the whole set S and takes the rows i = rank,
 * rank + nranks, ... of the pair triangle (i <= j); the hits are gathered on `root` and returned there in the order of
 * rsk_search_self with RSK_KEEP_HITS (a ascending, then b); *out is NULL on the other ranks. */
int rsk_search_self_sharded(rsk_ctx *ctx, rsk_comm *comm, const rsk_chainset *S, const rsk_search_opts *opts, int root, rsk_results **out);
/* `reseek -search Q -db DB -fast` with T_local = this rank's block of the DB (t_base = its first target index): local prefilter
 * kernels, triples all-gathered in rank order, the bag replayed on the merged stream, the candidates of the own block
 * post-filtered, hits gathered on `root` (hit.a = query, hit.b = target index in the whole DB).  cands_out (optional, every
 * rank) receives the merged candidate list = the reference's candidate TSV content. */
int rsk_search_fast_db_sharded(rsk_ctx *ctx, rsk_comm *comm, const rsk_chainset *Q, const rsk_chainset *T_local, uint32_t t_base,
		const rsk_prefilter_opts *popts, const rsk_search_opts *opts, int root, rsk_results **out, rsk_prefilter_result **cands_out);
/* order-independent 64-bit digest of a result set (records and path bytes): equal for the single-GPU and the sharded search */
uint64_t rsk_results_digest(const rsk_results *r);

/* -global: DSSAligner::AlignQueryTarget_Global (global.cpp:7-33) for explicit pairs (alignpair.cpp:110-114, runself.cpp:48-57,
 * scop40bench.cpp:313): the Mu filter when omega > 0, then ViterbiFastMem (viterbifastmem.cpp:33-193: three-state global
 * alignment, gap open -1, extend -0.05, terminal gaps free) with TraceBackBitMem.  One record per pair, in pair order:
 * flags has RSK_HIT_GLOBAL (and RSK_HIT_MU_REJECTED, score -9999, no path when the filter said no); score = m_GlobalScore,
 * lo_a = lo_b = 0, the path covers both chains completely; E-value fields stay at their ClearAlign values (FLT_MAX) as in
 * the reference, which does not compute them on this path.  LA*LB <= 1e8 per pair (the reference dies above that). */
int rsk_align_global(rsk_ctx *ctx, const rsk_chainset *A, const rsk_chainset *B, uint64_t npairs,
		const uint32_t *ia, const uint32_t *ib, rsk_results **out);

/* ---- results ---- */
uint64_t rsk_results_count(const rsk_results *r);
const rsk_hit *rsk_results_hits(const rsk_results *r);  /* [count], host memory owned by r */
const char *rsk_results_paths(const rsk_results *r);    /* path pool: 'M','D','I' bytes, hit k at path_off..+path_len */
uint64_t rsk_results_paths_bytes(const rsk_results *r);
void rsk_results_free(rsk_results *r);

/* ---- hit writers: DSSAligner::ToTsv / WriteUserField (dssaligner.cpp:1016, userfields.cpp:45-152), PathToCIGAR (cigar.cpp:95) ----
 * Host-only formatting, byte-compatible with the reference's -output TSV.  up != 0: query = A, target = B. */
typedef struct rsk_hit_view {
	const rsk_hit *hit;
	const char *path;               /* the hit's M/D/I path (path pool + hit->path_off), may be NULL without cigar/pctid */
	const char *label_a, *label_b;  /* chain labels */
	const char *seq_a, *seq_b;      /* amino-acid sequences (only for pctid), may be NULL */
	uint32_t len_a, len_b;
} rsk_hit_view;
int rsk_path_to_cigar(const char *path, uint32_t path_len, int up, char *out, size_t cap);
/* columns: '+'-separated names as for -columns (userfieldnames.h); NULL = the reference's default ("std", usage.h:49).
 * Returns the line length (no newline) or a negative rsk_status. */
int rsk_format_tsv(const rsk_hit_view *v, int up, const char *columns, char *out, size_t cap);
/* The block DSSAligner::ToAln appends to the -aln file for one hit (dssaligner.cpp:965-979 -> PrettyAln prettyaln.cpp:26-99,
 * WriteLocalAln writelocalaln.cpp:65-100); rowlen = 0 means the default of 80 columns (-rowlen).  Needs path, labels and both
 * sequences.  Returns the length written (excluding the NUL); when cap is too small, -(bytes needed) - 1, which is
 * never above -17; a negative rsk_status (-1 .. -4) on bad arguments. */
long long rsk_format_aln(const rsk_hit_view *v, int up, uint32_t rowlen, char *out, size_t cap);
/* The record DSSAligner::ToFasta2 appends to the -fasta2 file (dssaligner.cpp:981-1014): target row first, then the query
 * row, 80 residues per line, one empty line after; global != 0 is -unaligned (lower-case flanks, '.' padding).  Same return
 * convention as rsk_format_aln. */
long long rsk_format_fasta2(const rsk_hit_view *v, int up, int global, char *out, size_t cap);

/* ---- superposition: DSSAligner::GetKabsch (dssaligner.cpp:1371-1385) -> Kabsch (kabsch.cpp:330-387) ----
 * Host code in double precision, as in the reference.  xyz_* = [3][len] coordinate planes of the two chains (the layout of
 * rsk_chains_host for one chain), path/lo_* from the hit.  On return y ~ u x + t maps a QUERY residue x onto its target
 * partner (up != 0: query = A), u row-major 3x3; *msd = residual sum of squares / number of M columns, the value the
 * reference's Kabsch() returns.  Horn's quaternion method instead of the reference's TM-align routine: same optimum, results
 * agree to ~1e-6 (tests/test_cabi.py).  Applying it to PDB ATOM lines (-alignpair -output) is left to the caller. */
int rsk_kabsch(const float *xyz_a, uint32_t len_a, const float *xyz_b, uint32_t len_b, uint32_t lo_a, uint32_t lo_b,
		const char *path, uint32_t path_len, int up, double t[3], double u[9], double *msd);

/* ---- host-side statistics: StatSig (statsig.cpp:27-50, statsig.h:8-23); libm double pow, as the reference ---- */
double rsk_pvalue(double ts);
double rsk_evalue(double ts);
double rsk_qual(double ts);

#ifdef __cplusplus
}
#endif
#endif /* RESEEK_B200_H */
