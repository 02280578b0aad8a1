/* reseek_oracle.c - CPU restatement of the Reseek per-pair search hot path (see reseek_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: parity checker for the CUDA library; never part of the product path.
 * Parity status: PINNED against oracle/_ref (the unmodified reference) - tests/test_oracle_vs_reference.py.
 *
 * Written from the behaviour described in SURVEY.md Appendix A and the cited reference lines; it is a
 * restatement (flat arrays, explicit DP matrices), not a copy.  Build: -O2 -ffp-contract=off, no fast-math.
 */
#include "reseek_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../reseek_b200/csrc/score_tables_data.inc"

/* xdpmem.h:6 - a finite "minus infinity" */
#define ORC_NEG_INF (-9e9f)

int orc_feat_alpha(int f) { return rsk_tbl_feat_alpha[f]; }

int orc_feat_offset(int f)
{
	int off = 0;
	for (int k = 0; k < f; ++k)
		off += rsk_tbl_feat_alpha[k] * rsk_tbl_feat_alpha[k];
	return off;
}

const float *orc_bgfreq(int f)
{
	int off = 0;
	for (int k = 0; k < f; ++k)
		off += rsk_tbl_feat_alpha[k];
	return rsk_tbl_bgfreq + off;
}

const signed char *orc_mu_i8(void) { return rsk_tbl_mu_i8; }
const signed char *orc_mu_kmer_i8(void) { return rsk_tbl_mu_kmer_i8; }
const float *orc_mu_f32(void) { return rsk_tbl_mu_f32; }

/* namedparams.cpp:32-53 SetDefaults, dssparams.cpp:44-111 SetDSSParams, dssparams.cpp:344-364 ApplyWeights */
int orc_params_preset(orc_params *p, int mode)
{
	memset(p, 0, sizeof(*p));
	p->gap_open = rsk_tbl_gap_open;
	p->gap_ext = rsk_tbl_gap_ext;
	p->min_fwd_score = 7.0f;
	p->mu_gap_open = 2;
	p->mu_gap_ext = 1;
	switch (mode) {
	case ORC_MODE_FAST:
		p->omega = 22; p->omega_fwd = 50; p->mkfl = 500;
		p->mkf_x1 = 8; p->mkf_x2 = 8; p->mkf_min_hsp_score = 50; p->mkf_min_mega_hsp_score = -4;
		break;
	case ORC_MODE_SENSITIVE:
		p->omega = 12; p->omega_fwd = 20; p->mkfl = 600;
		p->mkf_x1 = 8; p->mkf_x2 = 8; p->mkf_min_hsp_score = 50; p->mkf_min_mega_hsp_score = -4;
		break;
	case ORC_MODE_VERYSENSITIVE:
		p->omega = 0; p->omega_fwd = 0; p->mkfl = 99999;
		p->mkf_x1 = 99999; p->mkf_x2 = 99999; p->mkf_min_hsp_score = 0; p->mkf_min_mega_hsp_score = -99999;
		p->min_fwd_score = 0;
		break;
	default:
		return -1;
	}
	int off = 0;
	for (int f = 0; f < ORC_NFEAT; ++f) {
		const float w = rsk_tbl_feat_weight[f];
		const int n = rsk_tbl_feat_alpha[f] * rsk_tbl_feat_alpha[f];
		p->weights[f] = w;
		for (int k = 0; k < n; ++k)
			p->tables[off + k] = w * rsk_tbl_logodds[off + k]; /* fp32 multiply */
		off += n;
	}
	return 0;
}

float orc_cell_score(const orc_params *p, const uint8_t *profA, uint32_t LA, uint32_t i,
		const uint8_t *profB, uint32_t LB, uint32_t j)
{
	float s = 0;
	int off = 0;
	for (int f = 0; f < ORC_NFEAT; ++f) {
		const int n = rsk_tbl_feat_alpha[f];
		const float t = p->tables[off + profA[(size_t)f * LA + i] * n + profB[(size_t)f * LB + j]];
		s = (f == 0) ? t : s + t; /* feature 0 assigns, the rest accumulate: dssaligner.cpp:557-595 */
		off += n * n;
	}
	return s;
}

/* trace codes per cell: 2 bits = where the match state came from, +4 = D opened from M, +8 = I opened from M */
enum { SRC_M = 0, SRC_D = 1, SRC_I = 2, SRC_START = 3, BIT_MD = 4, BIT_MI = 8 };

/* The recurrence of sw.cpp:119-197 on full matrices instead of rolling rows; scorefn supplies S[i][j]. */
typedef float (*cellfn)(const void *ctx, uint32_t i, uint32_t j);

static float sw_core(cellfn fn, const void *ctx, uint32_t LA, uint32_t LB, float open, float ext,
		uint32_t *lo_a, uint32_t *lo_b, char *path, uint32_t *path_len)
{
	*path_len = 0;
	if (path)
		path[0] = 0;
	*lo_a = UINT32_MAX;
	*lo_b = UINT32_MAX;
	if (LA == 0 || LB == 0)
		return 0.0f;
	float *Mprev = (float *)malloc(sizeof(float) * (LB + 1)); /* M[i][0..LB]: M[i][j] = match score ending at (i-1,j-1) */
	float *Mcur = (float *)malloc(sizeof(float) * (LB + 1));
	float *D = (float *)malloc(sizeof(float) * LB);           /* D[i][j] for the current i */
	uint8_t *tb = (uint8_t *)malloc((size_t)LA * LB);
	for (uint32_t j = 0; j <= LB; ++j)
		Mprev[j] = ORC_NEG_INF;
	Mprev[0] = 0.0f; /* sw.cpp:116: M0 = 0 for cell (0,0) only */
	for (uint32_t j = 0; j < LB; ++j)
		D[j] = ORC_NEG_INF;
	float best = 0.0f;
	uint32_t bi = UINT32_MAX, bj = UINT32_MAX;
	for (uint32_t i = 0; i < LA; ++i) {
		float I = ORC_NEG_INF; /* I[i][0] */
		Mcur[0] = ORC_NEG_INF; /* sw.cpp:196: M0 = -inf at the start of rows >= 1 */
		for (uint32_t j = 0; j < LB; ++j) {
			const float m = Mprev[j];
			uint8_t t = SRC_M;
			float x = m;
			if (D[j] > x) { x = D[j]; t = SRC_D; }
			if (I > x) { x = I; t = SRC_I; }
			if (0.0f >= x) { x = 0.0f; t = SRC_START; }
			x += fn(ctx, i, j);
			if (x > best) { best = x; bi = i; bj = j; } /* strict >, row-major first max */
			Mcur[j + 1] = x;
			const float mo = m + open;
			float d = D[j] + ext;
			if (mo >= d) { d = mo; t |= BIT_MD; }
			D[j] = d; /* D[i+1][j] */
			float ins = I + ext;
			if (mo >= ins) { ins = mo; t |= BIT_MI; }
			I = ins; /* I[i][j+1] */
			tb[(size_t)i * LB + j] = t;
		}
		float *tmp = Mprev; Mprev = Mcur; Mcur = tmp;
	}
	if (best == 0.0f) {
		free(Mprev); free(Mcur); free(D); free(tb);
		return 0.0f;
	}
	/* traceback, sw.cpp:8-77 (1-based i,j; state M first) */
	uint32_t i = bi + 1, j = bj + 1, n = 0;
	char state = 'M';
	char *rev = (char *)malloc((size_t)LA + LB + 2);
	for (;;) {
		rev[n++] = state;
		if (state == 'M') {
			const uint8_t t = tb[(size_t)(i - 1) * LB + (j - 1)];
			const int src = t & 3;
			--i; --j;
			if (src == SRC_D) state = 'D';
			else if (src == SRC_I) state = 'I';
			else if (src == SRC_START) break;
		} else if (state == 'D') {
			const uint8_t t = tb[(size_t)(i - 1) * LB + j];
			--i;
			state = (t & BIT_MD) ? 'M' : 'D';
		} else {
			const uint8_t t = tb[(size_t)i * LB + (j - 1)];
			--j;
			state = (t & BIT_MI) ? 'M' : 'I';
		}
	}
	*lo_a = i; /* = Besti+1-Leni */
	*lo_b = j;
	*path_len = n;
	if (path) {
		for (uint32_t k = 0; k < n; ++k)
			path[k] = rev[n - 1 - k];
		path[n] = 0;
	}
	free(rev); free(Mprev); free(Mcur); free(D); free(tb);
	return best;
}

typedef struct { const orc_params *p; const uint8_t *a, *b; uint32_t LA, LB; } prof_ctx;
static float prof_cell(const void *c, uint32_t i, uint32_t j)
{
	const prof_ctx *x = (const prof_ctx *)c;
	return orc_cell_score(x->p, x->a, x->LA, i, x->b, x->LB, j);
}

typedef struct { const float *S; uint32_t LB; } mx_ctx;
static float mx_cell(const void *c, uint32_t i, uint32_t j)
{
	const mx_ctx *x = (const mx_ctx *)c;
	return x->S[(size_t)i * x->LB + j];
}

float orc_sw_align(const orc_params *p, const uint8_t *profA, uint32_t LA, const uint8_t *profB, uint32_t LB,
		uint32_t *lo_a, uint32_t *lo_b, char *path, uint32_t *path_len)
{
	prof_ctx c = {p, profA, profB, LA, LB};
	return sw_core(prof_cell, &c, LA, LB, p->gap_open, p->gap_ext, lo_a, lo_b, path, path_len);
}

float orc_swfast_matrix(const float *S, uint32_t LA, uint32_t LB, float open, float ext,
		uint32_t *lo_a, uint32_t *lo_b, char *path, uint32_t *path_len)
{
	mx_ctx c = {S, LB};
	return sw_core(mx_cell, &c, LA, LB, open, ext, lo_a, lo_b, path, path_len);
}

/* Scalar Gotoh, every value floored at 0 (the int8 lanes are biased by -128 and saturate downwards:
 * parasail.cpp:580-583, 671-690).  E/F are derived from H (not from the other gap state). */
int orc_mu_sw_score(const uint8_t *a, uint32_t LA, const uint8_t *b, uint32_t LB, int open, int ext, int *saturated)
{
	int best = 0;
	if (saturated)
		*saturated = 0;
	if (LA == 0 || LB == 0)
		return 0;
	int *H = (int *)calloc(LA + 1, sizeof(int)); /* H[i] for the previous column, index i+1 */
	int *E = (int *)calloc(LA + 1, sizeof(int)); /* gap state running along b for each query row */
	for (uint32_t j = 0; j < LB; ++j) {
		const signed char *row = rsk_tbl_mu_i8 + 36 * b[j];
		int diag = 0, F = 0;
		for (uint32_t i = 0; i < LA; ++i) {
			int h = diag + row[a[i]];
			if (h < 0) h = 0;
			if (E[i] > h) h = E[i];
			if (F > h) h = F;
			diag = H[i + 1];
			H[i + 1] = h;
			if (h > best) best = h;
			int ho = h - open; if (ho < 0) ho = 0;
			int e = E[i] - ext; if (e < 0) e = 0;
			E[i] = e > ho ? e : ho;
			int f = F - ext; if (f < 0) f = 0;
			F = f > ho ? f : ho;
		}
	}
	free(H); free(E);
	if (saturated)
		*saturated = best > 250; /* maxp = 127-(4+1) biased: parasail.cpp:585, 728-733 */
	return best;
}

float orc_mu_filter_score(const orc_params *p, const uint8_t *a, uint32_t LA, const uint8_t *b, uint32_t LB,
		int *fwd_out, int *rev_out)
{
	int sat = 0;
	int fwd = orc_mu_sw_score(a, LA, b, LB, p->mu_gap_open, p->mu_gap_ext, &sat);
	if (sat)
		fwd = 777; /* parasail_mu.cpp:133-137 */
	if (fwd_out) *fwd_out = fwd;
	if (rev_out) *rev_out = 0;
	if ((float)fwd < p->omega_fwd)
		return 0; /* :139-144 */
	uint8_t *ar = (uint8_t *)malloc(LA ? LA : 1);
	for (uint32_t i = 0; i < LA; ++i)
		ar[i] = a[LA - 1 - i]; /* parasail_mu.cpp:174-177 */
	int rev = orc_mu_sw_score(ar, LA, b, LB, p->mu_gap_open, p->mu_gap_ext, &sat);
	free(ar);
	if (sat)
		rev = 255; /* score read before the 777 assignment (:149-155); saturated result->score = 127+128 */
	if (rev_out) *rev_out = rev;
	return (float)fwd - (float)rev;
}

static float dist2(const orc_chain *c, uint32_t p1, uint32_t p2)
{
	/* pdbchain.cpp:320-336 */
	const float dx = c->x[p1] - c->x[p2];
	const float dy = c->y[p1] - c->y[p2];
	const float dz = c->z[p1] - c->z[p2];
	return dx * dx + dy * dy + dz * dz;
}

float orc_lddt(const orc_chain *A, const orc_chain *B, const uint32_t *posA, const uint32_t *posB, uint32_t n)
{
	static const float R0sq = 15.0f * 15.0f;
	static const float thr[4] = {0.5f, 1.0f, 2.0f, 4.0f};
	if (n == 0)
		return 0;
	uint32_t *cons = (uint32_t *)calloc(n, sizeof(uint32_t));
	uint32_t *pres = (uint32_t *)calloc(n, sizeof(uint32_t));
	for (uint32_t ci = 0; ci < n; ++ci) {
		for (uint32_t cj = ci + 1; cj < n; ++cj) {
			const float d1s = dist2(A, posA[ci], posA[cj]);
			const float d2s = dist2(B, posB[ci], posB[cj]);
			if (d1s > R0sq && d2s > R0sq)
				continue;
			const float d1 = sqrtf(d1s), d2 = sqrtf(d2s);
			const float diff = fabsf(d1 - d2);
			for (int k = 0; k < 4; ++k)
				if (diff <= thr[k]) { pres[ci]++; pres[cj]++; }
			cons[ci] += 4;
			cons[cj] += 4;
		}
	}
	float total = 0;
	for (uint32_t c = 0; c < n; ++c) {
		float score = 0;
		if (cons[c] > 0)
			score = (float)pres[c] / (float)cons[c];
		total += score;
	}
	free(cons); free(pres);
	return total / (float)n;
}

double orc_pvalue(double ts)
{
	const double log10p = (ts < 0.11) ? (-80.0 * ts + -0.58) : (-52.0 * ts + -3.7);
	double P = pow(10, log10p);
	if (P > 1) P = 1;
	return P;
}

double orc_evalue(double ts) { return orc_pvalue(ts) * 8340; /* statsig.h:3 */ }

double orc_qual(double ts)
{
	const double logE = 5.0 + -40.0 * ts;
	if (logE < -20)
		return 1;
	const double x = pow(10, logE / 10);
	return 1 / (1 + x / 2);
}

void orc_calc_evalue(const orc_params *p, const orc_chain *A, const orc_chain *B, const char *path, orc_result *r)
{
	if (r->score < p->min_fwd_score)
		return; /* dssaligner.cpp:861-862: everything stays at its ClearAlign value */
	const uint32_t n = r->path_len;
	uint32_t M = 0, D = 0, I = 0;
	uint32_t *pa = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
	uint32_t *pb = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
	uint32_t ia = r->lo_a, ib = r->lo_b;
	for (uint32_t k = 0; k < n; ++k) {
		if (path[k] == 'M') { pa[M] = ia++; pb[M] = ib++; ++M; }
		else if (path[k] == 'D') { ++ia; ++D; }
		else { ++ib; ++I; }
	}
	r->hi_a = r->lo_a + M + D - 1;
	r->hi_b = r->lo_b + M + I - 1;
	r->ids = M;
	r->gaps = D + I;
	const float lddt = orc_lddt(A, B, pa, pb, M);
	free(pa); free(pb);
	float rev = 0;
	if (A->selfrev != FLT_MAX && B->selfrev != FLT_MAX)
		rev = (A->selfrev + B->selfrev) / 2;
	const float L = (float)(A->L + B->L) / 2;
	float ts = 0.13f * lddt;
	ts += (1.7f * r->score - 2.0f * rev) / (L + 250.0f);
	r->lddt = lddt;
	r->ts = ts;
	r->pvalue = (float)orc_pvalue(ts);
	r->qual = (float)orc_qual(ts);
	r->evalue = (float)orc_evalue(ts);
}

static void clear_result(orc_result *r)
{
	/* dssaligner.cpp:906-927 ClearAlign */
	memset(r, 0, sizeof(*r));
	r->lo_a = r->lo_b = r->hi_a = r->hi_b = r->ids = r->gaps = UINT32_MAX;
	r->pvalue = r->evalue = FLT_MAX;
	r->qual = FLT_MAX;
	r->ts = -FLT_MAX;
	r->lddt = 0;
}

/* Gapless variants: on every diagonal a running sum that restarts from 0 whenever it went negative. */
float orc_mu_gapless_profb(const uint8_t *a, uint32_t LA, const uint8_t *b, uint32_t LB)
{
	float bestf = 0, bestr = 0;
	float *prevf = (float *)malloc(sizeof(float) * (LB + 1)), *prevr = (float *)malloc(sizeof(float) * (LB + 1));
	float *curf = (float *)malloc(sizeof(float) * (LB + 1)), *curr = (float *)malloc(sizeof(float) * (LB + 1));
	for (uint32_t j = 0; j <= LB; ++j) prevf[j] = prevr[j] = ORC_NEG_INF; /* row "-1": every diagonal starts fresh */
	for (uint32_t i = 0; i < LA; ++i) {
		const float *rowf = rsk_tbl_mu_f32 + 36 * a[i], *rowr = rsk_tbl_mu_f32 + 36 * a[LA - 1 - i];
		curf[0] = curr[0] = 0; /* M0 = 0 at the start of a row (swgaplessprofb.cpp:55-56; the first row starts at 0 too) */
		for (uint32_t j = 0; j < LB; ++j) {
			float xf = (j == 0) ? 0.0f : prevf[j], xr = (j == 0) ? 0.0f : prevr[j]; /* diagonal predecessor (i-1,j-1) */
			if (xf < 0.0f) xf = 0.0f;
			if (xr < 0.0f) xr = 0.0f;
			xf += rowf[b[j]];
			xr += rowr[b[j]];
			if (xf > bestf) bestf = xf;
			if (xr > bestr) bestr = xr;
			curf[j + 1] = xf;
			curr[j + 1] = xr;
		}
		float *t = prevf; prevf = curf; curf = t;
		t = prevr; prevr = curr; curr = t;
	}
	free(prevf); free(prevr); free(curf); free(curr);
	return bestf - bestr;
}

int orc_mu_gapless_int(const uint8_t *a, uint32_t LA, const uint8_t *b, uint32_t LB)
{
	int best = 0;
	int *prev = (int *)calloc(LB + 1, sizeof(int)), *cur = (int *)calloc(LB + 1, sizeof(int));
	for (uint32_t i = 0; i < LA; ++i) {
		const signed char *row = rsk_tbl_mu_i8 + 36 * a[i];
		cur[0] = 0;
		for (uint32_t j = 0; j < LB; ++j) {
			int x = (j == 0) ? 0 : prev[j];
			if (x < 0) x = 0;
			x += row[b[j]];
			if (x > best) best = x;
			cur[j + 1] = x;
		}
		int *t = prev; prev = cur; cur = t;
	}
	free(prev); free(cur);
	return best;
}

/* ------------------------------------------------------------------------------------------------
 * Long-chain path
 * ------------------------------------------------------------------------------------------------ */
int orc_mu_xdrop(const uint8_t *Q, int LQ, const uint8_t *T, int LT, int PosQ, int PosT, int X, int *Loi, int *Loj, int *Len)
{
	*Loi = PosQ;
	*Loj = PosT;
	int fwd = 0, bestfwd = 0, fwdlen = 0;
	for (int i = PosQ, j = PosT; i < LQ && j < LT; ) {
		fwd += rsk_tbl_mu_i8[36 * Q[i] + T[j]];
		++i; ++j;
		if (fwd > bestfwd) { bestfwd = fwd; fwdlen = i - PosQ; }
		else if (fwd + X < bestfwd) break;
	}
	int rev = 0, bestrev = 0, revlen = 0;
	for (int i = PosQ - 1, j = PosT - 1; i >= 0 && j >= 0; --i, --j) {
		rev += rsk_tbl_mu_i8[36 * Q[i] + T[j]];
		if (rev > bestrev) { bestrev = rev; *Loi = i; *Loj = j; revlen = PosQ - i; }
		else if (rev + X < bestrev) break;
	}
	*Len = fwdlen + revlen;
	return bestfwd + bestrev;
}

/* 1-D chaining of intervals on the query axis (chainer.cpp:31-194).  Breakpoints sorted by position, interval
 * starts before interval ends at equal positions; remaining ties keep input order (stable). */
static float chain_hsps(const orc_hsp *h, int n, int *idx_out, int *nidx)
{
	*nidx = 0;
	if (n == 0)
		return 0;
	typedef struct { uint32_t pos; int is_lo; int index; } bp_t;
	bp_t *bp = (bp_t *)malloc(sizeof(bp_t) * 2 * n);
	for (int i = 0; i < n; ++i) {
		bp[2 * i].pos = (uint32_t)h[i].loi; bp[2 * i].is_lo = 1; bp[2 * i].index = i;
		bp[2 * i + 1].pos = (uint32_t)(h[i].loi + h[i].len - 1); bp[2 * i + 1].is_lo = 0; bp[2 * i + 1].index = i;
	}
	for (int i = 1; i < 2 * n; ++i) { /* stable insertion sort */
		bp_t t = bp[i];
		int k = i - 1;
		while (k >= 0 && (bp[k].pos > t.pos || (bp[k].pos == t.pos && !bp[k].is_lo && t.is_lo))) {
			bp[k + 1] = bp[k];
			--k;
		}
		bp[k + 1] = t;
	}
	float *cs = (float *)malloc(sizeof(float) * n);
	int *tb = (int *)malloc(sizeof(int) * n);
	int best_end = -1;
	for (int i = 0; i < 2 * n; ++i) {
		const int ix = bp[i].index;
		const float sc = (float)h[ix].score;
		if (bp[i].is_lo) {
			tb[ix] = best_end;
			cs[ix] = best_end < 0 ? sc : cs[best_end] + sc;
		} else if (best_end < 0 || cs[ix] > cs[best_end]) {
			best_end = ix;
		}
	}
	float total = 0;
	for (int ix = best_end; ix >= 0; ix = tb[ix]) {
		total += (float)h[ix].score;
		idx_out[(*nidx)++] = ix;
	}
	free(bp); free(cs); free(tb);
	return total;
}

int orc_mkf_align(const orc_params *p, const uint8_t *muQ, int LQ, const uint8_t *muT, int LT,
		orc_hsp *hsps, int cap, int *nhsp, int *best_hsp_score, int *best_chain_score, int *chain_idx, int *nchain)
{
	*nhsp = 0; *best_hsp_score = 0; *best_chain_score = 0; *nchain = 0;
	if (LQ < 3 || LT < 3)
		return 0;
	/* query hash: the first 4 positions of every 3-mer (mukmerfilter.cpp:208-230) */
	const int D = 36 * 36 * 36;
	uint16_t *ht = (uint16_t *)malloc(sizeof(uint16_t) * 4 * D);
	memset(ht, 0xff, sizeof(uint16_t) * 4 * D);
	for (int pos = 0; pos + 3 <= LQ; ++pos) {
		const int k = (muQ[pos] * 36 + muQ[pos + 1]) * 36 + muQ[pos + 2];
		for (int w = 0; w < 4; ++w)
			if (ht[4 * k + w] == 0xffff) { ht[4 * k + w] = (uint16_t)pos; break; }
	}
	int found = 0, best = 0, n = 0;
	for (int pt = 0; pt + 3 <= LT; ++pt) {
		const int k = (muT[pt] * 36 + muT[pt + 1]) * 36 + muT[pt + 2];
		for (int w = 0; w < 4; ++w) {
			const int pq = ht[4 * k + w];
			if (pq == 0xffff)
				continue;
			int loi, loj, len;
			const int sc = orc_mu_xdrop(muQ, LQ, muT, LT, pq, pt, p->mkf_x1, &loi, &loj, &len);
			if (sc >= p->mkf_min_hsp_score) {
				found = 1;
				if (sc > best) { /* order-dependent gating, mukmerfilter.cpp:354-377 */
					best = sc;
					int old = 0;
					for (int i = 0; i < n; ++i)
						if (hsps[i].loi == loi) { old = 1; break; }
					if (!old && n < cap) {
						hsps[n].loi = loi; hsps[n].loj = loj; hsps[n].len = len; hsps[n].score = sc;
						++n;
					}
				}
			}
		}
	}
	free(ht);
	*nhsp = n;
	*best_hsp_score = best;
	if (found)
		*best_chain_score = (int)chain_hsps(hsps, n, chain_idx, nchain);
	return found;
}

float orc_mega_hsp_score(const orc_params *p, const uint8_t *profA, uint32_t LA, const uint8_t *profB, uint32_t LB,
		uint32_t lo_i, uint32_t lo_j, uint32_t len)
{
	float total = 0;
	int off = 0;
	for (int f = 0; f < ORC_NFEAT; ++f) {
		const int n = rsk_tbl_feat_alpha[f];
		for (uint32_t k = 0; k < len; ++k)
			total += p->tables[off + profA[(size_t)f * LA + lo_i + k] * n + profB[(size_t)f * LB + lo_j + k]];
		off += n * n;
	}
	return total;
}

/* xdrophsp.cpp:8-33 SubstScore: starts from 0 and adds the 8 features in order */
static float subst(const orc_params *p, const uint8_t *profA, uint32_t LA, const uint8_t *profB, uint32_t LB,
		uint32_t pa, uint32_t pb)
{
	float t = 0;
	int off = 0;
	for (int f = 0; f < ORC_NFEAT; ++f) {
		const int n = rsk_tbl_feat_alpha[f];
		t += p->tables[off + profA[(size_t)f * LA + pa] * n + profB[(size_t)f * LB + pb]];
		off += n * n;
	}
	return t;
}

enum { XB_DM = 1, XB_IM = 2, XB_MD = 4, XB_MI = 8 };
#define XNONE 0xffffffffu
static uint32_t umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
static uint32_t umax(uint32_t a, uint32_t b) { return a > b ? a : b; }

float orc_xdrop_fwd(const orc_params *p, const uint8_t *profA, uint32_t LAfull, const uint8_t *profB, uint32_t LBfull,
		int reverse, uint32_t revLA, uint32_t revLB, float X, uint32_t LoA, uint32_t aLA, uint32_t LoB, uint32_t aLB,
		char *path, uint32_t *path_len)
{
#define SUB(pa, pb) (reverse ? subst(p, profA, LAfull, profB, LBfull, revLA - (pa) - 1, revLB - (pb) - 1) \
                             : subst(p, profA, LAfull, profB, LBfull, (pa), (pb)))
	*path_len = 0;
	path[0] = 0;
	const uint32_t LA = aLA - LoA, LB = aLB - LoB;
	const float open = p->gap_open, ext = p->gap_ext;
	if (LA == 1 || LB == 1) { /* xdropfwd.cpp:87-93 */
		const float sc = SUB(LoA, LoB);
		if (sc > 0) { path[0] = 'M'; path[1] = 0; *path_len = 1; }
		return sc;
	}
	const float absopen = -open, absext = -ext;
	float *Mbuf = (float *)malloc(sizeof(float) * (LB + 4));
	float *M = Mbuf + 1; /* M[-1] is addressable */
	float *Dr = (float *)malloc(sizeof(float) * (LB + 4));
	const size_t W = (size_t)LB + 3;
	uint8_t *tb = (uint8_t *)calloc((size_t)(LA + 3) * W, 1);
	M[-1] = ORC_NEG_INF;
	Dr[0] = ORC_NEG_INF;
	Dr[1] = ORC_NEG_INF;
	float best = 0;
	uint32_t besti = 0, bestj = 0;
	uint32_t prev_jlo = 0, prev_jhi = 0, jlo = 1, jhi = 1;
	float M0 = best;
	for (uint32_t i = 1; i <= LA; ++i) {
		if (jlo == prev_jlo) {
			M[jlo - 1] = ORC_NEG_INF;
			Dr[jlo] = ORC_NEG_INF;
		}
		uint32_t endj = umin(prev_jhi + 1, LB);
		for (uint32_t j = endj + 1; j <= umin(jhi + 1, LB); ++j) {
			M[j - 1] = ORC_NEG_INF;
			Dr[j] = ORC_NEG_INF;
		}
		uint32_t next_jlo = XNONE, next_jhi = XNONE;
		float I0 = ORC_NEG_INF;
		uint8_t *row = tb + (size_t)i * W;
		for (uint32_t j = jlo; j <= jhi; ++j) {
			uint8_t bits = 0;
			const float saved = M0;
			float x = M0;
			if (Dr[j] > x) { x = Dr[j]; bits = XB_DM; }
			if (I0 > x) { x = I0; bits = XB_IM; }
			M0 = M[j];
			float s = SUB(LoA + i - 1, LoB + j - 1);
			s += x;
			M[j] = s;
			float h = s - best + X;
			if (h > 0) { next_jlo = umin(next_jlo, j + 1); next_jhi = j + 1; }
			if (h > absopen) next_jlo = umin(next_jlo, j);
			if (h > absext && j == jhi && jhi + 1 < LB) { /* the match can be followed by an insert: widen this row */
				++jhi;
				const uint32_t ne = umax(umin(jhi + 1, LB), endj);
				for (uint32_t j2 = endj + 1; j2 <= ne; ++j2) {
					if (j2 - 1 > j) M[j2 - 1] = ORC_NEG_INF;
					Dr[j2] = ORC_NEG_INF;
				}
				endj = ne;
			}
			if (s >= best) { best = s; besti = i; bestj = j; }
			if (j != jlo) {
				const float md = saved + open;
				Dr[j] += ext;
				if (md >= Dr[j]) { Dr[j] = md; bits |= XB_MD; }
				h = Dr[j] - best + X;
				if (h > 0) { next_jlo = umin(next_jlo, j - 1); next_jhi = umax(next_jhi, j - 1); }
			}
			const float mi = saved + open;
			I0 += ext;
			if (mi >= I0) { I0 = mi; bits |= XB_MI; }
			h = I0 - best + X;
			if (h > 0) { next_jlo = umin(next_jlo, j + 1); next_jhi = umax(next_jhi, j + 1); }
			if (h > absext && j == jhi && jhi + 1 < LB) {
				++jhi;
				const uint32_t ne = umax(umin(jhi + 1, LB), endj);
				for (uint32_t j2 = endj + 1; j2 <= ne; ++j2) {
					M[j2 - 1] = ORC_NEG_INF;
					Dr[j2] = ORC_NEG_INF;
				}
				endj = ne;
			}
			row[j] = bits;
		}
		if (jhi < LB) { /* the D cell just right of the band */
			const uint32_t j1 = jhi + 1;
			row[j1] = 0;
			const float md = M0 + open;
			Dr[j1] += ext;
			if (md >= Dr[j1]) { Dr[j1] = md; row[j1] = XB_MD; }
		}
		if (next_jlo == XNONE)
			break;
		prev_jlo = jlo; prev_jhi = jhi;
		jlo = next_jlo; jhi = next_jhi;
		if (jlo > LB) jlo = LB;
		if (jhi > LB) jhi = LB;
		if (jlo == prev_jlo) { M0 = ORC_NEG_INF; Dr[jlo] = ORC_NEG_INF; }
		else M0 = M[jlo - 1];
	}
	float result = 0.0f;
	if (best > 0.0f) {
		result = best;
		uint32_t i = besti, j = bestj, n = 0;
		char st = 'M';
		char *rev = (char *)malloc((size_t)LA + LB + 4);
		for (;;) {
			rev[n++] = st;
			if (i == 1 || j == 1)
				break;
			char nx;
			if (st == 'M') {
				const uint8_t c = tb[(size_t)i * W + j];
				nx = (c & XB_DM) ? 'D' : (c & XB_IM) ? 'I' : 'M';
				--i; --j;
			} else if (st == 'D') {
				nx = (tb[(size_t)i * W + j + 1] & XB_MD) ? 'M' : 'D';
				--i;
			} else {
				nx = (tb[(size_t)(i + 1) * W + j] & XB_MI) ? 'M' : 'I';
				--j;
			}
			st = nx;
		}
		for (uint32_t k = 0; k < n; ++k)
			path[k] = rev[n - 1 - k];
		path[n] = 0;
		*path_len = n;
		free(rev);
	}
	free(Mbuf); free(Dr); free(tb);
	return result;
#undef SUB
}

int orc_do_mkf(const orc_params *p, const orc_chain *A, const orc_chain *B)
{
	if (!A->mu || !B->mu || A->L < 3 || B->L < 3)
		return 0;
	return A->L >= p->mkfl || B->L >= p->mkfl;
}

static void path_counts(const char *path, uint32_t n, uint32_t *M, uint32_t *D, uint32_t *I)
{
	*M = *D = *I = 0;
	for (uint32_t k = 0; k < n; ++k) {
		if (path[k] == 'M') ++*M;
		else if (path[k] == 'D') ++*D;
		else ++*I;
	}
}

void orc_align_mkf(const orc_params *p, const orc_chain *A, const orc_chain *B, orc_result *r, char *path,
		int *best_hsp_score, int *best_chain_score, float *xdrop_score)
{
	clear_result(r);
	path[0] = 0;
	*xdrop_score = 0;
	enum { CAP = 256 };
	orc_hsp hsps[CAP];
	int chain_idx[CAP], nhsp, nchain;
	orc_mkf_align(p, A->mu, (int)A->L, B->mu, (int)B->L, hsps, CAP, &nhsp, best_hsp_score, best_chain_score, chain_idx, &nchain);
	if (*best_chain_score <= 0)
		return;
	/* PostAlignMKF: dssaligner.cpp:1395-1437 */
	float mega_total = 0, best_mega = 0;
	int best_idx = 0;
	for (int k = 0; k < nchain; ++k) {
		const orc_hsp *h = &hsps[chain_idx[k]];
		const float ms = orc_mega_hsp_score(p, A->prof, A->L, B->prof, B->L, (uint32_t)h->loi, (uint32_t)h->loj, (uint32_t)h->len);
		if (ms > best_mega) { best_mega = ms; best_idx = k; }
		mega_total += ms;
	}
	if (mega_total < p->mkf_min_mega_hsp_score)
		return;
	/* XDropHSP: xdrophsp.cpp:42-150 */
	const orc_hsp *h = &hsps[chain_idx[best_idx]];
	const uint32_t loi_in = (uint32_t)h->loi, loj_in = (uint32_t)h->loj, len = (uint32_t)h->len;
	uint32_t LoA = loi_in + len / 2, LoB = loj_in + len / 2;
	const uint32_t K = 8;
	float *v = (float *)malloc(sizeof(float) * len);
	for (uint32_t c = 0; c < len; ++c)
		v[c] = subst(p, A->prof, A->L, B->prof, B->L, loi_in + c, loj_in + c);
	float best_mer = 0;
	for (uint32_t ms = 0; ms + K <= len; ++ms) {
		float sc = 0;
		for (uint32_t k = 0; k < K; ++k)
			sc += v[ms + k];
		if (sc > best_mer) { best_mer = sc; LoA = loi_in + ms; LoB = loj_in + ms; }
	}
	free(v);
	if ((LoA < LoB ? LoA : LoB) < K / 2) { LoA += K / 2; LoB += K / 2; }
	const float X = (float)p->mkf_x2;
	char *fwd = (char *)malloc((size_t)A->L + B->L + 4), *bwd = (char *)malloc((size_t)A->L + B->L + 4);
	uint32_t nf = 0, nb = 0;
	const float sf = orc_xdrop_fwd(p, A->prof, A->L, B->prof, B->L, 0, 0, 0, X, LoA, A->L, LoB, B->L, fwd, &nf);
	/* XDropBwd(HiA = LoA-1, HiB = LoB-1): forward DP on mirrored coordinates, then the path is reversed */
	const float sb = orc_xdrop_fwd(p, A->prof, A->L, B->prof, B->L, 1, LoA, LoB, X, 0, LoA, 0, LoB, bwd, &nb);
	for (uint32_t k = 0; k < nb / 2; ++k) { const char t = bwd[k]; bwd[k] = bwd[nb - 1 - k]; bwd[nb - 1 - k] = t; }
	const float total = sf + sb;
	if (total < 10) { /* xdrophsp.cpp:109-113: path cleared, score 0; Lo stays UINT_MAX and Hi = Lo + 0 - 1 wraps */
		free(fwd); free(bwd);
		r->score = 0;
		r->hi_a = r->lo_a - 1u;
		r->hi_b = r->lo_b - 1u;
		return;
	}
	/* MergeFwdBwd: mergefwdback.cpp:6-50 */
	uint32_t bM, bD, bI;
	path_counts(bwd, nb, &bM, &bD, &bI);
	r->lo_a = nb ? LoA - (bM + bD) : LoA;
	r->lo_b = nb ? LoB - (bM + bI) : LoB;
	memcpy(path, bwd, nb);
	memcpy(path + nb, fwd, nf);
	path[nb + nf] = 0;
	r->path_len = nb + nf;
	free(fwd); free(bwd);
	*xdrop_score = total;
	r->score = total;
	uint32_t nM, nD, nI;
	path_counts(path, r->path_len, &nM, &nD, &nI);
	r->hi_a = r->lo_a + nM + nD - 1; /* dssaligner.cpp:1432-1435 */
	r->hi_b = r->lo_b + nM + nI - 1;
	orc_calc_evalue(p, A, B, path, r);
}

void orc_align_pair(const orc_params *p, const orc_chain *A, const orc_chain *B, orc_result *r, char *path)
{
	if (orc_do_mkf(p, A, B)) { /* dssaligner.cpp:811-815 */
		int bh, bc;
		float xs;
		char *tmp = path ? path : (char *)malloc((size_t)A->L + B->L + 4);
		orc_align_mkf(p, A, B, r, tmp, &bh, &bc, &xs);
		if (!path)
			free(tmp);
		return;
	}
	clear_result(r);
	if (path)
		path[0] = 0;
	if (p->omega > 0 && A->mu && B->mu) {
		r->mu_score = orc_mu_filter_score(p, A->mu, A->L, B->mu, B->L, &r->mu_fwd, &r->mu_rev);
		if (r->mu_score < p->omega) {
			r->filtered = 1;
			return;
		}
	}
	char *tmp = path ? path : (char *)malloc((size_t)A->L + B->L + 2);
	r->score = orc_sw_align(p, A->prof, A->L, B->prof, B->L, &r->lo_a, &r->lo_b, tmp, &r->path_len);
	orc_calc_evalue(p, A, B, tmp, r);
	if (!path)
		free(tmp);
}

/* ------------------------------------------------------------------------------------------------
 * -global: DSSAligner::AlignQueryTarget_Global (global.cpp:7-33) = Mu filter, then ViterbiFastMem
 * (viterbifastmem.cpp:33-193) over the per-cell score of xdrophsp.cpp:8-33, traced back by TraceBackBitMem
 * (tracebackbitmem.cpp:8-69).  Three states with prefix-length indices: M[i][j] ends in a match of A[i-1], B[j-1];
 * D[i][j] in a deletion (A residue alone), I[i][j] in an insertion.  Gap parameters are file statics of the reference:
 * open -1, extend -0.05, terminal 0; the terminal values apply to column 0, to the column after the last B residue
 * (deletions only) and to the row after the last A residue (insertions only) - nowhere else, which is what the code does
 * rather than what the names suggest.  "Minus infinity" is the finite -9e9f (xdpmem.h:6).
 * ------------------------------------------------------------------------------------------------ */
#define VIT_NEG (-9e9f)
static const float vit_open = -1.0f, vit_ext = -0.05f, vit_topen = 0.0f, vit_text = 0.0f;

float orc_viterbi_global(const orc_params *p, const uint8_t *profA, uint32_t LA, const uint8_t *profB, uint32_t LB,
		char *path, uint32_t *path_len)
{
	*path_len = 0;
	if (path)
		path[0] = 0;
	if (LA == 0 || LB == 0)
		return 0; /* the reference reads Mrow[LB-1] out of range here; callers never pass empty chains */
	const size_t W = (size_t)LB + 1;
	uint8_t *tb = (uint8_t *)calloc((size_t)(LA + 1) * W, 1);
	float *m = (float *)malloc(sizeof(float) * W);   /* m[j] = M[i][j+1] while row i is open */
	float *d = (float *)malloc(sizeof(float) * W);   /* d[j] = D[i][j] */
	for (uint32_t j = 0; j <= LB; ++j)
		m[j] = d[j] = VIT_NEG;
	float mcorner = 0.0f; /* M[i][0]: 0 for the first row, "minus infinity" below */
	for (uint32_t i = 0; i < LA; ++i) {
		float open = vit_topen, ext = vit_text; /* column 0 only */
		float ins = VIT_NEG;                    /* I[i][j] */
		float mdiag = mcorner;                  /* M[i][j] */
		uint8_t *row = tb + (size_t)i * W;
		for (uint32_t j = 0; j < LB; ++j) {
			uint8_t bits = 0;
			const float mhere = mdiag;
			float best = mhere;
			if (d[j] > best) { best = d[j]; bits = XB_DM; }
			if (ins > best) { best = ins; bits = XB_IM; }
			mdiag = m[j];
			m[j] = best + subst(p, profA, LA, profB, LB, i, j);
			const float md = mhere + open;
			d[j] += ext;
			if (md >= d[j]) { d[j] = md; bits |= XB_MD; }
			const float mi = mhere + open;
			ins += ext;
			if (mi >= ins) { ins = mi; bits |= XB_MI; }
			open = vit_open;
			ext = vit_ext;
			row[j] = bits;
		}
		/* the column after the last B residue: deletions at the terminal price (:129-143) */
		row[LB] = 0;
		{
			const float md = mdiag + vit_topen;
			d[LB] += vit_text;
			if (md >= d[LB]) { d[LB] = md; row[LB] = XB_MD; }
		}
		mcorner = VIT_NEG;
	}
	/* the row after the last A residue: insertions at the terminal price, strict comparison (:151-167) */
	uint8_t *last = tb + (size_t)LA * W;
	float ins = VIT_NEG;
	for (uint32_t j = 1; j < LB; ++j) {
		last[j] = 0;
		const float mi = m[j - 1] + vit_topen;
		ins += vit_text;
		if (mi > ins) { ins = mi; last[j] = XB_MI; }
	}
	float score = m[LB - 1];
	char state = 'M';
	if (d[LB] > score) { score = d[LB]; state = 'D'; }
	if (ins > score) { score = ins; state = 'I'; }
	/* TraceBackBitMem */
	if (path) {
		size_t i = LA, j = LB;
		uint32_t n = 0;
		while (i != 0 || j != 0) {
			path[n++] = state;
			if (state == 'M') {
				const uint8_t t = tb[(i - 1) * W + (j - 1)];
				state = (t & XB_DM) ? 'D' : (t & XB_IM) ? 'I' : 'M';
				--i; --j;
			} else if (state == 'D') {
				const uint8_t t = tb[(i - 1) * W + j];
				state = (t & XB_MD) ? 'M' : 'D';
				--i;
			} else {
				const uint8_t t = tb[i * W + (j - 1)];
				state = (t & XB_MI) ? 'M' : 'I';
				--j;
			}
		}
		for (uint32_t a = 0, b = n ? n - 1 : 0; a < b; ++a, --b) {
			const char c = path[a]; path[a] = path[b]; path[b] = c;
		}
		path[n] = 0;
		*path_len = n;
	}
	free(tb); free(m); free(d);
	return score;
}

/* global.cpp:7-33.  r->score carries m_GlobalScore (-9999 after ClearAlign when the filter rejects), lo = 0 */
void orc_align_pair_global(const orc_params *p, const orc_chain *A, const orc_chain *B, orc_result *r, char *path)
{
	clear_result(r);
	r->score = -9999.0f;
	if (path)
		path[0] = 0;
	if (p->omega > 0 && A->mu && B->mu) {
		r->mu_score = orc_mu_filter_score(p, A->mu, A->L, B->mu, B->L, &r->mu_fwd, &r->mu_rev);
		if (r->mu_score < p->omega) {
			r->filtered = 1;
			return;
		}
	}
	r->score = orc_viterbi_global(p, A->prof, A->L, B->prof, B->L, path, &r->path_len);
	r->lo_a = 0;
	r->lo_b = 0;
}

/* ------------------------------------------------------------------------------------------------
 * -fast -db prefilter
 * ------------------------------------------------------------------------------------------------ */
static const int k5_off[5] = {0, 1, 2, 5, 6};

uint32_t orc_kmer5(const uint8_t *w)
{
	uint32_t k = 0;
	int self = 0;
	for (int c = 0; c < 5; ++c) {
		const uint8_t x = w[k5_off[c]];
		k = k * 36 + x;
		self += rsk_tbl_mu_kmer_i8[36 * x + x];
	}
	return self < 36 ? UINT32_MAX : k;
}

int orc_kmer5_pair_score(uint32_t k1, uint32_t k2)
{
	int s = 0;
	for (int c = 0; c < 5; ++c) {
		s += rsk_tbl_mu_kmer_i8[36 * (k1 % 36) + (k2 % 36)];
		k1 /= 36; k2 /= 36;
	}
	return s;
}

int orc_find_hsp(const uint8_t *q, uint32_t LQ, const uint8_t *t, uint32_t LT, int d)
{
	int i = (int)LQ - d - 1, j = 0;
	if (i < 0) { j = -i; i = 0; }
	int B = 0, F = 0;
	for (; i < (int)LQ && j < (int)LT; ++i, ++j) {
		F += rsk_tbl_mu_kmer_i8[36 * q[i] + t[j]];
		if (F > B) B = F;
		else if (F < 0) F = 0;
	}
	return B;
}

void orc_prefilter_target(const uint8_t *const *muQ, const uint32_t *LQ, uint32_t nQ, const uint8_t *muT, uint32_t LT,
		int query_neighborhood, uint16_t *best)
{
	for (uint32_t q = 0; q < nQ; ++q)
		best[q] = 0;
	if (LT < 7)
		return;
	const uint32_t nkt = LT - 6;
	uint32_t *kt = (uint32_t *)malloc(sizeof(uint32_t) * nkt);
	for (uint32_t j = 0; j < nkt; ++j)
		kt[j] = orc_kmer5(muT + j);
	for (uint32_t q = 0; q < nQ; ++q) {
		const uint32_t lq = LQ[q];
		if (lq < 7)
			continue;
		const uint32_t nkq = lq - 6;
		uint32_t *kq = (uint32_t *)malloc(sizeof(uint32_t) * nkq);
		for (uint32_t i = 0; i < nkq; ++i)
			kq[i] = orc_kmer5(muQ[q] + i);
		int bestq = 0;
		/* k-mer start positions (i, j) on diagonal d = lq + j - i - 1 */
		for (int d = 0; d <= (int)(lq + LT) - 2 && d <= 16383; ++d) { /* diag > 0x3fff is dropped (prefiltermu.cpp:254) */
			int i = (int)lq - d - 1, j = 0;
			if (i < 0) { j = -i; i = 0; }
			int seeds = 0;
			for (; i < (int)nkq && j < (int)nkt; ++i, ++j) {
				if (kq[i] == UINT32_MAX || kt[j] == UINT32_MAX)
					continue;
				if (orc_kmer5_pair_score(kq[i], kt[j]) >= 36)
					seeds += (query_neighborhood && kq[i] == kt[j]) ? 2 : 1;
			}
			if (seeds >= 2) {
				const int sc = orc_find_hsp(muQ[q], lq, muT, LT, d);
				if (sc > bestq)
					bestq = sc;
			}
		}
		free(kq);
		if (bestq >= 65535)
			bestq = 65534;
		best[q] = (uint16_t)bestq;
	}
	free(kt);
}

struct orc_rsb {
	uint32_t nQ, B;
	uint32_t *n, *cap;
	uint32_t **t;
	uint16_t **s;
	uint16_t *lo;
};

orc_rsb *orc_rsb_new(uint32_t nQ, uint32_t B)
{
	orc_rsb *r = (orc_rsb *)calloc(1, sizeof(orc_rsb));
	r->nQ = nQ; r->B = B;
	r->n = (uint32_t *)calloc(nQ, sizeof(uint32_t));
	r->cap = (uint32_t *)calloc(nQ, sizeof(uint32_t));
	r->t = (uint32_t **)calloc(nQ, sizeof(uint32_t *));
	r->s = (uint16_t **)calloc(nQ, sizeof(uint16_t *));
	r->lo = (uint16_t *)calloc(nQ, sizeof(uint16_t));
	return r;
}

/* sort.h:71-108: Hoare partition around the middle element, descending, on an index array */
static void order_desc(const uint16_t *v, int left, int right, uint32_t *order)
{
	int i = left, j = right;
	const uint16_t pivot = v[order[(left + right) / 2]];
	while (i <= j) {
		while (v[order[i]] > pivot) i++;
		while (v[order[j]] < pivot) j--;
		if (i <= j) {
			const uint32_t tmp = order[i]; order[i] = order[j]; order[j] = tmp;
			i++; j--;
		}
	}
	if (left < j) order_desc(v, left, j, order);
	if (i < right) order_desc(v, i, right, order);
}

static void rsb_truncate(orc_rsb *r, uint32_t q)
{
	const uint32_t n = r->n[q];
	if (n < r->B)
		return;
	uint32_t *order = (uint32_t *)malloc(sizeof(uint32_t) * n);
	for (uint32_t i = 0; i < n; ++i) order[i] = i;
	order_desc(r->s[q], 0, (int)n - 1, order);
	uint32_t *nt = (uint32_t *)malloc(sizeof(uint32_t) * r->cap[q]);
	uint16_t *ns = (uint16_t *)malloc(sizeof(uint16_t) * r->cap[q]);
	for (uint32_t k = 0; k < r->B; ++k) { nt[k] = r->t[q][order[k]]; ns[k] = r->s[q][order[k]]; }
	free(r->t[q]); free(r->s[q]); free(order);
	r->t[q] = nt; r->s[q] = ns;
	r->n[q] = r->B;
	r->lo[q] = ns[r->B - 1];
}

void orc_rsb_add(orc_rsb *r, uint32_t q, uint32_t t, uint16_t score)
{
	if (score < r->lo[q])
		return;
	if (r->n[q] == r->cap[q]) {
		r->cap[q] = r->cap[q] ? 2 * r->cap[q] : 64;
		r->t[q] = (uint32_t *)realloc(r->t[q], sizeof(uint32_t) * r->cap[q]);
		r->s[q] = (uint16_t *)realloc(r->s[q], sizeof(uint16_t) * r->cap[q]);
	}
	r->t[q][r->n[q]] = t;
	r->s[q][r->n[q]] = score;
	r->n[q]++;
	if (r->n[q] >= 2 * r->B)
		rsb_truncate(r, q);
}

void orc_rsb_finish(orc_rsb *r)
{
	for (uint32_t q = 0; q < r->nQ; ++q)
		rsb_truncate(r, q);
}

uint32_t orc_rsb_count(const orc_rsb *r, uint32_t q) { return r->n[q]; }
const uint32_t *orc_rsb_targets(const orc_rsb *r, uint32_t q) { return r->t[q]; }
const uint16_t *orc_rsb_scores(const orc_rsb *r, uint32_t q) { return r->s[q]; }

void orc_rsb_free(orc_rsb *r)
{
	for (uint32_t q = 0; q < r->nQ; ++q) { free(r->t[q]); free(r->s[q]); }
	free(r->n); free(r->cap); free(r->t); free(r->s); free(r->lo); free(r);
}

void orc_align_pairs(const orc_params *p, const orc_chain *chainsA, const orc_chain *chainsB,
		const uint32_t *ia, const uint32_t *ib, size_t npairs, orc_result *out)
{
	for (size_t k = 0; k < npairs; ++k)
		orc_align_pair(p, &chainsA[ia[k]], &chainsB[ib[k]], &out[k], NULL);
}
