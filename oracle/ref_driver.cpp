// ref_driver.cpp - thin C-ABI shim over the UNMODIFIED reference objects (oracle/_ref/libreseek_ref.so).
//
// TEST INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile against /root/reference/src (headers via -iquote),
// linked with every reference object except reseek_main.o.  It lets tests/ and tools/make_golden.py drive
// the reference's own DSSAligner / DSS / BCAData / parasail / SWFast code from Python (ctypes) so that the
// plain-C oracle (oracle/reseek_oracle.c) and the CUDA library are pinned against the real thing.
// Nothing here is product code and nothing in the product links against it.
#include "myutils.h"
#include "dss.h"
#include "dssaligner.h"
#include "dssparams.h"
#include "bcadata.h"
#include "pdbchain.h"
#include "xdpmem.h"
#include "statsig.h"
#include "mumx.h"
#include "parasail.h"
#include <thread>
#include <atomic>

// normally defined in reseek_main.cpp (which holds main() and is left out of the library)
int g_Frame = 0;
string g_Arg1;

float GetSelfRevScore(DSSAligner &DA, DSS &D, const PDBChain &Chain,
  const vector<vector<byte> > &Profile, const vector<byte> *ptrMuLetters,
  const vector<uint> *ptrMuKmers);
double GetLDDT_mu_fast(const PDBChain &Q, const PDBChain &T,
  const vector<uint> &PosQs, const vector<uint> &PosTs);

float SWFastGaplessProfb(float *DProw_, const float * const *ProfA, uint LA, const byte *B, uint LB);
uint SWFastPinopGapless(const int8_t * const *AP, uint LA, const int8_t *B, uint LB);

namespace {
DSSParams *g_Params = 0;      // search params (RunQuery / RunSelf / PostMuFilter view)
DSSParams *g_LoadParams = 0;  // ProfileLoader view: Omega=0, UsePara=false (profileloader.cpp:22-26)
DSSAligner *g_DA = 0;
DSSAligner *g_LoadDA = 0;
DSS *g_DSS = 0;
BCAData *g_BCA = 0;

void MakeChain(PDBChain &C, const char *label, uint L, const char *seq,
  const float *x, const float *y, const float *z)
	{
	C.m_Label = label ? label : "chain";
	C.m_Seq.assign(L, 'A');
	if (seq)
		for (uint i = 0; i < L; ++i)
			C.m_Seq[i] = seq[i];
	C.m_Xs.assign(x, x + L);
	C.m_Ys.assign(y, y + L);
	C.m_Zs.assign(z, z + L);
	}

void MakeProfile(vector<vector<byte> > &P, uint L, const uint8_t *prof)
	{
	P.resize(8);
	for (uint f = 0; f < 8; ++f)
		P[f].assign(prof + size_t(f)*L, prof + size_t(f + 1)*L);
	}
}

struct ref_result
	{
	float score;
	uint32_t lo_a, lo_b, hi_a, hi_b, ids, gaps;
	float lddt; // recomputed via GetLDDT() when a path exists and CalcEvalue ran
	float ts, pvalue, evalue, qual;
	float mu_score; // GetMuScore() when the filter applies, else 0
	int32_t mkf; // DoMKF() was true
	int32_t best_hsp_score, best_chain_score; // m_MKF fields
	float xdrop_score;
	uint32_t path_len;
	};

// kabsch.cpp:330 (C++ linkage)
double Kabsch(const PDBChain &ChainA, const PDBChain &ChainB, uint LoA, uint LoB, const string &Path,
  double t[3], double u[3][3]);

extern "C" {

// mode: 1 fast, 2 sensitive, 3 verysensitive (dssparams.cpp:52-81)
int ref_init(int mode)
	{
	DECIDE_MODE DM = mode == 1 ? DM_AlwaysFast : mode == 2 ? DM_AlwaysSensitive : DM_AlwaysVerysensitive;
	g_Params = new DSSParams;
	g_Params->SetDSSParams(DM);
	g_LoadParams = new DSSParams;
	*g_LoadParams = *g_Params;
	g_LoadParams->m_UsePara = false;
	g_LoadParams->m_Omega = 0;
	g_LoadParams->m_OwnScoreMxs = false;
	g_DA = new DSSAligner;
	g_DA->SetParams(*g_Params);
	g_LoadDA = new DSSAligner;
	g_LoadDA->SetParams(*g_LoadParams);
	g_DSS = new DSS;
	g_DSS->SetParams(*g_Params);
	return 0;
	}

// scalars: gap_open, gap_ext, min_fwd_score, omega, omega_fwd, mkfl, x1, x2, minhsp, minmegahsp, mu open, mu ext
void ref_get_params(float *scalars12, float *tables2192)
	{
	const DSSParams &P = *g_Params;
	scalars12[0] = P.m_GapOpen; scalars12[1] = P.m_GapExt; scalars12[2] = P.m_MinFwdScore;
	scalars12[3] = P.m_Omega; scalars12[4] = P.m_OmegaFwd; scalars12[5] = (float) P.m_MKFL;
	scalars12[6] = (float) P.m_MKF_X1; scalars12[7] = (float) P.m_MKF_X2;
	scalars12[8] = (float) P.m_MKF_MinHSPScore; scalars12[9] = P.m_MKF_MinMegaHSPScore;
	scalars12[10] = (float) P.m_ParaMuGapOpen; scalars12[11] = (float) P.m_ParaMuGapExt;
	uint k = 0;
	for (uint Idx = 0; Idx < P.GetFeatureCount(); ++Idx)
		{
		FEATURE F = P.m_Features[Idx];
		uint AS = g_AlphaSizes2[F];
		for (uint a = 0; a < AS; ++a)
			for (uint b = 0; b < AS; ++b)
				tables2192[k++] = P.m_ScoreMxs[F][a][b];
		}
	}

void ref_get_mu_matrices(float *f32_1296, int8_t *i8_1296, int8_t *kmer_i8_1296)
	{
	for (uint a = 0; a < 36; ++a)
		for (uint b = 0; b < 36; ++b)
			{
			f32_1296[a*36 + b] = ScoreMx_Mu[a][b];
			i8_1296[a*36 + b] = IntScoreMx_Mu[a][b];
			kmer_i8_1296[a*36 + b] = Mu_S_ij_i8[a][b];
			}
	}

// ---- .bca access (bcadata.cpp) ----
int ref_bca_open(const char *fn)
	{
	delete g_BCA;
	g_BCA = new BCAData;
	g_BCA->Open(fn);
	return (int) g_BCA->GetChainCount();
	}

int ref_bca_len(int idx) { return (int) g_BCA->GetSeqLength(idx); }

int ref_bca_chain(int idx, char *label, int label_cap, char *seq, float *x, float *y, float *z)
	{
	PDBChain C;
	g_BCA->ReadChain(idx, C);
	uint L = C.GetSeqLength();
	snprintf(label, label_cap, "%s", C.m_Label.c_str());
	for (uint i = 0; i < L; ++i)
		{
		seq[i] = C.m_Seq[i];
		x[i] = C.m_Xs[i]; y[i] = C.m_Ys[i]; z[i] = C.m_Zs[i];
		}
	return (int) L;
	}

// ---- DSS feature extraction (dss.cpp:716 GetProfile, :700 GetMuLetters, :659 GetMuKmers) ----
// prof: [8][L]; mu: [L]; kmers: [L] capacity, returns count in *nk
int ref_dss(uint L, const char *seq, const float *x, const float *y, const float *z,
  uint8_t *prof, uint8_t *mu, uint32_t *kmers, uint32_t *nk)
	{
	PDBChain C;
	MakeChain(C, "c", L, seq, x, y, z);
	vector<vector<byte> > P;
	vector<byte> Mu;
	vector<uint> K;
	g_DSS->Init(C);
	g_DSS->GetProfile(P);
	g_DSS->GetMuLetters(Mu);
	g_DSS->GetMuKmers(Mu, K, g_Params->m_MKFPatternStr);
	for (uint f = 0; f < 8; ++f)
		memcpy(prof + size_t(f)*L, P[f].data(), L);
	memcpy(mu, Mu.data(), L);
	*nk = SIZE(K);
	for (uint i = 0; i < SIZE(K); ++i)
		kmers[i] = K[i];
	return 0;
	}

// Self-reverse score (alignpair.cpp:7-25).  loader != 0: ProfileLoader semantics (Omega=0, no parasail);
// loader == 0: RunQuery semantics (full search params).  with_mu: pass Mu letters+kmers as the callers do.
float ref_selfrev(uint L, const char *seq, const float *x, const float *y, const float *z, int loader, int with_mu)
	{
	PDBChain C;
	MakeChain(C, "c", L, seq, x, y, z);
	vector<vector<byte> > P;
	vector<byte> Mu;
	vector<uint> K;
	g_DSS->Init(C);
	g_DSS->GetProfile(P);
	g_DSS->GetMuLetters(Mu);
	g_DSS->GetMuKmers(Mu, K, g_Params->m_MKFPatternStr);
	DSSAligner &DA = loader ? *g_LoadDA : *g_DA;
	const vector<byte> *pMu = (with_mu && !Mu.empty()) ? &Mu : 0;
	const vector<uint> *pK = (with_mu && !K.empty()) ? &K : 0;
	float s = GetSelfRevScore(DA, *g_DSS, C, P, pMu, pK);
	DA.UnsetQuery();
	return s;
	}

// Reversed-chain profile used by the self-reverse alignment (pdbchain.cpp:478 GetReverse + DSS)
int ref_rev_profile(uint L, const char *seq, const float *x, const float *y, const float *z, uint8_t *prof)
	{
	PDBChain C, R;
	MakeChain(C, "c", L, seq, x, y, z);
	C.GetReverse(R);
	vector<vector<byte> > P;
	g_DSS->Init(R);
	g_DSS->GetProfile(P);
	for (uint f = 0; f < 8; ++f)
		memcpy(prof + size_t(f)*L, P[f].data(), L);
	return 0;
	}

// ---- the per-pair aligner, driven exactly like runquery.cpp:45,70-71 ----
// mu / kmers may be NULL (then the Mu filter and MKF are skipped as in the reference).
// noaccel == 1 calls Align_NoAccel() directly (alignpair / self-rev style); noaccel == 2 calls
// AlignQueryTarget_Global() and reports m_GlobalScore as the score.
int ref_align_pair(
  uint LA, const uint8_t *profA, const uint8_t *muA, const uint32_t *kmA, uint nkA,
  const float *xA, const float *yA, const float *zA, float selfrevA,
  uint LB, const uint8_t *profB, const uint8_t *muB, const uint32_t *kmB, uint nkB,
  const float *xB, const float *yB, const float *zB, float selfrevB,
  int noaccel, ref_result *out, char *path, uint path_cap)
	{
	PDBChain CA, CB;
	MakeChain(CA, "A", LA, 0, xA, yA, zA);
	MakeChain(CB, "B", LB, 0, xB, yB, zB);
	vector<vector<byte> > PA, PB;
	MakeProfile(PA, LA, profA);
	MakeProfile(PB, LB, profB);
	vector<byte> MuA, MuB;
	vector<uint> KA, KB;
	if (muA) MuA.assign(muA, muA + LA);
	if (muB) MuB.assign(muB, muB + LB);
	if (kmA) KA.assign(kmA, kmA + nkA);
	if (kmB) KB.assign(kmB, kmB + nkB);
	DSSAligner &DA = *g_DA;
	DA.SetQuery(CA, &PA, muA ? &MuA : 0, kmA ? &KA : 0, selfrevA);
	DA.SetTarget(CB, &PB, muB ? &MuB : 0, kmB ? &KB : 0, selfrevB);
	memset(out, 0, sizeof(*out));
	out->mkf = DA.DoMKF() ? 1 : 0;
	if (noaccel == 2)
		DA.AlignQueryTarget_Global();  // global.cpp:7-33 (-global)
	else if (noaccel)
		DA.Align_NoAccel();
	else
		DA.AlignQueryTarget();
	out->score = noaccel == 2 ? DA.m_GlobalScore : DA.m_AlnFwdScore;
	out->lo_a = DA.m_LoA; out->lo_b = DA.m_LoB; out->hi_a = DA.m_HiA; out->hi_b = DA.m_HiB;
	out->ids = DA.m_Ids; out->gaps = DA.m_Gaps;
	out->ts = DA.m_NewTestStatisticA; out->pvalue = DA.m_PvalueA; out->evalue = DA.m_EvalueA;
	out->qual = DA.m_QualityA;
	out->xdrop_score = DA.m_XDropScore;
	out->best_hsp_score = DA.m_MKF.m_BestHSPScore;
	out->best_chain_score = DA.m_MKF.m_BestChainScore;
	out->path_len = SIZE(DA.m_Path);
	out->lddt = 0;
	if (!DA.m_Path.empty() && DA.m_EvalueA != FLT_MAX)
		out->lddt = DA.GetLDDT();
	if (path && path_cap > 0)
		snprintf(path, path_cap, "%s", DA.m_Path.c_str());
	out->mu_score = 0;
	if (!noaccel && !out->mkf && muA && muB && g_Params->m_Omega > 0)
		out->mu_score = DA.GetMuScore();
	DA.UnsetQuery();
	return 0;
	}

// AlignQueryTarget followed by the reference's own TSV writer (DSSAligner::ToTsv, userfields.cpp:45-152).
// Returns the number of bytes written to out (0 when the pair has no alignment).
int ref_align_pair_tsv(
  uint LA, const char *labelA, const char *seqA, const uint8_t *profA, const uint8_t *muA, const uint32_t *kmA, uint nkA,
  const float *xA, const float *yA, const float *zA, float selfrevA,
  uint LB, const char *labelB, const char *seqB, const uint8_t *profB, const uint8_t *muB, const uint32_t *kmB, uint nkB,
  const float *xB, const float *yB, const float *zB, float selfrevB,
  const char *columns, int up, char *out, uint cap)
	{
	PDBChain CA, CB;
	MakeChain(CA, labelA, LA, seqA, xA, yA, zA);
	MakeChain(CB, labelB, LB, seqB, xB, yB, zB);
	vector<vector<byte> > PA, PB;
	MakeProfile(PA, LA, profA);
	MakeProfile(PB, LB, profB);
	vector<byte> MuA, MuB;
	vector<uint> KA, KB;
	if (muA) MuA.assign(muA, muA + LA);
	if (muB) MuB.assign(muB, muB + LB);
	if (kmA) KA.assign(kmA, kmA + nkA);
	if (kmB) KB.assign(kmB, kmB + nkB);
	DSSAligner &DA = *g_DA;
	DA.m_UFs.clear();
	vector<string> Fields;
	Split(string(columns), Fields, '+');
	for (uint i = 0; i < SIZE(Fields); ++i)
		DA.m_UFs.push_back(StrToUF(Fields[i]));
	DA.SetQuery(CA, &PA, muA ? &MuA : 0, kmA ? &KA : 0, selfrevA);
	DA.SetTarget(CB, &PB, muB ? &MuB : 0, kmB ? &KB : 0, selfrevB);
	DA.AlignQueryTarget();
	out[0] = 0;
	uint n = 0;
	if (!DA.m_Path.empty() && DA.m_EvalueA != FLT_MAX)
		{
		char *buf = 0;
		size_t sz = 0;
		FILE *f = open_memstream(&buf, &sz);
		DA.ToTsv(f, up != 0);
		fclose(f);
		n = (uint) min(sz, (size_t) cap - 1);
		memcpy(out, buf, n);
		out[n] = 0;
		free(buf);
		}
	DA.UnsetQuery();
	return (int) n;
	}

// Mu filter score exactly as DSSAligner::GetMuScore() under the search params (parasail_mu.cpp:120-161)
float ref_mu_score(uint LA, const uint8_t *muA, uint LB, const uint8_t *muB)
	{
	PDBChain CA, CB;
	vector<float> z(max(LA, LB), 0.0f);
	MakeChain(CA, "A", LA, 0, z.data(), z.data(), z.data());
	MakeChain(CB, "B", LB, 0, z.data(), z.data(), z.data());
	vector<byte> MuA(muA, muA + LA), MuB(muB, muB + LB);
	DSSAligner &DA = *g_DA;
	DA.SetQuery(CA, 0, &MuA, 0, FLT_MAX);
	DA.SetTarget(CB, 0, &MuB, 0, FLT_MAX);
	float s = DA.GetMuScore();
	DA.UnsetQuery();
	return s;
	}

// gapless alternatives (declared in dssaligner.cpp:18-32; unreachable from the CLI)
float ref_gapless_profb(uint LA, const uint8_t *muA, uint LB, const uint8_t *muB)
	{
	vector<const float *> Prof(LA);
	for (uint i = 0; i < LA; ++i)
		Prof[i] = ScoreMx_Mu[muA[i]];
	vector<float> DProw(2*LB + 8);
	return SWFastGaplessProfb(DProw.data(), Prof.data(), LA, muB, LB);
	}
int ref_gapless_int(uint LA, const uint8_t *muA, uint LB, const uint8_t *muB)
	{
	vector<const int8_t *> AP(LA);
	for (uint i = 0; i < LA; ++i)
		AP[i] = IntScoreMx_Mu[muA[i]];
	return (int) SWFastPinopGapless(AP.data(), LA, (const int8_t *) muB, LB);
	}

// raw parasail int8 striped SW (parasail.cpp:515): returns score, *sat = saturated flag
int ref_parasail_sw(uint LA, const uint8_t *muA, uint LB, const uint8_t *muB, int open, int ext, int *sat)
	{
	extern parasail_matrix_t parasail_mu_matrix;
	parasail_profile_t *prof = parasail_profile_create_avx_256_8((const char *) muA, (int) LA, &parasail_mu_matrix);
	parasail_result_t *r = parasail_sw_striped_profile_avx2_256_8(prof, (const char *) muB, (int) LB, open, ext);
	int score = r->score;
	*sat = (r->flag & PARASAIL_FLAG_SATURATED) ? 1 : 0;
	parasail_result_free(r);
	parasail_profile_free(prof);
	return score;
	}

// sw.cpp:79 on an explicit score matrix
float ref_swfast(const float *S, uint LA, uint LB, float open, float ext,
  uint32_t *lo_a, uint32_t *lo_b, char *path, uint path_cap)
	{
	vector<const float *> rows(LA);
	for (uint i = 0; i < LA; ++i)
		rows[i] = S + size_t(i)*LB;
	XDPMem Mem;
	uint Loi = UINT_MAX, Loj = UINT_MAX, Leni = 0, Lenj = 0;
	string Path;
	float score = SWFast(Mem, rows.data(), LA, LB, open, ext, Loi, Loj, Leni, Lenj, Path);
	*lo_a = Loi; *lo_b = Loj;
	snprintf(path, path_cap, "%s", Path.c_str());
	return score;
	}

double ref_lddt(uint LA, const float *xA, const float *yA, const float *zA,
  uint LB, const float *xB, const float *yB, const float *zB,
  const uint32_t *posA, const uint32_t *posB, uint n)
	{
	PDBChain CA, CB;
	MakeChain(CA, "A", LA, 0, xA, yA, zA);
	MakeChain(CB, "B", LB, 0, xB, yB, zB);
	vector<uint> PA(posA, posA + n), PB(posB, posB + n);
	return GetLDDT_mu_fast(CA, CB, PA, PB);
	}

// Multi-threaded batch for CPU-baseline timing: the reference's own per-pair loop (one DSSAligner per thread,
// SetQuery when the A chain changes, SetTarget + AlignQueryTarget per pair - runquery.cpp:45,70-71), std::thread
// over contiguous pair ranges like dbsearcher's thread pool.  Chains are SoA: len[], prof [8][total] plane-major,
// mu [total] (may be NULL), xyz [3][total], selfrev[].  out_score / out_evalue: per pair.
// The reference's superposition (kabsch.cpp:330-387 over :21-327): rotation u, translation t mapping the aligned A residues
// onto their B partners; returns the residual sum of squares divided by the number of M columns.
double ref_kabsch(uint LA, const float *xA, const float *yA, const float *zA,
  uint LB, const float *xB, const float *yB, const float *zB,
  uint LoA, uint LoB, const char *path, double *t, double *u)
	{
	PDBChain CA, CB;
	MakeChain(CA, "A", LA, 0, xA, yA, zA);
	MakeChain(CB, "B", LB, 0, xB, yB, zB);
	double uu[3][3];
	double r = Kabsch(CA, CB, LoA, LoB, string(path), t, uu);
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j)
			u[3*i + j] = uu[i][j];
	return r;
	}

int ref_align_batch(int nthreads,
  uint nA, const uint32_t *lenA, const uint8_t *profA, const uint8_t *muA, const float *xyzA, const float *selfrevA,
  uint nB, const uint32_t *lenB, const uint8_t *profB, const uint8_t *muB, const float *xyzB, const float *selfrevB,
  uint64_t npairs, const uint32_t *ia, const uint32_t *ib, float *out_score, float *out_evalue, uint32_t *out_pathlen)
	{
	struct Set
		{
		vector<PDBChain *> Chains;
		vector<vector<vector<byte> > > Profs;
		vector<vector<byte> > Mus;
		};
	auto Build = [](Set &S, uint n, const uint32_t *len, const uint8_t *prof, const uint8_t *mu, const float *xyz)
		{
		uint64_t total = 0;
		for (uint i = 0; i < n; ++i) total += len[i];
		uint64_t off = 0;
		S.Profs.resize(n);
		S.Mus.resize(n);
		for (uint i = 0; i < n; ++i)
			{
			uint L = len[i];
			PDBChain *C = new PDBChain;
			MakeChain(*C, "c", L, 0, xyz + off, xyz + total + off, xyz + 2*total + off);
			S.Chains.push_back(C);
			S.Profs[i].resize(8);
			for (uint f = 0; f < 8; ++f)
				S.Profs[i][f].assign(prof + f*total + off, prof + f*total + off + L);
			if (mu)
				S.Mus[i].assign(mu + off, mu + off + L);
			off += L;
			}
		};
	Set SA, SB;
	Build(SA, nA, lenA, profA, muA, xyzA);
	Build(SB, nB, lenB, profB, muB, xyzB);
	if (nthreads < 1) nthreads = 1;
	vector<thread> ts;
	for (int t = 0; t < nthreads; ++t)
		{
		ts.emplace_back([&, t]()
			{
			DSSAligner DA;
			DA.SetParams(*g_Params);
			uint64_t k0 = npairs*t/nthreads, k1 = npairs*(t + 1)/nthreads;
			uint32_t cur = UINT_MAX;
			for (uint64_t k = k0; k < k1; ++k)
				{
				uint a = ia[k], b = ib[k];
				if (a != cur)
					{
					DA.SetQuery(*SA.Chains[a], &SA.Profs[a], muA ? &SA.Mus[a] : 0, 0, selfrevA ? selfrevA[a] : FLT_MAX);
					cur = a;
					}
				DA.SetTarget(*SB.Chains[b], &SB.Profs[b], muB ? &SB.Mus[b] : 0, 0, selfrevB ? selfrevB[b] : FLT_MAX);
				DA.AlignQueryTarget();
				if (out_score) out_score[k] = DA.m_AlnFwdScore;
				if (out_evalue) out_evalue[k] = DA.m_EvalueA;
				if (out_pathlen) out_pathlen[k] = SIZE(DA.m_Path);
				}
			DA.UnsetQuery();
			});
		}
	for (auto &t : ts) t.join();
	for (auto *C : SA.Chains) delete C;
	for (auto *C : SB.Chains) delete C;
	return 0;
	}

void ref_statsig(double ts, double *p, double *e, double *q)
	{
	*p = StatSig::GetPvalue(ts);
	*e = StatSig::GetEvalue(ts);
	*q = StatSig::GetQual(ts);
	}

} // extern "C"
