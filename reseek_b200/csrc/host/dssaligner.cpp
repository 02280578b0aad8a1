// dssaligner.cpp - DSSAligner / DSSParams look-alikes over the C ABI (see dssaligner.h).
#include "dssaligner.h"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

namespace reseek_b200 {

// ---- myutils.cpp:785-824 ----
void Die(const char *Format, ...)
	{
	char Msg[4096];
	va_list ArgList;
	va_start(ArgList, Format);
	vsnprintf(Msg, sizeof(Msg), Format, ArgList);
	va_end(ArgList);
	fprintf(stderr, "\n\n\n---Fatal error---\n%s\n", Msg);
	exit(1);
	}

void Warning(const char *Format, ...)
	{
	char Msg[4096];
	va_list ArgList;
	va_start(ArgList, Format);
	vsnprintf(Msg, sizeof(Msg), Format, ArgList);
	va_end(ArgList);
	fprintf(stderr, "\nWARNING: %s\n", Msg);
	}

static void Check(int rc)
	{
	if (rc != RSK_OK)
		Die("reseek_b200: %s", rsk_last_error());
	}

// ---- pdbchain.cpp:478-483 (+ Reverse) ----
void PDBChain::GetReverse(PDBChain &Rev) const
	{
	Rev = *this;
	std::reverse(Rev.m_Seq.begin(), Rev.m_Seq.end());
	std::reverse(Rev.m_Xs.begin(), Rev.m_Xs.end());
	std::reverse(Rev.m_Ys.begin(), Rev.m_Ys.end());
	std::reverse(Rev.m_Zs.begin(), Rev.m_Zs.end());
	Rev.m_Label += ".rev";
	}

// ---- dssparams.cpp:44-111 ----
void DSSParams::SetMode(ALGO_MODE AM)
	{
	int Mode = 0;
	switch (AM)
		{
	case AM_Fast: Mode = RSK_MODE_FAST; break;
	case AM_Sensitive: Mode = RSK_MODE_SENSITIVE; break;
	case AM_VerySensitive: Mode = RSK_MODE_VERYSENSITIVE; break;
	default: Die("Must set -fast, -sensitive or -verysensitive");  // dssparams.cpp:90
		}
	rsk_params R;
	Check(rsk_params_preset(&R, Mode));
	m_Mode = AM;
	m_Weights.assign(R.weights, R.weights + RSK_NFEAT);
	m_GapOpen = R.gap_open;
	m_GapExt = R.gap_ext;
	m_MinFwdScore = R.min_fwd_score;
	m_Omega = R.omega;
	m_OmegaFwd = R.omega_fwd;
	m_ParaMuGapOpen = R.mu_gap_open;
	m_ParaMuGapExt = R.mu_gap_ext;
	m_MKFL = R.mkfl;
	m_MKF_X1 = R.mkf_x1;
	m_MKF_X2 = R.mkf_x2;
	m_MKF_MinHSPScore = R.mkf_min_hsp_score;
	m_MKF_MinMegaHSPScore = R.mkf_min_mega_hsp_score;
	memcpy(m_Tables, R.tables, sizeof(m_Tables));
	}

void DSSParams::SetDSSParams(DECIDE_MODE DM)
	{
	switch (DM)
		{
	case DM_AlwaysFast: SetMode(AM_Fast); return;
	case DM_AlwaysSensitive: SetMode(AM_Sensitive); return;
	case DM_AlwaysVerysensitive: SetMode(AM_VerySensitive); return;
	case DM_DefaultFast: SetMode(AM_Fast); return;
	case DM_DefaultSensitive: SetMode(AM_Sensitive); return;
	default:
		Die("SetDSSParams(DM=%d): the command line is not visible to this layer, call SetMode(AM_*)", (int)DM);
		}
	}

void DSSParams::ToRsk(rsk_params &R, double MaxEvalue) const
	{
	memset(&R, 0, sizeof(R));
	if (m_GapOpen > 0 || m_GapExt > 0)
		Die("open=%.3g ext=%.3g, gap penalties must be >= 0", -m_GapOpen, -m_GapExt);  // dssparams.cpp:106-108
	R.gap_open = m_GapOpen;
	R.gap_ext = m_GapExt;
	R.min_fwd_score = m_MinFwdScore;
	R.omega = m_Omega;
	R.omega_fwd = m_OmegaFwd;
	R.mu_gap_open = m_ParaMuGapOpen;
	R.mu_gap_ext = m_ParaMuGapExt;
	R.mkfl = m_MKFL;
	R.mkf_x1 = m_MKF_X1;
	R.mkf_x2 = m_MKF_X2;
	R.mkf_min_hsp_score = m_MKF_MinHSPScore;
	R.mkf_min_mega_hsp_score = m_MKF_MinMegaHSPScore;
	R.max_evalue = MaxEvalue;
	for (uint i = 0; i < RSK_NFEAT && i < RSK_SIZE(m_Weights); ++i)
		R.weights[i] = m_Weights[i];
	memcpy(R.tables, m_Tables, sizeof(m_Tables));
	}

// ---- chains -> one SoA upload ----
rsk_chainset *UploadChains(rsk_ctx *Ctx, const vector<ChainData> &Chains, bool WithMu)
	{
	const uint N = RSK_SIZE(Chains);
	rsk_asserta(N > 0);
	vector<uint32_t> Len(N);
	uint64_t Total = 0;
	for (uint i = 0; i < N; ++i)
		{
		rsk_asserta(Chains[i].Chain != 0 && Chains[i].Profile != 0);
		Len[i] = Chains[i].Chain->GetSeqLength();
		Total += Len[i];
		if (WithMu && Chains[i].MuLetters == 0)
			WithMu = false;
		}
	vector<uint8_t> Prof((size_t)RSK_NFEAT * Total), Mu(WithMu ? Total : 0);
	vector<float> XYZ(3 * Total), SelfRev(N);
	uint64_t Off = 0;
	for (uint i = 0; i < N; ++i)
		{
		const ChainData &CD = Chains[i];
		const uint L = Len[i];
		rsk_asserta(RSK_SIZE(*CD.Profile) >= RSK_NFEAT);
		for (uint f = 0; f < RSK_NFEAT; ++f)
			{
			rsk_asserta(RSK_SIZE((*CD.Profile)[f]) == L);
			memcpy(&Prof[(size_t)f * Total + Off], (*CD.Profile)[f].data(), L);
			}
		if (WithMu)
			{
			rsk_asserta(RSK_SIZE(*CD.MuLetters) == L);
			memcpy(&Mu[Off], CD.MuLetters->data(), L);
			}
		rsk_asserta(RSK_SIZE(CD.Chain->m_Xs) == L && RSK_SIZE(CD.Chain->m_Ys) == L && RSK_SIZE(CD.Chain->m_Zs) == L);
		memcpy(&XYZ[Off], CD.Chain->m_Xs.data(), sizeof(float) * L);
		memcpy(&XYZ[Total + Off], CD.Chain->m_Ys.data(), sizeof(float) * L);
		memcpy(&XYZ[2 * Total + Off], CD.Chain->m_Zs.data(), sizeof(float) * L);
		SelfRev[i] = CD.SelfRevScore;
		Off += L;
		}
	rsk_chains_host H;
	H.n = N;
	H.total = Total;
	H.len = Len.data();
	H.prof = Prof.data();
	H.mu = WithMu ? Mu.data() : 0;
	H.xyz = XYZ.data();
	H.selfrev = SelfRev.data();
	rsk_chainset *CS = 0;
	Check(rsk_chainset_upload(Ctx, &H, &CS));
	return CS;
	}

// ---- DSSAligner ----
std::mutex DSSAligner::m_OutputLock;
bool DSSAligner::m_NoSelf = false;
std::atomic<uint> DSSAligner::m_AlnCount{0};
std::atomic<uint> DSSAligner::m_SWCount{0};
std::atomic<uint> DSSAligner::m_MuFilterDiscardCount{0};
std::atomic<uint> DSSAligner::m_MuFilterInputCount{0};
std::atomic<uint> DSSAligner::m_ParasailSaturateCount{0};
std::atomic<uint> DSSAligner::m_XDropAlnCount{0};

DSSAligner::DSSAligner() {}

DSSAligner::~DSSAligner()
	{
	rsk_chainset_free(m_SetA);
	rsk_chainset_free(m_SetB);
	if (m_OwnCtx)
		rsk_ctx_destroy(m_Ctx);
	}

void DSSAligner::SetParams(const DSSParams &Params)
	{
	m_Params = &Params;
	if (m_Ctx != 0 && m_OwnCtx)
		{
		rsk_params R;
		Params.ToRsk(R, DBL_MAX);
		Check(rsk_ctx_set_params(m_Ctx, &R));
		}
	}

void DSSAligner::UseContext(rsk_ctx *Ctx)
	{
	rsk_asserta(m_Ctx == 0);
	m_Ctx = Ctx;
	m_OwnCtx = false;
	}

rsk_ctx *DSSAligner::Ctx()
	{
	if (m_Ctx != 0)
		return m_Ctx;
	rsk_asserta(m_Params != 0);
	rsk_params R;
	m_Params->ToRsk(R, DBL_MAX);
	int Device = 0;
	if (const char *s = getenv("RSK_DEVICE"))
		Device = atoi(s);
	Check(rsk_ctx_create(Device, &R, 0, &m_Ctx));  // no CUDA device -> Die: there is no CPU fallback
	m_OwnCtx = true;
	return m_Ctx;
	}

void DSSAligner::UnsetQuery()
	{
	m_ChainA = 0;
	m_ProfileA = 0;
	m_MuLettersA = 0;
	m_MuKmersA = 0;
	m_SelfRevScoreA = 0;
	rsk_chainset_free(m_SetA);
	m_SetA = 0;
	}

void DSSAligner::SetQuery(const PDBChain &Chain, const vector<vector<byte> > *ptrProfile,
  const vector<byte> *ptrMuLetters, const vector<uint> *ptrMuKmers, float SelfRevScore)
	{
	if (ptrMuKmers != 0)
		rsk_asserta(ptrMuLetters != 0);  // dssaligner.cpp:681-685
	m_ChainA = &Chain;
	m_ProfileA = ptrProfile;
	m_MuLettersA = ptrMuLetters;
	m_MuKmersA = ptrMuKmers;
	m_SelfRevScoreA = SelfRevScore;
	rsk_chainset_free(m_SetA);
	m_SetA = 0;
	}

void DSSAligner::SetTarget(const PDBChain &Chain, const vector<vector<byte> > *ptrProfile,
  const vector<byte> *ptrMuLetters, const vector<uint> *ptrMuKmers, float SelfRevScore)
	{
	m_ChainB = &Chain;
	m_ProfileB = ptrProfile;
	m_MuKmersB = ptrMuKmers;
	m_MuLettersB = ptrMuLetters;
	m_SelfRevScoreB = SelfRevScore;
	rsk_chainset_free(m_SetB);
	m_SetB = 0;
	}

bool DSSAligner::DoMKF() const
	{
	if (m_MuLettersA == 0 || m_MuLettersB == 0)
		return false;
	if (m_MuKmersA == 0 || m_MuKmersB == 0)
		return false;
	if (m_MuLettersA->empty() || m_MuLettersB->empty())
		return false;
	if (m_MuKmersA->empty() || m_MuKmersB->empty())
		return false;
	uint LA = m_ChainA->GetSeqLength();
	uint LB = m_ChainB->GetSeqLength();
	if (LA >= m_Params->m_MKFL)
		return true;
	if (LB >= m_Params->m_MKFL)
		return true;
	return false;
	}

void DSSAligner::ClearAlign()
	{
	m_Path.clear();
	m_LoA = UINT_MAX;
	m_LoB = UINT_MAX;
	m_HiA = UINT_MAX;
	m_HiB = UINT_MAX;
	m_Ids = UINT_MAX;
	m_Gaps = UINT_MAX;
	m_PvalueA = FLT_MAX;
	m_PvalueB = FLT_MAX;
	m_EvalueA = FLT_MAX;
	m_EvalueB = FLT_MAX;
	// (m_QualityA/B are not cleared, as in the reference: dssaligner.cpp:906-927 leaves them, and -alignpair -global prints
	// the quality of the last local alignment, prettyaln.cpp:93)
	m_NewTestStatisticA = -FLT_MAX;
	m_NewTestStatisticB = -FLT_MAX;
	m_AlnFwdScore = 0;
	m_LDDT = 0;
	m_MuFwdScore = m_MuRevScore = m_BestHSPScore = m_BestChainScore = 0;
	m_MuFwdMinusRevScore = 0;
	m_GlobalScore = -9999;
	m_GlobalPath.clear();   // (m_XDropScore is not reset, as in the reference: it keeps the last long-chain score)
	m_Flags = 0;
	}

void DSSAligner::FromHit(const rsk_hit &H, const char *PathPool, const ChainData &A, const ChainData &B)
	{
	m_ChainA = A.Chain;
	m_ChainB = B.Chain;
	m_ProfileA = A.Profile;
	m_ProfileB = B.Profile;
	m_MuLettersA = A.MuLetters;
	m_MuLettersB = B.MuLetters;
	m_SelfRevScoreA = A.SelfRevScore;
	m_SelfRevScoreB = B.SelfRevScore;
	ClearAlign();
	m_Flags = H.flags;
	if (H.flags & RSK_HIT_MKF)
		{
		m_BestHSPScore = H.mu_fwd;
		m_BestChainScore = H.mu_rev;
		}
	else
		{
		m_MuFwdScore = H.mu_fwd;
		m_MuRevScore = H.mu_rev;
		m_MuFwdMinusRevScore = H.mu_score;
		}
	if (H.flags & RSK_HIT_MU_REJECTED)
		return;
	if (H.flags & RSK_HIT_GLOBAL)   // a record of rsk_align_global: what AlignQueryTarget_Global leaves (global.cpp:27-32)
		{
		m_GlobalScore = H.score;
		if (PathPool != 0)
			m_GlobalPath.assign(PathPool + H.path_off, H.path_len);
		m_LoA = 0;
		m_LoB = 0;
		m_Path = m_GlobalPath;
		return;
		}
	m_AlnFwdScore = H.score;
	if (H.flags & RSK_HIT_MKF)
		m_XDropScore = H.score;   // dssaligner.cpp:1427-1429
	if (H.path_len == 0)
		return;
	m_LoA = H.lo_a;
	m_LoB = H.lo_b;
	if (PathPool != 0)
		m_Path.assign(PathPool + H.path_off, H.path_len);
	else
		m_Path.assign(H.path_len, '?');
	if (H.flags & RSK_HIT_HAS_EVALUE)
		{
		m_HiA = H.hi_a;
		m_HiB = H.hi_b;
		m_Ids = H.ids;
		m_Gaps = H.gaps;
		m_LDDT = H.lddt;
		m_NewTestStatisticA = m_NewTestStatisticB = H.ts;
		m_PvalueA = m_PvalueB = H.pvalue;
		m_EvalueA = m_EvalueB = H.evalue;
		m_QualityA = m_QualityB = H.qual;
		}
	}

void DSSAligner::AlignOne(bool NoAccel, bool Bags)
	{
	rsk_asserta(m_Params != 0);
	rsk_asserta(m_ChainA != 0 && m_ChainB != 0 && m_ProfileA != 0 && m_ProfileB != 0);
	rsk_ctx *C = Ctx();
	ChainData A, B;
	A.Chain = m_ChainA; A.Profile = m_ProfileA; A.MuLetters = m_MuLettersA; A.SelfRevScore = m_SelfRevScoreA;
	B.Chain = m_ChainB; B.Profile = m_ProfileB; B.MuLetters = m_MuLettersB; B.SelfRevScore = m_SelfRevScoreB;
	const bool WithMu = (m_MuLettersA != 0 && m_MuLettersB != 0);
	if (m_SetA == 0)
		m_SetA = UploadChains(C, vector<ChainData>(1, A), WithMu);
	if (m_SetB == 0)
		m_SetB = UploadChains(C, vector<ChainData>(1, B), WithMu);

	// per-call parameters: Align_NoAccel skips the Mu filter and the k-mer path (dssaligner.cpp:833-850); the k-mer
	// path also needs both k-mer vectors (DoMKF, dssaligner.cpp:715-732)
	rsk_params R;
	m_Params->ToRsk(R, DBL_MAX);
	if (NoAccel)
		R.omega = 0;
	// (AlignBags decides on the Mu letters and the lengths alone, chainbag.cpp:6-21)
	if (NoAccel || (!Bags && (m_MuKmersA == 0 || m_MuKmersB == 0 || m_MuKmersA->empty() || m_MuKmersB->empty())))
		R.mkfl = UINT_MAX;
	Check(rsk_ctx_set_params(C, &R));

	rsk_search_opts O;
	memset(&O, 0, sizeof(O));
	O.keep = RSK_KEEP_ALL;
	O.want_paths = 1;
	const uint32_t Zero = 0;
	rsk_results *Res = 0;
	Check(rsk_search_pairs(C, m_SetA, m_SetB, 1, &Zero, &Zero, &O, &Res));
	rsk_asserta(rsk_results_count(Res) == 1);
	const rsk_hit &H = rsk_results_hits(Res)[0];
	FromHit(H, rsk_results_paths(Res), A, B);
	rsk_results_free(Res);

	rsk_stats S;
	Check(rsk_ctx_stats(C, &S));
	++m_AlnCount;
	m_SWCount += (uint)S.sw_pairs;
	m_MuFilterInputCount += (uint)S.mu_filter_in;
	m_MuFilterDiscardCount += (uint)S.mu_filter_rejected;
	m_ParasailSaturateCount += (uint)S.mu_saturated;
	m_XDropAlnCount += (uint)S.mkf_pairs;
	}

// global.cpp:7-33: Mu filter (when omega > 0), then the global Viterbi alignment; sets m_GlobalScore, m_GlobalPath,
// m_LoA = m_LoB = 0 and m_Path = m_GlobalPath
void DSSAligner::AlignQueryTarget_Global()
	{
	rsk_asserta(m_Params != 0);
	rsk_asserta(m_ChainA != 0 && m_ChainB != 0 && m_ProfileA != 0 && m_ProfileB != 0);
	rsk_ctx *C = Ctx();
	ChainData A, B;
	A.Chain = m_ChainA; A.Profile = m_ProfileA; A.MuLetters = m_MuLettersA; A.SelfRevScore = m_SelfRevScoreA;
	B.Chain = m_ChainB; B.Profile = m_ProfileB; B.MuLetters = m_MuLettersB; B.SelfRevScore = m_SelfRevScoreB;
	const bool WithMu = (m_MuLettersA != 0 && m_MuLettersB != 0);
	if (m_SetA == 0)
		m_SetA = UploadChains(C, vector<ChainData>(1, A), WithMu);
	if (m_SetB == 0)
		m_SetB = UploadChains(C, vector<ChainData>(1, B), WithMu);
	rsk_params R;
	m_Params->ToRsk(R, DBL_MAX);
	Check(rsk_ctx_set_params(C, &R));
	const uint32_t Zero = 0;
	rsk_results *Res = 0;
	Check(rsk_align_global(C, m_SetA, m_SetB, 1, &Zero, &Zero, &Res));
	rsk_asserta(rsk_results_count(Res) == 1);
	const rsk_hit &H = rsk_results_hits(Res)[0];
	ClearAlign();
	m_Flags = H.flags;
	m_MuFwdScore = H.mu_fwd;
	m_MuRevScore = H.mu_rev;
	m_MuFwdMinusRevScore = H.mu_score;
	++m_AlnCount;
	if (R.omega > 0 && WithMu)
		{
		++m_MuFilterInputCount;
		if (H.flags & RSK_HIT_MU_REJECTED)
			++m_MuFilterDiscardCount;
		}
	if (!(H.flags & RSK_HIT_MU_REJECTED))
		{
		m_GlobalScore = H.score;
		m_GlobalPath.assign(rsk_results_paths(Res) + H.path_off, H.path_len);
		m_LoA = 0;
		m_LoB = 0;
		m_Path = m_GlobalPath;
		}
	rsk_results_free(Res);
	}

void DSSAligner::AlignQueryTarget() { AlignOne(false); }

bool DSSAligner::DoMKF_Bags(const ChainBag &BagA, const ChainBag &BagB) const
	{
	if (BagA.m_ptrMuLetters == 0 || BagB.m_ptrMuLetters == 0)
		return false;
	const uint LA = BagA.m_ptrChain->GetSeqLength(), LB = BagB.m_ptrChain->GetSeqLength();
	return LA >= m_Params->m_MKFL || LB >= m_Params->m_MKFL;
	}

void DSSAligner::AlignBags(const ChainBag &BagA, const ChainBag &BagB)
	{
	rsk_asserta(BagA.m_ptrChain != 0 && BagB.m_ptrChain != 0);
	SetQuery(*BagA.m_ptrChain, BagA.m_ptrProfile, BagA.m_ptrMuLetters, BagA.m_ptrMuLetters ? BagA.m_ptrMuKmers : 0, BagA.m_SelfRevScore);
	SetTarget(*BagB.m_ptrChain, BagB.m_ptrProfile, BagB.m_ptrMuLetters, BagB.m_ptrMuLetters ? BagB.m_ptrMuKmers : 0, BagB.m_SelfRevScore);
	AlignOne(false, true);
	}
void DSSAligner::Align_NoAccel() { AlignOne(true); }

// the aligner's result members as the record + view the C-ABI writers take
void DSSAligner::GetHitView(rsk_hit &H, rsk_hit_view &V) const
	{
	memset(&H, 0, sizeof(H));
	H.score = (m_Flags & RSK_HIT_GLOBAL) ? m_GlobalScore : m_AlnFwdScore;
	H.lo_a = m_LoA; H.lo_b = m_LoB; H.hi_a = m_HiA; H.hi_b = m_HiB;
	H.ids = m_Ids; H.gaps = m_Gaps;
	H.lddt = m_LDDT;
	H.ts = m_NewTestStatisticA;
	H.pvalue = m_PvalueA; H.evalue = m_EvalueA; H.qual = m_QualityA;
	const bool Mkf = (m_Flags & RSK_HIT_MKF) != 0;
	H.mu_fwd = Mkf ? m_BestHSPScore : m_MuFwdScore;
	H.mu_rev = Mkf ? m_BestChainScore : m_MuRevScore;
	H.mu_score = m_MuFwdMinusRevScore;
	H.flags = m_Flags;
	H.path_len = RSK_SIZE(m_Path);
	V.hit = &H;
	V.path = m_Path.c_str();
	V.label_a = m_ChainA->m_Label.c_str();
	V.label_b = m_ChainB->m_Label.c_str();
	V.seq_a = m_ChainA->m_Seq.c_str();
	V.seq_b = m_ChainB->m_Seq.c_str();
	V.len_a = m_ChainA->GetSeqLength();
	V.len_b = m_ChainB->GetSeqLength();
	}

bool DSSAligner::FormatTsvColumns(std::string &Out, bool Up, const char *Columns) const
	{
	if (m_NoSelf && m_ChainA->m_Label == m_ChainB->m_Label)
		return false;
	rsk_hit H;
	rsk_hit_view V;
	GetHitView(H, V);
	static thread_local vector<char> Buf;  // reused: a search emits up to millions of lines
	if (Buf.size() < 4096 + 16 * m_Path.size())
		Buf.resize(4096 + 16 * m_Path.size());
	int n = rsk_format_tsv(&V, Up ? 1 : 0, Columns, Buf.data(), Buf.size());
	if (n < 0)
		Die("reseek_b200: %s", rsk_last_error());
	Out.append(Buf.data(), (size_t)n);
	Out.push_back('\n');
	return true;
	}

void DSSAligner::WriteTsvLine(FILE *f, const char *Line, size_t n)
	{
	if (f == 0 || n == 0)
		return;
	m_OutputLock.lock();  // dssaligner.cpp:1022
	fwrite(Line, 1, n, f);
	m_OutputLock.unlock();
	}

void DSSAligner::ToTsvColumns(FILE *f, bool Up, const char *Columns)
	{
	if (f == 0)
		return;
	static thread_local std::string Line;
	Line.clear();
	if (FormatTsvColumns(Line, Up, Columns))
		WriteTsvLine(f, Line.data(), Line.size());
	}

// dssaligner.cpp:965-979 (PrettyAln) and :981-1014; the text itself comes from the C ABI
void DSSAligner::WriteBlock(FILE *f, bool Up, int Kind, bool Global) const
	{
	if (f == 0)
		return;
	rsk_hit H;
	rsk_hit_view V;
	GetHitView(H, V);
	static thread_local vector<char> Buf;
	if (Buf.size() < 4096)
		Buf.resize(4096);
	for (;;)
		{
		long long n = Kind == 0 ? rsk_format_aln(&V, Up ? 1 : 0, m_RowLen, Buf.data(), Buf.size())
		  : rsk_format_fasta2(&V, Up ? 1 : 0, Global ? 1 : 0, Buf.data(), Buf.size());
		if (n >= 0)
			{
			if (Kind == 1) m_OutputLock.lock();  // ToFasta2 locks, ToAln relies on the caller's lock
			fwrite(Buf.data(), 1, (size_t)n, f);
			if (Kind == 1) m_OutputLock.unlock();
			return;
			}
		if (n >= -8)  // an rsk_status, not a size
			Die("reseek_b200: alignment writer failed (%lld)", n);
		Buf.resize((size_t)(-n) + 64);
		}
	}

// dssaligner.cpp:1371-1385: rotation u and translation t that put the query's aligned residues onto the target's
float DSSAligner::GetKabsch(double t[3], double u[3][3], bool Up) const
	{
	rsk_asserta(m_ChainA != 0 && m_ChainB != 0 && !m_Path.empty());
	auto Planes = [](const PDBChain &C, vector<float> &P)
		{
		const uint L = C.GetSeqLength();
		P.resize(3 * (size_t)L);
		for (uint i = 0; i < L; ++i)
			{
			P[i] = C.m_Xs[i];
			P[(size_t)L + i] = C.m_Ys[i];
			P[2 * (size_t)L + i] = C.m_Zs[i];
			}
		};
	vector<float> PA, PB;
	Planes(*m_ChainA, PA);
	Planes(*m_ChainB, PB);
	double uu[9], msd = 0;
	Check(rsk_kabsch(PA.data(), m_ChainA->GetSeqLength(), PB.data(), m_ChainB->GetSeqLength(), m_LoA, m_LoB, m_Path.c_str(),
	  RSK_SIZE(m_Path), Up ? 1 : 0, t, uu, &msd));
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j)
			u[i][j] = uu[3 * i + j];
	return (float)msd;
	}

void DSSAligner::ToAln(FILE *f, bool Up) const { WriteBlock(f, Up, 0, false); }
void DSSAligner::ToFasta2(FILE *f, bool Global, bool Up) const { WriteBlock(f, Up, 1, Global); }

void DSSAligner::ToTsv(FILE *f, bool Up) { ToTsvColumns(f, Up, 0); }

void DSSAligner::Stats()
	{
	uint Satn = m_ParasailSaturateCount;
	uint Disn = m_MuFilterDiscardCount;
	uint Inn = m_MuFilterInputCount;
	const double Pct = Inn == 0 ? 0.0 : 100.0 * Disn / Inn;
	fprintf(stderr, "DSSAligner::Stats() alns %u, mufil %u/%u %.1f%% (sat %u)\n", (uint)m_AlnCount, Inn, Disn, Pct, Satn);
	}

// alignpair.cpp:7-25: the reversed chain is aligned as the target with the FORWARD Mu letters and k-mers (:22)
float GetSelfRevScore(DSSAligner &DA, const PDBChain &Chain, const vector<vector<byte> > &Profile,
  const vector<vector<byte> > &RevProfile, const vector<byte> *ptrMuLetters, const vector<uint> *ptrMuKmers)
	{
	PDBChain RevChain;
	Chain.GetReverse(RevChain);
	DA.SetQuery(Chain, &Profile, ptrMuLetters, ptrMuKmers, FLT_MAX);
	DA.SetTarget(RevChain, &RevProfile, ptrMuLetters, ptrMuKmers, FLT_MAX);
	DA.AlignQueryTarget();
	return DA.m_AlnFwdScore;
	}

}  // namespace reseek_b200
