// dss.cpp - see dss.h: the reference's DSS class surface over the device kernel; no feature arithmetic on the host.
#include "dss.h"

#include <string.h>

namespace reseek_b200 {

static void Check(int rc)
	{
	if (rc != RSK_OK)
		Die("reseek_b200: %s", rsk_last_error());
	}

DSS::~DSS()
	{
	if (m_OwnCtx && m_Ctx != 0)
		rsk_ctx_destroy(m_Ctx);
	}

void DSS::UseContext(rsk_ctx *Ctx)
	{
	if (m_OwnCtx && m_Ctx != 0)
		rsk_ctx_destroy(m_Ctx);
	m_Ctx = Ctx;
	m_OwnCtx = false;
	}

void DSS::GetFeaturesBatch(rsk_ctx *Ctx, const vector<PDBChain *> &Chains, bool WithMu,
  vector<vector<vector<byte> > *> &Profiles, vector<vector<byte> *> &MuLetters, rsk_chainset **KeepSet)
	{
	const uint N = RSK_SIZE(Chains);
	rsk_asserta(N > 0);
	vector<uint32_t> Len(N);
	uint64_t Total = 0;
	for (uint i = 0; i < N; ++i)
		{
		Len[i] = Chains[i]->GetSeqLength();
		Total += Len[i];
		}
	string AA;
	AA.reserve(Total);
	vector<float> XYZ(3*Total);
	uint64_t Off = 0;
	for (uint i = 0; i < N; ++i)
		{
		const PDBChain &C = *Chains[i];
		const uint L = Len[i];
		rsk_asserta(RSK_SIZE(C.m_Xs) == L && RSK_SIZE(C.m_Ys) == L && RSK_SIZE(C.m_Zs) == L && RSK_SIZE(C.m_Seq) == L);
		AA += C.m_Seq;
		memcpy(&XYZ[Off], C.m_Xs.data(), sizeof(float)*L);
		memcpy(&XYZ[Total + Off], C.m_Ys.data(), sizeof(float)*L);
		memcpy(&XYZ[2*Total + Off], C.m_Zs.data(), sizeof(float)*L);
		Off += L;
		}
	rsk_coords_host H;
	H.n = N;
	H.total = Total;
	H.len = Len.data();
	H.aa = AA.data();
	H.xyz = XYZ.data();
	rsk_chainset *S = 0;
	Check(rsk_chainset_from_coords(Ctx, &H, WithMu ? 1 : 0, &S));
	vector<uint8_t> Prof((size_t) RSK_NFEAT*Total), Mu(WithMu ? Total : 0);
	Check(rsk_chainset_download_features(Ctx, S, Prof.data(), WithMu ? Mu.data() : 0, 0));
	Profiles.assign(N, 0);
	MuLetters.assign(N, 0);
	Off = 0;
	for (uint i = 0; i < N; ++i)
		{
		const uint L = Len[i];
		vector<vector<byte> > *P = new vector<vector<byte> >(RSK_NFEAT);
		for (uint f = 0; f < RSK_NFEAT; ++f)
			(*P)[f].assign(Prof.begin() + (size_t) f*Total + Off, Prof.begin() + (size_t) f*Total + Off + L);
		Profiles[i] = P;
		MuLetters[i] = new vector<byte>;
		if (WithMu)
			MuLetters[i]->assign(Mu.begin() + Off, Mu.begin() + Off + L);
		Off += L;
		}
	if (KeepSet != 0)
		*KeepSet = S;
	else
		rsk_chainset_free(S);
	}

void DSS::Init(const PDBChain &Chain)
	{
	m_Chain = &Chain;
	if (m_Ctx == 0)
		{
		rsk_params R;
		Check(rsk_params_preset(&R, RSK_MODE_SENSITIVE));  // the feature stage does not depend on the search mode
		Check(rsk_ctx_create(0, &R, 0, &m_Ctx));           // dies without a CUDA device
		m_OwnCtx = true;
		}
	vector<PDBChain *> One(1, const_cast<PDBChain *>(&Chain));
	vector<vector<vector<byte> > *> P;
	vector<vector<byte> *> M;
	GetFeaturesBatch(m_Ctx, One, true, P, M, 0);
	m_Profile.swap(*P[0]);
	m_MuLetters.swap(*M[0]);
	delete P[0];
	delete M[0];
	}

void DSS::GetProfile(vector<vector<byte> > &Profile)
	{
	Profile = m_Profile;
	}

void DSS::GetMuLetters(vector<byte> &Letters)
	{
	Letters = m_MuLetters;
	}

// dss.cpp:808-838; the Mu letter is SS3 + 3*NENSS3 + 9*RENDist4 (dss.cpp:629-644), so its parts are read off its digits
uint DSS::GetFeature(FEATURE Feature, uint Pos)
	{
	rsk_asserta(m_Chain != 0 && Pos < GetSeqLength());
	if ((uint) Feature < RSK_NFEAT)
		return m_Profile[(uint) Feature][Pos];
	const uint Mu = m_MuLetters[Pos];
	switch (Feature)
		{
	case FEATURE_Mu: return Mu;
	case FEATURE_SS3: return Mu%3;
	case FEATURE_NENSS3: return (Mu/3)%3;
	case FEATURE_RENDist4: return Mu/9;
	default: break;
		}
	Die("DSS::GetFeature(%d)", (int) Feature);
	return 0;
	}

void DSS::GetMuKmers(const vector<byte> &Letters, vector<uint> &Kmers, const string &PatternStr)
	{
	Kmers.clear();
	const uint PatternLength = RSK_SIZE(PatternStr);
	const uint L = RSK_SIZE(Letters);
	for (uint Pos = 0; Pos + PatternLength <= L; ++Pos)
		{
		uint Kmer = 0;
		for (uint j = 0; j < PatternLength; ++j)
			if (PatternStr[j] == '1')
				Kmer = Kmer*36 + Letters[Pos + j];
		Kmers.push_back(Kmer);
		}
	}

}  // namespace reseek_b200
