// profileloader.cpp - see profileloader.h
#include "profileloader.h"


#include "dssaligner.h"

namespace reseek_b200 {

static void Check(int rc)
	{
	if (rc != RSK_OK)
		Die("reseek_b200: %s", rsk_last_error());
	}

void ChainReader2::Open(const string &FN)
	{
	const size_t n = FN.size();
	if (n < 4 || FN.substr(n - 4) != ".bca")
		Die("%s: only .bca input is read by this layer (convert with `reseek -convert ... -bca`)", FN.c_str());
	m_BCA.Open(FN);
	m_NextIdx = 0;
	m_FirstIdx = 0;
	m_EndIdx = UINT_MAX;
	}

void ChainReader2::OpenRange(const string &FN, uint Rank, uint RankCount)
	{
	Open(FN);
	rsk_asserta(RankCount > 0 && Rank < RankCount);
	vector<uint32_t> Bounds(RankCount + 1);
	Check(rsk_partition_by_residues(m_BCA.m_SeqLengths.data(), m_BCA.GetChainCount(), (int) RankCount, Bounds.data()));
	m_FirstIdx = Bounds[Rank];
	m_NextIdx = Bounds[Rank];
	m_EndIdx = Bounds[Rank + 1];
	}

PDBChain *ChainReader2::GetNext()
	{
	for (;;)
		{
		uint Idx;
			{
			std::lock_guard<std::mutex> Guard(m_Lock);
			if (m_NextIdx >= m_BCA.GetChainCount() || m_NextIdx >= m_EndIdx)
				return 0;
			Idx = m_NextIdx++;
			}
		PDBChain *Chain = new PDBChain;
		m_BCA.ReadChain(Idx, *Chain);
		if (Chain->GetSeqLength() == 0)
			{
			delete Chain;
			continue;
			}
		return Chain;
		}
	}

void ChainFeatures::Release()
	{
	Chains.clear(); Profiles.clear(); MuLetters.clear(); MuKmers.clear(); RevProfiles.clear(); SelfRevScores.clear();
	}

void ChainFeatures::Free()
	{
	for (auto *p : Chains) delete p;
	for (auto *p : Profiles) delete p;
	for (auto *p : MuLetters) delete p;
	for (auto *p : MuKmers) delete p;
	Release();
	}

uint ProfileLoader::Load(const DSSParams &Params, ChainReader2 &CR, uint MaxChains, bool WithMu, rsk_ctx *Ctx,
  const DSSParams &SelfRevParams, double MaxEvalue, ChainFeatures &Out, uint ThreadCount)
	{
	Out.Free();
	// single reader keeps file order; the feature extraction below is what is worth threading
	for (;;)
		{
		if (MaxChains != 0 && RSK_SIZE(Out.Chains) >= MaxChains)
			break;
		PDBChain *Chain = CR.GetNext();
		if (Chain == 0)
			break;
		Out.Chains.push_back(Chain);
		}
	const uint N = RSK_SIZE(Out.Chains);
	if (N == 0)
		return 0;
	(void) ThreadCount;  // the feature stage runs on the GPU: one CTA per chain instead of one host thread per chain
	// DSS of the whole block on the device (dss_kernel.cu); the letters come back for the per-chain vectors of the reference's
	// class surface, the device set stays for the self-reverse scores
	rsk_chainset *S = 0;
	DSS::GetFeaturesBatch(Ctx, Out.Chains, WithMu, Out.Profiles, Out.MuLetters, &S);
	Out.MuKmers.assign(N, 0);
	for (uint i = 0; i < N; ++i)
		{
		Out.MuKmers[i] = new vector<uint>;
		if (WithMu)
			DSS::GetMuKmers(*Out.MuLetters[i], *Out.MuKmers[i], "111");  // m_MKFPatternStr (dssparams.cpp:88)
		}
	Out.RevProfiles.clear();

	// self-reverse scores (alignpair.cpp:7-25) for the whole block: chain i against its reversed self (PDBChain::GetReverse +
	// DSS, on the device), which carries the FORWARD Mu letters (:22)
	rsk_chainset *SR = 0;
	Check(rsk_chainset_reversed(Ctx, S, &SR));
	rsk_params Saved, R;
	SelfRevParams.ToRsk(R, MaxEvalue);
	Out.SelfRevScores.assign(N, FLT_MAX);
	Check(rsk_ctx_get_params(Ctx, &Saved));
	Check(rsk_ctx_set_params(Ctx, &R));
	Check(rsk_chainset_selfrev(Ctx, S, SR, Out.SelfRevScores.data()));
	Check(rsk_ctx_set_params(Ctx, &Saved));
	rsk_chainset_free(S);
	rsk_chainset_free(SR);
	return N;
	}

}  // namespace reseek_b200
