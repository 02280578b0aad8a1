// profileloader.cpp - see profileloader.h
#include "profileloader.h"

#include <atomic>
#include <thread>

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
	}

PDBChain *ChainReader2::GetNext()
	{
	for (;;)
		{
		uint Idx;
			{
			std::lock_guard<std::mutex> Guard(m_Lock);
			if (m_NextIdx >= m_BCA.GetChainCount())
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
	Out.Profiles.resize(N, 0);
	Out.MuLetters.resize(N, 0);
	Out.MuKmers.resize(N, 0);
	Out.RevProfiles.resize(N);
	if (ThreadCount == 0)
		ThreadCount = std::max(1u, std::thread::hardware_concurrency());
	ThreadCount = std::min(ThreadCount, N);
	std::atomic<uint> Next{0};
	auto Body = [&]()
		{
		DSS D;
		D.SetParams(Params);
		for (;;)
			{
			const uint i = Next.fetch_add(1);
			if (i >= N)
				return;
			const PDBChain &Chain = *Out.Chains[i];
			D.Init(Chain);
			Out.Profiles[i] = new vector<vector<byte> >;
			D.GetProfile(*Out.Profiles[i]);
			Out.MuLetters[i] = new vector<byte>;
			Out.MuKmers[i] = new vector<uint>;
			if (WithMu)
				{
				D.GetMuLetters(*Out.MuLetters[i]);
				D.GetMuKmers(*Out.MuLetters[i], *Out.MuKmers[i], "111");  // m_MKFPatternStr (dssparams.cpp:88)
				}
			PDBChain Rev;
			Chain.GetReverse(Rev);
			D.Init(Rev);
			D.GetProfile(Out.RevProfiles[i]);
			}
		};
	vector<std::thread> ts;
	for (uint t = 1; t < ThreadCount; ++t)
		ts.emplace_back(Body);
	Body();
	for (auto &t : ts)
		t.join();

	// self-reverse scores (alignpair.cpp:7-25) for the whole block in one GPU call: chain i against its reversed self,
	// which carries the FORWARD Mu letters (:22)
	vector<ChainData> Fwd(N), Rev(N);
	vector<PDBChain> RevChains(N);
	for (uint i = 0; i < N; ++i)
		{
		Fwd[i].Chain = Out.Chains[i];
		Fwd[i].Profile = Out.Profiles[i];
		Fwd[i].MuLetters = WithMu ? Out.MuLetters[i] : 0;
		Out.Chains[i]->GetReverse(RevChains[i]);
		Rev[i].Chain = &RevChains[i];
		Rev[i].Profile = &Out.RevProfiles[i];
		Rev[i].MuLetters = WithMu ? Out.MuLetters[i] : 0;
		}
	rsk_params Saved, R;
	SelfRevParams.ToRsk(R, MaxEvalue);
	rsk_chainset *S = UploadChains(Ctx, Fwd, WithMu);
	rsk_chainset *SR = UploadChains(Ctx, Rev, WithMu);
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
