// scop40bench.cpp - see scop40bench.h
#include "scop40bench.h"

#include <algorithm>
#include <numeric>

namespace reseek_b200 {

void SCOP40Bench::ReadLookup(const string &FN)
	{
	FILE *f = fopen(FN.c_str(), "r");
	if (f == 0)
		Die("Cannot open %s", FN.c_str());
	char Line[1024];
	while (fgets(Line, sizeof(Line), f) != 0)
		{
		string s(Line);
		while (!s.empty() && (s.back() == '\n' || s.back() == '\r'))
			s.pop_back();
		const size_t Tab = s.find('\t');
		if (Tab == string::npos)
			continue;
		m_DomToScopId[s.substr(0, Tab)] = s.substr(Tab + 1);
		}
	fclose(f);
	}

// scop40bench.cpp:217-289 (BuildDomSFIndexesFromDBChainLabels): label -> chain index, domain -> superfamily index
void SCOP40Bench::OnSetup()
	{
	const uint N = GetDBChainCount();
	m_Doms.clear();
	m_DomIdxToSFIdx.assign(N, UINT_MAX);
	m_LabelToChainIdx.clear();
	std::map<string, uint> SFToIdx;
	vector<uint> SFSizes;
	for (uint i = 0; i < N; ++i)
		{
		const string &Label = m_DBChains[i]->m_Label;
		m_Doms.push_back(Label);
		if (m_LabelToChainIdx.find(Label) == m_LabelToChainIdx.end())
			m_LabelToChainIdx[Label] = i;
		const string Dom = Label.substr(0, Label.find('/'));
		std::map<string, string>::const_iterator it = m_DomToScopId.find(Dom);
		if (it == m_DomToScopId.end())
			continue;
		// class.fold.superfamily.family -> class.fold.superfamily
		const string &Id = it->second;
		size_t p = 0;
		for (int k = 0; k < 3 && p != string::npos; ++k)
			p = Id.find('.', p + 1);
		const string SF = p == string::npos ? Id : Id.substr(0, p);
		std::map<string, uint>::const_iterator is = SFToIdx.find(SF);
		uint SFIdx;
		if (is == SFToIdx.end())
			{
			SFIdx = RSK_SIZE(SFSizes);
			SFToIdx[SF] = SFIdx;
			SFSizes.push_back(0);
			}
		else
			SFIdx = is->second;
		m_DomIdxToSFIdx[i] = SFIdx;
		++SFSizes[SFIdx];
		}
	uint64_t NT = 0;
	for (uint n : SFSizes)
		NT += (uint64_t) n*(n - 1);
	m_NT = (uint) NT;
	m_DomIdx1s.clear();
	m_DomIdx2s.clear();
	m_Scores.clear();
	}

void SCOP40Bench::StoreScore(uint ChainIdx1, uint ChainIdx2, float Score)
	{
	m_DomIdx1s.push_back(ChainIdx1);
	m_DomIdx2s.push_back(ChainIdx2);
	m_Scores.push_back(Score);
	}

// scop40bench.cpp:291-322 (called under DBSearcher::m_Lock by BaseOnAln)
void SCOP40Bench::OnAln(DSSAligner &DA, bool Up)
	{
	std::map<string, uint>::const_iterator iterA = m_LabelToChainIdx.find(DA.m_ChainA->m_Label);
	std::map<string, uint>::const_iterator iterB = m_LabelToChainIdx.find(DA.m_ChainB->m_Label);
	rsk_asserta(iterA != m_LabelToChainIdx.end() && iterB != m_LabelToChainIdx.end());
	const uint A = iterA->second, B = iterB->second;
	if (A == B)
		return;
	if (Up)
		StoreScore(A, B, DA.m_EvalueA);
	else
		StoreScore(B, A, DA.m_EvalueB);
	}

// SetStats (scop40bench.cpp:566-586) for the three numbers of the summary line
void SCOP40Bench::SetStats()
	{
	const size_t HitCount = m_Scores.size();
	vector<uint> Order(HitCount);
	std::iota(Order.begin(), Order.end(), 0u);
	std::stable_sort(Order.begin(), Order.end(), [&](uint x, uint y) { return m_Scores[x] < m_Scores[y]; });
	// ROC steps: counts are recorded whenever the score changes (scop40benchroc.cpp:454-511)
	vector<uint> NTPs, NFPs;
	uint NTP = 0, NFP = 0;
	if (HitCount > 0)
		{
		float Current = m_Scores[Order[0]];
		for (size_t k = 0; k < HitCount; ++k)
			{
			const uint i = Order[k];
			const uint D1 = m_DomIdx1s[i], D2 = m_DomIdx2s[i];
			if (D1 == D2)
				continue;
			if (m_Scores[i] != Current)
				{
				NTPs.push_back(NTP);
				NFPs.push_back(NFP);
				Current = m_Scores[i];
				}
			const uint S1 = m_DomIdxToSFIdx[D1], S2 = m_DomIdxToSFIdx[D2];
			if (S1 == UINT_MAX && S2 == UINT_MAX)
				continue;            // IsT == -1
			if (S1 != UINT_MAX && S1 == S2)
				++NTP;
			else
				++NFP;
			}
		NTPs.push_back(NTP);
		NFPs.push_back(NFP);
		}
	const uint QueryCount = RSK_SIZE(m_Doms);
	auto NTPAt = [&](float EPQThreshold)   // scop40benchroc.cpp:26-41
		{
		uint ntp = 0;
		for (size_t Idx = 0; Idx < NTPs.size(); ++Idx)
			{
			const float EPQ = float(NFPs[Idx])/QueryCount;
			if (Idx > 0)
				ntp = NTPs[Idx];
			if (EPQ >= EPQThreshold)
				break;
			}
		return ntp;
		};
	const float NT = m_NT == 0 ? 1.0f : float(m_NT);
	m_SensEPQ0_1 = float(NTPAt(0.1f))/NT;
	m_SensEPQ1 = float(NTPAt(1))/NT;
	m_SensEPQ10 = float(NTPAt(10))/NT;
	}

void SCOP40Bench::WriteSummary(FILE *f) const
	{
	fprintf(f, "SEPQ0.1=%.4f SEPQ1=%.4f SEPQ10=%.4f hits=%u NT=%u secs=%u\n", m_SensEPQ0_1, m_SensEPQ1, m_SensEPQ10,
	  (uint) m_Scores.size(), m_NT, m_Secs == UINT_MAX ? 0 : m_Secs);
	}

}  // namespace reseek_b200
