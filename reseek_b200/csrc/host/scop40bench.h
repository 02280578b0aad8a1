// scop40bench.h - SCOP40Bench look-alike (reference: scop40bench.h:9-178): the author's accuracy harness as an OnAln subclass of
// DBSearcher, so that it runs unchanged on the GPU engine.
//
// RunSelf() aligns all chain pairs; every reported alignment reaches OnAln (scop40bench.cpp:291-322), which stores
// (domain1, domain2, E-value); the summary sorts the hits by E-value, walks the ROC steps (one step per distinct score,
// scop40benchroc.cpp:454-511) and reports the sensitivity - true-positive pairs found / true pairs that exist - at 0.1, 1 and 10
// false positives per query (GetNTPAtEPQThreshold, scop40benchroc.cpp:26-41; WriteSummary, scop40bench.cpp:588-612).
// Truth is SCOP superfamily membership (level "sf": same superfamily = true, anything else = false; scop40benchroc.cpp:166-
// 193), read from a two-column lookup (domain <tab> class.fold.superfamily.family, test_data/dom_scopid.tsv).
#pragma once

#include <map>

#include "dbsearcher.h"

namespace reseek_b200 {

class SCOP40Bench : public DBSearcher
	{
public:
	vector<string> m_Doms;              // domain (chain) labels, index = chain index
	vector<uint> m_DomIdxToSFIdx;       // UINT_MAX when the lookup does not know the domain
	std::map<string, uint> m_LabelToChainIdx;
	vector<uint> m_DomIdx1s, m_DomIdx2s;
	vector<float> m_Scores;             // E-values
	uint m_NT = 0;                      // ordered pairs of different domains in the same superfamily
	float m_SensEPQ0_1 = 0, m_SensEPQ1 = 0, m_SensEPQ10 = 0;

public:
	void ReadLookup(const string &FN);  // before Setup()
	void OnSetup() override;
	void OnAln(DSSAligner &DA, bool Up) override;
	void StoreScore(uint ChainIdx1, uint ChainIdx2, float Score);
	void SetStats();
	void WriteSummary(FILE *f) const;

private:
	std::map<string, string> m_DomToScopId;
	};

}  // namespace reseek_b200
