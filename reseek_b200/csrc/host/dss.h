// dss.h - DSS look-alike: the reference's structure -> feature-letter stage (dss.h:14-119), SURVEY §8(f) row 1.
//
// Input: one PDBChain (amino-acid sequence + C-alpha coordinates).  Output: the 8-plane profile the aligner consumes
// (AA, NENDist, Conf, NENConf, RENDist, DstNxtHlx, StrandDens, NormDens; namedparams.cpp:36-43), the Mu letters
// (SS3 + 3*NENSS3 + 9*RENDist4, dssparams.cpp:7-14) and the Mu 3-mers.  Host code on purpose: the letters depend on double
// `exp`, float distances and first-minimum argmins, and a single flipped letter changes alignments, so every
// operation is performed in the reference's type and order with the same libm.  Threads give the parallelism
// (one DSS object per thread, like ProfileLoader::ThreadBody, profileloader.cpp:17-70).
#pragma once

#include "reseek_compat.h"

namespace reseek_b200 {

enum FEATURE
	{
	FEATURE_AA,
	FEATURE_NENDist,
	FEATURE_Conf,
	FEATURE_NENConf,
	FEATURE_RENDist,
	FEATURE_DstNxtHlx,
	FEATURE_StrandDens,
	FEATURE_NormDens,
	FEATURE_SS3,
	FEATURE_NENSS3,
	FEATURE_RENDist4,
	FEATURE_Mu
	};

class DSS
	{
public:
	const PDBChain *m_Chain = 0;
	string m_SS;                  // 'h' 's' 't' '~' per residue (getss.cpp:34-63)
	vector<uint> m_NENs;          // nearest "non-local" neighbour, UINT_MAX = none (dss.cpp:417-440)
	vector<uint> m_RENs;          // nearest neighbour on the other side of the chain (dss.cpp:374-415)
	vector<double> m_Density_ScaledValues;
	vector<uint> m_SSE_Mids;
	vector<char> m_SSE_cs;
	vector<double> m_ExpBand;     // exp(-d(i, i+1+k)/Radius), k < m_ExpBandW (see SetExpBand)
	uint m_ExpBandW = 0;
	vector<float> m_DistBand;     // d(i, i+1+k), k < m_NEN_W (see SetDistBand)
	vector<byte> m_ConfLetters;   // conformation letter of every position

	// dss.h:23-37
	int m_Density_W = 50;
	int m_Density_w = 3;
	int m_SSDensity_W = 50;
	int m_SSDensity_w = 8;
	double m_Density_Radius = 20.0;
	int m_NEN_W = 100;
	int m_NEN_w = 12;
	double m_DefaultNENDist = 10.0;
	double m_SSDensity_epsilon = 1;
	uint m_SSE_MinLength = 8;
	uint m_SSE_Margin = 8;

private:
	const DSSParams *m_Params = 0;
	bool m_SSEsSet = false;

public:
	void SetParams(const DSSParams &Params) { m_Params = &Params; }
	void Init(const PDBChain &Chain);
	uint GetSeqLength() const { return m_Chain->GetSeqLength(); }

	uint GetFeature(FEATURE Feature, uint Pos);
	void GetProfile(vector<vector<byte> > &Profile);
	void GetMuLetters(vector<byte> &Letters);
	void GetMuKmers(const vector<byte> &MuLetters, vector<uint> &Kmers, const string &PatternStr);

	void SetSS();
	void SetNENs();
	void SetSSEs();
	void SetDensity_ScaledValues();
	void SetExpBand();
	void SetDistBand();
	void SetConfLetters();
	float BandDist(uint Pos, uint Pos2) const   // |Pos - Pos2| in 1..m_NEN_W
		{
		const uint lo = Pos < Pos2 ? Pos : Pos2, d = Pos < Pos2 ? Pos2 - Pos : Pos - Pos2;
		return m_DistBand[(size_t) lo*(uint) m_NEN_W + (d - 1)];
		}
	double ExpFactor(uint Pos, uint Pos2) const
		{
		const uint lo = Pos < Pos2 ? Pos : Pos2, d = Pos < Pos2 ? Pos2 - Pos : Pos - Pos2;
		return m_ExpBand[(size_t) lo*m_ExpBandW + (d - 1)];
		}
	double GetDensity(uint Pos) const;
	double GetSSDensity(uint Pos, char c);
	uint CalcNEN(uint Pos) const;
	uint CalcREN(uint Pos, uint NEN) const;
	uint Get_Conf(uint Pos);
	uint Get_NENConf(uint Pos);
	double GetFloat_NENDist(uint Pos);
	double GetFloat_RENDist(uint Pos);
	double GetFloat_DstNxtHlx(uint Pos);

	static float GetDist(const PDBChain &Chain, uint Pos1, uint Pos2);  // pdbchain.cpp:310-318, abcxyz.h:116-126 (float)
	static void GetSS(const PDBChain &Chain, string &SS);               // getss.cpp:34-63
	};

}  // namespace reseek_b200
