// dss.cpp - see dss.h.  Every function cites the reference lines it restates; types (float distance, double
// accumulation) and evaluation order are the reference's.
#include "dss.h"

#include <math.h>

#include "dss_tables_data.inc"

namespace reseek_b200 {

static const uint WILDCARD = 0;  // dss.h:9

void DSS::Init(const PDBChain &Chain)
	{
	m_Chain = &Chain;
	m_Density_ScaledValues.clear();
	m_SS.clear();
	m_NENs.clear();
	m_RENs.clear();
	m_SSE_Mids.clear();
	m_SSE_cs.clear();
	m_SSEsSet = false;
	m_ExpBand.clear();
	m_DistBand.clear();
	m_ConfLetters.clear();
	}

// d(i, j) for every residue pair up to m_NEN_W apart, computed once per chain: the neighbour searches (CalcNEN, CalcREN)
// and the density sums all ask for these same float distances, several times each.
void DSS::SetDistBand()
	{
	if (!m_DistBand.empty())
		return;
	const uint L = GetSeqLength();
	const uint Wd = (uint) m_NEN_W;
	m_DistBand.assign((size_t) L*Wd + 1, 0.0f);
	for (uint i = 0; i < L; ++i)
		for (uint k = 0; k < Wd && i + 1 + k < L; ++k)
			m_DistBand[(size_t) i*Wd + k] = GetDist(*m_Chain, i, i + 1 + k);
	}

// exp(-d(i,j)/Radius) for every residue pair up to 50 apart, computed once per chain.  GetDensity and GetSSDensity of the
// reference evaluate this same expression for (Pos, Pos2) and again for (Pos2, Pos) and again for the second feature
// (dss.cpp:236-241, 362-367); d(i,j) is symmetric to the bit (float subtraction, squares) and exp is a function, so the
// stored factor is the value the reference computes each time, and the sums below add them in the reference's order.
void DSS::SetExpBand()
	{
	if (!m_ExpBand.empty())
		return;
	const uint L = GetSeqLength();
	const uint Wd = (uint) (m_Density_W > m_SSDensity_W ? m_Density_W : m_SSDensity_W);
	m_ExpBandW = Wd;
	SetDistBand();
	m_ExpBand.assign((size_t) L*Wd + 1, 0.0);
	for (uint i = 0; i < L; ++i)
		for (uint k = 0; k < Wd && i + 1 + k < L; ++k)
			{
			const double Dist = BandDist(i, i + 1 + k);
			m_ExpBand[(size_t) i*Wd + k] = exp(-Dist/m_Density_Radius);
			}
	}

// float subtraction, float sum of squares (left to right), float sqrt
float DSS::GetDist(const PDBChain &Chain, uint Pos1, uint Pos2)
	{
	const float dx = Chain.m_Xs[Pos1] - Chain.m_Xs[Pos2];
	const float dy = Chain.m_Ys[Pos1] - Chain.m_Ys[Pos2];
	const float dz = Chain.m_Zs[Pos1] - Chain.m_Zs[Pos2];
	const float d2 = dx*dx + dy*dy + dz*dz;
	return sqrtf(d2);
	}

// getss.cpp:6-32 (after sec_str() of TM-align): windows of five C-alphas
static char SSCharFromDists(double d13, double d14, double d15, double d24, double d25, double d35)
	{
	const double DH = 2.1;
	if (fabs(d15 - 6.37) < DH && fabs(d14 - 5.18) < DH && fabs(d25 - 5.18) < DH &&
	  fabs(d13 - 5.45) < DH && fabs(d24 - 5.45) < DH && fabs(d35 - 5.45) < DH)
		return 'h';
	const double DS = 1.42;
	if (fabs(d15 - 13) < DS && fabs(d14 - 10.4) < DS && fabs(d25 - 10.4) < DS &&
	  fabs(d13 - 6.1) < DS && fabs(d24 - 6.1) < DS && fabs(d35 - 6.1) < DS)
		return 's';
	if (d15 < 8.2)
		return 't';
	return '~';
	}

void DSS::GetSS(const PDBChain &Chain, string &SS)
	{
	SS.clear();
	const uint L = Chain.GetSeqLength();
	SS.reserve(L);
	for (uint Pos = 0; Pos < L; ++Pos)
		{
		if (Pos < 2 || Pos + 2 >= L)
			{
			SS += '~';
			continue;
			}
		const double d13 = GetDist(Chain, Pos - 2, Pos);
		const double d14 = GetDist(Chain, Pos - 2, Pos + 1);
		const double d15 = GetDist(Chain, Pos - 2, Pos + 2);
		const double d24 = GetDist(Chain, Pos - 1, Pos + 1);
		const double d25 = GetDist(Chain, Pos - 1, Pos + 2);
		const double d35 = GetDist(Chain, Pos, Pos + 2);
		SS += SSCharFromDists(d13, d14, d15, d24, d25, d35);
		}
	}

void DSS::SetSS()
	{
	if (m_SS.empty())
		GetSS(*m_Chain, m_SS);
	}

// dss.cpp:417-440: closest residue within +-m_NEN_W excluding +-m_NEN_w; first minimum wins (strict <), start 999
uint DSS::CalcNEN(uint Pos) const
	{
	const uint L = GetSeqLength();
	int iLo = int(Pos) - m_NEN_W;
	if (iLo < 0)
		iLo = 0;
	int iHi = int(Pos) + m_NEN_W;
	if (iHi >= int(L))
		iHi = int(L) - 1;
	double MinDist = 999;
	uint MinPos = UINT_MAX;
	for (uint Pos2 = uint(iLo); Pos2 <= uint(iHi); ++Pos2)
		{
		if (Pos2 + m_NEN_w >= Pos && Pos2 <= Pos + m_NEN_w)
			continue;
		const double Dist = BandDist(Pos, Pos2);
		if (Dist < MinDist)
			{
			MinDist = Dist;
			MinPos = Pos2;
			}
		}
	return MinPos;
	}

// dss.cpp:374-415: the same search restricted to the side of the chain the NEN is NOT on
uint DSS::CalcREN(uint Pos, uint NEN) const
	{
	if (NEN == UINT_MAX)
		return UINT_MAX;
	const uint L = GetSeqLength();
	int iLo, iHi;
	if (NEN > Pos)
		{
		iLo = int(Pos) - m_NEN_W;
		if (iLo < 0)
			iLo = 0;
		iHi = int(Pos) - 1;
		}
	else
		{
		iLo = int(Pos) + 1;
		iHi = int(Pos) + m_NEN_W;
		if (iHi >= int(L))
			iHi = int(L) - 1;
		}
	if (iHi < 0)
		return UINT_MAX;
	double MinDist = 999;
	uint MinPos = UINT_MAX;
	for (uint Pos2 = uint(iLo); Pos2 <= uint(iHi); ++Pos2)
		{
		if (Pos2 + m_NEN_w >= Pos && Pos2 <= Pos + m_NEN_w)
			continue;
		const double Dist = BandDist(Pos, Pos2);
		if (Dist < MinDist)
			{
			MinDist = Dist;
			MinPos = Pos2;
			}
		}
	return MinPos;
	}

void DSS::SetNENs()
	{
	if (!m_NENs.empty())
		return;
	SetDistBand();
	const uint L = GetSeqLength();
	m_NENs.reserve(L);
	m_RENs.reserve(L);
	for (uint Pos = 0; Pos < L; ++Pos)
		{
		const uint NEN = CalcNEN(Pos);
		m_NENs.push_back(NEN);
		m_RENs.push_back(CalcREN(Pos, NEN));
		}
	}

double DSS::GetFloat_NENDist(uint Pos)  // dss.cpp:496-503
	{
	SetNENs();
	const uint NEN = m_NENs[Pos];
	if (NEN == UINT_MAX)
		return m_DefaultNENDist;
	return GetDist(*m_Chain, Pos, NEN);
	}

double DSS::GetFloat_RENDist(uint Pos)  // dss.cpp:521-528
	{
	SetNENs();
	const uint REN = m_RENs[Pos];
	if (REN == UINT_MAX)
		return m_DefaultNENDist;
	return GetDist(*m_Chain, Pos, REN);
	}

// myss.cpp:142-183: nine distances around Pos -> nearest of 16 trained cluster centres (first minimum)
static uint ConfLetter(const PDBChain &Chain, uint Pos)
	{
	static const int is[9] = {-2, -2, -2, -1, -1, 0, -3, 0, -3};
	static const int js[9] = {0, 1, 2, 1, 2, 2, 3, 3, 0};
	const uint L = Chain.GetSeqLength();
	if (Pos < 3 || Pos + 3 >= L)
		return WILDCARD;
	double v[9];
	for (uint m = 0; m < 9; ++m)
		v[m] = DSS::GetDist(Chain, uint(int(Pos) + is[m]), uint(int(Pos) + js[m]));
	double MinDist = DBL_MAX;
	uint Best = WILDCARD;
	for (uint k = 0; k < 16; ++k)
		{
		double Sum2 = 0;
		for (uint m = 0; m < 9; ++m)
			{
			const double diff = v[m] - rsk_dss_conf_means[k][m];
			Sum2 += diff*diff;
			}
		const double d = sqrt(Sum2);
		if (k == 0 || d < MinDist)
			{
			Best = k;
			MinDist = d;
			}
		}
	return Best;
	}

// the conformation letter of a position is asked for by Conf (its own) and by NENConf (its neighbour's): once per position
void DSS::SetConfLetters()
	{
	if (!m_ConfLetters.empty())
		return;
	const uint L = GetSeqLength();
	m_ConfLetters.resize(L);
	for (uint Pos = 0; Pos < L; ++Pos)
		m_ConfLetters[Pos] = byte(ConfLetter(*m_Chain, Pos));
	}

uint DSS::Get_Conf(uint Pos)  // myss.cpp:185-193
	{
	SetConfLetters();
	return m_ConfLetters[Pos];
	}

uint DSS::Get_NENConf(uint Pos)  // myss.cpp:195-210
	{
	SetNENs();
	SetConfLetters();
	const uint NEN = m_NENs[Pos];
	if (NEN == UINT_MAX)
		return WILDCARD;
	return m_ConfLetters[NEN];
	}

// dss.cpp:217-244: sum of exp(-d/20) over +-50 residues excluding +-3; undefined (DBL_MAX) at the chain ends
double DSS::GetDensity(uint Pos) const
	{
	const uint L = GetSeqLength();
	if (Pos == 0 || Pos + 1 >= L)
		return DBL_MAX;
	int iLo = int(Pos) - m_Density_W;
	if (iLo < 0)
		iLo = 0;
	int iHi = int(Pos) + m_Density_W;
	if (iHi >= int(L))
		iHi = int(L) - 1;
	double D = 0;
	for (uint Pos2 = uint(iLo); Pos2 <= uint(iHi); ++Pos2)
		{
		if (Pos2 + m_Density_w >= Pos && Pos2 <= Pos + m_Density_w)
			continue;
		D += ExpFactor(Pos, Pos2);
		}
	return D;
	}

// dss.cpp:179-215: min-max scaling over the chain (range floored at 1)
void DSS::SetDensity_ScaledValues()
	{
	if (!m_Density_ScaledValues.empty())
		return;
	SetExpBand();
	const uint L = GetSeqLength();
	vector<double> Values;
	Values.reserve(L);
	double MinValue = 999;
	double MaxValue = 0;
	for (uint Pos = 0; Pos < L; ++Pos)
		{
		const double D = GetDensity(Pos);
		Values.push_back(D);
		if (D != DBL_MAX)
			{
			MinValue = (D < MinValue ? D : MinValue);
			MaxValue = (MaxValue < D ? D : MaxValue);
			}
		}
	double Range = MaxValue - MinValue;
	if (Range < 1)
		Range = 1;
	m_Density_ScaledValues.reserve(L);
	for (uint Pos = 0; Pos < L; ++Pos)
		{
		const double Value = Values[Pos];
		m_Density_ScaledValues.push_back(Value == DBL_MAX ? DBL_MAX : (Value - MinValue)/Range);
		}
	}

// dss.cpp:339-372: share of the exp-weighted neighbourhood (+-50, excluding +-8) that is in state c
double DSS::GetSSDensity(uint Pos, char c)
	{
	SetSS();
	SetExpBand();
	const uint L = GetSeqLength();
	if (Pos == 0 || Pos + 1 >= L)
		return DBL_MAX;
	int iLo = int(Pos) - m_SSDensity_W;
	if (iLo < 0)
		iLo = 0;
	int iHi = int(Pos) + m_SSDensity_W;
	if (iHi >= int(L))
		iHi = int(L) - 1;
	double D = 0;
	double Dc = 0;
	for (uint Pos2 = uint(iLo); Pos2 <= uint(iHi); ++Pos2)
		{
		if (Pos2 + m_SSDensity_w >= Pos && Pos2 <= Pos + m_SSDensity_w)
			continue;
		const double DistFactor = ExpFactor(Pos, Pos2);
		D += DistFactor;
		if (m_SS[Pos2] == c)
			Dc += DistFactor;
		}
	return Dc/(D + m_SSDensity_epsilon);
	}

// dss.cpp:78-110 + 138-155: runs of 'h' / 's' of at least 8 residues, their midpoints
void DSS::SetSSEs()
	{
	if (m_SSEsSet)
		return;
	m_SSEsSet = true;
	SetSS();
	const uint L = GetSeqLength();
	if (L == 0)
		return;
	char currc = m_SS[0];
	uint StartPos = 0;
	uint RunLength = 1;
	for (uint Pos = 1; Pos <= L; ++Pos)
		{
		const char ss = (Pos == L ? 0 : m_SS[Pos]);  // the reference reads the string's terminating NUL here
		if (ss == currc)
			++RunLength;
		else
			{
			if (RunLength >= m_SSE_MinLength && (currc == 'h' || currc == 's'))
				{
				m_SSE_Mids.push_back(StartPos + RunLength/2);
				m_SSE_cs.push_back(currc);
				}
			currc = ss;
			StartPos = Pos;
			RunLength = 1;
			}
		}
	}

// dss.cpp:866-881: distance to the midpoint of the next helix that starts more than 8 residues ahead, 0 if none
double DSS::GetFloat_DstNxtHlx(uint Pos)
	{
	SetSSEs();
	const uint SSECount = RSK_SIZE(m_SSE_Mids);
	for (uint i = 0; i < SSECount; ++i)
		{
		if (m_SSE_cs[i] != 'h')
			continue;
		const uint Mid = m_SSE_Mids[i];
		if (Mid <= Pos + m_SSE_Margin)
			continue;
		return GetDist(*m_Chain, Pos, Mid);
		}
	return 0;
	}

static uint Bin(const double *Ts, double Value)  // valuetoint.cpp: first threshold the value is below
	{
	for (uint i = 0; i < 15; ++i)
		if (Value < Ts[i])
			return i;
	return 15;
	}

static uint SS3Letter(char c)  // dss.cpp:64-76
	{
	switch (c)
		{
	case 'h': return 0;
	case 's': return 1;
	case 't': return 2;
	case '~': return 2;
		}
	return WILDCARD;
	}

// dss.cpp:808-838
uint DSS::GetFeature(FEATURE Feature, uint Pos)
	{
	switch (Feature)
		{
	case FEATURE_AA:
		{
		const uint Letter = rsk_dss_amino_letter[(unsigned char) m_Chain->m_Seq[Pos]];
		return Letter >= 20 ? WILDCARD : Letter;
		}
	case FEATURE_NENDist: return Bin(rsk_dss_bins_NENDist, GetFloat_NENDist(Pos));
	case FEATURE_Conf: return Get_Conf(Pos);
	case FEATURE_NENConf: return Get_NENConf(Pos);
	case FEATURE_RENDist: return Bin(rsk_dss_bins_RENDist, GetFloat_RENDist(Pos));
	case FEATURE_DstNxtHlx: return Bin(rsk_dss_bins_DstNxtHlx, GetFloat_DstNxtHlx(Pos));
	case FEATURE_StrandDens: return Bin(rsk_dss_bins_StrandDens, GetSSDensity(Pos, 's'));
	case FEATURE_NormDens:
		SetDensity_ScaledValues();
		return Bin(rsk_dss_bins_NormDens, m_Density_ScaledValues[Pos]);
	case FEATURE_SS3:
		SetSS();
		return SS3Letter(m_SS[Pos]);
	case FEATURE_NENSS3:  // dss.cpp:30-45
		{
		SetSS();
		SetNENs();
		const uint NEN = m_NENs[Pos];
		return NEN == UINT_MAX ? WILDCARD : SS3Letter(m_SS[NEN]);
		}
	case FEATURE_RENDist4: return GetFeature(FEATURE_RENDist, Pos)/4;  // dss.cpp:548-555
	case FEATURE_Mu:  // dss.cpp:629-644 with m_MuFeatures = SS3, NENSS3, RENDist4 and sizes 3, 3, 4 (dssparams.cpp:7-14)
		return GetFeature(FEATURE_SS3, Pos) + 3*GetFeature(FEATURE_NENSS3, Pos) + 9*GetFeature(FEATURE_RENDist4, Pos);
		}
	Die("DSS::GetFeature(%d)", (int) Feature);
	}

// dss.cpp:716-741 with the default feature list (namedparams.cpp:36-43)
void DSS::GetProfile(vector<vector<byte> > &Profile)
	{
	static const FEATURE Features[RSK_NFEAT] = { FEATURE_AA, FEATURE_NENDist, FEATURE_Conf, FEATURE_NENConf, FEATURE_RENDist,
	  FEATURE_DstNxtHlx, FEATURE_StrandDens, FEATURE_NormDens };
	const uint L = GetSeqLength();
	Profile.clear();
	Profile.resize(RSK_NFEAT);
	for (uint i = 0; i < RSK_NFEAT; ++i)
		{
		vector<byte> &Row = Profile[i];
		Row.reserve(L);
		for (uint Pos = 0; Pos < L; ++Pos)
			Row.push_back(byte(GetFeature(Features[i], Pos)));
		}
	}

void DSS::GetMuLetters(vector<byte> &Letters)  // dss.cpp:700-714
	{
	const uint L = GetSeqLength();
	Letters.clear();
	Letters.reserve(L);
	for (uint Pos = 0; Pos < L; ++Pos)
		Letters.push_back(byte(GetFeature(FEATURE_Mu, Pos)));
	}

void DSS::GetMuKmers(const vector<byte> &Letters, vector<uint> &Kmers, const string &PatternStr)  // dss.cpp:659-682
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
