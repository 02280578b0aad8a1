// reseek_compat.h - the handful of Reseek types the search hot path's call surface is written in.
//
// libreseek_b200 replaces the bodies of DSSAligner / DBSearcher; the classes in this directory keep
// the reference's names, method signatures, public result members and error behaviour so that its drivers
// (search.cpp, alignpair.cpp, scop40bench.cpp ...) compile against them unchanged.  Only what that surface needs is
// declared here; structure I/O, DSS feature extraction and the CLI stay in the reference.  Everything lives in
// namespace reseek_b200 so the shim can be linked next to the original classes during a migration.
#pragma once

#include <float.h>
#include <limits.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "../../../include/reseek_b200.h"

namespace reseek_b200 {

typedef unsigned char byte;  // myutils.h
typedef unsigned uint;

using std::string;
using std::vector;

#define RSK_SIZE(v) ((unsigned)(v).size())

// myutils.cpp:785-824: message to stderr, then exit(1).  No exceptions, no return codes.
[[noreturn]] void Die(const char *Format, ...);
void Warning(const char *Format, ...);
#define rsk_asserta(b) ((b) ? (void)0 : ::reseek_b200::Die("%s(%d) assert failed: %s", __FILE__, __LINE__, #b))

// pdbchain.h:10-17: label, amino-acid sequence and C-alpha coordinates are all the hot path reads
class PDBChain
	{
public:
	string m_Label;
	string m_Seq;
	vector<float> m_Xs;
	vector<float> m_Ys;
	vector<float> m_Zs;

public:
	uint GetSeqLength() const { return RSK_SIZE(m_Seq); }
	void GetReverse(PDBChain &Rev) const;  // pdbchain.cpp:478
	};

// dssparams.h:9-26
enum ALGO_MODE { AM_Invalid, AM_Fast, AM_Sensitive, AM_VerySensitive };
enum DECIDE_MODE
	{
	DM_Invalid,
	DM_AlwaysFast,
	DM_AlwaysSensitive,
	DM_AlwaysVerysensitive,
	DM_DefaultFast,
	DM_DefaultSensitive,
	DM_UseCommandLineOption
	};

// dssparams.h:28-120.  The scalars keep the reference's names; the weighted score matrices (m_ScoreMxs there) are
// the flat rsk_params::tables here.
class DSSParams
	{
public:
	vector<float> m_Weights;
	float m_GapOpen = FLT_MAX;
	float m_GapExt = FLT_MAX;
	float m_MinFwdScore = FLT_MAX;
	float m_Omega = FLT_MAX;
	float m_OmegaFwd = FLT_MAX;
	bool m_UsePara = true;
	int m_ParaMuGapOpen = 2;
	int m_ParaMuGapExt = 1;
	uint m_MKFL = UINT_MAX;
	int m_MKF_X1 = INT_MAX;
	int m_MKF_X2 = INT_MAX;
	int m_MKF_MinHSPScore = INT_MAX;
	float m_MKF_MinMegaHSPScore = FLT_MAX;
	ALGO_MODE m_Mode = AM_Invalid;
	float m_Tables[RSK_TABLE_FLOATS];

public:
	// dssparams.cpp:44-111.  DM_UseCommandLineOption / DM_Default* need the command line, which this layer does not
	// have: pass the mode that the command line selected (AM_*) through SetMode instead.
	void SetDSSParams(DECIDE_MODE DM);
	void SetMode(ALGO_MODE AM);
	uint GetFeatureCount() const { return RSK_NFEAT; }
	void ToRsk(rsk_params &R, double MaxEvalue) const;
	};

// One chain as the reference's loaders hand it to the aligner (profileloader.cpp:50-68): borrowed pointers.
struct ChainData
	{
	const PDBChain *Chain = 0;
	const vector<vector<byte> > *Profile = 0;  // [feature][pos]
	const vector<byte> *MuLetters = 0;         // may be 0
	float SelfRevScore = FLT_MAX;
	};

}  // namespace reseek_b200
