// dss.h - DSS look-alike (reference: dss.h:14-119) over the device DSS of libreseek_b200.
//
// Input: PDBChains (amino-acid sequence + C-alpha coordinates).  Output: the 8-plane profile the aligner consumes
// (AA, NENDist, Conf, NENConf, RENDist, DstNxtHlx, StrandDens, NormDens; namedparams.cpp:36-43), the Mu letters
// (SS3 + 3*NENSS3 + 9*RENDist4, dssparams.cpp:7-14) and the Mu 3-mers.  No feature arithmetic lives here: Init() ships the
// chain to the GPU (rsk_chainset_from_coords, dss_kernel.cu - one CTA per chain, letter-exact against the reference) and the
// getters read the planes back.  ProfileLoader uses the batch form (GetFeaturesBatch) for whole blocks of chains.
// Like the rest of this layer it has no CPU fallback: without a CUDA device Init() dies.
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

private:
	const DSSParams *m_Params = 0;
	rsk_ctx *m_Ctx = 0;
	bool m_OwnCtx = false;
	vector<vector<byte> > m_Profile;   // [feature][pos] of the current chain
	vector<byte> m_MuLetters;

public:
	~DSS();
	void SetParams(const DSSParams &Params) { m_Params = &Params; }
	void UseContext(rsk_ctx *Ctx);     // the GPU context to compute on; without it one is created on device 0 at the first Init
	void Init(const PDBChain &Chain);  // dss.cpp:46-62: binds the chain; here also computes its letters on the device
	uint GetSeqLength() const { return m_Chain->GetSeqLength(); }

	uint GetFeature(FEATURE Feature, uint Pos);              // dss.cpp:808-838
	void GetProfile(vector<vector<byte> > &Profile);         // dss.cpp:716-741
	void GetMuLetters(vector<byte> &Letters);                // dss.cpp:700-714
	static void GetMuKmers(const vector<byte> &MuLetters, vector<uint> &Kmers, const string &PatternStr);  // dss.cpp:659-682

	// Whole block at once: letters of every chain (Profiles[i], MuLetters[i] when WithMu) and, when KeepSet is given, the
	// device chain set they live in (self-reverse scores unset), so that the caller can go on without a second upload.
	static void GetFeaturesBatch(rsk_ctx *Ctx, const vector<PDBChain *> &Chains, bool WithMu,
	  vector<vector<vector<byte> > *> &Profiles, vector<vector<byte> *> &MuLetters, rsk_chainset **KeepSet);
	};

}  // namespace reseek_b200
