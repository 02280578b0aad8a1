// host_cabi.cpp - plain-C entry points into the host layer (ctypes-friendly): the DSS feature stage for one chain (computed on
// the GPU through the DSS look-alike; RSK_ERR_CUDA-style death without a device, as everywhere in this layer).
// Declared here only (they are helpers of the host library, not part of the device ABI in include/reseek_b200.h).
#include <string.h>

#include "dss.h"

using namespace reseek_b200;

// DSS::GetProfile + GetMuLetters (dss.cpp:716-741, 700-714) of one chain; rev_prof (optional) = profile of the
// coordinate-reversed chain (PDBChain::GetReverse, used by GetSelfRevScore).  prof / rev_prof are [8][L] plane-major.
extern "C" int rskh_dss_features(const char *seq, const float *x, const float *y, const float *z, uint32_t L,
  uint8_t *prof, uint8_t *mu, uint8_t *rev_prof)
	{
	if (seq == 0 || x == 0 || y == 0 || z == 0 || L == 0)
		return RSK_ERR_ARG;
	PDBChain Chain;
	Chain.m_Seq.assign(seq, L);
	Chain.m_Xs.assign(x, x + L);
	Chain.m_Ys.assign(y, y + L);
	Chain.m_Zs.assign(z, z + L);
	DSSParams Params;
	DSS D;
	D.SetParams(Params);
	D.Init(Chain);
	vector<vector<byte> > Profile;
	if (prof != 0)
		{
		D.GetProfile(Profile);
		for (uint f = 0; f < RSK_NFEAT; ++f)
			memcpy(prof + (size_t) f*L, Profile[f].data(), L);
		}
	if (mu != 0)
		{
		vector<byte> Letters;
		D.GetMuLetters(Letters);
		memcpy(mu, Letters.data(), L);
		}
	if (rev_prof != 0)
		{
		PDBChain Rev;
		Chain.GetReverse(Rev);
		D.Init(Rev);
		D.GetProfile(Profile);
		for (uint f = 0; f < RSK_NFEAT; ++f)
			memcpy(rev_prof + (size_t) f*L, Profile[f].data(), L);
		}
	return RSK_OK;
	}
