// profileloader.h - ChainReader2 / ProfileLoader look-alikes: .bca chains -> DSS features -> the aligner's inputs
// (reference: chainreader2.h, profileloader.cpp:17-110).  Structure formats other than .bca (PDB, mmCIF, .cal,
// directories) are not read by this layer.
#pragma once

#include <mutex>

#include "bcadata.h"
#include "dss.h"

namespace reseek_b200 {

// chainreader2.cpp:204-222 for .bca input: chains in file order, GetNext() hands ownership to the caller
class ChainReader2
	{
public:
	BCAData m_BCA;
	std::mutex m_Lock;
	uint m_NextIdx = 0;

	uint m_EndIdx = UINT_MAX; // one past the last chain this reader delivers

public:
	void Open(const string &FN);
	// Rank `Rank` of `RankCount` processes sharing one .bca: this reader delivers only the rank's contiguous, residue-balanced
	// block of chains (rsk_partition_by_residues over the file's length table, bcadata.cpp:140-168 - no chain data is read to
	// find it; each chain record is then fetched by its own file offset).  GetFirstIdx() is the block's a_base / t_base.
	void OpenRange(const string &FN, uint Rank, uint RankCount);
	uint GetFirstIdx() const { return m_FirstIdx; }
	uint m_FirstIdx = 0;
	PDBChain *GetNext();   // 0 at the end; chains of length 0 are skipped (chainreader2.cpp:103-107)
	uint GetChainCount() const { return m_BCA.GetChainCount(); }
	};

// Everything the aligner needs for a set of chains, in chain order (ProfileLoader pushes in thread arrival order,
// profileloader.cpp:62-69; file order is the deterministic choice and what -threads 1 gives).
struct ChainFeatures
	{
	vector<PDBChain *> Chains;                       // owned
	vector<vector<vector<byte> > *> Profiles;        // owned
	vector<vector<byte> *> MuLetters;                // owned; empty vector when Mu letters were not requested
	vector<vector<uint> *> MuKmers;                  // owned
	vector<vector<vector<byte> > > RevProfiles;      // unused: the reversed chains' DSS stays on the device (rsk_chainset_reversed)
	vector<float> SelfRevScores;
	void Release();                                  // forget the pointers (ownership moved elsewhere)
	void Free();
	};

class ProfileLoader
	{
public:
	// Reads up to MaxChains chains from CR (all when 0), runs DSS on the GPU (one CTA per chain) and computes the self-reverse scores on
	// the GPU with SelfRevParams (ProfileLoader: omega = 0, profileloader.cpp:22-26; RunQuery: the search parameters,
	// runquery.cpp:43).  Returns the number of chains loaded.
	static uint Load(const DSSParams &Params, ChainReader2 &CR, uint MaxChains, bool WithMu, rsk_ctx *Ctx,
	  const DSSParams &SelfRevParams, double MaxEvalue, ChainFeatures &Out, uint ThreadCount = 0);
	};

}  // namespace reseek_b200
