// dssaligner.h - DSSAligner look-alike over libreseek_b200 (reference: dssaligner.h:18-237).
//
// Same call surface: SetParams / SetQuery / SetTarget / UnsetQuery / AlignQueryTarget / Align_NoAccel, the public
// result members (m_Path, m_LoA ... m_AlnFwdScore), ToTsv and the static counters.  AlignQueryTarget() is the
// batch-of-one case of rsk_search_pairs; bulk searches go through DBSearcher, which fills a DSSAligner per hit as
// a *view* for OnAln / ToTsv, exactly the role the object plays in the reference's BaseOnAln (dbsearcher.cpp:267-278).
// Ownership: the aligner borrows every pointer handed to SetQuery/SetTarget (they must outlive the call;
// UnsetQuery forgets them, dssaligner.cpp:664-672).  One aligner per thread, never shared (dbsearcher.cpp:98-106).
#pragma once

#include <atomic>
#include <mutex>

#include "reseek_compat.h"

namespace reseek_b200 {

class DBSearcher;

// chainbag.h:5-20: what PostMuFilter keeps per chain.  Only the borrowed chain data matter on the GPU path; the parasail
// profiles and the k-mer hash table of the reference's bags are per-pair scratch that the library builds on the device.
class ChainBag
	{
public:
	const PDBChain *m_ptrChain = 0;
	const vector<vector<byte> > *m_ptrProfile = 0;
	const vector<byte> *m_ptrMuLetters = 0;
	const vector<uint> *m_ptrMuKmers = 0;
	const void *m_ptrProfPara = 0;               // unused here
	const void *m_ptrProfParaRev = 0;            // unused here
	const uint16_t *m_ptrKmerHashTableQ = 0;     // unused here
	float m_SelfRevScore = FLT_MAX;
	};

class DSSAligner
	{
private:
	const DSSParams *m_Params = 0;
	rsk_ctx *m_Ctx = 0;         // created on first use (device RSK_DEVICE or 0); shared when made by a DBSearcher
	bool m_OwnCtx = false;
	rsk_chainset *m_SetA = 0;   // device copy of the current query / target (batch-of-one path)
	rsk_chainset *m_SetB = 0;

public:
	const PDBChain *m_ChainA = 0;
	const PDBChain *m_ChainB = 0;
	const vector<vector<byte> > *m_ProfileA = 0;
	const vector<vector<byte> > *m_ProfileB = 0;
	const vector<byte> *m_MuLettersA = 0;
	const vector<byte> *m_MuLettersB = 0;
	const vector<uint> *m_MuKmersA = 0;   // accepted for signature compatibility; 3-mers are derived on the device
	const vector<uint> *m_MuKmersB = 0;

	string m_Path;
	uint m_LoA = UINT_MAX;
	uint m_LoB = UINT_MAX;
	uint m_HiA = UINT_MAX;
	uint m_HiB = UINT_MAX;
	float m_PvalueA = FLT_MAX;
	float m_PvalueB = FLT_MAX;
	float m_EvalueA = FLT_MAX;
	float m_EvalueB = FLT_MAX;
	float m_QualityA = FLT_MAX;
	float m_QualityB = FLT_MAX;
	float m_NewTestStatisticA = FLT_MAX;
	float m_NewTestStatisticB = FLT_MAX;
	uint m_Ids = UINT_MAX;
	uint m_Gaps = UINT_MAX;
	float m_SelfRevScoreA = FLT_MAX;
	float m_SelfRevScoreB = FLT_MAX;
	float m_AlnFwdScore = FLT_MAX;
	float m_XDropScore = 0;      // score of the banded x-drop alignment of a long-chain pair (dssaligner.cpp:1427-1429)
	float m_LDDT = 0;
	float m_GlobalScore = FLT_MAX;   // -global (global.cpp:27-32)
	string m_GlobalPath;
	// Mu filter / k-mer path by-products (m_MKF.m_BestHSPScore, m_MKF.m_BestChainScore, GetMuScore in the reference)
	int m_MuFwdScore = 0;
	int m_MuRevScore = 0;
	float m_MuFwdMinusRevScore = 0;  // what GetMuScore() returns (dssaligner.cpp:613-617)
	int m_BestHSPScore = 0;
	int m_BestChainScore = 0;
	uint m_Flags = 0;           // RSK_HIT_* of the last alignment

public:
	static std::mutex m_OutputLock;
	static bool m_NoSelf;        // opt(noself): ToTsv skips pairs of equally labelled chains (dssaligner.cpp:1020-1021)
	static std::atomic<uint> m_AlnCount;
	static std::atomic<uint> m_SWCount;
	static std::atomic<uint> m_MuFilterDiscardCount;
	static std::atomic<uint> m_MuFilterInputCount;
	static std::atomic<uint> m_ParasailSaturateCount;
	static std::atomic<uint> m_XDropAlnCount;

public:
	DSSAligner();
	~DSSAligner();
	DSSAligner(const DSSAligner &) = delete;
	DSSAligner &operator=(const DSSAligner &) = delete;

public:
	void SetParams(const DSSParams &Params);
	void UnsetQuery();
	void SetQuery(const PDBChain &Chain, const vector<vector<byte> > *ptrProfile, const vector<byte> *ptrMuLetters,
	  const vector<uint> *ptrMuKmers, float SelfRevScore);
	void SetTarget(const PDBChain &Chain, const vector<vector<byte> > *ptrProfile, const vector<byte> *ptrMuLetters,
	  const vector<uint> *ptrMuKmers, float SelfRevScore);
	bool DoMKF() const;          // dssaligner.cpp:715-732
	void ClearAlign();           // dssaligner.cpp:906-927
	void AlignQueryTarget();     // dssaligner.cpp:793-831
	void AlignQueryTarget_Global();  // global.cpp:7-33
	void AlignBags(const ChainBag &BagA, const ChainBag &BagB);  // chainbag.cpp:44-84 (PostMuFilter's per-candidate call)
	bool DoMKF_Bags(const ChainBag &BagA, const ChainBag &BagB) const;  // chainbag.cpp:6-21
	void Align_NoAccel();        // dssaligner.cpp:833-850: no Mu filter, no k-mer path
	const DSSParams &GetParams() const { return *m_Params; }

// Up is true  if alignment is Query=A, Target=B
// Up is false if alignment is Query=B, Target=A
	void ToTsv(FILE *f, bool Up);
	void ToTsvColumns(FILE *f, bool Up, const char *Columns);  // -columns a+b+c (userfieldnames.h)
	// the same line (with its newline) appended to Out instead of written; false if the reference would print nothing
	bool FormatTsvColumns(std::string &Out, bool Up, const char *Columns) const;
	static void WriteTsvLine(FILE *f, const char *Line, size_t n);  // under the output lock (dssaligner.cpp:1022)
	void ToAln(FILE *f, bool Up) const;                         // -aln (dssaligner.cpp:965-979)
	void ToFasta2(FILE *f, bool Global, bool Up) const;         // -fasta2 [-unaligned] (dssaligner.cpp:981-1014)
	float GetMuScore() const { return m_MuFwdMinusRevScore; }
	float GetKabsch(double t[3], double u[3][3], bool Up) const;  // dssaligner.cpp:1371-1385
	uint m_RowLen = 0;           // -rowlen (0: 80 columns)
	const char *GetLabel(bool Top) const { return Top ? m_ChainA->m_Label.c_str() : m_ChainB->m_Label.c_str(); }
	uint GetLo(bool Top) const { return Top ? m_LoA : m_LoB; }
	uint GetHi(bool Top) const { return Top ? m_HiA : m_HiB; }
	float GetNewTestStatistic(bool Top) const { return Top ? m_NewTestStatisticA : m_NewTestStatisticB; }
	float GetEvalue(bool Top) const { return Top ? m_EvalueA : m_EvalueB; }
	float GetPvalue(bool Top) const { return Top ? m_PvalueA : m_PvalueB; }
	float GetAQ(bool Top) const { return Top ? m_QualityA : m_QualityB; }
	float GetLDDT() const { return m_LDDT; }

	static void Stats();         // dssaligner.cpp:1088-1098

// shim plumbing
	void UseContext(rsk_ctx *Ctx);   // share a DBSearcher's context instead of creating one
	void GetHitView(rsk_hit &H, rsk_hit_view &V) const;
	void WriteBlock(FILE *f, bool Up, int Kind, bool Global) const;
	void FromHit(const rsk_hit &H, const char *PathPool, const ChainData &A, const ChainData &B);

private:
	rsk_ctx *Ctx();
	void AlignOne(bool NoAccel, bool Bags = false);
	};

// alignpair.cpp:7-25.  The reference re-runs DSS on the coordinate-reversed chain; feature extraction is not part of
// this layer, so the caller passes the reversed chain's profile (DSS::GetProfile of PDBChain::GetReverse).
float GetSelfRevScore(DSSAligner &DA, const PDBChain &Chain, const vector<vector<byte> > &Profile,
  const vector<vector<byte> > &RevProfile, const vector<byte> *ptrMuLetters, const vector<uint> *ptrMuKmers);

// helpers shared with DBSearcher
rsk_chainset *UploadChains(rsk_ctx *Ctx, const vector<ChainData> &Chains, bool WithMu);

}  // namespace reseek_b200
