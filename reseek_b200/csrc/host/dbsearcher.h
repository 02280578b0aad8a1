// dbsearcher.h - DBSearcher look-alike over libreseek_b200 (reference: dbsearcher.h:14-109).
//
// Same public surface: the per-chain vectors (m_DBChains, m_DBProfiles, m_DBMuLettersVec, m_DBSelfRevScores),
// m_Params, m_MaxEvalue, Setup / RunSelf / RunQuery, the virtual hooks OnSetup / OnAln and BaseOnAln / Reject.
// The pair loops of ThreadBodySelf / ThreadBodyQuery (runself.cpp:13-70, runquery.cpp:18-80) become one
// rsk_search_self / one rsk_search_cross per streamed block; every reported pair is then replayed through BaseOnAln
// with a DSSAligner filled from the hit record, so subclasses (scop40bench.cpp) keep working.
// LoadDB / RunQuery(ChainReader2&) read .bca files and run the DSS look-alike (dss.h) on host threads; other structure
// formats are upstream of this layer: fill the vectors (AddChain) or stream ChainData through a ChainSource.
#pragma once

#include <atomic>
#include <mutex>

#include "dssaligner.h"
#include "profileloader.h"

namespace reseek_b200 {

// What ChainReader2 + DSS deliver to ThreadBodyQuery (runquery.cpp:33-43), as a pull interface:
// GetNext returns false at the end of the stream.  The pointers in ChainData must stay valid until the next call
// with Release = true (the searcher reads a block of chains, aligns it, then releases it).
class ChainSource
	{
public:
	virtual ~ChainSource() {}
	virtual bool GetNext(ChainData &CD) = 0;
	};

class VectorChainSource : public ChainSource
	{
public:
	const vector<ChainData> *m_Chains = 0;
	size_t m_Next = 0;
	explicit VectorChainSource(const vector<ChainData> &Chains) : m_Chains(&Chains) {}
	bool GetNext(ChainData &CD) override;
	};

class DBSearcher
	{
public:
	virtual ~DBSearcher();

public:
	std::mutex m_Lock;
	const DSSParams *m_Params = 0;
	uint m_ThreadCount = UINT_MAX;
	vector<DSSAligner *> m_DAs;

	vector<PDBChain *> m_DBChains;
	bool m_QuerySelf = false;

// Per-chain vectors [ChainIdx]
	vector<vector<vector<byte> > *> m_DBProfiles;
	vector<vector<byte> *> m_DBMuLettersVec;
	vector<vector<uint> *> m_DBMuKmersVec;
	vector<float> m_DBSelfRevScores;

	std::atomic<uint> m_ProcessedQueryCount{0};
	std::atomic<uint> m_ProcessedPairCount{0};
	std::atomic<uint> m_HitCount{0};
	double m_MaxEvalue = 10;
	uint m_Secs = UINT_MAX;

// where BaseOnAln writes (g_fTsv in the reference, set from -output)
	FILE *m_fTsv = 0;
	FILE *m_fAln = 0;            // g_fAln (-aln), g_fFasta2 (-fasta2), opt(unaligned), opt(rowlen)
	FILE *m_fFasta2 = 0;
	bool m_Unaligned = false;
	bool m_Global = false;       // opt(global): RunSelf aligns globally (runself.cpp:48-57)
	uint m_RowLen = 0;
	const char *m_Columns = 0;   // -columns, 0 = default
	bool m_OwnsChains = false;   // the reference's destructor deletes chains/profiles (dbsearcher.cpp:12-22)
	uint m_BlockChains = 100000; // streamed chains per device block in RunQuery
	int m_Device = 0;
	// GPUs RunQuery shards every streamed block over (devices m_Device .. m_Device + m_GpuCount - 1).  0 = all visible
	// devices (RSK_GPUS in the environment overrides).  The reference's counterpart is m_ThreadCount worker threads pulling
	// chains from one reader (runquery.cpp:82-125); here one host thread per GPU takes a contiguous, residue-balanced share
	// of the block and the hits are gathered on the first GPU over NVLink (rsk_search_cross_sharded).
	uint m_GpuCount = 0;

public:
	void Setup();
	void LoadDB(const string &DBFN);   // dbsearcher.cpp:242-256 (.bca input): chains, DSS profiles, Mu letters, self-reverse scores
	uint GetDBChainCount() const { return RSK_SIZE(m_DBChains); }
	uint GetDBSize() const { return GetDBChainCount(); }
	void AddChain(PDBChain *ptrChain, vector<vector<byte> > *ptrProfile, vector<byte> *ptrMuLetters);

	void RunQuery(ChainSource &QCR);
	void RunQuery(ChainReader2 &QCR);  // runquery.cpp:82-130: streamed chains read from a .bca, DSS + self-reverse per block
	void RunSelf();
	void RunSelfGlobal();
	void RunStats() const;
	bool Reject(DSSAligner &DA, bool Up) const;

public:
	virtual void OnSetup() {}
	void BaseOnAln(DSSAligner &DA, bool Up);
	// BaseOnAln with the TSV line already formatted (Line/n; n = 0: nothing to print).  EmitHits replays the hits of a result block
	// through it: the lines of a block are formatted on the host threads first, then every hit goes through Reject / the
	// writers / OnAln on the calling thread, in result order, with `DA` holding the hit as in BaseOnAln.
	void BaseOnAlnLine(DSSAligner &DA, bool Up, const char *Line, size_t n);
	void EmitHits(const rsk_hit *Hits, uint64_t N, const char *Pool, const vector<ChainData> *BlockA, bool BothDirections);
	virtual void OnAln(DSSAligner &DA, bool Up) {}

// shim plumbing
	rsk_ctx *GetContext();
	const rsk_stats &GetLastStats() const { return m_LastStats; }

private:
	rsk_ctx *m_Ctx = 0;
	rsk_ctx *m_LoaderCtx = 0;    // second context: the next block's self-reverse scores while the current block is searched
	rsk_chainset *m_DBSet = 0;
	rsk_stats m_LastStats;
	// ranks 1 .. m_GpuCount-1 (rank 0 is m_Ctx / m_DBSet): context, communicator, replica of the in-memory chains
	vector<rsk_ctx *> m_RankCtx;
	vector<rsk_comm *> m_RankComm;
	vector<rsk_chainset *> m_RankDB;
	void SetupRanks();
	void RunQueryBlockSharded(const vector<ChainData> &Block);
	void UploadDB();
	void BeginRun();
	void RunQueryBlock(const vector<ChainData> &Block);
	ChainData GetDBChainData(uint Idx) const;
	void AddStats();
	};

// search.cpp:76-111 (`-search Q -db DB -fast`): MuPreFilter (muprefilter.cpp:64-133) writes the candidate TSV,
// PostMuFilter (postmufilter.cpp:211-301) aligns every listed (query, target) pair with the sensitive preset and
// writes the hits with Up = true.  The reference passes file names for the chains; here the chains come in memory.
void MuPreFilter(const DSSParams &Params, const vector<ChainData> &Query, const vector<ChainData> &DB,
  const string &OutputFN, int Device = 0);
void PostMuFilter(const DSSParams &Params, const string &MuFilterTsvFN, const vector<ChainData> &Query,
  const vector<ChainData> &DB, const string &HitsFN, const char *Columns = 0, int Device = 0,
  const string &AlnFN = string(), double MaxEvalue = 10);  // AlnFN: -aln (postmufilter.cpp:194, 244); MaxEvalue: -evalue (:217-220)

// Both stages on a DB block-partitioned over GpuCount GPUs (0 = all visible): one host thread per GPU, prefilter triples
// all-gathered and the merged bag replayed on every GPU, hits gathered on the first (rsk_search_fast_db_sharded).  Hits are
// written in the candidate TSV's line order (targets ascending, queries ascending), as the reference at -threads 1.
// CandTsvFN (optional) receives the merged candidate list.
void SearchFastDB(const DSSParams &Params, const vector<ChainData> &Query, const vector<ChainData> &DB, const string &HitsFN,
  const char *Columns = 0, const string &AlnFN = string(), double MaxEvalue = 10, uint GpuCount = 0,
  const string &CandTsvFN = string());
uint ResolveGpuCount(uint Requested, int FirstDevice);

// The same two stages with the chains named by file, PostMuFilter with the reference's own argument list
// (search.cpp:14-18, postmufilter.cpp:211-216): the .bca files are read, run through DSS and given the self-reverse scores
// of the sensitive preset (postmufilter.cpp:79-81, 166-167) before the in-memory versions above take over.  The reference's
// MuPreFilter takes a SeqDB and a MuSeqSource of Mu-letter sequences made from the same two files (search.cpp:82-101).
void MuPreFilter(const DSSParams &Params, const string &QueryCAFN, const string &DBBCAFN, const string &OutputFN);
void PostMuFilter(const DSSParams &Params, const string &MuFilterTsvFN, const string &QueryCAFN, const string &DBBCAFN,
  const string &HitsFN);

}  // namespace reseek_b200
