// dbsearcher.cpp - DBSearcher look-alike and the `-fast -db` drivers over the C ABI (see dbsearcher.h).
#include "dbsearcher.h"

#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <chrono>
#include <future>
#include <memory>
#include <numeric>
#include <thread>

namespace reseek_b200 {

static void Check(int rc)
	{
	if (rc != RSK_OK)
		Die("reseek_b200: %s", rsk_last_error());
	}

// RSK_TIMING=1: wall-clock phases of the host layer on stderr (developer aid)
static double NowMs()
	{
	return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
	}
static void Phase(const char *What, double &t0)
	{
	const double t1 = NowMs();
	if (getenv("RSK_TIMING") != 0)
		fprintf(stderr, "[reseek_b200 host] %-34s %9.1f ms\n", What, t1 - t0);
	t0 = t1;
	}

bool VectorChainSource::GetNext(ChainData &CD)
	{
	if (m_Next >= m_Chains->size())
		return false;
	CD = (*m_Chains)[m_Next++];
	return true;
	}

// dbsearcher.cpp:12-22
DBSearcher::~DBSearcher()
	{
	for (DSSAligner *DA : m_DAs)
		delete DA;
	for (rsk_chainset *S : m_RankDB)
		rsk_chainset_free(S);
	for (rsk_comm *C : m_RankComm)
		rsk_comm_destroy(C);
	for (size_t r = 1; r < m_RankCtx.size(); ++r)
		rsk_ctx_destroy(m_RankCtx[r]);
	rsk_chainset_free(m_DBSet);
	if (m_LoaderCtx != 0)
		rsk_ctx_destroy(m_LoaderCtx);
	if (m_Ctx != 0)
		rsk_ctx_destroy(m_Ctx);
	if (m_OwnsChains)
		{
		for (PDBChain *C : m_DBChains) delete C;
		for (auto *P : m_DBProfiles) delete P;
		for (auto *M : m_DBMuLettersVec) delete M;
		for (auto *K : m_DBMuKmersVec) delete K;
		}
	}

rsk_ctx *DBSearcher::GetContext()
	{
	if (m_Ctx != 0)
		return m_Ctx;
	rsk_asserta(m_Params != 0);
	rsk_params R;
	m_Params->ToRsk(R, m_MaxEvalue);
	Check(rsk_ctx_create(m_Device, &R, 0, &m_Ctx));  // Die()s when there is no CUDA device: no CPU fallback
	return m_Ctx;
	}

// dbsearcher.cpp:73-121.  -evalue is not visible here: set m_MaxEvalue before Setup() to override the default.
void DBSearcher::Setup()
	{
	rsk_asserta(m_Params != 0);
	rsk_asserta(m_DAs.empty());
	if (m_MaxEvalue == 10 && m_Params->m_Mode == AM_VerySensitive)
		m_MaxEvalue = DBL_MAX;
	m_ProcessedQueryCount = 0;
	m_ProcessedPairCount = 0;
	m_HitCount = 0;
	m_ThreadCount = 1;  // one GPU context replaces the reference's pool of per-thread aligners
	DSSAligner *DA = new DSSAligner;
	DA->SetParams(*m_Params);
	DA->UseContext(GetContext());
	m_DAs.push_back(DA);
	OnSetup();
	}

void DBSearcher::AddChain(PDBChain *ptrChain, vector<vector<byte> > *ptrProfile, vector<byte> *ptrMuLetters)
	{
	m_DBChains.push_back(ptrChain);
	m_DBProfiles.push_back(ptrProfile);
	m_DBMuLettersVec.push_back(ptrMuLetters);
	}

ChainData DBSearcher::GetDBChainData(uint Idx) const
	{
	ChainData CD;
	CD.Chain = m_DBChains[Idx];
	CD.Profile = m_DBProfiles[Idx];
	CD.MuLetters = m_DBMuLettersVec.empty() ? 0 : m_DBMuLettersVec[Idx];
	CD.SelfRevScore = m_DBSelfRevScores.empty() ? FLT_MAX : m_DBSelfRevScores[Idx];
	return CD;
	}

void DBSearcher::UploadDB()
	{
	if (m_DBSet != 0)
		return;
	const uint N = GetDBChainCount();
	if (N == 0)
		Die("DBSearcher: empty database");
	rsk_asserta(RSK_SIZE(m_DBProfiles) == N);
	vector<ChainData> Chains(N);
	for (uint i = 0; i < N; ++i)
		Chains[i] = GetDBChainData(i);
	m_DBSet = UploadChains(GetContext(), Chains, !m_DBMuLettersVec.empty());
	}

void DBSearcher::AddStats()
	{
	Check(rsk_ctx_stats(m_Ctx, &m_LastStats));
	const rsk_stats &S = m_LastStats;
	DSSAligner::m_AlnCount += (uint)S.pairs;
	DSSAligner::m_SWCount += (uint)S.sw_pairs;
	DSSAligner::m_MuFilterInputCount += (uint)S.mu_filter_in;
	DSSAligner::m_MuFilterDiscardCount += (uint)S.mu_filter_rejected;
	DSSAligner::m_ParasailSaturateCount += (uint)S.mu_saturated;
	DSSAligner::m_XDropAlnCount += (uint)S.mkf_pairs;
	m_ProcessedPairCount += (uint)S.pairs;
	}

// dbsearcher.cpp:258-265 (-mints / -scores_are_not_evalues are command-line switches of the reference)
bool DBSearcher::Reject(DSSAligner &DA, bool Up) const
	{
	if (DA.GetEvalue(Up) > m_MaxEvalue)
		return true;
	return false;
	}

// dbsearcher.cpp:267-278
void DBSearcher::BaseOnAln(DSSAligner &DA, bool Up)
	{
	if (Reject(DA, Up))
		return;
	m_Lock.lock();
	++m_HitCount;
	DA.ToTsvColumns(m_fTsv, Up, m_Columns);
	DA.m_RowLen = m_RowLen;
	DA.ToAln(m_fAln, Up);
	DA.ToFasta2(m_fFasta2, m_Unaligned, Up);
	OnAln(DA, Up);
	m_Lock.unlock();
	}

void DBSearcher::BaseOnAlnLine(DSSAligner &DA, bool Up, const char *Line, size_t n)
	{
	if (Reject(DA, Up))
		return;
	m_Lock.lock();
	++m_HitCount;
	DSSAligner::WriteTsvLine(m_fTsv, Line, n);
	DA.m_RowLen = m_RowLen;
	DA.ToAln(m_fAln, Up);
	DA.ToFasta2(m_fFasta2, m_Unaligned, Up);
	OnAln(DA, Up);
	m_Lock.unlock();
	}

// The hits of one result block through BaseOnAln (dbsearcher.cpp:267-278), A = BlockA[hit.a] (the streamed block) or the DB chain
// hit.a (RunSelf).  Formatting a TSV line (rsk_format_tsv: a dozen number conversions + the CIGAR) is the expensive part of
// the replay - 1.3 s for the 1.85e6 lines of the SCOP40 all-vs-all - and independent per hit, so slices of the block are
// formatted on host threads, each with its own DSSAligner view; the calling thread then walks the block in order.
void DBSearcher::EmitHits(const rsk_hit *Hits, uint64_t N, const char *Pool, const vector<ChainData> *BlockA, bool BothDirections)
	{
	DSSAligner &DA = *m_DAs[0];
	auto Load = [&](DSSAligner &D, const rsk_hit &H) -> bool
		{
		if (BothDirections && DSSAligner::m_NoSelf && H.a == H.b)
			return false;   // runself.cpp:39-40
		D.FromHit(H, Pool, BlockA ? (*BlockA)[H.a] : GetDBChainData(H.a), GetDBChainData(H.b));
		return !D.m_Path.empty();
		};
	const uint64_t Slice = 1u << 16;
	const uint T = (uint)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
	vector<string> Text(T);
	vector<vector<uint32_t> > Len(T);
	for (uint64_t k0 = 0; k0 < N; k0 += Slice)
		{
		const uint64_t k1 = std::min(N, k0 + Slice), n = k1 - k0;
		const bool Pre = m_fTsv != 0 && n >= 1024;
		if (Pre)
			{
			auto Work = [&](uint t)
				{
				DSSAligner W;
				W.SetParams(*m_Params);
				string &S = Text[t];
				vector<uint32_t> &L = Len[t];
				S.clear();
				L.clear();
				for (uint64_t k = k0 + n * t / T; k < k0 + n * (t + 1) / T; ++k)
					{
					uint32_t lu = 0, ld = 0;
					if (Load(W, Hits[k]))
						{
						size_t b = S.size();
						if (BothDirections && !Reject(W, true) && W.FormatTsvColumns(S, true, m_Columns))
							lu = (uint32_t)(S.size() - b);
						b = S.size();
						if ((!BothDirections || Hits[k].a != Hits[k].b) && !Reject(W, false) && W.FormatTsvColumns(S, false, m_Columns))
							ld = (uint32_t)(S.size() - b);
						}
					L.push_back(lu);
					L.push_back(ld);
					}
				};
			vector<std::thread> Th;
			for (uint t = 1; t < T; ++t)
				Th.emplace_back(Work, t);
			Work(0);
			for (std::thread &t : Th)
				t.join();
			}
		for (uint t = 0; t < (Pre ? T : 1u); ++t)
			{
			const uint64_t a = Pre ? k0 + n * t / T : k0, b = Pre ? k0 + n * (t + 1) / T : k1;
			const char *p = Pre ? Text[t].data() : 0;
			for (uint64_t k = a; k < b; ++k)
				{
				const rsk_hit &H = Hits[k];
				const uint32_t lu = Pre ? Len[t][2 * (k - a)] : 0, ld = Pre ? Len[t][2 * (k - a) + 1] : 0;
				if (Load(DA, H))
					{
					if (Pre)
						{
						if (BothDirections)
							BaseOnAlnLine(DA, true, p, lu);
						if (!BothDirections || H.a != H.b)
							BaseOnAlnLine(DA, false, p + lu, ld);
						}
					else
						{
						if (BothDirections)
							BaseOnAln(DA, true);
						if (!BothDirections || H.a != H.b)
							BaseOnAln(DA, false);
						}
					}
				if (Pre)
					p += lu + ld;
				}
			}
		}
	}

// runself.cpp:48-57 (-global): the same pairs through AlignQueryTarget_Global, emitted when the global path is not empty
// (i.e. unless the Mu filter rejected the pair).  Rows of the pair triangle go to the GPU in chunks of about a million pairs.
void DBSearcher::RunSelfGlobal()
	{
	DSSAligner &DA = *m_DAs[0];
	rsk_ctx *C = GetContext();
	const uint N = GetDBChainCount();
	vector<uint32_t> ia, ib;
	auto Flush = [&]()
		{
		if (ia.empty())
			return;
		rsk_results *Res = 0;
		Check(rsk_align_global(C, m_DBSet, m_DBSet, ia.size(), ia.data(), ib.data(), &Res));
		const uint64_t n = rsk_results_count(Res);
		const rsk_hit *Hits = rsk_results_hits(Res);
		const char *Pool = rsk_results_paths(Res);
		DSSAligner::m_AlnCount += (uint)n;
		for (uint64_t k = 0; k < n; ++k)
			{
			const rsk_hit &H = Hits[k];
			if (DSSAligner::m_NoSelf && H.a == H.b)
				continue;
			DA.FromHit(H, Pool, GetDBChainData(H.a), GetDBChainData(H.b));
			if (DA.m_GlobalPath.empty())
				continue;
			BaseOnAln(DA, true);
			if (H.a != H.b)
				BaseOnAln(DA, false);
			}
		rsk_results_free(Res);
		ia.clear();
		ib.clear();
		};
	for (uint i = 0; i < N; ++i)
		{
		for (uint j = i; j < N; ++j)
			{
			ia.push_back(i);
			ib.push_back(j);
			}
		if (ia.size() >= (1u << 20))
			Flush();
		}
	Flush();
	}

// runself.cpp:72-145: pairs (i, j >= i), A = chain i, B = chain j, both directions emitted (runself.cpp:60-66)
void DBSearcher::RunSelf()
	{
	rsk_asserta(!m_DAs.empty());
	DSSAligner &DA = *m_DAs[0];
	DA.SetParams(*m_Params);
	time_t t_start = time(0);
	rsk_ctx *C = GetContext();
	rsk_params R;
	m_Params->ToRsk(R, m_MaxEvalue);
	Check(rsk_ctx_set_params(C, &R));
	UploadDB();
	if (m_Global)
		{
		RunSelfGlobal();
		m_ProcessedQueryCount = GetDBChainCount();
		m_Secs = (uint)(time(0) - t_start);
		if (m_Secs == 0)
			m_Secs = 1;
		RunStats();
		return;
		}
	rsk_search_opts O;
	memset(&O, 0, sizeof(O));
	O.keep = RSK_KEEP_HITS;
	O.want_paths = 1;
	rsk_results *Res = 0;
	double tp = NowMs();
	SetupRanks();
	if (m_RankCtx.size() > 1)
		{
		// rows of the pair triangle interleaved over the GPUs, hits gathered on the first (rsk_search_self_sharded)
		const uint N = RSK_SIZE(m_RankCtx);
		vector<rsk_stats> Stats(N);
		vector<string> Errors(N);
		auto Work = [&](uint r)
			{
			rsk_results *Mine = 0;
			if (rsk_ctx_set_params(m_RankCtx[r], &R) != RSK_OK ||
			  rsk_search_self_sharded(m_RankCtx[r], m_RankComm[r], r == 0 ? m_DBSet : m_RankDB[r], &O, 0, &Mine) != RSK_OK)
				Errors[r] = rsk_last_error();
			rsk_ctx_stats(m_RankCtx[r], &Stats[r]);
			if (r == 0)
				Res = Mine;
			};
		vector<std::thread> Threads;
		for (uint r = 1; r < N; ++r)
			Threads.emplace_back(Work, r);
		Work(0);
		for (std::thread &t : Threads)
			t.join();
		for (uint r = 0; r < N; ++r)
			if (!Errors[r].empty())
				Die("reseek_b200 (GPU %u): %s", r, Errors[r].c_str());
		for (uint r = 0; r < N; ++r)
			{
			const rsk_stats &S = Stats[r];
			DSSAligner::m_AlnCount += (uint)S.pairs;
			DSSAligner::m_SWCount += (uint)S.sw_pairs;
			DSSAligner::m_MuFilterInputCount += (uint)S.mu_filter_in;
			DSSAligner::m_MuFilterDiscardCount += (uint)S.mu_filter_rejected;
			DSSAligner::m_ParasailSaturateCount += (uint)S.mu_saturated;
			DSSAligner::m_XDropAlnCount += (uint)S.mkf_pairs;
			m_ProcessedPairCount += (uint)S.pairs;
			}
		m_LastStats = Stats[0];
		Phase("RunSelf: rsk_search_self_sharded", tp);
		}
	else
		{
		// one GPU: the sharded entry with no communicator - hits are compacted on the device and only they cross PCIe
		// (rsk_search_self reads every pair record back: 4 GB for SCOP40's 6.3e7 pairs)
		Check(rsk_search_self_sharded(C, 0, m_DBSet, &O, 0, &Res));
		Phase("RunSelf: rsk_search_self_sharded (one rank)", tp);
		AddStats();
		}
	const uint64_t N = rsk_results_count(Res);
	const rsk_hit *Hits = rsk_results_hits(Res);
	const char *Pool = rsk_results_paths(Res);
	EmitHits(Hits, N, Pool, 0, true);
	Phase("RunSelf: BaseOnAln over the hits", tp);
	rsk_results_free(Res);
	m_ProcessedQueryCount = GetDBChainCount();
	m_Secs = (uint)(time(0) - t_start);
	if (m_Secs == 0)
		m_Secs = 1;
	RunStats();
	}

// runquery.cpp:18-80 for one block of streamed chains: every chain of the block (A, "query" slot) against every
// in-memory chain (B); hits are emitted with Up = false (runquery.cpp:72-73)
void DBSearcher::RunQueryBlock(const vector<ChainData> &Block)
	{
	if (m_RankCtx.size() > 1)
		{
		RunQueryBlockSharded(Block);
		return;
		}
	rsk_ctx *C = GetContext();
	rsk_search_opts O;
	memset(&O, 0, sizeof(O));
	O.keep = RSK_KEEP_HITS;
	O.want_paths = 1;
	const bool WithMu = !m_DBMuLettersVec.empty();
	rsk_chainset *A = UploadChains(C, Block, WithMu);
	rsk_results *Res = 0;
	Check(rsk_search_cross(C, A, m_DBSet, &O, &Res));
	AddStats();
	const uint64_t N = rsk_results_count(Res);
	const rsk_hit *Hits = rsk_results_hits(Res);
	const char *Pool = rsk_results_paths(Res);
	EmitHits(Hits, N, Pool, &Block, false);
	rsk_results_free(Res);
	rsk_chainset_free(A);
	m_ProcessedQueryCount += RSK_SIZE(Block);
	}

void DBSearcher::BeginRun()
	{
	rsk_asserta(!m_DAs.empty());
	m_DAs[0]->SetParams(*m_Params);
	rsk_params R;
	m_Params->ToRsk(R, m_MaxEvalue);
	Check(rsk_ctx_set_params(GetContext(), &R));
	UploadDB();
	SetupRanks();
	}

uint ResolveGpuCount(uint Requested, int FirstDevice)
	{
	const int Visible = rsk_device_count();
	uint N = Requested;
	if (N == 0)
		{
		const char *e = getenv("RSK_GPUS");
		N = (e != 0 && atoi(e) > 0) ? (uint)atoi(e) : (uint)std::max(1, Visible - FirstDevice);
		}
	if ((int)N + FirstDevice > Visible)
		Die("%u GPUs requested from device %d on, %d visible", N, FirstDevice, Visible);
	return N;
	}

// One context + communicator + replica of the in-memory chains per additional GPU (rank 0 = the searcher's own context)
void DBSearcher::SetupRanks()
	{
	if (!m_RankCtx.empty())
		return;
	const uint N = ResolveGpuCount(m_GpuCount, m_Device);
	m_RankCtx.assign(1, GetContext());
	if (N <= 1)
		return;
	rsk_params R;
	m_Params->ToRsk(R, m_MaxEvalue);
	const uint NC = GetDBChainCount();
	vector<ChainData> Chains(NC);
	for (uint i = 0; i < NC; ++i)
		Chains[i] = GetDBChainData(i);
	m_RankDB.assign(N, (rsk_chainset *)0);
	for (uint r = 1; r < N; ++r)
		{
		rsk_ctx *C = 0;
		Check(rsk_ctx_create(m_Device + (int)r, &R, 0, &C));
		m_RankCtx.push_back(C);
		m_RankDB[r] = UploadChains(C, Chains, !m_DBMuLettersVec.empty());
		}
	m_RankComm.assign(N, (rsk_comm *)0);
	Check(rsk_comm_create_all(m_RankCtx.data(), (int)N, m_RankComm.data()));
	}

// runquery.cpp:82-125 sharded: rank r searches its share of the block against its replica of the in-memory chains; rank 0
// receives every rank's hits (indices in the numbering of the whole block) and replays them through BaseOnAln.
void DBSearcher::RunQueryBlockSharded(const vector<ChainData> &Block)
	{
	const uint N = RSK_SIZE(m_RankCtx);
	const uint NB = RSK_SIZE(Block);
	vector<uint32_t> Len(NB);
	for (uint i = 0; i < NB; ++i)
		Len[i] = Block[i].Chain->GetSeqLength();
	vector<uint32_t> Bounds(N + 1);
	Check(rsk_partition_by_residues(Len.data(), NB, (int)N, Bounds.data()));
	const bool WithMu = !m_DBMuLettersVec.empty();
	rsk_results *Res = 0;
	vector<rsk_stats> Stats(N);
	vector<string> Errors(N);
	auto Work = [&](uint r)
		{
		rsk_ctx *C = m_RankCtx[r];
		rsk_chainset *A = 0;
		if (Bounds[r + 1] > Bounds[r])
			{
			vector<ChainData> Mine(Block.begin() + Bounds[r], Block.begin() + Bounds[r + 1]);
			A = UploadChains(C, Mine, WithMu);
			}
		rsk_search_opts O;
		memset(&O, 0, sizeof(O));
		O.keep = RSK_KEEP_HITS;
		O.want_paths = 1;
		rsk_results *Mine = 0;
		if (rsk_search_cross_sharded(C, m_RankComm[r], A, r == 0 ? m_DBSet : m_RankDB[r], Bounds[r], &O, 0, &Mine) != RSK_OK)
			Errors[r] = rsk_last_error();
		rsk_ctx_stats(C, &Stats[r]);
		if (r == 0)
			Res = Mine;
		rsk_chainset_free(A);
		};
	vector<std::thread> Threads;
	for (uint r = 1; r < N; ++r)
		Threads.emplace_back(Work, r);
	Work(0);
	for (std::thread &t : Threads)
		t.join();
	for (uint r = 0; r < N; ++r)
		if (!Errors[r].empty())
			Die("reseek_b200 (GPU %u): %s", r, Errors[r].c_str());
	for (uint r = 0; r < N; ++r)
		{
		const rsk_stats &S = Stats[r];
		DSSAligner::m_AlnCount += (uint)S.pairs;
		DSSAligner::m_SWCount += (uint)S.sw_pairs;
		DSSAligner::m_MuFilterInputCount += (uint)S.mu_filter_in;
		DSSAligner::m_MuFilterDiscardCount += (uint)S.mu_filter_rejected;
		DSSAligner::m_ParasailSaturateCount += (uint)S.mu_saturated;
		DSSAligner::m_XDropAlnCount += (uint)S.mkf_pairs;
		m_ProcessedPairCount += (uint)S.pairs;
		}
	m_LastStats = Stats[0];
	const uint64_t NH = rsk_results_count(Res);
	const rsk_hit *Hits = rsk_results_hits(Res);
	const char *Pool = rsk_results_paths(Res);
	EmitHits(Hits, NH, Pool, &Block, false);
	rsk_results_free(Res);
	m_ProcessedQueryCount += NB;
	}

// runquery.cpp:82-130.  The stream is consumed in blocks of m_BlockChains chains.
void DBSearcher::RunQuery(ChainSource &QCR)
	{
	time_t t_start = time(0);
	BeginRun();
	vector<ChainData> Block;
	for (;;)
		{
		Block.clear();
		ChainData CD;
		while (RSK_SIZE(Block) < m_BlockChains && QCR.GetNext(CD))
			Block.push_back(CD);
		if (Block.empty())
			break;
		RunQueryBlock(Block);
		}
	m_Secs = (uint)(time(0) - t_start);
	if (m_Secs == 0)
		m_Secs = 1;
	RunStats();
	}

// dbsearcher.cpp:242-256: Mu letters only when the filter is on; self-reverse scores with omega = 0 (profileloader.cpp:22-26)
void DBSearcher::LoadDB(const string &DBFN)
	{
	rsk_asserta(m_Params != 0);
	rsk_asserta(m_DBChains.empty());
	ChainReader2 CR;
	CR.Open(DBFN);
	DSSParams LoaderParams = *m_Params;
	LoaderParams.m_UsePara = false;
	LoaderParams.m_Omega = 0;
	ChainFeatures F;
	const bool WithMu = m_Params->m_Omega > 0;
	double tp = NowMs();
	rsk_ctx *C = GetContext();
	Phase("LoadDB: context (CUDA init)", tp);
	ProfileLoader::Load(*m_Params, CR, 0, WithMu, C, LoaderParams, m_MaxEvalue, F);
	Phase("LoadDB: read + DSS + self-reverse", tp);
	m_DBChains = F.Chains;
	m_DBProfiles = F.Profiles;
	if (WithMu)
		{
		m_DBMuLettersVec = F.MuLetters;
		m_DBMuKmersVec = F.MuKmers;
		}
	else
		{
		for (auto *p : F.MuLetters) delete p;
		for (auto *p : F.MuKmers) delete p;
		}
	m_DBSelfRevScores = F.SelfRevScores;
	F.Release();
	m_OwnsChains = true;
	}

// runquery.cpp:18-130 with the chains coming from a file: per block DSS on host threads, self-reverse scores with the
// search parameters (runquery.cpp:43), then the cross search of the block
void DBSearcher::RunQuery(ChainReader2 &QCR)
	{
	time_t t_start = time(0);
	BeginRun();
	const bool WithMu = !m_DBMuLettersVec.empty();
	// The next block is read, run through DSS and given its self-reverse scores (on a context of its own) while the
	// current one is being searched and its hits written.  RSK_NO_OVERLAP=1: one block at a time on the search context.
	const bool Overlap = getenv("RSK_NO_OVERLAP") == 0;
	rsk_ctx *LoadCtx = GetContext();
	if (Overlap)
		{
		if (m_LoaderCtx == 0)
			{
			rsk_params R;
			m_Params->ToRsk(R, m_MaxEvalue);
			Check(rsk_ctx_create(m_Device, &R, 0, &m_LoaderCtx));
			}
		LoadCtx = m_LoaderCtx;
		}
	struct Loaded
		{
		ChainFeatures F;
		vector<ChainData> Block;
		};
	auto LoadNext = [&]() -> std::unique_ptr<Loaded>
		{
		std::unique_ptr<Loaded> L(new Loaded);
		const uint N = ProfileLoader::Load(*m_Params, QCR, m_BlockChains, WithMu, LoadCtx, *m_Params, m_MaxEvalue, L->F);
		if (N == 0)
			return nullptr;
		L->Block.resize(N);
		for (uint i = 0; i < N; ++i)
			{
			L->Block[i].Chain = L->F.Chains[i];
			L->Block[i].Profile = L->F.Profiles[i];
			L->Block[i].MuLetters = WithMu ? L->F.MuLetters[i] : 0;
			L->Block[i].SelfRevScore = L->F.SelfRevScores[i];
			}
		return L;
		};
	std::unique_ptr<Loaded> Cur = LoadNext();
	while (Cur)
		{
		std::future<std::unique_ptr<Loaded> > Next;
		if (Overlap)
			Next = std::async(std::launch::async, LoadNext);
		RunQueryBlock(Cur->Block);
		Cur->F.Free();
		Cur = Overlap ? Next.get() : LoadNext();
		}
	m_Secs = (uint)(time(0) - t_start);
	if (m_Secs == 0)
		m_Secs = 1;
	RunStats();
	}

// dbsearcher.cpp:29-56
void DBSearcher::RunStats() const
	{
	uint Secs = m_Secs == 0 ? 1 : m_Secs;
	fprintf(stderr, "\n");
	if (m_MaxEvalue == DBL_MAX)
		fprintf(stderr, "%10u  Hits\n", (uint)m_HitCount);
	else
		fprintf(stderr, "%10u  Hits (max E-value %.3g)\n", (uint)m_HitCount, m_MaxEvalue);
	if (m_ProcessedQueryCount < 100)
		return;
	fprintf(stderr, "%10u  DB chains\n", (uint)m_ProcessedQueryCount);
	fprintf(stderr, "%10u  Query chains\n", GetDBChainCount());
	fprintf(stderr, "%10.1f  Chains/sec\n", (double)m_ProcessedQueryCount / Secs);
	fprintf(stderr, "%10.3g  Comparisons/sec\n", (double)m_ProcessedPairCount / Secs);
	DSSAligner::Stats();
	}

// ---- `-search Q -db DB -fast` (search.cpp:76-111) ----
void MuPreFilter(const DSSParams &Params, const vector<ChainData> &Query, const vector<ChainData> &DB,
  const string &OutputFN, int Device)
	{
	rsk_params R;
	Params.ToRsk(R, 10);
	rsk_ctx *C = 0;
	Check(rsk_ctx_create(Device, &R, 0, &C));
	rsk_chainset *Q = UploadChains(C, Query, true);
	rsk_chainset *T = UploadChains(C, DB, true);
	rsk_prefilter_opts PO;
	memset(&PO, 0, sizeof(PO));
	rsk_prefilter_result *PR = 0;
	Check(rsk_prefilter(C, Q, T, &PO, &PR));
	long long n = rsk_prefilter_to_tsv(PR, 0, 0);
	vector<char> Buf((size_t)(-n) + 1);
	n = rsk_prefilter_to_tsv(PR, Buf.data(), Buf.size());
	if (n < 0)
		Die("rsk_prefilter_to_tsv failed");
	FILE *f = fopen(OutputFN.c_str(), "w");
	if (f == 0)
		Die("Cannot create %s", OutputFN.c_str());
	fwrite(Buf.data(), 1, (size_t)n, f);
	fclose(f);
	rsk_prefilter_free(PR);
	rsk_chainset_free(Q);
	rsk_chainset_free(T);
	rsk_ctx_destroy(C);
	}

// postmufilter.cpp:211-301; scan loop :116-208; Accept :105-114 with the default thresholds (E <= 10)
void PostMuFilter(const DSSParams &Params, const string &MuFilterTsvFN, const vector<ChainData> &Query,
  const vector<ChainData> &DB, const string &HitsFN, const char *Columns, int Device, const string &AlnFN, double MaxEvalue)
	{
	FILE *fIn = fopen(MuFilterTsvFN.c_str(), "r");
	if (fIn == 0)
		Die("Cannot open %s", MuFilterTsvFN.c_str());
	vector<uint32_t> ia, ib;
	char Tag[64];
	uint TargetCount = 0;
	if (fscanf(fIn, "%63s %u", Tag, &TargetCount) != 2 || strcmp(Tag, "prefilter") != 0)
		Die("%s: bad prefilter TSV header", MuFilterTsvFN.c_str());  // postmufilter.cpp:236-242
	for (uint t = 0; t < TargetCount; ++t)
		{
		uint TargetIdx, FilHitCount;
		if (fscanf(fIn, "%u %u", &TargetIdx, &FilHitCount) != 2)
			Die("%s: truncated", MuFilterTsvFN.c_str());
		rsk_asserta(TargetIdx < RSK_SIZE(DB));
		for (uint k = 0; k < FilHitCount; ++k)
			{
			uint QueryIdx;
			if (fscanf(fIn, "%u", &QueryIdx) != 1)
				Die("%s: truncated", MuFilterTsvFN.c_str());
			rsk_asserta(QueryIdx < RSK_SIZE(Query));
			ia.push_back(QueryIdx);   // A = the query bag, B = the DB chain (postmufilter.cpp:190)
			ib.push_back(TargetIdx);
			}
		}
	fclose(fIn);

	rsk_params R;
	Params.ToRsk(R, MaxEvalue);
	rsk_ctx *C = 0;
	Check(rsk_ctx_create(Device, &R, 0, &C));
	rsk_chainset *Q = UploadChains(C, Query, true);
	rsk_chainset *T = UploadChains(C, DB, true);
	rsk_search_opts O;
	memset(&O, 0, sizeof(O));
	O.keep = RSK_KEEP_ALL;   // line order of the TSV, as the reference's single-threaded scan emits
	O.want_paths = 1;
	rsk_results *Res = 0;
	Check(rsk_search_pairs(C, Q, T, ia.size(), ia.data(), ib.data(), &O, &Res));
	FILE *fOut = HitsFN.empty() ? 0 : fopen(HitsFN.c_str(), "w");
	if (!HitsFN.empty() && fOut == 0)
		Die("Cannot create %s", HitsFN.c_str());
	FILE *fAln = AlnFN.empty() ? 0 : fopen(AlnFN.c_str(), "w");
	if (!AlnFN.empty() && fAln == 0)
		Die("Cannot create %s", AlnFN.c_str());
	DSSAligner DA;
	DA.SetParams(Params);
	DA.UseContext(C);
	const uint64_t N = rsk_results_count(Res);
	const rsk_hit *Hits = rsk_results_hits(Res);
	const char *Pool = rsk_results_paths(Res);
	for (uint64_t k = 0; k < N; ++k)
		{
		const rsk_hit &H = Hits[k];
		DA.FromHit(H, Pool, Query[H.a], DB[H.b]);
		if (DA.m_EvalueA <= MaxEvalue)  // Accept(): s_MaxEvalue = opt(evalue) or 10 (postmufilter.cpp:105-114, 217-220)
			{
			DA.ToTsvColumns(fOut, true, Columns);
			DA.ToAln(fAln, true);
			}
		}
	if (fOut != 0)
		fclose(fOut);
	if (fAln != 0)
		fclose(fAln);
	rsk_results_free(Res);
	rsk_chainset_free(Q);
	rsk_chainset_free(T);
	rsk_ctx_destroy(C);
	}

// `-search Q -db DB -fast` with the DB block-partitioned over the GPUs (see dbsearcher.h)
void SearchFastDB(const DSSParams &Params, const vector<ChainData> &Query, const vector<ChainData> &DB, const string &HitsFN,
  const char *Columns, const string &AlnFN, double MaxEvalue, uint GpuCount, const string &CandTsvFN)
	{
	const uint N = ResolveGpuCount(GpuCount, 0);
	const uint NT = RSK_SIZE(DB);
	vector<uint32_t> Len(NT);
	for (uint i = 0; i < NT; ++i)
		Len[i] = DB[i].Chain->GetSeqLength();
	vector<uint32_t> Bounds(N + 1);
	Check(rsk_partition_by_residues(Len.data(), NT, (int)N, Bounds.data()));
	rsk_params R;
	Params.ToRsk(R, MaxEvalue);
	vector<rsk_ctx *> Ctx(N, (rsk_ctx *)0);
	vector<rsk_comm *> Comm(N, (rsk_comm *)0);
	for (uint r = 0; r < N; ++r)
		Check(rsk_ctx_create((int)r, &R, 0, &Ctx[r]));
	Check(rsk_comm_create_all(Ctx.data(), (int)N, Comm.data()));
	rsk_results *Res = 0;
	rsk_prefilter_result *Cands = 0;
	vector<string> Errors(N);
	auto Work = [&](uint r)
		{
		rsk_chainset *Q = UploadChains(Ctx[r], Query, true);
		rsk_chainset *T = 0;
		if (Bounds[r + 1] > Bounds[r])
			{
			vector<ChainData> Mine(DB.begin() + Bounds[r], DB.begin() + Bounds[r + 1]);
			T = UploadChains(Ctx[r], Mine, true);
			}
		rsk_prefilter_opts PO;
		memset(&PO, 0, sizeof(PO));
		rsk_search_opts O;
		memset(&O, 0, sizeof(O));
		O.keep = RSK_KEEP_HITS;
		O.want_paths = 1;
		rsk_results *Mine = 0;
		rsk_prefilter_result *MyCands = 0;
		if (rsk_search_fast_db_sharded(Ctx[r], Comm[r], Q, T, Bounds[r], &PO, &O, 0, &Mine, r == 0 ? &MyCands : 0) != RSK_OK)
			Errors[r] = rsk_last_error();
		if (r == 0)
			{
			Res = Mine;
			Cands = MyCands;
			}
		rsk_chainset_free(T);
		rsk_chainset_free(Q);
		};
	vector<std::thread> Threads;
	for (uint r = 1; r < N; ++r)
		Threads.emplace_back(Work, r);
	Work(0);
	for (std::thread &t : Threads)
		t.join();
	for (uint r = 0; r < N; ++r)
		if (!Errors[r].empty())
			Die("reseek_b200 (GPU %u): %s", r, Errors[r].c_str());
	if (!CandTsvFN.empty() && Cands != 0)
		{
		long long n = rsk_prefilter_to_tsv(Cands, 0, 0);
		vector<char> Buf((size_t)(-n) + 1);
		n = rsk_prefilter_to_tsv(Cands, Buf.data(), Buf.size());
		FILE *f = fopen(CandTsvFN.c_str(), "w");
		if (f == 0 || n < 0)
			Die("Cannot create %s", CandTsvFN.c_str());
		fwrite(Buf.data(), 1, (size_t)n, f);
		fclose(f);
		}
	rsk_prefilter_free(Cands);
	FILE *fOut = HitsFN.empty() ? 0 : fopen(HitsFN.c_str(), "w");
	if (!HitsFN.empty() && fOut == 0)
		Die("Cannot create %s", HitsFN.c_str());
	FILE *fAln = AlnFN.empty() ? 0 : fopen(AlnFN.c_str(), "w");
	if (!AlnFN.empty() && fAln == 0)
		Die("Cannot create %s", AlnFN.c_str());
	DSSAligner DA;
	DA.SetParams(Params);
	DA.UseContext(Ctx[0]);
	const uint64_t NH = rsk_results_count(Res);
	const rsk_hit *Hits = rsk_results_hits(Res);
	const char *Pool = rsk_results_paths(Res);
	// line order of the candidate TSV: target ascending, query ascending
	vector<uint64_t> Order(NH);
	std::iota(Order.begin(), Order.end(), (uint64_t)0);
	std::sort(Order.begin(), Order.end(), [&](uint64_t x, uint64_t y)
		{
		return Hits[x].b != Hits[y].b ? Hits[x].b < Hits[y].b : Hits[x].a < Hits[y].a;
		});
	for (uint64_t k = 0; k < NH; ++k)
		{
		const rsk_hit &H = Hits[Order[k]];
		DA.FromHit(H, Pool, Query[H.a], DB[H.b]);
		if (DA.m_EvalueA <= MaxEvalue)
			{
			DA.ToTsvColumns(fOut, true, Columns);
			DA.ToAln(fAln, true);
			}
		}
	if (fOut != 0)
		fclose(fOut);
	if (fAln != 0)
		fclose(fAln);
	rsk_results_free(Res);
	for (uint r = 0; r < N; ++r)
		rsk_comm_destroy(Comm[r]);
	for (uint r = 0; r < N; ++r)
		rsk_ctx_destroy(Ctx[r]);
	}

namespace {
// both files as in-memory chains: DSS on host threads, self-reverse scores on the GPU with the given parameters
struct LoadedFiles
	{
	ChainFeatures Q, T;
	vector<ChainData> QD, TD;
	LoadedFiles(const DSSParams &Params, const string &QueryFN, const string &DBFN)
		{
		rsk_params R;
		Params.ToRsk(R, 10);
		rsk_ctx *C = 0;
		Check(rsk_ctx_create(0, &R, 0, &C));
		ChainReader2 QR, TR;
		QR.Open(QueryFN);
		TR.Open(DBFN);
		ProfileLoader::Load(Params, QR, 0, true, C, Params, 10, Q);
		ProfileLoader::Load(Params, TR, 0, true, C, Params, 10, T);
		rsk_ctx_destroy(C);
		Fill(Q, QD);
		Fill(T, TD);
		}
	~LoadedFiles() { Q.Free(); T.Free(); }
	static void Fill(const ChainFeatures &F, vector<ChainData> &v)
		{
		v.resize(F.Chains.size());
		for (size_t i = 0; i < v.size(); ++i)
			{
			v[i].Chain = F.Chains[i];
			v[i].Profile = F.Profiles[i];
			v[i].MuLetters = F.MuLetters[i];
			v[i].SelfRevScore = F.SelfRevScores.empty() ? FLT_MAX : F.SelfRevScores[i];
			}
		}
	};
}  // namespace

void MuPreFilter(const DSSParams &Params, const string &QueryCAFN, const string &DBBCAFN, const string &OutputFN)
	{
	LoadedFiles L(Params, QueryCAFN, DBBCAFN);
	MuPreFilter(Params, L.QD, L.TD, OutputFN);
	}

void PostMuFilter(const DSSParams &Params, const string &MuFilterTsvFN, const string &QueryCAFN, const string &DBBCAFN,
  const string &HitsFN)
	{
	LoadedFiles L(Params, QueryCAFN, DBBCAFN);
	PostMuFilter(Params, MuFilterTsvFN, L.QD, L.TD, HitsFN);
	}

}  // namespace reseek_b200
