// host_driver.cpp - small command-line harness over the DSSAligner / DBSearcher look-alikes.
// It exists for the parity tests (tests/test_host_shim.py) and as a usage example of the drop-in surface; it is not
// the reference's CLI.  Chains are read from ".rskc" dumps (feature letters, Mu letters, coordinates: the output of
// the reference's DSS stage, written by reseek_b200/chainio.py) because structure I/O and DSS are upstream of this layer.
//
//   rsk_host_demo self  <mode> <set.rskc> <out.tsv> [columns]       DBSearcher::RunSelf           (-search X)
//   rsk_host_demo query <mode> <stream.rskc> <db.rskc> <out.tsv>    DBSearcher::RunQuery          (-search Q -db DB)
//   rsk_host_demo pair  <mode> <set.rskc> <i> <j> <out.tsv>         DSSAligner::AlignQueryTarget  (-alignpair)
//   rsk_host_demo pairbags <mode> <set.rskc> <i> <j> <out.tsv>            DSSAligner::AlignBags (PostMuFilter's call)
//   rsk_host_demo pairglobal <mode> <set.rskc> <i> <j> <out.tsv> [columns]  AlignQueryTarget_Global (-global)
//   rsk_host_demo fastdb <q.rskc> <db.rskc> <cands.tsv> <out.tsv>   MuPreFilter + PostMuFilter    (-search Q -db DB -fast)
// From the reference's own .bca files, through the DSS look-alike (no precomputed features):
//   rsk_host_demo features   <in.bca> <out.rskc>                               DSS only (device DSS, dss_kernel.cu)
//   rsk_host_demo selfsearch <mode> <x.bca> <out.tsv> [columns]                search.cpp:20-38   (-search X)
//   rsk_host_demo search     <mode> <q.bca> <db.bca> <out.tsv> [columns]       search.cpp:40-63   (-search Q -db DB)
//   rsk_host_demo searchfast <q.bca> <db.bca> <cands.tsv> <out.tsv> [columns]  search.cpp:76-111  (-search Q -db DB -fast)
#include <stdlib.h>
#include <string.h>

#include "dbsearcher.h"
#include "scop40bench.h"

using namespace reseek_b200;

struct LoadedSet
	{
	vector<PDBChain> Chains;
	vector<vector<vector<byte> > > Profiles;
	vector<vector<byte> > Mus;
	vector<vector<uint> > Kmers;   // non-empty marker vectors: the k-mers themselves are derived on the device
	vector<float> SelfRevs;
	vector<ChainData> Data;
	};

static void ReadOrDie(void *p, size_t n, FILE *f, const char *fn)
	{
	if (n != 0 && fread(p, 1, n, f) != n)
		Die("%s: truncated", fn);
	}

static void Load(const char *fn, LoadedSet &S)
	{
	FILE *f = fopen(fn, "rb");
	if (f == 0)
		Die("Cannot open %s", fn);
	char magic[4];
	uint32_t ver, n, has_mu, has_sr;
	uint64_t total;
	ReadOrDie(magic, 4, f, fn);
	if (memcmp(magic, "RSKC", 4) != 0)
		Die("%s: not a .rskc file", fn);
	ReadOrDie(&ver, 4, f, fn); ReadOrDie(&n, 4, f, fn); ReadOrDie(&total, 8, f, fn);
	ReadOrDie(&has_mu, 4, f, fn); ReadOrDie(&has_sr, 4, f, fn);
	vector<uint32_t> len(n);
	ReadOrDie(len.data(), 4 * (size_t)n, f, fn);
	vector<uint8_t> prof(8 * total), mu(has_mu ? total : 0), seq(total);
	vector<float> xyz(3 * total), sr(n, FLT_MAX);
	ReadOrDie(prof.data(), prof.size(), f, fn);
	ReadOrDie(mu.data(), mu.size(), f, fn);
	ReadOrDie(xyz.data(), 4 * xyz.size(), f, fn);
	if (has_sr)
		ReadOrDie(sr.data(), 4 * (size_t)n, f, fn);
	ReadOrDie(seq.data(), total, f, fn);
	S.Chains.resize(n); S.Profiles.resize(n); S.Mus.resize(n); S.Kmers.resize(n); S.SelfRevs = sr; S.Data.resize(n);
	uint64_t off = 0;
	for (uint32_t i = 0; i < n; ++i)
		{
		string Label;
		for (int c; (c = fgetc(f)) > 0;)
			Label.push_back((char)c);
		const uint32_t L = len[i];
		PDBChain &C = S.Chains[i];
		C.m_Label = Label;
		C.m_Seq.assign((const char *)&seq[off], L);
		C.m_Xs.assign(&xyz[off], &xyz[off] + L);
		C.m_Ys.assign(&xyz[total + off], &xyz[total + off] + L);
		C.m_Zs.assign(&xyz[2 * total + off], &xyz[2 * total + off] + L);
		S.Profiles[i].resize(8);
		for (int ft = 0; ft < 8; ++ft)
			S.Profiles[i][ft].assign(&prof[ft * total + off], &prof[ft * total + off] + L);
		if (has_mu)
			{
			S.Mus[i].assign(&mu[off], &mu[off] + L);
			S.Kmers[i].assign(L >= 3 ? L - 2 : 0, 0u);
			}
		off += L;
		}
	fclose(f);
	for (uint32_t i = 0; i < n; ++i)
		{
		S.Data[i].Chain = &S.Chains[i];
		S.Data[i].Profile = &S.Profiles[i];
		S.Data[i].MuLetters = has_mu ? &S.Mus[i] : 0;
		S.Data[i].SelfRevScore = S.SelfRevs[i];
		}
	}

static ALGO_MODE ParseMode(const char *s)
	{
	if (!strcmp(s, "fast")) return AM_Fast;
	if (!strcmp(s, "sensitive")) return AM_Sensitive;
	if (!strcmp(s, "verysensitive")) return AM_VerySensitive;
	Die("Must set -fast, -sensitive or -verysensitive");
	}

static void FillSearcher(DBSearcher &DBS, LoadedSet &S)
	{
	const bool HasMu = !S.Mus.empty() && S.Data[0].MuLetters != 0;
	for (size_t i = 0; i < S.Chains.size(); ++i)
		{
		DBS.m_DBChains.push_back(&S.Chains[i]);
		DBS.m_DBProfiles.push_back(&S.Profiles[i]);
		if (HasMu)
			{
			DBS.m_DBMuLettersVec.push_back(&S.Mus[i]);
			DBS.m_DBMuKmersVec.push_back(&S.Kmers[i]);
			}
		DBS.m_DBSelfRevScores.push_back(S.SelfRevs[i]);
		}
	}

// an OnAln subclass in the style of scop40bench.cpp: counts what the searcher reports
class CountingSearcher : public DBSearcher
	{
public:
	uint m_OnAlnCount = 0;
	void OnAln(DSSAligner &DA, bool Up) override { ++m_OnAlnCount; }
	};

static void WriteRskc(const char *fn, const ChainFeatures &F, bool HasMu, bool HasSelfRev)
	{
	FILE *f = fopen(fn, "wb");
	if (f == 0)
		Die("Cannot create %s", fn);
	const uint32_t n = RSK_SIZE(F.Chains), ver = 1, has_mu = HasMu, has_sr = HasSelfRev;
	uint64_t total = 0;
	vector<uint32_t> len(n);
	for (uint32_t i = 0; i < n; ++i)
		total += (len[i] = F.Chains[i]->GetSeqLength());
	fwrite("RSKC", 1, 4, f);
	fwrite(&ver, 4, 1, f); fwrite(&n, 4, 1, f); fwrite(&total, 8, 1, f); fwrite(&has_mu, 4, 1, f); fwrite(&has_sr, 4, 1, f);
	fwrite(len.data(), 4, n, f);
	for (int ft = 0; ft < 8; ++ft)
		for (uint32_t i = 0; i < n; ++i)
			fwrite((*F.Profiles[i])[ft].data(), 1, len[i], f);
	if (HasMu)
		for (uint32_t i = 0; i < n; ++i)
			fwrite(F.MuLetters[i]->data(), 1, len[i], f);
	for (int c = 0; c < 3; ++c)
		for (uint32_t i = 0; i < n; ++i)
			{
			const vector<float> &v = c == 0 ? F.Chains[i]->m_Xs : c == 1 ? F.Chains[i]->m_Ys : F.Chains[i]->m_Zs;
			fwrite(v.data(), 4, len[i], f);
			}
	if (HasSelfRev)
		fwrite(F.SelfRevScores.data(), 4, n, f);
	for (uint32_t i = 0; i < n; ++i)
		fwrite(F.Chains[i]->m_Seq.data(), 1, len[i], f);
	for (uint32_t i = 0; i < n; ++i)
		fwrite(F.Chains[i]->m_Label.c_str(), 1, F.Chains[i]->m_Label.size() + 1, f);
	fclose(f);
	}

static vector<ChainData> ToChainData(const ChainFeatures &F)
	{
	vector<ChainData> v(F.Chains.size());
	for (size_t i = 0; i < v.size(); ++i)
		{
		v[i].Chain = F.Chains[i];
		v[i].Profile = F.Profiles[i];
		v[i].MuLetters = F.MuLetters[i];
		v[i].SelfRevScore = F.SelfRevScores.empty() ? FLT_MAX : F.SelfRevScores[i];
		}
	return v;
	}

// -aln / -fasta2 / -unaligned / -rowlen of the reference, taken from the environment (the positional arguments are the
// files the tests always need)
static void OpenAlnOutputs(DBSearcher &DBS)
	{
	if (const char *e = getenv("RSK_ALN"))
		if ((DBS.m_fAln = fopen(e, "w")) == 0) Die("Cannot create %s", e);
	if (const char *e = getenv("RSK_FASTA2"))
		if ((DBS.m_fFasta2 = fopen(e, "w")) == 0) Die("Cannot create %s", e);
	if (const char *e = getenv("RSK_UNALIGNED"))
		DBS.m_Unaligned = atoi(e) != 0;
	if (const char *e = getenv("RSK_ROWLEN"))
		DBS.m_RowLen = (uint)atoi(e);
	if (const char *e = getenv("RSK_GLOBAL"))
		DBS.m_Global = atoi(e) != 0;
	}
static void CloseAlnOutputs(DBSearcher &DBS)
	{
	if (DBS.m_fAln) fclose(DBS.m_fAln);
	if (DBS.m_fFasta2) fclose(DBS.m_fFasta2);
	}

// The search command with the reference's own command line (cmd_search, search.cpp:20-111; options from myopts.h):
//   rsk_host_demo -search Q.bca [-db DB.bca] -fast|-sensitive|-verysensitive [-output hits.tsv] [-columns a+b+c]
//                 [-aln FILE] [-fasta2 FILE] [-unaligned] [-rowlen N] [-global] [-evalue E] [-noself] [-threads N] [-gpus N]
// .bca inputs only; -threads is accepted and ignored (GPU contexts do the aligning); -gpus N shards the -db side over N GPUs
// (default: all visible devices).
static int ReseekSearch(int argc, char **argv)
	{
	string QFN, DBFN, OutFN, Columns, AlnFN, Fasta2FN;
	bool Unaligned = false, Global = false, NoSelf = false, HaveEvalue = false;
	uint RowLen = 0, Gpus = 0;
	double Evalue = 10;
	int Mode = -1;
	for (int i = 1; i < argc; ++i)
		{
		const string a = argv[i];
		auto Value = [&]() -> const char *
			{
			if (i + 1 >= argc)
				Die("Missing value for %s", a.c_str());
			return argv[++i];
			};
		if (a == "-search") QFN = Value();
		else if (a == "-db") DBFN = Value();
		else if (a == "-output") OutFN = Value();
		else if (a == "-columns") Columns = Value();
		else if (a == "-aln") AlnFN = Value();
		else if (a == "-fasta2") Fasta2FN = Value();
		else if (a == "-rowlen") RowLen = (uint)atoi(Value());
		else if (a == "-evalue") { Evalue = atof(Value()); HaveEvalue = true; }
		else if (a == "-threads") Value();
		else if (a == "-gpus") Gpus = (uint)atoi(Value());
		else if (a == "-unaligned") Unaligned = true;
		else if (a == "-global") Global = true;
		else if (a == "-noself") NoSelf = true;
		else if (a == "-fast") Mode = AM_Fast;
		else if (a == "-sensitive") Mode = AM_Sensitive;
		else if (a == "-verysensitive") Mode = AM_VerySensitive;
		else
			Die("Unknown option %s", a.c_str());
		}
	if (Mode < 0)
		Die("Must set -fast, -sensitive or -verysensitive");  // dssparams.cpp:90
	DSSAligner::m_NoSelf = NoSelf;
	DSSParams Params;
	Params.SetMode((ALGO_MODE)Mode);
	if (!DBFN.empty() && Mode == AM_Fast)
		{
		// search.cpp:76-111: prefilter, then the post-filter under the sensitive preset
		if (DBFN.size() < 4 || DBFN.compare(DBFN.size() - 4, 4, ".bca") != 0)
			Die(".bca format required for -db");
		// options this branch cannot honour are refused rather than dropped (the reference's PostMuFilter writes -output and
		// -aln only, postmufilter.cpp:190-194, 244)
		if (!Fasta2FN.empty() || NoSelf || Global)
			Die("-fasta2, -noself and -global are not supported with -db ... -fast");
		DSSParams Params2;
		Params2.SetDSSParams(DM_AlwaysSensitive);
		{
		rsk_params R;
		Params2.ToRsk(R, Evalue);
		rsk_ctx *C = 0;
		if (rsk_ctx_create(0, &R, 0, &C) != RSK_OK)
			Die("reseek_b200: %s", rsk_last_error());
		ChainReader2 QR, TR;
		QR.Open(QFN);
		TR.Open(DBFN);
		ChainFeatures Q, T;
		ProfileLoader::Load(Params2, QR, 0, true, C, Params2, Evalue, Q);
		ProfileLoader::Load(Params2, TR, 0, true, C, Params2, Evalue, T);
		rsk_ctx_destroy(C);
		const vector<ChainData> QD = ToChainData(Q), TD = ToChainData(T);
		// prefilter + merged bag + post-filter, the DB block-partitioned over the GPUs (-gpus N; default: all visible)
		SearchFastDB(Params2, QD, TD, OutFN, Columns.empty() ? 0 : Columns.c_str(), AlnFN, Evalue, Gpus);
		Q.Free();
		T.Free();
		}
		return 0;
		}
	CountingSearcher DBS;
	DBS.m_Params = &Params;
	if (HaveEvalue)
		DBS.m_MaxEvalue = (float)Evalue;   // dbsearcher.cpp:75-76
	DBS.LoadDB(QFN);               // the -search file is the in-memory side (search.cpp:53)
	DBS.Setup();
	if (HaveEvalue)
		DBS.m_MaxEvalue = (float)Evalue;
	if (!OutFN.empty() && (DBS.m_fTsv = fopen(OutFN.c_str(), "w")) == 0) Die("Cannot create %s", OutFN.c_str());
	if (!AlnFN.empty() && (DBS.m_fAln = fopen(AlnFN.c_str(), "w")) == 0) Die("Cannot create %s", AlnFN.c_str());
	if (!Fasta2FN.empty() && (DBS.m_fFasta2 = fopen(Fasta2FN.c_str(), "w")) == 0) Die("Cannot create %s", Fasta2FN.c_str());
	if (!Columns.empty())
		DBS.m_Columns = Columns.c_str();
	DBS.m_Unaligned = Unaligned;
	DBS.m_RowLen = RowLen;
	DBS.m_Global = Global;
	DBS.m_GpuCount = Gpus;
	if (const char *e = getenv("RSK_BLOCK_CHAINS"))
		DBS.m_BlockChains = (uint)atoi(e);
	if (DBFN.empty())
		DBS.RunSelf();             // SelfSearch, search.cpp:20-40
	else
		{
		ChainReader2 CR;
		CR.Open(DBFN);             // Search_NoMuFilter, search.cpp:42-63
		DBS.RunQuery(CR);
		}
	if (DBS.m_fTsv) fclose(DBS.m_fTsv);
	CloseAlnOutputs(DBS);
	return 0;
	}

// cmd_alignpair (alignpair.cpp:164-228) with the reference's command line:
//   rsk_host_demo -alignpair Q.bca -input2 T.bca [-aln FILE] [-global] [-output FILE] [-output2 FILE]
// Sensitive preset with the Mu filter off (alignpair.cpp:177-185); every chain of Q against every chain of T, the best
// scoring pair (first maximum, Q-major) is reported.  All pairs are one GPU batch here.  -output / -output2 carry the
// query's ATOM lines after the superposition; .bca files have no ATOM lines, so those files come out empty, as they do
// from the reference for .bca inputs.
static int ReseekAlignPair(int argc, char **argv)
	{
	string QFN, TFN, AlnFN, OutFN, Out2FN;
	bool Global = false;
	for (int i = 1; i < argc; ++i)
		{
		const string a = argv[i];
		auto Value = [&]() -> const char *
			{
			if (i + 1 >= argc)
				Die("Missing value for %s", a.c_str());
			return argv[++i];
			};
		if (a == "-alignpair") QFN = Value();
		else if (a == "-input2") TFN = Value();
		else if (a == "-aln") AlnFN = Value();
		else if (a == "-output") OutFN = Value();
		else if (a == "-output2") Out2FN = Value();
		else if (a == "-threads") Value();
		else if (a == "-global") Global = true;
		else if (a == "-fast" || a == "-sensitive" || a == "-verysensitive") {}  // overridden (alignpair.cpp:177-181)
		else
			Die("Unknown option %s", a.c_str());
		}
	if (TFN.empty())
		Die("Must specify -input2");
	DSSParams Params;
	Params.SetDSSParams(DM_AlwaysSensitive);
	Params.m_UsePara = false;
	Params.m_Omega = 0;
	rsk_params R;
	Params.ToRsk(R, DBL_MAX);
	rsk_ctx *C = 0;
	if (rsk_ctx_create(0, &R, 0, &C) != RSK_OK)
		Die("reseek_b200: %s", rsk_last_error());
	ChainReader2 QR, TR;
	QR.Open(QFN);
	TR.Open(TFN);
	ChainFeatures Q, T;
	ProfileLoader::Load(Params, QR, 0, false, C, Params, DBL_MAX, Q);
	ProfileLoader::Load(Params, TR, 0, false, C, Params, DBL_MAX, T);
	if (Q.Chains.empty()) Die("No chains found in %s", QFN.c_str());
	if (T.Chains.empty()) Die("No chains found in %s", TFN.c_str());
	const vector<ChainData> QD = ToChainData(Q), TD = ToChainData(T);
	rsk_chainset *QS = UploadChains(C, QD, false), *TS = UploadChains(C, TD, false);
	rsk_results *Res = 0;
	int rc;
	if (Global)
		{
		vector<uint32_t> ia, ib;
		for (uint32_t q = 0; q < QD.size(); ++q)
			for (uint32_t t = 0; t < TD.size(); ++t)
				{
				ia.push_back(q);
				ib.push_back(t);
				}
		rc = rsk_align_global(C, QS, TS, ia.size(), ia.data(), ib.data(), &Res);
		}
	else
		{
		rsk_search_opts O;
		memset(&O, 0, sizeof(O));
		O.keep = RSK_KEEP_ALL;
		O.want_paths = 1;
		rc = rsk_search_cross(C, QS, TS, &O, &Res);
		}
	if (rc != RSK_OK)
		Die("reseek_b200: %s", rsk_last_error());
	const uint64_t N = rsk_results_count(Res);
	const rsk_hit *Hits = rsk_results_hits(Res);
	float BestScore = -9999;
	uint64_t Best = 0;
	rsk_asserta(N == (uint64_t)QD.size() * TD.size());
	for (uint64_t k = 0; k < N; ++k)   // alignpair.cpp:203-218: Q-major order, first strict maximum
		{
		rsk_asserta(Hits[k].a == k / TD.size() && Hits[k].b == k % TD.size());
		if (Hits[k].score > BestScore)
			{
			BestScore = Hits[k].score;
			Best = k;
			}
		}
	if (BestScore == 0)
		Die("No alignment found");
	DSSAligner DA;
	DA.SetParams(Params);
	DA.UseContext(C);
	// AlignPair1 (alignpair.cpp:77-105) aligns each chain against its reversed self right before the pair; on the -global
	// path that is where the AQ of the printed block comes from (ClearAlign does not reset the quality)
	const uint bq = Hits[Best].a, bt = Hits[Best].b;
	{
	DSS D;
	D.UseContext(C);
	PDBChain Rev;
	vector<vector<byte> > RevProfile;
	Q.Chains[bq]->GetReverse(Rev);
	D.Init(Rev);
	D.GetProfile(RevProfile);
	GetSelfRevScore(DA, *Q.Chains[bq], *Q.Profiles[bq], RevProfile, 0, 0);
	T.Chains[bt]->GetReverse(Rev);
	D.Init(Rev);
	D.GetProfile(RevProfile);
	GetSelfRevScore(DA, *T.Chains[bt], *T.Profiles[bt], RevProfile, 0, 0);
	}
	DA.FromHit(Hits[Best], rsk_results_paths(Res), QD[bq], TD[bt]);
	if (!AlnFN.empty())
		{
		FILE *f = fopen(AlnFN.c_str(), "w");
		if (f == 0) Die("Cannot create %s", AlnFN.c_str());
		DA.ToAln(f, true);
		fclose(f);
		}
	double t3[3], u[3][3];
	DA.GetKabsch(t3, u, true);     // alignpair.cpp:129-131
	fprintf(stderr, "%s %s score %.4g, superposition t = %.3f %.3f %.3f\n", DA.GetLabel(true), DA.GetLabel(false), BestScore, t3[0], t3[1], t3[2]);
	for (const string *FN : {&OutFN, &Out2FN})
		if (!FN->empty())
			{
			FILE *f = fopen(FN->c_str(), "w");  // no ATOM lines in a .bca file
			if (f == 0) Die("Cannot create %s", FN->c_str());
			fclose(f);
			}
	rsk_results_free(Res);
	rsk_chainset_free(QS);
	rsk_chainset_free(TS);
	Q.Free();
	T.Free();
	rsk_ctx_destroy(C);
	return 0;
	}

int main(int argc, char **argv)
	{
	if (argc < 2)
		Die("usage: rsk_host_demo self|query|pair|fastdb|features|selfsearch|search|searchfast ...");
	const string Cmd = argv[1];
	if (Cmd == "-search")
		return ReseekSearch(argc, argv);
	if (Cmd == "-alignpair")
		return ReseekAlignPair(argc, argv);
	if (Cmd == "-scop40bench" && argc >= 4)
		{
		// cmd_scop40bench (scop40bench.cpp:767-823): rsk_host_demo -scop40bench X.bca -lookup dom_scopid.tsv -fast|-sensitive|-verysensitive
		string LookupFN;
		int Mode = -1;
		for (int i = 3; i < argc; ++i)
			{
			const string a = argv[i];
			if (a == "-lookup" && i + 1 < argc) LookupFN = argv[++i];
			else if (a == "-fast") Mode = AM_Fast;
			else if (a == "-sensitive") Mode = AM_Sensitive;
			else if (a == "-verysensitive") Mode = AM_VerySensitive;
			else Die("Unknown option %s", a.c_str());
			}
		if (Mode < 0)
			Die("Must set -fast, -sensitive or -verysensitive");
		if (LookupFN.empty())
			Die("-lookup FILE (domain <tab> scop id) required");
		DSSParams Params;
		Params.SetMode((ALGO_MODE) Mode);
		SCOP40Bench SB;
		SB.m_Params = &Params;
		SB.ReadLookup(LookupFN);
		SB.LoadDB(argv[2]);
		SB.Setup();
		SB.m_QuerySelf = true;
		SB.RunSelf();
		SB.SetStats();
		SB.WriteSummary(stdout);
		return 0;
		}
	if (Cmd == "bcarange" && argc >= 5)
		{
		// the block of a .bca that rank argv[3] of argv[4] processes reads (ChainReader2::OpenRange): "lo hi first-label last-label residues"
		ChainReader2 CR;
		CR.OpenRange(argv[2], (uint) atoi(argv[3]), (uint) atoi(argv[4]));
		uint Count = 0;
		uint64_t Residues = 0;
		string First, Last;
		while (PDBChain *Chain = CR.GetNext())
			{
			if (Count++ == 0)
				First = Chain->m_Label;
			Last = Chain->m_Label;
			Residues += Chain->GetSeqLength();
			delete Chain;
			}
		printf("%u %u %s %s %llu\n", CR.GetFirstIdx(), CR.GetFirstIdx() + Count, First.c_str(), Last.c_str(), (unsigned long long) Residues);
		return 0;
		}
	if (Cmd == "features" && argc >= 4)
		{
		// DSS stage only (on the GPU, chain by chain through the DSS look-alike; no self-reverse scores)
		ChainReader2 CR;
		CR.Open(argv[2]);
		ChainFeatures F;
		DSSParams Params;
		DSS D;
		D.SetParams(Params);
		while (PDBChain *Chain = CR.GetNext())
			{
			F.Chains.push_back(Chain);
			F.Profiles.push_back(new vector<vector<byte> >);
			F.MuLetters.push_back(new vector<byte>);
			F.MuKmers.push_back(new vector<uint>);
			D.Init(*Chain);
			D.GetProfile(*F.Profiles.back());
			D.GetMuLetters(*F.MuLetters.back());
			}
		WriteRskc(argv[3], F, true, false);
		F.Free();
		return 0;
		}
	if (Cmd == "selfsearch" && argc >= 5)
		{
		DSSParams Params;
		Params.SetMode(ParseMode(argv[2]));
		CountingSearcher DBS;
		DBS.m_Params = &Params;
		DBS.LoadDB(argv[3]);
		DBS.Setup();
		DBS.m_fTsv = fopen(argv[4], "w");
		if (argc > 5)
			DBS.m_Columns = argv[5];
		OpenAlnOutputs(DBS);
		DBS.RunSelf();
		fclose(DBS.m_fTsv);
		CloseAlnOutputs(DBS);
		fprintf(stderr, "OnAln calls %u, hits %u\n", DBS.m_OnAlnCount, (uint)DBS.m_HitCount);
		return 0;
		}
	if (Cmd == "search" && argc >= 6)
		{
		DSSParams Params;
		Params.SetMode(ParseMode(argv[2]));
		CountingSearcher DBS;
		DBS.m_Params = &Params;
		DBS.LoadDB(argv[3]);       // the -search file is the in-memory side (search.cpp:53)
		DBS.Setup();
		DBS.m_fTsv = fopen(argv[5], "w");
		if (argc > 6)
			DBS.m_Columns = argv[6];
		if (const char *e = getenv("RSK_BLOCK_CHAINS"))  // the tests ask for several blocks even on small sets
			DBS.m_BlockChains = (uint)atoi(e);
		ChainReader2 CR;
		CR.Open(argv[4]);          // the -db file is streamed (search.cpp:57-59)
		OpenAlnOutputs(DBS);
		DBS.RunQuery(CR);
		fclose(DBS.m_fTsv);
		CloseAlnOutputs(DBS);
		fprintf(stderr, "OnAln calls %u, hits %u\n", DBS.m_OnAlnCount, (uint)DBS.m_HitCount);
		return 0;
		}
	if (Cmd == "searchfast" && argc >= 6)
		{
		DSSParams Params;
		Params.SetMode(AM_Fast);
		DSSParams Params2;
		Params2.SetDSSParams(DM_AlwaysSensitive);  // search.cpp:106-108
		rsk_params R;
		Params2.ToRsk(R, 10);
		rsk_ctx *C = 0;
		if (rsk_ctx_create(0, &R, 0, &C) != RSK_OK)
			Die("reseek_b200: %s", rsk_last_error());
		// queries and targets with the self-reverse scores PostMuFilter computes for them (postmufilter.cpp:79-81, 166-167:
		// the sensitive parameters)
		ChainReader2 QR, TR;
		QR.Open(argv[2]);
		TR.Open(argv[3]);
		ChainFeatures Q, T;
		ProfileLoader::Load(Params2, QR, 0, true, C, Params2, 10, Q);
		ProfileLoader::Load(Params2, TR, 0, true, C, Params2, 10, T);
		const vector<ChainData> QD = ToChainData(Q), TD = ToChainData(T);
		MuPreFilter(Params, QD, TD, argv[4]);
		PostMuFilter(Params2, argv[4], QD, TD, argv[5], argc > 6 ? argv[6] : 0, 0, getenv("RSK_ALN") ? getenv("RSK_ALN") : "");
		Q.Free();
		T.Free();
		rsk_ctx_destroy(C);
		return 0;
		}
	if (Cmd == "searchfastfiles" && argc >= 6)   // search.cpp:76-111 through the file-name signatures
		{
		DSSParams Params;
		Params.SetMode(AM_Fast);
		DSSParams Params2;
		Params2.SetDSSParams(DM_AlwaysSensitive);
		MuPreFilter(Params, string(argv[2]), string(argv[3]), string(argv[4]));
		PostMuFilter(Params2, string(argv[4]), string(argv[2]), string(argv[3]), string(argv[5]));
		return 0;
		}
	if (Cmd == "self" && argc >= 5)
		{
		DSSParams Params;
		Params.SetMode(ParseMode(argv[2]));
		LoadedSet S;
		Load(argv[3], S);
		CountingSearcher DBS;
		DBS.m_Params = &Params;
		FillSearcher(DBS, S);
		DBS.m_fTsv = fopen(argv[4], "w");
		if (argc > 5)
			DBS.m_Columns = argv[5];
		DBS.Setup();
		OpenAlnOutputs(DBS);
		DBS.RunSelf();
		fclose(DBS.m_fTsv);
		CloseAlnOutputs(DBS);
		fprintf(stderr, "OnAln calls %u, hits %u\n", DBS.m_OnAlnCount, (uint)DBS.m_HitCount);
		return 0;
		}
	if (Cmd == "query" && argc >= 6)
		{
		DSSParams Params;
		Params.SetMode(ParseMode(argv[2]));
		LoadedSet Stream, DB;
		Load(argv[3], Stream);
		Load(argv[4], DB);
		CountingSearcher DBS;
		DBS.m_Params = &Params;
		FillSearcher(DBS, DB);
		DBS.m_fTsv = fopen(argv[5], "w");
		if (const char *e = getenv("RSK_BLOCK_CHAINS"))  // the tests ask for several blocks even on small sets
			DBS.m_BlockChains = (uint)atoi(e);
		DBS.Setup();
		VectorChainSource Src(Stream.Data);
		DBS.RunQuery(Src);
		fclose(DBS.m_fTsv);
		fprintf(stderr, "OnAln calls %u, hits %u\n", DBS.m_OnAlnCount, (uint)DBS.m_HitCount);
		return 0;
		}
	if (Cmd == "pair" && argc >= 7)
		{
		DSSParams Params;
		Params.SetMode(ParseMode(argv[2]));
		LoadedSet S;
		Load(argv[3], S);
		const uint i = (uint)atoi(argv[4]), j = (uint)atoi(argv[5]);
		rsk_asserta(i < S.Chains.size() && j < S.Chains.size());
		DSSAligner DA;
		DA.SetParams(Params);
		DA.SetQuery(S.Chains[i], &S.Profiles[i], S.Data[i].MuLetters, S.Data[i].MuLetters ? &S.Kmers[i] : 0, S.SelfRevs[i]);
		DA.SetTarget(S.Chains[j], &S.Profiles[j], S.Data[j].MuLetters, S.Data[j].MuLetters ? &S.Kmers[j] : 0, S.SelfRevs[j]);
		DA.AlignQueryTarget();
		FILE *f = fopen(argv[6], "w");
		if (!DA.m_Path.empty())   // alignpair.cpp:110-117
			DA.ToTsv(f, true);
		fclose(f);
		return 0;
		}
	if (Cmd == "pairbags" && argc >= 7)   // postmufilter.cpp:185-195: one candidate through DSSAligner::AlignBags
		{
		DSSParams Params;
		Params.SetMode(ParseMode(argv[2]));
		LoadedSet S;
		Load(argv[3], S);
		const uint i = (uint)atoi(argv[4]), j = (uint)atoi(argv[5]);
		rsk_asserta(i < S.Chains.size() && j < S.Chains.size());
		ChainBag BagA, BagB;
		BagA.m_ptrChain = &S.Chains[i]; BagA.m_ptrProfile = &S.Profiles[i]; BagA.m_ptrMuLetters = S.Data[i].MuLetters;
		BagA.m_ptrMuKmers = S.Data[i].MuLetters ? &S.Kmers[i] : 0; BagA.m_SelfRevScore = S.SelfRevs[i];
		BagB.m_ptrChain = &S.Chains[j]; BagB.m_ptrProfile = &S.Profiles[j]; BagB.m_ptrMuLetters = S.Data[j].MuLetters;
		BagB.m_ptrMuKmers = S.Data[j].MuLetters ? &S.Kmers[j] : 0; BagB.m_SelfRevScore = S.SelfRevs[j];
		DSSAligner DA;
		DA.SetParams(Params);
		DA.AlignBags(BagA, BagB);
		FILE *f = fopen(argv[6], "w");
		if (!DA.m_Path.empty())
			DA.ToTsv(f, true);
		fclose(f);
		return 0;
		}
	if (Cmd == "pairglobal" && argc >= 7)   // alignpair.cpp:110-114 with -global
		{
		DSSParams Params;
		Params.SetMode(ParseMode(argv[2]));
		LoadedSet S;
		Load(argv[3], S);
		const uint i = (uint)atoi(argv[4]), j = (uint)atoi(argv[5]);
		rsk_asserta(i < S.Chains.size() && j < S.Chains.size());
		DSSAligner DA;
		DA.SetParams(Params);
		DA.SetQuery(S.Chains[i], &S.Profiles[i], S.Data[i].MuLetters, S.Data[i].MuLetters ? &S.Kmers[i] : 0, S.SelfRevs[i]);
		DA.SetTarget(S.Chains[j], &S.Profiles[j], S.Data[j].MuLetters, S.Data[j].MuLetters ? &S.Kmers[j] : 0, S.SelfRevs[j]);
		DA.AlignQueryTarget_Global();
		FILE *f = fopen(argv[6], "w");
		if (!DA.m_GlobalPath.empty())   // runself.cpp:50-56
			DA.ToTsvColumns(f, true, argc > 7 ? argv[7] : "query+target+gscore+dpscore+cigar");
		fclose(f);
		return 0;
		}
	if (Cmd == "fastdb" && argc >= 6)
		{
		LoadedSet Q, DB;
		Load(argv[2], Q);
		Load(argv[3], DB);
		DSSParams Params;
		Params.SetMode(AM_Fast);
		MuPreFilter(Params, Q.Data, DB.Data, argv[4]);
		DSSParams Params2;
		Params2.SetDSSParams(DM_AlwaysSensitive);  // search.cpp:106-108
		PostMuFilter(Params2, argv[4], Q.Data, DB.Data, argv[5]);
		return 0;
		}
	Die("bad command line");
	}
