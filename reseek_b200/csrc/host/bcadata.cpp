// bcadata.cpp - see bcadata.h
#include "bcadata.h"

#include <string.h>

namespace reseek_b200 {

static void ReadAt(FILE *f, uint64_t Pos, void *Buf, size_t Bytes, const char *FN)
	{
	if (fseeko(f, (off_t) Pos, SEEK_SET) != 0 || (Bytes != 0 && fread(Buf, 1, Bytes, f) != Bytes))
		Die("%s: read of %llu bytes at %llu failed", FN, (unsigned long long) Bytes, (unsigned long long) Pos);
	}

void BCAData::Clear()
	{
	m_Labels.clear();
	m_Offsets.clear();
	m_SeqLengths.clear();
	m_FN.clear();
	if (m_f != 0)
		fclose(m_f);
	m_f = 0;
	m_Writing = false;
	m_Reading = false;
	}

void BCAData::Open(const string &FN)  // bcadata.cpp:60-120
	{
	if (FN == "")
		Die("Empty BCA filename");
	rsk_asserta(!m_Writing && !m_Reading && m_f == 0);
	m_FN = FN;
	m_f = fopen(FN.c_str(), "rb");
	if (m_f == 0)
		Die("Cannot open %s", FN.c_str());
	uint32_t Magic = 0;
	uint64_t Header[3];
	ReadAt(m_f, 0, &Magic, sizeof(Magic), FN.c_str());
	if (Magic != BCA_MAGIC)
		Die("Bad magic %08lx, invalid .bca file '%s'", (unsigned long) Magic, FN.c_str());
	ReadAt(m_f, sizeof(Magic), Header, sizeof(Header), FN.c_str());
	const uint64_t ChainCount = Header[0], SeqLengthsPos = Header[1], LabelDataSize = Header[2];
	rsk_asserta(ChainCount == uint64_t(uint(ChainCount)));
	m_SeqLengths.resize(ChainCount);
	ReadAt(m_f, SeqLengthsPos, m_SeqLengths.data(), sizeof(uint32_t)*ChainCount, FN.c_str());
	uint64_t Offset = sizeof(Magic) + sizeof(Header);
	m_Offsets.reserve(ChainCount);
	for (uint64_t i = 0; i < ChainCount; ++i)
		{
		m_Offsets.push_back(Offset);
		Offset += 7*uint64_t(m_SeqLengths[i]);
		}
	vector<char> LabelData(LabelDataSize + 1, 0);
	ReadAt(m_f, SeqLengthsPos + sizeof(uint32_t)*ChainCount, LabelData.data(), LabelDataSize, FN.c_str());
	m_Labels.clear();
	for (uint64_t Pos = 0; Pos < LabelDataSize && m_Labels.size() < ChainCount;)
		{
		m_Labels.push_back(string(LabelData.data() + Pos));
		Pos += m_Labels.back().size() + 1;
		}
	if (m_Labels.size() != ChainCount)
		Die("Bad BCA file, %u chains %u labels", uint(ChainCount), RSK_SIZE(m_Labels));
	m_Reading = true;
	}

void BCAData::ReadChain(uint64_t ChainIdx, PDBChain &Chain) const  // bcadata.cpp:191-232
	{
	rsk_asserta(m_Reading && !m_Writing && ChainIdx < m_SeqLengths.size());
	const uint L = m_SeqLengths[ChainIdx];
	vector<char> Seq(L + 1, 0);
	vector<uint16_t> ICs(3*size_t(L));
		{
		std::lock_guard<std::mutex> Guard(m_ReadLock);
		ReadAt(m_f, m_Offsets[ChainIdx], Seq.data(), L, m_FN.c_str());
		ReadAt(m_f, m_Offsets[ChainIdx] + L, ICs.data(), 6*size_t(L), m_FN.c_str());
		}
	Chain.m_Label = m_Labels[ChainIdx];
	Chain.m_Seq = string(Seq.data());  // a NUL inside the sequence truncates it, as in the reference
	Chain.m_Xs.clear(); Chain.m_Ys.clear(); Chain.m_Zs.clear();
	Chain.m_Xs.reserve(L); Chain.m_Ys.reserve(L); Chain.m_Zs.reserve(L);
	for (uint i = 0; i < L; ++i)  // pdbchain.cpp:435-449
		{
		Chain.m_Xs.push_back(ICToCoord(ICs[3*i]));
		Chain.m_Ys.push_back(ICToCoord(ICs[3*i+1]));
		Chain.m_Zs.push_back(ICToCoord(ICs[3*i+2]));
		}
	}

void BCAData::Create(const string &FN)  // bcadata.cpp:15-33
	{
	if (FN == "")
		Die("Empty BCA filename");
	rsk_asserta(!m_Writing && !m_Reading && m_f == 0);
	m_FN = FN;
	m_f = fopen(FN.c_str(), "wb");
	if (m_f == 0)
		Die("Cannot create %s", FN.c_str());
	const uint64_t Placeholder[3] = {0, 0, 0};
	fwrite(&BCA_MAGIC, sizeof(BCA_MAGIC), 1, m_f);
	fwrite(Placeholder, sizeof(Placeholder), 1, m_f);
	m_Writing = true;
	}

void BCAData::WriteChain(const PDBChain &Chain)  // bcadata.cpp:35-58
	{
	rsk_asserta(m_Writing && !m_Reading);
	const uint L = Chain.GetSeqLength();
	m_Offsets.push_back(uint64_t(ftello(m_f)));
	m_Labels.push_back(Chain.m_Label);
	m_SeqLengths.push_back(L);
	vector<uint16_t> ICs;
	ICs.reserve(3*size_t(L));
	for (uint i = 0; i < L; ++i)
		{
		ICs.push_back(CoordToIC(Chain.m_Xs[i]));
		ICs.push_back(CoordToIC(Chain.m_Ys[i]));
		ICs.push_back(CoordToIC(Chain.m_Zs[i]));
		}
	fwrite(Chain.m_Seq.c_str(), 1, L, m_f);
	fwrite(ICs.data(), sizeof(uint16_t), ICs.size(), m_f);
	}

void BCAData::Close()  // bcadata.cpp:5-13, 147-175
	{
	if (m_Writing && !m_Reading)
		{
		const uint64_t ChainCount = m_Labels.size();
		const uint64_t SeqLengthsPos = uint64_t(ftello(m_f));
		fwrite(m_SeqLengths.data(), sizeof(uint32_t), ChainCount, m_f);
		uint64_t LabelDataSize = 0;
		for (const string &Label : m_Labels)
			{
			fwrite(Label.c_str(), 1, Label.size() + 1, m_f);
			LabelDataSize += Label.size() + 1;
			}
		const uint64_t Header[3] = {ChainCount, SeqLengthsPos, LabelDataSize};
		fseeko(m_f, sizeof(BCA_MAGIC), SEEK_SET);
		fwrite(Header, sizeof(Header), 1, m_f);
		}
	else if (!(m_Reading && !m_Writing))
		Die("BCAData::Close(), not open");
	Clear();
	}

}  // namespace reseek_b200
