// bcadata.h - reader / writer of the reference's binary C-alpha format (.bca; bcadata.h:1-40, bcadata.cpp:15-232), SURVEY §8(f) row 2.
//
// Layout: u32 magic 0xBCABCA | u64 chain count | u64 offset of the length table | u64 size of the label data |
// per chain: L amino-acid chars + 3L uint16 integer coordinates (x+1000)*10+0.5 (pdbchain.h:89-90) |
// u32 lengths[count] | NUL-terminated labels.
#pragma once

#include <mutex>

#include "reseek_compat.h"

namespace reseek_b200 {

const uint32_t BCA_MAGIC = 0xBCABCA;

class BCAData
	{
public:
	// integer <-> float coordinates (pdbchain.h:89-90): 0.1 A resolution, offset 1000 A
	static uint16_t CoordToIC(float X) { return uint16_t((X + 1000)*10 + 0.5); }
	static float ICToCoord(uint16_t IC) { return float(IC/10.0f) - 1000; }

	~BCAData() { Clear(); }

	// reading
	void Open(const string &FN);
	void ReadChain(uint64_t ChainIdx, PDBChain &Chain) const;
	uint GetChainCount() const { return RSK_SIZE(m_Labels); }
	uint GetSeqLength(uint64_t ChainIdx) const { return m_SeqLengths[ChainIdx]; }

	// writing
	void Create(const string &FN);
	void WriteChain(const PDBChain &Chain);

	void Close();
	void Clear();

public:
	string m_FN;
	FILE *m_f = 0;
	bool m_Reading = false;
	bool m_Writing = false;
	vector<string> m_Labels;         // one per chain
	vector<uint32_t> m_SeqLengths;   // one per chain
	vector<uint64_t> m_Offsets;      // file offset of each chain record (sequence, then integer coordinates)
	mutable std::mutex m_ReadLock;   // ReadChain may be called from several loader threads
	};

}  // namespace reseek_b200
