// rsk_format.cu - host-side hit formatting, byte-compatible with the reference's writers.
// Replaces DSSAligner::ToTsv / WriteUserField (dssaligner.cpp:1016-1034, userfields.cpp:45-152), EvalueToStr
// (userfields.cpp:19-30), PathToCIGAR (cigar.cpp:95-139), GetQCovPct/GetTCovPct (dssaligner.cpp:1119-1141),
// GetPctId (dssaligner.cpp:1325-1369).  Pure host code: no device work happens here.
#include <ctype.h>
#include <float.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <string>

#include "../../include/reseek_b200.h"

namespace {

// cigar.cpp:95-139: run-length encode; D and I are exchanged unless FlipDI (= Up) is set
void path_to_cigar(const char *path, uint32_t n, bool up, std::string &out)
{
	out.clear();
	if (n == 0)
		return;
	char buf[32];
	uint32_t run = 1;
	char last = path[0];
	auto flush = [&]() {
		char c = last;
		if (!up) {
			if (c == 'D') c = 'I';
			else if (c == 'I') c = 'D';
		}
		snprintf(buf, sizeof(buf), "%u%c", run, c);
		out += buf;
	};
	for (uint32_t i = 1; i < n; ++i) {
		if (path[i] == last) {
			++run;
			continue;
		}
		flush();
		last = path[i];
		run = 1;
	}
	flush();
}

void fmt(std::string &s, const char *f, double v)
{
	char buf[64];
	snprintf(buf, sizeof(buf), f, v);
	s += buf;
}
void fmtu(std::string &s, unsigned v)
{
	char buf[32];
	snprintf(buf, sizeof(buf), "%u", v);
	s += buf;
}

void fmts(std::string &s, const char *f, ...) __attribute__((format(printf, 2, 3)));
void fmts(std::string &s, const char *f, ...)
{
	char buf[256];
	va_list ap;
	va_start(ap, f);
	const int n = vsnprintf(buf, sizeof(buf), f, ap);
	va_end(ap);
	if (n < (int)sizeof(buf)) {
		s += buf;
		return;
	}
	std::string big((size_t)n + 1, '\0');  // long labels
	va_start(ap, f);
	vsnprintf(&big[0], big.size(), f, ap);
	va_end(ap);
	s.append(big.c_str(), (size_t)n);
}

// DSSAligner::GetRow_A / GetRow_B (dssaligner.cpp:1161-1277): one gapped row of the alignment; with global set, the
// unaligned flanks in lower case and '.' padding.  which = 0: the A row, 1: the B row; always in the A/B frame of the
// stored path, as in the reference.
void aligned_row(const rsk_hit_view *v, int which, bool global, std::string &row)
{
	const rsk_hit &h = *v->hit;
	const char *X = which == 0 ? v->seq_a : v->seq_b;
	const uint32_t lo_x = which == 0 ? h.lo_a : h.lo_b, lo_y = which == 0 ? h.lo_b : h.lo_a;
	const uint32_t len_x = which == 0 ? v->len_a : v->len_b, len_y = which == 0 ? v->len_b : v->len_a;
	const char own_gap = which == 0 ? 'I' : 'D';  // the path letter that puts a gap into this row
	row.clear();
	if (global) {
		for (uint32_t i = lo_x; i < lo_y; ++i)
			row += '.';
		for (uint32_t i = 0; i < lo_x; ++i)
			row += (char)tolower((unsigned char)X[i]);
	}
	uint32_t px = lo_x, py = lo_y;
	for (uint32_t c = 0; c < h.path_len; ++c) {
		const char ch = v->path[c];
		if (ch == 'M') {
			row += X[px++];
			++py;
		} else if (ch == own_gap) {
			row += '-';
			++py;
		} else {
			row += X[px++];
		}
	}
	if (global) {
		while (px < len_x) {
			row += (char)tolower((unsigned char)X[px++]);
			++py;
		}
		while (py++ < len_y)
			row += '.';
	}
}

// identities over the M columns, exact character comparison (DSSAligner::GetPctId, dssaligner.cpp:1325-1369)
float pct_id(const rsk_hit_view *v)
{
	const rsk_hit &h = *v->hit;
	unsigned N = 0, n = 0;
	if (v->seq_a && v->seq_b && v->path) {
		uint32_t pa = h.lo_a, pb = h.lo_b;
		for (uint32_t c = 0; c < h.path_len; ++c) {
			const char ch = v->path[c];
			if (ch == 'M') {
				++N;
				if (v->seq_a[pa] == v->seq_b[pb]) ++n;
				++pa; ++pb;
			} else if (ch == 'D') ++pa;
			else ++pb;
		}
	}
	return N == 0 ? 0.0f : (n * 100.0f) / N;
}

// sfasta.cpp:5-26: nothing at all for an empty sequence
void seq_to_fasta(std::string &out, const std::string &label, const std::string &seq, uint32_t rowlen = 80)
{
	if (seq.empty())
		return;
	out += '>';
	out += label;
	out += '\n';
	for (size_t from = 0; from < seq.size(); from += rowlen) {
		out.append(seq, from, rowlen);
		out += '\n';
	}
}

long long hand_over(const std::string &text, char *out, size_t cap)
{
	if (text.size() + 1 > cap)  // "need this many bytes": at most -17, so that it cannot be taken for an rsk_status
		return -(long long)(text.size() < 16 ? 16 : text.size()) - 1;
	memcpy(out, text.c_str(), text.size() + 1);
	return (long long)text.size();
}

}  // namespace

extern "C" int rsk_path_to_cigar(const char *path, uint32_t path_len, int up, char *out, size_t cap)
{
	if (!out || cap == 0 || (path_len && !path))
		return RSK_ERR_ARG;
	std::string c;
	path_to_cigar(path, path_len, up != 0, c);
	if (c.size() + 1 > cap)
		return RSK_ERR_LIMIT;
	memcpy(out, c.c_str(), c.size() + 1);
	return (int)c.size();
}

// One TSV line (without the newline).  up != 0: query = A, target = B; up == 0: query = B, target = A.
extern "C" int rsk_format_tsv(const rsk_hit_view *v, int up_, const char *columns, char *out, size_t cap)
{
	if (!v || !v->hit || !out || cap == 0)
		return RSK_ERR_ARG;
	const bool up = up_ != 0;
	const rsk_hit &h = *v->hit;
	const char *cols = (columns && *columns) ? columns : "query+target+qlo+qhi+ql+tlo+thi+tl+pctid+pvalue";  // usage.h:49 "std"
	std::string line, name;
	const uint32_t ql = up ? v->len_a : v->len_b, tl = up ? v->len_b : v->len_a;
	const uint32_t qlo = up ? h.lo_a : h.lo_b, qhi = up ? h.hi_a : h.hi_b;
	const uint32_t tlo = up ? h.lo_b : h.lo_a, thi = up ? h.hi_b : h.hi_a;
	bool first = true;
	for (const char *p = cols;; ++p) {
		if (*p != '+' && *p != 0) {
			name += *p;
			continue;
		}
		if (!first)
			line += '\t';
		first = false;
		if (name == "query") line += (up ? v->label_a : v->label_b) ? (up ? v->label_a : v->label_b) : "";
		else if (name == "target") line += (up ? v->label_b : v->label_a) ? (up ? v->label_b : v->label_a) : "";
		else if (name == "evalue") {  // EvalueToStr
			double E = h.evalue;
			if (E > 10) E = 99;
			if (E > 1) fmt(line, "%.1f", E);
			else if (E > 0.001) fmt(line, "%.4f", E);
			else fmt(line, "%.3g", E);
		}
		else if (name == "pvalue") fmt(line, "%.3g", h.pvalue);
		else if (name == "ql") fmtu(line, ql);
		else if (name == "tl") fmtu(line, tl);
		else if (name == "qlo") fmtu(line, qlo + 1);
		else if (name == "qhi") fmtu(line, qhi + 1);
		else if (name == "tlo") fmtu(line, tlo + 1);
		else if (name == "thi") fmtu(line, thi + 1);
		else if (name == "qcovpct") {
			double pct = ql == 0 ? 0 : (100.0 * (qhi - qlo + 1)) / ql;
			if (pct > 100) pct = 100;
			fmt(line, "%.1f", pct);
		}
		else if (name == "tcovpct") {  // NB the reference divides by the QUERY length here (dssaligner.cpp:1134)
			double pct = (100.0 * (thi - tlo + 1)) / ql;
			if (pct > 100) pct = 100;
			fmt(line, "%.1f", pct);
		}
		else if (name == "pctid") fmt(line, "%.1f", pct_id(v));
		else if (name == "ts") fmt(line, "%.3g", -FLT_MAX);  // m_TestStatisticA is cleared to -FLT_MAX and never set (dssaligner.cpp:919)
		else if (name == "muscore") {
			// the reference re-runs AlignMuQP here; the record carries what the filter computed, and a pair that took the
			// long-chain path never saw the filter
			if (h.flags & RSK_HIT_MKF)
				return RSK_ERR_ARG;
			fmt(line, "%.3g", h.mu_score);
		}
		else if (name == "qrow" || name == "trow" || name == "qrowg" || name == "trowg") {
			if (!v->seq_a || !v->seq_b || (h.path_len && !v->path))
				return RSK_ERR_ARG;
			const bool top = name[0] == 'q', global = name.size() == 5;
			std::string row;
			aligned_row(v, (up == top) ? 0 : 1, global, row);  // GetRow (dssaligner.cpp:1143-1159)
			line += row;
		}
		else if (name == "newts") fmt(line, "%.3g", h.ts);
		// m_AlnFwdScore; a record of rsk_align_global carries m_GlobalScore in score, and the local score stays 0 there
		else if (name == "raw") fmt(line, "%.3g", (h.flags & RSK_HIT_GLOBAL) ? 0.0f : h.score);
		else if (name == "dpscore") fmt(line, "%.4g", (h.flags & RSK_HIT_GLOBAL) ? 0.0f : h.score);
		else if (name == "lddt") fmt(line, "%.4g", h.lddt);
		else if (name == "ids") fmtu(line, h.ids);
		else if (name == "gaps") fmtu(line, h.gaps);
		else if (name == "aq") fmt(line, "%.4f", h.qual);
		else if (name == "muhsp") { char b[32]; snprintf(b, sizeof(b), "%d", (h.flags & RSK_HIT_MKF) ? h.mu_fwd : 0); line += b; }
		else if (name == "muchain") { char b[32]; snprintf(b, sizeof(b), "%d", (h.flags & RSK_HIT_MKF) ? h.mu_rev : 0); line += b; }
		else if (name == "gscore") fmt(line, "%.1f", (h.flags & RSK_HIT_GLOBAL) ? h.score : -9999.0f);  // m_GlobalScore; ClearAlign leaves -9999 (dssaligner.cpp:925)
		else if (name == "cigar") {
			std::string c;
			path_to_cigar(v->path, v->path ? h.path_len : 0, up, c);
			line += c;
		}
		else
			return RSK_ERR_ARG;  // the reference dies: "Invalid user field name"
		name.clear();
		if (*p == 0)
			break;
	}
	if (line.size() + 1 > cap)
		return RSK_ERR_LIMIT;
	memcpy(out, line.c_str(), line.size() + 1);
	return (int)line.size();
}

// DSSAligner::ToAln -> PrettyAln (dssaligner.cpp:965-979, prettyaln.cpp:26-99) with WriteLocalAln (writelocalaln.cpp:65-100):
// the block the reference appends to the -aln file for one hit.  rowlen = 0: the default 80 columns (-rowlen).
extern "C" long long rsk_format_aln(const rsk_hit_view *v, int up_, uint32_t rowlen, char *out, size_t cap)
{
	if (!v || !v->hit || !v->path || !v->seq_a || !v->seq_b || v->hit->path_len == 0 || (cap && !out))
		return RSK_ERR_ARG;  // (the reference asserts on an empty path)
	const rsk_hit &h = *v->hit;
	const bool up = up_ != 0;
	if (rowlen == 0)
		rowlen = 80;
	// frame of the block: A = query; with up == 0 the chains change places and D/I are exchanged (InvertPath)
	const char *la = up ? v->label_a : v->label_b, *lb = up ? v->label_b : v->label_a;
	const char *A = up ? v->seq_a : v->seq_b, *B = up ? v->seq_b : v->seq_a;
	const uint32_t LA = up ? v->len_a : v->len_b, LB = up ? v->len_b : v->len_a;
	const uint32_t lo_a = up ? h.lo_a : h.lo_b, lo_b = up ? h.lo_b : h.lo_a;
	if (!la) la = "";
	if (!lb) lb = "";
	const uint32_t ncol = h.path_len;
	std::string path(v->path, ncol);
	if (!up)
		for (char &c : path)
			c = c == 'D' ? 'I' : c == 'I' ? 'D' : c;
	uint32_t pa = lo_a, pb = lo_b, ids = 0, gaps = 0;
	for (uint32_t c = 0; c < ncol; ++c) {
		switch (path[c]) {
		case 'M':
			if (pa >= LA || pb >= LB) return RSK_ERR_ARG;
			if (A[pa] == B[pb]) ++ids;
			++pa; ++pb;
			break;
		case 'D':
			if (pa >= LA) return RSK_ERR_ARG;
			++pa; ++gaps;
			break;
		case 'I':
			if (pb >= LB) return RSK_ERR_ARG;
			++pb; ++gaps;
			break;
		default:
			return RSK_ERR_ARG;
		}
	}
	std::string t;
	t.reserve((size_t)ncol * 4 + 512);
	t += "\n_____________________________________________________________________________________________________________\n";
	uint32_t i = lo_a, j = lo_b;
	for (uint32_t from = 0; from < ncol; from += rowlen) {
		const uint32_t to = from + rowlen < ncol ? from + rowlen : ncol;  // one past the block's last column
		const uint32_t i0 = i, j0 = j;
		fmts(t, "%5u ", i + 1);
		for (uint32_t k = from; k < to; ++k)
			t += (path[k] == 'I') ? '-' : A[i++];
		fmts(t, " %u  %s\n", i, la);
		t += "      ";
		for (uint32_t k = from, ii = i0, jj = j0; k < to; ++k) {
			if (path[k] == 'M') {
				t += toupper((unsigned char)A[ii]) == toupper((unsigned char)B[jj]) ? '|' : ' ';
				++ii; ++jj;
			} else {
				if (path[k] == 'D') ++ii; else ++jj;
				t += ' ';
			}
		}
		t += '\n';
		fmts(t, "%5u ", j + 1);
		for (uint32_t k = from; k < to; ++k)
			t += (path[k] == 'D') ? '-' : B[j++];
		fmts(t, " %u  %s\n", j, lb);
		t += '\n';
	}
	fmts(t, "%s %u-%u length %u\n", la, lo_a + 1, pa, LA);
	fmts(t, "%s %u-%u length %u\n", lb, lo_b + 1, pb, LB);
	const double pct_gaps = 100.0 * ((double)gaps / (double)ncol), pct_ids = 100.0 * ((double)ids / (double)ncol);
	fmts(t, "AQ %.4f, cols %u, gaps %u (%.1f%%), ids %u (%.1f%%)", h.qual, ncol, gaps, pct_gaps, ids, pct_ids);
	if (h.pvalue != FLT_MAX)
		fmts(t, ", P-value %.3g", h.pvalue);
	t += '\n';
	return hand_over(t, out, cap);
}

// DSSAligner::ToFasta2 (dssaligner.cpp:981-1014): the two gapped rows of a hit as FASTA, TARGET first (the reference flips
// Up on entry), its label extended with E-value, identity and the query's label.  global: -unaligned.
extern "C" long long rsk_format_fasta2(const rsk_hit_view *v, int up_, int global, char *out, size_t cap)
{
	if (!v || !v->hit || !v->seq_a || !v->seq_b || (v->hit->path_len && !v->path) || (cap && !out))
		return RSK_ERR_ARG;
	const bool up = !(up_ != 0);  // sic
	std::string first, second;
	aligned_row(v, up ? 0 : 1, global != 0, first);
	aligned_row(v, up ? 1 : 0, global != 0, second);
	const char *l1 = up ? v->label_a : v->label_b, *l2 = up ? v->label_b : v->label_a;
	if (!l1) l1 = "";
	if (!l2) l2 = "";
	std::string label = l1;
	fmts(label, " E=%.3g Id=%.1f%%", v->hit->evalue, pct_id(v));
	label += " (";
	label += l2;
	label += ")";
	std::string t;
	seq_to_fasta(t, label, first);
	seq_to_fasta(t, l2, second);
	t += '\n';
	return hand_over(t, out, cap);
}
