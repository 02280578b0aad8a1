// rsk_format.cu - host-side hit formatting, byte-compatible with the reference's writers.
// Replaces DSSAligner::ToTsv / WriteUserField (dssaligner.cpp:1016-1034, userfields.cpp:45-152), EvalueToStr
// (userfields.cpp:19-30), PathToCIGAR (cigar.cpp:95-139), GetQCovPct/GetTCovPct (dssaligner.cpp:1119-1141),
// GetPctId (dssaligner.cpp:1325-1369).  Pure host code: no device work happens here.
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
		else if (name == "pctid") {
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
			fmt(line, "%.1f", N == 0 ? 0.0f : (n * 100.0f) / N);
		}
		else if (name == "newts") fmt(line, "%.3g", h.ts);
		else if (name == "raw") fmt(line, "%.3g", h.score);
		else if (name == "dpscore") fmt(line, "%.4g", h.score);
		else if (name == "lddt") fmt(line, "%.4g", h.lddt);
		else if (name == "ids") fmtu(line, h.ids);
		else if (name == "gaps") fmtu(line, h.gaps);
		else if (name == "aq") fmt(line, "%.4f", h.qual);
		else if (name == "muhsp") { char b[32]; snprintf(b, sizeof(b), "%d", (h.flags & RSK_HIT_MKF) ? h.mu_fwd : 0); line += b; }
		else if (name == "muchain") { char b[32]; snprintf(b, sizeof(b), "%d", (h.flags & RSK_HIT_MKF) ? h.mu_rev : 0); line += b; }
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
