// dss_kernel.cu - K10: DSS feature extraction on the GPU: C-alpha coordinates + amino-acid sequence -> the 8
// feature planes the aligner consumes (AA, NENDist, Conf, NENConf, RENDist, DstNxtHlx, StrandDens, NormDens;
// namedparams.cpp:36-43) and the Mu letters (SS3 + 3*NENSS3 + 9*RENDist4; dss.cpp:629-644, dssparams.cpp:7-14).
//
// Replaces DSS::GetProfile / GetMuLetters (dss.cpp:716-741, 700-714) and what they call: GetSS (getss.cpp:6-63), the
// conformation letters (myss.cpp:142-210), CalcNEN / CalcREN (dss.cpp:417-440, 374-415), GetDensity + scaling (dss.cpp:179-244),
// GetSSDensity (dss.cpp:339-372), SetSSEs + DstNxtHlx (dss.cpp:78-155, 866-881) and the trained bins (valuetoint.cpp).
//
// One CTA per chain, one thread per residue (strided), four phases separated by block barriers.  One flipped letter changes
// alignments, so every value is computed in the reference's type and order: float coordinate differences, float sum of squares
// left to right, IEEE sqrtf, comparisons and density sums in double, first-minimum argmins (strict <), and exp() with the bits
// of the host's libm (exp_glibc.cuh).  The library is compiled with -fmad=false, so no multiply-add is contracted.
// Bound: FP64 pipe (about 100 exp evaluations and 200 distances per residue); algorithmic bytes per residue: 13 in (xyz + aa), 9 out.
#include <float.h>

#include "exp_glibc.cuh"
#include "rsk_internal.cuh"

namespace rsk {

namespace {

constexpr int kDssThreads = 128;
constexpr int kDensW = 50, kDensw = 3, kSSDensw = 8, kNenW = 100, kNenw = 12;  // dss.h:23-37
constexpr uint32_t kNoPos = 0xffffffffu;

struct ChainView {
	const float *x, *y, *z;
	uint32_t L;
	bool rev;
	__device__ __forceinline__ uint32_t ix(uint32_t p) const { return rev ? L - 1 - p : p; }
	// pdbchain.cpp:310-318 / abcxyz.h:116-126: float differences, float sum of squares left to right, float sqrt
	__device__ __forceinline__ float dist(uint32_t p1, uint32_t p2) const
	{
		const uint32_t a = ix(p1), b = ix(p2);
		const float dx = x[a] - x[b];
		const float dy = y[a] - y[b];
		const float dz = z[a] - z[b];
		const float d2 = dx * dx + dy * dy + dz * dz;
		return sqrtf(d2);
	}
};

__device__ __forceinline__ uint32_t bin15(const double *ts, double v)  // valuetoint.cpp: first threshold the value is below
{
	for (uint32_t i = 0; i < 15; ++i)
		if (v < ts[i])
			return i;
	return 15;
}

__device__ __forceinline__ uint32_t ss3(uint8_t c)  // dss.cpp:64-76
{
	return c == 'h' ? 0u : c == 's' ? 1u : 2u;
}

// getss.cpp:6-32
__device__ __forceinline__ uint8_t ss_char(const ChainView &C, uint32_t pos)
{
	if (pos < 2 || pos + 2 >= C.L)
		return '~';
	const double d13 = C.dist(pos - 2, pos), d14 = C.dist(pos - 2, pos + 1), d15 = C.dist(pos - 2, pos + 2);
	const double d24 = C.dist(pos - 1, pos + 1), d25 = C.dist(pos - 1, pos + 2), d35 = C.dist(pos, pos + 2);
	const double DH = 2.1;
	if (fabs(d15 - 6.37) < DH && fabs(d14 - 5.18) < DH && fabs(d25 - 5.18) < DH && fabs(d13 - 5.45) < DH && fabs(d24 - 5.45) < DH &&
		fabs(d35 - 5.45) < DH)
		return 'h';
	const double DS = 1.42;
	if (fabs(d15 - 13) < DS && fabs(d14 - 10.4) < DS && fabs(d25 - 10.4) < DS && fabs(d13 - 6.1) < DS && fabs(d24 - 6.1) < DS &&
		fabs(d35 - 6.1) < DS)
		return 's';
	if (d15 < 8.2)
		return 't';
	return '~';
}

// myss.cpp:142-183: nine distances around pos -> nearest of the 16 trained centres, first minimum
__device__ __forceinline__ uint32_t conf_letter(const ChainView &C, uint32_t pos, const double (*means)[9])
{
	if (pos < 3 || pos + 3 >= C.L)
		return 0;
	const int is[9] = {-2, -2, -2, -1, -1, 0, -3, 0, -3};
	const int js[9] = {0, 1, 2, 1, 2, 2, 3, 3, 0};
	double v[9];
#pragma unroll
	for (int m = 0; m < 9; ++m)
		v[m] = C.dist((uint32_t)((int)pos + is[m]), (uint32_t)((int)pos + js[m]));
	double mind = DBL_MAX;
	uint32_t best = 0;
	for (uint32_t k = 0; k < 16; ++k) {
		double sum2 = 0;
#pragma unroll
		for (int m = 0; m < 9; ++m) {
			const double diff = v[m] - means[k][m];
			sum2 += diff * diff;
		}
		const double d = sqrt(sum2);
		if (k == 0 || d < mind) {
			best = k;
			mind = d;
		}
	}
	return best;
}

// closest residue in [lo, hi] that is more than kNenw positions away; first minimum, start value 999 (dss.cpp:417-440)
__device__ __forceinline__ uint32_t nearest(const ChainView &C, uint32_t pos, int lo, int hi)
{
	double mind = 999;
	uint32_t minpos = kNoPos;
	for (int p2 = lo; p2 <= hi; ++p2) {
		if (p2 + kNenw >= (int)pos && p2 <= (int)pos + kNenw)
			continue;
		const double d = C.dist(pos, (uint32_t)p2);
		if (d < mind) {
			mind = d;
			minpos = (uint32_t)p2;
		}
	}
	return minpos;
}

__global__ void __launch_bounds__(kDssThreads) dss_kernel(const DssArgs a)
{
	__shared__ double s_red[2][kDssThreads / 32];
	__shared__ double s_min, s_max;
	__shared__ uint32_t s_nhelix;
	const DssTables &T = *a.tab;
	for (uint32_t c = blockIdx.x; c < a.n; c += gridDim.x) {
		const uint64_t off = a.off[c];
		ChainView C;
		C.x = a.x + off; C.y = a.y + off; C.z = a.z + off;
		C.L = a.len[c];
		C.rev = a.reverse != 0;
		const uint32_t L = C.L;
		uint8_t *ss = a.ss + off, *conf = a.conf + off;
		double *dens = a.dens + off;
		uint32_t *helix = a.helix_mid + off;
		uint8_t *pl = a.planes + off;
		const uint64_t ps = a.total;  // plane stride

		// ---- phase 1: secondary structure and conformation letters ----
		for (uint32_t pos = threadIdx.x; pos < L; pos += kDssThreads) {
			ss[pos] = ss_char(C, pos);
			conf[pos] = (uint8_t)conf_letter(C, pos, T.conf);
		}
		__syncthreads();

		// ---- phase 2: helix midpoints (one thread, dss.cpp:78-155) next to neighbours and densities ----
		if (threadIdx.x == 0) {
			uint32_t nh = 0;
			uint8_t currc = ss[0];
			uint32_t start = 0, run = 1;
			for (uint32_t pos = 1; pos <= L; ++pos) {
				const uint8_t s = pos == L ? 0 : ss[pos];  // the reference reads the string's terminating NUL
				if (s == currc) {
					++run;
				} else {
					if (run >= 8 && currc == 'h')  // m_SSE_MinLength; only helices are looked up afterwards (dss.cpp:866-881)
						helix[nh++] = start + run / 2;
					currc = s;
					start = pos;
					run = 1;
				}
			}
			s_nhelix = nh;
		}
		double lmin = 999, lmax = 0;
		for (uint32_t pos = threadIdx.x; pos < L; pos += kDssThreads) {
			// NEN / REN (dss.cpp:417-440, 374-415)
			const int lo = max(0, (int)pos - kNenW), hi = min((int)L - 1, (int)pos + kNenW);
			const uint32_t nen = nearest(C, pos, lo, hi);
			uint32_t ren = kNoPos;
			if (nen != kNoPos) {
				if (nen > pos) {
					if ((int)pos - 1 >= 0)
						ren = nearest(C, pos, lo, (int)pos - 1);
				} else {
					ren = nearest(C, pos, (int)pos + 1, hi);
				}
			}
			const double nend = nen == kNoPos ? 10.0 : (double)C.dist(pos, nen);  // m_DefaultNENDist
			const double rend = ren == kNoPos ? 10.0 : (double)C.dist(pos, ren);
			const uint32_t l_nend = bin15(T.bins[0], nend), l_rend = bin15(T.bins[1], rend);
			const uint32_t l_conf = conf[pos];
			const uint32_t l_nenconf = nen == kNoPos ? 0u : conf[nen];
			// densities (dss.cpp:217-244, 339-372): one exp per neighbour feeds both sums, each in ascending neighbour order
			double D = DBL_MAX, sdens = DBL_MAX;
			if (!(pos == 0 || pos + 1 >= L)) {
				const int dlo = max(0, (int)pos - kDensW), dhi = min((int)L - 1, (int)pos + kDensW);
				double d3 = 0, d8 = 0, dc8 = 0;
				for (int p2 = dlo; p2 <= dhi; ++p2) {
					if (p2 + kDensw >= (int)pos && p2 <= (int)pos + kDensw)
						continue;
					const double dist = C.dist(pos, (uint32_t)p2);
					const double e = rsk_exp_glibc(-dist / 20.0);  // m_Density_Radius
					d3 += e;
					if (!(p2 + kSSDensw >= (int)pos && p2 <= (int)pos + kSSDensw)) {
						d8 += e;
						if (ss[p2] == 's')
							dc8 += e;
					}
				}
				D = d3;
				sdens = dc8 / (d8 + 1.0);  // m_SSDensity_epsilon
				lmin = D < lmin ? D : lmin;
				lmax = lmax < D ? D : lmax;
			}
			dens[pos] = D;
			const uint32_t l_strand = bin15(T.bins[3], sdens);
			// amino-acid letter
			uint32_t l_aa;
			if (a.aa_char) {
				const uint32_t t = T.amino[a.aa_char[off + C.ix(pos)]];
				l_aa = t >= 20 ? 0u : t;
			} else {
				l_aa = (uint32_t)(a.aa_prof8[off + C.ix(pos)] & 0xffu);
			}
			pl[0 * ps + pos] = (uint8_t)l_aa;
			pl[1 * ps + pos] = (uint8_t)l_nend;
			pl[2 * ps + pos] = (uint8_t)l_conf;
			pl[3 * ps + pos] = (uint8_t)l_nenconf;
			pl[4 * ps + pos] = (uint8_t)l_rend;
			pl[6 * ps + pos] = (uint8_t)l_strand;
			if (a.mu)
				a.mu[off + pos] = (uint8_t)(ss3(ss[pos]) + 3u * (nen == kNoPos ? 0u : ss3(ss[nen])) + 9u * (l_rend / 4u));
		}
		// min / max of the defined densities (dss.cpp:179-215)
#pragma unroll
		for (int o = 16; o >= 1; o >>= 1) {
			lmin = fmin(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
			lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
		}
		if ((threadIdx.x & 31) == 0) {
			s_red[0][threadIdx.x >> 5] = lmin;
			s_red[1][threadIdx.x >> 5] = lmax;
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			double mn = s_red[0][0], mx = s_red[1][0];
			for (int w = 1; w < kDssThreads / 32; ++w) {
				mn = fmin(mn, s_red[0][w]);
				mx = fmax(mx, s_red[1][w]);
			}
			s_min = mn;
			s_max = mx;
		}
		__syncthreads();

		// ---- phase 3: scaled density and distance to the next helix ----
		const double mn = s_min;
		double range = s_max - mn;
		if (range < 1)
			range = 1;
		const uint32_t nh = s_nhelix;
		for (uint32_t pos = threadIdx.x; pos < L; pos += kDssThreads) {
			const double v = dens[pos];
			pl[7 * ps + pos] = (uint8_t)bin15(T.bins[4], v == DBL_MAX ? DBL_MAX : (v - mn) / range);
			double dnh = 0;
			for (uint32_t k = 0; k < nh; ++k) {
				const uint32_t mid = helix[k];
				if (mid <= pos + 8)  // m_SSE_Margin
					continue;
				dnh = C.dist(pos, mid);
				break;
			}
			pl[5 * ps + pos] = (uint8_t)bin15(T.bins[2], dnh);
		}
		__syncthreads();  // scratch and shared values are reused by the CTA's next chain
	}
}

// plane bytes back out of the packed e-letters (download for hosts that want the letters)
__global__ void dss_unpack_kernel(const uint64_t *__restrict__ prof8, uint64_t total, uint8_t *planes)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= total)
		return;
	const uint64_t e = prof8[i];
#pragma unroll
	for (int f = 0; f < RSK_NFEAT; ++f)
		planes[(uint64_t)f * total + i] = (uint8_t)(((e >> (8 * f)) & 0xff) - feat_base(f));
}

}  // namespace

int launch_dss(const DssArgs &a, int grid, cudaStream_t st)
{
	if (a.n == 0)
		return 0;
	dss_kernel<<<grid, kDssThreads, 0, st>>>(a);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_dss_unpack(const uint64_t *prof8, uint64_t total, uint8_t *planes, cudaStream_t st)
{
	if (total == 0)
		return 0;
	dss_unpack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(prof8, total, planes);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace rsk
