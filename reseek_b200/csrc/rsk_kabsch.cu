// rsk_kabsch.cu - least-squares superposition of the aligned residue pairs of a hit (host code, double precision).
// Replaces DSSAligner::GetKabsch (dssaligner.cpp:1371-1385) -> Kabsch(ChainA, ChainB, LoA, LoB, Path, t, u)
// (kabsch.cpp:330-387), whose core (kabsch.cpp:21-327) is the TM-align transcription of Kabsch's 1978 routine.
//
// This is not that routine: the optimal rotation is found with Horn's unit-quaternion formulation (the largest eigenvector
// of a symmetric 4x4 matrix built from the 3x3 correlation matrix, by cyclic Jacobi rotations).  Both solve the same
// least-squares problem, so u, t and the residual agree to rounding (the parity test states the tolerance); the reference
// runs this on the host in double precision as well, once per reported pair of -alignpair.
#include <math.h>
#include <string.h>

#include "../../include/reseek_b200.h"

int rsk_fail(int code, const char *fmt, ...);

namespace {

// eigen-decomposition of a symmetric 4x4 matrix by cyclic Jacobi sweeps: A -> diagonal, V = eigenvectors (columns)
void jacobi4(double A[4][4], double V[4][4])
{
	for (int i = 0; i < 4; ++i)
		for (int j = 0; j < 4; ++j)
			V[i][j] = i == j ? 1.0 : 0.0;
	for (int sweep = 0; sweep < 64; ++sweep) {
		double off = 0, diag = 0;
		for (int i = 0; i < 4; ++i) {
			diag += A[i][i] * A[i][i];
			for (int j = i + 1; j < 4; ++j)
				off += A[i][j] * A[i][j];
		}
		if (off <= 1e-32 * (diag + off) || off == 0)
			break;
		for (int p = 0; p < 3; ++p)
			for (int q = p + 1; q < 4; ++q) {
				if (A[p][q] == 0)
					continue;
				const double theta = (A[q][q] - A[p][p]) / (2 * A[p][q]);
				const double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
				const double c = 1 / sqrt(tt * tt + 1), s = tt * c;
				for (int k = 0; k < 4; ++k) {  // columns p, q
					const double akp = A[k][p], akq = A[k][q];
					A[k][p] = c * akp - s * akq;
					A[k][q] = s * akp + c * akq;
				}
				for (int k = 0; k < 4; ++k) {  // rows p, q
					const double apk = A[p][k], aqk = A[q][k];
					A[p][k] = c * apk - s * aqk;
					A[q][k] = s * apk + c * aqk;
				}
				for (int k = 0; k < 4; ++k) {
					const double vkp = V[k][p], vkq = V[k][q];
					V[k][p] = c * vkp - s * vkq;
					V[k][q] = s * vkp + c * vkq;
				}
			}
	}
}

}  // namespace

extern "C" int rsk_kabsch(const float *xyz_a, uint32_t len_a, const float *xyz_b, uint32_t len_b, uint32_t lo_a, uint32_t lo_b,
		const char *path, uint32_t path_len, int up, double t[3], double u[9], double *msd)
{
	if (!xyz_a || !xyz_b || !path || !t || !u || path_len == 0)
		return rsk_fail(RSK_ERR_ARG, "rsk_kabsch: null argument or empty path");
	// up == 0: the chains change places (query = B) and D/I are exchanged (dssaligner.cpp:1380-1384)
	const float *X = up ? xyz_a : xyz_b, *Y = up ? xyz_b : xyz_a;
	const uint32_t LX = up ? len_a : len_b, LY = up ? len_b : len_a;
	uint32_t px = up ? lo_a : lo_b, py = up ? lo_b : lo_a;
	const char adv_x = up ? 'D' : 'I';
	// pass 1: centroids of the paired residues
	double cx[3] = {0, 0, 0}, cy[3] = {0, 0, 0};
	uint32_t M = 0;
	{
		uint32_t i = px, j = py;
		for (uint32_t c = 0; c < path_len; ++c) {
			const char ch = path[c];
			if (ch == 'M') {
				if (i >= LX || j >= LY)
					return rsk_fail(RSK_ERR_ARG, "rsk_kabsch: path runs past the end of a chain");
				for (int d = 0; d < 3; ++d) {
					cx[d] += (double)X[(size_t)d * LX + i];
					cy[d] += (double)Y[(size_t)d * LY + j];
				}
				++i; ++j; ++M;
			} else if (ch == adv_x) {
				++i;
			} else if (ch == 'D' || ch == 'I') {
				++j;
			} else {
				return rsk_fail(RSK_ERR_ARG, "rsk_kabsch: bad path letter '%c'", ch);
			}
		}
	}
	if (M == 0)
		return rsk_fail(RSK_ERR_ARG, "rsk_kabsch: no aligned pair in the path");  // the reference asserts n > 0
	for (int d = 0; d < 3; ++d) {
		cx[d] /= M;
		cy[d] /= M;
	}
	// pass 2: correlation matrix S[a][b] = sum x'_a y'_b and the two sums of squares
	double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, e0 = 0;
	{
		uint32_t i = px, j = py;
		for (uint32_t c = 0; c < path_len; ++c) {
			const char ch = path[c];
			if (ch == 'M') {
				double x[3], y[3];
				for (int d = 0; d < 3; ++d) {
					x[d] = (double)X[(size_t)d * LX + i] - cx[d];
					y[d] = (double)Y[(size_t)d * LY + j] - cy[d];
					e0 += x[d] * x[d] + y[d] * y[d];
				}
				for (int a = 0; a < 3; ++a)
					for (int b = 0; b < 3; ++b)
						S[a][b] += x[a] * y[b];
				++i; ++j;
			} else if (ch == adv_x) {
				++i;
			} else {
				++j;
			}
		}
	}
	double N[4][4] = {
		{S[0][0] + S[1][1] + S[2][2], S[1][2] - S[2][1], S[2][0] - S[0][2], S[0][1] - S[1][0]},
		{S[1][2] - S[2][1], S[0][0] - S[1][1] - S[2][2], S[0][1] + S[1][0], S[2][0] + S[0][2]},
		{S[2][0] - S[0][2], S[0][1] + S[1][0], -S[0][0] + S[1][1] - S[2][2], S[1][2] + S[2][1]},
		{S[0][1] - S[1][0], S[2][0] + S[0][2], S[1][2] + S[2][1], -S[0][0] - S[1][1] + S[2][2]}};
	double V[4][4];
	jacobi4(N, V);
	int best = 0;
	for (int k = 1; k < 4; ++k)
		if (N[k][k] > N[best][best])
			best = k;
	double q0 = V[0][best], q1 = V[1][best], q2 = V[2][best], q3 = V[3][best];
	const double qn = sqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
	q0 /= qn; q1 /= qn; q2 /= qn; q3 /= qn;
	double R[3][3] = {
		{q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, 2 * (q1 * q2 - q0 * q3), 2 * (q1 * q3 + q0 * q2)},
		{2 * (q2 * q1 + q0 * q3), q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3, 2 * (q2 * q3 - q0 * q1)},
		{2 * (q3 * q1 - q0 * q2), 2 * (q3 * q2 + q0 * q1), q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3}};
	for (int a = 0; a < 3; ++a) {
		for (int b = 0; b < 3; ++b)
			u[3 * a + b] = R[a][b];
		t[a] = cy[a] - (R[a][0] * cx[0] + R[a][1] * cx[1] + R[a][2] * cx[2]);  // y ~ u x + t
	}
	if (msd) {
		double rss = e0 - 2 * N[best][best];  // residual sum of squares at the optimum
		if (rss < 0)
			rss = 0;
		*msd = rss / M;  // kabsch.cpp:386 returns rms/M with rms = the residual SUM
	}
	return RSK_OK;
}
