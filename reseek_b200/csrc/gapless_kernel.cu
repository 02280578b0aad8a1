// gapless_kernel.cu - K5: the alternative gapless Mu pre-scores, one warp per pair.
//
// Replaces SWFastGaplessProfb (swgaplessprofb.cpp:6-61: float ScoreMx_Mu, forward minus reversed-A) and
// SWFastPinopGapless (swfastpinopgapless.cpp:6-47: IntScoreMx_Mu, forward only).  Both reduce to: on every diagonal
// of the LA x LB matrix keep a running sum that restarts from 0 whenever it went negative, and take the maximum.
// The reference cannot reach these from its CLI (every caller that clears m_UsePara also sets omega = 0); they
// are exposed as a separate entry point with the same "Mu letters in, one score out" interface.
// Lanes own whole diagonals, so the float adds along a diagonal happen in the reference's order (bit-exact).
#include "rsk_internal.cuh"

namespace rsk {

namespace {

constexpr unsigned kFull = 0xffffffffu;

__global__ void __launch_bounds__(128) mu_gapless_kernel(const GaplessArgs a)
{
	__shared__ float s_f[36 * 36];
	__shared__ int s_i[36 * 36];
	for (int k = threadIdx.x; k < 36 * 36; k += blockDim.x) {
		s_f[k] = a.mu_f32[k];
		s_i[k] = a.mu_i32[k];
	}
	__syncthreads();
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t pair = blockIdx.x * 4 + warp;
	if (pair >= a.npairs)
		return;
	const uint32_t ia = a.pair_a[pair], ib = a.pair_b[pair];
	const uint8_t *A = a.muA + a.offA[ia], *B = a.muB + a.offB[ib];
	const int LA = (int)a.lenA[ia], LB = (int)a.lenB[ib];
	float bestf = 0.0f, bestr = 0.0f;
	int besti = 0;
	// diagonal d = j - i + (LA - 1), d in [0, LA + LB - 1)
	for (int d = lane; d < LA + LB - 1; d += 32) {
		int i = max(0, LA - 1 - d), j = max(0, d - (LA - 1));
		float xf = 0.0f, xr = 0.0f;
		int xi = 0;
		for (; i < LA && j < LB; ++i, ++j) {
			const int bj = B[j];
			if (xf < 0.0f) xf = 0.0f;
			if (xr < 0.0f) xr = 0.0f;
			if (xi < 0) xi = 0;
			xf += s_f[36 * A[i] + bj];
			xr += s_f[36 * A[LA - 1 - i] + bj];
			xi += s_i[36 * A[i] + bj];
			bestf = fmaxf(bestf, xf);
			bestr = fmaxf(bestr, xr);
			besti = max(besti, xi);
		}
	}
#pragma unroll
	for (int o = 16; o >= 1; o >>= 1) {
		bestf = fmaxf(bestf, __shfl_xor_sync(kFull, bestf, o));
		bestr = fmaxf(bestr, __shfl_xor_sync(kFull, bestr, o));
		besti = max(besti, __shfl_xor_sync(kFull, besti, o));
	}
	if (lane == 0) {
		a.out_f[pair] = bestf - bestr;
		a.out_i[pair] = besti;
	}
}

}  // namespace

int launch_mu_gapless(const GaplessArgs &args, cudaStream_t stream)
{
	if (args.npairs == 0)
		return 0;
	mu_gapless_kernel<<<(args.npairs + 3) / 4, 128, 0, stream>>>(args);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace rsk
