// sink_kernel.cu - device-side hit sink: the records a search would emit are compacted ON THE DEVICE, batch after batch, in
// schedule order, together with their paths.  Replaces, for DB-sharded and hits-only searches, the D2H of one 64-byte record per
// scheduled pair: what leaves the GPU (over NVLink to the root rank, or over PCIe to the host) is exactly the hit set.
//
// Reference: the emit test of DBSearcher (runquery.cpp:72-73 `if (!DA.m_Path.empty()) BaseOnAln`, dbsearcher.cpp:258-265 Reject:
// E > MaxEvalue).  E = P(TS)*8340 needs libm's double pow (statsig.cpp:27-50), which stays on the host; the device applies the
// equivalent monotone test TS >= ts_lo with ts_lo a few ulps BELOW the exact threshold, and the root re-applies the exact test.
#include <cub/cub.cuh>

#include "rsk_internal.cuh"

namespace rsk {

namespace {

constexpr int kSinkThreads = 256;

__device__ __forceinline__ bool sink_keep(const PairRec &r, const SinkArgs &a)
{
	if (a.keep_all)
		return true;
	if (r.path_len == 0)
		return false;
	if (r.flags & RSK_HIT_HAS_EVALUE)
		return r.ts >= a.ts_lo;
	return a.report_no_evalue != 0;  // ClearAlign E = FLT_MAX: reported only when MaxEvalue >= FLT_MAX (-verysensitive)
}

// pass 1: keep flag and path bytes per scheduled pair, plus the work counters the host used to take from the records
__global__ void __launch_bounds__(kSinkThreads) sink_flag_kernel(SinkArgs a)
{
	const uint32_t k = blockIdx.x * kSinkThreads + threadIdx.x;
	uint32_t keep = 0, plen = 0, ne = 0, nr = 0;
	if (k < a.npairs) {
		const PairRec r = a.rec[k];
		keep = sink_keep(r, a) ? 1u : 0u;
		plen = (keep && a.want_paths) ? r.path_len : 0u;
		ne = (r.flags & RSK_HIT_HAS_EVALUE) ? 1u : 0u;
		nr = (r.flags & RSK_HIT_MU_REJECTED) ? 1u : 0u;
		a.keep[k] = keep;
		a.plen[k] = plen;
	}
	typedef cub::BlockReduce<uint32_t, kSinkThreads> BR;
	__shared__ typename BR::TempStorage tmp;
	const uint32_t se = BR(tmp).Sum(ne);
	__syncthreads();
	const uint32_t sr = BR(tmp).Sum(nr);
	if (threadIdx.x == 0) {
		if (se)
			atomicAdd(&a.totals[2], (unsigned long long)se);
		if (sr)
			atomicAdd(&a.totals[3], (unsigned long long)sr);
	}
}

// pass 2: one warp per 32 scheduled pairs.  Lane l writes the record of pair 32w+l (if kept) at its scanned position; then the
// warp copies the 32 paths one after the other, lanes striding over the bytes.
__global__ void __launch_bounds__(kSinkThreads) sink_scatter_kernel(SinkArgs a)
{
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t k = blockIdx.x * kSinkThreads + threadIdx.x;
	const unsigned long long base_rec = a.totals[0], base_path = a.totals[1];
	uint32_t keep = 0, plen = 0;
	unsigned long long src = 0, dst = 0;
	if (k < a.npairs) {
		keep = a.keep[k];
		if (keep) {
			PairRec r = a.rec[k];
			plen = a.plen[k];
			src = r.path_off;
			dst = base_path + a.plen_scan[k];
			SinkRec o;
			r.path_off = dst;
			if (!a.want_paths)
				r.path_off = 0;
			o.r = r;
			if (a.cross) {
				o.a = a.a_begin + k / a.nB + a.a_base;
				o.b = k % a.nB + a.b_base;
			} else {
				o.a = a.pair_a[k] + a.a_base;
				o.b = a.pair_b[k] + a.b_base;
			}
			a.out_rec[base_rec + a.keep_scan[k]] = o;
		}
	}
	if (!a.want_paths)
		return;
	for (int l = 0; l < 32; ++l) {
		const uint32_t n = __shfl_sync(0xffffffffu, plen, l);
		if (n == 0)
			continue;
		const unsigned long long s = __shfl_sync(0xffffffffu, src, l), d = __shfl_sync(0xffffffffu, dst, l);
		for (uint32_t i = lane; i < n; i += 32)
			a.out_pool[d + i] = a.pool[s + i];
	}
}

// pass 3: advance the running totals by this batch's counts
__global__ void sink_advance_kernel(SinkArgs a)
{
	if (a.npairs == 0)
		return;
	const uint32_t last = a.npairs - 1;
	a.totals[0] += (unsigned long long)a.keep_scan[last] + a.keep[last];
	a.totals[1] += a.plen_scan[last] + a.plen[last];
}

}  // namespace

size_t sink_scan_tmp_bytes(uint32_t npairs)
{
	size_t b1 = 0, b2 = 0;
	cub::DeviceScan::ExclusiveSum(nullptr, b1, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)npairs);
	cub::DeviceScan::ExclusiveSum(nullptr, b2, (const uint32_t *)nullptr, (unsigned long long *)nullptr, (int)npairs);
	return (b1 > b2 ? b1 : b2) + 256;
}

int launch_sink_append(const SinkArgs &a, void *tmp, size_t tmp_bytes, cudaStream_t st)
{
	if (a.npairs == 0)
		return 0;
	const int blocks = (int)((a.npairs + kSinkThreads - 1) / kSinkThreads);
	sink_flag_kernel<<<blocks, kSinkThreads, 0, st>>>(a);
	size_t tb = tmp_bytes;
	if (cub::DeviceScan::ExclusiveSum(tmp, tb, a.keep, a.keep_scan, (int)a.npairs, st) != cudaSuccess)
		return -1;
	tb = tmp_bytes;
	if (cub::DeviceScan::ExclusiveSum(tmp, tb, a.plen, a.plen_scan, (int)a.npairs, st) != cudaSuccess)
		return -1;
	sink_scatter_kernel<<<blocks, kSinkThreads, 0, st>>>(a);
	sink_advance_kernel<<<1, 1, 0, st>>>(a);
	if (cudaGetLastError() != cudaSuccess)
		return -1;
	return 7;  // flag, 2 x (scan = 2 kernels), scatter, advance
}

}  // namespace rsk
