// rsk_dss.cu - host side of the device DSS (dss_kernel.cu): chain sets made from coordinates, reversed sets for the
// self-reverse scores, and the download of the feature letters for hosts that want them.
#include <float.h>

#include "rsk_host.cuh"
#include "host/dss_tables_data.inc"

// slab helpers of rsk_api.cu
void *rsk_slab_get(rsk_ctx *ctx, size_t bytes, size_t &cap);

namespace {

int dss_tables(rsk_ctx *ctx)
{
	if (ctx->d_dss_tables)
		return RSK_OK;
	static DssTables T;
	static bool init = false;
	if (!init) {
		memcpy(T.conf, rsk_dss_conf_means, sizeof(T.conf));
		memcpy(T.bins[0], rsk_dss_bins_NENDist, sizeof(T.bins[0]));
		memcpy(T.bins[1], rsk_dss_bins_RENDist, sizeof(T.bins[1]));
		memcpy(T.bins[2], rsk_dss_bins_DstNxtHlx, sizeof(T.bins[2]));
		memcpy(T.bins[3], rsk_dss_bins_StrandDens, sizeof(T.bins[3]));
		memcpy(T.bins[4], rsk_dss_bins_NormDens, sizeof(T.bins[4]));
		memcpy(T.amino, rsk_dss_amino_letter, sizeof(T.amino));
		init = true;
	}
	CK(cudaMalloc((void **)&ctx->d_dss_tables, sizeof(DssTables)));
	CK(cudaMemcpy(ctx->d_dss_tables, &T, sizeof(DssTables), cudaMemcpyHostToDevice));
	return RSK_OK;
}

// lay a chain set out in one slab (same layout as rsk_chainset_upload) and fill len/off; returns the set with device pointers
int make_set(rsk_ctx *ctx, uint32_t n, const uint32_t *len, bool with_mu, rsk_chainset **out)
{
	rsk_chainset *cs = new rsk_chainset();
	cs->ctx = ctx;
	cs->device = ctx->device;
	cs->hlen.assign(len, len + n);
	cs->hoff.resize(n);
	uint64_t tot = 0;
	for (uint32_t i = 0; i < n; ++i) {
		if (len[i] == 0) {
			delete cs;
			return fail(RSK_ERR_ARG, "chain %u has length 0 (the reference's reader skips such chains, chainreader2.cpp:103-107)", i);
		}
		cs->hoff[i] = tot;
		tot += len[i];
		cs->maxlen = std::max(cs->maxlen, len[i]);
	}
	DevChains &d = cs->d;
	d.n = n;
	d.total = tot;
	cs->has_mu = with_mu;
	auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
	size_t o_len = 0, o_off = o_len + up(sizeof(uint32_t) * n), o_prof = o_off + up(sizeof(uint64_t) * n),
		   o_x = o_prof + up(sizeof(uint64_t) * tot), o_y = o_x + up(sizeof(float) * tot), o_z = o_y + up(sizeof(float) * tot),
		   o_sr = o_z + up(sizeof(float) * tot), o_mu = o_sr + up(sizeof(float) * n), o_end = o_mu + (with_mu ? up(tot) : 0);
	cs->slab = rsk_slab_get(ctx, o_end, cs->slab_bytes);
	if (!cs->slab) {
		cudaGetLastError();
		rsk_chainset_free(cs);
		return fail(RSK_ERR_NOMEM, "device memory for %llu residues", (unsigned long long)tot);
	}
	unsigned char *base = (unsigned char *)cs->slab;
	d.len = (uint32_t *)(base + o_len); d.off = (uint64_t *)(base + o_off); d.prof8 = (uint64_t *)(base + o_prof);
	d.x = (float *)(base + o_x); d.y = (float *)(base + o_y); d.z = (float *)(base + o_z); d.selfrev = (float *)(base + o_sr);
	d.mu = with_mu ? (uint8_t *)(base + o_mu) : nullptr;
	cudaStream_t st = ctx->stream;
	cudaError_t e = cudaMemcpyAsync(d.len, cs->hlen.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st);
	if (e == cudaSuccess)
		e = cudaMemcpyAsync(d.off, cs->hoff.data(), sizeof(uint64_t) * n, cudaMemcpyHostToDevice, st);
	if (e != cudaSuccess) {
		rsk_chainset_free(cs);
		return fail(RSK_ERR_CUDA, "chain set upload: %s", cudaGetErrorString(e));
	}
	*out = cs;
	return RSK_OK;
}

// run K10 for the chains of cs (coordinates already in place); AA letters from characters or from another set's profile
int run_dss(rsk_ctx *ctx, rsk_chainset *cs, const float *x, const float *y, const float *z, uint8_t *mu_out,
		const uint8_t *d_aa_char, const uint64_t *d_aa_prof8, bool reverse)
{
	int rc = dss_tables(ctx);
	if (rc)
		return rc;
	const uint64_t tot = cs->d.total;
	if (ctx->dss_ss.ensure(tot) || ctx->dss_conf.ensure(tot) || ctx->dss_dens.ensure(tot) || ctx->dss_helix.ensure(tot) ||
		ctx->upload_stage.ensure((size_t)RSK_NFEAT * tot)) {
		cudaGetLastError();
		return fail(RSK_ERR_NOMEM, "DSS scratch for %llu residues", (unsigned long long)tot);
	}
	DssArgs a;
	memset(&a, 0, sizeof(a));
	a.n = cs->d.n; a.total = tot;
	a.len = cs->d.len; a.off = cs->d.off; a.x = x; a.y = y; a.z = z;
	a.aa_char = d_aa_char; a.aa_prof8 = d_aa_prof8;
	a.reverse = reverse ? 1 : 0;
	a.ss = ctx->dss_ss.p; a.conf = ctx->dss_conf.p; a.dens = ctx->dss_dens.p; a.helix_mid = ctx->dss_helix.p;
	a.planes = ctx->upload_stage.p;
	a.mu = mu_out;
	a.tab = ctx->d_dss_tables;
	const int grid = (int)std::min<uint64_t>(cs->d.n, (uint64_t)ctx->num_sms * 16);
	int nl = launch_dss(a, grid, ctx->stream);
	if (nl < 0)
		return fail(RSK_ERR_CUDA, "DSS kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
	int nl2 = launch_pack_profiles(ctx->upload_stage.p, tot, cs->d.prof8, ctx->stream);
	if (nl2 < 0)
		return fail(RSK_ERR_CUDA, "profile packing failed: %s", cudaGetErrorString(cudaGetLastError()));
	ctx->stats.kernel_launches += nl + nl2;
	return RSK_OK;
}

}  // namespace

extern "C" int rsk_chainset_from_coords(rsk_ctx *ctx, const rsk_coords_host *h, int with_mu, rsk_chainset **out)
{
	if (!ctx || !h || !out)
		return fail(RSK_ERR_ARG, "rsk_chainset_from_coords: null argument");
	*out = nullptr;
	if (h->n == 0 || !h->len || !h->aa || !h->xyz)
		return fail(RSK_ERR_ARG, "rsk_chainset_from_coords: empty chain set or missing len/aa/xyz");
	CK(cudaSetDevice(ctx->device));
	rsk_chainset *cs = nullptr;
	int rc = make_set(ctx, h->n, h->len, with_mu != 0, &cs);
	if (rc)
		return rc;
	const uint64_t tot = cs->d.total;
	if (tot != h->total) {
		rsk_chainset_free(cs);
		return fail(RSK_ERR_ARG, "rsk_chainset_from_coords: total=%llu but sum(len)=%llu", (unsigned long long)h->total, (unsigned long long)tot);
	}
	if (ctx->dss_aa.ensure(tot)) {
		cudaGetLastError();
		rsk_chainset_free(cs);
		return fail(RSK_ERR_NOMEM, "rsk_chainset_from_coords: device memory");
	}
	cudaStream_t st = ctx->stream;
	cudaError_t e = cudaSuccess;
	if (rsk_h2d(ctx, cs->d.x, h->xyz, sizeof(float) * tot, nullptr) || rsk_h2d(ctx, cs->d.y, h->xyz + tot, sizeof(float) * tot, nullptr) ||
		rsk_h2d(ctx, cs->d.z, h->xyz + 2 * tot, sizeof(float) * tot, nullptr) || rsk_h2d(ctx, ctx->dss_aa.p, h->aa, tot, nullptr))
		e = cudaErrorUnknown;
	std::vector<float> sr(cs->d.n, FLT_MAX);  // "unset" (dssaligner.cpp:876-877) until rsk_chainset_selfrev fills them
	if (e == cudaSuccess) e = cudaMemcpyAsync(cs->d.selfrev, sr.data(), sizeof(float) * cs->d.n, cudaMemcpyHostToDevice, st);
	if (e != cudaSuccess) {
		rsk_chainset_free(cs);
		return fail(RSK_ERR_CUDA, "rsk_chainset_from_coords: %s", cudaGetErrorString(e));
	}
	rc = run_dss(ctx, cs, cs->d.x, cs->d.y, cs->d.z, cs->d.mu, ctx->dss_aa.p, nullptr, false);
	if (!rc && cudaStreamSynchronize(st) != cudaSuccess)
		rc = fail(RSK_ERR_CUDA, "rsk_chainset_from_coords: %s", cudaGetErrorString(cudaGetLastError()));
	if (rc) {
		rsk_chainset_free(cs);
		return rc;
	}
	ctx->stats.h2d_bytes += tot * 13 + (uint64_t)cs->d.n * 16;
	*out = cs;
	return RSK_OK;
}

// PDBChain::GetReverse (pdbchain.cpp:478) + DSS of every reversed chain, with the FORWARD Mu letters (alignpair.cpp:22):
// the second argument of rsk_chainset_selfrev, made without leaving the device.
extern "C" int rsk_chainset_reversed(rsk_ctx *ctx, const rsk_chainset *S, rsk_chainset **out)
{
	if (!ctx || !S || !out)
		return fail(RSK_ERR_ARG, "rsk_chainset_reversed: null argument");
	*out = nullptr;
	if (S->ctx != ctx)
		return fail(RSK_ERR_ARG, "rsk_chainset_reversed: the chain set belongs to a different context");
	CK(cudaSetDevice(ctx->device));
	rsk_chainset *cs = nullptr;
	int rc = make_set(ctx, S->d.n, S->hlen.data(), S->has_mu, &cs);
	if (rc)
		return rc;
	cudaStream_t st = ctx->stream;
	const uint64_t tot = S->d.total;
	// the kernel reads S's coordinates back to front and takes the AA letters from S's packed profile; the Mu letters of the
	// new set are S's forward letters (copied below), not recomputed
	rc = run_dss(ctx, cs, S->d.x, S->d.y, S->d.z, nullptr, nullptr, S->d.prof8, true);
	cudaError_t e = cudaSuccess;
	if (!rc) {
		if (S->has_mu)
			e = cudaMemcpyAsync(cs->d.mu, S->d.mu, tot, cudaMemcpyDeviceToDevice, st);
		// coordinates: S's, unreversed - never read: rsk_chainset_selfrev stops after the SW score (alignpair.cpp:24)
		if (e == cudaSuccess) e = cudaMemcpyAsync(cs->d.x, S->d.x, sizeof(float) * tot, cudaMemcpyDeviceToDevice, st);
		if (e == cudaSuccess) e = cudaMemcpyAsync(cs->d.y, S->d.y, sizeof(float) * tot, cudaMemcpyDeviceToDevice, st);
		if (e == cudaSuccess) e = cudaMemcpyAsync(cs->d.z, S->d.z, sizeof(float) * tot, cudaMemcpyDeviceToDevice, st);
		std::vector<float> sr(cs->d.n, FLT_MAX);
		if (e == cudaSuccess) e = cudaMemcpyAsync(cs->d.selfrev, sr.data(), sizeof(float) * cs->d.n, cudaMemcpyHostToDevice, st);
		if (e == cudaSuccess) e = cudaStreamSynchronize(st);
		if (e != cudaSuccess)
			rc = fail(RSK_ERR_CUDA, "rsk_chainset_reversed: %s", cudaGetErrorString(e));
	}
	if (rc) {
		rsk_chainset_free(cs);
		return rc;
	}
	*out = cs;
	return RSK_OK;
}

extern "C" int rsk_chainset_download_features(rsk_ctx *ctx, const rsk_chainset *S, uint8_t *prof, uint8_t *mu, float *selfrev)
{
	if (!ctx || !S)
		return fail(RSK_ERR_ARG, "rsk_chainset_download_features: null argument");
	if (S->ctx != ctx)
		return fail(RSK_ERR_ARG, "rsk_chainset_download_features: the chain set belongs to a different context");
	CK(cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	const uint64_t tot = S->d.total;
	if (prof) {
		if (ctx->upload_stage.ensure((size_t)RSK_NFEAT * tot)) {
			cudaGetLastError();
			return fail(RSK_ERR_NOMEM, "rsk_chainset_download_features: device memory");
		}
		if (launch_dss_unpack(S->d.prof8, tot, ctx->upload_stage.p, st) < 0)
			return fail(RSK_ERR_CUDA, "unpack kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
		ctx->stats.kernel_launches += 1;
		CK(cudaMemcpyAsync(prof, ctx->upload_stage.p, (size_t)RSK_NFEAT * tot, cudaMemcpyDeviceToHost, st));
	}
	if (mu) {
		if (!S->has_mu)
			return fail(RSK_ERR_ARG, "rsk_chainset_download_features: the chain set has no Mu letters");
		CK(cudaMemcpyAsync(mu, S->d.mu, tot, cudaMemcpyDeviceToHost, st));
	}
	if (selfrev)
		CK(cudaMemcpyAsync(selfrev, S->d.selfrev, sizeof(float) * S->d.n, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	ctx->stats.d2h_bytes += (prof ? RSK_NFEAT * tot : 0) + (mu ? tot : 0) + (selfrev ? 4ull * S->d.n : 0);
	return RSK_OK;
}
