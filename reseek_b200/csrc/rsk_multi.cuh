// rsk_multi.cuh - internal interface between the search drivers (rsk_api.cu, rsk_prefilter.cu) and the communicator (rsk_multi.cu).
#pragma once

#include "rsk_host.cuh"

struct rsk_comm;

namespace rsk {

constexpr int kCommCountWords = 4;
// every rank contributes 4 words; all[r*4 + w] = rank r's word w (host memory owned by the communicator)
int comm_exchange_counts(rsk_comm *c, const unsigned long long *mine, const unsigned long long **all);
// exact-size gather of up to 3 device arrays per rank on `root` (out/out_total are set on the root only)
int comm_gather_parts(rsk_comm *c, int root, int nparts, const void *const *src, const unsigned long long *bytes,
		unsigned char **out, unsigned long long *out_total);
// exact-size all-gather of up to 3 device arrays per rank, rank order, result on every rank
int comm_allgather_parts(rsk_comm *c, int nparts, const void *const *src, const unsigned long long *bytes,
		unsigned char **out, unsigned long long *out_total);
// explicit pair list -> device hit sink -> gather on root (rsk_api.cu); a_base/b_base are added to the reported chain indices
int search_pairs_sharded(rsk_ctx *ctx, rsk_comm *comm, const rsk_chainset *A, const rsk_chainset *B, uint64_t npairs,
		const uint32_t *ia, const uint32_t *ib, uint32_t a_base, uint32_t b_base, const rsk_search_opts *opts, int root, rsk_results **out);
int comm_rank(const rsk_comm *c);
int comm_nranks(const rsk_comm *c);
rsk_ctx *comm_ctx(const rsk_comm *c);

}  // namespace rsk
