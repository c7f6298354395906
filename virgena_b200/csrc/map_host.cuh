#pragma once
#include "ctx.cuh"
// temporary stubs
extern "C" {
int vg_set_reference(vg_ctx* c, const uint8_t*, int32_t, const int32_t*, int32_t, const int32_t*, int32_t, int32_t, const uint8_t*, const int64_t*, const int32_t*, int32_t, int32_t) { return vg_fail(c, VG_ERR_STATE, "not implemented"); }
int vg_set_msa_index(vg_ctx* c, int32_t, float, float, const int32_t*, int32_t, const uint8_t*, const int64_t*, const int32_t*, int32_t, int32_t) { return vg_fail(c, VG_ERR_STATE, "not implemented"); }
int vg_upload_reads(vg_ctx* c, const uint8_t*, int64_t, const int64_t*, const int32_t*, const int64_t*, const int32_t*, int64_t) { return vg_fail(c, VG_ERR_STATE, "not implemented"); }
int vg_map_pairs(vg_ctx* c, const int64_t*, int64_t, vg_map_stats*) { return vg_fail(c, VG_ERR_STATE, "not implemented"); }
int vg_map_pairs_msa(vg_ctx* c, vg_map_stats*) { return vg_fail(c, VG_ERR_STATE, "not implemented"); }
int vg_result_sizes(vg_ctx* c, int64_t*, int64_t*) { return vg_fail(c, VG_ERR_STATE, "not implemented"); }
int vg_get_results(vg_ctx* c, vg_pair_result*, int64_t, uint8_t*, int64_t) { return vg_fail(c, VG_ERR_STATE, "not implemented"); }
int vg_pileup(vg_ctx* c, int32_t) { return vg_fail(c, VG_ERR_STATE, "not implemented"); }
int vg_pileup_counts_device(vg_ctx* c, void**, int64_t*) { return vg_fail(c, VG_ERR_STATE, "not implemented"); }
int vg_pileup_counts_host(vg_ctx* c, int32_t*, int64_t) { return vg_fail(c, VG_ERR_STATE, "not implemented"); }
int vg_consensus(vg_ctx* c, int32_t, int32_t, const uint8_t*, int32_t, uint8_t*) { return vg_fail(c, VG_ERR_STATE, "not implemented"); }
int vg_seed_batch(vg_ctx* c, int32_t, const uint8_t*, const int64_t*, int64_t, int32_t, int32_t*, int32_t*) { return vg_fail(c, VG_ERR_STATE, "not implemented"); }
}
