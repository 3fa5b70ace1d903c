// pairhmm_host.cuh -- internal (non-ABI) entry of K1 used by the fused allele-support pipeline.
#pragma once
#include "common.cuh"

namespace vlr {
// Pair-HMM over explicit haplotype sub-windows: slot k aligns read pair_read[k] against
// hap_base[win_start[k] .. win_start[k] + win_len[k]) with band med[k]; only slots with
// mode[k] == 2 are computed (out[k] of the others is left untouched).  Device pointers, enqueued
// on the context's stream.
int pairhmm_windows_dev(vlr_ctx *ctx, int64_t n_slots, const uint8_t *d_hap_base,
                        const int64_t *d_win_start, const int32_t *d_win_len,
                        const uint8_t *d_read_bases, const uint8_t *d_read_quals,
                        const int64_t *d_read_off, const int32_t *d_pair_read, const int32_t *d_med,
                        const uint8_t *d_mode, const vlr_gap_params *gap, uint32_t flags,
                        void *d_workspace, size_t workspace_bytes, double *d_out);
// Literal log-space evaluation (policy LogLit of pairhmm_kernel.cuh: the reference's operation order
// with the library's deterministic libm) of the listed pairs; d_list == nullptr: all n_pairs_cap
// pairs; d_count: device count of list entries.  Context scratch holds the per-warp prob_cols (sized by
// max_lx) and, for read windows above 248 rows, the anti-diagonals of pairhmm_diag.cuh (sized by max_ly).
int pairhmm_literal_dev(vlr_ctx *ctx, int64_t n_pairs_cap, const uint8_t *d_hap_bytes, const int64_t *d_hap_off,
                        const int64_t *d_win_start, const int32_t *d_win_len, const uint8_t *d_read_bases,
                        const uint8_t *d_read_quals, const int64_t *d_read_off, const int32_t *d_pair_hap,
                        const int32_t *d_pair_read, const int32_t *d_med, const int32_t *d_list,
                        const int32_t *d_count, const vlr_gap_params *gap, uint32_t flags, int64_t max_lx,
                        int64_t max_ly, double *d_out);
}  // namespace vlr
