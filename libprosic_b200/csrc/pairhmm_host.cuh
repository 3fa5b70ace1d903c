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
}  // namespace vlr
