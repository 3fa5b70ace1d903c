// pairhmm_diag.cuh -- the pair-HMM forward recursion as an anti-diagonal sweep with the LANES on the
// CELLS of the anti-diagonal, in natural-log space, for the pairs no other kernel takes: banded pairs
// and explicit haplotype windows with read windows above 248 rows (bio's Myers::Long territory:
// merged multi-locus regions, breakend alleles; realignment/mod.rs:201-223, edit_distance.rs:42-46).
//
// The wavefront kernel (pairhmm_kernel.cuh) fixes a lane to a block of read rows and is limited to
// 31 x 8 rows; its strip-wise long-read variant has no band.  Here nothing is tied to a lane: the
// three anti-diagonals d, d-1, d-2 (M, X, Y and the band's minimal edit distance per read row) live
// in a per-warp scratch in global memory, cell (i, j) with i + j = d reads its three predecessors
// from d-1 / d-2 and the lanes take the cells of d 32 at a time.  Under a band only the cells that
// can be alive are visited: the candidate rows of d are the hull of the rows whose predecessors had
// a distance <= k on d-1 / d-2 (+ row 0 while a new column starts), everything else stays at the
// reset state exactly as in the reference (a cell is computed iff one of its three predecessors has
// a minimal edit distance <= k), and the sweep stops when both previous diagonals are dead.
//
// Policy Log: CUDA libm, |d ln P| ~ 1e-12 against the oracle.  Policy LogLit: the reference's literal
// operation order with the deterministic libm, prob_cols kept per column and summed in slice order:
// bit-identical to the oracle (the near-tie decisions of allele support).
#pragma once
#include "pairhmm_kernel.cuh"

namespace vlr {

// final LogProb::ln_sum_exp(prob_cols) of the literal kernels: first maximum, then the sum of
// exp(p - max) over the other finite entries in slice order, accumulated by ONE lane in that order
__device__ __forceinline__ double literal_ln_sum_exp_cols(double *cols, int nc, int lane) {
    double pmax = -INFINITY;
    for (int k = lane; k < nc; k += 32) pmax = fmax(pmax, cols[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pmax = fmax(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
    int imax = nc;
    for (int k = lane; k < nc; k += 32)
        if (cols[k] == pmax) { imax = k; break; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) imax = min(imax, __shfl_xor_sync(0xffffffffu, imax, o));
    double res = pmax;
    if (pmax != -INFINITY && pmax != INFINITY) {
        for (int k = lane; k < nc; k += 32) {
            const double v = cols[k];
            cols[k] = (k == imax || v == -INFINITY) ? 0.0 : dm::exp(dm::sub(v, pmax));
        }
        __syncwarp();
        if (lane == 0) {
            double sum = 0.0;
            for (int k = 0; k < nc; ++k) sum = dm::add(sum, cols[k]);  // + 0.0 is exact
            res = dm::add(pmax, dm::log1p(sum));
        }
    }
    return res;  // valid in lane 0
}

struct DiagArgs {
    uint8_t *ws;          // per-warp scratch, ws_bytes each
    int64_t ws_bytes;
    int32_t min_rows;     // pairs with fewer read rows are left to the other kernels (list entry skipped)
};

template <int APPROX, typename A>
__global__ void __launch_bounds__(kWarps * 32, 2) pairhmm_diag_kernel(HmmArgs g, DiagArgs da) {
    static_assert(A::kLog, "log-space policies only");
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const HmmConsts c = g.ln;
    const int n = *g.count;
    const int32_t *list = g.list ? g.list + (g.list_off ? *g.list_off : 0) : nullptr;
    const int64_t wid = (int64_t)blockIdx.x * kWarps + warp;
    uint8_t *wsb = da.ws + wid * da.ws_bytes;
    // capacity in read rows: three diagonals of (M, X, Y, med) = 3 * 28 bytes per row
    const int cap = (int)min((int64_t)0x3fffffff, (da.ws_bytes - 64) / 84);
    double *bufM = reinterpret_cast<double *>(wsb);
    double *bufX = bufM + 3 * (int64_t)cap;
    double *bufY = bufX + 3 * (int64_t)cap;
    int *bufD = reinterpret_cast<int *>(bufY + 3 * (int64_t)cap);
    double *cols = A::kLit ? g.lit_ws + wid * 3 * g.lit_stride : nullptr;
    const double kNegInf = -INFINITY;
    const int kBig = 0x3fffffff;

    for (;;) {
        int p = 0;
        if (lane == 0) p = atomicAdd(g.cursor, 1);
        p = __shfl_sync(0xffffffffu, p, 0);
        if (p >= n) break;
        const int pair = list ? list[p] : p;
        if (pair < 0) continue;
        const int64_t hidx = g.pair_hap ? (int64_t)g.pair_hap[pair] : (int64_t)pair;
        const int64_t rd = g.pair_read ? (int64_t)g.pair_read[pair] : (int64_t)pair;
        const int64_t ho = g.win_start ? g.win_start[pair] : g.hap_off[hidx];
        const int lx = g.win_start ? g.win_len[pair] : (int)(g.hap_off[hidx + 1] - ho);
        const int64_t ro = g.read_off[rd];
        const int ly = (int)(g.read_off[rd + 1] - ro);
        if (ly < da.min_rows) continue;
        int band = -1;
        if (g.med) band = g.med[pair];
        const bool banded = band >= 0;
        if (lx <= 0 || ly <= 0) {
            if (lane == 0) g.out[pair] = -INFINITY;
            continue;
        }
        if (ly > cap || (A::kLit && lx > g.lit_stride)) {  // the caller sized the scratch for shorter windows
            if (lane == 0) g.out[pair] = NAN;
            continue;
        }
        if (A::kLit)
            for (int k = lane; k < 3 * lx; k += 32) cols[k] = kNegInf;
        // written and live row ranges of the two previous diagonals (empty: lo > hi)
        int w1lo = kBig, w1hi = -1, w2lo = kBig, w2hi = -1;
        int l1lo = kBig, l1hi = -1, l2lo = kBig, l2hi = -1;
        double amx = kNegInf, asum = 0.0;  // Log: online ln-sum-exp of the last-row values this lane saw
        int b0 = 0;                        // buffer index of diagonal d (d mod 3)
        __syncwarp();
        const int nd = lx + ly - 1;
        for (int d = 0; d < nd; ++d) {
            const int vlo = max(0, d - lx + 1), vhi = min(ly - 1, d);
            int clo = vlo, chi = vhi;
            if (banded) {
                clo = d < lx ? 0 : min(l1lo, l2lo == kBig ? kBig : l2lo + 1);
                chi = max(l1hi, l2hi) + 1;
                if (d < lx) chi = max(chi, 0);
                clo = max(clo, vlo);
                chi = min(chi, vhi);
                if (d >= lx && l1hi < 0 && l2hi < 0) break;  // everything further is at the reset state
            }
            const int b1 = b0 == 0 ? 2 : b0 - 1, b2 = b1 == 0 ? 2 : b1 - 1;
            const double *M1 = bufM + (int64_t)b1 * cap, *X1 = bufX + (int64_t)b1 * cap, *Y1 = bufY + (int64_t)b1 * cap;
            const double *M2 = bufM + (int64_t)b2 * cap, *X2 = bufX + (int64_t)b2 * cap, *Y2 = bufY + (int64_t)b2 * cap;
            const int *D1 = bufD + (int64_t)b1 * cap, *D2 = bufD + (int64_t)b2 * cap;
            double *M0 = bufM + (int64_t)b0 * cap, *X0 = bufX + (int64_t)b0 * cap, *Y0 = bufY + (int64_t)b0 * cap;
            int *D0 = bufD + (int64_t)b0 * cap;
            int llo = kBig, lhi = -1;
            for (int j = clo + lane; j <= chi; j += 32) {
                const int i = d - j;
                // predecessors: diag = (i-1, j-1) on d-2, left = (i-1, j) on d-1, up = (i, j-1) on d-1
                double dgM = kNegInf, dgX = kNegInf, dgY = kNegInf, lfM = kNegInf, lfX = kNegInf, upM = kNegInf, upY = kNegInf;
                int dgD = kMedMax, lfD = kMedMax, upD = kMedMax;
                if (j == 0) {  // fm[prev][0] = ln 1, min_edit_dist[prev][0] = 0 at the start of every column
                    dgM = 0.0;
                    dgD = 0;
                } else {
                    if (i > 0 && j - 1 >= w2lo && j - 1 <= w2hi) {
                        dgM = __ldcg(M2 + j - 1); dgX = __ldcg(X2 + j - 1); dgY = __ldcg(Y2 + j - 1); dgD = __ldcg(D2 + j - 1);
                    }
                    if (j - 1 >= w1lo && j - 1 <= w1hi) {
                        upM = __ldcg(M1 + j - 1); upY = __ldcg(Y1 + j - 1); upD = __ldcg(D1 + j - 1);
                    }
                }
                if (i > 0 && j >= w1lo && j <= w1hi) {
                    lfM = __ldcg(M1 + j); lfX = __ldcg(X1 + j); lfD = __ldcg(D1 + j);
                }
                double nM = kNegInf, nX = kNegInf, nY = kNegInf;
                int nD = kMedMax;
                if (!banded || min(dgD, min(upD, lfD)) <= band) {
                    const int q = g.rqual[ro + j];
                    const int rbase = g.rbase[ro + j];
                    int xb = g.hap[ho + i];
                    if ((unsigned)(xb - 'a') < 26u) xb -= 32;  // to_ascii_uppercase (pairhmm.rs:190)
                    const bool match = xb == rbase;
                    const double ln_eps = A::kLit ? dm::mul((double)q, -0.23025850929940456) : (double)q * -0.23025850929940456;
                    const double e = match ? g.qlut[q * kQlutStride + 3] : A::mul(ln_eps, kLnConfusion);
                    unsigned trk_unused = 0;
                    double wj_unused = 0.0;
                    nM = A::mul(e, A::template sum3<APPROX, true, false>(c, dgM, dgX, dgY, trk_unused, wj_unused));
                    nX = A::template dot2<true>(c.t_gy, lfM, c.t_gye, lfX);
                    nY = A::mul(ln_eps, A::add(A::mul(c.t_gx, upM), A::mul(c.t_gxe, upY)));
                    if (banded) {
                        nD = min(min(dgD + (match ? 0 : 1), lfD + 1), upD + 1);
                        nD = min(nD, kMedMax);
                        if (nD <= band) { llo = min(llo, j); lhi = max(lhi, j); }
                    }
                }
                __stcg(M0 + j, nM); __stcg(X0 + j, nX); __stcg(Y0 + j, nY); __stcg(D0 + j, nD);
                if (j == ly - 1) {  // free end: the last row of column i
                    if (A::kLit) {
                        cols[3 * i] = nM; cols[3 * i + 1] = nX; cols[3 * i + 2] = nY;
                    } else {
                        const double v[3] = {nM, nX, nY};
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            if (v[k] == kNegInf) continue;
                            if (v[k] > amx) { asum = asum * exp(amx - v[k]) + 1.0; amx = v[k]; }
                            else asum += exp(v[k] - amx);
                        }
                    }
                }
            }
            // live range of this diagonal (rows whose distance lets a successor be computed)
            if (banded) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    llo = min(llo, __shfl_xor_sync(0xffffffffu, llo, o));
                    lhi = max(lhi, __shfl_xor_sync(0xffffffffu, lhi, o));
                }
            }
            w2lo = w1lo; w2hi = w1hi; l2lo = l1lo; l2hi = l1hi;
            w1lo = clo <= chi ? clo : kBig; w1hi = clo <= chi ? chi : -1;
            l1lo = llo; l1hi = lhi;
            b0 = b0 == 2 ? 0 : b0 + 1;
            __syncwarp();
        }
        double res;
        if (A::kLit) {
            __syncwarp();
            res = literal_ln_sum_exp_cols(cols, 3 * lx, lane);
        } else {
            // combine the lanes' (max, sum) pairs
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double omx = __shfl_xor_sync(0xffffffffu, amx, o);
                const double osum = __shfl_xor_sync(0xffffffffu, asum, o);
                if (omx > amx) { asum = asum * (amx == kNegInf ? 0.0 : exp(amx - omx)) + osum; amx = omx; }
                else if (omx != kNegInf) asum += osum * exp(omx - amx);
            }
            res = amx == kNegInf ? kNegInf : amx + log(asum);
        }
        if (lane == 0) {
            if (res > 0.0) res = 0.0;  // "sum of paths can exceed 1" cap
            g.out[pair] = res;
        }
        __syncwarp();
    }
}

// log_space_literal: LogLit (needs g.lit_ws); else Log
template <int APPROX>
cudaError_t launch_pairhmm_diag(bool literal, const HmmArgs &a, const DiagArgs &da, int n_ctas, cudaStream_t st);

#define VLR_PAIRHMM_DIAG_INSTANTIATE(APPROX)                                                      \
    template <>                                                                                   \
    cudaError_t launch_pairhmm_diag<APPROX>(bool literal, const HmmArgs &a, const DiagArgs &da,  \
                                            int n_ctas, cudaStream_t st) {                        \
        if (literal) pairhmm_diag_kernel<APPROX, LogLit><<<n_ctas, kWarps * 32, 0, st>>>(a, da); \
        else pairhmm_diag_kernel<APPROX, Log><<<n_ctas, kWarps * 32, 0, st>>>(a, da);            \
        return cudaGetLastError();                                                                \
    }

}  // namespace vlr
