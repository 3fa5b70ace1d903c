// pairhmm_band.cuh -- K1b: banded pair-HMM that computes only the band (sm_100a).
//
// PairHMM::prob_related with `max_edit_dist = Some(k)` (realignment/mod.rs:439-443, Appendix A.2)
// skips every cell whose three predecessors have a minimal edit distance above k.  On the windows
// that `shrink_to_hit` leaves (len_x ~ len_y + dist + 4, k = dist + 2) the live cells lie on the
// diagonals i - j in [-(k+1), len_x - len_y + k + 1]: a path starts at a diagonal >= 0 (free start
// in x, first read row) and must end at one <= len_x - len_y, and every diagonal it moves costs a
// gap.  The anti-diagonal wavefront of pairhmm_kernel sweeps all len_x columns in every lane; this
// kernel maps the DIAGONALS to lanes and sweeps the read rows instead:
//
//   * lane l owns the G consecutive diagonals dlo + l*G + g; row j computes the cells
//     (i = j + diagonal, j).  The diagonal predecessor is the lane's own previous row (registers),
//     the row above is the next diagonal's previous row (one shuffle), the cell to the left is the
//     previous diagonal's CURRENT row.  With gap-extension probabilities of 0 (CLI default) only the
//     X state (one shuffle of the finished M values) and the band bookkeeping depend on the left
//     cell; the latter, L_i = min(A_i, L_{i-1} + 1 if L_{i-1} <= k), is a min-plus prefix scan over
//     the lanes (five shuffle rounds of 32-bit integers).
//   * per pair only two bytes per read row (quality, base code) and one byte per haplotype base
//     are staged in shared memory; the emission / Y-state factors come from a per-CTA table indexed
//     by the quality (6 KB), so that 6 to 12 CTAs fit an SM -- the row loop is one dependent chain
//     per warp (shuffles, the e^-10 rule, the scan) and needs the warps to hide its latency.
//
// What the band leaves out are paths (and liveness effects at the two edge diagonals) that need at
// least k + 1 = dist + 3 gap openings in one direction; with the gap probabilities this kernel
// accepts (<= 1e-4, the CLI defaults are 2.8e-6 / 5.1e-6) that mass is below 1e-9 of the result in
// the worst case met (homopolymers) and below fp64 resolution otherwise -- see the parity tests.
// Pairs it does not take (gap extension, unbanded, wide or long windows, non-ACGT read bases) stay
// with pairhmm_kernel<.., BANDED = true, ..>.
#pragma once
#include "pairhmm_kernel.cuh"

namespace vlr {

constexpr int kBandMaxLx = 480;   // haplotype window bytes staged per warp
constexpr int kBandMaxLy = 248;   // read rows staged per warp (same limit as pairhmm_kernel)
constexpr int kBandPad = 128;     // codes before / after the window (cells outside are masked)
constexpr int kBandInf = 1 << 28; // "no predecessor" in the min-plus scan (kMedMax + positions fit)
constexpr int kBandMaxG = 4;      // diagonals per lane: bands of up to 128 diagonals
// CTAs per SM by diagonals per lane (40 / 56 / 71 registers; 11 KB shared memory per CTA)
constexpr int band_blocks_per_sm(int G) { return G == 1 ? 12 : (G == 2 ? 8 : (G == 3 ? 6 : 5)); }

struct BandQ {  // per base quality: emission factors (already times t_mm) and the Y-state factor
    double em, emm, cy;
};

template <int G, int APPROX>
__global__ void __launch_bounds__(kWarps * 32, band_blocks_per_sm(G))
pairhmm_band_kernel(HmmArgs g, int32_t *list_rw, const int32_t *n_total) {
    __shared__ __align__(16) uint8_t s_hap[kWarps][kBandPad + kBandMaxLx + kBandPad];
    __shared__ __align__(4) uint16_t s_row[kWarps][kBandMaxLy];   // quality | base code << 8
    __shared__ __align__(16) BandQ s_q[256];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint8_t *hap = s_hap[warp];
    uint16_t *rows = s_row[warp];
    const HmmConsts c = g.lin;
    const int n = *n_total;
    for (int q = threadIdx.x; q < 256; q += kWarps * 32) {
        BandQ e;
        e.em = g.qlut[q * kQlutStride + 1] * c.t_mm;
        e.emm = g.qlut[q * kQlutStride + 2] * c.t_mm;
        e.cy = g.qlut[q * kQlutStride + 0] * c.t_gx_s;
        s_q[q] = e;
    }
    __syncthreads();

    for (;;) {
        int p = 0;
        if (lane == 0) p = atomicAdd(g.cursor, 1);
        p = __shfl_sync(0xffffffffu, p, 0);
        if (p >= n) break;
        const int pair = list_rw[p];
        if (pair < 0) continue;  // taken by an earlier band launch
        const int k = g.med[pair];
        if (k < 0) continue;     // unbanded pair
        const int64_t hidx = g.pair_hap ? (int64_t)g.pair_hap[pair] : (int64_t)pair;
        const int64_t rd = g.pair_read ? (int64_t)g.pair_read[pair] : (int64_t)pair;
        const int64_t ho = g.win_start ? g.win_start[pair] : g.hap_off[hidx];
        const int lx = g.win_start ? g.win_len[pair] : (int)(g.hap_off[hidx + 1] - ho);
        const int64_t ro = g.read_off[rd];
        const int ly = (int)(g.read_off[rd + 1] - ro);
        const int dhi = lx - ly;                  // highest diagonal a path can end on
        const int W = dhi + 2 * k + 3;            // diagonals -(k+1) .. dhi + k + 1
        if (dhi < 0 || ly < 1 || ly > kBandMaxLy || lx > kBandMaxLx || W > 32 * G || (G > 1 && W <= 32 * (G - 1)))
            continue;
        // ---- stage the read rows; a base outside ACGT leaves the pair to the byte-compare kernel
        bool odd = false;
        __syncwarp();
        for (int j = lane; j < ly; j += 32) {
            const int q = g.rqual[ro + j];
            const int code = base_code1(g.rbase[ro + j]);
            odd |= code == 4;
            rows[j] = (uint16_t)(q | (code << 8));
        }
        if (__any_sync(0xffffffffu, odd)) continue;
        for (int i = lane; i < lx + 2 * kBandPad; i += 32) {
            const int x = i - kBandPad;
            hap[i] = (x >= 0 && x < lx) ? (uint8_t)base_code1(g.hap[ho + x]) : (uint8_t)4;
        }
        if (lane == 0) list_rw[p] = -1 - pair;   // the class kernels skip this entry
        __syncwarp();

        const int dlo = -(k + 1);
        // state of the previous row on my diagonals.  Row -1 is the free start: mass 1 and edit
        // distance 0 on the diagonal predecessor of every cell (i, 0) of the matrix, nothing above it.
        // Cells left of the matrix stay dead by themselves (all three predecessors are), cells right
        // of it or beyond the band are masked by `i < ilim`.
        double Mp[G], Xp[G], Yp[G];
        int medp[G];
#pragma unroll
        for (int q = 0; q < G; ++q) {
            const int pos = lane * G + q, i0 = dlo + pos;
            const bool in0 = pos < W && i0 >= 0 && i0 < lx;
            Mp[q] = in0 ? c.one_s : 0.0; Xp[q] = 0.0; Yp[q] = 0.0; medp[q] = in0 ? 0 : kMedMax;
        }
        double accM = 0.0, accXY = 0.0;
        bool any_alive = true;
        unsigned trk = 0xffffffffu;  // three-way rule: closest approach of a high word to a threshold
        double wj = 0.0;
        // one read row; FIRST: the row above is the reset state of the column (no mass, no distance)
        auto row = [&](int j, auto first_tag) {
            constexpr bool FIRST = decltype(first_tag)::value;
            const int rq = rows[j];
            const BandQ qe = s_q[rq & 0xff];
            const int rcode = rq >> 8;
            double upM_next = 0.0;
            int upMed_next = kMedMax;
            if (!FIRST) {
                upM_next = __shfl_down_sync(0xffffffffu, Mp[0], 1);
                upMed_next = __shfl_down_sync(0xffffffffu, medp[0], 1);
                if (lane == 31) { upM_next = 0.0; upMed_next = kMedMax; }
            }
            const int ibase = j + dlo + lane * G;
            const int ilim = min(lx, j + dlo + W);
            double sumv[G], ev[G], upMv[G];
            int a_med[G], E[G];
            bool in[G], strong[G];
            int vmin = kBandInf;          // min over my diagonals of (a - position)
#pragma unroll
            for (int q = 0; q < G; ++q) {
                const int i = ibase + q;
                in[q] = i < ilim;
                upMv[q] = FIRST ? 0.0 : (q + 1 < G ? Mp[q + 1 < G ? q + 1 : 0] : upM_next);
                const int upMed = FIRST ? kMedMax : (q + 1 < G ? medp[q + 1 < G ? q + 1 : 0] : upMed_next);
                const bool match = hap[kBandPad + i] == rcode;
                ev[q] = match ? qe.em : qe.emm;
                sumv[q] = Lin::sum3<APPROX, false, true>(c, Mp[q], Xp[q], Yp[q], trk, wj);
                const int m_tl = medp[q];
                a_med[q] = min(m_tl + (match ? 0 : 1), upMed + 1);
                strong[q] = min(m_tl, upMed) <= k;
                E[q] = vmin;              // exclusive prefix inside the lane
                // a cell without a predecessor within k has a >= k + 1 and can never start a chain
                // (its candidate exceeds k + 1 one step later), so it enters the scan unmasked
                vmin = min(vmin, a_med[q] - (lane * G + q));
            }
            // exclusive min-scan of the lanes' minima
            int S = vmin;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, S, o);
                if (lane >= o) S = min(S, u);
            }
            S = __shfl_up_sync(0xffffffffu, S, 1);
            if (lane == 0) S = kBandInf;
            // liveness and band bookkeeping; dead cells are zeroed by a 0/1 factor on the fp64 pipe
            // (the kernel is bound by the ALU pipe: selects and integer min/compare)
            double nM[G], wv[G];
            bool row_alive = false;
#pragma unroll
            for (int q = 0; q < G; ++q) {
                const int pos = lane * G + q;
                const int Pq = min(S, E[q]) + pos;       // best predecessor chain ending left of me
                const bool chain = Pq <= k + 1;
                const bool alive = in[q] && (strong[q] || chain);
                const int L = chain ? min(a_med[q], Pq) : a_med[q];
                medp[q] = alive ? L : kMedMax;
                wv[q] = __hiloint2double(alive ? 0x3ff00000 : 0, 0);
                nM[q] = (ev[q] * wv[q]) * sumv[q];
                Yp[q] = (qe.cy * wv[q]) * upMv[q];
                row_alive |= alive;
            }
            double leftM = __shfl_up_sync(0xffffffffu, nM[G - 1], 1);
            if (lane == 0) leftM = 0.0;
#pragma unroll
            for (int q = 0; q < G; ++q) {
                const double lm = q == 0 ? leftM : nM[q > 0 ? q - 1 : 0];
                Xp[q] = (c.t_gy_s * wv[q]) * lm;
                Mp[q] = nM[q];
            }
            any_alive = __any_sync(0xffffffffu, row_alive);
        };
        row(0, std::true_type());
        #pragma unroll 1
        for (int j = 1; j < ly && any_alive; ++j) row(j, std::false_type());
        // ---- free end: the last read row of every column
        if (any_alive) {
#pragma unroll
            for (int q = 0; q < G; ++q) { accM += Mp[q]; accXY += Xp[q] + Yp[q]; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            accM += __shfl_xor_sync(0xffffffffu, accM, o);
            accXY += __shfl_xor_sync(0xffffffffu, accXY, o);
        }
        if (APPROX == 2 && __reduce_min_sync(0xffffffffu, trk) <= 8u) {
            // a doubtful decision of the three-way rule: the exact-decision kernel takes the pair
            if (lane == 0) g.gen_list[atomicAdd(g.gen_count, 1)] = pair;
            continue;
        }
        if (lane == 0) {
            const double acc = fma(accM, c.inv_tmm, accXY);
            double res;
            if (!(acc >= 4.3e-283 && acc < 1.0e307)) {   // outside the scaled range: log-space pass
                g.fb_list[atomicAdd(g.fb_count, 1)] = pair;
                res = NAN;
            } else {
                res = log(acc) - c_scale_ln;
            }
            if (res > 0.0) res = 0.0;
            g.out[pair] = res;
        }
    }
}

template <int APPROX>
cudaError_t launch_pairhmm_band(const HmmArgs &a, int32_t *list_rw, const int32_t *n_total, int32_t *cursors,
                                int sm_count, cudaStream_t st);

#define VLR_PAIRHMM_BAND_INSTANTIATE(APPROX)                                                      \
    template <>                                                                                   \
    cudaError_t launch_pairhmm_band<APPROX>(const HmmArgs &a, int32_t *list_rw, const int32_t *n_total, \
                                            int32_t *cursors, int sm_count, cudaStream_t st) {    \
        HmmArgs b = a;                                                                            \
        b.cursor = cursors + 0;                                                                   \
        pairhmm_band_kernel<1, APPROX><<<sm_count * band_blocks_per_sm(1), kWarps * 32, 0, st>>>(b, list_rw, n_total); \
        b.cursor = cursors + 1;                                                                   \
        pairhmm_band_kernel<2, APPROX><<<sm_count * band_blocks_per_sm(2), kWarps * 32, 0, st>>>(b, list_rw, n_total); \
        b.cursor = cursors + 2;                                                                   \
        pairhmm_band_kernel<3, APPROX><<<sm_count * band_blocks_per_sm(3), kWarps * 32, 0, st>>>(b, list_rw, n_total); \
        b.cursor = cursors + 3;                                                                   \
        pairhmm_band_kernel<4, APPROX><<<sm_count * band_blocks_per_sm(4), kWarps * 32, 0, st>>>(b, list_rw, n_total); \
        return cudaGetLastError();                                                                \
    }

}  // namespace vlr
