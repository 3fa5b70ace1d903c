// pairhmm_litband.cuh -- the LITERAL log-space pair-HMM (policy LogLit of pairhmm_kernel.cuh: the
// reference's operation order with the library's deterministic libm, bit-identical to the oracle) for
// BANDED pairs, mapped the way the band wants it.
//
// The anti-diagonal wavefront of pairhmm_kernel (lanes = blocks of read rows) spends a step on every
// column in every lane: under the band `dist + 2` of allele support only ~3 of 32 lanes hold live cells
// at any step, and a literal cell costs ~2000 instructions (msun exp / log1p in software), so one
// near-tie pair took ~9 ms of one warp -- and the near ties of a batch (0.1 % of the reads) cost as
// much as everything else together.  Here the warp sweeps the READ ROWS and the lanes take the live
// CELLS of the row: the state of the previous row lives in shared memory indexed by diagonal
// (p = i - j + len_y - 1), so any lane can pick up any diagonal; a row costs
// ceil(candidates / 32) cell evaluations instead of 8.
//
// The cell set is exactly the reference's (bio's PairHMM::prob_related with a band): a cell is computed
// iff one of its three predecessors has a minimal edit distance <= k.  The left predecessor is a cell
// of the same row, which makes the band bookkeeping a recurrence along the row; as in pairhmm_band.cuh
// it is a min-plus prefix scan with the validity rule folded in (a chain of left moves is valid as long
// as its value is <= k + 1 where it is used), here over all diagonals of the row in rounds of 32.
// Unlike pairhmm_band.cuh no diagonal is left out: cells far off the hit's diagonals are live in the
// first few rows (every column is a free start), and the literal evaluation includes them.
//
// Needs a gap-extension probability of 0 on the haplotype side (ln = -inf: the X state of a cell is then
// a function of the M state of its left neighbour alone, which removes the floating-point recurrence
// along the row; the CLI default, cli.rs:219-230) and len_x + len_y <= kLbMaxP + 1.  Pairs it does not
// take are appended to `rest_list` for the wavefront kernel.
#pragma once
#include "pairhmm_kernel.cuh"

namespace vlr {

constexpr int kLbMaxP = 512;      // diagonals per pair (len_x + len_y - 1)
constexpr int kLbPad = 2;         // entries before / after the diagonals (never-written sentinels)
constexpr int kLbStride = kLbMaxP + 2 * kLbPad;

struct LitBandShared {
    double M[2][kLbStride], X[2][kLbStride], Y[2][kLbStride];
    int med[2][kLbStride];
};

template <int APPROX>
__global__ void __launch_bounds__(32) pairhmm_litband_kernel(HmmArgs g, int32_t *rest_list, int32_t *rest_count) {
    extern __shared__ __align__(16) unsigned char lb_smem_raw[];
    LitBandShared &S = *reinterpret_cast<LitBandShared *>(lb_smem_raw);
    const int lane = threadIdx.x;
    const HmmConsts c = g.ln;
    const int n = *g.count;
    const int32_t *list = g.list ? g.list + (g.list_off ? *g.list_off : 0) : nullptr;
    double *cols = g.lit_ws + (int64_t)blockIdx.x * 3 * g.lit_stride;
    const double kNegInf = -INFINITY;
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
        const int k = g.med ? g.med[pair] : -1;
        const int P = lx + ly - 1;
        if (k < 0 || lx <= 0 || ly <= 0 || P > kLbMaxP || lx > g.lit_stride || c.t_gye != kNegInf) {
            if (lane == 0) rest_list[atomicAdd(rest_count, 1)] = pair;
            continue;
        }
        // ---- reset both row buffers: "no mass, no distance" everywhere
        for (int q = lane; q < kLbStride; q += 32) {
            S.M[0][q] = S.X[0][q] = S.Y[0][q] = kNegInf;
            S.M[1][q] = S.X[1][q] = S.Y[1][q] = kNegInf;
            S.med[0][q] = S.med[1][q] = kMedMax;
        }
        __syncwarp();
        // The row above row 0 is the free start: mass 1 (ln 0) and distance 0 on the diagonal
        // predecessor of every cell (0, i), nothing above it (fm[prev][0] = 1, pairhmm step of column 0).
        // Diagonal index of cell (j, i): p = i - j + ly - 1; its diagonal predecessor has the same p, the
        // cell above p + 1, the cell to the left p - 1.
        int ob = 0;  // buffer holding the previous row
        for (int i = lane; i < lx; i += 32) {
            S.M[ob][kLbPad + i + ly - 1] = 0.0;
            S.med[ob][kLbPad + i + ly - 1] = 0;
        }
        __syncwarp();
        int lo_prev = ly - 1, hi_prev = lx + ly - 2;  // diagonals alive in the previous row
        bool any_alive = true;
        for (int j = 0; j < ly && any_alive; ++j) {
            const int nb = ob ^ 1;
            const int q = g.rqual[ro + j];
            const int rb = g.rbase[ro + j];
            const double ln_eps = dm::mul((double)q, -0.23025850929940456);
            const double pm = g.qlut[q * kQlutStride + 3];
            const double pmm = dm::add(ln_eps, kLnConfusion);
            // candidates of this row: a diagonal or upper predecessor alive in the previous row
            // ([lo_prev - 1, hi_prev]), or reached by a chain of left moves (at most k + 1 further right);
            // inside the matrix: column i = p - (ly - 1) + j in [0, lx)
            const int pmin = ly - 1 - j, pmax = lx + ly - 2 - j;
            const int cl = max(lo_prev - 1, pmin), ch = min(hi_prev + k + 1, pmax);
            // the new buffer still holds the row before the previous one: clear what this row does not
            // rewrite (rows shrink and shift by one diagonal per row, a margin of k + 4 covers it)
            for (int pz = cl - (k + 4) + lane; pz <= ch + (k + 4); pz += 32)
                if ((pz < cl || pz > ch) && pz >= -kLbPad && pz < kLbMaxP + kLbPad) {
                    S.M[nb][kLbPad + pz] = S.X[nb][kLbPad + pz] = S.Y[nb][kLbPad + pz] = kNegInf;
                    S.med[nb][kLbPad + pz] = kMedMax;
                }
            int carry = kMedMax * 4;          // min over the cells left of this round of (a - p)
            double prev_M = kNegInf;          // M of the cell just left of this round (new row)
            int new_lo = 1 << 30, new_hi = -(1 << 30);
            for (int base = cl; base <= ch; base += 32) {
                const int pp = base + lane;
                const bool inr = pp <= ch;
                const int at = kLbPad + (inr ? pp : cl);
                const int i = pp - (ly - 1) + j;
                int xb = inr ? (int)g.hap[ho + i] : 0;
                if ((unsigned)(xb - 'a') < 26u) xb -= 32;  // to_ascii_uppercase (pairhmm.rs:190)
                const bool match = xb == rb;
                // row 0: nothing above (the slot p + 1 of the start row is the NEXT column's free start)
                const int dmed = S.med[ob][at], umed = j == 0 ? kMedMax : S.med[ob][at + 1];
                const int a = min(dmed + (match ? 0 : 1), umed + 1);
                const bool strong = min(dmed, umed) <= k;
                // exclusive prefix minimum of (a - p) over the row: rounds of 32 with a carry
                int sc = inr ? a - pp : kMedMax * 4;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_up_sync(0xffffffffu, sc, o);
                    if (lane >= o) sc = min(sc, u);
                }
                int ex = __shfl_up_sync(0xffffffffu, sc, 1);
                if (lane == 0) ex = kMedMax * 4;
                const int Pq = min(carry, ex) + pp;      // left neighbour's distance + 1, if its chain is valid
                carry = min(carry, __shfl_sync(0xffffffffu, sc, 31));
                const bool chain = Pq <= k + 1;
                const bool alive = inr && (strong || chain);
                const int L = min(chain ? min(a, Pq) : a, kMedMax);
                // ---- the cell (PairHMM::prob_related, literal operation order; see hmm_step)
                double nM = kNegInf, nY = kNegInf;
                if (alive) {
                    unsigned trk_unused = 0;
                    double wj_unused = 0.0;
                    const double e = match ? pm : pmm;
                    nM = LogLit::mul(e, LogLit::template sum3<APPROX, true, false>(c, S.M[ob][at], S.X[ob][at], S.Y[ob][at],
                                                                                   trk_unused, wj_unused));
                    const double upM = j == 0 ? kNegInf : S.M[ob][at + 1], upY = j == 0 ? kNegInf : S.Y[ob][at + 1];
                    nY = LogLit::mul(ln_eps, LogLit::add(LogLit::mul(c.t_gx, upM), LogLit::mul(c.t_gxe, upY)));
                }
                double left_M = __shfl_up_sync(0xffffffffu, nM, 1);
                if (lane == 0) left_M = prev_M;
                prev_M = __shfl_sync(0xffffffffu, nM, 31);
                // X: emit_x = 1; ln_add_exp(gap_y + M[left], gap_y_ext + X[left]) with gap_y_ext = ln 0
                const double nX = alive ? LogLit::template dot2<true>(c.t_gy, left_M, c.t_gye, kNegInf) : kNegInf;
                if (inr) {
                    S.M[nb][at] = nM; S.X[nb][at] = nX; S.Y[nb][at] = nY;
                    S.med[nb][at] = alive ? L : kMedMax;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, alive);
                if (bal) {
                    new_lo = min(new_lo, base + (__ffs(bal) - 1));
                    new_hi = max(new_hi, base + (31 - __clz(bal)));
                }
            }
            __syncwarp();
            any_alive = new_hi >= new_lo;
            lo_prev = new_lo; hi_prev = new_hi;
            ob = nb;
        }
        // ---- prob_cols: (M, X, Y) of the last read row per column (diagonal p = column), then
        // LogProb::ln_sum_exp over them: first maximum, then the sum of exp(p - max) over the other
        // finite entries in slice order, accumulated by ONE lane in that order
        const int nc = 3 * lx;
        for (int i = lane; i < lx; i += 32) {
            const bool have = any_alive && i >= lo_prev - (k + 4) && i <= hi_prev + (k + 4);
            cols[3 * i] = have ? S.M[ob][kLbPad + i] : kNegInf;
            cols[3 * i + 1] = have ? S.X[ob][kLbPad + i] : kNegInf;
            cols[3 * i + 2] = have ? S.Y[ob][kLbPad + i] : kNegInf;
        }
        __syncwarp();
        double pmax = -INFINITY;
        for (int t = lane; t < nc; t += 32) pmax = fmax(pmax, cols[t]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pmax = fmax(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
        int imax = nc;
        for (int t = lane; t < nc; t += 32)
            if (cols[t] == pmax) { imax = t; break; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) imax = min(imax, __shfl_xor_sync(0xffffffffu, imax, o));
        double res = pmax;
        if (pmax != -INFINITY && pmax != INFINITY) {
            for (int t = lane; t < nc; t += 32) {
                const double v = cols[t];
                cols[t] = (t == imax || v == -INFINITY) ? 0.0 : dm::exp(dm::sub(v, pmax));
            }
            __syncwarp();
            if (lane == 0) {
                double sum = 0.0;
                for (int t = 0; t < nc; ++t) sum = dm::add(sum, cols[t]);  // + 0.0 is exact
                res = dm::add(pmax, dm::log1p(sum));
            }
        }
        if (lane == 0) {
            if (res > 0.0) res = 0.0;
            g.out[pair] = res;
        }
        __syncwarp();
    }
}

template <int APPROX>
cudaError_t launch_pairhmm_litband(const HmmArgs &a, int32_t *rest_list, int32_t *rest_count, int n_ctas,
                                   cudaStream_t st);

#define VLR_PAIRHMM_LITBAND_INSTANTIATE(APPROX)                                                   \
    template <>                                                                                   \
    cudaError_t launch_pairhmm_litband<APPROX>(const HmmArgs &a, int32_t *rest_list, int32_t *rest_count, \
                                               int n_ctas, cudaStream_t st) {                     \
        cudaError_t e = cudaFuncSetAttribute(pairhmm_litband_kernel<APPROX>,                      \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                             (int)sizeof(LitBandShared));                         \
        if (e != cudaSuccess) return e;                                                           \
        pairhmm_litband_kernel<APPROX><<<n_ctas, 32, sizeof(LitBandShared), st>>>(a, rest_list, rest_count); \
        return cudaGetLastError();                                                                \
    }

}  // namespace vlr
