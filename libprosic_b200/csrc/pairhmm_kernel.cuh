// pairhmm_kernel.cuh -- K1: anti-diagonal wavefront pair-HMM forward kernel for sm_100a.
//
// Replaces bio::stats::pairhmm::PairHMM::prob_related as called from
// PairHMMRealigner::calculate_prob_allele (reference src/variants/evidence/realignment/mod.rs:430-444)
// with the emission model of realignment/pairhmm.rs:123-211 and the semiglobal start/end rules of
// pairhmm.rs:103-121.  Recurrence (x = haplotype column i, y = read row j):
//   M[i][j] = e_xy(i,j) * sum3( t_mm*M[i-1][j-1], t_xm*X[i-1][j-1], t_ym*Y[i-1][j-1] )
//   X[i][j] =             t_gy*M[i-1][j] + t_gye*X[i-1][j]            (hap-only state, emit 1)
//   Y[i][j] = eps_j *   ( t_gx*M[i][j-1] + t_gxe*Y[i][j-1] )          (read-only state)
//   M[i][0] = 1 for every column (free start), result = sum_i (M+X+Y)[i][len_y] (free end), capped at 1.
//
// Mapping: one warp per (read, haplotype) pair.  Lane l owns R consecutive read rows
// [l*R, l*R+R) with R = ceil(len_y / 31) (lane 31 is always padding); at step t it computes
// column i = t - l for its rows, so the warp sweeps
// anti-diagonals.  The bottom row of lane l-1 travels to lane l by __shfl_up_sync (M, X, Y, and the
// band's min-edit-distance); the haplotype window is staged into a per-warp shared-memory ring by
// 1-D TMA bulk copies (cp.async.bulk + mbarrier), read bases/qualities become per-row register
// constants from a 256-entry LUT.  Arithmetic is fp64 in SCALED LINEAR space (start mass 2^960,
// no per-cell transcendental, one log per pair); pairs whose result leaves the representable
// range are re-run by the same kernel instantiated with log-space arithmetic (still on the GPU).
#pragma once
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "detmath.cuh"

namespace vlr {

constexpr int kWarps = 4;          // warps per CTA; each warp owns one pair at a time
constexpr int kChunk = 256;        // haplotype bytes per TMA bulk copy
constexpr int kRing = 2 * kChunk;  // per-warp ring buffer (two chunks)
constexpr int kMedMax = 1 << 20;   // "usize::MAX" of the band bookkeeping
constexpr int kNumClasses = 8;     // R = 1,2,3,4,5,6,8 and "too long"
constexpr double kLnConfusion = -1.098712293668443;  // LogProb::from(Prob(0.3333)), pairhmm.rs:21-23

// 2^960 and its log; linear-space values are scaled by this (see header comment)
#define VLR_SCALE 9.7453140114e288
__device__ __constant__ double c_scale_ln = 665.4212933375474;  // 960 * ln 2

struct HmmConsts {  // linear or ln, depending on the arithmetic policy
    double t_mm, t_xm, t_ym, t_gy, t_gye, t_gx, t_gxe;
    // scaled-state variant (the M state carries the factor t_mm, see hmm_step<..., LUT = true>):
    // t_gy / t_mm, t_gx / t_mm, 1 / t_mm, 2^960 * t_mm  (linear space only)
    double t_gy_s, t_gx_s, inv_tmm, one_s;
};

struct HmmArgs {
    const uint8_t *hap;
    const int64_t *hap_off;
    const uint8_t *rbase;
    const uint8_t *rqual;
    const int64_t *read_off;
    const int32_t *pair_hap;
    const int32_t *pair_read;
    const int32_t *med;
    const int64_t *win_start;  // optional explicit haplotype windows (fused allele support)
    const int32_t *win_len;
    const int32_t *list;      // pair ids handled by this launch (device)
    const int32_t *list_off;  // device offset into `list` (nullable)
    const int32_t *count;  // number of entries of `list` (device)
    int32_t *cursor;       // dynamic work counter (device, zero at launch)
    int32_t *fb_list;      // pairs that need the log-space pass
    int32_t *fb_count;
    int32_t *gen_list;     // pairs whose read holds a non-ACGT base: generic (select) kernel
    int32_t *gen_count;
    double *long_ws;       // long-read variant: per-warp boundary rows, 6 * long_stride doubles
    int64_t long_stride;   // columns per boundary row
    double *lit_ws;        // literal log-space variant: per-warp prob_cols, 3 * lit_stride doubles
    int64_t lit_stride;    // columns per warp
    double *out;
    const double *qlut;
    HmmConsts lin, ln;
};

// ---------------------------------------------------------------- PTX helpers (TMA + mbarrier)
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "VLR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra VLR_DONE;\n"
        "bra VLR_WAIT;\n"
        "VLR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---------------------------------------------------------------- arithmetic policies
__device__ __forceinline__ double ln_add_exp_d(double a, double b) {
    double p0 = fmax(a, b), p1 = fmin(a, b);
    if (p0 == -INFINITY) return -INFINITY;
    if (p1 == -INFINITY) return p0;
    return p0 + log1p(exp(p1 - p0));
}

// EXT = false is the specialisation for gap-extension probabilities of exactly 0 (the CLI
// default, cli.rs:219-230): then t_xm = t_ym = 1, t_gye = t_gxe = 0 and three products vanish.
struct Lin {  // scaled linear space
    static constexpr bool kLog = false;
    static constexpr bool kLit = false;
    __device__ static double zero() { return 0.0; }
    __device__ static double one() { return VLR_SCALE; }
    __device__ static double mul(double a, double b) { return a * b; }
    __device__ static double add(double a, double b) { return a + b; }
    // a*b + c*d   (EXT = false: c*d == 0)
    template <bool EXT>
    __device__ static double dot2(double a, double b, double c, double d) {
        return EXT ? fma(a, b, c * d) : a * b;
    }
    // sum over the three predecessors of the M state (SCALED: dM already carries t_mm).
    // APPROX: 0 = exact sum, 1 = two-way, 2 = three-way ln_sum3_exp_approx (see vlr_gpu.h flags).
    __device__ static double sum3_exact3(double a, double b, double c) {
        const double kT = 4.5399929762484854e-05;
        const bool bc = b >= c;
        const double hi1 = bc ? b : c, lo1 = bc ? c : b;
        const bool ah = a >= hi1;
        const double p0 = ah ? a : hi1, m = ah ? hi1 : a;
        const bool ml = m >= lo1;
        const double p1 = ml ? m : lo1, p2 = ml ? lo1 : m;
        const double t2 = p2 >= p1 * kT ? p1 + p2 : p1;
        return p1 >= p0 * kT ? p0 + t2 : p0;
    }
    template <int APPROX, bool EXT, bool SCALED = false>
    __device__ static double sum3(const HmmConsts &k, double dM, double dX, double dY, unsigned &trk, double &wjunk) {
        if (APPROX == 1) {
            // two-way rule: return the maximum alone if the second largest term is
            // more than e^-10 below it, else the exact sum.  "second < max*e^-10" <=> exactly one of
            // the three terms reaches max*e^-10.  Written in PTX so that the three threshold
            // predicates feed one select (the C++ form compiles to nested 64-bit selects).
            const double a = SCALED ? dM : k.t_mm * dM;
            const double b = EXT ? k.t_xm * dX : dX;
            const double c = EXT ? k.t_ym * dY : dY;
            const double s = a + b + c;
            double r;
            asm("{\n\t"
                ".reg .pred p, q, pa, pb, pc, t, u;\n\t"
                ".reg .f64 hi, mx, thr;\n\t"
                "setp.gt.f64 p, %1, %2;\n\t"
                "selp.f64 hi, %1, %2, p;\n\t"
                "setp.gt.f64 q, hi, %3;\n\t"
                "selp.f64 mx, hi, %3, q;\n\t"
                "mul.f64 thr, mx, 0d3F07CD79B5647C9B;\n\t"  // e^-10
                "setp.ge.f64 pa, %1, thr;\n\t"
                "setp.ge.f64 pb, %2, thr;\n\t"
                "setp.ge.f64 pc, %3, thr;\n\t"
                "and.pred t, pa, pb;\n\t"
                "or.pred u, pa, pb;\n\t"
                "and.pred u, u, pc;\n\t"
                "or.pred t, t, u;\n\t"
                "selp.f64 %0, %4, mx, t;\n\t"
                "}" : "=d"(r) : "d"(a), "d"(b), "d"(c), "d"(s));
            return r;
        }
        if (APPROX == 3) {
            // three-way rule, exact decisions on the fp64 values: sorted p0 >= p1 >= p2;
            // p1 < p0*e^-10 -> p0; else p2 < p1*e^-10 -> p0 + p1; else all three
            const double a = SCALED ? dM : k.t_mm * dM;
            const double b = EXT ? k.t_xm * dX : dX;
            const double c = EXT ? k.t_ym * dY : dY;
            return sum3_exact3(a, b, c);
        }
        if (APPROX == 2 || APPROX == 4) {
            // three-way rule, decided on the HIGH WORDS of the terms (non-negative doubles order like
            // their bit patterns).  Sorted p0 >= p1 >= p2: p1 < p0*e^-10 -> p0; else p2 < p1*e^-10 ->
            // p0 + p1; else all three.  Equivalent without a sort: level 1 = the terms that reach
            // e^-10 * max; the SMALLEST level-1 term sets the second threshold; every term that reaches
            // it is summed (one level-1 term: the maximum alone; two: p2 is tested against p1; three:
            // everything passes).
            //   t = hi(fma((h, *), kTb, kF)) is the high word of the threshold that belongs to a term with
            // high word h.  The low word of the factor is left as it is (whatever the register pair holds),
            // kTb = e^-10 * (1 - 2.5 * 2^-20): the computed threshold lies 0.5 .. 6 units below the high
            // word of the exact one, so a term whose high word is below t is certainly excluded, one more
            // than 8 units above certainly included, and only [t, t + 8] needs the low words: `trk` keeps
            // the closest approach, and the caller re-evaluates with the exact rule -- the whole pair
            // (APPROX == 2: the fallback kernel is the exact APPROX == 3 instantiation) or this cell after
            // a warp vote (APPROX == 4).
            //   kF = 2^-1018 floors the thresholds: cells whose largest term is below ~2^-982 (the pair
            // is only valid above 2^-940 and a cell contributes at most its own value) get inflated
            // thresholds or are flushed to zero, and all-zero cells never look like ties.
            const double a = SCALED ? dM : k.t_mm * dM;
            const double b = EXT ? k.t_xm * dX : dX;
            const double c = EXT ? k.t_ym * dY : dY;
            const int ha = __double2hiint(a), hb = __double2hiint(b), hc = __double2hiint(c);
            const double kTb = 4.53998215206174e-05;                 // e^-10 * (1 - 2.5 * 2^-20)
            const double kF = 3.5601181736115222e-307;               // 2^-1018
            // w = (don't care, max(ha, hb, hc)): the maximum is written straight into the high half of a
            // register pair whose low half is whatever the last cell left there (`wjunk`, one pair per row)
            double w = wjunk;
            asm("{\n\t.reg .b32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\tmax.s32 hi, %1, %2;\n\tmax.s32 hi, hi, %3;\n\t"
                "mov.b64 %0, {lo, hi};\n\t}" : "+d"(w) : "r"(ha), "r"(hb), "r"(hc));
            w = fma(w, kTb, kF);
            const int t0 = __double2hiint(w);
            const unsigned um = __vimin3_u32((unsigned)(ha - t0), (unsigned)(hb - t0), (unsigned)(hc - t0));
            // high word += um, in place (the register pair keeps whatever low word it has)
            asm("{\n\t.reg .b32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\tadd.s32 hi, hi, %1;\n\tmov.b64 %0, {lo, hi};\n\t}"
                : "+d"(w) : "r"(um));
            w = fma(w, kTb, kF);
            const int t1 = __double2hiint(w);
            wjunk = w;
            const int va = ha - t1, vb = hb - t1, vc = hc - t1;
            const unsigned near = __vimin3_u32(um, (unsigned)va, (unsigned)vb);
            // excluded terms are cleared by zeroing their HIGH word (what is left of the low word is a
            // subnormal below 2^-1042, far under every threshold of this kernel)
            double a2 = a, b2 = b, c2 = c;
            asm("{\n\t.reg .b32 lo, hi;\n\t.reg .pred p;\n\tmov.b64 {lo, hi}, %0;\n\tsetp.lt.s32 p, %1, 0;\n\t"
                "@p mov.b32 hi, 0;\n\tmov.b64 %0, {lo, hi};\n\t}" : "+d"(a2) : "r"(va));
            asm("{\n\t.reg .b32 lo, hi;\n\t.reg .pred p;\n\tmov.b64 {lo, hi}, %0;\n\tsetp.lt.s32 p, %1, 0;\n\t"
                "@p mov.b32 hi, 0;\n\tmov.b64 %0, {lo, hi};\n\t}" : "+d"(b2) : "r"(vb));
            asm("{\n\t.reg .b32 lo, hi;\n\t.reg .pred p;\n\tmov.b64 {lo, hi}, %0;\n\tsetp.lt.s32 p, %1, 0;\n\t"
                "@p mov.b32 hi, 0;\n\tmov.b64 %0, {lo, hi};\n\t}" : "+d"(c2) : "r"(vc));
            double r = a2 + (b2 + c2);
            if (APPROX == 4) {
                if (__any_sync(0xffffffffu, min(near, (unsigned)vc) <= 8u)) r = sum3_exact3(a, b, c);
            } else {
                trk = __vimin3_u32(trk, near, (unsigned)vc);
            }
            return r;
        }
        if (SCALED) return EXT ? fma(k.t_ym, dY, fma(k.t_xm, dX, dM)) : dM + (dX + dY);
        if (EXT) return fma(k.t_ym, dY, fma(k.t_xm, dX, k.t_mm * dM));
        return fma(k.t_mm, dM, dX + dY);
    }
};

struct Log {  // natural-log space, the arithmetic of the reference (CUDA libm)
    static constexpr bool kLog = true;
    static constexpr bool kLit = false;
    __device__ static double zero() { return -INFINITY; }
    __device__ static double one() { return 0.0; }
    __device__ static double mul(double a, double b) { return a + b; }
    __device__ static double add(double a, double b) { return ln_add_exp_d(a, b); }
    template <bool EXT>
    __device__ static double dot2(double a, double b, double c, double d) {
        return ln_add_exp_d(a + b, c + d);
    }
    template <int APPROX, bool EXT, bool SCALED = false>
    __device__ static double sum3(const HmmConsts &k, double dM, double dX, double dY, unsigned &, double &) {
        double p0 = k.t_mm + dM, p1 = k.t_xm + dX, p2 = k.t_ym + dY, t;
        if (APPROX) {
            if (p1 < p2) { t = p1; p1 = p2; p2 = t; }
            if (p1 > p0) { t = p1; p1 = p0; p0 = t; }
            if (p2 > p1) { t = p2; p2 = p1; p1 = t; }
            if (p0 - p1 > 10.0) return p0;
            if (APPROX >= 2 && p1 - p2 > 10.0) return ln_add_exp_d(p0, p1);
        } else {
            // ln_sum_exp over the slice (M, X, Y): the first maximum leads, the others follow in order
            if (p2 > p0 && p2 > p1) { t = p2; p2 = p1; p1 = p0; p0 = t; }
            else if (p1 > p0) { t = p1; p1 = p0; p0 = t; }
        }
        if (p0 == -INFINITY) return -INFINITY;
        double s = 0.0;
        if (p1 != -INFINITY) s += exp(p1 - p0);
        if (p2 != -INFINITY) s += exp(p2 - p0);
        return p0 + log1p(s);
    }
};

// Natural-log space in the LITERAL operation order of the reference with the library's
// deterministic libm (detmath.cuh): every value is bit-identical to a host evaluation of the same
// formulas (bio::stats::pairhmm::PairHMM::prob_related as restated in SURVEY Appendix A.2).  Slow;
// used for the pairs whose result decides an exact comparison (allele-support near ties).
struct LogLit {
    static constexpr bool kLog = true;
    static constexpr bool kLit = true;
    __device__ static double zero() { return -INFINITY; }
    __device__ static double one() { return 0.0; }
    __device__ static double mul(double a, double b) { return dm::add(a, b); }
    __device__ static double add(double a, double b) { return dm::ln_add_exp(a, b); }
    template <bool EXT>
    __device__ static double dot2(double a, double b, double c, double d) {
        return dm::ln_add_exp(dm::add(a, b), dm::add(c, d));
    }
    template <int APPROX, bool EXT, bool SCALED = false>
    __device__ static double sum3(const HmmConsts &k, double dM, double dX, double dY, unsigned &, double &) {
        double p0 = dm::add(k.t_mm, dM), p1 = dm::add(k.t_xm, dX), p2 = dm::add(k.t_ym, dY), t;
        if (APPROX) {
            if (p1 < p2) { t = p1; p1 = p2; p2 = t; }
            if (p1 > p0) { t = p1; p1 = p0; p0 = t; }
            if (p2 > p1) { t = p2; p2 = p1; p1 = t; }
            if (dm::sub(p0, p1) > 10.0) return p0;
            if (APPROX >= 2 && dm::sub(p1, p2) > 10.0) return dm::ln_add_exp(p0, p1);
        } else {
            if (p2 > p0 && p2 > p1) { t = p2; p2 = p1; p1 = p0; p0 = t; }
            else if (p1 > p0) { t = p1; p1 = p0; p0 = t; }
        }
        if (p0 == -INFINITY) return -INFINITY;
        if (p0 == INFINITY) return INFINITY;
        double s = 0.0;
        if (p1 != -INFINITY) s = dm::add(s, dm::exp(dm::sub(p1, p0)));
        if (p2 != -INFINITY) s = dm::add(s, dm::exp(dm::sub(p2, p0)));
        return dm::add(p0, dm::log1p(s));
    }
};

// ---------------------------------------------------------------- the kernel
#ifndef VLR_OCC5
#define VLR_OCC5 4
#endif
template <int R>
struct Occ {
    static constexpr int kMinBlocks = R <= 3 ? 5 : (R <= 4 ? 4 : (R <= 5 ? VLR_OCC5 : (R <= 6 ? 3 : 2)));
};

// Per-pair state that lives in registers across the sweep.
template <int R>
struct PairState {
    double pm[R], pmm[R], cyo[R], cye[R];  // per-row emission / Y-transition constants
    int rb[R];                             // read base per row (-1 = padding row)
    double M[R], X[R], Y[R];               // previous column of my rows
    int med[R];                            // band bookkeeping (min edit distance)
    double bM, bX, bY, dM, dX, dY;         // bottom row of last column; diagonal inputs of row 0
    int bMed, dMed;
    double acc, acc2;                      // free-end sums (scaled state: M' and X + Y apart)
    double acc3;                           // literal variant: (acc, acc2, acc3) = last row (M, X, Y) of this column
    double one0;                           // lane 0: start mass of the row-0 boundary; else 0
    unsigned trk;                          // closest approach of a high word to a threshold (sum3, APPROX == 2)
    double wj[R];                          // scratch register pairs of the threshold products (low word: don't care)
};

// Ring of haplotype bytes: two 256-byte slots filled by TMA; `pos` is the ring position
// (column + misalignment shift).
struct HapRing {
    uint8_t *ring;
    uint64_t *bar;
    const uint8_t *hsrc;  // 16-byte aligned source
    int lxs;              // window length + shift
    int n_chunks;
    uint32_t parity0, parity1;
    double *cols;         // literal variant: prob_cols of this warp (3 doubles per column)
    int last_lane;
};

// LUT kernels keep base CODES in the ring (A, C, G, T -> 0..3, anything else -> 4, case folded as
// to_ascii_uppercase does, pairhmm.rs:190): the chunk is converted in place right after it landed,
// 8 bytes per lane.
__device__ __forceinline__ uint32_t base_codes4(uint32_t w) {
    const uint32_t u = w & 0xDFDFDFDFu;
    return 0x04040404u - (__vcmpeq4(u, 0x41414141u) & 0x04040404u) - (__vcmpeq4(u, 0x43434343u) & 0x03030303u) -
           (__vcmpeq4(u, 0x47474747u) & 0x02020202u) - (__vcmpeq4(u, 0x54545454u) & 0x01010101u);
}
__device__ __forceinline__ void encode_chunk(uint8_t *slot, int lane) {
    uint2 *p = reinterpret_cast<uint2 *>(slot) + lane;
    uint2 v = *p;
    v.x = base_codes4(v.x);
    v.y = base_codes4(v.y);
    *p = v;
    __syncwarp();
}
__device__ __forceinline__ int base_code1(int b) {
    b &= 0xDF;
    return b == 'A' ? 0 : (b == 'C' ? 1 : (b == 'G' ? 2 : (b == 'T' ? 3 : 4)));
}

constexpr int kLutCols = 5;  // emission LUT: [row][code][lane]

// One anti-diagonal step (lane `lane` computes column t - lane).  LR >= 0: compile-time index of
// the last read row inside lane `last_lane` (free-end accumulation); LR < 0: run-time `last_r`.
// LUT = true (scaled linear space only): the ring holds base codes, the emission factor comes
// from the per-warp shared-memory table lut[(r*5 + code)*32] (already multiplied by t_mm: the M
// state carries that factor, which removes one multiplication per cell), and the row-0 start mass
// is added instead of selected.
template <int R, int LR, bool BANDED, int APPROX, bool EXT, bool LUT, typename A>
__device__ __forceinline__ void hmm_step(PairState<R> &s, const HmmConsts &c, const uint8_t *ring,
                                         const double *lut, int t, int lane, int shift, int band,
                                         int last_r) {
    const int i = t - lane;
    int xb = ring[(i + shift) & (kRing - 1)];
    if (!LUT) {
        if ((unsigned)(xb - 'a') < 26u) xb -= 32;  // to_ascii_uppercase (pairhmm.rs:190)
    }
    const double *lcol = lut + xb * 32;

    // receive the bottom row of the lane above (its column i).  The rotation makes lane 0 read
    // lane 31, which is always a padding lane (R = ceil(len_y / 31)) whose state is identically
    // zero: exactly the row-0 boundary fm[curr][0] = fx = fy = 0, with no select.
    const int src = (lane + 31) & 31;
    const double uM = __shfl_sync(0xffffffffu, s.bM, src);
    const double uX = __shfl_sync(0xffffffffu, s.bX, src);
    const double uY = __shfl_sync(0xffffffffu, s.bY, src);
    int uMed = kMedMax;
    if (BANDED) {
        uMed = __shfl_sync(0xffffffffu, s.bMed, src);
        if (lane == 0) uMed = kMedMax;  // med[curr][0] = MAX
    }
    double diagM = s.dM, diagX = s.dX, diagY = s.dY, upM = uM, upY = uY;
    int diagMed = s.dMed, upMed = uMed;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        bool match = false;
        double e;
        if (LUT) {
            e = lcol[r * (kLutCols * 32)];
            if (BANDED) match = xb == s.rb[r];
        } else {
            match = xb == s.rb[r];
            e = match ? s.pm[r] : s.pmm[r];
        }
        double nM = A::mul(e, A::template sum3<APPROX, EXT, LUT>(c, diagM, diagX, diagY, s.trk, s.wj[r]));
        double nX = A::template dot2<EXT>(LUT ? c.t_gy_s : c.t_gy, s.M[r], c.t_gye, s.X[r]);
        double nY;
        if (A::kLit)  // emit_y + ln_add_exp(gap_x + fm[curr][j], gap_x_ext + fy[curr][j]); cyo = ln eps
            nY = A::mul(s.cyo[r], A::add(A::mul(c.t_gx, upM), A::mul(c.t_gxe, upY)));
        else
            nY = A::template dot2<EXT>(s.cyo[r], upM, s.cye[r], upY);
        if (BANDED) {
            // PairHMM band: skip the cell if all three predecessors exceed max_edit_dist
            const int leftMed = s.med[r];
            const bool skip = min(diagMed, min(upMed, leftMed)) > band;
            int nMed = min(min(diagMed + (match ? 0 : 1), leftMed + 1), upMed + 1);
            nMed = min(nMed, kMedMax);
            if (skip) {
                nM = A::zero(); nX = A::zero(); nY = A::zero();
                nMed = kMedMax;
            }
            diagMed = leftMed;
            s.med[r] = nMed;
            upMed = nMed;
        }
        diagM = s.M[r]; diagX = s.X[r]; diagY = s.Y[r];
        s.M[r] = nM; s.X[r] = nX; s.Y[r] = nY;
        upM = nM; upY = nY;
    }
    // next column's diagonal inputs of my row 0 are this column's "up" values; lane 0 sees the
    // free-start mass fm[prev][0] = 1 there (uM is exactly zero for lane 0)
    if (A::kLog) s.dM = lane == 0 ? A::one() : uM;
    else s.dM = uM + s.one0;
    s.dX = uX;
    s.dY = uY;
    s.bM = s.M[R - 1]; s.bX = s.X[R - 1]; s.bY = s.Y[R - 1];
    if (BANDED) { s.dMed = lane == 0 ? 0 : uMed; s.bMed = s.med[R - 1]; }

    // free end: add the last read row of this column (prob_cols).  Only the accumulator of lane
    // `last_lane` is used; its column is always inside [0, len_x) or still all-zero (i < 0).
    if (LR >= 0) {
        if (LUT) {
            s.acc += s.M[LR >= 0 ? LR : 0];
            s.acc2 += s.X[LR >= 0 ? LR : 0] + s.Y[LR >= 0 ? LR : 0];
        } else {
            s.acc = A::add(s.acc, A::add(A::add(s.M[LR >= 0 ? LR : 0], s.X[LR >= 0 ? LR : 0]),
                                         s.Y[LR >= 0 ? LR : 0]));
        }
    } else if (A::kLit) {
        // prob_cols keeps the three values apart: the final ln_sum_exp runs over all of them
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (r == last_r) { s.acc = s.M[r]; s.acc2 = s.X[r]; s.acc3 = s.Y[r]; }
    } else {
        double tsum = A::zero(), tm = A::zero();
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (r == last_r) {
                if (LUT) { tm = s.M[r]; tsum = s.X[r] + s.Y[r]; }
                else tsum = A::add(A::add(s.M[r], s.X[r]), s.Y[r]);
            }
        if (LUT) { s.acc += tm; s.acc2 += tsum; }
        else s.acc = A::add(s.acc, tsum);
    }
}

// Sweep all columns of one pair.  Ring maintenance happens only at segment boundaries
// (ring position == 0 or 32 mod 256), the inner loop is branch-free.
template <int R, int LR, bool BANDED, int APPROX, bool EXT, bool LUT, typename A>
__device__ __forceinline__ void hmm_sweep(PairState<R> &s, const HmmConsts &c, HapRing &h,
                                          const double *lut, int T, int lane, int shift, int band,
                                          int last_r) {
    int t = 0;
    while (t < T) {
        const int ts = t + shift;
        const int ph = ts & (kChunk - 1);
        if (ph == 32) {  // every lane has left the previous chunk: refill its slot
            const int cnext = (ts >> 8) + 1;
            if (cnext < h.n_chunks && lane == 0) {
                const int off = cnext * kChunk;
                const uint32_t bytes = (uint32_t)min(kChunk, ((h.lxs - off) + 15) & ~15);
                uint64_t *b = &h.bar[cnext & 1];
                // the slot was read (and, with LUT, rewritten) through the generic proxy
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(b, bytes);
                tma_load_1d(h.ring + (cnext & 1) * kChunk, h.hsrc + off, bytes, b);
            }
        } else if (ph == 0 && t > 0) {  // lane 0 enters a new chunk: it must have landed
            const int cidx = ts >> 8;
            if (cidx < h.n_chunks) {
                if (cidx & 1) { mbar_wait(&h.bar[1], h.parity1); h.parity1 ^= 1; }
                else { mbar_wait(&h.bar[0], h.parity0); h.parity0 ^= 1; }
                if (LUT) encode_chunk(h.ring + (cidx & 1) * kChunk, lane);
            }
        }
        const int next_ts = ph < 32 ? (ts & ~(kChunk - 1)) + 32 : (ts & ~(kChunk - 1)) + kChunk;
        const int t_end = min(T, next_ts - shift);
#pragma unroll 2
        for (; t < t_end; ++t) {
            hmm_step<R, LR, BANDED, APPROX, EXT, LUT, A>(s, c, h.ring, lut, t, lane, shift, band, last_r);
            if (A::kLit && lane == h.last_lane && t >= lane) {
                double *col = h.cols + 3 * (int64_t)(t - lane);
                col[0] = s.acc; col[1] = s.acc2; col[2] = s.acc3;
            }
        }
    }
}

template <int R, int LR, bool BANDED, int APPROX, bool EXT, bool LUT, typename A>
struct SweepDispatch {
    __device__ __forceinline__ static void run(PairState<R> &s, const HmmConsts &c, HapRing &h,
                                               const double *lut, int T, int lane, int shift,
                                               int band, int last_r) {
        if (last_r == LR)
            hmm_sweep<R, LR, BANDED, APPROX, EXT, LUT, A>(s, c, h, lut, T, lane, shift, band, last_r);
        else
            SweepDispatch<R, LR - 1, BANDED, APPROX, EXT, LUT, A>::run(s, c, h, lut, T, lane, shift,
                                                                       band, last_r);
    }
};
template <int R, bool BANDED, int APPROX, bool EXT, bool LUT, typename A>
struct SweepDispatch<R, -1, BANDED, APPROX, EXT, LUT, A> {
    __device__ __forceinline__ static void run(PairState<R> &s, const HmmConsts &c, HapRing &h,
                                               const double *lut, int T, int lane, int shift,
                                               int band, int last_r) {
        hmm_sweep<R, -1, BANDED, APPROX, EXT, LUT, A>(s, c, h, lut, T, lane, shift, band, last_r);
    }
};

template <int R, bool BANDED, int APPROX, bool EXT, bool LUT, typename A>
__global__ void __launch_bounds__(kWarps * 32, Occ<R>::kMinBlocks) pairhmm_kernel(HmmArgs g) {
    static_assert(!LUT || !A::kLog, "the emission LUT exists in scaled linear space only");
    __shared__ __align__(128) uint8_t s_hap[kWarps][kRing];
    __shared__ __align__(8) uint64_t s_bar[kWarps][2];
    __shared__ __align__(16) double s_lut[kWarps][LUT ? R * kLutCols * 32 : 1];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    HapRing h;
    h.ring = s_hap[warp];
    h.bar = s_bar[warp];
    double *lut = s_lut[warp] + (LUT ? lane : 0);
    // LUT: ring bytes index the emission table, so the ring must hold valid codes (0..4) at every
    // position a lane may touch, including the wrap-around reads of columns outside [0, len_x)
    if (LUT) reinterpret_cast<uint4 *>(h.ring)[lane] = make_uint4(0, 0, 0, 0);
    if (lane == 0) {
        mbar_init(&h.bar[0], 1);
        mbar_init(&h.bar[1], 1);
        fence_mbar_init();
    }
    __syncwarp();
    h.parity0 = 0;
    h.parity1 = 0;

    const HmmConsts c = A::kLog ? g.ln : g.lin;
    const int n = *g.count;
    const int32_t *list = g.list ? g.list + (g.list_off ? *g.list_off : 0) : nullptr;

    for (;;) {
        int p = 0;
        if (lane == 0) p = atomicAdd(g.cursor, 1);
        p = __shfl_sync(0xffffffffu, p, 0);
        if (p >= n) break;
        const int pair = list ? list[p] : p;
        if (pair < 0) continue;  // computed by the band-only kernel (pairhmm_band.cuh)
        const int64_t hidx = g.pair_hap ? (int64_t)g.pair_hap[pair] : (int64_t)pair;
        const int64_t rd = g.pair_read ? (int64_t)g.pair_read[pair] : (int64_t)pair;
        const int64_t ho = g.win_start ? g.win_start[pair] : g.hap_off[hidx];
        const int lx = g.win_start ? g.win_len[pair] : (int)(g.hap_off[hidx + 1] - ho);
        const int64_t ro = g.read_off[rd];
        const int ly = (int)(g.read_off[rd + 1] - ro);
        int band = kMedMax;
        if (BANDED) {
            const int k = g.med[pair];
            band = k < 0 ? kMedMax : k;
        }
        if (lx <= 0 || ly <= 0) {
            // len_x == 0: ln_sum_exp of no columns = ln 0; len_y == 0 cannot occur in the reference
            if (lane == 0) g.out[pair] = -INFINITY;
            continue;
        }
        if (A::kLit && ly > 31 * R) continue;  // taken by the anti-diagonal kernel (pairhmm_diag.cuh)

        // ---- per-row constants (ReadEmission, pairhmm.rs:160-200)
        PairState<R> s;
        bool odd_base = false;  // a read base outside ACGT: the 5-column LUT cannot express it
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int j = lane * R + r;
            if (j < ly) {
                const int q = g.rqual[ro + j];
                const int b = g.rbase[ro + j];
                s.rb[r] = LUT ? base_code1(b) : b;
                if (LUT) odd_base |= s.rb[r] == 4;
                if (A::kLit) {
                    const double ln_eps = dm::mul((double)q, -0.23025850929940456);
                    s.pm[r] = g.qlut[q * kQlutStride + 3];
                    s.pmm[r] = dm::add(ln_eps, kLnConfusion);
                    s.cyo[r] = ln_eps;   // emit_y; the transition is added in hmm_step
                    s.cye[r] = ln_eps;
                } else if (A::kLog) {
                    const double ln_eps = (double)q * -0.23025850929940456;
                    s.pm[r] = g.qlut[q * kQlutStride + 3];
                    s.pmm[r] = ln_eps + kLnConfusion;  // ln(0.3333), pairhmm.rs:21-23
                    s.cyo[r] = ln_eps + c.t_gx;
                    s.cye[r] = ln_eps + c.t_gxe;
                } else {
                    const double eps = g.qlut[q * kQlutStride + 0];
                    s.pm[r] = g.qlut[q * kQlutStride + 1];
                    s.pmm[r] = g.qlut[q * kQlutStride + 2];
                    s.cyo[r] = eps * (LUT ? c.t_gx_s : c.t_gx);
                    s.cye[r] = eps * c.t_gxe;
                }
            } else {  // padding row below the read: stays at zero mass
                s.rb[r] = -1;
                s.pm[r] = s.pmm[r] = s.cyo[r] = s.cye[r] = A::zero();
            }
            s.M[r] = s.X[r] = s.Y[r] = A::zero();
            s.med[r] = kMedMax;
        }
        if (LUT) {
            if (__any_sync(0xffffffffu, odd_base)) {
                if (lane == 0) g.gen_list[atomicAdd(g.gen_count, 1)] = pair;
                continue;
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double em = s.pm[r] * c.t_mm, emm = s.pmm[r] * c.t_mm;
#pragma unroll
                for (int k = 0; k < kLutCols; ++k) lut[(r * kLutCols + k) * 32] = k == s.rb[r] ? em : emm;
            }
        }

        // ---- stage the first haplotype chunk (TMA).  Bulk copies need 16-byte aligned sources:
        // copy from the aligned-down address and shift the ring index by the misalignment.
        const int shift = (int)((uintptr_t)(g.hap + ho) & 15);
        h.hsrc = g.hap + ho - shift;
        h.lxs = lx + shift;
        h.n_chunks = (h.lxs + kChunk - 1) / kChunk;
        __syncwarp();  // every lane is done with the ring contents of the previous pair
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)min(kChunk, (h.lxs + 15) & ~15);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&h.bar[0], bytes);
            tma_load_1d(h.ring, h.hsrc, bytes, &h.bar[0]);
        }
        const int last_lane = (ly - 1) / R;
        const int last_r = (ly - 1) - last_lane * R;
        s.bM = s.bX = s.bY = A::zero();
        s.one0 = lane == 0 ? (LUT ? c.one_s : A::one()) : A::zero();
        s.dM = s.one0;
        s.dX = s.dY = A::zero();
        s.bMed = kMedMax;
        s.dMed = lane == 0 ? 0 : kMedMax;
        s.acc = A::zero();
        s.acc2 = A::zero();
        s.acc3 = A::zero();
        s.trk = 0xffffffffu;
#pragma unroll
        for (int r = 0; r < R; ++r) s.wj[r] = 0.0;
        h.last_lane = last_lane;
        if (A::kLit) {
            if (lx > g.lit_stride) {  // the caller sized the workspace for shorter windows
                if (lane == 0) g.out[pair] = NAN;
                mbar_wait(&h.bar[0], h.parity0);
                h.parity0 ^= 1;
                continue;
            }
            h.cols = g.lit_ws + ((int64_t)blockIdx.x * kWarps + warp) * 3 * g.lit_stride;
        }

        mbar_wait(&h.bar[0], h.parity0);
        h.parity0 ^= 1;
        if (LUT) encode_chunk(h.ring, lane);

        const int T = lx + last_lane;
        // the log-space instantiation keeps one loop body (run-time last row); the linear one is
        // specialised on the last-row index so that the free-end sum costs three adds per step
        SweepDispatch<R, A::kLog ? -1 : R - 1, BANDED, APPROX, EXT, LUT, A>::run(s, c, h, lut, T, lane,
                                                                                 shift, band, last_r);

        if (A::kLit) {
            // LogProb::ln_sum_exp(prob_cols): first maximum, then the sum of exp(p - max) over the
            // other finite entries in slice order, accumulated by ONE lane in that order
            __syncwarp();
            const int nc = 3 * lx;
            double pmax = -INFINITY;
            for (int k = lane; k < nc; k += 32) pmax = fmax(pmax, h.cols[k]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) pmax = fmax(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
            int imax = nc;
            for (int k = lane; k < nc; k += 32)
                if (h.cols[k] == pmax) { imax = k; break; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) imax = min(imax, __shfl_xor_sync(0xffffffffu, imax, o));
            double res = pmax;
            if (pmax != -INFINITY && pmax != INFINITY) {
                for (int k = lane; k < nc; k += 32) {
                    const double v = h.cols[k];
                    h.cols[k] = (k == imax || v == -INFINITY) ? 0.0 : dm::exp(dm::sub(v, pmax));
                }
                __syncwarp();
                if (lane == 0) {
                    double sum = 0.0;
                    for (int k = 0; k < nc; ++k) sum = dm::add(sum, h.cols[k]);  // + 0.0 is exact
                    res = dm::add(pmax, dm::log1p(sum));
                }
            }
            if (lane == 0) {
                if (res > 0.0) res = 0.0;
                g.out[pair] = res;
            }
            __syncwarp();
            continue;
        }
        if (APPROX == 2 && !A::kLog) {
            // a decision of the three-way rule came within reach of its threshold: the pair is
            // re-evaluated by the exact-decision kernel (generic list)
            if (__reduce_min_sync(0xffffffffu, s.trk) <= 8u) {
                if (lane == 0) g.gen_list[atomicAdd(g.gen_count, 1)] = pair;
                continue;
            }
        }
        // ---- result lives in lane last_lane
        double acc = __shfl_sync(0xffffffffu, s.acc, last_lane);
        if (LUT) acc = fma(acc, c.inv_tmm, __shfl_sync(0xffffffffu, s.acc2, last_lane));
        if (lane == 0) {
            double res;
            if (A::kLog) {
                res = acc;
            } else {
                // representable? (2^-940 of scaled mass; also catches inf/nan)
                if (!(acc >= 4.3e-283 && acc < 1.0e307)) {
                    const int slot = atomicAdd(g.fb_count, 1);
                    g.fb_list[slot] = pair;
                    res = NAN;
                } else {
                    res = log(acc) - c_scale_ln;
                }
            }
            if (res > 0.0) res = 0.0;  // "sum of paths can exceed 1" cap
            g.out[pair] = res;
        }
    }
}


// ---------------------------------------------------------------- long reads (config C4)
// Read windows longer than 248 rows are swept in horizontal strips of 31 lanes x 8 rows.  The
// bottom row (M, X, Y per column) of a strip goes to a per-warp global buffer and comes back as
// the row-above inputs of lane 0 in the next strip (prefetched 32 columns at a time through
// registers into shared memory).  In scaled linear space every strip is renormalised: the
// largest value of the boundary row is brought back to 2^960 by an exact power of two that is
// applied while the boundary chunk is loaded, and the exponents are summed (a 15 kb read loses
// ~e^-4000, far outside fp64 otherwise).
constexpr int kLongR = 8;
constexpr int kStripRows = 31 * kLongR;

template <int APPROX, bool EXT, bool LUT, typename A>
__global__ void __launch_bounds__(kWarps * 32, 2) pairhmm_long_kernel(HmmArgs g) {
    static_assert(!LUT || !A::kLog, "the emission LUT exists in scaled linear space only");
    constexpr int R = kLongR;
    __shared__ __align__(128) uint8_t s_hap[kWarps][kRing];
    __shared__ __align__(8) uint64_t s_bar[kWarps][2];
    __shared__ double s_bnd[kWarps][3][32];
    __shared__ __align__(16) double s_lut[kWarps][LUT ? R * kLutCols * 32 : 1];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    HapRing h;
    h.ring = s_hap[warp];
    h.bar = s_bar[warp];
    double *lut = s_lut[warp] + (LUT ? lane : 0);
    if (LUT) reinterpret_cast<uint4 *>(h.ring)[lane] = make_uint4(0, 0, 0, 0);  // valid codes everywhere
    if (lane == 0) {
        mbar_init(&h.bar[0], 1);
        mbar_init(&h.bar[1], 1);
        fence_mbar_init();
    }
    __syncwarp();
    h.parity0 = 0;
    h.parity1 = 0;
    const HmmConsts c = A::kLog ? g.ln : g.lin;
    const int n = *g.count;
    const int32_t *list = g.list ? g.list + (g.list_off ? *g.list_off : 0) : nullptr;
    double *ws = g.long_ws + ((int64_t)blockIdx.x * kWarps + warp) * 6 * g.long_stride;

    for (;;) {
        int p = 0;
        if (lane == 0) p = atomicAdd(g.cursor, 1);
        p = __shfl_sync(0xffffffffu, p, 0);
        if (p >= n) break;
        const int pair = list ? list[p] : p;
        const int64_t hidx = g.pair_hap ? (int64_t)g.pair_hap[pair] : (int64_t)pair;
        const int64_t rd = g.pair_read ? (int64_t)g.pair_read[pair] : (int64_t)pair;
        const int64_t ho = g.hap_off[hidx];
        const int lx = (int)(g.hap_off[hidx + 1] - ho);
        const int64_t ro = g.read_off[rd];
        const int ly = (int)(g.read_off[rd + 1] - ro);
        if (lx <= 0 || ly <= 0 || lx > g.long_stride) {
            if (lane == 0) g.out[pair] = lx > g.long_stride ? NAN : -INFINITY;
            continue;
        }
        if (LUT) {
            // a read base outside ACGT cannot be expressed by the 5-column LUT: such pairs (rare in
            // long reads) take the byte-compare log-space instantiation
            bool odd_base = false;
            for (int j = lane; j < ly; j += 32) odd_base |= base_code1(g.rbase[ro + j]) == 4;
            if (__any_sync(0xffffffffu, odd_base)) {
                if (lane == 0) {
                    g.fb_list[atomicAdd(g.fb_count, 1)] = pair;
                    g.out[pair] = NAN;
                }
                continue;
            }
        }
        const int shift = (int)((uintptr_t)(g.hap + ho) & 15);
        h.hsrc = g.hap + ho - shift;
        h.lxs = lx + shift;
        h.n_chunks = (h.lxs + kChunk - 1) / kChunk;
        const int n_strips = (ly + kStripRows - 1) / kStripRows;
        int64_t e_total = 0;   // sum of the per-strip rescaling exponents (linear space)
        int kshift = 0;        // exponent applied to the incoming boundary row
        bool dead = false;
        double acc = A::zero();
        int last_lane = 0;

        for (int strip = 0; strip < n_strips; ++strip) {
            const int j0 = strip * kStripRows;
            const int ns = min(kStripRows, ly - j0);
            last_lane = (ns - 1) / R;
            const int last_r = (ns - 1) - last_lane * R;
            const bool last_strip = strip == n_strips - 1;
            double *bin = ws + (int64_t)(strip & 1) * 3 * g.long_stride;
            double *bout = ws + (int64_t)((strip + 1) & 1) * 3 * g.long_stride;
            __syncwarp();
            if (lane == 0) {
                const uint32_t bytes = (uint32_t)min(kChunk, (h.lxs + 15) & ~15);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&h.bar[0], bytes);
                tma_load_1d(h.ring, h.hsrc, bytes, &h.bar[0]);
            }
            PairState<R> s;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int j = j0 + lane * R + r;
                if (lane * R + r < ns) {
                    const int q = g.rqual[ro + j];
                    const int b = g.rbase[ro + j];
                    s.rb[r] = LUT ? base_code1(b) : b;
                    if (A::kLog) {
                        const double ln_eps = (double)q * -0.23025850929940456;
                        s.pm[r] = g.qlut[q * kQlutStride + 3];
                        s.pmm[r] = ln_eps + -1.098712293668443;
                        s.cyo[r] = ln_eps + c.t_gx;
                        s.cye[r] = ln_eps + c.t_gxe;
                    } else {
                        const double eps = g.qlut[q * kQlutStride + 0];
                        s.pm[r] = g.qlut[q * kQlutStride + 1];
                        s.pmm[r] = g.qlut[q * kQlutStride + 2];
                        s.cyo[r] = eps * (LUT ? c.t_gx_s : c.t_gx);
                        s.cye[r] = eps * c.t_gxe;
                    }
                } else {
                    s.rb[r] = -1;
                    s.pm[r] = s.pmm[r] = s.cyo[r] = s.cye[r] = A::zero();
                }
                s.M[r] = s.X[r] = s.Y[r] = A::zero();
                s.med[r] = kMedMax;
            }
            if (LUT) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const double em = s.pm[r] * c.t_mm, emm = s.pmm[r] * c.t_mm;
#pragma unroll
                    for (int k = 0; k < kLutCols; ++k) lut[(r * kLutCols + k) * 32] = k == s.rb[r] ? em : emm;
                }
            }
            s.bM = s.bX = s.bY = A::zero();
            s.one0 = (lane == 0 && strip == 0) ? (LUT ? c.one_s : A::one()) : A::zero();
            s.dM = s.one0;
            s.dX = s.dY = A::zero();
            s.acc = A::zero();
            s.acc2 = A::zero();
            // largest boundary value of the strip: only its exponent is needed, so the running maximum
            // is kept on the high words (non-negative doubles order like their bit patterns)
            int bmax_hi = 0;
            // prefetch registers for the incoming boundary chunk (columns t .. t+31)
            double pfM = A::zero(), pfX = A::zero(), pfY = A::zero();
            if (strip > 0 && lane < lx) {
                pfM = __ldcg(&bin[lane]); pfX = __ldcg(&bin[g.long_stride + lane]); pfY = __ldcg(&bin[2 * g.long_stride + lane]);
            }
            mbar_wait(&h.bar[0], h.parity0);
            h.parity0 ^= 1;
            if (LUT) encode_chunk(h.ring, lane);
            const int T = lx + last_lane;
            const int src = (lane + 31) & 31;
            for (int t = 0; t < T; ++t) {
                // ---- haplotype ring maintenance (as hmm_sweep)
                const int ts = t + shift;
                const int ph = ts & (kChunk - 1);
                if (ph == 32) {
                    const int cnext = (ts >> 8) + 1;
                    if (cnext < h.n_chunks && lane == 0) {
                        const int off = cnext * kChunk;
                        const uint32_t bytes = (uint32_t)min(kChunk, ((h.lxs - off) + 15) & ~15);
                        uint64_t *b = &h.bar[cnext & 1];
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        mbar_expect_tx(b, bytes);
                        tma_load_1d(h.ring + (cnext & 1) * kChunk, h.hsrc + off, bytes, b);
                    }
                } else if (ph == 0 && t > 0) {
                    const int cidx = ts >> 8;
                    if (cidx < h.n_chunks) {
                        if (cidx & 1) { mbar_wait(&h.bar[1], h.parity1); h.parity1 ^= 1; }
                        else { mbar_wait(&h.bar[0], h.parity0); h.parity0 ^= 1; }
                        if (LUT) encode_chunk(h.ring + (cidx & 1) * kChunk, lane);
                    }
                }
                // ---- incoming boundary chunk: registers -> shared, prefetch the next one
                if (strip > 0 && (t & 31) == 0) {
                    __syncwarp();
                    double vM = pfM, vX = pfX, vY = pfY;
                    if (!A::kLog && kshift != 0) { vM = scalbn(vM, kshift); vX = scalbn(vX, kshift); vY = scalbn(vY, kshift); }
                    s_bnd[warp][0][lane] = vM; s_bnd[warp][1][lane] = vX; s_bnd[warp][2][lane] = vY;
                    const int cn = t + 32 + lane;
                    pfM = pfX = pfY = A::zero();
                    if (cn < lx) {
                        pfM = __ldcg(&bin[cn]); pfX = __ldcg(&bin[g.long_stride + cn]); pfY = __ldcg(&bin[2 * g.long_stride + cn]);
                    }
                    __syncwarp();
                }
                const int i = t - lane;
                int xb = h.ring[(i + shift) & (kRing - 1)];
                if (!LUT) {
                    if ((unsigned)(xb - 'a') < 26u) xb -= 32;
                }
                const double *lcol = lut + xb * 32;
                double uM = __shfl_sync(0xffffffffu, s.bM, src);
                double uX = __shfl_sync(0xffffffffu, s.bX, src);
                double uY = __shfl_sync(0xffffffffu, s.bY, src);
                if (strip > 0 && lane == 0) {  // row above = bottom row of the previous strip
                    const bool in = t < lx;
                    uM = in ? s_bnd[warp][0][t & 31] : A::zero();
                    uX = in ? s_bnd[warp][1][t & 31] : A::zero();
                    uY = in ? s_bnd[warp][2][t & 31] : A::zero();
                }
                double diagM = s.dM, diagX = s.dX, diagY = s.dY, upM = uM, upY = uY;
                unsigned trk_unused = 0;
                double wj_unused = 0.0;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    double e;
                    if (LUT) e = lcol[r * (kLutCols * 32)];
                    else e = xb == s.rb[r] ? s.pm[r] : s.pmm[r];
                    // (three-way rule: a long pair has ~1e8 cells, so a doubtful decision is settled cell by
                    // cell after a warp vote instead of re-evaluating the pair)
                    const double nM = A::mul(e, A::template sum3<APPROX == 2 ? 4 : APPROX, EXT, LUT>(c, diagM, diagX, diagY, trk_unused, wj_unused));
                    const double nX = A::template dot2<EXT>(LUT ? c.t_gy_s : c.t_gy, s.M[r], c.t_gye, s.X[r]);
                    const double nY = A::template dot2<EXT>(s.cyo[r], upM, s.cye[r], upY);
                    diagM = s.M[r]; diagX = s.X[r]; diagY = s.Y[r];
                    s.M[r] = nM; s.X[r] = nX; s.Y[r] = nY;
                    upM = nM; upY = nY;
                }
                if (A::kLog) s.dM = (lane == 0 && strip == 0) ? A::one() : uM;
                else s.dM = uM + s.one0;
                s.dX = uX;
                s.dY = uY;
                s.bM = s.M[R - 1]; s.bX = s.X[R - 1]; s.bY = s.Y[R - 1];
                if (last_strip) {
                    double tsum = A::zero(), tm = A::zero();
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (r == last_r) {
                            if (LUT) { tm = s.M[r]; tsum = s.X[r] + s.Y[r]; }
                            else tsum = A::add(A::add(s.M[r], s.X[r]), s.Y[r]);
                        }
                    if (LUT) { s.acc += tm; s.acc2 += tsum; }
                    else s.acc = A::add(s.acc, tsum);
                } else if (lane == 30 && (unsigned)i < (unsigned)lx) {
                    // full strip: lane 30 holds its last row; hand it to the next strip
                    __stcg(&bout[i], s.bM); __stcg(&bout[g.long_stride + i], s.bX); __stcg(&bout[2 * g.long_stride + i], s.bY);
                    if (!A::kLog) bmax_hi = max(bmax_hi, max(__double2hiint(s.bM), max(__double2hiint(s.bX), __double2hiint(s.bY))));
                }
            }
            if (last_strip) {
                acc = __shfl_sync(0xffffffffu, s.acc, last_lane);
                if (LUT) acc = fma(acc, c.inv_tmm, __shfl_sync(0xffffffffu, s.acc2, last_lane));
            } else if (!A::kLog) {
                // renormalise: bring the largest boundary value back to 2^960
                bmax_hi = __shfl_sync(0xffffffffu, bmax_hi, 30);
                const int bexp = (bmax_hi >> 20) & 0x7ff;   // biased exponent of the maximum
                // zero / subnormal, above 1e307 (2^1020), Inf or NaN (a NaN state has sign 0 here or
                // shows up as a negative high word): the pair takes the log-space pass
                if (bmax_hi < 0 || bexp == 0 || bexp >= 1023 + 1019) { dead = true; break; }
                const int k = 960 - (bexp - 1023);
                kshift = k;
                e_total += k;
            }
            __threadfence_block();
        }
        if (lane == 0) {
            double res;
            if (A::kLog) {
                res = acc;
            } else if (dead || !(acc >= 4.3e-283 && acc < 1.0e307)) {
                const int slot = atomicAdd(g.fb_count, 1);
                g.fb_list[slot] = pair;
                res = NAN;
            } else {
                res = log(acc) - c_scale_ln - (double)e_total * 0.6931471805599453;
            }
            if (res > 0.0) res = 0.0;
            g.out[pair] = res;
        }
    }
}

// ---------------------------------------------------------------- launch table (pairhmm_inst_*.cu)
// cls: 0..6 -> R = 1,2,3,4,5,6,8
template <bool BANDED, int APPROX>
cudaError_t launch_pairhmm_lin(int cls, bool ext, const HmmArgs &a, int sm_count, cudaStream_t st);
template <bool BANDED, int APPROX>
cudaError_t launch_pairhmm_log(const HmmArgs &a, int sm_count, cudaStream_t st);
// scaled linear space without the emission LUT (any byte alphabet), R = 8, gap extension on
template <bool BANDED, int APPROX>
cudaError_t launch_pairhmm_generic(const HmmArgs &a, int sm_count, cudaStream_t st);
// long reads (unbanded only): linear pass and its log-space fallback
template <int APPROX>
cudaError_t launch_pairhmm_long(bool ext, bool log_space, const HmmArgs &a, int sm_count, cudaStream_t st);
constexpr int kLongBlocksPerSm = 2;
// literal log-space variant (LogLit): banded or not (a negative band entry = unbanded), any alphabet
template <int APPROX>
cudaError_t launch_pairhmm_literal(const HmmArgs &a, int sm_count, cudaStream_t st);
constexpr int kLitBlocksPerSm = 2;
// persistent-grid size per SM; VLR_K1_BLOCKS_PER_SM overrides it (tuning aid)
inline int blocks_per_sm(int dflt) {
    const char *e = getenv("VLR_K1_BLOCKS_PER_SM");
    const int v = e ? atoi(e) : 0;
    return v > 0 ? v : dflt;
}

#define VLR_PAIRHMM_LONG_INSTANTIATE(APPROX)                                                      \
    template <>                                                                                   \
    cudaError_t launch_pairhmm_long<APPROX>(bool ext, bool log_space, const HmmArgs &a,           \
                                            int sm_count, cudaStream_t st) {                      \
        const int grid = sm_count * kLongBlocksPerSm;                                             \
        if (log_space) pairhmm_long_kernel<APPROX, true, false, Log><<<grid, kWarps * 32, 0, st>>>(a); \
        else if (ext) pairhmm_long_kernel<APPROX, true, true, Lin><<<grid, kWarps * 32, 0, st>>>(a);   \
        else pairhmm_long_kernel<APPROX, false, true, Lin><<<grid, kWarps * 32, 0, st>>>(a);           \
        return cudaGetLastError();                                                                \
    }

#define VLR_PAIRHMM_LITERAL_INSTANTIATE(APPROX)                                                   \
    template <>                                                                                   \
    cudaError_t launch_pairhmm_literal<APPROX>(const HmmArgs &a, int sm_count, cudaStream_t st) { \
        pairhmm_kernel<8, true, APPROX, true, false, LogLit>                                      \
            <<<sm_count * kLitBlocksPerSm, kWarps * 32, 0, st>>>(a);                              \
        return cudaGetLastError();                                                                \
    }

#define VLR_PAIRHMM_INSTANTIATE(BANDED, APPROX)                                                   \
    template <int R, bool EXT>                                                                    \
    static cudaError_t launch_r(const HmmArgs &a, int sm_count, cudaStream_t st) {                \
        pairhmm_kernel<R, BANDED, APPROX, EXT, true, Lin>                                         \
            <<<sm_count * blocks_per_sm(Occ<R>::kMinBlocks), kWarps * 32, 0, st>>>(a);            \
        return cudaGetLastError();                                                                \
    }                                                                                             \
    template <>                                                                                   \
    cudaError_t launch_pairhmm_lin<BANDED, APPROX>(int cls, bool ext, const HmmArgs &a,           \
                                                   int sm_count, cudaStream_t st) {               \
        switch (cls) {                                                                            \
        case 0: return ext ? launch_r<1, true>(a, sm_count, st) : launch_r<1, false>(a, sm_count, st); \
        case 1: return ext ? launch_r<2, true>(a, sm_count, st) : launch_r<2, false>(a, sm_count, st); \
        case 2: return ext ? launch_r<3, true>(a, sm_count, st) : launch_r<3, false>(a, sm_count, st); \
        case 3: return ext ? launch_r<4, true>(a, sm_count, st) : launch_r<4, false>(a, sm_count, st); \
        case 4: return ext ? launch_r<5, true>(a, sm_count, st) : launch_r<5, false>(a, sm_count, st); \
        case 5: return ext ? launch_r<6, true>(a, sm_count, st) : launch_r<6, false>(a, sm_count, st); \
        case 6: return ext ? launch_r<8, true>(a, sm_count, st) : launch_r<8, false>(a, sm_count, st); \
        default: return cudaErrorInvalidValue;                                                    \
        }                                                                                         \
    }                                                                                             \
    template <>                                                                                   \
    cudaError_t launch_pairhmm_generic<BANDED, APPROX>(const HmmArgs &a, int sm_count,            \
                                                       cudaStream_t st) {                         \
        pairhmm_kernel<8, BANDED, (APPROX == 2 ? 3 : APPROX), true, false, Lin>                   \
            <<<sm_count * Occ<8>::kMinBlocks, kWarps * 32, 0, st>>>(a);                           \
        return cudaGetLastError();                                                                \
    }                                                                                             \
    template <>                                                                                   \
    cudaError_t launch_pairhmm_log<BANDED, APPROX>(const HmmArgs &a, int sm_count,                \
                                                   cudaStream_t st) {                             \
        pairhmm_kernel<8, BANDED, APPROX, true, false, Log>                                       \
            <<<sm_count * Occ<8>::kMinBlocks, kWarps * 32, 0, st>>>(a);                           \
        return cudaGetLastError();                                                                \
    }

}  // namespace vlr
