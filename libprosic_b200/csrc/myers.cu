// myers.cu -- K2: Myers bit-vector best hit with traceback, and the fused allele-support pipeline.
//
// Replaces EditDistanceCalculation::calc_best_hit (reference
// src/variants/evidence/realignment/edit_distance.rs:56-154, on top of
// bio::pattern_matching::myers) and, in vlr_allele_support_batch, the per-region part of
// Realigner::allele_support / prob_allele (realignment/mod.rs:225-319, 338-385):
//   best hit per (read window, haplotype) -> minimal-distance haplotypes per allele ->
//   dist == 0 ? certainty_est : banded pair-HMM on the window shrunk to the hit (pairhmm.rs:38-44)
//   -> max over haplotypes -> normalisation / 0.5 fallbacks.
//
// One thread per (read, haplotype) pair: the pattern is the read window (KW words of 64 bits: 256
// rows in the common instantiation, 1024 in the one for long windows -- bio's Myers::Long,
// edit_distance.rs:42-46), the text the haplotype window.  Two scans:
//   1. the whole text without storing anything: best distance d*, number and range of the best end
//      positions.  Exact occurrences (d* = 0) are finished here (start = end - m, no operations).
//   2. only the columns a traceback can look at, c0 = first_best - m - 2 d* - 1 .. last_best, this time
//      keeping the vertical-delta bit-vectors (Pv, Mv) of every column in a private global scratch
//      slot, so that the traceback can evaluate any D[j][i] with two masked popcounts per word.
//      Starting at c0 instead of 0 is exact: a traceback from (m, e) stays on the diagonals
//      e - m +- d* and consults their neighbours; a consulted cell only matters if its D is at most
//      d*, and an alignment of cost <= d* ending there starts at or after c0 (the truncated scan's D'
//      is never below D, so every "predecessor + cost == current" test has the same outcome).
// This cut the scratch traffic of the C3 read set from 18 GB to under 6 GB per launch (the kernel was
// bound by it).  Integer work only; results are bit-exact against the oracle.
#include <math.h>
#include <string.h>

#include <type_traits>
#include <vector>

#include <stdlib.h>

#include <chrono>

#include "pairhmm_host.cuh"
#include "detmath.cuh"

namespace vlr {

constexpr int kMaxIndelOps = VLR_MAX_INDEL_OPS;   // bytes per output record: one byte per run
constexpr int kMaxRuns = kMaxIndelOps;
constexpr int kMaxRunLen = 127;

struct HitOut {
    int32_t dist, start, end, n_best, n_indel;
};

struct MyersArgs {
    const uint8_t *hap;
    const int64_t *hap_off;
    const uint8_t *rbase;
    const int64_t *read_off;
    const int32_t *pair_hap, *pair_read;
    int64_t n_pairs;
    int32_t *cursor;
    uint64_t *scratch;        // per worker: (max_lx + 1) * 4 * KW words
    int32_t m_lo, m_hi;       // this launch takes the pairs with m_lo < rows <= m_hi
    int32_t exp_skip;         // experiments (VLR_K2_EXP): 1 = no traceback, 2 = no second scan either
    int32_t *scratch_dm;      // per worker: (max_lx + 1) bottom distances
    int64_t scratch_stride;   // columns per worker
    HitOut *out;
    uint8_t *out_ops;         // [n_pairs][kMaxIndelOps]
    // PathHMM ("fast" realignment mode, realignment/mod.rs:499-576): probability of the best
    // alignment path(s); needs the base qualities and the transition table (nullable)
    const uint8_t *rqual;
    const double *qlut;
    double *out_path;
    double no_gap, close_x, close_y, reopen_x, reopen_y, gap_x, gap_y;
};

__device__ __forceinline__ int up_char(int c) { return (unsigned)(c - 'a') < 26u ? c - 32 : c; }
__device__ __forceinline__ int sym_of(int c) {  // A C G T -> a code in 0..3, N -> 4, anything else 5
    const unsigned t = (unsigned)(c - 'A');
    const bool acgt = t < 20u && ((0x80045u >> t) & 1u);   // bits 0 (A), 2 (C), 6 (G), 19 (T)
    return acgt ? (c >> 1) & 3 : (c == 'N' ? 4 : 5);       // A 0, C 1, T 2, G 3
}

// D[j][i] from the stored column state (i >= 1): bottom distance minus the vertical deltas of
// rows j+1..m
// D[j][i] from a column record (W words of Pv, then W words of Mv) and the bottom distance
template <int kMyW>
__device__ __forceinline__ int dist_from(const uint64_t *rec, int dm, int j, int m, int W) {
    int d = dm;
#pragma unroll
    for (int w = 0; w < kMyW; ++w) {
        const int lo = w * 64, hi = min(m, lo + 64);
        if (w >= W || j >= hi) continue;
        const int from = max(j, lo) - lo;  // bit index of row from+1
        uint64_t mask = (hi - lo == 64 ? ~0ull : ((1ull << (hi - lo)) - 1ull)) & (~0ull << from);
        d -= __popcll(rec[w] & mask);
        d += __popcll(rec[kMyW + w] & mask);
    }
    return d;
}

struct WalkRes {   // result of one traceback, handed from the executing lane to the owner of the pair
    int32_t start, nindel, nruns, e;
    double path;
    uint8_t runs[kMaxRuns];
};

template <int kMyW>
// CTAs per SM the register allocation must allow.  More resident threads do NOT help this kernel (5 or 6
// CTAs at 96 / 80 registers: 4.2 / 4.1 ms against 3.7 ms on the C3 read set -- the scratch working set
// grows with them); no cap at all (125 registers, nothing spilled) is the fastest.
#ifndef VLR_K2_MINB
#define VLR_K2_MINB 1
#endif
__global__ void __launch_bounds__(128, VLR_K2_MINB) besthit_kernel(MyersArgs g) {
    __shared__ int2 s_task[4][32];
    __shared__ WalkRes s_res[4][32];
    const int64_t worker = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    // Scratch is interleaved across the lanes of a warp: word w of column c of lane l is one 32-byte
    // sector (Pv, Mv, Ph, Mh) at ((c * kMyW + w) * 32 + l) * 32 bytes, the bottom distance sits at
    // (c * 32 + l) * 4.  The lanes of a warp walk the columns in lock step, so a column's stores are whole
    // 1 KB rows, and a traceback step fetches exactly one sector.
    const int64_t warp_id = worker >> 5;
    const int lane_id = (int)(worker & 31);
    ulonglong2 *cols = reinterpret_cast<ulonglong2 *>(g.scratch) + warp_id * g.scratch_stride * (2 * kMyW * 32) + lane_id * 2;
    int32_t *dms_base = g.scratch_dm + warp_id * g.scratch_stride * 32 + lane_id;
    auto dms = [&](int c) -> int32_t & { return dms_base[(int64_t)c * 32]; };
    const bool want_path = g.out_path != nullptr;
    // A warp takes 32 consecutive pairs at a time and stays converged: every loop below that nests
    // data-dependent work (the tracebacks of several best end positions) runs warp-uniformly with
    // the lanes' own state predicated -- as a plain "for e { while (j > 0) }" per thread the lanes
    // drifted apart and the traceback ran with 5 of 32 lanes active.
    for (;;) {
        int base = 0;
        if (lane_id == 0) base = atomicAdd(g.cursor, 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if ((int64_t)base >= g.n_pairs) break;
        const int64_t k = (int64_t)base + lane_id;
        const bool valid = k < g.n_pairs;
        const int64_t h = !valid ? 0 : (g.pair_hap ? g.pair_hap[k] : k);
        const int64_t rd = !valid ? 0 : (g.pair_read ? g.pair_read[k] : k);
        const int64_t ho = g.hap_off[h], ro = g.read_off[rd];
        const int n = (int)(g.hap_off[h + 1] - ho), m = (int)(g.read_off[rd + 1] - ro);
        HitOut res = {0, 0, 0, 0, 0};
        // this instantiation's pair?  (rows in (m_lo, m_hi]; degenerate pairs belong to the first launch)
        const bool mine = valid && m <= g.m_hi && !(m <= g.m_lo && m > 0 && n > 0);
        const bool work = mine && n > 0 && m > 0;
        if (mine && !work) {
            res.dist = -1;  // no hit (calc_best_hit returns None for an empty text)
            g.out[k] = res;
            if (g.out_path) g.out_path[k] = NAN;
        }
        const uint8_t *pat = g.rbase + ro, *txt = g.hap + ho;
        // pattern masks for A C G T N
        uint64_t peq[5][kMyW];
#pragma unroll
        for (int s = 0; s < 5; ++s)
#pragma unroll
            for (int w = 0; w < kMyW; ++w) peq[s][w] = 0;
        if (work)
            for (int j = 0; j < m; ++j) {
                const int s = sym_of(pat[j]);
                if (s < 5) peq[s][j >> 6] |= 1ull << (j & 63);
            }
        const int W = (m + 63) >> 6, lastbit = (m - 1) & 63;
        int best = 1 << 30, n_best = 0, first_best = 0, last_best = 0;
        const int toff = (int)(reinterpret_cast<uintptr_t>(txt) & 7);
        const uint64_t *t8 = reinterpret_cast<const uint64_t *>(txt - toff);
        // One scan of the text columns [i0, i1) as if the text began at i0.  STORE: keep the column
        // records (second scan); else track the best bottom distance (first scan).
        auto scan = [&](int i0, int i1, auto store_tag) {
            constexpr bool STORE = decltype(store_tag)::value;
            uint64_t Pv[kMyW], Mv[kMyW];
#pragma unroll
            for (int w = 0; w < kMyW; ++w) { Pv[w] = ~0ull; Mv[w] = 0; }
            int score = m;
            // text characters arrive in aligned 8-byte chunks, one chunk ahead of their use (the
            // staged haplotype buffer carries slack on both sides of every window)
            uint64_t tw = 0, tw_next = 0;
            if (i0 < i1) { tw = t8[(i0 + toff) >> 3]; tw_next = t8[((i0 + toff) >> 3) + 1]; }
            for (int i = i0; i < i1; ++i) {
                const int tq = i + toff;
                if ((tq & 7) == 0 && i > i0) { tw = tw_next; tw_next = t8[(tq >> 3) + 1]; }
                const int c = up_char((int)((tw >> ((tq & 7) * 8)) & 0xffu));
                const int s = sym_of(c);
                int hin = 0;
                uint64_t PhS[STORE ? kMyW : 1], MhS[STORE ? kMyW : 1];
#pragma unroll
                for (int w = 0; w < kMyW; ++w) {
                    if (w >= W) break;
                    uint64_t Eq;
                    if (s < 5) Eq = peq[s][w];
                    else {  // rare symbol: build the mask on the fly
                        Eq = 0;
                        for (int j = w * 64; j < min(m, w * 64 + 64); ++j)
                            if (pat[j] == c) Eq |= 1ull << (j & 63);
                    }
                    const uint64_t pv = Pv[w], mv = Mv[w];
                    const uint64_t Xv = Eq | mv;
                    if (hin < 0) Eq |= 1ull;
                    const uint64_t Xh = (((Eq & pv) + pv) ^ pv) | Eq;
                    uint64_t Ph = mv | ~(Xh | pv);
                    uint64_t Mh = pv & Xh;
                    const int top = (w == W - 1) ? lastbit : 63;
                    const int hout = (int)((Ph >> top) & 1ull) - (int)((Mh >> top) & 1ull);
                    if (STORE) { PhS[w] = Ph; MhS[w] = Mh; }  // horizontal deltas of this column's rows
                    Ph <<= 1; Mh <<= 1;
                    if (hin < 0) Mh |= 1ull; else if (hin > 0) Ph |= 1ull;
                    Pv[w] = Mh | ~(Xv | Ph);
                    Mv[w] = Ph & Xv;
                    hin = hout;
                }
                score += hin;
                if (STORE) {
                    // column record: the words the pattern uses, at the column's index RELATIVE to the scan
                    // start -- the lanes of a warp then write the same scratch row in the same iteration
                    // although their windows start at different columns
                    ulonglong2 *col = cols + (int64_t)(i + 1 - i0) * (2 * kMyW * 32);
                    // Only the words a traceback can consult are written.  A walk from (m, e) stays on the
                    // diagonals (e - m) +- d*; at column c it reads the rows j and j - 1 of that band, and
                    // (one and two columns ahead of its position) row j of the columns c - 1 and c - 2:
                    // rows [c - (last_best - m) - d* - 1, c - (first_best - m) + d* + 2], one more on
                    // either side for good measure.  With a unique best end and a few edits that is one
                    // word of the three a 150-row read has (5.6 -> 2 GB of scratch writes per launch on the
                    // C3 read set).
                    const int cdp = i + 1;
                    const int w_lo = max(0, (cdp - (last_best - m) - best - 3) >> 6);
                    const int w_hi = (cdp - (first_best - m) + best + 2) >> 6;
#pragma unroll
                    for (int w = 0; w < kMyW; ++w)
                        if (w < W && w >= w_lo && w <= w_hi) {
                            col[w * 64] = make_ulonglong2(Pv[w], Mv[w]);          // vertical deltas
                            col[w * 64 + 1] = make_ulonglong2(PhS[w], MhS[w]);    // horizontal deltas
                        }
                    dms(i + 1 - i0) = score;  // bottom distance
                } else {
                    if (score < best) { best = score; n_best = 1; first_best = i + 1; last_best = i + 1; }
                    else if (score == best) { n_best++; last_best = i + 1; }
                }
            }
        };
        scan(0, work ? n : 0, std::false_type());
        // the columns a traceback can consult (see the header); exact occurrences need none
        const int c0 = work ? max(0, first_best - m - 2 * best - 1) : 0;
        const bool walks = work && (best > 0 || want_path) && g.exp_skip == 0;
        scan(c0, (work && (best > 0 || want_path) && g.exp_skip < 2) ? last_best : c0, std::true_type());
        // tracebacks of every best end position (edit_distance.rs:88-97), in increasing position
        int start_first = 0, start_last = 0, best_nindel = 1 << 30, best_nops = 0;
        uint8_t best_runs[kMaxRuns];  // best_indel_operations, run-length encoded (Ins << 7 | length), forward order
        double best_path = -INFINITY;
        bool have_path = false;
        if (work && best == 0 && !want_path) {
            // exact occurrences: D[m][e] = 0 forces the diagonal at every step, so the alignments start
            // at e - m and have no indel operation -- no second scan, no walk (41 % of the pairs of the
            // C3 read set)
            start_first = first_best - m;
            start_last = last_best - m;
            best_nindel = 0;
            best_nops = 0;
        }
        // ---- tracebacks, balanced over the lanes of the warp.  A lane OWNS a pair (its scratch
        // records, its result), but any lane can EXECUTE the walk from one of its best end positions:
        // pairs differ in the number of walks they need (none for exact occurrences, several for ends
        // that tie), and with one lane walking all of its own pair's alignments the warp waited for the
        // worst lane (6 of 32 lanes active on the C3 read set).  Per round the owners offer their next
        // best end positions, a prefix sum packs up to 32 of them into a task list in shared memory,
        // lane t walks task t with the owner's context fetched by shuffles, and the owners then fold
        // the results of their tasks in increasing end position.
        const int wib = (int)(threadIdx.x >> 5);
        int remaining = walks ? n_best : 0;
        int e_cursor = first_best - 1;
        for (;;) {
            const int offer = min(remaining, 8);
            int incl = offer;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane_id >= o) incl += v;
            }
            const int excl = incl - offer;
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (total == 0) break;  // warp-uniform
            const int acc = max(0, min(offer, 32 - excl));
            for (int t = 0; t < acc; ++t) {
                ++e_cursor;
                while (dms(e_cursor - c0) != best) ++e_cursor;
                s_task[wib][excl + t] = make_int2(lane_id, e_cursor);
            }
            remaining -= acc;
            const int n_tasks = min(total, 32);
            __syncwarp();
            // ---- executor side: the owner's context
            const bool has = lane_id < n_tasks;
            const int2 tk = has ? s_task[wib][lane_id] : make_int2(lane_id, 0);
            const int owner = tk.x, e = tk.y;
            const int t_m = __shfl_sync(0xffffffffu, m, owner);
            const int t_c0 = __shfl_sync(0xffffffffu, c0, owner);
            const int t_best = __shfl_sync(0xffffffffu, best, owner);
            const long long t_ro = __shfl_sync(0xffffffffu, (long long)ro, owner);
            const long long t_ho = __shfl_sync(0xffffffffu, (long long)ho, owner);
            const ulonglong2 *t_cols = cols + (owner - lane_id) * 2;
            const uint8_t *t_pat = g.rbase + t_ro, *t_txt = g.hap + t_ho;
            const int t_toff = (int)(reinterpret_cast<uintptr_t>(t_txt) & 7);
            const uint64_t *t_t8 = reinterpret_cast<const uint64_t *>(t_txt - t_toff);
            const int poff = (int)(reinterpret_cast<uintptr_t>(t_pat) & 7);
            const uint64_t *p8 = reinterpret_cast<const uint64_t *>(t_pat - poff);
            // indel ops in traceback (reverse) order, run-length encoded in one byte per run: bit 7 = Ins
            // (else Del), bits 0-6 = length (<= 127, longer runs continue in the next byte).  An
            // alignment has a handful of indel runs however long they are (a 40-base insertion is one);
            // more than kMaxRuns runs leave the tail uncounted, which the host can tell from the count.
            uint8_t runs[kMaxRuns];
            int nruns = 0, nindel = 0;
            auto push_op = [&](int o) {
                const int tag = o == 3 ? 0x80 : 0;
                if (nruns > 0 && (runs[nruns - 1] & 0x80) == tag && (runs[nruns - 1] & 0x7f) < kMaxRunLen) runs[nruns - 1]++;
                else if (nruns < kMaxRuns) runs[nruns++] = (uint8_t)(tag | 1);
            };
            // PathHMM: the walk runs backwards, so the transition INTO the operation found last
            // step (`next`) is added once its predecessor is known; emissions as found
            double path = 0.0;
            int next = -1;  // 0 match/subst, 2 del, 3 ins
            auto trans = [&](int prev, int cur) -> double {
                if (cur == 0) return prev == 2 ? g.close_y : (prev == 3 ? g.close_x : (prev == 0 ? g.no_gap : 0.0));
                if (cur == 2) return prev == 2 ? g.reopen_y : (prev == 3 ? g.close_x + g.gap_y : g.gap_y);
                return prev == 3 ? g.reopen_x : (prev == 2 ? g.close_y + g.gap_x : g.gap_x);
            };
            // D[j][i] of the current cell is carried along; its three predecessors follow from single
            // bits of the column's record: the vertical delta of row j (Pv / Mv, bit j-1) gives
            // D[j-1][i], the horizontal delta of row j (Ph / Mh) gives D[j][i-1], the horizontal delta of
            // row j-1 then D[j-1][i-1] -- no popcounts, one 64-bit word of each vector per step.
            // Tie-breaking among equally good predecessors (bio's Myers traceback, recollected and
            // confirmed by the reference's integration tests): a substitution
            // first, then Ins, then Del, a match last.
            // The words of column i (the one holding row j) are in registers, those of column i-1 are
            // loaded one step ahead, and the sectors of the columns further left are pulled into L1
            // four steps ahead (a prefetch has no destination register, so nothing waits for it): the
            // walk moves left by at most one column per step, so the scratch latency stays off the
            // dependent chain.
            struct Rec { uint64_t pv, mv, ph, mh; };
            auto load_rec = [&](int c, int w) {
                Rec r;
                const int64_t at = ((int64_t)max(c - t_c0, 1) * kMyW + w) * 64;
                const ulonglong2 v = t_cols[at], hh = t_cols[at + 1];
                r.pv = v.x; r.mv = v.y; r.ph = hh.x; r.mh = hh.y;
                return r;
            };
            auto prefetch_rec = [&](int c, int w) {
                const ulonglong2 *ptr = &t_cols[((int64_t)max(c - t_c0, 1) * kMyW + w) * 64];
                asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr));
            };
            int pchunk = -1, tchunk = -(1 << 28);   // pattern / text bytes in aligned 8-byte chunks
            uint64_t pw = 0, twd = 0;
            Rec cur = {0, 0, 0, 0}, left = {0, 0, 0, 0};
            int j = 0, i = 0, d = 0, wcur = 0;
            bool walking = has;
            if (has) {
                j = t_m; i = e; d = t_best;
                wcur = (j - 1) >> 6;
                prefetch_rec(i - 2, wcur); prefetch_rec(i - 3, wcur); prefetch_rec(i - 4, wcur);
                cur = load_rec(i, wcur);
                left = load_rec(i - 1, wcur);
            }
            while (__any_sync(0xffffffffu, walking)) {
                if (walking) {
                    const int rbit = j - 1;
                    if ((rbit >> 6) != wcur) {  // the row left the word in registers (once per 64 rows)
                        wcur = rbit >> 6;
                        cur = load_rec(i, wcur);
                        left = load_rec(i - 1, wcur);
                    }
                    prefetch_rec(i - 5, wcur);
                    const int bb = rbit & 63;
                    int op;
                    bool match = false;
                    if (i > 0) {
                        const int vd = (int)((cur.pv >> bb) & 1ull) - (int)((cur.mv >> bb) & 1ull);
                        const int hd = (int)((cur.ph >> bb) & 1ull) - (int)((cur.mh >> bb) & 1ull);
                        int hd1 = 0;  // horizontal delta of row j-1 (row 0: D[0][.] = 0)
                        if (bb > 0) hd1 = (int)((cur.ph >> (bb - 1)) & 1ull) - (int)((cur.mh >> (bb - 1)) & 1ull);
                        else if (rbit > 0) {
                            const ulonglong2 hh = t_cols[((int64_t)max(i - t_c0, 1) * kMyW + (wcur - 1)) * 64 + 1];
                            hd1 = (int)((hh.x >> 63) & 1ull) - (int)((hh.y >> 63) & 1ull);
                        }
                        const int dup = d - vd, dleft = d - hd, ddiag = dup - hd1;
                        const int pq = j - 1 + poff;
                        if ((pq >> 3) != pchunk) { pchunk = pq >> 3; pw = p8[pchunk]; }
                        const int tq = i - 1 + t_toff;
                        if ((tq >> 3) != tchunk) { tchunk = tq >> 3; twd = t_t8[tchunk]; }
                        match = (int)((pw >> ((pq & 7) * 8)) & 0xffu) ==
                                up_char((int)((twd >> ((tq & 7) * 8)) & 0xffu));  // pat[j-1] == T[i-1]
                        if (!match && ddiag + 1 == d) { op = 0; d = ddiag; }    // substitution
                        else if (d == dup + 1) { op = 3; d = dup; }             // Ins: vertical delta +1
                        else if (dleft + 1 == d) { op = 2; d = dleft; }         // Del
                        else { op = 0; d = ddiag; }                             // match
                    } else {
                        op = 3;  // text exhausted: D[j][0] = j
                        d -= 1;
                    }
                    if (op != 3) {  // moved one column to the left: the prefetched record becomes current
                        cur = left;
                        left = load_rec(i - 2, wcur);
                    }
                    if (op == 3) {
                        push_op(3);
                        nindel++; j--;
                        if (want_path) path += (double)g.rqual[t_ro + j] * -0.23025850929940456;  // prob_emit_y
                    } else if (op == 2) {  // Del (prob_emit_x = ln 1)
                        push_op(2);
                        nindel++; i--;
                    } else {               // Match / Subst
                        j--; i--;
                        if (want_path) {
                            const int q = g.rqual[t_ro + j];
                            path += match ? g.qlut[q * kQlutStride + 3]
                                          : (double)q * -0.23025850929940456 + -1.098712293668443;
                        }
                    }
                    if (want_path && next >= 0) path += trans(op, next);
                    next = op;
                    if (j == 0) {  // this alignment is complete
                        if (want_path && next >= 0) path += trans(-1, next);
                        walking = false;
                    }
                }
            }
            if (has) {
                WalkRes &r = s_res[wib][lane_id];
                r.start = i; r.nindel = nindel; r.nruns = nruns; r.e = e; r.path = path;
                for (int t = 0; t < nruns; ++t) r.runs[t] = runs[nruns - 1 - t];  // forward order
            }
            __syncwarp();
            // ---- owner side: fold the results of my tasks, in increasing end position
            for (int t = 0; t < acc; ++t) {
                const WalkRes &r = s_res[wib][excl + t];
                if (want_path && (!have_path || r.path > best_path)) { best_path = r.path; have_path = true; }
                if (r.e == first_best) start_first = r.start;
                start_last = r.start;
                if (r.nindel < best_nindel) {  // fewest indel ops, first on ties (min_by_key)
                    best_nindel = r.nindel;
                    best_nops = r.nruns;
                    for (int q = 0; q < r.nruns; ++q) best_runs[q] = r.runs[q];
                }
            }
            __syncwarp();
        }
        if (work) {
            res.dist = best;
            res.start = start_first;
            res.end = min(start_last + m + best, n);
            res.n_best = n_best;
            res.n_indel = best_nindel;
            g.out[k] = res;
            if (want_path) g.out_path[k] = best_path;
            if (g.out_ops)
                for (int t = 0; t < best_nops; ++t) g.out_ops[k * kMaxIndelOps + t] = best_runs[t];
        }
    }
}

// ---------------------------------------------------------------- fused allele support
// slot -> read index: every (region, haplotype) slot aligns the region's read window
__global__ void slot_read_kernel(int64_t n_regions, const int32_t *region_hap_off, const int32_t *region_read,
                                 int32_t *slot_read) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_regions; r += (int64_t)gridDim.x * blockDim.x)
        for (int s = region_hap_off[r]; s < region_hap_off[r + 1]; ++s) slot_read[s] = region_read[r];
}

struct SupportArgs {
    int64_t n_regions;
    const int32_t *region_read, *region_hap_off, *region_haps, *region_n_ref;
    const int32_t *shrink_extra;   // per haplotype (nullable)
    const int64_t *hap_off, *read_off;
    const uint8_t *rqual;
    const double *qlut;
    const HitOut *hits;            // per slot
    const uint8_t *hit_ops;
    const double *path_prob;       // per slot, fast mode only (else nullptr)
    // pair-HMM work list, per slot
    int64_t *win_start;
    int32_t *win_len, *pair_read, *med;
    double *slot_prob;             // certainty for dist == 0 slots; HMM output otherwise
    uint8_t *slot_mode;            // 0 skip, 1 certainty, 2 hmm
    // outputs
    double *prob_ref, *prob_alt, *raw_ref, *raw_alt;
    int32_t *ref_dist, *alt_dist, *n_indel;
    uint8_t *indel_ops;
    // near ties (first epilogue pass): pair-HMM slots to re-evaluate in literal log space
    const uint8_t *hap_base;
    int32_t *lit_list, *lit_count;
};

// prob_allele, first half (mod.rs:346-362): keep the haplotypes with minimal edit distance and
// prepare their pair-HMM windows (shrink_to_hit, pairhmm.rs:38-44)
__global__ void support_select_kernel(SupportArgs g) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < g.n_regions;
         r += (int64_t)gridDim.x * blockDim.x) {
        const int s0 = g.region_hap_off[r], s1 = g.region_hap_off[r + 1];
        const int nref = g.region_n_ref[r];
        const int rd = g.region_read[r];
        const int64_t ro = g.read_off[rd];
        const int ly = (int)(g.read_off[rd + 1] - ro);
        for (int allele = 0; allele < 2; ++allele) {
            const int a0 = allele == 0 ? s0 : s0 + nref, a1 = allele == 0 ? s0 + nref : s1;
            int best = -1;
            for (int s = a0; s < a1; ++s) {
                const int d = g.hits[s].dist;
                if (d >= 0 && (best < 0 || d < best)) best = d;
            }
            for (int s = a0; s < a1; ++s) {
                const HitOut hit = g.hits[s];
                const int h = g.region_haps[s];
                g.pair_read[s] = rd;
                g.win_start[s] = g.hap_off[h];
                g.win_len[s] = 0;
                g.med[s] = -1;
                g.slot_mode[s] = 0;
                g.slot_prob[s] = -INFINITY;
                if (hit.dist < 0 || hit.dist != best) continue;
                if (hit.dist == 0) {
                    // certainty_est (pairhmm.rs:208-210): sum of ln(1 - eps_j) in read order
                    double p = 0.0;
                    for (int j = 0; j < ly; ++j) p += g.qlut[g.rqual[ro + j] * kQlutStride + 3];
                    g.slot_prob[s] = p;
                    g.slot_mode[s] = 1;
                } else if (g.path_prob) {
                    // --pairhmm-mode fast: probability of the best alignment path(s)
                    g.slot_prob[s] = g.path_prob[s];
                    g.slot_mode[s] = 3;
                } else {
                    const int lx = (int)(g.hap_off[h + 1] - g.hap_off[h]);
                    int ws = 0, we = lx;
                    const int extra = g.shrink_extra ? g.shrink_extra[h] : 0;
                    if (extra >= 0) {  // < 0: not shrinkable (breakend alleles, quirk Q6)
                        const int ref_len = lx - extra;
                        ws = max(hit.start - 2, 0);
                        we = min(hit.end + 2, ref_len) + extra;
                    }
                    g.win_start[s] = g.hap_off[h] + ws;
                    g.win_len[s] = we - ws;
                    g.med[s] = hit.dist + 2;  // dist_upper_bound (edit_distance.rs:180-182)
                    g.slot_mode[s] = 2;
                }
            }
        }
    }
}

// prob_allele second half (mod.rs:364-383) + the region epilogue of allele_support (283-301).
// The reference keeps a read's strand / indel operations iff prob_ref != prob_alt EXACTLY (mod.rs:303)
// and observation.rs:384-388 drops reads without strand information, so the last bit of the two
// pair-HMM values decides a discrete output.  FIRST pass: regions whose raw values are nearly tied
// (relative 1e-9; scaled-linear and log-space arithmetic agree to ~1e-11) put their pair-HMM slots on
// the literal list -- unless both alleles are single windows with identical bytes and band, which
// give identical bits in any deterministic arithmetic.  The literal kernel (LogLit) then replaces
// those slot values by the reference's own operation order with the deterministic libm, and the
// second pass redoes the epilogue on them.
template <bool FIRST>
__global__ void support_epilogue_kernel(SupportArgs g) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < g.n_regions;
         r += (int64_t)gridDim.x * blockDim.x) {
        const int s0 = g.region_hap_off[r], s1 = g.region_hap_off[r + 1];
        const int nref = g.region_n_ref[r];
        double prob[2] = {-INFINITY, -INFINITY};
        int best_slot[2] = {-1, -1};
        bool any_hmm = false;
        for (int allele = 0; allele < 2; ++allele) {
            const int a0 = allele == 0 ? s0 : s0 + nref, a1 = allele == 0 ? s0 + nref : s1;
            bool have = false;
            for (int s = a0; s < a1; ++s) {
                const int mode = g.slot_mode[s];
                if (mode == 0) continue;
                any_hmm |= mode == 2;
                const double p = g.slot_prob[s];
                if (mode == 1) {  // perfect match: unconditional overwrite
                    prob[allele] = p; best_slot[allele] = s; have = true;
                } else if (!have || p > prob[allele]) {
                    prob[allele] = p; best_slot[allele] = s; have = true;
                }
            }
        }
        double pr = prob[0], pa = prob[1];
        if (FIRST && g.lit_list && any_hmm && pr != -INFINITY && pa != -INFINITY &&
            fabs(pr - pa) <= 1e-9 * fmax(1.0, fabs(pr))) {
            bool same = false;
            if (nref == 1 && s1 - s0 == 2 && g.slot_mode[s0] == 2 && g.slot_mode[s0 + 1] == 2 &&
                g.win_len[s0] == g.win_len[s0 + 1] && g.med[s0] == g.med[s0 + 1]) {
                const uint8_t *x = g.hap_base + g.win_start[s0], *y = g.hap_base + g.win_start[s0 + 1];
                same = true;
                for (int k = 0; k < g.win_len[s0]; ++k) same &= x[k] == y[k];
            }
            if (!same)
                for (int s = s0; s < s1; ++s)
                    if (g.slot_mode[s] == 2) {
                        g.slot_mode[s] = 4;  // literal value pending
                        g.lit_list[atomicAdd(g.lit_count, 1)] = s;
                    }
        }
        if (g.raw_ref) g.raw_ref[r] = pr;
        if (g.raw_alt) g.raw_alt[r] = pa;
        if (pr != -INFINITY && pa != -INFINITY) {  // normalise only if both are non-zero
            const double tot = dm::ln_add_exp(pa, pr);
            pr = dm::sub(pr, tot);
            pa = dm::sub(pa, tot);
        }
        if (pr == -INFINITY && pa == -INFINITY) pr = pa = -0.6931471805599453;  // Prob(0.5)
        g.prob_ref[r] = pr;
        g.prob_alt[r] = pa;
        g.ref_dist[r] = best_slot[0] >= 0 ? g.hits[best_slot[0]].dist : -1;
        g.alt_dist[r] = best_slot[1] >= 0 ? g.hits[best_slot[1]].dist : -1;
        int ni = 0;
        if (best_slot[1] >= 0) {
            ni = g.hits[best_slot[1]].n_indel;
            if (g.indel_ops && ni > 0)  // the run-length encoded record (zero padded)
                for (int t = 0; t < kMaxIndelOps; ++t)
                    g.indel_ops[r * kMaxIndelOps + t] = g.hit_ops[(int64_t)best_slot[1] * kMaxIndelOps + t];
        }
        g.n_indel[r] = ni;
    }
}

// PathHMMRealigner::new (realignment/mod.rs:459-486): transition table of the fast mode
static void set_path_consts(MyersArgs &a, const vlr_gap_params *gp) {
    auto l1me = [](double p) -> double { return p < -0.693 ? log1p(-exp(p)) : log(-expm1(p)); };
    auto lae = [](double x, double y) -> double {
        const double p0 = x > y ? x : y, p1 = x > y ? y : x;
        if (p0 == -INFINITY) return -INFINITY;
        if (p1 == -INFINITY) return p0;
        return p0 + log1p(exp(p1 - p0));
    };
    a.gap_x = gp->prob_gap_x;
    a.gap_y = gp->prob_gap_y;
    a.no_gap = l1me(lae(gp->prob_gap_x, gp->prob_gap_y));
    a.close_x = l1me(gp->prob_gap_x_extend);
    a.close_y = l1me(gp->prob_gap_y_extend);
    a.reopen_x = lae(gp->prob_gap_x_extend, a.close_x + gp->prob_gap_x);
    a.reopen_y = lae(gp->prob_gap_y_extend, a.close_y + gp->prob_gap_y);
}

// launches K2 over device-resident pairs: the 4-word instantiation (read windows up to 256 rows) and,
// if the batch holds longer windows, the 16-word one (up to 1024 rows; bio's Myers::Long)
constexpr int kMyWShort = 4, kMyWLong = 16;
template <int KW>
static int launch_besthit_w(vlr_ctx *ctx, MyersArgs &a, int64_t max_lx, int slot_base, int m_lo, int m_hi) {
    cudaStream_t st = ctx->stream;
    const size_t per_worker = (size_t)(max_lx + 1) * (4 * KW * 8 + 4);
    // one thread per pair, latency bound (scratch stores, local-memory masks): many resident warps
    const char *wenv = getenv("VLR_K2_THREADS_PER_SM");
    int64_t workers = (int64_t)ctx->sm_count * (wenv && atoi(wenv) > 0 ? atoi(wenv) : (KW > 4 ? 128 : 512));
    const size_t budget = (size_t)6 << 30;
    if ((size_t)workers * per_worker > budget) workers = (int64_t)(budget / per_worker);
    workers = workers / 128 * 128;
    if (workers < 128) return set_err(ctx, VLR_ERR_UNSUPPORTED, "haplotype window of %lld bytes is too long for the Myers scratch", (long long)max_lx);
    if (workers > a.n_pairs) workers = (a.n_pairs + 127) / 128 * 128;
    uint64_t *scratch = (uint64_t *)dev_scratch(ctx, slot_base, (size_t)workers * (max_lx + 1) * (4 * KW * 8));
    int32_t *dm = (int32_t *)dev_scratch(ctx, slot_base + 1, (size_t)workers * (max_lx + 1) * 4);
    int32_t *cnt = (int32_t *)dev_scratch(ctx, slot_base + 2, 64);
    if (!scratch || !dm || !cnt) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
    VLR_CUDA(ctx, cudaMemsetAsync(cnt, 0, 64, st));
    a.scratch = scratch;
    a.scratch_dm = dm;
    a.scratch_stride = max_lx + 1;
    a.cursor = cnt;
    a.m_lo = m_lo;
    { const char *e = getenv("VLR_K2_EXP"); a.exp_skip = e ? atoi(e) : 0; }
    a.m_hi = m_hi;
    besthit_kernel<KW><<<(int)(workers / 128), 128, 0, st>>>(a);
    VLR_LAUNCH_CHECK(ctx);
    return VLR_OK;
}
static int launch_besthit(vlr_ctx *ctx, MyersArgs &a, int64_t max_lx, int64_t max_ly, int slot_base) {
    int rc = launch_besthit_w<kMyWShort>(ctx, a, max_lx, slot_base, 0, 64 * kMyWShort);
    if (rc == VLR_OK && max_ly > 64 * kMyWShort)
        rc = launch_besthit_w<kMyWLong>(ctx, a, max_lx, slot_base, 64 * kMyWShort, 64 * kMyWLong);
    return rc;
}

}  // namespace vlr

using namespace vlr;

static int besthit_host(vlr_ctx *ctx, int64_t n_pairs, const uint8_t *hap_bytes, const int64_t *hap_off,
                        int64_t n_haps, const uint8_t *read_bases, const uint8_t *read_quals,
                        const int64_t *read_off, int64_t n_reads, const int32_t *pair_hap,
                        const int32_t *pair_read, const vlr_gap_params *gap, int32_t *out_dist,
                        int32_t *out_start, int32_t *out_end, int32_t *out_n_best,
                        int32_t *out_n_indel_ops, uint8_t *out_indel_ops, double *out_path) {
    if (!ctx) return VLR_ERR_INVALID;
    if (n_pairs < 0 || n_pairs > 0x7ff00000) return set_err(ctx, VLR_ERR_INVALID, "n_pairs out of range");
    if (n_pairs == 0) return VLR_OK;
    if (!hap_bytes || !hap_off || !read_bases || !read_off || n_haps <= 0 || n_reads <= 0 ||
        (out_path ? (!read_quals || !gap) : (!out_dist || !out_start || !out_end || !out_n_best)))
        return set_err(ctx, VLR_ERR_INVALID, "null/empty argument");
    if ((!pair_hap && n_haps < n_pairs) || (!pair_read && n_reads < n_pairs))
        return set_err(ctx, VLR_ERR_INVALID, "identity pair mapping needs n_haps/n_reads >= n_pairs");
    VLR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int64_t max_lx = 0, max_ly = 0;
    for (int64_t h = 0; h < n_haps; ++h) {
        if (hap_off[h + 1] < hap_off[h]) return set_err(ctx, VLR_ERR_INVALID, "hap_off not monotone");
        max_lx = std::max(max_lx, hap_off[h + 1] - hap_off[h]);
    }
    for (int64_t r = 0; r < n_reads; ++r) {
        if (read_off[r + 1] < read_off[r]) return set_err(ctx, VLR_ERR_INVALID, "read_off not monotone");
        max_ly = std::max(max_ly, read_off[r + 1] - read_off[r]);
    }
    if (max_ly > 64 * kMyWLong) return set_err(ctx, VLR_ERR_UNSUPPORTED, "read window of %lld rows exceeds %d", (long long)max_ly, 64 * kMyWLong);
    if (pair_hap || pair_read)
        for (int64_t k = 0; k < n_pairs; ++k) {
            if (pair_hap && (pair_hap[k] < 0 || pair_hap[k] >= n_haps)) return set_err(ctx, VLR_ERR_INVALID, "pair_hap out of range");
            if (pair_read && (pair_read[k] < 0 || pair_read[k] >= n_reads)) return set_err(ctx, VLR_ERR_INVALID, "pair_read out of range");
        }
    enum { S_HAP = 40, S_HOFF, S_RB, S_ROFF, S_PH, S_PR, S_OUT, S_OPS, S_SCR, S_SCR2, S_SCR3, S_RQ, S_PATH = 75 };
    const int64_t hap_lo = hap_off[0], rd_lo = read_off[0];
    const size_t hb = (size_t)(hap_off[n_haps] - hap_lo), rb = (size_t)(read_off[n_reads] - rd_lo);
    uint8_t *d_hap = (uint8_t *)dev_scratch(ctx, S_HAP, hb + 64);
    int64_t *d_hoff = (int64_t *)dev_scratch(ctx, S_HOFF, (size_t)(n_haps + 1) * 8);
    uint8_t *d_rb = (uint8_t *)dev_scratch(ctx, S_RB, rb + 16);
    int64_t *d_roff = (int64_t *)dev_scratch(ctx, S_ROFF, (size_t)(n_reads + 1) * 8);
    int32_t *d_ph = pair_hap ? (int32_t *)dev_scratch(ctx, S_PH, (size_t)n_pairs * 4) : nullptr;
    int32_t *d_pr = pair_read ? (int32_t *)dev_scratch(ctx, S_PR, (size_t)n_pairs * 4) : nullptr;
    HitOut *d_out = (HitOut *)dev_scratch(ctx, S_OUT, (size_t)n_pairs * sizeof(HitOut));
    uint8_t *d_ops = (uint8_t *)dev_scratch(ctx, S_OPS, (size_t)n_pairs * kMaxIndelOps);
    if (!d_hap || !d_hoff || !d_rb || !d_roff || !d_out || !d_ops || (pair_hap && !d_ph) || (pair_read && !d_pr))
        return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
    VLR_CUDA(ctx, cudaMemcpyAsync(d_hap, hap_bytes + hap_lo, hb, cudaMemcpyHostToDevice, st));
    VLR_CUDA(ctx, cudaMemcpyAsync(d_hoff, hap_off, (size_t)(n_haps + 1) * 8, cudaMemcpyHostToDevice, st));
    VLR_CUDA(ctx, cudaMemcpyAsync(d_rb, read_bases + rd_lo, rb, cudaMemcpyHostToDevice, st));
    VLR_CUDA(ctx, cudaMemcpyAsync(d_roff, read_off, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    if (pair_hap) VLR_CUDA(ctx, cudaMemcpyAsync(d_ph, pair_hap, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, st));
    if (pair_read) VLR_CUDA(ctx, cudaMemcpyAsync(d_pr, pair_read, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, st));
    VLR_CUDA(ctx, cudaMemsetAsync(d_ops, 0, (size_t)n_pairs * kMaxIndelOps, st));
    MyersArgs a;
    memset(&a, 0, sizeof(a));
    a.hap = d_hap - hap_lo; a.hap_off = d_hoff; a.rbase = d_rb - rd_lo; a.read_off = d_roff;
    a.pair_hap = d_ph; a.pair_read = d_pr; a.n_pairs = n_pairs; a.out = d_out; a.out_ops = d_ops;
    double *d_path = nullptr;
    if (out_path) {
        uint8_t *d_rq = (uint8_t *)dev_scratch(ctx, S_RQ, rb + 16);
        d_path = (double *)dev_scratch(ctx, S_PATH, (size_t)n_pairs * 8);
        if (!d_rq || !d_path) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
        VLR_CUDA(ctx, cudaMemcpyAsync(d_rq, read_quals + rd_lo, rb, cudaMemcpyHostToDevice, st));
        a.rqual = d_rq - rd_lo; a.qlut = ctx->d_qlut; a.out_path = d_path;
        set_path_consts(a, gap);
    }
    int rc = launch_besthit(ctx, a, max_lx, max_ly, S_SCR);
    if (rc != VLR_OK) return rc;
    if (out_path)
        VLR_CUDA(ctx, cudaMemcpyAsync(out_path, d_path, (size_t)n_pairs * 8, cudaMemcpyDeviceToHost, st));
    std::vector<HitOut> h((size_t)n_pairs);
    VLR_CUDA(ctx, cudaMemcpyAsync(h.data(), d_out, (size_t)n_pairs * sizeof(HitOut), cudaMemcpyDeviceToHost, st));
    if (out_indel_ops)
        VLR_CUDA(ctx, cudaMemcpyAsync(out_indel_ops, d_ops, (size_t)n_pairs * kMaxIndelOps, cudaMemcpyDeviceToHost, st));
    VLR_CUDA(ctx, cudaStreamSynchronize(st));
    for (int64_t k = 0; k < n_pairs; ++k) {
        if (out_dist) out_dist[k] = h[k].dist;
        if (out_start) out_start[k] = h[k].start;
        if (out_end) out_end[k] = h[k].end;
        if (out_n_best) out_n_best[k] = h[k].n_best;
        if (out_n_indel_ops) out_n_indel_ops[k] = h[k].n_indel;
    }
    return VLR_OK;
}

extern "C" {

static int vlr_besthit_batch_impl(vlr_ctx *ctx, int64_t n_pairs, const uint8_t *hap_bytes, const int64_t *hap_off,
                      int64_t n_haps, const uint8_t *read_bases, const int64_t *read_off,
                      int64_t n_reads, const int32_t *pair_hap, const int32_t *pair_read,
                      int32_t *out_dist, int32_t *out_start, int32_t *out_end, int32_t *out_n_best,
                      int32_t *out_n_indel_ops, uint8_t *out_indel_ops) {
    return besthit_host(ctx, n_pairs, hap_bytes, hap_off, n_haps, read_bases, nullptr, read_off, n_reads,
                        pair_hap, pair_read, nullptr, out_dist, out_start, out_end, out_n_best,
                        out_n_indel_ops, out_indel_ops, nullptr);
}

int vlr_besthit_batch(vlr_ctx *ctx, int64_t n_pairs, const uint8_t *hap_bytes, const int64_t *hap_off,
                      int64_t n_haps, const uint8_t *read_bases, const int64_t *read_off,
                      int64_t n_reads, const int32_t *pair_hap, const int32_t *pair_read,
                      int32_t *out_dist, int32_t *out_start, int32_t *out_end, int32_t *out_n_best,
                      int32_t *out_n_indel_ops, uint8_t *out_indel_ops) {
    return vlr::drain_on_error(ctx, vlr_besthit_batch_impl(ctx, n_pairs, hap_bytes, hap_off, n_haps, read_bases, read_off, n_reads, pair_hap, pair_read, out_dist, out_start, out_end, out_n_best, out_n_indel_ops, out_indel_ops));
}


static int vlr_pathhmm_batch_impl(vlr_ctx *ctx, int64_t n_pairs, const uint8_t *hap_bytes, const int64_t *hap_off,
                      int64_t n_haps, const uint8_t *read_bases, const uint8_t *read_quals,
                      const int64_t *read_off, int64_t n_reads, const int32_t *pair_hap,
                      const int32_t *pair_read, const vlr_gap_params *gap, double *out_lnprob,
                      int32_t *out_dist) {
    if (!ctx) return VLR_ERR_INVALID;
    if (!out_lnprob) return set_err(ctx, VLR_ERR_INVALID, "null/empty argument");
    return besthit_host(ctx, n_pairs, hap_bytes, hap_off, n_haps, read_bases, read_quals, read_off,
                        n_reads, pair_hap, pair_read, gap, out_dist, nullptr, nullptr, nullptr, nullptr,
                        nullptr, out_lnprob);
}

int vlr_pathhmm_batch(vlr_ctx *ctx, int64_t n_pairs, const uint8_t *hap_bytes, const int64_t *hap_off,
                      int64_t n_haps, const uint8_t *read_bases, const uint8_t *read_quals,
                      const int64_t *read_off, int64_t n_reads, const int32_t *pair_hap,
                      const int32_t *pair_read, const vlr_gap_params *gap, double *out_lnprob,
                      int32_t *out_dist) {
    return vlr::drain_on_error(ctx, vlr_pathhmm_batch_impl(ctx, n_pairs, hap_bytes, hap_off, n_haps, read_bases, read_quals, read_off, n_reads, pair_hap, pair_read, gap, out_lnprob, out_dist));
}


static int vlr_allele_support_batch_impl(vlr_ctx *ctx, int64_t n_regions, const int32_t *region_read,
                             const int32_t *region_hap_off, const int32_t *region_haps,
                             const int32_t *region_n_ref, const uint8_t *hap_bytes,
                             const int64_t *hap_off, int64_t n_haps, const int32_t *shrink_extra,
                             const uint8_t *read_bases, const uint8_t *read_quals,
                             const int64_t *read_off, int64_t n_reads, const vlr_gap_params *gap,
                             uint32_t flags, double *out_prob_ref, double *out_prob_alt,
                             double *out_raw_ref, double *out_raw_alt, int32_t *out_ref_dist,
                             int32_t *out_alt_dist, int32_t *out_n_indel_ops, uint8_t *out_indel_ops) {
    if (!ctx) return VLR_ERR_INVALID;
    const bool trace = getenv("VLR_TRACE") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_prev = now();
    auto lap = [&](const char *what, bool sync) {
        if (!trace) return;
        if (sync) cudaStreamSynchronize(ctx->stream);
        const double t = now();
        fprintf(stderr, "[vlr trace] allele_support %-12s %8.3f ms\n", what, t - t_prev);
        t_prev = t;
    };
    if (n_regions < 0 || n_regions > 0x3fffffff) return set_err(ctx, VLR_ERR_INVALID, "n_regions out of range");
    if (n_regions == 0) return VLR_OK;
    if (!region_read || !region_hap_off || !region_haps || !region_n_ref || !hap_bytes || !hap_off ||
        !read_bases || !read_quals || !read_off || !gap || !out_prob_ref || !out_prob_alt ||
        !out_ref_dist || !out_alt_dist || !out_n_indel_ops || n_haps <= 0 || n_reads <= 0)
        return set_err(ctx, VLR_ERR_INVALID, "null/empty argument");
    VLR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t n_slots = region_hap_off[n_regions];
    if (n_slots <= 0 || n_slots > 0x7ff00000) return set_err(ctx, VLR_ERR_INVALID, "bad region_hap_off");
    enum { S_HAP = 40, S_HOFF, S_RB, S_ROFF, S_PH, S_PR, S_OUT, S_OPS, S_SCR, S_SCR2, S_SCR3,
           S_RQ, S_RR, S_RHO, S_RNR, S_SE, S_WS0, S_WL, S_MED, S_SP, S_SM, S_O1, S_O2, S_O3, S_O4, S_O5,
           S_O6, S_O7, S_O8, S_PWS, S_LIT, S_PATH = 75 };
    const int64_t hap_lo = hap_off[0], rd_lo = read_off[0];
    if (hap_off[n_haps] < hap_lo || read_off[n_reads] < rd_lo)
        return set_err(ctx, VLR_ERR_INVALID, "offsets not monotone");
    const size_t hb = (size_t)(hap_off[n_haps] - hap_lo), rb = (size_t)(read_off[n_reads] - rd_lo);
    auto D = [&](int slot, size_t bytes) { return dev_scratch(ctx, slot, bytes ? bytes : 8); };
    uint8_t *d_hap = (uint8_t *)D(S_HAP, hb + 64);
    int64_t *d_hoff = (int64_t *)D(S_HOFF, (size_t)(n_haps + 1) * 8);
    uint8_t *d_rb = (uint8_t *)D(S_RB, rb + 16), *d_rq = (uint8_t *)D(S_RQ, rb + 16);
    int64_t *d_roff = (int64_t *)D(S_ROFF, (size_t)(n_reads + 1) * 8);
    int32_t *d_rr = (int32_t *)D(S_RR, (size_t)n_regions * 4), *d_rho = (int32_t *)D(S_RHO, (size_t)(n_regions + 1) * 4);
    int32_t *d_rh = (int32_t *)D(S_PH, (size_t)n_slots * 4), *d_rnr = (int32_t *)D(S_RNR, (size_t)n_regions * 4);
    int32_t *d_se = shrink_extra ? (int32_t *)D(S_SE, (size_t)n_haps * 4) : nullptr;
    int32_t *d_slot_read = (int32_t *)D(S_PR, (size_t)n_slots * 4);
    HitOut *d_hits = (HitOut *)D(S_OUT, (size_t)n_slots * sizeof(HitOut));
    uint8_t *d_ops = (uint8_t *)D(S_OPS, (size_t)n_slots * kMaxIndelOps);
    int64_t *d_ws = (int64_t *)D(S_WS0, (size_t)n_slots * 8);
    int32_t *d_wl = (int32_t *)D(S_WL, (size_t)n_slots * 4), *d_med = (int32_t *)D(S_MED, (size_t)n_slots * 4);
    double *d_sp = (double *)D(S_SP, (size_t)n_slots * 8);
    uint8_t *d_sm = (uint8_t *)D(S_SM, (size_t)n_slots);
    double *d_pref = (double *)D(S_O1, (size_t)n_regions * 8), *d_palt = (double *)D(S_O2, (size_t)n_regions * 8);
    double *d_rref = (double *)D(S_O3, (size_t)n_regions * 8), *d_ralt = (double *)D(S_O4, (size_t)n_regions * 8);
    int32_t *d_rd = (int32_t *)D(S_O5, (size_t)n_regions * 4), *d_ad = (int32_t *)D(S_O6, (size_t)n_regions * 4);
    int32_t *d_ni = (int32_t *)D(S_O7, (size_t)n_regions * 4);
    uint8_t *d_io = (uint8_t *)D(S_O8, (size_t)n_regions * kMaxIndelOps);
    if (!d_hap || !d_hoff || !d_rb || !d_rq || !d_roff || !d_rr || !d_rho || !d_rh || !d_rnr ||
        (shrink_extra && !d_se) || !d_slot_read || !d_hits || !d_ops || !d_ws || !d_wl || !d_med ||
        !d_sp || !d_sm || !d_pref || !d_palt || !d_rref || !d_ralt || !d_rd || !d_ad || !d_ni || !d_io)
        return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
    lap("scratch", false);
    // Copies first, checks while they fly: everything the Myers kernel needs goes on the main stream,
    // the base qualities (only used from the selection step on) on the copy stream, and the host
    // validates the offset / index arrays meanwhile.  The byte counts below come from the first and
    // last offsets only; a malformed array is rejected before any kernel reads the data.
    int rc = ensure_copy_stream(ctx);
    if (rc != VLR_OK) return rc;
    cudaStream_t cp = ctx->copy_stream;
    auto H2D = [&](void *d, const void *h, size_t b) { return cudaMemcpyAsync(d, h, b, cudaMemcpyHostToDevice, st); };
    VLR_CUDA(ctx, cudaMemcpyAsync(d_rq, read_quals + rd_lo, rb, cudaMemcpyHostToDevice, cp));
    VLR_CUDA(ctx, cudaEventRecord(ctx->ev_h2d[0], cp));
    VLR_CUDA(ctx, H2D(d_hap, hap_bytes + hap_lo, hb));
    VLR_CUDA(ctx, H2D(d_hoff, hap_off, (size_t)(n_haps + 1) * 8));
    VLR_CUDA(ctx, H2D(d_rb, read_bases + rd_lo, rb));
    VLR_CUDA(ctx, H2D(d_roff, read_off, (size_t)(n_reads + 1) * 8));
    VLR_CUDA(ctx, H2D(d_rr, region_read, (size_t)n_regions * 4));
    VLR_CUDA(ctx, H2D(d_rho, region_hap_off, (size_t)(n_regions + 1) * 4));
    VLR_CUDA(ctx, H2D(d_rh, region_haps, (size_t)n_slots * 4));
    VLR_CUDA(ctx, H2D(d_rnr, region_n_ref, (size_t)n_regions * 4));
    if (shrink_extra) VLR_CUDA(ctx, H2D(d_se, shrink_extra, (size_t)n_haps * 4));
    VLR_CUDA(ctx, cudaMemsetAsync(d_ops, 0, (size_t)n_slots * kMaxIndelOps, st));
    VLR_CUDA(ctx, cudaMemsetAsync(d_io, 0, (size_t)n_regions * kMaxIndelOps, st));
    // ---- validation (host), overlapped with the copies
    int64_t max_lx = 0, max_ly = 0;
    const char *bad = nullptr;
    for (int64_t h = 0; h < n_haps && !bad; ++h) {
        if (hap_off[h + 1] < hap_off[h]) bad = "hap_off not monotone";
        max_lx = std::max(max_lx, hap_off[h + 1] - hap_off[h]);
        if (shrink_extra && shrink_extra[h] > hap_off[h + 1] - hap_off[h]) bad = "shrink_extra exceeds the window";
    }
    for (int64_t r = 0; r < n_reads && !bad; ++r) {
        if (read_off[r + 1] < read_off[r]) bad = "read_off not monotone";
        max_ly = std::max(max_ly, read_off[r + 1] - read_off[r]);
    }
    for (int64_t r = 0; r < n_regions && !bad; ++r) {
        const int s0 = region_hap_off[r], s1 = region_hap_off[r + 1];
        if (s1 < s0 || s1 > n_slots || s0 < 0 || region_n_ref[r] < 0 || region_n_ref[r] > s1 - s0 || region_read[r] < 0 ||
            region_read[r] >= n_reads)
            bad = "a region is malformed";
    }
    for (int64_t k = 0; k < n_slots && !bad; ++k)
        if (region_haps[k] < 0 || region_haps[k] >= n_haps) bad = "region_haps out of range";
    // Myers takes read windows of up to 1024 rows (bio's Myers::Long); the banded pair-HMM behind it takes
    // windows above 248 rows through the anti-diagonal kernel (pairhmm_diag.cuh)
    const int64_t ly_limit = 64 * kMyWLong;
    if (bad || max_ly > ly_limit) {
        // the copies read caller memory: let them finish before the buffers go back to the caller
        cudaStreamSynchronize(st);
        cudaStreamSynchronize(cp);
        if (bad) return set_err(ctx, VLR_ERR_INVALID, "%s", bad);
        return set_err(ctx, VLR_ERR_UNSUPPORTED, "read window of %lld rows exceeds %lld", (long long)max_ly, (long long)ly_limit);
    }
    // workspace of the pair-HMM stage (sized now that the window lengths are known)
    const size_t pws_bytes = vlr_pairhmm_workspace_bytes(n_slots, max_lx, max_ly);
    void *d_pws = D(S_PWS, pws_bytes);
    if (!d_pws) {
        cudaStreamSynchronize(st);
        cudaStreamSynchronize(cp);
        return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
    }
    // slot -> read index (the Myers pairs), on the device
    {
        const int sgrid = (int)std::min<int64_t>((n_regions + 255) / 256, (int64_t)ctx->sm_count * 8);
        slot_read_kernel<<<sgrid, 256, 0, st>>>(n_regions, d_rho, d_rr, d_slot_read);
        VLR_LAUNCH_CHECK(ctx);
    }
    lap("h2d+checks", true);
    // 1. best hit of every (read window, haplotype) slot
    MyersArgs ma;
    memset(&ma, 0, sizeof(ma));
    ma.hap = d_hap - hap_lo; ma.hap_off = d_hoff; ma.rbase = d_rb - rd_lo; ma.read_off = d_roff;
    ma.pair_hap = d_rh; ma.pair_read = d_slot_read; ma.n_pairs = n_slots; ma.out = d_hits; ma.out_ops = d_ops;
    const bool fast = (flags & VLR_REALIGN_FAST) != 0;
    double *d_path = nullptr;
    if (fast) VLR_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_h2d[0], 0));  // the path probability reads the qualities
    if (fast) {
        d_path = (double *)D(S_PATH, (size_t)n_slots * 8);
        if (!d_path) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
        ma.rqual = d_rq - rd_lo; ma.qlut = ctx->d_qlut; ma.out_path = d_path;
        set_path_consts(ma, gap);
    }
    rc = launch_besthit(ctx, ma, max_lx, max_ly, S_SCR);
    if (rc != VLR_OK) return rc;
    lap("myers", true);
    // 2. minimal-distance selection, certainty short-circuit, HMM windows (from here on the base
    // qualities are read: join the copy stream)
    VLR_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_h2d[0], 0));
    SupportArgs sa;
    memset(&sa, 0, sizeof(sa));
    sa.n_regions = n_regions; sa.region_read = d_rr; sa.region_hap_off = d_rho; sa.region_haps = d_rh;
    sa.region_n_ref = d_rnr; sa.shrink_extra = d_se; sa.hap_off = d_hoff; sa.read_off = d_roff;
    sa.rqual = d_rq - rd_lo; sa.qlut = ctx->d_qlut; sa.hits = d_hits; sa.hit_ops = d_ops;
    sa.path_prob = d_path;
    sa.win_start = d_ws; sa.win_len = d_wl; sa.pair_read = d_slot_read; sa.med = d_med;
    sa.slot_prob = d_sp; sa.slot_mode = d_sm;
    sa.prob_ref = d_pref; sa.prob_alt = d_palt; sa.raw_ref = d_rref; sa.raw_alt = d_ralt;
    sa.ref_dist = d_rd; sa.alt_dist = d_ad; sa.n_indel = d_ni; sa.indel_ops = d_io;
    const int grid = (int)std::min<int64_t>((n_regions + 127) / 128, (int64_t)ctx->sm_count * 8);
    support_select_kernel<<<grid, 128, 0, st>>>(sa);
    VLR_LAUNCH_CHECK(ctx);
    // 3. banded pair-HMM on the shrunk windows (mode-2 slots; others have an empty window)
    if (!fast) {
        rc = pairhmm_windows_dev(ctx, n_slots, d_hap - hap_lo, d_ws, d_wl, d_rb - rd_lo, d_rq - rd_lo,
                                 d_roff, d_slot_read, d_med, d_sm, gap, flags, d_pws, pws_bytes, d_sp);
        if (rc != VLR_OK) return rc;
    }
    lap("select+hmm", true);
    // 4. max over haplotypes, normalisation, fallbacks; near ties go through the literal log-space
    // kernel and a second epilogue pass (no host round trip: the list length stays on the device)
    if (!fast) {
        int32_t *d_lit = (int32_t *)D(S_LIT, (size_t)n_slots * 4 + 64);
        if (!d_lit) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
        VLR_CUDA(ctx, cudaMemsetAsync(d_lit, 0, 64, st));
        sa.hap_base = d_hap - hap_lo;
        sa.lit_count = d_lit;
        sa.lit_list = d_lit + 16;
        support_epilogue_kernel<true><<<grid, 128, 0, st>>>(sa);
        VLR_LAUNCH_CHECK(ctx);
        rc = pairhmm_literal_dev(ctx, n_slots, d_hap - hap_lo, nullptr, d_ws, d_wl, d_rb - rd_lo, d_rq - rd_lo, d_roff,
                                 nullptr, d_slot_read, d_med, sa.lit_list, sa.lit_count, gap, flags, max_lx, max_ly, d_sp);
        if (rc != VLR_OK) return rc;
        lap("near ties", true);
    }
    support_epilogue_kernel<false><<<grid, 128, 0, st>>>(sa);
    VLR_LAUNCH_CHECK(ctx);
    auto D2H = [&](void *h, const void *d, size_t b) { return cudaMemcpyAsync(h, d, b, cudaMemcpyDeviceToHost, st); };
    VLR_CUDA(ctx, D2H(out_prob_ref, d_pref, (size_t)n_regions * 8));
    VLR_CUDA(ctx, D2H(out_prob_alt, d_palt, (size_t)n_regions * 8));
    if (out_raw_ref) VLR_CUDA(ctx, D2H(out_raw_ref, d_rref, (size_t)n_regions * 8));
    if (out_raw_alt) VLR_CUDA(ctx, D2H(out_raw_alt, d_ralt, (size_t)n_regions * 8));
    VLR_CUDA(ctx, D2H(out_ref_dist, d_rd, (size_t)n_regions * 4));
    VLR_CUDA(ctx, D2H(out_alt_dist, d_ad, (size_t)n_regions * 4));
    VLR_CUDA(ctx, D2H(out_n_indel_ops, d_ni, (size_t)n_regions * 4));
    if (out_indel_ops) VLR_CUDA(ctx, D2H(out_indel_ops, d_io, (size_t)n_regions * kMaxIndelOps));
    VLR_CUDA(ctx, cudaStreamSynchronize(st));
    lap("epilogue+d2h", false);
    if (trace && !fast) {
        int32_t n_lit = 0;
        cudaMemcpy(&n_lit, (const int32_t *)D(S_LIT, 64), 4, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[vlr trace] allele_support literal log-space slots: %d of %lld\n", n_lit, (long long)n_slots);
    }
    return VLR_OK;
}

int vlr_allele_support_batch(vlr_ctx *ctx, int64_t n_regions, const int32_t *region_read,
                             const int32_t *region_hap_off, const int32_t *region_haps,
                             const int32_t *region_n_ref, const uint8_t *hap_bytes,
                             const int64_t *hap_off, int64_t n_haps, const int32_t *shrink_extra,
                             const uint8_t *read_bases, const uint8_t *read_quals,
                             const int64_t *read_off, int64_t n_reads, const vlr_gap_params *gap,
                             uint32_t flags, double *out_prob_ref, double *out_prob_alt,
                             double *out_raw_ref, double *out_raw_alt, int32_t *out_ref_dist,
                             int32_t *out_alt_dist, int32_t *out_n_indel_ops, uint8_t *out_indel_ops) {
    return vlr::drain_on_error(ctx, vlr_allele_support_batch_impl(ctx, n_regions, region_read, region_hap_off, region_haps, region_n_ref, hap_bytes, hap_off, n_haps, shrink_extra, read_bases, read_quals, read_off, n_reads, gap, flags, out_prob_ref, out_prob_alt, out_raw_ref, out_raw_alt, out_ref_dist, out_alt_dist, out_n_indel_ops, out_indel_ops));
}


}  // extern "C"
