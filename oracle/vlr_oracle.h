/*
 * vlr_oracle.h -- CPU restatement (TEST INFRASTRUCTURE ONLY) of Varlociraptor 4.1.0's
 * per-site likelihood hot path.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product path (libprosic_b200) never links or calls this file.
 *
 * PARITY: the arithmetic of this path lives in the third-party crate `bio 0.37.0`
 * (Cargo.lock:181-184) whose source is not in /root/reference and there is no Rust toolchain
 * in the image, so the restatement follows the reference's call sites plus the published
 * algorithms (Durbin et al. pair-HMM forward recursion, Myers 1999 bit-vector edit distance,
 * log-space Simpson) and cannot be compared bit for bit with the real binary: PARITY UNPINNED
 * in that sense.  What IS pinned: (i) the reference's own likelihood identities
 * (src/variants/model/likelihood.rs:267-386) and a brute-force path enumeration; (ii) the
 * reference's integration tests -- with the restated preprocess/call glue
 * (libprosic_b200/fixtures.py) 31 of its 32 tumor/normal indel testcases satisfy their
 * `expected:` blocks, four of them on thresholds that are floors of the reference's output
 * (e.g. PROB_SOMATIC_TUMOR 590.2 vs >= 590.0), i.e. agreement to ~1e-3 relative end to end
 * (tests/golden/tn_results.txt); (iii) the observation wire format, byte for byte, against
 * records written by the reference (oracle/obs_codec.py).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).
 */
#ifndef VLR_ORACLE_H
#define VLR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- the oracle's libm (detmath.h: msun exp / expm1 / log / log1p, every operation a single
 *      rounded binary64 operation); fn: 0 exp, 1 expm1, 2 log, 3 log1p ---- */
void vlro_ref_math(int fn, long n, const double *x, double *out);

/* ---- bio::stats::probs::LogProb helpers ---- */
double vlro_ln_add_exp(double a, double b);
double vlro_ln_sum_exp(const double *p, size_t n);
double vlro_ln_one_minus_exp(double p);
/* returns NaN to signal the reference's panic */
double vlro_cap_numerical_overshoot(double p, double eps);

/* ---- src/variants/evidence/bases.rs:33-48 ---- */
double vlro_ln_prob_miscall(uint8_t qual); /* ln eps      */
double vlro_ln_prob_call(uint8_t qual);    /* ln (1-eps)  */

/* ---- src/utils/mod.rs:381-407 MiniLogProb wire quantisation ---- */
double vlro_minilogprob_roundtrip(double p);
uint16_t vlro_f64_to_f16_bits(double v);
double vlro_f16_bits_to_f64(uint16_t h);

/* ---- pair-HMM forward (bio::stats::pairhmm::PairHMM::prob_related, Appendix A.2;
 *      emissions src/variants/evidence/realignment/pairhmm.rs:123-211) ----
 * hap: upper-cased haplotype window (len_x), read/qual: read window (len_y).
 * gap_ln = { prob_gap_x (insertion artifact), prob_gap_y (deletion artifact),
 *            prob_gap_x_extend, prob_gap_y_extend }  (pairhmm.rs:73-101)
 * max_edit_dist < 0: unbanded.
 * flags: VLRO_HMM_* below. */
#define VLRO_HMM_SUM3_APPROX 1u  /* M-state sum uses bio's ln_sum3_exp_approx (drop terms e^-10 below max) */
#define VLRO_HMM_PATH_CLOSE  2u  /* close-gap pairing as PathHMMRealigner (realignment/mod.rs:469-470) */
#define VLRO_HMM_SUM3_APPROX3 8u /* three-way ln_sum3_exp_approx (p0; p0+p1; all three), the default */
#define VLRO_HMM_RESET_FM_ONLY 16u /* only fm[curr] is reset after a column swap (oracle only; sensitivity study) */
double vlro_pairhmm_prob_related(const uint8_t *hap, int len_x, const uint8_t *read,
                                 const uint8_t *qual, int len_y, const double gap_ln[4],
                                 int max_edit_dist, unsigned flags);

/* brute force: sum over all alignment paths by explicit recursion (tiny inputs only) */
double vlro_pairhmm_bruteforce(const uint8_t *hap, int len_x, const uint8_t *read,
                               const uint8_t *qual, int len_y, const double gap_ln[4],
                               unsigned flags);

/* ---- Myers best hit (src/variants/evidence/realignment/edit_distance.rs:56-154,
 *      bio::pattern_matching::myers, Appendix A.3) ----
 * ops: 0=Match 1=Subst 2=Del 3=Ins (forward order).  Returns 0 if no hit, else 1. */
typedef struct {
    int32_t dist;
    int32_t start;          /* alignments[0].start */
    int32_t end;            /* min(alignments.last().start + len_y + dist, len_x) */
    int32_t n_best;         /* number of best end positions */
    int32_t n_indel_ops;    /* length of best_indel_operations */
} vlro_hit_t;
int vlro_best_hit(const uint8_t *hap, int len_x, const uint8_t *read, int len_y,
                  vlro_hit_t *hit,
                  uint8_t *indel_ops, int indel_ops_cap,     /* best_indel_operations (2=Del,3=Ins) */
                  int32_t *aln_starts, int aln_cap,          /* start of each best alignment */
                  uint8_t *aln_ops, int32_t *aln_ops_off, int aln_ops_cap /* all alignments' ops */);

/* ---- PathHMM ("fast" mode) src/variants/evidence/realignment/mod.rs:499-576 ---- */
double vlro_pathhmm_prob(const uint8_t *hap, int len_x, const uint8_t *read, const uint8_t *qual,
                         int len_y, const double gap_ln[4], int n_aln, const int32_t *aln_starts,
                         const uint8_t *aln_ops, const int32_t *aln_ops_off);

/* ---- Realigner::prob_allele + allele_support epilogue for ONE merged region
 *      (src/variants/evidence/realignment/mod.rs:225-319, 338-385, 430-444) ----
 * Alleles: allele 0 = ref, allele 1 = alt; each has n_hap[a] candidate haplotypes, laid out
 * consecutively: hap bytes in `haps` at hap_off[h]..hap_off[h+1].
 * shrink_extra[h] >= 0: window is shrunk to the hit and then extended by shrink_extra bytes
 *   (insertion.rs:219-222);  < 0: haplotype is not shrinkable (breakends.rs:649-664, quirk Q6).
 * Outputs: prob_ref, prob_alt (normalised, with the 0.5 fallbacks), informative flag
 * (prob_ref != prob_alt), alt hit + its best_indel_operations. */
typedef struct {
    double prob_ref;
    double prob_alt;
    double raw_ref;      /* before normalisation */
    double raw_alt;
    int32_t informative;
    int32_t alt_dist;
    int32_t ref_dist;
    int32_t n_indel_ops;
} vlro_support_t;
int vlro_allele_support_region(const uint8_t *haps, const int32_t *hap_off, const int32_t n_hap[2],
                               const int32_t *shrink_extra, const uint8_t *read,
                               const uint8_t *qual, int len_y, const double gap_ln[4],
                               unsigned flags, int fast_mode, vlro_support_t *out,
                               uint8_t *indel_ops, int indel_ops_cap);

/* ---- pileup likelihood (src/variants/model/likelihood.rs, bias/ *.rs) ----
 * Observation columns, struct-of-arrays, n_obs rows:
 *   prob_mapping, prob_mismapping, prob_alt, prob_ref, prob_missed_allele, prob_sample_alt,
 *   prob_double_overlap, prob_single_overlap, prob_hit_base   (natural-log doubles)
 *   cat: packed categoricals per row (see VLRO_CAT_*). */
typedef struct {
    const double *prob_mapping, *prob_mismapping, *prob_alt, *prob_ref, *prob_missed_allele,
        *prob_sample_alt, *prob_double_overlap, *prob_single_overlap, *prob_hit_base;
    const uint16_t *cat;
    int32_t n_obs;
} vlro_pileup_t;

/* cat bit layout (mirrors observation.rs:40-45,118-134 and bio-types SequenceReadPairOrientation) */
#define VLRO_CAT_STRAND(c) ((c) & 3u)            /* 0 Forward 1 Reverse 2 Both 3 None      */
#define VLRO_CAT_ORIENT(c) (((c) >> 2) & 15u)    /* 0 F1R2 1 F2R1 2.. other, 8 None         */
#define VLRO_CAT_READPOS(c) (((c) >> 6) & 1u)    /* 0 Major 1 Some                          */
#define VLRO_CAT_SOFTCLIP(c) (((c) >> 7) & 1u)
#define VLRO_CAT_INDEL(c) (((c) >> 8) & 3u)      /* 0 Major 1 Other 2 None                  */
#define VLRO_CAT_PAIRED(c) (((c) >> 10) & 1u)

/* Biases (bias/mod.rs:86-99): packed config */
typedef struct {
    int32_t strand;        /* 0 None 1 Forward 2 Reverse        (strand_bias.rs:8-12)  */
    int32_t orientation;   /* 0 None 1 F1R2 2 F2R1              (read_orientation_bias.rs:9-13) */
    int32_t read_position; /* 0 None 1 Some                     (read_position_bias.rs:7-10) */
    int32_t softclip;      /* 0 None 1 Some                     (softclip_bias.rs:7-10) */
    int32_t divindel;      /* 0 None 1 Some{other_rate}         (divindel_bias.rs:10-16) */
    double other_rate;
    double min_other_rate;
} vlro_biases_t;

double vlro_biases_prob(const vlro_biases_t *b, const vlro_pileup_t *p, int i);
double vlro_biases_prob_any(const vlro_biases_t *b, const vlro_pileup_t *p, int i);
int vlro_biases_is_artifact(const vlro_biases_t *b);
int vlro_is_strong_obs(const vlro_pileup_t *p, int i);
/* gating: is_possible && is_informative && is_likely over all samples' pileups (generic.rs:258-262) */
int vlro_biases_gate(const vlro_biases_t *b, const vlro_pileup_t *pileups, int n_samples);
/* DivIndelBias::learn_parameters (divindel_bias.rs:98-132) */
void vlro_biases_learn(vlro_biases_t *b, const vlro_pileup_t *pileups, int n_samples);

/* likelihood_mapping (likelihood.rs:191-213); returns NaN where the reference would panic */
double vlro_likelihood_mapping(double ln_af, const vlro_biases_t *b, const vlro_pileup_t *p, int i);
/* SampleLikelihoodModel::compute (likelihood.rs:215-245) at allele freq `af` (linear) */
double vlro_likelihood_single(const vlro_pileup_t *p, double af, const vlro_biases_t *b);
/* ContaminatedSampleLikelihoodModel::compute (likelihood.rs:118-152) */
double vlro_likelihood_contaminated(const vlro_pileup_t *p, double purity, double af_primary,
                                    const vlro_biases_t *b_primary, double af_secondary,
                                    const vlro_biases_t *b_secondary);

/* ---- posterior: GenericPosterior / GenericLikelihood / Model (generic.rs:109-365,
 *      calling.rs:528-677, Appendix A.4) ----
 * VAF tree nodes, flattened: */
#define VLRO_NODE_SET 0
#define VLRO_NODE_RANGE 1
#define VLRO_NODE_FALSE 2
#define VLRO_NODE_SKIP 3 /* Variant node that passes (generic.rs:211-233) */
typedef struct {
    int32_t kind;
    int32_t sample;
    int32_t child_off;   /* into children[] */
    int32_t n_children;
    int32_t vaf_off;     /* into vafs[]: SET -> n_vafs values; RANGE -> start,end,left_excl,right_excl */
    int32_t n_vafs;
} vlro_node_t;

typedef struct {
    int32_t root_off;    /* into roots[] */
    int32_t n_roots;
    int32_t bias_off;    /* into biases[] */
    int32_t n_biases;
} vlro_event_t;

typedef struct {
    int32_t n_samples;
    const int32_t *resolution;     /* per sample */
    const int32_t *contaminated;   /* per sample: 0/1 */
    const int32_t *by;             /* per sample */
    const double *purity;          /* per sample (1 - contamination.fraction) */
    const vlro_node_t *nodes;
    const int32_t *children;
    const double *vafs;
    const int32_t *roots;
    const vlro_event_t *events;
    int32_t n_events;
    const vlro_biases_t *biases;   /* learn_parameters is applied per site to a copy */
    int32_t n_biases_total;
} vlro_model_t;

/* VAFRange::observable_min/max (src/grammar/formula.rs:1029-1071) */
double vlro_observable_min(double start, double end, int left_excl, int right_excl, int n_obs);
double vlro_observable_max(double start, double end, int left_excl, int right_excl, int n_obs);
/* GenericPosterior::grid_points (generic.rs:109-122) */
int vlro_grid_points(int n_obs, int resolution);

/* Model::compute for one site.  prior is flat (ln 1) unless prior_cb != NULL.
 * out_event_ln[n_events]: UNNORMALISED per-event ln posterior (before subtracting marginal)
 * out_marginal: ln marginal; out_map_af[n_samples], out_map_bias[n_samples] (index into biases[], -1 = none)
 * out_n_evals: number of distinct likelihood keys evaluated (cache misses), may be NULL. */
typedef double (*vlro_prior_fn)(const double *afs, int n_samples, void *user);
int vlro_model_compute(const vlro_model_t *m, const vlro_pileup_t *pileups, double *out_event_ln,
                       double *out_marginal, double *out_map_af, int32_t *out_map_bias,
                       int64_t *out_n_evals, vlro_prior_fn prior_cb, void *prior_user);
/* the same over a batch of sites (struct-of-arrays columns, rows obs_off[s*n_samples+m] ..), sites dealt
 * round-robin to n_threads threads; flat prior.  Used by the all-cores CPU legs of bench.py. */
int vlro_model_compute_batch(const vlro_model_t *m, const vlro_pileup_t *cols, const int64_t *obs_off,
                             int64_t n_sites, int n_threads, double *out_event_ln, double *out_marginal,
                             double *out_map_af, int32_t *out_map_bias, int64_t *out_n_evals);

#ifdef __cplusplus
}
#endif
#endif
