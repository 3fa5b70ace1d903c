/*
 * vlr_gpu.h -- C ABI of libvlr_b200.so: the B200 (sm_100a) implementation of Varlociraptor
 * 4.1.0's per-site likelihood hot path (pair-HMM realignment + AF-grid pileup likelihood with
 * event integration).  This is the drop-in boundary: plain pointers and sizes, `int` status
 * (0 = OK, <0 = error, message via vlr_last_error), no exceptions cross it, one opaque context
 * per GPU (not thread-safe per context, independent across contexts).
 *
 * The reference has no FFI of its own; the seams these entry points sit behind are Rust traits
 * (paths relative to the reference checkout):
 *   - trait Realigner::calculate_prob_allele / prob_allele / allele_support
 *       src/variants/evidence/realignment/mod.rs:164-335, 338-385, 387-389, 430-444
 *   - bio::stats::pairhmm::PairHMM::prob_related      (call site realignment/mod.rs:439-443)
 *   - EditDistanceCalculation::calc_best_hit          src/variants/evidence/realignment/edit_distance.rs:56-154
 *   - Likelihood::compute of SampleLikelihoodModel / ContaminatedSampleLikelihoodModel
 *       src/variants/model/likelihood.rs:118-152, 215-245
 *   - GenericPosterior::compute + Model::compute      src/variants/model/modes/generic.rs:238-282,
 *       src/calling/variants/calling.rs:568-602, 635-677
 * INTEGRATION.md shows the Rust `extern "C"` block and the trait impls that bind them.
 *
 * All probabilities are natural-log doubles (bio::stats::LogProb) unless stated otherwise.
 */
#ifndef VLR_GPU_H
#define VLR_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VLR_OK 0
#define VLR_ERR_INVALID (-1)   /* bad argument */
#define VLR_ERR_CUDA (-2)      /* CUDA runtime error (see vlr_last_error) */
#define VLR_ERR_NOMEM (-3)
#define VLR_ERR_UNSUPPORTED (-4)

typedef struct vlr_ctx vlr_ctx;

/* ---- context ------------------------------------------------------------------------ */
int vlr_version(void);
/* SHA-1 (hex) of the sources and compiler flags this library was built from; the host mirror compares it
 * with the stamp of the source tree next to it and refuses a stale binary (libprosic_b200/build.py). */
const char *vlr_build_stamp(void);
/* One context per GPU.  Fails with VLR_ERR_CUDA if `device` is not a CUDA device: there is no
 * CPU fallback anywhere in this library. */
int vlr_ctx_create(int device, vlr_ctx **out);
void vlr_ctx_destroy(vlr_ctx *ctx);
const char *vlr_last_error(const vlr_ctx *ctx);
/* Bind the context to a caller-owned CUDA stream (cudaStream_t as void*); NULL = the context's
 * own stream.  All *_dev entry points enqueue on it and do not synchronise. */
int vlr_ctx_set_stream(vlr_ctx *ctx, void *cuda_stream);
int vlr_ctx_synchronize(vlr_ctx *ctx);
/* number of kernels launched through this context so far */
int64_t vlr_ctx_launch_count(const vlr_ctx *ctx);

/* The library's reference-arithmetic libm (csrc/detmath.cuh: exp, expm1, log, log1p as fixed
 * sequences of correctly rounded binary64 operations, identical on host and device).  It backs
 * every result that must be reproducible bit for bit (VLR_HMM_LITERAL below).  fn: 0 exp, 1 expm1,
 * 2 log, 3 log1p over host arrays; ctx == NULL evaluates on the host, else on the device. */
int vlr_ref_math(vlr_ctx *ctx, int fn, int64_t n, const double *x, double *out);

/* ---- pair-HMM forward: replaces PairHMM::prob_related (realignment/mod.rs:439-443) ------
 * Gap parameters: GapParams, realignment/pairhmm.rs:73-101 (ln probabilities). */
typedef struct {
    double prob_gap_x;        /* prob_insertion_artifact        (cli.rs:207-212 default 2.8e-6) */
    double prob_gap_y;        /* prob_deletion_artifact         (cli.rs:213-218 default 5.1e-6) */
    double prob_gap_x_extend; /* prob_insertion_extend_artifact (default 0 -> ln = -inf)        */
    double prob_gap_y_extend; /* prob_deletion_extend_artifact  (default 0 -> ln = -inf)        */
} vlr_gap_params;

/* flags.  The M-state sum of bio 0.37's pair-HMM is `ln_sum3_exp_approx`; the crate source is not in
 * the reference checkout (un-vendored dependency), so both recollections of it are implemented and the
 * choice is a flag (DESIGN.md section 5 has the measured sensitivity):
 *   three-way (default): sorted p0 >= p1 >= p2;  p0 - p1 > 10 -> p0;  else p1 - p2 > 10 -> p0.ln_add_exp(p1);
 *                        else the exact sum of three
 *   two-way:             p0 - p1 > 10 -> p0, else the exact sum of three
 *   neither flag:        always the exact sum */
#define VLR_HMM_SUM3_APPROX 1u  /* two-way ln_sum3_exp_approx */
#define VLR_HMM_PATH_CLOSE 2u   /* close-gap constants paired as PathHMMRealigner (mod.rs:469-470)   */
#define VLR_REALIGN_FAST 4u     /* vlr_allele_support_batch: --pairhmm-mode fast (PathHMMRealigner)     */
#define VLR_HMM_SUM3_APPROX3 8u /* three-way ln_sum3_exp_approx (wins over VLR_HMM_SUM3_APPROX) */
/* vlr_pairhmm_batch only: evaluate every pair in natural-log space in the reference's literal
 * operation order with the library's deterministic libm (bit-reproducible; ~50x slower).  The fused
 * allele support uses this evaluation by itself for regions whose two allele probabilities are
 * nearly tied, because the reference decides `prob_ref != prob_alt` exactly (mod.rs:303). */
#define VLR_HMM_LITERAL 16u
#define VLR_HMM_DEFAULT VLR_HMM_SUM3_APPROX3

/*
 * Host-buffer entry point (H2D, kernels, D2H inside; synchronous).
 *   haps:  n_haps haplotype windows, bytes of window h at hap_bytes[hap_off[h] .. hap_off[h+1])
 *          (already spliced by the host: RefBaseEmission::ref_base(i), pairhmm.rs:213-237,
 *          deletion.rs:300-312, insertion.rs:200-214; lower case is upper-cased on device)
 *   reads: n_reads read windows; bases are decoded IUPAC upper-case (bam::record::Seq),
 *          quals are raw PHRED bytes (ReadEmission::new, pairhmm.rs:160-180)
 *   pairs: pair k aligns read pair_read[k] to haplotype pair_hap[k]; either may be NULL,
 *          meaning the identity mapping k -> k
 *   max_edit_dist: per pair band `hit.dist_upper_bound()` (edit_distance.rs:180-182);
 *          NULL or a negative entry = unbanded
 *   out_lnprob[k]: ln P(read | haplotype), capped at ln 1
 */
int vlr_pairhmm_batch(vlr_ctx *ctx, int64_t n_pairs,
                      const uint8_t *hap_bytes, const int64_t *hap_off, int64_t n_haps,
                      const uint8_t *read_bases, const uint8_t *read_quals,
                      const int64_t *read_off, int64_t n_reads,
                      const int32_t *pair_hap, const int32_t *pair_read,
                      const int32_t *max_edit_dist,
                      const vlr_gap_params *gap, uint32_t flags, double *out_lnprob);

/*
 * Device-resident entry point: every pointer is a device pointer, work is enqueued on the
 * context's stream, nothing is copied or synchronised.  Any byte alignment of the windows is
 * accepted (the kernel aligns its TMA bulk copies down to 16 bytes); the hap buffer must be
 * readable for 32 bytes past hap_off[n_haps] and 16 bytes before hap_off[0].  `workspace` must
 * hold vlr_pairhmm_workspace_bytes(n_pairs, max_len_x, max_len_y) bytes.  Read windows longer
 * than 248 rows take the strip-wise long-read variant when unbanded and the anti-diagonal
 * log-space variant when banded (NaN if the workspace was sized without max_len_x / max_len_y).
 */
size_t vlr_pairhmm_workspace_bytes(int64_t n_pairs, int64_t max_len_x, int64_t max_len_y);
int vlr_pairhmm_batch_dev(vlr_ctx *ctx, int64_t n_pairs,
                          const uint8_t *d_hap_bytes, const int64_t *d_hap_off, int64_t n_haps,
                          const uint8_t *d_read_bases, const uint8_t *d_read_quals,
                          const int64_t *d_read_off, int64_t n_reads,
                          const int32_t *d_pair_hap, const int32_t *d_pair_read,
                          const int32_t *d_max_edit_dist,
                          const vlr_gap_params *gap, uint32_t flags,
                          void *d_workspace, size_t workspace_bytes, double *d_out_lnprob);

/* fp64 DFMA peak of the device (register-resident FMA chains), in TFLOP/s; the roofline
 * denominator for the pair-HMM and likelihood kernels. */
int vlr_measure_fp64_peak(vlr_ctx *ctx, double *out_tflops);

/* ---- Myers best hit: replaces EditDistanceCalculation::calc_best_hit
 *      (edit_distance.rs:56-154); same haps/reads/pairs layout as above (host buffers).
 * out_dist = EditDistanceHit::dist (-1: no hit), out_start = alignments[0].start,
 * out_end = min(alignments.last().start + len_y + dist, len_x), out_n_best = number of best end
 * positions; best_indel_operations (131-144): their count in out_n_indel_ops (nullable) and the
 * operations RUN-LENGTH ENCODED in out_indel_ops[k*VLR_MAX_INDEL_OPS ..] (nullable): one byte per run in
 * forward order, bit 7 = Ins (else Del), bits 0-6 = run length (1..127; a longer run continues in the next
 * byte), zero padded -- 32 runs per alignment, however long the runs (a 60-base insertion is one byte).
 * An alignment with more runs (an unrelated read) has its tail uncounted: the sum of the run lengths is
 * then below out_n_indel_ops.  Read windows up to 1024 rows (Myers::Long). ------ */
#define VLR_MAX_INDEL_OPS 32
int vlr_besthit_batch(vlr_ctx *ctx, int64_t n_pairs,
                      const uint8_t *hap_bytes, const int64_t *hap_off, int64_t n_haps,
                      const uint8_t *read_bases, const int64_t *read_off, int64_t n_reads,
                      const int32_t *pair_hap, const int32_t *pair_read,
                      int32_t *out_dist, int32_t *out_start, int32_t *out_end,
                      int32_t *out_n_best, int32_t *out_n_indel_ops, uint8_t *out_indel_ops);

/* ---- PathHMM: replaces PathHMMRealigner::calculate_prob_allele (`--pairhmm-mode fast`,
 *      realignment/mod.rs:499-576): ln probability of the best Myers alignment path (maximum over
 *      the alignments of every best end position, transition table of mod.rs:459-486, emissions
 *      of pairhmm.rs:182-200).  Same layout as vlr_besthit_batch plus the base qualities;
 *      out_dist (nullable) = the hit's edit distance, -1 and NaN probability if there is no hit. */
int vlr_pathhmm_batch(vlr_ctx *ctx, int64_t n_pairs,
                      const uint8_t *hap_bytes, const int64_t *hap_off, int64_t n_haps,
                      const uint8_t *read_bases, const uint8_t *read_quals,
                      const int64_t *read_off, int64_t n_reads,
                      const int32_t *pair_hap, const int32_t *pair_read,
                      const vlr_gap_params *gap, double *out_lnprob, int32_t *out_dist);

/* ---- fused Realigner::allele_support for merged candidate regions (realignment/mod.rs:225-319
 *      with prob_allele 338-385): region r aligns read window region_read[r] against the
 *      haplotypes region_haps[region_hap_off[r] .. region_hap_off[r+1]) (indices into the hap
 *      CSR); the first region_n_ref[r] of them are candidates of the ref allele, the rest of the
 *      alt allele.  shrink_extra[h] (nullable = all 0): >= 0: the window is shrunk to the hit
 *      (pairhmm.rs:38-44) and extended by that many bytes (insertion.rs:219-222); < 0: not
 *      shrinkable (breakend alleles, breakends.rs:649-664).
 *      Device pipeline: K2 best hits -> minimal-distance haplotypes -> dist == 0 ? certainty_est
 *      : banded pair-HMM (band dist+2; with VLR_REALIGN_FAST in flags: best-path probability
 *      of PathHMMRealigner instead) -> max over haplotypes -> normalisation / 0.5 fallbacks.
 *      Outputs per region: normalised prob_ref / prob_alt, raw (unnormalised) values (nullable),
 *      edit distances of the chosen hits, and the alt hit's best_indel_operations.
 *      Read windows up to 1024 rows in both modes (VLR_ERR_UNSUPPORTED beyond).  Contexts are
 *      independent: calls on different contexts of one device overlap (one per worker thread). -- */
int vlr_allele_support_batch(vlr_ctx *ctx, int64_t n_regions, const int32_t *region_read,
                             const int32_t *region_hap_off, const int32_t *region_haps,
                             const int32_t *region_n_ref,
                             const uint8_t *hap_bytes, const int64_t *hap_off, int64_t n_haps,
                             const int32_t *shrink_extra,
                             const uint8_t *read_bases, const uint8_t *read_quals,
                             const int64_t *read_off, int64_t n_reads,
                             const vlr_gap_params *gap, uint32_t flags,
                             double *out_prob_ref, double *out_prob_alt,
                             double *out_raw_ref, double *out_raw_alt,
                             int32_t *out_ref_dist, int32_t *out_alt_dist,
                             int32_t *out_n_indel_ops, uint8_t *out_indel_ops);

/* ---- pileup likelihood + posterior: replaces Likelihood::compute (likelihood.rs:118-152,
 *      215-245) and GenericPosterior/Model::compute (generic.rs:124-365, calling.rs:568-677) --
 * Observation columns, struct of arrays over all rows of the batch (Observation,
 * observation.rs:172-215).  Pileup of (site s, sample m) = rows obs_off[s*n_samples+m] ..
 * obs_off[s*n_samples+m+1]. */
typedef struct {
    const double *prob_mapping;        /* Observation::prob_mapping()    */
    const double *prob_mismapping;     /* Observation::prob_mismapping() */
    const double *prob_alt;
    const double *prob_ref;
    const double *prob_missed_allele;
    const double *prob_sample_alt;
    const double *prob_double_overlap;
    const double *prob_single_overlap;
    const double *prob_hit_base;
    const uint16_t *cat;               /* packed categoricals, VLR_CAT_* */
} vlr_obs_columns;

/* cat bit layout: strand (observation.rs:40-45), read orientation (bio-types
 * SequenceReadPairOrientation), read_position (118-121), softclipped, indel_operations
 * (130-134), paired */
#define VLR_CAT_STRAND(c) ((c) & 3u)           /* 0 Forward 1 Reverse 2 Both 3 None */
#define VLR_CAT_ORIENT(c) (((c) >> 2) & 15u)   /* 0 F1R2 1 F2R1 2..7 other 8 None   */
#define VLR_CAT_READPOS(c) (((c) >> 6) & 1u)   /* 0 Major 1 Some                    */
#define VLR_CAT_SOFTCLIP(c) (((c) >> 7) & 1u)
#define VLR_CAT_INDEL(c) (((c) >> 8) & 3u)     /* 0 Major 1 Other 2 None            */
#define VLR_CAT_PAIRED(c) (((c) >> 10) & 1u)

/* Biases (bias/mod.rs:86-99) */
typedef struct {
    int32_t strand;        /* 0 None 1 Forward 2 Reverse */
    int32_t orientation;   /* 0 None 1 F1R2 2 F2R1 */
    int32_t read_position; /* 0 None 1 Some */
    int32_t softclip;      /* 0 None 1 Some */
    int32_t divindel;      /* 0 None 1 Some */
    int32_t _pad;
    double min_other_rate; /* DivIndelBias::Some::min_other_rate (cli.rs:420-430) */
} vlr_biases;

/* Likelihood work item: one (site, sample, AF[, secondary AF], bias config) evaluation =
 * one Likelihood::compute cache key (likelihood.rs:128, 221). */
typedef struct {
    int32_t site;
    int32_t sample;
    int32_t bias;          /* index into biases[] */
    int32_t contaminated;  /* 0: SampleLikelihoodModel; 1: ContaminatedSampleLikelihoodModel */
    double af;             /* primary allele frequency (linear) */
    double af_secondary;   /* contaminating sample's allele frequency */
    double purity;         /* 1 - contamination.fraction (generic.rs:303-305) */
    double other_rate;     /* learned DivIndelBias other_rate for this site (divindel_bias.rs:98-132) */
} vlr_lh_item;

int vlr_likelihood_batch(vlr_ctx *ctx, int64_t n_sites, int32_t n_samples,
                         const vlr_obs_columns *obs, const int64_t *obs_off,
                         const vlr_biases *biases, int32_t n_biases,
                         const vlr_lh_item *items, int64_t n_items, double *out_ln_lh);

/* Flattened scenario for the posterior (grammar::VAFTree, vaftree.rs:10-96; model::Event,
 * model/mod.rs:28-44).  Variant nodes are resolved by the host into SKIP/FALSE. */
#define VLR_NODE_SET 0
#define VLR_NODE_RANGE 1
#define VLR_NODE_FALSE 2
#define VLR_NODE_SKIP 3
typedef struct {
    int32_t kind;
    int32_t sample;
    int32_t child_off;  /* into children[] */
    int32_t n_children;
    int32_t vaf_off;    /* into vafs[]: SET: n_vafs ascending values; RANGE: start,end,left_excl,right_excl */
    int32_t n_vafs;
} vlr_node;
typedef struct {
    int32_t root_off;   /* into roots[] */
    int32_t n_roots;
    int32_t bias_off;   /* into biases[] */
    int32_t n_biases;
} vlr_event;
/* Prior over allele-frequency vectors as a small program (Prior::calc_prob, prior.rs:288-431): used
 * when tree paths bind Ranges, whose grid points depend on the site's depth so that no table over the
 * leaves exists.  The recursion of calc_prob sums over the germline alt-allele counts of the samples
 * with a somatic mutation rate; everything that does not depend on the allele frequencies themselves
 * (population prior 628-656, Mendelian alt-count terms 658-786, the germline equality of clonal
 * inheritance 465-467) is tabulated by the host per germline count vector, and the device adds the
 * somatic terms (prob_somatic_mutation, 433-455: ln rate - (2 ln |vaf| + ln genome size), a constant
 * for |vaf| <= 1e-4) at every leaf. */
#define VLR_PRIOR_UNIFORM 0        /* sample with an explicit universe: ln 1 inside it, germline VAF 0 */
#define VLR_PRIOR_SOMATIC 1        /* ploidy + somatic rate: sum over germline alt counts 0..ploidy */
#define VLR_PRIOR_GERMLINE 2       /* ploidy, no somatic rate: the VAF must be a germline VAF k/ploidy */
#define VLR_SOM_NONE 0
#define VLR_SOM_PLAIN 1            /* prob_somatic_mutation(rate, af - germline) */
#define VLR_SOM_CLONAL_DENOVO 2    /* (rate, af - germline - (af_parent - germline_parent)), prior.rs:474-480 */
#define VLR_SOM_CLONAL_SAME 3      /* effective somatic VAF must equal the parent's, prior.rs:481-492 */
typedef struct {
    int32_t kind[8];               /* VLR_PRIOR_* per sample */
    int32_t ploidy[8];             /* -1: none (UNIFORM samples only) */
    int32_t som_kind[8];           /* VLR_SOM_* per sample */
    int32_t parent[8];             /* clonal parent, else -1 */
    double ln_rate[8];             /* ln(somatic effective mutation rate x variant-type fraction) */
    double som_zero[8];            /* prob_somatic_mutation(rate, 0): ln(1 - Simpson integral), host-evaluated */
    double ln_genome_size;
    /* universes of the UNIFORM samples (VAFUniverse::contains): entries uni_off[s] .. uni_off[s+1] of
     * uni_spec, five doubles each: kind (0 single value, 1 range), start, end, left_excl, right_excl */
    int32_t uni_off[9];
    int32_t _pad;
    const double *uni_spec;
    /* ln of the allele-frequency independent factor per germline alt-count vector: mixed radix over
     * the samples, sample 0 slowest, radix ploidy + 1 (1 for UNIFORM samples and ploidy 0) */
    const double *germ_ln;
    int64_t n_germ;
} vlr_prior_prog;

typedef struct {
    int32_t n_samples;          /* <= 8 */
    int32_t n_events;
    int32_t n_nodes;
    int32_t n_biases;
    const int32_t *resolution;  /* per sample (grammar/mod.rs:425-427) */
    const int32_t *contaminated;
    const int32_t *by;
    const double *purity;
    const vlr_node *nodes;
    const int32_t *children;
    const double *vafs;
    const int32_t *roots;
    const vlr_event *events;
    const vlr_biases *biases;
    /* Prior (bio::stats::bayesian::model::Prior::compute, prior.rs:788-805; joint = prior +
     * likelihood, generic.rs:317-365): ln prior of every leaf AF vector in the order of
     * vlr_model_leaves, or NULL = flat (ln 1).  The prior is data-independent, so the host
     * evaluates its unchanged Prior once per event universe.  Only for models whose tree paths
     * bind VAF sets (ploidy-derived universes, e.g. pedigrees): the grid points of a Range depend
     * on the site's depth (generic.rs:109-122) -- such models pass `prior_prog` below. */
    const double *prior_ln;
    int64_t n_prior;
    /* ... or the prior as a program (any universe; required when a tree path binds a Range and the
     * prior is not flat).  NULL = none.  prior_ln wins when both are given and applicable. */
    const vlr_prior_prog *prior_prog;
} vlr_model;

/* Leaves (event, tree path, value combination; last sample fastest) of a set-only model:
 * out_event[leaf] and out_afs[leaf*n_samples + m], both nullable to query the count alone. */
int vlr_model_leaves(vlr_ctx *ctx, const vlr_model *model, int64_t capacity, int64_t *out_n_leaves,
                     int32_t *out_event, double *out_afs);

/*
 * Whole part 2 for a batch of sites: bias learning + gating (bias/mod.rs:30-83, 192-223),
 * AF-grid likelihoods, VAF-tree integration (Simpson, generic.rs:196-208), marginal, MAP.
 * Prior: model->prior_ln (flat = ln 1 when NULL).
 *   out_event_ln[s*n_events + e]  unnormalised ln posterior of event e (Posterior::compute)
 *   out_marginal[s]               ln marginal (Model::compute)
 *   out_map_af[s*n_samples + m]   MAP allele frequency (calling.rs:635-677)
 *   out_map_bias[s*n_samples + m] bias index of an artifact MAP, else -1
 */
int vlr_posterior_batch(vlr_ctx *ctx, int64_t n_sites, const vlr_model *model,
                        const vlr_obs_columns *obs, const int64_t *obs_off,
                        double *out_event_ln, double *out_marginal, double *out_map_af,
                        int32_t *out_map_bias);

/* Compact host-buffer variant: the seven STORED observation fields as f32 (30 bytes per
 * observation instead of 74 over PCIe).  The wire format holds every value as f16 or f32
 * (MiniLogProb, utils/mod.rs:381-407), so f32 is lossless for real `preprocess` output; the two
 * derived columns are computed on the device as ObservationBuilder does (observation.rs:218-226).
 * Both host-buffer variants stream the batch in chunks: H2D of chunk c+1 overlaps the kernels of
 * chunk c (pass pinned host memory for the copies to be asynchronous). */
typedef struct {
    const float *prob_mapping;
    const float *prob_alt;
    const float *prob_ref;
    const float *prob_missed_allele;
    const float *prob_sample_alt;
    const float *prob_double_overlap;
    const float *prob_hit_base;
    const uint16_t *cat;
} vlr_obs_columns_f32;
int vlr_posterior_batch_f32(vlr_ctx *ctx, int64_t n_sites, const vlr_model *model,
                            const vlr_obs_columns_f32 *obs, const int64_t *obs_off,
                            double *out_event_ln, double *out_marginal, double *out_map_af,
                            int32_t *out_map_bias);

/* device-resident variant (obs columns / obs_off / outputs are device pointers; `model` stays a
 * host struct).  max_obs_per_pileup: upper bound of observations per (site, sample), e.g.
 * --max-depth (cli.rs:246-252); sites exceeding it get NaN outputs. */
int vlr_posterior_batch_dev(vlr_ctx *ctx, int64_t n_sites, const vlr_model *model,
                            const vlr_obs_columns *d_obs, const int64_t *d_obs_off,
                            int32_t max_obs_per_pileup, double *d_out_event_ln,
                            double *d_out_marginal, double *d_out_map_af, int32_t *d_out_map_bias);

/* ---- values of the final call record: replaces the numeric part of the BCF writer of
 *      `varlociraptor call` (src/calling/variants/mod.rs:99-399, calling.rs:573-602).
 * event_is_artifact[e] marks the artifact twin of an event (model::Event::is_artifact); the
 * output has one column per non-artifact event, in order, plus a final "artifact" column:
 *   out_prob[s*(n_named+1) + k]  PROB_<EVENT> = |PHRED(posterior)| as f32 (mod.rs:326-333); BCF
 *                                float-missing bits when every pileup of the site is empty
 *   out_af[s*n_samples + m]      AF as f32 (mod.rs:185)
 *   out_dp[s*n_samples + m]      DP = round(exp(ln_sum_exp(prob_mapping))) (observation.rs:31-36)
 * Inputs are the outputs of vlr_posterior_batch and the prob_mapping column (host buffers). */
int vlr_call_records(vlr_ctx *ctx, int64_t n_sites, int32_t n_events, int32_t n_samples,
                     const uint8_t *event_is_artifact, const double *event_ln,
                     const double *marginal, const double *map_af, const double *prob_mapping,
                     const int64_t *obs_off, float *out_prob, float *out_af, int32_t *out_dp);

/* OBS / SOBS FORMAT strings and bias flag characters of the call record (calling/variants/mod.rs:
 * 147-260, 335-392; utils/mod.rs:101-146).  vlr_obs_codes derives, on the device, the packed
 * 7-character code of every observation row (host arrays in, host array out):
 *   bits 0-2 Kass-Raftery letter of exp(prob_alt - prob_ref) (0 N, 1 E, 2 B, 3 P, 4 S, 5 V), bit 3 lower
 *   case (prob_mapping < ln 0.95), bit 4 paired (p|s), bits 5-6 strand (0 *, 1 -, 2 +), bits 7-8 read
 *   orientation (0 >, 1 <, 2 *, 3 !), bit 9 read position (0 ^, 1 *), bit 10 softclip ($|.), bits 11-12
 *   indel operations (0 *, 1 #, 2 .)
 * vlr_obs_cigar (host, no context) formats the generalized CIGAR of one (site, sample) pileup from its
 * codes: simple = 0 -> OBS, 1 -> SOBS; returns the length (out is NUL terminated; "." for no rows) or a
 * negative error.  Most common first, E.. after the others, N.. last; equal counts in code order (the
 * reference's order among equal counts is a HashMap's).  vlr_bias_chars: SB ROB RPB SCB DIB. */
int vlr_obs_codes(vlr_ctx *ctx, int64_t n_rows, const double *prob_alt, const double *prob_ref,
                  const double *prob_mapping, const uint16_t *cat, uint16_t *out_code);
int64_t vlr_obs_cigar(const uint16_t *codes, int64_t n, int simple, char *out, int64_t capacity);
int vlr_bias_chars(const vlr_biases *biases, char out[5]);

/* ---- observation wire format: the INFO fields between `preprocess` and `call` -------------
 * Replaces read_observations / write_observations (src/calling/variants/preprocessing/mod.rs:
 * 452-598, OBSERVATION_FORMAT_VERSION "8"; format 7 = the same without INDEL_OPERATIONS) and
 * MiniLogProb (src/utils/mod.rs:381-407).  Each field is the BCF INFO integer vector of one
 * record (bincode bytes as u16 little endian); fields are indexed by VLR_OBS_*.  Host functions,
 * no context needed. */
#define VLR_OBS_PROB_MAPPING 0
#define VLR_OBS_PROB_REF 1
#define VLR_OBS_PROB_ALT 2
#define VLR_OBS_PROB_MISSED_ALLELE 3
#define VLR_OBS_PROB_SAMPLE_ALT 4
#define VLR_OBS_PROB_DOUBLE_OVERLAP 5
#define VLR_OBS_PROB_HIT_BASE 6
#define VLR_OBS_STRAND 7
#define VLR_OBS_READ_ORIENTATION 8
#define VLR_OBS_READ_POSITION 9
#define VLR_OBS_SOFTCLIPPED 10
#define VLR_OBS_INDEL_OPERATIONS 11
#define VLR_OBS_PAIRED 12
#define VLR_OBS_N_FIELDS 13
/* Decode one record into rows of the observation columns: out_cols[0..8] are the nine column
 * pointers in vlr_obs_columns order (already advanced to the record's first row), out_cat the
 * packed categoricals.  prob_mismapping / prob_single_overlap are derived as ObservationBuilder
 * does (observation.rs:218-226).  fields[VLR_OBS_INDEL_OPERATIONS] may be NULL (format 7):
 * IndelOperations::None.  With out_cols == NULL and out_cat == NULL only *out_n_obs is set. */
int vlr_obs_decode(const int32_t *const *fields, const int32_t *n_values, int64_t capacity,
                   int64_t *out_n_obs, double *const *out_cols, uint16_t *out_cat);
/* Encode n_obs rows (cols[0..8] in vlr_obs_columns order; the derived columns 1 and 7 are not
 * read) as format 8: each out_fields[f] receives at most `capacity` values
 * (vlr_obs_encoded_bound(n_obs) suffices), out_n_values[f] their count.  Applies
 * MiniLogProb::new (f16 iff p < -10 and the f16 value keeps the integer part, else f32). */
int64_t vlr_obs_encoded_bound(int64_t n_obs);
int vlr_obs_encode(int64_t n_obs, const double *const *cols, const uint16_t *cat,
                   int32_t *const *out_fields, int64_t capacity, int32_t *out_n_values);

/* ---- the wire format on the device -------------------------------------------------------
 * A batch of observation records AS STORED (what read_observations would parse, mod.rs:455-527): per
 * field the u16 words of all records back to back (the BCF INFO integer values narrowed to u16, as
 * mod.rs:473-479 does), record r = site * n_samples + sample occupying words[f][word_off[f][r] ..
 * word_off[f][r+1]).  words[VLR_OBS_INDEL_OPERATIONS] == NULL: format 7 (IndelOperations::None). */
typedef struct {
    const uint16_t *words[VLR_OBS_N_FIELDS];
    const int64_t *word_off[VLR_OBS_N_FIELDS];   /* n_records + 1 entries each */
} vlr_obs_wire;
/* Device-side read_observations: uploads the records, decodes them in one kernel (one thread per
 * record and field walks the variable-length MiniLogProb vector; prob_mismapping /
 * prob_single_overlap derived as ObservationBuilder does) and returns the struct-of-arrays columns
 * (out_cols[0..8] in vlr_obs_columns order, rows out_obs_off[r] .. out_obs_off[r+1] of record r).
 * out_cols == NULL and out_cat == NULL: only out_obs_off is filled (size query, no device work).
 * A malformed vector (bad variant index, length mismatch, trailing bytes) -> VLR_ERR_INVALID. */
int vlr_obs_decode_batch(vlr_ctx *ctx, int64_t n_records, const vlr_obs_wire *wire, int64_t capacity,
                         int64_t *out_obs_off, double *const *out_cols, uint16_t *out_cat);
/* Host-side write_observations for a batch (the counterpart a `preprocess` binding calls on the columns
 * part 1 produced): appends record r's field f at out_words[f] + out_word_off[f][r]; every out_words[f]
 * holds `capacity_words` u16 (sum over records of vlr_obs_encoded_bound(n_obs) suffices). */
int vlr_obs_encode_batch(int64_t n_records, const double *const *cols, const uint16_t *cat, const int64_t *obs_off,
                         uint16_t *const *out_words, int64_t *const *out_word_off, int64_t capacity_words);
/* Whole part 2 straight from stored records: upload (chunked, H2D of chunk c+1 under the kernels of
 * chunk c), device decode, vlr_posterior_batch_dev.  n_records = n_sites * model->n_samples. */
int vlr_posterior_batch_wire(vlr_ctx *ctx, int64_t n_sites, const vlr_model *model, const vlr_obs_wire *wire,
                             double *out_event_ln, double *out_marginal, double *out_map_af,
                             int32_t *out_map_bias);

#ifdef __cplusplus
}
#endif
#endif
