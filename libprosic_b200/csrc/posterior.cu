// posterior.cu -- host side of K3/K4: model flattening, shared-memory sizing, launch selection and
// the host-buffer pipelines behind vlr_posterior_batch[_dev|_f32], vlr_model_leaves and
// vlr_likelihood_batch.  The kernel lives in posterior_kernel.cuh and is instantiated per
// (team size, CTAs per SM, compile-time sample count) in posterior_inst_*.cu.
#include "posterior_kernel.cuh"

namespace vlr {

// ---------------------------------------------------------------- host: model flattening
struct HostModel {
    std::vector<DevSpectrum> spectra;
    std::vector<DevPath> paths;
    std::vector<DevEvent> events;
    std::vector<DevEventBias> eb;
    std::vector<vlr_biases> cfg;
    std::vector<double> vafs;
    std::vector<DevSeg> segs;
    std::vector<int64_t> path_leaf_off;  // per path: first leaf (prior-table entry), -1 if it binds a Range
    int64_t n_leaves = 0;                // leaves of all-Set paths
    bool any_range_path = false;
};

static bool same_cfg(const vlr_biases &a, const vlr_biases &b) {
    return a.strand == b.strand && a.orientation == b.orientation &&
           a.read_position == b.read_position && a.softclip == b.softclip &&
           a.divindel == b.divindel && a.min_other_rate == b.min_other_rate;
}

static int flatten_model(vlr_ctx *ctx, const vlr_model *mdl, HostModel &H) {
    const int NS = mdl->n_samples;
    if (NS < 1 || NS > kMaxSamples) return set_err(ctx, VLR_ERR_UNSUPPORTED, "n_samples must be 1..8");
    if (mdl->n_events < 1 || mdl->n_events > kMaxEvents)
        return set_err(ctx, VLR_ERR_UNSUPPORTED, "n_events must be 1..%d", kMaxEvents);
    // copy vafs referenced by nodes lazily: keep the caller's array layout
    int max_vaf = 0;
    for (int k = 0; k < mdl->n_nodes; ++k) {
        const vlr_node &nd = mdl->nodes[k];
        if (nd.kind == VLR_NODE_SET) max_vaf = std::max(max_vaf, nd.vaf_off + nd.n_vafs);
        if (nd.kind == VLR_NODE_RANGE) max_vaf = std::max(max_vaf, nd.vaf_off + 4);
    }
    H.vafs.assign(mdl->vafs, mdl->vafs + max_vaf);
    auto spectrum_of = [&](const vlr_node &nd) -> int {
        for (size_t s = 0; s < H.spectra.size(); ++s) {
            const DevSpectrum &sp = H.spectra[s];
            if (sp.sample != nd.sample || sp.kind != nd.kind) continue;
            const int n = nd.kind == VLR_NODE_RANGE ? 4 : nd.n_vafs;
            if (nd.kind == VLR_NODE_SET && sp.n_vafs != nd.n_vafs) continue;
            bool eq = true;
            for (int i = 0; i < n && eq; ++i) eq = H.vafs[sp.vaf_off + i] == H.vafs[nd.vaf_off + i];
            if (eq) return (int)s;
        }
        H.spectra.push_back(DevSpectrum{nd.sample, nd.kind, nd.vaf_off, nd.kind == VLR_NODE_RANGE ? 4 : nd.n_vafs});
        return (int)H.spectra.size() - 1;
    };
    for (int e = 0; e < mdl->n_events; ++e) {
        const vlr_event &ev = mdl->events[e];
        DevEvent de;
        de.path_off = (int)H.paths.size();
        de.eb_off = (int)H.eb.size();
        bool art = false;
        for (int k = 0; k < ev.n_biases; ++k) {
            const vlr_biases &b = mdl->biases[ev.bias_off + k];
            art |= b.strand || b.orientation || b.read_position || b.softclip || b.divindel;
            int id = -1;
            for (size_t c = 0; c < H.cfg.size(); ++c) if (same_cfg(H.cfg[c], b)) { id = (int)c; break; }
            if (id < 0) { H.cfg.push_back(b); id = (int)H.cfg.size() - 1; }
            H.eb.push_back(DevEventBias{e, id});
        }
        de.n_eb = ev.n_biases;
        de.is_artifact = art;
        de.bias_prior = art ? log(0.5) + log(1.0 / (double)ev.n_biases) : log(0.5);
        // depth-first path extraction (generic.rs:124-235): SET/RANGE bind a sample, SKIP passes,
        // FALSE kills the branch, >1 children fan out, no children = leaf
        struct Frame { int node; DevPath p; };
        std::vector<Frame> stack;
        for (int r = ev.n_roots - 1; r >= 0; --r) {
            Frame f;
            f.node = mdl->roots[ev.root_off + r];
            f.p.event = e;
            for (int s = 0; s < kMaxSamples; ++s) f.p.spec[s] = -1;
            stack.push_back(f);
        }
        // (stack order reversed above so that paths come out in root order)
        std::vector<DevPath> out;
        while (!stack.empty()) {
            Frame f = stack.back();
            stack.pop_back();
            if (f.node < 0 || f.node >= mdl->n_nodes) return set_err(ctx, VLR_ERR_INVALID, "bad node index");
            const vlr_node &nd = mdl->nodes[f.node];
            if (nd.kind == VLR_NODE_FALSE) continue;
            if (nd.kind == VLR_NODE_SET || nd.kind == VLR_NODE_RANGE) {
                if (nd.sample < 0 || nd.sample >= NS) return set_err(ctx, VLR_ERR_INVALID, "bad sample index");
                if (nd.kind == VLR_NODE_SET && nd.n_vafs < 1) return set_err(ctx, VLR_ERR_INVALID, "empty VAF set");
                f.p.spec[nd.sample] = spectrum_of(nd);
            } else if (nd.kind != VLR_NODE_SKIP) {
                return set_err(ctx, VLR_ERR_INVALID, "bad node kind");
            }
            if (nd.n_children == 0) {
                for (int s = 0; s < NS; ++s)
                    if (f.p.spec[s] < 0)
                        return set_err(ctx, VLR_ERR_UNSUPPORTED, "event %d: a tree path does not bind sample %d", e, s);
                out.push_back(f.p);
            } else {
                for (int c = nd.n_children - 1; c >= 0; --c) {
                    Frame g2 = f;
                    g2.node = mdl->children[nd.child_off + c];
                    stack.push_back(g2);
                }
            }
        }
        de.n_paths = (int)out.size();
        H.paths.insert(H.paths.end(), out.begin(), out.end());
        H.events.push_back(de);
    }
    // leaves (event, path, combination) in visiting order: the index space of the prior table
    H.path_leaf_off.assign(H.paths.size(), -1);
    for (size_t pi = 0; pi < H.paths.size(); ++pi) {
        int64_t combos = 1;
        bool all_set = true;
        for (int smp = 0; smp < NS; ++smp) {
            const DevSpectrum &sp = H.spectra[H.paths[pi].spec[smp]];
            if (sp.kind == VLR_NODE_RANGE) all_set = false;
            else combos *= sp.n_vafs;
        }
        if (all_set) { H.path_leaf_off[pi] = H.n_leaves; H.n_leaves += combos; }
        else H.any_range_path = true;
    }
    for (int e = 0; e < mdl->n_events; ++e) {
        DevEvent &de = H.events[e];
        de.seg_off = (int)H.segs.size();
        de.n_seg = de.n_eb * de.n_paths;
        for (int q = 0; q < de.n_eb; ++q)
            for (int pth = 0; pth < de.n_paths; ++pth) {
                DevSeg sg;
                memset(&sg, 0, sizeof(sg));
                sg.event = e; sg.ebi = de.eb_off + q; sg.cfg = H.eb[de.eb_off + q].cfg; sg.path = pth;
                const vlr_biases &b = H.cfg[sg.cfg];
                sg.art = (b.strand || b.orientation || b.read_position || b.softclip || b.divindel) ? 1 : 0;
                sg.leaf_off = H.path_leaf_off[de.path_off + pth];
                for (int smp = 0; smp < NS; ++smp) {
                    const int spi = H.paths[de.path_off + pth].spec[smp];
                    sg.spec[smp] = (int16_t)spi;
                    sg.nv[smp] = H.spectra[spi].kind == VLR_NODE_RANGE ? (int16_t)-1 : (int16_t)H.spectra[spi].n_vafs;
                    if (H.spectra[spi].kind != VLR_NODE_RANGE && H.spectra[spi].n_vafs > 32767)
                        return set_err(ctx, VLR_ERR_UNSUPPORTED, "VAF set with more than 32767 values");
                }
                H.segs.push_back(sg);
            }
    }
    if ((int)H.spectra.size() > kMaxSpectra) return set_err(ctx, VLR_ERR_UNSUPPORTED, "too many distinct VAF spectra");
    if ((int)H.cfg.size() > kMaxCfg) return set_err(ctx, VLR_ERR_UNSUPPORTED, "too many bias configurations");
    return VLR_OK;
}

// largest Simpson grid a Range spectrum of `sample` can get when no pileup holds more than max_obs
// observations: min(max(n_obs + 1, 5), resolution) made odd (generic.rs:109-122)
static int64_t grid_bound(const vlr_model *mdl, int sample, int64_t max_obs) {
    return std::min<int64_t>(std::max<int64_t>(max_obs + 1, 5), mdl->resolution[sample]) | 1;
}

// upper bound of the likelihood-table size (doubles) for one site
static int64_t table_bound(const vlr_model *mdl, const HostModel &H, int64_t max_obs) {
    int64_t nv[kMaxSamples] = {0};
    for (const DevSpectrum &sp : H.spectra) {
        // grid = min(max(n_obs + 1, 5), resolution) made odd (generic.rs:109-122) <= resolution | 1
        const int64_t c = sp.kind == VLR_NODE_RANGE ? grid_bound(mdl, sp.sample, max_obs) : sp.n_vafs;
        nv[sp.sample] += c;
    }
    int64_t tot = 0;
    for (int s = 0; s < mdl->n_samples; ++s) {
        int64_t keys = mdl->contaminated[s] ? nv[s] * nv[mdl->by[s]] : nv[s];
        tot += keys * (int64_t)H.cfg.size();
    }
    return tot;
}

}  // namespace vlr

using namespace vlr;

extern "C" {

int vlr_posterior_batch_dev(vlr_ctx *ctx, int64_t n_sites, const vlr_model *model,
                            const vlr_obs_columns *d_obs, const int64_t *d_obs_off,
                            int32_t max_obs_per_pileup, double *d_out_event_ln,
                            double *d_out_marginal, double *d_out_map_af, int32_t *d_out_map_bias) {
    if (!ctx) return VLR_ERR_INVALID;
    if (!model || !d_obs || !d_obs_off || !d_out_event_ln || !d_out_marginal || !d_out_map_af ||
        !d_out_map_bias || n_sites < 0 || n_sites > 0x7fffffff || max_obs_per_pileup < 0)
        return set_err(ctx, VLR_ERR_INVALID, "bad argument");
    if (n_sites == 0) return VLR_OK;
    VLR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    HostModel H;
    int rc = flatten_model(ctx, model, H);
    if (rc != VLR_OK) return rc;
    for (int s = 0; s < model->n_samples; ++s)
        if (model->contaminated[s] && (model->by[s] < 0 || model->by[s] >= model->n_samples || model->by[s] == s))
            return set_err(ctx, VLR_ERR_INVALID, "bad contamination source for sample %d", s);
    // the kernel admits pileups up to obs_cap observations: size every per-site bound for that depth
    const int obs_cap = std::max(8, (max_obs_per_pileup + 7) & ~7);
    const int64_t tbound = table_bound(model, H, obs_cap);
    // team size: whole CTA for big tables (tumor/normal-like), one warp per site for small ones
    size_t team_smem = (size_t)obs_cap * sizeof(double4) + (size_t)obs_cap * H.cfg.size() * sizeof(double);
    // blocked contaminated path: cfgs x observations x (values of the contaminant, padded to 8)
    int64_t gs_cap = 0;
    int val_cap = 0;
    {
        int64_t nvb[kMaxSamples] = {0}, nv_all = 0;
        for (const DevSpectrum &sp : H.spectra) {
            const int64_t c = sp.kind == VLR_NODE_RANGE ? grid_bound(model, sp.sample, obs_cap) : sp.n_vafs;
            nvb[sp.sample] += c;
            nv_all += c;
        }
        // allele-frequency values + ln weights of one site: sized for the deepest pileup the kernel
        // admits, so no site can exceed it (the shared-memory total is checked below)
        if (nv_all > kMaxValues)
            return set_err(ctx, VLR_ERR_UNSUPPORTED, "the model needs up to %lld allele-frequency values per site "
                           "(limit %d)", (long long)nv_all, kMaxValues);
        val_cap = (int)((nv_all + 3) & ~(int64_t)3);
        team_smem += (size_t)val_cap * 16;
        for (int smp = 0; smp < model->n_samples; ++smp)
            if (model->contaminated[smp])
                gs_cap = std::max(gs_cap, (int64_t)H.cfg.size() * obs_cap * ((nvb[model->by[smp]] + 7) & ~(int64_t)7));
        if (team_smem + (size_t)gs_cap * 8 > 72 * 1024) gs_cap = 0;  // does not fit: direct evaluation
    }
    team_smem += (size_t)gs_cap * 8;
    // likelihood table of one site in shared memory when it is small enough
    const int64_t tab_cap = (tbound <= 4096 && team_smem + (size_t)tbound * 8 <= 72 * 1024)
                                ? ((tbound + 3) & ~(int64_t)3) : 0;  // keeps the teams' carve-outs 32-byte aligned
    team_smem += (size_t)tab_cap * 8;
    const int seg_cap = ((int)H.segs.size() + 1 + 7) & ~7;  // ints (prefix of n_segs + 1 entries); keeps the carve-outs 32-byte aligned
    team_smem += (size_t)seg_cap * 4;
    const bool big = tbound > 2048 || 4 * team_smem > 64 * 1024;
    const int teams_per_cta = big ? 1 : 4;
    const size_t dyn_smem = (size_t)teams_per_cta * team_smem;
    if (dyn_smem > 160 * 1024)
        return set_err(ctx, VLR_ERR_UNSUPPORTED,
                       "%d observations x %d bias configurations do not fit the shared-memory staging",
                       max_obs_per_pileup, (int)H.cfg.size());
    // static shared memory: TeamShared per team + event partials (4 KB) + per-CTA reserve (1 KB).
    // One-warp teams are latency bound (serial per-site sections): up to 7 CTAs per SM = 28 warps at
    // 72 registers with 48-96 bytes of spills (measured best of 3..8 on the C2 / C5 workloads once the
    // prior program was out of the program-free instantiations: germline 157 -> 167 M sites/s, trio
    // 48 -> 56 M against 5 CTAs at 96 registers; 8 CTAs at 64 registers lose again); the register cap
    // of the instantiation follows the number of CTAs that fit.
    const size_t cta_smem = dyn_smem + (size_t)teams_per_cta * sizeof(TeamShared) + 4096 + 1024 + 1024;
    bool cont_or_generic = model->n_samples > 3;   // the instantiations with the contaminated paths / run-time
    for (int k = 0; k < model->n_samples; ++k)     // sample count spill heavily below 96 registers: 5 CTAs
        cont_or_generic |= model->contaminated[k] != 0;
    int minb = (int)std::max<size_t>(1, std::min<size_t>(big ? 3 : (cont_or_generic ? 5 : 7), (size_t)(224 * 1024) / cta_smem));
    if (!big) {
        if (const char *e = getenv("VLR_POST_MINB")) minb = std::max(1, std::min(minb, atoi(e)));
    }
    const int grid = ctx->sm_count * minb;
    const int64_t n_teams = (int64_t)grid * teams_per_cta;
    if (tbound > ((int64_t)1 << 24))
        return set_err(ctx, VLR_ERR_UNSUPPORTED, "likelihood table of %lld entries per site is too large", (long long)tbound);
    // the kernel enumerates a site's terms (event, bias cfg, path, grid combination) with 32-bit
    // counters: bound their number for the deepest pileup the kernel admits
    {
        double terms = 0.0;
        for (const DevSeg &sg : H.segs) {
            double c = 1.0;
            for (int smp = 0; smp < model->n_samples; ++smp)
                c *= sg.nv[smp] < 0 ? (double)grid_bound(model, smp, obs_cap) : (double)sg.nv[smp];
            terms += c;
        }
        if (terms > 2147483647.0)
            return set_err(ctx, VLR_ERR_UNSUPPORTED, "up to %.3g integration terms per site (limit 2^31 - 1): lower the "
                                                     "resolutions of the samples bound by Ranges", terms);
    }
    // prior: a table over the leaves (set-only universes) or a program evaluated at every leaf
    const bool has_prior = model->prior_ln != nullptr && !(H.any_range_path && model->prior_prog);
    const vlr_prior_prog *prog = has_prior ? nullptr : model->prior_prog;
    if (has_prior) {
        if (H.any_range_path)
            return set_err(ctx, VLR_ERR_UNSUPPORTED, "a prior table needs every tree path to bind VAF sets only "
                                                     "(grid points of a Range depend on the site's depth): pass prior_prog");
        if (model->n_prior != H.n_leaves)
            return set_err(ctx, VLR_ERR_INVALID, "prior table has %lld entries, the model has %lld leaves",
                           (long long)model->n_prior, (long long)H.n_leaves);
    }
    int64_t n_uni = 0;
    if (prog) {
        int64_t want = 1;
        for (int s = 0; s < model->n_samples; ++s) {
            const int k = prog->kind[s];
            if (k < VLR_PRIOR_UNIFORM || k > VLR_PRIOR_GERMLINE || prog->som_kind[s] < VLR_SOM_NONE ||
                prog->som_kind[s] > VLR_SOM_CLONAL_SAME || (k != VLR_PRIOR_UNIFORM && (prog->ploidy[s] < 0 || prog->ploidy[s] > 16)) ||
                (prog->som_kind[s] >= VLR_SOM_CLONAL_DENOVO && (prog->parent[s] < 0 || prog->parent[s] >= model->n_samples)) ||
                prog->uni_off[s + 1] < prog->uni_off[s] || prog->uni_off[0] != 0)
                return set_err(ctx, VLR_ERR_INVALID, "prior program: bad entry for sample %d", s);
            want *= (k == VLR_PRIOR_UNIFORM || prog->ploidy[s] <= 0) ? 1 : prog->ploidy[s] + 1;
        }
        n_uni = prog->uni_off[model->n_samples];
        if (want != prog->n_germ || !prog->germ_ln || (n_uni > 0 && !prog->uni_spec) || want > (1 << 20))
            return set_err(ctx, VLR_ERR_INVALID, "prior program: germline table has %lld entries, expected %lld",
                           (long long)prog->n_germ, (long long)want);
    }
    enum { S_MODEL = 16, S_TABLE, S_CNT };
    // model blob
    const size_t b_spec = H.spectra.size() * sizeof(DevSpectrum), b_path = H.paths.size() * sizeof(DevPath),
                 b_ev = H.events.size() * sizeof(DevEvent), b_eb = H.eb.size() * sizeof(DevEventBias),
                 b_cfg = H.cfg.size() * sizeof(vlr_biases), b_vaf = H.vafs.size() * sizeof(double),
                 b_seg = H.segs.size() * sizeof(DevSeg),
                 b_pri = has_prior ? (size_t)H.n_leaves * 8
                                   : (prog ? (size_t)(prog->n_germ + 5 * n_uni) * 8 + align_up(sizeof(vlr_prior_prog), 16) : 0);
    size_t off[9];
    off[0] = 0;
    off[1] = align_up(off[0] + b_spec, 16);
    off[2] = align_up(off[1] + b_path, 16);
    off[3] = align_up(off[2] + b_ev, 16);
    off[4] = align_up(off[3] + b_eb, 16);
    off[5] = align_up(off[4] + b_cfg, 16);
    off[6] = align_up(off[5] + b_vaf, 16);
    off[7] = align_up(off[6] + b_seg, 16);
    off[8] = align_up(off[7] + b_pri, 16);
    uint8_t *d_blob = (uint8_t *)dev_scratch(ctx, S_MODEL, off[8] + 16);
    double *d_table = (double *)dev_scratch(ctx, S_TABLE, (size_t)((tab_cap ? 0 : n_teams * tbound) + 8) * sizeof(double));
    int32_t *d_cnt = (int32_t *)dev_scratch(ctx, S_CNT, 256);
    if (!d_blob || !d_table || !d_cnt) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
    {
        std::vector<uint8_t> blob(off[8], 0);
        memcpy(blob.data() + off[0], H.spectra.data(), b_spec);
        memcpy(blob.data() + off[1], H.paths.data(), b_path);
        memcpy(blob.data() + off[2], H.events.data(), b_ev);
        memcpy(blob.data() + off[3], H.eb.data(), b_eb);
        memcpy(blob.data() + off[4], H.cfg.data(), b_cfg);
        memcpy(blob.data() + off[5], H.vafs.data(), b_vaf);
        memcpy(blob.data() + off[6], H.segs.data(), b_seg);
        if (has_prior) memcpy(blob.data() + off[7], model->prior_ln, b_pri);
        if (prog) {  // the program with device pointers, the germline table, the universes of the uniform samples
            const size_t hdr = align_up(sizeof(vlr_prior_prog), 16);
            vlr_prior_prog dp = *prog;
            dp.germ_ln = (const double *)(d_blob + off[7] + hdr);
            dp.uni_spec = dp.germ_ln + prog->n_germ;
            memcpy(blob.data() + off[7], &dp, sizeof(dp));
            memcpy(blob.data() + off[7] + hdr, prog->germ_ln, (size_t)prog->n_germ * 8);
            if (n_uni > 0) memcpy(blob.data() + off[7] + hdr + (size_t)prog->n_germ * 8, prog->uni_spec, (size_t)n_uni * 40);
        }
        // upload only when the model changed: the chunks of a batch and repeated calls of one
        // scenario reuse the device copy, and the launch needs no synchronisation with the stream
        if (ctx->model_dev != d_blob || ctx->model_copy != blob) {
            uint8_t *h_blob = (uint8_t *)pin_scratch(ctx, 4, off[8] + 16);
            if (!h_blob) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
            // the pinned staging copy may still be in flight from a previous upload on this stream
            VLR_CUDA(ctx, cudaStreamSynchronize(st));
            memcpy(h_blob, blob.data(), off[8]);
            VLR_CUDA(ctx, cudaMemcpyAsync(d_blob, h_blob, off[8], cudaMemcpyHostToDevice, st));
            ctx->model_copy.swap(blob);
            ctx->model_dev = d_blob;
        }
    }
    VLR_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, 256, st));
    PostArgs a;
    memset(&a, 0, sizeof(a));
    a.m.n_samples = model->n_samples;
    a.m.n_events = model->n_events;
    a.m.n_spectra = (int)H.spectra.size();
    a.m.n_paths = (int)H.paths.size();
    a.m.n_cfg = (int)H.cfg.size();
    a.m.n_eb = (int)H.eb.size();
    for (int s = 0; s < model->n_samples; ++s) {
        a.m.resolution[s] = model->resolution[s];
        a.m.contaminated[s] = model->contaminated[s];
        a.m.by[s] = model->by[s];
        a.m.purity[s] = model->purity[s];
        if (model->resolution[s] < 1) return set_err(ctx, VLR_ERR_INVALID, "resolution must be >= 1");
    }
    a.m.spectra = (const DevSpectrum *)(d_blob + off[0]);
    a.m.paths = (const DevPath *)(d_blob + off[1]);
    a.m.events = (const DevEvent *)(d_blob + off[2]);
    a.m.eb = (const DevEventBias *)(d_blob + off[3]);
    a.m.cfg = (const vlr_biases *)(d_blob + off[4]);
    a.m.vafs = (const double *)(d_blob + off[5]);
    a.m.segs = (const DevSeg *)(d_blob + off[6]);
    a.m.n_segs = (int)H.segs.size();
    a.m.prior = has_prior ? (const double *)(d_blob + off[7]) : nullptr;
    a.m.prog = prog ? (const vlr_prior_prog *)(d_blob + off[7]) : nullptr;
    a.obs = *d_obs;
    a.obs_off = d_obs_off;
    a.n_sites = n_sites;
    a.cursor = d_cnt;
    a.status = d_cnt + 1;
    a.table = d_table;
    a.table_stride = tbound;
    a.out_event_ln = d_out_event_ln;
    a.out_marginal = d_out_marginal;
    a.out_map_af = d_out_map_af;
    a.out_map_bias = d_out_map_bias;
    a.obs_cap = obs_cap;
    a.gs_cap = gs_cap;
    a.tab_cap = tab_cap;
    a.seg_cap = seg_cap;
    a.val_cap = val_cap;
    {
        const int NS = model->n_samples;
        const int team = big ? 4 : 1;
        const int mb = big ? 3 : (minb >= 7 ? 7 : (minb >= 6 ? 6 : (minb >= 5 ? 5 : 3)));
        bool any_cont = false;
        for (int k = 0; k < NS; ++k) any_cont |= model->contaminated[k] != 0;
        // instantiations: without contamination 1..3 samples; with contamination 2 samples or generic
        int nsc, cont;
        if (big) { cont = 1; nsc = NS == 2 ? 2 : 0; }
        else if (!any_cont && NS <= 3) { cont = 0; nsc = NS; }
        else { cont = 1; nsc = NS == 2 ? 2 : 0; }
        cudaError_t le = cudaErrorInvalidValue;
#define VLR_POST_TRY(TEAM, MINB, NSC, CONT) \
        if (team == TEAM && mb == MINB && nsc == NSC && cont == CONT)  \
            le = VLR_POSTERIOR_LAUNCHER(TEAM, MINB, NSC, CONT)(grid, dyn_smem, st, a);
        VLR_POSTERIOR_FOR_ALL(VLR_POST_TRY)
#undef VLR_POST_TRY
        if (le != cudaSuccess)
            return set_err(ctx, VLR_ERR_CUDA, "posterior kernel launch failed: %s", cudaGetErrorString(le));
        ctx->launches++;
    }
    return VLR_OK;
}

int vlr_model_leaves(vlr_ctx *ctx, const vlr_model *model, int64_t capacity, int64_t *out_n_leaves,
                     int32_t *out_event, double *out_afs) {
    if (!ctx) return VLR_ERR_INVALID;
    if (!model || !out_n_leaves) return set_err(ctx, VLR_ERR_INVALID, "bad argument");
    HostModel H;
    int rc = flatten_model(ctx, model, H);
    if (rc != VLR_OK) return rc;
    if (H.any_range_path)
        return set_err(ctx, VLR_ERR_UNSUPPORTED, "leaves are enumerable only when every tree path binds VAF sets");
    *out_n_leaves = H.n_leaves;
    if (!out_event && !out_afs) return VLR_OK;
    if (capacity < H.n_leaves) return set_err(ctx, VLR_ERR_INVALID, "capacity %lld < %lld leaves", (long long)capacity, (long long)H.n_leaves);
    const int NS = model->n_samples;
    for (int e = 0; e < model->n_events; ++e)
        for (int pth = 0; pth < H.events[e].n_paths; ++pth) {
            const int pi = H.events[e].path_off + pth;
            int64_t combos = 1;
            for (int smp = 0; smp < NS; ++smp) combos *= H.spectra[H.paths[pi].spec[smp]].n_vafs;
            for (int64_t combo = 0; combo < combos; ++combo) {
                const int64_t leaf = H.path_leaf_off[pi] + combo;
                if (out_event) out_event[leaf] = e;
                int64_t rem = combo;
                for (int smp = NS - 1; smp >= 0; --smp) {  // LAST sample is the fastest digit
                    const DevSpectrum &sp = H.spectra[H.paths[pi].spec[smp]];
                    if (out_afs) out_afs[leaf * NS + smp] = H.vafs[sp.vaf_off + rem % sp.n_vafs];
                    rem /= sp.n_vafs;
                }
            }
        }
    return VLR_OK;
}

// ---- host-buffer pipeline: the batch is cut into chunks of sites; H2D of chunk c+1 (copy
// stream) overlaps the kernels and D2H of chunk c (main stream), two sets of device buffers.
// T = double: the nine columns as given.  T = float: the seven stored columns (de-quantised
// MiniLogProb values are exactly representable in f32), widened on the device, where the two
// derived columns are computed as ObservationBuilder does (observation.rs:218-226).
}  // extern "C"

namespace vlr {

__global__ void widen_obs_kernel(int64_t n, const float *pm, const float *pa, const float *pr, const float *pmiss,
                                 const float *psa, const float *pdo, const float *phb, double *o_pm,
                                 double *o_pmm, double *o_pa, double *o_pr, double *o_pmiss, double *o_psa,
                                 double *o_pdo, double *o_pso, double *o_phb) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const double m = (double)pm[k], d = (double)pdo[k];
        o_pm[k] = m;
        o_pmm[k] = ln_one_minus_exp_d(m);
        o_pa[k] = (double)pa[k];
        o_pr[k] = (double)pr[k];
        o_pmiss[k] = (double)pmiss[k];
        o_psa[k] = (double)psa[k];
        o_pdo[k] = d;
        o_pso[k] = ln_one_minus_exp_d(d);
        o_phb[k] = (double)phb[k];
    }
}

template <typename T>
static int posterior_pipeline(vlr_ctx *ctx, int64_t n_sites, const vlr_model *model, const T *const *cols,
                              int n_cols_in, const uint16_t *cat, const int64_t *obs_off,
                              double *out_event_ln, double *out_marginal, double *out_map_af,
                              int32_t *out_map_bias) {
    if (!model || !cols || !cat || !obs_off || !out_event_ln || !out_marginal || !out_map_af || !out_map_bias ||
        n_sites < 0)
        return set_err(ctx, VLR_ERR_INVALID, "bad argument");
    if (n_sites == 0) return VLR_OK;
    VLR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int NS = model->n_samples, NE = model->n_events;
    if (NS < 1 || NS > kMaxSamples) return set_err(ctx, VLR_ERR_UNSUPPORTED, "n_samples must be 1..8");
    if (NE < 1 || NE > kMaxEvents) return set_err(ctx, VLR_ERR_UNSUPPORTED, "n_events must be 1..%d", kMaxEvents);
    const int64_t n_off = n_sites * NS + 1;
    const int64_t n_rows = obs_off[n_off - 1] - obs_off[0];
    int64_t max_obs = 0;
    for (int64_t k = 0; k + 1 < n_off; ++k) {
        const int64_t l = obs_off[k + 1] - obs_off[k];
        if (l < 0) return set_err(ctx, VLR_ERR_INVALID, "obs_off not monotone");
        if (l > max_obs) max_obs = l;
    }
    if (max_obs > 4096) return set_err(ctx, VLR_ERR_UNSUPPORTED, "pileup of %lld observations exceeds 4096", (long long)max_obs);
    for (int c = 0; c < n_cols_in; ++c)
        if (!cols[c] && n_rows > 0) return set_err(ctx, VLR_ERR_INVALID, "null observation column");
    int rc = ensure_copy_stream(ctx);
    if (rc != VLR_OK) return rc;
    // ---- chunk boundaries (in sites), ~24 MB of input per chunk
    const double in_bytes = (double)n_rows * (n_cols_in * sizeof(T) + 2) + (double)n_off * 8;
    int n_chunks = (int)std::min<double>(64.0, std::max<double>(1.0, ceil(in_bytes / (24.0 * 1048576.0))));
    if ((int64_t)n_chunks > n_sites) n_chunks = (int)n_sites;
    std::vector<int64_t> cs((size_t)n_chunks + 1);
    cs[0] = 0;
    for (int c = 1; c < n_chunks; ++c) {
        const int64_t want = obs_off[0] + n_rows * c / n_chunks;
        int64_t lo = cs[c - 1] + 1, hi = n_sites - (n_chunks - c);  // keep every chunk non-empty
        while (lo < hi) {
            const int64_t mid = (lo + hi) / 2;
            if (obs_off[mid * NS] < want) lo = mid + 1; else hi = mid;
        }
        cs[c] = lo;
    }
    cs[n_chunks] = n_sites;
    int64_t max_rows = 0, max_sites = 0;
    for (int c = 0; c < n_chunks; ++c) {
        max_rows = std::max(max_rows, obs_off[cs[c + 1] * NS] - obs_off[cs[c] * NS]);
        max_sites = std::max(max_sites, cs[c + 1] - cs[c]);
    }
    // ---- two sets of device buffers
    constexpr bool kWiden = sizeof(T) == 4;
    enum { S_BASE = 100, S_PER = 24 };  // per set: 0..8 fp64 columns, 9 cat, 10 off, 11..14 outputs, 15..21 f32 staging
    double *d_cols[2][9];
    T *d_in[2][9];
    uint16_t *d_cat[2];
    int64_t *d_off[2];
    double *d_ev[2], *d_marg[2], *d_af[2];
    int32_t *d_mb[2];
    for (int b = 0; b < 2; ++b) {
        const int base = S_BASE + b * S_PER;
        for (int c = 0; c < 9; ++c) {
            d_cols[b][c] = (double *)dev_scratch(ctx, base + c, (size_t)(max_rows + 1) * 8);
            if (!d_cols[b][c]) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
        }
        for (int c = 0; c < n_cols_in; ++c) {
            d_in[b][c] = kWiden ? (T *)dev_scratch(ctx, base + 15 + c, (size_t)(max_rows + 1) * sizeof(T))
                                : (T *)d_cols[b][c];
            if (!d_in[b][c]) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
        }
        d_cat[b] = (uint16_t *)dev_scratch(ctx, base + 9, (size_t)(max_rows + 1) * 2);
        d_off[b] = (int64_t *)dev_scratch(ctx, base + 10, (size_t)(max_sites * NS + 1) * 8);
        d_ev[b] = (double *)dev_scratch(ctx, base + 11, (size_t)max_sites * NE * 8);
        d_marg[b] = (double *)dev_scratch(ctx, base + 12, (size_t)max_sites * 8);
        d_af[b] = (double *)dev_scratch(ctx, base + 13, (size_t)max_sites * NS * 8);
        d_mb[b] = (int32_t *)dev_scratch(ctx, base + 14, (size_t)max_sites * NS * 4);
        if (!d_cat[b] || !d_off[b] || !d_ev[b] || !d_marg[b] || !d_af[b] || !d_mb[b])
            return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
    }
    // earlier work of this context may still read the buffers the copy stream is about to overwrite
    VLR_CUDA(ctx, cudaStreamSynchronize(st));
    int32_t *h_status = (int32_t *)pin_scratch(ctx, 5, (size_t)n_chunks * 4 + 16);
    if (!h_status) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
    memset(h_status, 0, (size_t)n_chunks * 4);
    cudaStream_t cp = ctx->copy_stream;
    for (int c = 0; c <= n_chunks; ++c) {
        if (c < n_chunks) {  // H2D of chunk c
            const int b = c & 1;
            const int64_t s0 = cs[c], s1 = cs[c + 1];
            const int64_t r0 = obs_off[s0 * NS], rows = obs_off[s1 * NS] - r0;
            if (c >= 2) VLR_CUDA(ctx, cudaStreamWaitEvent(cp, ctx->ev_done[b], 0));
            for (int k = 0; k < n_cols_in && rows > 0; ++k)
                VLR_CUDA(ctx, cudaMemcpyAsync(d_in[b][k], cols[k] + r0, (size_t)rows * sizeof(T), cudaMemcpyHostToDevice, cp));
            if (rows > 0) VLR_CUDA(ctx, cudaMemcpyAsync(d_cat[b], cat + r0, (size_t)rows * 2, cudaMemcpyHostToDevice, cp));
            VLR_CUDA(ctx, cudaMemcpyAsync(d_off[b], obs_off + s0 * NS, (size_t)((s1 - s0) * NS + 1) * 8, cudaMemcpyHostToDevice, cp));
            VLR_CUDA(ctx, cudaEventRecord(ctx->ev_h2d[b], cp));
        }
        if (c >= 1) {  // kernels + D2H of chunk c-1
            const int b = (c - 1) & 1;
            const int64_t s0 = cs[c - 1], s1 = cs[c];
            const int64_t r0 = obs_off[s0 * NS], rows = obs_off[s1 * NS] - r0;
            VLR_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_h2d[b], 0));
            if (kWiden && rows > 0) {
                const float *const *f = (const float *const *)d_in[b];
                const int grid = (int)std::min<int64_t>((rows + 255) / 256, (int64_t)ctx->sm_count * 8);
                widen_obs_kernel<<<grid, 256, 0, st>>>(rows, f[0], f[1], f[2], f[3], f[4], f[5], f[6], d_cols[b][0],
                                                       d_cols[b][1], d_cols[b][2], d_cols[b][3], d_cols[b][4],
                                                       d_cols[b][5], d_cols[b][6], d_cols[b][7], d_cols[b][8]);
                VLR_LAUNCH_CHECK(ctx);
            }
            vlr_obs_columns dc;  // offsets stay absolute: rebase the device pointers
            dc.prob_mapping = d_cols[b][0] - r0; dc.prob_mismapping = d_cols[b][1] - r0; dc.prob_alt = d_cols[b][2] - r0;
            dc.prob_ref = d_cols[b][3] - r0; dc.prob_missed_allele = d_cols[b][4] - r0; dc.prob_sample_alt = d_cols[b][5] - r0;
            dc.prob_double_overlap = d_cols[b][6] - r0; dc.prob_single_overlap = d_cols[b][7] - r0;
            dc.prob_hit_base = d_cols[b][8] - r0; dc.cat = d_cat[b] - r0;
            rc = vlr_posterior_batch_dev(ctx, s1 - s0, model, &dc, d_off[b], (int32_t)max_obs, d_ev[b], d_marg[b],
                                         d_af[b], d_mb[b]);
            if (rc != VLR_OK) return rc;
            // number of sites of this chunk that exceeded a kernel limit (their outputs are NaN)
            VLR_CUDA(ctx, cudaMemcpyAsync(&h_status[c - 1], (const int32_t *)dev_scratch(ctx, 18, 256) + 1, 4,
                                          cudaMemcpyDeviceToHost, st));
            const int64_t ns_ = s1 - s0;
            VLR_CUDA(ctx, cudaMemcpyAsync(out_event_ln + s0 * NE, d_ev[b], (size_t)ns_ * NE * 8, cudaMemcpyDeviceToHost, st));
            VLR_CUDA(ctx, cudaMemcpyAsync(out_marginal + s0, d_marg[b], (size_t)ns_ * 8, cudaMemcpyDeviceToHost, st));
            VLR_CUDA(ctx, cudaMemcpyAsync(out_map_af + s0 * NS, d_af[b], (size_t)ns_ * NS * 8, cudaMemcpyDeviceToHost, st));
            VLR_CUDA(ctx, cudaMemcpyAsync(out_map_bias + s0 * NS, d_mb[b], (size_t)ns_ * NS * 4, cudaMemcpyDeviceToHost, st));
            VLR_CUDA(ctx, cudaEventRecord(ctx->ev_done[b], st));
        }
    }
    VLR_CUDA(ctx, cudaStreamSynchronize(st));
    int64_t n_over = 0;
    for (int c = 0; c < n_chunks; ++c) n_over += h_status[c];
    if (n_over)
        return set_err(ctx, VLR_ERR_UNSUPPORTED, "%lld site(s) exceeded a kernel limit (observations per pileup, "
                       "allele-frequency values or likelihood-table entries per site); their outputs are NaN",
                       (long long)n_over);
    return VLR_OK;
}

}  // namespace vlr

extern "C" {

static int vlr_posterior_batch_impl(vlr_ctx *ctx, int64_t n_sites, const vlr_model *model,
                        const vlr_obs_columns *obs, const int64_t *obs_off, double *out_event_ln,
                        double *out_marginal, double *out_map_af, int32_t *out_map_bias) {
    if (!ctx) return VLR_ERR_INVALID;
    if (!obs) return set_err(ctx, VLR_ERR_INVALID, "bad argument");
    const double *cols[9] = {obs->prob_mapping, obs->prob_mismapping, obs->prob_alt, obs->prob_ref,
                             obs->prob_missed_allele, obs->prob_sample_alt, obs->prob_double_overlap,
                             obs->prob_single_overlap, obs->prob_hit_base};
    return posterior_pipeline<double>(ctx, n_sites, model, cols, 9, obs->cat, obs_off, out_event_ln, out_marginal,
                                      out_map_af, out_map_bias);
}

int vlr_posterior_batch(vlr_ctx *ctx, int64_t n_sites, const vlr_model *model,
                        const vlr_obs_columns *obs, const int64_t *obs_off, double *out_event_ln,
                        double *out_marginal, double *out_map_af, int32_t *out_map_bias) {
    return vlr::drain_on_error(ctx, vlr_posterior_batch_impl(ctx, n_sites, model, obs, obs_off, out_event_ln, out_marginal, out_map_af, out_map_bias));
}


static int vlr_posterior_batch_f32_impl(vlr_ctx *ctx, int64_t n_sites, const vlr_model *model,
                            const vlr_obs_columns_f32 *obs, const int64_t *obs_off, double *out_event_ln,
                            double *out_marginal, double *out_map_af, int32_t *out_map_bias) {
    if (!ctx) return VLR_ERR_INVALID;
    if (!obs) return set_err(ctx, VLR_ERR_INVALID, "bad argument");
    const float *cols[7] = {obs->prob_mapping, obs->prob_alt, obs->prob_ref, obs->prob_missed_allele,
                            obs->prob_sample_alt, obs->prob_double_overlap, obs->prob_hit_base};
    return posterior_pipeline<float>(ctx, n_sites, model, cols, 7, obs->cat, obs_off, out_event_ln, out_marginal,
                                     out_map_af, out_map_bias);
}

int vlr_posterior_batch_f32(vlr_ctx *ctx, int64_t n_sites, const vlr_model *model,
                            const vlr_obs_columns_f32 *obs, const int64_t *obs_off, double *out_event_ln,
                            double *out_marginal, double *out_map_af, int32_t *out_map_bias) {
    return vlr::drain_on_error(ctx, vlr_posterior_batch_f32_impl(ctx, n_sites, model, obs, obs_off, out_event_ln, out_marginal, out_map_af, out_map_bias));
}


}  // extern "C"

// ---------------------------------------------------------------- Likelihood::compute work items
// One warp per (site, sample, AF[, secondary AF], bias configuration) item: the literal log-space
// formulas of likelihood.rs (this entry point is the per-call seam; the batched table path above
// is the fast one).
namespace vlr {
__global__ void likelihood_items_kernel(vlr_obs_columns obs, const int64_t *obs_off, int32_t n_samples,
                                        const vlr_biases *biases, const vlr_lh_item *items,
                                        int64_t n_items, double *out) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t it = w; it < n_items; it += nw) {
        const vlr_lh_item item = items[it];
        const int64_t o0 = obs_off[(int64_t)item.site * n_samples + item.sample];
        const int64_t o1 = obs_off[(int64_t)item.site * n_samples + item.sample + 1];
        const vlr_biases b = biases[item.bias];
        double acc = 0.0;
        for (int64_t row = o0 + lane; row < o1; row += 32)
            acc += ln_T_exact(obs, row, b, item.other_rate, item.af, item.af_secondary,
                              item.contaminated != 0, item.purity);
        acc = warp_sum(acc);
        if (lane == 0) out[it] = acc;
    }
}
}  // namespace vlr

static int vlr_likelihood_batch_impl(vlr_ctx *ctx, int64_t n_sites, int32_t n_samples,
                                    const vlr_obs_columns *obs, const int64_t *obs_off,
                                    const vlr_biases *biases, int32_t n_biases,
                                    const vlr_lh_item *items, int64_t n_items, double *out_ln_lh) {
    if (!ctx) return VLR_ERR_INVALID;
    if (!obs || !obs_off || !biases || !items || !out_ln_lh || n_sites <= 0 || n_samples <= 0 ||
        n_biases <= 0 || n_items < 0)
        return set_err(ctx, VLR_ERR_INVALID, "bad argument");
    if (n_items == 0) return VLR_OK;
    VLR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    for (int64_t k = 0; k < n_items; ++k) {
        const vlr_lh_item &it = items[k];
        if (it.site < 0 || it.site >= n_sites || it.sample < 0 || it.sample >= n_samples ||
            it.bias < 0 || it.bias >= n_biases)
            return set_err(ctx, VLR_ERR_INVALID, "item %lld out of range", (long long)k);
        if (it.contaminated && !(it.purity > 0.0 && it.purity <= 1.0))
            return set_err(ctx, VLR_ERR_INVALID, "item %lld: purity must be in (0,1]", (long long)k);
    }
    const int64_t n_off = n_sites * n_samples + 1;
    const int64_t r0 = obs_off[0], n_rows = obs_off[n_off - 1] - r0;
    enum { S_COL0 = 20, S_CAT = 29, S_OFF, S_B = 36, S_IT, S_O };
    const double *cols[9] = {obs->prob_mapping, obs->prob_mismapping, obs->prob_alt, obs->prob_ref,
                             obs->prob_missed_allele, obs->prob_sample_alt, obs->prob_double_overlap,
                             obs->prob_single_overlap, obs->prob_hit_base};
    if (n_rows < 0) return set_err(ctx, VLR_ERR_INVALID, "obs_off not monotone");
    for (int c = 0; c < 9; ++c)
        if (n_rows > 0 && (!cols[c] || !obs->cat)) return set_err(ctx, VLR_ERR_INVALID, "null observation column");
    double *d_cols[9];
    for (int c = 0; c < 9; ++c) {
        d_cols[c] = (double *)dev_scratch(ctx, S_COL0 + c, (size_t)(n_rows + 1) * 8);
        if (!d_cols[c]) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
        if (n_rows > 0)
            VLR_CUDA(ctx, cudaMemcpyAsync(d_cols[c], cols[c] + r0, (size_t)n_rows * 8, cudaMemcpyHostToDevice, st));
    }
    uint16_t *d_cat = (uint16_t *)dev_scratch(ctx, S_CAT, (size_t)(n_rows + 1) * 2);
    int64_t *d_off = (int64_t *)dev_scratch(ctx, S_OFF, (size_t)n_off * 8);
    vlr_biases *d_b = (vlr_biases *)dev_scratch(ctx, S_B, (size_t)n_biases * sizeof(vlr_biases));
    vlr_lh_item *d_it = (vlr_lh_item *)dev_scratch(ctx, S_IT, (size_t)n_items * sizeof(vlr_lh_item));
    double *d_o = (double *)dev_scratch(ctx, S_O, (size_t)n_items * 8);
    if (!d_cat || !d_off || !d_b || !d_it || !d_o) return set_err(ctx, VLR_ERR_NOMEM, "scratch allocation failed");
    if (n_rows > 0)
        VLR_CUDA(ctx, cudaMemcpyAsync(d_cat, obs->cat + r0, (size_t)n_rows * 2, cudaMemcpyHostToDevice, st));
    VLR_CUDA(ctx, cudaMemcpyAsync(d_off, obs_off, (size_t)n_off * 8, cudaMemcpyHostToDevice, st));
    VLR_CUDA(ctx, cudaMemcpyAsync(d_b, biases, (size_t)n_biases * sizeof(vlr_biases), cudaMemcpyHostToDevice, st));
    VLR_CUDA(ctx, cudaMemcpyAsync(d_it, items, (size_t)n_items * sizeof(vlr_lh_item), cudaMemcpyHostToDevice, st));
    vlr_obs_columns dc;
    dc.prob_mapping = d_cols[0] - r0; dc.prob_mismapping = d_cols[1] - r0; dc.prob_alt = d_cols[2] - r0;
    dc.prob_ref = d_cols[3] - r0; dc.prob_missed_allele = d_cols[4] - r0; dc.prob_sample_alt = d_cols[5] - r0;
    dc.prob_double_overlap = d_cols[6] - r0; dc.prob_single_overlap = d_cols[7] - r0;
    dc.prob_hit_base = d_cols[8] - r0; dc.cat = d_cat - r0;
    const int64_t want = (n_items * 32 + 255) / 256;
    const int grid = (int)std::min<int64_t>(want, (int64_t)ctx->sm_count * 16);
    vlr::likelihood_items_kernel<<<grid, 256, 0, st>>>(dc, d_off, n_samples, d_b, d_it, n_items, d_o);
    VLR_LAUNCH_CHECK(ctx);
    VLR_CUDA(ctx, cudaMemcpyAsync(out_ln_lh, d_o, (size_t)n_items * 8, cudaMemcpyDeviceToHost, st));
    VLR_CUDA(ctx, cudaStreamSynchronize(st));
    return VLR_OK;
}

extern "C" int vlr_likelihood_batch(vlr_ctx *ctx, int64_t n_sites, int32_t n_samples,
                                    const vlr_obs_columns *obs, const int64_t *obs_off,
                                    const vlr_biases *biases, int32_t n_biases,
                                    const vlr_lh_item *items, int64_t n_items, double *out_ln_lh) {
    return vlr::drain_on_error(ctx, vlr_likelihood_batch_impl(ctx, n_sites, n_samples, obs, obs_off, biases, n_biases, items, n_items, out_ln_lh));
}

