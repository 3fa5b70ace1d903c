"""The two compute stages of libprosic_b200.fixtures on the CPU oracle (test infrastructure)."""
import numpy as np


class OracleBackend:
    def __init__(self, flags=None, fast=False):
        import libprosic_b200 as P
        import oracle
        self.O, self.gap, self.fast = oracle, P.GapParams().as_ln_array(), fast
        self.flags = P.HMM_DEFAULT if flags is None else flags

    def allele_support(self, jobs):
        out = []
        for j in jobs:
            w = self.O.allele_support_region([np.frombuffer(j["ref"], np.uint8)], [np.frombuffer(a, np.uint8) for a, _ in j["alts"]],
                                             np.frombuffer(j["bases"], np.uint8), np.frombuffer(j["quals"], np.uint8),
                                             self.gap, shrink_extra=[0] + [e for _, e in j["alts"]], flags=self.flags,
                                             fast_mode=self.fast)
            out.append(dict(prob_ref=w["prob_ref"], prob_alt=w["prob_alt"], indel_ops=[int(x) for x in w["indel_ops"]]))
        return out

    def posterior(self, scn, cols, cat, off):
        O = self.O
        spec = O.ModelSpec(scn["n_samples"], scn["resolution"], scn["contaminated"], scn["by"], scn["purity"],
                           scn["nodes"], scn["events"], [O.make_biases(**b) for b in scn["biases"]])
        ns = scn["n_samples"]
        res = [O.model_compute(spec, [O.PileupView(cols, cat, int(off[s * ns + k]), int(off[s * ns + k + 1]))
                                      for k in range(ns)], prior=scn.get("prior_fn"))
               for s in range((len(off) - 1) // ns)]
        return dict(event_ln=np.array([r["event_ln"] for r in res]), marginal=np.array([r["marginal"] for r in res]),
                    map_af=np.array([r["map_af"] for r in res]))

    def model_leaves(self, scn):
        return None, np.zeros((0, scn["n_samples"]))  # the oracle takes the prior as a callback

    def call_records(self, scn, post, prob_mapping, off):
        from oracle import call_records as ocr
        is_art = [any(scn["biases"][b][k] for k in ("strand", "orientation", "read_position", "softclip", "divindel"))
                  for b in (ev["biases"][0] for ev in scn["events"])]
        ns = scn["n_samples"]
        prob, af, dp = ocr.call_record(post["event_ln"][0], post["marginal"][0], post["map_af"][0], is_art,
                                       [prob_mapping[off[m]:off[m + 1]] for m in range(ns)])
        names = [e["name"] for e, a in zip(scn["events"], is_art) if not a] + ["artifact"]
        return dict(prob=prob[None, :], af=af[None, :], dp=dp[None, :], names=names)
