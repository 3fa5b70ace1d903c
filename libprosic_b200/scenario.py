"""Flattened calling scenarios (events -> VAF trees -> bias configurations) in the layout of
include/vlr_gpu.h (vlr_model).  The reference builds these on the host (grammar::Scenario ->
VAFTree, calling.rs:453-526); the GPU path consumes them unchanged, so this module is only a
hand-flattened stand-in for the grammar front end used by tests and
bench.py.

A scenario is a dict:
  n_samples, resolution[], contaminated[], by[], purity[],
  nodes: [{kind, sample, children:[node idx], vafs:[...]}],
  events: [{name, roots:[node idx], biases:[bias idx]}],
  biases: [dict(strand, orientation, read_position, softclip, divindel, min_other_rate)]
Sample indices are alphabetical by sample name (quirk Q12).
"""
NODE_SET, NODE_RANGE, NODE_FALSE, NODE_SKIP = 0, 1, 2, 3


def bias(strand=0, orientation=0, read_position=0, softclip=0, divindel=0, min_other_rate=0.0):
    return dict(strand=strand, orientation=orientation, read_position=read_position,
                softclip=softclip, divindel=divindel, min_other_rate=min_other_rate)


def artifact_biases(consider_strand=True, consider_orientation=True, consider_read_position=True,
                    consider_softclip=True, consider_divindel=False, min_divindel_other_rate=0.25):
    """Biases::all_artifact_combinations (bias/mod.rs:101-179): exactly one artifact component,
    cartesian order strand x orientation x read_position x softclip x divindel."""
    out = []
    S = [0, 1, 2] if consider_strand else [0]
    O = [0, 1, 2] if consider_orientation else [0]
    R = [0, 1] if consider_read_position else [0]
    C = [0, 1] if consider_softclip else [0]
    D = [0, 1] if consider_divindel else [0]
    for s in S:
        for o in O:
            for r in R:
                for c in C:
                    for d in D:
                        if (s != 0) + (o != 0) + (r != 0) + (c != 0) + (d != 0) == 1:
                            out.append(bias(s, o, r, c, d, min_divindel_other_rate if d else 0.0))
    return out


class _Builder:
    def __init__(self):
        self.nodes = []

    def node(self, kind, sample=0, vafs=(), children=()):
        self.nodes.append(dict(kind=kind, sample=sample, vafs=list(vafs), children=list(children)))
        return len(self.nodes) - 1

    def chain(self, specs):
        """specs: list of (sample, kind, vafs) from root to leaf -> root node index."""
        child = None
        for sample, kind, vafs in reversed(specs):
            child = self.node(kind, sample, vafs, [] if child is None else [child])
        return child


def _with_events(b, named_roots, n_samples, artifacts):
    """Event universe of calling.rs:472-504: absent + per scenario event {unbiased, artifact}."""
    biases = [bias()] + artifacts
    events = [dict(name="absent",
                   roots=[b.chain([(s, NODE_SET, [0.0]) for s in range(n_samples)])], biases=[0])]
    for name, roots in named_roots:
        events.append(dict(name=name, roots=roots, biases=[0]))
        if artifacts:
            events.append(dict(name=name, roots=roots, biases=list(range(1, len(biases)))))
    return events, biases


def tumor_normal(purity=0.75, indel=True, tumor_resolution=100, normal_resolution=5):
    """The built-in tumor/normal scenario of cli.rs:918-940.  Samples: normal = 0, tumor = 1
    (alphabetical, Q12); tumor is contaminated by normal with fraction 1 - purity."""
    b = _Builder()
    T, N = 1, 0
    tum = (T, NODE_RANGE, [0.0, 1.0, 1.0, 0.0])          # ]0.0,1.0]
    named = [
        ("somatic_tumor", [b.chain([tum, (N, NODE_SET, [0.0])])]),
        ("somatic_normal", [b.chain([tum, (N, NODE_RANGE, [0.0, 0.5, 1.0, 1.0])])]),  # ]0.0,0.5[
        ("germline_het", [b.chain([tum, (N, NODE_SET, [0.5])])]),
        ("germline_hom", [b.chain([tum, (N, NODE_SET, [1.0])])]),
    ]
    # indels: strand + divindel biases; SNVs: strand, orientation, read position, softclip
    # (calling.rs:398-412)
    arts = (artifact_biases(True, False, False, False, True) if indel
            else artifact_biases(True, True, True, True, False))
    events, biases = _with_events(b, named, 2, arts)
    return dict(n_samples=2, resolution=[normal_resolution, tumor_resolution], contaminated=[0, 1],
                by=[0, 0], purity=[1.0, purity], nodes=b.nodes, events=events, biases=biases,
                sample_names=["normal", "tumor"])


def single_sample(continuous=False, resolution=100, snv=True, with_biases=True):
    """Config C2: one sample, no contamination.  Scenario A (germline): universe {0, .5, 1};
    scenario B (continuous): universe [0,1] integrated on the Simpson grid."""
    b = _Builder()
    if continuous:
        named = [("present", [b.chain([(0, NODE_RANGE, [0.0, 1.0, 1.0, 0.0])])])]
    else:
        named = [("het", [b.chain([(0, NODE_SET, [0.5])])]),
                 ("hom", [b.chain([(0, NODE_SET, [1.0])])])]
    arts = []
    if with_biases:
        arts = (artifact_biases(True, True, True, True, False) if snv
                else artifact_biases(True, False, False, False, True))
    events, biases = _with_events(b, named, 1, arts)
    return dict(n_samples=1, resolution=[resolution], contaminated=[0], by=[0], purity=[1.0],
                nodes=b.nodes, events=events, biases=biases, sample_names=["sample"])


def trio(with_biases=False):
    """Config C5-like: three diploid samples (child = 0, father = 1, mother = 2), AF in {0,.5,1},
    one event per genotype class of the child plus a de-novo event with a branching tree."""
    b = _Builder()
    gts = [0.0, 0.5, 1.0]
    named = []
    # child het/hom with any parental genotype (a Set node with several values per parent)
    for name, caf in (("child_het", 0.5), ("child_hom", 1.0)):
        named.append((name, [b.chain([(0, NODE_SET, [caf]), (1, NODE_SET, gts), (2, NODE_SET, gts)])]))
    # de novo: child het, both parents absent -- written as a branching node to exercise fan-out
    leaf_a = b.node(NODE_SET, 2, [0.0])
    leaf_b = b.node(NODE_FALSE)
    mid = b.node(NODE_SET, 1, [0.0], [leaf_a, leaf_b])
    named.append(("denovo", [b.node(NODE_SKIP, 0, [], [b.node(NODE_SET, 0, [0.5], [mid])])]))
    arts = artifact_biases(True, True, True, True, False) if with_biases else []
    events, biases = _with_events(b, named, 3, arts)
    return dict(n_samples=3, resolution=[100, 100, 100], contaminated=[0, 0, 0], by=[0, 0, 0],
                purity=[1.0, 1.0, 1.0], nodes=b.nodes, events=events, biases=biases,
                sample_names=["child", "father", "mother"])
