"""Host-side restatement of the reference's breakend machinery: the alt-allele
haplotypes of breakend groups, and the variants that are encoded as breakend groups.

  * Breakend::new / from_operations       src/variants/types/breakends.rs:687-850
  * BreakendGroupBuilder::build           breakends.rs:58-139 (enclosable reference interval)
  * upstream_bnd / downstream_bnd         breakends.rs:146-176
  * is_insertion / is_deletion            breakends.rs:188-225
  * Variant::is_valid_evidence            breakends.rs:232-300 (MIN_REF_BASES)
  * SamplingBias                          breakends.rs:352-395
  * Realignable::alt_emission_params      breakends.rs:405-615 (alt allele assembly, one per breakend)
  * BreakendEmissionParams                breakends.rs:641-665 (ref_base ignores ref_offset: the
                                          haplotype is not shrinkable, quirk Q6)
  * Replacement / Duplication / Inversion types/{replacement,duplication,inversion}.rs

The kernels take the assembled alt alleles as ordinary haplotype windows with `shrink_extra < 0`
(include/vlr_gpu.h); nothing here runs on the hot path.  Single-contig events only."""
import re

MIN_REF_BASES = 10
_COMP = bytes.maketrans(b"ACGTNacgtn", b"TGCANtgcan")


def revcomp(seq):
    return bytes(seq).translate(_COMP)[::-1]


class Breakend:
    __slots__ = ("pos", "ref_allele", "replacement", "join", "left_to_right", "id")

    def __init__(self, pos, ref_allele, replacement, join, left_to_right, id_):
        # join: None | (pos, side, revcomp) with side "L" (LeftOfPos) / "R" (RightOfPos)
        self.pos, self.ref_allele, self.replacement = pos, bytes(ref_allele), bytes(replacement)
        self.join, self.left_to_right, self.id = join, left_to_right, id_

    def emits_revcomp(self):
        return self.join is not None and self.join[2]


_RE = re.compile(r"((?P<replacement>[ACGTN]+)|((?P<bracket1>[\]\[])(?P<a1><)?(?P<contig>[^\]\[:>]+)(?P<a2>>)?"
                 r"(:(?P<pos>[0-9]+))?(?P<bracket2>[\]\[])))")
_SINGLE_RE = re.compile(r"(\.(?P<from_right>[ACGTN]+))|((?P<from_left>[ACGTN]+)\.)")


def parse_breakend(contig, pos, ref_allele, spec, id_):
    """Breakend::new (breakends.rs:687-790).  Returns None for breakends into an assembly file."""
    single = list(_SINGLE_RE.finditer(spec))
    ops = list(_RE.finditer(spec))
    if len(single) == 1:
        m = single[0]
        if m.group("from_left"):
            return Breakend(pos, ref_allele.encode(), m.group("from_left").encode(), None, True, id_)
        return Breakend(pos, ref_allele.encode(), m.group("from_right").encode(), None, False, id_)
    if len(ops) != 2:
        raise ValueError("invalid BND record alt %s" % spec)
    replacement, join, left_to_right = None, None, False
    for m in ops:
        if m.group("replacement"):
            if join is None:
                left_to_right = True
            replacement = m.group("replacement").encode()
        else:
            b = m.group("bracket1")
            if b != m.group("bracket2"):
                raise ValueError("invalid BND record alt %s" % spec)
            a1, a2 = m.group("a1") is not None, m.group("a2") is not None
            if a1 and a2:
                return None
            if a1 != a2:
                raise ValueError("invalid BND record alt %s" % spec)
            if m.group("contig") != contig:
                raise NotImplementedError("breakend event across contigs")
            jpos = int(m.group("pos")) - 1
            side = "R" if b == "[" else "L"
            rc = (b != "[") if left_to_right else (b == "[")
            join = (jpos, side, rc)
    return Breakend(pos, ref_allele.encode(), replacement, join, left_to_right, id_)


class BreakendGroup:
    """BreakendGroup of one contig: `breakends` in locus order (the BTreeMap), `loci` in push order."""

    def __init__(self, bnds, report_indel_operations=False):
        self.loci = [(b.pos, b.pos + len(b.ref_allele)) for b in bnds]          # push order
        by_pos = {}
        for b in bnds:                                                             # BTreeMap::insert replaces
            by_pos[b.pos] = b
        self.breakends = [by_pos[p] for p in sorted(by_pos)]
        first, last = self.breakends[0], self.breakends[-1]
        self.enclosable_ref_interval = (first.pos, last.pos + (len(last.ref_allele) if not last.left_to_right else 0))
        self.report_indel_operations = report_indel_operations
        self._alts = {}

    # ---- navigation (breakends.rs:146-176)
    def upstream_bnd(self, pos):
        for b in reversed([b for b in self.breakends if b.pos < pos]):
            if not b.left_to_right:
                return b
        return None

    def downstream_bnd(self, pos):
        for b in [b for b in self.breakends if b.pos >= pos]:
            if b.pos > pos and b.left_to_right:
                return b
        return None

    def breakend_pair(self):
        return (self.breakends[0], self.breakends[1]) if len(self.breakends) == 2 else None

    def is_insertion(self):
        p = self.breakend_pair()
        if p is None:
            return False
        l, r = p
        return (l.pos + 1 == r.pos and not l.emits_revcomp() and not r.emits_revcomp() and l.left_to_right
                and len(l.replacement) > 1 and r.replacement[:-1] == l.replacement[1:] and not r.left_to_right)

    def is_deletion(self):
        p = self.breakend_pair()
        if p is None:
            return False
        l, r = p
        return (len(l.replacement) == 1 and len(r.replacement) == 1 and l.left_to_right and not r.left_to_right
                and not l.emits_revcomp() and not r.emits_revcomp())

    def enclosable_len(self):
        if self.is_deletion():
            l, r = self.breakend_pair()
            return r.pos - l.pos
        if self.is_insertion():
            return len(self.breakends[0].replacement) - 1
        return None

    def feasible_bases(self, read_len, props):
        def enclosable(max_op_len):
            ln = self.enclosable_len()
            return ln is not None and max_op_len is not None and ln <= max_op_len
        if self.is_deletion():
            if enclosable(props["max_del_cigar_len"]):
                return read_len
        elif self.is_insertion():
            if enclosable(props["max_ins_cigar_len"]):
                return read_len
        if props["frac_max_softclip"] is not None:
            return int(read_len * props["frac_max_softclip"])
        return None

    # ---- alt alleles (breakends.rs:405-615): one assembled sequence per breakend, in locus order
    def alt_alleles(self, ref, ref_window):
        key = ref_window
        if key in self._alts:
            return self._alts[key]
        out = []
        for first in self.breakends:
            out.append(self._assemble(first, ref, ref_window))
        self._alts[key] = out
        return out

    def _assemble(self, first, ref, ref_window):
        seq = []          # the VecDeque, as a list (front pushes insert at 0)

        def push(s, front):
            nonlocal seq
            seq = list(s) + seq if front else seq + list(s)

        def prefix_range(b):
            return max(0, b.pos - ref_window), b.pos

        def suffix_range(b, inclusive):
            start = b.pos + (0 if inclusive else 1)
            return start, min(start + ref_window, len(ref))
        if first.left_to_right:
            a, e = prefix_range(first)
            push(ref[a:e], False)
            push(first.replacement, False)
        else:
            a, e = suffix_range(first, False)
            push(ref[a:e], True)
            push(first.replacement, True)
        rc = False
        nxt = first
        visited = set()
        while nxt is not None:
            cur = nxt
            if cur.id in visited:
                if cur.left_to_right:
                    a, e = suffix_range(cur, False)
                    push(ref[a:e], False)
                else:
                    a, e = prefix_range(cur)
                    push(ref[a:e], True)
                break
            visited.add(cur.id)
            left_to_right = (not cur.left_to_right) if rc else cur.left_to_right
            if cur.join is None:
                break  # single breakend: the replacement has been added already
            jpos, side, jrc = cur.join
            if side == "L":
                nxt = self.upstream_bnd(jpos)
                s = ref[(nxt.pos + 1 if nxt is not None else 0):jpos + 1]
            else:
                nxt = self.downstream_bnd(jpos)
                s = ref[jpos:(nxt.pos if nxt is not None else len(ref))]
            if nxt is not None:
                s = (s + nxt.replacement) if side == "R" else (nxt.replacement + s)
            ext_rc = (not jrc) if rc else jrc
            if nxt is not None and len(seq) + len(s) > ref_window:
                nxt = None  # the addition exceeds the window: stop afterwards
            if nxt is not None:
                push(revcomp(s) if ext_rc else s, not left_to_right)
            elif left_to_right:
                t = revcomp(s) if ext_rc else s
                push(t[:min(ref_window, len(t))], False)
            else:
                t = revcomp(s) if ext_rc else s
                push(t[max(0, len(t) - ref_window):], True)
            rc = ext_rc
        return bytes(seq)

    # ---- evidence validity (breakends.rs:232-300)
    def is_valid_evidence(self, ev, overlap, end_pos):
        """ev = (left, right|None); overlap(locus, rec, consider_clips) -> str|None.  Returns the list of
        locus indices the evidence is valid for, or None."""
        lo, hi = self.enclosable_ref_interval

        def valid_ref_bases(r):
            return max(max(0, lo - r["pos"]), max(0, end_pos(r) - hi)) > MIN_REF_BASES
        left, right = ev
        if right is None:
            if not valid_ref_bases(left):
                return None
            idx = [i for i, l in enumerate(self.loci) if overlap(l, left, True) is not None]
        else:
            if not valid_ref_bases(left) and not valid_ref_bases(right):
                return None
            idx = [i for i, l in enumerate(self.loci)
                   if overlap(l, left, True) is not None or overlap(l, right, True) is not None]
        return idx or None


def replacement_group(ref, start, ref_allele, alt_allele):
    """Replacement::new (types/replacement.rs:24-74): interval = start .. start + len(ref_allele)."""
    end = start + len(ref_allele)
    repl = alt_allele.upper().encode() if isinstance(alt_allele, str) else bytes(alt_allele)
    bnds = [Breakend(start, ref[start:start + 1], repl, (end, "R", False), True, "u")]
    if end < len(ref):
        bnds.append(Breakend(end, ref[end:end + 1], repl + ref[end:end + 1], (start - 1, "L", False), False, "w"))
    return BreakendGroup(bnds, report_indel_operations=True)


def duplication_group(ref, start, length):
    """Duplication::new (types/duplication.rs:24-96)."""
    end = start + length
    r = lambda p: ref[p:p + 1]  # noqa: E731
    return BreakendGroup([
        Breakend(start, r(start), r(start), (end - 1, "L", False), False, "u"),
        Breakend(start - 1, r(start - 1), r(start - 1), (start, "R", False), True, "v"),
        Breakend(end - 1, r(end - 1), r(end - 1), (start, "R", False), True, "w"),
        Breakend(end, r(end), r(end), (end - 1, "L", False), False, "x")])


def inversion_group(ref, start, length):
    """Inversion::new (types/inversion.rs:24-94)."""
    end = start + length
    r = lambda p: ref[p:p + 1]  # noqa: E731
    return BreakendGroup([
        Breakend(start - 1, r(start - 1), r(start - 1), (end - 1, "L", True), True, "w"),
        Breakend(start, r(start), r(start), (end, "R", True), False, "v"),
        Breakend(end - 1, r(end - 1), r(end - 1), (start - 1, "L", True), True, "u"),
        Breakend(end, r(end), r(end), (start, "R", True), False, "x")])
