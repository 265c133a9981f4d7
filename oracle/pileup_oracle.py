"""CPU ORACLE (test infrastructure, NOT product code) for the live variant caller hot path.

This file is a pure-Python restatement of the reference algorithm.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it.  The product path (``covid-spings-variant-caller_b200/``) never does.

What it restates (all paths relative to /root/reference):

* ``variant_caller/utils.py:9-24``              -> from_phred_scale / to_phred_scale / genotype_likelihood
* ``variant_caller/live_variant_caller.py:54-103``  -> process_bam / process_pileup_column / process_svn
* ``variant_caller/live_variant_caller.py:120-231`` -> prepare_variants
* ``variant_caller/live_variant_caller.py:233-297`` -> write_vcf (text layout per SURVEY.md A8, [EXT])
* pysam/htslib ``bam_plp`` pileup engine (un-vendored third-party dependency, pysam UNPINNED in
  ``requirements.txt:1``; semantics restated from htslib ``sam.c``: ``bam_plp_push``, ``bam_plp_next``,
  ``resolve_cigar2`` and pysam ``libcalignedsegment.pyx: pileup_base_qual_skip``).

PARITY STATUS
  * GL / record half: PINNED.  ``oracle/gen_golden.py`` runs the REAL reference
    ``LiveVariantCaller.prepare_variants`` / ``process_pileup_column`` and ``utils.py`` (with a stub
    ``pysam`` module, because pysam is not installable here) and the vectors are committed under
    ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file against them.
  * pileup half (CIGAR walk, base-quality rule, max_depth admission): "parity unpinned" -- the
    arithmetic lives in htslib, which is absent from /root/reference and from this image, and the
    reference's own tests pin no result.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

NIBBLE_CHARS = "=ACMGRSVTWYHKDBN"          # htslib seq_nt16_str  [EXT B6]
CHAR_TO_NIBBLE = {c: i for i, c in enumerate(NIBBLE_CHARS)}
CIGAR_OPS = "MIDNSHP=XB"                    # BAM op codes 0..9
FLAG_FILTER = 0x4 | 0x100 | 0x200 | 0x400   # UNMAP|SECONDARY|QCFAIL|DUP  [EXT B1]
MAX_DEPTH = 8000                            # pysam default max_depth      [EXT B1]

# op classes
_REF_OPS = (0, 2, 3, 7, 8)      # M D N = X  consume reference
_QRY_OPS = (0, 1, 4, 7, 8)      # M I S = X  consume query
_MATCH_OPS = (0, 7, 8)


# --------------------------------------------------------------------------------------------
# utils.py restated
# --------------------------------------------------------------------------------------------
def from_phred_scale(score) -> float:
    """utils.py:9-10  ``math.pow(10, score / -10)``."""
    return math.pow(10, score / -10)


def to_phred_scale(probability: float, threshold: int = 99) -> int:
    """utils.py:12-13."""
    return min(round(-10 * math.log10(probability)), threshold) if probability > 0.0 else threshold


def genotype_likelihood(hypothesis: str, alleles: Dict[str, List[float]]):
    """utils.py:16-24: (prod of 1-e over the hypothesis allele) * prod over other alleles (dict order)
    of (prod of e), every product a left-to-right fp64 product (np.prod is sequential)."""
    hyp = _seq_prod([1.0 - e for e in alleles[hypothesis]])
    non = 1.0
    for allele in alleles:
        if allele != hypothesis:
            non = non * _seq_prod(alleles[allele])
    return np.float64(hyp * non)


def _seq_prod(values: Sequence[float]) -> float:
    """np.prod of a 1-D float64 array == plain left-to-right product (re-probed)."""
    p = 1.0
    for v in values:
        p *= v
    return p


# --------------------------------------------------------------------------------------------
# reads
# --------------------------------------------------------------------------------------------
class Read:
    """One alignment record as the pileup engine sees it (pos is 0-based)."""
    __slots__ = ("name", "flag", "pos", "mapq", "cigar", "seq", "qual", "mpos", "mref", "tlen")

    def __init__(self, flag: int, pos: int, mapq: int, cigar: List[Tuple[int, int]], seq: str,
                 qual: Sequence[int], name: str = "r", mpos: int = -1, mref: int = -1, tlen: int = 0):
        self.name = name
        self.mpos = mpos              # PNEXT, 0-based (-1: absent)
        self.mref = mref              # RNEXT: 1 = this contig, 0 = another contig, -1 = absent
        self.tlen = tlen
        self.flag = flag
        self.pos = pos
        self.mapq = mapq
        self.cigar = cigar            # list of (op_code, length)
        self.seq = seq                # str over NIBBLE_CHARS
        self.qual = list(qual)        # phred ints

    def ref_len(self) -> int:
        return sum(l for op, l in self.cigar if op in _REF_OPS)

    def end(self) -> int:
        """htslib bam_endpos: pos + rlen, rlen==0 -> pos+1."""
        r = self.ref_len()
        return self.pos + (r if r > 0 else 1)


def parse_cigar(text: str) -> List[Tuple[int, int]]:
    out, num = [], ""
    if text == "*":
        return out
    for ch in text:
        if ch.isdigit():
            num += ch
        else:
            out.append((CIGAR_OPS.index(ch), int(num)))
            num = ""
    return out


def read_sam(path: str, contig: Optional[str] = None) -> Tuple[List[Tuple[str, int]], List[Read]]:
    """Minimal SAM text reader: returns ([(contig, length)], reads on `contig` (default: first @SQ))."""
    contigs, reads = [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith("@"):
                if line.startswith("@SQ"):
                    f = dict(x.split(":", 1) for x in line.rstrip("\n").split("\t")[1:])
                    contigs.append((f["SN"], int(f["LN"])))
                continue
            t = line.rstrip("\n").split("\t")
            if len(t) < 11:
                continue
            if contig is None:
                contig = contigs[0][0]
            if t[2] != contig:
                continue
            qual = [ord(c) - 33 for c in t[10]] if t[10] != "*" else [255] * len(t[9])
            mref = 1 if t[6] in ("=", t[2]) else (-1 if t[6] == "*" else 0)
            reads.append(Read(int(t[1]), int(t[3]) - 1, int(t[4]), parse_cigar(t[5]), t[9].upper(), qual, t[0],
                              int(t[7]) - 1, mref, int(t[8])))
    return contigs, reads


def samtools_sort(reads: List[Read]) -> List[Read]:
    """pysam.sort order [EXT B7]: by position, forward strand before reverse, then input order."""
    return sorted(reads, key=lambda r: (r.pos, 1 if r.flag & 0x10 else 0))


def passes_read_filter(r: Read, min_mapq: int) -> bool:
    """pysam stepper 'samtools' (__advance_samtools) + htslib push-time skips  [EXT B2]."""
    if r.flag & FLAG_FILTER:
        return False
    if r.mapq < min_mapq:
        return False
    if (r.flag & 0x1) and not (r.flag & 0x2):      # ignore_orphans
        return False
    if not any(op in _REF_OPS for op, _ in r.cigar):
        # htslib resolve_cigar2 asserts on a read with no M/D/N/=/X op; such records are malformed.
        # The restatement (and the product) skip them.
        return False
    return True


# --------------------------------------------------------------------------------------------
# htslib mate-overlap handling (pysam pileup default ignore_overlaps=True)  [EXT B5, parity unpinned]
# --------------------------------------------------------------------------------------------
OVERLAP_OFF, OVERLAP_HTSLIB_1_10, OVERLAP_HTSLIB_1_13 = 0, 1, 2


def _x31_hash_string(name: str) -> int:
    """khash.h __ac_X31_hash_string."""
    b = name.encode("latin-1")
    if not b:
        return 0
    h = b[0]
    for c in b[1:]:
        h = ((h << 5) - h + c) & 0xFFFFFFFF
    return h


def _wang_hash(key: int) -> int:
    """khash.h __ac_Wang_hash (32 bit)."""
    m = 0xFFFFFFFF
    key = (key + (~(key << 15) & m)) & m
    key ^= key >> 10
    key = (key + (key << 3)) & m
    key ^= key >> 6
    key = (key + (~(key << 11) & m)) & m
    key ^= key >> 16
    return key


class _Cur:
    """state of htslib's cigar_iref2iseq_set / cigar_iref2iseq_next over one read"""
    __slots__ = ("cig", "k", "icig", "iseq", "iref")

    def __init__(self, cigar):
        self.cig, self.k, self.icig, self.iseq, self.iref = cigar, 0, 0, 0, 0


def _iref2iseq_set(c: _Cur, want: int) -> int:
    pos = want
    if pos < 0:
        return -1
    c.icig = c.iseq = c.iref = 0
    while c.k < len(c.cig):
        op, n = c.cig[c.k]
        if op == 4:
            c.k += 1; c.iseq += n; c.icig = 0
        elif op in (5, 6):
            c.k += 1; c.icig = 0
        elif op in _MATCH_OPS:
            pos -= n
            if pos < 0:
                c.icig = n + pos; c.iseq += c.icig; c.iref += c.icig
                return 0
            c.k += 1; c.iseq += n; c.icig = 0; c.iref += n
        elif op == 1:
            c.k += 1; c.iseq += n; c.icig = 0
        elif op in (2, 3):
            pos -= n
            if pos < 0:
                pos = 0
            c.k += 1; c.icig = 0; c.iref += n
        else:
            return -2
    c.iseq = -1
    return -1


def _iref2iseq_next(c: _Cur) -> int:
    while c.k < len(c.cig):
        op, n = c.cig[c.k]
        if op in _MATCH_OPS:
            if c.icig >= n - 1:
                c.icig = -1; c.k += 1
                continue
            c.iseq += 1; c.icig += 1; c.iref += 1
            return 0
        if op in (2, 3):
            c.k += 1; c.iref += n; c.icig = -1
        elif op in (1, 4):
            c.k += 1; c.iseq += n; c.icig = -1
        elif op in (5, 6):
            c.k += 1; c.icig = -1
        else:
            return -2
    c.iseq = -1
    c.iref = -1
    return -1


def tweak_overlap_quality(a: Read, b: Read, model: int) -> int:
    """htslib sam.c tweak_overlap_quality(a = buffered first read, b = read being pushed): rewrites a.qual / b.qual
    in place over the reference positions both reads cover.  Returns the number of rewritten positions."""
    ca, cb = _Cur(a.cigar), _Cur(b.cigar)
    iref = b.pos
    a_ret = _iref2iseq_set(ca, iref - a.pos)
    if a_ret < 0:
        return 0
    b_ret = _iref2iseq_set(cb, iref - b.pos)
    if b_ret < 0:
        return 0
    legacy = model == OVERLAP_HTSLIB_1_10
    if legacy:
        amul, bmul = 1, 0
    elif _wang_hash(_x31_hash_string(a.name)) & 1:
        amul, bmul = 1, 0
    else:
        amul, bmul = 0, 1
    touched = 0
    while True:
        while a_ret >= 0 and ca.iref >= 0 and ca.iref < iref - a.pos:
            a_ret = _iref2iseq_next(ca)
        if a_ret < 0:
            break
        if iref < ca.iref + a.pos:
            iref = ca.iref + a.pos
        while b_ret >= 0 and cb.iref >= 0 and cb.iref < iref - b.pos:
            b_ret = _iref2iseq_next(cb)
        if b_ret < 0:
            break
        if iref < cb.iref + b.pos:
            iref = cb.iref + b.pos
        iref += 1
        if ca.iref + a.pos != cb.iref + b.pos:
            if legacy:
                continue
            if ca.iref + a.pos < cb.iref + b.pos and cb.k > 0 and b.cigar[cb.k - 1][0] == 2:
                done = False
                while True:
                    if 0 <= ca.iseq < len(a.qual):
                        a.qual[ca.iseq] = int(a.qual[ca.iseq] * 0.8) if amul else 0
                        touched += 1
                    a_ret = _iref2iseq_next(ca)
                    if a_ret < 0:
                        done = True
                        break
                    if not ca.iref + a.pos < cb.iref + b.pos:
                        break
                if done:
                    return touched
            elif ca.k > 0 and a.cigar[ca.k - 1][0] == 2:
                done = False
                while True:
                    if 0 <= cb.iseq < len(b.qual):
                        b.qual[cb.iseq] = int(b.qual[cb.iseq] * 0.8) if bmul else 0
                        touched += 1
                    b_ret = _iref2iseq_next(cb)
                    if b_ret < 0:
                        done = True
                        break
                    if not cb.iref + b.pos < ca.iref + a.pos:
                        break
                if done:
                    return touched
            else:
                continue
        if ca.iseq < 0 or cb.iseq < 0 or ca.iseq >= len(a.qual) or cb.iseq >= len(b.qual):
            return touched
        qa, qb = a.qual[ca.iseq], b.qual[cb.iseq]
        touched += 1
        if CHAR_TO_NIBBLE.get(a.seq[ca.iseq], 15) == CHAR_TO_NIBBLE.get(b.seq[cb.iseq], 15):
            q = min(qa + qb, 200)
            if legacy:
                qa, qb = q, 0
            else:
                qa, qb = amul * q, bmul * q
        elif legacy:
            if qa >= qb:
                qa, qb = int(0.8 * qa), 0
            else:
                qa, qb = 0, int(0.8 * qb)
        else:
            if qa > qb:
                qa, qb = int(0.8 * qa), 0
            elif qa < qb:
                qa, qb = 0, int(0.8 * qb)
            else:
                qa, qb = int((amul * 0.8) * qa), int((bmul * 0.8) * qb)
        a.qual[ca.iseq], b.qual[cb.iseq] = qa, qb
    return touched


def _query_length(r: Read) -> int:
    return sum(l for op, l in r.cigar if op in _QRY_OPS)


def _overlap_push(olap: Dict[str, "_Node"], nd: "_Node", model: int) -> int:
    """htslib overlap_push for a node that has just been admitted.  Returns rewritten positions."""
    r = nd.read
    if (r.flag & 0x8) or not (r.flag & 0x2):
        return 0
    lq = _query_length(r)
    if model == OVERLAP_HTSLIB_1_10:
        if abs(r.tlen) >= 2 * lq:
            return 0
    else:
        if r.mref == 0 or (abs(r.tlen) >= 2 * lq and r.mpos >= nd.end):
            return 0
    other = olap.get(r.name)
    if other is None:
        if model == OVERLAP_HTSLIB_1_10 or r.mpos >= r.pos or ((r.flag & 0x1) and r.mpos == -1):
            olap[r.name] = nd
        return 0
    del olap[r.name]
    return tweak_overlap_quality(other.read, r, model)


# --------------------------------------------------------------------------------------------
# literal emulation of htslib bam_plp (push / next) -- small inputs only
# --------------------------------------------------------------------------------------------
class _Node:
    __slots__ = ("idx", "read", "beg", "end", "k", "x", "y")

    def __init__(self, idx, read, own_copy=False):
        if own_copy:       # htslib buffers a copy (bam_copy1); the overlap handling rewrites that copy's qualities
            read = Read(read.flag, read.pos, read.mapq, read.cigar, read.seq, list(read.qual), read.name, read.mpos,
                        read.mref, read.tlen)
        self.idx, self.read = idx, read
        self.beg, self.end = read.pos, read.end()
        self.k, self.x, self.y = -1, 0, 0


def _resolve_cigar2(node: _Node, pos: int):
    """htslib resolve_cigar2: returns (is_del, is_refskip, qpos) of `node` at reference `pos`."""
    cig = node.read.cigar
    n = len(cig)
    if node.k == -1:
        # htslib special-cases n_cigar == 1 (and leaves k = -1, i.e. undefined behaviour, for a lone
        # non-match op); the general scan below gives the same answer for every well-formed record.
        node.x, node.y = node.read.pos, 0
        k = 0
        while k < n:
            op, l = cig[k]
            if op in _REF_OPS:
                break
            elif op in (1, 4):
                node.y += l
            k += 1
        assert k < n
        node.k = k
    else:
        op, l = cig[node.k]
        if pos - node.x >= l:
            nop = cig[node.k + 1][0]
            if cig[node.k][0] in _MATCH_OPS:
                node.y += l
            node.x += l
            if nop in _REF_OPS:
                node.k += 1
            else:
                k = node.k + 1
                while k < n:
                    op2, l2 = cig[k]
                    if op2 in _REF_OPS:
                        break
                    elif op2 in (1, 4):
                        node.y += l2
                    k += 1
                node.k = k
            assert node.k < n
    op, l = cig[node.k]
    if op in _MATCH_OPS:
        return False, False, node.y + (pos - node.x)
    return True, op == 3, node.y


def pileup_columns(reads: Iterable[Read], min_mapq: int, max_depth: int = MAX_DEPTH,
                   admitted: Optional[List[int]] = None, overlap_model: int = OVERLAP_OFF,
                   tweaked: Optional[Dict[int, List[int]]] = None
                   ) -> Iterator[Tuple[int, List[Tuple[int, Read, bool, bool, int]]]]:
    """Literal restatement of bam_plp_auto/bam_plp_push/bam_plp_next for ONE contig (tid fixed).

    Yields (pos, [(read_index, read, is_del, is_refskip, qpos), ...]) for every column with >=1 entry,
    BEFORE the pysam base-quality filter.  `admitted`, when given, collects the indices of the reads
    that were admitted into the buffer (read-level filter passed and not dropped by max_depth).
    `overlap_model` != OVERLAP_OFF emulates bam_plp_init_overlaps: the entries then carry the engine's own copies
    of the reads with the rewritten qualities; `tweaked` collects {read index: rewritten quality list}.
    """
    buf: List[_Node] = []          # the linked list head..tail (without the sentinel)
    olap: Dict[str, _Node] = {}    # htslib iter->overlaps: QNAME -> buffered node
    state = {"pos": 0, "max_pos": -1, "eof": False, "started": False}

    def plp_next():
        # htslib bam_plp_next loop; returns a column or None
        while state["eof"] or state["max_pos"] > state["pos"]:
            if state["eof"] and not buf:
                return None
            pos = state["pos"]
            entries = []
            keep = []
            for nd in buf:
                if nd.end <= pos:
                    olap.pop(nd.read.name, None)  # overlap_remove, then freed (mp_free): cnt decreases
                    continue
                if nd.beg <= pos:
                    is_del, is_skip, qpos = _resolve_cigar2(nd, pos)
                    entries.append((nd.idx, nd.read, is_del, is_skip, qpos))
                keep.append(nd)
            buf[:] = keep
            if buf:
                if state["pos"] < buf[0].beg:
                    state["pos"] = buf[0].beg
                else:
                    state["pos"] += 1
            else:
                state["pos"] += 1
            if entries:
                return pos, entries
            if state["eof"] and not buf:
                return None
        return None

    it = enumerate(reads)
    while True:
        col = plp_next()
        if col is not None:
            yield col
            continue
        if state["eof"]:
            return
        # read alignments until a column can be produced
        produced = False
        for idx, r in it:
            if not passes_read_filter(r, min_mapq):
                continue
            if r.pos < state["max_pos"]:
                raise ValueError("reads are not coordinate sorted")
            # bam_plp_push (iter->tid == b->core.tid: single contig, tid 0 -- see DESIGN.md)
            cnt = len(buf) + 1                     # mempool count incl. the tail sentinel
            if r.pos == state["pos"] and cnt > max_depth:
                olap.pop(r.name, None)             # overlap_remove
                continue                           # dropped by maxcnt
            nd = _Node(idx, r, own_copy=overlap_model != OVERLAP_OFF)
            state["max_pos"] = nd.beg
            state["started"] = True
            if nd.end > state["pos"]:
                buf.append(nd)
                if admitted is not None:
                    admitted.append(idx)
                if overlap_model != OVERLAP_OFF:
                    partner = olap.get(nd.read.name)
                    if _overlap_push(olap, nd, overlap_model) and tweaked is not None:
                        tweaked[partner.idx] = partner.read.qual
                        tweaked[idx] = nd.read.qual
            col = plp_next()
            if col is not None:
                yield col
                produced = True
                break
        if not produced:
            state["eof"] = True


def admission_mask(pos: Sequence[int], end: Sequence[int], passes: Sequence[bool],
                   max_depth: int = MAX_DEPTH) -> List[bool]:
    """O(N + span) event simulation of the max_depth rule; must equal the `admitted`
    set of :func:`pileup_columns`.  Used to build the keep-mask for large inputs."""
    n = len(pos)
    keep = [False] * n
    if n == 0:
        return keep
    ends: Dict[int, int] = {}
    nbuf = 0                 # buffered (admitted, not yet swept) reads
    iter_pos = 0
    max_pos = -1
    started = False
    for i in range(n):
        if not passes[i]:
            continue
        p, e = pos[i], end[i]
        if p < max_pos:
            raise ValueError("reads are not coordinate sorted")
        if p == iter_pos and nbuf + 1 > max_depth:
            continue
        max_pos = p
        started = True
        keep[i] = True
        nbuf += 1
        ends[e] = ends.get(e, 0) + 1
        # bam_plp_next: emit columns while max_pos > iter_pos
        while max_pos > iter_pos:
            c = iter_pos
            # sweep nodes with end <= c.  Only ends in (previous c, c] can be non-empty, but a jump may
            # skip columns, so sweep by key.
            if c in ends:
                nbuf -= ends.pop(c)
            # buffer now holds: older reads with end > c, plus the new read (beg > c)
            if nbuf - 1 == 0:
                iter_pos = max_pos         # head is the new read: jump
            else:
                iter_pos = c + 1
        # ends <= iter_pos that were skipped by a jump cannot exist
    return keep


# --------------------------------------------------------------------------------------------
# the reference's state update + genotype stage, restated
# --------------------------------------------------------------------------------------------
class OracleCaller:
    """Restatement of LiveVariantCaller (live_variant_caller.py:21-297) over in-memory reads.

    `reference` is the contig sequence string (what ``fastaFile.fetch(reference=...)`` returns)."""

    def __init__(self, reference: str, minBaseQuality: int, minMappingQuality: int, minTotalDepth: int,
                 minAlleleDepth: int, minEvidenceRatio: float, maxVariants: int = 1,
                 contig: str = "NC_045512.2", max_depth: int = MAX_DEPTH, overlap_model: int = OVERLAP_OFF):
        self.reference = reference
        self.contig = contig
        self.minBaseQuality = minBaseQuality
        self.minMappingQuality = minMappingQuality
        self.minTotalDepth = minTotalDepth
        self.minAlleleDepth = minAlleleDepth
        self.minEvidenceRatio = minEvidenceRatio
        self.maxVariants = maxVariants
        self.max_depth = max_depth
        self.overlap_model = overlap_model     # pysam's default is ignore_overlaps=True (OVERLAP_HTSLIB_*)
        self.memory: Dict[int, dict] = {}

    def reset_memory(self):
        self.memory = {}

    # live_variant_caller.py:54-72 + :74-103
    def process_reads(self, reads: Iterable[Read]):
        for pos, entries in pileup_columns(reads, self.minMappingQuality, self.max_depth,
                                           overlap_model=self.overlap_model):
            # pysam PileupColumn.pileups: drop entries with qual[qpos] < min_base_quality (0 if qpos>=l_qseq)
            kept = []
            for idx, r, is_del, is_skip, qpos in entries:
                q = r.qual[qpos] if qpos < len(r.qual) else 0
                if q < self.minBaseQuality:
                    continue
                kept.append((r, is_del, is_skip, qpos))
            total = len(kept)                                        # :75
            if pos not in self.memory:                                # :77-85
                self.memory[pos] = {"reference": self.reference[pos], "totalDepth": total,
                                    "snvs": {}, "indels": {}}
            else:
                self.memory[pos]["totalDepth"] += total              # :87
            for r, is_del, is_skip, qpos in kept:                     # :89-103
                if not is_del and not is_skip:
                    base = r.seq[qpos]
                    self.memory[pos]["snvs"].setdefault(base, []).append(r.qual[qpos])

    # live_variant_caller.py:120-231 (the indel loop :187-229 never fires: 'indels' is always {})
    def prepare_variants(self) -> List[dict]:
        variants = []
        for position in self.memory:
            site = self.memory[position]
            if site["totalDepth"] >= self.minTotalDepth:
                snvs = {a: [from_phred_scale(q) for q in site["snvs"][a]] for a in site["snvs"]}
                gls = {a: genotype_likelihood(a, snvs) for a in snvs}
                s = 0.0
                for v in gls.values():
                    s = s + v
                s = s if s != 0 else 1.0
                for allele in snvs:
                    ad = len(snvs[allele])
                    if site["reference"] != allele and ad >= self.minAlleleDepth and \
                            ad / site["totalDepth"] >= self.minEvidenceRatio:
                        gl_lin = gls[allele]
                        if gl_lin != 0:
                            gl = math.log10(gl_lin)
                            pl = round(-10.0 * gl)
                        else:
                            gl, pl = 0, 0
                        score = to_phred_scale(1.0 - (gls[allele] / s))
                        qual = np.mean(snvs[allele])
                        variants.append({"start": position, "stop": position + 1,
                                         "alleles": (site["reference"], allele), "qual": qual,
                                         "info": {"DP": site["totalDepth"], "AD": ad, "GL": gl, "PL": pl,
                                                  "SCORE": score}})
        return variants

    def likelihoods(self) -> Dict[int, Dict[str, float]]:
        """L(a) for every allele of every site (no gates); used for the 1e-9 likelihood parity test."""
        out = {}
        for position, site in self.memory.items():
            snvs = {a: [from_phred_scale(q) for q in site["snvs"][a]] for a in site["snvs"]}
            out[position] = {a: float(genotype_likelihood(a, snvs)) for a in snvs}
        return out

    # live_variant_caller.py:233-297, text layout per SURVEY A8 [EXT]
    def vcf_text(self, contigs: Sequence[Tuple[str, int]]) -> str:
        return format_vcf(self.prepare_variants(), contigs)


# --------------------------------------------------------------------------------------------
# VCF text (htslib writer emulation, [EXT] SURVEY A8)
# --------------------------------------------------------------------------------------------
def fmt_g_float32(x) -> str:
    """htslib kputd(float32 value): C '%g' of the value after a round trip through float32."""
    return "%g" % float(np.float32(x))


VCF_INFO_META = [
    ("DP", "Integer", "Total Depth"),
    ("AD", "Integer", "Allele Depth"),
    ("GL", "Float", "Genotype likelihoods comprised of comma separated floating point log10-scaled "
                    "likelihoods for all possible genotypes given the set of alleles defined in the REF "
                    "and ALT fields"),
    ("PL", "Integer", "The phred-scaled genotype likelihoods rounded to the closest integer (and "
                      "otherwise defined precisely as the GL field)"),
    ("SCORE", "Float", "Custom scoring function"),
]


def format_vcf(variants: List[dict], contigs: Sequence[Tuple[str, int]]) -> str:
    lines = ["##fileformat=VCFv4.2", '##FILTER=<ID=PASS,Description="All filters passed">']
    for vid, typ, desc in VCF_INFO_META:
        lines.append(f'##INFO=<ID={vid},Number=1,Type={typ},Description="{desc}">')
    for name, length in contigs:
        lines.append(f"##contig=<ID={name},length={length}>")
    lines.append("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO")
    chrom = contigs[0][0]
    for v in sorted(variants, key=lambda v: (v["start"], v["info"]["SCORE"])):   # :285-286 (stable)
        info = v["info"]
        gl = fmt_g_float32(info["GL"])
        sc = fmt_g_float32(info["SCORE"])
        lines.append(f'{chrom}\t{v["start"] + 1}\t.\t{v["alleles"][0]}\t{v["alleles"][1]}\t'
                     f'{fmt_g_float32(v["qual"])}\t.\t'
                     f'DP={info["DP"]};AD={info["AD"]};GL={gl};PL={info["PL"]};SCORE={sc}')
    return "\n".join(lines) + "\n"


# --------------------------------------------------------------------------------------------
# order-free sufficient statistics (what the device tables must equal)
# --------------------------------------------------------------------------------------------
def tables_from_memory(memory: Dict[int, dict]):
    """memory -> (depth{pos}, hist{(pos, nibble_code, q): count}) for bit-exact table comparison."""
    depth = {p: s["totalDepth"] for p, s in memory.items()}
    hist: Dict[Tuple[int, int, int], int] = {}
    for p, s in memory.items():
        for base, quals in s["snvs"].items():
            code = CHAR_TO_NIBBLE[base]
            for q in quals:
                hist[(p, code, q)] = hist.get((p, code, q), 0) + 1
    return depth, hist
