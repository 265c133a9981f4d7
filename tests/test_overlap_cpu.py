"""Mate-overlap handling (pysam pileup default ignore_overlaps=True; reference call site
live_variant_caller.py:56-60): the product's admission pass (lvc_admit_overlaps, host only) against the oracle's
literal emulation of htslib's overlap hash + tweak_overlap_quality, for both release lines."""
import os
import random

import numpy as np
import pytest

from oracle import pileup_oracle as po

MODELS = [po.OVERLAP_HTSLIB_1_10, po.OVERLAP_HTSLIB_1_13]


def _rand_cigar(rng, l_ref_target, indels):
    """a CIGAR with soft clips, matches and (optionally) insertions / deletions / skips; returns ops, l_qseq, l_ref"""
    ops = []
    if rng.random() < 0.3:
        ops.append((4, rng.randint(1, 6)))
    left = l_ref_target
    while left > 0:
        m = min(left, rng.randint(3, 40))
        ops.append((rng.choice([0, 0, 0, 7, 8]), m))
        left -= m
        if left > 0 and indels and rng.random() < 0.5:
            kind = rng.choice([1, 2, 2, 3])
            n = rng.randint(1, 4)
            ops.append((kind, n))
            if kind in (2, 3):
                left -= min(left, n)
    if ops[-1][0] in (1, 2, 3):
        ops.append((0, 2))
    if rng.random() < 0.3:
        ops.append((4, rng.randint(1, 6)))
    # merge is not needed: adjacent match ops of different kinds are legal
    lq = sum(n for o, n in ops if o in (0, 1, 4, 7, 8))
    lr = sum(n for o, n in ops if o in (0, 2, 3, 7, 8))
    return ops, lq, lr


def make_pairs(seed, n_pairs=60, indels=True, same_start=False, with_extras=True):
    rng = random.Random(seed)
    reads = []
    for k in range(n_pairs):
        p1 = 50 if same_start else rng.randint(0, 400)
        c1, lq1, lr1 = _rand_cigar(rng, rng.randint(30, 90), indels)
        p2 = p1 + rng.randint(0, lr1 + 20)                    # mostly overlapping, sometimes just past the end
        c2, lq2, lr2 = _rand_cigar(rng, rng.randint(30, 90), indels)
        name = f"pair{seed}_{k}"
        s1 = "".join(rng.choice("ACGT") for _ in range(lq1))
        s2 = "".join(rng.choice("ACGT") for _ in range(lq2))
        # make the overlapping bases mostly agree: copy the reference-aligned part
        ref = {}
        q, r = 0, p1
        for o, n in c1:
            if o in (0, 7, 8):
                for j in range(n):
                    ref[r + j] = s1[q + j]
                q += n; r += n
            elif o in (1, 4):
                q += n
            elif o in (2, 3):
                r += n
        s2l = list(s2)
        q, r = 0, p2
        for o, n in c2:
            if o in (0, 7, 8):
                for j in range(n):
                    if r + j in ref and rng.random() < 0.85:
                        s2l[q + j] = ref[r + j]
                q += n; r += n
            elif o in (1, 4):
                q += n
            elif o in (2, 3):
                r += n
        s2 = "".join(s2l)
        q1 = [rng.choice([2, 12, 23, 37, 37, 37, 40, 93, 120]) for _ in range(lq1)]
        q2 = [rng.choice([2, 12, 23, 37, 37, 37, 40, 93, 120]) for _ in range(lq2)]
        end2 = p2 + lr2
        tlen = end2 - p1
        f1, f2 = 99, 147
        if rng.random() < 0.1:
            f1, f2 = 97, 145                                  # not a proper pair: orphans are dropped by the stepper
        reads.append(po.Read(f1, p1, 60, c1, s1, q1, name, p2, 1, tlen))
        reads.append(po.Read(f2, p2, 60, c2, s2, q2, name, p1, 1, -tlen))
        if with_extras and rng.random() < 0.15:               # a supplementary alignment sharing the name
            cs, lqs, lrs = _rand_cigar(rng, 20, False)
            ss = "".join(rng.choice("ACGT") for _ in range(lqs))
            reads.append(po.Read(2048 | 99, p1 + rng.randint(0, 30), 60, cs, ss, [30] * lqs, name, p2, 1, tlen))
        if with_extras and rng.random() < 0.1:                # mate on another contig / no mate information
            cs, lqs, lrs = _rand_cigar(rng, 25, False)
            ss = "".join(rng.choice("ACGT") for _ in range(lqs))
            reads.append(po.Read(99, p1 + 3, 60, cs, ss, [35] * lqs, f"lone{seed}_{k}", rng.choice([-1, 10]), rng.choice([0, -1]), 0))
    return po.samtools_sort(reads)


def product_rewrite(reads, model, min_mapq=0, max_depth=8000):
    from lvc_b200 import packing
    rows = [(r.flag, r.pos, r.mapq, r.cigar, r.seq, r.qual, r.name, r.mpos, r.mref, r.tlen) for r in reads]
    b = packing.pack_reads(rows, min_mapq, max_depth, overlap_model=model)
    quals = [b.qual[int(b.seq_off[i]):int(b.seq_off[i]) + len(r.qual)].tolist() for i, r in enumerate(reads)]
    return b, quals


def oracle_rewrite(reads, model, min_mapq=0, max_depth=8000):
    tweaked, admitted = {}, []
    for _ in po.pileup_columns(reads, min_mapq, max_depth, admitted=admitted, overlap_model=model, tweaked=tweaked):
        pass
    return [tweaked.get(i, list(r.qual)) for i, r in enumerate(reads)], set(admitted)


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("seed", range(12))
def test_rewrite_matches_htslib_emulation(lib, model, seed):
    reads = make_pairs(seed, indels=seed % 3 != 0)
    b, got = product_rewrite(reads, model)
    want, admitted = oracle_rewrite(reads, model)
    for i, r in enumerate(reads):
        assert got[i] == want[i], (i, r.name, r.flag, r.pos, r.cigar)
    assert {i for i in range(len(reads)) if b.keep[i] & 1} == admitted
    assert b.overlap_pairs > 0
    # the input reads themselves are untouched (htslib rewrites its buffered copies)
    assert any(got[i] != list(r.qual) for i, r in enumerate(reads))


@pytest.mark.parametrize("model", MODELS)
def test_depth_cap_drops_unpair(lib, model):
    """reads dropped by max_depth never enter the hash, and a dropped mate removes its partner's entry"""
    reads = make_pairs(99, n_pairs=40, indels=False, same_start=True, with_extras=False)
    b, got = product_rewrite(reads, model, max_depth=25)
    want, admitted = oracle_rewrite(reads, model, max_depth=25)
    assert {i for i in range(len(reads)) if b.keep[i] & 1} == admitted
    assert len(admitted) < len(reads)
    assert got == want


def test_known_answer_two_reads(lib):
    """a hand-checked pair: 10 bp overlap, one mismatch, equal qualities on it"""
    a = po.Read(99, 100, 60, [(0, 20)], "ACGTACGTACGTACGTACGT", [30] * 20, "q1", 110, 1, 30)
    b = po.Read(147, 110, 60, [(0, 20)], "GTACGTACGAACGTACGTAC", [30] * 10 + [20] * 10, "q1", 100, 1, -30)
    # legacy: first mate gets the sum, second 0; the mismatch (a[19]='T' vs b[9]='A', 30 vs 30): a keeps int(0.8*30)=24
    _, got = product_rewrite([a, b], po.OVERLAP_HTSLIB_1_10)
    assert got[0] == [30] * 10 + [60] * 9 + [24]
    assert got[1] == [0] * 10 + [20] * 10
    # 1.13: the name hash picks the mate that keeps the qualities
    keep_a = po._wang_hash(po._x31_hash_string("q1")) & 1
    _, got = product_rewrite([a, b], po.OVERLAP_HTSLIB_1_13)
    if keep_a:
        assert got[0] == [30] * 10 + [60] * 9 + [24] and got[1] == [0] * 10 + [20] * 10
    else:
        assert got[0] == [30] * 10 + [0] * 10 and got[1] == [60] * 9 + [24] + [20] * 10
    # off: untouched
    _, got = product_rewrite([a, b], po.OVERLAP_OFF)
    assert got == [a.qual, b.qual]


@pytest.mark.parametrize("model", MODELS)
def test_native_ingest_and_python_reader_agree_on_overlaps(lib, tmp_path, model):
    """the same BAM through the native ingest and through the pure-Python reader: identical packed qualities"""
    from lvc_b200 import samio, capi
    reads = make_pairs(7, n_pairs=80)
    for r in reads:
        r.qual = [min(q, 93) for q in r.qual]                 # SAM text cannot carry phred > 93
    recs = [(r.flag, r.pos, r.mapq, r.cigar, r.seq, r.qual, r.name, r.mpos, r.mref, r.tlen) for r in reads]
    bam = str(tmp_path / "pairs.bam")
    samio.write_bam(bam, [("chrT", 1000), ("chrU", 500)], recs)
    _, py = samio.read_bam(bam, None, 0, overlap_model=model)
    nat = capi.NativeReads(bam, None, 0, overlap_model=model)
    nb = nat.as_readbatch()
    assert nat.overlap_pairs == py.overlap_pairs > 0 and nat.overlap_bases == py.overlap_bases
    assert np.array_equal(nb.qual[:py.n_qual], py.qual[:py.n_qual])
    assert np.array_equal(nb.keep, py.keep)
    want, _ = oracle_rewrite(reads, model)
    for i, r in enumerate(reads):
        assert py.qual[int(py.seq_off[i]):int(py.seq_off[i]) + len(r.qual)].tolist() == want[i]
    nat.close()
    # SAM text, unsorted input: same result after the samtools-order sort
    sam = str(tmp_path / "pairs.sam")
    order = list(range(len(reads)))
    random.Random(3).shuffle(order)
    with open(sam, "w") as fh:
        fh.write("@HD\tVN:1.6\n@SQ\tSN:chrT\tLN:1000\n@SQ\tSN:chrU\tLN:500\n")
        for i in order:
            r = reads[i]
            cig = "".join(f"{n}{po.CIGAR_OPS[o]}" for o, n in r.cigar)
            rnext = "=" if r.mref == 1 else ("chrU" if r.mref == 0 else "*")
            fh.write("\t".join([r.name, str(r.flag), "chrT", str(r.pos + 1), str(r.mapq), cig, rnext, str(r.mpos + 1),
                                str(r.tlen), r.seq, "".join(chr(q + 33) for q in r.qual)]) + "\n")
    nat2 = capi.NativeReads(sam, None, 0, overlap_model=model)
    assert nat2.overlap_pairs > 0
    # (ties in the sort may order same-position reads differently from `reads`; compare as multisets per read name)
    nb2 = nat2.as_readbatch()
    _, py2 = samio.read_sam(sam, None, 0, overlap_model=model)
    assert np.array_equal(nb2.qual[:py2.n_qual], py2.qual[:py2.n_qual]) and np.array_equal(nb2.pos, py2.pos)
    nat2.close()


def test_native_ingest_many_overlapping_pairs(lib, tmp_path):
    """thousands of overlapping pairs: the native ingest lists the rewrites while the admission pass runs beside the payload
    copy and applies them on several threads afterwards; the result is the pure-Python reader's, byte for byte"""
    from lvc_b200 import samio, capi
    reads = make_pairs(11, n_pairs=4200, with_extras=True)
    for r in reads:
        r.qual = [min(q, 93) for q in r.qual]
    recs = [(r.flag, r.pos, r.mapq, r.cigar, r.seq, r.qual, r.name, r.mpos, r.mref, r.tlen) for r in reads]
    bam = str(tmp_path / "many.bam")
    samio.write_bam(bam, [("chrT", 1000), ("chrU", 500)], recs)
    _, py = samio.read_bam(bam, None, 0)
    for nt in (1, 2, 7):
        nat = capi.NativeReads(bam, None, 0, n_threads=nt)
        nb = nat.as_readbatch()
        assert nat.overlap_pairs == py.overlap_pairs > 2048 and nat.overlap_bases == py.overlap_bases, nt
        assert np.array_equal(nb.qual[:py.n_qual], py.qual[:py.n_qual]), nt
        assert np.array_equal(nb.keep, py.keep), nt
        nat.close()


@pytest.mark.parametrize("model", MODELS)
def test_rewrite_single_match_reads(lib, model):
    """mates whose CIGAR is [clips] one match op [clips] take a plain loop over the shared positions instead of the cursor
    walk: same result as the literal htslib emulation, for every geometry (mate inside the first read, starting at its
    last base, just past its end, equal starts) and both release models"""
    rng = random.Random(2026)
    reads = []
    for k in range(400):
        p1 = rng.randint(0, 300)
        l1, l2 = rng.randint(1, 80), rng.randint(1, 80)
        p2 = p1 + rng.choice([0, 0, l1 - 1, l1, l1 + 3, rng.randint(0, l1)])

        def cig(l):
            ops = []
            if rng.random() < 0.2: ops.append((5, rng.randint(1, 5)))
            if rng.random() < 0.4: ops.append((4, rng.randint(1, 9)))
            ops.append((rng.choice([0, 0, 7, 8]), l))
            if rng.random() < 0.4: ops.append((4, rng.randint(1, 9)))
            if rng.random() < 0.2: ops.append((5, rng.randint(1, 5)))
            return ops, sum(n for o, n in ops if o in (0, 4, 7, 8))
        c1, lq1 = cig(l1)
        c2, lq2 = cig(l2)
        s1 = "".join(rng.choice("ACGT") for _ in range(lq1))
        s2 = "".join(rng.choice("AACGT") for _ in range(lq2))
        q1 = [rng.choice([2, 12, 23, 37, 37, 40, 100, 120]) for _ in range(lq1)]
        q2 = [rng.choice([2, 12, 23, 37, 37, 40, 100, 120]) for _ in range(lq2)]
        tl = p2 + l2 - p1
        reads.append(po.Read(99, p1, 60, c1, s1, q1, f"s{k}", p2, 1, tl))
        reads.append(po.Read(147, p2, 60, c2, s2, q2, f"s{k}", p1, 1, -tl))
    reads = po.samtools_sort(reads)
    b, got = product_rewrite(reads, model)
    want, admitted = oracle_rewrite(reads, model)
    changed = 0
    for i, r in enumerate(reads):
        assert got[i] == want[i], (i, r.name, r.cigar)
        changed += got[i] != list(r.qual)
    assert changed > 300
