"""Seeded SMALL synthetic scenarios for the golden vectors (test infrastructure, not product code).

Used by oracle/gen_golden.py (to run the real reference on them) and stored verbatim inside
tests/golden/synthetic_golden.json so tests never depend on RNG reproducibility.
"""
from __future__ import annotations

import random
from typing import Dict, List, Tuple

from .pileup_oracle import Read, read_sam, samtools_sort, CIGAR_OPS

THRESHOLDS = {
    "strict": dict(minBQ=30, minMQ=20, minDP=10, minAD=5, ratio=0.10),
    "loose": dict(minBQ=13, minMQ=0, minDP=3, minAD=2, ratio=0.05),
    "zero": dict(minBQ=0, minMQ=0, minDP=1, minAD=1, ratio=0.0),
}

SCENARIOS: Dict[str, dict] = {
    "mixed_small": dict(seed=11, ref_len=400, n_reads=160, len_lo=30, len_hi=120, q_lo=0, q_hi=60,
                        indel_rate=0.06, weird=True),
    "ont_like": dict(seed=12, ref_len=900, n_reads=60, len_lo=300, len_hi=600, q_lo=2, q_hi=90,
                     indel_rate=0.05, weird=False, clip=(40, 70)),
    "deep_underflow": dict(seed=13, ref_len=64, n_reads=420, len_lo=64, len_hi=64, q_lo=28, q_hi=41,
                           indel_rate=0.0, weird=False, fixed_pos=0, snv_rate=0.08,
                           planted=((5, 0.5), (17, 0.15), (40, 1.0), (41, 0.3))),
    "amplicon_like": dict(seed=14, ref_len=700, n_reads=700, len_lo=100, len_hi=100, q_lo=2, q_hi=40,
                          indel_rate=0.01, weird=False, amplicon=(0, 300, 550), qbins=(2, 12, 23, 37),
                          planted=((10, 0.5), (50, 0.12), (320, 1.0), (321, 0.25), (600, 0.05))),
    "maxdepth": dict(seed=15, ref_len=200, n_reads=9300, len_lo=12, len_hi=14, q_lo=20, q_hi=41,
                     indel_rate=0.0, weird=False, amplicon=(5, 5, 5, 5, 5, 5, 5, 5, 9, 30), store_memory=False),
}


def _rand_seq(rng, n, weird):
    alphabet = "ACGT"
    out = []
    for _ in range(n):
        x = rng.random()
        if weird and x < 0.012:
            out.append("N")
        elif weird and x < 0.016:
            out.append(rng.choice("RYMKSW="))
        else:
            out.append(rng.choice(alphabet))
    return "".join(out)


def make_scenario(seed, ref_len, n_reads, len_lo, len_hi, q_lo, q_hi, indel_rate, weird,
                  clip=None, fixed_pos=None, snv_rate=0.01, amplicon=None, qbins=None,
                  store_memory=True, planted=()) -> Tuple[str, List[Read]]:
    rng = random.Random(seed)
    planted = dict(planted)
    ref = "".join(rng.choice("ACGT") for _ in range(ref_len))
    if weird:   # lower-case and N stretches in the FASTA are legal and reach the records
        ref = ref[:50] + ref[50:70].lower() + ref[70:90] + "NNNN" + ref[94:]
    reads = []
    for i in range(n_reads):
        span = rng.randint(len_lo, len_hi)
        if fixed_pos is not None:
            pos = fixed_pos
        elif amplicon is not None:
            pos = rng.choice(amplicon)
        else:
            pos = rng.randint(0, max(0, ref_len - span))
        span = min(span, ref_len - pos)
        # build CIGAR over `span` reference bases
        cigar: List[Tuple[int, int]] = []
        seq: List[str] = []
        if weird and rng.random() < 0.1:
            cigar.append((5, rng.randint(1, 9)))                       # H
        if clip is not None or (weird and rng.random() < 0.3):
            n = rng.randint(*clip) if clip else rng.randint(1, 12)
            cigar.append((4, n))
            seq.append(_rand_seq(rng, n, False))
        r = pos
        remaining = span
        first = True
        while remaining > 0:
            if first and weird and rng.random() < 0.03:
                n = min(remaining, rng.randint(1, 3))
                cigar.append((2, n))                                   # leading deletion (legal, odd)
                r += n
                remaining -= n
                first = False
                continue
            first = False
            run = remaining if rng.random() > indel_rate * 8 else rng.randint(1, max(1, min(remaining, 40)))
            op = 0
            if weird:
                x = rng.random()
                op = 7 if x < 0.1 else (8 if x < 0.15 else 0)
            cigar.append((op, run))
            for j in range(run):
                b = ref[r + j].upper()
                if b not in "ACGT":
                    b = rng.choice("ACGT")
                if op == 8 or (op == 0 and rng.random() < snv_rate):
                    b = rng.choice([c for c in "ACGT" if c != b])
                if (r + j) in planted and rng.random() < planted[r + j]:
                    b = "ACGT"[("ACGT".index(b) + 1 + (r + j) % 3) % 4]
                if weird and rng.random() < 0.012:
                    b = "N" if rng.random() < 0.7 else rng.choice("RYMKSW=")
                seq.append(b)
            r += run
            remaining -= run
            if remaining > 0:
                x = rng.random()
                if x < 0.4:
                    n = rng.randint(1, 4)
                    cigar.append((1, n))                               # I
                    seq.append(_rand_seq(rng, n, weird))
                    if weird and rng.random() < 0.2 and remaining > 3:  # I followed directly by D
                        n = rng.randint(1, 3)
                        cigar.append((2, n)); r += n; remaining -= n
                elif x < 0.8:
                    n = min(remaining, rng.randint(1, 5))
                    cigar.append((2, n)); r += n; remaining -= n       # D
                elif weird and x < 0.9:
                    n = min(remaining, rng.randint(2, 20))
                    cigar.append((3, n)); r += n; remaining -= n       # N (ref skip)
                elif weird:
                    cigar.append((6, rng.randint(1, 3)))               # P
        if clip is not None or (weird and rng.random() < 0.3):
            n = rng.randint(*clip) if clip else rng.randint(1, 12)
            cigar.append((4, n))
            seq.append(_rand_seq(rng, n, False))
        if weird and rng.random() < 0.1:
            cigar.append((5, rng.randint(1, 9)))
        s = "".join(seq)
        if qbins:
            qual = [rng.choices(qbins, weights=(1, 3, 6, 90))[0] for _ in s]
        else:
            qual = [rng.randint(q_lo, q_hi) for _ in s]
        flag = 16 if rng.random() < 0.5 else 0
        mapq = 60 if rng.random() < 0.9 else rng.randint(0, 25)
        if weird:
            x = rng.random()
            if x < 0.04:
                flag |= rng.choice([0x4, 0x100, 0x200, 0x400])
            elif x < 0.10:
                flag |= 0x1 | 0x2 | rng.choice([0x40, 0x80])          # proper pair (mate far away)
            elif x < 0.14:
                flag |= 0x1 | rng.choice([0x40, 0x80])                # orphan: dropped
            elif x < 0.17:
                flag |= 0x800                                          # supplementary: kept
        reads.append(Read(flag, pos, mapq, cigar, s, qual, f"r{i}"))
    return ref, samtools_sort(reads)


def reads_to_rows(reads: List[Read]):
    return [[r.name, r.flag, r.pos, r.mapq, "".join(f"{l}{CIGAR_OPS[o]}" for o, l in r.cigar), r.seq,
             "".join(chr(q + 33) for q in r.qual)] for r in reads]


def rows_to_reads(rows) -> List[Read]:
    from .pileup_oracle import parse_cigar
    return [Read(f, p, m, parse_cigar(c), s, [ord(x) - 33 for x in q], n) for n, f, p, m, c, s, q in rows]


def write_testfile_fasta(sam_path: str, fasta_path: str):
    """contig NC_045512.2, 29,903 bp, covered positions = majority base of the
    reads of test/testdata/testfile.sam at minBQ 0, 'N' elsewhere."""
    from .pileup_oracle import pileup_columns
    contigs, reads = read_sam(sam_path)
    name, length = contigs[0]
    ref = ["N"] * length
    for pos, entries in pileup_columns(reads, 0):
        cnt: Dict[str, int] = {}
        for idx, r, is_del, is_skip, qpos in entries:
            if not is_del:
                cnt[r.seq[qpos]] = cnt.get(r.seq[qpos], 0) + 1
        if cnt:
            best = max(cnt.values())
            ref[pos] = sorted(b for b, c in cnt.items() if c == best)[0]
    with open(fasta_path, "w") as fh:
        fh.write(f">{name}\n")
        s = "".join(ref)
        for i in range(0, length, 70):
            fh.write(s[i:i + 70] + "\n")
