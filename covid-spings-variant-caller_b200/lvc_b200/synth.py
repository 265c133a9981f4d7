"""Deterministic synthetic workloads of BASELINE.json's configs, written straight into the
packed structure-of-arrays batch (no BAM round trip).  numpy only; seeds as stated in SURVEY 8d.

config 2: Illumina 2x150 amplicon reads at 10,000x, single sample (SARS-CoV-2 geometry, G = 29,903)
config 4: 96 samples of config-2 geometry at 5,000x
config 5: 5 Mb genome at 1,000x, shotgun 150 bp reads
config 3: ONT-like reads in live batches of 1,000x
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

from . import packing
from .packing import ReadBatch

SARS2_LEN = 29903
QBINS = np.array([2, 12, 23, 37], dtype=np.uint8)           # RTA3 binned qualities
QBIN_P = np.array([0.01, 0.03, 0.06, 0.90])
_NIB = np.array([1, 2, 4, 8], dtype=np.uint8)               # A C G T nibble codes


def random_reference(length: int, seed: int) -> str:
    rng = np.random.default_rng(seed)
    return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, length)].tobytes().decode("ascii")


def _reference_codes(ref: str) -> np.ndarray:
    lut = np.full(256, 0, dtype=np.uint8)
    for i, c in enumerate("ACGT"):
        lut[ord(c)] = i
        lut[ord(c.lower())] = i
    return lut[np.frombuffer(ref.encode("latin-1"), dtype=np.uint8)]


def _illumina_block(rng, refc: np.ndarray, pos: np.ndarray, read_len: int, snv_pos: np.ndarray, snv_alt: np.ndarray,
                    snv_af: np.ndarray, indel_frac: float):
    """reads starting at pos[i] (int64), all `read_len` query bases.  Returns (cigar list per read as arrays,
    base codes (n, L) 0..3, qual (n, L))."""
    n = len(pos)
    L = read_len
    G = len(refc)
    qual = QBINS[rng.choice(4, size=(n, L), p=QBIN_P)]
    # CIGAR model: 99% LM, 1% one indel of 1-3 bp (xM yI zM / xM yD zM)
    kind = np.zeros(n, dtype=np.int8)                       # 0 none, 1 insertion, 2 deletion
    has = rng.random(n) < indel_frac
    kind[has] = rng.integers(1, 3, has.sum())
    y = rng.integers(1, 4, n)
    x = rng.integers(10, L - 20, n)
    j = np.arange(L)[None, :]
    shift = np.zeros((n, L), dtype=np.int64)
    ins_mask = np.zeros((n, L), dtype=bool)
    isdel = kind == 2
    isins = kind == 1
    shift[isdel] = np.where(j >= x[isdel, None], y[isdel, None], 0)
    shift[isins] = np.where(j >= (x[isins] + y[isins])[:, None], -y[isins, None], 0)
    ins_mask[isins] = (j >= x[isins, None]) & (j < (x[isins] + y[isins])[:, None])
    ridx = pos[:, None] + j + shift
    # deletion reads may run past the reference end: turn them into plain reads
    over = ridx.max(axis=1) >= G
    if over.any():
        kind[over] = 0
        shift[over] = 0
        ins_mask[over] = False
        ridx = pos[:, None] + j + shift
    base = refc[np.minimum(ridx, G - 1)]
    # planted SNVs: a read carries the ALT with probability AF at each planted position it covers
    if len(snv_pos):
        lo, hi = int(pos.min()), int(pos.max()) + L + 4
        sorted_pos = bool((np.diff(pos) >= 0).all())
        for k in np.nonzero((snv_pos >= lo) & (snv_pos < hi))[0]:
            # only reads starting in [snv - L - 3, snv] can cover it (pos is sorted: two binary searches)
            r0, r1 = (int(np.searchsorted(pos, snv_pos[k] - L - 4, "left")), int(np.searchsorted(pos, snv_pos[k], "right"))) \
                if sorted_pos else (0, n)
            hit = (ridx[r0:r1] == snv_pos[k]) & ~ins_mask[r0:r1]
            rows = np.nonzero(hit.any(axis=1))[0]
            if len(rows) == 0:
                continue
            carry = rng.random(len(rows)) < snv_af[k]
            rr = rows[carry]
            cc = hit[rr].argmax(axis=1)
            base[r0 + rr, cc] = snv_alt[k]
    base[ins_mask] = rng.integers(0, 4, int(ins_mask.sum()))
    # sequencing errors drawn from the base quality
    err = rng.random((n, L)) < np.power(10.0, qual.astype(np.float64) / -10.0)
    base = np.where(err, (base + rng.integers(1, 4, (n, L))) & 3, base).astype(np.uint8)
    return kind, x, y, base, qual


def _flags_mapq(rng, n: int, base_flag: np.ndarray):
    mapq = np.where(rng.random(n) < 0.98, 60, rng.integers(0, 20, n)).astype(np.uint8)
    flag = base_flag.astype(np.uint16).copy()
    bad = rng.random(n) < 0.005
    flag[bad] |= rng.choice(np.array([0x400, 0x100, 0x200], dtype=np.uint16), int(bad.sum()))
    return flag, mapq


def _assemble(pos, flag, mapq, kind, x, y, base, qual, min_mapq: int, max_depth: int) -> ReadBatch:
    n, L = base.shape
    ncig = np.where(kind == 0, 1, 3)
    coff = np.concatenate([[0], np.cumsum(ncig)]).astype(np.uint32)
    cig = np.zeros(int(coff[-1]), dtype=np.uint32)
    c0 = coff[:-1].astype(np.int64)
    plain = kind == 0
    cig[c0[plain]] = (L << 4) | 0
    ins = kind == 1
    cig[c0[ins]] = (x[ins] << 4) | 0
    cig[c0[ins] + 1] = (y[ins] << 4) | 1
    cig[c0[ins] + 2] = ((L - x[ins] - y[ins]) << 4) | 0
    dele = kind == 2
    cig[c0[dele]] = (x[dele] << 4) | 0
    cig[c0[dele] + 1] = (y[dele] << 4) | 2
    cig[c0[dele] + 2] = ((L - x[dele]) << 4) | 0
    nib = _NIB[base]
    Lp = L + (L & 1)
    if Lp != L:
        nib = np.concatenate([nib, np.zeros((n, 1), np.uint8)], axis=1)
        qual = np.concatenate([qual, np.zeros((n, 1), np.uint8)], axis=1)
    seq4 = ((nib[:, 0::2] << 4) | nib[:, 1::2]).astype(np.uint8).ravel()
    soff = (np.arange(n + 1, dtype=np.uint64) * np.uint64(Lp))
    return packing.finalize_batch(pos.astype(np.int32), flag, mapq, coff, cig, soff, seq4, qual.ravel(), min_mapq,
                                  max_depth, acgt_only=np.ones(n, dtype=bool))


def amplicon_sample(seed: int = 20260101, n_pairs: int = 996_767, ref: Optional[str] = None, ref_len: int = SARS2_LEN,
                    read_len: int = 150, amplicon_len: int = 400, amplicon_step: int = 300, n_snvs: int = 50,
                    min_mapq: int = 20, max_depth: int = 8000, ref_seed: int = 20260100) -> Tuple[str, ReadBatch]:
    """SURVEY 8d config 2 (and, with n_pairs halved and seed 20260300+k, one sample of config 4)."""
    if ref is None:
        ref = random_reference(ref_len, ref_seed)
    G = len(ref)
    refc = _reference_codes(ref)
    rng = np.random.default_rng(seed)
    n_amp = (G - amplicon_len) // amplicon_step + 1
    per = np.full(n_amp, n_pairs // n_amp)
    per[: n_pairs % n_amp] += 1
    snv_pos = np.sort(rng.choice(np.arange(50, G - 50), n_snvs, replace=False))
    snv_alt = ((refc[snv_pos] + rng.integers(1, 4, n_snvs)) & 3).astype(np.uint8)
    snv_af = rng.choice(np.array([0.02, 0.1, 0.5, 1.0]), n_snvs)
    parts = []
    for a in range(n_amp):
        start = a * amplicon_step
        for mate in (0, 1):
            n = int(per[a])
            p0 = start if mate == 0 else start + amplicon_len - read_len
            pos = np.full(n, p0, dtype=np.int64)
            kind, x, y, base, qual = _illumina_block(rng, refc, pos, read_len, snv_pos, snv_alt, snv_af, 0.01)
            flag, mapq = _flags_mapq(rng, n, np.full(n, 99 if mate == 0 else 147))
            parts.append((pos, flag, mapq, kind, x, y, base, qual))
    cat = [np.concatenate([p[i] for p in parts]) for i in range(8)]
    return ref, _assemble(*cat, min_mapq=min_mapq, max_depth=max_depth)


def amplicon_pairing(batch: ReadBatch, n_pairs: int = 996_767, ref_len: int = SARS2_LEN, read_len: int = 150,
                     amplicon_len: int = 400, amplicon_step: int = 300):
    """(name_id, mate_pos, tlen) of the reads of amplicon_sample(n_pairs=...): the k-th forward read of an amplicon
    and its k-th reverse read are mates (insert = amplicon length, so 2x150 mates of a 400 bp amplicon do not overlap)."""
    n_amp = (ref_len - amplicon_len) // amplicon_step + 1
    per = np.full(n_amp, n_pairs // n_amp)
    per[: n_pairs % n_amp] += 1
    pair_base = np.concatenate([[0], np.cumsum(per)])[:-1]
    ids, mpos, tl = [], [], []
    for a in range(n_amp):
        k = np.arange(per[a], dtype=np.int64) + pair_base[a]
        start = a * amplicon_step
        ids += [k, k]
        mpos += [np.full(per[a], start + amplicon_len - read_len), np.full(per[a], start)]
        tl += [np.full(per[a], amplicon_len), np.full(per[a], -amplicon_len)]
    ids, mpos, tl = np.concatenate(ids), np.concatenate(mpos), np.concatenate(tl)
    assert len(ids) == batch.n_reads
    return ids.astype(np.uint32), mpos.astype(np.int32), tl.astype(np.int32)


def shotgun_sample(seed: int = 20260400, ref_len: int = 5_000_000, depth: float = 1000.0, read_len: int = 150,
                   ref: Optional[str] = None, min_mapq: int = 20, max_depth: int = 8000, n_snvs: int = 500,
                   block: int = 1 << 18) -> Tuple[str, ReadBatch]:
    """single-end 150 bp reads with uniform random starts, coordinate sorted."""
    if ref is None:
        ref = random_reference(ref_len, seed + 1)
    G = len(ref)
    refc = _reference_codes(ref)
    rng = np.random.default_rng(seed)
    n = int(round(G * depth / read_len))
    pos_all = np.sort(rng.integers(0, G - read_len - 4, n)).astype(np.int64)
    strand = rng.random(n) < 0.5
    # samtools sort: forward before reverse on ties
    order = np.lexsort((strand, pos_all))
    pos_all, strand = pos_all[order], strand[order]
    snv_pos = np.sort(rng.choice(np.arange(50, G - 50), n_snvs, replace=False))
    snv_alt = ((refc[snv_pos] + rng.integers(1, 4, n_snvs)) & 3).astype(np.uint8)
    snv_af = rng.choice(np.array([0.02, 0.1, 0.5, 1.0]), n_snvs)
    parts = []
    for s in range(0, n, block):
        pos = pos_all[s:s + block]
        kind, x, y, base, qual = _illumina_block(rng, refc, pos, read_len, snv_pos, snv_alt, snv_af, 0.01)
        flag, mapq = _flags_mapq(rng, len(pos), np.where(strand[s:s + block], 16, 0))
        parts.append((pos, flag, mapq, kind, x, y, base, qual))
    cat = [np.concatenate([p[i] for p in parts]) for i in range(8)]
    return ref, _assemble(*cat, min_mapq=min_mapq, max_depth=max_depth)


def ont_batch(seed: int, ref: str, depth: float = 1000.0, ref_span: int = 400, min_mapq: int = 20,
              max_depth: int = 8000) -> ReadBatch:
    """one live batch of ONT-like reads modelled on test/testdata/testfile.sam:
    ref span ~400, 50-70 bp soft clips both ends, ~21 CIGAR ops (~5% indel ops), qualities 2..90 with
    mean ~20 (discretised log-normal), flag 0/16, mapq 60."""
    G = len(ref)
    refc = _reference_codes(ref)
    rng = np.random.default_rng(seed)
    n = int(round(G * depth / ref_span))
    pos = np.sort(rng.integers(0, G - ref_span - 40, n)).astype(np.int64)
    strand = rng.random(n) < 0.5
    order = np.lexsort((strand, pos))
    pos, strand = pos[order], strand[order]
    rows = []
    letters = "ACGT"
    for i in range(n):
        ops: List[Tuple[int, int]] = [(4, int(rng.integers(50, 71)))]
        seq = list(rng.integers(0, 4, ops[0][1]))
        r = int(pos[i])
        remaining = ref_span
        while remaining > 0:
            run = int(min(remaining, max(1, rng.geometric(1 / 38.0))))
            ops.append((0, run))
            seg = refc[r:r + run].copy()
            errs = rng.random(run) < 0.03
            seg[errs] = (seg[errs] + rng.integers(1, 4, int(errs.sum()))) & 3
            seq.extend(seg.tolist())
            r += run
            remaining -= run
            if remaining > 0:
                if rng.random() < 0.5:
                    k = int(rng.integers(1, 4))
                    ops.append((1, k))
                    seq.extend(rng.integers(0, 4, k).tolist())
                else:
                    k = int(min(remaining, rng.integers(1, 4)))
                    ops.append((2, k))
                    r += k
                    remaining -= k
        if ops[-1][0] == 2:                     # never end on a deletion
            ops.pop()
        k = int(rng.integers(50, 71))
        ops.append((4, k))
        seq.extend(rng.integers(0, 4, k).tolist())
        q = np.clip(np.round(np.exp(rng.normal(2.85, 0.55, len(seq)))), 2, 90).astype(np.uint8)
        rows.append((16 if strand[i] else 0, int(pos[i]), 60, ops, "".join(letters[b] for b in seq), q))
    return packing.pack_reads(rows, min_mapq, max_depth)


def ont_batch_fast(seed: int, ref: str, depth: float = 1000.0, ref_span: int = 400, min_mapq: int = 20,
                   max_depth: int = 8000, n_runs: int = 10) -> ReadBatch:
    """Vectorised SURVEY 8d config-3 batch (same model as ont_batch, fixed op count): every read is
    S, M, (I|D, M) x 9, S = 21 CIGAR ops, reference span `ref_span`, 50-70 bp soft clips, indels of 1-3 bp,
    3 % substitutions, qualities round(exp(N(2.85, 0.55))) clipped to 2..90, flag 0/16, mapq 60."""
    G = len(ref)
    refc = _reference_codes(ref)
    rng = np.random.default_rng(seed)
    n = int(round(G * depth / ref_span))
    pos = np.sort(rng.integers(0, G - ref_span - 40, n)).astype(np.int64)
    strand = rng.random(n) < 0.5
    order = np.lexsort((strand, pos))
    pos, strand = pos[order], strand[order]
    n_gap = n_runs - 1
    is_del = rng.random((n, n_gap)) < 0.5
    glen = rng.integers(1, 4, (n, n_gap)).astype(np.int64)
    m_total = ref_span - np.where(is_del, glen, 0).sum(axis=1)                 # reference bases in match runs
    # n_runs positive run lengths summing to m_total: sorted distinct cut points
    cuts = np.sort(rng.random((n, n_gap)), axis=1)
    edges = np.concatenate([np.zeros((n, 1)), cuts, np.ones((n, 1))], axis=1)
    runs = np.floor(np.diff(edges, axis=1) * (m_total - n_runs)[:, None]).astype(np.int64) + 1
    runs[:, -1] += m_total - runs.sum(axis=1)
    clipl, clipr = rng.integers(50, 71, n).astype(np.int64), rng.integers(50, 71, n).astype(np.int64)
    n_ops = 2 * n_runs + 1
    op = np.zeros((n, n_ops), dtype=np.int64)
    ln = np.zeros((n, n_ops), dtype=np.int64)
    op[:, 0], ln[:, 0], op[:, -1], ln[:, -1] = 4, clipl, 4, clipr
    op[:, 1:-1:2], ln[:, 1:-1:2] = 0, runs
    op[:, 2:-1:2], ln[:, 2:-1:2] = np.where(is_del, 2, 1), glen
    cigar = ((ln << 4) | op).astype(np.uint32).ravel()
    coff = (np.arange(n + 1, dtype=np.int64) * n_ops).astype(np.uint32)
    # per-op reference start and query length
    consumes_ref = (op == 0) | (op == 2)
    consumes_q = op != 2
    ref_start = pos[:, None] + np.cumsum(np.where(consumes_ref, ln, 0), axis=1) - np.where(consumes_ref, ln, 0)
    qlen_op = np.where(consumes_q, ln, 0)
    lq = qlen_op.sum(axis=1)
    lqp = lq + (lq & 1)                                                         # reads start at even offsets
    soff = np.concatenate([[0], np.cumsum(lqp)]).astype(np.uint64)
    # expand to bases
    tot = int(lq.sum())
    flat_q = qlen_op.ravel()
    op_of_base = np.repeat(np.arange(n * n_ops, dtype=np.int64), flat_q)
    starts = np.cumsum(flat_q) - flat_q
    within = np.arange(tot, dtype=np.int64) - starts[op_of_base]
    is_m = op.ravel()[op_of_base] == 0
    base = rng.integers(0, 4, tot).astype(np.uint8)
    rp = ref_start.ravel()[op_of_base] + within
    mb = refc[np.where(is_m, rp, 0)]
    err = rng.random(tot) < 0.03
    mb = np.where(err, (mb + rng.integers(1, 4, tot)) & 3, mb).astype(np.uint8)
    base = np.where(is_m, mb, base)
    qual_b = np.clip(np.round(np.exp(rng.normal(2.85, 0.55, tot))), 2, 90).astype(np.uint8)
    # scatter to padded per-read layout
    read_of_base = np.repeat(np.arange(n, dtype=np.int64), lq)
    rstart = np.cumsum(lq) - lq
    dst = soff[:-1].astype(np.int64)[read_of_base] + (np.arange(tot, dtype=np.int64) - rstart[read_of_base])
    nq = int(soff[-1])
    nib = np.zeros(nq, dtype=np.uint8)
    nib[dst] = _NIB[base]
    qual = np.zeros(nq, dtype=np.uint8)
    qual[dst] = qual_b
    seq4 = ((nib[0::2] << 4) | nib[1::2]).astype(np.uint8)
    flag = np.where(strand, 16, 0).astype(np.uint16)
    return packing.finalize_batch(pos.astype(np.int32), flag, np.full(n, 60, np.uint8), coff, cigar, soff, seq4, qual,
                                  min_mapq, max_depth, acgt_only=np.ones(n, dtype=bool))
