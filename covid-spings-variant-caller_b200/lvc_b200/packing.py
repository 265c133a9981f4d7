"""Host-side packing of alignments into the structure-of-arrays batch of include/lvc.h.

Replaces what pysam hands the reference inside process_bam (live_variant_caller.py:55-60): instead of
a Python PileupColumn iterator, reads are packed once into flat arrays and shipped to the device.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import capi

NIBBLE_CHARS = "=ACMGRSVTWYHKDBN"
_ASCII_TO_NIBBLE = np.full(256, 15, dtype=np.uint8)       # unknown letters -> N (htslib seq_nt16_table)
for _i, _c in enumerate(NIBBLE_CHARS):
    _ASCII_TO_NIBBLE[ord(_c)] = _i
    _ASCII_TO_NIBBLE[ord(_c.lower())] = _i
_IS_ACGT_NIBBLE = np.zeros(16, dtype=bool)
_IS_ACGT_NIBBLE[[1, 2, 4, 8]] = True
CIGAR_OPS = "MIDNSHP=XB"
_QUERY_OP = np.zeros(16, dtype=bool)
_QUERY_OP[[0, 1, 4, 7, 8]] = True
_REF_OP = np.zeros(16, dtype=bool)
_REF_OP[[0, 2, 3, 7, 8]] = True


class UnsupportedInput(ValueError):
    """Input the path cannot reproduce bit-exactly yet; raised instead of silently diverging."""


@dataclass
class ReadBatch:
    """One coordinate-sorted batch of alignments of one contig, packed (numpy, C-contiguous)."""
    pos: np.ndarray        # int32  [n]
    flag: np.ndarray       # uint16 [n]
    mapq: np.ndarray       # uint8  [n]
    keep: np.ndarray       # uint8  [n]   bit0 admitted (lvc_admit), bit1 ACGT-only hint
    cigar_off: np.ndarray  # uint32 [n+1]
    cigar: np.ndarray      # uint32 [n_cigar]
    seq_off: np.ndarray    # uint64 [n+1]  even byte offsets into qual; seq4 offset = seq_off/2
    seq4: np.ndarray       # uint8  [seq_off[n]/2 (+pad)]
    qual: np.ndarray       # uint8  [seq_off[n] (+pad)]
    # optional 2-bit quality-code form (include/lvc.h, qual_bits == 2): when set, as_capi() ships the codes instead of
    # the phred bytes; `qual` stays for the host side (oracle, BAM writer, tests)
    qcode: Optional[np.ndarray] = None   # uint8 [seq_off[n]/4 (+pad)]
    qdict: Optional[bytes] = None        # 4 phred values
    # optional 2-bit base codes (lvc_batch seq_form 2; only beside quality codes): when set, as_capi() ships them instead
    # of the nibbles in `seq4`, which stay for the host side
    scode: Optional[np.ndarray] = None   # uint8 [seq_off[n]/4 (+pad)]
    scode_min_bq: int = 0                # the base-quality threshold they were made for

    @property
    def n_reads(self) -> int:
        return int(len(self.pos))

    @property
    def n_cigar(self) -> int:
        return int(self.cigar_off[-1]) if len(self.cigar_off) else 0

    @property
    def n_qual(self) -> int:
        return int(self.seq_off[-1]) if len(self.seq_off) else 0

    def aligned_bases(self) -> int:
        """metric unit: query bases consumed by M/=/X ops of every read PRESENTED."""
        ops = self.cigar[: self.n_cigar] & 15
        lens = self.cigar[: self.n_cigar] >> 4
        m = (ops == 0) | (ops == 7) | (ops == 8)
        return int(lens[m].sum(dtype=np.int64))

    def algorithmic_bytes(self, ref_len: int) -> int:
        """SURVEY 8d formula: sum_reads[20 + 4*n_cigar + ceil(l/2) + l] + 52*G."""
        lq = query_lengths(self.cigar_off, self.cigar)
        return int(20 * self.n_reads + 4 * self.n_cigar + ((lq + 1) // 2).sum() + lq.sum() + 52 * ref_len)

    def as_capi(self) -> capi.Batch:
        if self.qcode is not None and self.scode is not None:
            b = capi.Handle.make_batch(self.n_reads, self.n_cigar, self.n_qual, self.pos, self.flag, self.mapq,
                                       self.keep, self.cigar_off, self.cigar, self.seq_off, self.scode, self.qcode,
                                       qual_dict=self.qdict, base_codes_min_bq=self.scode_min_bq)
        elif self.qcode is not None:
            b = capi.Handle.make_batch(self.n_reads, self.n_cigar, self.n_qual, self.pos, self.flag, self.mapq,
                                       self.keep, self.cigar_off, self.cigar, self.seq_off, self.seq4, self.qcode,
                                       qual_dict=self.qdict)
        else:
            b = capi.Handle.make_batch(self.n_reads, self.n_cigar, self.n_qual, self.pos, self.flag, self.mapq,
                                       self.keep, self.cigar_off, self.cigar, self.seq_off, self.seq4, self.qual)
        b._keepalive = self         # the struct only carries raw pointers: keep the arrays alive with it
        return b

    def with_quality_codes(self, n_threads: int = 0) -> "ReadBatch":
        """the same batch carrying 2-bit quality codes (lvc_pack_quality_codes) if the qualities of its admitted reads
        take at most four distinct values; otherwise the batch itself.  Arrays are shared, nothing is copied."""
        if self.qcode is not None or self.n_reads == 0 or self.n_qual == 0:
            return self
        got = capi.pack_quality_codes(self.qual, self.n_qual, self.keep, self.seq_off, self.cigar_off, self.cigar, n_threads)
        if got is None:
            return self
        out = ReadBatch(self.pos, self.flag, self.mapq, self.keep, self.cigar_off, self.cigar, self.seq_off, self.seq4,
                        self.qual, got[0], got[1])
        for extra in ("overlap_pairs", "overlap_bases", "_keepalive", "_pins"):
            if hasattr(self, extra):
                setattr(out, extra, getattr(self, extra))
        return out

    def with_base_codes(self, min_base_quality: int, n_threads: int = 0) -> "ReadBatch":
        """the same quality-code batch carrying 2-bit base codes as well (lvc_pack_base_codes) if every base that can reach
        the tables of a handle with this base-quality threshold is A, C, G or T; otherwise (or without quality codes) the
        batch itself.  Arrays are shared."""
        if self.qcode is None or self.scode is not None or self.n_reads == 0 or self.n_qual == 0:
            return self
        codes = capi.pack_base_codes(self.seq4, self.qual, self.n_qual, self.keep, self.seq_off, self.cigar_off, self.cigar,
                                     min_base_quality, n_threads)
        if codes is None:
            return self
        out = ReadBatch(self.pos, self.flag, self.mapq, self.keep, self.cigar_off, self.cigar, self.seq_off, self.seq4,
                        self.qual, self.qcode, self.qdict, codes, int(min_base_quality))
        for extra in ("overlap_pairs", "overlap_bases", "_keepalive", "_pins"):
            if hasattr(self, extra):
                setattr(out, extra, getattr(self, extra))
        return out

    def admitted_only(self) -> "ReadBatch":
        """the same batch without the reads the host admission dropped (keep bit0 clear: htslib's max_depth rule and
        the read-level filter): no kernel ever reads them, so a batch that leaves them out deposits the same counts,
        and first-seen ordinals number the admitted reads in the same order (record order is unchanged).  What crosses
        PCIe shrinks by their per-read arrays (config 2: 60 % of the reads).  Returns self if nothing is dropped."""
        n = self.n_reads
        live = (self.keep[:n] & 1) != 0
        if n == 0 or bool(live.all()):
            return self
        idx = np.nonzero(live)[0]
        co = self.cigar_off.astype(np.int64)
        so = self.seq_off.astype(np.int64)
        nc = (co[1:] - co[:-1])[idx]
        nq = (so[1:] - so[:-1])[idx]                       # even (pad included)
        new_co = np.concatenate([[0], np.cumsum(nc)])
        new_so = np.concatenate([[0], np.cumsum(nq)])

        def gather(starts, lens, new_off, src, scale=1):
            # concatenation of src[starts[i] : starts[i] + lens[i]] (in units of `scale` source items)
            total = int(new_off[-1]) // scale
            if total == 0:
                return np.zeros(0, dtype=src.dtype)
            rep = np.repeat(np.arange(len(lens)), lens // scale)
            within = np.arange(total) - np.repeat(new_off[:-1] // scale, lens // scale)
            return src[(starts // scale)[rep] + within]
        cigar = gather(co[:-1][idx], nc, new_co, self.cigar)
        qual = gather(so[:-1][idx], nq, new_so, self.qual)
        seq4 = gather(so[:-1][idx], nq, new_so, self.seq4, 2)
        out = ReadBatch(self.pos[:n][idx].copy(), self.flag[:n][idx].copy(), self.mapq[:n][idx].copy(),
                        self.keep[:n][idx].copy(), new_co.astype(np.uint32),
                        cigar if len(cigar) else np.zeros(1, np.uint32), new_so.astype(np.uint64),
                        _with_slack(seq4, int(new_so[-1]) // 2), _with_slack(qual, int(new_so[-1])))
        for extra in ("overlap_pairs", "overlap_bases"):
            if hasattr(self, extra):
                setattr(out, extra, getattr(self, extra))
        return out.with_quality_codes() if self.qcode is not None else out

    def without_quality_codes(self) -> "ReadBatch":
        if self.qcode is None:
            return self
        out = ReadBatch(self.pos, self.flag, self.mapq, self.keep, self.cigar_off, self.cigar, self.seq_off, self.seq4,
                        self.qual)
        for extra in ("overlap_pairs", "overlap_bases", "_keepalive", "_pins"):
            if hasattr(self, extra):
                setattr(out, extra, getattr(self, extra))
        return out

    def slice(self, a: int, b: int) -> "ReadBatch":
        """reads [a, b) as a new batch sharing the payload arrays (offsets rebased)."""
        co = self.cigar_off[a:b + 1]
        so = self.seq_off[a:b + 1]
        return ReadBatch(self.pos[a:b].copy(), self.flag[a:b].copy(), self.mapq[a:b].copy(), self.keep[a:b].copy(),
                         (co - co[0]).astype(np.uint32), self.cigar[int(co[0]):int(co[-1])].copy(),
                         (so - so[0]).astype(np.uint64),
                         np.concatenate([self.seq4[int(so[0]) // 2:(int(so[-1]) + 1) // 2], np.zeros(64, np.uint8)]),
                         np.concatenate([self.qual[int(so[0]):int(so[-1])], np.zeros(64, np.uint8)]))


def finalize_batch(pos, flag, mapq, cigar_off, cigar, seq_off, seq4, qual, min_mapq: int,
                   max_depth: int = capi.MAX_DEPTH_DEFAULT, acgt_only: Optional[np.ndarray] = None,
                   mates: Optional[dict] = None, overlap_model: int = capi.OVERLAP_DEFAULT) -> ReadBatch:
    """Validate, compute the keep mask (host admission, SURVEY B2+B4) and the ACGT-only hint.
    `mates` = {"names": [...], "mate_pos": [...], "mate_ref": [...], "tlen": [...]} switches htslib's mate-overlap
    quality rewrite on (pysam's default); without it, or with overlap_model OFF, qualities stay as given."""
    pos = np.ascontiguousarray(pos, dtype=np.int32)
    flag = np.ascontiguousarray(flag, dtype=np.uint16)
    mapq = np.ascontiguousarray(mapq, dtype=np.uint8)
    cigar_off = np.ascontiguousarray(cigar_off, dtype=np.uint32)
    cigar = np.ascontiguousarray(cigar, dtype=np.uint32)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
    n = len(pos)
    nq = int(seq_off[-1]) if n else 0
    # payload arrays carry 64 bytes of slack: the tiled kernel stages whole 16-byte groups
    seq4 = _with_slack(np.ascontiguousarray(seq4, dtype=np.uint8), (nq + 1) // 2)
    qual = _with_slack(np.ascontiguousarray(qual, dtype=np.uint8), nq)
    if n and (seq_off & np.uint64(1)).any():
        raise ValueError("seq_off entries must be even")
    cig_store = cigar if len(cigar) else np.zeros(1, dtype=np.uint32)
    overlap_stats = (0, 0)
    if mates is not None and overlap_model != capi.OVERLAP_OFF and n:
        qual = np.array(qual, dtype=np.uint8, copy=True)              # rewritten in place
        keep, pairs, nb = capi.admit_overlaps(pos, flag, mapq, cigar_off, cig_store, seq_off, seq4, qual, mates["names"],
                                              mates["mate_pos"], mates["mate_ref"], mates["tlen"], min_mapq, max_depth,
                                              overlap_model)
        overlap_stats = (pairs, nb)
    else:
        keep = capi.admit(pos, flag, mapq, cigar_off, cig_store, min_mapq, max_depth)
    if acgt_only is None:
        acgt_only = acgt_only_hint(seq4, seq_off, n, query_lengths(cigar_off, cig_store))
    keep = (keep | (acgt_only.astype(np.uint8) << 1)).astype(np.uint8)
    out = ReadBatch(pos, flag, mapq, keep, cigar_off, cig_store, seq_off, seq4, qual)
    out.overlap_pairs, out.overlap_bases = overlap_stats
    return out


def _with_slack(a: np.ndarray, used: int, slack: int = 64) -> np.ndarray:
    if len(a) >= used + slack:
        return a
    out = np.zeros(used + slack, dtype=np.uint8)
    out[:used] = a[:used]
    return out


def query_lengths(cigar_off: np.ndarray, cigar: np.ndarray) -> np.ndarray:
    """l_qseq per read = sum of the query-consuming CIGAR op lengths (BAM invariant)."""
    n = len(cigar_off) - 1
    if n <= 0:
        return np.zeros(0, dtype=np.int64)
    nc = int(cigar_off[-1])
    ops = cigar[:nc] & 15
    w = np.where(_QUERY_OP[ops], (cigar[:nc] >> 4).astype(np.int64), 0)
    csum = np.concatenate([[0], np.cumsum(w)])
    co = cigar_off.astype(np.int64)
    return csum[co[1:]] - csum[co[:-1]]


def acgt_only_hint(seq4: np.ndarray, seq_off: np.ndarray, n: int, lq: Optional[np.ndarray] = None) -> np.ndarray:
    """bit1 of keep: True iff every base nibble of the read is A, C, G or T (the pad nibble of an
    odd-length read is ignored)."""
    if n == 0:
        return np.zeros(0, dtype=bool)
    nb = int(seq_off[-1]) // 2
    by = seq4[:nb]
    bad_hi = ~_IS_ACGT_NIBBLE[by >> 4]
    bad_lo = ~_IS_ACGT_NIBBLE[by & 15]
    starts = (seq_off[:-1] // np.uint64(2)).astype(np.int64)
    ends = (seq_off[1:] // np.uint64(2)).astype(np.int64)
    if lq is not None:
        odd = (lq & 1) == 1
        bad_lo[starts[odd] + lq[odd] // 2] = False
    bad = (bad_hi | bad_lo).astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(bad)])
    return (csum[ends] - csum[starts]) == 0


def pack_reads(reads: Iterable[Tuple], min_mapq: int, max_depth: int = capi.MAX_DEPTH_DEFAULT,
               overlap_model: int = capi.OVERLAP_DEFAULT) -> ReadBatch:
    """reads: iterable of (flag, pos0, mapq, [(op, len)...], seq_str, qual_ints[, name, mate_pos0, mate_ref, tlen]) in
    coordinate order; the four optional mate fields (mate_ref: 1 same contig, 0 other, -1 absent) enable the
    mate-overlap handling."""
    pos, flag, mapq, coff, cig, soff = [], [], [], [0], [], [0]
    seq_parts: List[np.ndarray] = []
    qual_parts: List[np.ndarray] = []
    names, mpos, mref, tlen = [], [], [], []
    for rec in reads:
        f, p, m, ops, seq, qual = rec[:6]
        if len(rec) >= 10:
            names.append(rec[6]); mpos.append(rec[7]); mref.append(rec[8]); tlen.append(rec[9])
        ops = [(o, l) for o, l in ops if l > 0]                  # zero-length ops carry no information
        lq = sum(l for o, l in ops if o in (0, 1, 4, 7, 8))
        if len(seq) != lq or len(qual) != lq:
            if any(o in (0, 2, 3, 7, 8) for o, _ in ops):
                raise UnsupportedInput(f"read at {p}: CIGAR query length {lq} != SEQ/QUAL length "
                                       f"{len(seq)}/{len(qual)} (the reference needs both to be present)")
            ops, seq, qual, lq = [], "", [], 0                    # never reaches the pileup: keep the slot only
        pos.append(p); flag.append(f); mapq.append(m)
        for o, l in ops:
            cig.append((l << 4) | o)
        coff.append(len(cig))
        nib = _ASCII_TO_NIBBLE[np.frombuffer(seq.encode("ascii"), dtype=np.uint8)] if lq else np.zeros(0, np.uint8)
        q = np.asarray(qual, dtype=np.uint8)
        if lq & 1:
            nib = np.concatenate([nib, np.zeros(1, np.uint8)])
            q = np.concatenate([q, np.zeros(1, np.uint8)])
        seq_parts.append(((nib[0::2] << 4) | nib[1::2]).astype(np.uint8))
        qual_parts.append(q)
        soff.append(soff[-1] + len(q))
    seq4 = np.concatenate(seq_parts) if seq_parts else np.zeros(0, np.uint8)
    qual = np.concatenate(qual_parts) if qual_parts else np.zeros(0, np.uint8)
    mates = dict(names=names, mate_pos=mpos, mate_ref=mref, tlen=tlen) if len(names) == len(pos) and names else None
    return finalize_batch(pos, flag, mapq, coff, cig, soff, seq4, qual, min_mapq, max_depth, mates=mates,
                          overlap_model=overlap_model)


class _Pinned:
    """one cudaHostAlloc'ed buffer exposed as a numpy array; freed with the object."""

    def __init__(self, a: np.ndarray):
        import ctypes as C
        self._lib = capi.load_library()
        nbytes = max(int(a.nbytes), 1)
        self.ptr = self._lib.lvc_host_alloc(nbytes)
        if not self.ptr:
            raise MemoryError("lvc_host_alloc (cudaHostAlloc) failed")
        buf = (C.c_uint8 * nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=a.dtype, count=a.size).reshape(a.shape)
        self.array[...] = a

    def __del__(self):
        try:
            if self.ptr:
                self._lib.lvc_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def pin_batch(b: ReadBatch) -> ReadBatch:
    """Copy a batch into page-locked host memory so lvc_push_batch's H2D copies run at full PCIe speed.  A batch that
    carries quality codes gets those pinned (they are what is shipped); its phred bytes stay where they are."""
    fields = ["pos", "flag", "mapq", "keep", "cigar_off", "cigar", "seq_off", "seq4"]
    if b.qcode is not None and b.scode is not None:
        # quality codes and base codes are what is shipped: the nibbles and the phred bytes stay where they are
        pins = [_Pinned(getattr(b, f)) for f in fields[:7]] + [_Pinned(b.qcode), _Pinned(b.scode)]
        out = ReadBatch(*[p.array for p in pins[:7]], b.seq4, b.qual, pins[7].array, b.qdict, pins[8].array, b.scode_min_bq)
        out._pins = pins
        return out
    pins = [_Pinned(getattr(b, f)) for f in fields]
    if b.qcode is not None:
        pins.append(_Pinned(b.qcode))
        out = ReadBatch(*[p.array for p in pins[:8]], b.qual, pins[8].array, b.qdict)
    else:
        pins.append(_Pinned(b.qual))
        out = ReadBatch(*[p.array for p in pins])
    out._pins = pins            # keep the allocations alive as long as the batch
    return out
