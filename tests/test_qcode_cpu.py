"""2-bit quality codes (lvc_batch::qual_bits == 2, include/lvc.h) on the CPU: the integer transforms the generation-5
tiled kernel applies while it stages a quality-code batch (csrc/qcode.hpp, host + device functions, compiled here with
g++) against a base-by-base restatement, and the host packer lvc_pack_quality_codes through the product library."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import ROOT
from lvc_b200 import capi, packing

CSRC = os.path.join(ROOT, "covid-spings-variant-caller_b200", "csrc")

HARNESS = r"""
#include "qcode.hpp"
extern "C" {
void keys16(uint32_t s0, uint32_t s1, uint32_t w, uint32_t pcode, uint32_t* out) { lvc::qc_keys16(s0, s1, w, pcode, out[0], out[1]); }
uint32_t cold_flags(uint32_t w, uint32_t cold) { return lvc::qc_cold_flags(w, cold); }
uint32_t eq_flags(uint32_t w, uint32_t c) { return lvc::qc_eq_flags(w, c); }
uint32_t spread8(uint32_t v) { return lvc::qc_spread8(v); }
void onehot16(uint32_t w, uint32_t* out) { lvc::b2_onehot16(w, out[0], out[1]); }
}
"""


@pytest.fixture(scope="module")
def qc(tmp_path_factory):
    d = tmp_path_factory.mktemp("qcode")
    src, so = os.path.join(d, "h.cpp"), os.path.join(d, "h.so")
    with open(src, "w") as fh:
        fh.write(HARNESS)
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread", "-I", CSRC, "-o", so, src], check=True)
    lib = C.CDLL(so)
    for f in ("cold_flags", "eq_flags", "spread8"):
        getattr(lib, f).restype = C.c_uint32
        getattr(lib, f).argtypes = [C.c_uint32, C.c_uint32] if f != "spread8" else [C.c_uint32]
    lib.keys16.argtypes = [C.c_uint32] * 4 + [C.POINTER(C.c_uint32)]
    lib.onehot16.argtypes = [C.c_uint32, C.POINTER(C.c_uint32)]
    return lib


def test_key_transform_matches_base_by_base(qc):
    rng = np.random.default_rng(7)
    for it in range(3000):
        nib = rng.integers(0, 16, 16)
        code = rng.integers(0, 4, 16) if it % 5 else np.full(16, it % 4)
        pcode = int(rng.integers(0, 4))
        s0 = sum(int(nib[k]) << (4 * k) for k in range(8))
        s1 = sum(int(nib[8 + k]) << (4 * k) for k in range(8))
        w = sum(int(code[k]) << (2 * k) for k in range(16))
        out = (C.c_uint32 * 2)()
        qc.keys16(s0, s1, w, pcode, out)
        want = [int(nib[k]) if code[k] == pcode else 0 for k in range(16)]
        got = [(out[k // 8] >> (4 * (k % 8))) & 15 for k in range(16)]
        assert got == want
        for c in range(4):
            assert qc.eq_flags(w, c) == sum(1 << (2 * k) for k in range(16) if code[k] == c)
        cold = int(rng.integers(0, 16))
        assert qc.cold_flags(w, cold) == sum(1 << (2 * k) for k in range(16) if (cold >> int(code[k])) & 1)
    for v in range(0, 1 << 16, 37):
        flags = v & 0x5555
        assert qc.spread8(v & 0x5555 | (v << 16)) == sum(15 << (4 * k) for k in range(8) if (flags >> (2 * k)) & 1)


def _decode(codes, d, n):
    c = np.asarray(codes[: (n + 3) // 4])
    four = np.stack([(c >> s) & 3 for s in (0, 2, 4, 6)], axis=1).ravel()[:n]
    return np.frombuffer(d, dtype=np.uint8)[four]


def test_pack_quality_codes_roundtrip_and_refusal():
    rng = np.random.default_rng(11)
    for n in (0, 2, 6, 150, 4098, (1 << 21) + 2):          # the last one runs the threaded path
        q = np.array([2, 12, 23, 37], dtype=np.uint8)[rng.integers(0, 4, n)]
        got = capi.pack_quality_codes(q, n)
        assert got is not None
        codes, d = got
        assert list(d) == sorted(set(q.tolist())) + [0] * (4 - len(set(q.tolist()))) or n == 0
        assert np.array_equal(_decode(codes, d, n), q)
    q = rng.integers(0, 5, 1000).astype(np.uint8)                          # five distinct values: byte form stays
    assert capi.pack_quality_codes(q, 1000) is None
    # only the qualities of ADMITTED reads decide: a dropped read may carry anything
    # ... and so does the pad byte of an odd-length read (here 99 after the 9 bases of the third read)
    so = np.array([0, 10, 20, 30], dtype=np.uint64)
    q = np.concatenate([np.full(10, 37), np.arange(10) + 50, np.full(9, 12), [99]]).astype(np.uint8)
    assert capi.pack_quality_codes(q, 30) is None
    coff, cig = np.array([0, 1, 2, 4], dtype=np.uint32), np.array([10 << 4, 10 << 4, (4 << 4) | 4, 5 << 4], dtype=np.uint32)
    codes, d = capi.pack_quality_codes(q, 30, np.array([1, 0, 3], dtype=np.uint8), so, coff, cig)
    assert list(d) == [12, 37, 0, 0]
    dec = _decode(codes, d, 30)
    assert np.array_equal(dec[:10], q[:10]) and np.array_equal(dec[20:29], q[20:29])


def test_readbatch_code_form_is_optional_and_lossless():
    rows = [(99, 5, 60, [(0, 20)], "ACGTN" * 4, [37] * 10 + [12] * 10), (147, 9, 60, [(0, 7), (2, 2), (0, 8)], "ACGTACGTACGTACG", [23] * 15)]
    b = packing.pack_reads(rows, 20)
    c = b.with_quality_codes()
    assert c is not b and c.qcode is not None and list(c.qdict) == [12, 23, 37, 0]
    assert np.array_equal(_decode(c.qcode, c.qdict, b.n_qual)[:20], b.qual[:20])
    cb = c.as_capi()
    assert cb.qual_bits == 2 and list(cb.qual_dict) == [12, 23, 37, 0] and b.as_capi().qual_bits == 0
    assert c.without_quality_codes().qcode is None
    wide = packing.pack_reads([(0, 1, 60, [(0, 8)], "ACGTACGT", list(range(8)))], 20)
    assert wide.with_quality_codes() is wide


def test_admitted_only_keeps_exactly_the_admitted_reads():
    from helpers import synth_small
    ref, reads = synth_small.make_scenario(seed=301, ref_len=120, n_reads=700, len_lo=90, len_hi=101, q_lo=0, q_hi=0,
                                           indel_rate=0.05, weird=True, fixed_pos=3, qbins=(2, 12, 23, 37))
    b = packing.pack_reads([(r.flag, r.pos, r.mapq, r.cigar, r.seq, r.qual) for r in reads], 20, 300)
    c = b.admitted_only()
    idx = np.nonzero(b.keep & 1)[0]
    assert 0 < c.n_reads == len(idx) < b.n_reads and c.admitted_only() is c
    for j, i in enumerate(idx):
        assert (c.pos[j], c.flag[j], c.mapq[j], c.keep[j]) == (b.pos[i], b.flag[i], b.mapq[i], b.keep[i])
        assert np.array_equal(c.cigar[c.cigar_off[j]:c.cigar_off[j + 1]], b.cigar[b.cigar_off[i]:b.cigar_off[i + 1]])
        s0, s1, t0, t1 = int(c.seq_off[j]), int(c.seq_off[j + 1]), int(b.seq_off[i]), int(b.seq_off[i + 1])
        assert np.array_equal(c.qual[s0:s1], b.qual[t0:t1]) and np.array_equal(c.seq4[s0 // 2:s1 // 2], b.seq4[t0 // 2:t1 // 2])
    assert c.n_cigar == int(c.cigar_off[-1]) and c.n_qual == int(c.seq_off[-1]) and len(c.qual) >= c.n_qual + 64
    coded = b.with_quality_codes().admitted_only()
    assert coded.qcode is not None and coded.n_reads == c.n_reads


def test_native_ingest_compact_and_codes(tmp_path):
    """lvc_reads_compact == ReadBatch.admitted_only() of the same file; the code form follows the re-packed layout"""
    from helpers import synth_small
    from lvc_b200 import samio
    ref, reads = synth_small.make_scenario(seed=302, ref_len=300, n_reads=1500, len_lo=60, len_hi=151, q_lo=0, q_hi=0,
                                           indel_rate=0.05, weird=True, amplicon=(0, 0, 40, 120), qbins=(2, 12, 23, 37))
    bam = str(tmp_path / "c.bam")
    samio.write_bam(bam, [("chrS", len(ref))], [(r.flag, r.pos, r.mapq, r.cigar, r.seq, r.qual, r.name) for r in reads])
    nat = samio.read_alignments_native(bam, None, 20, max_depth=200)
    full = nat.as_readbatch()
    want = packing.ReadBatch(*[np.array(getattr(full, k)) for k in ("pos", "flag", "mapq", "keep", "cigar_off", "cigar",
                                                                   "seq_off", "seq4", "qual")]).admitted_only()
    assert nat.batch.qual_bits == 2 and nat.n_presented == len(reads)
    assert nat.compact() and not nat.compact()
    got = nat.as_readbatch()
    assert got.n_reads == want.n_reads < len(reads) and nat.batch.n_reads == want.n_reads
    for k in ("pos", "flag", "mapq", "keep"):
        assert np.array_equal(getattr(got, k)[:got.n_reads], getattr(want, k)[:want.n_reads]), k
    assert np.array_equal(got.cigar_off, want.cigar_off) and np.array_equal(got.seq_off, want.seq_off)
    assert np.array_equal(got.cigar[:got.n_cigar], want.cigar[:want.n_cigar])
    assert np.array_equal(got.qual[:got.n_qual], want.qual[:want.n_qual])
    assert np.array_equal(got.seq4[:got.n_qual // 2], want.seq4[:want.n_qual // 2])
    assert nat.batch.qual_bits == 2 and got.qcode is not None
    wc = want.with_quality_codes()
    assert bytes(got.qdict) == wc.qdict and np.array_equal(got.qcode[:(got.n_qual + 3) // 4], wc.qcode[:(got.n_qual + 3) // 4])
    nat.close()


def test_base_code_transform_matches_base_by_base(qc):
    """b2_onehot16: 16 two-bit base codes -> the one-hot BAM nibbles (A, C, G, T = 1, 2, 4, 8) the tiled kernel stages"""
    rng = np.random.default_rng(11)
    words = [0, 0xFFFFFFFF, 0x55555555, 0xAAAAAAAA, 0x1B1B1B1B] + [int(x) for x in rng.integers(0, 1 << 32, 4000, dtype=np.uint64)]
    for w in words:
        out = (C.c_uint32 * 2)()
        qc.onehot16(w, out)
        got = [(out[k // 8] >> (4 * (k % 8))) & 15 for k in range(16)]
        assert got == [1 << ((w >> (2 * k)) & 3) for k in range(16)], hex(w)


def _base_codes_restated(b, min_bq):
    """what lvc_pack_base_codes must produce, base by base; None if a base that can reach the tables is not A/C/G/T"""
    n_qual = b.n_qual
    codes = np.zeros((n_qual + 3) // 4, dtype=np.uint8)
    lq = packing.query_lengths(b.cigar_off, b.cigar)
    code_of = {1: 0, 2: 1, 4: 2, 8: 3}
    for i in range(b.n_reads):
        x0, x1 = int(b.seq_off[i]), int(b.seq_off[i + 1])
        live = bool(b.keep[i] & 1)
        for x in range(x0, x1):
            byte = int(b.seq4[x >> 1])
            nib = (byte & 15) if (x & 1) else (byte >> 4)
            if nib not in code_of:
                if live and x < x0 + int(lq[i]) and int(b.qual[x]) >= min_bq:
                    return None
                c = 0
            else:
                c = code_of[nib]
            codes[x >> 2] |= c << (2 * (x & 3))
    return codes


@pytest.mark.parametrize("min_bq", [0, 13, 30])
def test_base_code_packer_matches_restatement(lib, min_bq):
    """lvc_pack_base_codes (every thread count) against the base-by-base rule, on reads with odd lengths, ambiguity codes
    below and above the threshold, dropped reads"""
    rng = np.random.default_rng(100 + min_bq)
    rows = []
    pos = 0
    for i in range(3000):
        l = int(rng.integers(1, 60))
        seq = "".join(rng.choice(list("ACGT"), size=l))
        qual = [int(q) for q in rng.choice([2, 12, 23, 37], size=l)]
        if rng.random() < 0.2:                                  # an N with the lowest quality: allowed iff it cannot pass
            k = int(rng.integers(0, l))
            seq = seq[:k] + "N" + seq[k + 1:]
            qual[k] = 2
        flag = 1024 if rng.random() < 0.1 else 0               # a duplicate: dropped by the admission, may hold anything
        if flag and rng.random() < 0.5:
            seq = "R" * l
        rows.append((flag, pos, 60, [(0, l)], seq, qual))
        pos += int(rng.integers(0, 3))
    b = packing.pack_reads(rows, 20)
    want = _base_codes_restated(b, min_bq)
    assert (want is None) == (min_bq <= 2)                      # an admitted N at quality 2 passes a threshold of 0 ... 2
    for nt in (1, 2, 5):
        got = capi.pack_base_codes(b.seq4, b.qual, b.n_qual, b.keep, b.seq_off, b.cigar_off, b.cigar, min_bq, nt)
        if want is None:
            assert got is None
        else:
            assert got is not None and np.array_equal(got[:len(want)], want), nt
    # an admitted ambiguity code with a passing quality: the batch keeps its nibbles
    rows.append((0, pos, 60, [(0, 4)], "ACRT", [37, 37, 37, 37]))
    b2 = packing.pack_reads(rows, 20)
    assert capi.pack_base_codes(b2.seq4, b2.qual, b2.n_qual, b2.keep, b2.seq_off, b2.cigar_off, b2.cigar, 30, 3) is None
    wb = packing.pack_reads(rows[:-1], 20).with_quality_codes().with_base_codes(30)
    assert (wb.scode is not None) == (min_bq >= 0) and wb.as_capi().seq_form == (2 | (30 << 8))
